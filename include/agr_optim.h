/*
 * agr_optim.h — C ABI of the fused optimizer step over ONE flat parameter bucket.
 *
 * Replaces torch.optim.Adam(avatar_net.parameters(), lr) + optm.step() + optm.zero_grad()
 * (main_avatar.py:49-51,255-256) — a per-tensor loop over 650 tensors in the reference — by one
 * streaming pass over flat fp32 buffers (param, grad, exp_avg, exp_avg_sq); the gradient bucket is the same
 * buffer the single NCCL all-reduce of the view-sharded step works on.  Same update rule as torch Adam
 * (no amsgrad, weight_decay = 0):  m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
 *   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
 * grad_scale multiplies g first (1/world_size averaging or loss scaling); zero_grad != 0 clears g afterwards.
 */
#ifndef AGR_OPTIM_H_
#define AGR_OPTIM_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int agr_adam_step(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                  float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale,
                  int32_t zero_grad, void* cuda_stream);

/* Same update with the step counter living on the device (int32, incremented by the call itself), so the whole
 * training step including the optimizer can be captured once in a CUDA graph and replayed. */
int agr_adam_step_graph(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                        float lr, float beta1, float beta2, float eps, int32_t* device_step, float grad_scale,
                        int32_t zero_grad, void* cuda_stream);

/* Segmented form — what FlatAdam uses.  The bucket is a sequence of parameter SEGMENTS, each padded to a multiple of 256
 * elements; chunk_segment[c] names the segment of elements [256 c, 256 c + 256).  Per segment (device arrays of
 * num_segments entries): segment_active != 0 <=> the parameter received a gradient this step — inactive segments are
 * skipped entirely (parameter, moments and step untouched), which is torch.optim.Adam's behaviour for `p.grad is None`
 * (the trainer freezes position_net / colour nets per iteration, main_avatar.py:184-189, and pretraining never reaches
 * colour / viewdir nets); segment_step = per-parameter step counter (int32, incremented here for active segments);
 * segment_corr = 2 floats of scratch per segment (bias corrections).  hyper = 2 floats ON THE DEVICE: {lr, grad_scale},
 * read at execution time, so a CUDA graph captured around this call follows a learning-rate schedule
 * (main_avatar.py:61-68 update_lr) by refreshing those 8 bytes before each replay.  n must be a multiple of 256. */
int agr_adam_step_segments(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, int32_t num_segments,
                           const int32_t* chunk_segment, const int32_t* segment_active, int32_t* segment_step,
                           float* segment_corr, const float* hyper, float beta1, float beta2, float eps, int32_t zero_grad,
                           void* cuda_stream);
#ifdef __cplusplus
}
#endif
#endif
