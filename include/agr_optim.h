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
#ifdef __cplusplus
}
#endif
#endif
