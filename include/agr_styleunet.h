/*
 * agr_styleunet.h — C ABI of the StyleUNet operators (NHWC activations, fp32 or bf16).
 *
 * Replaces the reference's two native modules and the elementwise glue around its cuDNN convolutions:
 *   module `fused`     : fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 *                        (network/styleunet/fused_bias_act.cpp:18-31, fused_bias_act_kernel.cu:18-107)
 *   module `upfirdn2d` : upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
 *                        (network/styleunet/upfirdn2d.cpp:17-30, upfirdn2d_kernel.cu:49-369)
 *   ModulatedConv2d weight preparation (dual_styleunet.py:256-265), NoiseInjection (:303-313),
 *   HaarTransform / InverseHaarTransform (:387-425).
 * The dense contractions that replace the cuDNN calls (conv2d_gradfix.py:34,66) live in agr_conv.h.
 *
 * dtype: AGR_F32 = 0, AGR_BF16 = 1 (activations; reductions / parameters stay fp32).
 * Layout: activations are NHWC (torch channels_last), i.e. x[n][h][w][c] contiguous in c.
 * All pointers are device pointers owned by the caller except where marked HOST. Returns AgrStatus.
 */
#ifndef AGR_STYLEUNET_H_
#define AGR_STYLEUNET_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define AGR_F32 0
#define AGR_BF16 1

/* Per-channel FIR resampling: zero-insert upsample by `up`, pad (pad0 before / implied after), correlate with
 * `taps` (kh*kw HOST floats, ALREADY flipped, i.e. taps = flip(kernel)), decimate by `down`.
 * out[oy][ox][c] = sum_{ky,kx} taps[ky][kx] * U[oy*down + ky - pad_y0][ox*down + kx - pad_x0][c],
 * U[uy][ux] = x[uy/up][ux/up] when uy%up == ux%up == 0 and inside, else 0.   kh*kw <= 64. */
int agr_upfirdn2d(int32_t dtype, const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C,
                  int32_t out_h, int32_t out_w, const float* taps, int32_t kh, int32_t kw,
                  int32_t up, int32_t down, int32_t pad_x0, int32_t pad_y0, void* cuda_stream);

/* Haar analysis (C -> 4C at H/2 x W/2, sub-bands ll|lh|hl|hh) and synthesis (4C -> C at 2H x 2W), and their
 * adjoints (the backward of one is `transpose` of itself): dual_styleunet.py:387-425.
 * mode 0: dwt forward, 1: dwt backward (4C,H/2 -> C,H), 2: iwt forward, 3: iwt backward. H, W, C describe x. */
int agr_haar(int32_t dtype, int32_t mode, const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C,
             void* cuda_stream);

/* y = act(x + noise_w[0] * noise[h][w] + bias[c]),  act = identity (activate 0) | lrelu(.,0.2)*sqrt(2) (1) | lrelu(.,0.2) (2) | relu (3).
 * noise: (H,W) fp32 or NULL, indexed with pixel % noise_period (= H*W: one noise image shared by the batch);
 * noise_w: 1 fp32 on device or NULL; bias: (C) fp32 or NULL. */
int agr_bias_act_forward(int32_t dtype, const void* x, void* y, int64_t pixels, int32_t C, const float* bias,
                         const float* noise, const float* noise_w, int64_t noise_period, int32_t activate,
                         void* cuda_stream);
/* dx = dy * (activate ? (y > 0 ? gain : slope*gain) : 1), gain = sqrt2 (activate 1) or 1 (2, 3), slope = 0.2 (1, 2) or 0 (3); d_bias[c] += sum dx; d_noise_w[0] += sum dx*noise.
 * d_bias / d_noise_w (fp32) are ACCUMULATED into (caller zeroes); either may be NULL. y is the forward OUTPUT.
 * dx may be NULL when activate == 0 (dx == dy: only the reductions are computed). */
int agr_bias_act_backward(int32_t dtype, const void* dy, const void* y, void* dx, int64_t pixels, int32_t C,
                          const float* noise, int64_t noise_period, float* d_bias, float* d_noise_w, int32_t activate,
                          void* cuda_stream);

/* The reference's native op `fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)` element for element
 * (fused_bias_act.cpp:18-31, fused_bias_act_kernel.cu:18-107), for the drop-in module dropin/fused.py:
 *   v = x[i] + (bias ? bias[(i / step_b) % size_b] : 0);  r = refer ? refer[i] : 0
 *   act*10+grad: 10,11 -> v | 30 -> v > 0 ? v : v*alpha | 31 -> r > 0 ? v : v*alpha | 12,32 -> 0;   out[i] = y * scale
 * x, bias, refer, out share `dtype`; any contiguous layout (step_b = elements per bias index, size_b = bias length). */
int agr_fused_bias_act(int32_t dtype, const void* x, const void* bias, const void* refer, void* out, int64_t size_x,
                       int64_t step_b, int64_t size_b, int32_t act, int32_t grad, float alpha, float scale,
                       void* cuda_stream);

/* Modulated-convolution weight: w_out[co][ky][kx][ci] (KRSC, dtype) = scale*w[co][ci][ky][kx]*s[ci]*demod[co],
 * demod[co] = rsqrt(sum_{ci,k} (scale*w*s)^2 + 1e-8) if demodulate else 1.  w fp32 (Cout,Cin,k,k), s fp32 (Cin).
 * transpose_io != 0 writes w_out[ci][ky][kx][co] instead (operand of the transposed convolution).
 * demod_out (Cout) fp32 saved for the backward (may be NULL when demodulate == 0). */
int agr_modweight_forward(int32_t dtype, const float* w, const float* s, float scale, int32_t Cout, int32_t Cin,
                          int32_t k, int32_t demodulate, int32_t transpose_io, void* w_out, float* demod_out,
                          void* cuda_stream);
/* Backward: d_wout (same layout/dtype as w_out) -> d_w (Cout,Cin,k,k) fp32 (overwritten) and d_s (Cin) fp32
 * (ACCUMULATED; caller zeroes; NULL = not wanted). */
int agr_modweight_backward(int32_t dtype, const float* w, const float* s, float scale, int32_t Cout, int32_t Cin,
                           int32_t k, int32_t demodulate, int32_t transpose_io, const void* d_wout,
                           const float* demod, float* d_w, float* d_s, void* cuda_stream);

/* Grouped form of the two calls above: all conv layers of a U-Net (modulated and plain equalised ones, the latter with
 * s = ones and demodulate = 0) in ceil(count / 40) launches.  `items` is a HOST array; every pointer in it is a device
 * pointer with the meaning of the same-named argument above.  forward reads w, s and writes w_out, demod;
 * backward reads w, s, demod, d_wout and writes d_w, accumulates d_s (d_s may be NULL: no style gradient wanted). */
typedef struct AgrModWeightItem {
    const float* w;
    const float* s;
    void* w_out;
    float* demod;
    const void* d_wout;
    float* d_w;
    float* d_s;
    float scale;
    int32_t Cout, Cin, k, demodulate, transpose_io;
} AgrModWeightItem;
int agr_modweight_group_forward(int32_t dtype, const AgrModWeightItem* items, int32_t count, void* cuda_stream);
int agr_modweight_group_backward(int32_t dtype, const AgrModWeightItem* items, int32_t count, void* cuda_stream);

/* The dense contractions (every convolution of the path, forward / data gradient / weight gradient) are declared in
 * agr_conv.h. */

/* View-feature injection of the colour net (dual_styleunet.py:881-883,900-902):
 *   y[v] = base[v or 0] + bilinear_2x(vf[v])      F.interpolate(mode='bilinear', align_corners=False), exact 2x
 * vf (V,h,w,C) fp32 NHWC, base (Vb,2h,2w,C) dtype NHWC with Vb in {1,V}, y (V,2h,2w,C) dtype.  One pass instead of the
 * reference's resize -> add (two full-resolution fp32 intermediates per view).  The backward w.r.t. vf is the adjoint
 * resampling of g (V,2h,2w,C) dtype into d_vf (V,h,w,C) fp32, gather form (no atomics). */
int agr_bilinear2x_add_forward(int32_t dtype, const float* vf, const void* base, void* y, int32_t V, int32_t Vb, int32_t h,
                               int32_t w, int32_t C, void* cuda_stream);
int agr_bilinear2x_backward(int32_t dtype, const void* g, float* d_vf, int32_t V, int32_t h, int32_t w, int32_t C,
                            void* cuda_stream);

/* y[i] = sum_v x[v][i], i < n  (fp32 accumulate): the adjoint of broadcasting the shared colour-net prefix state to the
 * V views of a batch (ATen's strided reduction reaches ~0.3 TB/s on this shape; this streams at HBM rate). */
int agr_sum_batch(int32_t dtype, const void* x, void* y, int32_t V, int64_t n, void* cuda_stream);

/* ToRGB skip path  dwt(upsample(iwt(skip)))  (dual_styleunet.py:624-631; Upsample 30-50 = upfirdn2d(up=2, pad=(2,1)) with
 * a 4x4 FIR, HaarTransform / InverseHaarTransform 374-425) as ONE pass over NHWC tensors with 4*Ci channels (bands
 * ll|lh|hl|hh major, colour minor).  x (N,h,w,4Ci) -> y (N,2h,2w,4Ci).  The chain is linear, so the host collapses it
 * into a parity-dependent 2x2-tap filter bank (`taps256`, 256 floats):
 *   adjoint = 0:  y[2m+pi, 2n+pj, bo] = sum_{bi,di,dj} taps[pi][pj][bo][bi][di][dj] * x[m+di+pi-1, n+dj+pj-1, bi]
 *   adjoint = 1:  x is the gradient g (N,2h,2w,4Ci), y = d_skip (N,h,w,4Ci):
 *                 y[m, n, bi] = sum_{bo,a,b} taps[bi][bo][a][b] * g[2m-1+a, 2n-1+b, bo]
 * (zero outside the tensor; h, w are the SKIP size in both modes).  Replaces 3 launches and a (4h,4w,Ci)
 * intermediate each way. */
int agr_wavelet_upsample(int32_t dtype, int32_t adjoint, const void* x, void* y, int32_t N, int32_t h, int32_t w, int32_t Ci,
                         const float* taps256, void* cuda_stream);

/* EqualLinear on ONE style vector -- the modulation layer of every ModulatedConv2d (dual_styleunet.py:150-160, no
 * activation; called at dual_styleunet.py:244 with a (1, style_dim) input).  fp32:
 *   y[j] = lr_mul * bias[j] + scale * sum_i w[j][i] * x[i]                       w (out_dim, in_dim) row-major
 * backward: d_w[j][i] = scale * dy[j] * x[i];  d_bias[j] = lr_mul * dy[j];  d_x[i] += scale * sum_j w[j][i] * dy[j].
 * d_x must be zero-initialised by the caller (accumulated with atomics); d_bias / d_x / bias may be NULL.
 * One launch each instead of the reference's mul, mul, addmm (forward) and mm, mm, mul, mul, sum (autograd backward):
 * the three U-Nets hold 108 of these layers, ~900 tiny launches per step. */
int agr_equal_linear_forward(const float* w, const float* bias, const float* x, float scale, float lr_mul, int32_t out_dim,
                             int32_t in_dim, float* y, void* cuda_stream);
int agr_equal_linear_backward(const float* w, const float* x, const float* dy, float scale, float lr_mul, int32_t out_dim,
                              int32_t in_dim, float* d_w, float* d_bias, float* d_x, void* cuda_stream);

/* Grouped form: every modulation layer of a U-Net in ceil(count / 40) launches.  `items` is a HOST array of device
 * pointers with the meaning of the same-named arguments above; several items may share one d_x (their contributions
 * are accumulated). */
typedef struct AgrEqualLinearItem {
    const float* w;
    const float* bias;
    const float* x;
    const float* dy;
    float* y;
    float* d_w;
    float* d_bias;
    float* d_x;
    float scale, lr_mul;
    int32_t out_dim, in_dim;
} AgrEqualLinearItem;
int agr_equal_linear_group_forward(const AgrEqualLinearItem* items, int32_t count, void* cuda_stream);
int agr_equal_linear_group_backward(const AgrEqualLinearItem* items, int32_t count, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
