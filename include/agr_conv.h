/*
 * agr_conv.h — C ABI of the dense contractions of the StyleUNet path: every convolution / transposed convolution
 * (forward, data gradient, weight gradient) of network/styleunet/dual_styleunet.py and of AvatarNet.viewdir_net.
 *
 * Replaces the reference's cuDNN calls:
 *   F.conv2d / F.conv_transpose2d            network/styleunet/conv2d_gradfix.py:34,66 (called from
 *                                            dual_styleunet.py:114 EqualConv2d, :275-296 ModulatedConv2d's three branches)
 *   their autograd backward                  cudnn_convolution_backward_{input,weight} (ATen; conv2d_gradfix's custom
 *                                            functions are dead code on torch >= 2, conv2d_gradfix.py:85-92)
 *   nn.Conv2d(1,64,4,2,1), nn.Conv2d(64,128,4,2,1)   network/avatar.py:46-50 (viewdir_net)
 *
 * Layouts: activations NHWC (torch channels_last); weights "KRSC": w[co][ky][kx][ci] (torch channels_last of a
 * (Cout,Cin,k,k) tensor); the data-gradient calls take the transposed operand w_t[ci][ky][kx][co] (agr_weight_transpose).
 * dtype: AGR_F32 (0) or AGR_BF16 (1) for activations and weight operands; weight GRADIENTS are always fp32.
 *
 * Geometry (one struct for every layer of the path):
 *   transposed == 0:  y[n][oy][ox][co] = sum_{ky,kx,ci} x[n][oy*stride + ky - pad][ox*stride + kx - pad][ci] * w[co][ky][kx][ci]
 *   transposed == 1:  y[n][iy*stride + ky - pad][ix*stride + kx - pad][co] += x[n][iy][ix][ci] * w[co][ky][kx][ci]
 *   (x outside [0,H)x[0,W) reads as zero; y outside [0,OH)x[0,OW) is dropped).  The caller states OH, OW.
 *   StyleUNet layers:  3x3 "same" (stride 1, pad 1) | 1x1 | blur -> 3x3 stride 2 pad 0 (downsample, dual_styleunet.py:283-290,
 *   342-356) | transposed 3x3 stride 2 pad 0 -> blur (upsample, :266-282) | 4x4 stride 2 pad 1 (viewdir_net).
 *
 * Two implementations sit behind every call and are chosen by shape (agr_conv2d_path reports which):
 *   path 1  tcgen05: implicit GEMM on the 5th-gen tensor cores (bf16, Cin % 64 == 0, Cout % 64 == 0, stride <= 2):
 *           TMA boxes of the NHWC activation per filter tap (element strides for stride-2 layers, out-of-bounds zero
 *           fill = padding), tcgen05.mma with fp32 accumulators in TMEM, fused epilogue.  Transposed convolutions run
 *           as stride^2 output-phase sub-convolutions in one launch (no zero-insertion, no wasted MACs).  Weight
 *           gradients contract over pixels with MN-major operands straight from the NHWC tensors.
 *   path 2  direct: CUDA-core implicit GEMM (fp32 accumulate) for fp32 activations (the parity mode) and for the
 *           narrow layers the tensor-core tiles cannot express (3-channel inputs, 1-channel view map, 12/32-channel
 *           ToRGB outputs); these are HBM-bound layers.
 * There is no library (cuDNN / cuBLAS) path and no CPU fallback.
 */
#ifndef AGR_CONV_H_
#define AGR_CONV_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct AgrConvGeom {
    int32_t N, H, W, Cin;    /* input  x (N,H,W,Cin) */
    int32_t OH, OW, Cout;    /* output y (N,OH,OW,Cout) */
    int32_t ksize, stride, pad;
    int32_t transposed;
} AgrConvGeom;

/* Fused epilogue of the forward call:  y = act(acc + residual + noise_w*noise + bias).  Any pointer may be NULL. */
typedef struct AgrConvEpilogue {
    const float* bias;      /* (Cout) fp32 */
    const float* noise;     /* (OH*OW) fp32, one image shared by the N images (NoiseInjection, dual_styleunet.py:303-313) */
    const float* noise_w;   /* (1) fp32 */
    const float* residual;  /* (OH,OW,Cout) fp32 shared by the N images: the view-independent half of a split contraction
                               (path 1 only) */
    int32_t activate;       /* 0 none | 1 leaky-ReLU(0.2)*sqrt(2) (FusedLeakyReLU, fused_act.py:117-132) | 2 leaky-ReLU(0.2) |
                               3 ReLU (the VGG-16 trunk of LPIPS, network/lpips/pretrained_networks.py:96-134) */
    int32_t out_fp32;       /* != 0: store the raw fp32 accumulator in y (no bias / noise / activation; path 1 only) */
    int32_t w_cin_total;    /* weight holds w_cin_total >= Cin input channels per tap; 0 means Cin */
    int32_t w_cin_offset;   /* first input channel of the slice this call contracts with (multiple of 64 on path 1) */
} AgrConvEpilogue;

/* 1 = tcgen05, 2 = direct, 0 = invalid geometry.  `what`: 0 forward, 1 data gradient, 2 weight gradient. */
int agr_conv2d_path(int32_t dtype, const AgrConvGeom* g, int32_t what);

/* y = epilogue(conv(x, w_krsc)).  ep may be NULL (plain store). */
int agr_conv2d_forward(int32_t dtype, const AgrConvGeom* g, const void* x, const void* w_krsc, void* y,
                       const AgrConvEpilogue* ep, void* cuda_stream);
/* dx (N,H,W,Cin) = adjoint of the geometry applied to dy (N,OH,OW,Cout) with w_t[ci][ky][kx][co]. */
int agr_conv2d_dgrad(int32_t dtype, const AgrConvGeom* g, const void* dy, const void* w_t, void* dx, void* cuda_stream);
/* The same data gradient on path 1 straight from the layer's KRSC weight (no transposed copy: the tensor cores read it as
 * an MN-major operand).  The weight holds w_cin_total >= g->Cin input channels per tap (0 = g->Cin); dx receives the
 * g->Cin channels starting at w_cin_offset (multiple of 64).  AGR_ERR_INVALID_ARGUMENT when agr_conv2d_path(.., 1) != 1. */
int agr_conv2d_dgrad_krsc(int32_t dtype, const AgrConvGeom* g, const void* dy, const void* w_krsc, int32_t w_cin_total,
                          int32_t w_cin_offset, void* dx, void* cuda_stream);
/* dw[co][ky][kx][ci_offset + ci] (fp32, rows of ci_total floats) = sum over pixels of dy x x.  zero_first != 0 clears
 * the (Cout,k,k,ci_total) buffer before accumulating; partial sums are added with fp32 reductions (order not fixed). */
int agr_conv2d_wgrad(int32_t dtype, const AgrConvGeom* g, const void* x, const void* dy, float* dw, int32_t ci_total,
                     int32_t ci_offset, int32_t zero_first, void* cuda_stream);
/* w_out[ci][t][co] = w_krsc[co][t][ci]  (dtype elements). */
int agr_weight_transpose(int32_t dtype, const void* w_krsc, void* w_out, int32_t Cout, int32_t Cin, int32_t ksize,
                         void* cuda_stream);
/* Tuning / A-B switch of path 1's forward-form kernel (same results, different staging): 1 = one TMA box per tap
 * (conv_tc_kernel), 2 = tap groups sharing one haloed box, 3 = tap groups + CTA pairs (cta_group::2), 0 = chosen per layer
 * shape from measurements (the default; also the environment variable AGR_CONV_TC).  Returns the setting in force. */
int agr_conv2d_set_generation(int32_t generation);
/* Tuning of path 1's weight-gradient kernel, which splits the pixel range over CTAs (split-K with fp32 reductions into
 * dw): at most `max_ctas` CTAs, each contracting at least `min_boxes` 16x8-pixel boxes (more CTAs = more parallelism but
 * more reduction traffic per useful FLOP).  Values <= 0 leave a setting unchanged; returns max_ctas in force. */
int agr_conv2d_set_wgrad_split(int32_t max_ctas, int32_t min_boxes);

#ifdef __cplusplus
}
#endif
#endif
