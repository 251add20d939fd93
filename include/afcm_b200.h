/*
 * afcm_b200.h -- C ABI of libafcm_b200.so, the B200 (sm_100a) implementation of the AFCM generator
 * forward hot path.  Plain pointers and sizes only: no torch / ATen types cross this boundary.
 *
 * Each entry point replaces one native interface of the reference (paths relative to the
 * reference repo; OPS = models/networks/stylegan3/torch_utils/ops, NET =
 * models/networks/stylegan3/networks_stylegan3.py).  INTEGRATION.md shows the binding a reference
 * maintainer would add (ctypes, as used by afcm_b200/_lib.py).
 *
 * Conventions
 *   - every function returns int: 0 = ok, AFCM_ERR_UNSUPPORTED (-1) = no kernel for these arguments
 *     (same meaning as the reference's return code -1, OPS/filtered_lrelu.cpp:52-56: the caller may
 *     use the generic composition), AFCM_ERR_INVALID (-2) = bad arguments (the reference raises
 *     TORCH_CHECK), > 0 = a cudaError_t.  afcm_last_error() returns a thread-local message.
 *   - all data pointers are DEVICE pointers unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised.
 *   - tensors are NCHW; dtype codes AFCM_F32 / AFCM_F16 / AFCM_BF16.  Filter taps are HOST arrays:
 *     they are layer constants and travel in the kernel-parameter constant bank (the reference
 *     staged them through a global buffer + __constant__ copy on every call,
 *     OPS/filtered_lrelu.cu:87-117, which made the op unsafe on concurrent streams).
 *   - no entry point allocates device memory; the caller (PyTorch in this repo) owns every buffer.
 */
#ifndef AFCM_B200_H
#define AFCM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFCM_OK               0
#define AFCM_ERR_UNSUPPORTED (-1)
#define AFCM_ERR_INVALID     (-2)

#define AFCM_F32  0
#define AFCM_F16  1
#define AFCM_BF16 2

/* sign-tensor modes of afcm_filtered_lrelu (OPS/filtered_lrelu.cpp:83-96) */
#define AFCM_SIGN_NONE  0
#define AFCM_SIGN_WRITE 1
#define AFCM_SIGN_READ  2

/* ---------------------------------------------------------------------------------------------- */
/* library                                                                                         */

int         afcm_version(void);                 /* ABI version, currently 1 */
const char* afcm_last_error(void);              /* message of the last failing call on this thread */
int         afcm_device_check(void);            /* 0 iff the current device is sm_100 (B200)        */
long long   afcm_launch_count(void);            /* kernels launched by this library so far          */

/* ---------------------------------------------------------------------------------------------- */
/* filtered_lrelu -- replaces filtered_lrelu_plugin.filtered_lrelu (OPS/filtered_lrelu.cpp:16-209)   */
/*
 * y = down_fir(clamp(lrelu((up_fir(zero_insert(x + b)) ) * gain)))   (OPS/filtered_lrelu.py:59-80)
 *
 * x [N,C,xh,xw] with element strides xs[4] = {n,c,h,w};  y [N,C,yh,yw] with strides ys[4];
 * yh/yw must equal afcm_filtered_lrelu_out_size().  b: [C] or NULL.  skip: optional tensor with y's
 * shape and strides that is added to the result (fuses NET:376-377), out_scale multiplies it
 * (fuses NET:699-700); pass NULL / 1.0f for the plain op.
 * fu_host/fd_host: separable 1-D taps (NULL + taps 1 = identity).  Supported fused geometries:
 *   (up,fu,down,fd) in {(2,12,2,12), (2,12,4,24), (4,24,2,12)}  and  (1,1,1,1);
 * anything else returns AFCM_ERR_UNSUPPORTED and the caller composes afcm_upfirdn2d +
 * afcm_filtered_lrelu_act + afcm_upfirdn2d exactly like OPS/filtered_lrelu.py:223-229.
 * sign_mode WRITE: `signs` [N,C,sh,swb] uint8 is written (2 bits / up-sampled element: 1 = negative,
 * 2 = clamped; sh, swb from afcm_filtered_lrelu_sign_size()).  READ: `signs` is applied at element
 * offset (sx,sy) instead of lrelu/clamp (the backward pass, OPS/filtered_lrelu.py:252-266).
 */
int afcm_filtered_lrelu(const void* x, const int64_t* xs, void* y, const int64_t* ys,
                        const void* b, const void* skip, int dtype,
                        int N, int C, int xh, int xw, int yh, int yw,
                        const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                        int up, int down, int px0, int px1, int py0, int py1,
                        float gain, float slope, float clamp, float out_scale, int flip_filter,
                        int sign_mode, void* signs, int sign_h, int sign_wb, int sx, int sy,
                        void* stream);

int afcm_filtered_lrelu_out_size(int xh, int xw, int up, int down, int fu_taps, int fd_taps,
                                 int px0, int px1, int py0, int py1, int* yh, int* yw);
int afcm_filtered_lrelu_sign_size(int yh, int yw, int down, int fd_taps, int* sh, int* swb);

/* Tile override for tuning (0,0 = automatic).  Process-global, not part of the stable ABI. */
int afcm_filtered_lrelu_set_tile(int tow, int toh);

/* Tensor-core variant of afcm_filtered_lrelu (forward, no sign tensor): the four separable FIR passes run as
 * banded-Toeplitz products on mma.sync (fp16 operands, fp32 accumulation) chained in registers
 * (afcm_b200/csrc/flr_tc.cu).  Same semantics and argument meaning as afcm_filtered_lrelu; differences:
 * x and y may have different dtypes (x_dtype / y_dtype in {AFCM_F32, AFCM_F16}); `b` is always float32;
 * `skip` has y's dtype and strides; innermost strides must be 1.  Geometries: (up,down) in {(2,2),(4,2),(2,4)}
 * with 6*up / 6*down taps; anything else returns AFCM_ERR_UNSUPPORTED.  Results differ from the fp32 op by the
 * fp16 rounding of operands: max |err| <= 2e-3 * max|y| per call (tests/test_gpu_flr_tc.py). */
int afcm_filtered_lrelu_tc(const void* x, const int64_t* xs, int x_dtype, void* y, const int64_t* ys, int y_dtype,
                           const float* b, const void* skip,
                           int N, int C, int xh, int xw, int yh, int yw,
                           const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                           int up, int down, int px0, int px1, int py0, int py1,
                           float gain, float slope, float clamp, float out_scale, int flip_filter,
                           void* stream);
/* The same call writing, behind the last column of every output row, `zero_pad_cols` (0 or 2) zero samples: a result stored
 * at the row pitch yw + 2 (ys[2] >= yw + 2) is then the flat plane of row pitch W + 2 that afcm_conv2d_tc_nchw reads without
 * any row bookkeeping (x_pitch = W + 2). */
int afcm_filtered_lrelu_tc_padded(const void* x, const int64_t* xs, int x_dtype, void* y, const int64_t* ys, int y_dtype,
                           const float* b, const void* skip,
                           int N, int C, int xh, int xw, int yh, int yw,
                           const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                           int up, int down, int px0, int px1, int py0, int py1,
                           float gain, float slope, float clamp, float out_scale, int flip_filter,
                           int zero_pad_cols, void* stream);
/* The register-chained tensor-core kernel WITH the sign tensor (the generator training step): fp32 planes in and out, fp16
 * operands.  AFCM_SIGN_WRITE = forward (signs [N,C,sign_h,sign_wb] uint8 in the reference's format, OPS/filtered_lrelu.cpp:87-94,
 * bit-compatible with afcm_filtered_lrelu; needs a clamp in [2^-10, 2^10]); AFCM_SIGN_READ = backward (the op with up / down
 * exchanged, OPS/filtered_lrelu.py:252-263; sx, sy = sign offsets; the activation is the multiplier 1 / slope / 0 the stored code
 * selects, OPS/filtered_lrelu.cu:562-572).  amax: device pointer to max|x| written by afcm_absmax, or NULL -- the backward scales its
 * fp16 operands by the power of two that brings max|x| to ~256 (gradients lie far below the fp16 range) and scales the fp32
 * result back. */
int afcm_filtered_lrelu_tc_signs(const void* x, const int64_t* xs, void* y, const int64_t* ys, const float* b,
                                 int N, int C, int xh, int xw, int yh, int yw,
                                 const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                                 int up, int down, int px0, int px1, int py0, int py1,
                                 float gain, float slope, float clamp, int flip_filter,
                                 int sign_mode, void* signs, int sign_h, int sign_wb, int sx, int sy, const float* amax,
                                 void* stream);
/* out[0] = max|x| over n dense fp32 values (x 16-byte aligned), NaN / Inf ignored; stays on the device. */
int afcm_absmax(const float* x, int64_t n, float* out, void* stream);
/* Scheduling of afcm_filtered_lrelu_tc for tuning (process-global, not part of the stable ABI): n >= 1 = persistent warps,
 * at most n resident waves of CTAs (default 16, the measured optimum); 0 = one warp per 16-column strip. */
int afcm_filtered_lrelu_tc_set_waves(int waves);

/* filtered_lrelu on tcgen05 / tensor memory (afcm_b200/csrc/flr_t5.cu): the Blackwell-native inference kernel.  Replaces
 * the same reference entry point as afcm_filtered_lrelu (filtered_lrelu_plugin.filtered_lrelu, OPS/filtered_lrelu.cpp:16-209,
 * kernel OPS/filtered_lrelu.cu:139-1099) for the fast path: same arguments as afcm_filtered_lrelu_tc.  Input rows arrive by
 * TMA, every FIR pass is a banded-Toeplitz tcgen05.mma with the image tile resident in tensor memory.  Requirements (else
 * AFCM_ERR_UNSUPPORTED, and the caller falls back to afcm_filtered_lrelu_tc): x fp16 with a 16-byte aligned base and
 * row / channel / sample strides that are multiples of 8 elements (pad the row pitch), no bias (b == NULL: the convolution
 * epilogue adds it), (up,down) in {(2,2),(4,2),(2,4)} with 6*up / 6*down taps, clamp in [2^-10, 2^10]; y fp16 or fp32, any
 * strides with ys[3] == 1; skip (same dtype / strides as y) is added as skip * out_scale.  Tolerance: 2e-3 of max|y|. */
int afcm_filtered_lrelu_t5(const void* x, const int64_t* xs, int x_dtype, void* y, const int64_t* ys, int y_dtype,
                           const float* b, const void* skip,
                           int N, int C, int xh, int xw, int yh, int yw,
                           const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                           int up, int down, int px0, int px1, int py0, int py1,
                           float gain, float slope, float clamp, float out_scale, int flip_filter,
                           void* stream);
/* The plan afcm_filtered_lrelu_t5 would use (host only, no device access; not part of the stable ABI): fills `out` with
 * the fields of afcm::T5Plan (afcm_b200/csrc/flr_t5_plan.h), kw = 0 selects the strip width automatically.  Used by the
 * CPU-tier emulation test of the kernel's index algebra. */
int afcm_filtered_lrelu_t5_plan(int xh, int xw, int up, int down, int px0, int px1, int py0, int py1, int kw, int* out, int n_out);
/* Development aid (not part of the stable ABI): a device buffer of 4 x 2048 int64 that CTA 0 of the following
 * afcm_filtered_lrelu_t5 launches fills with (event code, clock64) records per warp role; NULL switches tracing off. */
int afcm_filtered_lrelu_t5_trace(void* dev_buffer);

/* Tensor-core filtered_lrelu WITH the sign tensor (afcm_b200/csrc/flr_tcs.cu): the training-step variant.  Same
 * arguments, sign-tensor format and sign_mode meaning as afcm_filtered_lrelu, so forward (SIGN_WRITE) and backward
 * (SIGN_READ, up/down exchanged by the caller as in OPS/filtered_lrelu.py:252-266) interoperate with the exact kernel.
 * x, y, b, skip are float32; `op_dtype` (AFCM_F16 / AFCM_BF16) is the operand type of the four banded-Toeplitz
 * m16n8k16 products (fp32 accumulation, fp32 activation).  Geometries: (up,down) in {(2,2),(2,4),(4,2)} with 6*up /
 * 6*down taps, else AFCM_ERR_UNSUPPORTED.  Tolerance: 2e-3 of max|y| with F16 operands (tests/test_gpu_flr_tcs.py). */
int afcm_filtered_lrelu_tcs(const void* x, const int64_t* xs, void* y, const int64_t* ys,
                            const void* b, const void* skip, int op_dtype,
                            int N, int C, int xh, int xw, int yh, int yw,
                            const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                            int up, int down, int px0, int px1, int py0, int py1,
                            float gain, float slope, float clamp, float out_scale, int flip_filter,
                            int sign_mode, void* signs, int sign_h, int sign_wb, int sx, int sy,
                            void* stream);

/* filtered_lrelu_act_ -- replaces filtered_lrelu_plugin.filtered_lrelu_act_ (OPS/filtered_lrelu.cpp:213-290):
 * in-place gain / lrelu / clamp with optional sign write or read on a dense [planes,h,w] tensor.     */
int afcm_filtered_lrelu_act(void* x, int dtype, int64_t planes, int h, int w,
                            float gain, float slope, float clamp,
                            int sign_mode, void* signs, int sign_h, int sign_wb, int sx, int sy,
                            void* stream);

/* ---------------------------------------------------------------------------------------------- */
/* upfirdn2d -- replaces upfirdn2d_plugin.upfirdn2d (OPS/upfirdn2d.cpp:16-98)                        */
/* x [planes,xh,xw] dense, y [planes,yh,yw] dense, f_host [fh][fw] dense 2-D taps (a separable filter */
/* is two calls, as in OPS/upfirdn2d.py:241-245).                                                    */
int afcm_upfirdn2d(const void* x, void* y, int dtype, int64_t planes, int xh, int xw, int yh, int yw,
                   const float* f_host, int fh, int fw,
                   int upx, int upy, int downx, int downy, int px0, int px1, int py0, int py1,
                   int flip_filter, float gain, void* stream);

/* ---------------------------------------------------------------------------------------------- */
/* bias_act -- replaces bias_act_plugin.bias_act (OPS/bias_act.cpp:32-90)                            */
/* grad 0: y = clamp(act(x + b) * gain).  grad 1: dx from dy (= x argument), xref, yref.  grad 2:    */
/* second-order term from dy-of-dy (= x), xref, yref, dy.  b indexes (i / step_b) % size_b.          */
int afcm_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy,
                  void* y, int dtype, int64_t n, int64_t step_b, int64_t size_b,
                  int grad, int act, float alpha, float gain, float clamp, void* stream);

/* ---------------------------------------------------------------------------------------------- */
/* convolution / modulated convolution -- replace modulated_conv2d (NET:25-64), the encoder conv    */
/* (NET:503-505) and Conv2dLayer's conv (models/networks/CoModGAN/layers.py:157) which the reference */
/* routes to cuDNN through conv2d_gradfix.conv2d (OPS/conv2d_gradfix.py:37-40).                       */
/*
 * Shared formulation:  y[n,o] = ocoef[n,o] * sum_{i,ky,kx} w[o,i,ky,kx] * (icoef[n,i] * x[n,i])
 * (stride 1, zero padding `pad`, correlation).  For modulated_conv2d icoef = normalised style *
 * input_gain and ocoef = demodulation coefficient (both from afcm_modconv_coefs), which is
 * algebraically the reference's per-sample weight w*s*d*g without materialising [N,O,I,k,k].
 * icoef / ocoef may be NULL (= 1).
 */

/* exact-fp32 SIMT path (parity path, |err| ~ 1e-6 relative). */
int afcm_conv2d_f32(const float* x, const float* w, const float* icoef, const float* ocoef, float* y,
                    int N, int Ci, int H, int W, int Co, int ksize, int pad, void* stream);

/* Per-layer weight preparation.  w [Co,Ci,k,k] fp32 (as stored in the reference state_dict).
 * pre_scale: multiply every weight (encoder: 1/sqrt(Ci*k*k), NET:464,503).  normalize != 0: divide each
 * output channel by its RMS (NET:42).  Writes any of (NULL = skip):
 *   w_f32  [Co,Ci,k,k]        prepared weights for afcm_conv2d_f32
 *   w_tc   [k*k,Co_pad,Ci_pad] 16-bit (tc_dtype AFCM_F16/AFCM_BF16) tap-major K-major tile source for the
 *                              tcgen05 kernel, zero padded (Co_pad = roundup(Co,16), Ci_pad = roundup(Ci,64))
 *   wsq    [Co,Ci]             sum over taps of the prepared weight squared (demodulation GEMV input) */
int afcm_conv_weight_prep(const float* w, int Co, int Ci, int ksize, float pre_scale, int normalize,
                          float* w_f32, void* w_tc, int tc_dtype, float* wsq, void* stream);

/* Modulation coefficients (NET:41-57).  styles [N,Ci]; demodulate != 0: s_hat = s * rsqrt(mean(s^2)) over
 * the whole [N,Ci] batch, ocoef[n,o] = rsqrt(sum_i wsq[o,i] * s_hat[n,i]^2 + 1e-8); else s_hat = s,
 * ocoef = 1.  icoef[n,i] = s_hat[n,i] * input_gain (input_gain: device scalar pointer or NULL). */
int afcm_modconv_coefs(const float* styles, const float* wsq, const float* input_gain,
                       float* icoef, float* ocoef, int N, int Ci, int Co, int demodulate, void* stream);
/* The same with the input gain given as the layer's magnitude_ema buffer: gain_rsqrt != 0 applies rsqrt to *input_gain in the
 * kernel (input_gain = magnitude_ema.rsqrt(), NET:346), which keeps a separate tiny kernel per layer out of the forward. */
int afcm_modconv_coefs_ema(const float* styles, const float* wsq, const float* input_gain, int gain_rsqrt,
                           float* icoef, float* ocoef, int N, int Ci, int Co, int demodulate, void* stream);

/* tcgen05 / TMEM implicit-GEMM path (16-bit operands, fp32 accumulation in tensor memory).
 * Step 1: pack activations into the conv-ready layout  xp [N, H*(W+2), Ci_pad8] 16-bit: channel-innermost
 * flat planes of row pitch W+2 (two trailing zero pixels per row, so horizontal taps wrap onto zeros;
 * vertical taps fall off the plane where TMA zero-fills); afcm_conv_tc_plane_elems(H,W,Ci) elements per
 * sample.  icoef is folded here (modulation on the activation side).
 * Step 2: afcm_conv2d_tc runs the GEMM  D[pixel, o] = sum_{tap,i} A_tap[pixel,i] * B_tap[o,i]  with
 * M = 128 flat pixels, N = up to 256 output channels per CTA, TMA-fed, and writes NCHW
 * y [N,Co,H+2*pad-2,W+2*pad-2] scaled by ocoef (y_dtype AFCM_F32 or AFCM_F16; x_dtype of the pack step
 * likewise: fp16 activations stay 16-bit between the layers of the fast inference path).  bias [Co] or NULL
 * is added after the scale: y = acc * ocoef + bias (the bias of the filtered_lrelu / bias_act that follows,
 * NET:371, NET:510, which then runs bias-free).  ksize must be 3, pad 0, 1 or 2 (pad 0 = the data gradient of a
 * full-padding convolution, see afcm_conv2d_wgrad_tc below). */
int64_t afcm_conv_tc_plane_elems(int H, int W, int Ci);
int afcm_conv_tc_pack(const void* x, int x_dtype, const float* icoef, void* xp, int tc_dtype,
                      int N, int Ci, int H, int W, void* stream);
/* the same with planes stored at a row pitch of x_pitch >= W elements (the W + 2 pitch with zero pad columns that
 * afcm_filtered_lrelu_tc_padded writes for the convolution that follows) */
int afcm_conv_tc_pack_pitched(const void* x, int x_dtype, int x_pitch, const float* icoef, void* xp, int tc_dtype,
                              int N, int Ci, int H, int W, void* stream);
int afcm_conv2d_tc(const void* xp, const void* w_tc, const float* ocoef, const float* bias, void* y, int y_dtype, int tc_dtype,
                   int N, int Ci, int H, int W, int Co, int pad, void* stream);
/* The same GEMM WITHOUT the pack step (NET:365-377 / NET:503-511: the convolution consumes what the preceding
 * filtered_lrelu wrote): x [N,Ci,H,W] fp16 NCHW contiguous is read directly (TMA ring of raw [64 channels][152 elements]
 * tiles), eight producer warps of the kernel transpose 8-channel x 8-pixel blocks in registers and store them into the
 * K-major swizzled tile the tensor core reads; icoef
 * [N,Ci] or NULL is applied on the way (fp32 product, one rounding -- bit-identical to afcm_conv_tc_pack + afcm_conv2d_tc).
 * x_pitch = W: dense planes.  x_pitch = W + 2: planes [N,Ci,H,W+2] whose last two columns are zeros (written by
 * afcm_filtered_lrelu_tc_padded) -- the flat plane of the GEMM formulation then exists in memory and the producers copy
 * without row bookkeeping (measured 15-20 % faster on the 64-channel layers).
 * Full padding (2), fp16 operands; returns AFCM_ERR_UNSUPPORTED for an odd W, H * x_pitch not a multiple of 8 or a base
 * address that is not 16-byte aligned (the raw tiles travel by TMA from the flat [N][Ci][H x_pitch] view of x). */
int afcm_conv2d_tc_nchw(const void* x, int x_pitch, const float* icoef, const void* w_tc, const float* ocoef, const float* bias, void* y,
                        int y_dtype, int N, int Ci, int H, int W, int Co, void* stream);

/* Debug / tuning aids, not part of the stable ABI: progress markers of the tcgen05 kernel in mapped host memory;
 * forced TMA->MMA ring depth (0 = automatic: as many stages as fit, at most 8). */
int afcm_conv_tc_set_stages(int stages);
int afcm_conv_tc_set_rowreuse(int mode);   /* A-tile reuse across the kx taps: -1 automatic, 0 off, 1 on, 2 on without resident weights */
void* afcm_conv_tc_debug_buffer(int enable);
int afcm_conv_tc_set_issuers(int n);         /* development switch: 1 (default) or 2 MMA-issuing warps per CTA */
int afcm_conv_tc_trace(void* dev_buffer);    /* timeline records of CTA 0 (3 roles x 4096 int64: code << 48 | clock64); NULL switches it off */

/* ---------------------------------------------------------------------------------------------- */
/* convolution gradients -- the backward pass the reference obtains from autograd of F.conv2d (cuDNN)  */
/* through conv2d_gradfix.conv2d (OPS/conv2d_gradfix.py:37-40) inside modulated_conv2d (NET:25-64), the */
/* encoder conv (NET:503-505) and Conv2dLayer.  (afcm_b200/csrc/conv2d_bwd.cu)                        */
/*
 * With y[n,o] = ocoef[n,o] * conv(icoef[n,i] * x[n,i], w, pad):
 *   data gradient   : dxm = conv(ocoef * dy, flip(w)^T, pad' = k-1-pad) is a FORWARD call (afcm_conv2d_f32 or
 *                     afcm_conv_tc_pack + afcm_conv2d_tc) on transposed, flipped weights with icoef := ocoef;
 *                     dx = icoef * dxm.
 *   afcm_plane_dot_scale : out[p] = <a[p,:], b[p,:]> (/ div[p] when div != NULL), then a[p,:] *= coef[p] in place when
 *                     coef != NULL; a, b [planes, hw] dense fp32.  Gives d_icoef = <dxm, x> followed by dx = icoef*dxm
 *                     in one launch, and d_ocoef = <dy, y> / ocoef.
 *   weight gradient : dw[o,i,ky,kx] = sum_{n,p} (ocoef dy)[n,o,p] * (icoef x)[n,i,p+(ky-pad,kx-pad)], fp32 [Co,Ci,k,k]
 *                     (overwritten).  _f32: exact SIMT path from the NCHW fp32 tensors (ksize 1 or 3).  _tc: mma.sync
 *                     tensor-core path (fp32 accumulation) from the two 16-bit flat-plane tensors of
 *                     afcm_conv_tc_pack: dyp = pack(dy [N,Co,OH,OW], ocoef), xp = pack(x [N,Ci,H,W], icoef) -- the
 *                     same xp the forward consumed and the same dyp the data gradient consumes; ksize 3, pad 0..2.
 */
int afcm_plane_dot_scale(float* a, const float* b, const float* div, const float* coef, float* out,
                         int64_t planes, int64_t hw, void* stream);
/* out[p] = sum over the plane a[p,:] (fp32, dense [planes, hw]): the per-plane part of the bias gradient
 * db = dx.sum([0,2,3]) of filtered_lrelu (OPS/filtered_lrelu.py:264-265) and bias_act (OPS/bias_act.py:169-170). */
int afcm_plane_sum(const float* a, float* out, int64_t planes, int64_t hw, void* stream);
int afcm_conv2d_wgrad_f32(const float* dy, const float* x, const float* icoef, const float* ocoef, float* dw,
                          int N, int Ci, int H, int W, int Co, int ksize, int pad, void* stream);
int afcm_conv2d_wgrad_tc(const void* dyp, const void* xp, float* dw, int tc_dtype,
                         int N, int Ci, int H, int W, int Co, int pad, void* stream);
/* The same weight gradient on tcgen05 / TMEM (afcm_b200/csrc/conv2d_wgrad_tc5.cu): pixels are the reduction dimension,
 * both operands are TMA-fed MN-major tiles of the packed planes, three accumulators (kx) of 128 x 128 per CTA in tensor
 * memory, split over pixels with per-split partial sums in `workspace` (device memory, at least
 * afcm_conv2d_wgrad_tc_workspace() bytes) reduced in a fixed order -- deterministic, unlike the atomics of the mma.sync
 * variant.  Same arguments otherwise. */
int64_t afcm_conv2d_wgrad_tc_workspace(int N, int Ci, int H, int W, int Co, int pad);
int afcm_conv2d_wgrad_tc5(const void* dyp, const void* xp, float* dw, void* workspace, int64_t workspace_bytes,
                          int tc_dtype, int N, int Ci, int H, int W, int Co, int pad, void* stream);

/* Fused Adam step on flat fp32 buffers (the generator optimiser of the reference training loop, train.py /
 * models/base_model.py: torch.optim.Adam; scrub != 0 applies the reference's gradient scrub
 * nan_to_num(nan=0, posinf=1e5, neginf=-1e5), train.py:67-77, before the update).  grad_scale multiplies the
 * gradient first (1/world_size after a sum all-reduce).  step counts from 1. */
int afcm_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   float lr, float beta1, float beta2, float eps, int step, float grad_scale, int scrub,
                   void* stream);
/* Generator EMA (train.py:67-77): param_ema[i] = lerp(param[i], param_ema[i], beta) = param + beta (param_ema - param) over a
 * flat buffer of n fp32 parameters. */
int afcm_ema_lerp(float* param_ema, const float* param, int64_t n, float beta, void* stream);

/* ---------------------------------------------------------------------------------------------- */
/* small fused kernels                                                                              */

/* FullyConnectedLayer.forward (NET:89-101): y[n,o] = act(sum_i x[n,i] * w[o,i] * weight_gain + b[o] *
 * bias_gain) * act_gain.  x row stride ldx (elements) lets the caller feed a concatenation view. */
int afcm_fully_connected(const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy,
                         int N, int in_features, int out_features,
                         float weight_gain, float bias_gain, int act, float alpha, float act_gain,
                         void* stream);
/* `groups` (<= 16) independent linear layers of the same input width in ONE launch: y[g] [N, out_features[g]] =
 * (x[g] [N, in_features] (row stride ldx) * w[g]^T * weight_gain + b[g] * bias_gain) * out_gain[g]; x, w, b, y are HOST arrays of
 * device pointers.  The 15 affine layers of the synthesis network (NET:349-352) depend only on the mapped styles, so the whole
 * set runs before the first synthesis layer instead of one small launch in front of every convolution. */
int afcm_fully_connected_grouped(int groups, const float* const* x, int64_t ldx, const float* const* w, const float* const* b,
                                 float* const* y, const int* out_features, const float* out_gain,
                                 int N, int in_features, float weight_gain, float bias_gain, void* stream);

/* normalize_2nd_moment of MappingNetwork (NET:142,146): y = x * rsqrt(mean(x^2, dim=1) + eps). */
int afcm_normalize_2nd_moment(const float* x, int64_t ldx, float* y, int64_t ldy, int N, int F, float eps,
                              void* stream);

/* AdaptiveAvgPool2d((oh,ow)) on [planes,H,W] (NET:636,683). */
int afcm_adaptive_avgpool(const float* x, float* y, int64_t planes, int H, int W, int oh, int ow, void* stream);

/* F.pad(img, margin) of NET:669, optionally fused with the uint8 -> [-1,1] input normalisation
 * (data/augment/transforms.py:604-616): when lut_host != NULL, x is uint8 and lut_host[256] holds the
 * float32 value of every code (computed on the host in float64 exactly like the reference transform). */
int afcm_pad_input(const void* x, float* y, const float* lut_host, int64_t planes, int H, int W, int margin,
                   void* stream);

/* ToRGB layer in one pass (NET:353-372 with is_torgb, NET:699-700): 1x1 modulated convolution with Co <= 4 output
 * channels, then the (up 1, down 1) filtered_lrelu = bias, gain, leaky slope, clamp, then the output scale:
 *   y[n,o,p] = clamp(lrelu((ocoef[n,o] * sum_i w[o,i] * icoef[n,i] * x[n,i,p] + b[o]) * gain, slope), +-clamp) * out_scale
 * x [N,Ci,HW] float32 or float16 (dense), y [N,Co,HW] float32.  icoef / ocoef / b may be NULL.  Returns
 * AFCM_ERR_UNSUPPORTED for Co > 4, Ci > 1024 or odd HW (the caller then composes conv + filtered_lrelu). */
int afcm_torgb(const void* x, int x_dtype, const float* w, const float* icoef, const float* ocoef, const float* b,
               float* y, int N, int Ci, int Co, int64_t HW, float gain, float slope, float clamp, float out_scale,
               void* stream);

/* SynthesisInput Fourier features (NET:198-243), API parity only (AFCM never instantiates it).
 * t [N,4] = affine(w); freqs [C,2]; phases [C]; weight [C,C]; y [N,C,size,size]. */
int afcm_fourier_features(const float* t, const float* freqs, const float* phases, const float* weight,
                          const float* transform3x3, float* y, int N, int C, int size_h, int size_w,
                          float sampling_rate, float bandwidth, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AFCM_B200_H */
