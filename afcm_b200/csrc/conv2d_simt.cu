// conv2d_simt.cu -- exact-fp32 convolution (the parity path), per-layer weight preparation and the
// modulation / demodulation coefficients of modulated_conv2d.
//
// Reference call sites (paths relative to the reference repo, NET = models/networks/stylegan3/
// networks_stylegan3.py): modulated_conv2d NET:25-64, encoder conv NET:503-505, Conv2dLayer
// models/networks/CoModGAN/layers.py:153-162 -- all end in F.conv2d (cuDNN).  Here the per-sample
// weight [N,O,I,k,k] is never materialised: modulation is a per-(n,i) scale on the activations,
// demodulation a per-(n,o) scale on the outputs (see include/afcm_b200.h).
#include "afcm_common.cuh"

namespace afcm {

constexpr int CV_TW = 16, CV_TH = 16;      // output pixels per CTA
constexpr int CV_CO = 64;                  // output channels per CTA
constexpr int CV_CK = 8;                   // input channels per smem stage
constexpr int CV_THREADS = 256;

template <int KS>
__global__ void __launch_bounds__(CV_THREADS)
conv2d_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ icoef,
                  const float* __restrict__ ocoef, float* __restrict__ y,
                  int Ci, int H, int W, int Co, int pad, int OH, int OW, int tiles_x)
{
    constexpr int IW = CV_TW + KS - 1, IH = CV_TH + KS - 1, IP = IW + 2;   // input tile, padded pitch
    __shared__ float xs[CV_CK][IH][IP];
    __shared__ __align__(16) float ws[CV_CK][KS * KS][CV_CO];
    const int tid = threadIdx.x;
    const int n = blockIdx.z, co0 = blockIdx.y * CV_CO;
    const int ty0 = (blockIdx.x / tiles_x) * CV_TH, tx0 = (blockIdx.x % tiles_x) * CV_TW;
    // thread -> 8 consecutive pixels of one tile row, 8 consecutive output channels
    const int pg = tid & 31, cg = tid >> 5;
    const int prow = pg >> 1, pcol = (pg & 1) * 8;
    float acc[8][8];
#pragma unroll
    for (int p = 0; p < 8; p++)
#pragma unroll
        for (int q = 0; q < 8; q++) acc[p][q] = 0.f;

    const float* xn = x + (long long)n * Ci * H * W;
    for (int ci0 = 0; ci0 < Ci; ci0 += CV_CK) {
        for (int i = tid; i < CV_CK * IH * IW; i += CV_THREADS) {
            const int c = i / (IH * IW), r = i % (IH * IW), iy = r / IW, ix = r % IW;
            const int gy = ty0 + iy - pad, gx = tx0 + ix - pad, ci = ci0 + c;
            float v = 0.f;
            if (ci < Ci && gy >= 0 && gy < H && gx >= 0 && gx < W) {
                v = xn[((long long)ci * H + gy) * W + gx];
                if (icoef) v *= icoef[n * Ci + ci];
            }
            xs[c][iy][ix] = v;
        }
        for (int i = tid; i < CV_CK * KS * KS * CV_CO; i += CV_THREADS) {
            const int co = i / (CV_CK * KS * KS), r = i % (CV_CK * KS * KS), c = r / (KS * KS), t = r % (KS * KS);
            const int o = co0 + co, ci = ci0 + c;
            ws[c][t][co] = (o < Co && ci < Ci) ? w[((long long)o * Ci + ci) * KS * KS + t] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CV_CK; c++) {
#pragma unroll
            for (int ky = 0; ky < KS; ky++) {
                float px[8 + KS - 1];
#pragma unroll
                for (int j = 0; j < 8 + KS - 1; j++) px[j] = xs[c][prow + ky][pcol + j];
#pragma unroll
                for (int kx = 0; kx < KS; kx++) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&ws[c][ky * KS + kx][cg * 8]);
                    const float4 w1 = *reinterpret_cast<const float4*>(&ws[c][ky * KS + kx][cg * 8 + 4]);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int p = 0; p < 8; p++)
#pragma unroll
                        for (int q = 0; q < 8; q++) acc[p][q] += px[p + kx] * wv[q];
                }
            }
        }
        __syncthreads();
    }
    const int oy = ty0 + prow;
    if (oy >= OH) return;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int o = co0 + cg * 8 + q;
        if (o >= Co) break;
        const float oc = ocoef ? ocoef[n * Co + o] : 1.f;
        float* yr = y + (((long long)n * Co + o) * OH + oy) * OW;
#pragma unroll
        for (int p = 0; p < 8; p++) {
            const int ox = tx0 + pcol + p;
            if (ox < OW) yr[ox] = acc[p][q] * oc;
        }
    }
}

// ---- weight preparation: one CTA per output channel ----------------------------------------------
template <typename TC>
__global__ void __launch_bounds__(256)
weight_prep_kernel(const float* __restrict__ w, int Co, int Ci, int KK, float pre_scale, int normalize,
                   float* __restrict__ w_f32, TC* __restrict__ w_tc, int co_pad, int ci_pad, float* __restrict__ wsq)
{
    __shared__ float red[8];
    __shared__ float s_scale;
    const int o = blockIdx.x, tid = threadIdx.x;
    const int n = Ci * KK;
    const float* wo = w + (long long)o * n;
    float scale = pre_scale;
    if (normalize) {                       // NET:42  w * rsqrt(mean(w^2, [1,2,3]))
        float s = 0.f;
        for (int i = tid; i < n; i += 256) { const float v = wo[i] * pre_scale; s += v * v; }
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
        if ((tid & 31) == 0) red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int k = 0; k < 8; k++) t += red[k];
            s_scale = pre_scale * rsqrtf(t / (float)n);
        }
        __syncthreads();
        scale = s_scale;
    }
    for (int i = tid; i < n; i += 256) {
        const float v = wo[i] * scale;
        if (w_f32) w_f32[(long long)o * n + i] = v;
        if (w_tc) { const int ci = i / KK, t = i % KK; w_tc[((long long)t * co_pad + o) * ci_pad + ci] = (TC)v; }
    }
    if (wsq)
        for (int ci = tid; ci < Ci; ci += 256) {
            float s = 0.f;
            for (int t = 0; t < KK; t++) { const float v = wo[ci * KK + t] * scale; s += v * v; }
            wsq[(long long)o * Ci + ci] = s;
        }
}

// ---- modulation coefficients ----------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
modconv_icoef_kernel(const float* __restrict__ styles, const float* __restrict__ input_gain, float* __restrict__ icoef,
                     int total, int demodulate, int gain_rsqrt)
{
    __shared__ float red[32];
    __shared__ float s_r;
    const int tid = threadIdx.x;
    float r = 1.f;
    if (demodulate) {                      // NET:43  s * rsqrt(mean(s^2)) over batch and channels
        float s = 0.f;
        for (int i = tid; i < total; i += 1024) { const float v = styles[i]; s += v * v; }
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
        if ((tid & 31) == 0) red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int k = 0; k < 32; k++) t += red[k];
            s_r = rsqrtf(t / (float)total);
        }
        __syncthreads();
        r = s_r;
    }
    const float g = input_gain ? (gain_rsqrt ? rsqrtf(*input_gain) : *input_gain) : 1.f;     // gain_rsqrt: the pointer is magnitude_ema (NET:346)
    for (int i = tid; i < total; i += 1024) icoef[i] = styles[i] * r * g;
}

__global__ void __launch_bounds__(256)
modconv_ocoef_kernel(const float* __restrict__ icoef, const float* __restrict__ wsq, const float* __restrict__ input_gain,
                     float* __restrict__ ocoef, int N, int Ci, int Co, int demodulate, int gain_rsqrt)
{
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5), n = blockIdx.y;
    if (o >= Co) return;
    if (!demodulate) { if (lane == 0) ocoef[n * Co + o] = 1.f; return; }
    const float g = input_gain ? (gain_rsqrt ? rsqrtf(*input_gain) : *input_gain) : 1.f;
    float s = 0.f;
    for (int i = lane; i < Ci; i += 32) { const float v = icoef[n * Ci + i]; s += wsq[(long long)o * Ci + i] * v * v; }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if (lane == 0) ocoef[n * Co + o] = rsqrtf(s / (g * g) + 1e-8f);      // NET:51
}

}  // namespace afcm

using namespace afcm;

extern "C" int afcm_conv2d_f32(const float* x, const float* w, const float* icoef, const float* ocoef, float* y,
                               int N, int Ci, int H, int W, int Co, int ksize, int pad, void* stream)
{
    AFCM_CHECK_ARG(x && w && y, "x, w and y must be given");
    AFCM_CHECK_ARG(N > 0 && Ci > 0 && Co > 0 && H > 0 && W > 0, "empty problem");
    AFCM_CHECK_ARG(pad >= 0, "negative padding");
    if (ksize != 1 && ksize != 3) { set_error("conv2d_f32: kernel size %d not supported (1 or 3)", ksize); return AFCM_ERR_UNSUPPORTED; }
    const int OH = H + 2 * pad - ksize + 1, OW = W + 2 * pad - ksize + 1;
    AFCM_CHECK_ARG(OH > 0 && OW > 0, "output must be at least 1x1");
    AFCM_CHECK_ARG(N <= 65535 && ceil_div(Co, CV_CO) <= 65535, "batch or channel count too large");
    const int tiles_x = ceil_div(OW, CV_TW), tiles_y = ceil_div(OH, CV_TH);
    dim3 grid(tiles_x * tiles_y, ceil_div(Co, CV_CO), N);
    cudaStream_t st = (cudaStream_t)stream;
    if (ksize == 3) conv2d_f32_kernel<3><<<grid, CV_THREADS, 0, st>>>(x, w, icoef, ocoef, y, Ci, H, W, Co, pad, OH, OW, tiles_x);
    else conv2d_f32_kernel<1><<<grid, CV_THREADS, 0, st>>>(x, w, icoef, ocoef, y, Ci, H, W, Co, pad, OH, OW, tiles_x);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_conv_weight_prep(const float* w, int Co, int Ci, int ksize, float pre_scale, int normalize,
                                     float* w_f32, void* w_tc, int tc_dtype, float* wsq, void* stream)
{
    AFCM_CHECK_ARG(w && Co > 0 && Ci > 0 && ksize > 0, "empty weight");
    AFCM_CHECK_ARG(!w_tc || tc_dtype == AFCM_F16 || tc_dtype == AFCM_BF16, "w_tc dtype must be F16 or BF16");
    const int co_pad = (Co + 15) & ~15, ci_pad = (Ci + 63) & ~63, KK = ksize * ksize;
    cudaStream_t st = (cudaStream_t)stream;
    if (w_tc) AFCM_CUDA(cudaMemsetAsync(w_tc, 0, (size_t)KK * co_pad * ci_pad * 2, st));
    if (w_tc && tc_dtype == AFCM_BF16)
        weight_prep_kernel<__nv_bfloat16><<<Co, 256, 0, st>>>(w, Co, Ci, KK, pre_scale, normalize, w_f32, (__nv_bfloat16*)w_tc, co_pad, ci_pad, wsq);
    else
        weight_prep_kernel<__half><<<Co, 256, 0, st>>>(w, Co, Ci, KK, pre_scale, normalize, w_f32, (__half*)w_tc, co_pad, ci_pad, wsq);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_modconv_coefs(const float* styles, const float* wsq, const float* input_gain,
                                  float* icoef, float* ocoef, int N, int Ci, int Co, int demodulate, void* stream)
{
    return afcm_modconv_coefs_ema(styles, wsq, input_gain, 0, icoef, ocoef, N, Ci, Co, demodulate, stream);
}

extern "C" int afcm_modconv_coefs_ema(const float* styles, const float* wsq, const float* input_gain, int gain_rsqrt,
                                      float* icoef, float* ocoef, int N, int Ci, int Co, int demodulate, void* stream)
{
    AFCM_CHECK_ARG(styles && icoef, "styles and icoef must be given");
    AFCM_CHECK_ARG(N > 0 && Ci > 0 && Co > 0, "empty problem");
    AFCM_CHECK_ARG(!demodulate || (wsq && ocoef), "demodulation needs wsq and ocoef");
    AFCM_CHECK_ARG(N <= 65535, "batch too large");
    cudaStream_t st = (cudaStream_t)stream;
    modconv_icoef_kernel<<<1, 1024, 0, st>>>(styles, input_gain, icoef, N * Ci, demodulate, gain_rsqrt);
    AFCM_LAUNCH_CHECK();
    count_launch();
    if (ocoef) {
        modconv_ocoef_kernel<<<dim3(ceil_div(Co, 8), N), 256, 0, st>>>(icoef, wsq, input_gain, ocoef, N, Ci, Co, demodulate, gain_rsqrt);
        AFCM_LAUNCH_CHECK();
        count_launch();
    }
    return AFCM_OK;
}
