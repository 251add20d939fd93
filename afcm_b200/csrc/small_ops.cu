// small_ops.cu -- the small fused kernels around the two heavy operators:
//   afcm_fully_connected      FullyConnectedLayer.forward            (NET:89-101)
//   afcm_normalize_2nd_moment MappingNetwork input normalisation      (NET:142,146)
//   afcm_adaptive_avgpool     AdaptiveAvgPool2d((4,4))                (NET:636,683)
//   afcm_pad_input            F.pad(img, margin) (NET:669) + uint8 -> [-1,1] (data/augment/transforms.py:604-616)
//   afcm_fourier_features     SynthesisInput.forward                  (NET:198-243)
// NET = models/networks/stylegan3/networks_stylegan3.py of the reference.  These are latency-bound
// (SURVEY.md 8(d)): one launch each, no intermediate tensors.
#include "afcm_common.cuh"

namespace afcm {

// ---- fully connected ---------------------------------------------------------------------------
// One warp per output feature, FC_NB batch rows at a time; lanes stride over the input features so the
// weight row is read once per batch tile with coalesced (128-bit when aligned) loads.
constexpr int FC_NB = 8;
constexpr int FC_WARPS = 8;

template <bool VEC>
__global__ void __launch_bounds__(FC_WARPS * 32)
fc_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w, const float* __restrict__ b,
          float* __restrict__ y, long long ldy, int N, int in_f, int out_f,
          float wg, float bg, int act, float alpha, float ag)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int o = blockIdx.x * FC_WARPS + warp;
    const int n0 = blockIdx.y * FC_NB;
    if (o >= out_f) return;
    float acc[FC_NB];
#pragma unroll
    for (int j = 0; j < FC_NB; j++) acc[j] = 0.f;
    const float* wr = w + (long long)o * in_f;
    if (VEC) {
        for (int i = lane * 4; i < in_f; i += 128) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + i));
#pragma unroll
            for (int j = 0; j < FC_NB; j++) {
                if (n0 + j < N) {
                    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)(n0 + j) * ldx + i));
                    acc[j] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
                }
            }
        }
    } else {
        for (int i = lane; i < in_f; i += 32) {
            const float wv = __ldg(wr + i);
#pragma unroll
            for (int j = 0; j < FC_NB; j++)
                if (n0 + j < N) acc[j] += __ldg(x + (long long)(n0 + j) * ldx + i) * wv;
        }
    }
#pragma unroll
    for (int j = 0; j < FC_NB; j++) {
        float v = acc[j];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        acc[j] = v;
    }
    if (lane < FC_NB && n0 + lane < N) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < FC_NB; j++) if (j == lane) v = acc[j];
        v = v * wg + (b ? b[o] * bg : 0.f);
        y[(long long)(n0 + lane) * ldy + o] = act_eval(v, act, alpha) * ag;
    }
}

__global__ void __launch_bounds__(256)
normalize_2nd_moment_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ y, long long ldy,
                            int N, int F, float eps)
{
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    const float* xr = x + (long long)row * ldx;
    float s = 0.f;
    for (int i = lane; i < F; i += 32) { const float v = xr[i]; s += v * v; }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    const float r = rsqrtf(s / (float)F + eps);
    for (int i = lane; i < F; i += 32) y[(long long)row * ldy + i] = xr[i] * r;
}

__global__ void __launch_bounds__(256)
adaptive_avgpool_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes, int H, int W, int oh, int ow)
{
    const long long total = planes * oh * ow;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % ow);
        const long long r = i / ow;
        const int oy = (int)(r % oh);
        const long long pl = r / oh;
        // torch adaptive pooling window: [floor(o*I/O), ceil((o+1)*I/O))
        const int y0 = (oy * H) / oh, y1 = ((oy + 1) * H + oh - 1) / oh;
        const int x0 = (ox * W) / ow, x1 = ((ox + 1) * W + ow - 1) / ow;
        const float* xp = x + pl * H * W;
        float s = 0.f;
        for (int yy = y0; yy < y1; yy++)
            for (int xx = x0; xx < x1; xx++) s += xp[yy * W + xx];
        y[i] = s / (float)((y1 - y0) * (x1 - x0));
    }
}

struct PadParams {
    const void* x; float* y; long long planes; int H, W, margin, use_lut;
    float lut[256];
};

__global__ void __launch_bounds__(256) pad_input_kernel(const __grid_constant__ PadParams p)
{
    const int OW = p.W + 2 * p.margin, OH = p.H + 2 * p.margin;
    const long long total = p.planes * OH * OW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW);
        const long long r = i / OW;
        const int oy = (int)(r % OH);
        const long long pl = r / OH;
        const int ix = ox - p.margin, iy = oy - p.margin;
        float v = 0.f;
        if (ix >= 0 && ix < p.W && iy >= 0 && iy < p.H) {
            const long long j = (pl * p.H + iy) * p.W + ix;
            v = p.use_lut ? p.lut[((const uint8_t*)p.x)[j]] : ((const float*)p.x)[j];
        }
        p.y[i] = v;
    }
}

// ---- Fourier features (SynthesisInput).  Per sample: build the inverse rotation/translation from
// t = (r_c, r_s, t_x, t_y) / |(r_c, r_s)|, transform freqs/phases, then y = W/sqrt(C) @ (sin(2pi(grid.f + ph)) * amp).
struct FourierParams {
    const float* t; const float* freqs; const float* phases; const float* weight; const float* transform;
    float* y; int N, C, H, W; float sampling_rate, bandwidth;
};

__global__ void __launch_bounds__(256) fourier_features_kernel(const __grid_constant__ FourierParams p)
{
    extern __shared__ float sm[];          // [C] fx, [C] fy, [C] phase, [C] amp
    float* fx = sm; float* fy = sm + p.C; float* ph = sm + 2 * p.C; float* amp = sm + 3 * p.C;
    const int n = blockIdx.y;
    const float* t = p.t + n * 4;
    const float nrm = sqrtf(t[0] * t[0] + t[1] * t[1]);
    const float rc = t[0] / nrm, rs = t[1] / nrm, tx = t[2] / nrm, ty = t[3] / nrm;
    // transforms = m_r @ m_t @ user   (NET:207-215); user transform is a 3x3 buffer (identity by default)
    float mr[9] = {rc, -rs, 0.f, rs, rc, 0.f, 0.f, 0.f, 1.f};
    float mt[9] = {1.f, 0.f, -tx, 0.f, 1.f, -ty, 0.f, 0.f, 1.f};
    float a[9], m[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        float s = 0.f; for (int k = 0; k < 3; k++) s += mr[i * 3 + k] * mt[k * 3 + j]; a[i * 3 + j] = s; }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        float s = 0.f; for (int k = 0; k < 3; k++) s += a[i * 3 + k] * p.transform[k * 3 + j]; m[i * 3 + j] = s; }
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
        const float f0 = p.freqs[c * 2], f1 = p.freqs[c * 2 + 1];
        ph[c] = p.phases[c] + f0 * m[2] + f1 * m[5];                    // freqs @ transforms[:, :2, 2:]
        const float gx = f0 * m[0] + f1 * m[3], gy = f0 * m[1] + f1 * m[4];   // freqs @ transforms[:, :2, :2]
        fx[c] = gx; fy[c] = gy;
        const float av = 1.f - (sqrtf(gx * gx + gy * gy) - p.bandwidth) / (p.sampling_rate / 2.f - p.bandwidth);
        amp[c] = fminf(fmaxf(av, 0.f), 1.f);
    }
    __syncthreads();
    const float sx = 0.5f * p.W / p.sampling_rate, sy = 0.5f * p.H / p.sampling_rate;
    const float wscale = rsqrtf((float)p.C);
    const long long hw = (long long)p.H * p.W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw * p.C; i += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(i / hw);
        const int pix = (int)(i % hw);
        const int yy = pix / p.W, xx = pix % p.W;
        // affine_grid(align_corners=False): base coordinate (2*i+1)/size - 1, scaled by theta
        const float gx = ((2.f * xx + 1.f) / p.W - 1.f) * sx;
        const float gy = ((2.f * yy + 1.f) / p.H - 1.f) * sy;
        float acc = 0.f;
        for (int c = 0; c < p.C; c++) {
            const float v = sinf((gx * fx[c] + gy * fy[c] + ph[c]) * 6.283185307179586f) * amp[c];
            acc += v * (p.weight[o * p.C + c] * wscale);
        }
        p.y[((long long)n * p.C + o) * hw + pix] = acc;
    }
}


// ---- ToRGB: modulated 1x1 convolution with a handful of output channels + bias + gain/lrelu/clamp + output scale
// (NET:353-372 with is_torgb, NET:699-700).  HBM-bound: every input plane is read once (pixel pairs, coalesced),
// the C_out <= 4 dot products stay in registers; the per-sample weights w[o,i] * icoef[n,i] are staged in shared memory.
constexpr int TORGB_MAX_CO = 4;
constexpr int TORGB_MAX_CI = 1024;

__device__ __forceinline__ float2 ld_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld_pair(const __half* p) { return __half22float2(*reinterpret_cast<const __half2*>(p)); }

template <typename T, int CO>
__global__ void __launch_bounds__(256)
torgb_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ icoef,
             const float* __restrict__ ocoef, const float* __restrict__ b, float* __restrict__ y,
             int Ci, long long HW, float gain, float slope, float clamp, float out_scale)
{
    __shared__ float cw[CO][TORGB_MAX_CI];
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < CO * Ci; i += blockDim.x) {
        const int o = i / Ci, c = i - o * Ci;
        cw[o][c] = w[o * Ci + c] * (icoef ? icoef[(long long)n * Ci + c] : 1.f);
    }
    __syncthreads();
    const T* xn = x + (long long)n * Ci * HW;
    for (long long p = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x); p < HW; p += 2LL * gridDim.x * blockDim.x) {
        float2 acc[CO];
#pragma unroll
        for (int o = 0; o < CO; o++) acc[o] = make_float2(0.f, 0.f);
#pragma unroll 8
        for (int c = 0; c < Ci; c++) {
            const float2 v = ld_pair(xn + c * HW + p);
#pragma unroll
            for (int o = 0; o < CO; o++) { acc[o].x = fmaf(v.x, cw[o][c], acc[o].x); acc[o].y = fmaf(v.y, cw[o][c], acc[o].y); }
        }
#pragma unroll
        for (int o = 0; o < CO; o++) {
            const float oc = ocoef ? ocoef[(long long)n * CO + o] : 1.f;
            const float bb = b ? b[o] : 0.f;
            float v0 = (acc[o].x * oc + bb) * gain, v1 = (acc[o].y * oc + bb) * gain;
            v0 = v0 < 0.f ? v0 * slope : v0; v1 = v1 < 0.f ? v1 * slope : v1;
            v0 = fminf(fmaxf(v0, -clamp), clamp) * out_scale; v1 = fminf(fmaxf(v1, -clamp), clamp) * out_scale;
            *reinterpret_cast<float2*>(y + ((long long)n * CO + o) * HW + p) = make_float2(v0, v1);
        }
    }
}

template <typename T>
static int launch_torgb(const void* x, const float* w, const float* icoef, const float* ocoef, const float* b, float* y,
                        int N, int Ci, int Co, long long HW, float gain, float slope, float clamp, float out_scale, cudaStream_t st)
{
    dim3 grid((unsigned)min((long long)ceil_div(HW / 2, 256), 4LL * sm_count()), N);
#define AFCM_TORGB(CO) torgb_kernel<T, CO><<<grid, 256, 0, st>>>((const T*)x, w, icoef, ocoef, b, y, Ci, HW, gain, slope, clamp, out_scale)
    switch (Co) {
    case 1: AFCM_TORGB(1); break;
    case 2: AFCM_TORGB(2); break;
    case 3: AFCM_TORGB(3); break;
    default: AFCM_TORGB(4); break;
    }
#undef AFCM_TORGB
    return AFCM_OK;
}

// ---- grouped fully connected: up to FCG_MAX independent linear layers with the same input width in ONE launch (the 15 affine
// layers of the synthesis network, NET:349-352: each maps its own [N, 1536] style input to the layer's channel count) ----------
constexpr int FCG_MAX = 16;
struct FcGroupParams {
    const float* x[FCG_MAX]; const float* w[FCG_MAX]; const float* b[FCG_MAX]; float* y[FCG_MAX];
    int out_f[FCG_MAX]; float ag[FCG_MAX];
    long long ldx;
    int N, in_f, groups;
    float wg, bg;
};

__global__ void __launch_bounds__(FC_WARPS * 32)
fc_grouped_kernel(const __grid_constant__ FcGroupParams p)
{
    const int grp = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int o = blockIdx.x * FC_WARPS + warp;
    const int n0 = blockIdx.y * FC_NB;
    const int out_f = p.out_f[grp];
    if (o >= out_f) return;
    const float* x = p.x[grp];
    const float* wr = p.w[grp] + (long long)o * p.in_f;
    float acc[FC_NB];
#pragma unroll
    for (int j = 0; j < FC_NB; j++) acc[j] = 0.f;
    for (int i = lane * 4; i < p.in_f; i += 128) {                    // in_f % 4 == 0 and 16-byte aligned rows (host-checked)
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + i));
#pragma unroll
        for (int j = 0; j < FC_NB; j++) {
            if (n0 + j < p.N) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)(n0 + j) * p.ldx + i));
                acc[j] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < FC_NB; j++) {
        float v = acc[j];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        acc[j] = v;
    }
    if (lane < FC_NB && n0 + lane < p.N) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < FC_NB; j++) if (j == lane) v = acc[j];
        const float* b = p.b[grp];
        v = v * p.wg + (b ? b[o] * p.bg : 0.f);
        p.y[grp][(long long)(n0 + lane) * out_f + o] = v * p.ag[grp];      // linear activation (NET:349: affine layers)
    }
}

static unsigned grid_for(long long total, int per_sm)
{
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace afcm

using namespace afcm;

extern "C" int afcm_fully_connected(const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy,
                                    int N, int in_features, int out_features,
                                    float weight_gain, float bias_gain, int act, float alpha, float act_gain,
                                    void* stream)
{
    AFCM_CHECK_ARG(x && w && y, "x, w and y must be given");
    AFCM_CHECK_ARG(N > 0 && in_features > 0 && out_features > 0, "empty problem");
    AFCM_CHECK_ARG(ldx >= in_features && ldy >= out_features, "row strides are smaller than the rows");
    AFCM_CHECK_ARG(act >= 1 && act <= 9, "unknown activation index %d", act);
    const bool vec = (in_features % 4 == 0) && (ldx % 4 == 0) && (((uintptr_t)x | (uintptr_t)w) % 16 == 0);
    dim3 grid(ceil_div(out_features, FC_WARPS), ceil_div(N, FC_NB));
    cudaStream_t st = (cudaStream_t)stream;
    if (vec) fc_kernel<true><<<grid, FC_WARPS * 32, 0, st>>>(x, ldx, w, b, y, ldy, N, in_features, out_features, weight_gain, bias_gain, act, alpha, act_gain);
    else fc_kernel<false><<<grid, FC_WARPS * 32, 0, st>>>(x, ldx, w, b, y, ldy, N, in_features, out_features, weight_gain, bias_gain, act, alpha, act_gain);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_fully_connected_grouped(int groups, const float* const* x, int64_t ldx, const float* const* w, const float* const* b,
                                            float* const* y, const int* out_features, const float* out_gain,
                                            int N, int in_features, float weight_gain, float bias_gain, void* stream)
{
    AFCM_CHECK_ARG(groups >= 1 && groups <= FCG_MAX, "1..%d groups", FCG_MAX);
    AFCM_CHECK_ARG(x && w && y && out_features && N > 0 && in_features > 0, "empty problem");
    AFCM_CHECK_ARG(in_features % 4 == 0 && ldx % 4 == 0 && ldx >= in_features, "the input width and row stride must be multiples of 4");
    FcGroupParams p;
    memset(&p, 0, sizeof(p));
    int max_out = 0;
    for (int g = 0; g < groups; g++) {
        AFCM_CHECK_ARG(x[g] && w[g] && y[g] && out_features[g] > 0, "group %d: null pointer or empty output", g);
        AFCM_CHECK_ARG((((uintptr_t)x[g] | (uintptr_t)w[g]) & 15) == 0, "group %d: x and w must be 16-byte aligned", g);
        p.x[g] = x[g]; p.w[g] = w[g]; p.b[g] = b ? b[g] : nullptr; p.y[g] = y[g]; p.out_f[g] = out_features[g];
        p.ag[g] = out_gain ? out_gain[g] : 1.f;
        if (out_features[g] > max_out) max_out = out_features[g];
    }
    p.ldx = ldx; p.N = N; p.in_f = in_features; p.groups = groups; p.wg = weight_gain; p.bg = bias_gain;
    dim3 grid(ceil_div(max_out, FC_WARPS), ceil_div(N, FC_NB), groups);
    fc_grouped_kernel<<<grid, FC_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_normalize_2nd_moment(const float* x, int64_t ldx, float* y, int64_t ldy, int N, int F, float eps,
                                         void* stream)
{
    AFCM_CHECK_ARG(x && y && N > 0 && F > 0, "empty problem");
    normalize_2nd_moment_kernel<<<ceil_div(N, 8), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, N, F, eps);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_adaptive_avgpool(const float* x, float* y, int64_t planes, int H, int W, int oh, int ow, void* stream)
{
    AFCM_CHECK_ARG(x && y && planes > 0 && H > 0 && W > 0 && oh > 0 && ow > 0, "empty problem");
    adaptive_avgpool_kernel<<<grid_for(planes * oh * ow, 8), 256, 0, (cudaStream_t)stream>>>(x, y, planes, H, W, oh, ow);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_pad_input(const void* x, float* y, const float* lut_host, int64_t planes, int H, int W, int margin,
                              void* stream)
{
    AFCM_CHECK_ARG(x && y && planes > 0 && H > 0 && W > 0 && margin >= 0, "empty problem");
    PadParams p;
    p.x = x; p.y = y; p.planes = planes; p.H = H; p.W = W; p.margin = margin; p.use_lut = lut_host != nullptr;
    if (lut_host) memcpy(p.lut, lut_host, sizeof(p.lut)); else memset(p.lut, 0, sizeof(p.lut));
    const long long total = planes * (H + 2 * margin) * (W + 2 * margin);
    pad_input_kernel<<<grid_for(total, 16), 256, 0, (cudaStream_t)stream>>>(p);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_fourier_features(const float* t, const float* freqs, const float* phases, const float* weight,
                                     const float* transform3x3, float* y, int N, int C, int size_h, int size_w,
                                     float sampling_rate, float bandwidth, void* stream)
{
    AFCM_CHECK_ARG(t && freqs && phases && weight && transform3x3 && y, "null argument");
    AFCM_CHECK_ARG(N > 0 && C > 0 && size_h > 0 && size_w > 0, "empty problem");
    AFCM_CHECK_ARG(C <= 8192, "too many channels");
    FourierParams p = {t, freqs, phases, weight, transform3x3, y, N, C, size_h, size_w, sampling_rate, bandwidth};
    dim3 grid(grid_for((long long)size_h * size_w * C, 4), N);
    fourier_features_kernel<<<grid, 256, 4 * C * sizeof(float), (cudaStream_t)stream>>>(p);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_torgb(const void* x, int x_dtype, const float* w, const float* icoef, const float* ocoef, const float* b,
                          float* y, int N, int Ci, int Co, int64_t HW, float gain, float slope, float clamp, float out_scale,
                          void* stream)
{
    AFCM_CHECK_ARG(x && w && y && N > 0 && Ci > 0 && Co > 0 && HW > 0, "empty problem");
    AFCM_CHECK_ARG(x_dtype == AFCM_F32 || x_dtype == AFCM_F16, "x must be float32 or float16");
    AFCM_CHECK_ARG(N <= 65535, "batch too large");
    if (Co > TORGB_MAX_CO || Ci > TORGB_MAX_CI || (HW & 1) || ((uintptr_t)x % (x_dtype == AFCM_F32 ? 8 : 4)) || ((uintptr_t)y % 8)) {
        set_error("torgb: needs Co <= %d, Ci <= %d, an even plane size and pair-aligned pointers", TORGB_MAX_CO, TORGB_MAX_CI);
        return AFCM_ERR_UNSUPPORTED;
    }
    if (!(clamp >= 0.f)) clamp = 3.4e38f;
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == AFCM_F32) launch_torgb<float>(x, w, icoef, ocoef, b, y, N, Ci, Co, HW, gain, slope, clamp, out_scale, st);
    else launch_torgb<__half>(x, w, icoef, ocoef, b, y, N, Ci, Co, HW, gain, slope, clamp, out_scale, st);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}
