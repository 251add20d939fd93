// filtered_lrelu.cu -- fused bias / up-FIR / leaky-ReLU / clamp / down-FIR for sm_100a.
// C-ABI: afcm_filtered_lrelu, afcm_filtered_lrelu_act, size helpers (include/afcm_b200.h).
// Design notes are in filtered_lrelu_core.h; this file holds the kernels, the tile chooser and the
// argument validation that the reference does in OPS/filtered_lrelu.cpp:16-209.
#include "afcm_common.cuh"
#include "filtered_lrelu_core.h"

namespace afcm {

constexpr int FLR_THREADS = 256;
constexpr int FLR_G = 8;

template <typename T, int UP, int FU, int DOWN, int FD, int G, int SIGN>
__global__ void __launch_bounds__(FLR_THREADS)
flr_fused_kernel(const __grid_constant__ FlrParams p)
{
    extern __shared__ __align__(16) float smem[];
    float* buf_a = smem;
    float* buf_b = smem + p.off_b;
    uint8_t* s_sign = reinterpret_cast<uint8_t*>(smem) + p.off_sign;
    const int tid = threadIdx.x;
    const FlrTile t = flr_tile3<UP, DOWN>(p, blockIdx.x, blockIdx.y, blockIdx.z);     // grid = (planes, tiles_x, tiles_y)

    flr_pass_load<T>(tid, FLR_THREADS, p, t, buf_a);
    if (SIGN == 2) flr_pass_sign_load(tid, FLR_THREADS, p, t, s_sign);
    __syncthreads();
    flr_pass_hup<UP, FU, G>(tid, FLR_THREADS, p, buf_a, buf_b);
    __syncthreads();
    flr_pass_vup<UP, FU, G, SIGN>(tid, FLR_THREADS, p, t, buf_b, buf_a, s_sign);
    __syncthreads();
    if (SIGN == 1) flr_pass_sign_flush<DOWN>(tid, FLR_THREADS, p, t, s_sign);
    flr_pass_hdown<DOWN, FD, G>(tid, FLR_THREADS, p, buf_a, buf_b);
    __syncthreads();
    flr_pass_vdown<T, DOWN, FD, G>(tid, FLR_THREADS, p, t, buf_b);
}

// Pointwise variant: no filters (ToRGB, NET:369-372) and the generic act step.  One thread = 4
// consecutive elements of a row = one sign byte.
struct ActParams {
    const void* x; void* y; const void* b; const void* skip;
    uint8_t* so; const uint8_t* si;
    long long xs_n, xs_c, xs_h, xs_w, ys_n, ys_c, ys_h, ys_w;
    int N, C, h, w, s_h, s_wb, s_ox, s_oy;
    float gain, slope, clamp, out_scale;
};

template <typename T, int SIGN>
__global__ void __launch_bounds__(256) flr_act_kernel(const __grid_constant__ ActParams p)
{
    const int w4 = (p.w + 3) >> 2;
    const long long total = (long long)p.N * p.C * p.h * w4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int xq = (int)(i % w4);
        long long r = i / w4;
        const int yy = (int)(r % p.h); r /= p.h;
        const int c = (int)(r % p.C);
        const int n = (int)(r / p.C);
        const long long plane = (long long)n * p.C + c;
        const T* xp = (const T*)p.x + n * p.xs_n + c * p.xs_c + yy * p.xs_h;
        T* yp = (T*)p.y + n * p.ys_n + c * p.ys_c + yy * p.ys_h;
        const T* kp = p.skip ? (const T*)p.skip + n * p.ys_n + c * p.ys_c + yy * p.ys_h : nullptr;
        const float bias = p.b ? (float)((const T*)p.b)[c] : 0.f;
        int code = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int xx = xq * 4 + k;
            if (xx >= p.w) break;
            float v = ((float)xp[xx * p.xs_w] + bias) * p.gain;
            if (SIGN == 2) {
                const int ex = xx + p.s_ox, ey = yy + p.s_oy;
                if (ex >= 0 && ex < (p.s_wb << 2) && ey >= 0 && ey < p.s_h) {
                    const int s = p.si[(plane * p.s_h + ey) * p.s_wb + (ex >> 2)] >> ((ex & 3) << 1);
                    if (s & 1) v *= p.slope;
                    if (s & 2) v = 0.f;
                }
            } else {
                int s = 0;
                if (v < 0.f) { v *= p.slope; s = 1; }
                if (v > p.clamp) { v = p.clamp; s = 2; }
                if (v < -p.clamp) { v = -p.clamp; s = 2; }
                code |= s << (2 * k);
            }
            if (kp) v += (float)kp[xx * p.ys_w];
            yp[xx * p.ys_w] = (T)(v * p.out_scale);
        }
        if (SIGN == 1 && yy < p.s_h && xq < p.s_wb) p.so[(plane * p.s_h + yy) * p.s_wb + xq] = (uint8_t)code;
    }
}

// ---------------------------------------------------------------------------------------------
// host side

static int g_tile_override[2] = {0, 0};

struct TileChoice { int tow, toh; size_t smem; double score; };

template <int UP, int FU, int DOWN, int FD>
static TileChoice choose_tile(FlrParams& p, int sign_stage)
{
    const int smem_cap = max_smem_optin();
    TileChoice best = {0, 0, 0, -1.0};
    auto try_tile = [&](int tow, int toh) {
        if ((tow * DOWN) % 4) return;
        FlrParams q = p;
        size_t bytes = flr_make_geom<UP, FU, DOWN, FD, FLR_G>(q, tow, toh, sign_stage);
        if ((long long)bytes > smem_cap - 1024) return;
        // useful outputs / computed up-res samples, derated by occupancy (CTAs per SM that fit)
        const double useful = (double)p.yw * p.yh;
        const double work = (double)q.tiles_x * q.tiles_y * ((double)q.ngx * UP) * ((double)q.ngy * UP);
        int per_sm = (int)((smem_cap + 1024) / (bytes + 1024)); if (per_sm > 4) per_sm = 4;
        const double occ = per_sm >= 3 ? 1.0 : (per_sm == 2 ? 0.92 : 0.7);
        const double score = useful * DOWN * DOWN / work * occ;
        if (score > best.score) best = {tow, toh, bytes, score};
    };
    if (g_tile_override[0] > 0 && g_tile_override[1] > 0) {
        try_tile(g_tile_override[0], g_tile_override[1]);
        if (best.score >= 0) return best;
    }
    const int tows[] = {96, 88, 80, 72, 64, 56, 48, 40, 32, 24, 16, 8};
    const int tohs[] = {64, 56, 48, 40, 32, 24, 16, 8};
    for (int tow : tows) for (int toh : tohs) {
        if (tow > ((p.yw + 7) & ~7) && tow != 8) continue;
        if (toh > ((p.yh + 7) & ~7) && toh != 8) continue;
        try_tile(tow, toh);
    }
    return best;
}

template <typename T, int UP, int FU, int DOWN, int FD>
static int launch_fused(FlrParams& p, int sign_mode, cudaStream_t stream)
{
    const int sign_stage = sign_mode == AFCM_SIGN_WRITE ? 1 : (sign_mode == AFCM_SIGN_READ ? 2 : 0);
    TileChoice tc = choose_tile<UP, FU, DOWN, FD>(p, sign_stage);
    if (tc.score < 0) { set_error("filtered_lrelu: no tile fits in shared memory"); return AFCM_ERR_UNSUPPORTED; }
    const size_t smem = flr_make_geom<UP, FU, DOWN, FD, FLR_G>(p, tc.tow, tc.toh, sign_stage);
    const long long planes = (long long)p.N * p.C;
    if (planes > 0x7fffffffLL || p.tiles_x > 65535 || p.tiles_y > 65535) { set_error("filtered_lrelu: too many tiles"); return AFCM_ERR_INVALID; }
    // offsets inside one plane are computed in 32 bits by the load / store passes
    auto labs64 = [](long long v) { return v < 0 ? -v : v; };
    const long long xspan = (long long)(p.xh + 256) * labs64(p.xs_h) + (long long)(p.xw + 256) * labs64(p.xs_w);   // + tile halo
    const long long yspan = (long long)p.yh * labs64(p.ys_h) + (long long)p.yw * labs64(p.ys_w);
    if (xspan > 0x3fffffffLL || yspan > 0x3fffffffLL) { set_error("filtered_lrelu: plane too large for 32-bit offsets"); return AFCM_ERR_UNSUPPORTED; }
    void (*kern)(const FlrParams) = nullptr;
    if (sign_mode == AFCM_SIGN_NONE)  kern = flr_fused_kernel<T, UP, FU, DOWN, FD, FLR_G, 0>;
    if (sign_mode == AFCM_SIGN_WRITE) kern = flr_fused_kernel<T, UP, FU, DOWN, FD, FLR_G, 1>;
    if (sign_mode == AFCM_SIGN_READ)  kern = flr_fused_kernel<T, UP, FU, DOWN, FD, FLR_G, 2>;
    AFCM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)planes, (unsigned)p.tiles_x, (unsigned)p.tiles_y), FLR_THREADS, smem, stream>>>(p);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

template <typename T>
static int launch_act(const ActParams& p, int sign_mode, cudaStream_t stream)
{
    const long long total = (long long)p.N * p.C * p.h * ((p.w + 3) >> 2);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (sign_mode == AFCM_SIGN_NONE)  flr_act_kernel<T, 0><<<(unsigned)blocks, 256, 0, stream>>>(p);
    if (sign_mode == AFCM_SIGN_WRITE) flr_act_kernel<T, 1><<<(unsigned)blocks, 256, 0, stream>>>(p);
    if (sign_mode == AFCM_SIGN_READ)  flr_act_kernel<T, 2><<<(unsigned)blocks, 256, 0, stream>>>(p);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

template <typename T>
static int dispatch_fused(FlrParams& p, int up, int fu, int down, int fd, int sign_mode, cudaStream_t stream)
{
    if (up == 2 && fu == 12 && down == 2 && fd == 12) return launch_fused<T, 2, 12, 2, 12>(p, sign_mode, stream);
    if (up == 2 && fu == 12 && down == 4 && fd == 24) return launch_fused<T, 2, 12, 4, 24>(p, sign_mode, stream);
    if (up == 4 && fu == 24 && down == 2 && fd == 12) return launch_fused<T, 4, 24, 2, 12>(p, sign_mode, stream);
    set_error("filtered_lrelu: no fused kernel for up=%d/%d taps, down=%d/%d taps", up, fu, down, fd);
    return AFCM_ERR_UNSUPPORTED;
}

}  // namespace afcm

using namespace afcm;

extern "C" int afcm_filtered_lrelu_out_size(int xh, int xw, int up, int down, int fu_taps, int fd_taps,
                                            int px0, int px1, int py0, int py1, int* yh, int* yw)
{
    // OPS/filtered_lrelu.cpp:69-78
    const long long cw = (long long)xw * up + px0 + px1 - (fu_taps - 1);
    const long long ch = (long long)xh * up + py0 + py1 - (fu_taps - 1);
    AFCM_CHECK_ARG(up >= 1 && down >= 1 && fu_taps >= 1 && fd_taps >= 1, "up, down and tap counts must be at least 1");
    AFCM_CHECK_ARG(cw > fd_taps - 1 && ch > fd_taps - 1, "upsampled buffer must be at least the size of downsampling filter");
    *yw = (int)((cw - (fd_taps - 1) + (down - 1)) / down);
    *yh = (int)((ch - (fd_taps - 1) + (down - 1)) / down);
    AFCM_CHECK_ARG(*yw > 0 && *yh > 0, "output must be at least 1x1");
    return AFCM_OK;
}

extern "C" int afcm_filtered_lrelu_sign_size(int yh, int yw, int down, int fd_taps, int* sh, int* swb)
{
    // OPS/filtered_lrelu.cpp:89-93
    const int sw_active = yw * down - (down - 1) + (fd_taps - 1);
    *sh = yh * down - (down - 1) + (fd_taps - 1);
    *swb = ((sw_active + 15) & ~15) >> 2;
    return AFCM_OK;
}

extern "C" int afcm_filtered_lrelu_set_tile(int tow, int toh)
{
    g_tile_override[0] = tow; g_tile_override[1] = toh;
    return AFCM_OK;
}

extern "C" int afcm_filtered_lrelu(const void* x, const int64_t* xs, void* y, const int64_t* ys,
                                   const void* b, const void* skip, int dtype,
                                   int N, int C, int xh, int xw, int yh, int yw,
                                   const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                                   int up, int down, int px0, int px1, int py0, int py1,
                                   float gain, float slope, float clamp, float out_scale, int flip_filter,
                                   int sign_mode, void* signs, int sign_h, int sign_wb, int sx, int sy,
                                   void* stream)
{
    AFCM_CHECK_ARG(x && y && xs && ys, "x, y and their strides must be given");
    AFCM_CHECK_ARG(dtype == AFCM_F32 || dtype == AFCM_F16, "x and b must be float16 or float32");
    AFCM_CHECK_ARG(N > 0 && C > 0 && xh > 0 && xw > 0, "x is empty");
    AFCM_CHECK_ARG((long long)N * C <= 0x7fffffffLL, "x is too large");
    AFCM_CHECK_ARG(up >= 1 && down >= 1, "up and down must be at least 1");
    AFCM_CHECK_ARG(fu_taps >= 1 && fd_taps >= 1, "fu and fd must not be empty");
    AFCM_CHECK_ARG(sign_mode >= 0 && sign_mode <= 2, "bad sign_mode");
    AFCM_CHECK_ARG(sign_mode == AFCM_SIGN_NONE || signs, "sign tensor missing");
    int eyh = 0, eyw = 0;
    int rc = afcm_filtered_lrelu_out_size(xh, xw, up, down, fu_taps, fd_taps, px0, px1, py0, py1, &eyh, &eyw);
    if (rc) return rc;
    AFCM_CHECK_ARG(eyh == yh && eyw == yw, "y has shape [%d,%d], expected [%d,%d]", yh, yw, eyh, eyw);
    if (sign_mode == AFCM_SIGN_WRITE) {
        int sh = 0, swb = 0;
        afcm_filtered_lrelu_sign_size(yh, yw, down, fd_taps, &sh, &swb);
        AFCM_CHECK_ARG(sign_h == sh && sign_wb == swb, "sign tensor has shape [%d,%d], expected [%d,%d]", sign_h, sign_wb, sh, swb);
        AFCM_CHECK_ARG(sx == 0 && sy == 0, "sign offsets must be zero when writing signs");
    }
    if (fu_taps > FLR_MAX_TAPS || fd_taps > FLR_MAX_TAPS) {
        set_error("filtered_lrelu: more than %d taps", FLR_MAX_TAPS);
        return AFCM_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;

    if (up == 1 && down == 1 && fu_taps == 1 && fd_taps == 1) {
        // Pointwise: a 1-tap filter is a scale (the reference turns it into a 1x1 "full" filter,
        // OPS/filtered_lrelu.py:184-195).  Fold both scales; padding crops / pads with zeros.
        const float fu0 = fu_host ? fu_host[0] : 1.f, fd0 = fd_host ? fd_host[0] : 1.f;
        if (px0 != 0 || px1 != 0 || py0 != 0 || py1 != 0 || fd0 != 1.f) {
            set_error("filtered_lrelu: pointwise path needs zero padding and unit down filter");
            return AFCM_ERR_UNSUPPORTED;
        }
        ActParams a;
        memset(&a, 0, sizeof(a));
        a.x = x; a.y = y; a.b = b; a.skip = skip;
        a.so = sign_mode == AFCM_SIGN_WRITE ? (uint8_t*)signs : nullptr;
        a.si = sign_mode == AFCM_SIGN_READ ? (const uint8_t*)signs : nullptr;
        a.xs_n = xs[0]; a.xs_c = xs[1]; a.xs_h = xs[2]; a.xs_w = xs[3];
        a.ys_n = ys[0]; a.ys_c = ys[1]; a.ys_h = ys[2]; a.ys_w = ys[3];
        a.N = N; a.C = C; a.h = xh; a.w = xw; a.s_h = sign_h; a.s_wb = sign_wb; a.s_ox = sx; a.s_oy = sy;
        a.gain = gain * fu0; a.slope = slope; a.clamp = clamp; a.out_scale = out_scale;
        return dtype == AFCM_F32 ? launch_act<float>(a, sign_mode, st) : launch_act<__half>(a, sign_mode, st);
    }

    FlrParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.y = y; p.b = b; p.skip = skip;
    p.so = sign_mode == AFCM_SIGN_WRITE ? (uint8_t*)signs : nullptr;
    p.si = sign_mode == AFCM_SIGN_READ ? (const uint8_t*)signs : nullptr;
    p.xs_n = xs[0]; p.xs_c = xs[1]; p.xs_h = xs[2]; p.xs_w = xs[3];
    p.ys_n = ys[0]; p.ys_c = ys[1]; p.ys_h = ys[2]; p.ys_w = ys[3];
    p.N = N; p.C = C; p.xh = xh; p.xw = xw; p.yh = yh; p.yw = yw;
    p.px0 = px0; p.py0 = py0;
    p.s_h = sign_h; p.s_wb = sign_wb; p.s_ox = sx; p.s_oy = sy;
    p.gain = gain; p.slope = slope; p.clamp = clamp; p.out_scale = out_scale;
    // Correlation-form taps: true convolution flips the filter unless flip_filter (OPS/upfirdn2d.py:198-199).
    // The up filter carries the zero-insertion gain `up` per 1-D pass (OPS/upfirdn2d.py:196 with gain=up^2).
    for (int t = 0; t < fu_taps; t++) p.ku[t] = (fu_host ? fu_host[flip_filter ? t : fu_taps - 1 - t] : 1.f) * (float)up;
    for (int t = 0; t < fd_taps; t++) p.kd[t] = fd_host ? fd_host[flip_filter ? t : fd_taps - 1 - t] : 1.f;
    if (dtype == AFCM_F32) return dispatch_fused<float>(p, up, fu_taps, down, fd_taps, sign_mode, st);
    return dispatch_fused<__half>(p, up, fu_taps, down, fd_taps, sign_mode, st);
}

extern "C" int afcm_filtered_lrelu_act(void* x, int dtype, int64_t planes, int h, int w,
                                       float gain, float slope, float clamp,
                                       int sign_mode, void* signs, int sign_h, int sign_wb, int sx, int sy,
                                       void* stream)
{
    AFCM_CHECK_ARG(x && planes > 0 && h > 0 && w > 0, "x is empty");
    AFCM_CHECK_ARG(planes <= 0x7fffffffLL, "x is too large");
    AFCM_CHECK_ARG(dtype == AFCM_F32 || dtype == AFCM_F16, "x must be float16 or float32");
    AFCM_CHECK_ARG(sign_mode >= 0 && sign_mode <= 2, "bad sign_mode");
    AFCM_CHECK_ARG(sign_mode == AFCM_SIGN_NONE || signs, "sign tensor missing");
    if (sign_mode == AFCM_SIGN_WRITE)
        AFCM_CHECK_ARG(sign_h == h && sign_wb == (((w + 15) & ~15) >> 2) && sx == 0 && sy == 0,
                       "sign tensor must be [planes,%d,%d] with zero offsets", h, ((w + 15) & ~15) >> 2);
    ActParams a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.y = x;
    a.so = sign_mode == AFCM_SIGN_WRITE ? (uint8_t*)signs : nullptr;
    a.si = sign_mode == AFCM_SIGN_READ ? (const uint8_t*)signs : nullptr;
    a.xs_n = 0; a.xs_c = (long long)h * w; a.xs_h = w; a.xs_w = 1;
    a.ys_n = 0; a.ys_c = (long long)h * w; a.ys_h = w; a.ys_w = 1;
    a.N = 1; a.C = (int)planes; a.h = h; a.w = w; a.s_h = sign_h; a.s_wb = sign_wb; a.s_ox = sx; a.s_oy = sy;
    a.gain = gain; a.slope = slope; a.clamp = clamp; a.out_scale = 1.f;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == AFCM_F32 ? launch_act<float>(a, sign_mode, st) : launch_act<__half>(a, sign_mode, st);
}
