// upfirdn2d.cu -- generic pad / zero-insert upsample / 2-D FIR / decimate (C-ABI afcm_upfirdn2d).
// Replaces upfirdn2d_kernel_small / _large (models/networks/stylegan3/torch_utils/ops/upfirdn2d.cu:29-200).
// On the AFCM inference path this op is only reached through the generic filtered_lrelu composition;
// in training it runs the 61-tap blur (models/stylegan3_model.py:28,102-103).  One thread per output
// sample, taps in the kernel-parameter constant bank (stream-safe, no device-side filter tensor), the
// polyphase structure skips the inserted zeros.
// 1-D passes without up-sampling (every separable call of the training step: the 61-tap blur of the loss images,
// models/stylegan3_model.py:28,102-103, and the [1,3,3,1] down-sampling of the discriminator, CM/generator.py:664-690) take
// tiled kernels: the horizontal pass stages a row segment in shared memory (each input is loaded once per block instead of
// once per tap), the vertical pass walks the taps with fully coalesced row reads.
#include "afcm_common.cuh"

namespace afcm {

template <int MAXT>
struct UpfirdnParams {
    const void* x; void* y;
    long long planes;
    int xh, xw, yh, yw, fh, fw, upx, upy, downx, downy, px0, py0;
    float f[MAXT];      // correlation-form taps (already flipped if needed) times gain
};

template <typename T, int MAXT>
__global__ void __launch_bounds__(256) upfirdn2d_kernel(const __grid_constant__ UpfirdnParams<MAXT> p)
{
    const long long total = p.planes * p.yh * p.yw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % p.yw);
        const long long r = i / p.yw;
        const int oy = (int)(r % p.yh);
        const long long plane = r / p.yh;
        const T* xp = (const T*)p.x + plane * p.xh * p.xw;
        // zero-inserted coordinate of tap (ty,tx): z = o*down + t - pad0 ; contributes iff z % up == 0
        const int by = oy * p.downy - p.py0, bx = ox * p.downx - p.px0;
        int ty0 = (-by) % p.upy; if (ty0 < 0) ty0 += p.upy;
        int tx0 = (-bx) % p.upx; if (tx0 < 0) tx0 += p.upx;
        float acc = 0.f;
        for (int ty = ty0; ty < p.fh; ty += p.upy) {
            const int zy = by + ty;
            if (zy < 0) continue;
            const int iy = zy / p.upy;
            if (iy >= p.xh) break;
            for (int tx = tx0; tx < p.fw; tx += p.upx) {
                const int zx = bx + tx;
                if (zx < 0) continue;
                const int ix = zx / p.upx;
                if (ix >= p.xw) break;
                acc += p.f[ty * p.fw + tx] * (float)xp[(long long)iy * p.xw + ix];
            }
        }
        ((T*)p.y)[i] = (T)acc;
    }
}

// ---- separable 1-D passes, up = 1 -------------------------------------------------------------------------------------------
constexpr int UF1_MAXT = 128;         // taps of a 1-D pass
constexpr int UF1_TILE = 256;         // outputs per block (horizontal) / threads per block
struct Upfirdn1dParams {
    const void* x; void* y;
    long long planes;
    int xh, xw, yh, yw, taps, down, pad0;
    float f[UF1_MAXT];                // correlation-form taps times gain
};

// y[row, ox] = sum_t f[t] x[row, ox down + t - pad0]; one block = UF1_TILE consecutive outputs of one row
template <typename T>
__global__ void __launch_bounds__(UF1_TILE) upfirdn1d_h_kernel(const __grid_constant__ Upfirdn1dParams p)
{
    extern __shared__ float uf_row[];                         // (UF1_TILE - 1) down + taps inputs
    const int tiles = (p.yw + UF1_TILE - 1) / UF1_TILE;
    const long long rows = p.planes * p.xh;
    for (long long b = blockIdx.x; b < rows * tiles; b += gridDim.x) {
        const long long row = b / tiles;
        const int ox0 = (int)(b - row * tiles) * UF1_TILE;
        const T* xr = (const T*)p.x + row * p.xw;
        const int ix0 = ox0 * p.down - p.pad0;
        const int span = (UF1_TILE - 1) * p.down + p.taps;
        for (int i = threadIdx.x; i < span; i += UF1_TILE) {
            const int ix = ix0 + i;
            uf_row[i] = (ix >= 0 && ix < p.xw) ? (float)xr[ix] : 0.f;
        }
        __syncthreads();
        const int ox = ox0 + threadIdx.x;
        if (ox < p.yw) {
            const float* s = uf_row + threadIdx.x * p.down;
            float acc = 0.f;
#pragma unroll 4
            for (int t = 0; t < p.taps; t++) acc = fmaf(p.f[t], s[t], acc);
            ((T*)p.y)[row * p.yw + ox] = (T)acc;
        }
        __syncthreads();
    }
}

// y[oy, x] = sum_t f[t] x[oy down + t - pad0, x]; threads along x, 4 output rows per thread
template <typename T>
__global__ void __launch_bounds__(256) upfirdn1d_v_kernel(const __grid_constant__ Upfirdn1dParams p)
{
    const int xt = (p.xw + 255) / 256, yt = (p.yh + 3) / 4;
    const long long total = p.planes * yt * xt;
    for (long long b = blockIdx.x; b < total; b += gridDim.x) {
        const int bx = (int)(b % xt);
        const long long r = b / xt;
        const int by = (int)(r % yt);
        const long long plane = r / yt;
        const int x = bx * 256 + threadIdx.x;
        if (x >= p.xw) continue;
        const T* xp = (const T*)p.x + plane * p.xh * p.xw + x;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const int oy0 = by * 4;
        const int iy0 = oy0 * p.down - p.pad0;
        // input row iy0 + i feeds output j with tap t = i - j down
        const int nin = 3 * p.down + p.taps;
        for (int i = 0; i < nin; i++) {
            const int iy = iy0 + i;
            if (iy < 0 || iy >= p.xh) continue;
            const float v = (float)xp[(long long)iy * p.xw];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int t = i - j * p.down;
                if (t >= 0 && t < p.taps) acc[j] = fmaf(p.f[t], v, acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (oy0 + j < p.yh) ((T*)p.y)[(plane * p.yh + oy0 + j) * p.xw + x] = (T)acc[j];
    }
}

template <typename T>
static int launch_upfirdn1d(bool horizontal, const void* x, void* y, long long planes, int xh, int xw, int yh, int yw,
                            const float* f_host, int taps, int down, int pad0, int flip, float gain, cudaStream_t st)
{
    Upfirdn1dParams q;
    q.x = x; q.y = y; q.planes = planes; q.xh = xh; q.xw = xw; q.yh = yh; q.yw = yw; q.taps = taps; q.down = down; q.pad0 = pad0;
    for (int t = 0; t < taps; t++) q.f[t] = (f_host ? (flip ? f_host[t] : f_host[taps - 1 - t]) : 1.f) * gain;
    const long long cap = (long long)sm_count() * 16;
    if (horizontal) {
        long long blocks = planes * xh * ((yw + UF1_TILE - 1) / UF1_TILE);
        if (blocks > cap) blocks = cap;
        const size_t smem = ((size_t)(UF1_TILE - 1) * down + taps) * sizeof(float);
        upfirdn1d_h_kernel<T><<<(unsigned)blocks, UF1_TILE, smem, st>>>(q);
    } else {
        long long blocks = planes * ((yh + 3) / 4) * ((xw + 255) / 256);
        if (blocks > cap) blocks = cap;
        upfirdn1d_v_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(q);
    }
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

template <typename T, int MAXT>
static int launch_upfirdn(const void* x, void* y, long long planes, int xh, int xw, int yh, int yw,
                          const float* f_host, int fh, int fw, int upx, int upy, int downx, int downy,
                          int px0, int py0, int flip, float gain, cudaStream_t st)
{
    static thread_local UpfirdnParams<MAXT> q_store;          // up to 16 KB: off the stack, no allocation per launch
    UpfirdnParams<MAXT>* q = &q_store;
    q->x = x; q->y = y; q->planes = planes; q->xh = xh; q->xw = xw; q->yh = yh; q->yw = yw; q->fh = fh; q->fw = fw;
    q->upx = upx; q->upy = upy; q->downx = downx; q->downy = downy; q->px0 = px0; q->py0 = py0;
    for (int ty = 0; ty < fh; ty++)
        for (int tx = 0; tx < fw; tx++) {
            const float v = f_host ? (flip ? f_host[ty * fw + tx] : f_host[(fh - 1 - ty) * fw + (fw - 1 - tx)]) : 1.f;
            q->f[ty * fw + tx] = v * gain;
        }
    const long long total = planes * yh * yw;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    upfirdn2d_kernel<T, MAXT><<<(unsigned)blocks, 256, 0, st>>>(*q);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("upfirdn2d launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    count_launch();
    return AFCM_OK;
}

}  // namespace afcm

using namespace afcm;

extern "C" int afcm_upfirdn2d(const void* x, void* y, int dtype, int64_t planes, int xh, int xw, int yh, int yw,
                              const float* f_host, int fh, int fw,
                              int upx, int upy, int downx, int downy, int px0, int px1, int py0, int py1,
                              int flip_filter, float gain, void* stream)
{
    // argument checks of OPS/upfirdn2d.cpp:20-37
    AFCM_CHECK_ARG(x && y && planes > 0 && xh > 0 && xw > 0, "x is empty");
    AFCM_CHECK_ARG(dtype == AFCM_F32 || dtype == AFCM_F16, "x must be float16 or float32");
    AFCM_CHECK_ARG(fh >= 1 && fw >= 1, "f must be at least 1x1");
    AFCM_CHECK_ARG(upx >= 1 && upy >= 1 && downx >= 1 && downy >= 1, "upsampling and downsampling factors must be at least 1");
    const long long ew = ((long long)xw * upx + px0 + px1 - fw + downx) / downx;
    const long long eh = ((long long)xh * upy + py0 + py1 - fh + downy) / downy;
    AFCM_CHECK_ARG(ew >= 1 && eh >= 1, "output must be at least 1x1");
    AFCM_CHECK_ARG(ew == yw && eh == yh, "y has shape [%d,%d], expected [%lld,%lld]", yh, yw, eh, ew);
    cudaStream_t st = (cudaStream_t)stream;
    const int taps = fh * fw;
    // separable 1-D pass without up-sampling: tiled kernels (down <= 4 keeps the shared-memory row segment under 5 KB)
    if (upx == 1 && upy == 1 && taps <= UF1_MAXT) {
        if (fh == 1 && downy == 1 && py0 == 0 && py1 == 0 && downx <= 4) {
            if (dtype == AFCM_F32) return launch_upfirdn1d<float>(true, x, y, planes, xh, xw, yh, yw, f_host, fw, downx, px0, flip_filter, gain, st);
            return launch_upfirdn1d<__half>(true, x, y, planes, xh, xw, yh, yw, f_host, fw, downx, px0, flip_filter, gain, st);
        }
        if (fw == 1 && fh > 1 && downx == 1 && px0 == 0 && px1 == 0 && downy <= 4) {
            if (dtype == AFCM_F32) return launch_upfirdn1d<float>(false, x, y, planes, xh, xw, yh, yw, f_host, fh, downy, py0, flip_filter, gain, st);
            return launch_upfirdn1d<__half>(false, x, y, planes, xh, xw, yh, yw, f_host, fh, downy, py0, flip_filter, gain, st);
        }
    }
#define AFCM_UPF(T, M) return launch_upfirdn<T, M>(x, y, planes, xh, xw, yh, yw, f_host, fh, fw, upx, upy, downx, downy, px0, py0, flip_filter, gain, st)
    if (dtype == AFCM_F32) {
        if (taps <= 64) AFCM_UPF(float, 64);
        if (taps <= 1024) AFCM_UPF(float, 1024);
        if (taps <= 4096) AFCM_UPF(float, 4096);
    } else {
        if (taps <= 64) AFCM_UPF(__half, 64);
        if (taps <= 1024) AFCM_UPF(__half, 1024);
        if (taps <= 4096) AFCM_UPF(__half, 4096);
    }
#undef AFCM_UPF
    set_error("upfirdn2d: filters with more than 4096 taps are not supported");
    return AFCM_ERR_UNSUPPORTED;
}
