// upfirdn2d.cu -- generic pad / zero-insert upsample / 2-D FIR / decimate (C-ABI afcm_upfirdn2d).
// Replaces upfirdn2d_kernel_small / _large (models/networks/stylegan3/torch_utils/ops/upfirdn2d.cu:29-200).
// On the AFCM inference path this op is only reached through the generic filtered_lrelu composition;
// in training it runs the 61-tap blur (models/stylegan3_model.py:28,102-103).  One thread per output
// sample, taps in the kernel-parameter constant bank (stream-safe, no device-side filter tensor), the
// polyphase structure skips the inserted zeros.
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

template <typename T, int MAXT>
static int launch_upfirdn(const void* x, void* y, long long planes, int xh, int xw, int yh, int yw,
                          const float* f_host, int fh, int fw, int upx, int upy, int downx, int downy,
                          int px0, int py0, int flip, float gain, cudaStream_t st)
{
    UpfirdnParams<MAXT>* q = new UpfirdnParams<MAXT>();      // up to 16 KB: keep it off the stack
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
    delete q;
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
