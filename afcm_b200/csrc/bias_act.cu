// bias_act.cu -- fused bias + activation + gain + clamp with first/second order gradients
// (C-ABI afcm_bias_act).  Replaces bias_act_kernel (models/networks/stylegan3/torch_utils/ops/
// bias_act.cu:23-147); the nine activations and their index follow OPS/bias_act.py:21-31.
#include "afcm_common.cuh"

namespace afcm {

struct BiasActParams {
    const void* x; const void* b; const void* xref; const void* yref; const void* dy; void* y;
    long long n, step_b, size_b;
    int grad, act;
    float alpha, gain, clamp;
};

// d/dx and d2/dx2 of the activation expressed through the saved forward output yy (= y / gain) where
// the reference does so ('ref' column of the activation table), or through xref for swish.
__device__ __forceinline__ float act_d1(int act, float yy, float xr, float alpha)
{
    const float ss = 1.0507009873554804934193349852946f, sa = 1.6732632423543772848170429916717f;
    switch (act) {
    case 1: return 1.f;
    case 2: return yy > 0.f ? 1.f : 0.f;
    case 3: return yy > 0.f ? 1.f : alpha;
    case 4: return 1.f - yy * yy;
    case 5: return yy * (1.f - yy);
    case 6: return yy >= 0.f ? 1.f : yy + 1.f;
    case 7: return yy >= 0.f ? ss : yy + ss * sa;
    case 8: return 1.f - expf(-yy);
    default: {
        if (xr > 40.f) return 1.f;
        const float c = expf(xr), d = c + 1.f;
        return c * (xr + d) / (d * d);
    }
    }
}

__device__ __forceinline__ float act_d2(int act, float yy, float xr)
{
    const float ss = 1.0507009873554804934193349852946f, sa = 1.6732632423543772848170429916717f;
    switch (act) {
    case 4: return (1.f - yy * yy) * (-2.f * yy);
    case 5: return yy * (1.f - yy) * (1.f - 2.f * yy);
    case 6: return yy >= 0.f ? 0.f : yy + 1.f;
    case 7: return yy >= 0.f ? 0.f : yy + ss * sa;
    case 8: { const float c = expf(-yy); return c * (1.f - c); }
    case 9: {
        if (xr > 40.f) return 0.f;
        const float c = expf(xr), d = c + 1.f;
        return c * (xr * (2.f - d) + 2.f * d) / (d * d * d);
    }
    default: return 0.f;      // linear, relu, lrelu have no second derivative
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bias_act_kernel(const __grid_constant__ BiasActParams p)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
        float x = (float)((const T*)p.x)[i];
        const float b = p.b ? (float)((const T*)p.b)[(i / p.step_b) % p.size_b] : 0.f;
        float y;
        if (p.grad == 0) {
            y = act_eval(x + b, p.act, p.alpha) * p.gain;
            if (p.clamp >= 0.f) y = fminf(fmaxf(y, -p.clamp), p.clamp);
        } else {
            const float xr = (p.xref ? (float)((const T*)p.xref)[i] : 0.f) + b;
            float yr = p.yref ? (float)((const T*)p.yref)[i] : 0.f;
            const float dy = p.dy ? (float)((const T*)p.dy)[i] : 1.f;
            const float yy = p.gain != 0.f ? yr / p.gain : 0.f;
            const float d = p.grad == 1 ? act_d1(p.act, yy, xr, p.alpha) : act_d2(p.act, yy, xr);
            y = x * d * p.gain * dy;
            if (p.act == 9) yr = act_eval(xr, 9, 0.f) * p.gain;      // swish saves x, not y
            if (p.clamp >= 0.f && !(yr > -p.clamp && yr < p.clamp)) y = 0.f;
        }
        ((T*)p.y)[i] = (T)y;
    }
}

}  // namespace afcm

using namespace afcm;

extern "C" int afcm_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy,
                             void* y, int dtype, int64_t n, int64_t step_b, int64_t size_b,
                             int grad, int act, float alpha, float gain, float clamp, void* stream)
{
    AFCM_CHECK_ARG(x && y, "x and y must be given");
    AFCM_CHECK_ARG(n > 0, "x is empty");
    AFCM_CHECK_ARG(dtype == AFCM_F32 || dtype == AFCM_F16, "x must be float16 or float32");
    AFCM_CHECK_ARG(grad >= 0 && grad <= 2, "grad must be 0, 1 or 2");
    AFCM_CHECK_ARG(act >= 1 && act <= 9, "unknown activation index %d", act);
    AFCM_CHECK_ARG(!b || (step_b > 0 && size_b > 0), "bad bias indexing");
    BiasActParams p = {x, b, xref, yref, dy, y, n, step_b > 0 ? step_b : 1, size_b > 0 ? size_b : 1, grad, act, alpha, gain, clamp};
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AFCM_F32) bias_act_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(p);
    else bias_act_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>(p);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}
