// conv2d_bwd.cu -- gradients of the shared convolution formulation (include/afcm_b200.h)
//
//     y[n,o] = ocoef[n,o] * sum_{i,ky,kx} w[o,i,ky,kx] * (icoef[n,i] * x[n,i])        (stride 1, zero padding)
//
// i.e. the backward pass of modulated_conv2d (reference models/networks/stylegan3/networks_stylegan3.py:25-64),
// the encoder convolution (:503-505) and Conv2dLayer (models/networks/CoModGAN/layers.py:157), which the
// reference obtains from cuDNN through autograd of F.conv2d (torch_utils/ops/conv2d_gradfix.py:37-40).
//
//   data gradient    dxm = conv(ocoef * dy, flip(w)^T, pad' = k-1-pad)  -> the FORWARD kernels (tcgen05 / fp32) with
//                    transposed, flipped weights; this file adds the per-plane reductions around it:
//                    d_icoef[n,i] = <dxm[n,i], x[n,i]>,  dx = icoef * dxm,  d_ocoef[n,o] = <dy[n,o], y[n,o]> / ocoef[n,o]
//   weight gradient  dw[o,i,ky,kx] = sum_{n,p} (ocoef dy)[n,o,p] * (icoef x)[n,i,p + (ky-pad, kx-pad)]
//                    * wgrad_f32_kernel : exact fp32 SIMT (the parity path)
//                    * wgrad_tc_kernel  : mma.sync m16n8k16 (bf16 / fp16 operands, fp32 accumulation) straight from the
//                      16-bit channel-innermost "flat plane" tensors the tcgen05 forward / data-gradient kernels consume
//                      (afcm_conv_tc_pack): a tap is a ROW shift of the shared-memory tile, so the nine taps read the
//                      same staged tile through ldmatrix.trans at different row offsets -- no im2col, no conversion.
//   afcm_adam_step   fused Adam update with the reference's gradient scrub (train.py:67-77 nan_to_num).
#include "afcm_common.cuh"

namespace afcm {

// ---- per-plane reductions ----------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float s, float* red)
{
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) t += red[k];
    return t;
}

// out[p] = <a[p,:], b[p,:]> * (div ? 1/div[p] : 1);  then (coef != null) a[p,:] *= coef[p] in place.
__global__ void __launch_bounds__(256)
plane_dot_scale_kernel(float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ div,
                       const float* __restrict__ coef, float* __restrict__ out, long long hw)
{
    __shared__ float red[8];
    const long long p = blockIdx.x;
    float* ap = a + p * hw;
    if (out) {
        const float* bp = b + p * hw;
        float s = 0.f;
        for (long long j = threadIdx.x; j < hw; j += 256) s = fmaf(ap[j], bp[j], s);
        const float t = block_sum_256(s, red);
        if (threadIdx.x == 0) out[p] = div ? t / div[p] : t;
    }
    if (coef) {
        const float c = coef[p];
        for (long long j = threadIdx.x; j < hw; j += 256) ap[j] *= c;
    }
}

// out[p] = sum_j a[p,j]  (the per-plane part of the bias gradient db = sum_{n,h,w} dx of filtered_lrelu / bias_act)
__global__ void __launch_bounds__(256)
plane_sum_kernel(const float* __restrict__ a, float* __restrict__ out, long long hw)
{
    __shared__ float red[8];
    const float* ap = a + (long long)blockIdx.x * hw;
    float s = 0.f;
    for (long long j = threadIdx.x; j < hw; j += 256) s += ap[j];
    const float t = block_sum_256(s, red);
    if (threadIdx.x == 0) out[blockIdx.x] = t;
}

// ---- exact fp32 weight gradient ------------------------------------------------------------------------
constexpr int WG_TO = 16, WG_TI = 16, WG_KB = 32;

template <int KS>
__global__ void __launch_bounds__(256)
wgrad_f32_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ icoef,
                 const float* __restrict__ ocoef, float* __restrict__ dw,
                 int N, int Ci, int H, int W, int Co, int pad, int OH, int OW, int nchunk, long long steps)
{
    constexpr int BP = WG_KB + KS - 1 + (KS == 1 ? 1 : 0);
    __shared__ float As[WG_TO][WG_KB];
    __shared__ float Bs[WG_TI][KS][BP];
    const int tid = threadIdx.x, to = tid >> 4, ti = tid & 15;
    const int o0 = blockIdx.x * WG_TO, i0 = blockIdx.y * WG_TI;
    const long long s0 = steps * blockIdx.z / gridDim.z, s1 = steps * (blockIdx.z + 1) / gridDim.z;
    float acc[KS * KS];
#pragma unroll
    for (int t = 0; t < KS * KS; t++) acc[t] = 0.f;
    for (long long s = s0; s < s1; s++) {
        const int ch = (int)(s % nchunk);
        const long long r = s / nchunk;
        const int oy = (int)(r % OH), n = (int)(r / OH);
        const int x0 = ch * WG_KB;
        for (int idx = tid; idx < WG_TO * WG_KB; idx += 256) {
            const int oo = idx / WG_KB, k = idx % WG_KB, o = o0 + oo, ox = x0 + k;
            float v = 0.f;
            if (o < Co && ox < OW) {
                v = dy[(((long long)n * Co + o) * OH + oy) * OW + ox];
                if (ocoef) v *= ocoef[(long long)n * Co + o];
            }
            As[oo][k] = v;
        }
        for (int idx = tid; idx < WG_TI * KS * (WG_KB + KS - 1); idx += 256) {
            const int j = idx % (WG_KB + KS - 1), rr = idx / (WG_KB + KS - 1), ky = rr % KS, ii = rr / KS;
            const int i = i0 + ii, iy = oy + ky - pad, ix = x0 + j - pad;
            float v = 0.f;
            if (i < Ci && iy >= 0 && iy < H && ix >= 0 && ix < W) {
                v = x[(((long long)n * Ci + i) * H + iy) * W + ix];
                if (icoef) v *= icoef[(long long)n * Ci + i];
            }
            Bs[ii][ky][j] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < WG_KB; k++) {
            const float a = As[to][k];
#pragma unroll
            for (int ky = 0; ky < KS; ky++)
#pragma unroll
                for (int kx = 0; kx < KS; kx++) acc[ky * KS + kx] = fmaf(a, Bs[ti][ky][k + kx], acc[ky * KS + kx]);
        }
        __syncthreads();
    }
    const int o = o0 + to, i = i0 + ti;
    if (o < Co && i < Ci && s1 > s0) {
        float* q = dw + ((long long)o * Ci + i) * (KS * KS);
#pragma unroll
        for (int t = 0; t < KS * KS; t++) atomicAdd(q + t, acc[t]);
    }
}

// ---- tensor-core weight gradient (3x3) -------------------------------------------------------------------
// GEMM per tap:  dW_t[o, i] = sum_k A[k, o] * B_t[k, i],  k = output pixels of one row segment.
//   A tile  = dyp[n][oy*(OW+2) + x0 .. +64][o0 .. +64]         (16-bit, channel innermost)    -> smem [64 px][72]
//   B tiles = xp [n][(oy+ky-pad)*(W+2) + x0-pad .. +66][i0..+64] for ky = 0..2              -> smem [3][66 px][72]
// Tap (ky,kx) pairs A row k with B row (ky, k + kx).  Both operands are stored [k][channel]; ldmatrix.trans delivers
// the m16n8k16 row/col fragments.  8 warps = 2 (o) x 4 (i), warp tile 32 o x 16 i x 9 taps = 144 fp32 accumulators.
// cp.async 16-byte copies with zero fill (masked pixels / rows off the plane), 4-stage ring, split-K over CTAs with
// fp32 atomics into dw.
constexpr int WT_BM = 64, WT_BN = 64, WT_KB = 64, WT_PITCH = 72, WT_STAGES = 4;
constexpr int WT_BPX = WT_KB + 2;
constexpr int WT_A_ELEMS = WT_KB * WT_PITCH;
constexpr int WT_B_ELEMS = 3 * WT_BPX * WT_PITCH;
constexpr int WT_STAGE_ELEMS = WT_A_ELEMS + WT_B_ELEMS;
constexpr int WT_SMEM_BYTES = WT_STAGES * WT_STAGE_ELEMS * 2;

__device__ __forceinline__ void cp_async16_zfill(void* dst, const void* src, bool valid)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
template <bool BF16>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    if (BF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct WtParams {
    const uint16_t* dyp;    // [N][OH*(OW+2)][co_pad]
    const uint16_t* xp;     // [N][H*(W+2)][ci_pad]
    float* dw;              // [Co][Ci][3][3]
    int N, Ci, Co, H, W, OH, OW, pad, co_pad, ci_pad, nchunk, tiles_i;
    long long steps;
};

template <bool BF16>
__global__ void __launch_bounds__(256, 1)
wgrad_tc_kernel(const WtParams p)
{
    extern __shared__ __align__(16) uint16_t wt_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int o0 = (blockIdx.x / p.tiles_i) * WT_BM, i0 = (blockIdx.x % p.tiles_i) * WT_BN;
    const long long s0 = p.steps * blockIdx.y / gridDim.y, s1 = p.steps * (blockIdx.y + 1) / gridDim.y;
    const int nsteps = (int)(s1 - s0);
    const long long rowsA = (long long)p.OH * (p.OW + 2), rowsB = (long long)p.H * (p.W + 2);

    float acc[9][2][2][4];
#pragma unroll
    for (int t = 0; t < 9; t++)
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[t][a][b][c] = 0.f;

    // Per-thread copy pattern, fixed for the whole kernel: 16-byte channel chunk c8 = tid & 7, pixel slots tid >> 3 (+32, +64).
    // Only the tile origin changes from stage to stage, so the address arithmetic per cp.async is one add.
    const int c8 = (tid & 7) * 8, j0 = tid >> 3;
    const bool a_ch_ok = o0 + c8 < p.co_pad, b_ch_ok = i0 + c8 < p.ci_pad;
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(wt_smem);
    const uint32_t a_dst0 = (uint32_t)((j0 * WT_PITCH + c8) * 2);
    const long long a_src0 = (long long)j0 * p.co_pad + o0 + c8;
    const long long b_rowstep = (long long)(p.W + 2) * p.ci_pad;
    auto cp16 = [&](uint32_t dst, const uint16_t* src, bool ok) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
    };
    auto load_stage = [&](int it) {
        const long long s = s0 + it;
        const int ch = (int)(s % p.nchunk);
        const long long r = s / p.nchunk;
        const int oy = (int)(r % p.OH), n = (int)(r / p.OH);
        const int x0 = ch * WT_KB;
        const uint32_t sA = smem0 + (uint32_t)((it % WT_STAGES) * WT_STAGE_ELEMS * 2), sB = sA + WT_A_ELEMS * 2;
        const uint16_t* an = p.dyp + ((long long)n * rowsA + (long long)oy * (p.OW + 2) + x0) * p.co_pad + a_src0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const bool ok = a_ch_ok && (x0 + j0 + h * 32 < p.OW);
            cp16(sA + a_dst0 + (uint32_t)(h * 32 * WT_PITCH * 2), ok ? an + (long long)h * 32 * p.co_pad : p.dyp, ok);
        }
        // B: rows iy = oy + ky - pad, pixels xi = x0 - pad + j, j = j0 + 32 h (h = 2 only for j0 < 2)
        const int xi0 = x0 - p.pad + j0;
        const long long flat0 = (long long)(oy - p.pad) * (p.W + 2) + xi0;
        const uint16_t* bn = p.xp + ((long long)n * rowsB + flat0) * p.ci_pad + i0 + c8;
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
            const int iy = oy + ky - p.pad;
            const bool row_ok = b_ch_ok && iy >= 0 && iy < p.H;
            const uint16_t* br = bn + ky * b_rowstep;
            const uint32_t d = sB + (uint32_t)(((ky * WT_BPX + j0) * WT_PITCH + c8) * 2);
#pragma unroll
            for (int h = 0; h < 3; h++) {
                if (h == 2 && j0 >= WT_BPX - 64) break;
                const int xi = xi0 + h * 32;
                const bool ok = row_ok && (flat0 + ky * (p.W + 2) + h * 32 >= 0) && xi < p.W + 2;
                cp16(d + (uint32_t)(h * 32 * WT_PITCH * 2), ok ? br + (long long)h * 32 * p.ci_pad : p.xp, ok);
            }
        }
    };

    for (int it = 0; it < WT_STAGES - 1; it++) {
        if (it < nsteps) load_stage(it);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int lj = lane >> 3, lr = lane & 7;
    for (int it = 0; it < nsteps; it++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(WT_STAGES - 2) : "memory");
        __syncthreads();
        if (it + WT_STAGES - 1 < nsteps) load_stage(it + WT_STAGES - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");

        const uint16_t* As = wt_smem + (size_t)(it % WT_STAGES) * WT_STAGE_ELEMS;
        const uint16_t* Bs = As + WT_A_ELEMS;
        const int ch = (int)((s0 + it) % p.nchunk);
        const int valid = min(WT_KB, p.OW - ch * WT_KB);
        const int nk = (valid + 15) >> 4;
        for (int kk = 0; kk < nk; kk++) {
            uint32_t af[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
                ldmatrix_x4_trans(af[mt], As + (kk * 16 + (lj >> 1) * 8 + lr) * WT_PITCH + wm * 32 + mt * 16 + (lj & 1) * 8);
#pragma unroll
            for (int ky = 0; ky < 3; ky++)
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    uint32_t bf[4];
                    ldmatrix_x4_trans(bf, Bs + (ky * WT_BPX + kk * 16 + kx + (lj & 1) * 8 + lr) * WT_PITCH + wn * 16 + (lj >> 1) * 8);
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        mma_16816<BF16>(acc[ky * 3 + kx][mt][0], af[mt], bf[0], bf[1]);
                        mma_16816<BF16>(acc[ky * 3 + kx][mt][1], af[mt], bf[2], bf[3]);
                    }
                }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (nsteps <= 0) return;
    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int t = 0; t < 9; t++)
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int o = o0 + wm * 32 + mt * 16 + g + (c >> 1) * 8;
                    const int i = i0 + wn * 16 + nt * 8 + t4 * 2 + (c & 1);
                    if (o < p.Co && i < p.Ci) atomicAdd(p.dw + ((long long)o * p.Ci + i) * 9 + t, acc[t][mt][nt][c]);
                }
}

// ---- fused Adam -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            float lr, float b1, float b2, float eps, float c1, float c2, float gscale, int scrub)
{
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        float gi = g[i] * gscale;
        if (scrub) {                                   // train.py:67-77  nan_to_num(nan=0, posinf=1e5, neginf=-1e5)
            if (gi != gi) gi = 0.f;
            else if (gi > 3.0e38f) gi = 1e5f;
            else if (gi < -3.0e38f) gi = -1e5f;
        }
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= lr * (mi / c1) / (sqrtf(vi) / c2 + eps);
    }
}

// ---- generator EMA (train.py:74-75): p_ema = lerp(p, p_ema, beta) = p + beta (p_ema - p), over the flat parameter buffers ------
__global__ void __launch_bounds__(256)
ema_lerp_kernel(float* __restrict__ p_ema, const float* __restrict__ p, long long n, float beta)
{
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float a = p[i];
        p_ema[i] = fmaf(beta, p_ema[i] - a, a);
    }
}

}  // namespace afcm

using namespace afcm;

extern "C" int afcm_ema_lerp(float* param_ema, const float* param, int64_t n, float beta, void* stream)
{
    AFCM_CHECK_ARG(param_ema && param && n > 0, "empty problem");
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    ema_lerp_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param_ema, param, (long long)n, beta);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_plane_dot_scale(float* a, const float* b, const float* div, const float* coef, float* out,
                                    int64_t planes, int64_t hw, void* stream)
{
    AFCM_CHECK_ARG(a && planes > 0 && hw > 0, "empty problem");
    AFCM_CHECK_ARG(!out || b, "the dot product needs b");
    AFCM_CHECK_ARG(planes <= 0x7fffffffLL, "too many planes");
    if (!out && !coef) return AFCM_OK;
    plane_dot_scale_kernel<<<(unsigned)planes, 256, 0, (cudaStream_t)stream>>>(a, b, div, coef, out, (long long)hw);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_plane_sum(const float* a, float* out, int64_t planes, int64_t hw, void* stream)
{
    AFCM_CHECK_ARG(a && out && planes > 0 && hw > 0, "empty problem");
    AFCM_CHECK_ARG(planes <= 0x7fffffffLL, "too many planes");
    plane_sum_kernel<<<(unsigned)planes, 256, 0, (cudaStream_t)stream>>>(a, out, (long long)hw);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_conv2d_wgrad_f32(const float* dy, const float* x, const float* icoef, const float* ocoef, float* dw,
                                     int N, int Ci, int H, int W, int Co, int ksize, int pad, void* stream)
{
    AFCM_CHECK_ARG(dy && x && dw, "dy, x and dw must be given");
    AFCM_CHECK_ARG(N > 0 && Ci > 0 && Co > 0 && H > 0 && W > 0 && pad >= 0, "empty problem");
    if (ksize != 1 && ksize != 3) { set_error("conv2d_wgrad_f32: kernel size %d not supported (1 or 3)", ksize); return AFCM_ERR_UNSUPPORTED; }
    const int OH = H + 2 * pad - ksize + 1, OW = W + 2 * pad - ksize + 1;
    AFCM_CHECK_ARG(OH > 0 && OW > 0, "output must be at least 1x1");
    cudaStream_t st = (cudaStream_t)stream;
    AFCM_CUDA(cudaMemsetAsync(dw, 0, (size_t)Co * Ci * ksize * ksize * sizeof(float), st));
    const int nchunk = ceil_div(OW, WG_KB);
    const long long steps = (long long)N * OH * nchunk;
    const int tiles = ceil_div(Co, WG_TO) * ceil_div(Ci, WG_TI);
    long long splits = (long long)sm_count() * 8 / tiles;
    if (splits < 1) splits = 1;
    if (splits > steps) splits = steps;
    if (splits > 65535) splits = 65535;
    dim3 grid(ceil_div(Co, WG_TO), ceil_div(Ci, WG_TI), (unsigned)splits);
    AFCM_CHECK_ARG(grid.y <= 65535, "too many input channels");
    if (ksize == 3) wgrad_f32_kernel<3><<<grid, 256, 0, st>>>(dy, x, icoef, ocoef, dw, N, Ci, H, W, Co, pad, OH, OW, nchunk, steps);
    else wgrad_f32_kernel<1><<<grid, 256, 0, st>>>(dy, x, icoef, ocoef, dw, N, Ci, H, W, Co, pad, OH, OW, nchunk, steps);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_conv2d_wgrad_tc(const void* dyp, const void* xp, float* dw, int tc_dtype,
                                    int N, int Ci, int H, int W, int Co, int pad, void* stream)
{
    AFCM_CHECK_ARG(dyp && xp && dw, "dyp, xp and dw must be given");
    AFCM_CHECK_ARG(N > 0 && Ci > 0 && Co > 0 && H > 0 && W > 0, "empty problem");
    AFCM_CHECK_ARG(tc_dtype == AFCM_F16 || tc_dtype == AFCM_BF16, "tc dtype must be F16 or BF16");
    if (pad < 0 || pad > 2) { set_error("conv2d_wgrad_tc: padding %d not supported (0..2)", pad); return AFCM_ERR_UNSUPPORTED; }
    WtParams p;
    p.dyp = (const uint16_t*)dyp; p.xp = (const uint16_t*)xp; p.dw = dw;
    p.N = N; p.Ci = Ci; p.Co = Co; p.H = H; p.W = W; p.pad = pad;
    p.OH = H + 2 * pad - 2; p.OW = W + 2 * pad - 2;
    AFCM_CHECK_ARG(p.OH > 0 && p.OW > 0, "output must be at least 1x1");
    p.co_pad = (Co + 7) & ~7; p.ci_pad = (Ci + 7) & ~7;
    p.nchunk = ceil_div(p.OW, WT_KB);
    p.steps = (long long)N * p.OH * p.nchunk;
    p.tiles_i = ceil_div(Ci, WT_BN);
    const int tiles = ceil_div(Co, WT_BM) * p.tiles_i;
    // split the pixel reduction so that the grid is a whole number of waves of one CTA per SM
    const int sms = sm_count();
    long long splits = tiles >= sms ? 1 : (2LL * sms) / tiles;
    if (splits < 1) splits = 1;
    if (splits > p.steps) splits = p.steps;
    if (splits > 65535) splits = 65535;
    cudaStream_t st = (cudaStream_t)stream;
    AFCM_CUDA(cudaMemsetAsync(dw, 0, (size_t)Co * Ci * 9 * sizeof(float), st));
    dim3 grid((unsigned)tiles, (unsigned)splits);
    if (tc_dtype == AFCM_BF16) {
        AFCM_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM_BYTES));
        wgrad_tc_kernel<true><<<grid, 256, WT_SMEM_BYTES, st>>>(p);
    } else {
        AFCM_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM_BYTES));
        wgrad_tc_kernel<false><<<grid, 256, WT_SMEM_BYTES, st>>>(p);
    }
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                              float lr, float beta1, float beta2, float eps, int step, float grad_scale, int scrub,
                              void* stream)
{
    AFCM_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "bad arguments");
    const double c1 = 1.0 - pow((double)beta1, (double)step), c2 = sqrt(1.0 - pow((double)beta2, (double)step));
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, (long long)n, lr, beta1, beta2, eps,
                                                                   (float)c1, (float)c2, grad_scale, scrub);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}
