// conv2d_wgrad_tc5.cu -- weight gradient of the 3x3 convolution on tcgen05 / TMEM (sm_100a).
//
//     dw[o,i,ky,kx] = sum_{n,oy,ox} dyp[n,(oy,ox),o] * xp[n,(oy+ky-pad, ox+kx-pad),i]
//
// (the cuDNN wgrad the reference reaches through autograd of F.conv2d, torch_utils/ops/conv2d_gradfix.py:37-40, inside
// modulated_conv2d networks_stylegan3.py:60-63 and the encoder conv :503-505).  dyp / xp are the 16-bit channel-innermost
// flat planes of afcm_conv_tc_pack (row pitch = width + 2) with the per-sample coefficients already folded in, so this is
// ONE GEMM over the whole batch with the PIXELS as the reduction dimension:
//
//     D_kx[o, i] += A[k, o]^T * B[k + kx, i]         k = 64 output pixels of one image row, per kernel row ky
//
//   * both operands are "MN-major" for the tensor core (the contiguous shared-memory dimension is the channel, not the
//     reduction index): TMA writes [pixel rows][64 channels = 128 B] tiles with SWIZZLE_128B, the UMMA descriptors address
//     them with LBO = distance between 64-channel column blocks, SBO = 1024 B (8 pixel rows), and a K step of 16 pixels is
//     16 rows = 2048 B further on.  A horizontal tap kx is the SAME B tile read one pixel row (128 B) later, so one staged
//     tile of 72 pixel rows serves the three taps of a kernel row (the swizzle is a function of the address bits only).
//   * masking is done by TMA: A comes from a [channel, x, row] view of dyp (x beyond the row is zero-filled), B from a
//     [channel, x, y, n] view of xp (negative / overflowing x and y are zero-filled), so rows of one image never mix.
//   * CTA = (128 output channels) x (64 or 128 input channels) x (one kernel row ky) x (one slice of the pixel range);
//     three fp32 accumulators (kx = 0,1,2) of 128 x BN live in tensor memory (3 BN <= 384 columns).
//   * warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2-5 = epilogue (tcgen05.ld -> partial sums).  Partials go
//     to a caller-provided workspace [split][ky,kx][Co][Ci] with plain coalesced stores and a small second kernel reduces
//     them in a fixed order into dw [Co,Ci,3,3] -- deterministic, no atomics.
#include "afcm_common.cuh"
#include "tc_ptx.cuh"

namespace afcm {

constexpr int W5_THREADS = 192;
constexpr int W5_KB = 64;                          // output pixels per stage
constexpr int W5_BROWS = 72;                       // B pixel rows per stage: 64 + 2 (taps), rounded to 8
constexpr int W5_A_BLK = W5_KB * 128;              // one 64-channel column block of A: 8 KB
constexpr int W5_B_BLK = W5_BROWS * 128;           // one 64-channel column block of B: 9 KB
constexpr int W5_STAGE_BYTES = 2 * W5_A_BLK + 2 * W5_B_BLK;   // 34 KB (1024-byte multiple)
constexpr int W5_MAX_STAGES = 6;
constexpr int W5_TMEM_COLS = 512;

struct W5Params {
    float* ws;                  // [splits][9][Co][Ci]
    int N, Ci, Co, H, W, OH, OW, pad;
    int co_pad, ci_pad;         // channel counts of the packed planes (multiples of 8)
    int BN, nblk_a, tiles_i, tiles_o, nchunk, stages;
    long long steps;            // N * OH * nchunk
    unsigned idesc;
};

// low descriptor word for an MN-major SWIZZLE_128B operand: start address | leading byte offset (between column blocks)
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t saddr, uint32_t lbo_bytes)
{
    return ((saddr >> 4) & 0x3fff) | (((lbo_bytes >> 4) & 0x3fff) << 16);
}

__global__ void __launch_bounds__(W5_THREADS, 1)
wgrad_tc5_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ W5Params p)
{
    extern __shared__ uint8_t w5_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(w5_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* tail = smem + p.stages * W5_STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);      // [W5_MAX_STAGES]
    uint64_t* empty = full + W5_MAX_STAGES;                  // [W5_MAX_STAGES]
    uint64_t* done = empty + W5_MAX_STAGES;                  // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x % p.tiles_i;
    const int r0 = blockIdx.x / p.tiles_i;
    const int to = r0 % p.tiles_o, ky = r0 / p.tiles_o;
    const int o0 = to * 128, i0 = ti * p.BN;
    const long long s0 = p.steps * blockIdx.y / gridDim.y, s1 = p.steps * (blockIdx.y + 1) / gridDim.y;
    const int nsteps = (int)(s1 - s0);
    const int nblk_b = p.BN >> 6;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(W5_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // which 64-channel column blocks exist (a block that lies wholly beyond the channel count is neither loaded nor
    // stored: its accumulator rows / columns are garbage that nobody reads)
    const int na = (o0 + 64 < p.co_pad) ? 2 : 1;
    const int nb = (nblk_b == 2 && i0 + 64 < p.ci_pad) ? 2 : 1;
    const uint32_t tx_bytes = (uint32_t)(na * W5_A_BLK + nb * W5_B_BLK);

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < nsteps; it++) {
                const long long s = s0 + it;
                const int ch = (int)(s % p.nchunk);
                const long long r = s / p.nchunk;
                const int oy = (int)(r % p.OH), n = (int)(r / p.OH);
                const int x0 = ch * W5_KB;
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], tx_bytes);
                uint8_t* sa = smem + stage * W5_STAGE_BYTES;
                uint8_t* sb = sa + 2 * W5_A_BLK;
                for (int b = 0; b < na; b++) tma_load_3d(sa + b * W5_A_BLK, &map_a, &full[stage], o0 + 64 * b, x0, (int)r);
                for (int b = 0; b < nb; b++)
                    tma_load_4d(sb + b * W5_B_BLK, &map_b, &full[stage], i0 + 64 * b, x0 - p.pad, oy + ky - p.pad, n);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        int stage = 0; uint32_t phase = 0;
        const uint32_t smem_base = smem_u32(smem);
        for (int it = 0; it < nsteps; it++) {
            const int ch = (int)((s0 + it) % p.nchunk);
            const int valid = min(W5_KB, p.OW - ch * W5_KB);
            const int nk = (valid + 15) >> 4;
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_base + (uint32_t)(stage * W5_STAGE_BYTES);
            const uint32_t a_lo = desc_lo_mn(sa, W5_A_BLK), b_lo = desc_lo_mn(sa + 2 * W5_A_BLK, W5_B_BLK);
            if (elect_one()) {
                for (int kk = 0; kk < nk; kk++) {
#pragma unroll
                    for (int kx = 0; kx < 3; kx++)
                        // 16 pixel rows = 2048 B = 128 descriptor units; one pixel row (tap) = 128 B = 8 units
                        umma_f16_lohi(tmem_base + (uint32_t)(kx * p.BN), a_lo + (uint32_t)(kk * 128),
                                      b_lo + (uint32_t)(kk * 128 + kx * 8), TC_DESC_HI, p.idesc, (it | kk) != 0);
                }
                umma_commit(&empty[stage]);
                if (it == nsteps - 1) umma_commit(done);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    } else {
        // ================= epilogue: warps 2..5, TMEM lane quadrant = warp & 3 =================
        const int wq = warp & 3;
        mbar_wait(done, 0);
        tc_fence_after();
        const int o = o0 + wq * 32 + lane;
        const bool vec = (p.Ci & 3) == 0;
#pragma unroll 1
        for (int kx = 0; kx < 3; kx++) {
            float* wrow = p.ws + (((long long)blockIdx.y * 9 + ky * 3 + kx) * p.Co + o) * p.Ci;
#pragma unroll 1
            for (int c0 = 0; c0 < p.BN; c0 += 32) {
                if (i0 + c0 >= p.Ci) break;                   // warp-uniform
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(kx * p.BN + c0), v);
                tmem_ld_wait();
                if (o < p.Co) {
                    const int i = i0 + c0;
                    if (vec && i + 32 <= p.Ci) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(wrow + i + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (i + j < p.Ci) wrow[i + j] = __uint_as_float(v[j]);
                    }
                }
            }
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(W5_TMEM_COLS) : "memory");
    }
}

// dw[o][i][t] = sum_s ws[s][t][o][i], fixed summation order
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int splits, int Co, int Ci)
{
    const long long oi = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long plane = (long long)Co * Ci;
    if (oi >= plane) return;
#pragma unroll
    for (int t = 0; t < 9; t++) {
        float acc = 0.f;
        for (int s = 0; s < splits; s++) acc += ws[((long long)s * 9 + t) * plane + oi];
        dw[oi * 9 + t] = acc;
    }
}

static int w5_splits(int N, int Ci, int H, int W, int Co, int pad)
{
    const int OH = H + 2 * pad - 2, OW = W + 2 * pad - 2;
    const int BN = Ci > 64 ? 128 : 64;
    const long long ctas = (long long)ceil_div(Co, 128) * ceil_div(Ci, BN) * 3;
    const long long steps = (long long)N * OH * ceil_div(OW, W5_KB);
    long long splits = (2LL * sm_count()) / ctas;                     // just under two full waves of one CTA per SM
    if (splits > steps) splits = steps;
    if (splits > 1024) splits = 1024;
    if (splits < 1) splits = 1;
    return (int)splits;
}

}  // namespace afcm

using namespace afcm;

extern "C" int64_t afcm_conv2d_wgrad_tc_workspace(int N, int Ci, int H, int W, int Co, int pad)
{
    if (N <= 0 || Ci <= 0 || Co <= 0 || H <= 0 || W <= 0 || pad < 0 || pad > 2) return 0;
    return (int64_t)w5_splits(N, Ci, H, W, Co, pad) * 9 * Co * Ci * (int64_t)sizeof(float);
}

extern "C" int afcm_conv2d_wgrad_tc5(const void* dyp, const void* xp, float* dw, void* workspace, int64_t workspace_bytes,
                                     int tc_dtype, int N, int Ci, int H, int W, int Co, int pad, void* stream)
{
    AFCM_CHECK_ARG(dyp && xp && dw && workspace, "dyp, xp, dw and workspace must be given");
    AFCM_CHECK_ARG(N > 0 && Ci > 0 && Co > 0 && H > 0 && W > 0, "empty problem");
    AFCM_CHECK_ARG(tc_dtype == AFCM_F16 || tc_dtype == AFCM_BF16, "tc dtype must be F16 or BF16");
    if (pad < 0 || pad > 2) { set_error("conv2d_wgrad_tc5: padding %d not supported (0..2)", pad); return AFCM_ERR_UNSUPPORTED; }
    W5Params p;
    memset(&p, 0, sizeof(p));
    p.ws = (float*)workspace;
    p.N = N; p.Ci = Ci; p.Co = Co; p.H = H; p.W = W; p.pad = pad;
    p.OH = H + 2 * pad - 2; p.OW = W + 2 * pad - 2;
    AFCM_CHECK_ARG(p.OH > 0 && p.OW > 0, "output must be at least 1x1");
    AFCM_CHECK_ARG((long long)N * p.OH < 0x7fffffffLL, "too many rows");
    p.co_pad = (Co + 7) & ~7; p.ci_pad = (Ci + 7) & ~7;
    p.BN = Ci > 64 ? 128 : 64;
    p.tiles_i = ceil_div(Ci, p.BN); p.tiles_o = ceil_div(Co, 128);
    p.nchunk = ceil_div(p.OW, W5_KB);
    p.steps = (long long)N * p.OH * p.nchunk;
    const int splits = w5_splits(N, Ci, H, W, Co, pad);
    AFCM_CHECK_ARG(workspace_bytes >= (int64_t)splits * 9 * Co * Ci * (int64_t)sizeof(float), "workspace too small (afcm_conv2d_wgrad_tc_workspace)");
    // instruction descriptor: D = F32, A/B = F16|BF16, A and B MN-major (bits 15, 16), N, M = 128
    const unsigned fmt = tc_dtype == AFCM_BF16 ? 1u : 0u;
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((unsigned)(p.BN >> 3) << 17) | ((128u >> 4) << 24);

    CUtensorMap map_a, map_b;
    {   // A: dyp viewed as [channel][x (OW+2)][row (N*OH)]
        const uint64_t dims[3] = {(uint64_t)p.co_pad, (uint64_t)(p.OW + 2), (uint64_t)N * p.OH};
        const uint64_t strides[2] = {(uint64_t)p.co_pad * 2, (uint64_t)p.co_pad * 2 * (p.OW + 2)};
        const uint32_t box[3] = {64, W5_KB, 1};
        int rc = encode_tiled(&map_a, tc_dtype, dyp, 3, dims, strides, box);
        if (rc) return rc;
    }
    {   // B: xp viewed as [channel][x (W+2)][y (H)][n]
        const uint64_t dims[4] = {(uint64_t)p.ci_pad, (uint64_t)(W + 2), (uint64_t)H, (uint64_t)N};
        const uint64_t strides[3] = {(uint64_t)p.ci_pad * 2, (uint64_t)p.ci_pad * 2 * (W + 2), (uint64_t)p.ci_pad * 2 * (W + 2) * H};
        const uint32_t box[4] = {64, W5_BROWS, 1, 1};
        int rc = encode_tiled(&map_b, tc_dtype, xp, 4, dims, strides, box);
        if (rc) return rc;
    }
    const int fixed = 256 + 1024;
    int stages = (max_smem_optin() - fixed) / W5_STAGE_BYTES;
    if (stages > W5_MAX_STAGES) stages = W5_MAX_STAGES;
    if (stages < 2) { set_error("conv2d_wgrad_tc5: not enough shared memory for the pipeline"); return AFCM_ERR_UNSUPPORTED; }
    p.stages = stages;
    const int smem = stages * W5_STAGE_BYTES + fixed;
    cudaStream_t st = (cudaStream_t)stream;
    AFCM_CUDA(cudaFuncSetAttribute(wgrad_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid((unsigned)(p.tiles_i * p.tiles_o * 3), (unsigned)splits);
    wgrad_tc5_kernel<<<grid, W5_THREADS, smem, st>>>(map_a, map_b, p);
    AFCM_LAUNCH_CHECK();
    count_launch();
    wgrad_reduce_kernel<<<ceil_div((long long)Co * Ci, 256), 256, 0, st>>>(p.ws, dw, splits, Co, Ci);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}
