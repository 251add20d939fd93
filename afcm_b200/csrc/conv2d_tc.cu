// conv2d_tc.cu -- 3x3 convolution / modulated convolution as a tcgen05 + TMEM implicit GEMM (sm_100a).
//
// Replaces the cuDNN grouped convolution the reference reaches through modulated_conv2d
// (models/networks/stylegan3/networks_stylegan3.py:25-64 -> torch_utils/ops/conv2d_gradfix.py:37-40)
// and the plain encoder convolution (networks_stylegan3.py:503-505).
//
// Formulation (see DESIGN.md "modulated_conv2d"):
//   * activations are pre-scaled by the modulation coefficient and packed to 16 bit in a channel-
//     innermost "flat plane" layout  xp[n][y*(W+2) + x][ci]  (ci padded to a multiple of 8) with two
//     zero pixels at the end of every row.  A 3x3 tap (ky,kx) is then a pure shift of the flat pixel
//     index by (ky-pad)*(W+2) + (kx-pad): horizontal taps wrap onto the zero pixels, vertical taps run
//     off the plane where TMA zero-fills.  The shift sits on an OUTER tensor-map dimension because TMA
//     requires the innermost start to be 16-byte aligned (measured: an unaligned pixel-innermost box
//     raises an illegal-instruction fault).
//   * GEMM:  D[p, o] = sum_{tap} sum_{ci} A_tap[p, ci] * B_tap[o, ci]
//       M = 128 consecutive flat output pixels p of one sample (all of them valid outputs for pad=2)
//       N = BN <= 256 output channels,  K = 9 taps x Ci (64 channels per pipeline stage)
//       A_tap tile : one TMA box   [128 px][64 ci] -> K-major SWIZZLE_128B operand
//       B_tap tile : one TMA box   [BN o][64 ci]   -> K-major SWIZZLE_128B operand
//       D          : fp32 accumulator in tensor memory, 2 stages x 256 columns (epilogue of tile i
//                    overlaps the main loop of tile i+1)
//   * warp roles: warp 0 TMA producer, warp 1 tcgen05.mma issuer (one lane), warp 2 TMEM allocator,
//     warps 4-7 epilogue (tcgen05.ld -> demodulation scale -> coalesced NCHW fp32 stores)
//   * persistent CTAs (one per SM), static tile striding, output-channel tiles innermost so CTAs that
//     run concurrently share their activation tiles in L2.
//   * ASW variant (afcm_conv2d_tc_nchw): the A operand is built IN the kernel from the fp16 NCHW planes the preceding
//     filtered_lrelu wrote -- no packed copy of the activations exists.  Warp 3 streams RAW tiles [64 channels][152 plane
//     elements] through a TMA ring (box on the flat [N][Ci][H pitch] view of the tensor, 16-byte aligned start, zero fill
//     outside the plane and beyond Ci); eight producer warps (12-19), one per 8-channel block, read aligned pixel pairs of two
//     channel rows per lane from it, byte-permute them into {channel pair, pixel} words, apply the modulation coefficient in
//     fp32 and store the words straight into the K-major SWIZZLE_128B tile the tensor core reads (the swizzle makes the 32
//     stores of a warp hit 32 different banks; the 304-byte raw rows do the same for the loads), then fence.proxy.async +
//     mbarrier arrive.  The row pitch W+2 of the flat-plane formulation is virtual: flat pixel p = y (W+2) + x lives at element
//     p - 2y of the plane; x >= W and rows outside the plane are zeros.
#include <cuda.h>
#include "afcm_common.cuh"
#include "tc_ptx.cuh"

namespace afcm {

constexpr int TC_BM = 128;            // pixels per tile (UMMA M)
constexpr int TC_BK = 64;             // channels per pipeline stage
constexpr int TC_MAX_STAGES = 8;      // pipeline depth is chosen per launch: small channel tiles need more stages in flight
constexpr int TC_THREADS = 384;         // warps 0-3: TMA producer, MMA issuer, TMEM allocator, idle; warps 4-11: two epilogue groups
constexpr int TC_ASW_WARPS = 8;         // ASW variant: warps 12-19 build the A tiles from NCHW planes (one 8-channel block each)
constexpr int TC_ASW_THREADS = TC_THREADS + 32 * TC_ASW_WARPS;
constexpr int TC_AROW_BLKS = 17;        // 8-pixel blocks of a row-reuse A tile (TC_AROW_PX / 8)
constexpr int TC_RAW_PX = 152;          // ASW raw tile: plane elements per channel row (136 + alignment slack; 304-byte rows: conflict-free)
constexpr int TC_RAW_BYTES = TC_RAW_PX * TC_BK * 2;    // 19 KB
constexpr int TC_RAW_MAX_STAGES = 8;    // depth of the raw ring is chosen per launch (p.raw_stages): it carries the global-memory latency
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;          // 16 KB
constexpr int TC_AROW_PX = TC_BM + 8;                  // row-reuse mode: 128 pixels + the two pixels to the right, rounded to 8
constexpr int TC_AROW_BYTES = TC_AROW_PX * TC_BK * 2;  // 17 KB
constexpr int TC_AROWP_BYTES = 18 * 8 * TC_BK * 2;     // ASW on pitched planes: 144 pixels from an 8-pixel boundary (18 KB), see the producers
constexpr int TC_TMEM_COLS = 512;
struct TcParams {
    const float* ocoef;     // [N, Co] or null
    const float* bias;      // [Co] or null: added after the demodulation scale (the bias of the following filtered_lrelu)
    void* y;                // [N, Co, OH, OW] fp32, or fp16 when y_half
    int y_half;
    int N, Ci, Co, H, W, Wp, OH, OW, pad;
    int BN, n_tiles, m_tiles, cblocks;     // channel tile, #channel tiles, #pixel tiles per sample, ceil(Ci/64)
    int total_tiles;
    int stages;                            // TMA -> MMA ring depth (2 .. TC_MAX_STAGES)
    int nk_last;                           // UMMA_K steps of the last 64-channel block that hold real channels: ceil((Ci mod 64) / 16), 1..4
    int issuers;                           // MMA-issuing warps: 1, or 2 taking alternate tiles (not in pipeline mode 2)
    int rowreuse;                          // 1: one A tile of 136 pixels per (ky, channel block) serves the three kx taps;
                                           // 2: the same with separate rings for the A rows and the per-tap weight tiles
    int a_stages;                          // mode 2: depth of the A ring (`stages` is the depth of the B ring)
    int raw_stages;                        // ASW: depth of the raw-tile ring
    int bres;                              // 1 (row-reuse mode, single channel tile): all weight tiles stay resident in shared memory
    unsigned idesc;
    const __half* xn;       // ASW: [N, Ci, H, W] fp16 activations (NCHW, contiguous)
    const float* icoef;     // ASW: [N, Ci] modulation coefficients or null
    int pitched;            // ASW: the planes are stored at the row pitch W + 2 with zero pad pixels: the flat plane exists in memory
    long long* trace;       // development aid: (code << 48 | clock64) records of CTA 0 per role (0 TMA, 1 MMA, 2 epilogue group 0), or null
    unsigned* dbg;          // mapped host memory for progress markers (AFCM_TC_DEBUG), or null
    int dbg_mode;           // debug bisection switches (see afcm_conv_tc_debug_buffer)
};

// Slot sequence of the TMA -> MMA ring, as every role of the kernel walks it.  With two MMA issuers the ring is split into two
// halves used by alternate tiles of the CTA: each slot then has one consumer that takes its fills in order (an mbarrier
// parity wait cannot tell fill f from fill f + 2, so two consumers must not share a slot).
struct TcRing {
    int depth, issuers, h, stage, st0, st1;
    uint32_t phase, ph0, ph1;
    __device__ TcRing(int stages, int issuers_)
        : depth(issuers_ == 2 ? stages >> 1 : stages), issuers(issuers_), h(0), stage(0), st0(0), st1(0), phase(0), ph0(0), ph1(0) {}
    __device__ __forceinline__ void begin_tile(int it) { h = issuers == 2 ? (it & 1) : 0; stage = h ? st1 : st0; phase = h ? ph1 : ph0; }
    __device__ __forceinline__ void end_tile() { if (h) { st1 = stage; ph1 = phase; } else { st0 = stage; ph0 = phase; } }
    __device__ __forceinline__ int slot() const { return h * depth + stage; }
    __device__ __forceinline__ void next() { if (++stage == depth) { stage = 0; phase ^= 1; } }
};

constexpr int TC_TRACE_SLOTS = 4096;
struct TcTrace {
    long long* base; int n;
    __device__ TcTrace(long long* t, int role, bool on) : base(t && on && blockIdx.x == 0 ? t + role * TC_TRACE_SLOTS : nullptr), n(0) {}
    __device__ __forceinline__ void mark(int code)
    {
        if (base && n < TC_TRACE_SLOTS) base[n++] = ((long long)code << 48) | (clock64() & 0xffffffffffffLL);
    }
};

__device__ __forceinline__ uint32_t pack_tc(float lo, float hi, __half*);
__device__ __forceinline__ uint32_t pack_tc(float lo, float hi, __nv_bfloat16*);

// ---- the kernel --------------------------------------------------------------------------------------
template <bool ASW>
__global__ void __launch_bounds__(ASW ? TC_ASW_THREADS : TC_THREADS, 1)
conv2d_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ TcParams p)
{
    // ASW: map_a is the RAW map (flat [H pitch] x Ci x N view of the NCHW tensor, no swizzle); map_a2 is unused
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.BN * TC_BK * 2;
    const int arow_bytes = (ASW && p.pitched) ? TC_AROWP_BYTES : TC_AROW_BYTES;       // A tile of a kernel row
    const int stage_bytes = p.rowreuse == 2 ? b_bytes
                          : (p.bres ? arow_bytes : (p.rowreuse ? arow_bytes + 3 * b_bytes : TC_A_BYTES + b_bytes));
    // in front of the ring: the resident weights (tile (tap, cb) at (tap * cblocks + cb) * b_bytes), or the A ring of mode 2
    const int res_bytes = p.rowreuse == 2 ? p.a_stages * arow_bytes : (p.bres ? 9 * p.cblocks * b_bytes : 0);
    uint8_t* raw = smem;                                                // ASW: TC_RAW_STAGES raw tiles in front of everything
    if (ASW) smem += p.raw_stages * TC_RAW_BYTES;                       // 19 x 1024 bytes each: the operand tiles stay 1024-byte aligned
    uint8_t* ring = smem + res_bytes;
    uint8_t* tail = ring + p.stages * stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);                 // [TC_MAX_STAGES]
    uint64_t* empty = full + TC_MAX_STAGES;                             // [TC_MAX_STAGES]
    uint64_t* tfull = empty + TC_MAX_STAGES;                            // [2]
    uint64_t* tempty = tfull + 2;                                       // [2]
    uint64_t* bfull = tempty + 2;                                       // [1]
    uint64_t* afull = bfull + 1;                                        // [8] mode 2: A ring
    uint64_t* aempty = afull + 8;                                       // [8]
    uint64_t* rfull = aempty + 8;                                       // [8] ASW: raw ring
    uint64_t* rempty = rfull + TC_RAW_MAX_STAGES;                       // [8]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rempty + TC_RAW_MAX_STAGES);
    float* s_ocoef = reinterpret_cast<float*>(tail + 512);              // [2][256]  (53 barriers + the TMEM slot take 428 bytes)
    float* s_bias = s_ocoef + 512;                                      // [2][256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        // ASW: an A tile is complete when every producer warp has arrived (+ the TMA thread's expect_tx arrive where the stage
        // also carries weight tiles)
        const uint32_t npw = (uint32_t)TC_ASW_WARPS;
        const uint32_t full_count = !ASW ? 1u : (p.rowreuse == 2 ? 1u : (p.bres ? npw : npw + 1u));
        for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], full_count); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        mbar_init(bfull, 1);
        for (int a = 0; a < 8; a++) { mbar_init(&afull[a], ASW ? npw : 1u); mbar_init(&aempty[a], 1); }
        for (int a = 0; a < TC_RAW_MAX_STAGES; a++) { mbar_init(&rfull[a], 1); mbar_init(&rempty[a], npw); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (blockIdx.x == 0 && threadIdx.x == 0) { dbg_mark(p.dbg, 5, tmem_base); dbg_mark(p.dbg, 6, 0xC0DE0001u); }

    const int kblocks = (p.rowreuse ? 3 : 9) * p.cblocks;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            TcTrace tr(p.trace, 0, true);
            int stage = 0; uint32_t phase = 0;
            if (p.bres) {
                mbar_expect_tx(bfull, (uint32_t)res_bytes);
                for (int tap = 0; tap < 9; tap++)
                    for (int cb = 0; cb < p.cblocks; cb++)
                        tma_load_3d(smem + (tap * p.cblocks + cb) * b_bytes, &map_b, bfull, cb * TC_BK, 0, tap);
            }
            if (p.rowreuse == 2) {
                // separate rings: an A row tile per (ky, channel block), a weight tile per (tap, channel block)
                int as = 0; uint32_t aph = 0;
                for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                    const int nt = tile % p.n_tiles;
                    const int r = tile / p.n_tiles;
                    const int mt = r % p.m_tiles, n = r / p.m_tiles;
                    const int p0 = mt * TC_BM, o0 = nt * p.BN;
                    for (int ky = 0; ky < 3; ky++) {
                        for (int cb = 0; cb < p.cblocks; cb++) {
                            if (!ASW) {
                                mbar_wait(&aempty[as], aph ^ 1, p.dbg, 0x600u | (unsigned)as);
                                mbar_expect_tx(&afull[as], (uint32_t)TC_AROW_BYTES);
                                tma_load_3d(smem + as * TC_AROW_BYTES, &map_a2, &afull[as], cb * TC_BK, p0 + (ky - p.pad) * p.Wp - p.pad, n);
                                if (++as == p.a_stages) { as = 0; aph ^= 1; }
                            }
                            for (int kx = 0; kx < 3; kx++) {
                                mbar_wait(&empty[stage], phase ^ 1, p.dbg, 0x100u | (unsigned)stage);
                                mbar_expect_tx(&full[stage], (uint32_t)b_bytes);
                                tma_load_3d(ring + stage * b_bytes, &map_b, &full[stage], cb * TC_BK, o0, ky * 3 + kx);
                                if (++stage == p.stages) { stage = 0; phase ^= 1; }
                            }
                        }
                    }
                }
            } else {
            TcRing rg(p.stages, p.issuers);
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, it++) {
                rg.begin_tile(it);
                const int nt = tile % p.n_tiles;
                const int r = tile / p.n_tiles;
                const int mt = r % p.m_tiles, n = r / p.m_tiles;
                const int p0 = mt * TC_BM, o0 = nt * p.BN;
                for (int kb = 0; kb < ((p.dbg_mode & 8) ? min(kblocks, p.stages) : kblocks); kb++) {
                    if (p.bres) {
                        if (ASW) break;                                  // the stage holds the A tile only: nothing for TMA to do
                        const int ky = kb / p.cblocks, cb = kb - ky * p.cblocks;
                        const int stage = rg.slot();
                        mbar_wait(&empty[stage], rg.phase ^ 1, p.dbg, 0x100u | (unsigned)stage);
                        tr.mark(1);
                        if ((p.dbg_mode & 64) && ky > 0) {              // timing experiment (wrong results): one A tile per output tile instead of three
                            mbar_expect_tx(&full[stage], 0u);
                        } else {
                            mbar_expect_tx(&full[stage], (uint32_t)TC_AROW_BYTES);
                            tma_load_3d(ring + stage * stage_bytes, &map_a2, &full[stage], cb * TC_BK, p0 + (ky - p.pad) * p.Wp - p.pad, n);
                        }
                        rg.next();
                        continue;
                    }
                    if (p.rowreuse) {
                        // stage = (ky, channel block): one A tile of 136 consecutive flat pixels starting at the kx = 0
                        // tap + the three B tiles of that kernel row; kx becomes a 128-byte offset of the A descriptor
                        const int ky = kb / p.cblocks, cb = kb - ky * p.cblocks;
                        const int stage = rg.slot();
                        mbar_wait(&empty[stage], rg.phase ^ 1, p.dbg, 0x100u | (unsigned)stage);
                        uint8_t* sa = ring + stage * stage_bytes;
                        const bool skip_a = !ASW && (p.dbg_mode & 64) && ky > 0;     // timing experiment, see above
                        mbar_expect_tx(&full[stage], (uint32_t)((ASW || skip_a) ? stage_bytes - arow_bytes : stage_bytes));
                        if (!ASW && !skip_a) tma_load_3d(sa, &map_a2, &full[stage], cb * TC_BK, p0 + (ky - p.pad) * p.Wp - p.pad, n);
#pragma unroll
                        for (int kx = 0; kx < 3; kx++)
                            tma_load_3d(sa + arow_bytes + kx * b_bytes, &map_b, &full[stage], cb * TC_BK, o0, ky * 3 + kx);
                        rg.next();
                        continue;
                    }
                    const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
                    const int ky = tap / 3, kx = tap - ky * 3;
                    const int shift = (ky - p.pad) * p.Wp + (kx - p.pad);
                    dbg_mark(p.dbg, 10, (unsigned)kb + 1);
                    const int stage = rg.slot();
                    mbar_wait(&empty[stage], rg.phase ^ 1, p.dbg, 0x100u | (unsigned)stage);
                    dbg_mark(p.dbg, 11, (unsigned)kb + 1);
                    uint8_t* sa = ring + stage * stage_bytes;
                    uint32_t bytes = (uint32_t)stage_bytes;
                    if (p.dbg_mode & 1) bytes -= TC_A_BYTES;
                    if (p.dbg_mode & 2) bytes -= (uint32_t)b_bytes;
                    const int pa = (p.dbg_mode & 4) ? p0 : p0 + shift;
                    mbar_expect_tx(&full[stage], bytes);
                    dbg_mark(p.dbg, 12, (unsigned)kb + 1);
                    if (!(p.dbg_mode & 1)) {
                        tma_load_3d(sa, &map_a, &full[stage], cb * TC_BK, pa, n);
                        dbg_mark(p.dbg, 13, (unsigned)kb + 1);
                    }
                    if (!(p.dbg_mode & 2)) {
                        tma_load_3d(sa + TC_A_BYTES, &map_b, &full[stage], cb * TC_BK, o0, tap);
                        dbg_mark(p.dbg, 15, (unsigned)kb + 1);
                    }
                    if (blockIdx.x == 0) dbg_mark(p.dbg, 0, (unsigned)(kb + 1) | ((unsigned)tile << 16));
                    rg.next();
                }
                rg.end_tile();
            }
            }
        }
    } else if (warp == 1 || (warp == 2 && p.issuers == 2)) {
        // ================= MMA issuer(s) =================
        // The whole warp runs the loops and the barrier waits (warp-uniform control flow); one elected lane issues.
        // With two issuers (warps 1 and 2) the CTA's tiles alternate between them, each with its own accumulator: a satisfied
        // mbarrier try_wait still costs ~90 clocks and the tcgen05.mma queue is only a few instructions deep, so a single
        // issuer leaves the tensor pipe idle between stages (measured timeline: profiles/r02_conv_tc_trace.txt).
        if (!(p.dbg_mode & 8)) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const int iss = warp - 1;
            const uint32_t smem_base = smem_u32(smem), ring_base = smem_u32(ring);
            const uint32_t b_step16 = (p.bres ? (uint32_t)(p.cblocks * b_bytes) : (uint32_t)b_bytes) >> 4;
            if (p.bres) { mbar_wait(bfull, 0, p.dbg, 0x500u); tc_fence_after(); }
            if (p.rowreuse == 2) {
                int as = 0; uint32_t aph = 0;
                const int nrows = 3 * p.cblocks;
                for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                    mbar_wait(&tempty[acc], acc_phase ^ 1, p.dbg, 0x200u | (unsigned)acc);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
                    const int pp_base = ((tile / p.n_tiles) % p.m_tiles) * TC_BM - p.pad * p.Wp - p.pad;   // first pixel of the ky = 0 tile
                    int cb2 = 0, ky2 = 0;
                    for (int rb = 0; rb < nrows; rb++) {
                        mbar_wait(&afull[as], aph, p.dbg, 0x700u | (unsigned)as);
                        // pitched ASW tiles start at an 8-pixel boundary: the first pixel is row (pp0 mod 8) of the tile
                        const uint32_t row0 = (ASW && p.pitched) ? (uint32_t)((pp_base + ky2 * p.Wp) & 7) : 0u;
                        const uint32_t a_lo = desc_lo(smem_base + (uint32_t)(as * arow_bytes)) + row0 * 8u;
                        const int nk = (cb2 == p.cblocks - 1) ? p.nk_last : TC_BK / 16;      // K steps of all-zero padding channels are skipped
                        if (++cb2 == p.cblocks) { cb2 = 0; ky2++; }
#pragma unroll
                        for (int kx = 0; kx < 3; kx++) {
                            mbar_wait(&full[stage], phase, p.dbg, 0x300u | (unsigned)stage);
                            tc_fence_after();
                            const uint32_t b_lo = desc_lo(ring_base + (uint32_t)(stage * b_bytes));
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < TC_BK / 16; k++)
                                    if (k < nk)
                                        umma_f16_lohi(tmem_d, a_lo + (uint32_t)(kx * 8 + k * 2), b_lo + (uint32_t)(k * 2), TC_DESC_HI, p.idesc,
                                                      (rb | kx | k) != 0);
                                umma_commit(&empty[stage]);
                                if (kx == 2) {
                                    umma_commit(&aempty[as]);
                                    if (rb == nrows - 1) umma_commit(&tfull[acc]);
                                }
                            }
                            __syncwarp();
                            if (++stage == p.stages) { stage = 0; phase ^= 1; }
                        }
                        if (++as == p.a_stages) { as = 0; aph ^= 1; }
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            } else {
            TcTrace tr(p.trace, 1, lane == 0 && iss == 0);
            TcRing rg(p.stages, p.issuers);
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, it++) {
                if (p.issuers == 2) {
                    if ((it & 1) != iss) continue;                      // the other issuer's tile (in the other half of the ring)
                    acc = iss; acc_phase = (uint32_t)(it >> 1) & 1u;
                }
                rg.begin_tile(it);
                tr.mark(10);
                mbar_wait(&tempty[acc], acc_phase ^ 1, p.dbg, 0x200u | (unsigned)acc);
                tc_fence_after();
                tr.mark(11);
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
                const int pp_base = ((tile / p.n_tiles) % p.m_tiles) * TC_BM - p.pad * p.Wp - p.pad;       // first pixel of the ky = 0 tile
                int ky_ = 0, cb_ = 0;                                   // row-reuse mode: kb = ky_ * cblocks + cb_
                for (int kb = 0; kb < kblocks; kb++) {
                    const int stage = rg.slot();
                    mbar_wait(&full[stage], rg.phase, p.dbg, 0x300u | (unsigned)stage);
                    tc_fence_after();
                    tr.mark(12);
                    const uint32_t sa = ring_base + (uint32_t)(stage * stage_bytes);
                    const bool last = kb == kblocks - 1;
                    if (p.rowreuse) {
                        // weights of kernel row ky: in the stage behind the A tile, or in the resident region
                        const uint32_t sb3 = p.bres ? smem_base + (uint32_t)((ky_ * 3 * p.cblocks + cb_) * b_bytes) : sa + (uint32_t)arow_bytes;
                        // pitched ASW tiles start at an 8-pixel boundary: the first pixel is row (pp0 mod 8) of the tile
                        const uint32_t row0 = (ASW && p.pitched) ? (uint32_t)((pp_base + ky_ * p.Wp) & 7) : 0u;
                        const uint32_t a_lo = desc_lo(sa) + row0 * 8u, b_lo = desc_lo(sb3);
                        const int nk = (cb_ == p.cblocks - 1) ? p.nk_last : TC_BK / 16;          // K steps of all-zero padding channels are skipped
                        if (elect_one()) {
#pragma unroll
                            for (int kx = 0; kx < 3; kx++) {
#pragma unroll
                                for (int k = 0; k < TC_BK / 16; k++) {
                                    if (k >= nk) continue;
                                    // the kx tap is the same tile read one pixel (= one 128-byte swizzled row) further on
                                    // (the swizzle is a function of the shared-memory address, so the descriptor's
                                    // base-offset field stays 0: measured bit-exact on B200, tests/test_gpu_tc.py); one
                                    // UMMA_K step of 16 channels = 32 bytes inside the swizzled row (16-byte units)
                                    umma_f16_lohi(tmem_d, a_lo + (uint32_t)(kx * 8 + k * 2), b_lo + kx * b_step16 + (uint32_t)(k * 2),
                                                  TC_DESC_HI, p.idesc, (kb | kx | k) != 0);
                                }
                            }
                            umma_commit(&empty[stage]);
                            if (last) umma_commit(&tfull[acc]);
                        }
                        if (++cb_ == p.cblocks) { cb_ = 0; ky_++; }
                    } else {
                        // A and B are both K-major SWIZZLE_128B tiles of 128-byte rows: one UMMA_K step of 16
                        // channels = 32 bytes inside the swizzled row, 8-row groups 1024 B apart (SBO)
                        const uint32_t a_lo = desc_lo(sa), b_lo = desc_lo(sa + TC_A_BYTES);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < TC_BK / 16; k++)
                                umma_f16_lohi(tmem_d, a_lo + (uint32_t)(k * 2), b_lo + (uint32_t)(k * 2), TC_DESC_HI, p.idesc, (kb | k) != 0);
                            umma_commit(&empty[stage]);
                            if (last) umma_commit(&tfull[acc]);
                        }
                    }
                    __syncwarp();
                    tr.mark(13);
                    rg.next();
                }
                rg.end_tile();
                if (p.issuers != 2 && ++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            }
        }
    } else if (warp >= 4 && warp < 12 && !(p.dbg_mode & 8)) {
        // ================= epilogue =================
        // Two groups of four warps (one warp per TMEM lane quadrant each) take alternate 32-channel column chunks of
        // the accumulator: with small channel tiles the store loop, not the MMA, bounds the tile time.
        const int wq = warp & 3, grp = (warp - 4) >> 2;
        const int et = threadIdx.x - 128;
        TcTrace tr(p.trace, 2, et == 0);
        int acc = 0; uint32_t acc_phase = 0;
        const long long ohw = (long long)p.OH * p.OW;
        const long long cstride = ohw * (p.y_half ? 2 : 4);          // bytes between output channels
        // fp16 planes with even row lengths: neighbouring lanes (an even / odd pixel of the same row) exchange halves, so that
        // every lane stores one 4-byte pixel pair of one channel -- half the store instructions and conversions of the
        // element-wise path below
        const bool paired = p.y_half && !(p.OW & 1) && !(p.Wp & 1) && !(p.OH * p.OW & 1) && !(reinterpret_cast<uintptr_t>(p.y) & 3);
        int coef_key0 = -1, coef_key1 = -1;                          // (sample, channel tile) whose coefficients each buffer holds
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int nt = tile % p.n_tiles;
            const int r = tile / p.n_tiles;
            const int mt = r % p.m_tiles, n = r / p.m_tiles;
            const int o0 = nt * p.BN;
            const int key = n * p.n_tiles + nt;
            const bool reload = key != (acc ? coef_key1 : coef_key0);      // the same for every epilogue thread
            if (reload) {
                // every epilogue warp has finished the tiles that read the old contents of this buffer (the warps are not
                // synchronised per tile otherwise; found by compute-sanitizer racecheck)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                for (int j = et; j < p.BN; j += 256) {
                    const int o = o0 + j;
                    s_ocoef[acc * 256 + j] = (o < p.Co) ? (p.ocoef ? p.ocoef[(long long)n * p.Co + o] : 1.f) : 0.f;
                    s_bias[acc * 256 + j] = (o < p.Co && p.bias) ? p.bias[o] : 0.f;
                }
                if (acc) coef_key1 = key; else coef_key0 = key;
            }
            tr.mark(20);
            mbar_wait(&tfull[acc], acc_phase, p.dbg, 0x400u | (unsigned)acc);
            tc_fence_after();
            tr.mark(21);
            if (reload) asm volatile("bar.sync 1, 256;" ::: "memory");
            tr.mark(23);
            const int pix = mt * TC_BM + wq * 32 + lane;
            const int oy = pix / p.Wp, ox = pix - oy * p.Wp;
            const bool ok = oy < p.OH && ox < p.OW;
            char* ybase = reinterpret_cast<char*>(p.y) +
                          (((long long)n * p.Co + o0) * ohw + (long long)oy * p.OW + ox) * (p.y_half ? 2 : 4);
            // paired path: the even lane's address, the odd lane one channel further on
            char* ypair = ybase - (lane & 1) * 2 + (lane & 1) * cstride;
            for (int c0 = grp * 32; c0 < p.BN; c0 += 64) {
                uint32_t v[32];
                if (!(p.dbg_mode & 32)) {
                    tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * 256 + c0), v);
                    tmem_ld_wait();
                }
                tr.mark(24);
                if (p.dbg_mode & 16) continue;
                const float4* oc4 = reinterpret_cast<const float4*>(s_ocoef + acc * 256 + c0);
                const float4* bi4 = reinterpret_cast<const float4*>(s_bias + acc * 256 + c0);
                const int nch = min(32, p.Co - o0 - c0);               // channels of this chunk that exist
                if (paired) {
                    char* q = ypair + c0 * cstride;
                    const int myc = lane & 1;                          // this lane stores channels 2 j + myc
#pragma unroll
                    for (int j4 = 0; j4 < 8; j4++) {
                        const float4 oc = oc4[j4], bi = bi4[j4];
                        const float r0 = fmaf(__uint_as_float(v[4 * j4 + 0]), oc.x, bi.x), r1 = fmaf(__uint_as_float(v[4 * j4 + 1]), oc.y, bi.y);
                        const float r2 = fmaf(__uint_as_float(v[4 * j4 + 2]), oc.z, bi.z), r3 = fmaf(__uint_as_float(v[4 * j4 + 3]), oc.w, bi.w);
                        const uint32_t w01 = pack_tc(r0, r1, (__half*)nullptr), w23 = pack_tc(r2, r3, (__half*)nullptr);
                        const uint32_t p01 = __shfl_xor_sync(0xffffffffu, w01, 1), p23 = __shfl_xor_sync(0xffffffffu, w23, 1);
                        // even lane: (own low half, partner's low half) = channel 2j of pixels (L, L + 1); odd lane: (partner's
                        // high half, own high half) = channel 2j + 1 of pixels (L - 1, L)
                        const uint32_t o01 = myc ? __byte_perm(p01, w01, 0x7632) : __byte_perm(w01, p01, 0x5410);
                        const uint32_t o23 = myc ? __byte_perm(p23, w23, 0x7632) : __byte_perm(w23, p23, 0x5410);
                        if (ok && 4 * j4 + myc < nch) *reinterpret_cast<uint32_t*>(q) = o01;
                        if (ok && 4 * j4 + 2 + myc < nch) *reinterpret_cast<uint32_t*>(q + 2 * cstride) = o23;
                        q += 4 * cstride;
                    }
                } else if (ok) {
                    char* q = ybase + c0 * cstride;
                    if (nch >= 32) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; j4++) {
                            const float4 oc = oc4[j4], bi = bi4[j4];
                            const float r0 = fmaf(__uint_as_float(v[4 * j4 + 0]), oc.x, bi.x), r1 = fmaf(__uint_as_float(v[4 * j4 + 1]), oc.y, bi.y);
                            const float r2 = fmaf(__uint_as_float(v[4 * j4 + 2]), oc.z, bi.z), r3 = fmaf(__uint_as_float(v[4 * j4 + 3]), oc.w, bi.w);
                            if (p.y_half) {
                                *reinterpret_cast<__half*>(q) = __float2half_rn(r0); q += cstride;
                                *reinterpret_cast<__half*>(q) = __float2half_rn(r1); q += cstride;
                                *reinterpret_cast<__half*>(q) = __float2half_rn(r2); q += cstride;
                                *reinterpret_cast<__half*>(q) = __float2half_rn(r3); q += cstride;
                            } else {
                                *reinterpret_cast<float*>(q) = r0; q += cstride;
                                *reinterpret_cast<float*>(q) = r1; q += cstride;
                                *reinterpret_cast<float*>(q) = r2; q += cstride;
                                *reinterpret_cast<float*>(q) = r3; q += cstride;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            if (j < nch) {
                                const float rj = fmaf(__uint_as_float(v[j]), s_ocoef[acc * 256 + c0 + j], s_bias[acc * 256 + c0 + j]);
                                if (p.y_half) *reinterpret_cast<__half*>(q + j * cstride) = __float2half_rn(rj);
                                else *reinterpret_cast<float*>(q + j * cstride) = rj;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            tr.mark(22);
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (blockIdx.x == 0 && et == 0) dbg_mark(p.dbg, 4, (unsigned)tile + 1);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    if (ASW && warp == 3 && lane == 0) {
        // ================= raw-tile TMA producer (ASW) =================
        // One box [64 channels][TC_RAW_PX plane elements] per (tile, ky, channel block), starting at the 8-element boundary at or
        // below the first element the A tile needs: element index of flat pixel p = p - 2 floor(p / Wp) (dense planes) or p (planes
        // stored at the pitch W + 2).
        int rs = 0; uint32_t rph = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int r = tile / p.n_tiles;
            const int mt = r % p.m_tiles, n = r / p.m_tiles;
            for (int ky = 0; ky < 3; ky++) {
                const int pp0 = mt * TC_BM + (ky - p.pad) * p.Wp - p.pad;
                const int y0 = (pp0 + 4 * p.Wp) / p.Wp - 4;
                const int a0 = ((p.pitched ? pp0 : pp0 - 2 * y0) >> 3) << 3;   // floor to a multiple of 8 (also for negatives)
                for (int cb = 0; cb < p.cblocks; cb++) {
                    mbar_wait(&rempty[rs], rph ^ 1, p.dbg, 0x900u | (unsigned)rs);
                    mbar_expect_tx(&rfull[rs], (uint32_t)TC_RAW_BYTES);
                    tma_load_3d(raw + rs * TC_RAW_BYTES, &map_a, &rfull[rs], a0, cb * TC_BK, n);
                    if (++rs == p.raw_stages) { rs = 0; rph ^= 1; }
                }
            }
        }
    }
    if (ASW && warp >= 12) {
        // ================= A-tile producers (ASW): raw [channel][element] tile -> K-major SWIZZLE_128B tile =================
        // Warp pw owns the 8-channel block pw of the 64-channel stage (= 16-byte chunk pw of every tile row).  Per iteration
        // the warp moves 16 pixels x 8 channels: lane (b2, m, q) reads pixel pair m of 8-pixel block 2 it + b2 from the channel
        // rows 2q and 2q + 1 of the raw tile (two 4-byte loads, conflict-free with the 304-byte raw rows), byte-permutes them
        // into the words {channel pair q of the even pixel, of the odd pixel} and stores each into its tile row (32 different
        // banks: the half-warps store even / odd pixels in opposite order).  Variants measured and rejected: an in-register
        // 8 x 8 transpose with movmatrix (~16 clocks per warp instruction per SM: 2900 clocks per tile); the producer warps
        // fetching the planes themselves with ld.global or cp.async rings (2-3 tiles of prefetch per warp do not cover the
        // global-memory latency: 50 ms over the 28 layers instead of 36).
        const int pw = warp - 12, q = lane & 3, m = (lane >> 2) & 3, b2 = lane >> 4;
        const uint32_t raw_lane = smem_u32(raw) + (uint32_t)((8 * pw + 2 * q) * (TC_RAW_PX * 2));
        const uint32_t off_even = (uint32_t)((8 * b2 + 2 * m) * 128 + ((pw ^ (2 * m)) << 4) + 4 * q);          // tile row & 7 == 2m
        const uint32_t off_odd = (uint32_t)((8 * b2 + 2 * m + 1) * 128 + ((pw ^ (2 * m + 1)) << 4) + 4 * q);   // tile row & 7 == 2m + 1
        const uint32_t off_first = b2 ? off_odd : off_even, off_second = b2 ? off_even : off_odd;
        constexpr int NIT = (TC_AROW_BLKS + 1) / 2;                                // 9 iterations of 16 pixels (the last one: 8)
        const bool two_rings = p.rowreuse == 2;
        const int nst = two_rings ? p.a_stages : p.stages;
        const int sbytes = two_rings ? arow_bytes : stage_bytes;
        uint8_t* abase = two_rings ? smem : ring;
        uint64_t* fullb = two_rings ? afull : full;
        uint64_t* emptyb = two_rings ? aempty : empty;
        const unsigned HWp = (unsigned)(p.H * p.Wp);
        auto tile_origin = [&](int tile, int ky, int& n, int& pp0, int& y0, int& a0) {
            const int r = tile / p.n_tiles;
            const int mt = r % p.m_tiles;
            n = r / p.m_tiles;
            pp0 = mt * TC_BM + (ky - p.pad) * p.Wp - p.pad;
            y0 = (pp0 + 4 * p.Wp) / p.Wp - 4;                                      // floor(pp0 / Wp): pp0 >= -2 Wp - 2
            a0 = ((p.pitched ? pp0 : pp0 - 2 * y0) >> 3) << 3;                    // floor to a multiple of 8 (also for negatives)
        };
        TcRing rg(nst, two_rings ? 1 : p.issuers);
        int rs = 0; uint32_t rph = 0;
        int tit = 0;
        if (p.pitched) {
            // Planes stored at the pitch W + 2 ARE the flat plane (pad pixels are stored zeros; outside the plane TMA zero-fills), so
            // the tile is a pure [channel][pixel] -> [pixel][channel] transposition, and it can start at the raw tile's own 8-pixel
            // boundary a0: the MMA descriptors skip the first (pp0 - a0) rows instead (a row offset of a K-major SWIZZLE_128B tile
            // is free).  ldmatrix.trans reads an 8 x 8 block (8 channel rows x 16 bytes, conflict-free with the 304-byte raw rows)
            // and hands every lane {channels 2t, 2t+1} of pixel lane / 4 -- exactly the fragment stmatrix stores as the 16-byte chunk
            // of 8 pixel rows of the swizzled tile: 10 shared-memory instructions per warp and stage instead of 54.  The loop is
            // software-pipelined over the flattened (tile, kernel row, channel block) sequence: the raw tile of the NEXT stage is
            // read into registers before this stage waits for its operand slot, so the two barrier round trips overlap.
            const uint32_t r8 = lane & 7u, mi = lane >> 3;
            const uint32_t rlane = smem_u32(raw) + (uint32_t)(8 * pw + (int)r8) * (TC_RAW_PX * 2) + mi * 16u;
            const uint32_t dlane = (mi * 8u + r8) * 128u + (((uint32_t)pw ^ r8) << 4);
            auto load_raw = [&](uint32_t (&f)[18], int slot) {
                const uint32_t rbase = rlane + (uint32_t)(slot * TC_RAW_BYTES);
#pragma unroll
                for (int g4 = 0; g4 < 4; g4++)
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(f[4 * g4]), "=r"(f[4 * g4 + 1]), "=r"(f[4 * g4 + 2]), "=r"(f[4 * g4 + 3]) : "r"(rbase + (uint32_t)(64 * g4)) : "memory");
                asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(f[16]), "=r"(f[17]) : "r"(rbase + 256u) : "memory");
            };
            uint32_t f[18], fn[18];
            int tile = blockIdx.x, ky = 0, cb = 0;
            bool valid = tile < p.total_tiles;
            int rs_cur = 0;
            if (valid) {
                mbar_wait(&rfull[rs], rph, p.dbg, 0xa00u | (unsigned)rs);
                load_raw(f, rs);
                rs_cur = rs;
                if (++rs == p.raw_stages) { rs = 0; rph ^= 1; }
            }
            while (valid) {
                if (ky == 0 && cb == 0) rg.begin_tile(tit);
                int ntile = tile, nky = ky, ncb = cb + 1;
                if (ncb == p.cblocks) { ncb = 0; if (++nky == 3) { nky = 0; ntile += gridDim.x; } }
                const bool nvalid = ntile < p.total_tiles;
                // The raw slot may be refilled once the loads have RETURNED, not merely issued (see the dense path below): the
                // arrive carries a data dependency on every loaded register that the compiler cannot fold.
                uint32_t loaded = 0u;
#pragma unroll
                for (int j = 0; j < 18; j++) loaded |= f[j];
                __syncwarp();
                if (lane == 0)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&rempty[rs_cur]) + (loaded & (uint32_t)p.dbg_mode)) : "memory");
                if (nvalid) {                                                      // prefetch the next stage's raw tile
                    mbar_wait(&rfull[rs], rph, p.dbg, 0xa00u | (unsigned)rs);
                    load_raw(fn, rs);
                    rs_cur = rs;
                    if (++rs == p.raw_stages) { rs = 0; rph ^= 1; }
                }
                const int stage = rg.slot();
                mbar_wait(&emptyb[stage], rg.phase ^ 1, p.dbg, 0x800u | (unsigned)stage);
                if (p.icoef) {                                                     // fp32 product, one rounding: as the pack kernel
                    const int n = (tile / p.n_tiles) / p.m_tiles;
                    const int cs = cb * TC_BK + 8 * pw + 2 * q;
                    const float s0 = cs < p.Ci ? p.icoef[(long long)n * p.Ci + cs] : 0.f;
                    const float s1 = cs + 1 < p.Ci ? p.icoef[(long long)n * p.Ci + cs + 1] : 0.f;
#pragma unroll
                    for (int j = 0; j < 18; j++) {
                        const float2 fv = __half22float2(*reinterpret_cast<const __half2*>(&f[j]));
                        f[j] = pack_tc(fv.x * s0, fv.y * s1, (__half*)nullptr);
                    }
                }
                const uint32_t dbase = smem_u32(abase + stage * sbytes) + dlane;
#pragma unroll
                for (int g4 = 0; g4 < 4; g4++)
                    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};"
                                 ::"r"(dbase + (uint32_t)(4096 * g4)), "r"(f[4 * g4]), "r"(f[4 * g4 + 1]), "r"(f[4 * g4 + 2]), "r"(f[4 * g4 + 3]) : "memory");
                asm volatile("stmatrix.sync.aligned.m8n8.x2.shared.b16 [%0], {%1, %2};" ::"r"(dbase + 16384u), "r"(f[16]), "r"(f[17]) : "memory");
                fence_proxy_async();                                               // generic-proxy stores -> visible to tcgen05.mma
                __syncwarp();
                if (lane == 0) mbar_arrive(&fullb[stage]);
                rg.next();
                if (nky == 0 && ncb == 0) { rg.end_tile(); tit++; }
#pragma unroll
                for (int j = 0; j < 18; j++) f[j] = fn[j];
                tile = ntile; ky = nky; cb = ncb; valid = nvalid;
            }
        } else
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, tit++) {
            rg.begin_tile(tit);
            for (int ky = 0; ky < 3; ky++) {
                int n, pp0, y0, a0;
                tile_origin(tile, ky, n, pp0, y0, a0);
                for (int cb = 0; cb < p.cblocks; cb++) {
                    // modulation coefficients of the channel pair this lane assembles
                    float s0 = 1.f, s1 = 1.f;
                    if (p.icoef) {
                        const int cs = cb * TC_BK + 8 * pw + 2 * q;
                        s0 = cs < p.Ci ? p.icoef[(long long)n * p.Ci + cs] : 0.f;
                        s1 = cs + 1 < p.Ci ? p.icoef[(long long)n * p.Ci + cs + 1] : 0.f;
                    }
                    mbar_wait(&rfull[rs], rph, p.dbg, 0xa00u | (unsigned)rs);
                    int pp = pp0 + 8 * b2 + 2 * m;                                 // flat pixel of this lane's pair in iteration 0 (even)
                    const uint32_t src = raw_lane + (uint32_t)(rs * TC_RAW_BYTES) - (uint32_t)(2 * a0);
                    uint32_t v0[NIT], v1[NIT];
                    {
                        int y = y0, x = pp - y0 * p.Wp;
                        if (x >= p.Wp) { x -= p.Wp; y++; }
#pragma unroll
                        for (int it = 0; it < NIT; it++) {
                            // rows outside the plane / channels >= Ci were zero-filled; pad pixels x >= W are zeros by definition
                            const bool ok = (unsigned)pp < HWp && x < p.W && (it < NIT - 1 || b2 == 0);
                            v0[it] = 0u; v1[it] = 0u;
                            if (ok) {
                                const uint32_t a = src + (uint32_t)(2 * (pp - 2 * y));  // element y W + x = pp - 2y of the plane
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v0[it]) : "r"(a) : "memory");
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v1[it]) : "r"(a + (uint32_t)(TC_RAW_PX * 2)) : "memory");
                            }
                            pp += 16; x += 16;
                            if (x >= p.Wp) { x -= p.Wp; y++; }
                        }
                    }
                    // The raw slot may be refilled once the loads have RETURNED, not merely issued: the arrive is handled by the
                    // barrier unit and overtakes loads still queued behind the tensor core's operand reads in the shared-memory
                    // pipe (measured: with a 2-deep operand ring the refill then lands under the last loads of a warp -- 16 of
                    // 16 runs of a 128 -> 128 channel layer wrong in a few dozen pixels).  Hence a data dependency the compiler
                    // cannot fold: the barrier address gets (OR of every loaded register) & dbg_mode added, and dbg_mode is a
                    // kernel parameter that is always 0 on this path.
                    uint32_t loaded = 0u;
#pragma unroll
                    for (int it = 0; it < NIT; it++) loaded |= v0[it] | v1[it];
                    __syncwarp();
                    if (lane == 0)
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&rempty[rs]) + (loaded & (uint32_t)p.dbg_mode)) : "memory");
                    if (++rs == p.raw_stages) { rs = 0; rph ^= 1; }
                    const int stage = rg.slot();
                    mbar_wait(&emptyb[stage], rg.phase ^ 1, p.dbg, 0x800u | (unsigned)stage);
                    const uint32_t dst = smem_u32(abase + stage * sbytes);
#pragma unroll
                    for (int it = 0; it < NIT; it++) {
                        uint32_t te = __byte_perm(v0[it], v1[it], 0x5410), to = __byte_perm(v0[it], v1[it], 0x7632);
                        if (p.icoef) {                                             // fp32 product, one rounding: as the pack kernel
                            const float2 fe = __half22float2(*reinterpret_cast<const __half2*>(&te));
                            const float2 fo = __half22float2(*reinterpret_cast<const __half2*>(&to));
                            te = pack_tc(fe.x * s0, fe.y * s1, (__half*)nullptr);
                            to = pack_tc(fo.x * s0, fo.y * s1, (__half*)nullptr);
                        }
                        if (it < NIT - 1 || b2 == 0) {
                            const uint32_t d = dst + (uint32_t)(it * 2048);
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(d + off_first), "r"(b2 ? to : te) : "memory");
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(d + off_second), "r"(b2 ? te : to) : "memory");
                        }
                    }
                    fence_proxy_async();                                           // generic-proxy stores -> visible to tcgen05.mma
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&fullb[stage]);
                    rg.next();
                }
            }
            rg.end_tile();
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

// ---- activation packing: fp32 / fp16 NCHW -> 16-bit [n][flat pixel][channel], modulation folded in -----
// One CTA transposes a tile of 128 flat pixels x 64 channels through shared memory.  Reads are aligned pixel
// PAIRS (float2 / half2: W, the row pitch W+2 and all strides are even, so a pair never straddles a row end),
// one warp per 8 channels; the tile is kept as [pixel][channel pair] words with a 33-word pitch (2-way store
// conflicts, conflict-free reads); writes are 16-byte stores, 4 pixels x 128 B per warp instruction.
__device__ __forceinline__ float2 load_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 load_pair(const __half* p) { return __half22float2(*reinterpret_cast<const __half2*>(p)); }
__device__ __forceinline__ uint32_t pack_tc(float lo, float hi, __half*)
{
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_tc(float lo, float hi, __nv_bfloat16*)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <typename TIN, typename TC>
__global__ void __launch_bounds__(256)
tc_pack_pairs_kernel(const TIN* __restrict__ x, const float* __restrict__ icoef, TC* __restrict__ xp,
                     int Ci, int c_pad, int H, int W, int Wp, int rows, int pitch)
{
    __shared__ uint32_t tile[128][33];
    const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long plane = (long long)H * pitch;              // pitch = W, or W + 2 for planes stored with zero pad columns
    int yy[2], xx[2];
    bool ok[2];
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const int pp = p0 + 2 * (lane + 32 * it);
        yy[it] = pp / Wp; xx[it] = pp - yy[it] * Wp;
        ok[it] = pp < rows && xx[it] < W;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int cpair = warp * 4 + j, c = c0 + 2 * cpair;
        const bool c_ok0 = c < Ci, c_ok1 = c + 1 < Ci;
        float s0 = 1.f, s1 = 1.f;
        if (icoef) { s0 = c_ok0 ? icoef[(long long)n * Ci + c] : 0.f; s1 = c_ok1 ? icoef[(long long)n * Ci + c + 1] : 0.f; }
        const TIN* xc = x + ((long long)n * Ci + c) * plane;
#pragma unroll
        for (int it = 0; it < 2; it++) {
            float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
            if (ok[it]) {
                const long long o = (long long)yy[it] * pitch + xx[it];
                if (c_ok0) a = load_pair(xc + o);
                if (c_ok1) b = load_pair(xc + plane + o);
            }
            const int r = 2 * (lane + 32 * it);
            tile[r][cpair] = pack_tc(a.x * s0, b.x * s1, (TC*)nullptr);
            tile[r + 1][cpair] = pack_tc(a.y * s0, b.y * s1, (TC*)nullptr);
        }
    }
    __syncthreads();
    const int g8 = tid & 7;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int pw = (tid >> 3) + 32 * j;
        const int pp = p0 + pw, cc = c0 + g8 * 8;
        if (pp < rows && cc < c_pad) {
            const uint4 v = make_uint4(tile[pw][g8 * 4], tile[pw][g8 * 4 + 1], tile[pw][g8 * 4 + 2], tile[pw][g8 * 4 + 3]);
            *reinterpret_cast<uint4*>(xp + ((long long)n * rows + pp) * c_pad + cc) = v;
        }
    }
}

// General variant (odd W or unaligned base): scalar reads, 64 pixels x 64 channels per CTA.
template <typename TIN, typename TC>
__global__ void __launch_bounds__(256)
tc_pack_kernel(const TIN* __restrict__ x, const float* __restrict__ icoef, TC* __restrict__ xp,
               int Ci, int c_pad, int H, int W, int Wp, int rows, int pitch)
{
    __shared__ float tile[64][65];
    const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
    const int tid = threadIdx.x;
    const int px = tid & 63, cq = tid >> 6;                 // read: 64 pixels x 4 channels per pass
    const int p = p0 + px;
    const int yy = p / Wp, xx = p - yy * Wp;
    const bool pix_ok = p < rows && xx < W;
#pragma unroll 4
    for (int j = 0; j < 16; j++) {
        const int c = c0 + cq * 16 + j;
        float v = 0.f;
        if (pix_ok && c < Ci) {
            v = (float)x[(((long long)n * Ci + c) * H + yy) * pitch + xx];
            if (icoef) v *= icoef[n * Ci + c];
        }
        tile[cq * 16 + j][px] = v;
    }
    __syncthreads();
    // write: thread -> (pixel, group of 8 channels) = one 16-byte store
    const int g8 = tid & 7;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int pw = (tid >> 3) + 32 * j;
        const int pp = p0 + pw, cc = c0 + g8 * 8;
        if (pp < rows && cc < c_pad) {
            alignas(16) TC v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = (TC)tile[g8 * 8 + k][pw];
            *reinterpret_cast<uint4*>(xp + ((long long)n * rows + pp) * c_pad + cc) = *reinterpret_cast<const uint4*>(v);
        }
    }
}

template <typename TIN, typename TC>
static void launch_pack(const void* x, int pitch, const float* icoef, void* xp, int N, int Ci, int H, int W, cudaStream_t st)
{
    const int Wp = W + 2, rows = H * Wp, c_pad = (Ci + 7) & ~7;
    const bool pairs = !(W & 1) && !(pitch & 1) && ((uintptr_t)x % (2 * sizeof(TIN))) == 0;
    if (pairs) {
        dim3 grid(ceil_div(rows, 128), ceil_div(c_pad, 64), N);
        tc_pack_pairs_kernel<TIN, TC><<<grid, 256, 0, st>>>((const TIN*)x, icoef, (TC*)xp, Ci, c_pad, H, W, Wp, rows, pitch);
    } else {
        dim3 grid(ceil_div(rows, 64), ceil_div(c_pad, 64), N);
        tc_pack_kernel<TIN, TC><<<grid, 256, 0, st>>>((const TIN*)x, icoef, (TC*)xp, Ci, c_pad, H, W, Wp, rows, pitch);
    }
}

// ---- host -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

int encode_tiled(CUtensorMap* map, int tc_dtype, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                 const uint32_t* box, bool swizzle128)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return AFCM_ERR_UNSUPPORTED; }
    cuuint64_t d[5], st[4];
    cuuint32_t b[5], estr[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; i++) { d[i] = dims[i]; b[i] = box[i]; }
    for (int i = 0; i + 1 < rank; i++) st[i] = strides[i];
    CUresult r = fn(map, tc_dtype == AFCM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                    const_cast<void*>(base), d, st, b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return AFCM_ERR_INVALID; }
    return AFCM_OK;
}

static int encode_3d(CUtensorMap* map, int tc_dtype, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                     uint64_t s1_bytes, uint64_t s2_bytes, uint32_t b0, uint32_t b1)
{
    const uint64_t dims[3] = {d0, d1, d2}, strides[2] = {s1_bytes, s2_bytes};
    const uint32_t box[3] = {b0, b1, 1};
    return encode_tiled(map, tc_dtype, base, 3, dims, strides, box);
}

static unsigned* g_dbg_host = nullptr;
static unsigned* g_dbg_dev = nullptr;
static int g_dbg_mode = 0;
static int g_force_stages = 0;
static int g_issuers = 1;            // two issuers measured: 64-channel packed layers 0.29 -> 0.26 ms, nothing on the direct path
static int g_rowreuse = -1;          // -1 automatic, 0 / 1 forced
static int g_bres = 1;               // resident weights allowed

}  // namespace afcm

using namespace afcm;

// Debug aid (not part of the stable ABI): enables progress markers written by the tcgen05 kernel into
// mapped host memory, readable even after a failed launch.  Returns the host pointer (64 words).
extern "C" void* afcm_conv_tc_debug_buffer(int enable)
{
    g_dbg_mode = enable >> 8;
    if (!(enable & 1)) { g_dbg_dev = nullptr; return g_dbg_host; }
    if (!g_dbg_host) {
        if (cudaHostAlloc((void**)&g_dbg_host, 64 * sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess) return nullptr;
        memset(g_dbg_host, 0, 64 * sizeof(unsigned));
    }
    if (cudaHostGetDevicePointer((void**)&g_dbg_dev, g_dbg_host, 0) != cudaSuccess) { g_dbg_dev = nullptr; return nullptr; }
    return g_dbg_host;
}

// Tuning aid (not part of the stable ABI): force the TMA->MMA ring depth (0 = automatic).
extern "C" int afcm_conv_tc_set_stages(int stages) { g_force_stages = stages; return AFCM_OK; }
extern "C" int afcm_conv_tc_set_issuers(int n) { g_issuers = n == 2 ? 2 : 1; return AFCM_OK; }
// -1 automatic, 0 per-tap pipeline, 1 row-reuse (resident weights where they fit), 2 row-reuse with streamed weights,
// 4 row-reuse with separate A / B rings for every layer
extern "C" int afcm_conv_tc_set_rowreuse(int mode)
{
    g_rowreuse = mode < 0 ? -1 : (mode == 4 ? 2 : (mode ? 1 : 0));
    g_bres = mode != 2;
    return AFCM_OK;
}

extern "C" int64_t afcm_conv_tc_plane_elems(int H, int W, int Ci)
{
    return (int64_t)H * (W + 2) * ((Ci + 7) & ~7);           // [H*(W+2) flat pixels][Ci padded to 8]
}

extern "C" int afcm_conv_tc_pack_pitched(const void* x, int x_dtype, int x_pitch, const float* icoef, void* xp, int tc_dtype,
                                         int N, int Ci, int H, int W, void* stream)
{
    AFCM_CHECK_ARG(x && xp && N > 0 && Ci > 0 && H > 0 && W > 0, "empty problem");
    AFCM_CHECK_ARG(x_dtype == AFCM_F32 || x_dtype == AFCM_F16, "x must be float32 or float16");
    AFCM_CHECK_ARG(tc_dtype == AFCM_F16 || tc_dtype == AFCM_BF16, "tc dtype must be F16 or BF16");
    AFCM_CHECK_ARG(N <= 65535, "batch too large");
    AFCM_CHECK_ARG(x_pitch >= W, "the row pitch must be at least the width");
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == AFCM_F32) {
        if (tc_dtype == AFCM_BF16) launch_pack<float, __nv_bfloat16>(x, x_pitch, icoef, xp, N, Ci, H, W, st);
        else launch_pack<float, __half>(x, x_pitch, icoef, xp, N, Ci, H, W, st);
    } else {
        if (tc_dtype == AFCM_BF16) launch_pack<__half, __nv_bfloat16>(x, x_pitch, icoef, xp, N, Ci, H, W, st);
        else launch_pack<__half, __half>(x, x_pitch, icoef, xp, N, Ci, H, W, st);
    }
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_conv_tc_pack(const void* x, int x_dtype, const float* icoef, void* xp, int tc_dtype,
                                 int N, int Ci, int H, int W, void* stream)
{
    return afcm_conv_tc_pack_pitched(x, x_dtype, W, icoef, xp, tc_dtype, N, Ci, H, W, stream);
}

static int conv2d_tc_launch(const void* xp, const void* xn, const float* icoef, const void* w_tc, const float* ocoef, const float* bias, void* y,
                            int y_dtype, int tc_dtype, int N, int Ci, int H, int W, int Co, int pad, void* stream);

extern "C" int afcm_conv2d_tc(const void* xp, const void* w_tc, const float* ocoef, const float* bias, void* y, int y_dtype, int tc_dtype,
                              int N, int Ci, int H, int W, int Co, int pad, void* stream)
{
    AFCM_CHECK_ARG(xp, "xp must be given");
    return conv2d_tc_launch(xp, nullptr, nullptr, w_tc, ocoef, bias, y, y_dtype, tc_dtype, N, Ci, H, W, Co, pad, stream);
}

// The same convolution reading the fp16 NCHW activations directly (no packed copy): y[n,o] = ocoef[n,o] * conv(icoef[n,i] * x[n,i], w) + bias[o].
// Full padding (2) only; W even and x 4-byte aligned (the planes are read as aligned pixel pairs); fp16 operands.
static long long* g_tc_trace = nullptr;
extern "C" int afcm_conv_tc_trace(void* dev_buffer) { g_tc_trace = (long long*)dev_buffer; return AFCM_OK; }

static int g_asw_pitch = 0;        // row pitch of the planes handed to conv2d_tc_launch by afcm_conv2d_tc_nchw (W or W + 2)

extern "C" int afcm_conv2d_tc_nchw(const void* x, int x_pitch, const float* icoef, const void* w_tc, const float* ocoef, const float* bias, void* y,
                                   int y_dtype, int N, int Ci, int H, int W, int Co, void* stream)
{
    AFCM_CHECK_ARG(x, "x must be given");
    AFCM_CHECK_ARG(x_pitch == W || x_pitch == W + 2, "x_pitch must be W (dense planes) or W + 2 (planes with two zero pad pixels per row)");
    // planes are read as aligned pixel pairs; the raw tiles arrive by TMA from the flat [N][Ci][H W] view (16-byte aligned strides)
    if ((W & 1) || (((long long)H * x_pitch) & 7) || ((uintptr_t)x & 15)) {
        set_error("conv2d_tc_nchw: needs an even W, H * x_pitch a multiple of 8 and a 16-byte aligned x");
        return AFCM_ERR_UNSUPPORTED;
    }
    g_asw_pitch = x_pitch;
    if ((long long)H * (W + 2) + 4LL * (W + 2) >= (1LL << 30)) { set_error("conv2d_tc_nchw: plane too large"); return AFCM_ERR_UNSUPPORTED; }
    return conv2d_tc_launch(nullptr, x, icoef, w_tc, ocoef, bias, y, y_dtype, AFCM_F16, N, Ci, H, W, Co, 2, stream);
}

static int conv2d_tc_launch(const void* xp, const void* xn, const float* icoef, const void* w_tc, const float* ocoef, const float* bias, void* y,
                            int y_dtype, int tc_dtype, int N, int Ci, int H, int W, int Co, int pad, void* stream)
{
    AFCM_CHECK_ARG(w_tc && y, "w_tc and y must be given");
    const bool asw = xn != nullptr;
    AFCM_CHECK_ARG(y_dtype == AFCM_F32 || y_dtype == AFCM_F16, "y must be float32 or float16");
    AFCM_CHECK_ARG(N > 0 && Ci > 0 && Co > 0 && H > 0 && W > 0, "empty problem");
    AFCM_CHECK_ARG(tc_dtype == AFCM_F16 || tc_dtype == AFCM_BF16, "tc dtype must be F16 or BF16");
    if (pad < 0 || pad > 2) { set_error("conv2d_tc: padding %d not supported (0, 1 or 2)", pad); return AFCM_ERR_UNSUPPORTED; }
    if (H + 2 * pad - 2 <= 0 || W + 2 * pad - 2 <= 0) { set_error("conv2d_tc: empty output"); return AFCM_ERR_INVALID; }
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.ocoef = ocoef; p.bias = bias; p.y = y; p.y_half = y_dtype == AFCM_F16; p.dbg = g_dbg_dev; p.dbg_mode = asw ? 0 : g_dbg_mode;
    p.xn = (const __half*)xn; p.icoef = icoef; p.trace = g_tc_trace;
    p.N = N; p.Ci = Ci; p.Co = Co; p.H = H; p.W = W; p.Wp = W + 2; p.pad = pad;
    p.OH = H + 2 * pad - 2; p.OW = W + 2 * pad - 2;
    p.n_tiles = ceil_div(Co, 256);
    p.BN = ((ceil_div(Co, p.n_tiles) + 31) / 32) * 32;
    p.m_tiles = ceil_div((long long)p.OH * p.Wp, TC_BM);
    p.cblocks = ceil_div(Ci, TC_BK);
    p.nk_last = ceil_div(Ci - (p.cblocks - 1) * TC_BK, 16);
    const long long total = (long long)N * p.m_tiles * p.n_tiles;
    AFCM_CHECK_ARG(total <= 0x7fffffffLL, "too many tiles");
    p.total_tiles = (int)total;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A/B = F16|BF16, A and B K-major, N, M
    const unsigned fmt = tc_dtype == AFCM_BF16 ? 1u : 0u;
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (0u << 15) | (0u << 16) | ((unsigned)(p.BN >> 3) << 17) | ((unsigned)(TC_BM >> 4) << 24);

    const int co_pad = (Co + 15) & ~15, ci_pad = (Ci + 63) & ~63;
    const uint64_t c_pad = (uint64_t)((Ci + 7) & ~7), rows = (uint64_t)H * p.Wp;
    CUtensorMap map_a, map_b;
    int rc = 0;
    if (!asw) {
        rc = encode_3d(&map_a, tc_dtype, xp, c_pad, rows, (uint64_t)N, c_pad * 2, rows * c_pad * 2, TC_BK, TC_BM);
        if (rc) return rc;
    }
    rc = encode_3d(&map_b, tc_dtype, w_tc, (uint64_t)ci_pad, (uint64_t)co_pad, 9, (uint64_t)ci_pad * 2, (uint64_t)ci_pad * 2 * co_pad,
                   TC_BK, (uint32_t)p.BN);
    if (rc) return rc;

    // one operand tile per kernel ROW serves its three taps (a third of the TMA loads, barrier waits and stages of the per-tap
    // scheme: measured 2.3x faster on the small channel tiles).  Large tiles keep the per-tap stages (B dominates).
    // BN <= 192: combined stages (A row + its three weight tiles); larger tiles: separate A / B rings (mode 2)
    p.rowreuse = (g_rowreuse >= 0 && !asw) ? g_rowreuse : (p.BN <= 192 ? 1 : 2);
    if (p.rowreuse == 1 && p.BN > 192) p.rowreuse = 2;
    p.a_stages = 3;
    if (asw) {
        // RAW map: dims {H pitch, Ci, N}, box {TC_RAW_PX, 64, 1}, no swizzle; out-of-range elements / channels are zero-filled
        const uint64_t hw = (uint64_t)H * g_asw_pitch;
        p.pitched = g_asw_pitch != W;
        const uint64_t dims[3] = {hw, (uint64_t)Ci, (uint64_t)N}, strides[2] = {hw * 2, hw * 2 * (uint64_t)Ci};
        const uint32_t box[3] = {(uint32_t)TC_RAW_PX, (uint32_t)TC_BK, 1};
        rc = encode_tiled(&map_a, AFCM_F16, xn, 3, dims, strides, box, false);
        if (rc) return rc;
    }
    CUtensorMap map_a2 = map_a;
    if (p.rowreuse && !asw) {
        rc = encode_3d(&map_a2, tc_dtype, xp, c_pad, rows, (uint64_t)N, c_pad * 2, rows * c_pad * 2, TC_BK, TC_AROW_PX);
        if (rc) return rc;
    }
    const int fixed = 512 + 4 * 256 * 4 + 1024;
    const int b_bytes = p.BN * TC_BK * 2;
    // a single channel tile whose 9 x cblocks weight tiles fit next to a 3-deep A ring: keep the weights resident
    const int res_bytes = 9 * p.cblocks * b_bytes;
    const int raw_min = asw ? 3 * TC_RAW_BYTES : 0;
    p.bres = p.rowreuse && g_bres != 0 && p.n_tiles == 1 &&
             res_bytes + 3 * ((asw && p.pitched) ? TC_AROWP_BYTES : TC_AROW_BYTES) + fixed + raw_min <= max_smem_optin();
    p.bres = p.bres && p.rowreuse == 1;
    int stages, smem;
    if (!asw) {
        const int front = p.rowreuse == 2 ? p.a_stages * TC_AROW_BYTES : (p.bres ? res_bytes : 0);
        const int stage_bytes = p.rowreuse == 2 ? b_bytes : (p.bres ? TC_AROW_BYTES : (p.rowreuse ? TC_AROW_BYTES + 3 * b_bytes : TC_A_BYTES + b_bytes));
        stages = (max_smem_optin() - fixed - front) / stage_bytes;
        if (g_force_stages > 0) stages = g_force_stages;
        if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
        smem = stages * stage_bytes + fixed + front;
    } else {
        // ASW: three raw tiles in flight carry the global-memory latency; the ring of transposed A tiles must be deep enough to
        // hide the release -> refill -> arrive round trip between the tensor core and the producer warps (~700 clocks) behind
        // the MMA time of the stages in between, which is short for small channel tiles (384 clocks per stage at BN = 64:
        // measured 0.79 ms instead of 0.53 ms on the 64-channel layers with a 3-deep ring); the weight ring keeps >= 4 stages
        const int avail = max_smem_optin() - fixed;
        const int arow_asw = p.pitched ? TC_AROWP_BYTES : TC_AROW_BYTES;       // pitched planes: 144-pixel tiles from an 8-pixel boundary
        p.raw_stages = 3;
        if (p.bres) {
            stages = (avail - res_bytes - p.raw_stages * TC_RAW_BYTES) / arow_asw;               // A ring
            if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
            smem = res_bytes + stages * arow_asw;
        } else if (p.rowreuse == 1 && (avail - 2 * TC_RAW_BYTES) / (arow_asw + 3 * b_bytes) >= 2) {
            // combined stages (A row + its three weight tiles), as the packed path uses for channel tiles up to 192: one barrier
            // pair per kernel row instead of four (measured on the packed path: separate rings cost 15-40 % on these layers)
            const int stage_bytes = arow_asw + 3 * b_bytes;
            stages = (avail - p.raw_stages * TC_RAW_BYTES) / stage_bytes;
            if (stages < 2) { p.raw_stages = 2; stages = (avail - p.raw_stages * TC_RAW_BYTES) / stage_bytes; }
            if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
            smem = stages * stage_bytes;
        } else {
            // the A tiles and the weight tiles get rings of their own
            p.rowreuse = 2;
            stages = 4;                                                                               // weight ring
            auto ast = [&]() { return (avail - stages * b_bytes - p.raw_stages * TC_RAW_BYTES) / arow_asw; };
            if (ast() < 2) p.raw_stages = 2;
            p.a_stages = ast() > 8 ? 8 : ast();
            if (p.a_stages < 2) stages = 0;
            else {
                stages = (avail - p.a_stages * arow_asw - p.raw_stages * TC_RAW_BYTES) / b_bytes;   // leftover to the weights
                if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
            }
            smem = p.a_stages * arow_asw + stages * b_bytes;
        }
        if (p.raw_stages > TC_RAW_MAX_STAGES) p.raw_stages = TC_RAW_MAX_STAGES;
        if (p.raw_stages < 2) stages = 0;
        smem += fixed + p.raw_stages * TC_RAW_BYTES;
    }
    if (stages < 2) { set_error("conv2d_tc: not enough shared memory for the pipeline"); return AFCM_ERR_UNSUPPORTED; }
    p.stages = stages;
    p.issuers = (p.rowreuse != 2 && g_issuers == 2 && stages >= 4) ? 2 : 1;   // each issuer needs at least a double-buffered half
    int grid = sm_count();
    if (grid > p.total_tiles) grid = p.total_tiles;
    if (asw) {
        AFCM_CUDA(cudaFuncSetAttribute(conv2d_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        conv2d_tc_kernel<true><<<grid, TC_ASW_THREADS, smem, (cudaStream_t)stream>>>(map_a, map_b, map_a2, p);
    } else {
        AFCM_CUDA(cudaFuncSetAttribute(conv2d_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        conv2d_tc_kernel<false><<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(map_a, map_b, map_a2, p);
    }
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}
