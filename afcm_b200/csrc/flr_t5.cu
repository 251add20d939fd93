// flr_t5.cu -- filtered_lrelu on tcgen05 / tensor memory (sm_100a): the Blackwell-native form of the fused
// up-FIR -> gain / leaky ReLU / clamp -> down-FIR chain (reference: models/networks/stylegan3/torch_utils/ops/
// filtered_lrelu.py:121-153 and the kernel it replaces, filtered_lrelu.cu:139-1099).
//
// Every separable FIR pass is a banded-Toeplitz matrix product on the 5th-generation tensor cores, and the image tile
// stays in tensor memory between the passes:
//
//   TMA      X[i, x]     K1 input rows x N1 columns of one plane -> shared memory (SWIZZLE_128B), zero-filled outside the plane
//   P1 (f16) D1[v, x]  = sum_i Tuy[v, i] X[i, x]     A = the vertical up-sampling Toeplitz matrix, resident in TMEM;
//                                                    B = the TMA tile as an MN-major operand.  M = 128 up-sampled rows = the
//                                                    128 TMEM lanes: from here on a lane is an up-sampled image row.
//   P2 (tf32) D2[v, j] = sum_x D1[v, x] T2[x, j]     A = D1 read IN PLACE from tensor memory (an fp32 accumulator is a valid
//                                                    tf32 operand), B = an 8 x 16U Toeplitz tile of the horizontal up filter;
//                                                    one instruction per 8 input columns, in groups of 64 up-sampled columns
//   E1        A3[v, j] = act(D2[v, j])               epilogue warps: tcgen05.ld -> packed half2 sat(u) - sat(-slope u)
//                                                    (leaky ReLU and both clamps in 3 instructions per pair) -> tcgen05.st in place
//   P3 (f16) D3[v, k] += sum_j A3[v, j] T3[j, k]     A = A3 from tensor memory, B = a 16 x 16 tile of the horizontal down filter
//   E2        ring[v, k] = fp16(D3[v, k])            epilogue warps -> shared-memory ring of down-sampled ROWS (SWIZZLE_128B)
//   P4 (f16) D4[k, w]  = sum_v ring[v, k] T4[v, w]   A = the ring as an MN-major operand (lanes = output columns), B = a
//                                                    16 x 16 tile of the vertical down filter; TMEM columns = output rows
//   E3        y[w, k]  = D4[k, w] (+ skip)           epilogue warps -> global memory
//
// A plane is cut into strips of KW output columns; one (plane, strip) unit streams top to bottom in steps of 128
// up-sampled rows (128/D output rows); the ring carries the FD-1 rows of vertical history between steps.  The Toeplitz
// windows are shift invariant, so the B tiles are built once per CTA and only TMEM / shared-memory addresses move.
// Index algebra (origins, windows, first-touch order of the overlapping accumulator windows): tools/flr_t5_emu.py, pinned
// against the oracle by tests/test_flr_t5_emu.py with the plan computed by t5_make_plan below.
//
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM allocator, warps 4-7 / 8-11 = two
// epilogue groups (each covers the 128 lanes).  Numerics: fp16 operands, tf32 for D1, fp32 accumulation; stated tolerance
// 2e-3 of max|y| per call (tests/test_gpu_flr_t5.py), the same as the mma.sync kernel flr_tc.cu.
#include <cuda.h>
#include <cuda_fp16.h>
#include "afcm_common.cuh"
#include "tc_ptx.cuh"
#include "flr_t5_plan.h"

namespace afcm {

constexpr int T5_THREADS = 384;
constexpr int T5_STAGES = 3;
constexpr int T5_RING_ROWS = 256;
constexpr int T5_RING_HALF = T5_RING_ROWS * 128;        // bytes of one 64-column block of the ring
constexpr int T5_TMEM_COLS = 512;
constexpr int T5_OUT_BYTES = 64 * 128 * 2;              // staging tile of the output rows of one step
constexpr int T5_C_TUY = 0, T5_C_D1 = 40, T5_C_D2 = 168, T5_C_D3 = 296, T5_C_D4 = 432;

struct T5Params {
    T5Plan pl;
    void* y; const void* skip;
    long long ys_n, ys_c;
    int ys_h, y_f32;
    int C, total_units;
    float slope, out_scale;
    long long* trace;                                   // development aid: per-role (code, clock) records of CTA 0, or null
    float kux[24], kuy[24], kdx[24], kdy[24];           // correlation-form taps (scales folded in, see the launcher)
};

enum { T5B_IN_FULL = 0, T5B_IN_EMPTY = 3, T5B_D1_FULL = 6, T5B_P2_DONE = 7, T5B_D2_FULL = 8, T5B_A3_FULL = 10, T5B_P3_DONE = 12,
       T5B_D3_FULL = 14, T5B_E2_DONE = 15, T5B_D4_FULL = 16, T5B_D4_EMPTY = 17, T5B_COUNT = 18 };

__device__ __forceinline__ float t5_tap(const float* k, int n, int e) { return (e >= 0 && e < n) ? k[e] : 0.f; }
// byte offset inside a K-major / MN-major SWIZZLE_128B tile whose rows are 128 bytes
__device__ __forceinline__ uint32_t t5_sw(int row, int byte) { return (uint32_t)(row * 128 + ((((byte >> 4) ^ (row & 7)) << 4) | (byte & 15))); }
__device__ __forceinline__ uint32_t t5_desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3fff) | (((lbo_bytes >> 4) & 0x3fff) << 16); }
__device__ __forceinline__ uint32_t t5_pack(float lo, float hi)
{
    uint32_t h;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(hi), "f"(lo));
    return h;
}
__device__ __forceinline__ uint32_t t5_act(uint32_t lo_bits, uint32_t hi_bits, uint32_t one, uint32_t nslope)
{
    uint32_t h = t5_pack(__uint_as_float(lo_bits), __uint_as_float(hi_bits)), s0, s1, o;
    asm("mul.rn.sat.f16x2 %0, %1, %2;" : "=r"(s0) : "r"(h), "r"(one));
    asm("mul.rn.sat.f16x2 %0, %1, %2;" : "=r"(s1) : "r"(h), "r"(nslope));
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(o) : "r"(s0), "r"(s1));
    return o;
}

// Timeline records of CTA 0 (afcm_filtered_lrelu_t5_trace): role r (0 TMA, 1 MMA, 2 / 3 epilogue groups) appends one word per
// event (code << 48 | clock64) to trace[r * 2048 ...]; the host zeroes the buffer beforehand.
constexpr int T5_TRACE_SLOTS = 2048;
struct T5Trace {
    long long* base; int n;
    __device__ T5Trace(long long* t, int role) : base(t && blockIdx.x == 0 ? t + role * T5_TRACE_SLOTS : nullptr), n(0) {}
    __device__ __forceinline__ void mark(int code)
    {
        if (base && n < T5_TRACE_SLOTS) base[n++] = ((long long)code << 48) | (clock64() & 0xffffffffffffLL);
    }
};

template <int U, int D>
__global__ void __launch_bounds__(T5_THREADS, 1)
flr_t5_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ T5Params p)
{
    constexpr int FU = 6 * U, FD = 6 * D;
    constexpr int K1 = 128 / U + 16, NKC1 = K1 / 16, RS = 128 / U, OS = 128 / D;
    constexpr int ADV = 16 / D;                       // accumulator column advance per 16-wide K chunk (P3 and P4)
    constexpr int NL = (FD - 1 + 15) / 16;            // lead chunks of P4 (rows of the previous step)
    constexpr int HW2 = 8 * U;                        // half window of P2
    constexpr int QPG = 8 / U;                        // P2 chunks per group of 64 up-sampled columns
    constexpr int T2_BYTES = 16 * U * 128, ZERO_BYTES = 4096, T3_BYTES = 2048, T4_BYTES = 2048;

    extern __shared__ uint8_t t5_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(t5_smem_raw) + 1023) & ~(uintptr_t)1023);
    const T5Plan& pl = p.pl;
    const int halves = pl.halves;
    const int stage_bytes = halves * K1 * 128;
    uint8_t* s_in = smem;
    uint8_t* s_ring = s_in + T5_STAGES * stage_bytes;
    uint8_t* s_t2 = s_ring + 2 * T5_RING_HALF;
    uint8_t* s_zero = s_t2 + T2_BYTES;
    uint8_t* s_t3 = s_zero + ZERO_BYTES;
    uint8_t* s_t4 = s_t3 + T3_BYTES;
    __half* s_out = reinterpret_cast<__half*>(s_t4 + (NL + 1) * T4_BYTES);     // [OS output rows][128 columns] staging of E3
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_out) + T5_OUT_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + T5B_COUNT);
    float* s_taps = reinterpret_cast<float*>(tmem_slot + 2);          // [4][24]: kux, kuy, kdx, kdy

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup: barriers, TMEM, the constant B tiles in shared memory, Tuy in tensor memory
    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_x) : "memory");
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < T5_STAGES; i++) { mbar_init(&bars[T5B_IN_FULL + i], 1); mbar_init(&bars[T5B_IN_EMPTY + i], 1); }
        mbar_init(&bars[T5B_D1_FULL], 1); mbar_init(&bars[T5B_P2_DONE], 1);
        for (int b = 0; b < 2; b++) { mbar_init(&bars[T5B_D2_FULL + b], 1); mbar_init(&bars[T5B_A3_FULL + b], 128); mbar_init(&bars[T5B_P3_DONE + b], 1); }
        mbar_init(&bars[T5B_D3_FULL], 1); mbar_init(&bars[T5B_E2_DONE], 256);
        mbar_init(&bars[T5B_D4_FULL], 1); mbar_init(&bars[T5B_D4_EMPTY], 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(T5_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 96; i += T5_THREADS) {
        const int a = i / 24, e = i - a * 24;
        s_taps[i] = a == 0 ? p.kux[e] : (a == 1 ? p.kuy[e] : (a == 2 ? p.kdx[e] : p.kdy[e]));
    }
    for (int i = threadIdx.x; i < ZERO_BYTES / 4; i += T5_THREADS) reinterpret_cast<uint32_t*>(s_zero)[i] = 0u;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    {
        const float* kux = s_taps; const float* kdx = s_taps + 48; const float* kdy = s_taps + 72;
        // T2[k][n] = kux[U k - n + FU - 1], tf32 (round to nearest), row n = 128 bytes, k = 8 x 4 bytes
        for (int i = threadIdx.x; i < 16 * U * 8; i += T5_THREADS) {
            const int n = i >> 3, k = i & 7;
            float v = t5_tap(kux, FU, U * k - n + FU - 1);
            uint32_t t;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
            *reinterpret_cast<uint32_t*>(s_t2 + t5_sw(n, k * 4)) = t;
        }
        // T3[k][n] = kdx[k - D n + t3_e];  T4_e[k][n] = kdy[k - D n + t4_e - 16 e]   (fp16, k = 16 x 2 bytes)
        for (int i = threadIdx.x; i < 256 * (NL + 2); i += T5_THREADS) {
            const int tile = i >> 8, n = (i >> 4) & 15, k = i & 15;
            const float v = tile == 0 ? t5_tap(kdx, FD, k - D * n + pl.t3_e) : t5_tap(kdy, FD, k - D * n + pl.t4_e - 16 * (tile - 1));
            uint8_t* base = tile == 0 ? s_t3 : s_t4 + (tile - 1) * T4_BYTES;
            *reinterpret_cast<__half*>(base + t5_sw(n, k * 2)) = __float2half_rn(v);
        }
    }
    if (warp >= 4 && warp < 8) {
        // Tuy[m][k] = kuy[tuy_e + U k - m] into tensor memory: lane m = this thread, 16 k per 8 packed columns
        const float* kuy = s_taps + 24;
        const int m = (warp & 3) * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + T5_C_TUY;
#pragma unroll
        for (int kc = 0; kc < NKC1; kc++) {
            uint32_t r[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = 16 * kc + 2 * j;
                r[j] = t5_pack(t5_tap(kuy, FU, pl.tuy_e + U * k - m), t5_tap(kuy, FU, pl.tuy_e + U * (k + 1) - m));
            }
            tmem_st8(taddr + 8 * kc, r);
        }
        tmem_st_wait();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int nsteps = pl.nsteps, NG = pl.NG;

    if (warp == 0) {
        // ================================================= TMA producer =================================================
        if (lane == 0) {
            T5Trace tr(p.trace, 0);
            int stage = 0; uint32_t phase = 0;
            for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
                const int plane = unit / pl.nstrips, strip = unit - plane * pl.nstrips;
                const int n = plane / p.C, c = plane - n * p.C;
                const int x0 = pl.iorg0 + pl.istep * strip;
                for (int s = 0; s < nsteps; s++) {
                    mbar_wait(&bars[T5B_IN_EMPTY + stage], phase ^ 1);
                    tr.mark(1);
                    mbar_expect_tx(&bars[T5B_IN_FULL + stage], (uint32_t)stage_bytes);
                    for (int h = 0; h < halves; h++)
                        tma_load_4d(s_in + stage * stage_bytes + h * K1 * 128, &map_x, &bars[T5B_IN_FULL + stage], x0 + 64 * h,
                                    pl.I0y + RS * s, c, n);
                    tr.mark(2);
                    if (++stage == T5_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================= MMA issuer ===================================================
        // The whole warp runs the control flow and the barrier waits; one elected lane issues (see tc_ptx.cuh: elect_one).
        const uint32_t id_p1 = (1u << 4) | (1u << 16) | ((uint32_t)(pl.N1 >> 3) << 17) | (8u << 24);          // B MN-major
        const uint32_t id_p2f = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * HW2) >> 3) << 17) | (8u << 24);   // tf32
        const uint32_t id_p2h = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HW2 >> 3) << 17) | (8u << 24);
        const uint32_t id_p3 = (1u << 4) | (2u << 17) | (8u << 24);
        const uint32_t id_p4 = (1u << 4) | (1u << 15) | (2u << 17) | (8u << 24);                              // A MN-major
        const uint32_t in_base = smem_u32(s_in), ring_base = smem_u32(s_ring);
        const uint32_t t2_lo = desc_lo(smem_u32(s_t2)), t2_hi_lo = desc_lo(smem_u32(s_t2) + HW2 * 128), zero_lo = desc_lo(smem_u32(s_zero));
        const uint32_t t3_lo = desc_lo(smem_u32(s_t3)), t4_lo = desc_lo(smem_u32(s_t4));
        const uint32_t tm_tuy = tmem_base + T5_C_TUY, tm_d1 = tmem_base + T5_C_D1, tm_d2 = tmem_base + T5_C_D2;
        const uint32_t tm_d3 = tmem_base + T5_C_D3, tm_d4 = tmem_base + T5_C_D4;
        uint32_t T = 0, nb[2] = {0u, 0u};
        int stage = 0; uint32_t in_phase = 0;
        T5Trace tr(lane == 0 ? p.trace : nullptr, 1);

        // P4 of global step Tp (its rows sit in ring half Tp & 1; the lead rows at the end of the other half)
        auto issue_p4 = [&](uint32_t Tp, bool first_step_of_unit) {
            mbar_wait(&bars[T5B_E2_DONE], Tp & 1);
            if (Tp >= 1) mbar_wait(&bars[T5B_D4_EMPTY], (Tp - 1) & 1);
            tc_fence_after();
            tr.mark(40);
            if (elect_one()) {
                const uint32_t rows = ring_base + (uint32_t)((Tp & 1) * 128 * 128);
                const uint32_t prev = ring_base + (uint32_t)(((Tp & 1) ^ 1) * 128 * 128);
                // init set: the windows that tile the accumulator columns overwrite
#pragma unroll
                for (int i = 0; i < 8; i += D)
                    umma_f16_ss2(tm_d4 + ADV * i, t5_desc_lo(rows + i * 2048, T5_RING_HALF), TC_DESC_HI, t4_lo, TC_DESC_HI, id_p4, false);
                if (!first_step_of_unit) {
#pragma unroll
                    for (int e = 1; e <= NL; e++)
                        umma_f16_ss2(tm_d4, t5_desc_lo(prev + (8 - e) * 2048, T5_RING_HALF), TC_DESC_HI, t4_lo + (uint32_t)((e * T4_BYTES) >> 4),
                                     TC_DESC_HI, id_p4, true);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (i % D == 0) continue;
                    umma_f16_ss2(tm_d4 + ADV * i, t5_desc_lo(rows + i * 2048, T5_RING_HALF), TC_DESC_HI, t4_lo, TC_DESC_HI, id_p4, true);
                }
                umma_commit(&bars[T5B_D4_FULL]);
            }
            __syncwarp();
            tr.mark(41);
        };
        // P3 of group gg of the current step
        auto issue_p3 = [&](int gg, bool last) {
            const int bb = gg & 1;
            mbar_wait(&bars[T5B_A3_FULL + bb], (nb[bb] - 1) & 1);
            if (gg == 0 && T > 0) mbar_wait(&bars[T5B_E2_DONE], (T - 1) & 1);          // D3 of the previous step has been read
            tc_fence_after();
            tr.mark(30 + gg);
            if (elect_one()) {
                const uint32_t a3 = tm_d2 + 64 * bb, d3 = tm_d3 + ADV * 4 * gg;
                if (gg == 0) umma_f16_ts(tm_d3, a3, zero_lo, TC_DESC_HI, id_p3, false);
                if (D == 2) {
                    umma_f16_ts(d3 + ADV * 1, a3 + 8, t3_lo, TC_DESC_HI, id_p3, false);
                    umma_f16_ts(d3 + ADV * 0, a3 + 0, t3_lo, TC_DESC_HI, id_p3, true);
                    umma_f16_ts(d3 + ADV * 3, a3 + 24, t3_lo, TC_DESC_HI, id_p3, false);
                    umma_f16_ts(d3 + ADV * 2, a3 + 16, t3_lo, TC_DESC_HI, id_p3, true);
                } else {
                    umma_f16_ts(d3 + ADV * 3, a3 + 24, t3_lo, TC_DESC_HI, id_p3, false);
                    umma_f16_ts(d3 + ADV * 0, a3 + 0, t3_lo, TC_DESC_HI, id_p3, true);
                    umma_f16_ts(d3 + ADV * 1, a3 + 8, t3_lo, TC_DESC_HI, id_p3, true);
                    umma_f16_ts(d3 + ADV * 2, a3 + 16, t3_lo, TC_DESC_HI, id_p3, true);
                }
                umma_commit(&bars[T5B_P3_DONE + bb]);
                if (last) umma_commit(&bars[T5B_D3_FULL]);
            }
            __syncwarp();
            tr.mark(35 + gg);
        };

        bool prev_first = false;
        for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
            for (int s = 0; s < nsteps; s++) {
                // ---- P1: D1 = Tuy x (input rows of this step)
                mbar_wait(&bars[T5B_IN_FULL + stage], in_phase);
                if (T > 0) mbar_wait(&bars[T5B_P2_DONE], (T - 1) & 1);                 // D1 of the previous step has been read
                tc_fence_after();
                tr.mark(10);
                if (elect_one()) {
                    const uint32_t sb = in_base + (uint32_t)(stage * stage_bytes);
#pragma unroll
                    for (int kc = 0; kc < NKC1; kc++)
                        umma_f16_ts(tm_d1, tm_tuy + 8 * kc, t5_desc_lo(sb + kc * 2048, K1 * 128), TC_DESC_HI, id_p1, kc > 0);
                    umma_commit(&bars[T5B_IN_EMPTY + stage]);
                    umma_commit(&bars[T5B_D1_FULL]);
                }
                __syncwarp();
                if (++stage == T5_STAGES) { stage = 0; in_phase ^= 1; }
                tr.mark(11);
                mbar_wait(&bars[T5B_D1_FULL], T & 1);
                tc_fence_after();
                tr.mark(12);
                // ---- groups of 64 up-sampled columns: P2 (g), P3 (g - 1); P4 of the previous step after the second P2
                for (int g = 0; g < NG; g++) {
                    const int b = g & 1;
                    if (nb[b] > 0) { mbar_wait(&bars[T5B_P3_DONE + b], (nb[b] - 1) & 1); tc_fence_after(); }
                    tr.mark(20 + g);
                    if (elect_one()) {
                        const uint32_t d2 = tm_d2 + 64 * b, a1 = tm_d1 + 8 * QPG * g;
                        // init set: lead (upper half of the previous chunk's window; group 0: zeros), odd full chunks, tail (lower half)
                        if (g > 0) umma_tf32_ts(d2, a1 - 8, t2_hi_lo, TC_DESC_HI, id_p2h, false);
                        else umma_tf32_ts(d2, a1, zero_lo, TC_DESC_HI, id_p2h, false);
#pragma unroll
                        for (int j = 1; j < QPG - 1; j += 2) umma_tf32_ts(d2 + HW2 * j, a1 + 8 * j, t2_lo, TC_DESC_HI, id_p2f, false);
                        umma_tf32_ts(d2 + 64 - HW2, a1 + 8 * (QPG - 1), t2_lo, TC_DESC_HI, id_p2h, false);
#pragma unroll
                        for (int j = 0; j < QPG - 1; j += 2) umma_tf32_ts(d2 + HW2 * j, a1 + 8 * j, t2_lo, TC_DESC_HI, id_p2f, true);
                        umma_commit(&bars[T5B_D2_FULL + b]);
                        if (g == NG - 1) umma_commit(&bars[T5B_P2_DONE]);
                    }
                    __syncwarp();
                    tr.mark(25 + g);
                    nb[b]++;
                    if (T > 0 && g == (NG > 1 ? 1 : 0)) issue_p4(T - 1, prev_first);
                    if (g >= 1) issue_p3(g - 1, false);
                }
                issue_p3(NG - 1, true);
                prev_first = (s == 0);
                T++;
            }
        }
        if (T > 0) issue_p4(T - 1, prev_first);
    } else if (warp >= 4) {
        // ================================================= epilogue groups ==============================================
        const int wg = (warp - 4) >> 2;                          // 0 / 1
        const int quad = warp & 3;
        const int m = quad * 32 + lane;                          // TMEM lane of this thread
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t h_one = 0x3c003c00u;
        uint32_t h_nslope;
        { __half2 t = __floats2half2_rn(-p.slope, -p.slope); h_nslope = *reinterpret_cast<uint32_t*>(&t); }
        uint32_t T = 0, nuse = 0;
        int pv_plane = 0, pv_strip = 0, pv_s = 0;
        T5Trace tr((quad == 0 && lane == 0) ? p.trace : nullptr, 2 + wg);

        // E3: D4 (lanes = output columns, TMEM columns = output rows) -> transposed fp16 staging tile in shared memory -> rows
        // written with 8-byte stores (a warp covers a whole row of the strip); the skip tensor is added on the way out.
        const int ew = warp - 4;
        const bool vec_ok = !p.y_f32 && ((p.ys_h | (int)(p.ys_c & 3) | (int)(p.ys_n & 3)) & 3) == 0 && (((uintptr_t)p.y | (uintptr_t)p.skip) & 7) == 0 &&
                            (pl.KW & 3) == 0 && (pl.yw & 3) == 0;
        auto e3 = [&](uint32_t Tp, int plane, int strip, int s) {
            mbar_wait(&bars[T5B_D4_FULL], Tp & 1);
            tc_fence_after();
            tr.mark(70);
            constexpr int NC = OS / 2;                            // D4 columns (output rows) per group: 32 or 16
            uint32_t r[NC];
            if constexpr (NC == 32) tmem_ld32(lane_base + T5_C_D4 + NC * wg, r);
            else tmem_ld16(lane_base + T5_C_D4 + NC * wg, r);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars[T5B_D4_EMPTY]);
            asm volatile("bar.sync 1, 256;" ::: "memory");       // the staging tile of the previous step has been read
            const int col = m - pl.m0;
            if (col >= 0 && col < pl.KW) {
                __half* so = s_out + (NC * wg) * 128 + col;
#pragma unroll
                for (int j = 0; j < NC; j++) so[j * 128] = __float2half_rn(__uint_as_float(r[j]));
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int k0 = pl.KW * strip;
            const int kwv = min(pl.KW, pl.yw - k0);
            const int n = plane / p.C, c = plane - n * p.C;
            const long long pofs = n * p.ys_n + c * p.ys_c + k0;
            const int cc = 4 * lane;
            if (cc < kwv) {
                for (int rr = ew; rr < OS; rr += 8) {
                    const int w = OS * s + pl.wlo0 + rr;
                    if (w < 0 || w >= pl.yh) continue;
                    const long long o = pofs + (long long)w * p.ys_h + cc;
                    const uint2 v = *reinterpret_cast<const uint2*>(s_out + rr * 128 + cc);
                    if (vec_ok) {
                        uint2 outv = v;
                        if (p.skip) {
                            const uint2 sk = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.skip) + o);
                            const __half2 sc = __floats2half2_rn(p.out_scale, p.out_scale);
                            __half2 a0 = __hfma2(*reinterpret_cast<const __half2*>(&sk.x), sc, *reinterpret_cast<const __half2*>(&v.x));
                            __half2 a1 = __hfma2(*reinterpret_cast<const __half2*>(&sk.y), sc, *reinterpret_cast<const __half2*>(&v.y));
                            outv.x = *reinterpret_cast<uint32_t*>(&a0); outv.y = *reinterpret_cast<uint32_t*>(&a1);
                        }
                        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.y) + o) = outv;
                    } else {
                        const __half* hv = reinterpret_cast<const __half*>(&v);
                        for (int e = 0; e < 4 && cc + e < kwv; e++) {
                            float f = __half2float(hv[e]);
                            if (p.y_f32) {
                                if (p.skip) f += reinterpret_cast<const float*>(p.skip)[o + e] * p.out_scale;
                                reinterpret_cast<float*>(p.y)[o + e] = f;
                            } else {
                                if (p.skip) f += __half2float(reinterpret_cast<const __half*>(p.skip)[o + e]) * p.out_scale;
                                reinterpret_cast<__half*>(p.y)[o + e] = __float2half_rn(f);
                            }
                        }
                    }
                }
            }
        };

        for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
            const int plane = unit / pl.nstrips, strip = unit - plane * pl.nstrips;
            for (int s = 0; s < nsteps; s++) {
                // ---- E1: activation of this group's D2 buffers (in place: 64 fp32 columns -> 32 packed half2 columns); the output
                // rows of the previous step (E3, its P4 is issued right after the second P2 of this step) go out after the first one
                bool e3_done = T == 0;
                for (int g = wg; g < NG; g += 2) {
                    mbar_wait(&bars[T5B_D2_FULL + wg], nuse & 1);
                    tc_fence_after();
                    tr.mark(50 + g);
                    const uint32_t d2 = lane_base + T5_C_D2 + 64 * wg;
                    uint32_t ra[32], rb[32], o[16];
                    tmem_ld32(d2, ra);
                    tmem_ld32(d2 + 32, rb);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) o[i] = t5_act(ra[2 * i], ra[2 * i + 1], h_one, h_nslope);
                    tmem_st16(d2, o);
#pragma unroll
                    for (int i = 0; i < 16; i++) o[i] = t5_act(rb[2 * i], rb[2 * i + 1], h_one, h_nslope);
                    tmem_st16(d2 + 16, o);
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&bars[T5B_A3_FULL + wg]);
                    tr.mark(55 + g);
                    nuse++;
                    if (!e3_done) { e3(T - 1, pv_plane, pv_strip, pv_s); tr.mark(71); e3_done = true; }
                }
                if (!e3_done) { e3(T - 1, pv_plane, pv_strip, pv_s); tr.mark(71); }
                // ---- E2: D3 columns [64 wg, 64 wg + 64) of this step -> ring rows (fp16, SWIZZLE_128B rows of 64 columns)
                mbar_wait(&bars[T5B_D3_FULL], T & 1);
                tc_fence_after();
                tr.mark(60);
                {
                    const int row = (int)((T & 1) * 128) + m;
                    uint8_t* rbase = s_ring + wg * T5_RING_HALF;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        uint32_t r[32];
                        tmem_ld32(lane_base + T5_C_D3 + 64 * wg + 32 * h, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            uint4 v;
                            v.x = t5_pack(__uint_as_float(r[8 * q + 0]), __uint_as_float(r[8 * q + 1]));
                            v.y = t5_pack(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3]));
                            v.z = t5_pack(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5]));
                            v.w = t5_pack(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7]));
                            *reinterpret_cast<uint4*>(rbase + t5_sw(row, (32 * h + 8 * q) * 2)) = v;
                        }
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&bars[T5B_E2_DONE]);
                tr.mark(61);
                pv_plane = plane; pv_strip = strip; pv_s = s;
                T++;
            }
        }
        if (T > 0) e3(T - 1, pv_plane, pv_strip, pv_s);
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T5_TMEM_COLS) : "memory");
    }
}

static long long* g_t5_trace = nullptr;

template <int U, int D>
static int launch_t5(const CUtensorMap& map, const T5Params& p, cudaStream_t st)
{
    constexpr int K1 = 128 / U + 16, NL = (6 * D - 1 + 15) / 16;
    const int smem = 1024 + T5_STAGES * p.pl.halves * K1 * 128 + 2 * T5_RING_HALF + 16 * U * 128 + 4096 + 2048 + (NL + 1) * 2048 +
                     T5_OUT_BYTES + T5B_COUNT * 8 + 16 + 96 * 4;
    if (smem > max_smem_optin()) { set_error("filtered_lrelu_t5: %d bytes of shared memory needed", smem); return AFCM_ERR_UNSUPPORTED; }
    AFCM_CUDA(cudaFuncSetAttribute(flr_t5_kernel<U, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));     // per device: not cached
    int blocks = p.total_units < sm_count() ? p.total_units : sm_count();
    flr_t5_kernel<U, D><<<blocks, T5_THREADS, smem, st>>>(map, p);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

}  // namespace afcm

using namespace afcm;

// Development aid (not part of the stable ABI): device buffer of 4 * 2 * 1024 int64 that CTA 0 of the following launches
// fills with (code, clock64) records per role (0 TMA, 1 MMA issuer, 2 / 3 epilogue groups; [0] = count); NULL switches it off.
extern "C" int afcm_filtered_lrelu_t5_trace(void* dev_buffer) { g_t5_trace = (long long*)dev_buffer; return AFCM_OK; }

extern "C" int afcm_filtered_lrelu_t5_plan(int xh, int xw, int up, int down, int px0, int px1, int py0, int py1, int kw, int* out, int n_out)
{
    T5Plan pl;
    int rc = t5_make_plan(xh, xw, up, down, px0, px1, py0, py1, kw, &pl);
    if (rc) return rc;
    const int n = (int)(sizeof(T5Plan) / sizeof(int));
    AFCM_CHECK_ARG(out && n_out >= n, "plan buffer needs %d ints", n);
    memcpy(out, &pl, sizeof(T5Plan));
    return AFCM_OK;
}

extern "C" int afcm_filtered_lrelu_t5(const void* x, const int64_t* xs, int x_dtype, void* y, const int64_t* ys, int y_dtype,
                                      const float* b, const void* skip,
                                      int N, int C, int xh, int xw, int yh, int yw,
                                      const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                                      int up, int down, int px0, int px1, int py0, int py1,
                                      float gain, float slope, float clamp, float out_scale, int flip_filter,
                                      void* stream)
{
    AFCM_CHECK_ARG(x && y && xs && ys, "x, y and their strides must be given");
    AFCM_CHECK_ARG(N > 0 && C > 0 && xh > 0 && xw > 0, "x is empty");
    const bool geo_ok = (up == 2 && down == 2) || (up == 4 && down == 2) || (up == 2 && down == 4);
    const bool clamp_ok = clamp >= 1.f / 1024.f && clamp <= 1024.f;
    if (!geo_ok || !fu_host || !fd_host || fu_taps != 6 * up || fd_taps != 6 * down || b || x_dtype != AFCM_F16 ||
        (y_dtype != AFCM_F16 && y_dtype != AFCM_F32) || !clamp_ok || !(slope >= 0.f) || !(gain > 0.f) || xs[3] != 1 || ys[3] != 1) {
        set_error("filtered_lrelu_t5: unsupported call (up=%d/%d taps, down=%d/%d taps; needs fp16 input, no bias, a clamp in [2^-10, 2^10])",
                  up, fu_taps, down, fd_taps);
        return AFCM_ERR_UNSUPPORTED;
    }
    // TMA: 16-byte aligned base and strides
    if (((uintptr_t)x & 15) || ((xs[0] * 2) & 15) || ((xs[1] * 2) & 15) || ((xs[2] * 2) & 15) || xs[2] < xw) {
        set_error("filtered_lrelu_t5: x needs a 16-byte aligned base address and row / plane / sample strides that are multiples of 8 elements");
        return AFCM_ERR_UNSUPPORTED;
    }
    T5Params p;
    memset(&p, 0, sizeof(p));
    int rc = t5_make_plan(xh, xw, up, down, px0, px1, py0, py1, 0, &p.pl);
    if (rc) return rc;
    AFCM_CHECK_ARG(p.pl.yh == yh && p.pl.yw == yw, "y has shape [%d,%d], expected [%d,%d]", yh, yw, p.pl.yh, p.pl.yw);
    AFCM_CHECK_ARG(ys[2] >= 0 && (long long)(yh + 1) * ys[2] < (1ll << 31), "row stride out of range");
    p.y = y; p.skip = skip; p.ys_n = ys[0]; p.ys_c = ys[1]; p.ys_h = (int)ys[2]; p.y_f32 = y_dtype == AFCM_F32;
    p.C = C;
    const long long units = (long long)N * C * p.pl.nstrips;
    AFCM_CHECK_ARG(units < 0x7fffffffLL, "too many strips");
    p.total_units = (int)units;
    p.slope = slope; p.out_scale = out_scale;
    p.trace = g_t5_trace;
    // the activation runs in units of `clamp`:  clamp(lrelu(u gain)) / clamp = sat(u') - sat(-slope u'),  u' = u gain / clamp
    const float u_scale = 1.f / clamp;
    for (int t = 0; t < fu_taps; t++) {
        const float f = fu_host[flip_filter ? t : fu_taps - 1 - t] * (float)up;
        p.kux[t] = f;
        p.kuy[t] = f * gain * u_scale;
    }
    for (int t = 0; t < fd_taps; t++) {
        const float f = fd_host[flip_filter ? t : fd_taps - 1 - t];
        p.kdx[t] = f * out_scale;
        p.kdy[t] = f / u_scale;
    }
    CUtensorMap map;
    {
        const uint64_t dims[4] = {(uint64_t)xw, (uint64_t)xh, (uint64_t)C, (uint64_t)N};
        const uint64_t strides[3] = {(uint64_t)xs[2] * 2, (uint64_t)xs[1] * 2, (uint64_t)xs[0] * 2};
        const uint32_t box[4] = {64, (uint32_t)p.pl.K1, 1, 1};
        rc = encode_tiled(&map, AFCM_F16, x, 4, dims, strides, box);
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (up == 2 && down == 2) return launch_t5<2, 2>(map, p, st);
    if (up == 4 && down == 2) return launch_t5<4, 2>(map, p, st);
    return launch_t5<2, 4>(map, p, st);
}
