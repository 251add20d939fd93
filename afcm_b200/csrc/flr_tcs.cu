// flr_tcs.cu -- filtered_lrelu on the tensor cores WITH the sign tensor: the training-step variant (forward with sign
// write, backward with sign read) of the op that flr_tc.cu serves for inference.
//
// Reference semantics: models/networks/stylegan3/torch_utils/ops/filtered_lrelu.py:121-153 (forward), :252-266
// (backward = the same op with up/down exchanged and the activation replaced by the sign mask), sign tensor format
// filtered_lrelu.cpp:87-94 / filtered_lrelu.cu:494-520 -- bit-compatible with afcm_filtered_lrelu (filtered_lrelu.cu).
//
// One CTA = one 32x32 (down 2) or 16x16 (down 4) output tile of one (n,c) plane, four passes through shared memory,
// each a set of independent m16n8k16 tensor-core products of a DATA tile with a constant banded-Toeplitz tile of the
// FIR taps (16-bit operands, fp32 accumulation):
//   pass 1  hup   T1[iy, ux] = sum_ix X[iy, ix]  ku[c0x + U ix - ux]     A = data (ldmatrix),        B = taps (registers)
//   pass 2  vup   T2[uy, ux] = sum_iy ku[c0y + U iy - uy] T1[iy, ux]     A = taps,                   B = data (ldmatrix.trans)
//           then gain, leaky ReLU, clamp in fp32 on the accumulator fragment, sign codes written to / read from the
//           the reference's sign tensor layout (forward: packed inside lane quads and stored directly; backward: staged
//           through shared memory)
//   pass 3  hdown T3[uy, ox] = sum_ux T2[uy, ux] kd[ux - D ox]           A = data,                   B = taps
//   pass 4  vdown  Y[oy, ox] = sum_uy kd[uy - D oy] T3[uy, ox]           A = taps,                   B = data (.trans)
// The Toeplitz tiles are shift invariant because every block origin advances by a fixed ratio (8 input columns <-> 8U
// up-sampled columns, 16/U input rows <-> 16 up-sampled rows, 8D up-sampled columns <-> 8 outputs, ...), so each thread
// keeps a few constant fragments per pass.  All window starts are multiples of 8 elements in the contiguous dimension
// (ldmatrix needs 16-byte aligned rows); row pitches are 8 * odd elements (conflict-free ldmatrix).
// Numerics: x (+ bias) and the three intermediates are rounded to the operand type (fp16 forward: same class of error
// as flr_tc, <= 2e-3 of max|y|; bf16 backward: gradients keep the fp32 exponent range, ~3 significant digits); the
// activation and every sum are fp32.  The sign of an up-sampled value within rounding distance of zero may differ from
// the exact kernel's; forward and backward of THIS kernel are consistent with each other (the backward applies the mask
// that the forward stored).
#include "afcm_common.cuh"
#include "filtered_lrelu_core.h"

namespace afcm {

constexpr int TS_THREADS = 256;

constexpr int ts_pitch(int w) { int p = (w + 7) / 8; return (p % 2 ? p : p + 1) * 8; }      // 8 * odd >= w
constexpr int ts_max(int a, int b) { return a > b ? a : b; }

template <int U, int D> struct TsGeo {
    static constexpr int FU = 6 * U, FD = 6 * D;
    static constexpr int TW = (D == 4) ? 16 : 32, TH = TW;          // output tile
    static constexpr int KKH = (7 * D + FD + 15) / 16;              // k16 steps per hdown product (8 outputs)
    static constexpr int KKV = (15 * D + FD + 15) / 16;             // k16 steps per vdown product (16 outputs)
    static constexpr int UWP = 8 * D * (TW / 8 - 1) + 16 * KKH;     // up-sampled columns read by hdown
    static constexpr int UHP = 16 * D * (TH / 16 - 1) + 16 * KKV;   // up-sampled rows read by vdown
    static constexpr int NWIN = (UWP + 8 * U - 1) / (8 * U);        // hup windows: 8 input columns -> 8U up-sampled columns
    static constexpr int UWC = NWIN * 8 * U;                        // up-sampled columns computed (>= UWP, multiple of 16)
    static constexpr int IW = 8 * (NWIN - 1) + 16;                  // input columns held
    static constexpr int MTV = UHP / 16;                            // vup row tiles (16 up-sampled rows <- 16/U + FU/U input rows)
    static constexpr int IH = (16 / U) * (MTV - 1) + 16;            // input rows read by vup
    static constexpr int IHP = (IH + 15) / 16 * 16;                 // input rows computed by hup
    static constexpr int P_X = ts_pitch(IW), P_T1 = ts_pitch(UWC), P_T2 = ts_pitch(UWC), P_T3 = ts_pitch(TW);
    static constexpr int A_ELEMS = ts_max(IHP * P_X, UHP * P_T2);   // buffer A: X, then T2
    static constexpr int B_ELEMS = ts_max(IHP * P_T1, UHP * P_T3);  // buffer B: T1, then T3
    static constexpr int UWT = (TW - 1) * D + FD, UHT = (TH - 1) * D + FD;   // up-sampled extent the outputs consume
    static_assert(UWC % 16 == 0 && UWP <= UWC && UHP % 16 == 0 && UWT <= UWP && UHT <= UHP, "tile geometry");
};

template <typename TH> struct TsOps;
template <> struct TsOps<__half> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) { __half2 h = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
    static __device__ __forceinline__ __half cvt(float v) { return __float2half_rn(v); }
    static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
    {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};
template <> struct TsOps<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
    static __device__ __forceinline__ __nv_bfloat16 cvt(float v) { return __float2bfloat16_rn(v); }
    static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
    {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};

__device__ __forceinline__ void ts_ldsm(uint32_t (&r)[4], const void* p)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ts_ldsm_t(uint32_t (&r)[4], const void* p)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ float ts_tap(const float* k, int n, int idx) { return (idx >= 0 && idx < n) ? k[idx] : 0.f; }

template <int U, int D, typename TH, int SIGN>
__global__ void __launch_bounds__(TS_THREADS)
flr_tcs_kernel(const __grid_constant__ FlrParams p)
{
    using G = TsGeo<U, D>;
    using Op = TsOps<TH>;
    extern __shared__ __align__(16) unsigned char ts_smem[];
    TH* bufA = reinterpret_cast<TH*>(ts_smem);                 // X, then T2
    TH* bufB = bufA + G::A_ELEMS;                              // T1, then T3
    uint8_t* s_sign = reinterpret_cast<uint8_t*>(bufB + G::B_ELEMS);
    // taps in shared memory: the fragment constants index them per lane, and a divergent index into the kernel-parameter
    // constant bank is replayed once per distinct address (measured: 20x slower than the whole rest of the kernel)
    __shared__ float s_ku[FLR_MAX_TAPS], s_kd[FLR_MAX_TAPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < FLR_MAX_TAPS) { s_ku[tid] = p.ku[tid]; s_kd[tid] = p.kd[tid]; }
    const int g = lane >> 2, t4 = lane & 3;
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lcol = (lane >> 4) * 8;     // ldmatrix.x4 address pattern
    const FlrTile t = flr_tile3<U, D>(p, blockIdx.x, blockIdx.y, blockIdx.z);
    // first input sample held, and the tap phase: up-sampled sample j = ux0 + ux depends on input i = ibx + ix through
    // tap ku[U i + px0 - j] = ku[c0x + U ix - ux]
    const int ibx = flr_floor_div(t.ux0 - p.px0 + U - 1, U), iby = flr_floor_div(t.uy0 - p.py0 + U - 1, U);
    const int c0x = U * ibx + p.px0 - t.ux0, c0y = U * iby + p.py0 - t.uy0;

    // ---- pass 0: global -> X (bias on real samples only, zero outside the plane), sign staging (backward) ----
    // Items are column PAIRS of a 64-column grid (power-of-two decode, one packed store per pair); columns >= P_X idle.
    {
        static_assert(G::P_X <= 64 && G::P_X % 2 == 0, "load pass assumes at most 64 columns");
        const float* xp = (const float*)p.x + t.n * p.xs_n + t.c * p.xs_c;
        const float bias = p.b ? ((const float*)p.b)[t.c] : 0.f;
        const int sh = (int)p.xs_h, sw = (int)p.xs_w;
        const float* x0 = xp + iby * sh + ibx * sw;                 // sample (0,0) of the tile (may lie outside the plane)
        constexpr int NIT = (G::IHP * 32 + TS_THREADS - 1) / TS_THREADS;
        // All global loads of the thread are issued before the first shared-memory store: interleaved, every store is a
        // possible alias of the next load for the compiler, and the pass ran as NIT dependent load -> store round trips
        // (57 % of the kernel's stall samples, profiles/r01_ncu_flr_tcs.txt).
        float v0[NIT], v1[NIT];
        // tiles whose input window lies inside the plane (most of them on the large planes) need only the tile-shape tests
        const bool inner = iby >= 0 && iby + G::IH <= p.xh && ibx >= 0 && ibx + G::IW <= p.xw;
        if (inner) {
            const float* xt = x0 + (tid >> 5) * sh + (tid & 31) * 2 * sw;
#pragma unroll
            for (int it = 0; it < NIT; it++) {
                const int iy = (tid >> 5) + it * (TS_THREADS / 32), ix = (tid & 31) * 2;
                const float* src = xt + it * (TS_THREADS / 32) * sh;
                v0[it] = (iy < G::IH && ix < G::IW) ? __ldg(src) + bias : 0.f;
                v1[it] = (iy < G::IH && ix + 1 < G::IW) ? __ldg(src + sw) + bias : 0.f;
            }
        } else {
#pragma unroll
            for (int it = 0; it < NIT; it++) {
                const int i = tid + it * TS_THREADS;
                const int iy = i >> 5, ix = (i & 31) * 2;
                const bool rok = iy < G::IH && (unsigned)(iby + iy) < (unsigned)p.xh;
                const float* src = x0 + iy * sh + ix * sw;
                const bool ok0 = rok && ix < G::IW && (unsigned)(ibx + ix) < (unsigned)p.xw;
                const bool ok1 = rok && ix + 1 < G::IW && (unsigned)(ibx + ix + 1) < (unsigned)p.xw;
                v0[it] = ok0 ? __ldg(src) + bias : 0.f;
                v1[it] = ok1 ? __ldg(src + sw) + bias : 0.f;
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int i = tid + it * TS_THREADS;
            const int iy = i >> 5, ix = (i & 31) * 2;
            if (iy < G::IHP && ix < G::P_X) *reinterpret_cast<uint32_t*>(bufA + iy * G::P_X + ix) = Op::pack(v0[it], v1[it]);
        }
        if (SIGN == 2) {
            // the tile's part of the packed sign tensor as aligned 32-bit words: staged row = 8 words = 32 bytes starting at
            // the sign byte eb0a (a multiple of 4) of sign row uy0 + ly + s_oy; words outside the tensor read as 0 = "leave
            // the value alone" (the tensor's row pitch s_wb is a multiple of 4, so a word is inside or outside as a whole)
            const int eb0a = flr_floor_div(t.ux0 + p.s_ox, 16) * 4;
            const uint8_t* base = p.si + (long long)t.plane * p.s_h * p.s_wb;
            uint32_t* sw32 = reinterpret_cast<uint32_t*>(s_sign);
            for (int i = tid; i < p.uht * 8; i += TS_THREADS) {
                const int ly = i >> 3, ey = t.uy0 + ly + p.s_oy, eb = eb0a + (i & 7) * 4;
                uint32_t v = 0;
                if ((unsigned)ey < (unsigned)p.s_h && (unsigned)eb < (unsigned)p.s_wb) v = *reinterpret_cast<const uint32_t*>(base + ey * p.s_wb + eb);
                sw32[i] = v;
            }
        }
    }
    __syncthreads();

    // ---- pass 1: horizontal up-FIR.  X[IHP][IW] -> T1[IHP][UWC] ----
    {
        uint32_t bu[U][2];
#pragma unroll
        for (int j = 0; j < U; j++)
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int k = 2 * t4 + 8 * r, n = 8 * j + g;
                bu[j][r] = Op::pack(ts_tap(s_ku, G::FU, c0x + U * k - n), ts_tap(s_ku, G::FU, c0x + U * (k + 1) - n));
            }
        constexpr int NJ = (G::IHP / 16) * G::NWIN;
        for (int job = warp; job < NJ; job += TS_THREADS / 32) {
            const int mt = job / G::NWIN, w = job - mt * G::NWIN;
            uint32_t a[4];
            ts_ldsm(a, bufA + (16 * mt + lrow) * G::P_X + 8 * w + lcol);
#pragma unroll
            for (int j = 0; j < U; j++) {
                float d[4] = {0.f, 0.f, 0.f, 0.f};
                Op::mma(d, a, bu[j][0], bu[j][1]);
                TH* dst = bufB + (16 * mt + g) * G::P_T1 + 8 * U * w + 8 * j + 2 * t4;
                *reinterpret_cast<uint32_t*>(dst) = Op::pack(d[0], d[1]);
                *reinterpret_cast<uint32_t*>(dst + 8 * G::P_T1) = Op::pack(d[2], d[3]);
            }
        }
    }
    __syncthreads();

    // ---- pass 2: vertical up-FIR + gain + leaky ReLU + clamp (+ signs).  T1[IH][UWC] -> T2[UHP][UWC] ----
    // The gain rides on the vertical taps.  Forward: sign codes are packed inside each quad of lanes (8 consecutive columns
    // of one row = 2 bytes) and stored straight into the sign tensor; a tile writes the codes of its own 16*... up-sampled
    // block only (the halo belongs to the neighbours), the last tile row / column also the rest of the tensor.
    {
        uint32_t av[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int m = g + (r & 1) * 8, k = 2 * t4 + (r >> 1) * 8;
            av[r] = Op::pack(ts_tap(s_ku, G::FU, c0y + U * k - m) * p.gain, ts_tap(s_ku, G::FU, c0y + U * (k + 1) - m) * p.gain);
        }
        constexpr int nbp = 32;                                     // staged sign bytes per row (backward)
        const int se0 = SIGN == 2 ? t.ux0 + p.s_ox - 16 * flr_floor_div(t.ux0 + p.s_ox, 16) : 0;   // 0..15
        // forward: extent of the codes this tile owns, and the sign-tensor origin of the tile
        const int own_w = (t.ox0 + G::TW >= p.yw) ? p.uwt : G::TW * D, own_h = (t.oy0 + G::TH >= p.yh) ? p.uht : G::TH * D;
        uint8_t* so = SIGN == 1 ? p.so + ((long long)t.plane * p.s_h + t.uy0) * p.s_wb + (t.ux0 >> 2) : nullptr;
        const int s_bytes = p.s_wb - (t.ux0 >> 2), row_lim = min(own_h, p.s_h - t.uy0);
        const float nclamp = -p.clamp;
        constexpr int NB2 = G::UWC / 16, NJ = G::MTV * NB2;
        for (int job = warp; job < NJ; job += TS_THREADS / 32) {
            const int mt = job / NB2, nb = job - mt * NB2;
            uint32_t b[4];
            ts_ldsm_t(b, bufB + ((16 / U) * mt + lrow) * G::P_T1 + 16 * nb + lcol);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float d[4] = {0.f, 0.f, 0.f, 0.f};
                Op::mma(d, av, b[2 * h], b[2 * h + 1]);
                const int lx = 16 * nb + 8 * h + 2 * t4;
                const bool col_st = SIGN == 1 && !(t4 & 1) && lx < own_w && (lx >> 2) < s_bytes;   // this lane stores the byte
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int ly = 16 * mt + g + 8 * rr;
                    float v0 = d[2 * rr], v1 = d[2 * rr + 1];
                    if (SIGN == 2) {
                        if (ly < p.uht && lx < p.uwt) {          // lx + 1 <= uwt - 1 as well: lx and uwt are even
                            const uint8_t* sr = s_sign + ly * nbp;
                            const int e0 = se0 + lx, e1 = e0 + 1;
                            const int c0 = sr[e0 >> 2] >> ((e0 & 3) << 1), c1 = sr[e1 >> 2] >> ((e1 & 3) << 1);
                            if (c0 & 1) v0 *= p.slope;
                            if (c0 & 2) v0 = 0.f;
                            if (c1 & 1) v1 *= p.slope;
                            if (c1 & 2) v1 = 0.f;
                        }
                    } else {
                        const bool n0 = v0 < 0.f, n1 = v1 < 0.f;
                        if (n0) v0 *= p.slope;
                        if (n1) v1 *= p.slope;
                        const float w0 = fminf(fmaxf(v0, nclamp), p.clamp), w1 = fminf(fmaxf(v1, nclamp), p.clamp);
                        if (SIGN == 1) {
                            // 2-bit codes: 1 = negative, 2 = clamped (overrides); 4 bits per lane, one byte per lane pair
                            const int c0 = (w0 != v0) ? 2 : (int)n0, c1 = (w1 != v1) ? 2 : (int)n1;
                            int c = c0 | (c1 << 2);
                            c |= __shfl_xor_sync(0xffffffffu, c, 1) << 4;
                            if (col_st && ly < row_lim) so[ly * p.s_wb + (lx >> 2)] = (uint8_t)c;
                        }
                        v0 = w0; v1 = w1;
                    }
                    *reinterpret_cast<uint32_t*>(bufA + ly * G::P_T2 + lx) = Op::pack(v0, v1);
                }
            }
        }
    }
    __syncthreads();

    // ---- pass 3: horizontal down-FIR with decimation.  T2[UHP][UWP] -> T3[UHP][TW] ----
    {
        uint32_t bd[G::KKH][2];
#pragma unroll
        for (int kk = 0; kk < G::KKH; kk++)
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int k = 16 * kk + 2 * t4 + 8 * r;
                bd[kk][r] = Op::pack(ts_tap(s_kd, G::FD, k - D * g), ts_tap(s_kd, G::FD, k + 1 - D * g));
            }
        constexpr int NBO = G::TW / 8, NJ = (G::UHP / 16) * NBO;
        for (int job = warp; job < NJ; job += TS_THREADS / 32) {
            const int mt = job / NBO, nb = job - mt * NBO;
            float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kk = 0; kk < G::KKH; kk++) {
                uint32_t a[4];
                ts_ldsm(a, bufA + (16 * mt + lrow) * G::P_T2 + 8 * D * nb + 16 * kk + lcol);
                Op::mma(d, a, bd[kk][0], bd[kk][1]);
            }
            TH* dst = bufB + (16 * mt + g) * G::P_T3 + 8 * nb + 2 * t4;
            *reinterpret_cast<uint32_t*>(dst) = Op::pack(d[0], d[1]);
            *reinterpret_cast<uint32_t*>(dst + 8 * G::P_T3) = Op::pack(d[2], d[3]);
        }
    }
    __syncthreads();

    // ---- pass 4: vertical down-FIR with decimation and the global store.  T3[UHP][TW] -> y ----
    {
        uint32_t ad[G::KKV][4];
#pragma unroll
        for (int kk = 0; kk < G::KKV; kk++)
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int m = g + (r & 1) * 8, k = 16 * kk + 2 * t4 + (r >> 1) * 8;
                ad[kk][r] = Op::pack(ts_tap(s_kd, G::FD, k - D * m), ts_tap(s_kd, G::FD, k + 1 - D * m));
            }
        float* yp = (float*)p.y + t.n * p.ys_n + t.c * p.ys_c;
        const float* kp = p.skip ? (const float*)p.skip + t.n * p.ys_n + t.c * p.ys_c : nullptr;
        const int ysh = (int)p.ys_h, ysw = (int)p.ys_w;
        constexpr int NB2 = G::TW / 16, NJ = (G::TH / 16) * NB2;
        for (int job = warp; job < NJ; job += TS_THREADS / 32) {
            const int mt = job / NB2, nb = job - mt * NB2;
            float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kk = 0; kk < G::KKV; kk++) {
                uint32_t b[4];
                ts_ldsm_t(b, bufB + (16 * D * mt + 16 * kk + lrow) * G::P_T3 + 16 * nb + lcol);
                Op::mma(d0, ad[kk], b[0], b[1]);
                Op::mma(d1, ad[kk], b[2], b[3]);
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float* d = h ? d1 : d0;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int gy = t.oy0 + 16 * mt + g + (e >> 1) * 8, gx = t.ox0 + 16 * nb + 8 * h + 2 * t4 + (e & 1);
                    if (gy < p.yh && gx < p.yw) {
                        const int o = gy * ysh + gx * ysw;
                        float acc = d[e];
                        if (kp) acc += kp[o];
                        yp[o] = acc * p.out_scale;
                    }
                }
            }
        }
    }
}

template <int U, int D>
static int launch_tcs(FlrParams& p, int op_dtype, int sign_mode, cudaStream_t stream)
{
    using G = TsGeo<U, D>;
    p.tow = G::TW; p.toh = G::TH;
    p.tiles_x = flr_ceil_div(p.yw, G::TW); p.tiles_y = flr_ceil_div(p.yh, G::TH);
    p.uwt = G::UWT; p.uht = G::UHT;
    const long long planes = (long long)p.N * p.C;
    if (planes > 0x7fffffffLL || p.tiles_x > 65535 || p.tiles_y > 65535) { set_error("filtered_lrelu_tcs: too many tiles"); return AFCM_ERR_INVALID; }
    auto labs64 = [](long long v) { return v < 0 ? -v : v; };
    const long long xspan = (long long)(p.xh + 256) * labs64(p.xs_h) + (long long)(p.xw + 256) * labs64(p.xs_w);
    const long long yspan = (long long)p.yh * labs64(p.ys_h) + (long long)p.yw * labs64(p.ys_w);
    if (xspan > 0x3fffffffLL || yspan > 0x3fffffffLL) { set_error("filtered_lrelu_tcs: plane too large for 32-bit offsets"); return AFCM_ERR_UNSUPPORTED; }
    size_t smem = (size_t)(G::A_ELEMS + G::B_ELEMS) * 2;
    if (sign_mode == AFCM_SIGN_READ) smem += (size_t)p.uht * 32;
    smem = (smem + 15) & ~(size_t)15;
    void (*kern)(const FlrParams) = nullptr;
    const bool bf = op_dtype == AFCM_BF16;
    if (sign_mode == AFCM_SIGN_NONE)  kern = bf ? flr_tcs_kernel<U, D, __nv_bfloat16, 0> : flr_tcs_kernel<U, D, __half, 0>;
    if (sign_mode == AFCM_SIGN_WRITE) kern = bf ? flr_tcs_kernel<U, D, __nv_bfloat16, 1> : flr_tcs_kernel<U, D, __half, 1>;
    if (sign_mode == AFCM_SIGN_READ)  kern = bf ? flr_tcs_kernel<U, D, __nv_bfloat16, 2> : flr_tcs_kernel<U, D, __half, 2>;
    AFCM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)planes, (unsigned)p.tiles_x, (unsigned)p.tiles_y), TS_THREADS, smem, stream>>>(p);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

}  // namespace afcm

using namespace afcm;

extern "C" int afcm_filtered_lrelu_tcs(const void* x, const int64_t* xs, void* y, const int64_t* ys,
                                       const void* b, const void* skip, int op_dtype,
                                       int N, int C, int xh, int xw, int yh, int yw,
                                       const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                                       int up, int down, int px0, int px1, int py0, int py1,
                                       float gain, float slope, float clamp, float out_scale, int flip_filter,
                                       int sign_mode, void* signs, int sign_h, int sign_wb, int sx, int sy,
                                       void* stream)
{
    AFCM_CHECK_ARG(x && y && xs && ys, "x, y and their strides must be given");
    AFCM_CHECK_ARG(op_dtype == AFCM_F16 || op_dtype == AFCM_BF16, "operand type must be F16 or BF16 (x, y, b, skip are float32)");
    AFCM_CHECK_ARG(N > 0 && C > 0 && xh > 0 && xw > 0, "x is empty");
    AFCM_CHECK_ARG(sign_mode >= 0 && sign_mode <= 2, "bad sign_mode");
    AFCM_CHECK_ARG(sign_mode == AFCM_SIGN_NONE || signs, "sign tensor missing");
    if (!(fu_host && fd_host && ((up == 2 && down == 2) || (up == 2 && down == 4) || (up == 4 && down == 2)) &&
          fu_taps == 6 * up && fd_taps == 6 * down)) {
        set_error("filtered_lrelu_tcs: no tensor-core kernel for up=%d/%d taps, down=%d/%d taps", up, fu_taps, down, fd_taps);
        return AFCM_ERR_UNSUPPORTED;
    }
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
    for (int t = 0; t < fu_taps; t++) p.ku[t] = fu_host[flip_filter ? t : fu_taps - 1 - t] * (float)up;
    for (int t = 0; t < fd_taps; t++) p.kd[t] = fd_host[flip_filter ? t : fd_taps - 1 - t];
    cudaStream_t st = (cudaStream_t)stream;
    if (up == 2 && down == 2) return launch_tcs<2, 2>(p, op_dtype, sign_mode, st);
    if (up == 2 && down == 4) return launch_tcs<2, 4>(p, op_dtype, sign_mode, st);
    return launch_tcs<4, 2>(p, op_dtype, sign_mode, st);
}
