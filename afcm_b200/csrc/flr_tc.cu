// flr_tc.cu -- filtered_lrelu on the tensor cores (sm_100a): the four separable FIR passes of
// bias -> up-FIR -> gain/leaky-ReLU/clamp -> down-FIR (reference: models/networks/stylegan3/torch_utils/
// ops/filtered_lrelu.py:121-153, filtered_lrelu.cu:139-1099) evaluated as a chain of small banded-Toeplitz
// matrix products with mma.sync.m16n8k16 (fp16 operands), entirely in registers.
//
// Why: the op does 17-46 FLOP per algorithmic byte, above the FP32-SIMT ridge of B200, so a CUDA-core
// kernel cannot approach the HBM roofline.  Here the MACs run on the tensor pipe,
// the nonlinearity on the (otherwise idle) FMA pipe as packed half2 arithmetic, and the ALU pipe only
// sees address arithmetic and the input conversion.
//
// Algorithm (one warp = one strip of 16 output columns, streamed top to bottom):
//   in[y][x]  --(1) R1[j,y] = sum_x Tu_x[j,x] in[y,x]      const A, data B (global -> regs)
//             --(2) R2[v,j] = sum_y Tu_y[v,y] R1[j,y]      const A, B = D-fragments of (1)   (transposing reuse)
//             --    leaky ReLU + clamp on half2 (see act2 below)
//             --(4) R3[k,v] = sum_j Td_x[k,j] R2[v,j]      const A, B = D-fragments of (2)
//             --(5) R4[k,w] = sum_v R3[k,v] Td_y[w,v]      A = D-fragments of (4), const B, fp32 accumulate
//   j,v = up-sampled coordinates, k,w = output coordinates.  Products (1), (2) and (4) use the f16
//   accumulator form of the instruction: its m16n8 result layout (two packed half2 registers) IS the B
//   (or A) operand layout of the next product, so there is no conversion, shared memory, shuffle or block
//   barrier between passes.  The Toeplitz operands are shift invariant, so each thread keeps a handful of
//   constant fragments (the FIR taps) in registers for the whole strip.  The polyphase structure of the
//   zero-insertion is folded into Tu (only every UP-th column of a row is non-zero).
//   Index conventions (s = phase shift, delta = input alignment, origins) are modelled in
//   tools/flr_tc_model.py and pinned against the oracle by tests/test_flr_tc_model.py.
//
// Activation: with u = x*gain/clamp (the scale is folded into the Tu_y taps),
//       clamp(lrelu(x*gain), +-clamp) / clamp = sat(u) - sat(-slope*u),      sat(.) = clamp to [0,1]
//   which is three half2 instructions on the FMA pipe (HMUL2.SAT x2, HADD2) for two samples, both clamps
//   included; `clamp` is folded back into the Td_x taps.  Without a clamp (or with one too large for the
//   fp16 range of u) the kernel uses max(u, slope*u) and an explicit min/max.
//
// Numerics: operands (activations, intermediates, taps) are rounded to fp16; sums are fp32 inside the
// tensor core and rounded once per product (product (4) rounds after each of its K chunks), the last
// product accumulates in fp32.  Measured against the fp32 oracle: see tests/test_gpu_flr_tc.py for the
// stated bound (2e-3 of max|y| per call).  This is the "tensor-core path with stated tolerance" of the north
// star; afcm_filtered_lrelu remains the exact-fp32 path.
#include <cuda_fp16.h>
#include "afcm_common.cuh"

namespace afcm {

constexpr int FTC_WARPS = 4;
constexpr int FTC_TAB = 192;          // tap tables: index t + 64, zero padded
constexpr int FTC_TAB_OFS = 64;
constexpr int FTC_STAGES = 4;         // per-warp shared-memory ring of input row blocks (cp.async)
constexpr int FTC_PD = 3;             // prefetch distance in row blocks

enum { FTC_ACT_SAT = 0, FTC_ACT_MINMAX = 1 };

struct FlrTcParams {
    const void* x; void* y; const float* b; const void* skip;
    long long xs_n, xs_c, ys_n, ys_c;                    // element strides, innermost stride 1
    int xs_h, ys_h;                                      // row strides (host-checked to fit 32 bits)
    int C, xh, xw, yh, yw;
    int strips, segs, seg_wblocks;                        // 16-column strips, row segments of 8*seg_wblocks rows
    long long total_warps;
    int ix0, iy0, iy_step;                                // input origin: col = strip*IXS + ix0, row = seg*iy_step + iy0
    int sx, sy, dx;
    float slope, out_scale, act_clamp;                    // act_clamp: clamp in the units of R2 (MINMAX mode)
    float kux[24], kuy[24], kdx[24], kdy[24];             // correlation-form taps (kuy, kdx carry gain and the clamp scale)
};

template <int U, int D> struct FtcGeo {
    static constexpr int FU = 6 * U, FD = 6 * D;
    static constexpr int KC4 = (D == 2) ? 3 : 6;          // 16-wide chunks of up-sampled columns per strip
    static constexpr int NJ8 = 2 * KC4;
    static constexpr int NC = (U == 2) ? KC4 + 1 : (KC4 - 1) / 2 + 2;   // 8-wide input column chunks per strip
    static constexpr int NPH = (U == 2) ? 1 : 2;          // distinct up-filter fragments (phases)
    static constexpr int NB5 = (D == 2) ? 2 : 4;          // 16-row chunks of R3 per block of 8 output rows
    static constexpr int IXS = 16 * D / U;                // input columns per strip step
};

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi)
{
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// D = A * B, f16 result (two packed half2 registers: rows g / g+8, columns 2t, 2t+1)
__device__ __forceinline__ void mma_h(uint32_t (&d)[2], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    const uint32_t z = 0u;
    asm("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%8,%8};"
        : "=r"(d[0]), "=r"(d[1])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(z));
}
// D += A * B, f16 accumulator
__device__ __forceinline__ void mma_h_acc(uint32_t (&d)[2], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
        : "+r"(d[0]), "+r"(d[1])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D += A * B, fp32 accumulator
__device__ __forceinline__ void mma_f(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t h2_mul_sat(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("mul.rn.sat.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t h2_sub(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t h2_mul(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t h2_max(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t h2_min(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

// 8- or 4-byte asynchronous global -> shared copy; src_bytes == 0 zero-fills without reading
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src, int src_bytes)
{
    if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T> struct Pair;
template <> struct Pair<float> { typedef float2 type; };
template <> struct Pair<__half> { typedef uint32_t type; };

template <int U, int D, typename TIN, typename TOUT, int ACT>
struct FtcWarp {
    using Geo = FtcGeo<U, D>;
    using RawT = typename Pair<TIN>::type;
    // constant fragments (the FIR taps)
    uint32_t a1[Geo::NPH][4], a2[Geo::NPH][4], a4[Geo::KC4][4], b5[Geo::NB5][2];
    uint32_t h_one, h_nslope, h_slope, h_cl, h_ncl, h_bias;
    const FlrTcParams& p;
    int g, t;
    // strip state
    const TIN* xc0;              // plane base + g rows + first column pair of the strip (chunk c at +8c)
    uint32_t ring;               // shared-memory address of this lane's slot in stage 0, chunk 0
    int ix;                      // first input column of the strip
    TOUT* yt;                    // plane base + (w0 + 2t) rows + k0 + g
    long long kofs;              // skip - y (elements), valid when has_skip
    bool has_skip;
    float bias;
    int iy, k0, w0, nwb;
    bool interior;               // every input column this strip reads and every output column it writes exists
    static constexpr int RAW_BYTES = (int)sizeof(RawT);
    static constexpr int CHUNK_BYTES = 32 * RAW_BYTES;               // one chunk of one row block, all lanes
    static constexpr int STAGE_BYTES = Geo::NC * CHUNK_BYTES;
    static constexpr int WARP_RING_BYTES = FTC_STAGES * STAGE_BYTES;

    __device__ FtcWarp(const FlrTcParams& p_, int lane) : p(p_), g(lane >> 2), t(lane & 3) {}

    __device__ void load_consts(const float (*tab)[FTC_TAB])
    {
        const float* tux = tab[0] + FTC_TAB_OFS; const float* tuy = tab[1] + FTC_TAB_OFS;
        const float* tdx = tab[2] + FTC_TAB_OFS; const float* tdy = tab[3] + FTC_TAB_OFS;
#pragma unroll
        for (int ph = 0; ph < Geo::NPH; ph++) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int row = g + (r & 1) * 8, col = 2 * t + (r >> 1) * 8;
                // Tu[row, col] = ku[(col - delta) * U - 16 * ph - row]
                const int ex = (col - p.dx) * U - 16 * ph - row;
                const int ey = col * U - 16 * ph - row;
                a1[ph][r] = pack_h2(tux[ex], tux[ex + U]);
                a2[ph][r] = pack_h2(tuy[ey], tuy[ey + U]);
            }
        }
#pragma unroll
        for (int kc = 0; kc < Geo::KC4; kc++) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int row = g + (r & 1) * 8, col = 2 * t + (r >> 1) * 8;
                // Td_x[row = k, col = j] = kd[16 kc + col - row * D - sx]
                const int e = 16 * kc + col - row * D - p.sx;
                a4[kc][r] = pack_h2(tdx[e], tdx[e + 1]);
            }
        }
#pragma unroll
        for (int q = 0; q < Geo::NB5; q++) {
#pragma unroll
            for (int r = 0; r < 2; r++) {
                // B[k = v, n = w] = kd[16 q + k - n * D - sy],  k = 2t + 8r + {0,1},  n = g
                const int e = 16 * q + 2 * t + 8 * r - g * D - p.sy;
                b5[q][r] = pack_h2(tdy[e], tdy[e + 1]);
            }
        }
        h_one = pack_h2(1.f, 1.f);
        h_nslope = pack_h2(-p.slope, -p.slope);
        h_slope = pack_h2(p.slope, p.slope);
        h_cl = pack_h2(p.act_clamp, p.act_clamp);
        h_ncl = pack_h2(-p.act_clamp, -p.act_clamp);
    }

    __device__ void begin_strip(long long wid)
    {
        const int strip = (int)(wid % p.strips);
        const long long r = wid / p.strips;
        const int seg = (int)(r % p.segs);
        const long long plane = r / p.segs;
        const int n = (int)(plane / p.C), c = (int)(plane - (long long)n * p.C);
        const TIN* xg = (const TIN*)p.x + n * p.xs_n + c * p.xs_c + (long long)g * p.xs_h;
        has_skip = p.skip != nullptr;
        kofs = has_skip ? (const TOUT*)p.skip - (const TOUT*)p.y : 0;
        bias = p.b ? p.b[c] : 0.f;
        h_bias = pack_h2(bias, bias);
        ix = strip * Geo::IXS + p.ix0;
        iy = seg * p.iy_step + p.iy0;
        k0 = strip * 16;
        w0 = seg * p.seg_wblocks * 8;
        yt = (TOUT*)p.y + n * p.ys_n + c * p.ys_c + (long long)(w0 + 2 * t) * p.ys_h + k0 + g;
        xc0 = xg + (ix + 2 * t);
        const int rows_left = p.yh - w0;
        nwb = (rows_left + 7) >> 3;
        if (nwb > p.seg_wblocks) nwb = p.seg_wblocks;
        // ix and xw are even (host-checked), so a column pair is either inside or outside the plane as a whole
        interior = __all_sync(0xffffffffu, col_mask() == (1u << Geo::NC) - 1u) && k0 + 16 <= p.yw;
    }

    // bit c: this lane's column pair of chunk c lies inside the plane
    __device__ __forceinline__ unsigned col_mask() const
    {
        unsigned m = 0;
#pragma unroll
        for (int c8 = 0; c8 < Geo::NC; c8++) {
            const int col = ix + 8 * c8 + 2 * t;
            if (col >= 0 && col < p.xw) m |= 1u << c8;
        }
        return m;
    }

    // ---- input: issue the asynchronous copies of one block of 8 input rows into the ring (this thread: row g,
    // NC column pairs).  EDGE: rows / columns outside the plane are zero-filled (src size 0, address clamped).
    template <bool EDGE>
    __device__ __forceinline__ void fetch(int yb) const
    {
        const int row0 = iy + 8 * yb;
        const uint32_t dst = ring + (uint32_t)((yb & (FTC_STAGES - 1)) * STAGE_BYTES);
        if (!EDGE) {
            const TIN* rp = xc0 + row0 * p.xs_h;
#pragma unroll
            for (int c8 = 0; c8 < Geo::NC; c8++) cp_async<RAW_BYTES>(dst + c8 * CHUNK_BYTES, rp + 8 * c8, RAW_BYTES);
        } else {
            const int row = row0 + g;
            const bool rok = row >= 0 && row < p.xh;
            const TIN* rp = xc0 + (rok ? row0 : -g) * p.xs_h - (ix + 2 * t);          // start of a valid row
#pragma unroll
            for (int c8 = 0; c8 < Geo::NC; c8++) {
                const int col = ix + 8 * c8 + 2 * t;
                const bool ok = rok && col >= 0 && col < p.xw;
                cp_async<RAW_BYTES>(dst + c8 * CHUNK_BYTES, rp + (ok ? col : 0), ok ? RAW_BYTES : 0);
            }
        }
        cp_async_commit();
    }
    __device__ __forceinline__ uint32_t cvt_pair(const float2& v) const { return pack_h2(v.x + bias, v.y + bias); }
    __device__ __forceinline__ uint32_t cvt_pair(const uint32_t& v) const
    {
        uint32_t r;
        asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(h_bias));
        return r;
    }
    // wait until row block yb has landed (at most FTC_PD - 1 younger groups may still be in flight), read this
    // lane's pairs back, add the bias and round to fp16.  EDGE: pairs outside the plane stay zero (no bias there).
    template <bool EDGE>
    __device__ __forceinline__ void convert(int yb, uint32_t (&in)[Geo::NC]) const
    {
        cp_async_wait<FTC_PD - 1>();
        const uint32_t src = ring + (uint32_t)((yb & (FTC_STAGES - 1)) * STAGE_BYTES);
        RawT raw[Geo::NC];
#pragma unroll
        for (int c8 = 0; c8 < Geo::NC; c8++) {
            if (RAW_BYTES == 8) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(src + c8 * CHUNK_BYTES) : "memory");
                raw[c8] = *reinterpret_cast<RawT*>(&v);
            } else {
                uint32_t v;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(src + c8 * CHUNK_BYTES) : "memory");
                raw[c8] = *reinterpret_cast<RawT*>(&v);
            }
        }
        if (!EDGE) {
#pragma unroll
            for (int c8 = 0; c8 < Geo::NC; c8++) in[c8] = cvt_pair(raw[c8]);
        } else {
            const int row = iy + 8 * yb + g;
            const unsigned m = (row >= 0 && row < p.xh) ? col_mask() : 0u;
#pragma unroll
            for (int c8 = 0; c8 < Geo::NC; c8++) in[c8] = ((m >> c8) & 1u) ? cvt_pair(raw[c8]) : 0u;
        }
    }

    // (1) horizontal up-FIR of one block of 8 input rows: r1[nb] = B-fragment half for column block nb of R2
    __device__ __forceinline__ void step1(const uint32_t (&in)[Geo::NC], uint32_t (&r1)[Geo::NJ8]) const
    {
#pragma unroll
        for (int b = 0; b < Geo::KC4; b++) {
            const int w = (U == 2) ? b : (b >> 1);          // first input chunk of the window
            const int ph = (U == 2) ? 0 : (b & 1);
            uint32_t d[2];
            mma_h(d, a1[ph], in[w], in[w + 1]);
            r1[2 * b] = d[0];
            r1[2 * b + 1] = d[1];
        }
    }

    // leaky ReLU + clamp of two packed samples (see the header)
    __device__ __forceinline__ uint32_t act2(uint32_t u) const
    {
        if (ACT == FTC_ACT_SAT) {
            return h2_sub(h2_mul_sat(u, h_one), h2_mul_sat(u, h_nslope));
        } else {
            const uint32_t s = h2_mul(u, h_slope);
            const uint32_t v = p.slope > 1.f ? h2_min(u, s) : h2_max(u, s);
            return h2_max(h2_min(v, h_cl), h_ncl);
        }
    }

    // (2)+(nonlinearity)+(4): one block of 16 up-sampled rows -> A-fragment chunk of R3 for step (5)
    __device__ __forceinline__ void vblock(const uint32_t (&ra)[Geo::NJ8], const uint32_t (&rb)[Geo::NJ8], int ph,
                                           uint32_t (&a5)[4]) const
    {
        uint32_t c0[2], c1[2];
#pragma unroll
        for (int kc = 0; kc < Geo::KC4; kc++) {
            uint32_t p2[2][2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int nb = 2 * kc + q;
                uint32_t d[2];
                mma_h(d, a2[ph], ra[nb], rb[nb]);
                p2[q][0] = act2(d[0]);
                p2[q][1] = act2(d[1]);
            }
            if (kc == 0) {
                mma_h(c0, a4[kc], p2[0][0], p2[1][0]);
                mma_h(c1, a4[kc], p2[0][1], p2[1][1]);
            } else {
                mma_h_acc(c0, a4[kc], p2[0][0], p2[1][0]);
                mma_h_acc(c1, a4[kc], p2[0][1], p2[1][1]);
            }
        }
        a5[0] = c0[0]; a5[1] = c0[1]; a5[2] = c1[0]; a5[3] = c1[1];
    }

    // out_scale is folded into the Td_y taps; a skip tensor is added as skip * out_scale
    template <bool EDGE>
    __device__ __forceinline__ void store(int wb, const float (&c)[4]) const
    {
        const int o = 8 * wb * p.ys_h;
        TOUT* q0 = yt + o;                                              // (row 2t, col g) of this block
        TOUT* q1 = yt + (o + p.ys_h);
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = c[i];
        if (!EDGE) {
            if (has_skip) {
                v[0] += (float)q0[kofs] * p.out_scale; v[1] += (float)q1[kofs] * p.out_scale;
                v[2] += (float)q0[kofs + 8] * p.out_scale; v[3] += (float)q1[kofs + 8] * p.out_scale;
            }
            q0[0] = (TOUT)v[0]; q1[0] = (TOUT)v[1]; q0[8] = (TOUT)v[2]; q1[8] = (TOUT)v[3];
        } else {
            const int y0 = w0 + 8 * wb + 2 * t;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int yy = y0 + (i & 1), xx = k0 + g + (i >> 1) * 8;
                if (yy < p.yh && xx < p.yw) {
                    TOUT* q = ((i & 1) ? q1 : q0) + (i >> 1) * 8;
                    if (has_skip) v[i] += (float)q[kofs] * p.out_scale;
                    *q = (TOUT)v[i];
                }
            }
        }
    }

    // (5) one block of 8 output rows from NB5 consecutive chunks of R3
    template <bool EDGE>
    __device__ __forceinline__ void wblock2(int wb, const uint32_t (&c0)[4], const uint32_t (&c1)[4]) const
    {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        mma_f(c, c0, b5[0][0], b5[0][1]);
        mma_f(c, c1, b5[1][0], b5[1][1]);
        store<EDGE>(wb, c);
    }
    template <bool EDGE>
    __device__ __forceinline__ void wblock4(int wb, const uint32_t (&c0)[4], const uint32_t (&c1)[4], const uint32_t (&c2)[4],
                                            const uint32_t (&c3)[4]) const
    {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        mma_f(c, c0, b5[0][0], b5[0][1]);
        mma_f(c, c1, b5[1][0], b5[1][1]);
        mma_f(c, c2, b5[Geo::NB5 - 2][0], b5[Geo::NB5 - 2][1]);
        mma_f(c, c3, b5[Geo::NB5 - 1][0], b5[Geo::NB5 - 1][1]);
        store<EDGE>(wb, c);
    }

    // ---- per-geometry pipelines.  Every strip runs  [0, lo) with EDGE handling, [lo, hi) without (all input
    // rows / columns read and all outputs written by those iterations are inside the planes), [hi, iters) with.
    static __device__ __forceinline__ int fdiv8(int a) { return a >> 3; }                 // floor(a / 8)

    // U == 2, D == 2.  iteration it: input block it -> R1[it]; v-block it-1 from R1[it-1], R1[it];
    // w-block it-2 from chunks it-2, it-1.
    template <bool EDGE>
    __device__ __forceinline__ void iter22(int it, uint32_t (&r1p)[Geo::NJ8], uint32_t (&r1c)[Geo::NJ8],
                                           uint32_t (&a5p)[4], uint32_t (&a5c)[4]) const
    {
        uint32_t in[Geo::NC];
        convert<EDGE>(it, in); step1(in, r1c); fetch<EDGE>(it + FTC_PD);
        if (!EDGE || it >= 1) vblock(r1p, r1c, 0, a5c);
        if (!EDGE || it >= 2) wblock2<EDGE>(it - 2, a5p, a5c);
    }
    __device__ void run22()
    {
        uint32_t r1a[Geo::NJ8], r1b[Geo::NJ8], a5a[4], a5b[4];
        const int iters = nwb + 2;
        int lo = 0, hi = 0;
        if (interior) {
            lo = max(2, fdiv8(-iy + 7));
            hi = min(min(fdiv8(p.xh - 8 - iy) - FTC_PD, fdiv8(p.yh - w0) + 1) + 1, iters);
        }
        if (hi < lo) hi = lo = 0;
        hi = lo + ((hi - lo) & ~1);                                     // the steady loop is unrolled by two
        int it = 0;
        for (; it < (hi > lo ? lo : iters); it++) {
            iter22<true>(it, r1a, r1b, a5a, a5b);
#pragma unroll
            for (int i = 0; i < Geo::NJ8; i++) r1a[i] = r1b[i];
#pragma unroll
            for (int i = 0; i < 4; i++) a5a[i] = a5b[i];
        }
        if (hi > lo) {
            for (; it < hi; it += 2) {                                  // "previous" buffers alternate, no copies
                iter22<false>(it, r1a, r1b, a5a, a5b);
                iter22<false>(it + 1, r1b, r1a, a5b, a5a);
            }
            for (; it < iters; it++) {
                iter22<true>(it, r1a, r1b, a5a, a5b);
#pragma unroll
                for (int i = 0; i < Geo::NJ8; i++) r1a[i] = r1b[i];
#pragma unroll
                for (int i = 0; i < 4; i++) a5a[i] = a5b[i];
            }
        }
    }

    // U == 4, D == 2.  iteration it: input block it; v-blocks 2it-2, 2it-1 (two phases) from R1[it-1], R1[it];
    // w-block 2it-3 from chunks (2it-3, 2it-2), w-block 2it-2 from chunks (2it-2, 2it-1)
    template <bool EDGE>
    __device__ __forceinline__ void iter42(int it, uint32_t (&r1p)[Geo::NJ8], uint32_t (&r1c)[Geo::NJ8], uint32_t (&a5l)[4]) const
    {
        uint32_t in[Geo::NC], a5a[4], a5b[4];
        convert<EDGE>(it, in); step1(in, r1c); fetch<EDGE>(it + FTC_PD);
        if (!EDGE || it >= 1) {
            vblock(r1p, r1c, 0, a5a);
            vblock(r1p, r1c, Geo::NPH - 1, a5b);
            if (!EDGE || (it >= 2 && 2 * it - 3 < nwb)) wblock2<EDGE>(2 * it - 3, a5l, a5a);
            if (!EDGE || 2 * it - 2 < nwb) wblock2<EDGE>(2 * it - 2, a5a, a5b);
#pragma unroll
            for (int i = 0; i < 4; i++) a5l[i] = a5b[i];
        }
    }
    __device__ void run42()
    {
        uint32_t r1a[Geo::NJ8], r1b[Geo::NJ8], a5l[4];
        const int iters = (nwb + 2) / 2 + 1;
        int lo = 0, hi = 0;
        if (interior) {
            lo = max(2, fdiv8(-iy + 7));
            hi = min(min(fdiv8(p.xh - 8 - iy) - FTC_PD, (fdiv8(p.yh - w0) + 1) >> 1) + 1, iters);
        }
        if (hi < lo) hi = lo = 0;
        hi = lo + ((hi - lo) & ~1);
        int it = 0;
        for (; it < (hi > lo ? lo : iters); it++) {
            iter42<true>(it, r1a, r1b, a5l);
#pragma unroll
            for (int i = 0; i < Geo::NJ8; i++) r1a[i] = r1b[i];
        }
        if (hi > lo) {
            for (; it < hi; it += 2) {
                iter42<false>(it, r1a, r1b, a5l);
                iter42<false>(it + 1, r1b, r1a, a5l);
            }
            for (; it < iters; it++) {
                iter42<true>(it, r1a, r1b, a5l);
#pragma unroll
                for (int i = 0; i < Geo::NJ8; i++) r1a[i] = r1b[i];
            }
        }
    }

    // U == 2, D == 4.  iteration it: input blocks 2it, 2it+1; v-block 2it-1 from R1[2it-1], R1[2it];
    // v-block 2it from R1[2it], R1[2it+1]; w-block it-2 from chunks 2it-4 .. 2it-1
    template <bool EDGE>
    __device__ __forceinline__ void iter24(int it, uint32_t (&r1l)[Geo::NJ8], uint32_t (&q0)[4], uint32_t (&q1)[4],
                                           uint32_t (&q2)[4]) const
    {
        uint32_t in[Geo::NC], r1a[Geo::NJ8], r1b[Geo::NJ8], qa[4], qb[4];
        convert<EDGE>(2 * it, in); step1(in, r1a); fetch<EDGE>(2 * it + FTC_PD);
        if (!EDGE || it >= 1) vblock(r1l, r1a, 0, qa);
        convert<EDGE>(2 * it + 1, in); step1(in, r1b); fetch<EDGE>(2 * it + 1 + FTC_PD);
        vblock(r1a, r1b, 0, qb);
        if (!EDGE || it >= 2) wblock4<EDGE>(it - 2, q0, q1, q2, qa);
#pragma unroll
        for (int i = 0; i < 4; i++) { q0[i] = q2[i]; q1[i] = qa[i]; q2[i] = qb[i]; }
#pragma unroll
        for (int i = 0; i < Geo::NJ8; i++) r1l[i] = r1b[i];
    }
    __device__ void run24()
    {
        uint32_t r1l[Geo::NJ8], q0[4], q1[4], q2[4];
        const int iters = nwb + 2;
        int lo = 0, hi = 0;
        if (interior) {
            lo = max(2, (-iy + 15) >> 4);
            hi = min(min((fdiv8(p.xh - 8 - iy) - FTC_PD - 1) >> 1, fdiv8(p.yh - w0) + 1) + 1, iters);
        }
        if (hi < lo) hi = lo = 0;
        int it = 0;
        for (; it < (hi > lo ? lo : iters); it++) iter24<true>(it, r1l, q0, q1, q2);
        if (hi > lo) {
            for (; it < hi; it++) iter24<false>(it, r1l, q0, q1, q2);
            for (; it < iters; it++) iter24<true>(it, r1l, q0, q1, q2);
        }
    }

    __device__ void run()
    {
        // the ring is primed with edge handling; these row blocks precede the steady range anyway
#pragma unroll
        for (int i = 0; i < FTC_PD; i++) fetch<true>(i);
        if (U == 2 && D == 2) run22();
        else if (U == 4 && D == 2) run42();
        else run24();
        cp_async_wait<0>();
    }
};

template <int U, int D, typename TIN, typename TOUT, int ACT>
__global__ void __launch_bounds__(FTC_WARPS * 32, (U == 2 && D == 4) ? 3 : (U == 4 ? 5 : 6))
flr_tc_kernel(const __grid_constant__ FlrTcParams p)
{
    __shared__ float tab[4][FTC_TAB];
    for (int i = threadIdx.x; i < 4 * FTC_TAB; i += blockDim.x) {
        const int k = i / FTC_TAB, e = i - k * FTC_TAB - FTC_TAB_OFS;
        const float* src = k == 0 ? p.kux : (k == 1 ? p.kuy : (k == 2 ? p.kdx : p.kdy));
        const int n = k < 2 ? 6 * U : 6 * D;
        tab[k][i - k * FTC_TAB] = (e >= 0 && e < n) ? src[e] : 0.f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * FTC_WARPS + warp;
    if (wid >= p.total_warps) return;
    extern __shared__ __align__(16) uint8_t ring_smem[];
    typedef FtcWarp<U, D, TIN, TOUT, ACT> W;
    W w(p, lane);
    w.ring = (uint32_t)__cvta_generic_to_shared(ring_smem) + (uint32_t)(warp * W::WARP_RING_BYTES + lane * W::RAW_BYTES);
    w.load_consts(tab);
    w.begin_strip(wid);
    w.run();
}

static int floor_mod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

template <int U, int D, typename TIN, typename TOUT, int ACT>
static int launch_tc(FlrTcParams& p, int N, cudaStream_t st)
{
    p.strips = ceil_div(p.yw, 16);
    const long long planes = (long long)N * p.C;
    int wblocks = ceil_div(p.yh, 8);
    if (wblocks & 1) wblocks++;                         // segments are multiples of 16 output rows
    int seg_wblocks = wblocks;
    // enough warps to fill the machine a few times over; long strips amortise the pipeline fill
    while (planes * p.strips * ceil_div(wblocks, seg_wblocks) < 16384 && seg_wblocks > 4) {
        seg_wblocks = ((seg_wblocks / 2) + 1) & ~1;
    }
    p.seg_wblocks = seg_wblocks;
    p.segs = ceil_div(wblocks, seg_wblocks);
    p.iy_step = seg_wblocks * 8 * D / U;
    p.total_warps = planes * p.strips * p.segs;
    const long long blocks = (p.total_warps + FTC_WARPS - 1) / FTC_WARPS;
    if (blocks > 0x7fffffffLL) { set_error("filtered_lrelu_tc: too many strips"); return AFCM_ERR_INVALID; }
    const int smem = FTC_WARPS * FtcWarp<U, D, TIN, TOUT, ACT>::WARP_RING_BYTES;
    flr_tc_kernel<U, D, TIN, TOUT, ACT><<<(unsigned)blocks, FTC_WARPS * 32, smem, st>>>(p);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

template <typename TIN, typename TOUT, int ACT>
static int dispatch_geo(FlrTcParams& p, int N, int up, int down, cudaStream_t st)
{
    if (up == 2 && down == 2) return launch_tc<2, 2, TIN, TOUT, ACT>(p, N, st);
    if (up == 4 && down == 2) return launch_tc<4, 2, TIN, TOUT, ACT>(p, N, st);
    if (up == 2 && down == 4) return launch_tc<2, 4, TIN, TOUT, ACT>(p, N, st);
    return AFCM_ERR_UNSUPPORTED;
}

template <typename TIN, typename TOUT>
static int dispatch_act(FlrTcParams& p, int N, int up, int down, int act, cudaStream_t st)
{
    if (act == FTC_ACT_SAT) return dispatch_geo<TIN, TOUT, FTC_ACT_SAT>(p, N, up, down, st);
    return dispatch_geo<TIN, TOUT, FTC_ACT_MINMAX>(p, N, up, down, st);
}

}  // namespace afcm

using namespace afcm;

extern "C" int afcm_filtered_lrelu_tc(const void* x, const int64_t* xs, int x_dtype, void* y, const int64_t* ys, int y_dtype,
                                      const float* b, const void* skip,
                                      int N, int C, int xh, int xw, int yh, int yw,
                                      const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                                      int up, int down, int px0, int px1, int py0, int py1,
                                      float gain, float slope, float clamp, float out_scale, int flip_filter,
                                      void* stream)
{
    AFCM_CHECK_ARG(x && y && xs && ys, "x, y and their strides must be given");
    AFCM_CHECK_ARG((x_dtype == AFCM_F32 || x_dtype == AFCM_F16) && (y_dtype == AFCM_F32 || y_dtype == AFCM_F16),
                   "x and y must be float16 or float32");
    AFCM_CHECK_ARG(N > 0 && C > 0 && xh > 0 && xw > 0, "x is empty");
    AFCM_CHECK_ARG(fu_host && fd_host, "the tensor-core path needs both filters");
    int eyh = 0, eyw = 0;
    int rc = afcm_filtered_lrelu_out_size(xh, xw, up, down, fu_taps, fd_taps, px0, px1, py0, py1, &eyh, &eyw);
    if (rc) return rc;
    AFCM_CHECK_ARG(eyh == yh && eyw == yw, "y has shape [%d,%d], expected [%d,%d]", yh, yw, eyh, eyw);
    const bool geo_ok = (up == 2 && down == 2) || (up == 4 && down == 2) || (up == 2 && down == 4);
    if (!geo_ok || fu_taps != 6 * up || fd_taps != 6 * down || xs[3] != 1 || ys[3] != 1 || !(slope >= 0.f) || !(gain > 0.f)) {
        set_error("filtered_lrelu_tc: unsupported geometry up=%d/%d taps down=%d/%d taps (or non-unit inner stride)",
                  up, fu_taps, down, fd_taps);
        return AFCM_ERR_UNSUPPORTED;
    }
    // the input is read as aligned column pairs (one 8- or 4-byte load per pair)
    const uintptr_t pair_bytes = x_dtype == AFCM_F32 ? 8 : 4;
    if ((xw & 1) || (xs[0] & 1) || (xs[1] & 1) || (xs[2] & 1) || ((uintptr_t)x % pair_bytes) != 0) {
        set_error("filtered_lrelu_tc: x must have an even width, even strides and a pair-aligned base address");
        return AFCM_ERR_UNSUPPORTED;
    }
    AFCM_CHECK_ARG(xs[2] >= 0 && ys[2] >= 0 && (long long)(xh + 64) * xs[2] < (1ll << 31) && (long long)(yh + 64) * ys[2] < (1ll << 31),
                   "row strides out of range");
    FlrTcParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.y = y; p.b = b; p.skip = skip;
    p.xs_n = xs[0]; p.xs_c = xs[1]; p.xs_h = (int)xs[2];
    p.ys_n = ys[0]; p.ys_c = ys[1]; p.ys_h = (int)ys[2];
    p.C = C; p.xh = xh; p.xw = xw; p.yh = yh; p.yw = yw;
    // phase shift s: the strip's first up-sampled sample is a multiple of `up` away from the padding origin;
    // delta: one extra input column in front so that the column pairs are aligned.
    p.sx = floor_mod(-px0, up);
    p.sy = floor_mod(-py0, up);
    int bx = (-p.sx - px0) / up;            // exact
    p.dx = (bx & 1) ? 1 : 0;
    p.ix0 = bx - p.dx;
    p.iy0 = (-p.sy - py0) / up;
    p.slope = slope; p.out_scale = out_scale;
    // activation scale: R2 is computed in units of `clamp` when the sat() form applies
    const bool finite_clamp = clamp > 0.f && clamp < 3.0e38f;
    int act = FTC_ACT_MINMAX;
    float u_scale = 1.f;                    // R2 = x * gain * u_scale
    if (finite_clamp && clamp <= 1024.f && clamp >= 1.f / 1024.f) { act = FTC_ACT_SAT; u_scale = 1.f / clamp; }
    p.act_clamp = finite_clamp ? fminf(clamp, 65504.f) : 65504.f;
    for (int t = 0; t < fu_taps; t++) {
        const float f = fu_host[flip_filter ? t : fu_taps - 1 - t] * (float)up;
        p.kux[t] = f;
        p.kuy[t] = f * gain * u_scale;
    }
    for (int t = 0; t < fd_taps; t++) {
        const float f = fd_host[flip_filter ? t : fd_taps - 1 - t];
        p.kdx[t] = f / u_scale;
        p.kdy[t] = f * out_scale;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == AFCM_F32 && y_dtype == AFCM_F32) return dispatch_act<float, float>(p, N, up, down, act, st);
    if (x_dtype == AFCM_F32 && y_dtype == AFCM_F16) return dispatch_act<float, __half>(p, N, up, down, act, st);
    if (x_dtype == AFCM_F16 && y_dtype == AFCM_F32) return dispatch_act<__half, float>(p, N, up, down, act, st);
    return dispatch_act<__half, __half>(p, N, up, down, act, st);
}
