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
    int zero_cols;                                        // columns [yw, yw + zero_cols) of every output row are written as zeros (0 or 2)
    // sign tensor (training step): uint8 [N, C, sign_h, 4 sign_nw], 2 bits per up-sampled sample (OPS/filtered_lrelu.cpp:87-94);
    // written by the forward (SIGN == 1), read by the backward (SIGN == 2) at the offset (s_ox, s_oy)
    uint32_t* signs;
    int sign_h, sign_nw, s_ox, s_oy;
    const float* amax;                                    // SIGN == 2: device pointer to max|x| (operand scaling into the fp16 range) or null
    int strips, segs, seg_wblocks;                        // 16-column strips, row segments of 8*seg_wblocks rows
    int units;                                            // strips * segs: warps per plane
    unsigned total_warps;
    int ix0, iy0, iy_step;                                // input origin: col = strip*IXS + ix0, row = seg*iy_step + iy0
    int sx, sy, dx;
    unsigned zero;                                        // always 0 (keeps duplicated constant fragments in registers of their own)
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
    // a last strip with at most 8 output columns left (every AFCM plane but the 256-pixel one ends with 4): output column K
    // reads the up-sampled columns D K + sx .. D K + sx + FD - 1, sx < U, so columns 0..7 need this many 16-wide blocks
    static constexpr int MBN = (D * 7 + (U - 1) + FD + 15) / 16;
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

// ---- sign tensor helpers (lane-level model: tools/flr_sign_pack_model.py; hardware check: tools/microbench/sign_pack_check.cu)
// codes of the two packed samples of u (pre-activation in units of the clamp) and a = sat(u) - sat(-slope u):
// bits 0-1 = code of the low half, bits 16-17 = code of the high half; 1 = negative, 2 = clamped (overrides)
__device__ __forceinline__ uint32_t ftc_codes(uint32_t u, uint32_t a)
{
    uint32_t neg, cl;
    const uint32_t zero = 0u, one = 0x3c003c00u;
    asm("set.lt.u32.f16x2 %0, %1, %2;" : "=r"(neg) : "r"(u), "r"(zero));
    const uint32_t mag = a & 0x7fff7fffu;
    asm("set.ge.u32.f16x2 %0, %1, %2;" : "=r"(cl) : "r"(mag), "r"(one));
    return (neg & ~cl & 0x00010001u) | (cl & 0x00020002u);
}
// c[q][h]: codes of row block q (rows 8 q + 2 t + {0, 1}) and register h (column 8 h + g) of one block of 16 up-sampled columns.
// Returns the 32-bit word (16 codes, column i at bits 2 i) of row 8 (g >> 2) + 2 t + ((g >> 1) & 1); valid on the lanes with even g.
__device__ __forceinline__ uint32_t ftc_sign_block_word(const uint32_t (&c)[2][2], unsigned lane)
{
    const uint32_t sh = 2u * (lane >> 2);
    uint32_t P[2][2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
        P[q][0] = (__byte_perm(c[q][0], c[q][1], 0x4400) & 0x00030003u) << sh;
        P[q][1] = (__byte_perm(c[q][0], c[q][1], 0x6622) & 0x00030003u) << sh;
    }
    const bool g2 = (lane & 16u) != 0, g1 = (lane & 8u) != 0;
    uint32_t R[2];
#pragma unroll
    for (int pp = 0; pp < 2; pp++) {
        const uint32_t send = g2 ? P[0][pp] : P[1][pp], keep = g2 ? P[1][pp] : P[0][pp];
        R[pp] = keep | __shfl_xor_sync(0xffffffffu, send, 16);
    }
    const uint32_t send = g1 ? R[0] : R[1], keep = g1 ? R[1] : R[0];
    const uint32_t S = keep | __shfl_xor_sync(0xffffffffu, send, 8);
    return S | __shfl_xor_sync(0xffffffffu, S, 4);
}
// read direction: the word of row (q, p) lives on lane 4 (4 q + 2 p) + t; every lane fetches its four rows and extracts the codes
// of its two columns.  m[q][h] = packed half2 multipliers (1, slope, 0) for register h of row block q.
__device__ __forceinline__ void ftc_sign_block_mult(uint32_t word_of_my_row, unsigned lane, uint32_t h_one_slope, uint32_t (&m)[2][2])
{
    const unsigned t = lane & 3u, sh = 2u * (lane >> 2);
#pragma unroll
    for (int q = 0; q < 2; q++) {
        uint32_t w[2];
#pragma unroll
        for (int pp = 0; pp < 2; pp++) w[pp] = __shfl_sync(0xffffffffu, word_of_my_row, (int)(4u * (4u * q + 2u * pp) + t)) >> sh;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t c0 = (w[0] >> (16 * h)) & 3u, c1 = (w[1] >> (16 * h)) & 3u;
            const uint32_t lo = c0 >= 2u ? 0u : __byte_perm(h_one_slope, 0u, c0 ? 0x4432 : 0x4410);
            const uint32_t hi = c1 >= 2u ? 0u : __byte_perm(h_one_slope, 0u, c1 ? 0x4432 : 0x4410);
            m[q][h] = lo | (hi << 16);
        }
    }
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

// SIGN: 0 = inference (no sign tensor), 1 = forward of the training step (writes the sign tensor), 2 = backward (the op with
// up / down exchanged; the activation is replaced by the multiplier 1 / slope / 0 the stored code selects, OPS/filtered_lrelu.cu:562-572)
template <int U, int D, typename TIN, typename TOUT, int ACT, bool FAST, int SIGN = 0>
struct FtcWarp {
    // FAST: no bias and no skip tensor (the fast inference path: the convolution epilogue has already added the
    // bias).  Zero-filled halo samples then need no masking after the copy and the epilogue has no skip loads.
    using Geo = FtcGeo<U, D>;
    using RawT = typename Pair<TIN>::type;
    using OutPair = typename Pair<TOUT>::type;
    static constexpr int MB = Geo::KC4;            // 16-wide blocks of up-sampled columns
    static constexpr int JB = 2 * MB;              // 8-wide blocks of up-sampled columns
    static constexpr int NAL = D / 2;              // 16-row chunks of up-sampled rows per window of 8 output rows
    static constexpr int NREL = D;                 // 16-column chunks of up-sampled columns per 8 output columns
    // constant fragments (the FIR taps); b2r = b2 with the two K halves exchanged (kept as separate registers so
    // that either order is a ready-made operand pair)
    uint32_t a1[Geo::NPH][4], b2[U][2], b2r[U][2], a3[NAL][4], b4[NREL][2];
    uint32_t h_one, h_nslope, h_slope, h_cl, h_ncl, h_bias, h_one_slope;
    const FlrTcParams& p;
    int g, t;
    unsigned lane_id;
    uint32_t* sgn_plane;         // SIGN: sign words of this plane
    int cur_blk;                 // SIGN: index of the input row block converted last
    float in_scale, post_scale;  // SIGN == 2: operand scale (a power of two) applied at the input conversion, and its inverse applied to the fp32 result
    // strip state
    const TIN* xf;               // next row block to fetch: plane base + (its first row + g) rows + ix + 2t
    const TIN* xsafe;            // row 0 of the plane at this lane's column origin (source of zero-size copies)
    int fy;                      // row index of xf
    int ccol[Geo::NC];           // EDGE: element offset (relative to xf) of chunk c's column pair, clamped into the plane
    uint32_t csz[Geo::NC];       // EDGE: RAW_BYTES if the column pair is inside the plane, else 0
    uint32_t ring;               // shared-memory address of this lane's slot in stage 0, chunk 0
    uint32_t rd;                 // byte offset of the stage holding the next row block to convert
    int ix;                      // first input column of the strip
    TOUT* yq;                    // output pointer of the next block pair to emit: row (w0 + 8 b + g), column k0 + 2t
    int eb;                      // b of yq
    float bias;
    int iy, k0, w0, nwb;
    bool interior;               // every input column this strip reads and every output column it writes exists
    static constexpr int RAW_BYTES = (int)sizeof(RawT);
    static constexpr int CHUNK_BYTES = 32 * RAW_BYTES;               // one chunk of one row block, all lanes
    static constexpr int pow2ceil(int v) { int r = 1; while (r < v) r <<= 1; return r; }
    static constexpr int STAGE_BYTES = pow2ceil(Geo::NC * CHUNK_BYTES);          // power of two: slot = offset & mask
    static constexpr int WARP_RING_BYTES = FTC_STAGES * STAGE_BYTES;
    static constexpr uint32_t RING_MASK = (uint32_t)(WARP_RING_BYTES - 1);
    // pipeline state.  Strip-local coordinates: X / Y input column / row, J / V up-sampled column / row,
    // K / W output column / row;   up_x[J] = sum_X kux[(X - dx) U - J] in[X],   up_y[V] = sum_Y kuy[U Y - V] in[Y],
    // out_x[K] = sum_J kdx[J - D K - sx] a[J],   out_y[W] = sum_V kdy[V - D W - sy] a[V].
    uint32_t P[2][MB][2];        // R1 of the last two input row blocks (slot = block & 1): D-fragments [J 16][Y 8]
    uint32_t carry[JB];          // lower half (rows +8..+15) of the last window of R3

    __device__ FtcWarp(const FlrTcParams& p_, int lane) : p(p_), g(lane >> 2), t(lane & 3), lane_id((unsigned)lane), sgn_plane(nullptr), cur_blk(0), in_scale(1.f), post_scale(1.f) {}

    __device__ void load_consts(const float (*tab)[FTC_TAB])
    {
        const float* tux = tab[0] + FTC_TAB_OFS; const float* tuy = tab[1] + FTC_TAB_OFS;
        const float* tdx = tab[2] + FTC_TAB_OFS; const float* tdy = tab[3] + FTC_TAB_OFS;
#pragma unroll
        for (int ph = 0; ph < Geo::NPH; ph++) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int row = g + (r & 1) * 8, col = 2 * t + (r >> 1) * 8;
                // (1) A[m = J, k = X] = kux[(X - dx) U - J];  window phase ph: J = 16 ph + row (U == 4)
                const int ex = (col - p.dx) * U - 16 * ph - row;
                a1[ph][r] = pack_h2(tux[ex], tux[ex + U]);
            }
        }
#pragma unroll
        for (int nb = 0; nb < U; nb++) {
#pragma unroll
            for (int r = 0; r < 2; r++) {
                // (2) B[k = Y, n = V] = kuy[U Y - V],  Y = 2t + 8r + {0,1} in the 16-row window, V = 8 nb + g
                const int e = U * (2 * t + 8 * r) - 8 * nb - g;
                b2[nb][r] = pack_h2(tuy[e], tuy[e + U]);
            }
            // the exchanged copy must not be merged with the original by value numbering: p.zero is 0 at run time
            b2r[nb][0] = b2[nb][1] + p.zero;
            b2r[nb][1] = b2[nb][0] + p.zero;
        }
#pragma unroll
        for (int al = 0; al < NAL; al++) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int row = g + (r & 1) * 8, col = 2 * t + (r >> 1) * 8;
                // (3) A[m = W, k = V] = kdy[V - D W - sy] for chunk `al` of the window: V = 16 al + col (+ window
                // origin), W = row - 8 (+ the same origin / D)
                const int e = 16 * al + col - D * row + 8 * D - p.sy;
                a3[al][r] = pack_h2(tdy[e], tdy[e + 1]);
            }
        }
#pragma unroll
        for (int rel = 0; rel < NREL; rel++) {
#pragma unroll
            for (int r = 0; r < 2; r++) {
                // (4) B[k = J, n = K] = kdx[J - D K - sx],  J = 16 rel + 2t + 8r + {0,1} (+ 8 D nb), K = g (+ 8 nb)
                const int e = 16 * rel + 2 * t + 8 * r - D * g - p.sx;
                b4[rel][r] = pack_h2(tdx[e], tdx[e + 1]);
            }
        }
        h_one = pack_h2(1.f, 1.f);
        h_nslope = pack_h2(-p.slope, -p.slope);
        h_slope = pack_h2(p.slope, p.slope);
        h_cl = pack_h2(p.act_clamp, p.act_clamp);
        h_ncl = pack_h2(-p.act_clamp, -p.act_clamp);
        h_one_slope = pack_h2(1.f, p.slope);
    }

    __device__ void begin_strip(int unit, int plane)
    {
        const int seg = unit / p.strips, strip = unit - seg * p.strips;
        const int n = plane / p.C, c = plane - n * p.C;
        const TIN* xplane = (const TIN*)p.x + n * p.xs_n + c * p.xs_c;
        bias = (!FAST && p.b) ? p.b[c] : 0.f;
        h_bias = pack_h2(bias, bias);
        if (SIGN) sgn_plane = p.signs + (long long)plane * p.sign_h * p.sign_nw;
        ix = strip * Geo::IXS + p.ix0;
        iy = seg * p.iy_step + p.iy0;
        k0 = strip * 16;
        w0 = seg * p.seg_wblocks * 8;
        eb = 0;
        yq = (TOUT*)p.y + n * p.ys_n + c * p.ys_c + (long long)(w0 + g) * p.ys_h + k0 + 2 * t;
        fy = iy + g;
        xf = xplane + (long long)fy * p.xs_h + (ix + 2 * t);
        xsafe = xplane + (ix + 2 * t);          // + ccol[c] is a readable address in row 0 of the plane
        rd = 0;
        const int rows_left = p.yh - w0;
        nwb = (rows_left + 7) >> 3;
        if (nwb > p.seg_wblocks) nwb = p.seg_wblocks;
        // ix and xw are even (host-checked), so a column pair is either inside or outside the plane as a whole
        unsigned cm = 0;
#pragma unroll
        for (int c8 = 0; c8 < Geo::NC; c8++) {
            const int col = ix + 8 * c8 + 2 * t;
            const bool ok = col >= 0 && col < p.xw;
            if (ok) cm |= 1u << c8;
            csz[c8] = ok ? (uint32_t)RAW_BYTES : 0u;
            ccol[c8] = (ok ? col : 0) - (ix + 2 * t);
        }
        interior = __all_sync(0xffffffffu, cm == (1u << Geo::NC) - 1u) && k0 + 16 <= p.yw;       // (a strip holding the zero pair is not interior)
#pragma unroll
        for (int mb = 0; mb < MB; mb++) { P[0][mb][0] = P[0][mb][1] = P[1][mb][0] = P[1][mb][1] = 0u; }
#pragma unroll
        for (int jb = 0; jb < JB; jb++) carry[jb] = 0u;
    }

    // ---- input: issue the asynchronous copies of the next block of 8 input rows into the ring stage that was
    // converted last (prefetch distance 3 of 4 stages).  This thread: row g, NC column pairs.
    // EDGE: rows / columns outside the plane are zero-filled (source size 0, address clamped into the plane).
    template <bool EDGE>
    __device__ __forceinline__ void fetch()
    {
        const uint32_t dst = ring + ((rd + (uint32_t)((FTC_PD - 1) * STAGE_BYTES)) & RING_MASK);   // == stage of block yb + PD
        if (!EDGE) {
#pragma unroll
            for (int c8 = 0; c8 < Geo::NC; c8++) cp_async<RAW_BYTES>(dst + c8 * CHUNK_BYTES, xf + 8 * c8, RAW_BYTES);
        } else {
            const bool rok = (unsigned)fy < (unsigned)p.xh;
            const TIN* rp = rok ? xf : xsafe;
            const uint32_t rmask = rok ? 0xffffffffu : 0u;
#pragma unroll
            for (int c8 = 0; c8 < Geo::NC; c8++)
                cp_async<RAW_BYTES>(dst + c8 * CHUNK_BYTES, rp + ccol[c8], (int)(csz[c8] & rmask));
        }
        cp_async_commit();
        xf += 8 * p.xs_h;
        fy += 8;
    }
    // the first FTC_PD blocks go to stages 0 .. PD-1
    __device__ __forceinline__ void prime()
    {
#pragma unroll
        for (int i = 0; i < FTC_PD; i++) {
            const uint32_t dst = ring + (uint32_t)(i * STAGE_BYTES);
            const bool rok = (unsigned)fy < (unsigned)p.xh;
            const TIN* rp = rok ? xf : xsafe;
            const uint32_t rmask = rok ? 0xffffffffu : 0u;
#pragma unroll
            for (int c8 = 0; c8 < Geo::NC; c8++)
                cp_async<RAW_BYTES>(dst + c8 * CHUNK_BYTES, rp + ccol[c8], (int)(csz[c8] & rmask));
            cp_async_commit();
            xf += 8 * p.xs_h;
            fy += 8;
        }
    }
    __device__ __forceinline__ uint32_t cvt_pair(const float2& v) const
    {
        if (SIGN == 2) return pack_h2(v.x * in_scale, v.y * in_scale);        // backward: no bias, gradients scaled into the fp16 range
        return FAST ? pack_h2(v.x, v.y) : pack_h2(v.x + bias, v.y + bias);
    }
    __device__ __forceinline__ uint32_t cvt_pair(const uint32_t& v) const
    {
        if (FAST) return v;
        uint32_t r;
        asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(h_bias));
        return r;
    }
    // wait until the oldest row block in flight has landed (at most FTC_PD - 1 younger groups may still be in
    // flight), read this lane's pairs back, add the bias and round to fp16.  EDGE (only with a bias): pairs outside
    // the plane stay zero.  `cy` = first row of the block (for the row test).
    template <bool EDGE>
    __device__ __forceinline__ void convert(int yb, uint32_t (&in)[Geo::NC])
    {
        cp_async_wait<FTC_PD - 1>();
        if (SIGN) cur_blk = yb;
        const uint32_t src = ring + rd;
        rd = (rd + (uint32_t)STAGE_BYTES) & RING_MASK;
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
        if (!EDGE || FAST) {
#pragma unroll
            for (int c8 = 0; c8 < Geo::NC; c8++) in[c8] = cvt_pair(raw[c8]);
        } else {
            const int row = iy + 8 * yb + g;
            const bool rok = row >= 0 && row < p.xh;
#pragma unroll
            for (int c8 = 0; c8 < Geo::NC; c8++) in[c8] = (rok && csz[c8]) ? cvt_pair(raw[c8]) : 0u;
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

    // (1) horizontal up-FIR of input row block yb (already converted) -> P[SLOT]
    // NARROW: the strip has at most 8 output columns, only the first MBN blocks of up-sampled columns and the first block of
    // output columns are computed
    template <int SLOT, bool NARROW>
    __device__ __forceinline__ void step1(const uint32_t (&in)[Geo::NC])
    {
#pragma unroll
        for (int b = 0; b < (NARROW ? Geo::MBN : MB); b++) {
            const int w = (U == 2) ? b : (b >> 1);          // first input chunk of the window
            const int ph = (U == 2) ? 0 : (b & 1);
            mma_h(P[SLOT][b], a1[ph], in[w], in[w + 1]);
        }
    }

    // (2) vertical up-FIR of the 16-row window (previous block | block in slot CUR) for the up-sampled row blocks
    // NB0, NB0 + 1 (= one 16-row chunk), activation, and (3) that chunk's contribution to the window of R3
    // ({ rows -8..-1, rows 0..7 } relative to the window's block of 8 output rows).
    // MODE 0 (D == 2, the chunk is the whole window): X[jb] = carry + upper half, carry = lower half.
    // MODE 1 (D == 4, first chunk): win = contribution.   MODE 2 (second chunk): win += contribution, then as MODE 0.
    template <int CUR, int NB0, int MODE, bool NARROW>
    __device__ __forceinline__ void chunk(int al, uint32_t (&win)[JB][2], uint32_t (&X)[JB])
    {
        // SIGN: the fragment element (J = 16 mb + 8 h + g, V = 8 (NB0 + q) + 2 t + p) of this chunk is the up-sampled sample
        //   row uy = D w0 + 8 U (cur_blk - 1) + 8 NB0 + V' - sy,   column ux = D k0 + J - sx     of the sign tensor
        // (tools/flr_tc_emu.py, tests/test_flr_tc_emu.py); after the packing butterfly the lanes with even g hold one row each
        const int sgn_row = SIGN ? D * w0 + 8 * U * (cur_blk - 1) + 8 * NB0 + 8 * (g >> 2) + 2 * t + ((g >> 1) & 1) - p.sy : 0;
        // SIGN == 1: word w of this lane's row = column blocks w and w + 1 funnel-shifted by 2 sx bits: a sliding pair of packed block
        // words is enough (no array of MB + 1 words: the up 2 / down 4 kernel has no registers to spare).  Ownership: a strip stores
        // the D aligned words of its own 16 D up-sampled columns (the last strip every word up to the end of the row), a segment
        // the rows of its own output rows (the last one the rest)
        uint32_t Tprev = 0u;
        uint32_t* sgn_wp = nullptr;              // where word 0 of this lane's row goes, or null (row not owned / odd g)
        int n_own = 0;
        if (SIGN == 1) {
            const int strip = k0 >> 4;
            n_own = (strip == p.strips - 1) ? p.sign_nw - D * strip : D;
            const int row_lo = max(D * w0, 0);
            const int row_hi = (w0 + 8 * p.seg_wblocks >= p.yh) ? p.sign_h : D * (w0 + 8 * p.seg_wblocks);
            if (!(g & 1) && sgn_row >= row_lo && sgn_row < row_hi) sgn_wp = sgn_plane + (long long)sgn_row * p.sign_nw + D * strip;
        }
        uint32_t sw[MB + 2];                     // SIGN == 2: the stored words this lane's row needs (column blocks + shift spill)
        int sshift = 0;
        if (SIGN == 2) {
            // columns ex = D k0 + 16 mb + j - sx + s_ox: word (ex0 >> 4) + mb, shifted by 2 (ex0 & 15) bits; outside the tensor = code 0
            const int ex0 = D * k0 - p.sx + p.s_ox;
            const int wi0 = ex0 >> 4;
            sshift = 2 * (ex0 & 15);
            const int ey = sgn_row + p.s_oy;
            const bool rok = !(g & 1) && ey >= 0 && ey < p.sign_h;
            const uint32_t* rowp = sgn_plane + (long long)(rok ? ey : 0) * p.sign_nw;
#pragma unroll
            for (int i = 0; i < MB + 2; i++) {
                const int wi = wi0 + i;
                sw[i] = (rok && wi >= 0 && wi < p.sign_nw) ? __ldg(rowp + wi) : 0u;
            }
        }
#pragma unroll
        for (int mb = 0; mb < (NARROW ? Geo::MBN : MB); mb++) {
            const uint32_t quad[4] = {P[0][mb][0], P[0][mb][1], P[1][mb][0], P[1][mb][1]};
            uint32_t e[2][2];
            uint32_t mlt[2][2];
            if (SIGN == 2) ftc_sign_block_mult(__funnelshift_r(sw[mb], sw[mb + 1], sshift), lane_id, h_one_slope, mlt);
            uint32_t cod[2][2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                uint32_t d[2];
                // slot 0 supplies k 0..7: if it holds the current block the two K halves of the constant swap
                if (CUR == 1) mma_h(d, quad, b2[NB0 + q][0], b2[NB0 + q][1]);
                else mma_h(d, quad, b2r[NB0 + q][0], b2r[NB0 + q][1]);
                if (SIGN == 2) {
                    e[q][0] = h2_mul(d[0], mlt[q][0]);
                    e[q][1] = h2_mul(d[1], mlt[q][1]);
                } else {
                    e[q][0] = act2(d[0]);
                    e[q][1] = act2(d[1]);
                    if (SIGN == 1) { cod[q][0] = ftc_codes(d[0], e[q][0]); cod[q][1] = ftc_codes(d[1], e[q][1]); }
                }
            }
            if (SIGN == 1) {
                const uint32_t Tcur = ftc_sign_block_word(cod, lane_id);
                if (mb >= 1 && sgn_wp && mb - 1 < n_own) sgn_wp[mb - 1] = __funnelshift_r(Tprev, Tcur, 2 * p.sx);
                Tprev = Tcur;
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int jb = 2 * mb + h;
                if (MODE == 0) {
                    uint32_t w[2];
                    mma_h(w, a3[al], e[0][h], e[1][h]);
                    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(X[jb]) : "r"(carry[jb]), "r"(w[0]));
                    carry[jb] = w[1];
                } else if (MODE == 1) {
                    mma_h(win[jb], a3[al], e[0][h], e[1][h]);
                } else {
                    mma_h_acc(win[jb], a3[al], e[0][h], e[1][h]);
                    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(X[jb]) : "r"(carry[jb]), "r"(win[jb][0]));
                    carry[jb] = win[jb][1];
                }
            }
        }
        if (SIGN == 1 && sgn_wp) {
            // the last computed block pairs with an all-zero block; words beyond the computed blocks (narrow strips) are zero
            constexpr int NB = NARROW ? Geo::MBN : MB;
            if (NB - 1 < n_own) sgn_wp[NB - 1] = __funnelshift_r(Tprev, 0u, 2 * p.sx);
#pragma unroll
            for (int w = NB; w < MB; w++)
                if (w < n_own) sgn_wp[w] = 0u;
        }
    }

    // (4) horizontal down-FIR of two blocks of 8 output rows (X0: rows 8b.., X1: rows 8b+8..), fp32 accumulate, store.
    // out_scale is folded into the Td_x taps; a skip tensor is added as skip * out_scale.
    // `rows`: 3 = both blocks, 2 = only the second (X1), 1 = only the first.  The block pair is the one yq points
    // at; ADV = number of 8-row blocks yq advances afterwards.
    template <bool EDGE, int ADV, bool NARROW>
    __device__ __forceinline__ void emit(const uint32_t (&X0)[JB], const uint32_t (&X1)[JB], int rows)
    {
        TOUT* q0 = yq;                                              // (row w0 + 8 eb + g, col k0 + 2t)
        TOUT* q1 = yq + 8 * p.ys_h;
        bool r0 = (rows & 1) != 0, r1 = (rows & 2) != 0;
        if (EDGE) {
            r0 = r0 && eb >= 0 && eb < nwb && w0 + 8 * eb + g < p.yh;
            r1 = r1 && eb + 1 >= 0 && eb + 1 < nwb && w0 + 8 * eb + 8 + g < p.yh;
        }
#pragma unroll
        for (int nb = 0; nb < (NARROW ? 1 : 2); nb++) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int rel = 0; rel < NREL; rel++) {
                const int kc = nb * (D / 2) + rel;
                const uint32_t quad[4] = {X0[2 * kc], X1[2 * kc], X0[2 * kc + 1], X1[2 * kc + 1]};
                mma_f(c, quad, b4[rel][0], b4[rel][1]);
            }
            if (SIGN == 2) { c[0] *= post_scale; c[1] *= post_scale; c[2] *= post_scale; c[3] *= post_scale; }
            const bool has_skip = !FAST && p.skip != nullptr;
            const long long kofs = (const TOUT*)p.skip - (const TOUT*)p.y;      // used only when has_skip
            if (!EDGE) {
                if (has_skip) {
                    c[0] += (float)q0[8 * nb + kofs] * p.out_scale; c[1] += (float)q0[8 * nb + kofs + 1] * p.out_scale;
                    c[2] += (float)q1[8 * nb + kofs] * p.out_scale; c[3] += (float)q1[8 * nb + kofs + 1] * p.out_scale;
                }
                if (r0) store_pair(q0 + 8 * nb, c[0], c[1]);
                if (r1) store_pair(q1 + 8 * nb, c[2], c[3]);
            } else {
                // yw is even and the strip origin is even (host-checked): a column pair is inside or outside as a whole
                const int col = k0 + 8 * nb + 2 * t;
                const bool cok = col < p.yw;
                if (r0 && cok) {
                    if (has_skip) { c[0] += (float)q0[8 * nb + kofs] * p.out_scale; c[1] += (float)q0[8 * nb + kofs + 1] * p.out_scale; }
                    store_pair(q0 + 8 * nb, c[0], c[1]);
                }
                if (r1 && cok) {
                    if (has_skip) { c[2] += (float)q1[8 * nb + kofs] * p.out_scale; c[3] += (float)q1[8 * nb + kofs + 1] * p.out_scale; }
                    store_pair(q1 + 8 * nb, c[2], c[3]);
                }
                // planes stored at the row pitch yw + 2 for the convolution that follows (its flat-plane formulation wants two
                // zero pixels behind every row, conv2d_tc.cu): the pair behind the last column is written as zeros
                if (!cok && col < p.yw + p.zero_cols) {
                    if (r0) store_pair(q0 + 8 * nb, 0.f, 0.f);
                    if (r1) store_pair(q1 + 8 * nb, 0.f, 0.f);
                }
            }
        }
        if (ADV) { yq += ADV * 8 * p.ys_h; eb += ADV; }
    }
    // an emit whose blocks all lie outside the segment: only the output cursor moves
    template <int ADV> __device__ __forceinline__ void skip_emit() { yq += ADV * 8 * p.ys_h; eb += ADV; }
    static __device__ __forceinline__ void store_pair(float* q, float a, float b) { *reinterpret_cast<float2*>(q) = make_float2(a, b); }
    static __device__ __forceinline__ void store_pair(__half* q, float a, float b) { *reinterpret_cast<__half2*>(q) = __floats2half2_rn(a, b); }

    static __device__ __forceinline__ int fdiv8(int a) { return a >> 3; }                 // floor(a / 8)

    // ---- U == 2, D == 2: super-iteration s = input row blocks 2s, 2s+1 -> up-sampled chunks 2s-1, 2s (= windows)
    // -> output row blocks 2s-2, 2s-1, emitted together.
    template <bool EDGE, bool NARROW>
    __device__ __forceinline__ void super22(int s, uint32_t (&X0)[JB], uint32_t (&X1)[JB])
    {
        uint32_t in[Geo::NC], win[JB][2];
        // Pipeline fill / drain (EDGE supers only; the interior supers emit both blocks by construction):
        //  * s == 0 produces blocks -2 and -1, which do not exist.  The first chunk only reaches block -1 (through the
        //    carry), so it is skipped together with the emit; the second chunk stays (its carry is part of block 0).
        //  * a last super whose second block lies beyond the segment (odd number of row blocks) stops after the first.
        convert<EDGE>(2 * s, in); step1<0, NARROW>(in); fetch<EDGE>();
        if (!EDGE || s > 0) chunk<0, 0, 0, NARROW>(0, win, X0);                 // block 2s-2
        if (!EDGE || eb + 1 < nwb) {
            convert<EDGE>(2 * s + 1, in); step1<1, NARROW>(in); fetch<EDGE>();
            chunk<1, 0, 0, NARROW>(0, win, X1);                                 // block 2s-1
        }
        if (!EDGE || s > 0) emit<EDGE, 2, NARROW>(X0, X1, 3);                   // EDGE: rows outside the segment are masked
        else skip_emit<2>();
    }
    template <bool NARROW>
    __device__ void run22()
    {
        uint32_t X0[JB], X1[JB];
#pragma unroll
        for (int jb = 0; jb < JB; jb++) X0[jb] = X1[jb] = 0u;
        const int S = ((nwb - 1) >> 1) + 2;
        int lo = 0, hi = 0;
        if (!NARROW && interior) {
            lo = max(1, (-iy + 15) >> 4);
            hi = min(min((fdiv8(p.xh - 8 - iy) - FTC_PD - 1) >> 1, min((p.yh - w0) >> 4, nwb >> 1)) + 1, S);
        }
        if (hi < lo) hi = lo = 0;
        eb = -2; yq -= 16 * p.ys_h;
        int s = 0;
        for (; s < (hi > lo ? lo : S); s++) super22<true, NARROW>(s, X0, X1);
        if (!NARROW && hi > lo) {
            for (; s < hi; s++) super22<false, false>(s, X0, X1);
            for (; s < S; s++) super22<true, false>(s, X0, X1);
        }
    }

    // ---- U == 4, D == 2: iteration it = input row block it -> chunks (= windows) 2it-2, 2it-1 -> output row blocks
    // 2it-3 (odd, completes the pair started by the previous iteration) and 2it-2 (even, kept in X0).
    template <bool EDGE, int CUR, bool NARROW>
    __device__ __forceinline__ void iter42(int it, uint32_t (&X0)[JB])
    {
        uint32_t in[Geo::NC], win[JB][2], X1[JB];
        // eb == 2it-4.  Pipeline fill / drain (EDGE iterations only): a chunk is needed if its block or the block its
        // carry reaches exists, the emit if one block of its pair exists; an iteration past the segment does nothing.
        if (EDGE && eb >= nwb) { skip_emit<2>(); return; }
        convert<EDGE>(it, in); step1<CUR, NARROW>(in); fetch<EDGE>();
        if (EDGE) {
#pragma unroll
            for (int jb = 0; jb < JB; jb++) X1[jb] = 0u;
        }
        if (!EDGE || (it >= 1 && eb + 1 < nwb)) chunk<CUR, 0, 0, NARROW>(0, win, X1);   // block 2it-3 (carry: 2it-2)
        if (!EDGE || it >= 2) emit<EDGE, 2, NARROW>(X0, X1, 3);                         // pair (2it-4, 2it-3)
        else skip_emit<2>();
        if (!EDGE || (it >= 1 && eb < nwb)) chunk<CUR, 2, 0, NARROW>(0, win, X0);       // block 2it-2 (== eb after the emit)
    }
    template <bool NARROW>
    __device__ void run42()
    {
        uint32_t X0[JB];
#pragma unroll
        for (int jb = 0; jb < JB; jb++) X0[jb] = 0u;
        const int iters = ((nwb - 1) >> 1) + 3;
        const int S = (iters + 1) >> 1;
        int lo = 0, hi = 0;
        if (!NARROW && interior) {
            // iterations 2s, 2s+1: input blocks up to 2s+1 (+PD), output pairs s-2.. : blocks 4s-4 .. 4s-1
            lo = max(1, (-iy + 15) >> 4);
            hi = min(min((fdiv8(p.xh - 8 - iy) - FTC_PD - 1) >> 1, min((p.yh - w0) >> 5, nwb >> 2)) + 1, S);
        }
        if (hi < lo) hi = lo = 0;
        eb = -4; yq -= 32 * p.ys_h;
        int s = 0;
        for (; s < (hi > lo ? lo : S); s++) { iter42<true, 0, NARROW>(2 * s, X0); iter42<true, 1, NARROW>(2 * s + 1, X0); }
        if (!NARROW && hi > lo) {
            for (; s < hi; s++) { iter42<false, 0, false>(2 * s, X0); iter42<false, 1, false>(2 * s + 1, X0); }
            for (; s < S; s++) { iter42<true, 0, false>(2 * s, X0); iter42<true, 1, false>(2 * s + 1, X0); }
        }
    }

    // ---- U == 2, D == 4: iteration it = input row blocks 2it, 2it+1 -> chunks 2it-1 (second half of window it-1)
    // and 2it (first half of window it) -> output row block it-2; blocks are emitted one at a time as the second
    // half of the pair (it-3, it-2).
    template <bool EDGE, bool NARROW>
    __device__ __forceinline__ void iter24(int it, uint32_t (&win)[JB][2], uint32_t (&Xp)[JB])
    {
        uint32_t in[Geo::NC], X[JB];
        // eb == it-3.  Pipeline fill / drain (EDGE iterations only): iteration 0 has no block of its own (its first chunk
        // only reaches block -1), iteration 1 emits block -1; the last iteration's second chunk opens a window past the
        // segment.
        convert<EDGE>(2 * it, in); step1<0, NARROW>(in); fetch<EDGE>();
        if (EDGE) {
#pragma unroll
            for (int jb = 0; jb < JB; jb++) X[jb] = 0u;
        }
        if (!EDGE || it >= 1) chunk<0, 0, 2, NARROW>(1, win, X);                // block it-2 (carry: it-1)
        if (!EDGE || it >= 2) emit<EDGE, 1, NARROW>(Xp, X, 2);                  // pair (it-3, it-2), second half only
        else skip_emit<1>();
#pragma unroll
        for (int jb = 0; jb < JB; jb++) Xp[jb] = X[jb];
        if (!EDGE || it - 1 < nwb) {
            convert<EDGE>(2 * it + 1, in); step1<1, NARROW>(in); fetch<EDGE>();
            chunk<1, 0, 1, NARROW>(0, win, X);
        }
    }
    template <bool NARROW>
    __device__ void run24()
    {
        uint32_t win[JB][2], Xp[JB];
#pragma unroll
        for (int jb = 0; jb < JB; jb++) { win[jb][0] = win[jb][1] = 0u; Xp[jb] = 0u; }
        const int iters = nwb + 2;
        int lo = 0, hi = 0;
        if (!NARROW && interior) {
            lo = max(3, (-iy + 15) >> 4);
            hi = min(min((fdiv8(p.xh - 8 - iy) - FTC_PD - 1) >> 1, fdiv8(p.yh - w0) + 1) + 1, iters);
        }
        if (hi < lo) hi = lo = 0;
        eb = -3; yq -= 24 * p.ys_h;
        int it = 0;
        for (; it < (hi > lo ? lo : iters); it++) iter24<true, NARROW>(it, win, Xp);
        if (!NARROW && hi > lo) {
            for (; it < hi; it++) iter24<false, false>(it, win, Xp);
            for (; it < iters; it++) iter24<true, false>(it, win, Xp);
        }
    }

    __device__ void run()
    {
        prime();
        const bool narrow = p.yw + p.zero_cols - k0 <= 8;             // warp-uniform (the zero pair behind the last column counts)
        if (U == 2 && D == 2) { if (narrow) run22<true>(); else run22<false>(); }
        else if (U == 4 && D == 2) { if (narrow) run42<true>(); else run42<false>(); }
        else { if (narrow) run24<true>(); else run24<false>(); }
        cp_async_wait<0>();
    }
};

#ifndef AFCM_FTC_MINB22
#define AFCM_FTC_MINB22 4
#endif
#ifndef AFCM_FTC_MINB24
#define AFCM_FTC_MINB24 2
#endif
#ifndef AFCM_FTC_MINB_SIGN          // CTAs per SM of the sign-tensor variants (up 2 / down 2 and up 4 / down 2)
#define AFCM_FTC_MINB_SIGN 4
#endif
template <int U, int D, typename TIN, typename TOUT, int ACT, bool FAST, int SIGN>
__global__ void __launch_bounds__(FTC_WARPS * 32, (U == 2 && D == 4) ? AFCM_FTC_MINB24 : (SIGN ? AFCM_FTC_MINB_SIGN : (U == 4 ? 4 : AFCM_FTC_MINB22)))
flr_tc_kernel(const __grid_constant__ FlrTcParams p)
{
    // zero-padded tap tables (index e + FTC_TAB_OFS): the fragments below index them with per-lane offsets
    __shared__ float tab[4][FTC_TAB];
    // SIGN == 2 (backward): gradients may lie far below the fp16 range; the operands are scaled by a power of two that brings
    // max|x| (a device-side scalar written by afcm_absmax) to ~256, the fp32 result is scaled back in emit()
    float in_scale = 1.f;
    if (SIGN == 2 && p.amax) {
        const float a = *p.amax;
        if (a > 0.f && a < 3.0e38f) in_scale = exp2f(fminf(fmaxf(floorf(log2f(256.f / a)), -60.f), 60.f));
    }
    for (int i = threadIdx.x; i < FTC_TAB; i += blockDim.x) {
        const int e = i - FTC_TAB_OFS;
        const bool u_ok = e >= 0 && e < 6 * U, d_ok = e >= 0 && e < 6 * D;
        tab[0][i] = u_ok ? p.kux[e] : 0.f;
        tab[1][i] = u_ok ? p.kuy[e] : 0.f;
        tab[2][i] = d_ok ? p.kdx[e] : 0.f;
        tab[3][i] = d_ok ? p.kdy[e] : 0.f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    extern __shared__ __align__(16) uint8_t ring_smem[];
    typedef FtcWarp<U, D, TIN, TOUT, ACT, FAST, SIGN> W;
    W w(p, lane);
    w.in_scale = in_scale;               // applied to the DATA at the input conversion (the taps are fp16 operands themselves)
    w.post_scale = 1.f / in_scale;
    w.ring = (uint32_t)__cvta_generic_to_shared(ring_smem) + (uint32_t)(warp * W::WARP_RING_BYTES + lane * W::RAW_BYTES);
    w.load_consts(tab);
    // Persistent warps: the grid is one resident wave, every warp walks over (plane, segment, strip) units with the tap
    // fragments built once (the table fill, the barrier and load_consts are ~250 of the ~500 instructions a warp used to
    // spend before its first product -- a third of all instructions on the 36-pixel planes).  Consecutive units are
    // neighbouring strips of one plane, so the warps of a CTA still share their halo columns in L1/L2.
    for (unsigned wid = blockIdx.x * FTC_WARPS + warp; wid < p.total_warps; wid += gridDim.x * FTC_WARPS) {
        const int plane = (int)(wid / (unsigned)p.units);            // n * C + c
        const int unit = (int)(wid - (unsigned)plane * (unsigned)p.units);
        w.begin_strip(unit, plane);
        w.run();
    }
}

static int floor_mod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

static int g_ftc_waves = 16;          // 0: one warp per unit (no persistence); n: at most n resident waves of CTAs (8..16 measured best)

template <int U, int D, typename TIN, typename TOUT, int ACT, bool FAST, int SIGN = 0>
static int launch_tc(FlrTcParams& p, int N, cudaStream_t st)
{
    p.strips = ceil_div(p.yw + p.zero_cols, 16);
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
    p.units = p.strips * p.segs;
    if (planes * p.units > 0x7fffffffLL) { set_error("filtered_lrelu_tc: too many strips"); return AFCM_ERR_INVALID; }
    p.total_warps = (unsigned)(planes * p.units);
    unsigned blocks = (p.total_warps + FTC_WARPS - 1) / FTC_WARPS;
    const int smem = FTC_WARPS * FtcWarp<U, D, TIN, TOUT, ACT, FAST, SIGN>::WARP_RING_BYTES;
    // one resident wave (persistent warps); g_ftc_waves > 1 launches that many waves' worth of CTAs (tuning switch)
    static int resident = 0;                  // per instantiation: CTAs per SM x SMs
    if (!resident) {
        int per_sm = 0, sms = 0, dev = 0;
        AFCM_CUDA(cudaGetDevice(&dev));
        AFCM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        AFCM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flr_tc_kernel<U, D, TIN, TOUT, ACT, FAST, SIGN>, FTC_WARPS * 32, smem));
        resident = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
    }
    if (g_ftc_waves > 0 && blocks > (unsigned)(resident * g_ftc_waves)) blocks = (unsigned)(resident * g_ftc_waves);
    flr_tc_kernel<U, D, TIN, TOUT, ACT, FAST, SIGN><<<blocks, FTC_WARPS * 32, smem, st>>>(p);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

template <typename TIN, typename TOUT, int ACT, bool FAST>
static int dispatch_geo(FlrTcParams& p, int N, int up, int down, cudaStream_t st)
{
    if (up == 2 && down == 2) return launch_tc<2, 2, TIN, TOUT, ACT, FAST>(p, N, st);
    if (up == 4 && down == 2) return launch_tc<4, 2, TIN, TOUT, ACT, FAST>(p, N, st);
    if (up == 2 && down == 4) return launch_tc<2, 4, TIN, TOUT, ACT, FAST>(p, N, st);
    return AFCM_ERR_UNSUPPORTED;
}

template <typename TIN, typename TOUT>
static int dispatch_act(FlrTcParams& p, int N, int up, int down, int act, cudaStream_t st)
{
    if (act == FTC_ACT_SAT) return dispatch_geo<TIN, TOUT, FTC_ACT_SAT, false>(p, N, up, down, st);
    return dispatch_geo<TIN, TOUT, FTC_ACT_MINMAX, false>(p, N, up, down, st);
}

// max|x| of a dense fp32 tensor into out[0] (a device scalar, zeroed by the launcher): |x| >= 0, so the float bit patterns order
// like unsigned integers and one atomicMax per block is enough; NaN / Inf inputs are ignored (the scale must stay finite)
__global__ void __launch_bounds__(256)
absmax_kernel(const float4* __restrict__ x4, const float* __restrict__ x, long long n4, long long n, unsigned* __restrict__ out)
{
    float m = 0.f;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
        const float4 v = x4[i];
        m = fmaxf(fmaxf(m, fminf(fabsf(v.x), 3.0e38f)), fmaxf(fminf(fabsf(v.y), 3.0e38f), fmaxf(fminf(fabsf(v.z), 3.0e38f), fminf(fabsf(v.w), 3.0e38f))));
    }
    if (blockIdx.x == 0)
        for (long long i = 4 * n4 + threadIdx.x; i < n; i += 256) m = fmaxf(m, fminf(fabsf(x[i]), 3.0e38f));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; w++) m = fmaxf(m, sm[w]);
        atomicMax(out, __float_as_uint(m));
    }
}

}  // namespace afcm

using namespace afcm;

extern "C" int afcm_absmax(const float* x, int64_t n, float* out, void* stream)
{
    AFCM_CHECK_ARG(x && out && n > 0, "empty problem");
    AFCM_CHECK_ARG(((uintptr_t)x & 15) == 0, "x must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    AFCM_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
    const long long n4 = n / 4;
    long long blocks = (n4 + 256 * 8 - 1) / (256 * 8);
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(x), x, n4, (long long)n, reinterpret_cast<unsigned*>(out));
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

extern "C" int afcm_filtered_lrelu_tc_set_waves(int waves) { g_ftc_waves = waves < 0 ? 0 : waves; return AFCM_OK; }

extern "C" int afcm_filtered_lrelu_tc(const void* x, const int64_t* xs, int x_dtype, void* y, const int64_t* ys, int y_dtype,
                                      const float* b, const void* skip,
                                      int N, int C, int xh, int xw, int yh, int yw,
                                      const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                                      int up, int down, int px0, int px1, int py0, int py1,
                                      float gain, float slope, float clamp, float out_scale, int flip_filter,
                                      void* stream)
{
    return afcm_filtered_lrelu_tc_padded(x, xs, x_dtype, y, ys, y_dtype, b, skip, N, C, xh, xw, yh, yw, fu_host, fu_taps, fd_host, fd_taps,
                                         up, down, px0, px1, py0, py1, gain, slope, clamp, out_scale, flip_filter, 0, stream);
}

static int flr_tc_run(const void* x, const int64_t* xs, int x_dtype, void* y, const int64_t* ys, int y_dtype,
                      const float* b, const void* skip,
                      int N, int C, int xh, int xw, int yh, int yw,
                      const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                      int up, int down, int px0, int px1, int py0, int py1,
                      float gain, float slope, float clamp, float out_scale, int flip_filter,
                      int zero_pad_cols, int sign_mode, void* signs, int sign_h, int sign_wb, int s_ox, int s_oy, const float* amax,
                      void* stream);

extern "C" int afcm_filtered_lrelu_tc_padded(const void* x, const int64_t* xs, int x_dtype, void* y, const int64_t* ys, int y_dtype,
                                             const float* b, const void* skip,
                                             int N, int C, int xh, int xw, int yh, int yw,
                                             const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                                             int up, int down, int px0, int px1, int py0, int py1,
                                             float gain, float slope, float clamp, float out_scale, int flip_filter,
                                             int zero_pad_cols, void* stream)
{
    return flr_tc_run(x, xs, x_dtype, y, ys, y_dtype, b, skip, N, C, xh, xw, yh, yw, fu_host, fu_taps, fd_host, fd_taps, up, down,
                      px0, px1, py0, py1, gain, slope, clamp, out_scale, flip_filter, zero_pad_cols, AFCM_SIGN_NONE, nullptr, 0, 0, 0, 0,
                      nullptr, stream);
}

// The training-step variant of the register-chained kernel: fp32 planes in and out, sign tensor written (forward) or read (backward,
// the op with up / down exchanged) in the reference's format, interchangeable with afcm_filtered_lrelu / afcm_filtered_lrelu_tcs.
// sign_mode AFCM_SIGN_WRITE: signs [N,C,sign_h,sign_wb] uint8 as sized by afcm_filtered_lrelu_sign_size, (sx, sy) = 0, a finite
// clamp in [2^-10, 2^10].  AFCM_SIGN_READ: (sx, sy) = sign offsets (OPS/filtered_lrelu.py:258-259); amax = device pointer to max|x|
// (afcm_absmax) or NULL: the fp16 operands are scaled by the power of two that brings it to ~256, the fp32 result is scaled back.
extern "C" int afcm_filtered_lrelu_tc_signs(const void* x, const int64_t* xs, void* y, const int64_t* ys, const float* b,
                                            int N, int C, int xh, int xw, int yh, int yw,
                                            const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                                            int up, int down, int px0, int px1, int py0, int py1,
                                            float gain, float slope, float clamp, int flip_filter,
                                            int sign_mode, void* signs, int sign_h, int sign_wb, int sx, int sy, const float* amax,
                                            void* stream)
{
    AFCM_CHECK_ARG(sign_mode == AFCM_SIGN_WRITE || sign_mode == AFCM_SIGN_READ, "sign_mode must be AFCM_SIGN_WRITE or AFCM_SIGN_READ");
    AFCM_CHECK_ARG(signs && sign_h > 0 && sign_wb > 0 && (sign_wb & 3) == 0 && ((uintptr_t)signs & 3) == 0,
                   "the sign tensor must be given, 4-byte aligned, with a row length that is a multiple of 4 bytes");
    if (sign_mode == AFCM_SIGN_WRITE) {
        int sh = 0, swb = 0;
        afcm_filtered_lrelu_sign_size(yh, yw, down, fd_taps, &sh, &swb);
        AFCM_CHECK_ARG(sign_h == sh && sign_wb == swb, "sign tensor has shape [%d,%d], expected [%d,%d]", sign_h, sign_wb, sh, swb);
        AFCM_CHECK_ARG(sx == 0 && sy == 0, "sign offsets must be zero when writing signs");
    }
    return flr_tc_run(x, xs, AFCM_F32, y, ys, AFCM_F32, b, nullptr, N, C, xh, xw, yh, yw, fu_host, fu_taps, fd_host, fd_taps, up, down,
                      px0, px1, py0, py1, gain, slope, clamp, 1.f, flip_filter, 0, sign_mode, signs, sign_h, sign_wb, sx, sy, amax, stream);
}

static int flr_tc_run(const void* x, const int64_t* xs, int x_dtype, void* y, const int64_t* ys, int y_dtype,
                      const float* b, const void* skip,
                      int N, int C, int xh, int xw, int yh, int yw,
                      const float* fu_host, int fu_taps, const float* fd_host, int fd_taps,
                      int up, int down, int px0, int px1, int py0, int py1,
                      float gain, float slope, float clamp, float out_scale, int flip_filter,
                      int zero_pad_cols, int sign_mode, void* signs, int sign_h, int sign_wb, int s_ox, int s_oy, const float* amax,
                      void* stream)
{
    AFCM_CHECK_ARG(zero_pad_cols == 0 || (zero_pad_cols == 2 && ys && ys[2] >= yw + 2), "zero_pad_cols must be 0, or 2 with a row pitch >= yw + 2");
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
    p.C = C; p.xh = xh; p.xw = xw; p.yh = yh; p.yw = yw; p.zero_cols = zero_pad_cols;
    // phase shift s: the strip's first up-sampled sample is a multiple of `up` away from the padding origin;
    // delta: one extra input column in front so that the column pairs are aligned.
    p.sx = floor_mod(-px0, up);
    p.sy = floor_mod(-py0, up);
    int bx = (-p.sx - px0) / up;            // exact
    p.dx = (bx & 1) ? 1 : 0;
    p.ix0 = bx - p.dx;
    p.iy0 = (-p.sy - py0) / up;
    p.slope = slope; p.out_scale = out_scale;
    {
        // the result is written as aligned column pairs
        const uintptr_t ypair = y_dtype == AFCM_F32 ? 8 : 4;
        if ((yw & 1) || (ys[0] & 1) || (ys[1] & 1) || (ys[2] & 1) || ((uintptr_t)y % ypair) || (skip && ((uintptr_t)skip % ypair))) {
            set_error("filtered_lrelu_tc: y (and skip) must have an even width, even strides and a pair-aligned base address");
            return AFCM_ERR_UNSUPPORTED;
        }
    }
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
        p.kdx[t] = f * out_scale;
        p.kdy[t] = f / u_scale;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (sign_mode != AFCM_SIGN_NONE) {
        p.signs = (uint32_t*)signs; p.sign_h = sign_h; p.sign_nw = sign_wb >> 2; p.s_ox = s_ox; p.s_oy = s_oy; p.amax = amax;
        if (sign_mode == AFCM_SIGN_WRITE) {
            if (act != FTC_ACT_SAT) { set_error("filtered_lrelu_tc_signs: the sign-write kernel needs a clamp in [2^-10, 2^10]"); return AFCM_ERR_UNSUPPORTED; }
            if (up == 2 && down == 2) return launch_tc<2, 2, float, float, FTC_ACT_SAT, false, 1>(p, N, st);
            if (up == 4 && down == 2) return launch_tc<4, 2, float, float, FTC_ACT_SAT, false, 1>(p, N, st);
            return launch_tc<2, 4, float, float, FTC_ACT_SAT, false, 1>(p, N, st);
        }
        if (up == 2 && down == 2) return launch_tc<2, 2, float, float, FTC_ACT_MINMAX, false, 2>(p, N, st);
        if (up == 4 && down == 2) return launch_tc<4, 2, float, float, FTC_ACT_MINMAX, false, 2>(p, N, st);
        return launch_tc<2, 4, float, float, FTC_ACT_MINMAX, false, 2>(p, N, st);
    }
#ifdef AFCM_MINI_U           // development switch: instantiate one kernel only (fast SASS inspection with tools/sass_loops.py)
    return launch_tc<AFCM_MINI_U, AFCM_MINI_D, __half, __half, FTC_ACT_SAT, true>(p, N, st);
#else
    if (x_dtype == AFCM_F32 && y_dtype == AFCM_F32) return dispatch_act<float, float>(p, N, up, down, act, st);
    if (x_dtype == AFCM_F32 && y_dtype == AFCM_F16) return dispatch_act<float, __half>(p, N, up, down, act, st);
    if (x_dtype == AFCM_F16 && y_dtype == AFCM_F32) return dispatch_act<__half, float>(p, N, up, down, act, st);
    // the fast inference path: fp16 planes, bias already added by the convolution epilogue, no skip tensor
    if (!b && !skip && act == FTC_ACT_SAT) return dispatch_geo<__half, __half, FTC_ACT_SAT, true>(p, N, up, down, st);
    return dispatch_act<__half, __half>(p, N, up, down, act, st);
#endif
}
