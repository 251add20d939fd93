// flr_tc.cu -- filtered_lrelu on the tensor cores (sm_100a): the four separable FIR passes of
// bias -> up-FIR -> gain/leaky-ReLU/clamp -> down-FIR (reference: models/networks/stylegan3/torch_utils/
// ops/filtered_lrelu.py:121-153, filtered_lrelu.cu:139-1099) evaluated as a chain of small banded-Toeplitz
// matrix products with mma.sync.m16n8k16 (fp16 operands, fp32 accumulation), entirely in registers.
//
// Why: the op does 17-46 FLOP per algorithmic byte, above the FP32-SIMT ridge of B200, so a CUDA-core
// kernel cannot approach the HBM roofline.  Here the MACs run on the tensor pipe
// and the FP32 pipe only does the nonlinearity.
//
// Algorithm (one warp = one strip of 16 output columns, streamed top to bottom):
//   in[y][x]  --(1) R1[j,y] = sum_x Tu_x[j,x] in[y,x]      const A, data B (global -> regs)
//             --(2) R2[v,j] = sum_y Tu_y[v,y] R1[j,y]      const A, B = C-fragments of (1)   (transposing reuse)
//             --    gain (folded into Tu_y), leaky ReLU in fp32, clamp by saturating fp16 conversion
//             --(4) R3[k,v] = sum_j Td_x[k,j] R2[v,j]      const A, B = C-fragments of (2)
//             --(5) R4[k,w] = sum_v R3[k,v] Td_y[w,v]      A = C-fragments of (4), const B
//   j,v = up-sampled coordinates, k,w = output coordinates.  The m16n8 accumulator layout of one product
//   is exactly the B (or A) operand layout of the next, so no shared memory, shuffles or block barriers
//   are needed between passes; the Toeplitz operands are shift invariant, so each thread keeps a handful
//   of constant fragments (the FIR taps) in registers for the whole strip.  The polyphase structure of
//   the zero-insertion is folded into Tu (only every UP-th column of a row is non-zero).
//   Index conventions (s = phase shift, delta = input alignment, origins) are modelled in
//   tools/flr_tc_model.py and pinned against the oracle by tests/test_flr_tc_model.py.
//
// Numerics: operands (activations, intermediates, taps) are rounded to fp16, sums are fp32.  Measured
// against the fp32 oracle: max error ~5e-4 of max|y| per call (tests/test_gpu_flr_tc.py states the bound).
// This is the "tensor-core path with stated tolerance" of the north star; afcm_filtered_lrelu remains the
// exact-fp32 path.
#include <cuda_fp16.h>
#include "afcm_common.cuh"

namespace afcm {

constexpr int FTC_WARPS = 4;
constexpr int FTC_TAB = 192;          // tap tables: index t + 64, zero padded
constexpr int FTC_TAB_OFS = 64;

struct FlrTcParams {
    const void* x; void* y; const float* b; const void* skip;
    long long xs_n, xs_c, xs_h, ys_n, ys_c, ys_h;        // element strides, innermost stride 1
    int C, xh, xw, yh, yw;
    int strips, segs, seg_wblocks;                        // 16-column strips, row segments of 8*seg_wblocks rows
    long long total_warps;
    int ix0, iy0, iy_step;                                // input origin: col = strip*IXS + ix0, row = seg*iy_step + iy0
    int sx, sy, dx;
    int slope_gt1;
    float slope, out_scale, sat_scale;                    // sat_scale: R2 is scaled so that clamp == fp16 max (or 1)
    float kux[24], kuy[24], kdx[24], kdy[24];             // correlation-form taps (kuy carries gain*sat_scale, kdx 1/sat_scale)
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
// saturating conversion: |v| >= 65504 -> +-65504 (implements the clamp, see sat_scale)
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int U, int D, typename TIN, typename TOUT>
struct FtcWarp {
    using Geo = FtcGeo<U, D>;
    // constant fragments (the FIR taps)
    uint32_t a1[Geo::NPH][4], a2[Geo::NPH][4], a4[Geo::KC4][4], b5[Geo::NB5][2];
    const FlrTcParams& p;
    int g, t;
    // strip state
    const TIN* xp; TOUT* yp; const TOUT* kp;
    float bias;
    int ix, iy, k0, w0, nwb;
    unsigned colmask;            // bit 2c / 2c+1: column validity of the two elements of chunk c

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
    }

    __device__ void begin_strip(long long wid)
    {
        const int strip = (int)(wid % p.strips);
        const long long r = wid / p.strips;
        const int seg = (int)(r % p.segs);
        const long long plane = r / p.segs;
        const int n = (int)(plane / p.C), c = (int)(plane - (long long)n * p.C);
        xp = (const TIN*)p.x + n * p.xs_n + c * p.xs_c;
        yp = (TOUT*)p.y + n * p.ys_n + c * p.ys_c;
        kp = p.skip ? (const TOUT*)p.skip + n * p.ys_n + c * p.ys_c : nullptr;
        bias = p.b ? p.b[c] : 0.f;
        ix = strip * Geo::IXS + p.ix0;
        iy = seg * p.iy_step + p.iy0;
        k0 = strip * 16;
        w0 = seg * p.seg_wblocks * 8;
        const int rows_left = p.yh - w0;
        nwb = (rows_left + 7) >> 3;
        if (nwb > p.seg_wblocks) nwb = p.seg_wblocks;
        colmask = 0;
#pragma unroll
        for (int c8 = 0; c8 < Geo::NC; c8++) {
            const int col = ix + 8 * c8 + 2 * t;
            if (col >= 0 && col < p.xw) colmask |= 1u << (2 * c8);
            if (col + 1 >= 0 && col + 1 < p.xw) colmask |= 2u << (2 * c8);
        }
    }

    // raw loads of one block of 8 input rows (this thread: row g, NC column pairs)
    __device__ __forceinline__ void load_raw(int yb, float2 (&raw)[Geo::NC]) const
    {
        const int row = iy + 8 * yb + g;
        const bool rok = row >= 0 && row < p.xh;
        const TIN* rp = xp + (long long)row * p.xs_h + ix + 2 * t;
#pragma unroll
        for (int c8 = 0; c8 < Geo::NC; c8++) {
            const unsigned m = rok ? (colmask >> (2 * c8)) & 3u : 0u;
            float2 v = make_float2(0.f, 0.f);
            const TIN* q = rp + 8 * c8;
            if (sizeof(TIN) == 4) {
                if (m == 3u && ((reinterpret_cast<uintptr_t>(q) & 7) == 0)) {
                    v = *reinterpret_cast<const float2*>(q);
                } else {
                    if (m & 1u) v.x = (float)q[0];
                    if (m & 2u) v.y = (float)q[1];
                }
            } else {
                if (m == 3u && ((reinterpret_cast<uintptr_t>(q) & 3) == 0)) {
                    v = __half22float2(*reinterpret_cast<const __half2*>(q));
                } else {
                    if (m & 1u) v.x = (float)q[0];
                    if (m & 2u) v.y = (float)q[1];
                }
            }
            if (m & 1u) v.x += bias;
            if (m & 2u) v.y += bias;
            raw[c8] = v;
        }
    }

    // (1) horizontal up-FIR of one block of 8 input rows: r1[nb] = B-fragment half for column block nb of R2
    __device__ __forceinline__ void step1(const float2 (&raw)[Geo::NC], uint32_t (&r1)[Geo::NJ8]) const
    {
        uint32_t in[Geo::NC];
#pragma unroll
        for (int c8 = 0; c8 < Geo::NC; c8++) in[c8] = pack_h2(raw[c8].x, raw[c8].y);
#pragma unroll
        for (int b = 0; b < Geo::KC4; b++) {
            const int w = (U == 2) ? b : (b >> 1);          // first input chunk of the window
            const int ph = (U == 2) ? 0 : (b & 1);
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            mma16816(c, a1[ph], in[w], in[w + 1]);
            r1[2 * b] = pack_h2(c[0], c[1]);
            r1[2 * b + 1] = pack_h2(c[2], c[3]);
        }
    }

    // (2)+(nonlinearity)+(4): one block of 16 up-sampled rows -> A-fragment chunk of R3 for step (5)
    __device__ __forceinline__ void vblock(const uint32_t (&ra)[Geo::NJ8], const uint32_t (&rb)[Geo::NJ8], int ph,
                                           uint32_t (&a5)[4]) const
    {
        float c3[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int kc = 0; kc < Geo::KC4; kc++) {
            uint32_t p2[2][2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int nb = 2 * kc + q;
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                mma16816(c, a2[ph], ra[nb], rb[nb]);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float s = c[i] * p.slope;
                    c[i] = p.slope_gt1 ? fminf(c[i], s) : fmaxf(c[i], s);
                }
                p2[q][0] = pack_h2_sat(c[0], c[1]);
                p2[q][1] = pack_h2_sat(c[2], c[3]);
            }
            mma16816(c3[0], a4[kc], p2[0][0], p2[1][0]);
            mma16816(c3[1], a4[kc], p2[0][1], p2[1][1]);
        }
        a5[0] = pack_h2(c3[0][0], c3[0][1]);
        a5[1] = pack_h2(c3[0][2], c3[0][3]);
        a5[2] = pack_h2(c3[1][0], c3[1][1]);
        a5[3] = pack_h2(c3[1][2], c3[1][3]);
    }

    __device__ __forceinline__ void store(int wb, const float (&c)[4]) const
    {
        const int y0 = w0 + 8 * wb + 2 * t;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int yy = y0 + (i & 1), xx = k0 + g + (i >> 1) * 8;
            if (yy < p.yh && xx < p.yw) {
                const long long o = (long long)yy * p.ys_h + xx;
                float v = c[i];
                if (kp) v += (float)kp[o];
                yp[o] = (TOUT)(v * p.out_scale);
            }
        }
    }

    // (5) one block of 8 output rows from NB5 consecutive chunks of R3
    __device__ __forceinline__ void wblock2(int wb, const uint32_t (&c0)[4], const uint32_t (&c1)[4]) const
    {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(c, c0, b5[0][0], b5[0][1]);
        mma16816(c, c1, b5[1][0], b5[1][1]);
        store(wb, c);
    }
    __device__ __forceinline__ void wblock4(int wb, const uint32_t (&c0)[4], const uint32_t (&c1)[4], const uint32_t (&c2)[4],
                                            const uint32_t (&c3)[4]) const
    {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(c, c0, b5[0][0], b5[0][1]);
        mma16816(c, c1, b5[1][0], b5[1][1]);
        mma16816(c, c2, b5[Geo::NB5 - 2][0], b5[Geo::NB5 - 2][1]);
        mma16816(c, c3, b5[Geo::NB5 - 1][0], b5[Geo::NB5 - 1][1]);
        store(wb, c);
    }

    __device__ void run()
    {
        float2 raw[Geo::NC];
        if (U == 2 && D == 2) {
            // iteration it: input block it -> R1[it]; v-block it-1 from R1[it-1], R1[it]; w-block it-2 from chunks it-2, it-1
            uint32_t r1p[Geo::NJ8], r1c[Geo::NJ8], a5p[4], a5c[4];
            const int iters = nwb + 2;
            load_raw(0, raw);
            for (int it = 0; it < iters; it++) {
                step1(raw, r1c);
                if (it + 1 < iters) load_raw(it + 1, raw);
                if (it >= 1) vblock(r1p, r1c, 0, a5c);
                if (it >= 2) wblock2(it - 2, a5p, a5c);
#pragma unroll
                for (int i = 0; i < Geo::NJ8; i++) r1p[i] = r1c[i];
#pragma unroll
                for (int i = 0; i < 4; i++) a5p[i] = a5c[i];
            }
        } else if (U == 4 && D == 2) {
            // iteration it: input block it; v-blocks 2it-2, 2it-1 (two phases) from R1[it-1], R1[it];
            // w-block 2it-3 from chunks (2it-3, 2it-2), w-block 2it-2 from chunks (2it-2, 2it-1)
            uint32_t r1p[Geo::NJ8], r1c[Geo::NJ8], a5l[4], a5a[4], a5b[4];
            const int iters = (nwb + 2) / 2 + 1;
            load_raw(0, raw);
            for (int it = 0; it < iters; it++) {
                step1(raw, r1c);
                if (it + 1 < iters) load_raw(it + 1, raw);
                if (it >= 1) {
                    vblock(r1p, r1c, 0, a5a);
                    vblock(r1p, r1c, Geo::NPH - 1, a5b);
                    if (it >= 2 && 2 * it - 3 < nwb) wblock2(2 * it - 3, a5l, a5a);
                    if (2 * it - 2 < nwb) wblock2(2 * it - 2, a5a, a5b);
#pragma unroll
                    for (int i = 0; i < 4; i++) a5l[i] = a5b[i];
                }
#pragma unroll
                for (int i = 0; i < Geo::NJ8; i++) r1p[i] = r1c[i];
            }
        } else {
            // U == 2, D == 4.  iteration it: input blocks 2it, 2it+1; v-block 2it-1 from R1[2it-1], R1[2it];
            // v-block 2it from R1[2it], R1[2it+1]; w-block it-2 from chunks 2it-4 .. 2it-1
            uint32_t r1l[Geo::NJ8], r1a[Geo::NJ8], r1b[Geo::NJ8], q0[4], q1[4], q2[4], qa[4], qb[4];
            const int iters = nwb + 2;
            for (int it = 0; it < iters; it++) {
                load_raw(2 * it, raw);
                step1(raw, r1a);
                load_raw(2 * it + 1, raw);
                step1(raw, r1b);
                if (it >= 1) vblock(r1l, r1a, 0, qa);
                vblock(r1a, r1b, 0, qb);
                if (it >= 2) wblock4(it - 2, q0, q1, q2, qa);
#pragma unroll
                for (int i = 0; i < 4; i++) { q0[i] = q2[i]; q1[i] = qa[i]; q2[i] = qb[i]; }
#pragma unroll
                for (int i = 0; i < Geo::NJ8; i++) r1l[i] = r1b[i];
            }
        }
    }
};

template <int U, int D, typename TIN, typename TOUT>
__global__ void __launch_bounds__(FTC_WARPS * 32)
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
    FtcWarp<U, D, TIN, TOUT> w(p, lane);
    w.load_consts(tab);
    w.begin_strip(wid);
    w.run();
}

static int floor_mod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

template <int U, int D, typename TIN, typename TOUT>
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
    flr_tc_kernel<U, D, TIN, TOUT><<<(unsigned)blocks, FTC_WARPS * 32, 0, st>>>(p);
    AFCM_LAUNCH_CHECK();
    count_launch();
    return AFCM_OK;
}

template <typename TIN, typename TOUT>
static int dispatch_tc(FlrTcParams& p, int N, int up, int down, cudaStream_t st)
{
    if (up == 2 && down == 2) return launch_tc<2, 2, TIN, TOUT>(p, N, st);
    if (up == 4 && down == 2) return launch_tc<4, 2, TIN, TOUT>(p, N, st);
    if (up == 2 && down == 4) return launch_tc<2, 4, TIN, TOUT>(p, N, st);
    return AFCM_ERR_UNSUPPORTED;
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
    FlrTcParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.y = y; p.b = b; p.skip = skip;
    p.xs_n = xs[0]; p.xs_c = xs[1]; p.xs_h = xs[2];
    p.ys_n = ys[0]; p.ys_c = ys[1]; p.ys_h = ys[2];
    p.C = C; p.xh = xh; p.xw = xw; p.yh = yh; p.yw = yw;
    // phase shift s: the strip's first up-sampled sample is a multiple of `up` away from the padding origin;
    // delta: one extra input column in front so that the fp32 pair loads are 8-byte aligned.
    p.sx = floor_mod(-px0, up);
    p.sy = floor_mod(-py0, up);
    int bx = (-p.sx - px0) / up;            // exact
    p.dx = (bx & 1) ? 1 : 0;
    p.ix0 = bx - p.dx;
    p.iy0 = (-p.sy - py0) / up;
    p.slope = slope; p.slope_gt1 = slope > 1.f; p.out_scale = out_scale;
    const bool finite_clamp = clamp > 0.f && clamp < 3.0e38f;
    p.sat_scale = finite_clamp ? 65504.f / clamp : 1.f;
    for (int t = 0; t < fu_taps; t++) {
        const float f = fu_host[flip_filter ? t : fu_taps - 1 - t] * (float)up;
        p.kux[t] = f;
        p.kuy[t] = f * gain * p.sat_scale;
    }
    for (int t = 0; t < fd_taps; t++) {
        const float f = fd_host[flip_filter ? t : fd_taps - 1 - t];
        p.kdx[t] = f / p.sat_scale;
        p.kdy[t] = f;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == AFCM_F32 && y_dtype == AFCM_F32) return dispatch_tc<float, float>(p, N, up, down, st);
    if (x_dtype == AFCM_F32 && y_dtype == AFCM_F16) return dispatch_tc<float, __half>(p, N, up, down, st);
    if (x_dtype == AFCM_F16 && y_dtype == AFCM_F32) return dispatch_tc<__half, float>(p, N, up, down, st);
    return dispatch_tc<__half, __half>(p, N, up, down, st);
}
