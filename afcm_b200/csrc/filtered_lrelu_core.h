// filtered_lrelu_core.h -- tile geometry and the four FIR passes of the fused filtered_lrelu kernel.
//
// Replaces the reference's filtered_lrelu_kernel (models/networks/stylegan3/torch_utils/ops/
// filtered_lrelu.cu:139-1099) with a from-scratch design:
//   * one CTA = one output tile of one (n,c) plane; runtime tile size, compile-time (UP,FU,DOWN,FD)
//   * taps live in the kernel-parameter constant bank and are indexed at compile time (every tap
//     loop is fully unrolled), so every MAC is one FFMA with a constant-bank operand
//   * the polyphase structure of the zero-insert up-sampling is made compile-time by anchoring the
//     tile's up-sampled coordinate on a multiple of UP relative to the padding (shift sxs/sys) --
//     no per-element modulo, no __constant__ staging kernel, no global filter buffer (stream-safe)
//   * horizontal passes: lanes run down rows, each thread slides a register window along x with
//     128-bit shared-memory loads (row pitches are 4 mod 8 words => conflict-free)
//   * vertical passes: lanes run along x (conflict-free, coalesced global stores)
//
// The passes are written as __host__ __device__ functions of (tid, nthreads) so the exact same code
// is executed by tests/emu (CPU emulation, one "thread" at a time) to validate the index math
// without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AFCM_HD __host__ __device__ __forceinline__
#else
#define AFCM_HD inline
#endif

namespace afcm {

constexpr int FLR_MAX_TAPS = 24;

struct alignas(16) F4 { float x, y, z, w; };

// Everything the kernel needs; passed by value (lives in the constant bank).
struct FlrParams {
    const void* x;            // input  [N,C,xh,xw], element strides below
    void*       y;            // output [N,C,yh,yw]
    const void* b;            // bias [C] (same dtype as x) or null
    const void* skip;         // optional tensor added to the result, same layout as y (engine fusion), or null
    uint8_t*       so;        // sign tensor to write  [N*C, s_h, s_wb] or null
    const uint8_t* si;        // sign tensor to read   [N*C, s_h, s_wb] or null
    long long xs_n, xs_c, xs_h, xs_w;   // x strides (elements)
    long long ys_n, ys_c, ys_h, ys_w;   // y strides (elements); skip uses the same
    int N, C, xh, xw, yh, yw;
    int px0, py0;
    int s_h, s_wb, s_ox, s_oy;          // sign tensor height, width in bytes, read/write offsets
    float gain, slope, clamp, out_scale;
    float ku[FLR_MAX_TAPS];   // up-filter taps as correlation kernel: v[u] = sum_t ku[t]*zp[u+t], already * UP
    float kd[FLR_MAX_TAPS];   // down-filter taps as correlation kernel
    // tile geometry (see flr_make_geom)
    int tow, toh, tiles_x, tiles_y;
    int uwt, uht, ngx, ngy, inw, inh;
    int p_in, p_uh, p_u, p_dh;
    int off_b;                // float offset of buffer B inside dynamic shared memory
    int off_sign;             // byte offset of the sign staging area (sign-write mode), else 0
};

AFCM_HD int flr_ceil_div(int a, int b) { return (a + b - 1) / b; }
AFCM_HD int flr_pitch(int w) { int p = (w + 3) & ~3; return (p & 4) ? p : p + 4; }   // multiple of 4, == 4 mod 8
AFCM_HD int flr_floor_mod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }
AFCM_HD int flr_floor_div(int a, int m) { int q = a / m; return (a % m != 0 && ((a < 0) != (m < 0))) ? q - 1 : q; }

// Fills the geometry fields for an output tile of tow x toh.  G = register-blocking factor.
AFCM_HD int flr_sign_pitch(int uwt) { return ((((uwt + 3) >> 2) + 2) + 3) & ~3; }   // staged sign bytes per row (read mode)

// sign_stage: 0 = no sign staging area, 1 = sign-write mode (one code byte per element), 2 = sign-read mode (the
// tile's part of the packed sign tensor).
template <int UP, int FU, int DOWN, int FD, int G>
inline size_t flr_make_geom(FlrParams& p, int tow, int toh, int sign_stage)
{
    constexpr int R = FU / UP;
    p.tow = tow; p.toh = toh;
    p.tiles_x = flr_ceil_div(p.yw, tow);
    p.tiles_y = flr_ceil_div(p.yh, toh);
    const int towp = flr_ceil_div(tow, G) * G;                 // outputs computed per tile row (padded to G)
    const int tohp = flr_ceil_div(toh, G) * G;
    p.uwt = (towp - 1) * DOWN + FD;                            // up-res extent consumed by the down passes
    p.uht = (tohp - 1) * DOWN + FD;
    p.ngx = flr_ceil_div(flr_ceil_div(p.uwt + UP - 1, UP), G) * G;   // up-res groups (of UP samples) computed
    p.ngy = flr_ceil_div(flr_ceil_div(p.uht + UP - 1, UP), G) * G;
    p.inw = p.ngx + R;                                         // input samples needed
    p.inh = p.ngy + R;
    p.p_in = flr_pitch(p.inw + 3);                             // +3: 128-bit window loads may over-read
    p.p_uh = flr_pitch(p.ngx * UP);
    p.p_u  = flr_pitch(p.uwt > p.ngx * UP ? p.uwt : p.ngx * UP);
    p.p_dh = flr_pitch(towp);
    const int rows_u = p.uht > p.ngy * UP ? p.uht : p.ngy * UP;
    const size_t a = (size_t)(p.inh * p.p_in > rows_u * p.p_u ? p.inh * p.p_in : rows_u * p.p_u);
    const size_t b = (size_t)(p.inh * p.p_uh > p.uht * p.p_dh ? p.inh * p.p_uh : p.uht * p.p_dh);
    p.off_b = (int)a;
    size_t bytes = (a + b) * sizeof(float);
    p.off_sign = 0;
    if (sign_stage == 1) { p.off_sign = (int)bytes; bytes += (size_t)p.uht * (size_t)((p.uwt + 3) & ~3); }
    if (sign_stage == 2) { p.off_sign = (int)bytes; bytes += (size_t)p.uht * (size_t)flr_sign_pitch(p.uwt); }
    return bytes;
}

// Per-tile runtime origin.
struct FlrTile {
    int plane, n, c;
    int ox0, oy0;         // first output sample of the tile
    int ux0, uy0;         // first up-res sample the down passes consume (= ox0*DOWN)
    int sxs, sys;         // extra up-res samples computed in front so phases are compile-time
    int ibx, iby;         // first input sample held in s_in
};

template <int UP, int DOWN>
AFCM_HD FlrTile flr_tile3(const FlrParams& p, int plane, int tx, int ty)
{
    FlrTile t;
    t.plane = plane;
    t.n = t.plane / p.C; t.c = t.plane - t.n * p.C;
    t.ox0 = tx * p.tow; t.oy0 = ty * p.toh;
    t.ux0 = t.ox0 * DOWN; t.uy0 = t.oy0 * DOWN;
    t.sxs = flr_floor_mod(t.ux0 - p.px0, UP);
    t.sys = flr_floor_mod(t.uy0 - p.py0, UP);
    t.ibx = flr_floor_div(t.ux0 - t.sxs - p.px0, UP);
    t.iby = flr_floor_div(t.uy0 - t.sys - p.py0, UP);
    return t;
}

template <int UP, int DOWN>
AFCM_HD FlrTile flr_tile(const FlrParams& p, int tile_linear)
{
    const int tx = tile_linear % p.tiles_x;
    const int r  = tile_linear / p.tiles_x;
    return flr_tile3<UP, DOWN>(p, r / p.tiles_y, tx, r % p.tiles_y);
}

// ---- pass 0: global -> s_in, bias added to real pixels only, zeros elsewhere -------------------
// Flat item index -> (row, column) without a division per item: one division at the start, then the pair advances by
// the (constant) thread count.
struct FlrRowCol {
    int row, col, drow, dcol, pitch;
    AFCM_HD FlrRowCol(int first, int step, int pitch_) : pitch(pitch_)
    {
        row = first / pitch_; col = first - row * pitch_;
        drow = step / pitch_; dcol = step - drow * pitch_;
    }
    AFCM_HD void next()
    {
        row += drow; col += dcol;
        if (col >= pitch) { col -= pitch; row++; }
    }
};

template <typename T>
AFCM_HD void flr_pass_load(int tid, int nthr, const FlrParams& p, const FlrTile& t, float* s_in)
{
    const T* xp = (const T*)p.x + t.n * p.xs_n + t.c * p.xs_c;
    const float bias = p.b ? (float)((const T*)p.b)[t.c] : 0.f;
    const int n = p.inh * p.p_in;
    FlrRowCol rc(tid, nthr, p.p_in);
    const T* x0 = xp + t.iby * p.xs_h + t.ibx * p.xs_w;          // element (0,0) of the tile (may lie outside the plane)
    const int sh = (int)p.xs_h, sw = (int)p.xs_w;                // offsets inside a plane fit 32 bits (checked by the host)
    // Four items per round, all loads before the first store: a shared-memory store between two global loads is a
    // possible alias for the compiler and serialises the pass into dependent load -> store round trips.
    constexpr int B = 4;
    if (t.iby >= 0 && t.iby + p.inh <= p.xh && t.ibx >= 0 && t.ibx + p.inw <= p.xw) {
        // interior tile: every input sample exists, only the pitch padding columns are zero
        for (int i = tid; i < n; i += B * nthr) {
            float v[B];
            FlrRowCol r2 = rc;
            for (int k = 0; k < B; k++, r2.next())
                v[k] = (i + k * nthr < n && r2.col < p.inw) ? (float)x0[r2.row * sh + r2.col * sw] + bias : 0.f;
            for (int k = 0; k < B; k++)
                if (i + k * nthr < n) s_in[i + k * nthr] = v[k];
            rc = r2;
        }
    } else {
        for (int i = tid; i < n; i += B * nthr) {
            float v[B];
            FlrRowCol r2 = rc;
            for (int k = 0; k < B; k++, r2.next()) {
                const int iy = r2.row, ix = r2.col;
                const int gy = t.iby + iy, gx = t.ibx + ix;
                v[k] = (i + k * nthr < n && ix < p.inw && gy >= 0 && gy < p.xh && gx >= 0 && gx < p.xw)
                           ? (float)x0[iy * sh + ix * sw] + bias : 0.f;
            }
            for (int k = 0; k < B; k++)
                if (i + k * nthr < n) s_in[i + k * nthr] = v[k];
            rc = r2;
        }
    }
}

// ---- pass 0b (sign-read mode): stage the tile's part of the packed sign tensor (4 codes per byte) in shared memory
// with coalesced byte loads; bytes outside the tensor read as 0 = "leave the value alone" (same as the bounds test of
// the reference kernel, filtered_lrelu.cu:562-572).  Staged byte (ly, b) = sign byte (uy0 + ly + s_oy, eb0 + b).
AFCM_HD void flr_pass_sign_load(int tid, int nthr, const FlrParams& p, const FlrTile& t, uint8_t* s_sign)
{
    const int nbp = flr_sign_pitch(p.uwt);
    const int eb0 = flr_floor_div(t.ux0 + p.s_ox, 4);
    const int items = p.uht * nbp;
    const uint8_t* base = p.si + (long long)t.plane * p.s_h * p.s_wb;
    FlrRowCol rc(tid, nthr, nbp);
    for (int it = tid; it < items; it += nthr, rc.next()) {
        const int ey = t.uy0 + rc.row + p.s_oy, eb = eb0 + rc.col;
        uint8_t v = 0;
        if (ey >= 0 && ey < p.s_h && eb >= 0 && eb < p.s_wb) v = base[ey * p.s_wb + eb];
        s_sign[it] = v;
    }
}

// ---- pass 1: horizontal polyphase up-FIR.  s_in[inh][p_in] -> s_uh[inh][p_uh] -------------------
// group g (UP consecutive up-res samples starting at local index g*UP):
//   q == 0 : sum_r ku[UP*r]        * in[g + r]
//   q  > 0 : sum_r ku[UP-q + UP*r] * in[g + 1 + r]
template <int UP, int FU, int G>
AFCM_HD void flr_pass_hup(int tid, int nthr, const FlrParams& p, const float* s_in, float* s_uh)
{
    constexpr int R = FU / UP;
    constexpr int NW = (G + R + 3) / 4 * 4;
    const int nchunk = p.ngx / G;
    const int items = nchunk * p.inh;
    FlrRowCol rc(tid, nthr, p.inh);
    for (int it = tid; it < items; it += nthr, rc.next()) {
        const int ch = rc.row, iy = rc.col;
        const float* src = s_in + iy * p.p_in + ch * G;
        float w[NW];
#pragma unroll
        for (int j = 0; j < NW; j += 4) {
            const F4 v = *reinterpret_cast<const F4*>(src + j);
            w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
        }
        float o[G * UP];
#pragma unroll
        for (int g = 0; g < G; g++) {
#pragma unroll
            for (int q = 0; q < UP; q++) {
                float acc = 0.f;
#pragma unroll
                for (int r = 0; r < R; r++)
                    acc += (q == 0 ? p.ku[UP * r] : p.ku[UP - q + UP * r]) * w[g + (q == 0 ? 0 : 1) + r];
                o[g * UP + q] = acc;
            }
        }
        float* dst = s_uh + iy * p.p_uh + ch * G * UP;
#pragma unroll
        for (int j = 0; j < G * UP; j += 4) {
            F4 v; v.x = o[j]; v.y = o[j + 1]; v.z = o[j + 2]; v.w = o[j + 3];
            *reinterpret_cast<F4*>(dst + j) = v;
        }
    }
}

// ---- pass 2: vertical polyphase up-FIR + gain + leaky ReLU + clamp (+ signs) ---------------------
// s_uh[inh][p_uh] -> s_u[uht][p_u], storing at (ly - sys, lx - sxs).
// SIGN: 0 none, 1 write (2-bit codes staged as bytes in s_sign[uht][uwt4]), 2 read from p.si.
// One item of pass 2: G*UP consecutive up-res rows of one column.  CHECK = the rows may fall outside [0, uht).
template <int UP, int FU, int G, int SIGN, bool CHECK>
AFCM_HD void flr_vup_rows(const FlrParams& p, const float* w, int ly0, float* dst, uint8_t* sdst, const uint8_t* srow,
                          int sshift, int nbp, int uwt4)
{
    constexpr int R = FU / UP;
#pragma unroll
    for (int g = 0; g < G; g++) {
#pragma unroll
        for (int q = 0; q < UP; q++) {
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < R; r++)
                acc += (q == 0 ? p.ku[UP * r] : p.ku[UP - q + UP * r]) * w[g + (q == 0 ? 0 : 1) + r];
            const int e = g * UP + q;                       // row offset inside the item (compile time)
            const bool ok = !CHECK || (ly0 + e >= 0 && ly0 + e < p.uht);
            float v = acc * p.gain;
            if (SIGN == 2) {
                if (ok) {
                    const int s = srow[e * nbp] >> sshift;
                    if (s & 1) v *= p.slope;
                    if (s & 2) v = 0.f;
                }
            } else {
                int s = 0;
                if (v < 0.f) { v *= p.slope; s = 1; }
                if (v > p.clamp) { v = p.clamp; s = 2; }
                if (v < -p.clamp) { v = -p.clamp; s = 2; }
                if (SIGN == 1 && ok) sdst[e * uwt4] = (uint8_t)s;
            }
            if (ok) dst[e * p.p_u] = v;
        }
    }
}

template <int UP, int FU, int G, int SIGN>
AFCM_HD void flr_pass_vup(int tid, int nthr, const FlrParams& p, const FlrTile& t,
                          const float* s_uh, float* s_u, uint8_t* s_sign)
{
    constexpr int R = FU / UP;
    const int ncols = p.ngx * UP;
    const int nchunk = p.ngy / G;
    const int items = nchunk * ncols;
    const int uwt4 = (p.uwt + 3) & ~3;
    const int nbp = SIGN == 2 ? flr_sign_pitch(p.uwt) : 0;
    // sign-read mode: element lx of a staged row sits in byte ((se0 + lx) >> 2), bit pair ((se0 + lx) & 3)
    const int se0 = SIGN == 2 ? t.ux0 + p.s_ox - 4 * flr_floor_div(t.ux0 + p.s_ox, 4) : 0;
    FlrRowCol rc(tid, nthr, ncols);
    for (int it = tid; it < items; it += nthr, rc.next()) {
        const int ch = rc.row, col = rc.col;
        const int lx = col - t.sxs;
        if (lx < 0 || lx >= p.uwt) continue;            // column outside the region the down passes consume
        const float* src = s_uh + (ch * G) * p.p_uh + col;
        float w[G + R];
#pragma unroll
        for (int j = 0; j < G + R; j++) w[j] = src[j * p.p_uh];
        const int ly0 = ch * G * UP - t.sys;            // first up-res row of the item (tile-local)
        // the pointers below are only dereferenced at rows that pass the range test
        float* dst = s_u + ly0 * p.p_u + lx;
        uint8_t* sdst = s_sign + ly0 * uwt4 + lx;
        const uint8_t* srow = s_sign + ly0 * nbp + ((se0 + lx) >> 2);
        const int sshift = ((se0 + lx) & 3) << 1;
        if (ly0 >= 0 && ly0 + G * UP <= p.uht)
            flr_vup_rows<UP, FU, G, SIGN, false>(p, w, ly0, dst, sdst, srow, sshift, nbp, uwt4);
        else
            flr_vup_rows<UP, FU, G, SIGN, true>(p, w, ly0, dst, sdst, srow, sshift, nbp, uwt4);
    }
}

// ---- sign flush: pack 4 staged codes per byte and write the tile's part of the sign tensor ------
template <int DOWN>
AFCM_HD void flr_pass_sign_flush(int tid, int nthr, const FlrParams& p, const FlrTile& t, const uint8_t* s_sign)
{
    const int uwt4 = (p.uwt + 3) & ~3;          // ux0 is a multiple of 4 (tow*DOWN % 4 == 0 enforced by the host)
    const int nb = uwt4 >> 2;
    const int items = p.uht * nb;
    const int eb0 = (t.ux0 + p.s_ox) >> 2;
    const bool last_col = t.ox0 + p.tow >= p.yw;
    uint8_t* base = p.so + (long long)t.plane * p.s_h * p.s_wb;
    FlrRowCol rc(tid, nthr, nb);
    for (int it = tid; it < items; it += nthr, rc.next()) {
        const int bx = rc.col;
        const int ey = t.uy0 + rc.row + p.s_oy, eb = eb0 + bx;
        if (ey < 0 || ey >= p.s_h || eb < 0 || eb >= p.s_wb) continue;
        // four staged code bytes (values 0..2) as one aligned word (it * 4 == ly * uwt4 + bx * 4); bytes at or beyond
        // uwt were never written
        uint32_t w = *reinterpret_cast<const uint32_t*>(s_sign + it * 4);
        const int nvalid = p.uwt - bx * 4;
        if (nvalid < 4) w &= (1u << (8 * nvalid)) - 1u;
        const int code = (int)((w & 3u) | ((w >> 6) & 0xcu) | ((w >> 12) & 0x30u) | ((w >> 18) & 0xc0u));
        // elements beyond this tile's uwt belong to the next tile; tiles overlap by FD-DOWN >= 4 samples, so only whole
        // bytes that this tile fully owns or that are the ragged end of the row are partial: partial bytes at a tile's
        // right edge are skipped unless this is the last tile column.
        if (nvalid >= 4 || last_col) base[ey * p.s_wb + eb] = (uint8_t)code;
    }
}

// ---- pass 3: horizontal down-FIR with decimation.  s_u[uht][p_u] -> s_dh[uht][p_dh] --------------
template <int DOWN, int FD, int G>
AFCM_HD void flr_pass_hdown(int tid, int nthr, const FlrParams& p, const float* s_u, float* s_dh)
{
    constexpr int NW = ((G - 1) * DOWN + FD + 3) / 4 * 4;
    const int nchunk = flr_ceil_div(p.tow, G);
    const int items = nchunk * p.uht;
    FlrRowCol rc(tid, nthr, p.uht);
    for (int it = tid; it < items; it += nthr, rc.next()) {
        const int ch = rc.row, uy = rc.col;
        const float* src = s_u + uy * p.p_u + ch * G * DOWN;
        float w[NW];
#pragma unroll
        for (int j = 0; j < NW; j += 4) {
            const F4 v = *reinterpret_cast<const F4*>(src + j);
            w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
        }
        float o[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < FD; k++) acc += p.kd[k] * w[g * DOWN + k];
            o[g] = acc;
        }
        float* dst = s_dh + uy * p.p_dh + ch * G;
#pragma unroll
        for (int j = 0; j < G; j += 4) {
            F4 v; v.x = o[j]; v.y = o[j + 1]; v.z = o[j + 2]; v.w = o[j + 3];
            *reinterpret_cast<F4*>(dst + j) = v;
        }
    }
}

// ---- pass 4: vertical down-FIR with decimation and the global store -----------------------------
template <typename T, int DOWN, int FD, int G>
AFCM_HD void flr_pass_vdown(int tid, int nthr, const FlrParams& p, const FlrTile& t, const float* s_dh)
{
    constexpr int NW = (G - 1) * DOWN + FD;
    const int towp = flr_ceil_div(p.tow, G) * G;
    const int nchunk = flr_ceil_div(p.toh, G);
    const int items = nchunk * towp;
    T* yp = (T*)p.y + t.n * p.ys_n + t.c * p.ys_c;
    const T* kp = p.skip ? (const T*)p.skip + t.n * p.ys_n + t.c * p.ys_c : nullptr;
    FlrRowCol rc(tid, nthr, towp);
    for (int it = tid; it < items; it += nthr, rc.next()) {
        const int ch = rc.row, ox = rc.col;
        const int gx = t.ox0 + ox;
        if (ox >= p.tow || gx >= p.yw) continue;         // padded tile column / beyond the plane: nothing to store
        const float* src = s_dh + (ch * G * DOWN) * p.p_dh + ox;
        float w[NW];
#pragma unroll
        for (int j = 0; j < NW; j++) w[j] = src[j * p.p_dh];
        const int r0 = ch * G, gy0 = t.oy0 + r0;
        const int nrow = (p.toh - r0 < p.yh - gy0 ? p.toh - r0 : p.yh - gy0);      // rows of this item that exist
        const int ysh = (int)p.ys_h;
        const int o0 = gy0 * ysh + gx * (int)p.ys_w;
#pragma unroll
        for (int g = 0; g < G; g++) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < FD; k++) acc += p.kd[k] * w[g * DOWN + k];
            if (g < nrow) {
                const int o = o0 + g * ysh;
                if (kp) acc += (float)kp[o];
                yp[o] = (T)(acc * p.out_scale);
            }
        }
    }
}

}  // namespace afcm
