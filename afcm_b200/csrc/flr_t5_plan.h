// flr_t5_plan.h -- host-side plan of the tcgen05 / TMEM filtered_lrelu (flr_t5.cu): origins of the tile coordinate systems,
// strip / step counts and the tap offsets of the Toeplitz tiles.  tools/flr_t5_emu.py states the same algebra in numpy;
// tests/test_flr_t5_emu.py runs the emulation with THIS plan (through afcm_filtered_lrelu_t5_plan) against the oracle.
//
// One axis of the operator (reference: filtered_lrelu.py:121-153, upfirdn2d.py:186-211), correlation-form taps ku / kd:
//     u[j] = sum_i ku[U i + p0 - j] x[i]         up-sampled sample j  (zero insertion, padding p0 in front, FU taps)
//     y[k] = sum_t kd[t] act(u[D k + t])         output sample k      (FD taps, every D-th kept)
#pragma once
#include "afcm_common.cuh"

namespace afcm {

struct T5Plan {
    int U, D, FU, FD, xh, xw, yh, yw;
    // rows: TMEM lanes = up-sampled rows 128 s .. 128 s + 127 of step s
    int RS;        // new input rows per step (128 / U)
    int K1;        // input rows of one step's window (RS + 16)
    int I0y;       // input row of window row 0 of step 0
    int tuy_e;     // Tuy[m][k] = kuy[tuy_e + U k - m]
    int OS;        // output rows per step (128 / D)
    int wlo0;      // first output row of step 0 (negative: dead rows)
    int nsteps;
    int NL;        // lead chunks of P4 (16 rows each, from the previous step)
    int t4_e;      // T4_e[k][n] = kdy[k - D n + t4_e - 16 e], e = 0 regular
    // columns
    int KW;        // output columns per strip
    int nstrips;
    int iorg0;     // input column of D1 column 0, strip 0 (multiple of 8: 16-byte aligned TMA box start)
    int istep;     // ... advance per strip
    int jorg0;     // up-sampled column of D2 column 0, strip 0
    int korg0;     // output column of D3 column 0, strip 0
    int m0;        // D3 column (= P4 lane) of a strip's first output
    int t3_e;      // T3[k][n] = kdx[k - D n + t3_e]
    int NG;        // groups of 64 up-sampled columns per step
    int N1;        // D1 columns (8 per P2 chunk)
    int halves;    // 64-column blocks of the input tile
};

static inline int t5_cdiv(int a, int b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }       // ceil, b > 0
static inline int t5_fdiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }         // floor, b > 0

static inline void t5_xplan(T5Plan* p, int px0, int kw)
{
    const int U = p->U, D = p->D, FU = p->FU, FD = p->FD;
    const int ineed0 = t5_cdiv(-px0, U);
    p->iorg0 = t5_fdiv(ineed0, 8) * 8;
    p->istep = D * kw / U;
    p->jorg0 = U * p->iorg0 + px0 - FU + 1;
    p->korg0 = t5_cdiv(p->jorg0 - FD + 1, D);
    p->m0 = -p->korg0;
    p->t3_e = p->jorg0 - D * p->korg0;
    const int jlast = D * (kw - 1) + FD - 1;
    const int nr = t5_cdiv(jlast - p->jorg0 + 1, 16);
    p->NG = t5_cdiv(nr, 4);
    p->N1 = 64 * p->NG / U;
    p->halves = (p->N1 + 63) / 64;
    p->KW = kw;
}

// kw = 0: the widest strip that fits (D1 <= 128 columns, valid outputs inside the first 128 D3 columns)
static inline int t5_make_plan(int xh, int xw, int up, int down, int px0, int px1, int py0, int py1, int kw, T5Plan* p)
{
    const bool geo_ok = (up == 2 && down == 2) || (up == 4 && down == 2) || (up == 2 && down == 4);
    if (!geo_ok) { set_error("filtered_lrelu_t5: up=%d down=%d not supported", up, down); return AFCM_ERR_UNSUPPORTED; }
    memset(p, 0, sizeof(*p));
    const int U = up, D = down, FU = 6 * up, FD = 6 * down;
    p->U = U; p->D = D; p->FU = FU; p->FD = FD; p->xh = xh; p->xw = xw;
    const int uw = xw * U + px0 + px1 - (FU - 1) - (FD - 1), uh = xh * U + py0 + py1 - (FU - 1) - (FD - 1);
    if (uw < 1 || uh < 1) { set_error("filtered_lrelu_t5: empty output"); return AFCM_ERR_INVALID; }
    p->yw = (uw + D - 1) / D; p->yh = (uh + D - 1) / D;
    p->RS = 128 / U; p->K1 = p->RS + 16;
    p->I0y = t5_cdiv(-py0, U);
    p->tuy_e = U * p->I0y + py0;
    p->OS = 128 / D;
    p->wlo0 = t5_cdiv(-(FD - 1), D);
    p->nsteps = t5_cdiv(p->yh - p->wlo0, p->OS);
    p->NL = (FD - 1 + 15) / 16;
    p->t4_e = -D * p->wlo0;
    const int kwq = (U == 2 && D == 2) ? 8 : (U == 4 ? 16 : 4);      // strip widths keep the TMA box start 16-byte aligned
    if (kw <= 0) {
        for (int nstrips = 1;; nstrips++) {
            kw = t5_cdiv(t5_cdiv(p->yw, nstrips), kwq) * kwq;
            t5_xplan(p, px0, kw);
            if (p->N1 <= 128 && p->m0 + kw <= 128) break;
            if (kw <= kwq) { set_error("filtered_lrelu_t5: no strip width fits"); return AFCM_ERR_UNSUPPORTED; }
        }
    } else {
        if (kw % kwq) { set_error("filtered_lrelu_t5: strip width %d must be a multiple of %d", kw, kwq); return AFCM_ERR_INVALID; }
        t5_xplan(p, px0, kw);
        if (p->N1 > 128 || p->m0 + kw > 128) { set_error("filtered_lrelu_t5: strip width %d does not fit", kw); return AFCM_ERR_UNSUPPORTED; }
    }
    p->nstrips = t5_cdiv(p->yw, p->KW);
    return AFCM_OK;
}

}  // namespace afcm
