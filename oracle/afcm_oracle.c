/*
 * afcm_oracle.c -- CPU restatement (plain C, float32) of the AFCM generator hot-path operators.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it, and only as the checker or the
 * CPU baseline.  The product path (afcm_b200/) never calls into this file.
 *
 * Each function restates the reference algorithm it cites (paths relative to /root/reference,
 * OPS = models/networks/stylegan3/torch_utils/ops):
 *   orc_bias_act        OPS/bias_act.py:91-120      (_bias_act_ref), grads OPS/bias_act.cu:23-147
 *   orc_upfirdn2d       OPS/upfirdn2d.py:167-211    (_upfirdn2d_ref)
 *   orc_filtered_lrelu  OPS/filtered_lrelu.py:121-153 (_filtered_lrelu_ref) plus the sign-tensor
 *                       semantics of the native op: OPS/filtered_lrelu.cpp:87-94 (shape),
 *                       OPS/filtered_lrelu.cu:1127-1146 (write: bit0 negative, bit1 clamped) and
 *                       OPS/filtered_lrelu.cu:1173-1189 (read: slope where bit0, zero where bit1)
 *   orc_conv2d          torch.nn.functional.conv2d semantics (correlation, zero padding) as called
 *                       from OPS/conv2d_gradfix.py:37-40
 *   orc_modulated_conv2d  models/networks/stylegan3/networks_stylegan3.py:25-64
 *   orc_fully_connected   networks_stylegan3.py:89-101
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * oracle is pinned against outputs of the reference's own Python `_ref` path executed in the
 * authoring container (tests/golden/gen_golden.py -> tests/golden/ npz files), see
 * tests/test_oracle_golden.py.
 *
 * All tensors are dense NCHW float32.  Accumulation is done in float32 in a fixed order so results
 * are deterministic; they agree with the torch CPU reference to float32 rounding (not bit-exact,
 * because torch's conv kernels use a different summation order).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* bias_act: y = clamp(act(x + b) * gain).  Activation indices follow OPS/bias_act.py:21-31.   */

static float orc_act_eval(float x, int act, float alpha)
{
    switch (act) {
    case 1: return x;                                            /* linear   */
    case 2: return x > 0.f ? x : 0.f;                            /* relu     */
    case 3: return x > 0.f ? x : x * alpha;                      /* lrelu    */
    case 4: return tanhf(x);                                     /* tanh     */
    case 5: return 1.f / (1.f + expf(-x));                       /* sigmoid  */
    case 6: return x > 0.f ? x : expf(x) - 1.f;                  /* elu      */
    case 7: {                                                    /* selu     */
        const float s = 1.0507009873554804934193349852946f, a = 1.6732632423543772848170429916717f;
        return x > 0.f ? s * x : s * a * (expf(x) - 1.f);
    }
    case 8: return x > 20.f ? x : log1pf(expf(x));               /* softplus */
    case 9: return x / (1.f + expf(-x));                         /* swish    */
    default: return NAN;
    }
}

/* grad==0: forward.  grad==1: dx = dy * gain * act'(.) gated by the clamp, evaluated from the
 * saved forward input (xref, pre-bias) and/or output (yref) exactly like the native op does
 * (OPS/bias_act.cu:60-140): for lrelu/relu the derivative is taken from the sign of yref.      */
ORC_API int orc_bias_act(const float* x, const float* b, const float* xref, const float* yref,
                         float* y, int64_t n, int64_t stepB, int64_t sizeB,
                         int grad, int act, float alpha, float gain, float clamp)
{
    if (act < 1 || act > 9 || grad < 0 || grad > 1) return -1;
    for (int64_t i = 0; i < n; i++) {
        float bias = b ? b[(i / stepB) % sizeB] : 0.f;
        if (grad == 0) {
            float v = orc_act_eval(x[i] + bias, act, alpha) * gain;
            if (clamp >= 0.f) v = v > clamp ? clamp : (v < -clamp ? -clamp : v);
            y[i] = v;
        } else {
            float dy = x[i];
            float yy = yref ? yref[i] / gain : 0.f;        /* activation output before gain */
            float xx = xref ? xref[i] + bias : 0.f;
            float d;
            switch (act) {
            case 1: d = 1.f; break;
            case 2: d = yy > 0.f ? 1.f : 0.f; break;
            case 3: d = yy > 0.f ? 1.f : alpha; break;
            case 4: d = 1.f - yy * yy; break;
            case 5: d = yy * (1.f - yy); break;
            case 6: d = yy > 0.f ? 1.f : yy + 1.f; break;
            case 7: {
                const float s = 1.0507009873554804934193349852946f, a = 1.6732632423543772848170429916717f;
                d = yy > 0.f ? s : yy + s * a; break;
            }
            case 8: d = 1.f - expf(-yy); break;
            default: { float sg = 1.f / (1.f + expf(-xx)); d = sg * (1.f + xx * (1.f - sg)); } break;
            }
            float v = dy * gain * d;
            if (clamp >= 0.f && yref && fabsf(yref[i]) >= clamp) v = 0.f;
            y[i] = v;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* upfirdn2d on one plane.  f is a dense [fh][fw] filter ALREADY scaled by the gain.            */
/* Definition (OPS/upfirdn2d.py:186-210): zero-insert by (upx,upy), pad/crop, correlate with the */
/* flipped filter (i.e. true convolution) unless flip_filter, keep every (downx,downy)-th sample. */

static void orc_upfirdn2d_plane(const float* x, int H, int W, float* y, int OH, int OW,
                                const float* f, int fh, int fw,
                                int upx, int upy, int downx, int downy,
                                int px0, int py0, int flip)
{
    for (int oy = 0; oy < OH; oy++)
    for (int ox = 0; ox < OW; ox++) {
        float acc = 0.f;
        for (int ty = 0; ty < fh; ty++) {
            int zy = oy * downy + ty - py0;               /* index into the zero-inserted image */
            if (zy < 0 || zy % upy) continue;
            int iy = zy / upy;
            if (iy >= H) continue;
            for (int tx = 0; tx < fw; tx++) {
                int zx = ox * downx + tx - px0;
                if (zx < 0 || zx % upx) continue;
                int ix = zx / upx;
                if (ix >= W) continue;
                float tap = flip ? f[ty * fw + tx] : f[(fh - 1 - ty) * fw + (fw - 1 - tx)];
                acc += tap * x[iy * W + ix];
            }
        }
        y[oy * OW + ox] = acc;
    }
}

static int orc_out_size(int in, int up, int p0, int p1, int taps, int down)
{
    return (in * up + p0 + p1 - taps + down) / down;      /* OPS/upfirdn2d.cpp output size rule */
}

/* f: [fh][fw] (fh==0 means separable 1-D filter of fw taps applied along x then along y, each pass
 * scaled by sqrt(gain), as the reference does for f.ndim == 1: OPS/upfirdn2d.py:196,205-207). */
ORC_API int orc_upfirdn2d(const float* x, int64_t planes, int H, int W, float* y,
                          const float* f, int fh, int fw,
                          int upx, int upy, int downx, int downy,
                          int px0, int px1, int py0, int py1, int flip, float gain)
{
    int separable = (fh == 0);
    int fhh = separable ? fw : fh;
    int OW = orc_out_size(W, upx, px0, px1, fw, downx);
    int OH = orc_out_size(H, upy, py0, py1, fhh, downy);
    if (OW <= 0 || OH <= 0) return -1;
    float* fs = (float*)malloc(sizeof(float) * (size_t)(separable ? fw : fh * fw));
    float g = separable ? sqrtf(gain) : gain;
    for (int i = 0; i < (separable ? fw : fh * fw); i++) fs[i] = f[i] * g;
    int err = 0;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < planes; p++) {
        const float* xp = x + p * (int64_t)H * W;
        float* yp = y + p * (int64_t)OH * OW;
        if (!separable) {
            orc_upfirdn2d_plane(xp, H, W, yp, OH, OW, fs, fh, fw, upx, upy, downx, downy, px0, py0, flip);
        } else {
            /* horizontal pass keeps all rows, then vertical pass */
            float* tmp = (float*)malloc(sizeof(float) * (size_t)H * OW);
            if (!tmp) { err = 1; continue; }
            orc_upfirdn2d_plane(xp, H, W, tmp, H, OW, fs, 1, fw, upx, 1, downx, 1, px0, 0, flip);
            orc_upfirdn2d_plane(tmp, H, OW, yp, OH, OW, fs, fw, 1, 1, upy, 1, downy, 0, py0, flip);
            free(tmp);
        }
    }
    free(fs);
    return err ? -2 : 0;
}

/* ------------------------------------------------------------------------------------------ */
/* filtered_lrelu.  fu/fd are separable 1-D filters (fu_n/fd_n taps; n==1 with NULL pointer =    */
/* identity).  Sign tensor layout: uint8 [planes][sh][swb] with 4 up-res elements per byte,       */
/* element e of a row in bits (2*(e&3)) of byte e>>2; value 1 = negative, 2 = clamped.           */
/*   so != NULL : write signs (forward of a training step)                                       */
/*   si != NULL : read signs at offset (sx,sy) instead of evaluating lrelu/clamp (backward)       */
/* preact (optional, [planes][UH][UW]) receives the up-FIR output times gain, before lrelu.      */

ORC_API int orc_filtered_lrelu_sizes(int H, int W, int up, int down, int fu_n, int fd_n,
                                     int px0, int px1, int py0, int py1,
                                     int* UH, int* UW, int* OH, int* OW, int* SH, int* SWB)
{
    int uw = W * up + px0 + px1 - (fu_n - 1);
    int uh = H * up + py0 + py1 - (fu_n - 1);
    if (uw < fd_n || uh < fd_n) return -1;
    int ow = (uw - (fd_n - 1) + down - 1) / down;
    int oh = (uh - (fd_n - 1) + down - 1) / down;
    int sw_active = ow * down - (down - 1) + (fd_n - 1);
    int sh = oh * down - (down - 1) + (fd_n - 1);
    *UH = uh; *UW = uw; *OH = oh; *OW = ow; *SH = sh; *SWB = ((sw_active + 15) & ~15) >> 2;
    return 0;
}

ORC_API int orc_filtered_lrelu(const float* x, int64_t N, int64_t C, int H, int W,
                               const float* b, const float* fu, int fu_n, const float* fd, int fd_n,
                               int up, int down, int px0, int px1, int py0, int py1,
                               float gain, float slope, float clamp, int flip,
                               float* y, uint8_t* so, const uint8_t* si, int si_h, int si_wb,
                               int sx, int sy, float* preact)
{
    int UH, UW, OH, OW, SH, SWB;
    if (orc_filtered_lrelu_sizes(H, W, up, down, fu_n, fd_n, px0, px1, py0, py1, &UH, &UW, &OH, &OW, &SH, &SWB)) return -1;
    const float one = 1.f;
    const float* fup = fu ? fu : &one;
    const float* fdn = fd ? fd : &one;
    /* per-pass filter scaling: up-filter carries gain up^2 split as sqrt per 1-D pass */
    float* fus = (float*)malloc(sizeof(float) * (size_t)fu_n);
    for (int i = 0; i < fu_n; i++) fus[i] = fup[i] * (float)up;
    int err = 0;
    if (so) memset(so, 0, (size_t)(N * C) * SH * SWB);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < N * C; p++) {
        int c = (int)(p % C);
        float* xb = (float*)malloc(sizeof(float) * (size_t)H * W);
        float* t1 = (float*)malloc(sizeof(float) * (size_t)H * UW);
        float* u  = (float*)malloc(sizeof(float) * (size_t)UH * UW);
        float* t2 = (float*)malloc(sizeof(float) * (size_t)UH * OW);
        if (!xb || !t1 || !u || !t2) { err = 1; free(xb); free(t1); free(u); free(t2); continue; }
        const float* xp = x + p * (int64_t)H * W;
        float bias = b ? b[c] : 0.f;
        for (int i = 0; i < H * W; i++) xb[i] = xp[i] + bias;                 /* step 1: bias      */
        /* steps 2-4: zero-insert, pad, up-FIR (horizontal then vertical pass)                      */
        orc_upfirdn2d_plane(xb, H, W, t1, H, UW, fus, 1, fu_n, up, 1, 1, 1, px0, 0, flip);
        orc_upfirdn2d_plane(t1, H, UW, u, UH, UW, fus, fu_n, 1, 1, up, 1, 1, 0, py0, flip);
        /* steps 5-7: gain, leaky ReLU, clamp (+ sign bookkeeping)                                  */
        for (int uy = 0; uy < UH; uy++)
        for (int ux = 0; ux < UW; ux++) {
            float v = u[uy * UW + ux] * gain;
            if (preact) preact[(p * UH + uy) * (int64_t)UW + ux] = v;
            if (si) {
                int ex = ux + sx, ey = uy + sy;
                if (ex >= 0 && ex < si_wb * 4 && ey >= 0 && ey < si_h) {
                    int s = (si[(p * si_h + ey) * (int64_t)si_wb + (ex >> 2)] >> ((ex & 3) << 1)) & 3;
                    if (s & 1) v *= slope;
                    if (s & 2) v = 0.f;
                }
            } else {
                int s = 0;
                if (v < 0.f) { v *= slope; s = 1; }
                if (fabsf(v) > clamp) { v = v < 0.f ? -clamp : clamp; s = 2; }
                if (so && uy < SH && ux < SWB * 4)
                    so[(p * SH + uy) * (int64_t)SWB + (ux >> 2)] |= (uint8_t)(s << ((ux & 3) << 1));
            }
            u[uy * UW + ux] = v;
        }
        /* steps 8-9: down-FIR and decimation                                                      */
        orc_upfirdn2d_plane(u, UH, UW, t2, UH, OW, fdn, 1, fd_n, 1, 1, down, 1, 0, 0, flip);
        orc_upfirdn2d_plane(t2, UH, OW, y + p * (int64_t)OH * OW, OH, OW, fdn, fd_n, 1, 1, 1, 1, down, 0, 0, flip);
        free(xb); free(t1); free(u); free(t2);
    }
    free(fus);
    return err ? -2 : 0;
}

/* ------------------------------------------------------------------------------------------ */
/* conv2d (stride 1, symmetric zero padding, correlation), optional groups == N folding is done  */
/* by the caller.  x [N,Ci,H,W], w [Co,Ci,kh,kw], y [N,Co,H+2p-kh+1,W+2p-kw+1].                  */

ORC_API int orc_conv2d(const float* x, int64_t N, int Ci, int H, int W,
                       const float* w, int Co, int kh, int kw, int pad, float* y)
{
    int OH = H + 2 * pad - kh + 1, OW = W + 2 * pad - kw + 1;
    if (OH <= 0 || OW <= 0) return -1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t n = 0; n < N; n++)
    for (int o = 0; o < Co; o++) {
        float* yp = y + (n * Co + o) * (int64_t)OH * OW;
        for (int i = 0; i < OH * OW; i++) yp[i] = 0.f;
        for (int c = 0; c < Ci; c++) {
            const float* xp = x + (n * Ci + c) * (int64_t)H * W;
            const float* wp = w + ((int64_t)o * Ci + c) * kh * kw;
            for (int ky = 0; ky < kh; ky++)
            for (int kx = 0; kx < kw; kx++) {
                float tap = wp[ky * kw + kx];
                int y0 = pad - ky > 0 ? pad - ky : 0, y1 = H + pad - ky < OH ? H + pad - ky : OH;
                int x0 = pad - kx > 0 ? pad - kx : 0, x1 = W + pad - kx < OW ? W + pad - kx : OW;
                for (int oy = y0; oy < y1; oy++) {
                    const float* xr = xp + (oy + ky - pad) * W + (kx - pad);
                    float* yr = yp + oy * OW;
                    for (int ox = x0; ox < x1; ox++) yr[ox] += tap * xr[ox];
                }
            }
        }
    }
    return 0;
}

/* modulated_conv2d (networks_stylegan3.py:25-64).  input_gain: NULL, or [gain_n] with gain_n in
 * {1, Ci, N*Ci} (the reference expands it to [N,Ci]).                                            */
ORC_API int orc_modulated_conv2d(const float* x, int64_t N, int Ci, int H, int W,
                                 const float* w, int Co, int kh, int kw,
                                 const float* s, int demodulate, int pad,
                                 const float* input_gain, int64_t gain_n, float* y)
{
    int64_t wn = (int64_t)Ci * kh * kw;
    float* wq = (float*)malloc(sizeof(float) * (size_t)(Co * wn));
    float* sq = (float*)malloc(sizeof(float) * (size_t)(N * Ci));
    float* wm = (float*)malloc(sizeof(float) * (size_t)(Co * wn));
    if (!wq || !sq || !wm) { free(wq); free(sq); free(wm); return -2; }
    memcpy(wq, w, sizeof(float) * (size_t)(Co * wn));
    memcpy(sq, s, sizeof(float) * (size_t)(N * Ci));
    if (demodulate) {                                   /* :41-43 pre-normalisation */
        for (int o = 0; o < Co; o++) {
            double m = 0; for (int64_t i = 0; i < wn; i++) m += (double)wq[o * wn + i] * wq[o * wn + i];
            float r = 1.f / sqrtf((float)(m / (double)wn));
            for (int64_t i = 0; i < wn; i++) wq[o * wn + i] *= r;
        }
        double m = 0; for (int64_t i = 0; i < N * Ci; i++) m += (double)sq[i] * sq[i];
        float r = 1.f / sqrtf((float)(m / (double)(N * Ci)));
        for (int64_t i = 0; i < N * Ci; i++) sq[i] *= r;
    }
    int OH = H + 2 * pad - kh + 1, OW = W + 2 * pad - kw + 1;
    int rc = 0;
    for (int64_t n = 0; n < N && !rc; n++) {
        for (int o = 0; o < Co; o++) {                  /* :46-52 modulate, demodulate */
            double acc = 0;
            for (int c = 0; c < Ci; c++)
                for (int k = 0; k < kh * kw; k++) {
                    float v = wq[((int64_t)o * Ci + c) * kh * kw + k] * sq[n * Ci + c];
                    wm[((int64_t)o * Ci + c) * kh * kw + k] = v;
                    acc += (double)v * v;
                }
            if (demodulate) {
                float d = 1.f / sqrtf((float)acc + 1e-8f);
                for (int64_t i = 0; i < wn; i++) wm[o * wn + i] *= d;
            }
            if (input_gain)                             /* :55-57 */
                for (int c = 0; c < Ci; c++) {
                    float g = gain_n == 1 ? input_gain[0] : (gain_n == Ci ? input_gain[c] : input_gain[n * Ci + c]);
                    for (int k = 0; k < kh * kw; k++) wm[((int64_t)o * Ci + c) * kh * kw + k] *= g;
                }
        }
        rc = orc_conv2d(x + n * (int64_t)Ci * H * W, 1, Ci, H, W, wm, Co, kh, kw, pad,
                        y + n * (int64_t)Co * OH * OW);
    }
    free(wq); free(sq); free(wm);
    return rc;
}

/* FullyConnectedLayer.forward (networks_stylegan3.py:89-101): y = act((x @ (W*wg)^T) + b*bg).  */
ORC_API int orc_fully_connected(const float* x, int64_t N, int in_f, const float* w, int out_f,
                                const float* b, float weight_gain, float bias_gain,
                                int act, float alpha, float gain, float* y)
{
    if (act < 1 || act > 9) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; n++)
        for (int o = 0; o < out_f; o++) {
            float acc = 0.f;
            for (int i = 0; i < in_f; i++) acc += x[n * in_f + i] * (w[(int64_t)o * in_f + i] * weight_gain);
            if (b) acc += b[o] * bias_gain;
            y[n * out_f + o] = orc_act_eval(acc, act, alpha) * gain;
        }
    return 0;
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
