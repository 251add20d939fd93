"""numpy model of the tensor-core filtered_lrelu strip algorithm (afcm_b200/csrc/flr_tc.cu): the same
closed-form local operators (phase shift s, input alignment delta, tile origins), evaluated densely per
16-column strip, optionally with fp16 rounding at the points where the kernel rounds.  Used by
tests/test_flr_tc_model.py to pin the index math against the oracle on the CPU."""
import numpy as np


def floor_mod(a, m):
    return a % m          # python % is a floor mod


def axis_consts(U, D, FU, FD, p0, want_even_origin):
    """-> s (up-res samples in front), delta (extra input samples in front), i0 (input origin of tile 0),
    i_step16 (input origin advance per 16 outputs)."""
    s = floor_mod(-p0, U)
    assert (-s - p0) % U == 0
    base = (-s - p0) // U
    delta = 0
    if want_even_origin and (base % 2) != 0:
        delta = 1
    return s, delta, base - delta, 16 * D // U


def taps_corr(f, n, flip, scale):
    if f is None:
        return np.ones(1, np.float64) * scale
    f = np.asarray(f, np.float64)
    return (f if flip else f[::-1]) * scale


def local_ops(U, D, ku, kd, s, delta, n_out, halo_rows=None):
    """Dense local matrices for n_out outputs starting at a tile origin: M3 [n_out, J], M1 [J, X]."""
    FU, FD = len(ku), len(kd)
    J = (n_out - 1) * D + FD + s
    J = (J + 15) // 16 * 16
    X = (J + U - 1) // U + FU // U + 2 + delta
    M1 = np.zeros((J, X)); M3 = np.zeros((n_out, J))
    for jl in range(J):
        for xl in range(X):
            t = (xl - delta) * U - jl
            if 0 <= t < FU:
                M1[jl, xl] = ku[t]
    for kl in range(n_out):
        for jl in range(J):
            t = jl - kl * D - s
            if 0 <= t < FD:
                M3[kl, jl] = kd[t]
    return M1, M3


def r16(a, on):
    return a.astype(np.float16).astype(np.float64) if on else a


def filtered_lrelu_model(x, fu, fd, b, up, down, padding, gain, slope, clamp, flip_filter=False, fp16=False,
                         strip=16, seg=16):
    x = np.asarray(x, np.float64)
    N, C, xh, xw = x.shape
    px0, px1, py0, py1 = padding
    FU = 1 if fu is None else len(fu); FD = 1 if fd is None else len(fd)
    yw = (xw * up + px0 + px1 - (FU - 1) - (FD - 1) + down - 1) // down
    yh = (xh * up + py0 + py1 - (FU - 1) - (FD - 1) + down - 1) // down
    kux = taps_corr(fu, FU, flip_filter, up); kuy = taps_corr(fu, FU, flip_filter, up * gain)
    kdx = taps_corr(fd, FD, flip_filter, 1.0); kdy = kdx
    sx, dx, ix0, ixs = axis_consts(up, down, FU, FD, px0, True)
    sy, dy, iy0, iys = axis_consts(up, down, FU, FD, py0, False)
    M1x, M3x = local_ops(up, down, r16(kux, fp16), r16(kdx, fp16), sx, dx, strip)
    M1y, M3y = local_ops(up, down, r16(kuy, fp16), r16(kdy, fp16), sy, dy, seg)
    y = np.zeros((N, C, yh, yw))
    bias = np.zeros(C) if b is None else np.asarray(b, np.float64)
    for si in range((yw + strip - 1) // strip):
        for sj in range((yh + seg - 1) // seg):
            cx = ix0 + si * ixs * strip // 16; cy = iy0 + sj * iys * seg // 16
            X = M1x.shape[1]; Y = M1y.shape[1]
            win = np.zeros((N, C, Y, X))
            for yl in range(Y):
                gy = cy + yl
                if gy < 0 or gy >= xh:
                    continue
                for xl in range(X):
                    gx = cx + xl
                    if 0 <= gx < xw:
                        win[:, :, yl, xl] = x[:, :, gy, gx] + bias[None, :]
            win = r16(win, fp16)
            r1 = r16(np.einsum('jx,ncyx->ncjy', M1x, win), fp16)         # [J, Y]
            r2 = np.einsum('vy,ncjy->ncvj', M1y, r1)                      # [V, J]
            r2 = np.where(r2 < 0, r2 * slope, r2)
            if clamp is not None:
                r2 = np.clip(r2, -clamp, clamp)
            r2 = r16(r2, fp16)
            r3 = r16(np.einsum('kj,ncvj->nckv', M3x, r2), fp16)           # [K, V]
            r4 = np.einsum('wv,nckv->ncwk', M3y, r3)                      # [W, K]
            h = min(seg, yh - sj * seg); w = min(strip, yw - si * strip)
            y[:, :, sj * seg:sj * seg + h, si * strip:si * strip + w] = r4[:, :, :h, :w]
    return y
