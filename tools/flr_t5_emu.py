"""numpy emulation of the tcgen05 / TMEM filtered_lrelu (afcm_b200/csrc/flr_t5.cu): the same plan (origins, windows, Toeplitz
tiles, step / group structure, shared-memory ring of the down-sampled rows) evaluated with dense matrix products, optionally
with the kernel's roundings (fp16 operands, tf32 truncation of the first accumulator).  tests/test_flr_t5_emu.py pins it
against the oracle on the CPU, with the plan taken from the library's own host code (afcm_filtered_lrelu_t5_plan), so the
index algebra of the kernel is checked before it reaches a GPU.

Data flow of one (plane, strip), streamed top to bottom in steps of 128 up-sampled rows (= the 128 TMEM lanes):
  P1  D1[v, x]   = sum_i  Tuy[v, i]  X[i, x]          Toeplitz = A (TMEM), input rows = B (shared memory, from TMA)
  P2  D2[v, j]   = sum_x  D1[v, x]  T2[x, j]          D1 read in place as a tf32 A operand; group of 64 up-sampled columns
      A3[v, j]   = act(D2[v, j])                      epilogue warps: fp32 -> packed half2, in place
  P3  D3[v, k]  += sum_j  A3[v, j]  T3[j, k]          accumulates over the groups of the strip
      ring[v, k] = fp16(D3[v, k])                     epilogue warps -> shared memory (MN-major A operand of P4)
  P4  D4[k, w]   = sum_v  ring[v, k]  T4[v, w]        lanes = output columns, TMEM columns = output rows
"""
import numpy as np


def cdiv(a, b):
    return -((-a) // b)


def t5_plan_py(xh, xw, up, down, padding, kw=None):
    """Python statement of the plan (the C++ host code computes the same numbers; the test compares them)."""
    U, D = up, down
    FU, FD = 6 * U, 6 * D
    px0, px1, py0, py1 = padding
    yw = (xw * U + px0 + px1 - (FU - 1) - (FD - 1) + D - 1) // D
    yh = (xh * U + py0 + py1 - (FU - 1) - (FD - 1) + D - 1) // D
    p = dict(U=U, D=D, FU=FU, FD=FD, xh=xh, xw=xw, yh=yh, yw=yw, px0=px0, py0=py0)
    # ---- rows (TMEM lanes = up-sampled rows 128 s .. 128 s + 127)
    p['RS'] = 128 // U                       # new input rows per step
    p['K1'] = p['RS'] + 16                   # input rows of one step's window
    p['I0y'] = cdiv(-py0, U)                 # input row of window row 0 of step 0
    p['tuy_e'] = U * p['I0y'] + py0          # Tuy[m][k] = kuy[tuy_e + U k - m]
    p['OS'] = 128 // D                       # output rows per step
    p['wlo0'] = cdiv(-(FD - 1), D)           # first output row of step 0 (negative: a few dead rows)
    p['nsteps'] = cdiv(yh - p['wlo0'], p['OS'])
    p['adv4'] = 16 // D                      # D4 column advance per 16-row chunk
    p['NL'] = cdiv(FD - 1, 16)               # lead chunks (rows of the previous step)
    p['t4_e'] = -D * p['wlo0']               # T4reg[k][n] = kdy[k - D n + t4_e];  lead e: kdy[k - D n + t4_e - 16 e]
    # ---- columns (TMEM columns)
    kwq = {(2, 2): 8, (4, 2): 16, (2, 4): 4}[(U, D)]
    if kw is None:
        # widest strip that fits: D1 <= 128 columns, D3 columns of valid outputs < 128
        nstrips = 1
        while True:
            kw = cdiv(cdiv(yw, nstrips), kwq) * kwq
            q = _xplan(p, kw)
            if q['N1'] <= 128 and q['m0'] + kw <= 128:
                break
            nstrips += 1
    p.update(_xplan(p, kw))
    assert p['N1'] <= 128 and p['m0'] + kw <= 128
    p['KW'] = kw
    p['nstrips'] = cdiv(yw, kw)
    return p


def _xplan(p, kw):
    U, D, FU, FD, px0 = p['U'], p['D'], p['FU'], p['FD'], p['px0']
    ineed0 = cdiv(-px0, U)                                   # strip 0: first input column the first output needs
    iorg0 = (ineed0 // 8) * 8                                # 8-aligned (TMA: 16-byte aligned box start)
    istep = D * kw // U
    assert (D * kw) % U == 0 and istep % 8 == 0
    jorg0 = U * iorg0 + px0 - FU + 1                         # up-sampled column of D2 column 0
    korg0 = cdiv(jorg0 - FD + 1, D)                          # output column of D3 column 0
    m0 = -korg0                                              # D3 column of the strip's first output
    t3_e = jorg0 - D * korg0                                 # T3[k][n] = kdx[k - D n + t3_e]
    jlast = D * (kw - 1) + FD - 1
    nr = cdiv(jlast - jorg0 + 1, 16)                         # P3 chunks of 16 up-sampled columns
    ng = cdiv(nr, 4)                                         # groups of 64 up-sampled columns
    n1 = 64 * ng // U                                        # D1 columns (= 8 per P2 chunk)
    return dict(iorg0=iorg0, istep=istep, jorg0=jorg0, korg0=korg0, m0=m0, t3_e=t3_e, NG=ng, N1=n1, adv3=16 // D)


def r16(a):
    return a.astype(np.float16).astype(np.float64)


def tf32_trunc(a):
    b = a.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return b.view(np.float32).astype(np.float64)


def tf32_round(a):
    b = a.astype(np.float32).view(np.uint32)
    b = (b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return b.view(np.float32).astype(np.float64)


def tap(k, e):
    e = np.asarray(e)
    ok = (e >= 0) & (e < len(k))
    return np.where(ok, k[np.clip(e, 0, len(k) - 1)], 0.0)


def build_tiles(p, fu, fd, gain, slope, clamp, out_scale, flip_filter, rounded, tf32_comp=0.0):
    """The constant operands.  Scaling as in flr_tc: the activation runs in units of `clamp`."""
    U, D, FU, FD = p['U'], p['D'], p['FU'], p['FD']
    fu = np.asarray(fu, np.float64); fd = np.asarray(fd, np.float64)
    fuc = fu if flip_filter else fu[::-1]
    fdc = fd if flip_filter else fd[::-1]
    u_scale = 1.0 / clamp
    kux = fuc * U * (1.0 + tf32_comp)
    kuy = fuc * U * gain * u_scale
    kdx = fdc * out_scale
    kdy = fdc / u_scale
    m = np.arange(128)[:, None]
    k = np.arange(p['K1'])[None, :]
    t = {}
    t['Tuy'] = tap(kuy, p['tuy_e'] + U * k - m)                                              # [128, K1]
    kk = np.arange(8)[:, None]; n = np.arange(16 * U)[None, :]
    t['T2'] = tap(kux, U * kk - n + FU - 1)                                                  # [8, 16U]
    kk = np.arange(16)[:, None]; n = np.arange(16)[None, :]
    t['T3'] = tap(kdx, kk - D * n + p['t3_e'])                                               # [16, 16]
    t['T4'] = [tap(kdy, kk - D * n + p['t4_e'] - 16 * e) for e in range(p['NL'] + 1)]        # e = 0: regular
    if rounded:
        t['Tuy'] = r16(t['Tuy']); t['T3'] = r16(t['T3']); t['T4'] = [r16(a) for a in t['T4']]
        t['T2'] = tf32_round(t['T2'])
    return t


def filtered_lrelu_t5_emu(x, fu, fd, up, down, padding, gain, slope, clamp, out_scale=1.0, flip_filter=False, rounded=False,
                          plan=None, kw=None, tf32_comp=0.0):
    x = np.asarray(x, np.float64)
    N, C, xh, xw = x.shape
    p = plan or t5_plan_py(xh, xw, up, down, padding, kw=kw)
    U, D = p['U'], p['D']
    t = build_tiles(p, fu, fd, gain, slope, clamp, out_scale, flip_filter, rounded, tf32_comp)
    yh, yw, KW = p['yh'], p['yw'], p['KW']
    y = np.zeros((N, C, yh, yw))
    xin = r16(x) if rounded else x
    NG, N1, adv3, adv4, OS = p['NG'], p['N1'], p['adv3'], p['adv4'], p['OS']
    for nn in range(N):
        for c in range(C):
            for st in range(p['nstrips']):
                iorg = p['iorg0'] + p['istep'] * st
                korg = p['korg0'] + KW * st
                ring = np.zeros((256, 128))
                for s in range(p['nsteps']):
                    # ---- TMA: K1 input rows x N1 columns, zero outside the plane
                    X = np.zeros((p['K1'], N1))
                    r0 = p['I0y'] + p['RS'] * s
                    for rr in range(p['K1']):
                        gy = r0 + rr
                        if 0 <= gy < xh:
                            lo = max(0, -iorg); hi = min(N1, xw - iorg)
                            if hi > lo:
                                X[rr, lo:hi] = xin[nn, c, gy, iorg + lo:iorg + hi]
                    # ---- P1
                    D1 = t['Tuy'] @ X
                    if rounded:
                        D1 = tf32_trunc(D1)
                    # ---- groups: P2, activation, P3
                    D3 = np.full((128, adv3 * (4 * NG - 1) + 16), np.nan)
                    D3[:, :16] = 0.0                                   # the zero-tile MMA of group 0
                    for g in range(NG):
                        D2 = np.full((128, 64), np.nan)
                        qpg = 8 // U                                   # P2 chunks per group (advance 8U columns each)
                        half = 8 * U
                        # chunk j of the group (j = -1 .. qpg-1) has the window [half j, half j + 2 half) in group columns.
                        # Init set (overwrite): lead j = -1 (upper half; group 0: zero tile), odd full chunks, tail j = qpg-1 (lower half)
                        def contrib(q, lo, hi):
                            return D1[:, 8 * q:8 * q + 8] @ t['T2'][:, lo:hi]
                        q0 = qpg * g
                        D2[:, 0:half] = contrib(q0 - 1, half, 2 * half) if g > 0 else 0.0
                        for j in range(1, qpg - 1, 2):
                            D2[:, half * j:half * j + 2 * half] = contrib(q0 + j, 0, 2 * half)
                        D2[:, 64 - half:64] = contrib(q0 + qpg - 1, 0, half)
                        for j in range(0, qpg - 1, 2):                 # even full chunks: accumulate
                            D2[:, half * j:half * j + 2 * half] += contrib(q0 + j, 0, 2 * half)
                        assert not np.isnan(D2).any()
                        if rounded:
                            D2 = r16(D2)
                        A3 = np.clip(D2, 0, 1) - np.clip(-slope * D2, 0, 1)
                        if rounded:
                            A3 = r16(A3)
                        # P3: first-touch chunks overwrite, the others accumulate
                        order = [(1, True), (0, False), (3, True), (2, False)] if D == 2 else [(3, True), (0, False), (1, False), (2, False)]
                        for r, ow in order:
                            R = 4 * g + r
                            v = A3[:, 16 * r:16 * r + 16] @ t['T3']
                            if ow:
                                D3[:, adv3 * R:adv3 * R + 16] = v
                            else:
                                D3[:, adv3 * R:adv3 * R + 16] += v
                    # ---- E2: D3 columns 0..127 -> ring rows of this step (fp16)
                    rows = (128 * s + np.arange(128)) % 256
                    d3 = D3[:, :128] if D3.shape[1] >= 128 else np.pad(D3, ((0, 0), (0, 128 - D3.shape[1])))
                    d3 = np.nan_to_num(d3, nan=777.0)                 # dead columns: any finite garbage
                    ring[rows, :] = r16(d3) if rounded else d3
                    # ---- P4: lanes = D3 columns (output columns), D4 columns = output rows of this step
                    D4 = np.full((128, adv4 * 7 + 16), np.nan)
                    for i in range(0, 8, 16 // adv4):                  # init set: windows that tile the columns
                        rws = (128 * s + 16 * i + np.arange(16)) % 256
                        D4[:, adv4 * i:adv4 * i + 16] = ring[rws, :].T @ t['T4'][0]
                    if s > 0:
                        for e in range(1, p['NL'] + 1):
                            rws = (128 * s - 16 * e + np.arange(16)) % 256
                            D4[:, 0:16] += ring[rws, :].T @ t['T4'][e]
                    for i in range(8):
                        if i % (16 // adv4) == 0:
                            continue
                        rws = (128 * s + 16 * i + np.arange(16)) % 256
                        D4[:, adv4 * i:adv4 * i + 16] += ring[rws, :].T @ t['T4'][0]
                    # ---- E3: valid outputs
                    for n in range(OS):
                        w = OS * s + p['wlo0'] + n
                        if w < 0 or w >= yh:
                            continue
                        for m in range(p['m0'], p['m0'] + KW):
                            k = korg + m
                            if 0 <= k < yw:
                                y[nn, c, w, k] = D4[m, n]
    return y
