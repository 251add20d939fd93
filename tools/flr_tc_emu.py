"""Block-level numpy emulation of csrc/flr_tc.cu: the same per-warp state machine (input row-block ring, P slots, carry,
super-iterations with the pipeline fill / drain skips, the narrow last strip) executed on 16x16 / 16x8 matrix blocks that
are built from the same closed-form index expressions as load_consts().  It restates the kernel's control flow and
fragment-to-coordinate mapping on the CPU, so edits to the schedule (which blocks are computed, which are skipped, which
rows an emit owns) can be checked against the oracle without a GPU (tests/test_flr_tc_emu.py).  Numerics are float64
(optionally with the kernel's fp16 rounding points); it is a checker for index math, not a timing model.

Coordinates (strip-local, as in the kernel): X / Y input column / row, J / V up-sampled column / row, K / W output column /
row;  up_x[J] = sum_X kux[(X - dx) U - J] in[X],  up_y[V] = sum_Y kuy[U Y - V] in[Y],
out_x[K] = sum_J kdx[J - D K - sx] a[J],  out_y[W] = sum_V kdy[V - D W - sy] a[V]."""
import numpy as np


def floor_mod(a, m):
    return a % m


def _tap(k, e):
    return k[e] if 0 <= e < len(k) else 0.0


def r16(a, on):
    return np.asarray(a, np.float64).astype(np.float16).astype(np.float64) if on else np.asarray(a, np.float64)


class Geo:
    def __init__(self, U, D):
        self.U, self.D = U, D
        self.FU, self.FD = 6 * U, 6 * D
        self.KC4 = 3 if D == 2 else 6
        self.MB, self.JB = self.KC4, 2 * self.KC4
        self.NC = self.KC4 + 1 if U == 2 else (self.KC4 - 1) // 2 + 2
        self.NAL, self.NREL = D // 2, D
        self.IXS = 16 * D // U
        self.MBN = (D * 7 + (U - 1) + self.FD + 15) // 16


class Warp:
    """One strip of 16 output columns of one plane (one unit of the kernel)."""

    def __init__(self, geo, x, y, prm, fp16):
        self.g, self.x, self.y, self.p, self.fp16 = geo, x, y, prm, fp16
        self.pre_map = None                                   # optional [SH][SW] map of pre-activation values (see record)
        self.sign_words = None                                # optional [SH][SWB / 4] uint32 sign tensor of the plane
        self.codes = {}
        self.sign_in, self.sign_ofs = None, (0, 0)            # sign-read mode: uint8 [SH][SWB] of the plane, (s_ox, s_oy)
        G, p = geo, prm
        U, D = G.U, G.D
        # constant blocks (load_consts): A1[ph][J 16][X 16], B2[nb][Y 16][V 8], A3[al][W 16][V 16], B4[rel][J 16][K 8]
        self.A1 = [np.array([[_tap(p['kux'], (col - p['dx']) * U - 16 * ph - row) for col in range(16)] for row in range(16)])
                   for ph in range(1 if U == 2 else 2)]
        self.B2 = [np.array([[_tap(p['kuy'], U * k - 8 * nb - n) for n in range(8)] for k in range(16)]) for nb in range(U)]
        self.A3 = [np.array([[_tap(p['kdy'], 16 * al + col - D * row + 8 * D - p['sy']) for col in range(16)] for row in range(16)])
                   for al in range(G.NAL)]
        self.B4 = [np.array([[_tap(p['kdx'], 16 * rel + k - D * n - p['sx']) for n in range(8)] for k in range(16)])
                   for rel in range(G.NREL)]
        for lst in (self.A1, self.B2, self.A3, self.B4):
            for i in range(len(lst)):
                lst[i] = r16(lst[i], fp16)

    def begin_strip(self, unit):
        G, p = self.g, self.p
        seg, strip = divmod(unit, p['strips'])
        self.ix = strip * G.IXS + p['ix0']
        self.iy = seg * p['iy_step'] + p['iy0']
        self.k0, self.w0 = strip * 16, seg * p['seg_wblocks'] * 8
        self.nwb = min((p['yh'] - self.w0 + 7) >> 3, p['seg_wblocks'])
        self.eb = 0
        self.next_block = 0                                   # input row block the next convert() returns
        self.P = [[np.zeros((16, 8)) for _ in range(G.MB)] for _ in range(2)]
        self.carry = [np.zeros((8, 8)) for _ in range(G.JB)]  # [W 8][J 8] lower half of the last window
        self.narrow = p['yw'] - self.k0 <= 8
        self.mb_n = G.MBN if self.narrow else G.MB
        self.nb_n = 1 if self.narrow else 2

    def convert(self):
        """Next block of 8 input rows x 8 NC columns, zero outside the plane (bias on real samples only)."""
        G, p = self.g, self.p
        b = self.next_block
        self.next_block += 1
        blk = np.zeros((8, 8 * G.NC))
        xh, xw = self.x.shape
        for r in range(8):
            gy = self.iy + 8 * b + r
            if not (0 <= gy < xh):
                continue
            for c in range(8 * G.NC):
                gx = self.ix + c
                if 0 <= gx < xw:
                    blk[r, c] = self.x[gy, gx] + p['bias']
        return r16(blk, self.fp16)

    def step1(self, slot, blk):
        G = self.g
        for b in range(self.mb_n):
            w, ph = (b, 0) if G.U == 2 else (b >> 1, b & 1)
            # D[J 16][Y 8] = A1[ph][J][X 16] * in[Y][8 w + X]^T
            self.P[slot][b] = r16(self.A1[ph] @ blk[:, 8 * w:8 * w + 16].T, self.fp16)

    def record(self, mb, nb0, pre):
        """Where the sign tensor would get its codes from: fragment element (J = 16 mb + m, V = 8 (nb0 + q) + n of the window
        whose current input row block is blk) is up-sampled sample (row, column) of the plane in sign-tensor coordinates
            uy = D w0 + 8 U (blk - 1) + 8 (nb0 + q) + n - sy,      ux = D k0 + 16 mb + m - sx."""
        G, p = self.g, self.p
        blk = self.next_block - 1
        H, W = self.pre_map.shape
        for q in range(2):
            for m in range(16):
                ux = G.D * self.k0 + 16 * mb + m - p['sx']
                for n in range(8):
                    uy = G.D * self.w0 + 8 * G.U * (blk - 1) + 8 * (nb0 + q) + n - p['sy']
                    if 0 <= uy < H and 0 <= ux < W:
                        self.pre_map[uy, ux] = pre[q][m, n]

    def act(self, v):
        p = self.p
        v = np.where(v < 0, v * p['slope'], v)
        return np.clip(v, -p['act_clamp'], p['act_clamp'])

    def sign_mult(self, mb, nb0, q):
        """Sign-read mode (backward): multiplier of every fragment element of column block mb, row block nb0 + q -- slope where
        the stored code is 1, 0 where it is 2, 1 elsewhere and outside the tensor (OPS/filtered_lrelu.cu:562-572)."""
        G, p = self.g, self.p
        si, (s_ox, s_oy) = self.sign_in, self.sign_ofs
        blk = self.next_block - 1
        m = np.ones((16, 8))
        for j in range(16):
            ex = G.D * self.k0 + 16 * mb + j - p['sx'] + s_ox
            for n in range(8):
                ey = G.D * self.w0 + 8 * G.U * (blk - 1) + 8 * (nb0 + q) + n - p['sy'] + s_oy
                if 0 <= ey < si.shape[0] and 0 <= ex < 4 * si.shape[1]:
                    code = (int(si[ey, ex >> 2]) >> (2 * (ex & 3))) & 3
                    m[j, n] = p['slope'] if code == 1 else (0.0 if code >= 2 else 1.0)
        return m

    def code_of(self, pre):
        """2-bit code of a pre-activation value (in units of u_scale): 1 = negative, 2 = clamped (overrides)."""
        p = self.p
        a = np.where(pre < 0, pre * p['slope'], pre)
        return np.where(np.abs(a) > p['act_clamp'], 2, (pre < 0).astype(np.int64)).astype(np.uint64)

    def flush_signs(self, nb0):
        """Sign write of one chunk (16 up-sampled rows x 16 mb_n columns): the warp recipe of tools/flr_sign_pack_model.py,
        then the ownership rule -- a strip stores the D aligned words of its own 16 D up-sampled columns (the last strip
        every word up to the end of the row), a segment the rows of its own output rows (the last one the rest)."""
        from tools.flr_sign_pack_model import pack_chunk
        G, p = self.g, self.p
        SH, NW = self.sign_words.shape
        blk = self.next_block - 1
        strip = self.k0 // 16
        last_strip = strip == p['strips'] - 1
        n_own = (NW - G.D * strip) if last_strip else G.D
        code = np.concatenate([self.codes[mb] for mb in range(self.mb_n)] + [np.zeros((16, 16), np.uint64)] * 2, axis=0)
        words = pack_chunk(code, p['sx'])                                   # [16 rows][mb_n + 1]
        assert n_own <= words.shape[1]
        row_lo = G.D * self.w0
        row_hi = SH if self.w0 + 8 * p['seg_wblocks'] >= p['yh'] else G.D * (self.w0 + 8 * p['seg_wblocks'])
        for n in range(16):
            uy = G.D * self.w0 + 8 * G.U * (blk - 1) + 8 * nb0 + n - p['sy']
            if row_lo <= uy < row_hi and uy >= 0:
                for wl in range(n_own):
                    self.sign_words[uy, G.D * strip + wl] = words[n, wl]

    def chunk(self, cur, nb0, mode, al, win, X):
        """Vertical up-FIR of the window (previous | current) for row blocks nb0, nb0 + 1, activation, and that chunk's
        contribution to the window of R3.  win / X: lists over jb of [16 W][8 J] / [8 W][8 J]."""
        for mb in range(self.mb_n):
            prev, curb = self.P[1 - cur][mb], self.P[cur][mb]
            quad = np.concatenate([prev, curb], axis=1)                     # [J 16][Y 16], natural row order
            pre = [quad @ self.B2[nb0 + q] for q in range(2)]               # [J 16][V 8] each, in units of u_scale
            if self.pre_map is not None:
                self.record(mb, nb0, pre)
            if self.sign_words is not None:
                self.codes[mb] = np.concatenate([self.code_of(pre[0]), self.code_of(pre[1])], axis=1)     # [J 16][V 16]
            if self.sign_in is not None:
                e = [r16(v * self.sign_mult(mb, nb0, q), self.fp16) for q, v in enumerate(pre)]
            else:
                e = [r16(self.act(v), self.fp16) for v in pre]
            for h in range(2):
                jb = 2 * mb + h
                Bv = np.concatenate([e[0][8 * h:8 * h + 8, :].T, e[1][8 * h:8 * h + 8, :].T], axis=0)    # [V 16][J 8]
                contrib = self.A3[al] @ Bv                                  # [W 16][J 8]: rows 0-7 = window rows -8..-1
                if mode == 0:
                    w = r16(contrib, self.fp16)
                    X[jb] = r16(self.carry[jb] + w[:8], self.fp16)
                    self.carry[jb] = w[8:]
                elif mode == 1:
                    win[jb] = r16(contrib, self.fp16)
                else:
                    win[jb] = r16(win[jb] + contrib, self.fp16)
                    X[jb] = r16(self.carry[jb] + win[jb][:8], self.fp16)
                    self.carry[jb] = win[jb][8:]
        if self.sign_words is not None:
            self.flush_signs(nb0)

    def emit(self, X0, X1, rows, adv):
        """Horizontal down-FIR of two blocks of 8 output rows and the masked store; advances the output cursor."""
        G, p = self.g, self.p
        for nb in range(self.nb_n):
            c = np.zeros((16, 8))
            for rel in range(G.NREL):
                kc = nb * (G.D // 2) + rel
                A = np.block([[X0[2 * kc], X0[2 * kc + 1]], [X1[2 * kc], X1[2 * kc + 1]]])      # [W 16][J 16]
                c += A @ self.B4[rel]
            for half, ok in ((0, rows & 1), (1, rows & 2)):
                e = self.eb + half
                if not ok or not (0 <= e < self.nwb):
                    continue
                for r in range(8):
                    gy = self.w0 + 8 * e + r
                    if gy >= p['yh']:
                        continue
                    for n in range(8):
                        gx = self.k0 + 8 * nb + n
                        if gx < p['yw']:
                            self.y[gy, gx] = c[8 * half + r, n]
        self.eb += adv

    # ---- U == 2, D == 2 -------------------------------------------------------------------------------------------
    def run22(self):
        G = self.g
        zero8 = lambda: [np.zeros((8, 8)) for _ in range(G.JB)]
        X0, X1 = zero8(), zero8()
        S = ((self.nwb - 1) >> 1) + 2
        self.eb = -2
        for s in range(S):
            blk = self.convert(); self.step1(0, blk)
            if s > 0:
                self.chunk(0, 0, 0, 0, None, X0)
            if self.eb + 1 < self.nwb:
                blk = self.convert(); self.step1(1, blk)
                self.chunk(1, 0, 0, 0, None, X1)
            if s > 0:
                self.emit(X0, X1, 3, 2)
            else:
                self.eb += 2

    # ---- U == 4, D == 2 -------------------------------------------------------------------------------------------
    def run42(self):
        G = self.g
        X0 = [np.zeros((8, 8)) for _ in range(G.JB)]
        iters = ((self.nwb - 1) >> 1) + 3
        S = (iters + 1) >> 1
        self.eb = -4
        for it in range(2 * S):
            cur = it & 1
            if self.eb >= self.nwb:
                self.eb += 2
                continue
            blk = self.convert(); self.step1(cur, blk)
            X1 = [np.zeros((8, 8)) for _ in range(G.JB)]
            if it >= 1 and self.eb + 1 < self.nwb:
                self.chunk(cur, 0, 0, 0, None, X1)
            if it >= 2:
                self.emit(X0, X1, 3, 2)
            else:
                self.eb += 2
            if it >= 1 and self.eb < self.nwb:
                self.chunk(cur, 2, 0, 0, None, X0)

    # ---- U == 2, D == 4 -------------------------------------------------------------------------------------------
    def run24(self):
        G = self.g
        win = [np.zeros((16, 8)) for _ in range(G.JB)]
        Xp = [np.zeros((8, 8)) for _ in range(G.JB)]
        iters = self.nwb + 2
        self.eb = -3
        for it in range(iters):
            blk = self.convert(); self.step1(0, blk)
            X = [np.zeros((8, 8)) for _ in range(G.JB)]
            if it >= 1:
                self.chunk(0, 0, 2, 1, win, X)
            if it >= 2:
                self.emit(Xp, X, 2, 1)
            else:
                self.eb += 1
            Xp = X
            if it - 1 < self.nwb:
                blk = self.convert(); self.step1(1, blk)
                self.chunk(1, 0, 1, 0, win, X)

    def run(self):
        {(2, 2): self.run22, (4, 2): self.run42, (2, 4): self.run24}[(self.g.U, self.g.D)]()


def filtered_lrelu_tc_emu(x, fu, fd, b, up, down, padding, gain, slope, clamp, flip_filter=False, fp16=False, seg_wblocks=None,
                          preact_shape=None, sign_shape=None, si=None, s_ofs=(0, 0)):
    """x: [N, C, H, W] -> y like afcm_filtered_lrelu_tc (host parameter set-up of the C entry point + launch_tc).
    preact_shape = (SH, SW): also return the pre-activation values (after the gain, before slope / clamp) the warps hold,
    placed at their sign-tensor coordinates -> (y, pre [N, C, SH, SW]); NaN where no warp computed the sample.
    sign_shape = (SH, SWB): also emulate the planned sign-write mode -> (y, signs uint8 [N, C, SH, SWB]).
    si [N, C, SH, SWB], s_ofs = (sx, sy): the planned sign-read mode (the backward pass of the op)."""
    x = np.asarray(x, np.float64)
    N, C, xh, xw = x.shape
    px0, px1, py0, py1 = padding
    G = Geo(up, down)
    assert len(fu) == G.FU and len(fd) == G.FD and xw % 2 == 0
    yw = (xw * up + px0 + px1 - (G.FU - 1) - (G.FD - 1) + down - 1) // down
    yh = (xh * up + py0 + py1 - (G.FU - 1) - (G.FD - 1) + down - 1) // down
    p = dict(yh=yh, yw=yw, slope=slope)
    p['sx'], p['sy'] = floor_mod(-px0, up), floor_mod(-py0, up)
    bx = (-p['sx'] - px0) // up
    p['dx'] = 1 if (bx & 1) else 0
    p['ix0'] = bx - p['dx']
    p['iy0'] = (-p['sy'] - py0) // up
    finite = clamp is not None and np.isfinite(clamp)
    u_scale = 1.0 / clamp if (finite and 1.0 / 1024 <= clamp <= 1024) else 1.0
    p['act_clamp'] = (clamp * u_scale) if finite else np.inf
    fuc = np.asarray(fu, np.float64)[::1 if flip_filter else -1] * up
    fdc = np.asarray(fd, np.float64)[::1 if flip_filter else -1]
    p['kux'], p['kuy'], p['kdx'], p['kdy'] = fuc, fuc * gain * u_scale, fdc, fdc / u_scale
    p['strips'] = (yw + 15) // 16
    wblocks = (yh + 7) // 8
    wblocks += wblocks & 1
    p['seg_wblocks'] = seg_wblocks or wblocks
    segs = -(-wblocks // p['seg_wblocks'])
    p['iy_step'] = p['seg_wblocks'] * 8 * down // up
    y = np.zeros((N, C, yh, yw))
    pre = np.full((N, C) + tuple(preact_shape), np.nan) if preact_shape else None
    sgn = np.zeros((N, C, sign_shape[0], sign_shape[1] // 4), np.uint32) if sign_shape else None
    for n in range(N):
        for c in range(C):
            p['bias'] = 0.0 if b is None else float(b[c])
            w = Warp(G, x[n, c], y[n, c], p, fp16)
            if pre is not None:
                w.pre_map = pre[n, c]
            if sgn is not None:
                w.sign_words = sgn[n, c]
            if si is not None:
                w.sign_in, w.sign_ofs = si[n, c], s_ofs
            for unit in range(p['strips'] * segs):
                w.begin_strip(unit)
                w.run()
    if sgn is not None:
        return y, sgn.astype('<u4').view(np.uint8).reshape(N, C, sign_shape[0], sign_shape[1])
    return y if pre is None else (y, pre / u_scale)
