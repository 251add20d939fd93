"""numpy model + cost model of the PLANNED tcgen05/TMEM filtered_lrelu -- design
groundwork for the next round, not a kernel.  It pins, on the CPU and against the oracle:

  * the tile geometry: for an output tile of TH x TW pixels, the up-sampled window (V x J) and the input window
    (Y x X) every pass needs, with the same phase-shift / origin formulas as csrc/flr_tc.cu (tools/flr_tc_model.py);
  * the pass order that keeps every operand MN-major without a transpose: vertical up (contract y, lines = x) ->
    horizontal up (contract x, lines = v) -> activation -> vertical down (contract v, lines = j) -> horizontal down
    (contract j, lines = w): each pass leaves its result with the NEXT contraction index on the TMEM lanes, so an
    epilogue thread writes its row of the accumulator as one shared-memory row of the next A operand;
  * the blocking of every pass into `tcgen05.mma` instructions D[128 lines, NB] += A[128 lines, 16] * T[16, NB]:
    for each block of NB output positions only the K slices of 16 input positions that meet the Toeplitz band are
    issued (the model computes exactly those products and nothing else, so a wrong band table shows as an error);
  * where the values are rounded to fp16 (every operand tile; accumulation is fp32 in tensor memory);
  * a cost per tile from the rates measured this round: tensor time 128*NB/256 clk per K slice, shared-memory time
    (4 KB of A + 32*NB bytes of T per instruction at 128 B/clk) -- an instruction costs the larger of the two (the
    64-channel convolution measured exactly this: profiles/r01_ncu_conv2d_tc_enc1_64ch.txt) -- and the epilogue work
    (TMEM -> fp16 -> shared memory, ~0.75 thread-instructions per value, 1.5 with the activation).

tests/test_flr_t5_model.py checks the numerics against the oracle and the cost summary against the numbers quoted in
DESIGN.md."""
import numpy as np

from tools.flr_tc_model import axis_consts, taps_corr, r16


class Axis:
    """One image axis of one tile: n_out outputs starting at output index o0.  Local index spaces: input i in [0, NI),
    up-sampled u in [0, NU), output o in [0, n_out).  up[u] = sum_i ku[(i - delta) U - u] in[i];
    out[o] = sum_u kd[u - D o - s] act(up[u])."""

    def __init__(self, U, D, ku, kd, p0, n_out, even_origin):
        self.U, self.D, self.n_out = U, D, n_out
        FU, FD = len(ku), len(kd)
        self.s, self.delta, self.i0, step16 = axis_consts(U, D, FU, FD, p0, even_origin)
        self.i_step = step16 * n_out // 16                      # input origin advance per tile
        assert (step16 * n_out) % 16 == 0
        self.NU = (n_out - 1) * D + FD + self.s                  # up-sampled samples the outputs read
        self.NI = (self.NU - 1) // U + FU // U + 1 + self.delta  # input samples those read
        self.Tu = np.zeros((self.NI, self.NU))                   # [contracted input, up-sampled]
        for i in range(self.NI):
            for u in range(self.NU):
                t = (i - self.delta) * U - u
                if 0 <= t < FU:
                    self.Tu[i, u] = ku[t]
        self.Td = np.zeros((self.NU, n_out))                     # [contracted up-sampled, output]
        for u in range(self.NU):
            for o in range(n_out):
                t = u - o * D - self.s
                if 0 <= t < FD:
                    self.Td[u, o] = kd[t]


def blocked_product(A, T, NB, stats, name, fp16):
    """D[lines, n] = sum_k A[lines, k] T[k, n] issued as tcgen05-sized instructions: lines in groups of 128 (M), n in
    blocks of NB, k in slices of 16; a (n-block, k-slice) pair is issued only if the Toeplitz block is non-zero.
    A: [..., lines, K] (operand tile, already fp16-rounded), T: [K, N]."""
    L, K = A.shape[-2:]
    N = T.shape[1]
    Kp = (K + 15) // 16 * 16
    Tp = np.zeros((Kp, N)); Tp[:K] = T
    Ap = np.zeros(A.shape[:-1] + (Kp,)); Ap[..., :K] = A
    Tp = r16(Tp, fp16)
    D = np.zeros(A.shape[:-1] + (N,))
    m_groups = (L + 127) // 128 * int(np.prod(A.shape[:-2]))        # per plane: ceil(lines / 128) instructions per (n, k) pair
    n_mma = 0
    for n0 in range(0, N, NB):
        nb = min(NB, N - n0)
        for k0 in range(0, Kp, 16):
            blk = Tp[k0:k0 + 16, n0:n0 + nb]
            if not blk.any():
                continue
            D[..., n0:n0 + nb] += Ap[..., k0:k0 + 16] @ blk
            n_mma += m_groups
    nbp = (min(NB, N) + 15) // 16 * 16                               # N of the instruction (multiple of 16 at M = 128)
    tensor_clk = 128 * nbp / 256.0
    smem_clk = (128 * 16 * 2 + 16 * nbp * 2) / 128.0
    st = stats.setdefault(name, dict(mma=0, clk=0.0, tensor_clk=0.0, smem_clk=0.0, lines=L, K=K, N=N, NB=nbp))
    st['mma'] += n_mma
    st['tensor_clk'] += n_mma * tensor_clk
    st['smem_clk'] += n_mma * smem_clk
    st['clk'] += n_mma * max(tensor_clk, smem_clk)
    return D


def filtered_lrelu_t5_model(x, fu, fd, b, up, down, padding, gain, slope, clamp, flip_filter=False, fp16=False,
                            tile=(56, 56), NB=(64, 64, 32, 32)):
    """-> (y, stats).  tile = (TH, TW) output pixels per CTA tile; NB = instruction N of the four passes
    (vertical up, horizontal up, vertical down, horizontal down)."""
    x = np.asarray(x, np.float64)
    N, C, xh, xw = x.shape
    px0, px1, py0, py1 = padding
    FU, FD = len(fu), len(fd)
    yw = (xw * up + px0 + px1 - (FU - 1) - (FD - 1) + down - 1) // down
    yh = (xh * up + py0 + py1 - (FU - 1) - (FD - 1) + down - 1) // down
    TH, TW = tile
    kux = taps_corr(fu, FU, flip_filter, up); kuy = taps_corr(fu, FU, flip_filter, up * gain)
    kd = taps_corr(fd, FD, flip_filter, 1.0)
    ax = Axis(up, down, kux, kd, px0, TW, False)
    ay = Axis(up, down, kuy, kd, py0, TH, False)
    bias = np.zeros(C) if b is None else np.asarray(b, np.float64)
    y = np.zeros((N, C, yh, yw))
    stats = {}
    tiles = 0
    for ty in range((yh + TH - 1) // TH):
        for tx in range((yw + TW - 1) // TW):
            tiles += 1
            cy, cx = ay.i0 + ty * ay.i_step, ax.i0 + tx * ax.i_step
            win = np.zeros((N, C, ay.NI, ax.NI))                       # [y, x]: the TMA box, zero filled outside the plane
            ys = np.arange(ay.NI) + cy; xs = np.arange(ax.NI) + cx
            oky = (ys >= 0) & (ys < xh); okx = (xs >= 0) & (xs < xw)
            iy, ix = np.where(oky)[0], np.where(okx)[0]
            if len(iy) and len(ix):
                win[:, :, iy[:, None], ix[None, :]] = x[:, :, ys[iy][:, None], xs[ix][None, :]] + bias[None, :, None, None]
            win = r16(win, fp16)
            # pass 1, vertical up: lines = x, contract y.  A = win^T as stored ([K = y rows][M = x contiguous] is the image
            # itself: MN-major A straight from the TMA box).  D1[x, v]
            d1 = blocked_product(np.swapaxes(win, -1, -2), ay.Tu, NB[0], stats, 'v_up', fp16)
            a2 = r16(d1, fp16)                                         # epilogue: thread x writes its row of v -> A2[K = x][M = v]
            # pass 2, horizontal up: lines = v, contract x.  D2[v, j]
            d2 = blocked_product(np.swapaxes(a2, -1, -2), ax.Tu, NB[1], stats, 'h_up', fp16)
            d2 = np.where(d2 < 0, d2 * slope, d2)
            if clamp is not None:
                d2 = np.clip(d2, -clamp, clamp)
            a3 = r16(d2, fp16)                                         # thread v writes its row of j -> A3[K = v][M = j]
            # pass 3, vertical down: lines = j, contract v.  D3[j, w]
            d3 = blocked_product(np.swapaxes(a3, -1, -2), ay.Td, NB[2], stats, 'v_down', fp16)
            a4 = r16(d3, fp16)                                         # thread j writes its row of w -> A4[K = j][M = w]
            # pass 4, horizontal down: lines = w, contract j.  D4[w, k]: thread w holds one output row -> row-major stores
            d4 = blocked_product(np.swapaxes(a4, -1, -2), ax.Td, NB[3], stats, 'h_down', fp16)
            h = min(TH, yh - ty * TH); w = min(TW, yw - tx * TW)
            y[:, :, ty * TH:ty * TH + h, tx * TW:tx * TW + w] = d4[:, :, :h, :w]
    planes = N * C
    summary = dict(tile=tile, NB=NB, geometry=dict(Y=ay.NI, X=ax.NI, V=ay.NU, J=ax.NU), passes={})
    tot_clk = tot_mma = 0.0
    for k, st in stats.items():
        per_tile = {q: st[q] / (tiles * planes) for q in ('mma', 'clk', 'tensor_clk', 'smem_clk')}
        summary['passes'][k] = dict(lines=st['lines'], K=st['K'], N=st['N'], NB=st['NB'], **per_tile)
        tot_clk += per_tile['clk']; tot_mma += per_tile['mma']
    # epilogue work per tile (thread-instructions): TMEM -> fp16 -> shared memory of the three intermediates + the output
    ep = 0.75 * (ax.NI * ay.NU) + 1.5 * (ay.NU * ax.NU) + 0.75 * (ax.NU * TH) + 0.75 * (TH * TW)
    summary.update(mma_per_tile=tot_mma, mma_clk_per_tile=tot_clk, mma_clk_per_256_outputs=tot_clk * 256.0 / (TH * TW),
                   epilogue_clk_per_256_outputs=ep / 128.0 * 256.0 / (TH * TW))      # 4 sub-partitions x 32 lanes per clock
    # bytes per 256 outputs at fp16 I/O: 256 outputs + their share of the input plane (ratio of the plane sizes)
    in_per_out = (down / up) ** 2
    summary['hbm_clk_per_256_outputs_fp16'] = 256.0 * 2 * (1.0 + in_per_out) / 26.0          # ~26 B/clk/SM at the measured copy rate
    return y, summary
