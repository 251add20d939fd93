"""Lane-level model of packing the 2-bit sign codes of filtered_lrelu straight out of `mma.sync.m16n8k16` accumulator
fragments into the reference sign-tensor layout (4 codes per byte along x, OPS/filtered_lrelu.cpp:87-94) -- the recipe
for giving csrc/flr_tc.cu a sign-write / sign-read mode (the register-chained kernel is ~6x
faster than the shared-memory-tiled flr_tcs on the forward, so it is the kernel the training step should run).  Design
groundwork verified on the CPU (tests/test_flr_sign_pack_model.py); no kernel uses it yet.

Fragment layout (PTX m16n8 accumulator, as used by the vertical up pass of flr_tc: M = up-sampled column J, N = up-sampled
row V): lane l = 4 g + t holds, for the column block mb, row block q and register h, the two rows p = 0, 1:
        code(J = 16 mb + 8 h + g,  V = 8 q + 2 t + p).
One "chunk" = 16 rows (q = 0, 1) x 16 MB columns.  Per column block a lane therefore holds 8 codes, and the 16 rows of the
block want one 32-bit word each (16 codes).  Steps per column block (4 SHFL + 4 OR + 6 SEL):
  1. partial words  P[q][p] = c[q][0][p] << 2g  |  c[q][1][p] << (2g + 16)
  2. butterfly over the three g bits (lane bits 4, 3, 2): exchange-and-OR, keeping the row whose (q, p) equals the lane's
     (g bit 2, g bit 1); after the third step lanes with even g hold the complete word of row V = 8 (g >> 2) + 2 t + ((g >> 1) & 1)
  3. the sign tensor counts columns from the first sample the down filter reads, u = J - sx (sx = phase shift, 0 .. up-1):
     word w of a row = funnel shift of the words of column blocks w and w + 1 by 2 sx bits
so a strip of 16 D output columns (MB = D + 1 column blocks) stores D aligned 32-bit words per up-sampled row.
The read direction (backward) runs the same steps in reverse order."""
import numpy as np


def shfl_xor(v, mask):
    """v: [32] per-lane values -> value of lane (l ^ mask)."""
    return v[np.arange(32) ^ mask]


def pack_chunk(code, sx):
    """code: [16 * MB columns J][16 rows V] of 2-bit codes -> words [16 rows][MB - 1] (uint32) in sign-tensor order
    (code of column u = J - sx at bits 2 (u % 16) of word u // 16), computed the way a warp would."""
    MB = code.shape[0] // 16
    lane = np.arange(32)
    g, t = lane >> 2, lane & 3
    g2, g1, g0 = (g >> 2) & 1, (g >> 1) & 1, g & 1
    T = np.zeros((MB, 32), np.uint64)
    for mb in range(MB):
        # registers of the lane: c[q][h][p]
        c = np.zeros((2, 2, 2, 32), np.uint64)
        for q in range(2):
            for h in range(2):
                for p in range(2):
                    c[q, h, p] = code[16 * mb + 8 * h + g, 8 * q + 2 * t + p]
        P = np.zeros((2, 2, 32), np.uint64)
        for q in range(2):
            for p in range(2):
                P[q, p] = (c[q, 0, p] << (2 * g).astype(np.uint64)) | (c[q, 1, p] << (2 * g + 16).astype(np.uint64))
        # step A: lane bit 4 (g bit 2) selects the row block q it keeps
        R = np.zeros((2, 32), np.uint64)
        for p in range(2):
            send = np.where(g2 == 1, P[0, p], P[1, p])
            keep = np.where(g2 == 1, P[1, p], P[0, p])
            R[p] = keep | shfl_xor(send, 16)
        # step B: lane bit 3 (g bit 1) selects the row p it keeps
        send = np.where(g1 == 1, R[0], R[1])
        keep = np.where(g1 == 1, R[1], R[0])
        S = keep | shfl_xor(send, 8)
        # step C: lane bit 2 (g bit 0): plain reduction
        T[mb] = S | shfl_xor(S, 4)
    out = np.zeros((16, MB - 1), np.uint32)
    for l in lane[g0 == 0]:
        V = 8 * (g[l] >> 2) + 2 * t[l] + ((g[l] >> 1) & 1)
        for w in range(MB - 1):
            lo, hi = int(T[w, l]), int(T[w + 1, l])
            out[V, w] = ((lo >> (2 * sx)) | (hi << (32 - 2 * sx))) & 0xffffffff if sx else lo
    return out


def pack_reference(code, sx):
    """The same words by definition."""
    MB = code.shape[0] // 16
    out = np.zeros((16, MB - 1), np.uint32)
    for V in range(16):
        for w in range(MB - 1):
            word = 0
            for i in range(16):
                word |= int(code[16 * w + i + sx, V]) << (2 * i)
            out[V, w] = word
    return out


def bytes_of(words):
    """[rows][words] uint32 -> [rows][4 * words] uint8, little endian = the byte order of the sign tensor row."""
    return words.astype('<u4').view(np.uint8).reshape(words.shape[0], -1)
