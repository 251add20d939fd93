// CPU emulation of the fused filtered_lrelu kernel: runs the exact pass functions of
// afcm_b200/csrc/filtered_lrelu_core.h one "thread" at a time (a __syncthreads() is the boundary
// between two loops over tid).  Test infrastructure: validates tile/index math without a GPU.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../afcm_b200/csrc/filtered_lrelu_core.h"

using namespace afcm;

template <int UP, int FU, int DOWN, int FD, int G>
static int run(FlrParams p, int tow, int toh, int sign_mode, int nthr)
{
    size_t bytes = flr_make_geom<UP, FU, DOWN, FD, G>(p, tow, toh, sign_mode);
    std::vector<float> smem(bytes / 4 + 64);
    const long long tiles = (long long)p.N * p.C * p.tiles_x * p.tiles_y;
    for (long long tile = 0; tile < tiles; tile++) {
        // poison shared memory so that reads of never-written words show up as NaN
        for (auto& v : smem) v = __builtin_nanf("");
        float* a = smem.data(); float* b = smem.data() + p.off_b;
        uint8_t* ss = reinterpret_cast<uint8_t*>(smem.data()) + p.off_sign;
        FlrTile t = flr_tile<UP, DOWN>(p, (int)tile);
        for (int tid = 0; tid < nthr; tid++) flr_pass_load<float>(tid, nthr, p, t, a);
        if (sign_mode == 2) for (int tid = 0; tid < nthr; tid++) flr_pass_sign_load(tid, nthr, p, t, ss);
        for (int tid = 0; tid < nthr; tid++) flr_pass_hup<UP, FU, G>(tid, nthr, p, a, b);
        for (int tid = 0; tid < nthr; tid++) {
            if (sign_mode == 0) flr_pass_vup<UP, FU, G, 0>(tid, nthr, p, t, b, a, ss);
            if (sign_mode == 1) flr_pass_vup<UP, FU, G, 1>(tid, nthr, p, t, b, a, ss);
            if (sign_mode == 2) flr_pass_vup<UP, FU, G, 2>(tid, nthr, p, t, b, a, ss);
        }
        if (sign_mode == 1) for (int tid = 0; tid < nthr; tid++) flr_pass_sign_flush<DOWN>(tid, nthr, p, t, ss);
        for (int tid = 0; tid < nthr; tid++) flr_pass_hdown<DOWN, FD, G>(tid, nthr, p, a, b);
        for (int tid = 0; tid < nthr; tid++) flr_pass_vdown<float, DOWN, FD, G>(tid, nthr, p, t, b);
    }
    return 0;
}

extern "C" int emu_filtered_lrelu(const float* x, float* y, const float* b, const float* skip,
                                  int N, int C, int xh, int xw, int yh, int yw,
                                  const float* fu, int fu_taps, const float* fd, int fd_taps,
                                  int up, int down, int px0, int py0,
                                  float gain, float slope, float clamp, float out_scale, int flip,
                                  int sign_mode, uint8_t* signs, int sign_h, int sign_wb, int sx, int sy,
                                  int tow, int toh, int nthr)
{
    FlrParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.y = y; p.b = b; p.skip = skip;
    p.so = sign_mode == 1 ? signs : nullptr; p.si = sign_mode == 2 ? signs : nullptr;
    p.xs_n = (long long)C * xh * xw; p.xs_c = (long long)xh * xw; p.xs_h = xw; p.xs_w = 1;
    p.ys_n = (long long)C * yh * yw; p.ys_c = (long long)yh * yw; p.ys_h = yw; p.ys_w = 1;
    p.N = N; p.C = C; p.xh = xh; p.xw = xw; p.yh = yh; p.yw = yw; p.px0 = px0; p.py0 = py0;
    p.s_h = sign_h; p.s_wb = sign_wb; p.s_ox = sx; p.s_oy = sy;
    p.gain = gain; p.slope = slope; p.clamp = clamp; p.out_scale = out_scale;
    for (int t = 0; t < fu_taps; t++) p.ku[t] = fu[flip ? t : fu_taps - 1 - t] * (float)up;
    for (int t = 0; t < fd_taps; t++) p.kd[t] = fd[flip ? t : fd_taps - 1 - t];
    if (up == 2 && fu_taps == 12 && down == 2 && fd_taps == 12) return run<2, 12, 2, 12, 8>(p, tow, toh, sign_mode, nthr);
    if (up == 2 && fu_taps == 12 && down == 4 && fd_taps == 24) return run<2, 12, 4, 24, 8>(p, tow, toh, sign_mode, nthr);
    if (up == 4 && fu_taps == 24 && down == 2 && fd_taps == 12) return run<4, 24, 2, 12, 8>(p, tow, toh, sign_mode, nthr);
    return -1;
}
