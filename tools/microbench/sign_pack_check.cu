// sign_pack_check.cu -- hardware check of the warp-level sign packing planned for a sign mode of csrc/flr_tc.cu (DESIGN.md
// section 7 item 2; lane-level model: tools/flr_sign_pack_model.py).  Not part of the library.  One warp plays the vertical
// up pass: every lane holds the m16n8 accumulator fragments of MB column blocks x 2 row blocks (packed half2 values),
// derives the 2-bit codes (1 = negative, 2 = clamped) from the packed pre-activation / activation words, packs them with the
// exchange-and-OR butterfly and the funnel shift, and stores one aligned 32-bit word per (row, 16 columns).  The host compares
// with the definition of the reference sign layout (4 codes per byte along x) for every phase shift.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sign_pack_check.bin sign_pack_check.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

// ---- device functions meant for flr_tc.cu -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t h2_mul_sat(uint32_t a, uint32_t b) { uint32_t r; asm("mul.rn.sat.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_sub(uint32_t a, uint32_t b) { uint32_t r; asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// codes of the two packed samples of u (pre-activation in units of the clamp) and a = sat(u) - sat(-slope u):
// bits 0-1 = code of the low half, bits 16-17 = code of the high half
__device__ __forceinline__ uint32_t ftc_codes(uint32_t u, uint32_t a)
{
    uint32_t neg, cl;
    const uint32_t zero = 0u, one = 0x3c003c00u;
    asm("set.lt.u32.f16x2 %0, %1, %2;" : "=r"(neg) : "r"(u), "r"(zero));                 // 0xffff per half where u < 0
    const uint32_t mag = a & 0x7fff7fffu;
    asm("set.ge.u32.f16x2 %0, %1, %2;" : "=r"(cl) : "r"(mag), "r"(one));                  // |a| reached the clamp
    return (neg & ~cl & 0x00010001u) | (cl & 0x00020002u);
}

// c[q][h]: codes (as returned by ftc_codes) of row block q (rows 8 q + 2 t + {0, 1}) and register h (column 8 h + g) of one
// block of 16 up-sampled columns.  Returns the 32-bit word (16 codes, column i at bits 2 i) of row 8 (g >> 2) + 2 t + ((g >> 1) & 1);
// the two lanes of a pair (g even / odd) hold the same word.
__device__ __forceinline__ uint32_t ftc_sign_block_word(const uint32_t (&c)[2][2], unsigned lane)
{
    const uint32_t sh = 2u * (lane >> 2);
    uint32_t P[2][2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
        P[q][0] = (__byte_perm(c[q][0], c[q][1], 0x4400) & 0x00030003u) << sh;
        P[q][1] = (__byte_perm(c[q][0], c[q][1], 0x6622) & 0x00030003u) << sh;
    }
    const bool g2 = (lane & 16u) != 0, g1 = (lane & 8u) != 0;
    uint32_t R[2];
#pragma unroll
    for (int p = 0; p < 2; p++) {
        const uint32_t send = g2 ? P[0][p] : P[1][p], keep = g2 ? P[1][p] : P[0][p];
        R[p] = keep | __shfl_xor_sync(0xffffffffu, send, 16);
    }
    const uint32_t send = g1 ? R[0] : R[1], keep = g1 ? R[1] : R[0];
    const uint32_t S = keep | __shfl_xor_sync(0xffffffffu, send, 8);
    return S | __shfl_xor_sync(0xffffffffu, S, 4);
}

// read direction: the word of row (q, p) lives on lane 4 (4 q + 2 p) + t; every lane fetches its four rows and extracts the
// codes of its two columns.  m[q][h] = packed half2 multipliers (1, slope, 0) for register h of row block q.
__device__ __forceinline__ void ftc_sign_block_mult(uint32_t word_of_my_row, unsigned lane, uint32_t h_one_slope, uint32_t (&m)[2][2])
{
    const unsigned t = lane & 3u, sh = 2u * (lane >> 2);
#pragma unroll
    for (int q = 0; q < 2; q++) {
        uint32_t w[2];
#pragma unroll
        for (int p = 0; p < 2; p++) w[p] = __shfl_sync(0xffffffffu, word_of_my_row, (int)(4u * (4u * q + 2u * p) + t)) >> sh;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t c0 = (w[0] >> (16 * h)) & 3u, c1 = (w[1] >> (16 * h)) & 3u;      // codes of rows p = 0, 1
            // h_one_slope = {1.0 (low half), slope (high half)}: code 0 -> bytes 0,1; code 1 -> bytes 2,3; code >= 2 -> zero
            const uint32_t lo = c0 >= 2u ? 0u : __byte_perm(h_one_slope, 0u, c0 ? 0x4432 : 0x4410);
            const uint32_t hi = c1 >= 2u ? 0u : __byte_perm(h_one_slope, 0u, c1 ? 0x4432 : 0x4410);
            m[q][h] = lo | (hi << 16);
        }
    }
}

// ---- test kernels -----------------------------------------------------------------------------------------------------
template <int MB>
__global__ void pack_kernel(const __half* pre, float slope, int sx, uint32_t* words /* [16][MB] */, uint32_t* mult /* [MB][32][4] */)
{
    const unsigned lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    const __half2 hs2 = __floats2half2_rn(-slope, -slope), h1 = __floats2half2_rn(1.f, 1.f), os = __floats2half2_rn(1.f, slope);
    const uint32_t nslope = *reinterpret_cast<const uint32_t*>(&hs2), one = *reinterpret_cast<const uint32_t*>(&h1);
    const uint32_t one_slope = *reinterpret_cast<const uint32_t*>(&os);
    uint32_t T[MB + 1];
    T[MB] = 0u;
#pragma unroll
    for (int mb = 0; mb < MB; mb++) {
        uint32_t c[2][2];
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int J = 16 * mb + 8 * h + g, V = 8 * q + 2 * t;                    // pre[J][V], [J][V + 1]
                const __half2 uu = __halves2half2(pre[J * 16 + V], pre[J * 16 + V + 1]);
                const uint32_t u = *reinterpret_cast<const uint32_t*>(&uu);
                const uint32_t a = h2_sub(h2_mul_sat(u, one), h2_mul_sat(u, nslope));
                c[q][h] = ftc_codes(u, a);
            }
        T[mb] = ftc_sign_block_word(c, lane);
    }
    const int V = 8 * (g >> 2) + 2 * t + ((g >> 1) & 1);
#pragma unroll
    for (int w = 0; w < MB; w++) {
        const uint32_t o = __funnelshift_r(T[w], T[w + 1], 2 * sx);                       // column u = J - sx
        if (!(g & 1)) words[V * MB + w] = o;
        // read direction on the UNSHIFTED block words: every lane gets the multipliers of its own fragment elements back
        uint32_t m[2][2];
        ftc_sign_block_mult(T[w], lane, one_slope, m);
        for (int q = 0; q < 2; q++) for (int h = 0; h < 2; h++) mult[((w * 32 + lane) * 2 + q) * 2 + h] = m[q][h];
    }
}

static int code_of(float v, float slope)
{
    const float a = v < 0 ? v * slope : v;
    return std::fabs(a) >= 1.f ? 2 : (v < 0 ? 1 : 0);
}

template <int MB>
static int run(int sx, float slope)
{
    std::vector<__half> pre(16 * MB * 16);
    std::vector<float> pf(pre.size());
    for (size_t i = 0; i < pre.size(); i++) {
        float v = ((rand() % 2001) - 1000) / 250.f;                 // [-4, 4]: negatives beyond -1 / slope clamp as well at slope 0.5
        if (i % 17 == 0) v = 0.f;
        pre[i] = __float2half(v); pf[i] = __half2float(pre[i]);
        if (std::fabs(std::fabs(pf[i] < 0 ? pf[i] * slope : pf[i]) - 1.f) < 2e-2f) { pf[i] = 0.25f; pre[i] = __float2half(0.25f); }   // away from the threshold
    }
    __half* d_pre; uint32_t *d_w, *d_m;
    cudaMalloc(&d_pre, pre.size() * sizeof(__half)); cudaMalloc(&d_w, 16 * MB * 4); cudaMalloc(&d_m, MB * 32 * 4 * 4);
    cudaMemcpy(d_pre, pre.data(), pre.size() * sizeof(__half), cudaMemcpyHostToDevice);
    pack_kernel<MB><<<1, 32>>>(d_pre, slope, sx, d_w, d_m);
    std::vector<uint32_t> w(16 * MB), m(MB * 32 * 4);
    cudaMemcpy(w.data(), d_w, w.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(m.data(), d_m, m.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int V = 0; V < 16; V++)
        for (int wd = 0; wd < MB; wd++) {
            uint32_t ref = 0;
            for (int i = 0; i < 16; i++) {
                const int J = 16 * wd + i + sx;
                if (J < 16 * MB) ref |= (uint32_t)code_of(pf[J * 16 + V], slope) << (2 * i);
            }
            if (ref != w[V * MB + wd]) { if (bad < 4) printf("  word mismatch V=%d w=%d got %08x want %08x\n", V, wd, w[V * MB + wd], ref); bad++; }
        }
    const __half hs = __float2half(slope);
    for (int mb = 0; mb < MB; mb++)
        for (int lane = 0; lane < 32; lane++)
            for (int q = 0; q < 2; q++)
                for (int h = 0; h < 2; h++) {
                    const int g = lane >> 2, t = lane & 3, J = 16 * mb + 8 * h + g;
                    uint32_t ref = 0;
                    for (int p = 0; p < 2; p++) {
                        const int c = code_of(pf[J * 16 + 8 * q + 2 * t + p], slope);
                        const __half mv = c == 0 ? __float2half(1.f) : (c == 1 ? hs : __float2half(0.f));
                        ref |= (uint32_t)(*reinterpret_cast<const unsigned short*>(&mv)) << (16 * p);
                    }
                    const uint32_t got = m[((mb * 32 + lane) * 2 + q) * 2 + h];
                    if (ref != got) { if (bad < 8) printf("  mult mismatch mb=%d lane=%d q=%d h=%d got %08x want %08x\n", mb, lane, q, h, got, ref); bad++; }
                }
    cudaFree(d_pre); cudaFree(d_w); cudaFree(d_m);
    printf("{\"MB\": %d, \"sx\": %d, \"slope\": %.2f, \"mismatches\": %d}\n", MB, sx, slope, bad);
    return bad;
}

int main()
{
    int bad = 0;
    for (int sx = 0; sx < 4; sx++) { bad += run<3>(sx, 0.2f); bad += run<6>(sx, 0.5f); }
    printf("%s (%s)\n", bad ? "FAIL" : "PASS", cudaGetErrorString(cudaGetLastError()));
    return bad != 0;
}
