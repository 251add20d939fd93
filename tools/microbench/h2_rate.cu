// h2_rate.cu -- issue rate of packed-half instructions on sm_100a (which of them run at half rate on the FMA pipe, which on the
// ALU pipe), alone and interleaved with mma.sync, for the activation of csrc/flr_tc.cu.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o h2_rate.bin h2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { OP_MUL, OP_MULSAT, OP_ADD, OP_FMA, OP_MAX, OP_MIN, OP_FMA_RELU, OP_ACT_SAT, OP_ACT_MNMX, OP_ACT_FMA_MNMX, MMA_ACT_SAT, MMA_ACT_MNMX, MMA_ACT_FMA_MNMX, OP_IADD, OP_LOP };

template <int OP>
__device__ __forceinline__ uint32_t op1(uint32_t h, uint32_t c0, uint32_t c1)
{
    uint32_t r = h, p, q;
    if (OP == OP_MUL) asm volatile("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(h), "r"(c0));
    if (OP == OP_MULSAT) asm volatile("mul.rn.sat.f16x2 %0, %1, %2;" : "=r"(r) : "r"(h), "r"(c0));
    if (OP == OP_ADD) asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(h), "r"(c0));
    if (OP == OP_FMA) asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(h), "r"(c0), "r"(c1));
    if (OP == OP_FMA_RELU) asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(h), "r"(c0), "r"(c1));
    if (OP == OP_MAX) asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(h), "r"(c0));
    if (OP == OP_MIN) asm volatile("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(h), "r"(c0));
    if (OP == OP_IADD) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(h), "r"(c0));
    if (OP == OP_LOP) asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(h), "r"(c0));
    if (OP == OP_ACT_SAT || OP == MMA_ACT_SAT) {          // sat(u) - sat(-slope u)
        asm volatile("mul.rn.sat.f16x2 %0, %1, %2;" : "=r"(p) : "r"(h), "r"(c0));
        asm volatile("mul.rn.sat.f16x2 %0, %1, %2;" : "=r"(q) : "r"(h), "r"(c1));
        asm volatile("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(p), "r"(q));
    }
    if (OP == OP_ACT_MNMX || OP == MMA_ACT_MNMX) {        // min(max(max(u, slope u), -1), 1)
        asm volatile("mul.rn.f16x2 %0, %1, %2;" : "=r"(p) : "r"(h), "r"(c1));
        asm volatile("max.f16x2 %0, %1, %2;" : "=r"(q) : "r"(h), "r"(p));
        asm volatile("max.f16x2 %0, %1, %2;" : "=r"(p) : "r"(q), "r"(c0));
        asm volatile("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(p), "r"(c0));
    }
    if (OP == OP_ACT_FMA_MNMX || OP == MMA_ACT_FMA_MNMX) { // a u + b |u| as fma(b, |u|, a u) would need abs: model with mul + fma + min + max
        asm volatile("mul.rn.f16x2 %0, %1, %2;" : "=r"(p) : "r"(h), "r"(c1));
        asm volatile("max.f16x2 %0, %1, %2;" : "=r"(q) : "r"(h), "r"(p));
        asm volatile("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(q), "r"(c0));
    }
    return r;
}

template <int OP, int ILP>
__global__ void __launch_bounds__(1024) rate_kernel(int iters, uint32_t* out, long long* clk)
{
    constexpr bool MMA = OP == MMA_ACT_SAT || OP == MMA_ACT_MNMX || OP == MMA_ACT_FMA_MNMX;
    uint32_t h[ILP], dh[ILP][2];
    uint32_t a[4] = {threadIdx.x + 1u, threadIdx.x * 3u, 0x3c003c00u, 0x38003800u};
    uint32_t b0 = 0x3c003c00u + threadIdx.x, b1 = 0x34003400u;
#pragma unroll
    for (int i = 0; i < ILP; i++) { h[i] = 0x3c003c00u + i + threadIdx.x; dh[i][0] = i; dh[i][1] = i; }
    const uint32_t c0 = 0x3c003c00u, c1 = 0xb266b266u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MMA) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
                             : "+r"(dh[i][0]), "+r"(dh[i][1]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
                h[i] = op1<OP>(h[i], c0, c1);              // flr_tc: ~0.8 activations (of one packed pair) per HMMA + 0.35 adds
            } else {
#pragma unroll
                for (int r = 0; r < 4; r++) h[i] = op1<OP>(h[i], c0, c1);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc += h[i] + dh[i][0] + dh[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const char* name, int per_group, int sms, uint32_t* out, long long* clk)
{
    const int iters = 2048, ILP = 4, warps = 16;
    rate_kernel<OP, ILP><<<sms, warps * 32>>>(16, out, clk);
    rate_kernel<OP, ILP><<<sms, warps * 32>>>(iters, out, clk);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; i++) mean += (double)h[i];
    mean /= sms;
    const double groups = (double)iters * ILP * warps / 4.0 * per_group;
    printf("{\"kind\": \"%s\", \"clk_per_unit_per_smsp\": %.2f}\n", name, mean / groups);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; long long* clk;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 1024);
    cudaMalloc(&clk, sizeof(long long) * 256);
    run<OP_MUL>("mul.f16x2 (per instr)", 4, sms, out, clk);
    run<OP_MULSAT>("mul.sat.f16x2", 4, sms, out, clk);
    run<OP_ADD>("add.f16x2", 4, sms, out, clk);
    run<OP_FMA>("fma.f16x2", 4, sms, out, clk);
    run<OP_FMA_RELU>("fma.relu.f16x2", 4, sms, out, clk);
    run<OP_MAX>("max.f16x2", 4, sms, out, clk);
    run<OP_MIN>("min.f16x2", 4, sms, out, clk);
    run<OP_IADD>("add.u32", 4, sms, out, clk);
    run<OP_LOP>("xor.b32", 4, sms, out, clk);
    run<OP_ACT_SAT>("act sat form (3 instr, per activation)", 4, sms, out, clk);
    run<OP_ACT_MNMX>("act mul+max+max+min (4 instr)", 4, sms, out, clk);
    run<OP_ACT_FMA_MNMX>("act mul+max+min (3 instr, one-sided clamp)", 4, sms, out, clk);
    run<MMA_ACT_SAT>("hmma + act sat (per hmma)", 1, sms, out, clk);
    run<MMA_ACT_MNMX>("hmma + act mul+max+max+min", 1, sms, out, clk);
    run<MMA_ACT_FMA_MNMX>("hmma + act mul+max+min", 1, sms, out, clk);
    printf("done: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
