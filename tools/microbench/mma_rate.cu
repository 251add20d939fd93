// mma_rate.cu -- issue-rate microbenchmark of the legacy tensor path on sm_100a (mma.sync = HMMA), alone and interleaved
// with packed-half FMA-pipe work, as used by csrc/flr_tc.cu.  Prints clocks per HMMA per SM sub-partition for a range of
// resident warps and independent accumulator chains.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate.bin mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { K_F16ACC = 0, K_F32ACC = 1, K_F16ACC_K8 = 2, K_MIX = 3, K_HFMA = 4, K_BF16 = 5 };

template <int KIND, int ILP>
__global__ void __launch_bounds__(1024) rate_kernel(int iters, uint32_t* out, long long* clk)
{
    uint32_t a[4] = {threadIdx.x + 1u, threadIdx.x * 3u, 0x3c003c00u, 0x38003800u};
    uint32_t b0 = 0x3c003c00u + threadIdx.x, b1 = 0x34003400u;
    uint32_t dh[ILP][2];
    float df[ILP][4];
    uint32_t h[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { dh[i][0] = i; dh[i][1] = 2 * i; df[i][0] = df[i][1] = df[i][2] = df[i][3] = (float)i; h[i] = 0x3c003c00u + i; }
    const uint32_t one = 0x3c003c00u, ns = 0xb266b266u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (KIND == K_F16ACC || KIND == K_MIX)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
                             : "+r"(dh[i][0]), "+r"(dh[i][1]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            if (KIND == K_F16ACC_K8)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3}, {%4}, {%0,%1};"
                             : "+r"(dh[i][0]), "+r"(dh[i][1]) : "r"(a[0]), "r"(a[1]), "r"(b0));
            if (KIND == K_F32ACC)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(df[i][0]), "+f"(df[i][1]), "+f"(df[i][2]), "+f"(df[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            if (KIND == K_BF16)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(df[i][0]), "+f"(df[i][1]), "+f"(df[i][2]), "+f"(df[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            if (KIND == K_MIX || KIND == K_HFMA) {
                // the activation of flr_tc: 2 x mul.sat + sub per packed pair, ~5 FMA-pipe instructions per HMMA there
                uint32_t p, q;
#pragma unroll
                for (int r = 0; r < (KIND == K_MIX ? 2 : 4); r++) {
                    asm volatile("mul.rn.sat.f16x2 %0, %1, %2;" : "=r"(p) : "r"(h[i]), "r"(one));
                    asm volatile("mul.rn.sat.f16x2 %0, %1, %2;" : "=r"(q) : "r"(h[i]), "r"(ns));
                    asm volatile("sub.rn.f16x2 %0, %1, %2;" : "=r"(h[i]) : "r"(p), "r"(q));
                }
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc += dh[i][0] + dh[i][1] + __float_as_uint(df[i][0] + df[i][1] + df[i][2] + df[i][3]) + h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int KIND, int ILP>
static void run(const char* name, int warps_per_sm, int sms, uint32_t* out, long long* clk)
{
    const int iters = 4096;
    rate_kernel<KIND, ILP><<<sms, warps_per_sm * 32>>>(64, out, clk);
    rate_kernel<KIND, ILP><<<sms, warps_per_sm * 32>>>(iters, out, clk);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; i++) mean += (double)h[i];
    mean /= sms;
    const double per_smsp = (double)iters * ILP * warps_per_sm / 4.0;       // instruction groups issued per sub-partition
    printf("{\"kind\": \"%s\", \"warps_per_sm\": %d, \"ilp\": %d, \"clk_per_group_per_smsp\": %.2f}\n", name, warps_per_sm, ILP, mean / per_smsp);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; long long* clk;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 1024);
    cudaMalloc(&clk, sizeof(long long) * 256);
    // a "group" = one HMMA (+ 6 FMA-pipe instructions for mix, 12 FMA-pipe instructions alone for hfma)
    for (int w : {4, 8, 16}) {
        if (w == 4)  { run<K_F16ACC, 4>("hmma.16816.f16acc", 4, sms, out, clk); run<K_F32ACC, 4>("hmma.16816.f32acc", 4, sms, out, clk); }
        if (w == 8)  { run<K_F16ACC, 4>("hmma.16816.f16acc", 8, sms, out, clk); run<K_F32ACC, 4>("hmma.16816.f32acc", 8, sms, out, clk); }
        if (w == 16) { run<K_F16ACC, 4>("hmma.16816.f16acc", 16, sms, out, clk); run<K_F32ACC, 4>("hmma.16816.f32acc", 16, sms, out, clk); }
    }
    run<K_F16ACC, 1>("hmma.16816.f16acc", 16, sms, out, clk);
    run<K_F16ACC, 2>("hmma.16816.f16acc", 16, sms, out, clk);
    run<K_F16ACC, 8>("hmma.16816.f16acc", 16, sms, out, clk);
    run<K_F16ACC_K8, 4>("hmma.1688.f16acc", 16, sms, out, clk);
    run<K_BF16, 4>("hmma.16816.bf16.f32acc", 16, sms, out, clk);
    run<K_MIX, 4>("hmma.f16acc+6xfma-pipe", 16, sms, out, clk);
    run<K_MIX, 4>("hmma.f16acc+6xfma-pipe", 8, sms, out, clk);
    run<K_HFMA, 4>("12xfma-pipe(h2)", 16, sms, out, clk);
    printf("done: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
