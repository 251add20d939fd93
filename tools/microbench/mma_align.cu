// mma_align.cu -- does a tcgen05.mma (kind::f16, A and B in shared memory, K-major SWIZZLE_128B) get slower when the A
// descriptor starts 1 or 2 rows (128 / 256 bytes) into the 8-row swizzle atom?  That is what the row-reuse trick of
// csrc/conv2d_tc.cu does for the kx = 1, 2 taps (one staged tile of 136 pixels serves three taps through 128-byte descriptor
// offsets).  Also: the same MMAs issued in the conv's order (3 taps x 4 K steps on one accumulator).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_align.bin mma_align.cu ; run on one B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    long long t0 = clock64();
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 2000000000LL) __trap();
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                 ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
}

// mode 0: every MMA reads A at row offset `rowofs`; mode 1: the conv's order -- taps kx = 0, 1, 2 (row offsets 0, 1, 2), four K steps each
__global__ void __launch_bounds__(128, 1) k_mma(long long* clk, int N, int iters, int rowofs, int mode)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (20480 + 3 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    if (warp == 0) {
        const uint32_t hi = (uint32_t)((1024u >> 4) & 0x3fff) | (1u << 14) | (2u << 29);
        const uint32_t a_lo = ((smem_u32(smem) >> 4) & 0x3fff) | (1u << 16), b_lo = ((smem_u32(smem + 20480) >> 4) & 0x3fff) | (1u << 16);
        const uint32_t b_tap = (uint32_t)((N * 128) >> 4);
        const long long t0 = clock64();
        for (int it = 0; it < iters; it += 12) {
            if (elect_one()) {
#pragma unroll
                for (int u = 0; u < 12; u++) {
                    const int kx = mode ? u / 4 : 0, k = u & 3;
                    mma_ss(base, a_lo + (uint32_t)((mode ? kx : rowofs) * 8 + k * 2), b_lo + (uint32_t)(kx * b_tap + k * 2), hi, idesc);
                }
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (threadIdx.x == 0) clk[blockIdx.x] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory"); }
}

int main()
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* clk;
    CK(cudaMallocManaged(&clk, sizeof(long long) * sms));
    const int smem = 1024 + 20480 + 3 * 32768;
    CK(cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int Ns[4] = {64, 96, 128, 256};
    const int iters = 4800;
    for (int ni = 0; ni < 4; ni++) {
        for (int cfg = 0; cfg < 4; cfg++) {
            const int mode = cfg == 3, rowofs = cfg == 3 ? 0 : cfg;
            for (int rep = 0; rep < 2; rep++) { k_mma<<<sms, 128, smem>>>(clk, Ns[ni], iters, rowofs, mode); CK(cudaDeviceSynchronize()); }
            long long mx = 0;
            for (int i = 0; i < sms; i++) if (clk[i] > mx) mx = clk[i];
            printf("f16 SS M=128 N=%3d %s: %.2f clk per MMA (floor N/2 = %d)\n", Ns[ni],
                   cfg == 0 ? "A at row offset 0        " : cfg == 1 ? "A at row offset 1 (128 B)" : cfg == 2 ? "A at row offset 2 (256 B)" : "conv order: taps 0, 1, 2 ",
                   (double)mx / iters, Ns[ni] / 2);
        }
    }
    return 0;
}
