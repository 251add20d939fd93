// tma_box_rate.cu -- how fast does one SM's TMA unit deliver the operand tiles of csrc/conv2d_tc.cu?  148 CTAs, one thread each
// streams boxes of an fp16 [N][C][H*pitch] tensor (L2-resident working set) into a 4-deep shared-memory ring and waits for them;
// nothing consumes the data.  Variants:
//   0  tensor box [64 channels][152 elements], no swizzle, 16-byte aligned start  (the direct kernel's raw tile, 19 KB)
//   1  tensor box [64 channels][128 elements], no swizzle, 16-byte aligned start  (256-byte rows)
//   2  tensor box [64 channels][64 elements] x 2, SWIZZLE_128B, 128-byte aligned start
//   3  64 x cp.async.bulk of 304 bytes (one per channel row), issued by 32 lanes
//   4  tensor box [128 rows][64 elements] SWIZZLE_128B on a [rows][64] K-major tensor (the packed path's A tile, 16 KB)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_box_rate.bin tma_box_rate.cu -lcuda ; run on one B200.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
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
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int DEPTH = 4, SLOT = 20480;

__global__ void __launch_bounds__(32, 1) k_stream(const __grid_constant__ CUtensorMap map, const __half* x, long long* clk, int variant, int iters,
                                                   int HW, int C, int N)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[DEPTH];
    const int lane = threadIdx.x;
    if (lane == 0) { for (int i = 0; i < DEPTH; i++) mbar_init(&full[i], variant == 3 ? 32 : 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    const int tiles_per_plane = HW / 128 - 2;
    long long t0 = 0;
    for (int it = 0; it < iters + DEPTH; it++) {
        const int s = it % DEPTH;
        if (it >= DEPTH) mbar_wait(&full[s], ((it / DEPTH) - 1) & 1);          // the fill of DEPTH iterations ago has landed
        if (it == DEPTH) t0 = clock64();
        if (it < iters) {
            const int tile = (blockIdx.x + it * gridDim.x) % (tiles_per_plane * N);
            const int n = tile / tiles_per_plane, p0 = (tile % tiles_per_plane) * 128 + 128;
            const int cb = (it & 1) * 64 % C;
            uint8_t* dst = smem + s * SLOT;
            if (variant == 0) { if (lane == 0) { mbar_expect_tx(&full[s], 152 * 64 * 2); tma3(dst, &map, &full[s], p0 - 8, cb, n); } }
            else if (variant == 1) { if (lane == 0) { mbar_expect_tx(&full[s], 128 * 64 * 2); tma3(dst, &map, &full[s], p0 - 8, cb, n); } }
            else if (variant == 2) { if (lane == 0) { mbar_expect_tx(&full[s], 128 * 64 * 2); tma3(dst, &map, &full[s], p0, cb, n); tma3(dst + 8192, &map, &full[s], p0 + 64, cb, n); } }
            else if (variant == 3) {
                mbar_expect_tx(&full[s], 2 * 304);
                for (int r = lane; r < 64; r += 32)
                    bulk(dst + r * 304, x + ((size_t)n * C + cb + r) * HW + p0 - 8, 304, &full[s]);
            } else { if (lane == 0) { mbar_expect_tx(&full[s], 128 * 64 * 2); tma3(dst, &map, &full[s], 0, p0, n); } }
        }
        __syncwarp();
    }
    if (lane == 0) clk[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const int HW = 276 * 278, C = 64, N = 8;                       // 8 x 64 planes of 153 KB = 78 MB: L2-resident after the first pass
    __half* x; long long* clk;
    CK(cudaMalloc(&x, (size_t)N * C * HW * 2)); CK(cudaMemset(x, 0, (size_t)N * C * HW * 2));
    CK(cudaMalloc(&clk, 148 * 8));
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fnp;
    const char* names[5] = {"box [64][152] no swizzle (raw tile, 19 KB)", "box [64][128] no swizzle (16 KB)", "2 x box [64][64] SWIZZLE_128B (16 KB)",
                            "64 x cp.async.bulk 304 B (19 KB)", "box [128 rows][64] SWIZZLE_128B, K-major tensor (16 KB)"};
    CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, DEPTH * SLOT + 1024));
    for (int v = 0; v < 5; v++) {
        CUtensorMap map;
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r;
        if (v < 4) {
            cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)N}, st[2] = {(cuuint64_t)HW * 2, (cuuint64_t)HW * 2 * C};
            cuuint32_t box[3] = {(cuuint32_t)(v == 0 ? 152 : (v == 2 ? 64 : 128)), 64, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, x, dims, st, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    v == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            cuuint64_t dims[3] = {64, (cuuint64_t)HW, (cuuint64_t)N}, st[2] = {128, (cuuint64_t)HW * 128};
            cuuint32_t box[3] = {64, 128, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, x, dims, st, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) { printf("encode failed %d for variant %d\n", (int)r, v); continue; }
        const int iters = 2000;
        for (int rep = 0; rep < 2; rep++) {
            k_stream<<<148, 32, DEPTH * SLOT + 1024>>>(map, x, clk, v, iters, HW, C, N);
            CK(cudaDeviceSynchronize());
        }
        long long h[148]; CK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
        double mean = 0; for (int i = 0; i < 148; i++) mean += (double)h[i]; mean /= 148;
        const double bytes = (v == 0 || v == 3) ? 19456.0 : 16384.0;
        printf("%-58s %7.0f clk per tile per SM, %5.1f B/clk/SM (ring depth %d, 148 CTAs, L2-resident source)\n", names[v], mean / (iters - DEPTH), bytes * (iters - DEPTH) / mean, DEPTH);
    }
    return 0;
}
