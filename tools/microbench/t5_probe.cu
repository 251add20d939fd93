// t5_probe.cu -- hardware facts the tcgen05 / TMEM filtered_lrelu (csrc/flr_t5.cu) is designed around, measured on B200:
//   A. tcgen05.ld / tcgen05.st throughput per SM (4 and 8 warps, several shapes / waits)
//   B. tcgen05.mma issue rate for small N with the A operand in tensor memory (kind::f16 and kind::tf32), M = 128
//   C. layout checks: f16 A operand in TMEM (packed pairs), an fp32 accumulator consumed directly as a tf32 A operand,
//      accumulator / A column offsets that are not multiples of 16, back-to-back dependent MMAs without a commit
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o t5_probe.bin t5_probe.cu ; run on one B200.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

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
// K-major SWIZZLE_128B descriptor: SBO 1024 B, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)(16u >> 4) << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
#define LD32_REGS(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
#define ST32_REGS(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), \
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), \
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : LD32_REGS(r) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
                 :: ST32_REGS(r), "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t* slot, int warp)
{
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *slot;
}
__device__ __forceinline__ void tmem_free_all(uint32_t base, int warp)
{
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ A: ld / st rates
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
                 :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr) : "memory");
}
// NLD x32 loads in flight per wait
template <int NLD>
__global__ void __launch_bounds__(256, 1) k_ld(long long* clk, uint32_t* sink, int iters)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t base = tmem_alloc_all(&slot, warp);
    const uint32_t lane_base = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    uint32_t acc = 0;
    uint32_t r[NLD][32];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 8; c += NLD) {
#pragma unroll
            for (int j = 0; j < NLD; j++) tmem_ld32(lane_base + (c + j) * 32, r[j]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < NLD; j++) acc ^= r[j][0] ^ r[j][31] ^ r[j][13];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    tmem_free_all(base, warp);
}
template <int NST>
__global__ void __launch_bounds__(256, 1) k_st(long long* clk, uint32_t* sink, int iters)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t base = tmem_alloc_all(&slot, warp);
    const uint32_t lane_base = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = i + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 8; c += NST) {
#pragma unroll
            for (int j = 0; j < NST; j++) tmem_st32(lane_base + (c + j) * 32, r);
            tmem_st_wait();
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    tmem_free_all(base, warp);
}
// the activation epilogue of the planned kernel: ld 32 fp32 columns -> packed half2 -> sat(u) - sat(-slope u) -> st 16 columns,
// the next load issued before the current chunk is processed (DB = 1) or not (DB = 0)
__device__ __forceinline__ uint32_t act_pair(uint32_t a, uint32_t b, uint32_t one, uint32_t nslope)
{
    uint32_t h, s0, s1, o;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(__uint_as_float(b)), "f"(__uint_as_float(a)));
    asm("mul.rn.sat.f16x2 %0, %1, %2;" : "=r"(s0) : "r"(h), "r"(one));
    asm("mul.rn.sat.f16x2 %0, %1, %2;" : "=r"(s1) : "r"(h), "r"(nslope));
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(o) : "r"(s0), "r"(s1));
    return o;
}
template <int DB>
__global__ void __launch_bounds__(256, 1) k_epi(long long* clk, uint32_t* sink, int iters)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t base = tmem_alloc_all(&slot, warp);
    const uint32_t lane_base = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    const uint32_t one = 0x3c003c00u, nslope = 0xb266b266u;      // 1.0, -0.2
    uint32_t ra[32], rb[32], o[16];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (DB) {
            tmem_ld32(lane_base, ra);
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                tmem_ld_wait();
                tmem_ld32(lane_base + (c + 1) * 32, rb);
#pragma unroll
                for (int i = 0; i < 16; i++) o[i] = act_pair(ra[2 * i], ra[2 * i + 1], one, nslope);
                tmem_st16(lane_base + c * 16, o);
                tmem_ld_wait();
                if (c + 2 < 8) tmem_ld32(lane_base + (c + 2) * 32, ra);
#pragma unroll
                for (int i = 0; i < 16; i++) o[i] = act_pair(rb[2 * i], rb[2 * i + 1], one, nslope);
                tmem_st16(lane_base + (c + 1) * 16, o);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                tmem_ld32(lane_base + c * 32, ra);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) o[i] = act_pair(ra[2 * i], ra[2 * i + 1], one, nslope);
                tmem_st16(lane_base + c * 16, o);
            }
        }
        tmem_st_wait();
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    tmem_free_all(base, warp);
}

// ------------------------------------------------------------------------------------------------ B: MMA issue rate
// mode 0: kind::f16, A and B in shared memory; 1: kind::f16, A in TMEM; 2: kind::tf32, A in TMEM.  drot: rotate the accumulator
// over 4 column ranges.  The whole warp runs the loop, one elected lane issues (warp-uniform control flow).
__device__ __forceinline__ void mma_lohi(int mode, uint32_t d, uint32_t a_lo_or_tmem, uint32_t b_lo, uint32_t hi, uint32_t idesc)
{
    if (mode == 0)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                     ::"r"(d), "r"(a_lo_or_tmem), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
    else if (mode == 1)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 db, {%2, %3};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
                     ::"r"(d), "r"(a_lo_or_tmem), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 db, {%2, %3};\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
                     ::"r"(d), "r"(a_lo_or_tmem), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_mma(long long* clk, int N, int iters, int drot)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t base = tmem_alloc_all(&slot, warp);
    const uint32_t fmt = MODE == 2 ? 2u : 0u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    if (warp == 0) {
        const uint32_t hi = (uint32_t)((1024u >> 4) & 0x3fff) | (1u << 14) | (2u << 29);
        const uint32_t a_lo = ((smem_u32(smem) >> 4) & 0x3fff) | (1u << 16), b_lo = ((smem_u32(smem + 16384) >> 4) & 0x3fff) | (1u << 16);
        const long long t0 = clock64();
        for (int it = 0; it < iters; it += 8) {
            if (elect_one()) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const uint32_t d = base + 256 + (drot ? (uint32_t)((u & 3) * 64) : 0u);
                    mma_lohi(MODE, d, MODE == 0 ? a_lo + (uint32_t)((u & 3) * 2) : base + (uint32_t)(u * 8), b_lo + (uint32_t)((u & 3) * 2), hi, idesc);
                }
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (threadIdx.x == 0) { clk[blockIdx.x * 2] = t2 - t0; clk[blockIdx.x * 2 + 1] = t1 - t0; }
    }
    tmem_free_all(base, warp);
}

// ------------------------------------------------------------------------------------------------ C: layout checks
// B tile in K-major SWIZZLE_128B form: row n at n * 128 B, 16-byte chunk c at position c ^ (n & 7)
__device__ __forceinline__ uint32_t sw128(int n, int byte_in_row) { return (uint32_t)(n * 128 + ((((byte_in_row >> 4) ^ (n & 7)) << 4) | (byte_in_row & 15))); }

struct ProbeOut {
    float d1[128][32];        // C1: f16 TS product, N = 16 (cols 0..15), D column offset 0
    float d1o[128][32];       // C5: same product written at accumulator column offset +8 and read back
    float d2[128][16];        // C2: tf32 TS product with A = d1 accumulator columns 0..7
    float d2o[128][16];       // C5b: tf32 TS with A = d1 columns 4..11 (A column offset 4)
    float d3[128][16];        // C3: dependent chain without commit between the MMAs
    float d4[128][16];        // C6: f16 A operand at TMEM column offset +4 (K = 16 starting at packed column 4)
};

__global__ void __launch_bounds__(128, 1) k_probe(ProbeOut* out)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, m = threadIdx.x;
    uint8_t* sB1 = smem;            // f16 [16 n][16 k]        (2 KB, two 8-row atoms)
    uint8_t* sB2 = smem + 4096;     // tf32 [16 n][8 k]
    uint8_t* sBig = smem + 8192;    // f16 [256 n][16 k] for the long first product of the chain test (32 KB)
    for (int i = threadIdx.x; i < (8192 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    __syncthreads();
    // B1[n][k] = ((n * 3 + k) % 5) - 2 ; B2[n][k] = ((n + 2 k) % 5) - 2 ; Big[n][k] = B1[n % 16][k]
    for (int i = threadIdx.x; i < 16 * 16; i += blockDim.x) {
        const int n = i / 16, k = i % 16;
        *reinterpret_cast<__half*>(sB1 + sw128(n, k * 2)) = __float2half((float)(((n * 3 + k) % 5) - 2));
    }
    for (int i = threadIdx.x; i < 16 * 8; i += blockDim.x) {
        const int n = i / 8, k = i % 8;
        *reinterpret_cast<float*>(sB2 + sw128(n, k * 4)) = (float)(((n + 2 * k) % 5) - 2);
    }
    for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) {
        const int n = i / 16, k = i % 16;
        *reinterpret_cast<__half*>(sBig + sw128(n, k * 2)) = __float2half((float)((((n % 16) * 3 + k) % 5) - 2));
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t base = tmem_alloc_all(&slot, warp);
    const uint32_t lane_base = base + ((uint32_t)(warp * 32) << 16);
    // TMEM map (columns): A f16 packed at 0..15 (A[m][k], k < 32: two K chunks); D1 at 64; D1o at 128 + 8; D2 at 192; D2o at 224;
    // chain: E1 at 256 (N = 256), E2 at 32
    // A[m][k] = ((m + k) % 7) - 3 for k < 16;  A[m][16 + k] = ((m * 2 + k) % 5) - 2   (second chunk, read at column offset 8)
    {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 32; j++) r[j] = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int k0 = 2 * j, k1 = 2 * j + 1;
            const float v0 = k0 < 16 ? (float)(((m + k0) % 7) - 3) : (float)(((m * 2 + k0 - 16) % 5) - 2);
            const float v1 = k1 < 16 ? (float)(((m + k1) % 7) - 3) : (float)(((m * 2 + k1 - 16) % 5) - 2);
            __half2 h = __floats2half2_rn(v0, v1);            // low half = even k
            r[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        tmem_st32(lane_base + 0, r);
        // poison the accumulators that the chain test reads as an operand
#pragma unroll
        for (int j = 0; j < 32; j++) r[j] = __float_as_uint(1.0e30f);
        tmem_st32(lane_base + 256, r);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    const uint32_t id_f16_n16 = (1u << 4) | (2u << 17) | (8u << 24);
    const uint32_t id_f16_n256 = (1u << 4) | (32u << 17) | (8u << 24);
    const uint32_t id_tf32_n16 = (1u << 4) | (2u << 7) | (2u << 10) | (2u << 17) | (8u << 24);
    uint32_t phase = 0;
    if (threadIdx.x == 0) {
        tc_fence_after();
        mma_f16_ts(base + 64, base + 0, make_desc(smem_u32(sB1)), id_f16_n16, 0);          // C1
        mma_f16_ts(base + 128 + 8, base + 0, make_desc(smem_u32(sB1)), id_f16_n16, 0);     // C5: D column offset 8
        mma_f16_ts(base + 160 + 4, base + 4, make_desc(smem_u32(sB1)), id_f16_n16, 0);     // C6: A column offset 4 (k = 8..23), D column offset 4
        umma_commit(&bar);
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        mma_tf32_ts(base + 192, base + 64, make_desc(smem_u32(sB2)), id_tf32_n16, 0);      // C2: A = D1[:, 0..7]
        mma_tf32_ts(base + 224, base + 64 + 4, make_desc(smem_u32(sB2)), id_tf32_n16, 0);  // C5b: A = D1[:, 4..11]
        umma_commit(&bar);
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        // C3: E1 = A * Big (N = 256, 8 accumulating repeats -> 8 * A * Big), then immediately E2 = E1[:, 0..7] (tf32) * B2
        for (int rep = 0; rep < 8; rep++) mma_f16_ts(base + 256, base + 0, make_desc(smem_u32(sBig)), id_f16_n256, rep > 0);
        mma_tf32_ts(base + 32, base + 256, make_desc(smem_u32(sB2)), id_tf32_n16, 0);
        umma_commit(&bar);
        mbar_wait(&bar, phase); phase ^= 1;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    {
        uint32_t r[32];
        tmem_ld32(lane_base + 64, r); tmem_ld_wait();
        for (int j = 0; j < 32; j++) out->d1[m][j] = __uint_as_float(r[j]);
        tmem_ld32(lane_base + 128, r); tmem_ld_wait();
        for (int j = 0; j < 32; j++) out->d1o[m][j] = __uint_as_float(r[j]);
        tmem_ld32(lane_base + 192, r); tmem_ld_wait();
        for (int j = 0; j < 16; j++) out->d2[m][j] = __uint_as_float(r[j]);
        tmem_ld32(lane_base + 224, r); tmem_ld_wait();
        for (int j = 0; j < 16; j++) out->d2o[m][j] = __uint_as_float(r[j]);
        tmem_ld32(lane_base + 32, r); tmem_ld_wait();
        for (int j = 0; j < 16; j++) out->d3[m][j] = __uint_as_float(r[j]);
        tmem_ld32(lane_base + 160, r); tmem_ld_wait();
        for (int j = 0; j < 16; j++) out->d4[m][j] = __uint_as_float(r[j + 4]);
    }
    (void)lane;
    tmem_free_all(base, warp);
}

static float A_val(int m, int k) { return k < 16 ? (float)(((m + k) % 7) - 3) : (float)(((m * 2 + k - 16) % 5) - 2); }
static float B1_val(int n, int k) { return (float)(((n * 3 + k) % 5) - 2); }
static float B2_val(int n, int k) { return (float)(((n + 2 * k) % 5) - 2); }

int main(int argc, char** argv)
{
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    printf("device %s, %d SMs, max clock %d MHz\n", prop.name, prop.multiProcessorCount, clk_khz / 1000);
    const int sms = prop.multiProcessorCount;
    long long* clk; CK(cudaMallocManaged(&clk, sizeof(long long) * 2 * sms));
    uint32_t* sink; CK(cudaMalloc(&sink, 4096));

    // ---- C: layout checks first (they decide whether the design works at all)
    {
        ProbeOut* out; CK(cudaMallocManaged(&out, sizeof(ProbeOut)));
        memset(out, 0, sizeof(ProbeOut));
        CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
        k_probe<<<1, 128, 48 * 1024>>>(out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("k_probe failed: %s\n", cudaGetErrorString(e)); return 1; }
        // references
        double e1 = 0, e1sw = 0, e1o = 0, e2 = 0, e2o = 0, e3 = 0, e4 = 0;
        for (int m = 0; m < 128; m++) {
            float d1ref[16];
            for (int n = 0; n < 16; n++) {
                float s = 0, ssw = 0;
                for (int k = 0; k < 16; k++) { s += A_val(m, k) * B1_val(n, k); ssw += A_val(m, k ^ 1) * B1_val(n, k); }
                d1ref[n] = s;
                e1 = fmax(e1, fabs(out->d1[m][n] - s));
                e1sw = fmax(e1sw, fabs(out->d1[m][n] - ssw));
                e1o = fmax(e1o, fabs(out->d1o[m][n + 8] - s));
                float s4 = 0;
                for (int k = 0; k < 16; k++) s4 += A_val(m, k + 8) * B1_val(n, k);
                e4 = fmax(e4, fabs(out->d4[m][n] - s4));
            }
            for (int n = 0; n < 16; n++) {
                float s = 0, so = 0, s3 = 0;
                for (int k = 0; k < 8; k++) { s += d1ref[k] * B2_val(n, k); so += d1ref[k + 4] * B2_val(n, k); s3 += 8.f * d1ref[k] * B2_val(n, k); }
                e2 = fmax(e2, fabs(out->d2[m][n] - s));
                e2o = fmax(e2o, fabs(out->d2o[m][n] - so));
                e3 = fmax(e3, fabs(out->d3[m][n] - s3));
            }
        }
        printf("C1 f16 A in TMEM (low half = even k): max err %.3g  (if halves swapped: %.3g)  -> %s\n", e1, e1sw, e1 == 0 ? "PASS" : "FAIL");
        printf("C5 accumulator at column offset 8:      max err %.3g -> %s\n", e1o, e1o == 0 ? "PASS" : "FAIL");
        printf("C6 f16 A at TMEM column offset 4, D at column offset 4: max err %.3g -> %s\n", e4, e4 == 0 ? "PASS" : "FAIL");
        printf("C2 fp32 accumulator as tf32 A operand:  max err %.3g -> %s\n", e2, e2 == 0 ? "PASS" : "FAIL");
        printf("C5b tf32 A at column offset 4:          max err %.3g -> %s\n", e2o, e2o == 0 ? "PASS" : "FAIL");
        printf("C3 dependent MMA chain without commit:  max err %.3g -> %s\n", e3, e3 == 0 ? "PASS (ordered)" : "FAIL (needs commit + wait)");
        printf("   sample d1[5][0..3] = %g %g %g %g ; d3[5][0..3] = %g %g %g %g\n", out->d1[5][0], out->d1[5][1], out->d1[5][2], out->d1[5][3],
               out->d3[5][0], out->d3[5][1], out->d3[5][2], out->d3[5][3]);
    }

    // ---- A: ld / st throughput
    {
        struct Case { const char* name; void (*fn)(long long*, uint32_t*, int); double bytes_per_iter_per_warp; };
        const Case cases[] = {
            {"ld x32, 1 per wait", k_ld<1>, 32.0 * 32 * 4 * 8}, {"ld x32, 2 per wait", k_ld<2>, 32.0 * 32 * 4 * 8},
            {"ld x32, 4 per wait", k_ld<4>, 32.0 * 32 * 4 * 8},
            {"st x32, 1 per wait", k_st<1>, 32.0 * 32 * 4 * 8}, {"st x32, 4 per wait", k_st<4>, 32.0 * 32 * 4 * 8},
            {"act epilogue ld32 -> half2 act -> st16, serial", k_epi<0>, 32.0 * 32 * 4 * 8},
            {"act epilogue ld32 -> half2 act -> st16, prefetched", k_epi<1>, 32.0 * 32 * 4 * 8}};
        for (const Case& c : cases) {
            for (int threads = 128; threads <= 256; threads += 128) {
                const int iters = 2000;
                c.fn<<<sms, threads>>>(clk, sink, iters);
                CK(cudaDeviceSynchronize());
                c.fn<<<sms, threads>>>(clk, sink, iters);
                CK(cudaDeviceSynchronize());
                long long mn = clk[0], mx = clk[0];
                for (int i = 1; i < sms; i++) { if (clk[i] < mn) mn = clk[i]; if (clk[i] > mx) mx = clk[i]; }
                const double bytes = (threads / 32) * c.bytes_per_iter_per_warp * iters;
                printf("A %-52s %d threads: %.1f B/clk/SM of TMEM columns read (or written)  (clk min %lld max %lld; %.1f clk per x32 op per warp)\n",
                       c.name, threads, bytes / (double)mx, mn, mx, (double)mx / (iters * 8.0));
            }
        }
    }

    // ---- B: MMA issue rates
    {
        CK(cudaFuncSetAttribute(k_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CK(cudaFuncSetAttribute(k_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CK(cudaFuncSetAttribute(k_mma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        const char* mnames[3] = {"f16 SS", "f16 TS (A in TMEM)", "tf32 TS (A in TMEM)"};
        const int Ns[7] = {16, 32, 48, 64, 96, 128, 256};
        for (int mode = 0; mode < 3; mode++) {
            for (int ni = 0; ni < 7; ni++) {
                for (int drot = 0; drot < 2; drot++) {
                    const int N = Ns[ni];
                    if (drot && N > 64) continue;
                    const int iters = 4000;
                    for (int rep = 0; rep < 2; rep++) {
                        if (mode == 0) k_mma<0><<<sms, 128, 64 * 1024>>>(clk, N, iters, drot);
                        else if (mode == 1) k_mma<1><<<sms, 128, 64 * 1024>>>(clk, N, iters, drot);
                        else k_mma<2><<<sms, 128, 64 * 1024>>>(clk, N, iters, drot);
                        CK(cudaDeviceSynchronize());
                    }
                    long long mx = 0, mxi = 0;
                    for (int i = 0; i < sms; i++) { if (clk[2 * i] > mx) mx = clk[2 * i]; if (clk[2 * i + 1] > mxi) mxi = clk[2 * i + 1]; }
                    printf("B %s N=%3d %s: %.2f clk per MMA (issue loop alone %.2f); floor 128*N/256 = %.1f%s\n", mnames[mode], N,
                           drot ? "4 accumulators" : "1 accumulator ", (double)mx / iters, (double)mxi / iters,
                           128.0 * N / 256.0, mode == 2 ? " (K = 8)" : " (K = 16)");
                }
            }
        }
    }
    printf("done\n");
    return 0;
}
