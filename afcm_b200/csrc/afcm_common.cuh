// afcm_common.cuh -- shared helpers for the afcm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/afcm_b200.h"

namespace afcm {

// Thread-local last error text, returned by afcm_last_error().
void set_error(const char* fmt, ...);

// C-ABI status convention: 0 ok, <0 argument / support errors, >0 a cudaError_t.
#define AFCM_CHECK_ARG(cond, ...)                                   \
    do { if (!(cond)) { ::afcm::set_error(__VA_ARGS__); return AFCM_ERR_INVALID; } } while (0)

#define AFCM_CUDA(call)                                             \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {          \
        ::afcm::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        return (int)e_; } } while (0)

#define AFCM_LAUNCH_CHECK()                                         \
    do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
        ::afcm::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
        return (int)e_; } } while (0)

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

int sm_count();           // cached multiprocessor count of the current device
int max_smem_optin();     // cached cudaDevAttrMaxSharedMemoryPerBlockOptin

// Running count of kernel launches issued through the library (bench.py reports it as gpu_launches).
void count_launch(int n = 1);

// Activation table shared by bias_act and the FC epilogues (indices follow the reference's cuda_idx,
// models/networks/stylegan3/torch_utils/ops/bias_act.py:21-31).
__device__ __forceinline__ float act_eval(float x, int act, float alpha)
{
    switch (act) {
    case 1: return x;
    case 2: return x > 0.f ? x : 0.f;
    case 3: return x > 0.f ? x : x * alpha;
    case 4: return tanhf(x);
    case 5: return 1.f / (1.f + expf(-x));
    case 6: return x > 0.f ? x : expf(x) - 1.f;
    case 7: return x > 0.f ? 1.0507009873554804934193349852946f * x
                           : 1.0507009873554804934193349852946f * 1.6732632423543772848170429916717f * (expf(x) - 1.f);
    case 8: return x > 20.f ? x : log1pf(expf(x));
    case 9: return x / (1.f + expf(-x));
    default: return x;
    }
}

}  // namespace afcm
