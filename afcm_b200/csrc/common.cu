// common.cu -- error text, device attribute cache, launch counter, library-level C-ABI entry points.
#include <stdarg.h>
#include <atomic>
#include "afcm_common.cuh"

namespace afcm {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int device_attr(cudaDeviceAttr attr)
{
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&v, attr, dev) != cudaSuccess) return 0;
    return v;
}

// attribute caches are per device: a process may drive several GPUs (DataParallel-style callers, `device=` arguments)
static int cached_attr(int (&cache)[64], cudaDeviceAttr attr)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return device_attr(attr);
    if (!cache[dev]) cache[dev] = device_attr(attr);
    return cache[dev];
}

int sm_count()
{
    static int v[64] = {0};
    const int r = cached_attr(v, cudaDevAttrMultiProcessorCount);
    return r > 0 ? r : 148;
}

int max_smem_optin()
{
    static int v[64] = {0};
    const int r = cached_attr(v, cudaDevAttrMaxSharedMemoryPerBlockOptin);
    return r > 0 ? r : 232448;
}

}  // namespace afcm

extern "C" int afcm_version(void) { return 1; }

extern "C" const char* afcm_last_error(void) { return afcm::g_err; }

extern "C" long long afcm_launch_count(void) { return afcm::g_launches.load(); }

extern "C" int afcm_device_check(void)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { afcm::set_error("no CUDA device: %s", cudaGetErrorString(e)); return (int)e; }
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        afcm::set_error("libafcm_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
        return AFCM_ERR_UNSUPPORTED;
    }
    return AFCM_OK;
}
