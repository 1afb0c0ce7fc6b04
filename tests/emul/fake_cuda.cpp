/*
 * tests/emul/fake_cuda.cpp -- TEST INFRASTRUCTURE: the handful of CUDA runtime entry points
 * ffgpu_api.cu calls, on plain host memory and without any asynchrony, so that the HOST side
 * of the library (handles, launch groups, routing over several devices, staging, packet
 * arenas, damage bookkeeping, error paths) can run in a container that has no GPU.  It is
 * linked -- together with fake_kernels.cpp -- only into tests/emul/cpu/libffgpu.so, which
 * only tests/test_host_pipeline_cpu.py loads.  The product links the real runtime and has no
 * CPU path.
 *
 *   FAKE_CUDA_DEVICES=n      devices cudaGetDeviceCount reports (default 1)
 *   FAKE_CUDA_FAIL_ALLOC=k   the k-th allocation from now on (cudaMalloc / cudaHostAlloc,
 *                            1-based, counted per process) fails with cudaErrorMemoryAllocation
 */
#include <cuda_runtime_api.h>

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>

namespace {
std::mutex g_lock;
struct Block { size_t bytes; int kind; };            /* kind 1: device, 2: pinned host */
std::map<uintptr_t, Block> g_blocks;
long g_allocs;
thread_local int t_device;
thread_local cudaError_t t_last = cudaSuccess;

cudaError_t set(cudaError_t e) { if (e != cudaSuccess) t_last = e; return e; }

long fail_at()
{
    const char *v = getenv("FAKE_CUDA_FAIL_ALLOC");
    return v && *v ? atol(v) : 0;
}

cudaError_t alloc(void **p, size_t bytes, int kind)
{
    std::lock_guard<std::mutex> g(g_lock);
    const long k = fail_at();
    *p = nullptr;
    if (k > 0 && ++g_allocs == k)
        return set(cudaErrorMemoryAllocation);
    /* 256-byte alignment like the real allocators; contents deliberately not zero (the big
     * arenas -- a handle reserves gigabytes it may never touch -- stay untouched pages) */
    void *q = nullptr;
    if (posix_memalign(&q, 256, bytes ? bytes : 1))
        return set(cudaErrorMemoryAllocation);
    memset(q, 0xA5, bytes < ((size_t)4 << 20) ? bytes : (size_t)4 << 20);
    g_blocks[(uintptr_t)q] = Block{ bytes, kind };
    *p = q;
    return cudaSuccess;
}

cudaError_t release(void *p, int kind)
{
    if (!p)
        return cudaSuccess;
    std::lock_guard<std::mutex> g(g_lock);
    auto it = g_blocks.find((uintptr_t)p);
    if (it == g_blocks.end() || it->second.kind != kind)
        return set(cudaErrorInvalidValue);          /* a free of something never allocated */
    g_blocks.erase(it);
    free(p);
    return cudaSuccess;
}

struct FakeEvent { int recorded; };
}  // namespace

extern "C" long fake_cuda_live_blocks(void)
{
    std::lock_guard<std::mutex> g(g_lock);
    return (long)g_blocks.size();
}
extern "C" void fake_cuda_reset_alloc_counter(void)
{
    std::lock_guard<std::mutex> g(g_lock);
    g_allocs = 0;
}

cudaError_t cudaGetDeviceCount(int *n)
{
    const char *v = getenv("FAKE_CUDA_DEVICES");
    *n = v && *v ? atoi(v) : 1;
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int d)
{
    int n = 0;
    cudaGetDeviceCount(&n);
    if (d < 0 || d >= n)
        return set(cudaErrorInvalidDevice);
    t_device = d;
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { cudaError_t e = t_last; t_last = cudaSuccess; return e; }
const char *cudaGetErrorString(cudaError_t e)
{
    switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorMemoryAllocation: return "out of memory";
    case cudaErrorInvalidValue: return "invalid argument";
    case cudaErrorInvalidDevice: return "invalid device ordinal";
    default: return "fake CUDA error";
    }
}

cudaError_t cudaMalloc(void **p, size_t bytes) { return alloc(p, bytes, 1); }
cudaError_t cudaFree(void *p) { return release(p, 1); }
cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned) { return alloc(p, bytes, 2); }
cudaError_t cudaFreeHost(void *p) { return release(p, 2); }

cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) { memcpy(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t)
{
    memmove(dst, src, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind, cudaStream_t)
{
    if (width > dpitch || width > spitch)
        return set(cudaErrorInvalidPitchValue);
    for (size_t y = 0; y < height; y++)
        memcpy((uint8_t *)dst + y * dpitch, (const uint8_t *)src + y * spitch, width);
    return cudaSuccess;
}
cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned)
{
    *s = (cudaStream_t)malloc(8);
    return *s ? cudaSuccess : set(cudaErrorMemoryAllocation);
}
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }

cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned)
{
    *e = (cudaEvent_t)calloc(1, sizeof(FakeEvent));
    return *e ? cudaSuccess : set(cudaErrorMemoryAllocation);
}
cudaError_t cudaEventCreate(cudaEvent_t *e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { if (e) ((FakeEvent *)e)->recorded = 1; return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.001f; return cudaSuccess; }

cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p)
{
    std::lock_guard<std::mutex> g(g_lock);
    memset(a, 0, sizeof(*a));
    a->type = cudaMemoryTypeUnregistered;
    auto it = g_blocks.upper_bound((uintptr_t)p);
    if (it != g_blocks.begin()) {
        --it;
        if ((uintptr_t)p < it->first + it->second.bytes)
            a->type = it->second.kind == 1 ? cudaMemoryTypeDevice : cudaMemoryTypeHost;
    }
    return cudaSuccess;
}

cudaError_t cudaGetDriverEntryPoint(const char *, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *st)
{
    *fn = nullptr;
    if (st)
        *st = cudaDriverEntryPointSymbolNotFound;
    return cudaSuccess;
}
