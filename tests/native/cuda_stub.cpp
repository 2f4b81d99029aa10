// A CUDA runtime made of host memory, for sanitizer runs of the library's HOST logic on a machine without a GPU (tests only):
// device memory is malloc'ed (so ASan sees every upload and copy), streams and events are opaque tokens, kernel launches do nothing.
// The nvcc-generated registration calls are answered too, so the test program links without libcudart.
#include <cuda_runtime_api.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

static std::mutex g_mu;
static std::map<void*, size_t> g_blocks;
static size_t g_in_use = 0;
size_t stub_device_bytes = (size_t)1 << 30;       // "device" capacity (tests may change it)
size_t stub_allocs = 0, stub_frees = 0, stub_launches = 0, stub_stream_waits = 0, stub_device_syncs = 0;
size_t stub_bytes_in_use() { return g_in_use; }

extern "C" {
cudaError_t cudaMalloc(void** p, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_in_use + bytes > stub_device_bytes) { *p = nullptr; return cudaErrorMemoryAllocation; }
    *p = malloc(bytes ? bytes : 1);
    if (!*p) return cudaErrorMemoryAllocation;
    g_blocks[*p] = bytes; g_in_use += bytes; stub_allocs++;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_blocks.find(p);
    if (it == g_blocks.end()) return cudaErrorInvalidValue;
    g_in_use -= it->second; g_blocks.erase(it); free(p); stub_frees++;
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemGetInfo(size_t* fr, size_t* tot) { *tot = stub_device_bytes; *fr = stub_device_bytes - g_in_use; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p);
    strcpy(p->name, "host stub"); p->multiProcessorCount = 148; p->major = 10; p->minor = 0;
    p->sharedMemPerBlockOptin = 227 * 1024; p->sharedMemPerMultiprocessor = 228 * 1024; p->totalGlobalMem = stub_device_bytes;
    return cudaSuccess;
}
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "stub error"; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { stub_device_syncs++; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { stub_stream_waits++; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(1); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)malloc(1); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaLaunchKernel(const void*, dim3, dim3, void**, size_t, cudaStream_t) { stub_launches++; return cudaSuccess; }
// what nvcc's host stubs call
void** __cudaRegisterFatBinary(void*) { static void* h; return &h; }
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
void __cudaRegisterFunction(void**, const char*, char*, const char*, int, uint3*, uint3*, dim3*, dim3*, int*) {}
char __cudaInitModule(void**) { return 0; }
void __cudaRegisterVar(void**, char*, char*, const char*, int, size_t, int, int) {}
unsigned __cudaPushCallConfiguration(dim3, dim3, size_t, void*) { return 0; }
cudaError_t __cudaPopCallConfiguration(dim3*, dim3*, size_t*, void*) { return cudaSuccess; }
}
