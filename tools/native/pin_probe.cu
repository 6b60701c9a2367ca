// Which kind of host memory gives the full PCIe rate for H2D copies?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
int main() {
  const size_t n = 6400000;
  char* d; CK(cudaMalloc(&d, n));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto rate = [&](const char* name, char* h) {
    memset(h, 1, n);
    for (int i = 0; i < 5; ++i) CK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < 50; ++i) CK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double up = n * 50.0 / (ms * 1e-3) / 1e9;
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < 50; ++i) CK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-50s H2D %6.1f GB/s   D2H %6.1f GB/s\n", name, up, n * 50.0 / (ms * 1e-3) / 1e9);
  };
  char* h;
  CK(cudaHostAlloc(&h, n, cudaHostAllocDefault)); rate("cudaHostAlloc default", h); CK(cudaFreeHost(h));
  CK(cudaHostAlloc(&h, n, cudaHostAllocMapped)); rate("cudaHostAlloc mapped", h); CK(cudaFreeHost(h));
  CK(cudaHostAlloc(&h, n, cudaHostAllocPortable)); rate("cudaHostAlloc portable", h); CK(cudaFreeHost(h));
  CK(cudaHostAlloc(&h, n, cudaHostAllocWriteCombined)); rate("cudaHostAlloc write-combined", h); CK(cudaFreeHost(h));
  h = (char*)aligned_alloc(4096, n); memset(h, 0, n); rate("malloc (pageable)", h);
  CK(cudaHostRegister(h, n, cudaHostRegisterDefault)); rate("malloc + cudaHostRegister default", h); CK(cudaHostUnregister(h));
  CK(cudaHostRegister(h, n, cudaHostRegisterMapped)); rate("malloc + cudaHostRegister mapped", h); CK(cudaHostUnregister(h));
  CK(cudaHostRegister(h, n, cudaHostRegisterPortable)); rate("malloc + cudaHostRegister portable", h); CK(cudaHostUnregister(h));
  // a large allocation, copy from an offset (what a caching host allocator hands out)
  CK(cudaHostAlloc(&h, 64 * n, cudaHostAllocDefault)); rate("cudaHostAlloc default, 410 MB block, offset 0", h); rate("   same block, offset 200 MB", h + 31 * n + 512); CK(cudaFreeHost(h));
  return 0;
}
