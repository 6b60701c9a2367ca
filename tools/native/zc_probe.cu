// PCIe probe: copy-engine transfers vs kernels that read / write mapped pinned host memory directly.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/native/zc_probe.cu -o tools/bin/zc_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void zc_read(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void zc_write(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// read-modify-write of 16-byte rows of mapped host memory, in place: coalesced (perm == nullptr) or through a permutation
__global__ void zc_rmw(float4* __restrict__ rows, const int* __restrict__ perm, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = perm ? (size_t)perm[i] : i;
    float4 v = rows[r];
    v.x += 1.0f; v.y += 1.0f; v.z += 1.0f;
    rows[r] = v;
  }
}

int main() {
  const size_t up = 6400000, down = 1600000;
  char *h_up, *h_down, *d_up, *d_down;
  CK(cudaHostAlloc(&h_up, up, cudaHostAllocMapped));
  CK(cudaHostAlloc(&h_down, down, cudaHostAllocMapped));
  CK(cudaMalloc(&d_up, up));
  CK(cudaMalloc(&d_down, down));
  memset(h_up, 1, up);
  char *m_up, *m_down;
  CK(cudaHostGetDevicePointer(&m_up, h_up, 0));
  CK(cudaHostGetDevicePointer(&m_down, h_down, 0));
  cudaStream_t s0, s1;
  CK(cudaStreamCreate(&s0));
  CK(cudaStreamCreate(&s1));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, auto fn, size_t bytes) {
    for (int i = 0; i < 5; ++i) fn();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, s0));
    const int n = 100;
    for (int i = 0; i < n; ++i) fn();
    CK(cudaEventRecord(e1, s0));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-60s %8.1f us  %6.1f GB/s\n", name, ms * 1e3 / n, bytes / (ms * 1e-3 / n) / 1e9);
  };
  timeit("H2D 6.4 MB copy engine", [&] { CK(cudaMemcpyAsync(d_up, h_up, up, cudaMemcpyHostToDevice, s0)); }, up);
  timeit("D2H 1.6 MB copy engine", [&] { CK(cudaMemcpyAsync(h_down, d_down, down, cudaMemcpyDeviceToHost, s0)); }, down);
  timeit("H2D 6.4 MB then D2H 1.6 MB, copy engine, one stream", [&] {
    CK(cudaMemcpyAsync(d_up, h_up, up, cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(h_down, d_down, down, cudaMemcpyDeviceToHost, s0)); }, up + down);
  timeit("H2D 6.4 MB as 2 halves on 2 streams", [&] {
    CK(cudaEventRecord(e1, s0)); CK(cudaStreamWaitEvent(s1, e1, 0));
    CK(cudaMemcpyAsync(d_up, h_up, up / 2, cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(d_up + up / 2, h_up + up / 2, up / 2, cudaMemcpyHostToDevice, s1));
    cudaEvent_t j; CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming)); CK(cudaEventRecord(j, s1)); CK(cudaStreamWaitEvent(s0, j, 0)); CK(cudaEventDestroy(j)); }, up);
  for (int grid : {16, 64, 148, 296, 592}) {
    char name[96];
    snprintf(name, sizeof name, "kernel reads 6.4 MB of mapped host memory, %d CTAs x 256", grid);
    timeit(name, [&] { zc_read<<<grid, 256, 0, s0>>>((const float4*)m_up, (float4*)d_up, up / 16); }, up);
  }
  for (int grid : {16, 64, 148, 296}) {
    char name[96];
    snprintf(name, sizeof name, "kernel writes 1.6 MB to mapped host memory, %d CTAs x 256", grid);
    timeit(name, [&] { zc_write<<<grid, 256, 0, s0>>>((const float4*)d_down, (float4*)m_down, down / 16); }, down);
  }
  timeit("kernel read 6.4 MB (148) then kernel write 1.6 MB (148)", [&] {
    zc_read<<<148, 256, 0, s0>>>((const float4*)m_up, (float4*)d_up, up / 16);
    zc_write<<<148, 256, 0, s0>>>((const float4*)d_down, (float4*)m_down, down / 16); }, up + down);
  {
    const size_t n = down / 16;
    int* hperm = (int*)malloc(n * sizeof(int));
    for (size_t i = 0; i < n; ++i) hperm[i] = (int)i;
    unsigned long long st = 88172645463325252ULL;
    for (size_t i = n - 1; i > 0; --i) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; size_t j = st % (i + 1); int t = hperm[i]; hperm[i] = hperm[j]; hperm[j] = t; }
    int* dperm; CK(cudaMalloc(&dperm, n * sizeof(int)));
    CK(cudaMemcpy(dperm, hperm, n * sizeof(int), cudaMemcpyHostToDevice));
    for (int grid : {148, 296}) {
      char name[112];
      snprintf(name, sizeof name, "kernel RMW 1.6 MB of mapped host rows in place, coalesced, %d CTAs", grid);
      timeit(name, [&] { zc_rmw<<<grid, 256, 0, s0>>>((float4*)m_down, nullptr, n); }, 2 * down);
      snprintf(name, sizeof name, "kernel RMW 1.6 MB of mapped host rows in place, random rows, %d CTAs", grid);
      timeit(name, [&] { zc_rmw<<<grid, 256, 0, s0>>>((float4*)m_down, dperm, n); }, 2 * down);
    }
  }
  // latency of one synchronous round trip (what psb_step_host pays per step on top of the bytes)
  {
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, s0));
    for (int i = 0; i < 100; ++i) {
      CK(cudaMemcpyAsync(d_up, h_up, up, cudaMemcpyHostToDevice, s0));
      CK(cudaMemcpyAsync(h_down, d_down, down, cudaMemcpyDeviceToHost, s0));
      CK(cudaStreamSynchronize(s0));
    }
    CK(cudaEventRecord(e1, s0));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-60s %8.1f us  %6.1f GB/s\n", "H2D + D2H + cudaStreamSynchronize per iteration", ms * 10, (up + down) / (ms * 1e-5) / 1e9);
  }
  return 0;
}
