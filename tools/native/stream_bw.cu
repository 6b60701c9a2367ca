// Ceiling probe for the slot-ordered streams: how fast can one launch read positions (N,4) f32 + images (N,3) i32
// (28 B / slot) with (a) per-warp TMA bulk rings, (b) register-staged LDG, and read-modify-write force rows (N,4)?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/native/stream_bw.cu -o tools/bin/stream_bw
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" :: "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void g2s(void* d, const void* s, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void s2g(void* d, const void* s, uint32_t n) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(d), "r"(s32(s)), "r"(n) : "memory");
}

// (a) per-warp TMA ring: CH slots per chunk, S stages
template <int CH, int S>
__global__ void __launch_bounds__(256, 2) tma_read(const float4* pos, const int* img, long long n, float* out) {
  extern __shared__ __align__(128) unsigned char ring_all[];
  __shared__ uint64_t bars[8 * S];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t SB = CH * 28;
  unsigned char* ring = ring_all + (size_t)warp * S * SB;
  uint64_t* full = bars + warp * S;
  const long long nw = (long long)gridDim.x * 8;
  long long run = (n + nw - 1) / nw;
  run = (run + CH - 1) / CH * CH;
  const long long lo = min(n, ((long long)blockIdx.x * 8 + warp) * run), hi = min(n, lo + run);
  const int nch = (int)((hi - lo + CH - 1) / CH);
  if (lane == 0) { for (int s = 0; s < S; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  auto issue = [&](int c) {
    if (lane == 0 && c < nch) {
      const long long base = lo + (long long)c * CH;
      const uint32_t cnt = (uint32_t)min((long long)CH, hi - base);
      unsigned char* st = ring + (size_t)(c % S) * SB;
      mbar_expect(&full[c % S], cnt * 28);
      g2s(st, pos + base, cnt * 16, &full[c % S]);
      g2s(st + CH * 16, img + 3 * base, cnt * 12, &full[c % S]);
    }
  };
  for (int c = 0; c < S; ++c) issue(c);
  float acc = 0.f;
  for (int c = 0; c < nch; ++c) {
    mbar_wait(&full[c % S], (c / S) & 1);
    const unsigned char* st = ring + (size_t)(c % S) * SB;
    for (int r = lane; r < CH; r += 32) {
      const float4 p = reinterpret_cast<const float4*>(st)[r];
      const int* q = reinterpret_cast<const int*>(st + CH * 16) + 3 * r;
      acc += p.x + p.y + p.z + (float)(q[0] + q[1] + q[2]);
    }
    __syncwarp();
    issue(c + S);
  }
  if (acc == 123.456f) out[0] = acc;
}

// (b) register-staged loads, U rows per thread in flight
template <int U>
__global__ void __launch_bounds__(256, 2) ldg_read(const float4* pos, const int* img, long long n, float* out) {
  long long chunk = (n + gridDim.x - 1) / gridDim.x;
  const long long lo = min(n, (long long)blockIdx.x * chunk), hi = min(n, lo + chunk);
  float acc = 0.f;
  for (long long base = lo + threadIdx.x; base < hi; base += U * 256) {
    float4 p[U]; int a[U], b[U], c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = min(base + u * 256, hi - 1);
      p[u] = __ldg(pos + i); a[u] = __ldg(img + 3 * i); b[u] = __ldg(img + 3 * i + 1); c[u] = __ldg(img + 3 * i + 2);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += p[u].x + p[u].y + p[u].z + (float)(a[u] + b[u] + c[u]);
  }
  if (acc == 123.456f) out[0] = acc;
}

// (c) force rows: bulk load -> add in shared memory -> bulk store
template <int CH, int S>
__global__ void __launch_bounds__(256, 2) tma_rmw(float4* f, long long n) {
  extern __shared__ __align__(128) unsigned char ring_all[];
  __shared__ uint64_t bars[8 * S];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t SB = CH * 16;
  unsigned char* ring = ring_all + (size_t)warp * S * SB;
  uint64_t* full = bars + warp * S;
  const long long nw = (long long)gridDim.x * 8;
  long long run = (n + nw - 1) / nw;
  run = (run + CH - 1) / CH * CH;
  const long long lo = min(n, ((long long)blockIdx.x * 8 + warp) * run), hi = min(n, lo + run);
  const int nch = (int)((hi - lo + CH - 1) / CH);
  if (lane == 0) { for (int s = 0; s < S; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  auto issue = [&](int c) {
    if (lane == 0 && c < nch) {
      const long long base = lo + (long long)c * CH;
      const uint32_t cnt = (uint32_t)min((long long)CH, hi - base);
      mbar_expect(&full[c % S], cnt * 16);
      g2s(ring + (size_t)(c % S) * SB, f + base, cnt * 16, &full[c % S]);
    }
  };
  for (int c = 0; c < S; ++c) issue(c);
  for (int c = 0; c < nch; ++c) {
    mbar_wait(&full[c % S], (c / S) & 1);
    unsigned char* st = ring + (size_t)(c % S) * SB;
    const long long base = lo + (long long)c * CH;
    const uint32_t cnt = (uint32_t)min((long long)CH, hi - base);
    for (int r = lane; r < CH; r += 32) {
      float4* q = reinterpret_cast<float4*>(st) + r;
      float4 v = *q; v.x += 1.f; v.y += 1.f; v.z += 1.f; *q = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      s2g(f + base, st, cnt * 16);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
    if (c >= 1) issue(c - 1 + S);
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__global__ void __launch_bounds__(256, 2) ldg_rmw(float4* f, long long n) {
  long long chunk = (n + gridDim.x - 1) / gridDim.x;
  const long long lo = min(n, (long long)blockIdx.x * chunk), hi = min(n, lo + chunk);
  for (long long base = lo + threadIdx.x; base < hi; base += 8 * 256) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = f[min(base + u * 256, hi - 1)];
#pragma unroll
    for (int u = 0; u < 8; ++u) if (base + u * 256 < hi) { v[u].x += 1.f; v[u].y += 1.f; v[u].z += 1.f; f[base + u * 256] = v[u]; }
  }
}

int main() {
  const long long n = 1000000;
  const int NB = 8;
  float4* pos[NB]; int* img[NB]; float4* frc[NB]; float* out;
  for (int i = 0; i < NB; ++i) { CK(cudaMalloc(&pos[i], n * 16)); CK(cudaMalloc(&img[i], n * 12)); CK(cudaMalloc(&frc[i], n * 16));
    CK(cudaMemset(pos[i], 0, n * 16)); CK(cudaMemset(img[i], 0, n * 12)); CK(cudaMemset(frc[i], 0, n * 16)); }
  CK(cudaMalloc(&out, 4));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, double bytes, auto launch) {
    for (int i = 0; i < NB; ++i) launch(i);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 40;
    for (int r = 0; r < reps; ++r) launch(r % NB);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-64s %7.2f us/launch  %7.1f GB/s\n", name, ms * 1e3 / reps, bytes / (ms * 1e-3 / reps) / 1e9);
  };
#define TMA_READ(CH, S) { CK(cudaFuncSetAttribute(tma_read<CH, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * S * CH * 28)); \
    char nm[96]; snprintf(nm, 96, "TMA ring read pos+img: chunk %d slots x %d stages/warp, 2 CTA/SM", CH, S); \
    timeit(nm, n * 28.0, [&](int i) { tma_read<CH, S><<<2 * sms, 256, 8 * S * CH * 28>>>(pos[i], img[i], n, out); }); }
  TMA_READ(128, 2) TMA_READ(128, 3) TMA_READ(64, 4) TMA_READ(64, 6) TMA_READ(32, 8) TMA_READ(256, 1)
  timeit("LDG read pos+img, 8 rows/thread in flight, 2 CTA/SM", n * 28.0, [&](int i) { ldg_read<8><<<2 * sms, 256>>>(pos[i], img[i], n, out); });
  timeit("LDG read pos+img, 4 rows/thread in flight, 4x SM CTAs", n * 28.0, [&](int i) { ldg_read<4><<<4 * sms, 256>>>(pos[i], img[i], n, out); });
  timeit("LDG read pos+img, 2 rows/thread, 16x SM CTAs", n * 28.0, [&](int i) { ldg_read<2><<<16 * sms, 256>>>(pos[i], img[i], n, out); });
#define TMA_RMW(CH, S) { CK(cudaFuncSetAttribute(tma_rmw<CH, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * S * CH * 16)); \
    char nm[96]; snprintf(nm, 96, "TMA ring RMW forces: chunk %d slots x %d stages/warp, 2 CTA/SM", CH, S); \
    timeit(nm, n * 32.0, [&](int i) { tma_rmw<CH, S><<<2 * sms, 256, 8 * S * CH * 16>>>(frc[i], n); }); }
  TMA_RMW(128, 3) TMA_RMW(128, 6) TMA_RMW(256, 3)
  timeit("LDG/STG RMW forces, 8 rows/thread, 2 CTA/SM", n * 32.0, [&](int i) { ldg_rmw<<<2 * sms, 256>>>(frc[i], n); });
  timeit("LDG/STG RMW forces, 8 rows/thread, 8x SM CTAs", n * 32.0, [&](int i) { ldg_rmw<<<8 * sms, 256>>>(frc[i], n); });
  return 0;
}
