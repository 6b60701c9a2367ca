// pairwise coordination kernels (rows per lane R, exponent n = 2H) and small utility kernels
#include "psb_launch.cuh"
#include "psb_pairs.cuh"
#include "psb_step.cuh"
namespace psb {

// OpenMM-CUDA: inverse of the slot -> atom permutation (openmm.py:106-107 `ids.argsort()`)
__global__ void inverse_permutation_kernel(const int* __restrict__ ids, uint32_t* __restrict__ inv,
                                           long long n) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n;
       s += (long long)gridDim.x * blockDim.x) {
    const int a = __ldg(ids + s);
    if (a >= 0 && a < n) inv[a] = (uint32_t)s;
  }
}

// grids.py:109-158 for a batch of points (parity tests for the bit-exact bins)
__global__ void grid_index_kernel(psb_grid_desc g, const double* __restrict__ x, long long n,
                                  uint32_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * g.ndim) return;
  const int d = (int)(i % g.ndim);
  out[i] = grid_index_1d(g.kind, x[i], g.lower[d], g.upper[d], g.shape[d]);
}

template <int R, int H>
static void pair_launch(int grid, cudaStream_t stream, const double* xa, int na_pad, const double* xb,
                        int nb_pad, int nA, int nB, int nrt, double inv_r02, double* gA, double* gB,
                        double* xi_part) {
  coord_pairs_kernel<R, H><<<grid, PAIR_BLOCK, 0, stream>>>(xa, na_pad, xb, nb_pad, nA, nB, nrt, 0, inv_r02,
                                                            gA, gB, xi_part);
}

PairLaunchFn pair_launcher(int R, int H) {
  if (R == 1 && H == 1) return pair_launch<1, 1>;
  if (R == 1 && H == 2) return pair_launch<1, 2>;
  if (R == 1 && H == 3) return pair_launch<1, 3>;
  if (R == 1 && H == 4) return pair_launch<1, 4>;
  if (R == 1 && H == 6) return pair_launch<1, 6>;
  if (R == 4 && H == 1) return pair_launch<4, 1>;
  if (R == 4 && H == 2) return pair_launch<4, 2>;
  if (R == 4 && H == 3) return pair_launch<4, 3>;
  if (R == 4 && H == 4) return pair_launch<4, 4>;
  if (R == 4 && H == 6) return pair_launch<4, 6>;
  return nullptr;
}

// Launch-latency floor: kernels that do nothing (optionally one grid-wide arrive/wait), with a small
// or a Model-sized parameter block, launched plainly or cooperatively.
struct BigParams { unsigned long long w[360]; };
__global__ void __launch_bounds__(BLOCK) floor_kernel_small(unsigned long long* counter, int use_barrier,
                                                             unsigned long long launch_index) {
  if (use_barrier && threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1ULL);
    const unsigned long long want = (launch_index + 1ULL) * gridDim.x;
    while (*((volatile unsigned long long*)counter) < want) {}
  }
}
__global__ void __launch_bounds__(BLOCK) floor_kernel_big(const __grid_constant__ BigParams P, unsigned long long* counter) {
  if (P.w[threadIdx.x] == 0x123456789ULL) *counter = 1;  // keep the parameter block alive
}

// Dependent-issue latency of the operations the single-warp finalizers are made of (cycles / op).
__global__ void latency_probe_kernel(double* out, double seed) {
  __shared__ double sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = (double)((threadIdx.x * 7 + 3) % 64);
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int N = 256;
  double x = seed, y = 1.0 + seed * 1e-9;
  float xf = (float)seed;
  int xi = (int)seed + 3;
  long long t0, t1;
#define PROBE(slot, STMT)                                    \
  t0 = clock64();                                            \
  _Pragma("unroll 16") for (int i = 0; i < N; ++i) { STMT; } \
  t1 = clock64();                                            \
  if (threadIdx.x == 0) out[slot] = (double)(t1 - t0) / N;
  PROBE(0, x = fma(x, y, 1e-30))
  PROBE(1, x = x + y)
  PROBE(2, xf = fmaf(xf, 1.0000001f, 1e-30f))
  PROBE(3, xi = xi * 3 + 1)
  PROBE(4, x = sm[((int)x) & 63])
  PROBE(5, x = sqrt(x + 2.0))
  PROBE(6, x = 1.0 / (x + 2.0))
  PROBE(7, x = rsqrt(x + 2.0))
  PROBE(8, x = floor(x * 1.5) + 0.25)
  PROBE(9, x = fmod(x + 37.5, 0.7))
  PROBE(10, x = atan2(x + 0.3, y))
  PROBE(11, x = exp(-x * 1e-3))
#undef PROBE
  if (threadIdx.x == 0) out[15] = x + (double)xf + (double)xi;
}

cudaError_t launch_latency_probe(double* out_dev, cudaStream_t stream) {
  latency_probe_kernel<<<1, 64, 0, stream>>>(out_dev, 1.5);
  latency_probe_kernel<<<1, 64, 0, stream>>>(out_dev, 1.5);
  return cudaGetLastError();
}

// Sustained issue throughput of the arithmetic the streaming and pair kernels are made of: every thread runs
// 8 independent chains (enough to cover the pipe latency at 8+ warps per scheduler), the whole device is filled.
// which: 0 DFMA, 1 FFMA, 2 F2F.F64.F32 (float -> double), 3 I2F.F64 (int -> double), 4 DADD
template <int WHICH>
__global__ void __launch_bounds__(256) throughput_kernel(double* out, int iters, double seed) {
  double x[8];
  float xf[8];
  int xi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    x[j] = seed + j + threadIdx.x * 1e-3;
    xf[j] = (float)x[j];
    xi[j] = (int)threadIdx.x + j;
  }
  const double y = 1.0 + seed * 1e-9;
  const float yf = (float)y;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (WHICH == 0) x[j] = fma(x[j], y, 1e-30);
      if (WHICH == 1) xf[j] = fmaf(xf[j], yf, 1e-30f);
      if (WHICH == 2) { x[j] = (double)xf[j]; xf[j] = __double2float_rn(x[j]) + yf; }  // one F2F up, one down, one FADD
      if (WHICH == 3) { x[j] = (double)xi[j]; xi[j] = __double2int_rn(x[j]) ^ i; }      // one I2F, one F2I, one LOP
      if (WHICH == 4) x[j] = x[j] + y;
    }
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j] + (double)xf[j] + (double)xi[j];
  if (s == 123.456) out[0] = s;  // keeps the chains alive
}

// -> out_host[5]: operations per second (DFMA, FFMA, F2F.F64.F32, I2F.F64, DADD), whole device
cudaError_t measure_throughput(double* out_host, int num_sms) {
  double* d = nullptr;
  cudaError_t e = cudaMalloc(&d, 8);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = num_sms * 8, iters = 4096;
  for (int which = 0; which < 5 && e == cudaSuccess; ++which) {
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      switch (which) {
        case 0: throughput_kernel<0><<<grid, 256>>>(d, iters, 1.5); break;
        case 1: throughput_kernel<1><<<grid, 256>>>(d, iters, 1.5); break;
        case 2: throughput_kernel<2><<<grid, 256>>>(d, iters, 1.5); break;
        case 3: throughput_kernel<3><<<grid, 256>>>(d, iters, 1.5); break;
        default: throughput_kernel<4><<<grid, 256>>>(d, iters, 1.5); break;
      }
      cudaEventRecord(e1);
      e = cudaEventSynchronize(e1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    out_host[which] = (double)grid * 256.0 * 8.0 * iters / (best * 1e-3);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return e;
}

cudaError_t launch_floor(int grid, int cooperative, int big_params, int use_barrier, long long n,
                         unsigned long long* counter, cudaStream_t stream) {
  BigParams P{};
  cudaError_t e = cudaMemsetAsync(counter, 0, 8, stream);
  for (long long i = 0; i < n && e == cudaSuccess; ++i) {
    unsigned long long li = (unsigned long long)i;
    int ub = use_barrier;
    void* args_small[] = {(void*)&counter, (void*)&ub, (void*)&li};
    void* args_big[] = {(void*)&P, (void*)&counter};
    const void* fn = big_params ? (const void*)floor_kernel_big : (const void*)floor_kernel_small;
    void** args = big_params ? args_big : args_small;
    if (cooperative) e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(BLOCK), args, 0, stream);
    else e = cudaLaunchKernel(fn, dim3(grid), dim3(BLOCK), args, 0, stream);
  }
  return e;
}

void launch_inverse_permutation(const int* ids, uint32_t* inv, long long n, int blocks, cudaStream_t stream) {
  inverse_permutation_kernel<<<blocks, 256, 0, stream>>>(ids, inv, n);
}

void launch_grid_index(const psb_grid_desc& g, const double* x, long long n, uint32_t* out) {
  const size_t cnt = (size_t)n * g.ndim;
  grid_index_kernel<<<(unsigned)((cnt + 255) / 256), 256>>>(g, x, n, out);
}

}  // namespace psb
