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

void launch_inverse_permutation(const int* ids, uint32_t* inv, long long n, int blocks, cudaStream_t stream) {
  inverse_permutation_kernel<<<blocks, 256, 0, stream>>>(ids, inv, n);
}

void launch_grid_index(const psb_grid_desc& g, const double* x, long long n, uint32_t* out) {
  const size_t cnt = (size_t)n * g.ndim;
  grid_index_kernel<<<(unsigned)((cnt + 255) / 256), 256>>>(g, x, n, out);
}

}  // namespace psb
