// psb_ermsd.cuh -- eRMSD collective variable (colvars/orientation.py:129-378, Bottaro et al., NAR 2014).
//
// n nucleotides, three ring atoms each (C2, C4/C6 in the reference's order): a local frame per base, an
// n x n table of 4-vectors G(r~_ij) of the position of base i seen from the frame of base j, and
//   eRMSD = sqrt( sum_ij |G_ij - Gref_ij|^2 / n ).
// The reference differentiates this with autodiff over a dense (n, n, 4) tensor.  Here ONE CTA does the whole
// CV (n <= ERMSD_MAX_N bases: RNA motifs; the step kernel that follows is the general class): gather + unwrap the
// 3n atoms into shared memory, thread j builds frame j, loops over i for the value (block sum), then loops again
// for the adjoints -- pair (i, j) feeds frame j and origin j, pair (j, i) feeds origin j with the opposite sign,
// both evaluated by thread j, so nothing is accumulated atomically -- and back-propagates through its own
// frame (normalisations and cross products written out).  Output: xi_part[0] and d xi / d r per ring atom in the
// CV's gradient workspace, which the step kernel consumes exactly like the pair kernel's (PSB_CV_COORDINATION).
#pragma once

#include "psb_device.cuh"
#include "psb_math.cuh"

namespace psb {

constexpr int ERMSD_MAX_N = 256;  // bases per CV (threads of the CTA)
constexpr int ERMSD_BLOCK = 256;

struct ErmsdFrame {
  V3 o, xh, yh, zh;
};

// orientation.py:177-216: x = cog -> site 0, z = x^ x (cog -> site 1), y = z^ x x^ -- and its adjoint.
struct FrameAux {
  double nu, nw, nv;  // |u|, |w|, |v| before the normalisations
  V3 cvec;            // cog -> site 1
};
__device__ __forceinline__ ErmsdFrame ermsd_make_frame(V3 r0, V3 r1, V3 r2, FrameAux* aux) {
  ErmsdFrame f;
  f.o = (1.0 / 3.0) * (r0 + r1 + r2);
  const V3 u = r0 - f.o;
  aux->nu = norm(u);
  f.xh = (1.0 / aux->nu) * u;
  aux->cvec = r1 - f.o;
  const V3 w = cross(f.xh, aux->cvec);
  aux->nw = norm(w);
  f.zh = (1.0 / aux->nw) * w;
  const V3 v = cross(f.zh, f.xh);
  aux->nv = norm(v);
  f.yh = (1.0 / aux->nv) * v;
  return f;
}
// adjoints of the axes (gxh, gyh, gzh) and of the origin (go) -> the three sites
__device__ __forceinline__ void ermsd_frame_backward(const ErmsdFrame& f, const FrameAux& aux, V3 gxh, V3 gyh, V3 gzh,
                                                     V3 go, V3* g0, V3* g1, V3* g2) {
  // y^ = v / |v|, v = z^ x x^; z^ = w / |w|, w = x^ x c; x^ = u / |u|; u = r0 - o, c = r1 - o;  d(a x b): g_a = b x g, g_b = g x a
  const V3 gv = (1.0 / aux.nv) * (gyh - dot(f.yh, gyh) * f.yh);
  gzh = gzh + cross(f.xh, gv);
  gxh = gxh + cross(gv, f.zh);
  const V3 gw = (1.0 / aux.nw) * (gzh - dot(f.zh, gzh) * f.zh);
  gxh = gxh + cross(aux.cvec, gw);
  const V3 gc = cross(gw, f.xh);
  const V3 gu = (1.0 / aux.nu) * (gxh - dot(f.xh, gxh) * f.xh);
  const V3 gshare = (1.0 / 3.0) * (go - gu - gc);  // every site's share through the centre of geometry
  *g0 = gshare + gu;
  *g1 = gshare + gc;
  *g2 = gshare;
}

// G(r~) of orientation.py:219-288 for p = o_i - o_j in frame F_j; returns |G - Gref|^2 contribution pieces.
// out: G[4]; aux for the adjoint: rt (r~), rho, inside (rho < cutoff)
__device__ __forceinline__ void ermsd_g(const ErmsdFrame& Fj, V3 p, double inv_a, double inv_b, double cutoff,
                                        double gamma, double* G, V3* rt, double* rho, bool* inside) {
  const V3 r = v3(dot(p, Fj.xh), dot(p, Fj.yh), dot(p, Fj.zh));
  *rt = v3(r.x * inv_a, r.y * inv_a, r.z * inv_b);
  *rho = norm(*rt);
  *inside = cutoff - *rho > 0.0;  // np.heaviside(cutoff - rho, 0)
  if (*inside) {
    const double inv = *rho != 0.0 ? 1.0 / *rho : 0.0;
    double sn, cs;
    sincos(gamma * *rho, &sn, &cs);
    const double s = sn * inv / gamma;
    G[0] = s * rt->x; G[1] = s * rt->y; G[2] = s * rt->z; G[3] = (1.0 + cs) / gamma;
  } else {
    G[0] = G[1] = G[2] = G[3] = 0.0;
  }
}

// adjoint of ermsd_g: gG (4) -> gradient with respect to p (returned) and to the frame axes of j (accumulated)
__device__ __forceinline__ V3 ermsd_g_adjoint(const ErmsdFrame& Fj, V3 p, V3 rt, double rho, double inv_a,
                                              double inv_b, double gamma, const double* gG, V3* gxh, V3* gyh,
                                              V3* gzh) {
  // G123 = s(rho) r~ / gamma with s = sin(gamma rho) / rho;  G4 = (1 + cos(gamma rho)) / gamma
  double sn, cs;
  sincos(gamma * rho, &sn, &cs);
  const double inv = 1.0 / rho;
  const double s = sn * inv / gamma;
  const double ds = (gamma * cs * inv - sn * inv * inv) / gamma;  // d s / d rho
  const double g_dot_rt = gG[0] * rt.x + gG[1] * rt.y + gG[2] * rt.z;
  const double radial = (ds * g_dot_rt - gG[3] * sn) * inv;  // multiplies r~ (d rho / d r~ = r~ / rho)
  const V3 grt = v3(s * gG[0] + radial * rt.x, s * gG[1] + radial * rt.y, s * gG[2] + radial * rt.z);
  const V3 gr = v3(grt.x * inv_a, grt.y * inv_a, grt.z * inv_b);
  if (gxh) {
    *gxh = *gxh + gr.x * p;
    *gyh = *gyh + gr.y * p;
    *gzh = *gzh + gr.z * p;
  }
  return gr.x * Fj.xh + gr.y * Fj.yh + gr.z * Fj.zh;
}

template <int LAYOUT, typename Real>
__global__ void __launch_bounds__(ERMSD_BLOCK)
ermsd_kernel(const Model* __restrict__ Mdev, const __grid_constant__ Snap S, int c) {
  using A = Access<LAYOUT, Real>;
  const Model& M = *Mdev;  // the handle's device copy (a by-value parameter would exceed 4 KB)
  __shared__ double P[3 * ERMSD_MAX_N][3];
  __shared__ ErmsdFrame F[ERMSD_MAX_N];
  __shared__ double red[ERMSD_BLOCK / 32];
  __shared__ double s_value;
  const CvDev& cv = M.cv[c];
  const int n = (int)(cv.t1 - cv.t0) / 3;
  const int tid = threadIdx.x;
  const double cutoff = cv.r0, inv_a = 1.0 / cv.sumQ[0], inv_b = 1.0 / cv.sumQ[1];
  const double gamma = 3.141592653589793 / cutoff;
  const double* Gref = cv.ref;  // (n, n, 4)

  for (int t = tid; t < 3 * n; t += ERMSD_BLOCK) {
    const uint32_t term = cv.t0 + t;
    const uint32_t slot = A::slot(S, M, __ldg(M.term_tag + term));
    M.term_slot[term] = slot;
    const V3 r = A::position(S, slot);
    P[t][0] = r.x; P[t][1] = r.y; P[t][2] = r.z;
  }
  __syncthreads();
  // local reference systems (orientation.py:177-216).  Coarse-grained variant (ERMSDCG, orientation.py:380-560): the
  // three sites first define a frame of their own in which the ideal C2 / C4 / C6 positions of the base are known
  // (local_coordinates[sequence]); the eRMSD frame is then built on those reconstructed atoms.
  const bool cg = cv.axis == 1;
  FrameAux aux{}, aux_cg{};
  ErmsdFrame f_cg{};
  double lc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (tid < n) {
    V3 r0 = v3(P[3 * tid][0], P[3 * tid][1], P[3 * tid][2]);
    V3 r1 = v3(P[3 * tid + 1][0], P[3 * tid + 1][1], P[3 * tid + 1][2]);
    V3 r2 = v3(P[3 * tid + 2][0], P[3 * tid + 2][1], P[3 * tid + 2][2]);
    if (cg) {
      f_cg = ermsd_make_frame(r0, r1, r2, &aux_cg);
      const int base = (int)__ldg(cv.w + 36 + tid);
      for (int q = 0; q < 9; ++q) lc[q] = __ldg(cv.w + 9 * base + q);
      r0 = f_cg.o + lc[0] * f_cg.xh + lc[1] * f_cg.yh + lc[2] * f_cg.zh;  // orientation.py infer_base_coordinates
      r1 = f_cg.o + lc[3] * f_cg.xh + lc[4] * f_cg.yh + lc[5] * f_cg.zh;
      r2 = f_cg.o + lc[6] * f_cg.xh + lc[7] * f_cg.yh + lc[8] * f_cg.zh;
    }
    F[tid] = ermsd_make_frame(r0, r1, r2, &aux);
  }
  __syncthreads();
  // value
  double sum2 = 0.0;
  if (tid < n) {
    const ErmsdFrame fj = F[tid];
    for (int i = 0; i < n; ++i) {
      if (i == tid) continue;  // the masked self entry is outside the cutoff on both sides (orientation.py:270-272)
      double G[4], rho;
      V3 rt;
      bool in;
      ermsd_g(fj, F[i].o - fj.o, inv_a, inv_b, cutoff, gamma, G, &rt, &rho, &in);
      const double* g0 = Gref + ((size_t)i * n + tid) * 4;
      for (int q = 0; q < 4; ++q) {
        const double d = G[q] - __ldg(g0 + q);
        sum2 += d * d;
      }
    }
  }
  sum2 = warp_sum(sum2);
  if ((tid & 31) == 0) red[tid >> 5] = sum2;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < ERMSD_BLOCK / 32; ++w) s += red[w];
    s_value = sqrt(s / (double)n);
    cv.xi_part[0] = s_value;
  }
  __syncthreads();
  // adjoints: d xi / d G_ij = (G_ij - Gref_ij) / (n xi)
  if (tid < n) {
    const int j = tid;
    const ErmsdFrame fj = F[j];
    const double wgt = 1.0 / ((double)n * s_value);
    V3 go = v3(0, 0, 0), gxh = v3(0, 0, 0), gyh = v3(0, 0, 0), gzh = v3(0, 0, 0);
    for (int i = 0; i < n; ++i) {
      if (i == j) continue;
      double G[4], gG[4], rho;
      V3 rt;
      bool in;
      // pair (i, j): base i in the frame of j -- feeds frame j and, through p = o_i - o_j, -g_p into o_j
      const V3 p = F[i].o - fj.o;
      ermsd_g(fj, p, inv_a, inv_b, cutoff, gamma, G, &rt, &rho, &in);
      if (in) {
        const double* g0 = Gref + ((size_t)i * n + j) * 4;
        for (int q = 0; q < 4; ++q) gG[q] = (G[q] - __ldg(g0 + q)) * wgt;
        go = go - ermsd_g_adjoint(fj, p, rt, rho, inv_a, inv_b, gamma, gG, &gxh, &gyh, &gzh);
      } else {  // G = 0 there, but the reference entry may not be: no derivative outside the cutoff
      }
      // pair (j, i): base j in the frame of i -- +g_p into o_j
      const ErmsdFrame fi = F[i];
      const V3 p2 = fj.o - fi.o;
      ermsd_g(fi, p2, inv_a, inv_b, cutoff, gamma, G, &rt, &rho, &in);
      if (in) {
        const double* g1 = Gref + ((size_t)j * n + i) * 4;
        for (int q = 0; q < 4; ++q) gG[q] = (G[q] - __ldg(g1 + q)) * wgt;
        go = go + ermsd_g_adjoint(fi, p2, rt, rho, inv_a, inv_b, gamma, gG, nullptr, nullptr, nullptr);
      }
    }
    V3 g0, g1, g2;
    ermsd_frame_backward(fj, aux, gxh, gyh, gzh, go, &g0, &g1, &g2);
    if (cg) {  // back through the reconstruction: V_m = o' + sum_l lc[m][l] axis'_l
      const V3 go2 = g0 + g1 + g2;
      const V3 gx2 = lc[0] * g0 + lc[3] * g1 + lc[6] * g2;
      const V3 gy2 = lc[1] * g0 + lc[4] * g1 + lc[7] * g2;
      const V3 gz2 = lc[2] * g0 + lc[5] * g1 + lc[8] * g2;
      ermsd_frame_backward(f_cg, aux_cg, gx2, gy2, gz2, go2, &g0, &g1, &g2);
    }
    double* out = cv.grad + 9 * (size_t)j;
    out[0] = g0.x; out[1] = g0.y; out[2] = g0.z;
    out[3] = g1.x; out[4] = g1.y; out[5] = g1.z;
    out[6] = g2.x; out[7] = g2.y; out[8] = g2.z;
  }
}

}  // namespace psb
