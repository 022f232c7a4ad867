// Undulator far-field amplitude integral for sm_100a.
//
// Replaces: kernel `undulator` (xrt/backends/raycing/cl/undulator.cl:54-121) launched from
// Undulator._build_I_map_CL (xrt/backends/raycing/sources/synchr.py:2110-2193, dispatch at :2174).
// Numerics follow the numpy path Undulator._sp_sum (synchr.py:1929-2033) and _build_I_map_conv
// (synchr.py:2048-2108) -- the oracle (oracle/undulator.py) -- NOT the .cl text: the two differ in beta'_z
// (conventions paragraph).
//
// One thread owns R samples (E, theta, psi); the Clenshaw-Curtis node table (7 derived doubles per node) is
// staged through shared memory and read as broadcasts.  FP64-pipe bound: ~65 DFMA-pipe instructions per
// (sample, node), 80 bytes of HBM traffic per sample.
//
// Rounding discipline: 1 - n.beta suffers a ~1e8 cancellation (gamma^-2 scale), so the reference's own value
// of 1/(1 - n.beta) carries ~4e-9 relative rounding noise.  To agree with it to 1e-13 the chain
// beta_z -> 1 - n.beta -> n_z - beta_z is evaluated with the reference's individually rounded operations
// (__dmul_rn/__dadd_rn, no FMA contraction); everything that is not cancellation-amplified uses FMAs.
#include "common.cuh"
#include "kernels.h"
#include "undulator_core.cuh"

namespace xrtb {

// MODE 0: the reference kernel interface (gamma, wu, w, ww1, ddphi, ddpsi per sample) -> raw (Is, Ip) of _sp_sum
// MODE 1: fused _build_I_map_conv: (w, theta, psi[, gamma]) -> (I, Es, Ep)
template <int R, int MODE, bool PLANAR, bool SYM, bool GC = false>
__global__ void __launch_bounds__(U_THREADS)
    undulator_kernel(size_t n, UParams p, UTables tab, const double* __restrict__ gamma, const double* __restrict__ wu,
                     const double* __restrict__ w, const double* __restrict__ ww1, const double* __restrict__ phi,
                     const double* __restrict__ psi, double2* __restrict__ Is, double2* __restrict__ Ip,
                     double* __restrict__ Iout) {
  __shared__ UNode tile[U_TILE];
  __shared__ double2 sctab[kSinTabEntries];
  const size_t base = (size_t)blockIdx.x * (U_THREADS * R);
  USample s[R];
  double acc[R][4];
  double abv[R], fin[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    size_t i = base + (size_t)r * U_THREADS + threadIdx.x;
    size_t ii = i < n ? i : n - 1;
    const double wi = w[ii], ph = phi[ii], ps = psi[ii];
    double g, wui, ww1i;
    if (MODE == 0) {
      g = gamma[ii];
      wui = wu[ii];
      ww1i = ww1[ii];
      abv[r] = 1.0;
    } else {
      g = gamma ? gamma[ii] : p.gamma0;
      sample_prefactors(p, g, wi, ph, ps, wui, ww1i, abv[r]);
    }
    s[r] = make_sample(p.Kx, p.Ky, g, wui, wi, ww1i, ph, ps);
    fin[r] = wui * XRTB_U_DIV(1.0, g);  // wuS*revgamma (synchr.py:2033)
    acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0;
  }

  node_loop<R, PLANAR, SYM, GC>(p, tab, tile, sctab, s, acc);

#pragma unroll
  for (int r = 0; r < R; ++r) {
    size_t i = base + (size_t)r * U_THREADS + threadIdx.x;
    if (i >= n) continue;
    const double f = fin[r];
    if (MODE == 0) {
      Is[i] = make_double2(f * acc[r][0], f * acc[r][1]);
      Ip[i] = make_double2(f * acc[r][2], f * acc[r][3]);
    } else {
      double I;
      double2 es, ep;
      imap_outputs(p, w[i], abv[r], f * acc[r][0], f * acc[r][1], f * acc[r][2], f * acc[r][3], I, es, ep);
      Iout[i] = I;
      Is[i] = es;
      Ip[i] = ep;
    }
  }
}

// (cos, sin)(k pi/512), k < 1024, correctly rounded (sincospi), once per device; the node loops stage it into shared memory
__global__ void sin_table_kernel(double2* tab) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= kSinTabEntries) return;
  double sn, cs;
  sincospi((double)k * (2.0 / (double)kSinTabEntries), &sn, &cs);
  tab[k] = make_double2(cs, sn);
}
const double2* sin_table_device(cudaStream_t st) {
  static double2* tabs[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!tabs[dev]) {
    double2* t = nullptr;
    if (cudaMalloc(&t, kSinTabEntries * sizeof(double2)) != cudaSuccess) return nullptr;
    sin_table_kernel<<<(kSinTabEntries + 255) / 256, 256, 0, st>>>(t);
    if (cudaStreamSynchronize(st) != cudaSuccess) {  // once per device: other streams may use the table next
      cudaFree(t);
      return nullptr;
    }
    tabs[dev] = t;
  }
  return tabs[dev];
}

static int launch(int mode, int R, bool symmetric, size_t n, const UParams& p, const UTables& t, const double* gamma,
                  const double* wu, const double* w, const double* ww1, const double* phi, const double* psi,
                  double* Is, double* Ip, double* I, cudaStream_t st) {
  if (n == 0) return 0;
  XRTB_REQUIRE(p.jend > 0, "undulator: the integration grid is empty");
  if (R == 0) R = 2;
  XRTB_REQUIRE(R == 1 || R == 2 || R == 4, "undulator: register blocking R must be 1, 2 or 4 (got %d)", R);
  const size_t per = (size_t)U_THREADS * R;
  const unsigned grid = (unsigned)((n + per - 1) / per);
  double2* is2 = reinterpret_cast<double2*>(Is);
  double2* ip2 = reinterpret_cast<double2*>(Ip);
  const bool planar = (p.Kx == 0.0);
  const bool sym = planar && symmetric;
  // fused I-map with one gamma for every sample (no per-sample gamma array): gamma-only node quantities from the tile
  const bool gconst = sym && mode == 1 && gamma == nullptr;
#define XRTB_U_LAUNCH(RR, MM)                                                                                         \
  do {                                                                                                                \
    if (gconst && MM == 1)                                                                                            \
      undulator_kernel<RR, 1, true, true, true><<<grid, U_THREADS, 0, st>>>(n, p, t, gamma, wu, w, ww1, phi, psi, is2, ip2, I); \
    else if (sym)                                                                                                     \
      undulator_kernel<RR, MM, true, true><<<grid, U_THREADS, 0, st>>>(n, p, t, gamma, wu, w, ww1, phi, psi, is2, ip2, I); \
    else if (planar)                                                                                                  \
      undulator_kernel<RR, MM, true, false><<<grid, U_THREADS, 0, st>>>(n, p, t, gamma, wu, w, ww1, phi, psi, is2, ip2, I); \
    else                                                                                                              \
      undulator_kernel<RR, MM, false, false><<<grid, U_THREADS, 0, st>>>(n, p, t, gamma, wu, w, ww1, phi, psi, is2, ip2, I); \
  } while (0)
  if (mode == 0) {
    if (R == 1) XRTB_U_LAUNCH(1, 0);
    else if (R == 2) XRTB_U_LAUNCH(2, 0);
    else XRTB_U_LAUNCH(4, 0);
  } else {
    if (R == 1) XRTB_U_LAUNCH(1, 1);
    else if (R == 2) XRTB_U_LAUNCH(2, 1);
    else XRTB_U_LAUNCH(4, 1);
  }
#undef XRTB_U_LAUNCH
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int undulator_raw_device(size_t n, int jend, double Kx, double Ky, const double* gamma, const double* wu,
                         const double* w, const double* ww1, const double* ddphi, const double* ddpsi,
                         const double* tg, const double* ag, const double* sintg, const double* costg,
                         const double* sintgph, const double* costgph, double* Is, double* Ip, int R,
                         int symmetric, cudaStream_t st) {
  UParams p{};
  p.Kx = Kx;
  p.Ky = Ky;
  p.jend = jend;
  UTables t{tg, ag, sintg, costg, sintgph, costgph, sin_table_device(st)};
  XRTB_REQUIRE(t.sintab != nullptr, "undulator: the sin table could not be set up");
  return launch(0, R, symmetric != 0, n, p, t, gamma, wu, w, ww1, ddphi, ddpsi, Is, Ip, nullptr, st);
}

int undulator_imap_device(size_t n, int jend, double Kx, double Ky, double L0, double Np, double gamma0,
                          double amp_const, double dstep, int bw_mode, const double* gamma, const double* w,
                          const double* theta, const double* psi, const double* tg, const double* ag,
                          const double* sintg, const double* costg, const double* sintgph, const double* costgph,
                          double* I, double* Es, double* Ep, int R, int symmetric, cudaStream_t st) {
  UParams p{};
  p.Kx = Kx;
  p.Ky = Ky;
  p.L0 = L0;
  p.Np = Np;
  p.gamma0 = gamma0;
  p.amp_const = amp_const;
  p.dstep = dstep;
  p.jend = jend;
  p.bw_mode = bw_mode;
  UTables t{tg, ag, sintg, costg, sintgph, costgph, sin_table_device(st)};
  XRTB_REQUIRE(t.sintab != nullptr, "undulator: the sin table could not be set up");
  return launch(1, R, symmetric != 0, n, p, t, gamma, nullptr, w, nullptr, theta, psi, Es, Ep, I, st);
}

}  // namespace xrtb
