// Tapered and near-field undulator amplitude integrals for sm_100a.
//
// Replaces: kernels `undulator_taper` (xrt/backends/raycing/cl/undulator.cl:123-202) and `undulator_nf`
// (undulator.cl:204-298), both launched from Undulator._build_I_map_CL (sources/synchr.py:2157-2176) with the
// scalar list [taper | R0*2pi/L0, Kx, Ky, jend, Np].  Numerics follow the numpy path Undulator._sp_sum
// (synchr.py:1929-2033: the `_taperVal` branch :1958-1971 and the `R0` branch :1972-1996, then the common
// velocity / acceleration block :2005-2031) -- the oracle (oracle/undulator.py) -- not the .cl text, which
// differs in beta'_z and in the sign of r_y (conventions paragraph).
//
// Both variants loop over the Np periods explicitly (no sin(N pi w)/sin(pi w) factor), so one sample costs
// Np * jend node evaluations; FP64-pipe bound like `undulator`.  One thread owns one sample; the node table
// (8 doubles per node incl. sin 2tg) is staged through shared memory and read as broadcasts.
//
// This translation unit is compiled with -fmad=false and every expression below is written in the reference's
// operation order: 1 - n.beta cancels ~1e8 (gamma^-2 scale) and, in the near field, n = dr/|dr| feeds that
// cancellation, so the chain has to be rounded exactly like numpy rounds it (IEEE sqrt and division included).
#include "common.cuh"
#include "kernels.h"
#include "undulator_core.cuh"

namespace xrtb {

struct __align__(16) ENode {
  double tg, ag, st, ct;    // node, weight, sin(tg), cos(tg)
  double sp, cp, s2, s2p;   // sin(tg+phase), cos(tg+phase), 2 sin cos (tg), 2 sin cos (tg+phase)
};

constexpr int E_THREADS = 128;
constexpr int E_TILE = 256;

__device__ __forceinline__ void load_ext_tile(ENode* tile, const UTables& t, int j0, int cnt) {
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
    const int g = j0 + j;
    ENode n;
    n.tg = t.tg[g];
    n.ag = t.ag[g];
    n.st = t.sintg[g];
    n.ct = t.costg[g];
    n.sp = t.sintgph[g];
    n.cp = t.costgph[g];
    n.s2 = (2.0 * n.st) * n.ct;    // synchr.py:1934
    n.s2p = (2.0 * n.sp) * n.cp;   // synchr.py:1935
    tile[j] = n;
  }
}

// VARIANT 1: taper (scalar0 = _taperVal), VARIANT 2: near field (scalar0 = R0 * pi * 2 / L0)
// MODE 0: reference kernel interface -> raw (Is, Ip) of _sp_sum; MODE 1: fused _build_I_map_conv -> (I, Es, Ep)
template <int VARIANT, int MODE>
__global__ void __launch_bounds__(E_THREADS)
    undulator_ext_kernel(size_t n, UParams p, double scalar0, int nmax, UTables tab,
                         const double* __restrict__ gamma, const double* __restrict__ wu_in,
                         const double* __restrict__ w_in, const double* __restrict__ ww1_in,
                         const double* __restrict__ phi_in, const double* __restrict__ psi_in,
                         double2* __restrict__ Is, double2* __restrict__ Ip, double* __restrict__ Iout) {
  __shared__ ENode tile[E_TILE];
  const double PI_ = 3.1415926535897932384626433832795;
  const double PI2_ = 6.283185307179586476925286766559;
  const size_t i = (size_t)blockIdx.x * E_THREADS + threadIdx.x;
  const size_t ii = i < n ? i : n - 1;
  const double wS = w_in[ii], ddphi = phi_in[ii], ddpsi = psi_in[ii];
  double gS, wuS, ww1S, ab = 1.0;
  if (MODE == 0) {
    gS = gamma[ii];
    wuS = wu_in[ii];
    ww1S = ww1_in[ii];
  } else {
    gS = gamma ? gamma[ii] : p.gamma0;
    double ab_ff;
    sample_prefactors(p, gS, wS, ddphi, ddpsi, wuS, ww1S, ab_ff);
    ab = 1.0 / PI2_ / wuS;  // synchr.py:2067-2068
  }
  const double Kx = p.Kx, Ky = p.Ky;
  const double Kx2 = Kx * Kx, Ky2 = Ky * Ky;
  const double revgamma = 1.0 / gS;
  const double revgamma2 = revgamma * revgamma;
  const double betam = 1.0 - (1.0 + 0.5 * Kx2 + 0.5 * Ky2) * 0.5 * revgamma2;  // synchr.py:1938
  const double wwuS = wS / wuS;
  const double wr = wwuS * revgamma;
  const double hrg = 0.5 * revgamma, erg = 0.125 * revgamma;
  double dirx = ddphi, diry = ddpsi;
  double dirz = 1.0 - 0.5 * (ddphi * ddphi + ddpsi * ddpsi);  // synchr.py:1945
  // taper constants (synchr.py:1960-1961)
  const double alphaS = VARIANT == 1 ? scalar0 / 5067.7309392068091 : 0.0;
  const double aw = alphaS / wuS, a2w = (2.0 * alphaS) / wuS;
  const double nKyd = (-Ky) * dirx, Kxd = Kx * diry;
  // near-field constants (synchr.py:2073-2077, :1949-1951)
  double R0x = 0.0, R0y = 0.0, R0z = 0.0, sinr0z = 0.0, cosr0z = 1.0;
  if (VARIANT == 2) {
    R0x = tan(ddphi) * scalar0;
    R0y = tan(ddpsi) * scalar0;
    R0z = scalar0;
    sincos(R0z, &sinr0z, &cosr0z);
  }
  const double wz = 1.0 - betam;
  double bs_re = 0.0, bs_im = 0.0, bp_re = 0.0, bp_im = 0.0;

  const bool single = p.jend <= E_TILE;
  if (single) {
    load_ext_tile(tile, tab, 0, p.jend);
    __syncthreads();
  }
  for (int ip = 0; ip < nmax; ++ip) {
    const double z0 = (-(double)(nmax - 1)) * PI_ + (double)ip * PI2_;  // synchr.py:1959 / :1975
    for (int j0 = 0; j0 < p.jend; j0 += E_TILE) {
      const int cnt = min(E_TILE, p.jend - j0);
      if (!single) {
        __syncthreads();
        load_ext_tile(tile, tab, j0, cnt);
        __syncthreads();
      }
#pragma unroll 1
      for (int j = 0; j < cnt; ++j) {
        const double4* q = reinterpret_cast<const double4*>(tile + j);
        const double4 n0 = q[0];  // tg, ag, sin, cos
        const double4 n1 = q[1];  // sin_ph, cos_ph, sin2x, sin2x_ph
        const double zloc = z0 + n0.x;
        double taperC = 1.0, ere, eim;
        if (VARIANT == 1) {
          taperC = 1.0 - (alphaS * zloc) / wuS;  // synchr.py:1961
          const double t1 = (1.0 - n0.w) - zloc * n0.z;
          const double A = nKyd * (n0.z + aw * t1);
          const double B = Kxd * n0.z;  // the reference uses sintg here, not sintgph (synchr.py:1967)
          const double t3 = (zloc * zloc + n0.w * n0.w) + zloc * n1.z;
          const double Cc = erg * (Kx2 * n1.w + Ky2 * (n1.z - a2w * t3));
          const double ucos = ww1S * zloc + wr * ((A + B) + Cc);
          sincos_big(ucos, eim, ere);
        } else {
          const double zterm = (0.5 * (Ky2 * n1.z + Kx2 * n1.w)) * revgamma;  // synchr.py:1973-1974
          const double qz = (0.25 * zterm) * revgamma;
          const double rx = (Ky * n0.z) * revgamma;
          const double ry = (Kx * n1.x) * revgamma;
          const double rz = betam * zloc - qz;
          const double drx = R0x - rx, dry = R0y - ry, drz = R0z - rz;
          const double dist = sqrt((drx * drx + dry * dry) + drz * drz);  // np.linalg.norm(dr, axis=0)
          const double drs = (0.5 * (drx * drx + dry * dry)) / drz;
          double sinzloc, coszloc, sindrs, cosdrs;
          sincos_big((wwuS * zloc) * wz, sinzloc, coszloc);
          sincos_big(wwuS * (drs + qz), sindrs, cosdrs);
          ere = (((-sinr0z) * sinzloc) * cosdrs - (sinr0z * coszloc) * sindrs - (cosr0z * sinzloc) * sindrs) +
                (cosr0z * coszloc) * cosdrs;  // synchr.py:1986-1987
          eim = ((((-sinr0z) * sinzloc) * sindrs + (sinr0z * coszloc) * cosdrs) + (cosr0z * sinzloc) * cosdrs) +
                (cosr0z * coszloc) * sindrs;  // synchr.py:1988-1989
          dirx = drx / dist;
          diry = dry / dist;
          dirz = drz / dist;
        }
        // velocity, acceleration (synchr.py:2005-2016)
        const double betax = ((taperC * Ky) * revgamma) * n0.w;
        const double betay = ((-Kx) * revgamma) * n1.y;
        const double betaz = 1.0 - 0.5 * ((revgamma2 + betax * betax) + betay * betay);
        const double betaPx = (-Ky) * (alphaS * n0.w + taperC * n0.z);
        const double betaPy = Kx * n1.x;
        const double betaPz = hrg * ((Ky2 * taperC) * ((alphaS * (n0.w * n0.w)) + taperC * n1.z) + Kx2 * n1.w);
        const double rkrel = 1.0 / (((1.0 - dirx * betax) - diry * betay) - dirz * betaz);
        const double wgt = n0.y * (rkrel * rkrel);
        ere *= wgt;
        eim *= wgt;
        const double bnx = dirx - betax, bny = diry - betay, bnz = dirz - betaz;
        const double ddbp = (dirx * betaPx + diry * betaPy) + dirz * betaPz;
        const double ddmb = (dirx * bnx + diry * bny) + dirz * bnz;
        const double ts = bnx * ddbp - betaPx * ddmb;
        const double tp = bny * ddbp - betaPy * ddmb;
        bs_re += ere * ts;
        bs_im += eim * ts;
        bp_re += ere * tp;
        bp_im += eim * tp;
      }
    }
  }
  if (i >= n) return;
  const double f = wuS * revgamma;  // synchr.py:2033
  if (MODE == 0) {
    Is[i] = make_double2(f * bs_re, f * bs_im);
    Ip[i] = make_double2(f * bp_re, f * bp_im);
  } else {
    double I;
    double2 es, ep;
    imap_outputs(p, wS, ab, f * bs_re, f * bs_im, f * bp_re, f * bp_im, I, es, ep);
    Iout[i] = I;
    Is[i] = es;
    Ip[i] = ep;
  }
}

// ---------------------------------------------------------------- custom field (SourceFromField)
// Replaces kernel `custom_field` (cl/undulator.cl:1001-...) launched from
// SourceFromField._build_I_map_custom_field_CL (sources/synchr.py:1227-1259) with scalars [jend, betam, R0], sliced
// [gamma, w, ddtheta, ddpsi], node tables [tg, ag, Bx, By, Bz, betax, betay, trajx, trajy, trajz].  Numerics follow
// the numpy path SourceFromField._sp_sum (synchr.py:888-973), not filamentBeam: far field, or near field for R0 > 0.
// The phase wc * (tg - n_z z(tg)) is a ~1e7 cancellation (wc ~ 5e7 / mm against a path difference of ~1e-4 mm out of
// ~1e3 mm), so everything is evaluated in the reference's rounding order (this file is compiled with -fmad=false).
struct __align__(16) FNode {
  double tg, ag, betax, betay;
  double trajx, trajy, trajz, Bx;
  double By, Bz, pad0, pad1;
};
constexpr int F_TILE = 128;

struct FTables {
  const double *tg, *ag, *Bx, *By, *Bz, *betax, *betay, *trajx, *trajy, *trajz;
};

template <bool NF>
__global__ void __launch_bounds__(E_THREADS)
    custom_field_kernel(size_t n, int jend, double betam, double R0, FTables tab, const double* __restrict__ gamma,
                        const double* __restrict__ w_in, const double* __restrict__ phi_in,
                        const double* __restrict__ psi_in, double2* __restrict__ Is, double2* __restrict__ Ip) {
  __shared__ FNode tile[F_TILE];
  const double EMC_ = 0.5866791802416487;
  const double K0 = 1.602176565e-19 / 9.109383701528e-31 / 2.99792458e10 / 10.;  // SIE0 / SIM0 / C / 10.
  const size_t i = (size_t)blockIdx.x * E_THREADS + threadIdx.x;
  const size_t ii = i < n ? i : n - 1;
  const double g = gamma[ii], wS = w_in[ii], ddphi = phi_in[ii], ddpsi = psi_in[ii];
  const double emcg = K0 / g;
  double dirx = ddphi, diry = ddpsi;
  double dirz = sqrt(1.0 - ddphi * ddphi - ddpsi * ddpsi);  // synchr.py:899
  const double revgamma2 = 1.0 / (g * g);
  const double wc = wS * 5067.7309392068091 / betam;  // w * E2WC / betam
  const double zf = 1.0 - 0.5 * revgamma2, zt = EMC_ * EMC_ * revgamma2;
  double R0x = 0.0, R0y = 0.0, R0z = 0.0, sinr0z = 0.0, cosr0z = 1.0;
  if (NF) {
    R0x = tan(ddphi) * R0;
    R0y = tan(ddpsi) * R0;
    R0z = 1.0 * R0;
    sincos_big(wc * R0z, sinr0z, cosr0z);
  }
  double bs_re = 0.0, bs_im = 0.0, bp_re = 0.0, bp_im = 0.0;
  for (int j0 = 0; j0 < jend; j0 += F_TILE) {
    const int cnt = min(F_TILE, jend - j0);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
      const int q = j0 + j;
      FNode nd;
      nd.tg = tab.tg[q];
      nd.ag = tab.ag[q];
      nd.betax = tab.betax[q];
      nd.betay = tab.betay[q];
      nd.trajx = tab.trajx[q];
      nd.trajy = tab.trajy[q];
      nd.trajz = tab.trajz[q];
      nd.Bx = tab.Bx[q];
      nd.By = tab.By[q];
      nd.Bz = tab.Bz[q];
      nd.pad0 = nd.pad1 = 0.0;
      tile[j] = nd;
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < cnt; ++j) {
      const double4* q4 = reinterpret_cast<const double4*>(tile + j);
      const double4 n0 = q4[0];  // tg ag betax betay
      const double4 n1 = q4[1];  // trajx trajy trajz Bx
      const double4 n2 = q4[2];  // By Bz
      const double betax_ = emcg * n0.z, betay_ = emcg * n0.w;
      const double trajx_ = emcg * n1.x, trajy_ = emcg * n1.y;
      const double trajz_ = n0.x * zf + zt * n1.z;  // synchr.py:919-920
      double ere, eim;
      if (NF) {
        const double drx = R0x - trajx_, dry = R0y - trajy_, drz = R0z - trajz_;
        const double dist = sqrt((drx * drx + dry * dry) + drz * drz);
        const double rdrz = 1.0 / drz;
        const double drs = (drx * drx + dry * dry) * rdrz;
        const double LRS = (0.5 * drs - (0.125 * (drs * drs)) * rdrz) + (0.0625 * (drs * drs * drs)) * (rdrz * rdrz);
        double sinzloc, coszloc, sindrs, cosdrs;
        sincos_big(wc * (n0.x - trajz_), sinzloc, coszloc);
        sincos_big(wc * LRS, sindrs, cosdrs);
        ere = (((-sinr0z) * sinzloc) * cosdrs - (sinr0z * coszloc) * sindrs - (cosr0z * sinzloc) * sindrs) +
              (cosr0z * coszloc) * cosdrs;
        eim = ((((-sinr0z) * sinzloc) * sindrs + (sinr0z * coszloc) * cosdrs) + (cosr0z * sinzloc) * cosdrs) +
              (cosr0z * coszloc) * sindrs;
        dirx = drx / dist;
        diry = dry / dist;
        dirz = drz / dist;
      } else {
        const double phz = wc * (n0.x - dirz * trajz_);
        const double phxy = wc * (dirx * trajx_ + diry * trajy_);
        double sinphz, cosphz, sinphxy, cosphxy;
        sincos_big(phz, sinphz, cosphz);
        sincos_big(phxy, sinphxy, cosphxy);
        ere = sinphz * cosphxy - cosphz * sinphxy;
        eim = cosphz * cosphxy + sinphz * sinphxy;
      }
      const double sm = (revgamma2 + betax_ * betax_) + betay_ * betay_;
      const double betaz = ((1.0 - 0.5 * sm) - 0.125 * (sm * sm)) - 0.0625 * (sm * sm * sm);
      const double betaPx = betay_ * n2.y - betaz * n2.x;
      const double betaPy = (-betax_) * n2.y + betaz * n1.w;
      const double betaPz = betax_ * n2.x - betay_ * n1.w;
      const double rkrel = 1.0 / (((1.0 - dirx * betax_) - diry * betay_) - dirz * betaz);
      const double wgt = n0.y * (rkrel * rkrel);
      ere *= wgt;
      eim *= wgt;
      const double bnx = dirx - betax_, bny = diry - betay_, bnz = dirz - betaz;
      const double ddbp = (dirx * betaPx + diry * betaPy) + dirz * betaPz;
      const double ddmb = (dirx * bnx + diry * bny) + dirz * bnz;
      const double ts = bnx * ddbp - betaPx * ddmb;
      const double tp = bny * ddbp - betaPy * ddmb;
      bs_re += ere * ts;
      bs_im += eim * ts;
      bp_re += ere * tp;
      bp_im += eim * tp;
    }
  }
  if (i >= n) return;
  Is[i] = make_double2(bs_re * emcg, bs_im * emcg);
  Ip[i] = make_double2(bp_re * emcg, bp_im * emcg);
}

int custom_field_device(size_t n, int jend, double betam, double R0, const double* gamma, const double* w,
                        const double* ddtheta, const double* ddpsi, const double* tg, const double* ag,
                        const double* Bx, const double* By, const double* Bz, const double* betax,
                        const double* betay, const double* trajx, const double* trajy, const double* trajz,
                        double* Is, double* Ip, cudaStream_t st) {
  if (n == 0) return 0;
  XRTB_REQUIRE(jend > 0, "custom_field: the integration grid is empty");
  XRTB_REQUIRE(gamma && w && ddtheta && ddpsi && tg && ag && Bx && By && Bz && betax && betay && trajx && trajy &&
                   trajz && Is && Ip,
               "custom_field: NULL argument");
  FTables t{tg, ag, Bx, By, Bz, betax, betay, trajx, trajy, trajz};
  const unsigned grid = (unsigned)((n + E_THREADS - 1) / E_THREADS);
  double2* is2 = reinterpret_cast<double2*>(Is);
  double2* ip2 = reinterpret_cast<double2*>(Ip);
  if (R0 != 0.0)  // `if self.R0:` (synchr.py:1319)
    custom_field_kernel<true><<<grid, E_THREADS, 0, st>>>(n, jend, betam, R0, t, gamma, w, ddtheta, ddpsi, is2, ip2);
  else
    custom_field_kernel<false><<<grid, E_THREADS, 0, st>>>(n, jend, betam, R0, t, gamma, w, ddtheta, ddpsi, is2, ip2);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int undulator_ext_device(int variant, int mode, size_t n, int jend, int nmax, double scalar0, double Kx, double Ky,
                         double L0, double gamma0, double amp_const, double dstep, int bw_mode, const double* gamma,
                         const double* wu, const double* w, const double* ww1, const double* ddphi,
                         const double* ddpsi, const double* tg, const double* ag, const double* sintg,
                         const double* costg, const double* sintgph, const double* costgph, double* Is, double* Ip,
                         double* I, cudaStream_t st) {
  if (n == 0) return 0;
  XRTB_REQUIRE(variant == 1 || variant == 2, "undulator_ext: variant must be 1 (taper) or 2 (near field), got %d",
               variant);
  XRTB_REQUIRE(jend > 0, "undulator_ext: the integration grid is empty");
  XRTB_REQUIRE(nmax > 0, "undulator_ext: the number of periods must be positive (got %d)", nmax);
  UParams p{};
  p.Kx = Kx;
  p.Ky = Ky;
  p.L0 = L0;
  p.Np = (double)nmax;
  p.gamma0 = gamma0;
  p.amp_const = amp_const;
  p.dstep = dstep;
  p.jend = jend;
  p.bw_mode = bw_mode;
  UTables t{tg, ag, sintg, costg, sintgph, costgph, nullptr};
  const unsigned grid = (unsigned)((n + E_THREADS - 1) / E_THREADS);
  double2* is2 = reinterpret_cast<double2*>(Is);
  double2* ip2 = reinterpret_cast<double2*>(Ip);
#define XRTB_E_LAUNCH(V, M) \
  undulator_ext_kernel<V, M><<<grid, E_THREADS, 0, st>>>(n, p, scalar0, nmax, t, gamma, wu, w, ww1, ddphi, ddpsi, is2, ip2, I)
  if (variant == 1 && mode == 0) XRTB_E_LAUNCH(1, 0);
  else if (variant == 1) XRTB_E_LAUNCH(1, 1);
  else if (mode == 0) XRTB_E_LAUNCH(2, 0);
  else XRTB_E_LAUNCH(2, 1);
#undef XRTB_E_LAUNCH
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace xrtb
