// Source-side kernels for sm_100a: Monte-Carlo sampling and Beam assembly of IntegratedSource.shine.
//
// Replaces the per-ray numpy body of IntegratedSource.shine (xrt/backends/raycing/sources/sybase.py:1580-1801):
//   * sample_kernel : uniform (E, theta, psi) in the source window (:1583-1617) and the four normal deviates
//                     used for source size / divergence smearing (:1669-1697) -- counter-based Philox4x32-10,
//                     one independent counter range per ray, so rank r of N just offsets the counter;
//   * fill_kernel   : Beam assembly from (E, theta, psi, I, mJs, mJp): positions ~ N(0, SIGMA(E))
//                     (Undulator.get_SIGMA, sources/synchr.py:2200-2249, eEspread = 0), directions, coherency
//                     matrix / amplitudes for uniformRayDensity or rejection-normalised rays (:1699-1721),
//                     normalisation (:1769-1772), pitch/yaw (:1785-1786), virgin_local_to_global (:1800).
// The field integral itself is undulator.cu.  Oracle: oracle/shine.py (same deviates fed to both).
// Compiled with -fmad=false (numpy rounding order).
#include "../../include/xrt_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "undulator_core.cuh"

namespace xrtb {

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al., SC'11)
struct Philox {
  uint32_t c[4];
  uint32_t k[2];
};
__device__ __forceinline__ void philox_round(uint32_t* c, const uint32_t* k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0;
  c[1] = n1;
  c[2] = n2;
  c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint64_t counter, uint32_t stream, uint64_t seed, uint32_t* out) {
  uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), stream, 0u};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  out[0] = c[0];
  out[1] = c[1];
  out[2] = c[2];
  out[3] = c[3];
}
// 53-bit uniform in [0, 1) from two 32-bit words (the construction of numpy's random_sample)
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

// Quotients and roots of ordinary positive numbers in the per-ray set-up of the undulator source (photon energies,
// gamma, beam sizes, the direction norm >= 1): the compiler's IEEE sequences without their range test (common.cuh,
// div_nr / sqrt_nr: same bits).  -DXRTB_SOURCE_DIV_CHECKED restores the plain operators.
#ifdef XRTB_SOURCE_DIV_CHECKED
#define XRTB_SRC_DIV(a, b) ((a) / (b))
#define XRTB_SRC_SQRT(a) sqrt(a)
#else
#define XRTB_SRC_DIV(a, b) div_nr(a, b)
#define XRTB_SRC_SQRT(a) sqrt_nr(a)
#endif

struct Draw {
  double E, theta, psi, u, n1, n2, n3, n4, n5;
};
// E, theta, psi uniform in their windows (sybase.py:1583-1617), u uniform (rejection test, :1637), n1..n4
// standard normal (Box-Muller on (0, 1] uniforms); four Philox blocks per ray
__device__ __forceinline__ Draw draw_ray(uint64_t ctr, uint64_t seed, double Emin, double Emax, double Tmin,
                                         double Tmax, double Pmin, double Pmax, bool normals, bool spread) {
  Draw d;
  uint32_t w[4];
  philox4x32_10(ctr, 0u, seed, w);
  const double r0 = u53(w[0], w[1]), r1 = u53(w[2], w[3]);
  philox4x32_10(ctr, 1u, seed, w);
  const double r2 = u53(w[0], w[1]), r3 = u53(w[2], w[3]);
  d.E = r0 * (Emax - Emin) + Emin;
  d.theta = r1 * (Tmax - Tmin) + Tmin;
  d.psi = r2 * (Pmax - Pmin) + Pmin;
  d.u = r3;
  d.n1 = d.n2 = d.n3 = d.n4 = 0.0;
  if (normals) {
    philox4x32_10(ctr, 2u, seed, w);
    const double a0 = 1.0 - u53(w[0], w[1]), a1 = u53(w[2], w[3]);
    philox4x32_10(ctr, 3u, seed, w);
    const double b0 = 1.0 - u53(w[0], w[1]), b1 = u53(w[2], w[3]);
    const double ra = XRTB_SRC_SQRT(-2.0 * log(a0)), rb = XRTB_SRC_SQRT(-2.0 * log(b0));  // arguments >= +0
    double sn, cs;
    sincospi(2.0 * a1, &sn, &cs);
    d.n1 = ra * cs;
    d.n2 = ra * sn;
    sincospi(2.0 * b1, &sn, &cs);
    d.n3 = rb * cs;
    d.n4 = rb * sn;
  }
  d.n5 = 0.0;
  if (spread) {  // electron energy deviate (one more Philox block)
    philox4x32_10(ctr, 4u, seed, w);
    const double c0 = 1.0 - u53(w[0], w[1]), c1 = u53(w[2], w[3]);
    double sn, cs;
    sincospi(2.0 * c1, &sn, &cs);
    d.n5 = sqrt(-2.0 * log(c0)) * cs;
  }
  return d;
}

__global__ void __launch_bounds__(256) sample_kernel(size_t n, uint64_t seed, uint64_t index_offset, double Emin,
                                                     double Emax, double Tmin, double Tmax, double Pmin, double Pmax,
                                                     double* __restrict__ E, double* __restrict__ theta,
                                                     double* __restrict__ psi, double* __restrict__ n1,
                                                     double* __restrict__ n2, double* __restrict__ n3,
                                                     double* __restrict__ n4, double* __restrict__ u,
                                                     double* __restrict__ n5) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Draw d = draw_ray(index_offset + i, seed, Emin, Emax, Tmin, Tmax, Pmin, Pmax, n1 != nullptr, n5 != nullptr);
  if (n5) n5[i] = d.n5;
  E[i] = d.E;
  theta[i] = d.theta;
  psi[i] = d.psi;
  if (u) u[i] = d.u;
  if (n1) {
    n1[i] = d.n1;
    n2[i] = d.n2;
    n3[i] = d.n3;
    n4[i] = d.n4;
  }
}

__device__ __forceinline__ cplx cmul_np3(cplx a, cplx b) {  // numpy SIMD complex multiply (see reflect.cu)
  const double ii = a.im * b.im, ir = a.im * b.re;
  return cplx{fma(a.re, b.re, -ii), fma(a.re, b.im, ir)};
}

// Beam assembly of one ray (sybase.py:1651-1721, :1769-1801) into slot `o` of the output beam
__device__ __forceinline__ void fill_ray(const xrt_source_params& p, double E, double rTheta, double rPsi, cplx js,
                                         cplx jp, double g1, double g2, double g3, double g4, const xrt_beam& out,
                                         size_t o) {
  // get_SIGMA (synchr.py:2240-2249) with get_sigma_r2 = get_sigma_r02 (eEspread = 0): 2*CHeVcm/E*10*L0*Np/PI2**2
  const double CHeVcm = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12;
  const double PI2 = 6.283185307179586476925286766559;
  double sr2 = XRTB_SRC_DIV(2.0 * CHeVcm, E) * 10.0 * p.L0 * p.Np / (PI2 * PI2);
  if (p.eEspread > 0.0) {
    // get_sigma_r2 with energy spread (synchr.py:2210-2223, onlyOddHarmonics=False as shine passes it) and
    // tanaka_kitamura_Qa2 (sybase.py:664-674)
    const double harmonic = floor(E / p.E1);  // np.floor_divide for positive operands
    const double enorm = PI2 * harmonic * p.Np * p.eEspread;
    const double xq = enorm / 4.0;
    double Qa2 = 1.0;
    if (xq > 1e-6) {
      const double yq = 1.4142135623730951 * xq;
      const double y2 = yq * yq;
      Qa2 = y2 / (exp(-y2) + 1.7724538509055159 * yq * erf(yq) - 1.0);
    }
    sr2 = sr2 * pow(Qa2, 2.0 / 3.0);
  }
  const double SX = XRTB_SRC_SQRT(p.dx * p.dx + sr2), SZ = XRTB_SRC_SQRT(p.dz * p.dz + sr2);
  double x = 0.0 + SX * g1;  // np.random.normal(0, SIGMA, n) = loc + scale * gauss
  double y = 0.0;
  double z = 0.0 + SZ * g2;
  double a = rTheta, b = 1.0, c = rPsi;
  if (p.dxprime > 0.0) a += 0.0 + p.dxprime * g3;
  if (p.dzprime > 0.0) c += 0.0 + p.dzprime * g4;
  const double mJs2 = cmul_np3(js, cconj(js)).re;
  const double mJp2 = cmul_np3(jp, cconj(jp)).re;
  cplx Jsp = cmul_np3(js, cconj(jp));
  double Jss, Jpp;
  cplx Es, Ep;
  if (p.uniform) {
    // sSP = 1. (sybase.py:1702-1707); x / 1. is exact
    Jss = mJs2;
    Jpp = mJp2;
    Es = js;
    Ep = jp;
  } else {
    const double sSP = mJs2 + mJp2;
    if (sSP != 0.0) {
      Jsp = cdiv(Jsp, cplx{sSP, 0.0});
      Jss = mJs2 / sSP;
      Jpp = mJp2 / sSP;
    } else {
      Jsp = cplx{0.0, 0.0};
      Jss = 0.0;
      Jpp = 0.0;
    }
    Es = cdiv(js, cplx{sqrt(mJs2), 0.0});  // mJs / mJs2**0.5 (sybase.py:1719-1720)
    Ep = cdiv(jp, cplx{sqrt(mJp2), 0.0});
  }
  // normalise (sybase.py:1769-1772)
  const double norm = XRTB_SRC_SQRT(a * a + b * b + c * c);  // b = 1: norm >= 1
#ifdef XRTB_SOURCE_DIV_CHECKED
  a /= norm;
  b /= norm;
  c /= norm;
#else
  {
    const Divisor dn = make_divisor(norm);  // three quotients, one refined reciprocal (same bits as a / norm etc.)
    a = div_by_nr(a, dn);
    b = div_by_nr(b, dn);
    c = div_by_nr(c, dn);
  }
#endif
  // pitch / yaw of the source (sybase.py:1785-1786): rotate_beam default sequence z then x
  if (p.has_yaw) {
    const double x1 = p.cosYaw * x - p.sinYaw * y, y1 = p.sinYaw * x + p.cosYaw * y;
    x = x1;
    y = y1;
    const double a1 = p.cosYaw * a - p.sinYaw * b, b1 = p.sinYaw * a + p.cosYaw * b;
    a = a1;
    b = b1;
  }
  if (p.has_pitch) {
    const double y1 = p.cosPitch * y - p.sinPitch * z, z1 = p.sinPitch * y + p.cosPitch * z;
    y = y1;
    z = z1;
    const double b1 = p.cosPitch * b - p.sinPitch * c, c1 = p.sinPitch * b + p.cosPitch * c;
    b = b1;
    c = c1;
  }
  if (p.to_global) {  // virgin_local_to_global (beamline.py:267-287)
    if (p.sinAz != 0.0) {
      const double a1 = p.cosAz * a - (-p.sinAz) * b, b1 = (-p.sinAz) * a + p.cosAz * b;
      a = a1;
      b = b1;
      const double x1 = p.cosAz * x - (-p.sinAz) * y, y1 = (-p.sinAz) * x + p.cosAz * y;
      x = x1;
      y = y1;
    }
    x += p.center[0];
    y += p.center[1];
    z += p.center[2];
  }
  out.x[o] = x;
  out.y[o] = y;
  out.z[o] = z;
  out.a[o] = a;
  out.b[o] = b;
  out.c[o] = c;
  out.path[o] = 0.0;
  out.E[o] = E;
  out.Jss[o] = Jss;
  out.Jpp[o] = Jpp;
  reinterpret_cast<double2*>(out.Jsp)[o] = make_double2(Jsp.re, Jsp.im);
  if (out.Es) {
    reinterpret_cast<double2*>(out.Es)[o] = make_double2(Es.re, Es.im);
    reinterpret_cast<double2*>(out.Ep)[o] = make_double2(Ep.re, Ep.im);
  }
  out.state[o] = 1;
}

// Beam assembly from sample arrays. idx: optional gather index (accepted samples of the rejection pass)
__global__ void __launch_bounds__(256) fill_kernel(size_t n, xrt_source_params p, const long long* __restrict__ idx,
                                                   const double* __restrict__ rE, const double* __restrict__ rTheta,
                                                   const double* __restrict__ rPsi, const double2* __restrict__ mJs,
                                                   const double2* __restrict__ mJp, const double* __restrict__ n1,
                                                   const double* __restrict__ n2, const double* __restrict__ n3,
                                                   const double* __restrict__ n4, xrt_beam out, size_t out_offset) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t j = idx ? (size_t)idx[i] : i;  // sample index
  const double2 s2 = mJs[j], p2 = mJp[j];
  fill_ray(p, rE[j], rTheta[j], rPsi[j], cplx{s2.x, s2.y}, cplx{p2.x, p2.y}, n1[j], n2[j], n3[j], n4[j], out,
           out_offset + i);
}

// ---------------------------------------------------------------- fused source for uniformRayDensity
// sample -> field integral -> Beam in one kernel: the (E, theta, psi) samples, the four deviates and
// (I, mJs, mJp) never travel through HBM (saves 64 + 64 + 40 + 40 B/ray and two launches); the Philox /
// Box-Muller integer work and the 132 B/ray of stores hide under the FP64-bound node loop.  Same device
// functions as the three separate kernels, so the beams are bit-identical (tests/test_shine.py).
// batches of rays per block (the node tile and the sin table staged once): 4 and 8 measured 15.0 ms instead of 14.1 ms per
// 1e8 rays -- the loop-carried state costs more than the 16 KB of table staging per 256 rays
#ifndef XRTB_SHINE_BATCHES
#define XRTB_SHINE_BATCHES 1
#endif
template <int R, bool PLANAR, bool SYM, bool GC = false>
__global__ void __launch_bounds__(U_THREADS, (R <= 2) ? 4 : 1)
    shine_uniform_kernel(size_t n, uint64_t seed, uint64_t index_offset, double Emin, double Emax, double Tmin,
                         double Tmax, double Pmin, double Pmax, UParams up, UTables tab, xrt_source_params sp,
                         xrt_beam out, unsigned long long* imax_slot) {
  __shared__ UNode tile[U_TILE];
  __shared__ double2 sctab[kSinTabEntries];
  double imax = 0.0;
  // XRTB_SHINE_BATCHES batches of U_THREADS * R rays per block: the node tile and the sin table are staged once
#pragma unroll 1
  for (int bt = 0; bt < XRTB_SHINE_BATCHES; ++bt) {
    const size_t base = ((size_t)blockIdx.x * XRTB_SHINE_BATCHES + bt) * (U_THREADS * R);
    if (base >= n) break;  // (the same for every thread of the block)
    Draw d[R];
    USample s[R];
    double acc[R][4];
    double abv[R], fin[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const size_t i = base + (size_t)r * U_THREADS + threadIdx.x;
      d[r] = draw_ray(index_offset + i, seed, Emin, Emax, Tmin, Tmax, Pmin, Pmax, true, sp.eEspread > 0.0);
      // gamma += gamma * eEspread * normal (synchr.py:2053-2059)
      const double g = (sp.eEspread > 0.0) ? up.gamma0 + up.gamma0 * sp.eEspread * d[r].n5 : up.gamma0;
      double wui, ww1i;
      sample_prefactors(up, g, d[r].E, d[r].theta, d[r].psi, wui, ww1i, abv[r]);
      s[r] = make_sample(up.Kx, up.Ky, g, wui, d[r].E, ww1i, d[r].theta, d[r].psi);
      fin[r] = wui * XRTB_SRC_DIV(1.0, g);
      acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0;
    }
    node_loop<R, PLANAR, SYM, GC>(up, tab, tile, sctab, s, acc, bt == 0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const size_t i = base + (size_t)r * U_THREADS + threadIdx.x;
      if (i >= n) continue;
      const double f = fin[r];
      double I;
      double2 es, ep;
      imap_outputs(up, d[r].E, abv[r], f * acc[r][0], f * acc[r][1], f * acc[r][2], f * acc[r][3], I, es, ep);
      imax = fmax(imax, I);
      fill_ray(sp, d[r].E, d[r].theta, d[r].psi, cplx{es.x, es.y}, cplx{ep.x, ep.y}, d[r].n1, d[r].n2, d[r].n3,
               d[r].n4, out, i);
    }
  }
  // beam-wide max intensity (Imax bookkeeping, sybase.py:1627-1631): non-negative doubles order like integers
  imax = warp_reduce(imax, OpMax());
  if ((threadIdx.x & 31) == 0 && imax > 0.0) atomicMax(imax_slot, (unsigned long long)__double_as_longlong(imax));
}

// ---------------------------------------------------------------- f3: bending magnet / wiggler
// Replaces the per-ray numpy body of BendingMagnet.build_I_map (sources/synchr.py:185-226, Wiggler = the isMPW
// branch) and of BendingMagnet.shine (synchr.py:321-480).  Oracle: oracle/synchrotron.py.

__device__ __forceinline__ bool finite_d(double v) { return fabs(v) <= 1.7976931348623157e308; }

// One sample of BendingMagnet.build_I_map.  ampS is purely imaginary and ampP purely real in the reference's
// complex arithmetic (-0.5j * ..., then 1j * ...): js = (0, s), jp = (p, 0).  Complex-by-real division in numpy
// multiplies by the reciprocal (Smith's algorithm with a zero imaginary divisor), reproduced here.
__device__ __forceinline__ void bm_sample(const xrt_bm_params& p, double gamma, double dde, double ddtheta,
                                          double ddpsi, double& I, cplx& js, cplx& jp) {
  const double SIE0 = 1.602176565e-19, SIM0 = 9.109383701528e-31, E2W = 1519267514747457.9195337718065469;
  const double PI_ = 3.1415926535897932384626433832795, SQ3 = 1.7320508075688772935274463415059;
  const double FINE_STR = 1 / 137.03599976;
  const double gamma2 = gamma * gamma;
  double w_cr = 1.5 * gamma2 * p.B * SIE0 / SIM0;
  if (p.isMPW) {
    const double thetaArg = ddtheta * gamma / p.K;
    double cl = 1.0 - thetaArg * thetaArg;
    if (cl < 0.0) cl = 0.0;
    w_cr *= sqrt(cl);
  }
  if (!finite_d(w_cr)) w_cr = 0.0;
  const double gammapsi = gamma * ddpsi;
  const double g2p1 = gammapsi * gammapsi + 1.0;
  const double eta = 0.5 * dde * E2W / w_cr * pow(g2p1, 1.5);
  const double c0 = (-0.5 * SQ3) / PI_;  // imaginary part of -0.5j * SQ3 / PI
  const double sp = (c0 * gamma * dde * E2W) * (1.0 / w_cr) * g2p1;  // Im(ampSP)
  double k13, k23;
  bessel_k13_k23(eta, k13, k23);
  double aS = sp * k23;                                       // Im(ampS)
  double aP = (-(gammapsi * sp)) * k13 * (1.0 / sqrt(g2p1));  // Re(ampP)
  if (!finite_d(aS)) aS = 0.0;
  if (!finite_d(aP)) aP = 0.0;
  const double bwFact = p.bw_mode ? 0.001 : 1.0 / dde;
  const double Amp2Flux = FINE_STR * bwFact * p.eI / SIE0 * 2.0 * p.Np;
  const double mS = fabs(aS), mP = fabs(aP);
  I = Amp2Flux * (mS * mS + mP * mP);
  const double sq = sqrt(Amp2Flux);
  js = cplx{0.0, sq * aS};
  jp = cplx{sq * aP, 0.0};
}

__global__ void __launch_bounds__(256) bm_imap_kernel(size_t n, xrt_bm_params p, const double* __restrict__ gamma,
                                                      const double* __restrict__ w, const double* __restrict__ theta,
                                                      const double* __restrict__ psi, double* __restrict__ Iout,
                                                      double2* __restrict__ Es, double2* __restrict__ Ep) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double I;
  cplx js, jp;
  bm_sample(p, gamma ? gamma[i] : p.gamma0, w[i], theta[i], psi[i], I, js, jp);
  Iout[i] = I;
  Es[i] = make_double2(js.re, js.im);
  Ep[i] = make_double2(jp.re, jp.im);
}

// Beam assembly of BendingMagnet.shine (synchr.py:370-480) for one accepted sample.  g1..g5: standard normal
// deviates (horizontal divergence, 1/gamma opening [bending magnet only], vertical divergence, horizontal size,
// vertical size); kint: the period index np.random.randint(-Np, Np) of a wiggler.
__device__ __forceinline__ void bm_fill_ray(const xrt_source_params& p, const xrt_bm_params& bm, double E,
                                            double Theta0, double Psi0, cplx js, cplx jp, double g1, double g2,
                                            double g3, double g4, double g5, double kint, const xrt_beam& out,
                                            size_t o) {
  const double PI_ = 3.1415926535897932384626433832795, PI2 = 6.283185307179586476925286766559;
  double dtheta = 0.0, dpsi = 0.0;
  if (p.dxprime > 0.0) dtheta = 0.0 + p.dxprime * g1;
  if (!bm.isMPW) dtheta += 0.0 + 1.0 / bm.gamma0 * g2;
  if (p.dzprime > 0.0) dpsi = 0.0 + p.dzprime * g3;
  double a = tan(Theta0 + dtheta), b = 1.0, c = tan(Psi0 + dpsi);
  const double intensS = cmul_np3(js, cconj(js)).re;
  const double intensP = cmul_np3(jp, cconj(jp)).re;
  const double sSP = p.uniform ? 1.0 : intensS + intensP;
  double x, y, z = 0.0;
  cplx Jsp{0.0, 0.0};
  if (bm.isMPW) {
    const double CHeVcm = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12;
    const double sigma_r2 = 2.0 * (CHeVcm / E * 10.0 * bm.L0 * bm.Np) / (PI2 * PI2);
    const double SX = sqrt(p.dx * p.dx + sigma_r2), SZ = sqrt(p.dz * p.dz + sigma_r2);
    double thetaArg = Theta0 * bm.gamma0 / bm.K;
    if (thetaArg < -1.0) thetaArg = -1.0;
    if (thetaArg > 1.0) thetaArg = 1.0;
    y = ((acos(thetaArg) / PI_) + kint - 0.5) * 0.5 * bm.L0;
    x = bm.X0 * sin(PI2 * y / bm.L0) + (0.0 + SX * g4);
    z = 0.0 + SZ * g5;
  } else {
    if (p.dz > 0.0) z = 0.0 + p.dz * g5;
    const double R1 = (p.dx > 0.0) ? bm.ro * 1e3 + p.dx * g4 : bm.ro * 1e3;
    x = -R1 * cos(Theta0) + bm.ro * 1000.0;
    y = R1 * sin(Theta0);
    if (sSP != 0.0) {
      const cplx jj = cmul_np3(js, cconj(jp));
      const double scl = 1.0 / sSP;  // complex / real in numpy: multiply by the reciprocal
      Jsp = cplx{jj.re * scl, jj.im * scl};
    }
  }
  double Jss = 0.0, Jpp = 0.0;
  if (sSP != 0.0) {
    Jss = intensS / sSP;
    Jpp = intensP / sSP;
  }
  const double norm = sqrt(a * a + 1.0 + c * c);
  a /= norm;
  b /= norm;
  c /= norm;
  if (p.has_yaw) {
    const double x1 = p.cosYaw * x - p.sinYaw * y, y1 = p.sinYaw * x + p.cosYaw * y;
    x = x1;
    y = y1;
    const double a1 = p.cosYaw * a - p.sinYaw * b, b1 = p.sinYaw * a + p.cosYaw * b;
    a = a1;
    b = b1;
  }
  if (p.has_pitch) {
    const double y1 = p.cosPitch * y - p.sinPitch * z, z1 = p.sinPitch * y + p.cosPitch * z;
    y = y1;
    z = z1;
    const double b1 = p.cosPitch * b - p.sinPitch * c, c1 = p.sinPitch * b + p.cosPitch * c;
    b = b1;
    c = c1;
  }
  if (p.to_global) {
    if (p.sinAz != 0.0) {
      const double a1 = p.cosAz * a - (-p.sinAz) * b, b1 = (-p.sinAz) * a + p.cosAz * b;
      a = a1;
      b = b1;
      const double x1 = p.cosAz * x - (-p.sinAz) * y, y1 = (-p.sinAz) * x + p.cosAz * y;
      x = x1;
      y = y1;
    }
    x += p.center[0];
    y += p.center[1];
    z += p.center[2];
  }
  out.x[o] = x;
  out.y[o] = y;
  out.z[o] = z;
  out.a[o] = a;
  out.b[o] = b;
  out.c[o] = c;
  out.path[o] = 0.0;
  out.E[o] = E;
  out.Jss[o] = Jss;
  out.Jpp[o] = Jpp;
  reinterpret_cast<double2*>(out.Jsp)[o] = make_double2(Jsp.re, Jsp.im);
  if (out.Es) {
    reinterpret_cast<double2*>(out.Es)[o] = make_double2(js.re, js.im);
    reinterpret_cast<double2*>(out.Ep)[o] = make_double2(jp.re, jp.im);
  }
  out.state[o] = 1;
}

__global__ void __launch_bounds__(256)
    bm_fill_kernel(size_t n, xrt_source_params p, xrt_bm_params bm, const long long* __restrict__ idx,
                   const double* __restrict__ rE, const double* __restrict__ rTheta, const double* __restrict__ rPsi,
                   const double2* __restrict__ mJs, const double2* __restrict__ mJp, const double* __restrict__ g1,
                   const double* __restrict__ g2, const double* __restrict__ g3, const double* __restrict__ g4,
                   const double* __restrict__ g5, const double* __restrict__ kint, xrt_beam out, size_t out_offset) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t j = idx ? (size_t)idx[i] : i;
  const double2 s2 = mJs[j], p2 = mJp[j];
  bm_fill_ray(p, bm, rE[j], rTheta[j], rPsi[j], cplx{s2.x, s2.y}, cplx{p2.x, p2.y}, g1[j], g2[j], g3[j], g4[j],
              g5[j], kint ? kint[j] : 0.0, out, out_offset + i);
}

// uniformRayDensity in one kernel: Philox sampling -> build_I_map -> Beam (no sample arrays in HBM)
__global__ void __launch_bounds__(256)
    bm_shine_uniform_kernel(size_t n, uint64_t seed, uint64_t index_offset, double Emin, double Emax, double Tmin,
                            double Tmax, double Pmin, double Pmax, xrt_source_params sp, xrt_bm_params bm,
                            xrt_beam out, unsigned long long* imax_slot) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double imax = 0.0;
  if (i < n) {
    const Draw d = draw_ray(index_offset + i, seed, Emin, Emax, Tmin, Tmax, Pmin, Pmax, true, false);
    // a second counter range (stream ids 5, 6) feeds the fifth normal deviate and the wiggler's period index
    uint32_t w[4];
    philox4x32_10(index_offset + i, 5u, seed, w);
    const double e0 = 1.0 - u53(w[0], w[1]), e1 = u53(w[2], w[3]);
    double sn, cs;
    sincospi(2.0 * e1, &sn, &cs);
    const double re = sqrt(-2.0 * log(e0));
    const double g5 = re * cs, g6 = re * sn;
    philox4x32_10(index_offset + i, 6u, seed, w);
    const double kint = floor(u53(w[0], w[1]) * (2.0 * floor(bm.Np))) - floor(bm.Np);
    const double gamma = (sp.eEspread > 0.0) ? bm.gamma0 + (0.0 + bm.gamma0 * sp.eEspread * g6) : bm.gamma0;
    double I;
    cplx js, jp;
    bm_sample(bm, gamma, d.E, d.theta, d.psi, I, js, jp);
    imax = I;
    bm_fill_ray(sp, bm, d.E, d.theta, d.psi, js, jp, d.n1, d.n2, d.n3, d.n4, g5, kint, out, i);
  }
  imax = warp_reduce(imax, OpMax());
  if ((threadIdx.x & 31) == 0 && imax > 0.0) atomicMax(imax_slot, (unsigned long long)__double_as_longlong(imax));
}

int bm_imap_device(size_t n, const xrt_bm_params* bm, const double* gamma, const double* w, const double* theta,
                   const double* psi, double* I, double* Es, double* Ep, cudaStream_t st) {
  if (n == 0) return 0;
  XRTB_REQUIRE(bm && w && theta && psi && I && Es && Ep, "bm_imap: NULL argument");
  bm_imap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, *bm, gamma, w, theta, psi, I,
                                                             reinterpret_cast<double2*>(Es),
                                                             reinterpret_cast<double2*>(Ep));
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int bm_fill_device(size_t n, const xrt_source_params* p, const xrt_bm_params* bm, const long long* idx,
                   const double* rE, const double* rTheta, const double* rPsi, const double* mJs, const double* mJp,
                   const double* g1, const double* g2, const double* g3, const double* g4, const double* g5,
                   const double* kint, const xrt_beam* out, size_t out_offset, cudaStream_t st) {
  if (n == 0) return 0;
  XRTB_REQUIRE(p && bm && rE && rTheta && rPsi && mJs && mJp && g1 && g2 && g3 && g4 && g5 && out,
               "bm_fill: NULL argument");
  XRTB_REQUIRE(!bm->isMPW || kint, "bm_fill: a wiggler needs the period indices");
  bm_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
      n, *p, *bm, idx, rE, rTheta, rPsi, reinterpret_cast<const double2*>(mJs), reinterpret_cast<const double2*>(mJp),
      g1, g2, g3, g4, g5, kint, *out, out_offset);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int bm_shine_uniform_device(size_t n, uint64_t seed, uint64_t index_offset, const double* window,
                            const xrt_source_params* sp, const xrt_bm_params* bm, const xrt_beam* out,
                            unsigned long long* imax_slot, cudaStream_t st) {
  if (n == 0) return 0;
  XRTB_REQUIRE(window && sp && bm && out && imax_slot, "bm_shine_uniform: NULL argument");
  XRTB_REQUIRE(sp->uniform, "bm_shine_uniform: the fused kernel is for uniformRayDensity sources");
  bm_shine_uniform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, seed, index_offset, window[0], window[1],
                                                                      window[2], window[3], window[4], window[5], *sp,
                                                                      *bm, *out, imax_slot);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- GeometricSource (sources/geoms.py:193-534)
struct GeomDev {
  const double* d[7];
};

// one standard deviate of stream k for ray i: supplied by the caller or drawn from Philox (uniform, or normal by
// Box-Muller on the (0, 1] / [0, 1) pair of the same block)
__device__ __forceinline__ double geom_deviate(const GeomDev& dv, int k, bool normal, uint64_t ctr, uint64_t seed,
                                               size_t i) {
  if (dv.d[k]) return dv.d[k][i];
  uint32_t w[4];
  philox4x32_10(ctr, 16u + (uint32_t)k, seed, w);
  const double u0 = u53(w[0], w[1]);
  if (!normal) return u0;
  double sn, cs;
  sincospi(2.0 * u53(w[2], w[3]), &sn, &cs);
  return sqrt(-2.0 * log(1.0 - u0)) * cs;
}

struct GeomAmp {
  double Jss, Jpp;
  cplx Jsp, Es, Ep;
};

// GeometricSource._apply_distribution (geoms.py:370-405) for one axis value
__device__ __forceinline__ double geom_axis(const xrt_geom_params& p, int ax, const GeomDev& dv, uint64_t ctr,
                                            uint64_t seed, size_t i, double cur, GeomAmp& A) {
  const int kind = p.dist[ax];
  if (kind == 0) return cur;
  const int stream = ax + 1;
  if (kind == 1) return 0.0 + p.p0[ax] * geom_deviate(dv, stream, true, ctr, seed, i);
  if (kind == 2) return p.p0[ax] + (p.p1[ax] - p.p0[ax]) * geom_deviate(dv, stream, false, ctr, seed, i);
  // uniformRayDensity: uniform in +-cutLim, Gaussian weight on the amplitudes (geoms.py:372-391)
  const double sigma = p.p0[ax], cut = p.p1[ax];
  const double v = -cut + (cut - (-cut)) * geom_deviate(dv, stream, false, ctr, seed, i);
  double amp = exp(-(v * v) / (sigma * sigma) / 2.0) / p.sqrt_pi2 / sigma * 2.0 * cut;
  A.Jss *= amp;
  A.Jpp *= amp;
  A.Jsp = cplx{A.Jsp.re * amp, A.Jsp.im * amp};
  amp = sqrt(amp);
  A.Es = cplx{A.Es.re * amp, A.Es.im * amp};
  A.Ep = cplx{A.Ep.re * amp, A.Ep.im * amp};
  return v;
}

// _set_annulus (geoms.py:409-418)
__device__ __forceinline__ void geom_annulus(const double* an, int s1, int s2, const GeomDev& dv, uint64_t ctr,
                                             uint64_t seed, size_t i, double& v1, double& v2) {
  const double rMin = an[0], rMax = an[1], phiMin = an[2], phiMax = an[3];
  double r = rMax;
  if (rMax > rMin) {
    const double A = 2.0 / (rMax * rMax - rMin * rMin);
    const double u = 0.0 + (1.0 - 0.0) * geom_deviate(dv, s1, false, ctr, seed, i);
    r = sqrt(2.0 * u / A + rMin * rMin);
  }
  const double phi = phiMin + (phiMax - phiMin) * geom_deviate(dv, s2, false, ctr, seed, i);
  v1 = r * cos(phi);
  v2 = r * sin(phi);
}

__global__ void __launch_bounds__(256) geom_sample_kernel(size_t n, xrt_geom_params p, uint64_t seed,
                                                          uint64_t index_offset, GeomDev dv, xrt_beam out,
                                                          unsigned long long* flag) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool over = false;
  if (i < n) {
    const uint64_t ctr = index_offset + i;
    GeomAmp A;
    A.Jss = p.Jss;
    A.Jpp = p.Jpp;
    A.Jsp = cplx{p.Jsp_re, p.Jsp_im};
    A.Es = cplx{p.Es_re, p.Es_im};
    A.Ep = cplx{p.Ep_re, p.Ep_im};
    if (p.ep_random && p.with_amp) A.Ep = cplx{geom_deviate(dv, 0, false, ctr, seed, i) * 0.70710678118654757, 0.0};
    // Beam() starts with x = y = z = a = c = 0, b = 1 (sources/beams.py:60-75)
    double x = 0.0, y = 0.0, z = 0.0, a = 0.0, c = 0.0;
    y = geom_axis(p, 0, dv, ctr, seed, i, y, A);
    if (p.annulus_xz) {
      geom_annulus(p.ann_xz, 2, 3, dv, ctr, seed, i, x, z);
    } else {
      x = geom_axis(p, 1, dv, ctr, seed, i, x, A);
      z = geom_axis(p, 2, dv, ctr, seed, i, z, A);
    }
    if (p.annulus_ac) {
      geom_annulus(p.ann_ac, 4, 5, dv, ctr, seed, i, a, c);
    } else {
      a = geom_axis(p, 3, dv, ctr, seed, i, a, A);
      c = geom_axis(p, 4, dv, ctr, seed, i, c, A);
    }
    double E = 0.0;
    if (p.distE == 1) {
      E = p.e0 + p.e1 * geom_deviate(dv, 6, true, ctr, seed, i);
    } else if (p.distE == 2) {
      E = p.e0 + (p.e1 - p.e0) * geom_deviate(dv, 6, false, ctr, seed, i);
    } else if (p.distE == 3) {
      int k;
      if (dv.d[6]) {
        k = (int)dv.d[6][i];  // the caller supplies the line index
      } else {
        const double u = geom_deviate(dv, 6, false, ctr, seed, i);
        if (p.weighted) {  // np.random.choice(p=...): cdf.searchsorted(u, side='right')
          k = 0;
          while (k < p.n_lines - 1 && p.cdf[k] <= u) ++k;
        } else {
          k = (int)(u * p.n_lines);
        }
      }
      if (k < 0) k = 0;
      if (k >= p.n_lines) k = p.n_lines - 1;
      E = p.lines[k];
    }
    over = (a * a + c * c) > 1.0;
    out.x[i] = x;
    out.y[i] = y;
    out.z[i] = z;
    out.a[i] = a;
    out.b[i] = 1.0;
    out.c[i] = c;
    out.path[i] = 0.0;
    out.E[i] = E;
    out.Jss[i] = A.Jss;
    out.Jpp[i] = A.Jpp;
    reinterpret_cast<double2*>(out.Jsp)[i] = make_double2(A.Jsp.re, A.Jsp.im);
    if (out.Es) {
      reinterpret_cast<double2*>(out.Es)[i] = make_double2(A.Es.re, A.Es.im);
      reinterpret_cast<double2*>(out.Ep)[i] = make_double2(A.Ep.re, A.Ep.im);
    }
    out.state[i] = 1;
  }
  const unsigned ball = __ballot_sync(0xffffffffu, over);
  if ((threadIdx.x & 31) == 0 && ball) atomicAdd(flag, (unsigned long long)__popc(ball));
}

// direction normalisation (geoms.py:503-511: one formula for the whole beam), rotate_beam, virgin_local_to_global
__global__ void __launch_bounds__(256) geom_finish_kernel(size_t n, xrt_geom_params p, xrt_beam out,
                                                          const unsigned long long* flag) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = out.x[i], y = out.y[i], z = out.z[i], a = out.a[i], c = out.c[i], b;
  const double ac = a * a + c * c;
  if (*flag > 0) {
    b = sqrt(ac + 1.0);
    a /= b;
    c /= b;
    b = 1.0 / b;
  } else {
    b = sqrt(1.0 - ac);
  }
  for (int k = 0; k < p.n_rot; ++k) {
    const double cA = p.rot_cos[k], sA = p.rot_sin[k];
    if (p.rot_axis[k] == 2) {  // rotate_z(x, y)
      const double x1 = cA * x - sA * y, y1 = sA * x + cA * y, a1 = cA * a - sA * b, b1 = sA * a + cA * b;
      x = x1, y = y1, a = a1, b = b1;
    } else if (p.rot_axis[k] == 1) {  // rotate_y(x, z)
      const double x1 = cA * x + sA * z, z1 = -sA * x + cA * z, a1 = cA * a + sA * c, c1 = -sA * a + cA * c;
      x = x1, z = z1, a = a1, c = c1;
    } else {  // rotate_x(y, z)
      const double y1 = cA * y - sA * z, z1 = sA * y + cA * z, b1 = cA * b - sA * c, c1 = sA * b + cA * c;
      y = y1, z = z1, b = b1, c = c1;
    }
  }
  if (p.to_global) {
    if (p.sinAz != 0.0) {
      const double a1 = p.cosAz * a - (-p.sinAz) * b, b1 = (-p.sinAz) * a + p.cosAz * b;
      a = a1;
      b = b1;
      const double x1 = p.cosAz * x - (-p.sinAz) * y, y1 = (-p.sinAz) * x + p.cosAz * y;
      x = x1;
      y = y1;
    }
    x += p.center[0];
    y += p.center[1];
    z += p.center[2];
  }
  out.x[i] = x;
  out.y[i] = y;
  out.z[i] = z;
  out.a[i] = a;
  out.b[i] = b;
  out.c[i] = c;
}

int geometric_source_device(size_t n, const xrt_geom_params* p, uint64_t seed, uint64_t index_offset,
                            const double* const* deviates, const xrt_beam* out, unsigned long long* flag, int phase,
                            cudaStream_t st) {
  XRTB_REQUIRE(p && out && flag, "geometric_source: NULL argument");
  XRTB_REQUIRE(p->n_lines >= 0 && p->n_lines <= XRT_GEOM_MAX_LINES, "geometric_source: at most %d energy lines",
               XRT_GEOM_MAX_LINES);
  XRTB_REQUIRE(p->distE != 3 || p->n_lines > 0, "geometric_source: distE = 'lines' needs at least one energy");
  XRTB_REQUIRE(p->n_rot >= 0 && p->n_rot <= 3, "geometric_source: at most 3 elementary rotations");
  GeomDev dv;
  for (int k = 0; k < 7; ++k) dv.d[k] = deviates ? deviates[k] : nullptr;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (phase & 1) {  // (the caller zeroes `flag`)
    if (n > 0) geom_sample_kernel<<<grid, 256, 0, st>>>(n, *p, seed, index_offset, dv, *out, flag);
    XRTB_CUDA_CHECK(cudaGetLastError());
  }
  if ((phase & 2) && n > 0) {
    geom_finish_kernel<<<grid, 256, 0, st>>>(n, *p, *out, flag);
    XRTB_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

int source_sample_device(size_t n, uint64_t seed, uint64_t index_offset, const double* window, double* E,
                         double* theta, double* psi, double* n1, double* n2, double* n3, double* n4, double* u,
                         double* n5, cudaStream_t st) {
  if (n == 0) return 0;
  XRTB_REQUIRE(window && E && theta && psi, "source_sample: NULL argument");
  XRTB_REQUIRE((n1 == nullptr) == (n2 == nullptr) && (n1 == nullptr) == (n3 == nullptr) &&
                   (n1 == nullptr) == (n4 == nullptr),
               "source_sample: give all four normal-deviate arrays or none");
  sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, seed, index_offset, window[0], window[1], window[2],
                                                            window[3], window[4], window[5], E, theta, psi, n1, n2,
                                                            n3, n4, u, n5);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int source_fill_device(size_t n, const xrt_source_params* p, const long long* idx, const double* rE,
                       const double* rTheta, const double* rPsi, const double* mJs, const double* mJp,
                       const double* n1, const double* n2, const double* n3, const double* n4, const xrt_beam* out,
                       size_t out_offset, cudaStream_t st) {
  if (n == 0) return 0;
  XRTB_REQUIRE(p && rE && rTheta && rPsi && mJs && mJp && n1 && n2 && n3 && n4 && out, "source_fill: NULL argument");
  fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, *p, idx, rE, rTheta, rPsi,
                                                          reinterpret_cast<const double2*>(mJs),
                                                          reinterpret_cast<const double2*>(mJp), n1, n2, n3, n4, *out,
                                                          out_offset);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int shine_uniform_device(size_t n, uint64_t seed, uint64_t index_offset, const double* window, int jend, double Kx,
                         double Ky, double L0, double Np, double gamma0, double amp_const, double dstep, int bw_mode,
                         int symmetric, const double* tg, const double* ag, const double* sintg, const double* costg,
                         const double* sintgph, const double* costgph, const xrt_source_params* sp,
                         const xrt_beam* out, unsigned long long* imax_slot, cudaStream_t st) {
  if (n == 0) return 0;
  XRTB_REQUIRE(window && sp && out && imax_slot, "shine_uniform: NULL argument");
  XRTB_REQUIRE(jend > 0, "shine_uniform: the integration grid is empty");
  XRTB_REQUIRE(sp->uniform, "shine_uniform: the fused kernel is for uniformRayDensity sources");
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
#ifndef XRTB_SHINE_R
#define XRTB_SHINE_R 2  // samples per thread (3 and 4 measured: profiles/r02_shine_variants.json)
#endif
  constexpr int R = XRTB_SHINE_R;
  const size_t per_block = (size_t)U_THREADS * R * XRTB_SHINE_BATCHES;
  const unsigned grid = (unsigned)((n + per_block - 1) / per_block);
  const bool planar = (Kx == 0.0);
#define XRTB_S_LAUNCH(PL, SY)                                                                                    \
  shine_uniform_kernel<R, PL, SY><<<grid, U_THREADS, 0, st>>>(n, seed, index_offset, window[0], window[1], window[2], \
                                                             window[3], window[4], window[5], p, t, *sp, *out, imax_slot)
#ifdef XRTB_SHINE_NO_GC  // (build switch for A / B comparisons: the per-sample evaluation everywhere)
  const bool gconst = false;
#else
  const bool gconst = !(sp->eEspread > 0.0);
#endif
  if (planar && symmetric && gconst)
    // one gamma for every sample: the gamma-dependent node quantities come from the node tile (GC)
    shine_uniform_kernel<R, true, true, true><<<grid, U_THREADS, 0, st>>>(n, seed, index_offset, window[0], window[1],
                                                                          window[2], window[3], window[4], window[5],
                                                                          p, t, *sp, *out, imax_slot);
  else if (planar && symmetric)
    XRTB_S_LAUNCH(true, true);
  else if (planar)
    XRTB_S_LAUNCH(true, false);
  else
    XRTB_S_LAUNCH(false, false);
#undef XRTB_S_LAUNCH
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace xrtb
