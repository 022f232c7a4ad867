// Shared device code of the undulator field integral (used by undulator.cu and by the fused source kernel in
// source.cu).  Everything here is written with explicit fma() / __dmul_rn / __dadd_rn, so the results do not
// depend on the -fmad setting of the including translation unit.  See undulator.cu for the derivation.
#pragma once
#include "common.cuh"

namespace xrtb {

struct __align__(16) UNode {
  double tg, ag, c1, c2;   // node, weight, cos(tg), cos(tg+phase)
  double p1, p2, qn, pad;  // -Ky sin(tg), Kx sin(tg+phase), Ky^2 sin(2tg) + Kx^2 sin(2(tg+phase))
  // r02: what the pair step derives from the node and gamma ALONE, when every sample has the same gamma (no energy
  // spread): beta_x, beta_z, beta'_z (synchr.py:2006-2016) -- computed once per node with the per-sample operations
  // (same bits), 6 of the 44 FP64 instructions of a (sample, node pair)
  double bx, bz, bpz, pad2;
};

constexpr int U_THREADS = 128;
constexpr int U_TILE = 256;  // nodes per shared-memory tile (24 KB)

struct USample {
  double ww1, b1, b2, a3;      // phase coefficients
  double kyr, kxrn, revg2, hrevg;
  double dirx, diry, dirz;
};

// per-sample prefactors: quotients / roots of ordinary positive numbers (gamma, photon energy, wu): the compiler's IEEE
// sequences without their range test (common.cuh: same bits); -DXRTB_SOURCE_DIV_CHECKED restores the plain operators
#ifdef XRTB_SOURCE_DIV_CHECKED
#define XRTB_U_DIV(a, b) ((a) / (b))
#define XRTB_U_SQRT(a) sqrt(a)
#else
#define XRTB_U_DIV(a, b) div_nr(a, b)
#define XRTB_U_SQRT(a) sqrt_nr(a)
#endif

__device__ __forceinline__ USample make_sample(double Kx, double Ky, double gamma, double wu, double w, double ww1,
                                               double phi, double psi) {
  USample s;
  const double revg = XRTB_U_DIV(1.0, gamma);  // synchr.py:1936
  s.revg2 = __dmul_rn(revg, revg);
  const double wwu = XRTB_U_DIV(w, wu);
  const double wr = wwu * revg;
  s.ww1 = ww1;
  s.b1 = wr * phi;
  s.b2 = wr * psi;
  s.a3 = wr * (0.125 * revg);
  s.kyr = __dmul_rn(Ky, revg);      // taperC*Ky*revgamma, taperC = 1 (synchr.py:2006)
  s.kxrn = __dmul_rn(-Kx, revg);    // -Kx*revgamma (synchr.py:2007)
  s.hrevg = 0.5 * revg;
  s.dirx = phi;
  s.diry = psi;
  // dirz = 1 - 0.5*(phi**2 + psi**2)  (synchr.py:1945)
  s.dirz = __dadd_rn(1.0, -0.5 * __dadd_rn(__dmul_rn(phi, phi), __dmul_rn(psi, psi)));
  return s;
}

// PLANAR (Kx == 0): beta_y, beta'_y and every product with them are exact zeros in the reference's arithmetic
// (adding/subtracting +-0 is exact), so dropping those operations leaves every result bit-identical.
template <bool PLANAR>
__device__ __forceinline__ void node_step(const USample& s, const double4 n0, const double4 n1, double& bs_re,
                                          double& bs_im, double& bp_re, double& bp_im) {
  // phase (synchr.py:1997-2002), reassociated: every term is O(1..100) rad, error ~1e-14 rad
  double ucos = s.ww1 * n0.x;
  ucos = fma(s.b1, n1.x, ucos);
  if (!PLANAR) ucos = fma(s.b2, n1.y, ucos);
  ucos = fma(s.a3, n1.z, ucos);
  double sn, cs;
  sincos_big(ucos, sn, cs);
  // velocity (synchr.py:2006-2008), reference rounding order
  const double bx = __dmul_rn(s.kyr, n0.z);
  const double by = PLANAR ? 0.0 : __dmul_rn(s.kxrn, n0.w);
  double t4 = __dadd_rn(s.revg2, __dmul_rn(bx, bx));
  if (!PLANAR) t4 = __dadd_rn(t4, __dmul_rn(by, by));
  const double bz = __dadd_rn(1.0, -0.5 * t4);
  // acceleration (synchr.py:2010-2016)
  const double bpz = __dmul_rn(s.hrevg, n1.z);
  // 1/(1 - n.beta) (synchr.py:2018), reference rounding order
  double den = __dadd_rn(1.0, -__dmul_rn(s.dirx, bx));
  if (!PLANAR) den = __dadd_rn(den, -__dmul_rn(s.diry, by));
  den = __dadd_rn(den, -__dmul_rn(s.dirz, bz));
  const double rk = rcp_cubic(den);
  const double wgt = n0.y * (rk * rk);  // ag * rkrel**2 (synchr.py:2019)
  const double ere = cs * wgt, eim = sn * wgt;
  const double bnx = s.dirx - bx;
  const double bny = PLANAR ? s.diry : s.diry - by;
  const double bnz = s.dirz - bz;  // exact (Sterbenz)
  const double ddbp = PLANAR ? fma(s.dirz, bpz, s.dirx * n1.x) : fma(s.dirz, bpz, fma(s.diry, n1.y, s.dirx * n1.x));
  const double ddmb = fma(s.dirz, bnz, fma(s.diry, bny, s.dirx * bnx));
  const double ts = fma(bnx, ddbp, -(n1.x * ddmb));
  const double tp = PLANAR ? bny * ddbp : fma(bny, ddbp, -(n1.y * ddmb));
  bs_re = fma(ere, ts, bs_re);
  bs_im = fma(eim, ts, bs_im);
  bp_re = fma(ere, tp, bp_re);
  bp_im = fma(eim, tp, bp_im);
}

// Mirror-pair step for a planar undulator on a grid that is symmetric about tg = 0 (Clenshaw-Curtis nodes on
// sub-intervals of [-pi, pi] are): under tg -> -tg the phase u and the bracket [..] are odd while 1/(1-n.beta)^2
// and the weight are even, so node j and its mirror contribute  w [..] (e^{iu} - e^{-iu}) = 2i w [..] sin(u).
// One sin (no cos), one velocity/denominator evaluation per PAIR of nodes; the sums are purely imaginary.
// The tables are symmetric to a few ulps only (np.sin/np.cos of mirrored arguments), which moves u by ~1e-15
// rad: orders below the parity tolerance.  The factor 2 is applied once at the end.
template <bool GC = false>
__device__ __forceinline__ void node_pair_step(const USample& s, const double4 n0, const double4 n1, const double4 n2,
                                               double& bs_im, double& bp_im, const double2* __restrict__ sctab) {
  double ucos = s.ww1 * n0.x;
  ucos = fma(s.b1, n1.x, ucos);
  ucos = fma(s.a3, n1.z, ucos);
#ifdef XRTB_UNDULATOR_SIN_POLY
  const double sn = sin_big(ucos);
#else
  const double sn = sin_tab(ucos, sctab);  // 13 instead of 17 FP64-pipe instructions (common.cuh)
#endif
  double bx, bz, bpz;
  if (GC) {  // the same gamma for every sample: taken from the node tile (load_node_tile)
    bx = n2.x;
    bz = n2.y;
    bpz = n2.z;
  } else {
    bx = __dmul_rn(s.kyr, n0.z);
    const double t4 = __dadd_rn(s.revg2, __dmul_rn(bx, bx));
    bz = __dadd_rn(1.0, -0.5 * t4);
    bpz = __dmul_rn(s.hrevg, n1.z);
  }
  double den = __dadd_rn(1.0, -__dmul_rn(s.dirx, bx));
  den = __dadd_rn(den, -__dmul_rn(s.dirz, bz));
  const double rk = rcp_cubic(den);
  const double wgt = n0.y * (rk * rk);
  const double eim = sn * wgt;
  const double bnx = s.dirx - bx;
  const double bnz = s.dirz - bz;
  const double ddbp = fma(s.dirz, bpz, s.dirx * n1.x);
  const double ddmb = fma(s.dirz, bnz, fma(s.diry, s.diry, s.dirx * bnx));
  const double ts = fma(bnx, ddbp, -(n1.x * ddmb));
  const double tp = s.diry * ddbp;
  bs_im = fma(eim, ts, bs_im);
  bp_im = fma(eim, tp, bp_im);
}

// node-loop unrolling: 1 = one node (pair) per iteration for the R samples of the thread (measured best)
#ifndef XRTB_NODE_UNROLL
#define XRTB_NODE_UNROLL 1
#endif
constexpr int kNodeUnroll = XRTB_NODE_UNROLL;

struct UTables {
  const double *tg, *ag, *sintg, *costg, *sintgph, *costgph;
  const double2* sintab;  // (cos, sin)(k pi/512), k < 1024, in device memory (sin_table_device): staged per block
};

// gamma0 > 0: every sample has this gamma -- the gamma-dependent node quantities of the pair step are filled in too
__device__ __forceinline__ void load_node_tile(UNode* tile, const UTables& t, int j0, int cnt, double Kx, double Ky,
                                               double gamma0 = 0.0) {
  const double Ky2 = Ky * Ky, Kx2 = Kx * Kx;
  // make_sample's operations on gamma (revg, revg2, kyr, hrevg), then node_pair_step's on the node
  const double revg = gamma0 > 0.0 ? XRTB_U_DIV(1.0, gamma0) : 0.0;
  const double revg2 = __dmul_rn(revg, revg), kyr = __dmul_rn(Ky, revg), hrevg = 0.5 * revg;
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
    const int g = j0 + j;
    const double st = t.sintg[g], ct = t.costg[g], sp = t.sintgph[g], cp = t.costgph[g];
    UNode n;
    n.tg = t.tg[g];
    n.ag = t.ag[g];
    n.c1 = ct;
    n.c2 = cp;
    n.p1 = -Ky * st;
    n.p2 = Kx * sp;
    const double sin2x = __dmul_rn(__dmul_rn(2.0, st), ct);    // synchr.py:1934
    const double sin2xph = __dmul_rn(__dmul_rn(2.0, sp), cp);  // synchr.py:1935
    n.qn = __dadd_rn(__dmul_rn(Ky2, sin2x), __dmul_rn(Kx2, sin2xph));
    n.pad = 0.0;
    n.bx = __dmul_rn(kyr, n.c1);
    n.bz = __dadd_rn(1.0, -0.5 * __dadd_rn(revg2, __dmul_rn(n.bx, n.bx)));
    n.bpz = __dmul_rn(hrevg, n.qn);
    n.pad2 = 0.0;
    tile[j] = n;
  }
}

struct UParams {
  double Kx, Ky, L0, Np, gamma0, amp_const, dstep;
  int jend;
  int bw_mode;  // 1: distE == 'BW'
};


// Per-sample prefactors of Undulator._build_I_map_conv (synchr.py:2060-2071): wu, ww1 and
// ab = sin(pi N ww1) / sin(pi ww1) / (2 pi wu), reference operation order.
__device__ __forceinline__ void sample_prefactors(const UParams& p, double g, double wi, double ph, double ps,
                                                  double& wui, double& ww1i, double& ab) {
  const double PI_ = 3.1415926535897932384626433832795;
  const double g2 = g * g;
  const double kx2h = 0.5 * (p.Kx * p.Kx), ky2h = 0.5 * (p.Ky * p.Ky);
  wui = XRTB_U_DIV(XRTB_U_DIV(XRTB_U_DIV(PI_, p.L0), g2) * (2.0 * g2 - 1.0 - kx2h - ky2h), 5067.7309392068091);
  ww1i = XRTB_U_DIV(wi * ((1.0 + kx2h + ky2h) + g2 * (ph * ph + ps * ps)), 2.0 * g2 * wui);
  // (the last quotient keeps the IEEE operator: sin(pi ww1) is zero at the harmonics -> inf / nan like numpy)
  ab = XRTB_U_DIV(1.0 / 6.283185307179586476925286766559, wui) * sin(PI_ * p.Np * ww1i) / sin(PI_ * ww1i);
}

// I, Es, Ep of synchr.py:2092-2108 from the raw node sums (already multiplied by wu/gamma)
__device__ __forceinline__ void imap_outputs(const UParams& p, double w, double ab, double is_re, double is_im,
                                             double ip_re, double ip_im, double& I, double2& Es, double2& Ep) {
  const double bw = p.bw_mode ? 0.001 : XRTB_U_DIV(1.0, w);
  const double amp2flux = p.amp_const * bw;  // FINE_STR * bwFact * eI / SIE0
  // np.abs of a purely imaginary sum (the node-pair path of a planar undulator) is |im| exactly
  const double as = (is_re == 0.0) ? fabs(is_im) : hypot(is_re, is_im);
  const double ap = (ip_re == 0.0) ? fabs(ip_im) : hypot(ip_re, ip_im);
  const double field = as * as + ap * ap;
  I = amp2flux * (ab * ab) * 0.25 * (p.dstep * p.dstep) * field;
  const double sc = XRTB_U_SQRT(amp2flux) * ab;
  const double hd = 0.5 * p.dstep;
  Es = make_double2(sc * is_re * hd, sc * is_im * hd);
  Ep = make_double2(sc * ip_re * hd, sc * ip_im * hd);
}

// The node loop for R samples of one thread: fills acc[r][0..3] = (Bs.re, Bs.im, Bp.re, Bp.im).
// Must be called by every thread of the block (it synchronises on the shared node tile).
template <int R, bool PLANAR, bool SYM, bool GC = false>
__device__ __forceinline__ void node_loop(const UParams& p, const UTables& tab, UNode* tile, double2* sctab,
                                          const USample* s, double (*acc)[4], bool first = true) {
  // first = false: a further batch of samples in the same block -- the sin table is staged already, and so is the
  // node tile when the grid fits one tile (the usual case: 64 nodes)
  // SYM: only the first half of the nodes is visited (pairs); an odd middle node goes through the general step
  const int jpair = SYM ? p.jend / 2 : 0;
  const int jlast = SYM ? jpair + (p.jend & 1) : p.jend;
  if (SYM && first) {  // the sin table of the pair step: 16 KB from L2 into shared memory (published by the barrier below)
    for (int k = threadIdx.x; k < kSinTabEntries; k += blockDim.x) sctab[k] = tab.sintab[k];
  }
  const bool reload = first || jlast > U_TILE;
  for (int j0 = 0; j0 < jlast; j0 += U_TILE) {
    const int cnt = min(U_TILE, jlast - j0);
    if (reload) {
      __syncthreads();
      load_node_tile(tile, tab, j0, cnt, p.Kx, p.Ky, GC ? p.gamma0 : 0.0);
      __syncthreads();
    }
#pragma unroll kNodeUnroll
    for (int j = 0; j < cnt; ++j) {
      const double4* q = reinterpret_cast<const double4*>(tile + j);
      const double4 n0 = q[0];
      const double4 n1 = q[1];
      if (SYM && (j0 + j) < jpair) {
        const double4 n2 = GC ? q[2] : make_double4(0.0, 0.0, 0.0, 0.0);
#pragma unroll
        for (int r = 0; r < R; ++r) node_pair_step<GC>(s[r], n0, n1, n2, acc[r][1], acc[r][3], sctab);
      } else if (SYM) {
        // middle node of an odd grid (tg = 0): counts once, i.e. half of the factor 2 applied below
#pragma unroll
        for (int r = 0; r < R; ++r) {
          double h[4] = {0.0, 0.0, 0.0, 0.0};
          node_step<PLANAR>(s[r], n0, n1, h[0], h[1], h[2], h[3]);
          acc[r][0] += 0.5 * h[0];
          acc[r][1] += 0.5 * h[1];
          acc[r][2] += 0.5 * h[2];
          acc[r][3] += 0.5 * h[3];
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) node_step<PLANAR>(s[r], n0, n1, acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      }
    }
  }
  if (SYM) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      acc[r][0] *= 2.0;
      acc[r][1] *= 2.0;
      acc[r][2] *= 2.0;
      acc[r][3] *= 2.0;
    }
  }
}

}  // namespace xrtb
