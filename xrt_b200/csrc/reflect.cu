// Fused ray-surface kernels for sm_100a: OE.reflect / DCM.double_reflect in one pass over the Beam SoA.
//
// Replaces the numpy pipeline OE.reflect -> _reflect_local (xrt/backends/raycing/oes/reflect.py:17-163,
// :551-1139) with everything it calls: global_to_virgin_local / virgin_local_to_global
// (beamline.py:230-287), rotate_beam (_rotate.py:23-57), _bracketing/_set_t (oes/base.py:1231-1295),
// find_intersection / find_dz / _use_my_method / _use_Brent_method (oes/base.py:801-1048, the semantics of
// the dead OpenCL kernel `find_intersection`, cl/OE.cl:313), rays_good (base.py:1094-1163), the surface
// classes (flat base.py:675-742, ToroidMirror oes/__init__.py:398-411, EllipticalMirrorParam
// oes/parametric.py:213-246), the direction / polarisation block (reflect.py:788-1064, :451-469, :568-612),
// rotate_coherency_matrix (sources/beams.py:448-479) and the amplitudes Material.get_amplitude
// (materials/material.py:348-493) / Crystal.get_amplitude thick Bragg (materials/crystal.py:492-645).
// Oracle: oracle/reflect.py, oracle/materials.py (bit-exact against the reference's outputs).
//
// One thread per ray, every Beam array is read once and written once with unit stride (SoA): 132 B in +
// 132 B (gb) + 140 B (lb) per ray with amplitudes.  Two small pre-passes produce the beam-wide scalars the
// numpy code derives with whole-array min/max: phase A (28 B/ray) the dominant direction component and the
// first good ray, phase B (52 B/ray) the global t bracket and the secant-vs-Brent switch for the (axis, sign)
// variant phase A selected.  They live in 64 device-resident slots, so the main pass needs no host round trip
// and ray shards on several GPUs all-reduce the slots in between.
//
// THIS FILE IS COMPILED WITH -fmad=false: numpy rounds every multiply and add separately, and the
// intersection iterates and grazing-incidence Fresnel terms amplify 1-ulp differences (cancellation ~1e6),
// so FMA contraction would break parity.  fma() is used explicitly only where it cannot matter.
#include "../../include/xrt_b200.h"
#include "common.cuh"
#include "kernels.h"

// ---------------------------------------------------------------- spliced surfaces (plug-in builds only)
// The CUDA counterpart of the reference's MY_LOCAL_Z / MY_LOCAL_N / MY_XYZPARAM text splicing
// (oes/base.py:548-569 into cl/OE.cl:60-66; the class attributes cl_plist, cl_local_z, cl_local_n of e.g.
// oes/__init__.py:245-256): xrt_b200/custom.py writes the element's `cu_local_z`, `cu_local_n` (and, for a
// parametric element, `cu_xyz_param`) function texts into a header and compiles THIS file once more with
//   -DXRTB_USER_SURFACE='"<header>"' -DXRTB_PLUGIN [-DXRTB_USER_PARAMETRIC]
// into a small shared object next to libxrt_b200.so.  The header defines, in namespace xrtb_user,
//   __device__ double local_z(const double* cu_plist, double x, double y);            // or local_r(plist, s, phi)
//   __device__ void   local_n(const double* cu_plist, double x, double y, double* n); // n[0..2], or n[0..5] = (n_B, n_s)
//   __device__ void   xyz_to_param(const double* cu_plist, double& x, double& y, double& z);   // parametric only
//   __device__ void   param_to_xyz(const double* cu_plist, double& s, double& phi, double& r); // parametric only
// and, with -DXRTB_USER_INTERSECTION, an element's OWN intersection method instead of the bracketed root search
// (the reference's BlazedGrating overrides OE.find_intersection this way, oes/gratings.py:492-520):
//   __device__ void   find_intersection(const double* cu_plist, double x, double y, double z, double a, double b,
//                                        double c, double& t2, double& x2, double& y2, double& z2);
// and the kernels below see it as surface kind 4 (z(x, y)) or 5 (parametric).  cu_plist = xrt_oe_params.sp.
#ifdef XRTB_USER_SURFACE
#include XRTB_USER_SURFACE
#endif
#if defined(XRTB_PLUGIN_KIND)
#define XRTB_USER_KIND XRTB_PLUGIN_KIND  // a built-in surface compiled as a plug-in variant (figure error, below)
#elif defined(XRTB_USER_PARAMETRIC)
#define XRTB_USER_KIND 5
#else
#define XRTB_USER_KIND 4
#endif
// Figure errors (OE(figureError=...): oes/base.py:681-686, :744-770, find_dz :825-829, reflect.py:767-776,
// figure_error.py:207-265) are a plug-in variant as well: -DXRTB_FIGURE_ERROR compiles the bicubic-spline height map
// into find_dz and the normal of ONE surface kind (-DXRTB_PLUGIN_KIND=<0|1|2> or a spliced kind 4 / 5), so the kernels of
// libxrt_b200.so keep their register budget.

namespace xrtb {

// ---------------------------------------------------------------- reduction slots (uint64, order-preserving)
enum {
  RS_MAXA = 0,  // max |a|, |b|, |c| over rays with state == 1 (base.py:1257-1263)
  RS_MAXB = 1,
  RS_MAXC = 2,
  RS_FIRST = 3,  // min over good rays of (global index << 3 | a>0 | b>0 << 1 | c>0 << 2)  (base.py:1241 `abc[0] > 0`)
  RS_VAR = 4,    // 6 variants (axis * 2 + neg) x {min t1, max t2, max |dz1|, max |dz2| masked} (base.py:861-871)
  RS_FOOT = 28,  // footprint min x,y,z / max x,y,z of the virgin-local output (reflect.py:1133-1136)
  RS_NGOOD = 34, // number of rays with state > 0 on entry (plain integer count)
  RS_NGOODN = 35, // number of rays with state == 1 after the surface (plain integer count); 36: out-of-table count
  RS_SIDE_NEG = 37,  // 37, 38: good rays that met an asymmetric crystal's planes with n_B.k < 0 / >= 0 (sgbdn check)
  RS_DERIV = 40,  // multiple_reflect, derivative stage: {min t1, max t2, max |dz1|, max |dz2| masked} (base.py:861-871)
  RS_DECISION = 44,  // 44..46: the decoded beam-wide decision of the main pass (reflect_decide_kernel)
  RS_HINT = 48    // 48..63: 32 ints, the last f1/f2 table interval found per element table (f1f2_locate)
};
constexpr double kZEps = 1e-12;      // raycing.zEps   (raycing/__init__.py:86)
constexpr double kDt = 1e-5;         // raycing.dt     (:89)
constexpr int kMaxIteration = 100;   // raycing.maxIteration (:92)
constexpr double kDs = 0.0;          // raycing.ds     (:90) margin used in multiple reflections
constexpr double kPi = 3.141592653589793;

__host__ __device__ inline unsigned long long enc(double d) {
  unsigned long long u;
  memcpy(&u, &d, 8);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline double dec(unsigned long long u) {
  u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
  double d;
  memcpy(&d, &u, 8);
  return d;
}

struct Ray {
  double x, y, z, a, b, c;
};

__device__ __forceinline__ void rot_axis(int axis, double cA, double sA, double& x, double& y, double& z) {
  // rotate_x / rotate_y / rotate_z of _rotate.py:5-20
  if (axis == 0) {
    const double y1 = cA * y - sA * z, z1 = sA * y + cA * z;
    y = y1;
    z = z1;
  } else if (axis == 1) {
    const double x1 = cA * x + sA * z, z1 = -sA * x + cA * z;
    x = x1;
    z = z1;
  } else {
    const double x1 = cA * x - sA * y, y1 = sA * x + cA * y;
    x = x1;
    y = y1;
  }
}

__device__ __forceinline__ void to_local(const xrt_oe_params& p, Ray& r) {
  if (p.from_global) {
    r.x -= p.center[0];
    r.y -= p.center[1];
    r.z -= p.center[2];
    if (p.sinAz != 0.0) {
      rot_axis(2, p.cosAz, p.sinAz, r.x, r.y, r.z);
      rot_axis(2, p.cosAz, p.sinAz, r.a, r.b, r.c);
    }
  }
  for (int i = 0; i < p.n_rot_in; ++i) {
    rot_axis(p.rot_in_axis[i], p.rot_in_cos[i], p.rot_in_sin[i], r.x, r.y, r.z);
    rot_axis(p.rot_in_axis[i], p.rot_in_cos[i], p.rot_in_sin[i], r.a, r.b, r.c);
  }
  if (p.shift[0] != 0.0) r.x -= p.shift[0];
  if (p.shift[1] != 0.0) r.y -= p.shift[1];
  if (p.shift[2] != 0.0) r.z -= p.shift[2];
}

__device__ __forceinline__ void to_virgin(const xrt_oe_params& p, Ray& r) {
  if (p.shift[0] != 0.0) r.x += p.shift[0];
  if (p.shift[1] != 0.0) r.y += p.shift[1];
  if (p.shift[2] != 0.0) r.z += p.shift[2];
  for (int i = 0; i < p.n_rot_out; ++i) {
    rot_axis(p.rot_out_axis[i], p.rot_out_cos[i], p.rot_out_sin[i], r.x, r.y, r.z);
    rot_axis(p.rot_out_axis[i], p.rot_out_cos[i], p.rot_out_sin[i], r.a, r.b, r.c);
  }
}

__device__ __forceinline__ void virgin_to_global(const xrt_oe_params& p, Ray& r) {
  if (p.sinAz != 0.0) {
    rot_axis(2, p.cosAz, -p.sinAz, r.a, r.b, r.c);
    rot_axis(2, p.cosAz, -p.sinAz, r.x, r.y, r.z);
  }
  r.x += p.center[0];
  r.y += p.center[1];
  r.z += p.center[2];
}

// ---------------------------------------------------------------- surfaces
// x**(-0.5) as glibc pow returns it (< 1 ulp): 1/sqrt with one residual correction
__device__ __forceinline__ double pow_m05(double x) {
#ifdef XRTB_SURF_DIV_CHECKED
  const double y0 = 1.0 / sqrt(x);
#else
  // 0 < x <= 1 here (toroid: 1 - (x/r)^2); x == 0 keeps the reference's 1 / 0
  const double y0 = (x == 0.0) ? 1.0 / x : div_nr(1.0, sqrt_nr(x));
#endif
  const double e = fma(-x * y0, y0, 1.0);
  return fma(0.5 * y0, e, y0);
}

#ifdef XRTB_FIGURE_ERROR
// ---------------------------------------------------------------- figure error: bicubic spline height map
// scipy's RectBivariateSpline(y1d, x1d, z_nm, kx=3, ky=3).ev(y, x[, dx, dy]) as FITPACK evaluates it (fpbisp with
// fpbspl; parder for the first derivatives): u = y + yShift runs along the first knot vector, v = x + xShift along
// the second; c is (nu - 4) x (nv - 4).  Same operation order as FITPACK (this file is compiled with -fmad=false).
__device__ __forceinline__ int fe_interval(const double* __restrict__ t, int n, int k, double& arg) {
  // fpbisp: clamp to [t(k+1), t(n-k)], then t(l) <= arg < t(l+1) or l = n-k-1 (1-based); returns the 0-based l
  const double tb = t[k], te = t[n - k - 1];
  if (arg < tb) arg = tb;
  if (arg > te) arg = te;
  int lo = k, hi = n - k - 2;  // 0-based candidates
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (t[mid] <= arg) lo = mid; else hi = mid - 1;
  }
  return lo;
}
template <int K>
__device__ __forceinline__ void fe_bspl(const double* __restrict__ t, double x, int L, double* h) {
  // fpbspl: the K+1 non-zero B-splines of degree K at t[L] <= x < t[L+1] (de Boor - Cox)
  double hh[K + 1];
  h[0] = 1.0;
#pragma unroll
  for (int j = 1; j <= K; ++j) {
#pragma unroll
    for (int i = 0; i < j; ++i) hh[i] = h[i];
    h[0] = 0.0;
#pragma unroll
    for (int i = 1; i <= j; ++i) {
      const double tli = t[L + i], tlj = t[L + i - j];
      const double f = hh[i - 1] / (tli - tlj);
      h[i - 1] = h[i - 1] + f * (tli - x);
      h[i] = f * (x - tlj);
    }
  }
}
// value (which = 0), d/du (1: first variable, y) or d/dv (2: second variable, x), in the map's units (nm, nm/mm)
__device__ __noinline__ double fe_eval_at(const double* __restrict__ tu, const double* __restrict__ tv,
                                          const double* __restrict__ c, int fe_nu, int fe_nv, double u, double v,
                                          int which) {
  const int nkv = fe_nv - 4;  // coefficients per row
  double hu[4], hv[4];
  double sp = 0.0;
  if (which == 0) {
    const int Lu = fe_interval(tu, fe_nu, 3, u), Lv = fe_interval(tv, fe_nv, 3, v);
    fe_bspl<3>(tu, u, Lu, hu);
    fe_bspl<3>(tv, v, Lv, hv);
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) sp = sp + c[(size_t)(Lu - 3 + a) * nkv + (Lv - 3 + b)] * hu[a] * hv[b];
  } else if (which == 1) {
    // parder, nux = 1: d[i][j] = (c[i+1][j] - c[i][j]) * 3 / (tu[i+4] - tu[i+1]) on the knots tu[1 .. nu-2], degree 2
    const double* __restrict__ tu1 = tu + 1;
    const int Lu = fe_interval(tu1, fe_nu - 2, 2, u), Lv = fe_interval(tv, fe_nv, 3, v);
    fe_bspl<2>(tu1, u, Lu, hu);
    fe_bspl<3>(tv, v, Lv, hv);
    for (int a = 0; a < 3; ++a) {
      const int i = Lu - 2 + a;
      const double fac = tu[i + 4] - tu[i + 1];
      for (int b = 0; b < 4; ++b) {
        const int j = Lv - 3 + b;
        const double c0 = c[(size_t)i * nkv + j], c1 = c[(size_t)(i + 1) * nkv + j];
        const double d = fac > 0.0 ? (c1 - c0) * 3.0 / fac : c0;
        sp = sp + d * hu[a] * hv[b];
      }
    }
  } else {
    const double* __restrict__ tv1 = tv + 1;
    const int Lu = fe_interval(tu, fe_nu, 3, u), Lv = fe_interval(tv1, fe_nv - 2, 2, v);
    fe_bspl<3>(tu, u, Lu, hu);
    fe_bspl<2>(tv1, v, Lv, hv);
    for (int a = 0; a < 4; ++a) {
      const int i = Lu - 3 + a;
      for (int b = 0; b < 3; ++b) {
        const int j = Lv - 2 + b;
        const double fac = tv[j + 4] - tv[j + 1];
        const double c0 = c[(size_t)i * nkv + j], c1 = c[(size_t)i * nkv + j + 1];
        const double d = fac > 0.0 ? (c1 - c0) * 3.0 / fac : c0;
        sp = sp + d * hu[a] * hv[b];
      }
    }
  }
  return sp;
}
__device__ __forceinline__ double fe_eval(const xrt_oe_params& p, double x, double y, int which) {
  return fe_eval_at(p.fe_tu, p.fe_tv, p.fe_c, p.fe_nu, p.fe_nv, y + p.fe_ushift, x + p.fe_vshift, which);
}
#endif

// the surface's loop-invariant divisors with their reciprocals refined once per thread (common.cuh, Divisor)
// quotients of a ray coordinate by a surface radius (toroid): operands are always in the range where the unchecked
// sequence is the IEEE quotient (common.cuh, div_by_nr); -DXRTB_SURF_DIV_CHECKED restores the range test
#ifdef XRTB_SURF_DIV_CHECKED
#define XRTB_SURF_DIV(a, d) div_by(a, d)
#define XRTB_SURF_SQRT(a) sqrt(a)
#define XRTB_SECANT_DIV(a, b) ((a) / (b))
#else
#define XRTB_SURF_DIV(a, d) div_by_nr(a, d)
#define XRTB_SURF_SQRT(a) sqrt_nr(a)  // zero or positive normal arguments (squares of coordinates, 1 - q^2)
// the secant step (t2 - t1) * dz1 / (dz2 - dz1): inside the loop dz1 and dz2 are non-zero with opposite signs (the pair
// brackets the root: solve_secant keeps that invariant), so |dz2 - dz1| >= max(|dz1|, |dz2|) > zEps, and the numerator is
// zero or >= ulp(t) * zEps -- the operands of the unchecked sequence
#define XRTB_SECANT_DIV(a, b) div_nr(a, b)
#endif
// Fresnel / Bragg amplitudes: quotients and roots of finite, non-zero optical quantities (refractive indices,
// susceptibilities, direction cosines) -- the unchecked primitives give the same correctly rounded values
#ifdef XRTB_SURF_DIV_CHECKED
#define XRTB_AMP_DIV(a, b) ((a) / (b))
#define XRTB_AMP_SQRT(a) sqrt(a)
#define XRTB_AMP_CDIV(a, b) cdiv(a, b)
#define XRTB_AMP_CSQRT(a) csqrt_(a)
#else
#define XRTB_AMP_DIV(a, b) div_nr(a, b)
#define XRTB_AMP_SQRT(a) sqrt_nr(a)
#define XRTB_AMP_CDIV(a, b) cdiv_nr(a, b)
#define XRTB_AMP_CSQRT(a) csqrt_nr(a)
#endif
struct SurfDiv {
  Divisor d0, d1;  // toroid: r, R; ellipse: A*A
};

template <int SURF>
__device__ __forceinline__ SurfDiv make_surf_div(const xrt_oe_params& p) {
  SurfDiv sd;
  sd.d0 = sd.d1 = Divisor{1.0, 1.0};
  if (SURF == 1) {
    sd.d0 = make_divisor(p.sp[1]);
    sd.d1 = make_divisor(p.sp[0]);
  }
  if (SURF == 2) sd.d0 = make_divisor(p.sp[0] * p.sp[0]);
  if (SURF == 3) {
    sd.d0 = make_divisor(4.0 * p.sp[0]);  // 4 * focus
    sd.d1 = make_divisor(2.0 * p.sp[0]);  // 2 * focus
  }
  return sd;
}

template <int SURF>
__device__ __forceinline__ double local_surface(const xrt_oe_params& p, const SurfDiv& sd, double x, double y) {
  if (SURF == 0) return 0.0;
#ifdef XRTB_USER_SURFACE
  if (SURF == 4 || SURF == 5) return xrtb_user::local_z(p.sp, x, y);
#endif
  if (SURF == 3) {
    // ParaboloidFlatLens.local_z1 (oes/refractive.py:394-399); the parabolic cylinder passes x = 0 (:613-614).
    // sp: focus, zmax, has_zmax, isCylinder
    const double xx = (p.sp[3] != 0.0) ? 0.0 : x * x;
    double z = div_by(xx + y * y, sd.d0);
    if (p.sp[2] != 0.0 && z > p.sp[1]) z = p.sp[1];
    return z;
  }
  if (SURF == 1) {
    // ToroidMirror.local_z (oes/__init__.py:398-401)
    const double r = p.sp[1];
    const double q = XRTB_SURF_DIV(x, sd.d0);  // x / r
    double rx = 1.0 - q * q;
    if (rx < 0.0) rx = 0.0;
    return XRTB_SURF_DIV(y * y / 2.0, sd.d1) + r * (1.0 - XRTB_SURF_SQRT(rx));  // y**2 / 2.0 / R
  }
  // EllipticalMirrorParam.local_r (parametric.py:228-234): x = s, y = phi
  const double B = p.sp[1];
  double rr = B * XRTB_SURF_SQRT(fabs(1.0 - XRTB_SURF_DIV(x * x, sd.d0)));  // s**2 / A**2
  if (p.sp[6] != 0.0) rr /= fabs(cos(y));
  return (fabs(y) > kPi / 2) ? rr : 1e20;
}

template <int SURF>
__device__ __forceinline__ void xyz_to_param(const xrt_oe_params& p, double& x, double& y, double& z) {
#ifdef XRTB_USER_PARAMETRIC
  if (SURF == 5) {
    xrtb_user::xyz_to_param(p.sp, x, y, z);
    return;
  }
#endif
  if (SURF != 2) return;
  // parametric.py:213-216: rotate_x(y - y0, z - z0, cosGamma, sinGamma); s, phi, r
  const double yy = y - p.sp[2], zz = z - p.sp[3];
  const double cg = p.sp[4], sg = p.sp[5];
  const double yN = cg * yy - sg * zz, zN = sg * yy + cg * zz;
  const double xx = x;
  x = yN;
  y = atan2(xx, zN);
  z = XRTB_SURF_SQRT(xx * xx + zN * zN);
}

template <int SURF>
__device__ __forceinline__ void param_to_xyz(const xrt_oe_params& p, double& x, double& y, double& z) {
#ifdef XRTB_USER_PARAMETRIC
  if (SURF == 5) {
    xrtb_user::param_to_xyz(p.sp, x, y, z);
    return;
  }
#endif
  if (SURF != 2) return;
  // parametric.py:218-224
  const double s = x, phi = y, r = z;
  const double xx = r * sin(phi), yy = s, zz = r * cos(phi);
  const double cg = p.sp[4], sg = -p.sp[5];
  const double yN = cg * yy - sg * zz, zN = sg * yy + cg * zz;
  x = xx;
  y = yN + p.sp[2];
  z = zN + p.sp[3];
}

template <int SURF>
__device__ __forceinline__ void local_normal(const xrt_oe_params& p, const SurfDiv& sd, double x, double y,
                                             double* n);

// find_dz (base.py:801-846); invertNormal = 1 from vacuum, -1 from inside a plate / lens.  DERIV = 0: the z (or r)
// difference between the ray and the surface; DERIV = 1 (multiple_reflect, base.py:818-820, :842-845): the ray
// direction along the surface normal at the point, whose root is where the ray is farthest from the surface
template <int SURF, bool REFR = true, int DERIV = 0>
__device__ __forceinline__ double find_dz(const xrt_oe_params& p, const SurfDiv& sd, double t, const Ray& r,
                                          double& x, double& y, double& z) {
  x = r.x + r.a * t;
  y = r.y + r.b * t;
  z = r.z + r.c * t;
  xyz_to_param<SURF>(p, x, y, z);
  if (DERIV == 1) {
    double n[6];
    local_normal<SURF>(p, sd, x, y, n);
    const double dn = r.a * n[3] + r.b * n[4] + r.c * n[5];
    return (REFR && p.from_vacuum == 0) ? -dn : dn;
  }
  double surf = local_surface<SURF>(p, sd, x, y);
#ifdef XRTB_FIGURE_ERROR
  if (p.fe_c) {
    if (SURF == 2 || SURF == 5) {
      // parametric element: local_r_distorted (base.py:790-797) -- the height map acts along z at the Cartesian image
      // of (s, phi, r) and the new point is taken back to r: surf += r1 - r
      double fx = x, fy = y, fz = surf;
      param_to_xyz<SURF>(p, fx, fy, fz);
      fz += fe_eval(p, fx, fy, 0) * 1e-6;
      xyz_to_param<SURF>(p, fx, fy, fz);
      surf = surf + (fz - surf);
    } else {
      surf = surf + fe_eval(p, x, y, 0) * 1e-6;  // surf += z_distorted, nm -> mm (base.py:825-829)
    }
  }
#endif
  if (isnan(surf)) surf = 0.0;
  double d = z - surf;
  if (SURF == 2 || SURF == 5) d = -d;  // parametric: local_f = local_r, diffSign = -1 (base.py:809-813, :840)
  // * invertNormal = -1 from inside the material (reflect.py:638-641); REFR = false (kernels of non-refractive
  // materials) compiles the test away
  return (REFR && p.from_vacuum == 0) ? -d : d;
}

// normal(s): n[0..2] = n_B (or the only normal), n[3..5] = surface normal
template <int SURF>
__device__ __forceinline__ void local_normal(const xrt_oe_params& p, const SurfDiv& sd, double x, double y,
                                             double* n) {
#ifdef XRTB_USER_SURFACE
  if (SURF == 4 || SURF == 5) {
    xrtb_user::local_n(p.sp, x, y, n);
    if (!p.asymmetric) {  // one normal returned: oeNormal[-3:] is the same vector (reflect.py:783-787)
      n[3] = n[0];
      n[4] = n[1];
      n[5] = n[2];
    }
    return;
  }
#endif
  if (SURF == 0) {
    n[3] = 0.0;
    n[4] = 0.0;
    n[5] = 1.0;
    if (p.asymmetric) {
      n[0] = p.nB[0];
      n[1] = p.nB[1];
      n[2] = p.nB[2];
    } else {
      n[0] = 0.0;
      n[1] = 0.0;
      n[2] = 1.0;
    }
    return;
  }
  if (SURF == 3) {
    // ParaboloidFlatLens.local_n1 (oes/refractive.py:405-419); cylinder: x = 0 (:616-617)
    double a = (p.sp[3] != 0.0) ? 0.0 : div_by(-x, sd.d1);  // -x / (2*focus)
    double b = div_by(-y, sd.d1);
    if (p.sp[2] != 0.0) {
      const double xx = (p.sp[3] != 0.0) ? 0.0 : x * x;
      const double z = div_by(xx + y * y, sd.d0);
      if (z > p.sp[1]) {
        if (p.sp[3] == 0.0) a = 0.0;  // the cylinder's a is a Python scalar 0 that the mask never touches
        b = 0.0;
      }
    }
    const double norm = sqrt(a * a + b * b + 1.0);
    n[0] = n[3] = a / norm;
    n[1] = n[4] = b / norm;
    n[2] = n[5] = 1.0 / norm;
    return;
  }
  if (SURF == 1) {
    // ToroidMirror.local_n (oes/__init__.py:403-411)
    const double q = XRTB_SURF_DIV(x, sd.d0);  // x / r
    const double rx = 1.0 - q * q;
    const double ax = (rx < 0.0) ? 0.0 : pow_m05(rx);
    const double a = XRTB_SURF_DIV(-x, sd.d0) * ax;
    const double b = XRTB_SURF_DIV(-y, sd.d1);  // -y / R
    const double norm = XRTB_SURF_SQRT(a * a + b * b + 1.0);
    const Divisor dn = make_divisor(norm);  // three quotients, one refined reciprocal (same bits as a / norm etc.)
    n[0] = n[3] = XRTB_SURF_DIV(a, dn);
    n[1] = n[4] = XRTB_SURF_DIV(b, dn);
    n[2] = n[5] = XRTB_SURF_DIV(1.0, dn);
    return;
  }
  // EllipticalMirrorParam.local_n (parametric.py:236-249): x = s, y = phi
  const double A = p.sp[0], B = p.sp[1];
  double A2s2 = A * A - x * x;
  if (A2s2 <= 0.0) A2s2 = 1e22;
  // (A2s2 > 0 and norm >= 1: the unchecked sequences give the IEEE quotients and roots, see XRTB_SURF_DIV)
  const double nr = XRTB_SURF_DIV(-B / A * x, make_divisor(XRTB_SURF_SQRT(A2s2)));
  const double norm = XRTB_SURF_SQRT(nr * nr + 1.0);
  const Divisor dn = make_divisor(norm);
  const double b = XRTB_SURF_DIV(nr, dn);
  double a, c;
  if (p.sp[6] != 0.0) {
    a = 0.0;
    c = XRTB_SURF_DIV(1.0, dn);
  } else {
    a = XRTB_SURF_DIV(-sin(y), dn);
    c = XRTB_SURF_DIV(-cos(y), dn);
  }
  const double cg = p.sp[4], sg = -(-p.sp[5]);  // rotate_x(b, c, cosGamma, -sinGamma): -sinangle*z = +sinGamma*z
  n[0] = n[3] = a;
  n[1] = n[4] = cg * b - (-p.sp[5]) * c;
  n[2] = n[5] = (-p.sp[5]) * b + cg * c;
  (void)sg;
}

// _set_t (base.py:1231-1245) followed by the tMin clamp of _bracketing (:1275)
__device__ __forceinline__ void bracket(const xrt_oe_params& p, int axis, int neg, const Ray& r, double& t1,
                                        double& t2) {
  const double xyz = axis == 0 ? r.x : (axis == 1 ? r.y : r.z);
  const double abc = axis == 0 ? r.a : (axis == 1 ? r.b : r.c);
  const double lo = p.brk_lim[axis][0], hi = p.brk_lim[axis][1];
  const Divisor dabc = make_divisor(abc);  // both quotients share the refined reciprocal (same bits as two divisions)
  if (!neg) {
    t1 = div_by(lo - xyz, dabc) - kDt;
    t2 = div_by(hi - xyz, dabc) + kDt;
  } else {
    t1 = div_by(hi - xyz, dabc) - kDt;
    t2 = div_by(lo - xyz, dabc) + kDt;
  }
  if (t1 < -1e6 * kZEps) t1 = -1e6 * kZEps;
}

__device__ __forceinline__ Ray load_ray(const xrt_beam& b, size_t i) {
  Ray r;
  r.x = b.x[i];
  r.y = b.y[i];
  r.z = b.z[i];
  r.a = b.a[i];
  r.b = b.b[i];
  r.c = b.c[i];
  return r;
}

// ---------------------------------------------------------------- pre-pass: beam-wide scalars
__global__ void reduction_init_kernel(unsigned long long* red) {
  const int i = threadIdx.x;
  if (i >= XRT_REDUCTION_SLOTS) return;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  unsigned long long v = enc(-inf);  // max slots
  if (i == RS_FIRST) v = ~0ull;
  if (i >= RS_VAR && i < RS_VAR + 24 && ((i - RS_VAR) & 3) == 0) v = enc(inf);  // min t1
  if (i >= RS_FOOT && i < RS_FOOT + 3) v = enc(inf);
  if (i == RS_DERIV) v = enc(inf);
  if (i == RS_NGOOD || i == RS_NGOODN || (i > RS_NGOODN && i < RS_DERIV) || i >= RS_DERIV + 4) v = 0ull;
  red[i] = v;
}

// Phase A: dominant direction component and first good ray (28 B/ray: state, a, b, c)
__global__ void __launch_bounds__(256) reflect_prepass_dir_kernel(size_t n, xrt_oe_params p, xrt_beam in,
                                                                  unsigned long long* red,
                                                                  unsigned long long index_offset) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double mx[3] = {-inf, -inf, -inf};
  unsigned long long first = ~0ull, ngood = 0;
  // U rays per thread and iteration with every load issued before the first use (dead rays included: their
  // directions are fetched and dropped), so that 28 B x U per thread are in flight instead of one dependent
  // state -> (a, b, c) chain: the pass is pure HBM streaming
  constexpr int U = 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * U) {
    int stv[U];
    double av[U], bv[U], cv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      const bool in_range = i < n;
      stv[u] = in_range ? in.state[i] : 0;
      av[u] = in_range ? in.a[i] : 0.0;
      bv[u] = in_range ? in.b[i] : 0.0;
      cv[u] = in_range ? in.c[i] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      const int st = stv[u];
      if (i >= n || !(p.good_mode ? (st == 1 || st == 2) : (st > 0))) continue;
      ++ngood;
      Ray r;
      r.x = r.y = r.z = 0.0;  // directions transform independently of positions
      r.a = av[u];
      r.b = bv[u];
      r.c = cv[u];
      // positions are not needed here: only the rotations act on (a, b, c)
      if (p.from_global && p.sinAz != 0.0) rot_axis(2, p.cosAz, p.sinAz, r.a, r.b, r.c);
      for (int k = 0; k < p.n_rot_in; ++k)
        rot_axis(p.rot_in_axis[k], p.rot_in_cos[k], p.rot_in_sin[k], r.a, r.b, r.c);
      if (st == 1) {
        mx[0] = fmax(mx[0], fabs(r.a));
        mx[1] = fmax(mx[1], fabs(r.b));
        mx[2] = fmax(mx[2], fabs(r.c));
      }
      const unsigned long long key = ((index_offset + i) << 3) | (r.a > 0.0 ? 1ull : 0ull) |
                                     (r.b > 0.0 ? 2ull : 0ull) | (r.c > 0.0 ? 4ull : 0ull);
      first = key < first ? key : first;
    }
  }
  __shared__ unsigned long long sh[8][5];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long e[5] = {enc(mx[0]), enc(mx[1]), enc(mx[2]), first, ngood};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    unsigned long long x = e[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, o);
      x = (k == 4) ? x + y : (k == 3 ? (y < x ? y : x) : (y > x ? y : x));
    }
    if (lane == 0) sh[warp][k] = x;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    const int k = threadIdx.x;
    unsigned long long x = sh[0][k];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      const unsigned long long y = sh[w][k];
      x = (k == 4) ? x + y : (k == 3 ? (y < x ? y : x) : (y > x ? y : x));
    }
    if (k == 4) {
      if (x) atomicAdd(&red[RS_NGOOD], x);
    } else if (k == 3) {
      atomicMin(&red[RS_FIRST], x);
    } else {
      atomicMax(&red[k], x);
    }
  }
}

__device__ __forceinline__ void decide_variant(const unsigned long long* red, int& axis, int& neg) {
  double maxa = dec(red[RS_MAXA]), maxb = dec(red[RS_MAXB]), maxc = dec(red[RS_MAXC]);
  if (maxa < 0.0 && maxb < 0.0 && maxc < 0.0) {  // no ray with state == 1: (0, 1, 0) (base.py:1261-1262)
    maxa = 0.0;
    maxb = 1.0;
    maxc = 0.0;
  }
  const double maxMax = fmax(maxa, fmax(maxb, maxc));
  axis = (maxMax == maxa) ? 0 : ((maxMax == maxb) ? 1 : 2);
  const unsigned long long key = red[RS_FIRST];
  neg = (int)((key >> axis) & 1ull) ? 0 : 1;
}

// Phase B: global t bracket and the secant-vs-Brent switch for the (axis, sign) variant phase A selected
// (52 B/ray: state, x, y, z, a, b, c)
template <int SURF>
__global__ void __launch_bounds__(256) reflect_prepass_kernel(size_t n, xrt_oe_params p, xrt_beam in,
                                                              unsigned long long* red) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  int axis, neg;
  decide_variant(red, axis, neg);
  double v[4] = {inf, -inf, -inf, -inf};
  const SurfDiv sd = make_surf_div<SURF>(p);
  // (one ray per iteration: this pass is bound by the two find_dz evaluations, not by load latency -- hoisting the
  // loads of two rays as in phase A measured 1.57 ms instead of 1.47 ms per 1e8 rays)
  const size_t stride = (size_t)gridDim.x * blockDim.x;
#ifndef XRTB_PREPASS_NO_PREFETCH
  // software pipeline: the next ray of this thread is requested before the current one is evaluated (80 registers,
  // three blocks per SM: 1.11 instead of 1.32 ms per 1e8 rays)
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int st_n = 0;
  Ray r_n{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (i < n) {
    st_n = in.state[i];
    r_n = load_ray(in, i);
  }
  for (; i < n; i += stride) {
    const int st = st_n;
    Ray r = r_n;
    if (i + stride < n) {
      st_n = in.state[i + stride];
      r_n = load_ray(in, i + stride);
    }
    if (!(p.good_mode ? (st == 1 || st == 2) : (st > 0))) continue;
#else
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int st = in.state[i];
#ifndef XRTB_PREPASS_DEPENDENT_LOADS
    Ray r = load_ray(in, i);  // requested together with the state (dead rays: fetched and dropped), one latency not two
    if (!(p.good_mode ? (st == 1 || st == 2) : (st > 0))) continue;
#else
    if (!(p.good_mode ? (st == 1 || st == 2) : (st > 0))) continue;
    Ray r = load_ray(in, i);
#endif
#endif
    to_local(p, r);
    double t1, t2, x, y, z;
    bracket(p, axis, neg, r, t1, t2);
    if (p.multi_tmin) t1 = p.multi_tmin[i] + kDs;  // base.py:1288: tMin = tMax(derivative root) + ds
    const double dz1 = find_dz<SURF>(p, sd, t1, r, x, y, z);
    double dz2 = find_dz<SURF>(p, sd, t2, r, x, y, z);
    if (dz1 <= 0.0 || dz2 >= 0.0) dz2 = 0.0;  // base.py:865 dz2[ind1 | ind2] = 0
    // numpy's min/max propagate NaN; keep that by letting NaN win
    v[0] = (t1 < v[0] || isnan(t1)) ? t1 : v[0];
    v[1] = (t2 > v[1] || isnan(t2)) ? t2 : v[1];
    v[2] = (fabs(dz1) > v[2] || isnan(dz1)) ? fabs(dz1) : v[2];
    v[3] = (fabs(dz2) > v[3] || isnan(dz2)) ? fabs(dz2) : v[3];
  }
  __shared__ unsigned long long sh[8][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned long long x = enc(v[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, o);
      x = (k == 0) ? (y < x ? y : x) : (y > x ? y : x);
    }
    if (lane == 0) sh[warp][k] = x;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const int k = threadIdx.x;
    unsigned long long x = sh[0][k];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      const unsigned long long y = sh[w][k];
      x = (k == 0) ? (y < x ? y : x) : (y > x ? y : x);
    }
    const int slot = RS_VAR + (axis * 2 + neg) * 4 + k;
    if (k == 0)
      atomicMin(&red[slot], x);
    else
      atomicMax(&red[slot], x);
  }
}

// ---------------------------------------------------------------- numpy-faithful complex arithmetic
// numpy >= 1.25 evaluates complex multiply/square with its SIMD kernel on every x86 CPU that has FMA3
// (numpy/_core/src/umath/loops_arithm_fp.dispatch.c.src, simd_cmul: ab = a_im * b_rev rounded, then
// fmaddsub(a_re, b, ab)), i.e. re = fma(ar, br, -RN(ai*bi)), im = fma(ar, bi, RN(ai*br)).  The grazing-incidence
// Fresnel chain amplifies 1-ulp differences by ~1e4, so the same fusion pattern is reproduced here.
__device__ __forceinline__ cplx cmul_np(cplx a, cplx b) {
  const double ii = a.im * b.im, ir = a.im * b.re;
  return cplx{fma(a.re, b.re, -ii), fma(a.re, b.im, ir)};
}
__device__ __forceinline__ cplx csquare_np(cplx a) { return cmul_np(a, a); }
// complex / real: numpy casts the real to (b, 0) and runs Smith's algorithm, where rat = 0/b = 0 and
// scl = 1/(b + 0*0): the quotient is (a.re*scl, a.im*scl) -- one division instead of two, same bits
__device__ __forceinline__ cplx cdiv_re(cplx a, double b) {
  const double scl = 1.0 / b;
  return cplx{a.re * scl, a.im * scl};
}
__device__ __forceinline__ cplx cdiv_re_amp(cplx a, double b) {  // b finite, non-zero (XRTB_AMP_DIV)
  const double scl = XRTB_AMP_DIV(1.0, b);
  return cplx{a.re * scl, a.im * scl};
}
__device__ __forceinline__ cplx cexp_i(double phi) {  // exp(0 + i phi)
  double s, c;
  sincos_big(phi, s, c);
  return cplx{c, s};
}

// np.interp (numpy/core/src/multiarray/compiled_base.c): binary search, slope form -- see f1f2() below
struct Tables {
  const double *E, *f1, *f2;
  int* hint;  // interval guesses per element table + a dirty flag (33 ints, block-shared; may be NULL): f1f2_locate
};

// slope form of np.interp at a located interval [lo, lo + 1]; `dx` = make_divisor(xp[lo + 1] - xp[lo]) is shared by
// the f1 and the f2 interpolation (same quotient bits as the plain division, common.cuh)
__device__ __forceinline__ double interp_at(const double* __restrict__ xp, const double* __restrict__ fp, int len,
                                            int lo, double x, const Divisor& dx) {
  if (lo < 0) return fp[0];
  if (lo >= len - 1) return fp[len - 1];
  if (xp[lo] == x) return fp[lo];
  const double slope = XRTB_SURF_DIV(fp[lo + 1] - fp[lo], dx);  // table differences over a table step
  double res = slope * (x - xp[lo]) + fp[lo];
  if (isnan(res)) {
    res = slope * (x - xp[lo + 1]) + fp[lo + 1];
    if (isnan(res) && fp[lo] == fp[lo + 1]) res = fp[lo];
  }
  return res;
}

// Element.get_f1f2 (element.py:252-263): f1 and f2 share the energy grid, so one binary search serves both
// np.interp calls; the out-of-range ValueError becomes NaN + an error flag.
// The interval of np.interp's binary search: xp[lo] <= E < xp[lo + 1] (lo = len - 1 at the last point), -2: E is
// outside the table.  The search is ten dependent loads for a 768-point table; beams are narrow-band far more often
// than not, so the last interval found is kept as a guess per element table (`hint`: 32 ints in the block's shared
// memory, seeded from and -- after a miss -- written back to reduction slots 48..63 once per block) and the search runs
// only when the energy is not in the guessed interval or one of its two neighbours (four independent loads).  The
// interval is unique, so the result does not depend on the hint.  hint[32] is the block's "a guess changed" flag.
__device__ __forceinline__ int f1f2_locate(const Tables& t, int off, int len, double E) {
  const double* __restrict__ xp = t.E + off;
  if (E < xp[0] || E > xp[len - 1]) return -2;
  if (E >= xp[len - 1]) return len - 1;
  volatile int* hp = t.hint ? t.hint + ((unsigned)off % 31u) : nullptr;  // one word per element table
  if (hp) {
    const int h = *hp;
    if (h >= 1 && h < len - 2) {
      const double e0 = xp[h - 1], e1 = xp[h], e2 = xp[h + 1], e3 = xp[h + 2];
      if (e1 <= E && E < e2) return h;
      if (e0 <= E && E < e1) return h - 1;
      if (e2 <= E && E < e3) return h + 1;
    }
  }
  int lo = 0, hi = len - 1;  // xp[lo] <= E < xp[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (E >= xp[mid])
      lo = mid;
    else
      hi = mid;
  }
  if (hp) {
    // move the guess halfway towards this interval (all the way when it was unset or far off): a band that covers
    // three intervals then settles on the middle one instead of flipping between the outer two
    const int h = *hp;
    *hp = (h >= 1 && h < len - 2 && abs(h - lo) <= 4) ? (h + lo + (lo > h ? 1 : 0)) / 2 : lo;
    t.hint[32] = 1;
  }
  return lo;
}
__device__ __forceinline__ cplx f1f2_at(const Tables& t, int off, int len, int lo, double E, int* oor) {
  if (lo == -2) {
    *oor = 1;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    return cplx{nan, nan};
  }
  const double* __restrict__ xp = t.E + off;
  Divisor dx{1.0, 1.0};
  if (lo >= 0 && lo < len - 1) dx = make_divisor(xp[lo + 1] - xp[lo]);
  return cplx{interp_at(xp, t.f1 + off, len, lo, E, dx), interp_at(xp, t.f2 + off, len, lo, E, dx)};
}
__device__ __forceinline__ cplx f1f2(const Tables& t, int off, int len, double E, int* oor) {
  return f1f2_at(t, off, len, f1f2_locate(t, off, len, E), E, oor);
}

// Material.get_refractive_index (material.py:372-378)
// lo0 >= -2: the located interval of the FIRST element's table (f1f2_locate), found earlier by the caller
__device__ __forceinline__ cplx refractive_index(const xrt_material_params& m, const Tables& t, double E, int* oor,
                                                 int lo0 = -3) {
  cplx xf = cplx{0.0, 0.0};
  for (int i = 0; i < m.nelem; ++i) {
    const cplx f = (i == 0 && lo0 != -3) ? f1f2_at(t, m.tab_off[0], m.tab_len[0], lo0, E, oor)
                                         : f1f2(t, m.tab_off[i], m.tab_len[i], E, oor);
    const cplx zf = cplx{m.Z[i] + f.re, f.im};            // elem.Z + (f1 + i f2)
    const cplx term = cmul_np(zf, cplx{m.quantity[i], 0.0});  // * xi
    xf = cadd(xf, term);
  }
  const double c0 = 1e-24 * 6.02214199e23 * 2.817940285e-5 / 6.283185307179586476925286766559;
  const double CH = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8;
  const double w = XRTB_AMP_DIV(CH, E);
  const double fac = c0 * (w * w) * m.rho;
  const cplx num = cmul_np(cplx{fac, 0.0}, xf);
  const cplx q = cdiv_re_amp(num, m.mass);
  return cplx{1.0 - q.re, 0.0 - q.im};
}

// Material.get_amplitude, fromVacuum=True, kinds 'mirror' / 'thin mirror' (material.py:455-493)
__device__ __forceinline__ void mirror_amplitude(const xrt_material_params& m, const Tables& t, double E,
                                                 double bdn, cplx& rs, cplx& rp, int* oor, int lo0 = -3) {
  const cplx n2 = refractive_index(m, t, E, oor, lo0);
  const double cosAlpha = fabs(bdn);
  double sinAlpha2 = 1.0 - bdn * bdn;
  if (sinAlpha2 < 0.0) sinAlpha2 = 0.0;
  const double n1cosAlpha = 1.0 * cosAlpha;
  const cplx n1n2 = XRTB_AMP_CDIV((cplx{1.0, 0.0}), n2);
  const cplx sq = cmul_np(csquare_np(n1n2), cplx{sinAlpha2, 0.0});
  const cplx cosBeta = XRTB_AMP_CSQRT((cplx{1.0 - sq.re, 0.0 - sq.im}));
  const cplx n2cosBeta = cmul_np(n2, cosBeta);
  rs = XRTB_AMP_CDIV((cplx{n1cosAlpha - n2cosBeta.re, 0.0 - n2cosBeta.im}),
                     (cplx{n1cosAlpha + n2cosBeta.re, 0.0 + n2cosBeta.im}));
  const cplx n2ca = cmul_np(n2, cplx{cosAlpha, 0.0});
  const cplx n1cb = cmul_np(cplx{1.0, 0.0}, cosBeta);
  rp = XRTB_AMP_CDIV(csub(n2ca, n1cb), cadd(n2ca, n1cb));
  if (m.kind == 2) {
    // thin mirror: p2 = exp(2j * E / CHBAR * n2cosBeta * t * 1e7) (material.py:479-482)
    const double CHBAR = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8 / 6.283185307179586476925286766559;
    cplx z = cmul_np(cplx{0.0, 2.0}, cplx{E, 0.0});
    z = cdiv_re(z, CHBAR);
    z = cmul_np(z, n2cosBeta);
    z = cmul_np(z, cplx{m.t, 0.0});
    z = cmul_np(z, cplx{1e7, 0.0});
    const double ex = exp(z.re);
    double s, c;
    sincos(z.im, &s, &c);
    const cplx p2 = cplx{ex * c, ex * s};
    const cplx one_m_p2 = cplx{1.0 - p2.re, 0.0 - p2.im};
    cplx den = cmul_np(csquare_np(rs), p2);
    rs = cmul_np(rs, cdiv(one_m_p2, cplx{1.0 - den.re, 0.0 - den.im}));
    den = cmul_np(csquare_np(rp), p2);
    rp = cmul_np(rp, cdiv(one_m_p2, cplx{1.0 - den.re, 0.0 - den.im}));
  }
}

// Material.get_amplitude, kinds 'plate' / 'lens' (transmittivity, material.py:455-493), from vacuum into the
// material or back.  mu = abs(n.imag)*E/CHBAR*2e8 [1/cm], nk = n.real*E/CHBAR*1e8 [1/cm].
__device__ __forceinline__ void plate_amplitude(const xrt_material_params& m, const Tables& t, double E, double bdn,
                                                bool fromVacuum, cplx& rs, cplx& rp, double& mu, double& nk,
                                                double& nreal, int* oor) {
  const cplx n = refractive_index(m, t, E, oor);
  const double CHBAR = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8 / 6.283185307179586476925286766559;
  mu = fabs(n.im) * E / CHBAR * 2e8;
  nk = n.re * E / CHBAR * 1e8;
  nreal = n.re;
  const double cosAlpha = fabs(bdn);
  double sinAlpha2 = 1.0 - bdn * bdn;
  if (sinAlpha2 < 0.0) sinAlpha2 = 0.0;
  if (fromVacuum) {
    // n1 = 1., n2 = n
    const double n1cosAlpha = 1.0 * cosAlpha;
    const cplx n1n2 = cdiv(cplx{1.0, 0.0}, n);
    const cplx sq = cmul_np(csquare_np(n1n2), cplx{sinAlpha2, 0.0});
    const cplx cosBeta = csqrt_(cplx{1.0 - sq.re, 0.0 - sq.im});
    const cplx n2cosBeta = cmul_np(n, cosBeta);
    const double tf = sqrt(n2cosBeta.re / cosAlpha) / 1.0;  // (n2cosBeta * 1.0.conjugate()).real / cosAlpha
    const cplx two = cplx{2.0 * n1cosAlpha, 0.0};
    rs = cmul_np(cdiv(two, cplx{n1cosAlpha + n2cosBeta.re, 0.0 + n2cosBeta.im}), cplx{tf, 0.0});
    const cplx n2ca = cmul_np(n, cplx{cosAlpha, 0.0});
    const cplx n1cb = cmul_np(cplx{1.0, 0.0}, cosBeta);
    rp = cmul_np(cdiv(two, cadd(n2ca, n1cb)), cplx{tf, 0.0});
  } else {
    // n1 = n, n2 = 1.
    const cplx n1cosAlpha = cmul_np(n, cplx{cosAlpha, 0.0});
    const cplx sq = cmul_np(csquare_np(n), cplx{sinAlpha2, 0.0});  // (n / 1.)**2 * sinAlpha2
    const cplx cosBeta = csqrt_(cplx{1.0 - sq.re, 0.0 - sq.im});
    const cplx n2cosBeta = cosBeta;  // 1. * cosBeta
    const double tf = sqrt(cmul_np(n2cosBeta, cconj(n)).re / cosAlpha) / hypot(n.re, n.im);
    const cplx two = cplx{2.0 * n1cosAlpha.re, 2.0 * n1cosAlpha.im};
    rs = cmul_np(cdiv(two, cadd(n1cosAlpha, n2cosBeta)), cplx{tf, 0.0});
    const cplx n1cb = cmul_np(n, cosBeta);
    rp = cmul_np(cdiv(two, cplx{1.0 * cosAlpha + n1cb.re, 0.0 + n1cb.im}), cplx{tf, 0.0});
  }
}

// np.exp / np.sin / np.cos / np.tan on complex128 = glibc cexp / csin / ccos / ctan (finite arguments of
// moderate size; the overflow-scaling branches of glibc for |im| > 709 (355 for ctan) are not reproduced)
__device__ __forceinline__ cplx cexp_np(cplx z) {
  double sn = z.im, cs = 1.0;
  if (fabs(z.im) > 2.2250738585072014e-308) sincos(z.im, &sn, &cs);
  const double e = exp(z.re);
  return cplx{e * cs, e * sn};
}
__device__ __forceinline__ cplx csin_np(cplx z) {
  double sn = z.re, cs = 1.0;
  if (fabs(z.re) > 2.2250738585072014e-308) sincos(z.re, &sn, &cs);
  return cplx{cosh(z.im) * sn, sinh(z.im) * cs};
}
__device__ __forceinline__ cplx ccos_np(cplx z) {  // ccosh(i z), i z = (-im, re)
  const double x = -z.im, y = z.re;
  double sn = y, cs = 1.0;
  if (fabs(y) > 2.2250738585072014e-308) sincos(y, &sn, &cs);
  return cplx{cosh(x) * cs, sinh(x) * sn};
}
__device__ __forceinline__ cplx ctan_np(cplx z) {
  double sn = z.re, cs = 1.0;
  if (fabs(z.re) > 2.2250738585072014e-308) sincos(z.re, &sn, &cs);
  double sh = z.im, ch = 1.0;
  if (fabs(z.im) > 2.2250738585072014e-308) {
    sh = sinh(z.im);
    ch = cosh(z.im);
  }
  double den = cs * cs;
  if (fabs(sh) > fabs(cs) * 2.220446049250313e-16) den = cs * cs + sh * sh;
  return cplx{sn * cs / den, sh * ch / den};
}

// Crystal.get_amplitude, one-element fcc/diamond structure (crystal.py:566-616): thick Bragg, and for a crystal
// of thickness t (mm) the Bragg / Laue, reflected / transmitted amplitudes of Belyakov & Dmitrienko.
// geom bit0 = Laue, bit1 = transmitted
__device__ __forceinline__ cplx bragg_one_pol(double polFactor, cplx alpha, cplx chih, cplx chih_, double b,
                                              double tmm, int geom, cplx chi0, double k02, double kHs, double k0s) {
  // delta = sqrt(alpha**2 + polFactor**2 * chih * chih_ / b)
  cplx t = cmul_np(cplx{polFactor * polFactor, 0.0}, chih);
  t = cmul_np(t, chih_);
  t = cdiv_re_amp(t, b);
  const cplx delta = XRTB_AMP_CSQRT(cadd(csquare_np(alpha), t));
  const cplx num = cmul_np(chih, cplx{polFactor, 0.0});
  if (tmm > 0.0) {
    const double tA = tmm * 1e7;
    // l = t * delta * k02 / 2. / kHs
    cplx l = cmul_np(cmul_np(cplx{tA, 0.0}, delta), cplx{k02, 0.0});
    l = cdiv_re(cdiv_re(l, 2.0), kHs);
    cplx ra;
    // exp(1j * k02 * t * (chi0 - alpha*b) / 2 / k0s)
    cplx ph = cplx{0.0, 0.0};
    if (geom != 0) {
      const cplx w = csub(chi0, cmul_np(alpha, cplx{b, 0.0}));
      cplx e = cmul_np(cmul_np(cmul_np(cplx{0.0, 1.0}, cplx{k02, 0.0}), cplx{tA, 0.0}), w);
      e = cdiv_re(cdiv_re(e, 2.0), k0s);
      ph = cexp_np(e);
    }
    const cplx ia = cmul_np(cplx{0.0, 1.0}, alpha);
    if (!(geom & 1)) {  // Bragg
      if (geom & 2) {
        const cplx q = cdiv(cmul_np(ia, csin_np(l)), delta);
        ra = cmul_np(cdiv(cplx{1.0, 0.0}, csub(ccos_np(l), q)), ph);
      } else {
        const cplx q = cdiv(cmul_np(cplx{0.0, 1.0}, delta), ctan_np(l));
        ra = cdiv(num, cadd(alpha, q));
      }
    } else {  // Laue
      if (geom & 2) {
        const cplx q = cdiv(cmul_np(ia, csin_np(l)), delta);
        ra = cmul_np(cadd(ccos_np(l), q), ph);
      } else {
        ra = cmul_np(cdiv(cmul_np(num, csin_np(l)), delta), ph);
      }
    }
    if (!(geom & 2)) ra = cdiv_re(ra, sqrt(fabs(b)));
    return ra;
  }
  cplx ra = XRTB_AMP_CDIV(num, cadd(alpha, delta));  // alpha + delta == 0: NaN either way, caught below
  cplx ad = csub(alpha, delta);
  if (ad.re == 0.0 && ad.im == 0.0) ad = cplx{1e-100, 0.0};
  const cplx rb = XRTB_AMP_CDIV(num, ad);
  if (isnan(ra.re) || isnan(ra.im)) ra = rb;
  // abs(rb) < abs(ra) (crystal.py:592-596) decided on the squared magnitudes: the two roots differ by the absorption
  // (percents), the values are far from over- / underflow, and two hypot calls per polarisation are ~100 instructions
  if (cabs2(rb) < cabs2(ra)) ra = rb;
  return cdiv_re_amp(ra, XRTB_AMP_SQRT(fabs(b)));
}

__device__ __forceinline__ void crystal_amplitude(const xrt_material_params& m, const Tables& t, double E,
                                                  double bInDotN, double bOutDotN, double bInDotH, cplx& rs, cplx& rp,
                                                  int* oor) {
  const double CH = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8;
  const double PI2 = 6.283185307179586476925286766559;
  const double waveLength = XRTB_AMP_DIV(CH, E);
  const double k = XRTB_AMP_DIV(PI2, waveLength);
  const double k0s = -bInDotN * k;
  double kHs = -bOutDotN * k;
  const double HH = PI2 / m.d;
  const double k0H = fabs(bInDotH) * HH * k;
  const double k02 = k * k;
  const double H2 = HH * HH;
  const bool kHs0 = (kHs == 0.0);
  if (kHs0) kHs = 1.0;
  double b = XRTB_AMP_DIV(k0s, kHs);
  if (kHs0) b = -1.0;
  // structure factors (crystals_basic.py:22-31, :76-80)
  const cplx anom = f1f2(t, m.tab_off[0], m.tab_len[0], E, oor);
  cplx F0 = cmul_np(cmul_np(cplx{4.0, 0.0}, cplx{m.Z[0] + anom.re, anom.im}), cplx{m.factDW, 0.0});
  cplx Fhkl = cplx{0.0, 0.0}, Fhkl_ = cplx{0.0, 0.0};
  if (m.hkl_allowed) {
    const double q = 0.5 / m.d;
    // f0 = c + sum(a_i exp(-b_i q^2)) depends on the crystal only: evaluated by the host (element.py:203-207)
    (void)q;
    const double f0 = m.f0c[0];
    Fhkl = cmul_np(cmul_np(cplx{4.0, 0.0}, cplx{f0 + anom.re, anom.im}), cplx{m.factDW, 0.0});
    Fhkl_ = Fhkl;
  }
  // diamond: F0 * 2, Fhkl * diamondToFcc (crystals_basic.py:76-80); in general, for a cell of like atoms at r_i with
  // occupancies o_i: f0scale = sum(o_i) / 4 and d2f = sum(o_i exp(2 pi i hkl.r_i)) / 4 (kernel `get_amplitude`,
  // cl/materials.cl:203-221, through XRT_CUDA.run_parallel('get_amplitude'))
  if (m.f0scale != 1.0 || m.d2f_re != 1.0 || m.d2f_im != 0.0) {
    F0 = cmul_np(F0, cplx{m.f0scale, 0.0});
    Fhkl = cmul_np(Fhkl, cplx{m.d2f_re, m.d2f_im});
    Fhkl_ = cmul_np(Fhkl_, cplx{m.d2f_re, -m.d2f_im});
  }
  // get_F_chi (crystal.py:297-306)
  const double c2f = m.chiToF * (waveLength * waveLength);
  const cplx chi0 = cmul_np(cconj(F0), cplx{c2f, 0.0});
  const cplx chih = cmul_np(cconj(Fhkl), cplx{c2f, 0.0});
  const cplx chih_ = cmul_np(cconj(Fhkl_), cplx{c2f, 0.0});
  // thetaB (crystal.py:1105-1120)
  double sa = XRTB_AMP_DIV(1.0 * CH, 2.0 * m.d * E);
  if (sa > 1.0) sa = 1.0 - 1e-16;
  if (sa < -1.0) sa = -1.0 + 1e-16;
  // polarisation factor of the p branch: cos(2 thetaB) with thetaB = arcsin(sa) is 1 - 2 sa^2 (agrees with
  // cos(2.0 * asin(sa)) to 2 ulp; it multiplies chi_h, no cancellation follows) -- asin + cos are ~130 instructions
  const double cos2ThetaB = fma(-2.0 * sa, sa, 1.0);
  // alpha = (H2/2 - k0H) / k02 + chi0/2 * (1/b - 1)
  const double a1 = XRTB_AMP_DIV(H2 / 2.0 - k0H, k02);
  const cplx half_chi0 = cplx{chi0.re * 0.5, chi0.im * 0.5};  // cdiv_re(chi0, 2.0): scl = 1 / 2 exactly
  const cplx a2 = cmul_np(half_chi0, cplx{XRTB_AMP_DIV(1.0, b) - 1.0, 0.0});
  const cplx alpha = cplx{a1 + a2.re, 0.0 + a2.im};
  rs = bragg_one_pol(1.0, alpha, chih, chih_, b, m.t, m.geom, chi0, k02, kHs, k0s);
  rp = bragg_one_pol(cos2ThetaB, alpha, chih, chih_, b, m.t, m.geom, chi0, k02, kHs, k0s);
}

// ---------------------------------------------------------------- f2: multilayer (materials/multilayer.py)
// one step of the Parratt recursion (multilayer.py:453-456): (rij + rj*p2) / (1 + rij*(rj*p2))
__device__ __forceinline__ cplx parratt_step(cplx rij, cplx rj, cplx p2) {
  const cplx rj2 = cmul_np(rj, p2);
  const cplx den = cmul_np(rij, rj2);
  return cdiv(cadd(rij, rj2), cplx{1.0 + den.re, 0.0 + den.im});
}

// p2i = exp(0.5j * (Q * d))**2 (multilayer.py:440-450)
__device__ __forceinline__ cplx layer_phase(cplx Q, double d) {
  const cplx iQT = cmul_np(Q, cplx{d, 0.0});
  const cplx p1 = cexp_np(cmul_np(cplx{0.0, 0.5}, iQT));
  return csquare_np(p1);
}

// the loop `for i in reversed(range(2*nPairs))` of Multilayer.get_amplitude, geom 'reflected' (multilayer.py:432-460)
// == kernel get_amplitude_graded_multilayer (cl/materials.cl:568-622)
__device__ __forceinline__ void multilayer_recursion(int npairs, int uniform, cplx rbs_s, cplx rbs_p, cplx rtb_s,
                                                     cplx rtb_p, cplx rvt_s, cplx rvt_p, cplx Qt, cplx Qb,
                                                     const double* __restrict__ dti,
                                                     const double* __restrict__ dbi, cplx& ri_s, cplx& ri_p) {
  cplx rj_s = rbs_s, rj_p = rbs_p;
  const cplx rbt_s = cplx{-rtb_s.re, -rtb_s.im}, rbt_p = cplx{-rtb_p.re, -rtb_p.im};
  cplx p2t = cplx{1.0, 0.0}, p2b = cplx{1.0, 0.0};
  if (uniform && npairs > 0) {
    p2t = layer_phase(Qt, dti[0]);
    p2b = layer_phase(Qb, dbi[0]);
  }
  for (int i = 2 * npairs - 1; i >= 0; --i) {
    cplx rij_s, rij_p, p2;
    if (i & 1) {
      rij_s = rtb_s;
      rij_p = rtb_p;
      p2 = uniform ? p2b : layer_phase(Qb, dbi[i >> 1]);
    } else {
      rij_s = (i == 0) ? rvt_s : rbt_s;
      rij_p = (i == 0) ? rvt_p : rbt_p;
      p2 = uniform ? p2t : layer_phase(Qt, dti[i >> 1]);
    }
    rj_s = parratt_step(rij_s, rj_s, p2);
    rj_p = parratt_step(rij_p, rj_p, p2);
  }
  ri_s = rj_s;
  ri_p = rj_p;
  if (npairs <= 0) ri_s = ri_p = cplx{0.0, 0.0};  // ri = zeros_like(rj) when the loop does not run
}

// geom 'transmitted' (multilayer.py:419-462) == kernel get_amplitude_graded_multilayer_tran (cl/materials.cl): the
// substrate is one more layer (thickness substThickness), then vacuum; r and t recurse together:
//   ri = (rij + rj p2) / (1 + rij rj p2),  ti = tij tj p1 / (1 + rij rj p2)
struct MLTran {
  cplx rvt_s, rvt_p, tvt_s, tvt_p, rbs_s, rbs_p, tbs_s, tbs_p, rsv_s, rsv_p, tsv_s, tsv_p, rtb_s, rtb_p, ttb_s, ttb_p,
      tbt_s, tbt_p, Qt, Qb, Qs;
};
__device__ __forceinline__ void multilayer_recursion_tran(int npairs, double substThickness, const MLTran& f,
                                                          const double* __restrict__ dti,
                                                          const double* __restrict__ dbi, cplx& ti_s, cplx& ti_p) {
  cplx rj_s = f.rsv_s, rj_p = f.rsv_p, tj_s = f.tsv_s, tj_p = f.tsv_p;
  const cplx rbt_s = cplx{-f.rtb_s.re, -f.rtb_s.im}, rbt_p = cplx{-f.rtb_p.re, -f.rtb_p.im};
  ti_s = ti_p = cplx{0.0, 0.0};
  for (int i = 2 * npairs; i >= 0; --i) {
    cplx rij_s, rij_p, tij_s, tij_p, iQT;
    if (i & 1) {
      rij_s = f.rtb_s, rij_p = f.rtb_p, tij_s = f.ttb_s, tij_p = f.ttb_p;
      iQT = cmul_np(f.Qb, cplx{dbi[i >> 1], 0.0});
    } else if (i == 0) {
      rij_s = f.rvt_s, rij_p = f.rvt_p, tij_s = f.tvt_s, tij_p = f.tvt_p;
      iQT = cmul_np(f.Qt, cplx{dti[0], 0.0});
    } else if (i == 2 * npairs) {
      rij_s = f.rbs_s, rij_p = f.rbs_p, tij_s = f.tbs_s, tij_p = f.tbs_p;
      iQT = cmul_np(f.Qs, cplx{substThickness, 0.0});
    } else {
      rij_s = rbt_s, rij_p = rbt_p, tij_s = f.tbt_s, tij_p = f.tbt_p;
      iQT = cmul_np(f.Qt, cplx{dti[i >> 1], 0.0});
    }
    const cplx p1 = cexp_np(cmul_np(cplx{0.0, 0.5}, iQT));
    const cplx p2 = csquare_np(p1);
    const cplx rj2_s = cmul_np(rj_s, p2), rj2_p = cmul_np(rj_p, p2);
    const cplx ds = cmul_np(rij_s, rj2_s), dp = cmul_np(rij_p, rj2_p);
    const cplx den_s = cplx{1.0 + ds.re, 0.0 + ds.im}, den_p = cplx{1.0 + dp.re, 0.0 + dp.im};
    const cplx ri_s = cdiv(cadd(rij_s, rj2_s), den_s), ri_p = cdiv(cadd(rij_p, rj2_p), den_p);
    ti_s = cdiv(cmul_np(cmul_np(tij_s, tj_s), p1), den_s);
    ti_p = cdiv(cmul_np(cmul_np(tij_p, tj_p), p1), den_p);
    tj_s = ti_s, tj_p = ti_p;
    rj_s = ri_s, rj_p = ri_p;
  }
}

// (Qi/ni*nj - Qj/nj*ni) / (Qi/ni*nj + Qj/nj*ni) (multilayer.py:391-392, :404-405)
__device__ __forceinline__ cplx fresnel_p(cplx Qi, cplx ni, cplx Qj, cplx nj) {
  const cplx a = cmul_np(cdiv(Qi, ni), nj);
  const cplx b = cmul_np(cdiv(Qj, nj), ni);
  return cdiv(csub(a, b), cadd(a, b));
}

// Multilayer.get_amplitude, geom 'reflected', all three materials present (multilayer.py:364-508)
__device__ __forceinline__ void multilayer_amplitude(const xrt_multilayer_params& ml, const Tables& tab, double E,
                                                     double bdn, cplx& rs, cplx& rp, int* oor) {
  const double CHBAR = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8 / 6.283185307179586476925286766559;
  const double k = E / CHBAR;
  // no top layer (Coated): nt = 1. (multilayer.py:373)
  const cplx nt = ml.tLayer.nelem > 0 ? cconj(refractive_index(ml.tLayer, tab, E, oor)) : cplx{1.0, 0.0};
  const cplx nb = cconj(refractive_index(ml.bLayer, tab, E, oor));
  const cplx ns = cconj(refractive_index(ml.substrate, tab, E, oor));
  const double Q = 2.0 * k * fabs(bdn);
  const double Q2 = Q * Q;
  const double k28 = 8.0 * (k * k);
  cplx z = cmul_np(cplx{nt.re - 1.0, nt.im}, cplx{k28, 0.0});
  const cplx Qt = csqrt_(cplx{Q2 + z.re, 0.0 + z.im});
  z = cmul_np(cplx{nb.re - 1.0, nb.im}, cplx{k28, 0.0});
  const cplx Qb = csqrt_(cplx{Q2 + z.re, 0.0 + z.im});
  z = cmul_np(cplx{ns.re - 1.0, ns.im}, cplx{k28, 0.0});
  const cplx Qs = csqrt_(cplx{Q2 + z.re, 0.0 + z.im});
  const cplx id2 = cplx{ml.idThickness * ml.idThickness, 0.0};
  const cplx Qc = cplx{Q, 0.0};
  // vacuum - top
  const cplx roughvt = cexp_np(cmul_np(cmul_np(cplx{-0.5 * Q, 0.0}, Qt), id2));
  const cplx rvt_s = cmul_np(cdiv(cplx{Q - Qt.re, 0.0 - Qt.im}, cplx{Q + Qt.re, 0.0 + Qt.im}), roughvt);
  cplx a = cmul_np(Qc, nt), b = cdiv(Qt, nt);
  const cplx rvt_p = cmul_np(cdiv(csub(a, b), cadd(a, b)), roughvt);
  // top - bottom
  const cplx roughtb = cexp_np(cmul_np(cmul_np(cmul_np(cplx{-0.5, 0.0}, Qt), Qb), id2));
  const cplx rtb_s = cmul_np(cdiv(csub(Qt, Qb), cadd(Qt, Qb)), roughtb);
  const cplx rtb_p = cmul_np(fresnel_p(Qt, nt, Qb, nb), roughtb);
  // bottom - substrate: rmsbs = id2 if the top layer exists, else substRoughness**2 (multilayer.py:401)
  const cplx rmsbs = ml.tLayer.nelem > 0 ? id2 : cplx{ml.substRoughness * ml.substRoughness, 0.0};
  const cplx roughbs = cexp_np(cmul_np(cmul_np(cmul_np(cplx{-0.5, 0.0}, Qb), Qs), rmsbs));
  const cplx rbs_s = cmul_np(cdiv(csub(Qb, Qs), cadd(Qb, Qs)), roughbs);
  const cplx rbs_p = cmul_np(fresnel_p(Qb, nb, Qs, ns), roughbs);
  if (ml.transmitted) {
    // the transmission coefficients of every interface (multilayer.py:388-417); numpy: 2*Q / (Q+Qt) * roughvt etc.
    MLTran f;
    f.rvt_s = rvt_s, f.rvt_p = rvt_p, f.rbs_s = rbs_s, f.rbs_p = rbs_p, f.rtb_s = rtb_s, f.rtb_p = rtb_p;
    f.Qt = Qt, f.Qb = Qb, f.Qs = Qs;
    const cplx two = cplx{2.0, 0.0};
    const cplx Q2c = cplx{2.0 * Q, 0.0};
    f.tvt_s = cmul_np(cdiv(Q2c, cplx{Q + Qt.re, 0.0 + Qt.im}), roughvt);
    f.tvt_p = cmul_np(cdiv(cmul_np(Q2c, nt), cadd(a, b)), roughvt);  // a = Q*nt, b = Qt/nt (above)
    const cplx tn = cmul_np(cdiv(Qt, nt), nb), bn = cmul_np(cdiv(Qb, nb), nt);  // Qt/nt*nb, Qb/nb*nt
    const cplx Qt2 = cmul_np(two, Qt), Qb2 = cmul_np(two, Qb), Qs2 = cmul_np(two, Qs);
    f.ttb_s = cmul_np(cdiv(Qt2, cadd(Qt, Qb)), roughtb);
    f.ttb_p = cmul_np(cdiv(cmul_np(cdiv(Qt2, nt), nb), cadd(tn, bn)), roughtb);
    f.tbt_s = cmul_np(cdiv(Qb2, cadd(Qt, Qb)), roughtb);
    f.tbt_p = cmul_np(cdiv(cmul_np(cdiv(Qb2, nb), nt), cadd(tn, bn)), roughtb);
    const cplx bs = cmul_np(cdiv(Qb, nb), ns), sb = cmul_np(cdiv(Qs, ns), nb);  // Qb/nb*ns, Qs/ns*nb
    f.tbs_s = cmul_np(cdiv(Qb2, cadd(Qb, Qs)), roughbs);
    f.tbs_p = cmul_np(cdiv(cmul_np(cdiv(Qb2, nb), ns), cadd(bs, sb)), roughbs);
    const cplx Qsn = cdiv(Qs, ns), Qns = cmul_np(Qc, ns);  // Qs/ns, Q*ns
    f.rsv_s = cmul_np(cdiv(cplx{Qs.re - Q, Qs.im}, cplx{Qs.re + Q, Qs.im}), roughbs);
    f.rsv_p = cmul_np(cdiv(csub(Qsn, Qns), cadd(Qsn, Qns)), roughbs);
    f.tsv_s = cmul_np(cdiv(Qs2, cplx{Qs.re + Q, Qs.im}), roughbs);
    f.tsv_p = cmul_np(cdiv(cdiv(Qs2, ns), cadd(Qsn, Qns)), roughbs);
    multilayer_recursion_tran(ml.nPairs, ml.substThickness, f, ml.dti, ml.dbi, rs, rp);
    return;
  }
  multilayer_recursion(ml.nPairs, ml.uniform, rbs_s, rbs_p, rtb_s, rtb_p, rvt_s, rvt_p, Qt, Qb, ml.dti, ml.dbi, rs,
                       rp);
  // multilayer.py:501-505: tabulation artefacts with n.real > 1 flip the sign convention.  The reference tests
  // the FIRST ray's nt (complex comparison = lexicographic); here per ray (identical for x-ray energies)
  const double nr = nt.re - 1.0;
  if (nr > 0.0 || (nr == 0.0 && nt.im > 0.0)) {
    rs = cconj(rs);
    rp = cconj(rp);
  }
}

__global__ void __launch_bounds__(256) multilayer_kernel(size_t n, int npairs, int uniform,
                                                         const double2* __restrict__ rbs_s,
                                                         const double2* __restrict__ rbs_p,
                                                         const double2* __restrict__ rtb_s,
                                                         const double2* __restrict__ rtb_p,
                                                         const double2* __restrict__ rvt_s,
                                                         const double2* __restrict__ rvt_p,
                                                         const double2* __restrict__ Qt,
                                                         const double2* __restrict__ Qb,
                                                         const double* __restrict__ dti,
                                                         const double* __restrict__ dbi, double2* __restrict__ ri_s,
                                                         double2* __restrict__ ri_p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto ld = [](const double2* p, size_t k) { return cplx{p[k].x, p[k].y}; };
  cplx s, q;
  multilayer_recursion(npairs, uniform, ld(rbs_s, i), ld(rbs_p, i), ld(rtb_s, i), ld(rtb_p, i), ld(rvt_s, i),
                       ld(rvt_p, i), ld(Qt, i), ld(Qb, i), dti, dbi, s, q);
  ri_s[i] = make_double2(s.re, s.im);
  ri_p[i] = make_double2(q.re, q.im);
}

int multilayer_device(size_t n, int npairs, int uniform, const double* rbs_s, const double* rbs_p,
                      const double* rtb_s, const double* rtb_p, const double* rvt_s, const double* rvt_p,
                      const double* Qt, const double* Qb, const double* dti, const double* dbi, double* ri_s,
                      double* ri_p, cudaStream_t st) {
  XRTB_REQUIRE(rbs_s && rbs_p && rtb_s && rtb_p && rvt_s && rvt_p && Qt && Qb && dti && dbi && ri_s && ri_p,
               "multilayer: NULL argument");
  XRTB_REQUIRE(npairs >= 0, "multilayer: nPairs must be >= 0");
  if (n == 0) return 0;
  auto c = [](const double* p) { return reinterpret_cast<const double2*>(p); };
  multilayer_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
      n, npairs, uniform, c(rbs_s), c(rbs_p), c(rtb_s), c(rtb_p), c(rvt_s), c(rvt_p), c(Qt), c(Qb), dti, dbi,
      reinterpret_cast<double2*>(ri_s), reinterpret_cast<double2*>(ri_p));
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// kernel get_amplitude_graded_multilayer_tran (cl/materials.cl; launched at multilayer.py:485-503): the 23 sliced
// complex inputs in the reference's order -- rvt_s, rvt_p, tvt_s, tvt_p, rbs_s, rbs_p, tbs_s, tbs_p, rsv_s, rsv_p,
// tsv_s, tsv_p, rbt_s, rbt_p, tbt_s, tbt_p, rtb_s, rtb_p, ttb_s, ttb_p, Qt, Qb, Qs -- as one table of device pointers
struct MLTranPtrs {
  const double2* a[23];
};
__global__ void __launch_bounds__(256) multilayer_tran_kernel(size_t n, int npairs, double substThickness,
                                                              MLTranPtrs in, const double* __restrict__ dti,
                                                              const double* __restrict__ dbi,
                                                              double2* __restrict__ ti_s, double2* __restrict__ ti_p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto ld = [&](int k) { return cplx{in.a[k][i].x, in.a[k][i].y}; };
  MLTran f;
  f.rvt_s = ld(0), f.rvt_p = ld(1), f.tvt_s = ld(2), f.tvt_p = ld(3), f.rbs_s = ld(4), f.rbs_p = ld(5);
  f.tbs_s = ld(6), f.tbs_p = ld(7), f.rsv_s = ld(8), f.rsv_p = ld(9), f.tsv_s = ld(10), f.tsv_p = ld(11);
  // 12, 13 = rbt (= -rtb, multilayer.py:394-395: the recursion forms it itself)
  f.tbt_s = ld(14), f.tbt_p = ld(15), f.rtb_s = ld(16), f.rtb_p = ld(17), f.ttb_s = ld(18), f.ttb_p = ld(19);
  f.Qt = ld(20), f.Qb = ld(21), f.Qs = ld(22);
  cplx s, q;
  multilayer_recursion_tran(npairs, substThickness, f, dti, dbi, s, q);
  ti_s[i] = make_double2(s.re, s.im);
  ti_p[i] = make_double2(q.re, q.im);
}

int multilayer_tran_device(size_t n, int npairs, double substThickness, const double* const* in23, const double* dti,
                           const double* dbi, double* ti_s, double* ti_p, cudaStream_t st) {
  XRTB_REQUIRE(in23 && dti && dbi && ti_s && ti_p && npairs >= 0, "multilayer_tran: NULL argument");
  if (n == 0) return 0;
  MLTranPtrs p;
  for (int k = 0; k < 23; ++k) {
    XRTB_REQUIRE(in23[k] != nullptr, "multilayer_tran: input %d is NULL", k);
    p.a[k] = reinterpret_cast<const double2*>(in23[k]);
  }
  multilayer_tran_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, npairs, substThickness, p, dti, dbi,
                                                                      reinterpret_cast<double2*>(ti_s),
                                                                      reinterpret_cast<double2*>(ti_p));
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- main pass
struct Decision {
  int axis, neg, brent;
  double tMin, tMax;
};

__device__ __forceinline__ Decision decide(const unsigned long long* red) {
  Decision d;
  decide_variant(red, d.axis, d.neg);
  const int var = d.axis * 2 + d.neg;
  d.tMin = dec(red[RS_VAR + var * 4 + 0]);
  d.tMax = dec(red[RS_VAR + var * 4 + 1]);
  const double mdz1 = dec(red[RS_VAR + var * 4 + 2]);
  const double mdz2 = dec(red[RS_VAR + var * 4 + 3]);
  d.brent = (mdz2 > mdz1 * 20.0) ? 1 : 0;  // base.py:871
  return d;
}

__device__ __forceinline__ int sgn(double v) { return (v > 0.0) - (v < 0.0); }

// The decision is decoded ONCE per reflect() by a one-thread kernel between the pre-passes and the main pass and left
// packed in slots 44..46; every thread of the main pass reads those three words itself (a broadcast load that runs in
// parallel with its geometry loads) instead of waiting at a block barrier for thread 0 to chase the slots.
__global__ void reflect_decide_kernel(unsigned long long* red) {
  const Decision d = decide(red);
  red[RS_DECISION] = (unsigned long long)d.axis | ((unsigned long long)d.neg << 8) | ((unsigned long long)d.brent << 16);
  red[RS_DECISION + 1] = (unsigned long long)__double_as_longlong(d.tMin);
  red[RS_DECISION + 2] = (unsigned long long)__double_as_longlong(d.tMax);
}
__device__ __forceinline__ Decision load_decision(const unsigned long long* red) {
  const unsigned long long w = red[RS_DECISION], a = red[RS_DECISION + 1], b = red[RS_DECISION + 2];
  Decision d;
  d.axis = (int)(w & 0xff);
  d.neg = (int)((w >> 8) & 0xff);
  d.brent = (int)((w >> 16) & 0xff);
  d.tMin = __longlong_as_double((long long)a);
  d.tMax = __longlong_as_double((long long)b);
  return d;
}

template <int SURF, bool REFR, int DERIV = 0>
__device__ __forceinline__ void solve_secant(const xrt_oe_params& p, const SurfDiv& sd, const Ray& r,
                                             const Decision& d, double& t1,
                                             double& t2, double& dz1, double& dz2, double& x2, double& y2,
                                             double& z2) {
  // _use_my_method (base.py:933-959).  The reference shifts (t1, dz1) <- (t2, dz2), takes the secant step from the old
  // pair and, when the new dz2 has the sign of the shifted dz1, restores the old pair.  Written here without the shift:
  // the new point always replaces (t2, dz2), and (t1, dz1) <- old (t2, dz2) only when the signs differ -- the same
  // values in the same operations, with two predicated moves per iteration instead of eight.
  int numit = 2;
  while (fabs(dz2) > kZEps && numit < kMaxIteration) {
    double t3 = t1 - XRTB_SECANT_DIV((t2 - t1) * dz1, dz2 - dz1);
    if (t3 < d.tMin) t3 = d.tMin;
    if (t3 > d.tMax) t3 = d.tMax;
    const double dz3 = find_dz<SURF, REFR, DERIV>(p, sd, t3, r, x2, y2, z2);
    // np.sign(dz2) == np.sign(dz1) (base.py:951): inside this loop the shifted dz1 (= dz2 here) is never 0 or NaN
    // (|dz2| > zEps), and when dz3 is 0 or NaN the loop ends right here with t1, dz1 unused -- so the comparison of
    // the sign bits decides every case that matters, on the integer pipe
    if ((__double2hiint(dz3) ^ __double2hiint(dz2)) < 0) {
      t1 = t2;
      dz1 = dz2;
    }
    t2 = t3;
    dz2 = dz3;
    ++numit;
  }
}

template <int SURF, bool REFR, int DERIV = 0>
__device__ __forceinline__ void solve_brent(const xrt_oe_params& p, const SurfDiv& sd, const Ray& r, double& t1,
                                            double& t2,
                                            double& dz1, double& dz2, double& x2, double& y2, double& z2) {
  // _use_Brent_method (base.py:961-1048)
  if (fabs(dz1) < fabs(dz2)) {
    double tt = t1;
    t1 = t2;
    t2 = tt;
    tt = dz1;
    dz1 = dz2;
    dz2 = tt;
  }
  double t3 = t1, dz3 = dz1, t4 = 0.0;
  bool mflag = true;
  int numit = 2;
  while (fabs(dz2) > kZEps && numit < kMaxIteration) {
    double xa = t1, xb = t2, xc = t3, xd = t4, fa = dz1, fb = dz2, fc = dz3, xs;
    if (fa != fc && fb != fc) {
      xs = xa * fb * fc / (fa - fb) / (fa - fc) + fa * xb * fc / (fb - fa) / (fb - fc) +
           fa * fb * xc / (fc - fa) / (fc - fb);
    } else {
      xs = xb - fb * (xb - xa) / (fb - fa);
    }
    const double q = (3.0 * xa + xb) / 4.0;
    const bool cond1 = ((xs < q) && (xs < xb)) || ((xs > q) && (xs > xb));
    const bool cond2 = mflag && (fabs(xs - xb) >= (fabs(xb - xc) / 2.0));
    const bool cond3 = (!mflag) && (fabs(xs - xb) >= (fabs(xc - xd) / 2.0));
    const bool cond4 = mflag && (fabs(xb - xc) < kZEps);
    const bool cond5 = (!mflag) && (fabs(xc - xd) < kZEps);
    const bool conds = cond1 || cond2 || cond3 || cond4 || cond5;
    if (conds) xs = (xa + xb) / 2.0;
    mflag = conds;
    const double fs = find_dz<SURF, REFR, DERIV>(p, sd, xs, r, x2, y2, z2);
    xd = xc;
    xc = xb;
    fc = fb;
    const bool neg = ((fa < 0.0) && (fs > 0.0)) || ((fa > 0.0) && (fs < 0.0));
    if (neg) {
      xb = xs;
      fb = fs;
    } else {
      xa = xs;
      fa = fs;
    }
    if (fabs(fa) < fabs(fb)) {
      double tt = xa;
      xa = xb;
      xb = tt;
      tt = fa;
      fa = fb;
      fb = tt;
    }
    t1 = xa;
    t2 = xb;
    t3 = xc;
    t4 = xd;
    dz1 = fa;
    dz2 = fb;
    dz3 = fc;
    ++numit;
  }
}

// ---------------------------------------------------------------- multiple_reflect: derivative stage
// _bracketing(isMulti=True) (base.py:1279-1289): before the next intersection is searched, the ray has to leave the
// surface -- find_intersection(local_n, tMin = 0, tMax, derivOrder = 1) finds the root of (ray direction . surface
// normal) along the ray.  Phase 1 reduces that search's beam-wide scalars (t2.max(), the secant-vs-Brent switch) into
// the RS_DERIV slots, phase 2 solves per ray and stores the root, which the z stage then uses as its tMin.
template <int SURF, int PHASE>
__global__ void __launch_bounds__(256) multi_deriv_kernel(size_t n, xrt_oe_params p, xrt_beam in,
                                                          unsigned long long* red, double* __restrict__ tD) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  int axis, neg;
  decide_variant(red, axis, neg);
  double v[4] = {inf, -inf, -inf, -inf};
  Decision d;
  if (PHASE == 2) {
    d.axis = axis;
    d.neg = neg;
    d.tMin = dec(red[RS_DERIV + 0]);
    d.tMax = dec(red[RS_DERIV + 1]);
    d.brent = (dec(red[RS_DERIV + 3]) > dec(red[RS_DERIV + 2]) * 20.0) ? 1 : 0;
  }
  const SurfDiv sd = make_surf_div<SURF>(p);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int st = in.state[i];
    if (!(p.good_mode ? (st == 1 || st == 2) : (st > 0))) continue;
    Ray r = load_ray(in, i);
    to_local(p, r);
    double t1, t2, x, y, z;
    bracket(p, axis, neg, r, t1, t2);
    t1 = 0.0;  // tMin[:] = 0 (base.py:1280)
    double dz1 = find_dz<SURF, true, 1>(p, sd, t1, r, x, y, z);
    double dz2 = find_dz<SURF, true, 1>(p, sd, t2, r, x, y, z);
    const bool ind1 = dz1 <= 0.0, ind2 = dz2 >= 0.0;
    if (ind1 || ind2) dz2 = 0.0;
    if (PHASE == 1) {
      v[0] = (t1 < v[0] || isnan(t1)) ? t1 : v[0];
      v[1] = (t2 > v[1] || isnan(t2)) ? t2 : v[1];
      v[2] = (fabs(dz1) > v[2] || isnan(dz1)) ? fabs(dz1) : v[2];
      v[3] = (fabs(dz2) > v[3] || isnan(dz2)) ? fabs(dz2) : v[3];
    } else {
      if (ind1) t2 = t1;
      if (!(ind1 || ind2)) {
        if (d.brent)
          solve_brent<SURF, true, 1>(p, sd, r, t1, t2, dz1, dz2, x, y, z);
        else
          solve_secant<SURF, true, 1>(p, sd, r, d, t1, t2, dz1, dz2, x, y, z);
      }
      tD[i] = t2;
    }
  }
  if (PHASE == 2) return;
  __shared__ unsigned long long sh[8][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned long long x = enc(v[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, o);
      x = (k == 0) ? (y < x ? y : x) : (y > x ? y : x);
    }
    if (lane == 0) sh[warp][k] = x;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const int k = threadIdx.x;
    unsigned long long x = sh[0][k];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      const unsigned long long y = sh[w][k];
      x = (k == 0) ? (y < x ? y : x) : (y > x ? y : x);
    }
    if (k == 0)
      atomicMin(&red[RS_DERIV + k], x);
    else
      atomicMax(&red[RS_DERIV + k], x);
  }
}

// needElevationMap (reflect.py:651-659; base.py:1285-1287): find_dz(None, tMax, ...) at the derivative root tD = the
// height of the ray above the surface where it is farthest from it, and that point (Cartesian for parametric surfaces)
template <int SURF>
__global__ void __launch_bounds__(256) multi_elevation_kernel(size_t n, xrt_oe_params p, xrt_beam in,
                                                              const double* __restrict__ tD, double* __restrict__ eD,
                                                              double* __restrict__ eX, double* __restrict__ eY,
                                                              double* __restrict__ eZ) {
  const SurfDiv sd = make_surf_div<SURF>(p);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int st = in.state[i];
    if (!(p.good_mode ? (st == 1 || st == 2) : (st > 0))) continue;
    Ray r = load_ray(in, i);
    to_local(p, r);
    double x, y, z;
    const double dz = find_dz<SURF, true, 0>(p, sd, tD[i], r, x, y, z);
    param_to_xyz<SURF>(p, x, y, z);
    eD[i] = dz;
    eX[i] = x;
    eY[i] = y;
    eZ[i] = z;
  }
}

__device__ __forceinline__ cplx ld_c(const double* p, size_t i) {
  const double2 v = reinterpret_cast<const double2*>(p)[i];
  return cplx{v.x, v.y};
}
__device__ __forceinline__ void st_c(double* p, size_t i, cplx v) {
  reinterpret_cast<double2*>(p)[i] = make_double2(v.re, v.im);
}

__device__ __forceinline__ void copy_ray(const xrt_beam& dst, const xrt_beam& src, size_t i, int state,
                                         bool withState) {
  dst.x[i] = src.x[i];
  dst.y[i] = src.y[i];
  dst.z[i] = src.z[i];
  dst.a[i] = src.a[i];
  dst.b[i] = src.b[i];
  dst.c[i] = src.c[i];
  dst.path[i] = src.path[i];
  dst.E[i] = src.E[i];
  dst.Jss[i] = src.Jss[i];
  dst.Jpp[i] = src.Jpp[i];
  st_c(dst.Jsp, i, ld_c(src.Jsp, i));
  if (dst.Es && src.Es) {
    st_c(dst.Es, i, ld_c(src.Es, i));
    st_c(dst.Ep, i, ld_c(src.Ep, i));
  }
  if (withState) dst.state[i] = state;
}

#ifndef XRTB_REFLECT_MINBLOCKS
#define XRTB_REFLECT_MINBLOCKS 2
#endif
#ifndef XRTB_REFLECT_THREADS
#define XRTB_REFLECT_THREADS 512
#endif
// r02 structure.  One thread per ray, every global access with unit stride.  The FIELD part of the tile (path, E, Jss,
// Jpp, Jsp, Es, Ep: 80 B/ray) is fetched by bulk-async copies (cp.async.bulk, the TMA engine; completion on an mbarrier)
// into shared memory while the threads load the geometry and run the root search, and is read from there when the
// amplitudes are applied: no register holds a field across the solver (r01 spilled them -- ncu: 13 % of the long-
// scoreboard stalls sat on the STL that waited for those loads and 50 % on the LDL -> STG reload chain of the final
// stores, with 488 B/ray of DRAM traffic for 404 algorithmic), and the load latency hides behind the search.  Every
// output is stored as soon as it is final.  A block-level class sort (dense warps for the rays that reach the surface)
// was measured and rejected: profiles/r02_reflect_variants.json.
struct FieldTile {
  double path[XRTB_REFLECT_THREADS], E[XRTB_REFLECT_THREADS], Jss[XRTB_REFLECT_THREADS], Jpp[XRTB_REFLECT_THREADS];
  double2 Jsp[XRTB_REFLECT_THREADS], Es[XRTB_REFLECT_THREADS], Ep[XRTB_REFLECT_THREADS];
  // ... and the geometry (its own mbarrier: needed at once): rays that are passed through (lost, dead) take their
  // position and direction from here instead of reading them from HBM a second time (20 B/ray of DRAM reads in
  // BASELINE config 2, where 41 % of the rays miss the mirror)
  double x[XRTB_REFLECT_THREADS], y[XRTB_REFLECT_THREADS], z[XRTB_REFLECT_THREADS], a[XRTB_REFLECT_THREADS],
      b[XRTB_REFLECT_THREADS], c[XRTB_REFLECT_THREADS];
  int state[XRTB_REFLECT_THREADS];
};

// gb (or lb) of a ray that is passed through: geometry from `src` in global memory, fields from the staged tile when
// `src` is the input beam itself
__device__ __forceinline__ void copy_ray_staged(const xrt_beam& dst, const xrt_beam& src, bool srcIsInput,
                                                const FieldTile& ft, int tid, bool amp, size_t i, int state) {
  // (every load before the first store: the arrays are not __restrict__, and a load behind a store would wait for it)
  const double x = srcIsInput ? ft.x[tid] : src.x[i], y = srcIsInput ? ft.y[tid] : src.y[i],
               z = srcIsInput ? ft.z[tid] : src.z[i], a = srcIsInput ? ft.a[tid] : src.a[i],
               b = srcIsInput ? ft.b[tid] : src.b[i], c = srcIsInput ? ft.c[tid] : src.c[i];
  if (srcIsInput) {
    dst.path[i] = ft.path[tid];
    dst.E[i] = ft.E[tid];
    dst.Jss[i] = ft.Jss[tid];
    dst.Jpp[i] = ft.Jpp[tid];
    reinterpret_cast<double2*>(dst.Jsp)[i] = ft.Jsp[tid];
    if (dst.Es && amp) {
      reinterpret_cast<double2*>(dst.Es)[i] = ft.Es[tid];
      reinterpret_cast<double2*>(dst.Ep)[i] = ft.Ep[tid];
    }
  } else {
    const double path = src.path[i], E = src.E[i], Jss = src.Jss[i], Jpp = src.Jpp[i];
    const cplx Jsp = ld_c(src.Jsp, i);
    const bool hasE = dst.Es && src.Es;
    cplx Es = cplx{0.0, 0.0}, Ep = cplx{0.0, 0.0};
    if (hasE) {
      Es = ld_c(src.Es, i);
      Ep = ld_c(src.Ep, i);
    }
    dst.path[i] = path;
    dst.E[i] = E;
    dst.Jss[i] = Jss;
    dst.Jpp[i] = Jpp;
    st_c(dst.Jsp, i, Jsp);
    if (hasE) {
      st_c(dst.Es, i, Es);
      st_c(dst.Ep, i, Ep);
    }
  }
  dst.x[i] = x;
  dst.y[i] = y;
  dst.z[i] = z;
  dst.a[i] = a;
  dst.b[i] = b;
  dst.c[i] = c;
  dst.state[i] = state;
}

// Eight per-thread keys -> the beam-wide slots: keys 0..2 minima and 3..5 maxima of order-preserving encodings
// (-> red[RS_FOOT + k]), keys 6 / 7 per-thread 0 / 1 counts (-> red[slot6], red[slot7]; a negative slot is skipped).
// Warp level on REDUX (two 32-bit passes per 64-bit key instead of five 64-bit shuffle rounds); block level without
// a barrier: every warp leaves its partial results in shared memory and the LAST warp to arrive combines them, so warps
// that finish early retire instead of waiting for the slowest one with their registers allocated.
template <int BT>
__device__ __forceinline__ void tail_combine(const unsigned long long (&e)[8], unsigned long long (*sh)[8], int* sdone,
                                             unsigned long long* red, int slot6, int slot7, int lane, int warp) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const unsigned hi = (unsigned)(e[k] >> 32), lo = (unsigned)e[k];
    unsigned long long x;
    if (k >= 6) {
      x = __reduce_add_sync(0xffffffffu, lo);
    } else if (k < 3) {
      const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
      const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
      x = ((unsigned long long)mh << 32) | ml;
    } else {
      const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
      const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
      x = ((unsigned long long)mh << 32) | ml;
    }
    if (lane == 0) sh[warp][k] = x;
  }
  int last = 0;
  if (lane == 0) {
    __threadfence_block();
    last = (atomicAdd(sdone, 1) == BT / 32 - 1);
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence_block();
  if (lane < 8) {
    const int k = lane;
    unsigned long long x = *(volatile unsigned long long*)&sh[0][k];
    for (int w = 1; w < BT / 32; ++w) {
      const unsigned long long y = *(volatile unsigned long long*)&sh[w][k];
      x = (k >= 6) ? x + y : (k < 3 ? (y < x ? y : x) : (y > x ? y : x));
    }
    if (k < 3) {
      if (x != enc(inf)) atomicMin(&red[RS_FOOT + k], x);
    } else if (k < 6) {
      if (x != enc(-inf)) atomicMax(&red[RS_FOOT + k], x);
    } else if (k == 6) {
      if (x && slot6 >= 0) atomicAdd(&red[slot6], x);
    } else {
      if (x && slot7 >= 0) atomicAdd(&red[slot7], x);
    }
  }
}

// Where the global beam goes.  FUSE = false: into `gb` (OE.reflect as the reference returns it).  FUSE = true: the
// kernel continues with Screen.expose (screens.py:225-302; the stand-alone kernel is expose_kernel in rays.cu, same
// operations in the same order) and stores the SCREEN beam instead -- gb never touches HBM (132 B/ray less written,
// 132 B/ray less read back, one launch less).  The stages arrive in the order the kernel finishes them: geometry,
// path / E, coherency matrix, amplitudes.
template <bool FUSE>
struct GlobalSink {
  const xrt_beam& gb;
  const xrt_beam& scr;
  const xrt_plane_params& sp;
  double pathS, E;
  double fx, fz;  // footprint contribution of this ray (x, 0, z on the screen) when cnt == 1
  unsigned cnt;
  __device__ __forceinline__ GlobalSink(const xrt_beam& g, const xrt_beam& s, const xrt_plane_params& q)
      : gb(g), scr(s), sp(q), pathS(0.0), E(0.0), fx(0.0), fz(0.0), cnt(0) {}
  __device__ __forceinline__ void geometry(size_t i, double x, double y, double z, double a, double b, double c,
                                           int st) {
    if (!FUSE) {
      gb.x[i] = x;
      gb.y[i] = y;
      gb.z[i] = z;
      gb.a[i] = a;
      gb.b[i] = b;
      gb.c[i] = c;
      gb.state[i] = st;
      return;
    }
    // global_to_virgin_local with an explicit basis (beamline.py:251-264): every ray, whatever its state
    const double x0 = x - sp.center[0], y0 = y - sp.center[1], z0 = z - sp.center[2];
    double xs = sp.ex[0] * x0 + sp.ex[1] * y0 + sp.ex[2] * z0;
    const double ys = sp.ey[0] * x0 + sp.ey[1] * y0 + sp.ey[2] * z0;
    double zs = sp.ez[0] * x0 + sp.ez[1] * y0 + sp.ez[2] * z0;
    const double as = sp.ex[0] * a + sp.ex[1] * b + sp.ex[2] * c;
    const double bs = sp.ey[0] * a + sp.ey[1] * b + sp.ey[2] * c;
    const double cs = sp.ez[0] * a + sp.ez[1] * b + sp.ez[2] * c;
    double path = div_nr(-ys, bs);  // as expose_kernel (rays.cu)
    bool bad = isnan(path) || isinf(path);
    if (sp.only_positive_path) bad = bad || (path < 0.0);
    if (bad) {
      path = 0.0;
      st = sp.lostNum;
    }
    xs = xs + as * path;
    zs = zs + cs * path;
    if (st == 1 || st == 2) {  // footprint (screens.py:262-268), before compression
      fx = xs;
      fz = zs;
      cnt = 1;
    }
    if (sp.has_compressX) xs *= sp.compressX;
    if (sp.has_compressZ) zs *= sp.compressZ;
    scr.x[i] = xs;
    scr.y[i] = 0.0;
    scr.z[i] = zs;
    scr.a[i] = as;
    scr.b[i] = bs;
    scr.c[i] = cs;
    scr.state[i] = st;
    pathS = path;
  }
  __device__ __forceinline__ void path_energy(size_t i, double path, double En) {
    if (!FUSE) {
      gb.path[i] = path;
      gb.E[i] = En;
      return;
    }
    scr.path[i] = path + pathS;
    scr.E[i] = En;
    E = En;
  }
  __device__ __forceinline__ void coherency(size_t i, double Jss, double Jpp, cplx Jsp) {
    const xrt_beam& o = FUSE ? scr : gb;
    o.Jss[i] = Jss;
    o.Jpp[i] = Jpp;
    st_c(o.Jsp, i, Jsp);
  }
  __device__ __forceinline__ bool wants_amplitudes() const { return (FUSE ? scr.Es : gb.Es) != nullptr; }
  __device__ __forceinline__ void amplitudes(size_t i, cplx Es, cplx Ep) {
    if (!FUSE) {
      st_c(gb.Es, i, Es);
      st_c(gb.Ep, i, Ep);
      return;
    }
    // propPhase = exp(1e7j * (E/CHBAR) * path) (screens.py:257-260)
    const double CHBAR = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8 / 6.283185307179586476925286766559;
    const double phi = (1e7 * div_nr(E, CHBAR)) * pathS;
    double sn, cs;
    sincos_big(phi, sn, cs);
    const cplx ph = cplx{cs, sn};
    st_c(scr.Es, i, cmul_np(Es, ph));
    st_c(scr.Ep, i, cmul_np(Ep, ph));
  }
};

// gb of a ray that is passed through (see copy_ray_staged), through the sink
template <bool FUSE>
__device__ __forceinline__ void pass_ray(GlobalSink<FUSE>& sink, const xrt_beam& src, bool srcIsInput,
                                         const FieldTile& ft, int tid, bool amp, size_t i, int state) {
  const double x = srcIsInput ? ft.x[tid] : src.x[i], y = srcIsInput ? ft.y[tid] : src.y[i],
               z = srcIsInput ? ft.z[tid] : src.z[i], a = srcIsInput ? ft.a[tid] : src.a[i],
               b = srcIsInput ? ft.b[tid] : src.b[i], c = srcIsInput ? ft.c[tid] : src.c[i];
  double path, E, Jss, Jpp;
  cplx Jsp, Es = cplx{0.0, 0.0}, Ep = cplx{0.0, 0.0};
  const bool hasE = sink.wants_amplitudes() && (srcIsInput ? amp : src.Es != nullptr);
  if (srcIsInput) {
    path = ft.path[tid];
    E = ft.E[tid];
    Jss = ft.Jss[tid];
    Jpp = ft.Jpp[tid];
    Jsp = cplx{ft.Jsp[tid].x, ft.Jsp[tid].y};
    if (hasE) {
      Es = cplx{ft.Es[tid].x, ft.Es[tid].y};
      Ep = cplx{ft.Ep[tid].x, ft.Ep[tid].y};
    }
  } else {
    path = src.path[i];
    E = src.E[i];
    Jss = src.Jss[i];
    Jpp = src.Jpp[i];
    Jsp = ld_c(src.Jsp, i);
    if (hasE) {
      Es = ld_c(src.Es, i);
      Ep = ld_c(src.Ep, i);
    }
  }
  sink.geometry(i, x, y, z, a, b, c, state);
  sink.path_energy(i, path, E);
  sink.coherency(i, Jss, Jpp, Jsp);
  if (hasE) sink.amplitudes(i, Es, Ep);
}

template <int SURF, int MAT, bool FUSE = false>
__global__ void __launch_bounds__(XRTB_REFLECT_THREADS, XRTB_REFLECT_MINBLOCKS) reflect_main_kernel(size_t n, xrt_oe_params p, xrt_material_params m,
                                                           xrt_multilayer_params ml, Tables tab, xrt_beam in,
                                                           xrt_beam fb, xrt_beam gb, xrt_beam lb,
                                                           unsigned long long* red, int bulk_ok, double cosRoll,
                                                           double sinRoll, xrt_plane_params scrp, xrt_beam scr,
                                                           unsigned long long* sred) {
  constexpr int BT = XRTB_REFLECT_THREADS;
  GlobalSink<FUSE> sink(gb, scr, scrp);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const size_t base = (size_t)blockIdx.x * BT;
  const size_t i = base + tid;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double fmin3[3] = {inf, inf, inf}, fmax3[3] = {-inf, -inf, -inf};
  unsigned long long ngoodN = 0;
  int oor = 0;
  int sideFlags = 0;  // asymmetric crystal: bit0 the ray met the planes with n_B.k < 0, bit1 with n_B.k >= 0
  const bool hasLb = lb.x != nullptr;
  const bool amp = in.Es != nullptr;
  const bool fbIsInput = fb.x == in.x;
  extern __shared__ __align__(16) unsigned char sdyn_raw[];
  FieldTile& ft = *reinterpret_cast<FieldTile*>(sdyn_raw);
  __shared__ __align__(8) uint64_t sbar, sbarG;
  __shared__ int shint[34];  // f1/f2 interval guesses per element table, [32] "changed" flag, [33] warps done
  // full tiles of 16-byte aligned arrays go through the bulk-copy engine; the last (ragged) tile and beams that are
  // views at odd offsets are staged with plain loads (each thread only ever reads its own slot)
  const bool bulk = bulk_ok && (base + BT <= n);
  if (tid == 0 && bulk) {
    mbar_init(&sbar, 1);
    mbar_init(&sbarG, 1);
    mbar_fence_init();
    mbar_expect_tx(&sbarG, (uint32_t)(BT * (6 * 8 + 4)));
    bulk_g2s(ft.state, in.state + base, BT * 4, &sbarG);
    bulk_g2s(ft.x, in.x + base, BT * 8, &sbarG);
    bulk_g2s(ft.y, in.y + base, BT * 8, &sbarG);
    bulk_g2s(ft.z, in.z + base, BT * 8, &sbarG);
    bulk_g2s(ft.a, in.a + base, BT * 8, &sbarG);
    bulk_g2s(ft.b, in.b + base, BT * 8, &sbarG);
    bulk_g2s(ft.c, in.c + base, BT * 8, &sbarG);
    mbar_expect_tx(&sbar, (uint32_t)(BT * (4 * 8 + (amp ? 3 : 1) * 16)));
    bulk_g2s(ft.path, in.path + base, BT * 8, &sbar);
    bulk_g2s(ft.E, in.E + base, BT * 8, &sbar);
    bulk_g2s(ft.Jss, in.Jss + base, BT * 8, &sbar);
    bulk_g2s(ft.Jpp, in.Jpp + base, BT * 8, &sbar);
    bulk_g2s(ft.Jsp, in.Jsp + 2 * base, BT * 16, &sbar);
    if (amp) {
      bulk_g2s(ft.Es, in.Es + 2 * base, BT * 16, &sbar);
      bulk_g2s(ft.Ep, in.Ep + 2 * base, BT * 16, &sbar);
    }
  }
  __shared__ int sdone, sdoneS, sdoneB;  // warps that delivered their tail results (OE slots, fused screen slots, sides)
  if (tid == 0) sdone = sdoneS = sdoneB = 0;  // (the block barrier below orders this against the tail's atomicAdd)
  int* const ghint = tab.hint;  // reduction slots 48..63
  const bool useHint = (MAT != 0 && MAT != 7) && ghint != nullptr;
  if (useHint) {
    if (tid >= BT - 32) shint[tid - (BT - 32)] = ghint[tid - (BT - 32)];
    if (tid < 2) shint[32 + tid] = 0;
    tab.hint = shint;
  } else {
    tab.hint = nullptr;
  }
  const Decision d = load_decision(red);
  if (!bulk && i < n) {
    ft.state[tid] = in.state[i];
    ft.x[tid] = in.x[i];
    ft.y[tid] = in.y[i];
    ft.z[tid] = in.z[i];
    ft.a[tid] = in.a[i];
    ft.b[tid] = in.b[i];
    ft.c[tid] = in.c[i];
    ft.path[tid] = in.path[i];
    ft.E[tid] = in.E[i];
    ft.Jss[tid] = in.Jss[i];
    ft.Jpp[tid] = in.Jpp[i];
    ft.Jsp[tid] = reinterpret_cast<const double2*>(in.Jsp)[i];
    if (amp) {
      ft.Es[tid] = reinterpret_cast<const double2*>(in.Es)[i];
      ft.Ep[tid] = reinterpret_cast<const double2*>(in.Ep)[i];
    }
  }
  // state and geometry come from the staged tile (the bulk copies were issued by thread 0 above; the barrier makes
  // the mbarrier initialisation visible before anybody waits on it)
  __syncthreads();
  if (bulk) mbar_wait(&sbarG, 0);
  int st0 = 0;
  Ray r;
  r.x = r.y = r.z = r.a = r.b = r.c = 0.0;
  if (i < n) {
    st0 = ft.state[tid];
    r.x = ft.x[tid];
    r.y = ft.y[tid];
    r.z = ft.z[tid];
    r.a = ft.a[tid];
    r.b = ft.b[tid];
    r.c = ft.c[tid];
  }
  const bool alive = i < n && (p.good_mode ? (st0 == 1 || st0 == 2) : (st0 > 0));
  // the geometric part that every traced ray shares runs BEFORE the block barrier, which then only has to publish
  // the mbarrier and the hints thread 0 / the last warp set up long ago
  const SurfDiv sd = make_surf_div<SURF>(p);
  double t1 = 0.0, t2 = 0.0, x1 = 0.0, y1 = 0.0, z1 = 0.0, x2 = 0.0, y2 = 0.0, z2 = 0.0, dz1 = 1.0, dz2 = -1.0;
  const bool search = alive && !p.no_intersection
#ifdef XRTB_USER_INTERSECTION
                      && SURF != 4
#endif
      ;
  if (alive) to_local(p, r);
  if (search) {
    // bracketing (base.py:848-885)
    bracket(p, d.axis, d.neg, r, t1, t2);
    if (p.multi_tmin) t1 = p.multi_tmin[i] + kDs;  // multiple_reflect, passes after the first (base.py:1288)
    dz1 = find_dz<SURF, MAT == 5>(p, sd, t1, r, x1, y1, z1);
    dz2 = find_dz<SURF, MAT == 5>(p, sd, t2, r, x2, y2, z2);
  }
#ifdef XRTB_SOLVER_COMPACTION
  // (-DXRTB_SOLVER_COMPACTION: measured and NOT the default -- 13.2 ms instead of 10.9 ms per 1e8 rays in config 2, the
  // two extra block barriers, the idle warps behind them and 230 B/thread more spills cost more than the dense loop
  // saves; profiles/r02_reflect_variants.json.)
  // Root search on DENSE warps.  Only the rays whose bracket ends straddle the surface are searched (59 % in BASELINE
  // config 2: the others miss the mirror), and in thread <-> ray order the secant loop ran at 17 of 32 lanes.  The
  // search is a pure function of ten doubles (local ray, bracket, dz at its ends) that returns four (t, point): those
  // rays are ranked (ballot + warp counts), their records go to shared memory in rank order, thread k searches the
  // k-th record, and the owner picks the result up -- the loads and stores of the Beam stay in thread <-> ray order
  // (coalesced, whole sectors; a full class sort of the rays was measured and rejected for what it does to the stores).
  const bool needSolve = search && !(dz1 <= 0.0 || dz2 >= 0.0) && fabs(dz2) > kZEps;
  const unsigned solveBallot = __ballot_sync(0xffffffffu, needSolve);
  __shared__ unsigned short swc[BT / 32];
  if (lane == 0) swc[warp] = (unsigned short)__popc(solveBallot);
  __syncthreads();
  {
    double(*sx)[BT] = reinterpret_cast<double(*)[BT]>(sdyn_raw + sizeof(FieldTile));  // [10][BT]
    int before = 0, n0 = 0;
#pragma unroll
    for (int w = 0; w < BT / 32; ++w) {
      const int c = swc[w];
      if (w < warp) before += c;
      n0 += c;
    }
    const int pos = before + __popc(solveBallot & ((1u << lane) - 1u));
    if (needSolve) {
      sx[0][pos] = r.x;
      sx[1][pos] = r.y;
      sx[2][pos] = r.z;
      sx[3][pos] = r.a;
      sx[4][pos] = r.b;
      sx[5][pos] = r.c;
      sx[6][pos] = t1;
      sx[7][pos] = t2;
      sx[8][pos] = dz1;
      sx[9][pos] = dz2;
    }
    __syncthreads();
    if (tid < n0) {
      Ray q;
      q.x = sx[0][tid];
      q.y = sx[1][tid];
      q.z = sx[2][tid];
      q.a = sx[3][tid];
      q.b = sx[4][tid];
      q.c = sx[5][tid];
      double a1 = sx[6][tid], a2 = sx[7][tid], f1 = sx[8][tid], f2 = sx[9][tid], px = 0.0, py = 0.0, pz = 0.0;
      if (d.brent)
        solve_brent<SURF, MAT == 5>(p, sd, q, a1, a2, f1, f2, px, py, pz);
      else
        solve_secant<SURF, MAT == 5>(p, sd, q, d, a1, a2, f1, f2, px, py, pz);
      sx[6][tid] = a2;  // (each thread only touches its own record)
      sx[0][tid] = px;
      sx[1][tid] = py;
      sx[2][tid] = pz;
    }
    __syncthreads();
    if (needSolve) {
      t2 = sx[6][pos];
      x2 = sx[0][pos];
      y2 = sx[1][pos];
      z2 = sx[2][pos];
    }
  }
#endif

  if (i < n) {
    if (!alive) {
      // rays that are not alive pass through untouched (gb = lb = copy of the input, reflect.py:104-109)
      if (bulk) mbar_wait(&sbar, 0);
      pass_ray(sink, in, true, ft, tid, amp, i, p.dead_state_lost ? p.lostNum : st0);  // Plate: dcm.py:311-312
      if (!FUSE && p.multi && gb.theta) gb.theta[i] = 0.0;  // lb.theta = zeros_like(lb.x) in every pass (reflect.py:789)
      if (hasLb) {
        copy_ray_staged(lb, in, true, ft, tid, amp, i, st0);
        if (p.zero_dead_local) {  // dcm.py:305-310
          lb.state[i] = 0;
          lb.x[i] = 0.0;
          lb.y[i] = 0.0;
          lb.z[i] = 0.0;
        }
        if (lb.theta) lb.theta[i] = 0.0;
        if (lb.s) {
          lb.s[i] = lb.x[i];
          lb.phi[i] = lb.y[i];
          lb.r[i] = lb.z[i];
        }
      }
    } else {
      const int lo0 = -3;
      bool ind1 = false;
      if (p.no_intersection) {
        // noIntersectionSearch (reflect.py:676-681): tMax = 0, the point stays where it is
        t2 = 0.0;
        x2 = r.x;
        y2 = r.y;
        z2 = r.z;
        xyz_to_param<SURF>(p, x2, y2, z2);
#ifdef XRTB_USER_INTERSECTION
      } else if (SURF == 4) {
        // the element's own find_intersection: no bracket, no lost-by-bracket rays (4 values returned, reflect.py:686-690)
        xrtb_user::find_intersection(p.sp, r.x, r.y, r.z, r.a, r.b, r.c, t2, x2, y2, z2);
#endif
      } else {
        // intersection (base.py:848-885)
        ind1 = dz1 <= 0.0;             // lost: solution t1
        const bool ind2 = dz2 >= 0.0;  // over: solution t2
        if (ind1 || ind2) dz2 = 0.0;
        if (ind1) {
          t2 = t1;
          x2 = x1;
          y2 = y1;
          z2 = z1;
        }
#ifndef XRTB_SOLVER_COMPACTION
        if (!(ind1 || ind2)) {
          if (d.brent)
            solve_brent<SURF, MAT == 5>(p, sd, r, t1, t2, dz1, dz2, x2, y2, z2);
          else
            solve_secant<SURF, MAT == 5>(p, sd, r, d, t1, t2, dz1, dz2, x2, y2, z2);
        }
#endif
      }
      if (bulk) mbar_wait(&sbar, 0);  // the field tile landed long ago
      const double tMax = t2;
      // Stores are issued as soon as a value is final (geometry first, fields last): outputs never alias inputs
      // (include/xrt_b200.h), and the short live ranges keep the 35 output values out of local memory.
      // lb.x,y,z = intersection point (parametric: s, phi, r)
      const double lx = x2, ly = y2, lz = z2;
      double cx = lx, cy = ly, cz = lz;  // Cartesian copy for rays_good
      param_to_xyz<SURF>(p, cx, cy, cz);
      // rays_good, shape 'rect' (base.py:1115-1139)
      int st = 1;
      if (p.shape_round) {
        // shape 'round' (base.py:1140-1156): an ellipse inscribed in limPhysX x limPhysY; outside -> lost.
        // sr: centerX, radiusX, centerY, radiusY formed by the host like the reference does
        const double ex = (cx - p.shape_r[0]) / p.shape_r[1], ey = (cy - p.shape_r[2]) / p.shape_r[3];
        if (ex * ex + ey * ey > 1.0) st = p.lostNum;
      } else {
        if (p.has_optX && (((p.physX[0] <= cx) && (cx < p.optX[0])) || ((p.optX[1] <= cx) && (cx < p.physX[1])))) st = 2;
        if (p.has_optY && (((p.physY[0] <= cy) && (cy < p.optY[0])) || ((p.optY[1] <= cy) && (cy < p.physY[1])))) st = 2;
        const bool outside = (cx < p.physX[0]) || (cx > p.physX[1]) || (cy < p.physY[0]) || (cy > p.physY[1]);
        bool over = false;
        if (p.over_mask & 1) over = over || (cx < p.physX[0]);
        if (p.over_mask & 2) over = over || (cx > p.physX[1]);
        if (p.over_mask & 4) over = over || (cy < p.physY[0]);
        if (p.over_mask & 8) over = over || (cy > p.physY[1]);
        if (outside) st = p.lostNum;
        if (over) st = 3;
      }
      double fzp_gx = 0.0, fzp_gy = 0.0;
      if ((MAT == 6 || MAT == 7) && p.grating == 4) {
        // NormalFZP.rays_good_gn (oes/gratings.py:120-137): zone index from r through scipy's linear interp1d
        // (rn -> zones, then astype(int)); rays in opaque zones, outside the last zone or not 'good' are lost; the
        // others get g = -(x, y)/r * rho with rho = 1 / (i_to_r(i+1) - i_to_r(i-1)), i_to_r the linear interp1d
        // zones -> rn evaluated AT the nodes: (rn[j] - rn[j-1]) + rn[j-1], and 0 outside [0, N]
        const double* __restrict__ rn = p.fzp_rn;
        const int nz = p.fzp_n;  // N + 1 radii
        const double r = sqrt(cx * cx + cy * cy);
        int lo = 0, hi = nz;  // searchsorted(rn, r, side='left')
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (rn[mid] < r) lo = mid + 1; else hi = mid;
        }
        int idx = lo < 1 ? 1 : (lo > nz - 1 ? nz - 1 : lo);
        double zi_f = (1.0 / (rn[idx] - rn[idx - 1])) * (r - rn[idx - 1]) + (double)(idx - 1);
        if (r > rn[nz - 1] || r < rn[0]) zi_f = 0.0;  // bounds_error=False, fill_value=0
        const int zi = (int)zi_f;
        const bool zgood = ((zi % 2) == p.fzp_black) && (r < rn[nz - 1]) && (st == 1);
        if (!zgood) {
          st = p.lostNum;
        } else {
          const double r_up = (rn[zi + 1] - rn[zi]) + rn[zi];  // zi + 1 <= N here
          const double r_dn = zi - 1 < 0 ? 0.0 : (zi - 1 == 0 ? rn[0] : (rn[zi - 1] - rn[zi - 2]) + rn[zi - 2]);
          const double rho = 1.0 / (r_up - r_dn);
          fzp_gx = -cx / r * rho;
          fzp_gy = -cy / r * rho;
        }
      }
      if (ind1) st = p.lostNum;  // reflect.py:706-707
      if (hasLb) {
        if (lb.s) {
          lb.s[i] = lx;
          lb.phi[i] = ly;
          lb.r[i] = lz;
        }
        lb.x[i] = cx;
        lb.y[i] = cy;
        lb.z[i] = cz;
        lb.state[i] = st;
      }

      double la = r.a, lbb = r.b, lc = r.c;
      double theta = 0.0;
      double bdn = 0.0, bds = 0.0, bos = 0.0, cY = 1.0, sY = 0.0;
      cplx ras5 = cplx{1.0, 0.0}, rap5 = cplx{1.0, 0.0};  // MAT 5: transmittivity, absorption, n.real * k
      double mu0 = 0.0, nk0 = 0.0;
      if (st == 1) {
        ++ngoodN;
        double nrm[6];
        local_normal<SURF>(p, sd, lx, ly, nrm);
#ifdef XRTB_FIGURE_ERROR
        if (p.fe_c) {
          // local_n_distorted -> (d_pitch, d_roll) = (atan(dz/dy), -atan(dz/dx)); the surface normal is turned by
          // rotate_x, then rotate_y (reflect.py:767-776); with one normal it is also the normal of the planes
          double fx = lx, fy = ly;
          if (SURF == 2 || SURF == 5) {  // slopes at the Cartesian image of (s, phi, local_r(s, phi)) (base.py:765-768)
            double fz = local_surface<SURF>(p, sd, lx, ly);
            param_to_xyz<SURF>(p, fx, fy, fz);
          }
          const double aa = fe_eval(p, fx, fy, 2) * 1e-6, bb = fe_eval(p, fx, fy, 1) * 1e-6;
          double sX, cX, sYY, cYY;
          sincos(atan(bb), &sX, &cX);
          sincos(-atan(aa), &sYY, &cYY);
          const double n4 = cX * nrm[4] - sX * nrm[5], n5 = sX * nrm[4] + cX * nrm[5];
          nrm[4] = n4;
          const double n3 = cYY * nrm[3] + sYY * n5;
          nrm[5] = -sYY * nrm[3] + cYY * n5;
          nrm[3] = n3;
          if (!p.asymmetric) {
            nrm[0] = nrm[3];
            nrm[1] = nrm[4];
            nrm[2] = nrm[5];
          }
        }
#endif
        bdn = la * nrm[0] + lbb * nrm[1] + lc * nrm[2];
        if (bdn < -1.0) bdn = -1.0;
        if (bdn > 1.0) bdn = 1.0;
        if ((hasLb && lb.theta) || (p.multi && gb.theta)) theta = acos(bdn) - kPi / 2;  // the local beam's theta
        bds = p.asymmetric ? (la * nrm[3] + lbb * nrm[4] + lc * nrm[5]) : bdn;
        double ao, bo, co;
        const bool transmitted = ((MAT == 3) && (m.geom & 2)) || (MAT == 4 && ml.transmitted);  // toWhere = 2 (reflect.py:749-757, :862-878)
        if ((MAT == 3 || MAT == 4) && !transmitted && !(MAT == 4 && ml.specular)) {
          // crystal / multilayer as a grating (reflect.py:568-612, :451-469); sgbdn is taken per ray here -- all
          // good rays hit the same side of the crystal in the supported geometries (documented in DESIGN.md)
          const double ndsn = nrm[0] * nrm[3] + nrm[1] * nrm[4] + nrm[2] * nrm[5];
          const double sgbdn = (bdn < 0.0) ? 1.0 : -1.0;
          if (p.asymmetric) sideFlags = (bdn < 0.0) ? 1 : 2;  // which side of the planes this ray met (tail: slots 37 / 38)
          const double wHd = (1.0 + 0.0) / ((MAT == 4 ? ml.d : m.d) * 1e-7);
          const double g0 = (nrm[0] - ndsn * nrm[3]) * wHd * sgbdn;
          const double g1 = (nrm[1] - ndsn * nrm[4]) * wHd * sgbdn;
          const double g2 = (nrm[2] - ndsn * nrm[5]) * wHd * sgbdn;
          const double CH = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8;
          const double bdg = la * g0 + lbb * g1 + lc * g2;
          const double G2 = g0 * g0 + g1 * g1 + g2 * g2;
          const double ol = XRTB_AMP_DIV(1.0 * CH, ft.E[tid]) * 1e-7;
          const double u = bds * bds - 2.0 * bdg * ol - G2 * (ol * ol);
          const double sg = (MAT == 3 && (m.geom & 1)) ? 1.0 : -1.0;  // reflect.py:608: +1 for Laue
          const double dn = bds + sg * XRTB_AMP_SQRT(fabs(u));
          ao = la - nrm[3] * dn + g0 * ol;
          bo = lbb - nrm[4] * dn + g1 * ol;
          co = lc - nrm[5] * dn + g2 * ol;
          const double nn = XRTB_AMP_SQRT(ao * ao + bo * bo + co * co);  // a direction: nn > 0
          ao = XRTB_AMP_DIV(ao, nn);
          bo = XRTB_AMP_DIV(bo, nn);
          co = XRTB_AMP_DIV(co, nn);
          bos = ao * nrm[3] + bo * nrm[4] + co * nrm[5];
        } else if (MAT == 5) {
          // toWhere = 1, refraction (reflect.py:894-919); the transmittivity is evaluated here once (same n)
          double nre;
          plate_amplitude(m, tab, ft.E[tid], bdn, p.from_vacuum != 0, ras5, rap5, mu0, nk0, nre, &oor);
          const double n1overn2 = p.from_vacuum ? 1.0 / nre : nre;
          const double signN = (-bdn > 0.0) ? 1.0 : ((-bdn < 0.0) ? -1.0 : 0.0);
          const double n1c = -n1overn2 * bdn;
          const double cosTheta2 = signN * sqrt(1.0 - n1overn2 * n1overn2 + n1c * n1c);
          const double dn = n1c - cosTheta2;
          ao = la * n1overn2 + nrm[0] * dn;
          bo = lbb * n1overn2 + nrm[1] * dn;
          co = lc * n1overn2 + nrm[2] * dn;
        } else if (MAT == 6 || MAT == 7) {
          // toWhere = 3 (reflect.py:840-861): g = OE.local_g(x, y) (base.py:688-717), then _grating_deflection
          // (reflect.py:451-469) with the surface normal, beamInDotSurfaceNormal, the order and sig = -1
          double g0 = 0.0, g1 = 0.0, g2 = 0.0;
          if (p.grating == 4) {  // FZP: the radial groove vector found with the zone (toWhere = 4)
            g0 = fzp_gx;
            g1 = fzp_gy;
          } else if (p.grating == 3) {
            g1 = p.g_rho0;
          } else {
            const double coord = (p.grating == 1) ? cx : cy;
            double poly = 0.0;
            for (int ic = 0; ic < p.g_ncoef; ++ic) {
              // (ic+1) * coeff * coord**ic: numpy takes x**0 = 1, x**1 = x, x**2 = x*x and pow() beyond
              const double pw = ic == 0 ? 1.0 : ic == 1 ? coord : ic == 2 ? coord * coord : pow(coord, (double)ic);
              poly += ((double)(ic + 1) * p.g_coef[ic]) * pw;
            }
            const double N = p.g_rho0 * poly;
            if (p.grating == 1) g0 = N; else g1 = N;
          }
          const double CH = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8;
          const double bdg = la * g0 + lbb * g1 + lc * g2;
          const double G2 = g0 * g0 + g1 * g1 + g2 * g2;
          const double ol = p.g_order * CH / ft.E[tid] * 1e-7;
          const double u = bds * bds - 2.0 * bdg * ol - G2 * (ol * ol);
          const double dn = bds + (p.grating == 4 ? 1.0 : -1.0) * sqrt(fabs(u));  // giveSign (reflect.py:857)
          ao = la - nrm[3] * dn + g0 * ol;
          bo = lbb - nrm[4] * dn + g1 * ol;
          co = lc - nrm[5] * dn + g2 * ol;
          const double nn = sqrt(ao * ao + bo * bo + co * co);
          ao = ao / nn;
          bo = bo / nn;
          co = co / nn;
        } else {
          ao = la - nrm[0] * 2.0 * bdn;
          bo = lbb - nrm[1] * 2.0 * bdn;
          co = lc - nrm[2] * 2.0 * bdn;
          if (MAT == 3) bos = ao * nrm[3] + bo * nrm[4] + co * nrm[5];
        }
        if (!transmitted) {  // 'transmitted': the ray passes straight, (ao, bo, co) only feeds the amplitude
          la = ao;
          lbb = bo;
          lc = co;
        }
        // polarisation frame (reflect.py:948-953; beams.py:448-479)
#ifndef XRTB_POLFRAME_ATAN2
        // cos / sin of rollAngle = roll + atan2(n_x, n_z) by the addition theorem: (n_z, n_x) / hypot(n_x, n_z) IS
        // (cos, sin) of atan2(n_x, n_z), and cos / sin of the beam-wide roll are taken once per block.  Differs from
        // numpy's sincos(float(rollAngle)) by <= 4e-16 absolute (cos / sin of the beam-wide roll come from the host) (numpy rounds the angle itself to 2e-16) -- 1e-12 is
        // the parity bar -- and replaces ~170 instructions (atan2 + sincos) by 12.
        {
          const double hx = nrm[3], hz = nrm[5];
          const double h2 = hx * hx + hz * hz;
          double cphi = 1.0, sphi = 0.0;  // atan2(0, 0) = 0
          if (h2 > 0.0) {
            double y0;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(h2));
            const double e = fma(-h2 * y0, y0, 1.0);
            const double inv = fma(y0, e * fma(0.375, e, 0.5), y0);
            cphi = hz * inv;
            sphi = hx * inv;
          }
          cY = fma(cosRoll, cphi, -(sinRoll * sphi));
          sY = fma(sinRoll, cphi, cosRoll * sphi);
        }
#else
        const double rollAngle = p.roll + atan2(nrm[3], nrm[5]);
        sincos(rollAngle, &sY, &cY);  // cos(-x) == cos(x) and sin(-x) == -sin(x) exactly (numpy and CUDA alike)
#endif
      }
      if (hasLb) {
        lb.a[i] = la;
        lb.b[i] = lbb;
        lb.c[i] = lc;
        if (lb.theta) lb.theta[i] = theta;
      }
      // virgin-local beam -> global (reflect.py:1112-1136, :125-135)
      // multiple_reflect (reflect.py:205-232): gb and lb are one beam there, so EVERY traced ray leaves in virgin-local
      // coordinates with its new state; rays that went over the edge get their entry point back (:224-227)
      const bool goodAfter = p.multi ? true : (p.created_by_diffract ? (st == 1) : (st == 1 || st == 2));
      {
        Ray v{cx, cy, cz, la, lbb, lc};
        to_virgin(p, v);
        fmin3[0] = v.x;
        fmin3[1] = v.y;
        fmin3[2] = v.z;
        fmax3[0] = v.x;
        fmax3[1] = v.y;
        fmax3[2] = v.z;
        if (p.multi && st == 3) {  // tmpX, tmpY, tmpZ: the entry point in virgin-local coordinates (after the footprint)
          double wx = ft.x[tid], wy = ft.y[tid], wz = ft.z[tid];
          if (p.from_global) {
            wx -= p.center[0];
            wy -= p.center[1];
            wz -= p.center[2];
            if (p.sinAz != 0.0) rot_axis(2, p.cosAz, p.sinAz, wx, wy, wz);
          }
          v.x = wx;
          v.y = wy;
          v.z = wz;
        }
        if (goodAfter) {
          if (p.to_global) virgin_to_global(p, v);
          sink.geometry(i, v.x, v.y, v.z, v.a, v.b, v.c, st);
          if (!FUSE && p.multi && gb.theta) gb.theta[i] = theta;
        }
      }

      // The field part of the ray comes from the staged tile, group by group, and every group is stored before the
      // next one is touched: path / E, then the amplitudes (nothing but E and three cosines live), then the coherency
      // matrix, then Es / Ep.  The arithmetic of each value is unchanged; only independent groups were reordered.
      const double E = ft.E[tid];
      {
        double path = ft.path[tid];
        if (st == 1) path += tMax;
        if (hasLb) {
          lb.path[i] = path;
          lb.E[i] = E;
        }
        if (goodAfter) {
          sink.path_energy(i, path, E);
        }
      }
      if (st == 1) {  // (then goodAfter holds in every mode)
        // amplitudes
        cplx ras = cplx{1.0, 0.0}, rap = cplx{1.0, 0.0};
        if (MAT == 1 || MAT == 2 || MAT == 6) mirror_amplitude(m, tab, E, bdn, ras, rap, &oor, lo0);
        if (MAT == 3) crystal_amplitude(m, tab, E, bds, bos, bdn, ras, rap, &oor);
        if (MAT == 4) multilayer_amplitude(ml, tab, E, bds, ras, rap, &oor);  // reflect.py:999-1003
        if (MAT == 5) {
          ras = ras5;
          rap = rap5;
        }
        if (MAT != 0 && MAT != 7) {
          if (isnan(ras.re) || isnan(ras.im)) ras = cplx{0.0, 0.0};
          if (isnan(rap.re) || isnan(rap.im)) rap = cplx{0.0, 0.0};
        }
        double att = 1.0;
        if (MAT == 5 && !p.from_vacuum) att = exp(-mu0 * tMax * 0.1);  // inside the material (reflect.py:1048-1059)
        {
          // coherency matrix into the local polarisation frame (reflect.py:948-953; beams.py:448-479)
          double Jss = ft.Jss[tid], Jpp = ft.Jpp[tid];
          cplx Jsp = cplx{ft.Jsp[tid].x, ft.Jsp[tid].y};
          {
            const double c = cY, s = -sY;
            const double c2 = c * c, s2 = s * s, cs = c * s;
            const double J0 = Jss * c2 + Jpp * s2 + 2.0 * Jsp.re * cs;
            const double J1 = Jss * s2 + Jpp * c2 - 2.0 * Jsp.re * cs;
            const double J2r = (Jpp - Jss) * cs + Jsp.re * (c2 - s2);
            Jss = J0;
            Jpp = J1;
            Jsp = cplx{J2r + (Jsp.im * 0.0), 0.0 + Jsp.im};
          }
          if (MAT != 0 && MAT != 7) {
            // reflect.py:1038-1040
            const cplx js = cmul_np(cmul_np(cplx{Jss, 0.0}, ras), cconj(ras));
            const cplx jp = cmul_np(cmul_np(cplx{Jpp, 0.0}, rap), cconj(rap));
            Jsp = cmul_np(cmul_np(Jsp, ras), cconj(rap));
            Jss = js.re;
            Jpp = jp.re;
          } else if (MAT == 7 && p.g_amp >= 0.0) {
            // grating efficiency (material.py:391-413): ras = rap = resI**0.5, a real array
            Jss = (Jss * p.g_amp) * p.g_amp;
            Jpp = (Jpp * p.g_amp) * p.g_amp;
            Jsp = cplx{(Jsp.re * p.g_amp) * p.g_amp, (Jsp.im * p.g_amp) * p.g_amp};
          } else {
            // refl = 1., 1.: J * 1. * conj(1.) with Python floats leaves the values unchanged
          }
          if (MAT == 5 && !p.from_vacuum) {
            Jss *= att;
            Jpp *= att;
            Jsp = cplx{Jsp.re * att, Jsp.im * att};
          }
          if (hasLb) {
            lb.Jss[i] = Jss;
            lb.Jpp[i] = Jpp;
            st_c(lb.Jsp, i, Jsp);
          }
          // ... and back for the virgin-local / global beam (reflect.py:1106-1110)
          const double c = cY, s = sY;
          const double c2 = c * c, s2 = s * s, cs = c * s;
          const double J2r = (Jpp - Jss) * cs + Jsp.re * (c2 - s2);
          sink.coherency(i, Jss * c2 + Jpp * s2 + 2.0 * Jsp.re * cs, Jss * s2 + Jpp * c2 - 2.0 * Jsp.re * cs,
                         cplx{J2r + (Jsp.im * 0.0), 0.0 + Jsp.im});
        }
        if (amp) {
          cplx Es = cplx{ft.Es[tid].x, ft.Es[tid].y}, Ep = cplx{ft.Ep[tid].x, ft.Ep[tid].y};
          {
            // rotate_y(Es, Ep, cosY, -sinY): (c*x + s*z, -s*x + c*z) with s = -sinY
            const double s = -sY;
            const cplx e1 = cplx{cY * Es.re + s * Ep.re, cY * Es.im + s * Ep.im};
            const cplx e2 = cplx{-s * Es.re + cY * Ep.re, -s * Es.im + cY * Ep.im};
            Es = e1;
            Ep = e2;
          }
          if (MAT != 0 && MAT != 7) {
            Es = cmul_np(Es, ras);
            Ep = cmul_np(Ep, rap);
          } else if (MAT == 7 && p.g_amp >= 0.0) {
            Es = cplx{Es.re * p.g_amp, Es.im * p.g_amp};
            Ep = cplx{Ep.re * p.g_amp, Ep.im * p.g_amp};
          }
          if (MAT == 5 && !p.from_vacuum) {
            // mPh = att**0.5 * exp(0.1j * refl[3] * tMax)
            const double sa = sqrt(att);
            const cplx ph = cexp_i((0.1 * nk0) * tMax);
            const cplx mPh = cplx{sa * ph.re, sa * ph.im};
            Es = cmul_np(Es, mPh);
            Ep = cmul_np(Ep, mPh);
          } else {
            // mPh = exp(1e7j * E/CHBAR * tMax) (reflect.py:1061-1064): im = ((1e7*E) * (1/CHBAR)) * tMax
            const double CHBAR = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8 / 6.283185307179586476925286766559;
            const double phi = ((1e7 * E) * (1.0 / CHBAR)) * tMax;
            const cplx mPh = cexp_i(phi);
            Es = cmul_np(Es, mPh);
            Ep = cmul_np(Ep, mPh);
          }
          if (hasLb && lb.Es) {
            st_c(lb.Es, i, Es);
            st_c(lb.Ep, i, Ep);
          }
          if (sink.wants_amplitudes())
            sink.amplitudes(i, cplx{cY * Es.re + sY * Ep.re, cY * Es.im + sY * Ep.im},
                            cplx{-sY * Es.re + cY * Ep.re, -sY * Es.im + cY * Ep.im});
        }
      } else {
        // alive but not good after the surface: the fields stay as they were
        if (hasLb) {
          lb.Jss[i] = ft.Jss[tid];
          lb.Jpp[i] = ft.Jpp[tid];
          reinterpret_cast<double2*>(lb.Jsp)[i] = ft.Jsp[tid];
          if (amp && lb.Es) {
            reinterpret_cast<double2*>(lb.Es)[i] = ft.Es[tid];
            reinterpret_cast<double2*>(lb.Ep)[i] = ft.Ep[tid];
          }
        }
        if (goodAfter) {
          sink.coherency(i, ft.Jss[tid], ft.Jpp[tid], cplx{ft.Jsp[tid].x, ft.Jsp[tid].y});
          if (amp && sink.wants_amplitudes())
            sink.amplitudes(i, cplx{ft.Es[tid].x, ft.Es[tid].y}, cplx{ft.Ep[tid].x, ft.Ep[tid].y});
        } else {
          // not intersected rays remain unchanged except their state (reflect.py:131-134)
          pass_ray(sink, fb, fbIsInput, ft, tid, amp, i, p.dead_state_lost ? p.lostNum : st);  // Plate: dcm.py:331-332
        }
      }
    }
  }

  // a table-interval guess changed in this block: the last warp to get here hands it to the blocks that follow
  if (useHint) {
    __syncwarp();
    int last = 0;
    if (lane == 0) last = (atomicAdd(&shint[33], 1) == BT / 32 - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last && *(volatile int*)&shint[32]) ghint[lane] = *(volatile int*)&shint[lane];
  }
  // block-level footprint / counters (skipped when nobody asked for them).  No block barrier: every warp leaves its
  // eight partial results in shared memory and the LAST warp to arrive combines them -- warps that finish early retire
  // instead of waiting for the slowest one with their registers allocated.
  if (FUSE) {
    // the screen's footprint and good-ray count (Screen.expose, screens.py:262-268) -> the screen's own slots
    __shared__ unsigned long long shs[BT / 32][8];
    unsigned long long es[8];
    es[0] = sink.cnt ? enc(sink.fx) : enc(inf);
    es[1] = sink.cnt ? enc(0.0) : enc(inf);
    es[2] = sink.cnt ? enc(sink.fz) : enc(inf);
    es[3] = sink.cnt ? enc(sink.fx) : enc(-inf);
    es[4] = sink.cnt ? enc(0.0) : enc(-inf);
    es[5] = sink.cnt ? enc(sink.fz) : enc(-inf);
    es[6] = sink.cnt;
    es[7] = 0;
    tail_combine<BT>(es, shs, &sdoneS, sred, RS_NGOOD, -1, lane, warp);
  }
  if (!p.want_footprint) return;
  if ((MAT == 3 || MAT == 4) && p.asymmetric) {
    // The reference takes the sign sgbdn of the asymmetric-cut deflection from the MEAN of n_B.k over the beam
    // (reflect.py:575-576), the kernel per ray: the same thing unless rays meet the planes from both sides.  Count the
    // two sides (slots 37, 38) so that the host can refuse that case instead of deviating silently.
    __shared__ unsigned long long shb[BT / 32][8];
    unsigned long long eb[8];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      eb[k] = enc(inf);
      eb[3 + k] = enc(-inf);
    }
    eb[6] = (sideFlags & 1) ? 1ull : 0ull;
    eb[7] = (sideFlags & 2) ? 1ull : 0ull;
    tail_combine<BT>(eb, shb, &sdoneB, red, RS_SIDE_NEG, RS_SIDE_NEG + 1, lane, warp);
  }
  __shared__ unsigned long long sh[BT / 32][8];
  unsigned long long e[8];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    e[k] = enc(fmin3[k]);
    e[3 + k] = enc(fmax3[k]);
  }
  e[6] = ngoodN;
  e[7] = (unsigned long long)oor;
  tail_combine<BT>(e, sh, &sdone, red, RS_NGOODN, RS_NGOODN + 1, lane, warp);
}

#ifndef XRTB_PLUGIN
// ================================================================ r02 "tile" structure of the main pass
// The same per-ray arithmetic as reflect_main_kernel (every expression below is taken from there, in the same order),
// with the work laid out differently:
//  * the block's 512 rays live in shared memory for the whole kernel: the input tile arrives by bulk-async copies (TMA),
//    is overwritten IN PLACE with the global beam, the local beam's own values sit next to it, and both beams leave by
//    bulk-async stores (cp.async.bulk shared -> global): whole 4-KB runs per array, no per-thread store, no address
//    arithmetic;
//  * phase 1 (thread <-> ray): frame change, bracket, find_dz at both ends -> class of the ray (searched / not
//    searched / dead);
//  * the rays are ranked by class (ballots, warp counts, two REDUX sums), and in phase 2 thread k takes the k-th ray of that
//    order: the warps that run the root search, the surface normal and the Fresnel / Bragg amplitudes are DENSE (in
//    BASELINE config 2, where 41 % of the rays miss the mirror, reflect_main_kernel ran them at 19 of 32 lanes), and
//    because results go to the ray's own slot of the tile the stores to HBM stay whole sectors -- the block-level sort
//    measured in profiles/r02_reflect_variants.json lost exactly there;
//  * phase 3 (thread <-> ray): the global beam's coherency matrix and amplitudes are turned back from the local
//    polarisation frame (the tile keeps the local-frame values, which are the local beam's) and stored, or, fused with
//    Screen.expose, the whole global ray goes through GlobalSink<true>.
// Built for flat / toroidal / elliptic surfaces with material kinds 0, 1, 3, single-pass reflect; everything else stays on
// reflect_main_kernel.  Which of the two runs is decided per element (launch_tile: mirrors here, crystals there -- a DCM
// traces every ray, there is nothing to compact); XRTB_REFLECT_TILE=0 / 1 in the environment forces one of them (tests
// compare the two bit for bit).  What the measurements say about this structure -- same time as reflect_main_kernel
// until the dependent chain of a ray was shortened as well -- is in DESIGN.md section 4 and
// profiles/r02_reflect_variants.json.  The header of reflect_main_kernel's tile notes ("a block-level class sort was
// measured and rejected") refers to the variant that scattered its stores to HBM; this one does not.
#ifndef XRTB_TILE_RAYS
#define XRTB_TILE_RAYS 512
#endif
#ifndef XRTB_TILE_MINBLOCKS
#define XRTB_TILE_MINBLOCKS (1024 / XRTB_TILE_RAYS)
#endif
constexpr int kTileRays = XRTB_TILE_RAYS;
struct RayTile {
  // the global beam, in place of the input
  double x[kTileRays], y[kTileRays], z[kTileRays], a[kTileRays], b[kTileRays], c[kTileRays];
  double path[kTileRays], E[kTileRays], Jss[kTileRays], Jpp[kTileRays];  // J, Es, Ep: LOCAL polarisation frame after phase 2
  double2 Jsp[kTileRays], Es[kTileRays], Ep[kTileRays];
  // the local beam's own values; la, lb, lc, theta carry (t1, t2, dz1, dz2) from phase 1 to phase 2
  double lx[kTileRays], ly[kTileRays], lz[kTileRays], la[kTileRays], lb[kTileRays], lc[kTileRays], theta[kTileRays];
  double2 cs[kTileRays];  // (cosY, sinY) of the polarisation frame; x = 2: fields as they are, x = 3: global ray from fb
  int state[kTileRays], lstate[kTileRays];
  unsigned short idx[kTileRays];  // ray slots in class order
};

__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// the point find_dz evaluates the surface at (base.py:831-839)
template <int SURF>
__device__ __forceinline__ void ray_point(const xrt_oe_params& p, const Ray& r, double t, double& x, double& y,
                                          double& z) {
  x = r.x + r.a * t;
  y = r.y + r.b * t;
  z = r.z + r.c * t;
  xyz_to_param<SURF>(p, x, y, z);
}

template <int SURF, int MAT, bool FUSE>
__global__ void __launch_bounds__(kTileRays, XRTB_TILE_MINBLOCKS)
    reflect_tile_kernel(size_t n, xrt_oe_params p, xrt_material_params m, Tables tab, xrt_beam in, xrt_beam fb,
                        xrt_beam gb, xrt_beam lb, unsigned long long* red, int bulk_ok, double cosRoll, double sinRoll,
                        xrt_plane_params scrp, xrt_beam scr, unsigned long long* sred) {
  constexpr int BT = kTileRays;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const size_t base = (size_t)blockIdx.x * BT;
  const size_t i = base + tid;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  extern __shared__ __align__(16) unsigned char sdyn_raw[];
  RayTile& T = *reinterpret_cast<RayTile*>(sdyn_raw);
  __shared__ __align__(8) uint64_t sbar, sbarG;
  __shared__ int shint[34];
  __shared__ unsigned swc[BT / 32];
  __shared__ int sdone, sdoneS, sdoneB;
  const bool hasLb = lb.x != nullptr;
  const bool amp = in.Es != nullptr;
  const bool fbIsInput = fb.x == in.x;
  const bool bulk = bulk_ok && (base + BT <= n);
  if (tid == 0 && bulk) {
    mbar_init(&sbar, 1);
    mbar_init(&sbarG, 1);
    mbar_fence_init();
    mbar_expect_tx(&sbarG, (uint32_t)(BT * (6 * 8 + 4)));
    bulk_g2s(T.state, in.state + base, BT * 4, &sbarG);
    bulk_g2s(T.x, in.x + base, BT * 8, &sbarG);
    bulk_g2s(T.y, in.y + base, BT * 8, &sbarG);
    bulk_g2s(T.z, in.z + base, BT * 8, &sbarG);
    bulk_g2s(T.a, in.a + base, BT * 8, &sbarG);
    bulk_g2s(T.b, in.b + base, BT * 8, &sbarG);
    bulk_g2s(T.c, in.c + base, BT * 8, &sbarG);
    mbar_expect_tx(&sbar, (uint32_t)(BT * (4 * 8 + (amp ? 3 : 1) * 16)));
    bulk_g2s(T.path, in.path + base, BT * 8, &sbar);
    bulk_g2s(T.E, in.E + base, BT * 8, &sbar);
    bulk_g2s(T.Jss, in.Jss + base, BT * 8, &sbar);
    bulk_g2s(T.Jpp, in.Jpp + base, BT * 8, &sbar);
    bulk_g2s(T.Jsp, in.Jsp + 2 * base, BT * 16, &sbar);
    if (amp) {
      bulk_g2s(T.Es, in.Es + 2 * base, BT * 16, &sbar);
      bulk_g2s(T.Ep, in.Ep + 2 * base, BT * 16, &sbar);
    }
  }
  if (tid == 0) sdone = sdoneS = sdoneB = 0;
  int* const ghint = tab.hint;
  const bool useHint = (MAT != 0) && ghint != nullptr;
  if (useHint) {
    if (tid >= BT - 32) shint[tid - (BT - 32)] = ghint[tid - (BT - 32)];
    if (tid < 2) shint[32 + tid] = 0;
    tab.hint = shint;
  } else {
    tab.hint = nullptr;
  }
  const Decision d = load_decision(red);
  if (!bulk && i < n) {
    T.state[tid] = in.state[i];
    T.x[tid] = in.x[i];
    T.y[tid] = in.y[i];
    T.z[tid] = in.z[i];
    T.a[tid] = in.a[i];
    T.b[tid] = in.b[i];
    T.c[tid] = in.c[i];
    T.path[tid] = in.path[i];
    T.E[tid] = in.E[i];
    T.Jss[tid] = in.Jss[i];
    T.Jpp[tid] = in.Jpp[i];
    T.Jsp[tid] = reinterpret_cast<const double2*>(in.Jsp)[i];
    if (amp) {
      T.Es[tid] = reinterpret_cast<const double2*>(in.Es)[i];
      T.Ep[tid] = reinterpret_cast<const double2*>(in.Ep)[i];
    }
  }
  __syncthreads();
  if (bulk) mbar_wait(&sbarG, 0);
  const SurfDiv sd = make_surf_div<SURF>(p);

  // ---- phase 1, thread <-> ray: class of the ray
  int cls = 0;  // 0 nothing to trace, 1 root search, 2 traced without a search
  T.cs[tid] = make_double2(2.0, 0.0);
  if (i < n) {
    const int st0 = T.state[tid];
    Ray r{T.x[tid], T.y[tid], T.z[tid], T.a[tid], T.b[tid], T.c[tid]};
    const bool alive = p.good_mode ? (st0 == 1 || st0 == 2) : (st0 > 0);
    if (alive) {
      to_local(p, r);
      double t1, t2, xx, yy, zz;
      bracket(p, d.axis, d.neg, r, t1, t2);  // base.py:848-885
      const double dz1 = find_dz<SURF, false>(p, sd, t1, r, xx, yy, zz);
      const double dz2 = find_dz<SURF, false>(p, sd, t2, r, xx, yy, zz);
      cls = (dz1 <= 0.0 || dz2 >= 0.0) ? 2 : 1;
      T.la[tid] = t1;
      T.lb[tid] = t2;
      T.lc[tid] = dz1;
      T.theta[tid] = dz2;
    } else {
      // rays that are not alive pass through untouched (gb = lb = copy of the input, reflect.py:104-109)
      T.state[tid] = p.dead_state_lost ? p.lostNum : st0;  // Plate: dcm.py:311-312
      const bool zero = p.zero_dead_local != 0;            // dcm.py:305-310
      T.lstate[tid] = zero ? 0 : st0;
      T.lx[tid] = zero ? 0.0 : r.x;
      T.ly[tid] = zero ? 0.0 : r.y;
      T.lz[tid] = zero ? 0.0 : r.z;
      T.la[tid] = r.a;
      T.lb[tid] = r.b;
      T.lc[tid] = r.c;
      T.theta[tid] = 0.0;
      if (hasLb && lb.s) {
        lb.s[i] = T.lx[tid];
        lb.phi[i] = T.ly[tid];
        lb.r[i] = T.lz[tid];
      }
    }
  }
  // rank within the class: searched rays first, then the other traced rays
  const unsigned bh = __ballot_sync(0xffffffffu, cls == 1), bc = __ballot_sync(0xffffffffu, cls == 2);
  if (lane == 0) swc[warp] = (unsigned)__popc(bh) | ((unsigned)__popc(bc) << 16);
  __syncthreads();
  int nH = 0, nC = 0;
  {
    // prefix over the (at most 32) warp counts: lane w holds warp w's packed pair, two warp-wide integer sums (REDUX)
    static_assert(BT / 32 <= 32, "one lane per warp count");
    const unsigned v = (lane < BT / 32) ? swc[lane] : 0u;
    const unsigned tot = __reduce_add_sync(0xffffffffu, v);
    const unsigned bef = __reduce_add_sync(0xffffffffu, lane < warp ? v : 0u);
    nH = (int)(tot & 0xffffu);
    nC = (int)(tot >> 16);
    const int beforeH = (int)(bef & 0xffffu), beforeC = (int)(bef >> 16);
    const unsigned lt = (1u << lane) - 1u;
    if (cls == 1) T.idx[beforeH + __popc(bh & lt)] = (unsigned short)tid;
    if (cls == 2) T.idx[nH + beforeC + __popc(bc & lt)] = (unsigned short)tid;
  }
  __syncthreads();
  if (bulk) mbar_wait(&sbar, 0);  // the field part of the tile landed long ago

  // ---- phase 2, thread k <-> k-th traced ray
  double fmin3[3] = {inf, inf, inf}, fmax3[3] = {-inf, -inf, -inf};
  unsigned long long ngoodN = 0;
  int oor = 0;
  int sideFlags = 0;
  if (tid < nH + nC) {
    const int j = T.idx[tid];
    const size_t ij = base + j;
    // (the frame change is repeated here: handing the local ray over through the tile -- six more values per ray --
    // measured 9.4 instead of 9.0 ms per 1e8 rays)
    Ray r{T.x[j], T.y[j], T.z[j], T.a[j], T.b[j], T.c[j]};
    to_local(p, r);
    double t1 = T.la[j], t2 = T.lb[j], dz1 = T.lc[j], dz2 = T.theta[j];
    // intersection (base.py:848-885)
    const bool ind1 = dz1 <= 0.0;  // lost: solution t1
    const bool ind2 = dz2 >= 0.0;  // over: solution t2
    if (ind1) t2 = t1;
    double x2, y2, z2;
    ray_point<SURF>(p, r, t2, x2, y2, z2);  // what find_dz left for that end of the bracket
    if (!(ind1 || ind2)) {
      if (d.brent)
        solve_brent<SURF, false>(p, sd, r, t1, t2, dz1, dz2, x2, y2, z2);
      else
        solve_secant<SURF, false>(p, sd, r, d, t1, t2, dz1, dz2, x2, y2, z2);
    }
    const double tMax = t2;
    const int lo0 = -3;
    const double lx = x2, ly = y2, lz = z2;
    double cx = lx, cy = ly, cz = lz;  // Cartesian copy for rays_good
    param_to_xyz<SURF>(p, cx, cy, cz);
    // rays_good (base.py:1115-1156)
    int st = 1;
    if (p.shape_round) {
      const double ex = (cx - p.shape_r[0]) / p.shape_r[1], ey = (cy - p.shape_r[2]) / p.shape_r[3];
      if (ex * ex + ey * ey > 1.0) st = p.lostNum;
    } else {
      if (p.has_optX && (((p.physX[0] <= cx) && (cx < p.optX[0])) || ((p.optX[1] <= cx) && (cx < p.physX[1])))) st = 2;
      if (p.has_optY && (((p.physY[0] <= cy) && (cy < p.optY[0])) || ((p.optY[1] <= cy) && (cy < p.physY[1])))) st = 2;
      const bool outside = (cx < p.physX[0]) || (cx > p.physX[1]) || (cy < p.physY[0]) || (cy > p.physY[1]);
      bool over = false;
      if (p.over_mask & 1) over = over || (cx < p.physX[0]);
      if (p.over_mask & 2) over = over || (cx > p.physX[1]);
      if (p.over_mask & 4) over = over || (cy < p.physY[0]);
      if (p.over_mask & 8) over = over || (cy > p.physY[1]);
      if (outside) st = p.lostNum;
      if (over) st = 3;
    }
    if (ind1) st = p.lostNum;  // reflect.py:706-707
    if (hasLb && lb.s) {
      lb.s[ij] = lx;
      lb.phi[ij] = ly;
      lb.r[ij] = lz;
    }
    T.lx[j] = cx;
    T.ly[j] = cy;
    T.lz[j] = cz;
    T.lstate[j] = st;

    double la = r.a, lbb = r.b, lc = r.c;
    double theta = 0.0;
    double bdn = 0.0, bds = 0.0, bos = 0.0, cY = 1.0, sY = 0.0;
    if (st == 1) {
      ++ngoodN;
      double nrm[6];
      local_normal<SURF>(p, sd, lx, ly, nrm);
      bdn = la * nrm[0] + lbb * nrm[1] + lc * nrm[2];
      if (bdn < -1.0) bdn = -1.0;
      if (bdn > 1.0) bdn = 1.0;
      if (hasLb && lb.theta) theta = acos(bdn) - kPi / 2;  // the local beam's theta
      bds = p.asymmetric ? (la * nrm[3] + lbb * nrm[4] + lc * nrm[5]) : bdn;
      double ao, bo, co;
      const bool transmitted = (MAT == 3) && (m.geom & 2);  // toWhere = 2 (reflect.py:749-757, :862-878)
      if (MAT == 3 && !transmitted) {
        // crystal as a grating (reflect.py:568-612, :451-469); sgbdn per ray, see reflect_main_kernel
        const double ndsn = nrm[0] * nrm[3] + nrm[1] * nrm[4] + nrm[2] * nrm[5];
        const double sgbdn = (bdn < 0.0) ? 1.0 : -1.0;
        if (p.asymmetric) sideFlags = (bdn < 0.0) ? 1 : 2;
        const double wHd = (1.0 + 0.0) / (m.d * 1e-7);
        const double g0 = (nrm[0] - ndsn * nrm[3]) * wHd * sgbdn;
        const double g1 = (nrm[1] - ndsn * nrm[4]) * wHd * sgbdn;
        const double g2 = (nrm[2] - ndsn * nrm[5]) * wHd * sgbdn;
        const double CH = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8;
        const double bdg = la * g0 + lbb * g1 + lc * g2;
        const double G2 = g0 * g0 + g1 * g1 + g2 * g2;
        const double ol = XRTB_AMP_DIV(1.0 * CH, T.E[j]) * 1e-7;
        const double u = bds * bds - 2.0 * bdg * ol - G2 * (ol * ol);
        const double sg = (m.geom & 1) ? 1.0 : -1.0;  // reflect.py:608: +1 for Laue
        const double dn = bds + sg * XRTB_AMP_SQRT(fabs(u));
        ao = la - nrm[3] * dn + g0 * ol;
        bo = lbb - nrm[4] * dn + g1 * ol;
        co = lc - nrm[5] * dn + g2 * ol;
        const double nn = XRTB_AMP_SQRT(ao * ao + bo * bo + co * co);
        ao = XRTB_AMP_DIV(ao, nn);
        bo = XRTB_AMP_DIV(bo, nn);
        co = XRTB_AMP_DIV(co, nn);
        bos = ao * nrm[3] + bo * nrm[4] + co * nrm[5];
      } else {
        ao = la - nrm[0] * 2.0 * bdn;
        bo = lbb - nrm[1] * 2.0 * bdn;
        co = lc - nrm[2] * 2.0 * bdn;
        if (MAT == 3) bos = ao * nrm[3] + bo * nrm[4] + co * nrm[5];
      }
      if (!transmitted) {
        la = ao;
        lbb = bo;
        lc = co;
      }
      // polarisation frame (reflect.py:948-953; beams.py:448-479), see reflect_main_kernel
#ifndef XRTB_POLFRAME_ATAN2
      {
        const double hx = nrm[3], hz = nrm[5];
        const double h2 = hx * hx + hz * hz;
        double cphi = 1.0, sphi = 0.0;  // atan2(0, 0) = 0
        if (h2 > 0.0) {
          double y0;
          asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(h2));
          const double e = fma(-h2 * y0, y0, 1.0);
          const double inv = fma(y0, e * fma(0.375, e, 0.5), y0);
          cphi = hz * inv;
          sphi = hx * inv;
        }
        cY = fma(cosRoll, cphi, -(sinRoll * sphi));
        sY = fma(sinRoll, cphi, cosRoll * sphi);
      }
#else
      const double rollAngle = p.roll + atan2(nrm[3], nrm[5]);
      sincos(rollAngle, &sY, &cY);
#endif
    }
    T.la[j] = la;
    T.lb[j] = lbb;
    T.lc[j] = lc;
    T.theta[j] = theta;
    // virgin-local beam -> global (reflect.py:1112-1136, :125-135)
    const bool goodAfter = p.created_by_diffract ? (st == 1) : (st == 1 || st == 2);
    {
      Ray v{cx, cy, cz, la, lbb, lc};
      to_virgin(p, v);
      fmin3[0] = v.x;
      fmin3[1] = v.y;
      fmin3[2] = v.z;
      fmax3[0] = v.x;
      fmax3[1] = v.y;
      fmax3[2] = v.z;
      if (goodAfter) {
        if (p.to_global) virgin_to_global(p, v);
        T.x[j] = v.x;
        T.y[j] = v.y;
        T.z[j] = v.z;
        T.a[j] = v.a;
        T.b[j] = v.b;
        T.c[j] = v.c;
        T.state[j] = st;
      } else {
        // not intersected rays remain unchanged except their state (reflect.py:131-134): the ray of fb
        T.state[j] = p.dead_state_lost ? p.lostNum : st;  // Plate: dcm.py:331-332
        if (!fbIsInput) {
          T.x[j] = fb.x[ij];
          T.y[j] = fb.y[ij];
          T.z[j] = fb.z[ij];
          T.a[j] = fb.a[ij];
          T.b[j] = fb.b[ij];
          T.c[j] = fb.c[ij];
          T.cs[j] = make_double2(3.0, 0.0);
        }
      }
    }
    const double E = T.E[j];
    if (st == 1) {
      T.path[j] = T.path[j] + tMax;
      // amplitudes
      cplx ras = cplx{1.0, 0.0}, rap = cplx{1.0, 0.0};
      if (MAT == 1) mirror_amplitude(m, tab, E, bdn, ras, rap, &oor, lo0);
      if (MAT == 3) crystal_amplitude(m, tab, E, bds, bos, bdn, ras, rap, &oor);
      if (MAT != 0) {
        if (isnan(ras.re) || isnan(ras.im)) ras = cplx{0.0, 0.0};
        if (isnan(rap.re) || isnan(rap.im)) rap = cplx{0.0, 0.0};
      }
      {
        // coherency matrix into the local polarisation frame (reflect.py:948-953; beams.py:448-479)
        double Jss = T.Jss[j], Jpp = T.Jpp[j];
        cplx Jsp = cplx{T.Jsp[j].x, T.Jsp[j].y};
        {
          const double c = cY, s = -sY;
          const double c2 = c * c, s2 = s * s, cs = c * s;
          const double J0 = Jss * c2 + Jpp * s2 + 2.0 * Jsp.re * cs;
          const double J1 = Jss * s2 + Jpp * c2 - 2.0 * Jsp.re * cs;
          const double J2r = (Jpp - Jss) * cs + Jsp.re * (c2 - s2);
          Jss = J0;
          Jpp = J1;
          Jsp = cplx{J2r + (Jsp.im * 0.0), 0.0 + Jsp.im};
        }
        if (MAT != 0) {
          // reflect.py:1038-1040
          const cplx js = cmul_np(cmul_np(cplx{Jss, 0.0}, ras), cconj(ras));
          const cplx jp = cmul_np(cmul_np(cplx{Jpp, 0.0}, rap), cconj(rap));
          Jsp = cmul_np(cmul_np(Jsp, ras), cconj(rap));
          Jss = js.re;
          Jpp = jp.re;
        }
        T.Jss[j] = Jss;
        T.Jpp[j] = Jpp;
        T.Jsp[j] = make_double2(Jsp.re, Jsp.im);
      }
      if (amp) {
        cplx Es = cplx{T.Es[j].x, T.Es[j].y}, Ep = cplx{T.Ep[j].x, T.Ep[j].y};
        {
          // rotate_y(Es, Ep, cosY, -sinY): (c*x + s*z, -s*x + c*z) with s = -sinY
          const double s = -sY;
          const cplx e1 = cplx{cY * Es.re + s * Ep.re, cY * Es.im + s * Ep.im};
          const cplx e2 = cplx{-s * Es.re + cY * Ep.re, -s * Es.im + cY * Ep.im};
          Es = e1;
          Ep = e2;
        }
        if (MAT != 0) {
          Es = cmul_np(Es, ras);
          Ep = cmul_np(Ep, rap);
        }
        {
          // mPh = exp(1e7j * E/CHBAR * tMax) (reflect.py:1061-1064): im = ((1e7*E) * (1/CHBAR)) * tMax
          const double CHBAR = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8 / 6.283185307179586476925286766559;
          const double phi = ((1e7 * E) * (1.0 / CHBAR)) * tMax;
          const cplx mPh = cexp_i(phi);
          Es = cmul_np(Es, mPh);
          Ep = cmul_np(Ep, mPh);
        }
        T.Es[j] = make_double2(Es.re, Es.im);
        T.Ep[j] = make_double2(Ep.re, Ep.im);
      }
      T.cs[j] = make_double2(cY, sY);
    }
  }
  // the tile is complete: make the generic-proxy writes visible to the bulk-copy engine, then everybody may read it
  fence_proxy_async_smem();
  __syncthreads();

  // ---- phase 3, thread <-> ray: what the bulk stores cannot do
  GlobalSink<FUSE> sink(gb, scr, scrp);
  if (i < n) {
    const double2 cs = T.cs[tid];
    const bool fromFb = cs.x > 2.5;
    double Jss = T.Jss[tid], Jpp = T.Jpp[tid];
    cplx Jsp = cplx{T.Jsp[tid].x, T.Jsp[tid].y};
    const bool hasE = sink.wants_amplitudes() && (fromFb ? fb.Es != nullptr : amp);
    cplx Es = cplx{0.0, 0.0}, Ep = cplx{0.0, 0.0};
    if (hasE && !fromFb) {
      Es = cplx{T.Es[tid].x, T.Es[tid].y};
      Ep = cplx{T.Ep[tid].x, T.Ep[tid].y};
    }
    double path = T.path[tid], E = T.E[tid];
    if (fromFb) {
      path = fb.path[i];
      E = fb.E[i];
      Jss = fb.Jss[i];
      Jpp = fb.Jpp[i];
      Jsp = ld_c(fb.Jsp, i);
      if (hasE) {
        Es = ld_c(fb.Es, i);
        Ep = ld_c(fb.Ep, i);
      }
    } else if (cs.x < 1.5) {
      // ... and back for the virgin-local / global beam (reflect.py:1106-1110)
      const double c = cs.x, s = cs.y;
      const double c2 = c * c, s2 = s * s, csn = c * s;
      const double J2r = (Jpp - Jss) * csn + Jsp.re * (c2 - s2);
      const double J0 = Jss * c2 + Jpp * s2 + 2.0 * Jsp.re * csn, J1 = Jss * s2 + Jpp * c2 - 2.0 * Jsp.re * csn;
      Jss = J0;
      Jpp = J1;
      Jsp = cplx{J2r + (Jsp.im * 0.0), 0.0 + Jsp.im};
      if (hasE) {
        const cplx e1 = cplx{c * Es.re + s * Ep.re, c * Es.im + s * Ep.im};
        const cplx e2 = cplx{-s * Es.re + c * Ep.re, -s * Es.im + c * Ep.im};
        Es = e1;
        Ep = e2;
      }
    }
    if (FUSE) {
      sink.geometry(i, T.x[tid], T.y[tid], T.z[tid], T.a[tid], T.b[tid], T.c[tid], T.state[tid]);
      sink.path_energy(i, path, E);
      sink.coherency(i, Jss, Jpp, Jsp);
      if (hasE) sink.amplitudes(i, Es, Ep);
    } else {
      sink.coherency(i, Jss, Jpp, Jsp);
      if (hasE) sink.amplitudes(i, Es, Ep);
      if (!bulk || !fbIsInput) sink.path_energy(i, path, E);
      if (!bulk) sink.geometry(i, T.x[tid], T.y[tid], T.z[tid], T.a[tid], T.b[tid], T.c[tid], T.state[tid]);
    }
    if (hasLb && !bulk) {
      lb.x[i] = T.lx[tid];
      lb.y[i] = T.ly[tid];
      lb.z[i] = T.lz[tid];
      lb.a[i] = T.la[tid];
      lb.b[i] = T.lb[tid];
      lb.c[i] = T.lc[tid];
      lb.state[i] = T.lstate[tid];
      if (lb.theta) lb.theta[i] = T.theta[tid];
      lb.path[i] = T.path[tid];
      lb.E[i] = T.E[tid];
      lb.Jss[i] = T.Jss[tid];
      lb.Jpp[i] = T.Jpp[tid];
      reinterpret_cast<double2*>(lb.Jsp)[i] = T.Jsp[tid];
      if (amp && lb.Es) {
        reinterpret_cast<double2*>(lb.Es)[i] = T.Es[tid];
        reinterpret_cast<double2*>(lb.Ep)[i] = T.Ep[tid];
      }
    }
  }
  if (bulk && tid == 0) {
    if (!FUSE) {
      bulk_s2g(gb.x + base, T.x, BT * 8);
      bulk_s2g(gb.y + base, T.y, BT * 8);
      bulk_s2g(gb.z + base, T.z, BT * 8);
      bulk_s2g(gb.a + base, T.a, BT * 8);
      bulk_s2g(gb.b + base, T.b, BT * 8);
      bulk_s2g(gb.c + base, T.c, BT * 8);
      bulk_s2g(gb.state + base, T.state, BT * 4);
      if (fbIsInput) {
        bulk_s2g(gb.path + base, T.path, BT * 8);
        bulk_s2g(gb.E + base, T.E, BT * 8);
      }
    }
    if (hasLb) {
      bulk_s2g(lb.x + base, T.lx, BT * 8);
      bulk_s2g(lb.y + base, T.ly, BT * 8);
      bulk_s2g(lb.z + base, T.lz, BT * 8);
      bulk_s2g(lb.a + base, T.la, BT * 8);
      bulk_s2g(lb.b + base, T.lb, BT * 8);
      bulk_s2g(lb.c + base, T.lc, BT * 8);
      bulk_s2g(lb.state + base, T.lstate, BT * 4);
      if (lb.theta) bulk_s2g(lb.theta + base, T.theta, BT * 8);
      bulk_s2g(lb.path + base, T.path, BT * 8);
      bulk_s2g(lb.E + base, T.E, BT * 8);
      bulk_s2g(lb.Jss + base, T.Jss, BT * 8);
      bulk_s2g(lb.Jpp + base, T.Jpp, BT * 8);
      bulk_s2g(lb.Jsp + 2 * base, T.Jsp, BT * 16);
      if (amp && lb.Es) {
        bulk_s2g(lb.Es + 2 * base, T.Es, BT * 16);
        bulk_s2g(lb.Ep + 2 * base, T.Ep, BT * 16);
      }
    }
    bulk_commit();
  }

  // a table-interval guess changed in this block: the last warp to get here hands it to the blocks that follow
  if (useHint) {
    __syncwarp();
    int last = 0;
    if (lane == 0) last = (atomicAdd(&shint[33], 1) == BT / 32 - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last && *(volatile int*)&shint[32]) ghint[lane] = *(volatile int*)&shint[lane];
  }
  if (FUSE) {
    __shared__ unsigned long long shs[BT / 32][8];
    unsigned long long es[8];
    es[0] = sink.cnt ? enc(sink.fx) : enc(inf);
    es[1] = sink.cnt ? enc(0.0) : enc(inf);
    es[2] = sink.cnt ? enc(sink.fz) : enc(inf);
    es[3] = sink.cnt ? enc(sink.fx) : enc(-inf);
    es[4] = sink.cnt ? enc(0.0) : enc(-inf);
    es[5] = sink.cnt ? enc(sink.fz) : enc(-inf);
    es[6] = sink.cnt;
    es[7] = 0;
    tail_combine<BT>(es, shs, &sdoneS, sred, RS_NGOOD, -1, lane, warp);
  }
  if (p.want_footprint) {
    if (MAT == 3 && p.asymmetric) {
      __shared__ unsigned long long shb[BT / 32][8];
      unsigned long long eb[8];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        eb[k] = enc(inf);
        eb[3 + k] = enc(-inf);
      }
      eb[6] = (sideFlags & 1) ? 1ull : 0ull;
      eb[7] = (sideFlags & 2) ? 1ull : 0ull;
      tail_combine<BT>(eb, shb, &sdoneB, red, RS_SIDE_NEG, RS_SIDE_NEG + 1, lane, warp);
    }
    __shared__ unsigned long long sh[BT / 32][8];
    unsigned long long e[8];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      e[k] = enc(fmin3[k]);
      e[3 + k] = enc(fmax3[k]);
    }
    e[6] = ngoodN;
    e[7] = (unsigned long long)oor;
    tail_combine<BT>(e, sh, &sdone, red, RS_NGOODN, RS_NGOODN + 1, lane, warp);
  }
  // the tile must stay until the bulk stores have read it
  if (bulk && tid == 0) bulk_wait_read_all();
}
#endif  // !XRTB_PLUGIN

// ---------------------------------------------------------------- launchers
// cp.async.bulk needs 16-byte aligned global addresses: true for whole torch / cudaMalloc allocations, not for a view
// that starts at an odd element
static int bulk_copy_ok(const xrt_beam& in) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  return al(in.path) && al(in.E) && al(in.Jss) && al(in.Jpp) && al(in.Jsp) && (in.Es == nullptr || (al(in.Es) && al(in.Ep)));
}
// the slot layout Screen.expose leaves (rays.cu, plane_reduction_init_kernel): footprint min / max, counters zero
__global__ void screen_reduction_init_kernel(unsigned long long* red) {
  const int i = threadIdx.x;
  if (i >= XRT_REDUCTION_SLOTS) return;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  unsigned long long v = 0ull;
  if (i >= RS_FOOT && i < RS_FOOT + 3) v = enc(inf);
  if (i >= RS_FOOT + 3 && i < RS_FOOT + 6) v = enc(-inf);
  red[i] = v;
}

// dynamic shared memory of the main pass: the staged tile (+ the records of the optional solver compaction)
#ifdef XRTB_SOLVER_COMPACTION
constexpr size_t kSolverSmem = sizeof(FieldTile) + 10 * sizeof(double) * XRTB_REFLECT_THREADS;
#else
constexpr size_t kSolverSmem = sizeof(FieldTile);  // the staged tile of the block (fields + geometry: 130 B/ray)
#endif

// everything a main-pass launch needs besides the material switch; scr != NULL: fused Screen.expose (GlobalSink)
struct MainLaunch {
  unsigned grid;
  cudaStream_t st;
  size_t n;
  const xrt_oe_params* p;
  const xrt_material_params* m;
  const xrt_multilayer_params* ml;
  Tables t;
  xrt_beam in, fb, gb, lb;
  unsigned long long* red;
  const xrt_plane_params* scrp;  // fused Screen.expose: plane, output beam, the screen's own reduction slots
  const xrt_beam* scr;
  unsigned long long* sred;
};

#ifndef XRTB_PLUGIN
// every array the tile kernel loads or stores by bulk copies starts on a 16-byte boundary
static int tile_bulk_ok(const xrt_beam& b, bool withState) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  return al(b.x) && al(b.y) && al(b.z) && al(b.a) && al(b.b) && al(b.c) && al(b.path) && al(b.E) && al(b.Jss) &&
         al(b.Jpp) && al(b.Jsp) && al(b.Es) && al(b.Ep) && al(b.theta) && (!withState || al(b.state));
}
// XRTB_REFLECT_TILE: 0 = reflect_main_kernel always, 1 = the tile kernel wherever it is built, unset = by measurement
// (profiles/r02_reflect_variants.json): mirrors go through the tile kernel (rays that miss the surface leave idle lanes
// in thread <-> ray order), crystals through reflect_main_kernel (a DCM meets every ray: nothing to compact, and the
// second evaluation of the bracket in phase 1 costs more than the bulk stores save -- 25.4 vs 23.8 ms per 1e8 rays)
static int tile_mode() {
  const char* e = getenv("XRTB_REFLECT_TILE");
  if (!e || !e[0]) return -1;
  return e[0] == '0' ? 0 : 1;
}

// the tile structure of the main pass where it is built (see reflect_tile_kernel); false: not served, use the old kernel
template <int SURF, int MAT>
static bool launch_tile(const MainLaunch& L, double cR, double sR) {
  if constexpr ((SURF == 0 || SURF == 1 || SURF == 2) && (MAT == 0 || MAT == 1 || MAT == 3)) {
    const xrt_oe_params& p = *L.p;
    const int mode = tile_mode();
    if (p.multi || p.multi_tmin || p.no_intersection || p.grating || p.fe_c || mode == 0) return false;
    if (mode < 0 && MAT == 3) return false;
    int bulk_ok = tile_bulk_ok(L.in, true);
    if (L.scr == nullptr) bulk_ok = bulk_ok && tile_bulk_ok(L.gb, true);
    if (L.lb.x) bulk_ok = bulk_ok && tile_bulk_ok(L.lb, true);
    const unsigned grid = (unsigned)((L.n + kTileRays - 1) / kTileRays);
    xrt_plane_params sp;
    xrt_beam scr;
    memset(&sp, 0, sizeof(sp));
    memset(&scr, 0, sizeof(scr));
    if (L.scr) {
      sp = *L.scrp;
      scr = *L.scr;
      auto k = reflect_tile_kernel<SURF, MAT, true>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RayTile));
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      k<<<grid, kTileRays, sizeof(RayTile), L.st>>>(L.n, p, *L.m, L.t, L.in, L.fb, L.gb, L.lb, L.red, bulk_ok, cR, sR, sp,
                                                   scr, L.sred);
    } else {
      auto k = reflect_tile_kernel<SURF, MAT, false>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RayTile));
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      k<<<grid, kTileRays, sizeof(RayTile), L.st>>>(L.n, p, *L.m, L.t, L.in, L.fb, L.gb, L.lb, L.red, bulk_ok, cR, sR, sp,
                                                   scr, nullptr);
    }
    return true;
  } else {
    return false;
  }
}
#endif

template <int SURF, int MAT>
static void launch_one(const MainLaunch& L) {
  const int bulk_ok = bulk_copy_ok(L.in);
  const double cR = cos(L.p->roll), sR = sin(L.p->roll);  // polarisation frame: cos / sin(roll + atan2(n_x, n_z))
  reflect_decide_kernel<<<1, 1, 0, L.st>>>(L.red);
#ifndef XRTB_PLUGIN
  if (launch_tile<SURF, MAT>(L, cR, sR)) return;
#endif
  xrt_plane_params sp;
  xrt_beam scr;
  memset(&sp, 0, sizeof(sp));
  memset(&scr, 0, sizeof(scr));
  if (L.scr) {
    // the fused variant exists for the elements of ray-tracing chains: flat / toroid / elliptic mirrors and crystals
    if constexpr ((SURF == 0 || SURF == 1 || SURF == 2) && (MAT == 0 || MAT == 1 || MAT == 3)) {
      sp = *L.scrp;
      scr = *L.scr;
      cudaFuncSetAttribute(reflect_main_kernel<SURF, MAT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)kSolverSmem);
      reflect_main_kernel<SURF, MAT, true><<<L.grid, XRTB_REFLECT_THREADS, kSolverSmem, L.st>>>(
          L.n, *L.p, *L.m, *L.ml, L.t, L.in, L.fb, L.gb, L.lb, L.red, bulk_ok, cR, sR, sp, scr, L.sred);
    }
    return;
  }
  cudaFuncSetAttribute(reflect_main_kernel<SURF, MAT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)kSolverSmem);
  reflect_main_kernel<SURF, MAT, false><<<L.grid, XRTB_REFLECT_THREADS, kSolverSmem, L.st>>>(
      L.n, *L.p, *L.m, *L.ml, L.t, L.in, L.fb, L.gb, L.lb, L.red, bulk_ok, cR, sR, sp, scr, nullptr);
}

template <int SURF>
static int launch_main_mat(int mat, const MainLaunch& L) {
  if (L.scr && !((SURF == 0 || SURF == 1 || SURF == 2) && (mat == 0 || mat == 1 || mat == 3))) {
    set_last_error("reflect_expose: the fused reflect + expose kernel is built for flat, toroidal and elliptic "
                   "surfaces with material kinds 0 (none), 1 (mirror), 3 (crystal); got surface %d, kind %d",
                   SURF, mat);
    return 4;
  }
  switch (mat) {
    case 0: launch_one<SURF, 0>(L); break;
    case 1: launch_one<SURF, 1>(L); break;
    case 2: launch_one<SURF, 2>(L); break;
    case 3: launch_one<SURF, 3>(L); break;
    case 4: launch_one<SURF, 4>(L); break;
    case 6: launch_one<SURF, 6>(L); break;
    case 7: launch_one<SURF, 7>(L); break;
    case 5:
      if constexpr (SURF == 0 || SURF == 4) {
        launch_one<SURF, 5>(L);
      } else {
        set_last_error("reflect: refraction is built for flat and parabolic surfaces");
        return 2;
      }
      break;
    default: set_last_error("reflect: unknown material kind %d", mat); return 2;
  }
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// grid of a grid-stride pre-pass: two waves of the blocks that are resident at once (by the occupancy calculator: the
// register count differs between surface kinds), never more blocks than rays need
template <typename K>
static unsigned prepass_grid(K kernel, size_t n) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
  unsigned grid = (unsigned)num_sms() * (unsigned)per_sm * 2u;
  const size_t need = (n + 255) / 256;
  if (need < grid) grid = (unsigned)(need ? need : 1);
  return grid;
}

static int check_multilayer(const xrt_material_params& m, xrt_multilayer_params& ml) {
  memset(&ml, 0, sizeof(ml));
  if (m.kind != 4) return 0;
  XRTB_REQUIRE(m.multilayer != nullptr, "multilayer: material kind 4 needs the layer stack");
  ml = *m.multilayer;
  XRTB_REQUIRE(ml.nPairs >= 0 && ml.dti && ml.dbi, "multilayer: nPairs >= 0 and the thickness arrays are required");
  XRTB_REQUIRE(ml.tLayer.nelem >= 0 && ml.bLayer.nelem > 0 && ml.substrate.nelem > 0,
               "multilayer: bottom layer and substrate materials are required (no top layer = a coated mirror)");
  XRTB_REQUIRE(ml.tLayer.nelem <= XRT_MAX_ELEM && ml.bLayer.nelem <= XRT_MAX_ELEM && ml.substrate.nelem <= XRT_MAX_ELEM,
               "multilayer: at most %d elements per material", XRT_MAX_ELEM);
  return 0;
}

// refractive elements: plates (flat) and parabolic lenses, bare or of kind 'plate' / 'lens'
template <int SURF>
static int launch_main_refr(int mat, const MainLaunch& L) {
  if (L.scr) {
    set_last_error("reflect_expose: no fused variant for parabolic lens surfaces");
    return 4;
  }
  switch (mat) {
    case 0: launch_one<SURF, 0>(L); break;
    case 5: launch_one<SURF, 5>(L); break;
    default: set_last_error("reflect: a parabolic lens takes a material of kind 'lens' / 'plate' or none (got %d)", mat); return 2;
  }
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int reflect_expose_device(size_t n, const xrt_oe_params* oe, const xrt_material_params* mat, const double* tables,
                          size_t table_len, const xrt_beam* in, const xrt_beam* fallback, const xrt_beam* gb,
                          const xrt_beam* lb, unsigned long long* red, const xrt_plane_params* scr_plane,
                          const xrt_beam* scr, unsigned long long* scr_red, int phase,
                          unsigned long long index_offset, cudaStream_t st);

int reflect_device(size_t n, const xrt_oe_params* oe, const xrt_material_params* mat, const double* tables,
                   size_t table_len, const xrt_beam* in, const xrt_beam* fallback, const xrt_beam* gb,
                   const xrt_beam* lb, unsigned long long* red, int phase, unsigned long long index_offset,
                   cudaStream_t st) {
  return reflect_expose_device(n, oe, mat, tables, table_len, in, fallback, gb, lb, red, nullptr, nullptr, nullptr,
                               phase, index_offset, st);
}

// OE.reflect, optionally continued by Screen.expose in the same kernel (scr_plane / scr / scr_red given, gb NULL):
// the global beam never reaches HBM
int reflect_expose_device(size_t n, const xrt_oe_params* oe, const xrt_material_params* mat, const double* tables,
                          size_t table_len, const xrt_beam* in, const xrt_beam* fallback, const xrt_beam* gb,
                          const xrt_beam* lb, unsigned long long* red, const xrt_plane_params* scr_plane,
                          const xrt_beam* scr, unsigned long long* scr_red, int phase,
                          unsigned long long index_offset, cudaStream_t st) {
  XRTB_REQUIRE(oe && in && red, "reflect: NULL argument");
  XRTB_REQUIRE((gb != nullptr) != (scr != nullptr), "reflect: exactly one of gb (OE.reflect) and scr (fused "
                                                    "Screen.expose) receives the global beam");
  XRTB_REQUIRE(scr == nullptr || (scr_plane && scr_red), "reflect_expose: the screen needs its plane parameters and "
                                                         "reduction slots");
  XRTB_REQUIRE(scr == nullptr || !oe->multi, "reflect_expose: multiple_reflect passes are not fused");
#ifdef XRTB_PLUGIN
  XRTB_REQUIRE(oe->surface == XRTB_USER_KIND, "reflect (plug-in): this object serves surface kind %d only (got %d)",
               XRTB_USER_KIND, oe->surface);
#else
  XRTB_REQUIRE(oe->surface != 4 && oe->surface != 5,
               "reflect: surface kinds 4 / 5 (spliced cu_local_z / cu_local_n) live in plug-in objects built by "
               "xrt_b200.custom, not in libxrt_b200.so");
  XRTB_REQUIRE(oe->surface >= 0 && oe->surface <= 3, "reflect: unknown surface kind %d", oe->surface);
#endif
  XRTB_REQUIRE(oe->n_rot_in >= 0 && oe->n_rot_in <= XRT_MAX_ROT && oe->n_rot_out >= 0 && oe->n_rot_out <= XRT_MAX_ROT,
               "reflect: too many elementary rotations");
#ifndef XRTB_FIGURE_ERROR
  XRTB_REQUIRE(oe->fe_c == nullptr, "reflect: a figure error (fe_c) is served by the figure-error plug-in variant of the "
                                    "surface kind (xrt_b200.custom), not by this object");
#else
  XRTB_REQUIRE(oe->fe_c == nullptr || (oe->fe_tu && oe->fe_tv && oe->fe_nu >= 8 && oe->fe_nv >= 8),
               "reflect: figure error needs both knot vectors (>= 8 knots each) and the coefficients");
#endif
  xrt_material_params m;
  memset(&m, 0, sizeof(m));
  if (mat) m = *mat;
  XRTB_REQUIRE(m.kind >= 0 && m.kind <= 7, "reflect: unknown material kind %d", m.kind);
  XRTB_REQUIRE(oe->grating != 4 || (oe->fzp_rn != nullptr && oe->fzp_n >= 3 && m.kind == 7),
               "reflect: a zone plate (grating = 4) needs its zone radii and goes with material kind 7");
  XRTB_REQUIRE((m.kind == 6 || m.kind == 7) == (oe->grating != 0),
               "reflect: material kinds 6 / 7 (grating) and xrt_oe_params.grating go together");
  XRTB_REQUIRE(oe->surface != 3 || (m.kind == 0 || m.kind == 5), "reflect: a parabolic lens refracts or is bare");
  XRTB_REQUIRE(m.kind != 5 || oe->surface == 0 || oe->surface == 3 || oe->surface == 4,
               "reflect: refraction (kind plate / lens) is built for flat, parabolic and spliced z(x, y) surfaces");
  xrt_multilayer_params ml;
  if (int rc = check_multilayer(m, ml)) return rc;
  XRTB_REQUIRE(m.kind == 0 || m.kind == 7 || (tables != nullptr && table_len > 0), "reflect: material needs f1/f2 tables");
  XRTB_REQUIRE(m.nelem >= 0 && m.nelem <= XRT_MAX_ELEM, "reflect: at most %d elements per material", XRT_MAX_ELEM);
  Tables t{tables, tables ? tables + table_len : nullptr, tables ? tables + 2 * table_len : nullptr,
           reinterpret_cast<int*>(red + RS_HINT)};
  xrt_beam lbv;
  memset(&lbv, 0, sizeof(lbv));
  if (lb) lbv = *lb;
  const xrt_beam fbv = fallback ? *fallback : *in;
  unsigned pgrid = (unsigned)num_sms() * 8;
  {
    const size_t need = (n + 255) / 256;
    if (need < pgrid) pgrid = (unsigned)(need ? need : 1);
  }
  if (phase & 1) {  // phase A: dominant direction component, first good ray
    reduction_init_kernel<<<1, XRT_REDUCTION_SLOTS, 0, st>>>(red);
    if (n > 0) reflect_prepass_dir_kernel<<<pgrid, 256, 0, st>>>(n, *oe, *in, red, index_offset);
    XRTB_CUDA_CHECK(cudaGetLastError());
  }
  if ((phase & 4) && n > 0) {  // phase B: global bracket + solver switch of the selected variant
#ifdef XRTB_PLUGIN
    reflect_prepass_kernel<XRTB_USER_KIND><<<prepass_grid(reflect_prepass_kernel<XRTB_USER_KIND>, n), 256, 0, st>>>(n, *oe, *in, red);
#else
    if (oe->surface == 0) reflect_prepass_kernel<0><<<prepass_grid(reflect_prepass_kernel<0>, n), 256, 0, st>>>(n, *oe, *in, red);
    if (oe->surface == 1) reflect_prepass_kernel<1><<<prepass_grid(reflect_prepass_kernel<1>, n), 256, 0, st>>>(n, *oe, *in, red);
    if (oe->surface == 2) reflect_prepass_kernel<2><<<prepass_grid(reflect_prepass_kernel<2>, n), 256, 0, st>>>(n, *oe, *in, red);
    if (oe->surface == 3) reflect_prepass_kernel<3><<<prepass_grid(reflect_prepass_kernel<3>, n), 256, 0, st>>>(n, *oe, *in, red);
#endif
    XRTB_CUDA_CHECK(cudaGetLastError());
  }
  if ((phase & 2) && n > 0) {
    xrt_beam gbv;
    memset(&gbv, 0, sizeof(gbv));
    if (gb) gbv = *gb;
    MainLaunch L{(unsigned)((n + XRTB_REFLECT_THREADS - 1) / XRTB_REFLECT_THREADS), st, n, oe, &m, &ml, t, *in, fbv,
                 gbv, lbv, red, scr_plane, scr, scr_red};
    if (scr) screen_reduction_init_kernel<<<1, XRT_REDUCTION_SLOTS, 0, st>>>(scr_red);
    int rc = 0;
#ifdef XRTB_PLUGIN
    if constexpr (XRTB_USER_KIND == 3) {
      rc = launch_main_refr<XRTB_USER_KIND>(m.kind, L);  // the figure-error variant of the parabolic lens surface
    } else {
      rc = launch_main_mat<XRTB_USER_KIND>(m.kind, L);
    }
#else
    if (oe->surface == 0) rc = launch_main_mat<0>(m.kind, L);
    if (oe->surface == 1) rc = launch_main_mat<1>(m.kind, L);
    if (oe->surface == 2) rc = launch_main_mat<2>(m.kind, L);
    if (oe->surface == 3) rc = launch_main_refr<3>(m.kind, L);
#endif
    if (rc) return rc;
  }
  return 0;
}

// ---------------------------------------------------------------- stand-alone amplitudes (a12, a13)
__global__ void __launch_bounds__(256) amplitude_kernel(size_t n, xrt_material_params m, xrt_multilayer_params ml,
                                                        Tables tab,
                                                        const double* __restrict__ E, const double* __restrict__ bIn,
                                                        const double* __restrict__ bOut, const double* __restrict__ bH,
                                                        double2* __restrict__ rs, double2* __restrict__ rp,
                                                        unsigned long long* oor_count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int oor = 0;
  cplx s, p;
  if (m.kind == 3) {
    const double in = bIn[i];
    const double out = bOut ? bOut[i] : -in;   // crystal.py:618-619
    const double h = bH ? bH[i] : in;          // crystal.py:621-622
    crystal_amplitude(m, tab, E[i], in, out, h, s, p, &oor);
  } else if (m.kind == 4) {
    multilayer_amplitude(ml, tab, E[i], bIn[i], s, p, &oor);
  } else {
    mirror_amplitude(m, tab, E[i], bIn[i], s, p, &oor);
  }
  rs[i] = make_double2(s.re, s.im);
  rp[i] = make_double2(p.re, p.im);
  if (oor) atomicAdd(oor_count, 1ull);
}

int amplitude_device(size_t n, const xrt_material_params* mat, const double* tables, size_t table_len,
                     const double* E, const double* bIn, const double* bOut, const double* bH, double* rs, double* rp,
                     unsigned long long* oor_count, cudaStream_t st) {
  XRTB_REQUIRE(mat && tables && table_len > 0, "get_amplitude: material and tables are required");
  XRTB_REQUIRE(mat->kind >= 1 && mat->kind <= 4, "get_amplitude: unsupported material kind %d", mat->kind);
  xrt_multilayer_params ml;
  if (int rc = check_multilayer(*mat, ml)) return rc;
  if (n == 0) return 0;
  Tables t{tables, tables + table_len, tables + 2 * table_len, nullptr};
  amplitude_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, *mat, ml, t, E, bIn, bOut, bH,
                                                               reinterpret_cast<double2*>(rs),
                                                               reinterpret_cast<double2*>(rp), oor_count);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- surface normals at given local points
// OE.local_n (with xyz_to_param for parametric surfaces) as waves.diffract and OE.local_to_global use it
// (waves.py:675-686, oes/base.py:1213-1217): out = the surface normal n[-3:].
template <int SURF>
__global__ void __launch_bounds__(256) normals_kernel(size_t n, xrt_oe_params p, const double* __restrict__ x,
                                                      const double* __restrict__ y, const double* __restrict__ z,
                                                      double* __restrict__ nx, double* __restrict__ ny,
                                                      double* __restrict__ nz) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double px = x[i], py = y[i], pz = z[i];
  xyz_to_param<SURF>(p, px, py, pz);
  double nrm[6];
  local_normal<SURF>(p, make_surf_div<SURF>(p), px, py, nrm);
  nx[i] = nrm[3];
  ny[i] = nrm[4];
  nz[i] = nrm[5];
}

// OE.prepare_wave's surface height at mesh / random (x, y) samples (oes/reflect.py:335-340)
template <int SURF>
__global__ void __launch_bounds__(256) surface_kernel(size_t n, xrt_oe_params p, const double* __restrict__ x,
                                                      const double* __restrict__ y, double* __restrict__ z) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double px = x[i], py = y[i], pz = 0.0;
  if (SURF == 2 || SURF == 5) {
    xyz_to_param<SURF>(p, px, py, pz);
    pz = local_surface<SURF>(p, make_surf_div<SURF>(p), px, py);
    param_to_xyz<SURF>(p, px, py, pz);
  } else {
    pz = local_surface<SURF>(p, make_surf_div<SURF>(p), px, py);
  }
  z[i] = pz;
}

int surface_device(size_t n, const xrt_oe_params* oe, const double* x, const double* y, double* z, cudaStream_t st) {
  XRTB_REQUIRE(oe && x && y && z, "oe_surface: NULL argument");
#ifdef XRTB_PLUGIN
  XRTB_REQUIRE(oe->surface == XRTB_USER_KIND, "oe_surface (plug-in): surface kind %d expected", XRTB_USER_KIND);
  if (n == 0) return 0;
  surface_kernel<XRTB_USER_KIND><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, *oe, x, y, z);
#else
  XRTB_REQUIRE(oe->surface >= 0 && oe->surface <= 2, "oe_surface: unknown surface kind %d", oe->surface);
  if (n == 0) return 0;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (oe->surface == 0) surface_kernel<0><<<grid, 256, 0, st>>>(n, *oe, x, y, z);
  if (oe->surface == 1) surface_kernel<1><<<grid, 256, 0, st>>>(n, *oe, x, y, z);
  if (oe->surface == 2) surface_kernel<2><<<grid, 256, 0, st>>>(n, *oe, x, y, z);
#endif
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int normals_device(size_t n, const xrt_oe_params* oe, const double* x, const double* y, const double* z, double* nx,
                   double* ny, double* nz, cudaStream_t st) {
  XRTB_REQUIRE(oe && x && y && z && nx && ny && nz, "oe_normals: NULL argument");
#ifdef XRTB_PLUGIN
  XRTB_REQUIRE(oe->surface == XRTB_USER_KIND, "oe_normals (plug-in): surface kind %d expected", XRTB_USER_KIND);
  if (n == 0) return 0;
  normals_kernel<XRTB_USER_KIND><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, *oe, x, y, z, nx, ny, nz);
#else
  XRTB_REQUIRE(oe->surface >= 0 && oe->surface <= 2, "oe_normals: unknown surface kind %d", oe->surface);
  if (n == 0) return 0;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (oe->surface == 0) normals_kernel<0><<<grid, 256, 0, st>>>(n, *oe, x, y, z, nx, ny, nz);
  if (oe->surface == 1) normals_kernel<1><<<grid, 256, 0, st>>>(n, *oe, x, y, z, nx, ny, nz);
  if (oe->surface == 2) normals_kernel<2><<<grid, 256, 0, st>>>(n, *oe, x, y, z, nx, ny, nz);
#endif
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// multiple_reflect, derivative stage (see multi_deriv_kernel): phase 1 = beam-wide scalars, 2 = per-ray roots -> tD
int reflect_multi_deriv_device(size_t n, const xrt_oe_params* oe, const xrt_beam* in, unsigned long long* red,
                               double* tD, int phase, cudaStream_t st) {
  XRTB_REQUIRE(oe && in && red && tD, "reflect_multi_deriv: NULL argument");
  XRTB_REQUIRE(phase == 1 || phase == 2, "reflect_multi_deriv: phase is 1 (scalars) or 2 (roots)");
  if (n == 0) return 0;
  unsigned grid = (unsigned)num_sms() * 8;
  const size_t need = (n + 255) / 256;
  if (need < grid) grid = (unsigned)need;
#define XRTB_LAUNCH_DERIV(S)                                                             \
  do {                                                                                   \
    if (phase == 1)                                                                      \
      multi_deriv_kernel<S, 1><<<grid, 256, 0, st>>>(n, *oe, *in, red, tD);             \
    else                                                                                 \
      multi_deriv_kernel<S, 2><<<grid, 256, 0, st>>>(n, *oe, *in, red, tD);             \
  } while (0)
#ifdef XRTB_PLUGIN
  XRTB_REQUIRE(oe->surface == XRTB_USER_KIND, "reflect_multi_deriv (plug-in): surface kind %d expected", XRTB_USER_KIND);
  XRTB_LAUNCH_DERIV(XRTB_USER_KIND);
#else
  XRTB_REQUIRE(oe->surface >= 0 && oe->surface <= 2, "reflect_multi_deriv: surface kind %d has no multiple_reflect",
               oe->surface);
  if (oe->surface == 0) XRTB_LAUNCH_DERIV(0);
  if (oe->surface == 1) XRTB_LAUNCH_DERIV(1);
  if (oe->surface == 2) XRTB_LAUNCH_DERIV(2);
#endif
#undef XRTB_LAUNCH_DERIV
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// multiple_reflect(needElevationMap=True): elevation distance and point of every traced ray at its derivative root tD
int reflect_multi_elevation_device(size_t n, const xrt_oe_params* oe, const xrt_beam* in, const double* tD, double* eD,
                                   double* eX, double* eY, double* eZ, cudaStream_t st) {
  XRTB_REQUIRE(oe && in && tD && eD && eX && eY && eZ, "reflect_multi_elevation: NULL argument");
  if (n == 0) return 0;
  unsigned grid = (unsigned)num_sms() * 8;
  const size_t need = (n + 255) / 256;
  if (need < grid) grid = (unsigned)need;
#ifdef XRTB_PLUGIN
  XRTB_REQUIRE(oe->surface == XRTB_USER_KIND, "reflect_multi_elevation (plug-in): surface kind %d expected",
               XRTB_USER_KIND);
  multi_elevation_kernel<XRTB_USER_KIND><<<grid, 256, 0, st>>>(n, *oe, *in, tD, eD, eX, eY, eZ);
#else
  XRTB_REQUIRE(oe->surface >= 0 && oe->surface <= 2, "reflect_multi_elevation: surface kind %d has no multiple_reflect",
               oe->surface);
  if (oe->surface == 0) multi_elevation_kernel<0><<<grid, 256, 0, st>>>(n, *oe, *in, tD, eD, eX, eY, eZ);
  if (oe->surface == 1) multi_elevation_kernel<1><<<grid, 256, 0, st>>>(n, *oe, *in, tD, eD, eX, eY, eZ);
  if (oe->surface == 2) multi_elevation_kernel<2><<<grid, 256, 0, st>>>(n, *oe, *in, tD, eD, eX, eY, eZ);
#endif
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

double reduction_decode(unsigned long long slot) { return dec(slot); }

}  // namespace xrtb

#ifdef XRTB_PLUGIN
// C ABI of a spliced-surface plug-in object: the three entry points of include/xrt_b200.h that depend on the
// surface, same argument lists (xrt_reflect_f64, xrt_oe_surface_f64, xrt_oe_normals_f64), plus the struct size the
// object was compiled against.  Everything else in this translation unit has hidden visibility.
#define XRTB_EXPORT extern "C" __attribute__((visibility("default")))
XRTB_EXPORT int xrt_plugin_reflect_f64(size_t n, const xrt_oe_params* oe, const xrt_material_params* mat,
                                       const double* tables, size_t table_len, const xrt_beam* in,
                                       const xrt_beam* fallback, const xrt_beam* gb, const xrt_beam* lb,
                                       uint64_t* reductions, int phase, uint64_t index_offset, void* stream) {
  return xrtb::reflect_device(n, oe, mat, tables, table_len, in, fallback, gb, lb,
                              reinterpret_cast<unsigned long long*>(reductions), phase, index_offset,
                              (cudaStream_t)stream);
}
XRTB_EXPORT int xrt_plugin_oe_surface_f64(size_t n, const xrt_oe_params* oe, const double* x, const double* y,
                                          double* z, void* stream) {
  return xrtb::surface_device(n, oe, x, y, z, (cudaStream_t)stream);
}
XRTB_EXPORT int xrt_plugin_oe_normals_f64(size_t n, const xrt_oe_params* oe, const double* x, const double* y,
                                          const double* z, double* nx, double* ny, double* nz, void* stream) {
  return xrtb::normals_device(n, oe, x, y, z, nx, ny, nz, (cudaStream_t)stream);
}
XRTB_EXPORT int xrt_plugin_reflect_multi_deriv_f64(size_t n, const xrt_oe_params* oe, const xrt_beam* in,
                                                   uint64_t* reductions, double* tD, int phase, void* stream) {
  return xrtb::reflect_multi_deriv_device(n, oe, in, reinterpret_cast<unsigned long long*>(reductions), tD, phase,
                                          (cudaStream_t)stream);
}
XRTB_EXPORT int xrt_plugin_reflect_multi_elevation_f64(size_t n, const xrt_oe_params* oe, const xrt_beam* in,
                                                       const double* tD, double* eD, double* eX, double* eY,
                                                       double* eZ, void* stream) {
  return xrtb::reflect_multi_elevation_device(n, oe, in, tD, eD, eX, eY, eZ, (cudaStream_t)stream);
}
XRTB_EXPORT size_t xrt_plugin_sizeof_oe_params(void) { return sizeof(xrt_oe_params); }
XRTB_EXPORT int xrt_plugin_surface_kind(void) { return XRTB_USER_KIND; }
XRTB_EXPORT int xrt_plugin_has_own_intersection(void) {
#ifdef XRTB_USER_INTERSECTION
  return 1;
#else
  return 0;
#endif
}
XRTB_EXPORT int xrt_plugin_has_figure_error(void) {
#ifdef XRTB_FIGURE_ERROR
  return 1;
#else
  return 0;
#endif
}
#endif
