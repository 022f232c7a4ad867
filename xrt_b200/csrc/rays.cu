// Plane elements of the ray path for sm_100a: Screen.expose and RectangularAperture.propagate.
//
// Replaces Screen.expose (xrt/backends/raycing/screens.py:225-302) and RectangularAperture.propagate
// (xrt/backends/raycing/apertures.py:333-413) with their frame changes (beamline.py:230-264).
// Oracle: oracle/planes.py.  HBM-bound elementwise kernels: 132 B read + 132 B written per ray (with
// amplitudes), one thread per ray, unit-stride SoA access.
//
// Compiled with -fmad=false (numpy rounding order, see reflect.cu); fma() appears only where numpy itself
// fuses (its SIMD complex multiply).
#include "../../include/xrt_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace xrtb {

enum { RS_FOOT = 28, RS_NGOOD = 34 };

__host__ __device__ inline unsigned long long enc2(double d) {
  unsigned long long u;
  memcpy(&u, &d, 8);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

__device__ __forceinline__ cplx cmul_np2(cplx a, cplx b) {  // numpy SIMD complex multiply (see reflect.cu)
  const double ii = a.im * b.im, ir = a.im * b.re;
  return cplx{fma(a.re, b.re, -ii), fma(a.re, b.im, ir)};
}
__device__ __forceinline__ cplx ldc(const double* p, size_t i) {
  const double2 v = reinterpret_cast<const double2*>(p)[i];
  return cplx{v.x, v.y};
}
__device__ __forceinline__ void stc(double* p, size_t i, cplx v) {
  reinterpret_cast<double2*>(p)[i] = make_double2(v.re, v.im);
}

__global__ void plane_reduction_init_kernel(unsigned long long* red) {
  const int i = threadIdx.x;
  if (i >= XRT_REDUCTION_SLOTS) return;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  unsigned long long v = 0ull;
  if (i >= RS_FOOT && i < RS_FOOT + 3) v = enc2(inf);
  if (i >= RS_FOOT + 3 && i < RS_FOOT + 6) v = enc2(-inf);
  red[i] = v;
}

__device__ __forceinline__ void block_minmax(double* mn, double* mx, unsigned long long cnt,
                                             unsigned long long* red) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  __shared__ unsigned long long sh[8][7];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long e[7];
  for (int k = 0; k < 3; ++k) {
    e[k] = enc2(mn[k]);
    e[3 + k] = enc2(mx[k]);
  }
  e[6] = cnt;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    unsigned long long x = e[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, o);
      x = (k == 6) ? x + y : (k < 3 ? (y < x ? y : x) : (y > x ? y : x));
    }
    if (lane == 0) sh[warp][k] = x;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    const int k = threadIdx.x;
    unsigned long long x = sh[0][k];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      const unsigned long long y = sh[w][k];
      x = (k == 6) ? x + y : (k < 3 ? (y < x ? y : x) : (y > x ? y : x));
    }
    if (k < 3) {
      if (x != enc2(inf)) atomicMin(&red[RS_FOOT + k], x);
    } else if (k < 6) {
      if (x != enc2(-inf)) atomicMax(&red[RS_FOOT + k], x);
    } else if (x) {
      atomicAdd(&red[RS_NGOOD], x);
    }
  }
}

// ---------------------------------------------------------------- Screen.expose
__global__ void __launch_bounds__(256) expose_kernel(size_t n, xrt_plane_params p, xrt_beam in, xrt_beam out,
                                                     unsigned long long* red) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
  unsigned long long cnt = 0;
  if (i < n) {
    // global_to_virgin_local with an explicit basis (beamline.py:251-264): every ray, whatever its state
    const double x0 = in.x[i] - p.center[0], y0 = in.y[i] - p.center[1], z0 = in.z[i] - p.center[2];
    const double a0 = in.a[i], b0 = in.b[i], c0 = in.c[i];
    double x = p.ex[0] * x0 + p.ex[1] * y0 + p.ex[2] * z0;
    const double y = p.ey[0] * x0 + p.ey[1] * y0 + p.ey[2] * z0;
    double z = p.ez[0] * x0 + p.ez[1] * y0 + p.ez[2] * z0;
    const double a = p.ex[0] * a0 + p.ex[1] * b0 + p.ex[2] * c0;
    const double b = p.ey[0] * a0 + p.ey[1] * b0 + p.ey[2] * c0;
    const double c = p.ez[0] * a0 + p.ez[1] * b0 + p.ez[2] * c0;
    // (the unchecked IEEE sequence, common.cuh: the same quotient for ordinary operands; b == 0 or an overflowing
    // quotient gives NaN / inf either way and the ray is lost below -- GlobalSink in reflect.cu does the same)
    double path = div_nr(-y, b);
    int st = in.state[i];
    bool bad = isnan(path) || isinf(path);
    if (p.only_positive_path) bad = bad || (path < 0.0);
    if (bad) {
      path = 0.0;
      st = p.lostNum;
    }
    const double E = in.E[i];
    x = x + a * path;
    z = z + c * path;
    if (st == 1 || st == 2) {  // footprint (screens.py:262-268), before compression
      mn[0] = mx[0] = x;
      mn[1] = mx[1] = 0.0;
      mn[2] = mx[2] = z;
      cnt = 1;
    }
    if (p.has_compressX) x *= p.compressX;
    if (p.has_compressZ) z *= p.compressZ;
    out.x[i] = x;
    out.y[i] = 0.0;
    out.z[i] = z;
    out.a[i] = a;
    out.b[i] = b;
    out.c[i] = c;
    out.path[i] = in.path[i] + path;
    out.E[i] = E;
    out.Jss[i] = in.Jss[i];
    out.Jpp[i] = in.Jpp[i];
    stc(out.Jsp, i, ldc(in.Jsp, i));
    out.state[i] = st;
    if (in.Es && out.Es) {
      // propPhase = exp(1e7j * (E/CHBAR) * path) (screens.py:257-260)
      const double CHBAR = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8 / 6.283185307179586476925286766559;
      const double phi = (1e7 * div_nr(E, CHBAR)) * path;
      double s, cs;
      sincos_big(phi, s, cs);
      const cplx ph = cplx{cs, s};
      stc(out.Es, i, cmul_np2(ldc(in.Es, i), ph));
      stc(out.Ep, i, cmul_np2(ldc(in.Ep, i), ph));
    }
  }
  block_minmax(mn, mx, cnt, red);
}

// ---------------------------------------------------------------- HemisphericScreen.expose (screens.py:501-545)
// Intersection with the sphere of radius R around the centre, in global coordinates; positions go to the screen's
// local frame, directions stay global (the reference copies them untouched); theta = asin(z/R) - thetaOffset,
// phi = atan2(y, x) - phiOffset land in out.theta / out.phi.
__global__ void __launch_bounds__(256) expose_hemi_kernel(size_t n, xrt_plane_params p, double R, double R2,
                                                          double thetaOffset, double phiOffset, xrt_beam in,
                                                          xrt_beam out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double bx = in.x[i], by = in.y[i], bz = in.z[i], a = in.a[i], b = in.b[i], c = in.c[i];
  const double dx = bx - p.center[0], dy = by - p.center[1], dz = bz - p.center[2];
  const double sqb_2 = a * dx + b * dy + c * dz;
  const double sqc = dx * dx + dy * dy + dz * dz - R2;
  double path = -sqb_2 + sqrt(sqb_2 * sqb_2 - sqc);
  int st = in.state[i];
  bool bad = isnan(path) || isinf(path);
  if (p.only_positive_path) bad = bad || (path < 0.0);
  if (bad) {
    path = 0.0;
    st = p.lostNum;
  }
  const double rx = bx + a * path - p.center[0], ry = by + b * path - p.center[1], rz = bz + c * path - p.center[2];
  const double lz = rx * p.ez[0] + ry * p.ez[1] + rz * p.ez[2];
  const double ly = rx * p.ey[0] + ry * p.ey[1] + rz * p.ey[2];
  const double lx = rx * p.ex[0] + ry * p.ex[1] + rz * p.ex[2];
  out.x[i] = lx;
  out.y[i] = ly;
  out.z[i] = lz;
  out.a[i] = a;
  out.b[i] = b;
  out.c[i] = c;
  out.path[i] = in.path[i] + path;
  const double E = in.E[i];
  out.E[i] = E;
  out.Jss[i] = in.Jss[i];
  out.Jpp[i] = in.Jpp[i];
  stc(out.Jsp, i, ldc(in.Jsp, i));
  out.state[i] = st;
  if (out.theta) out.theta[i] = asin(lz / R) - thetaOffset;
  if (out.phi) out.phi[i] = atan2(ly, lx) - phiOffset;
  if (in.Es && out.Es) {
    const double CHBAR = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8 / 6.283185307179586476925286766559;
    const double phi = (1e7 * (E / CHBAR)) * path;
    double s, cs;
    sincos_big(phi, s, cs);
    const cplx ph = cplx{cs, s};
    stc(out.Es, i, cmul_np2(ldc(in.Es, i), ph));
    stc(out.Ep, i, cmul_np2(ldc(in.Ep, i), ph));
  }
}

int expose_hemispheric_device(size_t n, const xrt_plane_params* p, double R, double thetaOffset, double phiOffset,
                              const xrt_beam* in, const xrt_beam* out, cudaStream_t st) {
  XRTB_REQUIRE(p && in && out, "expose_hemispheric: NULL argument");
  XRTB_REQUIRE(R > 0.0, "expose_hemispheric: the radius must be positive");
  if (n > 0)
    expose_hemi_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, *p, R, R * R, thetaOffset, phiOffset, *in,
                                                                    *out);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- RectangularAperture.propagate
__global__ void __launch_bounds__(256) propagate_rect_kernel(size_t n, xrt_plane_params p, xrt_beam in, xrt_beam out,
                                                             unsigned long long* red) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
  unsigned long long cnt = 0;
  if (i < n) {
    int st = in.state[i];
    double x = in.x[i], y = in.y[i], z = in.z[i], a = in.a[i], b = in.b[i], c = in.c[i];
    double path = in.path[i];
    const double E = in.E[i];
    cplx Es = cplx{0, 0}, Ep = cplx{0, 0};
    const bool amp = in.Es && out.Es;
    if (amp) {
      Es = ldc(in.Es, i);
      Ep = ldc(in.Ep, i);
    }
    if (st > 0) {
      // global_to_virgin_local (beamline.py:239-264): the aperture's `xyz` is always an explicit basis
      // (apertures.py:343 with :277-283); the azimuth form is kept for completeness
      x -= p.center[0];
      y -= p.center[1];
      z -= p.center[2];
      if (p.basis_mode) {
        const double x0 = x, y0 = y, z0 = z, a0 = a, b0 = b, c0 = c;
        x = p.ex[0] * x0 + p.ex[1] * y0 + p.ex[2] * z0;
        y = p.ey[0] * x0 + p.ey[1] * y0 + p.ey[2] * z0;
        z = p.ez[0] * x0 + p.ez[1] * y0 + p.ez[2] * z0;
        a = p.ex[0] * a0 + p.ex[1] * b0 + p.ex[2] * c0;
        b = p.ey[0] * a0 + p.ey[1] * b0 + p.ey[2] * c0;
        c = p.ez[0] * a0 + p.ez[1] * b0 + p.ez[2] * c0;
      } else if (p.sinAz != 0.0) {
        const double x1 = p.cosAz * x - p.sinAz * y, y1 = p.sinAz * x + p.cosAz * y;
        x = x1;
        y = y1;
        const double a1 = p.cosAz * a - p.sinAz * b, b1 = p.sinAz * a + p.cosAz * b;
        a = a1;
        b = b1;
      }
      const double dpath = -y / b;
      x += a * dpath;
      z += c * dpath;
      path += dpath;
      bool badI = false;
      for (int k = 0; k < p.n_blades; ++k) {  // apertures.py:352-361
        const double d = p.blade_pos[k];
        const int kind = p.blade_kind[k];
        if (kind == 0) badI = badI || (x < d);
        if (kind == 1) badI = badI || (x > d);
        if (kind == 2) badI = badI || (z < d);
        if (kind == 3) badI = badI || (z > d);
        if (kind == 4) badI = badI || (sqrt(x * x + z * z) > d);  // RoundAperture: lo.r > self.r (apertures.py:790-794)
      }
      if (p.is_beam_stop) badI = !badI;
      if (badI) st = p.lostNum;
      y = 0.0;
      if (amp) {
        const double CHBAR = 6.626069573e-27 * 2.99792458e10 / 1.602176565e-12 * 1e8 / 6.283185307179586476925286766559;
        const double phi = (1e7 * (E / CHBAR)) * dpath;
        double s, cs;
        sincos_big(phi, s, cs);
        const cplx ph = cplx{cs, s};
        Es = cmul_np2(Es, ph);
        Ep = cmul_np2(Ep, ph);
      }
      if (st > 0) {  // spotLimits (apertures.py:377-383)
        mn[0] = mx[0] = x;
        mn[1] = mx[1] = 0.0;
        mn[2] = mx[2] = z;
        cnt = 1;
      }
      in.state[i] = st;  // beam.state[badIndices] = self.lostNum: the INPUT beam is marked too (apertures.py:364)
    }
    out.x[i] = x;
    out.y[i] = y;
    out.z[i] = z;
    out.a[i] = a;
    out.b[i] = b;
    out.c[i] = c;
    out.path[i] = path;
    out.E[i] = E;
    out.Jss[i] = in.Jss[i];
    out.Jpp[i] = in.Jpp[i];
    stc(out.Jsp, i, ldc(in.Jsp, i));
    out.state[i] = st;
    if (amp) {
      stc(out.Es, i, Es);
      stc(out.Ep, i, Ep);
    }
  }
  block_minmax(mn, mx, cnt, red);
}

int expose_device(size_t n, const xrt_plane_params* p, const xrt_beam* in, const xrt_beam* out,
                  unsigned long long* red, cudaStream_t st) {
  XRTB_REQUIRE(p && in && out && red, "expose: NULL argument");
  plane_reduction_init_kernel<<<1, XRT_REDUCTION_SLOTS, 0, st>>>(red);
  if (n > 0) expose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, *p, *in, *out, red);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int propagate_rect_device(size_t n, const xrt_plane_params* p, const xrt_beam* in, const xrt_beam* out,
                          unsigned long long* red, cudaStream_t st) {
  XRTB_REQUIRE(p && in && out && red, "propagate: NULL argument");
  XRTB_REQUIRE(p->n_blades >= 0 && p->n_blades <= 4, "propagate: a rectangular aperture has at most 4 blades");
  plane_reduction_init_kernel<<<1, XRT_REDUCTION_SLOTS, 0, st>>>(red);
  if (n > 0) propagate_rect_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, *p, *in, *out, red);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- compaction of selected rays and fields
// What a consumer of a repeat reads from a Beam is a handful of fields of the rays raycing.get_output selects
// (raycing/__init__.py:170-300: `good = (state == 1) | ...`; x, z, E, Jss + Jpp -- 40 B of the 132 B/ray, and 59 % of
// the rays in BASELINE config 2).  Three small kernels pack those fields of the selected rays densely, in ray order,
// so that only they cross PCIe: per-block counts, an exclusive scan of the counts (one block), scatter with an
// in-block ballot scan.  Up to 8 arrays of 8-byte words per call (a complex128 field is two words per ray).
constexpr int kCompactBlock = 1024;
__device__ __forceinline__ bool ray_selected(int st, int state_mask) {
  const int bit = st == 1 ? 1 : (st == 2 ? 2 : (st == 3 ? 4 : (st < 0 ? 8 : (st > 0 ? 16 : 32))));
  return (bit & state_mask) != 0;
}
__global__ void __launch_bounds__(kCompactBlock) compact_count_kernel(size_t n, const int* __restrict__ state,
                                                                       int state_mask,
                                                                       unsigned long long* __restrict__ counts) {
  const size_t i = (size_t)blockIdx.x * kCompactBlock + threadIdx.x;
  const bool sel = i < n && ray_selected(state[i], state_mask);
  const int c = __syncthreads_count(sel);
  if (threadIdx.x == 0) counts[blockIdx.x] = (unsigned long long)c;
}
// counts[0 .. nb) -> exclusive prefix sums in place, counts[nb] = total
__global__ void __launch_bounds__(1024) compact_scan_kernel(size_t nb, unsigned long long* __restrict__ counts) {
  __shared__ unsigned long long sh[1024];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (size_t base = 0; base < nb; base += 1024) {
    const size_t k = base + threadIdx.x;
    const unsigned long long v = k < nb ? counts[k] : 0ull;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
      const unsigned long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0ull;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (k < nb) counts[k] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[nb] = carry;
}
struct CompactArrays {
  const double* src[8];
  double* dst[8];
  int words[8];
  int narr;
};
__global__ void __launch_bounds__(kCompactBlock) compact_scatter_kernel(size_t n, const int* __restrict__ state,
                                                                         int state_mask,
                                                                         const unsigned long long* __restrict__ offs,
                                                                         CompactArrays a) {
  __shared__ int wcount[kCompactBlock / 32];
  const size_t i = (size_t)blockIdx.x * kCompactBlock + threadIdx.x;
  const bool sel = i < n && ray_selected(state[i], state_mask);
  const unsigned b = __ballot_sync(0xffffffffu, sel);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) wcount[warp] = __popc(b);
  __syncthreads();
  if (!sel) return;
  int before = 0;
  for (int w = 0; w < warp; ++w) before += wcount[w];
  const size_t pos = (size_t)offs[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u));
  for (int k = 0; k < a.narr; ++k) {
    if (a.words[k] == 2)
      reinterpret_cast<double2*>(a.dst[k])[pos] = reinterpret_cast<const double2*>(a.src[k])[i];
    else
      a.dst[k][pos] = a.src[k][i];
  }
}

size_t compact_scratch_words(size_t n) { return (n + kCompactBlock - 1) / kCompactBlock + 1; }

int compact_rays_device(size_t n, const int* state, int state_mask, int narr, const double* const* src,
                        double* const* dst, const int* words, unsigned long long* scratch, cudaStream_t st) {
  XRTB_REQUIRE(state && src && dst && words && scratch, "compact_rays: NULL argument");
  XRTB_REQUIRE(narr >= 1 && narr <= 8, "compact_rays: 1..8 arrays per call (got %d)", narr);
  CompactArrays a;
  a.narr = narr;
  for (int k = 0; k < narr; ++k) {
    XRTB_REQUIRE(src[k] && dst[k] && (words[k] == 1 || words[k] == 2), "compact_rays: array %d", k);
    a.src[k] = src[k];
    a.dst[k] = dst[k];
    a.words[k] = words[k];
  }
  const size_t nb = (n + kCompactBlock - 1) / kCompactBlock;
  if (nb == 0) {
    XRTB_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(unsigned long long), st));
    return 0;
  }
  compact_count_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(n, state, state_mask, scratch);
  compact_scan_kernel<<<1, 1024, 0, st>>>(nb, scratch);
  compact_scatter_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(n, state, state_mask, scratch, a);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- f1: weighted 2-D histogram on the device
// np.histogram2d(y, x, bins=[ny, nx], range=..., weights=w) as xrt's do_hist2d calls it (xrt/multipro.py:119-171)
// on the rays raycing.get_output selects (state filter) with intensity = Jss + Jpp.  Bin edges come from the host
// (np.linspace, so they are numpy's edges bit for bit); a ray lands in bin b with e[b] <= v < e[b+1], the last
// edge inclusive (np.histogramdd: searchsorted(side='right') and the on_edge fix-up).
__device__ __forceinline__ int find_bin(const double* __restrict__ e, int nb, double v, double inv_w) {
  if (!(v >= e[0]) || !(v <= e[nb])) return -1;  // outliers and NaN
  int g = (int)((v - e[0]) * inv_w);
  g = g < 0 ? 0 : (g > nb - 1 ? nb - 1 : g);
  while (g > 0 && v < e[g]) --g;
  while (g < nb - 1 && v >= e[g + 1]) ++g;
  return g;
}

__global__ void __launch_bounds__(256) hist2d_kernel(size_t n, const double* __restrict__ x,
                                                     const double* __restrict__ y, const double* __restrict__ w1,
                                                     const double* __restrict__ w2, const int* __restrict__ state,
                                                     int state_mask, const double* __restrict__ ex, int nx,
                                                     const double* __restrict__ ey, int ny, double* __restrict__ out) {
  extern __shared__ double sedges[];  // ex[nx + 1] | ey[ny + 1]
  double* sx = sedges;
  double* sy = sedges + nx + 1;
  for (int i = threadIdx.x; i <= nx; i += blockDim.x) sx[i] = ex[i];
  for (int i = threadIdx.x; i <= ny; i += blockDim.x) sy[i] = ey[i];
  __syncthreads();
  const double inv_wx = nx / (sx[nx] - sx[0]), inv_wy = ny / (sy[ny] - sy[0]);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (state) {
      const int st = state[i];
      const int bit = st == 1 ? 1 : (st == 2 ? 2 : (st == 3 ? 4 : (st < 0 ? 8 : (st > 0 ? 16 : 32))));
      if (!(bit & state_mask)) continue;
    }
    const int bx = find_bin(sx, nx, x[i], inv_wx);
    if (bx < 0) continue;
    const int by = find_bin(sy, ny, y[i], inv_wy);
    if (by < 0) continue;
    double w = w1 ? w1[i] : 1.0;
    if (w2) w = w + w2[i];
    atomicAdd(&out[(size_t)by * nx + bx], w);
  }
}

int hist2d_device(size_t n, const double* x, const double* y, const double* w1, const double* w2, const int* state,
                  int state_mask, const double* ex, int nx, const double* ey, int ny, double* out, cudaStream_t st) {
  XRTB_REQUIRE(x && y && ex && ey && out, "hist2d: NULL argument");
  XRTB_REQUIRE(nx > 0 && ny > 0 && nx + ny + 2 <= 6000, "hist2d: 1 <= bins and nx + ny <= 5998");
  if (n == 0) return 0;
  unsigned grid = (unsigned)num_sms() * 8;
  const size_t need = (n + 255) / 256;
  if (need < grid) grid = (unsigned)need;
  const size_t smem = (size_t)(nx + ny + 2) * sizeof(double);
  hist2d_kernel<<<grid, 256, smem, st>>>(n, x, y, w1, w2, state, state_mask, ex, nx, ey, ny, out);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- f1: one XYCPlot in one pass
// flux by fluxKind (raycing/__init__.py:267-296), hue -> RGB (matplotlib.colors.hsv_to_rgb), 1-D histograms of
// the x, y and colour axes and the 2-D image, each with 4 channels (flux, R, G, B); xrt/multipro.py:296-340.
__device__ __forceinline__ void hsv_to_rgb(double h, double s, double v, double& r, double& g, double& b) {
  if (s == 0.0) {
    r = g = b = v;
    return;
  }
  const double h6 = h * 6.0;
  const int i = (int)h6;
  const double f = h6 - (double)i;
  const double p = v * (1.0 - s);
  const double q = v * (1.0 - s * f);
  const double t = v * (1.0 - s * (1.0 - f));
  switch (i % 6) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

__device__ __forceinline__ void add4(double* __restrict__ h, size_t bin, const double* w) {
#pragma unroll
  for (int k = 0; k < 4; ++k) atomicAdd(&h[bin * 4 + k], w[k]);
}

// PRIV: the three 1-D histograms (a few hundred hot bins hit by every ray) are privatised per block in shared
// memory and flushed once at the end; the 2-D image goes to global memory with RED.ADD.F64.
template <bool PRIV>
__global__ void __launch_bounds__(1024) plot_hist_kernel(size_t n, xrt_plot_params p) {
  extern __shared__ double sedges[];  // ex[nx + 1] | ey[ny + 1] | ec[nc + 1] | PRIV: hx[nx*4] | hy[ny*4] | hc[nc*4]
  double* sx = sedges;
  double* sy = sx + p.nx + 1;
  double* sc = sy + p.ny + 1;
  double* hx = sc + p.nc + 1;
  double* hy = hx + 4 * p.nx;
  double* hc = hy + 4 * p.ny;
  const bool hasC = p.c && p.ec && p.nc > 0;
  for (int i = threadIdx.x; i <= p.nx; i += blockDim.x) sx[i] = p.ex[i];
  for (int i = threadIdx.x; i <= p.ny; i += blockDim.x) sy[i] = p.ey[i];
  if (hasC)
    for (int i = threadIdx.x; i <= p.nc; i += blockDim.x) sc[i] = p.ec[i];
  if (PRIV)
    for (int i = threadIdx.x; i < 4 * (p.nx + p.ny + p.nc); i += blockDim.x) hx[i] = 0.0;
  __shared__ unsigned long long scnt[6];
  if (threadIdx.x < 6) scnt[threadIdx.x] = 0;
  __syncthreads();
  const double inv_wx = p.nx / (sx[p.nx] - sx[0]), inv_wy = p.ny / (sy[p.ny] - sy[0]);
  const double inv_wc = hasC ? p.nc / (sc[p.nc] - sc[0]) : 0.0;
  const double SIE0 = 1.602176565e-19;
  unsigned cnt[6] = {0, 0, 0, 0, 0, 0};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int st = p.state ? p.state[i] : 1;
    const int bit = st == 1 ? 1 : (st == 2 ? 2 : (st == 3 ? 4 : (st < 0 ? 8 : (st > 0 ? 16 : 32))));
    cnt[1] += st > 0;
    bool sel = (bit & p.state_mask) != 0;
    for (int k = 0; k < p.n_dead_flags; ++k) sel = sel || (st == p.dead_flags[k]);
    if (!sel) continue;
    cnt[0] += 1;
    cnt[2] += st == 1;
    cnt[3] += st == 2;
    cnt[4] += st == 3;
    cnt[5] += st < 0;
    double flux;
    switch (p.flux_kind) {
      case 1: flux = p.Jss[i] * p.srcWt; break;
      case 2: flux = p.Jpp[i] * p.srcWt; break;
      case 3: flux = 2.0 * p.Jsp[2 * i] * p.srcWt; break;
      case 4: flux = 2.0 * p.Jsp[2 * i + 1] * p.srcWt; break;
      case 5: flux = (p.Jss[i] + p.Jpp[i]) * p.E[i] * SIE0 * p.srcWt; break;
      default: flux = (p.Jss[i] + p.Jpp[i]) * p.srcWt; break;
    }
    double w[4];
    w[0] = flux;
    double cv = 0.0;
    if (p.c) {
      cv = p.c[i] * p.cfactor;
      double c01 = (cv - p.climits[0]) * p.colorFactor / (p.climits[1] - p.climits[0]);
      if (c01 < 0.0) c01 = 0.0;
      if (c01 > 1.0) c01 = 1.0;
      hsv_to_rgb(c01, p.colorSaturation, flux, w[1], w[2], w[3]);
    } else {
      w[1] = w[2] = w[3] = flux;
    }
    const int bx = find_bin(sx, p.nx, p.x[i] * p.xfactor, inv_wx);
    const int by = find_bin(sy, p.ny, p.y[i] * p.yfactor, inv_wy);
    if (bx >= 0 && p.histx) add4(PRIV ? hx : p.histx, (size_t)bx, w);
    if (by >= 0 && p.histy) add4(PRIV ? hy : p.histy, (size_t)by, w);
    if (bx >= 0 && by >= 0 && p.hist2d) add4(p.hist2d, (size_t)by * p.nx + bx, w);
    if (hasC && p.histc) {
      const int bc = find_bin(sc, p.nc, cv, inv_wc);
      if (bc >= 0) add4(PRIV ? hc : p.histc, (size_t)bc, w);
    }
  }
  if (p.counters) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      unsigned v = cnt[k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v) atomicAdd(&scnt[k], (unsigned long long)v);
    }
  }
  __syncthreads();
  if (p.counters && threadIdx.x < 6 && scnt[threadIdx.x]) atomicAdd(&p.counters[threadIdx.x], scnt[threadIdx.x]);
  if (PRIV) {
    for (int i = threadIdx.x; i < 4 * p.nx; i += blockDim.x)
      if (p.histx && hx[i] != 0.0) atomicAdd(&p.histx[i], hx[i]);
    for (int i = threadIdx.x; i < 4 * p.ny; i += blockDim.x)
      if (p.histy && hy[i] != 0.0) atomicAdd(&p.histy[i], hy[i]);
    for (int i = threadIdx.x; i < 4 * p.nc; i += blockDim.x)
      if (p.histc && hasC && hc[i] != 0.0) atomicAdd(&p.histc[i], hc[i]);
  }
}

int plot_hist_device(size_t n, const xrt_plot_params* p, cudaStream_t st) {
  XRTB_REQUIRE(p && p->x && p->y && p->ex && p->ey, "plot_hist: NULL axis data or edges");
  XRTB_REQUIRE(p->nx > 0 && p->ny > 0 && p->nc >= 0 && p->nx + p->ny + p->nc + 3 <= 6000,
               "plot_hist: 1 <= bins and nx + ny + nc <= 5997");
  XRTB_REQUIRE(p->flux_kind >= 0 && p->flux_kind <= 5, "plot_hist: unknown flux kind %d", p->flux_kind);
  XRTB_REQUIRE(p->n_dead_flags >= 0 && p->n_dead_flags <= 4, "plot_hist: at most 4 negative ray flags");
  XRTB_REQUIRE(p->Jss && p->Jpp && (p->flux_kind < 3 || p->flux_kind == 5 || p->Jsp) && (p->flux_kind != 5 || p->E),
               "plot_hist: the beam arrays the flux kind needs are missing");
  if (n == 0) return 0;
  unsigned grid = (unsigned)num_sms() * 8;
  const size_t need = (n + 255) / 256;
  if (need < grid) grid = (unsigned)need;
  const size_t nb = (size_t)p->nx + p->ny + p->nc;
  const size_t smem_priv = (nb + 3 + 4 * nb) * sizeof(double);
  if (smem_priv <= 46 * 1024) {
    plot_hist_kernel<true><<<grid, 256, smem_priv, st>>>(n, *p);
  } else if (smem_priv <= 200 * 1024) {
    // many bins: one 1024-thread block per SM with the opt-in shared-memory carve-out
    static bool attr_set = false;
    if (!attr_set) {
      XRTB_CUDA_CHECK(cudaFuncSetAttribute(plot_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           200 * 1024));
      attr_set = true;
    }
    unsigned g1 = (unsigned)num_sms();
    const size_t need1 = (n + 1023) / 1024;
    if (need1 < g1) g1 = (unsigned)need1;
    plot_hist_kernel<true><<<g1, 1024, smem_priv, st>>>(n, *p);
  } else {
    plot_hist_kernel<false><<<grid, 256, (nb + 3) * sizeof(double), st>>>(n, *p);
  }
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace xrtb
