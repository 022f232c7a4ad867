// C ABI of libxrt_b200.so (declarations: include/xrt_b200.h).
#include <math.h>
#include <stdarg.h>

#include <mutex>
#include <string>

#include "../../include/xrt_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace xrtb {

static thread_local char g_err[1024] = "";
static int g_sms = 0;

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  if (g_sms == 0) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess)
      g_sms = prop.multiProcessorCount;
    else
      g_sms = 148;
  }
  return g_sms;
}

// ---------------------------------------------------------------- fp64 ceilings
template <int MODE>
__global__ void __launch_bounds__(256) fp64_probe_kernel(int iters, double* sink) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double m = 1.0000000001, c = 1e-12;
  double d0[4], d1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) d0[i] = d1[i] = 0.0;
  const double fa = 1.0 + 1e-12 * threadIdx.x, fb = 1.0 - 1e-12 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(d0[i]), "+d"(d1[i])
                     : "d"(fa), "d"(fb));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += d0[i] + d1[i];
  if (s == 123.456) sink[0] = s;
}

static int fp64_probe(int mode, int iters, double* tflops, double* ms_out) {
  XRTB_REQUIRE(mode >= 0 && mode <= 2, "fp64_probe: mode must be 0, 1 or 2");
  double* sink = nullptr;
  XRTB_CUDA_CHECK(cudaMalloc(&sink, 8));
  const int ctas = num_sms() * 8;
  cudaEvent_t e0, e1;
  XRTB_CUDA_CHECK(cudaEventCreate(&e0));
  XRTB_CUDA_CHECK(cudaEventCreate(&e1));
  auto run = [&](int n) {
    if (mode == 0) fp64_probe_kernel<0><<<ctas, 256>>>(n, sink);
    if (mode == 1) fp64_probe_kernel<1><<<ctas, 256>>>(n, sink);
    if (mode == 2) fp64_probe_kernel<2><<<ctas, 256>>>(n, sink);
  };
  run(iters / 8 + 1);
  XRTB_CUDA_CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    XRTB_CUDA_CHECK(cudaEventRecord(e0));
    run(iters);
    XRTB_CUDA_CHECK(cudaEventRecord(e1));
    XRTB_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    XRTB_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  XRTB_CUDA_CHECK(cudaGetLastError());
  double threads = (double)ctas * 256;
  double flop = 0;
  if (mode == 0 || mode == 2) flop += threads * 8.0 * 2.0 * iters;
  if (mode == 1 || mode == 2) flop += (threads / 32.0) * 4.0 * (8 * 8 * 4 * 2.0) * iters;
  *tflops = flop / (best * 1e-3) * 1e-12;
  if (ms_out) *ms_out = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return 0;
}

// ---------------------------------------------------------------- device math helpers, exposed for known-answer tests
__global__ void math_probe_kernel(int which, size_t n, const double* __restrict__ x, double* __restrict__ y0,
                                  double* __restrict__ y1, const double2* __restrict__ sintab) {
  extern __shared__ __align__(16) unsigned char probe_smem[];
  double2* sctab = reinterpret_cast<double2*>(probe_smem);  // [kScTabEntries], filled for the table sincos only
  if (which == 8) sincos_tab_fill(sctab, threadIdx.x, blockDim.x);
  __syncthreads();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  double a = 0.0, b = 0.0;
  switch (which) {
    case 0: sincos_big(v, a, b); break;
    case 1: a = sin_big(v); break;
    case 2: a = rcp_fast(v); break;
    case 3: sqrt_rsqrt(v, a, b); break;
    case 4: bessel_k13_k23(v, a, b); break;
    case 6: a = rcp_cubic(v); break;
    case 7: sqrt_hrsqrt(v, a, b); break;
    case 8: sincos_tab(v, sctab, a, b); break;
    case 9: a = sin_tab(v, sintab); break;
    case 10:  // the unchecked IEEE sequences: x -> (1 / x, sqrt(|x|)); pairs (x[2i], x[2i+1]) -> quotient in y1 of the even slot
      a = div_nr(1.0, v);
      b = sqrt_nr(fabs(v));
      break;
    case 11: a = div_nr(v, x[i ^ 1]); break;  // x[i] / x[i ^ 1]
    default: break;
  }
  y0[i] = a;
  y1[i] = b;
}

// ---------------------------------------------------------------- host staging helper
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  int alloc(size_t bytes) {
    if (bytes == 0) bytes = 16;
    XRTB_CUDA_CHECK(cudaMalloc(&p, bytes));
    return 0;
  }
  int upload(const void* h, size_t bytes, cudaStream_t st) {
    if (alloc(bytes)) return 1;
    if (bytes) XRTB_CUDA_CHECK(cudaMemcpyAsync(p, h, bytes, cudaMemcpyHostToDevice, st));
    return 0;
  }
  template <typename T>
  T* as() {
    return reinterpret_cast<T*>(p);
  }
};

}  // namespace xrtb

using namespace xrtb;

extern "C" {

int xrt_b200_version(void) { return 100; }

int xrt_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int xrt_b200_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_last_error("xrt_b200_init: no CUDA device available (%s); this library has no CPU fallback",
                   e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return 1;
  }
  XRTB_REQUIRE(device >= 0 && device < n, "xrt_b200_init: device %d out of range [0, %d)", device, n);
  XRTB_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  XRTB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  XRTB_REQUIRE(prop.major >= 10, "xrt_b200_init: device %d is sm_%d%d; this build targets sm_100a (B200)", device,
               prop.major, prop.minor);
  g_sms = prop.multiProcessorCount;
  return 0;
}

int xrt_b200_sm_count(void) { return num_sms(); }

size_t xrt_b200_sizeof(int which) {
  return which == 0 ? sizeof(xrt_beam) : which == 1 ? sizeof(xrt_oe_params) : which == 2 ? sizeof(xrt_material_params) : which == 3 ? sizeof(xrt_plane_params) : which == 4 ? sizeof(xrt_source_params) : which == 5 ? sizeof(xrt_plot_params) : which == 6 ? sizeof(xrt_multilayer_params) : which == 7 ? sizeof(xrt_bm_params) : which == 8 ? sizeof(xrt_geom_params) : 0;
}

const char* xrt_b200_last_error(void) { return g_err; }

int xrt_b200_sync(void* stream) {
  XRTB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int xrt_b200_fp64_probe(int mode, int iters, double* tflops, double* ms) {
  return fp64_probe(mode, iters, tflops, ms);
}

// ------------------------------------------------------------------ Kirchhoff
static int k_R(int flags) {
  int R = (flags >> 4) & 7;
  return R == 0 ? 2 : R;
}

size_t xrt_kirchhoff_workspace_bytes(size_t M, size_t N, int flags) {
  return kirchhoff_workspace_bytes(M, N, k_R(flags));
}

int xrt_kirchhoff_f64(size_t M, size_t N, const double* xM, const double* yM, const double* zM, const double* nl,
                      const double* Es, const double* Ep, const double* k, const double* pos, int pos_stride,
                      const double* nrm, int nrm_stride, double* S, double* P, double* A, double* B, double* C,
                      int flags, void* workspace, void* stream) {
  XRTB_REQUIRE(pos_stride >= 3 && nrm_stride >= 3, "kirchhoff: pos/nrm strides must be >= 3");
  XRTB_REQUIRE(workspace != nullptr, "kirchhoff: workspace is NULL");
  return kirchhoff_device(M, N, xM, yM, zM, nl, Es, Ep, k, pos, pos_stride, nrm, nrm_stride, S, P, A, B, C, flags,
                          workspace, (cudaStream_t)stream);
}

int xrt_kirchhoff_pack_f64(size_t N, const double* nl, const double* Es, const double* Ep, const double* k,
                           const double* pos, int pos_stride, const double* nrm, int nrm_stride, void* records,
                           void* stream) {
  XRTB_REQUIRE(pos_stride >= 3 && nrm_stride >= 3, "kirchhoff_pack: pos/nrm strides must be >= 3");
  XRTB_REQUIRE(N == 0 || (nl && Es && Ep && k && pos && nrm && records), "kirchhoff_pack: NULL argument");
  return kirchhoff_pack_device(N, nl, Es, Ep, k, pos, pos_stride, nrm, nrm_stride, records, (cudaStream_t)stream);
}

int xrt_kirchhoff_packed_f64(size_t M, size_t N, const double* xM, const double* yM, const double* zM,
                             const void* records, double* S, double* P, double* A, double* B, double* C, int flags,
                             void* workspace, void* stream) {
  XRTB_REQUIRE(workspace != nullptr && (N == 0 || records != nullptr), "kirchhoff_packed: NULL argument");
  return kirchhoff_packed_device(M, N, xM, yM, zM, records, S, P, A, B, C, flags, workspace, (cudaStream_t)stream);
}

int xrt_kirchhoff_f64_host(size_t M, size_t N, const double* xM, const double* yM, const double* zM,
                           const double* nl, const double* Es, const double* Ep, const double* k, const double* pos,
                           int pos_stride, const double* nrm, int nrm_stride, double* S, double* P, double* A,
                           double* B, double* C, int flags) {
  XRTB_REQUIRE(pos_stride >= 3 && nrm_stride >= 3, "kirchhoff: pos/nrm strides must be >= 3");
  if (M == 0) return 0;
  cudaStream_t st = 0;
  DevBuf dx, dy, dz, dnl, des, dep, dk, dpos, dnrm, dout, dws;
  if (dx.upload(xM, M * 8, st) || dy.upload(yM, M * 8, st) || dz.upload(zM, M * 8, st) ||
      dnl.upload(nl, N * 8, st) || des.upload(Es, N * 16, st) || dep.upload(Ep, N * 16, st) ||
      dk.upload(k, N * 8, st) || dpos.upload(pos, N * 8 * pos_stride, st) ||
      dnrm.upload(nrm, N * 8 * nrm_stride, st))
    return 1;
  if (dout.alloc(5 * M * 16)) return 1;
  if (dws.alloc(xrt_kirchhoff_workspace_bytes(M, N, flags))) return 1;
  double* o = dout.as<double>();
  int rc = kirchhoff_device(M, N, dx.as<double>(), dy.as<double>(), dz.as<double>(), dnl.as<double>(),
                            des.as<double>(), dep.as<double>(), dk.as<double>(), dpos.as<double>(), pos_stride,
                            dnrm.as<double>(), nrm_stride, o, o + 2 * M, o + 4 * M, o + 6 * M, o + 8 * M, flags,
                            dws.p, st);
  if (rc) return rc;
  double* outs[5] = {S, P, A, B, C};
  for (int c = 0; c < 5; ++c)
    XRTB_CUDA_CHECK(cudaMemcpyAsync(outs[c], o + 2 * M * c, M * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

int xrt_b200_math_probe_host(int which, size_t n, const double* x, double* y0, double* y1) {
  if (n == 0) return 0;
  XRTB_REQUIRE(which >= 0 && which <= 11 && x && y0 && y1, "math_probe: bad arguments");
  XRTB_REQUIRE(which != 11 || n % 2 == 0, "math_probe: which = 11 takes pairs");
  cudaStream_t st = 0;
  DevBuf dx, out;
  if (dx.upload(x, n * 8, st) || out.alloc(2 * n * 8)) return 1;
  double* o = out.as<double>();
  const size_t psmem = (which == 8) ? kScTabEntries * sizeof(double2) : 0;
  if (psmem > 48 * 1024)
    XRTB_CUDA_CHECK(cudaFuncSetAttribute(math_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
  const double2* sintab = (which == 9) ? sin_table_device(st) : nullptr;
  XRTB_REQUIRE(which != 9 || sintab, "math_probe: no sin table");
  math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, psmem, st>>>(which, n, dx.as<double>(), o, o + n, sintab);
  XRTB_CUDA_CHECK(cudaGetLastError());
  XRTB_CUDA_CHECK(cudaMemcpyAsync(y0, o, n * 8, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaMemcpyAsync(y1, o + n, n * 8, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

// ------------------------------------------------------------------ undulator
int xrt_undulator_f64(size_t n, int jend, double alpha, double Kx, double Ky, const double* gamma, const double* wu,
                      const double* w, const double* ww1, const double* ddphi, const double* ddpsi, const double* tg,
                      const double* ag, const double* sintg, const double* costg, const double* sintgph,
                      const double* costgph, double* Is, double* Ip, int flags, void* stream) {
  (void)alpha;  // taper / near-field scalar of the sibling kernels; unused by `undulator` (undulator.cl:54)
  return undulator_raw_device(n, jend, Kx, Ky, gamma, wu, w, ww1, ddphi, ddpsi, tg, ag, sintg, costg, sintgph,
                              costgph, Is, Ip, (flags >> 4) & 7, (flags >> 10) & 1, (cudaStream_t)stream);
}

int xrt_undulator_f64_host(size_t n, int jend, double alpha, double Kx, double Ky, const double* gamma,
                           const double* wu, const double* w, const double* ww1, const double* ddphi,
                           const double* ddpsi, const double* tg, const double* ag, const double* sintg,
                           const double* costg, const double* sintgph, const double* costgph, double* Is,
                           double* Ip, int flags) {
  (void)alpha;
  if (n == 0) return 0;
  XRTB_REQUIRE(jend > 0, "undulator: the integration grid is empty");
  cudaStream_t st = 0;
  DevBuf samp[6], tab[6], out;
  const double* hs[6] = {gamma, wu, w, ww1, ddphi, ddpsi};
  const double* ht[6] = {tg, ag, sintg, costg, sintgph, costgph};
  for (int i = 0; i < 6; ++i) {
    if (samp[i].upload(hs[i], n * 8, st)) return 1;
    if (tab[i].upload(ht[i], (size_t)jend * 8, st)) return 1;
  }
  if (out.alloc(2 * n * 16)) return 1;
  double* o = out.as<double>();
  // mirror symmetry of the node tables about tg = 0 (to a few ulps): lets a planar undulator visit node pairs
  int symmetric = 1;
  for (int j = 0; j < jend / 2 && symmetric; ++j) {
    const int k = jend - 1 - j;
    const double tol = 4e-15;  // np.arange / np.sin / np.cos leave asymmetries of a few ulps of pi
    if (fabs(tg[j] + tg[k]) > tol * 4 || fabs(ag[j] - ag[k]) > tol * fabs(ag[j]) || fabs(sintg[j] + sintg[k]) > tol ||
        fabs(costg[j] - costg[k]) > tol)
      symmetric = 0;
  }
  if ((jend & 1) && fabs(tg[jend / 2]) > 4e-15 * 4) symmetric = 0;
  if (flags & (1 << 11)) symmetric = 0;  // caller opts out
  int rc = undulator_raw_device(n, jend, Kx, Ky, samp[0].as<double>(), samp[1].as<double>(), samp[2].as<double>(),
                                samp[3].as<double>(), samp[4].as<double>(), samp[5].as<double>(), tab[0].as<double>(),
                                tab[1].as<double>(), tab[2].as<double>(), tab[3].as<double>(), tab[4].as<double>(),
                                tab[5].as<double>(), o, o + 2 * n, (flags >> 4) & 7, symmetric, st);
  if (rc) return rc;
  XRTB_CUDA_CHECK(cudaMemcpyAsync(Is, o, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaMemcpyAsync(Ip, o + 2 * n, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

int xrt_undulator_imap_f64(size_t n, int jend, double Kx, double Ky, double L0, double Np, double gamma0,
                           double amp_const, double dstep, int bw_mode, const double* gamma, const double* w,
                           const double* theta, const double* psi, const double* tg, const double* ag,
                           const double* sintg, const double* costg, const double* sintgph, const double* costgph,
                           double* I, double* Es, double* Ep, int flags, void* stream) {
  return undulator_imap_device(n, jend, Kx, Ky, L0, Np, gamma0, amp_const, dstep, bw_mode, gamma, w, theta, psi, tg,
                               ag, sintg, costg, sintgph, costgph, I, Es, Ep, (flags >> 4) & 7, (flags >> 10) & 1,
                               (cudaStream_t)stream);
}

int xrt_undulator_ext_f64(int variant, size_t n, int jend, int nmax, double scalar0, double Kx, double Ky,
                          const double* gamma, const double* wu, const double* w, const double* ww1,
                          const double* ddphi, const double* ddpsi, const double* tg, const double* ag,
                          const double* sintg, const double* costg, const double* sintgph, const double* costgph,
                          double* Is, double* Ip, void* stream) {
  return undulator_ext_device(variant, 0, n, jend, nmax, scalar0, Kx, Ky, 0.0, 0.0, 0.0, 0.0, 0, gamma, wu, w, ww1,
                              ddphi, ddpsi, tg, ag, sintg, costg, sintgph, costgph, Is, Ip, nullptr,
                              (cudaStream_t)stream);
}

int xrt_undulator_ext_f64_host(int variant, size_t n, int jend, int nmax, double scalar0, double Kx, double Ky,
                               const double* gamma, const double* wu, const double* w, const double* ww1,
                               const double* ddphi, const double* ddpsi, const double* tg, const double* ag,
                               const double* sintg, const double* costg, const double* sintgph,
                               const double* costgph, double* Is, double* Ip) {
  if (n == 0) return 0;
  XRTB_REQUIRE(jend > 0, "undulator_ext: the integration grid is empty");
  cudaStream_t st = 0;
  DevBuf samp[6], tab[6], out;
  const double* hs[6] = {gamma, wu, w, ww1, ddphi, ddpsi};
  const double* ht[6] = {tg, ag, sintg, costg, sintgph, costgph};
  for (int i = 0; i < 6; ++i) {
    if (samp[i].upload(hs[i], n * 8, st)) return 1;
    if (tab[i].upload(ht[i], (size_t)jend * 8, st)) return 1;
  }
  if (out.alloc(2 * n * 16)) return 1;
  double* o = out.as<double>();
  int rc = undulator_ext_device(variant, 0, n, jend, nmax, scalar0, Kx, Ky, 0.0, 0.0, 0.0, 0.0, 0,
                                samp[0].as<double>(), samp[1].as<double>(), samp[2].as<double>(),
                                samp[3].as<double>(), samp[4].as<double>(), samp[5].as<double>(), tab[0].as<double>(),
                                tab[1].as<double>(), tab[2].as<double>(), tab[3].as<double>(), tab[4].as<double>(),
                                tab[5].as<double>(), o, o + 2 * n, nullptr, st);
  if (rc) return rc;
  XRTB_CUDA_CHECK(cudaMemcpyAsync(Is, o, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaMemcpyAsync(Ip, o + 2 * n, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

int xrt_undulator_ext_imap_f64(int variant, size_t n, int jend, int nmax, double scalar0, double Kx, double Ky,
                               double L0, double gamma0, double amp_const, double dstep, int bw_mode,
                               const double* gamma, const double* w, const double* theta, const double* psi,
                               const double* tg, const double* ag, const double* sintg, const double* costg,
                               const double* sintgph, const double* costgph, double* I, double* Es, double* Ep,
                               void* stream) {
  return undulator_ext_device(variant, 1, n, jend, nmax, scalar0, Kx, Ky, L0, gamma0, amp_const, dstep, bw_mode,
                              gamma, nullptr, w, nullptr, theta, psi, tg, ag, sintg, costg, sintgph, costgph, Es, Ep,
                              I, (cudaStream_t)stream);
}

int xrt_custom_field_f64(size_t n, int jend, double betam, double R0, const double* gamma, const double* w,
                         const double* ddtheta, const double* ddpsi, const double* tg, const double* ag,
                         const double* Bx, const double* By, const double* Bz, const double* betax,
                         const double* betay, const double* trajx, const double* trajy, const double* trajz,
                         double* Is, double* Ip, void* stream) {
  return custom_field_device(n, jend, betam, R0, gamma, w, ddtheta, ddpsi, tg, ag, Bx, By, Bz, betax, betay, trajx,
                             trajy, trajz, Is, Ip, (cudaStream_t)stream);
}

int xrt_custom_field_f64_host(size_t n, int jend, double betam, double R0, const double* gamma, const double* w,
                              const double* ddtheta, const double* ddpsi, const double* tg, const double* ag,
                              const double* Bx, const double* By, const double* Bz, const double* betax,
                              const double* betay, const double* trajx, const double* trajy, const double* trajz,
                              double* Is, double* Ip) {
  if (n == 0) return 0;
  XRTB_REQUIRE(jend > 0, "custom_field: the integration grid is empty");
  cudaStream_t st = 0;
  DevBuf samp[4], tab[10], out;
  const double* hs[4] = {gamma, w, ddtheta, ddpsi};
  const double* ht[10] = {tg, ag, Bx, By, Bz, betax, betay, trajx, trajy, trajz};
  for (int i = 0; i < 4; ++i)
    if (samp[i].upload(hs[i], n * 8, st)) return 1;
  for (int i = 0; i < 10; ++i)
    if (tab[i].upload(ht[i], (size_t)jend * 8, st)) return 1;
  if (out.alloc(2 * n * 16)) return 1;
  double* o = out.as<double>();
  int rc = custom_field_device(n, jend, betam, R0, samp[0].as<double>(), samp[1].as<double>(), samp[2].as<double>(),
                               samp[3].as<double>(), tab[0].as<double>(), tab[1].as<double>(), tab[2].as<double>(),
                               tab[3].as<double>(), tab[4].as<double>(), tab[5].as<double>(), tab[6].as<double>(),
                               tab[7].as<double>(), tab[8].as<double>(), tab[9].as<double>(), o, o + 2 * n, st);
  if (rc) return rc;
  XRTB_CUDA_CHECK(cudaMemcpyAsync(Is, o, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaMemcpyAsync(Ip, o + 2 * n, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

// ------------------------------------------------------------------ reflect
int xrt_reflect_f64(size_t n, const xrt_oe_params* oe, const xrt_material_params* mat, const double* tables,
                    size_t table_len, const xrt_beam* in, const xrt_beam* fallback, const xrt_beam* gb,
                    const xrt_beam* lb, uint64_t* reductions, int phase, uint64_t index_offset, void* stream) {
  return reflect_device(n, oe, mat, tables, table_len, in, fallback, gb, lb,
                        reinterpret_cast<unsigned long long*>(reductions), phase, index_offset, (cudaStream_t)stream);
}

int xrt_reflect_expose_f64(size_t n, const xrt_oe_params* oe, const xrt_material_params* mat, const double* tables,
                           size_t table_len, const xrt_beam* in, const xrt_beam* fallback, const xrt_beam* lb,
                           uint64_t* reductions, const xrt_plane_params* screen, const xrt_beam* out,
                           uint64_t* screen_reductions, int phase, uint64_t index_offset, void* stream) {
  return reflect_expose_device(n, oe, mat, tables, table_len, in, fallback, nullptr, lb,
                               reinterpret_cast<unsigned long long*>(reductions), screen, out,
                               reinterpret_cast<unsigned long long*>(screen_reductions), phase, index_offset,
                               (cudaStream_t)stream);
}

int xrt_reflect_multi_deriv_f64(size_t n, const xrt_oe_params* oe, const xrt_beam* in, uint64_t* reductions,
                                double* tD, int phase, void* stream) {
  return reflect_multi_deriv_device(n, oe, in, reinterpret_cast<unsigned long long*>(reductions), tD, phase,
                                    (cudaStream_t)stream);
}

int xrt_reflect_multi_elevation_f64(size_t n, const xrt_oe_params* oe, const xrt_beam* in, const double* tD,
                                    double* elevationD, double* elevationX, double* elevationY, double* elevationZ,
                                    void* stream) {
  return reflect_multi_elevation_device(n, oe, in, tD, elevationD, elevationX, elevationY, elevationZ,
                                        (cudaStream_t)stream);
}

double xrt_reduction_decode(uint64_t slot) { return reduction_decode(slot); }

// ------------------------------------------------------------------ plane elements
int xrt_expose_f64(size_t n, const xrt_plane_params* p, const xrt_beam* in, const xrt_beam* out,
                   uint64_t* reductions, void* stream) {
  return expose_device(n, p, in, out, reinterpret_cast<unsigned long long*>(reductions), (cudaStream_t)stream);
}

int xrt_expose_hemispheric_f64(size_t n, const xrt_plane_params* p, double R, double thetaOffset, double phiOffset,
                               const xrt_beam* in, const xrt_beam* out, void* stream) {
  return expose_hemispheric_device(n, p, R, thetaOffset, phiOffset, in, out, (cudaStream_t)stream);
}

int xrt_propagate_rect_f64(size_t n, const xrt_plane_params* p, const xrt_beam* in, const xrt_beam* out,
                           uint64_t* reductions, void* stream) {
  return propagate_rect_device(n, p, in, out, reinterpret_cast<unsigned long long*>(reductions),
                               (cudaStream_t)stream);
}

// ------------------------------------------------------------------ stand-alone amplitudes
int xrt_get_amplitude_f64_host(size_t n, const xrt_material_params* mat, const double* tables, size_t table_len,
                               const double* E, const double* bIn, const double* bOut, const double* bH, double* rs,
                               double* rp) {
  if (n == 0) return 0;
  cudaStream_t st = 0;
  DevBuf dt, dE, dIn, dOut, dH, dout, dcnt;
  if (dt.upload(tables, 3 * table_len * 8, st) || dE.upload(E, n * 8, st) || dIn.upload(bIn, n * 8, st)) return 1;
  if (bOut && dOut.upload(bOut, n * 8, st)) return 1;
  if (bH && dH.upload(bH, n * 8, st)) return 1;
  if (dout.alloc(2 * n * 16) || dcnt.alloc(8)) return 1;
  XRTB_CUDA_CHECK(cudaMemsetAsync(dcnt.p, 0, 8, st));
  double* o = dout.as<double>();
  // a multilayer's thickness arrays are host arrays at this entry point: stage them and patch the pointers
  xrt_material_params mcopy = *mat;
  xrt_multilayer_params mlcopy;
  DevBuf ddti, ddbi;
  if (mat->kind == 4 && mat->multilayer) {
    mlcopy = *mat->multilayer;
    XRTB_REQUIRE(mlcopy.nPairs >= 0 && mlcopy.dti && mlcopy.dbi, "get_amplitude: multilayer thickness arrays missing");
    if (ddti.upload(mlcopy.dti, (size_t)mlcopy.nPairs * 8, st) || ddbi.upload(mlcopy.dbi, (size_t)mlcopy.nPairs * 8, st))
      return 1;
    mlcopy.dti = ddti.as<double>();
    mlcopy.dbi = ddbi.as<double>();
    mcopy.multilayer = &mlcopy;
  }
  mat = &mcopy;
  int rc = amplitude_device(n, mat, dt.as<double>(), table_len, dE.as<double>(), dIn.as<double>(),
                            bOut ? dOut.as<double>() : nullptr, bH ? dH.as<double>() : nullptr, o, o + 2 * n,
                            dcnt.as<unsigned long long>(), st);
  if (rc) return rc;
  unsigned long long cnt = 0;
  XRTB_CUDA_CHECK(cudaMemcpyAsync(rs, o, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaMemcpyAsync(rp, o + 2 * n, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaMemcpyAsync(&cnt, dcnt.p, 8, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (cnt) {
    set_last_error("get_amplitude: %llu energies are out of the f1/f2 data table range", cnt);
    return 3;
  }
  return 0;
}

// ------------------------------------------------------------------ f2: multilayer recursion
static int all_equal(const double* v, int n) {
  for (int i = 1; i < n; ++i)
    if (v[i] != v[0]) return 0;
  return 1;
}

int xrt_multilayer_f64(size_t n, int npairs, const double* rbs_s, const double* rbs_p, const double* rtb_s,
                       const double* rtb_p, const double* rvt_s, const double* rvt_p, const double* Qt,
                       const double* Qb, const double* dti, const double* dbi, double* ri_s, double* ri_p,
                       void* stream) {
  return multilayer_device(n, npairs, 0, rbs_s, rbs_p, rtb_s, rtb_p, rvt_s, rvt_p, Qt, Qb, dti, dbi, ri_s, ri_p,
                           (cudaStream_t)stream);
}

int xrt_multilayer_f64_host(size_t n, int npairs, const double* rbs_s, const double* rbs_p, const double* rtb_s,
                            const double* rtb_p, const double* rvt_s, const double* rvt_p, const double* Qt,
                            const double* Qb, const double* dti, const double* dbi, double* ri_s, double* ri_p) {
  if (n == 0) return 0;
  XRTB_REQUIRE(npairs >= 0 && dti && dbi, "multilayer: nPairs >= 0 and the thickness arrays are required");
  cudaStream_t st = 0;
  const double* in[8] = {rbs_s, rbs_p, rtb_s, rtb_p, rvt_s, rvt_p, Qt, Qb};
  DevBuf d[8], ddti, ddbi, dout;
  for (int k = 0; k < 8; ++k) {
    XRTB_REQUIRE(in[k] != nullptr, "multilayer: NULL argument");
    if (d[k].upload(in[k], n * 16, st)) return 1;
  }
  if (ddti.upload(dti, (size_t)npairs * 8, st) || ddbi.upload(dbi, (size_t)npairs * 8, st)) return 1;
  if (dout.alloc(2 * n * 16)) return 1;
  double* o = dout.as<double>();
  const int uniform = all_equal(dti, npairs) && all_equal(dbi, npairs);
  int rc = multilayer_device(n, npairs, uniform, d[0].as<double>(), d[1].as<double>(), d[2].as<double>(),
                             d[3].as<double>(), d[4].as<double>(), d[5].as<double>(), d[6].as<double>(),
                             d[7].as<double>(), ddti.as<double>(), ddbi.as<double>(), o, o + 2 * n, st);
  if (rc) return rc;
  XRTB_CUDA_CHECK(cudaMemcpyAsync(ri_s, o, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaMemcpyAsync(ri_p, o + 2 * n, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

int xrt_multilayer_tran_f64(size_t n, int npairs, double substThickness, const double* const* in23, const double* dti,
                            const double* dbi, double* ti_s, double* ti_p, void* stream) {
  return multilayer_tran_device(n, npairs, substThickness, in23, dti, dbi, ti_s, ti_p, (cudaStream_t)stream);
}

int xrt_multilayer_tran_f64_host(size_t n, int npairs, double substThickness, const double* const* in23,
                                 const double* dti, const double* dbi, double* ti_s, double* ti_p) {
  if (n == 0) return 0;
  XRTB_REQUIRE(npairs >= 0 && in23 && dti && dbi, "multilayer_tran: nPairs >= 0, the inputs and the thickness arrays");
  cudaStream_t st = 0;
  DevBuf d[23], ddti, ddbi, dout;
  const double* dev[23];
  for (int k = 0; k < 23; ++k) {
    XRTB_REQUIRE(in23[k] != nullptr, "multilayer_tran: input %d is NULL", k);
    if (d[k].upload(in23[k], n * 16, st)) return 1;
    dev[k] = d[k].as<double>();
  }
  if (ddti.upload(dti, (size_t)npairs * 8, st) || ddbi.upload(dbi, (size_t)npairs * 8, st)) return 1;
  if (dout.alloc(2 * n * 16)) return 1;
  double* o = dout.as<double>();
  int rc = multilayer_tran_device(n, npairs, substThickness, dev, ddti.as<double>(), ddbi.as<double>(), o, o + 2 * n,
                                  st);
  if (rc) return rc;
  XRTB_CUDA_CHECK(cudaMemcpyAsync(ti_s, o, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaMemcpyAsync(ti_p, o + 2 * n, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

// ------------------------------------------------------------------ source
int xrt_source_sample_f64(size_t n, uint64_t seed, uint64_t index_offset, const double* window, double* E,
                          double* theta, double* psi, double* n1, double* n2, double* n3, double* n4, double* u,
                          double* n5, void* stream) {
  return source_sample_device(n, seed, index_offset, window, E, theta, psi, n1, n2, n3, n4, u, n5,
                              (cudaStream_t)stream);
}

int xrt_source_fill_f64(size_t n, const xrt_source_params* p, const int64_t* idx, const double* rE,
                        const double* rTheta, const double* rPsi, const double* mJs, const double* mJp,
                        const double* n1, const double* n2, const double* n3, const double* n4, const xrt_beam* out,
                        size_t out_offset, void* stream) {
  return source_fill_device(n, p, reinterpret_cast<const long long*>(idx), rE, rTheta, rPsi, mJs, mJp, n1, n2, n3,
                            n4, out, out_offset, (cudaStream_t)stream);
}

int xrt_shine_uniform_f64(size_t n, uint64_t seed, uint64_t index_offset, const double* window, int jend, double Kx,
                          double Ky, double L0, double Np, double gamma0, double amp_const, double dstep, int bw_mode,
                          int symmetric, const double* tg, const double* ag, const double* sintg, const double* costg,
                          const double* sintgph, const double* costgph, const xrt_source_params* p,
                          const xrt_beam* out, uint64_t* imax_slot, void* stream) {
  return shine_uniform_device(n, seed, index_offset, window, jend, Kx, Ky, L0, Np, gamma0, amp_const, dstep, bw_mode,
                              symmetric, tg, ag, sintg, costg, sintgph, costgph, p, out,
                              reinterpret_cast<unsigned long long*>(imax_slot), (cudaStream_t)stream);
}

int xrt_geometric_source_f64(size_t n, const xrt_geom_params* p, uint64_t seed, uint64_t index_offset,
                             const double* const* deviates, const xrt_beam* out, uint64_t* flag, int phase,
                             void* stream) {
  return geometric_source_device(n, p, seed, index_offset, deviates, out,
                                 reinterpret_cast<unsigned long long*>(flag), phase, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ bending magnet / wiggler
int xrt_bm_imap_f64(size_t n, const xrt_bm_params* bm, const double* gamma, const double* w, const double* theta,
                    const double* psi, double* I, double* Es, double* Ep, void* stream) {
  return bm_imap_device(n, bm, gamma, w, theta, psi, I, Es, Ep, (cudaStream_t)stream);
}

int xrt_bm_imap_f64_host(size_t n, const xrt_bm_params* bm, const double* gamma, const double* w,
                         const double* theta, const double* psi, double* I, double* Es, double* Ep) {
  if (n == 0) return 0;
  XRTB_REQUIRE(bm && w && theta && psi && I && Es && Ep, "bm_imap: NULL argument");
  cudaStream_t st = 0;
  DevBuf dg, dw, dt, dp, out;
  if (gamma && dg.upload(gamma, n * 8, st)) return 1;
  if (dw.upload(w, n * 8, st) || dt.upload(theta, n * 8, st) || dp.upload(psi, n * 8, st)) return 1;
  if (out.alloc(n * 40)) return 1;
  double* o = out.as<double>();
  int rc = bm_imap_device(n, bm, gamma ? dg.as<double>() : nullptr, dw.as<double>(), dt.as<double>(),
                          dp.as<double>(), o, o + n, o + 3 * n, st);
  if (rc) return rc;
  XRTB_CUDA_CHECK(cudaMemcpyAsync(I, o, n * 8, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaMemcpyAsync(Es, o + n, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaMemcpyAsync(Ep, o + 3 * n, n * 16, cudaMemcpyDeviceToHost, st));
  XRTB_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

int xrt_bm_fill_f64(size_t n, const xrt_source_params* p, const xrt_bm_params* bm, const long long* idx,
                    const double* rE, const double* rTheta, const double* rPsi, const double* mJs, const double* mJp,
                    const double* g1, const double* g2, const double* g3, const double* g4, const double* g5,
                    const double* kint, const xrt_beam* out, size_t out_offset, void* stream) {
  return bm_fill_device(n, p, bm, idx, rE, rTheta, rPsi, mJs, mJp, g1, g2, g3, g4, g5, kint, out, out_offset,
                        (cudaStream_t)stream);
}

int xrt_bm_shine_uniform_f64(size_t n, uint64_t seed, uint64_t index_offset, const double* window,
                             const xrt_source_params* sp, const xrt_bm_params* bm, const xrt_beam* out,
                             uint64_t* imax_slot, void* stream) {
  return bm_shine_uniform_device(n, seed, index_offset, window, sp, bm, out,
                                 reinterpret_cast<unsigned long long*>(imax_slot), (cudaStream_t)stream);
}

int xrt_oe_surface_f64(size_t n, const xrt_oe_params* oe, const double* x, const double* y, double* z,
                       void* stream) {
  return surface_device(n, oe, x, y, z, (cudaStream_t)stream);
}

int xrt_oe_normals_f64(size_t n, const xrt_oe_params* oe, const double* x, const double* y, const double* z,
                       double* nx, double* ny, double* nz, void* stream) {
  return normals_device(n, oe, x, y, z, nx, ny, nz, (cudaStream_t)stream);
}

int xrt_plot_hist_f64(size_t n, const xrt_plot_params* plot, void* stream) {
  return plot_hist_device(n, plot, (cudaStream_t)stream);
}

size_t xrt_compact_scratch_words(size_t n) { return compact_scratch_words(n); }

int xrt_compact_rays_f64(size_t n, const int32_t* state, int state_mask, int narr, const double* const* src,
                         double* const* dst, const int* words, uint64_t* scratch, void* stream) {
  return compact_rays_device(n, state, state_mask, narr, src, dst, words,
                             reinterpret_cast<unsigned long long*>(scratch), (cudaStream_t)stream);
}

int xrt_hist2d_f64(size_t n, const double* x, const double* y, const double* w1, const double* w2,
                   const int32_t* state, int state_mask, const double* ex, int nx, const double* ey, int ny,
                   double* out, void* stream) {
  return hist2d_device(n, x, y, w1, w2, state, state_mask, ex, nx, ey, ny, out, (cudaStream_t)stream);
}

}  // extern "C"
