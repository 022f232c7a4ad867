// Fresnel-Kirchhoff N x M propagation sum for sm_100a.
//
// Replaces: kernel `integrate_kirchhoff` (xrt/backends/raycing/cl/diffract.cl:80-151) as launched from
// waves._diffraction_integral_CL (xrt/backends/raycing/waves.py:854-896); numerically it follows the numpy
// path waves._diffraction_integral_conv (waves.py:834-851), which is the oracle (oracle/kirchhoff.py).
//
// For every output (receiving) point p_i and every input (diffracting) sample j
//     d = p_i - q_j,  r = |d|,  G_ij = (k_j / r) (n_j.d / r + nl_j) exp(i k_j r)
//     S_i = c1 sum_j G_ij Es_j          P_i = c1 sum_j G_ij Ep_j
//     (A,B,C)_i = c2 sum_j (k_j^2 / r) G_ij (Es_j + Ep_j) d
// with (c1, c2) = (i/4pi, i/(4pi)^2) in the numpy convention and (-i/4pi, (1+i)/4pi) in the OpenCL one.
//
// The matrix is generated and consumed on the fly: HBM traffic is 24 M + 112 N + 80 M bytes, i.e. nothing;
// the kernel is bound by the FP64 pipe (about 70 DFMA-pipe instructions per pair).  Design:
//   * inputs are packed once into 128-byte records and streamed through shared memory with 1-D bulk (TMA)
//     copies on an mbarrier double buffer; every lane reads the same record -> LDS broadcasts;
//   * each thread owns R outputs (register blocking) so one record read feeds R independent FP64 chains;
//   * |d|^2 is evaluated in numpy's rounding order with explicit non-contracted mul/add so that r and the
//     phase k*r are bit-identical to the oracle; sqrt and 1/sqrt share one MUFU seed; sin/cos use an exact
//     FMA Cody-Waite reduction valid for the 1e10..1e12 rad phases of hard X-rays;
//   * small M is handled by splitting N across blockIdx.y into a workspace reduced in a fixed order
//     (deterministic, no atomics).
#include "common.cuh"

namespace xrtb {

struct __align__(16) KInput {
  double qx, qy, qz, nl;
  double nx2, ny2, nz2, k;  // 2 n (the kernel works with 0.5/r, see sqrt_hrsqrt), k
  double es_re, es_im, ep_re, ep_im;
  double w_re, w_im, k2x, nq2;  // k (Es + Ep), 2 k, 2 n.q (uniform-normal fast path)
};
static_assert(sizeof(KInput) == 128, "KInput must be 128 bytes");

// ---------------------------------------------------------------- pack
__global__ void kirchhoff_pack_kernel(size_t N, const double* __restrict__ nl, const double2* __restrict__ Es,
                                      const double2* __restrict__ Ep, const double* __restrict__ k,
                                      const double* __restrict__ pos, int pos_stride,
                                      const double* __restrict__ nrm, int nrm_stride, KInput* __restrict__ out) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= N) return;
  KInput r;
  r.qx = pos[j * pos_stride + 0];
  r.qy = pos[j * pos_stride + 1];
  r.qz = pos[j * pos_stride + 2];
  const double nx = nrm[j * nrm_stride + 0], ny = nrm[j * nrm_stride + 1], nz = nrm[j * nrm_stride + 2];
  r.nx2 = 2.0 * nx;
  r.ny2 = 2.0 * ny;
  r.nz2 = 2.0 * nz;
  r.nl = nl[j];
  double kj = k[j];
  r.k = kj;
  double2 es = Es[j], ep = Ep[j];
  r.es_re = es.x;
  r.es_im = es.y;
  r.ep_re = ep.x;
  r.ep_im = ep.y;
  // numpy forms (Es + Ep) first as well (waves.py:847); the factor k of k^2 / r is folded in here, once per sample,
  // instead of once per pair (one FP64 instruction less in the pair loop; A, B, C carry no bit-exactness requirement)
  r.w_re = kj * (es.x + ep.x);
  r.w_im = kj * (es.y + ep.y);
  r.k2x = 2.0 * kj;
  r.nq2 = 2.0 * (nx * r.qx + ny * r.qy + nz * r.qz);
  out[j] = r;
}

// ---------------------------------------------------------------- main kernel
// UNIFORM_N: every sample has the same normal n (aperture, flat OE): n.d = n.p_i - n.q_j with n.p_i hoisted out
// of the loop (exactly d_y for the aperture normal (0, 1, 0)).
#ifndef XRTB_KIRCHHOFF_MINBLOCKS
#define XRTB_KIRCHHOFF_MINBLOCKS 2
#endif
template <int R, int THREADS, int TILE, bool STRICT, bool UNIFORM_N, bool TABSC>
__global__ void __launch_bounds__(THREADS, XRTB_KIRCHHOFF_MINBLOCKS)
    kirchhoff_kernel(size_t M, size_t N, const double* __restrict__ xM, const double* __restrict__ yM,
                     const double* __restrict__ zM, const KInput* __restrict__ in, size_t n_per_split,
                     double2* __restrict__ part, const double* __restrict__ un, double un_scale) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  KInput* tile = reinterpret_cast<KInput*>(smem_raw);                       // [2][TILE]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + 2 * TILE * sizeof(KInput));  // [2]
  double2* sctab = reinterpret_cast<double2*>(smem_raw + 2 * TILE * sizeof(KInput) + 16);  // [kScTabEntries] (TABSC)
  // the common normal of the uniform-normal variant is the first sample's, read here from device memory (r01 fetched
  // it to the host with a blocking copy inside an entry point that promises to be asynchronous)
  const double un_x = UNIFORM_N ? un[0] * un_scale : 0.0, un_y = UNIFORM_N ? un[1] * un_scale : 0.0,
               un_z = UNIFORM_N ? un[2] * un_scale : 0.0;

  const int tid = threadIdx.x;
  const size_t base = (size_t)blockIdx.x * (THREADS * R);
  const size_t j0 = (size_t)blockIdx.y * n_per_split;
  size_t j1 = j0 + n_per_split;
  if (j1 > N) j1 = N;
  const size_t nloc = (j1 > j0) ? (j1 - j0) : 0;
  const int ntiles = (int)((nloc + TILE - 1) / TILE);

  double px[R], py[R], pz[R], np2[R];
  double acc[R][10];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    size_t i = base + (size_t)r * THREADS + tid;
    size_t ii = i < M ? i : (M - 1);
    px[r] = xM[ii];
    py[r] = yM[ii];
    pz[r] = zM[ii];
    np2[r] = 2.0 * (un_x * px[r] + un_y * py[r] + un_z * pz[r]);
#pragma unroll
    for (int c = 0; c < 10; ++c) acc[r][c] = 0.0;
  }

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
  }
  if (TABSC) sincos_tab_fill(sctab, tid, THREADS);
  __syncthreads();

  auto issue = [&](int t) {
    size_t off = j0 + (size_t)t * TILE;
    size_t cnt = j1 - off;
    if (cnt > TILE) cnt = TILE;
    uint32_t bytes = (uint32_t)(cnt * sizeof(KInput));
    int s = t & 1;
    mbar_expect_tx(&full[s], bytes);
    bulk_g2s(tile + s * TILE, in + off, bytes, &full[s]);
  };
  if (tid == 0) {
    if (ntiles > 0) issue(0);
    if (ntiles > 1) issue(1);
  }

  for (int t = 0; t < ntiles; ++t) {
    const int s = t & 1;
    mbar_wait(&full[s], (t >> 1) & 1);
    const KInput* rec = tile + s * TILE;
    size_t off = (size_t)t * TILE;
    int cnt = (int)((nloc - off) < (size_t)TILE ? (nloc - off) : (size_t)TILE);

#pragma unroll 1
    for (int j = 0; j < cnt; ++j) {
      const double4* q4 = reinterpret_cast<const double4*>(rec + j);
      const double4 a0 = q4[0];  // qx qy qz nl
      const double4 a1 = q4[1];  // nx ny nz k
      const double4 a2 = q4[2];  // es ep
      const double4 a3 = q4[3];  // w  k2
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const double dx = px[r] - a0.x;
        const double dy = py[r] - a0.y;
        const double dz = pz[r] - a0.z;
        double r2;
        if (STRICT) {
          // numpy: (a**2 + b**2) + c**2 with one rounding per operation (waves.py:839)
          r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        } else {
          r2 = fma(dz, dz, fma(dy, dy, dx * dx));
        }
        double rr, hinv;  // hinv = 0.5 / r
        sqrt_hrsqrt(r2, rr, hinv);
        const double dn2 = UNIFORM_N ? (np2[r] - a3.w) : fma(dz, a1.z, fma(dy, a1.y, dx * a1.x));  // 2 n.d
        const double tcos = fma(dn2, hinv, a0.w);  // n.d/r + nl
        const double phase = STRICT ? __dmul_rn(a1.w, rr) : a1.w * rr;
        const double kip = a3.z * hinv;  // k / r
        const double cr = kip * tcos;
        double sn, cs;
        // (sincos_tab reduces with 2^11/pi: exact below 2^40 pi rad = 3.4e12 -- 100 keV over 6.8 m, 9 keV over 75 m.  The
        // CALLER vouches for that range with flag bit 11: a per-pair test in this loop costs 40 % of the kernel)
        if (TABSC)
          sincos_tab(phase, sctab, sn, cs);
        else
          sincos_big(phase, sn, cs);
        const double gre = cr * cs;
        const double gim = cr * sn;
        // S += g * Es ; P += g * Ep
        acc[r][0] = fma(gre, a2.x, acc[r][0]);
        acc[r][0] = fma(-gim, a2.y, acc[r][0]);
        acc[r][1] = fma(gre, a2.y, acc[r][1]);
        acc[r][1] = fma(gim, a2.x, acc[r][1]);
        acc[r][2] = fma(gre, a2.z, acc[r][2]);
        acc[r][2] = fma(-gim, a2.w, acc[r][2]);
        acc[r][3] = fma(gre, a2.w, acc[r][3]);
        acc[r][3] = fma(gim, a2.z, acc[r][3]);
        // h = (k / r) * g * (k (Es + Ep)) = (k^2 / r) * g * (Es + Ep)
        const double wre = fma(gre, a3.x, -gim * a3.y);
        const double wim = fma(gre, a3.y, gim * a3.x);
        const double hre = kip * wre;
        const double him = kip * wim;
        acc[r][4] = fma(hre, dx, acc[r][4]);
        acc[r][5] = fma(him, dx, acc[r][5]);
        acc[r][6] = fma(hre, dy, acc[r][6]);
        acc[r][7] = fma(him, dy, acc[r][7]);
        acc[r][8] = fma(hre, dz, acc[r][8]);
        acc[r][9] = fma(him, dz, acc[r][9]);
      }
    }
    __syncthreads();  // every thread is done with stage s
    if (tid == 0 && t + 2 < ntiles) issue(t + 2);
  }

  // partial sums: part[(split * 5 + c) * M + i]
#pragma unroll
  for (int r = 0; r < R; ++r) {
    size_t i = base + (size_t)r * THREADS + tid;
    if (i < M) {
#pragma unroll
      for (int c = 0; c < 5; ++c)
        part[((size_t)blockIdx.y * 5 + c) * M + i] = make_double2(acc[r][2 * c], acc[r][2 * c + 1]);
    }
  }
}

// ---------------------------------------------------------------- finalize: fixed-order split reduction + constants
__global__ void kirchhoff_finalize_kernel(size_t M, int nsplit, const double2* __restrict__ part, int convention,
                                          double2* __restrict__ S, double2* __restrict__ P, double2* __restrict__ A,
                                          double2* __restrict__ B, double2* __restrict__ C) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  double2 v[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    double re = 0.0, im = 0.0;
    for (int s = 0; s < nsplit; ++s) {
      double2 p = part[((size_t)s * 5 + c) * M + i];
      re += p.x;
      im += p.y;
    }
    v[c] = make_double2(re, im);
  }
  const double inv4pi = 0.07957747154594767;      // 1/(4 pi)
  const double inv4pi2 = 0.006332573977646111;    // 1/(4 pi)^2
  double2 o[5];
  if (convention == 0) {
    // numpy: U = k*1j/(4 pi) ..., abcU = k**2/(4 pi) ... U / r   (waves.py:844-847)
    for (int c = 0; c < 2; ++c) o[c] = make_double2(-v[c].y * inv4pi, v[c].x * inv4pi);
    for (int c = 2; c < 5; ++c) o[c] = make_double2(-v[c].y * inv4pi2, v[c].x * inv4pi2);
  } else {
    // OpenCL: cfactor = -i/(4 pi) for S,P; (1+i)/(4 pi) for A,B,C (diffract.cl:137-149 with :21-24)
    for (int c = 0; c < 2; ++c) o[c] = make_double2(v[c].y * inv4pi, -v[c].x * inv4pi);
    for (int c = 2; c < 5; ++c) o[c] = make_double2((v[c].x - v[c].y) * inv4pi, (v[c].x + v[c].y) * inv4pi);
  }
  S[i] = o[0];
  P[i] = o[1];
  A[i] = o[2];
  B[i] = o[3];
  C[i] = o[4];
}

// ---------------------------------------------------------------- host-side launcher
struct KLaunch {
  dim3 grid;
  size_t smem;
  cudaStream_t st;
  size_t M, N;
  const double *xM, *yM, *zM;
  const KInput* in;
  size_t nps;
  double2* part;
  const double* un;
  double un_scale;  // 1: un is a normal; 0.5: un points at the doubled normal of a packed record
};

template <int R, int THREADS, int TILE, bool STRICT, bool UNI, bool TABSC>
static int launch_one(const KLaunch& L) {
  XRTB_CUDA_CHECK(cudaFuncSetAttribute(kirchhoff_kernel<R, THREADS, TILE, STRICT, UNI, TABSC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  kirchhoff_kernel<R, THREADS, TILE, STRICT, UNI, TABSC><<<L.grid, THREADS, L.smem, L.st>>>(
      L.M, L.N, L.xM, L.yM, L.zM, L.in, L.nps, L.part, L.un, L.un_scale);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <int R, int THREADS, int TILE>
static int launch_main(bool strict, bool uni, bool tabsc, const KLaunch& L) {
  const int v = (strict ? 4 : 0) | (uni ? 2 : 0) | (tabsc ? 1 : 0);
  switch (v) {
    case 0: return launch_one<R, THREADS, TILE, false, false, false>(L);
    case 1: return launch_one<R, THREADS, TILE, false, false, true>(L);
    case 2: return launch_one<R, THREADS, TILE, false, true, false>(L);
    case 3: return launch_one<R, THREADS, TILE, false, true, true>(L);
    case 4: return launch_one<R, THREADS, TILE, true, false, false>(L);
    case 5: return launch_one<R, THREADS, TILE, true, false, true>(L);
    case 6: return launch_one<R, THREADS, TILE, true, true, false>(L);
    default: return launch_one<R, THREADS, TILE, true, true, true>(L);
  }
}

#ifndef XRTB_KIRCHHOFF_THREADS
#define XRTB_KIRCHHOFF_THREADS 128
#endif
constexpr int K_THREADS = XRTB_KIRCHHOFF_THREADS;
constexpr int K_TILE = 128;

size_t kirchhoff_packed_bytes(size_t N) { return N * sizeof(KInput); }

// number of N-splits used for a given problem (also sizes the workspace).  The grid is (M blocks) x (N splits) of
// equal CTAs that the hardware scheduler hands to 2 x 148 resident slots as they free up: the last wave is only
// partly full, so the loss is ~0.5 / waves.  r01 aimed at 4 CTAs per SM (2 - 3 waves): 768 CTAs = 2.59 waves for a
// 131072-point shard of config 4, i.e. the 0.87 strong-scaling efficiency of SCALE_r01.  Now the split is chosen so
// that there are >= 32 waves whenever every split still gets >= 16 tiles (2048 samples, ~1 ms of work per CTA), and
// among the next few candidates the one whose last wave is fullest.  The partial sums are reduced in a fixed order by
// kirchhoff_finalize_kernel, so the result does not depend on the split.
int kirchhoff_plan_splits(size_t M, size_t N, int R) {
  const size_t per_cta = (size_t)K_THREADS * R;
  const size_t mblocks = (M + per_cta - 1) / per_cta;
  const size_t slots = (size_t)num_sms() * XRTB_KIRCHHOFF_MINBLOCKS;
  const size_t target = slots * 32;
  size_t max_ns = N / (K_TILE * 16);
  if (max_ns < 1) max_ns = 1;
  if (max_ns > 4096) max_ns = 4096;
  if (mblocks == 0) return 1;
  size_t ns0 = (target + mblocks - 1) / mblocks;
  if (ns0 < 1) ns0 = 1;
  if (ns0 >= max_ns) return (int)max_ns;
  size_t best = ns0;
  double best_eff = 0.0;
  for (size_t ns = ns0; ns <= max_ns && ns < ns0 + 8; ++ns) {
    const double waves = (double)(mblocks * ns) / (double)slots;
    const double eff = waves / (double)(size_t)(waves + 0.999999);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = ns;
    }
  }
  return (int)best;
}

size_t kirchhoff_workspace_bytes(size_t M, size_t N, int R) {
  int ns = kirchhoff_plan_splits(M, N, R);
  return kirchhoff_packed_bytes(N) + (size_t)ns * 5 * M * sizeof(double2);
}

int kirchhoff_pack_device(size_t N, const double* nl, const double* Es, const double* Ep, const double* k,
                          const double* pos, int pos_stride, const double* nrm, int nrm_stride, void* records,
                          cudaStream_t st) {
  if (N == 0) return 0;
  kirchhoff_pack_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(
      N, nl, reinterpret_cast<const double2*>(Es), reinterpret_cast<const double2*>(Ep), k, pos, pos_stride, nrm,
      nrm_stride, reinterpret_cast<KInput*>(records));
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// the sum from packed records (xrt_kirchhoff_pack_f64; all-gathered over ranks when the samples are ray shards);
// `part` holds kirchhoff_plan_splits(M, N, R) * 5 * M complex partial sums
int kirchhoff_packed_device(size_t M, size_t N, const double* xM, const double* yM, const double* zM,
                            const void* records, double* S, double* P, double* A, double* B, double* C, int flags,
                            void* part_ws, cudaStream_t st) {
  if (M == 0) return 0;
  const bool strict = !(flags & 1);
  const int convention = (flags >> 1) & 1;
  const bool uni = (flags >> 9) & 1;  // caller guarantees that all N normals are equal
  const bool tabsc = (flags >> 11) & 1;  // sincos through the shared-memory table (sincos_tab)
  int R = (flags >> 4) & 7;
  if (R == 0) R = 2;
  XRTB_REQUIRE(R == 1 || R == 2 || R == 4, "kirchhoff: register blocking R must be 1, 2 or 4 (got %d)", R);
  const KInput* packed = reinterpret_cast<const KInput*>(records);
  double2* part = reinterpret_cast<double2*>(part_ws);
  const int nsplit = kirchhoff_plan_splits(M, N, R);
  const size_t per_cta = (size_t)K_THREADS * R;
  dim3 grid((unsigned)((M + per_cta - 1) / per_cta), (unsigned)nsplit);
  size_t nps = (N + nsplit - 1) / nsplit;
  nps = (nps + K_TILE - 1) / K_TILE * K_TILE;
  if (nps == 0) nps = K_TILE;
  const size_t smem = 2 * K_TILE * sizeof(KInput) + 2 * sizeof(uint64_t) + kScTabEntries * sizeof(double2);
  int rc = 0;
  // the common normal of the uniform-normal variant = the first record's (stored doubled), read by the kernel itself
  const KLaunch L{grid, smem, st, M, N, xM, yM, zM, packed, nps, part, N ? &packed[0].nx2 : nullptr, 0.5};
  if (R == 1)
    rc = launch_main<1, K_THREADS, K_TILE>(strict, uni && N > 0, tabsc, L);
  else if (R == 2)
    rc = launch_main<2, K_THREADS, K_TILE>(strict, uni && N > 0, tabsc, L);
  else
    rc = launch_main<4, K_THREADS, K_TILE>(strict, uni && N > 0, tabsc, L);
  if (rc) return rc;
  kirchhoff_finalize_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(
      M, nsplit, part, convention, reinterpret_cast<double2*>(S), reinterpret_cast<double2*>(P),
      reinterpret_cast<double2*>(A), reinterpret_cast<double2*>(B), reinterpret_cast<double2*>(C));
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// device pointers only; `ws` must hold kirchhoff_workspace_bytes(M, N, R) bytes.
int kirchhoff_device(size_t M, size_t N, const double* xM, const double* yM, const double* zM, const double* nl,
                     const double* Es, const double* Ep, const double* k, const double* pos, int pos_stride,
                     const double* nrm, int nrm_stride, double* S, double* P, double* A, double* B, double* C,
                     int flags, void* ws, cudaStream_t st) {
  if (M == 0) return 0;
  int rc = kirchhoff_pack_device(N, nl, Es, Ep, k, pos, pos_stride, nrm, nrm_stride, ws, st);
  if (rc) return rc;
  return kirchhoff_packed_device(M, N, xM, yM, zM, ws, S, P, A, B, C, flags,
                                 reinterpret_cast<unsigned char*>(ws) + kirchhoff_packed_bytes(N), st);
}

}  // namespace xrtb
