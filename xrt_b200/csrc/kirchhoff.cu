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
  double w_re, w_im, k2x, nq2;  // Es + Ep, 2 k, 2 n.q (uniform-normal fast path)
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
  r.w_re = es.x + ep.x;  // numpy forms (Es + Ep) first as well (waves.py:847)
  r.w_im = es.y + ep.y;
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
template <int R, int THREADS, int TILE, bool STRICT, bool UNIFORM_N>
__global__ void __launch_bounds__(THREADS, XRTB_KIRCHHOFF_MINBLOCKS)
    kirchhoff_kernel(size_t M, size_t N, const double* __restrict__ xM, const double* __restrict__ yM,
                     const double* __restrict__ zM, const KInput* __restrict__ in, size_t n_per_split,
                     double2* __restrict__ part, double un_x, double un_y, double un_z) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  KInput* tile = reinterpret_cast<KInput*>(smem_raw);                       // [2][TILE]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + 2 * TILE * sizeof(KInput));  // [2]

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
        // h = (k^2 / r) * g * (Es + Ep)
        const double kk = kip * a1.w;  // k^2 / r
        const double wre = fma(gre, a3.x, -gim * a3.y);
        const double wim = fma(gre, a3.y, gim * a3.x);
        const double hre = kk * wre;
        const double him = kk * wim;
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
template <int R, int THREADS, int TILE, bool STRICT, bool UNI>
static int launch_one(dim3 grid, size_t smem, cudaStream_t st, size_t M, size_t N, const double* xM, const double* yM,
                      const double* zM, const KInput* in, size_t nps, double2* part, const double* un) {
  XRTB_CUDA_CHECK(cudaFuncSetAttribute(kirchhoff_kernel<R, THREADS, TILE, STRICT, UNI>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kirchhoff_kernel<R, THREADS, TILE, STRICT, UNI><<<grid, THREADS, smem, st>>>(M, N, xM, yM, zM, in, nps, part, un[0],
                                                                              un[1], un[2]);
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <int R, int THREADS, int TILE>
static int launch_main(bool strict, bool uni, dim3 grid, size_t smem, cudaStream_t st, size_t M, size_t N,
                       const double* xM, const double* yM, const double* zM, const KInput* in, size_t nps,
                       double2* part, const double* un) {
  if (strict)
    return uni ? launch_one<R, THREADS, TILE, true, true>(grid, smem, st, M, N, xM, yM, zM, in, nps, part, un)
               : launch_one<R, THREADS, TILE, true, false>(grid, smem, st, M, N, xM, yM, zM, in, nps, part, un);
  return uni ? launch_one<R, THREADS, TILE, false, true>(grid, smem, st, M, N, xM, yM, zM, in, nps, part, un)
             : launch_one<R, THREADS, TILE, false, false>(grid, smem, st, M, N, xM, yM, zM, in, nps, part, un);
}

#ifndef XRTB_KIRCHHOFF_THREADS
#define XRTB_KIRCHHOFF_THREADS 128
#endif
constexpr int K_THREADS = XRTB_KIRCHHOFF_THREADS;
constexpr int K_TILE = 128;

size_t kirchhoff_packed_bytes(size_t N) { return N * sizeof(KInput); }

// number of N-splits used for a given problem (also sizes the workspace)
int kirchhoff_plan_splits(size_t M, size_t N, int R) {
  const size_t per_cta = (size_t)K_THREADS * R;
  size_t mblocks = (M + per_cta - 1) / per_cta;
  size_t target = (size_t)num_sms() * 4;  // >= 4 CTAs per SM keeps every SMSP busy and bounds the tail
  if (mblocks >= target) return 1;
  size_t ns = (target + mblocks - 1) / mblocks;
  size_t max_ns = (N + K_TILE * 4 - 1) / (K_TILE * 4);  // at least 4 tiles per split
  if (max_ns < 1) max_ns = 1;
  if (ns > max_ns) ns = max_ns;
  if (ns > 4096) ns = 4096;
  return (int)ns;
}

size_t kirchhoff_workspace_bytes(size_t M, size_t N, int R) {
  int ns = kirchhoff_plan_splits(M, N, R);
  return kirchhoff_packed_bytes(N) + (size_t)ns * 5 * M * sizeof(double2);
}

// device pointers only; `ws` must hold kirchhoff_workspace_bytes(M, N, R) bytes.
int kirchhoff_device(size_t M, size_t N, const double* xM, const double* yM, const double* zM, const double* nl,
                     const double* Es, const double* Ep, const double* k, const double* pos, int pos_stride,
                     const double* nrm, int nrm_stride, double* S, double* P, double* A, double* B, double* C,
                     int flags, void* ws, cudaStream_t st) {
  if (M == 0) return 0;
  const bool strict = !(flags & 1);
  const int convention = (flags >> 1) & 1;
  const bool uni = (flags >> 9) & 1;  // caller guarantees that all N normals are equal
  int R = (flags >> 4) & 7;
  if (R == 0) R = 2;
  XRTB_REQUIRE(R == 1 || R == 2 || R == 4, "kirchhoff: register blocking R must be 1, 2 or 4 (got %d)", R);
  KInput* packed = reinterpret_cast<KInput*>(ws);
  double2* part = reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(ws) + kirchhoff_packed_bytes(N));
  const int nsplit = kirchhoff_plan_splits(M, N, R);
  if (N > 0) {
    kirchhoff_pack_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(
        N, nl, reinterpret_cast<const double2*>(Es), reinterpret_cast<const double2*>(Ep), k, pos, pos_stride, nrm,
        nrm_stride, packed);
    XRTB_CUDA_CHECK(cudaGetLastError());
  }
  const size_t per_cta = (size_t)K_THREADS * R;
  dim3 grid((unsigned)((M + per_cta - 1) / per_cta), (unsigned)nsplit);
  size_t nps = (N + nsplit - 1) / nsplit;
  nps = (nps + K_TILE - 1) / K_TILE * K_TILE;
  if (nps == 0) nps = K_TILE;
  const size_t smem = 2 * K_TILE * sizeof(KInput) + 2 * sizeof(uint64_t);
  int rc = 0;
  double un[3] = {0.0, 0.0, 0.0};
  if (uni && N > 0) {  // the common normal = the first sample's (3 doubles, device -> host)
    XRTB_CUDA_CHECK(cudaMemcpyAsync(un, nrm, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    XRTB_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  if (R == 1)
    rc = launch_main<1, K_THREADS, K_TILE>(strict, uni, grid, smem, st, M, N, xM, yM, zM, packed, nps, part, un);
  else if (R == 2)
    rc = launch_main<2, K_THREADS, K_TILE>(strict, uni, grid, smem, st, M, N, xM, yM, zM, packed, nps, part, un);
  else
    rc = launch_main<4, K_THREADS, K_TILE>(strict, uni, grid, smem, st, M, N, xM, yM, zM, packed, nps, part, un);
  if (rc) return rc;
  kirchhoff_finalize_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(
      M, nsplit, part, convention, reinterpret_cast<double2*>(S), reinterpret_cast<double2*>(P),
      reinterpret_cast<double2*>(A), reinterpret_cast<double2*>(B), reinterpret_cast<double2*>(C));
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace xrtb
