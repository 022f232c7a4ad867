/* xrt_b200 -- C ABI of the B200-native (sm_100a) hot path of xrt's `raycing` backend.
 *
 * This header is the drop-in boundary.  Every entry point replaces one OpenCL kernel dispatch that the
 * reference issues through XRT_CL.run_parallel (xrt/backends/raycing/myopencl.py:414-583); the reference
 * file:line each one stands in for is cited next to it.  Conventions:
 *
 *   - plain C types only; complex128 arrays are interleaved (re, im) doubles;
 *   - `*_f64`       : every array pointer is a DEVICE pointer, the call is asynchronous on `stream`
 *                      (a cudaStream_t passed as void*, NULL = default stream);
 *   - `*_f64_host`  : every array pointer is a HOST pointer owned by the caller; the call stages the data
 *                      through device memory, runs the same kernels and returns after the results are back in
 *                      the caller's arrays (blocking, like cl_is_blocking=True, myopencl.py:97);
 *   - return value 0 = success; non-zero = failure, message in xrt_b200_last_error();
 *   - no CPU fallback exists: without a usable CUDA device every call fails.
 */
#ifndef XRT_B200_H
#define XRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ runtime */
int xrt_b200_version(void);
/* Select the CUDA device of this process (one process per GPU). Fails when no device is present.
 * Stands in for XRT_CL.set_cl device discovery (myopencl.py:155-412). */
int xrt_b200_init(int device);
int xrt_b200_device_count(void);
int xrt_b200_sm_count(void);
const char* xrt_b200_last_error(void);
int xrt_b200_sync(void* stream);
/* Measured fp64 ceilings used as roofline denominators (bench.py). mode 0: dependent-free DFMA stream,
 * mode 1: DMMA m8n8k4, mode 2: both interleaved. Writes TFLOP/s (2 flop per FMA). */
int xrt_b200_fp64_probe(int mode, int iters, double* tflops, double* ms);

/* ------------------------------------------------------------------ a17: Fresnel-Kirchhoff sum
 * Replaces kernel `integrate_kirchhoff` (cl/diffract.cl:80-151) launched at waves.py:892.
 *   M outputs (xM,yM,zM); N inputs: nl[N], Es[N] c128, Ep[N] c128, k[N], pos[N*pos_stride] (x,y,z,..),
 *   nrm[N*nrm_stride] (nx,ny,nz,..) -- stride 4 reproduces the reference's float3 buffers (waves.py:874-881).
 *   Outputs S,P,A,B,C: c128[M].
 * flags: bit0 = relaxed arithmetic (allow FMA contraction of |d|^2; default is numpy rounding order)
 *        bit1 = OpenCL sign/scale convention (default numpy convention, waves.py:844-850)
 *        bits4-6 = register blocking R in {1,2,4} (0 = default 2)
 * `workspace` must hold xrt_kirchhoff_workspace_bytes(M, N, flags) bytes of device memory. */
size_t xrt_kirchhoff_workspace_bytes(size_t M, size_t N, int flags);
int xrt_kirchhoff_f64(size_t M, size_t N, const double* xM, const double* yM, const double* zM, const double* nl,
                      const double* Es, const double* Ep, const double* k, const double* pos, int pos_stride,
                      const double* nrm, int nrm_stride, double* S, double* P, double* A, double* B, double* C,
                      int flags, void* workspace, void* stream);
int xrt_kirchhoff_f64_host(size_t M, size_t N, const double* xM, const double* yM, const double* zM,
                           const double* nl, const double* Es, const double* Ep, const double* k, const double* pos,
                           int pos_stride, const double* nrm, int nrm_stride, double* S, double* P, double* A,
                           double* B, double* C, int flags);

/* ------------------------------------------------------------------ a4: undulator far-field integral
 * Replaces kernel `undulator` (cl/undulator.cl:54-121) launched at sources/synchr.py:2174 with the argument
 * list of Undulator._build_I_map_CL (synchr.py:2134-2155): scalars (alpha ignored), Kx, Ky, jend; per-sample
 * gamma, wu, w, ww1, ddphi (theta), ddpsi (psi); node tables tg, ag, sintg, costg, sintgph, costgph[jend];
 * outputs Is, Ip: c128[n] = the return value of Undulator._sp_sum (synchr.py:1929-2033, numpy convention).
 * flags: bits4-6 = register blocking R in {1,2,4} (0 = default 2). */
int xrt_undulator_f64(size_t n, int jend, double alpha, double Kx, double Ky, const double* gamma, const double* wu,
                      const double* w, const double* ww1, const double* ddphi, const double* ddpsi, const double* tg,
                      const double* ag, const double* sintg, const double* costg, const double* sintgph,
                      const double* costgph, double* Is, double* Ip, int flags, void* stream);
int xrt_undulator_f64_host(size_t n, int jend, double alpha, double Kx, double Ky, const double* gamma,
                           const double* wu, const double* w, const double* ww1, const double* ddphi,
                           const double* ddpsi, const double* tg, const double* ag, const double* sintg,
                           const double* costg, const double* sintgph, const double* costgph, double* Is,
                           double* Ip, int flags);
/* a3+a4 fused: Undulator._build_I_map_conv (synchr.py:2048-2108) for device-resident sampling.
 * Inputs w (eV), theta, psi [n]; gamma [n] or NULL (then gamma0 is used for every sample).
 * amp_const = FINE_STR * eI / SIE0; bw_mode = 1 for distE == 'BW' (bwFact = 0.001) else bwFact = 1/w.
 * Outputs I f64[n], Es c128[n], Ep c128[n] exactly as the tuple returned at synchr.py:2106-2108. */
int xrt_undulator_imap_f64(size_t n, int jend, double Kx, double Ky, double L0, double Np, double gamma0,
                           double amp_const, double dstep, int bw_mode, const double* gamma, const double* w,
                           const double* theta, const double* psi, const double* tg, const double* ag,
                           const double* sintg, const double* costg, const double* sintgph, const double* costgph,
                           double* I, double* Es, double* Ep, int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XRT_B200_H */
