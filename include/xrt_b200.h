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

#ifdef __cplusplus
}
#endif
#endif /* XRT_B200_H */
