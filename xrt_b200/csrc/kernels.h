// Internal launcher prototypes shared between the .cu translation units of libxrt_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace xrtb {

// kirchhoff.cu
size_t kirchhoff_workspace_bytes(size_t M, size_t N, int R);
int kirchhoff_plan_splits(size_t M, size_t N, int R);
int kirchhoff_device(size_t M, size_t N, const double* xM, const double* yM, const double* zM, const double* nl,
                     const double* Es, const double* Ep, const double* k, const double* pos, int pos_stride,
                     const double* nrm, int nrm_stride, double* S, double* P, double* A, double* B, double* C,
                     int flags, void* ws, cudaStream_t st);

}  // namespace xrtb
