// Internal launcher prototypes shared between the .cu translation units of libxrt_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/xrt_b200.h"

namespace xrtb {

// kirchhoff.cu
size_t kirchhoff_workspace_bytes(size_t M, size_t N, int R);
int kirchhoff_plan_splits(size_t M, size_t N, int R);
int kirchhoff_pack_device(size_t N, const double* nl, const double* Es, const double* Ep, const double* k,
                          const double* pos, int pos_stride, const double* nrm, int nrm_stride, void* records,
                          cudaStream_t st);
int kirchhoff_packed_device(size_t M, size_t N, const double* xM, const double* yM, const double* zM,
                            const void* records, double* S, double* P, double* A, double* B, double* C, int flags,
                            void* part_ws, cudaStream_t st);
int kirchhoff_device(size_t M, size_t N, const double* xM, const double* yM, const double* zM, const double* nl,
                     const double* Es, const double* Ep, const double* k, const double* pos, int pos_stride,
                     const double* nrm, int nrm_stride, double* S, double* P, double* A, double* B, double* C,
                     int flags, void* ws, cudaStream_t st);

// undulator.cu
const double2* sin_table_device(cudaStream_t st);  // (cos, sin)(k pi/512) for sin_tab (common.cuh); NULL on failure
int undulator_raw_device(size_t n, int jend, double Kx, double Ky, const double* gamma, const double* wu,
                         const double* w, const double* ww1, const double* ddphi, const double* ddpsi,
                         const double* tg, const double* ag, const double* sintg, const double* costg,
                         const double* sintgph, const double* costgph, double* Is, double* Ip, int R,
                         int symmetric, cudaStream_t st);
int undulator_imap_device(size_t n, int jend, double Kx, double Ky, double L0, double Np, double gamma0,
                          double amp_const, double dstep, int bw_mode, const double* gamma, const double* w,
                          const double* theta, const double* psi, const double* tg, const double* ag,
                          const double* sintg, const double* costg, const double* sintgph, const double* costgph,
                          double* I, double* Es, double* Ep, int R, int symmetric, cudaStream_t st);
// undulator_ext.cu: variant 1 = taper (scalar0 = _taperVal), 2 = near field (scalar0 = R0*2pi/L0);
// mode 0 = raw (Is, Ip) of _sp_sum, 1 = fused _build_I_map_conv (I, Es, Ep)
int undulator_ext_device(int variant, int mode, size_t n, int jend, int nmax, double scalar0, double Kx, double Ky,
                         double L0, double gamma0, double amp_const, double dstep, int bw_mode, const double* gamma,
                         const double* wu, const double* w, const double* ww1, const double* ddphi,
                         const double* ddpsi, const double* tg, const double* ag, const double* sintg,
                         const double* costg, const double* sintgph, const double* costgph, double* Is, double* Ip,
                         double* I, cudaStream_t st);

// reflect.cu (compiled with -fmad=false)
int reflect_device(size_t n, const xrt_oe_params* oe, const xrt_material_params* mat, const double* tables,
                   size_t table_len, const xrt_beam* in, const xrt_beam* fallback, const xrt_beam* gb,
                   const xrt_beam* lb, unsigned long long* red, int phase, unsigned long long index_offset,
                   cudaStream_t st);
int reflect_expose_device(size_t n, const xrt_oe_params* oe, const xrt_material_params* mat, const double* tables,
                          size_t table_len, const xrt_beam* in, const xrt_beam* fallback, const xrt_beam* gb,
                          const xrt_beam* lb, unsigned long long* red, const xrt_plane_params* scr_plane,
                          const xrt_beam* scr, unsigned long long* scr_red, int phase,
                          unsigned long long index_offset, cudaStream_t st);
int reflect_multi_deriv_device(size_t n, const xrt_oe_params* oe, const xrt_beam* in, unsigned long long* red,
                               double* tD, int phase, cudaStream_t st);
int reflect_multi_elevation_device(size_t n, const xrt_oe_params* oe, const xrt_beam* in, const double* tD, double* eD,
                                   double* eX, double* eY, double* eZ, cudaStream_t st);
double reduction_decode(unsigned long long slot);
int surface_device(size_t n, const xrt_oe_params* oe, const double* x, const double* y, double* z, cudaStream_t st);
int normals_device(size_t n, const xrt_oe_params* oe, const double* x, const double* y, const double* z, double* nx,
                   double* ny, double* nz, cudaStream_t st);
int multilayer_device(size_t n, int npairs, int uniform, const double* rbs_s, const double* rbs_p,
                      const double* rtb_s, const double* rtb_p, const double* rvt_s, const double* rvt_p,
                      const double* Qt, const double* Qb, const double* dti, const double* dbi, double* ri_s,
                      double* ri_p, cudaStream_t st);
int multilayer_tran_device(size_t n, int npairs, double substThickness, const double* const* in23, const double* dti,
                           const double* dbi, double* ti_s, double* ti_p, cudaStream_t st);
int amplitude_device(size_t n, const xrt_material_params* mat, const double* tables, size_t table_len,
                     const double* E, const double* bIn, const double* bOut, const double* bH, double* rs, double* rp,
                     unsigned long long* oor_count, cudaStream_t st);

// rays.cu (compiled with -fmad=false)
int expose_device(size_t n, const xrt_plane_params* p, const xrt_beam* in, const xrt_beam* out,
                  unsigned long long* red, cudaStream_t st);
int expose_hemispheric_device(size_t n, const xrt_plane_params* p, double R, double thetaOffset, double phiOffset,
                              const xrt_beam* in, const xrt_beam* out, cudaStream_t st);
int propagate_rect_device(size_t n, const xrt_plane_params* p, const xrt_beam* in, const xrt_beam* out,
                          unsigned long long* red, cudaStream_t st);

int custom_field_device(size_t n, int jend, double betam, double R0, const double* gamma, const double* w,
                        const double* ddtheta, const double* ddpsi, const double* tg, const double* ag,
                        const double* Bx, const double* By, const double* Bz, const double* betax,
                        const double* betay, const double* trajx, const double* trajy, const double* trajz,
                        double* Is, double* Ip, cudaStream_t st);

// source.cu (compiled with -fmad=false)
int source_sample_device(size_t n, uint64_t seed, uint64_t index_offset, const double* window, double* E,
                         double* theta, double* psi, double* n1, double* n2, double* n3, double* n4, double* u,
                         double* n5, cudaStream_t st);
int source_fill_device(size_t n, const xrt_source_params* p, const long long* idx, const double* rE,
                       const double* rTheta, const double* rPsi, const double* mJs, const double* mJp,
                       const double* n1, const double* n2, const double* n3, const double* n4, const xrt_beam* out,
                       size_t out_offset, cudaStream_t st);

int bm_imap_device(size_t n, const xrt_bm_params* bm, const double* gamma, const double* w, const double* theta,
                   const double* psi, double* I, double* Es, double* Ep, cudaStream_t st);
int bm_fill_device(size_t n, const xrt_source_params* p, const xrt_bm_params* bm, const long long* idx,
                   const double* rE, const double* rTheta, const double* rPsi, const double* mJs, const double* mJp,
                   const double* g1, const double* g2, const double* g3, const double* g4, const double* g5,
                   const double* kint, const xrt_beam* out, size_t out_offset, cudaStream_t st);
int bm_shine_uniform_device(size_t n, uint64_t seed, uint64_t index_offset, const double* window,
                            const xrt_source_params* sp, const xrt_bm_params* bm, const xrt_beam* out,
                            unsigned long long* imax_slot, cudaStream_t st);

int geometric_source_device(size_t n, const xrt_geom_params* p, uint64_t seed, uint64_t index_offset,
                            const double* const* deviates, const xrt_beam* out, unsigned long long* flag, int phase,
                            cudaStream_t st);

size_t compact_scratch_words(size_t n);
int compact_rays_device(size_t n, const int* state, int state_mask, int narr, const double* const* src,
                        double* const* dst, const int* words, unsigned long long* scratch, cudaStream_t st);
int plot_hist_device(size_t n, const xrt_plot_params* p, cudaStream_t st);
int hist2d_device(size_t n, const double* x, const double* y, const double* w1, const double* w2, const int* state,
                  int state_mask, const double* ex, int nx, const double* ey, int ny, double* out, cudaStream_t st);
int shine_uniform_device(size_t n, uint64_t seed, uint64_t index_offset, const double* window, int jend, double Kx,
                         double Ky, double L0, double Np, double gamma0, double amp_const, double dstep, int bw_mode,
                         int symmetric, const double* tg, const double* ag, const double* sintg, const double* costg,
                         const double* sintgph, const double* costgph, const xrt_source_params* sp,
                         const xrt_beam* out, unsigned long long* imax_slot, cudaStream_t st);

}  // namespace xrtb
