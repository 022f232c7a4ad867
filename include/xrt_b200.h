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
/* sizeof() of the ABI structs below (0: xrt_beam, 1: xrt_oe_params, 2: xrt_material_params, 3: xrt_plane_params, 4: xrt_source_params,
 * 5: xrt_plot_params, 6: xrt_multilayer_params, 7: xrt_bm_params, 8: xrt_geom_params), for binding checks */
size_t xrt_b200_sizeof(int which);
/* Measured fp64 ceilings used as roofline denominators (bench.py). mode 0: dependent-free DFMA stream,
 * mode 1: DMMA m8n8k4, mode 2: both interleaved. Writes TFLOP/s (2 flop per FMA). */
int xrt_b200_fp64_probe(int mode, int iters, double* tflops, double* ms);

/* Known-answer access to the device math helpers the kernels are built from (host arrays; tests only):
 * which 0 sincos_big -> (sin, cos); 1 sin_big; 2 rcp_fast; 3 sqrt_rsqrt -> (sqrt, 1/sqrt); 4 (K_1/3, K_2/3);
 * 6 rcp_cubic; 7 sqrt_hrsqrt -> (sqrt, 0.5/sqrt); 8 sincos_tab -> (sin, cos); 9 sin_tab; 10 the unchecked IEEE
 * sequences div_nr / sqrt_nr -> (1/x, sqrt|x|); 11 div_nr(x[i], x[i^1]) (n even). */
int xrt_b200_math_probe_host(int which, size_t n, const double* x, double* y0, double* y1);

/* ------------------------------------------------------------------ a17: Fresnel-Kirchhoff sum
 * Replaces kernel `integrate_kirchhoff` (cl/diffract.cl:80-151) launched at waves.py:892.
 *   M outputs (xM,yM,zM); N inputs: nl[N], Es[N] c128, Ep[N] c128, k[N], pos[N*pos_stride] (x,y,z,..),
 *   nrm[N*nrm_stride] (nx,ny,nz,..) -- stride 4 reproduces the reference's float3 buffers (waves.py:874-881).
 *   Outputs S,P,A,B,C: c128[M].
 * flags: bit0 = relaxed arithmetic (allow FMA contraction of |d|^2; default is numpy rounding order)
 *        bit1 = OpenCL sign/scale convention (default numpy convention, waves.py:844-850)
 *        bits4-6 = register blocking R in {1,2,4} (0 = default 2)
 *        bit9 = all N normals are equal (aperture, flat OE): n.d = n.p - n.q with n.p hoisted per output
 *        bit11 = sin / cos of the phase through a 4096-entry (pi/2048) table in shared memory + two-term polynomials
 *                instead of the two pi/2-range polynomials (same <= 4e-16 accuracy, 16 % faster).  The caller
 *                vouches for |k r| < 2^40 pi rad (3.4e12: 100 keV over 6.8 m, 9 keV over 75 m) for every pair; the
 *                kernel does not test
 * `workspace` must hold xrt_kirchhoff_workspace_bytes(M, N, flags) bytes of device memory. */
size_t xrt_kirchhoff_workspace_bytes(size_t M, size_t N, int flags);
int xrt_kirchhoff_f64(size_t M, size_t N, const double* xM, const double* yM, const double* zM, const double* nl,
                      const double* Es, const double* Ep, const double* k, const double* pos, int pos_stride,
                      const double* nrm, int nrm_stride, double* S, double* P, double* A, double* B, double* C,
                      int flags, void* workspace, void* stream);
/* The same sum in two steps, for samples that are ray shards of several ranks (shard the M receiving
 * points, all-gather the N samples): xrt_kirchhoff_pack_f64 turns this rank's samples into 128-byte records
 * (`records`: N * 128 bytes of device memory), the host all-gathers the records (xrt_allgather_bytes), and
 * xrt_kirchhoff_packed_f64 sums over all of them.  workspace: xrt_kirchhoff_workspace_bytes(M, N, flags) bytes is
 * enough (it only holds the split partial sums here).  With bit9 the common normal is the first record's. */
#define XRT_KIRCHHOFF_RECORD_BYTES 128
int xrt_kirchhoff_pack_f64(size_t N, const double* nl, const double* Es, const double* Ep, const double* k,
                           const double* pos, int pos_stride, const double* nrm, int nrm_stride, void* records,
                           void* stream);
int xrt_kirchhoff_packed_f64(size_t M, size_t N, const double* xM, const double* yM, const double* zM,
                             const void* records, double* S, double* P, double* A, double* B, double* C, int flags,
                             void* workspace, void* stream);
int xrt_kirchhoff_f64_host(size_t M, size_t N, const double* xM, const double* yM, const double* zM,
                           const double* nl, const double* Es, const double* Ep, const double* k, const double* pos,
                           int pos_stride, const double* nrm, int nrm_stride, double* S, double* P, double* A,
                           double* B, double* C, int flags);

/* ------------------------------------------------------------------ a4: undulator far-field integral
 * Replaces kernel `undulator` (cl/undulator.cl:54-121) launched at sources/synchr.py:2174 with the argument
 * list of Undulator._build_I_map_CL (synchr.py:2134-2155): scalars (alpha ignored), Kx, Ky, jend; per-sample
 * gamma, wu, w, ww1, ddphi (theta), ddpsi (psi); node tables tg, ag, sintg, costg, sintgph, costgph[jend];
 * outputs Is, Ip: c128[n] = the return value of Undulator._sp_sum (synchr.py:1929-2033, numpy convention).
 * flags: bits4-6 = register blocking R in {1,2,4} (0 = default 2);
 *        bit10 = the node tables are mirror-symmetric about tg = 0 (Clenshaw-Curtis grids are, to a few ulps):
 *                a planar undulator (Kx == 0) then visits node PAIRS, 2 i w [..] sin(u) per pair (device-pointer
 *                entry points trust the caller; the _host entry point checks the tables itself, bit11 opts out). */
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

/* f3: tapered and near-field undulators.  Replace kernels `undulator_taper` (cl/undulator.cl:123-202) and
 * `undulator_nf` (cl/undulator.cl:204-298), launched at sources/synchr.py:2174 with the scalar list
 * [taper | R0*pi*2/L0, Kx, Ky, jend, Np] (synchr.py:2134-2146) and the sliced / non-sliced arrays of `undulator`.
 * variant 1: taper, scalar0 = Undulator._taperVal; variant 2: near field, scalar0 = R0*pi*2/L0.  nmax = Np periods,
 * summed explicitly.  Outputs Is, Ip c128[n] = Undulator._sp_sum(..., R0) (synchr.py:1929-2033, numpy convention). */
int xrt_undulator_ext_f64(int variant, size_t n, int jend, int nmax, double scalar0, double Kx, double Ky,
                          const double* gamma, const double* wu, const double* w, const double* ww1,
                          const double* ddphi, const double* ddpsi, const double* tg, const double* ag,
                          const double* sintg, const double* costg, const double* sintgph, const double* costgph,
                          double* Is, double* Ip, void* stream);
int xrt_undulator_ext_f64_host(int variant, size_t n, int jend, int nmax, double scalar0, double Kx, double Ky,
                               const double* gamma, const double* wu, const double* w, const double* ww1,
                               const double* ddphi, const double* ddpsi, const double* tg, const double* ag,
                               const double* sintg, const double* costg, const double* sintgph,
                               const double* costgph, double* Is, double* Ip);
/* ... and fused with Undulator._build_I_map_conv (synchr.py:2048-2108; ab = 1/(2 pi wu), :2067-2068) for
 * device-resident sampling: same arguments as xrt_undulator_imap_f64 plus variant, nmax = Np and scalar0. */
int xrt_undulator_ext_imap_f64(int variant, size_t n, int jend, int nmax, double scalar0, double Kx, double Ky,
                               double L0, double gamma0, double amp_const, double dstep, int bw_mode,
                               const double* gamma, const double* w, const double* theta, const double* psi,
                               const double* tg, const double* ag, const double* sintg, const double* costg,
                               const double* sintgph, const double* costgph, double* I, double* Es, double* Ep,
                               void* stream);

/* f3: radiation integral along the trajectory in a tabulated field.  Replaces kernel `custom_field`
 * (cl/undulator.cl) launched at sources/synchr.py:1259 with scalars [jend, betam, R0], sliced [gamma, w, ddtheta,
 * ddpsi], node tables [tg, ag, Bx, By, Bz, betax, betay, trajx, trajy, trajz] (synchr.py:1227-1257).  Outputs Is, Ip
 * c128[n] = SourceFromField._sp_sum (synchr.py:888-973, not filamentBeam); R0 = 0: far field. */
int xrt_custom_field_f64(size_t n, int jend, double betam, double R0, const double* gamma, const double* w,
                         const double* ddtheta, const double* ddpsi, const double* tg, const double* ag,
                         const double* Bx, const double* By, const double* Bz, const double* betax,
                         const double* betay, const double* trajx, const double* trajy, const double* trajz,
                         double* Is, double* Ip, void* stream);
int xrt_custom_field_f64_host(size_t n, int jend, double betam, double R0, const double* gamma, const double* w,
                              const double* ddtheta, const double* ddpsi, const double* tg, const double* ag,
                              const double* Bx, const double* By, const double* Bz, const double* betax,
                              const double* betay, const double* trajx, const double* trajy, const double* trajz,
                              double* Is, double* Ip);

/* ------------------------------------------------------------------ a1, a6-a15: ray kernels on the Beam SoA
 * The Beam container of the reference (sources/beams.py:96-186) is a struct of equal-length arrays; here it
 * is a struct of DEVICE pointers.  Es/Ep may be NULL (beam without amplitudes); theta, s, phi, r are outputs
 * of reflect() for the local beam only and may be NULL. */
typedef struct xrt_beam {
  double *x, *y, *z, *a, *b, *c, *path, *E, *Jss, *Jpp;
  double *Jsp; /* complex128, interleaved */
  double *Es, *Ep; /* complex128 or NULL */
  int32_t* state;
  double *theta, *s, *phi, *r; /* optional outputs (lb) */
} xrt_beam;

#define XRT_MAX_ROT 8
#define XRT_MAX_ELEM 4

/* Geometry of one optical surface = what OE.reflect / DCM.double_reflect read from `self`
 * (oes/reflect.py:104-135, :551-640; oes/dcm.py:273-335; oes/base.py:1094-1163, :1231-1295). */
typedef struct xrt_oe_params {
  double center[3];
  double sinAz, cosAz;    /* BeamLine azimuth (beamline.py:241-250); skipped when from_global == 0 */
  int from_global;        /* 1: input is global (reflect, 1st crystal); 0: input is virgin local (2nd crystal) */
  int to_global;          /* 1: gb leaves in global coordinates; 0: gb stays virgin local (1st crystal of a DCM) */
  int n_rot_in;           /* elementary rotations global->local in application order (_rotate.py:23-57) */
  int rot_in_axis[XRT_MAX_ROT]; /* 0 = Rx(pitch), 1 = Ry(roll), 2 = Rz(yaw) */
  double rot_in_cos[XRT_MAX_ROT], rot_in_sin[XRT_MAX_ROT];
  int n_rot_out;
  int rot_out_axis[XRT_MAX_ROT];
  double rot_out_cos[XRT_MAX_ROT], rot_out_sin[XRT_MAX_ROT];
  double shift[3];        /* dx, dy, dz (reflect.py:630-635) */
  double roll;            /* roll used for the polarisation frame, reflect.py:948 */
  int surface;            /* 0 flat, 1 toroid (R, r), 2 elliptical cylinder/ellipsoid, parametric,
                             3 paraboloid / parabolic cylinder lens surface (focus, zmax, has_zmax, isCylinder) */
  double sp[8];           /* surface parameters, cl_plist order of the reference classes */
  int asymmetric;         /* local_n returns (n_B, n_s): crystal with alpha != 0 */
  double nB[3];           /* n_B for a flat asymmetric crystal (base.py:736-739; dcm.py:234-238) */
  double physX[2], physY[2], optX[2], optY[2];
  int has_optX, has_optY;
  int over_mask;          /* bit0 xmin, bit1 xmax, bit2 ymin, bit3 ymax (overEdge, base.py:1127-1136) */
  int lostNum;
  double brk_lim[3][2];   /* _set_t limits per axis after the +-inf -> defSize replacement (base.py:1233-1239) */
  int created_by_diffract; /* goodAfter = (state == 1) only (reflect.py:125-128) */
  int zero_dead_local;    /* DCM 2nd crystal: lo2 of rays that never reached it is zeroed (dcm.py:305-310) */
  int good_mode;          /* which rays are traced: 0: state > 0 (reflect.py:110); 1: state in (1, 2) (dcm.py:303) */
  int want_footprint;     /* 1: reduce the footprint min/max and the good-ray / out-of-table counters into the
                             slots (reflect.py:1133-1136); 0 skips those block reductions (device pipelines) */
  int no_intersection;    /* noIntersectionSearch (reflect.py:676-681): the rays already sit on the surface (wave
                             samples): t = 0, the point is taken as it is (xyz_to_param for parametric OEs) */
  /* f4: refraction (materials of kind 'plate' / 'lens', toWhere = 1, reflect.py:894-919; oes/refractive.py) */
  int from_vacuum;        /* 1: the ray enters the material (invertNormal = 1); 0: it leaves it (invertNormal = -1,
                             attenuation and phase inside the material, reflect.py:638-641, :1048-1059) */
  int dead_state_lost;    /* Plate (dcm.py:311-312, :331-332): rays that are not good after a surface get lostNum */
  /* f4: grating (material kind 'grating', toWhere = 3, reflect.py:840-861 + _grating_deflection :451-469).
   * OE.local_g (base.py:688-717): 0 no grating; 1 / 2: groove density rho0 * sum_i (i+1) c_i coord^i along x / y
   * (gratingDensity = ['x'|'y', rho0, c0, c1, ...]); 3: constant (0, g_rho0, 0); 4: Fresnel zone plate (fzp_* below) */
  int grating;
  int g_ncoef;
  double g_rho0, g_coef[8];
  double g_order;         /* diffraction order (a single int in the reference's OE.order) */
  double g_amp;           /* >= 0: amplitude sqrt(efficiency) of that order (material.py:391-413); < 0: none */
  /* OE shape 'round' (rays_good, base.py:1140-1156): an ellipse in limPhysX x limPhysY; shape_r = centerX, radiusX,
   * centerY, radiusY; rays outside are lost (no 'out' / 'over' states) */
  int shape_round;
  double shape_r[4];
  /* f4: NormalFZP (oes/gratings.py:10-137; grating = 4, material kind 7): fzp_rn[fzp_n] zone radii r_0 = 0 .. r_N
   * (device pointer), fzp_black = int(isCentralZoneBlack): rays whose zone index i has i % 2 != fzp_black are lost,
   * the others are deflected by the local radial groove vector with giveSign = +1 (reflect.py:840-861) */
  const double* fzp_rn;
  int fzp_n, fzp_black;
  /* f4: OE.multiple_reflect (oes/reflect.py:165-264) */
  int multi;              /* 1: one pass of multiple_reflect -- gb receives EVERY traced ray in virgin-local coordinates
                             with its new state (gb and lb are one beam there), rays over the edge keep their entry
                             point (:224-227), gb.theta is written when present; input is virgin local */
  /* f4: figure error (OE(figureError=...), figure_error.py:207-265): height map in nm as the bicubic B-spline of scipy's
   * RectBivariateSpline(y1d, x1d, z): knots fe_tu[fe_nu] along y, fe_tv[fe_nv] along x, coefficients
   * fe_c[(fe_nu - 4) * (fe_nv - 4)] (device pointers), evaluated at (y + fe_ushift, x + fe_vshift).  NULL fe_c: none.
   * Served by figure-error plug-in variants only (xrt_b200/custom.py); libxrt_b200.so refuses a non-NULL fe_c. */
  const double *fe_tu, *fe_tv, *fe_c;
  int fe_nu, fe_nv;
  double fe_ushift, fe_vshift;
  const double* multi_tmin; /* passes after the first: per-ray lower bracket end = the derivative-stage root
                             (xrt_reflect_multi_deriv_f64; base.py:1279-1289); NULL: the usual _set_t bracket */
} xrt_oe_params;

/* Material = what get_amplitude reads (materials/material.py:348-493, materials/crystal.py:492-645). */
struct xrt_multilayer_params;
typedef struct xrt_material_params {
  int kind;               /* 0 none, 1 mirror, 2 thin mirror, 3 crystal (fcc/diamond, one element; thick Bragg or thickness t with geom),
                             4 multilayer (reflected; `multilayer` points to the stack),
                             5 plate / lens (refraction + transmittivity, material.py:482-486),
                             6 grating with a mirror coating, 7 grating without reflectivity or with tabulated
                             efficiencies (xrt_oe_params.grating != 0 goes with 6 / 7 and only with them) */
  int nelem;
  double Z[XRT_MAX_ELEM], quantity[XRT_MAX_ELEM];
  int tab_off[XRT_MAX_ELEM], tab_len[XRT_MAX_ELEM]; /* slices of the E/f1/f2 tables */
  double rho, mass, t;
  double d, chiToF, factDW;
  double f0c[11];
  double d2f_re, d2f_im;  /* diamondToFcc = 1 + exp(i pi/2 (h+k+l)); (1, 0) for fcc */
  double f0scale;         /* 2 for diamond (F0 * 2), 1 for fcc */
  int hkl_allowed;        /* h,k,l all even or all odd */
  int geom;               /* crystal: bit0 Laue (else Bragg), bit1 transmitted (else reflected); with t > 0 (mm) the
                             finite-thickness amplitudes of crystal.py:598-616, t == 0: thick Bragg */
  const struct xrt_multilayer_params* multilayer; /* kind 4: HOST pointer to the layer stack, else NULL */
} xrt_material_params;

/* Multilayer (materials/multilayer.py:14-515), geom 'reflected': top layer / bottom layer / substrate materials
 * (their element tables are slices of the same `tables` array the call receives), nPairs periods with the
 * per-period thicknesses dti[nPairs], dbi[nPairs] in Angstrom (set_dti / set_dbi: constant or depth graded),
 * interdiffusion rms.  dti/dbi live where the call's arrays live (device for xrt_reflect_f64, host for the
 * *_host entry points, which stage them). */
typedef struct xrt_multilayer_params {
  xrt_material_params tLayer, bLayer, substrate;  /* kind 1; only the refractive index is used */
  int nPairs;
  int uniform;            /* 1: every period has dti[0], dbi[0] -- the two phase factors are taken once per ray */
  double idThickness;     /* interdiffusion / roughness rms, Angstrom */
  double d;               /* tThickness + bThickness: the period (grating deflection, Bragg angle) */
  const double *dti, *dbi;
  /* Coated (materials/multilayer.py:528-584): tLayer.nelem == 0 means no top layer (vacuum, n = 1); the bottom-layer /
   * substrate interface then takes substRoughness instead of idThickness (multilayer.py:401) */
  double substRoughness;
  int transmitted;        /* 1: geom 'transmitted' (multilayer.py:388-516): the amplitude transmitted through the stack and
                             its substrate of thickness substThickness (Angstrom); the rays pass straight */
  double substThickness;
  int specular;           /* 1: kind 'mirror' -- specular reflection with this amplitude instead of the multilayer's
                             grating-like deflection (reflect.py:1003-1025 vs :999-1002) */
} xrt_multilayer_params;

/* Beam-wide scalars numpy takes over the whole beam. 64 slots of 8 bytes, device
 * memory.  Written by the pre-pass, read by the main pass; all-reduced (min/max) across ranks in between
 * when rays are sharded. Layout documented in xrt_b200/csrc/reflect.cu. */
#define XRT_REDUCTION_SLOTS 64

/* OE.reflect / one _reflect_local pass.  tables: [E | f1 | f2] each `table_len` doubles (float32 data promoted).
 * phase bit0: pre-pass A (dominant direction, first good ray), bit2: pre-pass B (global bracket, solver
 * switch), bit1: main pass; 7 = everything.  Ray shards all-reduce the slots after A and after B.
 * lb may be NULL (needLocal=False).
 * index_offset: global index of ray 0 of this shard (first-good-ray bookkeeping of _set_t, base.py:1241).
 * fallback: beam whose values non-intersected rays receive in gb (reflect.py:131-134; the ORIGINAL beam for the
 * 2nd crystal of a DCM, dcm.py:343-345); NULL = `in`.
 * f0c[0] of the material holds f0(sin(theta)/lambda = 0.5/d) evaluated on the host (element.py:203-207). */
int xrt_reflect_f64(size_t n, const xrt_oe_params* oe, const xrt_material_params* mat, const double* tables,
                    size_t table_len, const xrt_beam* in, const xrt_beam* fallback, const xrt_beam* gb,
                    const xrt_beam* lb, uint64_t* reductions, int phase, uint64_t index_offset, void* stream);
/* f4: derivative stage of OE.multiple_reflect -- `_bracketing(isMulti=True)` (oes/base.py:1279-1289) calls
 * find_intersection(local_n, tMin = 0, tMax, ..., derivOrder = 1) (base.py:848-885 with find_dz :818-820, :842-845):
 * the root of (ray direction . surface normal) along each ray, i.e. where the ray is farthest from the surface it has
 * just left.  Run after pre-pass A of xrt_reflect_f64 (phase 1) on the same `reductions`:
 *   phase 1: beam-wide scalars of that search (t2.max(), secant-vs-Brent switch) -> slots 40..43;
 *   phase 2: per-ray roots -> tD[n] (rays that are not traced keep their tD untouched).
 * Then xrt_reflect_f64 phases 4 | 2 with xrt_oe_params.multi_tmin = tD search the next intersection. */
int xrt_reflect_multi_deriv_f64(size_t n, const xrt_oe_params* oe, const xrt_beam* in, uint64_t* reductions,
                                double* tD, int phase, void* stream);
/* f4: OE.multiple_reflect(needElevationMap=True) (oes/reflect.py:212-216, :651-659; oes/base.py:1285-1287): after
 * phase 2 of xrt_reflect_multi_deriv_f64, find_dz(None, tMax = tD, ...) of every traced ray -- elevationD[n] the height
 * of the ray above the surface where it is farthest from it, elevationX/Y/Z[n] that point in local Cartesian
 * coordinates.  Rays that are not traced keep their entries (the reference initialises them to -1 and
 * -maxHalfSizeOfOE). */
int xrt_reflect_multi_elevation_f64(size_t n, const xrt_oe_params* oe, const xrt_beam* in, const double* tD,
                                    double* elevationD, double* elevationX, double* elevationY, double* elevationZ,
                                    void* stream);
/* decode one reduction slot to double (host helper; slots are order-preserving integer encodings) */
double xrt_reduction_decode(uint64_t slot);

/* ------------------------------------------------------------------ a15: plane elements
 * Screen.expose (screens.py:225-302) and RectangularAperture.propagate (apertures.py:333-413). */
typedef struct xrt_plane_params {
  double center[3];
  double ex[3], ey[3], ez[3]; /* local axes in the global system (raycing.xyz_from_xz, beamline.py:290-318) */
  double sinAz, cosAz;        /* used instead of the axes when basis_mode == 0 (beamline.py:241-250) */
  int basis_mode;             /* 1: explicit axes (screens; apertures, whose xyz is always a list), 0: azimuth */
  int lostNum;
  int only_positive_path;     /* Screen.expose(onlyPositivePath=...) */
  int has_compressX, has_compressZ;
  double compressX, compressZ;
  int n_blades;               /* aperture blades: kind 0 left (x < d lost), 1 right, 2 bottom, 3 top;
                                 kind 4: round opening of radius d (RoundAperture.propagate, apertures.py:769-825) */
  int blade_kind[4];
  double blade_pos[4];
  int is_beam_stop;
} xrt_plane_params;

/* out = Screen.expose(in): every ray is transported to the screen plane, local coordinates. Footprint min/max
 * of the good rays lands in reduction slots 28..33, their count in slot 34. */
int xrt_expose_f64(size_t n, const xrt_plane_params* p, const xrt_beam* in, const xrt_beam* out,
                   uint64_t* reductions, void* stream);
/* out = Screen.expose(OE.reflect(in)[0]) in ONE main-pass kernel: xrt_reflect_f64 whose global beam goes through
 * Screen.expose (screens.py:225-302) in registers instead of being written to `gb` and read back by xrt_expose_f64 --
 * 264 B/ray less HBM traffic and one launch less for chains that only consume the screen (xrt's run_process functions
 * that plot a screen behind a mirror, e.g. tests/speed/3_Softi_CXIw2D_speed.py:348-409).  Arguments as xrt_reflect_f64
 * (no gb) + the screen's plane, its output beam and its reduction slots (footprint, good-ray count as xrt_expose_f64
 * leaves them).  `out` equals the two-call result bit for bit.  Built for surfaces 0..2 with material kinds 0, 1, 3
 * (returns 4 otherwise: call the two entry points). */
int xrt_reflect_expose_f64(size_t n, const xrt_oe_params* oe, const xrt_material_params* mat, const double* tables,
                           size_t table_len, const xrt_beam* in, const xrt_beam* fallback, const xrt_beam* lb,
                           uint64_t* reductions, const xrt_plane_params* screen, const xrt_beam* out,
                           uint64_t* screen_reductions, int phase, uint64_t index_offset, void* stream);
/* out = RectangularAperture.propagate(in): alive rays are transported to the aperture plane; rays hitting a
 * blade get state = lostNum in `out` AND in `in` (the reference marks the incoming beam, apertures.py:364). */
/* HemisphericScreen.expose (screens.py:422-545): rays meet the sphere of radius R around p->center; positions in the
 * screen's local frame (p->ex, ey, ez), directions stay global, out->theta / out->phi receive the angular
 * coordinates minus their offsets.  Uses center, ex / ey / ez, lostNum, only_positive_path of the plane parameters. */
int xrt_expose_hemispheric_f64(size_t n, const xrt_plane_params* p, double R, double thetaOffset, double phiOffset,
                               const xrt_beam* in, const xrt_beam* out, void* stream);
int xrt_propagate_rect_f64(size_t n, const xrt_plane_params* p, const xrt_beam* in, const xrt_beam* out,
                           uint64_t* reductions, void* stream);

/* ------------------------------------------------------------------ a12 / a13: stand-alone amplitudes
 * Replaces kernel `get_amplitude` (cl/materials.cl:715-748; called from tests/raycing/test_materials.py:624) and
 * the numpy methods Material.get_amplitude (material.py:415-493) / Crystal.get_amplitude (crystal.py:492-645).
 * HOST pointers. mirror kinds: bIn = beamInDotNormal, bOut = bH = NULL. crystal: bIn = beamInDotNormal
 * (surface), bOut = beamOutDotNormal or NULL (= -bIn), bH = beamInDotHNormal or NULL (= bIn).
 * Returns 3 when an energy lies outside the f1/f2 table (the reference raises ValueError, element.py:254). */
int xrt_get_amplitude_f64_host(size_t n, const xrt_material_params* mat, const double* tables, size_t table_len,
                               const double* E, const double* bIn, const double* bOut, const double* bH, double* rs,
                               double* rp);

/* ------------------------------------------------------------------ a5: IntegratedSource.shine building blocks
 * (sources/sybase.py:1469-1812).  The host keeps the while-loop and the flux bookkeeping; the per-ray work is
 *   xrt_source_sample_f64  -> (E, theta, psi) uniform in `window` = {Emin, Emax, Thetamin, Thetamax, Psimin,
 *                             Psimax} (:1583-1617), four standard-normal deviates per ray and a uniform for
 *                             the rejection test (:1637), counter-based Philox4x32-10 keyed by
 *                             (seed, index_offset + i) so ray shards draw disjoint sub-streams;
 *   xrt_undulator_imap_f64 -> Intensity, mJs, mJp (:1619);
 *   xrt_source_fill_f64    -> the Beam (:1651-1721, :1769-1801), optionally gathering accepted samples. */
typedef struct xrt_source_params {
  double center[3];
  double sinAz, cosAz;
  int to_global;
  int uniform;            /* uniformRayDensity */
  double dx, dz, dxprime, dzprime; /* electron beam sizes [mm], divergences [rad] */
  double L0, Np;          /* undulator period [mm], number of periods (photon source size, synchr.py:2200-2204) */
  int has_pitch, has_yaw;
  double cosPitch, sinPitch, cosYaw, sinYaw;
  double eEspread, E1;    /* relative electron energy spread and first-harmonic energy: photon source size with the
                             Tanaka-Kitamura Qa factor (synchr.py:2210-2223, sybase.py:664-674); 0 = no spread */
  double gamma0;          /* Lorentz factor; per-ray gamma = gamma0 + gamma0*eEspread*n5 (synchr.py:2053-2059) */
} xrt_source_params;

int xrt_source_sample_f64(size_t n, uint64_t seed, uint64_t index_offset, const double* window, double* E,
                          double* theta, double* psi, double* n1, double* n2, double* n3, double* n4, double* u,
                          double* n5, void* stream);
int xrt_source_fill_f64(size_t n, const xrt_source_params* p, const int64_t* idx, const double* rE,
                        const double* rTheta, const double* rPsi, const double* mJs, const double* mJp,
                        const double* n1, const double* n2, const double* n3, const double* n4, const xrt_beam* out,
                        size_t out_offset, void* stream);

/* ------------------------------------------------------------------ GeometricSource
 * GeometricSource.shine (sources/geoms.py:193-534): the input generator of BASELINE.json configs 1 and 3.  Per ray
 * up to seven standard deviates (uniform [0, 1) or standard normal, depending on the distribution law): streams
 * 0 polarisation (unpolarised Ep), 1 y, 2 x, 3 z, 4 x', 5 z', 6 E -- drawn from Philox4x32-10 at counter
 * (seed, index_offset + i, stream) or read from `deviates[k]` when that pointer is non-NULL (parity tests feed
 * the numbers the numpy oracle drew; for distE = 'lines' stream 6 then holds the line index). */
#define XRT_GEOM_MAX_LINES 16
typedef struct xrt_geom_params {
  double center[3];
  double sinAz, cosAz;
  int to_global;
  int n_rot;              /* rotate_beam(pitch, roll, yaw) with the default sequence: Rz(yaw), Ry(roll), Rx(pitch) */
  int rot_axis[3];
  double rot_cos[3], rot_sin[3];
  int dist[5];            /* y, x, z, x', z': 0 None, 1 normal, 2 flat, 3 normal sampled uniformly (uniformRayDensity) */
  double p0[5], p1[5];    /* normal: sigma; flat: (aMin, aMax); uniform-normal: (sigma, cutLim) (geoms.py:370-405) */
  int annulus_xz, annulus_ac;   /* 1: (x, z) / (x', z') come from _set_annulus (geoms.py:409-418) */
  double ann_xz[4], ann_ac[4];  /* rMin, rMax, phiMin, phiMax */
  double sqrt_pi2;        /* PI2**0.5 as the host computes it */
  int with_amp;
  double Jss, Jpp, Jsp_re, Jsp_im, Es_re, Es_im, Ep_re, Ep_im;  /* make_polarization (geoms.py:64-175) */
  int ep_random;          /* unpolarised: Ep = uniform * 2**-0.5 */
  int distE;              /* 0 None, 1 normal (e0 mean, e1 sigma), 2 flat (e0, e1), 3 lines */
  double e0, e1;
  int n_lines, weighted;
  double lines[XRT_GEOM_MAX_LINES], cdf[XRT_GEOM_MAX_LINES];
} xrt_geom_params;
/* flag: one uint64 of device memory (zeroed by the caller): counts rays with a^2 + c^2 > 1, which switches the
 * normalisation of the direction for the WHOLE beam (geoms.py:503-511); ray shards all-reduce it between the two
 * phases (phase bit0: sample, bit1: normalise + rotate + to global; 3 = both). */
int xrt_geometric_source_f64(size_t n, const xrt_geom_params* p, uint64_t seed, uint64_t index_offset,
                             const double* const* deviates, const xrt_beam* out, uint64_t* flag, int phase,
                             void* stream);

/* ------------------------------------------------------------------ f3: bending magnet / wiggler
 * BendingMagnet / Wiggler (sources/synchr.py:69-609).  The reference computes these on the CPU only ("a GPU is not
 * required and is not implemented", synchr.py:70-73); here they are device kernels so that the beam is born in HBM. */
typedef struct xrt_bm_params {
  double gamma0;          /* electron Lorentz factor (sybase.py:155) */
  double B;               /* field [T] (synchr.py:88-103) */
  double ro;              /* bending radius [m] */
  double K, L0, Np, X0;   /* wiggler: deflection parameter, period [mm], periods, trajectory amplitude [mm]; a bending
                             magnet has Np = 0.5 (synchr.py:96) */
  double eI;              /* ring current [A] */
  int isMPW;              /* 1: wiggler branches (synchr.py:204-206, :432-448) */
  int bw_mode;            /* 1: distE == 'BW' */
} xrt_bm_params;
/* (I, Es, Ep) = BendingMagnet.build_I_map(w, theta, psi) (synchr.py:185-226); gamma[n] or NULL (then gamma0). */
int xrt_bm_imap_f64(size_t n, const xrt_bm_params* bm, const double* gamma, const double* w, const double* theta,
                    const double* psi, double* I, double* Es, double* Ep, void* stream);
int xrt_bm_imap_f64_host(size_t n, const xrt_bm_params* bm, const double* gamma, const double* w,
                         const double* theta, const double* psi, double* I, double* Es, double* Ep);
/* Beam assembly of BendingMagnet.shine (synchr.py:370-480, :492-505) from samples (idx: optional gather of the
 * accepted ones), (mJs, mJp) of xrt_bm_imap_f64, five standard normal deviates per sample (horizontal divergence,
 * 1/gamma opening, vertical divergence, horizontal size, vertical size) and, for a wiggler, the period index. */
int xrt_bm_fill_f64(size_t n, const struct xrt_source_params* p, const xrt_bm_params* bm, const long long* idx,
                    const double* rE, const double* rTheta, const double* rPsi, const double* mJs, const double* mJp,
                    const double* g1, const double* g2, const double* g3, const double* g4, const double* g5,
                    const double* kint, const xrt_beam* out, size_t out_offset, void* stream);
/* uniformRayDensity shine in one kernel (Philox sampling -> build_I_map -> Beam); imax_slot as xrt_shine_uniform_f64 */
int xrt_bm_shine_uniform_f64(size_t n, uint64_t seed, uint64_t index_offset, const double* window,
                             const struct xrt_source_params* sp, const xrt_bm_params* bm, const xrt_beam* out,
                             uint64_t* imax_slot, void* stream);

/* Fused uniformRayDensity source: xrt_source_sample_f64 + xrt_undulator_imap_f64 + xrt_source_fill_f64 in one
 * kernel (same Philox counters, same device functions => the same Beam bit for bit; samples and (I, mJs, mJp)
 * stay in registers).  *imax_slot (device, 8 bytes, zero-initialised by the caller) receives the bit pattern of
 * max I over the rays (non-negative doubles order like unsigned integers; all-reduce MAX across ranks). */
int xrt_shine_uniform_f64(size_t n, uint64_t seed, uint64_t index_offset, const double* window, int jend, double Kx,
                          double Ky, double L0, double Np, double gamma0, double amp_const, double dstep, int bw_mode,
                          int symmetric, const double* tg, const double* ag, const double* sintg, const double* costg,
                          const double* sintgph, const double* costgph, const xrt_source_params* p,
                          const xrt_beam* out, uint64_t* imax_slot, void* stream);

/* Surface normals of an OE at local Cartesian points (x, y, z): OE.local_n, preceded by xyz_to_param for
 * parametric surfaces, as waves.diffract (waves.py:675-686) and OE.local_to_global (oes/base.py:1213-1217) call
 * it. Device pointers; nx, ny, nz receive the surface normal. */
/* Surface height at local (x, y): z = OE.local_z(x, y), or for parametric surfaces the chain
 * xyz_to_param(x, y, 0) -> local_r(s, phi) -> param_to_xyz(...)[2] of OE.prepare_wave (oes/reflect.py:335-340). */
int xrt_oe_surface_f64(size_t n, const xrt_oe_params* oe, const double* x, const double* y, double* z, void* stream);

int xrt_oe_normals_f64(size_t n, const xrt_oe_params* oe, const double* x, const double* y, const double* z,
                       double* nx, double* ny, double* nz, void* stream);

/* ------------------------------------------------------------------ f2: multilayer recursion
 * Kernel `get_amplitude_graded_multilayer` of the reference (cl/materials.cl:568-622, called from
 * Multilayer.get_amplitude, materials/multilayer.py:471-484, through XRT_CL.run_parallel): the Parratt recursion
 * from the substrate up, per ray, given the interface Fresnel amplitudes rbs, rtb, rvt (s and p), the layer wave
 * vectors Qt, Qb (all complex128, n each, interleaved re/im) and the thicknesses dti, dbi [npairs].  Follows the
 * reference's numpy statements (:432-460: p1 = exp(0.5j Q d), p2 = p1**2, r = (r_ij + r_j p2)/(1 + r_ij r_j p2)).
 * Out: ri_s, ri_p.  Device pointers; the *_host variant stages host arrays. */
int xrt_multilayer_f64(size_t n, int npairs, const double* rbs_s, const double* rbs_p, const double* rtb_s,
                       const double* rtb_p, const double* rvt_s, const double* rvt_p, const double* Qt,
                       const double* Qb, const double* dti, const double* dbi, double* ri_s, double* ri_p,
                       void* stream);
int xrt_multilayer_f64_host(size_t n, int npairs, const double* rbs_s, const double* rbs_p, const double* rtb_s,
                            const double* rtb_p, const double* rvt_s, const double* rvt_p, const double* Qt,
                            const double* Qb, const double* dti, const double* dbi, double* ri_s, double* ri_p);
/* Kernel `get_amplitude_graded_multilayer_tran` (cl/materials.cl; Multilayer.get_amplitude with geom 'transmitted',
 * materials/multilayer.py:485-503): r and t recurse together from the vacuum behind the substrate up through the
 * substrate layer (thickness substThickness, Angstrom) and the stack (:419-462).  in23: table of the 23 sliced
 * complex128[n] inputs in the reference's order -- rvt_s, rvt_p, tvt_s, tvt_p, rbs_s, rbs_p, tbs_s, tbs_p, rsv_s,
 * rsv_p, tsv_s, tsv_p, rbt_s, rbt_p, tbt_s, tbt_p, rtb_s, rtb_p, ttb_s, ttb_p, Qt, Qb, Qs.  Out: ti_s, ti_p.
 * Device pointers (the table itself is a host array of device pointers); the *_host variant stages host arrays. */
int xrt_multilayer_tran_f64(size_t n, int npairs, double substThickness, const double* const* in23, const double* dti,
                            const double* dbi, double* ti_s, double* ti_p, void* stream);
int xrt_multilayer_tran_f64_host(size_t n, int npairs, double substThickness, const double* const* in23,
                                 const double* dti, const double* dbi, double* ti_s, double* ti_p);

/* ------------------------------------------------------------------ f1 (next row): device histogram
 * np.histogram2d(y, x, bins=[ny, nx], range=[[ey0, eyN], [ex0, exN]], weights=w1 (+ w2)) as do_hist2d
 * (xrt/multipro.py:119-171) accumulates it, over the rays whose state passes `state_mask` (bit0 state == 1 good,
 * bit1 == 2 out, bit2 == 3 over, bit3 < 0 dead; raycing.get_output's rayFlag).  ex[nx+1], ey[ny+1]: bin edges
 * (np.linspace on the host).  out[ny * nx] is ACCUMULATED into (zero it first), so repeats / ray shards add up;
 * w1/w2/state may be NULL (weight 1, no filter).  Device pointers. */
int xrt_hist2d_f64(size_t n, const double* x, const double* y, const double* w1, const double* w2,
                   const int32_t* state, int state_mask, const double* ex, int nx, const double* ey, int ny,
                   double* out, void* stream);

/* One XYCPlot's worth of histograms in a single pass over the beam (xrt/multipro.py:296-330 `run`: get_output ->
 * cData01 -> HSV -> RGB -> do_hist1d x 3 + do_hist2d): per selected ray the flux weight (fluxKind) and its three
 * RGB weights (hue = the colour-axis value mapped to [0, 1], saturation = colorSaturation, value = flux;
 * matplotlib.colors.hsv_to_rgb's sextant formula) are accumulated into
 *   hist2d[ny][nx][4], histx[nx][4], histy[ny][4], histc[nc][4]       (channel 0 = flux, 1..3 = R, G, B)
 * with np.histogram / np.histogram2d bin semantics (edges from np.linspace on the host, last edge inclusive).
 * All output arrays are ACCUMULATED into; any of them may be NULL.  counters[6] += {selected, alive, good, out,
 * over, dead} (raycing.get_output's locAlive, locGood, ...).  Device pointers. */
typedef struct xrt_plot_params {
  const double *x, *y, *c;       /* axis data: beam attributes (x, z, E, ...); c NULL: no colour axis (R=G=B=flux) */
  double xfactor, yfactor, cfactor; /* XYCAxis.factor (get_output multiplies the data by it) */
  const double *ex, *ey, *ec;    /* bin edges, nx+1 / ny+1 / nc+1 */
  int nx, ny, nc;
  const double *Jss, *Jpp, *Jsp; /* Jsp: interleaved (re, im) */
  const double *E;
  const int32_t* state;
  int state_mask;                /* rayFlag (raycing/__init__.py:203-218): bit0 state == 1, bit1 == 2, bit2 == 3,
                                    bit4 any other state > 0 (rayFlag 4 = alive sets bits 0, 1, 2, 4), bit3 every
                                    state < 0; exact negative flags go into dead_flags */
  int n_dead_flags;
  int dead_flags[4];             /* rayFlag < 0: rays with exactly this state */
  int flux_kind;                 /* 0 total, 1 's', 2 'p', 3 '+/-45', 4 'left-right', 5 'power' (__init__.py:267-296) */
  double srcWt;                  /* nrays * beam.sourceWeight, or 1 */
  double climits[2], colorFactor, colorSaturation;
  double *hist2d, *histx, *histy, *histc;
  unsigned long long* counters;
} xrt_plot_params;

int xrt_plot_hist_f64(size_t n, const xrt_plot_params* plot, void* stream);

/* ------------------------------------------------------------------ what leaves the GPU: selected rays, selected fields
 * Dense copies, in ray order, of up to 8 arrays restricted to the rays raycing.get_output would select
 * (raycing/__init__.py:170-300).  state_mask: bit0 state == 1 (good), bit1 2 (out), bit2 3 (over), bit3 < 0 (lost),
 * bit4 other positive, bit5 zero -- the mask of xrt_hist2d_f64.  src[k] / dst[k]: DEVICE arrays of words[k] (1 or 2)
 * 8-byte words per ray (2: complex128); the table of pointers itself is HOST memory.  scratch: device memory of
 * xrt_compact_scratch_words(n) uint64; afterwards its LAST word holds the number of selected rays (read it with an
 * 8-byte copy, then copy that many elements of each dst to the host). */
size_t xrt_compact_scratch_words(size_t n);
int xrt_compact_rays_f64(size_t n, const int32_t* state, int state_mask, int narr, const double* const* src,
                         double* const* dst, const int* words, uint64_t* scratch, void* stream);

/* ------------------------------------------------------------------ multi-GPU exchange steps (one process per GPU)
 * Replace the multi-device slicing of XRT_CL.run_parallel_max (myopencl.py:455-583: sliced arguments are cut into one
 * chunk per device, results copied back into the host slices) for hosts that shard rays / receiving points over
 * processes.  NCCL is loaded with dlopen at first use.  `comm` is an ncclComm_t: an existing one, or
 * one made by xrt_nccl_comm_init from the 128-byte id rank 0 obtains with xrt_nccl_unique_id and passes to the
 * other ranks over any host channel.  All buffers are DEVICE pointers; calls are asynchronous on `stream`. */
int xrt_nccl_unique_id(void* id128);
int xrt_nccl_comm_init(int world, int rank, const void* id128, void** comm);
int xrt_nccl_comm_destroy(void* comm);
/* min / max / sum of the 64 reduction slots over all ranks, in place (between the pre-passes and the main pass of
 * xrt_reflect_f64; base.py:861-871, :1257-1263 take these over the whole beam).  sums = 0 keeps this rank's counter
 * slots (they were summed by an earlier call).  scratch: world * 64 uint64 of device memory. */
int xrt_allreduce_slots(void* comm, int world, uint64_t* reductions, int sums, uint64_t* scratch, void* stream);
/* 0 = max, 1 = min, 2 = sum: how slot `slot` combines (the table xrt_allreduce_slots applies) */
int xrt_slot_kind(int slot);
/* The all-gather of the Beam SoA at a Screen: rank r contributes counts[r] rays (shards may be ragged), every rank
 * receives all of them in rank order in `out` (sum(counts) rays).  One grouped NCCL operation for all arrays. */
int xrt_allgather_beam(void* comm, int world, int rank, const uint64_t* counts, const xrt_beam* in,
                       const xrt_beam* out, void* stream);
/* all-gather-v of raw bytes (the packed input records of the Kirchhoff sum, xrt_kirchhoff_pack_f64): counts in bytes */
int xrt_allgather_bytes(void* comm, int world, int rank, const uint64_t* counts, const void* in, void* out,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XRT_B200_H */
