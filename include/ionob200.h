/*
 * ionob200.h -- C ABI of libionob200.so: the B200 (sm_100a) implementation of
 * IonoTomo's ray-integral forward model and its adjoint.
 *
 * This is the drop-in boundary for that path.  The reference has no FFI layer
 * (it is 100 % Python); each entry point below replaces the arithmetic of one
 * reference function, cited as file:line under /root/reference/src/ionotomo/.
 * The Python shims in ionotomo_b200/ bind these with ctypes and keep the
 * reference's signatures; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every array is fp64, C-order, DEVICE memory
 *     unless the parameter name ends in _host;
 *   - the caller owns every buffer; outputs are fully overwritten;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*,
 *     NULL = legacy default stream) and re-entrant; the only library-owned
 *     state is the opaque grid handle;
 *   - return value: IONO_OK, or an error code with text in iono_last_error()
 *     (thread-local).
 */
#ifndef IONOB200_H
#define IONOB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IONO_OK 0
#define IONO_EBADARG 1 /* NULL pointer, non-positive size, non-monotone axis ... */
#define IONO_EOOB 2    /* reserved for host-side wrappers: a sample left the grid */
#define IONO_ECUDA 3   /* CUDA runtime error, see iono_last_error() */

#define IONO_ABI_VERSION 1

/* Ray traversal order used by the forward/adjoint kernels (which rays the warps
 * of one CTA process together).  Results do not depend on it (forward) or only
 * to summation order (adjoint); it only steers L1/L2 locality. */
#define IONO_ORDER_NATURAL 0 /* (a,t,d) with d fastest: consecutive warps take consecutive directions */
#define IONO_ORDER_TIME 1    /* consecutive warps take consecutive times of one (antenna, direction) */
#define IONO_ORDER_ANTENNA 2 /* consecutive warps take consecutive antennas of one (time, direction) */

int iono_version(void);
const char *iono_last_error(void);

/* ---- grid handle -------------------------------------------------------
 * Axis vectors of a TriCubic (geometry/tri_cubic.py:13-47).  Host pointers:
 * they are a few KB and live on the host in the reference.  Builds the device
 * cell tables {g[i], 1/(g[i+1]-g[i])} used by every interpolating kernel and
 * records whether each axis is uniform (direct cell index) or needs bisection
 * (geometry/tri_cubic.py:105-132). */
typedef struct iono_grid *iono_grid_t;
int iono_grid_create(const double *xvec_host, const double *yvec_host, const double *zvec_host,
                     int nx, int ny, int nz, iono_grid_t *grid_out);
int iono_grid_destroy(iono_grid_t grid);
/* 1 if all three axes take the direct-index fast path */
int iono_grid_is_uniform(iono_grid_t grid);

/* ---- per-voxel transforms ----------------------------------------------
 * ne_out[v] = exp(m[v]) * scale, scale = K_ne/TECU:
 * inversion/forward_equation.py:41-43, inversion/gradient.py:49-51. */
int iono_ne_from_m_f64(const double *m, int64_t nvox, double scale, double *ne_out, void *stream);
/* out[v] = a[v] * b[v] (chain rule ne[v] * backprojection) */
int iono_mul_f64(const double *a, const double *b, int64_t n, double *out, void *stream);

/* ---- ray generation ----------------------------------------------------
 * Straight rays, independent variable z: Fermat.integrate_ray
 * (inversion/fermat.py:150-174 with euler_ode :48-84 at n=1) for every ray of
 * cast_ray (geometry/calc_rays.py:61-96).
 * origins, directions: (nrays,3); rays_out: (nrays,4,Ns) rows x,y,z,s. */
int iono_cast_rays_straight_f64(const double *origins, const double *directions, int64_t nrays,
                                double tmax, int Ns, double *rays_out, void *stream);

/* Arc length as the independent variable: Fermat(type='s') (inversion/fermat.py:74-82, :163-166) for
 * straight rays: s = linspace(0, smax, Ns), (x,y,z) = origin + unit(direction) * s. */
int iono_cast_rays_arclength_f64(const double *origins, const double *directions, int64_t nrays,
                                 double smax, int Ns, double *rays_out, void *stream);

/* Frame-aware variant: ITRS inputs, the reference's Pointing frame applied per ray on the fly
 * (astro/frames/pointing_frame.py:140-190 and the per-time loop of calc_rays, calc_rays.py:125-139):
 *   origin[a,t,k]    = R[t] . (ants_itrs_m[a] - p0_itrs_m) / 1000   (km)
 *   direction[a,t,k] = R[t] . dirs_itrs[t,k]
 * R: (Nt,3,3) rows east, north, up of the pointing frame at each obstime (see
 * ionotomo_b200.geometry.frames.pointing_rotation); dirs_itrs: (Nt,Nd,3) unit vectors. */
int iono_cast_rays_frames_f64(const double *ants_itrs_m, const double *p0_itrs_m, const double *R,
                              const double *dirs_itrs, int Na, int Nt, int Nd, double tmax_km, int Ns,
                              double *rays_out, void *stream);

/* "Curved" mode as the reference ships it (Fermat(straight_line_approx=False), fermat.py:48-84: the
 * index gradient is hard-coded to zero, so the geometry stays straight and only ds/dz = n/pz changes):
 * n_out = sqrt(1 - 8.980^2 ne / nu^2) (Fermat.ne2n, fermat.py:36-46), and the s row of `rays`
 * (nrays,4,Ns) is overwritten in place by the optical path int n dz / pz, integrated exactly over the
 * trilinear interpolant (piecewise cubic along the ray) instead of LSODA. */
int iono_ne_to_refractive_index_f64(const double *ne, int64_t nvox, double frequency_hz, double *n_out,
                                    void *stream);
int iono_optical_path_f64(iono_grid_t grid, const double *n_field, double *rays, int64_t nrays, int Ns,
                          unsigned long long *oob_count, void *stream);

/* ---- true tricubic interpolation and bent rays (BASELINE config 5; the reference's notebooks only) ------
 * notebooks/TricubicInterpolation.ipynb[cell 0]:138-299,1192-1257 + DeriveTricubic.ipynb[cell 0]:87-141: C1
 * Lekien-Marsden interpolant from (f, fx, fy, fz, fxy, fxz, fyz, fxyz) at the 8 cell corners, derivatives by
 * 4th-order central differences over the local spacing.  notebooks/FermatClass.ipynb[cell 0]:60-96: rays bent by
 * grad n, independent variable z: dp/dz = grad(n) n/pz, dx/dz = px/pz, dy/dz = py/pz, ds/dz = n/pz.
 * derivs: 8 grids of nx*ny*nz doubles in that order (iono_tricubic_derivs_f64 builds them once per field);
 * iono_tricubic_interp_f64: out[p] = f(x,y,z), grad_out (n,3) optional, *oob_count = points outside the grid;
 * iono_bent_rays_f64: classical RK4, `substeps` steps per sample interval, derivs of the REFRACTIVE INDEX
 * (iono_ne_to_refractive_index_f64); rays_out (nrays,4,Ns) rows x,y,z,s at z = linspace(z0, tmax, Ns);
 * *oob_count = rays that left the grid. */
int iono_tricubic_derivs_f64(iono_grid_t grid, const double *f, double *derivs, void *stream);
int iono_tricubic_interp_f64(iono_grid_t grid, const double *derivs, const double *x, const double *y,
                             const double *z, int64_t n, double *out, double *grad_out,
                             unsigned long long *oob_count, void *stream);
int iono_bent_rays_f64(iono_grid_t grid, const double *derivs, const double *origins, const double *directions,
                       int64_t nrays, double tmax, int Ns, int substeps, double *rays_out,
                       unsigned long long *oob_count, void *stream);

/* ---- point-wise interpolation -------------------------------------------
 * TriCubic.interp / .extrapolate (geometry/tri_cubic.py:69-75) == SciPy
 * RegularGridInterpolator(method='linear').  M: (nx,ny,nz).  oob_count (device,
 * 1 element) receives the number of points outside [g[0],g[-1]] on any axis or
 * NaN -- the caller raises ValueError when extrapolate==0 and it is non-zero. */
int iono_tci_interp_f64(iono_grid_t grid, const double *M, const double *x, const double *y,
                        const double *z, int64_t n, int extrapolate, double *out,
                        unsigned long long *oob_count, void *stream);

/* ---- TEC forward ---------------------------------------------------------
 * tec_out[ray] = simps(interp(ne; x,y,z), s) with simps = scipy's old
 * even='avg' rule: do_forward_equation (inversion/forward_equation.py:13-33).
 * rays: (Na,Nt,Nd,4,Ns); tec_out: (Na,Nt,Nd).  One warp per ray. */
int iono_tec_forward_f64(iono_grid_t grid, const double *ne, const double *rays, int Na, int Nt,
                         int Nd, int Ns, int order, double *tec_out,
                         unsigned long long *oob_count, void *stream);
/* Quad layout of a grid field for the forward gathers: record v = (ix*ny+iy)*nz+iz holds
 * { f[ix,iy,iz], f[ix,iy,iz+1], f[ix,iy+1,iz], f[ix,iy+1,iz+1] } (32 B, steps past the last node
 * clamped), so the 8 corners of a cell are two 256-bit loads.  quads_out: 4*nx*ny*nz doubles,
 * 32-byte aligned.  iono_ne_quads_from_m_f64 fuses ne = exp(m)*scale (forward_equation.py:41-43):
 * one launch writes the quad records and, if ne_out != NULL, the plain ne grid as well.
 * iono_tec_forward_f64 / iono_forwardprojector_apply_f64 (plain ne in) build the records themselves
 * in a stream-ordered temporary when that pays (IONO_FWD_LAYOUT=plain|quads overrides); the
 * *_quads_f64 entry points take records the caller keeps, e.g. across a forward and its line search. */
int iono_quads_from_ne_f64(const double *ne, int nx, int ny, int nz, double *quads_out, void *stream);
int iono_ne_quads_from_m_f64(const double *m, int nx, int ny, int nz, double scale, double *ne_out,
                             double *quads_out, void *stream);
int iono_tec_forward_quads_f64(iono_grid_t grid, const double *quads, const double *rays, int Na, int Nt,
                               int Nd, int Ns, int order, double *tec_out,
                               unsigned long long *oob_count, void *stream);
/* dtec_out = tec - tec[i0,:,:] (inversion/forward_equation.py:50); may alias tec */
int iono_dtec_f64(const double *tec, int Na, int Nt, int Nd, int i0, double *dtec_out, void *stream);

/* ---- adjoint (exact transpose of the dTEC forward) ----
 * coef_out[a,t,d] = dd[a,t,d] - [a==i0] * sum_i dd[i,t,d],
 * dd = (g-dobs)/(CdCt+1e-15) (inversion/gradient.py:33-37). */
int iono_adjoint_coef_f64(const double *g, const double *dobs, const double *CdCt, int Na, int Nt,
                          int Nd, int i0, double *coef_out, void *stream);
/* Everything between the forward and the adjoint in one launch: dtec_out = tec - tec[i0]
 * (forward_equation.py:50), misfit_out[0] = sum((dtec-dobs)^2/(CdCt+1e-15))/2 (line_search.py:48-49),
 * coef_out as iono_adjoint_coef_f64 (may be NULL), coef_perm_out = the same coefficients in the
 * back-projector's internal order (antenna, direction, time) for iono_backprojector_apply_permuted_f64
 * (may be NULL).  scratch: iono_residual_scratch_elems(Na, Nt, Nd) doubles.  Deterministic. */
int64_t iono_residual_scratch_elems(int Na, int Nt, int Nd);
int iono_residual_f64(const double *tec, const double *dobs, const double *CdCt, int Na, int Nt, int Nd,
                      int i0, double *dtec_out, double *coef_out, double *coef_perm_out, double *scratch,
                      double *misfit_out, void *stream);
/* acc[v] (+)= sum_ray coef[ray] sum_s w_s(ray) phi_v(x_s); zero_first!=0 clears
 * acc before accumulating.  acc: (nx,ny,nz).  The voxel gradient of the misfit
 * is ne[v]*acc[v] (iono_mul_f64) after the cross-GPU sum of acc. */
int iono_tec_adjoint_f64(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns,
                         const double *coef, int order, int zero_first, double *acc,
                         unsigned long long *oob_count, void *stream);

/* ---- prepared forward sweep (forward projector) -------------------------------------
 * For a ray geometry that is reused across iterations (the reference's drivers call
 * forward_equation with the same rays every iteration: bfgs_dask.py:207-340,
 * iterative_newton.py:954-1017): per sample the cell index, the in-cell fractions and the
 * Simpson weight are computed once (create) and streamed by apply, 36 B per sample, HBM owned
 * by the handle.  apply writes tec_out[a,t,d] = simps(interp(ne; ray), s), bit-identical to
 * iono_tec_forward_f64 (same device functions, same summation order).  When every ray's weights
 * are one common pattern times a per-ray factor to 2e-13 relative (s a linspace: every ray set the
 * casting entry points make), create stores the weights factored -- 28 B per sample,
 * iono_forwardprojector_factored() == 1 -- and apply agrees with iono_tec_forward_f64 to ~1e-14
 * relative instead of bitwise; the environment variable IONO_PREP_FACTOR=0 (read by create) keeps
 * per-sample weights.  create counts samples outside the grid in *oob_count (device) -- the caller
 * raises like the forward does. */
typedef struct iono_forwardprojector *iono_forwardprojector_t;
int iono_forwardprojector_create(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns,
                                 iono_forwardprojector_t *out, unsigned long long *oob_count,
                                 void *stream);
int iono_forwardprojector_apply_f64(iono_forwardprojector_t fp, const double *ne, double *tec_out,
                                    void *stream);
int iono_forwardprojector_apply_quads_f64(iono_forwardprojector_t fp, const double *quads, double *tec_out,
                                          void *stream);
/* quad records of ne = scale * exp(m) for the records this projector reads only (its rays touch a fraction
 * of the grid; the list is assembled by create): the per-iteration replacement of iono_ne_quads_from_m_f64
 * when this projector is the only consumer of quads_out.  Other records are left untouched. */
int iono_forwardprojector_quads_from_m_f64(iono_forwardprojector_t fp, const double *m, double scale,
                                           double *quads_out, void *stream);
long long iono_forwardprojector_n_records(iono_forwardprojector_t fp);
long long iono_forwardprojector_bytes(iono_forwardprojector_t fp);
int iono_forwardprojector_factored(iono_forwardprojector_t fp);
/* The transpose of the same operator (the adjoint of inversion/gradient.py:15-54 without a second copy of the
 * matrix): acc[v] += sum_ray coef_perm[(a*Nd + d)*Nt + t] * A[ray, v], A the matrix apply() applies, coefficients in
 * the time-fastest order iono_residual_f64 writes as coef_perm, acc the full (nx,ny,nz) grid.  The kernel walks the
 * time axis and aggregates the contributions of consecutive time steps to the same cell in registers; the sums reach
 * acc as fp64 reductions, so results are reproducible to rounding, not bitwise.  acc must be zero on entry at the
 * grid nodes the operator touches (n_voxels / voxels: ascending flat indices, int32); the finish_* calls consume the
 * accumulator at exactly those nodes and zero it again:
 *   finish_gradient: grad[v] = k * exp(m[v]) * acc[v]                (chain rule of ne = K exp(m), k = K/1e13)
 *   finish_compact : out[dst[i]] = acc[voxel_i]  (dst NULL: out[i])  (compact accumulator of the sharded adjoint) */
long long iono_forwardprojector_n_voxels(iono_forwardprojector_t fp);
int iono_forwardprojector_voxels(iono_forwardprojector_t fp, int *out, void *stream);
int iono_forwardprojector_adjoint_f64(iono_forwardprojector_t fp, const double *coef_perm, double *acc, void *stream);
int iono_forwardprojector_finish_gradient_f64(iono_forwardprojector_t fp, double *acc, const double *m, double k,
                                              double *grad, void *stream);
int iono_forwardprojector_finish_compact_f64(iono_forwardprojector_t fp, double *acc, const unsigned int *dst,
                                             double *out, void *stream);
int iono_forwardprojector_destroy(iono_forwardprojector_t fp);

/* ---- chord-length adjoint (the reference's generation-A gradient) ----------------
 * acc[v] (+)= sum_ray dd[ray] * l(ray,v): l = chord of the line first->last sample through the
 * box centred on node v, for the voxels within +-1 cell of any sample (geometry/ray_dirac.py:5-34,
 * geometry/slab_method.py:19-58); the gradient is ne[v]*acc[v] (inversion/gradient.py:15-20).
 * dd: (Na,Nt,Nd) weighted residuals.  Compatibility kernel, fp64 atomics. */
int iono_chord_adjoint_f64(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns,
                           const double *dd, int zero_first, double *acc, void *stream);

/* ---- Gaussian-covariance adjoint ("adjoint B", Cm.G^t.dd) -------------------------
 * acc[v] (+)= sum_ray dd[ray] * simps(sigma_m^2 exp(-|x_v-r(s)|^2/(2 L_m^2)) ne_rays[ray,s], s)
 * over the samples idx_min..idx_max of the ray whose +-Nkernel-cell boxes hold voxel v; the last
 * node of every axis receives nothing (inner loops of inversion/gradient_and_adjoint.py:12-103,
 * `do_adjoint`).  ne_rays: (Na,Nt,Nd,Ns) = K_ne exp(interp(m))/1e13 at the ray samples (:37), made
 * by the caller with iono_tci_interp_f64 + iono_ne_from_m_f64.  L_m = Nkernel*size_cell (:14).
 * dd: (Na,Nt,Nd) weighted residuals.  Compatibility kernel, fp64 atomics. */
int iono_gaussian_adjoint_f64(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns,
                              const double *ne_rays, const double *dd, double sigma_m, double L_m,
                              int Nkernel, int zero_first, double *acc, void *stream);

/* ---- phase-domain ray integrals (reference generation B) -------------------------
 * out[ray,f] = simps(g_f(ne(x_s)), s), n_f = sqrt(1 - ne/(1.2404e-2 nu_f^2)):
 *   dmu == NULL : g_f = 1 - n_f                (forward_equation, iterative_newton.py:108-119)
 *   dmu != NULL : g_f = (ne/n_f) * dmu(x_s)    (prior_penalty_mu, iterative_newton.py:157-179)
 * ne, dmu: (nx,ny,nz) grids; freqs_host: Nf <= 8 frequencies in Hz (host); out: (Na,Nt,Nd,Nf). */
int iono_phase_integrals_f64(iono_grid_t grid, const double *ne, const double *dmu, const double *rays,
                             int Na, int Nt, int Nd, int Ns, const double *freqs_host, int Nf, int order,
                             double *out, unsigned long long *oob_count, void *stream);
/* Simpson integration (old scipy even='avg') of integrands tabulated at the ray samples along the rays' s rows:
 * out[ray*out_stride] = simps(f(y[ray,:]), s[ray,:]);  mode 0: f = y;  mode 1: f = 1 - sqrt(1 + y*c);
 * mode 2: f = y/sqrt(1 + y*c) * y2.  y, y2: (nrays,Ns); rays: (nrays,4,Ns).  The integrate-after-interpolate
 * order of the reference's generation B (iterative_newton.py:108-119, :157-179), used for its bit-compatible
 * reproduction including the axis scramble of TriCubic.interp on 4-D inputs (geometry/tri_cubic.py:69-70). */
int iono_simps_rows_f64(const double *y, const double *y2, const double *rays, int64_t nrays, int Ns, int mode,
                        double c, double *out, int out_stride, void *stream);
/* penalty == 0: out = const[a] + 2 pi nu clock[a,t] - (2 pi nu/c)(I - I[i0])   (iterative_newton.py:107-123)
 * penalty != 0: out = -(2 pi nu/(2 n_p c))(I - I[i0])                          (iterative_newton.py:166-183) */
int iono_phase_assemble_f64(const double *integrals, int Na, int Nt, int Nd, int Nf, int i0,
                            const double *freqs_host, const double *clock, const double *konst, int penalty,
                            double *out, void *stream);

/* ---- voxel-binned back-projector (adjoint without atomics) ----------------------
 * The same linear map as iono_tec_adjoint_f64, assembled once per ray geometry in
 * voxel-major sparse form (sorted (voxel, ray, weight) triples; ~9.5 B per distinct (voxel, ray)
 * pair with run-compressed ray indices, ~520 pairs per ray at the LOFAR case) and then applied
 * as a gather: out[v] = scale[v] * sum_ray A[v,ray] * coef[ray]  (scale may be NULL).
 * Worth it when the rays are reused across iterations, as in every reference driver
 * (tests/test_inversion.py:30-39, bfgs_dask.py:207-340, iterative_newton.py:954-1017).
 * create() synchronises the stream; apply() is asynchronous and bit-reproducible. */
typedef struct iono_backprojector *iono_backprojector_t;
int iono_backprojector_create(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns,
                              iono_backprojector_t *bp_out, unsigned long long *oob_count, void *stream);
int iono_backprojector_apply_f64(iono_backprojector_t bp, const double *coef, const double *scale,
                                 double *out, void *stream);
/* The apply in sixteenths of the operator, for overlapping the cross-GPU sum with the computation:
 * chunks [c0,c1) (issue them in increasing order on one stream, starting at 0); afterwards
 * out[chunk_voxels(c0) : chunk_voxels(c1)) is final and may be all-reduced while later chunks run. */
int iono_backprojector_apply_chunks_f64(iono_backprojector_t bp, const double *coef, const double *scale,
                                        double *out, int c0, int c1, void *stream);
/* the same with the coefficients already in the internal (antenna, direction, time) order
 * (coef_perm_out of iono_residual_f64): no permutation pass */
int iono_backprojector_apply_permuted_f64(iono_backprojector_t bp, const double *coef_perm, const double *scale,
                                          double *out, int c0, int c1, void *stream);
/* The voxel gradient in one call: out[v] = k * exp(m[v]) * sum_ray A[v,ray] coef[ray] -- the chain-rule factor
 * ne[v] = K_ne exp(m[v]) / 1e13 (k = K_ne/1e13; inversion/gradient.py:49-51 with :19) is evaluated for the
 * rows the rays touch only, so no ne grid is needed. */
int iono_backprojector_apply_gradient_f64(iono_backprojector_t bp, const double *coef_perm, const double *m,
                                          double k, double *out, int c0, int c1, void *stream);
/* ne_out[v] = k * exp(m[v]) for the voxels of the operator's rows only (the other entries of ne_out are left
 * untouched): the `scale` of the next apply without a pass over the whole grid. */
int iono_backprojector_ne_rows_f64(iono_backprojector_t bp, const double *m, double k, double *ne_out, void *stream);
/* Sharded rays (one process per GPU): row r of this rank's operator is stored, unscaled, at
 * out_compact[row_dst[r]]; the caller numbers the voxels that ANY rank touches consecutively (from
 * iono_backprojector_row_voxels of every rank), clears out_compact, and sums that compact vector across
 * ranks instead of the whole grid (reference fan-out being replaced: inversion/gradient.py:52-54 da.sum). */
int iono_backprojector_apply_compact_f64(iono_backprojector_t bp, const double *coef_perm,
                                         const unsigned int *row_dst, double *out_compact, int c0, int c1,
                                         void *stream);
long long iono_backprojector_n_rows(iono_backprojector_t bp);
int iono_backprojector_row_voxels(iono_backprojector_t bp, unsigned int *row_voxels_out, void *stream);
long long iono_backprojector_chunk_voxels(iono_backprojector_t bp, int c);
long long iono_backprojector_nnz(iono_backprojector_t bp);
long long iono_backprojector_bytes(iono_backprojector_t bp);
int iono_backprojector_destroy(iono_backprojector_t bp);

/* ---- cross-GPU sum over NVLink peer memory (one process per GPU) ---------------------------------
 * Replaces the reference's da.sum over dask workers (inversion/gradient.py:52-54).  Buffers that peers
 * read or write are allocated with iono_peer_alloc (cudaMalloc + zero fill + CUDA IPC handle, 64 bytes,
 * to be sent to the other processes) and mapped on the other ranks with iono_peer_open.
 * iono_peer_reduce_expand_f64, called by EVERY rank with the same arguments but `me`:
 *   sum[k]   = sum over ranks r = 0..N-1 (in that order: same bits on every rank) of acc[r][k],  k < L
 *   grad[union_voxels[k]] = k_scale * exp(m[union_voxels[k]]) * sum[k]     for k < n_union   (local grid)
 *   misfit_out[0] = sum[n_union]                                           (if misfit_out != NULL)
 * as ONE kernel per rank: flag handshake, reduce-scatter by peer loads, all-gather by peer stores, flag
 * handshake, expansion.  acc, res, flags: arrays of N device pointers (index = rank; entry `me` is this
 * rank's own allocation): compact accumulators (L doubles, L even, n_union < L), result vectors (L doubles)
 * and flag blocks (iono_peer_flag_bytes(), zeroed).  The launch has no per-call arguments (the call
 * counter lives in the flag block), so it can be replayed from a CUDA graph. */
int iono_peer_alloc(int64_t bytes, void **ptr_out, void *ipc_handle_out64);
int iono_peer_open(const void *ipc_handle64, void **ptr_out);
int iono_peer_close(void *ptr);
int iono_peer_free(void *ptr);
int64_t iono_peer_flag_bytes(void);
int iono_peer_reduce_expand_f64(void *const *acc, void *const *res, void *const *flags, int N, int me,
                                int64_t L, const int *union_voxels, int64_t n_union, const double *m,
                                double k_scale, double *grad, double *misfit_out, void *stream);

/* ---- vector algebra of the device-resident inversion driver -------------------------------------
 * The reference's BFGS recursion evaluates one scalar product (a triple Simpson integral over the grid,
 * bfgs_dask.py:165-167) and one axpy per dask task (bfgs_dask.py:34-56, :165-194).  Here the history is ONE
 * matrix H (rows of length n, leading dimension ld >= n) and an iteration needs a constant number of passes:
 *   iono_multi_dot_f64 : out[r] = sum_i w[i] H[r][i] x[i], r < rows <= 32, one pass over x (w may be NULL;
 *                        with the grid's Simpson weights it is the reference's inner product); scratch:
 *                        iono_multi_dot_scratch_elems() doubles; deterministic
 *   iono_lincomb_f64   : out[i] = coef[0] x[i] + sum_r coef[r+1] H[r][i]; coef (rows+1 doubles) in DEVICE
 *                        memory, x may be NULL
 *   iono_gather_f64    : out[i] = src[idx[i]]                       (grid -> active voxels)
 *   iono_scatter_axpy_f64 : dst[idx[i]] = base[idx[i]] + alpha_dev[0] * x[i]   (active voxels -> grid) */
int64_t iono_multi_dot_scratch_elems(void);
int iono_multi_dot_f64(const double *H, int64_t ld, int rows, const double *x, const double *w, int64_t n,
                       double *scratch, double *out, void *stream);
/* three right-hand sides that are rows of H themselves, one pass over the history:
 * out[j*32 + r] = sum_i w[i] H[r][i] H[x_row_j][i], j < 3 (out: 96 doubles) */
int iono_multi_dot3_f64(const double *H, int64_t ld, int rows, int x_row0, int x_row1, int x_row2, const double *w,
                        int64_t n, double *scratch, double *out, void *stream);
int iono_lincomb_f64(const double *H, int64_t ld, int rows, const double *coef_dev, const double *x, int64_t n,
                     double *out, void *stream);
int iono_zero_f64(double *x, int64_t n, void *stream);   /* x[0..n) = 0 (memset node, no kernel) */
int iono_gather_f64(const double *src, const int *idx, int64_t n, double *out, void *stream);
int iono_scatter_set_f64(const double *x, const int *idx, int64_t n, double *dst, void *stream);   /* dst[idx[i]] = x[i] */
int iono_scatter_axpy_f64(const double *base, const double *alpha_dev, const double *x, const int *idx, int64_t n,
                          double *dst, void *stream);

/* ---- misfit ---------------------------------------------------------------
 * out[0] = sum((g-dobs)^2/(CdCt+1e-15))/2 (inversion/line_search.py:48-49).
 * Deterministic two-stage reduction; `scratch` needs iono_misfit_scratch_elems()
 * doubles. */
int64_t iono_misfit_scratch_elems(void);
int iono_misfit_f64(const double *g, const double *dobs, const double *CdCt, int64_t n,
                    double *scratch, double *out, void *stream);

/* ---- model-covariance smoothing ------------------------------------------------------
 * out = scipy.ndimage.convolve(phi, stencil, mode='nearest') for an (m,m,m) stencil, m odd:
 * Covariance.smooth (ionosphere/covariance.py:383-385), the Cm . (G^T r) step after the adjoint.
 * phi, out: (nx,ny,nz), must not alias. */
int iono_convolve3d_nearest_f64(const double *phi, int nx, int ny, int nz, const double *stencil, int m,
                                double *out, void *stream);

/* ---- host <-> device staging ---------------------------------------------------
 * Strided block copy of `height` rows of `width_bytes` from pinned or pageable HOST memory
 * to DEVICE memory (cudaMemcpy2DAsync).  Used to stream time blocks rays[:, t0:t1] of a
 * host-resident (Na,Nt,Nd,4,Ns) ray array -- the layout the reference's callers hold
 * (geometry/calc_rays.py:78-92) -- while the previous block is being integrated. */
int iono_copy2d_h2d(void *dst_dev, int64_t dst_pitch_bytes, const void *src_host, int64_t src_pitch_bytes,
                    int64_t width_bytes, int64_t height, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* IONOB200_H */
