// True tricubic interpolation and bent rays (BASELINE config 5, "where the reference implements it": its notebooks).
// Included by iono_kernels.cu.
//
// The shipped package interpolates trilinearly and never bends a ray (geometry/tri_cubic.py:22, fermat.py:53-55);
// the C1 interpolant and the Hamiltonian ray equations exist only as notebook code:
//   * notebooks/TricubicInterpolation.ipynb[cell 0]:138-299,1192-1257 and notebooks/DeriveTricubic.ipynb[cell 0]:87-141
//     -- Lekien-Marsden tricubic: per cell the 64 coefficients follow from (f, fx, fy, fz, fxy, fxz, fyz, fxyz) at the
//     8 corners; the derivatives are 4th-order central differences (8(f[i+1]-f[i-1]) - (f[i+2]-f[i-2]))/12 divided by
//     the local spacing (x[i+1]-x[i-1])/2, mixed ones by repeated application;
//   * notebooks/FermatClass.ipynb[cell 0]:60-96 -- state [p, x, s], independent variable z:
//     dp/dz = grad(n) n / pz,  dx/dz = px/pz,  dy/dz = py/pz,  ds/dz = n/pz.
// Here: the 8 derivative grids are built once per field (iono_tricubic_derivs_f64); the interpolant is evaluated as
// the tensor product of cubic Hermite bases -- the unique tricubic that matches those 64 corner values, i.e. the same
// polynomial the notebook's 64x64 matrix yields -- with the derivatives scaled by the cell size (the notebook feeds
// physical derivatives to a unit-cube formula, which is consistent only for unit spacing); nodes closer than two
// cells to a face use 2nd-order / one-sided differences so that every cell is usable (the notebook asserts
// 2 <= i <= n-3).  Rays are integrated with classical RK4 in z, `substeps` steps per sample interval, one thread per
// ray.  There are no reference numbers to pin; the CPU restatement used by the tests is validated against SciPy's
// odeint and against exact tricubic polynomials.
#pragma once

// d/d(axis) of `in` at every node: 4th-order central where two neighbours exist on both sides, 2nd-order central
// next to the faces, one-sided on the faces; spacing = half the distance between the neighbours used.
__global__ void __launch_bounds__(256) diff_axis_kernel(const double *__restrict__ in, double *__restrict__ out, Grid g,
                                                         int axis) {
    const int nx = g.ax[0].n, ny = g.ax[1].n, nz = g.ax[2].n;
    const long long n = (long long)nx * ny * nz;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long st = (axis == 0) ? (long long)ny * nz : (axis == 1 ? nz : 1);
    const int na = g.ax[axis].n;
    const double2 *tab = g.ax[axis].tab;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const int i = (axis == 0) ? (int)(v / ((long long)ny * nz)) : (axis == 1 ? (int)((v / nz) % ny) : (int)(v % nz));
        double d;
        if (i >= 2 && i <= na - 3) {
            const double h = 0.5 * (tab[i + 1].x - tab[i - 1].x);
            d = (8.0 * (in[v + st] - in[v - st]) - (in[v + 2 * st] - in[v - 2 * st])) / 12.0 / h;
        } else if (i >= 1 && i <= na - 2) {
            d = (in[v + st] - in[v - st]) / (tab[i + 1].x - tab[i - 1].x);
        } else if (i == 0) {
            d = (in[v + st] - in[v]) / (tab[1].x - tab[0].x);
        } else {
            d = (in[v] - in[v - st]) / (tab[na - 1].x - tab[na - 2].x);
        }
        out[v] = d;
    }
}

// derivs: 8 grids of V doubles: f, fx, fy, fz, fxy, fxz, fyz, fxyz
extern "C" int iono_tricubic_derivs_f64(iono_grid_t grid, const double *f, double *derivs, void *stream) {
    if (!grid || !f || !derivs) return fail(IONO_EBADARG, "iono_tricubic_derivs_f64: bad argument");
    if (device_check(grid->device, "iono_tricubic_derivs_f64")) return IONO_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const long long V = (long long)grid->nx * grid->ny * grid->nz;
    double *d = derivs;
    CU_CHECK(cudaMemcpyAsync(d, f, (size_t)V * sizeof(double), cudaMemcpyDeviceToDevice, st));
    const int blocks = ew_grid(V);
    diff_axis_kernel<<<blocks, 256, 0, st>>>(d, d + 1 * V, grid->dev, 0);           // fx
    diff_axis_kernel<<<blocks, 256, 0, st>>>(d, d + 2 * V, grid->dev, 1);           // fy
    diff_axis_kernel<<<blocks, 256, 0, st>>>(d, d + 3 * V, grid->dev, 2);           // fz
    diff_axis_kernel<<<blocks, 256, 0, st>>>(d + 1 * V, d + 4 * V, grid->dev, 1);   // fxy
    diff_axis_kernel<<<blocks, 256, 0, st>>>(d + 1 * V, d + 5 * V, grid->dev, 2);   // fxz
    diff_axis_kernel<<<blocks, 256, 0, st>>>(d + 2 * V, d + 6 * V, grid->dev, 2);   // fyz
    diff_axis_kernel<<<blocks, 256, 0, st>>>(d + 4 * V, d + 7 * V, grid->dev, 2);   // fxyz
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

// cubic Hermite bases on [0,1]: value at 0, value at 1, slope at 0, slope at 1 -- and their derivatives
__device__ __forceinline__ void hermite(double t, double h[4], double dh[4]) {
    const double t2 = t * t, t3 = t2 * t;
    h[0] = 2.0 * t3 - 3.0 * t2 + 1.0; h[1] = -2.0 * t3 + 3.0 * t2; h[2] = t3 - 2.0 * t2 + t; h[3] = t3 - t2;
    dh[0] = 6.0 * t2 - 6.0 * t; dh[1] = -6.0 * t2 + 6.0 * t; dh[2] = 3.0 * t2 - 4.0 * t + 1.0; dh[3] = 3.0 * t2 - 2.0 * t;
}

// f and its physical gradient at (x,y,z); returns true if the point is outside the grid (edge cell extrapolated)
__device__ bool tricubic_eval(const Grid &g, const double *__restrict__ D, long long V, double x, double y, double z,
                              double &f, double &fx, double &fy, double &fz) {
    int ix, iy, iz;
    double u, v, w;
    bool oob = false;
    locate<false>(g.ax[0].tab, g.ax[0], x, ix, u, oob);
    locate<false>(g.ax[1].tab, g.ax[1], y, iy, v, oob);
    locate<false>(g.ax[2].tab, g.ax[2], z, iz, w, oob);
    const int ny = g.ax[1].n, nz = g.ax[2].n;
    const double hx = g.ax[0].tab[ix + 1].x - g.ax[0].tab[ix].x, hy = g.ax[1].tab[iy + 1].x - g.ax[1].tab[iy].x,
                 hz = g.ax[2].tab[iz + 1].x - g.ax[2].tab[iz].x;
    double bu[4], du[4], bv[4], dv[4], bw[4], dw[4];
    hermite(u, bu, du); hermite(v, bv, dv); hermite(w, bw, dw);
    f = fx = fy = fz = 0.0;
    // derivative order (a,b,c) in (x,y,z) -> grid index in D: 000 f, 100 fx, 010 fy, 001 fz, 110 fxy, 101 fxz, 011 fyz, 111 fxyz
    const int which[2][2][2] = {{{0, 3}, {2, 6}}, {{1, 5}, {4, 7}}};
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const long long c = ((long long)(ix + i) * ny + (iy + j)) * nz + (iz + k);
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b)
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) {
                            const double val = __ldg(D + which[a][b][cc] * V + c) * (a ? hx : 1.0) * (b ? hy : 1.0) *
                                               (cc ? hz : 1.0);
                            const double Bx = bu[2 * a + i], By = bv[2 * b + j], Bz = bw[2 * cc + k];
                            f = fma(val, Bx * By * Bz, f);
                            fx = fma(val, du[2 * a + i] * By * Bz, fx);
                            fy = fma(val, Bx * dv[2 * b + j] * Bz, fy);
                            fz = fma(val, Bx * By * dw[2 * cc + k], fz);
                        }
            }
    fx /= hx; fy /= hy; fz /= hz;
    return oob;
}

__global__ void __launch_bounds__(128) tricubic_interp_kernel(Grid g, const double *__restrict__ D, long long V,
                                                               const double *__restrict__ x, const double *__restrict__ y,
                                                               const double *__restrict__ z, long long n,
                                                               double *__restrict__ out, double *__restrict__ grad,
                                                               unsigned long long *oob_count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned int n_oob = 0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        double f, fx, fy, fz;
        n_oob += tricubic_eval(g, D, V, x[p], y[p], z[p], f, fx, fy, fz);
        out[p] = f;
        if (grad) { grad[3 * p] = fx; grad[3 * p + 1] = fy; grad[3 * p + 2] = fz; }
    }
    if (n_oob) atomicAdd(oob_count, (unsigned long long)n_oob);
}

extern "C" int iono_tricubic_interp_f64(iono_grid_t grid, const double *derivs, const double *x, const double *y,
                                        const double *z, int64_t n, double *out, double *grad_out,
                                        unsigned long long *oob_count, void *stream) {
    if (!grid || !derivs || !oob_count || n < 0 || (n > 0 && (!x || !y || !z || !out)))
        return fail(IONO_EBADARG, "iono_tricubic_interp_f64: bad argument");
    if (device_check(grid->device, "iono_tricubic_interp_f64")) return IONO_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(oob_count, 0, sizeof(unsigned long long), st));
    if (n == 0) return IONO_OK;
    const long long V = (long long)grid->nx * grid->ny * grid->nz;
    long long blocks = (n + 127) / 128;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    tricubic_interp_kernel<<<(int)blocks, 128, 0, st>>>(grid->dev, derivs, V, x, y, z, n, out, grad_out, oob_count);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

// Bent rays, z as the independent variable (FermatClass.ipynb type 'z'): RK4 on [px, py, pz, x, y, s].
struct RayState { double px, py, pz, x, y, s; };

__device__ __forceinline__ bool ray_rhs(const Grid &g, const double *__restrict__ D, long long V, const RayState &q,
                                        double z, RayState &d) {
    double n, nx, ny, nz;
    const bool oob = tricubic_eval(g, D, V, q.x, q.y, z, n, nx, ny, nz);
    const double ipz = 1.0 / q.pz;
    d.px = nx * n * ipz; d.py = ny * n * ipz; d.pz = nz * n * ipz;
    d.x = q.px * ipz; d.y = q.py * ipz; d.s = n * ipz;
    return oob;
}
__device__ __forceinline__ RayState ray_axpy(const RayState &a, double h, const RayState &d) {
    RayState r;
    r.px = fma(h, d.px, a.px); r.py = fma(h, d.py, a.py); r.pz = fma(h, d.pz, a.pz);
    r.x = fma(h, d.x, a.x); r.y = fma(h, d.y, a.y); r.s = fma(h, d.s, a.s);
    return r;
}

__global__ void __launch_bounds__(128) bent_rays_kernel(Grid g, const double *__restrict__ D, long long V,
                                                         const double *__restrict__ origins,
                                                         const double *__restrict__ directions, long long nrays,
                                                         double tmax, int Ns, int substeps, double *__restrict__ rays,
                                                         unsigned long long *oob_count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned int n_oob = 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nrays; r += stride) {
        const double *o = origins + 3 * r, *dir = directions + 3 * r;
        const double nrm = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
        RayState q;
        q.px = dir[0] / nrm; q.py = dir[1] / nrm; q.pz = dir[2] / nrm; q.x = o[0]; q.y = o[1]; q.s = 0.0;
        const double z0 = o[2];
        const double dzs = (Ns > 1) ? (tmax - z0) / (double)(Ns - 1) : 0.0;
        double *out = rays + r * 4 * (long long)Ns;
        bool bad = false;
        for (int i = 0; i < Ns; ++i) {
            const double zi = (i == Ns - 1 && Ns > 1) ? tmax : z0 + i * dzs;
            out[i] = q.x; out[Ns + i] = q.y; out[2 * (long long)Ns + i] = zi; out[3 * (long long)Ns + i] = q.s;
            if (i == Ns - 1) break;
            const double h = dzs / substeps;
            for (int k = 0; k < substeps; ++k) {
                const double z = zi + k * h;
                RayState k1, k2, k3, k4;
                bad |= ray_rhs(g, D, V, q, z, k1);
                bad |= ray_rhs(g, D, V, ray_axpy(q, 0.5 * h, k1), z + 0.5 * h, k2);
                bad |= ray_rhs(g, D, V, ray_axpy(q, 0.5 * h, k2), z + 0.5 * h, k3);
                bad |= ray_rhs(g, D, V, ray_axpy(q, h, k3), z + h, k4);
                RayState sum;
                sum.px = k1.px + 2.0 * k2.px + 2.0 * k3.px + k4.px; sum.py = k1.py + 2.0 * k2.py + 2.0 * k3.py + k4.py;
                sum.pz = k1.pz + 2.0 * k2.pz + 2.0 * k3.pz + k4.pz; sum.x = k1.x + 2.0 * k2.x + 2.0 * k3.x + k4.x;
                sum.y = k1.y + 2.0 * k2.y + 2.0 * k3.y + k4.y; sum.s = k1.s + 2.0 * k2.s + 2.0 * k3.s + k4.s;
                q = ray_axpy(q, h / 6.0, sum);
            }
        }
        n_oob += bad;
    }
    if (n_oob) atomicAdd(oob_count, (unsigned long long)n_oob);
}

// derivs: the 8 grids of the REFRACTIVE INDEX field (iono_ne_to_refractive_index_f64 + iono_tricubic_derivs_f64);
// rays_out (nrays,4,Ns) rows x, y, z, s at z = linspace(z0, tmax, Ns); *oob_count = rays that left the grid.
extern "C" int iono_bent_rays_f64(iono_grid_t grid, const double *derivs, const double *origins,
                                  const double *directions, int64_t nrays, double tmax, int Ns, int substeps,
                                  double *rays_out, unsigned long long *oob_count, void *stream) {
    if (!grid || !derivs || !oob_count || nrays < 0 || Ns < 1 || substeps < 1 || (nrays > 0 && (!origins || !directions || !rays_out)))
        return fail(IONO_EBADARG, "iono_bent_rays_f64: bad argument");
    if (device_check(grid->device, "iono_bent_rays_f64")) return IONO_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(oob_count, 0, sizeof(unsigned long long), st));
    if (nrays == 0) return IONO_OK;
    const long long V = (long long)grid->nx * grid->ny * grid->nz;
    long long blocks = (nrays + 127) / 128;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    bent_rays_kernel<<<(int)blocks, 128, 0, st>>>(grid->dev, derivs, V, origins, directions, nrays, tmax, Ns, substeps,
                                                  rays_out, oob_count);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}
