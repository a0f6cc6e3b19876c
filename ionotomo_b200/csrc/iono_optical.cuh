// Optical path along straight rays: the reference's `straight_line_approx=False` mode as shipped.
// Included by iono_kernels.cu.
//
// Fermat.euler_ode with straight_line_approx=False (inversion/fermat.py:48-84) interpolates the
// refractive index n but hard-codes its gradient to zero (:53-55), so the trajectory stays the
// straight line of the closed form and only  ds/dz = n(x(z),y(z),z) / pz  changes: the fourth row
// of the ray array becomes the optical path  s_i = int_{z0}^{z_i} n dz / pz  (LSODA in the
// reference, rtol ~1.5e-8).  n is the trilinear interpolant of sqrt(1 - 8.98^2 ne / nu^2)
// (Fermat.ne2n, fermat.py:36-46), i.e. a cubic polynomial in z inside each grid cell along a
// straight line: every sample interval is cut at its cell-boundary crossings and each piece is
// integrated with 2-point Gauss-Legendre, which is exact for cubics.  Warp per ray, lanes take
// intervals, a warp scan turns interval integrals into the cumulative path.
#pragma once

// smallest k with g[k] > v  (n if none)
__device__ __forceinline__ int upper_index(const double2 *__restrict__ tab, int n, double v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tab[mid].x > v) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ double tril_at(const double *__restrict__ f, const Grid &g, double x, double y, double z,
                                          bool &oob) {
    int ix, iy, iz;
    double tx, ty, tz;
    locate<false>(g.ax[0].tab, g.ax[0], x, ix, tx, oob);
    locate<false>(g.ax[1].tab, g.ax[1], y, iy, ty, oob);
    locate<false>(g.ax[2].tab, g.ax[2], z, iz, tz, oob);
    const int ny = g.ax[1].n, nz = g.ax[2].n;
    return trilerp(f + ((long long)ix * ny + iy) * nz + iz, nz, ny * nz, tx, ty, tz);
}

__global__ void __launch_bounds__(256) optical_path_kernel(Grid g, const double *__restrict__ nfield,
                                                            double *__restrict__ rays, int R, int Ns,
                                                            unsigned long long *oob_count) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int nx = g.ax[0].n, ny = g.ax[1].n, nz = g.ax[2].n;
    unsigned int n_oob = 0;
    const double gl = 0.28867513459481288225;   // 1/(2 sqrt 3)
    for (int ray = warp_global; ray < R; ray += n_warps) {
        double *rp = rays + (long long)ray * 4 * Ns;
        // pz from the geometric path of the first interval: s_1 = (z_1 - z_0)/pz
        const double pz = (rp[2 * Ns + Ns - 1] - rp[2 * Ns]) / (rp[3 * Ns + Ns - 1] - rp[3 * Ns]);
        double carry = 0.0;
        for (int base = 0; base < Ns - 1; base += 32) {
            const int i = base + lane;
            double seg = 0.0;
            if (i < Ns - 1) {
                const double za = rp[2 * Ns + i], zb = rp[2 * Ns + i + 1];
                const double xa = rp[i], ya = rp[Ns + i];
                const double kx = (rp[i + 1] - xa) / (zb - za), ky = (rp[Ns + i + 1] - ya) / (zb - za);
                double zc = za;
                bool oob = false;
                for (int it = 0; it < 4096 && zc < zb; ++it) {
                    double zn = zb;
                    {   // next z plane
                        const int k = upper_index(g.ax[2].tab, nz, zc);
                        if (k < nz && g.ax[2].tab[k].x < zn) zn = g.ax[2].tab[k].x;
                    }
                    const double xc = xa + kx * (zc - za), yc = ya + ky * (zc - za);
                    if (kx > 0.0) {
                        const int k = upper_index(g.ax[0].tab, nx, xc);
                        if (k < nx) { const double zx = za + (g.ax[0].tab[k].x - xa) / kx; if (zx > zc && zx < zn) zn = zx; }
                    } else if (kx < 0.0) {
                        int k = upper_index(g.ax[0].tab, nx, xc) - 1;          // last node <= xc
                        if (k >= 0 && g.ax[0].tab[k].x >= xc) --k;              // strictly below
                        if (k >= 0) { const double zx = za + (g.ax[0].tab[k].x - xa) / kx; if (zx > zc && zx < zn) zn = zx; }
                    }
                    if (ky > 0.0) {
                        const int k = upper_index(g.ax[1].tab, ny, yc);
                        if (k < ny) { const double zy = za + (g.ax[1].tab[k].x - ya) / ky; if (zy > zc && zy < zn) zn = zy; }
                    } else if (ky < 0.0) {
                        int k = upper_index(g.ax[1].tab, ny, yc) - 1;
                        if (k >= 0 && g.ax[1].tab[k].x >= yc) --k;
                        if (k >= 0) { const double zy = za + (g.ax[1].tab[k].x - ya) / ky; if (zy > zc && zy < zn) zn = zy; }
                    }
                    if (!(zn > zc)) zn = zb;   // no progress (rounding): finish the interval in one piece
                    const double h = zn - zc, zm = 0.5 * (zc + zn);
                    const double z1 = zm - gl * h, z2 = zm + gl * h;
                    const double n1 = tril_at(nfield, g, xa + kx * (z1 - za), ya + ky * (z1 - za), z1, oob);
                    const double n2 = tril_at(nfield, g, xa + kx * (z2 - za), ya + ky * (z2 - za), z2, oob);
                    seg += 0.5 * h * (n1 + n2);
                    zc = zn;
                }
                n_oob += oob;
            }
            // inclusive warp scan of the interval integrals
            double incl = seg;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const double total = __shfl_sync(0xffffffffu, incl, 31);
            if (i < Ns - 1) rp[3 * Ns + i + 1] = (carry + incl) / pz;
            carry += total;
        }
        if (lane == 0) rp[3 * Ns] = 0.0;
    }
    if (n_oob) atomicAdd(oob_count, (unsigned long long)n_oob);
}

__global__ void __launch_bounds__(256) ne_to_n_kernel(const double *__restrict__ ne, int64_t n, double a,
                                                       double *__restrict__ out) {
    // Fermat.ne2n (fermat.py:36-46): M *= -8.980^2/nu^2; M += 1; sqrt
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __dsqrt_rn(__dadd_rn(__dmul_rn(ne[i], a), 1.0));
}

extern "C" int iono_ne_to_refractive_index_f64(const double *ne, int64_t nvox, double frequency_hz, double *n_out,
                                               void *stream) {
    if (!ne || !n_out || nvox < 0 || !(frequency_hz > 0)) return fail(IONO_EBADARG, "iono_ne_to_refractive_index_f64: bad argument");
    if (nvox == 0) return IONO_OK;
    const double a = -8.980 * 8.980 / (frequency_hz * frequency_hz);
    ne_to_n_kernel<<<ew_grid(nvox), 256, 0, (cudaStream_t)stream>>>(ne, nvox, a, n_out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

extern "C" int iono_optical_path_f64(iono_grid_t grid, const double *n_field, double *rays, int64_t nrays, int Ns,
                                     unsigned long long *oob_count, void *stream) {
    if (!grid || !n_field || !oob_count || nrays < 0 || Ns < 1 || (nrays > 0 && !rays))
        return fail(IONO_EBADARG, "iono_optical_path_f64: bad argument");
    if (sweep_size_check(grid, nrays, Ns)) return IONO_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(oob_count, 0, sizeof(unsigned long long), st));
    if (nrays == 0 || Ns < 2) return IONO_OK;
    long long ctas = (nrays + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    optical_path_kernel<<<(int)(ctas < cap ? ctas : cap), 256, 0, st>>>(grid->dev, n_field, rays, (int)nrays, Ns,
                                                                     oob_count);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}
