// Gaussian-covariance adjoint ("adjoint B", Cm.G^t.dd): the inner loops of
// inversion/gradient_and_adjoint.py:12-103 (`do_adjoint`).  Included by iono_kernels.cu.
//
//   acc[v] += dd[ray] * simps( sigma_m^2 exp(-|x_v - r(s)|^2 / (2 L_m^2)) ne(s), s )   over idx_min..idx_max
//
// where, per (ray, voxel), idx_min / idx_max are the first / last sample whose box holds the
// voxel.  The box of a sample is +-Nkernel cells around its `bisection` cell on every axis, and
// the reference's slice bound min(n-1, c+Nk+1) is exclusive (gradient_and_adjoint.py:42-44), so
// the last node of an axis never receives anything.  ALL samples idx_min..idx_max enter the
// Simpson sum (:80-92), with the old-SciPy even='avg' rule for an even count and 0 for one sample.
// ne(s) = K_ne exp(interp(m))/1e13 along the ray is an input (the host shim makes it with
// iono_tci_interp_f64 + iono_ne_from_m_f64, which also gives the reference's ValueError).
//
// Mapping: the one of the chord kernel (iono_chord.cuh) -- one warp per ray, cell indices of the
// samples in shared memory, lanes take z-levels, a (voxel, ray) pair is owned by the first sample
// that reaches it.  Compatibility kernel: the reference cannot run this beyond toy sizes.
#pragma once

__device__ __forceinline__ bool gauss_member(const int *cx, const int *cy, const int *cz, int q, int xi, int yi,
                                             int zi, int Nk) {
    const int dx = cx[q] - xi, dy = cy[q] - yi, dz = cz[q] - zi;
    return dx >= -Nk && dx <= Nk && dy >= -Nk && dy <= Nk && dz >= -Nk && dz <= Nk;
}

__global__ void __launch_bounds__(256) gaussian_adjoint_kernel(Grid g, const double *__restrict__ rays, int R, int Ns,
                                                                const double *__restrict__ ne_rays,
                                                                const double *__restrict__ dd, double sigma2,
                                                                double minus_two_L2, int Nk,
                                                                double *__restrict__ acc) {
    extern __shared__ int gauss_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int *cx = gauss_smem + warp * 3 * Ns, *cy = cx + Ns, *cz = cy + Ns;
    const int nx = g.ax[0].n, ny = g.ax[1].n, nz = g.ax[2].n;
    const double2 *tx = g.ax[0].tab, *ty = g.ax[1].tab, *tz = g.ax[2].tab;
    for (int ray = blockIdx.x * nwarp + warp; ray < R; ray += gridDim.x * nwarp) {
        const double *rp = rays + (long long)ray * 4 * Ns;
        const double *sp = rp + 3 * Ns;
        const double *nep = ne_rays + (long long)ray * Ns;
        __syncwarp();
        for (int s = lane; s < Ns; s += 32) {
            cx[s] = ref_bisection(tx, nx, rp[s]);
            cy[s] = ref_bisection(ty, ny, rp[Ns + s]);
            cz[s] = ref_bisection(tz, nz, rp[2 * Ns + s]);
        }
        __syncwarp();
        int mono = 1;
        for (int s = lane + 1; s < Ns; s += 32) mono &= (cz[s] >= cz[s - 1]);
        mono = __all_sync(0xffffffffu, mono);
        const double w = dd[ray];
        for (int zi = lane; zi <= nz - 2; zi += 32) {
            int s_lo = 0, s_hi = Ns;
            if (mono) {   // samples with |cz - zi| <= Nk
                int lo = 0, hi = Ns;
                while (lo < hi) { const int m = (lo + hi) >> 1; if (cz[m] < zi - Nk) lo = m + 1; else hi = m; }
                s_lo = lo;
                hi = Ns;
                while (lo < hi) { const int m = (lo + hi) >> 1; if (cz[m] <= zi + Nk) lo = m + 1; else hi = m; }
                s_hi = lo;
            }
            const double zv = tz[zi].x;
            for (int s = s_lo; s < s_hi; ++s) {
                const int dz = cz[s] - zi;
                if (dz < -Nk || dz > Nk) continue;
                const int xc = cx[s], yc = cy[s];
                for (int xi = max(0, xc - Nk); xi <= min(nx - 2, xc + Nk); ++xi)
                    for (int yi = max(0, yc - Nk); yi <= min(ny - 2, yc + Nk); ++yi) {
                        // owned by the first sample whose box holds (xi, yi, zi): that sample is idx_min
                        bool seen = false;
                        for (int q = s_lo; q < s && !seen; ++q) seen = gauss_member(cx, cy, cz, q, xi, yi, zi, Nk);
                        if (seen) continue;
                        int b = s;   // idx_max
                        for (int q = s + 1; q < s_hi; ++q)
                            if (gauss_member(cx, cy, cz, q, xi, yi, zi, Nk)) b = q;
                        const int n = b - s + 1;
                        if (n < 2) continue;   // simps over one sample is 0
                        const double xv = tx[xi].x, yv = ty[yi].x;
                        const bool n_odd = n & 1;
                        double sum = 0.0;
                        for (int i = 0; i < n; ++i) {
                            const int q = s + i;
                            const double ex = xv - rp[q], ey = yv - rp[Ns + q], ez = zv - rp[2 * Ns + q];
                            const double r2 = ex * ex + ey * ey + ez * ez;
                            const double f = exp(r2 / minus_two_L2) * sigma2 * nep[q];
                            // abscissae s[q-2..q+2], clamped to the segment (entries outside it are never used)
                            const double wq = simpson_weight(i, n, n_odd, sp[max(q - 2, s)], sp[max(q - 1, s)], sp[q],
                                                             sp[min(q + 1, b)], sp[min(q + 2, b)]);
                            sum = fma(wq, f, sum);
                        }
                        atomicAdd(acc + ((long long)xi * ny + yi) * nz + zi, w * sum);
                    }
            }
        }
    }
}

extern "C" int iono_gaussian_adjoint_f64(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns,
                                         const double *ne_rays, const double *dd, double sigma_m, double L_m,
                                         int Nkernel, int zero_first, double *acc, void *stream) {
    const long long R = (long long)Na * Nt * Nd;
    if (!grid || !acc || Na < 0 || Nt < 0 || Nd < 0 || Ns < 1 || Nkernel < 0 || !(L_m > 0.0) ||
        (R > 0 && (!rays || !ne_rays || !dd)))
        return fail(IONO_EBADARG, "iono_gaussian_adjoint_f64: bad argument");
    if (sweep_size_check(grid, R, Ns)) return IONO_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (zero_first)
        CU_CHECK(cudaMemsetAsync(acc, 0, (size_t)grid->nx * grid->ny * grid->nz * sizeof(double), st));
    if (R == 0) return IONO_OK;
    const size_t smem = (size_t)8 * 3 * Ns * sizeof(int);
    if (smem > 200 * 1024) return fail(IONO_EBADARG, "iono_gaussian_adjoint_f64: Ns too large");
    CU_CHECK(cudaFuncSetAttribute(gaussian_adjoint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long ctas = (R + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    gaussian_adjoint_kernel<<<(int)(ctas < cap ? ctas : cap), 256, smem, st>>>(
        grid->dev, rays, (int)R, Ns, ne_rays, dd, sigma_m * sigma_m, -2.0 * L_m * L_m, Nkernel, acc);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}
