// Chord-length ("ray dirac") adjoint: the reference's generation-A gradient
// (inversion/gradient.py:15-20 over geometry/ray_dirac.py:5-34 + geometry/slab_method.py:19-58).
// Included by iono_kernels.cu.
//
//   acc[v] += dd[ray] * l(ray, v)
// where, per ray, the voxels v are the +-1 neighbourhoods of the cells of its samples (cell =
// the reference's `bisection`), and l is the chord of the straight line first->last sample
// point through the box centred on node v with the spacing of the first cell, accepted only if
// the entry parameter is > 0 and < exit (slab_method.py:55-58).  The reference ASSIGNS the
// chord per (ray, voxel), so every voxel counts once per ray however many samples reach it.
//
// Mapping: one warp per ray.  Cell indices of all samples go to shared memory; lanes take
// z-levels; for a z-monotone ray (everything cast_ray produces) the samples that can reach a
// level form a contiguous range found by binary search, otherwise all samples are scanned.
// A (voxel, ray) pair is emitted by the first sample that reaches it.  This is a compatibility
// kernel (the reference cannot run its dense version beyond toy sizes): fp64 atomics, no tuning.
#pragma once

__device__ __forceinline__ int ref_bisection(const double2 *__restrict__ tab, int n, double v) {
    // geometry/tri_cubic.py:105-132: -1 below, n above, n-1 on the last node, else j with a[j] <= v < a[j+1]
    if (v < tab[0].x) return -1;
    if (v > tab[n - 1].x) return n;
    if (v == tab[n - 1].x) return n - 1;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (v >= tab[mid].x) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void slab_axis(double lo, double hi, double o, double inv, double &tmin, double &tmax) {
    double t1 = (lo - o) * inv, t2 = (hi - o) * inv;
    if (t1 != t1) t1 = 0.0;   // NaN (0 * inf) -> 0, slab_method.py:23-26
    if (t2 != t2) t2 = 0.0;
    tmin = fmin(t1, t2);
    tmax = fmax(t1, t2);
}

__global__ void __launch_bounds__(256) chord_adjoint_kernel(Grid g, const double *__restrict__ rays, int R, int Ns,
                                                             const double *__restrict__ dd, double *__restrict__ acc) {
    extern __shared__ int chord_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int *cx = chord_smem + warp * 3 * Ns, *cy = cx + Ns, *cz = cy + Ns;
    const int nx = g.ax[0].n, ny = g.ax[1].n, nz = g.ax[2].n;
    const double2 *tx = g.ax[0].tab, *ty = g.ax[1].tab, *tz = g.ax[2].tab;
    const double hx = 0.5 * (tx[1].x - tx[0].x), hy = 0.5 * (ty[1].x - ty[0].x), hz = 0.5 * (tz[1].x - tz[0].x);
    for (int ray = blockIdx.x * nwarp + warp; ray < R; ray += gridDim.x * nwarp) {
        const double *rp = rays + (long long)ray * 4 * Ns;
        __syncwarp();
        int mono = 1;
        for (int s = lane; s < Ns; s += 32) {
            cx[s] = ref_bisection(tx, nx, rp[s]);
            cy[s] = ref_bisection(ty, ny, rp[Ns + s]);
            cz[s] = ref_bisection(tz, nz, rp[2 * Ns + s]);
        }
        __syncwarp();
        for (int s = lane + 1; s < Ns; s += 32) mono &= (cz[s] >= cz[s - 1]);
        mono = __all_sync(0xffffffffu, mono);
        // the line (geometry/ray_dirac.py:21, slab_method.py:10-15)
        const double ox = rp[0], oy = rp[Ns], oz = rp[2 * Ns];
        double nxd = rp[Ns - 1] - ox, nyd = rp[2 * Ns - 1] - oy, nzd = rp[3 * Ns - 1] - oz;
        const double nrm = sqrt(nxd * nxd + nyd * nyd + nzd * nzd);
        nxd /= nrm; nyd /= nrm; nzd /= nrm;
        const double ix_ = 1.0 / nxd, iy_ = 1.0 / nyd, iz_ = 1.0 / nzd;
        const double nlen = sqrt(nxd * nxd + nyd * nyd + nzd * nzd);
        const double w = dd[ray];
        for (int zi = lane; zi < nz; zi += 32) {
            int s_lo = 0, s_hi = Ns;
            if (mono) {   // samples with |cz - zi| <= 1
                int lo = 0, hi = Ns;
                while (lo < hi) { const int m = (lo + hi) >> 1; if (cz[m] < zi - 1) lo = m + 1; else hi = m; }
                s_lo = lo;
                hi = Ns;
                while (lo < hi) { const int m = (lo + hi) >> 1; if (cz[m] <= zi + 1) lo = m + 1; else hi = m; }
                s_hi = lo;
            }
            const double zc = tz[zi].x;
            double tzmin, tzmax;
            slab_axis(zc - hz, zc + hz, oz, iz_, tzmin, tzmax);
            for (int s = s_lo; s < s_hi; ++s) {
                const int dz = cz[s] - zi;
                if (dz < -1 || dz > 1) continue;
                const int xc = cx[s], yc = cy[s];
                for (int xi = max(0, xc - 1); xi < min(nx, xc + 2); ++xi)
                    for (int yi = max(0, yc - 1); yi < min(ny, yc + 2); ++yi) {
                        // already emitted by an earlier sample that reaches (xi, yi, zi)?
                        bool seen = false;
                        for (int q = s_lo; q < s && !seen; ++q) {
                            const int dq = cz[q] - zi;
                            seen = dq >= -1 && dq <= 1 && xi >= cx[q] - 1 && xi <= cx[q] + 1 && yi >= cy[q] - 1 &&
                                   yi <= cy[q] + 1;
                        }
                        if (seen) continue;
                        double txmin, txmax, tymin, tymax;
                        const double xcn = tx[xi].x, ycn = ty[yi].x;
                        slab_axis(xcn - hx, xcn + hx, ox, ix_, txmin, txmax);
                        slab_axis(ycn - hy, ycn + hy, oy, iy_, tymin, tymax);
                        const double t_in = fmax(fmax(txmin, tymin), tzmin);
                        const double t_out = fmin(fmin(txmax, tymax), tzmax);
                        if (t_in < t_out && t_in > 0.0)
                            atomicAdd(acc + ((long long)xi * ny + yi) * nz + zi, w * (nlen * (t_out - t_in)));
                    }
            }
        }
    }
}

extern "C" int iono_chord_adjoint_f64(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns,
                                      const double *dd, int zero_first, double *acc, void *stream) {
    const long long R = (long long)Na * Nt * Nd;
    if (!grid || !acc || Na < 0 || Nt < 0 || Nd < 0 || Ns < 1 || (R > 0 && (!rays || !dd)))
        return fail(IONO_EBADARG, "iono_chord_adjoint_f64: bad argument");
    if (sweep_size_check(grid, R, Ns)) return IONO_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (zero_first)
        CU_CHECK(cudaMemsetAsync(acc, 0, (size_t)grid->nx * grid->ny * grid->nz * sizeof(double), st));
    if (R == 0) return IONO_OK;
    const size_t smem = (size_t)8 * 3 * Ns * sizeof(int);
    if (smem > 200 * 1024) return fail(IONO_EBADARG, "iono_chord_adjoint_f64: Ns too large");
    CU_CHECK(cudaFuncSetAttribute(chord_adjoint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long ctas = (R + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    chord_adjoint_kernel<<<(int)(ctas < cap ? ctas : cap), 256, smem, st>>>(grid->dev, rays, (int)R, Ns, dd, acc);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}
