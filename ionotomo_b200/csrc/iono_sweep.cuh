// The ray sweep: forward (gather + warp reduce) and adjoint (scatter with fp64 reductions).
// Included by iono_kernels.cu (one translation unit).
//
// Mapping
//   * one warp per ray; a CTA's warps take consecutive work indices of the chosen
//     traversal order (IONO_ORDER_*), persistent CTAs (one per SM) walk the rest with a
//     static stride, so the result is bit-reproducible;
//   * the ray's samples are streamed HBM -> shared memory in chunks of C samples by TMA 1-D
//     bulk copies (cp.async.bulk + mbarrier complete_tx) into a warp-private ring; lane 0
//     is the producer, the 32 lanes are the consumers;
//   * lanes take consecutive samples of the chunk.  The grid is z-fastest and rays run
//     near-vertically, so the corner reads of neighbouring lanes share 32-byte sectors;
//     the corner values come through L1 (ld.global.nc) and the 67 MB grid stays in L2
//     while the ray stream passes with an evict-first policy.
#pragma once

#define IONO_MAX_FREQS 8

struct SweepParams {
    Grid g;
    const double *field;   // forward: ne (nx,ny,nz)
    const double4 *quads;  // forward, quad layout of ne (iono_device.cuh); used instead of `field` when LAYOUT == 1
    double *acc;           // adjoint: accumulator (nx,ny,nz)
    const double *rays;    // (R,4,Ns)
    const double *coef;    // adjoint: per-ray coefficient (R)
    double *tec;           // forward: per-ray integral (R)
    unsigned long long *oob_count;
    int R;
    int Ns;
    int stages;            // ring depth per warp
    RayOrder order;
    // phase-domain integrals (MODE 2): out_nf[ray*nf + f] = simps(g_f(ne(x_s)) [* field2(x_s)], s)
    const double *field2;  // optional second grid (mu_prior - mu) for the prior penalty
    double *out_nf;
    int nf;
    double neg_inv_np[IONO_MAX_FREQS];   // -1/(1.2404e-2 nu_f^2)
    struct Freqs { double a[IONO_MAX_FREQS], s[IONO_MAX_FREQS]; };
};

// A stage holds x[C], y[C], z[C] and s[C+4] (two halo samples either side for the
// Simpson weights).
template <int C>
struct StageLayout {
    static constexpr int S_OFF = 3 * C;  // in doubles
    static constexpr int DOUBLES = 4 * C + 4;
    static constexpr int BYTES = ((DOUBLES * 8 + 127) / 128) * 128;
};

template <int C, bool BULK>
__device__ __forceinline__ void fill_stage(double *stage, uint64_t *bar, const double *ray, int Ns, int c0,
                                           int lane, uint64_t policy) {
    const int n_c = min(C, Ns - c0);
    const int s_lo = max(c0 - 2, 0), s_hi = min(c0 + C + 2, Ns);
    double *sdst = stage + StageLayout<C>::S_OFF + (s_lo - (c0 - 2));
    if (BULK) {
        if (elect_one()) {      // (callers are warp-converged here)
            mbar_expect_tx(bar, (uint32_t)((3 * n_c + (s_hi - s_lo)) * 8));
            bulk_g2s(stage, ray + c0, n_c * 8, bar, policy);
            bulk_g2s(stage + C, ray + Ns + c0, n_c * 8, bar, policy);
            bulk_g2s(stage + 2 * C, ray + 2 * Ns + c0, n_c * 8, bar, policy);
            bulk_g2s(sdst, ray + 3 * Ns + s_lo, (s_hi - s_lo) * 8, bar, policy);
        }
    } else {
        for (int i = lane; i < n_c; i += 32) {
            stage[i] = ld_stream(ray + c0 + i, policy);
            stage[C + i] = ld_stream(ray + Ns + c0 + i, policy);
            stage[2 * C + i] = ld_stream(ray + 2 * Ns + c0 + i, policy);
        }
        for (int i = lane; i < s_hi - s_lo; i += 32) sdst[i] = ld_stream(ray + 3 * Ns + s_lo + i, policy);
    }
}

// Axis constants held in registers inside the sweep.
struct AxisR {
    double inv_d, c_guess, g0, glast;
    int nm2;
};
__device__ __forceinline__ AxisR axis_regs(const Axis &a) {
    AxisR r;
    r.inv_d = a.inv_d; r.c_guess = a.c_guess; r.g0 = a.g0; r.glast = a.glast; r.nm2 = a.n - 2;
    return r;
}

// Same semantics as iono::locate(), split into the common case (direct index, one table read,
// t by multiplication) and a rarely taken repair step; table in shared memory, constants in
// registers.
template <bool UNIFORM>
__device__ __forceinline__ void locate_fast(const double2 *__restrict__ tab, const AxisR &a, double x, int &i,
                                            double &t) {
    if (UNIFORM) {
        const double v = fma(x, a.inv_d, a.c_guess);
        // unsigned min also catches negative guesses (they wrap to huge values); the repair step
        // walks back from the last cell in that (out-of-bounds) case
        i = (int)min((unsigned int)__double2loint(v + IONO_MAGIC), (unsigned int)a.nm2);
    } else {
        int lo = 0, hi = a.nm2 + 1;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (x >= tab[mid].x) lo = mid; else hi = mid;
        }
        i = min(lo, a.nm2);
    }
    const double2 e = tab[i];
    t = (x - e.x) * e.y;
}
// t in [0,1)  <=>  the high word, read as unsigned, is below that of 1.0 (negative numbers and
// NaN have larger high words)
__device__ __forceinline__ bool in_unit(double t) { return (unsigned int)__double2hiint(t) < 0x3FF00000u; }

__device__ __forceinline__ void locate_repair(const double2 *__restrict__ tab, int nm2, double g0, double glast,
                                           double x, int &i, double &t, bool &oob) {
    while (i > 0 && x < tab[i].x) --i;
    while (i < nm2 && x >= tab[i + 1].x) ++i;
    const double2 e = tab[i];
    t = (x - e.x) * e.y;
    oob = oob || !(x >= g0 && x <= glast);
}

// Exact-linspace axis (AXK == 2): no table read.  v = (x - g0)/d - 1/2 is rounded to the nearest integer
// by the magic-number add, t = (v - round(v)) + 1/2.  `ir` is the unclamped index.
__device__ __forceinline__ void locate_arith(const AxisR &a, double x, int &ir, double &t) {
    const double v = fma(x, a.inv_d, a.c_guess);
    const double r = v + IONO_MAGIC;
    ir = __double2loint(r);
    t = (v - (r - IONO_MAGIC)) + 0.5;
}
// 0 < t < 1 (strict; -0, +0, negatives, 1 and NaN fail): the fast path of locate_arith needs no check
__device__ __forceinline__ bool in_open_unit(double t) {
    return (unsigned int)__double2hiint(t) - 1u < 0x3FF00000u - 1u;
}
__device__ __forceinline__ void locate_arith_repair(const AxisR &a, double x, int ir, int &i, double &t, bool &oob) {
    i = min(max(ir, 0), a.nm2);
    if (i != ir) t = (fma(x, a.inv_d, a.c_guess) - (double)i) + 0.5;
    oob = oob || !(x >= a.g0 && x <= a.glast);
    if (!(x == x)) i = 0;   // NaN: any valid cell, the sample is counted as out of bounds
}

// AXK: 0 = bisection + table, 1 = direct index + table (nodes within a quarter cell of a linspace),
//      2 = exact linspace, arithmetic only (no shared-memory table reads in the hot loop).
// Cell (ix,iy,iz) and in-cell coordinates of one sample; returns true if the sample is outside the grid.
template <int AXK>
__device__ __forceinline__ bool locate3(const double2 *__restrict__ tabx, const double2 *__restrict__ taby,
                                        const double2 *__restrict__ tabz, const AxisR &ax, const AxisR &ay,
                                        const AxisR &az, double px, double py, double pz, int &ix, int &iy, int &iz,
                                        double &tx, double &ty, double &tz) {
    bool oob = false;
    if (AXK == 2) {
        locate_arith(ax, px, ix, tx);
        locate_arith(ay, py, iy, ty);
        locate_arith(az, pz, iz, tz);
        const bool fast = in_open_unit(tx) & in_open_unit(ty) & in_open_unit(tz) &
                          ((unsigned int)ix <= (unsigned int)ax.nm2) & ((unsigned int)iy <= (unsigned int)ay.nm2) &
                          ((unsigned int)iz <= (unsigned int)az.nm2);
        if (!fast) {   // on a node, at/over the grid faces, or NaN
            const int rx = ix, ry = iy, rz = iz;
            locate_arith_repair(ax, px, rx, ix, tx, oob);
            locate_arith_repair(ay, py, ry, iy, ty, oob);
            locate_arith_repair(az, pz, rz, iz, tz, oob);
        }
    } else {
        locate_fast<AXK == 1>(tabx, ax, px, ix, tx);
        locate_fast<AXK == 1>(taby, ay, py, iy, ty);
        locate_fast<AXK == 1>(tabz, az, pz, iz, tz);
        if (!(in_unit(tx) & in_unit(ty) & in_unit(tz))) {
            // rare: a sample on a cell edge with the guess one off, on the last node, outside the grid, or NaN
            locate_repair(tabx, ax.nm2, ax.g0, ax.glast, px, ix, tx, oob);
            locate_repair(taby, ay.nm2, ay.g0, ay.glast, py, iy, ty, oob);
            locate_repair(tabz, az.nm2, az.g0, az.glast, pz, iz, tz, oob);
        }
    }
    return oob;
}

// Adjoint scatter of one sample per lane: a * (trilinear hat weights) into the 8 corners of
// cell v.  Lanes hold consecutive samples of one ray, so the cell of lane l+1 is usually the
// one directly above the cell of lane l (v+1: same column, next z node); then the upper-node
// share of lane l and the lower-node share of lane l+1 hit the same four addresses.  They
// are combined with a shuffle so that the pair costs one fp64 reduction instead of two.
__device__ __forceinline__ void scatter_sample(double *__restrict__ accp, int v, int sy, int sx, double a,
                                               double tx, double ty, double tz, int lane, bool valid) {
    if (!valid) { a = 0.0; v = -2; }
    const double ax1 = a * tx, ax0 = a - ax1;
    const double a01 = ax0 * ty, a00 = ax0 - a01;
    const double a11 = ax1 * ty, a10 = ax1 - a11;
    const double h00 = a00 * tz, h01 = a01 * tz, h10 = a10 * tz, h11 = a11 * tz;
    double l00 = a00 - h00, l01 = a01 - h01, l10 = a10 - h10, l11 = a11 - h11;
    const unsigned full = 0xffffffffu;
    const int v_next = __shfl_down_sync(full, v, 1);
    const int v_prev = __shfl_up_sync(full, v, 1);
    const bool give = (lane < 31) && (v_next == v + 1);      // my upper share goes to lane+1
    const bool take = (lane > 0) && (v_prev == v - 1);       // I absorb lane-1's upper share
    const double p00 = __shfl_up_sync(full, h00, 1), p01 = __shfl_up_sync(full, h01, 1);
    const double p10 = __shfl_up_sync(full, h10, 1), p11 = __shfl_up_sync(full, h11, 1);
    if (take) { l00 += p00; l01 += p01; l10 += p10; l11 += p11; }
    if (valid) {
        double *c = accp + v;
        atomicAdd(c, l00); atomicAdd(c + sy, l01); atomicAdd(c + sx, l10); atomicAdd(c + sx + sy, l11);
        if (!give) {
            atomicAdd(c + 1, h00); atomicAdd(c + sy + 1, h01);
            atomicAdd(c + sx + 1, h10); atomicAdd(c + sx + sy + 1, h11);
        }
    }
}

__device__ __forceinline__ double trilerp(const double *__restrict__ c, int sy, int sx, double tx, double ty,
                                          double tz) {
    const double v000 = __ldg(c), v001 = __ldg(c + 1);
    const double v010 = __ldg(c + sy), v011 = __ldg(c + sy + 1);
    const double v100 = __ldg(c + sx), v101 = __ldg(c + sx + 1);
    const double v110 = __ldg(c + sx + sy), v111 = __ldg(c + sx + sy + 1);
    const double c00 = fma(tz, v001 - v000, v000), c01 = fma(tz, v011 - v010, v010);
    const double c10 = fma(tz, v101 - v100, v100), c11 = fma(tz, v111 - v110, v110);
    const double c0_ = fma(ty, c01 - c00, c00), c1_ = fma(ty, c11 - c10, c10);
    return fma(tx, c1_ - c0_, c0_);
}

// same arithmetic as trilerp() on the quad layout: two 256-bit loads
__device__ __forceinline__ double trilerp_quads(const double4 *__restrict__ q, int sxq, double tx, double ty,
                                                double tz) {
    double v000, v001, v010, v011, v100, v101, v110, v111;
    ld_quad(q, v000, v001, v010, v011);
    ld_quad(q + sxq, v100, v101, v110, v111);
    const double c00 = fma(tz, v001 - v000, v000), c01 = fma(tz, v011 - v010, v010);
    const double c10 = fma(tz, v101 - v100, v100), c11 = fma(tz, v111 - v110, v110);
    const double c0_ = fma(ty, c01 - c00, c00), c1_ = fma(ty, c11 - c10, c10);
    return fma(tx, c1_ - c0_, c0_);
}

// MODE 0: TEC forward, MODE 1: adjoint scatter, MODE 2: phase-domain integrals per frequency
// (inversion/iterative_newton.py:108-119 and :157-179)
// LAYOUT 0: plain (nx,ny,nz) field, 1: quad records (MODE 0 only)
template <int MODE, int AXK, int C, bool BULK, int MAXT, int LAYOUT>
__global__ void __launch_bounds__(MAXT, 1) ray_sweep_kernel(const SweepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // the warp index through a shuffle: the compiler then knows it is warp-uniform, keeps everything derived from it
    // (ring and barrier addresses, the producer's source pointers) in uniform registers and issues the bulk copies
    // without a per-lane address loop
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), nwarp = blockDim.x >> 5;
    const int nx = p.g.ax[0].n, ny = p.g.ax[1].n, nz = p.g.ax[2].n;
    const int stages = p.stages;

    // shared: [axis tables][per-warp mbarriers][per-warp stages]
    // (no axis tables in shared memory when the in-cell coordinates come from arithmetic, AXK == 2)
    double2 *tabx = reinterpret_cast<double2 *>(smem_raw);
    double2 *taby = tabx + nx;
    double2 *tabz = taby + ny;
    unsigned int off = (AXK == 2) ? 0u : ((unsigned int)(nx + ny + nz) * 16u + 127u) / 128u * 128u;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + off) + warp * stages;
    off += ((unsigned int)(nwarp * stages) * 8u + 127u) / 128u * 128u;
    unsigned char *ring = smem_raw + off + (unsigned int)(warp * stages) * StageLayout<C>::BYTES;

    if (AXK != 2) {
        for (int i = threadIdx.x; i < nx; i += blockDim.x) tabx[i] = p.g.ax[0].tab[i];
        for (int i = threadIdx.x; i < ny; i += blockDim.x) taby[i] = p.g.ax[1].tab[i];
        for (int i = threadIdx.x; i < nz; i += blockDim.x) tabz[i] = p.g.ax[2].tab[i];
    }
    if (BULK && lane == 0)
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
    if (BULK) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint64_t pol_stream = policy_evict_first();
    const AxisR ax = axis_regs(p.g.ax[0]), ay = axis_regs(p.g.ax[1]), az = axis_regs(p.g.ax[2]);
    const int Ns = p.Ns;
    const bool n_odd = Ns & 1;
    const int chunks = (Ns + C - 1) / C;
    const int G = gridDim.x;
    const double *const rays = p.rays;
    const long long ray_doubles = 4LL * Ns;
    const RayOrder order = p.order;

    // Static schedule: bundle b = nwarp consecutive work indices; this CTA takes
    // b = blockIdx.x, blockIdx.x + G, ...; this warp takes work index b*nwarp + warp.
    const int n_bundles = (p.R + nwarp - 1) / nwarp;
    int my_rays = (n_bundles > (int)blockIdx.x) ? (n_bundles - 1 - (int)blockIdx.x) / G + 1 : 0;
    if (my_rays > 0 && ((int)blockIdx.x + (my_rays - 1) * G) * nwarp + warp >= p.R) --my_rays;
    // work index of this warp's k-th ray: q_k = (blockIdx.x + k*G)*nwarp + warp; its (i0,i1,i2)
    // digits in the traversal order are advanced incrementally (no divisions in the loop)
    const int q0 = (int)blockIdx.x * nwarp + warp, qstep = G * nwarp;
    const int d0 = qstep % order.n0, d1 = (qstep / order.n0) % order.n1, d2 = qstep / (order.n0 * order.n1);
    struct Cursor { int i0, i1, i2; };
    auto cursor_init = [&]() { Cursor c; c.i0 = q0 % order.n0; c.i1 = (q0 / order.n0) % order.n1; c.i2 = q0 / (order.n0 * order.n1); return c; };
    auto cursor_ray = [&](const Cursor &c) -> long long { return (long long)(c.i0 * order.st0 + c.i1 * order.st1 + c.i2 * order.st2); };
    auto cursor_next = [&](Cursor &c) {
        c.i0 += d0; if (c.i0 >= order.n0) { c.i0 -= order.n0; ++c.i1; }
        c.i1 += d1; if (c.i1 >= order.n1) { c.i1 -= order.n1; ++c.i2; }
        c.i2 += d2;
    };
    Cursor pc = cursor_init(), cc = pc;   // producer / consumer

    // producer cursor (ray fk, chunk fc, stage fs) and consumer stage/phase
    int fk = 0, fc = 0, fs = 0;
    const double *fray = rays + ((my_rays > 0) ? cursor_ray(pc) * ray_doubles : 0);
    auto produce = [&]() {
        if (fk < my_rays) {
            fill_stage<C, true>(reinterpret_cast<double *>(ring + fs * StageLayout<C>::BYTES), &bars[fs], fray, Ns,
                                fc * C, lane, pol_stream);
            fs = (fs + 1 == stages) ? 0 : fs + 1;
            if (++fc == chunks) {
                fc = 0;
                if (++fk < my_rays) { cursor_next(pc); fray = rays + cursor_ray(pc) * ray_doubles; }
            }
        }
    };
    if (BULK)
        for (int s = 0; s < stages - 1; ++s) produce();

    unsigned int n_oob = 0;
    unsigned int phases = 0;   // bit s: parity to wait for on stage s
    int us = 0;                // stage to consume
    const int sy = nz, sx = ny * nz;   // element strides (nx*ny*nz < 2^31 checked on the host)

    for (int k = 0; k < my_rays; ++k) {
        if (k > 0) cursor_next(cc);
        const long long ray = cursor_ray(cc);
        const double *rayp = rays + ray * ray_doubles;
        double acc = 0.0;
        double coef = 0.0;
        double accf[IONO_MAX_FREQS];
        if (MODE == 2) {
#pragma unroll
            for (int f = 0; f < IONO_MAX_FREQS; ++f) accf[f] = 0.0;
        }
        if (MODE == 1) coef = __ldg(p.coef + ray);
        for (int chunk = 0; chunk < chunks; ++chunk) {
            const int c0 = chunk * C;
            double *stage = reinterpret_cast<double *>(ring + us * StageLayout<C>::BYTES);
            if (BULK) {
                produce();
                mbar_wait(&bars[us], (phases >> us) & 1u);
                phases ^= 1u << us;
            } else {
                fill_stage<C, false>(stage, nullptr, rayp, Ns, c0, lane, pol_stream);
                __syncwarp();
            }
            const double *sx_ = stage, *sy_ = stage + C, *sz_ = stage + 2 * C;
            const double *ss_ = stage + StageLayout<C>::S_OFF + 2;   // ss_[j] = s[c0 + j]
            const int n_c = min(C, Ns - c0);
#ifndef IONO_SIMPSON_SHFL
            if (MODE == 0 && n_c == C && (C % 64) == 0) {
                // full chunk of the forward: two samples per lane without per-sample guards, so both samples' corner
                // gathers are in flight before the first interpolation needs its values (same summation order)
                for (int jb = 0; jb < C; jb += 64) {
                    double f[2], w[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int j = jb + 32 * u + lane;
                        int ix, iy, iz;
                        double tx, ty, tz;
                        n_oob += locate3<AXK>(tabx, taby, tabz, ax, ay, az, sx_[j], sy_[j], sz_[j], ix, iy, iz, tx, ty, tz);
                        w[u] = simpson_weight(c0 + j, Ns, n_odd, ss_[j - 2], ss_[j - 1], ss_[j], ss_[j + 1], ss_[j + 2]);
                        const int v = (ix * ny + iy) * nz + iz;
                        if (LAYOUT == 1) f[u] = trilerp_quads(p.quads + v, sx, tx, ty, tz);
                        else f[u] = trilerp(p.field + v, sy, sx, tx, ty, tz);
                    }
                    acc = fma(w[0], f[0], acc);
                    acc = fma(w[1], f[1], acc);
                }
            } else
#endif
#pragma unroll 2
            for (int jb = 0; jb < n_c; jb += 32) {
                const int j = jb + lane;
                const bool valid = j < n_c;
#ifdef IONO_SIMPSON_SHFL
                // abscissae s[i-2..i+2]: one shared-memory read per lane, the neighbours by shuffle, the halo
                // (lanes 0, 1, 30, 31) by predicated reads -- 6 instead of 10 shared-memory wavefronts
                const double s_0 = ss_[j];
                double s_m1 = __shfl_up_sync(0xffffffffu, s_0, 1), s_m2 = __shfl_up_sync(0xffffffffu, s_0, 2);
                double s_p1 = __shfl_down_sync(0xffffffffu, s_0, 1), s_p2 = __shfl_down_sync(0xffffffffu, s_0, 2);
                if (lane < 2) { s_m2 = ss_[j - 2]; if (lane < 1) s_m1 = ss_[j - 1]; }
                if (lane > 29) { s_p2 = ss_[j + 2]; if (lane > 30) s_p1 = ss_[j + 1]; }
#endif
                if (MODE != 1 && !valid) continue;
                const int i = c0 + j;
                int ix, iy, iz;
                double tx, ty, tz;
                const double px = sx_[j], py = sy_[j], pz = sz_[j];
                const bool oob = locate3<AXK>(tabx, taby, tabz, ax, ay, az, px, py, pz, ix, iy, iz, tx, ty, tz);
                n_oob += (oob && valid);
#ifdef IONO_SIMPSON_SHFL
                const double w = simpson_weight(i, Ns, n_odd, s_m2, s_m1, s_0, s_p1, s_p2);
#else
                const double w = simpson_weight(i, Ns, n_odd, ss_[j - 2], ss_[j - 1], ss_[j], ss_[j + 1], ss_[j + 2]);
#endif
                const int v = (ix * ny + iy) * nz + iz;
                if (MODE == 0) {
                    if (LAYOUT == 1) acc = fma(w, trilerp_quads(p.quads + v, sx, tx, ty, tz), acc);
                    else acc = fma(w, trilerp(p.field + v, sy, sx, tx, ty, tz), acc);
                } else if (MODE == 2) {
                    const double ne_s = trilerp(p.field + v, sy, sx, tx, ty, tz);
                    const bool penalty = p.field2 != nullptr;
                    const double dmu_s = penalty ? trilerp(p.field2 + v, sy, sx, tx, ty, tz) : 0.0;
#pragma unroll
                    for (int f = 0; f < IONO_MAX_FREQS; ++f)
                        if (f < p.nf) {
                            // n = sqrt(1 - ne/n_p);  integrand (1 - n)  or  (ne/n) * dmu
                            const double n = sqrt(fma(ne_s, p.neg_inv_np[f], 1.0));
                            const double g = penalty ? (ne_s / n) * dmu_s : (1.0 - n);
                            accf[f] = fma(w, g, accf[f]);
                        }
                } else {
                    scatter_sample(p.acc, v, sy, sx, coef * w, tx, ty, tz, lane, valid);
                }
            }
            us = (us + 1 == stages) ? 0 : us + 1;
            __syncwarp();
        }
        if (MODE == 0) {
            const double tot = warp_sum(acc);
            if (lane == 0) p.tec[ray] = tot;
        }
        if (MODE == 2) {
#pragma unroll
            for (int f = 0; f < IONO_MAX_FREQS; ++f)
                if (f < p.nf) {
                    const double tot = warp_sum(accf[f]);
                    if (lane == 0) p.out_nf[ray * p.nf + f] = tot;
                }
        }
    }
    if (n_oob) atomicAdd(p.oob_count, (unsigned long long)n_oob);
}

struct SweepConfig {
    int warps;    // per CTA
    int stages;
    int chunk;    // 64 or 128
};

static SweepConfig sweep_config(int mode, int Ns) {
    SweepConfig c;
    c.warps = (mode == 0) ? 24 : 16;   // phase mode carries 8 accumulators: 16 warps
    c.stages = 2;
    c.chunk = (Ns <= 64) ? 64 : 128;
    const char *e;
    if ((e = getenv("IONO_SWEEP_WARPS"))) c.warps = atoi(e);
    if ((e = getenv("IONO_SWEEP_STAGES"))) c.stages = atoi(e);
    if ((e = getenv("IONO_SWEEP_CHUNK"))) c.chunk = atoi(e);
    if (c.warps < 1) c.warps = 1;
    if (c.warps > 24) c.warps = 24;
    if (mode == 2 && c.warps > 16) c.warps = 16;
    if (c.stages < 2) c.stages = 2;
    if (c.stages > 8) c.stages = 8;
    if (c.chunk != 64) c.chunk = 128;
    return c;
}

static RayOrder make_order(int order, int Na, int Nt, int Nd) {
    RayOrder o;
    const int sa = Nt * Nd, st = Nd, sd = 1;
    switch (order) {
        case IONO_ORDER_TIME:      // t fastest, then a, then d
            o.n0 = Nt; o.st0 = st; o.n1 = Na; o.st1 = sa; o.n2 = Nd; o.st2 = sd; break;
        case IONO_ORDER_ANTENNA:   // a fastest, then t, then d
            o.n0 = Na; o.st0 = sa; o.n1 = Nt; o.st1 = st; o.n2 = Nd; o.st2 = sd; break;
        default:                   // memory order
            o.n0 = Nd; o.st0 = sd; o.n1 = Nt; o.st1 = st; o.n2 = Na; o.st2 = sa; break;
    }
    return o;
}

template <int MODE, int AXK, int C, bool BULK, int MAXT, int LAYOUT>
static int launch_sweep_t(const SweepParams &p, const SweepConfig &cfg, size_t smem, int ctas, cudaStream_t st) {
    auto kern = ray_sweep_kernel<MODE, AXK, C, BULK, MAXT, LAYOUT>;
    CU_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ctas, cfg.warps * 32, smem, st>>>(p);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

template <int MODE, int AXK, int LAYOUT>
static int launch_sweep_a(const SweepParams &p, const SweepConfig &cfg, size_t smem, int ctas, bool bulk,
                          cudaStream_t st) {
    const bool wide = (MODE != 2 && cfg.warps > 16);
    if (cfg.chunk == 64) {
        if (bulk) return wide ? launch_sweep_t<MODE, AXK, 64, true, 768, LAYOUT>(p, cfg, smem, ctas, st)
                              : launch_sweep_t<MODE, AXK, 64, true, 512, LAYOUT>(p, cfg, smem, ctas, st);
        return wide ? launch_sweep_t<MODE, AXK, 64, false, 768, LAYOUT>(p, cfg, smem, ctas, st)
                    : launch_sweep_t<MODE, AXK, 64, false, 512, LAYOUT>(p, cfg, smem, ctas, st);
    }
    if (bulk) return wide ? launch_sweep_t<MODE, AXK, 128, true, 768, LAYOUT>(p, cfg, smem, ctas, st)
                          : launch_sweep_t<MODE, AXK, 128, true, 512, LAYOUT>(p, cfg, smem, ctas, st);
    return wide ? launch_sweep_t<MODE, AXK, 128, false, 768, LAYOUT>(p, cfg, smem, ctas, st)
                : launch_sweep_t<MODE, AXK, 128, false, 512, LAYOUT>(p, cfg, smem, ctas, st);
}

// LAYOUT 1 (quad records in p.quads) is offered for MODE 0 only
template <int MODE, int LAYOUT>
static int launch_sweep(SweepParams p, iono_grid_t grid, cudaStream_t st) {
    if (device_check(grid->device, "ray sweep")) return IONO_EBADARG;
    SweepConfig cfg = sweep_config(MODE, p.Ns);
    p.stages = cfg.stages;
    const int axk = grid->exact ? 2 : (grid->uniform ? 1 : 0);
    const size_t stage_bytes = cfg.chunk == 64 ? StageLayout<64>::BYTES : StageLayout<128>::BYTES;
    const size_t table_bytes =
        (axk == 2) ? 0 : (((size_t)(grid->nx + grid->ny + grid->nz) * sizeof(double2)) + 127) / 128 * 128;
    auto smem_for = [&](int warps) {
        return table_bytes + (((size_t)warps * cfg.stages * sizeof(uint64_t)) + 127) / 128 * 128 +
               (size_t)warps * cfg.stages * stage_bytes;
    };
    while (cfg.warps > 4 && smem_for(cfg.warps) > 227 * 1024) cfg.warps -= 4;   // very large axis tables: fewer warps
    const size_t smem = smem_for(cfg.warps);
    if (smem > 227 * 1024) return fail(IONO_EBADARG, "ray sweep: shared-memory configuration exceeds 227 KB");
    // TMA bulk copies need 16-byte aligned rows: even Ns and a 16-byte aligned base
    const bool bulk = (p.Ns % 2 == 0) && (((uintptr_t)p.rays & 15) == 0) && !getenv("IONO_SWEEP_NO_BULK");
    const int n_bundles = (p.R + cfg.warps - 1) / cfg.warps;
    int ctas = sm_count();
    if (ctas > n_bundles) ctas = n_bundles;
    if (axk == 2) return launch_sweep_a<MODE, 2, LAYOUT>(p, cfg, smem, ctas, bulk, st);
    if (axk == 1) return launch_sweep_a<MODE, 1, LAYOUT>(p, cfg, smem, ctas, bulk, st);
    return launch_sweep_a<MODE, 0, LAYOUT>(p, cfg, smem, ctas, bulk, st);
}

static int sweep_size_check(iono_grid_t grid, long long R, int Ns) {
    if (R * 4LL * Ns / 4 > 0x7fffffffLL * 1024LL) return fail(IONO_EBADARG, "ray sweep: ray array too large");
    if (R > 0x7fffffffLL - 1024) return fail(IONO_EBADARG, "ray sweep: more than 2^31 rays");
    if ((long long)grid->nx * grid->ny * grid->nz > 0x7fffffffLL)
        return fail(IONO_EBADARG, "ray sweep: more than 2^31 voxels");
    return IONO_OK;
}

// Policy of the plain-layout entry points: rewrite ne as quad records into a stream-ordered temporary when the
// quad grid stays <= 512 MB and the gathers that follow are long enough to pay for it (the rewrite moves
// 5 x the grid; IONO_FWD_LAYOUT=plain|quads overrides).  Returns NULL for "use the plain layout".
static double4 *quads_temporary(const double *ne, int nx, int ny, int nz, long long samples, cudaStream_t st) {
    const long long V = (long long)nx * ny * nz;
    bool use_quads = (V <= (1LL << 24)) && (samples >= 4 * V);
    if (const char *e = getenv("IONO_FWD_LAYOUT")) {
        if (!strcmp(e, "plain")) use_quads = false;
        else if (!strcmp(e, "quads")) use_quads = true;
    }
    if (!use_quads) return nullptr;
    double4 *q = nullptr;
    if (cudaMallocAsync((void **)&q, (size_t)V * sizeof(double4), st) != cudaSuccess) {
        cudaGetLastError();   // no memory for the temporary: the plain layout works too
        return nullptr;
    }
    quads_kernel<<<ew_grid(V), 256, 0, st>>>(ne, nullptr, 1.0, ny, nz, V, nullptr, q);
    return q;
}

static int tec_forward_common(iono_grid_t grid, const double *ne, const double4 *quads, const double *rays, int Na,
                              int Nt, int Nd, int Ns, int order, double *tec_out, unsigned long long *oob_count,
                              cudaStream_t st, const char *what) {
    const long long R = (long long)Na * Nt * Nd;
    if (!grid || (!ne && !quads) || !oob_count || Na < 0 || Nt < 0 || Nd < 0 || Ns < 1 || (R > 0 && (!rays || !tec_out)))
        return fail(IONO_EBADARG, "%s: bad argument", what);
    if (sweep_size_check(grid, R, Ns)) return IONO_EBADARG;
    CU_CHECK(cudaMemsetAsync(oob_count, 0, sizeof(unsigned long long), st));
    if (R == 0) return IONO_OK;
    if (Ns < 2) {  // simps of a single sample is 0
        CU_CHECK(cudaMemsetAsync(tec_out, 0, R * sizeof(double), st));
        return IONO_OK;
    }
    SweepParams p;
    memset(&p, 0, sizeof(p));
    p.g = grid->dev; p.field = ne; p.quads = quads; p.rays = rays; p.tec = tec_out; p.oob_count = oob_count;
    p.R = (int)R; p.Ns = Ns; p.order = make_order(order, Na, Nt, Nd);
    return quads ? launch_sweep<0, 1>(p, grid, st) : launch_sweep<0, 0>(p, grid, st);
}

// Forward on the quad layout of ne (iono_quads_from_ne_f64 / iono_ne_quads_from_m_f64).
extern "C" int iono_tec_forward_quads_f64(iono_grid_t grid, const double *quads, const double *rays, int Na, int Nt,
                                          int Nd, int Ns, int order, double *tec_out, unsigned long long *oob_count,
                                          void *stream) {
    if ((uintptr_t)quads & 31) return fail(IONO_EBADARG, "iono_tec_forward_quads_f64: quads must be 32-byte aligned");
    return tec_forward_common(grid, nullptr, reinterpret_cast<const double4 *>(quads), rays, Na, Nt, Nd, Ns, order,
                              tec_out, oob_count, (cudaStream_t)stream, "iono_tec_forward_quads_f64");
}

// Plain-layout entry point.  When the sweep is long enough to pay for it (IONO_FWD_LAYOUT=auto, the default),
// ne is first rewritten as quad records into a stream-ordered temporary (4 x the grid, cudaMallocAsync) and the
// sweep gathers from those; IONO_FWD_LAYOUT=plain keeps the 8 scalar corner loads.
extern "C" int iono_tec_forward_f64(iono_grid_t grid, const double *ne, const double *rays, int Na, int Nt, int Nd,
                                    int Ns, int order, double *tec_out, unsigned long long *oob_count,
                                    void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const long long R = (long long)Na * Nt * Nd;
    if (grid && ne && R > 0 && Ns >= 2) {
        double4 *q = quads_temporary(ne, grid->nx, grid->ny, grid->nz, R * Ns, st);
        if (q) {
            int rc = tec_forward_common(grid, nullptr, q, rays, Na, Nt, Nd, Ns, order, tec_out, oob_count, st,
                                        "iono_tec_forward_f64");
            cudaFreeAsync(q, st);
            return rc;
        }
    }
    return tec_forward_common(grid, ne, nullptr, rays, Na, Nt, Nd, Ns, order, tec_out, oob_count, st,
                              "iono_tec_forward_f64");
}

static int launch_adjoint_runs(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns, const double *coef,
                               double *acc, unsigned long long *oob_count, cudaStream_t st);

extern "C" int iono_tec_adjoint_f64(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns,
                                    const double *coef, int order, int zero_first, double *acc,
                                    unsigned long long *oob_count, void *stream) {
    const long long R = (long long)Na * Nt * Nd;
    if (!grid || !acc || !oob_count || Na < 0 || Nt < 0 || Nd < 0 || Ns < 1 || (R > 0 && (!rays || !coef)))
        return fail(IONO_EBADARG, "iono_tec_adjoint_f64: bad argument");
    if (sweep_size_check(grid, R, Ns)) return IONO_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(oob_count, 0, sizeof(unsigned long long), st));
    if (zero_first)
        CU_CHECK(cudaMemsetAsync(acc, 0, (size_t)grid->nx * grid->ny * grid->nz * sizeof(double), st));
    if (R == 0 || Ns < 2) return IONO_OK;
    if (device_check(grid->device, "iono_tec_adjoint_f64")) return IONO_EBADARG;
    // time-bundled scatter with run aggregation (iono_adjoint_runs.cuh) whenever there is a time axis to walk
    const int rc = launch_adjoint_runs(grid, rays, Na, Nt, Nd, Ns, coef, acc, oob_count, st);
    if (rc != -1) return rc;
    SweepParams p;
    memset(&p, 0, sizeof(p));
    p.g = grid->dev; p.acc = acc; p.rays = rays; p.coef = coef; p.oob_count = oob_count;
    p.R = (int)R; p.Ns = Ns; p.order = make_order(order, Na, Nt, Nd);
    return launch_sweep<1, 0>(p, grid, st);
}

// ---------------------------------------------------------------------------
// phase-domain ray integrals (generation B of the reference)
// ---------------------------------------------------------------------------
extern "C" int iono_phase_integrals_f64(iono_grid_t grid, const double *ne, const double *dmu, const double *rays,
                                        int Na, int Nt, int Nd, int Ns, const double *freqs_host, int Nf, int order,
                                        double *out, unsigned long long *oob_count, void *stream) {
    const long long R = (long long)Na * Nt * Nd;
    if (!grid || !ne || !oob_count || !freqs_host || Na < 0 || Nt < 0 || Nd < 0 || Ns < 1 || Nf < 1 ||
        Nf > IONO_MAX_FREQS || (R > 0 && (!rays || !out)))
        return fail(IONO_EBADARG, "iono_phase_integrals_f64: bad argument (1 <= Nf <= 8)");
    if (sweep_size_check(grid, R, Ns)) return IONO_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(oob_count, 0, sizeof(unsigned long long), st));
    if (R == 0) return IONO_OK;
    if (Ns < 2) {
        CU_CHECK(cudaMemsetAsync(out, 0, R * Nf * sizeof(double), st));
        return IONO_OK;
    }
    SweepParams p;
    memset(&p, 0, sizeof(p));
    p.g = grid->dev; p.field = ne; p.field2 = dmu; p.rays = rays; p.out_nf = out; p.oob_count = oob_count;
    p.R = (int)R; p.Ns = Ns; p.nf = Nf; p.order = make_order(order, Na, Nt, Nd);
    for (int f = 0; f < Nf; ++f) p.neg_inv_np[f] = -1.0 / (1.2404e-2 * freqs_host[f] * freqs_host[f]);
    return launch_sweep<2, 0>(p, grid, st);
}

// out[a,t,d,f] = base[a,t,f] - scale[f] * (I[a,t,d,f] - I[i0,t,d,f]),
// base = const[a] + 2 pi nu_f clock[a,t] (phase) or 0 (penalty: pass clock = konst = NULL)
__global__ void __launch_bounds__(256) phase_assemble_kernel(const double *__restrict__ I, int Na, int Nt, int Nd,
                                                              int Nf, int i0, const double *__restrict__ clock,
                                                              const double *__restrict__ konst, SweepParams::Freqs fr,
                                                              double *__restrict__ out) {
    const long long n = (long long)Na * Nt * Nd * Nf;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long per_a = (long long)Nt * Nd * Nf;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const int f = (int)(k % Nf);
        const long long a = k / per_a, rem = k - a * per_a;
        const long long t = rem / ((long long)Nd * Nf);
        const double ref = I[(long long)i0 * per_a + rem];
        double base = 0.0;
        if (clock) base = konst[a] + fr.a[f] * clock[a * Nt + t];
        out[k] = base - fr.s[f] * (I[k] - ref);
    }
}

extern "C" int iono_phase_assemble_f64(const double *integrals, int Na, int Nt, int Nd, int Nf, int i0,
                                       const double *freqs_host, const double *clock, const double *konst,
                                       int penalty, double *out, void *stream) {
    const long long n = (long long)Na * Nt * Nd * Nf;
    if (Na < 0 || Nt < 0 || Nd < 0 || Nf < 1 || Nf > IONO_MAX_FREQS || !freqs_host)
        return fail(IONO_EBADARG, "iono_phase_assemble_f64: bad argument");
    if (n == 0) return IONO_OK;
    if (!integrals || !out || i0 < 0 || i0 >= Na || (!penalty && (!clock || !konst)))
        return fail(IONO_EBADARG, "iono_phase_assemble_f64: bad argument");
    SweepParams::Freqs fr;
    const double c = 299792458.0, two_pi = 6.283185307179586476925286766559;
    for (int f = 0; f < Nf; ++f) {
        const double a_ = two_pi * freqs_host[f];
        const double n_p = 1.2404e-2 * freqs_host[f] * freqs_host[f];
        fr.a[f] = a_;
        fr.s[f] = penalty ? a_ / (2 * n_p * c) : a_ / c;   // iterative_newton.py:166-181 / :110-121
    }
    phase_assemble_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(integrals, Na, Nt, Nd, Nf, i0,
                                                                        penalty ? nullptr : clock,
                                                                        penalty ? nullptr : konst, fr, out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

// ---------------------------------------------------------------------------
// Simpson integration of tabulated integrands along the rays' s rows: out[ray*out_stride] =
// simps(f(y[ray,:]), s[ray,:]) with scipy's old even='avg' rule (the per-sample weights of the sweep).
//   mode 0: f = y;   mode 1: f = 1 - sqrt(1 + y * c)   (phase integrand, c = -1/n_p);
//   mode 2: f = y / sqrt(1 + y * c) * y2               (prior-penalty integrand)
// Used by the reference-compatible phase operators, which interpolate first (TriCubic.interp on 4-D inputs,
// with the reference's axis scramble, geometry/tri_cubic.py:69-70) and integrate afterwards
// (iterative_newton.py:108-119, :157-179).  Warp per ray.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) simps_rows_kernel(const double *__restrict__ y, const double *__restrict__ y2,
                                                         const double *__restrict__ rays, long long R, int Ns,
                                                         int mode, double c, double *__restrict__ out,
                                                         int out_stride) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const bool n_odd = Ns & 1;
    for (long long ray = warp_global; ray < R; ray += n_warps) {
        const double *sp = rays + ray * 4 * Ns + 3 * (long long)Ns;
        const double *yp = y + ray * Ns;
        double acc = 0.0;
        for (int i = lane; i < Ns; i += 32) {
            const double w = simpson_weight(i, Ns, n_odd, sp[max(i - 2, 0)], sp[max(i - 1, 0)], sp[i],
                                            sp[min(i + 1, Ns - 1)], sp[min(i + 2, Ns - 1)]);
            double f = yp[i];
            if (mode == 1) f = 1.0 - sqrt(fma(f, c, 1.0));
            else if (mode == 2) f = f / sqrt(fma(f, c, 1.0)) * y2[ray * Ns + i];
            acc = fma(w, f, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) out[ray * out_stride] = acc;
    }
}

extern "C" int iono_simps_rows_f64(const double *y, const double *y2, const double *rays, int64_t nrays, int Ns,
                                   int mode, double c, double *out, int out_stride, void *stream) {
    if (nrays < 0 || Ns < 1 || mode < 0 || mode > 2 || out_stride < 1)
        return fail(IONO_EBADARG, "iono_simps_rows_f64: bad argument");
    if (nrays == 0) return IONO_OK;
    if (!y || !rays || !out || (mode == 2 && !y2)) return fail(IONO_EBADARG, "iono_simps_rows_f64: NULL pointer");
    if (Ns < 2) {   // simps of a single sample is 0
        CU_CHECK(cudaMemset2DAsync(out, (size_t)out_stride * 8, 0, 8, (size_t)nrays, (cudaStream_t)stream));
        return IONO_OK;
    }
    simps_rows_kernel<<<ew_grid(nrays * 32), 256, 0, (cudaStream_t)stream>>>(y, y2, rays, nrays, Ns, mode, c, out,
                                                                            out_stride);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}
