// libionob200: hand-written sm_100a kernels + C ABI (include/ionob200.h) for
// IonoTomo's ray-integral forward model and its adjoint.
//
// Hot kernels (ray_sweep_kernel<...>):
//   one warp per ray; the ray's samples (x,y,z,s rows of rays[Na,Nt,Nd,4,Ns]) are
//   streamed HBM -> shared memory by TMA 1-D bulk copies (cp.async.bulk +
//   mbarrier) into a warp-private ring, lanes take consecutive samples (the grid
//   is z-fastest and rays are near-vertical, so the 8 corner reads of
//   neighbouring lanes fall in the same 32-byte sectors), the Simpson-weighted
//   integrand is reduced along the ray with warp shuffles (forward) or scattered
//   with fp64 reductions red.global.add.f64 (adjoint).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/ionob200.h"
#include "iono_device.cuh"

using namespace iono;

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, const char *a = "", const char *b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
#define CU_CHECK(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) return fail(IONO_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

static int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

extern "C" int iono_version(void) { return IONO_ABI_VERSION; }
extern "C" const char *iono_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------
// grid handle
// ---------------------------------------------------------------------------
struct iono_grid {
    Grid dev;          // by-value kernel argument
    double2 *tables;   // one allocation, x|y|z
    int nx, ny, nz;
    int uniform;
    int device;
};

static bool axis_monotone(const double *g, int n) {
    for (int i = 1; i < n; ++i)
        if (!(g[i] > g[i - 1])) return false;
    return true;
}

// uniform <=> every node is within a quarter cell of the ideal linspace node, so the
// direct index guess is off by at most one cell (fixed up exactly in locate()).
static bool axis_uniform(const double *g, int n) {
    double d = (g[n - 1] - g[0]) / (n - 1);
    for (int i = 0; i < n; ++i)
        if (fabs(g[i] - (g[0] + i * d)) > 0.25 * d) return false;
    return true;
}

extern "C" int iono_grid_create(const double *xv, const double *yv, const double *zv, int nx, int ny, int nz,
                                iono_grid_t *out) {
    if (!xv || !yv || !zv || !out) return fail(IONO_EBADARG, "iono_grid_create: NULL argument");
    if (nx < 2 || ny < 2 || nz < 2) return fail(IONO_EBADARG, "iono_grid_create: each axis needs >= 2 nodes");
    const double *g[3] = {xv, yv, zv};
    const int n[3] = {nx, ny, nz};
    for (int a = 0; a < 3; ++a)
        if (!axis_monotone(g[a], n[a])) return fail(IONO_EBADARG, "iono_grid_create: axis not strictly increasing");
    iono_grid *h = new iono_grid();
    h->nx = nx; h->ny = ny; h->nz = nz;
    CU_CHECK(cudaGetDevice(&h->device));
    size_t total = (size_t)nx + ny + nz;
    std::vector<double2> host(total);
    size_t off = 0;
    cudaError_t e = cudaMalloc(&h->tables, total * sizeof(double2));
    if (e != cudaSuccess) { delete h; return fail(IONO_ECUDA, "cudaMalloc(grid tables): %s", cudaGetErrorString(e)); }
    h->uniform = 1;
    for (int a = 0; a < 3; ++a) {
        Axis &A = h->dev.ax[a];
        for (int i = 0; i < n[a]; ++i) {
            host[off + i].x = g[a][i];
            host[off + i].y = (i < n[a] - 1) ? 1.0 / (g[a][i + 1] - g[a][i]) : 0.0;
        }
        A.tab = h->tables + off;
        A.n = n[a];
        A.uniform = axis_uniform(g[a], n[a]) ? 1 : 0;
        A.inv_d = (n[a] - 1) / (g[a][n[a] - 1] - g[a][0]);
        A.c_guess = -g[a][0] * A.inv_d - 0.5;
        A.g0 = g[a][0];
        A.glast = g[a][n[a] - 1];
        h->uniform &= A.uniform;
        off += n[a];
    }
    e = cudaMemcpy(h->tables, host.data(), total * sizeof(double2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(h->tables); delete h; return fail(IONO_ECUDA, "cudaMemcpy(grid tables): %s", cudaGetErrorString(e)); }
    *out = h;
    return IONO_OK;
}

extern "C" int iono_grid_destroy(iono_grid_t h) {
    if (!h) return IONO_OK;
    cudaFree(h->tables);
    delete h;
    return IONO_OK;
}

extern "C" int iono_grid_is_uniform(iono_grid_t h) { return h ? h->uniform : 0; }

// ---------------------------------------------------------------------------
// per-voxel kernels
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ne_from_m_kernel(const double *__restrict__ m, int64_t n, double scale,
                                                         double *__restrict__ ne) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) ne[i] = exp(m[i]) * scale;
}

__global__ void __launch_bounds__(256) mul_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                                   int64_t n, double *__restrict__ out) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = a[i] * b[i];
}

static int ew_grid(int64_t n) {
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

extern "C" int iono_ne_from_m_f64(const double *m, int64_t nvox, double scale, double *ne_out, void *stream) {
    if (!m || !ne_out || nvox < 0) return fail(IONO_EBADARG, "iono_ne_from_m_f64: bad argument");
    if (nvox == 0) return IONO_OK;
    ne_from_m_kernel<<<ew_grid(nvox), 256, 0, (cudaStream_t)stream>>>(m, nvox, scale, ne_out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

extern "C" int iono_mul_f64(const double *a, const double *b, int64_t n, double *out, void *stream) {
    if (!a || !b || !out || n < 0) return fail(IONO_EBADARG, "iono_mul_f64: bad argument");
    if (n == 0) return IONO_OK;
    mul_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(a, b, n, out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

// ---------------------------------------------------------------------------
// ray generation: one thread per ray derives the ray constants (unit momentum,
// slopes, z step) into shared memory, then the CTA's warps write the (4,Ns) block
// of each ray with coalesced stores.  IEEE ops without FMA contraction in the
// order NumPy evaluates the closed form, so the result is
// bit-identical to the oracle's cast_ray.
// ---------------------------------------------------------------------------
struct RayConst {
    double x0, y0, z0, kx, ky, pz, step;
};

constexpr int CAST_RAYS_PER_CTA = 64;

__global__ void __launch_bounds__(256) cast_rays_kernel(const double *__restrict__ origins,
                                                         const double *__restrict__ directions, int64_t nrays,
                                                         double tmax, int Ns, double *__restrict__ rays) {
    __shared__ RayConst rc[CAST_RAYS_PER_CTA];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int64_t base = (int64_t)blockIdx.x * CAST_RAYS_PER_CTA; base < nrays;
         base += (int64_t)gridDim.x * CAST_RAYS_PER_CTA) {
        __syncthreads();
        if (threadIdx.x < CAST_RAYS_PER_CTA && base + threadIdx.x < nrays) {
            const double *o = origins + (base + threadIdx.x) * 3;
            const double *d = directions + (base + threadIdx.x) * 3;
            double xd = d[0], yd = d[1], zd = d[2];
            double sdot = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(xd, xd), __dmul_rn(yd, yd)), __dmul_rn(zd, zd)));
            double px = __ddiv_rn(xd, sdot), py = __ddiv_rn(yd, sdot), pz = __ddiv_rn(zd, sdot);
            RayConst c;
            c.x0 = o[0]; c.y0 = o[1]; c.z0 = o[2];
            c.kx = __ddiv_rn(px, pz);
            c.ky = __ddiv_rn(py, pz);
            c.pz = pz;
            c.step = __ddiv_rn(__dadd_rn(tmax, -c.z0), (double)(Ns - 1));
            rc[threadIdx.x] = c;
        }
        __syncthreads();
        for (int r = warp; r < CAST_RAYS_PER_CTA && base + r < nrays; r += nwarp) {
            const RayConst c = rc[r];
            double *out = rays + (base + r) * 4 * (int64_t)Ns;
            for (int i = lane; i < Ns; i += 32) {
                double z = (i == Ns - 1 && Ns > 1) ? tmax : __dadd_rn(__dmul_rn((double)i, c.step), c.z0);
                double dz = __dadd_rn(z, -c.z0);
                __stcs(out + i, __dadd_rn(c.x0, __dmul_rn(c.kx, dz)));
                __stcs(out + Ns + i, __dadd_rn(c.y0, __dmul_rn(c.ky, dz)));
                __stcs(out + 2 * (int64_t)Ns + i, z);
                __stcs(out + 3 * (int64_t)Ns + i, __ddiv_rn(dz, c.pz));
            }
        }
    }
}

extern "C" int iono_cast_rays_straight_f64(const double *origins, const double *directions, int64_t nrays,
                                           double tmax, int Ns, double *rays_out, void *stream) {
    if (!origins || !directions || !rays_out || nrays < 0 || Ns < 1)
        return fail(IONO_EBADARG, "iono_cast_rays_straight_f64: bad argument");
    if (nrays == 0) return IONO_OK;
    int64_t ctas = (nrays + CAST_RAYS_PER_CTA - 1) / CAST_RAYS_PER_CTA;
    int64_t cap = (int64_t)sm_count() * 8;
    cast_rays_kernel<<<(int)(ctas < cap ? ctas : cap), 256, 0, (cudaStream_t)stream>>>(origins, directions, nrays,
                                                                                       tmax, Ns, rays_out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

// ---------------------------------------------------------------------------
// point-wise trilinear interpolation (TriCubic.interp / .extrapolate)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void exact_cell(const Axis &a, double x, int &i, double &t) {
    while (i < a.n - 2 && x >= a.tab[i + 1].x) ++i;
    while (i > 0 && x < a.tab[i].x) --i;
    const double gi = a.tab[i].x;
    t = __ddiv_rn(x - gi, a.tab[i + 1].x - gi);
}

template <bool UNIFORM>
__global__ void __launch_bounds__(256) interp_kernel(Grid g, const double *__restrict__ M,
                                                     const double *__restrict__ x, const double *__restrict__ y,
                                                     const double *__restrict__ z, int64_t n,
                                                     double *__restrict__ out, unsigned long long *oob_count) {
    const int ny = g.ax[1].n, nz = g.ax[2].n;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned int n_oob = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        int ix, iy, iz;
        double tx, ty, tz;
        bool oob = false;
        locate<UNIFORM>(g.ax[0].tab, g.ax[0], x[p], ix, tx, oob);
        locate<UNIFORM>(g.ax[1].tab, g.ax[1], y[p], iy, ty, oob);
        locate<UNIFORM>(g.ax[2].tab, g.ax[2], z[p], iz, tz, oob);
        n_oob += oob;
        // exact SciPy cell and division here (this kernel is not the hot path)
        exact_cell(g.ax[0], x[p], ix, tx);
        exact_cell(g.ax[1], y[p], iy, ty);
        exact_cell(g.ax[2], z[p], iz, tz);
        const double *c = M + ((int64_t)ix * ny + iy) * nz + iz;
        const int64_t sy = nz, sx = (int64_t)ny * nz;
        // SciPy's order: sum over corners 000..111 of M * ((wx*wy)*wz)
        const double wx0 = 1.0 - tx, wy0 = 1.0 - ty, wz0 = 1.0 - tz;
        double v = 0.0;
        v = __dadd_rn(v, __dmul_rn(c[0], __dmul_rn(__dmul_rn(wx0, wy0), wz0)));
        v = __dadd_rn(v, __dmul_rn(c[1], __dmul_rn(__dmul_rn(wx0, wy0), tz)));
        v = __dadd_rn(v, __dmul_rn(c[sy], __dmul_rn(__dmul_rn(wx0, ty), wz0)));
        v = __dadd_rn(v, __dmul_rn(c[sy + 1], __dmul_rn(__dmul_rn(wx0, ty), tz)));
        v = __dadd_rn(v, __dmul_rn(c[sx], __dmul_rn(__dmul_rn(tx, wy0), wz0)));
        v = __dadd_rn(v, __dmul_rn(c[sx + 1], __dmul_rn(__dmul_rn(tx, wy0), tz)));
        v = __dadd_rn(v, __dmul_rn(c[sx + sy], __dmul_rn(__dmul_rn(tx, ty), wz0)));
        v = __dadd_rn(v, __dmul_rn(c[sx + sy + 1], __dmul_rn(__dmul_rn(tx, ty), tz)));
        out[p] = v;
    }
    if (n_oob) atomicAdd(oob_count, (unsigned long long)n_oob);
}

extern "C" int iono_tci_interp_f64(iono_grid_t grid, const double *M, const double *x, const double *y,
                                   const double *z, int64_t n, int extrapolate, double *out,
                                   unsigned long long *oob_count, void *stream) {
    (void)extrapolate;  // the arithmetic is identical; the flag only decides whether the caller raises
    if (!grid || !M || !x || !y || !z || !out || !oob_count || n < 0)
        return fail(IONO_EBADARG, "iono_tci_interp_f64: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(oob_count, 0, sizeof(unsigned long long), st));
    if (n == 0) return IONO_OK;
    if (grid->uniform)
        interp_kernel<true><<<ew_grid(n), 256, 0, st>>>(grid->dev, M, x, y, z, n, out, oob_count);
    else
        interp_kernel<false><<<ew_grid(n), 256, 0, st>>>(grid->dev, M, x, y, z, n, out, oob_count);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

// ---------------------------------------------------------------------------
// The ray sweep: forward (gather + reduce) and adjoint (scatter)
// ---------------------------------------------------------------------------
struct SweepParams {
    Grid g;
    const double *field;   // forward: ne (nx,ny,nz).  adjoint: unused
    double *acc;           // adjoint: accumulator (nx,ny,nz)
    const double *rays;    // (R,4,Ns)
    const double *coef;    // adjoint: per-ray coefficient (R)
    double *tec;           // forward: per-ray integral (R)
    unsigned long long *oob_count;
    long long R;
    int Ns;
    int stages;            // ring depth per warp
    RayOrder order;
};

// Samples per ring stage.  A stage holds x[C], y[C], z[C] and s[C+4] (two halo
// samples either side for the Simpson weights).
template <int C>
struct StageLayout {
    static constexpr int S_OFF = 3 * C;              // in doubles
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
        if (lane == 0) {
            mbar_expect_tx(bar, (uint32_t)((3 * n_c + (s_hi - s_lo)) * 8));
            bulk_g2s(stage, ray + c0, n_c * 8, bar, policy);
            bulk_g2s(stage + C, ray + Ns + c0, n_c * 8, bar, policy);
            bulk_g2s(stage + 2 * C, ray + 2 * (int64_t)Ns + c0, n_c * 8, bar, policy);
            bulk_g2s(sdst, ray + 3 * (int64_t)Ns + s_lo, (s_hi - s_lo) * 8, bar, policy);
        }
    } else {
        for (int i = lane; i < n_c; i += 32) {
            stage[i] = ld_stream(ray + c0 + i, policy);
            stage[C + i] = ld_stream(ray + Ns + c0 + i, policy);
            stage[2 * C + i] = ld_stream(ray + 2 * (int64_t)Ns + c0 + i, policy);
        }
        for (int i = lane; i < s_hi - s_lo; i += 32) sdst[i] = ld_stream(ray + 3 * (int64_t)Ns + s_lo + i, policy);
    }
}

// MODE 0: forward, MODE 1: adjoint
template <int MODE, bool UNIFORM, int C, bool BULK>
__global__ void __launch_bounds__(512, 1) ray_sweep_kernel(const SweepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int nx = p.g.ax[0].n, ny = p.g.ax[1].n, nz = p.g.ax[2].n;

    // shared: [axis tables][per-warp mbarriers][per-warp stages]
    double2 *tabx = reinterpret_cast<double2 *>(smem_raw);
    double2 *taby = tabx + nx;
    double2 *tabz = taby + ny;
    size_t off = (((size_t)(nx + ny + nz) * sizeof(double2)) + 127) / 128 * 128;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + off) + (size_t)warp * p.stages;
    off += (((size_t)nwarp * p.stages * sizeof(uint64_t)) + 127) / 128 * 128;
    unsigned char *ring = smem_raw + off + (size_t)warp * p.stages * StageLayout<C>::BYTES;

    for (int i = threadIdx.x; i < nx; i += blockDim.x) tabx[i] = p.g.ax[0].tab[i];
    for (int i = threadIdx.x; i < ny; i += blockDim.x) taby[i] = p.g.ax[1].tab[i];
    for (int i = threadIdx.x; i < nz; i += blockDim.x) tabz[i] = p.g.ax[2].tab[i];
    if (BULK && lane == 0)
        for (int s = 0; s < p.stages; ++s) mbar_init(&bars[s], 1);
    if (BULK) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_grid = policy_evict_last();

    const int Ns = p.Ns;
    const int chunks = (Ns + C - 1) / C;
    // Static schedule: bundle b = nwarp consecutive work indices; CTA c takes b = c, c+G, ...
    const long long n_bundles = (p.R + nwarp - 1) / nwarp;
    long long my_rays = 0;
    for (long long b = blockIdx.x; b < n_bundles; b += gridDim.x)
        if (b * nwarp + warp < p.R) ++my_rays;
    const long long n_items = my_rays * chunks;
    auto item_ray = [&](long long item) -> const double * {
        long long k = item / chunks;
        long long q = ((long long)blockIdx.x + k * gridDim.x) * nwarp + warp;
        return p.rays + ray_of(p.order, q) * 4 * (long long)Ns;
    };
    auto stage_ptr = [&](long long item) -> double * {
        return reinterpret_cast<double *>(ring + (size_t)(item % p.stages) * StageLayout<C>::BYTES);
    };

    if (BULK)
        for (long long it = 0; it < p.stages - 1 && it < n_items; ++it)
            fill_stage<C, true>(stage_ptr(it), &bars[it % p.stages], item_ray(it), Ns, (int)(it % chunks) * C, lane,
                                pol_stream);

    unsigned int n_oob = 0;
    double acc = 0.0;
    double coef = 0.0;
    const long long sy = nz, sx = (long long)ny * nz;

    for (long long item = 0; item < n_items; ++item) {
        const int chunk = (int)(item % chunks);
        const int c0 = chunk * C;
        double *stage = stage_ptr(item);
        if (BULK) {
            const long long nxt = item + p.stages - 1;
            if (nxt < n_items)
                fill_stage<C, true>(stage_ptr(nxt), &bars[nxt % p.stages], item_ray(nxt), Ns, (int)(nxt % chunks) * C,
                                    lane, pol_stream);
            mbar_wait(&bars[item % p.stages], (uint32_t)((item / p.stages) & 1));
        } else {
            fill_stage<C, false>(stage, nullptr, item_ray(item), Ns, c0, lane, pol_stream);
            __syncwarp();
        }
        long long q = 0;
        if (chunk == 0) {
            acc = 0.0;
            if (MODE == 1) {
                q = ((long long)blockIdx.x + (item / chunks) * gridDim.x) * nwarp + warp;
                coef = p.coef[ray_of(p.order, q)];
            }
        }
        const double *sx_ = stage, *sy_ = stage + C, *sz_ = stage + 2 * C;
        const double *ss_ = stage + StageLayout<C>::S_OFF + 2;   // ss_[j] = s[c0 + j]
        const int n_c = min(C, Ns - c0);
#pragma unroll 2
        for (int j = lane; j < n_c; j += 32) {
            const int i = c0 + j;
            int ix, iy, iz;
            double tx, ty, tz;
            bool oob = false;
            locate<UNIFORM>(tabx, p.g.ax[0], sx_[j], ix, tx, oob);
            locate<UNIFORM>(taby, p.g.ax[1], sy_[j], iy, ty, oob);
            locate<UNIFORM>(tabz, p.g.ax[2], sz_[j], iz, tz, oob);
            n_oob += oob;
            const double w = simpson_weight(i, Ns, ss_[j - 2], ss_[j - 1], ss_[j], ss_[j + 1], ss_[j + 2]);
            const long long v = ((long long)ix * ny + iy) * nz + iz;
            if (MODE == 0) {
                const double *c = p.field + v;
                const double v000 = ld_grid(c, pol_grid), v001 = ld_grid(c + 1, pol_grid);
                const double v010 = ld_grid(c + sy, pol_grid), v011 = ld_grid(c + sy + 1, pol_grid);
                const double v100 = ld_grid(c + sx, pol_grid), v101 = ld_grid(c + sx + 1, pol_grid);
                const double v110 = ld_grid(c + sx + sy, pol_grid), v111 = ld_grid(c + sx + sy + 1, pol_grid);
                const double c00 = fma(tz, v001 - v000, v000), c01 = fma(tz, v011 - v010, v010);
                const double c10 = fma(tz, v101 - v100, v100), c11 = fma(tz, v111 - v110, v110);
                const double c0_ = fma(ty, c01 - c00, c00), c1_ = fma(ty, c11 - c10, c10);
                acc = fma(w, fma(tx, c1_ - c0_, c0_), acc);
            } else {
                double *c = p.acc + v;
                const double a = coef * w;
                const double ax1 = a * tx, ax0 = a - ax1;
                const double a01 = ax0 * ty, a00 = ax0 - a01;
                const double a11 = ax1 * ty, a10 = ax1 - a11;
                double hi;
                hi = a00 * tz; atomicAdd(c, a00 - hi); atomicAdd(c + 1, hi);
                hi = a01 * tz; atomicAdd(c + sy, a01 - hi); atomicAdd(c + sy + 1, hi);
                hi = a10 * tz; atomicAdd(c + sx, a10 - hi); atomicAdd(c + sx + 1, hi);
                hi = a11 * tz; atomicAdd(c + sx + sy, a11 - hi); atomicAdd(c + sx + sy + 1, hi);
            }
        }
        if (MODE == 0 && chunk == chunks - 1) {
            const double tot = warp_sum(acc);
            if (lane == 0) {
                q = ((long long)blockIdx.x + (item / chunks) * gridDim.x) * nwarp + warp;
                p.tec[ray_of(p.order, q)] = tot;
            }
        }
        __syncwarp();
    }
    if (n_oob) atomicAdd(p.oob_count, (unsigned long long)n_oob);
}

struct SweepConfig {
    int warps;    // per CTA
    int stages;
    int chunk;    // 64 or 128
    int ctas_per_sm;
};

static SweepConfig sweep_config(int Ns) {
    SweepConfig c;
    c.warps = 16;
    c.stages = 3;
    c.chunk = 64;
    c.ctas_per_sm = 1;
    const char *e;
    if ((e = getenv("IONO_SWEEP_WARPS"))) c.warps = atoi(e);
    if ((e = getenv("IONO_SWEEP_STAGES"))) c.stages = atoi(e);
    if ((e = getenv("IONO_SWEEP_CHUNK"))) c.chunk = atoi(e);
    if (c.warps < 1) c.warps = 1;
    if (c.warps > 16) c.warps = 16;
    if (c.stages < 2) c.stages = 2;
    if (c.stages > 8) c.stages = 8;
    if (c.chunk != 64) c.chunk = 128;
    return c;
}

static RayOrder make_order(int order, int Na, int Nt, int Nd) {
    RayOrder o;
    const long long sa = (long long)Nt * Nd, st = Nd, sd = 1;
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

template <int MODE, bool UNIFORM, int C, bool BULK>
static int launch_sweep_t(const SweepParams &p, const SweepConfig &cfg, size_t smem, int ctas, cudaStream_t st) {
    auto kern = ray_sweep_kernel<MODE, UNIFORM, C, BULK>;
    CU_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ctas, cfg.warps * 32, smem, st>>>(p);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

template <int MODE>
static int launch_sweep(SweepParams p, iono_grid_t grid, cudaStream_t st) {
    SweepConfig cfg = sweep_config(p.Ns);
    p.stages = cfg.stages;
    const size_t stage_bytes = cfg.chunk == 64 ? StageLayout<64>::BYTES : StageLayout<128>::BYTES;
    size_t smem = (((size_t)(grid->nx + grid->ny + grid->nz) * sizeof(double2)) + 127) / 128 * 128;
    smem += (((size_t)cfg.warps * cfg.stages * sizeof(uint64_t)) + 127) / 128 * 128;
    smem += (size_t)cfg.warps * cfg.stages * stage_bytes;
    if (smem > 227 * 1024) return fail(IONO_EBADARG, "ray sweep: shared-memory configuration exceeds 227 KB");
    // TMA bulk copies need 16-byte aligned rows: even Ns and a 16-byte aligned base
    const bool bulk = (p.Ns % 2 == 0) && (((uintptr_t)p.rays & 15) == 0) && !getenv("IONO_SWEEP_NO_BULK");
    const long long n_bundles = (p.R + cfg.warps - 1) / cfg.warps;
    long long ctas = (long long)sm_count() * cfg.ctas_per_sm;
    if (ctas > n_bundles) ctas = n_bundles;
    const bool uni = grid->uniform != 0;
#define IONO_DISPATCH(U, CC, B) return launch_sweep_t<MODE, U, CC, B>(p, cfg, smem, (int)ctas, st)
    if (cfg.chunk == 64) {
        if (uni) { if (bulk) IONO_DISPATCH(true, 64, true); else IONO_DISPATCH(true, 64, false); }
        else     { if (bulk) IONO_DISPATCH(false, 64, true); else IONO_DISPATCH(false, 64, false); }
    } else {
        if (uni) { if (bulk) IONO_DISPATCH(true, 128, true); else IONO_DISPATCH(true, 128, false); }
        else     { if (bulk) IONO_DISPATCH(false, 128, true); else IONO_DISPATCH(false, 128, false); }
    }
#undef IONO_DISPATCH
}

extern "C" int iono_tec_forward_f64(iono_grid_t grid, const double *ne, const double *rays, int Na, int Nt, int Nd,
                                    int Ns, int order, double *tec_out, unsigned long long *oob_count,
                                    void *stream) {
    if (!grid || !ne || !rays || !tec_out || !oob_count || Na < 0 || Nt < 0 || Nd < 0 || Ns < 1)
        return fail(IONO_EBADARG, "iono_tec_forward_f64: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(oob_count, 0, sizeof(unsigned long long), st));
    const long long R = (long long)Na * Nt * Nd;
    if (R == 0) return IONO_OK;
    if (Ns < 2) {  // simps of a single sample is 0
        CU_CHECK(cudaMemsetAsync(tec_out, 0, R * sizeof(double), st));
        return IONO_OK;
    }
    SweepParams p;
    memset(&p, 0, sizeof(p));
    p.g = grid->dev; p.field = ne; p.rays = rays; p.tec = tec_out; p.oob_count = oob_count;
    p.R = R; p.Ns = Ns; p.order = make_order(order, Na, Nt, Nd);
    return launch_sweep<0>(p, grid, st);
}

extern "C" int iono_tec_adjoint_f64(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns,
                                    const double *coef, int order, int zero_first, double *acc,
                                    unsigned long long *oob_count, void *stream) {
    if (!grid || !rays || !coef || !acc || !oob_count || Na < 0 || Nt < 0 || Nd < 0 || Ns < 1)
        return fail(IONO_EBADARG, "iono_tec_adjoint_f64: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(oob_count, 0, sizeof(unsigned long long), st));
    if (zero_first)
        CU_CHECK(cudaMemsetAsync(acc, 0, (size_t)grid->nx * grid->ny * grid->nz * sizeof(double), st));
    const long long R = (long long)Na * Nt * Nd;
    if (R == 0 || Ns < 2) return IONO_OK;
    SweepParams p;
    memset(&p, 0, sizeof(p));
    p.g = grid->dev; p.acc = acc; p.rays = rays; p.coef = coef; p.oob_count = oob_count;
    p.R = R; p.Ns = Ns; p.order = make_order(order, Na, Nt, Nd);
    return launch_sweep<1>(p, grid, st);
}

// ---------------------------------------------------------------------------
// small per-ray kernels
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dtec_kernel(const double *__restrict__ tec, int Na, long long ntd, int i0,
                                                    double *__restrict__ dtec) {
    // one thread per (t,d): read the reference value first so that dtec may alias tec
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < ntd; j += stride) {
        const double ref = tec[(long long)i0 * ntd + j];
        for (int a = 0; a < Na; ++a) dtec[a * ntd + j] = tec[a * ntd + j] - ref;
    }
}

extern "C" int iono_dtec_f64(const double *tec, int Na, int Nt, int Nd, int i0, double *dtec_out, void *stream) {
    if (!tec || !dtec_out || Na < 0 || Nt < 0 || Nd < 0 || (Na > 0 && (i0 < 0 || i0 >= Na)))
        return fail(IONO_EBADARG, "iono_dtec_f64: bad argument");
    const long long ntd = (long long)Nt * Nd;
    if (ntd == 0 || Na == 0) return IONO_OK;
    dtec_kernel<<<ew_grid(ntd), 256, 0, (cudaStream_t)stream>>>(tec, Na, ntd, i0, dtec_out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

__global__ void __launch_bounds__(256) adjoint_coef_kernel(const double *__restrict__ g,
                                                            const double *__restrict__ dobs,
                                                            const double *__restrict__ CdCt, int Na, long long ntd,
                                                            int i0, double *__restrict__ coef) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < ntd; j += stride) {
        double sum = 0.0;
        for (int a = 0; a < Na; ++a) {
            const long long k = a * ntd + j;
            const double dd = (g[k] - dobs[k]) / (CdCt[k] + 1e-15);
            sum += dd;
            coef[k] = dd;
        }
        coef[(long long)i0 * ntd + j] -= sum;
    }
}

extern "C" int iono_adjoint_coef_f64(const double *g, const double *dobs, const double *CdCt, int Na, int Nt, int Nd,
                                     int i0, double *coef_out, void *stream) {
    if (!g || !dobs || !CdCt || !coef_out || Na < 0 || Nt < 0 || Nd < 0 || (Na > 0 && (i0 < 0 || i0 >= Na)))
        return fail(IONO_EBADARG, "iono_adjoint_coef_f64: bad argument");
    const long long ntd = (long long)Nt * Nd;
    if (ntd == 0 || Na == 0) return IONO_OK;
    adjoint_coef_kernel<<<ew_grid(ntd), 256, 0, (cudaStream_t)stream>>>(g, dobs, CdCt, Na, ntd, i0, coef_out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

// ---------------------------------------------------------------------------
// misfit: deterministic two-stage sum
// ---------------------------------------------------------------------------
constexpr int MISFIT_BLOCKS = 1024;

__device__ __forceinline__ double block_sum_256(double v) {
    __shared__ double part[8];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < 8 ? part[threadIdx.x] : 0.0;
        t = warp_sum(t);
    }
    return t;
}

__global__ void __launch_bounds__(256) misfit_stage1(const double *__restrict__ g, const double *__restrict__ dobs,
                                                      const double *__restrict__ CdCt, int64_t n,
                                                      double *__restrict__ scratch) {
    double s = 0.0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double r = g[i] - dobs[i];
        s += r * r / (CdCt[i] + 1e-15);
    }
    s = block_sum_256(s);
    if (threadIdx.x == 0) scratch[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) misfit_stage2(const double *__restrict__ scratch, int nblocks,
                                                      double *__restrict__ out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) s += scratch[i];
    s = block_sum_256(s);
    if (threadIdx.x == 0) out[0] = 0.5 * s;
}

extern "C" int64_t iono_misfit_scratch_elems(void) { return MISFIT_BLOCKS; }

extern "C" int iono_misfit_f64(const double *g, const double *dobs, const double *CdCt, int64_t n, double *scratch,
                               double *out, void *stream) {
    if (!g || !dobs || !CdCt || !scratch || !out || n < 0) return fail(IONO_EBADARG, "iono_misfit_f64: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = (int)((n + 255) / 256);
    if (blocks > MISFIT_BLOCKS) blocks = MISFIT_BLOCKS;
    if (blocks < 1) blocks = 1;
    misfit_stage1<<<blocks, 256, 0, st>>>(g, dobs, CdCt, n, scratch);
    CU_CHECK(cudaGetLastError());
    misfit_stage2<<<1, 256, 0, st>>>(scratch, blocks, out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}
