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

// Handles cache device pointers: refuse to launch on another device than the one they were built on.
static int device_check(int handle_device, const char *what) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev != handle_device)
        return fail(IONO_EBADARG, "%s: the handle was created on another CUDA device than the current one", what);
    return IONO_OK;
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
    int uniform;       // all axes: direct cell index (tables in shared memory)
    int exact;         // all axes: exact linspace, in-cell coordinate by arithmetic (no tables in the sweep)
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

// exact <=> every node equals g0 + i*d to within a few ulps of the axis extent (what np.linspace
// produces): the in-cell coordinate (x - g[i])/(g[i+1] - g[i]) may then be computed as
// frac((x - g0)/d) with an absolute error of ~1e-13, without reading the node table.
static bool axis_exact(const double *g, int n) {
    const double d = (g[n - 1] - g[0]) / (n - 1);
    const double tol = 8.0 * 2.220446049250313e-16 * fmax(fabs(g[0]), fabs(g[n - 1]));
    for (int i = 0; i < n; ++i)
        if (fabs(g[i] - (g[0] + i * d)) > tol) return false;
    return !getenv("IONO_NO_EXACT_AXES");
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
    {
        cudaError_t e0 = cudaGetDevice(&h->device);
        if (e0 != cudaSuccess) { delete h; return fail(IONO_ECUDA, "cudaGetDevice: %s", cudaGetErrorString(e0)); }
    }
    size_t total = (size_t)nx + ny + nz;
    std::vector<double2> host(total);
    size_t off = 0;
    cudaError_t e = cudaMalloc(&h->tables, total * sizeof(double2));
    if (e != cudaSuccess) { delete h; return fail(IONO_ECUDA, "cudaMalloc(grid tables): %s", cudaGetErrorString(e)); }
    h->uniform = 1;
    h->exact = 1;
    for (int a = 0; a < 3; ++a) {
        Axis &A = h->dev.ax[a];
        for (int i = 0; i < n[a]; ++i) {
            host[off + i].x = g[a][i];
            host[off + i].y = (i < n[a] - 1) ? 1.0 / (g[a][i + 1] - g[a][i]) : 0.0;
        }
        A.tab = h->tables + off;
        A.n = n[a];
        A.uniform = axis_uniform(g[a], n[a]) ? 1 : 0;
        A.exact = (A.uniform && axis_exact(g[a], n[a])) ? 1 : 0;
        A.inv_d = (n[a] - 1) / (g[a][n[a] - 1] - g[a][0]);
        A.c_guess = -g[a][0] * A.inv_d - 0.5;
        A.g0 = g[a][0];
        A.glast = g[a][n[a] - 1];
        h->uniform &= A.uniform;
        h->exact &= A.exact;
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

// Quad layout (iono_device.cuh): q[v] = { f[v], f[v + dz], f[v + dy], f[v + dy + dz] } with the +1 steps in
// z and y clamped at the last node.  `m` != NULL: f = exp(m) * scale, computed on the fly (and written to
// ne_out if that is given), so that one launch turns the model into both layouts.
__global__ void __launch_bounds__(256) quads_kernel(const double *__restrict__ f, const double *__restrict__ m,
                                                     double scale, int ny, int nz, int64_t n,
                                                     double *__restrict__ ne_out, double4 *__restrict__ q) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const int iz = (int)(v % nz);
        const int iy = (int)((v / nz) % ny);
        const int64_t dz = (iz + 1 < nz) ? 1 : 0, dy = (iy + 1 < ny) ? nz : 0;
        double a, b, c, d;
        if (m) {
            a = exp(m[v]) * scale; b = exp(m[v + dz]) * scale;
            c = exp(m[v + dy]) * scale; d = exp(m[v + dy + dz]) * scale;
            if (ne_out) ne_out[v] = a;
        } else {
            a = f[v]; b = f[v + dz]; c = f[v + dy]; d = f[v + dy + dz];
        }
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(q + v), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
    }
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

extern "C" int iono_quads_from_ne_f64(const double *ne, int nx, int ny, int nz, double *quads_out, void *stream) {
    const int64_t n = (int64_t)nx * ny * nz;
    if (!ne || !quads_out || nx < 1 || ny < 1 || nz < 1 || ((uintptr_t)quads_out & 31))
        return fail(IONO_EBADARG, "iono_quads_from_ne_f64: bad argument (quads_out must be 32-byte aligned)");
    quads_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(ne, nullptr, 1.0, ny, nz, n, nullptr,
                                                              reinterpret_cast<double4 *>(quads_out));
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

extern "C" int iono_ne_quads_from_m_f64(const double *m, int nx, int ny, int nz, double scale, double *ne_out,
                                        double *quads_out, void *stream) {
    const int64_t n = (int64_t)nx * ny * nz;
    if (!m || !quads_out || nx < 1 || ny < 1 || nz < 1 || ((uintptr_t)quads_out & 31))
        return fail(IONO_EBADARG, "iono_ne_quads_from_m_f64: bad argument (quads_out must be 32-byte aligned)");
    quads_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(nullptr, m, scale, ny, nz, n, ne_out,
                                                              reinterpret_cast<double4 *>(quads_out));
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

// ARC: arc length as the independent variable (Fermat(type='s'), fermat.py:74-82,163-166):
// s = linspace(0, tmax, N), (x,y,z) = origin + p s.
template <bool ARC>
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
            c.kx = ARC ? px : __ddiv_rn(px, pz);
            c.ky = ARC ? py : __ddiv_rn(py, pz);
            c.pz = pz;
            c.step = ARC ? __ddiv_rn(tmax, (double)(Ns - 1)) : __ddiv_rn(__dadd_rn(tmax, -c.z0), (double)(Ns - 1));
            rc[threadIdx.x] = c;
        }
        __syncthreads();
        for (int r = warp; r < CAST_RAYS_PER_CTA && base + r < nrays; r += nwarp) {
            const RayConst c = rc[r];
            double *out = rays + (base + r) * 4 * (int64_t)Ns;
            for (int i = lane; i < Ns; i += 32) {
                if (ARC) {
                    const double sv = (i == Ns - 1 && Ns > 1) ? tmax : __dmul_rn((double)i, c.step);
                    __stcs(out + i, __dadd_rn(c.x0, __dmul_rn(c.kx, sv)));
                    __stcs(out + Ns + i, __dadd_rn(c.y0, __dmul_rn(c.ky, sv)));
                    __stcs(out + 2 * (int64_t)Ns + i, __dadd_rn(c.z0, __dmul_rn(c.pz, sv)));
                    __stcs(out + 3 * (int64_t)Ns + i, sv);
                    continue;
                }
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

// Same as cast_rays_kernel, with the model-frame origin and direction of every ray derived on the fly
// from ITRS inputs by the per-time rotation of the reference's Pointing frame
// (astro/frames/pointing_frame.py:151-183): origin = R_t (p_a - p0) / 1000 [km], direction = R_t d_{t,k}.
__global__ void __launch_bounds__(256) cast_rays_frames_kernel(const double *__restrict__ ants_itrs_m,
                                                                const double *__restrict__ p0_itrs_m,
                                                                const double *__restrict__ R,
                                                                const double *__restrict__ dirs_itrs, int Na, int Nt,
                                                                int Nd, double tmax, int Ns,
                                                                double *__restrict__ rays) {
    __shared__ RayConst rc[CAST_RAYS_PER_CTA];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int64_t nrays = (int64_t)Na * Nt * Nd;
    for (int64_t base = (int64_t)blockIdx.x * CAST_RAYS_PER_CTA; base < nrays;
         base += (int64_t)gridDim.x * CAST_RAYS_PER_CTA) {
        __syncthreads();
        if (threadIdx.x < CAST_RAYS_PER_CTA && base + threadIdx.x < nrays) {
            const int64_t ray = base + threadIdx.x;
            const int a = (int)(ray / ((int64_t)Nt * Nd));
            const int rem = (int)(ray - (int64_t)a * Nt * Nd);
            const int t = rem / Nd, k = rem - t * Nd;
            const double *Rt = R + (int64_t)t * 9;
            const double dx = __dadd_rn(ants_itrs_m[a * 3 + 0], -p0_itrs_m[0]);
            const double dy = __dadd_rn(ants_itrs_m[a * 3 + 1], -p0_itrs_m[1]);
            const double dz = __dadd_rn(ants_itrs_m[a * 3 + 2], -p0_itrs_m[2]);
            const double *d = dirs_itrs + ((int64_t)t * Nd + k) * 3;
            double o[3], v[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {   // row r of R_t times the vector, left to right, no FMA contraction
                o[r] = __dadd_rn(__dadd_rn(__dmul_rn(Rt[3 * r], dx), __dmul_rn(Rt[3 * r + 1], dy)),
                                 __dmul_rn(Rt[3 * r + 2], dz));
                v[r] = __dadd_rn(__dadd_rn(__dmul_rn(Rt[3 * r], d[0]), __dmul_rn(Rt[3 * r + 1], d[1])),
                                 __dmul_rn(Rt[3 * r + 2], d[2]));
                o[r] = __ddiv_rn(o[r], 1000.0);   // m -> km (calc_rays.py:132 `.to(au.km)`)
            }
            const double sdot = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])),
                                                     __dmul_rn(v[2], v[2])));
            const double px = __ddiv_rn(v[0], sdot), py = __ddiv_rn(v[1], sdot), pz = __ddiv_rn(v[2], sdot);
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

extern "C" int iono_cast_rays_frames_f64(const double *ants_itrs_m, const double *p0_itrs_m, const double *R,
                                         const double *dirs_itrs, int Na, int Nt, int Nd, double tmax_km, int Ns,
                                         double *rays_out, void *stream) {
    const int64_t nrays = (int64_t)Na * Nt * Nd;
    if (Na < 0 || Nt < 0 || Nd < 0 || Ns < 1) return fail(IONO_EBADARG, "iono_cast_rays_frames_f64: bad argument");
    if (nrays == 0) return IONO_OK;
    if (!ants_itrs_m || !p0_itrs_m || !R || !dirs_itrs || !rays_out)
        return fail(IONO_EBADARG, "iono_cast_rays_frames_f64: NULL pointer");
    int64_t ctas = (nrays + CAST_RAYS_PER_CTA - 1) / CAST_RAYS_PER_CTA;
    int64_t cap = (int64_t)sm_count() * 8;
    cast_rays_frames_kernel<<<(int)(ctas < cap ? ctas : cap), 256, 0, (cudaStream_t)stream>>>(
        ants_itrs_m, p0_itrs_m, R, dirs_itrs, Na, Nt, Nd, tmax_km, Ns, rays_out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

extern "C" int iono_cast_rays_straight_f64(const double *origins, const double *directions, int64_t nrays,
                                           double tmax, int Ns, double *rays_out, void *stream) {
    if (!origins || !directions || !rays_out || nrays < 0 || Ns < 1)
        return fail(IONO_EBADARG, "iono_cast_rays_straight_f64: bad argument");
    if (nrays == 0) return IONO_OK;
    int64_t ctas = (nrays + CAST_RAYS_PER_CTA - 1) / CAST_RAYS_PER_CTA;
    int64_t cap = (int64_t)sm_count() * 8;
    cast_rays_kernel<false><<<(int)(ctas < cap ? ctas : cap), 256, 0, (cudaStream_t)stream>>>(origins, directions,
                                                                                              nrays, tmax, Ns, rays_out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

// Fermat(type='s') with straight rays: s = linspace(0, smax, Ns), position = origin + unit direction * s.
extern "C" int iono_cast_rays_arclength_f64(const double *origins, const double *directions, int64_t nrays,
                                            double smax, int Ns, double *rays_out, void *stream) {
    if (!origins || !directions || !rays_out || nrays < 0 || Ns < 1)
        return fail(IONO_EBADARG, "iono_cast_rays_arclength_f64: bad argument");
    if (nrays == 0) return IONO_OK;
    int64_t ctas = (nrays + CAST_RAYS_PER_CTA - 1) / CAST_RAYS_PER_CTA;
    int64_t cap = (int64_t)sm_count() * 8;
    cast_rays_kernel<true><<<(int)(ctas < cap ? ctas : cap), 256, 0, (cudaStream_t)stream>>>(origins, directions,
                                                                                             nrays, smax, Ns, rays_out);
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
    if (!grid || !oob_count || n < 0 || (n > 0 && (!M || !x || !y || !z || !out)))
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

#include "iono_sweep.cuh"
#include "iono_adjoint_runs.cuh"
#include "iono_prepared.cuh"
#include "iono_prepared_adjoint.cuh"
#include "iono_backproject.cuh"
#include "iono_chord.cuh"
#include "iono_gaussian.cuh"
#include "iono_optical.cuh"
#include "iono_peer.cuh"
#include "iono_optim.cuh"
#include "iono_tricubic.cuh"

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
    const long long ntd = (long long)Nt * Nd;
    if (Na < 0 || Nt < 0 || Nd < 0) return fail(IONO_EBADARG, "iono_dtec_f64: bad argument");
    if (ntd == 0 || Na == 0) return IONO_OK;
    if (!tec || !dtec_out || i0 < 0 || i0 >= Na) return fail(IONO_EBADARG, "iono_dtec_f64: bad argument");
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
    const long long ntd = (long long)Nt * Nd;
    if (Na < 0 || Nt < 0 || Nd < 0) return fail(IONO_EBADARG, "iono_adjoint_coef_f64: bad argument");
    if (ntd == 0 || Na == 0) return IONO_OK;
    if (!g || !dobs || !CdCt || !coef_out || i0 < 0 || i0 >= Na)
        return fail(IONO_EBADARG, "iono_adjoint_coef_f64: bad argument");
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

// ---------------------------------------------------------------------------
// Fused per-ray step between the forward and the adjoint: dTEC (forward_equation.py:50), weighted residual
// and adjoint coefficients (gradient.py:33-37 + the reference-antenna term of the exact transpose), misfit
// (line_search.py:48-49) -- one launch instead of dtec + misfit x 2 + adjoint_coef + permute_coef, with the
// arithmetic of those kernels (bit-identical g, coef; the misfit is summed in a different, still fixed,
// order).  A CTA owns 8 times x 32 directions and walks the antennas, so the coefficients can also be written
// in the back-projector's internal (antenna, direction, time) order through a shared-memory transpose.
// ---------------------------------------------------------------------------
constexpr int RES_TT = 8, RES_TD = 32, RES_AG = 8;   // tile of 8 times x 32 directions, 8 antennas per work item
constexpr int RES_MAX_CTAS = 2048;

// Work item = (tile, antenna group): items are independent except for the sum over antennas in the reference
// antenna's coefficient, which the LAST item of a tile to finish forms from the per-group partial sums in group
// order (fixed order => reproducible).  With one item per CTA the kernel is one memory round trip deep instead
// of Na/8 (it was latency-bound: 44 us at the LOFAR case whatever the number of rays).
struct ResidualScratch {
    double *partial;          // [RES_MAX_CTAS] misfit partials
    double *psum;             // [groups][tiles][256] sum of dd over the group's antennas
    double *ddref;            // [tiles][256] dd of the reference antenna
    unsigned int *tile_count; // [tiles] arrivals per tile
    unsigned int *counter;    // arrivals of CTAs
};

__global__ void __launch_bounds__(256) residual_kernel(const double *__restrict__ tec, const double *__restrict__ dobs,
                                                        const double *__restrict__ CdCt, int Na, int Nt, int Nd, int i0,
                                                        double *__restrict__ dtec, double *__restrict__ coef,
                                                        double *__restrict__ coef_perm, ResidualScratch sc,
                                                        double *__restrict__ out_S) {
    __shared__ double tile[RES_AG][RES_TT][RES_TD + 1];
    __shared__ bool last_of_tile, last;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int dl = threadIdx.x >> 3, tl = threadIdx.x & 7;   // transposed role: direction, time within the tile
    const int tiles_d = (Nd + RES_TD - 1) / RES_TD, tiles_t = (Nt + RES_TT - 1) / RES_TT;
    const int tiles = tiles_d * tiles_t, groups = (Na + RES_AG - 1) / RES_AG;
    const long long ntd = (long long)Nt * Nd;
    double S = 0.0;
    for (int item = blockIdx.x; item < tiles * groups; item += gridDim.x) {
        const int b = item / groups, grp = item - b * groups, a0 = grp * RES_AG;
        const int t0 = (b / tiles_d) * RES_TT, d0 = (b % tiles_d) * RES_TD;
        const int t = t0 + ty, d = d0 + tx;
        const bool ok = t < Nt && d < Nd;
        const bool okT = (d0 + dl < Nd) && (t0 + tl < Nt);
        const long long j = (long long)t * Nd + d;
        const double ref = ok ? tec[(long long)i0 * ntd + j] : 0.0;
        double tv[RES_AG], ov[RES_AG], cv[RES_AG];
#pragma unroll
        for (int u = 0; u < RES_AG; ++u) {   // 8 x 3 independent loads in flight per thread
            const bool in = ok && (a0 + u < Na);
            const long long k = (long long)(a0 + u) * ntd + j;
            tv[u] = in ? tec[k] : 0.0;
            ov[u] = in ? dobs[k] : 0.0;
            cv[u] = in ? CdCt[k] : 1.0;
        }
        double sum = 0.0, dd_ref = 0.0;
        __syncthreads();     // the previous item's reads of `tile` are done
#pragma unroll
        for (int u = 0; u < RES_AG; ++u) {
            const int a = a0 + u;
            double dd = 0.0;
            if (ok && a < Na) {
                const long long k = (long long)a * ntd + j;
                const double g = tv[u] - ref;
                const double r = g - ov[u];
                dd = r / (cv[u] + 1e-15);
                S = fma(r, dd, S);
                sum += dd;
                dtec[k] = g;
                if (a == i0) dd_ref = dd;
                else if (coef) coef[k] = dd;
            }
            tile[u][ty][tx] = dd;
        }
        sc.psum[((long long)grp * tiles + b) * 256 + threadIdx.x] = sum;
        if (i0 >= a0 && i0 < a0 + RES_AG) sc.ddref[(long long)b * 256 + threadIdx.x] = dd_ref;
        __syncthreads();
        if (coef_perm) {
#pragma unroll
            for (int u = 0; u < RES_AG; ++u) {
                const int a = a0 + u;
                if (okT && a < Na && a != i0)
                    coef_perm[((long long)a * Nd + d0 + dl) * Nt + t0 + tl] = tile[u][tl][dl];
            }
        }
        // reference antenna: c = dd - sum over ALL antennas (the -tec[i0] term of dTEC, transposed), formed by the
        // last group of this tile to arrive
        if (threadIdx.x == 0) {
            __threadfence();
            last_of_tile = (atomicAdd(sc.tile_count + b, 1u) == (unsigned int)groups - 1u);
        }
        __syncthreads();
        if (last_of_tile) {
            __threadfence();
            double total = 0.0;
            for (int q = 0; q < groups; ++q) total += __ldcg(sc.psum + ((long long)q * tiles + b) * 256 + threadIdx.x);
            const double c_ref = __ldcg(sc.ddref + (long long)b * 256 + threadIdx.x) - total;
            if (ok && coef) coef[(long long)i0 * ntd + j] = c_ref;
            if (coef_perm) {
                tile[0][ty][tx] = c_ref;
                __syncthreads();
                if (okT) coef_perm[((long long)i0 * Nd + d0 + dl) * Nt + t0 + tl] = tile[0][tl][dl];
            }
            if (threadIdx.x == 0) sc.tile_count[b] = 0u;
        }
    }
    // misfit: per-CTA partial, the last CTA to arrive adds the partials in CTA order
    __syncthreads();
    S = block_sum_256(S);
    if (threadIdx.x == 0) {
        sc.partial[blockIdx.x] = S;
        __threadfence();
        last = (atomicAdd(sc.counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(sc.partial + i);
        s = block_sum_256(s);
        if (threadIdx.x == 0) { out_S[0] = 0.5 * s; *sc.counter = 0u; }
    }
}

static void residual_layout(int Na, int Nt, int Nd, long long &tiles, long long &groups, long long &elems) {
    tiles = (long long)((Nd + RES_TD - 1) / RES_TD) * ((Nt + RES_TT - 1) / RES_TT);
    groups = (Na + RES_AG - 1) / RES_AG;
    elems = RES_MAX_CTAS + (groups + 1) * tiles * 256 + (tiles + 2 + 1) / 2 + 1;
}

extern "C" int64_t iono_residual_scratch_elems(int Na, int Nt, int Nd) {
    long long tiles, groups, elems;
    residual_layout(Na < 1 ? 1 : Na, Nt < 1 ? 1 : Nt, Nd < 1 ? 1 : Nd, tiles, groups, elems);
    return elems;
}

extern "C" int iono_residual_f64(const double *tec, const double *dobs, const double *CdCt, int Na, int Nt, int Nd,
                                 int i0, double *dtec_out, double *coef_out, double *coef_perm_out, double *scratch,
                                 double *misfit_out, void *stream) {
    if (Na < 0 || Nt < 0 || Nd < 0 || !scratch || !misfit_out)
        return fail(IONO_EBADARG, "iono_residual_f64: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if ((long long)Na * Nt * Nd == 0) {
        CU_CHECK(cudaMemsetAsync(misfit_out, 0, sizeof(double), st));
        return IONO_OK;
    }
    if (!tec || !dobs || !CdCt || !dtec_out || i0 < 0 || i0 >= Na)
        return fail(IONO_EBADARG, "iono_residual_f64: bad argument");
    long long tiles, groups, elems;
    residual_layout(Na, Nt, Nd, tiles, groups, elems);
    if (tiles * groups > 0x7fffffffLL) return fail(IONO_EBADARG, "iono_residual_f64: too many rays");
    ResidualScratch sc;
    sc.partial = scratch;
    sc.psum = scratch + RES_MAX_CTAS;
    sc.ddref = sc.psum + groups * tiles * 256;
    sc.tile_count = reinterpret_cast<unsigned int *>(sc.ddref + tiles * 256);
    sc.counter = sc.tile_count + tiles;
    CU_CHECK(cudaMemsetAsync(sc.tile_count, 0, (size_t)(tiles + 1) * sizeof(unsigned int), st));
    const long long items = tiles * groups;
    const int blocks = (int)(items < RES_MAX_CTAS ? items : RES_MAX_CTAS);
    residual_kernel<<<blocks, 256, 0, st>>>(tec, dobs, CdCt, Na, Nt, Nd, i0, dtec_out, coef_out, coef_perm_out, sc,
                                            misfit_out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

// ---------------------------------------------------------------------------
// host -> device strided staging
// ---------------------------------------------------------------------------
extern "C" int iono_copy2d_h2d(void *dst, int64_t dpitch, const void *src, int64_t spitch, int64_t width,
                               int64_t height, void *stream) {
    if (width < 0 || height < 0 || dpitch < width || spitch < width)
        return fail(IONO_EBADARG, "iono_copy2d_h2d: bad argument");
    if (width == 0 || height == 0) return IONO_OK;
    if (!dst || !src) return fail(IONO_EBADARG, "iono_copy2d_h2d: NULL pointer");
    CU_CHECK(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)height,
                               cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return IONO_OK;
}

// ---------------------------------------------------------------------------
// model-covariance smoothing: Covariance.smooth == scipy.ndimage.convolve(phi, stencil, mode='nearest')
// (ionosphere/covariance.py:383-385) -- the Cm . (G^T r) step that follows the adjoint in the
// reference's solvers.  out[i,j,k] = sum_{a,b,c} w[a,b,c] * phi[clamp(i + h - a), clamp(j + h - b),
// clamp(k + h - c)], h = m/2 (convolution flips the stencil).  One thread per voxel, z fastest; the
// stencil sits in shared memory, the input is read through L1 (each value is reused m^3 times).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) convolve3d_nearest_kernel(const double *__restrict__ phi, int nx, int ny,
                                                                  int nz, const double *__restrict__ w, int m,
                                                                  double *__restrict__ out) {
    extern __shared__ double w_s[];
    for (int i = threadIdx.x; i < m * m * m; i += blockDim.x) w_s[i] = w[i];
    __syncthreads();
    const int h = m / 2;
    const long long n = (long long)nx * ny * nz;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const int k = (int)(v % nz);
        const int j = (int)((v / nz) % ny);
        const int i = (int)(v / ((long long)nz * ny));
        double acc = 0.0;
        for (int a = 0; a < m; ++a) {
            const int ii = min(max(i + h - a, 0), nx - 1);
            for (int b = 0; b < m; ++b) {
                const int jj = min(max(j + h - b, 0), ny - 1);
                const double *row = phi + ((long long)ii * ny + jj) * nz;
                const double *wr = w_s + (a * m + b) * m;
                for (int c = 0; c < m; ++c) {
                    const int kk = min(max(k + h - c, 0), nz - 1);
                    acc = fma(wr[c], __ldg(row + kk), acc);
                }
            }
        }
        out[v] = acc;
    }
}

extern "C" int iono_convolve3d_nearest_f64(const double *phi, int nx, int ny, int nz, const double *stencil, int m,
                                           double *out, void *stream) {
    // 29^3 doubles = 195 KB of shared memory; 31^3 would exceed the 227 KB a CTA can opt into
    if (nx < 0 || ny < 0 || nz < 0 || m < 1 || (m & 1) == 0 || m > 29)
        return fail(IONO_EBADARG, "iono_convolve3d_nearest_f64: bad argument (odd stencil size 1..29)");
    const long long n = (long long)nx * ny * nz;
    if (n == 0) return IONO_OK;
    if (!phi || !stencil || !out || phi == out) return fail(IONO_EBADARG, "iono_convolve3d_nearest_f64: bad pointer");
    const size_t smem = (size_t)m * m * m * sizeof(double);
    CU_CHECK(cudaFuncSetAttribute(convolve3d_nearest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    convolve3d_nearest_kernel<<<ew_grid(n), 256, smem, (cudaStream_t)stream>>>(phi, nx, ny, nz, stencil, m, out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}
