// Prepared forward sweep ("forward projector"): the TEC forward for a ray geometry that is reused
// across iterations (every driver of the reference calls forward_equation with the same rays in each
// iteration: bfgs_dask.py:207-340, iterative_newton.py:954-1017, tests/test_inversion.py:30-39).
// Included by iono_kernels.cu.  The forward-side twin of the voxel-binned back-projector.
//
//   assembly (once per geometry): per sample the cell index v, the in-cell fractions (tx,ty,tz) and
//     the Simpson 'avg' weight w -- computed by the SAME device functions the stateless sweep
//     (iono_sweep.cuh, MODE 0) evaluates per launch -- are written as a record stream, rays stored in
//     the sweep's time-fastest traversal order and every ray cut into 64-sample BLOCKS that hold
//     tx[n] ty[n] tz[n] (w[n]) cell[n] back to back, so that one chunk of one ray is ONE contiguous
//     bulk copy (2304 bytes) -- the producer's instruction count per chunk was the kernel's largest
//     cost with one copy per row (profiles/r02_ncu_step_summary.txt: 158 warp instructions per 32
//     samples, 35 of them the interpolation);
//   factored weights: when every ray's weights are one common pattern times a per-ray factor,
//     w[q][i] = c[q] P[i] to 2e-13 relative (true for every ray set the casting kernels make: s is a
//     linspace, so w = h/3 * {1,4,2,...} up to the rounding of the differences), the w row is dropped
//     -- 28 bytes per sample instead of 36 -- and the apply multiplies by P[i] from shared memory and
//     by c[q] once per ray.  Checked per geometry at create time; IONO_PREP_FACTOR=0 disables it.
//   apply: tec[ray] = sum_s w_s * trilerp(ne; v_s, t_s).  One warp per ray, lanes = consecutive
//     samples, records streamed by TMA 1-D bulk copies into a warp-private ring (as the sweep), but no
//     cell search, no table reads and no weight arithmetic in the loop: 8 corner gathers, 14 fp64
//     operations and one fma per sample.
//
// v, t and w are bit-identical to what MODE 0 derives and the per-lane accumulation order is the
// same, so with per-sample weights the result is bit-identical to iono_tec_forward_f64 (tested); with
// factored weights it agrees to ~1e-14 relative.  The roofline fraction is quoted on the algorithmic
// bytes of the reference's API boundary (32 B per sample).
#pragma once
#include <cub/cub.cuh>

constexpr int PREP_C = 64;      // samples per block and per ring stage

struct iono_forwardprojector {
    unsigned char *rec;   // (R, Nsp) samples in blocks of PREP_C: tx[n] ty[n] tz[n] (w[n]) doubles, cell[n] ints
    double *wscale;       // [R] per-ray weight factor c[q], slot order (factored weights only)
    double *pattern;      // [Ns] common weight pattern P[i] (factored weights only)
    int factored;
    int *records;     // quad records (iono_device.cuh) the samples read: cells v and v + ny*nz, ascending
    long long n_records;
    int *voxels;      // grid nodes that are a corner of a visited cell, ascending (the support of the adjoint)
    long long n_voxels;
    long long R;
    int Na, Nt, Nd, Ns, Nsp;   // Nsp = Ns rounded up to a multiple of 4 (16-byte pieces for the bulk copies)
    int nx, ny, nz;
    int device;
};
typedef struct iono_forwardprojector *iono_forwardprojector_t;

// slot q (time fastest, then antenna, then direction -- IONO_ORDER_TIME) -> ray index in (Na,Nt,Nd)
__device__ __forceinline__ long long prepared_ray_of(int q, int Na, int Nt, int Nd) {
    const int t = q % Nt;
    const int r = q / Nt;
    const int a = r % Na;
    const int d = r / Na;
    return ((long long)a * Nt + t) * Nd + d;
}

__device__ __forceinline__ double prepared_weight(const double *sp, int i, int Ns, bool n_odd) {
    // neighbours outside [0, Ns) are never used by simpson_weight; clamp the reads
    return simpson_weight(i, Ns, n_odd, sp[max(i - 2, 0)], sp[max(i - 1, 0)], sp[i], sp[min(i + 1, Ns - 1)],
                          sp[min(i + 2, Ns - 1)]);
}

// P[i] = w[slot 0][i] * Ns / sum_i w[slot 0][i]   (one warp)
__global__ void __launch_bounds__(32) weight_pattern_kernel(const double *__restrict__ rays, int Na, int Nt, int Nd,
                                                             int Ns, double *__restrict__ pattern) {
    const int lane = threadIdx.x;
    const double *sp = rays + prepared_ray_of(0, Na, Nt, Nd) * 4 * Ns + 3 * Ns;
    const bool n_odd = Ns & 1;
    double tot = 0.0;
    for (int i = lane; i < Ns; i += 32) tot += prepared_weight(sp, i, Ns, n_odd);
    tot = warp_sum(tot);
    tot = __shfl_sync(0xffffffffu, tot, 0);
    for (int i = lane; i < Ns; i += 32) pattern[i] = (tot != 0.0) ? prepared_weight(sp, i, Ns, n_odd) * (double)Ns / tot : 0.0;
}

// c[q] = sum_i w[q][i] / Ns, and the check |w[q][i] - c[q] P[i]| <= tol |c[q] P[i]| for every sample (warp per ray)
__global__ void __launch_bounds__(256) weight_factor_kernel(const double *__restrict__ rays, int R, int Na, int Nt,
                                                             int Nd, int Ns, const double *__restrict__ pattern,
                                                             double tol, double *__restrict__ wscale,
                                                             int *__restrict__ mismatch) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const bool n_odd = Ns & 1;
    bool bad = false;
    for (int q = warp_global; q < R; q += n_warps) {
        const double *sp = rays + prepared_ray_of(q, Na, Nt, Nd) * 4 * Ns + 3 * Ns;
        double tot = 0.0;
        for (int i = lane; i < Ns; i += 32) tot += prepared_weight(sp, i, Ns, n_odd);
        tot = warp_sum(tot);
        const double c = __shfl_sync(0xffffffffu, tot, 0) / (double)Ns;
        for (int i = lane; i < Ns; i += 32) {
            const double want = c * pattern[i];
            if (!(fabs(prepared_weight(sp, i, Ns, n_odd) - want) <= tol * fabs(want))) bad = true;   // (NaN -> bad)
        }
        if (lane == 0) wscale[q] = c;
    }
    if (bad) atomicOr(mismatch, 1);
}

template <int AXK, bool FACT>
__global__ void __launch_bounds__(256) prepare_samples_kernel(Grid g, const double *__restrict__ rays, int R, int Na,
                                                               int Nt, int Nd, int Ns, int Nsp,
                                                               unsigned char *__restrict__ rec,
                                                               unsigned char *__restrict__ used,
                                                               unsigned long long *oob_count) {
    constexpr int RB = FACT ? 28 : 36;          // bytes per sample
    const int ny = g.ax[1].n, nz = g.ax[2].n;
    const AxisR ax = axis_regs(g.ax[0]), ay = axis_regs(g.ax[1]), az = axis_regs(g.ax[2]);
    const double2 *tabx = g.ax[0].tab, *taby = g.ax[1].tab, *tabz = g.ax[2].tab;
    const bool n_odd = Ns & 1;
    const long long total = (long long)R * Nsp;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned int n_oob = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
        const int q = (int)(k / Nsp), i = (int)(k - (long long)q * Nsp);
        int v = 0;
        double tx = 0.0, ty = 0.0, tz = 0.0, w = 0.0;
        if (i < Ns) {   // the pad samples (i >= Ns) are never read back as data
            const double *rp = rays + prepared_ray_of(q, Na, Nt, Nd) * 4 * Ns;
            const double px = rp[i], py = rp[Ns + i], pz = rp[2 * Ns + i];
            int ix, iy, iz;
            n_oob += locate3<AXK>(tabx, taby, tabz, ax, ay, az, px, py, pz, ix, iy, iz, tx, ty, tz);
            if (!FACT) w = prepared_weight(rp + 3 * Ns, i, Ns, n_odd);
            v = (ix * ny + iy) * nz + iz;
            used[v] = 1;                 // quad records this sample gathers from (benign race: all writers store 1)
            used[v + ny * nz] = 1;
        }
        const int c0 = i / PREP_C * PREP_C, j = i - c0, n4 = min(PREP_C, Nsp - c0);
        unsigned char *blk = rec + ((long long)q * Nsp + c0) * RB;
        double *f = reinterpret_cast<double *>(blk);
        f[j] = tx; f[n4 + j] = ty; f[2 * n4 + j] = tz;
        if (!FACT) f[3 * n4 + j] = w;
        reinterpret_cast<int *>(f + (FACT ? 3 : 4) * n4)[j] = v;
    }
    if (n_oob) atomicAdd(oob_count, (unsigned long long)n_oob);
}

template <bool FACT>
struct PreparedStage {
    static constexpr int RB = FACT ? 28 : 36;
    static constexpr int BYTES = ((PREP_C * RB + 127) / 128) * 128;
};

template <bool FACT, bool BULK, int MAXT, int LAYOUT>
__global__ void __launch_bounds__(MAXT, 1) prepared_forward_kernel(const unsigned char *__restrict__ rec,
                                                                    const double *__restrict__ wscale,
                                                                    const double *__restrict__ pattern,
                                                                    const double *__restrict__ field,
                                                                    double *__restrict__ tec, int R, int Na, int Nt,
                                                                    int Nd, int Ns, int Nsp, int stages, int sy,
                                                                    int sx) {
    constexpr int C = PREP_C, RB = PreparedStage<FACT>::RB, SB = PreparedStage<FACT>::BYTES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // the warp index through a shuffle: the compiler then knows it is warp-uniform, keeps everything derived from it
    // (ring and barrier addresses, the producer's source pointers) in uniform registers and issues the bulk copies
    // without a per-lane address loop
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), nwarp = blockDim.x >> 5;
    // shared: [weight pattern (FACT)][per-warp mbarriers][per-warp stages]
    const double *pat = reinterpret_cast<const double *>(smem_raw);
    unsigned int off = FACT ? ((unsigned int)Ns * 8u + 127u) / 128u * 128u : 0u;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + off) + warp * stages;
    off += ((unsigned int)(nwarp * stages) * 8u + 127u) / 128u * 128u;
    unsigned char *ring = smem_raw + off + (unsigned int)(warp * stages) * SB;
    if (FACT)
        for (int i = threadIdx.x; i < Ns; i += blockDim.x) reinterpret_cast<double *>(smem_raw)[i] = pattern[i];
    if (BULK && lane == 0)
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
    if (BULK) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint32_t ring_s = smem_u32(ring), bars_s = smem_u32(bars);     // shared-window addresses, computed once

    const uint64_t pol_stream = policy_evict_first();
    const int chunks = (Ns + C - 1) / C;
    const int G = gridDim.x;
    // static schedule of the sweep: bundle b = nwarp consecutive slots; this CTA takes bundles
    // blockIdx.x, blockIdx.x + G, ...; this warp takes slot b*nwarp + warp
    const int n_bundles = (R + nwarp - 1) / nwarp;
    int my_rays = (n_bundles > (int)blockIdx.x) ? (n_bundles - 1 - (int)blockIdx.x) / G + 1 : 0;
    if (my_rays > 0 && ((int)blockIdx.x + (my_rays - 1) * G) * nwarp + warp >= R) --my_rays;
    const int q0 = (int)blockIdx.x * nwarp + warp, qstep = G * nwarp;
    const long long ray_bytes = (long long)Nsp * RB;

    // producer (lane 0): the block of (ray slot fq, chunk fc) -> stage fs, ONE bulk copy
    int fk = 0, fc = 0, fs = 0;
    const unsigned char *fsrc = rec + (long long)q0 * ray_bytes;
    auto produce = [&]() {
        if (fk < my_rays) {
            if (elect_one()) {
                const uint32_t bytes = (uint32_t)(min(C, Nsp - fc * C) * RB);
                mbar_expect_tx_s(bars_s + 8u * fs, bytes);
                bulk_g2s_s(ring_s + (uint32_t)SB * fs, fsrc + (long long)fc * (C * RB), bytes, bars_s + 8u * fs, pol_stream);
            }
            fs = (fs + 1 == stages) ? 0 : fs + 1;
            if (++fc == chunks) { fc = 0; ++fk; fsrc += (long long)qstep * ray_bytes; }
        }
    };
    if (BULK)
        for (int s = 0; s < stages - 1; ++s) produce();

    unsigned int phases = 0;   // bit s: parity to wait for on stage s
    int us = 0;                // stage to consume
    // (t, a, d) of the slot, advanced by the digits of qstep instead of three divisions per ray
    int rt = q0 % Nt, ra = (q0 / Nt) % Na, rd = (q0 / Nt) / Na;
    const int st_t = qstep % Nt, st_a = (qstep / Nt) % Na, st_d = (qstep / Nt) / Na;
    for (int k = 0; k < my_rays; ++k) {
        const int q = q0 + k * qstep;
        double acc = 0.0;
        for (int chunk = 0; chunk < chunks; ++chunk) {
            const int c0 = chunk * C;
            const int n4 = min(C, Nsp - c0);
            double *stage = reinterpret_cast<double *>(ring + us * SB);
            if (BULK) {
                produce();
                mbar_wait_s(bars_s + 8u * us, (phases >> us) & 1u);
                phases ^= 1u << us;
            } else {
                const unsigned char *src = rec + (long long)q * ray_bytes + (long long)c0 * RB;
                for (int i = lane; i < n4 * RB / 4; i += 32)
                    reinterpret_cast<int *>(stage)[i] = __ldcs(reinterpret_cast<const int *>(src) + i);
                __syncwarp();
            }
            const double *tx_ = stage, *ty_ = stage + n4, *tz_ = stage + 2 * n4, *w_ = FACT ? pat + c0 : stage + 3 * n4;
            const int *cell_ = reinterpret_cast<const int *>(stage + (FACT ? 3 : 4) * n4);
            const int n_c = min(C, Ns - c0);
            if (n_c == C) {
                // full chunk (every chunk when Ns is a multiple of 64): no per-sample guard, so the gathers of both
                // samples of a lane are issued before the first interpolation waits for its corners
                double f[C / 32];
#pragma unroll
                for (int u = 0; u < C / 32; ++u) {
                    const int j = 32 * u + lane;
                    if (LAYOUT == 1)
                        f[u] = trilerp_quads(reinterpret_cast<const double4 *>(field) + cell_[j], sx, tx_[j], ty_[j], tz_[j]);
                    else
                        f[u] = trilerp(field + cell_[j], sy, sx, tx_[j], ty_[j], tz_[j]);
                }
#pragma unroll
                for (int u = 0; u < C / 32; ++u) acc = fma(w_[32 * u + lane], f[u], acc);
            } else {
#pragma unroll 2
                for (int jb = 0; jb < n_c; jb += 32) {
                    const int j = jb + lane;
                    if (j < n_c) {
                        if (LAYOUT == 1)
                            acc = fma(w_[j], trilerp_quads(reinterpret_cast<const double4 *>(field) + cell_[j], sx, tx_[j],
                                                           ty_[j], tz_[j]), acc);
                        else
                            acc = fma(w_[j], trilerp(field + cell_[j], sy, sx, tx_[j], ty_[j], tz_[j]), acc);
                    }
                }
            }
            us = (us + 1 == stages) ? 0 : us + 1;
            __syncwarp();
        }
        double tot = warp_sum(acc);
        if (lane == 0) {
            if (FACT) tot *= __ldg(wscale + q);
            tec[((long long)ra * Nt + rt) * Nd + rd] = tot;
        }
        rt += st_t;
        if (rt >= Nt) { rt -= Nt; ++ra; }
        ra += st_a;
        if (ra >= Na) { ra -= Na; ++rd; }
        rd += st_d;
    }
}

// flag[u] = 1 for the grid nodes that are a corner of a cell the rays visit: `used` marks the quad records
// r in {v, v + ny*nz}; record r covers the nodes r, r+1, r+nz, r+nz+1
__global__ void __launch_bounds__(256) touched_nodes_kernel(const unsigned char *__restrict__ used, long long V, int ny,
                                                            int nz, unsigned char *__restrict__ flag) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < V; u += stride) {
        const int iz = (int)(u % nz), iy = (int)((u / nz) % ny);
        unsigned char f = used[u];
        if (iz >= 1) f |= used[u - 1];
        if (iy >= 1) f |= used[u - nz];
        if (iz >= 1 && iy >= 1) f |= used[u - nz - 1];
        flag[u] = f ? 1 : 0;
    }
}

// quad records of ne = k * exp(m) for the listed records only (the rays of a shard touch ~12-20 % of the grid)
__global__ void __launch_bounds__(256) quads_list_kernel(const double *__restrict__ m, double k,
                                                          const int *__restrict__ list, long long n, int ny, int nz,
                                                          double4 *__restrict__ q) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int v = list[i];
        const int iz = v % nz, iy = (v / nz) % ny;
        const int dz = (iz + 1 < nz) ? 1 : 0, dy = (iy + 1 < ny) ? nz : 0;
        const double a = exp(m[v]) * k, b = exp(m[v + dz]) * k, c = exp(m[v + dy]) * k, d = exp(m[v + dy + dz]) * k;
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(q + v), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
    }
}

extern "C" long long iono_forwardprojector_n_records(iono_forwardprojector_t h) { return h ? h->n_records : 0; }

// quads_out[v] for the records this projector reads (other records are left untouched): the per-iteration
// replacement of iono_ne_quads_from_m_f64 when only this projector consumes the quad grid.
extern "C" int iono_forwardprojector_quads_from_m_f64(iono_forwardprojector_t h, const double *m, double scale,
                                                      double *quads_out, void *stream) {
    if (!h || !m || !quads_out || ((uintptr_t)quads_out & 31))
        return fail(IONO_EBADARG, "iono_forwardprojector_quads_from_m_f64: bad argument");
    if (device_check(h->device, "iono_forwardprojector_quads_from_m_f64")) return IONO_EBADARG;
    if (h->n_records == 0) return IONO_OK;
    quads_list_kernel<<<ew_grid(h->n_records), 256, 0, (cudaStream_t)stream>>>(
        m, scale, h->records, h->n_records, h->ny, h->nz, reinterpret_cast<double4 *>(quads_out));
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

extern "C" int iono_forwardprojector_destroy(iono_forwardprojector_t h) {
    if (!h) return IONO_OK;
    cudaFree(h->records);
    cudaFree(h->voxels);
    cudaFree(h->rec);
    cudaFree(h->wscale);
    cudaFree(h->pattern);
    delete h;
    return IONO_OK;
}

extern "C" long long iono_forwardprojector_bytes(iono_forwardprojector_t h) {
    if (!h) return 0;
    return h->R * h->Nsp * (h->factored ? 28 : 36) + (h->factored ? (h->R + h->Ns) * 8 : 0);
}

// 1 when the operator stores factored Simpson weights (28 bytes per sample), 0 for per-sample weights (36)
extern "C" int iono_forwardprojector_factored(iono_forwardprojector_t h) { return h ? h->factored : 0; }

template <int AXK>
static void launch_prepare_samples(iono_forwardprojector *h, iono_grid_t grid, const double *rays,
                                   unsigned char *used, unsigned long long *oob_count, cudaStream_t st) {
    const long long n = h->R * h->Nsp;
    if (h->factored)
        prepare_samples_kernel<AXK, true><<<ew_grid(n), 256, 0, st>>>(grid->dev, rays, (int)h->R, h->Na, h->Nt, h->Nd,
                                                                      h->Ns, h->Nsp, h->rec, used, oob_count);
    else
        prepare_samples_kernel<AXK, false><<<ew_grid(n), 256, 0, st>>>(grid->dev, rays, (int)h->R, h->Na, h->Nt, h->Nd,
                                                                       h->Ns, h->Nsp, h->rec, used, oob_count);
}

extern "C" int iono_forwardprojector_create(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns,
                                            iono_forwardprojector_t *out, unsigned long long *oob_count,
                                            void *stream) {
    const long long R = (long long)Na * Nt * Nd;
    if (!grid || !out || !oob_count || Na < 0 || Nt < 0 || Nd < 0 || Ns < 1 || (R > 0 && !rays))
        return fail(IONO_EBADARG, "iono_forwardprojector_create: bad argument");
    if (sweep_size_check(grid, R, Ns)) return IONO_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(oob_count, 0, sizeof(unsigned long long), st));
    iono_forwardprojector *h = new iono_forwardprojector();
    h->rec = nullptr; h->wscale = nullptr; h->pattern = nullptr; h->factored = 0;
    h->records = nullptr; h->n_records = 0; h->voxels = nullptr; h->n_voxels = 0; h->R = R; h->Na = Na; h->Nt = Nt; h->Nd = Nd; h->Ns = Ns;
    h->Nsp = (Ns + 3) / 4 * 4;
    h->nx = grid->nx; h->ny = grid->ny; h->nz = grid->nz;
    cudaGetDevice(&h->device);
    if (R > 0 && Ns >= 2) {   // simps of a single sample is 0: nothing to store
        const long long n = R * h->Nsp;
        const long long V = (long long)grid->nx * grid->ny * grid->nz;
        cudaError_t e = cudaSuccess;
        unsigned char *used = nullptr;
        long long *d_n = nullptr;
        void *tmp = nullptr;
        int *d_bad = nullptr;
        auto bail = [&](const char *what) {
            cudaFree(used); cudaFree(d_n); cudaFree(tmp); cudaFree(d_bad);
            iono_forwardprojector_destroy(h);
            return fail(IONO_ECUDA, "iono_forwardprojector_create: %s: %s", what, cudaGetErrorString(e));
        };
        // common weight pattern x per-ray factor?  (one pass over the s rows)
        const char *env = getenv("IONO_PREP_FACTOR");
        if (!(env && atoi(env) == 0)) {
            if ((e = cudaMalloc(&h->wscale, (size_t)R * sizeof(double))) != cudaSuccess) return bail("cudaMalloc");
            if ((e = cudaMalloc(&h->pattern, (size_t)Ns * sizeof(double))) != cudaSuccess) return bail("cudaMalloc");
            if ((e = cudaMalloc(&d_bad, sizeof(int))) != cudaSuccess) return bail("cudaMalloc");
            if ((e = cudaMemsetAsync(d_bad, 0, sizeof(int), st)) != cudaSuccess) return bail("cudaMemsetAsync");
            weight_pattern_kernel<<<1, 32, 0, st>>>(rays, Na, Nt, Nd, Ns, h->pattern);
            weight_factor_kernel<<<ew_grid(R * 32), 256, 0, st>>>(rays, (int)R, Na, Nt, Nd, Ns, h->pattern, 2e-13,
                                                                  h->wscale, d_bad);
            int bad = 1;
            if ((e = cudaGetLastError()) != cudaSuccess) return bail("launch");
            if ((e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
                (e = cudaStreamSynchronize(st)) != cudaSuccess)
                return bail("weight check");
            cudaFree(d_bad); d_bad = nullptr;
            h->factored = bad ? 0 : 1;
            if (!h->factored) {
                cudaFree(h->wscale); cudaFree(h->pattern);
                h->wscale = nullptr; h->pattern = nullptr;
            }
        }
        if ((e = cudaMalloc(&h->rec, (size_t)n * (h->factored ? 28 : 36))) != cudaSuccess) return bail("cudaMalloc");
        if ((e = cudaMalloc(&used, (size_t)V)) != cudaSuccess) return bail("cudaMalloc");
        if ((e = cudaMemsetAsync(used, 0, (size_t)V, st)) != cudaSuccess) return bail("cudaMemsetAsync");
        if (grid->exact) launch_prepare_samples<2>(h, grid, rays, used, oob_count, st);
        else if (grid->uniform) launch_prepare_samples<1>(h, grid, rays, used, oob_count, st);
        else launch_prepare_samples<0>(h, grid, rays, used, oob_count, st);
        if ((e = cudaGetLastError()) != cudaSuccess) return bail("launch");
        // list of the quad records in use (ascending)
        if ((e = cudaMalloc(&d_n, sizeof(long long))) != cudaSuccess) return bail("cudaMalloc");
        if ((e = cudaMalloc(&h->records, (size_t)V * sizeof(int))) != cudaSuccess) return bail("cudaMalloc");
        size_t tmp_bytes = 0;
        cub::CountingInputIterator<int> ids(0);
        if ((e = cub::DeviceSelect::Flagged(nullptr, tmp_bytes, ids, used, h->records, d_n, (int)V, st)) != cudaSuccess)
            return bail("select");
        if ((e = cudaMalloc(&tmp, tmp_bytes)) != cudaSuccess) return bail("cudaMalloc");
        if ((e = cub::DeviceSelect::Flagged(tmp, tmp_bytes, ids, used, h->records, d_n, (int)V, st)) != cudaSuccess)
            return bail("select");
        if ((e = cudaMemcpyAsync(&h->n_records, d_n, sizeof(long long), cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
            (e = cudaStreamSynchronize(st)) != cudaSuccess)
            return bail("select");
        // list of the grid nodes those records cover (ascending): the support of the adjoint
        {
            unsigned char *flag = nullptr;
            if ((e = cudaMalloc(&flag, (size_t)V)) != cudaSuccess) return bail("cudaMalloc");
            touched_nodes_kernel<<<ew_grid(V), 256, 0, st>>>(used, V, grid->ny, grid->nz, flag);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMalloc(&h->voxels, (size_t)V * sizeof(int));
            if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(tmp, tmp_bytes, ids, flag, h->voxels, d_n, (int)V, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(&h->n_voxels, d_n, sizeof(long long), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            cudaFree(flag);
            if (e != cudaSuccess) return bail("voxel list");
            if (h->n_voxels < V) {
                int *small = nullptr;
                if (cudaMalloc(&small, (size_t)(h->n_voxels > 0 ? h->n_voxels : 1) * sizeof(int)) == cudaSuccess) {
                    cudaMemcpy(small, h->voxels, (size_t)h->n_voxels * sizeof(int), cudaMemcpyDeviceToDevice);
                    cudaFree(h->voxels);
                    h->voxels = small;
                } else {
                    cudaGetLastError();
                }
            }
        }
        cudaFree(used); cudaFree(d_n); cudaFree(tmp);
        // shrink the list to its size
        if (h->n_records < V) {
            int *small = nullptr;
            if (cudaMalloc(&small, (size_t)(h->n_records > 0 ? h->n_records : 1) * sizeof(int)) == cudaSuccess) {
                cudaMemcpy(small, h->records, (size_t)h->n_records * sizeof(int), cudaMemcpyDeviceToDevice);
                cudaFree(h->records);
                h->records = small;
            } else {
                cudaGetLastError();
            }
        }
    }
    *out = h;
    return IONO_OK;
}

template <bool FACT, bool BULK, int MAXT, int LAYOUT>
static int launch_prepared_t(iono_forwardprojector_t h, const double *ne, double *tec, int warps, int stages,
                             size_t smem, int ctas, cudaStream_t st) {
    auto kern = prepared_forward_kernel<FACT, BULK, MAXT, LAYOUT>;
    CU_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ctas, warps * 32, smem, st>>>(h->rec, h->wscale, h->pattern, ne, tec, (int)h->R, h->Na, h->Nt, h->Nd, h->Ns,
                                         h->Nsp, stages, h->nz, h->ny * h->nz);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

template <int LAYOUT>
static int forwardprojector_apply(iono_forwardprojector_t h, const double *field, double *tec_out, cudaStream_t st) {
    if (device_check(h->device, "iono_forwardprojector_apply")) return IONO_EBADARG;
    if (h->R == 0) return IONO_OK;
    if (h->Ns < 2) {
        CU_CHECK(cudaMemsetAsync(tec_out, 0, (size_t)h->R * sizeof(double), st));
        return IONO_OK;
    }
    // 32 warps x 64-sample stages: the kernel waits on its gathers (long-scoreboard stalls), so more warps per SM
    // beat longer chunks (profiles/r02_kernel_bench.json)
    // factored records (1792 B per stage) leave room for a third stage: 0.94 ms against 0.96 ms at the LOFAR case
    int warps = 32, stages = h->factored ? 3 : 2;
    const char *e;
    if ((e = getenv("IONO_PREP_WARPS"))) warps = atoi(e);
    if ((e = getenv("IONO_PREP_STAGES"))) stages = atoi(e);
    if (warps < 1) warps = 1;
    if (warps > 32) warps = 32;
    if (stages < 2) stages = 2;
    if (stages > 8) stages = 8;
    const size_t stage_bytes = h->factored ? PreparedStage<true>::BYTES : PreparedStage<false>::BYTES;
    const size_t pat_bytes = h->factored ? ((size_t)h->Ns * 8 + 127) / 128 * 128 : 0;
    auto smem_for = [&](int w) {
        return pat_bytes + (((size_t)w * stages * sizeof(uint64_t)) + 127) / 128 * 128 + (size_t)w * stages * stage_bytes;
    };
    while (warps > 4 && smem_for(warps) > 227 * 1024) warps -= 4;
    const size_t smem = smem_for(warps);
    if (smem > 227 * 1024) return fail(IONO_EBADARG, "prepared sweep: shared-memory configuration exceeds 227 KB");
    const bool bulk = !getenv("IONO_SWEEP_NO_BULK");   // blocks are multiples of 16 bytes, cudaMalloc bases are aligned
    const int n_bundles = (int)((h->R + warps - 1) / warps);
    int ctas = sm_count();
    if (ctas > n_bundles) ctas = n_bundles;
#define IONO_PREP_DISPATCH(F, B)                                                                                \
    do {                                                                                                         \
        if (warps > 24) return launch_prepared_t<F, B, 1024, LAYOUT>(h, field, tec_out, warps, stages, smem, ctas, st); \
        if (warps > 16) return launch_prepared_t<F, B, 768, LAYOUT>(h, field, tec_out, warps, stages, smem, ctas, st);  \
        return launch_prepared_t<F, B, 512, LAYOUT>(h, field, tec_out, warps, stages, smem, ctas, st);          \
    } while (0)
    if (h->factored) { if (bulk) IONO_PREP_DISPATCH(true, true); else IONO_PREP_DISPATCH(true, false); }
    else             { if (bulk) IONO_PREP_DISPATCH(false, true); else IONO_PREP_DISPATCH(false, false); }
#undef IONO_PREP_DISPATCH
}

extern "C" int iono_forwardprojector_apply_f64(iono_forwardprojector_t h, const double *ne, double *tec_out,
                                               void *stream) {
    if (!h || !ne || (h->R > 0 && !tec_out))
        return fail(IONO_EBADARG, "iono_forwardprojector_apply_f64: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (h->R > 0 && h->Ns >= 2) {
        double4 *q = quads_temporary(ne, h->nx, h->ny, h->nz, h->R * h->Ns, st);
        if (q) {
            int rc = forwardprojector_apply<1>(h, reinterpret_cast<const double *>(q), tec_out, st);
            cudaFreeAsync(q, st);
            return rc;
        }
    }
    return forwardprojector_apply<0>(h, ne, tec_out, st);
}

// the same on the quad layout of ne (iono_quads_from_ne_f64 / iono_ne_quads_from_m_f64)
extern "C" int iono_forwardprojector_apply_quads_f64(iono_forwardprojector_t h, const double *quads, double *tec_out,
                                                     void *stream) {
    if (!h || !quads || ((uintptr_t)quads & 31) || (h->R > 0 && !tec_out))
        return fail(IONO_EBADARG, "iono_forwardprojector_apply_quads_f64: bad argument");
    return forwardprojector_apply<1>(h, quads, tec_out, (cudaStream_t)stream);
}
