// Prepared forward sweep ("forward projector"): the TEC forward for a ray geometry that is reused
// across iterations (every driver of the reference calls forward_equation with the same rays in each
// iteration: bfgs_dask.py:207-340, iterative_newton.py:954-1017, tests/test_inversion.py:30-39).
// Included by iono_kernels.cu.  The forward-side twin of the voxel-binned back-projector.
//
//   assembly (once per geometry): per sample the cell index v, the in-cell fractions (tx,ty,tz) and
//     the Simpson 'avg' weight w -- computed by the SAME device functions the stateless sweep
//     (iono_sweep.cuh, MODE 0) evaluates per launch -- are written as a 36-byte record stream, rays
//     stored in the sweep's time-fastest traversal order;
//   apply: tec[ray] = sum_s w_s * trilerp(ne; v_s, t_s).  One warp per ray, lanes = consecutive
//     samples, records streamed by TMA 1-D bulk copies into a warp-private ring (as the sweep), but no
//     cell search, no table reads and no weight arithmetic in the loop: 8 corner gathers, 14 fp64
//     operations and one fma per sample.
//
// v, t and w are bit-identical to what MODE 0 derives and the per-lane accumulation order is the
// same, so the result is bit-identical to iono_tec_forward_f64 (tested).  The stream is 36 B per
// sample against the 32 B of the raw ray rows; the roofline fraction is still quoted on the
// algorithmic bytes of the reference's API boundary.
#pragma once
#include <cub/cub.cuh>

struct iono_forwardprojector {
    int *cell;        // (R, Nsp) flat index of the cell's low corner
    double *frac;     // (R, 4, Nsp) rows tx, ty, tz, w
    int *records;     // quad records (iono_device.cuh) the samples read: cells v and v + ny*nz, ascending
    long long n_records;
    long long R;
    int Na, Nt, Nd, Ns, Nsp;   // Nsp = Ns rounded up to a multiple of 4 (16-byte rows for the bulk copies)
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

template <int AXK>
__global__ void __launch_bounds__(256) prepare_samples_kernel(Grid g, const double *__restrict__ rays, int R, int Na,
                                                               int Nt, int Nd, int Ns, int Nsp,
                                                               int *__restrict__ cell, double *__restrict__ frac,
                                                               unsigned char *__restrict__ used,
                                                               unsigned long long *oob_count) {
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
            const double *sp = rp + 3 * Ns;
            const double px = rp[i], py = rp[Ns + i], pz = rp[2 * Ns + i];
            int ix, iy, iz;
            n_oob += locate3<AXK>(tabx, taby, tabz, ax, ay, az, px, py, pz, ix, iy, iz, tx, ty, tz);
            // neighbours outside [0, Ns) are never used by simpson_weight; clamp the reads
            w = simpson_weight(i, Ns, n_odd, sp[max(i - 2, 0)], sp[max(i - 1, 0)], sp[i], sp[min(i + 1, Ns - 1)],
                               sp[min(i + 2, Ns - 1)]);
            v = (ix * ny + iy) * nz + iz;
            used[v] = 1;                 // quad records this sample gathers from (benign race: all writers store 1)
            used[v + ny * nz] = 1;
        }
        cell[k] = v;
        double *f = frac + (long long)q * 4 * Nsp + i;
        f[0] = tx; f[Nsp] = ty; f[2 * Nsp] = tz; f[3 * Nsp] = w;
    }
    if (n_oob) atomicAdd(oob_count, (unsigned long long)n_oob);
}

// A stage holds tx[C], ty[C], tz[C], w[C] and cell[C] (int32).
template <int C>
struct PreparedStage {
    static constexpr int CELL_OFF = 4 * C;   // in doubles
    static constexpr int BYTES = ((4 * C * 8 + C * 4 + 127) / 128) * 128;
};

// n_cp: samples to copy (a multiple of 4, so every piece is a multiple of 16 bytes)
template <int C, bool BULK>
__device__ __forceinline__ void fill_prepared(double *stage, uint64_t *bar, const double *frac_q, const int *cell_q,
                                              int Nsp, int c0, int n_cp, int lane, uint64_t policy) {
    if (BULK) {
        if (lane == 0) {
            mbar_expect_tx(bar, (uint32_t)(n_cp * 36));
            bulk_g2s(stage, frac_q + c0, n_cp * 8, bar, policy);
            bulk_g2s(stage + C, frac_q + Nsp + c0, n_cp * 8, bar, policy);
            bulk_g2s(stage + 2 * C, frac_q + 2 * Nsp + c0, n_cp * 8, bar, policy);
            bulk_g2s(stage + 3 * C, frac_q + 3 * Nsp + c0, n_cp * 8, bar, policy);
            bulk_g2s(stage + PreparedStage<C>::CELL_OFF, cell_q + c0, n_cp * 4, bar, policy);
        }
    } else {
        int *cdst = reinterpret_cast<int *>(stage + PreparedStage<C>::CELL_OFF);
        for (int i = lane; i < n_cp; i += 32) {
            stage[i] = ld_stream(frac_q + c0 + i, policy);
            stage[C + i] = ld_stream(frac_q + Nsp + c0 + i, policy);
            stage[2 * C + i] = ld_stream(frac_q + 2 * Nsp + c0 + i, policy);
            stage[3 * C + i] = ld_stream(frac_q + 3 * Nsp + c0 + i, policy);
            cdst[i] = __ldcs(cell_q + c0 + i);
        }
    }
}

template <int C, bool BULK, int MAXT, int LAYOUT>
__global__ void __launch_bounds__(MAXT, 1) prepared_forward_kernel(const double *__restrict__ frac,
                                                                    const int *__restrict__ cell,
                                                                    const double *__restrict__ field,
                                                                    double *__restrict__ tec, int R, int Na, int Nt,
                                                                    int Nd, int Ns, int Nsp, int stages, int sy,
                                                                    int sx) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    // shared: [per-warp mbarriers][per-warp stages]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw) + warp * stages;
    const unsigned int off = ((unsigned int)(nwarp * stages) * 8u + 127u) / 128u * 128u;
    unsigned char *ring = smem_raw + off + (unsigned int)(warp * stages) * PreparedStage<C>::BYTES;
    if (BULK && lane == 0)
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
    if (BULK) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint64_t pol_stream = policy_evict_first();
    const int chunks = (Ns + C - 1) / C;
    const int G = gridDim.x;
    // static schedule of the sweep: bundle b = nwarp consecutive slots; this CTA takes bundles
    // blockIdx.x, blockIdx.x + G, ...; this warp takes slot b*nwarp + warp
    const int n_bundles = (R + nwarp - 1) / nwarp;
    int my_rays = (n_bundles > (int)blockIdx.x) ? (n_bundles - 1 - (int)blockIdx.x) / G + 1 : 0;
    if (my_rays > 0 && ((int)blockIdx.x + (my_rays - 1) * G) * nwarp + warp >= R) --my_rays;
    const int q0 = (int)blockIdx.x * nwarp + warp, qstep = G * nwarp;

    // producer cursor (ray fk = slot fq, chunk fc, stage fs)
    int fk = 0, fc = 0, fs = 0, fq = q0;
    auto produce = [&]() {
        if (fk < my_rays) {
            const int c0 = fc * C;
            fill_prepared<C, true>(reinterpret_cast<double *>(ring + fs * PreparedStage<C>::BYTES), &bars[fs],
                                   frac + (long long)fq * 4 * Nsp, cell + (long long)fq * Nsp, Nsp, c0,
                                   min(C, Nsp - c0), lane, pol_stream);
            fs = (fs + 1 == stages) ? 0 : fs + 1;
            if (++fc == chunks) { fc = 0; ++fk; fq += qstep; }
        }
    };
    if (BULK)
        for (int s = 0; s < stages - 1; ++s) produce();

    unsigned int phases = 0;   // bit s: parity to wait for on stage s
    int us = 0;                // stage to consume
    for (int k = 0; k < my_rays; ++k) {
        const int q = q0 + k * qstep;
        double acc = 0.0;
        for (int chunk = 0; chunk < chunks; ++chunk) {
            const int c0 = chunk * C;
            double *stage = reinterpret_cast<double *>(ring + us * PreparedStage<C>::BYTES);
            if (BULK) {
                produce();
                mbar_wait(&bars[us], (phases >> us) & 1u);
                phases ^= 1u << us;
            } else {
                fill_prepared<C, false>(stage, nullptr, frac + (long long)q * 4 * Nsp, cell + (long long)q * Nsp, Nsp,
                                        c0, min(C, Nsp - c0), lane, pol_stream);
                __syncwarp();
            }
            const double *tx_ = stage, *ty_ = stage + C, *tz_ = stage + 2 * C, *w_ = stage + 3 * C;
            const int *cell_ = reinterpret_cast<const int *>(stage + PreparedStage<C>::CELL_OFF);
            const int n_c = min(C, Ns - c0);
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
            us = (us + 1 == stages) ? 0 : us + 1;
            __syncwarp();
        }
        const double tot = warp_sum(acc);
        if (lane == 0) tec[prepared_ray_of(q, Na, Nt, Nd)] = tot;
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
    cudaFree(h->cell);
    cudaFree(h->frac);
    delete h;
    return IONO_OK;
}

extern "C" long long iono_forwardprojector_bytes(iono_forwardprojector_t h) {
    return h ? h->R * h->Nsp * 36 : 0;
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
    h->cell = nullptr; h->frac = nullptr; h->records = nullptr; h->n_records = 0; h->R = R; h->Na = Na; h->Nt = Nt; h->Nd = Nd; h->Ns = Ns;
    h->Nsp = (Ns + 3) / 4 * 4;
    h->nx = grid->nx; h->ny = grid->ny; h->nz = grid->nz;
    cudaGetDevice(&h->device);
    if (R > 0 && Ns >= 2) {   // simps of a single sample is 0: nothing to store
        const long long n = R * h->Nsp;
        cudaError_t e = cudaMalloc(&h->cell, (size_t)n * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&h->frac, (size_t)n * 4 * sizeof(double));
        if (e != cudaSuccess) {
            iono_forwardprojector_destroy(h);
            return fail(IONO_ECUDA, "iono_forwardprojector_create: cudaMalloc: %s", cudaGetErrorString(e));
        }
        const long long V = (long long)grid->nx * grid->ny * grid->nz;
        unsigned char *used = nullptr;
        long long *d_n = nullptr;
        void *tmp = nullptr;
        auto bail = [&](const char *what) {
            cudaFree(used); cudaFree(d_n); cudaFree(tmp);
            iono_forwardprojector_destroy(h);
            return fail(IONO_ECUDA, "iono_forwardprojector_create: %s: %s", what, cudaGetErrorString(e));
        };
        if ((e = cudaMalloc(&used, (size_t)V)) != cudaSuccess) return bail("cudaMalloc");
        if ((e = cudaMemsetAsync(used, 0, (size_t)V, st)) != cudaSuccess) return bail("cudaMemsetAsync");
        if (grid->exact)
            prepare_samples_kernel<2><<<ew_grid(n), 256, 0, st>>>(grid->dev, rays, (int)R, Na, Nt, Nd, Ns, h->Nsp,
                                                                  h->cell, h->frac, used, oob_count);
        else if (grid->uniform)
            prepare_samples_kernel<1><<<ew_grid(n), 256, 0, st>>>(grid->dev, rays, (int)R, Na, Nt, Nd, Ns, h->Nsp,
                                                                  h->cell, h->frac, used, oob_count);
        else
            prepare_samples_kernel<0><<<ew_grid(n), 256, 0, st>>>(grid->dev, rays, (int)R, Na, Nt, Nd, Ns, h->Nsp,
                                                                  h->cell, h->frac, used, oob_count);
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

template <int C, bool BULK, int MAXT, int LAYOUT>
static int launch_prepared_t(iono_forwardprojector_t h, const double *ne, double *tec, int warps, int stages,
                             size_t smem, int ctas, cudaStream_t st) {
    auto kern = prepared_forward_kernel<C, BULK, MAXT, LAYOUT>;
    CU_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ctas, warps * 32, smem, st>>>(h->frac, h->cell, ne, tec, (int)h->R, h->Na, h->Nt, h->Nd, h->Ns, h->Nsp,
                                         stages, h->nz, h->ny * h->nz);
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
    // 32 warps x 64-sample chunks: the kernel waits on its gathers (long-scoreboard stalls), so more warps per SM
    // beat longer chunks -- 1.18 ms against 1.26 ms for 24 x 128 at the LOFAR case (profiles/r02_kernel_bench.json)
    int warps = 32, stages = 2, chunk = 64;
    const char *e;
    if ((e = getenv("IONO_PREP_WARPS"))) warps = atoi(e);
    if ((e = getenv("IONO_PREP_STAGES"))) stages = atoi(e);
    if ((e = getenv("IONO_PREP_CHUNK"))) chunk = atoi(e);
    if (warps < 1) warps = 1;
    if (warps > 32) warps = 32;
    if (stages < 2) stages = 2;
    if (stages > 8) stages = 8;
    if (chunk != 64) chunk = 128;
    const size_t stage_bytes = chunk == 64 ? PreparedStage<64>::BYTES : PreparedStage<128>::BYTES;
    auto smem_for = [&](int w) {
        return (((size_t)w * stages * sizeof(uint64_t)) + 127) / 128 * 128 + (size_t)w * stages * stage_bytes;
    };
    while (warps > 4 && smem_for(warps) > 227 * 1024) warps -= 4;
    const size_t smem = smem_for(warps);
    if (smem > 227 * 1024) return fail(IONO_EBADARG, "prepared sweep: shared-memory configuration exceeds 227 KB");
    const bool bulk = !getenv("IONO_SWEEP_NO_BULK");   // rows are padded to 16 bytes, cudaMalloc bases are aligned
    const int n_bundles = (int)((h->R + warps - 1) / warps);
    int ctas = sm_count();
    if (ctas > n_bundles) ctas = n_bundles;
#define IONO_PREP_DISPATCH(CC, B)                                                                                \
    do {                                                                                                         \
        if (warps > 24) return launch_prepared_t<CC, B, 1024, LAYOUT>(h, field, tec_out, warps, stages, smem, ctas, st); \
        if (warps > 16) return launch_prepared_t<CC, B, 768, LAYOUT>(h, field, tec_out, warps, stages, smem, ctas, st);  \
        return launch_prepared_t<CC, B, 512, LAYOUT>(h, field, tec_out, warps, stages, smem, ctas, st);          \
    } while (0)
    if (chunk == 64) { if (bulk) IONO_PREP_DISPATCH(64, true); else IONO_PREP_DISPATCH(64, false); }
    else             { if (bulk) IONO_PREP_DISPATCH(128, true); else IONO_PREP_DISPATCH(128, false); }
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
