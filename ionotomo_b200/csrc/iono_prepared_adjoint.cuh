// Adjoint of the prepared forward operator: the TRANSPOSE of the same record stream ("forward projector" of
// iono_prepared.cuh), walked along the time axis with run aggregation (iono_adjoint_runs.cuh).
// Included by iono_kernels.cu after iono_prepared.cuh.
//
//   acc[v] += sum_ray coef[ray] sum_s w_s(ray) phi_v(x_s(ray))          (reference: inversion/gradient.py:15-54,
//                                                                         the quantity the dask graph sums per ray)
// The forward's records already hold, per sample, the cell, the in-cell fractions and (factored) the Simpson weight:
// no ray rows, no cell search and no weight arithmetic here -- ~65 warp instructions per 32 samples against ~265 in
// the stateless run-aggregated kernel.  The slot order of the records is time-fastest, so the blocks one warp needs
// for (antenna, direction, 64-sample chunk) at consecutive times are Nsp*RB bytes apart and each is ONE bulk copy.
// A lane owns two samples of the chunk; per sample it keeps the 8 corner contributions of the current cell in
// registers while the cell stays the same from one time step to the next and queues what is finished in shared
// memory when it changes (see below: usually one face of the cell, 4 values); 32 queued faces are drained as 4
// full-warp fp64 reductions.  One operator in HBM serves both directions (the voxel-binned back-projector is a second, 6.2 GB,
// copy of the same matrix in voxel order); the price is that the sums are formed by reductions in arrival order, so
// the result is reproducible to rounding (~1e-15 relative), not bitwise.
//
// The accumulator is the full grid and must be zero on entry for the voxels the operator touches; the finishing
// kernels below read it for exactly those voxels (list assembled at create time), apply the chain-rule factor or
// compact it for the cross-GPU sum, and zero it again, so a session never clears or scans the whole grid.
#pragma once

constexpr int PA_SPL = PREP_C / 32;     // samples per lane and block
constexpr int PA_QCAP = 64;             // queue slots per warp
constexpr int PA_QBYTES = PA_QCAP * (4 * 8 + 2 * 4);      // a queued FACE: 4 values, base index, stride

// The unit that is queued and drained is a FACE of a cell -- 4 of its 8 corners: {base, base+1, base+stride,
// base+stride+1} (the z-neighbours are always together; stride = nz for a face normal to x, ny*nz for a face normal
// to y).  From one time step to the next a sample usually moves to a cell that shares a face with the old one
// (the rays of the casting kernels keep z per sample index, so it moves in x or in y): then only the FAR face of the
// old cell is finished -- 4 reductions -- and the shared face's sums are carried over into the new cell's
// accumulators.  A diagonal move is an x move followed by a y move (2 faces), anything else finishes both x faces.
// An SM retires one warp-level fp64 reduction per ~39 cycles whatever the lane count (tools/probes/atomics_probe.cu),
// so the number of reduction INSTRUCTIONS is what bounds a scatter adjoint: 8 per 32 samples x steps in the plain
// kernel, 8 per 32 finished cells with run aggregation (a cell lasts ~4.2 steps at the LOFAR case), ~4 per 32 here.
template <bool FACT, bool BULK, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) prepared_adjoint_kernel(const unsigned char *__restrict__ rec,
                                                                    const double *__restrict__ wscale,
                                                                    const double *__restrict__ pattern,
                                                                    const double *__restrict__ coef_perm,
                                                                    double *__restrict__ acc, int Na, int Nt, int Nd,
                                                                    int Ns, int Nsp, int stages, int sy, int sx,
                                                                    int tsplit) {
    constexpr int C = PREP_C, RB = PreparedStage<FACT>::RB, SB = PreparedStage<FACT>::BYTES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), nwarp = blockDim.x >> 5;
    // shared: [per-warp mbarriers][per-warp rings][per-warp queues: 4 x QCAP doubles, QCAP bases, QCAP strides]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw) + warp * stages;
    unsigned int off = ((unsigned int)(nwarp * stages) * 8u + 127u) / 128u * 128u;
    unsigned char *ring = smem_raw + off + (unsigned int)(warp * stages) * SB;
    off += (unsigned int)(nwarp * stages) * SB;
    double *qval = reinterpret_cast<double *>(smem_raw + off + (unsigned int)warp * PA_QBYTES);      // [4][QCAP]
    int *qbase = reinterpret_cast<int *>(qval + 4 * PA_QCAP);                                          // [QCAP]
    int *qstride = qbase + PA_QCAP;                                                                    // [QCAP]
    if (BULK && lane == 0)
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
    if (BULK) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint32_t ring_s = smem_u32(ring), bars_s = smem_u32(bars);     // shared-window addresses, computed once

    const uint64_t pol_stream = policy_evict_first();
    const int chunks = (Ns + C - 1) / C;
    // a task = (slot block, chunk, time segment): the time axis is cut into `tsplit` segments when there are too few
    // (antenna, direction, chunk) triples to fill the machine (a shard of a multi-GPU job)
    const int tseg = (Nt + tsplit - 1) / tsplit;
    const long long n_tasks = (long long)Na * Nd * chunks * tsplit;
    const long long ray_bytes = (long long)Nsp * RB;
    const unsigned full = 0xffffffffu;
    unsigned int phases = 0;
    int qhead = 0, qtail = 0;      // queue indices (monotone; slot = index % QCAP)

    // drain `count` (<= 32) faces: one per lane, 4 reductions issued by the lanes that hold one
    auto drain = [&](int count) {
        __syncwarp();
        if (lane < count) {
            const int slot = (qhead + lane) % PA_QCAP;
            double *c = acc + qbase[slot];
            const int st = qstride[slot];
            atomicAdd(c, qval[0 * PA_QCAP + slot]);      atomicAdd(c + 1, qval[1 * PA_QCAP + slot]);
            atomicAdd(c + st, qval[2 * PA_QCAP + slot]); atomicAdd(c + st + 1, qval[3 * PA_QCAP + slot]);
        }
        qhead += count;
        __syncwarp();
    };
    // queue slot of this lane among the lanes of ballot `m` that push in this round
    auto slot_of = [&](unsigned m) { return (qtail + __popc(m & ((1u << lane) - 1u))) % PA_QCAP; };
    auto put = [&](int slot, int base, int stride, double f0, double f1, double f2, double f3) {
        qbase[slot] = base;
        qstride[slot] = stride;
        qval[0 * PA_QCAP + slot] = f0; qval[1 * PA_QCAP + slot] = f1;
        qval[2 * PA_QCAP + slot] = f2; qval[3 * PA_QCAP + slot] = f3;
    };
    // Leave cell `vc` for cell `vc + delta` (whole == true: leave it for good).  a[4x + 2y + z] are the sums of the
    // corner (x,y,z) of the current cell; on return they are the sums the NEW cell inherits (0 where it shares nothing).
    // All samples of a warp belong to one ray bundle and drift the same way, so the branches on the sign of the move
    // are nearly warp-uniform.
    auto leave = [&](bool moved, bool whole, int vc, int delta, double (&a)[8]) {
        int dx = 0, dy = 0;
        bool jump = whole;
        if (moved && !whole) {
            dx = (abs(delta - sx) <= sy) ? 1 : ((abs(delta + sx) <= sy) ? -1 : 0);
            const int ry = delta - dx * sx;
            dy = (ry == sy) ? 1 : ((ry == -sy) ? -1 : 0);
            jump = (ry != dy * sy);
        }
        moved = moved || whole;
        // round A: a face normal to x -- the far one of an x move, or the low-x half of a cell that is left for good
        {
            const bool go = moved && (jump || dx != 0);
            const unsigned m = __ballot_sync(full, go);
            if (m) {
                if (go) {
                    const int slot = slot_of(m);
                    if (jump) {
                        put(slot, vc, sy, a[0], a[1], a[2], a[3]);
                        a[0] = a[1] = a[2] = a[3] = 0.0;
                    } else if (dx > 0) {
                        put(slot, vc, sy, a[0], a[1], a[2], a[3]);
                        a[0] = a[4]; a[1] = a[5]; a[2] = a[6]; a[3] = a[7]; a[4] = a[5] = a[6] = a[7] = 0.0;
                    } else {
                        put(slot, vc + sx, sy, a[4], a[5], a[6], a[7]);
                        a[4] = a[0]; a[5] = a[1]; a[6] = a[2]; a[7] = a[3]; a[0] = a[1] = a[2] = a[3] = 0.0;
                    }
                }
                qtail += __popc(m);
                if (qtail - qhead >= 32) drain(32);
            }
        }
        // round B: the far face of a y move (normal to y, relative to the cell after the x move), or the high-x half
        {
            const bool go = moved && (jump || dy != 0);
            const unsigned m = __ballot_sync(full, go);
            if (m) {
                if (go) {
                    const int slot = slot_of(m);
                    const int vm = vc + dx * sx;
                    if (jump) {
                        put(slot, vc + sx, sy, a[4], a[5], a[6], a[7]);
                        a[4] = a[5] = a[6] = a[7] = 0.0;
                    } else if (dy > 0) {
                        put(slot, vm, sx, a[0], a[1], a[4], a[5]);
                        a[0] = a[2]; a[1] = a[3]; a[4] = a[6]; a[5] = a[7]; a[2] = a[3] = a[6] = a[7] = 0.0;
                    } else {
                        put(slot, vm + sy, sx, a[2], a[3], a[6], a[7]);
                        a[2] = a[0]; a[3] = a[1]; a[6] = a[4]; a[7] = a[5]; a[0] = a[1] = a[4] = a[5] = 0.0;
                    }
                }
                qtail += __popc(m);
                if (qtail - qhead >= 32) drain(32);
            }
        }
    };

    // tasks are dealt to the CTAs round-robin (task -> CTA task % G, then warp), so that a shard with fewer tasks than
    // warps (1/8 of the LOFAR case: 3 100 tasks, 3 552 warps) still loads every SM equally: the kernel is bound by the
    // instruction issue of an SM, i.e. by the number of tasks it hosts
    for (long long task = (long long)warp * gridDim.x + blockIdx.x; task < n_tasks; task += (long long)gridDim.x * nwarp) {
        // chunk index slowest: the far chunks of a ray move through more cells than the near ones, and with the chunk
        // fastest a CTA of an even-sized grid would only ever see one kind
        const int seg = (int)(task % tsplit);
        const int ad = (int)((task / tsplit) % ((long long)Na * Nd));         // slot block: d * Na + a (time fastest inside)
        const int c = (int)(task / ((long long)tsplit * Na * Nd));
        const int a_ = ad % Na, d_ = ad / Na;
        const int t_lo = seg * tseg, n_t = min(Nt, t_lo + tseg) - t_lo;       // this task: time steps t_lo .. t_lo + n_t - 1
        const int c0 = c * C;
        const int n4 = min(C, Nsp - c0), n_c = min(C, Ns - c0);
        const long long q0 = (long long)ad * Nt + t_lo;
        const unsigned char *src0 = rec + q0 * ray_bytes + (long long)c0 * RB;
        const double *cp = coef_perm + ((long long)a_ * Nd + d_) * Nt + t_lo;
        const uint32_t bytes = (uint32_t)(n4 * RB);
        // ring positions advance by counters (stages is a run-time value: t % stages would be a division per step)
        int ps = 0, us = 0;       // stage the next produce fills / the next step consumes
        const unsigned char *src_next = src0;        // block of the next time step to fetch
        auto produce = [&](int t) {
            if (t < n_t && elect_one()) {
                mbar_expect_tx_s(bars_s + 8u * ps, bytes);
                bulk_g2s_s(ring_s + (uint32_t)SB * ps, src_next, bytes, bars_s + 8u * ps, pol_stream);
            }
            src_next += ray_bytes;
            ps = (ps + 1 == stages) ? 0 : ps + 1;
        };
        if (BULK)
            for (int t = 0; t < stages - 1; ++t) produce(t);
        // lanes beyond the end of a partial chunk shadow its last sample with weight 0 (no per-sample guards in the
        // time loop; their faces carry zeros)
        int jj[PA_SPL];
        double pw[PA_SPL];
        int v_cur[PA_SPL];
        double a[PA_SPL][8];
#pragma unroll
        for (int s = 0; s < PA_SPL; ++s) {
            const int j = lane + 32 * s;
            jj[s] = min(j, n_c - 1);
            pw[s] = (j < n_c) ? (FACT ? __ldg(pattern + c0 + j) : 1.0) : 0.0;
            v_cur[s] = -1;
#pragma unroll
            for (int e = 0; e < 8; ++e) a[s][e] = 0.0;
        }
        double cw = 0.0;
        for (int t = 0; t < n_t; ++t) {
            if ((t & 31) == 0) {      // coefficients (x the per-ray weight factor) of the next 32 time steps, one per lane
                cw = 0.0;
                if (t + lane < n_t) {
                    cw = __ldg(cp + t + lane);
                    if (FACT) cw *= __ldg(wscale + q0 + t + lane);
                }
            }
            const double coef = __shfl_sync(full, cw, t & 31);
            double *stage = reinterpret_cast<double *>(ring + us * SB);
            if (BULK) {
                produce(t + stages - 1);
                mbar_wait_s(bars_s + 8u * us, (phases >> us) & 1u);
                phases ^= 1u << us;
            } else {
                const unsigned char *src = src0 + (long long)t * ray_bytes;
                for (int i = lane; i < (int)bytes / 4; i += 32)
                    reinterpret_cast<int *>(stage)[i] = __ldcs(reinterpret_cast<const int *>(src) + i);
                __syncwarp();
            }
            const double *tx_ = stage, *ty_ = stage + n4, *tz_ = stage + 2 * n4, *w_ = stage + 3 * n4;
            const int *cell_ = reinterpret_cast<const int *>(stage + (FACT ? 3 : 4) * n4);
#pragma unroll
            for (int s = 0; s < PA_SPL; ++s) {
                const int j = jj[s];
                double l[8];
                const int v = cell_[j];
                {
                    const double tx = tx_[j], ty = ty_[j], tz = tz_[j];
                    const double aw = FACT ? coef * pw[s] : (coef * pw[s]) * w_[j];
                    const double ax1 = aw * tx, ax0 = aw - ax1;
                    const double a01 = ax0 * ty, a00 = ax0 - a01;
                    const double a11 = ax1 * ty, a10 = ax1 - a11;
                    l[1] = a00 * tz; l[3] = a01 * tz; l[5] = a10 * tz; l[7] = a11 * tz;
                    l[0] = a00 - l[1]; l[2] = a01 - l[3]; l[4] = a10 - l[5]; l[6] = a11 - l[7];
                }
                leave((v != v_cur[s]) && (v_cur[s] >= 0), false, v_cur[s], v - v_cur[s], a[s]);
                v_cur[s] = v;
#pragma unroll
                for (int e = 0; e < 8; ++e) a[s][e] += l[e];
            }
            us = (us + 1 == stages) ? 0 : us + 1;
            __syncwarp();     // the stage is free for the producer again
        }
#pragma unroll
        for (int s = 0; s < PA_SPL; ++s) leave(false, v_cur[s] >= 0, v_cur[s], 0, a[s]);    // what the lanes still hold
    }
    while (qtail > qhead) drain(min(32, qtail - qhead));
}

// grad[v] = k exp(m[v]) acc[v], acc[v] = 0 for the listed voxels
__global__ void __launch_bounds__(256) finish_gradient_kernel(const int *__restrict__ voxels, long long n,
                                                              double *__restrict__ acc, const double *__restrict__ m,
                                                              double k, double *__restrict__ grad) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int v = voxels[i];
        grad[v] = k * exp(m[v]) * acc[v];
        acc[v] = 0.0;
    }
}

// out[dst[i]] = acc[v], acc[v] = 0 for the listed voxels (dst == NULL: out[i])
__global__ void __launch_bounds__(256) finish_compact_kernel(const int *__restrict__ voxels, long long n,
                                                             double *__restrict__ acc, const unsigned int *__restrict__ dst,
                                                             double *__restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int v = voxels[i];
        out[dst ? (long long)dst[i] : i] = acc[v];
        acc[v] = 0.0;
    }
}

extern "C" long long iono_forwardprojector_n_voxels(iono_forwardprojector_t h) { return h ? h->n_voxels : 0; }

// grid nodes the operator touches, ascending (device array of n_voxels int32)
extern "C" int iono_forwardprojector_voxels(iono_forwardprojector_t h, int *out, void *stream) {
    if (!h || (h->n_voxels > 0 && !out)) return fail(IONO_EBADARG, "iono_forwardprojector_voxels: bad argument");
    if (h->n_voxels > 0)
        CU_CHECK(cudaMemcpyAsync(out, h->voxels, (size_t)h->n_voxels * sizeof(int), cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream));
    return IONO_OK;
}

template <bool FACT, bool BULK, int MAXT>
static int launch_prepared_adjoint_t(iono_forwardprojector_t h, const double *coef_perm, double *acc, int warps,
                                     int stages, size_t smem, int ctas, int tsplit, cudaStream_t st) {
    auto kern = prepared_adjoint_kernel<FACT, BULK, MAXT>;
    CU_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ctas, warps * 32, smem, st>>>(h->rec, h->wscale, h->pattern, coef_perm, acc, h->Na, h->Nt, h->Nd, h->Ns,
                                         h->Nsp, stages, h->nz, h->ny * h->nz, tsplit);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

// acc[v] += sum_ray coef_perm[(a*Nd + d)*Nt + t] * A[ray, v] with A the matrix iono_forwardprojector_apply applies
// (coefficients in the time-fastest order iono_residual_f64 writes as `coef_perm`); acc: the full (nx,ny,nz) grid.
extern "C" int iono_forwardprojector_adjoint_f64(iono_forwardprojector_t h, const double *coef_perm, double *acc,
                                                 void *stream) {
    if (!h || (h->R > 0 && (!coef_perm || !acc))) return fail(IONO_EBADARG, "iono_forwardprojector_adjoint_f64: bad argument");
    if (device_check(h->device, "iono_forwardprojector_adjoint_f64")) return IONO_EBADARG;
    if (h->R == 0 || h->Ns < 2) return IONO_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // the kernel is bound by instruction issue with few warps and by its register budget with many: 24 warps x 80
    // registers (a few spilled words) is the measured best -- 1.20 ms against 1.46 ms with 16 x 103 and 1.6 ms with
    // 32 x 64 at the LOFAR case (profiles/r02_kernel_bench.json)
    int warps = 24, stages = 2;
    const char *e;
    if ((e = getenv("IONO_PADJ_WARPS"))) warps = atoi(e);
    if ((e = getenv("IONO_PADJ_STAGES"))) stages = atoi(e);
    if (warps < 1) warps = 1;
    if (warps > 32) warps = 32;
    if (stages < 2) stages = 2;
    if (stages > 8) stages = 8;
    const size_t stage_bytes = h->factored ? PreparedStage<true>::BYTES : PreparedStage<false>::BYTES;
    auto smem_for = [&](int w, int s) {
        return (((size_t)w * s * sizeof(uint64_t)) + 127) / 128 * 128 + (size_t)w * s * stage_bytes + (size_t)w * PA_QBYTES;
    };
    while (stages > 2 && smem_for(warps, stages) > 227 * 1024) --stages;
    while (warps > 1 && smem_for(warps, stages) > 227 * 1024) --warps;
    const size_t smem = smem_for(warps, stages);
    if (smem > 227 * 1024) return fail(IONO_EBADARG, "prepared adjoint: shared-memory configuration exceeds 227 KB");
    const bool bulk = !getenv("IONO_SWEEP_NO_BULK");
    long long n_tasks = (long long)h->Na * h->Nd * ((h->Ns + PREP_C - 1) / PREP_C);
    // IONO_PADJ_TSPLIT cuts the time axis of every task into segments (more, shorter tasks for a small shard).  Off by
    // default: every segment ends with a whole-cell flush, and at 1/8 of the LOFAR case (0.87 tasks per warp) 2, 4 and
    // 8 segments measured 0.186, 0.191 and 0.209 ms against 0.182 ms unsplit (profiles/README.md)
    int tsplit = 1;
    if ((e = getenv("IONO_PADJ_TSPLIT"))) {
        tsplit = atoi(e);
        if (tsplit < 1) tsplit = 1;
        if (tsplit > h->Nt) tsplit = h->Nt;
    }
    n_tasks *= tsplit;
    int ctas = sm_count();            // (tasks are dealt round-robin: every SM gets its share even when tasks < warps)
    if (ctas > n_tasks) ctas = (int)n_tasks;
#define IONO_PADJ_DISPATCH(F, B)                                                                              \
    do {                                                                                                       \
        if (warps > 28) return launch_prepared_adjoint_t<F, B, 1024>(h, coef_perm, acc, warps, stages, smem, ctas, tsplit, st); \
        if (warps > 24) return launch_prepared_adjoint_t<F, B, 896>(h, coef_perm, acc, warps, stages, smem, ctas, tsplit, st); \
        if (warps > 16) return launch_prepared_adjoint_t<F, B, 768>(h, coef_perm, acc, warps, stages, smem, ctas, tsplit, st);  \
        return launch_prepared_adjoint_t<F, B, 512>(h, coef_perm, acc, warps, stages, smem, ctas, tsplit, st);         \
    } while (0)
    if (h->factored) { if (bulk) IONO_PADJ_DISPATCH(true, true); else IONO_PADJ_DISPATCH(true, false); }
    else             { if (bulk) IONO_PADJ_DISPATCH(false, true); else IONO_PADJ_DISPATCH(false, false); }
#undef IONO_PADJ_DISPATCH
}

// grad[v] = k * exp(m[v]) * acc[v] and acc[v] = 0 for the voxels the operator touches (grad elsewhere untouched)
extern "C" int iono_forwardprojector_finish_gradient_f64(iono_forwardprojector_t h, double *acc, const double *m,
                                                         double k, double *grad, void *stream) {
    if (!h || (h->n_voxels > 0 && (!acc || !m || !grad)))
        return fail(IONO_EBADARG, "iono_forwardprojector_finish_gradient_f64: bad argument");
    if (device_check(h->device, "iono_forwardprojector_finish_gradient_f64")) return IONO_EBADARG;
    if (h->n_voxels == 0) return IONO_OK;
    finish_gradient_kernel<<<ew_grid(h->n_voxels), 256, 0, (cudaStream_t)stream>>>(h->voxels, h->n_voxels, acc, m, k, grad);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

// out[dst[i]] (dst NULL: out[i]) = acc[voxel i] and acc[voxel i] = 0: the compact accumulator of the sharded adjoint
extern "C" int iono_forwardprojector_finish_compact_f64(iono_forwardprojector_t h, double *acc, const unsigned int *dst,
                                                        double *out, void *stream) {
    if (!h || (h->n_voxels > 0 && (!acc || !out)))
        return fail(IONO_EBADARG, "iono_forwardprojector_finish_compact_f64: bad argument");
    if (device_check(h->device, "iono_forwardprojector_finish_compact_f64")) return IONO_EBADARG;
    if (h->n_voxels == 0) return IONO_OK;
    finish_compact_kernel<<<ew_grid(h->n_voxels), 256, 0, (cudaStream_t)stream>>>(h->voxels, h->n_voxels, acc, dst, out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}
