// Stateless adjoint with run aggregation ("time-bundled scatter").  Included by iono_kernels.cu after iono_sweep.cuh.
//
// The plain scatter adjoint (ray_sweep_kernel<1,...>) issues 8 fp64 reductions per sample and is bound by the rate at
// which an SM retires global reductions (7.2 ms at the LOFAR case, 0.11 of the HBM roofline).  The rays of one
// (antenna, direction) at consecutive times are nearly identical: a given sample index stays in the same grid cell
// for ~9 consecutive time steps (the run structure the binned operator's index compression exploits,
// iono_backproject.cuh).  So here a WARP owns (antenna, direction, 32 consecutive samples) and walks the TIME axis:
// every lane keeps the 8 corner contributions of its current cell in registers and adds to them while the cell stays
// the same; when the cell changes, the finished (cell, 8 values) tuple goes into a warp-private shared-memory queue,
// and whenever 32 tuples are waiting the warp drains them with 8 FULL reduction instructions (one tuple per lane).
// ~9x fewer reductions, all of them issued with 32 active lanes.  Nothing is assumed about the rays: for arbitrary
// input every step simply changes cell and the kernel degenerates to the plain scatter.
//
// Rows are streamed by the same TMA 1-D bulk copies as the sweep (one 32-sample chunk of the four rows per stage,
// 4-deep ring per warp); cell lookup, Simpson weights and the trilinear split are the sweep's device functions, so the
// result equals iono_tec_adjoint_f64's plain kernel up to the order of the additions.
#pragma once

constexpr int AR_C = 32;          // samples per task and stage
constexpr int AR_STAGES = 4;      // ring depth (time steps in flight per warp)
constexpr int AR_QCAP = 64;       // queue slots per warp

struct AdjRunsParams {
    Grid g;
    const double *rays;
    const double *coef;
    double *acc;
    unsigned long long *oob_count;
    int Na, Nt, Nd, Ns;
};

template <int AXK, bool BULK, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) adjoint_runs_kernel(const AdjRunsParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // the warp index through a shuffle: the compiler then knows it is warp-uniform, keeps everything derived from it
    // (ring and barrier addresses, the producer's source pointers) in uniform registers and issues the bulk copies
    // without a per-lane address loop
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), nwarp = blockDim.x >> 5;
    const int nx = p.g.ax[0].n, ny = p.g.ax[1].n, nz = p.g.ax[2].n;
    // shared: [axis tables (AXK != 2)][per-warp mbarriers][per-warp rings][per-warp queues: 8 x QCAP doubles + QCAP ints]
    double2 *tabx = reinterpret_cast<double2 *>(smem_raw);
    double2 *taby = tabx + nx;
    double2 *tabz = taby + ny;
    unsigned int off = (AXK == 2) ? 0u : ((unsigned int)(nx + ny + nz) * 16u + 127u) / 128u * 128u;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + off) + warp * AR_STAGES;
    off += ((unsigned int)(nwarp * AR_STAGES) * 8u + 127u) / 128u * 128u;
    unsigned char *ring = smem_raw + off + (unsigned int)(warp * AR_STAGES) * StageLayout<AR_C>::BYTES;
    off += (unsigned int)(nwarp * AR_STAGES) * StageLayout<AR_C>::BYTES;
    constexpr unsigned int QBYTES = AR_QCAP * (8 * 8 + 4);
    double *qval = reinterpret_cast<double *>(smem_raw + off + (unsigned int)warp * QBYTES);      // [8][QCAP]
    int *qcell = reinterpret_cast<int *>(qval + 8 * AR_QCAP);                                        // [QCAP]
    if (AXK != 2) {
        for (int i = threadIdx.x; i < nx; i += blockDim.x) tabx[i] = p.g.ax[0].tab[i];
        for (int i = threadIdx.x; i < ny; i += blockDim.x) taby[i] = p.g.ax[1].tab[i];
        for (int i = threadIdx.x; i < nz; i += blockDim.x) tabz[i] = p.g.ax[2].tab[i];
    }
    if (BULK && lane == 0)
        for (int s = 0; s < AR_STAGES; ++s) mbar_init(&bars[s], 1);
    if (BULK) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint64_t pol_stream = policy_evict_first();
    const AxisR ax = axis_regs(p.g.ax[0]), ay = axis_regs(p.g.ax[1]), az = axis_regs(p.g.ax[2]);
    const int Ns = p.Ns, Nt = p.Nt, Nd = p.Nd;
    const bool n_odd = Ns & 1;
    const int chunks = (Ns + AR_C - 1) / AR_C;
    const long long n_tasks = (long long)p.Na * Nd * chunks;
    const long long ray_doubles = 4LL * Ns;
    const int sy = nz, sx = ny * nz;
    const unsigned full = 0xffffffffu;
    unsigned int n_oob = 0;
    unsigned int phases = 0;
    int qhead = 0, qtail = 0;      // queue indices (monotone; slot = index % QCAP)

    // drain 32 tuples (or all that are left, at the end): one tuple per lane, 8 reductions with the lanes that hold one
    auto drain = [&](int count) {
        __syncwarp();
        if (lane < count) {
            const int slot = (qhead + lane) % AR_QCAP;
            double *c = p.acc + qcell[slot];
            atomicAdd(c, qval[0 * AR_QCAP + slot]);           atomicAdd(c + 1, qval[1 * AR_QCAP + slot]);
            atomicAdd(c + sy, qval[2 * AR_QCAP + slot]);      atomicAdd(c + sy + 1, qval[3 * AR_QCAP + slot]);
            atomicAdd(c + sx, qval[4 * AR_QCAP + slot]);      atomicAdd(c + sx + 1, qval[5 * AR_QCAP + slot]);
            atomicAdd(c + sx + sy, qval[6 * AR_QCAP + slot]); atomicAdd(c + sx + sy + 1, qval[7 * AR_QCAP + slot]);
        }
        qhead += count;
        __syncwarp();
    };

    for (long long task = (long long)blockIdx.x * nwarp + warp; task < n_tasks; task += (long long)gridDim.x * nwarp) {
        const int c = (int)(task % chunks);
        const long long ad = task / chunks;
        const int d = (int)(ad % Nd), a = (int)(ad / Nd);
        const int c0 = c * AR_C;
        const int n_c = min(AR_C, Ns - c0);
        const bool valid = lane < n_c;
        const int i = c0 + lane;
        const long long ray0 = ((long long)a * Nt) * Nd + d;       // ray index at t = 0; + t * Nd
        // producer: stage of time step t lives in ring slot t % STAGES
        auto produce = [&](int t) {
            if (t < Nt)
                fill_stage<AR_C, true>(reinterpret_cast<double *>(ring + (t % AR_STAGES) * StageLayout<AR_C>::BYTES),
                                       &bars[t % AR_STAGES], p.rays + (ray0 + (long long)t * Nd) * ray_doubles, Ns, c0,
                                       lane, pol_stream);
        };
        if (BULK)
            for (int t = 0; t < AR_STAGES - 1; ++t) produce(t);
        const SimpsonCoef sk = simpson_coef(i, Ns, n_odd);      // i is fixed for the whole walk along the time axis
        int v_cur = -1;
        double a000 = 0, a001 = 0, a010 = 0, a011 = 0, a100 = 0, a101 = 0, a110 = 0, a111 = 0;
        for (int t = 0; t < Nt; ++t) {
            const int us = t % AR_STAGES;
            double *stage = reinterpret_cast<double *>(ring + us * StageLayout<AR_C>::BYTES);
            const long long ray = ray0 + (long long)t * Nd;
            if (BULK) {
                produce(t + AR_STAGES - 1);
                mbar_wait(&bars[us], (phases >> us) & 1u);
                phases ^= 1u << us;
            } else {
                fill_stage<AR_C, false>(stage, nullptr, p.rays + ray * ray_doubles, Ns, c0, lane, pol_stream);
                __syncwarp();
            }
            const double coef = __ldg(p.coef + ray);
            int v = -1;
            double h00 = 0, h01 = 0, h10 = 0, h11 = 0, l00 = 0, l01 = 0, l10 = 0, l11 = 0;
            if (valid) {
                const double *ss_ = stage + StageLayout<AR_C>::S_OFF + 2;
                int ix, iy, iz;
                double tx, ty, tz;
                const bool oob = locate3<AXK>(tabx, taby, tabz, ax, ay, az, stage[lane], stage[AR_C + lane],
                                              stage[2 * AR_C + lane], ix, iy, iz, tx, ty, tz);
                n_oob += oob;
                const double w = simpson_weight_c(sk, ss_[lane - 2], ss_[lane - 1], ss_[lane], ss_[lane + 1], ss_[lane + 2]);
                v = (ix * ny + iy) * nz + iz;
                const double aw = coef * w;
                const double ax1 = aw * tx, ax0 = aw - ax1;
                const double a01 = ax0 * ty, a00 = ax0 - a01;
                const double a11 = ax1 * ty, a10 = ax1 - a11;
                h00 = a00 * tz; h01 = a01 * tz; h10 = a10 * tz; h11 = a11 * tz;
                l00 = a00 - h00; l01 = a01 - h01; l10 = a10 - h10; l11 = a11 - h11;
            }
            // cell changed: the finished tuple goes to the queue
            const bool flush = (v != v_cur) && (v_cur >= 0);
            const unsigned m = __ballot_sync(full, flush);
            if (m) {
                if (flush) {
                    const int slot = (qtail + __popc(m & ((1u << lane) - 1u))) % AR_QCAP;
                    qcell[slot] = v_cur;
                    qval[0 * AR_QCAP + slot] = a000; qval[1 * AR_QCAP + slot] = a001;
                    qval[2 * AR_QCAP + slot] = a010; qval[3 * AR_QCAP + slot] = a011;
                    qval[4 * AR_QCAP + slot] = a100; qval[5 * AR_QCAP + slot] = a101;
                    qval[6 * AR_QCAP + slot] = a110; qval[7 * AR_QCAP + slot] = a111;
                }
                qtail += __popc(m);
                if (qtail - qhead >= 32) drain(32);
            }
            // same cell: add; new cell: start over -- as 8 fmas with keep = 1 or 0 (fma(1, a, l) == a + l exactly)
            const double keep = (v == v_cur) ? 1.0 : 0.0;
            v_cur = v;
            a000 = fma(keep, a000, l00); a001 = fma(keep, a001, h00); a010 = fma(keep, a010, l01);
            a011 = fma(keep, a011, h01); a100 = fma(keep, a100, l10); a101 = fma(keep, a101, h10);
            a110 = fma(keep, a110, l11); a111 = fma(keep, a111, h11);
            __syncwarp();     // the stage is free for the producer again
        }
        // end of the task: what the lanes still hold
        {
            const bool flush = v_cur >= 0;
            const unsigned m = __ballot_sync(full, flush);
            if (flush) {
                const int slot = (qtail + __popc(m & ((1u << lane) - 1u))) % AR_QCAP;
                qcell[slot] = v_cur;
                qval[0 * AR_QCAP + slot] = a000; qval[1 * AR_QCAP + slot] = a001;
                qval[2 * AR_QCAP + slot] = a010; qval[3 * AR_QCAP + slot] = a011;
                qval[4 * AR_QCAP + slot] = a100; qval[5 * AR_QCAP + slot] = a101;
                qval[6 * AR_QCAP + slot] = a110; qval[7 * AR_QCAP + slot] = a111;
            }
            qtail += __popc(m);
            if (qtail - qhead >= 32) drain(32);
        }
    }
    while (qtail > qhead) drain(min(32, qtail - qhead));
    if (n_oob) atomicAdd(p.oob_count, (unsigned long long)n_oob);
}

template <int AXK, bool BULK, int MAXT>
static int launch_adjoint_runs_m(const AdjRunsParams &p, int warps, size_t smem, int ctas, cudaStream_t st) {
    auto kern = adjoint_runs_kernel<AXK, BULK, MAXT>;
    CU_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ctas, warps * 32, smem, st>>>(p);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}
template <int AXK, bool BULK>
static int launch_adjoint_runs_t(const AdjRunsParams &p, int warps, size_t smem, int ctas, cudaStream_t st) {
    // beyond 16 warps the register budget is 80 per thread
    if (warps > 16) return launch_adjoint_runs_m<AXK, BULK, 768>(p, warps, smem, ctas, st);
    return launch_adjoint_runs_m<AXK, BULK, 512>(p, warps, smem, ctas, st);
}

// returns IONO_OK, an error, or -1 for "not applicable, use the plain scatter kernel"
static int launch_adjoint_runs(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns, const double *coef,
                               double *acc, unsigned long long *oob_count, cudaStream_t st) {
    if (const char *e = getenv("IONO_ADJOINT_RUNS")) {
        if (atoi(e) == 0) return -1;
    }
    if (Nt < 2) return -1;       // nothing to aggregate along the time axis
    const int axk = grid->exact ? 2 : (grid->uniform ? 1 : 0);
    const size_t table_bytes =
        (axk == 2) ? 0 : (((size_t)(grid->nx + grid->ny + grid->nz) * sizeof(double2)) + 127) / 128 * 128;
    int warps = 20;      // 1.69 ms; 16: 1.71, 24: 1.78 at the LOFAR case (profiles/r02_kernel_bench.json)
    if (const char *e = getenv("IONO_ADJOINT_RUNS_WARPS")) { int v = atoi(e); if (v >= 1 && v <= 24) warps = v; }
    auto smem_for = [&](int w) {
        return table_bytes + (((size_t)w * AR_STAGES * sizeof(uint64_t)) + 127) / 128 * 128 +
               (size_t)w * AR_STAGES * StageLayout<AR_C>::BYTES + (size_t)w * AR_QCAP * (8 * 8 + 4);
    };
    while (warps > 2 && smem_for(warps) > 227 * 1024) warps -= 2;
    const size_t smem = smem_for(warps);
    if (smem > 227 * 1024) return -1;
    AdjRunsParams p;
    memset(&p, 0, sizeof(p));
    p.g = grid->dev; p.rays = rays; p.coef = coef; p.acc = acc; p.oob_count = oob_count;
    p.Na = Na; p.Nt = Nt; p.Nd = Nd; p.Ns = Ns;
    const bool bulk = (Ns % 2 == 0) && (((uintptr_t)rays & 15) == 0) && !getenv("IONO_SWEEP_NO_BULK");
    const long long n_tasks = (long long)Na * Nd * ((Ns + AR_C - 1) / AR_C);
    long long want = (n_tasks + warps - 1) / warps;
    int ctas = sm_count();
    if (ctas > want) ctas = (int)want;
    if (axk == 2) return bulk ? launch_adjoint_runs_t<2, true>(p, warps, smem, ctas, st) : launch_adjoint_runs_t<2, false>(p, warps, smem, ctas, st);
    if (axk == 1) return bulk ? launch_adjoint_runs_t<1, true>(p, warps, smem, ctas, st) : launch_adjoint_runs_t<1, false>(p, warps, smem, ctas, st);
    return bulk ? launch_adjoint_runs_t<0, true>(p, warps, smem, ctas, st) : launch_adjoint_runs_t<0, false>(p, warps, smem, ctas, st);
}
