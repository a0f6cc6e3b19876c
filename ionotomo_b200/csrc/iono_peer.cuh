// Cross-GPU sum of the sharded adjoint over NVLink peer memory, fused with the chain-rule scaling and the
// expansion to the grid.  Included by iono_kernels.cu.
//
// One process per GPU holds a shard of the rays (antennas x its block of directions or times) and the whole
// grid.  The reference sums the per-shard gradients with da.sum over dask workers
// (inversion/gradient.py:52-54, gradient_and_adjoint.py:125-127); here every rank's back-projector writes its
// partial sums into a COMPACT accumulator that numbers only the voxels any rank touches (a fifth of the grid at
// the LOFAR case: 13 MB instead of 67 MB), and one kernel per rank then
//   1. signals "my accumulator is complete" to every peer and waits for theirs      (flags in peer memory),
//   2. reduce-scatter by PULL: sums its 1/N slice over the ranks' accumulators in rank order -- peer loads
//      over NVLink, a fixed order, so all ranks get the same bits run after run --
//   3. all-gather by PUSH: stores the summed slice into every rank's result vector (peer stores) and signals,
//   4. expands: grad[voxel[k]] = K exp(m[voxel[k]]) * sum[k]  (voxels no ray touches keep their zero) -- its own
//      slice straight from the registers that summed it, every other slice as soon as that slice's owner has
//      signalled, so the expansion overlaps the arrival of the later slices -- and hands out the summed misfit
//      that travels as the last element of the vector.
// Nothing but this kernel touches the link; no NCCL call sits on the step's critical path.
//
// Epoch counters instead of flag resets: call e writes e, waits for >= e; the epoch itself lives in device memory
// and is advanced by the kernel, so the launch is identical every time (CUDA-graph replay).  Every rank must
// make the same sequence of calls.  The buffers are cudaMalloc'ed by
// iono_peer_alloc and mapped into the peers with CUDA IPC handles (exchanged by the Python layer through
// torch.distributed).
#pragma once

constexpr int IONO_MAX_PEERS = 16;

struct PeerTable {
    double *acc[IONO_MAX_PEERS];                 // compact accumulators (L doubles), [me] is local
    double *res[IONO_MAX_PEERS];                 // result vectors (L doubles), [me] is local
    unsigned long long *flags[IONO_MAX_PEERS];   // per rank: [2][IONO_MAX_PEERS] epochs, arrival counter, call counter,
                                                 // 5 phase time stamps of the last call
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Peer data: relaxed system-scope loads (never served from a stale L1 line).  `asm volatile` keeps them below the
// flag wait (which ends in __syncthreads()) and in program order; there is no memory clobber and the callers
// issue all loads of a step before the first use, so 2N requests per thread are in flight over the link.
__device__ __forceinline__ double2 ld_peer_v2(const double *p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_peer(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// all CTAs: wait until every rank has published `epoch` in row `phase` of the local flag block
__device__ __forceinline__ void wait_all_ranks(const unsigned long long *my_flags, int phase, int N,
                                               unsigned long long epoch) {
    if ((int)threadIdx.x < N) {
        const unsigned long long *f = my_flags + phase * IONO_MAX_PEERS + threadIdx.x;
        while (ld_acquire_sys(f) < epoch) __nanosleep(64);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024, 1) peer_reduce_expand_kernel(PeerTable T, int N, int me, long long L,
                                                                  const int *__restrict__ voxel, long long n_union,
                                                                  const double *__restrict__ m, double k,
                                                                  double *__restrict__ grad,
                                                                  double *__restrict__ misfit_out) {
    unsigned long long *my_flags = T.flags[me];
    unsigned int *arrivals = reinterpret_cast<unsigned int *>(my_flags + 2 * IONO_MAX_PEERS);
    unsigned long long *calls = my_flags + 2 * IONO_MAX_PEERS + 1;
    __shared__ bool last;
    // every CTA reads the call counter before its arrival below; CTA 0 advances it after the second wait,
    // which cannot complete before all CTAs of this grid have arrived
    const unsigned long long epoch = ld_acquire_sys(calls) + 1ull;
    // phase time stamps of CTA 0 (ns, %globaltimer) for the last call: start, accumulators complete everywhere,
    // my slice reduced, pushed and expanded, first foreign slice arrived, all slices expanded (read by the bench)
    unsigned long long *stamps = my_flags + 2 * IONO_MAX_PEERS + 2;
    auto stamp = [&](int i) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            stamps[i] = t;
        }
    };
    stamp(0);
    // 1. my accumulator was finished by the previous kernel of this stream: tell everybody (one CTA does)
    if (blockIdx.x == 0 && (int)threadIdx.x < N) {
        __threadfence_system();
        st_release_sys(T.flags[threadIdx.x] + 0 * IONO_MAX_PEERS + me, epoch);
    }
    wait_all_ranks(my_flags, 0, N, epoch);
    stamp(1);
    // 2. + 3. my slice (pairs of doubles; L is padded to an even length by the caller).  The thread that summed a
    // pair also expands it -- the sum is in its registers -- so the own slice never waits for anything
    const long long pairs = L / 2;
    const long long p0 = pairs * me / N, p1 = pairs * (me + 1) / N;
    const long long stride = (long long)gridDim.x * blockDim.x;
    auto expand_pair = [&](long long p, double2 s) {
        const long long i = 2 * p;
        if (i + 1 < n_union) {
            const int2 v = __ldg(reinterpret_cast<const int2 *>(voxel + i));
            const double e0 = exp(__ldg(m + v.x)), e1 = exp(__ldg(m + v.y));
            grad[v.x] = k * e0 * s.x;
            grad[v.y] = k * e1 * s.y;
        } else if (i < n_union) {
            const int v = __ldg(voxel + i);
            grad[v] = k * exp(__ldg(m + v)) * s.x;
        }
    };
    for (long long p = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; p < p1; p += stride) {
        double2 s = make_double2(0.0, 0.0);
        for (int r0 = 0; r0 < N; r0 += 8) {          // 8 peer loads in flight per thread, summed in rank order
            double2 v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r0 + r < N) v[r] = ld_peer_v2(T.acc[r0 + r] + 2 * p);
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r0 + r < N) { s.x += v[r].x; s.y += v[r].y; }
        }
        for (int r = 0; r < N; ++r) *reinterpret_cast<double2 *>(T.res[r] + 2 * p) = s;
        expand_pair(p, s);
    }
    // every CTA's peer stores must be out before the slice is announced: the CTA barrier orders the threads'
    // stores before thread 0's system-scope fence (cumulative), the last CTA to arrive signals
    __syncthreads();
    stamp(2);
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = (atomicAdd(arrivals, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        if ((int)threadIdx.x < N) {
            __threadfence_system();
            st_release_sys(T.flags[threadIdx.x] + 1 * IONO_MAX_PEERS + me, epoch);
        }
        if (threadIdx.x == 0) *arrivals = 0u;      // nobody of this launch reads it again
    }
    // 4. the other slices, each as soon as ITS owner has announced it (no all-ranks wait: the expansion of the
    // early slices hides the arrival of the late ones): grad[voxel[i]] = K exp(m[voxel[i]]) * sum[i]
    const double *res = T.res[me];
    for (int dr = 1; dr < N; ++dr) {
        const int r = (me + dr) % N;
        if (threadIdx.x == 0) {
            const unsigned long long *f = my_flags + 1 * IONO_MAX_PEERS + r;
            while (ld_acquire_sys(f) < epoch) __nanosleep(32);
        }
        __syncthreads();
        if (dr == 1) stamp(3);
        const long long q0 = pairs * r / N, q1 = pairs * (r + 1) / N;
        for (long long p = q0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; p < q1; p += stride)
            expand_pair(p, ld_peer_v2(res + 2 * p));
    }
    // the call counter advances once every CTA of this grid has read it, i.e. after this rank's own announcement
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long *f = my_flags + 1 * IONO_MAX_PEERS + me;
        while (ld_acquire_sys(f) < epoch) __nanosleep(32);
        *calls = epoch;
        // the last element of the vector is the summed misfit (reduced by the owner of the last slice; every
        // slice has been waited for above)
        if (misfit_out) misfit_out[0] = ld_peer(res + n_union);
    }
    stamp(4);
}

// ---- host side -----------------------------------------------------------------------------------------
extern "C" int iono_peer_alloc(int64_t bytes, void **ptr_out, void *ipc_handle_out64) {
    if (bytes <= 0 || !ptr_out || !ipc_handle_out64) return fail(IONO_EBADARG, "iono_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    CU_CHECK(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(IONO_ECUDA, "iono_peer_alloc: %s", cudaGetErrorString(e));
    }
    memcpy(ipc_handle_out64, &h, 64);
    *ptr_out = p;
    return IONO_OK;
}

extern "C" int iono_peer_open(const void *ipc_handle64, void **ptr_out) {
    if (!ipc_handle64 || !ptr_out) return fail(IONO_EBADARG, "iono_peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, 64);
    void *p = nullptr;
    CU_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = p;
    return IONO_OK;
}

extern "C" int iono_peer_close(void *ptr) {
    if (ptr) CU_CHECK(cudaIpcCloseMemHandle(ptr));
    return IONO_OK;
}

extern "C" int iono_peer_free(void *ptr) {
    if (ptr) CU_CHECK(cudaFree(ptr));
    return IONO_OK;
}

extern "C" int64_t iono_peer_flag_bytes(void) { return (2 * IONO_MAX_PEERS + 2 + 6) * 8; }

// acc, res, flags: arrays of N device pointers (entry `me` local, the others opened with iono_peer_open);
// L: length of the compact vectors (even; element n_union carries the misfit).
extern "C" int iono_peer_reduce_expand_f64(void *const *acc, void *const *res, void *const *flags, int N, int me,
                                           int64_t L, const int *union_voxels, int64_t n_union,
                                           const double *m, double k, double *grad, double *misfit_out,
                                           void *stream) {
    if (!acc || !res || !flags || N < 1 || N > IONO_MAX_PEERS || me < 0 || me >= N || L < 2 || (L & 1) ||
        n_union < 0 || n_union >= L || (n_union > 0 && (!union_voxels || !m || !grad)))
        return fail(IONO_EBADARG, "iono_peer_reduce_expand_f64: bad argument");
    PeerTable T;
    memset(&T, 0, sizeof(T));
    for (int r = 0; r < N; ++r) {
        if (!acc[r] || !res[r] || !flags[r]) return fail(IONO_EBADARG, "iono_peer_reduce_expand_f64: NULL peer pointer");
        T.acc[r] = (double *)acc[r];
        T.res[r] = (double *)res[r];
        T.flags[r] = (unsigned long long *)flags[r];
    }
    // all CTAs spin on flags: the grid must be co-resident -- one CTA per SM is.  1024 threads per CTA: the loops are
    // chains of dependent long-latency accesses (peer loads; voxel index -> model gather -> scattered store), so the
    // kernel's time is latency / (threads in flight)
    int ctas = sm_count(), threads = 1024;
    if (const char *e = getenv("IONO_PEER_CTAS")) { int v = atoi(e); if (v >= 1 && v <= ctas) ctas = v; }
    if (const char *e = getenv("IONO_PEER_THREADS")) { int v = atoi(e); if (v >= 32 && v <= 1024 && v % 32 == 0) threads = v; }
    peer_reduce_expand_kernel<<<ctas, threads, 0, (cudaStream_t)stream>>>(T, N, me, (long long)L, union_voxels, (long long)n_union, m, k, grad,
                                                                     misfit_out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}
