// Voxel-binned back-projector: the adjoint without atomics.
// Included by iono_kernels.cu (one translation unit).
//
// The scatter formulation of the adjoint (ray_sweep_kernel<ADJ>) is bound by the rate at
// which an SM can issue global reductions (measured: one warp-level RED per ~39 cycles,
// whatever the type, lane count or sector count -- tools/probes/atomics_probe.cu).  Inside
// an inversion the ray geometry is fixed while the model and the residuals change every
// iteration (reference driver loop: tests/test_inversion.py:30-39, bfgs_dask.py:207-340), so
// the linear operator  acc[v] = sum_ray coef[ray] * A[v,ray],
//     A[v,ray] = sum_s w_s(ray) * phi_v(x_s(ray))     (Simpson-avg weights x trilinear hats)
// can be assembled ONCE per geometry in voxel-major (CSC-like) form and then applied as a
// pure gather: one warp per voxel streams that voxel's (ray, weight) list, gathers coef[ray]
// from L2 and reduces with shuffles.  No atomics, bit-reproducible, and the result is the
// same matrix the forward kernel applies, so <Gx,y> = <x,G^T y> still holds to rounding.
//
// Assembly (all on the GPU): emit 8 (voxel<<32|ray, weight) pairs per sample, radix-sort by
// key (CUB), sum runs of equal key (a ray visits a voxel through 1-3 consecutive samples),
// split the keys into a ray-index array and voxel offsets.
#pragma once
#include <cub/cub.cuh>
#include <cuda/std/functional>

struct iono_backprojector {
    unsigned int *ray_idx;   // [nnz]
    double *weight;          // [nnz]
    long long *ptr;          // [n_rows+1] entry offsets of the non-empty rows
    unsigned int *row_voxel; // [n_rows] voxel index of each non-empty row
    long long n_rows;
    int *long_rows;          // rows whose entries span 2..8 segments (thread each in the combine step)
    int n_long;
    int *vlong_rows;         // rows spanning more than 8 segments (warp each)
    int n_vlong;
    double *partial;         // [2 * nseg] per-segment sums of the straddling rows
    int2 *items;             // [nseg] first and last row of every segment
    long long nnz;
    long long V;
    long long R;
    int Na, Nt, Nd;
    int seg;                 // entries per segment
    // 16 equal chunks of the segment range (for overlapping the cross-GPU sum with the apply):
    // chunk c completes rows [chunk_row[c], chunk_row[c+1]) = voxels [chunk_vox[c], chunk_vox[c+1])
    long long chunk_seg[17], chunk_row[17], chunk_vox[17];
    int chunk_short[17], chunk_vlong[17];   // index ranges in the (sorted) straddler lists
    double *coef_perm;       // [R] coefficients in the internal ray order (a, d, t)
    int device;
    // optional run-compressed ray indices (IONO_BP_RUNS=1 at create time, warp-private segments only):
    // per segment a 256-bit mask of run heads + the ray index of every head, see backproject_wruns_kernel
    unsigned char *runs;           // records, 16-byte units
    unsigned long long *run_ptr;   // [nseg+1] record offsets in 16-byte units
    long long run_bytes;
    int use_runs;
};

// Per-row factor applied when a finished row is stored: nothing, a grid array (scale[v]), or the chain-rule
// factor of the log-density model computed on the spot, k * exp(p[v]) (ne[v] = K_ne exp(m[v]) / 1e13) -- the
// operator touches a fifth of the voxels, so this replaces a full-grid ne pass by ~1.6 M exps.
// `index` (optional) replaces row_voxel as the destination index of a row: the compact accumulator of the
// sharded adjoint (rows of all ranks numbered consecutively, ionotomo_b200/inversion/session.py).
struct RowScale {
    const double *p;
    double k;
    int mode;   // 0: 1, 1: p[v], 2: k * exp(p[v])
};
__device__ __forceinline__ double row_scale(const RowScale &rs, unsigned int v) {
    if (rs.mode == 1) return __ldg(rs.p + v);
    if (rs.mode == 2) return rs.k * exp(__ldg(rs.p + v));
    return 1.0;
}

// Internal ray numbering of the back-projector: time fastest, (a*Nd + d)*Nt + t.  Rays of one
// (antenna, direction) at consecutive times are nearly identical and meet in the same voxels,
// so with this order a voxel's sorted entry list references runs of adjacent coefficients
// (4 per 32-byte sector) instead of one sector per entry.
__global__ void __launch_bounds__(256) permute_coef_kernel(const double *__restrict__ coef, int Na, int Nt, int Nd,
                                                            double *__restrict__ out) {
    __shared__ double tile[32][33];
    // per antenna: (Nt, Nd) -> (Nd, Nt), 32x32 tiles
    const int tiles_d = (Nd + 31) / 32, tiles_t = (Nt + 31) / 32;
    const int per_a = tiles_d * tiles_t;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int b = blockIdx.x; b < Na * per_a; b += gridDim.x) {
        const int a = b / per_a, r = b % per_a;
        const int t0 = (r / tiles_d) * 32, d0 = (r % tiles_d) * 32;
        const double *src = coef + (long long)a * Nt * Nd;
        double *dst = out + (long long)a * Nt * Nd;
        __syncthreads();
        for (int j = ty; j < 32; j += 8)
            if (t0 + j < Nt && d0 + tx < Nd) tile[j][tx] = src[(long long)(t0 + j) * Nd + d0 + tx];
        __syncthreads();
        for (int j = ty; j < 32; j += 8)
            if (d0 + j < Nd && t0 + tx < Nt) dst[(long long)(d0 + j) * Nt + t0 + tx] = tile[tx][j];
    }
}

template <int AXK>
__global__ void __launch_bounds__(256) emit_entries_kernel(Grid g, const double *__restrict__ rays, int R, int Nt,
                                                            int Nd, int Ns,
                                                            unsigned long long *__restrict__ keys,
                                                            double *__restrict__ vals,
                                                            unsigned long long *oob_count) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int ny = g.ax[1].n, nz = g.ax[2].n;
    const bool n_odd = Ns & 1;
    const AxisR ax = axis_regs(g.ax[0]), ay = axis_regs(g.ax[1]), az = axis_regs(g.ax[2]);
    unsigned int n_oob = 0;
    for (int ray = warp_global; ray < R; ray += n_warps) {
        const double *rp = rays + (long long)ray * 4 * Ns;
        const double *sp = rp + 3 * Ns;
        for (int i = lane; i < Ns; i += 32) {
            int ix, iy, iz;
            double tx, ty, tz;
            // the same cell and in-cell coordinates as the forward sweep derives (one operator, two directions)
            n_oob += locate3<AXK>(g.ax[0].tab, g.ax[1].tab, g.ax[2].tab, ax, ay, az, __ldg(rp + i), __ldg(rp + Ns + i),
                                  __ldg(rp + 2 * Ns + i), ix, iy, iz, tx, ty, tz);
            const double sm2 = (i >= 2) ? __ldg(sp + i - 2) : 0.0, sm1 = (i >= 1) ? __ldg(sp + i - 1) : 0.0;
            const double s0 = __ldg(sp + i);
            const double sp1 = (i + 1 < Ns) ? __ldg(sp + i + 1) : 0.0, sp2 = (i + 2 < Ns) ? __ldg(sp + i + 2) : 0.0;
            const double a = simpson_weight(i, Ns, n_odd, sm2, sm1, s0, sp1, sp2);
            const unsigned long long v = (unsigned long long)((ix * ny + iy) * nz + iz);
            const unsigned long long sy = nz, sx = (unsigned long long)ny * nz;
            const double ax1 = a * tx, ax0 = a - ax1;
            const double a01 = ax0 * ty, a00 = ax0 - a01;
            const double a11 = ax1 * ty, a10 = ax1 - a11;
            const long long e = ((long long)ray * Ns + i) * 8;
            const int ra = ray / (Nt * Nd), rem = ray - ra * (Nt * Nd);
            const int rt = rem / Nd, rd = rem - rt * Nd;
            const unsigned long long r = (unsigned long long)((ra * Nd + rd) * Nt + rt);
            double hi;
            hi = a00 * tz; keys[e + 0] = ((v) << 32) | r;              vals[e + 0] = a00 - hi;
                           keys[e + 1] = ((v + 1) << 32) | r;          vals[e + 1] = hi;
            hi = a01 * tz; keys[e + 2] = ((v + sy) << 32) | r;         vals[e + 2] = a01 - hi;
                           keys[e + 3] = ((v + sy + 1) << 32) | r;     vals[e + 3] = hi;
            hi = a10 * tz; keys[e + 4] = ((v + sx) << 32) | r;         vals[e + 4] = a10 - hi;
                           keys[e + 5] = ((v + sx + 1) << 32) | r;     vals[e + 5] = hi;
            hi = a11 * tz; keys[e + 6] = ((v + sx + sy) << 32) | r;    vals[e + 6] = a11 - hi;
                           keys[e + 7] = ((v + sx + sy + 1) << 32) | r; vals[e + 7] = hi;
        }
    }
    if (n_oob) atomicAdd(oob_count, (unsigned long long)n_oob);
}

__global__ void __launch_bounds__(256) split_keys_kernel(const unsigned long long *__restrict__ keys, long long n,
                                                          unsigned int *__restrict__ ray_idx) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride)
        ray_idx[e] = (unsigned int)(keys[e] & 0xffffffffull);
}

struct KeyVoxel {
    __host__ __device__ __forceinline__ unsigned int operator()(unsigned long long k) const {
        return (unsigned int)(k >> 32);
    }
};

// Apply = a balanced sweep over the ENTRY stream, not over voxels: row lengths range from 0 to
// ~1e6 (heavy voxels sit under the array core, where every ray of a station starts in the same
// cell), so the entry index space is cut into segments of BP_WSEG entries, one WARP each:
//   1. products weight[k]*coef[ray(k)] of the segment -> shared memory (8 gathers in flight per lane);
//   2. every row (voxel) that intersects the segment is summed from shared memory by the warp;
//      rows completely inside are written to out[], the (at most two) rows that continue into
//      a neighbouring segment go to partial[2*seg + slot];
//   3. a small kernel adds the partials of each straddling row in segment order.
// slot 0: the row covers the segment's first entry; slot 1: the row begins inside the segment.
// Everything is a fixed reduction tree: bit-reproducible.

// voxel (row) that contains entry k: largest v with ptr[v] <= k
__device__ __forceinline__ long long row_of_entry(const long long *__restrict__ ptr, long long V, long long k) {
    long long lo = 0, hi = V;   // invariant: ptr[lo] <= k < ptr[hi]
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(ptr + mid) <= k) lo = mid; else hi = mid;
    }
    return lo;
}

// Build time: first and last row of every segment.
__global__ void __launch_bounds__(256) segment_rows_kernel(const long long *__restrict__ ptr, long long V,
                                                            long long nnz, int BP_SEG, int2 *__restrict__ seg_rows) {
    const long long nseg = (nnz + BP_SEG - 1) / BP_SEG;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long seg = (long long)blockIdx.x * blockDim.x + threadIdx.x; seg < nseg; seg += stride) {
        const long long k0 = seg * BP_SEG, k1 = min(k0 + (long long)BP_SEG, nnz);
        seg_rows[seg] = make_int2((int)row_of_entry(ptr, V, k0), (int)row_of_entry(ptr, V, k1 - 1));
    }
}

// Build time: rows that span more than one segment.
__global__ void __launch_bounds__(256) find_straddling_rows_kernel(const long long *__restrict__ ptr, long long V,
                                                                    int BP_SEG, int *__restrict__ rows, int *count,
                                                                    int cap, int *__restrict__ vrows, int *vcount,
                                                                    int vcap) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += stride) {
        const long long b = ptr[v], e = ptr[v + 1];
        if (e <= b) continue;
        const long long span = (e - 1) / BP_SEG - b / BP_SEG;
        if (span >= 8) {
            const int k = atomicAdd(vcount, 1);
            if (k < vcap) vrows[k] = (int)v;
        } else if (span >= 1) {
            const int k = atomicAdd(count, 1);
            if (k < cap) rows[k] = (int)v;
        }
    }
}

// Warp-private variant: every WARP owns segments of BP_WSEG entries (static stride over all warps of
// the grid), with its own 2-deep TMA ring and mbarriers -- no __syncthreads in the loop, so a warp
// that waits for its gathers or sums a long row never stalls the other warps of the CTA.
// Rows of a segment are summed one after the other by the whole warp (lanes stride, shuffle
// reduction); lane q remembers the total of row q and the lanes store their rows together.
constexpr int BP_WSEG = 256;

__global__ void __launch_bounds__(256) backproject_wsegments_kernel(const int2 *__restrict__ seg_rows,
                                                                     const long long *__restrict__ ptr,
                                                                     const unsigned int *__restrict__ row_voxel,
                                                                     const unsigned int *__restrict__ ray_idx,
                                                                     const double *__restrict__ weight,
                                                                     const double *__restrict__ coef,
                                                                     const RowScale scale, long long nnz,
                                                                     long long seg_begin, long long seg_end,
                                                                     double *__restrict__ out,
                                                                     double *__restrict__ partial) {
    extern __shared__ __align__(128) unsigned char bp_smem[];
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), nwarp = blockDim.x >> 5;   // (warp-uniform for the compiler)
    constexpr int STAGE = BP_WSEG * 12;                       // weights then ray indices
    unsigned char *mine = bp_smem + (size_t)warp * (2 * STAGE);
    uint64_t *bar = reinterpret_cast<uint64_t *>(bp_smem + (size_t)nwarp * 2 * STAGE) + warp * 2;
    const long long nseg = seg_end;
    const long long gw = seg_begin + (long long)blockIdx.x * nwarp + warp, gstride = (long long)gridDim.x * nwarp;
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const uint64_t pol = policy_evict_first();
    auto issue = [&](long long seg, int buf) {
        if (elect_one()) {
            mbar_expect_tx(&bar[buf], STAGE);
            bulk_g2s(mine + buf * STAGE, weight + seg * BP_WSEG, BP_WSEG * 8, &bar[buf], pol);
            bulk_g2s(mine + buf * STAGE + BP_WSEG * 8, ray_idx + seg * BP_WSEG, BP_WSEG * 4, &bar[buf], pol);
        }
    };
    if (gw < nseg) issue(gw, 0);
    unsigned int phase = 0;
    int buf = 0;
    constexpr int PER = BP_WSEG / 32;
    for (long long seg = gw; seg < nseg; seg += gstride, buf ^= 1) {
        const long long k0 = seg * BP_WSEG, k1 = min(k0 + (long long)BP_WSEG, nnz);
        const int2 rr = __ldg(seg_rows + seg);
        if (seg + gstride < nseg) issue(seg + gstride, buf ^ 1);
        const int n_rows = rr.y - rr.x + 1;
        // row table of the first 31 rows: lane q holds ptr[rr.x + q]; lanes < n_rows also voxel and scale
        long long myptr = 0;
        unsigned int myvox = 0;
        double myscale = 1.0;
        if (lane <= min(n_rows, 31)) myptr = __ldg(ptr + rr.x + lane);
        if (lane < min(n_rows, 31)) {
            myvox = __ldg(row_voxel + rr.x + lane);
            myscale = row_scale(scale, myvox);
        }
        mbar_wait(&bar[buf], (phase >> buf) & 1u);
        phase ^= 1u << buf;
        double *prod = reinterpret_cast<double *>(mine + buf * STAGE);
        const unsigned int *rs = reinterpret_cast<const unsigned int *>(mine + buf * STAGE + BP_WSEG * 8);
        double c[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) c[u] = __ldg(coef + rs[lane + u * 32]);
        if (n_rows == 1) {
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < PER; ++u) s = fma(prod[lane + u * 32], c[u], s);   // padding has weight 0
            s = warp_sum(s);
            const long long b = __shfl_sync(0xffffffffu, myptr, 0), e = __shfl_sync(0xffffffffu, myptr, 1);
            if (lane == 0) {
                if (b >= k0 && e <= k1) out[myvox] = s * myscale;
                else partial[2 * seg + (b > k0 ? 1 : 0)] = s;
            }
        } else {
#pragma unroll
            for (int u = 0; u < PER; ++u) prod[lane + u * 32] *= c[u];
            __syncwarp();
            int r0 = 0;   // first row of the current batch (relative to rr.x)
            while (true) {
                const int nb = min(n_rows - r0, 31);
                double mysum = 0.0;
                for (int q0 = 0; q0 < nb; q0 += 4) {   // as in backproject_wruns_kernel: four rows at a time
                    const int q = q0 + (lane >> 3);
                    const long long b = __shfl_sync(0xffffffffu, myptr, min(q, 31)),
                                    e = __shfl_sync(0xffffffffu, myptr, min(q + 1, 31));
                    const int lo = (int)(max(b, k0) - k0), hi = (q < nb) ? (int)(min(e, k1) - k0) : lo;
                    double s = 0.0;
                    for (int j = lo + (lane & 7); j < hi; j += 8) s += prod[j];
                    s += __shfl_xor_sync(0xffffffffu, s, 4);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    const double t = __shfl_sync(0xffffffffu, s, ((lane - q0) & 3) << 3);
                    if (lane >= q0 && lane < q0 + 4) mysum = t;
                }
                const long long mye = __shfl_down_sync(0xffffffffu, myptr, 1);
                if (lane < nb) {
                    if (myptr >= k0 && mye <= k1) out[myvox] = mysum * myscale;
                    else partial[2 * seg + (myptr > k0 ? 1 : 0)] = mysum;
                }
                r0 += nb;
                if (r0 >= n_rows) break;
                // next batch of rows (segments made of many tiny rows)
                myptr = 0; myvox = 0; myscale = 1.0;
                if (lane <= min(n_rows - r0, 31)) myptr = __ldg(ptr + rr.x + r0 + lane);
                if (lane < min(n_rows - r0, 31)) {
                    myvox = __ldg(row_voxel + rr.x + r0 + lane);
                    myscale = row_scale(scale, myvox);
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Run-compressed ray indices (the default; IONO_BP_RUNS=0 at create time keeps the plain 4-byte indices).
// In the internal ray order (time fastest) a voxel's sorted entry list consists of RUNS of consecutive
// ray numbers -- the rays of one (antenna, direction) at successive times cross the same voxels;
// measured on the LOFAR-like geometry: 8.6 entries per run -- so 4 bytes of index per entry are mostly
// redundant.  Per segment of BP_WSEG entries the index array is replaced by a record
//     [BP_WSEG x u8 run number, lane-major: byte lane*8 + u belongs to entry 32u + lane]
//     [n_runs x u32 base = ray index of the run's head - position of the head], padded to 16 bytes
// and the ray of entry j is  base[run(j)] + j : one shared-memory byte extract, one table read and one add
// per entry (a first version with a 256-bit head mask decoded by popc/clz was 0.6 B smaller per entry
// and instruction-bound: 20 instructions per 32 entries for the decode alone).  ~9.5 B per entry
// instead of 12.
// ---------------------------------------------------------------------------------------------
constexpr int BP_RUNREC_MAX = BP_WSEG + BP_WSEG * 4;   // run numbers + a base for every entry

// head flags of segment `seg` as ballot words: lane b, word u <-> entry 32u+b
__device__ __forceinline__ unsigned int run_head_word(const unsigned int *__restrict__ ray_idx, long long k0, int u,
                                                      int lane) {
    const int j = 32 * u + lane;
    const bool head = (j == 0) || (ray_idx[k0 + j] != ray_idx[k0 + j - 1] + 1u);
    return __ballot_sync(0xffffffffu, head);
}

// Build time, pass 1: record size of every segment in 16-byte units (warp per segment).
__global__ void __launch_bounds__(256) run_record_units_kernel(const unsigned int *__restrict__ ray_idx, long long nseg,
                                                                unsigned long long *__restrict__ units) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long seg = warp_global; seg < nseg; seg += n_warps) {
        int nh = 0;
        for (int u = 0; u < BP_WSEG / 32; ++u) nh += __popc(run_head_word(ray_idx, seg * BP_WSEG, u, lane));
        if (lane == 0) units[seg] = (unsigned long long)((BP_WSEG + 4 * nh + 15) / 16);
    }
    if (warp_global == 0 && lane == 0) units[nseg] = 0;
}

// Build time, pass 2: write the records.
__global__ void __launch_bounds__(256) run_record_fill_kernel(const unsigned int *__restrict__ ray_idx, long long nseg,
                                                               const unsigned long long *__restrict__ run_ptr,
                                                               unsigned char *__restrict__ runs) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long seg = warp_global; seg < nseg; seg += n_warps) {
        unsigned char *rec = runs + run_ptr[seg] * 16ull;
        unsigned int *bases = reinterpret_cast<unsigned int *>(rec + BP_WSEG);
        const long long k0 = seg * BP_WSEG;
        int cum = 0;
        for (int u = 0; u < BP_WSEG / 32; ++u) {
            const unsigned int m = run_head_word(ray_idx, k0, u, lane);
            const int run = cum + __popc(m & (0xffffffffu >> (31 - lane))) - 1;   // heads at positions <= mine
            rec[lane * 8 + u] = (unsigned char)run;
            if ((m >> lane) & 1u) bases[run] = ray_idx[k0 + 32 * u + lane] - (unsigned int)(32 * u + lane);
            cum += __popc(m);
        }
    }
}

// backproject_wsegments_kernel with the ray indices reconstructed from the run records, the row tables
// of the NEXT segment and the record offsets of the one after prefetched into registers (the dependent
// seg_rows -> ptr / row_voxel -> scale loads otherwise sit exposed at the head of every segment), and
// 32-bit offsets relative to the segment in the row loop.
__global__ void __launch_bounds__(256) backproject_wruns_kernel(const int2 *__restrict__ seg_rows,
                                                                 const long long *__restrict__ ptr,
                                                                 const unsigned int *__restrict__ row_voxel,
                                                                 const unsigned char *__restrict__ runs,
                                                                 const unsigned long long *__restrict__ run_ptr,
                                                                 const double *__restrict__ weight,
                                                                 const double *__restrict__ coef,
                                                                 const RowScale scale, long long nnz,
                                                                 long long seg_begin, long long seg_end,
                                                                 double *__restrict__ out,
                                                                 double *__restrict__ partial) {
    extern __shared__ __align__(128) unsigned char bp_smem[];
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), nwarp = blockDim.x >> 5;   // (warp-uniform for the compiler)
    constexpr int STAGE = BP_WSEG * 8 + BP_RUNREC_MAX;         // weights then the run record
    unsigned char *mine = bp_smem + (size_t)warp * (2 * STAGE);
    uint64_t *bar = reinterpret_cast<uint64_t *>(bp_smem + (size_t)nwarp * 2 * STAGE) + warp * 2;
    const long long nseg = seg_end;
    const long long gw = seg_begin + (long long)blockIdx.x * nwarp + warp, gstride = (long long)gridDim.x * nwarp;
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (gw >= nseg) return;
    const uint64_t pol = policy_evict_first();
    auto issue = [&](long long seg, int buf, unsigned long long u0, unsigned long long u1) {
        if (elect_one()) {
            const unsigned int rec_bytes = (unsigned int)(u1 - u0) * 16u;
            mbar_expect_tx(&bar[buf], BP_WSEG * 8 + rec_bytes);
            bulk_g2s(mine + buf * STAGE, weight + seg * BP_WSEG, BP_WSEG * 8, &bar[buf], pol);
            bulk_g2s(mine + buf * STAGE + BP_WSEG * 8, runs + u0 * 16ull, rec_bytes, &bar[buf], pol);
        }
    };
    // row table of up to 31 rows starting at row `first`: lane q holds ptr[first + q] (q <= n) and the voxel (q < n)
    auto load_rows = [&](int first, int n, long long &p, unsigned int &vox) {
        p = 0; vox = 0;
        if (lane <= min(n, 31)) p = __ldg(ptr + first + lane);
        if (lane < min(n, 31)) vox = __ldg(row_voxel + first + lane);
    };
    // software pipeline: rr/rr1 = row ranges of this and the next segment, (v0,v1) = record offsets of the next
    int2 rr = __ldg(seg_rows + gw);
    int2 rr1 = make_int2(0, -1);
    unsigned long long v0 = 0, v1 = 0;
    issue(gw, 0, __ldg(run_ptr + gw), __ldg(run_ptr + gw + 1));
    if (gw + gstride < nseg) {
        rr1 = __ldg(seg_rows + gw + gstride);
        v0 = __ldg(run_ptr + gw + gstride);
        v1 = __ldg(run_ptr + gw + gstride + 1);
    }
    long long myptr;
    unsigned int myvox;
    load_rows(rr.x, rr.y - rr.x + 1, myptr, myvox);
    unsigned int phase = 0;
    int buf = 0;
    constexpr int PER = BP_WSEG / 32;
    for (long long seg = gw; seg < nseg; seg += gstride, buf ^= 1) {
        const long long k0 = seg * BP_WSEG;
        const int kend = (int)(min(k0 + (long long)BP_WSEG, nnz) - k0);
        const long long nxt = seg + gstride, nxt2 = nxt + gstride;
        if (nxt < nseg) issue(nxt, buf ^ 1, v0, v1);
        // prefetches: rows of the next segment, row range and record offsets of the one after
        long long nptr = 0;
        unsigned int nvox = 0;
        if (nxt < nseg) load_rows(rr1.x, rr1.y - rr1.x + 1, nptr, nvox);
        int2 rr2 = make_int2(0, -1);
        unsigned long long w0 = 0, w1 = 0;
        if (nxt2 < nseg) {
            rr2 = __ldg(seg_rows + nxt2);
            w0 = __ldg(run_ptr + nxt2);
            w1 = __ldg(run_ptr + nxt2 + 1);
        }
        const int n_rows = rr.y - rr.x + 1;
        double myscale = 1.0;
        if (lane < min(n_rows, 31)) myscale = row_scale(scale, myvox);
        // start of row q relative to the segment, clipped to [-1, BP_WSEG + 1] (-1: begins before the segment)
        int pb = (int)max(min(myptr - k0, (long long)(BP_WSEG + 1)), -1LL);
        mbar_wait(&bar[buf], (phase >> buf) & 1u);
        phase ^= 1u << buf;
        double *prod = reinterpret_cast<double *>(mine + buf * STAGE);
        const unsigned char *rec = mine + buf * STAGE + BP_WSEG * 8;
        const uint2 idw = *reinterpret_cast<const uint2 *>(rec + lane * 8);
        const unsigned int *bases = reinterpret_cast<const unsigned int *>(rec + BP_WSEG);
        double c[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const unsigned int run = ((u < 4 ? idw.x : idw.y) >> (8 * (u & 3))) & 0xffu;
            c[u] = __ldg(coef + (bases[run] + (unsigned int)(32 * u + lane)));
        }
        if (n_rows == 1) {
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < PER; ++u) s = fma(prod[lane + u * 32], c[u], s);   // padding has weight 0
            s = warp_sum(s);
            const int b = __shfl_sync(0xffffffffu, pb, 0), e = __shfl_sync(0xffffffffu, pb, 1);
            if (lane == 0) {
                if (b >= 0 && e <= kend) out[myvox] = s * myscale;
                else partial[2 * seg + (b > 0 ? 1 : 0)] = s;
            }
        } else {
#pragma unroll
            for (int u = 0; u < PER; ++u) prod[lane + u * 32] *= c[u];
            __syncwarp();
            int r0 = 0;
            while (true) {
                const int nb = min(n_rows - r0, 31);
                double mysum = 0.0;
                // four rows at a time, one per group of 8 lanes (rows of a multi-row segment are short: a whole
                // warp per row spent most of its time in the shuffle tree)
                for (int q0 = 0; q0 < nb; q0 += 4) {
                    const int q = q0 + (lane >> 3);
                    const int b = __shfl_sync(0xffffffffu, pb, min(q, 31)), e = __shfl_sync(0xffffffffu, pb, min(q + 1, 31));
                    const int lo = max(b, 0), hi = (q < nb) ? min(e, kend) : lo;
                    double s = 0.0;
                    for (int j = lo + (lane & 7); j < hi; j += 8) s += prod[j];
                    s += __shfl_xor_sync(0xffffffffu, s, 4);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    const double t = __shfl_sync(0xffffffffu, s, ((lane - q0) & 3) << 3);   // row q0+i sits in group i
                    if (lane >= q0 && lane < q0 + 4) mysum = t;
                }
                const int mye = __shfl_down_sync(0xffffffffu, pb, 1);
                if (lane < nb) {
                    if (pb >= 0 && mye <= kend) out[myvox] = mysum * myscale;
                    else partial[2 * seg + (pb > 0 ? 1 : 0)] = mysum;
                }
                r0 += nb;
                if (r0 >= n_rows) break;
                // next batch of rows (segments made of many tiny rows): loaded on demand
                load_rows(rr.x + r0, n_rows - r0, myptr, myvox);
                myscale = 1.0;
                if (lane < min(n_rows - r0, 31)) myscale = row_scale(scale, myvox);
                pb = (int)max(min(myptr - k0, (long long)(BP_WSEG + 1)), -1LL);
            }
        }
        __syncwarp();
        rr = rr1; rr1 = rr2; v0 = w0; v1 = w1; myptr = nptr; myvox = nvox;
    }
}

// One THREAD per straddling row whose partials lie in at most 8 segments (the common case).
__global__ void __launch_bounds__(256) backproject_combine_short_kernel(const int *__restrict__ rows, int n_rows,
                                                                         int seg, const long long *__restrict__ ptr,
                                                                         const unsigned int *__restrict__ row_voxel,
                                                                         const double *__restrict__ partial,
                                                                         const RowScale scale,
                                                                         double *__restrict__ out) {
    const int stride = gridDim.x * blockDim.x;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const long long row = rows[r];
        const long long b = ptr[row], e = ptr[row + 1];
        const long long s_first = b / seg, s_last = (e - 1) / seg;
        double s = 0.0;
        for (long long sg = s_first; sg <= s_last; ++sg)
            s += partial[2 * sg + ((sg == s_first && b > sg * seg) ? 1 : 0)];
        const long long v = row_voxel[row];
        out[v] = s * row_scale(scale, (unsigned int)v);
    }
}

// One warp per straddling row with partials in more than 8 segments (the heaviest voxels sit under the array core
// and collect ~1e6 entries = thousands of segments): every lane keeps 8 independent loads in flight, fixed
// reduction tree.
__global__ void __launch_bounds__(256) backproject_combine_kernel(const int *__restrict__ rows, int n_rows,
                                                                   int BP_SEG,
                                                                   const long long *__restrict__ ptr,
                                                                   const unsigned int *__restrict__ row_voxel,
                                                                   const double *__restrict__ partial,
                                                                   const RowScale scale,
                                                                   double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp_global; r < n_rows; r += n_warps) {
        const long long row = rows[r];
        const long long v = row_voxel[row];
        const long long b = ptr[row], e = ptr[row + 1];
        const long long s_first = b / BP_SEG, s_last = (e - 1) / BP_SEG;
        const int first_slot = (b > s_first * BP_SEG) ? 1 : 0;
        double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (long long sg = s_first + lane; sg <= s_last; sg += 256) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const long long q = sg + 32 * u;
                if (q <= s_last) a[u] += partial[2 * q + (q == s_first ? first_slot : 0)];
            }
        }
        double s = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
        s = warp_sum(s);
        if (lane == 0) out[v] = s * row_scale(scale, (unsigned int)v);
    }
}

// Build time: boundaries of 16 equal chunks of the segment range.  Chunk c owns the rows whose LAST
// entry lies in its segments, i.e. rows [row(c), row(c+1)) with row(c) = #rows ending at or before
// the chunk's first entry; voxels likewise (rows are sorted by voxel).
__global__ void chunk_table_kernel(const long long *__restrict__ ptr, const unsigned int *__restrict__ row_voxel,
                                   long long n_rows, long long nseg, int seg, long long V,
                                   const int *__restrict__ short_rows, int n_short, const int *__restrict__ vlong_rows,
                                   int n_vlong, long long *__restrict__ tab) {
    const int c = threadIdx.x;
    if (c > 16) return;
    const long long sb = (c == 16) ? nseg : nseg * c / 16;
    const long long first_entry = sb * seg;
    // number of rows with ptr[r+1] <= first_entry
    long long lo = 0, hi = n_rows;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (ptr[mid + 1] <= first_entry) lo = mid + 1; else hi = mid;
    }
    const long long row = (c == 16) ? n_rows : lo;
    long long vox = (c == 0) ? 0 : ((row < n_rows) ? (long long)row_voxel[row] : V);
    if (c == 16) vox = V;
    auto lower = [&](const int *list, int n) {
        int a = 0, b = n;
        while (a < b) { const int m = (a + b) >> 1; if ((long long)list[m] < row) a = m + 1; else b = m; }
        return (long long)a;
    };
    tab[c * 5 + 0] = sb; tab[c * 5 + 1] = row; tab[c * 5 + 2] = vox;
    tab[c * 5 + 3] = lower(short_rows, n_short); tab[c * 5 + 4] = lower(vlong_rows, n_vlong);
}

extern "C" int iono_backprojector_destroy(iono_backprojector_t h) {
    if (!h) return IONO_OK;
    cudaFree(h->ray_idx);
    cudaFree(h->weight);
    cudaFree(h->ptr);
    cudaFree(h->row_voxel);
    cudaFree(h->long_rows);
    cudaFree(h->vlong_rows);
    cudaFree(h->partial);
    cudaFree(h->items);
    cudaFree(h->coef_perm);
    cudaFree(h->runs);
    cudaFree(h->run_ptr);
    delete h;
    return IONO_OK;
}

extern "C" long long iono_backprojector_nnz(iono_backprojector_t h) { return h ? h->nnz : 0; }
extern "C" long long iono_backprojector_bytes(iono_backprojector_t h) {
    if (!h) return 0;
    if (h->use_runs) return h->nnz * 8 + h->run_bytes + ((h->nnz + BP_WSEG - 1) / BP_WSEG + 1) * 8 + (h->n_rows + 1) * 12;
    return h->nnz * 12 + (h->n_rows + 1) * 12;
}

extern "C" int iono_backprojector_create(iono_grid_t grid, const double *rays, int Na, int Nt, int Nd, int Ns,
                                         iono_backprojector_t *out, unsigned long long *oob_count, void *stream) {
    const long long R = (long long)Na * Nt * Nd;
    if (!grid || !out || !oob_count || Na < 0 || Nt < 0 || Nd < 0 || Ns < 1 || (R > 0 && !rays))
        return fail(IONO_EBADARG, "iono_backprojector_create: bad argument");
    if (sweep_size_check(grid, R, Ns)) return IONO_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const long long V = (long long)grid->nx * grid->ny * grid->nz;
    const long long N = (Ns >= 2) ? R * Ns * 8 : 0;
    CU_CHECK(cudaMemsetAsync(oob_count, 0, sizeof(unsigned long long), st));

    iono_backprojector *h = new iono_backprojector();
    for (int c = 0; c <= 16; ++c) { h->chunk_seg[c] = 0; h->chunk_row[c] = 0; h->chunk_vox[c] = (c == 16) ? V : 0; h->chunk_short[c] = 0; h->chunk_vlong[c] = 0; }
    h->seg = BP_WSEG;
    const int BP_SEG = h->seg;
    h->ray_idx = nullptr; h->weight = nullptr; h->ptr = nullptr; h->row_voxel = nullptr; h->n_rows = 0; h->long_rows = nullptr; h->n_long = 0; h->vlong_rows = nullptr; h->n_vlong = 0; h->partial = nullptr; h->items = nullptr;
    h->nnz = 0; h->V = V; h->R = R; h->Na = Na; h->Nt = Nt; h->Nd = Nd; h->coef_perm = nullptr;
    h->runs = nullptr; h->run_ptr = nullptr; h->run_bytes = 0; h->use_runs = 0;
    cudaGetDevice(&h->device);
    unsigned long long *k0 = nullptr, *k1 = nullptr, *uk = nullptr;
    double *v0 = nullptr, *v1 = nullptr;
    long long *d_runs = nullptr;
    void *tmp = nullptr;
    cudaError_t e = cudaSuccess;
    auto cleanup = [&]() {
        cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(uk); cudaFree(d_runs); cudaFree(tmp);
    };
#define BP_TRY(expr)                                                                              \
    do {                                                                                          \
        e = (expr);                                                                               \
        if (e != cudaSuccess) {                                                                   \
            cleanup();                                                                            \
            iono_backprojector_destroy(h);                                                        \
            return fail(IONO_ECUDA, "iono_backprojector_create: %s: %s", #expr, cudaGetErrorString(e)); \
        }                                                                                         \
    } while (0)

    BP_TRY(cudaMalloc(&h->coef_perm, (size_t)(R > 0 ? R : 1) * sizeof(double)));
    long long M = 0;
    if (N > 0) {
        BP_TRY(cudaMalloc(&k0, N * 8));
        BP_TRY(cudaMalloc(&v0, N * 8));
        BP_TRY(cudaMalloc(&k1, N * 8));
        BP_TRY(cudaMalloc(&v1, N * 8));
        const int ctas = sm_count() * 8;
        if (grid->exact)
            emit_entries_kernel<2><<<ctas, 256, 0, st>>>(grid->dev, rays, (int)R, Nt, Nd, Ns, k0, v0, oob_count);
        else if (grid->uniform)
            emit_entries_kernel<1><<<ctas, 256, 0, st>>>(grid->dev, rays, (int)R, Nt, Nd, Ns, k0, v0, oob_count);
        else
            emit_entries_kernel<0><<<ctas, 256, 0, st>>>(grid->dev, rays, (int)R, Nt, Nd, Ns, k0, v0, oob_count);
        BP_TRY(cudaGetLastError());
        // sort by (voxel, ray): only the populated key bits
        int rbits = 1, vbits = 1;
        while ((1LL << rbits) < R) ++rbits;
        while ((1LL << vbits) < V) ++vbits;
        (void)rbits;
        const int end_bit = 32 + vbits;
        cub::DoubleBuffer<unsigned long long> dk(k0, k1);
        cub::DoubleBuffer<double> dv(v0, v1);
        size_t tmp_bytes = 0;
        BP_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, N, 0, end_bit, st));
        BP_TRY(cudaMalloc(&tmp, tmp_bytes));
        BP_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, N, 0, end_bit, st));
        cudaFree(tmp); tmp = nullptr;
        unsigned long long *ks = dk.Current(), *kalt = dk.Alternate();
        double *vs = dv.Current(), *valt = dv.Alternate();
        // merge runs of equal (voxel, ray): unique keys -> kalt, sums -> valt
        BP_TRY(cudaMalloc(&d_runs, sizeof(long long)));
        tmp_bytes = 0;
        BP_TRY(cub::DeviceReduce::ReduceByKey(nullptr, tmp_bytes, ks, kalt, vs, valt, d_runs, ::cuda::std::plus<>{}, N, st));
        BP_TRY(cudaMalloc(&tmp, tmp_bytes));
        BP_TRY(cub::DeviceReduce::ReduceByKey(tmp, tmp_bytes, ks, kalt, vs, valt, d_runs, ::cuda::std::plus<>{}, N, st));
        BP_TRY(cudaMemcpyAsync(&M, d_runs, sizeof(long long), cudaMemcpyDeviceToHost, st));
        BP_TRY(cudaStreamSynchronize(st));
        cudaFree(tmp); tmp = nullptr;
        // DeviceRunLengthEncode (row table below) counts with a plain int in this CUB, and row/segment
        // indices are ints throughout the apply: refuse larger operators instead of truncating
        if (M > 0x7fffffffLL - 2 * BP_SEG) {
            cleanup();
            iono_backprojector_destroy(h);
            return fail(IONO_EBADARG, "iono_backprojector_create: more than 2^31 operator entries; shard the rays "
                                      "(or use the stateless adjoint)");
        }
        // release the sorted inputs before allocating the final arrays
        if (ks == k0) { cudaFree(k0); k0 = nullptr; cudaFree(v0); v0 = nullptr; }
        else          { cudaFree(k1); k1 = nullptr; cudaFree(v1); v1 = nullptr; }
        const long long Mpad = ((M + BP_SEG - 1) / BP_SEG) * BP_SEG + BP_SEG;   // whole segments (TMA copies)
        BP_TRY(cudaMalloc(&h->ray_idx, Mpad * sizeof(unsigned int)));
        BP_TRY(cudaMalloc(&h->weight, Mpad * sizeof(double)));
        BP_TRY(cudaMemsetAsync(h->ray_idx, 0, Mpad * sizeof(unsigned int), st));
        BP_TRY(cudaMemsetAsync(h->weight, 0, Mpad * sizeof(double), st));
        BP_TRY(cudaMemcpyAsync(h->weight, valt, M * sizeof(double), cudaMemcpyDeviceToDevice, st));
        split_keys_kernel<<<ew_grid(M), 256, 0, st>>>(kalt, M, h->ray_idx);
        BP_TRY(cudaGetLastError());
        // non-empty rows: run-length encode the voxel part of the sorted unique keys
        {
            const long long max_rows = (M < V ? M : V) + 1;
            long long *d_len = nullptr;
            BP_TRY(cudaMalloc(&h->row_voxel, (size_t)max_rows * sizeof(unsigned int)));
            BP_TRY(cudaMalloc(&h->ptr, (size_t)(max_rows + 1) * sizeof(long long)));
            e = cudaMalloc(&d_len, (size_t)max_rows * sizeof(long long));
            if (e != cudaSuccess) { cleanup(); iono_backprojector_destroy(h); return fail(IONO_ECUDA, "cudaMalloc(row lengths): %s", cudaGetErrorString(e)); }
            cub::TransformInputIterator<unsigned int, KeyVoxel, const unsigned long long *> vox(kalt, KeyVoxel());
            tmp_bytes = 0;
            e = cub::DeviceRunLengthEncode::Encode(nullptr, tmp_bytes, vox, h->row_voxel, d_len, d_runs, M, st);
            if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes);
            if (e == cudaSuccess) e = cub::DeviceRunLengthEncode::Encode(tmp, tmp_bytes, vox, h->row_voxel, d_len, d_runs, M, st);
            long long nr = 0;
            if (e == cudaSuccess) e = cudaMemcpyAsync(&nr, d_runs, sizeof(long long), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            cudaFree(tmp); tmp = nullptr;
            if (e == cudaSuccess) {
                h->n_rows = nr;
                tmp_bytes = 0;
                e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len, h->ptr, nr, st);
                if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes);
                if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_len, h->ptr, nr, st);
                if (e == cudaSuccess) e = cudaMemcpyAsync(h->ptr + nr, &M, sizeof(long long), cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                cudaFree(tmp); tmp = nullptr;
            }
            cudaFree(d_len);
            if (e != cudaSuccess) { cleanup(); iono_backprojector_destroy(h); return fail(IONO_ECUDA, "iono_backprojector_create: row table: %s", cudaGetErrorString(e)); }
        }
        // segment tables
        const long long nseg = (M + BP_SEG - 1) / BP_SEG;
        int *d_count = reinterpret_cast<int *>(d_runs);
        BP_TRY(cudaMalloc(&h->partial, (size_t)(2 * nseg + 2) * sizeof(double)));
        BP_TRY(cudaMalloc(&h->items, (size_t)(nseg + 1) * sizeof(int2)));
        BP_TRY(cudaMalloc(&h->long_rows, (size_t)(nseg + 1) * sizeof(int)));   // <= one straddler per boundary
        BP_TRY(cudaMalloc(&h->vlong_rows, (size_t)(nseg / 7 + 2) * sizeof(int)));
        BP_TRY(cudaMemsetAsync(d_count, 0, 2 * sizeof(int), st));
        if (nseg > 0) {
            segment_rows_kernel<<<ew_grid(nseg), 256, 0, st>>>(h->ptr, h->n_rows, M, BP_SEG, h->items);
            BP_TRY(cudaGetLastError());
            find_straddling_rows_kernel<<<ew_grid(h->n_rows), 256, 0, st>>>(h->ptr, h->n_rows, BP_SEG, h->long_rows, d_count,
                                                                          (int)nseg + 1, h->vlong_rows, d_count + 1,
                                                                          (int)(nseg / 7 + 2));
            BP_TRY(cudaGetLastError());
        }
        BP_TRY(cudaMemcpyAsync(&h->n_long, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        BP_TRY(cudaMemcpyAsync(&h->n_vlong, d_count + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        BP_TRY(cudaStreamSynchronize(st));
        if (h->n_long > (int)nseg + 1) h->n_long = (int)nseg + 1;            // cannot happen (one straddler per boundary)
        if (h->n_vlong > (int)(nseg / 7 + 2)) h->n_vlong = (int)(nseg / 7 + 2);
        // sort the straddler lists by row and cut everything into 16 chunks of segments
        {
            int *lists[2] = {h->long_rows, h->vlong_rows};
            const int counts[2] = {h->n_long, h->n_vlong};
            for (int li = 0; li < 2; ++li) {
                if (counts[li] < 2) continue;
                int *alt = nullptr;
                BP_TRY(cudaMalloc(&alt, (size_t)counts[li] * sizeof(int)));
                tmp_bytes = 0;
                e = cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, lists[li], alt, counts[li], 0, 32, st);
                if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes);
                if (e == cudaSuccess) e = cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, lists[li], alt, counts[li], 0, 32, st);
                if (e == cudaSuccess) e = cudaMemcpyAsync(lists[li], alt, (size_t)counts[li] * sizeof(int), cudaMemcpyDeviceToDevice, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                cudaFree(tmp); tmp = nullptr;
                cudaFree(alt);
                if (e != cudaSuccess) { cleanup(); iono_backprojector_destroy(h); return fail(IONO_ECUDA, "iono_backprojector_create: sort: %s", cudaGetErrorString(e)); }
            }
            long long *d_tab = nullptr;
            BP_TRY(cudaMalloc(&d_tab, 17 * 5 * sizeof(long long)));
            chunk_table_kernel<<<1, 32, 0, st>>>(h->ptr, h->row_voxel, h->n_rows, nseg, BP_SEG, V, h->long_rows, h->n_long,
                                                 h->vlong_rows, h->n_vlong, d_tab);
            long long tab[17 * 5];
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(tab, d_tab, sizeof(tab), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            cudaFree(d_tab);
            if (e != cudaSuccess) { cleanup(); iono_backprojector_destroy(h); return fail(IONO_ECUDA, "iono_backprojector_create: chunk table: %s", cudaGetErrorString(e)); }
            for (int c = 0; c <= 16; ++c) {
                h->chunk_seg[c] = tab[c * 5 + 0]; h->chunk_row[c] = tab[c * 5 + 1]; h->chunk_vox[c] = tab[c * 5 + 2];
                h->chunk_short[c] = (int)tab[c * 5 + 3]; h->chunk_vlong[c] = (int)tab[c * 5 + 4];
            }
        }
        // optional: run-compressed ray indices for the warp-private apply
        const char *er = getenv("IONO_BP_RUNS");
        if (!(er && atoi(er) == 0) && nseg > 0) {
            BP_TRY(cudaMalloc(&h->run_ptr, (size_t)(nseg + 1) * sizeof(unsigned long long)));
            run_record_units_kernel<<<ew_grid(nseg * 32), 256, 0, st>>>(h->ray_idx, nseg, h->run_ptr);
            BP_TRY(cudaGetLastError());
            tmp_bytes = 0;
            BP_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->run_ptr, h->run_ptr, nseg + 1, st));
            BP_TRY(cudaMalloc(&tmp, tmp_bytes));
            BP_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, h->run_ptr, h->run_ptr, nseg + 1, st));
            unsigned long long total_units = 0;
            BP_TRY(cudaMemcpyAsync(&total_units, h->run_ptr + nseg, sizeof(total_units), cudaMemcpyDeviceToHost, st));
            BP_TRY(cudaStreamSynchronize(st));
            cudaFree(tmp); tmp = nullptr;
            h->run_bytes = (long long)total_units * 16;
            BP_TRY(cudaMalloc(&h->runs, (size_t)h->run_bytes + 16));
            BP_TRY(cudaMemsetAsync(h->runs, 0, (size_t)h->run_bytes + 16, st));
            run_record_fill_kernel<<<ew_grid(nseg * 32), 256, 0, st>>>(h->ray_idx, nseg, h->run_ptr, h->runs);
            BP_TRY(cudaGetLastError());
            BP_TRY(cudaStreamSynchronize(st));
            cudaFree(h->ray_idx); h->ray_idx = nullptr;   // not read again in this mode
            h->use_runs = 1;
        }
    }
#undef BP_TRY
    cleanup();
    h->nnz = M;
    *out = h;
    return IONO_OK;
}

// combine the straddling rows completed in chunks [c0, c1)
static cudaError_t bp_combine(iono_backprojector_t h, const unsigned int *row_dst, const RowScale scale, double *out,
                              int c0, int c1, cudaStream_t st) {
    const int ns = h->chunk_short[c1] - h->chunk_short[c0], nv = h->chunk_vlong[c1] - h->chunk_vlong[c0];
    if (ns > 0)
        backproject_combine_short_kernel<<<ew_grid(ns), 256, 0, st>>>(h->long_rows + h->chunk_short[c0], ns, h->seg,
                                                                    h->ptr, row_dst, h->partial, scale, out);
    if (nv > 0)
        backproject_combine_kernel<<<(nv + 7) / 8, 256, 0, st>>>(h->vlong_rows + h->chunk_vlong[c0], nv, h->seg, h->ptr,
                                                                 row_dst, h->partial, scale, out);
    return cudaGetLastError();
}

// Chunks c0..c1-1 (sixteenths of the entry stream).  Chunk 0 also clears `out` and permutes the
// coefficients, so the chunks of one apply must be issued in increasing order on one stream.  After the
// call, out[chunk_voxels(c0) : chunk_voxels(c1)) is final -- the caller may start summing that
// slice across GPUs while the next chunks are computed.
// row_dst == NULL: rows are stored at their voxel index and `out` (V doubles) is cleared first;
// row_dst != NULL: row r is stored at out[row_dst[r]] and the caller has cleared `out`.
static int bp_apply_chunks(iono_backprojector_t h, const double *coef, bool permuted, const RowScale scale,
                           const unsigned int *row_dst, double *out, int c0, int c1, cudaStream_t st) {
    if (!h || !out || (h->R > 0 && !coef) || c0 < 0 || c1 > 16 || c0 >= c1)
        return fail(IONO_EBADARG, "iono_backprojector_apply: bad argument");
    if (device_check(h->device, "iono_backprojector_apply")) return IONO_EBADARG;
    const int ctas = sm_count() * 8;
    if (c0 == 0) {
        if (!row_dst) CU_CHECK(cudaMemsetAsync(out, 0, (size_t)h->V * sizeof(double), st));   // voxels no ray touches
        if (h->nnz == 0) return IONO_OK;
        if (!permuted) {
            permute_coef_kernel<<<ctas, 256, 0, st>>>(coef, h->Na, h->Nt, h->Nd, h->coef_perm);
            CU_CHECK(cudaGetLastError());
        }
    }
    if (h->nnz == 0) return IONO_OK;
    const double *coef_int = permuted ? coef : h->coef_perm;
    const unsigned int *dst = row_dst ? row_dst : h->row_voxel;
    const long long sb = h->chunk_seg[c0], se = h->chunk_seg[c1];
    const long long nseg = se - sb;
    if (nseg > 0) {
        int warps = 8, per_sm = 4;
        if (const char *ew = getenv("IONO_BP_WARPS")) warps = atoi(ew);
        if (const char *ec = getenv("IONO_BP_CTAS")) per_sm = atoi(ec);
        if (warps < 1 || warps > 8) warps = 8;
        if (per_sm < 1) per_sm = 1;
        const long long cap = (long long)sm_count() * per_sm;
        const long long want = (nseg + warps - 1) / warps;
        const int ctas_seg = (int)(want < cap ? want : cap);
        if (h->use_runs) {
            const int smem_r = warps * 2 * (BP_WSEG * 8 + BP_RUNREC_MAX) + warps * 16 + 64;
            CU_CHECK(cudaFuncSetAttribute(backproject_wruns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r));
            backproject_wruns_kernel<<<ctas_seg, warps * 32, smem_r, st>>>(
                h->items, h->ptr, dst, h->runs, h->run_ptr, h->weight, coef_int, scale, h->nnz, sb, se,
                out, h->partial);
        } else {
            const int smem = warps * 2 * BP_WSEG * 12 + warps * 16 + 64;
            CU_CHECK(cudaFuncSetAttribute(backproject_wsegments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            backproject_wsegments_kernel<<<ctas_seg, warps * 32, smem, st>>>(
                h->items, h->ptr, dst, h->ray_idx, h->weight, coef_int, scale, h->nnz, sb, se, out,
                h->partial);
        }
        CU_CHECK(cudaGetLastError());
    }
    CU_CHECK(bp_combine(h, dst, scale, out, c0, c1, st));
    return IONO_OK;
}

extern "C" int iono_backprojector_apply_chunks_f64(iono_backprojector_t h, const double *coef, const double *scale,
                                                   double *out, int c0, int c1, void *stream) {
    return bp_apply_chunks(h, coef, false, RowScale{scale, 1.0, scale ? 1 : 0}, nullptr, out, c0, c1, (cudaStream_t)stream);
}

// `coef_perm` already in the operator's internal ray order (antenna, direction, time) -- what
// iono_residual_f64 writes -- so no permutation pass; chunks as above.
extern "C" int iono_backprojector_apply_permuted_f64(iono_backprojector_t h, const double *coef_perm,
                                                     const double *scale, double *out, int c0, int c1, void *stream) {
    return bp_apply_chunks(h, coef_perm, true, RowScale{scale, 1.0, scale ? 1 : 0}, nullptr, out, c0, c1, (cudaStream_t)stream);
}

// The voxel gradient in one call: out[v] = k * exp(m[v]) * sum_ray A[v,ray] coef[ray] -- the chain-rule factor
// ne[v] = K_ne exp(m[v]) / 1e13 (k = K_ne/1e13) is evaluated for the touched rows only, no ne grid needed.
extern "C" int iono_backprojector_apply_gradient_f64(iono_backprojector_t h, const double *coef_perm, const double *m,
                                                     double k, double *out, int c0, int c1, void *stream) {
    if (!m) return fail(IONO_EBADARG, "iono_backprojector_apply_gradient_f64: bad argument");
    return bp_apply_chunks(h, coef_perm, true, RowScale{m, k, 2}, nullptr, out, c0, c1, (cudaStream_t)stream);
}

// Sharded adjoint: row r of this rank's operator is stored (unscaled) at out_compact[row_dst[r]], where the
// caller numbers the voxels any rank touches consecutively (row voxels: iono_backprojector_row_voxels) and has
// cleared out_compact; the cross-rank sum then moves that compact vector instead of the whole grid.
extern "C" int iono_backprojector_apply_compact_f64(iono_backprojector_t h, const double *coef_perm,
                                                    const unsigned int *row_dst, double *out_compact, int c0, int c1,
                                                    void *stream) {
    if (!row_dst) return fail(IONO_EBADARG, "iono_backprojector_apply_compact_f64: bad argument");
    return bp_apply_chunks(h, coef_perm, true, RowScale{nullptr, 1.0, 0}, row_dst, out_compact, c0, c1,
                           (cudaStream_t)stream);
}

__global__ void __launch_bounds__(256) ne_rows_kernel(const unsigned int *__restrict__ row_voxel, long long n_rows,
                                                      const double *__restrict__ m, double k, double *__restrict__ ne) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const unsigned int v = row_voxel[r];
        ne[v] = k * exp(m[v]);
    }
}

// ne_out[v] = k * exp(m[v]) for the voxels of the operator's rows only (the chain-rule factor the apply needs as
// `scale`; the other voxels of ne_out are left untouched): ~1.6 M exps instead of a pass over the 8.4 M-voxel grid.
extern "C" int iono_backprojector_ne_rows_f64(iono_backprojector_t h, const double *m, double k, double *ne_out,
                                              void *stream) {
    if (!h || !m || !ne_out) return fail(IONO_EBADARG, "iono_backprojector_ne_rows_f64: bad argument");
    if (device_check(h->device, "iono_backprojector_ne_rows_f64")) return IONO_EBADARG;
    if (h->n_rows == 0) return IONO_OK;
    ne_rows_kernel<<<ew_grid(h->n_rows), 256, 0, (cudaStream_t)stream>>>(h->row_voxel, h->n_rows, m, k, ne_out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

extern "C" long long iono_backprojector_n_rows(iono_backprojector_t h) { return h ? h->n_rows : 0; }

// voxel index of every non-empty row, ascending (device array of n_rows uint32)
extern "C" int iono_backprojector_row_voxels(iono_backprojector_t h, unsigned int *out, void *stream) {
    if (!h || (h->n_rows > 0 && !out)) return fail(IONO_EBADARG, "iono_backprojector_row_voxels: bad argument");
    if (h->n_rows > 0)
        CU_CHECK(cudaMemcpyAsync(out, h->row_voxel, (size_t)h->n_rows * sizeof(unsigned int), cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream));
    return IONO_OK;
}

extern "C" long long iono_backprojector_chunk_voxels(iono_backprojector_t h, int c) {
    return (h && c >= 0 && c <= 16) ? h->chunk_vox[c] : -1;
}

extern "C" int iono_backprojector_apply_f64(iono_backprojector_t h, const double *coef, const double *scale,
                                            double *out, void *stream) {
    if (!h || !out || (h->R > 0 && !coef)) return fail(IONO_EBADARG, "iono_backprojector_apply_f64: bad argument");
    return iono_backprojector_apply_chunks_f64(h, coef, scale, out, 0, 16, stream);
}
