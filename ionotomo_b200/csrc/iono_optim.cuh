// Vector algebra of the device-resident inversion driver (L-BFGS in compact form): every product an
// iteration needs in a constant number of passes over the history, no host round trip per dot.
// Included by iono_kernels.cu.
//
// The reference builds its BFGS recursion as a chain of dask tasks, one scalar product (a triple Simpson
// integral over the grid, bfgs_dask.py:165-167 `scalarProduct`) and one axpy at a time
// (bfgs_dask.py:34-56, :165-194).  Here the history lives in ONE matrix H of `rows` vectors of length n
// (rows 0..k-1 = s_i, rows k..2k-1 = y_i), and
//   iono_multi_dot_f64   : out[r] = sum_i w[i] * H[r][i] * x[i]   for all rows in one pass over x
//                          (w optional: the Simpson quadrature weights of the grid make this the reference's
//                          inner product, TriCubic.inner / scalarProduct; NULL = Euclidean)
//   iono_lincomb_f64     : out[i] = c0 * x[i] + sum_r c[r] * H[r][i]     (coefficients read from device memory)
//   iono_gather_f64 / iono_scatter_axpy_f64 : between the grid and the ACTIVE voxels -- the ones some ray
//                          touches; the gradient is identically zero elsewhere, so the optimiser's vectors
//                          only need those entries (a fifth of the grid at the LOFAR case).
// All reductions are two-stage with a fixed tree: bit-reproducible.
#pragma once

constexpr int OPT_MAX_ROWS = 32;
constexpr int OPT_BLOCKS = 592;    // 4 x 148

template <int ROWS>
__global__ void __launch_bounds__(256) multi_dot_kernel(const double *__restrict__ H, long long ld,
                                                         const double *__restrict__ x, const double *__restrict__ w,
                                                         long long n, int rows, double *__restrict__ partial) {
    __shared__ double red[8][ROWS];
    double acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double xi = x[i];
        if (w) xi *= w[i];
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
            if (r < rows) acc[r] = fma(H[r * ld + i], xi, acc[r]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const double t = warp_sum(acc[r]);
        if (lane == 0) red[warp][r] = t;
    }
    __syncthreads();
    if (threadIdx.x < ROWS && (int)threadIdx.x < rows) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x];
        partial[(long long)blockIdx.x * OPT_MAX_ROWS + threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) multi_dot_final_kernel(const double *__restrict__ partial, int blocks, int rows,
                                                               double *__restrict__ out) {
    // one warp per row: lanes stride over the per-CTA partials in a fixed order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < rows; r += 8) {
        double t = 0.0;
        for (int b = lane; b < blocks; b += 32) t += partial[(long long)b * OPT_MAX_ROWS + r];
        t = warp_sum(t);
        if (lane == 0) out[r] = t;
    }
}

// Three right-hand sides at once: out[j][r] = sum_i w[i] H[r][i] H[xr[j]][i] for rows r0 <= r < r0 + ROWS -- the
// products of the whole history with the new gradient, the new y and the new s in ONE pass over the history.
template <int ROWS>
__global__ void __launch_bounds__(256) multi_dot3_kernel(const double *__restrict__ H, long long ld, int r0, int rows,
                                                          int x0, int x1, int x2, const double *__restrict__ w,
                                                          long long n, double *__restrict__ partial) {
    __shared__ double red[8][3 * ROWS];
    double acc[3][ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[0][r] = acc[1][r] = acc[2][r] = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double a = H[x0 * ld + i], b = H[x1 * ld + i], c = H[x2 * ld + i];
        if (w) { const double wi = w[i]; a *= wi; b *= wi; c *= wi; }
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
            if (r0 + r < rows) {
                const double h = H[(r0 + r) * ld + i];
                acc[0][r] = fma(h, a, acc[0][r]);
                acc[1][r] = fma(h, b, acc[1][r]);
                acc[2][r] = fma(h, c, acc[2][r]);
            }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const double t = warp_sum(acc[j][r]);
            if (lane == 0) red[warp][j * ROWS + r] = t;
        }
    __syncthreads();
    if (threadIdx.x < 3 * ROWS) {
        const int j = threadIdx.x / ROWS, r = threadIdx.x % ROWS;
        if (r0 + r < rows) {
            double t = 0.0;
            for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x];
            partial[((long long)blockIdx.x * 3 + j) * OPT_MAX_ROWS + r0 + r] = t;
        }
    }
}

__global__ void __launch_bounds__(256) multi_dot3_final_kernel(const double *__restrict__ partial, int blocks, int rows,
                                                                double *__restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = warp; q < 3 * rows; q += 8) {
        const int j = q / rows, r = q % rows;
        double t = 0.0;
        for (int b = lane; b < blocks; b += 32) t += partial[((long long)b * 3 + j) * OPT_MAX_ROWS + r];
        t = warp_sum(t);
        if (lane == 0) out[j * OPT_MAX_ROWS + r] = t;
    }
}

extern "C" int64_t iono_multi_dot_scratch_elems(void) { return (int64_t)OPT_BLOCKS * OPT_MAX_ROWS * 3; }

// out[j*32 + r] = sum_i w[i] H[r][i] H[x_rows[j]][i], j < 3, r < rows <= 32: the history is read once (in groups of
// 8 rows) for all three right-hand sides, which are rows of H themselves.
extern "C" int iono_multi_dot3_f64(const double *H, int64_t ld, int rows, int x_row0, int x_row1, int x_row2,
                                   const double *w, int64_t n, double *scratch, double *out, void *stream) {
    if (rows < 0 || rows > OPT_MAX_ROWS || n < 0 || ld < n || !out || !scratch || x_row0 < 0 || x_row1 < 0 || x_row2 < 0)
        return fail(IONO_EBADARG, "iono_multi_dot3_f64: bad argument (rows <= 32)");
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0) return IONO_OK;
    if (n == 0) {
        CU_CHECK(cudaMemsetAsync(out, 0, 3 * OPT_MAX_ROWS * sizeof(double), st));
        return IONO_OK;
    }
    if (!H) return fail(IONO_EBADARG, "iono_multi_dot3_f64: NULL pointer");
    long long want = (n + 255) / 256;
    const int blocks = (int)(want < OPT_BLOCKS ? want : OPT_BLOCKS);
    for (int r0 = 0; r0 < rows; r0 += 8)
        multi_dot3_kernel<8><<<blocks, 256, 0, st>>>(H, ld, r0, rows, x_row0, x_row1, x_row2, w, n, scratch);
    CU_CHECK(cudaGetLastError());
    multi_dot3_final_kernel<<<1, 256, 0, st>>>(scratch, blocks, rows, out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

extern "C" int iono_multi_dot_f64(const double *H, int64_t ld, int rows, const double *x, const double *w, int64_t n,
                                  double *scratch, double *out, void *stream) {
    if (rows < 0 || rows > OPT_MAX_ROWS || n < 0 || ld < n || !out || !scratch)
        return fail(IONO_EBADARG, "iono_multi_dot_f64: bad argument (rows <= 32)");
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0) return IONO_OK;
    if (n == 0) {
        CU_CHECK(cudaMemsetAsync(out, 0, rows * sizeof(double), st));
        return IONO_OK;
    }
    if (!H || !x) return fail(IONO_EBADARG, "iono_multi_dot_f64: NULL pointer");
    long long want = (n + 255) / 256;
    const int blocks = (int)(want < OPT_BLOCKS ? want : OPT_BLOCKS);
    if (rows <= 4) multi_dot_kernel<4><<<blocks, 256, 0, st>>>(H, ld, x, w, n, rows, scratch);
    else if (rows <= 12) multi_dot_kernel<12><<<blocks, 256, 0, st>>>(H, ld, x, w, n, rows, scratch);
    else if (rows <= 22) multi_dot_kernel<22><<<blocks, 256, 0, st>>>(H, ld, x, w, n, rows, scratch);
    else multi_dot_kernel<32><<<blocks, 256, 0, st>>>(H, ld, x, w, n, rows, scratch);
    CU_CHECK(cudaGetLastError());
    multi_dot_final_kernel<<<1, 256, 0, st>>>(scratch, blocks, rows, out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

template <int ROWS>
__global__ void __launch_bounds__(256) lincomb_kernel(const double *__restrict__ H, long long ld, int rows,
                                                       const double *__restrict__ coef, const double *__restrict__ x,
                                                       long long n, double *__restrict__ out) {
    __shared__ double c[ROWS + 1];
    if (threadIdx.x <= ROWS && (int)threadIdx.x <= rows) c[threadIdx.x] = coef[threadIdx.x];   // c[0] scales x
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v = x ? c[0] * x[i] : 0.0;
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
            if (r < rows) v = fma(c[r + 1], H[r * ld + i], v);
        out[i] = v;
    }
}

// out[i] = coef[0] * x[i] + sum_{r < rows} coef[r + 1] * H[r][i]; coef: rows + 1 doubles in DEVICE memory.
extern "C" int iono_lincomb_f64(const double *H, int64_t ld, int rows, const double *coef, const double *x, int64_t n,
                                double *out, void *stream) {
    if (rows < 0 || rows > OPT_MAX_ROWS || n < 0 || ld < n || !coef || (rows > 0 && !H))
        return fail(IONO_EBADARG, "iono_lincomb_f64: bad argument (rows <= 32)");
    if (n == 0) return IONO_OK;
    if (!out) return fail(IONO_EBADARG, "iono_lincomb_f64: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = ew_grid(n);
    if (rows <= 4) lincomb_kernel<4><<<blocks, 256, 0, st>>>(H, ld, rows, coef, x, n, out);
    else if (rows <= 12) lincomb_kernel<12><<<blocks, 256, 0, st>>>(H, ld, rows, coef, x, n, out);
    else if (rows <= 22) lincomb_kernel<22><<<blocks, 256, 0, st>>>(H, ld, rows, coef, x, n, out);
    else lincomb_kernel<32><<<blocks, 256, 0, st>>>(H, ld, rows, coef, x, n, out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

__global__ void __launch_bounds__(256) gather_kernel(const double *__restrict__ src, const int *__restrict__ idx,
                                                      long long n, double *__restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = src[idx[i]];
}

// dst[idx[i]] = base[idx[i]] + (*alpha) * x[i]   (alpha in device memory; dst may alias base)
__global__ void __launch_bounds__(256) scatter_axpy_kernel(const double *__restrict__ base, const double *alpha,
                                                            const double *__restrict__ x, const int *__restrict__ idx,
                                                            long long n, double *__restrict__ dst) {
    const double a = *alpha;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int v = idx[i];
        dst[v] = fma(a, x[i], base[v]);
    }
}

__global__ void __launch_bounds__(256) scatter_set_kernel(const double *__restrict__ x, const int *__restrict__ idx,
                                                           long long n, double *__restrict__ dst) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[idx[i]] = x[i];
}

// dst[idx[i]] = x[i]
extern "C" int iono_scatter_set_f64(const double *x, const int *idx, int64_t n, double *dst, void *stream) {
    if (n < 0 || (n > 0 && (!x || !idx || !dst))) return fail(IONO_EBADARG, "iono_scatter_set_f64: bad argument");
    if (n == 0) return IONO_OK;
    scatter_set_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, idx, n, dst);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

extern "C" int iono_gather_f64(const double *src, const int *idx, int64_t n, double *out, void *stream) {
    if (n < 0 || (n > 0 && (!src || !idx || !out))) return fail(IONO_EBADARG, "iono_gather_f64: bad argument");
    if (n == 0) return IONO_OK;
    gather_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(src, idx, n, out);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

extern "C" int iono_scatter_axpy_f64(const double *base, const double *alpha_dev, const double *x, const int *idx,
                                     int64_t n, double *dst, void *stream) {
    if (n < 0 || (n > 0 && (!base || !alpha_dev || !x || !idx || !dst)))
        return fail(IONO_EBADARG, "iono_scatter_axpy_f64: bad argument");
    if (n == 0) return IONO_OK;
    scatter_axpy_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(base, alpha_dev, x, idx, n, dst);
    CU_CHECK(cudaGetLastError());
    return IONO_OK;
}

// x[0..n) = 0 (cudaMemsetAsync: a memset node in a captured graph, no kernel)
extern "C" int iono_zero_f64(double *x, int64_t n, void *stream) {
    if (n < 0 || (n > 0 && !x)) return fail(IONO_EBADARG, "iono_zero_f64: bad argument");
    if (n > 0) CU_CHECK(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), (cudaStream_t)stream));
    return IONO_OK;
}
