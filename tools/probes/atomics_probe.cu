// Micro-benchmark: throughput of fire-and-forget global reductions (RED) on B200 for the
// access shapes of the adjoint scatter.  Each warp instruction touches `runs` distinct
// column runs of `32/runs` consecutive elements (so runs*ceil(32/runs*8/32) sectors),
// at pseudo-random places of a 64 MiB array.  Types: f64, u64, f32.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <typename T> __device__ void red(T *p, T v);
template <> __device__ void red<double>(double *p, double v) { atomicAdd(p, v); }
template <> __device__ void red<unsigned long long>(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }
template <> __device__ void red<float>(float *p, float v) { atomicAdd(p, v); }

template <typename T>
__global__ void probe(T *a, size_t n_elems, int runs, int iters, int active_mask_mod) {
    const int lane = threadIdx.x & 31;
    const int per = 32 / runs;
    const int run = lane / per, within = lane % per;
    unsigned long long warp_id = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    unsigned long long state = warp_id * 0x9E3779B97F4A7C15ull + 12345;
    for (int it = 0; it < iters; ++it) {
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        // each run starts at an independent pseudo-random, sector-aligned place
        unsigned long long h = (state >> 11) + (unsigned long long)run * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        size_t base = (h % (n_elems / 64)) * 64;   // 64-element aligned
        if (active_mask_mod <= 1 || (lane % active_mask_mod) == 0) red<T>(a + base + within, (T)1);
    }
}

template <typename T>
void run(const char *name, int runs, int mod) {
    const size_t bytes = 64ull << 20;
    T *a;
    cudaMalloc(&a, bytes);
    cudaMemset(a, 0, bytes);
    const int iters = 2000, blocks = 148 * 4, threads = 512;
    probe<T><<<blocks, threads>>>(a, bytes / sizeof(T), runs, 50, mod);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<T><<<blocks, threads>>>(a, bytes / sizeof(T), runs, iters, mod);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)blocks * threads / 32 * iters;
    int per = 32 / runs;
    int active = (mod <= 1) ? 32 : (32 + mod - 1) / mod;
    double sectors_per = runs * ((per * sizeof(T) + 31) / 32);
    printf("%-4s runs=%2d lanes_active=%2d sectors/inst=%4.0f : %7.3f ms  %7.2f G warp-inst/s  %7.2f G lane-ops/s  %7.2f G sector-ops/s\n",
           name, runs, active, sectors_per, ms, winst / ms / 1e6, winst * active / ms / 1e6, winst * sectors_per / ms / 1e6);
    cudaFree(a);
}

int main() {
    int runs_list[] = {1, 2, 4, 8, 16, 32};
    for (int r : runs_list) run<double>("f64", r, 1);
    for (int r : runs_list) run<unsigned long long>("u64", r, 1);
    for (int r : runs_list) run<float>("f32", r, 1);
    // fewer active lanes, same spread (every 2nd / 4th lane)
    run<double>("f64", 32, 2);
    run<double>("f64", 32, 4);
    run<double>("f64", 8, 2);
    return 0;
}
