// sd_bdot.cuh -- block dot for the check pass of the full reorthogonalisation (Lanczos.jl:142-153).
// The reference evaluates  overlap_k = |dot(v_k, w / beta)|  for k = 1..j one at a time and only modifies w when an
// overlap exceeds orthogonalize_tol -- which after the Gram-Schmidt sweep of :116-122 is the rare case.  As long as w
// is not modified the j dots are independent, so they can be taken eight at a time in one pass over w (9 vector reads
// per 8 dots instead of 16) and fetched with ONE synchronisation instead of j; the first overlap above the tolerance
// (if any) is then handled exactly like the reference does, and the pass resumes behind it (sd_lanczos_groundstate,
// SD_BATCH_CHECK=1).  f64 only (the ground-state basis is real).
#pragma once
#include "sd_kernels.cuh"

// partials[j * nparts + block] = sum over the block's elements of V_j[i] * w[i], j < m <= SD_BDOT_MAX
__global__ void __launch_bounds__(SD_BLAS_THREADS)
sd_bdot_f64_kernel(uint64_t n, const __grid_constant__ SdPtrBlock V, int m, const double *w, double *partials, unsigned nparts) {
    __shared__ double scratch[SD_BDOT_MAX][16];
    double red[SD_BDOT_MAX];
#pragma unroll
    for (int j = 0; j < SD_BDOT_MAX; ++j) red[j] = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double wi = w[i];
#pragma unroll
        for (int j = 0; j < SD_BDOT_MAX; ++j)
            if (j < m) red[j] += V.v[j][i] * wi;
    }
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31u) >> 5;
#pragma unroll
    for (int j = 0; j < SD_BDOT_MAX; ++j) {
        const double s = sd_warp_sum(red[j]);
        if (lane == 0) scratch[j][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < (unsigned)m) {
        double t = 0.0;
        for (unsigned k = 0; k < nwarp; ++k) t += scratch[threadIdx.x][k];
        partials[(size_t)threadIdx.x * nparts + blockIdx.x] = t;
    }
}
// result[j] = sum over blocks of partials[j * nparts + block] in block order (run-to-run identical), j < m
__global__ void __launch_bounds__(SD_BDOT_MAX * 32) sd_bdot_reduce_kernel(const double *partials, unsigned nparts, int m, double *result) {
    const unsigned j = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if ((int)j >= m) return;
    double t = 0.0;
    for (unsigned i = lane; i < nparts; i += 32) t += partials[(size_t)j * nparts + i];
    t = sd_warp_sum(t);
    if (lane == 0) result[j] = t;
}
