// sd_reorth.cuh -- one Lanczos step of lanczos_groundstate (Lanczos.jl:116-155) behind the apply, as ONE cooperative
// kernel: the modified Gram-Schmidt sweep against V_1 .. V_{j-1}, alpha_j, the three-term update, beta_j, the check pass
// over V_1 .. V_j and V_{j+1} = w / beta_j.  The reference does these as 2 (j - 1) + 2 + 2 j + 1 BLAS-1 calls with a
// scalar on the host between each pair; here the scalars never leave the device and a grid-wide barrier stands where
// the reference has a sequence point.  The survey's block_dot / block_axpy: the sweep is sequential by definition
// (dot k uses the w updated by k - 1), so the update by c_k is fused with the dot for k + 1 -- one pass over w per basis
// vector (3 reads + 1 write) instead of two (2 reads; 2 reads + 1 write) -- and the check pass, whose dots are
// independent while w is not modified, takes them SD_RCHK at a time in one pass (SD_RCHK + 1 reads).
//
// Every sum is taken in a fixed order (fixed shuffle tree per warp, warps in order, CTAs in order by every CTA for
// itself), so all CTAs see bit-identical scalars, take the same branches and the result is the same from run to run
// (test_Lanczos.jl:122-166).  f64, single GPU (a sharded model keeps the one-call-per-operation path with NCCL sums).
#pragma once
#include <cooperative_groups.h>
#include "sd_common.h"

#define SD_RTH_THREADS 256
#define SD_RCHK 8                       // overlaps per check pass

struct SdReorthArgs {
    double *w;                          // in: H v_j; out: the residual (unnormalised)
    const double *const *V;             // device table of V_1 .. V_j (V[j-1] = v_j)
    double *vnext;                      // out: V_{j+1} = w / beta_j (nullptr: last step, j = m)
    int j;
    uint64_t n;
    const double *beta_prev;            // device: beta_{j-1}, the out[1] of the previous step (nullptr for j = 1)
    double *stop;                       // device flag: set by the step that hits the breakdown test at :136; later steps return at once
    double tol, orth_tol;
    double *partials;                   // [2][SD_RCHK][gridDim.x]
    double *out;                        // out[0] = alpha_j, out[1] = beta_j, out[2] = 1: beta_j < tol at :136 (stop), 2: inside the check pass (m_actual = j, go on), out[3] = 1
};

namespace sd_rth {
namespace cg = cooperative_groups;

// block sum of NV values per thread -> partials[v * gridDim.x + blockIdx.x]
template <int NV>
__device__ __forceinline__ void block_partials(const double (&r)[NV], int nv, double *partials, double (*scratch)[SD_RTH_THREADS / 32]) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        if (v >= nv) break;
        double t = r[v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (lane == 0) scratch[v][warp] = t;
    }
    __syncthreads();
    if ((int)threadIdx.x < nv) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < SD_RTH_THREADS / 32; ++k) t += scratch[threadIdx.x][k];
        partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = t;
    }
}
// after the grid barrier: total[v] = sum over CTAs in CTA order, computed identically by every CTA (warp v sums value v)
__device__ __forceinline__ void grid_totals(int nv, const double *partials, double *total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if ((int)warp < nv) {
        double t = 0.0;
        for (unsigned i = lane; i < gridDim.x; i += 32) t += partials[(size_t)warp * gridDim.x + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (lane == 0) total[warp] = t;
    }
    __syncthreads();
}
}  // namespace sd_rth

__global__ void __launch_bounds__(SD_RTH_THREADS) sd_reorth_step_kernel(const __grid_constant__ SdReorthArgs A) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double scratch[SD_RCHK][SD_RTH_THREADS / 32];
    __shared__ double total[SD_RCHK];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // A solve enqueues all its steps without waiting for their scalars; once a step has met the reference's `break` the
    // remaining ones do nothing.  Every CTA reads the flag before the first grid barrier, the flag is written behind the
    // last one, so all CTAs of a launch see the same value.
    if (*A.stop != 0.0) return;
    double *w = A.w;
    const int j = A.j;
    const size_t pstride = (size_t)SD_RCHK * gridDim.x;
    int phase = 0;
    auto pbuf = [&]() { return A.partials + (size_t)(phase & 1) * pstride; };
    // ---- sweep (Lanczos.jl:116-124): d_1 = <V_1, w>; for k = 1 .. j-1: w -= d_k V_k, d_{k+1} = <V_{k+1}, w>; alpha_j = d_j
    {
        double r[1] = {0.0};
        const double *v0 = A.V[0];
        for (uint64_t i = i0; i < A.n; i += stride) r[0] += v0[i] * w[i];
        sd_rth::block_partials<1>(r, 1, pbuf(), scratch);
        grid.sync();
        sd_rth::grid_totals(1, pbuf(), total);
        ++phase;
    }
    for (int k = 0; k + 1 < j; ++k) {
        const double c = total[0];
        const double *vk = A.V[k], *vn = A.V[k + 1];
        double r[1] = {0.0};
        for (uint64_t i = i0; i < A.n; i += stride) {
            const double x = w[i] - c * vk[i];
            w[i] = x;
            r[0] += vn[i] * x;
        }
        sd_rth::block_partials<1>(r, 1, pbuf(), scratch);
        grid.sync();
        sd_rth::grid_totals(1, pbuf(), total);
        ++phase;
    }
    const double alpha = total[0];
    // ---- three-term update with the fused norm (:126-133)
    double beta;
    {
        const double *vj = A.V[j - 1], *vp = j >= 2 ? A.V[j - 2] : nullptr;
        const double bp = vp ? *A.beta_prev : 0.0;
        double r[1] = {0.0};
        for (uint64_t i = i0; i < A.n; i += stride) {
            double x = w[i] - alpha * vj[i];
            if (vp) x -= bp * vp[i];
            w[i] = x;
            r[0] += x * x;
        }
        sd_rth::block_partials<1>(r, 1, pbuf(), scratch);
        grid.sync();
        sd_rth::grid_totals(1, pbuf(), total);
        ++phase;
        beta = sqrt(total[0]);
    }
    // The reference leaves the outer loop only at the first breakdown test (:136-139).  A breakdown found INSIDE the check
    // pass (:148-151) records m_actual = j, leaves the check pass and still forms V_{j+1} = w / beta_j.
    double flag = 0.0;
    if (A.vnext != nullptr) {                                      // j < m (:132)
        if (beta < A.tol) flag = 1.0;                              // :136-139
        // ---- check pass (:142-153): overlaps SD_RCHK at a time while w is unchanged; the first one above the tolerance is
        // corrected as the reference does (w -= <v_k, w> v_k, beta = ||w||) and the pass resumes behind it
        int kfirst = 0;
        while (flag == 0.0 && kfirst < j) {
            const int nb = min(SD_RCHK, j - kfirst);
            double r[SD_RCHK];
#pragma unroll
            for (int t = 0; t < SD_RCHK; ++t) r[t] = 0.0;
            for (uint64_t i = i0; i < A.n; i += stride) {
                const double wi = w[i];
#pragma unroll
                for (int t = 0; t < SD_RCHK; ++t)
                    if (t < nb) r[t] += A.V[kfirst + t][i] * wi;
            }
            sd_rth::block_partials<SD_RCHK>(r, nb, pbuf(), scratch);
            grid.sync();
            sd_rth::grid_totals(nb, pbuf(), total);
            ++phase;
            int viol = -1;
            for (int t = 0; t < nb && viol < 0; ++t)
                if (fabs(total[t]) / beta > A.orth_tol) viol = t;
            if (viol < 0) { kfirst += nb; continue; }
            const double d = total[viol];
            const double *vk = A.V[kfirst + viol];
            double r2[1] = {0.0};
            for (uint64_t i = i0; i < A.n; i += stride) {
                const double x = w[i] - d * vk[i];
                w[i] = x;
                r2[0] += x * x;
            }
            sd_rth::block_partials<1>(r2, 1, pbuf(), scratch);
            grid.sync();
            sd_rth::grid_totals(1, pbuf(), total);
            ++phase;
            beta = sqrt(total[0]);
            if (beta < A.tol) flag = 2.0;
            kfirst += viol + 1;
        }
        if (flag != 1.0) {                                         // :155
            double *vn = A.vnext;
            for (uint64_t i = i0; i < A.n; i += stride) vn[i] = w[i] / beta;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        A.out[0] = alpha; A.out[1] = beta; A.out[2] = flag; A.out[3] = 1.0;   // [3]: this step ran
        if (flag == 1.0) *A.stop = 1.0;
    }
}
