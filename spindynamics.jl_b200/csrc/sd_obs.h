// sd_obs.h -- device-resident observables of a state vector (reference Observables.jl:14-109):
// what a caller needs to consume psi(t) at L >= 32 without downloading 5 - 145 GB.
//
//   mags[i] = sum_states |psi|^2 s_i(state)                                  magnetization_per_site, :14-37
//   zz[r]   = sum_i sum_states |psi|^2 s_i s_{(i+r) mod L}                   the i-sum of SzSz[i, mod1(i+r, L)], :48-93
// The reference accumulates the full L x L matrix SzSz per thread (L^2 multiply-adds per state) and only ever uses
// its cyclic diagonals; here  sum_i s_i s_{i+r} = (L - 2 popc(state xor rot_r(state))) / 4  -- one rotate, one xor
// and one popcount per (state, r) -- so lane r of a warp owns mags[r] and zz[r] and the 32 states a warp has loaded
// are broadcast with shuffles.  connected_correlations' C_r = (zz[r] - sum_i mags[i] mags[(i+r) mod L]) / L and
// the FFT of structure_factor_Sq stay on the host (L numbers).
// States come from unranking the element's basis rank (sd_common.h), so the kernel runs on rank-ordered data.
#pragma once
#include "sd_common.h"

// contribution of one basis state with weight w to site / distance r (r < L <= 63)
SD_HD void sd_obs_accum(int L, int r, uint64_t s, double w, double &mag, double &zz) {
    const uint64_t mask = (L >= 64) ? ~0ULL : ((1ULL << L) - 1ULL);
    const uint64_t rot = r == 0 ? s : (((s << r) | (s >> (L - r))) & mask);
    mag += w * (((s >> r) & 1ULL) ? 0.5 : -0.5);
    zz += w * (0.25 * (double)(L - 2 * (int)SD_POPC64(s ^ rot)));
}

#if defined(__CUDACC__)
// partials: [nwarps_total][128] = mags[0..63] | zz[0..63] of each warp; summed in warp order by sd_obs_reduce_kernel
// (run-to-run identical results).  first_rank: basis rank of local element 0.
template <int NC>
__global__ void __launch_bounds__(256) sd_obs_kernel(int L, int k, const uint64_t *binom, uint64_t first_rank, uint64_t n,
                                                     const double *v, double *partials) {
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    double mag[2] = {0.0, 0.0}, zz[2] = {0.0, 0.0};
    for (uint64_t c0 = gw * 32; c0 < n; c0 += nw * 32) {
        const uint64_t i = c0 + lane;
        double w = 0.0;
        uint64_t s = 0;
        if (i < n) {
            if (NC == 2) { const double re = v[2 * i], im = v[2 * i + 1]; w = re * re + im * im; }
            else { const double re = v[i]; w = re * re; }
            if (w != 0.0) s = (k < 0) ? first_rank + i : sd_unrank_state(first_rank + i, L, k, binom, SD_BINOM_DIM);
        }
        for (int j = 0; j < 32; ++j) {
            const double wj = __shfl_sync(0xffffffffu, w, j);
            if (wj == 0.0) continue;                                   // Observables.jl:21,58 skip zero weights
            const uint64_t sj = __shfl_sync(0xffffffffu, s, j);
            if ((int)lane < L) sd_obs_accum(L, (int)lane, sj, wj, mag[0], zz[0]);
            if ((int)lane + 32 < L) sd_obs_accum(L, (int)lane + 32, sj, wj, mag[1], zz[1]);
        }
    }
    double *p = partials + gw * 128;
    p[lane] = mag[0]; p[32 + lane] = mag[1]; p[64 + lane] = zz[0]; p[96 + lane] = zz[1];
}
__global__ void __launch_bounds__(128) sd_obs_reduce_kernel(const double *partials, unsigned nwarps, double *result) {
    double t = 0.0;
    for (unsigned w = 0; w < nwarps; ++w) t += partials[(size_t)w * 128 + threadIdx.x];
    result[threadIdx.x] = t;
}
#endif
