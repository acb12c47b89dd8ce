// sd_blkv.h -- vector operations that need the basis STATE of every element, run natively on block-layout vectors
// (sd_blk.h): Sz_q_vector (Hamiltonian.jl:307-337) and the device-resident observables (Observables.jl:14-109).
// In block layout the state of a stored element is known without unranking: prefix bits from the tile key, mid
// configuration from the item table, tail configuration from (class, row) -- so these kernels stream the vector once
// (round 1 staged block-layout vectors through a rank-ordered copy: two or three extra passes and a 4.8 - 9.6 GB
// scratch buffer per call at L = 32, paid once per momentum by lanczos_sqw / kpm_sqw).
#pragma once
#include "sd_blk.h"
#include "sd_obs.h"

#if defined(__CUDACC__)
struct SdBlkSzq {
    double ph_re[SD_MAX_L + 1], ph_im[SD_MAX_L + 1];   // e^{i q r}
    double normfact;                                    // L^-1/2
};
// phi (c128, block layout) = L^-1/2 (sum_r e^{iqr} s_r(state)) * ComplexF64(psi0)  for a block-layout psi0 of NCIN
// components; optional ||phi||^2 -> partials[3 * nparts + cta].  One CTA per tile, grid-stride over the shard's keys.
template <int NCIN>
__global__ void __launch_bounds__(256) sd_blk_szq_kernel(const __grid_constant__ SdBlkParams P, const __grid_constant__ SdBlkSzq Z,
                                                         const double *psi_local, double *phi_local, double *partials, unsigned nparts) {
    __shared__ uint64_t s_base;
    __shared__ double s_pre[2];
    __shared__ double scratch[SD_NSLOT][16];
    const int A = P.A, k = P.k, L = P.L;
    double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
    double half_re = 0.0, half_im = 0.0;                               // sum_r e^{iqr} / 2
    for (int r = 0; r < L; ++r) { half_re += 0.5 * Z.ph_re[r]; half_im += 0.5 * Z.ph_im[r]; }
    for (uint64_t key = P.key_lo + blockIdx.x; key < P.key_hi; key += gridDim.x) {
        const uint64_t Pb = __brevll(~key) >> (64 - A);
        const int js = k - __popcll(Pb);
        if (js < 0 || js > SD_BLK_B) continue;                         // uniform over the CTA
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t pb = 0;
            double pr = 0.0, pi = 0.0;
            for (int q = 0; q < A; ++q) {
                if ((Pb >> q) & 1ULL) { pr += Z.ph_re[q]; pi += Z.ph_im[q]; continue; }
                pb += P.W[q * (A + 1) + __popcll(Pb & ((1ULL << q) - 1ULL))];
            }
            s_base = pb - P.shards.pstart[P.shards.rank];
            s_pre[0] = pr - half_re; s_pre[1] = pi - half_im;          // sum over set bits minus half the full sum = sum_r e^{iqr} s_r
        }
        __syncthreads();
        const SdBlkJs &I = P.js[js];
        const uint64_t base = s_base;
        for (uint32_t p = threadIdx.x; p < I.size_pad; p += blockDim.x) {
            int jt; uint32_t e, u;
            double vr = 0.0, vi = 0.0;
            if (sd_blk_decode(I, 2, p, jt, e, u)) {
                const unsigned cm = P.items[I.cls[jt].item_off + u].c;
                const unsigned tau = sd_tail_cfg(SD_BLK_T, jt, (int)e);
                double sr = s_pre[0], si = s_pre[1];
                for (int q = 0; q < SD_BLK_M; ++q) if ((cm >> q) & 1u) { sr += Z.ph_re[A + q]; si += Z.ph_im[A + q]; }
                for (int q = 0; q < SD_BLK_T; ++q) if ((tau >> q) & 1u) { sr += Z.ph_re[A + SD_BLK_M + q]; si += Z.ph_im[A + SD_BLK_M + q]; }
                sr *= Z.normfact; si *= Z.normfact;
                double pr, pi;
                if (NCIN == 2) { pr = psi_local[(base + p) * 2]; pi = psi_local[(base + p) * 2 + 1]; }
                else { pr = psi_local[base + sd_blk_encode(I, 1, jt, e, u)]; pi = 0.0; }
                vr = sr * pr - si * pi; vi = sr * pi + si * pr;
            }
            *(double2 *)(phi_local + (base + p) * 2) = make_double2(vr, vi);   // padding := 0
            red[3] += vr * vr + vi * vi;
        }
    }
    if (partials) sd_block_reduce_store(red, 8, scratch, partials, nparts, blockIdx.x);
}

// mags / zz partial sums of a block-layout vector (sd_obs.h for the formulas and the partials layout); one CTA per tile,
// grid-stride over the keys, per-warp accumulators written once at the end (fixed tile -> warp map: deterministic).
template <int NC>
__global__ void __launch_bounds__(256) sd_blk_obs_kernel(const __grid_constant__ SdBlkParams P, const double *v_local, double *partials) {
    __shared__ uint64_t s_base;
    const int A = P.A, k = P.k, L = P.L;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    double mag[2] = {0.0, 0.0}, zz[2] = {0.0, 0.0};
    for (uint64_t key = P.key_lo + blockIdx.x; key < P.key_hi; key += gridDim.x) {
        const uint64_t Pb = __brevll(~key) >> (64 - A);
        const int js = k - __popcll(Pb);
        if (js < 0 || js > SD_BLK_B) continue;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t pb = 0;
            for (int q = 0; q < A; ++q)
                if (!((Pb >> q) & 1ULL)) pb += P.W[q * (A + 1) + __popcll(Pb & ((1ULL << q) - 1ULL))];
            s_base = pb - P.shards.pstart[P.shards.rank];
        }
        __syncthreads();
        const SdBlkJs &I = P.js[js];
        const uint64_t base = s_base;
        for (uint32_t p0 = warp * 32u; p0 < I.size_pad; p0 += nwarp * 32u) {
            const uint32_t p = p0 + lane;
            double w = 0.0;
            uint64_t s = 0;
            int jt; uint32_t e, u;
            if (p < I.size_pad && sd_blk_decode(I, NC, p, jt, e, u)) {
                if (NC == 2) { const double re = v_local[(base + p) * 2], im = v_local[(base + p) * 2 + 1]; w = re * re + im * im; }
                else { const double re = v_local[base + p]; w = re * re; }
                s = Pb | ((uint64_t)P.items[I.cls[jt].item_off + u].c << A) | ((uint64_t)sd_tail_cfg(SD_BLK_T, jt, (int)e) << (A + SD_BLK_M));
            }
            for (int j = 0; j < 32; ++j) {
                const double wj = __shfl_sync(0xffffffffu, w, j);
                if (wj == 0.0) continue;                               // Observables.jl:21,58 skip zero weights
                const uint64_t sj = __shfl_sync(0xffffffffu, s, j);
                if ((int)lane < L) sd_obs_accum(L, (int)lane, sj, wj, mag[0], zz[0]);
                if ((int)lane + 32 < L) sd_obs_accum(L, (int)lane + 32, sj, wj, mag[1], zz[1]);
            }
        }
    }
    double *o = partials + ((size_t)blockIdx.x * nwarp + warp) * 128;
    o[lane] = mag[0]; o[32 + lane] = mag[1]; o[64 + lane] = zz[0]; o[96 + lane] = zz[1];
}

#endif
