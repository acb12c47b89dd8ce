// sd_mv.cuh -- q-batched recurrences on an interleaved [state][q] multi-vector.
//
// lanczos_sqw / kpm_sqw (LanczosSqw.jl:65-77, KPM_Sqw.jl:218-253) run the SAME recurrence for every momentum q on
// phi_q = S^z_q psi0 under Threads.@threads.  Here the nq vectors are the columns of one multi-vector, element (r, c)
// (basis rank r, column c) as a complex number at doubles [(r * nqp + c) * 2 .. +1], nqp = nq padded (padding columns are
// zero and stay zero).  One kernel applies H to all columns: the index work of a state (unrank, diagonal, neighbour
// ranks) is done once per QPT columns, every neighbour access is a contiguous run of nqp * 16 bytes, and a Lanczos or
// Chebyshev step of ALL momenta is two kernels (one) with per-column scalars that never leave the device.  That is what
// makes config 1 (L = 16: one apply of one vector fills 2 of 148 SMs) fill the GPU.
//
// Thread map: a state is handled by LPS = nqp / QPT lanes (a power of two <= 32), each owning QPT consecutive columns
// (QPT * 16 contiguous bytes per access).  Any model the generic kernel takes (full / sector basis, arbitrary bond lists).
// Reductions are per column and deterministic: fixed xor-shuffle tree over the lanes that own the same columns, warps in
// order, CTAs in order by the last CTA to finish (threadfence + ticket), so a step needs no separate reduce kernel.
// Single GPU.
#pragma once
#include "sd_common.h"

#define SD_MV_THREADS 256
#define SD_MV_MAXQ 128                  // padded columns per multi-vector
#define SD_MV_NS 2                      // reduction slots per kernel

struct SdMvModel {                      // what the generic kernel needs of a model (SdGenericParams without the shard map)
    int L, k;                           // k = -1: full basis
    int nhop, nzz;
    uint64_t N;
    const int *hop_a, *hop_b;
    const double *hop_J;
    const int *zz_a, *zz_b;
    const double *zz_J;
    const double *field;
    const uint64_t *binom;
    int lin_h;
    const uint64_t *linA, *linB;
};
struct SdMvRed {
    double *partials;                   // [gridDim.x][SD_MV_NS][nqp]
    unsigned *ticket;                   // zero before the first launch; the last CTA resets it
    double *result;                     // [SD_MV_NS][nqp], written by the last CTA; slots a kernel does not use are left alone
};

// per-thread sums red[s][q] (columns chunk * QPT + q) -> result[s * nqp + column]
template <int QPT>
__device__ __forceinline__ void sd_mv_reduce(double (&red)[SD_MV_NS][QPT], int nslots, unsigned lps, unsigned nqp, const SdMvRed &R) {
    __shared__ double scratch[SD_MV_THREADS / 32][SD_MV_NS * SD_MV_MAXQ];
    __shared__ unsigned is_last;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < SD_MV_NS; ++s) {
        if (s >= nslots) break;
#pragma unroll
        for (int q = 0; q < QPT; ++q) {
            double t = red[s][q];
            for (unsigned o = 16; o >= lps; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);   // lanes with equal (lane % lps)
            if (lane < lps) scratch[warp][s * nqp + lane * QPT + q] = t;
        }
    }
    __syncthreads();
    const unsigned nval = (unsigned)nslots * nqp;
    for (unsigned t = threadIdx.x; t < nval; t += blockDim.x) {
        double a = 0.0;
#pragma unroll
        for (int wv = 0; wv < SD_MV_THREADS / 32; ++wv) a += scratch[wv][t];
        R.partials[(size_t)blockIdx.x * (SD_MV_NS * nqp) + t] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicInc(R.ticket, gridDim.x - 1) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (unsigned t = threadIdx.x; t < nval; t += blockDim.x) {
        double a = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) a += __ldcg(R.partials + (size_t)b * (SD_MV_NS * nqp) + t);
        R.result[t] = a;
    }
}

// ------------------------------------------------------------------ phi_c = S^z_{q_c} psi0 for all columns
// Hamiltonian.jl:321-334 per column, same summation order as sd_szq_kernel; result slot 1 = ||phi_c||^2.
struct SdMvSzq {
    int L, k, lps_log2, nqp, ncin;
    uint64_t N;
    double normfact;
    const uint64_t *binom;
    const double *ph;                   // [nqp][L][2]: cos(q r), sin(q r); zero for padding columns
    const double *psi0;                 // rank order, ncin components
    double *phi;
    SdMvRed R;
};
template <int QPT>
__global__ void __launch_bounds__(SD_MV_THREADS) sd_mv_szq_kernel(const __grid_constant__ SdMvSzq A) {
    const unsigned lps = 1u << A.lps_log2, chunk = threadIdx.x & (lps - 1u);
    const uint64_t gt = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    double red[SD_MV_NS][QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) red[0][q] = red[1][q] = 0.0;
    for (uint64_t r = gt >> A.lps_log2; r < A.N; r += nthr >> A.lps_log2) {
        const uint64_t s = A.k < 0 ? r : sd_unrank_state(r, A.L, A.k, A.binom, SD_BINOM_DIM);
        double sr[QPT], si[QPT];
#pragma unroll
        for (int q = 0; q < QPT; ++q) sr[q] = si[q] = 0.0;
        for (int site = 0; site < A.L; ++site) {
            const double sz = ((s >> site) & 1ULL) ? 0.5 : -0.5;
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                const double2 ph = __ldg((const double2 *)A.ph + (size_t)(chunk * QPT + q) * A.L + site);
                sr[q] += ph.x * sz;
                si[q] += ph.y * sz;
            }
        }
        double pr, pi;
        if (A.ncin == 2) { pr = A.psi0[2 * r]; pi = A.psi0[2 * r + 1]; }
        else { pr = A.psi0[r]; pi = 0.0; }
        double2 *o = (double2 *)A.phi + r * (uint64_t)A.nqp + chunk * QPT;
#pragma unroll
        for (int q = 0; q < QPT; ++q) {
            const double a = sr[q] * A.normfact, b = si[q] * A.normfact;
            const double vr = a * pr - b * pi, vi = a * pi + b * pr;
            o[q] = make_double2(vr, vi);
            red[1][q] += vr * vr + vi * vi;
        }
    }
    // slot 0 is unused here; it is reduced (as zeros) so that the result layout is uniform
    sd_mv_reduce<QPT>(red, 2, lps, (unsigned)A.nqp, A.R);
}

// ------------------------------------------------------------------ H on all columns
// EK 1 (Lanczos): w = H u / sqrt(n_prev[c]);            slot 0 = Re <u_c, w_c>                 (Lanczos.jl:217-219)
// EK 2 (KPM v1):  w = (H u - b u) / a;                  slot 0 = Re <u_c, w_c>   (u = phi)     (KPM_Sqw.jl:106-107)
// EK 3 (KPM):     w = 2 (H u - b u) / a - vprev;        slot 0 = Re <phi_c, w_c>, slot 1 = ||w_c||^2   (:111-117); w may be vprev
struct SdMvApply {
    SdMvModel G;
    int lps_log2, nqp;
    const double *u;
    double *w;
    const double *vprev, *phi;
    const double *n_prev;               // EK 1: [nqp]
    double a, b;
    SdMvRed R;
};
template <int QPT, int EK>
__global__ void __launch_bounds__(SD_MV_THREADS) sd_mv_apply_kernel(const __grid_constant__ SdMvApply A) {
    const SdMvModel &G = A.G;
    const unsigned lps = 1u << A.lps_log2, chunk = threadIdx.x & (lps - 1u);
    const uint64_t gt = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    const bool full = G.k < 0;
    double red[SD_MV_NS][QPT], hs[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        red[0][q] = red[1][q] = 0.0;
        hs[q] = 1.0;
        if (EK == 1) { const double n = A.n_prev[chunk * QPT + q]; hs[q] = n > 0.0 ? 1.0 / sqrt(n) : 0.0; }   // zero / padding column stays zero
    }
    const double2 *U = (const double2 *)A.u;
    for (uint64_t r = gt >> A.lps_log2; r < G.N; r += nthr >> A.lps_log2) {
        const uint64_t s = full ? r : sd_unrank_state(r, G.L, G.k, G.binom, SD_BINOM_DIM);
        double diag = 0.0;
        for (int i = 0; i < G.L; ++i) diag += G.field[i] * (((s >> i) & 1ULL) ? 0.5 : -0.5);
        for (int b = 0; b < G.nzz; ++b)
            diag += G.zz_J[b] * (((s >> G.zz_a[b]) & 1ULL) ? 0.5 : -0.5) * (((s >> G.zz_b[b]) & 1ULL) ? 0.5 : -0.5);
        const size_t e0 = (size_t)r * A.nqp + chunk * QPT;
        double2 p[QPT], h[QPT];
#pragma unroll
        for (int q = 0; q < QPT; ++q) { p[q] = U[e0 + q]; h[q] = make_double2(diag * p[q].x, diag * p[q].y); }
        for (int b = 0; b < G.nhop; ++b) {
            const int a = G.hop_a[b], bb = G.hop_b[b];
            const uint64_t ba = (s >> a) & 1ULL, bq = (s >> bb) & 1ULL;
            if (ba == bq) continue;
            uint64_t nr;
            if (full) nr = s ^ (1ULL << a) ^ (1ULL << bb);
            else if (bb == a + 1) {
                const uint64_t d = G.binom[(G.L - 2 - a) * SD_BINOM_DIM + SD_POPC64(a + 2 < 64 ? (s >> (a + 2)) : 0ULL)];
                nr = ba ? r + d : r - d;
            } else {
                const uint64_t ns = s ^ (1ULL << a) ^ (1ULL << bb);
                nr = G.lin_h > 0 ? G.linA[ns & ((1ULL << G.lin_h) - 1)] + G.linB[ns >> G.lin_h]
                                 : sd_rank_state(ns, G.L, G.k, G.binom, SD_BINOM_DIM);
            }
            const double J = G.hop_J[b];
            const double2 *nb = U + (size_t)nr * A.nqp + chunk * QPT;
#pragma unroll
            for (int q = 0; q < QPT; ++q) { const double2 t = nb[q]; h[q].x += J * t.x; h[q].y += J * t.y; }
        }
        double2 *W = (double2 *)A.w + e0;
#pragma unroll
        for (int q = 0; q < QPT; ++q) {
            double2 o;
            if (EK == 1) {
                o = make_double2(hs[q] * h[q].x, hs[q] * h[q].y);
                red[0][q] += p[q].x * o.x + p[q].y * o.y;
            } else {
                o = make_double2((h[q].x - A.b * p[q].x) / A.a, (h[q].y - A.b * p[q].y) / A.a);
                if (EK == 2) red[0][q] += p[q].x * o.x + p[q].y * o.y;
                if (EK == 3) {
                    const double2 vp = ((const double2 *)A.vprev)[e0 + q], f = ((const double2 *)A.phi)[e0 + q];
                    o.x = 2.0 * o.x - vp.x; o.y = 2.0 * o.y - vp.y;
                    red[0][q] += f.x * o.x + f.y * o.y;
                    red[1][q] += o.x * o.x + o.y * o.y;
                }
            }
            W[q] = o;
        }
    }
    sd_mv_reduce<QPT>(red, EK == 3 ? 2 : 1, lps, (unsigned)A.nqp, A.R);
}

// ------------------------------------------------------------------ Lanczos update on all columns (the three-term recurrence of
// sd_lanczos_update_kernel with per-column scalars):  w_c -= (alpha_c / b1_c) u_c + (b1_c / b2_c) uo_c,  slot 1 = ||w_c||^2,
// b1 = sqrt(n1), alpha = d / b1, b2 = sqrt(n2)   (deferred normalisation, see sd_lanczos_engine in sd_api.cu)
struct SdMvUpdate {
    int lps_log2, nqp;
    uint64_t N;
    double *w;
    const double *u, *uo;               // uo null at the first step
    const double *d, *n1, *n2;          // [nqp] each
    SdMvRed R;
};
template <int QPT>
__global__ void __launch_bounds__(SD_MV_THREADS) sd_mv_update_kernel(const __grid_constant__ SdMvUpdate A) {
    const unsigned lps = 1u << A.lps_log2, chunk = threadIdx.x & (lps - 1u);
    const uint64_t gt = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    double red[SD_MV_NS][QPT], ca[QPT], cb[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        red[0][q] = red[1][q] = 0.0;
        const unsigned c = chunk * QPT + q;
        const double b1 = sqrt(A.n1[c]), alpha = A.d[c] / b1;
        ca[q] = b1 > 0.0 ? alpha / b1 : 0.0;                        // zero / padding column stays zero
        cb[q] = (A.uo && b1 > 0.0) ? b1 / sqrt(A.n2[c]) : 0.0;
    }
    for (uint64_t r = gt >> A.lps_log2; r < A.N; r += nthr >> A.lps_log2) {
        const size_t e0 = (size_t)r * A.nqp + chunk * QPT;
        double2 *W = (double2 *)A.w + e0;
        const double2 *U = (const double2 *)A.u + e0;
#pragma unroll
        for (int q = 0; q < QPT; ++q) {
            double2 x = W[q];
            const double2 y = U[q];
            x.x -= ca[q] * y.x; x.y -= ca[q] * y.y;
            if (A.uo) { const double2 z = ((const double2 *)A.uo)[e0 + q]; x.x -= cb[q] * z.x; x.y -= cb[q] * z.y; }
            W[q] = x;
            red[1][q] += x.x * x.x + x.y * x.y;
        }
    }
    // one slot: R.result points at the ||w_c||^2 row of this step (the host keeps the apply's dot in the row before it)
    double r1[SD_MV_NS][QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) { r1[0][q] = red[1][q]; r1[1][q] = 0.0; }
    sd_mv_reduce<QPT>(r1, 1, lps, (unsigned)A.nqp, A.R);
}

// ------------------------------------------------------------------ phi_c /= sqrt(n0[c])  (KPM_Sqw.jl:231), slot 0 = ||phi_c||^2 = mu_0
// A zero column (norm(phi) == 0, :226-229) stays zero.
struct SdMvScale {
    int lps_log2, nqp;
    uint64_t N;
    double *phi;
    const double *n0;
    SdMvRed R;
};
template <int QPT>
__global__ void __launch_bounds__(SD_MV_THREADS) sd_mv_colnorm_kernel(const __grid_constant__ SdMvScale A) {
    const unsigned lps = 1u << A.lps_log2, chunk = threadIdx.x & (lps - 1u);
    const uint64_t gt = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    double red[SD_MV_NS][QPT], nrm[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) { red[0][q] = red[1][q] = 0.0; nrm[q] = sqrt(A.n0[chunk * QPT + q]); }
    for (uint64_t r = gt >> A.lps_log2; r < A.N; r += nthr >> A.lps_log2) {
        double2 *P = (double2 *)A.phi + (size_t)r * A.nqp + chunk * QPT;
#pragma unroll
        for (int q = 0; q < QPT; ++q) {
            double2 x = P[q];
            if (nrm[q] > 0.0) { x.x /= nrm[q]; x.y /= nrm[q]; }
            P[q] = x;
            red[0][q] += x.x * x.x + x.y * x.y;
        }
    }
    sd_mv_reduce<QPT>(red, 1, lps, (unsigned)A.nqp, A.R);
}
