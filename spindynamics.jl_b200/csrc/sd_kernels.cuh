// sd_kernels.cuh -- every __global__ kernel of libspindyn_cuda (sm_100a).
// Tensor cores are not used: nothing on this path is a dense contraction; the
// kernels are HBM/L2/shared-memory gathers with integer index algebra.
#pragma once
#include <cuda_runtime.h>
#include "sd_common.h"
#include "sd_tile.h"

#define SD_TILE_THREADS 512
#define SD_GEN_THREADS 256
#define SD_BLAS_THREADS 256
#define SD_FULL_MASK 0xffffffffu

// ------------------------------------------------------------ reductions
// Deterministic: fixed shuffle tree inside a warp, warps summed in order by
// thread 0, CTAs summed in a fixed-shape tree by sd_reduce_partials_kernel
// (test_Lanczos.jl:122-166 demands run-to-run identical results).
__device__ __forceinline__ double sd_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(SD_FULL_MASK, v, o);
    return v;
}

// red[] per thread -> partials[slot*nparts + part]; scratch is [SD_NSLOT][32] smem.
__device__ __forceinline__ void sd_block_reduce_store(double (&red)[SD_NSLOT], int mask,
                                                      double (*scratch)[16], double *partials,
                                                      unsigned nparts, unsigned part) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31u) >> 5;
#pragma unroll
    for (int s = 0; s < SD_NSLOT; ++s) {
        if (!((mask >> s) & 1)) continue;
        const double w = sd_warp_sum(red[s]);
        if (lane == 0) scratch[s][warp] = w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < SD_NSLOT; ++s) {
            if (!((mask >> s) & 1)) continue;
            double t = 0.0;
            for (unsigned w = 0; w < nwarp; ++w) t += scratch[s][w];
            partials[(size_t)s * nparts + part] = t;
        }
    }
}

// slot mask of an epilogue's reductions (dot_self uses slots 0 and 1)
__host__ __device__ __forceinline__ int sd_epi_slotmask(int red) {
    int m = 0;
    if (red & SD_RED_DOT_SELF) m |= 3;
    if (red & SD_RED_DOT_PHI) m |= 4;
    if (red & SD_RED_NORM2) m |= 8;
    return m;
}

// One CTA: result[s] = sum_i partials[s*nparts + i], fixed order.
__global__ void __launch_bounds__(1024) sd_reduce_partials_kernel(const double *partials, unsigned nparts,
                                                                   int slotmask, double *result) {
    __shared__ double sh[32];
    for (int s = 0; s < SD_NSLOT; ++s) {
        if (!((slotmask >> s) & 1)) { if (threadIdx.x == 0) result[s] = 0.0; continue; }
        double t = 0.0;
        for (unsigned i = threadIdx.x; i < nparts; i += blockDim.x) t += partials[(size_t)s * nparts + i];
        t = sd_warp_sum(t);
        if ((threadIdx.x & 31u) == 0) sh[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x < 32) {
            double u = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
            u = sd_warp_sum(u);
            if (threadIdx.x == 0) result[s] = u;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------ tiled apply
// Optional per-phase cycle counters (thread 0 of every CTA), enabled with -DSD_PHASE_TIMING.
#ifdef SD_PHASE_TIMING
#define SD_TICK(i)                                                          \
    do {                                                                    \
        if (threadIdx.x == 0) {                                             \
            const long long now_ = clock64();                               \
            atomicAdd(&sd_phase_cycles[i], (unsigned long long)(now_ - tick_)); \
            tick_ = now_;                                                   \
        }                                                                   \
    } while (0)
#define SD_TICK_INIT() long long tick_ = clock64()
#else
#define SD_TICK(i) ((void)0)
#define SD_TICK_INIT() ((void)0)
#endif

// registers: own[] + acc[] take 4*C(T,T/2)*NC; two CTAs per SM only when that fits 64 regs/thread.
// PLAIN = no fused epilogue (out = H psi): phase 3 is a pure streaming store.
template <int NC, int T, bool PLAIN, int NTHR>
__global__ void __launch_bounds__(NTHR, ((4 * sd_cbinom(T, T / 2) * NC <= 48) ? 1024 : 512) / NTHR)
sd_tile_apply_kernel(const __grid_constant__ SdTileParams P, const __grid_constant__ SdVecView psi,
                     double *out_vbase, const __grid_constant__ SdEpi epi, uint32_t cap) {
    extern __shared__ __align__(16) unsigned char sd_smem[];
    // One CTA per tile (a persistent variant was measured slower: all CTAs fall into lockstep and
    // the memory system idles during the compute phases).  The CTA size is a compile-time constant
    // so element strides become immediates after inlining.
    const SdTileView<NC> v = sd_tile_carve<NC>(sd_smem, cap);
    SdTileHdr &H = *v.hdr;
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint64_t key = P.key_lo + blockIdx.x;
    SD_TICK_INIT();
    if (warp == 0) {
        const SdHdrRegs r = sd_tile_hdr_issue(P, key, lane);
        sd_tile_hdr_finish<NC>(P, key, H, psi, lane, r);
    } else if (warp == 1 && P.pf_dist > 0 && key + (uint64_t)P.pf_dist < P.key_hi) {
        // L2 prefetch of the own tile and far neighbour tiles of the tile pf_dist keys ahead
        SdTileHdr &Hp = *(SdTileHdr *)((char *)v.hdr + ((sizeof(SdTileHdr) + 15) & ~(size_t)15));
        const uint64_t key2 = key + (uint64_t)P.pf_dist;
        const SdHdrRegs r = sd_tile_hdr_issue(P, key2, lane);
        sd_tile_hdr_finish<NC>(P, key2, Hp, psi, lane, r);
        __syncwarp();
        sd_tile_hdr_prefetch<NC>(P, Hp, psi, lane);
    }
    for (int i = (int)tid; i < (P.M + 1) * (P.M + 1); i += NTHR) v.binomM[i] = P.binomM[i];
    // the first phase-2 item of this thread depends on js and tid only: prefetch it now
    SdItem item0;
    item0.c = 0xFFFFu;
    {
        const int js = P.k - __popcll(sd_tile_prefix_bits(key, P.A));
        if (js >= 0 && js <= P.B && tid < P.js[js].nslots) item0 = P.items[P.js[js].item_off + tid];
    }
    SD_TICK(0);
    __syncthreads();
    SD_TICK(1);
    const int slotmask = PLAIN ? 0 : sd_epi_slotmask(epi.red);
    double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
    if (H.valid) {
        sd_tile_phase1<NC>(P, v, psi, tid, NTHR);
        SD_TICK(2);                                       // thread 0's own phase-1 time
        __syncthreads();
        SD_TICK(3);                                       // wait for the slowest warp
        sd_tile_phase2<NC, T>(P, v, tid, NTHR, item0);
        SD_TICK(4);
        __syncthreads();
        SD_TICK(5);
        sd_tile_phase3<NC, PLAIN>(P, v, out_vbase, epi, tid, NTHR, red);
        SD_TICK(6);
    }
    if (!PLAIN && slotmask) sd_block_reduce_store(red, slotmask, H.red, epi.partials, epi.nparts, blockIdx.x);
}

// ------------------------------------------------------------ generic apply
// One thread per output state, arbitrary bond lists, full or sector basis;
// same term order as the reference loop (field, zz list, hop list).
struct SdGenericParams {
    int L, k;                         // k = -1: full basis
    int nhop, nzz;
    uint64_t N;
    const int *hop_a, *hop_b;         // 0-based positions, a < b
    const double *hop_J;
    const int *zz_a, *zz_b;
    const double *zz_J;
    const double *field;
    const uint64_t *binom;            // [65*65]
    int lin_h;                        // rank = linA[s & (2^h-1)] + linB[s >> h]; 0 = tables absent
    const uint64_t *linA, *linB;
    SdShardMap shards;
};

__device__ __forceinline__ uint64_t sd_dev_rank(const SdGenericParams &G, uint64_t s) {
    if (G.lin_h > 0) return G.linA[s & ((1ULL << G.lin_h) - 1)] + G.linB[s >> G.lin_h];
    return sd_rank_state(s, G.L, G.k, G.binom, SD_BINOM_DIM);
}

template <int NC>
__global__ void __launch_bounds__(SD_GEN_THREADS)
sd_generic_apply_kernel(const __grid_constant__ SdGenericParams G, const __grid_constant__ SdVecView psi,
                        double *out_local, const __grid_constant__ SdEpi epi) {
    __shared__ double scratch[SD_NSLOT][16];
    const uint64_t lstart = G.shards.start[G.shards.rank];
    const uint64_t ln = G.shards.start[G.shards.rank + 1] - lstart;
    const bool full = G.k < 0;
    const bool single = G.shards.world == 1;
    double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
    for (uint64_t li = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; li < ln;
         li += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = lstart + li;
        const uint64_t s = full ? r : sd_unrank_state(r, G.L, G.k, G.binom, SD_BINOM_DIM);
        double diag = 0.0;
        for (int i = 0; i < G.L; ++i) diag += G.field[i] * (((s >> i) & 1ULL) ? 0.5 : -0.5);
        for (int b = 0; b < G.nzz; ++b)
            diag += G.zz_J[b] * (((s >> G.zz_a[b]) & 1ULL) ? 0.5 : -0.5) * (((s >> G.zz_b[b]) & 1ULL) ? 0.5 : -0.5);
        SdVal<NC> p, h;
        const double *pl = psi.base[G.shards.rank] + (size_t)NC * r;
#pragma unroll
        for (int c = 0; c < NC; ++c) { p.c[c] = pl[c]; h.c[c] = diag * p.c[c]; }
        for (int b = 0; b < G.nhop; ++b) {
            const int a = G.hop_a[b], bb = G.hop_b[b];
            const uint64_t ba = (s >> a) & 1ULL, bq = (s >> bb) & 1ULL;
            if (ba != bq) {
                const uint64_t ns = s ^ (1ULL << a) ^ (1ULL << bb);
                uint64_t nr;
                if (full) nr = ns;
                else if (bb == a + 1) {
                    const uint64_t d = G.binom[(G.L - 2 - a) * SD_BINOM_DIM + SD_POPC64(a + 2 < 64 ? (s >> (a + 2)) : 0ULL)];
                    nr = ba ? r + d : r - d;
                } else nr = sd_dev_rank(G, ns);
                const double *q = psi.base[single ? 0 : sd_owner(G.shards, nr)] + (size_t)NC * nr;
                const double J = G.hop_J[b];
#pragma unroll
                for (int c = 0; c < NC; ++c) h.c[c] += J * q[c];
            }
        }
        const SdVal<NC> o = sd_epilogue<NC>(epi, h, p, li, red);
#pragma unroll
        for (int c = 0; c < NC; ++c) out_local[(size_t)li * NC + c] = o.c[c];
    }
    const int slotmask = sd_epi_slotmask(epi.red);
    if (slotmask) sd_block_reduce_store(red, slotmask, scratch, epi.partials, epi.nparts, blockIdx.x);
}

// ------------------------------------------------------------ basis kernels
__global__ void sd_unrank_kernel(int L, int k, const uint64_t *binom, uint64_t first, uint64_t count,
                                 uint64_t *states) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (uint64_t)gridDim.x * blockDim.x)
        states[i] = (k < 0) ? first + i : sd_unrank_state(first + i, L, k, binom, SD_BINOM_DIM);
}

// idx1 = 1-based rank, 0 if the state is outside the basis (get(idxmap, s, 0))
__global__ void sd_rank_kernel(int L, int k, const uint64_t *binom, const uint64_t *states, uint64_t count,
                               int64_t *idx1) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = states[i];
        int64_t r;
        if (L < 64 && (s >> L)) r = 0;
        else if (k < 0) r = (int64_t)s + 1;
        else if (SD_POPC64(s) != k) r = 0;
        else r = (int64_t)sd_rank_state(s, L, k, binom, SD_BINOM_DIM) + 1;
        idx1[i] = r;
    }
}

// ------------------------------------------------------------ Sz_q_vector
struct SdSzqParams {
    int L, k;
    double normfact;
    double ph_re[SD_MAX_L + 1], ph_im[SD_MAX_L + 1];
    const uint64_t *binom;
};
// phi[idx] = L^-1/2 (sum_r e^{iqr} s_r(state)) * ComplexF64(psi0[idx])   Hamiltonian.jl:321-334
template <int NCIN>
__global__ void __launch_bounds__(SD_BLAS_THREADS)
sd_szq_kernel(const __grid_constant__ SdSzqParams Z, uint64_t lstart, uint64_t ln, const double *psi0,
              double *phi, double *partials, unsigned nparts) {
    __shared__ double scratch[SD_NSLOT][16];
    double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
    for (uint64_t li = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; li < ln;
         li += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = lstart + li;
        const uint64_t s = (Z.k < 0) ? r : sd_unrank_state(r, Z.L, Z.k, Z.binom, SD_BINOM_DIM);
        double sr = 0.0, si = 0.0;
        for (int q = 0; q < Z.L; ++q) {
            const double sz = ((s >> q) & 1ULL) ? 0.5 : -0.5;
            sr += Z.ph_re[q] * sz;
            si += Z.ph_im[q] * sz;
        }
        sr *= Z.normfact; si *= Z.normfact;
        double pr, pi;
        if (NCIN == 2) { pr = psi0[li * 2]; pi = psi0[li * 2 + 1]; }
        else { pr = psi0[li]; pi = 0.0; }
        const double vr = sr * pr - si * pi, vi = sr * pi + si * pr;
        phi[li * 2] = vr; phi[li * 2 + 1] = vi;
        red[3] += vr * vr + vi * vi;
    }
    if (partials) sd_block_reduce_store(red, 8, scratch, partials, nparts, blockIdx.x);
}

// ------------------------------------------------------------ BLAS-1
__global__ void sd_fill_seeded_kernel(double *v, int nc, uint64_t first, uint64_t n, uint64_t seed, double scale) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        if (nc == 2) {
            v[2 * i] = scale * sd_seeded_value(seed, first + i);
            v[2 * i + 1] = scale * sd_seeded_value(seed + 1, first + i);
        } else v[i] = scale * sd_seeded_value(seed, first + i);
    }
}

__global__ void sd_set_one_kernel(double *v, uint64_t off) { v[off] = 1.0; }
// out[i*nc + c] = v[off[i]*nc + c]  (sd_vec_get: a handful of elements by stored offset)
__global__ void sd_gather_kernel(const double *v, const uint64_t *off, unsigned n, int nc, double *out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) for (int c = 0; c < nc; ++c) out[(size_t)i * nc + c] = v[off[i] * nc + c];
}

// y = ComplexF64.(x) (nc_in=1 -> nc_out=2) or real part (2 -> 1)
__global__ void sd_convert_kernel(double *y, const double *x, uint64_t n, int nc_in) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        if (nc_in == 1) { y[2 * i] = x[i]; y[2 * i + 1] = 0.0; }
        else y[i] = x[2 * i];
    }
}

// Scalars may come from the host (by value) or from device memory (dev != null:
// s = dev[0] + i dev[1], then multiplied by (fr, fi)), so recurrences can chain
// kernels without a host round trip.
struct SdScalar {
    double re, im;
    const double *dev;
    int dev_mode;    // 0: (dev[0], dev[1]);  1: (dev[0], 0);  2: (1/sqrt(dev[0]), 0);  3: (-dev[0], -dev[1]); 4: (-dev[0], 0)
};
__device__ __forceinline__ void sd_load_scalar(const SdScalar &s, double &re, double &im) {
    if (!s.dev) { re = s.re; im = s.im; return; }
    const double a = s.dev[0], b = s.dev[1];
    switch (s.dev_mode) {
        case 0: re = a; im = b; break;
        case 1: re = a; im = 0.0; break;
        case 2: re = 1.0 / sqrt(a); im = 0.0; break;
        case 3: re = -a; im = -b; break;
        default: re = -a; im = 0.0; break;
    }
}

// ---- streaming BLAS-1 kernels.  All of them move 16-byte units (double2: one c128 element or two f64 elements),
// four units per thread and loop trip with the loads issued before the first store, grid-stride over the vector; an
// odd f64 tail element is handled by one thread.  n = elements, units = n * NC / 2.
#define SD_BLAS_UNROLL 4
__device__ __forceinline__ double2 sd_ld2(const double *p, uint64_t unit) { return *(const double2 *)(p + 2 * unit); }
__device__ __forceinline__ void sd_st2(double *p, uint64_t unit, double2 v) { *(double2 *)(p + 2 * unit) = v; }

// x *= s
template <int NC>
__global__ void __launch_bounds__(SD_BLAS_THREADS) sd_scale_kernel(double *x, uint64_t n, const __grid_constant__ SdScalar s) {
    double sr, si;
    sd_load_scalar(s, sr, si);
    const uint64_t units = n * NC / 2, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < units; i0 += SD_BLAS_UNROLL * stride) {
        double2 v[SD_BLAS_UNROLL];
#pragma unroll
        for (int k = 0; k < SD_BLAS_UNROLL; ++k) if (i0 + k * stride < units) v[k] = sd_ld2(x, i0 + k * stride);
#pragma unroll
        for (int k = 0; k < SD_BLAS_UNROLL; ++k) {
            if (i0 + k * stride >= units) continue;
            const double2 a = v[k];
            sd_st2(x, i0 + k * stride, NC == 2 ? make_double2(sr * a.x - si * a.y, sr * a.y + si * a.x) : make_double2(sr * a.x, sr * a.y));
        }
    }
    if (NC == 1 && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) x[n - 1] *= sr;
}
// y = x / s (true division by a real scalar, as `w / beta` in Lanczos.jl:71,155,233)
template <int NC>
__global__ void __launch_bounds__(SD_BLAS_THREADS) sd_divide_kernel(double *y, const double *x, uint64_t n, const __grid_constant__ SdScalar s) {
    double sr, si;
    sd_load_scalar(s, sr, si);
    const uint64_t units = n * NC / 2, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < units; i0 += SD_BLAS_UNROLL * stride) {
        double2 v[SD_BLAS_UNROLL];
#pragma unroll
        for (int k = 0; k < SD_BLAS_UNROLL; ++k) if (i0 + k * stride < units) v[k] = sd_ld2(x, i0 + k * stride);
#pragma unroll
        for (int k = 0; k < SD_BLAS_UNROLL; ++k)
            if (i0 + k * stride < units) sd_st2(y, i0 + k * stride, make_double2(v[k].x / sr, v[k].y / sr));
    }
    if (NC == 1 && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) y[n - 1] = x[n - 1] / sr;
}

// y += a x  [+ b z]; optionally ||y||^2 (slot 3)
template <int NC>
__global__ void __launch_bounds__(SD_BLAS_THREADS)
sd_axpy_kernel(double *y, uint64_t n, const __grid_constant__ SdScalar a, const double *x,
               const __grid_constant__ SdScalar b, const double *z, double *partials, unsigned nparts) {
    __shared__ double scratch[SD_NSLOT][16];
    double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
    double ar, ai, br = 0.0, bi = 0.0;
    sd_load_scalar(a, ar, ai);
    if (z) sd_load_scalar(b, br, bi);
    const uint64_t units = n * NC / 2, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < units; i0 += SD_BLAS_UNROLL * stride) {
        double2 vy[SD_BLAS_UNROLL], vx[SD_BLAS_UNROLL], vz[SD_BLAS_UNROLL];
#pragma unroll
        for (int k = 0; k < SD_BLAS_UNROLL; ++k) {
            const uint64_t i = i0 + k * stride;
            vy[k] = vx[k] = vz[k] = make_double2(0.0, 0.0);
            if (i >= units) continue;
            vy[k] = sd_ld2(y, i); vx[k] = sd_ld2(x, i);
            if (z) vz[k] = sd_ld2(z, i);
        }
#pragma unroll
        for (int k = 0; k < SD_BLAS_UNROLL; ++k) {
            const uint64_t i = i0 + k * stride;
            if (i >= units) continue;
            double2 r = vy[k];
            if (NC == 2) {
                r.x += ar * vx[k].x - ai * vx[k].y; r.y += ar * vx[k].y + ai * vx[k].x;
                if (z) { r.x += br * vz[k].x - bi * vz[k].y; r.y += br * vz[k].y + bi * vz[k].x; }
            } else {
                r.x += ar * vx[k].x; r.y += ar * vx[k].y;
                if (z) { r.x += br * vz[k].x; r.y += br * vz[k].y; }
            }
            sd_st2(y, i, r);
            red[3] += r.x * r.x + r.y * r.y;
        }
    }
    if (NC == 1 && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double r = y[n - 1] + ar * x[n - 1];
        if (z) r += br * z[n - 1];
        y[n - 1] = r;
        red[3] += r * r;
    }
    if (partials) sd_block_reduce_store(red, 8, scratch, partials, nparts, blockIdx.x);
}

// dot(x, y) = sum conj(x) y (conj != 0) or sum x y (conj == 0): slots 0,1
template <int NC>
__global__ void __launch_bounds__(SD_BLAS_THREADS)
sd_dot_kernel(const double *x, const double *y, uint64_t n, int conj, double *partials, unsigned nparts) {
    __shared__ double scratch[SD_NSLOT][16];
    double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
    const uint64_t units = n * NC / 2, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < units; i0 += SD_BLAS_UNROLL * stride) {
        double2 vx[SD_BLAS_UNROLL], vy[SD_BLAS_UNROLL];
#pragma unroll
        for (int k = 0; k < SD_BLAS_UNROLL; ++k) {
            vx[k] = vy[k] = make_double2(0.0, 0.0);
            if (i0 + k * stride < units) { vx[k] = sd_ld2(x, i0 + k * stride); vy[k] = sd_ld2(y, i0 + k * stride); }
        }
#pragma unroll
        for (int k = 0; k < SD_BLAS_UNROLL; ++k) {
            if (NC == 2) {
                const double xi = conj ? vx[k].y : -vx[k].y;
                red[0] += vx[k].x * vy[k].x + xi * vy[k].y;
                red[1] += vx[k].x * vy[k].y - xi * vy[k].x;
            } else red[0] += vx[k].x * vy[k].x + vx[k].y * vy[k].y;
        }
    }
    if (NC == 1 && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) red[0] += x[n - 1] * y[n - 1];
    sd_block_reduce_store(red, 3, scratch, partials, nparts, blockIdx.x);
}

// Fused tail of a Lanczos step with deferred normalisation (Lanczos.jl:52-65,124-155,219-233 in one 3R+1W pass, no
// `w / beta` pass).  The vectors are kept UNNORMALISED: u_1 = v0, u_{j+1} = w_j, v_j = u_j / beta_{j-1} with
// beta_0 = ||v0||.  The apply before this kernel wrote  w = H v_j = (1 / beta_{j-1}) H u_j  (epilogue hscale) and
// d = <u_j, w>; here
//     alpha_j = d / beta_{j-1},   w -= (alpha_j / beta_{j-1}) u_j + (beta_{j-1} / beta_{j-2}) u_{j-1},   n_j = ||w||^2
// and, for the second pass of the memory-lean ground state, out += (y_j / beta_{j-1}) u_j.
// beta = sqrt(n), the divisions and the sqrt are IEEE operations, so a host that recomputes alpha and beta from the
// fetched (d, n) gets the same bits and pass 2 (host scalars) regenerates the vectors of pass 1 (device scalars) exactly.
struct SdLanczosScal {
    const double *d_dev, *n1_dev, *n2_dev;   // device: <u_j, w>, ||u_j||^2 = beta_{j-1}^2, ||u_{j-1}||^2; d_dev null -> host values:
    double alpha, b1, b2;                     // alpha_j, beta_{j-1}, beta_{j-2}
    double yj;                                // pass 2: Ritz coefficient of v_j (0: no accumulation)
};
template <int NC>
__global__ void __launch_bounds__(SD_BLAS_THREADS)
sd_lanczos_update_kernel(double *w, const double *u, const double *uo, double *out, uint64_t n,
                         const __grid_constant__ SdLanczosScal S, double *partials, unsigned nparts) {
    __shared__ double scratch[SD_NSLOT][16];
    double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
    double alpha = S.alpha, b1 = S.b1, b2 = S.b2;
    if (S.d_dev) {
        b1 = sqrt(*S.n1_dev);
        alpha = *S.d_dev / b1;
        if (uo) b2 = sqrt(*S.n2_dev);
    }
    const double ca = alpha / b1, cb = uo ? b1 / b2 : 0.0, co = S.yj / b1;
    const uint64_t units = n * NC / 2, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < units; i0 += SD_BLAS_UNROLL * stride) {
        double2 vw[SD_BLAS_UNROLL], vu[SD_BLAS_UNROLL], vo[SD_BLAS_UNROLL], va[SD_BLAS_UNROLL];
#pragma unroll
        for (int k = 0; k < SD_BLAS_UNROLL; ++k) {
            const uint64_t i = i0 + k * stride;
            vw[k] = vu[k] = vo[k] = va[k] = make_double2(0.0, 0.0);
            if (i >= units) continue;
            vw[k] = sd_ld2(w, i); vu[k] = sd_ld2(u, i);
            if (uo) vo[k] = sd_ld2(uo, i);
            if (out) va[k] = sd_ld2(out, i);
        }
#pragma unroll
        for (int k = 0; k < SD_BLAS_UNROLL; ++k) {
            const uint64_t i = i0 + k * stride;
            if (i >= units) continue;
            double2 r = vw[k];
            r.x -= ca * vu[k].x; r.y -= ca * vu[k].y;
            if (uo) { r.x -= cb * vo[k].x; r.y -= cb * vo[k].y; }
            sd_st2(w, i, r);
            red[3] += r.x * r.x + r.y * r.y;
            if (out) sd_st2(out, i, make_double2(va[k].x + co * vu[k].x, va[k].y + co * vu[k].y));
        }
    }
    if (NC == 1 && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double r = w[n - 1] - ca * u[n - 1];
        if (uo) r -= cb * uo[n - 1];
        w[n - 1] = r;
        red[3] += r * r;
        if (out) out[n - 1] += co * u[n - 1];
    }
    sd_block_reduce_store(red, 8, scratch, partials, nparts, blockIdx.x);
}


// block dot: c[j] = dot(V_j, w) for j < m (full reorthogonalisation, Lanczos.jl:116-122
// computes these one at a time; the block form reads w once).  partials: [m][nparts].
#define SD_BDOT_MAX 8
struct SdPtrBlock { const double *v[SD_BDOT_MAX]; };
template <int NC>
__global__ void __launch_bounds__(SD_BLAS_THREADS)
sd_lincomb_kernel(double *out, uint64_t n, const __grid_constant__ SdPtrBlock V, int m, const double *y /*dev, (re,im) pairs*/,
                  int accumulate, double *partials, unsigned nparts) {
    __shared__ double scratch[SD_NSLOT][16];
    double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        if (NC == 2) {
            double orr = accumulate ? out[2 * i] : 0.0, oi = accumulate ? out[2 * i + 1] : 0.0;
            for (int j = 0; j < m; ++j) {
                const double yr = y[2 * j], yi = y[2 * j + 1];
                const double vr = V.v[j][2 * i], vi = V.v[j][2 * i + 1];
                orr += yr * vr - yi * vi; oi += yr * vi + yi * vr;
            }
            out[2 * i] = orr; out[2 * i + 1] = oi;
            red[3] += orr * orr + oi * oi;
        } else {
            double o = accumulate ? out[i] : 0.0;
            for (int j = 0; j < m; ++j) o += y[2 * j] * V.v[j][i];
            out[i] = o;
            red[3] += o * o;
        }
    }
    if (partials) sd_block_reduce_store(red, 8, scratch, partials, nparts, blockIdx.x);
}

// out(C128) = sum_j y_j * V_j with REAL V_j (Krylov.jl:185-188 with T = Float64)
__global__ void __launch_bounds__(SD_BLAS_THREADS)
sd_lincomb_r2c_kernel(double *out, uint64_t n, const __grid_constant__ SdPtrBlock V, int m, const double *y,
                      int accumulate, double *partials, unsigned nparts) {
    __shared__ double scratch[SD_NSLOT][16];
    double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        double orr = accumulate ? out[2 * i] : 0.0, oi = accumulate ? out[2 * i + 1] : 0.0;
        for (int j = 0; j < m; ++j) {
            const double v = V.v[j][i];
            orr += y[2 * j] * v; oi += y[2 * j + 1] * v;
        }
        out[2 * i] = orr; out[2 * i + 1] = oi;
        red[3] += orr * orr + oi * oi;
    }
    if (partials) sd_block_reduce_store(red, 8, scratch, partials, nparts, blockIdx.x);
}
