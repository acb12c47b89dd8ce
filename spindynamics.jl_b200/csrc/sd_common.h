// sd_common.h -- host/device helpers shared by every kernel of libspindyn_cuda.
//
// Basis convention (reference Basis.jl:37-53): site i (1-based) is bit i-1 of the
// state; the sector basis is ordered like `combinations(1:L, nup)`, i.e.
// lexicographically in the ascending site lists.  Reading a state from bit 0
// upwards, "1 comes before 0" and bit 0 is the most significant digit of the
// rank.  With rem = number of set bits strictly above position q,
//     rank(s) = sum over CLEAR bits q (rem > 0) of C(L-1-q, rem-1)
// and a nearest-neighbour hop of bond p (bits p, p+1) moves the rank by
//     delta = C(L-2-p, popc(s >> (p+2)))        ((1,0) -> (0,1): +delta).
// Nothing here stores states[] or a Dict (Basis.jl:49-52 is replaced).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define SD_HD __host__ __device__ __forceinline__
#define SD_HDC __host__ __device__ constexpr
#else
#define SD_HD inline
#define SD_HDC constexpr
#endif

#if defined(__CUDA_ARCH__)
#define SD_POPC32(x) __popc((unsigned)(x))
#define SD_POPC64(x) __popcll((unsigned long long)(x))
#else
#define SD_POPC32(x) __builtin_popcount((unsigned)(x))
#define SD_POPC64(x) __builtin_popcountll((unsigned long long)(x))
#endif

#define SD_MAX_L 63
#define SD_BINOM_DIM 65           // rows/cols of the binomial table
#define SD_MAX_WORLD 8

// Binomial table C(n, r), n, r < 65, row-major [n*65 + r]; 0 outside 0<=r<=n.
// C(64,32) < 2^64, so u64 holds every entry.
static inline void sd_fill_binom(uint64_t *c) {
    for (int n = 0; n < SD_BINOM_DIM; ++n)
        for (int r = 0; r < SD_BINOM_DIM; ++r) {
            uint64_t v;
            if (r > n) v = 0;
            else if (r == 0 || r == n) v = 1;
            else v = c[(n - 1) * SD_BINOM_DIM + r - 1] + c[(n - 1) * SD_BINOM_DIM + r];
            c[n * SD_BINOM_DIM + r] = v;
        }
}

// C(n, r) with r possibly -1 (-> 0).  `tab` is any [n*stride + r] table.
template <typename TabT>
SD_HD uint64_t sd_binom_at(const TabT *tab, int stride, int n, int r) {
    return (r < 0 || r > n) ? 0 : (uint64_t)tab[n * stride + r];
}

// idx0 -> state  (the idx-th element of build_sector_basis, 0-based)
template <typename TabT>
SD_HD uint64_t sd_unrank_state(uint64_t idx, int L, int k, const TabT *tab, int stride) {
    uint64_t s = 0;
    int r = k;
    for (int p = 0; p < L && r > 0; ++p) {
        uint64_t c = (uint64_t)tab[(L - 1 - p) * stride + (r - 1)];
        if (idx < c) { s |= 1ULL << p; --r; }
        else idx -= c;
    }
    return s;
}

// state -> idx0; caller guarantees popcount(s) == k and s < 2^L.
template <typename TabT>
SD_HD uint64_t sd_rank_state(uint64_t s, int L, int k, const TabT *tab, int stride) {
    uint64_t idx = 0;
    int r = k;
    for (int p = 0; p < L && r > 0; ++p) {
        if ((s >> p) & 1ULL) --r;
        else idx += (uint64_t)tab[(L - 1 - p) * stride + (r - 1)];
    }
    return idx;
}

// ---- compile-time tables for the register-resident tail block ------------
SD_HDC int sd_cbinom(int n, int r) {
    if (r < 0 || r > n) return 0;
    int v = 1;
    for (int i = 1; i <= r; ++i) v = v * (n - r + i) / i;
    return v;
}
// t-th configuration of T sites with J set bits, same "1 first" lexicographic order
SD_HDC unsigned sd_tail_cfg(int T, int J, int t) {
    unsigned s = 0;
    int r = J, idx = t;
    for (int p = 0; p < T && r > 0; ++p) {
        int c = sd_cbinom(T - 1 - p, r - 1);
        if (idx < c) { s |= 1u << p; --r; }
        else idx -= c;
    }
    return s;
}
SD_HDC int sd_tail_rank(int T, int J, unsigned s) {
    int idx = 0, r = J;
    for (int p = 0; p < T && r > 0; ++p) {
        if ((s >> p) & 1u) --r;
        else idx += sd_cbinom(T - 1 - p, r - 1);
    }
    return idx;
}

// ---- counter-based synthetic psi; oracle.c:seeded_value ---
SD_HD double sd_seeded_value(uint64_t seed, uint64_t r) {
    uint64_t z = (seed ^ r) + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    z ^= z >> 31;
    return 2.0 * ((double)(z >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}

// ---- sharded vectors -------------------------------------------------------
// A vector is split by contiguous rank range over `world` GPUs.  base[g] is a
// VIRTUAL base pointer: element r (global rank) of component c lives at
// base[g][r*NC + c] for start[g] <= r < start[g+1].  base[rank] is local HBM,
// the others are CUDA-IPC mappings of peer HBM (plain ld.global over NVLink).
struct SdShardMap {
    int world, rank;
    uint64_t start[SD_MAX_WORLD + 1];
};
struct SdVecView {
    const double *base[SD_MAX_WORLD];
};
SD_HD int sd_owner(const SdShardMap &m, uint64_t r) {
    int g = 0;
#pragma unroll
    for (int i = 1; i < SD_MAX_WORLD; ++i) g += (i < m.world && r >= m.start[i]) ? 1 : 0;
    return g;
}

// ---- fused epilogue of every apply kernel ---------------------------------
// h = hscale * (H psi)[r]            (hscale = -1 runs -H, Lanczos.jl:261-265)
// mode 0: out = h                                   Hamiltonian.jl:211-273
// mode 1: out = (h - b psi)/a                       Hamiltonian.jl:286-301
// mode 2: out = 2 (h - b psi)/a - vprev             KPM_Sqw.jl:111-112, Chebyshev.jl:112-115
// then, optionally:  acc += ck * out                Chebyshev.jl:116
// reductions (per-CTA partial sums, slot-major [slot*nparts + cta]):
//   slot 0,1: dot(psi, out) = sum conj(psi) out     Lanczos.jl:50,124,219; Krylov.jl:155
//   slot 2  : Re dot(phi, out)                      KPM_Sqw.jl:114
//   slot 3  : ||out||^2                             KPM_Sqw.jl:117
// All auxiliary pointers address the LOCAL shard, indexed by local element.
enum { SD_EPI_PLAIN = 0, SD_EPI_RESCALED = 1, SD_EPI_CHEB = 2 };
enum { SD_RED_DOT_SELF = 1, SD_RED_DOT_PHI = 2, SD_RED_NORM2 = 4 };
#define SD_NSLOT 4
struct SdEpi {
    int mode;
    int red;                 // bitmask of SD_RED_*
    double hscale, a, b;
    const double *hscale_dev; // non-null: hscale / sqrt(*hscale_dev) is used  (deferred normalisation of the Lanczos vectors:
                              // *hscale_dev = ||u_j||^2 from the previous fused reduction, no host round trip)
    double ck_re, ck_im;
    const double *vprev;     // mode 2
    const double *phi;       // SD_RED_DOT_PHI
    double *acc;             // optional psi_t accumulation
    double *partials;        // [SD_NSLOT * nparts]
    unsigned nparts;
};

template <int NC>
struct SdVal {
    double c[NC];
};

// Applies the epilogue to one element.  `h` is H psi (before hscale `hs`), `p` is
// psi at the same index, `li` the local element index.  Accumulates the
// requested reductions into red[SD_NSLOT].
// The value part of the epilogue for one real component (mode arithmetic incl. the true division of
// Hamiltonian.jl:296-298): NOT inlined on the device -- sixty inlined copies of the FP64 division sequence made the
// fused block kernel twice the size of the plain one and instruction-fetch bound.
#if defined(__CUDACC__)
#define SD_EPI_VALUE_FN __device__ __noinline__
#else
#define SD_EPI_VALUE_FN inline
#endif
static SD_EPI_VALUE_FN double sd_epi_value(int mode, double hs, double a, double b, double h, double p, double vprev) {
    double v = hs * h;
    if (mode != SD_EPI_PLAIN) {
        v = (v - b * p) / a;
        if (mode == SD_EPI_CHEB) v = 2.0 * v - vprev;
    }
    return v;
}
template <int NC>
SD_HD SdVal<NC> sd_epilogue_hs(const SdEpi &e, double hs, SdVal<NC> h, SdVal<NC> p, uint64_t li,
                               double (&red)[SD_NSLOT]) {
    SdVal<NC> o;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (e.mode == SD_EPI_PLAIN) o.c[c] = hs * h.c[c];
        else o.c[c] = sd_epi_value(e.mode, hs, e.a, e.b, h.c[c], p.c[c], e.mode == SD_EPI_CHEB ? e.vprev[li * NC + c] : 0.0);
    }
    if (e.acc) {
        if (NC == 2) {
            e.acc[li * 2 + 0] += e.ck_re * o.c[0] - e.ck_im * o.c[NC - 1];
            e.acc[li * 2 + 1] += e.ck_re * o.c[NC - 1] + e.ck_im * o.c[0];
        } else {
            e.acc[li] += e.ck_re * o.c[0];
        }
    }
    if (e.red & SD_RED_DOT_SELF) {
        if (NC == 2) {
            red[0] += p.c[0] * o.c[0] + p.c[NC - 1] * o.c[NC - 1];
            red[1] += p.c[0] * o.c[NC - 1] - p.c[NC - 1] * o.c[0];
        } else {
            red[0] += p.c[0] * o.c[0];
        }
    }
    if (e.red & SD_RED_DOT_PHI) {
        if (NC == 2) red[2] += e.phi[li * 2] * o.c[0] + e.phi[li * 2 + 1] * o.c[NC - 1];
        else red[2] += e.phi[li] * o.c[0];
    }
    if (e.red & SD_RED_NORM2) {
#pragma unroll
        for (int c = 0; c < NC; ++c) red[3] += o.c[c] * o.c[c];
    }
    return o;
}
// hscale resolved per call (tiled / generic kernels); the block kernel resolves it once per CTA and calls sd_epilogue_hs
template <int NC>
SD_HD SdVal<NC> sd_epilogue(const SdEpi &e, SdVal<NC> h, SdVal<NC> p, uint64_t li, double (&red)[SD_NSLOT]) {
    return sd_epilogue_hs<NC>(e, e.hscale_dev ? e.hscale / sqrt(*e.hscale_dev) : e.hscale, h, p, li, red);
}
