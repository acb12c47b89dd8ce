// sd_tile_host.h -- host-side construction of the lookup tables of the tiled
// kernel (sd_tile.h) and of the tile-aligned shard boundaries.  Pure C++ (no
// CUDA) so the CPU emulation used by the tests builds the very same tables.
#pragma once
#include <vector>
#include <cstring>
#include "sd_tile.h"

struct SdTileHost {
    SdTileParams P;                 // table pointers left null; caller points them at host or device copies
    std::vector<uint64_t> binom;    // [65*65]
    std::vector<uint16_t> perm, midcfg, urank;
    std::vector<double> dmid;
    std::vector<uint32_t> cls_base; // [(B+1)*(SD_TILE_MAXT+2)]
    uint32_t cap_max = 0;           // largest padded tile, in elements
};

// Jhop/Jz: coefficient of bond p (positions p,p+1), length L-1 (0 if absent); h: length L.
// Returns false if the split is not representable (then the generic kernel is used).
static inline bool sd_tile_build(int L, int k, int B, int T, const double *Jhop, const double *Jz,
                                 const double *h, SdTileHost &o) {
    if (k < 0 || k > L || T < 1 || T > SD_TILE_MAXT || B > L || B - T < 2 || B > 30) return false;
    const int M = B - T, A = L - B;
    if (M > 15 || A > SD_TILE_MAXNB - 1) return false;   // u16 mid tables; neighbour list capacity
    o.binom.assign(SD_BINOM_DIM * SD_BINOM_DIM, 0);
    sd_fill_binom(o.binom.data());
    const uint64_t *C = o.binom.data();
    SdTileParams &P = o.P;
    std::memset(&P, 0, sizeof(P));
    P.L = L; P.k = k; P.A = A; P.B = B; P.M = M; P.T = T;
    for (int p = 0; p + 1 < L; ++p) { P.Jhop[p] = Jhop[p]; P.Jz[p] = Jz[p]; }
    for (int p = 0; p < L; ++p) P.h[p] = h[p];
    // mid configurations, class-sorted (class = popcount), "1 first" lexicographic inside a class
    o.midcfg.assign((size_t)1 << M, 0);
    o.urank.assign((size_t)1 << M, 0);
    o.dmid.assign((size_t)1 << M, 0.0);
    uint32_t off = 0;
    for (int jm = 0; jm <= M; ++jm) {
        P.mid_off[jm] = off;
        const uint32_t n = (uint32_t)C[M * SD_BINOM_DIM + jm];
        for (uint32_t u = 0; u < n; ++u) {
            const unsigned c = (unsigned)sd_unrank_state(u, M, jm, C, SD_BINOM_DIM);
            o.midcfg[off + u] = (uint16_t)c;
            o.urank[c] = (uint16_t)u;
        }
        off += n;
    }
    P.mid_off[M + 1] = off;
    for (unsigned c = 0; c < (1u << M); ++c) {
        double d = 0.0;
        for (int q = 0; q < M; ++q) {
            const double s = ((c >> q) & 1u) ? 0.5 : -0.5;
            d += h[A + q] * s;
            if (q + 1 < M) d += Jz[A + q] * s * (((c >> (q + 1)) & 1u) ? 0.5 : -0.5);
        }
        o.dmid[c] = d;
    }
    for (unsigned t = 0; t < (1u << T); ++t) {
        double d = 0.0;
        for (int q = 0; q < T; ++q) {
            const double s = ((t >> q) & 1u) ? 0.5 : -0.5;
            d += h[A + M + q] * s;
            if (q + 1 < T) d += Jz[A + M + q] * s * (((t >> (q + 1)) & 1u) ? 0.5 : -0.5);
        }
        P.dtail[t] = d;
    }
    // class-major smem layout per suffix popcount js
    o.cls_base.assign((size_t)(B + 1) * (SD_TILE_MAXT + 2), 0);
    o.cap_max = 0;
    for (int js = 0; js <= B; ++js) {
        uint32_t run = 0;
        for (int jt = 0; jt <= T; ++jt) {
            o.cls_base[js * (SD_TILE_MAXT + 2) + jt] = run;
            const int jm = js - jt;
            if (jm >= 0 && jm <= M)
                run += (uint32_t)C[M * SD_BINOM_DIM + jm] * ((uint32_t)C[T * SD_BINOM_DIM + jt] | 1u);
        }
        for (int jt = T + 1; jt < SD_TILE_MAXT + 2; ++jt) o.cls_base[js * (SD_TILE_MAXT + 2) + jt] = run;
        if (run > o.cap_max) o.cap_max = run;
    }
    if (o.cap_max > 65535u) return false;
    // permutation flat suffix index -> smem position
    o.perm.assign((size_t)1 << B, 0);
    uint32_t poff = 0;
    for (int js = 0; js <= B; ++js) {
        P.perm_off[js] = poff;
        const uint32_t n = (uint32_t)C[B * SD_BINOM_DIM + js];
        for (uint32_t l = 0; l < n; ++l) {
            const uint64_t s = sd_unrank_state(l, B, js, C, SD_BINOM_DIM);
            const unsigned c = (unsigned)(s & ((1ULL << M) - 1));
            const unsigned tau = (unsigned)(s >> M);
            const int jt = __builtin_popcount(tau);
            const uint32_t ntp = (uint32_t)C[T * SD_BINOM_DIM + jt] | 1u;
            const uint32_t t = (uint32_t)sd_rank_state(tau, T, jt, C, SD_BINOM_DIM);
            o.perm[poff + l] = (uint16_t)(o.cls_base[js * (SD_TILE_MAXT + 2) + jt] + o.urank[c] * ntp + t);
        }
        poff += n;
    }
    P.perm_off[B + 1] = poff;
    P.key_lo = 0;
    P.key_hi = 1ULL << A;
    P.shards.world = 1; P.shards.rank = 0;
    P.shards.start[0] = 0;
    for (int g = 1; g <= SD_MAX_WORLD; ++g) P.shards.start[g] = C[L * SD_BINOM_DIM + k];
    return true;
}

// First tile key whose base rank is >= r (r = N maps to 2^A).  Tiles are in
// rank order, so this is the key of the tile containing r, +1 unless r is its base.
static inline uint64_t sd_tile_key_of_rank(const SdTileHost &o, uint64_t r, uint64_t *base_out) {
    const SdTileParams &P = o.P;
    const uint64_t *C = o.binom.data();
    const uint64_t N = C[P.L * SD_BINOM_DIM + P.k];
    if (r >= N) { if (base_out) *base_out = N; return 1ULL << P.A; }
    const uint64_t s = sd_unrank_state(r, P.L, P.k, C, SD_BINOM_DIM);
    const uint64_t Pb = s & ((P.A >= 64) ? ~0ULL : ((1ULL << P.A) - 1));
    uint64_t key = 0;
    for (int q = 0; q < P.A; ++q)
        if (!((Pb >> q) & 1ULL)) key |= 1ULL << (P.A - 1 - q);
    // base rank of that tile = rank of (prefix, first suffix configuration)
    const int js = P.k - __builtin_popcountll(Pb);
    uint64_t first = Pb;
    for (int i = 0; i < js; ++i) first |= 1ULL << (P.A + i);
    const uint64_t base = sd_rank_state(first, P.L, P.k, C, SD_BINOM_DIM);
    if (base_out) *base_out = base;
    return key;
}

// Tile-aligned shard boundaries: bounds[g] = base of the tile holding rank g*N/world.
static inline void sd_tile_shard_bounds(const SdTileHost &o, int world, uint64_t *bounds, uint64_t *keys) {
    const uint64_t *C = o.binom.data();
    const uint64_t N = C[o.P.L * SD_BINOM_DIM + o.P.k];
    for (int g = 0; g <= world; ++g) {
        const uint64_t r = (g == world) ? N : (uint64_t)(((unsigned __int128)N * g) / world);
        uint64_t base = 0;
        const uint64_t key = sd_tile_key_of_rank(o, r, &base);
        bounds[g] = base;
        if (keys) keys[g] = key;
    }
}
