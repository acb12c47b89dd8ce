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
    std::vector<uint16_t> perm;     // flat suffix index -> smem position, all js concatenated
    std::vector<SdItem> items;      // phase-2 work items, all js concatenated
    std::vector<SdJsInfo> jsinfo;   // [B+1]
    std::vector<uint16_t> binomM;   // [(M+1)*(M+1)]
    uint32_t cap_max = 0;           // largest padded tile, in elements
    uint32_t nslots_max = 0;
};

// Jhop/Jz: coefficient of bond p (positions p,p+1), length L-1 (0 if absent); h: length L.
// Returns false if the split is not representable (then the generic kernel is used).
// Prefix bonds q < qfar can shift a tile by more than `far_bytes` (their largest shift is
// C(L-2-q, (L-2-q)/2) elements): such neighbour tiles cannot be reused from L2.
static inline int sd_tile_qfar(int L, int A, const uint64_t *C, uint64_t far_bytes, int elem_bytes) {
    int q = 0;
    while (q + 1 < A && q < 31) {
        const int n = L - 2 - q;
        if (C[n * SD_BINOM_DIM + n / 2] * (uint64_t)elem_bytes <= far_bytes) break;
        ++q;
    }
    return q;
}

static inline bool sd_tile_build(int L, int k, int B, int T, const double *Jhop, const double *Jz,
                                 const double *h, SdTileHost &o) {
    if (k < 0 || k > L || T < 1 || T > SD_TILE_MAXT || B > L || B - T < 2 || B > SD_TILE_MAXB) return false;
    const int M = B - T, A = L - B;
    if (M > 15 || A > SD_TILE_MAXNB) return false;       // u16 mid tables; neighbour list capacity
    o.binom.assign(SD_BINOM_DIM * SD_BINOM_DIM, 0);
    sd_fill_binom(o.binom.data());
    const uint64_t *C = o.binom.data();
    SdTileParams &P = o.P;
    std::memset(&P, 0, sizeof(P));
    P.L = L; P.k = k; P.A = A; P.B = B; P.M = M; P.T = T;
    P.qfar = 0;
    P.hop_mask = 0;
    for (int p = 0; p + 1 < L && p < 32; ++p) if (Jhop[p] != 0.0) P.hop_mask |= 1u << p;
    for (int p = 0; p + 1 < L; ++p) { P.Jhop[p] = Jhop[p]; P.Jz[p] = Jz[p]; }
    for (int p = 0; p < L; ++p) P.h[p] = h[p];
    // mid configurations, class-sorted (class = popcount), "1 first" lexicographic inside a class
    std::vector<uint16_t> midcfg((size_t)1 << M, 0), urank((size_t)1 << M, 0);
    std::vector<uint32_t> mid_off(M + 2, 0);
    std::vector<double> dmid((size_t)1 << M, 0.0);
    uint32_t off = 0;
    for (int jm = 0; jm <= M; ++jm) {
        mid_off[jm] = off;
        const uint32_t n = (uint32_t)C[M * SD_BINOM_DIM + jm];
        for (uint32_t u = 0; u < n; ++u) {
            const unsigned c = (unsigned)sd_unrank_state(u, M, jm, C, SD_BINOM_DIM);
            midcfg[off + u] = (uint16_t)c;
            urank[c] = (uint16_t)u;
        }
        off += n;
    }
    for (unsigned c = 0; c < (1u << M); ++c) {
        double d = 0.0;
        for (int q = 0; q < M; ++q) {
            const double s = ((c >> q) & 1u) ? 0.5 : -0.5;
            d += h[A + q] * s;
            if (q + 1 < M) d += Jz[A + q] * s * (((c >> (q + 1)) & 1u) ? 0.5 : -0.5);
        }
        dmid[c] = d;
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
    // per-js layout: class-major smem offsets, phase-2 slots (classes padded to warps), work items
    o.jsinfo.assign(B + 1, SdJsInfo());
    o.items.clear();
    o.cap_max = 0; o.nslots_max = 0;
    uint32_t poff = 0;
    for (int js = 0; js <= B; ++js) {
        SdJsInfo &I = o.jsinfo[js];
        std::memset(&I, 0, sizeof(I));
        uint32_t run = 0, slot = 0;
        I.item_off = (uint32_t)o.items.size();
        for (int jt = 0; jt <= T; ++jt) {
            I.cls_base[jt] = run;
            const int jm = js - jt;
            if (jm < 0 || jm > M) continue;
            const uint32_t ni = (uint32_t)C[M * SD_BINOM_DIM + jm];
            run += ni * ((uint32_t)C[T * SD_BINOM_DIM + jt] | 1u);
            const uint32_t padded = (ni + 31u) & ~31u;
            for (uint32_t u = 0; u < padded; ++u) {
                SdItem it;
                std::memset(&it, 0, sizeof(it));
                it.c = 0xFFFFu;
                if (u < ni) {
                    const unsigned c = midcfg[mid_off[jm] + u];
                    it.c = (uint16_t)c; it.u = (uint16_t)u; it.jt = (uint16_t)jt;
                    it.u2 = urank[c ^ (1u << (M - 1))];
                    it.dmid = dmid[c];
                }
                o.items.push_back(it);
            }
            slot += padded;
        }
        for (int jt = T + 1; jt < SD_TILE_MAXT + 2; ++jt) I.cls_base[jt] = run;
        I.nslots = slot;
        I.size = (uint32_t)C[B * SD_BINOM_DIM + js];
        I.n1 = js >= 1 ? (uint32_t)C[(B - 1) * SD_BINOM_DIM + js - 1] : 0u;
        I.ncross = (uint32_t)C[(B - 1) * SD_BINOM_DIM + js];
        I.perm_off = poff;
        poff += I.size;
        if (run > o.cap_max) o.cap_max = run;
        if (slot > o.nslots_max) o.nslots_max = slot;
    }
    if (o.cap_max > 65535u) return false;
    for (int js = 0; js <= B; ++js) P.js[js] = o.jsinfo[js];
    o.binomM.assign((size_t)(M + 1) * (M + 1), 0);
    for (int nn = 0; nn <= M; ++nn)
        for (int r = 0; r <= nn; ++r) o.binomM[nn * (M + 1) + r] = (uint16_t)C[nn * SD_BINOM_DIM + r];
    // permutation flat suffix index -> smem position
    o.perm.assign((size_t)1 << B, 0);
    for (int js = 0; js <= B; ++js) {
        const SdJsInfo &I = o.jsinfo[js];
        for (uint32_t l = 0; l < I.size; ++l) {
            const uint64_t s = sd_unrank_state(l, B, js, C, SD_BINOM_DIM);
            const unsigned c = (unsigned)(s & ((1ULL << M) - 1));
            const unsigned tau = (unsigned)(s >> M);
            const int jt = __builtin_popcount(tau);
            const uint32_t ntp = (uint32_t)C[T * SD_BINOM_DIM + jt] | 1u;
            const uint32_t t = (uint32_t)sd_rank_state(tau, T, jt, C, SD_BINOM_DIM);
            o.perm[I.perm_off + l] = (uint16_t)(I.cls_base[jt] + urank[c] * ntp + t);
        }
    }
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

// The same with given cut positions: cum[g] = basis rank at which shard g starts (cum[0] = 0, cum[world] = N,
// non-decreasing); every cut is moved down to the base of the tile that holds it.
static inline void sd_tile_shard_bounds_at(const SdTileHost &o, int world, const uint64_t *cum, uint64_t *bounds, uint64_t *keys) {
    for (int g = 0; g <= world; ++g) {
        uint64_t base = 0;
        const uint64_t key = sd_tile_key_of_rank(o, cum[g], &base);
        bounds[g] = base;
        if (keys) keys[g] = key;
    }
}
