// sd_tile.h -- the smem-tiled H.psi kernel body for the sector basis of an open
// nearest-neighbour chain (reference Hamiltonian.jl:211-273 restricted to the
// bond lists XXZChain builds, SpinModel.jl:63-90, with arbitrary per-bond J/Jz
// and per-site field).  Pull style, atomic-free, one output per state.
//
// The chain is cut into  prefix (A sites) | mid (M sites) | tail (T sites),
// B = M + T = suffix.  Rank order is "site 1 most significant", so
//   * one prefix configuration P owns a CONTIGUOUS rank range (a tile) holding
//     every suffix configuration with js = k - popc(P) set bits;
//   * a hop on a prefix bond maps the whole tile onto another whole tile with
//     the same local order (uniform rank shift): coalesced streaming reads;
//   * the prefix|suffix crossing bond is a uniform shift on a sub-range;
//   * hops on suffix bonds stay inside the tile: shared-memory gathers.
// Inside a tile, one thread owns a whole TAIL BLOCK: all C(T,jt) tail
// configurations of one mid configuration c (contiguous ranks).  Tail-internal
// hops are register moves fixed at compile time; a mid bond moves the whole
// block by a uniform shift, so bit tests / binomial lookups are paid once per
// block, not once per state.  Blocks are stored in shared memory class-major
// (class = jt) with an odd pitch, so a warp (32 consecutive blocks of one
// class) reads conflict-free.
//
// Phases (separated by CTA barriers):
//   0  header: prefix bits, tile base rank, neighbour-tile list        (1 thread)
//   1  flat, coalesced: own psi -> smem, g = sum_J psi[neighbour tiles] -> smem
//   2  block-mapped: diag + g + tail/mid/crossing hops -> result in smem
//   3  flat, coalesced: epilogue (rescale / Chebyshev / dots) -> out
//
// Everything here is __host__ __device__ so tests/emul can run the identical
// body on the CPU (test infrastructure; the product only runs it on the GPU).
#pragma once
#include "sd_common.h"

#define SD_TILE_MAXNB 32
#define SD_TILE_MAXT 6

struct SdTileParams {
    int L, k, A, B, M, T;
    uint64_t key_lo;                 // tile key of blockIdx.x == 0
    uint64_t key_hi;                 // one past the last key of this launch
    double Jhop[SD_MAX_L];           // hop coefficient of bond p (positions p, p+1)
    double Jz[SD_MAX_L];             // zz coefficient of bond p
    double h[SD_MAX_L + 1];          // field at position p
    double dtail[1 << SD_TILE_MAXT]; // diag of the tail sites + tail-internal zz, by tail bits
    const uint64_t *binom;           // [65*65]
    const uint16_t *perm;            // [perm_off[js] + l] -> smem element position
    const uint16_t *midcfg;          // [mid_off[jm] + u] -> mid bits c
    const uint16_t *urank;           // [c] -> class-local index u
    const double *dmid;              // [c] -> diag of mid sites + mid-internal zz
    const uint32_t *cls_base;        // [js*(SD_TILE_MAXT+2) + jt] smem start of class jt; [.. + T+1] = cap
    uint32_t perm_off[32];           // B <= 30
    uint32_t mid_off[32];            // M <= 30
    SdShardMap shards;
};

struct SdTileHdr {
    uint64_t base;                   // rank of the tile's first state
    uint32_t size;                   // C(B, js)
    int js, valid, nnb;
    uint32_t nslots;
    double dP[2];                    // prefix diag + crossing zz, by first mid bit
    uint32_t cls_base[SD_TILE_MAXT + 2];
    uint32_t slot_base[SD_TILE_MAXT + 2];
    uint32_t n_items[SD_TILE_MAXT + 1];
    int64_t nb_off[SD_TILE_MAXNB];   // rank shift of neighbour tile
    double nb_J[SD_TILE_MAXNB];
    uint32_t nb_lo[SD_TILE_MAXNB], nb_hi[SD_TILE_MAXNB];
    const double *nb_ptr[SD_TILE_MAXNB];  // virtual base if the range sits in one shard, else null
    uint64_t t_base[SD_TILE_MAXNB];  // phase-0 scratch: per-position rank terms
    double t_diag[SD_TILE_MAXNB];    //                  per-position diagonal terms
    double red[SD_NSLOT][32];
};

template <int NC>
struct SdTileView {
    SdTileHdr *hdr;
    double *spsi;                    // [cap*NC]
    double *sg;                      // [cap*NC]
    uint16_t *binomM;                // [(M+1)*(M+1)]
};

SD_HD size_t sd_tile_smem_bytes(int NC, uint32_t cap, int M) {
    size_t b = sizeof(SdTileHdr);
    b = (b + 15) & ~(size_t)15;
    b += (size_t)2 * cap * NC * sizeof(double);
    b += (size_t)(M + 1) * (M + 1) * sizeof(uint16_t);
    return (b + 15) & ~(size_t)15;
}

template <int NC>
SD_HD SdTileView<NC> sd_tile_carve(void *smem, uint32_t cap) {
    SdTileView<NC> v;
    char *p = (char *)smem;
    v.hdr = (SdTileHdr *)p;
    p += (sizeof(SdTileHdr) + 15) & ~(size_t)15;
    v.spsi = (double *)p;
    p += (size_t)cap * NC * sizeof(double);
    v.sg = (double *)p;
    p += (size_t)cap * NC * sizeof(double);
    v.binomM = (uint16_t *)p;
    return v;
}

// prefix bits of tile `key`: keys enumerate prefixes in rank order, bit q of the
// prefix is the complement of key digit A-1-q ("1 first").
SD_HD uint64_t sd_tile_prefix_bits(uint64_t key, int A) {
    uint64_t Pb = 0;
    for (int q = 0; q < A; ++q)
        if (!((key >> (A - 1 - q)) & 1ULL)) Pb |= 1ULL << q;
    return Pb;
}

// ---------------------------------------------------------------- phase 0
// 0a runs on every thread: thread q owns prefix position q and bond q (one
// binomial lookup each, all in flight together); 0b (one thread) sums the
// per-position terms in order, compacts the neighbour list and lays out the
// classes.  A single serial thread doing all of it cost ~35 us per tile.
template <int NC>
SD_HD void sd_tile_phase0a(const SdTileParams &P, uint64_t key, const SdTileView<NC> &v,
                           unsigned tid, unsigned nthreads) {
    SdTileHdr &H = *v.hdr;
    const int L = P.L, k = P.k, A = P.A, B = P.B, M = P.M, T = P.T;
    const uint64_t *C = P.binom;
    const uint64_t Pb = sd_tile_prefix_bits(key, A);
    const int js = k - SD_POPC64(Pb);
    const bool valid = (js >= 0 && js <= B);
    if (tid == 0) {
        H.valid = valid ? 1 : 0;
        H.js = js;
        H.size = valid ? (uint32_t)sd_binom_at(C, SD_BINOM_DIM, B, js) : 0u;
    }
    if (!valid) return;
    const uint32_t size = (uint32_t)sd_binom_at(C, SD_BINOM_DIM, B, js);
    for (int q = (int)tid; q < A; q += (int)nthreads) {
        const int bit = (int)((Pb >> q) & 1ULL);
        const int below = SD_POPC64(Pb & ((1ULL << q) - 1));             // set bits at positions < q
        const double sq = bit ? 0.5 : -0.5;
        double d = P.h[q] * sq;
        H.t_base[q] = bit ? 0ULL : sd_binom_at(C, SD_BINOM_DIM, L - 1 - q, k - below - 1);
        int64_t off = 0;
        uint32_t lo = 0, hi = 0;
        const double J = P.Jhop[q];
        if (q + 1 < A) {                                                  // prefix-internal bond q
            const int bn = (int)((Pb >> (q + 1)) & 1ULL);
            d += P.Jz[q] * sq * (bn ? 0.5 : -0.5);
            if (bit != bn && J != 0.0) {
                const uint64_t dl = sd_binom_at(C, SD_BINOM_DIM, L - 2 - q, k - (below + bit + bn));
                off = bit ? (int64_t)dl : -(int64_t)dl;
                hi = size;
            }
        } else if (J != 0.0) {                                            // prefix|suffix crossing bond
            const uint32_t n1 = (uint32_t)sd_binom_at(C, SD_BINOM_DIM, B - 1, js - 1);   // first suffix bit = 1
            if (bit) {                       // (1,0) -> (0,1): states with first suffix bit 0 move up
                if (n1 < size) { off = (int64_t)sd_binom_at(C, SD_BINOM_DIM, B - 1, js); lo = n1; hi = size; }
            } else if (n1 > 0) {             // (0,1) -> (1,0)
                off = -(int64_t)n1; lo = 0; hi = n1;
            }
        }
        H.t_diag[q] = d;
        H.nb_off[q] = off; H.nb_J[q] = J; H.nb_lo[q] = lo; H.nb_hi[q] = hi;
    }
    for (int i = (int)tid; i < (M + 1) * (M + 1); i += (int)nthreads) {
        const int nn = i / (M + 1), r = i - nn * (M + 1);
        v.binomM[i] = (uint16_t)sd_binom_at(C, SD_BINOM_DIM, nn, r);
    }
    for (int jt = (int)tid; jt <= T + 1; jt += (int)nthreads) {
        H.cls_base[jt] = P.cls_base[js * (SD_TILE_MAXT + 2) + jt];
        if (jt <= T) {
            const int jm = js - jt;
            H.n_items[jt] = (jm >= 0 && jm <= M) ? (uint32_t)sd_binom_at(C, SD_BINOM_DIM, M, jm) : 0u;
        }
    }
}

template <int NC>
SD_HD void sd_tile_phase0b(const SdTileParams &P, uint64_t key, const SdTileView<NC> &v,
                           const SdVecView &psi) {
    SdTileHdr &H = *v.hdr;
    if (!H.valid) return;
    const int A = P.A, T = P.T;
    uint64_t base = 0;
    double dpre = 0.0;
    for (int q = 0; q < A; ++q) { base += H.t_base[q]; dpre += H.t_diag[q]; }
    H.base = base;
    if (A > 0) {
        const uint64_t Pb = sd_tile_prefix_bits(key, A);
        const double sl = ((Pb >> (A - 1)) & 1ULL) ? 0.5 : -0.5;
        H.dP[0] = dpre + P.Jz[A - 1] * sl * (-0.5);
        H.dP[1] = dpre + P.Jz[A - 1] * sl * (0.5);
    } else {
        H.dP[0] = H.dP[1] = dpre;
    }
    int n = 0;
    const bool single = P.shards.world == 1;
    for (int q = 0; q < A; ++q) {
        const uint32_t lo = H.nb_lo[q], hi = H.nb_hi[q];
        if (hi <= lo) continue;
        const int64_t off = H.nb_off[q];
        const double J = H.nb_J[q];
        const double *ptr;
        if (single) {
            ptr = psi.base[0] + (int64_t)NC * ((int64_t)base + off);
        } else {
            const uint64_t r0 = (uint64_t)((int64_t)(base + lo) + off);
            const uint64_t r1 = (uint64_t)((int64_t)(base + hi - 1) + off);
            const int g0 = sd_owner(P.shards, r0), g1 = sd_owner(P.shards, r1);
            ptr = (g0 == g1) ? psi.base[g0] + (int64_t)NC * ((int64_t)base + off) : nullptr;
        }
        H.nb_off[n] = off; H.nb_J[n] = J; H.nb_lo[n] = lo; H.nb_hi[n] = hi; H.nb_ptr[n] = ptr;
        ++n;
    }
    H.nnb = n;
    uint32_t slot = 0;
    for (int jt = 0; jt <= T; ++jt) {
        H.slot_base[jt] = slot;
        slot += (H.n_items[jt] + 31u) & ~31u;
    }
    H.slot_base[T + 1] = slot;
    H.nslots = slot;
}

// ---------------------------------------------------------------- phase 1
template <int NC>
SD_HD void sd_tile_phase1(const SdTileParams &P, const SdTileView<NC> &v, const SdVecView &psi,
                          unsigned tid, unsigned nthreads) {
    const SdTileHdr &H = *v.hdr;
    const uint16_t *perm = P.perm + P.perm_off[H.js];
    const double *own = psi.base[P.shards.rank] + (int64_t)NC * (int64_t)H.base;
    const int nnb = H.nnb;
    for (uint32_t l = tid; l < H.size; l += nthreads) {
        const uint32_t pos = perm[l];
        double g[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) g[c] = 0.0;
        for (int i = 0; i < nnb; ++i) {
            if (l >= H.nb_lo[i] && l < H.nb_hi[i]) {
                const double *q = H.nb_ptr[i];
                if (q) {
                    q += (size_t)NC * l;
                } else {
                    const uint64_t r = (uint64_t)((int64_t)(H.base + l) + H.nb_off[i]);
                    q = psi.base[sd_owner(P.shards, r)] + (size_t)NC * r;
                }
                const double J = H.nb_J[i];
#pragma unroll
                for (int c = 0; c < NC; ++c) g[c] += J * q[c];
            }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            v.spsi[(size_t)pos * NC + c] = own[(size_t)l * NC + c];
            v.sg[(size_t)pos * NC + c] = g[c];
        }
    }
}

// ---------------------------------------------------------------- phase 2
// compile-time unrolled helpers over the tail block of class (T, JT)
template <int NC, int T, int JT, int t, int q>
struct SdTailHop {
    static SD_HD void run(double (&acc)[sd_cbinom(T, JT) * NC], const double (&own)[sd_cbinom(T, JT) * NC],
                          const double *Jt) {
        constexpr unsigned cfg = sd_tail_cfg(T, JT, t);
        constexpr bool act = (((cfg >> q) ^ (cfg >> (q + 1))) & 1u) != 0;
        if constexpr (act) {
            constexpr unsigned cf2 = cfg ^ (3u << q);
            constexpr int t2 = sd_tail_rank(T, JT, cf2);
            const double J = Jt[q];
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[t * NC + c] += J * own[t2 * NC + c];
        }
        if constexpr (q + 2 < T) SdTailHop<NC, T, JT, t, q + 1>::run(acc, own, Jt);
    }
};
template <int NC, int T, int JT, int t>
struct SdTailRow {
    static SD_HD void run(double (&acc)[sd_cbinom(T, JT) * NC], const double (&own)[sd_cbinom(T, JT) * NC],
                          const double *Jt) {
        if constexpr (T >= 2) SdTailHop<NC, T, JT, t, 0>::run(acc, own, Jt);
        if constexpr (t + 1 < sd_cbinom(T, JT)) SdTailRow<NC, T, JT, t + 1>::run(acc, own, Jt);
    }
};
// diag of state t of the block: dthread + dtail[cfg] +- qx (crossing zz with the last mid bit)
template <int NC, int T, int JT, int t>
struct SdTailInit {
    static SD_HD void run(double (&acc)[sd_cbinom(T, JT) * NC], const double (&own)[sd_cbinom(T, JT) * NC],
                          const double *g, const double *dtail, double dthread, double dx) {
        constexpr unsigned cfg = sd_tail_cfg(T, JT, t);
        const double d = dthread + dtail[cfg] + ((cfg & 1u) ? dx : -dx);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[t * NC + c] = g[t * NC + c] + d * own[t * NC + c];
        if constexpr (t + 1 < sd_cbinom(T, JT))
            SdTailInit<NC, T, JT, t + 1>::run(acc, own, g, dtail, dthread, dx);
    }
};

template <int NC, int T, int JT>
SD_HD void sd_tile_block(const SdTileParams &P, const SdTileView<NC> &v, uint32_t u) {
    constexpr int NT = sd_cbinom(T, JT);
    constexpr int NTP = NT | 1;                      // odd pitch: conflict-free across a warp
    const SdTileHdr &H = *v.hdr;
    const int M = P.M, A = P.A;
    const int jm = H.js - JT;
    const unsigned c = P.midcfg[P.mid_off[jm] + u];
    const uint32_t cb = H.cls_base[JT];
    const uint32_t pos = cb + u * NTP;
    double own[NT * NC], acc[NT * NC];
    {
        const double *sp = v.spsi + (size_t)pos * NC;
#pragma unroll
        for (int i = 0; i < NT * NC; ++i) own[i] = sp[i];
    }
    // diagonal + neighbour-tile sum
    const int c_first = (int)(c & 1u), c_last = (int)((c >> (M - 1)) & 1u);
    const double dthread = H.dP[c_first] + P.dmid[c];
    const double qx = P.Jz[A + M - 1] * 0.25;
    const double dx = c_last ? qx : -qx;             // +qx when tail bit 0 equals the last mid bit
    SdTailInit<NC, T, JT, 0>::run(acc, own, v.sg + (size_t)pos * NC, P.dtail, dthread, dx);
    // tail-internal hops: registers only
    SdTailRow<NC, T, JT, 0>::run(acc, own, P.Jhop + A + M);
    // mid-internal hops: the whole block shifts by a class-local index delta
    for (int pm = 0; pm + 1 < M; ++pm) {
        const unsigned b0 = (c >> pm) & 1u, b1 = (c >> (pm + 1)) & 1u;
        if (b0 != b1) {
            const double J = P.Jhop[A + pm];
            const int mm = SD_POPC32(c >> (pm + 2));
            const uint32_t du = v.binomM[(M - 2 - pm) * (M + 1) + mm];
            const uint32_t u2 = b0 ? u + du : u - du;
            const double *sp = v.spsi + (size_t)(cb + u2 * NTP) * NC;
#pragma unroll
            for (int i = 0; i < NT * NC; ++i) acc[i] += J * sp[i];
        }
    }
    // mid|tail crossing bond (positions A+M-1, A+M)
    {
        const double J = P.Jhop[A + M - 1];
        constexpr int n1 = sd_cbinom(T - 1, JT - 1);          // tail configs with first bit 1
        if (c_last) {
            if constexpr (JT < T && NT - n1 > 0) {                       // our first bit 0 -> partner class JT+1, first part
                constexpr int NTP2 = sd_cbinom(T, JT + 1) | 1;
                const unsigned c2 = c ^ (1u << (M - 1));
                const uint32_t u2 = P.urank[c2];
                const double *sp = v.spsi + (size_t)(H.cls_base[JT + 1] + u2 * NTP2) * NC;
#pragma unroll
                for (int t = n1; t < NT; ++t)
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) acc[t * NC + cc] += J * sp[(t - n1) * NC + cc];
            }
        } else {
            if constexpr (JT > 0 && n1 > 0) {                            // our first bit 1 -> partner class JT-1, second part
                constexpr int NTP2 = sd_cbinom(T, JT - 1) | 1;
                constexpr int n1p = sd_cbinom(T - 1, JT - 2);
                const unsigned c2 = c | (1u << (M - 1));
                const uint32_t u2 = P.urank[c2];
                const double *sp = v.spsi + (size_t)(H.cls_base[JT - 1] + u2 * NTP2) * NC;
#pragma unroll
                for (int t = 0; t < n1; ++t)
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) acc[t * NC + cc] += J * sp[(n1p + t) * NC + cc];
            }
        }
    }
    {
        double *sp = v.sg + (size_t)pos * NC;
#pragma unroll
        for (int i = 0; i < NT * NC; ++i) sp[i] = acc[i];
    }
}

template <int NC, int T, int JT>
struct SdTileDispatch {
    static SD_HD void run(const SdTileParams &P, const SdTileView<NC> &v, int jt, uint32_t u) {
        if (jt == JT) sd_tile_block<NC, T, JT>(P, v, u);
        else if constexpr (JT > 0) SdTileDispatch<NC, T, JT - 1>::run(P, v, jt, u);
    }
};

template <int NC, int T>
SD_HD void sd_tile_phase2(const SdTileParams &P, const SdTileView<NC> &v, unsigned tid,
                          unsigned nthreads) {
    const SdTileHdr &H = *v.hdr;
    for (uint32_t s = tid; s < H.nslots; s += nthreads) {
        int jt = 0;
#pragma unroll
        for (int j = 1; j <= T; ++j) jt += (s >= H.slot_base[j]) ? 1 : 0;
        const uint32_t u = s - H.slot_base[jt];
        if (u < H.n_items[jt]) SdTileDispatch<NC, T, T>::run(P, v, jt, u);
    }
}

// ---------------------------------------------------------------- phase 3
template <int NC>
SD_HD void sd_tile_phase3(const SdTileParams &P, const SdTileView<NC> &v, double *out_vbase,
                          const SdEpi &epi, unsigned tid, unsigned nthreads,
                          double (&red)[SD_NSLOT]) {
    const SdTileHdr &H = *v.hdr;
    const uint16_t *perm = P.perm + P.perm_off[H.js];
    const uint64_t lstart = P.shards.start[P.shards.rank];
    double *o = out_vbase + (int64_t)NC * (int64_t)H.base;
    for (uint32_t l = tid; l < H.size; l += nthreads) {
        const uint32_t pos = perm[l];
        SdVal<NC> h, p;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            h.c[c] = v.sg[(size_t)pos * NC + c];
            p.c[c] = v.spsi[(size_t)pos * NC + c];
        }
        const SdVal<NC> r = sd_epilogue<NC>(epi, h, p, H.base + l - lstart, red);
#pragma unroll
        for (int c = 0; c < NC; ++c) o[(size_t)l * NC + c] = r.c[c];
    }
}
