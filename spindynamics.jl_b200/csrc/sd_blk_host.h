// sd_blk_host.h -- host-side tables of the block-layout kernel (sd_blk.h): padded tile
// offsets, per-suffix-popcount class layout, unit lists, per-mid-configuration work items.
// Pure C++ (no CUDA).
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>
#include "sd_blk.h"

struct SdBlkHost {
    SdBlkParams P;                    // table pointers left null; the caller points them at device copies
    std::vector<uint64_t> binom;      // [65*65]
    std::vector<uint64_t> W;          // [A*(A+1)]
    std::vector<SdBlkJs> js;          // [B+1]
    std::vector<uint16_t> units;      // [2][(B+1)*MAXUNITS] item codes jt << 12 | chunk << 8 | unit-in-class
    std::vector<SdBlkItem> items;
    std::vector<double> dmid;         // [1 << M], in item order (dmid[item index])
    std::vector<uint16_t> urank;      // [1 << M] class-local index of a mid configuration
    uint64_t n_store = 0;             // stored elements of the whole vector (all shards)
};

static inline bool sd_blk_build(int L, int k, const double *Jhop, const double *Jz, const double *h, SdBlkHost &o) {
    constexpr int M = SD_BLK_M, T = SD_BLK_T, B = SD_BLK_B;
    const int A = L - B;
    if (A < 1 || A > SD_BLK_MAXA || L > SD_MAX_L || k < 0 || k > L) return false;
    o.binom.assign(SD_BINOM_DIM * SD_BINOM_DIM, 0);
    sd_fill_binom(o.binom.data());
    const uint64_t *C = o.binom.data();
    SdBlkParams &P = o.P;
    std::memset(&P, 0, sizeof(P));
    P.L = L; P.k = k; P.A = A;
    for (int p = 0; p + 1 < L; ++p) { P.Jhop[p] = Jhop[p]; P.Jz[p] = Jz[p]; }
    for (int p = 0; p < L; ++p) P.h[p] = h[p];
    // mid configurations: class = popcount, "1 first" lexicographic inside a class
    std::vector<std::vector<uint16_t>> midcfg(M + 1);
    o.urank.assign((size_t)1 << M, 0);
    for (int jm = 0; jm <= M; ++jm) {
        const uint32_t n = (uint32_t)C[M * SD_BINOM_DIM + jm];
        midcfg[jm].resize(n);
        for (uint32_t u = 0; u < n; ++u) {
            const unsigned c = (unsigned)sd_unrank_state(u, M, jm, C, SD_BINOM_DIM);
            midcfg[jm][u] = (uint16_t)c;
            o.urank[c] = (uint16_t)u;
        }
    }
    std::vector<double> dmid_cfg((size_t)1 << M, 0.0);             // by mid configuration bits; stored in item order below
    for (unsigned c = 0; c < (1u << M); ++c) {
        double d = 0.0;
        for (int q = 0; q < M; ++q) {
            const double s = ((c >> q) & 1u) ? 0.5 : -0.5;
            d += h[A + q] * s;
            if (q + 1 < M) d += Jz[A + q] * s * (((c >> (q + 1)) & 1u) ? 0.5 : -0.5);
        }
        dmid_cfg[c] = d;
    }
    for (int q = 0; q < M; ++q) P.Jmid[q] = Jhop[A + q];             // [M-1]: the mid|tail bond
    for (int q = 0; q + 1 < T; ++q) P.Jtail[q] = Jhop[A + M + q];
    P.qx = Jz[A + M - 1] * 0.25;
    for (unsigned t = 0; t < (1u << T); ++t) {
        double d = 0.0;
        for (int q = 0; q < T; ++q) {
            const double s = ((t >> q) & 1u) ? 0.5 : -0.5;
            d += h[A + M + q] * s;
            if (q + 1 < T) d += Jz[A + M + q] * s * (((t >> (q + 1)) & 1u) ? 0.5 : -0.5);
        }
        P.dtail[t] = d;
    }
    // work items: one list per mid popcount jm (shared by every (js, jt) with js - jt == jm)
    std::vector<uint32_t> item_off(M + 1, 0);
    o.items.clear();
    o.dmid.clear();
    for (int jm = 0; jm <= M; ++jm) {
        item_off[jm] = (uint32_t)o.items.size();
        for (uint32_t u = 0; u < midcfg[jm].size(); ++u) {
            const unsigned c = midcfg[jm][u];
            SdBlkItem it;
            std::memset(&it, 0xFF, sizeof(it));
            it.c = (uint16_t)c;
            it.u2x = o.urank[c ^ (1u << (M - 1))];
            unsigned amask = 0;                                // active mid bonds (body variant 1 iterates these)
            for (int pm = 0; pm + 1 < M; ++pm) {
                const unsigned b0 = (c >> pm) & 1u, b1 = (c >> (pm + 1)) & 1u;
                if (b0 != b1 && Jhop[A + pm] != 0.0) {
                    const unsigned u2 = o.urank[c ^ (3u << pm)];
                    if (u2 >= 0xFFu) return false;
                    it.nb[pm] = (uint8_t)u2;
                    amask |= 1u << pm;
                }
            }
            static_assert(SD_BLK_M - 1 <= 9, "nb[0..8] partners, nb[9..10] active mask");
            it.nb[9] = (uint8_t)(amask & 0xFFu);
            it.nb[10] = (uint8_t)(amask >> 8);
            o.items.push_back(it);
            o.dmid.push_back(dmid_cfg[c]);
        }
    }
    // per-js class layout
    o.js.assign(B + 1, SdBlkJs());
    o.units.assign((size_t)2 * (B + 1) * SD_BLK_MAXUNITS, 0);
    uint32_t cap = 0;
    for (int js = 0; js <= B; ++js) {
        SdBlkJs &I = o.js[js];
        std::memset(&I, 0, sizeof(I));
        I.size = (uint32_t)C[B * SD_BINOM_DIM + js];
        uint32_t run = 0;
        for (int jt = 0; jt <= T; ++jt) {
            SdBlkCls &c = I.cls[jt];
            const int jm = js - jt;
            c.cb = run;
            if (jm < 0 || jm > M) continue;
            c.nblk = (uint32_t)C[M * SD_BINOM_DIM + jm];
            c.pitch = (c.nblk + 3u) & ~3u;
            c.n1 = jm >= 1 ? (uint32_t)C[(M - 1) * SD_BINOM_DIM + jm - 1] : 0u;
            c.item_off = item_off[jm];
            run += c.pitch * (uint32_t)C[T * SD_BINOM_DIM + jt];
        }
        I.size_pad = (run + 15u) & ~15u;
        cap = std::max(cap, I.size_pad);
        for (int w = 0; w < 2; ++w) {                         // item lists: w = 0 f64, 1 c128 (units of 32 mid configurations)
            const uint32_t uw = 32u;
            std::vector<std::pair<int, uint16_t>> list;       // (-NT, code): heavy classes first
            for (int jt = 0; jt <= T; ++jt) {
                const SdBlkCls &c = I.cls[jt];
                const uint32_t nu = (c.pitch + uw - 1) / uw;
                const int nt = (int)C[T * SD_BINOM_DIM + jt];
                const int nchunk = (w == 1 && nt > 5) ? 2 : 1;    // sd_blk_dispatch: c128, NT = 10 -> two chunks of 5
                for (uint32_t j = 0; j < nu; ++j)
                    for (int ch = 0; ch < nchunk; ++ch)
                        list.push_back({-(nt / nchunk), (uint16_t)((jt << 12) | (ch << 8) | j)});
            }
            std::stable_sort(list.begin(), list.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
            if (list.size() > SD_BLK_MAXUNITS) return false;
            I.nunits[w] = (uint32_t)list.size();
            for (size_t i = 0; i < list.size(); ++i)
                o.units[((size_t)w * (B + 1) + js) * SD_BLK_MAXUNITS + i] = list[i].second;
        }
    }
    P.cap = cap;
    // W[q][below]: stored elements of every tile whose prefix has a 1 at q after `below` ones
    auto padsz = [&](int js) -> uint64_t { return (js < 0 || js > B) ? 0 : o.js[js].size_pad; };
    o.W.assign((size_t)A * (A + 1), 0);
    for (int q = 0; q < A; ++q)
        for (int below = 0; below <= q; ++below) {
            uint64_t w = 0;
            const int n = A - 1 - q;
            for (int x = 0; x <= n; ++x) w += C[n * SD_BINOM_DIM + x] * padsz(k - below - 1 - x);
            o.W[(size_t)q * (A + 1) + below] = w;
        }
    {   // total stored elements = sum over prefix popcounts
        uint64_t tot = 0;
        for (int x = 0; x <= A; ++x) tot += C[A * SD_BINOM_DIM + x] * padsz(k - x);
        o.n_store = tot;
    }
    P.key_lo = 0; P.key_hi = 1ULL << A;
    P.shards.world = 1; P.shards.rank = 0;
    P.shards.pstart[0] = 0;
    for (int g = 1; g <= SD_MAX_WORLD; ++g) P.shards.pstart[g] = o.n_store;
    return true;
}

// stored-element offset of the tile with prefix bits Pb (valid or not: invalid tiles have size 0)
static inline uint64_t sd_blk_tile_base(const SdBlkHost &o, uint64_t Pb) {
    const int A = o.P.A;
    uint64_t b = 0;
    for (int q = 0; q < A; ++q) {
        if ((Pb >> q) & 1ULL) continue;
        const int below = __builtin_popcountll(Pb & ((1ULL << q) - 1ULL));
        b += o.W[(size_t)q * (A + 1) + below];
    }
    return b;
}
// stored-element offset of tile `key` (key = 2^A: the end)
static inline uint64_t sd_blk_key_base(const SdBlkHost &o, uint64_t key) {
    if (key >= (1ULL << o.P.A)) return o.n_store;
    return sd_blk_tile_base(o, sd_blk_prefix_bits(key, o.P.A));
}
// stored-element offset of basis state s (popcount k) in a vector of nc components per element
static inline uint64_t sd_blk_pos_of_state(const SdBlkHost &o, uint64_t s, int nc) {
    constexpr int M = SD_BLK_M, T = SD_BLK_T;
    const int A = o.P.A;
    const uint64_t Pb = s & ((1ULL << A) - 1ULL);
    const unsigned c = (unsigned)((s >> A) & ((1u << M) - 1u));
    const unsigned tau = (unsigned)((s >> (A + M)) & ((1u << T) - 1u));
    const int jt = __builtin_popcount(tau);
    const int js = o.P.k - __builtin_popcountll(Pb);
    const SdBlkCls &cl = o.js[js].cls[jt];
    const uint32_t e = (uint32_t)sd_tail_rank(T, jt, tau);
    const uint32_t u = o.urank[c];
    const uint32_t nt = (uint32_t)sd_cbinom(T, jt);
    if (nc == 1 && (nt & 1u) && e == nt - 1u) return sd_blk_tile_base(o, Pb) + cl.cb + (uint64_t)e * cl.pitch + u;
    if (nc == 1) return sd_blk_tile_base(o, Pb) + cl.cb + (uint64_t)(e >> 1) * 2u * cl.pitch + 2u * u + (e & 1u);
    return sd_blk_tile_base(o, Pb) + cl.cb + (uint64_t)e * cl.pitch + u;
}

// L2-friendly tile order of the keys [key_lo, key_hi): breadth-first (Cuthill-McKee) order of each popcount group of
// the adjacent-swap graph on the top e prefix sites (the slow index; the remaining prefix sites run in rank order).
// Every bond partner of a configuration then lies in the same or a neighbouring BFS level, i.e. within about two
// level widths of the traversal, so most partner tiles are still in L2 when they are needed.  Model (scripts/l2_sim.py,
// 63 MB LRU, e = 16): 13.5 GB of DRAM reads per L = 32 apply against 22.8 GB in rank order; measured on B200: 14.7 GB
// against 22.4 GB (profiles/round2_a_ab.txt).  Any order is correct (tiles are independent); only valid tiles are listed.
static inline void sd_blk_tile_order(const SdBlkHost &o, uint64_t key_lo, uint64_t key_hi, int e, std::vector<uint32_t> &out) {
    const int A = o.P.A, k = o.P.k;
    if (e > A) e = A;
    if (e > 24) e = 24;
    if (e < 1) e = 1;
    const unsigned ne = 1u << e;
    auto lex = [&](unsigned c) {                                    // "1 first" lexicographic key of a top configuration
        unsigned v = 0;
        for (int q = 0; q < e; ++q) v = (v << 1) | (((c >> q) & 1u) ? 0u : 1u);
        return v;
    };
    std::vector<uint32_t> slow(ne, 0);                              // visiting position of a top configuration
    uint32_t next_pos = 0;
    for (int p = e; p >= 0; --p) {
        std::vector<unsigned> cfgs;
        for (unsigned c = 0; c < ne; ++c) if (__builtin_popcount(c) == p) cfgs.push_back(c);
        std::sort(cfgs.begin(), cfgs.end(), [&](unsigned a, unsigned b) { return lex(a) < lex(b); });
        std::vector<unsigned char> seen(ne, 0);
        std::vector<unsigned> queue;
        queue.reserve(cfgs.size());
        for (unsigned start : cfgs) {
            if (seen[start]) continue;
            size_t head = queue.size();
            queue.push_back(start); seen[start] = 1;
            while (head < queue.size()) {
                const unsigned c = queue[head++];
                slow[c] = next_pos++;
                for (int q = e - 2; q >= 0; --q)
                    if (((c >> q) ^ (c >> (q + 1))) & 1u) {
                        const unsigned n = c ^ (3u << q);
                        if (!seen[n]) { seen[n] = 1; queue.push_back(n); }
                    }
            }
        }
    }
    std::vector<std::pair<uint64_t, uint32_t>> keyed;
    keyed.reserve((size_t)(key_hi - key_lo));
    for (uint64_t key = key_lo; key < key_hi; ++key) {
        const uint64_t Pb = sd_blk_prefix_bits(key, A);
        const int js = k - __builtin_popcountll(Pb);
        if (js < 0 || js > SD_BLK_B) continue;
        const unsigned top = (unsigned)(Pb & (ne - 1u));
        const uint64_t low = key & ((1ULL << (A - e)) - 1ULL);       // the low prefix sites are the low key bits (rank order)
        keyed.push_back({((uint64_t)slow[top] << (A - e)) | low, (uint32_t)key});
    }
    std::sort(keyed.begin(), keyed.end());
    out.resize(keyed.size());
    for (size_t i = 0; i < keyed.size(); ++i) out[i] = keyed[i].second;
}
