// sd_shard_host.h -- host-side model of a sharded block-layout apply's NVLink traffic, and shards weighted by it.
//
// A sharded apply reads the partner tiles of the cut prefix bonds straight from peer HBM inside the kernel (plain 16-byte
// loads through CUDA-IPC mappings).  With equal rank ranges the inbound volume is very uneven: at 8 ranks the ranks whose
// top prefix bits are 101 / 010 gather 2.5 shards' worth, the edge ranks 0.5.  This file computes, from the
// same tile-header code the kernel runs (sd_blk_hdr_host), which peer ranges a rank's tiles point at, and moves the cut
// positions until the slowest rank's modelled time stops improving (sd_shard_balance).  Measured at L = 32 on 8 B200:
// 2.62 ms per apply with equal shards, 1.81 ms weighted (profiles/round2_e_8gpu.txt).
// (Round 2 also measured the copy-engine "halo mirror" this plan was first written for -- chunked cudaMemcpyAsync of the
// peer ranges into a local sparse mapping, kernels on local pointers only: 4.47 ms against 3.43 ms for direct peer loads
// at 2 ranks, so it was removed.)  Pure C++ (no CUDA); checked on the CPU by tests/emul.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>
#include "sd_blk_host.h"
#include "sd_tile_host.h"

struct SdShardSeg {
    int peer;
    uint64_t lo, hi;                 // stored elements [lo, hi), global offsets (all shards)
};
struct SdShardPlan {
    std::vector<uint64_t> chunk_key;                                    // K + 1 tile-key bounds of the rank's range
    std::vector<std::vector<SdShardSeg>> segs;                           // per chunk: ranges not yet copied by earlier chunks
    std::vector<std::pair<uint64_t, uint64_t>> need[SD_MAX_WORLD];      // per peer: union of all ranges, merged, sorted
    uint64_t remote_elems = 0;                                          // sum of the ranges (stored elements)
};

// the remote partner tiles of tile `key`: (peer, [lo, hi)) appended to out.  Runs the kernel's header code with base
// pointers that encode the owner in the top byte, so the owner / offset of every entry is exactly what the kernel derefs.
static inline void sd_shard_tile_remotes(const SdBlkHost &bh, const SdBlkParams &P, uint64_t key, int qfar,
                                        std::vector<SdShardSeg> &out) {
    const uint64_t Pb = sd_blk_key_prefix(key, P.A);
    const int js = P.k - SD_POPC64(Pb);
    if (js < 0 || js > SD_BLK_B) return;
    SdVecView fake;
    for (int g = 0; g < SD_MAX_WORLD; ++g) fake.base[g] = (const double *)(uintptr_t)((uint64_t)(g + 1) << 56);
    SdBlkHdr H;
    std::memset(&H, 0, sizeof(H));
    sd_blk_hdr_host<1>(P, bh.W.data(), key, qfar, fake, H);
    for (int n = 0; n < H.ntot; ++n) {
        const uint64_t v = (uint64_t)(uintptr_t)H.nb[n].p;
        const int g = (int)(v >> 56) - 1;
        const uint64_t nbase = (v & ((1ULL << 56) - 1ULL)) / sizeof(double);
        if (g == P.shards.rank) continue;
        const uint32_t sz = bh.js[n == H.nnb ? H.jsx : js].size_pad;       // entry nnb (if any) is the crossing partner
        out.push_back({g, nbase, nbase + sz});
    }
}

static inline void sd_shard_merge(std::vector<std::pair<uint64_t, uint64_t>> &v) {
    std::sort(v.begin(), v.end());
    size_t w = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        if (w > 0 && v[i].first <= v[w - 1].second) v[w - 1].second = std::max(v[w - 1].second, v[i].second);
        else v[w++] = v[i];
    }
    v.resize(w);
}
// a minus b (both merged and sorted)
static inline std::vector<std::pair<uint64_t, uint64_t>> sd_shard_subtract(const std::vector<std::pair<uint64_t, uint64_t>> &a,
                                                                         const std::vector<std::pair<uint64_t, uint64_t>> &b) {
    std::vector<std::pair<uint64_t, uint64_t>> r;
    size_t j = 0;
    for (auto [lo, hi] : a) {
        while (j < b.size() && b[j].second <= lo) ++j;
        uint64_t cur = lo;
        for (size_t t = j; t < b.size() && b[t].first < hi; ++t) {
            if (b[t].first > cur) r.push_back({cur, b[t].first});
            cur = std::max(cur, b[t].second);
            if (cur >= hi) break;
        }
        if (cur < hi) r.push_back({cur, hi});
    }
    return r;
}

// P: the rank's launch parameters (shards, key_lo / key_hi, host table pointers set).  nchunks >= 1.
static inline bool sd_shard_plan(const SdBlkHost &bh, const SdBlkParams &P, int nchunks, int qfar, SdShardPlan &out) {
    out = SdShardPlan();
    if (nchunks < 1) nchunks = 1;
    const uint64_t klo = P.key_lo, khi = P.key_hi;
    // chunk bounds: equal shares of the rank's stored elements, moved to tile keys (the stored base is monotone in the key)
    const uint64_t b0 = sd_blk_key_base(bh, klo), b1 = sd_blk_key_base(bh, khi);
    out.chunk_key.assign(nchunks + 1, khi);
    out.chunk_key[0] = klo;
    for (int j = 1; j < nchunks; ++j) {
        const uint64_t target = b0 + (uint64_t)(((unsigned __int128)(b1 - b0) * j) / nchunks);
        uint64_t lo = klo, hi = khi;                                   // smallest key with base >= target
        while (lo < hi) {
            const uint64_t mid = lo + (hi - lo) / 2;
            if (sd_blk_key_base(bh, mid) >= target) hi = mid; else lo = mid + 1;
        }
        out.chunk_key[j] = std::max(lo, out.chunk_key[j - 1]);
    }
    out.segs.assign(nchunks, {});
    std::vector<std::pair<uint64_t, uint64_t>> have[SD_MAX_WORLD];
    std::vector<SdShardSeg> raw;
    for (int j = 0; j < nchunks; ++j) {
        raw.clear();
        for (uint64_t key = out.chunk_key[j]; key < out.chunk_key[j + 1]; ++key) sd_shard_tile_remotes(bh, P, key, qfar, raw);
        for (int g = 0; g < P.shards.world; ++g) {
            std::vector<std::pair<uint64_t, uint64_t>> want;
            for (const SdShardSeg &s : raw) if (s.peer == g) want.push_back({s.lo, s.hi});
            if (want.empty()) continue;
            sd_shard_merge(want);
            for (auto [lo, hi] : sd_shard_subtract(want, have[g])) {
                if (lo < P.shards.pstart[g] || hi > P.shards.pstart[g + 1]) return false;   // a tile never straddles shards
                out.segs[j].push_back({g, lo, hi});
                out.remote_elems += hi - lo;
            }
            have[g].insert(have[g].end(), want.begin(), want.end());
            sd_shard_merge(have[g]);
        }
    }
    for (int g = 0; g < SD_MAX_WORLD; ++g) out.need[g] = have[g];
    return true;
}

// ---- shards weighted by their remote volume (default for more than two ranks; SD_SHARD_BALANCE=0 switches it off)
// A rank's apply time is modelled as max(local elements, remote_cost * remote elements): remote_cost = time to receive
// one element over NVLink in units of the time to process one local element (default 1.3: 8 B at ~750 GB/s against
// ~8 ps per local state).
// This fixed-point iteration moves the cut positions until the largest such time stops improving; every rank runs
// it on the same inputs, so all ranks arrive at the same bounds.  cost[0] = largest time with equal shards,
// cost[1] = with the returned bounds (both in local-element units).
static inline bool sd_shard_rank_cost(const SdBlkHost &bh, const uint64_t *keys, int world, int qfar, double remote_cost,
                                     double *tmax, std::vector<double> *per_rank) {
    SdBlkParams P = bh.P;
    P.W = bh.W.data(); P.js = bh.js.data(); P.units = bh.units.data(); P.items = bh.items.data(); P.dmid = bh.dmid.data();
    P.shards.world = world;
    for (int g = 0; g <= SD_MAX_WORLD; ++g) P.shards.pstart[g] = sd_blk_key_base(bh, keys[g < world ? g : world]);
    *tmax = 0.0;
    if (per_rank) per_rank->assign(world, 0.0);
    for (int r = 0; r < world; ++r) {
        P.shards.rank = r; P.key_lo = keys[r]; P.key_hi = keys[r + 1];
        SdShardPlan plan;
        if (!sd_shard_plan(bh, P, 1, qfar, plan)) return false;
        const double local = (double)(P.shards.pstart[r + 1] - P.shards.pstart[r]);
        const double t = std::max(local, remote_cost * (double)plan.remote_elems);
        if (per_rank) (*per_rank)[r] = t;
        *tmax = std::max(*tmax, t);
    }
    return true;
}
static inline bool sd_shard_balance(const SdBlkHost &bh, const SdTileHost &th, int world, int qfar, double remote_cost, int iters,
                                   uint64_t *bounds, uint64_t *keys, double *cost) {
    const uint64_t N = th.binom[(size_t)th.P.L * SD_BINOM_DIM + th.P.k];
    std::vector<double> frac(world, 1.0 / world), t(world);
    std::vector<uint64_t> cum(world + 1), b(world + 1), kk(world + 1);
    double best = -1.0;
    for (int it = 0; it <= iters; ++it) {
        double acc = 0.0;
        cum[0] = 0;
        for (int g = 1; g < world; ++g) { acc += frac[g - 1]; cum[g] = std::max(cum[g - 1], (uint64_t)(acc * (double)N)); }
        cum[world] = N;
        sd_tile_shard_bounds_at(th, world, cum.data(), b.data(), kk.data());
        double tmax = 0.0;
        if (!sd_shard_rank_cost(bh, kk.data(), world, qfar, remote_cost, &tmax, &t)) return false;
        if (it == 0) cost[0] = tmax;
        if (best < 0.0 || tmax < best) {
            best = tmax;
            for (int g = 0; g <= world; ++g) { bounds[g] = b[g]; keys[g] = kk[g]; }
        }
        double mean = 0.0;
        for (int g = 0; g < world; ++g) mean += t[g] / world;
        double sum = 0.0;
        for (int g = 0; g < world; ++g) {                               // a slow rank gets a smaller share
            const double sz = std::max(frac[g], 1e-6);
            frac[g] = sz * std::pow(mean / std::max(t[g], 1e-30), 0.6);
            sum += frac[g];
        }
        for (int g = 0; g < world; ++g) frac[g] /= sum;
    }
    cost[1] = best;
    return true;
}
