// emul_blk.cpp -- CPU emulation of the block-layout CUDA kernel (sd_blk.h).
//
// TEST INFRASTRUCTURE ONLY.  The product (libspindyn_cuda.so) never loads this; it exists so
// `pytest -m "not gpu"` can check, in a container without a GPU, everything of the block kernel
// that is arithmetic rather than plumbing: the padded block layout (f64 pair rows / c128 rows), the
// tile header (tile bases, neighbour-tile pointers, owner GPU of a neighbour, crossing partner), the
// item body (prefix streams, tail/mid/crossing hops, diagonal, fused epilogues and their reductions)
// and the tile-aligned sharding.  It runs the SAME __host__ __device__ functions the kernel runs
// (sd_blk_hdr_lane / sd_blk_hdr_fill / sd_blk_dispatch), lane by lane, one tile at a time; what it does
// not cover is the TMA/mbarrier producer-consumer pipeline, which moves bytes but computes nothing.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../spindynamics.jl_b200/csrc/sd_tile_host.h"
#include "../../spindynamics.jl_b200/csrc/sd_blk_host.h"
#include "../../spindynamics.jl_b200/csrc/sd_blkl.h"
#include "../../spindynamics.jl_b200/csrc/sd_shard_host.h"

namespace {

struct AlignedBuf {                      // 128-byte aligned doubles (the kernel moves 16-byte slots)
    double *p = nullptr;
    size_t n = 0;
    void alloc(size_t count, double fill) {
        n = count;
        void *q = nullptr;
        if (posix_memalign(&q, 128, (count + 16) * sizeof(double)) != 0) q = nullptr;
        p = (double *)q;
        for (size_t i = 0; i < count + 16; ++i) p[i] = fill;
    }
    ~AlignedBuf() { free(p); }
};

template <int NC, int EK, bool WRAP = false>
void run_tiles(const SdBlkHost &bh, const SdBlkParams &P, const SdVecView &psi, double *out_local, const SdEpi &epi,
               int qfar, double *red_total) {
    AlignedBuf tile;
    tile.alloc((size_t)P.cap * NC, NAN);
    // tables + context of the kernel's static shared-memory block SD_SH
    std::memcpy(SD_SH.js, bh.js.data(), sizeof(SdBlkJs) * (SD_BLK_B + 1));
    std::memcpy(SD_SH.dmid, bh.dmid.data(), sizeof(double) << SD_BLK_M);
    sd_blkl_ctx_init(epi);
    for (uint64_t key = P.key_lo; key < P.key_hi; ++key) {
        const uint64_t Pb = sd_blk_key_prefix(key, P.A);
        const int js = P.k - SD_POPC64(Pb);
        if (js < 0 || js > SD_BLK_B) continue;                     // impossible suffix popcount
        // ---- header: what the producer warp computes with shuffles
        SdBlkHdr H;
        std::memset(&H, 0, sizeof(H));
        for (SdBlkEnt &e : H.nb) { e.p = nullptr; e.J = NAN; }       // entries the header does not write must never be used
        sd_blk_hdr_host<NC, WRAP>(P, bh.W.data(), key, qfar, psi, H);
        // pipeline overrun entries: valid pointer, J = 0 (the kernel never dereferences them: ok_ is false)
        // ---- own tile -> "shared memory" (the TMA bulk copy); the rest of the buffer stays NaN
        const uint32_t size_pad = bh.js[H.js].size_pad;
        for (size_t i = 0; i < (size_t)P.cap * NC; ++i) tile.p[i] = NAN;
        std::memcpy(tile.p, psi.base[P.shards.rank] + (size_t)NC * H.base, (size_t)size_pad * NC * sizeof(double));
        // ---- work items in list order, 32 lanes each
        const unsigned nunits = bh.js[H.js].nunits[NC - 1];
        const uint16_t *ut = bh.units.data() + ((size_t)(NC - 1) * (SD_BLK_B + 1) + H.js) * SD_BLK_MAXUNITS;
        for (unsigned un = 0; un < nunits; ++un) {
            const unsigned code = ut[un];
            for (unsigned lane = 0; lane < 32; ++lane) {
                const uint32_t u = (code & 0xFFu) * 32u + lane;
                double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
                sd_blkl_dispatch<NC, EK, WRAP>(P, epi, out_local, H, tile.p, code, u, red);
                for (int s = 0; s < SD_NSLOT; ++s) red_total[s] += red[s];
            }
        }
    }
}

}  // namespace

static double g_wrap_J = 0.0, g_wrap_Jz = 0.0;                      // periodic wrap bond of the next emul_blk_apply calls

extern "C" {

// hop coefficient and Jz of the wrap bond (sites L-1, 0); (0, 0) = open chain
void emul_blk_set_wrap(double J, double Jz) { g_wrap_J = J; g_wrap_Jz = Jz; }

// Runs rank `rank` of `world` of the block kernel over full-length RANK-ORDERED host vectors (the
// layout conversion the library does in sd_vec_upload/download is done here with sd_blk_pos_of_state).
// states[N]: the basis in rank order (from the oracle).  psi/out (and vprev/phi/acc) hold NC doubles per
// state.  out/acc are written for the states of this rank's shard only.
// Returns 0; -1 model does not qualify; -2 a padding element of out became nonzero; -3 a position
// collision / out-of-range position in the layout map.
int emul_blk_apply(int L, int k, const double *Jhop, const double *Jz, const double *h, int NC,
                   const uint64_t *states, uint64_t N, const double *psi, double *out,
                   int world, int rank, int mode, int redmask, double hscale, double a, double b,
                   const double *vprev, const double *phi, double *acc, double ck_re, double ck_im,
                   double *red_out, uint64_t *bounds_out, uint64_t far_bytes, uint64_t *n_store_out, int variant) {
    SdBlkHost bh;
    if (!sd_blk_build(L, k, Jhop, Jz, h, bh)) return -1;
    SdTileHost th;                                                  // shard bounds come from the tiled split (same tile keys)
    if (!sd_tile_build(L, k, SD_BLK_B, 5, Jhop, Jz, h, th)) return -1;
    uint64_t bounds[SD_MAX_WORLD + 1], keys[SD_MAX_WORLD + 1];
    sd_tile_shard_bounds(th, world, bounds, keys);
    if (variant & 512) {
        double cost[2];
        if (!sd_shard_balance(bh, th, world, sd_tile_qfar(L, bh.P.A, bh.binom.data(), far_bytes, 8 * NC), 0.7, 5, bounds, keys, cost)) return -9;
    }
    SdBlkParams P = bh.P;
    P.W = bh.W.data(); P.js = bh.js.data(); P.units = bh.units.data(); P.items = bh.items.data(); P.dmid = bh.dmid.data();
    P.nbuf = 3;
    const bool halo = (variant & 256) != 0;                          // + 256: through the halo mirror (sd_shard_host.h), 3 chunks
    // + 512: remote-volume-weighted shard bounds (sd_shard_balance), applied above where the bounds are computed
    P.key_lo = keys[rank]; P.key_hi = keys[rank + 1];
    P.shards.world = world; P.shards.rank = rank;
    uint64_t pstart[SD_MAX_WORLD + 1];
    for (int g = 0; g <= world; ++g) pstart[g] = sd_blk_key_base(bh, keys[g]);
    for (int g = 0; g <= SD_MAX_WORLD; ++g) P.shards.pstart[g] = pstart[g < world ? g : world];
    if (bounds_out) for (int g = 0; g <= world; ++g) bounds_out[g] = bounds[g];
    if (n_store_out) *n_store_out = bh.n_store;
    const int qfar = sd_tile_qfar(L, P.A, bh.binom.data(), far_bytes, 8 * NC);

    // ---- layout map: stored position of every basis state, checked to be a bijection onto non-padding slots
    std::vector<uint64_t> pos(N);
    {
        std::vector<unsigned char> used(bh.n_store, 0);
        for (uint64_t r = 0; r < N; ++r) {
            const uint64_t p = sd_blk_pos_of_state(bh, states[r], NC);
            if (p >= bh.n_store || used[p]) return -3;
            used[p] = 1;
            pos[r] = p;
        }
    }
    auto to_blk = [&](const double *src, std::vector<AlignedBuf> &shard, SdVecView *view) {
        shard.clear();
        shard.resize(world);
        for (int g = 0; g < world; ++g) {
            shard[g].alloc((size_t)(pstart[g + 1] - pstart[g]) * NC, 0.0);
            if (view) view->base[g] = shard[g].p - (int64_t)pstart[g] * NC;
        }
        for (uint64_t r = 0; r < N; ++r) {
            int g = 0;
            while (g + 1 < world && pos[r] >= pstart[g + 1]) ++g;
            for (int c = 0; c < NC; ++c) shard[g].p[(pos[r] - pstart[g]) * NC + c] = src[r * NC + c];
        }
    };
    SdVecView view;
    for (int g = 0; g < SD_MAX_WORLD; ++g) view.base[g] = nullptr;
    std::vector<AlignedBuf> s_psi, s_prev, s_phi, s_acc;
    to_blk(psi, s_psi, &view);
    const size_t nloc = (size_t)(pstart[rank + 1] - pstart[rank]) * NC;
    AlignedBuf o;
    o.alloc(nloc, 0.0);                                             // sd_vec_alloc zero-fills block-layout vectors
    SdEpi epi;
    std::memset(&epi, 0, sizeof(epi));
    epi.mode = mode; epi.red = redmask; epi.hscale = hscale; epi.a = a; epi.b = b;
    epi.ck_re = ck_re; epi.ck_im = ck_im;
    if (vprev) { to_blk(vprev, s_prev, nullptr); epi.vprev = s_prev[rank].p; }
    if (phi) { to_blk(phi, s_phi, nullptr); epi.phi = s_phi[rank].p; }
    if (acc) { to_blk(acc, s_acc, nullptr); epi.acc = s_acc[rank].p; }
    double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
    const bool plain = epi.mode == SD_EPI_PLAIN && epi.red == 0 && !epi.acc && epi.hscale == 1.0;
#define RUN(NC_, EK_)                                                                         \
    do {                                                                                      \
        if (P.wrap_on) run_tiles<NC_, EK_, true>(bh, P, view, o.p, epi, qfar, red);   /* as sd_blk_launch_range */ \
        else run_tiles<NC_, EK_>(bh, P, view, o.p, epi, qfar, red);                           \
    } while (0)
    // halo mirror: the peers' shards are replaced by NaN-filled mirrors that only hold what the plan copies, chunk by
    // chunk, before the tiles of that chunk run (what sd_apply_blk_halo does with the copy engines and one event per chunk)
    SdShardPlan plan;
    std::vector<AlignedBuf> mirror(world);
    int nchunks = 1;
    if (halo && world > 1) {
        nchunks = 3;
        if (!sd_shard_plan(bh, P, nchunks, qfar, plan)) return -9;
        for (int g = 0; g < world; ++g) {
            if (g == rank) continue;
            mirror[g].alloc((size_t)(pstart[g + 1] - pstart[g]) * NC, NAN);
            view.base[g] = mirror[g].p - (int64_t)pstart[g] * NC;
        }
    }
    // periodic chain: the WRAP variant of the item body; the header carries the wrap partner tile
    if (g_wrap_J != 0.0 || g_wrap_Jz != 0.0) { P.wrap_on = 1; P.wrapJ = g_wrap_J; P.wrapJz4 = 0.25 * g_wrap_Jz; }
    const uint64_t klo_all = P.key_lo, khi_all = P.key_hi;
    for (int j = 0; j < nchunks; ++j) {
        if (halo && world > 1) {
            for (const SdShardSeg &sg : plan.segs[j])
                std::memcpy(mirror[sg.peer].p + (sg.lo - pstart[sg.peer]) * NC, s_psi[sg.peer].p + (sg.lo - pstart[sg.peer]) * NC,
                            (size_t)(sg.hi - sg.lo) * NC * sizeof(double));
            P.key_lo = plan.chunk_key[j]; P.key_hi = plan.chunk_key[j + 1];
        }
        const int ek = plain ? 0 : ((epi.mode == SD_EPI_PLAIN && (epi.red == SD_RED_DOT_SELF || epi.red == 0) && !epi.acc) ? 1 : 2);   // as sd_blk_launch_range
        if (NC == 1) { if (ek == 0) RUN(1, 0); else if (ek == 1) RUN(1, 1); else RUN(1, 2); }
        else { if (ek == 0) RUN(2, 0); else if (ek == 1) RUN(2, 1); else RUN(2, 2); }
    }
    P.key_lo = klo_all; P.key_hi = khi_all;
#undef RUN
    if (red_out) for (int s = 0; s < SD_NSLOT; ++s) red_out[s] = red[s];
    // ---- back to rank order; padding must still be zero
    std::vector<unsigned char> real(nloc / NC, 0);
    for (uint64_t r = bounds[rank]; r < bounds[rank + 1]; ++r) {
        const uint64_t lp = pos[r] - pstart[rank];
        real[lp] = 1;
        for (int c = 0; c < NC; ++c) {
            out[r * NC + c] = o.p[lp * NC + c];
            if (acc) acc[r * NC + c] = s_acc[rank].p[lp * NC + c];
        }
    }
    for (size_t i = 0; i < nloc / NC; ++i)
        if (!real[i])
            for (int c = 0; c < NC; ++c)
                if (o.p[i * NC + c] != 0.0) return -2;
    return 0;
}

// Host-side plan of the sharded block layout at ANY size (no vectors are touched): rank bounds, stored-
// element bounds and tile-key bounds per rank, total stored elements, largest tile.  Returns 0 or -1.
int emul_blk_plan(int L, int k, int world, uint64_t *bounds, uint64_t *pstart, uint64_t *keys, uint64_t *n_store,
                  uint32_t *cap, uint64_t *n_tiles) {
    SdBlkHost bh;
    std::vector<double> J(L, 0.5), Jz(L, 1.0), h(L, 0.0);
    if (!sd_blk_build(L, k, J.data(), Jz.data(), h.data(), bh)) return -1;
    SdTileHost th;
    if (!sd_tile_build(L, k, SD_BLK_B, 5, J.data(), Jz.data(), h.data(), th)) return -1;
    sd_tile_shard_bounds(th, world, bounds, keys);
    for (int g = 0; g <= world; ++g) pstart[g] = sd_blk_key_base(bh, keys[g]);
    *n_store = bh.n_store;
    *cap = bh.P.cap;
    uint64_t nt = 0;                                               // tiles with a possible suffix popcount
    for (int x = 0; x <= bh.P.A; ++x)
        if (k - x >= 0 && k - x <= SD_BLK_B) nt += bh.binom[(size_t)bh.P.A * SD_BINOM_DIM + x];
    *n_tiles = nt;
    return 0;
}

// Halo-mirror plan of rank `rank` (sd_shard_host.h) at any size, checked: every remote partner tile of every tile header
// of chunk j lies inside the segments of chunks 0..j; segments are disjoint, inside their peer's shard and tile aligned
// (multiples of 16 elements).  stats: [0] segments, [1] remote stored elements, [2] local stored elements, [3] peers used,
// [4] largest number of segments in one chunk.  Returns 0, -1 (model), -2 (plan failed), -3 (coverage), -4 (segment shape).
int emul_halo_plan(int L, int k, int world, int rank, int nchunks, uint64_t *stats) {
    SdBlkHost bh;
    std::vector<double> J(L, 0.5), Jz(L, 1.0), h(L, 0.0);
    if (!sd_blk_build(L, k, J.data(), Jz.data(), h.data(), bh)) return -1;
    SdTileHost th;
    if (!sd_tile_build(L, k, SD_BLK_B, 5, J.data(), Jz.data(), h.data(), th)) return -1;
    uint64_t bounds[SD_MAX_WORLD + 1], keys[SD_MAX_WORLD + 1];
    sd_tile_shard_bounds(th, world, bounds, keys);
    SdBlkParams P = bh.P;
    P.W = bh.W.data(); P.js = bh.js.data(); P.units = bh.units.data(); P.items = bh.items.data(); P.dmid = bh.dmid.data();
    P.key_lo = keys[rank]; P.key_hi = keys[rank + 1];
    P.shards.world = world; P.shards.rank = rank;
    for (int g = 0; g <= SD_MAX_WORLD; ++g) P.shards.pstart[g] = sd_blk_key_base(bh, keys[g < world ? g : world]);
    const int qfar = sd_tile_qfar(L, P.A, bh.binom.data(), (uint64_t)100 << 20, 8);
    SdShardPlan plan;
    if (!sd_shard_plan(bh, P, nchunks, qfar, plan)) return -2;
    if ((int)plan.chunk_key.size() != nchunks + 1 || plan.chunk_key.front() != P.key_lo || plan.chunk_key.back() != P.key_hi) return -2;
    std::vector<std::pair<uint64_t, uint64_t>> have[SD_MAX_WORLD];
    uint64_t nseg = 0, maxseg = 0, total = 0;
    for (int j = 0; j < nchunks; ++j) {
        if (plan.chunk_key[j] > plan.chunk_key[j + 1]) return -2;
        maxseg = std::max<uint64_t>(maxseg, plan.segs[j].size());
        for (const SdShardSeg &s : plan.segs[j]) {
            if (s.peer == rank || s.peer < 0 || s.peer >= world || s.lo >= s.hi || (s.lo & 15u) || (s.hi & 15u)) return -4;
            if (s.lo < P.shards.pstart[s.peer] || s.hi > P.shards.pstart[s.peer + 1]) return -4;
            for (auto [lo, hi] : have[s.peer]) if (s.lo < hi && lo < s.hi) return -4;             // disjoint from everything before
            have[s.peer].push_back({s.lo, s.hi});
            ++nseg; total += s.hi - s.lo;
        }
        for (int g = 0; g < world; ++g) sd_shard_merge(have[g]);
        std::vector<SdShardSeg> raw;
        for (uint64_t key = plan.chunk_key[j]; key < plan.chunk_key[j + 1]; ++key) sd_shard_tile_remotes(bh, P, key, qfar, raw);
        for (const SdShardSeg &r : raw) {
            bool ok = false;
            for (auto [lo, hi] : have[r.peer]) if (lo <= r.lo && r.hi <= hi) { ok = true; break; }
            if (!ok) return -3;
        }
    }
    if (total != plan.remote_elems) return -4;
    uint64_t peers = 0;
    for (int g = 0; g < world; ++g) {
        if (have[g] != plan.need[g]) return -4;
        if (!have[g].empty()) ++peers;
    }
    stats[0] = nseg; stats[1] = total; stats[2] = P.shards.pstart[rank + 1] - P.shards.pstart[rank]; stats[3] = peers; stats[4] = maxseg;
    stats[5] = 0;
    return 0;
}

// Remote-volume-weighted shard bounds (sd_shard_balance): bounds / keys [world + 1], cost[2] (largest per-rank time with
// equal shards / with the returned bounds, in local-element units), per_rank[world] times at the returned bounds.
int emul_halo_balance(int L, int k, int world, double remote_cost, int iters, uint64_t *bounds, uint64_t *keys, double *cost, double *per_rank) {
    SdBlkHost bh;
    std::vector<double> J(L, 0.5), Jz(L, 1.0), h(L, 0.0);
    if (!sd_blk_build(L, k, J.data(), Jz.data(), h.data(), bh)) return -1;
    SdTileHost th;
    if (!sd_tile_build(L, k, SD_BLK_B, 5, J.data(), Jz.data(), h.data(), th)) return -1;
    const int qfar = sd_tile_qfar(L, bh.P.A, bh.binom.data(), (uint64_t)100 << 20, 8);
    if (!sd_shard_balance(bh, th, world, qfar, remote_cost, iters, bounds, keys, cost)) return -2;
    double tmax = 0.0;
    std::vector<double> t;
    if (!sd_shard_rank_cost(bh, keys, world, qfar, remote_cost, &tmax, &t)) return -2;
    for (int g = 0; g < world; ++g) per_rank[g] = t[g];
    return 0;
}

// The optional L2-friendly tile order of rank `rank` (sd_blk_tile_order).  Returns the number of keys written
// (<= cap), or -1.
long emul_blk_order(int L, int k, int world, int rank, int e, uint32_t *out, long cap, uint64_t *key_lo, uint64_t *key_hi) {
    SdBlkHost bh;
    std::vector<double> J(L, 0.5), Jz(L, 1.0), h(L, 0.0);
    if (!sd_blk_build(L, k, J.data(), Jz.data(), h.data(), bh)) return -1;
    SdTileHost th;
    if (!sd_tile_build(L, k, SD_BLK_B, 5, J.data(), Jz.data(), h.data(), th)) return -1;
    uint64_t bounds[SD_MAX_WORLD + 1], keys[SD_MAX_WORLD + 1];
    sd_tile_shard_bounds(th, world, bounds, keys);
    std::vector<uint32_t> ord;
    sd_blk_tile_order(bh, keys[rank], keys[rank + 1], e, ord);
    if ((long)ord.size() > cap) return -1;
    for (size_t i = 0; i < ord.size(); ++i) out[i] = ord[i];
    *key_lo = keys[rank]; *key_hi = keys[rank + 1];
    return (long)ord.size();
}

}  // extern "C"
