// sd_blk.h -- the "block layout" H.psi kernel for the sector basis of an open
// nearest-neighbour chain (reference Hamiltonian.jl:211-273 restricted to the bond
// lists XXZChain builds, SpinModel.jl:63-90; arbitrary per-bond J/Jz, per-site field).
//
// Device vectors of a block-layout model are NOT stored in rank order.  The chain is
// cut into  prefix (A sites) | mid (M sites) | tail (T sites); a TILE is one prefix
// configuration (a contiguous rank range of the reference basis, Basis.jl:37-53).
// Inside a tile the suffix configurations are stored class-major (class jt = tail popcount);
// inside a class, with e = index of the tail configuration in its class (NT = C(T, jt) of them)
// and u = index of the mid configuration among those with js - jt set bits (row pitch = their
// number rounded up to 4):
//     c128:  element (e, u) at  cb + e*pitch + u                       (one 16-byte value)
//     f64 :  element (e, u) at  cb + (e/2)*2*pitch + 2u + (e & 1)      (pair rows: a 16-byte SLOT
//            holds tail configurations 2s and 2s+1 of one mid configuration); the last
//            configuration of an odd class is a plain row at  cb + (NT-1)*pitch + u  (half slot)
// so that both types move 16-byte slots at  cb*NC + s*2*pitch + 2u  doubles, lanes = consecutive u.
// Tiles are padded to a multiple of 16 elements; padding is zero and stays zero under every
// linear operation.  With this order
//   * a hop on a prefix bond maps a whole tile onto another whole tile, element by
//     element: one thread that owns mid configuration u streams  acc[e] += J psi'[e][u]
//     with fully coalesced 16-byte loads (lanes = consecutive u), no staging;
//   * the tile itself is copied verbatim into shared memory by one TMA bulk copy
//     (cp.async.bulk + mbarrier), so hops on suffix bonds are shared-memory gathers;
//   * tail-internal hops are register moves fixed at compile time;
//   * the result is stored straight from registers, again coalesced 16-byte stores.
// There is no flat phase, no permutation table and no CTA barrier in the steady state:
// consumer warps pull (tile, unit) work items from a per-tile counter, a producer warp
// computes tile headers and keeps NBUF tiles in flight.
// This file holds the layout, the tables, the tile header, the producer warp and the layout-conversion kernel;
// the apply kernel itself (consumer warps, item body) is sd_blkl.h.
//
// f64: a lane owns two adjacent mid configurations (a double2 = blocks u, u+1).
// c128: a lane owns one mid configuration (a double2 = re, im).  H is real, so both are
// "two independent real columns" for everything but the gather addresses.
#pragma once
#include "sd_common.h"

#define SD_BLK_M 10
#define SD_BLK_T 5
#define SD_BLK_B (SD_BLK_M + SD_BLK_T)
#define SD_BLK_NCLS (SD_BLK_T + 1)
#define SD_BLK_MAXA 32
#define SD_BLK_MAXUNITS 64      // work items per tile: (unit of 32 mid configurations, element chunk); <= 32 (f64), <= 47 (c128)
struct SdBlkCls {
    uint32_t cb;         // element offset of the class inside the tile
    uint32_t pitch;      // row pitch (elements) = nblk rounded up to 4
    uint32_t nblk;       // mid configurations in the class = C(M, js - jt)
    uint32_t n1;         // of which the first mid bit is set = C(M-1, js - jt - 1)
    uint32_t item_off;   // first SdBlkItem of the class
    uint32_t pad_;
};
struct SdBlkJs {
    uint32_t size;       // C(B, js) real elements
    uint32_t size_pad;   // stored elements (multiple of 16)
    uint32_t nunits[2];  // work items per tile, [0]: f64 (one chunk per class), [1]: c128 (classes of 10 in two chunks)
    SdBlkCls cls[SD_BLK_NCLS];
};
// one mid configuration of one class
struct alignas(16) SdBlkItem {
    uint8_t nb[12];      // nb[pm]: class-local index of c with mid bits pm, pm+1 swapped; 0xFF = parallel
    uint16_t c;          // mid configuration bits
    uint16_t u2x;        // class-local index (in class jt +- 1) of c with its last bit flipped
};

struct SdBlkShards {
    int world, rank;
    uint64_t pstart[SD_MAX_WORLD + 1];   // stored-element offset of each shard (tile aligned)
};

struct SdBlkParams {
    int L, k, A;
    int nbuf;                        // tile buffers in shared memory
    int pfp;                         // producer L2 prefetch of partner tiles: bit 0 on, bit 1 one tile late, bit 3 two tiles late, bit 2 all prefix entries (else the far ones), bit 4 + the crossing partner
    uint64_t key_lo, key_hi;         // tile keys of this launch (this shard)
    double Jhop[SD_MAX_L + 1];       // hop coefficient of bond p (positions p, p+1)
    double Jz[SD_MAX_L + 1];
    double h[SD_MAX_L + 1];
    double dtail[1 << SD_BLK_T];     // diag of the tail sites + tail-internal zz, by tail bits
    double Jmid[SD_BLK_M];           // = Jhop[A + pm]: mid bonds pm = 0 .. M-2, [M-1] the mid|tail bond (constant-bank operands)
    double Jtail[SD_BLK_T];          // = Jhop[A + M + q]: tail bonds
    double qx;                       // Jz of the mid|tail bond * 0.25
    const uint64_t *W;               // [A*(A+1)] stored elements of all tiles "1 at q, `below` ones before"
    const SdBlkJs *js;               // [B+1]
    const uint16_t *units;           // [2][(B+1)*MAXUNITS]: jt << 8 | unit-in-class, heavy classes first
    const SdBlkItem *items;
    const double *dmid;              // [1 << M] diag of the mid sites + mid-internal zz, in ITEM order (same index as items[])
    uint32_t cap;                    // largest size_pad
    int wrap_on;                     // periodic chain (SpinModel.jl:71-78): the bond between sites L-1 and 0 is present
    double wrapJ, wrapJz4;           // its hop coefficient and Jz / 4
    const uint32_t *order;           // optional tile order of this shard (keys, norder of them); nullptr: rank order
    uint32_t norder;
    SdBlkShards shards;
};

SD_HD int sd_blk_owner(const SdBlkShards &m, uint64_t p) {
    int g = 0;
#pragma unroll
    for (int i = 1; i < SD_MAX_WORLD; ++i) g += (i < m.world && p >= m.pstart[i]) ? 1 : 0;
    return g;
}

// prefix bits of tile `key`: keys enumerate prefixes in rank order ("1 first").
SD_HD uint64_t sd_blk_prefix_bits(uint64_t key, int A) {
    uint64_t Pb = 0;
    for (int q = 0; q < A; ++q)
        if (!((key >> (A - 1 - q)) & 1ULL)) Pb |= 1ULL << q;
    return Pb;
}

#if !defined(__CUDACC__)
// host build (tests/emul/emul_blk.cpp runs the item body on the CPU): the CUDA vector types it uses
#include <cstring>
struct alignas(16) double2 { double x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }
#endif


// ------------------------------------------------------------------ stored position <-> (class, tail configuration, mid configuration)
// Stored position p of a tile with class table I, for a vector of nc components -> (jt, e, u); false for padding.
SD_HD bool sd_blk_decode(const SdBlkJs &I, int nc, uint32_t p, int &jt, uint32_t &e, uint32_t &u) {
    jt = -1;
    uint32_t rel = 0;
#pragma unroll
    for (int j = 0; j < SD_BLK_NCLS; ++j) {
        const uint32_t len = I.cls[j].pitch * (uint32_t)sd_cbinom(SD_BLK_T, j);
        if (jt < 0 && p >= I.cls[j].cb && p < I.cls[j].cb + len) { jt = j; rel = p - I.cls[j].cb; }
    }
    if (jt < 0) return false;
    const uint32_t pitch = I.cls[jt].pitch;
    const uint32_t nt = (uint32_t)sd_cbinom(SD_BLK_T, jt);
    if (nc == 1 && (nt & 1u) && rel >= (nt - 1u) * pitch) {        // f64, odd class: last row is plain
        e = nt - 1u; u = rel - (nt - 1u) * pitch;
    } else if (nc == 1) {                                          // f64: pair rows (2s, 2s+1) of double2 per block
        const uint32_t pr = rel / (2u * pitch), r2 = rel % (2u * pitch);
        u = r2 >> 1; e = 2u * pr + (r2 & 1u);
    } else {                                                       // c128: one row per tail configuration
        e = rel / pitch; u = rel % pitch;
    }
    return u < I.cls[jt].nblk && e < nt;
}
// position of (jt, e, u) inside the tile for a vector of nc components (inverse of sd_blk_decode)
SD_HD uint32_t sd_blk_encode(const SdBlkJs &I, int nc, int jt, uint32_t e, uint32_t u) {
    const SdBlkCls &c = I.cls[jt];
    const uint32_t nt = (uint32_t)sd_cbinom(SD_BLK_T, jt);
    if (nc == 1 && (nt & 1u) && e == nt - 1u) return c.cb + e * c.pitch + u;
    if (nc == 1) return c.cb + (e >> 1) * 2u * c.pitch + 2u * u + (e & 1u);
    return c.cb + e * c.pitch + u;
}


// ------------------------------------------------------------------ periodic wrap bond (sites L-1 and 0; SpinModel.jl:71-78)
// The bond between the last tail site and the first prefix site is the one bond of a periodic chain that is neither
// inside a tile nor between two tiles of equal shape: flipping prefix bit 0 changes the prefix popcount, so the partner
// tile has suffix popcount js +- 1, and flipping tail bit T-1 moves the element to tail class jt +- 1 -- the mid
// configuration keeps its popcount, hence its index u in the class (item lists are shared by mid popcount).  The tile
// header carries the partner tile (wptr, jsw) and prefix bit 0; the item body of the WRAP kernel variant (sd_blkl.h) adds,
// per tail configuration whose last bit differs from prefix bit 0, one element of the partner tile, and +-Jz/4 on the
// diagonal.  The open chain's kernels are separate instantiations and carry none of it.

// ------------------------------------------------------------------ tile header
struct alignas(16) SdBlkEnt {
    const double *p;                     // stored base of the neighbour tile (component 0)
    double J;                            // hop coefficient of the bond
};
struct alignas(16) SdBlkHdr {
    uint64_t base;                       // stored-element offset of the tile (global, all shards)
    int js, jsx;                         // suffix popcount of the tile / of the crossing partner tile
    int valid;                           // 1: tile, -1: end of this CTA's tile list
    int nnb, nfar;                       // active prefix-internal bonds; the first nfar are beyond L2 reach
    int ntot;                            // nnb + 1 if the prefix|mid crossing bond is active: entry nnb of nb[]
    int bP;                              // last prefix bit
    unsigned next_unit;                  // work counter of the consumer warps
    unsigned tile_index;                 // key - key_lo (slot of the per-tile partial sums)
    double dP[2];                        // prefix diag + prefix|mid zz, by first mid bit
    double Jx;                           // hop coefficient of the prefix|mid bond (0: none)
    const double *xptr;                  // stored base of the crossing partner tile (component 0)
    const double *wptr;                  // periodic chain: stored base of the wrap partner tile (prefix bit 0 flipped), nullptr: none / no hop
    int jsw, b0;                         // its suffix popcount; prefix bit 0 of this tile
    int wloc;                            // the wrap partner tile lives in this rank's shard
    SdBlkEnt nb[SD_BLK_MAXA + 1];        // neighbour tiles of the active prefix bonds (one LDS.128 per entry); entry nnb: the crossing bond (if active)
    unsigned char nbloc[SD_BLK_MAXA + 3];   // entry lives in this rank's shard (the producer's L2 prefetch skips peer memory); [nnb]: the crossing partner
    double usum[SD_NSLOT][SD_BLK_MAXUNITS];   // per-unit reduction results (deterministic: summed in unit order)
};

// Header of tile `key`, one LANE per prefix position q (the kernel runs it warp-wide and sums with
// shuffles, tests/emul loops over the lanes).  W[q*(A+1) + below] lives in shared memory.
SD_HD uint64_t sd_blk_key_prefix(uint64_t key, int A) {
#if defined(__CUDA_ARCH__)
    return __brevll(~key) >> (64 - A);                            // A >= 1
#else
    return sd_blk_prefix_bits(key, A);
#endif
}
struct SdBlkHdrLane {
    uint64_t term, wq, wn;       // rank-base contribution; W of this site / of the bond partner
    uint64_t wterm;              // periodic chain: rank-base contribution of this site in the wrap partner tile (prefix bit 0 flipped)
    double d;                    // diagonal contribution of site q and bond (q, q+1)
    int bit;
    bool act;                    // bond (q, q+1) is an active prefix-internal hop
};
template <bool WRAP = false>
SD_HD SdBlkHdrLane sd_blk_hdr_lane(const SdBlkParams &P, const uint64_t *W, uint64_t Pb, int q) {
    const int A = P.A;
    SdBlkHdrLane l;
    const int bit = (int)((Pb >> q) & 1ULL), bn = (int)((Pb >> (q + 1)) & 1ULL);
    const int below = SD_POPC64(Pb & ((1ULL << q) - 1ULL));
    l.term = 0; l.wq = 0; l.wn = 0; l.d = 0.0; l.act = false; l.bit = bit; l.wterm = 0;
    if (q < A) {
        if constexpr (WRAP) {
            const uint64_t Pw = Pb ^ 1ULL;
            if (!((Pw >> q) & 1ULL)) l.wterm = W[q * (A + 1) + SD_POPC64(Pw & ((1ULL << q) - 1ULL))];
        }
        l.wq = W[q * (A + 1) + below];
        if (!bit) l.term = l.wq;
        const double sq = bit ? 0.5 : -0.5;
        l.d = P.h[q] * sq;
        if (q + 1 < A) {
            l.d += P.Jz[q] * sq * (bn ? 0.5 : -0.5);
            l.act = (bit != bn) && (P.Jhop[q] != 0.0);
            if (l.act) l.wn = W[(q + 1) * (A + 1) + below + 1];
        }
    }
    return l;
}
// second half, after the warp-wide sums base = sum(term), dpre = sum(d), actmask = ballot(act).
// Entry order: partner tiles beyond L2 reach (q < qfar; on other GPUs when sharded) before the near ones.
// (Round 2 tried the remote tiles last with an L1 prefetch at the start of the item: peer-memory prefetches are
// pathologically slow -- 237 ms per L = 32 apply on 2 GPUs instead of 3.4 -- so remote tiles are plain .cg loads; an
// L2 prefetch of the later LOCAL entries at the start of the item cost 10 % too: profiles/round2_h_ab.txt.)
template <int NC, class HDR = SdBlkHdr, bool WRAP = false>
SD_HD void sd_blk_hdr_fill(const SdBlkParams &P, const SdBlkHdrLane &l, uint64_t Pb, uint64_t key, uint64_t base, double dpre,
                           unsigned actmask, int qfar, int q, HDR &H, const SdVecView &psi, uint64_t wbase = 0) {
    const int A = P.A, js = P.k - SD_POPC64(Pb);
    const int bit = l.bit;
    const unsigned farmask = (qfar >= 32) ? 0xffffffffu : ((1u << qfar) - 1u);
    const int nfar = SD_POPC32(actmask & farmask);
    if (l.act) {
        // (1,0) -> (0,1): + (W[q][below] - W[q+1][below+1]);  (0,1) -> (1,0): the negative
        const uint64_t dl = l.wq - l.wn;
        const uint64_t nbase = bit ? base + dl : base - dl;
        const unsigned lt = (1u << q) - 1u;
        const int slot = ((farmask >> q) & 1u) ? SD_POPC32(actmask & farmask & lt) : nfar + SD_POPC32(actmask & ~farmask & lt);
        const int owner = sd_blk_owner(P.shards, nbase);
        H.nb[slot].p = psi.base[owner] + (size_t)NC * nbase;
        H.nb[slot].J = P.Jhop[q];
        H.nbloc[slot] = owner == P.shards.rank ? 1 : 0;
    }
    if (q == A - 1) {                                             // prefix|mid crossing bond
        const int jsx = bit ? js + 1 : js - 1;
        const double J = P.Jhop[q];
        const bool ok = (J != 0.0) && jsx >= 0 && jsx <= SD_BLK_B;
        const uint64_t nbase = bit ? base + l.wq : base - l.wq;
        H.jsx = jsx;
        H.Jx = ok ? J : 0.0;
        const int xowner = ok ? sd_blk_owner(P.shards, nbase) : P.shards.rank;
        H.xptr = ok ? psi.base[xowner] + (size_t)NC * nbase : nullptr;
        if (ok) {                                                 // the crossing bond as entry nnb of the stream list
            const int nnb = SD_POPC32(actmask);
            H.nb[nnb].p = H.xptr;
            H.nb[nnb].J = J;
            H.nbloc[nnb] = xowner == P.shards.rank ? 1 : 0;
        }
        H.bP = bit;
        const double sl = bit ? 0.5 : -0.5;
        H.dP[0] = dpre + P.Jz[q] * sl * (-0.5);
        H.dP[1] = dpre + P.Jz[q] * sl * (0.5);
    }
    if (q == 0) {
        if constexpr (WRAP) {
            H.wptr = nullptr; H.jsw = js; H.b0 = (int)(Pb & 1ULL); H.wloc = 0;
            const int jsw = P.k - SD_POPC64(Pb ^ 1ULL);
            if (P.wrapJ != 0.0 && jsw >= 0 && jsw <= SD_BLK_B) {
                const int wowner = sd_blk_owner(P.shards, wbase);
                H.wptr = psi.base[wowner] + (size_t)NC * wbase;
                H.jsw = jsw;
                H.wloc = wowner == P.shards.rank ? 1 : 0;
            }
        }
        H.base = base;
        H.js = js;
        H.valid = 1;
        H.nnb = SD_POPC32(actmask);
        H.nfar = nfar;
        {
            const int bl = (int)((Pb >> (A - 1)) & 1ULL), jsx = bl ? js + 1 : js - 1;
            const bool okx = (P.Jhop[A - 1] != 0.0) && jsx >= 0 && jsx <= SD_BLK_B;
            const int ntot = H.nnb + (okx ? 1 : 0);
            H.ntot = ntot;
        }
        H.next_unit = 0;
        H.tile_index = (unsigned)(key - P.key_lo);
    }
}
#if defined(__CUDACC__)
template <int NC, class HDR = SdBlkHdr, bool WRAP = false>
__device__ __forceinline__ void sd_blk_make_hdr(const SdBlkParams &P, const uint64_t *W, uint64_t key, HDR &H,
                                                const SdVecView &psi, int qfar, unsigned lane) {
    const uint64_t Pb = sd_blk_key_prefix(key, P.A);
    SdBlkHdrLane l = sd_blk_hdr_lane<WRAP>(P, W, Pb, (int)lane);
    uint64_t base = l.term;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
    double dpre = l.d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dpre += __shfl_xor_sync(0xffffffffu, dpre, o);
    const unsigned actmask = __ballot_sync(0xffffffffu, l.act);
    uint64_t wbase = l.wterm;
    if constexpr (WRAP) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wbase += __shfl_xor_sync(0xffffffffu, wbase, o);
    }
    sd_blk_hdr_fill<NC, HDR, WRAP>(P, l, Pb, key, base, dpre, actmask, qfar, (int)lane, H, psi, wbase);
}
#endif
// the same header on the host (tests/emul, sd_halo_host.h): the 32 "lanes" in a loop
template <int NC, bool WRAP = false>
inline void sd_blk_hdr_host(const SdBlkParams &P, const uint64_t *W, uint64_t key, int qfar, const SdVecView &psi, SdBlkHdr &H) {
    const uint64_t Pb = sd_blk_prefix_bits(key, P.A);
    SdBlkHdrLane lanes[32];
    uint64_t base = 0, wbase = 0;
    double dpre = 0.0;
    unsigned actmask = 0;
    for (int q = 0; q < 32; ++q) {
        lanes[q] = sd_blk_hdr_lane<WRAP>(P, W, Pb, q);
        base += lanes[q].term;
        wbase += lanes[q].wterm;
        dpre += lanes[q].d;
        if (lanes[q].act) actmask |= 1u << q;
    }
    for (int q = 0; q < 32; ++q) sd_blk_hdr_fill<NC, SdBlkHdr, WRAP>(P, lanes[q], Pb, key, base, dpre, actmask, qfar, q, H, psi, wbase);
}

// ------------------------------------------------------------------ per-item body
// A work ITEM is (unit of 32 mid configurations, slot chunk): a lane owns ONE mid configuration u of
// class jt.  All data moves as 16-byte SLOTS:
//   f64 : slot s = tail configurations (2s, 2s+1) of the block  (pair rows, see the layout above)
//   c128: slot s = tail configuration s, (re, im)
// H is real, so a slot is two independent real columns for everything but the tail-internal hops and
// the mid|tail crossing bond, which address single tail configurations.
// The helpers below are shared by the item body of sd_blkl.h.
// Everything compiles for host and device: tests/emul/emul_blk.cpp runs it on the CPU, lane by lane, against
// the oracle (index algebra, layout, epilogues, shard ownership without a GPU).
SD_HD double2 sd_blk_ldg(const double *p) {
    double2 v;
#if defined(__CUDA_ARCH__)
    // .cg: cached in L2 only (normal eviction priority); L1::no_allocate loads were measured to be treated
    // as streaming by L2 as well (+6 GB of DRAM reads per apply at L = 32)
    asm("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
#else
    v.x = p[0]; v.y = p[1];
#endif
    return v;
}
SD_HD double2 sd_blk_ldg_half(const double *p) {
    double2 v;
#if defined(__CUDA_ARCH__)
    asm("ld.global.cg.f64 %0, [%1];" : "=d"(v.x) : "l"(p));
#else
    v.x = p[0];
#endif
    v.y = 0.0;
    return v;
}
SD_HD void sd_blk_stg(double *p, double2 v) {
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
#else
    p[0] = v.x; p[1] = v.y;
#endif
}
SD_HD void sd_blk_stg_half(double *p, double v) {
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
#else
    p[0] = v;
#endif
}
SD_HD uint4 sd_blk_ld_item(const SdBlkItem *p) {
#if defined(__CUDA_ARCH__)
    return __ldg((const uint4 *)p);
#else
    uint4 v;
    memcpy(&v, p, sizeof(v));
    return v;
#endif
}

// tail-configuration accessors (e is a compile-time element index relative to the first element of
// the chunk for acc, absolute for own)
#define SD_BLK_EL(arr, e, NC_) (*((NC_) == 1 ? (((e) & 1) ? &arr[(e) >> 1].y : &arr[(e) >> 1].x) : &arr[(e)].x))

template <int NC, int JT, int E0, int EC, int NO, int t, int q>
struct SdBlkTailHop {
    static SD_HD void run(double2 (&acc)[EC], const double2 (&own)[NO], const double *Jt) {
        constexpr unsigned cfg = sd_tail_cfg(SD_BLK_T, JT, E0 + t);
        constexpr bool act = (((cfg >> q) ^ (cfg >> (q + 1))) & 1u) != 0;
        if constexpr (act) {
            constexpr int t2 = sd_tail_rank(SD_BLK_T, JT, cfg ^ (3u << q));
            const double J = Jt[q];
            if (NC == 1) {
                SD_BLK_EL(acc, t, 1) += J * ((t2 & 1) ? own[t2 >> 1].y : own[t2 >> 1].x);
            } else {
                acc[t].x += J * own[t2].x;
                acc[t].y += J * own[t2].y;
            }
        }
        if constexpr (q + 2 < SD_BLK_T) SdBlkTailHop<NC, JT, E0, EC, NO, t, q + 1>::run(acc, own, Jt);
    }
};
// NE = tail configurations in the chunk
template <int NC, int JT, int E0, int NE, int EC, int NO, int t>
struct SdBlkTailRow {
    static SD_HD void run(double2 (&acc)[EC], const double2 (&own)[NO], const double *Jt,
                                               const double *dtail, double d0, double dx0) {
        constexpr unsigned cfg = sd_tail_cfg(SD_BLK_T, JT, E0 + t);
        // + dx when tail bit 0 equals the last mid bit (dx already carries the sign of the last mid bit)
        const double d = d0 + dtail[cfg] + ((cfg & 1u) ? dx0 : -dx0);
        if (NC == 1) {
            SD_BLK_EL(acc, t, 1) += d * (((E0 + t) & 1) ? own[(E0 + t) >> 1].y : own[(E0 + t) >> 1].x);
        } else {
            acc[t].x += d * own[E0 + t].x;
            acc[t].y += d * own[E0 + t].y;
        }
        SdBlkTailHop<NC, JT, E0, EC, NO, t, 0>::run(acc, own, Jt);
        if constexpr (t + 1 < NE) SdBlkTailRow<NC, JT, E0, NE, EC, NO, t + 1>::run(acc, own, Jt, dtail, d0, dx0);
    }
};
// own block: diagonal + tail-internal hops.  The only part specialised on (class, chunk): the tail
// configurations are compile-time constants, so tail hops are register moves.
// E0/NE: first tail configuration / number of tail configurations of the chunk (f64: the whole class).
template <int NC, int JT, int E0, int NE, int EC>
SD_HD void sd_blk_tail(double2 (&acc)[EC], const double *own_ptr, uint32_t ss, uint32_t u,
                                            const double *Jt, const double *dtail, double d0, double dx0) {
    constexpr int NT = sd_cbinom(SD_BLK_T, JT);
    constexpr int NO = NC == 1 ? (NT + 1) / 2 : NT;              // slots of the whole block
    static_assert(NC == 2 || (E0 == 0 && NE == NT && EC == NO), "f64 items cover the whole class");
    static_assert(NC == 1 || NE == EC, "c128 slots are tail configurations");
    double2 own[NO];
#pragma unroll
    for (int s = 0; s < NO; ++s) {
        if (NC == 1 && (NT & 1) && s == NO - 1) own[s] = make_double2(*(own_ptr + s * ss - u), 0.0);   // half slot
        else own[s] = *(const double2 *)(own_ptr + s * ss);
    }
    SdBlkTailRow<NC, JT, E0, NE, EC, NO, 0>::run(acc, own, Jt, dtail, d0, dx0);
}

// the producer warp's view of the CTA's shared memory
struct SdBlkSmem {
    uint64_t *full, *empty;      // [nbuf] mbarriers
    SdBlkHdr *hdr;               // [nbuf]
    const uint64_t *W;           // [A*(A+1)] (global memory, L2 resident)
    const SdBlkJs *js;           // [B+1]
    double *tiles;               // [nbuf][cap*NC]
};

#if defined(__CUDACC__)
// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ unsigned sd_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sd_mbar_init(uint64_t *b, unsigned cnt) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sd_smem_u32(b)), "r"(cnt) : "memory");
}
__device__ __forceinline__ void sd_mbar_expect_tx(uint64_t *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sd_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sd_mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sd_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void sd_mbar_wait(uint64_t *b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SD_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SD_DONE_%=;\n"
        "bra SD_WAIT_%=;\n"
        "SD_DONE_%=:\n"
        "}" ::"r"(sd_smem_u32(b)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void sd_bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sd_smem_u32(dst)), "l"(src), "r"(bytes), "r"(sd_smem_u32(bar)) : "memory");
}
// ------------------------------------------------------------------ producer warp and reduction tail (kernel: sd_blkl.h)
// Fused reductions, deterministic and without fences in the consumer warps: lane 0 of a consumer warp leaves the warp sum
// of each item in the tile header (usum[slot][item]) and releases the buffer with its usual mbarrier arrive (release
// semantics); the producer warp, which acquires the buffer through the same mbarrier before it reuses it, adds the item
// sums in item order and writes the tile's partial sums.  (Round 2: a __threadfence_block + atomic counter per item in
// the consumers cost 22 % of a Lanczos step -- the fence waits for the item's global stores.)
template <int NC>
__device__ __forceinline__ void sd_blk_flush_sums(const SdBlkSmem &S, SdBlkHdr &H, const SdEpi &epi, int slotmask, unsigned lane) {
    if (H.valid == 1 && lane < SD_NSLOT && ((slotmask >> lane) & 1)) {
        const unsigned nunits = S.js[H.js].nunits[NC - 1];
        double t = 0.0;
        for (unsigned j = 0; j < nunits; ++j) t += ((volatile double *)H.usum[lane])[j];
        epi.partials[(size_t)lane * epi.nparts + H.tile_index] = t;
    }
    __syncwarp();
}
// producer warp: tile keys from the global counter, tile headers, TMA of the own tiles
template <int NC, bool WRAP = false>
__device__ __forceinline__ void sd_blk_producer(const SdBlkParams &P, const SdBlkSmem &S, const SdVecView &psi, int qfar,
                                                unsigned long long *tile_ctr, unsigned lane, const SdEpi &epi, int slotmask) {
    const int nbuf = P.nbuf;
    const size_t tile_doubles = (size_t)P.cap * NC;
    for (unsigned i = 0;; ++i) {
        const int b = (int)(i % (unsigned)nbuf);
        const unsigned round = i / (unsigned)nbuf;
        uint64_t key;
        for (;;) {                                             // next valid tile of this shard
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(tile_ctr, 1ULL);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (P.order != nullptr) {                          // explicit tile order (valid tiles only, sd_blk_tile_order)
                key = t < (unsigned long long)P.norder ? (uint64_t)P.order[t] : P.key_hi;
                break;
            }
            key = P.key_lo + t;
            if (key >= P.key_hi) break;
            const uint64_t Pb = __brevll(~key) >> (64 - P.A);
            const int js = P.k - __popcll(Pb);
            if (js >= 0 && js <= SD_BLK_B) break;              // else: impossible suffix popcount
        }
        sd_mbar_wait(&S.empty[b], (round & 1u) ^ 1u);          // consumers released this buffer
        SdBlkHdr &H = S.hdr[b];
        if (slotmask && i >= (unsigned)nbuf) sd_blk_flush_sums<NC>(S, H, epi, slotmask, lane);   // of the tile that used this buffer
        if (key >= P.key_hi) {
            if (lane == 0) { H.valid = -1; }
            __syncwarp();
            if (lane == 0) sd_mbar_arrive(&S.full[b]);
            // the tiles still in the other buffers: wait until their consumers are done, then their sums
            for (unsigned k = 1; slotmask && k < (unsigned)nbuf; ++k) {
                const unsigned ii = i + k;
                if (ii < (unsigned)nbuf) continue;                 // that buffer never held a tile
                const int bb = (int)(ii % (unsigned)nbuf);
                sd_mbar_wait(&S.empty[bb], ((ii / (unsigned)nbuf) & 1u) ^ 1u);
                sd_blk_flush_sums<NC>(S, S.hdr[bb], epi, slotmask, lane);
            }
            break;
        }
        sd_blk_make_hdr<NC, SdBlkHdr, WRAP>(P, S.W, key, H, psi, qfar, lane);
        __syncwarp();
        const uint32_t bytes = S.js[H.js].size_pad * (uint32_t)(NC * 8);
        const char *src = (const char *)(psi.base[P.shards.rank] + (size_t)NC * H.base);
        char *dst = (char *)(S.tiles + (size_t)b * tile_doubles);
        if (lane == 0) sd_mbar_expect_tx(&S.full[b], bytes);
        __syncwarp();
        constexpr uint32_t CH = 8192;
        for (uint32_t o = lane * CH; o < bytes; o += 32 * CH)
            sd_bulk_g2s(dst + o, src + o, (bytes - o < CH) ? bytes - o : CH, &S.full[b]);
        // L2 bulk prefetch of the partner tiles by the producer (no consumer instruction).  Measured at L = 32
        // (profiles/round2_n_ab.txt): issued for the tile just handed to the consumers' queue it thrashes (all entries:
        // 20.9 GB of DRAM reads, 5.91 ms); one tile later -- for the tile the consumers reach next -- it lifts the L2 hit
        // rate from 53 to 62 % at 15.3 GB and takes 5.77 -> 5.40 ms.  Peer memory is never prefetched (pathologically slow): sharded, only the partner tiles in
        // this rank's own shard are (nbloc).
        if (P.pfp & 1) {                                       // partner tiles of the prefix bonds have the tile's own js, i.e. its size
            unsigned late = (P.pfp & 8) ? 2u : ((P.pfp & 2) ? 1u : 0u);
            if (late > (unsigned)nbuf - 1u) late = (unsigned)nbuf - 1u;   // the header of that tile must still be in its buffer
            if (i >= late) {
                const SdBlkHdr &Hp = S.hdr[((unsigned)b + (unsigned)nbuf - late) % (unsigned)nbuf];
                const int cnt = (P.pfp & 4) ? Hp.nnb : Hp.nfar;
                if ((int)lane < cnt) {
                    const uint32_t pb = Hp.nbloc[lane] ? S.js[Hp.js].size_pad * (uint32_t)(NC * 8) : 0u;   // local tiles only
                    if (pb)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(Hp.nb[lane].p), "r"(pb) : "memory");
                } else if (WRAP && (int)lane == cnt + 1 && Hp.wptr != nullptr && Hp.wloc) {     // periodic chain: the wrap partner tile (its own js)
                    const uint32_t pb = S.js[Hp.jsw].size_pad * (uint32_t)(NC * 8);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(Hp.wptr), "r"(pb) : "memory");
                } else if ((P.pfp & 16) && (int)lane == cnt && Hp.xptr != nullptr && Hp.nbloc[Hp.nnb]) {   // the prefix|mid crossing partner (its own js)
                    const uint32_t pb = S.js[Hp.jsx].size_pad * (uint32_t)(NC * 8);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(Hp.xptr), "r"(pb) : "memory");
                }
            }
        }
    }
}
// per-item tail of a consumer warp: warp sums of the fused reductions into the tile header (see sd_blk_flush_sums)
__device__ __forceinline__ void sd_blk_item_reduce(SdBlkHdr &H, int slotmask, unsigned un, const double (&red)[SD_NSLOT], unsigned lane) {
#pragma unroll
    for (int s = 0; s < SD_NSLOT; ++s) {
        if (!((slotmask >> s) & 1)) continue;
        double w = red[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
        if (lane == 0) H.usum[s][un] = w;
    }
}

// ------------------------------------------------------------------ layout conversion
// dir 0: blk[pos] = rankvec[rank]   (scatter a rank-ordered vector into block layout, padding := 0)
// dir 1: rankvec[rank] = blk[pos]
// One CTA per tile (grid-stride).  rank_local points at the local shard of the rank-ordered vector
// (element r - rstart), blk_local at the local shard of the block-layout vector.
// seeded != 0: dir 0 writes scale*sd_seeded_value(seed (+c), rank) instead of reading rankvec.
struct SdBlkPermute {
    int dir, nc_blk, nc_rank, seeded;
    uint64_t seed;
    double scale;
    uint64_t rstart;             // first rank of the local shard
    const uint64_t *binom;       // [65*65]
};
#if !defined(SD_NO_KERNELS)   // a second translation unit (sd_batch.cu) includes this header for the types only
__global__ void __launch_bounds__(256) sd_blk_permute_kernel(const __grid_constant__ SdBlkParams P, SdBlkPermute Q,
                                                             double *blk_local, double *rank_local) {
    __shared__ uint64_t s_base[2];
    const int A = P.A, k = P.k, L = P.L;
    for (uint64_t key = P.key_lo + blockIdx.x; key < P.key_hi; key += gridDim.x) {
        const uint64_t Pb = __brevll(~key) >> (64 - A);
        const int js = k - __popcll(Pb);
        if (js < 0 || js > SD_BLK_B) continue;                     // uniform over the CTA
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t pb = 0, rb = 0;
            for (int q = 0; q < A; ++q) {
                if ((Pb >> q) & 1ULL) continue;
                const int below = __popcll(Pb & ((1ULL << q) - 1ULL));
                pb += P.W[q * (A + 1) + below];
                rb += sd_binom_at(Q.binom, SD_BINOM_DIM, L - 1 - q, k - below - 1);
            }
            s_base[0] = pb; s_base[1] = rb;
        }
        __syncthreads();
        const uint64_t pbase = s_base[0] - P.shards.pstart[P.shards.rank], rbase = s_base[1] - Q.rstart;
        const SdBlkJs &I = P.js[js];
        for (uint32_t p = threadIdx.x; p < I.size_pad; p += blockDim.x) {
            int jt = -1;
            uint32_t rel = 0;
#pragma unroll
            for (int j = 0; j < SD_BLK_NCLS; ++j) {
                const uint32_t len = I.cls[j].pitch * (uint32_t)sd_cbinom(SD_BLK_T, j);
                if (jt < 0 && p >= I.cls[j].cb && p < I.cls[j].cb + len) { jt = j; rel = p - I.cls[j].cb; }
            }
            bool real = jt >= 0;
            uint32_t e = 0, u = 0;
            if (real) {
                const uint32_t pitch = I.cls[jt].pitch;
                const uint32_t nt = (uint32_t)sd_cbinom(SD_BLK_T, jt);
                if (Q.nc_blk == 1 && (nt & 1u) && rel >= (nt - 1u) * pitch) {   // f64, odd class: last row is plain
                    e = nt - 1u; u = rel - (nt - 1u) * pitch;
                } else if (Q.nc_blk == 1) {                        // f64: pair rows (2s, 2s+1) of double2 per block
                    const uint32_t pr = rel / (2u * pitch), r2 = rel % (2u * pitch);
                    u = r2 >> 1; e = 2u * pr + (r2 & 1u);
                } else {                                           // c128: one row per tail configuration
                    e = rel / pitch; u = rel % pitch;
                }
                real = u < I.cls[jt].nblk && e < (uint32_t)sd_cbinom(SD_BLK_T, jt);
            }
            double *bp = blk_local + (size_t)(pbase + p) * Q.nc_blk;
            if (!real) {
                if (Q.dir == 0) for (int c = 0; c < Q.nc_blk; ++c) bp[c] = 0.0;
                continue;
            }
            const unsigned cm = P.items[I.cls[jt].item_off + u].c;
            const unsigned tau = sd_tail_cfg(SD_BLK_T, jt, (int)e);
            const uint64_t suf = (uint64_t)cm | ((uint64_t)tau << SD_BLK_M);
            const uint64_t lr = sd_rank_state(suf, SD_BLK_B, js, Q.binom, SD_BINOM_DIM);
            double *rp = rank_local + (size_t)(rbase + lr) * Q.nc_rank;
            if (Q.dir == 0) {
                if (Q.seeded) {
                    for (int c = 0; c < Q.nc_blk; ++c) bp[c] = Q.scale * sd_seeded_value(Q.seed + (uint64_t)c, Q.rstart + rbase + lr);
                } else {
                    for (int c = 0; c < Q.nc_blk; ++c) bp[c] = c < Q.nc_rank ? rp[c] : 0.0;
                }
            } else {
                for (int c = 0; c < Q.nc_rank; ++c) rp[c] = c < Q.nc_blk ? bp[c] : 0.0;
            }
        }
    }
}
#endif  // SD_NO_KERNELS
#endif  // __CUDACC__
