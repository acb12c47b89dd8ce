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
//   0a every thread: one prefix position / bond each (one binomial lookup, all
//      in flight together); item records of phase 2 are prefetched to registers
//   0b one thread: ordered sums, neighbour-list compaction
//   1  flat, coalesced, U elements per thread in flight per neighbour tile:
//      own psi -> smem, g = sum_J psi[neighbour tiles] -> smem
//   2  block-mapped: diag + g + tail/mid/crossing hops -> result in smem
//   3  flat, coalesced: epilogue (rescale / Chebyshev / dots) -> out
//
// Everything here is __host__ __device__ so tests/emul can run the identical
// body on the CPU (test infrastructure; the product only runs it on the GPU).
#pragma once
#include "sd_common.h"

#define SD_TILE_MAXNB 32      // prefix sites A <= 32
#define SD_TILE_MAXT 6
#define SD_TILE_MAXB 19     // suffix sites (tile = up to C(19,9) states would not fit smem anyway)
#define SD_TILE_U(NC) ((NC) == 1 ? 7 : 3)   // elements per thread per phase-1 warp chunk (x3 tiles in flight)

#if defined(__CUDA_ARCH__)
#define SD_ATOMIC_OR(p, v) atomicOr(p, v)
#define SD_LD_NEAR(p) __ldg(p)
#define SD_LD_FAR(p) __ldcs(p)
#define SD_ST_STREAM(p, v) __stcs(p, v)
// nc doubles (8 or 16 bytes) global -> shared without a register landing zone
#define SD_CP_ASYNC(dst, src, nc)                                                                   \
    do {                                                                                            \
        const unsigned sd_sa_ = (unsigned)__cvta_generic_to_shared(dst);                            \
        if ((nc) == 1) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sd_sa_), "l"(src) : "memory");  \
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sd_sa_), "l"(src) : "memory");           \
    } while (0)
#define SD_CP_ASYNC_WAIT_ALL() asm volatile("cp.async.wait_all;" ::: "memory")
// bulk L2 prefetch of [p, p+bytes): address aligned down to 16 B, size rounded up (vectors carry slack)
#define SD_PREFETCH_L2(p, bytes)                                                                    \
    do {                                                                                            \
        const unsigned long long sd_a_ = (unsigned long long)(p);                                   \
        const unsigned long long sd_b_ = sd_a_ & ~15ULL;                                            \
        const unsigned sd_n_ = (unsigned)(((sd_a_ - sd_b_) + (bytes) + 15ULL) & ~15ULL);            \
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(sd_b_), "r"(sd_n_) : "memory"); \
    } while (0)
#else
#define SD_ATOMIC_OR(p, v) (*(p) |= (v))
#define SD_LD_NEAR(p) (*(p))
#define SD_LD_FAR(p) (*(p))
#define SD_ST_STREAM(p, v) (*(p) = (v))
#define SD_CP_ASYNC(dst, src, nc)                          \
    do {                                                   \
        for (int sd_c_ = 0; sd_c_ < (nc); ++sd_c_) (dst)[sd_c_] = (src)[sd_c_]; \
    } while (0)
#define SD_CP_ASYNC_WAIT_ALL() ((void)0)
#define SD_PREFETCH_L2(p, bytes) ((void)(p))
#endif

// optional fine-grained cycle counters inside the phases (thread 0 only; -DSD_PHASE_TIMING)
#if defined(SD_PHASE_TIMING) && defined(__CUDA_ARCH__)
__device__ unsigned long long sd_phase_cycles[16];
#define SD_PTICK_INIT() long long ptick_ = clock64()
#define SD_PTICK(i)                                                             \
    do {                                                                        \
        if (threadIdx.x == 0) {                                                 \
            const long long now_ = clock64();                                   \
            atomicAdd(&sd_phase_cycles[i], (unsigned long long)(now_ - ptick_)); \
            ptick_ = now_;                                                      \
        }                                                                       \
    } while (0)
#elif defined(__CUDACC__)
__device__ unsigned long long sd_phase_cycles[16];
#define SD_PTICK_INIT() ((void)0)
#define SD_PTICK(i) ((void)0)
#else
#define SD_PTICK_INIT() ((void)0)
#define SD_PTICK(i) ((void)0)
#endif

// One phase-2 work item (a tail block), precomputed per (js, slot) on the host.
struct alignas(16) SdItem {
    uint16_t c;        // mid configuration bits; 0xFFFF marks a padding slot
    uint16_t u;        // class-local index of c
    uint16_t u2;       // class-local index of the crossing partner (c with its last bit flipped)
    uint16_t jt;       // tail popcount = class
    double dmid;       // diag of the mid sites + mid-internal zz
};

// Per suffix-popcount layout, precomputed on the host.
struct SdJsInfo {
    uint32_t cls_base[SD_TILE_MAXT + 2];   // smem element offset of class jt; [T+1] = padded tile size
    uint32_t nslots;                       // phase-2 slots (classes padded to warps)
    uint32_t item_off;                     // first SdItem of this js
    uint32_t perm_off;                     // first perm entry of this js
    uint32_t size;                         // C(B, js)
    uint32_t n1;                           // C(B-1, js-1): suffix configurations whose first bit is set
    uint32_t ncross;                       // C(B-1, js): rank shift of the (1,0)->(0,1) crossing hop
};

struct SdTileParams {
    int L, k, A, B, M, T;
    uint64_t key_lo;                 // tile key of blockIdx.x == 0
    uint64_t key_hi;                 // one past the last key of this launch
    int qfar;                        // prefix bonds q < qfar are far streams (shift beyond L2 reach): read evict-first
    uint32_t hop_mask;               // bit q: Jhop[q] != 0 (q < 32)
    double Jhop[SD_MAX_L];           // hop coefficient of bond p (positions p, p+1)
    double Jz[SD_MAX_L];             // zz coefficient of bond p
    double h[SD_MAX_L + 1];          // field at position p
    double dtail[1 << SD_TILE_MAXT]; // diag of the tail sites + tail-internal zz, by tail bits
    const uint64_t *binom;           // [65*65]
    const uint16_t *perm;            // [perm_off + l] -> smem element position
    const SdItem *items;             // [item_off + slot]
    SdJsInfo js[SD_TILE_MAXB + 1];   // per suffix popcount layout (kernel parameter space: no load latency)
    int pf_dist;                     // L2 prefetch distance in tiles (0 = off)
    const uint16_t *binomM;          // [(M+1)*(M+1)] C(n, r) for n, r <= M
    SdShardMap shards;
};

struct SdNbEntry {
    const double *ptr;               // virtual base: element l of the neighbour tile is ptr[l*NC]
    double J;
};

struct SdTileHdr {
    uint64_t base;                   // rank of the tile's first state
    uint32_t size;                   // C(B, js)
    int js, valid;
    int nfar, nfull;                 // nbf[0..nfar) far streams, nbf[nfar..nfull) near streams
    uint32_t mixed_mask;             // bit q: neighbour range straddles a shard boundary (slow path)
    uint32_t nslots, item_off, perm_off;
    double dP[2];                    // prefix diag + crossing zz, by first mid bit
    uint32_t cls_base[SD_TILE_MAXT + 2];
    SdNbEntry nbf[SD_TILE_MAXNB];    // compacted neighbour tiles of the active prefix-internal bonds
    SdNbEntry cross;                 // prefix|suffix crossing bond (ptr == null: inactive)
    // per prefix position q: rank shift and element range of bond q (crossing bond and slow path)
    int64_t t_off[SD_TILE_MAXNB];
    uint32_t t_lo[SD_TILE_MAXNB], t_hi[SD_TILE_MAXNB];
    double red[SD_NSLOT][16];        // CTA reduction scratch (<= 16 warps)
};

// Scratch of the host-callable two-step header (sd_tile_phase0a/0b, CPU emulation only; the device
// keeps these terms in registers of warp 0).
struct SdTileScratch {
    uint64_t t_base[SD_TILE_MAXNB];  // per prefix position: rank term
    double t_diag[SD_TILE_MAXNB];    //                      diagonal term
};

template <int NC>
struct SdTileView {
    SdTileHdr *hdr;
    double *spsi;                    // [cap*NC]
    double *sg;                      // [cap*NC]
    uint16_t *binomM;                // [(M+1)*(M+1)]
};

// smem: two headers (current tile / next tile of the persistent loop), psi tile, g/result tile, binomM
SD_HD size_t sd_tile_smem_bytes(int NC, uint32_t cap, int M) {
    size_t b = sizeof(SdTileHdr);
    b = 2 * ((b + 15) & ~(size_t)15);
    b += (size_t)2 * cap * NC * sizeof(double);
    b += (size_t)(M + 1) * (M + 1) * sizeof(uint16_t);
    return (b + 15) & ~(size_t)15;
}

template <int NC>
SD_HD SdTileView<NC> sd_tile_carve(void *smem, uint32_t cap) {
    SdTileView<NC> v;
    char *p = (char *)smem;
    v.hdr = (SdTileHdr *)p;                               // second header follows the first
    p += 2 * ((sizeof(SdTileHdr) + 15) & ~(size_t)15);
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
    if (A == 0) return 0;
#if defined(__CUDA_ARCH__)
    return __brevll(~key) >> (64 - A);
#else
    uint64_t Pb = 0;
    for (int q = 0; q < A; ++q)
        if (!((key >> (A - 1 - q)) & 1ULL)) Pb |= 1ULL << q;
    return Pb;
#endif
}

// ---------------------------------------------------------------- phase 0
// 0a: thread q < A computes the rank / diagonal terms of prefix position q and
//     the rank shift of bond q; binomM copy and the first phase-2 item prefetch
//     are spread over the CTA.  0b: thread q < A sums the A rank terms itself
//     (broadcast smem reads) and writes its own neighbour entry into a slot
//     computed from the bond-activity mask (popc prefix) -- no serial work.
//     Far bonds (q < qfar: shift larger than anything L2 can hold) come first.
SD_HD uint32_t sd_tile_active_mask(const SdTileParams &P, uint64_t Pb) {
    // bit q: prefix-internal bond q (positions q, q+1) is antiparallel and has J != 0
    if (P.A < 2) return 0u;
    const uint64_t m = (Pb ^ (Pb >> 1)) & ((1ULL << (P.A - 1)) - 1);
    return (uint32_t)m & P.hop_mask;
}

template <int NC>
SD_HD void sd_tile_phase0a(const SdTileParams &P, uint64_t key, const SdTileView<NC> &v,
                           unsigned tid, unsigned nthreads, SdItem &item0, SdTileScratch &X) {
    SdTileHdr &H = *v.hdr;
    const int L = P.L, k = P.k, A = P.A, B = P.B, M = P.M, T = P.T;
    const uint64_t *C = P.binom;
    const uint64_t Pb = sd_tile_prefix_bits(key, A);
    const int js = k - SD_POPC64(Pb);
    const bool valid = (js >= 0 && js <= B);
    if (tid == 0) {
        H.valid = valid ? 1 : 0;
        H.js = js;
        H.mixed_mask = 0;
    }
    item0.c = 0xFFFFu;
    if (!valid) return;
    const SdJsInfo &I = P.js[js];
    const uint32_t size = I.size;
    if (tid < I.nslots) item0 = P.items[I.item_off + tid];   // first phase-2 item: depends on js and tid only
    if (tid == 0) {
        H.size = size; H.nslots = I.nslots; H.item_off = I.item_off; H.perm_off = I.perm_off;
    }
    for (int q = (int)tid; q < A; q += (int)nthreads) {
        const int bit = (int)((Pb >> q) & 1ULL);
        const int below = SD_POPC64(Pb & ((1ULL << q) - 1));             // set bits at positions < q
        const double sq = bit ? 0.5 : -0.5;
        double d = P.h[q] * sq;
        X.t_base[q] = bit ? 0ULL : sd_binom_at(C, SD_BINOM_DIM, L - 1 - q, k - below - 1);
        int64_t off = 0;
        uint32_t lo = 0, hi = 0;
        if (q + 1 < A) {                                                  // prefix-internal bond q
            const int bn = (int)((Pb >> (q + 1)) & 1ULL);
            d += P.Jz[q] * sq * (bn ? 0.5 : -0.5);
            if (bit != bn && P.Jhop[q] != 0.0) {
                const uint64_t dl = sd_binom_at(C, SD_BINOM_DIM, L - 2 - q, k - (below + bit + bn));
                off = bit ? (int64_t)dl : -(int64_t)dl;
                hi = size;
            }
        } else if (P.Jhop[q] != 0.0) {                                    // prefix|suffix crossing bond
            const uint32_t n1 = I.n1;                                     // first suffix bit = 1
            if (bit) {                       // (1,0) -> (0,1): states with first suffix bit 0 move up
                if (n1 < size) { off = (int64_t)I.ncross; lo = n1; hi = size; }
            } else if (n1 > 0) {             // (0,1) -> (1,0)
                off = -(int64_t)n1; lo = 0; hi = n1;
            }
        }
        X.t_diag[q] = d;
        H.t_off[q] = off; H.t_lo[q] = lo; H.t_hi[q] = hi;
    }
    for (int i = (int)tid; i < (M + 1) * (M + 1); i += (int)nthreads) v.binomM[i] = P.binomM[i];
    for (int jt = (int)tid; jt <= T + 1; jt += (int)nthreads) H.cls_base[jt] = I.cls_base[jt];
}

template <int NC>
SD_HD void sd_tile_phase0b(const SdTileParams &P, uint64_t key, const SdTileView<NC> &v,
                           const SdVecView &psi, unsigned tid, unsigned nthreads, const SdTileScratch &X) {
    SdTileHdr &H = *v.hdr;
    const int A = P.A;
    if (!H.valid) return;
    if ((int)tid >= A && tid != 0) return;
    uint64_t base = 0;
    for (int q = 0; q < A; ++q) base += X.t_base[q];
    const uint64_t Pb = sd_tile_prefix_bits(key, A);
    const uint32_t act = sd_tile_active_mask(P, Pb);
    const uint32_t farbits = (P.qfar >= 32) ? ~0u : ((1u << P.qfar) - 1u);
    const uint32_t act_far = act & farbits, act_near = act & ~farbits;
    const int nfar = SD_POPC32(act_far);
    if (tid == 0) {
        double dpre = 0.0;
        for (int q = 0; q < A; ++q) dpre += X.t_diag[q];
        H.base = base;
        if (A > 0) {
            const double sl = ((Pb >> (A - 1)) & 1ULL) ? 0.5 : -0.5;
            H.dP[0] = dpre + P.Jz[A - 1] * sl * (-0.5);
            H.dP[1] = dpre + P.Jz[A - 1] * sl * (0.5);
        } else {
            H.dP[0] = H.dP[1] = dpre;
        }
        H.nfar = nfar;
        H.nfull = nfar + SD_POPC32(act_near);
        H.cross.ptr = nullptr;
    }
    const bool single = P.shards.world == 1;
    for (int q = (int)tid; q < A; q += (int)nthreads) {
        const uint32_t lo = H.t_lo[q], hi = H.t_hi[q];
        if (hi <= lo) continue;
        const int64_t off = H.t_off[q];
        int g0 = 0, g1 = 0;
        if (!single) {
            g0 = sd_owner(P.shards, (uint64_t)((int64_t)(base + lo) + off));
            g1 = sd_owner(P.shards, (uint64_t)((int64_t)(base + hi - 1) + off));
        }
        SdNbEntry e;
        e.J = P.Jhop[q];
        e.ptr = psi.base[g0] + (int64_t)NC * ((int64_t)base + off);
        if (g0 != g1) {                                  // range straddles a shard boundary: slow path,
            SD_ATOMIC_OR(&H.mixed_mask, 1u << q);        // the entry stays in its slot with J = 0
            e.J = 0.0;
            e.ptr = psi.base[P.shards.rank] + (int64_t)NC * (int64_t)base;
        }
        if (q + 1 < A) {
            const uint32_t below = (1u << q) - 1u;
            const int slot = ((farbits >> q) & 1u) ? SD_POPC32(act_far & below) : nfar + SD_POPC32(act_near & below);
            H.nbf[slot] = e;
        } else {
            H.cross = e;
        }
    }
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------- tile header, device version
// The kernel is persistent (one CTA walks tiles key, key+grid, ...), so the header of the NEXT tile
// is computed by warp 0 in two halves around phase 1 of the current tile: sd_tile_hdr_issue starts
// the two binomial loads of every prefix site, sd_tile_hdr_finish reduces them with shuffles and
// writes the header.  The L2 latency of the lookups (~2000 cycles under load) hides behind phase 1.
// Same arithmetic as the host-callable sd_tile_phase0a/0b used by the CPU emulation.
struct SdHdrRegs {
    uint64_t tb, dl;                                      // C(L-1-q, .) rank term, C(L-2-q, .) bond shift
};
__device__ __forceinline__ uint64_t sd_ldg_u64_pinned(const uint64_t *p) {
    uint64_t x;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(x) : "l"(p) : "memory");
    return x;
}
__device__ __forceinline__ uint64_t sd_warp_sum_u64(uint64_t x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double sd_warp_sum_f64(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// warp 0, lane q = prefix position q
__device__ __forceinline__ SdHdrRegs sd_tile_hdr_issue(const SdTileParams &P, uint64_t key, unsigned lane) {
    SdHdrRegs r;
    r.tb = 0; r.dl = 0;
    const int L = P.L, k = P.k, A = P.A, q = (int)lane;
    if (q >= A) return r;
    const uint64_t Pb = sd_tile_prefix_bits(key, A);
    const int js = k - __popcll(Pb);
    if (js < 0 || js > P.B) return r;
    const int bit = (int)((Pb >> q) & 1ULL);
    const int below = __popcll(Pb & ((1ULL << q) - 1));
    if (!bit) {
        const int n = L - 1 - q, rr = k - below - 1;
        if (rr >= 0 && rr <= n) r.tb = sd_ldg_u64_pinned(P.binom + n * SD_BINOM_DIM + rr);
    }
    if (q + 1 < A) {
        const int bn = (int)((Pb >> (q + 1)) & 1ULL);
        if (bit != bn && P.Jhop[q] != 0.0) {
            const int n = L - 2 - q, rr = k - (below + bit + bn);
            if (rr >= 0 && rr <= n) r.dl = sd_ldg_u64_pinned(P.binom + n * SD_BINOM_DIM + rr);
        }
    }
    return r;
}

template <int NC>
__device__ __forceinline__ void sd_tile_hdr_finish(const SdTileParams &P, uint64_t key, SdTileHdr &H,
                                                   const SdVecView &psi, unsigned lane, const SdHdrRegs &r) {
    const int k = P.k, A = P.A, B = P.B, T = P.T, q = (int)lane;
    const uint64_t Pb = sd_tile_prefix_bits(key, A);
    const int js = k - __popcll(Pb);
    const bool valid = (js >= 0 && js <= B);
    if (lane == 0) {
        H.valid = valid ? 1 : 0;
        H.js = js;
        H.mixed_mask = 0;
    }
    if (!valid) return;
    const SdJsInfo &I = P.js[js];
    const uint32_t size = I.size;
    if (lane == 0) {
        H.size = size; H.nslots = I.nslots; H.item_off = I.item_off; H.perm_off = I.perm_off;
    }
    if (q <= T + 1) H.cls_base[q] = I.cls_base[q];
    double d = 0.0;
    int64_t off = 0;
    uint32_t lo = 0, hi = 0;
    if (q < A) {
        const int bit = (int)((Pb >> q) & 1ULL);
        const double sq = bit ? 0.5 : -0.5;
        d = P.h[q] * sq;
        if (q + 1 < A) {
            const int bn = (int)((Pb >> (q + 1)) & 1ULL);
            d += P.Jz[q] * sq * (bn ? 0.5 : -0.5);
            if (bit != bn && P.Jhop[q] != 0.0) {
                off = bit ? (int64_t)r.dl : -(int64_t)r.dl;
                hi = size;
            }
        } else if (P.Jhop[q] != 0.0) {
            const uint32_t n1 = I.n1;
            if (bit) {
                if (n1 < size) { off = (int64_t)I.ncross; lo = n1; hi = size; }
            } else if (n1 > 0) {
                off = -(int64_t)n1; lo = 0; hi = n1;
            }
        }
        H.t_off[q] = off; H.t_lo[q] = lo; H.t_hi[q] = hi;
    }
    const uint64_t base = sd_warp_sum_u64(r.tb);
    const double dpre = sd_warp_sum_f64(d);
    const uint32_t act = sd_tile_active_mask(P, Pb);
    const uint32_t farbits = (P.qfar >= 32) ? ~0u : ((1u << P.qfar) - 1u);
    const uint32_t act_far = act & farbits, act_near = act & ~farbits;
    const int nfar = __popc(act_far);
    if (lane == 0) {
        H.base = base;
        if (A > 0) {
            const double sl = ((Pb >> (A - 1)) & 1ULL) ? 0.5 : -0.5;
            H.dP[0] = dpre + P.Jz[A - 1] * sl * (-0.5);
            H.dP[1] = dpre + P.Jz[A - 1] * sl * (0.5);
        } else {
            H.dP[0] = H.dP[1] = dpre;
            H.cross.ptr = nullptr;
        }
        H.nfar = nfar;
        H.nfull = nfar + __popc(act_near);
    }
    if (q < A) {
        SdNbEntry e;
        e.ptr = nullptr;
        e.J = P.Jhop[q];
        if (hi > lo) {
            int g0 = 0, g1 = 0;
            if (P.shards.world != 1) {
                g0 = sd_owner(P.shards, (uint64_t)((int64_t)(base + lo) + off));
                g1 = sd_owner(P.shards, (uint64_t)((int64_t)(base + hi - 1) + off));
            }
            e.ptr = psi.base[g0] + (int64_t)NC * ((int64_t)base + off);
            if (g0 != g1) {
                atomicOr(&H.mixed_mask, 1u << q);
                e.J = 0.0;
                e.ptr = psi.base[P.shards.rank] + (int64_t)NC * (int64_t)base;
            }
        }
        if (q + 1 < A) {
            if (hi > lo) {
                const uint32_t below = (1u << q) - 1u;
                const int slot = ((farbits >> q) & 1u) ? __popc(act_far & below) : nfar + __popc(act_near & below);
                H.nbf[slot] = e;
            }
        } else {
            H.cross = e;                                  // ptr == null when the crossing bond is inactive
        }
    }
}

// warp 1: pull the own tile and the far neighbour tiles of an already finished header into L2
template <int NC>
__device__ __forceinline__ void sd_tile_hdr_prefetch(const SdTileParams &P, const SdTileHdr &H, const SdVecView &psi,
                                                     unsigned lane) {
    if (!H.valid) return;
    const size_t bytes = (size_t)NC * 8 * H.size;
    if ((int)lane < H.nfar) {
        const double *p = H.nbf[lane].ptr;
        const bool local = (P.shards.world == 1) ||
            (p >= psi.base[P.shards.rank] + (size_t)NC * P.shards.start[P.shards.rank] &&
             p < psi.base[P.shards.rank] + (size_t)NC * P.shards.start[P.shards.rank + 1]);
        if (local) SD_PREFETCH_L2(p, bytes);
    } else if ((int)lane == H.nfar) {
        SD_PREFETCH_L2(psi.base[P.shards.rank] + (size_t)NC * H.base, bytes);
    }
}
#endif

// ---------------------------------------------------------------- phase 1
// Warp-blocked mapping: a warp owns 32*U consecutive elements, a thread's U elements are 256 B
// apart (element l0 + 32 u).  With the obvious CTA-strided mapping (elements 4 KB apart) all of a
// thread's loads fall into the same L1 set and throughput stops scaling with U; measured on B200
// (scripts/mb_streams.cu): 10 streams, 32 warps/SM: strided U=7 8.9 ms, U=13 11.0 ms; blocked U=13 7.2 ms.
//
// The apply is bound by BYTES IN FLIGHT (Little's law): load destinations are registers and the
// register file is shared with phase 2, so the partial sums g live in shared memory (where they
// must end up anyway) and ALL landing registers hold loads: two neighbour tiles x U elements are in
// flight per thread, then  sg[pos] (+)= J0 t0 + J1 t1.
template <int NC, int CNT, bool FULL, bool FAR>
SD_HD void sd_tile_issue(double (&t)[CNT][NC], const double *q, const int32_t (&idx)[CNT]) {
#pragma unroll
    for (int u = 0; u < CNT; ++u)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double *a = FULL ? q + (size_t)NC * 32 * u + c : q + (int64_t)NC * idx[u] + c;
            t[u][c] = FAR ? SD_LD_FAR(a) : SD_LD_NEAR(a);
        }
}

// one round: R neighbour tiles x CNT elements in flight, then sg[pos] (+)= sum_r J_r t_r
template <int NC, int CNT, bool FULL, bool FAR, int R>
SD_HD void sd_tile_round(double *sg, const uint32_t (&pos)[CNT], const int32_t (&idx)[CNT], uint32_t ok,
                         const SdNbEntry *e, uint32_t l0, bool first) {
    double t[R][CNT][NC];
    double J[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const SdNbEntry er = e[r];
        J[r] = er.J;
        sd_tile_issue<NC, CNT, FULL, FAR>(t[r], er.ptr + (size_t)NC * l0, idx);
    }
#pragma unroll
    for (int u = 0; u < CNT; ++u) {
        if (!FULL && !((ok >> u) & 1u)) continue;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            double s = J[0] * t[0][u][c];
#pragma unroll
            for (int r = 1; r < R; ++r) s += J[r] * t[r][u][c];
            double *d = sg + (size_t)pos[u] * NC + c;
            *d = first ? s : *d + s;
        }
    }
}

// entries e[0..n) of one cache policy, SD_TILE_EPR per round; `first` = sg not yet initialised
template <int NC, int CNT, bool FULL, bool FAR>
SD_HD void sd_tile_stream(double *sg, const uint32_t (&pos)[CNT], const int32_t (&idx)[CNT], uint32_t ok,
                          const SdNbEntry *e, int n, uint32_t l0, bool &first) {
    int i = 0;
#pragma unroll 1
    for (; i + 3 <= n; i += 3) {
        sd_tile_round<NC, CNT, FULL, FAR, 3>(sg, pos, idx, ok, e + i, l0, first);
        first = false;
    }
    if (n - i == 2) {
        sd_tile_round<NC, CNT, FULL, FAR, 2>(sg, pos, idx, ok, e + i, l0, first);
        first = false;
    } else if (n - i == 1) {
        sd_tile_round<NC, CNT, FULL, FAR, 1>(sg, pos, idx, ok, e + i, l0, first);
        first = false;
    }
}

// One warp chunk: elements l0 + 32 u, u < CNT.  FULL: every element of every lane is inside the tile
// (unpredicated loads, immediate offsets); otherwise indices are clamped into the tile and only the
// stores are predicated.
template <int NC, int CNT, bool FULL>
SD_HD void sd_tile_phase1_chunk(const SdTileParams &P, const SdTileView<NC> &v, const SdVecView &psi, uint32_t l0) {
    const SdTileHdr &H = *v.hdr;
    const uint32_t size = H.size;
    int32_t idx[CNT];                                     // element offset relative to l0 (clamped lanes: negative)
    uint32_t ok = 0u;                                     // bit u: element u of this lane exists
#pragma unroll
    for (int u = 0; u < CNT; ++u) {
        const uint32_t l = l0 + 32u * u;
        idx[u] = FULL ? 32 * u : (int32_t)(l < size ? l : size - 1) - (int32_t)l0;
        if (FULL || l < size) ok |= 1u << u;
    }
    SD_PTICK_INIT();
    const uint16_t *perm = P.perm + H.perm_off + l0;
    uint32_t pos[CNT];
#pragma unroll
    for (int u = 0; u < CNT; ++u) pos[u] = perm[idx[u]];
    {                                                     // own tile: global -> smem (class-major position)
        const double *own = psi.base[P.shards.rank] + (int64_t)NC * (int64_t)(H.base + l0);
#pragma unroll
        for (int u = 0; u < CNT; ++u)
            if (FULL || ((ok >> u) & 1u)) SD_CP_ASYNC(v.spsi + (size_t)pos[u] * NC, own + (int64_t)NC * idx[u], NC);
    }
    SD_PTICK(8);                                          // perm round trip + cp.async issue
    bool first = true;
    const int nfar = H.nfar, nfull = H.nfull;
    sd_tile_stream<NC, CNT, FULL, true>(v.sg, pos, idx, ok, H.nbf, nfar, l0, first);
    SD_PTICK(9);                                          // far streams
    sd_tile_stream<NC, CNT, FULL, false>(v.sg, pos, idx, ok, H.nbf + nfar, nfull - nfar, l0, first);
    SD_PTICK(10);                                         // near streams
    if (first) {                                          // no active prefix bond at all
#pragma unroll
        for (int u = 0; u < CNT; ++u)
            if (FULL || ((ok >> u) & 1u)) {
#pragma unroll
                for (int c = 0; c < NC; ++c) v.sg[(size_t)pos[u] * NC + c] = 0.0;
            }
    }
    const SdNbEntry ce = H.cross;                         // prefix|suffix crossing bond: sub-range [lo, hi)
    if (ce.ptr) {
        const int qc = P.A - 1;
        const uint32_t lo = H.t_lo[qc], hi = H.t_hi[qc];
        const double *q = ce.ptr + (size_t)NC * l0;
#pragma unroll
        for (int u = 0; u < CNT; ++u) {
            const uint32_t l = l0 + 32u * u;
            if (l >= lo && l < hi) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    v.sg[(size_t)pos[u] * NC + c] += ce.J * SD_LD_NEAR(q + (size_t)NC * 32 * u + c);
            }
        }
    }
    SD_PTICK(11);                                         // crossing bond
    for (uint32_t mm = H.mixed_mask; mm; mm &= mm - 1) {  // neighbour range straddles a shard boundary (rare)
        int qq = 0;
        while (!((mm >> qq) & 1u)) ++qq;
        const int64_t off = H.t_off[qq];
        const double J = P.Jhop[qq];
        const uint32_t lo = H.t_lo[qq], hi = H.t_hi[qq];
#pragma unroll
        for (int u = 0; u < CNT; ++u) {
            const uint32_t l = l0 + 32u * u;
            if (l >= lo && l < hi) {
                const uint64_t r = (uint64_t)((int64_t)(H.base + l) + off);
                const double *q = psi.base[sd_owner(P.shards, r)] + (size_t)NC * r;
#pragma unroll
                for (int c = 0; c < NC; ++c) v.sg[(size_t)pos[u] * NC + c] += J * q[c];
            }
        }
    }
}

template <int NC>
SD_HD void sd_tile_phase1(const SdTileParams &P, const SdTileView<NC> &v, const SdVecView &psi,
                          unsigned tid, unsigned nthreads) {
    constexpr int U = SD_TILE_U(NC);
    const uint32_t size = v.hdr->size;
    const uint32_t lane = tid & 31u, warp = tid >> 5, nwarps = nthreads >> 5;
    for (uint32_t c0 = warp * 32u * U; c0 < size; c0 += nwarps * 32u * U) {
        if (c0 + 32u * U <= size) sd_tile_phase1_chunk<NC, U, true>(P, v, psi, c0 + lane);
        else sd_tile_phase1_chunk<NC, U, false>(P, v, psi, c0 + lane);
    }
    SD_PTICK_INIT();
    SD_CP_ASYNC_WAIT_ALL();                               // own-tile copies landed before the CTA barrier
    SD_PTICK(12);
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
SD_HD void sd_tile_block(const SdTileParams &P, const SdTileView<NC> &v, const SdItem &it) {
    constexpr int NT = sd_cbinom(T, JT);
    constexpr int NTP = NT | 1;                      // odd pitch: conflict-free across a warp
    const SdTileHdr &H = *v.hdr;
    const int M = P.M, A = P.A;
    const unsigned c = it.c;
    const uint32_t u = it.u;
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
    const double dthread = H.dP[c_first] + it.dmid;
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
            if constexpr (JT < T && NT - n1 > 0) {             // our first bit 0 -> partner class JT+1, first part
                constexpr int NTP2 = sd_cbinom(T, JT + 1) | 1;
                const double *sp = v.spsi + (size_t)(H.cls_base[JT + 1] + (uint32_t)it.u2 * NTP2) * NC;
#pragma unroll
                for (int t = n1; t < NT; ++t)
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) acc[t * NC + cc] += J * sp[(t - n1) * NC + cc];
            }
        } else {
            if constexpr (JT > 0 && n1 > 0) {                  // our first bit 1 -> partner class JT-1, second part
                constexpr int NTP2 = sd_cbinom(T, JT - 1) | 1;
                constexpr int n1p = sd_cbinom(T - 1, JT - 2);
                const double *sp = v.spsi + (size_t)(H.cls_base[JT - 1] + (uint32_t)it.u2 * NTP2) * NC;
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
    static SD_HD void run(const SdTileParams &P, const SdTileView<NC> &v, const SdItem &it) {
        if (it.jt == JT) sd_tile_block<NC, T, JT>(P, v, it);
        else if constexpr (JT > 0) SdTileDispatch<NC, T, JT - 1>::run(P, v, it);
    }
};

template <int NC, int T>
SD_HD void sd_tile_phase2(const SdTileParams &P, const SdTileView<NC> &v, unsigned tid,
                          unsigned nthreads, const SdItem &item0) {
    const SdTileHdr &H = *v.hdr;
    const uint32_t nslots = H.nslots;
    const SdItem *items = P.items + H.item_off;
    SdItem cur = item0;                                   // prefetched in phase 0
#pragma unroll 1
    for (uint32_t s = tid; s < nslots; s += nthreads) {
        SdItem nxt;
        nxt.c = 0xFFFFu;
        if (s + nthreads < nslots) nxt = items[s + nthreads];             // lookahead hides the L2 latency
        if (cur.c != 0xFFFFu) SdTileDispatch<NC, T, T>::run(P, v, cur);
        cur = nxt;
    }
}

// ---------------------------------------------------------------- phase 3
template <int NC, bool PLAIN, int CNT, bool FULL>
SD_HD void sd_tile_phase3_chunk(const SdTileParams &P, const SdTileView<NC> &v, double *out_vbase,
                                const SdEpi &epi, uint32_t l0, double (&red)[SD_NSLOT]) {
    const SdTileHdr &H = *v.hdr;
    const uint32_t size = H.size;
    const uint16_t *perm = P.perm + H.perm_off + l0;
    double *o = out_vbase + (int64_t)NC * (int64_t)(H.base + l0);
    uint32_t pos[CNT];
#pragma unroll
    for (int u = 0; u < CNT; ++u) pos[u] = (FULL || l0 + 32u * u < size) ? perm[32 * u] : 0u;
#pragma unroll
    for (int u = 0; u < CNT; ++u) {
        if (!FULL && l0 + 32u * u >= size) continue;
        SdVal<NC> h;
#pragma unroll
        for (int c = 0; c < NC; ++c) h.c[c] = v.sg[(size_t)pos[u] * NC + c];
        if (PLAIN) {
#pragma unroll
            for (int c = 0; c < NC; ++c) SD_ST_STREAM(o + (size_t)NC * 32 * u + c, h.c[c]);
        } else {
            SdVal<NC> p;
#pragma unroll
            for (int c = 0; c < NC; ++c) p.c[c] = v.spsi[(size_t)pos[u] * NC + c];
            const uint64_t li = H.base + l0 + 32u * u - P.shards.start[P.shards.rank];
            const SdVal<NC> r = sd_epilogue<NC>(epi, h, p, li, red);
#pragma unroll
            for (int c = 0; c < NC; ++c) o[(size_t)NC * 32 * u + c] = r.c[c];
        }
    }
}

template <int NC, bool PLAIN>
SD_HD void sd_tile_phase3(const SdTileParams &P, const SdTileView<NC> &v, double *out_vbase,
                          const SdEpi &epi, unsigned tid, unsigned nthreads,
                          double (&red)[SD_NSLOT]) {
    constexpr int U = 8;
    const uint32_t size = v.hdr->size;
    const uint32_t lane = tid & 31u, warp = tid >> 5, nwarps = nthreads >> 5;
    for (uint32_t c0 = warp * 32u * U; c0 < size; c0 += nwarps * 32u * U) {
        if (c0 + 32u * U <= size) sd_tile_phase3_chunk<NC, PLAIN, U, true>(P, v, out_vbase, epi, c0 + lane, red);
        else sd_tile_phase3_chunk<NC, PLAIN, U, false>(P, v, out_vbase, epi, c0 + lane, red);
    }
}
