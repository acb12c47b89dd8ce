// sd_blkr.h -- "ring" variant of the block-layout H.psi kernel (f64; same layout, tables and tile
// header as sd_blk.h; reference Hamiltonian.jl:211-273 on the bond lists XXZChain builds).
//
// What changes against sd_blk_apply_kernel is WHO moves the neighbour tiles and WHERE the partial sums live:
//   * the accumulators of a whole tile stay in registers from the first neighbour tile to the store: a
//     consumer warp owns a fixed set of work items of the tile (one item of <= 5 slots in register group A,
//     one item of 3 slots or up to three single-slot items in group B; the packing per suffix popcount is a
//     host table, sd_blkr_host.h), i.e. 8 double2 accumulators per lane;
//   * every tile a tile needs -- the partner tiles of its active prefix bonds, the prefix|mid crossing
//     partner, and finally the tile itself -- is brought into a ring of SD_BLKR_NB shared-memory buffers by TMA
//     bulk copies (cp.async.bulk + mbarrier complete_tx) issued by the producer warp, so the bytes in flight
//     (three buffers ~ 150 KB per SM) do not depend on how far the consumer warps have got, and a neighbour
//     tile on another GPU is fetched over NVLink by the same bulk copies;
//   * the consumers turn each ring entry into  acc[s] += J * buf[slot s]  (one LDS.128 and two DFMAs per
//     slot, addresses fixed per tile), then run diagonal / tail / mid / mid|tail hops, the fused epilogue and
//     the store when the tile itself arrives (last entry of the tile).
// There is no per-item scheduling, no per-load address arithmetic or predicate and no global-load latency in
// the consumer warps.  Cost: neighbour bytes cross the shared-memory port twice (TMA write + LDS).
//
// Ring protocol.  Entries are numbered e = 0, 1, 2, ... over all tiles of the CTA: tile t contributes
// ntot(t) neighbour entries followed by one own entry.  Entry e lives in slot e % NB; full[slot] completes when
// the TMA bytes of the entry have landed (phase parity (e / NB) & 1), empty[slot] when all consumer warps have
// released it.  The header of tile number t (per CTA) lives in hdr[t % NB]; the producer writes it after the
// wait on empty[] of the tile's FIRST entry, which implies every consumer warp has released the own entry of
// tile t - NB (each tile has at least one entry and warps release entries in order), and consumers read it after
// the wait on full[] of the same first entry.  The end of the tile list is a header with valid = -1 in an
// otherwise empty entry.
#pragma once
#include "sd_blk.h"

#define SD_BLKR_NB 4                 // ring slots (and tile headers) per CTA
#define SD_BLKR_NSLOT_A 5            // register group A: one item of up to 5 slots
#define SD_BLKR_NSLOT_B 3            // register group B: one 3-slot item or up to three 1-slot items
#define SD_BLKR_NONE 0xFFFFu

// Work of one consumer warp on a tile of suffix popcount js: item codes jt << 12 | unit-in-class.
struct SdBlkrWarp {
    uint16_t a;                      // group A item (any class), SD_BLKR_NONE: none
    uint16_t b[3];                   // group B: b[0] of a 3-slot class (jt = 1, 4) alone, or up to three 1-slot items (jt = 0, 5)
};

// Tile header: SdBlkHdr with one reduction entry per consumer warp (shared memory is the scarce resource here).
struct SdBlkrHdr {
    uint64_t base;
    int js, jsx;
    int valid;
    int nnb, nfar;
    int ntot;
    int bP;
    unsigned next_unit;              // unused (static work assignment)
    unsigned done_units;             // consumer warps that finished the tile
    unsigned tile_index;
    double dP[2];
    double Jx;
    const double *xptr;
    const double *nb_ptr[SD_BLK_MAXA + 8];
    double nb_J[SD_BLK_MAXA + 8];
    double usum[SD_NSLOT][16];       // per-warp reduction results, summed in warp order by the last warp
};

SD_HD int sd_blkr_ec(int jt) {       // slots of an f64 item of class jt: (C(5, jt) + 1) / 2
    return (jt == 0 || jt == SD_BLK_T) ? 1 : ((jt == 1 || jt == SD_BLK_T - 1) ? 3 : 5);
}

// ------------------------------------------------------------------ per-lane state of one tile
struct SdBlkrLane {
    double2 acc[SD_BLKR_NSLOT_A + SD_BLKR_NSLOT_B];
    int jtA;                         // -1: no item, or lane beyond the class
    int jtB[3];
    int ecA, ecB0;                   // slots of the group A item / of item b[0] (0: none); b[1], b[2] are 1-slot items
    uint32_t uA, uB[3];              // the lane's mid configuration (class-local index)
    uint32_t baseA, ssA;             // doubles from the tile start to slot 0 of the block / between slots
    uint32_t baseB[3], ssB;
};

SD_HD void sd_blkr_begin(SdBlkrLane &S, const SdBlkJs &I, SdBlkrWarp w, unsigned lane) {
#pragma unroll
    for (int s = 0; s < SD_BLKR_NSLOT_A + SD_BLKR_NSLOT_B; ++s) S.acc[s] = make_double2(0.0, 0.0);
    S.jtA = -1; S.uA = 0; S.baseA = 0; S.ssA = 0; S.ssB = 0; S.ecA = 0; S.ecB0 = 0;
    if (w.a != SD_BLKR_NONE) {
        const int jt = (int)(w.a >> 12);
        const SdBlkCls c = I.cls[jt];
        const uint32_t u = (uint32_t)(w.a & 0xFFFu) * 32u + lane;
        if (u < c.nblk) { S.jtA = jt; S.ecA = sd_blkr_ec(jt); S.uA = u; S.baseA = c.cb + 2u * u; S.ssA = 2u * c.pitch; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        S.jtB[i] = -1; S.uB[i] = 0; S.baseB[i] = 0;
        if (w.b[i] != SD_BLKR_NONE) {
            const int jt = (int)(w.b[i] >> 12);
            const SdBlkCls c = I.cls[jt];
            const uint32_t u = (uint32_t)(w.b[i] & 0xFFFu) * 32u + lane;
            if (u < c.nblk) {
                S.jtB[i] = jt; S.uB[i] = u; S.baseB[i] = c.cb + 2u * u;
                if (i == 0) { S.ssB = 2u * c.pitch; S.ecB0 = sd_blkr_ec(jt); }
            }
        }
    }
}

// acc[O .. O+EC) += J * (slots of one block in `buf`);  HALF: the last slot is a plain double at base + (EC-1)*ss - u
template <int O, int EC, bool HALF>
SD_HD void sd_blkr_axpy(SdBlkrLane &S, const double *buf, uint32_t base, uint32_t ss, uint32_t u, double J) {
#pragma unroll
    for (int s = 0; s < EC; ++s) {
        if (HALF && s == EC - 1) {
            S.acc[O + s].x += J * *(buf + base + s * ss - u);
        } else {
            const double2 t = *(const double2 *)(buf + base + s * ss);
            S.acc[O + s].x += J * t.x;
            S.acc[O + s].y += J * t.y;
        }
    }
}
template <int O>
SD_HD void sd_blkr_axpy_any(SdBlkrLane &S, const double *buf, int ec, uint32_t base, uint32_t ss, uint32_t u, double J) {
    if (O == 0 && ec == 5) sd_blkr_axpy<O, (O == 0 ? 5 : 1), false>(S, buf, base, ss, u, J);
    else if (O <= 5 && ec == 3) sd_blkr_axpy<O, (O <= 5 ? 3 : 1), true>(S, buf, base, ss, u, J);
    else sd_blkr_axpy<O, 1, true>(S, buf, base, ss, u, J);
}

// One neighbour entry: n < nnb is the partner tile of an active prefix bond (same suffix popcount, same element
// order: the lane's own offsets apply); n == nnb is the prefix|mid crossing partner (suffix popcount jsx, same class,
// block index shifted by the number of blocks whose first mid bit is set), which only concerns lanes whose first
// mid bit differs from the last prefix bit.
SD_HD void sd_blkr_stream(SdBlkrLane &S, const SdBlkJs *jstab, const SdBlkrHdr &H, const double *buf, int n) {
    const double J = H.nb_J[n];
    if (n < H.nnb) {
        if (S.ecA > 0) sd_blkr_axpy_any<0>(S, buf, S.ecA, S.baseA, S.ssA, S.uA, J);
        if (S.ecB0 > 0) sd_blkr_axpy_any<5>(S, buf, S.ecB0, S.baseB[0], S.ssB, S.uB[0], J);
        if (S.jtB[1] >= 0) sd_blkr_axpy<6, 1, true>(S, buf, S.baseB[1], 0u, S.uB[1], J);
        if (S.jtB[2] >= 0) sd_blkr_axpy<7, 1, true>(S, buf, S.baseB[2], 0u, S.uB[2], J);
        return;
    }
    const SdBlkJs &I = jstab[H.js];
    const SdBlkJs &Ix = jstab[H.jsx];
    const bool bP = H.bP != 0;
#define SD_BLKR_CROSS(O_, jt_, u_, EC_)                                                           \
    do {                                                                                          \
        if ((jt_) >= 0) {                                                                         \
            const uint32_t n1_ = I.cls[(jt_)].n1;                                                 \
            if (((u_) < n1_) != bP) {                                                             \
                const SdBlkCls cx_ = Ix.cls[(jt_)];                                               \
                const uint32_t xu_ = bP ? (u_) - n1_ : cx_.n1 + (u_);                             \
                if ((EC_) >= 0) sd_blkr_axpy_any<O_>(S, buf, (EC_), cx_.cb + 2u * xu_, 2u * cx_.pitch, xu_, J); \
                else sd_blkr_axpy<O_, 1, true>(S, buf, cx_.cb + 2u * xu_, 0u, xu_, J);            \
            }                                                                                     \
        }                                                                                         \
    } while (0)
    SD_BLKR_CROSS(0, S.jtA, S.uA, S.ecA);
    SD_BLKR_CROSS(5, S.jtB[0], S.uB[0], S.ecB0);
    SD_BLKR_CROSS(6, S.jtB[1], S.uB[1], -1);
    SD_BLKR_CROSS(7, S.jtB[2], S.uB[2], -1);
#undef SD_BLKR_CROSS
}

// The same for a partner tile read straight from global memory / L2 (no ring slot): the `ndirect` nearest prefix
// entries of a tile (the last ones of the far-first list, most likely L2 hits) can bypass shared memory, which takes
// their bytes off the shared-memory port (TMA write + LDS) at the price of global-load latency in the consumer warps
// once per tile, while the ring entries of the tile are already in flight (SD_BLKR_DIRECT, unmeasured).
template <int O, int EC, bool HALF>
SD_HD void sd_blkr_axpy_g(SdBlkrLane &S, const double *g, uint32_t base, uint32_t ss, uint32_t u, double J) {
    double2 t[EC];
#pragma unroll
    for (int s = 0; s < EC; ++s) {
        if (HALF && s == EC - 1) t[s] = sd_blk_ldg_half(g + base + s * ss - u);
        else t[s] = sd_blk_ldg(g + base + s * ss);
    }
#pragma unroll
    for (int s = 0; s < EC; ++s) {
        S.acc[O + s].x += J * t[s].x;
        if (!(HALF && s == EC - 1)) S.acc[O + s].y += J * t[s].y;
    }
}
SD_HD void sd_blkr_stream_direct(SdBlkrLane &S, const SdBlkrHdr &H, int n) {
    const double J = H.nb_J[n];
    const double *g = H.nb_ptr[n];
    if (S.ecA == 5) sd_blkr_axpy_g<0, 5, false>(S, g, S.baseA, S.ssA, S.uA, J);
    else if (S.ecA == 3) sd_blkr_axpy_g<0, 3, true>(S, g, S.baseA, S.ssA, S.uA, J);
    else if (S.ecA == 1) sd_blkr_axpy_g<0, 1, true>(S, g, S.baseA, S.ssA, S.uA, J);
    if (S.ecB0 == 3) sd_blkr_axpy_g<5, 3, true>(S, g, S.baseB[0], S.ssB, S.uB[0], J);
    else if (S.ecB0 == 1) sd_blkr_axpy_g<5, 1, true>(S, g, S.baseB[0], 0u, S.uB[0], J);
    if (S.jtB[1] >= 0) sd_blkr_axpy_g<6, 1, true>(S, g, S.baseB[1], 0u, S.uB[1], J);
    if (S.jtB[2] >= 0) sd_blkr_axpy_g<7, 1, true>(S, g, S.baseB[2], 0u, S.uB[2], J);
}
// ring entries of a tile: prefix entries [0, nring), then the crossing entry (if any), then the tile itself;
// prefix entries [nring, nnb) are read directly by the consumers
SD_HD int sd_blkr_nring(const SdBlkrHdr &H, int ndirect) {
    const int nd = ndirect < H.nnb ? ndirect : H.nnb;
    return H.nnb - (nd > 0 ? nd : 0);
}

// ---- own tile.  Everything that depends on the class JT (tail popcount) is resolved at compile time: the tail
// configurations of the class, which of them hop where inside the tail, and which of them cross the mid|tail bond.
template <int JT, int e, int q>
struct SdBlkrTailHop {                                              // tail-internal bond q of tail configuration e
    template <int EC>
    static SD_HD void run(double2 (&a)[EC], const double2 (&own)[EC], const double (&Jt)[SD_BLK_T - 1]) {
        constexpr unsigned cfg = sd_tail_cfg(SD_BLK_T, JT, e);
        if constexpr ((((cfg >> q) ^ (cfg >> (q + 1))) & 1u) != 0) {
            constexpr int e2 = sd_tail_rank(SD_BLK_T, JT, cfg ^ (3u << q));
            SD_BLK_EL(a, e, 1) += Jt[q] * ((e2 & 1) ? own[e2 >> 1].y : own[e2 >> 1].x);
        }
        if constexpr (q + 2 < SD_BLK_T) SdBlkrTailHop<JT, e, q + 1>::run(a, own, Jt);
    }
};
template <int JT, int e>
struct SdBlkrTailRow {                                              // diagonal + tail hops of tail configuration e
    template <int EC>
    static SD_HD void run(double2 (&a)[EC], const double2 (&own)[EC], const double (&Jt)[SD_BLK_T - 1], const double *dtail,
                          double dplus, double dminus) {
        constexpr unsigned cfg = sd_tail_cfg(SD_BLK_T, JT, e);
        // dplus / dminus: tail bit 0 equal to / different from the last mid bit (zz of the mid|tail bond)
        const double d = ((cfg & 1u) ? dplus : dminus) + dtail[cfg];
        SD_BLK_EL(a, e, 1) += d * ((e & 1) ? own[e >> 1].y : own[e >> 1].x);
        SdBlkrTailHop<JT, e, 0>::run(a, own, Jt);
        if constexpr (e + 1 < sd_cbinom(SD_BLK_T, JT)) SdBlkrTailRow<JT, e + 1>::run(a, own, Jt, dtail, dplus, dminus);
    }
};
// mid|tail crossing bond.  Tail configurations with bit 0 set come first in a class (n1 = C(T-1, JT-1) of them).
// UP: last mid bit set & tail bit 0 clear -> class JT+1, configuration e - n1;
// else: last mid bit clear & tail bit 0 set -> class JT-1, configuration C(T-1, JT-2) + e.
template <int JT, bool UP, int e>
struct SdBlkrCross {
    template <int EC>
    static SD_HD void run(double2 (&a)[EC], const double *sp, uint32_t s2, uint32_t u2x, double J) {
        constexpr int T = SD_BLK_T;
        constexpr int NT = sd_cbinom(T, JT), n1 = sd_cbinom(T - 1, JT - 1);
        constexpr int JT2 = UP ? JT + 1 : JT - 1;
        constexpr int NT2 = sd_cbinom(T, JT2);
        constexpr bool act = UP ? (e >= n1) : (e < n1);
        if constexpr (act && JT2 >= 0 && JT2 <= T) {
            constexpr int e2 = UP ? e - n1 : e + sd_cbinom(T - 1, JT - 2);
            static_assert(e2 >= 0 && e2 < NT2, "partner tail configuration");
            // the last configuration of an odd class is a plain row of doubles
            const double t = ((NT2 & 1) && e2 == NT2 - 1) ? *(sp + (e2 >> 1) * s2 - u2x) : *(sp + (e2 >> 1) * s2 + (e2 & 1));
            SD_BLK_EL(a, e, 1) += J * t;
        }
        if constexpr (e + 1 < NT) SdBlkrCross<JT, UP, e + 1>::run(a, sp, s2, u2x, J);
    }
};

// The tile itself has arrived in `tb`: diagonal, tail-internal hops, mid-internal hops, mid|tail crossing bond,
// fused epilogue and store of ONE item of class JT whose neighbour sums are in a[].
template <int JT, bool PLAIN>
SD_HD void sd_blkr_own_item(const SdBlkCtx &X, const SdBlkrHdr &H, const double *tb, uint32_t u,
                            double2 (&a)[(sd_cbinom(SD_BLK_T, JT) + 1) / 2], double (&red)[SD_NSLOT]) {
    constexpr int T = SD_BLK_T, M = SD_BLK_M;
    constexpr int NT = sd_cbinom(T, JT), EC = (NT + 1) / 2;
    constexpr bool HALF = (NT & 1) != 0;
    static_assert(SD_BLK_T == 5, "item shapes are written for T = 5");
    const SdBlkParams &P = *X.P;
    const SdBlkJs &I = X.js[H.js];
    const SdBlkCls cls = I.cls[JT];
    const uint32_t ss = 2u * cls.pitch;
    const uint32_t off0 = cls.cb + 2u * u;
    const uint4 it = sd_blk_ld_item(P.items + cls.item_off + u);
    const bool c0 = u < cls.n1;
    const unsigned cmid = it.w & ((1u << M) - 1u);
    const bool clast = (cmid >> (M - 1)) & 1u;
    double2 own[EC];
#pragma unroll
    for (int s = 0; s < EC; ++s) {
        if (HALF && s == EC - 1) own[s] = make_double2(*(tb + off0 + s * ss - u), 0.0);
        else own[s] = *(const double2 *)(tb + off0 + s * ss);
    }
    {   // diagonal + tail-internal hops (registers)
        const double d0 = H.dP[c0 ? 1 : 0] + X.dmid[cmid];
        const double dplus = d0 + X.qx, dminus = d0 - X.qx;
        double Jt[T - 1];
#pragma unroll
        for (int q = 0; q < T - 1; ++q) Jt[q] = X.Jhop[P.A + M + q];
        // tail bit 0 set: equal to the last mid bit iff clast
        SdBlkrTailRow<JT, 0>::run(a, own, Jt, X.dtail, clast ? dplus : dminus, clast ? dminus : dplus);
    }
    {   // mid-internal hops: the whole block moves to block nb[pm] of the same class; active bonds from the item mask
        const double *cbp = tb + cls.cb;
        const uint64_t lo = (uint64_t)it.x | ((uint64_t)it.y << 32);
        uint32_t am = (P.dbg & 2) ? 0u : ((it.z >> 8) & 0xFFFFu);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        while (am != 0u) {
            const int pm = SD_POPC32((am & (0u - am)) - 1u);
            am &= am - 1u;
            const unsigned nbu = pm < 8 ? (unsigned)((lo >> (8 * pm)) & 0xFFu) : (it.z & 0xFFu);
            const double J = X.Jhop[P.A + pm];
            const double *sp = cbp + 2u * nbu;
#pragma unroll
            for (int s = 0; s < EC; ++s) {
                if (HALF && s == EC - 1) {
                    a[s].x += J * *(sp + s * ss - nbu);
                } else {
                    const double2 t = *(const double2 *)(sp + s * ss);
                    a[s].x += J * t.x;
                    a[s].y += J * t.y;
                }
            }
        }
    }
    {   // mid|tail crossing bond
        const double J = X.Jhop[P.A + M - 1];
        const uint32_t u2x = it.w >> 16;
        if (clast) {
            if constexpr (JT + 1 <= T) {
                const SdBlkCls c2 = I.cls[JT + 1];
                SdBlkrCross<JT, true, 0>::run(a, tb + c2.cb + 2u * u2x, 2u * c2.pitch, u2x, J);
            }
        } else {
            if constexpr (JT - 1 >= 0) {
                const SdBlkCls c2 = I.cls[JT - 1];
                SdBlkrCross<JT, false, 0>::run(a, tb + c2.cb + 2u * u2x, 2u * c2.pitch, u2x, J);
            }
        }
    }
    // fused epilogue + store
    const uint64_t ld0 = (H.base - X.pstart_local) + off0;           // doubles from the start of the local shard
    double *o = X.out_local + ld0;
    if (PLAIN) {
        if (P.dbg & 4) { if (a[0].x == 1.2345e300) sd_blk_stg(o, a[0]); return; }
#pragma unroll
        for (int s = 0; s < EC; ++s) {
            if (HALF && s == EC - 1) sd_blk_stg_half(o + s * ss - u, a[s].x);
            else sd_blk_stg(o + s * ss, a[s]);
        }
    } else {
        const SdEpi &E = *X.epi;
#pragma unroll
        for (int s = 0; s < EC; ++s) {
            SdVal<1> hh, pp;
            if (HALF && s == EC - 1) {
                const uint64_t ld = ld0 + (uint64_t)s * ss - u;
                hh.c[0] = a[s].x; pp.c[0] = own[s].x;
                const SdVal<1> r0 = sd_epilogue<1>(E, hh, pp, ld, red);
                *(o + s * ss - u) = r0.c[0];
                continue;
            }
            const uint64_t ld = ld0 + (uint64_t)s * ss;
            hh.c[0] = a[s].x; pp.c[0] = own[s].x;
            const SdVal<1> r0 = sd_epilogue<1>(E, hh, pp, ld, red);
            hh.c[0] = a[s].y; pp.c[0] = own[s].y;
            const SdVal<1> r1 = sd_epilogue<1>(E, hh, pp, ld + 1, red);
            *(double2 *)(o + s * ss) = make_double2(r0.c[0], r1.c[0]);
        }
    }
}
template <int O, int JT, bool PLAIN>
SD_HD void sd_blkr_own_at(SdBlkrLane &S, const SdBlkCtx &X, const SdBlkrHdr &H, const double *tb, uint32_t u,
                          double (&red)[SD_NSLOT]) {
    constexpr int EC = (sd_cbinom(SD_BLK_T, JT) + 1) / 2;
    static_assert(O + EC <= SD_BLKR_NSLOT_A + SD_BLKR_NSLOT_B, "register group too small for the class");
    double2 a[EC];
#pragma unroll
    for (int s = 0; s < EC; ++s) a[s] = S.acc[O + s];
    sd_blkr_own_item<JT, PLAIN>(X, H, tb, u, a, red);
}
template <bool PLAIN>
SD_HD void sd_blkr_own(SdBlkrLane &S, const SdBlkCtx &X, const SdBlkrHdr &H, const double *tb, double (&red)[SD_NSLOT]) {
    switch (S.jtA) {                                                 // group A: any class
        case 0: sd_blkr_own_at<0, 0, PLAIN>(S, X, H, tb, S.uA, red); break;
        case 1: sd_blkr_own_at<0, 1, PLAIN>(S, X, H, tb, S.uA, red); break;
        case 2: sd_blkr_own_at<0, 2, PLAIN>(S, X, H, tb, S.uA, red); break;
        case 3: sd_blkr_own_at<0, 3, PLAIN>(S, X, H, tb, S.uA, red); break;
        case 4: sd_blkr_own_at<0, 4, PLAIN>(S, X, H, tb, S.uA, red); break;
        case 5: sd_blkr_own_at<0, 5, PLAIN>(S, X, H, tb, S.uA, red); break;
        default: break;
    }
    switch (S.jtB[0]) {                                              // group B, first item: a 3-slot or a 1-slot class
        case 0: sd_blkr_own_at<5, 0, PLAIN>(S, X, H, tb, S.uB[0], red); break;
        case 1: sd_blkr_own_at<5, 1, PLAIN>(S, X, H, tb, S.uB[0], red); break;
        case 4: sd_blkr_own_at<5, 4, PLAIN>(S, X, H, tb, S.uB[0], red); break;
        case 5: sd_blkr_own_at<5, 5, PLAIN>(S, X, H, tb, S.uB[0], red); break;
        default: break;
    }
    if (S.jtB[1] == 0) sd_blkr_own_at<6, 0, PLAIN>(S, X, H, tb, S.uB[1], red);
    else if (S.jtB[1] == SD_BLK_T) sd_blkr_own_at<6, SD_BLK_T, PLAIN>(S, X, H, tb, S.uB[1], red);
    if (S.jtB[2] == 0) sd_blkr_own_at<7, 0, PLAIN>(S, X, H, tb, S.uB[2], red);
    else if (S.jtB[2] == SD_BLK_T) sd_blkr_own_at<7, SD_BLK_T, PLAIN>(S, X, H, tb, S.uB[2], red);
}

// The crossing entry only needs, of every slot row of the partner tile, the blocks whose first mid bit is the
// opposite of the tile's last prefix bit: a contiguous range of each row (blocks with the first mid bit set come first
// in a class).  Piece i = (class jt, slot row s) in class-major order, i < SD_BLKR_NPIECE; returns false if the piece is
// empty.  off / len in doubles from the start of the partner tile, both even (TMA bulk copies move 16-byte units: the
// plain last row of an odd class is widened by at most one element on either side, inside the padded row).
#define SD_BLKR_NPIECE 18            // sum over jt of (C(5, jt) + 1) / 2
SD_HD bool sd_blkr_cross_piece(const SdBlkJs *jstab, int js, int jsx, int bP, int i, uint32_t *off, uint32_t *len) {
    int jt = 0, s = i;
#pragma unroll
    for (int j = 0; j <= SD_BLK_T; ++j) {
        const int ec = sd_blkr_ec(j);
        if (jt == j && s >= ec) { s -= ec; jt = j + 1; }
    }
    if (jt > SD_BLK_T) return false;
    const SdBlkCls c = jstab[js].cls[jt];
    const SdBlkCls cx = jstab[jsx].cls[jt];
    if (c.nblk == 0u || cx.nblk == 0u) return false;
    // bP: lanes with the first mid bit clear (u >= n1) read partner blocks [0, nblk - n1); else lanes u < n1 read [cx.n1, cx.n1 + n1)
    uint32_t x0 = bP ? 0u : cx.n1;
    uint32_t x1 = bP ? c.nblk - c.n1 : cx.n1 + c.n1;
    if (x1 <= x0) return false;
    const int ec = sd_blkr_ec(jt);
    const bool half = s == ec - 1 && ec != 5;                      // f64: classes of 1 and 5 end in a plain row
    if (half) {
        x0 &= ~1u; x1 = (x1 + 1u) & ~1u;
        *off = cx.cb + (uint32_t)(2 * ec - 2) * cx.pitch + x0;
        *len = x1 - x0;
    } else {
        *off = cx.cb + (uint32_t)s * 2u * cx.pitch + 2u * x0;
        *len = 2u * (x1 - x0);
    }
    return true;
}

// The bulk copies of ring entry n of a tile (n < ntot: neighbour entries, n == ntot: the tile itself), as producer
// LANE `lane` issues them: copy number i (i = 0, 1, ...) of that lane moves len bytes from src + off to slot + off.
// Returns false when the lane has no copy number i.  The kernel's producer warp and tests/emul walk the same list.
// Whole tiles go in 8 KB chunks dealt round robin to the lanes; the crossing entry is one row range per lane.
#define SD_BLKR_CHUNK 8192u
SD_HD bool sd_blkr_copy(const SdBlkJs *jstab, const SdBlkrHdr &H, const double *own_src, int dbg, int n, int ntot,
                        unsigned lane, unsigned i, const char **src, uint32_t *off, uint32_t *len) {
    // n: index into the header's entry list (n == ntot: the tile itself); the caller skips the direct entries
    const bool cross = n < ntot && n == H.nnb;
    *src = (const char *)(n == ntot ? own_src : H.nb_ptr[n]);
    if (cross && !(dbg & 16)) {
        uint32_t o = 0, l = 0;
        if (i != 0u || lane >= (unsigned)SD_BLKR_NPIECE || !sd_blkr_cross_piece(jstab, H.js, H.jsx, H.bP, (int)lane, &o, &l)) return false;
        *off = o * 8u; *len = l * 8u;
        return true;
    }
    const uint32_t bytes = jstab[cross ? H.jsx : H.js].size_pad * 8u;
    const uint32_t o = (lane + 32u * i) * SD_BLKR_CHUNK;
    if (o >= bytes) return false;
    *off = o; *len = (bytes - o < SD_BLKR_CHUNK) ? bytes - o : SD_BLKR_CHUNK;
    return true;
}

// shared-memory carve-up of the ring kernel
struct SdBlkrSmem {
    uint64_t *full, *empty;      // [NB] mbarriers
    SdBlkrHdr *hdr;              // [NB]
    uint64_t *W;                 // [A*(A+1)]
    SdBlkJs *js;                 // [B+1]
    SdBlkrWarp *rw;              // [(B+1)*CWARPS]
    double *dmid;                // [1 << M]
    double *dtail;               // [1 << T]
    double *Jhop;                // [L + 1]
    double *ring;                // [NB][cap]
};
SD_HD size_t sd_blkr_smem_carve(SdBlkrSmem *s, void *base, int A, int L, uint32_t cap) {
    size_t o = 0;
    auto take = [&](size_t bytes, size_t align) {
        o = (o + align - 1) & ~(align - 1);
        const size_t at = o;
        o += bytes;
        return at;
    };
    const size_t a_full = take(8 * (size_t)SD_BLKR_NB, 8), a_empty = take(8 * (size_t)SD_BLKR_NB, 8);
    const size_t a_hdr = take(sizeof(SdBlkrHdr) * (size_t)SD_BLKR_NB, 16);
    const size_t a_W = take(8 * (size_t)A * (A + 1) + 8, 8);
    const size_t a_js = take(sizeof(SdBlkJs) * (SD_BLK_B + 1), 16);
    const size_t a_rw = take(sizeof(SdBlkrWarp) * (size_t)(SD_BLK_B + 1) * SD_BLK_CWARPS, 8);
    const size_t a_dmid = take(8 * ((size_t)1 << SD_BLK_M), 16);
    const size_t a_dtail = take(8 * ((size_t)1 << SD_BLK_T), 16);
    const size_t a_J = take(8 * (size_t)(L + 1), 16);
    const size_t a_ring = take((size_t)SD_BLKR_NB * cap * 8, 128);
    if (s) {
        char *b = (char *)base;
        s->full = (uint64_t *)(b + a_full); s->empty = (uint64_t *)(b + a_empty);
        s->hdr = (SdBlkrHdr *)(b + a_hdr); s->W = (uint64_t *)(b + a_W); s->js = (SdBlkJs *)(b + a_js);
        s->rw = (SdBlkrWarp *)(b + a_rw); s->dmid = (double *)(b + a_dmid); s->dtail = (double *)(b + a_dtail);
        s->Jhop = (double *)(b + a_J); s->ring = (double *)(b + a_ring);
    }
    return (o + 127) & ~(size_t)127;
}

#if defined(__CUDACC__)
// mbarrier wait with a watchdog: a protocol error ends the launch instead of hanging the GPU.  The clock is read once
// per 256 failed polls; 2^32 cycles are about two seconds, a thousand times the longest legitimate wait.  Default: trap
// (the launch fails with an error the host reports).  Diagnostic mode (SD_BLK_DBG & 64): the first warp to time out
// records where it was stuck in ctr[2..6] (CTA, warp, entry number, tag of the wait, tile number) and raises the abort
// flag ctr[1]; every wait loop polls the flag and returns false, every caller returns, the kernel ends normally and
// the host turns the record into an error message (sd_apply_impl).
__device__ __forceinline__ bool sd_blkr_wait(uint64_t *b, unsigned parity, unsigned long long *ctr, int diag,
                                             unsigned tag, unsigned e, unsigned t) {
    const unsigned addr = sd_smem_u32(b);
    long long t0 = 0;
    for (unsigned spins = 1;; ++spins) {
        unsigned ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return true;
        if ((spins & 255u) == 0u) {
            if (diag && *(volatile unsigned long long *)(ctr + 1) != 0ULL) return false;
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > (1LL << 32)) {
                if (!diag) __trap();
                if (atomicCAS(ctr + 1, 0ULL, 1ULL) == 0ULL) {
                    ctr[2] = blockIdx.x; ctr[3] = threadIdx.x >> 5; ctr[4] = e; ctr[5] = tag; ctr[6] = t;
                    __threadfence();
                }
                return false;
            }
        }
    }
}

// bulk copy with an L2 evict-first hint: partner tiles of the far prefix bonds (entries n < nfar in rank order) miss
// L2 anyway and are not needed again soon, so they should not push the near tiles out (SD_BLK_DBG & 32, unmeasured)
__device__ __forceinline__ void sd_bulk_g2s_evict_first(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(sd_smem_u32(dst)), "l"(src), "r"(bytes), "r"(sd_smem_u32(bar)), "l"(pol) : "memory");
}

// grid = one persistent CTA per SM; tiles are handed out by a global counter (rank order, or the optional order table).
// partials: [SD_NSLOT][ntiles] per-tile sums, zero-filled by the host; each is the sum, in warp order, of the tile's
// per-warp sums, and the item -> warp assignment is a fixed table, so results are run-to-run identical.
template <bool PLAIN>
__global__ void __launch_bounds__(SD_BLK_THREADS, 1)
sd_blkr_apply_kernel(const __grid_constant__ SdBlkParams P, const __grid_constant__ SdVecView psi, double *out_local,
                     const __grid_constant__ SdEpi epi, int qfar, unsigned long long *tile_ctr, const SdBlkrWarp *rwtab,
                     int ndirect) {
    extern __shared__ __align__(128) unsigned char sd_blkr_smem[];
    SdBlkrSmem S;
    sd_blkr_smem_carve(&S, sd_blkr_smem, P.A, P.L, P.cap);
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    constexpr unsigned NB = SD_BLKR_NB;
    static_assert((NB & (NB - 1)) == 0, "ring size must be a power of two");
    // ---- one-time setup
    for (int i = (int)tid; i < P.A * (P.A + 1); i += SD_BLK_THREADS) S.W[i] = P.W[i];
    {
        const uint32_t *src = (const uint32_t *)P.js;
        uint32_t *dst = (uint32_t *)S.js;
        for (int i = (int)tid; i < (int)(sizeof(SdBlkJs) * (SD_BLK_B + 1) / 4); i += SD_BLK_THREADS) dst[i] = src[i];
    }
    for (int i = (int)tid; i < (SD_BLK_B + 1) * SD_BLK_CWARPS; i += SD_BLK_THREADS) S.rw[i] = rwtab[i];
    for (int i = (int)tid; i < (1 << SD_BLK_M); i += SD_BLK_THREADS) S.dmid[i] = P.dmid[i];
    for (int i = (int)tid; i < (1 << SD_BLK_T); i += SD_BLK_THREADS) S.dtail[i] = P.dtail[i];
    for (int i = (int)tid; i <= P.L; i += SD_BLK_THREADS) S.Jhop[i] = P.Jhop[i];
    if (tid == 0) {
        for (unsigned b = 0; b < NB; ++b) { sd_mbar_init(&S.full[b], 1); sd_mbar_init(&S.empty[b], SD_BLK_CWARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool nostream = (P.dbg & 1) != 0;
    const int diag = P.dbg & 64;

    if (warp == SD_BLK_CWARPS) {
        // ================= producer warp: tile keys, headers, TMA of neighbour tiles and own tiles into the ring
        unsigned e = 0;
        for (unsigned t = 0;; ++t) {
            uint64_t key;
            for (;;) {                                             // next valid tile of this shard
                unsigned long long c = 0;
                if (lane == 0) c = atomicAdd(tile_ctr, 1ULL);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (P.order != nullptr) {
                    key = c < (unsigned long long)P.norder ? (uint64_t)P.order[c] : P.key_hi;
                    break;
                }
                key = P.key_lo + c;
                if (key >= P.key_hi) break;
                const uint64_t Pb = __brevll(~key) >> (64 - P.A);
                const int js = P.k - __popcll(Pb);
                if (js >= 0 && js <= SD_BLK_B) break;
            }
            if (!sd_blkr_wait(&S.empty[e & (NB - 1)], ((e / NB) & 1u) ^ 1u, tile_ctr, diag, 1u, e, t)) return;   // first entry of the tile: see the protocol above
            SdBlkrHdr &H = S.hdr[t & (NB - 1)];
            if (key >= P.key_hi) {
                if (lane == 0) H.valid = -1;
                __syncwarp();
                if (lane == 0) sd_mbar_arrive(&S.full[e & (NB - 1)]);
                break;
            }
            sd_blk_make_hdr<1, SdBlkrHdr>(P, S.W, key, H, psi, qfar, lane);
            __syncwarp();
            const int ntot = nostream ? 0 : H.ntot;
            const int nring = sd_blkr_nring(H, ndirect), nnb = H.nnb;
            const double *own_src = psi.base[P.shards.rank] + H.base;
            bool first = true;
            for (int n = 0; n <= ntot; ++n) {
                if (n >= nring && n < nnb && n < ntot) continue;    // read directly by the consumers
                const unsigned slot = e & (NB - 1);
                if (!first && !sd_blkr_wait(&S.empty[slot], ((e / NB) & 1u) ^ 1u, tile_ctr, diag, 2u, e, t)) return;
                first = false;
                char *dst = (char *)(S.ring + (size_t)slot * P.cap);
                // the lane's copies of this entry (sd_blkr_copy); the transaction count is their byte total
                uint32_t tot = 0;
                {
                    const char *src; uint32_t off, len;
                    for (unsigned i = 0; sd_blkr_copy(S.js, H, own_src, P.dbg, n, ntot, lane, i, &src, &off, &len); ++i) tot += len;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                if (lane == 0) {
                    if (tot) sd_mbar_expect_tx(&S.full[slot], tot);
                    else sd_mbar_arrive(&S.full[slot]);              // nothing to read: an empty entry
                }
                __syncwarp();
                const bool evict_first = (P.dbg & 32) && n < H.nfar;
                const char *src; uint32_t off, len;
                for (unsigned i = 0; sd_blkr_copy(S.js, H, own_src, P.dbg, n, ntot, lane, i, &src, &off, &len); ++i) {
                    if (evict_first) sd_bulk_g2s_evict_first(dst + off, src + off, len, &S.full[slot]);
                    else sd_bulk_g2s(dst + off, src + off, len, &S.full[slot]);
                }
                ++e;
            }
        }
    } else {
        // ================= consumer warps
        SdBlkCtx X;
        X.P = &P; X.js = S.js; X.dmid = S.dmid; X.dtail = S.dtail; X.Jhop = S.Jhop;
        X.qx = P.Jz[P.A + SD_BLK_M - 1] * 0.25;
        X.pstart_local = P.shards.pstart[P.shards.rank];
        X.out_local = out_local;
        X.epi = &epi;
        const int slotmask = PLAIN ? 0 : sd_epi_slotmask(epi.red);
        unsigned e = 0;
        for (unsigned t = 0;; ++t) {
            if (!sd_blkr_wait(&S.full[e & (NB - 1)], (e / NB) & 1u, tile_ctr, diag, 3u, e, t)) return;
            SdBlkrHdr &H = S.hdr[t & (NB - 1)];
            if (H.valid < 0) break;
            SdBlkrLane Ln;
            sd_blkr_begin(Ln, S.js[H.js], S.rw[H.js * SD_BLK_CWARPS + warp], lane);
            const int ntot = nostream ? 0 : H.ntot;
            const int nring = sd_blkr_nring(H, ndirect), nnb = H.nnb;
            for (int n = nring; n < nnb && n < ntot; ++n) sd_blkr_stream_direct(Ln, H, n);   // while the ring entries fly
            bool first = true;
            for (int n = 0; n < ntot; ++n) {
                if (n >= nring && n < nnb) continue;
                const unsigned slot = e & (NB - 1);
                if (!first && !sd_blkr_wait(&S.full[slot], (e / NB) & 1u, tile_ctr, diag, 4u, e, t)) return;
                first = false;
                sd_blkr_stream(Ln, S.js, H, S.ring + (size_t)slot * P.cap, n);
                __syncwarp();
                if (lane == 0) sd_mbar_arrive(&S.empty[slot]);
                ++e;
            }
            const unsigned slot = e & (NB - 1);
            if (!first && !sd_blkr_wait(&S.full[slot], (e / NB) & 1u, tile_ctr, diag, 5u, e, t)) return;
            double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
            sd_blkr_own<PLAIN>(Ln, X, H, S.ring + (size_t)slot * P.cap, red);
            if (!PLAIN && slotmask) {
#pragma unroll
                for (int s = 0; s < SD_NSLOT; ++s) {
                    if (!((slotmask >> s) & 1)) continue;
                    double w = red[s];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
                    if (lane == 0) H.usum[s][warp] = w;
                }
                if (lane == 0) {
                    __threadfence_block();
                    const unsigned done = atomicAdd(&H.done_units, 1u);
                    if (done + 1 == SD_BLK_CWARPS) {               // last warp of the tile: sum in warp order
                        __threadfence_block();
                        for (int s = 0; s < SD_NSLOT; ++s) {
                            if (!((slotmask >> s) & 1)) continue;
                            double tsum = 0.0;
                            for (unsigned j = 0; j < SD_BLK_CWARPS; ++j) tsum += ((volatile double *)H.usum[s])[j];
                            epi.partials[(size_t)s * epi.nparts + H.tile_index] = tsum;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) sd_mbar_arrive(&S.empty[slot]);      // after the last use of the header (protocol above)
            ++e;
        }
    }
}
#endif  // __CUDACC__
