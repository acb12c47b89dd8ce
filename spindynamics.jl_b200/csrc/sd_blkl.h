// sd_blkl.h -- the lean block-layout apply kernel (the default for block-layout models).
//
// Same vector layout, tile keys, tile header and tables as sd_blk.h (which documents them); what differs is the
// consumer side.  Round 1's kernel was issue bound: 233 thread instructions per state at 25 % occupancy (512 threads x
// 128 registers), profiles/round1_b_apply_full.txt.  This one is written for instruction count and occupancy:
//   * the item body is fully specialised on the class (JT, chunk): slot counts, half slots, tail hops and the
//     mid|tail crossing are compile-time, nothing is selected per load
//   * per-slot tile offsets o[s] are computed once per item; a neighbour-tile load is one IMAD.WIDE + LDG.128
//   * stream entries are walked with warp-uniform control flow (no zero fill, no per-entry selects), two entries deep;
//     the prefix|mid crossing partner is loaded first and consumed under the first stream load
//   * every table, the tile headers and the per-launch context live in ONE STATIC shared-memory block (SD_SH), so
//     their addresses are immediates and no pointer to them occupies a register between items; only the tile buffers
//     are dynamic shared memory.  That is what brings the kernel from 128 to <= 96 registers per thread and lets
//     SD_BLKL_THREADS exceed 512.
// The body compiles for the host too (tests/emul runs it lane by lane against the oracle): there SD_SH is a plain
// static object.
// Tried and dropped (profiles/round2_s_x4_*.txt): a separate body for the f64 classes with ONE tail configuration in which a
// lane owns four mid configurations (32-byte stream loads; 28 instead of 32 work items per tile).  Correct, but 5.94 ms
// against 5.52: four item-table reads and four serial mid-hop loops per lane cost more than the saved load rounds, and
// the extra code hurts a kernel that is instruction-cache sensitive.
#pragma once
#include <type_traits>
#include "sd_blk.h"

#define SD_BLKL_NBUF 3           // tile buffers (f64); c128 uses 2

struct SdBlkShared {
    uint64_t full[SD_BLKL_NBUF], empty[SD_BLKL_NBUF];     // mbarriers
    SdBlkHdr hdr[SD_BLKL_NBUF];
    SdBlkJs js[SD_BLK_B + 1];
    uint16_t units[(SD_BLK_B + 1) * SD_BLK_MAXUNITS];
    double dmid[1 << SD_BLK_M];  // diagonal of the mid sites, in item order: lanes = consecutive u read consecutive entries
    double hs;                   // hscale of the epilogue, resolved once per CTA (SdEpi::hscale_dev)
};
// Everything that is the same for every lane of a warp and fixed for the launch (mid / tail hop coefficients, tail
// diagonal, qx, the epilogue's mode / coefficients / pointers, the out pointer) is NOT in shared memory: it is read from
// the __grid_constant__ kernel parameters P / epi, i.e. from the constant bank.  A warp-uniform LDS costs a full 128-byte wavefront of the LSU data pipe, which is this kernel's scarcest
// resource (profiles/round2_b_apply_full.txt: 71 % busy, 12 % of it uniform table reads).
#if defined(__CUDACC__)
__shared__ SdBlkShared sd_blkl_sh;
#define SD_SH sd_blkl_sh
#define SD_BLKL_FN __device__ __forceinline__
#else
static SdBlkShared sd_blkl_sh;
#define SD_SH sd_blkl_sh
#define SD_BLKL_FN inline
#endif
// compile-time loop: f(std::integral_constant<int, 0>) ... f(std::integral_constant<int, N - 1>)
template <int N, int I = 0, class F>
SD_BLKL_FN void sd_static_for(F &&f) {
    if constexpr (I < N) { f(std::integral_constant<int, I>{}); sd_static_for<N, I + 1>(f); }
}
// resolves the epilogue's hscale (device: one thread, before the CTA barrier; host: the emulation)
SD_BLKL_FN void sd_blkl_ctx_init(const SdEpi &epi) {
    SD_SH.hs = epi.hscale_dev ? epi.hscale / sqrt(*epi.hscale_dev) : epi.hscale;
}

// EK: epilogue kind, compile-time so that each instantiation only carries the code it runs (the kernel is instruction-
// cache sensitive): 0 plain out = H psi; 1 Lanczos: out = hs * H psi with the fused <psi, out> (every Lanczos flavour);
// 2 generic (rescaled / Chebyshev step, psi_t accumulation, phi dot, norm: sd_epilogue_hs).
template <int NC, int JT, int S0, int EK, bool WRAP = false>
SD_BLKL_FN void sd_blkl_item(const SdBlkParams &P, const SdEpi &E, double *out_local, const SdBlkHdr &H, const double *tb, uint32_t u,
                             double (&red)[SD_NSLOT]) {
    constexpr bool PLAIN = EK == 0;
    constexpr int T = SD_BLK_T, M = SD_BLK_M;
    constexpr int NT = sd_cbinom(T, JT);
    constexpr int NO = NC == 1 ? (NT + 1) / 2 : NT;                  // slots of the whole block
    constexpr int EC = NC == 1 ? NO : (NT > 5 ? 5 : NT);             // slots of the chunk
    constexpr bool HALF = NC == 1 && (NT & 1) != 0;
    constexpr int NE = NC == 1 ? NT : EC;                            // tail configurations of the chunk
    constexpr int E0 = NC == 1 ? 0 : S0;
    static_assert(SD_BLK_T == 5 && (NC == 2 || S0 == 0) && (S0 == 0 || (S0 == 5 && NT == 10)), "chunking is written for T = 5");
    const SdBlkJs &I = SD_SH.js[H.js];
    const SdBlkCls cls = I.cls[JT];
    if (u >= cls.nblk) return;
    const uint32_t ss = 2u * cls.pitch;                              // doubles between slots (both dtypes)
    const uint32_t off0 = cls.cb * NC + 2u * u;                      // doubles, slot 0 of the block
    uint32_t o[EC];                                                  // doubles, slot s of the chunk (half slot: its plain row)
#pragma unroll
    for (int s = 0; s < EC; ++s) o[s] = off0 + (uint32_t)(S0 + s) * ss - ((HALF && s == EC - 1) ? u : 0u);
    double2 acc[EC], t0[EC], t1[EC];
#pragma unroll
    for (int s = 0; s < EC; ++s) acc[s] = t0[s] = t1[s] = make_double2(0.0, 0.0);   // t0/t1: conditionally loaded below; left undefined they end up on the stack
#define SD_LEAN_LOAD(t_, p_)                                                                  \
    do {                                                                                      \
        const double *q_ = (p_);                                                              \
        _Pragma("unroll") for (int s = 0; s < EC; ++s)                                        \
            t_[s] = (HALF && s == EC - 1) ? sd_blk_ldg_half(q_ + o[s]) : sd_blk_ldg(q_ + o[s]); \
    } while (0)
#define SD_LEAN_FMA(t_, J_)                                                                   \
    do {                                                                                      \
        const double j_ = (J_);                                                               \
        _Pragma("unroll") for (int s = 0; s < EC; ++s) { acc[s].x += j_ * t_[s].x; acc[s].y += j_ * t_[s].y; } \
    } while (0)
    // ---- prefix|mid crossing bond: partner tile with js +- 1, same class, uniform block shift; only the lanes whose
    // first mid bit differs from the last prefix bit
    const bool c0 = u < cls.n1;                                      // first mid bit (blocks with it set come first)
    const int nnb = H.nnb;
    bool xl = false;
    if (H.xptr != nullptr) {
        xl = c0 != (bool)H.bP;
        if (xl) {
            const SdBlkCls cx = SD_SH.js[H.jsx].cls[JT];
            const uint32_t xu = H.bP ? u - cls.n1 : cx.n1 + u;
            const uint32_t xs = 2u * cx.pitch;
            const double *xp = H.xptr + (cx.cb * NC + 2u * xu);
#pragma unroll
            for (int s = 0; s < EC; ++s)
                t1[s] = (HALF && s == EC - 1) ? sd_blk_ldg_half(xp + (uint32_t)(S0 + s) * xs - xu) : sd_blk_ldg(xp + (uint32_t)(S0 + s) * xs);
        }
    }
    // ---- prefix-internal bonds: whole neighbour tiles in the same element order
    // (Three / four entries in flight per lane on 512 threads were measured slower, 7.3 / 8.9 ms against 5.8 at two entries
    // on 640 threads: profiles/round2_n_ab.txt.)
    SdBlkEnt e0 = H.nb[0], e1;                                       // {tile base, J}: one LDS.128 per entry
    if (nnb > 0) SD_LEAN_LOAD(t0, e0.p);
    if (xl) SD_LEAN_FMA(t1, H.Jx);
    int n = 0;
#pragma unroll 1
    while (n + 1 < nnb) {
        e1 = H.nb[n + 1];
        SD_LEAN_LOAD(t1, e1.p);
        SD_LEAN_FMA(t0, e0.J);
        if (n + 2 < nnb) { e0 = H.nb[n + 2]; SD_LEAN_LOAD(t0, e0.p); }
        SD_LEAN_FMA(t1, e1.J);
        n += 2;
    }
    if (n < nnb) SD_LEAN_FMA(t0, e0.J);
#undef SD_LEAN_LOAD
#undef SD_LEAN_FMA
    // ---- periodic chain (WRAP variant): the bond between tail site T-1 and prefix site 0.  +-Jz/4 on the diagonal of every
    // element; the hop acts on the tail configurations whose last bit differs from prefix bit 0 (tile-uniform) and reads one
    // element of the wrap partner tile: class jt -+ 1, the configuration with the last bit flipped, the same u.  The loads
    // come after the streams, when the stream registers are free again.
    if constexpr (WRAP) {
        const int b0 = H.b0;
        const double dw = b0 ? P.wrapJz4 : -P.wrapJz4;
        const double *wp = H.wptr;
        const SdBlkJs &Iw = SD_SH.js[H.jsw];
        sd_static_for<NE>([&](auto tc) {
            constexpr int t = decltype(tc)::value;
            constexpr unsigned cfgt = sd_tail_cfg(T, JT, E0 + t);
            constexpr int tbit = (int)((cfgt >> (T - 1)) & 1u);
            constexpr int jt2 = tbit ? JT - 1 : JT + 1;
            double wr = 0.0, wi = 0.0;
            if constexpr (jt2 >= 0 && jt2 <= T) {
                if (wp != nullptr && tbit != b0) {
                    constexpr int e2 = sd_tail_rank(T, jt2, cfgt ^ (1u << (T - 1)));
                    constexpr int NT2 = sd_cbinom(T, jt2);
                    const SdBlkCls c2 = Iw.cls[jt2];
                    if (NC == 1) {
                        const uint32_t p2 = ((NT2 & 1) && e2 == NT2 - 1) ? c2.cb + (uint32_t)e2 * c2.pitch + u
                                                                         : c2.cb + (uint32_t)(e2 >> 1) * 2u * c2.pitch + 2u * u + (uint32_t)(e2 & 1);
                        wr = P.wrapJ * sd_blk_ldg_half(wp + p2).x;
                    } else {
                        const double2 v = sd_blk_ldg(wp + 2u * (c2.cb + (uint32_t)e2 * c2.pitch + u));
                        wr = P.wrapJ * v.x; wi = P.wrapJ * v.y;
                    }
                }
            }
            const double dsg = tbit ? dw : -dw;
            if (NC == 1) {
                constexpr int sl = (E0 + t) >> 1;
                const double ownv = (HALF && sl == EC - 1) ? tb[o[sl]] : tb[o[sl] + (uint32_t)((E0 + t) & 1)];
                SD_BLK_EL(acc, t, 1) += wr + dsg * ownv;
            } else {
                const double2 ownv = *(const double2 *)(tb + o[t]);
                acc[t].x += wr + dsg * ownv.x;
                acc[t].y += wi + dsg * ownv.y;
            }
        });
    }
    // ---- own block: diagonal + tail-internal hops (registers, compile-time permutation)
    const uint4 it = sd_blk_ld_item(P.items + cls.item_off + u);     // x,y,z = nb[12]; w = c | u2x << 16
    const unsigned cmid = it.w & ((1u << M) - 1u);
    const bool clast = (cmid >> (M - 1)) & 1u;
    sd_blk_tail<NC, JT, E0, NE, EC>(acc, tb + off0, ss, u, P.Jtail, P.dtail, H.dP[c0 ? 1 : 0] + SD_SH.dmid[cls.item_off + u],
                                    clast ? P.qx : -P.qx);
    // ---- mid-internal hops: the whole block moves to block nb[pm] of the same class
    {
        const double *cbp = tb + cls.cb * NC + (uint32_t)S0 * ss;
        // Plain kernel: fully unrolled (byte extract and J compile-time selected; 41 -> 54 KB of code, 5 % faster).
        // Fused kernels: a loop -- nine unrolled copies in each of the six class bodies are 13 - 20 KB, and with their
        // epilogues those kernels were instruction-fetch bound ("no instruction" was the top stall at 99 KB).
        uint64_t lo = (uint64_t)it.x | ((uint64_t)it.y << 32);
        uint32_t hi = it.z;
        constexpr int MU = PLAIN ? (M - 1) : 1;
#pragma unroll MU
        for (int pm = 0; pm + 1 < M; ++pm) {
            unsigned nbu;
            if (PLAIN) {
                nbu = ((pm < 4 ? it.x : (pm < 8 ? it.y : it.z)) >> (8 * (pm & 3))) & 0xFFu;
            } else {
                nbu = (unsigned)(lo & 0xFFu);
                lo = (lo >> 8) | ((uint64_t)hi << 56);
                hi >>= 8;
            }
            if (nbu != 0xFFu) {
                const double J = P.Jmid[pm];
                const double *sp = cbp + 2u * nbu;
#pragma unroll
                for (int s = 0; s < EC; ++s) {
                    if (HALF && s == EC - 1) {
                        acc[s].x += J * *(sp + (uint32_t)s * ss - nbu);
                    } else {
                        const double2 t = *(const double2 *)(sp + (uint32_t)s * ss);
                        acc[s].x += J * t.x;
                        acc[s].y += J * t.y;
                    }
                }
            }
        }
    }
    // ---- mid|tail crossing bond, per tail configuration.  Tail configurations with bit 0 set come first in a
    // class (n1 of them).  Last mid bit set & tail bit 0 clear -> class JT+1, configuration e - n1; last mid bit
    // clear & tail bit 0 set -> class JT-1, configuration C(T-1, JT-2) + e.  u2x: the block with the last mid bit flipped.
    {
        constexpr int n1 = sd_cbinom(T - 1, JT - 1);
        const double J = P.Jmid[M - 1];
        const uint32_t u2x = it.w >> 16;
#define SD_LEAN_CROSS(JT2_, ELO_, EHI_, SHIFT_)                                               \
    do {                                                                                      \
        constexpr int NT2 = sd_cbinom(T, (JT2_));                                             \
        const SdBlkCls c2 = I.cls[(JT2_)];                                                    \
        const double *sp = tb + c2.cb * NC + 2u * u2x;                                        \
        const uint32_t s2 = 2u * c2.pitch;                                                    \
        _Pragma("unroll") for (int e = (ELO_); e < (EHI_); ++e) {                             \
            const int e2 = e + (SHIFT_);                                                      \
            if (NC == 1) {                                                                    \
                const double t = ((NT2 & 1) && e2 == NT2 - 1) ? *(sp + (uint32_t)(e2 >> 1) * s2 - u2x)   \
                                                              : *(sp + (uint32_t)(e2 >> 1) * s2 + (e2 & 1)); \
                SD_BLK_EL(acc, e - E0, 1) += J * t;                                           \
            } else {                                                                          \
                const double2 t = *(const double2 *)(sp + (uint32_t)e2 * s2);                 \
                acc[e - E0].x += J * t.x;                                                     \
                acc[e - E0].y += J * t.y;                                                     \
            }                                                                                 \
        }                                                                                     \
    } while (0)
        if (clast) {
            if constexpr (JT + 1 <= T) {
                constexpr int elo = E0 > n1 ? E0 : n1, ehi = E0 + NE < NT ? E0 + NE : NT;
                SD_LEAN_CROSS(JT + 1, elo, ehi, -n1);
            }
        } else {
            if constexpr (JT >= 1) {
                constexpr int ehi = E0 + NE < n1 ? E0 + NE : n1;
                SD_LEAN_CROSS(JT - 1, E0, ehi, sd_cbinom(T - 1, JT - 2));
            }
        }
#undef SD_LEAN_CROSS
    }
    // ---- epilogue + store
    const uint64_t ld0 = (H.base - P.shards.pstart[P.shards.rank]) * NC;  // doubles from the start of the local shard
    double *ob = out_local + ld0;
    if (PLAIN) {
#pragma unroll
        for (int s = 0; s < EC; ++s) {
            if (HALF && s == EC - 1) sd_blk_stg_half(ob + o[s], acc[s].x);
            else sd_blk_stg(ob + o[s], acc[s]);
        }
    } else if (EK == 1) {                                            // out = hs * H psi, red[0..1] += conj(psi) out
        const double hs = SD_SH.hs;
#pragma unroll
        for (int s = 0; s < EC; ++s) {
            if (HALF && s == EC - 1) {
                const double r = hs * acc[s].x;
                red[0] += tb[o[s]] * r;
                sd_blk_stg_half(ob + o[s], r);
                continue;
            }
            const double2 p = *(const double2 *)(tb + o[s]);
            const double2 r = make_double2(hs * acc[s].x, hs * acc[s].y);
            if (NC == 2) { red[0] += p.x * r.x + p.y * r.y; red[1] += p.x * r.y - p.y * r.x; }
            else red[0] += p.x * r.x + p.y * r.y;
            sd_blk_stg(ob + o[s], r);
        }
    } else {
        const double hs = SD_SH.hs;
#pragma unroll
        for (int s = 0; s < EC; ++s) {
            const uint64_t ld = ld0 + o[s];
            if (HALF && s == EC - 1) {
                SdVal<1> hh, pp;
                hh.c[0] = acc[s].x; pp.c[0] = tb[o[s]];
                const SdVal<1> r0 = sd_epilogue_hs<1>(E, hs, hh, pp, ld, red);
                sd_blk_stg_half(ob + o[s], r0.c[0]);
                continue;
            }
            const double2 p = *(const double2 *)(tb + o[s]);
            double2 r;
            if (NC == 2) {
                SdVal<2> hh, pp;
                hh.c[0] = acc[s].x; hh.c[1] = acc[s].y; pp.c[0] = p.x; pp.c[1] = p.y;
                const SdVal<2> rr = sd_epilogue_hs<2>(E, hs, hh, pp, ld / 2, red);
                r = make_double2(rr.c[0], rr.c[1]);
            } else {
                SdVal<1> hh, pp;
                hh.c[0] = acc[s].x; pp.c[0] = p.x;
                const SdVal<1> r0 = sd_epilogue_hs<1>(E, hs, hh, pp, ld, red);
                hh.c[0] = acc[s].y; pp.c[0] = p.y;
                const SdVal<1> r1 = sd_epilogue_hs<1>(E, hs, hh, pp, ld + 1, red);
                r = make_double2(r0.c[0], r1.c[0]);
            }
            sd_blk_stg(ob + o[s], r);                                 // evict-first like the plain path: out must not displace psi in L2
        }
    }
}
template <int NC, int EK, bool WRAP = false>
SD_BLKL_FN void sd_blkl_dispatch(const SdBlkParams &P, const SdEpi &E, double *out_local, const SdBlkHdr &H, const double *tb, unsigned code, uint32_t u,
                                double (&red)[SD_NSLOT]) {
    const int jt = (int)(code >> 12);
    const bool hi = ((code >> 8) & 0xFu) != 0;                       // c128, classes of 10: second chunk of five
    switch (jt) {
        case 0: sd_blkl_item<NC, 0, 0, EK, WRAP>(P, E, out_local, H, tb, u, red); break;
        case 1: sd_blkl_item<NC, 1, 0, EK, WRAP>(P, E, out_local, H, tb, u, red); break;
        case 2:
            if constexpr (NC == 2) { if (hi) { sd_blkl_item<NC, 2, 5, EK, WRAP>(P, E, out_local, H, tb, u, red); break; } }
            sd_blkl_item<NC, 2, 0, EK, WRAP>(P, E, out_local, H, tb, u, red); break;
        case 3:
            if constexpr (NC == 2) { if (hi) { sd_blkl_item<NC, 3, 5, EK, WRAP>(P, E, out_local, H, tb, u, red); break; } }
            sd_blkl_item<NC, 3, 0, EK, WRAP>(P, E, out_local, H, tb, u, red); break;
        case 4: sd_blkl_item<NC, 4, 0, EK, WRAP>(P, E, out_local, H, tb, u, red); break;
        default: sd_blkl_item<NC, 5, 0, EK, WRAP>(P, E, out_local, H, tb, u, red); break;
    }
}


#if defined(__CUDACC__)
// grid = one persistent CTA per SM of NTHR threads; the last warp is the producer (sd_blk_producer: tile keys from the global
// counter or the order table, headers, TMA of the own tiles), the others pull (tile, unit) items.
template <int NC, int EK, int NTHR, bool WRAP = false>
__global__ void __launch_bounds__(NTHR, 1)
sd_blkl_apply_kernel(const __grid_constant__ SdBlkParams P, const __grid_constant__ SdVecView psi, double *out_local,
                     const __grid_constant__ SdEpi epi, int qfar, unsigned long long *tile_ctr) {
    extern __shared__ __align__(128) unsigned char sd_blk_smem[];    // [nbuf][cap * NC] doubles: the tile buffers
    constexpr bool PLAIN = EK == 0;
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const unsigned nbuf = (unsigned)P.nbuf;
    constexpr unsigned NCONS = NTHR / 32 - 1;                        // consumer warps; the last warp is the producer
    {
        const uint32_t *src = (const uint32_t *)P.js;
        uint32_t *dst = (uint32_t *)SD_SH.js;
        for (int i = (int)tid; i < (int)(sizeof(SdBlkJs) * (SD_BLK_B + 1) / 4); i += NTHR) dst[i] = src[i];
    }
    for (int i = (int)tid; i < (SD_BLK_B + 1) * SD_BLK_MAXUNITS; i += NTHR)
        SD_SH.units[i] = P.units[(NC - 1) * (SD_BLK_B + 1) * SD_BLK_MAXUNITS + i];
    for (int i = (int)tid; i < (1 << SD_BLK_M); i += NTHR) SD_SH.dmid[i] = P.dmid[i];
    if (tid == 0) {
        sd_blkl_ctx_init(epi);
        for (unsigned b = 0; b < nbuf; ++b) { sd_mbar_init(&SD_SH.full[b], 1); sd_mbar_init(&SD_SH.empty[b], NCONS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned tile_doubles = P.cap * NC;
    if (warp == NCONS) {
        SdBlkSmem S;
        S.full = SD_SH.full; S.empty = SD_SH.empty; S.hdr = SD_SH.hdr; S.W = P.W; S.js = SD_SH.js;
        S.tiles = (double *)sd_blk_smem;
        sd_blk_producer<NC, WRAP>(P, S, psi, qfar, tile_ctr, lane, epi, PLAIN ? 0 : sd_epi_slotmask(epi.red));
    } else {
        const int slotmask = PLAIN ? 0 : sd_epi_slotmask(epi.red);
        unsigned b = 0, phase = 0;
        for (;;) {
            sd_mbar_wait(&SD_SH.full[b], phase);
            SdBlkHdr &H = SD_SH.hdr[b];
            if (H.valid < 0) break;
            const double *tb = (const double *)sd_blk_smem + (size_t)b * tile_doubles;
            for (;;) {
                unsigned un = 0;
                if (lane == 0) un = atomicAdd(&H.next_unit, 1u);
                un = __shfl_sync(0xffffffffu, un, 0);
                const unsigned nunits = SD_SH.js[H.js].nunits[NC - 1];
                if (un >= nunits) break;
                const unsigned code = SD_SH.units[H.js * SD_BLK_MAXUNITS + un];
                const uint32_t u = (code & 0xFFu) * 32u + lane;      // the lane's mid configuration
                double red[SD_NSLOT] = {0.0, 0.0, 0.0, 0.0};
                sd_blkl_dispatch<NC, EK, WRAP>(P, epi, out_local, H, tb, code, u, red);
                if (!PLAIN && slotmask) sd_blk_item_reduce(H, slotmask, un, red, lane);
            }
            __syncwarp();
            if (lane == 0) sd_mbar_arrive(&SD_SH.empty[b]);
            if (++b == nbuf) { b = 0; phase ^= 1u; }
        }
    }
}
#endif  // __CUDACC__
