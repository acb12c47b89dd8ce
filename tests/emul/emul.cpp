// emul.cpp -- CPU emulation of the tiled CUDA kernel body (sd_tile.h).
//
// TEST INFRASTRUCTURE ONLY.  The product (libspindyn_cuda.so) never loads this;
// it exists so `pytest -m "not gpu"` can check the kernel's index algebra
// (tile bases, neighbour-tile shifts, class-major permutation, tail/mid/crossing
// hops, epilogue, shard ownership) against the oracle in a container without a
// GPU.  It runs the SAME __host__ __device__ phase functions the kernel runs,
// sequentially over the threads of one CTA, one CTA (tile) at a time.
#include <cstdlib>
#include <cstring>
#include <array>
#include <vector>
#include "../../spindynamics.jl_b200/csrc/sd_tile_host.h"

template <int NC, int T>
static void run_tiles(const SdTileHost &th, SdTileParams P, const SdVecView &psi, double *out_vbase,
                      const SdEpi &epi, unsigned nthreads, double *red_total) {
    const size_t bytes = sd_tile_smem_bytes(NC, th.cap_max, P.M);
    std::vector<unsigned char> smem(bytes + 16);
    void *sm = (void *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    for (uint64_t key = P.key_lo; key < P.key_hi; ++key) {
        std::memset(sm, 0xA5, bytes);              // poison: catches reads of unwritten smem
        SdTileView<NC> v = sd_tile_carve<NC>(sm, th.cap_max);
        std::vector<SdItem> regs(nthreads);        // per-thread register held across phases
        SdTileScratch scratch;
        for (unsigned tid = 0; tid < nthreads; ++tid) sd_tile_phase0a<NC>(P, key, v, tid, nthreads, regs[tid], scratch);
        for (unsigned tid = 0; tid < nthreads; ++tid) sd_tile_phase0b<NC>(P, key, v, psi, tid, nthreads, scratch);
        if (!v.hdr->valid) continue;
        for (unsigned tid = 0; tid < nthreads; ++tid) sd_tile_phase1<NC>(P, v, psi, tid, nthreads);
        for (unsigned tid = 0; tid < nthreads; ++tid)
            sd_tile_phase2<NC, T>(P, v, tid, nthreads, regs[tid]);
        const bool plain = epi.mode == SD_EPI_PLAIN && epi.red == 0 && !epi.acc && epi.hscale == 1.0;
        for (unsigned tid = 0; tid < nthreads; ++tid) {
            double red[SD_NSLOT] = {0, 0, 0, 0};
            if (plain) sd_tile_phase3<NC, true>(P, v, out_vbase, epi, tid, nthreads, red);
            else sd_tile_phase3<NC, false>(P, v, out_vbase, epi, tid, nthreads, red);
            for (int s = 0; s < SD_NSLOT; ++s) red_total[s] += red[s];
        }
    }
}

extern "C" {

// Runs rank `rank` of `world` over the full-length host vectors psi/out (and
// the optional epilogue vectors, also full length).  psi is split into
// per-shard copies so that the owner lookup really selects different buffers.
// Returns 0, or -1 if the tiled split is not representable.
int emul_tile_apply(int L, int k, int B, int T, const double *Jhop, const double *Jz, const double *h,
                    int NC, const double *psi, double *out, unsigned nthreads, int world, int rank,
                    int mode, int redmask, double hscale, double a, double b, const double *vprev,
                    const double *phi, double *acc, double ck_re, double ck_im, double *red_out,
                    uint64_t *bounds_out, uint64_t far_elems) {
    SdTileHost th;
    if (!sd_tile_build(L, k, B, T, Jhop, Jz, h, th)) return -1;
    SdTileParams P = th.P;
    P.binom = th.binom.data();
    P.perm = th.perm.data();
    P.items = th.items.data();
    P.pf_dist = 3;                                 // exercises the prefetch bookkeeping (no-op on the host)
    P.binomM = th.binomM.data();
    P.qfar = sd_tile_qfar(L, P.A, th.binom.data(), far_elems, 8 * NC);
    uint64_t bounds[SD_MAX_WORLD + 1], keys[SD_MAX_WORLD + 1];
    sd_tile_shard_bounds(th, world, bounds, keys);
    P.shards.world = world; P.shards.rank = rank;
    for (int g = 0; g <= world; ++g) P.shards.start[g] = bounds[g];
    for (int g = world + 1; g <= SD_MAX_WORLD; ++g) P.shards.start[g] = bounds[world];
    if (bounds_out) for (int g = 0; g <= world; ++g) bounds_out[g] = bounds[g];
    P.key_lo = keys[rank]; P.key_hi = keys[rank + 1];
    std::vector<std::vector<double>> shard(world);
    SdVecView view;
    for (int g = 0; g < SD_MAX_WORLD; ++g) view.base[g] = nullptr;
    for (int g = 0; g < world; ++g) {
        const uint64_t n = bounds[g + 1] - bounds[g];
        shard[g].assign(psi + bounds[g] * NC, psi + (bounds[g] + n) * NC);
        view.base[g] = shard[g].data() - (int64_t)bounds[g] * NC;
    }
    const uint64_t ls = bounds[rank];
    SdEpi epi;
    std::memset(&epi, 0, sizeof(epi));
    epi.mode = mode; epi.red = redmask; epi.hscale = hscale; epi.a = a; epi.b = b;
    epi.ck_re = ck_re; epi.ck_im = ck_im;
    epi.vprev = vprev ? vprev + ls * NC : nullptr;
    epi.phi = phi ? phi + ls * NC : nullptr;
    epi.acc = acc ? acc + ls * NC : nullptr;
    double red[SD_NSLOT] = {0, 0, 0, 0};
    if (NC == 1 && T == 5) run_tiles<1, 5>(th, P, view, out, epi, nthreads, red);
    else if (NC == 2 && T == 5) run_tiles<2, 5>(th, P, view, out, epi, nthreads, red);
    else if (NC == 1 && T == 4) run_tiles<1, 4>(th, P, view, out, epi, nthreads, red);
    else if (NC == 2 && T == 4) run_tiles<2, 4>(th, P, view, out, epi, nthreads, red);
    else if (NC == 1 && T == 6) run_tiles<1, 6>(th, P, view, out, epi, nthreads, red);
    else if (NC == 1 && T == 3) run_tiles<1, 3>(th, P, view, out, epi, nthreads, red);
    else return -2;
    if (red_out) for (int s = 0; s < SD_NSLOT; ++s) red_out[s] = red[s];
    return 0;
}

}  // extern "C"
