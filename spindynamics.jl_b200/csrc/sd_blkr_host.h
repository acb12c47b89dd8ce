// sd_blkr_host.h -- host tables of the ring variant of the block kernel (sd_blkr.h): the static packing of a
// tile's f64 work items into the register groups of the 15 consumer warps, one packing per suffix popcount.
// Pure C++ (no CUDA).
#pragma once
#include <algorithm>
#include <vector>
#include "sd_blk_host.h"
#include "sd_blkr.h"

// rw[(js * SD_BLK_CWARPS) + warp].  Every f64 item of a tile (class jt, unit of 32 mid configurations) appears in
// exactly one register group of exactly one warp:
//   group A (5 slots): one item of any class;
//   group B (3 slots): one item of a 3-slot class (jt = 1, 4) in b[0], or up to three items of 1-slot classes (jt = 0, 5).
// Items are placed heaviest first on the warp with the fewest slots in use, so the warps of a tile finish a ring
// entry at about the same time (at js = 7, 8 of an Sz = 0 chain: 114 slots on 15 warps, 7 or 8 each).
// Returns false if some suffix popcount cannot be packed (the model then stays on sd_blk_apply_kernel).
static inline bool sd_blkr_build(const SdBlkHost &o, std::vector<SdBlkrWarp> &rw) {
    constexpr int B = SD_BLK_B, T = SD_BLK_T, NW = SD_BLK_CWARPS;
    rw.assign((size_t)(B + 1) * NW, SdBlkrWarp{SD_BLKR_NONE, {SD_BLKR_NONE, SD_BLKR_NONE, SD_BLKR_NONE}});
    for (int js = 0; js <= B; ++js) {
        const SdBlkJs &I = o.js[js];
        struct It { int ec; uint16_t code; };
        std::vector<It> items;
        for (int jt = 0; jt <= T; ++jt) {
            const SdBlkCls &c = I.cls[jt];
            if (c.nblk == 0) continue;
            const uint32_t nu = (c.nblk + 31u) / 32u;
            if (nu > 0xFFFu) return false;
            for (uint32_t j = 0; j < nu; ++j) items.push_back({sd_blkr_ec(jt), (uint16_t)((jt << 12) | j)});
        }
        std::stable_sort(items.begin(), items.end(), [](const It &a, const It &b) { return a.ec > b.ec; });
        int load[NW] = {0};
        int nb1[NW] = {0};               // 1-slot items in group B
        bool b3[NW] = {false};           // group B holds a 3-slot item
        SdBlkrWarp *W = &rw[(size_t)js * NW];
        for (const It &it : items) {
            int best = -1, where = -1;   // where: 0 = A, 1 = B
            for (int w = 0; w < NW; ++w) {
                int pos = -1;
                // prefer the group that wastes the fewest registers: small items go to B when it has room
                const bool b_ok = (it.ec == 3 && !b3[w] && nb1[w] == 0) || (it.ec == 1 && !b3[w] && nb1[w] < 3);
                const bool a_ok = W[w].a == SD_BLKR_NONE;
                if (it.ec == 5) { if (a_ok) pos = 0; }
                else if (b_ok) pos = 1;
                else if (a_ok) pos = 0;
                if (pos < 0) continue;
                if (best < 0 || load[w] < load[best]) { best = w; where = pos; }
            }
            if (best < 0) return false;
            if (where == 0) W[best].a = it.code;
            else if (it.ec == 3) { W[best].b[0] = it.code; b3[best] = true; }
            else W[best].b[nb1[best]++] = it.code;
            load[best] += it.ec;
        }
    }
    return true;
}
