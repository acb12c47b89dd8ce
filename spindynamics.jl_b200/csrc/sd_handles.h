// sd_handles.h -- what the translation units of libspindyn_cuda share: error plumbing, the NCCL function table, the
// opaque handles of include/spindyn.h (sd_ctx, sd_model, sd_vec, sd_vecset) and the helpers defined in sd_api.cu that
// sd_batch.cu (fused reorthogonalisation, q-batched recurrences) calls.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/spindyn.h"
#include "sd_common.h"
#include "sd_tile.h"
#include "sd_tile_host.h"
#include "sd_blk.h"
#include "sd_blk_host.h"

#define SD_VERSION 100

// ----------------------------------------------------------------- errors
int sd_fail(int code, const char *fmt, ...);                      // sd_api.cu: formats into the thread-local error string
#define SD_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return sd_fail(e_ == cudaErrorMemoryAllocation ? SD_ERR_NOMEM : SD_ERR_CUDA,    \
                           "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define SD_TRY(call)              \
    do {                          \
        int r_ = (call);          \
        if (r_ != SD_OK) return r_; \
    } while (0)
#define SD_LOCK(ctxptr) std::lock_guard<std::recursive_mutex> sd_ctx_lock_((ctxptr)->mu)
#define SD_ARG(cond, ...)                                  \
    do {                                                   \
        if (!(cond)) return sd_fail(SD_ERR_ARG, __VA_ARGS__); \
    } while (0)

// ----------------------------------------------------------------- NCCL (dlopen, only when world > 1)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclUint8_ = 1, ncclFloat64_ = 8 };   // ncclDataType_t values (nccl.h)
enum { ncclSum_ = 0 };
struct SdNccl {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
extern SdNccl g_nccl;
int sd_nccl_load();
#define SD_NCCL(call)                                                                      \
    do {                                                                                   \
        int e_ = (call);                                                                   \
        if (e_ != 0) return sd_fail(SD_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(e_)); \
    } while (0)

// ----------------------------------------------------------------- handles
#define SD_NSCAL 4096
#define SD_HIST 4096                  // d_scal[SD_HIST + 8 j ..]: reductions of Lanczos step j (kept on the device, fetched in blocks)
#define SD_HIST_MAX 4096             // steps
struct sd_ctx {
    // Threading contract: calls on one context serialise.  Every entry point that touches the context's
    // stream or scratch state takes this lock (recursive: entry points call each other), so the reference's
    // Threads.@threads q-loops stay correct when they share a context -- they become sequential device work; for
    // concurrency use one context per host thread (spindyn's q_threads).
    mutable std::recursive_mutex mu;
    int device = 0;
    int rank = 0, world = 1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    ncclComm_t comm = nullptr;
    uint64_t *d_binom = nullptr;
    double *d_scal = nullptr;       // device scalars [SD_NSCAL]
    double *h_scal = nullptr;       // pinned mirror
    double *d_partials = nullptr;
    size_t partials_cap = 0;        // doubles
    unsigned char *d_ipc = nullptr; // [(world + 1) * 128] exchange buffer of sd_exchange
    std::vector<unsigned char> h_ipc;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t launches = 0;
    std::vector<uint64_t> binom;
    // ---- copy engine of sd_vec_upload_async / sd_vec_download_async on block-layout vectors (sd_api.cu): pinned host
    // <-> rank-ordered staging on two copy streams, chunk by chunk, the layout permute of chunk i on the compute stream
    // while chunk i + 1 is on the wire; an upload and a download run concurrently (both PCIe directions).
    cudaStream_t h2d = nullptr, d2h = nullptr;
    std::vector<cudaEvent_t> ev_copy;         // [2 * SD_COPY_CHUNKS + 4]: per-chunk events up / down, fork, joins, d2h done
    bool copy_pending = false;                // work may still be running on h2d / d2h that the compute stream has not waited for
    bool d2h_pending = false;                 // the download staging buffer is still being read by the d2h stream
    bool up_event = false;                    // ev_copy[fork] marks the end of the previous upload's permutes (the upload staging buffer is free after it)
    bool stage_touched[2] = {false, false};   // a compute-stream user (sd_scratch) has had the staging buffer since the last asynchronous copy
    const double **d_vtab = nullptr;          // sd_reorth_step: device table of the basis vectors' pointers (+ host mirror)
    std::vector<const double *> h_vtab;
    int vtab_count = 0;
    double *d_rth_partials = nullptr;
    unsigned rth_grid_max = 0;
    unsigned long long *d_tilectr = nullptr;  // tile counter of the block kernel's dynamic scheduler
    void *scratch[2] = {nullptr, nullptr};   // rank-ordered staging of block-layout vectors (upload/download/szq)
    size_t scratch_cap[2] = {0, 0};
    // ---- cross-rank ordering of sharded vectors (world > 1).  Every rank makes the same API calls in the same order
    // (SPMD), vector ids are handed out by the collective sd_vec_alloc, so these sets evolve identically on all ranks and
    // the barriers they trigger pair up.  Both are emptied by every collective (all ranks' earlier kernels have finished).
    std::vector<uint64_t> dirty_ids;         // vectors written since the last collective: peers must not gather them yet
    std::vector<uint64_t> read_ids;          // vectors an apply gathered from since the last collective: peers may still read the local shard
    uint64_t next_vec_id = 1;
    // ---- deferred release of IPC-exported shards (world > 1): sd_vec_free is LOCAL (finalizers run at different times
    // on different ranks); a shard is cudaFree'd once every rank has announced the free of that vector id, which the
    // ranks tell each other inside the next collective sd_vec_alloc / sd_ctx_collect.
    struct Dead { uint64_t id; double *d; };
    std::vector<Dead> dead;                  // local shards waiting for the peers to unmap them
    std::vector<uint64_t> outbox;            // locally freed ids not yet announced
    std::vector<std::pair<uint64_t, int>> freed_count;   // id -> ranks that announced it
};

struct SdBlkDev {
    bool ok = false;
    SdBlkHost host;
    uint64_t *d_W = nullptr;
    SdBlkJs *d_js = nullptr;
    uint16_t *d_units = nullptr;
    SdBlkItem *d_items = nullptr;
    double *d_dmid = nullptr;
    uint64_t pstart[SD_MAX_WORLD + 1];
    int nbuf[2] = {0, 0};
    size_t smem[2] = {0, 0};
    int qfar[2] = {0, 0};
    int pfp = 0;                    // SD_BLK_PFP (experiment, SdBlkParams::pfp)
    int threads = 640;              // CTA size of sd_blkl_apply_kernel (SD_BLKL_THREADS = 512 | 640 | 768, read once at model creation)
    uint32_t *d_order = nullptr;    // breadth-first tile order of this rank's shard (vectors larger than the L2)
    uint32_t norder = 0;
};

struct SdTileDev {
    bool ok = false;
    SdTileHost host;
    uint32_t cap = 0;
    size_t smem = 0;
    void *d_perm = nullptr, *d_items = nullptr, *d_binomM = nullptr;
    uint64_t keys[SD_MAX_WORLD + 1];
};

#define SD_COPY_CHUNKS 8
struct sd_model {
    sd_ctx *ctx = nullptr;
    std::vector<uint64_t> cp_keys, cp_ranks;  // chunk boundaries of the copy engine: tile keys / local basis ranks, [chunks + 1]
    int L = 0, k = -1;
    uint64_t N = 0;
    std::vector<int> hop_a, hop_b, zz_a, zz_b;
    std::vector<double> hop_J, zz_J, field;
    int *d_hop_a = nullptr, *d_hop_b = nullptr, *d_zz_a = nullptr, *d_zz_b = nullptr;
    double *d_hop_J = nullptr, *d_zz_J = nullptr, *d_field = nullptr;
    uint64_t *d_linA = nullptr, *d_linB = nullptr;
    int lin_h = 0;
    int path = SD_PATH_GENERIC;
    bool tile_capable = false;
    int tile_T[2] = {5, 4};
    int tile_threads = 512;
    SdTileDev tile[2];              // [0]: F64, [1]: C128
    SdShardMap shards;
    SdBlkDev blk;                   // block-layout kernel (sd_blk.h)
    bool blk_layout = false;        // vectors of this model are stored in block layout
    bool has_wrap = false;          // periodic chain: nearest-neighbour bonds plus the wrap bond (sites L-1, 0)
    double wrap_hop = 0.0, wrap_zz = 0.0;   // its hop coefficient and Jz (the WRAP variant of the block kernel, sd_blkl.h)
    int live_vecs = 0;
    std::vector<sd_vec *> pool;     // idle work vectors of the recurrences (SdVecGuard)
    bool free_pending = false;      // sd_model_free was called while vectors were alive
};

struct sd_vec {
    sd_model *model = nullptr;
    int dtype = SD_F64, nc = 1;
    uint64_t local_n = 0;           // STORED elements of the local shard (block layout: padded)
    uint64_t logical_n = 0;         // basis states of the local shard
    int layout = 0;                 // 0: rank order, 1: block layout
    uint64_t id = 0;                // collective allocation number (same on every rank)
    double *d = nullptr;
    SdVecView view;
    void *peer[SD_MAX_WORLD];
    bool owned = true;
};

struct sd_vecset {
    std::vector<sd_vec *> v;
};

// ----------------------------------------------------------------- helpers defined in sd_api.cu
int sd_launch_check(sd_ctx *c, const char *what);
int sd_use(const sd_ctx *c);
int sd_partials_reserve(sd_ctx *c, size_t doubles);
int sd_fetch(sd_ctx *c, int slot, int n, double *out);
int sd_env_int(const char *name, int dflt);
int sd_scratch(sd_ctx *c, int which, size_t bytes, double **p);
void sd_scratch_release(sd_ctx *c);
int sd_blk_permute(const sd_vec *v, double *rank_local, int nc_rank, int dir, int seeded, uint64_t seed, double scale);

// ----------------------------------------------------------------- sd_batch.cu
int sd_reorth_begin(sd_ctx *c, int mm);
int sd_reorth_step(sd_ctx *c, sd_vec *w, sd_vec *const *V, int j, sd_vec *vnext, double tol, double orth_tol);
int sd_reorth_finish(sd_ctx *c, int mm, double *rec);
