// sd_api.cu -- the C ABI of libspindyn_cuda (include/spindyn.h): contexts,
// models, device vectors, the H.psi operator and the recurrences that call it.
// Host code only orchestrates; all vector work runs in the kernels of
// sd_kernels.cuh.  There is no CPU fallback: without a device every compute
// entry point returns SD_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
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
#include "sd_kernels.cuh"
#include "sd_blk.h"
#include "sd_blk_host.h"
#include "sd_blkl.h"
#include "sd_obs.h"
#include "sd_blkv.h"
#include "sd_bdot.cuh"
#include "sd_shard_host.h"

#include "sd_handles.h"

static int sd_copy_join(sd_ctx *c);                               // copy engine (below): compute stream waits for the copy streams

// ----------------------------------------------------------------- errors
static thread_local char g_err[512] = "";
int sd_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
SdNccl g_nccl;
static std::mutex g_nccl_mu;
int sd_nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.h) return SD_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return sd_fail(SD_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SD_SYM(field, name)                                                   \
    *(void **)(&g_nccl.field) = dlsym(h, name);                               \
    if (!g_nccl.field) return sd_fail(SD_ERR_NCCL, "libnccl lacks %s", name);
    SD_SYM(GetUniqueId, "ncclGetUniqueId")
    SD_SYM(CommInitRank, "ncclCommInitRank")
    SD_SYM(CommDestroy, "ncclCommDestroy")
    SD_SYM(AllReduce, "ncclAllReduce")
    SD_SYM(AllGather, "ncclAllGather")
    SD_SYM(GetErrorString, "ncclGetErrorString")
#undef SD_SYM
    g_nccl.h = h;
    return SD_OK;
}

static inline unsigned sd_blas_grid(const sd_ctx *c, uint64_t n) {
    uint64_t g = (n + SD_BLAS_THREADS - 1) / SD_BLAS_THREADS;
    const uint64_t cap = (uint64_t)c->sm_count * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}
int sd_launch_check(sd_ctx *c, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sd_fail(SD_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    c->launches++;
    return SD_OK;
}
int sd_use(const sd_ctx *c) {
    SD_CUDA(cudaSetDevice(c->device));
    return SD_OK;
}
int sd_partials_reserve(sd_ctx *c, size_t doubles) {
    if (doubles <= c->partials_cap) return SD_OK;
    if (c->d_partials) { SD_CUDA(cudaStreamSynchronize(c->stream)); SD_CUDA(cudaFree(c->d_partials)); c->d_partials = nullptr; }
    size_t cap = std::max(doubles, (size_t)1 << 16);
    SD_CUDA(cudaMalloc(&c->d_partials, cap * sizeof(double)));
    c->partials_cap = cap;
    return SD_OK;
}
static inline void sd_collective_done(sd_ctx *c) { c->dirty_ids.clear(); c->read_ids.clear(); }
// partials[slot*nparts + i] -> d_scal[slot_out + slot] (+ NCCL sum over ranks)
static int sd_finish_reduce(sd_ctx *c, unsigned nparts, int slotmask, int slot_out) {
    sd_reduce_partials_kernel<<<1, 1024, 0, c->stream>>>(c->d_partials, nparts, slotmask, c->d_scal + slot_out);
    SD_TRY(sd_launch_check(c, "sd_reduce_partials_kernel"));
    if (c->world > 1) {
        SD_NCCL(g_nccl.AllReduce(c->d_scal + slot_out, c->d_scal + slot_out, SD_NSLOT, ncclFloat64_, ncclSum_,
                                 c->comm, c->stream));
        sd_collective_done(c);
    }
    return SD_OK;
}
int sd_fetch(sd_ctx *c, int slot, int n, double *out) {
    SD_CUDA(cudaMemcpyAsync(c->h_scal + slot, c->d_scal + slot, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SD_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) out[i] = c->h_scal[slot + i];
    return SD_OK;
}
// stream-ordered barrier across ranks: peers' earlier kernels have finished
// before any later kernel of this rank runs.
static int sd_rank_barrier(sd_ctx *c) {
    if (c->world <= 1) return SD_OK;
    SD_NCCL(g_nccl.AllReduce(c->d_scal + SD_NSCAL - 8, c->d_scal + SD_NSCAL - 8, 1, ncclFloat64_, ncclSum_,
                             c->comm, c->stream));
    sd_collective_done(c);
    return SD_OK;
}
static inline bool sd_id_in(const std::vector<uint64_t> &v, uint64_t id) {
    for (uint64_t x : v) if (x == id) return true;
    return false;
}
// Call before launching anything that WRITES v: if a peer may still be gathering v (an apply read it since the last
// collective), wait for the peers first; then v counts as not yet visible to the peers.
static int sd_before_write(sd_ctx *c, const sd_vec *v) {
    if (c->world <= 1) return SD_OK;
    if (sd_id_in(c->read_ids, v->id)) SD_TRY(sd_rank_barrier(c));
    if (!sd_id_in(c->dirty_ids, v->id)) c->dirty_ids.push_back(v->id);
    return SD_OK;
}
// Call before an apply gathers psi from the peers' shards and writes out (and acc): psi must be complete everywhere
// (not written since the last collective) and nobody may still be reading out / acc.
static int sd_before_apply(sd_ctx *c, const sd_vec *out, const sd_vec *psi, const sd_vec *acc) {
    if (c->world <= 1) return SD_OK;
    if (sd_id_in(c->dirty_ids, psi->id) || sd_id_in(c->read_ids, out->id) || (acc && sd_id_in(c->read_ids, acc->id)))
        SD_TRY(sd_rank_barrier(c));
    if (!sd_id_in(c->read_ids, psi->id)) c->read_ids.push_back(psi->id);
    if (!sd_id_in(c->dirty_ids, out->id)) c->dirty_ids.push_back(out->id);
    if (acc && !sd_id_in(c->dirty_ids, acc->id)) c->dirty_ids.push_back(acc->id);
    return SD_OK;
}

// All exported functions get C linkage from their declarations in spindyn.h.

const char *sd_last_error(void) { return g_err; }
int sd_version(void) { return SD_VERSION; }
int sd_device_count(int *n) {
    SD_ARG(n, "n is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *n = 0; return sd_fail(SD_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *n = c;
    return SD_OK;
}

// ------------------------------------------------------------------ context
static int sd_ctx_init(int device, int rank, int world, const void *id128, sd_ctx **out) {
    SD_ARG(out, "ctx is NULL");
    *out = nullptr;
    SD_ARG(world >= 1 && world <= SD_MAX_WORLD, "world must be 1..%d", SD_MAX_WORLD);
    SD_ARG(rank >= 0 && rank < world, "rank out of range");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return sd_fail(SD_ERR_CUDA, "no CUDA device available (libspindyn_cuda has no CPU fallback): %s",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    SD_ARG(device >= 0 && device < ndev, "device %d out of range (0..%d)", device, ndev - 1);
    sd_ctx *c = new (std::nothrow) sd_ctx;
    if (!c) return sd_fail(SD_ERR_NOMEM, "out of host memory");
    struct Guard { sd_ctx *c; ~Guard() { if (c) sd_ctx_free(c); } } guard{c};   // early returns below release what was allocated
    c->device = device; c->rank = rank; c->world = world;
    SD_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SD_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    SD_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SD_CUDA(cudaEventCreate(&c->ev0));
    SD_CUDA(cudaEventCreate(&c->ev1));
    c->binom.assign(SD_BINOM_DIM * SD_BINOM_DIM, 0);
    sd_fill_binom(c->binom.data());
    SD_CUDA(cudaMalloc(&c->d_binom, c->binom.size() * sizeof(uint64_t)));
    SD_CUDA(cudaMemcpy(c->d_binom, c->binom.data(), c->binom.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    SD_CUDA(cudaMalloc(&c->d_scal, (SD_NSCAL + 8 * (SD_HIST_MAX + 2)) * sizeof(double)));
    SD_CUDA(cudaMemset(c->d_scal, 0, (SD_NSCAL + 8 * (SD_HIST_MAX + 2)) * sizeof(double)));
    SD_CUDA(cudaMallocHost(&c->h_scal, (SD_NSCAL + 8 * (SD_HIST_MAX + 2)) * sizeof(double)));
    SD_CUDA(cudaMalloc(&c->d_tilectr, 16 * sizeof(unsigned long long)));   // [0]: tile counter of the block kernels' dynamic scheduler
    if (world > 1) {
        SD_ARG(id128, "id128 is NULL");
        SD_TRY(sd_nccl_load());
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        SD_NCCL(g_nccl.CommInitRank(&c->comm, world, id, rank));
        SD_CUDA(cudaMalloc(&c->d_ipc, (size_t)(world + 1) * 128));
    }
    guard.c = nullptr;
    *out = c;
    return SD_OK;
}
int sd_ctx_create(int device, sd_ctx **ctx) { return sd_ctx_init(device, 0, 1, nullptr, ctx); }
int sd_nccl_unique_id(void *id128) {
    SD_ARG(id128, "id128 is NULL");
    SD_TRY(sd_nccl_load());
    ncclUniqueId id;
    SD_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return SD_OK;
}
int sd_ctx_create_rank(int device, int rank, int world, const void *id128, sd_ctx **ctx) {
    return sd_ctx_init(device, rank, world, id128, ctx);
}
int sd_ctx_free(sd_ctx *c) {
    if (!c) return SD_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    cudaFree(c->d_vtab); cudaFree(c->d_rth_partials);
    cudaFree(c->d_binom); cudaFree(c->d_scal); cudaFreeHost(c->h_scal); cudaFree(c->d_partials); cudaFree(c->d_ipc);
    cudaFree(c->scratch[0]); cudaFree(c->scratch[1]); cudaFree(c->d_tilectr);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (cudaEvent_t e : c->ev_copy) if (e) cudaEventDestroy(e);
    if (c->h2d) cudaStreamDestroy(c->h2d);
    if (c->d2h) cudaStreamDestroy(c->d2h);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return SD_OK;
}
int sd_ctx_sync(sd_ctx *c) {
    SD_ARG(c, "ctx is NULL");
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_TRY(sd_copy_join(c));                                        // asynchronous uploads / downloads included
    SD_CUDA(cudaStreamSynchronize(c->stream));
    return SD_OK;
}
int sd_ctx_rank(const sd_ctx *c, int *rank, int *world) {
    SD_ARG(c, "ctx is NULL");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return SD_OK;
}
int sd_timer_start(sd_ctx *c) {
    SD_ARG(c, "ctx is NULL");
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_CUDA(cudaEventRecord(c->ev0, c->stream));
    return SD_OK;
}
int sd_timer_stop(sd_ctx *c, float *ms) {
    SD_ARG(c && ms, "NULL argument");
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_TRY(sd_copy_join(c));                                        // the stopwatch covers the copy streams too
    SD_CUDA(cudaEventRecord(c->ev1, c->stream));
    SD_CUDA(cudaEventSynchronize(c->ev1));
    SD_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return SD_OK;
}
// debug: per-phase cycle sums of the tiled kernel (all zero unless built with -DSD_PHASE_TIMING)
extern "C" int sd_debug_phase_cycles(sd_ctx *c, uint64_t *out8, int reset) {
    SD_ARG(c && out8, "NULL argument");
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long h[16];
    SD_CUDA(cudaMemcpyFromSymbol(h, sd_phase_cycles, sizeof(h)));
    for (int i = 0; i < 16; ++i) out8[i] = h[i];
    if (reset) { memset(h, 0, sizeof(h)); SD_CUDA(cudaMemcpyToSymbol(sd_phase_cycles, h, sizeof(h))); }
    return SD_OK;
}
int sd_launch_count(const sd_ctx *c, uint64_t *n) {
    SD_ARG(c && n, "NULL argument");
    *n = c->launches;
    return SD_OK;
}

// -------------------------------------------------------------------- model
template <typename T>
static int sd_to_device(T **dst, const std::vector<T> &src) {
    *dst = nullptr;
    const size_t n = std::max<size_t>(src.size(), 1);
    SD_CUDA(cudaMalloc(dst, n * sizeof(T)));
    if (!src.empty()) SD_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SD_OK;
}

int sd_env_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

static int sd_tile_setup(sd_model *m, int which, int B) {
    SdTileDev &t = m->tile[which];
    t.ok = false;
    const int L = m->L;
    std::vector<double> Jhop(L, 0.0), Jz(L, 0.0);
    for (size_t b = 0; b < m->hop_a.size(); ++b) if (m->hop_b[b] == m->hop_a[b] + 1) Jhop[m->hop_a[b]] += m->hop_J[b];   // not the wrap bond
    for (size_t b = 0; b < m->zz_a.size(); ++b) if (m->zz_b[b] == m->zz_a[b] + 1) Jz[m->zz_a[b]] += m->zz_J[b];
    if (!sd_tile_build(L, m->k, B, m->tile_T[which], Jhop.data(), Jz.data(), m->field.data(), t.host)) return SD_OK;
    const int nc = which + 1;
    t.cap = t.host.cap_max;
    t.smem = sd_tile_smem_bytes(nc, t.cap, t.host.P.M);
    if (t.smem > 227 * 1024) return SD_OK;
    SD_TRY(sd_to_device((uint16_t **)&t.d_perm, t.host.perm));
    SD_TRY(sd_to_device((SdItem **)&t.d_items, t.host.items));
    SD_TRY(sd_to_device((uint16_t **)&t.d_binomM, t.host.binomM));
    SdTileParams &P = t.host.P;
    P.binom = m->ctx->d_binom;
    P.perm = (const uint16_t *)t.d_perm;
    P.items = (const SdItem *)t.d_items;
    P.pf_dist = sd_env_int("SD_PF_DIST", 128);
    P.binomM = (const uint16_t *)t.d_binomM;
    // neighbour tiles further away than this are streamed evict-first (they cannot be reused from L2)
    P.qfar = sd_tile_qfar(L, P.A, t.host.binom.data(), (uint64_t)sd_env_int("SD_FAR_MB", 100) << 20, 8 * nc);
    t.ok = true;
    return SD_OK;
}

// The compute stream waits for everything the copy streams have been given so far (no host synchronisation).
static int sd_copy_join(sd_ctx *c) {
    if (!c->copy_pending) return SD_OK;
    const size_t e = 2 * SD_COPY_CHUNKS;
    SD_CUDA(cudaEventRecord(c->ev_copy[e + 1], c->h2d)); SD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy[e + 1], 0));
    SD_CUDA(cudaEventRecord(c->ev_copy[e + 2], c->d2h)); SD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy[e + 2], 0));
    c->copy_pending = false; c->d2h_pending = false;
    return SD_OK;
}
static int sd_copy_init(sd_ctx *c) {
    if (c->h2d) return SD_OK;
    SD_CUDA(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
    SD_CUDA(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
    c->ev_copy.assign(2 * SD_COPY_CHUNKS + 4, nullptr);
    for (cudaEvent_t &e : c->ev_copy) SD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return SD_OK;
}
static int sd_scratch_raw(sd_ctx *c, int which, size_t bytes, double **p);
// staging buffer for work on the compute stream: whatever the copy streams still do with it comes first
int sd_scratch(sd_ctx *c, int which, size_t bytes, double **p) {
    SD_TRY(sd_copy_join(c));
    c->stage_touched[which] = true;
    return sd_scratch_raw(c, which, bytes, p);
}
static int sd_scratch_raw(sd_ctx *c, int which, size_t bytes, double **p) {
    if (bytes > c->scratch_cap[which]) {
        SD_TRY(sd_copy_join(c));
        if (c->scratch[which]) { SD_CUDA(cudaStreamSynchronize(c->stream)); SD_CUDA(cudaFree(c->scratch[which])); c->scratch[which] = nullptr; c->scratch_cap[which] = 0; }
        cudaError_t e = cudaMalloc(&c->scratch[which], bytes);
        if (e != cudaSuccess) return sd_fail(SD_ERR_NOMEM, "cudaMalloc of %zu staging bytes failed: %s", bytes, cudaGetErrorString(e));
        c->scratch_cap[which] = bytes;
    }
    *p = (double *)c->scratch[which];
    return SD_OK;
}
void sd_scratch_release(sd_ctx *c) {          // staging is only kept while it is small
    for (int w = 0; w < 2; ++w)
        if (c->scratch_cap[w] > ((size_t)256 << 20)) {
            sd_copy_join(c);
            cudaStreamSynchronize(c->stream);
            cudaFree(c->scratch[w]); c->scratch[w] = nullptr; c->scratch_cap[w] = 0;
        }
}

// block-layout kernel tables (sd_blk.h); needs the tiled tables too (same tile keys: B = 15)
static int sd_blk_setup(sd_model *m) {
    SdBlkDev &b = m->blk;
    b.ok = false;
    const int L = m->L;
    if (!m->tile_capable || m->tile[0].host.P.B != SD_BLK_B || L - SD_BLK_B < 1) return SD_OK;
    std::vector<double> Jhop(L + 1, 0.0), Jz(L + 1, 0.0);
    for (size_t i = 0; i < m->hop_a.size(); ++i) if (m->hop_b[i] == m->hop_a[i] + 1) Jhop[m->hop_a[i]] += m->hop_J[i];   // not the wrap bond
    for (size_t i = 0; i < m->zz_a.size(); ++i) if (m->zz_b[i] == m->zz_a[i] + 1) Jz[m->zz_a[i]] += m->zz_J[i];
    if (!sd_blk_build(L, m->k, Jhop.data(), Jz.data(), m->field.data(), b.host)) return SD_OK;
    b.threads = sd_env_int("SD_BLKL_THREADS", 640);
    b.pfp = sd_env_int("SD_BLK_PFP", 7);      // all prefix partners of the tile the consumers reach next (profiles/round2_n_ab.txt, round2_o_blkenv.txt)
    if (b.threads != 512 && b.threads != 768) b.threads = 640;
    for (int w = 0; w < 2; ++w) {
        const int nc = w + 1;
        int nbuf = w == 0 ? 3 : 2;
        for (; nbuf >= 2; --nbuf) {
            // the tile buffers are the dynamic part; tables, headers and context are static (sizeof(SdBlkShared))
            b.smem[w] = (size_t)nbuf * b.host.P.cap * nc * sizeof(double);
            if (b.smem[w] + sizeof(SdBlkShared) + 128 <= 227 * 1024) break;
        }
        if (nbuf < 2) return SD_OK;
        b.nbuf[w] = nbuf;
        b.qfar[w] = sd_tile_qfar(L, b.host.P.A, b.host.binom.data(), (uint64_t)sd_env_int("SD_FAR_MB", 100) << 20, 8 * nc);
    }
    SD_TRY(sd_to_device(&b.d_W, b.host.W));
    SD_TRY(sd_to_device(&b.d_js, b.host.js));
    SD_TRY(sd_to_device(&b.d_units, b.host.units));
    SD_TRY(sd_to_device(&b.d_items, b.host.items));
    SD_TRY(sd_to_device(&b.d_dmid, b.host.dmid));
    SdBlkParams &P = b.host.P;
    P.W = b.d_W; P.js = b.d_js; P.units = b.d_units; P.items = b.d_items; P.dmid = b.d_dmid;
    b.ok = true;
    return SD_OK;
}
// parameters of a launch on this rank's shard
static SdBlkParams sd_blk_params(const sd_model *m, int nc) {
    SdBlkParams P = m->blk.host.P;
    const sd_ctx *c = m->ctx;
    P.nbuf = m->blk.nbuf[nc - 1];
    P.pfp = m->blk.pfp;
    P.wrap_on = m->has_wrap ? 1 : 0; P.wrapJ = m->wrap_hop; P.wrapJz4 = 0.25 * m->wrap_zz;
    P.order = m->blk.d_order; P.norder = m->blk.norder;
    P.key_lo = m->tile[0].keys[c->rank];
    P.key_hi = m->tile[0].keys[c->rank + 1];
    P.shards.world = c->world; P.shards.rank = c->rank;
    for (int g = 0; g <= SD_MAX_WORLD; ++g) P.shards.pstart[g] = m->blk.pstart[std::min(g, c->world)];
    return P;
}


int sd_model_create(sd_ctx *ctx, int L, int nup, const sd_bond *hop, int nhop, const sd_bond *zz, int nzz,
                    const double *field, sd_model **model) {
    SD_ARG(ctx && model, "NULL argument");
    *model = nullptr;
    // Basis.jl:9-20
    SD_ARG(L >= 1, "L must be at least 1");
    SD_ARG(L <= SD_MAX_L, "L must be at most 63 when using UInt64 basis states");
    SD_ARG(nup >= -1 && nup <= L, "nup must satisfy 0 <= nup <= L");
    SD_ARG(nhop >= 0 && nzz >= 0 && (nhop == 0 || hop) && (nzz == 0 || zz) && field, "bad bond lists");
    SD_LOCK(ctx); SD_TRY(sd_use(ctx));
    sd_model *m = new (std::nothrow) sd_model;
    if (!m) return sd_fail(SD_ERR_NOMEM, "out of host memory");
    struct Guard { sd_model *m; ~Guard() { if (m) sd_model_free(m); } } guard{m};   // early returns below release the model and its device tables
    m->ctx = ctx; m->L = L; m->k = nup;
    m->N = nup < 0 ? (1ULL << L) : ctx->binom[L * SD_BINOM_DIM + nup];
    bool all_nn = true, wrap = false;                               // every bond nearest-neighbour, except possibly the wrap bond
    for (int b = 0; b < nhop; ++b) {
        const int64_t i = hop[b].i, j = hop[b].j;
        if (i < 1 || i > L || j < 1 || j > L) return sd_fail(SD_ERR_ARG, "hopping site outside 1..L");
        if (i == j) continue;                       // bits never differ: the reference skips it
        const int a = (int)std::min(i, j) - 1, c = (int)std::max(i, j) - 1;
        m->hop_a.push_back(a); m->hop_b.push_back(c); m->hop_J.push_back(hop[b].J);
        if (a == 0 && c == L - 1 && L > 2) { wrap = true; m->wrap_hop += hop[b].J; }   // SpinModel.jl:71-78 periodic chain
        else if (c != a + 1) all_nn = false;
    }
    for (int b = 0; b < nzz; ++b) {
        const int64_t i = zz[b].i, j = zz[b].j;
        if (i < 1 || i > L || j < 1 || j > L) return sd_fail(SD_ERR_ARG, "zz site outside 1..L");
        const int a = (int)std::min(i, j) - 1, c = (int)std::max(i, j) - 1;
        m->zz_a.push_back(a); m->zz_b.push_back(c); m->zz_J.push_back(zz[b].J);
        if (a == 0 && c == L - 1 && L > 2) { wrap = true; m->wrap_zz += zz[b].J; }
        else if (c != a + 1) all_nn = false;
    }
    m->field.assign(field, field + L);
    SD_TRY(sd_to_device(&m->d_hop_a, m->hop_a));
    SD_TRY(sd_to_device(&m->d_hop_b, m->hop_b));
    SD_TRY(sd_to_device(&m->d_hop_J, m->hop_J));
    SD_TRY(sd_to_device(&m->d_zz_a, m->zz_a));
    SD_TRY(sd_to_device(&m->d_zz_b, m->zz_b));
    SD_TRY(sd_to_device(&m->d_zz_J, m->zz_J));
    SD_TRY(sd_to_device(&m->d_field, m->field));
    // two-table ranking for non-nearest-neighbour hops in a sector (L <= 40)
    bool need_lin = false;
    for (size_t b = 0; b < m->hop_a.size(); ++b) if (m->hop_b[b] != m->hop_a[b] + 1) need_lin = true;
    if (nup >= 0 && need_lin && L >= 2 && L <= 40) {
        const int h = L / 2, g = L - h;
        const uint64_t *C = ctx->binom.data();
        std::vector<uint64_t> A((size_t)1 << h, 0), Bt((size_t)1 << g, 0);
        for (uint64_t lo = 0; lo < (1ULL << h); ++lo) {
            const int pl = __builtin_popcountll(lo);
            if (pl > nup || nup - pl > g) continue;
            uint64_t acc = 0;
            int rem = nup;                           // set bits at positions >= q
            for (int q = 0; q < h; ++q) {
                if ((lo >> q) & 1ULL) --rem;
                else acc += sd_binom_at(C, SD_BINOM_DIM, L - 1 - q, rem - 1);
            }
            A[lo] = acc;
        }
        for (uint64_t hi = 0; hi < (1ULL << g); ++hi) {
            int rem = __builtin_popcountll(hi);
            if (rem > nup) continue;
            uint64_t acc = 0;
            for (int q = 0; q < g; ++q) {
                if ((hi >> q) & 1ULL) --rem;
                else acc += sd_binom_at(C, SD_BINOM_DIM, L - 1 - (h + q), rem - 1);
            }
            Bt[hi] = acc;
        }
        SD_TRY(sd_to_device(&m->d_linA, A));
        SD_TRY(sd_to_device(&m->d_linB, Bt));
        m->lin_h = h;
    }
    // tiled path: sector basis, every bond nearest-neighbour
    m->tile_threads = sd_env_int("SD_TILE_THREADS", 512) == 256 ? 256 : 512;
    m->tile_T[0] = sd_env_int("SD_TILE_T", 5);
    m->tile_T[1] = sd_env_int("SD_TILE_T_C128", 4);
    m->tile_capable = false;
    if (nup >= 0 && all_nn && L >= 10 && !sd_env_int("SD_FORCE_GENERIC", 0)) {
        const int B1 = std::min(L, sd_env_int("SD_TILE_B", 15));
        const int B2 = std::min(L, sd_env_int("SD_TILE_B_C128", 13));
        SD_TRY(sd_tile_setup(m, 0, B1));
        SD_TRY(sd_tile_setup(m, 1, std::min(B1, B2)));
        m->tile_capable = m->tile[0].ok && m->tile[1].ok;
    }
    m->has_wrap = wrap && all_nn;
    m->path = m->tile_capable ? SD_PATH_TILED : SD_PATH_GENERIC;
    if (m->tile_capable && sd_env_int("SD_BLK", 1)) {
        SD_TRY(sd_blk_setup(m));
        if (m->blk.ok) m->path = SD_PATH_BLOCK;
        // A chain of L = 16 .. 21 has 2 .. 64 tiles, i.e. that many CTAs on 148 SMs; below SD_BLK_MIN_TILES prefixes the
        // model starts on the one-thread-per-state kernel instead (0 = off, the measured default; round-2 A/B).
        const int min_tiles = sd_env_int("SD_BLK_MIN_TILES", 0);
        if (m->blk.ok && min_tiles > 0 && (1ULL << m->blk.host.P.A) < (uint64_t)min_tiles) m->path = SD_PATH_GENERIC;
    }
    // a periodic chain runs on the block kernel (WRAP variant) or on the generic one; the tiled kernel has no wrap bond
    if (m->has_wrap && m->path == SD_PATH_TILED) m->path = SD_PATH_GENERIC;
    // shards: tile-aligned to the coarser (F64) tiling when tiled, plain equal split otherwise
    uint64_t bounds[SD_MAX_WORLD + 1];
    if (m->tile_capable) {
        sd_tile_shard_bounds(m->tile[0].host, ctx->world, bounds, m->tile[0].keys);
        if (m->blk.ok && ctx->world > 2 && sd_env_int("SD_SHARD_BALANCE", 1)) {
            // Shards weighted by their remote volume (sd_shard_balance): with equal rank ranges the ranks whose top prefix
            // bits are 101 / 010 gather 2.5 shards' worth over NVLink, the edge ranks 0.5; measured at L = 32 on 8 GPUs:
            // 2.62 ms per apply with equal shards, 1.81 ms weighted (profiles/round2_e_8gpu.txt).  SD_SHARD_REMOTE_COST
            // (percent): NVLink time per remote element / compute time per local one.  SD_SHARD_BALANCE=0: equal shards.
            double cost[2] = {0.0, 0.0};
            uint64_t wb[SD_MAX_WORLD + 1], wk[SD_MAX_WORLD + 1];
            if (sd_shard_balance(m->blk.host, m->tile[0].host, ctx->world, m->blk.qfar[0], sd_env_int("SD_SHARD_REMOTE_COST", 130) / 100.0,
                                std::max(1, std::min(20, sd_env_int("SD_SHARD_BALANCE_ITERS", 6))), wb, wk, cost)) {
                for (int g = 0; g <= ctx->world; ++g) { bounds[g] = wb[g]; m->tile[0].keys[g] = wk[g]; }
            }
        }
        for (int g = 0; g <= ctx->world; ++g) {
            uint64_t base = 0;
            m->tile[1].keys[g] = sd_tile_key_of_rank(m->tile[1].host, bounds[g], &base);
            if (base != bounds[g]) return sd_fail(SD_ERR_UNSUPPORTED, "internal: shard bound not tile aligned");
        }
    } else {
        for (int g = 0; g <= ctx->world; ++g)
            bounds[g] = (g == ctx->world) ? m->N : (uint64_t)(((unsigned __int128)m->N * g) / ctx->world);
    }
    m->shards.world = ctx->world; m->shards.rank = ctx->rank;
    for (int g = 0; g <= SD_MAX_WORLD; ++g) m->shards.start[g] = bounds[std::min(g, ctx->world)];
    if (m->blk.ok)
        for (int g = 0; g <= ctx->world; ++g) m->blk.pstart[g] = sd_blk_key_base(m->blk.host, m->tile[0].keys[g]);
    m->blk_layout = (m->path == SD_PATH_BLOCK);
    // Breadth-first tile order (sd_blk_tile_order) once a vector no longer fits the L2: measured 14.7 GB instead of 22.4 GB
    // of DRAM reads per L = 32 apply (profiles/round2_a_ab.txt).  SD_BLK_ORDER=0 keeps rank order (A/B only).
    if (m->blk.ok && m->blk.host.n_store * sizeof(double) > ((size_t)48 << 20) && sd_env_int("SD_BLK_ORDER", 2) != 0) {
        std::vector<uint32_t> ord;
        sd_blk_tile_order(m->blk.host, m->tile[0].keys[ctx->rank], m->tile[0].keys[ctx->rank + 1], m->blk.host.P.A - 1, ord);
        if (!ord.empty()) {
            SD_TRY(sd_to_device(&m->blk.d_order, ord));
            m->blk.norder = (uint32_t)ord.size();
        }
    }
    guard.m = nullptr;
    *model = m;
    return SD_OK;
}

static void sd_pool_drain(sd_model *m) {
    std::vector<sd_vec *> p;
    p.swap(m->pool);
    for (sd_vec *v : p) { m->live_vecs++; sd_vec_free(v); }
}
// Finalizer order is unspecified (Julia, Python shutdown): a model that still has live vectors is only marked and
// released by the sd_vec_free of its last vector.
int sd_model_free(sd_model *m) {
    if (!m) return SD_OK;
    {
        SD_LOCK(m->ctx);
        sd_pool_drain(m);
        if (m->live_vecs > 0) { m->free_pending = true; return SD_OK; }
    }
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->d_hop_a); cudaFree(m->d_hop_b); cudaFree(m->d_hop_J);
    cudaFree(m->d_zz_a); cudaFree(m->d_zz_b); cudaFree(m->d_zz_J); cudaFree(m->d_field);
    cudaFree(m->d_linA); cudaFree(m->d_linB);
    for (int w = 0; w < 2; ++w) {
        SdTileDev &t = m->tile[w];
        cudaFree(t.d_perm); cudaFree(t.d_items); cudaFree(t.d_binomM);
    }
    cudaFree(m->blk.d_order);
    cudaFree(m->blk.d_W); cudaFree(m->blk.d_js); cudaFree(m->blk.d_units); cudaFree(m->blk.d_items); cudaFree(m->blk.d_dmid);
    delete m;
    return SD_OK;
}
int sd_model_dim(const sd_model *m, uint64_t *dim) {
    SD_ARG(m && dim, "NULL argument");
    *dim = m->N;
    return SD_OK;
}
int sd_model_local_range(const sd_model *m, uint64_t *first, uint64_t *count) {
    SD_ARG(m, "NULL argument");
    const uint64_t a = m->shards.start[m->shards.rank], b = m->shards.start[m->shards.rank + 1];
    if (first) *first = a;
    if (count) *count = b - a;
    return SD_OK;
}
int sd_model_shard_bounds(const sd_model *m, int world, uint64_t *bounds) {
    SD_ARG(m && bounds, "NULL argument");
    SD_ARG(world >= 1 && world <= SD_MAX_WORLD, "world must be 1..%d", SD_MAX_WORLD);
    if (world == m->shards.world)                                    // this context's split (may be weighted: SD_SHARD_BALANCE)
        for (int g = 0; g <= world; ++g) bounds[g] = m->shards.start[g];
    else if (m->tile_capable) sd_tile_shard_bounds(m->tile[0].host, world, bounds, nullptr);
    else
        for (int g = 0; g <= world; ++g)
            bounds[g] = (g == world) ? m->N : (uint64_t)(((unsigned __int128)m->N * g) / world);
    return SD_OK;
}
int sd_model_info(const sd_model *m, int *kernel_path, int *tile_sites, int *rank_bits) {
    SD_ARG(m, "NULL argument");
    if (kernel_path) *kernel_path = m->path;
    if (tile_sites) *tile_sites = m->tile_capable ? m->tile[0].host.P.B : 0;
    if (rank_bits) *rank_bits = (m->N > 0xffffffffULL) ? 64 : 32;
    return SD_OK;
}
int sd_model_set_path(sd_model *m, int kernel_path) {
    SD_ARG(m, "NULL argument");
    // the block kernel works on block-layout vectors, the other two on rank-ordered ones: the
    // layout can only change while no vector of the model is alive
    const bool want_blk = kernel_path == SD_PATH_BLOCK;
    if (want_blk != m->blk_layout) { SD_LOCK(m->ctx); sd_pool_drain(m); }   // idle work vectors have the old layout
    if (want_blk != m->blk_layout && m->live_vecs > 0)
        return sd_fail(SD_ERR_UNSUPPORTED, "cannot switch between the block kernel and the rank-ordered kernels while %d vectors of the model are alive", m->live_vecs);
    if (kernel_path == SD_PATH_GENERIC) { m->path = SD_PATH_GENERIC; m->blk_layout = false; return SD_OK; }
    if (kernel_path == SD_PATH_TILED) {
        if (!m->tile_capable || m->has_wrap)
            return sd_fail(SD_ERR_UNSUPPORTED, "model does not qualify for the tiled kernel (sector basis, open chain of nearest-neighbour bonds, L >= 10)");
        m->path = SD_PATH_TILED; m->blk_layout = false;
        return SD_OK;
    }
    if (kernel_path == SD_PATH_BLOCK) {
        if (!m->blk.ok)
            return sd_fail(SD_ERR_UNSUPPORTED, "model does not qualify for the block kernel (sector basis, nearest-neighbour bonds, L >= 16)");
        m->path = SD_PATH_BLOCK; m->blk_layout = true;
        return SD_OK;
    }
    return sd_fail(SD_ERR_ARG, "unknown kernel path %d", kernel_path);
}

// -------------------------------------------------------------------- basis
int sd_unrank(sd_model *m, uint64_t first, uint64_t count, uint64_t *states) {
    SD_ARG(m && (states || count == 0), "NULL argument");
    SD_ARG(first <= m->N && count <= m->N - first, "range outside the basis");
    if (count == 0) return SD_OK;
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    uint64_t *d = nullptr;
    SD_CUDA(cudaMalloc(&d, count * sizeof(uint64_t)));
    sd_unrank_kernel<<<sd_blas_grid(c, count), SD_BLAS_THREADS, 0, c->stream>>>(m->L, m->k, c->d_binom, first, count, d);
    int rc = sd_launch_check(c, "sd_unrank_kernel");
    if (rc == SD_OK) {
        cudaError_t e = cudaMemcpyAsync(states, d, count * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = sd_fail(SD_ERR_CUDA, "sd_unrank copy: %s", cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}
int sd_rank(sd_model *m, const uint64_t *states, uint64_t count, int64_t *idx1) {
    SD_ARG(m && ((states && idx1) || count == 0), "NULL argument");
    if (count == 0) return SD_OK;
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    uint64_t *ds = nullptr;
    int64_t *di = nullptr;
    SD_CUDA(cudaMalloc(&ds, count * sizeof(uint64_t)));
    cudaError_t e = cudaMalloc(&di, count * sizeof(int64_t));
    if (e != cudaSuccess) { cudaFree(ds); return sd_fail(SD_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); }
    int rc = SD_OK;
    e = cudaMemcpyAsync(ds, states, count * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        sd_rank_kernel<<<sd_blas_grid(c, count), SD_BLAS_THREADS, 0, c->stream>>>(m->L, m->k, c->d_binom, ds, count, di);
        rc = sd_launch_check(c, "sd_rank_kernel");
        if (rc == SD_OK) {
            e = cudaMemcpyAsync(idx1, di, count * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        }
    }
    if (rc == SD_OK && e != cudaSuccess) rc = sd_fail(SD_ERR_CUDA, "sd_rank: %s", cudaGetErrorString(e));
    cudaFree(ds); cudaFree(di);
    return rc;
}

// ------------------------------------------------------------------ vectors
// One all-gather of 128 bytes per rank: [0,64) the IPC handle of *dptr (zeros if dptr is null), [64,120) up to 7 vector
// ids this rank has freed since its last announcement, [120,128) how many more it still has to announce.  Every rank
// counts the announcements; a shard is released when all ranks have freed its vector.
static int sd_exchange(sd_ctx *c, double **dptr) {
    unsigned char mine[128];
    memset(mine, 0, sizeof(mine));
    if (dptr) {
        cudaIpcMemHandle_t hnd;
        SD_CUDA(cudaIpcGetMemHandle(&hnd, *dptr));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        memcpy(mine, &hnd, 64);
    }
    const size_t nsend = std::min<size_t>(7, c->outbox.size());
    for (size_t i = 0; i < nsend; ++i) memcpy(mine + 64 + 8 * i, &c->outbox[i], 8);
    c->outbox.erase(c->outbox.begin(), c->outbox.begin() + nsend);
    const uint64_t pending = c->outbox.size();
    memcpy(mine + 120, &pending, 8);
    unsigned char *d_mine = c->d_ipc + (size_t)c->world * 128;
    SD_CUDA(cudaMemcpyAsync(d_mine, mine, 128, cudaMemcpyHostToDevice, c->stream));
    SD_NCCL(g_nccl.AllGather(d_mine, c->d_ipc, 128, ncclUint8_, c->comm, c->stream));
    sd_collective_done(c);
    c->h_ipc.resize((size_t)c->world * 128);
    SD_CUDA(cudaMemcpyAsync(c->h_ipc.data(), c->d_ipc, (size_t)c->world * 128, cudaMemcpyDeviceToHost, c->stream));
    SD_CUDA(cudaStreamSynchronize(c->stream));
    for (int g = 0; g < c->world; ++g)
        for (int i = 0; i < 7; ++i) {
            uint64_t id;
            memcpy(&id, c->h_ipc.data() + (size_t)g * 128 + 64 + 8 * i, 8);
            if (!id) continue;
            size_t k = 0;
            while (k < c->freed_count.size() && c->freed_count[k].first != id) ++k;
            if (k == c->freed_count.size()) c->freed_count.push_back({id, 0});
            if (++c->freed_count[k].second == c->world) {            // every rank has unmapped it: release the local shard
                for (size_t j = 0; j < c->dead.size(); ++j)
                    if (c->dead[j].id == id) { cudaFree(c->dead[j].d); c->dead.erase(c->dead.begin() + j); break; }
                c->freed_count.erase(c->freed_count.begin() + k);
            }
        }
    return SD_OK;
}
int sd_vec_alloc(sd_model *m, int dtype, sd_vec **vec) {
    SD_ARG(m && vec, "NULL argument");
    *vec = nullptr;
    SD_ARG(dtype == SD_F64 || dtype == SD_C128, "dtype must be SD_F64 or SD_C128");
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    sd_vec *v = new (std::nothrow) sd_vec;
    if (!v) return sd_fail(SD_ERR_NOMEM, "out of host memory");
    v->model = m; v->dtype = dtype; v->nc = dtype == SD_C128 ? 2 : 1;
    v->layout = m->blk_layout ? 1 : 0;
    const uint64_t *starts = v->layout ? m->blk.pstart : m->shards.start;
    const uint64_t ls = starts[c->rank];
    v->local_n = starts[c->rank + 1] - ls;
    v->logical_n = m->shards.start[c->rank + 1] - m->shards.start[c->rank];
    for (int g = 0; g < SD_MAX_WORLD; ++g) { v->peer[g] = nullptr; v->view.base[g] = nullptr; }
    // +2 elements of slack so 16-byte vector accesses at the ends stay inside the allocation
    const size_t bytes = (size_t)(v->local_n + 2) * v->nc * sizeof(double);
    cudaError_t e = cudaMalloc(&v->d, bytes);
    if (e != cudaSuccess && !m->pool.empty() && c->world == 1) {     // idle work vectors are the first thing to give back
        cudaGetLastError();
        sd_pool_drain(m);
        e = cudaMalloc(&v->d, bytes);
    }
    if (e != cudaSuccess) {
        delete v;
        return sd_fail(SD_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    if (v->layout) {                                  // block layout: padding must be (and stays) zero
        e = cudaMemsetAsync(v->d, 0, bytes, c->stream);
        if (e != cudaSuccess) { cudaFree(v->d); delete v; return sd_fail(SD_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e)); }
    }
    v->view.base[c->rank] = v->d - (int64_t)ls * v->nc;
    v->id = c->next_vec_id++;
    m->live_vecs++;
    if (c->world > 1) {
        // collective: exchange CUDA IPC handles (64 bytes) and up to 7 locally freed vector ids (64 bytes), map every peer shard
        SD_TRY(sd_exchange(c, &v->d));
        for (int g = 0; g < c->world; ++g) {
            if (g == c->rank) continue;
            cudaIpcMemHandle_t hnd;
            memcpy(&hnd, c->h_ipc.data() + (size_t)g * 128, 64);
            void *p = nullptr;
            SD_CUDA(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
            v->peer[g] = p;
            v->view.base[g] = (const double *)p - (int64_t)starts[g] * v->nc;
        }
    }
    *vec = v;
    return SD_OK;
}
// Not collective.  With world > 1 the local shard stays allocated (peers may still have it mapped and may still read
// it) until every rank has freed the same vector and said so in a later collective (sd_vec_alloc / sd_ctx_collect);
// finalizers may therefore call this at any time and in any order.
int sd_vec_free(sd_vec *v) {
    if (!v) return SD_OK;
    sd_ctx *c = v->model->ctx;
    SD_LOCK(c);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int g = 0; g < SD_MAX_WORLD; ++g)
        if (v->peer[g]) cudaIpcCloseMemHandle(v->peer[g]);
    if (c->world > 1) {
        c->outbox.push_back(v->id);
        if (v->owned) c->dead.push_back({v->id, v->d});
    } else if (v->owned) {
        cudaFree(v->d);
    }
    sd_model *m = v->model;
    const bool last = --m->live_vecs == 0 && m->free_pending;
    delete v;
    if (last) { m->free_pending = false; return sd_model_free(m); }   // the lock is recursive
    return SD_OK;
}
// Collective (world > 1): announces locally freed vectors until every rank's list is empty and releases the shards
// that all ranks have freed.  sd_vec_alloc does one round of the same exchange; call this to return memory earlier.
int sd_ctx_collect(sd_ctx *c) {
    SD_ARG(c, "ctx is NULL");
    SD_LOCK(c); SD_TRY(sd_use(c));
    if (c->world <= 1) return SD_OK;
    for (;;) {
        SD_TRY(sd_exchange(c, nullptr));
        bool more = false;
        for (int g = 0; g < c->world; ++g) {
            uint64_t pending;
            memcpy(&pending, c->h_ipc.data() + (size_t)g * 128 + 120, 8);
            if (pending) more = true;
        }
        if (!more) break;
    }
    return SD_OK;
}
int sd_vec_dtype(const sd_vec *v, int *dtype) {
    SD_ARG(v && dtype, "NULL argument");
    *dtype = v->dtype;
    return SD_OK;
}
int sd_vec_local_len(const sd_vec *v, uint64_t *n) {
    SD_ARG(v && n, "NULL argument");
    *n = v->logical_n;
    return SD_OK;
}
static inline size_t sd_vec_bytes(const sd_vec *v) { return (size_t)v->local_n * v->nc * sizeof(double); }
static inline size_t sd_vec_logical_bytes(const sd_vec *v) { return (size_t)v->logical_n * v->nc * sizeof(double); }
// block layout <-> rank order on the local shard.  dir 0: blk := rank-ordered (or seeded values), 1: rank-ordered := blk
int sd_blk_permute(const sd_vec *v, double *rank_local, int nc_rank, int dir, int seeded, uint64_t seed, double scale) {
    sd_model *m = v->model;
    sd_ctx *c = m->ctx;
    SdBlkParams P = sd_blk_params(m, v->nc);
    const uint64_t nkeys = P.key_hi - P.key_lo;
    if (nkeys == 0) return SD_OK;
    SdBlkPermute Q;
    Q.dir = dir; Q.nc_blk = v->nc; Q.nc_rank = nc_rank; Q.seeded = seeded; Q.seed = seed; Q.scale = scale;
    Q.rstart = m->shards.start[c->rank];
    Q.binom = c->d_binom;
    const unsigned grid = (unsigned)std::min<uint64_t>(nkeys, (uint64_t)c->sm_count * 16);
    sd_blk_permute_kernel<<<grid, 256, 0, c->stream>>>(P, Q, v->d, rank_local);
    return sd_launch_check(c, "sd_blk_permute_kernel");
}

static int sd_blk_permute_range(const sd_vec *v, double *rank_local, int nc_rank, int dir, uint64_t key_lo, uint64_t key_hi) {
    sd_model *m = v->model;
    sd_ctx *c = m->ctx;
    SdBlkParams P = sd_blk_params(m, v->nc);
    P.key_lo = key_lo; P.key_hi = key_hi;
    if (key_hi <= key_lo) return SD_OK;
    SdBlkPermute Q;
    Q.dir = dir; Q.nc_blk = v->nc; Q.nc_rank = nc_rank; Q.seeded = 0; Q.seed = 0; Q.scale = 0.0;
    Q.rstart = m->shards.start[c->rank];
    Q.binom = c->d_binom;
    const unsigned grid = (unsigned)std::min<uint64_t>(key_hi - key_lo, (uint64_t)c->sm_count * 16);
    sd_blk_permute_kernel<<<grid, 256, 0, c->stream>>>(P, Q, v->d, rank_local);
    return sd_launch_check(c, "sd_blk_permute_kernel");
}
// chunk boundaries of the copy engine on this rank's shard: tile keys and the basis ranks (relative to the shard) they start at
static void sd_copy_chunks(sd_model *m) {
    if (!m->cp_keys.empty()) return;
    const sd_ctx *c = m->ctx;
    const uint64_t r0 = m->shards.start[c->rank], r1 = m->shards.start[c->rank + 1];
    const uint64_t k0 = m->tile[0].keys[c->rank], k1 = m->tile[0].keys[c->rank + 1];
    const int K = (r1 - r0) * sizeof(double) < ((uint64_t)64 << 20) ? 1 : SD_COPY_CHUNKS;
    m->cp_keys.assign(1, k0); m->cp_ranks.assign(1, 0);
    for (int i = 1; i < K; ++i) {
        uint64_t base = 0;
        const uint64_t key = sd_tile_key_of_rank(m->tile[0].host, r0 + (uint64_t)(((unsigned __int128)(r1 - r0) * i) / K), &base);
        if (key <= m->cp_keys.back() || key >= k1) continue;
        m->cp_keys.push_back(key); m->cp_ranks.push_back(base - r0);
    }
    m->cp_keys.push_back(k1); m->cp_ranks.push_back(r1 - r0);
}

int sd_vec_upload(sd_vec *v, const void *host) {
    SD_ARG(v && host, "NULL argument");
    sd_ctx *c = v->model->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_TRY(sd_before_write(c, v));
    if (v->layout) {
        double *st = nullptr;
        SD_TRY(sd_scratch(c, 0, sd_vec_logical_bytes(v) + 16, &st));
        SD_CUDA(cudaMemcpyAsync(st, host, sd_vec_logical_bytes(v), cudaMemcpyHostToDevice, c->stream));
        SD_TRY(sd_blk_permute(v, st, v->nc, 0, 0, 0, 0.0));
        SD_CUDA(cudaStreamSynchronize(c->stream));
        sd_scratch_release(c);
        return SD_OK;
    }
    SD_CUDA(cudaMemcpyAsync(v->d, host, sd_vec_bytes(v), cudaMemcpyHostToDevice, c->stream));
    SD_CUDA(cudaStreamSynchronize(c->stream));
    return SD_OK;
}
int sd_vec_download(sd_vec *v, void *host) {
    SD_ARG(v && host, "NULL argument");
    sd_ctx *c = v->model->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    if (v->layout) {
        double *st = nullptr;
        SD_TRY(sd_scratch(c, 1, sd_vec_logical_bytes(v) + 16, &st));
        SD_TRY(sd_blk_permute(v, st, v->nc, 1, 0, 0, 0.0));
        SD_CUDA(cudaMemcpyAsync(host, st, sd_vec_logical_bytes(v), cudaMemcpyDeviceToHost, c->stream));
        SD_CUDA(cudaStreamSynchronize(c->stream));
        sd_scratch_release(c);
        return SD_OK;
    }
    SD_CUDA(cudaMemcpyAsync(host, v->d, sd_vec_bytes(v), cudaMemcpyDeviceToHost, c->stream));
    SD_CUDA(cudaStreamSynchronize(c->stream));
    return SD_OK;
}
int sd_vec_upload_async(sd_vec *v, const void *host) {
    SD_ARG(v && host, "NULL argument");
    sd_ctx *c = v->model->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_TRY(sd_before_write(c, v));
    if (v->layout) {
        // pinned host -> rank-ordered staging on the h2d stream, chunk by chunk; the permute of chunk i into block layout
        // runs on the compute stream while chunk i + 1 is on the wire.
        sd_model *m = v->model;
        SD_TRY(sd_copy_init(c));
        sd_copy_chunks(m);
        double *st = nullptr;
        SD_TRY(sd_scratch_raw(c, 0, sd_vec_logical_bytes(v) + 16, &st));
        const size_t e = 2 * SD_COPY_CHUNKS, esz = (size_t)v->nc * sizeof(double);
        // the staging buffer is free once the permutes of the previous upload are done: wait for exactly that, not for the
        // whole compute stream (which may be parked behind a download), so uploads run one step ahead of the kernels
        if (c->stage_touched[0] || !c->up_event) SD_CUDA(cudaEventRecord(c->ev_copy[e], c->stream));
        c->stage_touched[0] = false;
        SD_CUDA(cudaStreamWaitEvent(c->h2d, c->ev_copy[e], 0));
        for (size_t i = 0; i + 1 < m->cp_keys.size(); ++i) {
            const size_t off = m->cp_ranks[i] * esz, len = (m->cp_ranks[i + 1] - m->cp_ranks[i]) * esz;
            if (len) SD_CUDA(cudaMemcpyAsync((char *)st + off, (const char *)host + off, len, cudaMemcpyHostToDevice, c->h2d));
            SD_CUDA(cudaEventRecord(c->ev_copy[i], c->h2d));
            SD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy[i], 0));
            SD_TRY(sd_blk_permute_range(v, st, v->nc, 0, m->cp_keys[i], m->cp_keys[i + 1]));
        }
        SD_CUDA(cudaEventRecord(c->ev_copy[e], c->stream));         // staging free again
        c->up_event = true;
        return SD_OK;                                               // the compute stream has waited for every chunk: nothing pending
    }
    SD_CUDA(cudaMemcpyAsync(v->d, host, sd_vec_bytes(v), cudaMemcpyHostToDevice, c->stream));
    return SD_OK;
}
int sd_vec_download_async(sd_vec *v, void *host) {
    SD_ARG(v && host, "NULL argument");
    sd_ctx *c = v->model->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    if (v->layout) {
        // permute of chunk i into the rank-ordered staging on the compute stream, its copy to the pinned host buffer on
        // the d2h stream.  v itself is free again when the permutes are done (stream order); the staging buffer is not
        // until the copies are, so the next download's permutes wait for this one's last copy (d2h_pending) -- an upload
        // issued in between does not, which is what lets both PCIe directions run at once.  The data is on the host
        // after sd_ctx_sync (or sd_timer_stop).
        sd_model *m = v->model;
        SD_TRY(sd_copy_init(c));
        sd_copy_chunks(m);
        double *st = nullptr;
        SD_TRY(sd_scratch_raw(c, 1, sd_vec_logical_bytes(v) + 16, &st));
        const size_t e = 2 * SD_COPY_CHUNKS, esz = (size_t)v->nc * sizeof(double);
        if (c->d2h_pending) SD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy[e + 3], 0));
        for (size_t i = 0; i + 1 < m->cp_keys.size(); ++i) {
            const size_t off = m->cp_ranks[i] * esz, len = (m->cp_ranks[i + 1] - m->cp_ranks[i]) * esz;
            SD_TRY(sd_blk_permute_range(v, st, v->nc, 1, m->cp_keys[i], m->cp_keys[i + 1]));
            SD_CUDA(cudaEventRecord(c->ev_copy[SD_COPY_CHUNKS + i], c->stream));
            SD_CUDA(cudaStreamWaitEvent(c->d2h, c->ev_copy[SD_COPY_CHUNKS + i], 0));
            if (len) SD_CUDA(cudaMemcpyAsync((char *)host + off, (const char *)st + off, len, cudaMemcpyDeviceToHost, c->d2h));
        }
        SD_CUDA(cudaEventRecord(c->ev_copy[e + 3], c->d2h));
        c->d2h_pending = true; c->copy_pending = true;
        return SD_OK;
    }
    SD_CUDA(cudaMemcpyAsync(host, v->d, sd_vec_bytes(v), cudaMemcpyDeviceToHost, c->stream));
    return SD_OK;
}
int sd_host_alloc(void **p, uint64_t bytes) {
    SD_ARG(p, "NULL argument");
    SD_CUDA(cudaMallocHost(p, bytes ? bytes : 1));
    return SD_OK;
}
int sd_host_free(void *p) {
    if (p) SD_CUDA(cudaFreeHost(p));
    return SD_OK;
}
int sd_vec_zero(sd_vec *v) {
    SD_ARG(v, "NULL argument");
    sd_ctx *c = v->model->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_TRY(sd_before_write(c, v));
    SD_CUDA(cudaMemsetAsync(v->d, 0, sd_vec_bytes(v), c->stream));
    return SD_OK;
}
int sd_vec_set_onehot(sd_vec *v, uint64_t idx0) {
    SD_ARG(v, "NULL argument");
    SD_ARG(idx0 < v->model->N, "index outside the basis");
    sd_ctx *c = v->model->ctx;
    SD_LOCK(c);
    SD_TRY(sd_vec_zero(v));
    const uint64_t ls = v->model->shards.start[c->rank];
    if (idx0 >= ls && idx0 < ls + v->logical_n) {
        uint64_t off = idx0 - ls;
        if (v->layout) {
            const sd_model *m = v->model;
            const uint64_t st = sd_unrank_state(idx0, m->L, m->k, c->binom.data(), SD_BINOM_DIM);
            off = sd_blk_pos_of_state(m->blk.host, st, v->nc) - m->blk.pstart[c->rank];
        }
        sd_set_one_kernel<<<1, 1, 0, c->stream>>>(v->d, off * v->nc);
        SD_TRY(sd_launch_check(c, "sd_set_one_kernel"));
    }
    return SD_OK;
}
// Elements of a (possibly sharded, possibly block-layout) vector by basis rank: out[i] = v[idx0[i]] for the ranks this
// shard holds (present[i] = 1), untouched otherwise (present[i] = 0).  Host output; meant for a few sampled rows.
int sd_vec_get(sd_vec *v, const uint64_t *idx0, uint64_t count, void *out, unsigned char *present) {
    SD_ARG(v && idx0 && out && present, "NULL argument");
    SD_ARG(count <= 65536, "at most 65536 elements per call");
    sd_model *m = v->model;
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    const uint64_t ls = m->shards.start[c->rank];
    std::vector<uint64_t> off;
    std::vector<uint64_t> where;
    for (uint64_t i = 0; i < count; ++i) {
        SD_ARG(idx0[i] < m->N, "index outside the basis");
        present[i] = (idx0[i] >= ls && idx0[i] < ls + v->logical_n) ? 1 : 0;
        if (!present[i]) continue;
        uint64_t o = idx0[i] - ls;
        if (v->layout) {
            const uint64_t st = sd_unrank_state(idx0[i], m->L, m->k, c->binom.data(), SD_BINOM_DIM);
            o = sd_blk_pos_of_state(m->blk.host, st, v->nc) - m->blk.pstart[c->rank];
        }
        off.push_back(o);
        where.push_back(i);
    }
    if (off.empty()) return SD_OK;
    uint64_t *d_off = nullptr;
    double *d_out = nullptr;
    SD_CUDA(cudaMalloc(&d_off, off.size() * sizeof(uint64_t)));
    cudaError_t e = cudaMalloc(&d_out, off.size() * v->nc * sizeof(double));
    if (e != cudaSuccess) { cudaFree(d_off); return sd_fail(SD_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); }
    std::vector<double> h(off.size() * v->nc);
    e = cudaMemcpyAsync(d_off, off.data(), off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        sd_gather_kernel<<<(unsigned)((off.size() + 127) / 128), 128, 0, c->stream>>>(v->d, d_off, (unsigned)off.size(), v->nc, d_out);
        c->launches++;
        e = cudaMemcpyAsync(h.data(), d_out, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_off); cudaFree(d_out);
    if (e != cudaSuccess) return sd_fail(SD_ERR_CUDA, "sd_vec_get: %s", cudaGetErrorString(e));
    for (size_t j = 0; j < where.size(); ++j)
        for (int cc = 0; cc < v->nc; ++cc) ((double *)out)[where[j] * v->nc + cc] = h[j * v->nc + cc];
    return SD_OK;
}
int sd_vec_fill_seeded(sd_vec *v, uint64_t seed, double scale) {
    SD_ARG(v, "NULL argument");
    sd_ctx *c = v->model->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_TRY(sd_before_write(c, v));
    if (v->local_n == 0) return SD_OK;
    if (v->layout) return sd_blk_permute(v, nullptr, v->nc, 0, 1, seed, scale);
    sd_fill_seeded_kernel<<<sd_blas_grid(c, v->local_n), SD_BLAS_THREADS, 0, c->stream>>>(
        v->d, v->nc, v->model->shards.start[c->rank], v->local_n, seed, scale);
    return sd_launch_check(c, "sd_fill_seeded_kernel");
}
int sd_vec_copy(sd_vec *dst, const sd_vec *src) {
    SD_ARG(dst && src, "NULL argument");
    SD_ARG(dst->model == src->model && dst->dtype == src->dtype, "vectors differ in model or dtype");
    sd_ctx *c = dst->model->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    if (dst->d != src->d) {
        SD_TRY(sd_before_write(c, dst));
        SD_CUDA(cudaMemcpyAsync(dst->d, src->d, sd_vec_bytes(dst), cudaMemcpyDeviceToDevice, c->stream));
    }
    return SD_OK;
}
int sd_vec_convert(sd_vec *dst, const sd_vec *src) {
    SD_ARG(dst && src, "NULL argument");
    SD_ARG(dst->model == src->model, "vectors belong to different models");
    if (dst->dtype == src->dtype) return sd_vec_copy(dst, src);
    sd_ctx *c = dst->model->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_TRY(sd_before_write(c, dst));
    if (dst->local_n == 0) return SD_OK;
    if (dst->layout || src->layout) {                // f64 and c128 block layouts order a class differently
        SD_ARG(dst->layout && src->layout, "vectors differ in layout");
        double *st = nullptr;
        SD_TRY(sd_scratch(c, 0, sd_vec_logical_bytes(src) + 16, &st));
        SD_TRY(sd_blk_permute(src, st, src->nc, 1, 0, 0, 0.0));
        SD_TRY(sd_blk_permute(dst, st, src->nc, 0, 0, 0, 0.0));
        sd_scratch_release(c);
        return SD_OK;
    }
    sd_convert_kernel<<<sd_blas_grid(c, dst->local_n), SD_BLAS_THREADS, 0, c->stream>>>(dst->d, src->d, dst->local_n, src->nc);
    return sd_launch_check(c, "sd_convert_kernel");
}

static SdScalar sd_host_scalar(double re, double im) {
    SdScalar s; s.re = re; s.im = im; s.dev = nullptr; s.dev_mode = 0; return s;
}
static SdScalar sd_dev_scalar(const double *p, int mode) {
    SdScalar s; s.re = 0; s.im = 0; s.dev = p; s.dev_mode = mode; return s;
}

static int sd_scale_impl(sd_vec *x, SdScalar s) {
    sd_ctx *c = x->model->ctx;
    SD_TRY(sd_before_write(c, x));
    if (x->local_n == 0) return SD_OK;
    const unsigned g = sd_blas_grid(c, x->local_n);
    if (x->nc == 2) sd_scale_kernel<2><<<g, SD_BLAS_THREADS, 0, c->stream>>>(x->d, x->local_n, s);
    else sd_scale_kernel<1><<<g, SD_BLAS_THREADS, 0, c->stream>>>(x->d, x->local_n, s);
    return sd_launch_check(c, "sd_scale_kernel");
}
int sd_vec_scale(sd_vec *x, sd_complex s) {
    SD_ARG(x, "NULL argument");
    SD_ARG(x->nc == 2 || s.im == 0.0, "complex scale of a real vector (InexactError)");
    SD_LOCK(x->model->ctx); SD_TRY(sd_use(x->model->ctx));
    return sd_scale_impl(x, sd_host_scalar(s.re, s.im));
}
// y = x / s (s real; device or host scalar)
static int sd_divide_impl(sd_vec *y, const sd_vec *x, SdScalar s) {
    sd_ctx *c = y->model->ctx;
    SD_TRY(sd_before_write(c, y));
    if (y->local_n == 0) return SD_OK;
    const unsigned g = sd_blas_grid(c, y->local_n * y->nc);
    if (y->nc == 2) sd_divide_kernel<2><<<g, SD_BLAS_THREADS, 0, c->stream>>>(y->d, x->d, y->local_n, s);
    else sd_divide_kernel<1><<<g, SD_BLAS_THREADS, 0, c->stream>>>(y->d, x->d, y->local_n, s);
    return sd_launch_check(c, "sd_divide_kernel");
}
// y += a x [+ b z]; if slot_out >= 0: ||y||^2 -> d_scal[slot_out+3]
static int sd_axpy_impl(sd_vec *y, SdScalar a, const sd_vec *x, SdScalar b, const sd_vec *z, int slot_out) {
    sd_ctx *c = y->model->ctx;
    SD_TRY(sd_before_write(c, y));
    const unsigned g = sd_blas_grid(c, y->local_n);
    double *partials = nullptr;
    if (slot_out >= 0) { SD_TRY(sd_partials_reserve(c, (size_t)SD_NSLOT * g)); partials = c->d_partials; }
    if (y->nc == 2)
        sd_axpy_kernel<2><<<g, SD_BLAS_THREADS, 0, c->stream>>>(y->d, y->local_n, a, x->d, b, z ? z->d : nullptr, partials, g);
    else
        sd_axpy_kernel<1><<<g, SD_BLAS_THREADS, 0, c->stream>>>(y->d, y->local_n, a, x->d, b, z ? z->d : nullptr, partials, g);
    SD_TRY(sd_launch_check(c, "sd_axpy_kernel"));
    if (slot_out >= 0) SD_TRY(sd_finish_reduce(c, g, 8, slot_out));
    return SD_OK;
}
int sd_vec_axpy(sd_vec *y, sd_complex a, const sd_vec *x) {
    SD_ARG(y && x, "NULL argument");
    SD_ARG(y->model == x->model && y->dtype == x->dtype, "vectors differ in model or dtype");
    SD_ARG(y->nc == 2 || a.im == 0.0, "complex axpy into a real vector (InexactError)");
    SD_LOCK(y->model->ctx); SD_TRY(sd_use(y->model->ctx));
    return sd_axpy_impl(y, sd_host_scalar(a.re, a.im), x, sd_host_scalar(0, 0), nullptr, -1);
}
// dot -> d_scal[slot_out + 0,1]
static int sd_dot_impl(const sd_vec *x, const sd_vec *y, int conj, int slot_out) {
    sd_ctx *c = x->model->ctx;
    const unsigned g = sd_blas_grid(c, x->local_n);
    SD_TRY(sd_partials_reserve(c, (size_t)SD_NSLOT * g));
    if (x->nc == 2) sd_dot_kernel<2><<<g, SD_BLAS_THREADS, 0, c->stream>>>(x->d, y->d, x->local_n, conj, c->d_partials, g);
    else sd_dot_kernel<1><<<g, SD_BLAS_THREADS, 0, c->stream>>>(x->d, y->d, x->local_n, conj, c->d_partials, g);
    SD_TRY(sd_launch_check(c, "sd_dot_kernel"));
    return sd_finish_reduce(c, g, 3, slot_out);
}
int sd_vec_dot(const sd_vec *x, const sd_vec *y, sd_complex *result) {
    SD_ARG(x && y && result, "NULL argument");
    SD_ARG(x->model == y->model && x->dtype == y->dtype, "vectors differ in model or dtype");
    sd_ctx *c = x->model->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_TRY(sd_dot_impl(x, y, 1, 0));
    double r[2];
    SD_TRY(sd_fetch(c, 0, 2, r));
    result->re = r[0]; result->im = r[1];
    return SD_OK;
}
int sd_vec_dotu(const sd_vec *x, const sd_vec *y, sd_complex *result) {
    SD_ARG(x && y && result, "NULL argument");
    SD_ARG(x->model == y->model && x->dtype == y->dtype, "vectors differ in model or dtype");
    sd_ctx *c = x->model->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_TRY(sd_dot_impl(x, y, 0, 0));
    double r[2];
    SD_TRY(sd_fetch(c, 0, 2, r));
    result->re = r[0]; result->im = r[1];
    return SD_OK;
}
int sd_vec_norm(const sd_vec *x, double *result) {
    SD_ARG(x && result, "NULL argument");
    sd_ctx *c = x->model->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_TRY(sd_dot_impl(x, x, 1, 0));
    double r[2];
    SD_TRY(sd_fetch(c, 0, 2, r));
    *result = sqrt(r[0]);
    return SD_OK;
}

// ----------------------------------------------------------------- operator
// A rank whose shard holds no tile (fewer tiles than ranks, tiny models only) launches nothing but must
// still contribute zeros to the cross-rank sum of the fused reductions: every rank calls the same
// sequence of NCCL collectives.
static int sd_empty_shard_reduce(sd_ctx *c, int slotmask, int slot_out) {
    if (!slotmask) return SD_OK;
    SD_CUDA(cudaMemsetAsync(c->d_scal + slot_out, 0, SD_NSLOT * sizeof(double), c->stream));
    if (c->world > 1) {
        SD_NCCL(g_nccl.AllReduce(c->d_scal + slot_out, c->d_scal + slot_out, SD_NSLOT, ncclFloat64_, ncclSum_,
                                 c->comm, c->stream));
        sd_collective_done(c);
    }
    return SD_OK;
}
// one launch of the block-layout apply over the tile keys [P.key_lo, P.key_hi) (or P.order)
static int sd_blk_launch_range(sd_model *m, int nc, const SdBlkParams &P, const SdVecView &view, double *out_local,
                               const SdEpi &epi, bool plain) {
    sd_ctx *c = m->ctx;
    const uint64_t nkeys = P.key_hi - P.key_lo;
    if (nkeys == 0) return SD_OK;
    const unsigned grid = (unsigned)std::min<uint64_t>(nkeys, (uint64_t)c->sm_count);
    SD_CUDA(cudaMemsetAsync(c->d_tilectr, 0, 16 * sizeof(unsigned long long), c->stream));
    const int qfar = m->blk.qfar[nc - 1];
    const size_t smem = m->blk.smem[nc - 1];
    // the shared-memory attribute is per device and cheap to set: no process-global "already set" cache
#define SD_HL(KERNEL_, THREADS_)                                                                             \
    do {                                                                                                     \
        SD_CUDA(cudaFuncSetAttribute(KERNEL_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        KERNEL_<<<grid, THREADS_, smem, c->stream>>>(P, view, out_local, epi, qfar, c->d_tilectr);           \
    } while (0)
#define SD_HL_LEAN(NC_, EK_)                                                                                 \
    do {                                                                                                     \
        if (P.wrap_on) SD_HL((sd_blkl_apply_kernel<NC_, EK_, 640, true>), 640);   /* periodic chain: WRAP variant */ \
        else if (m->blk.threads == 512) SD_HL((sd_blkl_apply_kernel<NC_, EK_, 512>), 512);                   \
        else if (m->blk.threads == 768) SD_HL((sd_blkl_apply_kernel<NC_, EK_, 768>), 768);                   \
        else SD_HL((sd_blkl_apply_kernel<NC_, EK_, 640>), 640);                                              \
    } while (0)
    // epilogue kind (sd_blkl.h): 0 plain, 1 Lanczos (hscale + fused <psi, out>), 2 generic
    const int ek = plain ? 0 : ((epi.mode == SD_EPI_PLAIN && (epi.red == SD_RED_DOT_SELF || epi.red == 0) && !epi.acc) ? 1 : 2);
    if (nc == 1) { if (ek == 0) SD_HL_LEAN(1, 0); else if (ek == 1) SD_HL_LEAN(1, 1); else SD_HL_LEAN(1, 2); }
    else { if (ek == 0) SD_HL_LEAN(2, 0); else if (ek == 1) SD_HL_LEAN(2, 1); else SD_HL_LEAN(2, 2); }
#undef SD_HL_LEAN
#undef SD_HL
    return sd_launch_check(c, "sd_blkl_apply_kernel");
}
// Launches one apply kernel with the given epilogue; reductions (if any) land
// in d_scal[slot_out .. slot_out+3].
static int sd_apply_impl(sd_model *m, sd_vec *out, const sd_vec *psi, SdEpi epi, int slot_out, const sd_vec *acc = nullptr) {
    sd_ctx *c = m->ctx;
    SD_ARG(out && psi, "NULL argument");
    SD_ARG(out->model == m && psi->model == m, "vector does not belong to this model");
    SD_ARG(out->dtype == psi->dtype, "out and psi differ in element type");
    SD_ARG(out->d != psi->d, "out must not alias psi");
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_TRY(sd_before_apply(c, out, psi, acc));
    const int nc = psi->nc;
    const int slotmask = sd_epi_slotmask(epi.red);
    SD_ARG(out->layout == psi->layout && psi->layout == (m->path == SD_PATH_BLOCK ? 1 : 0),
           "vector layout does not match the model's kernel path");
    if (m->path == SD_PATH_BLOCK) {
        SdBlkParams P = sd_blk_params(m, nc);
        const uint64_t nkeys = P.key_hi - P.key_lo;
        if (nkeys == 0) return sd_empty_shard_reduce(c, slotmask, slot_out);
        SD_ARG(nkeys < 0x7fffffffULL, "too many tiles for one launch");
        const unsigned grid = (unsigned)std::min<uint64_t>(nkeys, (uint64_t)c->sm_count);
        if (slotmask) {
            SD_TRY(sd_partials_reserve(c, (size_t)SD_NSLOT * nkeys));
            SD_CUDA(cudaMemsetAsync(c->d_partials, 0, (size_t)SD_NSLOT * nkeys * sizeof(double), c->stream));
        }
        epi.partials = c->d_partials;
        epi.nparts = (unsigned)nkeys;
        const bool plain = epi.mode == SD_EPI_PLAIN && epi.red == 0 && !epi.acc && epi.hscale == 1.0 && !epi.hscale_dev;
        SD_TRY(sd_blk_launch_range(m, nc, P, psi->view, out->d, epi, plain));
        if (slotmask) SD_TRY(sd_finish_reduce(c, (unsigned)nkeys, slotmask, slot_out));
    } else if (m->path == SD_PATH_TILED) {
        SdTileDev &t = m->tile[nc - 1];
        SdTileParams P = t.host.P;
        P.shards = m->shards;
        P.key_lo = t.keys[c->rank];
        P.key_hi = t.keys[c->rank + 1];
        const uint64_t nkeys = P.key_hi - P.key_lo;
        if (nkeys == 0) return sd_empty_shard_reduce(c, slotmask, slot_out);
        SD_ARG(nkeys < 0x7fffffffULL, "too many tiles for one launch");
        const unsigned grid = (unsigned)nkeys;           // one CTA per tile, in rank order
        if (slotmask) SD_TRY(sd_partials_reserve(c, (size_t)SD_NSLOT * grid));
        epi.partials = c->d_partials;
        epi.nparts = grid;
        double *out_vbase = out->d - (int64_t)m->shards.start[c->rank] * nc;
        const int T = m->tile_T[nc - 1];
#define SD_LAUNCH_TILE3(NC_, T_, PLAIN_, NTHR_)                                                             \
    do {                                                                                                    \
        /* per device and cheap: set on every launch, no process-global cache (several contexts / GPUs per process) */ \
        SD_CUDA(cudaFuncSetAttribute(sd_tile_apply_kernel<NC_, T_, PLAIN_, NTHR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem)); \
        sd_tile_apply_kernel<NC_, T_, PLAIN_, NTHR_><<<grid, NTHR_, t.smem, c->stream>>>(P, psi->view, out_vbase, epi, t.cap); \
    } while (0)
#define SD_LAUNCH_TILE2(NC_, T_, PLAIN_)                                                                    \
    do {                                                                                                    \
        if (m->tile_threads == 256) SD_LAUNCH_TILE3(NC_, T_, PLAIN_, 256);                                  \
        else SD_LAUNCH_TILE3(NC_, T_, PLAIN_, 512);                                                         \
    } while (0)
#define SD_LAUNCH_TILE(NC_, T_)                                                                             \
    do {                                                                                                    \
        if (plain) SD_LAUNCH_TILE2(NC_, T_, true);                                                          \
        else SD_LAUNCH_TILE2(NC_, T_, false);                                                               \
    } while (0)
        const bool plain = epi.mode == SD_EPI_PLAIN && epi.red == 0 && !epi.acc && epi.hscale == 1.0 && !epi.hscale_dev;
        if (nc == 1 && T == 5) SD_LAUNCH_TILE(1, 5);
        else if (nc == 2 && T == 4) SD_LAUNCH_TILE(2, 4);
        else if (nc == 1 && T == 4) SD_LAUNCH_TILE(1, 4);
        else if (nc == 1 && T == 6) SD_LAUNCH_TILE(1, 6);
        else if (nc == 2 && T == 3) SD_LAUNCH_TILE(2, 3);
        else if (nc == 2 && T == 5) SD_LAUNCH_TILE(2, 5);
        else return sd_fail(SD_ERR_UNSUPPORTED, "tail size %d not compiled", T);
#undef SD_LAUNCH_TILE3
#undef SD_LAUNCH_TILE2
#undef SD_LAUNCH_TILE
        SD_TRY(sd_launch_check(c, "sd_tile_apply_kernel"));
        if (slotmask) SD_TRY(sd_finish_reduce(c, grid, slotmask, slot_out));
    } else {
        SdGenericParams G;
        G.L = m->L; G.k = m->k; G.nhop = (int)m->hop_a.size(); G.nzz = (int)m->zz_a.size(); G.N = m->N;
        G.hop_a = m->d_hop_a; G.hop_b = m->d_hop_b; G.hop_J = m->d_hop_J;
        G.zz_a = m->d_zz_a; G.zz_b = m->d_zz_b; G.zz_J = m->d_zz_J; G.field = m->d_field;
        G.binom = c->d_binom; G.lin_h = m->lin_h; G.linA = m->d_linA; G.linB = m->d_linB;
        G.shards = m->shards;
        uint64_t g64 = (psi->logical_n + SD_GEN_THREADS - 1) / SD_GEN_THREADS;
        g64 = std::max<uint64_t>(1, std::min<uint64_t>(g64, (uint64_t)c->sm_count * 16));
        const unsigned g = (unsigned)g64;
        if (slotmask) SD_TRY(sd_partials_reserve(c, (size_t)SD_NSLOT * g));
        epi.partials = c->d_partials;
        epi.nparts = g;
        if (nc == 2) sd_generic_apply_kernel<2><<<g, SD_GEN_THREADS, 0, c->stream>>>(G, psi->view, out->d, epi);
        else sd_generic_apply_kernel<1><<<g, SD_GEN_THREADS, 0, c->stream>>>(G, psi->view, out->d, epi);
        SD_TRY(sd_launch_check(c, "sd_generic_apply_kernel"));
        if (slotmask) SD_TRY(sd_finish_reduce(c, g, slotmask, slot_out));
    }
    return SD_OK;
}
static SdEpi sd_epi_plain(double hscale) {
    SdEpi e;
    memset(&e, 0, sizeof(e));
    e.mode = SD_EPI_PLAIN; e.hscale = hscale; e.a = 1.0; e.b = 0.0;
    return e;
}

int sd_apply_H(sd_model *m, sd_vec *out, const sd_vec *psi) {
    SD_ARG(m, "NULL argument");
    return sd_apply_impl(m, out, psi, sd_epi_plain(1.0), 0);
}
int sd_apply_H_dot(sd_model *m, sd_vec *out, const sd_vec *psi, sd_complex *dot) {
    SD_ARG(m && dot, "NULL argument");
    SdEpi e = sd_epi_plain(1.0);
    e.red = SD_RED_DOT_SELF;
    SD_TRY(sd_apply_impl(m, out, psi, e, 0));
    double r[2];
    SD_TRY(sd_fetch(m->ctx, 0, 2, r));
    dot->re = r[0]; dot->im = r[1];
    return SD_OK;
}
int sd_apply_rescaled_H(sd_model *m, sd_vec *out, const sd_vec *psi, double a, double b) {
    SD_ARG(m, "NULL argument");
    SdEpi e = sd_epi_plain(1.0);
    e.mode = SD_EPI_RESCALED; e.a = a; e.b = b;
    return sd_apply_impl(m, out, psi, e, 0);
}
static int sd_cheb_step_impl(sd_model *m, sd_vec *vnext, const sd_vec *v, const sd_vec *vprev, double a, double b,
                             const sd_vec *phi, sd_vec *acc, sd_complex ck, int slot_out) {
    SD_ARG(vnext && v && vprev, "NULL argument");
    SD_ARG(vprev->model == m && vprev->dtype == v->dtype, "vprev mismatch");
    SdEpi e = sd_epi_plain(1.0);
    e.mode = SD_EPI_CHEB; e.a = a; e.b = b; e.vprev = vprev->d;
    if (phi) {
        SD_ARG(phi->model == m && phi->dtype == v->dtype, "phi mismatch");
        e.red = SD_RED_DOT_PHI | SD_RED_NORM2; e.phi = phi->d;
    }
    if (acc) {
        SD_ARG(acc->model == m && acc->dtype == v->dtype, "acc mismatch");
        SD_ARG(acc->nc == 2 || ck.im == 0.0, "complex coefficient into a real accumulator (InexactError)");
        SD_ARG(acc->d != vnext->d && acc->d != v->d, "acc must not alias vnext or v");
        e.acc = acc->d; e.ck_re = ck.re; e.ck_im = ck.im;
    }
    return sd_apply_impl(m, vnext, v, e, slot_out, acc);
}
int sd_cheb_step(sd_model *m, sd_vec *vnext, const sd_vec *v, const sd_vec *vprev, double a, double b,
                 const sd_vec *phi, double *mu, double *norm2, sd_vec *acc, sd_complex ck) {
    SD_ARG(m, "NULL argument");
    SD_TRY(sd_cheb_step_impl(m, vnext, v, vprev, a, b, phi, acc, ck, 0));
    if (phi && (mu || norm2)) {
        double r[4];
        SD_TRY(sd_fetch(m->ctx, 0, 4, r));
        if (mu) *mu = r[2];
        if (norm2) *norm2 = r[3];
    }
    return SD_OK;
}
// phi = (sum_r w_r s_r(state)) * ComplexF64(psi0) with complex per-site weights w[L]: Sz_q_vector is w_r = e^{iqr} / sqrt(L)
// (Hamiltonian.jl:307-337), the single-site S^z_i of the site-resolved KPM (TimeEvolution/KPM.jl:197-213) is w = delta_{r,i}.
static int sd_sz_weights_impl(sd_model *m, sd_vec *phi, const sd_vec *psi0, const double *wre, const double *wim, double normfact, double *norm2);
int sd_apply_sz_weights(sd_model *m, sd_vec *phi, const sd_vec *psi0, const sd_complex *w, double *norm2) {
    SD_ARG(m && phi && psi0 && w, "NULL argument");
    double wre[SD_MAX_L + 1], wim[SD_MAX_L + 1];
    for (int r = 0; r <= SD_MAX_L; ++r) { wre[r] = r < m->L ? w[r].re : 0.0; wim[r] = r < m->L ? w[r].im : 0.0; }
    return sd_sz_weights_impl(m, phi, psi0, wre, wim, 1.0, norm2);
}
int sd_szq(sd_model *m, sd_vec *phi, const sd_vec *psi0, double q, double *norm2) {
    SD_ARG(m && phi && psi0, "NULL argument");
    double wre[SD_MAX_L + 1], wim[SD_MAX_L + 1];
    for (int r = 0; r <= SD_MAX_L; ++r) { wre[r] = r < m->L ? cos(q * (double)r) : 0.0; wim[r] = r < m->L ? sin(q * (double)r) : 0.0; }
    return sd_sz_weights_impl(m, phi, psi0, wre, wim, 1.0 / sqrt((double)m->L), norm2);
}
static int sd_sz_weights_impl(sd_model *m, sd_vec *phi, const sd_vec *psi0, const double *wre, const double *wim, double normfact, double *norm2) {
    SD_ARG(phi->model == m && psi0->model == m, "vector does not belong to this model");
    SD_ARG(phi->dtype == SD_C128, "phi must be SD_C128");
    SD_ARG(phi->d != psi0->d, "phi must not alias psi0");
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_TRY(sd_before_write(c, phi));
    if (phi->layout) {                               // block layout: the state of an element is known from its position (sd_blkv.h)
        SD_ARG(psi0->layout, "vectors differ in layout");
        SdBlkParams P = sd_blk_params(m, 2);
        SdBlkSzq ZB;
        ZB.normfact = normfact;
        for (int r = 0; r <= SD_MAX_L; ++r) { ZB.ph_re[r] = wre[r]; ZB.ph_im[r] = wim[r]; }
        const uint64_t nkeys = P.key_hi - P.key_lo;
        const unsigned g = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(nkeys, (uint64_t)c->sm_count * 16));
        double *partials = nullptr;
        if (norm2) {
            SD_TRY(sd_partials_reserve(c, (size_t)SD_NSLOT * g));
            partials = c->d_partials;
            if (nkeys == 0) SD_CUDA(cudaMemsetAsync(partials, 0, (size_t)SD_NSLOT * g * sizeof(double), c->stream));
        }
        if (nkeys > 0) {
            if (psi0->nc == 2) sd_blk_szq_kernel<2><<<g, 256, 0, c->stream>>>(P, ZB, psi0->d, phi->d, partials, g);
            else sd_blk_szq_kernel<1><<<g, 256, 0, c->stream>>>(P, ZB, psi0->d, phi->d, partials, g);
            SD_TRY(sd_launch_check(c, "sd_blk_szq_kernel"));
        }
        if (norm2) {
            SD_TRY(sd_finish_reduce(c, g, 8, 0));
            double r[4];
            SD_TRY(sd_fetch(c, 0, 4, r));
            *norm2 = r[3];
        }
        return SD_OK;
    }
    SdSzqParams Z;
    Z.L = m->L; Z.k = m->k; Z.normfact = normfact; Z.binom = c->d_binom;
    for (int r = 0; r < m->L; ++r) { Z.ph_re[r] = wre[r]; Z.ph_im[r] = wim[r]; }
    const unsigned g = sd_blas_grid(c, phi->logical_n);
    double *partials = nullptr;
    if (norm2) { SD_TRY(sd_partials_reserve(c, (size_t)SD_NSLOT * g)); partials = c->d_partials; }
    const uint64_t ls = m->shards.start[c->rank];
    if (phi->logical_n > 0) {
        if (psi0->nc == 2) sd_szq_kernel<2><<<g, SD_BLAS_THREADS, 0, c->stream>>>(Z, ls, phi->logical_n, psi0->d, phi->d, partials, g);
        else sd_szq_kernel<1><<<g, SD_BLAS_THREADS, 0, c->stream>>>(Z, ls, phi->logical_n, psi0->d, phi->d, partials, g);
        SD_TRY(sd_launch_check(c, "sd_szq_kernel"));
    } else if (partials) {                           // empty shard: contribute zeros to the cross-rank sum
        SD_CUDA(cudaMemsetAsync(partials, 0, (size_t)SD_NSLOT * g * sizeof(double), c->stream));
    }
    if (norm2) {
        SD_TRY(sd_finish_reduce(c, g, 8, 0));
        double r[4];
        SD_TRY(sd_fetch(c, 0, 4, r));
        *norm2 = r[3];
    }
    return SD_OK;
}
// Observables.jl:14-109 on a device-resident vector: mags[L], zz[L] (see sd_obs.h); summed over ranks.
#define SD_OBS_SLOT 3584                                               // d_scal[3584 .. 3711] (sd_lincomb stages coefficients at 1024 .. 3071)
int sd_vec_observables(const sd_vec *psi, double *mags, double *zz) {
    SD_ARG(psi && mags && zz, "NULL argument");
    sd_model *m = psi->model;
    sd_ctx *c = m->ctx;
    SD_ARG(m->L <= 63, "L must be at most 63");
    SD_LOCK(c); SD_TRY(sd_use(c));
    unsigned nwarps;
    if (psi->layout) {                               // block layout: states from the element positions (sd_blkv.h)
        SdBlkParams P = sd_blk_params(m, psi->nc);
        const uint64_t nkeys = P.key_hi - P.key_lo;
        const unsigned g = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(nkeys, (uint64_t)c->sm_count * 8));
        nwarps = g * (256 / 32);
        SD_TRY(sd_partials_reserve(c, (size_t)nwarps * 128));
        if (psi->nc == 2) sd_blk_obs_kernel<2><<<g, 256, 0, c->stream>>>(P, psi->d, c->d_partials);
        else sd_blk_obs_kernel<1><<<g, 256, 0, c->stream>>>(P, psi->d, c->d_partials);
        SD_TRY(sd_launch_check(c, "sd_blk_obs_kernel"));
    } else {
        const unsigned g = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((psi->logical_n + 255) / 256, (uint64_t)c->sm_count * 4));
        nwarps = g * (256 / 32);
        SD_TRY(sd_partials_reserve(c, (size_t)nwarps * 128));
        const uint64_t ls = m->shards.start[c->rank];
        if (psi->nc == 2) sd_obs_kernel<2><<<g, 256, 0, c->stream>>>(m->L, m->k, c->d_binom, ls, psi->logical_n, psi->d, c->d_partials);
        else sd_obs_kernel<1><<<g, 256, 0, c->stream>>>(m->L, m->k, c->d_binom, ls, psi->logical_n, psi->d, c->d_partials);
        SD_TRY(sd_launch_check(c, "sd_obs_kernel"));
    }
    sd_obs_reduce_kernel<<<1, 128, 0, c->stream>>>(c->d_partials, nwarps, c->d_scal + SD_OBS_SLOT);
    SD_TRY(sd_launch_check(c, "sd_obs_reduce_kernel"));
    if (c->world > 1)
        { SD_NCCL(g_nccl.AllReduce(c->d_scal + SD_OBS_SLOT, c->d_scal + SD_OBS_SLOT, 128, ncclFloat64_, ncclSum_, c->comm, c->stream)); sd_collective_done(c); }
    double r[128];
    SD_TRY(sd_fetch(c, SD_OBS_SLOT, 128, r));
    for (int i = 0; i < m->L; ++i) { mags[i] = r[i]; zz[i] = r[64 + i]; }
    return SD_OK;
}
int sd_apply_H_host(sd_model *m, int dtype, void *out, const void *psi) {
    SD_ARG(m && out && psi, "NULL argument");
    SD_ARG(m->ctx->world == 1, "sd_apply_H_host needs a single-rank context");
    SD_ARG(out != psi, "out must not alias psi");
    sd_vec *vi = nullptr, *vo = nullptr;
    int rc = sd_vec_alloc(m, dtype, &vi);
    if (rc == SD_OK) rc = sd_vec_alloc(m, dtype, &vo);
    if (rc == SD_OK) rc = sd_vec_upload_async(vi, psi);
    if (rc == SD_OK) rc = sd_apply_H(m, vo, vi);
    if (rc == SD_OK) rc = sd_vec_download_async(vo, out);
    if (rc == SD_OK) rc = sd_ctx_sync(m->ctx);
    sd_vec_free(vi); sd_vec_free(vo);
    return rc;
}

// -------------------------------------------------------------- recurrences
int sd_vecset_free(sd_vecset *s) {
    if (!s) return SD_OK;
    for (sd_vec *v : s->v) sd_vec_free(v);
    delete s;
    return SD_OK;
}
int sd_vecset_size(const sd_vecset *s, int *m) {
    SD_ARG(s && m, "NULL argument");
    *m = (int)s->v.size();
    return SD_OK;
}
int sd_vecset_get(sd_vecset *s, int k, sd_vec **vec) {
    SD_ARG(s && vec, "NULL argument");
    SD_ARG(k >= 0 && k < (int)s->v.size(), "index outside the vector set");
    *vec = s->v[k];
    return SD_OK;
}
int sd_lincomb(sd_vecset *s, const sd_complex *y, int mcount, sd_vec *out, double *norm2) {
    SD_ARG(s && y && out, "NULL argument");
    SD_ARG(mcount >= 1 && mcount <= (int)s->v.size(), "m outside the vector set");
    sd_vec *v0 = s->v[0];
    sd_model *m = v0->model;
    SD_ARG(out->model == m, "out belongs to a different model");
    SD_ARG(out->nc >= v0->nc, "cannot combine complex vectors into a real output");
    for (int j = 0; j < mcount; ++j) {
        SD_ARG(s->v[j]->d != out->d, "out aliases a member of the set");
        if (out->nc == 1) SD_ARG(y[j].im == 0.0, "complex coefficient into a real output (InexactError)");
    }
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    SD_ARG(2 * mcount <= 2048, "too many vectors");
    if (out->layout && out->nc == 2 && v0->nc == 1) {
        // the f64 and c128 block layouts order a class differently (pair rows): combine through a
        // converted copy of each member (Krylov recombination of a real basis, once per solve)
        sd_vec *tmp = nullptr;
        SD_TRY(sd_vec_alloc(m, SD_C128, &tmp));
        int rc = sd_vec_zero(out);
        for (int j = 0; j < mcount && rc == SD_OK; ++j) {
            rc = sd_vec_convert(tmp, s->v[j]);
            if (rc == SD_OK) rc = sd_vec_axpy(out, y[j], tmp);
        }
        sd_vec_free(tmp);
        SD_TRY(rc);
        if (norm2) {
            double nrm = 0.0;
            SD_TRY(sd_vec_norm(out, &nrm));
            *norm2 = nrm * nrm;
        }
        return SD_OK;
    }
    SD_TRY(sd_before_write(c, out));
    // coefficients -> device (pinned staging, ordered on the stream)
    SD_CUDA(cudaStreamSynchronize(c->stream));
    for (int j = 0; j < mcount; ++j) { c->h_scal[1024 + 2 * j] = y[j].re; c->h_scal[1024 + 2 * j + 1] = y[j].im; }
    SD_CUDA(cudaMemcpyAsync(c->d_scal + 1024, c->h_scal + 1024, 2 * mcount * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const unsigned g = sd_blas_grid(c, out->local_n);
    SD_TRY(sd_partials_reserve(c, (size_t)SD_NSLOT * g));
    for (int j0 = 0; j0 < mcount; j0 += SD_BDOT_MAX) {
        const int nb = std::min(SD_BDOT_MAX, mcount - j0);
        SdPtrBlock pb;
        for (int j = 0; j < SD_BDOT_MAX; ++j) pb.v[j] = (j < nb) ? s->v[j0 + j]->d : nullptr;
        const bool last = j0 + nb >= mcount;
        double *partials = last ? c->d_partials : nullptr;
        const double *yd = c->d_scal + 1024 + 2 * j0;
        if (out->nc == 2 && v0->nc == 2)
            sd_lincomb_kernel<2><<<g, SD_BLAS_THREADS, 0, c->stream>>>(out->d, out->local_n, pb, nb, yd, j0 > 0, partials, g);
        else if (out->nc == 1)
            sd_lincomb_kernel<1><<<g, SD_BLAS_THREADS, 0, c->stream>>>(out->d, out->local_n, pb, nb, yd, j0 > 0, partials, g);
        else
            sd_lincomb_r2c_kernel<<<g, SD_BLAS_THREADS, 0, c->stream>>>(out->d, out->local_n, pb, nb, yd, j0 > 0, partials, g);
        SD_TRY(sd_launch_check(c, "sd_lincomb_kernel"));
    }
    SD_TRY(sd_finish_reduce(c, g, 8, 0));
    if (norm2) {
        double r[4];
        SD_TRY(sd_fetch(c, 0, 4, r));
        *norm2 = r[3];
    }
    return SD_OK;
}

// Work vectors of the recurrences.  They come from a small per-model pool and go back to it on every exit path: a
// solver call at L = 32 otherwise spends more time in cudaMalloc / cudaFree of its 4.8 GB temporaries (and, sharded, in
// the collective handle exchange of sd_vec_alloc) than in its kernels.  The pool holds at most SD_POOL_MAX vectors, is
// emptied by sd_model_free and by an allocation that runs out of memory.  Pooled vectors do not count as live.
#define SD_POOL_MAX 6
struct SdVecGuard {
    std::vector<sd_vec *> v;
    ~SdVecGuard() {
        for (sd_vec *p : v) {
            sd_model *m = p->model;
            if ((int)m->pool.size() < SD_POOL_MAX) { m->pool.push_back(p); m->live_vecs--; }
            else sd_vec_free(p);
        }
    }
    int make(sd_model *m, int dtype, sd_vec **out) {
        for (size_t i = 0; i < m->pool.size(); ++i)
            if (m->pool[i]->dtype == dtype) {
                *out = m->pool[i];
                m->pool.erase(m->pool.begin() + i);
                m->live_vecs++;
                v.push_back(*out);
                return SD_OK;
            }
        int rc = sd_vec_alloc(m, dtype, out);
        if (rc == SD_OK) v.push_back(*out);
        return rc;
    }
    void release(sd_vec *p) { v.erase(std::remove(v.begin(), v.end(), p), v.end()); }
};

// v <- v0 / ||v0||; returns the norm
static int sd_normalised_copy(sd_vec *dst, const sd_vec *src, double *norm_out) {
    sd_ctx *c = dst->model->ctx;
    SD_TRY(sd_dot_impl(src, src, 1, 0));
    double r[2];
    SD_TRY(sd_fetch(c, 0, 2, r));
    const double nrm = sqrt(r[0]);
    if (norm_out) *norm_out = nrm;
    if (nrm == 0.0) return SD_OK;
    return sd_divide_impl(dst, src, sd_host_scalar(nrm, 0));
}

// One fused 3R+1W pass after an apply (sd_lanczos_update_kernel); ||w||^2 -> d_scal[slot_out + 3] (+ NCCL sum).
static int sd_lanczos_update(sd_vec *w, const sd_vec *u, const sd_vec *uo, sd_vec *out, const SdLanczosScal &S, int slot_out) {
    sd_ctx *c = w->model->ctx;
    SD_TRY(sd_before_write(c, w));
    if (out) SD_TRY(sd_before_write(c, out));
    const unsigned g = sd_blas_grid(c, w->local_n);
    SD_TRY(sd_partials_reserve(c, (size_t)SD_NSLOT * g));
    if (w->nc == 2)
        sd_lanczos_update_kernel<2><<<g, SD_BLAS_THREADS, 0, c->stream>>>(w->d, u->d, uo ? uo->d : nullptr, out ? out->d : nullptr, w->local_n, S, c->d_partials, g);
    else
        sd_lanczos_update_kernel<1><<<g, SD_BLAS_THREADS, 0, c->stream>>>(w->d, u->d, uo ? uo->d : nullptr, out ? out->d : nullptr, w->local_n, S, c->d_partials, g);
    SD_TRY(sd_launch_check(c, "sd_lanczos_update_kernel"));
    return sd_finish_reduce(c, g, 8, slot_out);
}

// The three-term Lanczos recurrence of Lanczos.jl:27-84 (extremal), :196-246 (tridiag) and of the memory-lean ground
// state, on three work vectors with DEFERRED NORMALISATION and no host synchronisation inside the recurrence.  The
// vectors are kept unnormalised, u_1 = v0, u_{j+1} = w_j, v_j = u_j / beta_{j-1}, beta_0 = ||v0||; a step is two kernels:
//     apply   w = (hsign / beta_{j-1}) H u_j  with the fused dot  d_j = <u_j, w>         (1 / beta from device memory)
//     update  w -= (alpha_j / beta_{j-1}) u_j + (beta_{j-1} / beta_{j-2}) u_{j-1},  n_j = ||w||^2   (3R + 1W, fused norm)
// with alpha_j = Re d_j / beta_{j-1}, beta_j = sqrt(n_j); the last step (j = mm) is the apply alone.  The reductions of
// step j stay on the device in d_scal[SD_HIST + 4 j ..] and are fetched every 16 steps; beta_j < tol then truncates
// (m_eff = j), which returns exactly what the step-by-step test of the reference returns.
// Pass 2 (y != NULL, f64 ground state only): alpha / beta are INPUTS; the recurrence is regenerated from them -- every
// coefficient is an IEEE sqrt / division of the same numbers, so the vectors are bit-identical to pass 1 -- and
// out = sum_j y_j v_j is accumulated inside the update pass.
static int sd_lanczos_engine(sd_model *m, const sd_vec *v0, int dtype, int mm, double tol, double hsign,
                             double *alpha, double *beta, int *m_eff, double *norm0, const double *y, sd_vec *out) {
    sd_ctx *c = m->ctx;
    SD_ARG(mm >= 1 && mm <= SD_HIST_MAX, "lanc_m must be in 1 .. %d", SD_HIST_MAX);
    SdVecGuard G;
    sd_vec *u, *uo, *w;
    SD_TRY(G.make(m, dtype, &u)); SD_TRY(G.make(m, dtype, &uo)); SD_TRY(G.make(m, dtype, &w));
    double *hist = c->d_scal + SD_HIST;                                       // record j: [8 j + 0] = Re d_j (apply), [8 j + 7] = n_j (update: its own
    //                                                                         4 slots, a cross-rank sum always covers all four); n_0 = ||v0||^2
    SD_TRY(sd_vec_copy(u, v0));
    SD_TRY(sd_dot_impl(v0, v0, 1, SD_HIST));                                  // n_0 -> hist[0]
    SD_CUDA(cudaMemcpyAsync(hist + 7, hist + 0, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    double n0;
    SD_TRY(sd_fetch(c, SD_HIST, 1, &n0));
    if (norm0) *norm0 = sqrt(n0);
    if (n0 == 0.0) return sd_fail(SD_ERR_ZERO_NORM, "starting vector has zero norm");
    const double beta0 = sqrt(n0);
    if (y) SD_TRY(sd_vec_zero(out));
    int eff = mm, fetched = 0;
    std::vector<double> hh((size_t)8 * (mm + 1), 0.0);
    const int scratch_slot = SD_HIST + 8 * SD_HIST_MAX;                       // pass 2 does not need its reductions
    for (int j = 1; j <= mm; ++j) {
        SdEpi e = sd_epi_plain(hsign);
        SdLanczosScal S;
        memset(&S, 0, sizeof(S));
        if (!y) {                                                             // pass 1: every scalar from device memory
            e.red = SD_RED_DOT_SELF;
            e.hscale_dev = hist + 8 * (j - 1) + 7;                            // hsign / sqrt(n_{j-1})
            S.d_dev = hist + 8 * j + 0; S.n1_dev = hist + 8 * (j - 1) + 7; S.n2_dev = j >= 2 ? hist + 8 * (j - 2) + 7 : nullptr;
        } else {                                                              // pass 2: the same numbers from alpha / beta
            S.b1 = j >= 2 ? beta[j - 2] : beta0;
            S.b2 = j >= 3 ? beta[j - 3] : beta0;
            S.alpha = alpha[j - 1];
            S.yj = y[j - 1];
            e.hscale = hsign / S.b1;
            if (j == mm) {                                                    // last Ritz term: out += y_m v_m, no further apply
                SD_TRY(sd_axpy_impl(out, sd_host_scalar(S.yj / S.b1, 0), u, sd_host_scalar(0, 0), nullptr, -1));
                break;
            }
        }
        const int slot = y ? scratch_slot : SD_HIST + 8 * j;
        SD_TRY(sd_apply_impl(m, w, u, e, slot));                              // w = H v_j, d_j = <u_j, w>
        if (j < mm) SD_TRY(sd_lanczos_update(w, u, j >= 2 ? uo : nullptr, y ? out : nullptr, S, slot + 4));
        if (!y && (j % 16 == 0 || j == mm)) {                                 // block fetch + breakdown test (Lanczos.jl:65-69,148,228-231)
            SD_TRY(sd_fetch(c, SD_HIST + 8 * fetched, 8 * (j - fetched + 1), hh.data() + 8 * fetched));
            bool stop = false;
            for (int t = fetched + 1; t <= j; ++t) {
                const double b1 = sqrt(hh[8 * (t - 1) + 7]);
                alpha[t - 1] = hh[8 * t + 0] / b1;
                if (t < mm) beta[t - 1] = sqrt(hh[8 * t + 7]);
                if (t < mm && beta[t - 1] < tol) { eff = t; stop = true; break; }
            }
            fetched = j;
            if (stop) break;
        }
        sd_vec *t = uo; uo = u; u = w; w = t;                                 // u_{j+1} = w_j
    }
    *m_eff = eff;
    return SD_OK;
}

int sd_lanczos_extremal(sd_model *m, const sd_vec *v0, int lanc_m, double tol, int negate,
                        double *alpha, double *beta, int *m_eff) {
    SD_ARG(m && v0 && alpha && beta && m_eff, "NULL argument");
    SD_ARG(v0->model == m && v0->dtype == SD_C128, "v0 must be an SD_C128 vector of this model");
    SD_ARG(lanc_m >= 1, "lanc_m must be >= 1");
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    const int mm = (int)std::min<uint64_t>((uint64_t)lanc_m, m->N);          // Lanczos.jl:36
    return sd_lanczos_engine(m, v0, SD_C128, mm, tol, negate ? -1.0 : 1.0, alpha, beta, m_eff, nullptr, nullptr, nullptr);
}

int sd_lanczos_tridiag(sd_model *m, const sd_vec *v, int lanc_m, double tol, double *alpha, double *beta,
                       int *m_eff, double *normv) {
    SD_ARG(m && v && alpha && beta && m_eff && normv, "NULL argument");
    SD_ARG(v->model == m && v->dtype == SD_C128, "v must be an SD_C128 vector of this model");
    SD_ARG(lanc_m >= 1, "lanc_m must be >= 1");
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    const int mm = (int)std::min<uint64_t>((uint64_t)lanc_m, m->N);          // Lanczos.jl:200
    return sd_lanczos_engine(m, v, SD_C128, mm, tol, 1.0, alpha, beta, m_eff, normv, nullptr, nullptr);   // :209-239
}

int sd_lanczos_groundstate(sd_model *m, const sd_vec *v0, int lanc_m, double tol, double orth_tol,
                           double *alpha, double *beta, int *m_actual, sd_vecset **Vout) {
    SD_ARG(m && v0 && alpha && beta && m_actual && Vout, "NULL argument");
    *Vout = nullptr;
    SD_ARG(v0->model == m && v0->dtype == SD_F64, "v0 must be an SD_F64 vector of this model");
    SD_ARG(lanc_m >= 1, "lanc_m must be >= 1");
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    const int mm = (int)std::min<uint64_t>((uint64_t)lanc_m, m->N);          // Lanczos.jl:97
    sd_vecset *S = new (std::nothrow) sd_vecset;
    if (!S) return sd_fail(SD_ERR_NOMEM, "out of host memory");
    struct SetGuard { sd_vecset *s; ~SetGuard() { if (s) sd_vecset_free(s); } } sg{S};
    SdVecGuard G;
    sd_vec *w;
    SD_TRY(G.make(m, SD_F64, &w));
    {
        sd_vec *v1;
        SD_TRY(sd_vec_alloc(m, SD_F64, &v1));
        S->v.push_back(v1);
        double n0;
        SD_TRY(sd_normalised_copy(v1, v0, &n0));                            // :99-100
        if (n0 == 0.0) return sd_fail(SD_ERR_ZERO_NORM, "starting vector has zero norm");
    }
    int mact = mm;
    const bool batch_check = sd_env_int("SD_BATCH_CHECK", 0) != 0;
    // Single GPU: everything behind the apply is one cooperative kernel per step (sd_reorth.cuh), scalars fetched once at the end.
    // SD_REORTH_FUSED=0 keeps the one-call-per-BLAS-operation path below (the only one for sharded models).
    const bool fused = c->world == 1 && sd_env_int("SD_REORTH_FUSED", 1) != 0;
    if (fused) {
        // all steps enqueued back to back: apply, cooperative step kernel, no host synchronisation in between (the next
        // basis vector is allocated while the device works); the scalars of all steps come back in one fetch.  A step behind
        // the reference's `break` (:136-139) returns at once, so what it leaves in its vectors is never looked at.
        SD_TRY(sd_reorth_begin(c, mm));
        for (int j = 1; j <= mm; ++j) {
            sd_vec *vj = S->v[j - 1];
            SD_TRY(sd_apply_impl(m, w, vj, sd_epi_plain(1.0), 0));          // :113
            sd_vec *vn = nullptr;
            if (j < mm) { SD_TRY(sd_vec_alloc(m, SD_F64, &vn)); S->v.push_back(vn); }
            SD_TRY(sd_reorth_step(c, w, S->v.data(), j, vn, tol, orth_tol));
            if (j % 16 == 0 && j < mm) {                                    // bound the work (and the vectors) behind a break
                double stop = 0.0;
                SD_TRY(sd_fetch(c, 3712 /* SD_RTH_SLOT */, 1, &stop));
                if (stop != 0.0) break;
            }
        }
        std::vector<double> rec((size_t)8 * (mm + 1), 0.0);
        SD_TRY(sd_reorth_finish(c, mm, rec.data()));
        for (int j = 1; j <= mm; ++j) {
            const double *r = rec.data() + 8 * j;
            if (r[3] != 1.0) break;                                         // behind a break: did not run
            alpha[j - 1] = r[0];
            if (j < mm) {
                beta[j - 1] = r[1];
                if (r[2] == 1.0) { mact = j; break; }                       // :136-139
                if (r[2] == 2.0) mact = j;                                  // :148-151 leaves the check pass only
            }
        }
        while ((int)S->v.size() > std::max(mact, 1)) { sd_vec_free(S->v.back()); S->v.pop_back(); }   // vectors behind m_actual
    }
    for (int j = 1; !fused && j <= mm; ++j) {
        sd_vec *vj = S->v[j - 1];
        SD_TRY(sd_apply_impl(m, w, vj, sd_epi_plain(1.0), 0));              // :113
        for (int k = 1; k < j; ++k) {                                       // :116-122 sequential MGS
            SD_TRY(sd_dot_impl(S->v[k - 1], w, 1, 8));
            SD_TRY(sd_axpy_impl(w, sd_dev_scalar(c->d_scal + 8, 4), S->v[k - 1], sd_host_scalar(0, 0), nullptr, -1));
        }
        SD_TRY(sd_dot_impl(vj, w, 1, 8));                                   // :124
        SdScalar sa = sd_dev_scalar(c->d_scal + 8, 4);
        if (j == 1) SD_TRY(sd_axpy_impl(w, sa, vj, sd_host_scalar(0, 0), nullptr, 12));
        else SD_TRY(sd_axpy_impl(w, sa, vj, sd_host_scalar(-beta[j - 2], 0), S->v[j - 2], 12));   // :126-130
        double r[8];
        SD_TRY(sd_fetch(c, 8, 8, r));
        alpha[j - 1] = r[0];
        if (j < mm) {
            beta[j - 1] = sqrt(r[7]);                                       // :133
            if (beta[j - 1] < tol) { mact = j; break; }                     // :136-139
            // :142-153 check pass.  SD_BATCH_CHECK=1: the overlaps are taken eight at a time and fetched once
            // (sd_bdot.cuh); an overlap above the tolerance is corrected exactly as below and the pass resumes behind it
            // with the modified w and beta -- the same sequence of operations as the one-at-a-time loop.
            int kfirst = 1;
            while (batch_check && kfirst <= j) {
                const int cnt = j - kfirst + 1;
                const unsigned g = sd_blas_grid(c, w->local_n);
                SD_TRY(sd_partials_reserve(c, (size_t)SD_BDOT_MAX * g));
                std::vector<double> d((size_t)cnt, 0.0);
                for (int k0 = 0; k0 < cnt; k0 += SD_BDOT_MAX) {
                    const int nb = std::min(SD_BDOT_MAX, cnt - k0);
                    SdPtrBlock pb;
                    for (int t = 0; t < SD_BDOT_MAX; ++t) pb.v[t] = (t < nb) ? S->v[kfirst - 1 + k0 + t]->d : nullptr;
                    sd_bdot_f64_kernel<<<g, SD_BLAS_THREADS, 0, c->stream>>>(w->local_n, pb, nb, w->d, c->d_partials, g);
                    SD_TRY(sd_launch_check(c, "sd_bdot_f64_kernel"));
                    sd_bdot_reduce_kernel<<<1, SD_BDOT_MAX * 32, 0, c->stream>>>(c->d_partials, g, nb, c->d_scal + 2048 + (k0 % 1024));
                    SD_TRY(sd_launch_check(c, "sd_bdot_reduce_kernel"));
                    if (c->world > 1) {
                        SD_NCCL(g_nccl.AllReduce(c->d_scal + 2048 + (k0 % 1024), c->d_scal + 2048 + (k0 % 1024), nb, ncclFloat64_, ncclSum_, c->comm, c->stream));
                        sd_collective_done(c);
                    }
                    if ((k0 + SD_BDOT_MAX) % 1024 == 0 || k0 + SD_BDOT_MAX >= cnt) {       // drain the staged block of results
                        const int base = (k0 / 1024) * 1024, have = std::min(cnt - base, 1024);
                        SD_TRY(sd_fetch(c, 2048, have, d.data() + base));
                    }
                }
                int viol = -1;
                for (int t = 0; t < cnt && viol < 0; ++t)
                    if (fabs(d[t]) / beta[j - 1] > orth_tol) viol = t;
                if (viol < 0) { kfirst = j + 1; break; }
                const int k = kfirst + viol;
                SD_TRY(sd_axpy_impl(w, sd_host_scalar(-d[viol], 0), S->v[k - 1], sd_host_scalar(0, 0), nullptr, 12));
                double nn[4];
                SD_TRY(sd_fetch(c, 12, 4, nn));
                beta[j - 1] = sqrt(nn[3]);
                if (beta[j - 1] < tol) { mact = j; kfirst = j + 1; break; }   // like the loop below: leaves the check pass only
                kfirst = k + 1;
            }
            for (int k = kfirst; k <= j; ++k) {                             // one at a time (the default)
                SD_TRY(sd_dot_impl(S->v[k - 1], w, 1, 8));
                double d[2];
                SD_TRY(sd_fetch(c, 8, 2, d));
                if (fabs(d[0]) / beta[j - 1] > orth_tol) {
                    SD_TRY(sd_axpy_impl(w, sd_host_scalar(-d[0], 0), S->v[k - 1], sd_host_scalar(0, 0), nullptr, 12));
                    double nn[4];
                    SD_TRY(sd_fetch(c, 12, 4, nn));
                    beta[j - 1] = sqrt(nn[3]);
                    if (beta[j - 1] < tol) { mact = j; break; }
                }
            }
            sd_vec *vn;
            SD_TRY(sd_vec_alloc(m, SD_F64, &vn));
            S->v.push_back(vn);
            SD_TRY(sd_divide_impl(vn, w, sd_host_scalar(beta[j - 1], 0)));    // :155
        }
    }
    *m_actual = mact;
    *Vout = S;
    sg.s = nullptr;
    return SD_OK;
}

// Memory-lean ground state: NOT a reference function.  lanczos_groundstate (Lanczos.jl:87-181) keeps
// the N x m basis V (481 GB at L = 32, m = 100) for full reorthogonalisation and for the Ritz vector V*y.  This is the
// plain three-term recurrence on three work vectors, run twice from the same v0:
//   pass 1 (y == NULL): alpha[0 .. m_eff-1], beta[0 .. m_eff-2] (apply with the fused <v,Hv>, axpy with the fused ||w||^2);
//   pass 2 (y != NULL): the SAME kernels in the same order with the stored alpha / beta as host scalars -- the
//       regenerated v_j are bit-identical to pass 1 -- accumulating out = sum_{j < m_eff} y[j] v_j; *norm2 = ||out||^2.
// Without reorthogonalisation converged Ritz values reappear as copies, which leaves the lowest Ritz value and the
// direction of V*y alone; callers gate on E0 against the faithful path (tests: 1e-10).
// Memory-lean ground state: sd_lanczos_engine on three f64 vectors, called twice by the host.  Pass 1
// (y == NULL) returns alpha / beta; pass 2 (y = the Ritz coefficients, alpha / beta of pass 1 as inputs) accumulates
// out = sum_j y_j v_j from the regenerated vectors; norm2 = ||out||^2.
int sd_lanczos_lean(sd_model *m, const sd_vec *v0, int lanc_m, double tol, double *alpha, double *beta, int *m_eff,
                    const double *y, sd_vec *out, double *norm2) {
    SD_ARG(m && v0 && alpha && beta && m_eff, "NULL argument");
    SD_ARG(v0->model == m && v0->dtype == SD_F64, "v0 must be an SD_F64 vector of this model");
    SD_ARG(lanc_m >= 1, "lanc_m must be >= 1");
    SD_ARG(!y || (out && norm2 && out->model == m && out->dtype == SD_F64 && out->d != v0->d), "pass 2 needs y, an SD_F64 out and norm2");
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    const int mm = (int)std::min<uint64_t>((uint64_t)lanc_m, m->N);
    SD_TRY(sd_lanczos_engine(m, v0, SD_F64, mm, tol, 1.0, alpha, beta, m_eff, nullptr, y, out));
    if (y) {
        SD_TRY(sd_dot_impl(out, out, 1, 0));
        double r[2];
        SD_TRY(sd_fetch(c, 0, 2, r));
        *norm2 = r[0];
    }
    return SD_OK;
}

int sd_kpm_moments(sd_model *m, const sd_vec *phi, int M, double a, double b, double *mu) {
    SD_ARG(m && phi && mu, "NULL argument");
    SD_ARG(phi->model == m && phi->dtype == SD_C128, "phi must be an SD_C128 vector of this model");
    SD_ARG(M >= 1, "M must be >= 1");
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    SdVecGuard G;
    sd_vec *vp, *vc;
    SD_TRY(G.make(m, SD_C128, &vp)); SD_TRY(G.make(m, SD_C128, &vc));
    for (int i = 0; i < M; ++i) mu[i] = 0.0;
    SD_TRY(sd_vec_copy(vp, phi));                                           // KPM_Sqw.jl:100
    SD_TRY(sd_dot_impl(phi, vp, 1, 0));                                     // :104
    double r[4];
    SD_TRY(sd_fetch(c, 0, 2, r));
    mu[0] = r[0];
    if (M == 1) return SD_OK;
    {                                                                       // :106-107
        SdEpi e = sd_epi_plain(1.0);
        e.mode = SD_EPI_RESCALED; e.a = a; e.b = b;
        e.red = SD_RED_DOT_PHI; e.phi = phi->d;
        SD_TRY(sd_apply_impl(m, vc, vp, e, 0));
        SD_TRY(sd_fetch(c, 0, 4, r));
        mu[1] = r[2];
    }
    // :109-126.  The reductions of moment n stay on the device (d_scal[SD_HIST + 8 (n mod SD_HIST_MAX) ..]) and are fetched
    // 32 moments at a time: no host synchronisation per moment.  The reference renormalises v_next when ||v_next|| > 1e3
    // (:117-121), which only happens with wrong rescaling bounds; the norms arrive with the same block, and the first
    // one above the threshold sends the rest of the loop down the step-by-step path from the vectors of that moment --
    // which are gone by then, so the speculative block is re-run from a checkpoint taken at its start.
    SdVecGuard CK;
    sd_vec *cp = nullptr, *cc = nullptr;                                    // checkpoint of (v_prev, v_curr) at the block start
    const int BLK = 32;
    int n = 2;
    bool careful = false;
    while (n < M) {
        const int n1 = std::min(M, n + BLK);
        if (!careful) {
            if (!cp) { SD_TRY(CK.make(m, SD_C128, &cp)); SD_TRY(CK.make(m, SD_C128, &cc)); }
            SD_TRY(sd_vec_copy(cp, vp)); SD_TRY(sd_vec_copy(cc, vc));
            sd_vec *p0 = vp, *c0 = vc;
            for (int t = n; t < n1; ++t) {
                sd_complex zero = {0, 0};
                SD_TRY(sd_cheb_step_impl(m, vp, vc, vp, a, b, phi, nullptr, zero, SD_HIST + 8 * (t - n)));   // v_next overwrites v_prev
                std::swap(vp, vc);
            }
            std::vector<double> hh((size_t)8 * (n1 - n));
            SD_TRY(sd_fetch(c, SD_HIST, 8 * (n1 - n), hh.data()));
            bool blown = false;
            for (int t = n; t < n1; ++t) if (sqrt(hh[8 * (t - n) + 3]) > 1e3) blown = true;
            if (!blown) {
                for (int t = n; t < n1; ++t) mu[t] = hh[8 * (t - n) + 2];
                n = n1;
                continue;
            }
            vp = p0; vc = c0;                                               // restore and redo this block step by step
            SD_TRY(sd_vec_copy(vp, cp)); SD_TRY(sd_vec_copy(vc, cc));
            careful = true;
        }
        for (int t = n; t < n1; ++t) {
            sd_complex zero = {0, 0};
            SD_TRY(sd_cheb_step_impl(m, vp, vc, vp, a, b, phi, nullptr, zero, 0));
            SD_TRY(sd_fetch(c, 0, 4, r));
            mu[t] = r[2];
            const double nv = sqrt(r[3]);
            if (nv > 1e3) SD_TRY(sd_divide_impl(vp, vp, sd_host_scalar(nv, 0)));   // :118-121
            std::swap(vp, vc);
        }
        n = n1;
    }
    return SD_OK;
}

int sd_krylov_basis(sd_model *m, const sd_vec *psi0, int kry_m, sd_complex *alpha, double *beta, int *m_eff,
                    double *norm0, sd_vecset **Vout) {
    SD_ARG(m && psi0 && alpha && beta && m_eff && norm0 && Vout, "NULL argument");
    *Vout = nullptr;
    SD_ARG(psi0->model == m, "psi0 belongs to a different model");
    SD_ARG(kry_m >= 1, "kry_m must be >= 1");
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    const int dt = psi0->dtype;
    sd_vecset *S = new (std::nothrow) sd_vecset;
    if (!S) return sd_fail(SD_ERR_NOMEM, "out of host memory");
    struct SetGuard { sd_vecset *s; ~SetGuard() { if (s) sd_vecset_free(s); } } sg{S};
    SdVecGuard G;
    sd_vec *w;
    SD_TRY(G.make(m, dt, &w));
    sd_vec *v1;
    SD_TRY(sd_vec_alloc(m, dt, &v1));
    S->v.push_back(v1);
    SD_TRY(sd_normalised_copy(v1, psi0, norm0));                            // Krylov.jl:147-150
    int eff = kry_m;
    if (*norm0 == 0.0) { *m_eff = 0; *Vout = S; sg.s = nullptr; return SD_OK; }
    for (int j = 1; j <= kry_m; ++j) {                                      // :153-172
        sd_vec *vj = S->v[j - 1];
        SdEpi e = sd_epi_plain(1.0);
        e.red = SD_RED_DOT_SELF;
        SD_TRY(sd_apply_impl(m, w, vj, e, 0));                              // alpha_j = dot(V_j, w), complex
        SdScalar sa = sd_dev_scalar(c->d_scal + 0, 3);
        if (j == 1) SD_TRY(sd_axpy_impl(w, sa, vj, sd_host_scalar(0, 0), nullptr, 4));
        else SD_TRY(sd_axpy_impl(w, sa, vj, sd_host_scalar(-beta[j - 2], 0), S->v[j - 2], 4));
        double r[8];
        SD_TRY(sd_fetch(c, 0, 8, r));
        alpha[j - 1].re = r[0]; alpha[j - 1].im = r[1];
        if (j < kry_m) {
            beta[j - 1] = sqrt(r[7]);
            if (fabs(beta[j - 1]) < 1e-14) { eff = j; break; }              // :163-168
            sd_vec *vn;
            SD_TRY(sd_vec_alloc(m, dt, &vn));
            S->v.push_back(vn);
            SD_TRY(sd_divide_impl(vn, w, sd_host_scalar(beta[j - 1], 0)));
        }
    }
    *m_eff = eff;
    *Vout = S;
    sg.s = nullptr;
    return SD_OK;
}

int sd_chebyshev_evolve(sd_model *m, const sd_vec *psi0, const sd_complex *cf, int n, double a, double b,
                        sd_vec *out) {
    SD_ARG(m && psi0 && cf && out, "NULL argument");
    SD_ARG(psi0->model == m && out->model == m, "vector belongs to a different model");
    SD_ARG(psi0->dtype == SD_C128 && out->dtype == SD_C128, "psi0 and out must be SD_C128 (InexactError otherwise)");
    SD_ARG(n >= 1, "cheb_n must be >= 1");
    SD_ARG(out->d != psi0->d, "out must not alias psi0");
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    SdVecGuard G;
    sd_vec *vp, *vc;
    SD_TRY(G.make(m, SD_C128, &vp)); SD_TRY(G.make(m, SD_C128, &vc));
    SD_TRY(sd_vec_copy(vp, psi0));                                          // Chebyshev.jl:95
    SD_TRY(sd_apply_rescaled_H(m, vc, vp, a, b));                           // :98
    SD_TRY(sd_vec_zero(out));                                               // :100-107
    SD_TRY(sd_axpy_impl(out, sd_host_scalar(cf[0].re, cf[0].im), vp, sd_host_scalar(0, 0), nullptr, -1));
    if (n >= 2) SD_TRY(sd_axpy_impl(out, sd_host_scalar(cf[1].re, cf[1].im), vc, sd_host_scalar(0, 0), nullptr, -1));
    for (int k = 2; k < n; ++k) {                                           // :110-121
        SD_TRY(sd_cheb_step_impl(m, vp, vc, vp, a, b, nullptr, out, cf[k], 0));
        std::swap(vp, vc);
    }
    SD_CUDA(cudaStreamSynchronize(c->stream));
    return SD_OK;
}

