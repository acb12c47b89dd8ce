/*
 * spindyn.h -- C ABI of libspindyn_cuda: the B200 (sm_100a) drop-in for the
 * matrix-free H.psi hot path of javahedi/SpinDynamics.jl.
 *
 * The reference is pure Julia and has no FFI for this path; its de-facto
 * operator interface is the callback  applyH!(out, psi, model) -> out  taken by
 * every recurrence (reference Lanczos.jl:27-33,87-94,196-198,255-256;
 * Hamiltonian.jl:286-288; Krylov.jl:136-137; Chebyshev.jl:61-64) and hard-coded
 * as Hamiltonian.apply_H! by the public layer (PublicAPI.jl:28,62,70,80).
 * Each entry point below names the reference function (file:line under
 * /root/reference/src) whose device-side work it replaces.  INTEGRATION.md shows
 * the Julia `ccall` stubs that bind these symbols.
 *
 * Conventions
 *   - plain C, no exceptions: every function returns 0 on success and a negative
 *     sd_status otherwise; sd_last_error() returns a thread-local message.
 *   - handles are opaque, owned by the library until the matching *_free.
 *   - sites and basis indices that cross the ABI follow Julia: bonds carry
 *     1-based sites (exactly the memory of Vector{Tuple{Int,Int,Float64}}),
 *     sd_rank returns 1-based indices with 0 = absent (get(idxmap, s, 0)).
 *     Element offsets inside vectors (first/count, onehot) are 0-based.
 *   - dtype SD_F64 = Vector{Float64}; SD_C128 = Vector{ComplexF64}, interleaved
 *     (re, im) doubles.
 *   - vectors live in device memory.  In a multi-rank context every vector is
 *     sharded by contiguous basis-rank range; host buffers passed to
 *     upload/download hold the LOCAL shard (sd_model_local_range).
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with SD_ERR_CUDA.
 *   - calls on one context run on that context's stream.  Entry points that
 *     return scalars synchronise; the others only enqueue (sd_ctx_sync waits).
 */
#ifndef SPINDYN_H
#define SPINDYN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sd_ctx sd_ctx;
typedef struct sd_model sd_model;
typedef struct sd_vec sd_vec;
typedef struct sd_vecset sd_vecset;

/* One entry of hopping_list / zz_list (SpinModel.jl:6-15): Tuple{Int,Int,Float64}. */
typedef struct { int64_t i, j; double J; } sd_bond;

typedef struct { double re, im; } sd_complex;

enum sd_dtype { SD_F64 = 0, SD_C128 = 1 };

enum sd_status {
    SD_OK = 0,
    SD_ERR_ARG = -1,        /* Julia ArgumentError / DimensionMismatch / AssertionError */
    SD_ERR_CUDA = -2,       /* CUDA runtime failure, or no device                       */
    SD_ERR_NOMEM = -3,
    SD_ERR_NCCL = -4,
    SD_ERR_ZERO_NORM = -5,  /* error("starting vector has zero norm")  Lanczos.jl:210-212 */
    SD_ERR_UNSUPPORTED = -6
};

/* Which apply kernel a model was routed to (sd_model_info). */
enum sd_kernel_path {
    SD_PATH_GENERIC = 0,    /* one thread per state, arbitrary bond lists, full or sector  */
    SD_PATH_TILED = 1,      /* sector basis, open nearest-neighbour chain: smem-tiled gather */
    SD_PATH_BLOCK = 2       /* same models, L >= 16: block-layout vectors, TMA tiles, no staging (default) */
};

const char *sd_last_error(void);
int sd_version(void);
int sd_device_count(int *n);

/* ------------------------------------------------------------------ context */
/* One context = one GPU + one stream.  sd_ctx_create_rank makes this process
 * rank `rank` of `world` cooperating processes on one node (one per GPU):
 * vectors are sharded by basis-rank range, remote hop targets are read through
 * CUDA-IPC-mapped peer shards over NVLink, scalars are all-reduced with NCCL
 * (id128 = the 128-byte ncclUniqueId from sd_nccl_unique_id on rank 0). */
int sd_ctx_create(int device, sd_ctx **ctx);
int sd_nccl_unique_id(void *id128);
int sd_ctx_create_rank(int device, int rank, int world, const void *id128, sd_ctx **ctx);
int sd_ctx_free(sd_ctx *ctx);
int sd_ctx_sync(sd_ctx *ctx);
/* Multi-rank contract (world > 1): every rank makes the same library calls in the same order (SPMD).  Collective
 * calls are sd_model_create, sd_vec_alloc, sd_ctx_collect, every call that returns a reduced scalar, and the solvers
 * built on them.  sd_vec_free is NOT collective (finalizers may call it at any time): the local shard stays allocated
 * until every rank has freed the same vector and a later sd_vec_alloc / sd_ctx_collect has told the peers.  The
 * library orders peer reads against writes itself (it tracks which vectors were written or gathered since the last
 * collective and inserts a stream-ordered barrier where needed); callers never synchronise ranks by hand. */
int sd_ctx_collect(sd_ctx *ctx);
int sd_ctx_rank(const sd_ctx *ctx, int *rank, int *world);
/* CUDA-event stopwatch on the context's stream (bench.py uses it so the timed
 * region is measured on the stream the kernels are launched on). */
int sd_timer_start(sd_ctx *ctx);
int sd_timer_stop(sd_ctx *ctx, float *ms);
/* Number of kernels this library launched on the context so far. */
int sd_launch_count(const sd_ctx *ctx, uint64_t *n);

/* -------------------------------------------------------------------- model */
/* Replaces build_model / XXZChain's device-relevant part (SpinModel.jl:23-38,
 * 63-90) WITHOUT materialising states[] or idxmap (Basis.jl:37-53): the basis is
 * the combinatorial number system of the `combinations(1:L,nup)` order.
 * nup = -1 selects the full basis (mode :full).  Validation follows
 * Basis.jl:9-20 (1 <= L <= 63, 0 <= nup <= L -> SD_ERR_ARG).  `field` has L
 * entries.  Bond sites are 1-based. */
int sd_model_create(sd_ctx *ctx, int L, int nup,
                    const sd_bond *hop, int nhop,
                    const sd_bond *zz, int nzz,
                    const double *field, sd_model **model);
int sd_model_free(sd_model *model);
int sd_model_dim(const sd_model *model, uint64_t *dim);
/* This rank's shard [first, first+count) of the basis (whole basis if world=1). */
int sd_model_local_range(const sd_model *model, uint64_t *first, uint64_t *count);
/* Shard boundaries for any world size: bounds[0..world], host-only arithmetic. */
int sd_model_shard_bounds(const sd_model *model, int world, uint64_t *bounds);
int sd_model_info(const sd_model *model, int *kernel_path, int *tile_sites, int *rank_bits);
/* Force a kernel path (tests compare the two); SD_ERR_UNSUPPORTED if the model
 * does not qualify for it. */
int sd_model_set_path(sd_model *model, int kernel_path);

/* -------------------------------------------------------------------- basis */
/* states[i] = model.states[first+i+1]   (Basis.jl:37-53 / :23-34), computed on
 * the device by unranking; host output. */
int sd_unrank(sd_model *model, uint64_t first, uint64_t count, uint64_t *states);
/* idx1[i] = get(model.idxmap, states[i], 0)   (Hamiltonian.jl:260,
 * InitialStates.jl:_one_hot): 1-based, 0 when the state is outside the basis. */
int sd_rank(sd_model *model, const uint64_t *states, uint64_t count, int64_t *idx1);

/* ------------------------------------------------------------------ vectors */
int sd_vec_alloc(sd_model *model, int dtype, sd_vec **vec);   /* collective when world > 1 */
int sd_vec_free(sd_vec *vec);                                 /* local; see the multi-rank contract above */
int sd_vec_dtype(const sd_vec *vec, int *dtype);
int sd_vec_local_len(const sd_vec *vec, uint64_t *n);
int sd_vec_upload(sd_vec *vec, const void *host);       /* local shard, synchronous   */
int sd_vec_download(sd_vec *vec, void *host);
/* Asynchronous copies between a PINNED host buffer (sd_host_alloc) and the local shard.  They return at once; the
 * library orders them against its own kernels.  Block-layout vectors go through the copy engine: two dedicated copy
 * streams, the transfer cut into chunks whose layout permutes overlap the next chunk's transfer, and an upload issued
 * after a download runs concurrently with it (both PCIe directions busy).  A downloaded buffer is valid after
 * sd_ctx_sync (or sd_timer_stop, whose stopwatch covers the copy streams); the host buffer of an upload must stay
 * untouched until then as well. */
int sd_vec_upload_async(sd_vec *vec, const void *pinned_host);
int sd_vec_download_async(sd_vec *vec, void *pinned_host);
int sd_vec_zero(sd_vec *vec);
int sd_vec_set_onehot(sd_vec *vec, uint64_t idx0);      /* InitialStates.jl one-hot   */
/* out[i] = vec[idx0[i]] (0-based basis ranks) for the ranks the local shard holds (present[i] = 1), untouched
 * otherwise (present[i] = 0); host output, any layout.  The Julia side's getindex on a device vector, and how
 * bench.py samples rows of H.psi at sizes nobody can download (count <= 65536). */
int sd_vec_get(sd_vec *vec, const uint64_t *idx0, uint64_t count, void *out, unsigned char *present);
/* psi[r] = 2*u(splitmix64(seed ^ r)) - 1 (re: seed, im: seed+1), then * scale.
 * Counter-based bench/test input; oracle.c:seeded_value is the same formula. */
int sd_vec_fill_seeded(sd_vec *vec, uint64_t seed, double scale);
int sd_vec_copy(sd_vec *dst, const sd_vec *src);
/* dst = ComplexF64.(src)  (LanczosSqw.jl:56, KPM_Sqw.jl:205) or Float64<-real part */
int sd_vec_convert(sd_vec *dst, const sd_vec *src);
int sd_vec_scale(sd_vec *x, sd_complex s);              /* x *= s                      */
int sd_vec_axpy(sd_vec *y, sd_complex a, const sd_vec *x);   /* y += a x               */
/* dot(x, y) = sum conj(x_i) y_i  (LinearAlgebra.dot); all-reduced over ranks;
 * deterministic (fixed-order reduction, test_Lanczos.jl:122-166). */
int sd_vec_dot(const sd_vec *x, const sd_vec *y, sd_complex *result);
/* unconjugated sum x_i y_i (LanczosSqw.jl:59 takes dot(conj(psi), H psi)) */
int sd_vec_dotu(const sd_vec *x, const sd_vec *y, sd_complex *result);
int sd_vec_norm(const sd_vec *x, double *result);
/* Host pointers need page-locked memory for async copies. */
int sd_host_alloc(void **p, uint64_t bytes);
int sd_host_free(void *p);

/* ----------------------------------------------------------------- operator */
/* apply_H!(out, psi, model)                    Hamiltonian.jl:211-273
 * out and psi must be distinct vectors of the same dtype. */
int sd_apply_H(sd_model *model, sd_vec *out, const sd_vec *psi);
/* apply_H! fused with dot(psi, out)  (Lanczos alpha: Lanczos.jl:50,124,219;
 * E0: LanczosSqw.jl:58-59, KPM_Sqw.jl:208-209). */
int sd_apply_H_dot(sd_model *model, sd_vec *out, const sd_vec *psi, sd_complex *dot);
/* apply_rescaled_H!: out = (H psi - b psi)/a       Hamiltonian.jl:286-301 */
int sd_apply_rescaled_H(sd_model *model, sd_vec *out, const sd_vec *psi, double a, double b);
/* Chebyshev step, one kernel:  vnext = 2 (H v - b v)/a - vprev
 * (KPM_Sqw.jl:111-112, Chebyshev.jl:112-115).  vnext may alias vprev.
 * Optional fusions (NULL to skip):
 *   phi  -> *mu = Re dot(phi, vnext) and *norm2 = ||vnext||^2   (KPM_Sqw.jl:114-117)
 *   acc  -> acc += ck * vnext                                   (Chebyshev.jl:116) */
int sd_cheb_step(sd_model *model, sd_vec *vnext, const sd_vec *v, const sd_vec *vprev,
                 double a, double b,
                 const sd_vec *phi, double *mu, double *norm2,
                 sd_vec *acc, sd_complex ck);
/* Sz_q_vector: phi = L^-1/2 sum_r e^{iqr} s_r psi0 (phi C128, psi0 F64 or C128),
 * *norm2 = ||phi||^2 if non-NULL               Hamiltonian.jl:307-337 */
int sd_szq(sd_model *model, sd_vec *phi, const sd_vec *psi0, double q, double *norm2);
/* phi (SD_C128) = (sum_r w[r] s_r(state)) * ComplexF64(psi0) with complex per-site weights w[L] (site r = bit r, 0-based):
 * the single-site S^z_i that TimeEvolution/KPM.jl:197-213 builds with create_spin_operator(i, :z) is w = delta_{r,i};
 * Sz_q_vector is w[r] = e^{iqr} / sqrt(L).  norm2 (optional) receives ||phi||^2. */
int sd_apply_sz_weights(sd_model *model, sd_vec *phi, const sd_vec *psi0, const sd_complex *w, double *norm2);
/* Observables.jl:14-109 on a device-resident vector, summed over ranks:
 *   mags[i] = sum |psi|^2 s_i                      magnetization_per_site (:14-37)
 *   zz[r]   = sum_i sum |psi|^2 s_i s_{(i+r) mod L}  the cyclic diagonals of SzSz (:48-93); the caller forms
 *   C_r = (zz[r] - sum_i mags[i] mags[(i+r) mod L]) / L  (connected_correlations) and its FFT (structure_factor_Sq).
 * mags and zz are host arrays of L doubles. */
int sd_vec_observables(const sd_vec *psi, double *mags, double *zz);
/* Host-pointer convenience backing apply_H!(::Vector, ::Vector, ::Model): H2D,
 * apply, D2H, synchronous.  Whole-basis vectors; world must be 1. */
int sd_apply_H_host(sd_model *model, int dtype, void *out, const void *psi);

/* -------------------------------------------------------------- recurrences */
/* A set of m device vectors (V::Matrix N x m of Lanczos.jl:104; Vector{Vector}
 * of Lanczos.jl:202, Krylov.jl:140). */
int sd_vecset_free(sd_vecset *set);
int sd_vecset_size(const sd_vecset *set, int *m);
int sd_vecset_get(sd_vecset *set, int k, sd_vec **vec);    /* borrowed handle          */
/* out = sum_{k<m} y[k] V_k, *norm2 = ||out||^2  (Ritz vector Lanczos.jl:170-171;
 * Krylov recombination Krylov.jl:185-190).  y has m sd_complex entries. */
int sd_lincomb(sd_vecset *set, const sd_complex *y, int m, sd_vec *out, double *norm2);

/* lanczos_extremal's recurrence (Lanczos.jl:27-84): v0 is the C128 start vector
 * (randn stays on the host); it is normalised internally.  alpha[m], beta[m-1]
 * with m = min(lanc_m, N); *m_eff = entries of alpha that are valid.  negate!=0
 * runs on -H (estimate_energy_bounds' wrapper, Lanczos.jl:261-265). */
int sd_lanczos_extremal(sd_model *model, const sd_vec *v0, int lanc_m, double tol, int negate,
                        double *alpha, double *beta, int *m_eff);
/* lanczos_groundstate's recurrence (Lanczos.jl:87-157): F64 start vector, full
 * reorthogonalisation and the second check pass; keeps V on the device. */
int sd_lanczos_groundstate(sd_model *model, const sd_vec *v0, int lanc_m, double tol,
                           double orthogonalize_tol, double *alpha, double *beta,
                           int *m_actual, sd_vecset **V);
/* lanczos_tridiag (Lanczos.jl:196-246): C128 start vector; returns alpha[m_eff],
 * beta[m_eff-1], normv.  SD_ERR_ZERO_NORM mirrors :210-212. */
int sd_lanczos_tridiag(sd_model *model, const sd_vec *v, int lanc_m, double tol,
                       double *alpha, double *beta, int *m_eff, double *normv);
/* Memory-lean ground state (an extension, not a reference function): the plain three-term
 * Lanczos recurrence on three work vectors instead of the N x m basis of Lanczos.jl:104, run twice from the same v0.
 * Pass 1 (y == NULL; out, norm2 unused): alpha[lanc_m], beta[lanc_m], *m_eff.  Pass 2 (y[m_eff] = lowest
 * eigenvector of the tridiagonal matrix, lanc_m = m_eff of pass 1): out = sum_j y[j] v_j, *norm2 = ||out||^2. */
int sd_lanczos_lean(sd_model *model, const sd_vec *v0_f64, int lanc_m, double tol, double *alpha, double *beta,
                    int *m_eff, const double *y, sd_vec *out, double *norm2);
/* compute_chebyshev_moments (KPM_Sqw.jl:95-128): mu[M], phi C128. */
int sd_kpm_moments(sd_model *model, const sd_vec *phi, int M, double a, double b, double *mu);
/* q-batched S(q,w) recurrences: the bodies of the Threads.@threads q-loops of lanczos_sqw
 * (LanczosSqw.jl:65-77) and kpm_sqw (KPM_Sqw.jl:218-253) for nq momenta at once, on one interleaved [state][q]
 * multi-vector: phi_q = Sz_q_vector(model, psi0, q) (Hamiltonian.jl:307-337) for every q, then ONE fused kernel pair
 * per Lanczos step (one per Chebyshev moment) for all momenta, per-momentum scalars kept on the device.  psi0 F64 or
 * C128; 1 <= nq <= 128; single-GPU contexts (SD_ERR_ARG otherwise: loop over q with sd_szq + sd_lanczos_tridiag).
 * Any model (full / sector basis, arbitrary bond lists).  Needs 3 * 16 * nq' * N bytes of device memory, nq' = nq
 * rounded up to 2, 3 or 4 times a power of two (SD_ERR_NOMEM if that does not fit: pass fewer momenta per call).
 *   sd_lanczos_tridiag_szq_batch: lanczos_tridiag (Lanczos.jl:196-246) per momentum.  alpha[c * lanc_m + t],
 *     beta[c * lanc_m + t] (both nq * lanc_m doubles), m_eff[c] valid entries (0: norm(phi_c) == 0, the reference's
 *     `continue`), norm_phi[c].
 *   sd_kpm_moments_szq_batch: compute_chebyshev_moments (KPM_Sqw.jl:95-128) of phi_c / norm(phi_c) per momentum.
 *     mu[c * M + n]; norm_phi[c] (0: row of zeros).  *blown = 1 if some ||v_next|| exceeded 1e3, where the reference
 *     renormalises (:117-121, only with wrong rescaling bounds): the moments are then not the reference's -- use the
 *     per-momentum sd_kpm_moments. */
int sd_lanczos_tridiag_szq_batch(sd_model *model, const sd_vec *psi0, const double *q, int nq, int lanc_m, double tol,
                                 double *alpha, double *beta, int *m_eff, double *norm_phi);
int sd_kpm_moments_szq_batch(sd_model *model, const sd_vec *psi0, const double *q, int nq, int M, double a, double b,
                             double *mu, double *norm_phi, int *blown);
/* Free and total device memory of the context's GPU in bytes (sizing the momenta per batch). */
int sd_ctx_mem_info(sd_ctx *ctx, uint64_t *free_bytes, uint64_t *total_bytes);
/* krylov_time_evolve's basis build (Krylov.jl:136-173): alpha complex (the
 * reference keeps dot(V_j, w) complex, :155), beta real; host does the small
 * eigen problem and calls sd_lincomb. */
int sd_krylov_basis(sd_model *model, const sd_vec *psi0, int kry_m,
                    sd_complex *alpha, double *beta, int *m_eff, double *norm0,
                    sd_vecset **V);
/* chebyshev_time_evolve's recurrence (Chebyshev.jl:95-123): psi0 C128, c[n] the
 * host-computed coefficients (:70-79), out = psi_t. */
int sd_chebyshev_evolve(sd_model *model, const sd_vec *psi0, const sd_complex *c, int n,
                        double a, double b, sd_vec *out);

#ifdef __cplusplus
}
#endif
#endif /* SPINDYN_H */
