// sd_batch.cu -- second translation unit of libspindyn_cuda: the recurrences whose device work is ONE kernel per step.
//   * sd_reorth_step: the full reorthogonalisation of lanczos_groundstate as a cooperative kernel (sd_reorth.cuh)
//   * the q-batched S(q, w) recurrences on an interleaved [state][q] multi-vector (sd_mv.cuh)
// Host code only orchestrates.  No CPU fallback.
#define SD_NO_KERNELS
#include "sd_handles.h"
#include "sd_reorth.cuh"

#define SD_RTH_SLOT 3712                                              // d_scal[3712]: the stop flag of the running solve

// One step of Lanczos.jl:116-155 behind the apply (see sd_reorth.cuh).  V[0 .. j-1] are v_1 .. v_j, w = H v_j on entry;
// vnext (j < m) receives v_{j+1}.  Nothing is fetched: the step's scalars (alpha_j, beta_j, breakdown flag, "ran") stay in
// d_scal[SD_HIST + 8 j ..], beta_{j-1} is read from the previous step's record, and sd_reorth_finish returns them all at
// once -- a solve is m applies and m cooperative launches enqueued back to back.  Single GPU.
int sd_reorth_begin(sd_ctx *c, int mm) {
    SD_ARG(c->world == 1, "the fused reorthogonalisation is single-GPU");
    SD_ARG(mm >= 1 && mm <= SD_HIST_MAX, "lanc_m must be in 1 .. %d", SD_HIST_MAX);
    SD_CUDA(cudaMemsetAsync(c->d_scal + SD_RTH_SLOT, 0, sizeof(double), c->stream));
    SD_CUDA(cudaMemsetAsync(c->d_scal + SD_HIST, 0, (size_t)8 * (mm + 1) * sizeof(double), c->stream));
    return SD_OK;
}
int sd_reorth_step(sd_ctx *c, sd_vec *w, sd_vec *const *V, int j, sd_vec *vnext, double tol, double orth_tol) {
    SD_ARG(c->world == 1, "sd_reorth_step is single-GPU");
    SD_ARG(j >= 1 && j <= SD_HIST_MAX, "too many basis vectors");
    if (!c->d_vtab) {
        SD_CUDA(cudaMalloc(&c->d_vtab, (size_t)(SD_HIST_MAX + 1) * sizeof(double *)));
        int per_sm = 0;
        SD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sd_reorth_step_kernel, SD_RTH_THREADS, 0));
        SD_ARG(per_sm >= 1, "sd_reorth_step_kernel does not fit an SM");
        c->rth_grid_max = (unsigned)(c->sm_count * std::min(per_sm, 4));
        SD_CUDA(cudaMalloc(&c->d_rth_partials, (size_t)2 * SD_RCHK * c->rth_grid_max * sizeof(double)));
        c->vtab_count = 0;
    }
    // the table only grows while a solve runs; entry t is rewritten when a new solve starts with other vectors
    for (int t = 0; t < j; ++t) {
        if (t < c->vtab_count && c->h_vtab[t] == V[t]->d) continue;
        if ((int)c->h_vtab.size() <= t) c->h_vtab.resize(t + 1, nullptr);
        c->h_vtab[t] = V[t]->d;
        SD_CUDA(cudaMemcpyAsync(c->d_vtab + t, &c->h_vtab[t], sizeof(double *), cudaMemcpyHostToDevice, c->stream));   // pageable source: staged before the call returns
        c->vtab_count = std::max(c->vtab_count, t + 1);
    }
    SdReorthArgs A;
    A.w = w->d; A.V = c->d_vtab; A.vnext = vnext ? vnext->d : nullptr; A.j = j; A.n = w->local_n;
    A.beta_prev = j >= 2 ? c->d_scal + SD_HIST + 8 * (j - 1) + 1 : nullptr;
    A.stop = c->d_scal + SD_RTH_SLOT;
    A.tol = tol; A.orth_tol = orth_tol;
    A.partials = c->d_rth_partials; A.out = c->d_scal + SD_HIST + 8 * j;
    const uint64_t want = (A.n + SD_RTH_THREADS - 1) / SD_RTH_THREADS;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, c->rth_grid_max));
    void *args[] = {&A};
    SD_CUDA(cudaLaunchCooperativeKernel((const void *)sd_reorth_step_kernel, dim3(grid), dim3(SD_RTH_THREADS), args, 0, c->stream));
    return sd_launch_check(c, "sd_reorth_step_kernel");
}
// records of steps 1 .. mm: rec[8 j + 0..3] = alpha_j, beta_j, flag, ran (one fetch, one synchronisation)
int sd_reorth_finish(sd_ctx *c, int mm, double *rec) {
    return sd_fetch(c, SD_HIST, 8 * (mm + 1), rec);
}

// =====================================================================================================================
// q-batched S(q, w) recurrences (sd_mv.cuh)
#include "sd_mv.cuh"

namespace {
struct MvPlan {
    int nq = 0, nqp = 0, qpt = 0, lps_log2 = 0;
    unsigned grid = 1;
};
// smallest padded column count QPT * 2^e >= nq with QPT in {2, 3, 4}, LPS = 2^e <= 32 lanes per state
bool mv_plan(const sd_ctx *c, uint64_t N, int nq, MvPlan &P) {
    if (nq < 1 || nq > SD_MV_MAXQ) return false;
    int best = 1 << 30;
    for (int qpt = 4; qpt >= 2; --qpt)                              // ties go to the wider per-thread run
        for (int e = 0; e <= 5; ++e) {
            const int nqp = qpt << e;
            if (nqp >= nq && nqp < best && nqp <= SD_MV_MAXQ) { best = nqp; P.qpt = qpt; P.lps_log2 = e; }
        }
    if (best == (1 << 30)) return false;
    P.nq = nq; P.nqp = best;
    const uint64_t thr = N << P.lps_log2;
    P.grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((thr + SD_MV_THREADS - 1) / SD_MV_THREADS, (uint64_t)c->sm_count * 8));
    return true;
}
struct DevBuf {                                                      // cudaFree on scope exit (error paths included)
    std::vector<void *> p;
    ~DevBuf() { for (void *q : p) cudaFree(q); }
    int make(void **out, size_t bytes) {
        *out = nullptr;
        cudaError_t e = cudaMalloc(out, bytes);
        if (e != cudaSuccess) return sd_fail(e == cudaErrorMemoryAllocation ? SD_ERR_NOMEM : SD_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        p.push_back(*out);
        return SD_OK;
    }
};
SdMvModel mv_model(const sd_model *m) {
    SdMvModel G;
    G.L = m->L; G.k = m->k; G.nhop = (int)m->hop_a.size(); G.nzz = (int)m->zz_a.size(); G.N = m->N;
    G.hop_a = m->d_hop_a; G.hop_b = m->d_hop_b; G.hop_J = m->d_hop_J;
    G.zz_a = m->d_zz_a; G.zz_b = m->d_zz_b; G.zz_J = m->d_zz_J; G.field = m->d_field;
    G.binom = m->ctx->d_binom; G.lin_h = m->lin_h; G.linA = m->d_linA; G.linB = m->d_linB;
    return G;
}
#define SD_MV_QPT(P_, CALL_)                          \
    do {                                              \
        if ((P_).qpt == 4) { constexpr int QPT = 4; CALL_; } \
        else if ((P_).qpt == 3) { constexpr int QPT = 3; CALL_; } \
        else { constexpr int QPT = 2; CALL_; }        \
    } while (0)

// common front end: checks, the plan, psi0 in rank order, phi_c = S^z_{q_c} psi0 into `phi`; hist rows 0 / 1 of step 0 =
// (unused, ||phi_c||^2).  Buffers: three multi-vectors mv[0..2], hist[(steps + 1) * 2 * nqp], partials, ticket, phases.
struct MvWork {
    MvPlan P;
    DevBuf bufs;
    double *mv[3] = {nullptr, nullptr, nullptr};
    double *hist = nullptr, *partials = nullptr, *ph = nullptr;
    unsigned *ticket = nullptr;
    SdMvRed red(int step, int slot) const {
        SdMvRed R;
        R.partials = partials; R.ticket = ticket; R.result = hist + ((size_t)step * 2 + slot) * P.nqp;
        return R;
    }
    const double *row(int step, int slot) const { return hist + ((size_t)step * 2 + slot) * P.nqp; }
};
int mv_begin(sd_model *m, const sd_vec *psi0, const double *q, int nq, int steps, MvWork &W) {
    sd_ctx *c = m->ctx;
    SD_ARG(c->world == 1, "the q-batched recurrences are single-GPU: loop over q on a sharded model");
    SD_ARG(psi0->model == m, "psi0 belongs to a different model");
    SD_ARG(mv_plan(c, m->N, nq, W.P), "1 <= nq <= %d momenta per batch", SD_MV_MAXQ);
    const MvPlan &P = W.P;
    const size_t mvbytes = (size_t)m->N * P.nqp * 2 * sizeof(double);
    for (int i = 0; i < 3; ++i) SD_TRY(W.bufs.make((void **)&W.mv[i], mvbytes));
    SD_TRY(W.bufs.make((void **)&W.hist, (size_t)(steps + 2) * 2 * P.nqp * sizeof(double)));
    SD_TRY(W.bufs.make((void **)&W.partials, (size_t)P.grid * SD_MV_NS * P.nqp * sizeof(double)));
    SD_TRY(W.bufs.make((void **)&W.ticket, 64));
    SD_TRY(W.bufs.make((void **)&W.ph, (size_t)P.nqp * m->L * 2 * sizeof(double)));
    SD_CUDA(cudaMemsetAsync(W.ticket, 0, 64, c->stream));
    SD_CUDA(cudaMemsetAsync(W.hist, 0, (size_t)(steps + 2) * 2 * P.nqp * sizeof(double), c->stream));
    std::vector<double> ph((size_t)P.nqp * m->L * 2, 0.0);
    for (int cidx = 0; cidx < nq; ++cidx)
        for (int r = 0; r < m->L; ++r) {                              // phases = exp.(im * q * (0:L-1)), Hamiltonian.jl:316
            ph[((size_t)cidx * m->L + r) * 2 + 0] = cos(q[cidx] * (double)r);
            ph[((size_t)cidx * m->L + r) * 2 + 1] = sin(q[cidx] * (double)r);
        }
    SD_CUDA(cudaMemcpyAsync(W.ph, ph.data(), ph.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));   // pageable: staged
    const double *src = psi0->d;
    if (psi0->layout) {                                               // block layout -> rank order (one pass)
        double *st = nullptr;
        SD_TRY(sd_scratch(c, 1, (size_t)psi0->logical_n * psi0->nc * sizeof(double) + 16, &st));
        SD_TRY(sd_blk_permute(psi0, st, psi0->nc, 1, 0, 0, 0.0));
        src = st;
    }
    SdMvSzq Z;
    Z.L = m->L; Z.k = m->k; Z.lps_log2 = P.lps_log2; Z.nqp = P.nqp; Z.ncin = psi0->nc; Z.N = m->N;
    Z.normfact = 1.0 / sqrt((double)m->L); Z.binom = c->d_binom; Z.ph = W.ph; Z.psi0 = src; Z.phi = W.mv[0];
    Z.R = W.red(0, 0);
    SD_MV_QPT(P, (sd_mv_szq_kernel<QPT><<<P.grid, SD_MV_THREADS, 0, c->stream>>>(Z)));
    return sd_launch_check(c, "sd_mv_szq_kernel");
}
int mv_fetch(sd_ctx *c, const MvWork &W, int rows, std::vector<double> &h) {
    h.resize((size_t)rows * W.P.nqp);
    SD_CUDA(cudaMemcpyAsync(h.data(), W.hist, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SD_CUDA(cudaStreamSynchronize(c->stream));
    return SD_OK;
}
}  // namespace

// lanczos_tridiag (Lanczos.jl:196-246) of phi_q = S^z_q psi0 for nq momenta at once: the body of the q-loop of
// lanczos_sqw (LanczosSqw.jl:65-77) up to the tridiagonal matrix.  alpha[c * lanc_m + t], beta[c * lanc_m + t],
// m_eff[c] (0: norm(phi_c) == 0, the reference's `continue`), norm_phi[c].
int sd_lanczos_tridiag_szq_batch(sd_model *m, const sd_vec *psi0, const double *q, int nq, int lanc_m, double tol,
                                 double *alpha, double *beta, int *m_eff, double *norm_phi) {
    SD_ARG(m && psi0 && q && alpha && beta && m_eff && norm_phi, "NULL argument");
    SD_ARG(lanc_m >= 1, "lanc_m must be >= 1");
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    const int mm = (int)std::min<uint64_t>((uint64_t)lanc_m, m->N);          // Lanczos.jl:200
    SD_ARG(mm <= SD_HIST_MAX, "lanc_m must be <= %d", SD_HIST_MAX);
    MvWork W;
    SD_TRY(mv_begin(m, psi0, q, nq, mm, W));
    const MvPlan &P = W.P;
    double *u = W.mv[0], *uo = W.mv[1], *w = W.mv[2];
    // deferred normalisation exactly as sd_lanczos_engine: u_1 = phi, u_{j+1} = w_j, v_j = u_j / beta_{j-1};
    // step j: row (j, 0) = d_j = <u_j, H u_j> / beta_{j-1}, row (j, 1) = n_j = ||w_j||^2; n_0 = ||phi||^2
    for (int j = 1; j <= mm; ++j) {
        SdMvApply A;
        A.G = mv_model(m); A.lps_log2 = P.lps_log2; A.nqp = P.nqp; A.u = u; A.w = w; A.vprev = nullptr; A.phi = nullptr;
        A.n_prev = W.row(j - 1, 1); A.a = 1.0; A.b = 0.0; A.R = W.red(j, 0);
        SD_MV_QPT(P, (sd_mv_apply_kernel<QPT, 1><<<P.grid, SD_MV_THREADS, 0, c->stream>>>(A)));
        SD_TRY(sd_launch_check(c, "sd_mv_apply_kernel"));
        if (j < mm) {
            SdMvUpdate U;
            U.lps_log2 = P.lps_log2; U.nqp = P.nqp; U.N = m->N; U.w = w; U.u = u; U.uo = j >= 2 ? uo : nullptr;
            U.d = W.row(j, 0); U.n1 = W.row(j - 1, 1); U.n2 = j >= 2 ? W.row(j - 2, 1) : nullptr; U.R = W.red(j, 1);
            SD_MV_QPT(P, (sd_mv_update_kernel<QPT><<<P.grid, SD_MV_THREADS, 0, c->stream>>>(U)));
            SD_TRY(sd_launch_check(c, "sd_mv_update_kernel"));
        }
        double *t = uo; uo = u; u = w; w = t;
    }
    std::vector<double> h;
    SD_TRY(mv_fetch(c, W, 2 * (mm + 1), h));
    auto H = [&](int step, int slot, int col) { return h[((size_t)step * 2 + slot) * P.nqp + col]; };
    for (int cidx = 0; cidx < nq; ++cidx) {
        const double n0 = H(0, 1, cidx);
        norm_phi[cidx] = sqrt(n0);
        for (int t = 0; t < lanc_m; ++t) alpha[(size_t)cidx * lanc_m + t] = beta[(size_t)cidx * lanc_m + t] = 0.0;
        if (n0 == 0.0) { m_eff[cidx] = 0; continue; }                 // LanczosSqw.jl:69-72
        int eff = mm;
        for (int t = 1; t <= mm; ++t) {                               // Lanczos.jl:218-231 / sd_lanczos_engine
            alpha[(size_t)cidx * lanc_m + t - 1] = H(t, 0, cidx) / sqrt(H(t - 1, 1, cidx));
            if (t < mm) {
                const double bt = sqrt(H(t, 1, cidx));
                beta[(size_t)cidx * lanc_m + t - 1] = bt;
                if (bt < tol) { eff = t; break; }
            }
        }
        m_eff[cidx] = eff;
    }
    return SD_OK;
}

// compute_chebyshev_moments (KPM_Sqw.jl:95-128) of phi_q / ||phi_q|| for nq momenta at once: the body of the q-loop of
// kpm_sqw (:218-253) up to the moments.  mu[c * M + n]; norm_phi[c] (0: the reference's `continue`, mu row = 0).
// *blown = 1 if some ||v_next|| exceeded 1e3 (:117-121, wrong rescaling bounds): the moments are then NOT those of the
// reference, which renormalises there -- the caller falls back to the per-q path.
int sd_kpm_moments_szq_batch(sd_model *m, const sd_vec *psi0, const double *q, int nq, int M, double a, double b,
                             double *mu, double *norm_phi, int *blown) {
    SD_ARG(m && psi0 && q && mu && norm_phi && blown, "NULL argument");
    SD_ARG(M >= 1 && M <= 1 << 20, "M must be in 1 .. 2^20");
    sd_ctx *c = m->ctx;
    SD_LOCK(c); SD_TRY(sd_use(c));
    MvWork W;
    SD_TRY(mv_begin(m, psi0, q, nq, M, W));
    const MvPlan &P = W.P;
    double *phi = W.mv[0], *b1 = W.mv[1], *b2 = W.mv[2];
    // hist rows: (0, 1) = ||phi_c||^2 from mv_begin; (0, 0) = mu_0 after the normalisation; (n, 0) = mu_n, (n, 1) = ||v_n||^2
    {
        SdMvScale S;
        S.lps_log2 = P.lps_log2; S.nqp = P.nqp; S.N = m->N; S.phi = phi; S.n0 = W.row(0, 1); S.R = W.red(0, 0);
        SD_MV_QPT(P, (sd_mv_colnorm_kernel<QPT><<<P.grid, SD_MV_THREADS, 0, c->stream>>>(S)));
        SD_TRY(sd_launch_check(c, "sd_mv_colnorm_kernel"));
    }
    SdMvApply A;
    A.G = mv_model(m); A.lps_log2 = P.lps_log2; A.nqp = P.nqp; A.n_prev = nullptr; A.a = a; A.b = b; A.phi = phi;
    if (M >= 2) {                                                     // :106-107
        A.u = phi; A.w = b1; A.vprev = nullptr; A.R = W.red(1, 0);
        SD_MV_QPT(P, (sd_mv_apply_kernel<QPT, 2><<<P.grid, SD_MV_THREADS, 0, c->stream>>>(A)));
        SD_TRY(sd_launch_check(c, "sd_mv_apply_kernel"));
    }
    // v_2 goes to the third buffer (v_prev = phi is kept for the dots); from n = 3 on v_next overwrites v_prev in place:
    // the kernel reads v_prev only at the element it writes
    const double *vp = phi;
    double *vc = b1;
    for (int n = 2; n < M; ++n) {                                     // :109-126
        double *vn = (n == 2) ? b2 : const_cast<double *>(vp);
        A.u = vc; A.w = vn; A.vprev = vp; A.R = W.red(n, 0);
        SD_MV_QPT(P, (sd_mv_apply_kernel<QPT, 3><<<P.grid, SD_MV_THREADS, 0, c->stream>>>(A)));
        SD_TRY(sd_launch_check(c, "sd_mv_apply_kernel"));
        vp = vc; vc = vn;
    }
    std::vector<double> h;
    SD_TRY(mv_fetch(c, W, 2 * (M + 1), h));
    auto H = [&](int step, int slot, int col) { return h[((size_t)step * 2 + slot) * P.nqp + col]; };
    *blown = 0;
    for (int cidx = 0; cidx < nq; ++cidx) {
        norm_phi[cidx] = sqrt(H(0, 1, cidx));
        for (int n = 0; n < M; ++n) mu[(size_t)cidx * M + n] = (norm_phi[cidx] == 0.0) ? 0.0 : H(n, 0, cidx);
        for (int n = 2; n < M; ++n) if (norm_phi[cidx] != 0.0 && !(sqrt(H(n, 1, cidx)) <= 1e3)) *blown = 1;
    }
    return SD_OK;
}

int sd_ctx_mem_info(sd_ctx *c, uint64_t *free_bytes, uint64_t *total_bytes) {
    SD_ARG(c && free_bytes && total_bytes, "NULL argument");
    SD_LOCK(c); SD_TRY(sd_use(c));
    size_t f = 0, t = 0;
    SD_CUDA(cudaMemGetInfo(&f, &t));
    *free_bytes = f; *total_bytes = t;
    return SD_OK;
}
