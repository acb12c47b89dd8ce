/*
 * oracle.c -- CPU restatement of SpinDynamics.jl's matrix-free H.psi path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under spindynamics.jl_b200/ may link,
 * load or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker
 * or as the timed CPU baseline.
 *
 * The reference is pure Julia and `julia` is not installed in this image, so
 * the reference itself cannot be executed here.  Every
 * function below restates one reference function; citations are file:line
 * relative to /root/reference.  Pins: tests/test_oracle.py checks this file
 * against the reference tests' known answers (test_PublicAPI.jl:5-28 L=2
 * matrix, test_Basis.jl:11-18) and against an independent Kronecker-product
 * dense construction of the XXZ Hamiltonian.  The ordering *inside* a sector
 * is whatever Combinatorics.jl's `combinations` yields (lexicographic in the
 * ascending site lists); no reference test pins it -> "parity unpinned" for
 * that one property (it is pinned against Python's itertools.combinations,
 * which documents the same order).
 *
 * Data layout mirrors Julia: Vector{Float64} = double[], Vector{ComplexF64} =
 * interleaved (re,im) double pairs, Vector{Tuple{Int,Int,Float64}} = 24-byte
 * records {int64 i; int64 j; double J} with 1-based sites.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <complex.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { int64_t i, j; double J; } orc_bond;   /* 1-based sites */

typedef struct {
    int L;
    int nup;              /* -1 = full basis (mode :full)      SpinModel.jl:6-15 */
    uint64_t N;           /* length(model.states)                                */
    uint64_t *states;     /* model.states                                        */
    /* model.idxmap :: Dict{UInt64,Int}; here an open-addressing table         */
    uint64_t *keys;       /* key+1 (0 = empty slot)                              */
    int64_t  *vals;       /* 1-based index                                       */
    uint64_t cap_mask;
    int nhop, nzz;
    orc_bond *hop, *zz;   /* hopping_list, zz_list                               */
    double *field;        /* onsite_field, length L                              */
} orc_model;

/* ------------------------------------------------------------------ helpers */

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31; return x;
}

/* Hamiltonian.jl:19-29 */
static inline uint64_t bit_at(uint64_t state, int i) { return (state >> i) & 1ULL; }
static inline double   sz_value(uint64_t bit)        { return bit == 1 ? 0.5 : -0.5; }
static inline uint64_t flip_bits(uint64_t s, int i, int j) {
    return s ^ (1ULL << i) ^ (1ULL << j);
}

/* The table stands in for Julia's Dict{UInt64,Int}: Base.Dict keeps its slot count a power of two and grows when
 * more than 2/3 of the slots are in use, so N keys occupy the smallest power of two >= 1.5 N slots (17 bytes per slot
 * there, 16 here).  L = 32, nup = 16: 2^30 slots = 17 GB.  Large bases are inserted by all OpenMP threads (compare-
 * and-swap on the key slot); the result does not depend on the insertion order because keys are unique. */
static void map_build(orc_model *m) {
    uint64_t cap = 16;
    while (2 * cap < 3 * m->N) cap <<= 1;
    m->cap_mask = cap - 1;
    m->keys = (uint64_t *)calloc(cap, sizeof(uint64_t));
    m->vals = (int64_t *)malloc(cap * sizeof(int64_t));
    const int64_t N = (int64_t)m->N;
#pragma omp parallel for schedule(static) if (N > (1 << 22))
    for (int64_t i = 0; i < N; ++i) {               /* Basis.jl:49-52, idxmap[s] = i */
        uint64_t s = m->states[i];
        uint64_t h = mix64(s) & m->cap_mask;
        for (;;) {
            uint64_t expect = 0;
            if (__atomic_compare_exchange_n(&m->keys[h], &expect, s + 1, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) break;
            h = (h + 1) & m->cap_mask;
        }
        m->vals[h] = (int64_t)i + 1;
    }
}

/* get(model.idxmap, s, 0)   Hamiltonian.jl:260 */
static inline int64_t map_get(const orc_model *m, uint64_t s) {
    uint64_t h = mix64(s) & m->cap_mask;
    for (;;) {
        uint64_t k = m->keys[h];
        if (k == s + 1) return m->vals[h];
        if (k == 0) return 0;
        h = (h + 1) & m->cap_mask;
    }
}

static uint64_t binom_u64(int n, int k) {
    if (k < 0 || k > n) return 0;
    if (k > n - k) k = n - k;
    unsigned __int128 r = 1;
    for (int i = 1; i <= k; ++i) r = r * (unsigned)(n - k + i) / (unsigned)i;
    return (uint64_t)r;
}

/* ---------------------------------------------------------------- basis */

/* Basis.jl:9-20  _validate_basis_args; returns 0 ok, -1 ArgumentError */
int orc_validate_basis_args(int L, int nup /* -1 = nothing */) {
    if (L < 1) return -1;
    if (L > 63) return -1;
    if (nup != -1 && (nup < 0 || nup > L)) return -1;
    return 0;
}

uint64_t orc_sector_dim(int L, int nup) { return binom_u64(L, nup); }

/* Basis.jl:37-53  build_sector_basis: `for comb in combinations(1:L, nup)`,
 * site i sets bit i-1, states pushed in iteration order.  Combinatorics.jl
 * yields the k-subsets in lexicographic order of their ascending site lists.
 * out must hold binomial(L,nup) entries. */
static void sector_basis_range(int L, int nup, uint64_t first, uint64_t count, uint64_t *out) {
    /* the `first`-th combination (0-based) of combinations(1:L, nup), then `count` successors in iteration order */
    int c[64];
    uint64_t idx = first;
    int prev = 0;
    for (int i = 0; i < nup; ++i) {
        int v = prev + 1;
        for (;; ++v) {
            uint64_t below = binom_u64(L - v, nup - 1 - i);        /* combinations whose i-th site is v */
            if (idx < below) break;
            idx -= below;
        }
        c[i] = v; prev = v;
    }
    for (uint64_t n = 0; n < count; ++n) {
        uint64_t s = 0;
        for (int i = 0; i < nup; ++i) s |= 1ULL << (c[i] - 1);
        out[n] = s;
        int i = nup - 1;
        while (i >= 0 && c[i] == L - nup + i + 1) --i;
        if (i < 0) break;
        ++c[i];
        for (int j = i + 1; j < nup; ++j) c[j] = c[j - 1] + 1;
    }
}
int orc_build_sector_basis(int L, int nup, uint64_t *out) {
    if (orc_validate_basis_args(L, nup) || nup < 0) return -1;
    const uint64_t N = binom_u64(L, nup);
    if (N > (1u << 22)) {              /* same enumeration, cut into chunks that start at an unranked combination */
        const int64_t nchunk = 1024;
#pragma omp parallel for schedule(dynamic)
        for (int64_t ch = 0; ch < nchunk; ++ch) {
            const uint64_t lo = (uint64_t)(((unsigned __int128)N * (uint64_t)ch) / (uint64_t)nchunk);
            const uint64_t hi = (uint64_t)(((unsigned __int128)N * (uint64_t)(ch + 1)) / (uint64_t)nchunk);
            if (hi > lo) sector_basis_range(L, nup, lo, hi - lo, out + lo);
        }
        return 0;
    }
    int c[64];
    for (int i = 0; i < nup; ++i) c[i] = i + 1;      /* first combination 1..k */
    uint64_t n = 0;
    for (;;) {
        uint64_t s = 0;
        for (int i = 0; i < nup; ++i) s |= 1ULL << (c[i] - 1);
        out[n++] = s;
        int i = nup - 1;
        while (i >= 0 && c[i] == L - nup + i + 1) --i;
        if (i < 0) break;
        ++c[i];
        for (int j = i + 1; j < nup; ++j) c[j] = c[j - 1] + 1;
    }
    return 0;
}

/* Basis.jl:23-34  build_full_basis: states[i+1] = i */
int orc_build_full_basis(int L, uint64_t *out) {
    if (orc_validate_basis_args(L, -1)) return -1;
    uint64_t N = 1ULL << L;
    for (uint64_t i = 0; i < N; ++i) out[i] = i;
    return 0;
}

/* ---------------------------------------------------------------- model */

/* SpinModel.jl:23-38 build_model (states + idxmap + copies of the lists) */
orc_model *orc_model_create(int L, int nup, const orc_bond *hop, int nhop,
                            const orc_bond *zz, int nzz, const double *field) {
    if (orc_validate_basis_args(L, nup)) return NULL;
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    m->L = L; m->nup = nup;
    m->N = nup < 0 ? (1ULL << L) : binom_u64(L, nup);
    m->states = (uint64_t *)malloc(m->N * sizeof(uint64_t));
    if (nup < 0) orc_build_full_basis(L, m->states);
    else         orc_build_sector_basis(L, nup, m->states);
    map_build(m);
    m->nhop = nhop; m->nzz = nzz;
    m->hop = (orc_bond *)malloc((nhop ? nhop : 1) * sizeof(orc_bond));
    m->zz  = (orc_bond *)malloc((nzz ? nzz : 1) * sizeof(orc_bond));
    if (nhop) memcpy(m->hop, hop, nhop * sizeof(orc_bond));
    if (nzz)  memcpy(m->zz, zz, nzz * sizeof(orc_bond));
    m->field = (double *)malloc(L * sizeof(double));
    memcpy(m->field, field, L * sizeof(double));
    return m;
}

void orc_model_free(orc_model *m) {
    if (!m) return;
    free(m->states); free(m->keys); free(m->vals);
    free(m->hop); free(m->zz); free(m->field); free(m);
}

uint64_t orc_model_dim(const orc_model *m) { return m->N; }
const uint64_t *orc_model_states(const orc_model *m) { return m->states; }

/* get(model.idxmap, s, 0) for an array of states; 1-based, 0 = absent */
void orc_rank(const orc_model *m, const uint64_t *s, uint64_t n, int64_t *idx1) {
    for (uint64_t i = 0; i < n; ++i) idx1[i] = map_get(m, s[i]);
}

/* bench.py: the OpenMP team is set explicitly (torchrun exports OMP_NUM_THREADS=1 to its children) */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* -------------------------------------------------------------- apply_H! */

/* Hamiltonian.jl:211-273, T = Float64.  One output per iteration
 * (Threads.@threads for idx in 1:N  ->  omp parallel for, static schedule). */
void orc_apply_H_f64(const orc_model *m, double *out, const double *psi) {
    const int L = m->L;
    const int full = m->nup < 0;
    const int64_t N = (int64_t)m->N;
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < N; ++idx) {
        uint64_t state = full ? (uint64_t)idx : m->states[idx];     /* :223 */
        double diag = 0.0;
        for (int i = 1; i <= L; ++i)                                 /* :228-233 */
            diag += m->field[i - 1] * sz_value(bit_at(state, i - 1));
        for (int b = 0; b < m->nzz; ++b) {                           /* :235-241 */
            const orc_bond *z = &m->zz[b];
            diag += z->J * sz_value(bit_at(state, (int)z->i - 1))
                         * sz_value(bit_at(state, (int)z->j - 1));
        }
        double value = diag * psi[idx];                              /* :243 */
        for (int b = 0; b < m->nhop; ++b) {                          /* :248-267 */
            const orc_bond *h = &m->hop[b];
            uint64_t bi = bit_at(state, (int)h->i - 1);
            uint64_t bj = bit_at(state, (int)h->j - 1);
            if (bi != bj) {
                uint64_t ns = flip_bits(state, (int)h->i - 1, (int)h->j - 1);
                if (full) {
                    value += h->J * psi[ns];                         /* :255-257 */
                } else {
                    int64_t ni = map_get(m, ns);                     /* :260 */
                    if (ni != 0) value += h->J * psi[ni - 1];
                }
            }
        }
        out[idx] = value;                                            /* :269 */
    }
}

/* NOT the reference's algorithm: the "second, stronger CPU baseline" of BASELINE.md.  Same loop as above for the
 * sector basis, but the Dict probe `get(idxmap, new_state, 0)` (:260) is replaced by combinatorial ranking of the
 * flipped state (orc_rank_closed_form's sum, evaluated incrementally: a hop that swaps bits a < b of the state moves
 * the rank by a sum of binomials over the positions a..b), so no hash table is touched.  Used only by bench.py as a
 * reported figure and checked against orc_apply_H_f64 in tests/test_oracle.py. */
void orc_apply_H_f64_ranked(const orc_model *m, double *out, const double *psi) {
    const int L = m->L;
    if (m->nup < 0) { orc_apply_H_f64(m, out, psi); return; }
    uint64_t C[65][65];
    for (int n = 0; n < 65; ++n)
        for (int r = 0; r < 65; ++r) C[n][r] = (r > n) ? 0 : binom_u64(n, r);
    const int64_t N = (int64_t)m->N;
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < N; ++idx) {
        const uint64_t state = m->states[idx];
        double diag = 0.0;
        for (int i = 1; i <= L; ++i) diag += m->field[i - 1] * sz_value(bit_at(state, i - 1));
        for (int b = 0; b < m->nzz; ++b) {
            const orc_bond *z = &m->zz[b];
            diag += z->J * sz_value(bit_at(state, (int)z->i - 1)) * sz_value(bit_at(state, (int)z->j - 1));
        }
        double value = diag * psi[idx];
        for (int b = 0; b < m->nhop; ++b) {
            const orc_bond *h = &m->hop[b];
            int pa = (int)h->i - 1, pb = (int)h->j - 1;
            if (pa > pb) { int t = pa; pa = pb; pb = t; }
            const uint64_t ba = bit_at(state, pa), bb = bit_at(state, pb);
            if (ba == bb) continue;
            /* rank = sum over clear bits q of C(L-1-q, rem(q)-1), rem(q) = set bits at positions >= q (of the
             * remaining ones); only positions pa..pb change their contribution */
            const uint64_t ns = flip_bits(state, pa, pb);
            int rem_old = __builtin_popcountll(state >> pa), rem_new = __builtin_popcountll(ns >> pa);
            int64_t delta = 0;
            for (int q = pa; q <= pb; ++q) {
                if ((state >> q) & 1ULL) --rem_old; else if (rem_old > 0) delta -= (int64_t)C[L - 1 - q][rem_old - 1];
                if ((ns >> q) & 1ULL) --rem_new; else if (rem_new > 0) delta += (int64_t)C[L - 1 - q][rem_new - 1];
            }
            value += h->J * psi[idx + delta];
        }
        out[idx] = value;
    }
}

/* Hamiltonian.jl:211-273, T = ComplexF64 (interleaved re,im). */
void orc_apply_H_c128(const orc_model *m, double *out, const double *psi) {
    const int L = m->L;
    const int full = m->nup < 0;
    const int64_t N = (int64_t)m->N;
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < N; ++idx) {
        uint64_t state = full ? (uint64_t)idx : m->states[idx];
        double diag = 0.0;
        for (int i = 1; i <= L; ++i)
            diag += m->field[i - 1] * sz_value(bit_at(state, i - 1));
        for (int b = 0; b < m->nzz; ++b) {
            const orc_bond *z = &m->zz[b];
            diag += z->J * sz_value(bit_at(state, (int)z->i - 1))
                         * sz_value(bit_at(state, (int)z->j - 1));
        }
        double vr = diag * psi[2 * idx], vi = diag * psi[2 * idx + 1];
        for (int b = 0; b < m->nhop; ++b) {
            const orc_bond *h = &m->hop[b];
            uint64_t bi = bit_at(state, (int)h->i - 1);
            uint64_t bj = bit_at(state, (int)h->j - 1);
            if (bi != bj) {
                uint64_t ns = flip_bits(state, (int)h->i - 1, (int)h->j - 1);
                int64_t ni = full ? (int64_t)ns + 1 : map_get(m, ns);
                if (ni != 0) {
                    vr += h->J * psi[2 * (ni - 1)];
                    vi += h->J * psi[2 * (ni - 1) + 1];
                }
            }
        }
        out[2 * idx] = vr; out[2 * idx + 1] = vi;
    }
}

/* Hamiltonian.jl:286-301 apply_rescaled_H!: out = (H psi - b psi)/a, true
 * division, serial second pass.  `cplx` selects ComplexF64. */
void orc_apply_rescaled_H(const orc_model *m, double *out, const double *psi,
                          double a, double b, int cplx) {
    if (cplx) orc_apply_H_c128(m, out, psi); else orc_apply_H_f64(m, out, psi);
    uint64_t n = cplx ? 2 * m->N : m->N;
    for (uint64_t i = 0; i < n; ++i) out[i] = (out[i] - b * psi[i]) / a;
}

/* Hamiltonian.jl:307-337 Sz_q_vector: phi[idx] = L^-1/2 (sum_r e^{iqr} s_r) psi0[idx];
 * phases = exp.(im*q*(0:L-1)); psi0 real (cplx=0) or complex (cplx=1); phi complex. */
void orc_szq(const orc_model *m, double *phi, const double *psi0, int cplx, double q) {
    const int L = m->L;
    const int full = m->nup < 0;
    const int64_t N = (int64_t)m->N;
    const double normfact = 1.0 / sqrt((double)L);
    double complex phases[64];
    for (int r = 0; r < L; ++r) phases[r] = cexp(I * q * (double)r);
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < N; ++idx) {
        uint64_t state = full ? (uint64_t)idx : m->states[idx];
        double complex sq = 0.0;
        for (int r = 0; r < L; ++r) sq += phases[r] * sz_value(bit_at(state, r));
        double complex p = cplx ? (psi0[2 * idx] + I * psi0[2 * idx + 1])
                                : (double complex)psi0[idx];
        double complex v = normfact * sq * p;
        phi[2 * idx] = creal(v); phi[2 * idx + 1] = cimag(v);
    }
}

/* ------------------------------------------------ bench-input generator ----
 * Not from the reference: the counter-based synthetic psi of SURVEY.md 8(d),
 * psi[r] = 2*u(splitmix64(seed ^ r)) - 1 with u from the top 53 bits.  The
 * CUDA library implements the same formula (sd_vec_fill_seeded) so CPU and
 * GPU can regenerate any element without storing psi. */
static inline double seeded_value(uint64_t seed, uint64_t r) {
    uint64_t z = (seed ^ r) + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    z ^= z >> 31;
    return 2.0 * ((double)(z >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}

void orc_fill_seeded(double *v, uint64_t first, uint64_t n, uint64_t seed, int cplx) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        if (cplx) {
            v[2 * i]     = seeded_value(seed, first + i);
            v[2 * i + 1] = seeded_value(seed + 1, first + i);
        } else {
            v[i] = seeded_value(seed, first + i);
        }
    }
}

/* One row of H.psi for a state given explicitly, with psi supplied by the
 * seeded generator: lets the tests check sampled rows of an L=32..36 apply
 * without materialising states[], the Dict or psi on the host.  Ranking of the
 * hop targets uses the closed form of the combinations order (validated
 * against orc_build_sector_basis for every (L<=16,nup) in tests). */
uint64_t orc_rank_closed_form(int L, int nup, uint64_t s) {
    /* idx0 = sum over chosen sites c_t (ascending, 0-based) of the number of
     * combinations that precede because they pick a smaller site at slot t. */
    uint64_t idx = 0; int t = 0; int prev = -1;
    for (int p = 0; p < L && t < nup; ++p) {
        if ((s >> p) & 1ULL) {
            for (int j = prev + 1; j < p; ++j) idx += binom_u64(L - 1 - j, nup - 1 - t);
            prev = p; ++t;
        }
    }
    return idx;
}

double orc_row_seeded_f64(int L, int nup, const orc_bond *hop, int nhop,
                          const orc_bond *zz, int nzz, const double *field,
                          uint64_t state, uint64_t seed, double scale) {
    double diag = 0.0;
    for (int i = 1; i <= L; ++i) diag += field[i - 1] * sz_value(bit_at(state, i - 1));
    for (int b = 0; b < nzz; ++b)
        diag += zz[b].J * sz_value(bit_at(state, (int)zz[b].i - 1))
                        * sz_value(bit_at(state, (int)zz[b].j - 1));
    uint64_t idx = nup < 0 ? state : orc_rank_closed_form(L, nup, state);
    double value = diag * (scale * seeded_value(seed, idx));
    for (int b = 0; b < nhop; ++b) {
        uint64_t bi = bit_at(state, (int)hop[b].i - 1), bj = bit_at(state, (int)hop[b].j - 1);
        if (bi != bj) {
            uint64_t ns = flip_bits(state, (int)hop[b].i - 1, (int)hop[b].j - 1);
            uint64_t ni = nup < 0 ? ns : orc_rank_closed_form(L, nup, ns);
            value += hop[b].J * (scale * seeded_value(seed, ni));
        }
    }
    return value;
}
