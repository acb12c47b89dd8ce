// mb_streams.cu -- microbenchmark of the phase-1 access pattern of the tiled H.psi kernel:
// out[l] = sum_k J * in[l + off_k] over K shifted coalesced streams (some beyond L2 reach, some
// within), to find what the B200 memory system sustains for this pattern independent of the rest
// of the kernel.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb_streams mb_streams.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Offs { long long off[16]; int k; };

// A: flat grid-stride, E elements per thread, K loads per element in flight
template <int E>
__global__ void __launch_bounds__(256) flat_kernel(const double *__restrict__ in, double *__restrict__ out, long long n, Offs o) {
    const long long stride = (long long)gridDim.x * blockDim.x * E;
    for (long long i0 = ((long long)blockIdx.x * blockDim.x) * E + threadIdx.x; i0 < n; i0 += stride) {
        double s[E];
#pragma unroll
        for (int e = 0; e < E; ++e) s[e] = 0.0;
        for (int k = 0; k < o.k; ++k) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                long long i = i0 + (long long)e * blockDim.x;
                if (i < n) { long long j = i + o.off[k]; if (j >= n) j -= n; if (j < 0) j += n; s[e] += 0.5 * __ldg(in + j); }
            }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) { long long i = i0 + (long long)e * blockDim.x; if (i < n) __stcs(out + i, s[e]); }
    }
}

// B: tile kernel: one CTA per tile of TS elements, 512 threads, rolling double buffer of U loads
template <int U, int DEPTH, int NT = 512>
__global__ void __launch_bounds__(NT, 1) tile_kernel(const double *__restrict__ in, double *__restrict__ out, long long n, int ts, Offs o, int use_smem) {
    extern __shared__ double sm[];
    const long long base = (long long)blockIdx.x * ts;
    if (base >= n) return;
    const int size = (int)((n - base) < ts ? (n - base) : ts);
    for (int l0 = threadIdx.x; l0 < size; l0 += U * NT) {
        double g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) g[u] = 0.0;
        if (DEPTH == 1) {
            for (int k = 0; k < o.k; ++k) {
                long long b = base + o.off[k]; if (b >= n) b -= n; if (b < 0) b += n;
                if (b + ts > n) b = 0;
                const double *q = in + b + l0;
                double t[U];
#pragma unroll
                for (int u = 0; u < U; ++u) t[u] = (l0 + u * NT < size) ? __ldg(q + u * NT) : 0.0;
#pragma unroll
                for (int u = 0; u < U; ++u) g[u] += 0.5 * t[u];
            }
        } else {
            double t0[U], t1[U];
            auto ptr = [&](int k) { long long b = base + o.off[k]; if (b >= n) b -= n; if (b < 0) b += n; if (b + ts > n) b = 0; return in + b + l0; };
            const double *q = ptr(0);
#pragma unroll
            for (int u = 0; u < U; ++u) t0[u] = (l0 + u * NT < size) ? __ldg(q + u * NT) : 0.0;
            int k = 0;
            for (; k + 2 < o.k; k += 2) {
                q = ptr(k + 1);
#pragma unroll
                for (int u = 0; u < U; ++u) t1[u] = (l0 + u * NT < size) ? __ldg(q + u * NT) : 0.0;
#pragma unroll
                for (int u = 0; u < U; ++u) g[u] += 0.5 * t0[u];
                q = ptr(k + 2);
#pragma unroll
                for (int u = 0; u < U; ++u) t0[u] = (l0 + u * NT < size) ? __ldg(q + u * NT) : 0.0;
#pragma unroll
                for (int u = 0; u < U; ++u) g[u] += 0.5 * t1[u];
            }
            if (k + 1 < o.k) {
                q = ptr(k + 1);
#pragma unroll
                for (int u = 0; u < U; ++u) t1[u] = (l0 + u * NT < size) ? __ldg(q + u * NT) : 0.0;
#pragma unroll
                for (int u = 0; u < U; ++u) g[u] += 0.5 * t0[u] + 0.5 * t1[u];
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) g[u] += 0.5 * t0[u];
            }
        }
        if (use_smem) {
#pragma unroll
            for (int u = 0; u < U; ++u) if (l0 + u * NT < size) sm[l0 + u * NT] = g[u];
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) if (l0 + u * NT < size) __stcs(out + base + l0 + u * NT, g[u]);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int l = threadIdx.x; l < size; l += NT) __stcs(out + base + l, sm[size - 1 - l] + sm[l]);
    }
}

// B2: like B (DEPTH 1) but warp-blocked mapping: a warp owns 32*U consecutive elements, lane's elements are 256 B apart
template <int U, int NT>
__global__ void __launch_bounds__(NT, 1) tileb_kernel(const double *__restrict__ in, double *__restrict__ out, long long n, int ts, Offs o, int use_smem) {
    extern __shared__ double sm[];
    const long long base = (long long)blockIdx.x * ts;
    if (base >= n) return;
    const int size = (int)((n - base) < ts ? (n - base) : ts);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c0 = warp * 32 * U; c0 < size; c0 += (NT / 32) * 32 * U) {
        const int l0 = c0 + lane;
        double g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) g[u] = 0.0;
        for (int k = 0; k < o.k; ++k) {
            long long b = base + o.off[k]; if (b >= n) b -= n; if (b < 0) b += n;
            if (b + ts > n) b = 0;
            const double *q = in + b + l0;
            double t[U];
#pragma unroll
            for (int u = 0; u < U; ++u) t[u] = (l0 + u * 32 < size) ? __ldg(q + u * 32) : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) g[u] += 0.5 * t[u];
        }
        if (use_smem) {
#pragma unroll
            for (int u = 0; u < U; ++u) if (l0 + u * 32 < size) sm[l0 + u * 32] = g[u];
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) if (l0 + u * 32 < size) __stcs(out + base + l0 + u * 32, g[u]);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int l = threadIdx.x; l < size; l += NT) __stcs(out + base + l, sm[size - 1 - l] + sm[l]);
    }
}

// C: TMA bulk (cp.async.bulk) staging: one elected thread copies each stream's slice into a smem ring,
// all threads accumulate from smem.  SL = slice elements, NS = ring slots.
__device__ __forceinline__ void mbar_init(uint64_t *b, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned phase) {
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
template <int SL, int NS>
__global__ void __launch_bounds__(512, 2) tma_kernel(const double *__restrict__ in, double *__restrict__ out, long long n, int ts, Offs o) {
    extern __shared__ __align__(128) unsigned char smraw[];
    double *ring = (double *)smraw;                       // [NS][SL]
    uint64_t *full = (uint64_t *)(smraw + (size_t)NS * SL * 8);
    const long long base = ((long long)blockIdx.x * ts) & ~1LL;      // 16-byte aligned tiles for the bulk copy
    if (base >= n) return;
    const int size = (int)((n - base) < ts ? (n - base) : ts) & ~1;
    const int nsl = (size + SL - 1) / SL;                 // slices per stream
    const int total = nsl * o.k;                          // (slice, stream) jobs, stream fastest
    if (threadIdx.x == 0) for (int s = 0; s < NS; ++s) mbar_init(full + s, 1);
    __syncthreads();
    auto issue = [&](int job) {
        const int sl = job / o.k, k = job - sl * o.k;
        long long b = base + o.off[k]; if (b >= n) b -= n; if (b < 0) b += n; if (b + ts > n) b = 0;
        b &= ~1LL;
        const int cnt = (size - sl * SL) < SL ? (size - sl * SL) : SL;
        const int slot = job % NS;
        mbar_expect(full + slot, cnt * 8);
        bulk_g2s(ring + (size_t)slot * SL, in + b + (long long)sl * SL, cnt * 8, full + slot);
    };
    if (threadIdx.x == 0) for (int j = 0; j < NS && j < total; ++j) issue(j);
    constexpr int EPT = SL / 512;
    double g[EPT];
    for (int job = 0; job < total; ++job) {
        const int sl = job / o.k, k = job - sl * o.k;
        const int slot = job % NS;
        if (k == 0) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) g[e] = 0.0;
        }
        mbar_wait(full + slot, (job / NS) & 1);
#pragma unroll
        for (int e = 0; e < EPT; ++e) g[e] += 0.5 * ring[(size_t)slot * SL + threadIdx.x + e * 512];
        __syncthreads();                                  // slot free
        if (threadIdx.x == 0 && job + NS < total) issue(job + NS);
        if (k == o.k - 1) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) { const int l = sl * SL + threadIdx.x + e * 512; if (l < size) __stcs(out + base + l, g[e]); }
        }
    }
}

// D: hybrid: far streams (k < kfar) arrive through TMA bulk copies into a smem ring, near streams through LDG
template <int SL, int KFAR, int NT>
__global__ void __launch_bounds__(NT, 1) hybrid_kernel(const double *__restrict__ in, double *__restrict__ out, long long n, int ts, Offs o) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int NS = 2 * KFAR;
    double *ring = (double *)smraw;                       // [NS][SL]
    uint64_t *full = (uint64_t *)(smraw + (size_t)NS * SL * 8);
    const long long base = ((long long)blockIdx.x * ts) & ~1LL;
    if (base >= n) return;
    const int size = (int)((n - base) < ts ? (n - base) : ts) & ~1;
    const int nsl = (size + SL - 1) / SL;
    if (threadIdx.x == 0) for (int s = 0; s < NS; ++s) mbar_init(full + s, 1);
    __syncthreads();
    auto issue = [&](int sl) {                            // all far streams of slice sl
        for (int k = 0; k < KFAR; ++k) {
            long long b = base + o.off[k]; if (b >= n) b -= n; if (b < 0) b += n; if (b + ts > n) b = 0;
            b &= ~1LL;
            const int cnt = (size - sl * SL) < SL ? (size - sl * SL) : SL;
            const int slot = (sl & 1) * KFAR + k;
            mbar_expect(full + slot, cnt * 8);
            bulk_g2s(ring + (size_t)slot * SL, in + b + (long long)sl * SL, cnt * 8, full + slot);
        }
    };
    if (threadIdx.x == 0) { issue(0); if (nsl > 1) issue(1); }
    constexpr int EPT = SL / NT;
    for (int sl = 0; sl < nsl; ++sl) {
        double g[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) g[e] = 0.0;
        const int l0 = sl * SL + threadIdx.x;
        for (int k = KFAR; k < o.k; ++k) {
            long long b = base + o.off[k]; if (b >= n) b -= n; if (b < 0) b += n; if (b + ts > n) b = 0;
            const double *q = in + b + l0;
            double t[EPT];
#pragma unroll
            for (int e = 0; e < EPT; ++e) t[e] = (l0 + e * NT < size) ? __ldg(q + e * NT) : 0.0;
#pragma unroll
            for (int e = 0; e < EPT; ++e) g[e] += 0.5 * t[e];
        }
        for (int k = 0; k < KFAR; ++k) {
            const int slot = (sl & 1) * KFAR + k;
            mbar_wait(full + slot, (sl >> 1) & 1);
#pragma unroll
            for (int e = 0; e < EPT; ++e) g[e] += 0.5 * ring[(size_t)slot * SL + threadIdx.x + e * NT];
        }
#pragma unroll
        for (int e = 0; e < EPT; ++e) if (l0 + e * NT < size) __stcs(out + base + l0 + e * NT, g[e]);
        __syncthreads();
        if (threadIdx.x == 0 && sl + 2 < nsl) issue(sl + 2);
    }
}

template <typename F>
static float timeit(F f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char **argv) {
    const long long n = 601080390LL;
    double *in, *out;
    CK(cudaMalloc(&in, (n + 8) * sizeof(double)));
    CK(cudaMalloc(&out, (n + 8) * sizeof(double)));
    CK(cudaMemset(in, 0, (n + 8) * sizeof(double)));
    // shifts of the L=32 prefix bonds (elements): bonds 0..15 at half filling
    const long long all[16] = {155117520LL, -77558760LL, 40116600LL, -20058300LL, 10400600LL, -5200300LL, 2704156LL, -1352078LL,
                               705432LL, -352716LL, 184756LL, -92378LL, 48620LL, -24310LL, 12870LL, -6435LL};
    int sm_count = 148;
    printf("variant,K,which,ms,read_GBps_streams\n");
    for (int cfg = 0; cfg < 2; ++cfg) {
        Offs o;
        const char *name;
        if (cfg == 0) { o.k = 10; for (int i = 0; i < 10; ++i) o.off[i] = all[i]; name = "bonds0-9(4far)"; }
        else if (cfg == 1) { o.k = 10; for (int i = 0; i < 10; ++i) o.off[i] = all[6 + i]; name = "near-only"; }
        else if (cfg == 2) { o.k = 4; for (int i = 0; i < 4; ++i) o.off[i] = all[i]; name = "far-only4"; }
        else if (cfg == 3) { o.k = 1; o.off[0] = 0; name = "copy"; }
        else { o.k = 10; for (int i = 0; i < 10; ++i) o.off[i] = all[2 * (i % 8)] + i; name = "mixed"; }
        const double gb = (double)n * 8 * o.k / 1e9;
        float ms;
        ms = timeit([&] { flat_kernel<4><<<sm_count * 8, 256>>>(in, out, n, o); }, 5);
        printf("flat4,%d,%s,%.3f,%.0f\n", o.k, name, ms, gb / ms);
        ms = timeit([&] { flat_kernel<8><<<sm_count * 8, 256>>>(in, out, n, o); }, 5);
        printf("flat8,%d,%s,%.3f,%.0f\n", o.k, name, ms, gb / ms);
        const int ts = 6435;
        const int ntile = (int)((n + ts - 1) / ts);
        CK(cudaFuncSetAttribute(tile_kernel<7, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
        CK(cudaFuncSetAttribute(tile_kernel<7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
        CK(cudaFuncSetAttribute(tile_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
        ms = timeit([&] { tile_kernel<7, 1><<<ntile, 512, 52 * 1024>>>(in, out, n, ts, o, 0); }, 5);
        printf("tile_u7_d1_4cta,%d,%s,%.3f,%.0f\n", o.k, name, ms, gb / ms);
        ms = timeit([&] { tile_kernel<7, 1><<<ntile, 512, 113 * 1024>>>(in, out, n, ts, o, 1); }, 5);
        printf("tile_u7_d1_2cta_smem,%d,%s,%.3f,%.0f\n", o.k, name, ms, gb / ms);
        ms = timeit([&] { tile_kernel<7, 2><<<ntile, 512, 113 * 1024>>>(in, out, n, ts, o, 1); }, 5);
        printf("tile_u7_d2_2cta_smem,%d,%s,%.3f,%.0f\n", o.k, name, ms, gb / ms);
        ms = timeit([&] { tile_kernel<4, 2><<<ntile, 512, 113 * 1024>>>(in, out, n, ts, o, 1); }, 5);
        printf("tile_u4_d2_2cta_smem,%d,%s,%.3f,%.0f\n", o.k, name, ms, gb / ms);
#define RUN(U_, D_, NT_, SMEMKB_, USESM_, LABEL_)                                                               \
    do {                                                                                                         \
        CK(cudaFuncSetAttribute(tile_kernel<U_, D_, NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
        ms = timeit([&] { tile_kernel<U_, D_, NT_><<<ntile, NT_, SMEMKB_ * 1024>>>(in, out, n, ts, o, USESM_); }, 5); \
        printf("%s,%d,%s,%.3f,%.0f\n", LABEL_, o.k, name, ms, gb / ms);                                          \
    } while (0)
#define RUNB(U_, NT_, SMEMKB_, USESM_, LABEL_)                                                                  \
    do {                                                                                                         \
        CK(cudaFuncSetAttribute(tileb_kernel<U_, NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
        ms = timeit([&] { tileb_kernel<U_, NT_><<<ntile, NT_, SMEMKB_ * 1024>>>(in, out, n, ts, o, USESM_); }, 5); \
        printf("%s,%d,%s,%.3f,%.0f\n", LABEL_, o.k, name, ms, gb / ms);                                          \
    } while (0)
        RUNB(7, 512, 113, 0, "B_u7_512thr_2cta_nosmem");
        RUNB(13, 512, 113, 0, "B_u13_512thr_2cta_nosmem");
        RUNB(13, 512, 113, 1, "B_u13_512thr_2cta_smem");
        RUNB(7, 512, 113, 1, "B_u7_512thr_2cta_smem");
        RUNB(7, 512, 52, 0, "B_u7_512thr_4cta_nosmem");
        RUNB(13, 512, 52, 0, "B_u13_512thr_4cta_nosmem");
        RUNB(4, 512, 52, 0, "B_u4_512thr_4cta_nosmem");
        RUNB(13, 256, 52, 1, "B_u13_256thr_4cta_smem");
        RUNB(26, 256, 52, 1, "B_u26_256thr_4cta_smem");
        RUNB(16, 1024, 113, 1, "B_u16_1024thr_2cta_smem");
        RUN(7, 1, 512, 113, 0, "X_512thr_2cta_nosmem");
        RUN(7, 1, 512, 52, 1, "X_512thr_4cta_smem");
        RUN(7, 1, 1024, 113, 1, "X_1024thr_2cta_smem");
        RUN(4, 1, 1024, 113, 1, "X_1024thr_u4_2cta_smem");
        RUN(7, 1, 256, 52, 1, "X_256thr_4cta_smem");
        RUN(7, 1, 256, 28, 0, "X_256thr_8cta_nosmem");
        RUN(7, 1, 128, 14, 0, "X_128thr_16cta_nosmem");
        RUN(4, 1, 512, 52, 0, "X_512thr_u4_4cta_nosmem");
        RUN(2, 1, 512, 52, 0, "X_512thr_u2_4cta_nosmem");
        RUN(2, 1, 1024, 113, 0, "X_1024thr_u2_2cta_nosmem");
        if (cfg == 0) {
#define RUNH(SL_, NT_, EXTRAKB_, LABEL_)                                                                       \
    do {                                                                                                         \
        const int smem = 8 * SL_ * 8 + 128 + EXTRAKB_ * 1024;                                                    \
        CK(cudaFuncSetAttribute(hybrid_kernel<SL_, 4, NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
        ms = timeit([&] { hybrid_kernel<SL_, 4, NT_><<<ntile, NT_, smem>>>(in, out, n, ts, o); }, 5);           \
        printf("%s_%dKB,%d,%s,%.3f,%.0f\n", LABEL_, smem / 1024, o.k, name, ms, gb / ms);                        \
    } while (0)
            RUNH(1024, 512, 0, "H_sl1024_512thr");
            RUNH(1024, 256, 0, "H_sl1024_256thr");
            RUNH(2048, 512, 0, "H_sl2048_512thr");
            RUNH(2048, 256, 0, "H_sl2048_256thr");
            RUNH(1024, 256, 50, "H_sl1024_256thr_2cta");
            RUNH(512, 256, 0, "H_sl512_256thr");
        }
        CK(cudaFuncSetAttribute(tile_kernel<13, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
        ms = timeit([&] { tile_kernel<13, 1><<<ntile, 512, 113 * 1024>>>(in, out, n, ts, o, 1); }, 5);
        printf("tile_u13_d1_2cta_smem,%d,%s,%.3f,%.0f\n", o.k, name, ms, gb / ms);
        ms = timeit([&] { tile_kernel<13, 1><<<ntile, 512, 52 * 1024>>>(in, out, n, ts, o, 0); }, 5);
        printf("tile_u13_d1_4cta,%d,%s,%.3f,%.0f\n", o.k, name, ms, gb / ms);
        ms = timeit([&] { tile_kernel<7, 2><<<ntile, 512, 52 * 1024>>>(in, out, n, ts, o, 0); }, 5);
        printf("tile_u7_d2_4cta,%d,%s,%.3f,%.0f\n", o.k, name, ms, gb / ms);
        {
            constexpr int SL = 1024, NS = 6;
            const int smem = NS * SL * 8 + 64;
            CK(cudaFuncSetAttribute(tma_kernel<SL, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
            ms = timeit([&] { tma_kernel<SL, NS><<<ntile, 512, smem>>>(in, out, n, ts, o); }, 5);
            printf("tma_sl1024_ns6_%dKB,%d,%s,%.3f,%.0f\n", smem / 1024, o.k, name, ms, gb / ms);
            ms = timeit([&] { tma_kernel<SL, NS><<<ntile, 512, 110 * 1024>>>(in, out, n, ts, o); }, 5);
            printf("tma_sl1024_ns6_2cta,%d,%s,%.3f,%.0f\n", o.k, name, ms, gb / ms);
        }
        {
            constexpr int SL = 2048, NS = 6;
            const int smem = NS * SL * 8 + 64;
            CK(cudaFuncSetAttribute(tma_kernel<SL, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
            ms = timeit([&] { tma_kernel<SL, NS><<<ntile, 512, smem>>>(in, out, n, ts, o); }, 5);
            printf("tma_sl2048_ns6_%dKB,%d,%s,%.3f,%.0f\n", smem / 1024, o.k, name, ms, gb / ms);
        }
        CK(cudaGetLastError());
    }
    return 0;
}
