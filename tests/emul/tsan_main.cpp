// tsan_main.cpp -- ThreadSanitizer driver for the threaded host run of the ring kernel's protocol (emul_blk.cpp,
// run_tiles_ring_threaded): builds a few sector models, runs the threaded variant (1 producer + 15 consumer threads on a
// reused 4-slot ring) and compares it bit for bit with the sequential emulation.  TEST INFRASTRUCTURE ONLY.
//   make -C tests/emul tsan && tests/emul/tsan_ring
#include <cstdio>
#include <random>
#include "emul_blk.cpp"

int main() {
    int bad = 0;
    const int cases[][3] = {{16, 8, 1}, {17, 8, 2}, {18, 9, 1}, {19, 7, 3}, {20, 10, 1}};
    for (const auto &cs : cases) {
        const int L = cs[0], k = cs[1], world = cs[2];
        std::vector<uint64_t> binom(SD_BINOM_DIM * SD_BINOM_DIM);
        sd_fill_binom(binom.data());
        const uint64_t N = binom[(size_t)L * SD_BINOM_DIM + k];
        std::vector<uint64_t> states(N);
        for (uint64_t r = 0; r < N; ++r) states[r] = sd_unrank_state(r, L, k, binom.data(), SD_BINOM_DIM);
        std::mt19937_64 rng(L * 100 + k);
        std::normal_distribution<double> nd;
        std::vector<double> J(L), Jz(L), h(L), psi(N), vprev(N), phi(N);
        for (int i = 0; i < L; ++i) { J[i] = 0.3 + 0.1 * (i % 5); Jz[i] = 0.5 - 0.07 * (i % 7); h[i] = 0.01 * i; }
        for (uint64_t i = 0; i < N; ++i) { psi[i] = nd(rng); vprev[i] = nd(rng); phi[i] = nd(rng); }
        for (int ndirect : {0, 3}) {
            for (int mode : {0, 2}) {
                std::vector<double> a(N, 0.0), b(N, 0.0);
                double ra[4] = {0}, rb[4] = {0};
                for (int rank = 0; rank < world; ++rank) {
                    double r1[4] = {0}, r2[4] = {0};
                    const int v_seq = 2 + 4096 * ndirect, v_thr = v_seq + 1024;
                    const int rc1 = emul_blk_apply(L, k, J.data(), Jz.data(), h.data(), 1, states.data(), N, psi.data(), a.data(), world, rank,
                                                   mode, mode ? 7 : 0, 1.0, 2.5, 0.3, mode ? vprev.data() : nullptr, mode ? phi.data() : nullptr,
                                                   nullptr, 0.0, 0.0, r1, nullptr, 1 << 20, nullptr, v_seq);
                    const int rc2 = emul_blk_apply(L, k, J.data(), Jz.data(), h.data(), 1, states.data(), N, psi.data(), b.data(), world, rank,
                                                   mode, mode ? 7 : 0, 1.0, 2.5, 0.3, mode ? vprev.data() : nullptr, mode ? phi.data() : nullptr,
                                                   nullptr, 0.0, 0.0, r2, nullptr, 1 << 20, nullptr, v_thr);
                    if (rc1 || rc2) { printf("L=%d k=%d world=%d rank=%d: rc %d %d\n", L, k, world, rank, rc1, rc2); ++bad; }
                    for (int s = 0; s < 4; ++s) { ra[s] += r1[s]; rb[s] += r2[s]; }
                }
                bool same = true;
                for (uint64_t i = 0; i < N; ++i) if (a[i] != b[i]) { same = false; break; }
                for (int s = 0; s < 4; ++s) if (std::fabs(ra[s] - rb[s]) > 1e-9 * (1.0 + std::fabs(ra[s]))) same = false;
                printf("L=%d k=%d world=%d ndirect=%d mode=%d: %s\n", L, k, world, ndirect, mode, same ? "threaded == sequential" : "MISMATCH");
                if (!same) ++bad;
            }
        }
    }
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad ? 1 : 0;
}
