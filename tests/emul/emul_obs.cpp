// emul_obs.cpp -- CPU run of the observables kernel's per-state arithmetic (sd_obs.h sd_obs_accum: the
// rotate / xor / popcount form of sum_i s_i s_{i+r}) over an explicit list of states.  TEST INFRASTRUCTURE ONLY.
#include <cstdint>
#include "../../spindynamics.jl_b200/csrc/sd_obs.h"
extern "C" int emul_obs(int L, const uint64_t *states, uint64_t n, const double *psi, int nc, double *mags, double *zz) {
    if (L < 1 || L > 63) return -1;
    for (int r = 0; r < L; ++r) { mags[r] = 0.0; zz[r] = 0.0; }
    for (uint64_t i = 0; i < n; ++i) {
        const double w = nc == 2 ? psi[2 * i] * psi[2 * i] + psi[2 * i + 1] * psi[2 * i + 1] : psi[i] * psi[i];
        if (w == 0.0) continue;
        for (int r = 0; r < L; ++r) sd_obs_accum(L, r, states[i], w, mags[r], zz[r]);
    }
    return 0;
}
