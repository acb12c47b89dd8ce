"""GPU parity of the device-resident observables (sd_vec_observables; Observables.jl:14-109, PublicAPI.jl:94-106)
against the oracle's restatement: every kernel path (block-layout vectors go through the layout conversion), f64 and
c128, sector and full basis, host and device inputs.  Tolerance 1e-12 absolute (sums of |psi|^2-weighted +-1/4)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle.oracle as orc  # noqa: E402
from conftest import sd  # noqa: E402


@pytest.mark.parametrize("L,nup", [(8, 4), (12, 5), (16, 8), (18, 9), (10, None), (33, 1)])
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_observables_match_oracle(L, nup, dtype):
    m, om = sd.XXZChain(L, nup=nup), orc.XXZChain(L, nup=nup)
    rng = np.random.default_rng(L)
    psi = rng.standard_normal(m.dim).astype(dtype)
    if dtype == np.complex128:
        psi = psi + 1j * rng.standard_normal(m.dim)
    psi /= np.linalg.norm(psi)
    mags = sd.magnetization_per_site(psi, m)
    assert np.allclose(mags, orc.magnetization_per_site(psi, om), atol=1e-12)
    C = sd.connected_correlations(m.to_device(psi), m)                      # device-resident input
    assert np.allclose(C, orc.connected_correlations(psi, om), atol=1e-12)
    Sq, ref = sd.structure_factor(m, psi), orc.structure_factor_Sq(psi, om)
    assert set(Sq) == set(ref) and all(abs(Sq[q] - ref[q]) < 1e-12 for q in ref)
    again = sd.magnetization_per_site(psi, m)
    assert np.array_equal(mags, again)                                      # deterministic reduction order


def test_structure_factor_public_api_L4():
    """test_PublicAPI.jl:135-150."""
    m = sd.XXZChain(4, Jxy=1.0, Jz=1.0, nup=2)
    _, psi0 = sd.groundstate(m, lanc_m=6, rng=np.random.default_rng(2))
    a, b = sd.structure_factor(m, psi0), sd.structure_factor_Sq(psi0, m)
    assert set(a) == set(b) and all(abs(a[q] - b[q]) < 1e-12 for q in a)
    assert len(a) == m.L and all(np.isfinite(v) for v in a.values())
    assert abs(a[0.0]) < 1e-12                                              # total Sz is sharp in a sector


def test_time_evolved_neel_state_observables_stay_on_device():
    """What the row is for: psi(t) from the Chebyshev stepper is consumed without a download."""
    L = 20
    m, om = sd.XXZChain(L, nup=L // 2), orc.XXZChain(L, nup=L // 2)
    psi0 = sd.neel_state(m, device=True).astype(np.complex128)
    pt = sd.chebyshev_time_evolve(psi0, 0.4, sd.apply_H_, m, cheb_n=40, Ebounds=(-9.2, 4.8), device=True)
    mags = sd.magnetization_per_site(pt, m)
    ref = orc.magnetization_per_site(pt.to_host(), om)
    assert np.allclose(mags, ref, atol=1e-12) and abs(mags.sum()) < 1e-12   # Sz_tot = 0 is conserved
    assert np.all(np.abs(mags) < 0.5) and abs(mags[0] + mags[-1]) < 1e-10   # reflection maps the Neel state to its flip
