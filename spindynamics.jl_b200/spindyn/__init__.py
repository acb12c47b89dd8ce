"""spindyn -- Python host mirror of SpinDynamics.jl's H.psi hot path on top of
libspindyn_cuda (B200, sm_100a).  Re-exports the reference's export list for
this path (SpinDynamics.jl:1-75) plus the site-resolved KPM drivers of TimeEvolution/KPM.jl;
the general create_spin_operator stays in the Julia package (only the single-site S^z it is used with is here)."""
from ._lib import LIB_PATH, SIGNATURES, OTHER_SYMBOLS, SpinDynError, ZeroNormError, lib
from .core import (Context, DeviceVector, Model, PinnedBuffer, VecSet, default_context, device_count,
                   set_default_context)
from .api import (XXZChain, Sz_q_vector, apply_H_, apply_H_neg_, apply_rescaled_H_, bit_at, build_full_basis,
                  build_model, build_sector_basis, chebyshev_coefficients, chebyshev_time_evolve,
                  compute_chebyshev_moments, domain_wall_state, dynamical_structure_factor,
                  estimate_energy_bounds, flip_bits, get_kernel, get_rescaling_params, groundstate,
                  kpm_sqw, kpm_sw, krylov_time_evolve, krylov_time_evolve_, KrylovWorkspace, ChebyshevWorkspace,
                  lanczos_extremal, lanczos_groundstate, lanczos_groundstate_lean, lanczos_sqw, magnetization_per_site,
                  connected_correlations, structure_factor_Sq, structure_factor,
                  lanczos_tridiag, long_range_hopping, momenta, neel_state, nn_hopping, polarized_state,
                  polarized_state_with_flips, randn_complex, spectral_from_tridiagonal, sz_value,
                  time_evolve, _rescaling_from_bounds,
                  site_sz_operator, kpm_get_rescaling_params, get_jackson_kernel, evaluate_chebyshev_series,
                  compute_cross_chebyshev_moments, kpm_dynamical_correlation, kpm_correlation_matrix, Sqw)
