"""Independent dense construction of spin-1/2 Hamiltonians from Kronecker
products of the spin matrices.  Shares no code with oracle/ or the CUDA
library: it is the physics-textbook definition the oracle is pinned against.

Convention (reference Hamiltonian.jl:19-29, Basis.jl:41-46): site i (1-based)
is bit i-1 of the basis integer, bit 1 = up = +1/2.  In a Kronecker product the
LAST factor is the least-significant bit, so site 1 is the last factor.
"""
import itertools

import numpy as np

SZ = np.array([[-0.5, 0.0], [0.0, 0.5]])        # index 0 = bit 0 = down
SP = np.array([[0.0, 0.0], [1.0, 0.0]])         # S+ |down> = |up>: row 1, col 0
SM = SP.T


def _site_op(op, site, L):
    mats = [np.eye(2)] * L
    mats = list(mats)
    mats[L - site] = op                          # site 1 -> last factor
    out = mats[0]
    for m in mats[1:]:
        out = np.kron(out, m)
    return out


def dense_H_full(L, hopping, zz, field):
    """H = sum_(i,j,J) J (S+_i S-_j + S-_i S+_j) + sum Jz Sz_i Sz_j + sum h_i Sz_i."""
    D = 1 << L
    H = np.zeros((D, D))
    for (i, j, J) in hopping:
        H += J * (_site_op(SP, i, L) @ _site_op(SM, j, L) + _site_op(SM, i, L) @ _site_op(SP, j, L))
    for (i, j, Jz) in zz:
        H += Jz * (_site_op(SZ, i, L) @ _site_op(SZ, j, L))
    for i in range(1, L + 1):
        H += field[i - 1] * _site_op(SZ, i, L)
    return H


def sector_states(L, nup):
    """Lexicographic combinations order (Combinatorics.jl / itertools)."""
    return [sum(1 << (i - 1) for i in comb) for comb in itertools.combinations(range(1, L + 1), nup)]


def dense_H(L, nup, hopping, zz, field):
    H = dense_H_full(L, hopping, zz, field)
    if nup is None:
        return H
    st = sector_states(L, nup)
    return H[np.ix_(st, st)]


def xxz_lists(L, Jxy=1.0, Jz=1.0, hz=0.0, periodic=False):
    hop = [(i, i + 1, Jxy / 2) for i in range(1, L)]
    zz = [(i, i + 1, Jz) for i in range(1, L)]
    if periodic and L > 2:
        hop.append((L, 1, Jxy / 2))
        zz.append((L, 1, Jz))
    return hop, zz, [hz] * L
