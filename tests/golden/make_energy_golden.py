"""Writes tests/golden/energy_golden.json: ground-state energies of the open Heisenberg chain
(Jxy = Jz = 1, Sz = 0) at sizes beyond dense diagonalisation, from ARPACK (scipy eigsh, tol 0 =
machine precision) driving the CPU oracle's apply_H! restatement as a LinearOperator.  Independent of
the Lanczos code under test (different eigensolver, CPU matvec).  Also the energy of the Neel state
and sum rules used by the full-size property tests.  Run from the repo root (about 2 minutes):
    python tests/golden/make_energy_golden.py
"""
import json
import os
import sys

import numpy as np
from scipy.sparse.linalg import LinearOperator, eigsh

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

out = {"model": "XXZChain(L, Jxy=1, Jz=1, hz=0, nup=L/2, boundary=open)", "E0": {}, "residual": {}}
for L in (18, 20, 22, 24):
    m = orc.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=L // 2)
    N = len(m)

    def mv(x, m=m):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        y = np.empty_like(x)
        orc.apply_H_(y, x, m)
        return y

    op = LinearOperator((N, N), matvec=mv, dtype=np.float64)
    v0 = orc.fill_seeded(N, 7)
    w, v = eigsh(op, k=1, which="SA", tol=0, v0=v0, ncv=40, maxiter=5000)
    r = np.linalg.norm(mv(v[:, 0]) - w[0] * v[:, 0])
    out["E0"][str(L)] = float(w[0])
    out["residual"][str(L)] = float(r)
    print(L, N, repr(float(w[0])), r, flush=True)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "energy_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
