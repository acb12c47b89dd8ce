"""Writes tests/golden/apply_golden.npz from the CPU oracle.

The reference is Julia and cannot run in this image, so the golden vectors are
outputs of the oracle (itself pinned against the reference tests' known answers
and an independent Kronecker construction, tests/test_oracle.py).  Inputs are
the counter-based seeded psi, so only outputs are stored.  Run from the repo
root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

SEED, JXY, JZ, HZ = 20261018, 1.0, 0.75, 0.1
out = {"seed": SEED, "Jxy": JXY, "Jz": JZ, "hz": HZ}
for L, nup in [(10, 5), (12, 6), (13, 4), (14, 7), (16, 8), (17, 7)]:      # L >= 16: the block-layout kernel's sizes
    m = orc.XXZChain(L, Jxy=JXY, Jz=JZ, hz=HZ, nup=nup)
    for kind, cplx in (("f64", False), ("c128", True)):
        psi = orc.fill_seeded(len(m), SEED, cplx=cplx)
        ref = np.empty_like(psi)
        orc.apply_H_(ref, psi, m)
        out[f"out_{L}_{nup}_{kind}"] = ref
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "apply_golden.npz"), **out)
print("wrote", len(out) - 4, "vectors")
