"""The CUDA library against the committed golden outputs of the reference's own kernels
(tests/golden/ref3d_kernels.npz): works on a GPU box without /root/reference and without oracle/_ref."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import TOL_STEP, Case3D, rel_linf, run_gpu_symbol
from test_golden_cpu import GOLD, _gen

pytestmark = pytest.mark.gpu


def test_library_matches_reference_kernel_golden_vectors(cuda, oracle):
    if not os.path.exists(GOLD):
        pytest.skip("golden file not generated yet")
    from gpufluidsimulation_b200 import load_library
    lib = C.CDLL(load_library()._name)
    m = _gen()
    c = Case3D(m.NI, m.NJ, m.NK, m.H, seed=11)
    gold = np.load(GOLD)
    worst = 0.0
    for name, (args, outs) in m.calls(c).items():
        got = run_gpu_symbol(lib, name, args)
        for q in outs:
            err = rel_linf(got[q], gold[f"{name}:out{q}"])
            worst = max(worst, err)
            assert err <= TOL_STEP, (name, q, err)
    print(f"library vs golden reference-kernel outputs: worst rel Linf {worst:.2e}")
