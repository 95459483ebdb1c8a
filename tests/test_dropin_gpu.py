"""Literal drop-in: the REFERENCE'S OWN host code -- Mapping.cpp (MapperBaseGPU) and the gpuMapper
class of GPU_Advection.h, compiled unmodified -- linked against libbimocq_b200.so instead of the
reference's GPU_kernel.cu, must produce the same fields.  oracle/Makefile builds the same driver
(oracle/dropin_driver.cpp) twice; this test runs both binaries and compares the dumps."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFBIN = os.path.join(ROOT, "oracle", "_ref", "dropin_ref")
OURBIN = os.path.join(ROOT, "oracle", "_ref", "dropin_ours")


@pytest.mark.parametrize("ni,nj,nk,L", [(40, 36, 44, 1.25), (37, 41, 35, 0.2)])
def test_reference_host_code_runs_unchanged_on_our_library(cuda, tmp_path, ni, nj, nk, L):
    if not (os.path.exists(REFBIN) and os.path.exists(OURBIN)):
        pytest.skip("oracle/_ref/dropin_* not built (needs /root/reference at build time)")
    outs = []
    for exe in (REFBIN, OURBIN):
        out = tmp_path / (os.path.basename(exe) + ".bin")
        r = subprocess.run([exe, str(ni), str(nj), str(nk), str(L), "7", str(out)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(np.fromfile(out, dtype=np.float32))
    ref, ours = outs
    assert ref.shape == ours.shape and np.isfinite(ref).all()
    assert np.abs(ref).max() > 0.1
    err = float(np.abs(ours - ref).max() / np.abs(ref).max())
    print(f"drop-in ({ni}x{nj}x{nk}, L={L}): max rel difference over 7 frames of all dumped fields = {err:.2e}, "
          f"bitwise equal = {bool(np.array_equal(ref, ours))}")
    assert err <= 1e-5
