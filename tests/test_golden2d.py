"""2D golden vectors (tests/golden/ref2d_steps.npz, produced by the reference's own 2D code via
tests/golden/make_golden_2d.py):
* CPU: the committed file is reproduced bit for bit by re-running the reference here (when
  oracle/_ref/libref2d.so is built) -- pins the fixture to the reference;
* GPU: bmq2d_* reproduces every recorded step from the recorded pre-state within 1e-5."""
import importlib.util
import os

import numpy as np
import pytest

import ref2d
from helpers import rel_linf

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "ref2d_steps.npz")


def _gen():
    spec = importlib.util.spec_from_file_location("make_golden_2d", os.path.join(HERE, "golden", "make_golden_2d.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_golden_file_is_what_the_reference_computes():
    if not ref2d.available():
        pytest.skip("oracle/_ref/libref2d.so not built")
    m = _gen()
    gold = np.load(GOLD)
    ref = ref2d.Ref2D(m.NI, m.NJ, m.L, m.BLEND)
    u, v, rho, T = m.initial(m.NI, m.NJ, m.L)
    for mem, a in (("u", u), ("v", v), ("u_init", u), ("v_init", v), ("rho", rho), ("rho_init", rho), ("temperature", T), ("T_init", T)):
        ref.field(mem)[...] = a
    remaps = 0
    for frame in range(4):
        ref.phase_a(m.DT, frame)
        for mem in m.AFTER_A:
            assert np.array_equal(ref.field(mem), gold[f"f{frame}:a:{mem}"]), (frame, mem)
        adv = [ref.field(x).copy() for x in ("u", "v", "rho", "temperature")]
        ref.phase_b(m.DT, frame, *m.forcing(adv, m.DT))
        for mem in m.AFTER_B:
            assert np.array_equal(ref.field(mem), gold[f"f{frame}:b:{mem}"]), (frame, mem)
        remaps += int(gold[f"f{frame}:b:flags"].sum())
    assert sum(int(gold[f"f{f}:b:flags"].sum()) for f in range(m.FRAMES)) >= 1
    ref.close()


@pytest.mark.gpu
def test_gpu_reproduces_golden_steps(cuda):
    from gpufluidsimulation_b200.solver2d import BimocqAdvection2D
    m = _gen()
    gold = np.load(GOLD)
    h = float(np.float32(m.L) / np.float32(m.NI))
    gpu = BimocqAdvection2D(m.NI, m.NJ, h, m.BLEND)
    worst = 0.0
    for frame in range(m.FRAMES):
        for mem, name in ref2d.MEMBERS.items():
            gpu.upload(name, gold[f"f{frame}:pre:{mem}"])
        c = gold[f"f{frame}:pre:counters"]
        gpu.set_counters(int(c[0]), int(c[1]))
        gpu.advect(frame, m.DT)
        for mem in m.AFTER_A:
            got, want = gpu.download(ref2d.MEMBERS[mem]), gold[f"f{frame}:a:{mem}"]
            e = rel_linf(got[1:-1], want[1:-1])
            worst = max(worst, e)
            assert e <= 1e-5, (frame, "A", mem, e)
        adv = [gold[f"f{frame}:a:{x}"] for x in ("u", "v", "rho", "temperature")]
        for nme, a in zip(("U", "V", "RHO", "T", "U_SAVE", "V_SAVE", "RHO_SAVE", "T_SAVE"), adv + adv):
            gpu.upload(nme, a)
        gpu.accumulate_host(frame, m.DT, *m.forcing([a.copy() for a in adv], m.DT))
        st = gpu.stats()
        assert [st["vel_remap"], st["scalar_remap"]] == list(gold[f"f{frame}:b:flags"]), frame
        for mem in m.AFTER_B:
            e = rel_linf(gpu.download(ref2d.MEMBERS[mem]), gold[f"f{frame}:b:{mem}"])
            worst = max(worst, e)
            assert e <= 1e-5, (frame, "B", mem, e)
    print(f"2D golden steps: worst rel Linf {worst:.2e}")
    gpu.close()
