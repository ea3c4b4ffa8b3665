"""z-slab engines with halo exchange through peer memory, against the single-domain oracle.
Runs with several slabs on ONE GPU (always) and across two GPUs when the box has them."""
import numpy as np
import pytest

from oracle.pyoracle import BC_PEC, BC_PMC, BC_MUR, BC_PML
from tests import cases
from tests.gpu_util import operator_from_oracle
from openems_b200.slabs import slab_range, held_range, link_engines_in_process

pytestmark = pytest.mark.gpu


def run_slabs(s, bounds, devices, steps=(1, 2, 40), fused=None):
    op = operator_from_oracle(s)
    nz = s.N[2]
    engines = [op.CreateEngine(device=devices[r], slab=(bounds[r], bounds[r + 1])) for r in range(len(bounds) - 1)]
    if fused is not None:
        for e in engines:
            e.SetOption("fused", fused)
    link_engines_in_process(engines)
    if fused:
        assert all("fused_EH" in [n for n, _ in e.TimeSchedule(0)] for e in engines)
    total = 0
    for n in steps:
        s.iterate(n)
        for _ in range(n):
            for e in engines:
                e.IterateTS(1)
        total += n
        for e in engines:
            e.Synchronize()
        for w, ref in ((0, s.volt), (1, s.curr)):
            got = np.zeros_like(ref)
            for r, e in enumerate(engines):
                zb, ze = bounds[r], bounds[r + 1]
                h0, h1 = held_range(nz, zb, ze)
                f = e.GetFields(w)
                got[..., zb:ze] = f[..., zb - h0: ze - h0]
            bad = int((got.view(np.uint32) != ref.view(np.uint32)).sum())
            assert bad == 0, "%d values differ after %d steps (field %d)" % (bad, total, w)
    assert np.abs(s.volt).max() > 0
    return engines


def test_two_slabs_one_gpu_pml():
    s = cases.uniform_box(n=(24, 20, 36), bc=(BC_PML,) * 6, pml=5)
    run_slabs(s, [0, 18, 36], [0, 0])


def test_three_uneven_slabs_one_gpu_mixed_bc():
    s = cases.engine_cavity()
    run_slabs(s, [0, 9, 20, 33], [0, 0, 0], steps=(1, 3, 60))


@pytest.mark.parametrize("fused", [0, 1])
def test_slabs_one_pass_and_two_pass(fused):
    """both timestep schedules on z-slabs (the one-pass schedule does the slab's top plane after
    the neighbour's E plane has arrived), UPML + Mur + PMC mix and a Mur-only mesh"""
    s = cases.engine_cavity()
    run_slabs(s, [0, 12, 21, 33], [0, 0, 0], steps=(1, 2, 45), fused=fused)
    s = cases.uniform_box(n=(20, 22, 30), bc=(BC_MUR,) * 6)
    run_slabs(s, [0, 15, 30], [0, 0], steps=(1, 3, 50), fused=fused)


@pytest.mark.parametrize("fused", [0, 1])
def test_two_gpus(fused):
    """two engines on two real GPUs in one process (peer access, no IPC): both schedules, the one-pass one with
    its ping-pong ghost-plane targets and the per-device shared-memory opt-in of the TMA kernel"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = cases.uniform_box(n=(40, 36, 44), bc=(BC_PML,) * 6, pml=8)
    engines = run_slabs(s, [0, 22, 44], [0, 1], steps=(1, 5, 80), fused=fused)
    if fused:
        assert all(e.GetOption("tma") == 1 for e in engines)


def test_dispersive_block_across_slabs_one_pass():
    """C4 in small on 3 slabs, one-pass schedule: the Drude block crosses both interfaces (the ADE of a slab's top
    plane is applied by the list kernel after update_H_top, everything else inside the one-pass kernel)"""
    from tests import configs
    s = configs.c4_drude_block(block=(12, 35))
    engines = run_slabs(s, [0, 17, 30, 48], [0, 0, 0], steps=(1, 2, 60), fused=1)
    for e in engines:
        names = [n for n, _ in e.TimeSchedule(0)]
        assert e.GetOption("fused") == 1 and "lorentz_pre_V" in names
    assert "lorentz_apply_I_top" in [n for n, _ in engines[0].TimeSchedule(0)]
