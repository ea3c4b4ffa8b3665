"""Host-side multi-GPU logic on CPU: slab partitioning, and the halo-exchange PROTOCOL the CUDA
engines use (which plane, which components, which direction, at which point of the timestep),
emulated with numpy on world_size-2 gloo ranks and compared bit for bit with the
single-domain oracle.  (The device kernels themselves are covered by tests/test_gpu_multi.py.)"""
import os

import numpy as np
import pytest

from openems_b200.slabs import slab_range, held_range


def test_slab_range_covers_domain():
    for nz, world in ((1024, 8), (33, 3), (44, 2), (100, 7)):
        prev = 0
        for r in range(world):
            zb, ze = slab_range(nz, world, r)
            assert zb == prev and ze - zb >= 2
            prev = ze
        assert prev == nz
    assert held_range(100, 0, 50) == (0, 51)
    assert held_range(100, 50, 100) == (49, 100)
    assert held_range(100, 25, 50) == (24, 51)


def test_slab_range_pml_weighting_shrinks_end_slabs():
    plain = [slab_range(1024, 8, r) for r in range(8)]
    w = [slab_range(1024, 8, r, pml_lo=8, pml_hi=8, pml_weight=1.0) for r in range(8)]
    assert w[0][1] - w[0][0] < plain[0][1] - plain[0][0]
    assert w[7][1] - w[7][0] < plain[7][1] - plain[7][0]
    assert w[-1][1] == 1024 and w[0][0] == 0
    with pytest.raises(ValueError):
        slab_range(5, 4, 0)


def _np_update_E(V, I, vv, vi, k0, k1, gz0):
    """Engine::UpdateVoltages (engine.cpp:110-168) on local planes [k0,k1); arrays [3][x][y][zl];
    gz0 = global z of local plane 0 (the k-1 clamp applies at GLOBAL z = 0 only)"""
    f = np.float32

    def sh(a, axis):  # a[pos - (pos>0)] along axis
        b = np.roll(a, 1, axis)
        idx = [slice(None)] * 3
        idx[axis] = 0
        b[tuple(idx)] = a[tuple(idx)]
        return b
    I0, I1, I2 = I
    zs = slice(k0, k1)

    def zm(a):  # a[k-1] with the clamp only at global z=0
        b = np.empty_like(a[..., zs])
        for q, k in enumerate(range(k0, k1)):
            km = k - 1 if (k + gz0) > 0 else k
            b[..., q] = a[..., km]
        return b
    c0 = ((I2 - sh(I2, 1))[..., zs] - I1[..., zs]) + zm(I1)
    c1 = ((I0[..., zs] - zm(I0)) - I2[..., zs]) + sh(I2, 0)[..., zs]
    c2 = ((I1 - sh(I1, 0))[..., zs] - I0[..., zs]) + sh(I0, 1)[..., zs]
    for n, c in enumerate((c0, c1, c2)):
        V[n][..., zs] = (V[n][..., zs] * vv[n][..., zs]).astype(f) + (vi[n][..., zs] * c.astype(f)).astype(f)


def _np_update_H(V, I, ii, iv, k0, k1):
    """Engine::UpdateCurrents (engine.cpp:170-222) on local planes [k0,k1), i<Nx-1, j<Ny-1"""
    V0, V1, V2 = V
    X, Y = slice(0, -1), slice(0, -1)
    Xp, Yp = slice(1, None), slice(1, None)
    zs, zp = slice(k0, k1), slice(k0 + 1, k1 + 1)
    c0 = ((V2[X, Y, zs] - V2[X, Yp, zs]) - V1[X, Y, zs]) + V1[X, Y, zp]
    c1 = ((V0[X, Y, zs] - V0[X, Y, zp]) - V2[X, Y, zs]) + V2[Xp, Y, zs]
    c2 = ((V1[X, Y, zs] - V1[Xp, Y, zs]) - V0[X, Y, zs]) + V0[X, Yp, zs]
    for n, c in enumerate((c0, c1, c2)):
        I[n][X, Y, zs] = (I[n][X, Y, zs] * ii[n][X, Y, zs]) + (iv[n][X, Y, zs] * c)


def _rank_main(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import BC_PEC
    from tests import cases
    s = cases.uniform_box(n=(14, 12, 20), bc=(BC_PEC,) * 6)
    nz = s.N[2]
    zb, ze = slab_range(nz, world, rank)
    h0, h1 = held_range(nz, zb, ze)
    co = {w: s.coeff(w)[..., h0:h1].copy() for w in ("vv", "vi", "ii", "iv")}
    V = np.zeros((3,) + co["vv"].shape[1:], np.float32)
    I = np.zeros_like(V)
    idx, d, amp, delay = s.excitation(0)
    sig, _, _ = s.signal()
    steps = 45
    for ts in range(steps):
        # E half-step on the owned planes, then the excitation (Apply2Voltages)
        _np_update_E(V, I, co["vv"], co["vi"], zb - h0, ze - h0, h0)
        for e in range(len(d)):
            z = int(idx[2][e])
            if zb <= z < ze:
                pos = ts - int(delay[e])
                pos = pos if 0 < pos < len(sig) else 0
                V[d[e]][idx[0][e], idx[1][e], z - h0] = np.float32(V[d[e]][idx[0][e], idx[1][e], z - h0] + np.float32(amp[e] * sig[pos]))
        # halo: tangential E (x, y) of my lowest owned plane -> lower rank's ghost-E plane
        reqs = []
        if rank > 0:
            t = torch.from_numpy(np.ascontiguousarray(V[:2, :, :, zb - h0]))
            reqs.append(dist.isend(t, rank - 1))
        if rank < world - 1:
            r = torch.empty((2,) + V.shape[1:3], dtype=torch.float32)
            dist.recv(r, rank + 1)
            V[:2, :, :, ze - h0] = r.numpy()
        for q_ in reqs:
            q_.wait()
        # H half-step on the owned planes below the global top
        _np_update_H(V, I, co["ii"], co["iv"], zb - h0, min(ze, nz - 1) - h0)
        # halo: tangential H of my highest owned plane -> upper rank's ghost-H plane
        reqs = []
        if rank < world - 1:
            t = torch.from_numpy(np.ascontiguousarray(I[:2, :, :, ze - 1 - h0]))
            reqs.append(dist.isend(t, rank + 1))
        if rank > 0:
            r = torch.empty((2,) + I.shape[1:3], dtype=torch.float32)
            dist.recv(r, rank - 1)
            I[:2, :, :, zb - 1 - h0] = r.numpy()
        for q_ in reqs:
            q_.wait()
    s.iterate(steps)
    okV = np.array_equal(V[..., zb - h0: ze - h0].view(np.uint32), s.volt[..., zb:ze].view(np.uint32))
    okI = np.array_equal(I[..., zb - h0: ze - h0].view(np.uint32), s.curr[..., zb:ze].view(np.uint32))
    q.put((rank, bool(okV), bool(okI), float(np.abs(s.volt).max())))
    dist.barrier()
    dist.destroy_process_group()


def test_halo_protocol_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, okV, okI, vmax in res:
        assert okV and okI, "rank %d differs from the single-domain oracle" % rank
        assert vmax > 0


# ---------------------------------------------------------------------------------------------
# read-out on slabs: the host-side merge logic of openems_b200/slabs.py on two gloo ranks.  Every rank
# stands in for a slab engine with a stub that evaluates the ORACLE's dump / mode-matching / steady-state
# terms on the planes the rank owns (what the device kernels are tested to do in
# tests/test_gpu_slab_readout.py); the merged results must equal the single-domain oracle.
# ---------------------------------------------------------------------------------------------
class _StubSlabEngine:
    def __init__(self, s, zb, ze, box, mode):
        self.s, self.zb, self.ze, self.box, self.mode = s, zb, ze, box, mode

    def ReadDump(self, dump_id):
        start, stop = self.box
        full = self.s.dump_field(0, 1, start, stop)               # {3, nz, ny, nx}
        zs = [z for z in range(start[2], stop[2] + 1) if self.zb <= z < self.ze]
        return full[:, [z - start[2] for z in zs]].copy()

    def ReadModeMatchRaw(self, mode_id):
        is_H, ny, start, stop, d0, d1, area = self.mode
        nP, nPP = (ny + 1) % 3, (ny + 2) % 3
        value = purity = 0.0
        for a in range(stop[nP] - start[nP] + 1):
            for b in range(stop[nPP] - start[nPP] + 1):
                pos = [0, 0, 0]
                pos[ny] = start[ny]; pos[nP] = start[nP] + a; pos[nPP] = start[nPP] + b
                if not (self.zb <= pos[2] < self.ze):
                    continue
                f = self.s.dump_field(is_H, 1, tuple(pos), tuple(pos))[:, 0, 0, 0].astype(np.float64)
                value += f[nP] * d0[a, b] * area[a, b] + f[nPP] * d1[a, b] * area[a, b]
                purity += f[nP] * f[nP] * area[a, b] + f[nPP] * f[nPP] * area[a, b]
        return value, (value * value / purity if purity else 0.0), purity


def _readout_rank(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import BC_PEC, BC_MUR
    from tests import cases
    from openems_b200 import slabs
    s = cases.uniform_box(n=(14, 12, 20), bc=(BC_MUR, BC_PEC, BC_PEC, BC_MUR, BC_PEC, BC_MUR))
    s.iterate(40)
    nz = s.N[2]
    zb, ze = slab_range(nz, world, rank)
    box = ((2, 1, 3), (11, 9, 17))
    ny, start, stop = 0, (6, 1, 2), (6, 10, 18)
    rng = np.random.default_rng(3)
    d0, d1, area = rng.random((10, 17)), rng.random((10, 17)), 1e-6 * (1 + rng.random((10, 17)))
    eng = _StubSlabEngine(s, zb, ze, box, (0, ny, start, stop, d0, d1, area))
    got = slabs.read_dump_distributed(eng, 0, dist, world)
    ok_dump = np.array_equal(got.view(np.uint32), s.dump_field(0, 1, *box).view(np.uint32))
    mm = slabs.mode_match_distributed(eng, 0, dist, world)
    ref = _StubSlabEngine(s, 0, nz, box, (0, ny, start, stop, d0, d1, area)).ReadModeMatchRaw(0)
    ok_mode = abs(mm[0] - ref[0]) <= 1e-12 * abs(ref[0]) and abs(mm[1] - ref[1]) <= 1e-12 * abs(ref[1]) and ref[0] != 0
    q.put((rank, bool(ok_dump), bool(ok_mode)))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_readout_merge_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_readout_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok_dump, ok_mode in res:
        assert ok_dump, "rank %d: concatenated slab dumps differ from the single-domain dump" % rank
        assert ok_mode, "rank %d: merged mode-matching partial sums differ" % rank


def test_steadystate_eval_and_merge():
    """oems_cuda_steadystate_eval (host arithmetic of Engine_Ext_SteadyState::Apply2Voltages,
    engine_ext_steadystate.cpp:62-106) against a numpy restatement, and the slab merge: energies add up,
    records are concatenated; the criterion is a maximum over the probes, so the slab order is free"""
    from openems_b200 import slabs
    rng = np.random.default_rng(9)
    p, cnt = 7, 5
    snap = rng.standard_normal((2 * p, cnt))
    snap[p:] = snap[:p] * (1 + 1e-3 * rng.standard_normal((p, cnt)))       # nearly periodic
    snap[:, 3] *= 1e-4                                                     # a probe below 1 % of the strongest: ignored
    info = np.array([3, 5 * p], np.uint32)                                 # rel_pos = 5p % 2p = p <= p: new = rows 0..p-1
    en = np.array([0.0, 0.0, 2.0, 2.002])
    new, old = snap[:p], snap[p:]
    cur = (new ** 2).sum(0)
    dif = ((old - new) ** 2).sum(0)
    want = max(abs(en[3] - en[2]) / en[2], max(dif[n] / cur[n] for n in range(cnt) if cur[n] > cur.max() * 1e-2))
    whole = slabs.combine_steadystate([(info, en, snap)], p)
    assert whole[1] == 3 and whole[0] == pytest.approx(want, rel=1e-12)
    # two slabs: probes 0..1 on the first, 2..4 on the second, each with a part of the energy
    parts = [(info, en * 0.25, snap[:, :2]), (info, en * 0.75, snap[:, 2:])]
    assert slabs.combine_steadystate(parts, p)[0] == pytest.approx(want, rel=1e-12)
    assert slabs.combine_steadystate(parts[::-1], p)[0] == pytest.approx(want, rel=1e-12)
    # no completed check yet -> 1
    assert slabs.combine_steadystate([(np.array([0, 0], np.uint32), en, snap)], p)[0] == 1.0
    assert slabs.combine_mode_match([(1.0, 0, 2.0), (3.0, 0, 6.0)]) == (4.0, 2.0)
