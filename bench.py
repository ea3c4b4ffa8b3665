#!/usr/bin/env python
"""bench.py -- FDTD MCells/s of the B200 engine on BASELINE.json's headline config.

Workload (config C5 of BASELINE.md): synthetic uniform Cartesian vacuum mesh 1024^3, unit 1 mm,
PML_8 on all six faces, centre E_z soft Gauss source, 3 voltage + 3 current + 6 field probes.
A "step" is one FDTD timestep (E half-step + H half-step + all extension hooks) over the whole
mesh.  MCells/s = Nx*Ny*Nz*timesteps / seconds / 1e6 (openems.cpp:1485 definition).

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                   the reference algorithm on the host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fdtd_mcells_per_s"
UNIT = "MCells/s"
PML = 8


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """samples SM clock + throttle reasons with nvidia-smi while the timed region runs"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_c5(n, slab=None, reduce_min=None):
    """host side: build the compressed operator of the C5 mesh.  With a slab (one process per GPU) only the planes
    the rank holds are built; the ranks agree on the timestep through `reduce_min` (a MIN all-reduce of a float)"""
    from openems_b200 import SyntheticOperator
    from openems_b200.synthetic import BC_PML, EXC_E_SOFT
    C0 = 299792458.0
    nx, ny, nz = n
    lines = tuple(np.arange(m, dtype=np.float64) for m in (nx, ny, nz))  # drawing unit mm
    so = SyntheticOperator(*lines, 1e-3)
    so.set_bc([BC_PML] * 6, (PML,) * 6)
    fc = C0 / (20 * 1e-3)
    so.set_excite_gauss(fc / 2, fc / 2)
    c = (nx // 2, ny // 2, nz // 2)
    so.add_excitation((c[0], c[1], c[2] + 0.5), (c[0], c[1], c[2] + 0.5), EXC_E_SOFT, (0, 0, 1))
    t0 = time.time()
    if slab is not None:
        so.set_slab(*slab)
        so.set_timestep(reduce_min(so.local_timestep()))
    so.build()
    t_build = time.time() - t0
    return so, t_build


def add_c5_probes(eng, n):
    nx, ny, nz = n
    c = (nx // 2, ny // 2, nz // 2)
    q = (nx // 4, ny // 4, nz // 4)
    for a in range(3):  # 3 voltage probes: 16-edge lines through the quarter point
        stop = list(q)
        stop[a] += 16
        eng.AddVoltageProbe(q, stop)
    for a in range(3):  # 3 current probes: 16x16 loops around the centre
        start, stop = list(c), list(c)
        for b in range(3):
            if b != a:
                start[b] -= 8
                stop[b] += 8
        eng.AddCurrentProbe(start, stop, a)
    for k in range(3):  # 6 field probes
        p = (q[0] + 5 * k, q[1] + 3 * k, q[2] + 7 * k)
        eng.AddFieldProbe(0, p)
        eng.AddFieldProbe(1, p)


def algorithmic_bytes(n, pml_cells, index_bytes, one_pass=False):
    """SURVEY 8(d).  Two-pass schedule: per half-step 36 B field traffic + index per cell, + 24 B
    per PML cell (flux r/w).  One-pass schedule (k_fused_EH): E and H read once and written once
    = 48 B + index per cell and launch; per step the UPML flux r/w of both half-steps on top."""
    cells = n[0] * n[1] * n[2]
    if one_pass:
        per_launch = cells * (48 + index_bytes)
        return per_launch, per_launch + pml_cells * 48
    per_half = cells * (36 + index_bytes) + pml_cells * 24
    return per_half, 2 * per_half


class quiet_stdout:
    """the reference prints banners to stdout (\"Create FDTD operator\" ...): route fd 1 to stderr meanwhile, the bench
    line must be the only thing on stdout"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_baseline(sample_n, steps, threads, warm=3):
    with quiet_stdout():
        return _cpu_baseline(sample_n, steps, threads, warm)


def _cpu_baseline(sample_n, steps, threads, warm=3):
    """the reference's CPU engine on a bounded sample of the same workload (same BC, source, timestep rule).
    kind "reference": oracle/_ref/libopenems_ref.so = the UNMODIFIED reference translation units (operator build,
    Engine_Multithread -- the reference's default engine --, UPML extension) compiled by oracle/Makefile.ref, run with
    `threads` threads.  Fallback kind "port": the sse-compressed multithreaded restatement (oracle/fdtd_oracle_sse.c).
    -> (MCells/s, seconds, kind, description)"""
    from oracle.pyoracle import OracleSim, OracleSSE, BC_PML, EXC_E_SOFT
    C0 = 299792458.0
    lines = tuple(np.arange(m, dtype=np.float64) for m in sample_n)
    kind, s = "port", None
    try:
        from oracle import pyref
        if pyref.available():
            s = pyref.RefSim(*lines, 1e-3, engine=pyref.ENGINE_MULTITHREADED, threads=threads)
            kind = "reference"
    except Exception:
        s = None
    if s is None:
        s = OracleSim(*lines, 1e-3)
    s.set_bc([BC_PML] * 6, (PML,) * 6)
    fc = C0 / (20 * 1e-3)
    s.set_excite_gauss(fc / 2, fc / 2)
    c = tuple(m // 2 for m in sample_n)
    s.add_excitation((c[0], c[1], c[2] + 0.5), (c[0], c[1], c[2] + 0.5), EXC_E_SOFT, (0, 0, 1))
    t0 = time.time()
    s.build()
    t_build = time.time() - t0
    eng = s if kind == "reference" else OracleSSE(s, threads=threads)
    eng.iterate(warm)
    t0 = time.time()
    eng.iterate(steps)
    dt = time.time() - t0
    cells = sample_n[0] * sample_n[1] * sample_n[2]
    what = ("unmodified reference (oracle/_ref: Operator_Multithread + Engine_Multithread, %d threads)" % threads) if kind == "reference" \
        else ("sse-compressed multithreaded restatement (oracle/fdtd_oracle_sse.c, %d threads)" % threads)
    return cells * steps / dt / 1e6, dt, kind, what + ", operator build %.1f s" % t_build


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = tuple(args.cpu_sample)
    # each "step" of this arm = one timestep on the bounded sample mesh
    steps = max(1, min(args.steps, 400))
    val, dt, kind, what = cpu_baseline(sample, steps, threads, warm=max(1, args.warmup))
    n = tuple(args.n)
    try:
        import psutil
        ram_gb = round(psutil.virtual_memory().total / 2 ** 30, 1)
    except Exception:
        ram_gb = None
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C5 uniform vacuum %dx%dx%d PML_8x6 centre Ez Gauss source, 12 probes" % n,
                   "timed_on": "bounded sample %dx%dx%d of the same workload (same BC, source, timestep rule), all host cores; "
                               "the reference's own operator build of the full 1024^3 mesh needs ~150 B/cell of host memory and "
                               "tens of minutes (single-threaded parts), so the headline size is not run on the CPU: "
                               "same_config is false by construction, MCells/s is per cell" % sample,
                   "host_ram_gb": ram_gb},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "%dx%dx%d PML_8 mesh, %d timesteps; %s" % (sample + (steps, what))},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def parity_check(world, rank, local_rank, dist, reduce_min, reduce_sum_u64):
    """the timed path on a small mesh, in this very run: N>1: z-slab engines in N processes (CUDA-IPC halos) against ONE
    single-GPU engine of the same mesh on rank 0; N=1: the one-pass schedule against the two-pass schedule.  Both
    sides start from the same deterministic pre-fill and must end with identical E/H digests.  (That the single-GPU
    engine equals the reference bit for bit is what tests/ establish.)"""
    from openems_b200.slabs import slab_range
    n = (192, 160, max(64, 48 * world))
    steps = 24
    slab = slab_range(n[2], world, rank, pml_lo=PML, pml_hi=PML, pml_weight=1.7) if world > 1 else None
    so, _ = build_c5(n, slab=slab, reduce_min=reduce_min)
    eng = so.operator().CreateEngine(device=local_rank, slab=slab)
    eng.SetOption("fused", 1)
    if world > 1:
        blobs = [None] * world
        dist.all_gather_object(blobs, eng.ExportIPC())
        eng.OpenPeers(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None)
        dist.barrier()
    one_pass = bool(eng.GetOption("fused"))
    eng.FillFields(7)
    eng.IterateTS(steps)
    eng.Synchronize()
    got = reduce_sum_u64(eng.FieldDigest())
    eng.close()
    if world > 1:
        dist.barrier()
    want = None
    if rank == 0:
        so1, _ = build_c5(n)
        ref = so1.operator().CreateEngine(device=local_rank)
        ref.SetOption("fused", 1 if world > 1 else 0)
        ref.FillFields(7)
        ref.IterateTS(steps)
        want = list(ref.FieldDigest())
        ref.close()
    return {"mesh": "%dx%dx%d PML_8" % n, "timesteps": steps,
            "compared": ("%d z-slab processes (CUDA-IPC halos, %s schedule) vs one single-GPU engine" % (world, "one-pass" if one_pass else "two-pass"))
            if world > 1 else "one-pass schedule vs two-pass schedule, single GPU",
            "digest_E": "%016x" % got[0], "digest_H": "%016x" % got[1],
            "equal": None if want is None else bool(want == list(got))}


def run_gpu(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:  # keep stdout to the one JSON line (NCCL prints its version there at VERSION/INFO level)
        os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1 and "BENCH_KEEP_OMP" not in os.environ:
        # the host-side operator build is OpenMP-parallel and very uneven between the ranks (the end slabs hold the
        # graded z-PML planes: 12-15 distinct xy planes to evaluate, a middle slab 1-2): every rank may use all cores
        # (the middle ranks are done in a fraction of a second), waiting threads sleep.  torchrun presets
        # OMP_NUM_THREADS=1, which made the end ranks build single-threaded (16 s at N=2).
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        os.environ["OMP_WAIT_POLICY"] = "passive"
    import torch
    from openems_b200 import load_library
    load_library()  # fail loudly when the CUDA library is missing
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = tuple(args.n)
    nz = n[2]
    slab = None
    if world > 1:
        # strong scaling: z-slabs of the same mesh.  A z-UPML plane (all cells UPML: shell E with its second store +
        # shell H = 136 B/cell) costs about 2.7x a plain plane (50 B/cell), so the two end slabs get fewer planes;
        # the first split uses that estimate (weight 1.7), N > 2 then re-deals the planes from measured busy times
        from openems_b200.slabs import slab_range
        slab = slab_range(nz, world, rank, pml_lo=PML, pml_hi=PML, pml_weight=args.pml_weight)

    def reduce_min(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    def reduce_sum_u64(vals):
        """sum mod 2^64 over the ranks (digests of slab engines add up to the single-GPU digest)"""
        if world == 1:
            return list(vals)
        parts = [None] * world
        dist.all_gather_object(parts, list(vals))
        return [sum(p[i] for p in parts) % (1 << 64) for i in range(len(vals))]

    # every rank builds only the planes it holds (the ranks agree on the timestep by a MIN reduction)
    so, t_build = build_c5(n, slab=slab, reduce_min=reduce_min if world > 1 else None)
    op = so.operator()

    # ---------------- slab sizes from MEASURED per-rank kernel times (N > 2): a short un-timed run on the first split
    # gives every rank's busy time per timestep (all kernels but the halo waits); middle ranks give the cost of a
    # plain plane, the end ranks the extra cost of a z-PML plane; the planes are re-dealt with that weight and the
    # ranks whose range changed rebuild their operator
    balance = None
    if world > 2 and not args.no_rebalance:
        from openems_b200.slabs import slab_range
        e0 = op.CreateEngine(device=local_rank, slab=slab)
        for kv in args.opt:
            k, v = kv.split("=")
            e0.SetOption(k, int(v))
        blobs = [None] * world
        dist.all_gather_object(blobs, e0.ExportIPC())
        e0.OpenPeers(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None)
        dist.barrier()
        e0.FillFields(0)
        e0.IterateTS(3)
        e0.Synchronize()
        # (join_side = the main stream waiting for the side stream, i.e. for the neighbour's E plane: a wait as well)
        busy = sum(ms for name, ms in e0.TimeSchedule(4) if not name.startswith("halo_wait") and name != "join_side")
        e0.close()
        del e0
        parts = [None] * world
        dist.all_gather_object(parts, (busy, slab[1] - slab[0]))
        mid = [b / m for b, m in parts[1:-1]]
        a = sum(mid) / len(mid)                                            # ms per plain plane
        b = sum(p[0] - a * p[1] for p in (parts[0], parts[-1])) / 2 / (PML + 1)   # extra ms per z-PML plane
        w = max(0.0, b / a)
        new_slab = slab_range(nz, world, rank, pml_lo=PML, pml_hi=PML, pml_weight=w)
        balance = {"first_split_weight": args.pml_weight, "busy_ms_per_rank": [round(p[0], 4) for p in parts],
                   "planes_first_split": [p[1] for p in parts], "measured_weight": round(w, 3)}
        changed = torch.tensor([int(tuple(new_slab) != tuple(slab))], device="cuda")
        dist.all_reduce(changed)
        if int(changed.item()):
            slab = new_slab
            so, t_build2 = build_c5(n, slab=slab, reduce_min=reduce_min)
            t_build = max(t_build, t_build2)
            op = so.operator()
        allp = [None] * world
        dist.all_gather_object(allp, slab[1] - slab[0])
        balance["planes"] = allp

    # ---------------- e2e leg: engine creation from HOST buffers + K timesteps with probe readback
    def make_engine():
        eng = op.CreateEngine(device=local_rank, slab=slab)
        for kv in args.opt:                      # experiments: engine options, e.g. --opt overlap_halo=0
            k, v = kv.split("=")
            eng.SetOption(k, int(v))
        add_c5_probes(eng, n)
        return eng

    burst = max(1, so.nyquist // 4)  # Processing interval: Nyquist / OverSampling(4), openems.cpp:568

    def link(eng):
        if world == 1:
            return
        blob = eng.ExportIPC()
        blobs = [None] * world
        dist.all_gather_object(blobs, blob)
        eng.OpenPeers(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None)
        dist.barrier()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- parity gate of this run: the path that is timed below, on a small mesh
    parity = parity_check(world, rank, local_rank, dist, reduce_min if world > 1 else None, reduce_sum_u64)

    eng = make_engine()
    link(eng)
    stats0 = eng.GetStats()
    pml_cells = stats0["pml_cells"]
    index_bytes = stats0["index_bytes"]

    # ---------------- device-resident leg: the timed window starts from a FILLED domain (deterministic pre-fill,
    # a function of the global cell index: the same state at every N), not from the zeros a point source leaves
    eng.FillFields(0)
    eng.IterateTS(args.warmup)
    eng.Synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    k0 = eng.GetStats()["kernels_launched"]
    # timed on the device: CUDA events on the engine's own stream around exactly K timesteps
    t_dev = eng.IterateTimed(args.steps) * 1e-3
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.GetStats()["kernels_launched"] - k0
    # digest of E and H after warmup + K timesteps: identical for every N (z-slab digests add up)
    dig = reduce_sum_u64(eng.FieldDigest())
    energy_after = eng.CalcFastEnergy() if world == 1 else None
    if world > 1:
        t = torch.tensor([t_dev], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev = float(t.item())
    cells = n[0] * n[1] * n[2]
    value = cells * args.steps / t_dev / 1e6

    # ---------------- per-kernel timing (CUDA events on the engine's stream) for the roofline
    sched = eng.TimeSchedule(min(10, max(3, args.steps // 10)))
    local_cells = cells if slab is None else n[0] * n[1] * (slab[1] - slab[0])
    peak, peak_src = measured_peak()
    kern = {}
    for name, ms in sched:
        kern[name] = kern.get(name, 0.0) + ms
    one_pass = "fused_EH" in kern
    per_half, per_step = algorithmic_bytes((n[0], n[1], local_cells // (n[0] * n[1])), pml_cells, index_bytes, one_pass)
    if one_pass:
        # the one-pass kernel works on the interior rows x planes only (all-UPML planes / rows at the mesh ends go
        # through the shell launches): its algorithmic bytes are those cells x (48 + index)
        rows, planes = eng.GetOption("onepass_rows"), eng.GetOption("onepass_planes")
        per_half = n[0] * rows * planes * (48 + index_bytes)
    if one_pass:
        dom, t_dom = "fused_EH", kern["fused_EH"]
    else:
        t_E, t_H = kern.get("update_E", 0.0), kern.get("update_H", 0.0)
        dom = "update_E" if t_E >= t_H else "update_H"
        t_dom = max(t_E, t_H)
    achieved = per_half / (t_dom * 1e-3) / 1e9 if t_dom > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
            key = "%dx%dx%d" % n
            if world == 1 and key in tj and dom in tj[key]:
                traffic = tj[key][dom]
    except Exception:
        pass
    step_ms_sched = sum(ms for _, ms in sched)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "algorithmic_bytes_per_launch": per_half, "kernel_ms": t_dom, "peak_source": peak_src,
                "kernel_cells": (n[0] * rows * planes) if one_pass else local_cells,
                "kernels_ms": {k: round(v, 5) for k, v in kern.items()}, "step_ms_from_events": step_ms_sched,
                "step_frac_of_peak": (per_step / (step_ms_sched * 1e-3) / 1e9 / peak) if step_ms_sched else None}

    # ---------------- e2e: new engine from host buffers, K steps in bursts, probes read to host
    eng.close()
    del eng
    barrier()
    # The engine that was just closed held 55 GB; the driver hands freed memory back asynchronously, and an allocation
    # of the same size right behind the free waits for that (engine creation then measured 0.03 ... 0.35 s from run to
    # run).  An application creates its engine on an idle device: the CPU baseline leg (N = 1, ~10 s of host work) or a
    # short pause goes in between, untimed.
    cpu_leg = None
    if world == 1 and not args.no_cpu and rank == 0:
        threads = os.cpu_count() or 1
        cpu_leg = cpu_baseline(tuple(args.cpu_sample), args.cpu_steps, threads) + (threads,)
    else:
        time.sleep(2.0)
    barrier()
    t0 = time.perf_counter()
    eng = make_engine()
    link(eng)
    t_upload = time.perf_counter() - t0
    # bytes copied host->device while creating the engine: coefficient tuples + the operator index as the host
    # builder holds it -- unique xy planes and one plane id per z, expanded on the device
    # (oems_cuda_set_operator_planes) -- per rank (+ a few KB of signal / excitation / probe lists, ignored)
    # counted by the engine at its copy calls (option "h2d_bytes"), summed over the ranks
    h2d = eng.GetOption("h2d_bytes")
    if world > 1:
        t = torch.tensor([h2d], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        h2d = int(t.item())
    done, d2h = 0, 0
    while done < args.steps:
        m = min(burst, args.steps - done)
        eng.IterateTS(m)
        vals = eng.ReadProbes()  # D2H of the probe values, synchronises
        d2h += vals.nbytes
        done += m
    eng.Synchronize()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_val = cells * args.steps / t_e2e / 1e6
    probe_sample = [float(v) for v in vals[:4]]

    # ---------------- asynchronous field dumps (north_star (4), N = 1): an NF2FF-style box -- 6 faces x E and H, cell
    # interpolation -- sampled after EVERY burst with oems_cuda_read_dump_async while the time loop goes on; the host
    # waits for a sample (and copies it out of the page-locked buffer) only while the next burst is running
    dump_info = None
    if world == 1 and not args.no_dump_leg:
        m = 32
        lo, hi = [m, m, m], [n[0] - 1 - m, n[1] - 1 - m, n[2] - 1 - m]
        el = [np.full(k, 1e-3) for k in n]
        ids = []
        for a in range(3):
            for side in (lo[a], hi[a]):
                rng = [np.arange(lo[b], hi[b] + 1) if b != a else np.array([side]) for b in range(3)]
                for is_H in (0, 1):
                    ids.append(eng.AddDump(is_H, 2, rng[0], rng[1], rng[2], el, el))
        nbytes = sum(int(np.prod(eng._dump_shapes[i])) * 4 for i in ids)
        steps_d = max(burst * 10, (args.steps // burst) * burst)

        def run(with_dumps):
            eng.Synchronize()
            t0 = time.perf_counter()
            pending, done, got = [], 0, 0
            while done < steps_d:
                eng.IterateTS(burst)
                done += burst
                if with_dumps:
                    # the PREVIOUS sample is taken out of the page-locked buffers while this burst runs on the GPU
                    for t in pending:
                        got += eng.WaitDump(t).nbytes
                    # this burst's sample: evaluated on the engine stream behind the burst, copied on the copy stream
                    pending = [eng.ReadDumpAsync(i) for i in ids]
                else:
                    eng.ReadProbes()
            eng.Synchronize()                       # the time loop (and the last sample's evaluation) is through here
            t_loop = time.perf_counter() - t0
            for t in pending:                       # the last sample still has to arrive and be taken out
                got += eng.WaitDump(t).nbytes
            return t_loop / steps_d * 1e3, (time.perf_counter() - t0) / steps_d * 1e3, got

        run(True)
        ms_plain, _, _ = run(False)
        ms_dump, ms_dump_drained, got = run(True)
        dump_info = {"box": "6 faces %d cells inside the mesh, E and H, cell interpolation" % m, "samples": steps_d // burst,
                     "bytes_per_sample": nbytes, "every_timesteps": burst, "ms_per_step_without": ms_plain, "ms_per_step_with": ms_dump,
                     "overhead_frac": ms_dump / ms_plain - 1.0,
                     "ms_per_step_with_incl_last_sample_drain": ms_dump_drained, "d2h_bytes_total": got,
                     "note": "overhead_frac = what the dumps cost the time loop (their 12 interpolation kernels per sample + copy-stream "
                             "contention); the last sample's D2H copy and host memcpy after the loop is reported separately"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5 uniform vacuum %dx%dx%d PML_8x6 centre Ez Gauss source, 12 probes" % n,
                       "parallelism": "z-slabs x%d over NVLink peer memory" % world if world > 1 else "single GPU",
                       "schedule": "one-pass (k_fused_EH + UPML shell)" if one_pass else "two-pass (k_update_E, k_update_H)",
                       "l2": "inputs (%.1f GB of fields+index per GPU) far larger than the 126 MB L2; no flush needed"
                             % ((24 + index_bytes) * local_cells / 1e9),
                       "n_unique_coeff_tuples": so.n_unique, "index_bytes": index_bytes, "pml_cells": pml_cells,
                       "host_operator_build_s": round(t_build, 2), "burst_ts": burst,
                       "operator_index": "%d unique xy planes + plane ids (%.1f MB), expanded on the device" % (so.unique_planes, so.unique_planes * n[0] * n[1] * index_bytes / 1e6)},
            "roofline": roofline,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                    "includes": "engine creation from host buffers (operator H2D %.2f s) + %d timesteps in bursts of %d with "
                                "probe read-back" % (t_upload, args.steps, burst), "probe_sample": probe_sample},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "slab_balance": balance,
            "async_dumps": dump_info,
            "parity_check": dict(parity, full_size_digest={"timesteps": args.warmup + args.steps, "prefill_seed": 0,
                                                           "E": "%016x" % dig[0], "H": "%016x" % dig[1],
                                                           "note": "same value at every N for the same --warmup/--steps (bit-exact slabs)"}),
        }
        if energy_after is not None:
            out["parity_check"]["full_size_digest"]["energy_estimate_J"] = energy_after
        if cpu_leg is not None:
            v, dt, kind, what, threads = cpu_leg
            sample = tuple(args.cpu_sample)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                                   "sample": "%dx%dx%d PML_8 mesh, %d timesteps in %.1f s; %s" % (sample + (args.cpu_steps, dt, what))}
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def build_c4(n):
    """BASELINE config C4: uniform mesh, PML_8 x6, central Drude eps+mue block of half the mesh width (f_p 5 GHz, tau 5 ns:
    matlab/examples/other/Metamaterial_PlaneWave_Drude.m:28-33,75-76), plane E_y source below the block"""
    from openems_b200 import SyntheticOperator
    from openems_b200.synthetic import BC_PML, EXC_E_SOFT
    lines = tuple(np.arange(m, dtype=np.float64) for m in n)
    so = SyntheticOperator(*lines, 1e-3)
    so.set_bc([BC_PML] * 6, (PML,) * 6)
    so.set_excite_gauss(5e9, 5e9)
    a = [m // 4 for m in n]
    b = [m - m // 4 for m in n]
    so.add_lorentz(tuple(a), tuple(b), eps_fp=(5e9,), eps_tau=(5e-9,), mue_fp=(5e9,), mue_tau=(5e-9,))
    so.add_excitation((10, 10, 10), (n[0] - 11, n[1] - 11, 10), EXC_E_SOFT, (0, 1, 0))
    t0 = time.time()
    so.build()
    return so, time.time() - t0


def c4_cpu_baseline(sample_n, steps, threads):
    """the reference's own engine (oracle/_ref, Engine_Multithread + Engine_Ext_LorentzMaterial) on a bounded sample"""
    from oracle.pyoracle import OracleSim, OracleSSE, BC_PML, EXC_E_SOFT
    lines = tuple(np.arange(m, dtype=np.float64) for m in sample_n)
    kind, s = "port", None
    with quiet_stdout():
        try:
            from oracle import pyref
            if pyref.available():
                s = pyref.RefSim(*lines, 1e-3, engine=pyref.ENGINE_MULTITHREADED, threads=threads)
                kind = "reference"
        except Exception:
            s = None
        if s is None:
            s = OracleSim(*lines, 1e-3)
        s.set_bc([BC_PML] * 6, (PML,) * 6)
        s.set_excite_gauss(5e9, 5e9)
        a = tuple(m // 4 for m in sample_n)
        b = tuple(m - m // 4 for m in sample_n)
        s.add_lorentz(a, b, eps_fp=(5e9,), eps_tau=(5e-9,), mue_fp=(5e9,), mue_tau=(5e-9,))
        s.add_excitation((10, 10, 10), (sample_n[0] - 11, sample_n[1] - 11, 10), EXC_E_SOFT, (0, 1, 0))
        s.build()
        eng = s if kind == "reference" else OracleSSE(s, threads=threads)
        eng.iterate(3)
        t0 = time.time()
        eng.iterate(steps)
        dt = time.time() - t0
    return sample_n[0] * sample_n[1] * sample_n[2] * steps / dt / 1e6, dt, kind


def run_gpu_c4(args):
    """--config c4 (single GPU): the extension-heavy path.  One-pass schedule with the ADE applied inside the kernel.
    Traffic model per timestep: (48 B + index) per cell + 48 B per UPML cell + 96 B per dispersive cell and ADE pair
    (auxiliary value read + written and its two coefficients read, for E and for H)."""
    import torch
    from openems_b200 import load_library
    load_library()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("bench.py --config c4 is a single-GPU line")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    n = tuple(args.n)
    so, t_build = build_c4(n)
    op = so.operator()
    cells = n[0] * n[1] * n[2]
    eng = op.CreateEngine(device=0)
    st = eng.GetStats()
    disp = so.lorentz_counts()[0]
    eng.FillFields(0)
    eng.IterateTS(args.warmup)
    eng.Synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    k0 = eng.GetStats()["kernels_launched"]
    t_dev = eng.IterateTimed(args.steps) * 1e-3
    clocks = sampler.stop()
    launches = eng.GetStats()["kernels_launched"] - k0
    dig = eng.FieldDigest()
    value = cells * args.steps / t_dev / 1e6
    # the other schedule from the same state: digests must agree (parity of the timed path, in this run)
    one_pass = bool(eng.GetOption("fused"))
    sched = eng.TimeSchedule(min(10, max(3, args.steps // 10)))
    kern = {}
    for name, ms in sched:
        kern[name] = kern.get(name, 0.0) + ms
    eng2 = op.CreateEngine(device=0)
    eng2.SetOption("fused", 0)
    eng2.FillFields(0)
    t_two = eng2.IterateTimed(args.warmup + args.steps) * 1e-3 / (args.warmup + args.steps)
    dig2 = eng2.FieldDigest()
    eng2.close()
    peak, peak_src = measured_peak()
    ib = st["index_bytes"]
    per_step = cells * (48 + ib) + st["pml_cells"] * 48 + disp * 96
    dom = "fused_EH" if one_pass else "update_E"
    per_launch = cells * (48 + ib) + disp * 24 if one_pass else cells * (36 + ib) + st["pml_cells"] * 24
    t_dom = kern.get(dom, 0.0)
    achieved = per_launch / (t_dom * 1e-3) / 1e9 if t_dom else 0.0
    step_ms = sum(kern.values())
    # e2e: engine from host buffers + K steps (the CPU leg goes between the close and the new engine: see run_gpu)
    eng.close()
    torch.cuda.synchronize()
    cpu_leg = None
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        cpu_leg = c4_cpu_baseline(tuple(args.cpu_sample), min(args.cpu_steps, 60), threads) + (threads,)
    else:
        time.sleep(2.0)
    t0 = time.perf_counter()
    eng = op.CreateEngine(device=0)
    h2d = eng.GetOption("h2d_bytes")
    p = (n[0] // 2, n[1] // 2, n[2] // 2)
    eng.AddFieldProbe(0, p)
    eng.AddFieldProbe(1, p)
    done, d2h, burst = 0, 0, max(1, so.nyquist // 4)
    while done < args.steps:
        m = min(burst, args.steps - done)
        eng.IterateTS(m)
        d2h += eng.ReadProbes().nbytes
        done += m
    eng.Synchronize()
    t_e2e = time.perf_counter() - t0
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4 %dx%dx%d PML_8x6, central Drude eps+mue block %d^3 (%d dispersive cells), plane Ey Gauss source" % (n + (n[0] // 2, disp)),
                   "parallelism": "single GPU",
                   "schedule": "one-pass, ADE applied inside k_fused_tma (LOR instance) + lorentz_pre list kernels" if one_pass else "two-pass + Lorentz list kernels",
                   "l2": "inputs far larger than the 126 MB L2; no flush needed", "pml_cells": st["pml_cells"], "index_bytes": ib,
                   "host_operator_build_s": round(t_build, 2), "two_pass_ms_per_step": t_two * 1e3},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "traffic": None, "algorithmic_bytes_per_launch": per_launch, "kernel_ms": t_dom, "peak_source": peak_src,
                     "kernels_ms": {k: round(v, 5) for k, v in kern.items()}, "step_ms_from_events": step_ms,
                     "algorithmic_bytes_per_step": per_step,
                     "step_frac_of_peak": per_step / (step_ms * 1e-3) / 1e9 / peak if step_ms else None},
        "e2e": {"value": cells * args.steps / t_e2e / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                "includes": "engine creation from host buffers + %d timesteps in bursts of %d with probe read-back" % (args.steps, burst)},
        "gpu_launches": int(launches), "clocks": clocks,
        "parity_check": {"compared": "one-pass (ADE in the kernel) vs two-pass schedule, %d timesteps from the same pre-fill" % (args.warmup + args.steps),
                         "digest_E": "%016x" % dig[0], "digest_H": "%016x" % dig[1], "equal": bool(list(dig) == list(dig2))},
    }
    if cpu_leg is not None:
        sample = tuple(args.cpu_sample)
        v, dt, kind, threads = cpu_leg
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                               "sample": "%dx%dx%d C4 mesh (Drude block %d^3), %d timesteps in %.1f s, %s" % (sample + (sample[0] // 2, min(args.cpu_steps, 60), dt, "unmodified reference engine (oracle/_ref)" if kind == "reference" else "sse restatement"))}
    print(json.dumps(out))
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--mesh", dest="n", type=int, nargs=3, default=[1024, 1024, 1024], help="mesh lines (default: the 1024^3 headline config)")
    ap.add_argument("--cpu-sample", type=int, nargs=3, default=None,
                    help="bounded sample mesh of the CPU arm (default 192^3 inside the GPU line, 256^3 for --impl reference)")
    ap.add_argument("--cpu-steps", type=int, default=100)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value for the timed engine (experiments)")
    ap.add_argument("--no-dump-leg", action="store_true", help="skip the asynchronous-dump overhead measurement (N = 1)")
    ap.add_argument("--pml-weight", type=float, default=1.7,
                    help="first z-slab split: extra cost of a z-PML plane relative to a plain plane (136 vs 50 bytes per cell)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="label of the line: weak when --n grows with --gpus (e.g. 1024 1024 1024*N)")
    ap.add_argument("--no-rebalance", action="store_true", help="keep the first split (N > 2: otherwise re-dealt from measured busy times)")
    ap.add_argument("--config", default="c5", choices=["c5", "c4"], help="c5: the headline (default); c4: 512^3 Drude block")
    args = ap.parse_args()
    if args.config == "c4" and args.n == [1024, 1024, 1024]:
        args.n = [512, 512, 512]
    if args.warmup < 3:
        args.warmup = 3
    if args.cpu_sample is None:
        args.cpu_sample = [256, 256, 256] if args.impl == "reference" else [192, 192, 192]
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c4":
        run_gpu_c4(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
