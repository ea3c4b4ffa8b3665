"""One rank of the multi-PROCESS z-slab path that bench.py --gpus N uses: engine per process, halo planes pushed into
the neighbour's memory through CUDA-IPC mappings (Engine::export_ipc / open_peers).  Launched by
tests/test_gpu_ipc.py through torch.distributed.run; the ranks share one GPU when the box has fewer GPUs than
ranks (CUDA IPC works between processes on one device; the halo waits then rely on time slicing).
Every rank steps its slab; rank 0 gathers the owned planes and compares E and H bit for bit with the oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle.pyoracle import BC_PML, BC_MUR, BC_PMC, BC_PEC
    from tests import cases
    from tests.gpu_util import operator_from_oracle
    from openems_b200.slabs import slab_range, held_range

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    fused = int(os.environ.get("IPC_FUSED", "1"))
    case = os.environ.get("IPC_CASE", "allpml")
    steps = [int(x) for x in os.environ.get("IPC_STEPS", "1,3,40").split(",")]
    dist.init_process_group("gloo")   # rendezvous only (two ranks may share one device, which NCCL refuses)
    dev = rank % torch.cuda.device_count()
    if case == "cavity":
        s = cases.engine_cavity()
    else:
        s = cases.uniform_box(n=(40, 36, 16 * world + 12), bc=(BC_PML,) * 6, pml=8)
    nz = s.N[2]
    zb, ze = slab_range(nz, world, rank, pml_lo=8, pml_hi=8, pml_weight=2.4)
    eng = operator_from_oracle(s).CreateEngine(device=dev, slab=(zb, ze))
    eng.SetOption("fused", fused)
    eng.SetOption("halo_timeout_s", 120)
    blobs = [None] * world
    dist.all_gather_object(blobs, eng.ExportIPC())
    eng.OpenPeers(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None)
    dist.barrier()
    result = {"ok": True, "schedule": [n for n, _ in eng.TimeSchedule(0)], "checked": [], "dumps": []}
    # readout across the process boundary: node-/cell-interpolated E and H dumps of the whole mesh (they read the
    # neighbour rank's plane, completed through the IPC mappings by oems_cuda_exchange_ghosts)
    from openems_b200.slabs import read_dump_distributed
    el = [[s.edge_length(a, [p if b == a else 0 for b in range(3)], False) for p in range(s.N[a])] for a in range(3)]
    dl = [[s.edge_length(a, [p if b == a else 0 for b in range(3)], True) for p in range(s.N[a])] for a in range(3)]
    dump_kinds = [(0, 1), (1, 1), (0, 2), (1, 2)]
    dump_ids = [eng.AddDump(is_H, interp, np.arange(s.N[0]), np.arange(s.N[1]), np.arange(nz), el, dl) for is_H, interp in dump_kinds]
    for n in steps:
        eng.IterateTS(n)          # a whole burst per process: the flags order the halos, not the hosts
        eng.Synchronize()
        h0, h1 = held_range(nz, zb, ze)
        mine = [eng.GetFields(w)[..., zb - h0: ze - h0].copy() for w in (0, 1)]
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((zb, ze, mine), gathered, dst=0)
        if rank == 0:
            s.iterate(n)
            for w, ref in ((0, s.volt), (1, s.curr)):
                got = np.zeros_like(ref)
                for b, e, f in gathered:
                    got[..., b:e] = f[w]
                bad = int((got.view(np.uint32) != ref.view(np.uint32)).sum())
                result["checked"].append({"ts": int(s.num_ts), "field": w, "differing_values": bad, "max_abs": float(np.abs(ref).max())})
                if bad:
                    result["ok"] = False
        for (is_H, interp), d in zip(dump_kinds, dump_ids):
            got = read_dump_distributed(eng, d, dist, world)
            if rank == 0:
                ref = s.dump_field(is_H, interp, (0, 0, 0), tuple(m - 1 for m in s.N))
                bad = int((got.view(np.uint32) != ref.view(np.uint32)).sum())
                result["dumps"].append({"ts": int(s.num_ts), "is_H": is_H, "interp": interp, "differing_values": bad, "max_abs": float(np.abs(ref).max())})
                if bad:
                    result["ok"] = False
        dist.barrier()
    eng.close()
    if rank == 0:
        with open(os.environ["IPC_RESULT"], "w") as f:
            json.dump(result, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
