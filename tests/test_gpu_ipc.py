"""The multi-process CUDA-IPC slab path (what `bench.py --gpus N` and SCALE time) against the single-domain oracle:
2 and 3 ranks launched with torch.distributed.run, both timestep schedules.  Uses as many GPUs as the box has
(ranks share a device when there are fewer)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,fused,case", [(2, 1, "allpml"), (2, 0, "allpml"), (3, 1, "cavity")])
def test_ipc_slabs_equal_oracle(tmp_path, world, fused, case):
    out = tmp_path / "result.json"
    env = dict(os.environ, IPC_FUSED=str(fused), IPC_CASE=case, IPC_RESULT=str(out), IPC_STEPS="1,3,40", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "tests", "ipc_worker.py")]
    res = subprocess.run(cmd, env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    r = json.loads(out.read_text())
    assert r["ok"], r
    assert len(r["checked"]) == 6 and all(c["differing_values"] == 0 for c in r["checked"])
    assert r["checked"][-1]["max_abs"] > 0
    # field dumps across the process boundary (ghost planes completed through the IPC mappings)
    assert len(r["dumps"]) == 12 and all(c["differing_values"] == 0 for c in r["dumps"]) and r["dumps"][-1]["max_abs"] > 0
    assert ("fused_EH" in r["schedule"]) == bool(fused)
