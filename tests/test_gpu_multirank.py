"""The N > 1 path on real GPUs and real NCCL: launches torchrun with one rank per GPU (2 ranks when the box has at least
two GPUs; skipped on a single-GPU box — the driver's scaling run then covers it through bench.py's per-rank parity
report).  Per rank: forces of the functor path and of the device-resident tree step against the fp64 oracle."""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_ranks_nccl_let_exchange_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_gpu_multirank_worker.py"), "60000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    lines = [json.loads(l[len("MULTIRANK "):]) for l in out.stdout.splitlines() if l.startswith("MULTIRANK ")]
    print(out.stdout[-4000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert len(lines) == 2 and all(r["pass"] for r in lines)
    assert all(r["let_ep"] > 0 and r["let_sp"] > 0 and r["nccl_bytes"] > 0 for r in lines)
