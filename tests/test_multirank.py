"""CPU test of the N>1 path: world_size 2 and 4 over gloo (host logic of petar_b200/multigpu.py)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_domain_decomposition_and_let_exchange_gloo(world):
    port = _free_port()
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_multirank_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(" OK ") == world, out.stdout[-2000:]


def test_domain_split_balanced():
    import numpy as np
    from petar_b200 import multigpu
    rng = np.random.default_rng(0)
    pos = rng.normal(size=(10001, 3))
    for w in (1, 2, 4, 8, 3):
        o = multigpu.domain_split(pos, w)
        c = np.bincount(o, minlength=w)
        assert len(c) == w and c.max() - c.min() <= w
