"""world_size-2 (and 3) gloo runs of the multi-GPU exchange code on CPU."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from sequencedetectionqueryexecutor_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 60, size=1000)
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    for world in (1, 2, 3, 8):
        b = D.shard_bounds(off, world)
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0) and len(b) == world + 1
        ev = np.array([off[b[r + 1]] - off[b[r]] for r in range(world)])
        assert ev.sum() == off[-1] and ev.max() - ev.min() <= 2 * 60
    assert D.shard_bounds(np.zeros(1, dtype=np.int64), 4).tolist() == [0, 0, 0, 0, 0]


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_allgather_and_allreduce(world, tmp_path):
    out = tmp_path / "r.json"
    port = 29650 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(out.read_text())
    assert res["world"] == world and res["ok_detect"], res
    assert res["ok_declare"] and res["ok_stats"] and res["n_traces"] > 10
