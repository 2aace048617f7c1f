"""N > 1 path: host-side sharding logic on CPU with world_size 2 over gloo; the NCCL exchange inside
the library on 2 GPUs (skipped when fewer are visible)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from theia_b200 import sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(mode, world, out_dir, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multi_worker.py"), mode, str(out_dir)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_owner_rank_mirror():
    t = synth.make_flows(500, 20, seed=3)
    for world in (1, 2, 4, 8):
        own = sharding.owner_rank(t, world)
        assert own.min() >= 0 and own.max() < world
        if world > 1:
            assert len(np.unique(own)) == world
    parts = [sharding.shard_rows(t, r, 4) for r in range(4)]
    assert sum(len(p["value"]) for p in parts) == len(t["value"])


def test_two_ranks_gloo_cpu(tmp_path):
    _launch("cpu", 2, tmp_path, 29611)


@pytest.mark.gpu
def test_two_ranks_nccl_exchange(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests.util import assert_same_rows, oracle_rows
    _launch("gpu", 2, tmp_path, 29612)
    table = synth.make_flows(4000, 40, seed=77, dup_frac=0.05, ragged=True)
    t2 = synth.make_flows(2, 6000, seed=78)
    table = {k: np.concatenate([table[k], t2[k]]) for k in table}
    for algo in ("EWMA", "DBSCAN"):
        parts = [np.load(os.path.join(tmp_path, "res_%s_%d.npz" % (algo, r))) for r in range(2)]
        got = {k: np.concatenate([p[k] for p in parts]) for k in parts[0].files}
        want, _, _ = oracle_rows(table, algo, emit_all=True)
        assert_same_rows(got, want, what="2-rank " + algo)
