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


@pytest.mark.parametrize("world", [2, 4])
def test_ranks_gloo_cpu(tmp_path, world):
    _launch("cpu", world, tmp_path, 29611 + world)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_on_gpus_match_the_oracle(tmp_path, world):
    """Union of the ranks' result rows == the oracle's rows on the whole table (anomaly_detection.py:680-684: the grouping must
    survive the sharding), for the peer-pull path, its fallback and the NCCL exchange; every result column bit-exact."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from tests.multi_worker import case_table
    from tests.util import assert_same_rows, oracle_rows
    _launch("gpu", world, tmp_path, 29620 + world)
    for name, algos in (("spread", ("EWMA", "DBSCAN")), ("skewed", ("EWMA", "DBSCAN")), ("xpull", ("EWMA", "DBSCAN")),
                        ("nccl", ("EWMA",)), ("ncclskew", ("EWMA",))):
        table = case_table({"nccl": "spread", "xpull": "skewed", "ncclskew": "skewed"}.get(name, name))
        for algo in algos:
            parts = [np.load(os.path.join(tmp_path, "res_%s_%s_%d.npz" % (name, algo, r))) for r in range(world)]
            got = {k: np.concatenate([p[k] for p in parts]) for k in parts[0].files}
            want, _, _ = oracle_rows(table, algo, emit_all=True)
            assert_same_rows(got, want, what="%d ranks, %s, %s" % (world, name, algo))
