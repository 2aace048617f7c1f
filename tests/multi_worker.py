"""Worker of the multi-rank tests (launched with torch.distributed.run, one process per rank).

    python -m torch.distributed.run --nproc-per-node N tests/multi_worker.py <mode> <out_dir>

mode = cpu : gloo only; checks the host-side sharding logic (no GPU).
mode = gpu : every rank runs its row shard through the engine (NCCL exchange inside the library) and
             saves its result rows; the parent test compares the union with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from theia_b200 import sharding, synth  # noqa: E402


def main():
    mode, out_dir = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    table = synth.make_flows(4000, 40, seed=77, dup_frac=0.05, ragged=True)      # same table on every rank
    t2 = synth.make_flows(2, 6000, seed=78)                                      # two long connections -> spill path
    table = {k: np.concatenate([table[k], t2[k]]) for k in table}
    total = len(table["value"])
    mine = sharding.shard_rows(table, rank, world)
    own = sharding.owner_rank(table, world, total)
    expect_owned = int((own == rank).sum())
    counts = [None] * world
    dist.all_gather_object(counts, (len(mine["value"]), expect_owned))
    assert sum(c[0] for c in counts) == total and sum(c[1] for c in counts) == total
    # a connection has exactly one owner
    h = sharding.key_hash(table)
    order = np.argsort(h, kind="stable")
    same = h[order][1:] == h[order][:-1]
    assert (own[order][1:][same] == own[order][:-1][same]).all()
    if mode == "gpu":
        from theia_b200.engine import TadEngine
        os.environ["TAD_EXCHANGE_MIN_ROWS"] = "0"        # exercise the chunked (overlapped) exchange on this small table
        uid = [TadEngine.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        local = int(os.environ.get("LOCAL_RANK", rank))
        eng = TadEngine(device=local, world_size=world, rank=rank, nccl_unique_id=uid[0])
        for algo in ("EWMA", "DBSCAN"):
            got, st = eng.run(mine, algo=algo, tad_id="multi", emit_all=True)
            assert st["rows_owned"] == expect_owned, (rank, st["rows_owned"], expect_owned)
            assert st["rows_kept"] == len(mine["value"])
            np.savez(os.path.join(out_dir, "res_%s_%d.npz" % (algo, rank)), **got)
        eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
