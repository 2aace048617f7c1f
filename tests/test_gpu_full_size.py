"""BASELINE configs[1] at full size (100M records / 1M connections x 100 points) on one GPU: size-independent
properties + an exact oracle cross-check on a sample of connections (the whole table would take the oracle minutes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(engine, cols_t, rows, **kw):
    from theia_b200.engine import DeviceColumns
    dcols = DeviceColumns(rows, {k: v.data_ptr() for k, v in cols_t.items()})
    dcols.keepalive = cols_t
    job = engine.submit(dcols, tad_id="full", **kw)
    st = job.wait()
    res = job.result()
    job.release()
    return res, st


def test_full_size_properties(engine):
    import torch
    from theia_b200 import synth
    from tests.util import assert_same_rows, oracle_rows
    free, total = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~40 GB of free HBM")
    S, n = 1_000_000, 100
    dev = torch.device("cuda", 0)
    cols_t = synth.make_flows_torch(S, n, seed=7, device=dev)
    rows = S * n
    torch.cuda.synchronize()
    res, st = _run(engine, cols_t, rows, algo="EWMA")
    # counts: the generator draws S distinct keys (a collision would merge two connections)
    assert st["rows_in"] == rows and st["rows_kept"] == rows and st["points"] == rows
    assert abs(int(st["series"]) - S) <= 2
    assert st["result_rows"] == len(res["flow_end"]) > 0
    # every emitted row re-satisfies the flag rule of calculate_ewma_anomaly on its own columns
    assert (res["anomaly"] == 1).all()
    assert (np.abs(res["throughput"] - res["algo_calc"]) > res["stddev"]).all()
    assert (res["flow_end"] > res["flow_start"]).all() and ((res["flow_end"] - res["flow_start"]) % 60 == 0).all()
    # e_1 = x_1 / 2: the first point of a connection is flagged iff x_1 / 2 > stddev
    first = (res["flow_end"] - res["flow_start"]) == 60
    assert np.array_equal(res["algo_calc"][first], res["throughput"][first] * 0.5)
    # idempotence: the same job again gives the same set of rows
    res2, st2 = _run(engine, cols_t, rows, algo="EWMA")
    assert st2["result_rows"] == st["result_rows"]
    assert_same_rows(res2, res, what="idempotence")
    # exact oracle cross-check on the connections of 3000 sampled source ports x ips
    src_ip = cols_t["src_ip"]
    pick = torch.unique(src_ip)[::400][:3000]
    mask = torch.isin(src_ip, pick)
    sub = {k: v[mask].cpu().numpy() for k, v in cols_t.items()}
    sub = {"src_ip": sub["src_ip"].view(np.uint32), "dst_ip": sub["dst_ip"].view(np.uint32),
           "src_port": sub["src_port"].view(np.uint16), "dst_port": sub["dst_port"].view(np.uint16),
           "proto": sub["proto"], "flow_start": sub["flow_start"].view(np.uint32),
           "flow_end": sub["flow_end"].view(np.uint32), "value": sub["value"].view(np.uint64)}
    want, ns, npts = oracle_rows(sub, "EWMA")
    sel = np.isin(res["src_ip"], pick.cpu().numpy().view(np.uint32))
    got = {k: v[sel] for k, v in res.items()}
    assert ns >= 2500
    assert_same_rows(got, want, what="sampled connections at full size")
    # DBSCAN at the same size: counts + flag sanity (algoCalc placeholder is 0)
    res3, st3 = _run(engine, cols_t, rows, algo="DBSCAN")
    assert st3["points"] == rows and (res3["algo_calc"] == 0.0).all() and st3["result_rows"] > 0
    want3, _, _ = oracle_rows(sub, "DBSCAN")
    sel3 = np.isin(res3["src_ip"], pick.cpu().numpy().view(np.uint32))
    assert_same_rows({k: v[sel3] for k, v in res3.items()}, want3, what="sampled connections at full size, DBSCAN")
