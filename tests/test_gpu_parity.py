"""Parity of the CUDA path (through the C ABI) with the oracle: bit-exact on every column."""
import json
import os

import numpy as np
import pytest

from theia_b200 import synth
from tests.util import assert_same_rows, oracle_rows

pytestmark = pytest.mark.gpu
REF = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_test_vectors.json")))


def run_both(engine, table, algo="EWMA", **kw):
    got, st = engine.run(table, algo=algo, tad_id="test", **kw)
    okw = {k: v for k, v in kw.items() if k in ("emit_all", "reducer", "start_time", "end_time")}
    want, ns, npts = oracle_rows(table, algo=algo, **okw)
    assert st["state"] == "COMPLETED"
    assert st["series"] == ns and st["points"] == npts, (st, ns, npts)
    assert st["result_rows"] == len(want["flow_end"])
    assert st["gpu_launches"] > 0
    assert_same_rows(got, want, what="%s %s" % (algo, kw))
    return got, st


def test_e2e_fixture_ewma(engine):
    """The connection the reference e2e test inserts; known answers from the reference's own goldens."""
    t = synth.golden_e2e_table(REF["throughput_list"], duplicates=2)
    got, st = run_both(engine, t, "EWMA")
    assert list((np.sort(got["flow_end"]) - (synth.T0 + 3600)) // 60) == [68, 69, 70]
    order = np.argsort(got["flow_end"])
    assert list(got["algo_calc"][order]) == [27003756818.20375, 15504576757.601875, 9754862525.800938]
    # every EWMA value of the golden series (anomaly_detection_test.py:219-249), via emit_all
    allp, _ = run_both(engine, synth.golden_e2e_table(REF["throughput_list"]), "EWMA", emit_all=True)
    order = np.argsort(allp["flow_end"])
    assert list(allp["algo_calc"][order]) == REF["expected_ewma_row_list"]
    assert list(allp["anomaly"][order].astype(bool)) == REF["expected_anomaly_list_ewma"]


@pytest.mark.parametrize("kw", [
    dict(n_series=100, points_per_series=100, seed=1),                       # BASELINE configs[0]: 10K / 100
    dict(n_series=40, points_per_series=25, seed=2, dup_frac=0.2, ragged=True),
    dict(n_series=3000, points_per_series=30, seed=3, ragged=True),
    dict(n_series=5000, points_per_series=1, seed=4),                        # every stddev NULL
    dict(n_series=20000, points_per_series=100, seed=5),                     # 2M rows, many buckets
    dict(n_series=7, points_per_series=13, seed=6, shuffle=False),
])
@pytest.mark.parametrize("emit_all", [False, True])
def test_synthetic_ewma(engine, kw, emit_all):
    run_both(engine, synth.make_flows(**kw), "EWMA", emit_all=emit_all)


def test_reducer_sum(engine):
    t = synth.make_flows(300, 20, seed=7, dup_frac=0.5)
    run_both(engine, t, "EWMA", reducer=1, emit_all=True)
    run_both(engine, t, "EWMA", reducer=0, emit_all=True)


def test_time_filters(engine):
    t = synth.make_flows(500, 40, seed=8)
    got, st = run_both(engine, t, "EWMA", start_time=synth.T0 + 600, end_time=synth.T0 + 3600, emit_all=True)
    assert 0 < st["rows_kept"] < st["rows_in"]


def test_empty_and_tiny(engine):
    empty = {k: np.zeros(0, dtype=v) for k, v in synth.COLUMN_DTYPES.items()}
    got, st = engine.run(empty, algo="EWMA")
    assert st["state"] == "COMPLETED" and st["result_rows"] == 0 and len(got["flow_end"]) == 0
    run_both(engine, synth.make_flows(1, 1, seed=9), "EWMA", emit_all=True)
    run_both(engine, synth.make_flows(1, 2, seed=10), "EWMA", emit_all=True)
    run_both(engine, synth.make_flows(3, 5, seed=11), "EWMA", emit_all=True)     # odd row count: scalar tail


def test_missing_key_columns(engine):
    """Aggregated-flow style key: only two key slots populated (NULL columns = 0)."""
    t = synth.make_flows(200, 30, seed=12)
    t2 = dict(t)
    for c in ("src_ip", "src_port", "proto", "flow_start"):
        t2[c] = None
    run_both(engine, t2, "EWMA", reducer=1, emit_all=True)


def test_invalid_requests(engine):
    from theia_b200.engine import TadError
    t = synth.make_flows(2, 2, seed=13)
    cols = engine.columns_from_numpy(t)
    with pytest.raises(TadError) as e:
        engine.submit(cols, algo=7)
    assert "algorithm type should be 'EWMA' or 'ARIMA' or 'DBSCAN'" in str(e.value)     # controller.go:527-529
    with pytest.raises(TadError) as e:
        engine.submit(cols, algo="EWMA", start_time=100, end_time=50)
    assert "EndInterval should be after StartInterval" in str(e.value)                   # controller.go:535-539
    cols.free()


def test_job_state_machine(engine):
    t = synth.make_flows(2000, 50, seed=14)
    cols = engine.columns_from_numpy(t)
    job = engine.submit(cols, algo="EWMA", tad_id="5ca1ab1e-0000-4000-8000-000000000001")
    st = job.poll()
    assert st["state"] in ("SCHEDULED", "RUNNING", "COMPLETED") and st["total_stages"] == 6
    st = job.wait()
    assert st["state"] == "COMPLETED" and st["completed_stages"] == st["total_stages"]
    job.release()
    cols.free()


# ------------------------------------------------------------------------------------------------
# DBSCAN (anomaly_detection.py:325-349)
# ------------------------------------------------------------------------------------------------
def test_e2e_fixture_dbscan(engine):
    t = synth.golden_e2e_table(REF["throughput_list"], duplicates=1)
    got, st = run_both(engine, t, "DBSCAN")
    assert list((np.sort(got["flow_end"]) - (synth.T0 + 3600)) // 60) == [58, 60, 68, 80, 88]
    assert (got["algo_calc"] == 0.0).all()
    allp, _ = run_both(engine, synth.golden_e2e_table(REF["throughput_list"]), "DBSCAN", emit_all=True)
    order = np.argsort(allp["flow_end"])
    assert list(allp["anomaly"][order].astype(bool)) == REF["expected_dbscan_anomaly_list"]


@pytest.mark.parametrize("kw", [
    dict(n_series=100, points_per_series=100, seed=21),
    dict(n_series=4000, points_per_series=24, seed=22),                      # configs[3] shape: series length 24
    dict(n_series=3000, points_per_series=8, seed=23, ragged=True),          # n <= 11: sklearn brute-force regime
    dict(n_series=500, points_per_series=40, seed=24, dup_frac=0.3, ragged=True),
    dict(n_series=20000, points_per_series=100, seed=25),
])
@pytest.mark.parametrize("emit_all", [False, True])
def test_synthetic_dbscan(engine, kw, emit_all):
    run_both(engine, synth.make_flows(**kw), "DBSCAN", emit_all=emit_all)


def test_dbscan_golden_udf_cases(engine):
    """Every seeded series of tests/golden/udf_cases.json (outputs of the reference UDF itself, incl. exact-eps
    ties and > 2^53 values) as one table: one connection per case."""
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "udf_cases.json")))["cases"]
    cols = {k: [] for k in synth.COLUMN_DTYPES}
    for ci, c in enumerate(cases):
        n = len(c["values"])
        cols["src_ip"].append(np.full(n, 0x0A000000 + ci, np.uint32)); cols["dst_ip"].append(np.full(n, 0x0A010000, np.uint32))
        cols["src_port"].append(np.full(n, 1000 + ci, np.uint16)); cols["dst_port"].append(np.full(n, 443, np.uint16))
        cols["proto"].append(np.full(n, 6, np.uint8)); cols["flow_start"].append(np.full(n, synth.T0, np.uint32))
        cols["flow_end"].append((synth.T0 + 60 * (1 + np.arange(n))).astype(np.uint32))
        cols["value"].append(np.array(c["values"], dtype=np.uint64))
    t = {k: np.concatenate(v) for k, v in cols.items()}
    perm = np.random.default_rng(5).permutation(len(t["value"]))
    t = {k: v[perm] for k, v in t.items()}
    for algo, key in (("DBSCAN", "dbscan_flags"), ("EWMA", "ewma")):
        got, st = run_both(engine, t, algo, emit_all=True)
        for ci, c in enumerate(cases):
            m = got["src_ip"] == 0x0A000000 + ci
            order = np.argsort(got["flow_end"][m])
            if algo == "DBSCAN":
                assert list(got["anomaly"][m][order].astype(bool)) == c["dbscan_flags"], c["tag"]
            else:
                assert list(got["algo_calc"][m][order]) == c["ewma"], c["tag"]


# ------------------------------------------------------------------------------------------------
# oversized buckets -> spill path
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ["EWMA", "DBSCAN"])
def test_long_series_spill(engine, algo):
    """Connections longer than the shared-memory bucket capacity (2048 rows) take the global path."""
    t1 = synth.make_flows(3, 5000, seed=31, dup_frac=0.05)
    t2 = synth.make_flows(3000, 20, seed=32)
    t = {k: np.concatenate([t1[k], t2[k]]) for k in t1}
    got, st = run_both(engine, t, algo, emit_all=True)
    assert st["spill_rows"] >= 15000


@pytest.fixture(scope="module")
def tiny_bucket_engine():
    """Engine forced to 2 hash buckets: every table of more than a few thousand rows spills."""
    from theia_b200.engine import TadEngine
    os.environ["TAD_DEBUG_LOGB"] = "1"
    try:
        eng = TadEngine(device=0)
    finally:
        del os.environ["TAD_DEBUG_LOGB"]
    yield eng
    eng.close()


@pytest.mark.parametrize("algo", ["EWMA", "DBSCAN"])
def test_forced_spill_everything(tiny_bucket_engine, algo):
    t = synth.make_flows(800, 30, seed=33, dup_frac=0.1, ragged=True)
    got, st = run_both(tiny_bucket_engine, t, algo, emit_all=True)
    assert st["spill_rows"] == st["rows_kept"]
    got, st = run_both(tiny_bucket_engine, t, algo, reducer=1)
