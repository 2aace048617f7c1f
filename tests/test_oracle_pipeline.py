"""Stages A-D: the C port of the oracle against the straightforward numpy/Python restatement,
including duplicates (stage-A max/sum), ragged series, filters and the e2e fixture of the reference
(test/e2e/throughputanomalydetection_test.go:401-492; expected rows in SURVEY.md appendix A)."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle, tad_oracle as o
from theia_b200 import synth

REF = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_test_vectors.json")))


def _same(pr, cc):
    cc = o.canonicalize(cc)
    for k in pr.cols:
        assert np.array_equal(pr.cols[k], cc[k], equal_nan=True), k


@pytest.mark.parametrize("algo", [o.ALGO_EWMA, o.ALGO_DBSCAN])
@pytest.mark.parametrize("emit_all", [False, True])
@pytest.mark.parametrize("kw", [
    dict(n_series=40, points_per_series=25, seed=1, dup_frac=0.2, ragged=True),
    dict(n_series=200, points_per_series=30, seed=2),
    dict(n_series=1, points_per_series=500, seed=3),
    dict(n_series=300, points_per_series=1, seed=4),
])
def test_c_port_pipeline(algo, emit_all, kw):
    t = synth.make_flows(**kw)
    for reducer in (o.REDUCE_MAX, o.REDUCE_SUM):
        pr = o.run_job(t, o.JobSpec(algo=algo, emit_all=emit_all, reducer=reducer))
        cc, ns, npts = c_oracle.run_job(t, algo=algo, emit_all=emit_all, reducer=reducer, threads=3)
        assert ns == pr.n_series and npts == pr.n_points
        _same(pr, cc)


def test_filters():
    t = synth.make_flows(100, 40, seed=5)
    start, end = synth.T0 + 600, synth.T0 + 3600
    pr = o.run_job(t, o.JobSpec(start_time=start, end_time=end, emit_all=True))
    cc, ns, npts = c_oracle.run_job(t, start_time=start, end_time=end, emit_all=True)
    assert 0 < npts < len(t["value"])
    _same(pr, cc)
    assert (pr.cols["flow_start"] >= start).all() and (pr.cols["flow_end"] < end).all()


def test_empty_table():
    t = {k: np.zeros(0, dtype=v) for k, v in synth.COLUMN_DTYPES.items()}
    pr = o.run_job(t, o.JobSpec())
    cc, ns, npts = c_oracle.run_job(t)
    assert len(pr) == 0 and len(cc["flow_end"]) == 0 and ns == 0
    rows = o.tadetector_rows(pr, "EWMA", "abc")
    assert len(rows) == 1 and rows[0]["anomaly"] == "NO ANOMALY DETECTED"     # anomaly_detection.py:395-420


def test_e2e_fixture_known_answers():
    # duplicates (the e2e test re-inserts the rows per sub-test) + shuffle; stage A's max collapses them
    t = synth.golden_e2e_table(REF["throughput_list"], duplicates=2)
    r = o.run_job(t, o.JobSpec(algo=o.ALGO_EWMA))
    assert r.n_series == 1 and r.n_points == 90
    assert list((r.cols["flow_end"] - (synth.T0 + 3600)) // 60) == [68, 69, 70]
    assert list(r.cols["algo_calc"]) == [27003756818.20375, 15504576757.601875, 9754862525.800938]
    assert list(r.cols["throughput"]) == [50007861276.0, 4005396697.0, 4005148294.0]
    assert abs(r.cols["stddev"][0] - 4919851535.682699) < 1e-5
    d = o.run_job(t, o.JobSpec(algo=o.ALGO_DBSCAN))
    assert list((d.cols["flow_end"] - (synth.T0 + 3600)) // 60) == [58, 60, 68, 80, 88]
    assert (d.cols["algo_calc"] == 0.0).all()
    rows = o.tadetector_rows(r, "EWMA", "1234")
    assert rows[0]["sourceIP"] == "10.10.1.25" and rows[0]["destinationTransportPort"] == 5201
    assert rows[0]["anomaly"] == "true" and rows[0]["aggType"] == "None"
