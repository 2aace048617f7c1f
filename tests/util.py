"""Shared helpers of the parity tests: run the oracle and the CUDA engine on the same table and
compare every result column bit-for-bit after canonical ordering."""
import numpy as np

from oracle import c_oracle, tad_oracle

ALGO_ID = {"EWMA": 0, "ARIMA": 1, "DBSCAN": 2}
EXACT_COLS = ("src_ip", "src_port", "dst_ip", "dst_port", "proto", "flow_start", "flow_end", "throughput", "anomaly")
SCORE_COLS = ("stddev", "algo_calc")


def oracle_rows(table, algo="EWMA", **kw):
    cols, n_series, n_points = c_oracle.run_job(table, algo=ALGO_ID[algo], **kw)
    return tad_oracle.canonicalize(cols), n_series, n_points


def assert_same_rows(got: dict, want: dict, score_rtol: float = 0.0, what: str = ""):
    """id/key/timestamp/throughput/flag columns must be identical; score columns identical when
    score_rtol == 0 (EWMA, DBSCAN, stddev: same FP64 operation order as the oracle), else within
    score_rtol relative (north_star: 1e-6)."""
    got = tad_oracle.canonicalize(got)
    want = tad_oracle.canonicalize(want)
    assert len(got["flow_end"]) == len(want["flow_end"]), "%s: row count %d != %d" % (
        what, len(got["flow_end"]), len(want["flow_end"]))
    for c in EXACT_COLS:
        assert np.array_equal(got[c], want[c]), "%s: column %s differs" % (what, c)
    for c in SCORE_COLS:
        if score_rtol == 0.0:
            assert np.array_equal(got[c], want[c], equal_nan=True), "%s: score column %s differs (%d rows)" % (
                what, c, int(np.sum(~((got[c] == want[c]) | (np.isnan(got[c]) & np.isnan(want[c]))))))
        else:
            assert np.array_equal(np.isnan(got[c]), np.isnan(want[c])), "%s: NULLs of %s differ" % (what, c)
            np.testing.assert_allclose(got[c], want[c], rtol=score_rtol, atol=0, equal_nan=True,
                                       err_msg="%s: %s" % (what, c))
