"""The host mirror's job-level tests (tests/test_host_mirror.py, GPU-marked there) run here against an engine double that
answers with the CPU oracle: the string predicates, dictionaries, key packing, UNION ALL branches and result decoding of
theia_b200/anomaly_detection.py are exercised end to end without a GPU.  (The oracle is the checker on both sides; what
this adds is coverage of the host code in the CPU suite.)"""
import numpy as np

from oracle import tad_oracle as o
from theia_b200 import _lib as L

from . import test_host_mirror as thm


class OracleEngine:
    """Same ``run`` contract as theia_b200.engine.TadEngine, computed by oracle/tad_oracle.py."""

    def run(self, table, algo="EWMA", reducer=0, start_time=0, end_time=0, tad_id="", ns_ignore=(), flags=0):
        t = {k: v for k, v in table.items() if v is not None}
        spec = o.JobSpec(algo={"EWMA": o.ALGO_EWMA, "DBSCAN": o.ALGO_DBSCAN}[algo], reducer=reducer, start_time=start_time,
                         end_time=end_time, ns_ignore=tuple(ns_ignore))
        n = len(t["flow_end"])
        keep = np.ones(n, dtype=bool)
        if start_time:
            keep &= np.asarray(t.get("flow_start", np.zeros(n))) >= start_time
        if end_time:
            keep &= np.asarray(t["flow_end"]) < end_time
        if ns_ignore:
            keep &= ~np.isin(t["src_ns"], list(ns_ignore)) & ~np.isin(t["dst_ns"], list(ns_ignore))
        res = o.run_job(t, spec)
        return dict(res.cols), {"rows_kept": int(keep.sum()), "state": "COMPLETED", "completed_stages": 6, "total_stages": 6,
                                "err_msg": ""}


def test_agg_modes_on_the_oracle_engine():
    thm.test_agg_modes(OracleEngine())


def test_agg_modes_with_time_window_on_the_oracle_engine():
    thm.test_agg_modes_with_time_window(OracleEngine())


def test_per_connection_with_namespace_ignore_and_window_on_the_oracle_engine():
    thm.test_per_connection_with_namespace_ignore_and_window(OracleEngine())


def test_sentinel_row_on_the_oracle_engine():
    thm.test_sentinel_row_when_nothing_is_anomalous(OracleEngine())


def test_controller_state_machine_on_the_oracle_engine():
    thm.test_controller_state_machine(OracleEngine())


def test_reducer_codes_match():
    assert (o.REDUCE_MAX, o.REDUCE_SUM) == (L.TAD_REDUCE_MAX, L.TAD_REDUCE_SUM)
