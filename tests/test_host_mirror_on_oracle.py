"""The host mirror's job-level tests (tests/test_host_mirror.py, GPU-marked there) run here against an engine double that
answers with the CPU oracle: the string predicates, dictionaries, key packing, UNION ALL branches and result decoding of
theia_b200/anomaly_detection.py are exercised end to end without a GPU.  (The oracle is the checker on both sides; what
this adds is coverage of the host code in the CPU suite.)"""
import numpy as np

from oracle import tad_oracle as o
from theia_b200 import _lib as L

from . import test_host_mirror as thm


class _FakeJob:
    """tad_poll's view of a job, scripted."""

    def __init__(self):
        self.completed, self.cancelled = 0, False

    def stage(self, k):
        self.completed = k

    def poll(self):
        return {"state": "RUNNING", "completed_stages": self.completed, "total_stages": 6, "err_msg": ""}

    def cancel(self):
        self.cancelled = True


class OracleEngine:
    """Same ``run`` contract as theia_b200.engine.TadEngine, computed by oracle/tad_oracle.py."""

    def run(self, table, algo="EWMA", reducer=0, start_time=0, end_time=0, tad_id="", ns_ignore=(), flags=0, on_job=None):
        if on_job is not None:            # the controller's progress hook: a job handle that reports the stages (a double too)
            on_job(self.job)
            self.job.stage(3)
            self.gate.wait(30)            # "the job is running": held until the test has looked at the CR
        try:
            return self._run(table, algo, reducer, start_time, end_time, ns_ignore)
        finally:
            if on_job is not None:
                on_job(None)

    def __init__(self):
        import threading
        self.gate = threading.Event()
        self.gate.set()
        self.job = _FakeJob()

    def _run(self, table, algo, reducer, start_time, end_time, ns_ignore):
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


def test_controller_reports_intermediate_progress():
    """controller_test.go:285-306: while the application runs, the CR shows RUNNING with the stages completed so far (the
    reference's fake Spark UI says 3 of 5; the engine has 6 stages and the scripted job stands at 3)."""
    from theia_b200 import controller as ctl
    eng = OracleEngine()
    eng.gate.clear()                                   # hold the job "in flight"
    c = ctl.AnomalyDetectorController(eng)
    name = "tad-5ca1ab1e-0000-4000-8000-00000000cafe"
    c.create(name, ctl.TADSpec(jobType="EWMA", aggFlow="svc"))
    assert c.sync(name, flows=thm._flows(seed=7)).state == "SCHEDULED"
    import time
    for _ in range(2000):
        st = c.sync(name)
        if st.state == "RUNNING" and st.completedStages == 3:
            break
        time.sleep(0.001)
    assert (st.state, st.completedStages, st.totalStages) == ("RUNNING", 3, 6)
    assert name[4:] not in c.results                   # nothing is published before the job has ended
    eng.gate.set()
    st = c.wait(name)
    assert (st.state, st.completedStages, st.totalStages) == ("COMPLETED", 6, 6) and c.results[name[4:]]
    # deleting a CR whose job is still in flight cancels the job (DeleteSparkApplication, controller.go:385-398)
    eng.gate.clear()
    name2 = "tad-5ca1ab1e-0000-4000-8000-00000000f00d"
    c.create(name2, ctl.TADSpec(jobType="EWMA"))
    c.sync(name2, flows=thm._flows(seed=8))
    for _ in range(2000):
        if c.sync(name2).state == "RUNNING":
            break
        time.sleep(0.001)
    import threading
    threading.Timer(0.05, eng.gate.set).start()
    c.delete(name2)
    assert eng.job.cancelled and name2 not in c.crs


def test_reducer_codes_match():
    assert (o.REDUCE_MAX, o.REDUCE_SUM) == (L.TAD_REDUCE_MAX, L.TAD_REDUCE_SUM)
