"""Host-side mirror of the reference interface: query planning (pinned to the reference's 12 SQL test cases),
label clean-up, argument parsing, the controller's validation messages (controller_test.go:318-446), and --
on the GPU -- the aggregated-flow modes and the job state machine."""
import json
import os
from collections import defaultdict

import numpy as np
import pytest

from oracle import tad_oracle as o
from theia_b200 import anomaly_detection as job
from theia_b200 import controller as ctl

G = os.path.join(os.path.dirname(__file__), "golden")
SQL_CASES = json.load(open(os.path.join(G, "sql_cases.json")))


@pytest.mark.parametrize("case", SQL_CASES, ids=[str(i) for i in range(len(SQL_CASES))])
def test_query_plan_renders_reference_sql(case):
    """anomaly_detection_test.py:46-195 (test_generate_sql_query)."""
    plan = job.plan_query(*case["args"])
    assert plan.to_sql() == case["sql"]
    assert plan.reducer == (0 if not case["args"][3] else 1)          # max() per connection, sum() when aggregated


def test_remove_meaningless_labels():
    s = '{"app":"x","pod-template-hash":"abc","controller-revision-hash":"1","pod-template-generation":"2","z":"1"}'
    assert job.remove_meaningless_labels(s) == '{"app": "x", "z": "1"}'
    assert job.remove_meaningless_labels("not json") == ""


def test_parse_args_accepts_controller_spelling():
    a = job.parse_args(["--algo", "EWMA", "--start_time", "2022-01-01 00:00:00", "--id", "abc",
                        "--ns-ignore-list", '["kube-system","flow-visibility"]', "--agg-flow", "pod", "--pod-label", "app"])
    assert a["algo"] == "EWMA" and a["ns_ignore_list"] == ["kube-system", "flow-visibility"] and a["agg_flow"] == "pod"
    with pytest.raises(SystemExit):
        job.parse_args(["--algo", "LSTM"])
    with pytest.raises(SystemExit):
        job.parse_args(["--start_time", "yesterday"])


UUID = "tad-5ca1ab1e-0000-4000-8000-000000000001"
INVALID = [
    (UUID, ctl.TADSpec(jobType="nonexistent-job-type"),
     "invalid request: Throughput Anomaly Detector algorithm type should be 'EWMA' or 'ARIMA' or 'DBSCAN'"),
    (UUID, ctl.TADSpec(jobType="ARIMA", startInterval="2023-01-01 00:00:10", endInterval="2023-01-01 00:00:00"),
     "invalid request: EndInterval should be after StartInterval"),
    (UUID, ctl.TADSpec(jobType="ARIMA", executorInstances=-1), "invalid request: ExecutorInstances should be an integer >= 0"),
    (UUID, ctl.TADSpec(jobType="ARIMA", driverCoreRequest="m200"),
     "invalid request: DriverCoreRequest should conform to the Kubernetes resource quantity convention"),
    (UUID, ctl.TADSpec(jobType="ARIMA", driverMemory="m512"),
     "invalid request: DriverMemory should conform to the Kubernetes resource quantity convention"),
    (UUID, ctl.TADSpec(jobType="ARIMA", executorCoreRequest="m200"),
     "invalid request: ExecutorCoreRequest should conform to the Kubernetes resource quantity convention"),
    (UUID, ctl.TADSpec(jobType="ARIMA", executorMemory="m512"),
     "invalid request: ExecutorMemory should conform to the Kubernetes resource quantity convention"),
    (UUID, ctl.TADSpec(jobType="ARIMA", aggFlow="pod", podNameSpace="podNameSpace"),
     "invalid request: 'pod-namespace' argument can not be used alone"),
    (UUID, ctl.TADSpec(jobType="ARIMA", aggFlow="nonexistent-agg-flow"),
     "invalid request: Throughput Anomaly Detector aggregated flow type should be 'pod' or 'external' or 'svc'"),
    ("tad-not-a-uuid", ctl.TADSpec(jobType="EWMA"), "invalid request: Throughput Anomaly Detector Querier job name is invalid"),
]


@pytest.mark.parametrize("name,spec,msg", INVALID, ids=[m[20:60] for _, _, m in INVALID])
def test_controller_invalid_requests(name, spec, msg):
    """controller_test.go:318-446: illegal arguments are terminal FAILED with these messages."""
    c = ctl.AnomalyDetectorController(engine=None)
    c.create(name, spec)
    st = c.sync(name)
    assert st.state == "FAILED"
    assert st.errorMsg.startswith("error in creating AnomalyDetector: ") and msg in st.errorMsg


def test_controller_job_args():
    a = ctl.build_job_args(UUID, ctl.TADSpec(jobType="DBSCAN", aggFlow="svc", servicePortName="web", nsIgnoreList=["kube-system"]))
    assert a == {"algo_type": "DBSCAN", "ns_ignore_list": ["kube-system"], "agg_flow": "svc", "svc_port_name": "web",
                 "tad_id": "5ca1ab1e-0000-4000-8000-000000000001"}


# ------------------------------------------------------------------------------------------------
# GPU: aggregated-flow modes against a direct dictionary-based evaluation of the same SQL semantics
# ------------------------------------------------------------------------------------------------
def _flows(seed=3, n_pods=12, n_rows=6000):
    rng = np.random.default_rng(seed)
    ns = np.array(["default", "kube-system", "shop", "flow-visibility"])
    labels = np.array(['{"app":"web","pod-template-hash":"%d"}' % i if i % 3 else "" for i in range(n_pods)])
    names = np.array(["pod-%d" % i for i in range(n_pods)])
    svc = np.array(["", "shop/web:http", "shop/db:pg", "default/dns:udp"])
    sp, dp = rng.integers(0, n_pods, n_rows), rng.integers(0, n_pods, n_rows)
    t_end = 1660200000 + 60 * rng.integers(0, 40, n_rows)
    base = 1e6 * (1 + sp + dp)
    thr = np.rint(base * (1 + 0.01 * rng.normal(size=n_rows)) * np.where(rng.random(n_rows) < 0.01, 30, 1)).astype(np.uint64)
    return {
        "sourceIP": (0x0A000000 + sp).astype(np.uint32), "destinationIP": (0x0A000100 + dp % 5).astype(np.uint32),
        "sourceTransportPort": rng.integers(1024, 1030, n_rows).astype(np.uint16),
        "destinationTransportPort": np.full(n_rows, 443, np.uint16), "protocolIdentifier": np.full(n_rows, 6, np.uint8),
        "flowStartSeconds": np.full(n_rows, 1660199000, np.uint32), "flowEndSeconds": t_end.astype(np.uint32),
        "throughput": thr, "flowType": rng.choice([1, 2, 3], n_rows).astype(np.uint8),
        "sourcePodNamespace": ns[sp % 4], "destinationPodNamespace": ns[dp % 4],
        "sourcePodLabels": labels[sp], "destinationPodLabels": labels[dp],
        "sourcePodName": names[sp], "destinationPodName": names[dp],
        "destinationServicePortName": svc[dp % 4],
    }


def _expected(flows, key_rows, algo="EWMA"):
    """key_rows: list of (key tuple, row index).  GROUP BY key, flowEndSeconds -> sum; series by time; EWMA flags."""
    agg = defaultdict(lambda: defaultdict(int))
    for key, i in key_rows:
        agg[key][int(flows["flowEndSeconds"][i])] += int(flows["throughput"][i])
    out = set()
    for key, pts in agg.items():
        ts = sorted(pts)
        vals = [pts[t] for t in ts]
        sd = o.stddev_samp(vals)
        calc = o.calculate_ewma(vals)
        for t, c, f in zip(ts, calc, o.calculate_ewma_anomaly(vals, sd)):
            if f:
                out.add((key, t, float(c)))
    return out


@pytest.mark.gpu
def test_agg_modes(engine):
    fl = _flows()
    n = len(fl["throughput"])
    # svc: destinationServicePortName <> '' ; key = service port name
    rows, st = job.anomaly_detection(engine, "EWMA", fl, tad_id="t1", agg_flow="svc")
    exp = _expected(fl, [((fl["destinationServicePortName"][i],), i) for i in range(n) if fl["destinationServicePortName"][i] != ""])
    got = {((r["destinationServicePortName"],), r["flowEndSeconds"], r["algoCalc"]) for r in rows if r["anomaly"] == "true"}
    assert got == exp and all(r["aggType"] == "svc" for r in rows)
    # external: flowType = 3 ; key = (destinationIP, flowType)
    rows, st = job.anomaly_detection(engine, "EWMA", fl, tad_id="t2", agg_flow="external")
    exp = _expected(fl, [((job.u32_to_ip(int(fl["destinationIP"][i])),), i) for i in range(n) if fl["flowType"][i] == 3])
    got = {((r["destinationIP"],), r["flowEndSeconds"], r["algoCalc"]) for r in rows if r["anomaly"] == "true"}
    assert got == exp
    # pod, by label: inbound UNION ALL outbound, labels cleaned in the output
    rows, st = job.anomaly_detection(engine, "EWMA", fl, tad_id="t3", agg_flow="pod", pod_label="web")
    kr = []
    for i in range(n):
        if "web" in fl["destinationPodLabels"][i].lower():
            kr.append(((fl["destinationPodNamespace"][i], fl["destinationPodLabels"][i], "inbound"), i))
        if "web" in fl["sourcePodLabels"][i].lower():
            kr.append(((fl["sourcePodNamespace"][i], fl["sourcePodLabels"][i], "outbound"), i))
    exp = {((k[0], job.remove_meaningless_labels(k[1]), k[2]), t, c) for k, t, c in _expected(fl, kr)}
    got = {((r["podNamespace"], r["podLabels"], r["direction"]), r["flowEndSeconds"], r["algoCalc"]) for r in rows if r["anomaly"] == "true"}
    assert got == exp and len(got) > 0
    # pod, by name + namespace
    rows, st = job.anomaly_detection(engine, "EWMA", fl, tad_id="t4", agg_flow="pod", pod_name="pod-3", pod_namespace="flow-visibility")
    kr = [((fl["destinationPodNamespace"][i], "pod-3", "inbound"), i) for i in range(n)
          if fl["destinationPodName"][i] == "pod-3" and fl["destinationPodNamespace"][i] == "flow-visibility"]
    kr += [((fl["sourcePodNamespace"][i], "pod-3", "outbound"), i) for i in range(n)
           if fl["sourcePodName"][i] == "pod-3" and fl["sourcePodNamespace"][i] == "flow-visibility"]
    exp = _expected(fl, kr)
    got = {((r["podNamespace"], r["podName"], r["direction"]), r["flowEndSeconds"], r["algoCalc"]) for r in rows if r["anomaly"] == "true"}
    assert got == exp


@pytest.mark.gpu
def test_agg_modes_with_time_window(engine):
    """external / svc with --start_time / --end_time: flowStartSeconds is not part of their key, yet the reference's
    WHERE clause filters on it (anomaly_detection.py:581-586)."""
    fl = _flows(seed=9)
    n = len(fl["throughput"])
    fl["flowStartSeconds"] = (1660199000 + 60 * np.random.default_rng(1).integers(0, 40, n)).astype(np.uint32)
    start, end = "2022-08-11 06:35:00", "2022-08-11 07:00:00"
    lo, hi = job._epoch(start), job._epoch(end)
    win = [i for i in range(n) if fl["flowStartSeconds"][i] >= lo and fl["flowEndSeconds"][i] < hi]
    assert 0 < len(win) < n and (fl["flowStartSeconds"] < lo).any()
    rows, st = job.anomaly_detection(engine, "EWMA", fl, start_time=start, end_time=end, tad_id="w1", agg_flow="svc")
    exp = _expected(fl, [((fl["destinationServicePortName"][i],), i) for i in win if fl["destinationServicePortName"][i] != ""])
    got = {((r["destinationServicePortName"],), r["flowEndSeconds"], r["algoCalc"]) for r in rows if r["anomaly"] == "true"}
    assert got == exp and len(exp) > 0
    rows, st = job.anomaly_detection(engine, "EWMA", fl, start_time=start, end_time=end, tad_id="w2", agg_flow="external")
    exp = _expected(fl, [((job.u32_to_ip(int(fl["destinationIP"][i])),), i) for i in win if fl["flowType"][i] == 3])
    got = {((r["destinationIP"],), r["flowEndSeconds"], r["algoCalc"]) for r in rows if r["anomaly"] == "true"}
    assert got == exp and len(exp) > 0
    # start only: the lower bound alone must not empty the job either
    rows, st = job.anomaly_detection(engine, "EWMA", fl, start_time=start, tad_id="w3", agg_flow="svc")
    win2 = [i for i in range(n) if fl["flowStartSeconds"][i] >= lo]
    exp = _expected(fl, [((fl["destinationServicePortName"][i],), i) for i in win2 if fl["destinationServicePortName"][i] != ""])
    got = {((r["destinationServicePortName"],), r["flowEndSeconds"], r["algoCalc"]) for r in rows if r["anomaly"] == "true"}
    assert got == exp and len(exp) > 0


@pytest.mark.gpu
def test_per_connection_with_namespace_ignore_and_window(engine):
    fl = _flows(seed=5)
    n = len(fl["throughput"])
    ignore = ["kube-system"]
    rows, st = job.anomaly_detection(engine, "EWMA", fl, start_time="2022-08-11 06:00:00", end_time="2022-08-11 07:00:00",
                                     tad_id="t5", ns_ignore_list=ignore)
    lo, hi = job._epoch("2022-08-11 06:00:00"), job._epoch("2022-08-11 07:00:00")
    kr = [((job.u32_to_ip(int(fl["sourceIP"][i])), int(fl["sourceTransportPort"][i]), job.u32_to_ip(int(fl["destinationIP"][i])),
            int(fl["destinationTransportPort"][i]), int(fl["protocolIdentifier"][i]), int(fl["flowStartSeconds"][i])), i)
          for i in range(n) if fl["sourcePodNamespace"][i] not in ignore and fl["destinationPodNamespace"][i] not in ignore
          and fl["flowStartSeconds"][i] >= lo and fl["flowEndSeconds"][i] < hi]
    # per-connection mode reduces duplicates with max(), not sum()
    agg = defaultdict(lambda: defaultdict(int))
    for key, i in kr:
        t = int(fl["flowEndSeconds"][i])
        agg[key][t] = max(agg[key][t], int(fl["throughput"][i]))
    exp = set()
    for key, pts in agg.items():
        ts = sorted(pts)
        vals = [pts[t] for t in ts]
        sd = o.stddev_samp(vals)
        for t, c, f in zip(ts, o.calculate_ewma(vals), o.calculate_ewma_anomaly(vals, sd)):
            if f:
                exp.add((key, t, float(c)))
    got = {((r["sourceIP"], r["sourceTransportPort"], r["destinationIP"], r["destinationTransportPort"],
             r["protocolIdentifier"], r["flowStartSeconds"]), r["flowEndSeconds"], r["algoCalc"]) for r in rows if r["anomaly"] == "true"}
    assert got == exp and st["rows_kept"] == len(kr)


@pytest.mark.gpu
def test_sentinel_row_when_nothing_is_anomalous(engine):
    fl = _flows(seed=6, n_rows=50)
    fl["throughput"][:] = 1000                                    # constant -> stddev 0 ... |x - ewma| > 0 flags the ramp-up
    rows, st = job.anomaly_detection(engine, "DBSCAN", fl, tad_id="t6", agg_flow="svc", svc_port_name="no-such-service")
    assert len(rows) == 1 and rows[0]["anomaly"] == "NO ANOMALY DETECTED" and rows[0]["aggType"] == "svc"      # :395-420
    assert rows[0]["sourceIP"] == "None" and rows[0]["algoType"] == "DBSCAN" and rows[0]["id"] == "t6"


@pytest.mark.gpu
def test_controller_state_machine(engine):
    """NEW -> SCHEDULED -> (RUNNING with stage progress) -> COMPLETED, result retrieval by id, deletion
    (controller.go:354-424).  Dispatch and observation are decoupled: sync() never blocks on the job."""
    c = ctl.AnomalyDetectorController(engine)
    name = "tad-5ca1ab1e-0000-4000-8000-00000000beef"
    c.create(name, ctl.TADSpec(jobType="EWMA", aggFlow="svc"))
    st = c.sync(name, flows=_flows(seed=7))
    assert st.state == "SCHEDULED" and st.sparkApplication == name[4:]
    seen = []
    st = c.wait(name, observe=lambda s: seen.append((s.state, s.completedStages, s.totalStages)))
    assert st.state == "COMPLETED" and st.completedStages == st.totalStages == 6
    assert all(a in ("SCHEDULED", "RUNNING", "COMPLETED") for a, _, _ in seen)
    prog = [k for a, k, _ in seen if a == "RUNNING"]
    assert prog == sorted(prog) and all(0 <= k <= 6 for k in prog)            # progress never goes backwards
    rows = c.results[name[4:]]
    assert rows and all(r["id"] == name[4:] for r in rows)
    c.delete(name)
    assert name not in c.crs and name[4:] not in c.results


@pytest.mark.gpu
def test_controller_observes_intermediate_progress_on_the_gpu(engine):
    """A job long enough to be caught in flight (ARIMA, the slowest detector): the controller reports RUNNING with
    0 < completedStages < totalStages before COMPLETED (controller_test.go:303-306 asserts an intermediate count too)."""
    from theia_b200 import synth
    fl = synth.make_flows(3000, 60, seed=12)
    flows = {"sourceIP": fl["src_ip"], "destinationIP": fl["dst_ip"], "sourceTransportPort": fl["src_port"],
             "destinationTransportPort": fl["dst_port"], "protocolIdentifier": fl["proto"], "flowStartSeconds": fl["flow_start"],
             "flowEndSeconds": fl["flow_end"], "throughput": (fl["value"] % 500 + 50).astype(np.uint64)}
    c = ctl.AnomalyDetectorController(engine)
    name = "tad-5ca1ab1e-0000-4000-8000-0000000000a1"
    c.create(name, ctl.TADSpec(jobType="ARIMA"))
    assert c.sync(name, flows=flows).state == "SCHEDULED"
    seen = []
    st = c.wait(name, observe=lambda s: seen.append((s.state, s.completedStages, s.totalStages)))
    assert st.state == "COMPLETED" and st.completedStages == st.totalStages == 6
    mid = [(k, n) for a, k, n in seen if a == "RUNNING" and 0 < k < n]
    assert mid, seen[-5:]
    c.delete(name)
