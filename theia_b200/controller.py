"""A double of the Go controller's TAD state machine (pkg/controller/anomalydetector/controller.go) with the
GPU job runner behind the seam where the reference creates a SparkApplication -- what INTEGRATION.md's cgo shim
does, executable here because Python is the only host toolchain in the image.

Kept identical to the reference: the validation order and every "invalid request: ..." message of
startSparkApplication (controller.go:527-622, asserted by controller_test.go:318-446), the states
(types.go:33-37), "error in creating AnomalyDetector: ..." for illegal arguments (controller.go:505-514, terminal
FAILED, not retried), CompletedStages/TotalStages progress (controller.go:426-453), the id = name minus "tad-"
(controller.go:622-623)."""
from __future__ import annotations

import re
import threading
import uuid
from dataclasses import dataclass, field

K8S_QUANTITIES_REG = re.compile(r"^([+-]?[0-9.]+)([eEinumkKMGTP]*[-+]?[0-9]*)$")      # pkg/controller/util.go:46

STATE_NEW, STATE_SCHEDULED, STATE_RUNNING, STATE_COMPLETED, STATE_FAILED = "NEW", "SCHEDULED", "RUNNING", "COMPLETED", "FAILED"


class IllegalArgumentError(Exception):
    """illeagelArguementError of the reference (controller.go:63-65)."""


@dataclass
class TADSpec:
    """ThroughputAnomalyDetectorSpec, pkg/apis/crd/v1alpha1/types.go:97-113 (times as 'YYYY-MM-DD hh:mm:ss' or '')."""
    jobType: str = ""
    startInterval: str = ""
    endInterval: str = ""
    nsIgnoreList: list = field(default_factory=list)
    aggFlow: str = ""
    podLabel: str = ""
    podName: str = ""
    podNameSpace: str = ""
    externalIp: str = ""
    servicePortName: str = ""
    executorInstances: int = 1
    driverCoreRequest: str = "200m"
    driverMemory: str = "512M"
    executorCoreRequest: str = "200m"
    executorMemory: str = "512M"


@dataclass
class TADStatus:
    """ThroughputAnomalyDetectorStatus, types.go:114-122."""
    state: str = ""
    sparkApplication: str = ""           # kept for API compatibility: carries the job id
    completedStages: int = 0
    totalStages: int = 0
    errorMsg: str = ""


def parse_ad_algorithm_id(name: str) -> str:
    """pkg/util/utils.go ParseADAlgorithmID."""
    if not name.startswith("tad-"):
        raise ValueError("input name %s is not a valid Throughput Anomaly Detection job name" % name)
    try:
        uuid.UUID(name[4:])
    except ValueError as e:
        raise ValueError("input name %s does not contain a valid UUID, parsing error: %s" % (name, e))
    return name[4:]


def build_job_args(name: str, spec: TADSpec) -> dict:
    """startSparkApplication's validation + argv construction (controller.go:527-623), returning the keyword
    arguments of theia_b200.anomaly_detection.anomaly_detection instead of a SparkApplication."""
    if spec.jobType not in ("EWMA", "ARIMA", "DBSCAN"):
        raise IllegalArgumentError("invalid request: Throughput Anomaly Detector algorithm type should be 'EWMA' or 'ARIMA' or 'DBSCAN'")
    args = {"algo_type": spec.jobType}
    if spec.startInterval:
        args["start_time"] = spec.startInterval
    if spec.endInterval:
        if not (spec.endInterval > spec.startInterval):          # same format on both sides: lexical == chronological
            raise IllegalArgumentError("invalid request: EndInterval should be after StartInterval")
        args["end_time"] = spec.endInterval
    if spec.nsIgnoreList:
        args["ns_ignore_list"] = list(spec.nsIgnoreList)
    if spec.aggFlow:
        if spec.aggFlow == "pod":
            args["agg_flow"] = "pod"
            if spec.podLabel:
                args["pod_label"] = spec.podLabel
            if spec.podName:
                args["pod_name"] = spec.podName
            if spec.podNameSpace:
                if not spec.podName and not spec.podLabel:
                    raise IllegalArgumentError("invalid request: 'pod-namespace' argument can not be used alone, should be specified along pod-label or pod-name")
                args["pod_namespace"] = spec.podNameSpace
        elif spec.aggFlow == "external":
            args["agg_flow"] = "external"
            if spec.externalIp:
                args["external_ip"] = spec.externalIp
        elif spec.aggFlow == "svc":
            args["agg_flow"] = "svc"
            if spec.servicePortName:
                args["svc_port_name"] = spec.servicePortName
        else:
            raise IllegalArgumentError("invalid request: Throughput Anomaly Detector aggregated flow type should be 'pod' or 'external' or 'svc'")
    # the Spark resource knobs are meaningless for the GPU engine but are validated identically (controller.go:581-616)
    if spec.executorInstances < 0:
        raise IllegalArgumentError("invalid request: ExecutorInstances should be an integer >= 0")
    for attr, label in (("driverCoreRequest", "DriverCoreRequest"), ("driverMemory", "DriverMemory"),
                        ("executorCoreRequest", "ExecutorCoreRequest"), ("executorMemory", "ExecutorMemory")):
        if not K8S_QUANTITIES_REG.match(getattr(spec, attr)):
            raise IllegalArgumentError("invalid request: %s should conform to the Kubernetes resource quantity convention" % label)
    try:
        args["tad_id"] = parse_ad_algorithm_id(name)
    except ValueError as e:
        raise IllegalArgumentError("invalid request: Throughput Anomaly Detector Querier job name is invalid: %s" % e)
    return args


class _Run:
    """One dispatched job: the thread that drives it (what the Spark operator's driver pod is to the reference) and the
    engine job handle the controller polls for progress."""

    def __init__(self, args, flows):
        self.args, self.flows = args, flows
        self.lock = threading.Lock()
        self.handle = None            # engine Job while it is alive (set / cleared by the engine's on_job hook)
        self.rows = self.status = self.error = None
        self.done = threading.Event()
        self.thread = None


class AnomalyDetectorController:
    """syncTADetector (controller.go:354-383) over an in-memory CR store; one GPU engine instead of the Spark operator.

    Like the reference, dispatch and observation are decoupled: the first sync validates and dispatches the job
    (startJob, controller.go:499-523 -- here a thread that runs the job on the engine instead of a SparkApplication CR),
    every later sync only OBSERVES it: while the job runs it reports RUNNING with the engine's completed / total stages
    (updateProgress, controller.go:426-453, which scrapes the Spark UI; here tad_poll), and once it has ended it finishes
    the CR (finishJob, controller.go:400-424).  One deviation, on purpose: on completion the stage counters are set to the
    job's final n/n, where the reference keeps whatever its last RUNNING poll saw (controller_test.go:303-306: 3 of 5)."""

    def __init__(self, engine):
        self.engine = engine
        self.crs: dict = {}          # name -> (spec, status)
        self._jobs: dict = {}        # name -> _Run
        self.results: dict = {}      # id -> tadetector rows (stands in for the ClickHouse table)

    def create(self, name: str, spec: TADSpec):
        self.crs[name] = (spec, TADStatus(state=""))

    def _drive(self, run: _Run):
        from . import anomaly_detection as job

        def on_job(handle):          # called by the engine right after tad_submit, and with None before tad_release
            with run.lock:
                run.handle = handle
        try:
            run.rows, run.status = job.anomaly_detection(self.engine, flows=run.flows, on_job=on_job, **run.args)
        except Exception as e:       # the job's failure is the CR's failure (controller.go:455-497)
            run.error = str(e)
        finally:
            with run.lock:
                run.handle = None
            run.done.set()

    def sync(self, name: str, flows: dict | None = None) -> TADStatus:
        spec, st = self.crs[name]
        if st.state in ("", STATE_NEW):                               # startJob, controller.go:499-523
            try:
                args = build_job_args(name, spec)
            except IllegalArgumentError as e:
                st.state, st.errorMsg = STATE_FAILED, "error in creating AnomalyDetector: %s" % e
                return st
            st.state, st.sparkApplication = STATE_SCHEDULED, args["tad_id"]
            run = self._jobs[name] = _Run(args, flows)
            run.thread = threading.Thread(target=self._drive, args=(run,), daemon=True)
            run.thread.start()
        elif st.state in (STATE_SCHEDULED, STATE_RUNNING):            # checkSparkApplicationStatus / updateProgress
            run = self._jobs[name]
            if not run.done.is_set():
                with run.lock:
                    p = run.handle.poll() if run.handle is not None else None
                if p is not None and p["state"] in ("RUNNING", "COMPLETED"):
                    st.state = STATE_RUNNING
                    st.completedStages, st.totalStages = p["completed_stages"], p["total_stages"]
                return st
            if run.error is not None or run.status["state"] != "COMPLETED":
                st.state, st.errorMsg = STATE_FAILED, run.error if run.error is not None else run.status["err_msg"]
            else:
                st.completedStages, st.totalStages = run.status["completed_stages"], run.status["total_stages"]
                self.results[run.args["tad_id"]] = run.rows
                st.state = STATE_COMPLETED                            # finishJob, controller.go:400-424
        return st

    def wait(self, name: str, timeout: float = 600.0, observe=None) -> TADStatus:
        """Poll sync() until the CR is terminal (the informer's resync loop); ``observe`` sees every status."""
        import time
        t0 = time.time()
        while True:
            st = self.sync(name)
            if observe is not None:
                observe(st)
            if st.state in (STATE_COMPLETED, STATE_FAILED) or time.time() - t0 > timeout:
                return st
            self._jobs[name].done.wait(0.0005)

    def delete(self, name: str):
        """cleanupTADetector (controller.go:385-398): cancel a job that is still running (DeleteSparkApplication), then drop
        it and its rows (ALTER TABLE ... DELETE WHERE id)."""
        spec, st = self.crs.pop(name)
        run = self._jobs.pop(name, None)
        if run is not None and not run.done.is_set():
            with run.lock:
                if run.handle is not None:
                    run.handle.cancel()
            run.done.wait(60)
        self.results.pop(st.sparkApplication, None)
