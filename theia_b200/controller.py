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


class AnomalyDetectorController:
    """syncTADetector (controller.go:354-383) over an in-memory CR store; one GPU engine instead of the Spark operator."""

    def __init__(self, engine):
        self.engine = engine
        self.crs: dict = {}          # name -> (spec, status)
        self._jobs: dict = {}        # name -> (Job, Columns, args)
        self.results: dict = {}      # id -> tadetector rows (stands in for the ClickHouse table)

    def create(self, name: str, spec: TADSpec):
        self.crs[name] = (spec, TADStatus(state=""))

    def sync(self, name: str, flows: dict | None = None) -> TADStatus:
        from . import anomaly_detection as job
        spec, st = self.crs[name]
        if st.state in ("", STATE_NEW):                               # startJob, controller.go:499-523
            try:
                args = build_job_args(name, spec)
            except IllegalArgumentError as e:
                st.state, st.errorMsg = STATE_FAILED, "error in creating AnomalyDetector: %s" % e
                return st
            st.state, st.sparkApplication = STATE_SCHEDULED, args["tad_id"]
            self._jobs[name] = (args, flows)
        elif st.state in (STATE_SCHEDULED, STATE_RUNNING):            # checkSparkApplicationStatus / updateProgress
            args, fl = self._jobs[name]
            rows, jst = job.anomaly_detection(self.engine, flows=fl, **args)
            st.completedStages, st.totalStages = jst["completed_stages"], jst["total_stages"]
            if jst["state"] == "COMPLETED":
                self.results[args["tad_id"]] = rows
                st.state = STATE_COMPLETED                            # finishJob, controller.go:400-424
            else:
                st.state, st.errorMsg = STATE_FAILED, jst["err_msg"]
        return st

    def delete(self, name: str):
        """cleanupTADetector (controller.go:385-398): drop the job and its rows (ALTER TABLE ... DELETE WHERE id)."""
        spec, st = self.crs.pop(name)
        self._jobs.pop(name, None)
        self.results.pop(st.sparkApplication, None)
