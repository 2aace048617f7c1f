#!/usr/bin/env python3
"""bench.py -- flow records/s through EWMA throughput-anomaly detection (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one complete TAD job (stage A filter/reduce, group, time sort, stddev_samp, EWMA,
anomaly compaction) over one synthetic flow table of BASELINE.json configs[1]: 100M records /
1M connections x 100 points per GPU (weak scaling: every rank holds that many rows of a table
whose connections are spread over all ranks).

ours:       `value`  = rows / device time per step, inputs resident in HBM (CUDA events recorded by
                       the library on its own stream: first kernel start -> last kernel end).
            `e2e`    = same job through the C ABI with HOST (pinned) column buffers: H2D of the
                       29 B/row inputs and D2H of the result rows inside the timed region.
reference:  the CPU oracle port (oracle/tad_oracle.c, the restated reference job; the reference
            itself is PySpark and cannot run in this image) on all host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "flow records/sec through EWMA anomaly detection"
BYTES_PER_ROW = 29          # src_ip4 dst_ip4 src_port2 dst_port2 proto1 flow_start4 flow_end4 value8
RESULT_ROW_BYTES = 46


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons: one streaming `nvidia-smi -lms 20` process; only the samples whose
    timestamps fall inside the timed region are kept."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.lines, self.proc, self.th = index, [], None, None
        self.t0 = self.t1 = None

    def _pump(self):
        try:
            for line in self.proc.stdout:
                self.lines.append((time.time(), line))
        except Exception:
            pass

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
        rows = []
        for ts, line in self.lines:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) >= 7:
                rows.append((ts, f[1:]))
        inside = [f for ts, f in rows if self.t0 is not None and self.t0 - 0.02 <= ts <= self.t1 + 0.02]
        use = inside if inside else [f for _, f in rows[-3:]]
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(f[0]) for f in use if f[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(f[2 + i].lower().startswith("active") for f in use)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(use[0][1]) if use[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(use), "samples_in_timed_region": len(inside)}


def ncu_traffic(kernel, rows):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), valid for
    the 1e8-row workload it was taken on; None otherwise."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        d = json.load(open(p))
        if d.get("rows") == rows and kernel in d["kernels"]:
            return d["kernels"][kernel]["dram_bytes_read"] + d["kernels"][kernel]["dram_bytes_write"]
    except Exception:
        pass
    return None


def python_port_rate(points):
    """Single-core rate of the pure-Python restatement (numpy stages A/B + the per-series UDF loops the Spark workers
    would run, oracle/tad_oracle.py) on a 2e5-row sample: the 'local[1]'-style figure SURVEY section 8(d) asks for.
    Reported next to the C port; never fatal."""
    try:
        from oracle import tad_oracle
        from theia_b200 import synth
        t = synth.make_flows(max(1, 200_000 // points), points, seed=1)
        tad_oracle.run_job(t, tad_oracle.JobSpec())
        t0 = time.perf_counter()
        tad_oracle.run_job(t, tad_oracle.JobSpec())
        dt = time.perf_counter() - t0
        return {"value": len(t["value"]) / dt, "unit": "records/s", "cores": 1,
                "sample": "%d rows, oracle/tad_oracle.py" % len(t["value"])}
    except Exception as e:       # the bench line must not depend on this side figure
        return {"unavailable": repr(e)[:120]}


def native_decode_rate():
    """Host-side rate of the ClickHouse Native-format ingest (theia_b200/clickhouse_native.py): 1e6 flow rows of the eight
    engine columns, IPs as text.  A side figure for the transport that replaces JDBC; never fatal."""
    try:
        from theia_b200 import clickhouse_native as chn
        from theia_b200 import synth
        t = synth.make_flows(10_000, 100, seed=2)
        cols = [("sourceIP", "String", t["src_ip"]), ("destinationIP", "String", t["dst_ip"]),
                ("sourceTransportPort", "UInt16", t["src_port"]), ("destinationTransportPort", "UInt16", t["dst_port"]),
                ("protocolIdentifier", "UInt8", t["proto"]), ("flowStartSeconds", "DateTime", t["flow_start"]),
                ("flowEndSeconds", "DateTime", t["flow_end"]), ("throughput", "UInt64", t["value"])]
        n = len(t["value"])
        stream = b"".join(chn.write_native([(k, ty, v[lo:lo + 65536]) for k, ty, v in cols]) for lo in range(0, n, 65536))
        chn.flows_from_native(stream)
        t0 = time.perf_counter()
        chn.flows_from_native(stream)
        dt = time.perf_counter() - t0
        return {"value": n / dt, "unit": "records/s", "bytes_per_record": len(stream) / n,
                "sample": "%d rows in 64 Ki-row blocks, IPv4 addresses as text" % n}
    except Exception as e:
        return {"unavailable": repr(e)[:120]}


def run_reference(args):
    """CPU arm: the oracle port with all host threads on a bounded sample of the same workload."""
    import numpy as np
    from oracle import c_oracle
    from theia_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    series = args.ref_series
    t = synth.make_flows(series, args.points, seed=1)
    rows = len(t["value"])
    c_oracle.build()
    for _ in range(args.warmup):
        c_oracle.run_job(t, algo=0, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cols, ns, npts = c_oracle.run_job(t, algo=0, threads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = rows / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "records/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "EWMA, %d records / %d connections x %d points per GPU (BASELINE configs[1])" % (args.series * args.points, args.series, args.points),
                   "sample": "each step runs the CPU path on a bounded sample: %d connections x %d points = %d rows" % (series, args.points, rows),
                   "rows_per_step": rows},
        "cpu_baseline": {"value": v, "unit": "records/s", "cores": cores, "kind": "port",
                         "sample": "%d rows per step, oracle/tad_oracle.c with OpenMP on %d threads" % (rows, cores),
                         "python_port_1core": python_port_rate(args.points)},
        "e2e": {"value": v, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_ours(args):
    import numpy as np
    import torch
    from theia_b200 import synth
    from theia_b200.engine import DeviceColumns, TadEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        buf = [TadEngine.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(buf, src=0)
        uid = buf[0]
    eng = TadEngine(device=local, world_size=world, rank=rank, nccl_unique_id=uid)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()              # early: nvidia-smi takes a second or more to enumerate the GPUs

    S, n = args.series, args.points
    cols_t = synth.make_flows_torch(S, n, seed=1 + rank, device=dev) if world == 1 else \
        synth.make_flows_torch_sharded(S, n, seed=1, device=dev, rank=rank, world=world)
    rows = int(cols_t["value"].numel())
    torch.cuda.synchronize()
    dcols = DeviceColumns(rows, {k: v.data_ptr() for k, v in cols_t.items()})
    dcols.keepalive = cols_t

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(cols):
        job = eng.submit(cols, algo=args.algo, tad_id="bench")
        st = job.wait()
        return job, st

    # ---- value: inputs resident in HBM ---------------------------------------------------------
    for _ in range(args.warmup):
        job, st = step(dcols)
        job.release()
    barrier()
    sampler.mark_begin()
    t0 = time.perf_counter()
    dev_ms, phase, launches, result_rows = 0.0, {}, 0, 0
    for _ in range(args.steps):
        job, st = step(dcols)
        dev_ms += st["device_ms"]
        launches += st["gpu_launches"]
        result_rows = st["result_rows"]
        for k, v in st["phase_ms"].items():
            phase[k] = phase.get(k, 0.0) + v
        job.release()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    stats = torch.tensor([dev_ms / args.steps, wall_ms / args.steps], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_dev, ms_wall = float(stats[0]), float(stats[1])
    tr = torch.tensor([rows], device=dev, dtype=torch.int64)
    if dist is not None:
        dist.all_reduce(tr)
    total_rows = int(tr[0])

    # ---- e2e: host (pinned) buffers through the C ABI, H2D + D2H inside the timed region --------
    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"value": total_rows / (ms_dev * 1e-3), "ms_per_step": ms_dev,
                              "phase_ms": {k: v / args.steps for k, v in phase.items()}, "note": "profiling run"}))
        eng.close()
        if dist is not None:
            dist.destroy_process_group()
        return
    hcols = eng.alloc_columns(rows)
    for name, tns in cols_t.items():
        hv = hcols.view(name)
        hv[:rows] = tns.cpu().numpy().view(hv.dtype)
    hcols.c.rows = rows
    for _ in range(max(1, args.warmup // 2)):
        job, st = step(hcols)
        job.result(copy=False)
        job.release()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, args.steps // 2)
    for _ in range(e2e_steps):
        job, st = step(hcols)
        res = job.result(copy=False)
        d2h = int(st["result_rows"]) * RESULT_ROW_BYTES
        h2d_ms = st["phase_ms"]["h2d"]
        job.release()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    e2e_t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_t[0])

    # ---- e2e, two jobs in flight: a second context (own stream + workspace) on the same GPU, jobs submitted
    # alternately so that the H2D of job i+1 overlaps the compute and D2H of job i (how the controller's 4
    # workers drive tad_submit).  Reported separately; `e2e.value` above is one job at a time.
    pipe_ms = None
    if world == 1 and not args.no_pipelined:
        eng2 = TadEngine(device=local)
        hcols2 = eng2.alloc_columns(rows)
        for name in cols_t:
            hcols2.view(name)[:rows] = hcols.view(name)[:rows]
        hcols2.c.rows = rows
        engines = [(eng, hcols), (eng2, hcols2)]
        for e_, c_ in engines:                       # warm both workspaces
            j_ = e_.submit(c_, algo="EWMA", tad_id="bench")
            j_.wait(); j_.result(copy=False); j_.release()
        torch.cuda.synchronize()
        n_pipe = max(4, args.steps)
        t0 = time.perf_counter()
        inflight = []
        for k in range(n_pipe):
            e_, c_ = engines[k % 2]
            if len(inflight) == 2:
                j_ = inflight.pop(0)
                j_.wait(); j_.result(copy=False); j_.release()
            inflight.append(e_.submit(c_, algo="EWMA", tad_id="bench"))
        for j_ in inflight:
            j_.wait(); j_.result(copy=False); j_.release()
        torch.cuda.synchronize()
        pipe_ms = (time.perf_counter() - t0) * 1e3 / n_pipe
        hcols2.free()
        eng2.close()

    if rank == 0:
        peak, peak_src = peaks()
        per = {k: v / args.steps for k, v in phase.items()}
        kern = {k: per[k] for k in ("hist", "scatter", "group", "detect") if per.get(k, 0) > 0}
        dom = max(kern, key=kern.get)
        alg_bytes = {"hist": 17, "scatter": 29, "group": 29, "detect": 29}[dom] * rows
        achieved = alg_bytes / (kern[dom] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": total_rows / (ms_dev * 1e-3), "unit": "records/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "EWMA, %d records / %d connections x %d points per GPU (BASELINE configs[1])" % (rows, S, n),
                       "rows_per_gpu": rows, "l2": "inputs (%.1f GB) larger than L2" % (rows * BYTES_PER_ROW / 1e9),
                       "timing": "library CUDA events on its stream; wall clock per step %.3f ms" % ms_wall},
            "e2e": {"value": total_rows / (e2e_ms * 1e-3), "unit": "records/s",
                    "h2d_bytes_per_step": rows * BYTES_PER_ROW, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "h2d_ms": h2d_ms,
                    "two_jobs_in_flight": None if pipe_ms is None else
                    {"value": rows / (pipe_ms * 1e-3), "ms_per_step": pipe_ms,
                     "note": "two contexts on one GPU, H2D of job i+1 overlaps compute + D2H of job i"}},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(dom, rows), "peak_source": peak_src,
                         "note": "achieved = algorithmic bytes (29 B/row; hist 17) / CUDA-event time of the phase; the group phase is "
                                 "three launches of one kernel template (capacity classes)",
                         "pipeline_frac": rows * BYTES_PER_ROW / (ms_dev * 1e-3) / 1e9 / peak},
            "phase_ms": per, "result_rows": result_rows, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            from oracle import c_oracle
            cores = os.cpu_count() or 1
            t = synth.make_flows(args.ref_series, n, seed=1)
            c_oracle.build()
            c_oracle.run_job(t, algo=0, threads=cores)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                c_oracle.run_job(t, algo=0, threads=cores)
            dt = (time.perf_counter() - t0) / reps
            line["cpu_baseline"] = {"value": len(t["value"]) / dt, "unit": "records/s", "cores": cores, "kind": "port",
                                    "sample": "%d rows x %d reps, oracle/tad_oracle.c OpenMP" % (len(t["value"]), reps),
                                    "python_port_1core": python_port_rate(n)}
            line["native_ingest"] = native_decode_rate()
        print(json.dumps(line))
    hcols.free()
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--series", type=int, default=1_000_000)
    ap.add_argument("--points", type=int, default=100)
    ap.add_argument("--ref-series", type=int, default=100_000, help="connections in the CPU sample")
    ap.add_argument("--algo", default="EWMA", choices=["EWMA", "DBSCAN", "ARIMA"],
                    help="detector (the BASELINE metric is EWMA; the others are side measurements of configs[2]/[3])")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-buffer leg")
    ap.add_argument("--no-pipelined", action="store_true", help="skip the two-jobs-in-flight e2e figure")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
