"""Python host over the C ABI: the calls the reference-side shim makes, in the order it makes them.

    eng = TadEngine(device=0)
    cols = eng.columns_from_numpy(table)          # fill library-owned pinned buffers
    job = eng.submit(cols, algo="EWMA", tad_id=...)   # non-blocking, like CreateSparkApplication
    st = job.wait()                               # or poll(): state / completed_stages / total_stages
    rows = job.result()                           # dict of numpy arrays (tadetector columns)
    job.release(); cols.free(); eng.close()
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

_COLS = (("src_ip", np.uint32), ("dst_ip", np.uint32), ("src_port", np.uint16), ("dst_port", np.uint16),
         ("proto", np.uint8), ("flow_start", np.uint32), ("flow_end", np.uint32), ("value", np.uint64))
_NS_COLS = (("src_ns", np.uint32), ("dst_ns", np.uint32))
_OUT = (("src_ip", np.uint32), ("dst_ip", np.uint32), ("src_port", np.uint16), ("dst_port", np.uint16),
        ("proto", np.uint8), ("flow_start", np.uint32), ("flow_end", np.uint32), ("stddev", np.float64),
        ("algo_calc", np.float64), ("throughput", np.float64), ("anomaly", np.uint8))


class TadError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        name = L.load().tad_strerror(code).decode()
        super().__init__("%s (%d)%s" % (name, code, (": " + msg) if msg else ""))


def _status_dict(st: L.TadStatus) -> dict:
    d = {k: getattr(st, k) for k in ("completed_stages", "total_stages", "error", "rows_in", "rows_kept", "rows_owned",
                                     "points", "series", "result_rows", "spill_rows", "gpu_launches", "device_ms",
                                     "total_ms")}
    d["state"] = L.STATE_NAMES[st.state]
    d["err_msg"] = st.err_msg.decode(errors="replace")
    d["phase_ms"] = {n: st.phase_ms[i] for i, n in enumerate(L.PHASE_NAMES)}
    return d


class Columns:
    """Library-owned pinned-host (or device) column buffers (tad_alloc_columns)."""

    def __init__(self, eng: "TadEngine", capacity: int, mem: int = L.TAD_MEM_HOST):
        self.eng = eng
        self.c = L.TadColumns()
        rc = eng.lib.tad_alloc_columns(eng.ctx, capacity, mem, C.byref(self.c))
        if rc != 0:
            raise TadError(rc, "tad_alloc_columns(%d)" % capacity)
        self.owned = True

    def view(self, name: str) -> np.ndarray:
        """numpy view of a host column (capacity elements)."""
        assert self.c.mem == L.TAD_MEM_HOST
        dt = dict(_COLS + _NS_COLS)[name]
        n = int(self.c.capacity)
        ptr = getattr(self.c, name)
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt, count=n)

    def fill(self, table: dict):
        n = len(table["flow_end"])
        assert n <= self.c.capacity
        for name, dt in _COLS:
            a = table.get(name)
            if a is None:
                self.view(name)[:n] = 0
            else:
                self.view(name)[:n] = np.asarray(a, dtype=dt)
        if table.get("src_ns") is not None or table.get("dst_ns") is not None:
            rc = self.eng.lib.tad_alloc_ns_columns(self.eng.ctx, C.byref(self.c))
            if rc != 0:
                raise TadError(rc, "tad_alloc_ns_columns")
            for name, dt in _NS_COLS:
                a = table.get(name)
                self.view(name)[:n] = 0 if a is None else np.asarray(a, dtype=dt)
        self.c.rows = n
        return self

    def free(self):
        if self.owned and self.c.capacity:
            self.eng.lib.tad_free_columns(self.eng.ctx, C.byref(self.c))
            self.owned = False


class DeviceColumns:
    """Columns that already live in HBM (raw device pointers, e.g. torch tensors' data_ptr())."""

    def __init__(self, rows: int, ptrs: dict):
        self.c = L.TadColumns()
        self.c.rows = rows
        self.c.capacity = rows
        self.c.mem = L.TAD_MEM_DEVICE
        for name, _ in _COLS:
            setattr(self.c, name, ptrs.get(name))
        self.keepalive = ptrs

    def free(self):
        pass


class Job:
    def __init__(self, eng, handle, keep):
        self.eng = eng
        self.h = handle
        self._keep = keep

    def poll(self) -> dict:
        st = L.TadStatus()
        self.eng.lib.tad_poll(self.h, C.byref(st))
        return _status_dict(st)

    def wait(self, timeout_ms: int = -1, check: bool = True) -> dict:
        st = L.TadStatus()
        self.eng.lib.tad_wait(self.h, timeout_ms, C.byref(st))
        d = _status_dict(st)
        if check and d["state"] == "FAILED":
            raise TadError(st.error, d["err_msg"])
        return d

    def result(self, copy: bool = True) -> dict:
        r = L.TadRows()
        rc = self.eng.lib.tad_result(self.h, C.byref(r))
        if rc != 0:
            raise TadError(rc, "tad_result")
        n = int(r.rows)
        out = {}
        for name, dt in _OUT:
            if n == 0:
                out[name] = np.zeros(0, dtype=dt)
                continue
            buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(getattr(r, name))
            a = np.frombuffer(buf, dtype=dt, count=n)
            out[name] = a.copy() if copy else a
        return out

    def cancel(self):
        self.eng.lib.tad_cancel(self.h)

    def release(self):
        if self.h:
            self.eng.lib.tad_release(self.h)
            self.h = None


class TadEngine:
    def __init__(self, device: int = 0, world_size: int = 1, rank: int = 0, nccl_unique_id: bytes | None = None):
        self.lib = L.load()
        cfg = L.TadConfig()
        cfg.device, cfg.world_size, cfg.rank = device, world_size, rank
        self._uid = None
        if nccl_unique_id is not None:
            self._uid = C.create_string_buffer(bytes(nccl_unique_id), 128)
            cfg.nccl_unique_id = C.cast(self._uid, C.c_void_p)
            cfg.nccl_unique_id_bytes = 128
        ctx = C.c_void_p()
        rc = self.lib.tad_init(C.byref(cfg), C.byref(ctx))
        if rc != 0:
            raise TadError(rc, "tad_init(device=%d)" % device)
        self.ctx = ctx

    @staticmethod
    def get_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = L.load().tad_get_unique_id(buf, 128)
        if rc != 0:
            raise TadError(rc, "tad_get_unique_id")
        return buf.raw

    def alloc_columns(self, capacity: int, mem: int = L.TAD_MEM_HOST) -> Columns:
        return Columns(self, capacity, mem)

    def columns_from_numpy(self, table: dict) -> Columns:
        return self.alloc_columns(max(1, len(table["flow_end"]))).fill(table)

    def submit(self, cols, algo="EWMA", reducer: int = L.TAD_REDUCE_MAX, start_time: int = 0, end_time: int = 0,
               tad_id: str = "", emit_all: bool = False, ns_ignore=(), check: bool = True, global_rows: int = 0) -> Job:
        spec = L.TadJobSpec()
        spec.algo = L.ALGOS[algo] if isinstance(algo, str) else int(algo)
        spec.reducer = reducer
        spec.start_time, spec.end_time = int(start_time), int(end_time)
        spec.flags = L.TAD_FLAG_EMIT_ALL if emit_all else 0
        spec.id = tad_id.encode()[:39]
        spec.global_rows = int(global_rows)
        keep = [cols]
        if len(ns_ignore):
            arr = np.asarray(ns_ignore, dtype=np.uint32)
            spec.n_ns_ignore = len(arr)
            spec.ns_ignore = arr.ctypes.data
            keep.append(arr)
        h = C.c_void_p()
        rc = self.lib.tad_submit(self.ctx, C.byref(spec), C.byref(cols.c), C.byref(h))
        job = Job(self, h, keep)
        if rc != 0 and check:
            msg = job.poll()["err_msg"] if h else ""
            job.release()
            raise TadError(rc, msg)
        return job

    def run(self, table: dict, on_job=None, **kw):
        """Convenience: numpy table in, (result dict, status dict) out.  ``on_job(job)`` is called right after tad_submit
        (the caller may poll it from another thread) and ``on_job(None)`` before the handle is released."""
        cols = self.columns_from_numpy(table)
        job = self.submit(cols, **kw)
        if on_job is not None:
            on_job(job)
        try:
            st = job.wait()
            return job.result(), st
        finally:
            if on_job is not None:
                on_job(None)
            job.release()
            cols.free()

    def close(self):
        if self.ctx:
            self.lib.tad_shutdown(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
