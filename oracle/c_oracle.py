"""ctypes wrapper of the C oracle port (oracle/tad_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libtad_oracle.so")
_lib = None


class _Table(C.Structure):
    _fields_ = [("rows", C.c_uint64),
                ("src_ip", C.c_void_p), ("dst_ip", C.c_void_p), ("flow_start", C.c_void_p),
                ("flow_end", C.c_void_p), ("src_port", C.c_void_p), ("dst_port", C.c_void_p),
                ("proto", C.c_void_p), ("value", C.c_void_p)]


class _Spec(C.Structure):
    _fields_ = [("algo", C.c_int32), ("reducer", C.c_int32), ("start_time", C.c_uint32),
                ("end_time", C.c_uint32), ("emit_all", C.c_int32), ("threads", C.c_int32)]


class _Result(C.Structure):
    _fields_ = [("n", C.c_uint64), ("n_series", C.c_uint64), ("n_points", C.c_uint64),
                ("src_ip", C.POINTER(C.c_uint32)), ("dst_ip", C.POINTER(C.c_uint32)),
                ("flow_start", C.POINTER(C.c_uint32)), ("flow_end", C.POINTER(C.c_uint32)),
                ("src_port", C.POINTER(C.c_uint16)), ("dst_port", C.POINTER(C.c_uint16)),
                ("proto", C.POINTER(C.c_uint8)), ("anomaly", C.POINTER(C.c_uint8)),
                ("stddev", C.POINTER(C.c_double)), ("algo_calc", C.POINTER(C.c_double)),
                ("throughput", C.POINTER(C.c_double))]


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in ("tad_oracle.c", "Makefile")]
    if (force or not os.path.exists(LIB_PATH)
            or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.tad_oracle_run.argtypes = [C.POINTER(_Table), C.POINTER(_Spec), C.POINTER(_Result)]
        _lib.tad_oracle_run.restype = C.c_int
        _lib.tad_oracle_free.argtypes = [C.POINTER(_Result)]
    return _lib


_DT = {"src_ip": np.uint32, "dst_ip": np.uint32, "flow_start": np.uint32, "flow_end": np.uint32,
       "src_port": np.uint16, "dst_port": np.uint16, "proto": np.uint8, "value": np.uint64}


def run_job(table: dict, algo: int = 0, reducer: int = 0, start_time: int = 0, end_time: int = 0,
            emit_all: bool = False, threads: int = 0):
    """Returns (cols dict in engine column naming, n_series, n_points); rows unordered."""
    keep = {}
    t = _Table()
    t.rows = len(table["flow_end"])
    for name, dt in _DT.items():
        a = table.get(name)
        if a is None:
            setattr(t, name, None)
            continue
        a = np.ascontiguousarray(a, dtype=dt)
        keep[name] = a
        setattr(t, name, a.ctypes.data)
    sp = _Spec(algo, reducer, start_time, end_time, int(emit_all), threads)
    r = _Result()
    rc = lib().tad_oracle_run(C.byref(t), C.byref(sp), C.byref(r))
    if rc != 0:
        raise RuntimeError("tad_oracle_run failed: %d" % rc)
    n = int(r.n)

    def grab(ptr, dt):
        return np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].astype(dt, copy=True)

    cols = {"src_ip": grab(r.src_ip, np.uint32), "src_port": grab(r.src_port, np.uint16),
            "dst_ip": grab(r.dst_ip, np.uint32), "dst_port": grab(r.dst_port, np.uint16),
            "proto": grab(r.proto, np.uint8), "flow_start": grab(r.flow_start, np.uint32),
            "flow_end": grab(r.flow_end, np.uint32), "stddev": grab(r.stddev, np.float64),
            "algo_calc": grab(r.algo_calc, np.float64), "throughput": grab(r.throughput, np.float64),
            "anomaly": grab(r.anomaly, np.uint8)}
    out = (cols, int(r.n_series), int(r.n_points))
    lib().tad_oracle_free(C.byref(r))
    return out


def ewma_series(values):
    v = np.ascontiguousarray(values, dtype=np.uint64)
    n = len(v)
    calc = np.zeros(max(n, 1), dtype=np.float64)
    flag = np.zeros(max(n, 1), dtype=np.uint8)
    sd = C.c_double()
    L = lib()
    L.tad_oracle_ewma(v.ctypes.data_as(C.c_void_p), C.c_uint32(n), calc.ctypes.data_as(C.c_void_p),
                      flag.ctypes.data_as(C.c_void_p), C.byref(sd))
    return calc[:n], flag[:n].astype(bool), sd.value


def dbscan_series(values):
    v = np.ascontiguousarray(values, dtype=np.uint64)
    n = len(v)
    flag = np.zeros(max(n, 1), dtype=np.uint8)
    lib().tad_oracle_dbscan(v.ctypes.data_as(C.c_void_p), C.c_uint32(n), flag.ctypes.data_as(C.c_void_p))
    return flag[:n].astype(bool)
