"""CPU oracle for the Theia throughput-anomaly-detection (TAD) hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product package ``theia_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and there
only as the checker / reported baseline.

What this restates (paths relative to the reference repository root):

* stage A  ``plugins/anomaly-detection/anomaly_detection.py:566-613``  -- the SQL
  that ClickHouse executes: optional ``flowStartSeconds >= start`` /
  ``flowEndSeconds < end`` filters, ``GROUP BY <key cols>, flowEndSeconds`` with
  ``max(throughput)`` (``sum(throughput)`` for the aggregated-flow modes,
  ``:63-106``).
* stage B  ``anomaly_detection.py:680-684`` -- ``groupby(key).agg(collect_list(
  flowEndSeconds), stddev_samp(max(throughput)), collect_list(max(throughput)))``.
  ``collect_list`` order is undefined in the reference; the engine contract (and
  this oracle) orders every series by ``flowEndSeconds`` ascending.
  ``stddev_samp`` is Spark's ``CentralMomentAgg`` (Apache Spark 3.x, not vendored in
  the reference tree): Welford update ``n+=1; d=x-avg; dn=d/n; avg+=dn;
  m2+=d*(d-dn)``, result ``sqrt(m2/(n-1))``, NULL for n == 1.
* stage C  ``anomaly_detection.py:146-165`` (calculate_ewma), ``:168-212``
  (calculate_ewma_anomaly), ``:312-349`` (calculate_dbscan[_anomaly]);
  DBSCAN is scikit-learn 1.3.0 ``DBSCAN(min_samples=4, eps=250000000)`` on the 1-D
  values (``requirements.txt``), restated as the exact 1-D rule including the
  brute-force distance expansion sklearn uses for n <= 11.
* stage D  ``anomaly_detection.py:352-421`` -- explode, keep ``anomaly`` rows.
* stage E  column order of ``default.tadetector``
  (``build/charts/theia/provisioning/datasources/create_table.sh:363-384``).

Pinned by ``tests/test_oracle_golden.py`` against (i) the reference's own golden
vectors (``anomaly_detection_test.py:199-402``) and (ii) outputs of the reference
UDFs themselves, generated in the build container by
``tests/golden/make_golden.py`` and committed under ``tests/golden/``.

ARIMA lives in ``oracle/arima_oracle.py`` (parity status documented there).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

ALGO_EWMA, ALGO_ARIMA, ALGO_DBSCAN = 0, 1, 2
ALGO_NAMES = {"EWMA": ALGO_EWMA, "ARIMA": ALGO_ARIMA, "DBSCAN": ALGO_DBSCAN}
REDUCE_MAX, REDUCE_SUM = 0, 1

EWMA_ALPHA = 0.5            # anomaly_detection.py:157
DBSCAN_EPS = 250000000.0    # anomaly_detection.py:342
DBSCAN_MIN_SAMPLES = 4      # anomaly_detection.py:342

KEY_COLUMNS = ("src_ip", "src_port", "dst_ip", "dst_port", "proto", "flow_start")
KEY_DTYPES = {"src_ip": np.uint32, "src_port": np.uint16, "dst_ip": np.uint32,
              "dst_port": np.uint16, "proto": np.uint8, "flow_start": np.uint32}


# --------------------------------------------------------------------------- #
# stage C: per-series arithmetic
# --------------------------------------------------------------------------- #
def u64_to_f64(values) -> np.ndarray:
    """``float(Decimal(u64))`` of anomaly_detection.py:161 -- round-to-nearest-even."""
    return np.asarray(values, dtype=np.uint64).astype(np.float64)


def calculate_ewma(values) -> np.ndarray:
    """anomaly_detection.py:146-165.  e_0 = 0; e_i = (1-a)*e_{i-1} + a*x_i, a = 0.5."""
    x = u64_to_f64(values)
    out = np.empty(len(x), dtype=np.float64)
    prev = 0.0
    one_minus = 1.0 - EWMA_ALPHA
    for i in range(len(x)):
        prev = one_minus * prev + EWMA_ALPHA * float(x[i])
        out[i] = prev
    return out


def stddev_samp(values):
    """Spark CentralMomentAgg restated (see module docstring).  None for n < 2."""
    x = u64_to_f64(values)
    n = 0.0
    avg = 0.0
    m2 = 0.0
    for i in range(len(x)):
        n += 1.0
        d = float(x[i]) - avg
        dn = d / n
        avg = avg + dn
        m2 = m2 + d * (d - dn)
    if n < 2.0:
        return None
    return math.sqrt(m2 / (n - 1.0))


def calculate_ewma_anomaly(values, stddev) -> np.ndarray:
    """anomaly_detection.py:168-212.  flag_i = abs(x_i - e_i) > stddev (strict);
    stddev None -> all False (:198-201)."""
    x = u64_to_f64(values)
    if stddev is None:
        return np.zeros(len(x), dtype=bool)
    e = calculate_ewma(values)
    return np.abs(x - e) > float(stddev)


def _sk_brute_d2(xi: float, xj: float) -> float:
    """Squared distance as sklearn's brute-force radius search computes it for
    float64 inputs: ||x||^2 - 2 x.y + ||y||^2 evaluated as
    fl(fl(fl(xi*xi) + fl(-2*fl(xi*xj))) + fl(xj*xj)), clipped at 0
    (scikit-learn 1.3.0 ``_middle_term_computer`` / ``euclidean_distances``)."""
    d = (xi * xi + (-2.0 * (xi * xj))) + xj * xj
    return d if d > 0.0 else 0.0


def calculate_dbscan_anomaly(values) -> np.ndarray:
    """anomaly_detection.py:325-349: DBSCAN(min_samples=4, eps=2.5e8) labels == -1.

    1-D rule.  Neighbourhood is inclusive (<= eps) and contains the point itself.
    n >= 12 -> sklearn picks a KD-tree (``n_neighbors=5 < n // 2``) whose reduced
    distance test is equivalent to ``abs(xi - xj) <= eps`` in float64; n <= 11 ->
    brute force with the expansion of ``_sk_brute_d2`` compared against eps^2
    (self-distance forced to 0).  core <=> >= 4 neighbours; noise <=> not core and
    no core neighbour.
    """
    x = u64_to_f64(values)
    n = len(x)
    nb = np.zeros((n, n), dtype=bool)
    if n >= 12:
        for i in range(n):
            nb[i] = np.abs(x[i] - x) <= DBSCAN_EPS
    else:
        r2 = DBSCAN_EPS * DBSCAN_EPS
        for i in range(n):
            for j in range(n):
                nb[i, j] = True if i == j else (_sk_brute_d2(float(x[i]), float(x[j])) <= r2)
    core = nb.sum(axis=1) >= DBSCAN_MIN_SAMPLES
    reach = (nb & core[None, :]).any(axis=1)
    return ~(core | reach)


def calculate_dbscan(values) -> np.ndarray:
    """anomaly_detection.py:312-322 -- placeholder zeros."""
    return np.zeros(len(values), dtype=np.float64)


# --------------------------------------------------------------------------- #
# stages A, B, D, E on a columnar table
# --------------------------------------------------------------------------- #
@dataclass
class JobSpec:
    algo: int = ALGO_EWMA
    reducer: int = REDUCE_MAX
    start_time: int = 0          # 0 = unbounded; filter flow_start >= start (:581-583)
    end_time: int = 0            # 0 = unbounded; filter flow_end   <  end   (:584-586)
    emit_all: bool = False       # debugging / parity: emit every point, with its flag
    ns_ignore: tuple = ()        # namespace ids; needs src_ns / dst_ns columns (:576-580)


@dataclass
class Result:
    """Rows of stage D in canonical order (key cols, then flow_end)."""
    cols: dict = field(default_factory=dict)
    n_series: int = 0
    n_points: int = 0            # rows after stage A

    def __len__(self):
        return len(self.cols["flow_end"]) if self.cols else 0


OUT_COLUMNS = KEY_COLUMNS + ("flow_end", "stddev", "algo_calc", "throughput", "anomaly")


def _full_columns(table: dict) -> dict:
    n = len(table["flow_end"])
    out = {}
    for c in KEY_COLUMNS:
        out[c] = (np.asarray(table[c], dtype=KEY_DTYPES[c]) if c in table and table[c] is not None
                  else np.zeros(n, dtype=KEY_DTYPES[c]))
    out["flow_end"] = np.asarray(table["flow_end"], dtype=np.uint32)
    out["value"] = np.asarray(table["value"], dtype=np.uint64)
    return out


def run_job(table: dict, spec: JobSpec, arima_fn=None) -> Result:
    """Stages A-D.  ``table`` maps column name -> 1-D array (absent key columns are
    zero).  Returns the anomalous points (or all points when ``spec.emit_all``)."""
    t = _full_columns(table)
    n = len(t["flow_end"])
    keep = np.ones(n, dtype=bool)
    if spec.start_time:
        keep &= t["flow_start"] >= np.uint32(spec.start_time)
    if spec.end_time:
        keep &= t["flow_end"] < np.uint32(spec.end_time)
    if spec.ns_ignore:
        ign = np.asarray(spec.ns_ignore, dtype=np.uint32)
        keep &= ~np.isin(np.asarray(table["src_ns"], dtype=np.uint32), ign)
        keep &= ~np.isin(np.asarray(table["dst_ns"], dtype=np.uint32), ign)
    t = {k: v[keep] for k, v in t.items()}
    n = len(t["flow_end"])

    # canonical sort: key columns (most significant first), then flow_end
    order = np.lexsort((t["flow_end"],) + tuple(t[c] for c in reversed(KEY_COLUMNS)))
    t = {k: v[order] for k, v in t.items()}
    res = Result()
    out = {c: [] for c in OUT_COLUMNS}
    if n == 0:
        res.cols = {c: np.zeros(0, dtype=_out_dtype(c)) for c in OUT_COLUMNS}
        return res

    key_change = np.zeros(n, dtype=bool)
    key_change[0] = True
    for c in KEY_COLUMNS:
        key_change[1:] |= t[c][1:] != t[c][:-1]
    point_change = key_change.copy()
    point_change[1:] |= t["flow_end"][1:] != t["flow_end"][:-1]
    # stage A: reduce duplicates of (key, flow_end)
    pstart = np.flatnonzero(point_change)
    if spec.reducer == REDUCE_MAX:
        pvalue = np.maximum.reduceat(t["value"], pstart)
    else:
        pvalue = np.add.reduceat(t["value"], pstart)        # wraps mod 2^64 like UInt64
    pts = {k: v[pstart] for k, v in t.items()}
    pts["value"] = pvalue
    skey = key_change[pstart]
    sstart = np.flatnonzero(skey)
    send = np.append(sstart[1:], len(pstart))
    res.n_series = len(sstart)
    res.n_points = len(pstart)

    for s0, s1 in zip(sstart, send):
        vals = pts["value"][s0:s1]
        x = u64_to_f64(vals)
        sd = stddev_samp(vals)
        if spec.algo == ALGO_EWMA:
            calc = calculate_ewma(vals)
            flag = calculate_ewma_anomaly(vals, sd)
        elif spec.algo == ALGO_DBSCAN:
            calc = calculate_dbscan(vals)
            flag = calculate_dbscan_anomaly(vals)
        elif spec.algo == ALGO_ARIMA:
            if arima_fn is None:
                raise ValueError("ARIMA needs arima_fn (oracle/arima_oracle.py)")
            calc = arima_fn(vals)
            if calc is None:                      # :232-234 / :260-264 -> no rows
                continue
            flag = (np.zeros(len(x), dtype=bool) if sd is None
                    else np.abs(x - calc) > float(sd))
        else:
            raise ValueError("bad algo")
        sel = np.arange(s1 - s0) if spec.emit_all else np.flatnonzero(flag)
        if len(sel) == 0:
            continue
        for c in KEY_COLUMNS:
            out[c].append(np.repeat(pts[c][s0], len(sel)))
        out["flow_end"].append(pts["flow_end"][s0:s1][sel])
        out["stddev"].append(np.repeat(np.nan if sd is None else sd, len(sel)))
        out["algo_calc"].append(np.asarray(calc)[sel])
        out["throughput"].append(x[sel])
        out["anomaly"].append(np.asarray(flag)[sel].astype(np.uint8))
    res.cols = {c: (np.concatenate(out[c]).astype(_out_dtype(c)) if out[c]
                    else np.zeros(0, dtype=_out_dtype(c))) for c in OUT_COLUMNS}
    return res


def _out_dtype(c):
    if c in KEY_DTYPES:
        return KEY_DTYPES[c]
    return {"flow_end": np.uint32, "stddev": np.float64, "algo_calc": np.float64,
            "throughput": np.float64, "anomaly": np.uint8}[c]


def canonicalize(cols: dict) -> dict:
    """Sort result columns into the canonical (key cols, flow_end) order."""
    if len(cols["flow_end"]) == 0:
        return {k: np.asarray(v) for k, v in cols.items()}
    order = np.lexsort((cols["flow_end"],) + tuple(cols[c] for c in reversed(KEY_COLUMNS)))
    return {k: np.asarray(v)[order] for k, v in cols.items()}


# --------------------------------------------------------------------------- #
# stage E: tadetector rows (create_table.sh:363-384; anomaly_detection.py:385-420,500-503)
# --------------------------------------------------------------------------- #
def ip_to_str(ip: int) -> str:
    return "%d.%d.%d.%d" % ((ip >> 24) & 255, (ip >> 16) & 255, (ip >> 8) & 255, ip & 255)


def tadetector_rows(res: Result, algo_name: str, tad_id: str, agg_type: str = "None",
                    now_str: str = "1970-01-01 00:00:00") -> list:
    """Row dicts as the reference appends them to ``default.tadetector``; one
    sentinel row when no anomaly was found (anomaly_detection.py:395-420)."""
    c = res.cols
    rows = []
    for i in range(len(res)):
        if not c["anomaly"][i]:
            continue
        rows.append({
            "sourceIP": ip_to_str(int(c["src_ip"][i])),
            "sourceTransportPort": int(c["src_port"][i]),
            "destinationIP": ip_to_str(int(c["dst_ip"][i])),
            "destinationTransportPort": int(c["dst_port"][i]),
            "protocolIdentifier": int(c["proto"][i]),
            "flowStartSeconds": int(c["flow_start"][i]),
            "flowEndSeconds": int(c["flow_end"][i]),
            "throughputStandardDeviation": float(c["stddev"][i]),
            "aggType": agg_type, "algoType": algo_name,
            "algoCalc": float(c["algo_calc"][i]),
            "throughput": float(c["throughput"][i]),
            "anomaly": "true", "id": tad_id})
    if not rows:
        rows.append({
            "sourceIP": "None", "sourceTransportPort": 0, "destinationIP": "None",
            "destinationTransportPort": 0, "protocolIdentifier": 0,
            "flowStartSeconds": now_str, "podNamespace": "None", "podLabels": "None",
            "podName": "None", "destinationServicePortName": "None", "direction": "None",
            "flowEndSeconds": 0, "throughputStandardDeviation": 0,
            "aggType": agg_type, "algoType": algo_name, "algoCalc": 0.0,
            "throughput": 0.0, "anomaly": "NO ANOMALY DETECTED", "id": tad_id})
    return rows
