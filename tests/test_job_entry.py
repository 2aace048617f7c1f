"""The job's entry point (theia_b200.anomaly_detection.main) as a drop-in for ``python anomaly_detection.py <argv>``:
same options, ClickHouse over the HTTP port of --db_jdbc_url, results appended to default.tadetector."""
import http.server
import threading
import urllib.parse

import numpy as np

from theia_b200 import anomaly_detection as ad
from theia_b200 import clickhouse_native as chn

from .test_clickhouse_native import _RecordingEngine, _flows_as_native, _named_flows


class _FakeTransport:
    def __init__(self, stream: bytes):
        self.stream, self.selects, self.inserts = stream, [], []

    def select_native(self, sql):
        self.selects.append(sql)
        return self.stream

    def insert_native(self, table, block):
        self.inserts.append((table, block))


def test_main_reads_flows_and_appends_to_tadetector():
    flows = _named_flows(seed=21, series=8, points=6)
    tr, eng = _FakeTransport(_flows_as_native(flows, block_rows=20)), _RecordingEngine()
    rc = ad.main(["--algo", "EWMA", "--id", "abc", "--start_time", "2020-01-01 00:00:00", "--ns-ignore-list", '["kube-system"]'],
                 engine=eng, transport=tr)
    assert rc == 0
    assert tr.selects == ["SELECT flowEndSeconds, throughput, sourceIP, sourceTransportPort, destinationIP, "
                          "destinationTransportPort, protocolIdentifier, flowStartSeconds, sourcePodNamespace, "
                          "destinationPodNamespace FROM default.flows WHERE flowStartSeconds >= '2020-01-01 00:00:00'"]
    table, kw = eng.calls[0]
    assert kw["algo"] == "EWMA" and kw["tad_id"] == "abc" and kw["start_time"] > 0 and len(kw["ns_ignore"]) == 1
    assert np.array_equal(table["value"], np.asarray(flows["throughput"]))
    (name, body), = tr.inserts
    got = chn.read_native(body)
    assert name == "default.tadetector" and got["anomaly"].tolist() == ["NO ANOMALY DETECTED"] and got["id"].tolist() == ["abc"]
    assert got["algoType"].tolist() == ["EWMA"] and got["aggType"].tolist() == ["None"]


def test_main_generates_an_id_and_handles_an_empty_table():
    tr, eng = _FakeTransport(b""), _RecordingEngine()
    assert ad.main(["--algo", "DBSCAN", "--agg-flow", "svc"], engine=eng, transport=tr) == 0
    assert tr.selects == ["SELECT flowEndSeconds, throughput, destinationServicePortName FROM default.flows"]
    got = chn.read_native(tr.inserts[0][1])
    assert len(got["id"][0]) == 36 and got["aggType"].tolist() == ["svc"] and got["algoType"].tolist() == ["DBSCAN"]


def test_bad_options_exit_like_the_reference():
    assert ad.main(["--no-such-option"], engine=_RecordingEngine(), transport=_FakeTransport(b"")) == 2
    try:
        ad.main(["--algo", "KMEANS"], engine=_RecordingEngine(), transport=_FakeTransport(b""))
    except SystemExit as e:
        assert e.code == 2
    else:
        raise AssertionError("invalid --algo must exit 2")


def test_clickhouse_http_speaks_to_the_port_of_the_jdbc_url():
    seen = []

    class H(http.server.BaseHTTPRequestHandler):
        def do_POST(self):
            body = self.rfile.read(int(self.headers.get("Content-Length", 0)))
            q = urllib.parse.parse_qs(urllib.parse.urlparse(self.path).query)
            seen.append((q, self.headers.get("X-ClickHouse-User"), self.headers.get("X-ClickHouse-Key"), body))
            out = b"NATIVE" if q["query"][0].startswith("SELECT") else b""
            self.send_response(200)
            self.send_header("Content-Length", str(len(out)))
            self.end_headers()
            self.wfile.write(out)

        def log_message(self, *a):
            pass

    srv = http.server.HTTPServer(("127.0.0.1", 0), H)
    th = threading.Thread(target=srv.serve_forever, daemon=True)
    th.start()
    try:
        ch = ad.ClickHouseHTTP("jdbc:clickhouse://127.0.0.1:%d/default" % srv.server_address[1], user="u", password="p")
        assert ch.select_native("SELECT 1 ") == b"NATIVE"
        ch.insert_native("default.tadetector", b"\\x01\\x02")
    finally:
        srv.shutdown()
    (q1, u1, k1, b1), (q2, u2, k2, b2) = seen
    assert q1 == {"query": ["SELECT 1 FORMAT Native"], "database": ["default"], "wait_end_of_query": ["1"]} and (u1, k1, b1) == ("u", "p", b"")
    assert q2["query"] == ["INSERT INTO default.tadetector FORMAT Native"] and b2 == b"\\x01\\x02"


def test_ipv6_addresses_group_like_text_and_come_back_as_text():
    """Dual-stack tables: a column holding any non-IPv4 text is dictionary-encoded as a whole (ids and u32 addresses never
    share a key column) and decoded through the same dictionary on the way out."""
    n = 12
    flows = {"sourceIP": np.array(["fd00::1", "10.0.0.1", "fd00::1"] * 4, dtype=object),
             "destinationIP": np.array(["10.0.0.9"] * n, dtype=object),
             "sourceTransportPort": np.full(n, 1000, np.uint16), "destinationTransportPort": np.full(n, 80, np.uint16),
             "protocolIdentifier": np.full(n, 6, np.uint8), "flowStartSeconds": np.full(n, 100, np.uint32),
             "flowEndSeconds": np.arange(200, 200 + n, dtype=np.uint32), "throughput": np.arange(1, n + 1, dtype=np.uint64)}
    eng = _RecordingEngine()
    ad.anomaly_detection(eng, "EWMA", flows, tad_id="t")
    table, _ = eng.calls[0]
    assert table["src_ip"].tolist() == [0, 1, 0] * 4                      # dictionary ids, first appearance
    assert table["dst_ip"].tolist() == [(10 << 24) | 9] * n               # pure IPv4 column stays numeric
    # decode: what the engine returns for series id 0 / address 10.0.0.9
    plan = ad.plan_query()
    dicts = {slot: ad.Dictionary() for slot in ad.KEY_SLOTS}
    dicts["src_ip"].encode(flows["sourceIP"])
    got = {"src_ip": np.array([0, 1]), "dst_ip": np.array([(10 << 24) | 9] * 2), "src_port": np.array([1000, 1000]),
           "dst_port": np.array([80, 80]), "proto": np.array([6, 6]), "flow_start": np.array([100, 100]),
           "flow_end": np.array([200, 201]), "stddev": np.array([1.0, 1.0]), "algo_calc": np.array([2.0, 2.0]),
           "throughput": np.array([3.0, 3.0]), "anomaly": np.array([1, 1])}
    rows = ad._result_rows(got, plan, dicts, "EWMA", "t", None)
    assert [r["sourceIP"] for r in rows] == ["fd00::1", "10.0.0.1"] and rows[0]["destinationIP"] == "10.0.0.9"


def test_main_end_to_end_on_the_oracle_engine_matches_the_direct_call():
    """argv -> SELECT ... FORMAT Native -> decode -> plan / masks / dictionaries -> engine -> tadetector block, against the
    rows the in-memory call produces (engine double backed by the CPU oracle: no GPU, no ClickHouse)."""
    from . import test_host_mirror as thm
    from .test_host_mirror_on_oracle import OracleEngine
    fl = thm._flows(seed=4)
    types = dict(chn_types(fl))
    stream = chn.write_native([(k, types[k], np.asarray(v).astype(np.uint32) if types[k] == "String" and np.asarray(v).dtype.kind in "iu" else v)
                               for k, v in fl.items()])
    for argv, kw in ((["--algo", "EWMA", "--id", "j1"], {}),
                     (["--algo", "EWMA", "--id", "j2", "--agg-flow", "svc"], {"agg_flow": "svc"}),
                     (["--algo", "DBSCAN", "--id", "j3", "--agg-flow", "pod", "--pod-label", "web"], {"agg_flow": "pod", "pod_label": "web"})):
        tr = _FakeTransport(stream)
        assert ad.main(argv, engine=OracleEngine(), transport=tr) == 0
        want, _ = ad.anomaly_detection(OracleEngine(), argv[1], fl, tad_id=argv[3], **kw)
        got = chn.read_native(tr.inserts[0][1])
        assert len(want) == len(got["id"]) and len(want) > 1
        keyf = lambda r: (str(r.get("sourceIP", "")), str(r.get("podNamespace", "")), str(r.get("podLabels", "")), str(r.get("direction", "")),
                          str(r.get("destinationServicePortName", "")), int(r.get("sourceTransportPort", 0)), int(r["flowEndSeconds"]), float(r["algoCalc"]))
        rows_got = [{c: got[c][i] for c in got} for i in range(len(got["id"]))]
        assert sorted(map(keyf, want)) == sorted(map(keyf, rows_got))


def chn_types(fl):
    for k, v in fl.items():
        a = np.asarray(v)
        if k in ("sourceIP", "destinationIP") or a.dtype.kind in "OUS":
            yield k, "String"
        elif k in ("flowStartSeconds", "flowEndSeconds"):
            yield k, "DateTime"
        else:
            yield k, {1: "UInt8", 2: "UInt16", 4: "UInt32", 8: "UInt64"}[a.dtype.itemsize]


def test_ipv4_pushdown_selects_numeric_addresses_and_falls_back_to_text(monkeypatch):
    plan = ad.plan_query()
    sql = ad.raw_select_sql(plan, ipv4_pushdown=True)
    assert "IPv4StringToNum(sourceIP) AS sourceIP" in sql and "IPv4StringToNum(destinationIP) AS destinationIP" in sql
    flows = _named_flows(seed=2, series=4, points=5)
    numeric = chn.write_native([("sourceIP", "UInt32", flows["sourceIP"]), ("destinationIP", "UInt32", flows["destinationIP"])]
                               + [(k, t, flows[k]) for k, t in (("sourceTransportPort", "UInt16"), ("destinationTransportPort", "UInt16"),
                                  ("protocolIdentifier", "UInt8"), ("flowStartSeconds", "DateTime"), ("flowEndSeconds", "DateTime"),
                                  ("throughput", "UInt64"))])
    text = _flows_as_native(flows, block_rows=1000)

    class T(_FakeTransport):
        def select_native(self, q):
            self.selects.append(q)
            if "IPv4StringToNum" in q and self.fail_pushdown:
                raise RuntimeError("Code: 441. DB::Exception: Invalid IPv4 value")
            return numeric if "IPv4StringToNum" in q else text

    monkeypatch.setenv("TAD_IPV4_PUSHDOWN", "1")
    tables = []
    for fail in (False, True):
        tr, eng = T(b""), _RecordingEngine()
        tr.fail_pushdown = fail
        assert ad.main(["--algo", "EWMA", "--id", "p"], engine=eng, transport=tr) == 0
        assert len(tr.selects) == (2 if fail else 1)
        tables.append(eng.calls[0][0])
    for k in tables[0]:
        assert (tables[0][k] is None) == (tables[1][k] is None)
        if tables[0][k] is not None:
            assert np.array_equal(tables[0][k], tables[1][k]), k


def test_column_wise_insert_body_is_byte_identical_to_the_row_wise_one():
    from . import test_host_mirror as thm
    from .test_host_mirror_on_oracle import OracleEngine
    fl = thm._flows(seed=4)
    for kw in ({}, {"agg_flow": "svc"}, {"agg_flow": "external"}, {"agg_flow": "pod", "pod_label": "web"},
               {"agg_flow": "pod", "pod_name": "pod-3", "pod_namespace": "flow-visibility"}):
        got, _st, plan, dicts = ad.run_engine(OracleEngine(), "EWMA", fl, tad_id="T", **kw)
        rows = ad._result_rows(got, plan, dicts, "EWMA", "T", kw.get("pod_label"))
        assert len(rows) > 1
        assert chn.tadetector_block_from_result(got, plan, dicts, "EWMA", "T") == chn.tadetector_block(rows), kw


def test_main_svc_with_a_window_does_not_refilter_on_the_gpu():
    """with --agg-flow svc / external the window's lower bound is a WHERE condition on a column that is not
    part of the key; the SELECT pushes it down and the engine must not see start_time (its flow_start column is absent)."""
    from . import test_host_mirror as thm
    from .test_host_mirror_on_oracle import OracleEngine
    fl = thm._flows(seed=4)
    start = "2022-08-11 06:00:00"                     # every flowStartSeconds of the fixture is inside the window
    assert (fl["flowStartSeconds"] >= ad._epoch(start)).all()
    sel = {k: v for k, v in fl.items() if k in ("flowEndSeconds", "throughput", "destinationServicePortName")}
    types = dict(chn_types(sel))
    tr, eng = _FakeTransport(chn.write_native([(k, types[k], v) for k, v in sel.items()])), _RecordingEngine()
    assert ad.main(["--algo", "EWMA", "--id", "w", "--agg-flow", "svc", "--start_time", start], engine=eng, transport=tr) == 0
    assert tr.selects[0].endswith("WHERE flowStartSeconds >= '%s'" % start) and "flowStartSeconds" not in tr.selects[0].split(" FROM ")[0]
    table, kw = eng.calls[0]
    assert table["flow_start"] is None and kw["start_time"] == 0
    # and the rows: same as without the window (it keeps everything here)
    tr2 = _FakeTransport(tr.stream)
    assert ad.main(["--algo", "EWMA", "--id", "w", "--agg-flow", "svc", "--start_time", start], engine=OracleEngine(), transport=tr2) == 0
    want, _ = ad.anomaly_detection(OracleEngine(), "EWMA", fl, tad_id="w", agg_flow="svc")
    got = chn.read_native(tr2.inserts[0][1])
    assert len(got["id"]) == len(want) > 1


import pytest  # noqa: E402


@pytest.mark.gpu
def test_main_on_the_gpu_engine_matches_the_oracle_engine(engine):
    """main() with the REAL engine behind it (fake ClickHouse transport): argv -> SELECT ... FORMAT Native -> decode -> plan /
    masks / dictionaries -> C ABI -> CUDA kernels -> tadetector INSERT block, byte-compared column by column with the block
    the same argv produces on the oracle-backed engine double."""
    from . import test_host_mirror as thm
    from .test_host_mirror_on_oracle import OracleEngine
    fl = thm._flows(seed=4)
    types = dict(chn_types(fl))
    stream = chn.write_native([(k, types[k], np.asarray(v).astype(np.uint32) if types[k] == "String" and np.asarray(v).dtype.kind in "iu" else v)
                               for k, v in fl.items()])
    for argv in (["--algo", "EWMA", "--id", "g1", "--ns-ignore-list", '["kube-system"]'],
                 ["--algo", "EWMA", "--id", "g2", "--agg-flow", "svc", "--end_time", "2022-08-11 07:00:00"],
                 ["--algo", "DBSCAN", "--id", "g3", "--agg-flow", "pod", "--pod-label", "web"],
                 ["--algo", "EWMA", "--id", "g4", "--agg-flow", "external"]):
        blocks = []
        for eng in (engine, OracleEngine()):
            tr = _FakeTransport(stream)
            assert ad.main(argv, engine=eng, transport=tr) == 0
            (name, body), = tr.inserts
            assert name == "default.tadetector"
            blocks.append(chn.read_native(body))
        got, want = blocks
        assert set(got) == set(want) and len(got["id"]) == len(want["id"]) > 1, argv
        cols = [c for c in got if c != "timeCreated"] if "timeCreated" in got else list(got)
        order = lambda b: np.lexsort(tuple(np.asarray(b[c]).astype(str) for c in sorted(cols)))
        og, ow = order(got), order(want)
        for c in cols:
            assert np.array_equal(np.asarray(got[c])[og], np.asarray(want[c])[ow]), (argv, c)


def test_clickhouse_errors_surface_and_only_ipv4_failures_fall_back(monkeypatch):
    """a 200 response that carries X-ClickHouse-Exception-Code is an error, and the IPv4 push-down retries in
    text mode only when IPv4StringToNum itself failed (not on auth / network errors, which would double the scan)."""
    class H(http.server.BaseHTTPRequestHandler):
        def do_POST(self):
            self.rfile.read(int(self.headers.get("Content-Length", 0)))
            out = b"Code: 441. DB::Exception: Invalid IPv4 value"
            self.send_response(200)
            self.send_header("X-ClickHouse-Exception-Code", "441")
            self.send_header("Content-Length", str(len(out)))
            self.end_headers()
            self.wfile.write(out)

        def log_message(self, *a):
            pass

    srv = http.server.HTTPServer(("127.0.0.1", 0), H)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        ch = ad.ClickHouseHTTP("jdbc:clickhouse://127.0.0.1:%d" % srv.server_address[1])
        with pytest.raises(ad.ClickHouseError) as ei:
            ch.select_native("SELECT 1")
        assert ei.value.code == 441 and ad._is_ipv4_pushdown_failure(ei.value)
    finally:
        srv.shutdown()
    assert not ad._is_ipv4_pushdown_failure(ConnectionRefusedError("nope"))
    assert not ad._is_ipv4_pushdown_failure(ad.ClickHouseError(516, "Authentication failed"))

    class T(_FakeTransport):
        def select_native(self, q):
            self.selects.append(q)
            raise ad.ClickHouseError(516, "default: Authentication failed")

    monkeypatch.setenv("TAD_IPV4_PUSHDOWN", "1")
    tr = T(b"")
    with pytest.raises(ad.ClickHouseError):
        ad.main(["--algo", "EWMA", "--id", "p"], engine=_RecordingEngine(), transport=tr)
    assert len(tr.selects) == 1                       # no second scan


def test_null_stddev_goes_into_the_block_as_the_column_default():
    """a single-point series has stddev_samp NULL; throughputStandardDeviation is a non-nullable Float64, the
    reference's NULL lands there as 0."""
    plan = ad.plan_query()
    dicts = {slot: ad.Dictionary() for slot in ad.KEY_SLOTS}
    got = {"src_ip": np.array([1, 2], np.uint32), "dst_ip": np.array([3, 4], np.uint32), "src_port": np.array([5, 6], np.uint16),
           "dst_port": np.array([7, 8], np.uint16), "proto": np.array([6, 6], np.uint8), "flow_start": np.array([100, 100], np.uint32),
           "flow_end": np.array([200, 201], np.uint32), "stddev": np.array([np.nan, 2.5]), "algo_calc": np.array([0.0, 0.0]),
           "throughput": np.array([3.0, 3.0]), "anomaly": np.array([1, 1], np.uint8)}
    a = chn.read_native(chn.tadetector_block_from_result(got, plan, dicts, "DBSCAN", "t"))
    b = chn.read_native(chn.tadetector_block(ad._result_rows(got, plan, dicts, "DBSCAN", "t", None)))
    assert a["throughputStandardDeviation"].tolist() == [0.0, 2.5] == b["throughputStandardDeviation"].tolist()


def test_streamed_select_is_decoded_block_by_block_and_mid_stream_errors_surface():
    """The job entry reads `SELECT ... FORMAT Native` through ClickHouseHTTP.select_native_stream: pieces of the body are
    decoded into blocks as they arrive (same table as decoding the whole body), main() produces the same INSERT as with the
    buffered transport, and ClickHouse's `Code: N. DB::Exception` trailer of a broken stream is raised as ClickHouseError."""
    from . import test_host_mirror as thm
    from .test_host_mirror_on_oracle import OracleEngine
    fl = thm._flows(seed=4)
    types = dict(chn_types(fl))
    cols = [(k, types[k], np.asarray(v).astype(np.uint32) if types[k] == "String" and np.asarray(v).dtype.kind in "iu" else v)
            for k, v in fl.items()]
    n = len(fl["throughput"])
    stream = b"".join(chn.write_native([(k, t, v[lo:lo + 700]) for k, t, v in cols]) for lo in range(0, n, 700))
    state = {"fail": False}

    class H(http.server.BaseHTTPRequestHandler):
        def do_POST(self):
            self.rfile.read(int(self.headers.get("Content-Length", 0)))
            q = urllib.parse.parse_qs(urllib.parse.urlparse(self.path).query)["query"][0]
            if q.startswith("INSERT"):
                H.inserted = True
                self.send_response(200); self.send_header("Content-Length", "0"); self.end_headers()
                return
            assert "wait_end_of_query" not in self.path            # the streamed SELECT must not make the server buffer the result
            body = stream if not state["fail"] else stream[: len(stream) // 2] + b"Code: 241. DB::Exception: Memory limit (total) exceeded"
            self.send_response(200)
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            for i in range(0, len(body), 4096):                      # many small writes: blocks straddle the pieces
                self.wfile.write(body[i:i + 4096])

        def log_message(self, *a):
            pass

    srv = http.server.HTTPServer(("127.0.0.1", 0), H)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        ch = ad.ClickHouseHTTP("jdbc:clickhouse://127.0.0.1:%d" % srv.server_address[1])
        got = chn.flows_from_native(ch.select_native_stream("SELECT 1", piece=1500))
        want = chn.flows_from_native(stream)
        assert set(got) == set(want) and all(np.array_equal(np.asarray(got[k]), np.asarray(want[k])) for k in want)
        # main() over the streaming transport == main() over the buffered fake transport
        class Rec(ad.ClickHouseHTTP):
            def insert_native(self, table, block):
                self.block = block
        tr = Rec("jdbc:clickhouse://127.0.0.1:%d" % srv.server_address[1])
        assert ad.main(["--algo", "EWMA", "--id", "s1", "--agg-flow", "svc"], engine=OracleEngine(), transport=tr) == 0
        tr2 = _FakeTransport(stream)
        assert ad.main(["--algo", "EWMA", "--id", "s1", "--agg-flow", "svc"], engine=OracleEngine(), transport=tr2) == 0
        a, b = chn.read_native(tr.block), chn.read_native(tr2.inserts[0][1])
        assert len(a["id"]) == len(b["id"]) > 1
        key = lambda d: sorted(zip(d["destinationServicePortName"].tolist(), d["flowEndSeconds"].tolist(), d["algoCalc"].tolist()))
        assert key(a) == key(b)
        state["fail"] = True
        with pytest.raises(ad.ClickHouseError) as ei:
            chn.flows_from_native(ch.select_native_stream("SELECT 1", piece=1500))
        assert ei.value.code == 241 and "Memory limit" in ei.value.text
    finally:
        srv.shutdown()
