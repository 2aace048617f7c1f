"""ClickHouse Native-format codec (theia_b200/clickhouse_native.py + the tad_ch_* helpers of the library).

No ClickHouse binary exists in this image, so the byte layout is pinned with hand-assembled blocks that follow the
published format description (n_columns, n_rows, then name / type / data per column)."""
import numpy as np
import pytest

from theia_b200 import clickhouse_native as chn
from theia_b200 import synth
from theia_b200.anomaly_detection import anomaly_detection, ip_to_u32, u32_to_ip


def _s(b: bytes) -> bytes:                      # a String value with a one-byte length
    assert len(b) < 128
    return bytes([len(b)]) + b


HAND_BLOCK = (
    b"\x03" b"\x03"                                                         # 3 columns, 3 rows
    + _s(b"n") + _s(b"UInt16") + b"\x01\x00\x02\x00\xff\xff"
    + _s(b"s") + _s(b"String") + _s(b"a") + _s(b"") + _s(b"10.0.0.1")
    + _s(b"t") + _s(b"DateTime") + (1700000000).to_bytes(4, "little") + (0).to_bytes(4, "little") + (2 ** 32 - 1).to_bytes(4, "little")
)


def test_hand_assembled_block_decodes():
    got = chn.read_native(HAND_BLOCK)
    assert list(got) == ["n", "s", "t"]
    assert got["n"].dtype == np.uint16 and got["n"].tolist() == [1, 2, 65535]
    assert got["s"].tolist() == ["a", "", "10.0.0.1"]
    assert got["t"].dtype == np.uint32 and got["t"].tolist() == [1700000000, 0, 2 ** 32 - 1]


def test_writer_reproduces_the_hand_assembled_bytes():
    out = chn.write_native([("n", "UInt16", [1, 2, 65535]), ("s", "String", ["a", "", "10.0.0.1"]),
                            ("t", "DateTime", [1700000000, 0, 2 ** 32 - 1])])
    assert out == HAND_BLOCK


def test_long_strings_use_multi_byte_varuint_and_blocks_concatenate():
    long = "x" * 200 + "é"                                   # 202 bytes: VarUInt 0xCA 0x01
    blk = chn.write_native([("s", "String", [long, "y"])])
    assert blk[:2] == b"\x01\x02" and b"\xca\x01" in blk
    got = chn.read_native(blk + blk)
    assert got["s"].tolist() == [long, "y", long, "y"]


@pytest.mark.parametrize("cut", [1, 5, 12, len(HAND_BLOCK) - 1])
def test_truncated_streams_raise(cut):
    with pytest.raises(ValueError):
        chn.read_native(HAND_BLOCK[:cut])


def test_unsupported_types_are_refused_not_misread():
    blk = b"\x01\x01" + _s(b"c") + _s(b"LowCardinality(String)") + b"\x00" * 32
    with pytest.raises(NotImplementedError):
        chn.read_native(blk)


def test_nullable_and_fixed_string():
    blk = (b"\x02\x02" + _s(b"v") + _s(b"Nullable(UInt8)") + b"\x00\x01" + b"\x07\x00"
           + _s(b"f") + _s(b"FixedString(3)") + b"abcxyz")
    got = chn.read_native(blk)
    assert got["v"].tolist() == [7, None]
    assert got["f"].tolist() == [b"abc", b"xyz"]


def test_ipv4_text_to_u32_and_back():
    texts = ["10.0.0.1", "255.255.255.255", "0.0.0.0", "1.2.3", "1.2.3.4.5", "256.1.1.1", "1..2.3", "::1", "", "01.2.3.4",
             "1.2.3.4 ", "a.b.c.d", "192.168.1.100"]
    blk = chn.write_native([("ip", "String", texts)])
    (cols, _types, rows), = list(chn.read_blocks(blk))
    ip, ok = cols["ip"].ipv4()
    want_ok = [True, True, True, False, False, False, False, False, False, False, False, False, True]
    assert ok.tolist() == want_ok
    for t, v, k in zip(texts, ip.tolist(), ok.tolist()):
        if k:
            assert v == int(ip_to_u32([t])[0]) and u32_to_ip(v) == t
        else:
            assert v == 0
    # u32 -> String column bytes == the text encoding of the same addresses
    good = [t for t, k in zip(texts, want_ok) if k]
    assert chn.write_native([("ip", "String", ip_to_u32(good))]) == chn.write_native([("ip", "String", good)])


def test_string_codes_are_first_appearance_dictionary_ids():
    blk = chn.write_native([("ns", "String", ["b", "a", "b", "", "a", "c"])])
    (cols, _t, _r), = list(chn.read_blocks(blk))
    ids, names = cols["ns"].codes()
    assert names == ["b", "a", "", "c"] and ids.tolist() == [0, 1, 0, 2, 1, 3]


FLOW_TYPES = {"sourceIP": "String", "destinationIP": "String", "sourceTransportPort": "UInt16",
              "destinationTransportPort": "UInt16", "protocolIdentifier": "UInt8", "flowStartSeconds": "DateTime",
              "flowEndSeconds": "DateTime", "throughput": "UInt64", "sourcePodNamespace": "String",
              "destinationPodNamespace": "String", "flowType": "UInt8"}


def _flows_as_native(flows: dict, block_rows: int) -> bytes:
    n = len(flows["flowEndSeconds"])
    out = []
    for lo in range(0, n, block_rows):
        hi = min(n, lo + block_rows)
        cols = []
        for k, typ in FLOW_TYPES.items():
            if k not in flows:
                continue
            v = np.asarray(flows[k])[lo:hi]
            if typ == "String" and v.dtype.kind in "iu":
                v = v.astype(np.uint32)                       # addresses: written as dotted quads
            cols.append((k, typ, v))
        out.append(chn.write_native(cols))
    return b"".join(out)


class _RecordingEngine:
    """Stands in for TadEngine: records the packed table the job hands to the GPU."""

    def __init__(self):
        self.calls = []

    def run(self, table, **kw):
        self.calls.append(({k: (None if v is None else np.asarray(v).copy()) for k, v in table.items()}, kw))
        z = np.zeros(0)
        return ({"src_ip": z, "dst_ip": z, "src_port": z, "dst_port": z, "proto": z, "flow_start": z, "flow_end": z,
                 "stddev": z, "algo_calc": z, "throughput": z, "anomaly": z}, None)


def _named_flows(seed=5, series=40, points=12):
    t = synth.make_flows(series, points, seed=seed)
    rng = np.random.default_rng(seed)
    n = len(t["value"])
    return {"sourceIP": t["src_ip"], "destinationIP": t["dst_ip"], "sourceTransportPort": t["src_port"],
            "destinationTransportPort": t["dst_port"], "protocolIdentifier": t["proto"], "flowStartSeconds": t["flow_start"],
            "flowEndSeconds": t["flow_end"], "throughput": t["value"],
            "sourcePodNamespace": rng.choice(np.array(["default", "kube-system", "antrea"], dtype=object), n),
            "destinationPodNamespace": rng.choice(np.array(["default", "flow-visibility"], dtype=object), n),
            "flowType": np.ones(n, dtype=np.uint8)}


def test_native_transport_hands_the_engine_the_same_table_as_the_dict_path():
    flows = _named_flows()
    stream = _flows_as_native(flows, block_rows=100)           # several blocks
    decoded = chn.flows_from_native(stream)
    assert decoded["sourceIP"].dtype == np.uint32              # parsed by the library, no per-row Python strings
    a, b = _RecordingEngine(), _RecordingEngine()
    kw = dict(start_time="", end_time="", tad_id="x", ns_ignore_list=["kube-system"])
    rows_a, _ = anomaly_detection(a, "EWMA", flows, **kw)
    rows_b, _ = anomaly_detection(b, "EWMA", decoded, **kw)
    (ta, kwa), (tb, kwb) = a.calls[0], b.calls[0]
    assert kwa == kwb and set(ta) == set(tb)
    for k in ta:
        assert (ta[k] is None) == (tb[k] is None)
        if ta[k] is not None:
            assert np.array_equal(ta[k], tb[k]), k
    assert rows_a[0]["anomaly"] == rows_b[0]["anomaly"] == "NO ANOMALY DETECTED"


def test_tadetector_block_round_trip_including_the_sentinel_row():
    flows = _named_flows(seed=9, series=3, points=5)
    eng = _RecordingEngine()
    sentinel, _ = anomaly_detection(eng, "DBSCAN", flows, tad_id="job-1")
    rows = [{"sourceIP": "10.1.2.3", "sourceTransportPort": 443, "destinationIP": "10.9.8.7", "destinationTransportPort": 80,
             "protocolIdentifier": 6, "flowStartSeconds": 1700000000, "flowEndSeconds": 1700000600,
             "throughputStandardDeviation": 12.5, "aggType": "None", "algoType": "EWMA", "algoCalc": 3.25,
             "throughput": 1e9, "anomaly": "true", "id": "job-1"}] + sentinel
    got = chn.read_native(chn.tadetector_block(rows))
    assert list(got) == [n for n, _ in chn.TADETECTOR_SCHEMA]
    assert got["sourceIP"].tolist() == ["10.1.2.3", "None"] and got["podName"].tolist() == ["", "None"]
    assert got["protocolIdentifier"].dtype == np.uint16 and got["protocolIdentifier"].tolist() == [6, 0]
    assert got["throughput"].tolist() == [1e9, 0.0] and got["algoCalc"].tolist() == [3.25, 0.0]
    assert got["anomaly"].tolist() == ["true", "NO ANOMALY DETECTED"] and got["id"].tolist() == ["job-1", "job-1"]
    assert got["flowEndSeconds"].tolist() == [1700000600, 0] and got["flowStartSeconds"][0] == 1700000000


def test_decode_rate_is_reported():
    """Not an assertion on speed -- prints the host-side decode rate of a 1e6-row flows block for the record."""
    import time
    t = synth.make_flows(10_000, 100, seed=2)
    flows = {"sourceIP": t["src_ip"], "destinationIP": t["dst_ip"], "sourceTransportPort": t["src_port"],
             "destinationTransportPort": t["dst_port"], "protocolIdentifier": t["proto"], "flowStartSeconds": t["flow_start"],
             "flowEndSeconds": t["flow_end"], "throughput": t["value"]}
    stream = _flows_as_native(flows, block_rows=65_536)
    t0 = time.perf_counter()
    d = chn.flows_from_native(stream)
    dt = time.perf_counter() - t0
    assert np.array_equal(d["sourceIP"], t["src_ip"]) and np.array_equal(d["throughput"], t["value"])
    print("native decode: %.1f M rows/s, %.0f MB/s" % (len(t["value"]) / dt / 1e6, len(stream) / dt / 1e6))


def test_mutated_streams_never_crash_the_decoder():
    """Bit flips and truncations of a valid stream either decode or raise ValueError / NotImplementedError -- the C index
    walks attacker-controlled lengths, so it must never read outside the buffer."""
    from hypothesis import given, settings, strategies as st
    flows = _named_flows(seed=3, series=6, points=4)
    good = _flows_as_native(flows, block_rows=10)

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.tuples(st.integers(0, len(good) - 1), st.integers(0, 255)), min_size=1, max_size=6),
           st.integers(0, len(good)))
    def run(muts, cut):
        b = bytearray(good)
        for pos, val in muts:
            b[pos] = val
        data = bytes(b[:cut]) if cut else bytes(b)
        try:
            chn.flows_from_native(data)
        except (ValueError, NotImplementedError):
            pass

    run()


def test_dictionary_matches_a_python_dict_on_many_values():
    rng = np.random.default_rng(11)
    vocab = ["ns-%d" % i for i in range(5000)] + ["", "x" * 300, "é"]
    vals = [vocab[i] for i in rng.integers(0, len(vocab), 60000)]
    (cols, _t, _r), = list(chn.read_blocks(chn.write_native([("ns", "String", vals)])))
    ids, names = cols["ns"].codes()
    want, order = {}, []
    for v in vals:
        if v not in want:
            want[v] = len(order)
            order.append(v)
    assert names == order and ids.tolist() == [want[v] for v in vals]


def test_stream_reader_equals_whole_body_reader_for_any_split():
    """read_blocks_stream decodes a Native stream that arrives in arbitrary pieces (down to single bytes) into exactly the table
    the whole-body reader gives, and a stream that ends inside a block is an error."""
    import random
    flows = _named_flows(seed=3, series=30, points=25)
    stream = _flows_as_native(flows, block_rows=97)
    whole = chn.flows_from_native(stream)
    rng = random.Random(1)
    for _ in range(20):
        cuts = sorted(rng.sample(range(1, len(stream)), rng.randint(1, 30)))
        pieces = [stream[a:b] for a, b in zip([0] + cuts, cuts + [len(stream)])]
        got = chn.flows_from_native(iter(pieces))
        assert set(got) == set(whole)
        for k in whole:
            assert np.array_equal(np.asarray(got[k]), np.asarray(whole[k])), k
    got = chn.flows_from_native(stream[i:i + 1] for i in range(len(stream)))
    assert all(np.array_equal(np.asarray(got[k]), np.asarray(whole[k])) for k in whole)
    assert chn.flows_from_native(iter([])) == {}
    import pytest
    with pytest.raises(ValueError):
        chn.flows_from_native(iter([stream[:-5]]))
