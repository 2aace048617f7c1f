"""Property test: the C port of the oracle equals the Python restatement on arbitrary small tables
(few distinct keys so that groups, duplicate (key, time) pairs and equal values all occur)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import c_oracle, tad_oracle as o

row = st.tuples(st.integers(0, 3), st.integers(0, 2), st.integers(0, 1), st.integers(0, 12),
                st.one_of(st.integers(0, 5), st.integers(10**9, 10**9 + 3), st.integers(2**63, 2**63 + 2)))


@settings(max_examples=60, deadline=None)
@given(st.lists(row, min_size=0, max_size=60), st.sampled_from([o.ALGO_EWMA, o.ALGO_DBSCAN]), st.booleans(), st.booleans())
def test_c_port_equals_python(rows, algo, emit_all, use_sum):
    n = len(rows)
    t = {
        "src_ip": np.array([r[0] for r in rows], dtype=np.uint32), "dst_ip": np.zeros(n, dtype=np.uint32),
        "src_port": np.array([r[1] for r in rows], dtype=np.uint16), "dst_port": np.zeros(n, dtype=np.uint16),
        "proto": np.array([r[2] for r in rows], dtype=np.uint8), "flow_start": np.full(n, 100, dtype=np.uint32),
        "flow_end": np.array([200 + r[3] for r in rows], dtype=np.uint32),
        "value": np.array([r[4] for r in rows], dtype=np.uint64),
    }
    reducer = o.REDUCE_SUM if use_sum else o.REDUCE_MAX
    pr = o.run_job(t, o.JobSpec(algo=algo, emit_all=emit_all, reducer=reducer))
    cc, ns, npts = c_oracle.run_job(t, algo=algo, emit_all=emit_all, reducer=reducer, threads=2)
    cc = o.canonicalize(cc)
    assert ns == pr.n_series and npts == pr.n_points
    for k in pr.cols:
        assert np.array_equal(pr.cols[k], cc[k], equal_nan=True), k
