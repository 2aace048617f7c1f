"""Import the UNMODIFIED reference job module in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/tad_oracle.py).  ``/root/reference`` exists
only in the build container, never on the GPU box, so this loader is used by
``tests/golden/make_golden.py`` (fixture generation) and by CPU tests that skip
when the reference tree is absent.  ``pyspark`` and ``statsmodels`` are not
installed in the image; they are stubbed in ``sys.modules`` so that
``plugins/anomaly-detection/anomaly_detection.py`` imports and its pure-Python
UDFs (calculate_ewma*, calculate_dbscan*, generate_tad_sql_query) run on the
installed numpy / scipy / scikit-learn.  ARIMA cannot run (statsmodels absent).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("THEIA_REFERENCE_ROOT", "/root/reference")
_JOB_DIR = os.path.join(REFERENCE_ROOT, "plugins", "anomaly-detection")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_JOB_DIR, "anomaly_detection.py"))


class _Stub:
    def __init__(self, *a, **k):
        pass


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_job():
    """Return the reference ``anomaly_detection`` module (UDFs usable, Spark not)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _stub("pyspark")
    _stub("pyspark.sql", SparkSession=_Stub)
    _stub("pyspark.sql.functions")
    _stub("pyspark.sql.types", **{n: _Stub for n in (
        "BooleanType ArrayType StructField DecimalType DoubleType StringType "
        "LongType TimestampType StructType").split()})
    _stub("statsmodels")
    _stub("statsmodels.tsa")
    _stub("statsmodels.tsa.arima")
    _stub("statsmodels.tsa.arima.model", ARIMA=_Stub)
    if _JOB_DIR not in sys.path:
        sys.path.insert(0, _JOB_DIR)
    import anomaly_detection  # noqa: E402  (the reference module)
    return anomaly_detection


def load_reference_test_vectors():
    """Golden vectors held by the reference's own unit test
    (plugins/anomaly-detection/anomaly_detection_test.py:199-402)."""
    load_reference_job()
    import importlib
    mod = importlib.import_module("anomaly_detection_test")
    return {
        "throughput_list": list(mod.throughput_list),
        "expected_ewma_row_list": list(mod.expected_ewma_row_list),
        "expected_arima_row_list_5digits": list(mod.expected_arima_row_list),
        "expanded_arima_row_list": list(mod.expanded_arima_row_list),
        "stddev": float(mod.stddev),
        "expected_anomaly_list_arima": [bool(b) for b in mod.expected_anomaly_list_arima],
        "expected_anomaly_list_ewma": [bool(b) for b in mod.expected_anomaly_list_ewma],
        "expected_dbscan_anomaly_list": [bool(b) for b in mod.expected_dbscan_anomaly_list],
    }
