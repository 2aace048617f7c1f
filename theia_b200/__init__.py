"""theia_b200 -- B200-native throughput-anomaly-detection engine behind Theia's TAD job API.

    theia_b200.engine             ctypes host over the C ABI (include/theia_tad.h)
    theia_b200.anomaly_detection  mirror of plugins/anomaly-detection/anomaly_detection.py's interface
    theia_b200.controller         double of pkg/controller/anomalydetector's state machine and validation
    theia_b200.clickhouse_native  ClickHouse Native-format column blocks <-> the engine's columns / tadetector rows
    theia_b200.sharding, .synth   ownership mirror / synthetic flow tables
    theia_b200.build              in-tree nvcc build of libtheia_tad.so (sm_100a)

There is no CPU fallback: without the CUDA library or a GPU the engine refuses to start.
"""
__version__ = "0.1.0"
