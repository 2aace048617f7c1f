/*
 * theia_tad.h -- C ABI of the B200-native throughput-anomaly-detection (TAD) engine.
 *
 * This is the drop-in boundary for ONE path of antrea-io/theia: the job that
 * pkg/controller/anomalydetector/controller.go launches as a SparkApplication
 * (startSparkApplication, controller.go:525-698) and that
 * plugins/anomaly-detection/anomaly_detection.py executes.  The reference has no FFI for
 * this path (its boundary is "create a SparkApplication CR, poll it, read ClickHouse"),
 * so every entry point below cites the reference function it replaces; INTEGRATION.md
 * shows the cgo binding a maintainer adds behind the controller's function-variable seam
 * (controller.go:54-59).
 *
 * Conventions: plain C, no callbacks into the host language, no exceptions across the
 * boundary; every function returns 0 (TAD_OK) or a negative tad_error; strings are
 * NUL-terminated UTF-8.  A tad_ctx may be used from several threads (the controller
 * runs 4 workers, pkg/controller/util.go:43); a tad_job belongs to the thread that polls
 * it.  All pinned-host and device memory is owned by the library (Go's GC may move Go
 * memory, so the host shim fills library-owned buffers obtained from tad_alloc_columns).
 */
#ifndef THEIA_TAD_H
#define THEIA_TAD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAD_ABI_VERSION 2   /* 2: tad_job_spec.global_rows, TAD_PHASE_SYNC */

typedef enum {
    TAD_OK = 0,
    TAD_ERR_INVALID_ARG = -1,  /* controller.go:527-622 "invalid request: ..." class      */
    TAD_ERR_CUDA = -2,
    TAD_ERR_NOMEM = -3,
    TAD_ERR_NCCL = -4,
    TAD_ERR_STATE = -5,        /* call not legal in the job's current state                */
    TAD_ERR_CANCELLED = -6,
    TAD_ERR_UNSUPPORTED = -7,
    TAD_ERR_INTERNAL = -8
} tad_error;

/* anomaly_detection.py:811 valid_algos; controller.go:527 */
typedef enum { TAD_ALGO_EWMA = 0, TAD_ALGO_ARIMA = 1, TAD_ALGO_DBSCAN = 2 } tad_algo;

/* Stage-A reducer over duplicates of (key, flowEndSeconds): max(throughput) for the
 * per-connection query (anomaly_detection.py:52-61), sum(throughput) for the pod /
 * external / svc aggregated-flow queries (:63-106). */
typedef enum { TAD_REDUCE_MAX = 0, TAD_REDUCE_SUM = 1 } tad_reducer;

/* Job states = ThroughputAnomalyDetectorStatus.State, pkg/apis/crd/v1alpha1/types.go:33-37 */
typedef enum {
    TAD_STATE_NEW = 0, TAD_STATE_SCHEDULED = 1, TAD_STATE_RUNNING = 2,
    TAD_STATE_COMPLETED = 3, TAD_STATE_FAILED = 4
} tad_state;

typedef enum { TAD_MEM_HOST = 0, TAD_MEM_DEVICE = 1 } tad_mem;

/* tad_job_spec.flags */
#define TAD_FLAG_EMIT_ALL 1u   /* emit every point with its flag (parity/debug), not only anomalies */
#define TAD_FLAG_PROFILE  2u   /* record per-phase CUDA-event times into tad_status.phase_ms         */

typedef struct tad_ctx tad_ctx;
typedef struct tad_job tad_job;

typedef struct {
    int32_t device;            /* CUDA device ordinal of this process/rank                          */
    int32_t world_size;        /* 1 = single GPU; N = this process is one of N ranks on one host    */
    int32_t rank;
    uint32_t flags;            /* reserved, 0                                                       */
    const void *nccl_unique_id;/* world_size > 1: the 128-byte ncclUniqueId shared by all ranks      */
    size_t nccl_unique_id_bytes;
} tad_config;

/*
 * Columnar flow records: the columns the TAD query selects from the ClickHouse `flows`
 * table (create_table.sh:31-85; anomaly_detection.py:52-61).  IPs are IPv4 packed in host
 * byte order, or host-assigned dictionary ids for anything else (the aggregated-flow
 * modes put their dictionary ids -- podNamespace/podLabels/podName/direction/
 * destinationServicePortName/flowType -- into these same six key slots).  A NULL key
 * column means "all zero".  src_ns/dst_ns are optional namespace dictionary ids used only
 * by the ns_ignore filter (anomaly_detection.py:576-580).
 */
typedef struct {
    uint64_t rows;
    uint64_t capacity;
    int32_t mem;               /* tad_mem: where the column pointers live                           */
    int32_t reserved;
    uint32_t *src_ip;          /* sourceIP                                                          */
    uint32_t *dst_ip;          /* destinationIP                                                     */
    uint16_t *src_port;        /* sourceTransportPort                                               */
    uint16_t *dst_port;        /* destinationTransportPort                                          */
    uint8_t *proto;            /* protocolIdentifier                                                */
    uint32_t *flow_start;      /* flowStartSeconds (DateTime, seconds)                              */
    uint32_t *flow_end;        /* flowEndSeconds                                                    */
    uint64_t *value;           /* throughput (UInt64)                                               */
    uint32_t *src_ns;          /* optional                                                          */
    uint32_t *dst_ns;          /* optional                                                          */
} tad_columns;

/* Job description = the argv the controller builds (controller.go:530-623) after its own
 * validation, with strings already mapped to ids by the host shim. */
typedef struct {
    int32_t algo;              /* tad_algo         --algo                                           */
    int32_t reducer;           /* tad_reducer      implied by --agg-flow                            */
    uint32_t start_time;       /* --start_time as epoch seconds, 0 = unbounded: flow_start >= start */
    uint32_t end_time;         /* --end_time, 0 = unbounded: flow_end < end                         */
    uint32_t flags;            /* TAD_FLAG_*                                                        */
    uint32_t n_ns_ignore;      /* --ns-ignore-list mapped to namespace ids                          */
    const uint32_t *ns_ignore;
    char id[40];               /* --id (uuid, 36 chars + NUL)                                       */
    uint64_t global_rows;      /* world_size > 1: rows of the whole table over all ranks (the host  */
                               /* knows it from its SELECT count()); every rank must pass the same   */
                               /* value.  0 = the ranks agree on it with a blocking all-gather at    */
                               /* job start.  Ignored on a single GPU.                               */
} tad_job_spec;

enum {
    TAD_PHASE_H2D = 0,         /* host -> device copy of the columns (host-resident input only)     */
    TAD_PHASE_HIST,            /* key pack + hash + bucket histogram                                */
    TAD_PHASE_SCAN,            /* bucket offsets                                                    */
    TAD_PHASE_SCATTER,         /* hash partition into 32-byte packed rows                           */
    TAD_PHASE_EXCHANGE,        /* multi-GPU: exposed part of the NCCL all-to-all (exact partition), or the  */
                               /* peer counter gather (optimistic partition: rows are pulled inside GROUP)  */
    TAD_PHASE_GROUP,           /* per-bucket group + time sort + reduce -> per-series arrays        */
    TAD_PHASE_SPILL,           /* oversized buckets through the global-memory path                  */
    TAD_PHASE_DETECT,          /* stddev_samp + EWMA/ARIMA/DBSCAN + anomaly compaction              */
    TAD_PHASE_D2H,             /* result rows device -> host                                        */
    TAD_PHASE_SYNC,            /* multi-GPU: time spent waiting for the other ranks at the job's     */
                               /* barriers (arrival skew + barrier latency; 0 on one GPU)            */
    TAD_NPHASES
};

/* = ThroughputAnomalyDetectorStatus (types.go:114-122) + counters. */
typedef struct {
    int32_t state;             /* tad_state                                                         */
    int32_t completed_stages;  /* maps onto Status.CompletedStages (controller.go:426-453)          */
    int32_t total_stages;      /* maps onto Status.TotalStages                                      */
    int32_t error;             /* tad_error when state == FAILED                                    */
    char err_msg[256];         /* Status.ErrorMsg                                                   */
    uint64_t rows_in;          /* rows handed in (this rank)                                        */
    uint64_t rows_kept;        /* after the stage-A filters (this rank, before exchange)            */
    uint64_t rows_owned;       /* rows this rank owns after the exchange                            */
    uint64_t points;           /* after the stage-A reduce                                          */
    uint64_t series;           /* distinct keys owned by this rank                                  */
    uint64_t result_rows;      /* anomalous points (all points with TAD_FLAG_EMIT_ALL)              */
    uint64_t spill_rows;       /* rows that went through the oversized-bucket path                  */
    uint64_t gpu_launches;     /* kernels of this library launched for the job                      */
    double device_ms;          /* first kernel start -> last kernel end (CUDA events)               */
    double total_ms;           /* submit -> results on host (wall clock)                            */
    double phase_ms[TAD_NPHASES];
} tad_status;

/* Result rows = the columns the reference appends to default.tadetector
 * (anomaly_detection.py:385-393; create_table.sh:363-384), structure-of-arrays, host
 * memory owned by the job, valid until tad_release().  aggType / algoType / id /
 * anomaly="true" are per-job constants the host shim adds, as is the
 * "NO ANOMALY DETECTED" sentinel row when rows == 0 (anomaly_detection.py:395-420). */
typedef struct {
    uint64_t rows;
    const uint32_t *src_ip;
    const uint32_t *dst_ip;
    const uint16_t *src_port;
    const uint16_t *dst_port;
    const uint8_t *proto;
    const uint32_t *flow_start;
    const uint32_t *flow_end;
    const double *stddev;      /* throughputStandardDeviation; NaN = SQL NULL (series of 1 point)   */
    const double *algo_calc;   /* algoCalc                                                          */
    const double *throughput;  /* throughput as Float64                                             */
    const uint8_t *anomaly;    /* 1 = anomalous (always 1 without TAD_FLAG_EMIT_ALL)                */
} tad_rows;

/* Replaces: cluster validation + Spark session bring-up (controller.go:499-503;
 * anomaly_detection.py:651).  Creates streams, NCCL communicator (world_size > 1). */
int tad_init(const tad_config *cfg, tad_ctx **out);
void tad_shutdown(tad_ctx *ctx);

/* Replaces: the JDBC read target (anomaly_detection.py:655-662).  Allocates library-owned
 * pinned host (mem = TAD_MEM_HOST) or device (TAD_MEM_DEVICE) column buffers. */
int tad_alloc_columns(tad_ctx *ctx, uint64_t capacity, int32_t mem, tad_columns *cols);
/* Adds the two optional namespace-id columns (src_ns, dst_ns) to buffers obtained from
 * tad_alloc_columns; needed only when the job carries an ns_ignore list (anomaly_detection.py:576-580). */
int tad_alloc_ns_columns(tad_ctx *ctx, tad_columns *cols);
int tad_free_columns(tad_ctx *ctx, tad_columns *cols);

/* Replaces: CreateSparkApplication (controller.go:685, pkg/controller/util.go:223-233).
 * Non-blocking: validates, enqueues, returns a job handle in state SCHEDULED. */
int tad_submit(tad_ctx *ctx, const tad_job_spec *spec, const tad_columns *cols, tad_job **out);

/* Replaces: GetSparkApplication + Spark-UI stage scraping (controller.go:426-497). */
int tad_poll(tad_job *job, tad_status *status);
/* Convenience: block until COMPLETED/FAILED or timeout_ms (<0 = forever) elapses. */
int tad_wait(tad_job *job, int64_t timeout_ms, tad_status *status);

/* Replaces: SELECT ... FROM tadetector WHERE id = ? (rest.go:249-315) for a finished job. */
int tad_result(tad_job *job, tad_rows *rows);

/* Replaces: DeleteSparkApplication (controller.go:385-424). */
int tad_cancel(tad_job *job);
int tad_release(tad_job *job);

const char *tad_strerror(int err);
int tad_abi_version(void);

/* Multi-GPU bring-up: rank 0 obtains a 128-byte ncclUniqueId here, the host distributes it
 * (the Go shim over its own channel, bench.py over torch.distributed) and every rank passes
 * it to tad_init.  Replaces: Spark's driver/executor rendezvous (no reference call site). */
int tad_get_unique_id(void *out, size_t bytes);

/* ---- ClickHouse Native-format helpers (host only: no GPU, no context) -----------------------------------------
 * The reference job moves rows over JDBC (read: anomaly_detection.py:655-662; write: :713-726).  The shim streams
 * `SELECT ... FORMAT Native` column blocks instead (clickhouse-go, go.mod:7) and copies every fixed-width column
 * (UInt8/16/64, DateTime = UInt32; create_table.sh:31-85) straight into the tad_columns buffers.  Only String
 * columns need work: these calls index them, turn IPv4 text into the u32 key column and back, and dictionary-encode the rest.
 * theia_b200/clickhouse_native.py is the block reader / writer built on them. */

/* A String column of `rows` values is rows x (VarUInt length, bytes).  Writes the payload position and length of
 * every value and the number of bytes consumed.  TAD_ERR_INVALID_ARG if the buffer ends early or a length >= 2^32. */
int tad_ch_string_index(const uint8_t *buf, size_t len, uint64_t rows, uint64_t *offsets, uint32_t *lengths,
                        size_t *consumed);
/* "a.b.c.d" -> a<<24 | b<<16 | c<<8 | d.  is_v4[i] = 0 (and out[i] = 0) for anything else (IPv6, empty, names):
 * the caller gives those rows dictionary ids.  sourceIP / destinationIP are String columns (create_table.sh:38-39). */
int tad_ch_parse_ipv4(const uint8_t *buf, const uint64_t *offsets, const uint32_t *lengths, uint64_t rows,
                      uint32_t *out, uint8_t *is_v4);
/* u32 -> String column bytes (VarUInt length + "a.b.c.d" per row) for the tadetector INSERT (create_table.sh:363-384).
 * Needs at most 16 bytes per row; TAD_ERR_INVALID_ARG if out_cap is too small. */
int tad_ch_format_ipv4(const uint32_t *ips, uint64_t rows, uint8_t *out, size_t out_cap, size_t *written);

/* Dense dictionary ids (order of first appearance) for a String column: ids[i] in [0, *n_unique); first_row[k] is the
 * row at which id k first occurs (the caller reads the k-th name from there).  first_row needs room for `rows` entries.
 * This is how pod namespaces / labels / service-port names become the u32 key columns of the aggregated-flow modes
 * (anomaly_detection.py:511-609) and the namespace ids of the ignore list (:576-580). */
int tad_ch_dictionary(const uint8_t *buf, const uint64_t *offsets, const uint32_t *lengths, uint64_t rows, uint32_t *ids,
                      uint64_t *first_row, uint32_t *n_unique);

#ifdef __cplusplus
}
#endif
#endif /* THEIA_TAD_H */
