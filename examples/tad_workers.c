/* Four host threads on ONE tad_ctx -- the controller's worker pool (pkg/controller/util.go:43: DefaultWorkers = 4; every
 * worker runs syncTADetector and would call tad_submit / tad_poll / tad_result / tad_release through the cgo shim).
 * Each thread owns a table of its own, submits it ROUNDS times, polls it to completion (tad_poll, never tad_wait: the
 * controller only ever polls), reads the rows and releases the job, while the other three do the same; one thread
 * also cancels a job now and then.  Every result must equal the one the same table gave when it ran alone.
 *
 *   gcc -std=c99 -pthread -Iinclude examples/tad_workers.c -Ltheia_b200 -ltheia_tad -Wl,-rpath,$PWD/theia_b200 -o tad_workers
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include "theia_tad.h"

enum { WORKERS = 4, ROUNDS = 6 };

typedef struct {
    tad_ctx *ctx;
    int id;
    tad_columns cols;
    uint64_t want_rows, want_sum;     /* from the solo run */
    int algo;
    int failures;
    int cancelled_seen;
} worker;

static uint64_t mix(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

/* order-independent digest of the result rows (the engine emits them in no particular order) */
static uint64_t digest(const tad_rows *r)
{
    uint64_t s = 0;
    for (uint64_t i = 0; i < r->rows; i++) {
        uint64_t c, d;
        memcpy(&c, &r->algo_calc[i], 8);
        memcpy(&d, &r->stddev[i], 8);
        s += mix(((uint64_t)r->src_ip[i] << 32 | r->flow_end[i]) ^ mix(c ^ mix(d + r->src_port[i])));
    }
    return s;
}

static void fill(worker *w, uint64_t series, uint64_t points)
{
    tad_columns *c = &w->cols;
    uint64_t n = 0;
    for (uint64_t s = 0; s < series; s++) {
        const uint64_t base = 1000000ull + mix(s * 7919 + (uint64_t)w->id) % 900000000ull;
        for (uint64_t k = 0; k < points; k++, n++) {
            const uint64_t j = (n * 2654435761ull + (uint64_t)w->id) % (series * points);      /* scrambled row order */
            (void)j;
            c->src_ip[n] = 0x0A000000u + (uint32_t)(w->id << 20) + (uint32_t)s;
            c->dst_ip[n] = 0x0A640000u + (uint32_t)(s % 251);
            c->src_port[n] = (uint16_t)(1024 + s % 60000);
            c->dst_port[n] = 443;
            c->proto[n] = 6;
            c->flow_start[n] = 1660199214u;
            c->flow_end[n] = 1660202814u + 60u * (uint32_t)k;
            c->value[n] = base + mix(n + 17 * (uint64_t)w->id) % (base / 500 + 1) * ((mix(n) % 97 == 0) ? 40 : 1);
        }
    }
    c->rows = n;
}

static int run_once(worker *w, int cancel, uint64_t *rows_out, uint64_t *sum_out)
{
    tad_job_spec spec;
    memset(&spec, 0, sizeof(spec));
    spec.algo = w->algo;
    spec.reducer = TAD_REDUCE_MAX;
    snprintf(spec.id, sizeof(spec.id), "5ca1ab1e-0000-4000-8000-%012d", w->id);
    tad_job *job = NULL;
    int rc = tad_submit(w->ctx, &spec, &w->cols, &job);
    if (rc != TAD_OK) return rc;
    if (cancel) tad_cancel(job);
    tad_status st;
    struct timespec nap = {0, 200000};          /* 0.2 ms between polls */
    int last = -1;
    for (;;) {
        tad_poll(job, &st);
        if (st.completed_stages < last) { fprintf(stderr, "worker %d: progress went backwards\n", w->id); w->failures++; }
        last = st.completed_stages;
        if (st.state == TAD_STATE_COMPLETED || st.state == TAD_STATE_FAILED) break;
        nanosleep(&nap, NULL);
    }
    if (st.state == TAD_STATE_FAILED) {
        rc = st.error;
        if (cancel && rc == TAD_ERR_CANCELLED) w->cancelled_seen++;
        tad_release(job);
        return rc;
    }
    tad_rows r;
    rc = tad_result(job, &r);
    if (rc == TAD_OK) { *rows_out = r.rows; *sum_out = digest(&r); }
    tad_release(job);
    return rc;
}

static void *worker_main(void *arg)
{
    worker *w = (worker *)arg;
    for (int round = 0; round < ROUNDS; round++) {
        const int cancel = w->id == 3 && (round & 1);          /* worker 3 cancels every other job right after submit */
        uint64_t rows = 0, sum = 0;
        int rc = run_once(w, cancel, &rows, &sum);
        if (cancel && rc == TAD_ERR_CANCELLED) continue;        /* a cancelled job fails with TAD_ERR_CANCELLED ...      */
        if (rc != TAD_OK) { fprintf(stderr, "worker %d round %d: %s\n", w->id, round, tad_strerror(rc)); w->failures++; continue; }
        if (rows != w->want_rows || sum != w->want_sum) {       /* ... or had already finished: then it must be right    */
            fprintf(stderr, "worker %d round %d: %llu rows / digest %016llx, alone it gave %llu / %016llx\n", w->id, round,
                    (unsigned long long)rows, (unsigned long long)sum, (unsigned long long)w->want_rows, (unsigned long long)w->want_sum);
            w->failures++;
        }
    }
    return NULL;
}

int main(void)
{
    tad_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.world_size = 1;
    tad_ctx *ctx = NULL;
    int rc = tad_init(&cfg, &ctx);
    if (rc != TAD_OK) { fprintf(stderr, "tad_init: %s\n", tad_strerror(rc)); return 1; }
    static worker w[WORKERS];
    const uint64_t series[WORKERS] = {3000, 500, 12000, 40}, points[WORKERS] = {50, 400, 16, 5000};   /* 40 x 5000: spill path */
    for (int i = 0; i < WORKERS; i++) {
        w[i].ctx = ctx; w[i].id = i; w[i].algo = i == 1 ? TAD_ALGO_DBSCAN : TAD_ALGO_EWMA;
        if (tad_alloc_columns(ctx, series[i] * points[i], TAD_MEM_HOST, &w[i].cols) != TAD_OK) return 1;
        fill(&w[i], series[i], points[i]);
        rc = run_once(&w[i], 0, &w[i].want_rows, &w[i].want_sum);             /* alone */
        if (rc != TAD_OK) { fprintf(stderr, "solo run %d: %s\n", i, tad_strerror(rc)); return 1; }
    }
    pthread_t th[WORKERS];
    for (int i = 0; i < WORKERS; i++) pthread_create(&th[i], NULL, worker_main, &w[i]);
    int failures = 0, cancelled = 0;
    for (int i = 0; i < WORKERS; i++) { pthread_join(th[i], NULL); failures += w[i].failures; cancelled += w[i].cancelled_seen; }
    for (int i = 0; i < WORKERS; i++) {
        printf("worker %d: %llu rows in, %llu anomalous rows, digest %016llx\n", i, (unsigned long long)w[i].cols.rows,
               (unsigned long long)w[i].want_rows, (unsigned long long)w[i].want_sum);
        tad_free_columns(ctx, &w[i].cols);
    }
    tad_shutdown(ctx);
    printf("%d workers x %d rounds on one context: %d failures, %d jobs cancelled\n", WORKERS, ROUNDS, failures, cancelled);
    return failures ? 1 : 0;
}
