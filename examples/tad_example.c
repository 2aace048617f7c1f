/* Minimal C host of the TAD engine: the same call sequence the cgo shim of INTEGRATION.md makes.
 *
 *   gcc -std=c99 -Iinclude examples/tad_example.c -Ltheia_b200 -ltheia_tad -Wl,-rpath,$PWD/theia_b200 -o tad_example
 *   ./tad_example            # needs a B200; prints the anomalous points of one synthetic connection
 */
#include <stdio.h>
#include <string.h>

#include "theia_tad.h"

int main(void)
{
    tad_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.world_size = 1;
    tad_ctx *ctx = NULL;
    int rc = tad_init(&cfg, &ctx);
    if (rc != TAD_OK) {
        fprintf(stderr, "tad_init: %s\n", tad_strerror(rc));
        return 1;
    }
    enum { N = 64 };
    tad_columns cols;
    rc = tad_alloc_columns(ctx, N, TAD_MEM_HOST, &cols);
    if (rc != TAD_OK) return 1;
    for (int i = 0; i < N; i++) {                       /* one connection, one outlier */
        cols.src_ip[i] = 0x0A0A0119u; cols.dst_ip[i] = 0x0A0A0121u;
        cols.src_port[i] = 58076; cols.dst_port[i] = 5201; cols.proto[i] = 6;
        cols.flow_start[i] = 1660199214u; cols.flow_end[i] = 1660202814u + 60u * (uint32_t)i;
        cols.value[i] = i == 40 ? 50007861276ull : 4005000000ull + (uint64_t)(i * 977 % 4001);
    }
    cols.rows = N;
    tad_job_spec spec;
    memset(&spec, 0, sizeof(spec));
    spec.algo = TAD_ALGO_EWMA;
    spec.reducer = TAD_REDUCE_MAX;
    strncpy(spec.id, "5ca1ab1e-0000-4000-8000-000000000001", sizeof(spec.id) - 1);
    tad_job *job = NULL;
    rc = tad_submit(ctx, &spec, &cols, &job);
    tad_status st;
    if (rc == TAD_OK) rc = tad_wait(job, -1, &st);
    if (rc != TAD_OK || st.state != TAD_STATE_COMPLETED) {
        tad_poll(job, &st);
        fprintf(stderr, "job failed: %s\n", st.err_msg);
        return 1;
    }
    tad_rows rows;
    tad_result(job, &rows);
    printf("%llu series, %llu points, %llu anomalous rows (%d/%d stages, %.3f ms on the device)\n",
           (unsigned long long)st.series, (unsigned long long)st.points, (unsigned long long)rows.rows,
           st.completed_stages, st.total_stages, st.device_ms);
    for (uint64_t i = 0; i < rows.rows; i++)
        printf("  flowEnd %u throughput %.0f algoCalc %.3f stddev %.3f\n", rows.flow_end[i], rows.throughput[i],
               rows.algo_calc[i], rows.stddev[i]);
    tad_release(job);
    tad_free_columns(ctx, &cols);
    tad_shutdown(ctx);
    return 0;
}
