/*
 * CPU oracle (C port) for the Theia throughput-anomaly-detection hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the
 * product package (theia_b200/).  Used by tests/ as the bulk checker and by
 * bench.py's cpu_baseline / --impl reference legs as the reported CPU baseline.
 *
 * Restates, stage by stage, the reference job
 * plugins/anomaly-detection/anomaly_detection.py (paths relative to the reference root):
 *   stage A  :566-613  filter + GROUP BY (key, flowEndSeconds) -> max()/sum()
 *   stage B  :680-684  groupby(key): series ordered by flowEndSeconds (engine contract),
 *                      stddev_samp = Spark CentralMomentAgg (Welford), NULL for n == 1
 *   stage C  :146-165 calculate_ewma, :168-212 calculate_ewma_anomaly,
 *            :312-349 calculate_dbscan / calculate_dbscan_anomaly
 *            (scikit-learn DBSCAN(min_samples=4, eps=250000000) on 1-D values)
 *   stage D  :352-421  explode + keep anomalous points
 * It is the same arithmetic, in the same order, as oracle/tad_oracle.py, which is pinned
 * to the reference's golden vectors and to outputs of the reference UDFs
 * (tests/test_oracle_golden.py); tests/test_oracle_c.py pins this file to the Python one.
 *
 * Build: oracle/Makefile  (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).
 * -ffp-contract=off: every FP64 operation is individually rounded, as in CPython.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ALGO_EWMA 0
#define ALGO_ARIMA 1
#define ALGO_DBSCAN 2
#define REDUCE_MAX 0
#define REDUCE_SUM 1

#define DBSCAN_EPS 250000000.0
#define DBSCAN_MIN_SAMPLES 4

typedef struct {
    uint64_t rows;
    const uint32_t *src_ip, *dst_ip, *flow_start, *flow_end;
    const uint16_t *src_port, *dst_port;
    const uint8_t *proto;
    const uint64_t *value;
} oracle_table;

typedef struct {
    int32_t algo, reducer;
    uint32_t start_time, end_time;
    int32_t emit_all, threads;
} oracle_spec;

typedef struct {
    uint64_t n, n_series, n_points;
    uint32_t *src_ip, *dst_ip, *flow_start, *flow_end;
    uint16_t *src_port, *dst_port;
    uint8_t *proto, *anomaly;
    double *stddev, *algo_calc, *throughput;
} oracle_result;

typedef struct {
    uint64_t a;      /* src_ip << 32 | dst_ip */
    uint64_t b;      /* flow_start << 32 | src_port << 16 | dst_port */
    uint64_t value;
    uint32_t t;      /* flow_end */
    uint32_t proto;
} prow;

typedef struct {
    prow *rows;
    double *sd, *calc;
    uint8_t *flag;
    uint64_t n, cap;
} outvec;

static inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

static int cmp_prow(const void *pa, const void *pb)
{
    const prow *x = (const prow *)pa, *y = (const prow *)pb;
    if (x->a != y->a) return x->a < y->a ? -1 : 1;
    if (x->b != y->b) return x->b < y->b ? -1 : 1;
    if (x->proto != y->proto) return x->proto < y->proto ? -1 : 1;
    if (x->t != y->t) return x->t < y->t ? -1 : 1;
    return 0;
}

static void out_push(outvec *o, const prow *key, uint32_t t, double sd, double calc, double x, uint8_t flag)
{
    if (o->n == o->cap) {
        o->cap = o->cap ? o->cap * 2 : 256;
        o->rows = (prow *)realloc(o->rows, o->cap * sizeof(prow));
        o->sd = (double *)realloc(o->sd, o->cap * sizeof(double));
        o->calc = (double *)realloc(o->calc, o->cap * sizeof(double));
        o->flag = (uint8_t *)realloc(o->flag, o->cap);
    }
    prow r = *key;
    r.t = t;
    memcpy(&r.value, &x, 8);            /* throughput as f64 bits */
    o->rows[o->n] = r; o->sd[o->n] = sd; o->calc[o->n] = calc; o->flag[o->n] = flag;
    o->n++;
}

/* sklearn brute-force squared distance (n <= 11), see oracle/tad_oracle.py:_sk_brute_d2 */
static inline double sk_brute_d2(double xi, double xj)
{
    double d = (xi * xi + (-2.0 * (xi * xj))) + xj * xj;
    return d > 0.0 ? d : 0.0;
}

typedef struct { double v; uint32_t i; } dv;
static int cmp_dv(const void *a, const void *b)
{
    double x = ((const dv *)a)->v, y = ((const dv *)b)->v;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* anomaly_detection.py:325-349, exact 1-D DBSCAN noise rule */
static void dbscan_flags(const double *x, uint32_t n, uint8_t *flag, dv *tmp, uint32_t *pc)
{
    if (n <= 11) {
        const double r2 = DBSCAN_EPS * DBSCAN_EPS;
        uint8_t core[11];
        for (uint32_t i = 0; i < n; i++) {
            uint32_t c = 0;
            for (uint32_t j = 0; j < n; j++) c += (i == j) || (sk_brute_d2(x[i], x[j]) <= r2);
            core[i] = c >= DBSCAN_MIN_SAMPLES;
        }
        for (uint32_t i = 0; i < n; i++) {
            uint8_t reach = core[i];
            for (uint32_t j = 0; j < n && !reach; j++)
                reach = core[j] && ((i == j) || (sk_brute_d2(x[i], x[j]) <= r2));
            flag[i] = !reach;
        }
        return;
    }
    for (uint32_t i = 0; i < n; i++) { tmp[i].v = x[i]; tmp[i].i = i; }
    qsort(tmp, n, sizeof(dv), cmp_dv);
    /* pass 1: core flags via two monotone pointers; pc = prefix count of cores */
    uint32_t lo = 0, hi = 0;
    pc[0] = 0;
    for (uint32_t k = 0; k < n; k++) {
        while (fabs(tmp[k].v - tmp[lo].v) > DBSCAN_EPS) lo++;
        if (hi < k) hi = k;
        while (hi + 1 < n && fabs(tmp[hi + 1].v - tmp[k].v) <= DBSCAN_EPS) hi++;
        pc[k + 1] = pc[k] + ((hi - lo + 1) >= DBSCAN_MIN_SAMPLES);
    }
    lo = 0; hi = 0;
    for (uint32_t k = 0; k < n; k++) {
        while (fabs(tmp[k].v - tmp[lo].v) > DBSCAN_EPS) lo++;
        if (hi < k) hi = k;
        while (hi + 1 < n && fabs(tmp[hi + 1].v - tmp[k].v) <= DBSCAN_EPS) hi++;
        flag[tmp[k].i] = (pc[hi + 1] - pc[lo]) == 0;     /* no core within eps (incl. itself) */
    }
}

/* one series: rows[0..n) share a key, sorted by t, duplicates already reduced */
static void do_series(const prow *rows, uint32_t n, const oracle_spec *sp, outvec *o,
                      double *x, double *calc, uint8_t *flag, dv *tmp, uint32_t *pc)
{
    /* stddev_samp: Spark CentralMomentAgg */
    double cnt = 0.0, avg = 0.0, m2 = 0.0;
    for (uint32_t i = 0; i < n; i++) {
        x[i] = (double)rows[i].value;                 /* round-to-nearest-even */
        cnt += 1.0;
        double d = x[i] - avg;
        double dn = d / cnt;
        avg = avg + dn;
        m2 = m2 + d * (d - dn);
    }
    int has_sd = n >= 2;
    double sd = has_sd ? sqrt(m2 / (cnt - 1.0)) : NAN;
    if (sp->algo == ALGO_EWMA) {
        double prev = 0.0;
        for (uint32_t i = 0; i < n; i++) {            /* :146-165, alpha = 0.5 */
            prev = (1.0 - 0.5) * prev + 0.5 * x[i];
            calc[i] = prev;
            flag[i] = has_sd && (fabs(x[i] - prev) > sd);   /* :208-210, strict > */
        }
    } else if (sp->algo == ALGO_DBSCAN) {
        for (uint32_t i = 0; i < n; i++) calc[i] = 0.0;     /* :312-322 */
        dbscan_flags(x, n, flag, tmp, pc);
    } else {
        return;                                             /* ARIMA: oracle/arima_oracle.py */
    }
    for (uint32_t i = 0; i < n; i++)
        if (sp->emit_all || flag[i]) out_push(o, &rows[0], rows[i].t, sd, calc[i], x[i], flag[i]);
}

int tad_oracle_run(const oracle_table *tb, const oracle_spec *sp, oracle_result *res)
{
    const uint64_t R = tb->rows;
    const uint32_t LOGP = R > (1u << 20) ? 12 : (R > 4096 ? 6 : 0);
    const uint32_t P = 1u << LOGP;
    int nthreads = sp->threads;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    memset(res, 0, sizeof(*res));
    prow *part = (prow *)malloc((R ? R : 1) * sizeof(prow));
    uint64_t *hist = (uint64_t *)calloc((size_t)P * nthreads + 1, sizeof(uint64_t));
    uint64_t *pstart = (uint64_t *)calloc(P + 1, sizeof(uint64_t));
    if (!part || !hist || !pstart) return -1;

#define ROW_KEEP(i) ((!sp->start_time || tb->flow_start[i] >= sp->start_time) && \
                     (!sp->end_time || tb->flow_end[i] < sp->end_time))
#define ROW_A(i) (((uint64_t)(tb->src_ip ? tb->src_ip[i] : 0) << 32) | (tb->dst_ip ? tb->dst_ip[i] : 0))
#define ROW_B(i) (((uint64_t)(tb->flow_start ? tb->flow_start[i] : 0) << 32) | \
                  ((uint64_t)(tb->src_port ? tb->src_port[i] : 0) << 16) | (tb->dst_port ? tb->dst_port[i] : 0))
#define ROW_P(i) ((uint32_t)(tb->proto ? tb->proto[i] : 0))
#define ROW_PART(a, b, p) (LOGP ? (uint32_t)(mix64((a) ^ mix64((b) + 0x9e3779b97f4a7c15ULL * ((p) + 1))) >> (64 - LOGP)) : 0u)

    /* stage A filter + hash partition (parallel counting sort) */
#pragma omp parallel num_threads(nthreads)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        int tid = 0, nt = 1;
#endif
        uint64_t lo = R * tid / nt, hi = R * (tid + 1) / nt;
        uint64_t *h = hist + (size_t)tid * P;
        for (uint64_t i = lo; i < hi; i++) {
            if (!ROW_KEEP(i)) continue;
            uint64_t a = ROW_A(i), b = ROW_B(i);
            h[ROW_PART(a, b, ROW_P(i))]++;
        }
#pragma omp barrier
#pragma omp single
        {
            uint64_t run = 0;
            for (uint32_t p = 0; p < P; p++) {
                pstart[p] = run;
                for (int t = 0; t < nt; t++) {
                    uint64_t c = hist[(size_t)t * P + p];
                    hist[(size_t)t * P + p] = run;
                    run += c;
                }
            }
            pstart[P] = run;
        }
        for (uint64_t i = lo; i < hi; i++) {
            if (!ROW_KEEP(i)) continue;
            prow r;
            r.a = ROW_A(i); r.b = ROW_B(i); r.proto = ROW_P(i);
            r.value = tb->value[i]; r.t = tb->flow_end[i];
            part[h[ROW_PART(r.a, r.b, r.proto)]++] = r;
        }
    }

    outvec *outs = (outvec *)calloc(P, sizeof(outvec));
    uint64_t n_series = 0, n_points = 0;
#pragma omp parallel num_threads(nthreads) reduction(+ : n_series, n_points)
    {
        uint32_t scap = 1024;
        double *x = (double *)malloc(scap * sizeof(double));
        double *calc = (double *)malloc(scap * sizeof(double));
        uint8_t *flag = (uint8_t *)malloc(scap);
        dv *tmp = (dv *)malloc(scap * sizeof(dv));
        uint32_t *pc = (uint32_t *)malloc((scap + 1) * sizeof(uint32_t));
#pragma omp for schedule(dynamic, 1)
        for (uint32_t p = 0; p < P; p++) {
            prow *rows = part + pstart[p];
            uint64_t n = pstart[p + 1] - pstart[p];
            if (!n) continue;
            qsort(rows, n, sizeof(prow), cmp_prow);
            /* stage A reduce duplicates of (key, t) in place */
            uint64_t w = 0;
            for (uint64_t i = 1; i < n; i++) {
                if (cmp_prow(&rows[w], &rows[i]) == 0) {
                    if (sp->reducer == REDUCE_MAX) { if (rows[i].value > rows[w].value) rows[w].value = rows[i].value; }
                    else rows[w].value += rows[i].value;
                } else rows[++w] = rows[i];
            }
            n = w + 1;
            n_points += n;
            uint64_t s0 = 0;
            for (uint64_t i = 1; i <= n; i++) {
                if (i == n || rows[i].a != rows[s0].a || rows[i].b != rows[s0].b || rows[i].proto != rows[s0].proto) {
                    uint32_t len = (uint32_t)(i - s0);
                    if (len > scap) {
                        scap = len * 2;
                        x = (double *)realloc(x, scap * sizeof(double));
                        calc = (double *)realloc(calc, scap * sizeof(double));
                        flag = (uint8_t *)realloc(flag, scap);
                        tmp = (dv *)realloc(tmp, scap * sizeof(dv));
                        pc = (uint32_t *)realloc(pc, (scap + 1) * sizeof(uint32_t));
                    }
                    do_series(rows + s0, len, sp, &outs[p], x, calc, flag, tmp, pc);
                    n_series++;
                    s0 = i;
                }
            }
        }
        free(x); free(calc); free(flag); free(tmp); free(pc);
    }

    uint64_t total = 0;
    for (uint32_t p = 0; p < P; p++) total += outs[p].n;
    uint64_t cap = total ? total : 1;
    res->n = total; res->n_series = n_series; res->n_points = n_points;
    res->src_ip = (uint32_t *)malloc(cap * 4); res->dst_ip = (uint32_t *)malloc(cap * 4);
    res->flow_start = (uint32_t *)malloc(cap * 4); res->flow_end = (uint32_t *)malloc(cap * 4);
    res->src_port = (uint16_t *)malloc(cap * 2); res->dst_port = (uint16_t *)malloc(cap * 2);
    res->proto = (uint8_t *)malloc(cap); res->anomaly = (uint8_t *)malloc(cap);
    res->stddev = (double *)malloc(cap * 8); res->algo_calc = (double *)malloc(cap * 8);
    res->throughput = (double *)malloc(cap * 8);
    uint64_t k = 0;
    for (uint32_t p = 0; p < P; p++) {
        outvec *o = &outs[p];
        for (uint64_t i = 0; i < o->n; i++, k++) {
            const prow *r = &o->rows[i];
            res->src_ip[k] = (uint32_t)(r->a >> 32); res->dst_ip[k] = (uint32_t)r->a;
            res->flow_start[k] = (uint32_t)(r->b >> 32);
            res->src_port[k] = (uint16_t)(r->b >> 16); res->dst_port[k] = (uint16_t)r->b;
            res->proto[k] = (uint8_t)r->proto; res->flow_end[k] = r->t;
            memcpy(&res->throughput[k], &r->value, 8);
            res->stddev[k] = o->sd[i]; res->algo_calc[k] = o->calc[i]; res->anomaly[k] = o->flag[i];
        }
        free(o->rows); free(o->sd); free(o->calc); free(o->flag);
    }
    free(outs); free(part); free(hist); free(pstart);
    return 0;
}

void tad_oracle_free(oracle_result *r)
{
    free(r->src_ip); free(r->dst_ip); free(r->flow_start); free(r->flow_end);
    free(r->src_port); free(r->dst_port); free(r->proto); free(r->anomaly);
    free(r->stddev); free(r->algo_calc); free(r->throughput);
    memset(r, 0, sizeof(*r));
}

/* Per-series entry points for direct comparison with the Python restatement. */
void tad_oracle_ewma(const uint64_t *v, uint32_t n, double *calc, uint8_t *flag, double *sd_out)
{
    oracle_spec sp = {ALGO_EWMA, REDUCE_MAX, 0, 0, 1, 1};
    outvec o = {0};
    prow *rows = (prow *)calloc(n ? n : 1, sizeof(prow));
    double *x = (double *)malloc((n + 1) * 8);
    for (uint32_t i = 0; i < n; i++) { rows[i].value = v[i]; rows[i].t = i; }
    do_series(rows, n, &sp, &o, x, calc, flag, NULL, NULL);
    *sd_out = n ? o.sd[0] : NAN;
    free(o.rows); free(o.sd); free(o.calc); free(o.flag); free(rows); free(x);
}

void tad_oracle_dbscan(const uint64_t *v, uint32_t n, uint8_t *flag)
{
    double *x = (double *)malloc((n + 1) * 8);
    dv *tmp = (dv *)malloc((n + 1) * sizeof(dv));
    uint32_t *pc = (uint32_t *)malloc((n + 2) * 4);
    for (uint32_t i = 0; i < n; i++) x[i] = (double)v[i];
    dbscan_flags(x, n, flag, tmp, pc);
    free(x); free(tmp); free(pc);
}

/* ARIMA has no C port: the authoritative ARIMA oracle is oracle/arima_oracle.py (SciPy's own L-BFGS-B and Brent, the routines
 * statsmodels itself calls); tad_oracle_run() returns no rows for algo = ARIMA.  Callers ask here before they rely on it. */
int tad_oracle_arima_available(void) { return 0; }
