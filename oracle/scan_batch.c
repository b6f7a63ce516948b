/*
 * oracle/scan_batch.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; also bench.py's
 * cpu_baseline / --impl reference leg, kind "port").
 *
 * Batch drivers around oracle/jsre.c that restate the reference's two scan loops:
 *   - policy semantics  : gov/src/conditions/context.ts:9-25 (matchesAny: RegExp.test
 *                         per pattern) -> one boolean per (message, rule)
 *   - redaction semantics: gov/src/redaction/registry.ts:212-242 (findMatches: every
 *                         pattern in iteration order, global exec loop) followed by
 *                         registry.ts:288-316 (resolveOverlaps: stable sort by
 *                         start asc, length desc, category order asc; greedy keep)
 * Messages arrive as UTF-8 (bytes + offsets[n+1]) and are decoded to UTF-16 code
 * units first, because the reference matches JS strings.
 */
#define _GNU_SOURCE   /* sched_getaffinity, pthread_attr_setaffinity_np */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <sched.h>

typedef struct jsre jsre;
int jsre_exec(const jsre *re, const uint16_t *s, int n, int last_index, int *start, int *end);
int jsre_test(const jsre *re, const uint16_t *s, int n);
int jsre_find_all(const jsre *re, const uint16_t *s, int n, int *spans, int cap);

/* WHATWG-style UTF-8 decode to UTF-16 units; ill-formed bytes become U+FFFD */
static int utf8_to_utf16(const uint8_t *p, size_t n, uint16_t *out) {
    int k = 0; size_t i = 0;
    while (i < n) {
        uint32_t c = p[i];
        if (c < 0x80) { out[k++] = (uint16_t)c; i++; continue; }
        int need = 0; uint32_t cp = 0, lo = 0x80, hi = 0xbf;
        if (c >= 0xc2 && c <= 0xdf) { need = 1; cp = c & 0x1f; }
        else if (c >= 0xe0 && c <= 0xef) { need = 2; cp = c & 0x0f; if (c == 0xe0) lo = 0xa0; if (c == 0xed) hi = 0x9f; }
        else if (c >= 0xf0 && c <= 0xf4) { need = 3; cp = c & 0x07; if (c == 0xf0) lo = 0x90; if (c == 0xf4) hi = 0x8f; }
        else { out[k++] = 0xfffd; i++; continue; }
        size_t j = i + 1; int ok = 1;
        for (int t = 0; t < need; t++, j++) {
            if (j >= n || p[j] < lo || p[j] > hi) { ok = 0; break; }
            cp = (cp << 6) | (p[j] & 0x3f); lo = 0x80; hi = 0xbf;
        }
        if (!ok) { out[k++] = 0xfffd; i = j > i + 1 ? j : i + 1; continue; }
        i = j;
        if (cp >= 0x10000) { cp -= 0x10000; out[k++] = (uint16_t)(0xd800 + (cp >> 10)); out[k++] = (uint16_t)(0xdc00 + (cp & 0x3ff)); }
        else out[k++] = (uint16_t)cp;
    }
    return k;
}

typedef struct {
    jsre **res; int nres;
    const int *catrank;                /* per pattern: index in CATEGORY_ORDER (redaction mode) */
    const uint8_t *bytes; const uint64_t *off;
    size_t lo, hi;                     /* message range of this worker */
    int mode;                          /* 0 = policy booleans, 1 = redaction spans */
    uint8_t *hit_bits; size_t row_bytes;
    uint64_t *words;
    /* redaction output (per worker, merged afterwards) */
    int32_t *spans; size_t nspans, cspans;   /* quadruples msg, pat, start16, end16 */
    int error;
} Job;

typedef struct { int pat, start, end, rank, seq; } Cand;

static int cand_cmp(const void *a, const void *b) {
    const Cand *x = (const Cand *)a, *y = (const Cand *)b;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    int lx = x->end - x->start, ly = y->end - y->start;
    if (lx != ly) return lx > ly ? -1 : 1;
    if (x->rank != y->rank) return x->rank < y->rank ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq ? 1 : 0);   /* stable, like V8's sort */
}

static void push_span(Job *j, int32_t msg, int32_t pat, int32_t s, int32_t e) {
    if (j->nspans == j->cspans) { j->cspans = j->cspans ? j->cspans * 2 : 1024; j->spans = (int32_t *)realloc(j->spans, sizeof(int32_t) * 4 * j->cspans); }
    int32_t *q = j->spans + 4 * j->nspans++;
    q[0] = msg; q[1] = pat; q[2] = s; q[3] = e;
}

static void *worker(void *arg) {
    Job *j = (Job *)arg;
    size_t maxlen = 0;
    for (size_t i = j->lo; i < j->hi; i++) { size_t l = (size_t)(j->off[i + 1] - j->off[i]); if (l > maxlen) maxlen = l; }
    uint16_t *u = (uint16_t *)malloc(sizeof(uint16_t) * (maxlen + 1));
    int capsp = 4096; int *sp = (int *)malloc(sizeof(int) * 2 * capsp);
    Cand *cands = NULL; size_t ccap = 0;
    for (size_t i = j->lo; i < j->hi; i++) {
        int n = utf8_to_utf16(j->bytes + j->off[i], (size_t)(j->off[i + 1] - j->off[i]), u);
        if (j->mode == 0) {
            uint8_t *row = j->hit_bits ? j->hit_bits + i * j->row_bytes : NULL;
            uint32_t cnt = 0, first = 0xffffffffu;
            for (int r = 0; r < j->nres; r++) {
                int t = jsre_test(j->res[r], u, n);
                if (t < 0) { j->error = 1; t = 0; }
                if (t) { if (row) row[r >> 3] |= (uint8_t)(1u << (r & 7)); if (!cnt) first = (uint32_t)r; cnt++; }
            }
            if (j->words) j->words[i] = cnt ? ((1ull << 63) | ((uint64_t)cnt << 32) | first) : 0ull;
        } else {
            size_t nc = 0;
            for (int r = 0; r < j->nres; r++) {
                int c = jsre_find_all(j->res[r], u, n, sp, capsp);
                if (c < 0) { j->error = 1; continue; }
                if (c > capsp) { capsp = c; sp = (int *)realloc(sp, sizeof(int) * 2 * capsp); c = jsre_find_all(j->res[r], u, n, sp, capsp); }
                for (int q = 0; q < c; q++) {
                    if (nc == ccap) { ccap = ccap ? ccap * 2 : 256; cands = (Cand *)realloc(cands, sizeof(Cand) * ccap); }
                    cands[nc].pat = r; cands[nc].start = sp[2 * q]; cands[nc].end = sp[2 * q + 1];
                    cands[nc].rank = j->catrank ? j->catrank[r] : 0; cands[nc].seq = (int)nc; nc++;
                }
            }
            if (nc > 1) qsort(cands, nc, sizeof(Cand), cand_cmp);
            int last_end = -1;
            for (size_t q = 0; q < nc; q++) if (cands[q].start >= last_end) { push_span(j, (int32_t)i, cands[q].pat, cands[q].start, cands[q].end); last_end = cands[q].end; }
        }
    }
    free(u); free(sp); free(cands);
    return NULL;
}

static int run_jobs(Job *proto, size_t nmsg, int nthreads, Job **out_jobs) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > nmsg && nmsg > 0) nthreads = (int)nmsg;
    Job *jobs = (Job *)calloc((size_t)nthreads, sizeof(Job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    /* one thread per CPU of the caller's affinity mask, pinned (thread t -> the t-th allowed CPU, round robin): the timing
     * of the CPU arm should not depend on where the scheduler happens to put 128 threads of a cgroup-limited process */
    cpu_set_t allowed; int cpus[1024], ncpu = 0;
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0) for (int c = 0; c < 1024 && c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = *proto;
        jobs[t].lo = nmsg * (size_t)t / (size_t)nthreads; jobs[t].hi = nmsg * (size_t)(t + 1) / (size_t)nthreads;
        pthread_attr_t at; pthread_attr_init(&at); pthread_attr_setstacksize(&at, (size_t)512 << 20);
        if (ncpu > 0) { cpu_set_t one; CPU_ZERO(&one); CPU_SET(cpus[t % ncpu], &one); pthread_attr_setaffinity_np(&at, sizeof one, &one); }
        pthread_create(&th[t], &at, worker, &jobs[t]);
        pthread_attr_destroy(&at);
    }
    int err = 0;
    for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); err |= jobs[t].error; }
    free(th);
    *out_jobs = jobs;
    return err ? -nthreads : nthreads;
}

/* policy semantics.  hit_bits: nmsg rows of ceil(nres/8) bytes (may be NULL, must be
 * zeroed); words: nmsg result words (may be NULL):
 *   hit ? 1<<63 | count<<32 | lowest hit rule index : 0.  returns 0 / -1. */
int oracle_scan_policy(jsre **res, int nres, const uint8_t *bytes, const uint64_t *off, size_t nmsg,
                       int nthreads, uint8_t *hit_bits, uint64_t *words) {
    Job p; memset(&p, 0, sizeof p);
    p.res = res; p.nres = nres; p.bytes = bytes; p.off = off; p.mode = 0;
    p.hit_bits = hit_bits; p.row_bytes = (size_t)(nres + 7) / 8; p.words = words;
    Job *jobs; int r = run_jobs(&p, nmsg, nthreads, &jobs);
    free(jobs);
    return r < 0 ? -1 : 0;
}

/* redaction semantics.  res[] must already be in findMatches iteration order and
 * catrank[] gives each pattern's CATEGORY_ORDER index.  *out_spans is malloc'd
 * quadruples (msg, pat, start16, end16) in message order; free with oracle_free. */
long oracle_find_matches_batch(jsre **res, const int *catrank, int nres, const uint8_t *bytes, const uint64_t *off,
                               size_t nmsg, int nthreads, int32_t **out_spans) {
    Job p; memset(&p, 0, sizeof p);
    p.res = res; p.nres = nres; p.catrank = catrank; p.bytes = bytes; p.off = off; p.mode = 1;
    Job *jobs; int nt = run_jobs(&p, nmsg, nthreads, &jobs);
    int bad = nt < 0; if (bad) nt = -nt;
    size_t total = 0;
    for (int t = 0; t < nt; t++) total += jobs[t].nspans;
    int32_t *all = (int32_t *)malloc(sizeof(int32_t) * 4 * (total ? total : 1));
    size_t k = 0;
    for (int t = 0; t < nt; t++) { memcpy(all + 4 * k, jobs[t].spans, sizeof(int32_t) * 4 * jobs[t].nspans); k += jobs[t].nspans; free(jobs[t].spans); }
    free(jobs);
    *out_spans = all;
    return bad ? -1 : (long)total;
}

void oracle_free(void *p) { free(p); }
