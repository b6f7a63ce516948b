/*
 * oracle/sha256_merkle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * SHA-256 exactly as FIPS 180-4 section 6.2 states it.  The reference gets this
 * primitive from node:crypto (OpenSSL inside Node >= 22, not vendored):
 *     gov/src/util.ts:77-79, gov/src/redaction/vault.ts:26-28,
 *     nats/src/hooks.ts:90-94  -- createHash("sha256").update(str).digest("hex")
 * Pinned against hashlib, the NIST "abc"/448-bit/million-'a' vectors and the two
 * digests quoted in gov/RFC.md:1561-1562 (tests/test_oracle_golden.py).
 *
 * Merkle tree: the reference has NO implementation (only the
 * unimplemented hash chain of gov/RFC.md:782-800).  PARITY UNPINNED: the
 * convention below is this repository's own and is frozen here:
 *     leaf  = SHA-256(0x00 || leaf_bytes)                   (RFC 6962 2.1)
 *     node  = SHA-256(0x01 || left || right)
 *     level-wise pairing, an unpaired last node is promoted unchanged
 *     (same tree shape as RFC 6962 MTH for every n); empty tree = SHA-256("").
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const uint32_t K[64] = {
    0x428a2f98,0x71374491,0xb5c0fbcf,0xe9b5dba5,0x3956c25b,0x59f111f1,0x923f82a4,0xab1c5ed5,
    0xd807aa98,0x12835b01,0x243185be,0x550c7dc3,0x72be5d74,0x80deb1fe,0x9bdc06a7,0xc19bf174,
    0xe49b69c1,0xefbe4786,0x0fc19dc6,0x240ca1cc,0x2de92c6f,0x4a7484aa,0x5cb0a9dc,0x76f988da,
    0x983e5152,0xa831c66d,0xb00327c8,0xbf597fc7,0xc6e00bf3,0xd5a79147,0x06ca6351,0x14292967,
    0x27b70a85,0x2e1b2138,0x4d2c6dfc,0x53380d13,0x650a7354,0x766a0abb,0x81c2c92e,0x92722c85,
    0xa2bfe8a1,0xa81a664b,0xc24b8b70,0xc76c51a3,0xd192e819,0xd6990624,0xf40e3585,0x106aa070,
    0x19a4c116,0x1e376c08,0x2748774c,0x34b0bcb5,0x391c0cb3,0x4ed8aa4a,0x5b9cca4f,0x682e6ff3,
    0x748f82ee,0x78a5636f,0x84c87814,0x8cc70208,0x90befffa,0xa4506ceb,0xbef9a3f7,0xc67178f2
};
#define ROTR(x,n) (((x) >> (n)) | ((x) << (32 - (n))))

typedef struct { uint32_t h[8]; uint8_t buf[64]; uint64_t len; int fill; } sha_ctx;

static void sha_block(uint32_t h[8], const uint8_t *p) {
    uint32_t w[64];
    for (int t = 0; t < 16; t++) w[t] = ((uint32_t)p[4*t] << 24) | ((uint32_t)p[4*t+1] << 16) | ((uint32_t)p[4*t+2] << 8) | p[4*t+3];
    for (int t = 16; t < 64; t++) {
        uint32_t s0 = ROTR(w[t-15], 7) ^ ROTR(w[t-15], 18) ^ (w[t-15] >> 3);
        uint32_t s1 = ROTR(w[t-2], 17) ^ ROTR(w[t-2], 19) ^ (w[t-2] >> 10);
        w[t] = s1 + w[t-7] + s0 + w[t-16];
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int t = 0; t < 64; t++) {
        uint32_t S1 = ROTR(e, 6) ^ ROTR(e, 11) ^ ROTR(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K[t] + w[t];
        uint32_t S0 = ROTR(a, 2) ^ ROTR(a, 13) ^ ROTR(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
static void sha_init(sha_ctx *c) {
    static const uint32_t iv[8] = {0x6a09e667,0xbb67ae85,0x3c6ef372,0xa54ff53a,0x510e527f,0x9b05688c,0x1f83d9ab,0x5be0cd19};
    memcpy(c->h, iv, sizeof iv); c->len = 0; c->fill = 0;
}
static void sha_update(sha_ctx *c, const uint8_t *p, size_t n) {
    c->len += n;
    while (n) {
        if (c->fill == 0 && n >= 64) { sha_block(c->h, p); p += 64; n -= 64; continue; }
        size_t k = 64 - c->fill; if (k > n) k = n;
        memcpy(c->buf + c->fill, p, k); c->fill += (int)k; p += k; n -= k;
        if (c->fill == 64) { sha_block(c->h, c->buf); c->fill = 0; }
    }
}
static void sha_final(sha_ctx *c, uint8_t out[32]) {
    uint64_t bits = c->len * 8;
    uint8_t pad = 0x80; sha_update(c, &pad, 1);
    uint8_t z = 0; while (c->fill != 56) sha_update(c, &z, 1);
    uint8_t lb[8]; for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
    sha_update(c, lb, 8);
    for (int i = 0; i < 8; i++) { out[4*i] = c->h[i] >> 24; out[4*i+1] = c->h[i] >> 16; out[4*i+2] = c->h[i] >> 8; out[4*i+3] = c->h[i]; }
}

void oracle_sha256(const uint8_t *p, size_t n, uint8_t out[32]) {
    sha_ctx c; sha_init(&c); sha_update(&c, p, n); sha_final(&c, out);
}

/* digests of n variable-length items laid out as bytes + offsets[n+1] */
void oracle_sha256_batch(const uint8_t *bytes, const uint64_t *off, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; i++) oracle_sha256(bytes + off[i], (size_t)(off[i + 1] - off[i]), out + 32 * i);
}

void oracle_merkle_leaf(const uint8_t *p, size_t n, uint8_t out[32]) {
    sha_ctx c; sha_init(&c); uint8_t t = 0x00; sha_update(&c, &t, 1); sha_update(&c, p, n); sha_final(&c, out);
}
void oracle_merkle_node(const uint8_t l[32], const uint8_t r[32], uint8_t out[32]) {
    sha_ctx c; sha_init(&c); uint8_t t = 0x01; sha_update(&c, &t, 1); sha_update(&c, l, 32); sha_update(&c, r, 32); sha_final(&c, out);
}

/* fold n 32-byte nodes (in place, clobbers `nodes`) level by level into one root */
void oracle_merkle_fold(uint8_t *nodes, size_t n, uint8_t out[32]) {
    if (n == 0) { oracle_sha256((const uint8_t *)"", 0, out); return; }
    while (n > 1) {
        size_t m = 0;
        for (size_t i = 0; i + 1 < n; i += 2) { uint8_t t[32]; oracle_merkle_node(nodes + 32 * i, nodes + 32 * (i + 1), t); memcpy(nodes + 32 * m, t, 32); m++; }
        if (n & 1) { memmove(nodes + 32 * m, nodes + 32 * (n - 1), 32); m++; }
        n = m;
    }
    memcpy(out, nodes, 32);
}

/* root over n leaves given as bytes + offsets[n+1]; returns 0, or -1 on OOM */
int oracle_merkle_root(const uint8_t *bytes, const uint64_t *off, size_t n, uint8_t out[32]) {
    if (n == 0) { oracle_sha256((const uint8_t *)"", 0, out); return 0; }
    uint8_t *nodes = (uint8_t *)malloc(32 * n);
    if (!nodes) return -1;
    for (size_t i = 0; i < n; i++) oracle_merkle_leaf(bytes + off[i], (size_t)(off[i + 1] - off[i]), nodes + 32 * i);
    oracle_merkle_fold(nodes, n, out);
    free(nodes);
    return 0;
}

/* fixed-size leaves (the C3 layout: implicit offsets) */
int oracle_merkle_root_fixed(const uint8_t *bytes, size_t leaf_len, size_t n, uint8_t out[32]) {
    if (n == 0) { oracle_sha256((const uint8_t *)"", 0, out); return 0; }
    uint8_t *nodes = (uint8_t *)malloc(32 * n);
    if (!nodes) return -1;
    for (size_t i = 0; i < n; i++) oracle_merkle_leaf(bytes + leaf_len * i, leaf_len, nodes + 32 * i);
    oracle_merkle_fold(nodes, n, out);
    free(nodes);
    return 0;
}

/* The same tree with the leaf level spread over threads: aligned blocks of 2^16 leaves are reduced to their roots
 * independently (level-k nodes of the level-wise tree are exactly the roots of aligned 2^k-leaf blocks), the block roots
 * are folded at the end.  Only there so that the full-size C3 tree can be checked in seconds. */
#include <pthread.h>
typedef struct { const uint8_t *bytes; size_t leaf_len, n, nblocks; uint8_t *roots; volatile long *next; } MtJob;
static void *mt_worker(void *arg) {
    MtJob *j = (MtJob *)arg;
    for (;;) {
        long b = __sync_fetch_and_add(j->next, 1);
        if ((size_t)b >= j->nblocks) break;
        size_t lo = (size_t)b << 16, cnt = j->n - lo < ((size_t)1 << 16) ? j->n - lo : ((size_t)1 << 16);
        oracle_merkle_root_fixed(j->bytes + j->leaf_len * lo, j->leaf_len, cnt, j->roots + 32 * (size_t)b);
    }
    return NULL;
}
int oracle_merkle_root_fixed_mt(const uint8_t *bytes, size_t leaf_len, size_t n, int nthreads, uint8_t out[32]) {
    if (n <= ((size_t)1 << 16) || nthreads <= 1) return oracle_merkle_root_fixed(bytes, leaf_len, n, out);
    size_t nblocks = (n + 65535) >> 16;
    uint8_t *roots = (uint8_t *)malloc(32 * nblocks);
    if (!roots) return -1;
    volatile long next = 0;
    MtJob job = { bytes, leaf_len, n, nblocks, roots, &next };
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, mt_worker, &job);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    oracle_merkle_fold(roots, nblocks, out);
    free(roots);
    return 0;
}

/* RFC 6962 2.1 MTH stated recursively (split at the largest power of two < n):
 * used by the tests to show the level-wise fold above has the same tree shape. */
static void mth_rec(const uint8_t *leaf_hashes, size_t n, uint8_t out[32]) {
    if (n == 1) { memcpy(out, leaf_hashes, 32); return; }
    size_t k = 1; while (k * 2 < n) k *= 2;
    uint8_t l[32], r[32];
    mth_rec(leaf_hashes, k, l); mth_rec(leaf_hashes + 32 * k, n - k, r);
    oracle_merkle_node(l, r, out);
}
int oracle_merkle_root_rfc6962(const uint8_t *bytes, const uint64_t *off, size_t n, uint8_t out[32]) {
    if (n == 0) { oracle_sha256((const uint8_t *)"", 0, out); return 0; }
    uint8_t *nodes = (uint8_t *)malloc(32 * n);
    if (!nodes) return -1;
    for (size_t i = 0; i < n; i++) oracle_merkle_leaf(bytes + off[i], (size_t)(off[i + 1] - off[i]), nodes + 32 * i);
    mth_rec(nodes, n, out);
    free(nodes);
    return 0;
}
