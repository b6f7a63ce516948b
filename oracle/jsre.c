/*
 * oracle/jsre.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).
 *
 * A plain-C restatement of the ECMAScript (ECMA-262 22.2.2, non-unicode mode,
 * Annex B web syntax) backtracking RegExp semantics that the reference obtains
 * from V8/Irregexp (a runtime dependency of Node >= 22, not vendored under
 * /root/reference).  The reference's scan path calls it at:
 *     gov/src/redaction/registry.ts:222-236   new RegExp(src,"g"[+"i"]) + exec loop
 *     gov/src/conditions/context.ts:9-25      RegExp.test per rule pattern
 *     gov/src/policy-loader.ts:119-128        new RegExp(pattern) (no flags)
 * It works on UTF-16 code units exactly like a JS string does, uses the spec's
 * continuation-passing matcher structure (RepeatMatcher with the empty check,
 * leftmost-first alternation, direction-aware lookbehind, backreferences), and
 * is deliberately a different algorithm from the product's byte-level Pike VM /
 * DFA so that agreement between the two is meaningful.
 *
 * Parity pin: tests/test_oracle_golden.py runs this against every vector of
 * gov/test/redaction/registry.test.ts (tests/golden/registry_vectors.json).
 *
 * Case-insensitive matching: Canonicalize() is exact for code units < 128 and the
 * identity above (non-unicode mode never maps a non-ASCII unit onto an ASCII one,
 * ECMA-262 22.2.2.7.3 step 5).  Patterns with flag i that *contain* non-ASCII
 * literals would need the Unicode toUpperCase table and are rejected.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define JSRE_INF 0x7fffffff

enum {
    N_EMPTY, N_CHAR, N_ANY, N_CLASS, N_SEQ, N_ALT, N_GROUP, N_REPEAT,
    N_BOL, N_EOL, N_WORDB, N_NWORDB, N_LOOK, N_BACKREF
};

typedef struct Node Node;
struct Node {
    int type;
    int ch;                 /* N_CHAR: code unit; N_BACKREF: group number */
    int negate;             /* N_CLASS / N_LOOK */
    int nranges;            /* N_CLASS */
    int *ranges;            /* pairs lo,hi (inclusive) */
    int nkids;
    Node **kids;            /* N_SEQ/N_ALT children; [0] for GROUP/REPEAT/LOOK */
    int cap;                /* N_GROUP: capture index (>=1) or 0 for (?: ) */
    int min, max, greedy;   /* N_REPEAT */
    int cap_lo, cap_hi;     /* N_REPEAT: capture indices inside the body [lo,hi) */
    int behind;             /* N_LOOK */
};

typedef struct jsre {
    Node *root;
    int ncap;               /* number of capture groups + 1 */
    int icase;
    Node **all; int nall, call; /* for freeing */
} jsre;

/* ------------------------------------------------------------------ parser */

typedef struct {
    const uint16_t *p; int n; int i;
    jsre *re;
    int ncap;
    int total_groups;   /* pre-scan count of capturing groups (for \N decisions) */
    char *err; int errlen; int failed;
} Parser;

static void perr(Parser *ps, const char *msg) {
    if (!ps->failed) { ps->failed = 1; if (ps->err) snprintf(ps->err, ps->errlen, "%s (at %d)", msg, ps->i); }
}

static Node *mknode(Parser *ps, int type) {
    Node *nd = (Node *)calloc(1, sizeof(Node));
    nd->type = type;
    jsre *re = ps->re;
    if (re->nall == re->call) { re->call = re->call ? re->call * 2 : 64; re->all = (Node **)realloc(re->all, sizeof(Node *) * re->call); }
    re->all[re->nall++] = nd;
    return nd;
}
static void addkid(Node *nd, Node *k) {
    nd->kids = (Node **)realloc(nd->kids, sizeof(Node *) * (nd->nkids + 1));
    nd->kids[nd->nkids++] = k;
}
static void addrange(Node *nd, int lo, int hi) {
    nd->ranges = (int *)realloc(nd->ranges, sizeof(int) * 2 * (nd->nranges + 1));
    nd->ranges[2 * nd->nranges] = lo; nd->ranges[2 * nd->nranges + 1] = hi; nd->nranges++;
}

static int peek(Parser *ps) { return ps->i < ps->n ? ps->p[ps->i] : -1; }
static int peek2(Parser *ps, int k) { return ps->i + k < ps->n ? ps->p[ps->i + k] : -1; }
static int isdig(int c) { return c >= '0' && c <= '9'; }
static int hexval(int c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

/* JS WhiteSpace + LineTerminator code units (ECMA-262 12.2, 12.3): the \s set */
static const int WS_RANGES[] = {
    0x09, 0x0d, 0x20, 0x20, 0xa0, 0xa0, 0x1680, 0x1680, 0x2000, 0x200a,
    0x2028, 0x2029, 0x202f, 0x202f, 0x205f, 0x205f, 0x3000, 0x3000, 0xfeff, 0xfeff
};
static const int WORD_RANGES[] = { '0', '9', 'A', 'Z', '_', '_', 'a', 'z' };
static const int DIGIT_RANGES[] = { '0', '9' };

static void add_set(Node *cls, const int *r, int nr, int invert) {
    if (!invert) { for (int i = 0; i < nr; i++) addrange(cls, r[2 * i], r[2 * i + 1]); return; }
    int lo = 0;
    for (int i = 0; i < nr; i++) { if (r[2 * i] > lo) addrange(cls, lo, r[2 * i] - 1); lo = r[2 * i + 1] + 1; }
    if (lo <= 0xffff) addrange(cls, lo, 0xffff);
}
/* returns 1 if escape was a class escape added to cls */
static int class_escape_set(Node *cls, int c) {
    switch (c) {
    case 'd': add_set(cls, DIGIT_RANGES, 1, 0); return 1;
    case 'D': add_set(cls, DIGIT_RANGES, 1, 1); return 1;
    case 'w': add_set(cls, WORD_RANGES, 4, 0); return 1;
    case 'W': add_set(cls, WORD_RANGES, 4, 1); return 1;
    case 's': add_set(cls, WS_RANGES, 10, 0); return 1;
    case 'S': add_set(cls, WS_RANGES, 10, 1); return 1;
    }
    return 0;
}

/* Parses the body of a character escape after the backslash (shared by atom and
 * class contexts).  Returns the code unit, or -1 on error. */
static int parse_char_escape(Parser *ps, int in_class) {
    int c = peek(ps);
    if (c < 0) { perr(ps, "\\ at end of pattern"); return -1; }
    ps->i++;
    switch (c) {
    case 't': return 0x09; case 'n': return 0x0a; case 'v': return 0x0b;
    case 'f': return 0x0c; case 'r': return 0x0d;
    case 'b': if (in_class) return 0x08; break;
    case '0': if (!isdig(peek(ps))) return 0; break;
    case 'c': {
        int l = peek(ps);
        if ((l >= 'a' && l <= 'z') || (l >= 'A' && l <= 'Z')) { ps->i++; return l % 32; }
        if (in_class && (isdig(l) || l == '_')) { ps->i++; return l % 32; }
        ps->i--; return '\\';   /* Annex B: "\c" is a literal backslash, 'c' re-read */
    }
    case 'x': {
        int h1 = hexval(peek(ps)), h2 = hexval(peek2(ps, 1));
        if (h1 >= 0 && h2 >= 0) { ps->i += 2; return h1 * 16 + h2; }
        return 'x';
    }
    case 'u': {
        int v = 0, ok = 1;
        for (int k = 0; k < 4; k++) { int h = hexval(peek2(ps, k)); if (h < 0) { ok = 0; break; } v = v * 16 + h; }
        if (ok) { ps->i += 4; return v; }
        return 'u';
    }
    }
    if (c >= '0' && c <= '7') {            /* legacy octal escape */
        int v = c - '0';
        int d = peek(ps);
        if (d >= '0' && d <= '7') { v = v * 8 + (d - '0'); ps->i++;
            d = peek(ps);
            if (c <= '3' && d >= '0' && d <= '7') { v = v * 8 + (d - '0'); ps->i++; } }
        return v;
    }
    return c;                               /* identity escape */
}

static Node *parse_disjunction(Parser *ps);

static Node *parse_class(Parser *ps) {
    Node *cls = mknode(ps, N_CLASS);
    if (peek(ps) == '^') { cls->negate = 1; ps->i++; }
    for (;;) {
        int c = peek(ps);
        if (c < 0) { perr(ps, "unterminated character class"); return cls; }
        if (c == ']') { ps->i++; break; }
        int lo = -1, lo_is_set = 0;
        ps->i++;
        if (c == '\\') {
            int e = peek(ps);
            if (e == 'd' || e == 'D' || e == 'w' || e == 'W' || e == 's' || e == 'S') {
                ps->i++; class_escape_set(cls, e); lo_is_set = 1;
            } else lo = parse_char_escape(ps, 1);
        } else lo = c;
        if (ps->failed) return cls;
        /* range? */
        if (peek(ps) == '-' && peek2(ps, 1) != ']' && peek2(ps, 1) >= 0) {
            int save = ps->i;
            ps->i++;
            int c2 = peek(ps), hi = -1, hi_is_set = 0;
            ps->i++;
            if (c2 == '\\') {
                int e = peek(ps);
                if (e == 'd' || e == 'D' || e == 'w' || e == 'W' || e == 's' || e == 'S') {
                    ps->i++; hi_is_set = 1;
                    if (!lo_is_set) addrange(cls, lo, lo);
                    addrange(cls, '-', '-');
                    class_escape_set(cls, e);
                    continue;
                } else hi = parse_char_escape(ps, 1);
            } else hi = c2;
            (void)hi_is_set;
            if (lo_is_set) {            /* Annex B: [\d-x] => set, '-', x */
                addrange(cls, '-', '-'); addrange(cls, hi, hi); continue;
            }
            if (hi < lo) { ps->i = save; perr(ps, "range out of order in character class"); return cls; }
            addrange(cls, lo, hi); if (hi >= 128) cls->ch = 1;
            continue;
        }
        if (!lo_is_set) { addrange(cls, lo, lo); if (lo >= 128) cls->ch = 1; }
    }
    return cls;
}

/* try to parse {m}, {m,}, {m,n}; returns 1 and advances if it is a quantifier */
static int parse_braces(Parser *ps, int *mn, int *mx) {
    int j = ps->i;
    if (j >= ps->n || ps->p[j] != '{') return 0;
    j++;
    if (j >= ps->n || !isdig(ps->p[j])) return 0;
    long a = 0; while (j < ps->n && isdig(ps->p[j])) { a = a * 10 + (ps->p[j] - '0'); if (a > JSRE_INF) a = JSRE_INF; j++; }
    long b = a;
    if (j < ps->n && ps->p[j] == ',') {
        j++;
        if (j < ps->n && isdig(ps->p[j])) { b = 0; while (j < ps->n && isdig(ps->p[j])) { b = b * 10 + (ps->p[j] - '0'); if (b > JSRE_INF) b = JSRE_INF; j++; } }
        else b = JSRE_INF;
    }
    if (j >= ps->n || ps->p[j] != '}') return 0;
    ps->i = j + 1; *mn = (int)a; *mx = (int)b;
    return 1;
}

static Node *parse_atom_escape(Parser *ps) {
    int c = peek(ps);
    if (c == 'd' || c == 'D' || c == 'w' || c == 'W' || c == 's' || c == 'S') {
        ps->i++; Node *cls = mknode(ps, N_CLASS); class_escape_set(cls, c); return cls;
    }
    if (c == 'b') { ps->i++; return mknode(ps, N_WORDB); }
    if (c == 'B') { ps->i++; return mknode(ps, N_NWORDB); }
    if (c >= '1' && c <= '9') {
        /* DecimalEscape: backreference if that many groups exist in the whole pattern */
        int j = ps->i; long v = 0;
        while (j < ps->n && isdig(ps->p[j])) { v = v * 10 + (ps->p[j] - '0'); if (v > 100000) v = 100000; j++; }
        if (v <= ps->total_groups) { ps->i = j; Node *br = mknode(ps, N_BACKREF); br->ch = (int)v; return br; }
        if (c >= '8') { ps->i++; Node *ch = mknode(ps, N_CHAR); ch->ch = c; return ch; }
    }
    if (c == 'k') { /* named backreference only if the pattern has named groups; we treat as identity */ }
    int v = parse_char_escape(ps, 0);
    Node *ch = mknode(ps, N_CHAR); ch->ch = v; return ch;
}

static int count_groups(const uint16_t *p, int n) {
    int cnt = 0, in_class = 0;
    for (int i = 0; i < n; i++) {
        int c = p[i];
        if (c == '\\') { i++; continue; }
        if (in_class) { if (c == ']') in_class = 0; continue; }
        if (c == '[') { in_class = 1; continue; }
        if (c == '(') {
            if (i + 1 < n && p[i + 1] == '?') {
                if (i + 2 < n && p[i + 2] == '<' && i + 3 < n && p[i + 3] != '=' && p[i + 3] != '!') cnt++;
            } else cnt++;
        }
    }
    return cnt;
}

static Node *parse_term(Parser *ps, int *is_assertion) {
    *is_assertion = 0;
    int c = peek(ps);
    Node *atom = NULL;
    ps->i++;
    switch (c) {
    case '^': *is_assertion = 1; return mknode(ps, N_BOL);
    case '$': *is_assertion = 1; return mknode(ps, N_EOL);
    case '.': atom = mknode(ps, N_ANY); break;
    case '[': atom = parse_class(ps); break;
    case '\\': atom = parse_atom_escape(ps);
        if (atom && (atom->type == N_WORDB || atom->type == N_NWORDB)) { *is_assertion = 1; return atom; }
        break;
    case '(': {
        if (peek(ps) == '?') {
            int d = peek2(ps, 1);
            if (d == ':') { ps->i += 2; atom = mknode(ps, N_GROUP); atom->cap = 0; addkid(atom, parse_disjunction(ps)); }
            else if (d == '=' || d == '!') {
                ps->i += 2; atom = mknode(ps, N_LOOK); atom->negate = (d == '!'); atom->behind = 0;
                addkid(atom, parse_disjunction(ps));
                /* Annex B: lookaheads are quantifiable; lookbehinds are not */
            } else if (d == '<' && (peek2(ps, 2) == '=' || peek2(ps, 2) == '!')) {
                int neg = peek2(ps, 2) == '!';
                ps->i += 3; atom = mknode(ps, N_LOOK); atom->negate = neg; atom->behind = 1;
                addkid(atom, parse_disjunction(ps));
                *is_assertion = 1;
            } else if (d == '<') {
                ps->i += 2;
                while (peek(ps) >= 0 && peek(ps) != '>') ps->i++;
                if (peek(ps) != '>') { perr(ps, "invalid capture group name"); return mknode(ps, N_EMPTY); }
                ps->i++;
                atom = mknode(ps, N_GROUP); atom->cap = ++ps->ncap; addkid(atom, parse_disjunction(ps));
            } else { perr(ps, "invalid group"); return mknode(ps, N_EMPTY); }
        } else {
            atom = mknode(ps, N_GROUP); atom->cap = ++ps->ncap; addkid(atom, parse_disjunction(ps));
        }
        if (ps->failed) return atom;
        if (peek(ps) != ')') { perr(ps, "unterminated group"); return atom; }
        ps->i++;
        if (*is_assertion) return atom;
        break;
    }
    case '*': case '+': case '?': perr(ps, "nothing to repeat"); return mknode(ps, N_EMPTY);
    case '{': {
        ps->i--; int mn, mx, save = ps->i;
        if (parse_braces(ps, &mn, &mx)) { ps->i = save; perr(ps, "nothing to repeat"); return mknode(ps, N_EMPTY); }
        ps->i++; atom = mknode(ps, N_CHAR); atom->ch = '{'; break;   /* Annex B literal */
    }
    default: atom = mknode(ps, N_CHAR); atom->ch = c; break;
    }
    return atom;
}

static void cap_range(Node *nd, int *lo, int *hi) {
    if (nd->type == N_GROUP && nd->cap) { if (nd->cap < *lo) *lo = nd->cap; if (nd->cap + 1 > *hi) *hi = nd->cap + 1; }
    for (int i = 0; i < nd->nkids; i++) cap_range(nd->kids[i], lo, hi);
}

static Node *parse_alternative(Parser *ps) {
    Node *seq = mknode(ps, N_SEQ);
    while (!ps->failed) {
        int c = peek(ps);
        if (c < 0 || c == '|' || c == ')') break;
        int is_assert = 0;
        Node *t = parse_term(ps, &is_assert);
        if (ps->failed) break;
        if (!is_assert) {
            int mn = -1, mx = -1;
            int q = peek(ps);
            if (q == '*') { mn = 0; mx = JSRE_INF; ps->i++; }
            else if (q == '+') { mn = 1; mx = JSRE_INF; ps->i++; }
            else if (q == '?') { mn = 0; mx = 1; ps->i++; }
            else if (q == '{') { if (!parse_braces(ps, &mn, &mx)) mn = -1; }
            if (mn >= 0) {
                if (mx < mn) { perr(ps, "numbers out of order in {} quantifier"); break; }
                Node *r = mknode(ps, N_REPEAT);
                r->min = mn; r->max = mx; r->greedy = 1;
                if (peek(ps) == '?') { r->greedy = 0; ps->i++; }
                addkid(r, t);
                r->cap_lo = 1 << 30; r->cap_hi = 0; cap_range(t, &r->cap_lo, &r->cap_hi);
                if (r->cap_hi == 0) r->cap_lo = 0;
                t = r;
                int q2 = peek(ps);
                if (q2 == '*' || q2 == '+' || q2 == '?') { perr(ps, "nothing to repeat"); break; }
                { int a, b, sv = ps->i; if (q2 == '{' && parse_braces(ps, &a, &b)) { ps->i = sv; perr(ps, "nothing to repeat"); break; } }
            }
        }
        addkid(seq, t);
    }
    return seq;
}

static Node *parse_disjunction(Parser *ps) {
    Node *first = parse_alternative(ps);
    if (peek(ps) != '|') return first;
    Node *alt = mknode(ps, N_ALT);
    addkid(alt, first);
    while (!ps->failed && peek(ps) == '|') { ps->i++; addkid(alt, parse_alternative(ps)); }
    return alt;
}


void jsre_free(jsre *re) {
    if (!re) return;
    for (int i = 0; i < re->nall; i++) { free(re->all[i]->ranges); free(re->all[i]->kids); free(re->all[i]); }
    free(re->all); free(re);
}

/* flags: bit0 = ignoreCase */
jsre *jsre_compile(const uint16_t *pat, int n, int flags, char *err, int errlen) {
    jsre *re = (jsre *)calloc(1, sizeof(jsre));
    Parser ps; memset(&ps, 0, sizeof ps);
    ps.p = pat; ps.n = n; ps.re = re; ps.err = err; ps.errlen = errlen;
    ps.total_groups = count_groups(pat, n);
    if (err && errlen) err[0] = 0;
    re->root = parse_disjunction(&ps);
    if (!ps.failed && ps.i < n) perr(&ps, ps.p[ps.i] == ')' ? "unmatched ')'" : "unexpected character");
    re->ncap = ps.ncap + 1;
    re->icase = flags & 1;
    if (ps.failed) { jsre_free(re); return NULL; }
    return re;
}

/* ----------------------------------------------------------------- matcher */

typedef struct Cont Cont;
typedef struct M {
    const jsre *re;
    const uint16_t *s; int n;
    int *caps;              /* 2*ncap, -1 = undefined */
    long steps, step_limit;
    int aborted;
    int end;                /* end index of overall match */
} M;
typedef int (*ContFn)(M *, const Cont *, int);
struct Cont { ContFn fn; const Cont *next; const Node *node; int a, b, x, dir; };

/* ECMA-262 22.2.2.7.3 Canonicalize without the unicode flag: toUppercase when that is ONE code unit and does not turn a
 * non-ASCII unit into an ASCII one.  The non-identity part for units >= 0x80 is generated from the Unicode tables
 * (vainplex_openclaw_b200/csrc/gen_case_table.py); pinned by the reference's language-pack tests
 * (tests/golden/cortex_pack_vectors.json). */
#include "../vainplex_openclaw_b200/csrc/case_canon.inc"
static int canon_unit(int c) {
    if (c >= 'a' && c <= 'z') return c - 32;
    if (c < 0x80) return c;
    int lo = 0, hi = kCaseCanonCount - 1;
    while (lo <= hi) { int mid = (lo + hi) / 2; if (kCaseCanon[mid][0] == c) return kCaseCanon[mid][1]; if (kCaseCanon[mid][0] < c) lo = mid + 1; else hi = mid - 1; }
    return c;
}
static inline int canon(const M *m, int c) { return m->re->icase ? canon_unit(c) : c; }
static inline int isword(int c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || c == '_' || (c >= 'a' && c <= 'z'); }
static inline int is_lineterm(int c) { return c == 0x0a || c == 0x0d || c == 0x2028 || c == 0x2029; }

static int class_has(const M *m, const Node *cls, int c) {
    int found = 0;
    for (int i = 0; i < cls->nranges && !found; i++) if (c >= cls->ranges[2 * i] && c <= cls->ranges[2 * i + 1]) found = 1;
    if (!found && m->re->icase) {
        /* exists member a with Canonicalize(a) == Canonicalize(c) (22.2.2.9 CharacterSetMatcher) */
        int alt = -1;
        if (c >= 'a' && c <= 'z') alt = c - 32; else if (c >= 'A' && c <= 'Z') alt = c + 32;
        if (alt >= 0) for (int i = 0; i < cls->nranges && !found; i++) if (alt >= cls->ranges[2 * i] && alt <= cls->ranges[2 * i + 1]) found = 1;
        if (!found && c >= 0x80) {
            const int cc = canon_unit(c);
            for (int i = 0; i < cls->nranges && !found; i++) if (cc >= cls->ranges[2 * i] && cc <= cls->ranges[2 * i + 1]) found = 1;
            for (int k = 0; k < kCaseCanonCount && !found; k++) if (kCaseCanon[k][1] == cc) {
                const int u = kCaseCanon[k][0];
                for (int i = 0; i < cls->nranges && !found; i++) if (u >= cls->ranges[2 * i] && u <= cls->ranges[2 * i + 1]) found = 1;
            }
        }
    }
    return cls->negate ? !found : found;
}

static int unit_matches(const M *m, const Node *nd, int c) {
    switch (nd->type) {
    case N_CHAR: return canon(m, c) == canon(m, nd->ch);
    case N_ANY: return !is_lineterm(c);
    case N_CLASS: return class_has(m, nd, c);
    }
    return 0;
}

static int match_node(M *m, const Node *nd, int pos, int dir, const Cont *k);

static int k_seq(M *m, const Cont *k, int pos) {
    const Node *seq = k->node; int idx = k->a, dir = k->dir;
    if (idx < 0 || idx >= seq->nkids) return k->next->fn(m, k->next, pos);
    Cont c2 = *k; c2.a = idx + dir;
    return match_node(m, seq->kids[idx], pos, dir, &c2);
}

static int k_group_end(M *m, const Cont *k, int pos) {
    int cap = k->node->cap, startpos = k->x;
    int o0 = m->caps[2 * cap], o1 = m->caps[2 * cap + 1];
    if (k->dir > 0) { m->caps[2 * cap] = startpos; m->caps[2 * cap + 1] = pos; }
    else { m->caps[2 * cap] = pos; m->caps[2 * cap + 1] = startpos; }
    if (k->next->fn(m, k->next, pos)) return 1;
    m->caps[2 * cap] = o0; m->caps[2 * cap + 1] = o1;
    return 0;
}

static int rep(M *m, const Node *r, int min, int max, int pos, int dir, const Cont *k);

static int k_rep(M *m, const Cont *k, int pos) {
    /* continuation d of RepeatMatcher (22.2.2.3.1 step 2) */
    int min = k->a, max = k->b;
    if (min == 0 && pos == k->x) return 0;          /* empty check */
    int min2 = min == 0 ? 0 : min - 1;
    int max2 = max == JSRE_INF ? JSRE_INF : max - 1;
    return rep(m, k->node, min2, max2, pos, k->dir, k->next);
}

static int rep(M *m, const Node *r, int min, int max, int pos, int dir, const Cont *k) {
    if (max == 0) return k->fn(m, k, pos);
    const Node *body = r->kids[0];
    if (body->type == N_CHAR || body->type == N_ANY || body->type == N_CLASS) {
        /* single-unit atom: no internal choice points, so iterate instead of recursing */
        int cnt = 0;
        if (r->greedy) {
            while (cnt < max) {
                int p = dir > 0 ? pos + cnt : pos - cnt - 1;
                if (p < 0 || p >= m->n || !unit_matches(m, body, m->s[p])) break;
                cnt++;
            }
            for (int j = cnt; j >= min; j--) {
                if (++m->steps > m->step_limit) { m->aborted = 1; return 0; }
                if (k->fn(m, k, pos + dir * j)) return 1;
                if (m->aborted) return 0;
            }
            return 0;
        } else {
            for (int j = 0; j <= max; j++) {
                if (j >= min) {
                    if (++m->steps > m->step_limit) { m->aborted = 1; return 0; }
                    if (k->fn(m, k, pos + dir * j)) return 1;
                    if (m->aborted) return 0;
                }
                if (j == max) break;
                int p = dir > 0 ? pos + j : pos - j - 1;
                if (p < 0 || p >= m->n || !unit_matches(m, body, m->s[p])) break;
            }
            return 0;
        }
    }
    Cont d; d.fn = k_rep; d.next = k; d.node = r; d.a = min; d.b = max; d.x = pos; d.dir = dir;
    /* step 3-4: clear the captures inside the body */
    int ncl = r->cap_hi - r->cap_lo;
    int saved_buf[64]; int *saved = saved_buf;
    if (ncl > 32) saved = (int *)malloc(sizeof(int) * 2 * ncl);
    for (int i = 0; i < ncl; i++) {
        saved[2 * i] = m->caps[2 * (r->cap_lo + i)]; saved[2 * i + 1] = m->caps[2 * (r->cap_lo + i) + 1];
        m->caps[2 * (r->cap_lo + i)] = -1; m->caps[2 * (r->cap_lo + i) + 1] = -1;
    }
    int res;
    if (min != 0) res = match_node(m, body, pos, dir, &d);
    else if (!r->greedy) {
        /* lazy: try the continuation with the ORIGINAL captures first */
        for (int i = 0; i < ncl; i++) { m->caps[2 * (r->cap_lo + i)] = saved[2 * i]; m->caps[2 * (r->cap_lo + i) + 1] = saved[2 * i + 1]; }
        res = k->fn(m, k, pos);
        if (!res && !m->aborted) {
            for (int i = 0; i < ncl; i++) { m->caps[2 * (r->cap_lo + i)] = -1; m->caps[2 * (r->cap_lo + i) + 1] = -1; }
            res = match_node(m, body, pos, dir, &d);
        }
    } else {
        res = match_node(m, body, pos, dir, &d);
        if (!res && !m->aborted) {
            for (int i = 0; i < ncl; i++) { m->caps[2 * (r->cap_lo + i)] = saved[2 * i]; m->caps[2 * (r->cap_lo + i) + 1] = saved[2 * i + 1]; }
            res = k->fn(m, k, pos);
        }
    }
    if (!res) for (int i = 0; i < ncl; i++) { m->caps[2 * (r->cap_lo + i)] = saved[2 * i]; m->caps[2 * (r->cap_lo + i) + 1] = saved[2 * i + 1]; }
    if (saved != saved_buf) free(saved);
    return res;
}

static int k_accept(M *m, const Cont *k, int pos) { (void)k; m->end = pos; return 1; }

static int match_node(M *m, const Node *nd, int pos, int dir, const Cont *k) {
    if (++m->steps > m->step_limit) { m->aborted = 1; return 0; }
    switch (nd->type) {
    case N_EMPTY: return k->fn(m, k, pos);
    case N_CHAR: case N_ANY: case N_CLASS: {
        int p = dir > 0 ? pos : pos - 1;
        if (p < 0 || p >= m->n) return 0;
        if (!unit_matches(m, nd, m->s[p])) return 0;
        return k->fn(m, k, pos + dir);
    }
    case N_SEQ: {
        if (nd->nkids == 0) return k->fn(m, k, pos);
        Cont c; c.fn = k_seq; c.next = k; c.node = nd; c.dir = dir; c.b = 0; c.x = 0;
        c.a = dir > 0 ? 0 : nd->nkids - 1;
        return k_seq(m, &c, pos);
    }
    case N_ALT:
        for (int i = 0; i < nd->nkids; i++) {
            if (match_node(m, nd->kids[i], pos, dir, k)) return 1;
            if (m->aborted) return 0;
        }
        return 0;
    case N_GROUP: {
        if (!nd->cap) return match_node(m, nd->kids[0], pos, dir, k);
        Cont c; c.fn = k_group_end; c.next = k; c.node = nd; c.x = pos; c.dir = dir; c.a = c.b = 0;
        return match_node(m, nd->kids[0], pos, dir, &c);
    }
    case N_REPEAT: return rep(m, nd, nd->min, nd->max, pos, dir, k);
    case N_BOL: return pos == 0 ? k->fn(m, k, pos) : 0;
    case N_EOL: return pos == m->n ? k->fn(m, k, pos) : 0;
    case N_WORDB: case N_NWORDB: {
        int a = pos > 0 && isword(m->s[pos - 1]);
        int b = pos < m->n && isword(m->s[pos]);
        int is_b = a != b;
        if ((nd->type == N_WORDB) == is_b) return k->fn(m, k, pos);
        return 0;
    }
    case N_LOOK: {
        Cont acc; acc.fn = k_accept; acc.next = NULL; acc.node = NULL; acc.a = acc.b = acc.x = 0; acc.dir = 0;
        int ncap2 = 2 * m->re->ncap;
        int buf[64]; int *sv = ncap2 > 64 ? (int *)malloc(sizeof(int) * ncap2) : buf;
        memcpy(sv, m->caps, sizeof(int) * ncap2);
        int saved_end = m->end;
        int r = match_node(m, nd->kids[0], pos, nd->behind ? -1 : 1, &acc);
        m->end = saved_end;
        int res = 0;
        if (!m->aborted) {
            if (nd->negate) {
                memcpy(m->caps, sv, sizeof(int) * ncap2);
                if (!r) res = k->fn(m, k, pos);
            } else if (r) {
                res = k->fn(m, k, pos);
                if (!res) memcpy(m->caps, sv, sizeof(int) * ncap2);
            }
        }
        if (sv != buf) free(sv);
        return res;
    }
    case N_BACKREF: {
        int g = nd->ch;
        int s0 = m->caps[2 * g], e0 = m->caps[2 * g + 1];
        if (s0 < 0 || e0 < 0) return k->fn(m, k, pos);
        int len = e0 - s0;
        int f = dir > 0 ? pos + len : pos - len;
        if (f < 0 || f > m->n) return 0;
        int g0 = f < pos ? f : pos;
        for (int i = 0; i < len; i++) if (canon(m, m->s[s0 + i]) != canon(m, m->s[g0 + i])) return 0;
        return k->fn(m, k, f);
    }
    }
    return 0;
}

/* RegExpBuiltinExec (22.2.7.2) without the sticky flag: first index >= last_index
 * at which the pattern matches.  returns 1/0, -1 when the step limit was hit. */
int jsre_exec(const jsre *re, const uint16_t *s, int n, int last_index, int *start, int *end) {
    int capbuf[64];
    int *caps = 2 * re->ncap > 64 ? (int *)malloc(sizeof(int) * 2 * re->ncap) : capbuf;
    M m; m.re = re; m.s = s; m.n = n; m.caps = caps; m.steps = 0; m.step_limit = 200000000L; m.aborted = 0; m.end = -1;
    Cont acc; acc.fn = k_accept; acc.next = NULL; acc.node = NULL; acc.a = acc.b = acc.x = 0; acc.dir = 0;
    int res = 0;
    for (int i = last_index; i <= n; i++) {
        for (int c = 0; c < 2 * re->ncap; c++) caps[c] = -1;
        if (match_node(&m, re->root, i, 1, &acc)) { *start = i; *end = m.end; res = 1; break; }
        if (m.aborted) { res = -1; break; }
    }
    if (caps != capbuf) free(caps);
    return res;
}

int jsre_test(const jsre *re, const uint16_t *s, int n) { int a, b; return jsre_exec(re, s, n, 0, &a, &b); }

/* The global-exec loop of PatternRegistry.findMatches (registry.ts:225-236):
 * collect (start,end) of every match; lastIndex = end, +1 after an empty match.
 * spans = int[2*cap]; returns the number of matches (may exceed cap; extra dropped),
 * or -1 on step-limit abort. */
int jsre_find_all(const jsre *re, const uint16_t *s, int n, int *spans, int cap) {
    int li = 0, cnt = 0;
    while (li <= n) {
        int a, b;
        int r = jsre_exec(re, s, n, li, &a, &b);
        if (r < 0) return -1;
        if (!r) break;
        if (cnt < cap) { spans[2 * cnt] = a; spans[2 * cnt + 1] = b; }
        cnt++;
        li = b;
        if (a == b) li++;
    }
    return cnt;
}
