"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU oracle for the openclaw-governance hot path: a ctypes front-end over
oracle/_build/liboracle.so (jsre.c = ECMA-262 backtracking RegExp over UTF-16 units,
scan_batch.c = the reference's two scan loops, sha256_merkle.c = FIPS 180-4 SHA-256
and this repository's Merkle convention).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg may import this module; the product
package (vainplex_openclaw_b200) never does.

Reference lines restated here:
  findMatches / resolveOverlaps  gov/src/redaction/registry.ts:212-242, 288-316
  CATEGORY_ORDER                  gov/src/redaction/registry.ts:17-22
  BUILTIN_PATTERNS                gov/src/redaction/registry.ts:31-151 (data)
  matchesAny                      gov/src/conditions/context.ts:9-25
  vault placeholder               gov/src/redaction/vault.ts:26-35,75-104
Parity status: scan = pinned by tests/golden/registry_vectors.json (extracted from
gov/test/redaction/registry.test.ts); SHA-256 = pinned by FIPS/NIST vectors + hashlib
(the reference tests hold no digest constants); Merkle = PARITY UNPINNED (no reference
implementation exists; convention frozen in sha256_merkle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

CATEGORY_ORDER = ("credential", "financial", "pii", "custom")

# (id, category, source, flags) in table order -- registry.ts:31-151
BUILTIN_PATTERNS = (
    ("openai-api-key", "credential", r"sk-[a-zA-Z0-9]{20,}", ""),
    ("anthropic-api-key", "credential", r"sk-ant-[a-zA-Z0-9-]{80,}", ""),
    ("aws-key", "credential", r"(?<![A-Z0-9])AKIA[0-9A-Z]{16}(?![A-Z0-9])", ""),
    ("generic-api-key", "credential", r"sk-[a-zA-Z0-9_-]{20,}", ""),
    ("google-api-key", "credential", r"AIza[0-9A-Za-z_-]{35}", ""),
    ("github-pat", "credential", r"ghp_[a-zA-Z0-9]{36}", ""),
    ("github-server-token", "credential", r"ghs_[a-zA-Z0-9]{36}", ""),
    ("gitlab-pat", "credential", r"glpat-[a-zA-Z0-9_-]{20,}", ""),
    ("private-key-header", "credential", r"-----BEGIN (?:RSA |EC |OPENSSH )?PRIVATE KEY-----", ""),
    ("bearer-token", "credential", r"Bearer [a-zA-Z0-9_./-]{20,}", ""),
    ("basic-auth", "credential", r"Basic [A-Za-z0-9+/]{16,}={0,2}", ""),
    ("key-value-credential", "credential",
     r"""(?:password|passwd|pwd|secret|token|api_key|apikey)\s*[:=]\s*['"]?[^\s'"]{8,64}""", "i"),
    ("email-address", "pii", r"\b[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-zA-Z]{2,}\b", ""),
    ("phone-number", "pii", r"(?<!\d)\+?[1-9]\d{6,14}(?!\d)", ""),
    ("ssn-us", "pii", r"\b\d{3}-\d{2}-\d{4}\b", ""),
    ("credit-card", "financial", r"\b[45]\d{3}[\s-]?\d{4}[\s-]?\d{4}[\s-]?\d{4}\b", ""),
    ("iban", "financial", r"\b[A-Z]{2}\d{2}\s?[A-Z0-9]{4}\s?(?:\d{4}\s?){2,7}\d{1,4}\b", ""),
)


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("jsre.c", "scan_batch.c", "sha256_merkle.c", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.jsre_compile.restype = C.c_void_p
        L.jsre_compile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.jsre_free.argtypes = [C.c_void_p]
        L.oracle_scan_policy.restype = C.c_int
        L.oracle_scan_policy.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_find_matches_batch.restype = C.c_long
        L.oracle_find_matches_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_sha256.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_sha256_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_merkle_root.restype = C.c_int
        L.oracle_merkle_root.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_merkle_root_fixed.restype = C.c_int
        L.oracle_merkle_root_fixed.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.oracle_merkle_root_fixed_mt.restype = C.c_int
        L.oracle_merkle_root_fixed_mt.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
        L.oracle_merkle_root_rfc6962.restype = C.c_int
        L.oracle_merkle_root_rfc6962.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_merkle_fold.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


class RegexSyntaxError(ValueError):
    """What `new RegExp(src)` would throw as SyntaxError."""


class OracleUnsupported(ValueError):
    """Valid JS the oracle itself cannot evaluate exactly (flag i with non-ASCII literals)."""


def js_units(s: str) -> np.ndarray:
    """A JS string as its UTF-16 code units."""
    return np.frombuffer(s.encode("utf-16-le", "surrogatepass"), dtype=np.uint16)


def js_to_utf8(s: str) -> bytes:
    """What Buffer.from(str,'utf8') / TextEncoder produce: lone surrogates -> U+FFFD."""
    return s.encode("utf-16-le", "surrogatepass").decode("utf-16-le", "replace").encode("utf-8")


class Regex:
    """One compiled JS RegExp (source + flags subset 'i'; 'g' is implied by the callers)."""

    def __init__(self, source: str, flags: str = ""):
        u = js_units(source)
        err = C.create_string_buffer(256)
        buf = u.ctypes.data_as(C.c_void_p) if len(u) else None
        self.handle = lib().jsre_compile(buf, len(u), 1 if "i" in flags else 0, err, 256)
        self.source, self.flags = source, flags
        if not self.handle:
            msg = err.value.decode() or "invalid regular expression"
            if msg.startswith("oracle:"):
                raise OracleUnsupported(msg)
            raise RegexSyntaxError(msg)

    def __del__(self):
        h = getattr(self, "handle", None)
        if h and _lib is not None:
            _lib.jsre_free(h)
            self.handle = None


def pack(messages: Sequence[bytes]):
    """bytes + offsets[n+1] (uint64) layout used by every batch entry point."""
    off = np.zeros(len(messages) + 1, dtype=np.uint64)
    if len(messages):
        off[1:] = np.cumsum([len(m) for m in messages], dtype=np.uint64)
    data = np.frombuffer(b"".join(messages) + b"\0", dtype=np.uint8).copy()
    return data, off


def _handles(regexes: Sequence[Regex]):
    arr = (C.c_void_p * max(1, len(regexes)))(*[r.handle for r in regexes])
    return arr


def scan_policy(regexes: Sequence[Regex], data: np.ndarray, off: np.ndarray, threads: int = 0,
                want_bits: bool = True):
    """matchesAny semantics per (message, rule): returns (hit_bits[n, ceil(R/8)], words[n])."""
    n = len(off) - 1
    R = len(regexes)
    threads = threads or host_threads()
    bits = np.zeros((n, (R + 7) // 8), dtype=np.uint8) if want_bits else None
    words = np.zeros(n, dtype=np.uint64)
    rc = lib().oracle_scan_policy(_handles(regexes), R, data.ctypes.data, off.ctypes.data, n, threads,
                                  bits.ctypes.data if want_bits else None, words.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle step limit hit")
    return bits, words


def find_matches_batch(patterns: Sequence[tuple], data: np.ndarray, off: np.ndarray, threads: int = 0):
    """PatternRegistry.findMatches over a batch.

    patterns: (Regex, category) in *storage* order (built-ins then custom); iteration order
    (CATEGORY_ORDER, stable within category) is applied here exactly like registry.ts:216-220.
    Returns int32 array [k,4] of (msg, storage_index, start16, end16), message-major."""
    order = [i for c in CATEGORY_ORDER for i, p in enumerate(patterns) if p[1] == c]
    regs = [patterns[i][0] for i in order]
    ranks = np.array([CATEGORY_ORDER.index(patterns[i][1]) for i in order], dtype=np.int32)
    n = len(off) - 1
    threads = threads or host_threads()
    out = C.c_void_p()
    k = lib().oracle_find_matches_batch(_handles(regs), ranks.ctypes.data, len(regs), data.ctypes.data,
                                        off.ctypes.data, n, threads, C.byref(out))
    try:
        if k < 0:
            raise RuntimeError("oracle step limit hit")
        arr = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int32)), shape=(max(k, 1), 4))[:k].copy()
    finally:
        lib().oracle_free(out)
    if k:
        arr[:, 1] = np.array(order, dtype=np.int32)[arr[:, 1]]
    return arr


def builtin_registry(categories: Iterable[str] = ("credential", "pii", "financial"),
                     custom: Sequence[dict] = ()):
    """new PatternRegistry(categories, custom): list of (Regex, category, id) in storage order.
    Custom patterns that fail to compile are dropped (registry.ts:272-279); the
    timing-based ReDoS rejection (registry.ts:253-265) is non-deterministic and not restated."""
    enabled = set(categories)
    pats = []
    for pid, cat, src, fl in BUILTIN_PATTERNS:
        if cat in enabled:
            pats.append((Regex(src, fl), cat, pid))
    for cp in custom:
        try:
            pats.append((Regex(cp["regex"], ""), cp["category"], "custom-" + cp["name"]))
        except RegexSyntaxError:
            pass
    return pats


def find_matches(patterns: Sequence[tuple], text: str):
    """findMatches(text) -> list of dict(id, category, match, start, end) (UTF-16 offsets)."""
    data, off = pack([js_to_utf8(text)])
    spans = find_matches_batch([(p[0], p[1]) for p in patterns], data, off, threads=1)
    u = js_units(text)
    out = []
    for _, pi, s, e in spans.tolist():
        m = u[s:e].tobytes().decode("utf-16-le", "surrogatepass")
        out.append({"id": patterns[pi][2], "category": patterns[pi][1], "match": m, "start": s, "end": e})
    return out


def matches_any(patterns, texts: Sequence[str], regex_cache: dict) -> bool:
    """conditions/context.ts:9-25."""
    lst = patterns if isinstance(patterns, (list, tuple)) else [patterns]
    for p in lst:
        re_ = regex_cache.get(p)
        if re_ is None:
            try:
                re_ = Regex(p, "")
            except RegexSyntaxError:
                if any(p in t for t in texts):
                    return True
                continue
        for t in texts:
            data, off = pack([js_to_utf8(t)])
            _, words = scan_policy([re_], data, off, threads=1, want_bits=False)
            if words[0]:
                return True
    return False


# ---------------------------------------------------------------- SHA-256 / Merkle

def sha256(data: bytes) -> bytes:
    out = np.zeros(32, dtype=np.uint8)
    buf = np.frombuffer(data + b"\0", dtype=np.uint8)
    lib().oracle_sha256(buf.ctypes.data, len(data), out.ctypes.data)
    return out.tobytes()


def sha256_hex_of_js_string(s: str) -> str:
    """createHash('sha256').update(str).digest('hex') -- util.ts:77-79."""
    return sha256(js_to_utf8(s)).hex()


def vault_placeholder(original: str, category: str, long_form: bool = False) -> str:
    """vault.ts:33-35,75-104: [REDACTED:<category>:<hash8|hash12>]."""
    h = sha256_hex_of_js_string(original)
    return "[REDACTED:%s:%s]" % (category, h[:12] if long_form else h[:8])


def sha256_batch(data: np.ndarray, off: np.ndarray) -> np.ndarray:
    n = len(off) - 1
    out = np.zeros((n, 32), dtype=np.uint8)
    lib().oracle_sha256_batch(data.ctypes.data, off.ctypes.data, n, out.ctypes.data)
    return out


def merkle_root(data: np.ndarray, off: np.ndarray) -> bytes:
    out = np.zeros(32, dtype=np.uint8)
    if lib().oracle_merkle_root(data.ctypes.data, off.ctypes.data, len(off) - 1, out.ctypes.data) != 0:
        raise MemoryError
    return out.tobytes()


def merkle_root_fixed(data: np.ndarray, leaf_len: int, n: int, threads: int = 1) -> bytes:
    out = np.zeros(32, dtype=np.uint8)
    if threads > 1:
        rc = lib().oracle_merkle_root_fixed_mt(data.ctypes.data, leaf_len, n, threads, out.ctypes.data)
    else:
        rc = lib().oracle_merkle_root_fixed(data.ctypes.data, leaf_len, n, out.ctypes.data)
    if rc != 0:
        raise MemoryError
    return out.tobytes()


def host_threads() -> int:
    """CPUs this process may run on (its affinity mask -- a cgroup-limited container often has fewer than os.cpu_count())"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def merkle_root_rfc6962(data: np.ndarray, off: np.ndarray) -> bytes:
    out = np.zeros(32, dtype=np.uint8)
    lib().oracle_merkle_root_rfc6962(data.ctypes.data, off.ctypes.data, len(off) - 1, out.ctypes.data)
    return out.tobytes()


def merkle_fold(nodes: np.ndarray) -> bytes:
    """Fold an [m,32] array of subtree roots (equal-size aligned blocks) into the root."""
    buf = np.ascontiguousarray(nodes, dtype=np.uint8).copy()
    out = np.zeros(32, dtype=np.uint8)
    lib().oracle_merkle_fold(buf.ctypes.data, buf.shape[0], out.ctypes.data)
    return out.tobytes()


# ---------------------------------------------------------------- Merkle log: audit paths (RFC 6962 2.1.1 / RFC 9162 2.1.3.2)
# Restated from the RFC text on top of the oracle's own SHA-256; small inputs only (pure Python recursion).

def merkle_leaf_hash(leaf: bytes) -> bytes:
    return sha256(b"\x00" + leaf)


def merkle_node_hash(left: bytes, right: bytes) -> bytes:
    return sha256(b"\x01" + left + right)


def _mth(hashes):
    """MTH over a list of leaf hashes, RFC 6962 2.1: split at the largest power of two smaller than n."""
    n = len(hashes)
    if n == 1:
        return hashes[0]
    k = 1
    while k * 2 < n:
        k *= 2
    return merkle_node_hash(_mth(hashes[:k]), _mth(hashes[k:]))


def merkle_audit_path(leaves, index: int):
    """PATH(m, D[n]) of RFC 6962 2.1.1 (leaf level first)."""
    def path(m, hs):
        n = len(hs)
        if n == 1:
            return []
        k = 1
        while k * 2 < n:
            k *= 2
        if m < k:
            return path(m, hs[:k]) + [_mth(hs[k:])]
        return path(m - k, hs[k:]) + [_mth(hs[:k])]
    return path(index, [merkle_leaf_hash(x) for x in leaves])


def merkle_verify_path(leaf: bytes, index: int, tree_size: int, path, root: bytes) -> bool:
    """RFC 9162 2.1.3.2."""
    if index >= tree_size:
        return False
    fn, sn, r = index, tree_size - 1, merkle_leaf_hash(leaf)
    for p in path:
        if sn == 0:
            return False
        if (fn & 1) or fn == sn:
            r = merkle_node_hash(p, r)
            if not (fn & 1):
                while not (fn & 1) and fn != 0:
                    fn >>= 1
                    sn >>= 1
        else:
            r = merkle_node_hash(r, p)
        fn >>= 1
        sn >>= 1
    return sn == 0 and r == root


def merkle_consistency_proof(leaves, first_size: int):
    """PROOF(m, D[n]) of RFC 6962 2.1.2."""
    hs = [merkle_leaf_hash(x) for x in leaves]

    def sub(m, d, b):
        n = len(d)
        if m == n:
            return [] if b else [_mth(d)]
        k = 1
        while k * 2 < n:
            k *= 2
        if m <= k:
            return sub(m, d[:k], b) + [_mth(d[k:])]
        return sub(m - k, d[k:], False) + [_mth(d[:k])]
    if not 0 < first_size <= len(hs):
        raise ValueError("first_size out of range")
    return sub(first_size, hs, True)


def merkle_verify_consistency(first_size: int, second_size: int, root_first: bytes, root_second: bytes, path) -> bool:
    """RFC 9162 2.1.4.2."""
    path = list(path)
    if first_size == 0 or first_size > second_size:
        return False
    if first_size == second_size:
        return not path and root_first == root_second
    if not path:
        return False
    if first_size & (first_size - 1) == 0:
        path = [root_first] + path
    fn, sn = first_size - 1, second_size - 1
    while fn & 1:
        fn >>= 1
        sn >>= 1
    fr = sr = path[0]
    for c in path[1:]:
        if sn == 0:
            return False
        if (fn & 1) or fn == sn:
            fr = merkle_node_hash(c, fr)
            sr = merkle_node_hash(c, sr)
            while not (fn & 1) and fn != 0:
                fn >>= 1
                sn >>= 1
        else:
            sr = merkle_node_hash(sr, c)
        fn >>= 1
        sn >>= 1
    return sn == 0 and fr == root_first and sr == root_second
