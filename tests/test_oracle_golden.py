"""The oracle is pinned before it is trusted (CPU only):
  * every vector of gov/test/redaction/registry.test.ts (tests/golden/registry_vectors.json)
  * SHA-256: NIST/FIPS 180-4 vectors, hashlib, and the digests quoted in gov/RFC.md:1561-1562
  * Merkle: level-wise fold == RFC 6962 recursive MTH for every n (tree shape), fold of aligned blocks
  * a second, unrelated engine (Python `re` on UTF-16-unit strings) agrees on the built-in rules
"""
import hashlib
import json
import os
import re

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
VECTORS = json.load(open(os.path.join(HERE, "golden", "registry_vectors.json")))["vectors"]


def run_checks(matches, checks):
    for c in checks:
        if c["op"] == "some":
            assert any(m["id"] in c["ids"] for m in matches) == c["value"], c
        elif c["op"] in ("len_eq", "len_ge"):
            n = len([m for m in matches if "filter_id" not in c or m["id"] == c["filter_id"]])
            assert (n == c["n"]) if c["op"] == "len_eq" else (n >= c["n"]), (c, matches)
        elif c["op"] == "id_at":
            assert matches[c["i"]]["id"] == c["id"], c
        elif c["op"] == "match_at":
            assert matches[c["i"]]["match"] == c["match"], c
        elif c["op"] == "find_defined":
            assert any(m["id"] == c["id"] for m in matches), c


@pytest.mark.parametrize("v", VECTORS, ids=lambda v: "L%d" % v["line"])
def test_oracle_matches_reference_vectors(oracle, v):
    reg = oracle.builtin_registry(v["categories"], v["custom"])
    run_checks(oracle.find_matches(reg, v["input"]), v["checks"])


def test_vector_count_is_complete():
    assert len(VECTORS) >= 125 and sum(len(v["checks"]) for v in VECTORS) >= 140


def test_survey_smoke_vectors(oracle):
    reg = oracle.builtin_registry()
    m = oracle.find_matches(reg, "api_key=sk-abcdefghijklmnopqrstuvwxyz")
    assert [(x["id"], x["start"], x["end"]) for x in m] == [("key-value-credential", 0, 37)]
    # Appendix B.1: identical spans from openai-api-key and generic-api-key -> iteration order wins
    m = oracle.find_matches(reg, "sk-" + "a" * 26)
    assert [x["id"] for x in m] == ["openai-api-key"]
    assert oracle.find_matches(reg, "") == []


def test_invalid_custom_pattern_is_dropped(oracle):
    reg = oracle.builtin_registry(["custom"], [{"name": "bad", "regex": "[invalid", "category": "custom"}])
    assert reg == []


def test_matches_any_semantics(oracle):
    cache = {"secret\\d+": oracle.Regex("secret\\d+")}
    assert oracle.matches_any("secret\\d+", ["my secret42"], cache)
    assert not oracle.matches_any("secret\\d+", ["my secret"], cache)
    assert oracle.matches_any(["nope", "se+cret"], ["my seeecret"], {})          # compiled on the fly
    assert oracle.matches_any("a(b", ["xa(by"], {})                               # syntax error -> includes()
    assert not oracle.matches_any("a(b", ["xaby"], {})


# ---------------------------------------------------------------------------------------- SHA-256

NIST = [
    (b"abc", "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"),
    (b"", "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"),
    (b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq",
     "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"),
    (b"a" * 1000000, "cdc76e5c9914fb9281a1c7e284d73e67f1809a48a497200e046d39ccc7112cd0"),
]


@pytest.mark.parametrize("msg,hexd", NIST, ids=["abc", "empty", "448bit", "million_a"])
def test_sha256_known_answers(oracle, msg, hexd):
    assert oracle.sha256(msg).hex() == hexd


def test_sha256_rfc_md_digests(oracle):
    # gov/RFC.md:1561-1562 quotes these two digests
    assert oracle.sha256_hex_of_js_string("") == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"
    assert oracle.sha256_hex_of_js_string("Hello World!") == "7f83b1657ff1fc53b92dc18148a1d65dfc2d4b1fa3d677284addd200126d9069"


def test_sha256_every_length_against_hashlib(oracle):
    rng = np.random.default_rng(1)
    for n in list(range(0, 200)) + [255, 256, 257, 1000, 4095]:
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.sha256(b) == hashlib.sha256(b).digest()


def test_vault_placeholder(oracle):
    h = hashlib.sha256("s3cr3t-ü".encode()).hexdigest()
    assert oracle.vault_placeholder("s3cr3t-ü", "credential") == "[REDACTED:credential:%s]" % h[:8]
    assert oracle.vault_placeholder("s3cr3t-ü", "pii", long_form=True) == "[REDACTED:pii:%s]" % h[:12]
    # lone surrogate hashes as U+FFFD
    assert oracle.sha256_hex_of_js_string("\ud800") == hashlib.sha256(b"\xef\xbf\xbd").hexdigest()


# ---------------------------------------------------------------------------------------- Merkle

def _leaves(rng, n, lo=0, hi=300):
    msgs = [rng.integers(0, 256, int(rng.integers(lo, hi + 1)), dtype=np.uint8).tobytes() for _ in range(n)]
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(m) for m in msgs])
    return np.frombuffer(b"".join(msgs) + b"\0", dtype=np.uint8).copy(), off, msgs


def test_merkle_small_trees_by_hand(oracle):
    H = lambda b: hashlib.sha256(b).digest()
    rng = np.random.default_rng(2)
    data, off, m = _leaves(rng, 5)
    L = [H(b"\x00" + x) for x in m]
    N = lambda a, b: H(b"\x01" + a + b)
    assert oracle.merkle_root(data[:1], np.zeros(1, dtype=np.uint64)) == H(b"")
    assert oracle.merkle_root(data, off[:2]) == L[0]
    assert oracle.merkle_root(data, off[:3]) == N(L[0], L[1])
    assert oracle.merkle_root(data, off[:4]) == N(N(L[0], L[1]), L[2])
    assert oracle.merkle_root(data, off[:6]) == N(N(N(L[0], L[1]), N(L[2], L[3])), L[4])


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 9, 31, 32, 33, 100, 255, 1000, 1025])
def test_merkle_levelwise_equals_rfc6962_shape(oracle, n):
    rng = np.random.default_rng(n)
    data, off, _ = _leaves(rng, n, 0, 80)
    assert oracle.merkle_root(data, off) == oracle.merkle_root_rfc6962(data, off)


@pytest.mark.parametrize("n,k", [(1000, 4), (1024, 5), (1025, 6), (37, 3), (5, 4)])
def test_merkle_block_fold_composes(oracle, n, k):
    """roots of aligned 2^k-leaf blocks folded with the same rule == the root (multi-GPU sharding rule)."""
    rng = np.random.default_rng(n * 31 + k)
    data, off, _ = _leaves(rng, n, 1, 64)
    bs = 1 << k
    roots = []
    for s in range(0, n, bs):
        e = min(n, s + bs)
        sub_off = (off[s:e + 1] - off[s]).astype(np.uint64)
        roots.append(np.frombuffer(oracle.merkle_root(data[int(off[s]):], sub_off), dtype=np.uint8))
    assert oracle.merkle_fold(np.stack(roots)) == oracle.merkle_root(data, off)


def test_merkle_fixed_equals_variable(oracle):
    rng = np.random.default_rng(5)
    n, ll = 300, 64
    data = rng.integers(0, 256, n * ll + 1, dtype=np.uint8)
    off = (np.arange(n + 1, dtype=np.uint64) * ll)
    assert oracle.merkle_root_fixed(data, ll, n) == oracle.merkle_root(data, off)


# ------------------------------------------------------------------ second opinion: Python `re`

PY_WS = "\t\n\x0b\x0c\r \xa0  -     　﻿"


def js_to_py(src: str) -> str:
    """Enough of a JS->Python translation for the 17 built-ins (ASCII \\d \\w \\b via re.ASCII)."""
    out, i, in_class = [], 0, False
    while i < len(src):
        c = src[i]
        if c == "\\" and i + 1 < len(src):
            n = src[i + 1]
            if n == "s":
                out.append(PY_WS if in_class else "[" + PY_WS + "]")
                i += 2
                continue
            out.append(src[i:i + 2])
            i += 2
            continue
        if c == "[":
            in_class = True
        elif c == "]":
            in_class = False
        out.append(c)
        i += 1
    return "".join(out)


def test_python_re_agrees_on_builtins(oracle):
    from vainplex_openclaw_b200 import workload as W
    rules = W.make_rules(17)
    data, off, _ = W.make_messages(1500, 192, rules, p_hit=0.3, utf8_frac=0.0, seed=77)
    buf = data.numpy()
    msgs = [bytes(buf[int(off[i]):int(off[i + 1])]) for i in range(1500)]
    for r in rules:
        rx = oracle.Regex(r["source"], "i" if r["flags"] else "")
        py = re.compile(js_to_py(r["source"]), re.ASCII | (re.IGNORECASE if r["flags"] else 0))
        d, o = oracle.pack(msgs)
        got = oracle.find_matches_batch([(rx, "custom")], d, o)
        exp = []
        for mi, m in enumerate(msgs):
            s = m.decode("utf-8", "replace")
            exp += [(mi, 0, x.start(), x.end()) for x in py.finditer(s)]
        assert [tuple(x) for x in got.tolist()] == exp, r["id"]


def test_merkle_audit_paths_of_the_oracle(oracle):
    """RFC 6962 2.1.1 paths restated in the oracle verify against the oracle's root for every leaf of every tree size
    1..40 (RFC 9162 2.1.3.2), have the expected lengths, and reject a wrong leaf / index."""
    import numpy as np
    rng = np.random.default_rng(11)
    leaves = [bytes(rng.integers(0, 256, int(rng.integers(0, 90)), dtype=np.uint8)) for _ in range(40)]
    for n in range(1, 41):
        ls = leaves[:n]
        off = np.zeros(n + 1, dtype=np.uint64); off[1:] = np.cumsum([len(x) for x in ls])
        data = np.frombuffer(b"".join(ls) + b"\0" * 64, dtype=np.uint8).copy()
        root = oracle.merkle_root(data, off)
        assert root == oracle.merkle_root_rfc6962(data, off)
        for i in range(n):
            p = oracle.merkle_audit_path(ls, i)
            assert len(p) <= max(1, (n - 1).bit_length())
            assert oracle.merkle_verify_path(ls[i], i, n, p, root)
            assert not oracle.merkle_verify_path(ls[i] + b"!", i, n, p, root)
            if n > 1:
                assert not oracle.merkle_verify_path(ls[i], (i + 1) % n, n, p, root) or ls[i] == ls[(i + 1) % n]


# ------------------------------------------------------------------- RFC 6962 known answers (CT reference vectors)

def _ct():
    import json
    v = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ct_merkle_vectors.json")))
    return v, [bytes.fromhex(x) for x in v["leaves_hex"]]


def test_merkle_ct_reference_roots(oracle):
    """The oracle's tree convention pinned on the published Certificate Transparency test vectors: the root over the
    first k of the 8 fixed leaves, k = 1..8, from both the level-wise builder and the recursive RFC 6962 restatement."""
    v, leaves = _ct()
    for k in range(1, 9):
        off = np.zeros(k + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(x) for x in leaves[:k]])
        data = np.frombuffer(b"".join(leaves[:k]) + b"\0" * 64, dtype=np.uint8).copy()
        assert oracle.merkle_root(data, off).hex() == v["roots_hex"][k - 1]
        assert oracle.merkle_root_rfc6962(data, off).hex() == v["roots_hex"][k - 1]
    assert oracle.merkle_root(np.zeros(64, dtype=np.uint8), np.zeros(1, dtype=np.uint64)).hex() == \
        "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"                # RFC 6962 2.1: MTH({}) = SHA-256()


def test_merkle_ct_reference_proofs(oracle):
    v, leaves = _ct()
    roots = [bytes.fromhex(x) for x in v["roots_hex"]]
    for c in v["inclusion"]:
        p = oracle.merkle_audit_path(leaves[:c["size"]], c["index"])
        assert [x.hex() for x in p] == c["path_hex"]
        assert oracle.merkle_verify_path(leaves[c["index"]], c["index"], c["size"], p, roots[c["size"] - 1])
    for c in v["consistency"]:
        p = oracle.merkle_consistency_proof(leaves[:c["second"]], c["first"])
        assert [x.hex() for x in p] == c["path_hex"]
        assert oracle.merkle_verify_consistency(c["first"], c["second"], roots[c["first"] - 1], roots[c["second"] - 1], p)


def test_merkle_consistency_every_pair_small(oracle):
    rng = np.random.default_rng(9162)
    leaves = [bytes(rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8)) for _ in range(41)]
    hs = [oracle.merkle_leaf_hash(x) for x in leaves]
    roots = [None] + [oracle._mth(hs[:k]) for k in range(1, 42)]
    for n in range(1, 42):
        for m in range(1, n + 1):
            p = oracle.merkle_consistency_proof(leaves[:n], m)
            assert oracle.merkle_verify_consistency(m, n, roots[m], roots[n], p)
            if m < n:
                assert not oracle.merkle_verify_consistency(m, n, roots[m], roots[n - 1] if n - 1 >= 1 and n - 1 != m else roots[n][::-1], p)
                assert not oracle.merkle_verify_consistency(m, n, roots[m][::-1], roots[n], p)
                assert not oracle.merkle_verify_consistency(m, n, roots[m], roots[n], p[:-1])
                assert not oracle.merkle_verify_consistency(m, n, roots[m], roots[n], p + [p[-1]])


# ------------------------------------------------------- language packs (flag i on non-ASCII literals; CJK; Hangul)

def _cortex():
    import json
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cortex_pack_vectors.json"), encoding="utf-8"))


def test_cortex_language_pack_vectors(oracle):
    """Every anyMatch assertion of openclaw-cortex test/patterns-lang-*.test.ts (100 of them, 8 languages) on the packs'
    own regex literals (src/patterns/lang-*.ts): pins Canonicalize for non-ASCII units (Cyrillic, Latin-1) in the oracle."""
    c = _cortex()
    assert len(c["vectors"]) >= 100 and sum(len(p) for cats in c["packs"].values() for p in cats.values()) >= 70
    cache = {}
    for v in c["vectors"]:
        pats = c["packs"][v["lang"]][v["category"]]
        got = False
        for p in pats:
            key = (p["source"], "i" in p["flags"])
            rx = cache.get(key) or cache.setdefault(key, oracle.Regex(p["source"], "i" if "i" in p["flags"] else ""))
            data, off = oracle.pack([oracle.js_to_utf8(v["text"])])
            _, words = oracle.scan_policy([rx], data, off, threads=1, want_bits=False)
            got = got or bool(words[0])
        assert got == v["expect"], (v["lang"], v["category"], v["text"], v["line"])
    # a second opinion on case folding: Python's re with IGNORECASE on the same vectors
    import re
    for v in c["vectors"]:
        pats = c["packs"][v["lang"]][v["category"]]
        try:
            got = any(re.search(p["source"], v["text"], re.I if "i" in p["flags"] else 0) for p in pats)
        except re.error:
            continue
        assert got == v["expect"], ("python re", v["lang"], v["text"])


def test_napi_shim_compiles_against_declarations():
    """napi/openclaw_gov_napi.c (the reference-side binding; Node.js is absent here) type-checks against N-API declarations
    (napi/stub/node_api.h) and the C ABI header, without warnings -- and binds every entry point a JS caller needs."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I" + os.path.join(root, "napi", "stub"), "-I" + os.path.join(root, "include"),
                        os.path.join(root, "napi", "openclaw_gov_napi.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(root, "napi", "openclaw_gov_napi.c")).read()
    for fn in ("cg_init", "cg_shutdown", "cg_get_stats", "cg_rule_check", "cg_ruleset_create", "cg_scan_batch", "cg_scan_one", "cg_find_matches_batch",
               "cg_redact_batch", "cg_ruleset_set_policy", "cg_policy_verdict_batch", "cg_sha256_batch", "cg_merkle_root", "cg_merkle_log_create",
               "cg_merkle_log_restore", "cg_merkle_log_append", "cg_merkle_log_append_jsonl", "cg_merkle_log_reserve", "cg_merkle_log_size", "cg_merkle_log_root",
               "cg_merkle_log_frontier", "cg_merkle_log_proof", "cg_merkle_log_consistency", "cg_merkle_verify_proof", "cg_merkle_verify_consistency"):
        assert fn + "(" in src, fn
