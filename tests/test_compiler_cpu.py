"""CPU tier for the product's host logic: the rule compiler, the prefilter tables and the Pike-VM
*source* (compiled for the host by tests/native/vm_harness.cpp) against the oracle.  No GPU needed;
the GPU tier (test_gpu_parity.py) repeats the comparisons through the C ABI and the real kernels."""
import json
import os
import re

import numpy as np
import pytest

from helpers import Harness, PACK_RULES, PACK_TEXTS, random_regex, oracle_regexes
from vainplex_openclaw_b200 import workload as W

HERE = os.path.dirname(os.path.abspath(__file__))
VECTORS = json.load(open(os.path.join(HERE, "golden", "registry_vectors.json")))["vectors"]


def oracle_spans(O, rx, msgs):
    d, o = O.pack(msgs)
    got = O.find_matches_batch([(rx, "custom")], d, o)
    out = {}
    for mi, _, s, e in got.tolist():
        out.setdefault(mi, []).append((s, e))
    return out


def u16_to_byte_offsets(msg: bytes):
    """map UTF-16 index -> byte offset for a valid UTF-8 message (tests only)."""
    s = msg.decode("utf-8")
    m, b = [0], 0
    for ch in s:
        enc = ch.encode("utf-8")
        if len(enc) == 4:
            m.append(b + 2)
            b += 4
            m.append(b)
        else:
            b += len(enc)
            m.append(b)
    return m


def test_builtins_on_reference_vectors(oracle, harness_lib):
    rules = W.rules_as_tuples(W.make_rules(17))
    h = Harness(harness_lib, rules)
    assert (h.status == 0).all()
    regs = [oracle.Regex(r[0], "i" if r[1] else "") for r in rules]
    msgs = [oracle.js_to_utf8(v["input"]) for v in VECTORS]
    for ri, rx in enumerate(regs):
        exp = oracle_spans(oracle, rx, msgs)
        for mi, m in enumerate(msgs):
            got = h.find_all(ri, m)
            assert [(g[2], g[3]) for g in got] == exp.get(mi, []), (rules[ri][0], m)
            if exp.get(mi):
                assert ri in h.candidates(m), ("prefilter missed", rules[ri][0], m)
                assert h.test(ri, m)
            else:
                assert not h.test(ri, m)
    h.close()


@pytest.mark.parametrize("stride", [0, 2, 4])
@pytest.mark.parametrize("utf8_frac", [0.0, 0.15])
def test_synthetic_rules_spans_and_prefilter_soundness(oracle, harness_lib, stride, utf8_frac):
    """gram filter (every stride; the message at every alignment inside a buffer of other bytes, and at the very start
    of the scanned range: the head check) -> exact factors -> VM, against the oracle"""
    rl = W.make_rules(160)
    rules = W.rules_as_tuples(rl)
    h = Harness(harness_lib, rules, stride=stride)
    assert (h.status == 0).all(), [h.L.harness_rule_error(h.h, i) for i in np.nonzero(h.status)[0]]
    assert h.info()["stride"] == (stride or 2)          # the built-ins hold factors of 2-3 elements: stride 4 cannot cover them
    data, off, inj = W.make_messages(1200, 200, rl, p_hit=0.25, utf8_frac=utf8_frac, seed=1234 + stride)
    buf = data.numpy()
    msgs = [bytes(buf[int(off[i]):int(off[i + 1])]) for i in range(1200)]
    n_hits = 0
    c2 = [h.candidates2(m, lead=(0, 1, 2, 3, 13, 16, 18, 23)[i % 8], seed=i) for i, m in enumerate(msgs)]
    cands = [c[0] | c[1] for c in c2]
    direct = [c[1] for c in c2]
    for ri, r in enumerate(rules):
        exp = oracle_spans(oracle, oracle.Regex(r[0], "i" if r[1] else ""), msgs)
        for mi, m in enumerate(msgs):
            e = exp.get(mi, [])
            if e or ri in cands[mi]:
                got = h.find_all(ri, m)
                assert [(g[2], g[3]) for g in got] == e, (r[0], m)
                if e:
                    n_hits += 1
                    assert ri in cands[mi], ("prefilter missed", r[0], m)
                else:
                    assert ri not in direct[mi], ("direct hit without a match", r[0], m)
    assert n_hits >= 200          # the injected tokens were really found
    # policy semantics through the pipeline's own logic (level 1 -> confirm -> island-restricted VM)
    hits_by_msg = {}
    for ri, r in enumerate(rules):
        for mi in oracle_spans(oracle, oracle.Regex(r[0], "i" if r[1] else ""), msgs):
            hits_by_msg.setdefault(mi, set()).add(ri)
    for mi, m in enumerate(msgs):
        assert h.policy_hits(m, lead=mi % 21, seed=mi) == hits_by_msg.get(mi, set()), (mi, m)
    h.close()


def test_byte_offsets_match_utf16_offsets(oracle, harness_lib):
    rules = [(r"\d+", 0, 3), (r"[^\s]+@[a-z]+", 0, 3), (r".", 0, 3), (r"\b\w+\b", 0, 3), (r"", 0, 3), (r"x*", 0, 3)]
    h = Harness(harness_lib, rules)
    texts = ["héllo 123 wörld 45", "😀a@b 😀😀 c@d", "日本語 12 テスト", "", "a", "😀", "ab cd 7"]
    for t in texts:
        m = t.encode("utf-8")
        mp = u16_to_byte_offsets(m)
        for ri, r in enumerate(rules):
            exp = oracle_spans(oracle, oracle.Regex(r[0]), [m]).get(0, [])
            got = h.find_all(ri, m)
            assert [(g[2], g[3]) for g in got] == exp, (r[0], t)
            assert [(g[0], g[1]) for g in got] == [(mp[s], mp[e]) for s, e in exp], (r[0], t)
    h.close()


SYNTAX_ERRORS = ["a(b", "[abc", "a)", "*a", "a**", "a{2,1}", "(?<n", "x{3}{2}", "\\", "(?", "[b-a]", "+"]
UNSUPPORTED = [r"(a)\1", r"(?=ab)c", r"(?<!ab)c", r"(?<n>a)\k<n>", r"(?=a)*b"]
VALID = ["a{", "a{,3}", "x{1,2}y", "[]a]?", "[^]", r"\c", r"é+", r"a|", "()", "(?:)", r"[\d-x]", r"\x4", r"\08", "}{", "]"]


def test_syntax_and_support_classification(oracle, harness_lib):
    from vainplex_openclaw_b200 import _native as N
    for s in SYNTAX_ERRORS:
        assert N.rule_check(s) == N.CG_ERR_SYNTAX, s
        with pytest.raises(oracle.RegexSyntaxError):
            oracle.Regex(s)
    for s in UNSUPPORTED:
        assert N.rule_check(s) == N.CG_ERR_UNSUPPORTED, s
        oracle.Regex(s)                       # valid JS: the oracle accepts it
    for s in VALID:
        assert N.rule_check(s) == 0, s
        oracle.Regex(s)
    assert N.rule_check("[a-z]{2000}") == N.CG_ERR_TOO_LARGE
    assert N.rule_check("É", N.FLAG_ICASE) == 0          # (round 1: unsupported; Canonicalize now covers the BMP, test_cortex_language_packs_*)


def test_closure_stack_is_bounded_at_compile_time(oracle, harness_lib):
    """A rule either provably fits the VM's closure stack or is rejected when it is compiled (it used to pass the compile
    step and then fail every batch at scan time).  An alternation keeps one pending branch per alternative on that stack."""
    import itertools, time
    from vainplex_openclaw_b200 import _native as N
    words = ["".join(p) for p in itertools.product("abcdefghijklmnopqrst", repeat=2)]        # 400 two-letter alternatives
    for n_alt, expect_ok in ((60, True), (150, True), (190, True), (200, False), (240, False)):
        src = "|".join(words[:n_alt])
        rc = N.rule_check(src)
        assert (rc == 0) == expect_ok and rc in (0, N.CG_ERR_TOO_LARGE), (n_alt, rc)
        if expect_ok:
            h = Harness(harness_lib, [(src, 0, 3)])
            assert h.status[0] == 0
            last = words[n_alt - 1].encode()
            msgs = [b"xx " + last + b" yy", b"ab0", b"zz zz zz", words[7].encode() + b"7"]
            exp = oracle_spans(oracle, oracle.Regex(src), msgs)
            for mi, m in enumerate(msgs):
                assert h.test(0, m) == bool(exp.get(mi)), (n_alt, m)          # (h.test asserts that the VM did not overflow)
            h.close()
    # an empty body repeated two billion times must not stall the compiler (it emits nothing: once is enough)
    t0 = time.time()
    for src in ("(?:){2000000000}", "(){2000000000}a", "(?:(?:){2000000000}){2000000000}b"):
        assert N.rule_check(src) in (0, N.CG_ERR_TOO_LARGE, N.CG_ERR_UNSUPPORTED)
    assert time.time() - t0 < 2.0


def test_random_regex_differential(oracle, harness_lib):
    """product compiler + VM vs oracle on random patterns / texts over a tiny alphabet (dense matches,
    empty matches, astral characters, lazy quantifiers, lookarounds)."""
    rng = np.random.default_rng(20260921)
    alphabet = ["a", "b", "c", "1", " ", "é", "😀", "\n", "A"]
    texts = ["".join(alphabet[int(k)] for k in rng.integers(0, len(alphabet), int(rng.integers(0, 24)))) for _ in range(60)]
    texts += ["", "a", "aaaa", "abcabc", "😀", "a😀b", "1 a1"]
    msgs = [t.encode("utf-8") for t in texts]
    checked = 0
    for it in range(700):
        src = random_regex(rng)
        fl = 1 if rng.random() < 0.2 else 0
        try:
            rx = oracle.Regex(src, "i" if fl else "")
        except oracle.RegexSyntaxError:
            rx = None
        except oracle.OracleUnsupported:
            continue
        h = Harness(harness_lib, [(src, fl, 3)])
        st = int(h.status[0])
        if rx is None:
            assert st == -1, ("oracle says syntax error, product says", st, src)
            h.close()
            continue
        assert st != -1, ("product says syntax error, oracle accepts", src, harness_lib.harness_rule_error(h.h, 0))
        if st != 0:
            h.close()
            continue
        try:
            exp = oracle_spans(oracle, rx, msgs)
        except RuntimeError:          # exponential backtracking in the oracle: not a usable vector
            h.close()
            continue
        for mi, m in enumerate(msgs):
            got = [(g[2], g[3]) for g in h.find_all(0, m)]
            assert got == exp.get(mi, []), (src, fl, texts[mi])
            if exp.get(mi):
                assert 0 in h.candidates(m), ("prefilter missed", src, texts[mi])
            assert (0 in h.policy_hits(m)) == bool(exp.get(mi)), ("policy pipeline differs", src, fl, texts[mi])
        checked += 1
        h.close()
    assert checked > 400


def test_gram_filter_geometry_and_alignment_independence(harness_lib):
    """What the compiler promises the scan kernel: stride 4 only when every factor is coverable at four alignments;
    short factors and key budgets move a set to stride 2; the result never depends on where a message sits in the buffer,
    on the bitmap size or on the stride."""
    # long literals only: stride 4, every gram inside its factor (no wildcard bytes -> few keys)
    lits = [("alphabetagamma", 0, 3), ("thequickbrownfox", 0, 3), ("zeppelin_airship", 1, 3)]
    h4 = Harness(harness_lib, lits)
    i4 = h4.info()
    assert i4["stride"] == 4 and i4["entries"] == 4 * i4["n_factors"] and i4["keys"] <= 4 * i4["n_factors"] and i4["triggers"] == 0
    # a two-byte keyword cannot be covered at four alignments: stride 2; a lone '@' between classes becomes a trigger byte
    h2 = Harness(harness_lib, lits + [("ok", 0, 3), (r"[a-z]+@[a-z]+", 0, 3)])
    i2 = h2.info()
    assert i2["stride"] == 2 and i2["triggers"] == 1 and i2["n_always"] == 0
    rl = W.make_rules(300)
    rules = W.rules_as_tuples(rl)
    data, off, _ = W.make_messages(300, 256, rl, p_hit=0.3, seed=77)
    buf = data.numpy()
    msgs = [bytes(buf[int(off[i]):int(off[i + 1])]) for i in range(300)]
    hs = [Harness(harness_lib, rules), Harness(harness_lib, rules, bitmap_kb=16), Harness(harness_lib, rules, bitmap_kb=128), Harness(harness_lib, rules, max_keys=1 << 20)]
    assert hs[1].info()["bitmap_bytes"] == 16384 and hs[2].info()["bitmap_bytes"] == 131072
    ref = [(hs[0].candidates2(m)[:2], hs[0].policy_hits(m)) for m in msgs]
    assert sum(1 for r in ref if r[1]) >= 60
    for h in hs:
        for lead in (0, 1, 2, 3, 5, 16, 31):
            got = [(h.candidates2(m, lead=lead, seed=7 * lead + i)[:2], h.policy_hits(m, lead=lead, seed=i)) for i, m in enumerate(msgs)]
            assert got == ref, (h.info(), lead)
    # the bitmap only ever costs time: a denser one flags more grams, never fewer candidates
    fl16, occ16 = hs[1].batch_rates(buf, 300, 256)
    fl128, occ128 = hs[2].batch_rates(buf, 300, 256)
    assert occ16 == occ128 and fl16 >= fl128
    for h in hs + [h4, h2]:
        h.close()


def test_non_ascii_rule_packs_and_lazy_block(oracle, harness_lib):
    """the forms of the reference's other rule packs -- CJK / Cyrillic / Hangul literal alternations, classes
    mixing CJK ranges with \\w, `.*` between literals, the lazy PEM block -- through the product compiler + VM on the host."""
    h = Harness(harness_lib, PACK_RULES)
    assert (h.status == 0).all(), [h.L.harness_rule_error(h.h, i) for i in np.nonzero(h.status)[0]]
    msgs = [t.encode("utf-8") for t in PACK_TEXTS]
    total = 0
    for ri, r in enumerate(PACK_RULES):
        reg = oracle.Regex(r[0], "i" if r[1] else "")
        exp = oracle_spans(oracle, reg, msgs)
        for mi, m in enumerate(msgs):
            got = h.find_all(ri, m)
            assert [(g[2], g[3]) for g in got] == exp.get(mi, []), (r[0], PACK_TEXTS[mi])
            total += len(got)
            if exp.get(mi):
                assert ri in h.candidates(m) and ri in h.policy_hits(m)
            else:
                assert ri not in h.policy_hits(m)
    assert total >= 14
    h.close()


def test_abi_exports_every_declared_symbol():
    """every CG_API symbol of include/openclaw_gov.h is exported by the built library (no compute calls)."""
    from vainplex_openclaw_b200 import _native as N
    hdr = open(os.path.join(HERE, "..", "include", "openclaw_gov.h")).read()
    declared = set(re.findall(r"CG_API\s+[\w\s\*]+?\b(cg_\w+)\s*\(", hdr))
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    L = N.load()
    for name in declared:
        assert hasattr(L, name), name
    assert L.cg_version() >= 100


def test_no_device_fails_loudly():
    """Without a CUDA device the product refuses to compute (no CPU fallback)."""
    from vainplex_openclaw_b200 import _native as N
    if N.load().cg_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(N.GovError) as ei:
        N.init()
    assert ei.value.code == -3
    with pytest.raises(N.GovError):
        N.Ruleset([("abc", 0, 3)])


def test_bit_parallel_matcher_agrees_with_the_vm(harness_lib, oracle):
    """bitprog.h: for every confirmed factor occurrence of the tests above's kind of traffic AND of random regexes, the
    bit-parallel matcher (resolve_kernel's fast path) either declines (non-ASCII island) or answers what the Pike VM
    answers; most rules are eligible, and it does decide the bulk of the occurrences."""
    import ctypes as C
    from vainplex_openclaw_b200 import workload as W
    stats = (C.c_uint64 * 3)()
    harness_lib.harness_bitprog_stats(stats)
    base = list(stats)
    rl = W.make_rules(300)
    rules = W.rules_as_tuples(rl)
    h = Harness(harness_lib, rules)
    assert harness_lib.harness_bitprog_eligible(h.h) >= 0.8 * len(rules)
    d, o, _ = W.make_messages(3000, 200, rl, p_hit=0.3, seed=11, utf8_frac=0.1) if "utf8_frac" in W.make_messages.__code__.co_varnames else W.make_messages(3000, 200, rl, p_hit=0.3, seed=11)
    buf = d.numpy(); off = o.numpy()
    for i in range(3000):
        h.policy_hits(bytes(buf[int(off[i]):int(off[i + 1])]), lead=i % 19, seed=i)
    h.close()
    rng = np.random.default_rng(77)
    for k in range(300):
        src = random_regex(rng)
        hh = Harness(harness_lib, [(src, int(rng.integers(0, 2)), 0)])
        if int(hh.status[0]) == 0:
            for j in range(40):
                n = int(rng.integers(0, 60))
                m = bytes(rng.choice(np.frombuffer(b"abcxyz019 _-.@\n", dtype=np.uint8), n)) if n else b""
                hh.policy_hits(m, lead=j % 5, seed=j)
        hh.close()
    harness_lib.harness_bitprog_stats(stats)
    decided, mismatch, fallback = (int(stats[i]) - int(base[i]) for i in range(3))
    assert mismatch == 0, (decided, mismatch, fallback)
    assert decided > 800 and decided > fallback


def test_island_matcher_two_words_lookaround_and_factor_skip(harness_lib, oracle):
    """bitprog.h beyond one word: programs with 64 .. 127 consuming instructions (two state words), single-unit lookaround over
    ASCII sets (token-boundary guards: 32 contexts), and the resume behind a confirmed factor (factor_skip).  Every text is
    built around the places where those differ from the plain walk: the guard bytes on both sides, a match ending exactly at
    the end of the message, class runs one short of / at / beyond their bounds, a state that crosses the 64-bit word.  The
    harness decides each occurrence with island_test AND the Pike VM and counts disagreements; hits must equal the oracle's."""
    import ctypes as C
    rules = [
        (r"(?<![A-Z0-9])PZULOYQ[0-9A-Z]{16}(?![A-Z0-9])", 0, 0),
        (r"(?<!\d)\+?[1-9]\d{6,14}(?!\d)", 0, 0),
        (r"(?:password|passwd|pwd|secret|token|api_key|apikey)\s*[:=]\s*['\"]?[^\s'\"]{8,64}", 1, 0),
        (r"sk-ant-[a-zA-Z0-9-]{80,}", 0, 0),
        (r"hgisr[a-zA-Z0-9_-]{36,}", 0, 0),
        (r"(?<=[a-f])zq[0-9]{3}(?=[xy])", 0, 0),
        (r"\bkey_[a-z]{70,90}\b", 0, 0),
        (r"abc-[A-Za-z0-9]{16,40}", 0, 0),
    ]
    h = Harness(harness_lib, rules)
    assert all(int(x) == 0 for x in h.status[:len(rules)])
    assert harness_lib.harness_bitprog_eligible(h.h) == len(rules)
    regs = oracle_regexes(oracle, rules)
    stats = (C.c_uint64 * 3)()
    harness_lib.harness_bitprog_stats(stats)
    base = list(stats)
    tok = "PZULOYQ" + "A1B2C3D4E5F6G7H8"
    texts = [tok, " " + tok, tok + " ", "x" + tok + ".", "X" + tok, tok + "9", "-" + tok + "-", tok[:-1], tok + tok,
             "+4915112345678", "call +4915112345678 now", "1234567", "123456", "a1234567b", "91234567890123456", "x+1234567",
             "token=abcdefgh", "TOKEN : 'abcdefghij'", "secret=short", "password = \"" + "p" * 64 + "\"", "api_key=" + "k" * 70, "pwd:1234567", "pwd:12345678",
             "sk-ant-" + "a" * 79, "sk-ant-" + "a" * 80, "say sk-ant-" + "Ab-9" * 25 + " end", "sk-ant-" + "a" * 200,
             "hgisr" + "_" * 35, "hgisr" + "_" * 36, "hgisr" + "x" * 120,
             "azq123x", "gzq123x", "fzq123y!", "azq12x", "zq123x", "azq123",
             "key_" + "a" * 69, "key_" + "a" * 70, "key_" + "a" * 90, "key_" + "a" * 91, "a key_" + "b" * 80 + " z", "xkey_" + "b" * 80,
             "abc-" + "Z" * 15, "abc-" + "Z" * 16, "abc-" + "Z" * 41, "abc-abc-" + "Q" * 16, "\u00e9abc-" + "Z" * 16, "abc-" + "Z" * 10 + "\u00e9" + "Z" * 16]
    msgs = [t.encode("utf-8") for t in texts]
    data, off = oracle.pack(msgs)
    bits, _ = oracle.scan_policy(regs, data, off)
    for i, (t, m) in enumerate(zip(texts, msgs)):
        got = h.policy_hits(m, lead=len(m) % 7, seed=3)
        row = np.unpackbits(bits[i], bitorder="little")[:len(rules)]
        exp = {int(r) for r in np.nonzero(row)[0]}
        assert got == exp, (t, got, exp)
    harness_lib.harness_bitprog_stats(stats)
    decided, mismatch, fallback = (int(stats[i]) - int(base[i]) for i in range(3))
    assert mismatch == 0 and decided >= 40, (decided, mismatch, fallback)
    h.close()


def test_cortex_language_packs_compile_and_match(harness_lib, oracle):
    """SURVEY 8 f4 with the real packs (tests/golden/cortex_pack_vectors.json): all 70 regex literals of the ten language
    packs compile (flag i on Cyrillic / Latin-1 literals included), and the product's pipeline (gram filter -> factors ->
    matchers, host harness) answers the reference's 100 anyMatch assertions."""
    c = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cortex_pack_vectors.json"), encoding="utf-8"))
    rules, where = [], {}
    for lang, cats in c["packs"].items():
        for cat, pats in cats.items():
            for p in pats:
                where.setdefault((lang, cat), []).append(len(rules))
                rules.append((p["source"], 1 if "i" in p["flags"] else 0, 3))
    h = Harness(harness_lib, rules)
    assert [int(x) for x in h.status[:len(rules)]] == [0] * len(rules), [(rules[i][0], harness_lib.harness_rule_error(h.h, i)) for i in range(len(rules)) if h.status[i]]
    for v in c["vectors"]:
        m = v["text"].encode("utf-8")
        hits = h.policy_hits(m)
        assert any(r in hits for r in where[(v["lang"], v["category"])]) == v["expect"], (v["lang"], v["category"], v["text"])
        assert all(h.test(r, m) == (r in hits) for r in where[(v["lang"], v["category"])])
    # case folding beyond ASCII, both directions, classes and negated classes (ECMA-262 22.2.2.7.3 / 22.2.2.9)
    extra = [("привет", 1, "ПРИВЕТ мир", True), ("ÄRGER", 1, "so ein ärger", True), ("straße", 1, "STRASSE", False), ("[а-я]+ть", 1, "ДЕЛАТЬ", True),
             ("[^а-я]", 1, "Я", False), ("ς", 1, "Σ", True), ("ǆ", 1, "ǅ", True), ("\\u212a", 1, "k", False), ("ſ", 1, "s", False), ("é", 0, "É", False)]
    for src, fl, text, want in extra:
        hh = Harness(harness_lib, [(src, fl, 3)])
        assert int(hh.status[0]) == 0, src
        assert hh.test(0, text.encode("utf-8")) == want == bool(oracle_spans(oracle, oracle.Regex(src, "i" if fl else ""), [text.encode("utf-8")])), (src, text)
        hh.close()
    h.close()
