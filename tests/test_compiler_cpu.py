"""CPU tier for the product's host logic: the rule compiler, the prefilter tables and the Pike-VM
*source* (compiled for the host by tests/native/vm_harness.cpp) against the oracle.  No GPU needed;
the GPU tier (test_gpu_parity.py) repeats the comparisons through the C ABI and the real kernels."""
import json
import os
import re

import numpy as np
import pytest

from helpers import Harness, PACK_RULES, PACK_TEXTS, random_regex
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


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("utf8_frac", [0.0, 0.15])
def test_synthetic_rules_spans_and_prefilter_soundness(oracle, harness_lib, mode, utf8_frac):
    rl = W.make_rules(160)
    rules = W.rules_as_tuples(rl)
    h = Harness(harness_lib, rules, mode=mode)
    assert (h.status == 0).all(), [h.L.harness_rule_error(h.h, i) for i in np.nonzero(h.status)[0]]
    data, off, inj = W.make_messages(1200, 200, rl, p_hit=0.25, utf8_frac=utf8_frac, seed=1234 + mode)
    buf = data.numpy()
    msgs = [bytes(buf[int(off[i]):int(off[i + 1])]) for i in range(1200)]
    n_hits = 0
    c2 = [h.candidates2(m) for m in msgs]
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
        assert h.policy_hits(m) == hits_by_msg.get(mi, set()), (mi, m)
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
    assert N.rule_check("É", N.FLAG_ICASE) == N.CG_ERR_UNSUPPORTED


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


def test_trap_table_image_and_profile_guided_ranking(harness_lib):
    """The shared-memory image of scan_kernel: plain state indices, accepting / non-resident transitions -> the absorbing
    trap row; and profile-guided residency (rank_states_by_visits) is a pure renumbering: same candidates, same hits,
    fewer visits to non-resident rows."""
    rl = W.make_rules(500)
    rules = W.rules_as_tuples(rl)
    h = Harness(harness_lib, rules, mode=2, budget_kb=48)          # small budget: most states are not resident
    data, off, _ = W.make_messages(400, 256, rl, p_hit=0.2, seed=77)
    buf = data.numpy()
    msgs = [bytes(buf[int(off[i]):int(off[i + 1])]) for i in range(400)]

    def check_image():
        img, hot, stride, lut_off, tab = h.image()
        ns, nc = tab.shape
        assert stride == 2 * nc + 4 and hot < ns and lut_off % 16 == 0 and len(img) == lut_off + 256
        rows = np.stack([np.frombuffer(img[r * stride:r * stride + 2 * nc].tobytes(), dtype=np.uint16) for r in range(hot + 1)])
        assert (rows[hot] == hot).all()                              # absorbing trap row
        nxt, acc = tab[:hot] & 0x3fff, (tab[:hot] & 0x8000) != 0
        trapped = acc | (nxt >= hot)
        assert (rows[:hot][trapped] == hot).all() and (rows[:hot][~trapped] == nxt[~trapped]).all()
        return hot

    hot = check_image()
    before = [(h.candidates2(m), h.policy_hits(m)) for m in msgs]
    hist, _ = h.l1_hist(buf, 400, 256)
    cold_before = int(hist[hot:].sum())
    h.rank(np.minimum(hist, 0xffffffff))
    assert check_image() == hot
    after = [(h.candidates2(m), h.policy_hits(m)) for m in msgs]
    assert before == after
    hist2, _ = h.l1_hist(buf, 400, 256)
    assert int(hist2.sum()) == int(hist.sum()) and hist2[0] == hist[0]      # the start state stays state 0
    cold_after = int(hist2[hot:].sum())
    assert cold_after <= cold_before and (cold_before == 0 or cold_after < cold_before)
    assert (np.diff(hist2[1:].astype(np.int64)) <= 0).all()                  # most visited first
    h.close()


def test_non_ascii_rule_packs_and_lazy_block(oracle, harness_lib):
    """the forms of the reference's other rule packs -- CJK / Cyrillic / Hangul literal alternations, classes
    mixing CJK ranges with \\w, `.*` between literals, the lazy PEM block -- through the product compiler + VM on the host."""
    h = Harness(harness_lib, PACK_RULES, mode=2)
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


def test_napi_shim_source_compiles_against_a_stub_header():
    """Node.js is absent, so the N-API addon cannot be built here; at least its C source must parse and type-check
    against the C ABI header (a stand-in node_api.h declares the handful of N-API calls it uses)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-I", os.path.join(root, "tests", "native"),
                        os.path.join(root, "napi", "openclaw_gov_napi.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
