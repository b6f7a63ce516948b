"""Shared helpers for the test-suite (CPU harness wrapper, oracle adapters)."""
import ctypes as C

import numpy as np


class Harness:
    """Host-compiled product compiler + VM (tests/native/vm_harness.cpp)."""

    def __init__(self, L, rules, stride=0, bitmap_kb=0, max_keys=0):
        self.L = L
        self.srcs = [r[0] if isinstance(r[0], bytes) else r[0].encode("utf-8", "surrogatepass") for r in rules]
        arr = (C.c_char_p * max(1, len(rules)))(*self.srcs)
        lens = np.array([len(s) for s in self.srcs] or [0], dtype=np.uint32)
        flags = np.array([r[1] for r in rules] or [0], dtype=np.uint32)
        self.status = np.zeros(max(1, len(rules)), dtype=np.int32)
        self.h = L.harness_create(arr, lens.ctypes.data, flags.ctypes.data, len(rules), stride, bitmap_kb, max_keys,
                                  self.status.ctypes.data)
        assert self.h
        self.n = len(rules)
        self.rw = max(1, (self.n + 31) // 32)

    def info(self):
        o = np.zeros(16, dtype=np.uint32)
        self.L.harness_info(self.h, o.ctypes.data)
        return dict(keys=int(o[0]), n_factors=int(o[1]), stride=int(o[2]), factor_min=int(o[3]) & 0xff, factor_max=int(o[3]) >> 8,
                    n_always=int(o[4]), image_bytes=int(o[5]), prog_words=int(o[6]), entries=int(o[7]), shapes=int(o[8]),
                    triggers=int(o[9]), bitmap_bytes=int(o[10]), tables_resident=int(o[11]))

    def find_all(self, rule, msg: bytes):
        cap = 64
        while True:
            out = np.zeros(4 * cap, dtype=np.uint32)
            k = self.L.harness_find_all(self.h, rule, msg, len(msg), out.ctypes.data, cap)
            assert k >= 0, "VM overflow"
            if k <= cap:
                return [tuple(int(x) for x in out[4 * j:4 * j + 4]) for j in range(k)]
            cap = k

    def test(self, rule, msg: bytes) -> bool:
        r = self.L.harness_test(self.h, rule, msg, len(msg))
        assert r >= 0
        return bool(r)

    def candidates2(self, msg: bytes, want_spans=False, lead=16, seed=1):
        """-> (queued-for-VM rules, direct-hit rules, number of flagged grams).  The message sits `lead` bytes into a
        buffer of seeded filler (its neighbours in a batch); lead < 3 exercises the scan kernel's head check."""
        c = np.zeros(self.rw, dtype=np.uint32)
        d = np.zeros(self.rw, dtype=np.uint32)
        l1 = np.zeros(1, dtype=np.uint32)
        self.L.harness_candidates(self.h, msg, len(msg), c.ctypes.data, d.ctypes.data, 1 if want_spans else 0, l1.ctypes.data, lead, seed)
        f = lambda bits: {r for r in range(self.n) if (bits[r >> 5] >> (r & 31)) & 1}
        return f(c), f(d), int(l1[0])

    def policy_hits(self, msg: bytes, lead=16, seed=1):
        """rules with RegExp.test(msg) true, computed the way the device pipeline does (gram filter -> exact factor -> island VM)"""
        bits = np.zeros(self.rw, dtype=np.uint32)
        self.L.harness_policy_hits(self.h, msg, len(msg), bits.ctypes.data, lead, seed)
        return {r for r in range(self.n) if (bits[r >> 5] >> (r & 31)) & 1}

    def policy_hits_batch(self, data: np.ndarray, off: np.ndarray):
        """[(msg, rule)] sorted: the device pipeline restated over a packed batch (buffer-relative gram positions, head
        check, message lookup); data must extend 16 bytes past off[-1]"""
        n = len(off) - 1
        o = np.ascontiguousarray(off, dtype=np.uint32)
        bits = np.zeros(max(1, n) * self.rw, dtype=np.uint32)
        self.L.harness_policy_hits_batch(self.h, data.ctypes.data, o.ctypes.data, n, bits.ctypes.data)
        bits = bits.reshape(max(1, n), self.rw)
        return [(m, r) for m in range(n) for r in range(self.n) if (bits[m, r >> 5] >> (r & 31)) & 1]

    def batch_rates(self, data: np.ndarray, n_msgs: int, msg_len: int):
        """(flagged grams, confirmed factor occurrences) over n_msgs fixed-length messages laid out back to back"""
        o = np.zeros(2, dtype=np.uint64)
        self.L.harness_batch_rates(self.h, data.ctypes.data, n_msgs, msg_len, o.ctypes.data)
        return int(o[0]), int(o[1])

    def dump(self):
        self.L.harness_dump_factors(self.h)

    def candidates(self, msg: bytes, lead=16, seed=1):
        c, d, _ = self.candidates2(msg, lead=lead, seed=seed)
        return c | d
    def close(self):
        if self.h:
            self.L.harness_destroy(self.h)
            self.h = None


def oracle_regexes(O, rules):
    """rules: (source, flags, category) -> list of oracle Regex or None when the oracle rejects it."""
    out = []
    for src, fl, _cat in rules:
        try:
            out.append(O.Regex(src, "i" if fl & 1 else ""))
        except O.RegexSyntaxError:
            out.append(None)
    return out


def words_to_tuple(w):
    w = int(w)
    return (w >> 63, (w >> 32) & 0x7fffffff, w & 0xffffffff) if w else (0, 0, 0xffffffff)

# Rule packs in the FORMS of the reference's other packs: alternations of CJK / Cyrillic literals,
# classes mixing CJK ranges with \w, a greedy wildcard between literals, a capture group after optional blanks
# (cortex/src/patterns/lang-*.ts), and the lazy PEM block of cortex/src/trace-analyzer/redactor.ts.  The words are ours.
PACK_RULES = [
    (r"(?:已经决定|方案确定|我们选用|就这样办)", 0, 3),
    (r"(?:关于|回到|讨论一下)\s*([\u4e00-\u9fff\w]{2,20})", 0, 3),
    (r"(?:等待|被.*卡住|需要.*才可以)", 0, 3),
    (r"(?:решили|договорились|утвердили план)", 0, 3),
    (r"(?:насчёт|по поводу)\s+([а-яёА-ЯЁ\w]{3,24})", 0, 3),
    (r"(?:결정했|확정했|하기로 했)", 0, 3),
    (r"(?:すごい|完璧|やった)[!！]*", 0, 3),
    (r"-----BEGIN [A-Z ]+-----[\s\S]*?-----END [A-Z ]+-----", 0, 0),
    (r"(?:schluessel|geheimnis)\s*[:=]\s*\S{6,}", 1, 0),
]
PACK_TEXTS = [
    "今天开会，我们选用新的部署流程，关于 数据库迁移_v2 的事情稍后再说。",
    "被上游的评审卡住了，需要安全团队签字才可以继续；讨论一下预算",
    "Вчера договорились и утвердили план, а насчёт  бюджета2025 решим позже",
    "по поводу релиза: решили выкатывать в пятницу",
    "우리는 금요일에 배포하기로 했습니다. 확정했어요",
    "すごい！！ 完璧 やった!",
    "key follows\n-----BEGIN RSA PRIVATE KEY-----\nMIIB\n😀\n-----END RSA PRIVATE KEY-----\ntrailer -----END X-----",
    "Schluessel = abc123def and GEHEIMNIS:xyz (short) ünd GeheimNis=лмнопрст",
    "nothing here 😀 plain ascii and ünïcödé",
    "",
]


# random ECMAScript patterns over a tiny alphabet (dense matches, empty matches, astral characters, lazy quantifiers, look-arounds)
ATOMS = ["a", "b", "c", "1", " ", "é", "😀", ".", r"\d", r"\w", r"\s", r"\S", r"\W", "[ab]", "[^a]", "[a-c1]", r"[^\s1]", r"\b", r"\B",
         "^", "$", "(?!a)", "(?=b)", "(?<!a)", "(?<=b)", "(?<![a1])", r"(?!\d)"]
QUANTS = ["", "", "", "*", "+", "?", "{2}", "{1,2}", "{2,}", "*?", "+?", "??", "{1,3}?"]


def random_regex(rng, depth=0):
    n = int(rng.integers(1, 5))
    parts = []
    for _ in range(n):
        u = rng.random()
        if depth < 2 and u < 0.25:
            inner = random_regex(rng, depth + 1)
            if rng.random() < 0.5:
                inner = inner + "|" + random_regex(rng, depth + 1)
            atom = ("(?:%s)" if rng.random() < 0.6 else "(%s)") % inner
        else:
            atom = ATOMS[int(rng.integers(0, len(ATOMS)))]
        q = QUANTS[int(rng.integers(0, len(QUANTS)))]
        if atom in ("^", "$", r"\b", r"\B") or atom.startswith("(?<") :
            q = ""
        if atom.startswith("(?=") or atom.startswith("(?!"):
            q = ""
        parts.append(atom + q)
    return "".join(parts)


