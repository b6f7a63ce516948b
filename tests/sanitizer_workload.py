"""Mixed workload for compute-sanitizer runs (not collected by pytest): ragged and long messages through the policy,
span, redaction and Merkle-log entry points.  Usage on the GPU box:
  compute-sanitizer --tool memcheck  --error-exitcode 3 python tests/sanitizer_workload.py
  compute-sanitizer --tool racecheck --error-exitcode 3 python tests/sanitizer_workload.py"""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vainplex_openclaw_b200 import _native as N, workload as W
N.init(0)
rl = W.make_rules(300); rules = W.rules_as_tuples(rl)
rs = N.Ruleset(rules, strict=True)
# ragged short messages + a few long ones (segmented path), policy + spans + redact + verdict-free paths
data_t, off_t, _ = W.make_messages(3000, 200, rl, p_hit=0.1, utf8_frac=0.1, seed=3)
buf, off0 = data_t.numpy(), off_t.numpy()
rng = np.random.default_rng(1)
msgs = [bytes(buf[int(off0[i]):int(off0[i]) + int(rng.integers(0, 201))]) for i in range(3000)]
data, off = N.pack(msgs)
w1, h1 = rs.scan_batch(data, off)
sp = rs.find_matches_batch(data, off)
out, oo, s2, dg = rs.redact_batch(data, off)
longm = [bytes(buf[:5000]), bytes(buf[5000:5000 + 70000]), b"", bytes(buf[100:3300])] + msgs[:50]
d2, o2 = N.pack(longm)
w2, h2 = rs.scan_batch(d2, o2)
w3, _ = rs.scan_batch(d2, o2)
assert np.array_equal(w2, w3)
lg = N.MerkleLog(); lg.append([bytes(rng.integers(0, 256, int(rng.integers(0, 200)), dtype=np.uint8)) for _ in range(300)]); r = lg.root(); p = lg.proof(17)
print("ok", int((w1 != 0).sum()), len(sp), len(s2), int((w2 != 0).sum()), len(p))
