"""Re-derives tests/golden/ct_merkle_vectors.json from first principles (hashlib + the RFC 6962 text) and checks that
the committed file says the same.  The leaves, roots and the inclusion / consistency cases are the Certificate
Transparency reference test vectors (RFC 6962 trees over 8 fixed leaves); nothing here reads /root/reference -- the
governance plugin has no Merkle code, so this published known-answer set is what pins the oracle's
tree convention.  Run: python tests/golden/make_ct_vectors.py  (exit 0 = file and derivation agree)."""
import hashlib
import json
import os
import sys


def H(b):
    return hashlib.sha256(b).digest()


def mth(ls):
    if len(ls) == 0:
        return H(b"")
    if len(ls) == 1:
        return H(b"\x00" + ls[0])
    k = 1
    while k * 2 < len(ls):
        k *= 2
    return H(b"\x01" + mth(ls[:k]) + mth(ls[k:]))


def path(m, ls):
    if len(ls) == 1:
        return []
    k = 1
    while k * 2 < len(ls):
        k *= 2
    return path(m, ls[:k]) + [mth(ls[k:])] if m < k else path(m - k, ls[k:]) + [mth(ls[:k])]


def proof(m, ls, b=True):
    if m == len(ls):
        return [] if b else [mth(ls)]
    k = 1
    while k * 2 < len(ls):
        k *= 2
    return proof(m, ls[:k], b) + [mth(ls[k:])] if m <= k else proof(m - k, ls[k:], False) + [mth(ls[:k])]


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    v = json.load(open(os.path.join(here, "ct_merkle_vectors.json")))
    leaves = [bytes.fromhex(x) for x in v["leaves_hex"]]
    bad = 0
    for i, r in enumerate(v["roots_hex"]):
        bad += mth(leaves[:i + 1]).hex() != r
    for c in v["inclusion"]:
        bad += [x.hex() for x in path(c["index"], leaves[:c["size"]])] != c["path_hex"]
    for c in v["consistency"]:
        bad += [x.hex() for x in proof(c["first"], leaves[:c["second"]])] != c["path_hex"]
    print("mismatches:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
