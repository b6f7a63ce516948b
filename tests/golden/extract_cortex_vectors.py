#!/usr/bin/env python3
"""Extract the language-pack golden vectors the reference's own tests hold ("other rule packs on the same kernels").

Sources : /root/reference/packages/openclaw-cortex/src/patterns/lang-*.ts        (the packs: regex literals per category)
          /root/reference/packages/openclaw-cortex/test/patterns-lang-*.test.ts  (anyMatch(reg.getPatterns().<category>, "<text>") -> true / false)
Output  : tests/golden/cortex_pack_vectors.json (committed; /root/reference is not on the GPU box)

Only data is extracted: each pack's pattern sources + flags per category, and every test assertion as
{lang, category, text, expect}.  Run from the repo root inside the build container."""
import glob
import json
import os
import re

ROOT = "/root/reference/packages/openclaw-cortex"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cortex_pack_vectors.json")


def regex_literals(block: str):
    """regex literals `/.../flags` of a TS array body (handles escaped slashes and slashes inside classes)"""
    out, i = [], 0
    while i < len(block):
        if block[i] == "/" and (i == 0 or block[i - 1] in " \t\n,[("):
            j, in_class = i + 1, False
            while j < len(block):
                c = block[j]
                if c == "\\":
                    j += 2
                    continue
                if c == "[":
                    in_class = True
                elif c == "]":
                    in_class = False
                elif c == "/" and not in_class:
                    break
                elif c == "\n":
                    break
                j += 1
            if j < len(block) and block[j] == "/" and j > i + 1:
                k = j + 1
                while k < len(block) and block[k].isalpha():
                    k += 1
                out.append((block[i + 1:j], block[j + 1:k]))
                i = k
                continue
        i += 1
    return out


def js_string(tok: str) -> str:
    body = tok[1:-1]
    return re.sub(r"\\(u[0-9a-fA-F]{4}|.)", lambda m: {"n": "\n", "t": "\t"}.get(m.group(1), chr(int(m.group(1)[1:], 16)) if m.group(1)[0] == "u" and len(m.group(1)) == 5 else m.group(1)), body)


def main():
    packs, vectors = {}, []
    for path in sorted(glob.glob(ROOT + "/src/patterns/lang-*.ts")):
        lang = os.path.basename(path)[5:-3]
        src = open(path, encoding="utf-8").read()
        m = re.search(r"patterns:\s*\{(.*?)\n  \},", src, re.S)
        cats = {}
        for cm in re.finditer(r"(\w+):\s*\[(.*?)\n    \],", m.group(1), re.S):
            cats[cm.group(1)] = [{"source": s, "flags": f} for s, f in regex_literals(cm.group(2))]
        packs[lang] = cats
    for path in sorted(glob.glob(ROOT + "/test/patterns-lang-*.test.ts")):
        lang = os.path.basename(path)[len("patterns-lang-"):-len(".test.ts")]
        src = open(path, encoding="utf-8").read()
        for m in re.finditer(r"anyMatch\(\s*reg\.getPatterns\(\)\.(\w+),\s*(\"(?:[^\"\\]|\\.)*\"|'(?:[^'\\]|\\.)*')\s*\)\s*\)\s*\.toBe\((true|false)\)", src):
            vectors.append({"lang": lang, "category": m.group(1), "text": js_string(m.group(2)), "expect": m.group(3) == "true",
                            "line": src[:m.start()].count("\n") + 1})
    json.dump({"_source": "openclaw-cortex src/patterns/lang-*.ts + test/patterns-lang-*.test.ts (data only)", "packs": packs, "vectors": vectors},
              open(OUT, "w", encoding="utf-8"), ensure_ascii=False, indent=1)
    print("packs:", {k: sum(len(v) for v in c.values()) for k, c in packs.items()}, "vectors:", len(vectors))


if __name__ == "__main__":
    main()
