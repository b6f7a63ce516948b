#!/usr/bin/env python3
"""Extract the scan golden vectors the reference's own tests hold for the hot path.

Source : /root/reference/packages/openclaw-governance/test/redaction/registry.test.ts
Output : tests/golden/registry_vectors.json  (committed; /root/reference is not on the GPU box)

Every `it(...)` block that calls `reg.findMatches(<string expr>)` becomes one vector:
  {line, name, categories, custom, input, checks:[...]}
with the block's `expect(...)` assertions translated to data:
  {"op":"some","ids":[...],"value":bool}            matches.some(m => m.pattern.id === ...)
  {"op":"len_eq"|"len_ge","n":N[,"filter_id":id]}     matches.length / filtered length
  {"op":"id_at","i":I,"id":ID}, {"op":"match_at","i":I,"match":S}
  {"op":"find_defined","id":ID}
Only the test *data* (inputs / expected values) is extracted -- no reference code is copied.
Run from the repo root inside the build container:  python tests/golden/extract_registry_vectors.py
"""
import json
import os
import re
import sys

SRC = "/root/reference/packages/openclaw-governance/test/redaction/registry.test.ts"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "registry_vectors.json")


def js_str_literal(tok: str, env: dict) -> str:
    q = tok[0]
    body = tok[1:-1]
    if q == "`":
        body = re.sub(r"\$\{(\w+)\}", lambda m: env[m.group(1)], body)
    out, i = [], 0
    while i < len(body):
        c = body[i]
        if c == "\\":
            n = body[i + 1]
            i += 2
            if n == "n":
                out.append("\n")
            elif n == "t":
                out.append("\t")
            elif n == "r":
                out.append("\r")
            elif n == "u":
                out.append(chr(int(body[i:i + 4], 16)))
                i += 4
            elif n == "x":
                out.append(chr(int(body[i:i + 2], 16)))
                i += 2
            else:
                out.append(n)
        else:
            out.append(c)
            i += 1
    return "".join(out)


TOKEN = re.compile(r"""\s*(?:("(?:[^"\\]|\\.)*"|'(?:[^'\\]|\\.)*'|`(?:[^`\\]|\\.)*`)|(\.repeat\((\d+)\))|(\+)|(\w+))""")


def eval_str_expr(expr: str, env: dict) -> str:
    pos, parts = 0, []
    expr = expr.strip()
    while pos < len(expr):
        m = TOKEN.match(expr, pos)
        if not m:
            raise ValueError("cannot parse string expr: %r at %d" % (expr, pos))
        pos = m.end()
        if m.group(1):
            parts.append(js_str_literal(m.group(1), env))
        elif m.group(2):
            parts[-1] = parts[-1] * int(m.group(3))
        elif m.group(4):
            pass
        elif m.group(5):
            parts.append(env[m.group(5)])
    return "".join(parts)


def parse_make_registry(argtext: str):
    argtext = argtext.strip()
    cats = ["credential", "pii", "financial"]
    custom = []
    if not argtext:
        return cats, custom
    m = re.match(r"\[([^\]]*)\]", argtext)
    cats = re.findall(r'"(\w+)"', m.group(1))
    rest = argtext[m.end():]
    for cm in re.finditer(r'\{\s*name:\s*("(?:[^"\\]|\\.)*"),\s*regex:\s*("(?:[^"\\]|\\.)*"),\s*category:\s*("(?:[^"\\]|\\.)*")\s*\}', rest):
        custom.append({"name": js_str_literal(cm.group(1), {}), "regex": js_str_literal(cm.group(2), {}),
                       "category": js_str_literal(cm.group(3), {})})
    return cats, custom


def balanced_call(text: str, start: int) -> int:
    """index just past the ')' matching the '(' at text[start]; skips string literals."""
    depth, i = 0, start
    while i < len(text):
        c = text[i]
        if c in "\"'`":
            j = i + 1
            while text[j] != c:
                j += 2 if text[j] == "\\" else 1
            i = j + 1
            continue
        if c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced")


def main():
    src = open(SRC, encoding="utf-8").read()
    lines = src.split("\n")
    vectors = []
    describe_reg = None          # (cats, custom) of the enclosing describe's `const reg`
    i = 0
    while i < len(lines):
        ln = lines[i]
        md = re.match(r"^  describe\(", ln)
        if md:
            describe_reg = None
        mr = re.match(r"^    const reg = makeRegistry\((.*)\);\s*$", ln)
        if mr:
            describe_reg = parse_make_registry(mr.group(1))
        mi = re.match(r'^    it\((".*?"|\'.*?\'),', ln)
        if not mi:
            i += 1
            continue
        name = js_str_literal(mi.group(1), {})
        j = i + 1
        while not re.match(r"^    \}\);\s*$", lines[j]):
            j += 1
        block = "\n".join(lines[i + 1:j])
        block_nc = re.sub(r"^\s*//.*$", "", block, flags=re.M)
        env, regs, inputs, filters, finds = {}, {"reg": describe_reg}, {}, {}, {}
        checks_by_var = {}
        # local registries (possibly multi-line)
        for m in re.finditer(r"const (\w+) = makeRegistry\(", block_nc):
            end = balanced_call(block_nc, m.end() - 1)
            regs[m.group(1)] = parse_make_registry(block_nc[m.end():end - 1])
        for stmt in re.finditer(r"const (\w+) = ([^;]*?);", block_nc, flags=re.S):
            var, rhs = stmt.group(1), stmt.group(2).strip()
            if rhs.startswith("makeRegistry("):
                continue
            fm = re.match(r"(\w+)\.findMatches\((.*)\)$", rhs, flags=re.S)
            if fm:
                inputs[var] = (fm.group(1), eval_str_expr(fm.group(2), env), stmt.start())
                continue
            fl = re.match(r'(\w+)\.filter\(\(m\) => m\.pattern\.id === "([\w-]+)"\)$', rhs)
            if fl:
                filters[var] = (fl.group(1), fl.group(2))
                continue
            fd = re.match(r'(\w+)\.find\(\(m\) => m\.pattern\.id === "([\w-]+)"\)$', rhs)
            if fd:
                finds[var] = (fd.group(1), fd.group(2))
                continue
            try:
                env[var] = eval_str_expr(rhs, env)
            except Exception:
                pass
        inline_n = 0
        for ex in re.finditer(r"expect\(", block_nc):
            end = balanced_call(block_nc, ex.end() - 1)
            inner = block_nc[ex.end():end - 1].strip()
            tail = re.match(r"\.(\w+)\((.*?)\);", block_nc[end:], flags=re.S)
            if not tail:
                continue
            matcher, arg = tail.group(1), tail.group(2).strip()
            chk, var = None, None
            m1 = re.match(r"(\w+)\.some\(\(m\) => (.*)\)$", inner, flags=re.S)
            m2 = re.match(r"(\w+)\.length$", inner)
            m3 = re.match(r"(\w+)\[(\d+)\]!\.pattern\.id$", inner)
            m4 = re.match(r"(\w+)\[(\d+)\]!\.match$", inner)
            m5 = re.match(r"(\w+)\.findMatches\((.*)\)\.length$", inner, flags=re.S)
            if m1 and matcher == "toBe":
                var = m1.group(1)
                chk = {"op": "some", "ids": re.findall(r'm\.pattern\.id === "([\w-]+)"', m1.group(2)), "value": arg == "true"}
            elif m2 and matcher in ("toBe", "toBeGreaterThanOrEqual"):
                var = m2.group(1)
                chk = {"op": "len_eq" if matcher == "toBe" else "len_ge", "n": int(arg)}
                if var in filters:
                    chk["filter_id"] = filters[var][1]
                    var = filters[var][0]
            elif m3 and matcher == "toBe":
                var = m3.group(1)
                chk = {"op": "id_at", "i": int(m3.group(2)), "id": js_str_literal(arg, {})}
            elif m4 and matcher == "toBe":
                var = m4.group(1)
                chk = {"op": "match_at", "i": int(m4.group(2)), "match": js_str_literal(arg, {})}
            elif m5 and matcher == "toBe":
                var = "__inline%d" % inline_n
                inline_n += 1
                inputs[var] = (m5.group(1), eval_str_expr(m5.group(2), env), ex.start())
                chk = {"op": "len_eq", "n": int(arg)}
            elif inner in finds and matcher == "toBeDefined":
                var = finds[inner][0]
                chk = {"op": "find_defined", "id": finds[inner][1]}
            if chk is not None and var in inputs:
                checks_by_var.setdefault(var, []).append(chk)
        for var, (regname, text, pos) in inputs.items():
            cats_custom = regs.get(regname)
            if cats_custom is None or var not in checks_by_var:
                continue
            line = i + 2 + block_nc[:pos].count("\n")
            vectors.append({"line": line, "name": name, "categories": cats_custom[0], "custom": cats_custom[1],
                            "input": text, "checks": checks_by_var[var]})
        i = j + 1
    vectors.sort(key=lambda v: v["line"])
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump({"source": "packages/openclaw-governance/test/redaction/registry.test.ts",
                   "reference_commit": "105d2b9", "vectors": vectors}, f, indent=1, ensure_ascii=True)
        f.write("\n")
    print("wrote %d vectors (%d checks) to %s" % (len(vectors), sum(len(v["checks"]) for v in vectors), OUT))


if __name__ == "__main__":
    sys.exit(main())
