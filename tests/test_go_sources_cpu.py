"""No Go toolchain exists in the build container or on the GPU box, so the Go sources this repo ships — the replay
harness under baseline/go and the cgo stub in INTEGRATION.md — were written blind.  This is the little that can be checked
without a compiler: brackets balance outside strings / runes / comments, and every imported package is used (an unused
import is a compile error in Go)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip(src):
    """Go source with comments, string / raw-string / rune literals blanked out."""
    out, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
        elif src.startswith("/*", i):
            j = src.find("*/", i + 2)
            assert j >= 0, "unterminated block comment"
            out.append("\n" * src.count("\n", i, j))
            i = j + 2
        elif c == '"':
            j = i + 1
            while src[j] != '"':
                assert src[j] != "\n", "newline in interpreted string literal"
                j += 2 if src[j] == "\\" else 1
            out.append('""')
            i = j + 1
        elif c == "`":
            j = src.index("`", i + 1)
            out.append('""' + "\n" * src.count("\n", i, j))
            i = j + 1
        elif c == "'":
            j = i + 1
            while src[j] != "'":
                j += 2 if src[j] == "\\" else 1
            out.append("' '")
            i = j + 1
        else:
            out.append(c)
            i += 1
    return "".join(out)


def _check(name, src):
    code = _strip(src)
    stack = []
    pairs = {")": "(", "]": "[", "}": "{"}
    line = 1
    for ch in code:
        if ch == "\n":
            line += 1
        elif ch in "([{":
            stack.append((ch, line))
        elif ch in ")]}":
            assert stack and stack[-1][0] == pairs[ch], f"{name}:{line}: unbalanced {ch!r}"
            stack.pop()
    assert not stack, f"{name}: unclosed {stack[-1][0]!r} opened at line {stack[-1][1]}"
    # imports: `import "x/y"`, `import alias "x/y"` and the parenthesised block
    body_start = 0
    imports = []
    for m in re.finditer(r'^import\s*\(\s*\n(.*?)^\)', src, flags=re.S | re.M):
        body_start = max(body_start, m.end())
        for ln in m.group(1).splitlines():
            ln = ln.split("//")[0].strip()
            if ln:
                mm = re.match(r'(?:([A-Za-z_.][A-Za-z0-9_]*)\s+)?"([^"]+)"$', ln)
                assert mm, f"{name}: odd import line {ln!r}"
                imports.append((mm.group(1), mm.group(2)))
    for m in re.finditer(r'^import\s+(?:([A-Za-z_.][A-Za-z0-9_]*)\s+)?"([^"]+)"\s*$', src, flags=re.M):
        body_start = max(body_start, m.end())
        imports.append((m.group(1), m.group(2)))
    body = _strip(src[body_start:])
    for alias, path in imports:
        if path == "C" or alias in ("_", "."):
            continue
        pkg = alias or path.rstrip("/").split("/")[-1]
        if re.fullmatch(r"v\d+", pkg):                       # module major-version suffix
            pkg = path.rstrip("/").split("/")[-2]
        assert re.search(r"\b" + re.escape(pkg) + r"\.", body), f"{name}: import {path!r} ({pkg}) is never used"


def test_go_replay_harness_sources_are_well_formed():
    n = 0
    for d, _, files in os.walk(os.path.join(ROOT, "baseline", "go")):
        for f in files:
            if f.endswith(".go"):
                _check(f, open(os.path.join(d, f)).read())
                n += 1
    assert n >= 5


def test_cgo_stub_in_integration_md_is_well_formed():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```go\n(.*?)```", md, flags=re.S)
    assert blocks
    for i, b in enumerate(blocks):
        if b.lstrip().startswith("package "):
            _check(f"INTEGRATION.md go block {i}", b)          # a whole file: brackets and imports
        else:
            code = _strip(b)                                      # a fragment: brackets only
            for o, c in ("()", "[]", "{}"):
                assert code.count(o) == code.count(c), f"INTEGRATION.md go block {i}: unbalanced {o}{c}"
