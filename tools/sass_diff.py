"""Compare the SASS of every device function of the working tree's build (pairec_b200/csrc/build/*.o) with a git
revision's: `python tools/sass_diff.py <rev>`.  Used to show that opt-in code (new template instantiations, new
kernels) leaves the kernels of a build that was verified on the GPU untouched.  Template arguments added since <rev>
are stripped with --strip (repeatable), e.g. --strip ELb0EE=EE."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
        "--expt-relaxed-constexpr", "-c"]


def sass(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[cur].append(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip())   # drop addresses and encodings
    return funcs


def main():
    rev = sys.argv[1]
    strips = [a.split("=", 1) for i, a in enumerate(sys.argv) if i > 1 and sys.argv[i - 1] == "--strip"]
    tmp = tempfile.mkdtemp(prefix="sassdiff_")
    subprocess.run(f"git -C {ROOT} archive {rev} pairec_b200/csrc include | tar x -C {tmp}", shell=True, check=True)
    src = os.path.join(tmp, "pairec_b200", "csrc")
    bad = 0
    for f in sorted(os.listdir(src)):
        if not f.endswith(".cu"):
            continue
        new_obj = os.path.join(ROOT, "pairec_b200", "csrc", "build", f[:-3] + ".o")
        if not os.path.exists(new_obj):
            print(f"{f}: not in the working tree's build")
            continue
        old_obj = os.path.join(tmp, f[:-3] + ".o")
        subprocess.run(NVCC + [f, "-o", old_obj], cwd=src, check=True, stderr=subprocess.DEVNULL)
        a, b = sass(old_obj), sass(new_obj)
        for k, v in strips:   # a new function also answers to its name with the added template argument removed
            b.update({name.replace(k, v): body for name, body in list(b.items()) if name.replace(k, v) not in b})
        diff = {}
        for name, body in a.items():
            if name not in b:
                diff[name] = "missing"
            elif body != b[name]:
                n = sum(x != y for x, y in zip(body, b[name])) + abs(len(body) - len(b[name]))
                diff[name] = f"{n} of {len(body)} instructions differ"
        bad += len(diff)
        print(f"{f}: {len(a)} device functions, " + ("all identical" if not diff else str(diff)))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
