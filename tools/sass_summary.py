"""Per-kernel SASS opcode summary of libpairec_gpu.so: the mnemonics that prove the Blackwell paths
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UBLKCP = TMA, SYNCS = mbarrier,
UTCBAR = tcgen05.commit).  usage: python tools/sass_summary.py [lib] > profiles/rNN_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "pairec_b200/libpairec_gpu.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMXQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP",
         "SYNCS", "HMMA", "DFMA", "DMUL", "DADD", "REDUX", "LDGSTS"]
cur, per = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m:
        op, mod = m.group(1), m.group(2) or ""
        per[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                per[cur][w] += 1
                if w == "UTCHMMA" and ".2CTA" in mod:
                    per[cur]["UTCHMMA.2CTA"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
print(f"# SASS opcode summary of {lib} (cuobjdump -sass; {len(per)} kernels)")
for name, c in zip(demangle, per.values()):
    keys = [k for k in c if k != "_total"]
    for k in keys:
        tot[k] += c[k]
    if not keys:
        continue
    short = re.sub(r"\(.*", "", name).replace("prg::", "")[:90]
    print(f"{short:92s} instr={c['_total']:6d}  " + "  ".join(f"{k}={c[k]}" for k in sorted(keys)))
print("# totals: " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items())))
