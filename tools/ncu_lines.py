"""Per-source-line stall samples of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py REPORT KERNEL_REGEX [TOP_N]"""
import csv, subprocess, sys
from collections import defaultdict

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
done_first = False
cur_file, hdr, ix = None, None, None
acc = defaultdict(lambda: [0.0, 0.0, ""])   # (file, line) -> [samples, instructions executed, text]
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        if done_first and hdr is not None and any(acc):
            pass
        continue
    if r[0] == "Line No":
        if hdr is not None and cur_file is None:
            break
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] != "":        # a source line: its totals
        try:
            acc[(cur_file, int(r[0]))][0] += float(r[ix["# Samples"]] or 0)
            acc[(cur_file, int(r[0]))][1] += float(r[ix["Instructions Executed"]] or 0)
            acc[(cur_file, int(r[0]))][2] = r[1].strip()[:110]
        except ValueError:
            pass
tot = sum(v[0] for v in acc.values())
print("total samples", tot)
for (f, ln), v in sorted(acc.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{v[0]:8.0f} {100 * v[0] / max(tot, 1):5.1f}%  inst {v[1]:10.0f}  {f}:{ln}  {v[2]}")
