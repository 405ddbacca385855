"""Summarise an .ncu-rep (raw + source pages) into the few numbers the design discussion needs."""
import csv, subprocess, sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second"]
for r in rows[2:]:
    print("=" * 100)
    for k in KEYS:
        if k in ix:
            print(f"  {k:85s} {r[ix[k]][:80]}")
    st = {h.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue")[0]: float(r[ix[h]] or 0)
          for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio")}
    print("  stalls/issue:", ", ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 3]
    ix = {h: i for i, h in enumerate(hdr)}
    g = lambda r, k: (float(r[ix[k]] or 0) if r[ix[k]].replace(".", "").isdigit() else 0.0) if ix[k] < len(r) else 0.0
    tot = sum(g(r, "# Samples") for r in data)
    c = Counter()
    for r in data:
        s = r[ix["Source"]].split()
        if s:
            c[s[1] if s[0].startswith("@") else s[0]] += g(r, "# Samples")
    print("samples", tot, [(k, int(v)) for k, v in c.most_common(14)])
    for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:int(sys.argv[3]) if len(sys.argv) > 3 else 12]:
        st = sorted(((k, g(r, k)) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and g(r, k) > 0), key=lambda x: -x[1])[:2]
        print(f"  {r[ix['Source']][:72]:72s} {int(g(r, '# Samples')):6d} {st}")
