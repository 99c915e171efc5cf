"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares for
the LAST train step (kernels after the last `adamw`-closed step boundary are grouped by demangled name)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((int(r["ID"]), r["Kernel Name"], ns))
    # step boundaries: interleave_gather_kernel opens the LLM part of each step; take from the last
    # im2col (tower start) to the end
    starts = [i for i, (_, n, _) in enumerate(rows) if "im2col_patch14" in n]
    begin = starts[-1] if starts else 0
    step = rows[begin:]
    agg = defaultdict(lambda: [0, 0.0])
    for _, n, ns in step:
        short = re.sub(r"\(.*", "", n)
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", short)
        agg[short][0] += 1
        agg[short][1] += ns
    total = sum(v[1] for v in agg.values())
    lines = [f"# launch list summary of the last train step in {path}",
             f"# kernels: {len(step)}  total device time (serialised, cold-cache): {total/1e6:.1f} ms", ""]
    lines.append(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k[:70]:70s} {c:8d} {ns/1e6:10.2f} {100*ns/total:6.1f}%")
    txt = "\n".join(lines)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
