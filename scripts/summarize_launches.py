"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into per-kernel shares of ONE step
(the launches between two consecutive adam_kernel launches).  Usage: summarize_launches.py list.csv [step_index]"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    n = re.sub(r"<unnamed>::", "", name)
    m = re.match(r"(?:void )?([\w:]+(?:<[^(]*>)?)\(", n)
    return m.group(1) if m else n[:60]


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append((short(r["Kernel Name"]), r["Grid Size"], float(r["Metric Value"].replace(",", "")) * 1e-6))
    ends = [i for i, r in enumerate(rows) if r[0].startswith("adam_kernel")]
    starts = [i for i, r in enumerate(rows) if r[0].startswith("fill_kernel")]     # canvas fill opens every step
    if not ends:
        print("no adam_kernel launch: cannot delimit a step"); return
    b = ends[which] + 1
    prev = [e for e in ends if e < ends[which]]
    # a step opens with the canvas fill that follows the previous step's Adam (fill_kernel also zeroes the CSD sums mid-step)
    a = min((i for i in starts if i > prev[-1]), default=prev[-1] + 1) if prev else min(starts)
    step = rows[a:b]
    agg = defaultdict(lambda: [0, 0.0])
    for k, g, ms in step:
        agg[k][0] += 1; agg[k][1] += ms
    tot = sum(v[1] for v in agg.values())
    print(f"one step: {len(step)} launches, {tot:.3f} ms serialised ({len(ends)} steps in the list)\n")
    print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {ms:.3f} | {100 * ms / tot:.1f}% |")
    if "--gemm" in sys.argv:
        print("\nGEMM launches by grid:")
        g = defaultdict(lambda: [0, 0.0])
        for k, gr, ms in step:
            if k.startswith("tc_gemm"):
                g[(k, gr)][0] += 1; g[(k, gr)][1] += ms
        for (k, gr), (c, ms) in sorted(g.items(), key=lambda kv: -kv[1][1])[:40]:
            print(f"  {k:42s} grid {gr:16s} x{c:3d}  {ms:8.3f} ms  ({ms / c * 1e3:7.1f} us each)")


if __name__ == "__main__":
    main()
