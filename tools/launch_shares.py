"""Per-kernel share of one hot-path pass from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file X python bench.py ...`).

    python tools/launch_shares.py gpurun_out/launches.csv > profiles/rNN_launch_shares.md

The passes at the head of the capture are averaged (the tail of the list is the roofline probe's back-to-back decode tails)."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, per_pass=30):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = val / 1e3 if unit in ("ns", "nsecond") else val * (1e3 if unit in ("ms", "msecond") else 1.0)
        name = re.sub(r"^void |\(.*$", "", r["Kernel Name"]).replace("ldiff::", "")
        mine = "ldiff::" in r["Kernel Name"] or r["Kernel Name"].startswith("tc::")     # head_tc.cu's namespace
        rows.append((name, us, mine))
    ours = [(n, t) for n, t, mine in rows if mine]
    # passes in the capture = launches of a once-per-pass kernel; the launch order inside a pass is the
    # graph's, not the code's, so kernels are counted over the region of whole passes (everything before
    # the roofline probe, i.e. up to the last launch that is not a decode tail) instead of cut by position
    P = sum(n.startswith("lift_argmax") for n, _ in ours)        # (once per pass: the envelope or the per-pixel kernel)
    if P == 0:
        sys.exit("no pass found in the launch list")
    last = max(i for i, (n, _) in enumerate(ours) if not n.startswith("decode_tail"))
    agg = OrderedDict()
    for i, (n, t) in enumerate(ours):
        c = agg.setdefault(n, [0, 0, 0.0])           # launches in the pass region, launches, total us
        c[0] += i <= last
        c[1] += 1
        c[2] += t
    table = {n: (round(c[0] / P), c[2] / c[1]) for n, c in agg.items() if round(c[0] / P) > 0}
    total = sum(k * avg for k, avg in table.values())
    print(f"One pass of the hot path as ncu sees it (`{path.split('/')[-1]}`: `ncu --metrics gpu__time_duration.sum "
          f"--clock-control none -c 400 python bench.py --steps 2 --warmup 3`), from the {P} passes in the capture: "
          f"launches per pass x the kernel's mean duration.  ncu serialises the launches and runs them cold, so only "
          f"the SHARES are meaningful: the live pass is shorter because the five chains (and consecutive passes) overlap.\n")
    print("| launches / pass | µs / pass | share | kernel |\n|---|---|---|---|")
    for n, (k, avg) in sorted(table.items(), key=lambda kv: -kv[1][0] * kv[1][1]):
        print(f"| {k} | {k * avg:.1f} | {100 * k * avg / total:.1f}% | `{n}` |")
    print(f"| {sum(k for k, _ in table.values())} | {total:.1f} | 100% | total |")


if __name__ == "__main__":
    main(sys.argv[1])
