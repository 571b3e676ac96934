"""Turn an .ncu-rep (ncu --set full) into a small markdown/CSV summary for profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_kernels
"""
import csv
import io
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf"),
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, s) for m, s in WANT if m in idx]
    lines = ["| kernel | " + " | ".join(s for _, s in cols) + " |", "|---|" + "---|" * len(cols)]
    seen = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("ldiff::", "")
        seen[name] = r                                   # keep the last (warm) launch of each kernel
    for name, r in seen.items():
        cells = []
        for m, _ in cols:
            v, u = r[idx[m]], units[idx[m]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip())
        lines.append(f"| `{name[:70]}` | " + " | ".join(cells) + " |")
    open(out + ".md", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
