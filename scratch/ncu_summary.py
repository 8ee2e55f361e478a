"""Markdown summary of an ncu --set full report: python scratch/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys

COLS = [
    ("duration", "gpu__time_duration.sum"),
    ("dram read", "dram__bytes_read.sum"),
    ("dram write", "dram__bytes_write.sum"),
    ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe %", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("warp insts", "smsp__inst_executed.sum"),
    ("long-scoreboard / issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("no-instruction / issue", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
    ("grid", "launch__grid_size"),
]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]


def find(name):
    exact = [i for i, h in enumerate(hdr) if h == name]
    if exact:
        return exact[0]
    for i, h in enumerate(hdr):
        if h.endswith("." + name) and any(r[i] for r in rows[2:]):
            return i
    return None


idx = [(t, find(m)) for t, m in COLS]
ki = hdr.index("Kernel Name")
print("| kernel | " + " | ".join(t for t, _ in idx) + " |")
print("|---|" + "---:|" * len(idx))
for r in rows[2:]:
    cells = []
    for t, i in idx:
        if i is None:
            cells.append("n/a")
            continue
        v = r[i]
        try:
            f = float(v.replace(",", ""))
            v = f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}"
        except ValueError:
            pass
        cells.append(f"{v} {units[i]}".strip() if units[i] not in ("%", "", "inst", "register/thread", "warp") else v)
    name = r[ki].replace("odecol::tc::", "").replace("odecol::", "")
    name = name.split("(")[0]
    print(f"| `{name}` | " + " | ".join(cells) + " |")
