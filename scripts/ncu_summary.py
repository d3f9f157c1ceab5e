"""Per-kernel key metrics (time, DRAM bytes, L1/L2 hit rates, issue utilisation, top stall reasons) of every launch in an
.ncu-rep captured with --set full.  usage: python scripts/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}


def g(r, name, default=None):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"): return default
    try: return float(r[i].replace(",", ""))
    except ValueError: return r[i]


STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "mio_throttle", "lg_throttle", "branch_resolving",
          "no_instruction", "not_selected", "dispatch_stall", "membar", "sleeping", "drain", "tex_throttle", "imc_miss"]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    name = r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")
    t = g(r, "gpu__time_duration.sum", 0.0)
    unit = rows[1][col["gpu__time_duration.sum"]]
    t_us = t / 1e3 if unit in ("ns", "nsecond") else (t * 1e3 if unit in ("ms", "msecond") else t)
    rd, wr = g(r, "dram__bytes_read.sum", 0.0), g(r, "dram__bytes_write.sum", 0.0)
    def to_bytes(v, name):
        u = rows[1][col[name]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    rd, wr = to_bytes(rd, "dram__bytes_read.sum"), to_bytes(wr, "dram__bytes_write.sum")
    st = []
    for s in STALLS:
        v = g(r, f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio")
        if isinstance(v, float): st.append((v, s))
    st.sort(reverse=True)
    print(f"{name}: {t_us:.1f} us | grid {r[col['Grid Size']]} x block {r[col['Block Size']]} | regs {g(r, 'launch__registers_per_thread')} | "
          f"DRAM {rd / 1e6:.1f} MB rd + {wr / 1e6:.1f} MB wr = {(rd + wr) / (t_us * 1e-6) / 1e9:.0f} GB/s | "
          f"L1 hit {g(r, 'l1tex__t_sector_hit_rate.pct')} % L2 hit {g(r, 'lts__t_sector_hit_rate.pct')} % | "
          f"warp inst {g(r, 'smsp__inst_executed.sum')} | issue active {g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')} % | "
          f"warps active {g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')} % | stalls/issue: "
          + ", ".join(f"{s} {v:.1f}" for v, s in st[:4]))
