"""Markdown table of the per-kernel counters of an `ncu --set full` report (reads `ncu -i REP --page raw --csv`)."""
import csv, subprocess, sys

COLS = [("time ms", "gpu__time_duration.sum", 1.0), ("dram rd GB", "dram__bytes_read.sum", 1.0), ("dram wr GB", "dram__bytes_write.sum", 1.0),
        ("L2->SM GB", "l1tex__m_xbar2l1tex_read_bytes.sum", 1.0), ("lts %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("issue %", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1.0),
        ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("xu %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
        ("fma %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
        ("lsu %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1.0),
        ("regs", "launch__registers_per_thread", 1.0), ("grid", "launch__grid_size", 1.0),
        ("GHz", "sm__cycles_elapsed.avg.per_second", 1.0)]
UNIT_SCALE = {"us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Kbyte": 1e-6, "byte": 1e-9, "Mhz": 1e-3, "Ghz": 1.0}

def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("| # | kernel | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---|" + "---:|" * len(COLS))
    for n, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("b200vad::", "")
        cells = []
        for _, metric, _ in COLS:
            if metric not in hdr:
                cells.append("n/a"); continue
            i = hdr.index(metric)
            try:
                v = float(r[i].replace(",", "")) * UNIT_SCALE.get(units[i], 1.0)
                cells.append(f"{v:.3g}")
            except ValueError:
                cells.append(r[i] or "n/a")
        print(f"| {n} | {name} | " + " | ".join(cells) + " |")

if __name__ == "__main__":
    main(sys.argv[1])
