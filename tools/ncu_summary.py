"""Condenses an `ncu -i X.ncu-rep --page raw --csv` dump into the handful of counters the roofline
discussion needs (one block per profiled launch): duration, DRAM bytes, pipe utilisation, issue
rate, occupancy and the top warp-stall reasons.

    python tools/ncu_summary.py gpurun_out/prof_r02/nbody.raw.csv [more.raw.csv ...]"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA sub-pipe %"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA cycles active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path, newline="")))
        if len(rows) < 3:
            print(f"## {path}: empty")
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"## {path}")
        for r in rows[2:]:
            name = r[col["Kernel Name"]].split("(")[0]
            print(f"kernel {name}  grid {r[col['Grid Size']]}  block {r[col['Block Size']]}")
            for key, label in KEYS:
                if key in col and r[col[key]] != "":
                    print(f"    {label:36s} {r[col[key]]:>16s} {units[col[key]]}")
            stalls = []
            for h, i in col.items():
                if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                    try:
                        stalls.append((float(r[i].replace(",", "")), h[len("smsp__pcsamp_warps_issue_stalled_"):]))
                    except ValueError:
                        pass
            tot = sum(s for s, _ in stalls)
            if tot > 0:
                top = sorted(stalls, reverse=True)[:5]
                print("    warp-state samples: " + ", ".join(f"{n} {100 * s / tot:.0f}%" for s, n in top))
        print()


if __name__ == "__main__":
    main()
