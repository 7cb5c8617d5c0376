"""profiles/rNN_lu_dominant_launch.json from an `ncu --set full ... --page raw --csv` dump of the triangular
bulk update (update_kernel_t<true>): the longest launch = the first outer panel of the 20k system.  bench.py
reads `traffic` / `dominant_launch` of its JSON line from that file (never literals).

    python tools/lu_dominant_launch.py gpurun_out/prof/lu_update_tri.raw.csv profiles/r02_lu_dominant_launch.json"""
import csv, json, subprocess, sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def main(raw, out):
    rows = list(csv.reader(open(raw, newline="")))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        s = r[col[key]].replace(",", "")
        return float(s) * SCALE.get(units[col[key]], 1.0)

    best = max(rows[2:], key=lambda r: val(r, "gpu__time_duration.sum"))
    grid = best[col["Grid Size"]]
    gx, gy = [int(x) for x in grid.strip("()").replace(" ", "").split(",")[:2]]
    tiles = sum(min(gx, 2 * by + 2) for by in range(gy))  # tiles that intersect the lower triangle
    K = 1024
    flop = tiles * 128 * 64 * K * 2
    ms = val(best, "gpu__time_duration.sum")
    rd, wr = val(best, "dram__bytes_read.sum"), val(best, "dram__bytes_write.sum")
    alg = tiles * 128 * 64 * 8 * 2 + (gy * 128 * K + (gx * 64) * K) * 8  # C read + written once, packed operands once
    stalls = {}
    for h, i in col.items():
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                stalls[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(best[i].replace(",", ""))
            except ValueError:
                pass
    tot = sum(stalls.values()) or 1.0
    top = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:4]}

    def opt(key):
        return float(best[col[key]].replace(",", "")) if key in col and best[col[key]] != "" else None

    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    rec = {
        "kernel": "scb::update_kernel_t<true> (triangular bulk update, K = 1024), first outer panel of the 19968 system",
        "grid": grid, "tiles_updated": tiles, "duration_ms": ms, "flop": flop, "tflops": flop / (ms * 1e-3) * 1e-12,
        "dram_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "algorithmic_bytes": alg,
        "traffic_over_algorithmic": (rd + wr) / alg,
        "l2_hit_rate_pct": opt("lts__t_sector_hit_rate.pct"),
        "dmma_subpipe_pct": opt("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
        "sm_throughput_pct": opt("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": opt("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": opt("launch__registers_per_thread"),
        "top_stalls": top,
        "capture": "ncu --set full --clock-control none --kernel-name-base mangled -k regex:update_kernel_tILb1 -c 2 "
                   "python tools/run_stage.py --stage getrf --n 20164 --sym 1 (tools/collect_profiles.sh lu)",
        "git_head": head,
    }
    json.dump({"symmetric": rec}, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
