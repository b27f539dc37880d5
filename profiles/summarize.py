"""Turns the ncu reports brought back in gpurun_out/ into the tracked summaries of profiles/.
Usage: python profiles/summarize.py r01   (reads gpurun_out/r01_*.ncu-rep, writes profiles/r01_*.txt and traffic.json)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main(prefix):
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for name in sorted(os.listdir(os.path.join(ROOT, "gpurun_out"))):
        if not (name.startswith(prefix) and name.endswith(".ncu-rep")):
            continue
        traffic = {k: v for k, v in traffic.items() if v.get("report") != name}
        hdr, units, rows = raw(os.path.join(ROOT, "gpurun_out", name))
        lines = [f"# ncu --set full --clock-control none, report {name} (one launch of the dominant kernel)"]
        for r in rows:
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    lines.append(f"{k}: {r[i]} {units[i]}")
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            total = float(r[rd]) * UNIT[units[rd]] + float(r[wr]) * UNIT[units[wr]]
            lines.append(f"dram bytes per launch (read+write): {total:.0f}")
            wl = name[len(prefix):].split(".")[0].split("_")[-1]
            kernel = r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1].strip()
            num = lambda key: float(r[hdr.index(key)])
            dur_unit = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[hdr.index("gpu__time_duration.sum")], 1.0)
            rec = {"bytes_per_launch": total, "kernel": kernel, "report": name,
                   "grid": int(num("launch__grid_size")), "duration_us": num("gpu__time_duration.sum") * dur_unit,
                   "warp_inst_executed": num("smsp__inst_executed.sum"), "sm_cycles_elapsed": num("sm__cycles_elapsed.max"),
                   "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"), "launches_in_report": 1}
            prev = traffic.get(wl)
            if prev and prev.get("report") == name:
                # several launches of one sweep in the same report (the two passes of the fused pedigree sweep): totals of the sweep,
                # issue activity weighted by duration
                w0, w1 = prev["duration_us"], rec["duration_us"]
                rec["issue_active_pct"] = (prev["issue_active_pct"] * w0 + rec["issue_active_pct"] * w1) / (w0 + w1)
                for key in ("bytes_per_launch", "duration_us", "warp_inst_executed", "sm_cycles_elapsed"):
                    rec[key] += prev[key]
                rec["launches_in_report"] = prev["launches_in_report"] + 1
            traffic[wl] = rec
        open(os.path.join(ROOT, "profiles", name.replace(".ncu-rep", ".txt")), "w").write("\n".join(lines) + "\n")
        print("\n".join(lines))
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
