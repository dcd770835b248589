"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (run here, no GPU needed).

    python scripts/summarize_profiles.py r1
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list (ncu --metrics gpu__time_duration.sum of `bench.py --steps 3 --warmup 3 --no-cpu-baseline`)
launches = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = next(r for r in rows if r[0] == "ID")
    data = [dict(zip(hdr, r)) for r in rows if r[0].isdigit()]
    agg = collections.OrderedDict()
    for d in data:
        agg.setdefault(d["Kernel Name"].split("(")[0][:70], []).append(float(d["Metric Value"].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_bench_launches.md"), "w") as f:
        f.write(f"# ncu launch list, `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-aux --no-eager` ({tag})\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and "
                "serialised: compare SHARES.\n\n| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / 1e3:.1f} | {100 * sum(v) / total:.1f} % |\n")
    print("wrote launches summary", len(data), "launches")

WANT = None


def summarize(rep, title, cmd, out_md, out_csv):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\n`{cmd}`\nDurations under ncu are cold-cache.\n\n")
        names = [d[idx["Kernel Name"]].split("(")[0].replace("void ", "") for d in data]
        f.write("| metric | " + " | ".join(f"`{n}`" for n in names) + " |\n|---|" + "---|" * len(names) + "\n")
        for key, label in WANT:
            if key not in idx:
                continue
            cells = []
            for d in data:
                v = d[idx[key]]
                try:
                    cells.append(f"{float(v.replace(',', '')):.4g} {units[idx[key]]}")
                except ValueError:
                    cells.append(v)
            f.write(f"| {label} | " + " | ".join(cells) + " |\n")
    with open(out_csv, "w") as f:
        keep = [i for i, h in enumerate(hdr) if any(h.startswith(p) for p in (
            "Kernel Name", "gpu__time", "launch__", "sm__", "smsp__issue", "smsp__warps", "smsp__average_warps_issue_stalled",
            "dram__bytes", "gpu__dram", "l1tex__data_bank", "lts__t_bytes", "smsp__inst_executed.sum"))]
        w = csv.writer(f)
        for r in [hdr, units] + data:
            w.writerow([r[i] for i in keep])
    print("wrote", out_md, len(data), "kernels")


# ---- full capture of the BL kernels
rep = os.path.join(ROOT, "gpurun_out", f"bl_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    want = [
        ("gpu__time_duration.sum", "duration"),
        ("launch__grid_size", "grid"),
        ("launch__registers_per_thread", "regs/thread"),
        ("launch__waves_per_multiprocessor", "waves/SM"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("dram__bytes_read.sum", "dram read"),
        ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps/scheduler"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle (warps/issue)"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ]
    want += [
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "TC pipe inst %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("lts__t_bytes.sum", "L2 bytes"),
    ]
    WANT = want
    aux = os.path.join(ROOT, "gpurun_out", f"aux_{tag}.ncu-rep")
    if os.path.exists(aux):
        summarize(aux, f"ncu --set full, ISW module fwd+bwd (B=8, C=256, HW=6400) and one density map (2048x2048, 25 000 heads; adaptive, then fixed sigma) ({tag})",
                  "ncu --set full --clock-control none --import-source on -k 'regex:isw_|dmap_' -s 23 -c 23 python scripts/profile_aux.py",
                  os.path.join(out_dir, f"{tag}_aux_ncu_summary.md"), os.path.join(out_dir, f"{tag}_aux_ncu_raw.csv"))
    with open(os.path.join(out_dir, f"{tag}_bl_ncu_summary.md"), "w") as f:
        f.write(f"# ncu --set full, fused Bayesian loss, BASELINE config 3 ({tag})\n\n"
                "`ncu --set full --clock-control none --import-source on -k regex:bl_ -s 7 -c 7 python scripts/profile_bl.py`\n"
                "(second step of two; 16 images, 192x256 grid, 49 697 heads).  Durations under ncu are cold-cache.\n\n")
        names = [d[idx["Kernel Name"]].split("(")[0].replace("void ", "") for d in data]
        f.write("| metric | " + " | ".join(f"`{n}`" for n in names) + " |\n|---|" + "---|" * len(names) + "\n")
        for key, label in want:
            if key not in idx:
                continue
            cells = []
            for d in data:
                v = d[idx[key]]
                try:
                    cells.append(f"{float(v.replace(',', '')):.4g} {units[idx[key]]}")
                except ValueError:
                    cells.append(v)
            f.write(f"| {label} | " + " | ".join(cells) + " |\n")
    with open(os.path.join(out_dir, f"{tag}_bl_ncu_raw.csv"), "w") as f:
        keep = [i for i, h in enumerate(hdr) if any(h.startswith(p) for p in (
            "Kernel Name", "gpu__time", "launch__", "sm__", "smsp__issue", "smsp__warps", "smsp__average_warps_issue_stalled",
            "dram__bytes", "gpu__dram", "l1tex__data_bank", "lts__t_bytes", "smsp__inst_executed.sum"))]
        w = csv.writer(f)
        for r in [hdr, units] + data:
            w.writerow([r[i] for i in keep])
    print("wrote BL ncu summary for", len(data), "kernels")
