#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

    python profiles/summarize.py r01        # reads gpurun_out/launches_r01.csv, prof_r01_{fwd,bwd}.ncu-rep
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
]


def launch_table(path, out):
    rows = list(csv.reader(open(path)))
    k = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[k]
    iN, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[k + 1:]:
        if len(r) <= iV:
            continue
        try:
            v = float(r[iV].replace(",", ""))
        except ValueError:
            continue
        name = r[iN].split("(")[0][:70]
        tot[name] += v
        cnt[name] += 1
    S = sum(tot.values())
    out.write(f"## launch list ({path}; gpu__time_duration.sum, ncu serialised / cold cache: compare SHARES)\n\n")
    out.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for n, v in tot.most_common(30):
        out.write(f"| `{n}` | {cnt[n]} | {v / 1e3:.1f} | {100 * v / S:.1f}% |\n")
    own = sum(v for n, v in tot.items() if "void k_" in n or n.startswith("k_"))
    out.write(f"\nown kernels (k_*): {100 * own / S:.1f}% of the captured GPU time; the rest are the library GEMMs of the two "
              "feature projections (cuBLAS fp32 SIMT, TF32 off for parity) and elementwise glue.\n\n")


def full_table(rep, out):
    """rep: an .ncu-rep, or the `ncu -i rep --page raw --csv` dump of one (reports can exceed gpurun's 64 MiB return limit)"""
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out.write(f"## ncu --set full ({rep})\n\n")
    for r in rows[2:]:
        out.write(f"### `{r[hdr.index('Kernel Name')]}`\n\n| metric | value |\n|---|---|\n")
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                out.write(f"| {label} | {r[i]} {units[i]} |\n")
        out.write("\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    with open(f"profiles/{tag}_summary.md", "w") as out:
        out.write(f"# ncu summary, round {tag}\n\n"
                  "Launch list and per-slot DRAM traffic: one step of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (the\n"
                  "benchmark sizes, batch 32).  Full capture: `python tools/fwd_stage_times.py --scales 0 --reps 1 --bwd` - the 17\n"
                  "streaming kernels of one forward + backward of the scale-0 block [32,48,256,256].\n"
                  "Numbers under ncu are never bench values (cold cache, serialised); they explain where the time goes.\n\n")
        launch_table(f"gpurun_out/launches_{tag}b.csv", out)
        import json, os
        sp = f"profiles/{tag}_step_slots.json"
        if os.path.exists(sp):
            d = json.load(open(sp))["slots"]
            out.write("## per-slot totals of one step (tools/ncu_slots.py; DRAM bytes = dram__bytes_read.sum + dram__bytes_write.sum)\n\n"
                      "| slot | launches | ncu ms | share of own kernels | DRAM GB |\n|---|---|---|---|---|\n")
            for k, v in sorted(d.items(), key=lambda kv: -kv[1]["ncu_ms"]):
                out.write(f"| {k} | {v['launches']} | {v['ncu_ms']:.3f} | {100 * v['share_of_own_kernels']:.1f}% | {v['dram_bytes'] / 1e9:.3f} |\n")
            out.write("\n")
        full_table(f"gpurun_out/prof_{tag}_stream_raw.csv", out)
    print("wrote", f"profiles/{tag}_summary.md")
