#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

    python profiles/summarize.py r01        # reads gpurun_out/launches_r01.csv, prof_r01_{fwd,bwd}.ncu-rep
"""
import collections
import csv
import re
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
    own = sum(v for n, v in tot.items() if re.search(r"(^|::|\s)k_", n))
    lib = [n for n in tot if "cutlass" in n or "sgemm" in n or "cublas" in n.lower() or "gemm" in n.lower()]
    out.write(f"\nown kernels (k_*): {100 * own / S:.1f}% of the captured GPU time; library GEMM kernels in the list: "
              f"{', '.join(lib) if lib else 'none'}; the rest is torch elementwise glue (fills of the gradient accumulators, adds).\n\n")


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
    import json, os
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    # r01: gpurun_out/launches_r01b.csv + prof_r01_stream_raw.csv; r02: r02_launches_all.csv + r02_{bw2,fwd,proj}_raw.csv
    launches = f"gpurun_out/launches_{tag}b.csv" if tag == "r01" else f"gpurun_out/{tag}_launches_all.csv"
    raws = [f"gpurun_out/prof_{tag}_stream_raw.csv"] if tag == "r01" else [f"gpurun_out/{tag}_{k}_raw.csv" for k in ("bw2", "fwd", "ww", "proj")]
    with open(f"profiles/{tag}_summary.md", "w") as out:
        out.write(f"# ncu summary, round {tag}\n\n"
                  "Launch list: the whole process `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-infer4k`\n"
                  "(benchmark sizes, batch 32: 3 warm-up + 1 timed + 3 end-to-end steps).  Per-slot DRAM traffic: the timed step of the same\n"
                  "command.  Full captures (`ncu --set full --clock-control none --import-source on`): the scale-0 launches of the same\n"
                  "command (`tools/gpu_r02_final.sh`).  Numbers under ncu are never bench values (cold cache, serialised); they\n"
                  "explain where the time goes.\n\n")
        launch_table(launches, out)
        sp = f"profiles/{tag}_step_slots.json"
        if os.path.exists(sp):
            d = json.load(open(sp))["slots"]
            out.write("## per-slot totals of one step (tools/ncu_slots.py; DRAM bytes = dram__bytes_read.sum + dram__bytes_write.sum)\n\n"
                      "| slot | launches | ncu ms | share of own kernels | DRAM GB |\n|---|---|---|---|---|\n")
            for k, v in sorted(d.items(), key=lambda kv: -kv[1]["ncu_ms"]):
                out.write(f"| {k} | {v['launches']} | {v['ncu_ms']:.3f} | {100 * v['share_of_own_kernels']:.1f}% | {v['dram_bytes'] / 1e9:.3f} |\n")
            out.write("\n")
        for r in raws:
            if os.path.exists(r):
                full_table(r, out)
    print("wrote", f"profiles/{tag}_summary.md")
