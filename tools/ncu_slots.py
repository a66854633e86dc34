#!/usr/bin/env python
"""Fold an ncu launch list (--csv, one row per kernel and metric) of ONE bench step into the bench's kernel slots.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:"^k_" -s 492 -c 164 --csv --log-file gpurun_out/step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-infer4k
    python tools/ncu_slots.py gpurun_out/step.csv profiles/r02_step_slots.json

(164 = own k_* kernels per step at the benchmark sizes in round 2; the first 3 steps are warm-up.)  The JSON gives, per slot, the
number of launches, the summed ncu duration and the summed DRAM bytes: bench.py reads it for `roofline.traffic`."""
import collections
import csv
import json
import re
import sys

FWD = ["fwd_BA", "fwd_X1", "fwd_X2", "fwd_X3"]
BWD = ["bwd_X3", "bwd_X2", "bwd_X2", "bwd_X1", "bwd_BA"]
GW = ["gw_X3", "gw_X2", "gw_X2", "gw_X1", "gw_BA"]
PLANE_BWD = ["bwd_X3", "bwd_X2", "bwd_X1", "bwd_BA"]


def slot_of(name):
    m = re.search(r"\b(k_\w+)(?:<(?:\(int\))?(\d+))?", name)      # also "void <unnamed>::k_proj_tc<0>(...)"
    if not m:
        return None
    k, a = m.group(1), int(m.group(2)) if m.group(2) else 0
    if k == "k_stream_fwd":
        return FWD[a]
    if k == "k_stream_bwd":
        return BWD[a]
    if k in ("k_gw_stage", "k_gw_quad", "k_gw_stream"):
        return GW[a]
    if k == "k_bw2":                      # k_bw2<MODE, COARSE, ...>: X3, X2A, X2B, X1, BA
        return BWD[a]
    if k == "k_fw2":                      # k_fw2<MODE, COARSE, ...>: BA, X1, X2, X3
        return FWD[a]
    if k == "k_proj_tc":                  # MODE 0: activations (forward / dgrad share the kernel), MODE 1: weight gradient
        return "proj_wgrad" if a == 1 else "proj_act"
    if k == "k_proj_wprep":
        return "proj_act"
    if k == "k_block_stage":
        return FWD[a]
    if k == "k_block_bwd_stage":
        return PLANE_BWD[a]
    if k in ("k_block_weights", "k_gtv_coeffs", "k_weights_walk"):
        return "fwd_weights"
    if k in ("k_block_weights_bwd", "k_weights_walk_bwd"):
        return "bwd_weights"
    if k == "k_space_to_depth":
        return "proj_act"
    return "other:" + k


def main(src, dst):
    rows = list(csv.reader(open(src)))
    k = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[k]
    iN, iM, iV, iU, iID = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0}
    out = collections.defaultdict(lambda: {"launches": 0, "ncu_ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
    seen = set()
    for r in rows[k + 1:]:
        if len(r) <= iV:
            continue
        s = slot_of(r[iN])
        if s is None:
            continue
        try:
            v = float(r[iV].replace(",", "")) * scale.get(r[iU], 1.0)
        except ValueError:
            continue
        if (r[iID], "n") not in seen:
            seen.add((r[iID], "n"))
            out[s]["launches"] += 1
        if r[iM].startswith("gpu__time_duration"):
            out[s]["ncu_ms"] += v
        elif r[iM].startswith("dram__bytes_read"):
            out[s]["dram_read_bytes"] += v
        elif r[iM].startswith("dram__bytes_write"):
            out[s]["dram_write_bytes"] += v
    tot = sum(v["ncu_ms"] for v in out.values())
    for v in out.values():
        v["share_of_own_kernels"] = round(v["ncu_ms"] / tot, 4) if tot else None
        v["dram_bytes"] = v["dram_read_bytes"] + v["dram_write_bytes"]
    json.dump({"source": src, "note": "one bench step under ncu (cold cache, serialised): compare shares, not absolutes",
               "slots": dict(sorted(out.items()))}, open(dst, "w"), indent=1)
    for s, v in sorted(out.items(), key=lambda kv: -kv[1]["ncu_ms"]):
        print(f"{s:12s} x{v['launches']:3d} {v['ncu_ms']:8.3f} ms {100 * v['ncu_ms'] / tot:5.1f}%  dram {v['dram_bytes'] / 1e9:7.3f} GB")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
