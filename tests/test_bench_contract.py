"""bench.py's reference arm (CPU, no GPU needed) and its bookkeeping helpers: the JSON contract the driver parses."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_Mpix_per_s" and d["unit"] == "Mpix/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    # the real reference module when oracle/_ref travelled with the snapshot (oracle/vendor_ref.sh), else the oracle port
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "deep_multiscale_GGLR_GGTV_v1x0.py"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def _bench():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def test_slot_names_follow_the_header_enum():
    """bench.py indexes glrgtv_profile_read()'s arrays by position: SLOTS must be include/glrgtv.h's enum, in order."""
    bench = _bench()
    hdr = open(os.path.join(ROOT, "include", "glrgtv.h")).read()
    enum = dict((n.lower(), int(v)) for n, v in re.findall(r"GLRGTV_SLOT_(\w+)\s*=\s*(\d+)", hdr))
    count = enum.pop("count")
    assert count == len(bench.SLOTS)
    assert [enum[s.lower()] for s in bench.SLOTS] == list(range(count))


def test_algorithmic_bytes_bookkeeping():
    bench = _bench()
    assert (bench.DIMS, bench.NGRAPHS, bench.RES) == ([48, 96, 192, 384], [8, 16, 16, 32], 256)
    # hand count for fwd_BA at B = 1: reads y (C) and the half set cT (0.5 x 4G fine + a quarter of that coarse),
    # writes bA (C), 4 bytes each, per pixel of each scale
    by_hand = 4 * sum((2 * C + 0.625 * 4 * G) * (256 >> s) ** 2 for s, (C, G) in enumerate(zip([48, 96, 192, 384], [8, 16, 16, 32])))
    assert bench.algorithmic_bytes("fwd_BA", 1) == by_hand
    for s in bench.SLOTS:                                   # every slot has a model; it is linear in the batch
        b1 = bench.algorithmic_bytes(s, 1)
        assert (b1 > 0 or s == "gw") and bench.algorithmic_bytes(s, 32) == 32 * b1
    # the X2 stages move the most (two solver passes share the slot in the backward)
    assert max(bench.SLOTS, key=lambda s: bench.algorithmic_bytes(s, 1)) == "bwd_X2"
