#!/bin/bash
# Barrier-discipline check of the kernel SOURCES without a GPU: the emulation runs every CUDA thread of a block as a cooperative fiber;
# GLRGTV_EMU_SCHED permutes the order in which the fibers run between two synchronisation points (reverse, or a fresh random
# permutation per sweep).  A kernel whose shared-memory hand-overs are all separated by a barrier or a warp shuffle produces the same
# result under every order; one that relies on "the lower thread ran first" fails its parity test under some order.
#   tools/emu_races.sh [pytest args ...]      default: the fiber-based emulation test modules
set -e
cd "$(dirname "$0")/.."
python -m imagerestoration_development_unrolling_b200.build --emu > /dev/null
if [ $# -eq 0 ]; then set -- tests/test_emu_stream.py tests/test_emu_host_cnn.py; fi
for sched in reverse random:1 random:2 random:3; do
    echo "== GLRGTV_EMU_SCHED=$sched"
    GLRGTV_EMU_SCHED=$sched python -m pytest "$@" -x -q -p no:cacheprovider | tail -1
done
# asynchronous copies completing at the LATEST legal moment (at the covering wait_group / the first mbarrier wait) instead of at issue:
# catches a missing or too-shallow wait, which the default mode (earliest completion: catches write-after-read hazards) cannot see
for sched in forward reverse random:4; do
    echo "== GLRGTV_EMU_ASYNC=late GLRGTV_EMU_SCHED=$sched"
    GLRGTV_EMU_ASYNC=late GLRGTV_EMU_SCHED=$sched python -m pytest "$@" tests/test_emu_block.py -x -q -p no:cacheprovider | tail -1
done
