#!/bin/bash
# Memory check of the kernel SOURCES without a GPU: the g++ emulation build under AddressSanitizer.  Global tensors are torch CPU
# allocations (redzoned by the preloaded runtime) and each launch's dynamic shared memory is a heap block of exactly the requested
# size (-DGLRGTV_EMU_EXACT_SMEM), so out-of-bounds reads / writes of either kind abort the test that triggered them.
#   tools/emu_asan.sh [pytest args ...]      default: all emulation test modules
set -e
cd "$(dirname "$0")/.."
python -m imagerestoration_development_unrolling_b200.build --emu --asan
ASAN=$(gcc -print-file-name=libasan.so)
if [ $# -eq 0 ]; then set -- tests/test_emu_ops.py tests/test_emu_block.py tests/test_emu_host_cnn.py tests/test_emu_stream.py; fi
LOG=$(mktemp -d)/asan
set +e
GLRGTV_EMU_ASAN=1 LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1:detect_stack_use_after_return=0:log_path=$LOG \
    python -m pytest "$@" -x -q -p no:cacheprovider
rc=$?
if grep -qs "ERROR: AddressSanitizer" $LOG.*; then   # the sanitizer aborts the interpreter: its report is in the log, not in pytest's output
    echo "---- AddressSanitizer report ----"; head -40 $LOG.*; [ $rc -eq 0 ] && rc=1
fi
exit $rc
