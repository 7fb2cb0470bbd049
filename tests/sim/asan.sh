#!/bin/bash
# The gpu tests and the multi-rank check on an AddressSanitizer build of the simulated library: every out-of-bounds access of
# a kernel (or of the host code) on these cases is reported -- the closest thing to compute-sanitizer memcheck without a GPU.
# usage: bash tests/sim/asan.sh            (from the repository root; ~3 min)
set -u
export MOKAB_SIM_ASAN=1 MOKAB_SIM_BUILD=${MOKAB_SIM_BUILD:-/tmp/mokab_sim_asan}
export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:log_path=/tmp/mokab_sim_asan_report
rm -f /tmp/mokab_sim_asan_report*
# libstdc++ has to be there when ASan initialises (python does not link it), or throwing an exception trips ASan itself
PRE="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so)"
python tests/sim/build.py > /dev/null || exit 1
LD_PRELOAD="$PRE" MOKAB_SIM=1 python -m pytest tests -x -q -m gpu -p no:cacheprovider -k "not full_size and not large_mesh" | tail -n 2
LD_PRELOAD="$PRE" python tests/sim/check_decomposed.py --cases suite --seeds 1 --policies lazy,random | tail -n 1
if ls /tmp/mokab_sim_asan_report* > /dev/null 2>&1; then echo "ASAN REPORTS:"; head -n 30 /tmp/mokab_sim_asan_report*; exit 1; fi
echo "no AddressSanitizer report"
