#!/bin/bash
# One GPU round: parity tests, end-to-end parity line, bench line, launch list of one batch-32 forward.
# usage (under gpurun): bash tools/gpu_check.sh <tag>
tag=${1:-x}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest.log
timeout 300 python tools/stage_parity.py --precision bf16 --B 1 --L 264600 > gpurun_out/sp_bf16.log 2>&1; grep -E "^OUT" gpurun_out/sp_bf16.log
ATHTD_PROFILE_DUMP=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-160 gpurun_out/bench.log
timeout 300 python tools/one_forward.py 32 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${tag}.csv python tools/one_forward.py 32 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"
