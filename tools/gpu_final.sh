#!/bin/bash
# Round-end evidence run (one GPU): smoke, tests (-s: parity figures), bench lines (own arm with cpu_baseline, STFT round trip, reference
# arm) and ONE ncu pass (launch durations + DRAM traffic of a batch-32 forward).   usage (under gpurun): bash tools/gpu_final.sh <tag>
tag=${1:-final}
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_${tag}.log
timeout 900 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_${tag}.json
timeout 300 python bench.py --config stft > gpurun_out/bench_stft_${tag}.json 2> gpurun_out/bench_stft_${tag}.err; echo "stft exit $?"; cut -c1-200 gpurun_out/bench_stft_${tag}.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_ref_${tag}.json
timeout 300 python tools/one_forward.py 32 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_traffic_${tag}.csv python tools/one_forward.py 32 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"
