#!/bin/bash
# Round-end evidence run (one GPU): smoke, tests, bench lines (own arm with cpu_baseline, reference arm), STFT microbench,
# launch list and DRAM-traffic capture of one batch-32 forward.   usage (under gpurun): bash tools/gpu_final.sh <tag>
tag=${1:-final}
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_ref.log
timeout 300 python tools/stft_bench.py > gpurun_out/stft_bench.json 2> gpurun_out/stft_bench.err; echo "stft exit $?"
timeout 300 python tools/one_forward.py 32 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${tag}.csv python tools/one_forward.py 32 > gpurun_out/ncu.log 2>&1; echo "ncu launches exit $?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/traffic_${tag}.csv python tools/one_forward.py 32 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
