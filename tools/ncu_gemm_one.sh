# usage: bash tools/ncu_gemm_one.sh <gemm_tc launch index> [kernel regex]   -- source-level ncu capture of one launch of a batch-32 forward
set -e
s=$1; k=${2:-gemm_tc}
python tools/one_forward.py 32 1 > gpurun_out/of.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s $s -c 1 -o /tmp/g$s -f python tools/one_forward.py 32 1 > gpurun_out/ncu_g$s.log 2>&1
ncu -i /tmp/g$s.ncu-rep --page source --csv > gpurun_out/g${s}_source.csv
ncu -i /tmp/g$s.ncu-rep --page raw --csv > gpurun_out/g${s}_raw.csv
