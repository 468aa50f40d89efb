set -e
python tools/one_forward.py 32 1 > gpurun_out/of.log 2>&1
for s in 1 14 64; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc -s $s -c 1 -o /tmp/g$s -f python tools/one_forward.py 32 1 > gpurun_out/ncu_g$s.log 2>&1
  ncu -i /tmp/g$s.ncu-rep --page source --csv > gpurun_out/g${s}_source.csv
  ncu -i /tmp/g$s.ncu-rep --page raw --csv > gpurun_out/g${s}_raw.csv
done
ls -la gpurun_out/g*_source.csv
