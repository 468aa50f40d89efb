# usage (under gpurun): bash tools/ncu_flash_one.sh <tag>   -- source-level ncu capture of the frequency self-attention launch (B = 32)
set -e
t=${1:-fa}
python tools/flash_one.py > gpurun_out/fa_one.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 2 -c 1 -o /tmp/$t -f python tools/flash_one.py > gpurun_out/ncu_$t.log 2>&1
ncu -i /tmp/$t.ncu-rep --page source --csv > gpurun_out/${t}_source.csv
ncu -i /tmp/$t.ncu-rep --page raw --csv > gpurun_out/${t}_raw.csv
