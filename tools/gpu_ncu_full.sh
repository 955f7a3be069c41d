#!/bin/sh
# one `ncu --set full` capture of the kernels of one evaluation (after the same command has run clean), TAG=$1
TAG=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:k_(prep|blocklist|tree|born|gb|deriv|finish)" --launch-skip 240 --launch-count 8 -f -o gpurun_out/prof_${TAG}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}_full.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_full_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out/prof_${TAG}_full.ncu-rep
