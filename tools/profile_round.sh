#!/bin/sh
# Round-end measurement on the GPU box (run from the repo root under gpurun): bench lines, launch list, one full ncu
# capture of the nine kernels of one evaluation.  Outputs in gpurun_out/; tools/ncu_summary.py condenses the raw CSV.
TAG=${1:-r2}          # tools/profile_round.sh r2a  -> gpurun_out/bench_r2a.json ...
mkdir -p gpurun_out
set -x
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_20_5.json 2>> gpurun_out/bench_${TAG}.err
for w in 2clr:1.2 1dwc:1.2 rnaseh:1.2; do n=${w%%:*}; c=${w##*:}; python bench.py --workload $n --cutoff $c --no-cpu-baseline > gpurun_out/bench_${TAG}_${n}_cut12.json 2> gpurun_out/bench_${TAG}_${n}.err; done
python bench.py --workload trpcage --no-cpu-baseline > gpurun_out/bench_${TAG}_trpcage.json 2> gpurun_out/bench_${TAG}_trpcage.err
for w in 2clr 1dwc rnaseh; do python bench.py --md 2000 --workload $w --cutoff 1.2 > gpurun_out/md_${TAG}_${w}_cut12.json 2> gpurun_out/md_${TAG}_${w}.err; done
python bench.py --md 1000 > gpurun_out/md_${TAG}_hivrt.json 2> gpurun_out/md_${TAG}_hivrt.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_(prep|blocklist|tree|born|gb|deriv|finish)" --launch-skip 240 --launch-count 8 -f -o gpurun_out/prof_${TAG}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}_full.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_full_raw.csv 2>/dev/null
tail -c 600 gpurun_out/bench_${TAG}.json
