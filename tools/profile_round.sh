#!/bin/sh
# Round-end measurement on the GPU box (run from the repo root under gpurun): bench lines, launch list, one full ncu
# capture of the nine kernels of one evaluation.  Outputs in gpurun_out/; tools/ncu_summary.py condenses the raw CSV.
set -x
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err
for w in 2clr:1.2 1dwc:1.2 rnaseh:1.2; do n=${w%%:*}; c=${w##*:}; python bench.py --workload $n --cutoff $c --no-cpu-baseline > gpurun_out/bench_r1f_${n}_cut12.json 2> gpurun_out/bench_r1f_${n}.err; done
python bench.py --workload trpcage --no-cpu-baseline > gpurun_out/bench_r1f_trpcage.json 2> gpurun_out/bench_r1f_trpcage.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_r1f.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_(prep|blocklist|tree|born|gb|deriv|finish)" --launch-skip 45 --launch-count 9 -f -o gpurun_out/prof_r1f_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1f.log 2>&1
ncu -i gpurun_out/prof_r1f_full.ncu-rep --page raw --csv > gpurun_out/prof_r1f_full_raw.csv 2>/dev/null
tail -c 600 gpurun_out/bench_r1f.json
