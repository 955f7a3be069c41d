#!/bin/sh
# `ncu --set full` capture of selected kernels of one settled evaluation: tools/gpu_ncu_kernels.sh TAG "born|deriv" NKERNELS
# (30 evaluations are skipped: NKERNELS = how many kernels of one evaluation the regex matches)
TAG=${1:-r2}; RE=${2:-"born|deriv"}; NK=${3:-2}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:^k_(${RE})$" --launch-skip $((30*NK)) --launch-count $NK -f -o gpurun_out/prof_${TAG} python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log
ls -la gpurun_out/prof_${TAG}.ncu-rep
