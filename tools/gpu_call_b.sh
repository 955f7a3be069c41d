#!/bin/sh
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.txt
tail -5 gpurun_out/r2n_pytest.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2n_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.1f ms %.4f e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']), d['step_ms']['median'], d['settle']['capacities'], d.get('parity',{}).get('ok'), d['kernels_us'], d['path_roofline']['frac'])
    except Exception as e:
        print(f,'ERR',e)
PY
