#!/bin/sh
# round 2, call A: new parity tests + reproducibility of the driver's bench invocation
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.txt
tail -5 gpurun_out/r2a_pytest.txt
for i in 1 2 3; do
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 $( [ $i -gt 1 ] && echo --no-cpu-baseline ) > gpurun_out/r2a_bench_20_$i.json 2> gpurun_out/r2a_bench_20_$i.err
done
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2a_bench_200.json 2> gpurun_out/r2a_bench_200.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_torchrun1.json 2> gpurun_out/r2a_bench_torchrun1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.1f ms %.4f e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']), d['step_ms'], d['settle'], d.get('parity',{}).get('ok'), d['kernels_us'], d['path_roofline']['frac'])
    except Exception as e:
        print(f,'ERR',e)
PY
