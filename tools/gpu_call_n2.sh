#!/bin/sh
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
echo "rc=$?"
tail -3 gpurun_out/r2_scale_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_scale_n$N.json').read().strip().splitlines()[-1])
print('N=$N value %.1f ms %.4f e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']), d['step_ms'], d.get('parity'), d['kernels_us'], d.get('replica_mode'), d.get('collectives_per_step'))
PY
