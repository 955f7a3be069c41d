#!/bin/sh
mkdir -p gpurun_out
timeout 600 python bench.py --md 1000 > gpurun_out/r2_md_hivrt.json 2> gpurun_out/r2_md_hivrt.err
AGBNP_B200_TREE_GROUP=0 timeout 600 python bench.py --md 1000 > gpurun_out/r2_md_hivrt_nogroup.json 2> gpurun_out/r2_md_hivrt_ng.err
python tools/quick_time.py hivrt 2>&1 | grep -A1 "method=0" | grep k_tree
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_md_hivrt*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.1f ns/day  %.4f ms/step  T=%.0f K'%(d['value'],d['ms_per_step'],d['temperature_K']), d['during_timed_region'], d['capacities'], d['kernels_us'])
    except Exception as e: print(f,'ERR',e)
PY
