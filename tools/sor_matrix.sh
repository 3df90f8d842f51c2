#!/bin/bash
tag=$1
for v in 0 1 2 3; do PBSM3D_SOR_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('sor variant $v', round(d['ms_per_step'],3), {k:round(v,3) for k,v in c['phases_ms'].items()}, c['suspension_iterations'][:3], c['deposition_iterations'][:3], 'launches', d['gpu_launches'], 'calm', round(c['calm_step_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'roof', round(d['roofline']['frac'],3), d['roofline'].get('per_launch'))
"; done | tee gpurun_out/${tag}_sor_matrix.txt
PBSM3D_FP32_X=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('no fp32 x', round(d['ms_per_step'],3), {k:round(v,3) for k,v in c['phases_ms'].items()}, c['suspension_iterations'][:3], 'roof', round(d['roofline']['frac'],3), d['roofline'].get('per_launch'))
" | tee -a gpurun_out/${tag}_sor_matrix.txt
