#!/bin/bash
tag=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
SHORT="python bench.py --steps 12 --warmup 3 --no-c4 --no-variants --no-cpu-baseline --no-parity"
for v in 512 384 256; do
  PBSM3D_SOR_RES_NT=$v timeout 300 $SHORT > gpurun_out/${tag}_bench_nt$v.json 2> gpurun_out/${tag}_bench_nt$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench_nt$v.json').read().strip().splitlines()[-1]); c=d['config']; print('nt=$v', round(d['ms_per_step'],3), c['phases_ms'], c['deposition_iterations'][:3], c['deposition_residual'])"
done
