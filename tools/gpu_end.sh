#!/bin/bash
# Capture "ae" (1 GPU), the shipped code of the round: all GPU tests, the full bench line, and ncu --set full of the persistent
# suspension solve with and without the active set (DRAM traffic per launch for bench.py's roofline.traffic).
tag=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
SHORT="python bench.py --steps 3 --warmup 3 --no-c4 --no-variants --no-cpu-baseline --no-parity"
for as in 1 0; do
  PBSM3D_ACTIVE_SET=$as timeout 200 ncu --set full --clock-control none --import-source on -k regex:gs_persistent_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_gs_persistent_as$as $SHORT > gpurun_out/${tag}_ncu_as$as.log 2>&1
  echo "ncu as=$as rc=$?"
done
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1]); c=d['config']
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), c['phases_ms'], c['suspension_iterations'][:3], c['deposition_iterations'][:3], d['roofline']['frac'], d['roofline'].get('active_set'), c.get('calm_step_ms'), (c.get('strong_c4') or {}).get('ms_per_step'), (c.get('variants') or {}).get('default_block', {}).get('ms_per_step'))
"
