#!/bin/bash
# Capture "ab" (1 GPU): the four pass distributions of the persistent line solver (PBSM3D_GS_SCHED) at c2, with the active set.
tag=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
SHORT="--steps 12 --warmup 3 --no-c4 --no-variants --no-cpu-baseline --no-parity"
for cfg in "1 1" "2 1" "3 1" "3 0"; do
  set -- $cfg
  PBSM3D_GS_SCHED=$1 PBSM3D_ACTIVE_SET=$2 timeout 200 python bench.py $SHORT > gpurun_out/${tag}_bench_s$1_as$2.json 2> gpurun_out/${tag}_bench_s$1_as$2.err; echo "bench sched=$1 as=$2 rc=$?"
done
for sc in 2 3; do
  PBSM3D_GS_SCHED=$sc timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "active_set or step_matches or golden_sequence or layer_count" > gpurun_out/${tag}_pytest_s$sc.log 2>&1; echo "pytest sched=$sc rc=$?"; tail -2 gpurun_out/${tag}_pytest_s$sc.log
done
python -c "
import json, glob
for f in sorted(glob.glob('gpurun_out/${tag}_bench_s*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d['config']; r=d['roofline']
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k: round(v,3) for k,v in c['phases_ms'].items()}, c['suspension_iterations'][:2], round(r['frac'],3), (r.get('active_set') or {}).get('column_updates_executed'))
    except Exception as e: print(f, 'ERR', e)
"
