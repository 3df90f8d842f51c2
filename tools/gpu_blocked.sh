#!/bin/bash
# Capture "ag" (1 GPU): the plain sweep with one contiguous run of columns per block (PBSM3D_GS_BLOCKED, default) against the
# strided 512-column chunks, same box; parity tests on the default.
tag=$1
mkdir -p gpurun_out
SHORT="--steps 12 --warmup 3 --no-c4 --no-variants --no-cpu-baseline --no-parity"
for b in 1 0; do
  PBSM3D_GS_BLOCKED=$b timeout 120 python bench.py $SHORT > gpurun_out/${tag}_bench_blk$b.json 2> gpurun_out/${tag}_bench_blk$b.err; echo "bench blocked=$b rc=$?"
done
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
python -c "
import json
for b in (1,0):
    d=json.loads(open('gpurun_out/${tag}_bench_blk%d.json'%b).read().strip().splitlines()[-1]); c=d['config']
    print('blocked',b, round(d['ms_per_step'],3), {k: round(v,3) for k,v in c['phases_ms'].items()}, c['suspension_iterations'][:2], round(d['roofline']['frac'],3))
"
