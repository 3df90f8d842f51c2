#!/bin/bash
# Capture "ah" (1 GPU, the round's last GPU seconds): neighbour slots of a thread's next column loaded one column ahead
# (PBSM3D_NBS_PREFETCH, default on) against loading them when needed, same box; a parity subset on the default.
tag=$1
mkdir -p gpurun_out
SHORT="--steps 12 --warmup 3 --no-c4 --no-variants --no-cpu-baseline --no-parity"
for b in 1 0; do
  PBSM3D_NBS_PREFETCH=$b timeout 60 python bench.py $SHORT > gpurun_out/${tag}_bench_pf$b.json 2> gpurun_out/${tag}_bench_pf$b.err; echo "bench prefetch=$b rc=$?"
done
python -c "
import json
for b in (1,0):
    d=json.loads(open('gpurun_out/${tag}_bench_pf%d.json'%b).read().strip().splitlines()[-1]); c=d['config']
    print('prefetch',b, round(d['ms_per_step'],3), {k: round(v,3) for k,v in c['phases_ms'].items()}, c['suspension_iterations'][:2], round(d['roofline']['frac'],3))
"
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_matches or golden_sequence or active_set or fp32_sweep or layer_count" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
