#!/bin/bash
tag=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "not c3" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
SHORT="python bench.py --steps 12 --warmup 3 --no-c4 --no-variants --no-cpu-baseline --no-parity"
for cfg in "PBSM3D_SOR_BLOCKED=1 PBSM3D_SOR_RES_NT=384" "PBSM3D_SOR_BLOCKED=1 PBSM3D_SOR_RES_NT=512" "PBSM3D_SOR_BLOCKED=0 PBSM3D_SOR_RES_NT=384" "PBSM3D_SOR_BLOCKED=0 PBSM3D_SOR_RES_NT=512" "PBSM3D_SOR_BLOCKED=0 PBSM3D_SOR_RES_NT=256"; do
  env $cfg PBSM3D_VERBOSE=1 timeout 300 $SHORT > gpurun_out/${tag}_b.json 2> gpurun_out/${tag}_b.err
  grep 'blocked SOR' gpurun_out/${tag}_b.err | head -1
  python -c "
import json; d=json.loads(open('gpurun_out/${tag}_b.json').read().strip().splitlines()[-1]); c=d['config']; print('$cfg', round(d['ms_per_step'],3), 'dep', round(c['phases_ms']['ms_deposition'],4), c['deposition_iterations'][:3], c['deposition_residual'])"
done
