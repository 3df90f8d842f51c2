#!/bin/bash
tag=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "not c3" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
SHORT="python bench.py --steps 12 --warmup 3 --no-c4 --no-variants --no-cpu-baseline --no-parity"
for v in 1 0; do
  PBSM3D_SOR_RESIDENT=$v timeout 300 $SHORT > gpurun_out/${tag}_bench_res$v.json 2> gpurun_out/${tag}_bench_res$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench_res$v.json').read().strip().splitlines()[-1]); c=d['config']; print('resident=$v', round(d['ms_per_step'],3), c['phases_ms'], c['deposition_iterations'][:3], c['deposition_residual'])"
done
