#!/bin/bash
# Capture "af" (1 GPU): the automatic switch of the line solver's active set (used when <= 40 % of the faces have a right-hand
# side): parity tests and a bench line with the patchy-saltation variant.
tag=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 12 --warmup 3 --no-c4 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1]); c=d['config']
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), c['phases_ms'], d['roofline']['frac'], d['roofline'].get('active_set',{}).get('on'))
print(json.dumps(c['variants'].get('patchy_saltation'), indent=1))
"
