#!/bin/bash
tag=$1; N=$2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-c4 > gpurun_out/${tag}_bench1.json 2> gpurun_out/${tag}_bench1.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 24 --warmup 3 > gpurun_out/${tag}_bench$N.json 2> gpurun_out/${tag}_bench$N.err; echo "bench$N rc=$?"
python -c "
import json
for f in ('gpurun_out/${tag}_bench1.json','gpurun_out/${tag}_bench$N.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); c=d['config']; s=c.get('strong_c4') or {}
    print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), c['phases_ms'], c['suspension_iterations'][:3], c['deposition_iterations'][:3], round(d['roofline']['frac'],3), s.get('ms_per_step'), s.get('efficiency_vs_1gpu'), (d.get('parity_check') or {}).get('ok'))
"
