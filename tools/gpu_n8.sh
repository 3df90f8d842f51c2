#!/bin/bash
tag=$1; N=$2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 24 --warmup 3 --no-c4 --no-variants --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1]); c=d['config']; print('N=$N', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), c['phases_ms'], c['suspension_iterations'][:3], c['deposition_iterations'][:3], (d.get('parity_check') or {}).get('ok'), c.get('providers'))"
