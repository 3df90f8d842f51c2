#!/bin/bash
tag=$1; N=$2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -k "not c3" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for v in 1 0; do
PBSM3D_SOR_RESIDENT=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$v bench.py --gpus $N --steps 12 --warmup 3 --no-c4 --no-variants --no-cpu-baseline > gpurun_out/${tag}_bench_res$v.json 2> gpurun_out/${tag}_bench_res$v.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench_res$v.json').read().strip().splitlines()[-1]); c=d['config']; print('N=$N resident=$v', round(d['ms_per_step'],3), c['phases_ms'], c['deposition_iterations'][:3], (d.get('parity_check') or {}).get('ok'), (d.get('parity_check') or {}).get('vs_single_rank_rel_l2'))"
done
