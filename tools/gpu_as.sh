#!/bin/bash
# Capture "aa" (2-GPU box): parity + multi-GPU tests of the active-set line solver, 1-GPU bench with and without it, 2-GPU bench.
tag=$1; N=${2:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
SHORT="--steps 12 --warmup 3 --no-c4 --no-variants --no-cpu-baseline --no-parity"
for as in 1 0; do
  CUDA_VISIBLE_DEVICES=0 PBSM3D_ACTIVE_SET=$as timeout 300 python bench.py $SHORT > gpurun_out/${tag}_bench1_as$as.json 2> gpurun_out/${tag}_bench1_as$as.err; echo "bench1 as=$as rc=$?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 24 --warmup 3 --no-c4 --no-variants --no-cpu-baseline > gpurun_out/${tag}_bench$N.json 2> gpurun_out/${tag}_bench$N.err; echo "bench$N rc=$?"
python -c "
import json
for f in ('gpurun_out/${tag}_bench1_as1.json','gpurun_out/${tag}_bench1_as0.json','gpurun_out/${tag}_bench$N.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d['config']; r=d['roofline']
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k: round(v,3) for k,v in c['phases_ms'].items()}, c['suspension_iterations'][:3], c['deposition_iterations'][:3], round(r['frac'],3), (r.get('active_set') or {}).get('column_updates_executed'), (d.get('parity_check') or {}).get('ok'))
    except Exception as e: print(f, 'ERR', e)
"
