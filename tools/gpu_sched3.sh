#!/bin/bash
# Capture "ad" (1 GPU): one-byte active-set test (act[]) with the flag of the next column prefetched, under three compile-time
# pass distributions (PBSM3D_GS_SCHED_MODE 0 = in-tree library, 2 and 3 = scratch/ builds of the same sources).
tag=$1
mkdir -p gpurun_out
SHORT="--steps 12 --warmup 3 --no-c4 --no-variants --no-cpu-baseline --no-parity"
timeout 200 python bench.py $SHORT > gpurun_out/${tag}_bench_s0.json 2> gpurun_out/${tag}_bench_s0.err; echo "bench s0 rc=$?"
for v in s2 s3; do
  PBSM3D_LIB=$PWD/scratch/libpbsm3d_$v.so timeout 200 python bench.py $SHORT > gpurun_out/${tag}_bench_$v.json 2> gpurun_out/${tag}_bench_$v.err; echo "bench $v rc=$?"
done
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_pytest_s0.log 2>&1; echo "pytest s0 rc=$?"; tail -2 gpurun_out/${tag}_pytest_s0.log
for v in s2 s3; do
  PBSM3D_LIB=$PWD/scratch/libpbsm3d_$v.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "active_set" > gpurun_out/${tag}_pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -2 gpurun_out/${tag}_pytest_$v.log
done
python -c "
import json, glob
for f in sorted(glob.glob('gpurun_out/${tag}_bench_s*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d['config']; r=d['roofline']
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k: round(v,3) for k,v in c['phases_ms'].items()}, c['suspension_iterations'][:2], round(r['frac'],3), (r.get('active_set') or {}).get('column_updates_executed'))
    except Exception as e: print(f, 'ERR', e)
"
