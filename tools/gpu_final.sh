#!/bin/bash
# Final 1-GPU capture of the round: GPU tests, the full bench line, the steady-state launch list, ncu --set full of the kernels
# that changed since capture C, and one GPU's share of config c5.
tag=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
SHORT="python bench.py --steps 3 --warmup 3 --no-c4 --no-variants --no-cpu-baseline --no-parity"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 60 --csv --log-file gpurun_out/${tag}_launches.csv $SHORT > gpurun_out/${tag}_ncu_launches.log 2>&1
for k in sor_resident_kernel gs_persistent_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${tag}_$k $SHORT > gpurun_out/${tag}_ncu_$k.log 2>&1
  echo "$k rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:slide_sweep_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_slide_sweep_kernel python tools/time_slide.py 708 0 1.0 > gpurun_out/${tag}_ncu_slide.log 2>&1; echo "slide ncu rc=$?"
timeout 600 python bench.py --workload c5shard --steps 6 --warmup 3 --no-cpu-baseline --no-variants --no-c4 > gpurun_out/${tag}_bench_c5shard.json 2> gpurun_out/${tag}_bench_c5shard.err; echo "c5shard rc=$?"
python -c "
import json
for f in ('gpurun_out/${tag}_bench.json','gpurun_out/${tag}_bench_c5shard.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d['config']; print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), c['phases_ms'], c['suspension_iterations'][:3], c['deposition_iterations'][:3], d['roofline']['frac'], c.get('next_rows'))
    except Exception as e: print(f, 'ERR', e)
"
