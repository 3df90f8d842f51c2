#!/bin/bash
# One gpurun call (1 GPU): GPU tests, the bench line, the ncu launch list of a short bench run and --set full captures of the
# kernels that make up the step.  Usage: tools/gpu_r2prof.sh <tag>
tag=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
SHORT="python bench.py --steps 3 --warmup 3 --no-c4 --no-variants --no-cpu-baseline --no-parity"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $SHORT > gpurun_out/${tag}_ncu_launches.log 2>&1
for k in gs_persistent_kernel sor_persistent_kernel assemble_kernel face_prelude_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${tag}_$k $SHORT > gpurun_out/${tag}_ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out | tail -12
