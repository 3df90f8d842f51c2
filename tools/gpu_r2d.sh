#!/bin/bash
# 2-GPU call: full GPU suite (multi-GPU tests included), snow_slide timing, steady-state launch list.
tag=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python tools/time_slide.py 708 354 1.0 > gpurun_out/${tag}_slide.json 2> gpurun_out/${tag}_slide.err; echo "slide rc=$?"; cat gpurun_out/${tag}_slide.json
timeout 300 python tools/time_slide.py 708 0 3.0 > gpurun_out/${tag}_slide_deep.json 2>> gpurun_out/${tag}_slide.err; cat gpurun_out/${tag}_slide_deep.json
SHORT="python bench.py --steps 3 --warmup 3 --no-c4 --no-variants --no-cpu-baseline --no-parity"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 120 --csv --log-file gpurun_out/${tag}_launches.csv $SHORT > gpurun_out/${tag}_ncu_launches.log 2>&1
tail -3 gpurun_out/${tag}_slide.err
