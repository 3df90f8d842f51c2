#!/bin/bash
tag=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -14 gpurun_out/${tag}_pytest.log
timeout 300 python tools/time_slide.py 708 0 1.0 > gpurun_out/${tag}_slide.json 2> gpurun_out/${tag}_slide.err; cat gpurun_out/${tag}_slide.json
timeout 300 python tools/time_slide.py 708 0 3.0 > gpurun_out/${tag}_slide_deep.json 2>> gpurun_out/${tag}_slide.err; cat gpurun_out/${tag}_slide_deep.json
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['next_rows'])"
