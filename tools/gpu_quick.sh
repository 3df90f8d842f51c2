#!/bin/bash
tag=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_slide.py tests/test_gpu_snobal.py tests/test_adaptor_cpp.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python tools/time_slide.py 708 0 1.0 > gpurun_out/${tag}_slide.json 2> gpurun_out/${tag}_slide.err; cat gpurun_out/${tag}_slide.json
timeout 300 python tools/time_slide.py 708 0 3.0 > gpurun_out/${tag}_slide_deep.json 2>> gpurun_out/${tag}_slide.err; cat gpurun_out/${tag}_slide_deep.json
tail -3 gpurun_out/${tag}_slide.err
