#!/bin/bash
# One gpurun call: GPU tests + kernel timings + a bench line.  Usage: tools/gpu_round.sh <tag> [pytest args...]
tag=$1; shift
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
