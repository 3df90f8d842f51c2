#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/dbg_build.log 2>&1
for W in 4 3; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 2961$W tests/mp_worker.py gpurun_out/slide_w$W.npz alpine70 2 0 2 > gpurun_out/dbg_w$W.log 2>&1; echo "w$W rc=$?"
done
PBSM3D_HALO=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29617 tests/mp_worker.py gpurun_out/slide_w4_nccl.npz alpine70 2 0 2 > gpurun_out/dbg_w4n.log 2>&1; echo "w4 nccl rc=$?"
