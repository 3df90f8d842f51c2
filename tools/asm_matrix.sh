#!/bin/bash
# assembly kernel variants on config c2: prints one line per variant
mkdir -p gpurun_out
tag=$1
run() { env "$@" python tools/time_kernels.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', 'asm ms', round(d['assembly']['ms'],4), 'step asm', round(d['step']['ms_assembly'],3))"; }
{
run PBSM3D_ASSEMBLY=tile PBSM3D_ASM_MINB=2
run PBSM3D_ASSEMBLY=tile PBSM3D_ASM_MINB=3
run PBSM3D_ASSEMBLY=column PBSM3D_ASM_MINB=3
run PBSM3D_ASSEMBLY=column PBSM3D_ASM_MINB=4
run PBSM3D_ASSEMBLY=column PBSM3D_ASM_MINB=5
run PBSM3D_ASSEMBLY=column PBSM3D_ASM_MINB=6
run PBSM3D_ASSEMBLY=column PBSM3D_ASM_MINB=2 PBSM3D_ASM_UNROLL=2
run PBSM3D_ASSEMBLY=column PBSM3D_ASM_MINB=3 PBSM3D_ASM_UNROLL=2
run PBSM3D_ASSEMBLY=column PBSM3D_ASM_MINB=4 PBSM3D_ASM_UNROLL=2
} | tee gpurun_out/${tag}_asm_matrix.txt
