#!/bin/bash
# compile the single-warp marching kernel alone and print the static opcode histogram of its row loop
# usage: scripts/m2_sass.sh [C] [S] [extra nvcc flags]
C=${1:-1}; S=${2:-2}; shift; shift
OUT=/tmp/exp/m2_${C}${S}
mkdir -p /tmp/exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DCC=$C -DSS=$S "$@" -Xptxas -v -cubin -o $OUT.cubin /root/repo/scripts/exp/m2only.cu 2>&1 | grep -E "registers|spill|error" 
cuobjdump -sass $OUT.cubin > $OUT.sass
nvdisasm -g -c $OUT.cubin > $OUT.dis 2>/dev/null
python3 /root/repo/scripts/m2_rowloop.py $OUT.sass $OUT.dis
