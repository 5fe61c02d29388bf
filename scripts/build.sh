#!/bin/bash
# in-tree build of libmpc_b200.so for sm_100a (same flags as mpconstellation_b200/_lib.py build())
set -e
cd "$(dirname "$0")/../mpconstellation_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -shared -Xcompiler -fPIC \
     -o libmpc_b200.so mpc_b200.cu mpc_b200_drag.cu "$@" 2>&1 | grep -E "error|warning: |Compiling entry|spill|Used" | sed 's/ptxas info    : //' | paste - - - | sed 's/Compiling entry function//; s/for .sm_100a.//' | awk '{print}' 
