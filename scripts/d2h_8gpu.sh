#!/bin/bash
# concurrent D2H from all GPUs of the box: is the 8-rank e2e leg limited by the platform (PCIe switches / host memory)?
nvidia-smi topo -m 2>/dev/null | head -14
nproc; lscpu | grep -i "numa\|socket\|model name" | head -8
N=$(nvidia-smi -L | wc -l)
echo "== one GPU alone"
CUDA_VISIBLE_DEVICES=0 scripts/micro/d2h_probe | head -2
echo "== $N GPUs at once (first two lines of each)"
for i in $(seq 0 $((N-1))); do CUDA_VISIBLE_DEVICES=$i scripts/micro/d2h_probe > /tmp/d2h_$i.log 2>&1 & done; wait
for i in $(seq 0 $((N-1))); do echo "gpu $i: $(sed -n 1p /tmp/d2h_$i.log)"; done
echo "== 4 GPUs at once (0,2,4,6)"
for i in 0 2 4 6; do CUDA_VISIBLE_DEVICES=$i scripts/micro/d2h_probe > /tmp/d2h_$i.log 2>&1 & done; wait
for i in 0 2 4 6; do echo "gpu $i: $(sed -n 1p /tmp/d2h_$i.log)"; done
echo "== 2 GPUs at once (0,1)"
for i in 0 1; do CUDA_VISIBLE_DEVICES=$i scripts/micro/d2h_probe > /tmp/d2h_$i.log 2>&1 & done; wait
for i in 0 1; do echo "gpu $i: $(sed -n 1p /tmp/d2h_$i.log)"; done
