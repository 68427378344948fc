#!/bin/bash
# halo_sweep.sh N [quick] -- correctness + timing of the subdomain (halo) mode on N GPUs, fused peer-window kernel vs NCCL path.
# Output goes to gpurun_out/halo_sweep_N.log
N=${1:-2}
OUT=gpurun_out/halo_sweep_${N}.log
mkdir -p gpurun_out
: > $OUT
run() { echo "### $*" >> $OUT; timeout 400 "$@" >> $OUT 2>&1; echo "### rc=$?" >> $OUT; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
run $TR scripts/halo_check.py --g 256
run $TR scripts/halo_check.py --g 1024 --iters 5
run env SPB_PCG_TIMING=1 $TR scripts/halo_check.py --g 1024 --iters 2 --no-reference
if [ "$2" != "quick" ]; then
run env SPB_HALO_PEER=0 $TR scripts/halo_check.py --g 1024 --iters 5 --no-reference
run env SPB_HALO_PEER=0 $TR scripts/halo_check.py --g 4472 --iters 3 --no-reference
fi
run $TR scripts/halo_check.py --g 4472 --iters 3 --no-reference
run env SPB_PCG_TIMING=1 $TR scripts/halo_check.py --g 4472 --iters 1 --no-reference
grep -v "^\[W\|^W0\|Warning\|warn\|^\*\*\*\|OMP_NUM_THREADS\|^$" $OUT | cut -c1-900 | tail -60
