#!/bin/bash
# ncu captures of the persistent PCG kernels on LF-1M (one launch each, --set full), to gpurun_out/.
# usage: bash scripts/ncu_pcg.sh   (on the GPU box; run AFTER the same bench command has exited 0 without ncu)
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-direct --no-extras > gpurun_out/ncu_pre_classic.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:pcg_persistent_kernel -s 3 -c 1 -o gpurun_out/r2_pcg_classic -f \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-direct --no-extras > gpurun_out/ncu_classic.log 2>&1
SPB_PCG_VARIANT=sr python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-direct --no-extras > gpurun_out/ncu_pre_sr.log 2>&1 || exit 1
SPB_PCG_VARIANT=sr ncu --set full --clock-control none --import-source on -k regex:pcg_sr_kernel -s 3 -c 1 -o gpurun_out/r2_pcg_sr -f \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-direct --no-extras > gpurun_out/ncu_sr.log 2>&1
ls -la gpurun_out/*.ncu-rep
