#!/bin/bash
# ncu capture of the persistent PCG kernel on LF-1M as the bench runs it (one launch, --set full), to gpurun_out/.
# usage: bash scripts/ncu_pcg.sh   (on the GPU box; the same bench command runs first WITHOUT ncu and must exit 0)
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-direct --no-extras"
$B > gpurun_out/ncu_pre_classic.log 2>&1 || exit 1
# launch 4 of the kernel = first loop body of the timed solve (81 PCG iterations; the solve's iteration counts are printed by SPB_PCG_TIMING=1)
ncu --set full --clock-control none --import-source on -k regex:pcg_stream_kernel -s 3 -c 1 -o gpurun_out/r2_pcg_stream_relaxed -f $B > gpurun_out/ncu_classic.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv
ncu --set full --clock-control none --import-source on -k regex:assemble_pairs -s 3 -c 1 -o gpurun_out/r2_asm_pairs -f $B > gpurun_out/ncu_asm_pairs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_cols -s 3 -c 1 -o gpurun_out/r2_asm_cols -f $B > gpurun_out/ncu_asm_cols.log 2>&1
ls -la gpurun_out/r2_asm_*.ncu-rep
