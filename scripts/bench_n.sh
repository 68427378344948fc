#!/bin/bash
# bench_n.sh N [steps] -- the driver's N-GPU bench command, plus the halo check job with per-phase timing; logs under gpurun_out/
N=${1:-2}; K=${2:-10}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 bench.py --gpus $N --steps $K --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench rc=$?"
grep "^{" gpurun_out/bench_n$N.log | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print("N=%d value=%.2f it/s ms/step=%.2f e2e=%.2f us/pcg_it=%.1f frac=%.3f check=%s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["us_per_pcg_iteration"], d["roofline"]["frac"], json.dumps(d["check"])[:600]))
PY
mkdir -p gpurun_out/timing_n$N; rm -f gpurun_out/timing_n$N/*; timeout 400 env SPB_PCG_TIMING=1 SPB_PCG_TIMING_DIR=gpurun_out/timing_n$N $TR --master-port 29702 scripts/halo_check.py --g 4472 --iters 1 --no-reference > gpurun_out/halo_timing_n$N.log 2>&1; tail -n 1 gpurun_out/timing_n$N/pcg_timing_rank0.txt | cut -c1-700
grep "rank=0 iters" gpurun_out/halo_timing_n$N.log | tail -1 | cut -c1-700; grep "HALO_CHECK\|halo mode" gpurun_out/halo_timing_n$N.log
timeout 400 $TR --master-port 29703 scripts/halo_check.py --g 1024 > gpurun_out/halo_check_n$N.log 2>&1
grep "world=\|HALO_CHECK\|halo mode" gpurun_out/halo_check_n$N.log
