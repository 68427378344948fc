"""Reads bench.py JSON lines on stdin and prints one compact line each: value, ms/step, PCG iterations per step,
us per PCG iteration, roofline fraction and the per-stage times.  Usage: python bench.py ... | python scripts/fmt_bench.py LABEL"""
import sys,json
for line in sys.stdin:
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],3), d["config"].get("pcg_iterations_per_step"), round(r["us_per_pcg_iteration"],2), round(r["frac"],3), {k:round(v,3) for k,v in r["stage_ms_per_step"].items() if v})
