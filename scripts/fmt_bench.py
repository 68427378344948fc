import sys,json
for line in sys.stdin:
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],3), d["config"].get("pcg_iterations_per_step"), round(r["us_per_pcg_iteration"],2), round(r["frac"],3), {k:round(v,3) for k,v in r["stage_ms_per_step"].items() if v})
