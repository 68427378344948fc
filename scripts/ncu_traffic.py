"""Reads an `ncu --set full` report (.ncu-rep) of one bench run and writes the per-launch DRAM traffic of a kernel as a
small JSON under profiles/ (bench.py's roofline.traffic reads the newest profiles/*_pcg_traffic.json).
    python scripts/ncu_traffic.py gpurun_out/r2_pcg.ncu-rep pcg_sr_kernel profiles/r2_pcg_traffic.json [--iters-per-launch 91.6]"""
import csv
import io
import json
import subprocess
import sys

rep, kernel, out = sys.argv[1], sys.argv[2], sys.argv[3]
iters = None
if "--iters-per-launch" in sys.argv:
    iters = float(sys.argv[sys.argv.index("--iters-per-launch") + 1])
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]
col = {name: i for i, name in enumerate(hdr)}
want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
units = rows[1]
launches = []
for r in rows[2:]:
    if len(r) != len(hdr) or kernel not in r[col["Kernel Name"]]:
        continue
    d = {"kernel": r[col["Kernel Name"]][:80]}
    for w in want:
        if w in col:
            try:
                d[w] = float(r[col[w]].replace(",", ""))
                d[w + ".unit"] = units[col[w]]
            except ValueError:
                pass
    launches.append(d)
if not launches:
    raise SystemExit("no launch of %s in %s" % (kernel, rep))


def to_bytes(v, unit):
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)


tot = [to_bytes(d["dram__bytes_read.sum"], d["dram__bytes_read.sum.unit"]) + to_bytes(d["dram__bytes_write.sum"], d["dram__bytes_write.sum.unit"])
       for d in launches]
res = {"report": rep, "kernel": kernel, "launches": len(launches), "dram_bytes_per_launch": sum(tot) / len(tot), "per_launch": launches}
if iters:
    res["pcg_iterations_per_launch"] = iters
    res["dram_bytes_per_pcg_iteration"] = res["dram_bytes_per_launch"] / iters
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "per_launch"}))
