"""Key metrics of every launch in an `ncu --set full` report, as text (for profiles/).
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt
Also: python scripts/ncu_summary.py --launches gpurun_out/r2_launches.csv   (per-kernel totals of a gpu__time_duration launch list)"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"]

if sys.argv[1] == "--launches":
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    col = {n: i for i, n in enumerate(rows[hdr])}
    tot = {}
    for r in rows[hdr + 1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[col["Kernel Name"]].split("(")[0][:70]
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += v
    allus = sum(t[1] for t in tot.values())
    print("launch list %s (ncu --metrics gpu__time_duration.sum; cold-cache, serialised: shares, not absolutes)" % sys.argv[2])
    for name, (cnt, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-72s launches %5d  total %10.1f us  share %5.1f %%  mean %9.1f us" % (name, cnt, us, 100 * us / allus, us / cnt))
    sys.exit(0)

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
print("report %s (ncu --set full --clock-control none)" % rep)
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    print("kernel: %s" % r[col["Kernel Name"]][:160])
    for w in WANT:
        if w in col and r[col[w]] != "":
            print("  %-85s %s %s" % (w, r[col[w]], units[col[w]]))
    stalls = [(h, float(r[col[h]])) for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled") and h.endswith(".ratio") is False and r[col[h]] not in ("", "n/a")]
    st = [(h, float(r[col[h]] or 0)) for h in hdr if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct")]
    for h, v in sorted(st, key=lambda kv: -kv[1])[:6]:
        print("  %-85s %.2f" % (h, v))
