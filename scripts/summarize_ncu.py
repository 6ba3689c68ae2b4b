#!/usr/bin/env python
"""Summarise `ncu --set full` reports (raw page) into a small JSON + markdown table for profiles/.

    python scripts/summarize_ncu.py profiles/ncu_r1 gpurun_out/knn_full.ncu-rep gpurun_out/gemm_tc_full.ncu-rep
"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_inst_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "smsp__inst_executed.sum": "warp_insts",
    "sm__cycles_elapsed.max": "cycles",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "Grid Size": "grid", "Block Size": "block", "Kernel Name": "kernel",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h in KEYS:
                v = r[i]
                try:
                    v = float(v.replace(",", "")) * SCALE.get(units[i], 1.0)
                except ValueError:
                    pass
                d[KEYS[h]] = v
        res.append(d)
    return res


def main(prefix, reps):
    allk = {}
    for rep in reps:
        launches = load(rep)
        byname = {}
        for l in launches:
            name = l["kernel"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            byname.setdefault(name, []).append(l)
        for name, ls in byname.items():
            agg = {"launches_captured": len(ls), "report": rep.split("/")[-1]}
            for k in ls[0]:
                if isinstance(ls[0][k], float):
                    agg[k] = sum(x[k] for x in ls) / len(ls)
                else:
                    agg[k] = ls[0][k]
            agg["dram_traffic_bytes"] = agg.get("dram_read", 0.0) + agg.get("dram_write", 0.0)
            allk[name if name not in allk else "%s [%s]" % (name, agg["report"])] = agg
    json.dump(allk, open(prefix + ".json", "w"), indent=1, sort_keys=True)
    with open(prefix + ".md", "w") as f:
        f.write("# ncu --set full summaries (per launch averages; durations are under the profiler: cold, serialised)\n\n")
        f.write("| kernel | grid x block | ms | DRAM read MB | DRAM write MB | DRAM % | tensor pipe % | FMA pipe % | warps active % | regs |\n")
        f.write("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for name, a in allk.items():
            f.write("| `%s` | %s x %s | %.3f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d |\n" % (
                name, a.get("grid"), a.get("block"), a.get("duration", 0), a.get("dram_read", 0) / 1e6,
                a.get("dram_write", 0) / 1e6, a.get("dram_pct", 0), a.get("tensor_pipe_pct", 0), a.get("fma_pipe_pct", 0),
                a.get("warps_active_pct", 0), int(a.get("regs", 0))))
    print("wrote", prefix + ".json", prefix + ".md")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
