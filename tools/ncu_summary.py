"""Condense an `ncu --page raw --csv` dump (one row per profiled launch) into the few metrics the roofline
argument needs.  Usage: python tools/ncu_summary.py gpurun_out/x_raw.csv [--json out.json] > profiles/x.md"""
import csv
import json
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_uniform.sum", "smsp__cycles_active.avg",
]


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name"), "metrics": {}}
        for k in KEYS:
            if k in d and d[k] != "":
                rec["metrics"][k] = (d[k], u[k])
        stalls = {h: float(d[h]) for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and d[h]}
        if not stalls:
            stalls = {h: float(d[h].replace(",", "")) for h in hdr if "warp_issue_stalled" in h and "pct" in h and d[h]}
        rec["top_stalls"] = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        out.append(rec)
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
    for rec in out:
        print(f"### {rec['kernel']}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k, (v, u) in rec["metrics"].items():
            print(f"| `{k}` | {v} | {u} |")
        if rec["top_stalls"]:
            print("\nTop warp-stall reasons: " + ", ".join(f"`{k.split('issue_stalled_')[-1].split('_per')[0]}`={v:.2f}" for k, v in rec["top_stalls"]))
        print()


if __name__ == "__main__":
    main()
