"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON: duration, DRAM traffic, pipe utilisation,
registers, top stall reasons.  Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.json"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_mufu_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "sm__cycles_elapsed.avg": "sm_cycles",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
}


def main(rep, out):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for d in data:
        e = {"kernel": d[col["Kernel Name"]][:80]}
        for k, name in KEYS.items():
            if k in col:
                e[name] = f"{d[col[k]]} {units[col[k]]}".strip()
        stalls = {}
        for h, i in col.items():
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    v = float(d[i].replace(",", ""))
                except ValueError:
                    continue
                if v > 0:
                    stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = v
        tot = sum(stalls.values()) or 1.0
        e["stall_samples_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
        launches.append(e)
    json.dump({"report": rep, "launches": launches}, open(out, "w"), indent=1)
    print(json.dumps(launches[-1], indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
