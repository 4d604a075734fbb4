"""Extract the counters quoted in DESIGN.md from an ncu report (read on the CPU box):
   python profiles/ncu_summary.py gpurun_out/X.ncu-rep profiles/X_ncu_metrics.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active", "sm__pipe_tensor_subpipe_hmma", "sm__inst_executed_pipe_xu",
        "sm__inst_executed_pipe_tensor", "smsp__issue_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block", "sm__cycles_elapsed.max", "sm__mem_tensor_cycles_active",
        "sm__throughput.avg.pct", "lts__t_bytes.sum", "lts__throughput", "sm__warps_active.avg.pct",
        "smsp__average_warp_latency_issue_stalled", "smsp__warp_issue_stalled", "l1tex__throughput", "sm__inst_executed_pipe_lsu",
        "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_uniform", "smsp__cycles_active.avg")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
        for i, h in enumerate(hdr):
            if any(k in h for k in KEEP) and r[i] not in ("", "n/a"):
                d[h + (" [%s]" % units[i] if units[i] else "")] = r[i]
        res.append(d)
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", out, "with", len(res), "launch(es),", len(res[0]) if res else 0, "counters")


if __name__ == "__main__":
    main()
