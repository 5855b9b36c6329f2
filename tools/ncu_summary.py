#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU):  python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.json]
One record per profiled launch with the metrics the roofline discussion needs."""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sectors.sum": "l2_sectors_M",
    "lts__t_sectors_op_read.sum": "l2_read_sectors_M",
    "lts__t_sectors_op_write.sum": "l2_write_sectors_M",
    "lts__t_sectors_op_atom.sum": "l2_atom_sectors_M",
    "lts__t_sectors_op_red.sum": "l2_red_sectors_M",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_wavefront_pct",
    "l1tex__data_pipe_lsu_wavefronts.sum": "l1_wavefronts_M",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__inst_executed.sum": "warp_insts_M",
    "sm__inst_executed_pipe_fma.sum": "fma_pipe_insts_M",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio": "stall_long_scoreboard",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio": "stall_short_scoreboard",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio": "stall_lg_throttle",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio": "stall_mio_throttle",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio": "stall_barrier",
    "smsp__average_warp_latency_issue_stalled_wait.ratio": "stall_wait",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio": "stall_math_throttle",
}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"kernel": d.get("Kernel Name", "")[:100], "id": d.get("ID")}
        for k, name in KEYS.items():
            if k in d and d[k] not in ("", "n/a"):
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                u = units[hdr.index(k)]
                if name.endswith("_MB"):
                    v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                elif name.endswith("_M"):
                    v = v * 1e-6
                elif name == "time_us":
                    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
                rec[name] = round(v, 3)
        out.append(rec)
    text = "\n".join(json.dumps(r) for r in out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
