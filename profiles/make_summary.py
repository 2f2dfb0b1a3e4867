#!/usr/bin/env python
"""Turn ncu reports (--set full) into the tracked summary files:  python profiles/make_summary.py gpurun_out/prof_r1.ncu-rep r1 [more reports]
writes profiles/<tag>_ncu_summary.md and profiles/traffic.json (dram bytes per launch per kernel, read by bench.py)."""
import csv
import io
import json
import os
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
HERE = os.path.dirname(os.path.abspath(__file__))
reports = []
for rp in [rep] + sys.argv[3:]:
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    reports.append((rows[0], rows[1], rows[2:]))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "sm__icc_requests.sum.pct_of_peak_sustained_elapsed",
        "sm__icc_request_hit_rate.pct", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "second": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}
out, traffic = [f"# ncu --set full summary ({tag}); source reports: {', '.join(os.path.basename(x) for x in [rep] + sys.argv[3:])} "
                f"(scratch, not tracked)\n"], {}
for hdr, units, data in reports:
  ik = hdr.index("Kernel Name")
  for r in data:
      name = r[ik].split("(")[0].replace("void ", "").strip()
      out.append(f"\n## {r[ik][:110]}\n\n| metric | value | unit |\n|---|---|---|")
      vals = {}
      for m in want:
          if m in hdr:
              i = hdr.index(m)
              out.append(f"| {m} | {r[i]} | {units[i]} |")
              try:
                  vals[m] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
              except ValueError:
                  pass
      if "dram__bytes_read.sum" in vals:
          tot = vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
          key = ("fused_pass_kernel" if "fused_pass" in name else "sample_gather_kernel" if "sample_gather" in name else
                 "tqc_loss_kernel" if "tqc_loss" in name else name.split("<")[0])
          traffic[key] = tot
          if "smsp__inst_executed.sum" in vals:
              traffic.setdefault("warp_instructions", {})[key] = vals["smsp__inst_executed.sum"]
          dur = vals.get("gpu__time_duration.sum")
          out.append(f"| dram bytes read+write per launch | {tot:.4g} | byte |")
          if dur:
              out.append(f"| dram GB/s under ncu (cold-cache replay) | {tot / dur / 1e9:.1f} | GB/s |")
open(os.path.join(HERE, f"{tag}_ncu_summary.md"), "w").write("\n".join(out) + "\n")
json.dump(traffic, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
print("\n".join(out))
