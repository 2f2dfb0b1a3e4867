#!/usr/bin/env python
"""Per-source-line view of an ncu --set full --import-source on report:  python profiles/make_line_table.py <rep> <tag>
writes profiles/<tag>_lines.md: warp instructions, shared-memory wavefronts and stall samples per CUDA source line and per
opcode, divided by the number of units (transitions / windows) of the launch."""
import collections
import csv
import io
import os
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
units = int(sys.argv[3]) if len(sys.argv) > 3 else 262144
min_inst = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0  # lines below this many warp instructions per unit are left out
HERE = os.path.dirname(os.path.abspath(__file__))
out = [f"# per-line instruction / shared-memory view ({tag}); report {os.path.basename(rep)}; all figures per unit "
       f"(= per transition / window, {units} per launch)\n"]
for kern, label in (("tqc_loss_group", "tqc_loss_group_kernel"), ("sample_gather_tile", "sample_gather_tile_kernel"),
                    ("sample_gather_lean", "sample_gather_lean_kernel"), ("fused_pass", "fused_pass_kernel")):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, fil, cur = None, None, None
    inst, wf, wfi, smp, src = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter(), {}
    ops, seen = collections.Counter(), set()
    for r in rows:
        if r and r[0] == "File Path":
            fil = os.path.basename(r[1]); continue
        if r and r[0] == "Line No":
            hdr = r
            iI, iW, iD, iS, iA = (hdr.index(x) for x in ("Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
                                                         "# Samples", "Address"))
            continue
        if not hdr or len(r) < len(hdr):
            continue
        if r[0] != "":
            cur = (fil, int(r[0])); src[cur] = r[1].strip(); continue
        if r[2] == "..." or r[iA] in seen:   # an instruction is listed under every line it is attributed to: count it once
            continue
        seen.add(r[iA])
        try:
            n = int(r[iI])
        except ValueError:
            continue
        inst[cur] += n; wf[cur] += int(r[iW]); wfi[cur] += int(r[iD]); smp[cur] += int(r[iS])
        t = r[3].split()
        ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += n
    if not inst:
        continue
    ti, tw, twi, ts = sum(inst.values()), sum(wf.values()), sum(wfi.values()), max(sum(smp.values()), 1)
    out.append(f"\n## {label}\n\n{ti / units:.1f} warp instructions, {tw / units:.1f} shared-memory wavefronts "
               f"({twi / units:.1f} ideal) per unit from instructions (bulk-copy engine writes come on top)\n")
    out.append("opcodes: " + ", ".join(f"{k} {v / units:.1f}" for k, v in ops.most_common(16)) + "\n")
    out.append("| file:line | warp inst | smem wavefronts (ideal) | stall samples | source |\n|---|---|---|---|---|")
    for k in sorted(inst, key=lambda k: -inst[k]):
        if inst[k] / units < min_inst and wf[k] / units < 1.0 and smp[k] / ts < 0.01:
            continue
        out.append(f"| {k[0]}:{k[1]} | {inst[k] / units:.1f} | {wf[k] / units:.1f} ({wfi[k] / units:.1f}) | {100 * smp[k] / ts:.1f} % | "
                   f"`{src[k][:90].replace('|', '/')}` |")
open(os.path.join(HERE, f"{tag}_lines.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))
