"""Turn an `ncu --set full --import-source on` report into the text summary kept under profiles/.

  python tools/ncu_summary.py gpurun_out/k3_r2k.ncu-rep profiles/k3_r2k_summary.txt [kernel-name-substring ...]

Per selected launch (the longest launch of every kernel whose name contains one of the substrings; all kernels if none
given): duration, grid, registers, issue-slot and FP64-pipe utilisation, DRAM bytes, L2 hit rate, the warp-stall shares
(smsp__average_warps_issue_stalled_*_per_issue_active), the opcode mix of the executed warp instructions and, from the
SASS page, every loop of more than 64 instructions with its share of the stall samples (the hot loops of K3 are the
per-knot loops of linearisation, Riccati step and rollout).
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, out_path = sys.argv[1], sys.argv[2]
subs = sys.argv[3:]


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, rows = raw[0], raw[1], raw[2:]
col = {h: i for i, h in enumerate(hdr)}
best = {}
for i, r in enumerate(rows):
    name = r[col["Kernel Name"]].split("(")[0]
    if subs and not any(s in name for s in subs):
        continue
    try:
        dur = float(r[col["gpu__time_duration.sum"]])
    except ValueError:
        continue
    u = units[col["gpu__time_duration.sum"]]
    dur_ms = dur * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(u, 1.0)
    if name not in best or dur_ms > best[name][1]:
        best[name] = (i, dur_ms, r[col["ID"]])

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum"]
lines = ["summary of %s (tools/ncu_summary.py)" % rep, ""]
SRC = None
for name, (i, dur_ms, kid) in best.items():
    r = rows[i]
    lines.append("=" * 100)
    lines.append("%s   launch id %s   duration %.3f ms" % (name, kid, dur_ms))
    for k in KEYS:
        if k in col and r[col[k]] not in ("", "nan", "-nan"):
            lines.append("  %-70s %s %s" % (k, r[col[k]], units[col[k]]))
    st = {}
    for h, j in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[j])
            except ValueError:
                pass
    tot = sum(st.values()) or 1.0
    lines.append("  warp states per issue (share of all warp-cycles): " + ", ".join(
        "%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda x: -x[1]) if v / tot > 0.005))
    # ---- SASS page of this launch (the export holds one section per launch, each opened by a "Kernel Name" row)
    if SRC is None:
        SRC = []
        for row in csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))):
            if row and row[0] == "Kernel Name":
                SRC.append([])
            elif SRC:
                SRC[-1].append(row)
    h2 = None
    ins = []
    per = max(1, len(SRC) // max(1, len(rows)))   # the export repeats every launch `per` times
    for row in (SRC[per * i] if per * i < len(SRC) else []):
        if row and row[0] == "Address":
            h2 = {h: j for j, h in enumerate(row)}
            continue
        if h2 and len(row) > h2["Instructions Executed"]:
            try:
                ins.append((int(row[h2["Address"]], 16), row[h2["Source"]].strip(), int(row[h2["# Samples"]] or 0), int(row[h2["Instructions Executed"]] or 0)))
            except ValueError:
                pass
    if not ins:
        continue
    base = ins[0][0]
    ins = [(a - base, t, s, e) for a, t, s, e in ins]
    tot_s = sum(x[2] for x in ins) or 1
    tot_e = sum(x[3] for x in ins) or 1
    mix = collections.Counter()
    for a, t, s, e in ins:
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
        mix[op] += e
    lines.append("  executed warp instructions by opcode: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot_e) for k, v in mix.most_common(14)))
    lines.append("  loops (> 64 instructions) by share of stall samples:")
    for a, t, s, e in ins:
        if "BRA" in t:
            m = re.search(r"(0x[0-9a-f]+)\s*$", t)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            tgt = tgt - base if tgt >= base else tgt
            if tgt < a and 64 < (a - tgt) // 16 < 3000:
                seg = [x for x in ins if tgt <= x[0] <= a]
                ss = sum(x[2] for x in seg)
                if ss / tot_s < 0.01:
                    continue
                first = seg[0][3]
                m2 = collections.Counter()
                for x in seg:
                    m2[re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0]] += 1
                lines.append("    0x%05x-0x%05x  %4d instr  %5.1f%% of samples  trip count %d  static mix: %s" % (
                    tgt, a, (a - tgt) // 16 + 1, 100 * ss / tot_s, first, ", ".join("%s %d" % kv for kv in m2.most_common(7))))
open(out_path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
