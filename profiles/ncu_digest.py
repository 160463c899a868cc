"""Digest an ncu report: key raw metrics per kernel + SASS opcode / stall summary of one kernel (run in the build container).
usage: python profiles/ncu_digest.py <report.ncu-rep> [kernel-regex]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "sm__cycles_elapsed.avg", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k} [{units[i]}]:", [d[i][:60] for d in data])
if kre:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    ix = {n: i for i, n in enumerate(h)}
    d = rows[2:]
    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    ti = sum(f(r, "Instructions Executed") for r in d)
    ops, opw, ops_smp = collections.Counter(), collections.Counter(), collections.Counter()
    for r in d:
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        op = ".".join((m.group(2) if m else "?").split(".")[:2])
        ops[op] += f(r, "Instructions Executed")
        opw[op] += f(r, "L1 Wavefronts Shared")
        ops_smp[op] += f(r, "# Samples")
    ts = sum(ops_smp.values())
    print(f"-- SASS mix of {kre}: {len(d)} lines, {ti:.0f} warp instructions")
    for op, c in ops.most_common(16):
        print(f"  {op:22s} {c / ti * 100:6.2f}% inst  smem wavefronts {opw[op]:12.0f}  stall samples {ops_smp[op] / ts * 100:5.1f}%")
    st = collections.Counter()
    for k in h:
        if k.startswith("stall_") and "Not Issued" not in k:
            st[k] = sum(f(r, k) for r in d)
    t = sum(st.values())
    print("-- stall reasons:", ", ".join(f"{k[6:]} {c / t * 100:.1f}%" for k, c in st.most_common(9)))
