"""Summarise an .ncu-rep (raw + source pages) into a short text for profiles/ (run on the CPU box)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
m = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum", "local_load", "smsp__inst_executed_op_local"]
units = dict(zip(hdr, rows[1])) if len(rows) > 2 else {}
print("kernel:", m.get("Kernel Name"))
for k in keys:
    for h in hdr:
        if h == k or (k in h and len(k) < 25):
            print(f"{h} = {m[h]} {units.get(h,'')}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot_inst = sum(int(r[ix['Instructions Executed']]) for r in data)
tot_thr = sum(int(r[ix['Thread Instructions Executed']]) for r in data)
print(f"\nSASS instructions: {len(data)}; warp-instructions executed: {tot_inst}; avg active threads/instr: {tot_thr/tot_inst:.2f}")
ops = {}
for r in data:
    op = r[1].split()[0] if not r[1].strip().startswith('@') else r[1].split()[1]
    op = op.split('.')[0]
    ops[op] = ops.get(op, 0) + int(r[ix['Instructions Executed']])
print("opcode mix (share of executed warp-instructions):")
for op, c in sorted(ops.items(), key=lambda kv: -kv[1])[:22]:
    print(f"  {op:10s} {100*c/tot_inst:6.2f}%")
seg = []; cur = None
for k, r in enumerate(data):
    ie = int(r[ix['Instructions Executed']]); te = int(r[ix['Thread Instructions Executed']])
    if ie == 0: continue
    a = te / ie
    if cur and abs(cur['ie'] - ie) / max(ie, 1) < 0.02 and abs(cur['a'] - a) < 0.7:
        cur['n'] += 1; cur['sum_ie'] += ie; cur['sum_te'] += te; cur['samples'] += int(r[ix['# Samples']])
    else:
        cur = dict(start=k, n=1, ie=ie, a=a, sum_ie=ie, sum_te=te, samples=int(r[ix['# Samples']]), first=r[1].strip()[:44]); seg.append(cur)
ts = sum(s['samples'] for s in seg)
print("\ncode segments (runs of SASS with equal execution count):")
print("  sass#  len  execs/instr  avg_threads  %instr  %stall_samples  first instruction")
for s in seg:
    if s['sum_ie'] / tot_inst > 0.004:
        print(f"  {s['start']:5d} {s['n']:4d} {s['ie']:11d} {s['sum_te']/s['sum_ie']:8.1f} {100*s['sum_ie']/tot_inst:8.2f} {100*s['samples']/ts:10.2f}      {s['first']}")
