#!/usr/bin/env python
"""bench.py -- stiff-ODE ensemble systems solved/sec (Robertson, RADAU_IIA_53) on N B200.

Workload (BASELINE.json configs[1]): 2^20 swept-rate Robertson members PER GPU, RADAU_IIA_53,
rtol = atol = 1e-6, newton tol 1e-7, dx_delta 1e-4, maxit 5000, refresh_jac 10, dt0 = 1e-6,
y0 = {1,0,0}, t in [0,40] (--span long: [0,1e11]).  One "step" = one solve of the whole ensemble.

  value  = members/s with inputs resident in HBM (device-resident C-ABI entry, CUDA-event timed)
  e2e    = members/s through rehuel_irk_ensemble() with HOST buffers (pinned), H2D + D2H inside
  roofline = FP64: flops from the per-member attempt / newton_iters counters the kernel returns
             (SURVEY.md 8d dense model, plus the executed-algorithm model) / kernel time,
             against the FP64 FMA peak measured on this device in this run
  cpu_baseline = the unmodified reference (oracle/_ref) on the host cores, one process per core

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--span short|long]
N > 1: launched by torchrun, one rank per GPU; members are sharded by index (rank r owns members
[r*M, (r+1)*M) of the sweep: weak scaling), the only collective is an all-reduce of statistics.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "stiff-ODE ensemble systems solved/sec (Robertson, RADAU_IIA_53)"
UNIT = "systems/s"
M_PER_GPU = 1 << 20
RTOL = ATOL = 1e-6
DT0 = 1e-6
SPANS = {"short": (0.0, 40.0), "long": (0.0, 1e11)}
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2: SMs x FP64 lanes x 2 x max SM clock


def workload_name(span: str, M: int) -> str:
    t0, t1 = SPANS[span]
    return (f"robertson-sweep M={M}/GPU (k1,k2,k3 = nominal*10^(u-1/2), splitmix64 seed 0x5EED), RADAU_IIA_53, "
            f"rtol=atol=1e-6, t=[{t0:g},{t1:g}], dt0=1e-6, y0={{1,0,0}}")


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores, one process per core
# ---------------------------------------------------------------------------------------------
_WORKER = r"""
import sys, json, time
sys.path.insert(0, {root!r})
import numpy as np
from oracle.refsolver import RefSolver, make_options
from rehuel_b200 import ensembles
kind, start, stride, total, t1 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
s = RefSolver(kind)
P = ensembles.robertson_params(total)
o = make_options(rel_tol=1e-6, abs_tol=1e-6)
tic = time.perf_counter()
r = s.solve("robertson", 211, 0.0, t1, 1e-6, ensembles.ROBERTSON_Y0, P, o, start=start, stride=stride)
el = time.perf_counter() - tic
n = len(range(start, total, stride))
ok = int((r.status[start::stride] == 0).sum())
print(json.dumps(dict(n=n, ok=ok, seconds=el, attempts=int(r.counters["attempt"][start::stride].sum()))))
"""


def cpu_reference_run(span: str, sample: int, cores: int) -> dict:
    """Members 0..sample-1 of the sweep striped over `cores` processes; wall = slowest process."""
    kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "librehuel_ref.so")) else "port"
    if kind == "port" and not os.path.exists(os.path.join(ROOT, "oracle", "libirk_port.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)
    t1 = SPANS[span][1]
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as fh:
        fh.write(_WORKER.format(root=ROOT))
        script = fh.name
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    tic = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, script, kind, str(c), str(cores), str(sample), repr(t1)],
                              stdout=subprocess.PIPE, env=env, text=True) for c in range(cores)]
    outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
    wall = time.perf_counter() - tic
    os.unlink(script)
    slowest = max(o["seconds"] for o in outs)
    n = sum(o["n"] for o in outs)
    assert n == sample and all(o["ok"] == o["n"] for o in outs)
    return dict(value=n / slowest, unit=UNIT, cores=cores, kind=kind,
                sample=f"members 0..{sample - 1} of the same sweep, striped over {cores} processes "
                       f"(one per core); time = slowest process ({slowest:.2f} s; wall incl. start-up {wall:.2f} s)",
                per_core=n / slowest / cores, seconds=slowest, attempts=sum(o["attempts"] for o in outs))


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


_JSON_OUT = None


def claim_stdout() -> None:
    """stdout carries the ONE JSON line and nothing else: file descriptor 1 is pointed at stderr for the rest of
    the process (NCCL writes its version banner to fd 1 from C, below Python's sys.stdout) and the line goes to a
    duplicate of the original descriptor."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_step = (2048 if args.span == "short" else 32) * cores
    for _ in range(args.warmup):
        cpu_reference_run(args.span, max(cores, per_step // 8), cores)
    vals, secs = [], 0.0
    for _ in range(args.steps):
        r = cpu_reference_run(args.span, per_step, cores)
        vals.append(r["value"])
        secs += r["seconds"]
    value = per_step * args.steps / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.span, M_PER_GPU), "sample_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": r["kind"],
                         "sample": f"each step = members 0..{per_step - 1} of the sweep, one process per core, time = slowest process"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.fh = open(self.path, "w")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        self.fh.close()
        sm, mx, power, reasons = [], [], [], set()
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [c for c, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Pin this rank (and the pinned buffers it allocates afterwards: first touch) to the NUMA node
    its GPU hangs off; with several ranks per box the host<->device copies of the e2e path otherwise
    cross sockets.  Best effort: returns a note for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        bdf = f"{dom:04x}:{bus:02x}:{dev:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"numa: gpu {bdf} -> node {node}, {len(allowed)} cpus"
        return f"numa: node {node} has no allowed cpus"
    except Exception as e:  # noqa: BLE001
        return f"numa: not bound ({type(e).__name__})"


# dram__bytes_read.sum + dram__bytes_write.sum of ensemble_irk_thread<RobertsonF,...> at 2^20 members, t in [0,40]
# (profiles/r1_v8_ncu_summary.txt: 248.26 MB + 290.20 MB)
NCU_DRAM_BYTES_PER_LAUNCH = 538.46e6


def run_gpu_arm(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    from rehuel_b200 import capi, ensembles, sharding
    from rehuel_b200.device import DeviceEnsemble

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    numa_note = bind_to_gpu_numa_node(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"

    M = args.members
    t0, t1 = SPANS[args.span]
    handout = {"index": capi.ORDER_INDEX, "locality": capi.ORDER_LOCALITY, "locality_desc": capi.ORDER_LOCALITY_DESC}[args.handout]
    opts = capi.make_options(RTOL, ATOL, handout_order=handout)
    # rank r owns members [r*M, (r+1)*M) of the sweep (weak scaling; any rank can generate its slice)
    P_host = torch.from_numpy(ensembles.robertson_params(M, start=rank * M)).pin_memory()
    y0_host = torch.from_numpy(ensembles.ROBERTSON_Y0.copy()).pin_memory()
    P_dev = P_host.to(dev)
    y0_dev = y0_host.to(dev)
    ens = DeviceEnsemble(capi.ROBERTSON, capi.RADAU_IIA_53, M, y0_dev, P_dev, handout_order=handout)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_tflops = capi.fp64_peak_tflops()

    # ---------------- value: inputs resident in HBM, device-timed
    for _ in range(args.warmup):
        ens.solve(t0, t1, DT0, opts)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = capi.kernel_launches()
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for a, m, b in ev:
        flush.fill_(1)                 # L2 flush between timed iterations (outside the event pair)
        a.record()
        ens.reorder()                  # hand-out order from the parameters (2 kernels + radix sort); part of the step
        m.record()
        ens.launch(t0, t1, DT0, opts)  # ONE launch of the integration kernel
        b.record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = capi.kernel_launches() - launches0
    step_ms = [a.elapsed_time(b) for a, m, b in ev]
    kernel_ms = [m.elapsed_time(b) for a, m, b in ev]   # the integration kernel alone (roofline)
    total_ms = float(sum(step_ms))
    tt = torch.tensor([total_ms, float(sum(kernel_ms))], dtype=torch.float64, device=dev)
    ll = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ll, op=dist.ReduceOp.SUM)
    total_ms_max, kernel_ms_max = tt[0].item(), tt[1].item()

    # ---------------- statistics from the kernel's own counters (the one collective of the path)
    att = ens.counter("attempt").double()
    its = ens.counter("newton_iters").double()
    fun = ens.counter("fun_evals").double()
    nok = ens.counter("newton_success").double()
    alt_uses = fun - 3.0 * (nok + its) - nok  # attempts that also evaluated formula 8.20
    fixed, extra820, per_newton = ensembles.flops_per_attempt_dense_model(3, 3, 11, 8)
    flops_dense = fixed * att.sum() + per_newton * its.sum() + extra820 * alt_uses.sum()
    # DESIGN.md section 6; the last term: final residual evaluations the kernel skips as dead work (measured share of
    # the successful Newton solves whose step test fires first: REHUEL_COUNT_SKIPS build, 0.962 short span / 0.999 long)
    skip_share = 0.962 if args.span == "short" else 0.999
    flops_exec = 259.0 * att.sum() + 369.0 * its.sum() + 30.0 * alt_uses.sum() - 132.0 * skip_share * nok.sum()
    local = sharding.pack_stats(members=M, attempt=att, steps=ens.counter("steps").double(), newton_iters=its,
                                fun_evals=fun, jac_evals=ens.counter("jac_evals").double(),
                                reject_err=ens.counter("reject_err").double(),
                                reject_newton=ens.counter("reject_newton").double(),
                                n_failed=(ens.status != 0).double(), flops_dense_model=flops_dense,
                                flops_executed_model=flops_exec)
    tot = sharding.allreduce_stats(local, device=dev)
    mass_err = float((ens.y.sum(0) - 1.0).abs().max())

    # ---------------- e2e: the public C-ABI call with host buffers
    def e2e_call():
        res = capi.Result()
        capi.check(capi.lib.rehuel_irk_ensemble(
            capi.ROBERTSON, capi.RADAU_IIA_53, t0, t1, DT0, M, 3,
            capi.C.cast(y0_host.data_ptr(), capi._PD), 1, capi.C.cast(P_host.data_ptr(), capi._PD),
            capi.C.byref(opts), 0, 1, capi.C.byref(res)))
        d2h = (3 + 1 + 1) * 8 * M + 4 * M + 4 * M * capi.NCOUNTERS
        checksum = res.sum.attempt
        capi.lib.rehuel_result_free(capi.C.byref(res))
        return d2h, checksum

    for _ in range(max(1, args.warmup)):
        e2e_call()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        d2h_bytes, chk = e2e_call()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - e0
    et = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_value = world * M * args.steps / et.item()
    clocks = sampler.stop() if sampler else None  # sampled over the device-timed AND the e2e region

    # ---------------- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        cpu = cpu_reference_run(args.span, (4096 if args.span == "short" else 64) * cores, cores)

    if rank == 0:
        ms_per_step = total_ms_max / args.steps
        value = world * M * args.steps / (total_ms_max * 1e-3)
        kern_s = kernel_ms_max / args.steps * 1e-3
        ach_dense = tot["flops_dense_model"] / world / kern_s / 1e12   # per GPU, per launch
        ach_exec = tot["flops_executed_model"] / world / kern_s / 1e12
        info = capi.kernel_info(capi.ROBERTSON, capi.RADAU_IIA_53)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.span, M), "members_total": world * M, "span": args.span,
                       "parallelism": f"member-sharded x{world} (no data-path collective)",
                       "l2": "flushed between timed steps (256 MiB fill, outside the per-step event pair)",
                       "timing": "sum of per-step CUDA-event durations on the launching stream, max over ranks",
                       "handout_order": args.handout + (" (Morton order of the rate constants, recomputed on the device inside every timed step"
                                                       + ("; walked from the large-k1 end, whose members need up to 39 attempts against 23, so the "
                                                          "queue drains on cheap members" if args.handout == "locality_desc" else "") + ")" if handout else ""),
                       "wall_s_timed_region_rank0": wall, "host_binding_rank0": numa_note},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(P_host.numel() * 8 + 24),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": 1e3 * et.item() / args.steps,
                    "api": "rehuel_irk_ensemble (C ABI, pinned host buffers, results copied back)"},
            "gpu_launches": int(ll.item()),
            "clocks": clocks,
            "roofline": {
                "bound": "fp64", "achieved": ach_dense, "peak": peak_tflops, "unit": "TFLOP/s",
                "frac": ach_dense / peak_tflops, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                "traffic_note": "DRAM bytes read + written per launch of this kernel, ncu --set full capture of this command "
                                "(profiles/r1_v8_ncu_summary.txt): 5x the 109 MB of algorithmic I/O because members are handed "
                                "out in parameter order, so every 8-byte result lands in its own 32-byte sector; 2% of HBM peak",
                "kernel_ms_per_launch": kernel_ms_max / args.steps,
                "model": "SURVEY.md 8(d) dense-algorithm flops: 873*attempts + 321*newton_iters + 41*uses_of_8.20, from the kernel's per-member counters",
                "peak_source": "measured in this run: register-resident DFMA chains (rehuel_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                "peak_nominal": FP64_NOMINAL_TFLOPS, "frac_of_nominal": ach_dense / FP64_NOMINAL_TFLOPS,
                "achieved_executed": ach_exec, "frac_executed": ach_exec / peak_tflops,
                "model_executed": "flops of the eigen-decoupled algorithm actually run: 259*attempts + 369*newton_iters + 30*uses_of_8.20 - 132*skipped_final_residuals (DESIGN.md section 6)",
                "kernel": "ensemble_irk_thread<RobertsonF,3,1,1,EST_REUSE>", "regs": info["regs"],
                "spill_bytes": info["local_bytes"], "ctas_per_sm": info["ctas_per_sm"], "block": info["block"],
            },
            "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
            "stats": {"attempts_per_member": tot["attempt"] / tot["members"], "accepted_per_member": tot["steps"] / tot["members"],
                      "rejected_err": tot["reject_err"], "rejected_newton": tot["reject_newton"],
                      "newton_iters_per_attempt": tot["newton_iters"] / max(tot["attempt"], 1),
                      "fun_evals": tot["fun_evals"], "jac_evals": tot["jac_evals"], "failed_members": tot["n_failed"],
                      "max_attempts": tot["max_attempt"], "mass_conservation_max_abs_err_rank0": mass_err,
                      "step_ms": step_ms, "kernel_ms": kernel_ms},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--span", choices=tuple(SPANS), default="short")
    ap.add_argument("--members", type=int, default=M_PER_GPU, help="members per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--handout", choices=("index", "locality", "locality_desc"), default="locality_desc",
                    help="order in which members are handed to the GPU lanes (results do not depend on it)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
