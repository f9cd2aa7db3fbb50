"""BASELINE config 5 (mixed-stiffness, mixed-tolerance ensemble) and config 3 (Van der Pol sweep), sharded over the
ranks of a torchrun launch by member index modulo world size; single process = one GPU.
  python tools/config5.py [--members 10485760] [--variants]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/config5.py
Device-timed (CUDA events, max over ranks); prints one JSON line on rank 0."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from rehuel_b200 import capi, ensembles
from rehuel_b200.device import DeviceEnsemble

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=10485760)
ap.add_argument("--vdp-members", type=int, default=1 << 20)
ap.add_argument("--variants", action="store_true", help="time alternative hand-out orders (1 GPU)")
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)

def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    t = torch.tensor([best], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

o = capi.make_options(1e-6, 1e-6)
out = {"n_gpus": world}
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

# ---- config 5: even members Robertson (config-2 sweep), odd members Van der Pol (config-3 sweep), rtol = atol = 10^(-3-6v)
M = args.members
half = M // 2
tol = ensembles.mixed_tolerances(M)
Pr_all, Pv_all = ensembles.robertson_params(half), ensembles.vdpol_params(half)
sel = np.arange(rank, half, world)                      # members i = 2*sel (+1) of this rank: round-robin over the ranks
tr, tv = tol[0::2][sel], tol[1::2][sel]
Pr, Pv = Pr_all[:, sel], Pv_all[:, sel]
m = len(sel)
def make(order_mode, keyed):
    ens = []
    for prob, method, t1, y0h, P, tl in ((capi.ROBERTSON, 211, 40.0, ensembles.ROBERTSON_Y0, Pr, tr),
                                         (capi.VAN_DER_POL, 230, 2.0, ensembles.VDPOL_Y0, Pv, tv)):
        keys = None
        if keyed:  # the tolerance is a cost dimension like any parameter
            k = np.vstack([tl[None, :], P[:2]]) if prob == capi.ROBERTSON else np.vstack([tl[None, :], P])
            keys = cu(k)
        tl_d = cu(tl)
        ens.append((DeviceEnsemble(prob, method, m, cu(y0h.copy()), cu(P), rtol=tl_d, atol=tl_d.clone(),
                                   newton_tol=(0.1 * tl_d).contiguous(), handout_order=order_mode, handout_keys=keys), t1))
    return ens
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def run(ens):
    cur = torch.cuda.current_stream()
    for (de, t1), st in zip(ens, streams):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            de.solve(0.0, t1, 1e-6, o)
    for st in streams:
        cur.wait_stream(st)
variants = [("index", capi.ORDER_INDEX, False), ("locality(params)", capi.ORDER_LOCALITY, False),
            ("locality(tol,params)", capi.ORDER_LOCALITY, True), ("locality_desc(tol,params)", capi.ORDER_LOCALITY_DESC, True)]
if not args.variants:
    variants = variants[2:3]
for name, mode, keyed in variants:
    ens = make(mode, keyed)
    ms = timed(lambda: run(ens), reps=1)
    failed = torch.tensor([float(sum(int((de.status != 0).sum()) for de, _ in ens))], device=dev)
    att = torch.tensor([float(de.counter("attempt").double().sum()) for de, _ in ens], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(failed); dist.all_reduce(att)
    out[f"config5_mixed_{M}_{name}"] = dict(ms=ms, systems_per_s=M / ms * 1e3, failed=int(failed.item()),
                                           attempts_per_member=[float(att[0]) / half, float(att[1]) / half])
    del ens

# ---- config 3: Van der Pol, mu_ref = 10^(-6 i/(M-1)), LOBATTO_IIIC_43, sharded i mod world
Mv = args.vdp_members
Pv3 = ensembles.vdpol_params(Mv)[:, rank::world]
for name, mode in (("locality", capi.ORDER_LOCALITY),) + ((("index", capi.ORDER_INDEX), ("locality_desc", capi.ORDER_LOCALITY_DESC)) if args.variants else ()):
    de = DeviceEnsemble(capi.VAN_DER_POL, 230, Pv3.shape[1], cu(ensembles.VDPOL_Y0.copy()), cu(Pv3), handout_order=mode)
    ms = timed(lambda: de.solve(0.0, 2.0, 1e-6, o))
    out[f"config3_vdpol_{Mv}_{name}"] = dict(ms=ms, systems_per_s=Mv / ms * 1e3)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
