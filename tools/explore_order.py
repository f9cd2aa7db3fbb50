"""Dev aid: does handing out members in a parameter-locality order (Morton code of the log-parameters)
shorten the Robertson config-2 kernel?  Times index order vs several orders on the device."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rehuel_b200 import capi, ensembles
from rehuel_b200.device import DeviceEnsemble

def timed(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def morton(u, bits):
    q = np.minimum((u * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    key = np.zeros(u.shape[0], dtype=np.uint64)
    for b in range(bits):
        for d in range(u.shape[1]):
            key |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * u.shape[1] + d)
    return key

M = 1 << 20
o = capi.make_options(1e-6, 1e-6)
Ph = ensembles.robertson_params(M)
u = np.log10(Ph.T / np.array([0.04, 1e4, 3e7])) + 0.5
P = torch.from_numpy(Ph).cuda(); y0 = torch.from_numpy(ensembles.ROBERTSON_Y0.copy()).cuda()
orders = {"index": None}
for bits in (4, 7, 10):
    orders[f"morton{bits}"] = np.argsort(morton(u, bits), kind="stable").astype(np.int32)
orders["sort_k1"] = np.argsort(u[:, 0]).astype(np.int32)
orders["sort_k2"] = np.argsort(u[:, 1]).astype(np.int32)
orders["sort_k3"] = np.argsort(u[:, 2]).astype(np.int32)
for t1, span in ((40.0, "short"), (1e11, "long")):
    Mu = M if span == "short" else M >> 2
    for name, od in orders.items():
        if od is not None and Mu != M:
            od = od[od < Mu]
        de = DeviceEnsemble(capi.ROBERTSON, 211, Mu, y0, P[:, :Mu].contiguous(), order=None if od is None else torch.from_numpy(od.copy()).cuda())
        ms = timed(lambda: de.solve(0.0, t1, 1e-6, o))
        att = de.counter("attempt").cpu().numpy().astype(np.int64); it = de.counter("newton_iters").cpu().numpy().astype(np.int64)
        if name == "index":
            ref_y = de.y.clone()
        same = bool(torch.equal(ref_y, de.y))
        print(f"{span:5s} {name:9s} {ms:8.3f} ms  attempts/member {att.mean():.2f} iters/attempt {it.sum()/att.sum():.3f} identical_to_index={same}", flush=True)
