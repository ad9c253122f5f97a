"""GPU experiment: bx_spt_pnt alone on the real patches of a C2 pair (one timing per scale).
    python tools/spt_bench.py [reps]
Under ncu:  ncu --set full --import-source on -k regex:spt_pnt --launch-skip 6 -c 3 python tools/spt_bench.py 1"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bufferx_b200 as bx
from bufferx_b200 import ops
from bufferx_b200.synth import init_synthetic_weights, make_pair, workload_cfg

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
cfg = workload_cfg("C2")
cfg.match.iter_n = 1000
model = init_synthetic_weights(bx.BufferX(cfg)).cuda()
data = make_pair("C2", 0)
np.random.seed(0)
with torch.no_grad():
    model(data, ransac_seed=0, debug=True)
dbg = model.last_debug
prep = model.Desc.prepared(torch.device("cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i, sc in enumerate(dbg["scales"]):
    delta = sc["s"]["patches"].contiguous()
    nz = (delta.abs().sum(-1) > 0).float().sum(1).mean().item()
    ts = []
    for r in range(reps + 1):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.spt_pnt(delta, prep["voxels"], prep["rot"], 0.8 / 3, 10, prep["w_pnt"], prep["b_pnt"], 20)
        b.record()
        torch.cuda.synchronize()
        if r:
            ts.append(a.elapsed_time(b))
    print(f"scale {i}: K={delta.shape[0]} non-zero points/patch {nz:.0f}  spt_pnt {1e3 * np.mean(ts):.1f} us")
