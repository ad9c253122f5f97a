"""GPU experiment (uses the oracle as the CPU side): SURVEY 8(f) row 1, geometric bootstrapping of one KITTI-sized pair.
    python tests/tools/bootstrap_bench.py
Times sphericity_based_voxel_analysis + two voxel_down_sample calls on the GPU (CUDA events, clouds resident) against
the float64 NumPy restatement of the reference path (sklearn PCA / Open3D semantics) on the host."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bufferx_b200 as bx
from bufferx_b200 import ops
from bufferx_b200.synth import make_pair
from bufferx_b200.bootstrap import sphericity_based_voxel_analysis, voxel_down_sample
from oracle import oracle as O

data = make_pair("C3", 0)
src, tgt = data["src_fds_pcd"], data["tgt_fds_pcd"]
st = np.random.RandomState(0)
i_s = st.choice(len(src), len(src) // 10, replace=False)
i_t = st.choice(len(tgt), len(tgt) // 10, replace=False)
dev = torch.device("cuda")
ds, dt = torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev)
dis, dit = torch.from_numpy(i_s.astype(np.int32)).to(dev), torch.from_numpy(i_t.astype(np.int32)).to(dev)
for r in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    v, sph, al = sphericity_based_voxel_analysis(ds, dt, dis, dit, device=dev)
    ks, xs, _ = ops.voxel_down_sample(ds, v)
    kt, xt, _ = ops.voxel_down_sample(dt, v)
    b.record()
    torch.cuda.synchronize()
    gpu_ms = a.elapsed_time(b)
t0 = time.perf_counter()
ov, osph, oal = O.sphericity_based_voxel_analysis(src, tgt, i_s, i_t)
ek, em, _ = O.voxel_down_sample(src, ov)
ek2, em2, _ = O.voxel_down_sample(tgt, ov)
cpu_ms = (time.perf_counter() - t0) * 1e3
assert v == ov and al == oal and len(ks) == len(ek) and len(kt) == len(ek2)
print(f"C3 pair (2 x {len(src)} points): voxel {v} m, sphericity {sph:.4f}, aligned {al}; {len(ks)} + {len(kt)} voxels")
print(f"GPU {gpu_ms:.3f} ms (incl. three small host reads)   CPU NumPy restatement {cpu_ms:.1f} ms   ratio {cpu_ms / gpu_ms:.0f}x")
