"""Times the fit path on the binary Fe-C workload (B2 lattice, basis of the reference's
tests/test_representation.py:605-648) and checks a small frame against the oracle.
    python scripts/binary_step.py [reps_x reps_y reps_z] [--a 2.87] [--check]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uf3_b200 import geometry, synthetic  # noqa: E402
from uf3_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("reps", nargs="*", type=int, default=[10, 20, 25])
ap.add_argument("--a", type=float, default=2.87)
ap.add_argument("--check", action="store_true")
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
basis = synthetic.fec_basis()
dev = torch.device("cuda", 0)
if args.check:
    from oracle import uf3_oracle as orc
    pos, numbers, cell, pbc = synthetic.b2_fec((3, 3, 4), a=args.a, seed=5)
    images = geometry.image_table(cell, pbc, basis.r_cut)
    eng = Engine(basis, device=0)
    eng.build_neighbors(pos, numbers, images=images)
    xe, xf = eng.featurize()
    t0 = time.perf_counter()
    we, wf = orc.featurize(orc.PackedBasis(basis), pos, numbers, images[1])
    print("oracle s", time.perf_counter() - t0, "n", len(pos))
    print("energy row rel err", np.abs(xe - we).max() / np.abs(we).max(), "force rows rel err", np.abs(xf - wf).max() / np.abs(wf).max())
    eng.close()
pos, numbers, cell, pbc = synthetic.b2_fec(tuple(args.reps), a=args.a, seed=0)
n = len(pos)
images = geometry.image_table(cell, pbc, basis.r_cut)
eng = Engine(basis, device=0)
F = eng.n_feats
stream = torch.cuda.current_stream().cuda_stream
d_pos = torch.from_numpy(pos).to(dev)
d_num = torch.from_numpy(numbers).to(dev)
d_xe = torch.empty(F, dtype=torch.float64, device=dev)
d_xf = torch.empty((3 * n, F), dtype=torch.float64, device=dev)
for _ in range(2):
    eng.build_neighbors_device(d_pos.data_ptr(), d_num.data_ptr(), n, images, stream)
    eng.featurize_device(d_xe.data_ptr(), d_xf.data_ptr(), F, stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    eng.build_neighbors_device(d_pos.data_ptr(), d_num.data_ptr(), n, images, stream)
    eng.featurize_device(d_xe.data_ptr(), d_xf.data_ptr(), F, stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
eng.set_timing(True)
eng.featurize_device(d_xe.data_ptr(), d_xf.data_ptr(), F, stream)
print({"n": n, "F": F, "ms_per_frame": ms, "kernel_ms": eng.last_kernel_ms(), "list3_per_atom": eng.neighbor_count(3) / n,
       "pairs_per_atom": eng.neighbor_count(2) / n, "atom_steps_per_s": n / ms * 1e3})
