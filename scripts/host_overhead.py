"""Host-side cost of the C-ABI calls of one frame (10k-atom W, demo basis): how long each call keeps
the calling thread, with the GPU work of featurize / gram left asynchronous."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uf3_b200 import geometry, least_squares as ls
from uf3_b200.engine import Engine
basis = bench.make_basis("demo")
pos, numbers, cell, pbc = bench.frame(0)
n = len(pos)
images = geometry.image_table(cell, pbc, basis.r_cut)
eng = Engine(basis, device=0)
F = eng.n_feats
d_pos = torch.from_numpy(pos).cuda(); d_num = torch.from_numpy(numbers).cuda()
d_xe = torch.zeros(F, dtype=torch.float64, device="cuda"); d_xf = torch.zeros((3 * n, F), dtype=torch.float64, device="cuda")
d_y = torch.zeros(3 * n, dtype=torch.float64, device="cuda")
acc = ls.GramAccumulator(F)
stream = torch.cuda.current_stream().cuda_stream
tb = tf = tg = ts = 0.0
steps = 50
for k in range(steps + 5):
    if k == 5:
        tb = tf = tg = ts = 0.0
    t0 = time.perf_counter()
    eng.build_neighbors_device(d_pos.data_ptr(), d_num.data_ptr(), n, images, stream)
    t1 = time.perf_counter()
    eng.featurize_device(d_xe.data_ptr(), d_xf.data_ptr(), F, stream)
    t2 = time.perf_counter()
    acc.add_force_rows_device(d_xf.data_ptr(), d_y.data_ptr(), 3 * n, F, stream, y_moments=(0.0, 0.0))
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    tb += t1 - t0; tf += t2 - t1; tg += t3 - t2; ts += t4 - t3
print(json.dumps({"build_call_us": 1e6 * tb / steps, "featurize_call_us": 1e6 * tf / steps, "gram_call_us": 1e6 * tg / steps,
                  "final_sync_us": 1e6 * ts / steps, "cpus": os.cpu_count()}))
