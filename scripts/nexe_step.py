"""Ne/Xe 50k inference step (configs[2]) alone, for launch lists."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uf3_b200 import geometry, synthetic
from uf3_b200.engine import Engine
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
data = np.load(os.path.join(ROOT, "tests", "golden", "calc_syn_nexe64_pair.npz"))
basis = synthetic.nexe_basis()
pos, numbers, cell, pbc = synthetic.nexe((25, 25, 10), seed=0)
n = len(pos)
images = geometry.image_table(cell, pbc, basis.r_cut)
eng = Engine(basis, device=0)
eng.set_coefficients(np.array(data["coefficients"]))
d_pos = torch.from_numpy(pos).cuda(); d_num = torch.from_numpy(numbers).cuda()
d_e = torch.zeros(1, dtype=torch.float64, device="cuda"); d_f = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(3):
    eng.build_neighbors_device(d_pos.data_ptr(), d_num.data_ptr(), n, images, stream); eng.energy_forces_device(d_e.data_ptr(), d_f.data_ptr(), stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    eng.build_neighbors_device(d_pos.data_ptr(), d_num.data_ptr(), n, images, stream); eng.energy_forces_device(d_e.data_ptr(), d_f.data_ptr(), stream)
e1.record(); torch.cuda.synchronize()
print("ms/step", e0.elapsed_time(e1) / steps, "pairs/atom", eng.neighbor_count(2) / n, "images", len(images[1]))
