"""W 100k-atom MD force step (configs[4]) alone on one GPU: neighbour lists + energy + forces, with
event timings of the two halves.  UF3B_FAKE_WORLD=k: the share of rank 0 of k."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uf3_b200 import geometry, synthetic
from uf3_b200.distributed import atom_range
from uf3_b200.engine import Engine
basis, coeff = bench.w_model23()
pos, numbers, cell, pbc = synthetic.bcc_w((25, 40, 50), a=3.206, sigma=0.15, seed=0)
n = len(pos)
images = geometry.image_table(cell, pbc, basis.r_cut)
eng = Engine(basis, device=0, deferred_lists=bool(int(os.environ.get('UF3B_DEFERRED', '1'))))
eng.set_coefficients(coeff)
d_pos = torch.from_numpy(pos).cuda(); d_num = torch.from_numpy(numbers).cuda()
d_e = torch.zeros(1, dtype=torch.float64, device="cuda"); d_f = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
world = int(os.environ.get("UF3B_FAKE_WORLD", "1"))
centres = atom_range(n, 0, world) if world > 1 else None
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
def build():
    eng.build_neighbors_device(d_pos.data_ptr(), d_num.data_ptr(), n, images, stream, centres=centres)
def evalf():
    eng.energy_forces_device(d_e.data_ptr(), d_f.data_ptr(), stream)
for _ in range(3):
    build(); evalf()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
ev[0].record()
for k in range(steps):
    build(); ev[2 * k + 1].record(); evalf(); ev[2 * k + 2].record()
torch.cuda.synchronize()
tb = sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(steps)) / steps
te = sum(ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(steps)) / steps
eng.set_timing(True); build(); evalf(); k_ms = eng.last_kernel_ms(); eng.set_timing(False)
print(json.dumps({"world": world, "ms_per_step": ev[0].elapsed_time(ev[-1]) / steps, "build_ms": tb, "eval_ms": te,
                  "k_energy_forces_ms": k_ms, "list2_per_atom": eng.neighbor_count(2) / n, "list3_per_atom": eng.neighbor_count(3) / n}))
