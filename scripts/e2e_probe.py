"""Diagnostic: ms per frame of the host-buffer pipeline under a few settings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uf3_b200 import geometry, synthetic
from uf3_b200.pipeline import FramePipeline

basis = synthetic.w_basis("demo")
frames = [synthetic.bcc_w((10, 20, 25), seed=i) for i in range(8)]
images = geometry.image_table(frames[0][2], frames[0][3], basis.r_cut)
h_pos = [torch.from_numpy(fr[0]).pin_memory().numpy() for fr in frames]
h_num = torch.from_numpy(frames[0][1]).pin_memory().numpy()

def run(depth, steps=60, forces=True):
    pipe = FramePipeline(basis, 10000, device=0, depth=depth, forces=forces)
    def go(n):
        prev = []
        for k in range(n):
            prev.append(pipe.submit(h_pos[k % 8], h_num, images))
            if len(prev) == depth:
                pipe.result(prev.pop(0))
        for s in prev:
            pipe.result(s)
        torch.cuda.synchronize()
    go(10)
    t0 = time.perf_counter(); go(steps); ms = (time.perf_counter() - t0) * 1e3 / steps
    pipe.close()
    return ms

for depth in (2, 3, 4):
    print(f"depth {depth}: {run(depth):.3f} ms/frame", flush=True)
print(f"depth 2, energy rows only: {run(2, forces=False):.3f} ms/frame")
# raw D2H bandwidth of one 17.5 MB row block
d = torch.empty((30000, 73), dtype=torch.float64, device="cuda"); h = torch.empty((30000, 73), dtype=torch.float64).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print(f"D2H alone: {dt*1e3:.3f} ms per block, {d.numel()*8/dt/1e9:.1f} GB/s")

# native pipeline (C worker threads, no torch on the path)
from uf3_b200.pipeline import NativePipeline
def run_native(depth, steps=60):
    pipe = NativePipeline(basis, depth=depth, device=0)
    F = pipe.n_feats
    outs = [(torch.empty(F, dtype=torch.float64).pin_memory().numpy(),
             torch.empty((30000, F), dtype=torch.float64).pin_memory().numpy()) for _ in range(depth)]
    def go(n):
        tickets = []
        for k in range(n):
            xe, xf = outs[k % depth]
            tickets.append(pipe.submit(h_pos[k % 8], h_num, images, xe, xf))
            if len(tickets) == depth:
                pipe.wait(tickets.pop(0))
        for t in tickets:
            pipe.wait(t)
    go(10)
    t0 = time.perf_counter(); go(steps); ms = (time.perf_counter() - t0) * 1e3 / steps
    pipe.close()
    return ms
for depth in (2, 3, 4, 6):
    print(f"native pipeline depth {depth}: {run_native(depth):.3f} ms/frame", flush=True)
