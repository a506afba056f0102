"""Diagnostic: per-frame stage timestamps of the host-buffer pipeline (CUDA events)."""
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
pipe = FramePipeline(basis, 10000, device=0, depth=2)
ev = []
t_ref = torch.cuda.Event(enable_timing=True)
def submit(k):
    slot = pipe.turn; pipe.turn = (pipe.turn + 1) % 2
    n = 10000
    if pipe.busy[slot]: pipe.copied[slot].synchronize()
    pipe.busy[slot] = True; pipe.n_atoms[slot] = n
    compute, eng = pipe.compute[slot], pipe.engines[slot]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    h0 = time.perf_counter()
    with torch.cuda.stream(compute):
        e[0].record(compute)
        pipe.d_pos[slot][:n].copy_(torch.from_numpy(h_pos[k % 8]), non_blocking=True)
        pipe.d_num[slot][:n].copy_(torch.from_numpy(h_num), non_blocking=True)
        eng.build_neighbors_device(pipe.d_pos[slot].data_ptr(), pipe.d_num[slot].data_ptr(), n, images, compute.cuda_stream)
        e[1].record(compute)
        eng.featurize_device(pipe.d_xe[slot].data_ptr(), pipe.d_xf[slot].data_ptr(), pipe.F, compute.cuda_stream)
        e[2].record(compute)
        pipe.computed[slot].record(compute)
    h1 = time.perf_counter()
    with torch.cuda.stream(pipe.copy):
        pipe.copy.wait_event(pipe.computed[slot])
        e[3].record(pipe.copy)
        pipe.h_xf[slot][:3 * n].copy_(pipe.d_xf[slot][:3 * n], non_blocking=True)
        pipe.h_xe[slot].copy_(pipe.d_xe[slot], non_blocking=True)
        e[4].record(pipe.copy)
        pipe.copied[slot].record(pipe.copy)
    ev.append((e, h0, h1))
    return slot
t_ref.record(); torch.cuda.synchronize(); host_ref = time.perf_counter()
prev = None
for k in range(14):
    s = submit(k)
    if prev is not None: pipe.result(prev)
    prev = s
pipe.result(prev); torch.cuda.synchronize()
print("frame: build_start build_end feat_end | d2h_start d2h_end | host submit start/end (ms since start)")
for k, (e, h0, h1) in enumerate(ev):
    ts = [t_ref.elapsed_time(x) for x in e]
    print(k, " ".join(f"{t:7.3f}" for t in ts[:3]), "|", " ".join(f"{t:7.3f}" for t in ts[3:]), "|", f"{(h0-host_ref)*1e3:7.3f} {(h1-host_ref)*1e3:7.3f}")
