#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: atom-steps/s featurized (2+3-body) on 10k-atom bulk W.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--basis demo|manuscript]
    python bench.py --impl reference ...       # CPU arm: the oracle port on the host cores

One step = one 10 000-atom frame taken through the whole hot path: neighbour lists
(Kernel A) + energy row + 3N force rows (Kernel B).  `value` has the frame resident in HBM
and leaves the rows in HBM.  `e2e` is the path north_star describes, through the library's own
frame pipeline with HOST buffers (uf3b_pipeline_submit_fit): pinned positions and force targets
go up every step, the rows stay in HBM and are folded into the normal equations
(uf3b_gram_accumulate), the frame's energy row comes back, and the timed region ends with the ONE
all-reduce of the normal equations and the regularised solve (coefficients on the host).
`e2e_rows_to_host` is the same frames with every row copied to the host instead (the copy, not the
kernels, bounds it).  Under torchrun every rank takes its own frames (weak scaling, no collective
on the data path); time = max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "bulk bcc W 10x20x25 cells (10000 atoms/frame), sigma=0.05 A, 2+3-body featurization"
# --basis binary: the Fe-C basis of the reference's own golden rows (tests/test_representation.py:605-648,
# six trios, two of symmetry 1, F = 609) on a B2 lattice of the same 10x20x25 cells
WORKLOAD_BINARY = ("B2 Fe-C 10x20x25 cells (10000 atoms/frame), a=2.87 A, sigma=0.05 A, 2+3-body featurization "
                   "(55 neighbours inside the 5 A three-body cutoff)")
N_POOL = 8          # distinct frames (inputs AND output row buffers) per rank, cycled
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernels, from the ncu captures
# summarised under profiles/ (profiles/ncu_constants.json says which capture each number comes from)
NCU_TRAFFIC_BYTES = {"demo": None, "manuscript": None, "binary": None}
# executed FP64 flops per launch of the dominant kernels from the same captures (2 per fused, 1 per
# non-fused thread instruction, predicated-off lanes excluded)
NCU_FP64_FLOP = {"demo": None, "manuscript": None, "binary": None}
try:
    with open(os.path.join(ROOT, "profiles", "ncu_constants.json")) as _fh:
        _c = json.load(_fh)
    NCU_TRAFFIC_BYTES.update(_c.get("dram_bytes_per_launch", {}))
    NCU_FP64_FLOP.update(_c.get("fp64_flop_per_launch", {}))
except (OSError, ValueError):
    pass


def bench_config(basis_name, n_feats):
    """`config` of the JSON line: the same dictionary in both arms (the driver compares them)."""
    return {"workload": WORKLOAD_BINARY if basis_name == "binary" else WORKLOAD, "basis": basis_name, "n_feats": int(n_feats),
            "step": "one 10000-atom frame per rank: neighbour lists + energy row + 3N force rows"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons WHILE the timed region runs: an NVML polling thread
    (5 ms period; ctypes calls release the GIL, so the bench loop is not held up), with
    `nvidia-smi -lms` as the fallback when NVML cannot be loaded."""
    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
               ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
               ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.sm, self.smax, self.reasons = [], None, set()
        self.thread = self.proc = self.tmp = None
        self.stop_flag = False

    def _poll(self, nv, handle):
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                for name, attr in self.REASONS:
                    if mask & getattr(nv, attr):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() \
                else self.index
            handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.smax = nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out.update(sm_mhz=statistics.median(self.sm), sm_max_mhz=self.smax, samples=len(self.sm),
                           source="nvml, 5 ms period")
            out["reasons"] = sorted(self.reasons)
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[4:8]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(smax)
            out["samples"] = len(sm)
            out["source"] = "nvidia-smi -lms 100"
        out["reasons"] = sorted(reasons)
        return out


BASIS_KIND = "demo"      # set from --basis: frame() follows it


def make_basis(kind):
    from uf3_b200 import synthetic
    return synthetic.fec_basis() if kind == "binary" else synthetic.w_basis(kind)


def frame(seed):
    from uf3_b200 import synthetic
    if BASIS_KIND == "binary":
        return synthetic.b2_fec((10, 20, 25), seed=seed)
    return synthetic.bcc_w((10, 20, 25), seed=seed)


# --------------------------------------------------------------------------- CPU arms
def oracle_frames_per_second(basis, n_frames, threads):
    """Oracle port (oracle/uf3_oracle.c) on `n_frames` full frames, `threads` at a time."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import uf3_oracle as orc
    from uf3_b200 import geometry
    packed = orc.PackedBasis(basis)
    frames = [frame(1000 + i) for i in range(n_frames)]
    images = geometry.image_table(frames[0][2], frames[0][3], basis.r_cut)

    def work(fr):
        orc.featurize(packed, fr[0], fr[1], images[1], energy=True, forces=True)

    t0 = time.perf_counter()
    if threads == 1:
        for fr in frames:
            work(fr)
    else:
        with ThreadPoolExecutor(max_workers=threads) as pool:
            list(pool.map(work, frames))
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    """CPU arm.  The reference is pure Python and holds dense O(M^2) matrices (it cannot
    represent a 10k-atom frame: SURVEY.md fact 2) and cannot travel to the GPU box, so the
    arm times the C restatement of its algorithm (the oracle port) on all host threads."""
    if rank != 0:
        return
    basis = make_basis(args.basis)
    cores = os.cpu_count() or 1
    n_atoms = 10000
    for _ in range(min(args.warmup, 1)):
        oracle_frames_per_second(basis, cores, cores)
    times = [oracle_frames_per_second(basis, cores, cores) for _ in range(args.steps)]
    total = sum(times)
    value = cores * n_atoms * args.steps / total
    line = {
        "impl": "reference", "metric": "atom-steps/s featurized (2+3-body) on 10k-atom W",
        "value": value, "unit": "atom-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": bench_config(args.basis, basis.n_feats),
        "notes": {"step": f"{cores} frames of 10000 atoms per step, one per host thread"},
        "cpu_baseline": {"value": value, "unit": "atom-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} full 10k-atom frames per step (energy row + 3N force "
                                   "rows), oracle/uf3_oracle.c, one frame per thread"},
        "e2e": {"value": value, "unit": "atom-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- inference configs
def w_model23():
    """The shipped W 2+3-body model (examples/tungsten_extxyz/model_2and3.json), carried as a test
    fixture because /root/reference does not exist on the GPU box."""
    from uf3_b200 import bspline, composition
    data = np.load(os.path.join(ROOT, "tests", "golden", "calc_syn_w54_model23.npz"))
    cfg = json.loads(str(data["config"]))
    knots = {}
    for key, val in cfg["kwargs"]["knots_map"].items():
        parts = tuple(key.split("-"))
        knots[parts] = np.array(val) if len(parts) == 2 else [np.array(v) for v in val]
    chem = composition.ChemicalSystem(cfg["element_list"], degree=cfg["degree"])
    lead = {int(k): v for k, v in cfg["kwargs"]["leading_trim"].items()}
    trail = {int(k): v for k, v in cfg["kwargs"]["trailing_trim"].items()}
    basis = bspline.BSplineBasis(chem, knots_map=knots, leading_trim=lead, trailing_trim=trail)
    return basis, np.array(data["coefficients"])


def inference_extras(torch, dev, stream, steps=20):
    """BASELINE.json configs[2] and configs[4] at one GPU: energy + forces per step through
    the C ABI (neighbour lists + evaluator), inputs resident / via host buffers.  Reported
    beside the headline; not part of `value`."""
    from uf3_b200 import geometry, synthetic
    from uf3_b200.engine import Engine

    def nexe_model():
        data = np.load(os.path.join(ROOT, "tests", "golden", "calc_syn_nexe64_pair.npz"))
        return synthetic.nexe_basis(), np.array(data["coefficients"])

    out = {}
    for tag, (basis, coeff), fr in (
            ("nexe_50k_energy_forces", nexe_model(), synthetic.nexe((25, 25, 10), seed=0)),
            ("w_100k_md_step_energy_forces", w_model23(), synthetic.bcc_w((25, 40, 50), a=3.206, sigma=0.15, seed=0))):
        # list builds as an MD loop issues them: the previous cell grid is reused and the status of a build is
        # checked behind the evaluator's launch (uf3b_basis_set_deferred_lists), one host wait per step
        eng = Engine(basis, device=dev.index, deferred_lists=True)
        eng.set_coefficients(coeff)
        pos, numbers, cell, pbc = fr
        n = len(pos)
        images = geometry.image_table(cell, pbc, basis.r_cut)
        d_pos = torch.from_numpy(pos).to(dev)
        d_num = torch.from_numpy(numbers).to(dev)
        d_e = torch.zeros(1, dtype=torch.float64, device=dev)
        d_f = torch.zeros((n, 3), dtype=torch.float64, device=dev)
        h_pos = torch.from_numpy(pos).pin_memory().numpy()

        def resident():
            eng.build_neighbors_device(d_pos.data_ptr(), d_num.data_ptr(), n, images, stream)
            eng.energy_forces_device(d_e.data_ptr(), d_f.data_ptr(), stream)

        eng_h = Engine(basis, device=dev.index)      # host arrays in and out: checked builds
        eng_h.set_coefficients(coeff)

        def host():
            eng_h.build_neighbors(h_pos, numbers, images=images, stream=stream)
            eng_h.energy_forces(stream=stream)

        res = {}
        for name, fn in (("resident", resident), ("e2e", host)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res[name] = {"ms_per_step": ms, "atom_steps_per_s": n / (ms * 1e-3)}
        eng.set_timing(True)
        resident()
        res["k_energy_forces_ms"] = eng.last_kernel_ms()
        eng.set_timing(False)
        res["n_atoms"] = n
        res["n_feats"] = int(basis.n_feats)
        res["pairs_per_atom"] = eng.neighbor_count(2) / n
        # algorithmic bytes of the evaluator launch: positions + species + both CSR lists + forces
        alg = n * 28 + 4 * (eng.neighbor_count(2) + eng.neighbor_count(3)) + 8 * (n + 1) + 24 * n + 8
        res["hbm_GBps_algorithmic"] = alg / (res["k_energy_forces_ms"] * 1e-3) / 1e9
        out[tag] = res
        eng.close()
        eng_h.close()
    return out


# --------------------------------------------------------------------------- GPU arm
def md_extra(torch, dist, dev, rank, world, steps=20, warmup=3, dt=0.2):
    """BASELINE.json configs[4] beside the headline when several GPUs run: the 100 000-atom W MD
    step (neighbour lists + energy + forces, atom ranges over the ranks), strong scaling."""
    from uf3_b200 import geometry, synthetic
    from uf3_b200.distributed import ShardedEvaluator
    basis, coeff = w_model23()
    pos, numbers, cell, pbc = synthetic.bcc_w((25, 40, 50), a=3.206, sigma=0.15, seed=0)
    n = len(pos)
    images = geometry.image_table(cell, pbc, basis.r_cut)
    ev = ShardedEvaluator(basis, coeff, device=dev.index)
    x = torch.from_numpy(pos).to(dev)
    z = torch.from_numpy(numbers).to(dev)
    v = torch.zeros_like(x)
    inv_m = 9.648533e-3 / 183.84

    def step(f):
        v.add_(f, alpha=0.5 * dt * inv_m)
        x.add_(v, alpha=dt)
        e, f_new = ev.energy_forces(x, z, images)
        v.add_(f_new, alpha=0.5 * dt * inv_m)
        return e, f_new

    e0, f = ev.energy_forces(x, z, images)
    e0, f = float(e0), f.clone()
    for _ in range(warmup):
        e, f = step(f)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        e, f = step(f)
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ke = 0.5 * 183.84 / 9.648533e-3 * float((v * v).sum())
    drift = float(e) + ke - e0
    ev.close()
    return {"workload": "bulk bcc W 25x40x50 cells (100000 atoms), velocity Verlet, model_2and3 (2+3-body)",
            "scaling": "strong", "n_gpus": world, "ms_per_step": ms, "atom_steps_per_s": n / (ms * 1e-3),
            "dt_fs": dt, "energy0_eV": e0, "energy_drift_eV": drift, "energy_drift_rel": abs(drift) / abs(e0),
            "energy_conserved": bool(abs(drift) / abs(e0) < 1e-6)}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from uf3_b200 import _native, distributed, geometry, least_squares as ls
    from uf3_b200.engine import Engine
    from uf3_b200.pipeline import NativePipeline

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # Host waits spin by default.  Sleeping waits (uf3b_set_blocking_sync, UF3B_BLOCKING_SYNC=1) are for hosts
    # where ranks x threads outnumber the cores; on the 32-core 8-GPU box they cost their wake-up latency on
    # every wait: e2e 201 M (sleeping) against 255 M atom-steps/s (spinning), resident arm 0.42 against 0.26 ms
    # per step (profiles/README.md).
    blocking_e2e = os.environ.get("UF3B_BLOCKING_SYNC") == "1"

    def blocking(enabled):
        _native.check(_native.lib().uf3b_set_blocking_sync(1 if (enabled and blocking_e2e) else 0))

    basis = make_basis(args.basis)
    eng = Engine(basis, device=local_rank)
    F = eng.n_feats
    frames = [frame(rank * N_POOL + i) for i in range(N_POOL)]
    n_atoms = len(frames[0][0])
    images = geometry.image_table(frames[0][2], frames[0][3], basis.r_cut)
    stream = torch.cuda.current_stream().cuda_stream

    d_pos = [torch.from_numpy(fr[0]).to(dev) for fr in frames]
    d_num = torch.from_numpy(frames[0][1]).to(dev)
    d_xe = torch.empty(F, dtype=torch.float64, device=dev)
    d_xf = torch.empty((3 * n_atoms, F), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step_resident(i):
        eng.build_neighbors_device(d_pos[i % N_POOL].data_ptr(), d_num.data_ptr(), n_atoms, images,
                                   stream)
        eng.featurize_device(d_xe.data_ptr(), d_xf.data_ptr(), F, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---------------------------------------------------------------- host-buffer arms
    h_pos_np = [torch.from_numpy(fr[0]).pin_memory().numpy() for fr in frames]
    h_num_np = torch.from_numpy(frames[0][1]).pin_memory().numpy()
    # frames in flight in the host-buffer arms: 6 slots (depth sweep on the 8-GPU box: 3 -> 270 M, 4 -> 303 M,
    # 6 -> 345 M atom-steps/s; one GPU: 4 -> 33-42 M, 6 -> 40 M, 8 -> 37 M, run-to-run spread of the 25 ms window)
    # (wide rows: four — six slots of 110 MB rows each measured 5.1 against 5.8 M with the 456-column basis)
    e2e_depth = int(os.environ.get("UF3B_E2E_DEPTH", "6" if F <= 128 else "4"))
    # synthetic targets of the fit (outside the timed region): y = rows @ c_true, E = x_e @ c_true
    model = ls.WeightedLinearModel(basis, solver="cusolver", ridge_1b=1e-10, ridge_2b=1e-10, ridge_3b=1e-10)
    free = np.zeros(F)
    free[np.asarray(model.mask)] = 1.0
    c_true = np.random.default_rng(11).normal(size=F) * 0.1 * free        # frozen columns carry no signal
    c_dev = torch.from_numpy(c_true).to(dev)
    h_y, e_target = [], []
    for i in range(N_POOL):
        step_resident(i)
        h_y.append((d_xf @ c_dev).cpu().pin_memory().numpy())
        e_target.append(float(d_xe @ c_dev))
    h_xf_check = d_xf.cpu().numpy()         # rows of the last pool frame, for the sanity check of the fit
    blocking(True)          # the pipeline's events take their wait mode when they are created
    pipe = NativePipeline(basis, depth=e2e_depth, device=local_rank)
    blocking(False)
    h_xe = [torch.empty(F, dtype=torch.float64).pin_memory().numpy() for _ in range(e2e_depth)]
    fit_info = {}

    def run_e2e_fit(steps):
        """wall-clock ms for `steps` frames per rank from host buffers to fitted coefficients on the
        host: uf3b_pipeline_submit_fit per frame, then ONE all-reduce and the cuSOLVER solve"""
        stats = ls.GramStats(F)
        barrier()
        t0 = time.perf_counter()
        pending = []
        for k in range(steps):
            xe = h_xe[k % e2e_depth]
            pending.append((pipe.submit_fit(h_pos_np[k % N_POOL], h_num_np, images, h_y[k % N_POOL], xe), xe, k))
            if len(pending) == e2e_depth:          # every frame's energy row is read on the host
                ticket, xe_done, kk = pending.pop(0)
                pipe.wait(ticket)
                stats.add_energy_row(xe_done, e_target[kk % N_POOL], n_atoms)
        for ticket, xe_done, kk in pending:
            pipe.wait(ticket)
            stats.add_energy_row(xe_done, e_target[kk % N_POOL], n_atoms)
        gram_f, ord_f, moments = pipe.export_gram()
        t1 = time.perf_counter()
        stats.gram_f += gram_f
        stats.ord_f += ord_f
        stats.moments[3:6] += moments
        distributed.all_reduce_stats(stats)
        t2 = time.perf_counter()
        model.fit_from_accumulator(stats, weight=0.5)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        # sanity of the fit: the targets were generated from c_true, so the fitted model must reproduce
        # the force targets of a frame (columns without signal — trimmed, frozen — are free to differ)
        y_fit = h_xf_check @ model.coefficients
        fit_info.update(frames_ms=(t1 - t0) * 1e3, all_reduce_ms=(t2 - t1) * 1e3, solve_ms=(t3 - t2) * 1e3,
                        frames_only_value=world * n_atoms * steps / (t1 - t0),
                        force_prediction_error_rel=float(np.linalg.norm(y_fit - h_y[N_POOL - 1])
                                                         / np.linalg.norm(h_y[N_POOL - 1])))
        return max_over_ranks((t3 - t0) * 1e3), dict(fit_info)

    # rows-to-host: the same frames with every row copied out (secondary figure)
    h_out = [(torch.empty(F, dtype=torch.float64).pin_memory().numpy(),
              torch.empty((3 * n_atoms, F), dtype=torch.float64).pin_memory().numpy()) for _ in range(e2e_depth)]

    rows_window = min(e2e_depth, 4)     # row copies in flight: more of them only queue on the copy engine (28 -> 21 M at 6)

    def run_e2e_rows(steps):
        barrier()
        t0 = time.perf_counter()
        pending = []
        checksum = 0.0
        for k in range(steps):
            xe, xf = h_out[k % e2e_depth]
            pending.append((pipe.submit(h_pos_np[k % N_POOL], h_num_np, images, xe, xf), xe, xf))
            if len(pending) == rows_window:
                ticket, xe, xf = pending.pop(0)
                pipe.wait(ticket)
                checksum += float(xe[1]) + float(xf[-1, -1])
        for ticket, xe, xf in pending:
            pipe.wait(ticket)
            checksum += float(xe[1]) + float(xf[-1, -1])
        torch.cuda.synchronize()
        return max_over_ranks((time.perf_counter() - t0) * 1e3)

    def d2h_ceiling():
        """GB/s of a plain pinned device->host copy of one frame's rows (per rank, all ranks at once)"""
        src = d_xf
        dst = torch.from_numpy(h_out[0][1])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return 10 * src.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9

    # ---------------------------------------------------------------- resident arm
    # L2: every step reads its own frame and writes its own row buffer out of a pool; the pool
    # (positions + rows) is larger than the 126 MB L2, so no step finds its data in cache and no
    # explicit flush is needed.  Three slots (engine + stream each) alternate frames, each launch
    # taking half of an SM's resources (frames_in_flight=2): two frames' kernels share every SM, so
    # one frame's tail and the next frame's list build (which ends in a host wait) fill each other's
    # gaps.  (The block-per-atom kernel of the manuscript basis measured slower that way.)
    n_slots = int(os.environ.get("UF3B_BENCH_SLOTS", "4" if args.basis == "demo" else ("1" if args.basis == "binary" else "2")))
    in_flight = int(os.environ.get("UF3B_BENCH_IN_FLIGHT", "2" if args.basis == "demo" else "1"))
    # Deferred list builds: a build that reuses its slot's cell grid returns without a host wait, the feature
    # kernels run behind it, and its status is checked when the slot's NEXT build is issued (an invalid build
    # there raises: UF3B_RETRY) and, for the last frame of every slot, right after the timed region.  The host
    # never waits inside the loop, so a descheduled rank does not stall its GPU (0.27 -> 0.25 ms per step at
    # eight ranks).
    deferred_slots = os.environ.get("UF3B_BENCH_DEFERRED", "1") == "1"
    slots = [(Engine(basis, device=local_rank, frames_in_flight=in_flight, deferred_lists=deferred_slots),
              torch.cuda.Stream(dev)) for _ in range(n_slots)]
    out_pool = [(torch.empty(F, dtype=torch.float64, device=dev),
                 torch.empty((3 * n_atoms, F), dtype=torch.float64, device=dev)) for _ in range(N_POOL)]
    pool_bytes = N_POOL * (3 * n_atoms * F * 8 + n_atoms * 24)
    while pool_bytes < 160e6:                      # small bases: pad the pool past the L2 size
        out_pool.append((torch.empty(F, dtype=torch.float64, device=dev),
                         torch.empty((3 * n_atoms, F), dtype=torch.float64, device=dev)))
        pool_bytes += 3 * n_atoms * F * 8

    def timed(steps, warmup):
        main = torch.cuda.current_stream()

        def run(count, first):
            for k in range(count):
                e, st = slots[k % n_slots]
                xe_k, xf_k = out_pool[(first + k) % len(out_pool)]
                with torch.cuda.stream(st):
                    e.build_neighbors_device(d_pos[(first + k) % N_POOL].data_ptr(), d_num.data_ptr(),
                                             n_atoms, images, st.cuda_stream)
                    e.featurize_device(xe_k.data_ptr(), xf_k.data_ptr(), F, st.cuda_stream)

        # priming (untimed, before the W warm-up steps): every slot runs two frames, so that no slot meets its
        # first list build — a checked one, with its buffer allocations — inside the timed region when W < slots
        run(2 * n_slots, 0)
        run(warmup, 0)
        barrier()
        launches0 = eng.launch_count()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(main)
        for _, st in slots:
            st.wait_stream(main)
        run(steps, warmup)
        for _, st in slots:
            main.wait_stream(st)
        stop.record(main)
        barrier()
        for e, _ in slots:          # the last frame of every slot: raises if its deferred build was invalid
            e.neighbor_count(3)
        launches = eng.launch_count() - launches0
        return max_over_ranks(start.elapsed_time(stop)), launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, launches = timed(args.steps, args.warmup)
    blocking(True)
    run_e2e_fit(max(args.warmup, 2 * e2e_depth))    # warm-up: every slot's buffers and cell grid, NCCL connections, cuSOLVER handle
    # the host-buffer arms are wall-clock windows of ~25 ms: median of three repetitions of K steps each
    e2e_reps = sorted((run_e2e_fit(args.steps) for _ in range(3)), key=lambda r: r[0])
    e2e_ms, e2e_info = e2e_reps[1]
    e2e_runs = [r[0] for r in e2e_reps]
    clocks = sampler.stop() if rank == 0 else None
    run_e2e_rows(max(args.warmup, 2 * e2e_depth))
    rows_ms = sorted(run_e2e_rows(args.steps) for _ in range(3))[1]
    blocking(False)
    copy_gbs = d2h_ceiling()

    # dominant kernels (the row kernels of Kernel B), timed alone with CUDA events on their stream
    eng.set_timing(True)
    kernel_ms = []
    for i in range(5):
        flush.zero_()
        step_resident(i)
        kernel_ms.append(eng.last_kernel_ms())
    eng.set_timing(False)
    k_ms = statistics.mean(kernel_ms[1:])
    e2, e3 = eng.neighbor_count(2), eng.neighbor_count(3)
    # algorithmic bytes of one launch (DESIGN.md "Kernel B"): positions + species, both lists
    # (start, count, entries), and the 3N x F force rows + energy row written once
    alg_bytes = n_atoms * 28 + 4 * (e2 + e3) + 8 * (n_atoms + 1) + 24 * n_atoms * F + 8 * F
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9

    value = world * n_atoms * args.steps / (total_ms * 1e-3)
    e2e_value = world * n_atoms * args.steps / (e2e_ms * 1e-3)
    rows_value = world * n_atoms * args.steps / (rows_ms * 1e-3)

    extra = None
    if args.extra:
        if world == 1:
            extra = inference_extras(torch, dev, stream)
        else:
            extra = {"md": md_extra(torch, dist, dev, rank, world)}
    if rank != 0:
        pipe.close()
        if world > 1:
            dist.destroy_process_group()
        return

    fp64_peak = eng.probe_fp64_tflops()
    fp64_flop = NCU_FP64_FLOP.get(args.basis)
    fp64_achieved = fp64_flop / (k_ms * 1e-3) / 1e12 if fp64_flop else None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu_frames = 1 if args.basis == "binary" else 4
        t_cpu = oracle_frames_per_second(basis, n_cpu_frames, 1)
        cpu = {"value": n_cpu_frames * n_atoms / t_cpu, "unit": "atom-steps/s", "cores": 1,
               "kind": "port",
               "sample": f"{n_cpu_frames} full 10k-atom frames (energy row + 3N force rows), "
                         "oracle/uf3_oracle.c, single thread"}
    tiled = args.basis == "demo"
    reduce_doubles = 2 * F * F + 2 * F + 6
    line = {
        "metric": "atom-steps/s featurized (2+3-body) on 10k-atom W",
        "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.basis, F),
        "notes": {"frames_per_rank_pool": N_POOL, "pairs_per_atom": e2 / n_atoms, "list3_per_atom": e3 / n_atoms,
                  "l2": "inputs larger than L2: every step uses its own frame and row buffer out of a pool of "
                        f"{len(out_pool)} ({pool_bytes / 1e6:.0f} MB > 126 MB L2); no explicit flush",
                  "streams": f"{n_slots} slots alternate frames (one engine + stream each), "
                             f"each launch on 1/{in_flight} of the SM resources; {2 * n_slots} untimed priming frames before "
                             "the W warm-up steps (every slot's first, checked list build and its allocations); list builds "
                             + ("verified when the slot's next build is issued (deferred status check)" if deferred_slots
                                else "verified by a host wait per frame")},
        "e2e": {"value": e2e_value, "unit": "atom-steps/s", "ms_per_step": e2e_ms / args.steps,
                "how": f"uf3b_pipeline_submit_fit ({e2e_depth} slots): pinned host positions and force targets in "
                       "every step, rows folded into the normal equations on the device, the frame's energy row "
                       "read on the host every step; the timed region ends with ONE all-reduce of the normal "
                       "equations and the cuSOLVER solve (coefficients on the host); wall clock, max over ranks, "
                       "median of three repetitions of K steps",
                "repetitions_ms": e2e_runs,
                "h2d_bytes_per_step": n_atoms * 28 + 24 * n_atoms + images[1].nbytes + images[0].size * 4,
                "d2h_bytes_per_step": 8 * F,
                "d2h_bytes_at_end": 8 * (2 * F * F + F + 3), "all_reduce_doubles": reduce_doubles,
                "all_reduce_bytes": 8 * reduce_doubles, **e2e_info},
        "e2e_rows_to_host": {"value": rows_value, "unit": "atom-steps/s", "ms_per_step": rows_ms / args.steps,
                             "how": f"uf3b_pipeline_submit ({rows_window} frames in flight): pinned host positions in, energy row + "
                                    "3N force rows copied to pinned host memory every step",
                             "d2h_bytes_per_step": (3 * n_atoms + 1) * F * 8,
                             "d2h_GBps_per_rank": (3 * n_atoms + 1) * F * 8 / (rows_ms / args.steps * 1e-3) / 1e9,
                             "plain_pinned_copy_GBps_per_rank": copy_gbs},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm",
                     "kernel": ("k_centre_legs + k_rows_nbr<3,9> + k_rows_ctr<3,9>" if tiled
                                else ("k_rows_multi2<4>" if args.basis == "binary" else "k_leg_cache + k_featurize_coop<8>")),
                     "achieved": achieved, "peak": peak,
                     "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC_BYTES.get(args.basis),
                     "traffic_source": "ncu, dram__bytes_read.sum + dram__bytes_write.sum of the kernels named "
                                       "(profiles/ncu_constants.json: demo = steady state of the frame loop, "
                                       "--replay-mode application; others = kernel replay)",
                     "kernel_ms": k_ms, "algorithmic_bytes": alg_bytes,
                     "kernel_note": "the kernels named are timed together; the path is bound by the shared-memory "
                                    "/ FP64 pipes, not by HBM (roofline_fp64, DESIGN.md)"},
        "roofline_fp64": {"bound": "fp64", "achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                          "frac": fp64_achieved / fp64_peak if fp64_achieved else None,
                          "flop_per_launch": fp64_flop,
                          "flop_source": "ncu: 2 x fused + non-fused FP64 thread instructions of the kernels named "
                                         "(profiles/ncu_constants.json)",
                          "peak_source": "uf3b_probe_fp64_tflops (DFMA chains)",
                          "binding_roof": "fp64: min(HBM time, FP64 time) of the algorithm is the FP64 one "
                                          "(SURVEY.md 8d)"},
        "cpu_baseline": cpu,
        "clocks": clocks,
        "extra": extra,
    }
    print(json.dumps(line), flush=True)
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- MD strong scaling
def run_md(args, rank, world, local_rank):
    """BASELINE.json configs[4]: velocity-Verlet loop on ONE 100 000-atom W frame, forces from
    the 2+3-body model every step, atoms split over the ranks by index range
    (uf3_b200.distributed.ShardedEvaluator: own-range neighbour rows + evaluator, one NCCL
    all-reduce of 3N+1 doubles per step).  Strong scaling: total work fixed."""
    import torch
    import torch.distributed as dist
    from uf3_b200 import bspline, composition, geometry, synthetic
    from uf3_b200.distributed import ShardedEvaluator

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    data = np.load(os.path.join(ROOT, "tests", "golden", "calc_syn_w54_model23.npz"))
    cfg = json.loads(str(data["config"]))
    knots = {}
    for key, val in cfg["kwargs"]["knots_map"].items():
        parts = tuple(key.split("-"))
        knots[parts] = np.array(val) if len(parts) == 2 else [np.array(v) for v in val]
    chem = composition.ChemicalSystem(cfg["element_list"], degree=cfg["degree"])
    lead = {int(k): v for k, v in cfg["kwargs"]["leading_trim"].items()}
    trail = {int(k): v for k, v in cfg["kwargs"]["trailing_trim"].items()}
    basis = bspline.BSplineBasis(chem, knots_map=knots, leading_trim=lead, trailing_trim=trail)
    pos, numbers, cell, pbc = synthetic.bcc_w((25, 40, 50), a=3.206, sigma=0.15, seed=0)
    n = len(pos)
    images = geometry.image_table(cell, pbc, basis.r_cut)
    ev = ShardedEvaluator(basis, np.array(data["coefficients"]), device=local_rank)
    x = torch.from_numpy(pos).to(dev)
    z = torch.from_numpy(numbers).to(dev)
    v = torch.zeros_like(x)
    dt, inv_m = args.dt, 9.648533e-3 / 183.84      # fs; (eV/A)/amu -> A/fs^2

    def step(f):
        v.add_(f, alpha=0.5 * dt * inv_m)
        x.add_(v, alpha=dt)
        e, f_new = ev.energy_forces(x, z, images)
        v.add_(f_new, alpha=0.5 * dt * inv_m)
        return e, f_new

    e0, f = ev.energy_forces(x, z, images)
    e0 = float(e0)              # the evaluator reuses its output buffer: keep the VALUE, not a view
    f = f.clone()
    for _ in range(args.warmup):
        e, f = step(f)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = ev.engine.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        e, f = step(f)
    stop.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = start.elapsed_time(stop)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = ev.engine.launch_count() - launches0
    if rank == 0:
        ke = 0.5 * 183.84 / 9.648533e-3 * float((v * v).sum())
        print(json.dumps({
            "metric": "atom-steps/s, MD loop (neighbour lists + energy + forces per step) on 100k-atom W",
            "value": n * args.steps / (ms * 1e-3), "unit": "atom-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "bulk bcc W 25x40x50 cells (100000 atoms), a=3.206 A, sigma=0.15 A, "
                                   f"velocity Verlet dt={args.dt} fs, model_2and3 (2+3-body)",
                       "partition": "atom ranges, replicated positions, one all-reduce of 3N+1 doubles per step"},
            "gpu_launches": launches,
            "energy_drift_eV": float(e) + ke - e0, "energy0_eV": e0,
            "energy_drift_rel": abs(float(e) + ke - e0) / abs(e0)}), flush=True)
    ev.close()
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- fit at scale
def run_fit(args, rank, world, local_rank):
    """BASELINE.json configs[3]: stream 10 000-atom W frames through neighbour lists -> feature
    rows (left in HBM) -> normal-equation accumulation (csrc/gram.cu) on every rank, then ONE
    all-reduce of 2F^2 + 2F + 6 doubles and the regularised solve on cuSOLVER.  A step is one
    frame per rank; targets are synthetic (rows @ c_true + noise, prepared before timing)."""
    import torch
    import torch.distributed as dist
    from uf3_b200 import distributed, geometry, least_squares as ls
    from uf3_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    basis = make_basis(args.basis)
    eng = Engine(basis, device=local_rank)
    F = eng.n_feats
    frames = [frame(rank * N_POOL + i) for i in range(N_POOL)]
    n_atoms = len(frames[0][0])
    images = geometry.image_table(frames[0][2], frames[0][3], basis.r_cut)
    stream = torch.cuda.current_stream().cuda_stream
    d_pos = [torch.from_numpy(fr[0]).to(dev) for fr in frames]
    d_num = torch.from_numpy(frames[0][1]).to(dev)
    rows = torch.empty((3 * n_atoms, F), dtype=torch.float64, device=dev)
    c_true = torch.from_numpy(np.random.default_rng(11).normal(size=F) * 0.1).to(dev)
    n_total = args.steps + args.warmup
    xe_all = torch.zeros((n_total, F), dtype=torch.float64, device=dev)
    y_pool, y_mom = [], []
    for i in range(N_POOL):                       # synthetic targets, outside the timed region
        eng.build_neighbors_device(d_pos[i].data_ptr(), d_num.data_ptr(), n_atoms, images, stream)
        eng.featurize_device(xe_all[0].data_ptr(), rows.data_ptr(), F, stream)
        gen = torch.Generator(device=dev).manual_seed(100 * rank + i)
        y = rows @ c_true + 1e-3 * torch.randn(3 * n_atoms, dtype=torch.float64, device=dev, generator=gen)
        y_pool.append(y)
        y_mom.append((float(y.sum()), float((y * y).sum())))
    acc = ls.GramAccumulator(F)

    def step(k):
        i = k % N_POOL
        eng.build_neighbors_device(d_pos[i].data_ptr(), d_num.data_ptr(), n_atoms, images, stream)
        eng.featurize_device(xe_all[k].data_ptr(), rows.data_ptr(), F, stream)
        acc.add_force_rows_device(rows.data_ptr(), y_pool[i].data_ptr(), 3 * n_atoms, F, stream,
                                  y_moments=y_mom[i])

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = eng.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for k in range(args.warmup, n_total):
        step(k)
    stop.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = start.elapsed_time(stop)
    launches = eng.launch_count() - launches0
    # energy rows: one small copy, folded on the host (targets = x_e @ c_true)
    xe_host = xe_all.cpu().numpy()
    c_host = c_true.cpu().numpy()
    for k in range(n_total):
        acc.add_energy_row(xe_host[k], float(xe_host[k] @ c_host), n_atoms)
    t0 = time.perf_counter()
    distributed.all_reduce_stats(acc)
    model = ls.WeightedLinearModel(basis, solver="cusolver", ridge_1b=1e-10, ridge_2b=1e-10, ridge_3b=1e-10)
    model.fit_from_accumulator(acc, weight=0.5)
    torch.cuda.synchronize()
    tail_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms, tail_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, tail_ms = float(t[0]), float(t[1])
    if rank == 0:
        # sanity of the device solve: residual of the regularised normal equations it solved,
        # rebuilt on the host from the all-reduced statistics
        stats = acc.export()
        w_e, w_f = ls.calc_E_F_weights(stats["n_e"], stats["n_f"], stats["std_e"], stats["std_f"])
        gram, ordinate = model.combine_weighted_gram(stats["gram_e"], stats["gram_f"], stats["ord_e"],
                                                     stats["ord_f"], w_e, w_f, 0.5)
        mask = model.mask
        reg = ls.freeze_regularizer(model.regularizer, mask)
        lhs = gram[np.ix_(mask, mask)] + reg.T @ reg
        res = lhs @ model.coefficients[mask] - ordinate[mask]
        err = float(np.linalg.norm(res) / np.linalg.norm(ordinate[mask]))
        print(json.dumps({
            "metric": "atom-steps/s, fit pipeline (neighbour lists + feature rows + normal equations) on 10k-atom W",
            "value": world * n_atoms * args.steps / (ms * 1e-3), "unit": "atom-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD + " + Gram accumulation", "basis": args.basis, "n_feats": F,
                       "frames_total": world * n_total},
            "gpu_launches": launches,
            "all_reduce_and_solve_ms": tail_ms, "all_reduce_doubles": 2 * F * F + 2 * F + 6,
            "normal_equation_residual_rel": err}), flush=True)
    acc.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--basis", default="demo", choices=["demo", "manuscript", "binary"],
                    help="demo / manuscript: the two W bases of BASELINE.json configs[1]; binary: the reference's Fe-C "
                         "test basis (six trios) on a B2 lattice — the several-species kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle timing (profiling runs)")
    ap.add_argument("--extra", dest="extra", action="store_true", default=True,
                    help="also time the inference configs: Ne/Xe 50k and W 100k at one GPU, the 100k-atom MD "
                         "step over all ranks under torchrun (default)")
    ap.add_argument("--no-extra", dest="extra", action="store_false")
    ap.add_argument("--dt", type=float, default=1.0, help="MD time step in fs (--workload md)")
    ap.add_argument("--workload", default="featurize", choices=["featurize", "md", "fit"],
                    help="featurize = BASELINE.json headline (default); md = configs[4] MD loop, strong scaling; "
                         "fit = configs[3] featurize + normal equations + all-reduce + solve")
    args = ap.parse_args()
    global BASIS_KIND
    BASIS_KIND = args.basis
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "md":
        run_md(args, rank, world, local_rank)
    elif args.workload == "fit":
        run_fit(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
