"""world_size-2 run of the frame-sharded fit on CPU (gloo): every rank folds its share of
the rows into Gram statistics, one all-reduce, identical coefficients everywhere and
equal to the single-process fit."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import golden_util as gu
from uf3_b200 import distributed, least_squares as ls

CASES = ["syn_w16_demo", "syn_w54_demo", "syn_w128_demo", "syn_w36_slab"]


def _frames():
    rng = np.random.default_rng(5)
    frames = []
    coeff = None
    for name in CASES:
        case = gu.Case(name)
        if coeff is None:
            coeff = rng.normal(size=case["x_energy"].shape[0])
        n = len(case.numbers)
        frames.append((n, case["x_energy"], float(case["x_energy"] @ coeff + rng.normal(0, 1e-2)),
                       case["x_forces"], case["x_forces"] @ coeff + rng.normal(0, 1e-2, 3 * n)))
    return frames


def _accumulate(frames, n_feats):
    stats = ls.GramStats(n_feats)
    for n, xe, e, xf, f in frames:
        stats.add_energy_row(xe, e, n)
        stats.add_force_rows(xf, f)
    return stats


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    basis = gu.Case(CASES[0]).basis()
    frames = _frames()
    mine = distributed.shard(frames, rank, world)
    assert len(mine) == len(frames) // world
    stats = _accumulate(mine, basis.n_feats)
    distributed.all_reduce_stats(stats)
    model = ls.WeightedLinearModel(basis, ridge_1b=1e-4, ridge_2b=1e-4, ridge_3b=1e-4, curvature_2b=1e-4)
    model.fit_from_accumulator(stats, weight=0.5)
    np.save(os.path.join(out_dir, f"coeff_{rank}.npy"), model.coefficients)
    dist.destroy_process_group()


def test_two_rank_fit_matches_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    basis = gu.Case(CASES[0]).basis()
    single = ls.WeightedLinearModel(basis, ridge_1b=1e-4, ridge_2b=1e-4, ridge_3b=1e-4, curvature_2b=1e-4)
    single.fit_from_accumulator(_accumulate(_frames(), basis.n_feats), weight=0.5)
    c0 = np.load(tmp_path / "coeff_0.npy")
    c1 = np.load(tmp_path / "coeff_1.npy")
    assert np.array_equal(c0, c1)
    assert np.allclose(c0, single.coefficients, rtol=1e-6, atol=1e-7)


def test_shard_is_a_partition():
    items = list(range(11))
    parts = [distributed.shard(items, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == items
    assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _range_worker(rank, world, port, out_dir):
    """Atom-range partition of one frame (MD strong scaling): every rank contributes a partial
    [forces, energy] buffer, one all-reduce completes it.  The per-rank partials are stood in
    by the oracle's full result restricted to the rank's atom range (forces) and an equal
    share of the energy — on the GPU they come from uf3b_neighbors_build_range +
    uf3b_energy_forces (tests/test_gpu_api.py::test_centre_ranges_sum_to_the_full_frame)."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = gu.Case("calc_syn_w54_model23")
    n = len(case.numbers)
    first, count = distributed.atom_range(n, rank, world)
    buf = torch.zeros(3 * n + 1, dtype=torch.float64)
    buf[3 * first:3 * (first + count)] = torch.from_numpy(np.asarray(case["forces"]).reshape(-1)[3 * first:3 * (first + count)])
    buf[3 * n] = float(case["energy"]) / world
    distributed.all_reduce_partials(buf)
    np.save(os.path.join(out_dir, f"partial_{rank}.npy"), buf.numpy())
    dist.destroy_process_group()


def test_atom_ranges_cover_and_reduce(tmp_path):
    for n, world in ((10, 3), (100000, 8), (5, 8), (0, 2)):
        ranges = [distributed.atom_range(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and sum(c for _, c in ranges) == n
        assert all(ranges[r][0] + ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
        assert max(c for _, c in ranges) - min(c for _, c in ranges) <= 1
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_range_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    case = gu.Case("calc_syn_w54_model23")
    want = np.concatenate([np.asarray(case["forces"]).reshape(-1), [float(case["energy"])]])
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), f"partial_{rank}.npy"))
        assert np.allclose(got, want, rtol=1e-14, atol=0)
