"""N > 1 host logic on CPU (gloo, world_size 2): block sharding, the single all-reduce of the totals and of
the per-block Gram pieces, leave-one-out on the owning rank -- must reproduce the one-rank result.
The per-block numbers come from the fp64 numpy model of the device algorithm (tests/device_model.py);
the sharding / exchange functions are the product's (pyrhe_b200.engine)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank_main(rank, world, port, name, out_dir, weights=None):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import oracle_problem
    from device_model import run_model
    from test_algebra_model import plan_for
    from pyrhe_b200.engine import allreduce_sum, shard_blocks
    from pyrhe_b200.hostmath import host_terms
    p = oracle_problem(name)
    plan = plan_for(p)
    ht, Y_res = host_terms(plan, p.Z, p.W, p.y, p.env)
    full = run_model(p.packed, p.n_indv_original, p.annot, p.Z, Y_res, p.W, p.env, p.missing_indv, p.num_jack,
                     p.impute, p.seed, plan)
    J, E, E_reg = p.num_jack, plan.E, plan.E_reg
    j0, j1 = shard_blocks(J, world, rank, weights)
    # rank-local partials: only the own blocks are "computed"
    P_own = torch.from_numpy(full["P"][j0:j1].copy())
    S = P_own.sum(dim=0)
    G_blk = torch.zeros((J,) + full["G_blk"].shape[1:], dtype=torch.float64)
    G_blk[j0:j1] = torch.from_numpy(full["G_blk"][j0:j1])
    allreduce_sum([S, G_blk])
    if plan.has_nxe:
        S[E_reg] = torch.from_numpy(full["S"][E_reg])
    XX = torch.zeros((J + 1, E, E), dtype=torch.float64)
    for jl, j in enumerate(range(j0, j1)):
        L = (S - P_own[jl]).reshape(E, -1)
        XX[j] = L @ L.T
    if rank == world - 1:
        Sf = S.reshape(E, -1)
        XX[J] = Sf @ Sf.T
    allreduce_sum([XX])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), XX=XX.numpy(), G_blk=G_blk.numpy(), XX_ref=full["XX"],
             G_ref=full["G_blk"])
    dist.destroy_process_group()


@pytest.mark.parametrize("name,weights", [("rhe_cov_binary", None), ("genie_full_cov", None),
                                          ("rhe_cov_binary", [1.0, 3.0])])       # unequal shares (rate-weighted shards)
def test_two_rank_sharding_reproduces_single_rank(name, weights, tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(2, port, name, str(tmp_path), weights), nprocs=2, join=True)
    for r in range(2):
        d = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        np.testing.assert_allclose(d["XX"], d["XX_ref"], rtol=1e-12, atol=1e-12 * np.abs(d["XX_ref"]).max())
        np.testing.assert_allclose(d["G_blk"], d["G_ref"], rtol=0, atol=0)


def test_shard_blocks_matches_reference_distribution():
    from pyrhe_b200.engine import shard_blocks
    # base.py:530-533 with num_jobs = 100
    assert [shard_blocks(100, 8, r) for r in range(8)] == [(0, 13), (13, 26), (26, 39), (39, 52), (52, 65),
                                                           (65, 78), (78, 91), (91, 100)]
    assert [shard_blocks(3, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert shard_blocks(10, 1, 0) == (0, 10)


def test_weighted_shards_are_contiguous_and_proportional():
    """`shard_sizes(weights=...)`: block shares proportional to per-rank weights (the measured H2D rates of an
    upload-bound pass), every block owned exactly once, nobody idle while another rank holds several blocks."""
    from pyrhe_b200.engine import shard_blocks, shard_sizes
    rates = [23.2, 23.23, 23.29, 23.28, 35.3, 35.47, 35.49, 35.27]
    assert shard_sizes(100, 8, rates) == [10, 10, 10, 10, 15, 15, 15, 15]
    ranges = [shard_blocks(100, 8, r, rates) for r in range(8)]
    assert ranges[0][0] == 0 and ranges[-1][1] == 100
    assert all(ranges[r][1] == ranges[r + 1][0] for r in range(7))
    assert shard_sizes(8, 4, [1, 1, 1, 50]) == [1, 1, 1, 5]
    assert sum(shard_sizes(3, 4, [1, 1, 1, 5])) == 3
    assert shard_sizes(10, 1, [3.0]) == [10]
    assert shard_sizes(100, 8, [1.0] * 8) in ([13, 13, 13, 13, 12, 12, 12, 12], [12, 13, 12, 13, 12, 13, 12, 13])
    import pytest
    with pytest.raises(ValueError):
        shard_sizes(10, 2, [1.0, 0.0])


def test_weighted_e2e_leg_is_refused_when_a_share_does_not_fit():
    """bench.py only runs its rate-weighted e2e leg when every rank has room for its share next to the resident engine."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    block = 10112 * 125056 + 8 * 10 * 500224 * 4             # rows + stored partial of one config-5 block
    rates = [23.2, 23.23, 23.29, 23.28, 35.3, 35.47, 35.49, 35.27]
    assert all(bench.weighted_shares_fit(100, 8, r, rates, block, 140e9) for r in range(8))
    assert not bench.weighted_shares_fit(100, 2, 1, [45.0, 55.0], block, 40e9)
    assert not bench.weighted_shares_fit(100, 2, 0, [0.0, 55.0], block, 400e9)       # invalid weights: refused, not raised


def test_numa_cpulist_parser_and_noop_binding():
    from pyrhe_b200.util.numa import _parse_cpulist, bind_to_gpu_node
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()
    rep = bind_to_gpu_node(0)                                   # no GPU here: reports why and leaves the affinity alone
    assert rep["bound"] is False and "why" in rep
