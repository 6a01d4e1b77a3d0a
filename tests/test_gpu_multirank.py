"""GPU, two ranks over NCCL: the engine with its blocks sharded over two B200s (one process per GPU, the all-reduce of the
totals and Gram pieces over NVLink, leave-one-out on the owning rank) reproduces the one-GPU result and the golden
T, q of the unmodified reference.  Skipped on a box with a single GPU (`gpurun --gpus 2` runs it)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank_main(rank, world, port, name, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from helpers import oracle_problem
    from test_gpu_parity import make_engine, plan_for
    p = oracle_problem(name)
    eng, _, _ = make_engine(p, plan_for(p), rank=rank, world=world, device=torch.device("cuda", rank))
    out = eng.run()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), XX=out["XX"], G_blk=out["G_blk"], own=np.array(eng.own))
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["rhe_cov_binary", "dom_cov", "genie_full_cov"])
def test_two_gpus_reproduce_one_gpu_and_the_reference(name, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from helpers import load_golden, oracle_problem
    from device_model import assemble_all
    from test_gpu_parity import make_engine, plan_for
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    p = oracle_problem(name)
    plan = plan_for(p)
    eng, ht, _ = make_engine(p, plan)
    one = eng.run()
    eng.close()
    r0, r1 = (np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(2))
    assert len(r0["own"]) + len(r1["own"]) == p.num_jack and set(r0["own"]).isdisjoint(set(r1["own"]))
    np.testing.assert_array_equal(r0["XX"], r1["XX"])          # identical on all ranks after the exchange
    np.testing.assert_array_equal(r0["G_blk"], r1["G_blk"])
    # the totals are summed in a different order (per rank, then across ranks): fp32 round-off of S
    np.testing.assert_allclose(r0["XX"], one["XX"], rtol=2e-6, atol=1e-8 * np.abs(one["XX"]).max())
    np.testing.assert_allclose(r0["G_blk"], one["G_blk"], rtol=1e-12, atol=1e-12 * np.abs(one["G_blk"]).max())
    g = load_golden(name)
    T, q = assemble_all(plan, ht, dict(XX=r0["XX"], G_blk=r0["G_blk"], M=one["M"]), p.num_jack)
    np.testing.assert_allclose(T, g["T"][0], rtol=1e-5, atol=1e-6 * np.abs(g["T"][0]).max())
    np.testing.assert_allclose(q, g["q"][0], rtol=1e-5, atol=1e-6 * np.abs(g["q"][0]).max())
