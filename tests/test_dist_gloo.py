"""Particle sharding across ranks (the N>1 path) on CPU with the gloo backend, world_size 2:
every rank draws the full-N noise from identically seeded generators, keeps its own rows, and the only
collective is the all-reduce of the policy gradient -- the result must equal the single-process run."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def _run(world_size, rank, port, out):
    os.environ["PROB_MBRL_BACKEND"] = "eager"
    os.environ["PMB_NO_PBAR"] = "1"
    torch.set_num_threads(1)
    import golden_util as gu
    import prob_mbrl_b200 as pm
    if world_size > 1:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world_size)
    ops, g = gu.load("dcartpole_48x3_n24_h30")
    dyn, pol = gu.modules_from_ops(ops)
    opt = torch.optim.Adam(pol.parameters(), 1e-3)
    torch.manual_seed(17)
    losses = []
    pm.mc_pilco(g["x0"], dyn, pol, 8, opt, None, 3, pegasus=True, maximize=True, clip_grad=1.0,
                resampling_period=2, init_state_noise=0.01,
                on_iteration=lambda i, loss, *a: losses.append(float(loss)))
    params = torch.cat([p.detach().flatten() for p in pol.parameters()])
    # the modules hold the full-N buffers again afterwards
    assert pol.model.drop0.noise.shape[0] == 24 and dyn.output_density.z.shape[0] == 24
    if world_size > 1:
        gathered = [torch.zeros_like(params) for _ in range(world_size)]
        dist.all_gather(gathered, params)
        assert torch.equal(gathered[0], gathered[1])          # identical update on every rank
        dist.destroy_process_group()
    if rank == 0:
        torch.save({"params": params, "losses": losses}, out)


def _worker(rank, world_size, port, out):
    _run(world_size, rank, port, out)


def test_two_rank_sharded_mc_pilco_equals_single_process(tmp_path):
    single, double = str(tmp_path / "single.pt"), str(tmp_path / "double.pt")
    _run(1, 0, 0, single)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, double), nprocs=2, join=True)
    a, b = torch.load(single), torch.load(double)
    # global loss = mean over ranks of the local means; parameters agree to fp32 summation order
    assert (a["params"] - b["params"]).abs().max() < 2e-7
    assert len(b["losses"]) == 3


def test_sharded_noise_narrow_widen_round_trip():
    import golden_util as gu
    from prob_mbrl_b200 import dist as pd
    ops, g = gu.load("dcartpole_48x3_n24_h30")
    dyn, pol = gu.modules_from_ops(ops)
    full = pol.model.drop1.noise.clone()
    cfull = dyn.model.drop0.concrete_noise.clone()
    sh = pd.ShardedNoise(dyn, pol, 24, rank=1, world_size=2)
    sh.narrow()
    assert torch.equal(pol.model.drop1.noise, full[12:24])
    assert torch.equal(dyn.model.drop0.concrete_noise, cfull[12:24])
    assert dyn.output_density.z.shape[0] == 12 and pol.model.fc_nonlin.z.shape[0] == 12
    sh.widen()
    assert torch.equal(pol.model.drop1.noise, full) and torch.equal(dyn.model.drop0.concrete_noise, cfull)
    with pytest.raises(ValueError):
        pd.shard_rows(25, 0, 2)
