"""N-rank NCCL check (torchrun; 24 particles: N = 2, 4 or 8): particle-sharded fused iteration == single-GPU iteration on the same
global batch.  Rank 0 prints PASS/FAIL."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["PMB_NO_PBAR"] = "1"
import torch, torch.distributed as dist
import golden_util as gu
import prob_mbrl_b200 as pm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
MM = os.environ.get("DIST_MM", "0") == "1"      # moment matching of states and rewards across the ranks
ops, g = gu.load("cartpole_200x2_n25_h40" if MM else "dcartpole_48x3_n24_h30")

def run(distributed):
    dyn, pol = gu.modules_from_ops(ops, dev)
    opt = torch.optim.Adam(pol.parameters(), 1e-3)
    torch.manual_seed(17)
    losses = []
    if not distributed:
        # hide the process group from mc_pilco: run the full batch locally
        saved = pm.dist.world
        pm.dist.world = lambda: (0, 1)
        sys.modules["prob_mbrl_b200.mc_pilco"].dist.world = pm.dist.world
    try:
        x0 = g["x0"][:24].to(dev)
        pm.mc_pilco(x0, dyn, pol, 8, opt, None, 4, pegasus=True, maximize=True, clip_grad=1.0,
                    resampling_period=3, init_state_noise=0.01, mm_states=MM, mm_rewards=MM,
                    on_iteration=lambda i, loss, *a: losses.append(float(loss)))
    finally:
        if not distributed:
            pm.dist.world = saved
    return torch.cat([p.detach().flatten() for p in pol.parameters()]), losses

import datetime
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
p_single, l_single = run(False)
p_shard, l_shard = run(True)
err = float((p_single - p_shard).abs().max())
pkeys = [k for k in ("pol_W0", "pol_b0", "pol_W1", "pol_b1", "pol_W2", "pol_b2", "pol_W3", "pol_b3") if k in ops]
moved = float((p_single - torch.cat([ops[k].flatten() for k in pkeys]).to(dev)).abs().max())
t = torch.tensor([err], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
# every rank ends with the same parameters: bitwise with the peer-memory exchange (slots added in rank order everywhere)
lo, hi = p_shard.clone(), p_shard.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
spread = float((hi - lo).abs().max())
sync = os.environ.get("PMB_GRAD_SYNC", "peer")
if rank == 0:
    # with moment matching the statistics are summed in another grouping (per-rank tiles): fp32-rounding-level differences
    ptol, ltol = (5e-6, 1e-5) if MM else (5e-7, 1e-6)
    ok = float(t) < ptol and moved > 1e-4 and max(abs(a - b) for a, b in zip(l_single, l_shard)) < ltol
    ok = ok and (spread == 0.0 if sync == "peer" else spread < 1e-7)
    print("DIST", "PASS" if ok else "FAIL", "mm=%d sync=%s max|dparam| %.2e  moved %.2e  rank spread %.1e" % (int(MM), sync, float(t), moved, spread),
          l_single, l_shard)
dist.destroy_process_group()
