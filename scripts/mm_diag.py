import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, golden_util as gu
from oracle import rollout_oracle as orc
import test_gpu_parity as T
for name, tag, groups in [("cartpole_200x2_n25_h40", "mm", None), ("cartpole_37x2_n7_h12", "mm", None),
                          ("dcartpole_48x3_n24_h30", "mm", None), ("dcartpole_48x3_n24_h30", "mmg", 2)]:
    ops, g = gu.load(name); H = int(g["H"])
    mm = dict(mm_states=True, mm_rewards=True, mm_groups=groups, z_mm=g["z_mm"], z_rr=g["z_rr"])
    r = T._run(ops, g["x0"], H, mm=mm)
    ops64, g64 = gu.load(name, torch.float64)
    r64 = orc.loss_and_grads(ops64, g64["x0"], H, mm_states=True, mm_rewards=True, z_mm=g64["z_mm"], z_rr=g64["z_rr"], mm_groups=groups)
    keys = orc.policy_param_keys(ops64); g64l = [r64["grads"][k] for k in keys]; gold = gu.policy_grad_list(g, tag, ops)
    S64 = torch.stack(r64["states"])
    print(name, tag, "status", r["status"])
    print("   states: ours-vs-fp64 %.2e  ref32-vs-fp64 %.2e" % ((r["S"].double()-S64).abs().max(), (g[tag+"_states"].double()-S64).abs().max()))
    print("   loss:   ours-vs-fp64 %.2e  ref32-vs-fp64 %.2e" % (abs(float(r["obj"])-float(r64["loss"])), abs(float(g[tag+"_loss"])-float(r64["loss"]))))
    print("   grad:   ours-vs-fp64 %.2e  ref32-vs-fp64 %.2e  ours-vs-ref32 %.2e" % (gu.rel_l2(r["grads"], g64l), gu.rel_l2(gold, g64l), gu.rel_l2(r["grads"], gold)))
    print("   dx0:    ours-vs-fp64 %.2e  ref32-vs-fp64 %.2e" % (gu.rel_l2(r["dx0"], r64["dx0"]), gu.rel_l2(g[tag+"_dx0"], r64["dx0"])))
# single-flag: states only / rewards only on N=25
for flags in (dict(mm_states=True, mm_rewards=False), dict(mm_states=False, mm_rewards=True)):
    name = "cartpole_200x2_n25_h40"
    ops, g = gu.load(name); H = int(g["H"])
    r = T._run(ops, g["x0"], H, mm=dict(flags, mm_groups=None, z_mm=g["z_mm"], z_rr=g["z_rr"]))
    ops64, g64 = gu.load(name, torch.float64)
    r64 = orc.loss_and_grads(ops64, g64["x0"], H, z_mm=g64["z_mm"], z_rr=g64["z_rr"], **flags)
    r32 = orc.loss_and_grads(ops, g["x0"], H, z_mm=g["z_mm"], z_rr=g["z_rr"], **flags)
    keys = orc.policy_param_keys(ops64); g64l = [r64["grads"][k] for k in keys]
    print(flags, "grad ours-vs-fp64 %.2e  oracle32-vs-fp64 %.2e" % (gu.rel_l2(r["grads"], g64l), gu.rel_l2([r32["grads"][k] for k in keys], g64l)))
