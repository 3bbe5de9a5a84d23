"""Drop-in replacement for ``prob_mbrl.algorithms.mc_pilco`` (reference algorithms/mc_pilco.py:13-267).

Same keyword list and the same per-iteration semantics (PEGASUS noise management, x0 perturbation,
rollout, discounted return, backward, clip, optimiser step, x0 re-sampling, RuntimeError => resample
and skip), with the rollout + back-propagation-through-time executed by the sm_100a library.

Two execution paths:
  * ``FusedIteration`` (default flag set of the deep_pilco examples: known reward, no value function,
    no CVaR, torch.optim.Adam): the whole iteration -- weight packing, forward sweep, reverse sweep,
    batched weight gradient, [NCCL all-reduce when torch.distributed is initialised], gradient clip
    and Adam -- stays on the device with static buffers and no host synchronisation other than the
    progress-bar read of the loss the reference does as well (algorithms/mc_pilco.py:215-216);
  * the generic path: fused rollout as an autograd.Function + torch autograd for the loss variants
    (value_func tail, CVaR subset, regulariser, arbitrary optimisers).
"""
import ctypes as C
import os
import weakref
from collections import defaultdict

import numpy as np
import torch
import tqdm

from . import _lib, dist, operands
from .operands import NotEligible
from .rollout import backend, fused_rollout_tensors, rollout

policy_update_counter = defaultdict(lambda: 0)   # persists across calls like the reference's (mc_pilco.py:8)
x0_tree = None          # prioritized replay: sum tree over every state seen so far (reference mc_pilco.py:9), made on first use
episode_counter = 0     # episodes of `exp` already in the tree (reference mc_pilco.py:10)
_ENGINES = {}   # fused-iteration engines survive across mc_pilco calls (the examples call it once per episode)


def tile(tensor, n):
    """Repeat every row n times contiguously (reference utils/core.py:188-190)."""
    return tensor.repeat_interleave(n, dim=0)


def _discount_fn(discount, steps):
    if discount is None:
        return lambda i: 1.0 / steps
    if callable(discount):
        return discount
    return lambda i: discount ** i


class FusedIteration:
    """Device-resident state of the fused policy-gradient iteration for one (N, H, module pair)."""

    def __init__(self, dynamics, policy, x0, H, opt, g_rewards, clip_grad, mm=None, grad_sync=None):
        self.lib = _lib.load()
        N = x0.shape[0]
        self.dynamics, self.policy, self.N, self.H = dynamics, policy, int(N), int(H)
        self.opt, self.clip = opt, (float(clip_grad) if clip_grad is not None else 0.0)
        self.mm = mm or dict(mm_states=False, mm_rewards=False, mm_groups=None, z_mm=None, z_rr=None)
        self.grad_sync = grad_sync
        self.params = [p for p in policy.parameters() if p.requires_grad]
        self.dev = self.params[0].device
        self.g_rewards = g_rewards.to(self.dev, torch.float32).contiguous()
        self.tune = _lib.make_tuning()
        self.static = {}
        self.prob = None
        f32 = dict(device=self.dev, dtype=torch.float32)
        self.x0 = x0.detach().to(**f32).clone()
        self.refresh()
        D, U = self.ops.D, self.ops.U
        self.states = torch.empty(H + 1, N, D, **f32)
        self.actions = torch.empty(H, N, U, **f32)
        self.rewards = torch.empty(H, N, **f32)
        self.weighted = torch.empty(H, N, **f32)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.nbytes = self.lib.pmb_workspace_bytes(C.byref(self.prob), C.byref(self.tune))
        self.ws = torch.empty(self.nbytes, dtype=torch.uint8, device=self.dev)
        self.nparam = int(self.lib.pmb_policy_param_count(C.byref(self.prob)))
        # flat policy gradient + the loss in one buffer: ONE all-reduce per iteration when sharded
        self.reduced = torch.zeros(self.nparam + 1, **f32)
        if self.grad_sync == "auto":
            # sharded run: the gradient all-reduce over NVLink peer memory (in-stream kernels, part of the iteration's
            # CUDA graph), or NCCL with PMB_GRAD_SYNC=nccl; a collective construction on every rank
            from . import dist as _dist
            self.grad_sync = _dist.gradient_sync(self.nparam + 1, self.dev)
        self.sync_in_graph = type(self.grad_sync).__name__ == "PeerAllReduce"
        self.grad_flat = self.reduced[:self.nparam]
        self.dx0 = torch.empty(N, D, **f32)
        self.scratch = torch.zeros(1024, **f32)
        self.loss = self.reduced[self.nparam]
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        # p.grad become views of the flat gradient (what NCCL reduces and clip+Adam consume)
        off = 0
        self.grad_views = []
        for p in self.params:
            self.grad_views.append(self.grad_flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        if off != self.nparam:
            raise RuntimeError("policy has %d trainable scalars, the library expects %d" % (off, self.nparam))
        self._init_adam()

    # -- operands ------------------------------------------------------------------------------
    def _static_copy(self, key, t):
        """Masks / noise are re-assigned (``.data =``) by the modules on resample (reference
        models/modules.py:44, densities.py:85): keep device-stable copies so captured graphs and the
        problem descriptor stay valid."""
        buf = self.static.get(key)
        if buf is None or buf.shape != t.shape:
            buf = torch.empty_like(t, memory_format=torch.contiguous_format)
            self.static[key] = buf
            self.prob = None
        buf.copy_(t)
        return buf

    @staticmethod
    def _signature(ops):
        """Everything the problem descriptor bakes in besides the engine-owned static buffers: the storage of every
        weight / bias (the modules re-bind them with ``.data =`` in Regressor.load / Policy.load, reference
        models/core.py:154-159,214-219) and the scalar fields."""
        sig = []
        for net in (ops.pol, ops.dyn):
            for w, b in zip(net.W, net.b):
                sig += [w.data_ptr(), tuple(w.shape), None if b is None else b.data_ptr()]
            sig += [tuple(float(x) for x in net.p), bool(net.has_density), float(net.lmax)]
        sig += [float(ops.rew.scale), float(ops.rew.offset), tuple(ops.rew.C.shape)]
        return tuple(sig)

    def refresh(self):
        """(Re)read the modules (after a resample(), at the start of another mc_pilco call).  Masks, noise, scalers
        and the reward operands are COPIED into engine-owned buffers, so a module that re-binds them (resample():
        models/modules.py:44; set_dataset(): models/core.py:142-149) only changes values the captured graph reads;
        the descriptor and the graph are rebuilt when a weight moved or a shape / scalar changed."""
        N = self.N
        try:
            ops = operands.extract(self.dynamics, self.policy, N)
        except NotEligible as e:
            if "buffer" not in str(e):
                raise
            operands.materialize_noise(self.dynamics, self.policy, self.x0)
            ops = operands.extract(self.dynamics, self.policy, N)
        for tag, net in (("pol", ops.pol), ("dyn", ops.dyn)):
            for i, m in enumerate(net.mask):
                if m is not None:
                    net.mask[i] = self._static_copy("%s_mask%d" % (tag, i), m[:N])
            if net.z is not None:
                net.z = self._static_copy(tag + "_z", net.z[:N])
        mm = dict(self.mm)
        rank, world = dist.world()
        self.sharded_mm = world > 1 and bool(mm.get("mm_states") or mm.get("mm_rewards"))
        for k in ("z_mm", "z_rr"):
            if mm.get(k) is not None:
                # rows 0 .. N-1 of the tables are used; across GPUs the N * world particles of all ranks are matched together
                mm[k] = self._static_copy(k, mm[k][:N * world if self.sharded_mm else N])
        for name in ("act_scale", "act_bias", "mx", "iSx", "my", "Sy"):
            setattr(ops, name, self._static_copy(name, getattr(ops, name)))
        for name in ("C", "c0", "Q", "R"):
            setattr(ops.rew, name, self._static_copy("rew_" + name, getattr(ops.rew, name)))
        sig = self._signature(ops)
        if sig != getattr(self, "_sig", None):
            self._sig = sig
            self.prob = None
        self.ops = ops
        if self.prob is None:
            self.prob, self.keep = _lib.make_problem(ops, N, self.H, shard=(rank, world), **mm)
            _lib.check_problem(self.prob, self.tune)
            self._bind_exchange()
            self.graph = None

    def _bind_exchange(self):
        """Moment matching across GPUs: the exchange areas of the per-step statistics (records + arrival counters) and of
        the whole-horizon reward exchange live in peer-mapped memory (collective allocation, once per engine)."""
        if not self.sharded_mm:
            return
        if getattr(self, "mm_exchange", None) is None:
            sizes = (C.c_size_t * 3)()
            _lib.check(self.lib.pmb_mm_exchange_bytes(C.byref(self.prob), C.byref(self.tune), sizes))
            self.mm_exchange = (dist.PeerBuffer(sizes[0], self.dev), dist.PeerBuffer(sizes[1], self.dev),
                                torch.zeros(max(int(sizes[2]) // 8, 8), dtype=torch.int64, device=self.dev))
            ok = [b.all_ok(self.dev) for b in self.mm_exchange[:2]]
            if not all(ok):
                raise NotEligible("moment matching across ranks needs peer-mapped memory between the GPUs of one node "
                                  "(%s)" % (self.mm_exchange[0].error or self.mm_exchange[1].error,))
        rec, gather, state = self.mm_exchange
        for r in range(rec.world):
            self.prob.mm_peer_rec[r] = rec.ptrs[r]
            self.prob.mm_peer_gather[r] = gather.ptrs[r]
        self.prob.mm_local_state = state.data_ptr()

    # -- optimiser -----------------------------------------------------------------------------
    def _hyper_now(self):
        g = self.opt.param_groups[0]
        return (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]))

    def _adam_sig_now(self):
        sig = []
        for p in self.params:
            st = self.opt.state[p]
            if len(st) == 0:
                return None
            sig += [p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()]
        return tuple(sig)

    def _init_adam(self):
        opt = self.opt
        self._hyper = self._hyper_now()
        self.lr, self.b1, self.b2, self.eps = self._hyper
        entries = (_lib.PmbAdamTensor * len(self.params))()
        for i, p in enumerate(self.params):
            st = opt.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            entries[i].param = p.data_ptr()
            entries[i].grad = self.grad_views[i].data_ptr()
            entries[i].exp_avg = st["exp_avg"].data_ptr()
            entries[i].exp_avg_sq = st["exp_avg_sq"].data_ptr()
            entries[i].n = p.numel()
        raw = torch.frombuffer(bytearray(bytes(entries)), dtype=torch.uint8).clone()
        self.adam_table = raw.to(self.dev)
        self.adam_step = int(float(opt.state[self.params[0]]["step"]))
        self.step_dev.fill_(self.adam_step)
        self._adam_sig = self._adam_sig_now()
        self.graph = None

    @staticmethod
    def adam_is_plain(opt, params):
        if type(opt) is not torch.optim.Adam or len(opt.param_groups) != 1:
            return False
        g = opt.param_groups[0]
        if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
            return False
        if g.get("capturable", False) or g.get("differentiable", False) or g.get("fused", None):
            return False
        if torch.is_tensor(g["lr"]):
            return False
        return [id(p) for p in g["params"]] == [id(p) for p in params]

    # -- one iteration -------------------------------------------------------------------------
    def _enqueue_sweeps(self):
        """Forward sweep, reverse sweep + weight gradient, loss: everything before the gradient is complete."""
        lib, st = self.lib, _lib.current_stream_ptr()
        pb, tb = C.byref(self.prob), C.byref(self.tune)
        _lib.check(lib.pmb_rollout_forward(pb, tb, self.x0.data_ptr(), self.states.data_ptr(),
                                           self.actions.data_ptr(), self.rewards.data_ptr(), self.ws.data_ptr(),
                                           self.nbytes, self.status.data_ptr(), st))
        _lib.check(lib.pmb_rollout_backward(pb, tb, self.states.data_ptr(), self.actions.data_ptr(),
                                            self.rewards.data_ptr(), None, None, self.g_rewards.data_ptr(),
                                            self.grad_flat.data_ptr(), self.dx0.data_ptr(), None, self.ws.data_ptr(),
                                            self.nbytes, st))
        torch.mul(self.rewards, self.g_rewards, out=self.weighted)
        torch.sum(self.weighted.view(-1), 0, out=self.loss)

    def _enqueue_update(self):
        """Gradient clip + Adam on the (reduced) flat gradient.  With moment matching the update is predicated on
        the device status word: a failed Cholesky leaves parameters, moments and the step counter untouched, like
        the reference which raises inside rollout and skips the iteration (algorithms/mc_pilco.py:122-131)."""
        guard = self.status.data_ptr() if (self.mm.get("mm_states") or self.mm.get("mm_rewards")) else None
        _lib.check(self.lib.pmb_clip_adam_step(self.adam_table.data_ptr(), len(self.params), self.clip, self.lr,
                                               self.b1, self.b2, self.eps, 0, self.step_dev.data_ptr(),
                                               self.scratch.data_ptr(), guard, _lib.current_stream_ptr()))

    def _enqueue(self):
        self._enqueue_sweeps()
        if self.grad_sync is not None:
            self.grad_sync(self.reduced, None)
        self._enqueue_update()

    def _capture(self, fns):
        """Warm up outside capture (cudaFuncSetAttribute, lazy module load, NCCL communicator), then capture one
        CUDA graph per function of `fns`; the optimiser / parameter state is restored afterwards."""
        self._sync_adam_counter()
        saved = [t.clone() for t in self._mutable_state()]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._enqueue()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for t, v in zip(self._mutable_state(), saved):
            t.copy_(v)
        graphs = []
        for fn in fns:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            graphs.append(g)
        for t, v in zip(self._mutable_state(), saved):
            t.copy_(v)
        return graphs

    def step(self, x0):
        """Run one iteration from particles ``x0`` (device tensor [N, D]); returns the loss tensor."""
        self.x0.copy_(x0, non_blocking=True)
        if self._hyper_now() != self._hyper:
            # lr / betas / eps are by-value kernel arguments baked into the captured graph: re-read and re-capture
            # when the user (or an LR scheduler) changed them
            self._hyper = self._hyper_now()
            self.lr, self.b1, self.b2, self.eps = self._hyper
            self.graph = None
        # The whole iteration replays from ONE CUDA graph -- on one GPU, and sharded with the peer-memory gradient
        # exchange (two kernels on the stream).  With the NCCL all-reduce (PMB_GRAD_SYNC=nccl): two graphs (sweeps /
        # clip+Adam) around the collective, which stays an ordinary stream-ordered call (capturing it hung with torch
        # 2.11 / NCCL 2.28 on 2xB200).  PMB_CUDA_GRAPH=0: plain launches.
        if os.environ.get("PMB_CUDA_GRAPH", "1") != "0":
            one_graph = self.grad_sync is None or self.sync_in_graph
            if self.graph is None:
                if one_graph:
                    self.graph = self._capture([self._enqueue])
                else:
                    self.graph = self._capture([self._enqueue_sweeps, self._enqueue_update])
            self.graph[0].replay()
            if not one_graph:
                self.grad_sync(self.reduced, None)
                self.graph[1].replay()
        else:
            self._enqueue()
        self.adam_step += 1
        for p, gview in zip(self.params, self.grad_views):
            p.grad = gview
            self.opt.state[p]["step"] += 1
        return self.loss

    def undo_step_count(self):
        """The device skipped the update of the last step() (status word set): take the host-side step counters
        back as well."""
        self.adam_step -= 1
        for p in self.params:
            self.opt.state[p]["step"] -= 1

    def _mutable_state(self):
        out = [self.step_dev]
        for p in self.params:
            st = self.opt.state[p]
            out += [p.data, st["exp_avg"], st["exp_avg_sq"]]
        return out

    def _sync_adam_counter(self):
        self.step_dev.fill_(self.adam_step)

    def resync(self):
        """Re-attach to the modules / optimiser at the start of another mc_pilco call."""
        self.refresh()
        if self._adam_sig_now() != self._adam_sig:
            self._init_adam()        # optimiser state reset / reloaded, or the parameters were re-bound (Policy.load)
        self.adam_step = int(float(self.opt.state[self.params[0]]["step"]))
        self._sync_adam_counter()


def _on_device_loop_ok(value_func, cvar_eps, reg_weight, prioritized_replay, opt, policy, on_rollout, debug,
                       rollout_kwargs):
    if value_func is not None or prioritized_replay or reg_weight > 0 or debug or rollout_kwargs:
        return False
    if cvar_eps > -1.0 and cvar_eps < 1.0 and cvar_eps != 0:
        return False
    if callable(on_rollout):
        return False   # runs between rollout and backward in the reference
    params = [p for p in policy.parameters() if p.requires_grad]
    return FusedIteration.adam_is_plain(opt, params)


def mc_pilco(init_states, dynamics, policy, steps, opt=None, exp=None, opt_iters=1000, value_func=None,
             pegasus=True, mm_states=False, mm_rewards=False, mm_groups=None, maximize=True, clip_grad=1.0,
             cvar_eps=0.0, reg_weight=0.0, discount=None, on_rollout=None, on_iteration=None,
             step_idx_to_sample=None, init_state_noise=0.0, resampling_period=99, prioritized_replay=False,
             priority_alpha=0.6, priority_eps=1e-8, init_priority_beta=1.0, priority_beta_increase=0.0,
             debug=False, rollout_kwargs={}):
    """MC-PILCO policy search: ``opt_iters`` policy-gradient iterations on imagined particle rollouts."""
    global policy_update_counter, x0_tree, episode_counter
    if prioritized_replay and x0_tree is None:
        from .replay import SumTree
        x0_tree = SumTree(2 ** 20)
    dynamics.eval()
    policy.train()
    H = int(steps)
    disc = _discount_fn(discount, H)
    msg = "Pred. Cumm. rewards: %f" if maximize else "Pred. Cumm. costs: %f"
    if opt is None:
        opt = torch.optim.Adam([p for p in policy.parameters() if p.requires_grad])
    dev, dt = dynamics.X.device, dynamics.X.dtype
    D = init_states.shape[-1]
    # noise tables for moment matching: drawn on the host generator like the reference (mc_pilco.py:57-62)
    z_mm = torch.randn(H + init_states.shape[0], *init_states.shape[1:]).reshape(-1, D).to(dev, dt)
    z_rr = torch.randn(H + init_states.shape[0], 1).reshape(-1, 1).to(dev, dt)

    def resample():
        seed = torch.randint(2 ** 32, [1])
        dynamics.resample(seed=seed)
        policy.resample(seed=seed)
        if value_func is not None:
            value_func.resample(seed=seed)
        z_mm.normal_()
        z_rr.normal_()

    resample()
    x0 = init_states
    N_particles = init_states.shape[0]
    n_opt_steps = policy_update_counter[policy]
    if prioritized_replay:      # reference mc_pilco.py:79-83
        x0_idxs = None
        x0_weights = torch.ones_like(x0)
        priority_beta = init_priority_beta
    mode = backend()
    rank, world = dist.world()
    sharder = None
    if world > 1:
        if prioritized_replay:
            raise NotEligible("prioritized replay keeps one host-side sum tree; not sharded across ranks")
        if (mm_states or mm_rewards) and (backend() == "eager" or not pegasus or mm_groups):
            raise NotEligible("moment matching across ranks runs on the fused engine (pegasus=True, one matching group): "
                              "the per-step statistics are exchanged over peer memory inside the sweeps")
        sharder = dist.ShardedNoise(dynamics, policy, N_particles, rank, world)
    fast = (mode != "eager" and dev.type == "cuda"
            and _on_device_loop_ok(value_func, cvar_eps, reg_weight, prioritized_replay, opt, policy,
                                   on_rollout, debug, rollout_kwargs))
    if world > 1 and (mm_states or mm_rewards) and not fast:
        raise NotEligible("moment matching across ranks runs on the fused engine only (CUDA modules, plain Adam, no "
                          "per-iteration python hooks)")
    engine = None
    readback, pending = None, None
    pbar = tqdm.tqdm(range(opt_iters), total=opt_iters, disable=os.environ.get("PMB_NO_PBAR") == "1")
    pbar_every = max(1, int(os.environ.get("PMB_PBAR_EVERY", "1")))
    sign = -1.0 if maximize else 1.0

    for i in pbar:
        need_resample = (not pegasus) or n_opt_steps % resampling_period == 0
        if need_resample:
            if sharder is not None:
                sharder.widen()
            resample()
        x0_ = x0
        if mm_groups is not None and x0_.shape[0] == mm_groups:
            x0_ = tile(x0_, int(N_particles / mm_groups))
        # non_blocking: a pinned host batch (exp.sample_states(...)) is uploaded without stalling the host on the
        # previous iteration still running on the stream
        x0_ = x0_.to(dev, dt, non_blocking=True)
        x0_ = x0_ + init_state_noise * torch.randn_like(x0_)
        if sharder is not None:
            # every rank holds the full-N noise (identical seeds); take this rank's rows of everything
            if sharder.full is None:
                try:
                    operands.extract(dynamics, policy, x0_.shape[0])
                except NotEligible as e:
                    if "buffer" not in str(e):
                        raise
                    operands.materialize_noise(dynamics, policy, x0_.detach())
                sharder.narrow()
            x0_ = x0_[sharder.row0:sharder.row0 + sharder.n]
        Nloc = x0_.shape[0]
        # Only a NUMERICAL failure of the rollout is caught (the reference wraps just utils.rollout,
        # algorithms/mc_pilco.py:101-131): non-PD particle covariance -> resample all random numbers and skip the
        # iteration without an update.  Library / CUDA errors and ineligible configurations propagate.
        failed = None
        if fast and pegasus:
            if engine is None or engine.N != Nloc:
                weights = tuple(sign * disc(t) for t in range(H))
                key = (id(policy), id(dynamics), id(opt), Nloc, H, clip_grad, weights, world,
                       bool(mm_states), bool(mm_rewards), mm_groups)
                hit = _ENGINES.get(key)
                if hit is not None and hit[0]() is policy and hit[1]() is opt:
                    engine = hit[2]
                    engine.mm.update(z_mm=z_mm if mm_states else None, z_rr=z_rr if mm_rewards else None)
                    engine.x0.copy_(x0_)
                    engine.resync()
                else:
                    g_r = torch.tensor(weights, dtype=torch.float32, device=dev)
                    g_r = (g_r / (Nloc * world))[:, None].expand(H, Nloc).contiguous()
                    sync = "auto" if world > 1 else None
                    engine = FusedIteration(dynamics, policy, x0_, H, opt, g_r, clip_grad,
                                            dict(mm_states=mm_states, mm_rewards=mm_rewards, mm_groups=mm_groups,
                                                 z_mm=z_mm if mm_states else None,
                                                 z_rr=z_rr if mm_rewards else None), sync)
                    if len(_ENGINES) > 8:
                        _ENGINES.clear()
                    _ENGINES[key] = (weakref.ref(policy), weakref.ref(opt), engine)
            elif need_resample:
                engine.refresh()
            loss = engine.step(x0_)
            S, A, R = engine.states, engine.actions, engine.rewards
            if mm_states or mm_rewards:
                bad = int(engine.status.item())
                if bad:
                    # the device predicated clip + Adam on the status word: nothing was updated
                    engine.undo_step_count()
                    failed = "moment matching: covariance not positive-definite at step %d" % (bad - 1)
        else:
            policy.zero_grad()
            dynamics.zero_grad()
            opt.zero_grad()
            lists = None
            want_prio = prioritized_replay and x0_idxs is not None
            extras = {"want_action_grads": True} if want_prio else None
            try:
                if mode == "eager":
                    lists = rollout(x0_, dynamics, policy, H, resample_state_noise=not pegasus,
                                    resample_action_noise=not pegasus, mm_states=mm_states, mm_rewards=mm_rewards,
                                    z_mm=z_mm if pegasus else None, z_rr=z_rr if pegasus else None,
                                    mm_groups=mm_groups, **rollout_kwargs)
                    S, A = torch.stack(lists[0]), torch.stack(lists[1])
                    R = torch.stack(lists[2]).squeeze(-1)
                else:
                    S, A, R, status = fused_rollout_tensors(
                        x0_, dynamics, policy, H, mm_states, mm_rewards, z_mm if pegasus else None,
                        z_rr if pegasus else None, mm_groups, not pegasus, not pegasus, extras)
                    if (mm_states or mm_rewards) and int(status.item()):
                        failed = ("moment matching: covariance not positive-definite at step %d"
                                  % (int(status.item()) - 1))
            except RuntimeError as e:
                if isinstance(e, (_lib.LibraryMissing, _lib.LibraryError)):
                    raise
                import traceback
                traceback.print_exc()
                failed = str(e)
            if failed is None:
                if callable(on_rollout):
                    lists = lists or _as_lists(S, A, R)
                    on_rollout(i, lists[0], lists[1], lists[2], disc)
                w = torch.tensor([disc(t) for t in range(R.shape[0])], dtype=R.dtype, device=R.device)
                total = (R * w[:, None]).sum(0)
                if value_func is not None:
                    Vend = value_func(S[-1], resample=False, return_samples=True)
                    total = total + disc(H) * Vend.reshape(-1)
                returns = -total if maximize else total
                if cvar_eps > -1.0 and cvar_eps < 1.0 and cvar_eps != 0:
                    rd = returns.detach().cpu().numpy()
                    if cvar_eps > 0:
                        returns = returns[torch.as_tensor(rd < np.quantile(rd, cvar_eps), device=returns.device)]
                    else:
                        returns = returns[torch.as_tensor(rd > np.quantile(rd, -cvar_eps), device=returns.device)]
                step_norms = []
                if want_prio:
                    # importance-sampling weights (same broadcast as the reference, mc_pilco.py:158-160)
                    returns = returns.reshape(-1, 1) * x0_weights
                    if lists is not None:       # module loop: hooks on the per-step actions like the reference (:184-188)
                        for a_t in lists[1]:
                            a_t.register_hook(lambda g: step_norms.append(g.norm(dim=-1)))
                loss = returns.mean()
                if reg_weight > 0:
                    loss = loss + reg_weight * policy.regularization_loss()
                loss.backward()
                if want_prio:
                    # score every initial state by the mean norm of dL/da_t over its particles' imagined steps and
                    # refresh its priority (reference mc_pilco.py:165-182); the fused reverse sweep exports the
                    # total dL/da_t of every step in one tensor
                    m_norms = torch.stack(step_norms) if lists is not None else extras["action_grads"].norm(dim=-1)
                    if mm_groups is not None:
                        m_norms = m_norms.view(-1, mm_groups, int(N_particles / mm_groups)).mean(-1)
                    scores = m_norms.mean(0).detach().cpu().numpy() / x0_tree.counts[x0_idxs - x0_tree.max_size + 1]
                    for node, prio in zip(x0_idxs, (scores + priority_eps) ** priority_alpha):
                        x0_tree.update(node, prio)
                    x0_tree.renormalize()
                if world > 1:
                    for p in policy.parameters():
                        if p.grad is not None:
                            torch.distributed.all_reduce(p.grad)
                            p.grad /= world
                if clip_grad is not None:
                    torch.nn.utils.clip_grad_norm_(policy.parameters(), clip_grad)
                opt.step()
        if failed is not None:
            print("RuntimeError: %s" % failed)
            if sharder is not None:
                sharder.widen()
            resample()
            if sharder is not None:
                sharder.narrow()
            if engine is not None:
                engine.refresh()
            policy.zero_grad()
            dynamics.zero_grad()
            opt.zero_grad()
            continue
        n_opt_steps += 1
        if (i + 1) % pbar_every == 0 or i + 1 == opt_iters:
            if R.is_cuda:
                # progress line without a pipeline bubble: the predicted return goes to pinned host memory
                # asynchronously and is shown once the NEXT iteration is already queued (the loop end flushes)
                if readback is None:
                    readback = [(torch.empty((), dtype=torch.float32, pin_memory=True), torch.cuda.Event())
                                for _ in range(2)]
                buf, ev = readback[i & 1]
                buf.copy_(R.detach().sum(0).mean(), non_blocking=True)
                ev.record()
                if pending is not None:
                    pending[1].synchronize()
                    pbar.set_description((msg % float(pending[0])) + " [{0}]".format(R.shape[0]))
                pending = (buf, ev)
            else:
                pbar.set_description((msg % float(R.detach().sum(0).mean())) + " [{0}]".format(R.shape[0]))
        if callable(on_iteration):
            lists = _as_lists(S, A, R)
            on_iteration(i, loss, lists[0], lists[1], lists[2], disc)
        if exp is not None and prioritized_replay:
            # reference mc_pilco.py:223-246: new episodes enter the tree with the largest priority seen so far, then
            # the next batch of initial states is drawn from it together with its importance weights
            if exp.n_samples() > x0_tree.size:
                for ep in range(episode_counter, exp.n_episodes()):
                    for x in torch.tensor(exp.states[ep]):
                        x0_tree.append(x, x0_tree.max_p)
                        x0_tree.renormalize()
                episode_counter = exp.n_episodes()
            nsamp = mm_groups if mm_groups is not None else N_particles
            picked, x0_idxs, w_is = x0_tree.sample(nsamp, beta=priority_beta)
            priority_beta = max(1.0, priority_beta + priority_beta_increase)
            x0 = torch.stack(list(picked)).to(dev, dt)
            x0_weights = torch.tensor(np.stack(w_is)).to(x0.device, x0.dtype)
        elif exp is not None:
            nsamp = mm_groups if mm_groups is not None else N_particles
            x0 = exp.sample_states(nsamp, timestep=step_idx_to_sample).to(dev, dt, non_blocking=True)
            init_states = x0
        else:
            x0 = init_states.detach()

    if pending is not None:
        pending[1].synchronize()
        pbar.set_description((msg % float(pending[0])) + " [{0}]".format(H))
    if sharder is not None:
        sharder.widen()
    policy.eval()
    dynamics.eval()
    policy_update_counter[policy] = n_opt_steps


def _as_lists(S, A, R):
    return [list(S.unbind(0)), list(A.unbind(0)), list(R.unsqueeze(-1).unbind(0))]
