"""Drop-in replacement for ``prob_mbrl.utils.rollout`` (reference utils/rollout.py:62-163).

Same signature, same return value -- ``[states, actions, rewards]`` as python lists of per-step
tensors (``len = H+1, H, H``), every element differentiable w.r.t. the policy parameters and the
initial particles -- but the H-step loop runs as two persistent sm_100a kernels (forward sweep,
reverse sweep) plus the batched weight-gradient GEMMs, entered through the C ABI of
``include/pmb_b200.h``.

Backend selection (environment variable ``PROB_MBRL_BACKEND``, no new call-site flags so the
reference examples run unchanged):
  fused  (default)  CUDA library or an exception -- never a silent fallback;
  auto              fused when the configuration is eligible, otherwise warn
                    and run the eager module loop;
  eager             the reference algorithm through the modules' own ``forward`` (host-side logic
                    tests in the GPU-less build container).
"""
import ctypes as C
import os
import warnings

import torch

from . import _lib, operands
from .operands import NotEligible


def backend():
    b = os.environ.get("PROB_MBRL_BACKEND", "fused").lower()
    if b not in ("fused", "auto", "eager"):
        raise ValueError("PROB_MBRL_BACKEND must be fused, auto or eager (got %r)" % b)
    return b


# ----------------------------------------------------------------------------------------------
# eager module loop (opt-in)
# ----------------------------------------------------------------------------------------------
def cyclic_rows(z, i, shape, device=None):
    """Rows (i + arange(N)) mod N of a pre-drawn noise table, or fresh noise when there is none
    (reference utils/rollout.py:53-59)."""
    if z is None:
        return torch.randn(*shape, device=device)
    n = shape[0]
    return z[(torch.arange(n, device=z.device) + i) % n]


def moment_match(x, z, jitter=1e-12):
    """Replace particles by m + zhat chol(cov)^T with zhat the per-column standardised z
    (reference utils/rollout.py:20-29)."""
    n = x.shape[-2]
    m = x.mean(-2, keepdim=True)
    dx = x - m
    cov = dx.transpose(-1, -2) @ dx / (n - 1) + jitter * torch.eye(x.shape[-1], device=x.device, dtype=x.dtype)
    L = torch.linalg.cholesky(cov)
    zh = ((z - z.mean(-2, keepdim=True)) / z.std(-2, keepdim=True)).detach()
    return m + zh @ L.transpose(-1, -2)


def _eager_rollout(states, dynamics, policy, steps, resample_model, resample_policy, resample_state_noise,
                   resample_action_noise, mm_states, mm_rewards, z_mm, z_rr, mm_groups, breaking_condition,
                   on_step, on_pol_eval):
    traj = []
    next_states = states
    for i in range(steps):
        try:
            z1 = cyclic_rows(z_mm, i, states.shape, states.device)
            z2 = cyclic_rows(z_rr, i, (states.shape[0], 1), states.device)
            actions = policy(states, resample=resample_policy, return_samples=True,
                             resample_noise=resample_action_noise)
            if callable(on_pol_eval):
                states, actions = on_pol_eval(i, states, actions)
            next_states, rewards = dynamics((states, actions), return_samples=True, separate_outputs=True,
                                            deltas=False, resample=resample_model,
                                            resample_noise=resample_state_noise)
            G = mm_groups
            if mm_states:
                D = next_states.shape[-1]
                next_states = (moment_match(next_states.view(G, -1, D), z1.view(G, -1, D)).view(-1, D)
                               if G is not None else moment_match(next_states, z1))
            if mm_rewards:
                rewards = (moment_match(rewards.view(G, -1, 1), z2.view(G, -1, 1)).view(-1, 1)
                           if G is not None else moment_match(rewards, z2))
            traj.append((states, actions, rewards))
            states = next_states
            if callable(breaking_condition) and breaking_condition(traj):
                break
            if callable(on_step):
                on_step(traj)
        except RuntimeError:
            # numerical failure (e.g. non-PD particle covariance): keep what we have if it is long enough
            if len(traj) > 5:
                break
            raise
    out = [list(x) for x in zip(*traj)]
    out[0].append(next_states)
    return out


# ----------------------------------------------------------------------------------------------
# fused path
# ----------------------------------------------------------------------------------------------
class FusedRolloutFunction(torch.autograd.Function):
    """states, actions, rewards = f(x0, *policy_parameters) through libpmb_b200."""

    @staticmethod
    def forward(ctx, x0, problem_pack, *params):
        ops, N, H, mm = problem_pack[:4]
        # optional side channel: {"want_action_grads": True} -> backward leaves the TOTAL dL/da_t [H, N, U] of every
        # step in extras["action_grads"] (what hooks on actions[t] see in the reference, mc_pilco.py:160-188)
        ctx.extras = problem_pack[4] if len(problem_pack) > 4 else None
        lib = _lib.load()
        prob, keep = _lib.make_problem(ops, N, H, **mm)
        tune = _lib.make_tuning()
        _lib.check_problem(prob, tune)
        nbytes = lib.pmb_workspace_bytes(C.byref(prob), C.byref(tune))
        ctx.set_materialize_grads(False)
        dev = x0.device
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        x0c = x0.detach().contiguous()
        states = torch.empty(H + 1, N, ops.D, device=dev, dtype=torch.float32)
        actions = torch.empty(H, N, ops.U, device=dev, dtype=torch.float32)
        rewards = torch.empty(H, N, device=dev, dtype=torch.float32)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(lib.pmb_rollout_forward(C.byref(prob), C.byref(tune), x0c.data_ptr(), states.data_ptr(),
                                           actions.data_ptr(), rewards.data_ptr(), ws.data_ptr(), nbytes,
                                           status.data_ptr(), _lib.current_stream_ptr()))
        ctx.pack = (prob, tune, keep, ws, nbytes, states, actions, rewards, x0c)
        ctx.nparam = int(lib.pmb_policy_param_count(C.byref(prob)))
        ctx.param_shapes = [p.shape for p in params]
        ctx.status = status
        ctx.mark_non_differentiable(status)
        return states, actions, rewards, status

    @staticmethod
    def backward(ctx, g_states, g_actions, g_rewards, _g_status):
        prob, tune, keep, ws, nbytes, states, actions, rewards, x0c = ctx.pack
        lib = _lib.load()

        def cot(g):
            return None if g is None else g.contiguous().float()

        g_states, g_actions, g_rewards = cot(g_states), cot(g_actions), cot(g_rewards)
        grad_flat = torch.empty(ctx.nparam, device=states.device, dtype=torch.float32)
        dx0 = torch.empty_like(x0c)
        da_total = None
        if ctx.extras is not None and ctx.extras.get("want_action_grads"):
            da_total = torch.empty_like(actions)
            ctx.extras["action_grads"] = da_total
        _lib.check(lib.pmb_rollout_backward(
            C.byref(prob), C.byref(tune), states.data_ptr(), actions.data_ptr(), rewards.data_ptr(),
            g_states.data_ptr() if g_states is not None else None,
            g_actions.data_ptr() if g_actions is not None else None,
            g_rewards.data_ptr() if g_rewards is not None else None,
            grad_flat.data_ptr(), dx0.data_ptr(), da_total.data_ptr() if da_total is not None else None,
            ws.data_ptr(), nbytes, _lib.current_stream_ptr()))
        grads, off = [], 0
        for shp in ctx.param_shapes:
            n = 1
            for s in shp:
                n *= s
            grads.append(grad_flat[off:off + n].view(shp))
            off += n
        return (dx0, None) + tuple(grads)


def fused_rollout_tensors(states, dynamics, policy, steps, mm_states=False, mm_rewards=False, z_mm=None,
                          z_rr=None, mm_groups=None, resample_state_noise=False, resample_action_noise=False,
                          extras=None):
    """Fused rollout returning stacked tensors (states [H+1,N,D], actions [H,N,U], rewards [H,N], status).
    Raises NotEligible for module graphs outside the fused scope."""
    if not (torch.is_tensor(states) and states.is_cuda):
        raise NotEligible("fused rollout needs CUDA tensors (got %s); set PROB_MBRL_BACKEND=eager for the "
                          "module loop" % (states.device if torch.is_tensor(states) else type(states)))
    if states.dim() != 2:
        raise NotEligible("states must be [N, D]")
    states = states.float()
    N, D = states.shape
    try:
        ops = operands.extract(dynamics, policy, N, D)
    except NotEligible as e:
        if "buffer" not in str(e):
            raise
        operands.materialize_noise(dynamics, policy, states.detach())
        ops = operands.extract(dynamics, policy, N, D)
    if ops.pol.W[0].device != states.device:
        raise NotEligible("policy parameters live on %s, states on %s" % (ops.pol.W[0].device, states.device))
    # fresh per-step output noise (rollout defaults; reference models/densities.py:113-116 redraws z every
    # call): pre-draw it in the reference's per-step order so the RNG stream is consumed identically
    if resample_state_noise or resample_action_noise or (mm_states and z_mm is None) or (mm_rewards and z_rr is None):
        zp, zd = [], []
        for _ in range(steps):
            if z_mm is None:
                torch.randn(N, D, device=states.device)       # stream parity with get_z_rnd
            if z_rr is None:
                torch.randn(N, 1, device=states.device)
            if ops.pol.has_density:
                zp.append(torch.randn(N, ops.U, device=states.device) if resample_action_noise else ops.pol.z[:N])
            if ops.dyn.has_density:
                zd.append(torch.randn(N, D, device=states.device) if resample_state_noise else ops.dyn.z[:N])
        if (mm_states and z_mm is None) or (mm_rewards and z_rr is None):
            raise NotEligible("moment matching without pre-drawn z_mm/z_rr is not fused")
        if ops.pol.has_density and resample_action_noise:
            ops.pol.z = torch.stack(zp)
            policy.model[-1].z.data = zp[-1]
        if ops.dyn.has_density and resample_state_noise:
            ops.dyn.z = torch.stack(zd)
            dynamics.output_density.z.data = zd[-1]
    mm = dict(mm_states=mm_states, mm_rewards=mm_rewards, mm_groups=mm_groups, z_mm=z_mm, z_rr=z_rr)
    params = ops.policy_parameters()
    return FusedRolloutFunction.apply(states, (ops, N, int(steps), mm, extras), *params)


_warned = set()


def rollout(states, dynamics, policy, steps, resample_model=False, resample_policy=False,
            resample_state_noise=True, resample_action_noise=True, mm_states=False, mm_rewards=False,
            infer_noise_variables=False, z_mm=None, z_rr=None, mm_groups=None, breaking_condition=None,
            on_step=None, on_pol_eval=None, **kwargs):
    """Trajectory distribution (s_0, a_0, r_0, s_1, ...) of ``policy`` on ``dynamics`` from ``states``.

    Signature and semantics of reference utils/rollout.py:62-79 (unknown keywords are swallowed like
    there).  Returns ``[states, actions, rewards]``, lists of per-step tensors."""
    mode = backend()
    eager_args = (states, dynamics, policy, steps, resample_model, resample_policy, resample_state_noise,
                  resample_action_noise, mm_states, mm_rewards, z_mm, z_rr, mm_groups, breaking_condition,
                  on_step, on_pol_eval)
    if mode == "eager":
        if infer_noise_variables:
            raise NotImplementedError("infer_noise_variables is outside this package's scope")
        return _eager_rollout(*eager_args)
    try:
        if resample_model or resample_policy:
            raise NotEligible("resample_model/resample_policy=True redraw the dropout masks every step")
        if infer_noise_variables:
            raise NotEligible("infer_noise_variables=True")
        if callable(breaking_condition) or callable(on_step) or callable(on_pol_eval):
            raise NotEligible("per-step python callbacks need the module loop")
        S, A, R, status = fused_rollout_tensors(states, dynamics, policy, steps, mm_states, mm_rewards, z_mm, z_rr,
                                                mm_groups, resample_state_noise, resample_action_noise)
    except NotEligible as e:
        if mode == "fused":
            raise
        key = str(e)
        if key not in _warned:
            _warned.add(key)
            warnings.warn("prob_mbrl_b200: falling back to the eager module loop: %s" % key)
        return _eager_rollout(*eager_args)
    if mm_states or mm_rewards:
        bad = int(status.item())
        if bad:
            # the reference raises from cholesky() at that step (utils/rollout.py:25,154-157)
            H_ok = bad - 1
            if H_ok > 5:
                return [list(S[:H_ok + 1].unbind(0)), list(A[:H_ok].unbind(0)),
                        list(R[:H_ok].unsqueeze(-1).unbind(0))]
            raise RuntimeError("moment matching: particle covariance is not positive-definite at step %d" % H_ok)
    return [list(S.unbind(0)), list(A.unbind(0)), list(R.unsqueeze(-1).unbind(0))]
