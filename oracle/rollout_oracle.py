"""CPU ORACLE for the imagined-rollout hot path of mcgillmrl/prob_mbrl.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this file, and
only as the checker / timed CPU baseline -- never as the thing shipped.  The product
path (``prob_mbrl_b200``) does not import it and fails loudly without its CUDA library.

It is a plain-PyTorch (CPU, fp32 or fp64) restatement of the reference algorithm on
*extracted operand tensors* (the flat dict written by
``prob_mbrl_b200.operands.RolloutOperands.to_flat`` and stored in ``tests/golden/*.npz``).
Each function cites the reference file:line it follows (paths relative to the reference
root).  The arithmetic itself lives in PyTorch (reference setup.py:15-18 leaves torch
unpinned; the build container has torch 2.11.0+cu128).

PARITY PINNING: the reference ships no tests / golden vectors / known answers for this
path, so parity is "unpinned by the reference's own tests".  We pin it
ourselves: ``tests/golden/make_golden.py`` imports the UNMODIFIED reference in the build
container, runs its ``utils.rollout`` + ``backward`` and ``algorithms.mc_pilco`` on seeded
fixtures and commits inputs+outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
holds this oracle to those files (bit-exact or <= a few ulp), and the SURVEY App. C.3
known-answer losses are re-checked there as well.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# helpers on the flat operand dict
# --------------------------------------------------------------------------- #
def as_torch(d, dtype=torch.float32):
    """Convert a loaded npz / flat dict into torch tensors of ``dtype`` (scalars stay python)."""
    out = {}
    for k, v in d.items():
        if hasattr(v, "shape") and getattr(v, "shape", ()) != ():
            out[k] = torch.as_tensor(v).to(dtype).clone()
        else:
            v = v.item() if hasattr(v, "item") else v
            out[k] = v
    return out


def policy_param_keys(d):
    keys = []
    for i in range(int(d["pol_L"]) + 1):
        keys.append("pol_W%d" % i)
        if ("pol_b%d" % i) in d:
            keys.append("pol_b%d" % i)
    return keys


def _noise(z, t, n):
    """PEGASUS noise buffers are [N, k] constant over steps (reference models/densities.py:113-119
    re-uses ``self.z``); a [H, N, k] tensor is per-step pre-drawn noise."""
    if z.dim() == 3:
        return z[t, :n]
    return z[:n]


# --------------------------------------------------------------------------- #
# one step
# --------------------------------------------------------------------------- #
def net_forward(d, tag, x, keep=None):
    """BSequential.forward over (Linear, ReLU, dropout)xL + Linear -- reference
    models/modules.py:215-232; BDropout eval/pegasus branch ``x*noise[:N]/p`` modules.py:61;
    CDropout eval branch ``x*concrete_noise[:N]`` (no /p) modules.py:160."""
    n = x.shape[0]
    L = int(d[tag + "_L"])
    for i in range(L):
        x = F.relu(F.linear(x, d["%s_W%d" % (tag, i)], d.get("%s_b%d" % (tag, i))))
        mk = "%s_mask%d" % (tag, i)
        if mk in d:
            x = x * d[mk][:n]
        p = float(d["%s_p%d" % (tag, i)])
        if p != 1.0:
            x = x / p
        if keep is not None:
            keep.append(x)
    return F.linear(x, d["%s_W%d" % (tag, L)], d.get("%s_b%d" % (tag, L)))


def _clamped_log_std(log_std, lmax):
    """Smooth upper clamp at log(max_noise_std) -- reference models/densities.py:97-98."""
    return -F.softplus(-log_std + lmax) + lmax


def policy_step(d, s, t=0, keep=None):
    """Policy.forward with resample=False, return_samples=True, resample_noise=False --
    reference models/core.py:221-248 + DiagGaussianDensity.forward densities.py:87-121."""
    U = int(d["U"])
    o = net_forward(d, "pol", s, keep)
    if int(d["pol_has_density"]):
        mean, log_std = o.split(U, -1)
        log_std = _clamped_log_std(log_std, d["pol_lmax"])
        u = mean + _noise(d["pol_z"], t, s.shape[0]) * log_std.exp()
    else:
        u = o
    return d["act_scale"] * u.tanh() + d["act_bias"]          # models/core.py:243


def dynamics_step(d, s, a, t=0, keep=None):
    """DynamicsModel.forward((s,a), separate_outputs=True, deltas=False) up to the next state --
    reference models/core.py:265-303, Regressor.forward core.py:169-187, density scaling
    densities.py:100-107."""
    D = int(d["D"])
    x = (torch.cat([s, a], -1) - d["mx"]) * d["iSx"]          # core.py:269,177
    o = net_forward(d, "dyn", x, keep)
    if int(d["dyn_has_density"]):
        mean, log_std = o.split(D, -1)
        log_std = _clamped_log_std(log_std, d["dyn_lmax"])
        log_std = log_std + d["Sy"].log()                       # densities.py:105
        mean = mean * d["Sy"] + d["my"]                         # densities.py:106
        delta = mean + _noise(d["dyn_z"], t, s.shape[0]) * log_std.exp()
    else:
        delta = o * d["Sy"] + d["my"]                           # core.py:185
    return s + delta                                            # core.py:293,298


def reward(d, s_next, a):
    """exp(-0.5*(delta^T Q delta + a^T R a)), delta = normalised tip-to-target distance --
    reference envs/cartpole/env.py:41-86 (and double_cartpole/env.py:45-90,
    cart_acrobot/env.py:45-89, pendulum/env.py:41-79), with the constant tip map folded into
    (C, c0) by ``operands.read_reward``.  Evaluated on the NEXT state (models/core.py:293)."""
    delta = s_next @ d["rew_C"].t() + d["rew_c0"]
    cost = 0.5 * (((delta @ d["rew_Q"]) * delta).sum(-1, keepdim=True)
                  + ((a @ d["rew_R"]) * a).sum(-1, keepdim=True))
    return float(d.get("rew_scale", 1.0)) * (-cost).exp() + float(d.get("rew_offset", 0.0))


def get_z_rnd(z, i, n):
    """Cyclic rotation of the first n rows -- reference utils/rollout.py:53-59."""
    idx = (torch.arange(i, i + n) % n).long()
    return z[idx]


def mm_resample(x, z, jitter=1e-12):
    """Moment-matching re-draw m + zhat chol(cov)^T -- reference utils/rollout.py:20-29
    (unbiased covariance, 1e-12 jitter, zhat standardised per column with unbiased std, detached)."""
    M = x.shape[-2]
    m = x.mean(-2, keepdim=True)
    dx = x - m
    S = dx.transpose(-1, -2).matmul(dx) / (M - 1) + jitter * torch.eye(x.shape[-1], dtype=x.dtype)
    L = torch.linalg.cholesky(S)
    zh = ((z - z.mean(-2, keepdim=True)) / z.std(-2, keepdim=True)).detach()
    return m + zh.matmul(L.transpose(-1, -2))


# --------------------------------------------------------------------------- #
# the rollout and the mc_pilco iteration
# --------------------------------------------------------------------------- #
def rollout(d, x0, H, mm_states=False, mm_rewards=False, z_mm=None, z_rr=None, mm_groups=None,
            keep=None):
    """H-step particle rollout -- reference utils/rollout.py:93-163.
    Returns (states[H+1], actions[H], rewards[H]) as python lists like the reference."""
    N, D = x0.shape
    states, actions, rewards = [x0], [], []
    s = x0
    for i in range(H):
        kp = kd = None
        if keep is not None:
            kp, kd = [], []
            keep.append((kp, kd))
        a = policy_step(d, s, i, kp)
        s_next = dynamics_step(d, s, a, i, kd)
        r = reward(d, s_next, a)
        if mm_states:                                           # rollout.py:121-132
            z1 = get_z_rnd(z_mm, i, N)
            if mm_groups is not None:
                s_next = mm_resample(s_next.view(mm_groups, -1, D), z1.view(mm_groups, -1, D)).view(-1, D)
            else:
                s_next = mm_resample(s_next, z1)
        if mm_rewards:                                          # rollout.py:135-145
            z2 = get_z_rnd(z_rr, i, N)
            if mm_groups is not None:
                r = mm_resample(r.view(mm_groups, -1, 1), z2.view(mm_groups, -1, 1)).view(-1, 1)
            else:
                r = mm_resample(r, z2)
        actions.append(a)
        rewards.append(r)
        states.append(s_next)
        s = s_next
    return states, actions, rewards


def mc_pilco_loss(rewards, discount=None, maximize=True):
    """loss = mean_n( -sum_t disc(t) r_t ) -- reference algorithms/mc_pilco.py:46-50,134-144,190."""
    H = len(rewards)
    if discount is None:
        disc = [1.0 / H] * H
    elif callable(discount):
        disc = [discount(i) for i in range(H)]
    else:
        disc = [discount ** i for i in range(H)]
    total = torch.stack([r * w for r, w in zip(rewards, disc)]).sum(0)
    returns = -total if maximize else total
    return returns.mean()


def loss_and_grads(d, x0, H, discount=None, maximize=True, **mm):
    """Rollout + loss + reverse-mode gradient w.r.t. the policy parameters and x0 (autograd
    stands in for ``loss.backward()``, reference algorithms/mc_pilco.py:197)."""
    keys = policy_param_keys(d)
    d = dict(d)
    for k in keys:
        d[k] = d[k].detach().clone().requires_grad_(True)
    x0 = x0.detach().clone().requires_grad_(True)
    states, actions, rewards = rollout(d, x0, H, **mm)
    loss = mc_pilco_loss(rewards, discount, maximize)
    grads = torch.autograd.grad(loss, [d[k] for k in keys] + [x0])
    return {"loss": loss.detach(), "grads": dict(zip(keys, grads[:-1])), "dx0": grads[-1],
            "states": [s.detach() for s in states], "actions": [a.detach() for a in actions],
            "rewards": [r.detach() for r in rewards]}


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ semantics (reference algorithms/mc_pilco.py:209-210):
    total L2 norm, coefficient max_norm/(norm+1e-6) clamped to 1."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    coef = min(1.0, float(max_norm / (total + 1e-6)))
    return [g * coef for g in grads], total


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam single-tensor update (reference uses torch.optim.Adam(lr=1e-4),
    examples/deep_pilco_no_mm.py:163-166; step at algorithms/mc_pilco.py:213)."""
    b1, b2 = betas
    out = []
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** step
        bc2 = 1 - b2 ** step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        out.append(p - (lr / bc1) * (m / denom))
    return out


def mc_pilco_iterations(d, x0, H, iters, lr, clip=1.0, discount=None, maximize=True, **mm):
    """``iters`` policy-gradient iterations with fixed (PEGASUS) noise, fixed x0, Adam --
    the loop body of reference algorithms/mc_pilco.py:86-263 with exp=None, callbacks off."""
    keys = policy_param_keys(d)
    d = dict(d)
    m = [torch.zeros_like(d[k]) for k in keys]
    v = [torch.zeros_like(d[k]) for k in keys]
    losses = []
    for it in range(iters):
        res = loss_and_grads(d, x0, H, discount, maximize, **mm)
        grads = [res["grads"][k] for k in keys]
        if clip is not None:
            grads, _ = clip_grad_norm(grads, clip)
        new = adam_step([d[k] for k in keys], grads, m, v, it + 1, lr)
        for k, p in zip(keys, new):
            d[k] = p.detach()
        losses.append(float(res["loss"]))
    return d, losses


# --------------------------------------------------------------------------- #
# hand-derived reverse sweep: the executable spec of the CUDA backward kernel
# --------------------------------------------------------------------------- #
def manual_backward(d, states, actions, rewards_pre, saved, g_states=None, g_actions=None,
                    g_rewards=None):
    """Reverse-time sweep with explicit adjoints (no autograd), no-mm mode.

    Inputs are what the CUDA forward stores: states[t], actions[t], per-step saved activations
    ``saved[t] = (pol_h[l], pol_o, dyn_g[l], dyn_o)``; cotangents g_* are dense [H(+1), N, .]
    tensors or None.  Returns (policy grads dict, dL/dx0).  Mirrors csrc/rollout_bwd.cu step for
    step; checked against autograd in tests/test_oracle_backward.py."""
    D, U = int(d["D"]), int(d["U"])
    H = len(actions)
    N = states[0].shape[0]
    Lp, Ld = int(d["pol_L"]), int(d["dyn_L"])
    keys = policy_param_keys(d)
    grads = {k: torch.zeros_like(d[k]) for k in keys}
    Qs = d["rew_Q"] + d["rew_Q"].t()
    Rs = d["rew_R"] + d["rew_R"].t()
    rs = float(d.get("rew_scale", 1.0))
    roff = float(d.get("rew_offset", 0.0))
    gs = torch.zeros(N, D, dtype=states[0].dtype) if g_states is None else g_states[H].clone()
    for t in reversed(range(H)):
        s, a, s1 = states[t], actions[t], states[t + 1]
        pol_h, pol_o, dyn_g, dyn_o = saved[t]
        ga = torch.zeros(N, U, dtype=s.dtype) if g_actions is None else g_actions[t].clone()
        # reward adjoint
        if g_rewards is not None:
            gr = g_rewards[t].reshape(N, 1)
            e = (rewards_pre[t].reshape(N, 1) - roff)              # = scale*exp(-cost)
            delta = s1 @ d["rew_C"].t() + d["rew_c0"]
            gs = gs + (-0.5 * gr * e) * ((delta @ Qs) @ d["rew_C"])
            ga = ga + (-0.5 * gr * e) * (a @ Rs)
        # dynamics density adjoint: s1 = s + mean*Sy + my + z*exp(lstd)
        gs_prev = gs.clone()
        if int(d["dyn_has_density"]):
            mu, ls = dyn_o.split(D, -1)
            lmax = d["dyn_lmax"]
            lst = _clamped_log_std(ls, lmax) + d["Sy"].log()
            z = _noise(d["dyn_z"], t, N)
            dmu = gs * d["Sy"]
            dls = gs * z * lst.exp() * torch.sigmoid(lmax - ls)
            do = torch.cat([dmu, dls], -1)
        else:
            do = gs * d["Sy"]
        # dynamics net backward (data only; its weight grads are never consumed, App. D-4)
        delta_l = do
        for l in reversed(range(Ld)):
            dg = delta_l @ d["dyn_W%d" % (l + 1)]
            gate = (dyn_g[l] != 0).to(s.dtype)
            mk = "dyn_mask%d" % l
            if mk in d:
                gate = gate * d[mk][:N]
            gate = gate / float(d["dyn_p%d" % l])
            delta_l = dg * gate
        dx = (delta_l @ d["dyn_W0"]) * d["iSx"]
        gs_prev = gs_prev + dx[:, :D]
        ga = ga + dx[:, D:]
        # policy squash + density adjoint: a = scale*tanh(u)+bias, u = mu + z*exp(lstd)
        if int(d["pol_has_density"]):
            mu, ls = pol_o.split(U, -1)
            lmax = d["pol_lmax"]
            lst = _clamped_log_std(ls, lmax)
            z = _noise(d["pol_z"], t, N)
            u = mu + z * lst.exp()
            du = ga * d["act_scale"] * (1 - u.tanh() ** 2)
            dls = du * z * lst.exp() * torch.sigmoid(lmax - ls)
            do = torch.cat([du, dls], -1)
        else:
            du = ga * d["act_scale"] * (1 - pol_o.tanh() ** 2)
            do = du
        # policy net backward: data + weight grads
        delta_l = do
        for l in reversed(range(Lp + 1)):
            inp = pol_h[l - 1] if l > 0 else s
            grads["pol_W%d" % l] += delta_l.t() @ inp
            if ("pol_b%d" % l) in d:
                grads["pol_b%d" % l] += delta_l.sum(0)
            dh = delta_l @ d["pol_W%d" % l]
            if l > 0:
                gate = (pol_h[l - 1] != 0).to(s.dtype)
                mk = "pol_mask%d" % (l - 1)
                if mk in d:
                    gate = gate * d[mk][:N]
                gate = gate / float(d["pol_p%d" % (l - 1)])
                delta_l = dh * gate
            else:
                gs_prev = gs_prev + dh
        gs = gs_prev
        if g_states is not None:
            gs = gs + g_states[t]
    return grads, gs


def forward_with_saved(d, x0, H):
    """No-mm forward that also returns what the CUDA forward kernel stores per step."""
    states, actions, rewards, saved = [x0], [], [], []
    s = x0
    D, U = int(d["D"]), int(d["U"])
    for t in range(H):
        kp, kd = [], []
        o_p = net_forward(d, "pol", s, kp)
        if int(d["pol_has_density"]):
            mean, ls = o_p.split(U, -1)
            u = mean + _noise(d["pol_z"], t, s.shape[0]) * _clamped_log_std(ls, d["pol_lmax"]).exp()
        else:
            u = o_p
        a = d["act_scale"] * u.tanh() + d["act_bias"]
        x = (torch.cat([s, a], -1) - d["mx"]) * d["iSx"]
        o_d = net_forward(d, "dyn", x, kd)
        if int(d["dyn_has_density"]):
            mean, ls = o_d.split(D, -1)
            lst = _clamped_log_std(ls, d["dyn_lmax"]) + d["Sy"].log()
            s1 = s + mean * d["Sy"] + d["my"] + _noise(d["dyn_z"], t, s.shape[0]) * lst.exp()
        else:
            s1 = s + o_d * d["Sy"] + d["my"]
        r = reward(d, s1, a)
        saved.append((kp, o_p, kd, o_d))
        states.append(s1)
        actions.append(a)
        rewards.append(r)
        s = s1
    return states, actions, rewards, saved


# --------------------------------------------------------------------------- #
# hand-derived adjoint of moment matching: the spec of the CUDA mm reverse step
# --------------------------------------------------------------------------- #
def mm_forward_parts(x, z, jitter=1e-12):
    """mm_resample (reference utils/rollout.py:20-29) returning the pieces the reverse step needs."""
    M = x.shape[0]
    m = x.mean(0)
    dx = x - m
    S = dx.t() @ dx / (M - 1) + jitter * torch.eye(x.shape[1], dtype=x.dtype)
    L = torch.linalg.cholesky(S)
    zh = (z - z.mean(0)) / z.std(0)
    return m + zh @ L.t(), m, L, zh


def mm_backward(g_out, x, m, L, zh):
    """Adjoint of x' = m + zh L^T w.r.t. the particles x (zh detached), with
    S = (x-m)^T (x-m)/(M-1) + jitter, L = chol(S):
        dm = sum_n g_n,  dL = tril(sum_n g_n zh_n^T),
        Sbar = sym( L^-T Phi(L^T dL) L^-1 )   (Phi: lower triangle, diagonal halved),
        dx_n = dm/M + 2/(M-1) * Sbar (x_n - m)."""
    M = x.shape[0]
    dm = g_out.sum(0)
    dL = torch.tril(g_out.t() @ zh)
    A = L.t() @ dL
    Pm = torch.tril(A)
    Pm = Pm - 0.5 * torch.diag(torch.diagonal(A))
    X = torch.linalg.solve_triangular(L.t(), Pm, upper=True)            # L^-T P
    Sb = torch.linalg.solve_triangular(L.t(), X.t(), upper=True).t()   # (L^-T P) L^-1 = (L^-T X^T)^T
    Sb = 0.5 * (Sb + Sb.t())
    return dm / M + (2.0 / (M - 1)) * (x - m) @ Sb
