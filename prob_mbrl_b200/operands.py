"""Module -> operand extraction: the host half of the drop-in boundary.

The reference has no operator/FFI interface; its seam is a Python call with
duck-typed ``nn.Module`` arguments:
``utils.rollout(states, dynamics, policy, steps, ...)`` (reference
utils/rollout.py:62-79).  This file reads, *without importing the reference*,
everything the fused sm_100a rollout needs out of those modules:

  policy.model / dynamics.model   BSequential of (Linear, ReLU, [B|C]Dropout)xL, Linear
                                  (reference models/core.py:39-73)
  dropout buffers                 ``noise``/``p`` (BDropout, models/modules.py:19-61),
                                  ``concrete_noise`` (CDropout eval branch, modules.py:120-160)
  output densities                ``z``, ``max_log_std`` (models/densities.py:70-121)
  input/output scalers            ``mx, iSx, my, Sy`` (models/core.py:121-187)
  action squashing                ``scale, bias`` (models/core.py:190-248)
  reward                          ``Q, R, target, pole*_length`` of the env ``*Reward`` modules
                                  (envs/cartpole/env.py:27-86, envs/double_cartpole/env.py:27-90,
                                  envs/cart_acrobot/env.py:27-89, envs/pendulum/env.py:27-79)

Both the reference's own modules and this package's mirror (``models.py``) satisfy
the protocol, which is what makes ``rollout()`` a drop-in.  A module graph that does
not match raises :class:`NotEligible` with the reason.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import torch


class NotEligible(ValueError):
    """The module graph / flag combination is outside the fused rollout's scope.

    Deliberately NOT a RuntimeError: ``mc_pilco`` treats RuntimeError as a numerical failure of one
    iteration (reference algorithms/mc_pilco.py:122-131) and would silently skip every iteration."""


@dataclass
class NetOperands:
    """One MLP: ``L`` hidden layers + the output projection, nn.Linear layout [out, in]."""
    W: List[torch.Tensor]
    b: List[Optional[torch.Tensor]]
    mask: List[Optional[torch.Tensor]]   # per hidden layer, [rows >= N, h] or None (no dropout)
    p: List[float]                       # per hidden layer divisor (BDropout: 1-rate; CDropout: 1)
    has_density: bool = True
    z: Optional[torch.Tensor] = None     # [N, out_dims] (constant over steps) or [H, N, out_dims]
    lmax: float = 0.0
    # every mask value is exactly 0 or 1: true for BDropout.noise (Bernoulli draws) and for CDropout.concrete_noise,
    # whose (b - probs) + probs rounds to b exactly in fp32 (reference models/modules.py:61,113-116)
    masks_binary: bool = False

    @property
    def hidden(self):
        return [w.shape[0] for w in self.W[:-1]]


@dataclass
class RewardOperands:
    """r = scale*exp(-0.5*(d^T Q d + a^T R a)) + offset, d = C s' + c0  (KR rows)."""
    C: torch.Tensor
    c0: torch.Tensor
    Q: torch.Tensor
    R: torch.Tensor
    scale: float = 1.0
    offset: float = 0.0


@dataclass
class RolloutOperands:
    D: int
    U: int
    pol: NetOperands
    dyn: NetOperands
    act_scale: torch.Tensor            # [U]
    act_bias: torch.Tensor             # [U]
    mx: torch.Tensor                   # [D+U]
    iSx: torch.Tensor                  # [D+U]
    my: torch.Tensor                   # [D]
    Sy: torch.Tensor                   # [D]
    rew: RewardOperands = None
    extras: dict = field(default_factory=dict)

    # ---- (de)serialisation used by the golden fixtures and the oracle ----
    def to_flat(self):
        out = {"D": self.D, "U": self.U}
        for tag, net in (("pol", self.pol), ("dyn", self.dyn)):
            out[tag + "_L"] = len(net.W) - 1
            for i, w in enumerate(net.W):
                out["%s_W%d" % (tag, i)] = w.detach()
                if net.b[i] is not None:
                    out["%s_b%d" % (tag, i)] = net.b[i].detach()
            for i, m in enumerate(net.mask):
                if m is not None:
                    out["%s_mask%d" % (tag, i)] = m.detach()
                out["%s_p%d" % (tag, i)] = float(net.p[i])
            out[tag + "_has_density"] = int(net.has_density)
            if net.z is not None:
                out[tag + "_z"] = net.z.detach()
            out[tag + "_lmax"] = float(net.lmax)
        for k in ("act_scale", "act_bias", "mx", "iSx", "my", "Sy"):
            out[k] = getattr(self, k).detach()
        out["rew_C"], out["rew_c0"] = self.rew.C, self.rew.c0
        out["rew_Q"], out["rew_R"] = self.rew.Q, self.rew.R
        out["rew_scale"], out["rew_offset"] = float(self.rew.scale), float(self.rew.offset)
        return out

    @staticmethod
    def from_flat(d, dtype=torch.float32, device="cpu"):
        def t(x):
            return torch.as_tensor(x).to(dtype=dtype, device=device).clone()

        nets = {}
        for tag in ("pol", "dyn"):
            L = int(d[tag + "_L"])
            W = [t(d["%s_W%d" % (tag, i)]) for i in range(L + 1)]
            b = [t(d["%s_b%d" % (tag, i)]) if ("%s_b%d" % (tag, i)) in d else None for i in range(L + 1)]
            mask = [t(d["%s_mask%d" % (tag, i)]) if ("%s_mask%d" % (tag, i)) in d else None for i in range(L)]
            p = [float(d["%s_p%d" % (tag, i)]) for i in range(L)]
            nets[tag] = NetOperands(W, b, mask, p, bool(int(d[tag + "_has_density"])),
                                    t(d[tag + "_z"]) if (tag + "_z") in d else None,
                                    float(d[tag + "_lmax"]),
                                    masks_binary=all(bool(((m == 0) | (m == 1)).all()) for m in mask if m is not None))
        rew = RewardOperands(t(d["rew_C"]), t(d["rew_c0"]), t(d["rew_Q"]), t(d["rew_R"]),
                             float(d["rew_scale"]), float(d["rew_offset"]))
        return RolloutOperands(int(d["D"]), int(d["U"]), nets["pol"], nets["dyn"],
                               t(d["act_scale"]), t(d["act_bias"]), t(d["mx"]), t(d["iSx"]),
                               t(d["my"]), t(d["Sy"]), rew)

    def policy_parameters(self):
        """Flat list in ``policy.parameters()`` order: W0, b0, W1, b1, ..., Wout, bout."""
        out = []
        for w, b in zip(self.pol.W, self.pol.b):
            out.append(w)
            if b is not None:
                out.append(b)
        return out


# --------------------------------------------------------------------------- #
# duck-typed readers
# --------------------------------------------------------------------------- #
def _is_linear(m):
    return isinstance(m, torch.nn.Linear)


def _is_relu(m):
    return isinstance(m, torch.nn.ReLU)


def _is_cdropout(m):
    return hasattr(m, "concrete_noise") and hasattr(m, "logit_p")


def _is_bdropout(m):
    return hasattr(m, "noise") and hasattr(m, "p") and hasattr(m, "rate") and not _is_cdropout(m)


def _is_diag_gaussian(m):
    return (m is not None and hasattr(m, "max_log_std") and hasattr(m, "z")
            and hasattr(m, "output_dims") and not hasattr(m, "n_components"))


def _scalar(x, what):
    x = torch.as_tensor(x).detach().float().reshape(-1)
    if x.numel() != 1 and not bool((x == x[0]).all()):
        raise NotEligible("%s is not a single value" % what)
    return float(x[0])


def read_net(seq, n_rows, what):
    """Walk a BSequential-like container (reference models/core.py:39-73 naming) into NetOperands.

    Dropout mask semantics mirror models/modules.py:46-61 (BDropout: x*noise[:N]/p) and
    modules.py:120-160 (CDropout in eval mode: x*concrete_noise[:N], detached, no /p).
    """
    children = list(seq._modules.values())
    W, b, mask, p = [], [], [], []
    density = None
    i = 0
    while i < len(children):
        m = children[i]
        if not _is_linear(m):
            raise NotEligible("%s: unsupported layer %s at position %d" % (what, type(m).__name__, i))
        W.append(m.weight)
        b.append(m.bias)
        i += 1
        if i == len(children):
            break
        if _is_diag_gaussian(children[i]):
            if i != len(children) - 1:
                raise NotEligible("%s: density is not the last module" % what)
            density = children[i]
            break
        if not _is_relu(children[i]):
            raise NotEligible("%s: only ReLU hidden activations are fused (got %s)"
                              % (what, type(children[i]).__name__))
        i += 1
        if i < len(children) and (_is_cdropout(children[i]) or _is_bdropout(children[i])):
            drop = children[i]
            h = m.weight.shape[0]
            if _is_cdropout(drop):
                if drop.training:
                    raise NotEligible("%s: CDropout in train mode re-relaxes every call "
                                      "(reference models/modules.py:151-153); call .eval()" % what)
                noise = drop.concrete_noise
                div = 1.0
            else:
                noise = drop.noise
                div = _scalar(drop.p, "%s dropout keep-probability" % what)
            if noise.dim() != 2 or noise.shape[1] != h or noise.shape[0] < n_rows:
                raise NotEligible("%s: dropout mask buffer has shape %s, need [>=%d, %d]; run one "
                                  "step through the module so it allocates it (reference "
                                  "models/modules.py:48-54,140-149)" % (what, tuple(noise.shape), n_rows, h))
            mask.append(noise.detach())
            p.append(div)
            i += 1
        else:
            mask.append(None)
            p.append(1.0)
    if len(W) < 1:
        raise NotEligible("%s: empty network" % what)
    if len(mask) != len(W) - 1:
        raise NotEligible("%s: output projection must not be followed by an activation" % what)
    net = NetOperands(W, b, mask, p, has_density=density is not None, masks_binary=True)
    return net, density


def _density_fields(net, density, n_rows, out_dims, what):
    if density is None:
        net.has_density = False
        return
    if int(density.output_dims) != out_dims:
        raise NotEligible("%s: density output_dims %s != %d" % (what, density.output_dims, out_dims))
    z = density.z
    if z.dim() != 2 or z.shape[1] != out_dims or z.shape[0] < n_rows:
        raise NotEligible("%s: density noise buffer z has shape %s, need [>=%d, %d]"
                          % (what, tuple(z.shape), n_rows, out_dims))
    net.has_density = True
    net.z = z.detach()
    net.lmax = float(density.max_log_std)


def _expand_angles(x, dims):
    """[others, sin(angles), cos(angles)] -- the layout of reference utils/angles.py:39-42."""
    others = [i for i in range(x.shape[-1]) if i not in dims]
    return torch.cat([x[..., others], x[..., dims].sin(), x[..., dims].cos()], -1)


def read_reward(reward_func, D, U):
    """Reduce one of the env ``*Reward`` modules to the (C, c0, Q, R) tip-distance form.

    All four reference rewards are r = exp(-0.5*(d^T Q d + u^T R u)) with d a fixed linear map
    of the (angle-expanded) next state minus a constant target tip:
      Cartpole        envs/cartpole/env.py:41-86       tip = [x + l sin(th), -l cos(th)],   / (2 l)
      DoubleCartpole  envs/double_cartpole/env.py:45-90 tip = [x - l1 s1 - l2 s2, l1 c1 + l2 c2], / (2 (l1+l2))
      CartAcrobot     envs/cart_acrobot/env.py:45-89   same tip map as DoubleCartpole
      Pendulum        envs/pendulum/env.py:41-79       tip = [l sin(th), -l cos(th)],        / (2 l)
    A custom reward can opt in by exposing ``tip_quadratic_form() -> (C, c0, Q, R[, scale, offset])``.
    """
    if hasattr(reward_func, "tip_quadratic_form"):
        form = reward_func.tip_quadratic_form()
        C, c0, Q, R = [torch.as_tensor(v).detach().float() for v in form[:4]]
        scale, offset = (float(form[4]), float(form[5])) if len(form) >= 6 else (1.0, 0.0)
    else:
        if not (hasattr(reward_func, "Q") and hasattr(reward_func, "R") and hasattr(reward_func, "target")):
            raise NotEligible("reward_func %s is not a known tip-distance reward" % type(reward_func).__name__)
        Q = reward_func.Q.detach().float().cpu()
        R = reward_func.R.detach().float().cpu()
        target = reward_func.target.detach().float().cpu().reshape(1, -1)
        C = torch.zeros(2, D)
        if hasattr(reward_func, "pole1_length"):
            l1 = float(reward_func.pole1_length)
            l2 = float(reward_func.pole2_length)
            ta = _expand_angles(target, [2, 4])
            if ta.shape[-1] != D or D != 8:
                raise NotEligible("double-pole reward expects 8 angle-expanded state dims, got %d" % D)
            C[0, 0], C[0, 4], C[0, 5] = 1.0, -l1, -l2
            C[1, 6], C[1, 7] = l1, l2
            norm = 2.0 * (l1 + l2)
        elif hasattr(reward_func, "pole_length"):
            lp = float(reward_func.pole_length)
            if target.shape[-1] == 4:          # cartpole: [x, xdot, theta, thetadot]
                ta = _expand_angles(target, [2])
                if ta.shape[-1] != D:
                    raise NotEligible("cartpole reward expects 5 angle-expanded state dims, got %d" % D)
                C[0, 0], C[0, 3] = 1.0, lp
                C[1, 4] = -lp
            elif target.shape[-1] == 2:        # pendulum: [theta, thetadot]
                ta = _expand_angles(target, [0])
                if ta.shape[-1] != D:
                    raise NotEligible("pendulum reward expects 3 angle-expanded state dims, got %d" % D)
                C[0, 1] = lp
                C[1, 2] = -lp
            else:
                raise NotEligible("unrecognised single-pole reward target of size %d" % target.shape[-1])
            norm = 2.0 * lp
        else:
            raise NotEligible("reward_func %s is not a known tip-distance reward" % type(reward_func).__name__)
        tgt_tip = (ta @ C.t()).reshape(-1)
        C = C / norm
        c0 = -tgt_tip / norm
        scale, offset = 1.0, 0.0
    if Q.shape != (C.shape[0], C.shape[0]) or R.shape != (U, U) or C.shape[1] != D:
        raise NotEligible("reward operand shapes do not match D=%d U=%d" % (D, U))
    return RewardOperands(C.contiguous(), c0.contiguous(), Q.contiguous(), R.contiguous(), scale, offset)


def extract(dynamics, policy, n_rows, D=None):
    """Read a (DynamicsModel, Policy) pair into :class:`RolloutOperands` for ``n_rows`` particles.

    Mirrors what one step of the reference consumes: Policy.forward (models/core.py:221-248),
    DynamicsModel.forward with separate_outputs=True, deltas=False (models/core.py:265-303).
    """
    for mod, what in ((policy, "policy"), (dynamics, "dynamics")):
        ad = getattr(mod, "angle_dims", None)
        if ad is not None and len(ad) > 0:
            raise NotEligible("%s.angle_dims must be empty (broken in the reference too, "
                              "utils/angles.py:31-35)" % what)
    pol, pol_density = read_net(policy.model, n_rows, "policy.model")
    dyn, dyn_density = read_net(dynamics.model, n_rows, "dynamics.model")
    if dyn_density is not None:
        raise NotEligible("dynamics.model must not end in a density; pass output_density=")
    dyn_density = getattr(dynamics, "output_density", None)
    if dyn_density is not None and not _is_diag_gaussian(dyn_density):
        raise NotEligible("dynamics.output_density %s is not DiagGaussianDensity"
                          % type(dyn_density).__name__)
    if D is None:
        D = pol.W[0].shape[1]
    n_pol_out = pol.W[-1].shape[0]
    U = n_pol_out // 2 if pol_density is not None else n_pol_out
    if pol.W[0].shape[1] != D or dyn.W[0].shape[1] != D + U:
        raise NotEligible("layer input sizes do not match D=%d, U=%d" % (D, U))
    n_dyn_out = dyn.W[-1].shape[0]
    if n_dyn_out != (2 * D if dyn_density is not None else D):
        raise NotEligible("dynamics output size %d does not match D=%d (learned-reward heads are "
                          "broken in the reference rollout, models/core.py:286-296)" % (n_dyn_out, D))
    _density_fields(pol, pol_density, n_rows, U, "policy")
    _density_fields(dyn, dyn_density, n_rows, D, "dynamics")
    if not callable(getattr(dynamics, "reward_func", None)):
        raise NotEligible("dynamics.reward_func must be a known reward module")

    dev, dt = pol.W[0].device, pol.W[0].dtype

    def vec(x, n, what):
        x = torch.as_tensor(x).detach().to(device=dev, dtype=dt).reshape(-1)
        if x.numel() == 1:
            x = x.expand(n)
        if x.numel() != n:
            raise NotEligible("%s has %d entries, expected %d" % (what, x.numel(), n))
        return x.contiguous()

    rew = read_reward(dynamics.reward_func, D, U)
    for k in ("C", "c0", "Q", "R"):
        setattr(rew, k, getattr(rew, k).to(device=dev, dtype=dt).contiguous())
    return RolloutOperands(
        D=D, U=U, pol=pol, dyn=dyn,
        act_scale=vec(policy.scale, U, "policy.scale"), act_bias=vec(policy.bias, U, "policy.bias"),
        mx=vec(dynamics.mx, D + U, "dynamics.mx"), iSx=vec(dynamics.iSx, D + U, "dynamics.iSx"),
        my=vec(dynamics.my, D, "dynamics.my"), Sy=vec(dynamics.Sy, D, "dynamics.Sy"), rew=rew)


def materialize_noise(dynamics, policy, states):
    """Let the modules allocate their [N, h] masks / [N, .] z buffers exactly as step 0 of the
    reference rollout would (lazy allocation from the global RNG in the order pol drop0.., pol z,
    dyn drop0.., dyn z -- SURVEY.md App. B.2; reference models/modules.py:48-54,140-149,
    models/densities.py:113-116).  No-op when the buffers already fit."""
    with torch.no_grad():
        a = policy(states, resample=False, return_samples=True, resample_noise=False)
        dynamics((states, a), return_samples=True, separate_outputs=True, deltas=False,
                 resample=False, resample_noise=False)
