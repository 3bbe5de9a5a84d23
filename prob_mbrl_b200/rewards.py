"""Host-side mirror of the env reward modules the rollout evaluates every step.

All four reference rewards are the same function: the exponentiated negative quadratic cost of the
normalised distance between the pendulum tip and a fixed target tip, plus a control penalty,
    r = exp(-0.5 * (d^T Q d + u^T R u)),   d = (tip(x) - tip(target)) / norm
(reference envs/cartpole/env.py:27-86, envs/double_cartpole/env.py:27-90,
envs/cart_acrobot/env.py:27-89, envs/pendulum/env.py:27-79).  Here they share one base class that
states the tip map once as a matrix on the angle-expanded state; ``tip_quadratic_form()`` hands
exactly that matrix to the fused kernel (see operands.read_reward).  Parameter names match the
reference modules so state_dicts interchange.
"""
import numpy as np
import torch
from torch import nn


def to_complex(x, dims):
    """[non-angle dims, sin(angle dims), cos(angle dims)] -- layout of reference utils/angles.py:39-42."""
    dims = list(dims)
    if not dims:
        return x
    others = [i for i in range(x.shape[-1]) if i not in dims]
    return torch.cat([x[..., others], x[..., dims].sin(), x[..., dims].cos()], -1)


class TipQuadraticReward(nn.Module):
    angle_dims = ()

    def __init__(self, target, Q, R):
        super().__init__()
        self.Q = nn.Parameter(Q, requires_grad=False)
        self.R = nn.Parameter(R, requires_grad=False)
        self.target = nn.Parameter(target.unsqueeze(0) if target.dim() == 1 else target, requires_grad=False)

    def tip_matrix(self):
        """(M [2, D_expanded], norm): tip = M @ expanded_state, d = (tip - tip_target) / norm."""
        raise NotImplementedError

    def tip_quadratic_form(self):
        M, norm = self.tip_matrix()
        M = M.to(self.Q.dtype)
        tgt = to_complex(self.target.detach().cpu(), self.angle_dims) @ M.t()
        return M / norm, (-tgt / norm).reshape(-1), self.Q.detach().cpu(), self.R.detach().cpu()

    def forward(self, x, u):
        x = torch.as_tensor(x).to(device=self.Q.device, dtype=self.Q.dtype)
        u = torch.as_tensor(u).to(device=self.Q.device, dtype=self.Q.dtype)
        x = x.unsqueeze(0) if x.dim() == 1 else x
        u = u.unsqueeze(0) if u.dim() == 1 else u
        M, norm = self.tip_matrix()
        M = M.to(device=self.Q.device, dtype=self.Q.dtype)
        if x.shape[-1] != M.shape[1]:
            x = to_complex(x, self.angle_dims)
        tgt = to_complex(self.target, self.angle_dims) @ M.t()
        delta = (x @ M.t() - tgt) / norm
        cost = 0.5 * ((delta.mm(self.Q) * delta).sum(-1, keepdim=True) + (u.mm(self.R) * u).sum(-1, keepdim=True))
        return (-cost).exp()


class CartpoleReward(TipQuadraticReward):
    angle_dims = (2,)

    def __init__(self, pole_length=0.5, target=torch.tensor([0, 0, np.pi, 0]), Q=16.0 * torch.eye(2),
                 R=1e-4 * torch.eye(1)):
        super().__init__(target.float(), Q, R)
        self.pole_length = nn.Parameter(torch.as_tensor(pole_length, dtype=torch.float32), requires_grad=False)

    def tip_matrix(self):   # expanded state [x, xdot, thetadot, sin, cos]
        lp = float(self.pole_length)
        return torch.tensor([[1.0, 0, 0, lp, 0], [0, 0, 0, 0, -lp]]), 2.0 * lp


class PendulumReward(TipQuadraticReward):
    angle_dims = (0,)

    def __init__(self, pole_length=1.0, target=torch.tensor([np.pi, 0]), Q=4.0 * torch.eye(2),
                 R=1e-4 * torch.eye(1)):
        super().__init__(target.float(), Q, R)
        self.pole_length = nn.Parameter(torch.as_tensor(pole_length, dtype=torch.float32), requires_grad=False)

    def tip_matrix(self):   # expanded state [thetadot, sin, cos]
        lp = float(self.pole_length)
        return torch.tensor([[0.0, lp, 0], [0, 0, -lp]]), 2.0 * lp


class DoubleCartpoleReward(TipQuadraticReward):
    angle_dims = (2, 4)

    def __init__(self, pole1_length=0.6, pole2_length=0.6, target=torch.zeros(6), Q=8.0 * torch.eye(2),
                 R=1e-3 * torch.eye(1)):
        super().__init__(target.float(), Q, R)
        self.pole1_length = nn.Parameter(torch.as_tensor(pole1_length, dtype=torch.float32), requires_grad=False)
        self.pole2_length = nn.Parameter(torch.as_tensor(pole2_length, dtype=torch.float32), requires_grad=False)

    def tip_matrix(self):   # expanded state [x, xdot, th1dot, th2dot, sin1, sin2, cos1, cos2]
        l1, l2 = float(self.pole1_length), float(self.pole2_length)
        return torch.tensor([[1.0, 0, 0, 0, -l1, -l2, 0, 0], [0, 0, 0, 0, 0, 0, l1, l2]]), 2.0 * (l1 + l2)


class CartAcrobotReward(DoubleCartpoleReward):
    def __init__(self, pole1_length=0.6, pole2_length=0.6, target=torch.zeros(6), Q=8.0 * torch.eye(2),
                 R=1e-4 * torch.eye(1)):
        super().__init__(pole1_length, pole2_length, target, Q, R)


# --------------------------------------------------------------------------------------------------
# Generic quadratic costs (reference losses.py:67-75)
# --------------------------------------------------------------------------------------------------
def quadratic_loss(states, target, Q):
    """(x - t)^T Q (x - t), shape [N, 1]  (reference losses.py:67-71)."""
    target, Q = target.to(states.device), Q.to(states.device)
    d = states - target
    return (d.mm(Q) * d).sum(-1)[:, None]


def quadratic_saturating_loss(states, target, Q):
    """1 - exp(-0.5 (x - t)^T Q (x - t))  (reference losses.py:74-75)."""
    return 1 - (-0.5 * quadratic_loss(states, target, Q)).exp()


class QuadraticSaturatingCost(nn.Module):
    """``reward_func`` built on the reference's ``losses.quadratic_saturating_loss``: a full D x D quadratic form on
    the next state around ``target`` (plus an optional control penalty u^T R u inside the exponent), as a COST in
    [0, 1) (``reward=False``, use with ``mc_pilco(..., maximize=False)``) or as the reward exp(-0.5 q) = 1 - cost.
    It is the same function family as the env tip rewards with the distance map C = I, so the fused sweeps evaluate
    it in-kernel: ``tip_quadratic_form()`` hands (I, -target, Q, R, scale, offset) to operands.read_reward."""

    def __init__(self, target, Q, R=None, reward=False):
        super().__init__()
        target = torch.as_tensor(target, dtype=torch.float32).reshape(1, -1)
        Q = torch.as_tensor(Q, dtype=torch.float32)
        if Q.shape != (target.shape[1], target.shape[1]):
            raise ValueError("Q must be [D, D] for a target of D entries")
        self.target = nn.Parameter(target, requires_grad=False)
        self.Q = nn.Parameter(Q, requires_grad=False)
        self.R = None if R is None else nn.Parameter(torch.as_tensor(R, dtype=torch.float32), requires_grad=False)
        self.reward = bool(reward)

    def forward(self, x, u):
        x = torch.as_tensor(x).to(device=self.Q.device, dtype=self.Q.dtype)
        x = x.unsqueeze(0) if x.dim() == 1 else x
        q = quadratic_loss(x, self.target, self.Q)
        if self.R is not None:
            u = torch.as_tensor(u).to(device=self.Q.device, dtype=self.Q.dtype)
            u = u.unsqueeze(0) if u.dim() == 1 else u
            q = q + (u.mm(self.R) * u).sum(-1, keepdim=True)
        e = (-0.5 * q).exp()
        return e if self.reward else 1 - e

    def tip_quadratic_form(self):
        D = self.target.shape[1]
        U = 1 if self.R is None else self.R.shape[0]
        R = torch.zeros(U, U) if self.R is None else self.R.detach().cpu()
        scale, offset = (1.0, 0.0) if self.reward else (-1.0, 1.0)
        return torch.eye(D), -self.target.detach().cpu().reshape(-1), self.Q.detach().cpu(), R, scale, offset
