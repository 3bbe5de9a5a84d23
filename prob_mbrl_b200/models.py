"""Host-side mirror of the reference's model API for the rollout path.

Same class names, constructor arguments, buffer/parameter names (so ``state_dict``s are
interchangeable with the reference's) and call semantics as ``prob_mbrl.models``:

  mlp                  reference models/core.py:15-99
  BDropout / CDropout  reference models/modules.py:14-171
  BSequential          reference models/modules.py:198-274
  DiagGaussianDensity  reference models/densities.py:70-148
  Regressor / DynamicsModel / Policy   reference models/core.py:121-303

These are the *eager PyTorch* modules the examples build and that stay on the CPU side of the
boundary (env interaction via ``apply_controller``, dynamics fitting via ``train_regressor``).
The fused sm_100a rollout never calls their ``forward``: ``operands.extract`` reads their
tensors and ``rollout()`` runs the CUDA sweeps.  They exist here because the reference package
is not importable on the GPU box; when the reference IS installed its own modules satisfy the
same duck-typed protocol and can be passed to ``prob_mbrl_b200.rollout`` directly.
"""
import copy
import inspect
from collections import OrderedDict
from collections.abc import Iterable
from functools import partial

import numpy as np
import torch
from torch import nn


class StochasticModule(nn.Module):
    """Marker base: BSequential forwards resample/seed keywords only to these."""


class BDropout(StochasticModule):
    """Bernoulli dropout with a persistent mask (PEGASUS-style reuse of the first N rows)."""

    def __init__(self, rate=0.5, name=None, regularizer_scale=1.0, **kwargs):
        super().__init__(**kwargs)
        self.name = name
        self.register_buffer("regularizer_scale", torch.tensor(0.5 * regularizer_scale))
        self.register_buffer("rate", rate if torch.is_tensor(rate) else torch.tensor(rate))
        self.register_buffer("p", 1 - self.rate)
        self.register_buffer("noise", torch.bernoulli(self.p))

    def weights_regularizer(self, weights):
        self.p = 1 - self.rate
        return self.regularizer_scale * (self.p * weights.pow(2).sum(0)).sum()

    def biases_regularizer(self, biases):
        return self.regularizer_scale * biases.pow(2).sum()

    def resample(self, seed=None):
        self.update_noise(self.noise, seed)

    def update_noise(self, like, seed=None):
        if seed is not None:
            torch.manual_seed(seed)
        self.p = 1 - self.rate
        self.noise.data = torch.bernoulli(self.p.expand(like.shape))

    def _needs_new_mask(self, x, mask_dims, stored):
        shape = x.shape[-mask_dims:]
        return shape[1:] != stored.shape[1:] or shape[0] > stored.shape[0]

    def forward(self, x, resample=True, mask_dims=2, seed=None, **kwargs):
        if self._needs_new_mask(x, mask_dims, self.noise):
            self.update_noise(x.reshape(-1, *x.shape[-mask_dims:])[0], seed)
        elif resample:
            if seed is not None:
                torch.manual_seed(seed)
            return x * torch.bernoulli(self.p.expand(x.shape)) / self.p
        rows = x.shape[-mask_dims]
        return x * self.noise[..., :rows, :].detach() / self.p

    def extra_repr(self):
        return "rate={}".format(self.rate.mean().item() if self.rate.dim() else self.rate.item())


class CDropout(BDropout):
    """Concrete dropout (learnable keep-probability); eval mode freezes a hard mask."""

    def __init__(self, rate=0.5, name=None, regularizer_scale=1.0, dropout_regularizer=1.0,
                 temperature=0.1, **kwargs):
        super().__init__(rate, name, regularizer_scale, **kwargs)
        self.register_buffer("temp", torch.tensor(temperature))
        self.register_buffer("dropout_regularizer", torch.tensor(dropout_regularizer))
        self.logit_p = nn.Parameter(-torch.log(1.0 / self.p - 1.0))
        self.register_buffer("concrete_noise", torch.bernoulli(self.p))

    def weights_regularizer(self, weights):
        p = self.p
        reg = self.regularizer_scale * (p * weights.pow(2).sum(0))
        reg = reg + self.dropout_regularizer * (p * p.log() + (1 - p) * (1 - p).log())
        return reg.sum()

    def update_noise(self, like, seed=None):
        if seed is not None:
            torch.manual_seed(seed)
        self.noise.data = torch.rand_like(like)
        if not self.training:
            self.update_concrete_noise(self.noise)

    def update_concrete_noise(self, noise):
        logits = self.logit_p + ((noise + 1e-7) / (1 - (noise - 1e-7))).log()
        probs = (logits / self.temp).sigmoid()
        hard = torch.bernoulli(probs)
        # straight-through: hard sample forward, relaxed gradient backward
        self.concrete_noise = (hard - probs).detach() + probs
        self.p = self.logit_p.sigmoid()

    def forward(self, x, resample=False, mask_dims=2, seed=None, **kwargs):
        noise, fresh = self.noise, False
        if resample:
            if seed is not None:
                torch.manual_seed(seed)
            noise, fresh = torch.rand_like(x), True
        elif (self._needs_new_mask(x, mask_dims, self.noise)
              or self._needs_new_mask(x, mask_dims, self.concrete_noise)):
            self.update_noise(x.reshape(-1, *x.shape[-mask_dims:])[0], seed)
            noise, fresh = self.noise, True
        if self.training:
            self.update_concrete_noise(noise)
            mask = self.concrete_noise
        else:
            if fresh:
                self.update_concrete_noise(noise)
            mask = self.concrete_noise.detach()
        rows = x.shape[-mask_dims]
        return x * mask[..., :rows, :]


class BSequential(nn.Sequential):
    """nn.Sequential that routes resample/seed keywords to its stochastic children."""

    def __init__(self, *args):
        super().__init__(*args)
        self.modules_to_regularize = []

    def resample(self, seed=None):
        i = 0
        for m in self._modules.values():
            if isinstance(m, BDropout):
                m.resample(None if seed is None else seed + i)
                i += 1

    def forward(self, input, resample=True, repeat_mask=False, **kwargs):
        for m in self._modules.values():
            if isinstance(m, StochasticModule):
                input = m(input, resample=resample, repeat_mask=repeat_mask, **kwargs)
            else:
                input = m(input)
        return input

    def regularization_loss(self):
        mods = list(self._modules.values())
        if not self.modules_to_regularize:
            for i, m in enumerate(mods):
                if hasattr(m, "weights_regularizer"):
                    for nxt in mods[i:]:
                        if isinstance(nxt, (nn.Linear, nn.modules.conv._ConvNd)):
                            entry = {"module": m, "weight": nxt.weight}
                            if nxt.bias is not None and hasattr(m, "biases_regularizer"):
                                entry["bias"] = nxt.bias
                            self.modules_to_regularize.append(entry)
                            break
                elif hasattr(m, "regularization_loss"):
                    m.regularization_loss()
                    self.modules_to_regularize.extend(m.modules_to_regularize)
        loss = 0
        for e in self.modules_to_regularize:
            loss = loss + e["module"].weights_regularizer(e["weight"])
            if "bias" in e:
                loss = loss + e["module"].biases_regularizer(e["bias"])
        return loss


class DiagGaussianDensity(StochasticModule):
    """Splits the incoming features into (mean, log_std) of a diagonal Gaussian."""

    def __init__(self, output_dims, max_noise_std=5.0):
        super().__init__()
        self.output_dims = output_dims
        self.register_buffer("z", torch.ones([1, 1]))
        self.register_buffer("max_log_std", torch.tensor(max_noise_std).log())

    def resample(self, seed=None):
        if seed is not None:
            torch.manual_seed(seed)
        self.z.data = torch.randn_like(self.z)

    def forward(self, x, scaling_params=None, return_samples=False, resample_noise=True, seed=None, **kwargs):
        mean, log_std = x.split(int(self.output_dims), -1)
        log_std = self.max_log_std - nn.functional.softplus(self.max_log_std - log_std)
        if scaling_params is not None and len(scaling_params) == 2:
            my, Sy = scaling_params
            log_std = log_std + Sy.log()
            mean = mean * Sy + my
        if not return_samples:
            return mean, log_std
        if mean.shape != self.z.shape or resample_noise:
            if seed is not None:
                torch.manual_seed(seed)
            self.z.data = torch.randn_like(mean)
        return mean + self.z * log_std.exp()

    def log_prob(self, z, mean, log_std=None):
        d = mean - z
        if log_std is None:
            return -0.5 * d.pow(2).sum(-1)
        return (-0.5 * (d * (-log_std).exp()).pow(2).sum(-1) - log_std.sum(-1)
                - self.output_dims * 0.5 * float(np.log(2 * np.pi)))

    def __repr__(self):
        return "DiagGaussianDensity(output_dims=%d)" % self.output_dims


def mlp(input_dims, output_dims, hidden_dims=[200, 200], nonlin=nn.ReLU, output_nonlin=None,
        weights_initializer=partial(nn.init.xavier_normal_, gain=nn.init.calculate_gain("relu")),
        biases_initializer=partial(nn.init.uniform_, a=-1e-1, b=1e-1),
        hidden_biases=True, output_biases=True, dropout_layers=BDropout, input_dropout=None,
        layer_norm=False):
    """Multilayer perceptron as a BSequential with children fc{i}, nonlin{i}, drop{i}, fc_out[, fc_nonlin]."""
    widths = [input_dims] + list(hidden_dims)
    n_hidden = len(hidden_dims)
    if not isinstance(dropout_layers, Iterable):
        dropout_layers = [copy.deepcopy(dropout_layers)] * n_hidden
    if not isinstance(nonlin, Iterable):
        nonlin = [nonlin] * n_hidden
    layers = OrderedDict()
    if inspect.isclass(input_dropout):
        input_dropout = input_dropout(name="drop_input")
    if input_dropout is not None:
        layers["drop_input"] = input_dropout
    for i in range(n_hidden):
        drop = dropout_layers[i]
        if inspect.isclass(drop):
            drop = drop(name="drop%d" % i)
        layers["fc%d" % i] = nn.Linear(widths[i], widths[i + 1], bias=hidden_biases)
        if layer_norm:
            layers["ln%d" % i] = nn.LayerNorm(widths[i + 1])
        if callable(nonlin[i]):
            layers["nonlin%d" % i] = nonlin[i]()
        if drop is not None:
            layers["drop%d" % i] = drop
    layers["fc_out"] = nn.Linear(widths[-1], output_dims, bias=output_biases)
    if callable(output_nonlin):
        layers["fc_nonlin"] = output_nonlin()
    net = BSequential(layers)
    if callable(weights_initializer):
        for m in net.modules():
            if hasattr(m, "weight") and not isinstance(m, nn.LayerNorm):
                weights_initializer(m.weight)
    if callable(biases_initializer):
        for m in net.modules():
            if getattr(m, "bias", None) is not None:
                biases_initializer(m.bias)
    return net.float()


class Regressor(nn.Module):
    """Input/output-whitened regressor around a BSequential + optional output density."""

    def __init__(self, model, output_density=None, angle_dims=[]):
        super().__init__()
        self.model = model
        self.output_density = output_density
        self.register_buffer("angle_dims", torch.tensor(angle_dims).long())
        for name, fill in (("X", 1.0), ("Y", 1.0), ("mx", 0.0), ("Sx", 1.0), ("iSx", 1.0),
                           ("my", 0.0), ("Sy", 1.0), ("iSy", 1.0)):
            self.register_buffer(name, torch.full([1, 1], fill))

    def set_dataset(self, X, Y, **kwargs):
        if len(self.angle_dims):
            raise NotImplementedError("angle_dims on Regressor is broken in the reference as well "
                                      "(utils/angles.py:31-35); expand angles in the environment")
        self.X.data, self.Y.data = X, Y
        self.mx.data = X.mean(0, keepdim=True)
        self.Sx.data = 4.0 * X.std(0, keepdim=True)
        self.Sx.data[self.Sx == 0] = 4.0
        self.iSx.data = self.Sx.reciprocal()
        self.my.data = Y.mean(0, keepdim=True)
        self.Sy.data = 4.0 * Y.std(0, keepdim=True)
        self.Sy.data[self.Sy == 0] = 4.0
        self.iSy.data = self.Sy.reciprocal()

    def load(self, state_dict):
        own = dict(self.named_parameters())
        own.update(self.named_buffers())
        for k, v in state_dict.items():
            if k in own:
                own[k].data = v.data.clone()

    def regularization_loss(self):
        return self.model.regularization_loss()

    def resample(self, *args, **kwargs):
        self.model.resample(*args, **kwargs)
        if self.output_density is not None:
            self.output_density.resample(*args, **kwargs)

    def forward(self, x, normalize=True, **kwargs):
        if normalize:
            x = (x - self.mx) * self.iSx
        outs = self.model(x, **kwargs)
        if callable(self.output_density):
            scaling = (self.my, self.Sy) if normalize else None
            return self.output_density(outs, scaling_params=scaling, **kwargs)
        return outs * self.Sy + self.my


class DynamicsModel(Regressor):
    """(state, action) -> next state (or state delta) and reward from a known reward function."""

    def __init__(self, model, reward_func=None, predict_done=False, **kwargs):
        super().__init__(model, **kwargs)
        self.register_buffer("maxR", torch.ones([1, 1]))
        self.register_buffer("minR", torch.ones([1, 1]))
        self.reward_func = reward_func

    def set_dataset(self, X, Y):
        super().set_dataset(X, Y)
        R = self.Y[..., -1]
        self.maxR.data, self.minR.data = R.max(), R.min()

    def forward(self, inputs, separate_outputs=False, deltas=True, **kwargs):
        paired = isinstance(inputs, (tuple, list))
        if paired:
            prev_states, actions = inputs[0], inputs[1]
            inputs = torch.cat([prev_states, actions], -1)
        outs = super().forward(inputs, **kwargs)
        if not kwargs.get("return_samples", False):
            return outs
        if not paired:
            D = outs.shape[-1] - (0 if callable(self.reward_func) else 1)
            prev_states, actions = inputs[..., :D], inputs[..., D:]
        if callable(self.reward_func):
            dstates = outs
            rewards = self.reward_func(prev_states + dstates, actions)
        else:
            dstates, rewards = outs[..., :-1], outs[..., -1:]
        states = dstates if deltas else prev_states + dstates
        if separate_outputs:
            return states, rewards
        return torch.cat([states, rewards], -1)


class Policy(nn.Module):
    """Stochastic NN policy with outputs squashed into [minU, maxU]."""

    def __init__(self, model, maxU=1.0, minU=None, angle_dims=[]):
        super().__init__()
        self.model = model
        self.register_buffer("angle_dims", torch.tensor(angle_dims).long())
        if minU is None:
            minU = -maxU
        self.register_buffer("scale", torch.tensor(0.5 * (maxU - minU)).squeeze())
        self.register_buffer("bias", torch.tensor(0.5 * (maxU + minU)).squeeze())

    def regularization_loss(self):
        return self.model.regularization_loss()

    def resample(self, *args, **kwargs):
        self.model.resample(*args, **kwargs)

    load = Regressor.load

    def forward(self, x, **kwargs):
        as_numpy = isinstance(x, np.ndarray)
        kwargs.setdefault("resample", True)
        kwargs.setdefault("return_samples", True)
        x = torch.as_tensor(x).to(dtype=self.scale.dtype, device=self.scale.device)
        if x.dim() == 1:
            x = x[None, :]
        if len(self.angle_dims) > 0:
            raise NotImplementedError("angle_dims on Policy is broken in the reference as well")
        u = self.model(x, **kwargs)
        if isinstance(u, tuple):
            u = u[0] + u[1]
        u = self.scale * u.tanh() + self.bias
        return u.detach().cpu().numpy() if as_numpy else u
