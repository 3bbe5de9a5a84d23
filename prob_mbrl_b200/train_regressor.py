"""Drop-in replacement for ``prob_mbrl.utils.train_regressor`` (reference utils/train_regressor.py:58-165):
minibatch maximum-likelihood fit of the Bayesian-NN dynamics model (train-mode concrete dropout + Gaussian NLL +
dropout / weight regulariser), SURVEY.md section 8f row 1.

Two paths behind the reference's signature:
  * fused (CUDA model of the pattern (Linear, ReLU, CDropout) x L, Linear + DiagGaussianDensity, plain Adam,
    default flags): every iteration = ``pmb_fit_gradient`` (forward + backward + regulariser in three kernels
    plus the weight-gradient tiles) + ``pmb_clip_adam_step`` on the optimiser's own state; the minibatch indices
    come from numpy and the dropout noise from torch's generator in the reference's order
    (train_regressor.py:14-22, models/modules.py:102-118,135-139), so seeded runs consume the same random
    streams as the reference;
  * the module loop (everything else: CPU models, prioritized sampling, decoupled regulariser, other optimisers,
    or ``PROB_MBRL_BACKEND=eager``): the reference's loop on the modules' own forward.
"""
import ctypes as C
import os

import numpy as np
import torch
from tqdm import tqdm

from . import _lib
from .operands import NotEligible, _is_cdropout, _is_diag_gaussian, _is_linear, _is_relu
from .rollout import backend


def iterate_minibatches(inputs, targets, batchsize):
    """Endless stream of shuffled minibatches (reference utils/train_regressor.py:14-22; numpy global RNG)."""
    assert len(inputs) == len(targets)
    N = len(inputs)
    while True:
        order = np.arange(0, max(N, batchsize)) % N
        np.random.shuffle(order)
        for i in range(0, len(inputs), batchsize):
            idx = order[i:i + batchsize]
            yield inputs[idx], targets[idx], idx


def gaussian_log_likelihood(targets, means, log_stds=None):
    """Diagonal-Gaussian log-likelihood per row (reference losses.py:16-36)."""
    D = means.shape[-1]
    deltas = means - targets
    if log_stds is None:
        return -0.5 * (deltas ** 2).sum(-1) - D * 0.5 * np.log(2 * np.pi)
    lml = -0.5 * ((deltas * log_stds.exp().reciprocal()) ** 2).sum(-1) - log_stds.sum(-1) - D * 0.5 * np.log(2 * np.pi)
    return lml


def _read_fit_net(model):
    """(linears, dropouts) of a Regressor whose net is (Linear, ReLU, CDropout) x L, Linear."""
    children = list(model.model._modules.values())
    lin, drop = [], []
    i = 0
    while i < len(children):
        if not _is_linear(children[i]):
            raise NotEligible("dynamics.model: unsupported layer %s" % type(children[i]).__name__)
        lin.append(children[i])
        i += 1
        if i == len(children):
            break
        if not (_is_relu(children[i]) and i + 1 < len(children) and _is_cdropout(children[i + 1])):
            raise NotEligible("dynamics.model: the fused fit needs (Linear, ReLU, CDropout) blocks")
        drop.append(children[i + 1])
        i += 2
    if len(lin) != len(drop) + 1 or not drop:
        raise NotEligible("dynamics.model: the fused fit needs hidden layers followed by one output projection")
    if not _is_diag_gaussian(getattr(model, "output_density", None)):
        raise NotEligible("dynamics.output_density must be DiagGaussianDensity")
    return lin, drop


class FusedFit:
    """Device-resident state of the fused fit for one (model, optimiser, batch size)."""

    def __init__(self, model, opt, M, reg_weight):
        self.lib = _lib.load()
        self.model, self.opt, self.M = model, opt, int(M)
        self.lin, self.drop = _read_fit_net(model)
        dev = self.lin[0].weight.device
        self.dev = dev
        L = len(self.drop)
        self.params = []
        for l in range(L):
            self.params += [self.lin[l].weight, self.lin[l].bias, self.drop[l].logit_p]
        self.params += [self.lin[L].weight, self.lin[L].bias]
        if any(p is None or not p.requires_grad or p.dtype != torch.float32 for p in self.params):
            raise NotEligible("the fused fit needs float32 trainable weights, biases and dropout logits")
        want = [p for p in model.parameters() if p.requires_grad]
        if [id(p) for p in want] != [id(p) for p in self.params]:
            raise NotEligible("model.parameters() is not (fc.weight, fc.bias, drop.logit_p) x L, fc_out.weight, fc_out.bias")
        g = opt.param_groups[0]
        if (type(opt) is not torch.optim.Adam or len(opt.param_groups) != 1 or g.get("weight_decay", 0) != 0
                or g.get("amsgrad", False) or g.get("maximize", False) or g.get("capturable", False) or g.get("fused", None)
                or torch.is_tensor(g["lr"])
                or [id(p) for p in g["params"] if p.requires_grad] != [id(p) for p in self.params]):
            raise NotEligible("the fused fit drives a plain torch.optim.Adam over model.parameters()")
        self.X = ((model.X - model.mx) * model.iSx).detach().float().contiguous()       # train_regressor.py:75-76
        self.Y = ((model.Y - model.my) * model.iSy).detach().float().contiguous()
        f32 = dict(device=dev, dtype=torch.float32)
        self.u = [torch.empty(self.M, d.logit_p.numel(), **f32) for d in self.drop]
        self.hard = [torch.empty_like(u) for u in self.u]
        self.mask = [torch.empty_like(u) for u in self.u]
        self.keep_p = [torch.empty_like(d.logit_p.detach()) for d in self.drop]
        self.idx = torch.empty(self.M, dtype=torch.int64, device=dev)
        self.loglik = torch.zeros(1, **f32)
        self.scratch = torch.zeros(1024, **f32)
        p = _lib.PmbFitProblem()
        p.N, p.M = int(self.X.shape[0]), self.M
        net = p.net
        net.n_linear = L + 1
        net.dims[0] = self.lin[0].weight.shape[1]
        for l, fc in enumerate(self.lin):
            net.dims[l + 1] = fc.weight.shape[0]
            net.W[l] = fc.weight.data_ptr()
            net.b[l] = fc.bias.data_ptr()
        net.max_log_std = float(model.output_density.max_log_std)
        for l, d in enumerate(self.drop):
            p.logit_p[l] = d.logit_p.data_ptr()
            p.u[l], p.hard[l], p.mask_out[l] = self.u[l].data_ptr(), self.hard[l].data_ptr(), self.mask[l].data_ptr()
            p.p_out[l] = self.keep_p[l].data_ptr()
            p.temp[l] = float(d.temp)
            p.reg_scale[l] = float(d.regularizer_scale)
            p.drop_reg[l] = float(d.dropout_regularizer)
        p.reg_weight = float(reg_weight)
        p.Xw, p.Yw = self.X.data_ptr(), self.Y.data_ptr()
        self.prob = p
        self.nbytes = int(self.lib.pmb_fit_workspace_bytes(C.byref(p)))
        self.nparam = int(self.lib.pmb_fit_param_count(C.byref(p)))
        if self.nbytes == 0 or self.nparam != sum(q.numel() for q in self.params):
            raise NotEligible("fit descriptor rejected: %s" % self.lib.pmb_fit_last_error().decode())
        self.ws = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)
        self.grad_flat = torch.zeros(self.nparam, **f32)
        self.grad_views, off = [], 0
        for q in self.params:
            self.grad_views.append(self.grad_flat[off:off + q.numel()].view_as(q))
            off += q.numel()
        entries = (_lib.PmbAdamTensor * len(self.params))()
        for i, q in enumerate(self.params):
            st = opt.state[q]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(q, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(q, memory_format=torch.preserve_format)
            entries[i].param, entries[i].grad = q.data_ptr(), self.grad_views[i].data_ptr()
            entries[i].exp_avg, entries[i].exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            entries[i].n = q.numel()
        self.adam_table = torch.frombuffer(bytearray(bytes(entries)), dtype=torch.uint8).clone().to(dev)
        self.step_no = int(float(opt.state[self.params[0]]["step"]))
        self.lr, (self.b1, self.b2), self.eps = float(g["lr"]), g["betas"], float(g["eps"])

    def draw_noise(self, m):
        """Per dropout layer, in forward order: u = rand_like, probs, b = bernoulli(probs) -- the draws of a train-mode
        CDropout.forward(resample=True) on an [m, h] batch (reference models/modules.py:135-139,102-118), from
        torch's generator."""
        for l, d in enumerate(self.drop):
            u = torch.rand_like(self.u[l][:m])
            probs = ((d.logit_p.detach() + ((u + 1e-7) / (1 - (u - 1e-7))).log()) / d.temp).sigmoid()
            self.u[l][:m].copy_(u)
            self.hard[l][:m].copy_(torch.bernoulli(probs))

    def step(self, idx, noise=None):
        """One iteration on the minibatch rows `idx` (numpy int array, <= M rows); returns the mean log-likelihood
        tensor.  `noise` = per dropout layer (u, b): use these instead of drawing (parity tests replay the
        reference's recorded draws)."""
        m = len(idx)
        if m > self.M:
            raise ValueError("minibatch of %d rows, engine built for %d" % (m, self.M))
        self.prob.M = m
        self.last_m = m
        self.idx[:m].copy_(torch.as_tensor(np.asarray(idx), dtype=torch.int64), non_blocking=True)
        if noise is None:
            self.draw_noise(m)
        else:
            for l, (u, b) in enumerate(noise):
                self.u[l][:m].copy_(u)
                self.hard[l][:m].copy_(b)
        st = _lib.current_stream_ptr()
        rc = self.lib.pmb_fit_gradient(C.byref(self.prob), self.idx.data_ptr(), self.grad_flat.data_ptr(),
                                       self.loglik.data_ptr(), self.ws.data_ptr(), self.nbytes, st)
        if rc != 0:
            raise _lib.LibraryError(rc, self.lib.pmb_fit_last_error().decode("utf-8", "replace"))
        self.step_no += 1
        _lib.check(self.lib.pmb_clip_adam_step(self.adam_table.data_ptr(), len(self.params), 0.0, self.lr, self.b1,
                                               self.b2, self.eps, self.step_no, None, self.scratch.data_ptr(), None, st))
        for q, gv in zip(self.params, self.grad_views):
            q.grad = gv
            self.opt.state[q]["step"] += 1
        return self.loglik

    def finish(self):
        """Leave the modules' buffers as the reference's last train-mode forward would (modules.py:114-118)."""
        for l, d in enumerate(self.drop):
            d.concrete_noise = self.mask[l][:getattr(self, "last_m", self.M)].clone()
            d.p = self.keep_p[l].clone()


def train_regressor(model, iters=2000, batchsize=100, resample=True, optimizer=None,
                    log_likelihood=gaussian_log_likelihood, reg_weight=1.0, pbar_class=tqdm, summary_writer=None,
                    summary_scope='', decoupled_reg=False, prioritized_sampling=False, priority_eps=1e-3,
                    priority_alpha=0.6):
    """Fit ``model`` (a Regressor / DynamicsModel with a dataset set by ``set_dataset``) by minibatch MLE.
    Signature and semantics of reference utils/train_regressor.py:58-165."""
    model.train()
    N, M = model.X.shape[0], batchsize
    print('train_regressor >', 'Dataset size [%d]' % int(N))
    if optimizer is None:
        optimizer = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), 1e-4)
    mode = backend()
    fit = None
    density_ll = getattr(getattr(model, "output_density", None), "log_prob", None)
    same_ll = log_likelihood is gaussian_log_likelihood or (
        density_ll is not None and getattr(log_likelihood, "__func__", None) is getattr(density_ll, "__func__", object()))
    if mode != "eager" and model.X.is_cuda:
        try:
            if decoupled_reg or prioritized_sampling or not resample or not same_ll:
                raise NotEligible("prioritized sampling / decoupled regulariser / custom likelihood / frozen masks "
                                  "run on the module loop")
            fit = FusedFit(model, optimizer, min(M, max(N, M)), reg_weight)
        except NotEligible:
            if mode == "fused":
                raise
            fit = None
    elif mode == "fused" and not model.X.is_cuda:
        raise NotEligible("the fused fit needs a CUDA model; set PROB_MBRL_BACKEND=eager (or auto) for the module loop")

    if fit is None and (decoupled_reg or prioritized_sampling):
        raise NotImplementedError("decoupled_reg / prioritized_sampling are outside this package's scope")

    if fit is not None:
        # indices only (same numpy stream as iterate_minibatches); the log-likelihood of every iteration stays on the
        # device and is read back for the progress line every PMB_FIT_PBAR_EVERY iterations (the reference syncs every
        # iteration for it, train_regressor.py:143) and once at the end for the summary writer
        rows = np.empty((N, 0))
        stream = iterate_minibatches(rows, rows, M)
        trace = torch.zeros(iters + 1, device=model.X.device)
        every = max(1, int(os.environ.get("PMB_FIT_PBAR_EVERY", "50")))
        pbar = pbar_class(range(iters + 1), total=iters)
        for i in pbar:
            idx = next(stream)[2]
            trace[i] = fit.step(idx)[0]
            if i % every == 0 or i == iters:
                pbar.set_description('log-likelihood of data: %f' % float(trace[i]))
            if i == iters:
                pbar.close() if hasattr(pbar, "close") else None
                break
        if summary_writer is not None:
            scope = (summary_scope + '/') if summary_scope else ''
            for i, v in enumerate(trace.tolist()):
                summary_writer.add_scalar(scope + 'E_lml', v, i)
        fit.finish()
        model.eval()
        print(model)
        return
    X = (model.X - model.mx) * model.iSx
    Y = (model.Y - model.my) * model.iSy
    pbar = pbar_class(enumerate(iterate_minibatches(X, Y, M)), total=iters)
    for i, (x, y, idx) in pbar:
        model.zero_grad()
        outs = model(x, normalize=False, resample=resample)
        log_probs = log_likelihood(y, *outs)
        Enlml = -log_probs.mean()
        reg = reg_weight * model.regularization_loss()
        loss = Enlml + reg / N
        loss.backward()
        optimizer.step()
        pbar.set_description('log-likelihood of data: %f' % (-Enlml))
        if summary_writer is not None:
            scope = (summary_scope + '/') if summary_scope else ''
            summary_writer.add_scalar(scope + 'training_loss', loss, i)
            summary_writer.add_scalar(scope + 'E_lml', -Enlml, i)
            summary_writer.add_scalar(scope + 'reg_loss', reg, i)
        if i == iters:
            pbar.close()
            break
    model.eval()
    print(model)
