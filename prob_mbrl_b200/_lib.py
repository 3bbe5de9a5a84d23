"""ctypes binding of libpmb_b200.so (C ABI: include/pmb_b200.h).

There is no CPU fallback behind this module: if the library is missing or a call fails, the
product path raises.  PyTorch only supplies device memory and the stream.
"""
import ctypes as C
import os

import torch

from .operands import RolloutOperands

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libpmb_b200.so")

PMB_MAX_LINEAR = 6
PMB_MAX_WIDTH = 1024
PMB_MAX_STATE = 16
PMB_MAX_REWARD_ROWS = 16
ABI_VERSION = 6

_fp = C.POINTER(C.c_float)


class PmbNet(C.Structure):
    _fields_ = [
        ("n_linear", C.c_int),
        ("dims", C.c_int * (PMB_MAX_LINEAR + 1)),
        ("W", C.c_void_p * PMB_MAX_LINEAR),
        ("b", C.c_void_p * PMB_MAX_LINEAR),
        ("mask", C.c_void_p * PMB_MAX_LINEAR),
        ("keep", C.c_float * PMB_MAX_LINEAR),
        ("has_density", C.c_int),
        ("z", C.c_void_p),
        ("z_step_stride", C.c_longlong),
        ("max_log_std", C.c_float),
    ]


class PmbProblem(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("H", C.c_int), ("D", C.c_int), ("U", C.c_int),
        ("pol", PmbNet), ("dyn", PmbNet),
        ("act_scale", C.c_void_p), ("act_bias", C.c_void_p),
        ("mx", C.c_void_p), ("iSx", C.c_void_p), ("my", C.c_void_p), ("Sy", C.c_void_p),
        ("rew_rows", C.c_int),
        ("rew_C", C.c_void_p), ("rew_c0", C.c_void_p), ("rew_Q", C.c_void_p), ("rew_R", C.c_void_p),
        ("rew_scale", C.c_float), ("rew_offset", C.c_float),
        ("mm_states", C.c_int), ("mm_rewards", C.c_int), ("mm_groups", C.c_int),
        ("z_mm", C.c_void_p), ("z_rr", C.c_void_p),
        ("n_global", C.c_int),
        ("masks_binary", C.c_int),
        ("mm_world", C.c_int),
        ("mm_rank", C.c_int),
        ("mm_peer_rec", C.c_void_p * 16),
        ("mm_peer_gather", C.c_void_p * 16),
        ("mm_local_state", C.c_void_p),
    ]


class PmbTuning(C.Structure):
    _fields_ = [
        ("particles_per_cta", C.c_int), ("stream_mode", C.c_int), ("wgrad_splits", C.c_int),
        ("reserved", C.c_int * 5),
    ]


class PmbAdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p),
                ("exp_avg_sq", C.c_void_p), ("n", C.c_longlong)]


class PmbFitProblem(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("M", C.c_int),
        ("net", PmbNet),
        ("logit_p", C.c_void_p * PMB_MAX_LINEAR), ("u", C.c_void_p * PMB_MAX_LINEAR), ("hard", C.c_void_p * PMB_MAX_LINEAR),
        ("temp", C.c_float * PMB_MAX_LINEAR), ("reg_scale", C.c_float * PMB_MAX_LINEAR), ("drop_reg", C.c_float * PMB_MAX_LINEAR),
        ("reg_weight", C.c_float),
        ("Xw", C.c_void_p), ("Yw", C.c_void_p),
        ("mask_out", C.c_void_p * PMB_MAX_LINEAR),
        ("p_out", C.c_void_p * PMB_MAX_LINEAR),
    ]


class PmbPlanInfo(C.Structure):
    _fields_ = [("variant", C.c_int), ("ctas", C.c_int), ("threads_per_cta", C.c_int), ("cluster_size", C.c_int),
                ("particles_per_group", C.c_int), ("smem_fwd_bytes", C.c_int), ("smem_bwd_bytes", C.c_int),
                ("launches_fwd", C.c_int), ("launches_bwd", C.c_int)]


EXPORTS = ("pmb_abi_version", "pmb_last_error", "pmb_check_problem", "pmb_workspace_bytes", "pmb_policy_param_count",
           "pmb_plan_describe",
           "pmb_rollout_forward", "pmb_rollout_backward", "pmb_clip_adam_step",
           "pmb_fit_workspace_bytes", "pmb_fit_param_count", "pmb_fit_last_error", "pmb_fit_gradient",
           "pmb_peer_last_error", "pmb_peer_buffer_bytes", "pmb_peer_alloc", "pmb_peer_open", "pmb_peer_close",
           "pmb_peer_free", "pmb_peer_allreduce", "pmb_mm_exchange_bytes")

_lib = None


class LibraryMissing(RuntimeError):
    """libpmb_b200.so is absent / has the wrong ABI.  Never a numerical failure: mc_pilco re-raises it."""


class LibraryError(RuntimeError):
    """A library call returned a PMB_E_* code (bad descriptor, workspace, CUDA runtime failure).
    Distinct from the numerical failures the reference's ``except RuntimeError`` is meant for
    (algorithms/mc_pilco.py:122-131): mc_pilco re-raises it instead of skipping the iteration."""

    def __init__(self, code, msg):
        super().__init__("libpmb_b200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """Load the CUDA library; raise (never fall back) when it is absent or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            "prob_mbrl_b200: %s not found -- build it with `python -m prob_mbrl_b200.build` "
            "(or __graft_entry__.build()); there is no CPU fallback for the fused rollout" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise LibraryMissing("%s does not export %s" % (LIB_PATH, name))
    lib.pmb_abi_version.restype = C.c_int
    lib.pmb_last_error.restype = C.c_char_p
    lib.pmb_check_problem.restype = C.c_int
    lib.pmb_check_problem.argtypes = [C.POINTER(PmbProblem), C.POINTER(PmbTuning)]
    lib.pmb_workspace_bytes.restype = C.c_size_t
    lib.pmb_workspace_bytes.argtypes = [C.POINTER(PmbProblem), C.POINTER(PmbTuning)]
    lib.pmb_plan_describe.restype = C.c_int
    lib.pmb_plan_describe.argtypes = [C.POINTER(PmbProblem), C.POINTER(PmbTuning), C.POINTER(PmbPlanInfo)]
    lib.pmb_policy_param_count.restype = C.c_size_t
    lib.pmb_policy_param_count.argtypes = [C.POINTER(PmbProblem)]
    lib.pmb_rollout_forward.restype = C.c_int
    lib.pmb_rollout_forward.argtypes = [C.POINTER(PmbProblem), C.POINTER(PmbTuning), C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.pmb_rollout_backward.restype = C.c_int
    lib.pmb_rollout_backward.argtypes = [C.POINTER(PmbProblem), C.POINTER(PmbTuning)] + [C.c_void_p] * 10 + \
                                        [C.c_size_t, C.c_void_p]
    lib.pmb_clip_adam_step.restype = C.c_int
    lib.pmb_clip_adam_step.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pmb_fit_workspace_bytes.restype = C.c_size_t
    lib.pmb_fit_workspace_bytes.argtypes = [C.POINTER(PmbFitProblem)]
    lib.pmb_fit_param_count.restype = C.c_size_t
    lib.pmb_fit_param_count.argtypes = [C.POINTER(PmbFitProblem)]
    lib.pmb_peer_last_error.restype = C.c_char_p
    lib.pmb_peer_buffer_bytes.restype = C.c_size_t
    lib.pmb_peer_buffer_bytes.argtypes = [C.c_longlong, C.c_int]
    lib.pmb_peer_alloc.restype = C.c_int
    lib.pmb_peer_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
    lib.pmb_peer_open.restype = C.c_int
    lib.pmb_peer_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.pmb_peer_close.restype = C.c_int
    lib.pmb_peer_close.argtypes = [C.c_void_p]
    lib.pmb_peer_free.restype = C.c_int
    lib.pmb_peer_free.argtypes = [C.c_void_p]
    lib.pmb_peer_allreduce.restype = C.c_int
    lib.pmb_peer_allreduce.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                       C.c_void_p, C.c_void_p]
    lib.pmb_mm_exchange_bytes.restype = C.c_int
    lib.pmb_mm_exchange_bytes.argtypes = [C.POINTER(PmbProblem), C.POINTER(PmbTuning), C.POINTER(C.c_size_t)]
    lib.pmb_fit_last_error.restype = C.c_char_p
    lib.pmb_fit_gradient.restype = C.c_int
    lib.pmb_fit_gradient.argtypes = [C.POINTER(PmbFitProblem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                     C.c_void_p]
    if lib.pmb_abi_version() != ABI_VERSION:
        raise LibraryMissing("ABI version mismatch: library %d, binding %d" % (lib.pmb_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().pmb_last_error().decode("utf-8", "replace")
        raise LibraryError(rc, msg)


PMB_E_UNSUPPORTED = -2


def check_problem(prob, tune):
    """Validate a descriptor; shapes outside the fused scope raise NotEligible, the rest RuntimeError."""
    from .operands import NotEligible
    lib = load()
    rc = lib.pmb_check_problem(C.byref(prob), C.byref(tune))
    if rc == PMB_E_UNSUPPORTED:
        raise NotEligible(lib.pmb_last_error().decode("utf-8", "replace"))
    check(rc)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t, what):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise LibraryError(-1, "%s must be a CUDA float32 tensor (got %s on %s)" % (what, t.dtype, t.device))
    return t if t.is_contiguous() else t.contiguous()


def _fill_net(dst, net, n_rows, out_dims, keepalive):
    L = len(net.W) - 1
    if L + 1 > PMB_MAX_LINEAR:
        raise LibraryError(-2, "network has %d linear layers, the fused path supports %d" % (L + 1, PMB_MAX_LINEAR))
    dst.n_linear = L + 1
    dst.dims[0] = net.W[0].shape[1]
    for i, w in enumerate(net.W):
        w = _f32c(w.detach(), "weight")
        keepalive.append(w)
        dst.dims[i + 1] = w.shape[0]
        dst.W[i] = w.data_ptr()
        if net.b[i] is not None:
            b = _f32c(net.b[i].detach(), "bias")
            keepalive.append(b)
            dst.b[i] = b.data_ptr()
        else:
            dst.b[i] = None
    for i in range(L):
        if net.mask[i] is not None:
            m = _f32c(net.mask[i], "dropout mask")
            if m.shape[0] < n_rows or m.shape[1] != net.W[i].shape[0]:
                raise LibraryError(-1, "dropout mask %d has shape %s" % (i, tuple(m.shape)))
            keepalive.append(m)
            dst.mask[i] = m.data_ptr()
        else:
            dst.mask[i] = None
        dst.keep[i] = float(net.p[i])
    dst.has_density = int(net.has_density)
    if net.has_density:
        z = _f32c(net.z, "density noise")
        keepalive.append(z)
        dst.z = z.data_ptr()
        if z.dim() == 3:
            dst.z_step_stride = z.shape[1] * z.shape[2]
        else:
            dst.z_step_stride = 0
        if z.shape[-1] != out_dims or z.shape[-2] < n_rows:
            raise LibraryError(-1, "density noise has shape %s" % (tuple(z.shape),))
        if z.dim() == 2 and z.shape[0] != n_rows:
            # rows beyond N are never read, but the row stride must be out_dims: fine as is
            pass
    else:
        dst.z = None
        dst.z_step_stride = 0
    dst.max_log_std = float(net.lmax)


def make_problem(ops: RolloutOperands, N, H, mm_states=False, mm_rewards=False, mm_groups=None,
                 z_mm=None, z_rr=None, shard=None):
    """Operand bundle -> (pmb_problem, keepalive list).  Tensors in `keepalive` must outlive the call.
    shard = (rank, world): N is this rank's equal share of a batch of N * world particles that is moment-matched as a
    whole (the exchange areas are bound later, FusedIteration._bind_exchange)."""
    keep = []
    p = PmbProblem()
    p.N, p.H, p.D, p.U = int(N), int(H), int(ops.D), int(ops.U)
    _fill_net(p.pol, ops.pol, N, ops.U, keep)
    _fill_net(p.dyn, ops.dyn, N, ops.D, keep)
    for name in ("act_scale", "act_bias", "mx", "iSx", "my", "Sy"):
        t = _f32c(getattr(ops, name), name)
        keep.append(t)
        setattr(p, name, t.data_ptr())
    p.rew_rows = int(ops.rew.C.shape[0])
    for name in ("C", "c0", "Q", "R"):
        t = _f32c(getattr(ops.rew, name), "reward " + name)
        keep.append(t)
        setattr(p, "rew_" + name, t.data_ptr())
    p.rew_scale, p.rew_offset = float(ops.rew.scale), float(ops.rew.offset)
    p.mm_states, p.mm_rewards = int(bool(mm_states)), int(bool(mm_rewards))
    p.mm_groups = int(mm_groups) if mm_groups else 0
    if mm_states:
        z = _f32c(z_mm, "z_mm")
        keep.append(z)
        p.z_mm = z.data_ptr()
    if mm_rewards:
        z = _f32c(z_rr, "z_rr")
        keep.append(z)
        p.z_rr = z.data_ptr()
    p.n_global = int(N)
    if shard is not None and shard[1] > 1 and (mm_states or mm_rewards):
        p.mm_rank, p.mm_world = int(shard[0]), int(shard[1])
        p.n_global = int(N) * int(shard[1])
    p.masks_binary = int(bool(ops.pol.masks_binary and ops.dyn.masks_binary))
    return p, keep


def make_tuning(particles_per_cta=0, stream_mode=0, wgrad_splits=0, phases=0):
    t = PmbTuning()
    t.reserved[0] = int(phases)
    t.reserved[1] = int(os.environ.get("PMB_STAGES", "0"))
    t.reserved[4] = int(os.environ.get("PMB_WGRAD_UMMA", "0"))
    t.particles_per_cta = int(particles_per_cta or int(os.environ.get("PMB_PARTICLES_PER_CTA", "0")))
    # 0 = auto, 1/2 = streaming sweeps, 3 = cluster-resident FFMA2 sweeps (required), 4 = tensor-core cluster
    # sweeps (required), 5 = wide cluster-resident FFMA2 sweeps (required)
    t.stream_mode = int(stream_mode or int(os.environ.get("PMB_STREAM_MODE", "0")))
    if t.stream_mode == 3:
        # particles per cluster (1..8, 0 = auto) and CTAs per cluster (4 or 8, 0 = 8)
        # PMB_CLUSTER_PINGPONG=0 lets the two particle tiles of a CTA run unsynchronised (default: they alternate on
        # the shared-memory-bound phases)
        pp = os.environ.get("PMB_CLUSTER_PINGPONG")
        t.reserved[1] = (int(os.environ.get("PMB_CLUSTER_PG", "0")) | (int(os.environ.get("PMB_CLUSTER_C", "0")) << 4)
                         | ((0 if pp is None else int(pp) + 1) << 8))
    if t.stream_mode == 4:
        # tensor-core cluster sweeps required; tuning aid: k-blocks per ring stage (bits 0-7), ring stages (bits 8-11)
        t.reserved[1] = (int(os.environ.get("PMB_TC_KBS", "0")) | (int(os.environ.get("PMB_TC_STAGES", "0")) << 8)
                         | (int(os.environ.get("PMB_TC_DBG", "0")) << 12))       # DBG: timing experiments, wrong results
    t.wgrad_splits = int(wgrad_splits or int(os.environ.get("PMB_WGRAD_SPLITS", "0")))
    return t


def describe_plan(prob, tune):
    """Sweep variant and launch geometry the planner chose for `prob` (dict of the pmb_plan_info fields)."""
    from .operands import NotEligible
    info = PmbPlanInfo()
    lib = load()
    rc = lib.pmb_plan_describe(C.byref(prob), C.byref(tune), C.byref(info))
    if rc == PMB_E_UNSUPPORTED:
        raise NotEligible(lib.pmb_last_error().decode("utf-8", "replace"))
    check(rc)
    return {f: getattr(info, f) for f, _ in PmbPlanInfo._fields_}


def current_stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
