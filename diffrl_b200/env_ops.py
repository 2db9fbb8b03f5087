"""Fused env epilogue ops (observation + reward + termination flags of the free-root walkers) as ONE
autograd.Function over the C ABI (``dfx_walker_obs_forward / _backward`` in include/dfx.h)."""
import ctypes

import torch

from . import _capi


class DfxWalkerParams(ctypes.Structure):
    """ctypes mirror of ``DfxWalkerParams`` in include/dfx.h (keep field order in sync)."""

    _fields_ = [
        ("num_q", ctypes.c_int), ("num_qd", ctypes.c_int), ("num_act", ctypes.c_int), ("num_obs", ctypes.c_int),
        ("obs_has_actions", ctypes.c_int), ("height_mode", ctypes.c_int), ("action_penalty_abs", ctypes.c_int),
        ("early_termination", ctypes.c_int), ("check_invalid", ctypes.c_int), ("zero_reward_on_invalid", ctypes.c_int),
        ("episode_length", ctypes.c_int),
        ("joint_vel_scale", ctypes.c_float), ("termination_height", ctypes.c_float),
        ("termination_tolerance", ctypes.c_float), ("height_rew_scale", ctypes.c_float), ("action_penalty", ctypes.c_float),
        ("target", ctypes.c_float * 3), ("inv_start_rot", ctypes.c_float * 4),
        ("basis_heading", ctypes.c_float * 3), ("basis_up", ctypes.c_float * 3),
    ]


def _bind(lib):
    if getattr(lib, "_walker_bound", False):
        return
    P = ctypes.POINTER(DfxWalkerParams)
    V = ctypes.c_void_p
    lib.dfx_walker_obs_forward.argtypes = [P, ctypes.c_int, V, V, V, V, V, V, V, V]
    lib.dfx_walker_obs_backward.argtypes = [P, ctypes.c_int, V, V, V, V, V, V, V, V, V]
    lib._walker_bound = True


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _c(t):
    return t if (t.is_contiguous() and t.dtype == torch.float32) else t.contiguous().float()


class WalkerObsFunction(torch.autograd.Function):
    """(q, qd, actions) -> (obs, rew, reset).  ``want_reward=False`` computes only the observation."""

    @staticmethod
    def forward(ctx, params, n, want_reward, progress, q, qd, actions):
        lib = _capi.lib()
        _bind(lib)
        q, qd, actions = _c(q.detach()), _c(qd.detach()), _c(actions.detach())
        dev = q.device
        obs = torch.empty((n, params.num_obs), dtype=torch.float32, device=dev)
        rew = torch.empty(n, dtype=torch.float32, device=dev) if want_reward else None
        reset = torch.empty(n, dtype=torch.long, device=dev) if want_reward else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            code = lib.dfx_walker_obs_forward(ctypes.byref(params), n, _ptr(q), _ptr(qd), _ptr(actions),
                                              _ptr(progress) if want_reward else None, _ptr(obs), _ptr(rew), _ptr(reset), stream)
        _capi.check(code, "dfx_walker_obs_forward")
        ctx.params, ctx.n = params, n
        ctx.save_for_backward(q, qd, actions)
        ctx.shapes = None
        if want_reward:
            ctx.mark_non_differentiable(reset)
            return obs, rew, reset
        return obs

    @staticmethod
    def backward(ctx, g_obs, g_rew=None, g_reset=None):
        lib = _capi.lib()
        q, qd, actions = ctx.saved_tensors
        dev = q.device
        gq, gqd = torch.empty_like(q), torch.empty_like(qd)
        gact = torch.empty_like(actions) if ctx.needs_input_grad[6] else None
        g_obs = None if g_obs is None else _c(g_obs)
        g_rew = None if g_rew is None else _c(g_rew)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            code = lib.dfx_walker_obs_backward(ctypes.byref(ctx.params), ctx.n, _ptr(q), _ptr(qd), _ptr(actions),
                                               _ptr(g_obs), _ptr(g_rew), _ptr(gq), _ptr(gqd), _ptr(gact), stream)
        _capi.check(code, "dfx_walker_obs_backward")
        return None, None, None, None, gq, gqd, gact
