"""Channels-last GroupNorm (+ channel bias, + SiLU) — autograd binding over dm4d_groupnorm_nhwc_forward / _backward.

The glue between the tensor-core convolutions of the Zero123 networks in the SDS step: one fused
pass replaces torch's NCHW-only GroupNorm, the separate SiLU, the time-embedding broadcast add and the NHWC<->NCHW
transposes cuDNN otherwise inserts around every convolution (ResBlock, openaimodel.py:269-289; ResnetBlock,
model.py:118-138 of the reference).  CUDA only; on other devices ``GroupNormAct`` evaluates the same formula with torch
ops (the CPU parity tests of ``zero123.py`` run that way).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import check, ptr

_DTYPES = {torch.float32: 0, torch.float16: 1}


class _GroupNormNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, chan_bias, gamma, beta, groups, eps, silu):
        N, C, H, W = x.shape
        if x.dtype not in _DTYPES:
            raise TypeError(f"groupnorm_nhwc: fp16 or fp32 activations, got {x.dtype}")
        x = x.contiguous(memory_format=torch.channels_last)
        cb = None if chan_bias is None else chan_bias.detach().float().contiguous()
        f32 = dict(dtype=torch.float32, device=x.device)
        stats = torch.empty(N, groups, 2, **f32)
        scratch = torch.empty(N * groups * 2 + N + 1, **f32)
        y = torch.empty_like(x, memory_format=torch.channels_last)
        check(_lib.lib().dm4d_groupnorm_nhwc_forward(ptr(x), ptr(cb), ptr(gamma), ptr(beta), N, H * W, C, groups, float(eps),
                                                     int(silu), _DTYPES[x.dtype], ptr(stats), ptr(scratch), ptr(y),
                                                     torch.cuda.current_stream().cuda_stream), "dm4d_groupnorm_nhwc_forward")
        ctx.save_for_backward(x, cb, gamma, beta, stats)
        ctx.cfg = (groups, float(eps), int(silu))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, cb, gamma, beta, stats = ctx.saved_tensors
        groups, eps, silu = ctx.cfg
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("groupnorm_nhwc: no gradient for the channel bias (frozen in the SDS step)")
        N, C, H, W = x.shape
        dy = dy.to(x.dtype).contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        scratch = torch.empty(N * groups * 2 + N + 1, dtype=torch.float32, device=x.device)
        check(_lib.lib().dm4d_groupnorm_nhwc_backward(ptr(x), ptr(cb), ptr(dy), ptr(gamma), ptr(beta), N, H * W, C, groups, eps, silu,
                                                      _DTYPES[x.dtype], ptr(stats), ptr(scratch), ptr(dx),
                                                      torch.cuda.current_stream().cuda_stream), "dm4d_groupnorm_nhwc_backward")
        return dx, None, None, None, None, None, None


def groupnorm_nhwc(x: torch.Tensor, gamma32: torch.Tensor, beta32: torch.Tensor, groups: int, eps: float, silu: bool,
                   chan_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``act(GroupNorm(x + chan_bias[:, :, None, None]))`` on a channels-last [N,C,H,W] CUDA tensor (fp16 / fp32);
    ``gamma32`` / ``beta32`` are fp32 [C].  Differentiable w.r.t. ``x`` only."""
    return _GroupNormNHWC.apply(x, chan_bias, gamma32, beta32, groups, eps, silu)


class GroupNormAct(nn.GroupNorm):
    """``nn.GroupNorm`` (same parameters / state_dict keys) followed by an optional SiLU, with an optional per-sample
    channel bias added in front; fused channels-last kernel on CUDA, torch ops elsewhere."""

    def __init__(self, num_groups: int, num_channels: int, eps: float = 1e-5, silu: bool = False):
        super().__init__(num_groups, num_channels, eps=eps, affine=True)
        self.silu = silu
        self._p32 = None

    def _params32(self):
        key = (self.weight._version, self.bias._version, self.weight.data_ptr(), self.weight.device)
        if self._p32 is None or self._p32[0] != key:
            self._p32 = (key, self.weight.detach().float().contiguous(), self.bias.detach().float().contiguous())
        return self._p32[1], self._p32[2]

    def forward(self, x: torch.Tensor, chan_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x.is_cuda and x.dtype in _DTYPES and not (self.weight.requires_grad or self.bias.requires_grad) and \
                self.num_channels % 4 == 0 and self.num_groups <= 64:
            g32, b32 = self._params32()
            return groupnorm_nhwc(x, g32, b32, self.num_groups, self.eps, self.silu, chan_bias)
        if chan_bias is not None:
            x = x + chan_bias.to(x.dtype)[:, :, None, None]
        y = F.group_norm(x, self.num_groups, self.weight, self.bias, self.eps)
        return F.silu(y) if self.silu else y


class _BiasResidualNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, residual, bias32):
        N, C, H, W = h.shape
        h = h.contiguous(memory_format=torch.channels_last)
        r = None if residual is None else residual.to(h.dtype).contiguous(memory_format=torch.channels_last)
        out = torch.empty_like(h, memory_format=torch.channels_last)
        check(_lib.lib().dm4d_bias_residual_add_nhwc(ptr(h), ptr(r), ptr(bias32), N * H * W, C, _DTYPES[h.dtype], ptr(out),
                                                     torch.cuda.current_stream().cuda_stream), "dm4d_bias_residual_add_nhwc")
        ctx.has_res = residual is not None
        return out

    @staticmethod
    def backward(ctx, g):
        return g, (g if ctx.has_res else None), None


def conv_nobias(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """The convolution WITHOUT its bias: PyTorch adds a cuDNN convolution's bias in a separate broadcast pass over the
    output; the callers fold it into the next fused pass instead (GroupNorm's channel bias, or ``bias_residual_add``)."""
    return F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)


def bias_residual_add(h: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``h + bias[None, :, None, None] (+ residual)`` in one channels-last pass (bias frozen: no gradient)."""
    if bias is None:
        return h if residual is None else residual + h
    if h.is_cuda and h.dtype in _DTYPES and h.shape[1] % 4 == 0 and not bias.requires_grad:
        return _BiasResidualNHWC.apply(h, residual, bias.detach().float().contiguous())
    out = h + bias.to(h.dtype)[None, :, None, None]
    return out if residual is None else residual + out


class _GEGLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proj):
        D = proj.shape[-1] // 2
        p2 = proj.contiguous()
        out = torch.empty(*proj.shape[:-1], D, dtype=proj.dtype, device=proj.device)
        check(_lib.lib().dm4d_geglu(ptr(p2), p2.numel() // (2 * D), D, _DTYPES[proj.dtype], ptr(out),
                                    torch.cuda.current_stream().cuda_stream), "dm4d_geglu")
        ctx.save_for_backward(p2)
        return out

    @staticmethod
    def backward(ctx, g):
        (proj,) = ctx.saved_tensors      # the UNet runs without gradient in the SDS step; kept for completeness
        with torch.enable_grad():
            p = proj.detach().requires_grad_(True)
            a, gate = p.chunk(2, dim=-1)
            (a * F.gelu(gate)).backward(g)
        return p.grad


def geglu(proj: torch.Tensor) -> torch.Tensor:
    """``a * gelu(gate)`` for ``a, gate = proj.chunk(2, -1)`` (attention.py:37-65), one pass on CUDA."""
    if proj.is_cuda and proj.dtype in _DTYPES and proj.shape[-1] % 8 == 0:
        return _GEGLU.apply(proj)
    a, gate = proj.chunk(2, dim=-1)
    return a * F.gelu(gate)


def add_layernorm(x: torch.Tensor, delta: Optional[torch.Tensor], ln: nn.LayerNorm):
    """``x_new = x + delta`` (``delta`` [B,N,C] or broadcast [B,1,C]; None: ``x_new = x``) and ``LayerNorm(x_new)`` in one
    pass (``dm4d_add_layernorm``) when no gradient is recorded — the UNet of the SDS step — else with torch ops.
    Returns ``(x_new, y)``."""
    C = x.shape[-1]
    fast = (x.is_cuda and x.dtype in _DTYPES and C % 4 == 0 and C <= 1536 and x.dim() == 3 and
            not (torch.is_grad_enabled() and (x.requires_grad or (delta is not None and delta.requires_grad) or ln.weight.requires_grad)))
    if not fast:
        x_new = x if delta is None else x + delta
        return x_new, F.layer_norm(x_new, (C,), ln.weight, ln.bias, ln.eps)
    B, N, _ = x.shape
    xc = x.contiguous()
    bcast, d = 0, None
    if delta is not None:
        if delta.shape[1] == 1 or delta.stride(1) == 0:
            d, bcast = delta[:, :1].to(x.dtype).contiguous(), N
        else:
            d = delta.to(x.dtype).contiguous()
    cache = getattr(ln, "_dm4d_p32", None)
    key = (ln.weight._version, ln.bias._version, ln.weight.data_ptr(), ln.weight.device)
    if cache is None or cache[0] != key:
        cache = (key, ln.weight.detach().float().contiguous(), ln.bias.detach().float().contiguous())
        ln._dm4d_p32 = cache
    x_new = torch.empty_like(xc) if d is not None else xc
    y = torch.empty_like(xc)
    check(_lib.lib().dm4d_add_layernorm(ptr(xc), ptr(d), bcast, ptr(cache[1]), ptr(cache[2]), B * N, C, float(ln.eps),
                                        _DTYPES[x.dtype], ptr(x_new) if d is not None else None, ptr(y),
                                        torch.cuda.current_stream().cuda_stream), "dm4d_add_layernorm")
    return x_new, y
