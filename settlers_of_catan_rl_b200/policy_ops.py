"""Autograd functions on the two sm_100a kernels the policy network uses where its inner dimensions are tiny
(``csrc/policy_kernels.cu``): the tile encoder's 19-token self-attention and LayerNorm over rows of at most 64 elements.
CUDA only — ``CatanPolicy`` keeps the plain torch ops for CPU tensors (its parity tests against the reference run there)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _p(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class _TileAttention(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, qkv: torch.Tensor) -> torch.Tensor:
        qkv = qkv.contiguous()
        B = qkv.shape[0]
        assert qkv.shape[1:] == (19, 192) and qkv.dtype == torch.float32
        y = torch.empty((B, 19, 64), dtype=torch.float32, device=qkv.device)
        with torch.cuda.device(qkv.device):
            _lib.check(_lib.load().catan_tile_attention_fwd(_p(qkv), _p(y), B, _stream(qkv)))
        ctx.save_for_backward(qkv)
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy: torch.Tensor):
        (qkv,) = ctx.saved_tensors
        dy = dy.contiguous().float()
        dqkv = torch.empty_like(qkv)
        with torch.cuda.device(qkv.device):
            _lib.check(_lib.load().catan_tile_attention_bwd(_p(qkv), _p(dy), _p(dqkv), qkv.shape[0], _stream(qkv)))
        return dqkv


def tile_attention(qkv: torch.Tensor) -> torch.Tensor:
    """qkv [B, 19, 192] (q | k | v; 4 heads of 16) -> [B, 19, 64]: softmax(q k^T / sqrt(16)) v per head, heads concatenated"""
    return _TileAttention.apply(qkv)


class _LayerNormSmall(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float) -> torch.Tensor:
        dim = x.shape[-1]
        x2 = x.contiguous().view(-1, dim)
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        need = x.requires_grad or weight.requires_grad
        stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device) if need else None
        w, b = weight.contiguous(), bias.contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().catan_ln_small_fwd(_p(x2), _p(w), _p(b), _p(y), C.c_void_p(0) if stats is None else _p(stats), rows, dim,
                                                      float(eps), _stream(x2)))
        if need:
            ctx.save_for_backward(x2, w, stats)
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy: torch.Tensor):
        x2, w, stats = ctx.saved_tensors
        dim = x2.shape[1]
        dy2 = dy.contiguous().float().view(-1, dim)
        dx = torch.empty_like(x2)
        dw = torch.zeros(dim, dtype=torch.float32, device=x2.device)
        db = torch.zeros(dim, dtype=torch.float32, device=x2.device)
        with torch.cuda.device(x2.device):
            _lib.check(_lib.load().catan_ln_small_bwd(_p(x2), _p(w), _p(stats), _p(dy2), _p(dx), _p(dw), _p(db), x2.shape[0], dim, _stream(x2)))
        return dx.view(ctx.shape), dw, db, None


def layer_norm_small(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm over a last dimension of at most 64 elements (fp32 result)"""
    return _LayerNormSmall.apply(x, weight, bias, eps)
