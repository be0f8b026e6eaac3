"""Tensor-level wrappers over the C ABI (one function per entry point of include/mmer.h).

Every function takes CUDA torch tensors, passes raw device pointers plus the current
stream, and raises on error.  PyTorch is only the allocator and the stream provider here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import BF16, F32, GemmArgs, call


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"unsupported activation dtype {t.dtype} (float32 or bfloat16)")


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.MmerError("mmer_b200 ops need CUDA tensors; there is no CPU fallback")
    if not t.is_contiguous():
        raise _lib.MmerError("mmer_b200 ops need contiguous tensors")
    _lib.bind_device(t.device.index)
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t: Optional[torch.Tensor]):
    if t is not None and t.dtype != torch.float32:
        raise TypeError("expected a float32 tensor")
    return _p(t)


def gemm(A: torch.Tensor, B: torch.Tensor, *, M: int, N: int, K: int, a_major=_lib.MAJOR_K, b_major=_lib.MAJOR_K,
         bias=None, residual=None, gate=None, gate_scale=1.0, relu=False, drop_p=0.0, seed=0, site=0,
         out: Optional[torch.Tensor] = None, out_dtype=None, accumulate=False,
         a_rowsum: Optional[torch.Tensor] = None, relu_mask_out: Optional[torch.Tensor] = None,
         gate_bits: Optional[torch.Tensor] = None, d_colsum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """D[M,N] = epilogue(A[M,K] . B[N,K]^T); see mmer_gemm in include/mmer.h."""
    for t in (A, B, bias, residual, gate, out):
        if t is not None and not t.is_cuda:
            raise _lib.MmerError("mmer_b200 ops need CUDA tensors; there is no CPU fallback")
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=out_dtype or A.dtype)
    a = GemmArgs()
    a.A, a.B, a.D = A.data_ptr(), B.data_ptr(), out.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    a.residual = residual.data_ptr() if residual is not None else None
    a.gate = gate.data_ptr() if gate is not None else None
    a.M, a.N, a.K = M, N, K
    a.lda = A.stride(0)
    a.ldb = B.stride(0)
    a.ldd = out.stride(0)
    a.a_major, a.b_major = a_major, b_major
    a.in_dtype, a.out_dtype = _dt(A), _dt(out)
    a.accumulate, a.relu = int(accumulate), int(relu)
    if a_rowsum is not None:
        if a_rowsum.dtype != torch.float32 or not a_rowsum.is_cuda:
            raise TypeError("a_rowsum must be a CUDA float32 tensor")
        a.a_rowsum = a_rowsum.data_ptr()
    for name, t in (("relu_mask_out", relu_mask_out), ("gate_bits", gate_bits)):
        if t is not None:
            if t.dtype != torch.uint8 or not t.is_cuda or t.numel() != M * N // 8:
                raise TypeError(f"{name} must be a CUDA uint8 tensor of M*N/8 bytes")
            setattr(a, name, t.data_ptr())
    if d_colsum is not None:
        if d_colsum.dtype != torch.float32 or not d_colsum.is_cuda or d_colsum.numel() != N:
            raise TypeError("d_colsum must be a CUDA float32 tensor of N elements")
        a.d_colsum = d_colsum.data_ptr()
    a.drop_p, a.gate_scale, a.seed, a.drop_site = float(drop_p), float(gate_scale), int(seed), int(site)
    call("mmer_gemm", C.byref(a), _stream())
    return out


def linear_fwd(x, w, bias=None, relu=False, drop_p=0.0, seed=0, site=0):
    M, K = x.shape
    return gemm(x, w, M=M, N=w.shape[0], K=K, bias=bias, relu=relu, drop_p=drop_p, seed=seed, site=site)


def linear_dgrad(dy, w, residual=None, gate=None, gate_scale=1.0):
    M, N = dy.shape
    return gemm(dy, w, M=M, N=w.shape[1], K=N, b_major=_lib.MAJOR_MN, residual=residual, gate=gate,
                gate_scale=gate_scale)


def linear_wgrad(dy, x, out: torch.Tensor, dbias: Optional[torch.Tensor] = None):
    """out[N,K] (fp32) += dy[M,N]^T x[M,K];  dbias[N] (fp32, optional) += column sums of dy, from the same kernel"""
    M, N = dy.shape
    return gemm(dy, x, M=N, N=x.shape[1], K=M, a_major=_lib.MAJOR_MN, b_major=_lib.MAJOR_MN, out=out, accumulate=True,
                a_rowsum=dbias)


def embed_fwd(pv, pa, gv, bv, ga, ba, pos, B, T, drop_p=0.0, seed=0, site=0):
    F = pv.shape[-1]
    x0 = torch.empty((B * (T + 1), F), device=pv.device, dtype=pv.dtype)
    stats = torch.empty((B * (T + 1), 2), device=pv.device, dtype=torch.float32)
    call("mmer_embed_fwd", _p(pv), _p(pa), _f32(gv), _f32(bv), _f32(ga), _f32(ba), _f32(pos), _p(x0), _p(stats), B, T, F,
         _dt(pv), drop_p, seed, site, _stream())
    return x0, stats


def embed_bwd(dx0, pv, pa, stats, gv, ga, B, T, dgv, dbv, dga, dba, dpos, drop_p=0.0, seed=0, site=0, dbias_v=None,
              dbias_a=None):
    """dbias_v / dbias_a (optional fp32 [F]) accumulate the column sums of dpv / dpa (input-projection bias gradients)."""
    F = pv.shape[-1]
    dpv, dpa = torch.empty_like(pv), torch.empty_like(pa)
    call("mmer_embed_bwd", _p(dx0), _p(pv), _p(pa), _f32(stats), _f32(gv), _f32(ga), _p(dpv), _p(dpa), _f32(dgv),
         _f32(dbv), _f32(dga), _f32(dba), _f32(dpos), _f32(dbias_v), _f32(dbias_a), B, T, F, _dt(pv), drop_p, seed, site,
         _stream())
    return dpv, dpa


def add_ln_fwd(x, a, gamma, beta, relu=False, drop_a_p=0.0, site_a=0, drop_y_p=0.0, site_y=0, seed=0):
    M, F = a.shape
    y = torch.empty_like(a)
    stats = torch.empty((M, 2), device=a.device, dtype=torch.float32)
    call("mmer_add_ln_fwd", _p(x), _p(a), _f32(gamma), _f32(beta), _p(y), _p(stats), M, F, _dt(a), int(relu), drop_a_p,
         site_a, drop_y_p, site_y, seed, _stream())
    return y, stats


def add_ln_bwd(dy, x, a, stats, gamma, beta, dgamma, dbeta, dbias=None, relu=False, drop_a_p=0.0, site_a=0,
               drop_y_p=0.0, site_y=0, seed=0):
    M, F = a.shape
    dz = torch.empty_like(a)
    da = torch.empty_like(a) if drop_a_p > 0 else None
    call("mmer_add_ln_bwd", _p(dy), _p(x), _p(a), _f32(stats), _f32(gamma), _f32(beta), _p(dz), _p(da), _f32(dgamma),
         _f32(dbeta), _f32(dbias), M, F, _dt(a), int(relu), drop_a_p, site_a, drop_y_p, site_y, seed, _stream())
    return dz, da


def add_ln_bwd_z(dy, z, stats, gamma, dgamma, dbeta, dbias=None, drop_a_p=0.0, site_a=0, seed=0):
    """Backward of y = LN(z) given the stored z = x + dropout(a) (see gemm_ln_fwd): returns (dz, da)."""
    M, F = z.shape
    dz = torch.empty_like(z)
    da = torch.empty_like(z) if drop_a_p > 0 else None
    call("mmer_add_ln_bwd_z", _p(dy), _p(z), _f32(stats), _f32(gamma), _p(dz), _p(da), _f32(dgamma), _f32(dbeta),
         _f32(dbias), M, F, _dt(z), drop_a_p, site_a, seed, _stream())
    return dz, da


def gemm_ln_fwd(a, w, bias, residual, gamma, beta, drop_p=0.0, site=0, seed=0):
    """(z, y, stats) = fused Linear + bias + dropout + residual + LayerNorm (N = 512, bf16); mmer_gemm_ln_fwd."""
    M, K = a.shape
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or w.shape != (512, K):
        raise TypeError("gemm_ln_fwd: bf16 a [M,K] and w [512,K]")
    z = torch.empty((M, 512), device=a.device, dtype=torch.bfloat16)
    y = torch.empty_like(z)
    stats = torch.empty((M, 2), device=a.device, dtype=torch.float32)
    call("mmer_gemm_ln_fwd", _p(a), _p(w), _f32(bias), _p(residual), _f32(gamma), _f32(beta), _p(z), _p(y), _p(stats), M, K,
         drop_p, site, seed, _stream())
    return z, y, stats


def pool_ln_fwd(x, mask, gamma, beta, B, T):
    F = x.shape[-1]
    pooled = torch.empty((B, F), device=x.device, dtype=torch.float32)
    fused = torch.empty((B, F), device=x.device, dtype=x.dtype)
    stats = torch.empty((B, 2), device=x.device, dtype=torch.float32)
    call("mmer_pool_ln_fwd", _p(x), _p(mask), _f32(gamma), _f32(beta), _p(pooled), _p(fused), _p(stats), B, T, F, _dt(x),
         _stream())
    return fused, pooled, stats


def pool_ln_bwd(dfused, pooled, stats, gamma, mask, B, T, dgamma, dbeta):
    F = dfused.shape[-1]
    dx = torch.empty((B * (T + 1), F), device=dfused.device, dtype=dfused.dtype)
    call("mmer_pool_ln_bwd", _p(dfused), _f32(pooled), _f32(stats), _f32(gamma), _p(mask), _p(dx), _f32(dgamma),
         _f32(dbeta), B, T, F, _dt(dfused), _stream())
    return dx


def colsum(x, out):
    M, N = x.shape
    call("mmer_colsum", _p(x), _f32(out), M, N, x.stride(0), _dt(x), _stream())
    return out


def mha_fwd(qkv, mask, B, T, H, d, want_probs=False, drop_p=0.0, seed=0, site=0):
    S = T + 1
    out = torch.empty((B * S, H * d), device=qkv.device, dtype=qkv.dtype)
    probs = torch.empty((B, H, S, S), device=qkv.device, dtype=torch.float32) if want_probs else None
    call("mmer_mha_fwd", _p(qkv), _p(mask), _p(out), _p(probs), B, T, H, d, _dt(qkv), drop_p, seed, site, _stream())
    return out, probs


def mha_bwd(qkv, mask, dout, B, T, H, d, drop_p=0.0, seed=0, site=0, dbias=None):
    """dbias (optional fp32 [3*H*d]) accumulates the column sums of dqkv (gradient of in_proj_bias)."""
    dqkv = torch.empty_like(qkv)
    call("mmer_mha_bwd", _p(qkv), _p(mask), _p(dout), _p(dqkv), _f32(dbias), B, T, H, d, _dt(qkv), drop_p, seed, site,
         _stream())
    return dqkv


def head_out_fwd(h, W, b):
    B, K = h.shape
    Cn = W.shape[0]
    logits = torch.empty((B, Cn), device=h.device, dtype=torch.float32)
    probs = torch.empty_like(logits)
    call("mmer_head_out_fwd", _p(h), _f32(W), _f32(b), _p(logits), _p(probs), B, K, Cn, _dt(h), _stream())
    return logits, probs


def head_out_bwd(dlogits, h, W, dW, db):
    B, K = h.shape
    dh = torch.empty_like(h)
    call("mmer_head_out_bwd", _f32(dlogits), _p(h), _f32(W), _p(dh), _f32(dW), _f32(db), B, K, W.shape[0], _dt(h),
         _stream())
    return dh


def loss_fwd_bwd(logits, labels, alpha=None, kind=_lib.LOSS_FOCAL, gamma=2.0, reduction=_lib.REDUCE_MEAN,
                 want_grad=True, grad_scale=1.0):
    """Returns (loss scalar tensor [1] or per-sample [B] for REDUCE_NONE, dlogits or None)."""
    B, Cn = logits.shape
    if labels.dtype != torch.int64:
        raise TypeError("labels must be int64")
    loss = torch.empty(1, device=logits.device, dtype=torch.float32)
    per = torch.empty(B, device=logits.device, dtype=torch.float32) if reduction == _lib.REDUCE_NONE else None
    dlogits = torch.empty_like(logits) if want_grad else None
    scratch = torch.empty(2, device=logits.device, dtype=torch.float32)
    call("mmer_loss_fwd_bwd", _f32(logits), _p(labels), _f32(alpha), kind, float(gamma), reduction, _p(loss), _p(per),
         _p(dlogits), _p(scratch), B, Cn, float(grad_scale), _stream())
    return (per if reduction == _lib.REDUCE_NONE else loss), dlogits


def adam_step(p, g, m, v, shadow, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, grad_scale=1.0,
              sumsq=None, max_norm=0.0):
    call("mmer_adam_step", _f32(p), _f32(g), _f32(m), _f32(v), _p(shadow), p.numel(), float(lr), float(beta1),
         float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale), _f32(sumsq), float(max_norm),
         _stream())


def grad_sumsq(g, out=None):
    if out is None:
        out = torch.empty(1, device=g.device, dtype=torch.float32)
    call("mmer_grad_sumsq", _f32(g), g.numel(), _f32(out), _stream())
    return out


def cast_bf16(src, dst=None):
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16)
    call("mmer_cast_bf16", _f32(src), _p(dst), src.numel(), _stream())
    return dst


def cast_f32(src, dst=None):
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=torch.float32)
    call("mmer_cast_f32", _p(src), _f32(dst), src.numel(), _stream())
    return dst


def bn_fwd(x, gamma, beta, running_mean, running_var, training=True, relu=False, momentum=0.1, drop_p=0.0, seed=0,
           site=0):
    N, Cn = x.shape
    y = torch.empty_like(x)
    stats = torch.empty((2, Cn), device=x.device, dtype=torch.float32)
    call("mmer_bn_fwd", _p(x), _f32(gamma), _f32(beta), _f32(running_mean), _f32(running_var), _p(y), _p(stats), N, Cn,
         _dt(x), int(training), int(relu), momentum, drop_p, seed, site, _stream())
    return y, stats


def bn_bwd(dy, x, stats, gamma, beta, dgamma, dbeta, training=True, relu=False, drop_p=0.0, seed=0, site=0):
    N, Cn = x.shape
    dx = torch.empty_like(x)
    scratch = torch.empty(2 * Cn, device=x.device, dtype=torch.float32)
    call("mmer_bn_bwd", _p(dy), _p(x), _f32(stats), _f32(gamma), _f32(beta), _p(dx), _f32(dgamma), _f32(dbeta),
         _p(scratch), N, Cn, _dt(x), int(training), int(relu), drop_p, seed, site, _stream())
    return dx
