"""Thin torch-tensor wrappers over the C ABI (``include/pmgt_b200.h``).

Every function enqueues work on torch's current CUDA stream and returns
immediately; torch is used only for device memory and streams.  bf16 tensors
are passed as raw ``uint16`` storage.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import AttnArgs, EmbedArgs, GemmArgs, GsrArgs, NfrArgs, ResLnArgs, check, cur_stream, ptr

EPI_BIAS, EPI_GELU, EPI_GELU_BWD, EPI_ADDEND, EPI_OUT_F32, EPI_ATOMIC = 1, 2, 4, 8, 16, 32
BF16 = torch.bfloat16


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.PMGTError(f"{name} must be a CUDA tensor: pmgt_b200 has no CPU fallback")


def _num_sms(dev) -> int:
    return torch.cuda.get_device_properties(dev).multi_processor_count


def gemm(a, b, out, *, M, N, K, lda, ldb, ldo, a_mn=False, b_mn=False, a_rows=None, a_src_rows=0, b_rows=None,
         b_src_rows=0, bias=None, addend=None, ld_addend=0, aux=None, ld_aux=0, alpha=1.0, epi=0, split_k=1):
    """``pmgt_gemm_bf16``: D[M,N] (+)= op(A) op(B); see the header for operand layouts."""
    _require_cuda(a, "a")
    g = GemmArgs(M, N, K, ptr(a), lda, int(a_mn), ptr(a_rows), a_src_rows, ptr(b), ldb, int(b_mn), ptr(b_rows),
                 b_src_rows, ptr(out), ldo, ptr(bias), ptr(addend), ld_addend, ptr(aux), ld_aux, alpha, epi, split_k)
    check(_lib.lib().pmgt_gemm_bf16(C.byref(g), cur_stream()), "pmgt_gemm_bf16")


def linear_fwd(x, w, bias, out, *, rows=None, src_rows=0, gelu_aux=None):
    """out[T,N] = x[T,K] @ w[N,K]^T + bias   (x optionally gathered through ``rows``)."""
    N, K = w.shape
    T = out.shape[0]
    epi = EPI_BIAS if bias is not None else 0
    if gelu_aux is not None:
        epi |= EPI_GELU
    gemm(x, w, out, M=T, N=N, K=K, lda=x.stride(0), ldb=w.stride(0), ldo=out.stride(0), a_rows=rows,
         a_src_rows=src_rows, bias=bias, aux=gelu_aux, ld_aux=(gelu_aux.stride(0) if gelu_aux is not None else 0), epi=epi)


def linear_dx(dy, w, out, *, addend=None, gelu_bwd_aux=None):
    """out[T,K] = dy[T,N] @ w[N,K] (+ addend) (* gelu'(aux))."""
    N, K = w.shape
    T = dy.shape[0]
    epi = 0
    if addend is not None:
        epi |= EPI_ADDEND
    if gelu_bwd_aux is not None:
        epi |= EPI_GELU_BWD
    gemm(dy, w, out, M=T, N=K, K=N, lda=dy.stride(0), ldb=w.stride(0), ldo=out.stride(0), b_mn=True,
         addend=addend, ld_addend=(addend.stride(0) if addend is not None else 0),
         aux=gelu_bwd_aux, ld_aux=(gelu_bwd_aux.stride(0) if gelu_bwd_aux is not None else 0), epi=epi)


def linear_dw(dy, x, dw_f32, *, rows=None, src_rows=0, x_cols=None):
    """dw[N,K] += dy[T,N]^T @ x[T,K]  (fp32 atomic accumulation, split over T).

    ``x`` may be a table whose rows are fetched through ``rows`` (T int64 ids)."""
    T, N = dy.shape
    K = x_cols if x_cols is not None else x.shape[1]
    tiles = ((N + 127) // 128) * ((K + 127) // 128)
    num_kb = (T + 63) // 64
    split = max(1, min(num_kb, (2 * _num_sms(dy.device) + tiles - 1) // tiles))
    gemm(dy, x, dw_f32, M=N, N=K, K=T, lda=dy.stride(0), ldb=x.stride(0), ldo=dw_f32.stride(0), a_mn=True, b_mn=True,
         b_rows=rows, b_src_rows=src_rows, epi=EPI_ATOMIC, split_k=split)


def embed_args(rows, L, H, ev, et, w_att, b_att, pos, role, ln_g, ln_b, eps, p, seed, site, **kw):
    a = EmbedArgs()
    a.rows, a.L, a.H = rows, L, H
    a.ev, a.et = ptr(ev), ptr(et)
    a.w_att, a.b_att, a.pos, a.role, a.ln_g, a.ln_b = ptr(w_att), ptr(b_att), ptr(pos), ptr(role), ptr(ln_g), ptr(ln_b)
    a.ln_eps, a.dropout_p, a.dropout_seed, a.dropout_site = eps, p, seed, site
    for k, v in kw.items():
        setattr(a, k, ptr(v))
    return a


def embed_fuse_fwd(a: EmbedArgs):
    check(_lib.lib().pmgt_embed_fuse_fwd(C.byref(a), cur_stream()), "pmgt_embed_fuse_fwd")


def embed_fuse_bwd(a: EmbedArgs):
    check(_lib.lib().pmgt_embed_fuse_bwd(C.byref(a), cur_stream()), "pmgt_embed_fuse_bwd")


def attn_args(rows, L, H, heads, beta, qkvc, mask, p, seed, site, **kw):
    a = AttnArgs()
    a.rows, a.L, a.H, a.heads, a.beta = rows, L, H, heads, beta
    a.qkvc, a.mask = ptr(qkvc), ptr(mask)
    a.dropout_p, a.dropout_seed, a.dropout_site = p, seed, site
    for k, v in kw.items():
        setattr(a, k, ptr(v))
    return a


def attn_core_fwd(a: AttnArgs):
    check(_lib.lib().pmgt_attn_core_fwd(C.byref(a), cur_stream()), "pmgt_attn_core_fwd")


def attn_core_bwd(a: AttnArgs):
    check(_lib.lib().pmgt_attn_core_bwd(C.byref(a), cur_stream()), "pmgt_attn_core_bwd")


def resln_args(T, H, o, res, ln_g, ln_b, eps, p, seed, site, **kw):
    a = ResLnArgs()
    a.T, a.H = T, H
    a.o, a.res, a.ln_g, a.ln_b = ptr(o), ptr(res), ptr(ln_g), ptr(ln_b)
    a.ln_eps, a.dropout_p, a.dropout_seed, a.dropout_site = eps, p, seed, site
    for k, v in kw.items():
        setattr(a, k, ptr(v))
    return a


def res_ln_fwd(a: ResLnArgs):
    check(_lib.lib().pmgt_res_ln_fwd(C.byref(a), cur_stream()), "pmgt_res_ln_fwd")


def res_ln_bwd(a: ResLnArgs):
    check(_lib.lib().pmgt_res_ln_bwd(C.byref(a), cur_stream()), "pmgt_res_ln_bwd")


def colsum(x, out_f32):
    T, N = x.shape
    check(_lib.lib().pmgt_colsum_bf16(ptr(x), T, N, x.stride(0), ptr(out_f32), cur_stream()), "pmgt_colsum_bf16")


def gsr(fwd: bool, B, SP, H, tgt_h, ld_t, pair_h, ld_p, pair_off, labels, logits=None, loss_out=None, grad_out=None,
        d_tgt=None, d_pair=None):
    a = GsrArgs(B, SP, H, ptr(tgt_h), ld_t, ptr(pair_h), ld_p, ptr(pair_off), ptr(labels), ptr(logits), ptr(loss_out),
                ptr(grad_out), ptr(d_tgt), ptr(d_pair))
    fn = _lib.lib().pmgt_gsr_fwd if fwd else _lib.lib().pmgt_gsr_bwd
    check(fn(C.byref(a), cur_stream()), "pmgt_gsr")


def nfr_mse(fwd: bool, Mm, D, proj, table, target_ids, weight, loss_out=None, grad_out=None, dproj=None):
    a = NfrArgs(Mm, D, ptr(proj), proj.stride(0) if proj is not None else 0, ptr(table), table.stride(0),
                ptr(target_ids), weight, ptr(loss_out), ptr(grad_out), ptr(dproj))
    fn = _lib.lib().pmgt_nfr_mse_fwd if fwd else _lib.lib().pmgt_nfr_mse_bwd
    check(fn(C.byref(a), cur_stream()), "pmgt_nfr_mse")


def cast_f32_bf16(src, dst):
    check(_lib.lib().pmgt_cast_f32_bf16(ptr(src), ptr(dst), src.numel(), cur_stream()), "pmgt_cast_f32_bf16")


def sumsq(x, out):
    check(_lib.lib().pmgt_sumsq_f32(ptr(x), x.numel(), ptr(out), cur_stream()), "pmgt_sumsq_f32")


def gather_rows(src, idx, out):
    check(_lib.lib().pmgt_gather_rows_bf16(ptr(src), src.stride(0), ptr(idx), idx.numel(), src.shape[1], ptr(out),
                                           out.stride(0), cur_stream()), "pmgt_gather_rows_bf16")


def adamw_step(p, g, m, v, decay_mask, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, grad_scale_dev=None):
    check(_lib.lib().pmgt_adamw_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(decay_mask), p.numel(), lr, beta1, beta2, eps,
                                     weight_decay, step, grad_scale, ptr(grad_scale_dev), cur_stream()), "pmgt_adamw_step")


def sample_contexts(graph_handle, roots, keys, hops, max_ctx, seed, out_ids, out_mask, out_visited_deg=None):
    h = (C.c_int32 * len(hops))(*hops)
    check(_lib.lib().pmgt_sample_contexts(graph_handle, ptr(roots), ptr(keys), roots.numel(), h, len(hops), max_ctx,
                                          seed, ptr(out_ids), ptr(out_mask), ptr(out_visited_deg), cur_stream()),
          "pmgt_sample_contexts")


def sample_pairs(graph_handle, targets, keys, max_pos, min_neg, max_total, stride, seed, out_pairs, out_labels, out_num):
    check(_lib.lib().pmgt_sample_pairs(graph_handle, ptr(targets), ptr(keys), targets.numel(), max_pos, min_neg,
                                       max_total, stride, seed, ptr(out_pairs), ptr(out_labels), ptr(out_num),
                                       cur_stream()), "pmgt_sample_pairs")
