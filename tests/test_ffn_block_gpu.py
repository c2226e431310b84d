"""Parity of the fused post-attention block kernels (pmgt_block_fwd / pmgt_block_bwd, csrc/ffn_block.cu) against fp32
torch autograd of the reference compositions -- BertSelfOutput (modeling_pmgt.py:358-375: dense + dropout + residual +
LayerNorm) and BertIntermediate + BertOutput (modeling_pmgt.py:296-325) -- on identical bf16-rounded inputs, and, with
dropout on, against the unfused token-tile chain, which draws the same dropout stream.  Sizes: ragged last tile, fewer
tiles than SMs, more tiles than SMs (persistent wrap-around), and the BASELINE config-2 token count (294,912 tokens:
2,304 tiles, ~16 per CTA)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
SIZES = [200, 128 * 5, 40000 + 37, 294912]
EPS = 1e-12


def _ops():
    from pmgt_b200 import ops
    return ops


def _r(*shape, s=1.0):
    return (torch.randn(*shape, device="cuda") * s).to(BF16)


def _close(got, want, tol=1e-2, name=""):
    scale = want.abs().max().clamp_min(1e-6)
    assert torch.isfinite(got.float()).all(), f"{name}: non-finite"
    err = float((got.float() - want).abs().max() / scale)
    assert err < tol, f"{name}: max scaled error {err:.4g} >= {tol}"


def _cos(got, want, lo=0.999, name=""):
    c = float(F.cosine_similarity(got.float().flatten(), want.float().flatten(), dim=0))
    assert c > lo, f"{name}: cosine {c:.6f} <= {lo}"


def _params():
    torch.manual_seed(5)
    w1, w2 = _r(128, 128, s=0.1), _r(128, 128, s=0.1)
    b1 = torch.randn(128, device="cuda") * 0.3
    b2 = torch.randn(128, device="cuda") * 0.3
    g = 1 + 0.1 * torch.randn(128, device="cuda")
    be = 0.1 * torch.randn(128, device="cuda")
    return w1, b1, w2, b2, g, be


def _ref(x, res, w1, b1, w2, b2, g, be, ffn):
    if ffn:
        x = F.gelu(F.linear(x, w1.float(), b1)).to(BF16).float()      # the kernel feeds bf16 h to the second product
    return F.layer_norm(F.linear(x, w2.float(), b2) + res, (128,), g, be, EPS)


def _args(ops, ffn, x, res, prm, p=0.0, seed=0, site=0):
    w1, b1, w2, b2, g, be = prm
    if ffn:
        return ops.block_args(x, w2, b2, g, be, EPS, p, seed, site, w1=w1, b1=b1)
    return ops.block_args(x, w2, b2, g, be, EPS, p, seed, site, res=res)


@pytest.mark.parametrize("T", SIZES)
@pytest.mark.parametrize("ffn", [True, False])
@pytest.mark.parametrize("f32", [False, True])
def test_block_fwd(T, ffn, f32):
    ops = _ops()
    prm = _params()
    x = _r(T, 128)
    res = x if ffn else _r(T, 128)
    out = torch.full((T, 128), 7.0, device="cuda", dtype=BF16)
    o32 = torch.empty(T, 128, device="cuda") if f32 else None
    ops.block_fwd(_args(ops, ffn, x, res, prm), out, o32)
    want = _ref(x.float(), res.float(), *prm, ffn)
    _close(out, want, 1.5e-2, name="out")
    if f32:
        _close(o32, want, 1.5e-2, name="out_f32")
        assert torch.equal(o32.to(BF16), out)


@pytest.mark.parametrize("T", SIZES)
@pytest.mark.parametrize("ffn", [True, False])
@pytest.mark.parametrize("two_terms", [False, True])
def test_block_bwd(T, ffn, two_terms):
    ops = _ops()
    prm = _params()
    w1, b1, w2, b2, g, be = prm
    x = _r(T, 128)
    res = x if ffn else _r(T, 128)
    dy = _r(T, 128, s=0.5)
    dy_b = _r(T, 128, s=0.5) if two_terms else None
    dx = torch.full((T, 128), 3.0, device="cuda", dtype=BF16)
    dz = None if ffn else torch.full((T, 128), 3.0, device="cuda", dtype=BF16)
    dw1 = torch.full((128, 128), 0.25, device="cuda")   # accumulated on top of what is there
    dw2 = torch.full((128, 128), -0.5, device="cuda")
    db1, db2 = torch.ones(128, device="cuda"), torch.ones(128, device="cuda")
    dg, dbe = torch.ones(128, device="cuda"), torch.ones(128, device="cuda")
    fa = _args(ops, ffn, x, res, prm)
    out = torch.empty(T, 128, device="cuda", dtype=BF16)
    sv = ops.BlockSaved(T, ffn, 0.0, "cuda")
    out2 = torch.empty_like(out)
    ops.block_fwd(_args(ops, ffn, x, res, prm), out2)
    ops.block_fwd(fa, out, None, sv)
    assert torch.equal(out, out2), "saving activations must not change the output"
    if ffn:
        ops.block_bwd(fa, sv, dy, dx, dw2, db2, dg, dbe, dy_b=dy_b, dw1=dw1, db1=db1)
    else:
        ops.block_bwd(fa, sv, dy, dx, dw2, db2, dg, dbe, dy_b=dy_b, dz=dz)
    x32 = x.float().requires_grad_(True)
    r32 = x32 if ffn else res.float().requires_grad_(True)
    ps = [t.float().clone().requires_grad_(True) for t in (w1, b1, w2, b2, g, be)]
    hcur = F.gelu(F.linear(x32, ps[0], ps[1])) if ffn else x32
    y = F.layer_norm(F.linear(hcur, ps[2], ps[3]) + r32, (128,), ps[4], ps[5], EPS)
    y.backward(dy.float() + (dy_b.float() if two_terms else 0))
    _close(dx, x32.grad, 2e-2, name="dx")
    _cos(dx, x32.grad, name="dx")
    if not ffn:
        _close(dz, r32.grad, 2e-2, name="dz")
        _cos(dz, r32.grad, name="dz")
    checks = [(dw2, ps[2], -0.5, "dw2"), (db2, ps[3], 1.0, "db2"), (dg, ps[4], 1.0, "d_gamma"), (dbe, ps[5], 1.0, "d_beta")]
    if ffn:
        checks += [(dw1, ps[0], 0.25, "dw1"), (db1, ps[1], 1.0, "db1")]
    for got, ref, base, nm in checks:
        _close(got - base, ref.grad, 1e-2, name=nm)
        _cos(got - base, ref.grad, name=nm)


@pytest.mark.parametrize("T", [640, 40037])
def test_ffn_dropout_matches_the_unfused_chain(T):
    """Same seed / site => the fused kernels draw the dropout mask of the token-tile chain (LT_GELU + LT_RES_LN forward,
    pmgt_ln_bwd + fused dX/dW backward): outputs agree to bf16 rounding, and about p of the dense outputs are dropped."""
    ops = _ops()
    prm = _params()
    w1, b1, w2, b2, g, be = prm
    a, dy = _r(T, 128), _r(T, 128, s=0.5)
    p, seed, site = 0.1, 0x1234567890ABCDEF, 14
    out = torch.empty(T, 128, device="cuda", dtype=BF16)
    fa = _args(ops, True, a, a, prm, p, seed, site)
    sv = ops.BlockSaved(T, True, p, "cuda")
    ops.block_fwd(fa, out, None, sv)
    # unfused chain
    h_pre, h, z, y = (torch.empty(T, 128, device="cuda", dtype=BF16) for _ in range(4))
    ops.linear_tile(a, w1, h, ops.LT_GELU, bias=b1, aux_out=h_pre)
    ops.linear_tile(h, w2, y, ops.LT_RES_LN, bias=b2, aux_out=z, e_in=a, ln_g=g, ln_b=be, ln_eps=EPS, p=p, seed=seed, site=site)
    _close(out, y.float(), 1e-2, name="out vs unfused")
    _close(sv.h, h.float(), 1e-2, name="h vs unfused")
    p32 = h_pre.float().requires_grad_(True)
    F.gelu(p32).sum().backward()
    _close(sv.gp, p32.grad, 1e-2, name="gelu' vs autograd")
    xhat = F.layer_norm(z.float(), (128,), None, None, EPS)
    _close(sv.xhat, xhat, 1e-2, name="xhat vs unfused")
    # without dropout the result differs: the mask is really applied
    out0 = torch.empty_like(out)
    ops.block_fwd(_args(ops, True, a, a, prm, 0.0, seed, site), out0)
    assert float((out0.float() - out.float()).abs().max()) > 0.05
    # backward
    da = torch.empty(T, 128, device="cuda", dtype=BF16)
    grads = [torch.zeros(128, 128, device="cuda"), torch.zeros(128, 128, device="cuda")] + [torch.zeros(128, device="cuda") for _ in range(4)]
    ops.block_bwd(fa, sv, dy, da, grads[1], grads[3], grads[4], grads[5], dw1=grads[0], db1=grads[2])
    dz, do = torch.empty(T, 128, device="cuda", dtype=BF16), torch.empty(T, 128, device="cuda", dtype=BF16)
    rg = [torch.zeros(128, 128, device="cuda"), torch.zeros(128, 128, device="cuda")] + [torch.zeros(128, device="cuda") for _ in range(4)]
    ops.ln_bwd(T, 128, z, g, EPS, p, seed, site, dz, do, rg[4], rg[5], dy_a=dy)
    dh_pre, da_ffn = torch.empty(T, 128, device="cuda", dtype=BF16), torch.empty(T, 128, device="cuda", dtype=BF16)
    ops.linear_tile(do, w2, dh_pre, ops.LT_GELU_BWD, w_mn=True, e_in=h_pre, dw_x=h, dw=rg[1], dbias=rg[3])
    ops.linear_tile(dh_pre, w1, da_ffn, ops.LT_PLAIN, w_mn=True, dw_x=a, dw=rg[0], dbias=rg[2])
    want_da = da_ffn.float() + dz.float()
    _close(da, want_da, 2e-2, name="da vs unfused")
    for got, ref, nm in zip(grads, rg, ("dw1", "dw2", "db1", "db2", "d_gamma", "d_beta")):
        _close(got, ref, 1e-2, name=nm + " vs unfused")
    kept = float((do.float() != 0).float().mean())
    assert abs(kept - (1 - p)) < 0.01


@pytest.mark.parametrize("T", [640, 40037])
def test_dense_block_dropout_matches_the_unfused_chain(T):
    ops = _ops()
    prm = _params()
    w1, b1, w2, b2, g, be = prm
    ctx, x, dy, dyb = _r(T, 128), _r(T, 128), _r(T, 128, s=0.5), _r(T, 128, s=0.5)
    p, seed, site = 0.1, 0xFEDCBA9876543210, 13
    out = torch.empty(T, 128, device="cuda", dtype=BF16)
    fa = _args(ops, False, ctx, x, prm, p, seed, site)
    sv = ops.BlockSaved(T, False, p, "cuda")
    ops.block_fwd(fa, out, None, sv)
    z, y = torch.empty_like(out), torch.empty_like(out)
    ops.linear_tile(ctx, w2, y, ops.LT_RES_LN, bias=b2, aux_out=z, e_in=x, ln_g=g, ln_b=be, ln_eps=EPS, p=p, seed=seed, site=site)
    _close(out, y.float(), 1e-2, name="out vs unfused")
    dctx, dz = torch.empty_like(out), torch.empty_like(out)
    grads = [torch.zeros(128, 128, device="cuda")] + [torch.zeros(128, device="cuda") for _ in range(3)]
    ops.block_bwd(fa, sv, dy, dctx, grads[0], grads[1], grads[2], grads[3], dy_b=dyb, dz=dz)
    rdz, rdo, rdctx = torch.empty_like(out), torch.empty_like(out), torch.empty_like(out)
    rg = [torch.zeros(128, 128, device="cuda")] + [torch.zeros(128, device="cuda") for _ in range(3)]
    ops.ln_bwd(T, 128, z, g, EPS, p, seed, site, rdz, rdo, rg[2], rg[3], dy_a=dy, dy_b=dyb)
    ops.linear_tile(rdo, w2, rdctx, ops.LT_PLAIN, w_mn=True, dw_x=ctx, dw=rg[0], dbias=rg[1])
    _close(dz, rdz.float(), 2e-2, name="dz vs unfused")
    _close(dctx, rdctx.float(), 2e-2, name="dctx vs unfused")
    for got, ref, nm in zip(grads, rg, ("dw", "db", "d_gamma", "d_beta")):
        _close(got, ref, 1e-2, name=nm + " vs unfused")
