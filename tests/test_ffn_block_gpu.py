"""Parity of the fused feed-forward block kernels (pmgt_ffn_fwd / pmgt_ffn_bwd, csrc/ffn_block.cu) against fp32 torch
autograd of the reference composition (BertIntermediate + BertOutput, modeling_pmgt.py:296-325) on identical
bf16-rounded inputs, and -- with dropout on -- against the unfused token-tile chain, which draws the same dropout
stream.  Sizes: ragged last tile, fewer tiles than SMs, more tiles than SMs (persistent wrap-around), and the
BASELINE config-2 token count (294,912 tokens: 2,304 tiles, ~16 per CTA)."""
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


def _ref(a, w1, b1, w2, b2, g, be):
    h = F.gelu(F.linear(a, w1.float(), b1)).to(BF16).float()      # the kernel feeds bf16 h to the second product
    return F.layer_norm(F.linear(h, w2.float(), b2) + a, (128,), g, be, EPS)


@pytest.mark.parametrize("T", SIZES)
@pytest.mark.parametrize("f32", [False, True])
def test_ffn_fwd(T, f32):
    ops = _ops()
    w1, b1, w2, b2, g, be = _params()
    a = _r(T, 128)
    out = torch.full((T, 128), 7.0, device="cuda", dtype=BF16)
    o32 = torch.empty(T, 128, device="cuda") if f32 else None
    ops.ffn_fwd(ops.ffn_args(a, w1, b1, w2, b2, g, be, EPS, 0.0, 0, 0), out, o32)
    want = _ref(a.float(), w1, b1, w2, b2, g, be)
    _close(out, want, 1.5e-2, name="out")
    if f32:
        _close(o32, want, 1.5e-2, name="out_f32")
        assert torch.equal(o32.to(BF16), out)


@pytest.mark.parametrize("T", SIZES)
@pytest.mark.parametrize("two_terms", [False, True])
def test_ffn_bwd(T, two_terms):
    ops = _ops()
    w1, b1, w2, b2, g, be = _params()
    a = _r(T, 128)
    dy = _r(T, 128, s=0.5)
    dy_b = _r(T, 128, s=0.5) if two_terms else None
    da = torch.full((T, 128), 3.0, device="cuda", dtype=BF16)
    dw1 = torch.full((128, 128), 0.25, device="cuda")   # accumulated on top of what is there
    dw2 = torch.full((128, 128), -0.5, device="cuda")
    db1, db2 = torch.ones(128, device="cuda"), torch.ones(128, device="cuda")
    dg, dbe = torch.ones(128, device="cuda"), torch.ones(128, device="cuda")
    fa = ops.ffn_args(a, w1, b1, w2, b2, g, be, EPS, 0.0, 0, 0)
    out, h, gp = (torch.empty(T, 128, device="cuda", dtype=BF16) for _ in range(3))
    ops.ffn_fwd(fa, out, None, h, gp)
    ops.ffn_bwd(fa, h, gp, dy, da, dw1, dw2, db1, db2, dg, dbe, dy_b=dy_b)
    a32 = a.float().requires_grad_(True)
    ps = [t.float().clone().requires_grad_(True) for t in (w1, b1, w2, b2, g, be)]
    h = F.gelu(F.linear(a32, ps[0], ps[1]))
    y = F.layer_norm(F.linear(h, ps[2], ps[3]) + a32, (128,), ps[4], ps[5], EPS)
    y.backward(dy.float() + (dy_b.float() if two_terms else 0))
    _close(da, a32.grad, 2e-2, name="da")
    _cos(da, a32.grad, name="da")
    for got, ref, base, nm in ((dw1, ps[0], 0.25, "dw1"), (db1, ps[1], 1.0, "db1"), (dw2, ps[2], -0.5, "dw2"),
                               (db2, ps[3], 1.0, "db2"), (dg, ps[4], 1.0, "d_gamma"), (dbe, ps[5], 1.0, "d_beta")):
        _close(got - base, ref.grad, 1e-2, name=nm)
        _cos(got - base, ref.grad, name=nm)


@pytest.mark.parametrize("T", [640, 40037])
def test_ffn_dropout_matches_the_unfused_chain(T):
    """Same seed / site => the fused kernels draw the dropout mask of the token-tile chain (LT_GELU + LT_RES_LN forward,
    pmgt_ln_bwd + fused dX/dW backward): outputs agree to bf16 rounding, and about p of the dense outputs are dropped."""
    ops = _ops()
    w1, b1, w2, b2, g, be = _params()
    a, dy = _r(T, 128), _r(T, 128, s=0.5)
    p, seed, site = 0.1, 0x1234567890ABCDEF, 14
    out = torch.empty(T, 128, device="cuda", dtype=BF16)
    fa = ops.ffn_args(a, w1, b1, w2, b2, g, be, EPS, p, seed, site)
    hs, gps = torch.empty_like(out), torch.empty_like(out)
    ops.ffn_fwd(fa, out, None, hs, gps)
    # unfused chain
    h_pre, h, z, y = (torch.empty(T, 128, device="cuda", dtype=BF16) for _ in range(4))
    ops.linear_tile(a, w1, h, ops.LT_GELU, bias=b1, aux_out=h_pre)
    ops.linear_tile(h, w2, y, ops.LT_RES_LN, bias=b2, aux_out=z, e_in=a, ln_g=g, ln_b=be, ln_eps=EPS, p=p, seed=seed, site=site)
    _close(out, y.float(), 1e-2, name="out vs unfused")
    _close(hs, h.float(), 1e-2, name="h vs unfused")
    p32 = h_pre.float().requires_grad_(True)
    F.gelu(p32).sum().backward()
    _close(gps, p32.grad, 1e-2, name="gelu' vs autograd")
    # without dropout the result differs: the mask is really applied
    out0 = torch.empty_like(out)
    ops.ffn_fwd(ops.ffn_args(a, w1, b1, w2, b2, g, be, EPS, 0.0, seed, site), out0)
    assert float((out0.float() - out.float()).abs().max()) > 0.05
    # backward
    da = torch.empty(T, 128, device="cuda", dtype=BF16)
    grads = [torch.zeros(128, 128, device="cuda"), torch.zeros(128, 128, device="cuda")] + [torch.zeros(128, device="cuda") for _ in range(4)]
    ops.ffn_bwd(fa, hs, gps, dy, da, *grads)
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
