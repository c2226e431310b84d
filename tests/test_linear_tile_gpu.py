"""Parity of the persistent tcgen05 token-tile kernels (pmgt_linear_tile / pmgt_dw_tile / pmgt_ln_bwd) against fp32
torch on identical bf16-rounded inputs.  Sizes cover a ragged last tile, fewer tiles than SMs and more tiles than SMs
(persistent loop wrap-around, ring / TMEM-slot phase flips)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
SIZES = [200, 128 * 5, 40000 + 37]


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


@pytest.mark.parametrize("T", SIZES)
@pytest.mark.parametrize("N", [128, 512])
def test_bias(T, N):
    ops = _ops()
    x, w = _r(T, 128), _r(N, 128, s=0.1)
    b = torch.randn(N, device="cuda")
    out = torch.full((T, N), 7.0, device="cuda", dtype=BF16)
    ops.linear_tile(x, w, out, ops.LT_BIAS, bias=b)
    _close(out, F.linear(x.float(), w.float(), b), name="bias")


@pytest.mark.parametrize("T", SIZES)
def test_gelu(T):
    ops = _ops()
    x, w = _r(T, 128), _r(128, 128, s=0.1)
    b = torch.randn(128, device="cuda") * 0.3
    out = torch.empty(T, 128, device="cuda", dtype=BF16)
    pre = torch.empty_like(out)
    ops.linear_tile(x, w, out, ops.LT_GELU, bias=b, aux_out=pre)
    want_pre = F.linear(x.float(), w.float(), b)
    _close(pre, want_pre, name="pre")
    _close(out, F.gelu(want_pre), name="gelu")


@pytest.mark.parametrize("T", SIZES)
@pytest.mark.parametrize("f32", [False, True])
def test_res_ln(T, f32):
    ops = _ops()
    x, w, res = _r(T, 128), _r(128, 128, s=0.1), _r(T, 128)
    b = torch.randn(128, device="cuda") * 0.3
    g = 1 + 0.1 * torch.randn(128, device="cuda")
    be = 0.1 * torch.randn(128, device="cuda")
    out = torch.empty(T, 128, device="cuda", dtype=BF16)
    z = torch.empty_like(out)
    o32 = torch.empty(T, 128, device="cuda") if f32 else None
    ops.linear_tile(x, w, out, ops.LT_RES_LN, bias=b, aux_out=z, e_in=res, ln_g=g, ln_b=be, ln_eps=1e-12, out_f32=o32)
    want_z = F.linear(x.float(), w.float(), b) + res.float()
    _close(z, want_z, name="z")
    want = F.layer_norm(want_z, (128,), g, be, 1e-12)
    _close(out, want, 1.5e-2, name="y")
    if f32:
        _close(o32, want, 1.5e-2, name="y_f32")
        assert torch.equal(o32.to(BF16), out)


@pytest.mark.parametrize("T", SIZES)
@pytest.mark.parametrize("K", [128, 512])
def test_plain_dx(T, K):
    ops = _ops()
    dy, w = _r(T, K), _r(K, 128, s=0.1)  # w is the nn.Linear weight [N_w = K][K_w = 128]
    out = torch.empty(T, 128, device="cuda", dtype=BF16)
    ops.linear_tile(dy, w, out, ops.LT_PLAIN, w_mn=True)
    _close(out, dy.float() @ w.float(), name="dx")


@pytest.mark.parametrize("T", SIZES)
def test_gelu_bwd(T):
    ops = _ops()
    dy, w, pre = _r(T, 128), _r(128, 128, s=0.1), _r(T, 128)
    out = torch.empty(T, 128, device="cuda", dtype=BF16)
    ops.linear_tile(dy, w, out, ops.LT_GELU_BWD, w_mn=True, e_in=pre)
    p32 = pre.float().requires_grad_(True)
    F.gelu(p32).backward(dy.float() @ w.float())
    _close(out, p32.grad, name="dh_pre")


@pytest.mark.parametrize("T", SIZES)
@pytest.mark.parametrize("N", [128, 512])
def test_dw(T, N):
    ops = _ops()
    dy, x = _r(T, N), _r(T, 128)
    dw = torch.ones(N, 128, device="cuda")
    db = torch.ones(N, device="cuda")
    ops.dw_tile(dy, x, dw, db)
    _close(dw, dy.float().t() @ x.float() + 1, 2e-3, name="dw")
    _close(db, dy.float().sum(0) + 1, 2e-3, name="db")
    dw2 = torch.zeros(N, 128, device="cuda")
    ops.dw_tile(dy, x, dw2, None)
    _close(dw2, dy.float().t() @ x.float(), 2e-3, name="dw (no bias)")


@pytest.mark.parametrize("T", SIZES)
@pytest.mark.parametrize("gelu", [False, True])
def test_dx_with_fused_dw(T, gelu):
    """dX call that also forms dW += dY^T X and dbias += colsum(dY) (pmgt_linear_tile with dw != NULL)."""
    ops = _ops()
    dy, x_in = _r(T, 128, s=0.5), _r(T, 128, s=0.5)
    w = _r(128, 128, s=0.1)                      # nn.Linear weight [N = out][K = in]; dX = dY W
    pre = _r(T, 128) if gelu else None
    out = torch.empty(T, 128, device="cuda", dtype=BF16)
    dw = torch.full((128, 128), 0.25, device="cuda")  # accumulates on top of what is there
    db = torch.full((128,), -1.0, device="cuda")
    ops.linear_tile(dy, w, out, ops.LT_GELU_BWD if gelu else ops.LT_PLAIN, w_mn=True, e_in=pre, dw_x=x_in, dw=dw, dbias=db)
    want = dy.float() @ w.float()
    if gelu:
        p = pre.float().requires_grad_(True)
        F.gelu(p).sum().backward()
        want = want * p.grad
    _close(out, want, name="dx")
    _close(dw - 0.25, dy.float().t() @ x_in.float(), 2e-3, "dw")
    _close(db + 1.0, dy.float().sum(0), 2e-3, "dbias")
    # without dbias
    dw2 = torch.zeros(128, 128, device="cuda")
    out2 = torch.empty_like(out)
    ops.linear_tile(dy, w, out2, ops.LT_GELU_BWD if gelu else ops.LT_PLAIN, w_mn=True, e_in=pre, dw_x=x_in, dw=dw2)
    assert torch.equal(out, out2)
    _close(dw2, dy.float().t() @ x_in.float(), 2e-3, "dw (no dbias)")


@pytest.mark.parametrize("N", [128, 512])
@pytest.mark.parametrize("n", [1, 3, 5, 9])
def test_dw_batch(N, n):
    """pmgt_dw_tile_batch: n independent problems (different token counts, one of them a single ragged tile) in one launch."""
    ops = _ops()
    sizes = [40000 + 37, 77, 128 * 9, 5000, 200, 300, 1000, 129, 64][:n]   # n = 9 > 8: two launches
    probs, want = [], []
    for i, T in enumerate(sizes):
        dy, x = _r(T, N, s=0.5), _r(T, 128, s=0.5)
        dw = torch.full((N, 128), float(i), device="cuda")
        db = torch.zeros(N, device="cuda") if i % 2 == 0 else None
        probs.append((dy, x, dw, db))
        want.append((dy.float().t() @ x.float() + float(i), dy.float().sum(0)))
    ops.dw_tile_batch(probs)
    for (dy, x, dw, db), (wdw, wdb) in zip(probs, want):
        _close(dw, wdw, 2e-3, "dw")
        if db is not None:
            _close(db, wdb, 2e-3, "dbias")


@pytest.mark.parametrize("T", [77, 5000])
def test_ln_bwd(T):
    ops = _ops()
    z, dya, dyb = _r(T, 128), _r(T, 128), _r(T, 128)
    dy32 = torch.randn(T, 128, device="cuda")
    g = 1 + 0.1 * torch.randn(128, device="cuda")
    zf = z.float().requires_grad_(True)
    gg = g.clone().requires_grad_(True)
    bb = torch.zeros(128, device="cuda", requires_grad=True)
    F.layer_norm(zf, (128,), gg, bb, 1e-12).backward(dya.float() + dyb.float() + dy32)
    dz = torch.empty_like(z)
    dg, db = torch.zeros(128, device="cuda"), torch.zeros(128, device="cuda")
    ops.ln_bwd(T, 128, z, g, 1e-12, 0.0, 0, 0, dz, dz, dg, db, dy_a=dya, dy_b=dyb, dy_f32=dy32)
    _close(dz, zf.grad, name="dz")
    _close(dg, gg.grad, 5e-3, "d_g")
    _close(db, bb.grad, 5e-3, "d_b")


@pytest.mark.parametrize("T", [1, 63, 64, 77, 5000, 20011])
@pytest.mark.parametrize("combo", ["a", "ab", "b", "f32"])
def test_ln_bwd_streamed_variants(T, combo):
    """The bulk-copy-staged kernel covers (dy_a), (dy_a + dy_b), (dy_b) and (dy_f32) -- the combinations the encoder
    issues; ragged tails (T % 64 != 0) must not read the stale part of a stage."""
    ops = _ops()
    z = _r(T, 128)
    dya = _r(T, 128) if "a" in combo else None
    dyb = _r(T, 128) if "b" in combo else None
    dy32 = torch.randn(T, 128, device="cuda") if combo == "f32" else None
    g = 1 + 0.1 * torch.randn(128, device="cuda")
    zf = z.float().requires_grad_(True)
    gg = g.clone().requires_grad_(True)
    bb = torch.zeros(128, device="cuda", requires_grad=True)
    total = sum(t.float() for t in (dya, dyb, dy32) if t is not None)
    F.layer_norm(zf, (128,), gg, bb, 1e-12).backward(total)
    dz = torch.empty_like(z)
    dg, db = torch.zeros(128, device="cuda"), torch.zeros(128, device="cuda")
    ops.ln_bwd(T, 128, z, g, 1e-12, 0.0, 0, 0, dz, dz, dg, db, dy_a=dya, dy_b=dyb, dy_f32=dy32)
    _close(dz, zf.grad, name="dz")
    _close(dg, gg.grad, 5e-3, "d_g")
    _close(db, bb.grad, 5e-3, "d_b")


def test_res_ln_dropout_matches_ln_bwd_mask():
    """The RES_LN epilogue and pmgt_ln_bwd must regenerate the same Philox mask: with x = 0 and bias = 1 the
    dense output is 1 everywhere, so z - res is 0 exactly where the forward dropped."""
    ops = _ops()
    T, p = 1000, 0.25
    x, w = torch.zeros(T, 128, device="cuda", dtype=BF16), _r(128, 128)
    b = torch.ones(128, device="cuda")
    res = torch.zeros(T, 128, device="cuda", dtype=BF16)
    g, be = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
    out, z = torch.empty(T, 128, device="cuda", dtype=BF16), torch.empty(T, 128, device="cuda", dtype=BF16)
    ops.linear_tile(x, w, out, ops.LT_RES_LN, bias=b, aux_out=z, e_in=res, ln_g=g, ln_b=be, ln_eps=1e-12, p=p, seed=99, site=3)
    dropped = z.float() == 0
    rate = float(dropped.float().mean())
    assert abs(rate - p) < 0.01, rate
    assert torch.allclose(z.float()[~dropped], torch.tensor(1 / (1 - p), device="cuda"), rtol=1e-2)
    dy = _r(T, 128)
    dz, d_o = torch.empty_like(z), torch.empty_like(z)
    ops.ln_bwd(T, 128, z, g, 1e-12, p, 99, 3, dz, d_o, torch.zeros(128, device="cuda"), torch.zeros(128, device="cuda"),
               dy_a=dy)
    assert bool((d_o.float()[dropped] == 0).all())
    assert torch.allclose(d_o.float()[~dropped], (dz.float() / (1 - p))[~dropped], rtol=2e-2, atol=1e-3)


def test_launch_options_do_not_change_results():
    """pmgt_set_pdl / pmgt_set_alternate_order only change HOW the chain is launched and walked: outputs are identical."""
    ops = _ops()
    T = 40000 + 37
    x, w = _r(T, 128, s=0.5), _r(128, 128, s=0.1)
    b = torch.randn(128, device="cuda")
    outs = []
    prev_pdl, prev_alt = ops.set_pdl(True), ops.set_alternate_order(True)
    try:
        for pdl, alt in ((True, True), (False, True), (True, False), (False, False)):
            ops.set_pdl(pdl)
            ops.set_alternate_order(alt)
            for _ in range(3):  # odd and even positions of the alternation
                out, pre = torch.empty(T, 128, device="cuda", dtype=BF16), torch.empty(T, 128, device="cuda", dtype=BF16)
                ops.linear_tile(x, w, out, ops.LT_GELU, bias=b, aux_out=pre)
                outs.append((out, pre))
        assert ops.set_pdl(True) is False and ops.set_alternate_order(True) is False   # previous settings are returned
    finally:
        ops.set_pdl(prev_pdl)
        ops.set_alternate_order(prev_alt)
    for out, pre in outs[1:]:
        assert torch.equal(out, outs[0][0]) and torch.equal(pre, outs[0][1])
