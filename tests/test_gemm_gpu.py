"""tcgen05 GEMM (pmgt_gemm_bf16) against torch fp32 matmul on the same bf16 inputs.
Tolerance: fp32 accumulation of bf16 products, result rounded to bf16 -> 1e-2
relative to the row scale; fp32 outputs 2e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16


def _ops():
    from pmgt_b200 import ops
    return ops


def _close(got, want, tol):
    scale = want.abs().max().clamp_min(1e-6)
    err = (got.float() - want).abs().max() / scale
    assert torch.isfinite(got.float()).all(), "non-finite output"
    assert err < tol, f"max scaled error {float(err):.4g} >= {tol}"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 128), (256, 128, 256), (300, 128, 128), (77, 32, 40),
                                   (1000, 512, 128), (512, 128, 1536), (129, 264, 200)])
def test_linear_fwd(M, N, K):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    x = torch.randn(M, K, device="cuda", generator=g).to(BF16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.1).to(BF16)
    b = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=BF16)
    ops.linear_fwd(x, w, b, out)
    _close(out, x.float() @ w.float().t() + b, 1e-2)


def test_linear_fwd_gelu_epilogue():
    ops = _ops()
    M, N, K = 333, 128, 128
    x = torch.randn(M, K, device="cuda").to(BF16)
    w = (torch.randn(N, K, device="cuda") * 0.1).to(BF16)
    b = torch.randn(N, device="cuda") * 0.1
    out = torch.empty(M, N, device="cuda", dtype=BF16)
    pre = torch.empty(M, N, device="cuda", dtype=BF16)
    ops.linear_fwd(x, w, b, out, gelu_aux=pre)
    want_pre = x.float() @ w.float().t() + b
    _close(pre, want_pre, 1e-2)
    _close(out, torch.nn.functional.gelu(want_pre), 1e-2)


@pytest.mark.parametrize("M,D,K,nrows", [(256, 128, 1536, 500), (700, 128, 768, 50), (100, 64, 128, 30)])
def test_linear_fwd_gathered_rows(M, D, K, nrows):
    """A rows fetched through an index vector (the fused feature gather)."""
    ops = _ops()
    table = torch.randn(nrows, K, device="cuda").to(BF16)
    table[0] = 0
    idx = torch.randint(0, nrows, (M,), device="cuda", dtype=torch.int64)
    w = (torch.randn(D, K, device="cuda") * 0.05).to(BF16)
    b = torch.randn(D, device="cuda") * 0.1
    out = torch.empty(M, D, device="cuda", dtype=BF16)
    ops.linear_fwd(table, w, b, out, rows=idx, src_rows=nrows)
    _close(out, table[idx].float() @ w.float().t() + b, 1e-2)


@pytest.mark.parametrize("M,N,K", [(256, 128, 128), (300, 128, 512), (1000, 128, 128), (130, 96, 64), (257, 1536, 128)])
def test_linear_dx(M, N, K):
    """dX[M,K] = dY[M,N] @ W[N,K]  (B operand MN-major)."""
    ops = _ops()
    dy = torch.randn(M, N, device="cuda").to(BF16)
    w = (torch.randn(N, K, device="cuda") * 0.1).to(BF16)
    add = torch.randn(M, K, device="cuda").to(BF16)
    out = torch.empty(M, K, device="cuda", dtype=BF16)
    ops.linear_dx(dy, w, out)
    _close(out, dy.float() @ w.float(), 1e-2)
    ops.linear_dx(dy, w, out, addend=add)
    _close(out, dy.float() @ w.float() + add.float(), 1e-2)


def test_linear_dx_gelu_bwd():
    ops = _ops()
    M, N, K = 200, 128, 96
    dy = torch.randn(M, N, device="cuda").to(BF16)
    w = (torch.randn(N, K, device="cuda") * 0.1).to(BF16)
    pre = torch.randn(M, K, device="cuda").to(BF16)
    out = torch.empty(M, K, device="cuda", dtype=BF16)
    ops.linear_dx(dy, w, out, gelu_bwd_aux=pre)
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).backward(dy.float() @ w.float())
    _close(out, x.grad, 1e-2)


@pytest.mark.parametrize("T,N,K", [(256, 128, 128), (1000, 128, 128), (5000, 512, 128), (333, 96, 64), (4096, 128, 1536)])
def test_linear_dw(T, N, K):
    """dW[N,K] += dY[T,N]^T @ X[T,K]  (both operands MN-major, split over T, fp32 atomics)."""
    ops = _ops()
    dy = (torch.randn(T, N, device="cuda") * 0.1).to(BF16)
    x = torch.randn(T, K, device="cuda").to(BF16)
    dw = torch.ones(N, K, device="cuda", dtype=torch.float32)  # accumulates on top
    ops.linear_dw(dy, x, dw)
    _close(dw, dy.float().t() @ x.float() + 1.0, 2e-3)


def test_linear_dw_gathered_rows():
    ops = _ops()
    T, N, K, nrows = 3000, 128, 768, 400
    table = torch.randn(nrows, K, device="cuda").to(BF16)
    idx = torch.randint(0, nrows, (T,), device="cuda", dtype=torch.int64)
    dy = (torch.randn(T, N, device="cuda") * 0.1).to(BF16)
    dw = torch.zeros(N, K, device="cuda", dtype=torch.float32)
    ops.linear_dw(dy, table, dw, rows=idx, src_rows=nrows, x_cols=K)
    _close(dw, dy.float().t() @ table[idx].float(), 2e-3)


@pytest.mark.parametrize("T,K,nrows", [(40000, 1536, 3000), (20001, 768, 777), (130, 256, 20), (5000, 512, 64), (33, 1536, 9)])
def test_gather_proj_kernels_match_torch(T, K, nrows):
    """The persistent gather-fused projection kernels (csrc/gather_proj.cu): forward wraps around the 148 CTAs at
    T = 40000 (313 tiles), ragged last tiles / stages, ids outside the table read as zero rows; dW over every slab
    count (K = 512 / 768 / 256 -> 8 / 6 / 4 slabs)."""
    ops = _ops()
    assert ops.gather_proj_supported(128, K)
    g = torch.Generator(device="cuda").manual_seed(T + K)
    table = torch.randn(nrows, K, device="cuda", generator=g).to(BF16)
    idx = torch.randint(0, nrows, (T,), device="cuda", dtype=torch.int64, generator=g)
    idx[::17] = nrows + 5  # out of range: zero row
    idx[3::29] = -1
    w = (torch.randn(128, K, device="cuda", generator=g) * 0.05).to(BF16)
    b = torch.randn(128, device="cuda", generator=g) * 0.1
    x = table[idx.clamp(0, nrows - 1)].float()
    x[(idx < 0) | (idx >= nrows)] = 0
    out = torch.full((T, 128), float("nan"), device="cuda", dtype=BF16)
    ops.linear_fwd(table, w, b, out, rows=idx, src_rows=nrows)
    _close(out, x @ w.float().t() + b, 1e-2)
    dy = (torch.randn(T, 128, device="cuda", generator=g) * 0.1).to(BF16)
    dw = torch.ones(128, K, device="cuda", dtype=torch.float32)
    ops.linear_dw(dy, table, dw, rows=idx, src_rows=nrows, x_cols=K)
    _close(dw, dy.float().t() @ x + 1.0, 3e-3)
    # the general GEMM path (pmgt_gemm_bf16 with a_rows) agrees
    ops.GATHER_PROJ = False
    try:
        out2 = torch.empty_like(out)
        ops.linear_fwd(table, w, b, out2, rows=idx, src_rows=nrows)
    finally:
        ops.GATHER_PROJ = True
    assert (out2.float() - out.float()).abs().max() < 2e-2

@pytest.mark.parametrize("T,K,nrows", [(40000, 1536, 3000), (20001, 768, 777), (4100, 256, 50)])
def test_gather_proj_tma_and_cp_async_paths_agree(T, K, nrows, monkeypatch):
    """csrc/gather_proj.cu fetches the table rows either with TMA tile::gather4 (one issuing lane in each of 8 warps,
    stages completed by the copies' own byte counts) or with per-thread 16-byte cp.async (PMGT_GATHER_TMA bit 0 forward,
    bit 1 weight gradient).  Same stage order and operands: the forward results are bit-identical, the weight gradients
    differ by the order of the fp32 reductions only.  Ids outside the table are zero rows on both paths."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(T * 3 + K)
    table = torch.randn(nrows, K, device="cuda", generator=g).to(BF16)
    idx = torch.randint(0, nrows, (T,), device="cuda", dtype=torch.int64, generator=g)
    idx[5::13] = nrows
    idx[2::31] = -7
    w = (torch.randn(128, K, device="cuda", generator=g) * 0.05).to(BF16)
    b = torch.randn(128, device="cuda", generator=g) * 0.1
    dy = (torch.randn(T, 128, device="cuda", generator=g) * 0.1).to(BF16)
    res = {}
    for mode in ("0", "3"):
        monkeypatch.setenv("PMGT_GATHER_TMA", mode)
        out = torch.full((T, 128), float("nan"), device="cuda", dtype=BF16)
        ops.linear_fwd(table, w, b, out, rows=idx, src_rows=nrows)
        dw = torch.zeros(128, K, device="cuda", dtype=torch.float32)
        ops.linear_dw(dy, table, dw, rows=idx, src_rows=nrows, x_cols=K)
        torch.cuda.synchronize()
        res[mode] = (out, dw)
    assert torch.equal(res["0"][0], res["3"][0])
    _close(res["3"][1], res["0"][1], 1e-4)
    x = table[idx.clamp(0, nrows - 1)].float()
    x[(idx < 0) | (idx >= nrows)] = 0
    _close(res["3"][0], x @ w.float().t() + b, 1e-2)
    _close(res["3"][1], dy.float().t() @ x, 3e-3)


@pytest.mark.parametrize("M,N,K", [(40037, 768, 256), (40000, 384, 320), (38000, 640, 256), (20001, 3072, 192)])
def test_persistent_gemm_kernels_match_torch(M, N, K):
    """Large dense GEMMs take the persistent kernel (umma_gemm_persist_kernel: more tiles than one wave): 128-column
    tiles for N < 512, 256-column tiles otherwise, ragged last row tile, a last column tile that is half out of range
    (N = 640), every epilogue of the Linear forward / dX paths."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = torch.randn(M, K, device="cuda", generator=g).to(BF16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(BF16)
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    out = torch.full((M, N), float("nan"), device="cuda", dtype=BF16)
    ops.linear_fwd(x, w, b, out)
    sl = torch.cat([torch.arange(0, 300, device="cuda"), torch.arange(M - 300, M, device="cuda"),
                    torch.randint(0, M, (400,), device="cuda", generator=g)])
    ref = x[sl].float() @ w.float().t() + b
    _close(out[sl], ref, 1e-2)
    assert torch.isfinite(out.float()).all()
    aux = torch.empty(M, N, device="cuda", dtype=BF16)
    ops.linear_fwd(x, w, b, out, gelu_aux=aux)
    _close(aux[sl], ref, 1e-2)
    _close(out[sl], torch.nn.functional.gelu(ref), 1e-2)
    # dX = dY W (+ addend) (* gelu'(aux)): B operand MN-major
    dy = torch.randn(M, N, device="cuda", generator=g).to(BF16)
    add = torch.randn(M, K, device="cuda", generator=g).to(BF16)
    dx = torch.full((M, K), float("nan"), device="cuda", dtype=BF16)
    if K >= 512 or M * ((K + 127) // 128) // 128 > 2 * 148:
        ops.linear_dx(dy, w, dx, addend=add)
        _close(dx[sl], dy[sl].float() @ w.float() + add[sl].float(), 1e-2)
        assert torch.isfinite(dx.float()).all()


def test_gemm_rejects_bad_arguments():
    from pmgt_b200 import _lib
    ops = _ops()
    x = torch.randn(64, 60, device="cuda").to(BF16)
    w = torch.randn(64, 60, device="cuda").to(BF16)
    out = torch.empty(64, 64, device="cuda", dtype=BF16)
    with pytest.raises(_lib.PMGTError):
        ops.linear_fwd(x, w, None, out)  # lda not a multiple of 8
