"""tcgen05 (3xTF32) MLP-layer kernel against an fp64 reference; also asserts the tensor-core path
is the one dispatched (p2c_linear_path), so a silent SIMT fallback cannot pass these tests."""
import pytest
import torch

from point2cyl_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


SHAPES = [  # M, N, K, pool
    (128, 64, 64, 0), (1024, 64, 64, 0), (4096, 128, 64, 64), (2048, 128, 131, 0), (8192, 128, 128, 0),
    (4096, 256, 128, 64), (1000, 128, 128, 0), (3 * 128 + 17, 64, 40, 0), (2048, 128, 128, 128), (1024, 32, 128, 32),
    (148 * 128 * 3, 128, 128, 0),
]


@pytest.mark.parametrize("M,N,K,pool", SHAPES)
@pytest.mark.parametrize("prologue", [False, True])
def test_linear_tc(M, N, K, pool, prologue):
    g = torch.Generator().manual_seed(M * 7 + N + K)
    ld = ops.pad4(K)
    assert _lib.load().p2c_linear_path(ld, M, N, K, 0, pool, _lib.PREC_3XTF32, 0) == 1
    X = torch.zeros(M, ld)
    X[:, :K] = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    sc = torch.randn(K, generator=g) if prologue else None
    sh = torch.randn(K, generator=g) if prologue else None
    A = X[:, :K].double()
    if prologue:
        A = torch.relu(A * sc.double() + sh.double())
    ref = A @ W.double().t() + b.double()
    stats = torch.zeros(2 * N, dtype=torch.float64, device=DEV)
    res = ops.linear(X.to(DEV), W.to(DEV), b.to(DEV), K=K, in_scale=None if sc is None else sc.to(DEV),
                     in_shift=None if sh is None else sh.to(DEV), stats=stats, pool_group=pool, want_y=True,
                     precision=_lib.PREC_3XTF32)
    torch.cuda.synchronize()
    Y = res[0] if pool else res
    assert rel_err(Y, ref) <= 2e-6
    assert rel_err(stats[:N], ref.sum(0)) <= 1e-5
    assert rel_err(stats[N:], (ref ** 2).sum(0)) <= 1e-5
    if pool:
        assert rel_err(res[1], ref.reshape(M // pool, pool, N).max(1).values) <= 2e-6
        assert rel_err(res[2], ref.reshape(M // pool, pool, N).min(1).values) <= 2e-6


SS_SHAPES = [(4096, 256, 259, 0), (4096, 512, 256, 0), (4096, 1024, 512, 128), (4096, 256, 1280, 0),
             (16384, 256, 384, 0), (16384, 128, 256, 0), (1000, 200, 300, 0), (512, 320, 224, 64)]


@pytest.mark.parametrize("M,N,K,pool", SS_SHAPES)
@pytest.mark.parametrize("prologue", [False, True])
def test_linear_tc_streamed_weights(M, N, K, pool, prologue):
    g = torch.Generator().manual_seed(M + 3 * N + K)
    ld = ops.pad4(K)
    assert _lib.load().p2c_linear_path(ld, M, N, K, 0, pool, _lib.PREC_3XTF32, 1) == 2
    X = torch.zeros(M, ld)
    X[:, :K] = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    sc = torch.randn(K, generator=g) if prologue else None
    sh = torch.randn(K, generator=g) if prologue else None
    A = X[:, :K].double()
    if prologue:
        A = torch.relu(A * sc.double() + sh.double())
    ref = A @ W.double().t() + b.double()
    stats = torch.zeros(2 * N, dtype=torch.float64, device=DEV)
    Wd = W.to(DEV)
    res = ops.linear(X.to(DEV), Wd, b.to(DEV), K=K, in_scale=None if sc is None else sc.to(DEV),
                     in_shift=None if sh is None else sh.to(DEV), stats=stats, pool_group=pool, want_y=True,
                     precision=_lib.PREC_3XTF32, w_split=ops.split_tf32(Wd))
    torch.cuda.synchronize()
    Y = res[0] if pool else res
    assert rel_err(Y, ref) <= 1e-5      # 3xTF32 error grows ~sqrt(K); K is up to 1280 here
    assert rel_err(stats[:N], ref.sum(0)) <= 1e-5
    assert rel_err(stats[N:], (ref ** 2).sum(0)) <= 3e-5
    if pool:
        assert rel_err(res[1], ref.reshape(M // pool, pool, N).max(1).values) <= 1e-5
        assert rel_err(res[2], ref.reshape(M // pool, pool, N).min(1).values) <= 1e-5


@pytest.mark.parametrize("M,K,N,pool", [(4096, 128, 128, 0), (5000, 64, 64, 0), (8192, 259, 256, 64), (4096, 1280, 256, 0)])
def test_linear_bf16_path(M, K, N, pool):
    """P2C_PREC_BF16: one kind::f16 pass on bf16 operands (BASELINE.json configs[2]); compared with the same product of
    bf16-rounded operands in float64 (tight) and with the fp32 product (bf16-level tolerance)."""
    g = torch.Generator().manual_seed(K)
    ld = ops.pad4(K)
    X = torch.randn(M, ld, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    sc, sh = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.2
    Xd, Wd = X.to(DEV), W.to(DEV)
    wop = ops.weight_operand(Xd, Wd, N, K, False, pool, _lib.PREC_BF16)
    assert wop is not None and wop.dtype == torch.bfloat16
    stats = torch.zeros(2 * N, dtype=torch.float64, device=DEV)
    res = ops.linear(Xd, Wd, b.to(DEV), K=K, in_scale=sc.to(DEV), in_shift=sh.to(DEV), stats=stats, pool_group=pool,
                     precision=_lib.PREC_BF16, w_split=wop)
    Y = res[0] if pool else res
    A = torch.relu(X[:, :K] * sc + sh)
    ref_bf = A.bfloat16().double() @ W.bfloat16().double().T + b.double()
    ref32 = A.double() @ W.double().T + b.double()
    # (the operand fold is an fma on the device and mul+add on the host: a one-ulp fp32 difference now and then flips
    # a bf16 rounding, hence 2e-3 rather than accumulation-order noise)
    assert rel_err(Y, ref_bf) <= 2e-3
    assert rel_err(Y, ref32) <= 2e-2
    assert rel_err(stats[:N], Y.double().sum(0)) <= 1e-5
    if pool:
        assert rel_err(res[1], Y.reshape(M // pool, pool, N).max(1)[0]) == 0.0
