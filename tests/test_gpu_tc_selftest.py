"""GPU: the tcgen05/TMEM building blocks (operand layout, descriptors, commit, tcgen05.ld)
against a float64 product.  bf16: operands rounded to bf16, fp32 accumulate; tf32x3: fp32-grade."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,n,k", [("bf16", 256, 256), ("bf16", 64, 64), ("bf16", 16, 32),
                                           ("tf32x3", 64, 128), ("tf32x3", 256, 64), ("tf32x3", 32, 32)])
@pytest.mark.parametrize("a_in_tmem", [False, True])
def test_umma_selftest(engine_factory, precision, n, k, a_in_tmem):
    eng = engine_factory(batch_size=64)
    rng = np.random.default_rng(n * 1000 + k)
    A = rng.standard_normal((128, k)).astype(np.float32)
    B = rng.standard_normal((n, k)).astype(np.float32)
    # make layout mistakes loud: every row/column gets its own scale
    A *= (1 + np.arange(128, dtype=np.float32))[:, None] / 64
    B *= (1 + np.arange(n, dtype=np.float32))[:, None] / 32
    D = eng.selftest_umma(A, B, precision, a_in_tmem=a_in_tmem)
    if precision == "bf16":
        import torch
        Ab = torch.from_numpy(A).bfloat16().double().numpy()
        Bb = torch.from_numpy(B).bfloat16().double().numpy()
        ref = Ab @ Bb.T
        tol = 2e-5       # exact bf16 products, fp32 accumulation
    else:
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        tol = 2e-6       # 3-term tf32 split: ~2^-21 per product
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    err = np.max(np.abs(D - ref) / scale)
    assert err <= tol, err
