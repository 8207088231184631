"""GPU tests of the CTA-pair Lee kernel (``lee_tc2_kernel``, tcgen05 ``cta_group::2``): it issues the same
products in the same accumulation order as the single-CTA kernel (``SC_LEE_TC_CTA2=0``) and as the variant that
rewrites the staged tile as its TF32 hi part (``SC_LEE_TC_MASK_HI=1``), so all three must agree BIT FOR BIT --
on ragged shapes too (cells not a multiple of the stage, genes not a multiple of the tile, one tile pair only)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from spatialcore_b200 import engine

    return engine


@pytest.mark.parametrize("n,g", [(8192, 256), (20011, 333), (70001, 1030), (513, 40), (9, 5)])
def test_pair_kernel_bit_identical_to_single_cta_kernel(eng, monkeypatch, n, g):
    gen = torch.Generator(device="cuda").manual_seed(n + g)
    ld = eng.padded_ld(g)
    a = torch.randn((n, ld), device="cuda", generator=gen)
    b = 0.3 * torch.randn((n, ld), device="cuda", generator=gen) + 0.05 * a
    a[:, g:] = 0
    b[:, g:] = 0
    monkeypatch.delenv("SC_LEE_TC_CTA2", raising=False)
    monkeypatch.delenv("SC_LEE_TC_MASK_HI", raising=False)
    pair = eng.lee_gemm(a, b, g, impl=2).clone()
    monkeypatch.setenv("SC_LEE_TC_CTA2", "0")
    single = eng.lee_gemm(a, b, g, impl=2).clone()
    monkeypatch.setenv("SC_LEE_TC_MASK_HI", "1")
    masked = eng.lee_gemm(a, b, g, impl=2).clone()
    assert torch.equal(pair[:g, :g], single[:g, :g])
    assert torch.equal(single[:g, :g], masked[:g, :g])
    # and all of them are the contraction: backward-error bar of an FP32 evaluation
    ref = (a.double().T @ b.double())[:g, :g]
    mag = (a.double().abs().T @ b.double().abs())[:g, :g]
    err = (pair.double()[:g, :g] - ref).abs()
    bar = 1e-5 * ref.abs() + 4.0 * 2.0 ** -24 * mag + 1e-30
    assert bool((err <= bar).all()), float((err / bar).max())


def test_pair_kernel_chunk_override(eng, monkeypatch):
    """SC_LEE_TC_CHUNK is rounded to a multiple of every stage length; shorter accumulations stay exact enough."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn((30000, 256), device="cuda", generator=gen)
    b = torch.randn((30000, 256), device="cuda", generator=gen)
    ref = a.double().T @ b.double()
    mag = a.double().abs().T @ b.double().abs()
    for chunk in ("8", "40", "64", "250"):
        monkeypatch.setenv("SC_LEE_TC_CHUNK", chunk)
        L = eng.lee_gemm(a, b, 256, impl=2)
        err = (L.double() - ref).abs()
        assert bool((err <= 1e-5 * ref.abs() + 8.0 * 2.0 ** -24 * mag).all()), chunk
