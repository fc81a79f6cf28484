"""GPU parity: dense row kernels (tcgen05 bf16 / FFMA fp32) and the FinalAttention row transform.

The floating-point reference for a single contraction is plain torch in fp32/fp64 on the SAME
(already rounded) operands; the row transform is checked against the CPU oracle.
"""
import math

import numpy as np
import pytest
import torch

from oracle import oracle
from news_recommendation_project_v2_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from news_recommendation_project_v2_b200 import ops as _ops
    return _ops


def _ref_linear(a, w, bias, epi, res):
    y = a.double() @ w.double().T
    if bias is not None:
        y = y + bias.double()
    if epi == 1:
        y = torch.relu(y)
    elif epi == 2:
        y = torch.exp(y)
    elif epi == 3:
        y = y + res.double()
    elif epi == 4:
        y4 = y.reshape(y.shape[0], -1, 4)  # interleaved in pairs (a0,a1,g0,g1,a2,a3,g2,g3,...)
        a_, g_ = y4[:, :, 0:2].reshape(y.shape[0], -1), y4[:, :, 2:4].reshape(y.shape[0], -1)
        y = a_ * 0.5 * g_ * (1 + torch.erf(g_ / math.sqrt(2)))
    return y


SHAPES = [(128, 256, 64), (1, 256, 128), (300, 768, 1024), (1000, 4096, 768), (257, 128, 256), (129, 64, 64),
          (4099, 1024, 4096), (640, 6144, 768)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("epi", [0, 1, 2, 3, 4])
def test_linear_bf16_tcgen05(ops, M, N, K, epi):
    g = torch.Generator().manual_seed(M * 7 + N + K + epi)
    a = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, generator=g) * 0.1
    res = torch.randn(M, N, generator=g) if epi == 3 else None
    want = _ref_linear(a, w, bias, epi, res)
    for out_dtype in (torch.float32, torch.bfloat16):
        y = ops.linear(a.cuda(), w.cuda(), bias.cuda(), epi, None if res is None else res.cuda(), out_dtype)
        assert y.dtype == out_dtype and y.shape == want.shape
        tol = 2e-4 if out_dtype == torch.float32 else 1.2e-2  # fp32 accumulate / one bf16 rounding (2^-8 rel)
        torch.testing.assert_close(y.cpu().double(), want, atol=tol, rtol=tol)
    # no-bias path
    y = ops.linear(a.cuda(), w.cuda(), None, 0 if epi != 4 else 4, None, torch.float32)
    torch.testing.assert_close(y.cpu().double(), _ref_linear(a, w, None, 0 if epi != 4 else 4, None), atol=2e-4, rtol=2e-4)


@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (1, 8, 16), (300, 768, 1024), (257, 4096, 768), (131, 136, 48)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3, 4])
def test_linear_fp32_ffma(ops, M, N, K, epi):
    g = torch.Generator().manual_seed(M + N * 3 + K + epi)
    a = torch.randn(M, K, generator=g) * 0.5
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    bias = torch.randn(N, generator=g) * 0.1
    res = torch.randn(M, N, generator=g) if epi == 3 else None
    want = _ref_linear(a, w, bias, epi, res)
    y = ops.linear(a.cuda(), w.cuda(), bias.cuda(), epi, None if res is None else res.cuda(), torch.float32)
    torch.testing.assert_close(y.cpu().double(), want, atol=2e-5, rtol=2e-5)  # fp32 FFMA accumulation
    yb = ops.linear(a.cuda(), w.cuda(), bias.cuda(), epi, None if res is None else res.cuda(), torch.bfloat16)
    torch.testing.assert_close(yb.cpu().double(), want, atol=1.2e-2, rtol=1.2e-2)


@pytest.mark.parametrize("precision,tol", [(torch.float32, 2e-5), (torch.bfloat16, 4e-2)])
def test_final_attention_rows_vs_oracle(ops, precision, tol):
    dim, hidden, n = 768, 4096, 700
    sd = syn.make_final_attention_state_dict(dim, hidden, seed=21)
    table = syn.make_table(n, dim, seed=22)
    want_x, want_logit = oracle.final_attention_rows(sd, table, dtype=torch.float64)
    w = {k: (v.to(precision) if k.endswith("weight") else v.float()).cuda().contiguous() for k, v in sd.items()}
    x, e = ops.final_attention_rows(table.to(precision).cuda(), w, torch.float32)
    # fp32: FFMA accumulation noise; bf16: three to five chained bf16 contractions (rel 2^-8 each)
    torch.testing.assert_close(x.cpu().double(), want_x, atol=tol, rtol=tol)
    torch.testing.assert_close(e.cpu().double(), torch.exp(want_logit), atol=tol, rtol=tol)
    xb, eb = ops.final_attention_rows(table.to(precision).cuda(), w, torch.bfloat16)
    torch.testing.assert_close(xb.float().cpu().double(), want_x, atol=max(tol, 1e-2), rtol=max(tol, 1e-2))
    assert eb.dtype == torch.bfloat16


@pytest.mark.parametrize("M,N,K,group,valid", [
    (300, 512, 128, 32, 32), (300, 512, 128, 64, 40), (257, 512, 64, 128, 128), (130, 512, 128, 128, 100),
    (300, 512, 128, 256, 256), (300, 768, 128, 256, 130),          # group == one 256-column tile
    (700, 1024, 256, 512, 512), (129, 4096, 768, 512, 512),        # cluster of 2 (BASELINE L=512)
    (520, 2048, 128, 1024, 1000),                                  # cluster of 4 (BASELINE cfg 5, L=1024)
])
def test_linear_softmax_epilogue_cluster(ops, M, N, K, group, valid):
    """Per-group softmax fused behind the contraction; groups wider than a tile span a CTA cluster (DSMEM)."""
    g = torch.Generator().manual_seed(M + N + K + group + valid)
    a = (torch.randn(M, K, generator=g)).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * (4.0 / math.sqrt(K))).to(torch.bfloat16)  # logits with a real spread
    logits = (a.double() @ w.double().T).reshape(M, N // group, group)
    logits[:, :, valid:] = -float("inf")
    want = torch.softmax(logits, dim=-1).reshape(M, N)
    y = ops.linear(a.cuda(), w.cuda(), None, 5, None, torch.bfloat16, group=group, group_valid=valid)
    got = y.float().cpu().double()
    assert torch.all(got.reshape(M, N // group, group)[:, :, valid:] == 0)
    torch.testing.assert_close(got, want, atol=2e-3, rtol=1.0e-2)  # one bf16 rounding of p in [0,1]
    torch.testing.assert_close(got.reshape(M, N // group, group).sum(-1), torch.ones(M, N // group, dtype=torch.float64),
                               atol=6e-3, rtol=0)
