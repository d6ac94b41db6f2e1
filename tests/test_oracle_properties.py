"""CPU: algebraic invariants of the oracle on random shapes -- the same properties the GPU tests assert at
full BASELINE sizes, checked here on the checker itself."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

import nfdpf_oracle as O


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 6), st.integers(1, 300), st.integers(0, 10 ** 6))
def test_cascade_row_sum_equals_torch_sum(B, N, seed):
    if torch.backends.cpu.get_cpu_capability() not in ("AVX2", "AVX512"):
        return
    g = torch.Generator().manual_seed(seed)
    q = torch.rand(B, N, generator=g) * 0.9 + 1e-3
    assert np.array_equal(q.sum(-1).numpy(), O.cascade_row_sum(q.numpy()))


@settings(max_examples=15, deadline=None)
@given(st.sampled_from([(2, 4), (2, 36), (4, 3), (32, 32), (2, 0)]), st.integers(1, 40), st.integers(0, 10 ** 6))
def test_flow_roundtrip_and_logdet(shape, P, seed):
    D, C = shape
    g = torch.Generator().manual_seed(seed)
    flows = O.unpack_stack(O.init_stack(g, D, C, std=0.2, bias_std=0.1), D, C)
    x = torch.randn(P, D, generator=g)
    ctx = torch.randn(P, C, generator=g) if C else None
    z, ld = O.stack_forward(x, ctx, flows)
    xr, ldi = O.stack_inverse(z, ctx, flows)
    assert torch.allclose(xr, x, atol=1e-4) and torch.allclose(ldi, -ld, atol=1e-5)


@settings(max_examples=15, deadline=None)
@given(st.integers(1, 5), st.integers(1, 200), st.floats(0.05, 1.0), st.integers(0, 10 ** 6))
def test_soft_resample_invariants(B, N, alpha, seed):
    g = torch.Generator().manual_seed(seed)
    w = torch.softmax(torch.randn(B, N, generator=g) * 3, -1)
    p = torch.randn(B, N, 2, generator=g)
    off = torch.rand(B, generator=g) / N
    pr, wr, idx = O.soft_resample(p, w, alpha, off)
    loc = idx - N * torch.arange(B)[:, None]
    assert int(loc.min()) >= 0 and int(loc.max()) < N and bool((loc[:, 1:] >= loc[:, :-1]).all())
    assert torch.allclose(wr.sum(-1), torch.ones(B), atol=1e-5)
    assert torch.equal(pr, p.reshape(B * N, 2)[idx])


def test_ot_plan_invariants():
    g = torch.Generator().manual_seed(3)
    B, N = 3, 50
    w = torch.softmax(torch.randn(B, N, generator=g) * 2, -1)
    x = torch.randn(B, N, 2, generator=g) * 10
    T, iters = O.ot_transport(x, w.log())
    assert 3 <= iters <= 100
    assert torch.allclose(T.sum(1), (w * N).double(), rtol=1e-6)          # column sums = N w_j
    assert float((T.sum(2) - 1).abs().max()) < 0.2                        # row sums ~ 1
    p, wr, idx = O.ot_resample(x, w)
    assert torch.allclose(p.mean(1), (w[..., None] * x).sum(1), atol=1e-3)
    assert torch.equal(wr, torch.full_like(w, 1.0 / N))
