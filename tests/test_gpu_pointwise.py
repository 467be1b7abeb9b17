"""Kernel-level parity of the 1x1 convolution kernels (tcgen05 3xFP16 / 3xTF32 and CUDA-core fp32) against a
float64 torch reference, on every (K, N) pair the EfficientNet-B3 trunk uses plus ragged M."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    from cosypose_b200.engine import Engine
    return Engine(0, max_batch=1)


def _trunk_shapes():
    from cosypose_b200 import effnet_spec as spec
    shapes = set()
    for b in spec.BLOCKS:
        if b.e != 1:
            shapes.add((b.cin, b.cexp, 'expand'))
        shapes.add((b.cexp, b.cout, 'project'))
    shapes.add((384, 1536, 'head'))
    return sorted(shapes)


def _ref(A, W, bias, gate, rows, resid, swish):
    A64 = A.double()
    if gate is not None:
        A64 = A64 * gate.double().repeat_interleave(rows, dim=0)[:A.shape[0]]
    y = A64 @ W.double().t() + bias.double()
    if swish:
        y = y * torch.sigmoid(y)
    if resid is not None:
        y = y + resid.double()
    return y


@pytest.mark.parametrize('impl,groups', [(2, 0), (1, 1), (1, 2), (0, 0)])
def test_all_trunk_shapes(eng, impl, groups):
    """groups: producer-warpgroup variant of the tensor-core kernel (1: two CTAs per SM, 2: one CTA per SM)."""
    dev = eng.device
    eng.set_option('tc_groups', groups)
    gen = torch.Generator().manual_seed(0)
    worst = 0.0
    for K, N, kind in _trunk_shapes():
        rows = 70
        M = 3 * rows + 17                     # ragged: not a multiple of the 128-row tile
        A = torch.randn((M, K), generator=gen)
        W = torch.randn((N, K), generator=gen) / np.sqrt(K)
        bias = torch.randn(N, generator=gen)
        gate = torch.rand((4, K), generator=gen) if kind == 'project' else None
        resid = torch.randn((M, N), generator=gen) if kind == 'project' and K == 6 * N else None
        swish = kind != 'project'
        out = eng.debug_pointwise(impl, A.to(dev), W, bias, gate.to(dev) if gate is not None else None, rows,
                                  resid.to(dev) if resid is not None else None, swish)
        ref = _ref(A, W, bias, gate, rows, resid, swish)
        err = (out.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
        worst = max(worst, err)
        assert err < 2e-6, (K, N, kind, err)
    eng.set_option('tc_groups', 0)
    print('worst relative error', worst)


@pytest.mark.parametrize('impl,groups', [(2, 0), (1, 1), (1, 2)])
@pytest.mark.parametrize('M', [1, 127, 128, 129, 4480, 19200 * 2 + 5])
def test_row_counts(eng, M, impl, groups):
    dev = eng.device
    eng.set_option('tc_groups', groups)
    gen = torch.Generator().manual_seed(M)
    K, N = 48, 288
    A = torch.randn((M, K), generator=gen)
    W = torch.randn((N, K), generator=gen) / np.sqrt(K)
    bias = torch.randn(N, generator=gen)
    out = eng.debug_pointwise(impl, A.to(dev), W, bias, swish=True)
    ref = _ref(A, W, bias, None, 1, None, True)
    eng.set_option('tc_groups', 0)
    assert (out.cpu().double() - ref).abs().max().item() < 2e-6 * ref.abs().max().item()


@pytest.mark.parametrize('impl,groups', [(2, 0), (1, 1), (1, 2)])
@pytest.mark.parametrize('M,K,N,kind', [(19200, 816, 136, 'project'), (4480, 1392, 232, 'project'),
                                        (4480, 232, 1392, 'expand'), (76800, 192, 32, 'project'),
                                        (19200, 96, 576, 'expand')])
def test_repeatable(eng, M, K, N, kind, impl, groups):
    """Bit-identical results over repeated launches at the benchmark sizes (every barrier hand-off in the
    pipelined kernel is exercised thousands of times per launch: a missing dependency shows up here)."""
    dev = eng.device
    eng.set_option('tc_groups', groups)
    gen = torch.Generator().manual_seed(K + N)
    rows = 300 if M == 19200 else 70
    A = torch.randn((M, K), generator=gen).to(dev)
    W = torch.randn((N, K), generator=gen) / np.sqrt(K)
    bias = torch.randn(N, generator=gen)
    gate = torch.rand((-(-M // rows), K), generator=gen).to(dev) if kind == 'project' else None
    resid = torch.randn((M, N), generator=gen).to(dev) if kind == 'project' and K == 6 * N else None
    first = eng.debug_pointwise(impl, A, W, bias, gate, rows, resid, kind != 'project').clone()
    ref = _ref(A.cpu(), W, bias, gate.cpu() if gate is not None else None, rows,
               resid.cpu() if resid is not None else None, kind != 'project')
    assert (first.cpu().double() - ref).abs().max().item() < 2e-6 * ref.abs().max().item()
    for _ in range(8):
        again = eng.debug_pointwise(impl, A, W, bias, gate, rows, resid, kind != 'project')
        assert torch.equal(first, again)
    eng.set_option('tc_groups', 0)
