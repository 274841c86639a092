"""GPU test of rb_noisy_compose (model.py:43-44, W = mu + sigma * eps): the one exported entry point no class of the product
calls (it exists for a reference-side NoisyLinear.forward binding).  Bit-exact against the two-rounding
numpy expression.  Written after the round's GPU budget was spent, so it sorts last among the GPU test files."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("count,offset", [(512 * 3136, 0), (1001, 0), (1000, 1), (4, 0), (1, 0)])
def test_noisy_compose_bit_exact(count, offset):
    from rainbow_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(count + offset)
    host = [rs.standard_normal(count + offset).astype(np.float32) * s for s in (0.02, 0.005, 1.0)]
    dev = [torch.from_numpy(h).cuda()[offset:] for h in host]          # offset 1: not 16-byte aligned -> scalar path
    out = torch.empty(count, dtype=torch.float32, device="cuda")
    _lib.check(lib.rb_noisy_compose(_lib.ptr(dev[0]), _lib.ptr(dev[1]), _lib.ptr(dev[2]), count, _lib.ptr(out), _lib.stream()))
    mu, sigma, eps = (h[offset:] for h in host)
    want = mu + sigma * eps                                              # fl32(mu + fl32(sigma * eps)), like the reference on CPU
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert lib.rb_noisy_compose(_lib.ptr(dev[0]), _lib.ptr(dev[1]), _lib.ptr(dev[2]), 0, _lib.ptr(out), _lib.stream()) == -22
