"""the prototype head einsum('bchw,nc->bnhw', feats, unify_prototype) (lib/models/semseg.py:325-333,
lib/loss/loss_cross_datasets.py:940-969) on the tcgen05 tensor cores, forward and both gradients, against a float64
einsum.  Bars: 1e-5 relative (fp32 inputs, three bf16 terms per operand), 2e-2 (bf16 / fp16 inputs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from mdseg_b200 import ops
    return ops


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("B,K,N,h,w", [(2, 512, 358, 24, 40), (1, 64, 19, 16, 16), (3, 96, 600, 9, 13), (2, 512, 150, 32, 32),
                                      (1, 40, 5, 8, 8)])
@pytest.mark.parametrize("dt,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2), (torch.float16, 2e-2)])
def test_prototype_head_forward_and_gradients(ops, B, K, N, h, w, dt, tol):
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + N)
    feats = (torch.randn(B, K, h, w, generator=g, device=DEV)).to(dt).requires_grad_(True)
    proto = (torch.randn(N, K, generator=g, device=DEV) * 0.2).requires_grad_(True)
    dy = torch.randn(B, N, h, w, generator=g, device=DEV)
    y = ops.prototype_head(feats, proto)
    assert y.dtype == torch.float32 and tuple(y.shape) == (B, N, h, w)
    y.backward(dy)
    f64, p64 = feats.detach().double(), proto.detach().double()
    if dt != torch.float32:  # what autocast does to the reference einsum: the weights rounded to the input type
        p64 = proto.detach().to(dt).double()
    want = torch.einsum("bchw,nc->bnhw", f64, p64)
    assert rel(y, want) <= tol
    assert rel(feats.grad, torch.einsum("bnhw,nc->bchw", dy.double(), p64)) <= tol
    assert rel(proto.grad, torch.einsum("bnhw,bchw->nc", dy.double(), f64)) <= tol
    ops.check_errors(DEV)
