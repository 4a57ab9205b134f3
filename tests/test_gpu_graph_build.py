"""Index lists of column-one-hot 0/1 bi_graphs built on the device (mdseg_graph_build_onehot, no host copy of the
matrix) against the host builder of ops.BipartiteGraphs: the lists themselves, the fused loss and its gradient
(bit-identical: same lists, same kernels), and the error flag for a matrix that is not of the declared kind."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from mdseg_b200 import ops
    return ops


def onehot_graph(c_ds, c_uni, gen, empty_cols=0):
    idx = torch.randint(0, c_ds, (c_uni,), generator=gen)
    idx[:c_ds] = torch.arange(c_ds) if c_uni >= c_ds else idx[:c_ds]
    m = torch.zeros(c_ds, c_uni)
    m[idx, torch.arange(c_uni)] = 1
    if empty_cols:
        m[:, torch.randperm(c_uni, generator=gen)[:empty_cols]] = 0
    return m


@pytest.mark.parametrize("c_ds,c_uni,empty", [(19, 358, 0), (150, 358, 7), (5, 11, 2), (64, 64, 0), (5, 1200, 100), (8, 9, 0)])
def test_device_lists_equal_host_lists(ops, c_ds, c_uni, empty):
    gen = torch.Generator().manual_seed(c_ds * 7 + c_uni)
    g = onehot_graph(c_ds, c_uni, gen, empty).to(DEV)
    host = ops.BipartiteGraphs()._entry(0, g)
    dev = ops.BipartiteGraphs(assume_onehot01=True)._entry(0, g)
    torch.cuda.synchronize()
    ops.check_errors(DEV)
    assert not dev["dense"] and dev["col_onehot"] == 1 and host["col_onehot"] == 1
    nnz = host["nnz"]
    for name in ("csr_ptr", "csc_ptr", "csr4_ptr"):
        assert torch.equal(dev[name].cpu(), host[name].cpu()), name
    assert torch.equal(dev["csr_col"][:nnz].cpu(), host["csr_col"].cpu())
    assert torch.equal(dev["csc_row"][:nnz].cpu(), host["csc_row"].cpu())
    n4 = 4 * int(host["csr4_ptr"][-1])
    assert torch.equal(dev["csr4_col"][:n4].cpu(), host["csr4_col"][:n4].cpu())


def test_loss_and_gradient_identical_with_device_built_graphs(ops):
    gen = torch.Generator().manual_seed(3)
    n_cats, c_uni, ids = [19, 7, 33], 48, [0, 0, 1, 2, 2]
    graphs = [onehot_graph(c, c_uni, gen).to(DEV) for c in n_cats]
    x = torch.randn(len(ids), c_uni, 8, 32, generator=gen).to(DEV)
    labels = torch.stack([torch.randint(0, n_cats[d], (32, 128), generator=gen) for d in ids]).to(DEV)
    ids_t = torch.tensor(ids, dtype=torch.int32, device=DEV)
    out = []
    for cache in (ops.BipartiteGraphs(), ops.BipartiteGraphs(assume_onehot01=True)):
        xi = x.clone().requires_grad_(True)
        loss = ops.mds_proj_ohem_ce(xi, labels, ids_t, graphs, ops.neg_log(0.4), cache=cache)
        loss.backward()
        out.append((loss.detach().clone(), xi.grad.clone()))
    ops.check_errors(DEV)
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])


@pytest.mark.parametrize("kind", ["two_in_a_column", "value_not_one"])
def test_wrong_kind_raises_at_check_errors(ops, kind):
    gen = torch.Generator().manual_seed(5)
    g = onehot_graph(6, 20, gen)
    if kind == "two_in_a_column":
        g[0, 3] = 1; g[1, 3] = 1
    else:
        g[g.argmax(0)[4], 4] = 0.5
    ops.check_errors(DEV)
    ops.BipartiteGraphs(assume_onehot01=True)._entry(0, g.to(DEV))
    with pytest.raises(RuntimeError, match="column-one-hot"):
        ops.check_errors(DEV)
