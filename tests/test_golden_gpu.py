"""CUDA path vs the fixtures produced by the reference's own files (tests/golden/make_golden.py)."""
import pytest
import torch

import efficient_gnns_b200  # noqa: F401
from conftest import rel_err
from efficient_gnns_b200 import ops
from efficient_gnns_b200.engine import GCNStudentTrainer
from efficient_gnns_b200.sparse import SparseTensor

pytestmark = pytest.mark.gpu


def test_kd_kernel_reproduces_reference_kd_criterion(golden_criterion):
    i, case = golden_criterion["inputs"], golden_criterion["cases"]["kd"]
    out, dz = ops.kd_loss_fwd_bwd(i["logits"].cuda(), i["labels"].cuda(), None, i["t_logits"].cuda(), 0.9, 4.0)
    out = out.cpu()
    assert abs(out[0] - case["loss"]) < 1e-5 * abs(case["loss"])
    assert abs(out[1] - case["loss_cls"]) < 1e-5 * abs(case["loss_cls"])
    assert abs(out[2] - case["loss_aux"]) < 1e-5 * abs(case["loss_aux"])
    assert rel_err(dz, case["d_logits"]) < 1e-5


def test_engine_reproduces_reference_gcn_module(golden_model):
    """The reference's GCN class (arxiv_pyg/gnn.py:23-53), dropout 0, train mode: logits, out_feat, loss, grads."""
    G, m = golden_model, golden_model["models"]["gcn"]
    n = G["x"].shape[0]
    adj = SparseTensor(row=G["sym_row"].cuda(), col=G["sym_col"].cuda(), sparse_sizes=(n, n), is_sorted=True)
    tr = GCNStudentTrainer(adj, [16, 32, 32, 8], dropout=0.0)
    tr.load_state_dict({k: v.cuda() for k, v in m["state"].items() if "num_batches" not in k})
    loss = tr.train_step(G["x"].cuda(), m["y"].cuda(), G["train_idx"].cuda(), None).cpu()
    assert rel_err(tr.Y[-1], m["logits_train"]) < 1e-5
    assert rel_err(tr.A[-1], m["out_feat"]) < 1e-5
    assert abs(loss[0] - m["loss"]) < 1e-5 * abs(m["loss"])
    for l in range(3):
        assert rel_err(tr.gW[l], m["grads"][f"convs.{l}.weight"]) < 5e-5
    assert rel_err(tr.gb[2], m["grads"]["convs.2.bias"]) < 5e-5
    for l in range(2):
        assert rel_err(tr.ggamma[l], m["grads"][f"bns.{l}.weight"]) < 5e-5
        assert rel_err(tr.gbeta[l], m["grads"][f"bns.{l}.bias"]) < 5e-5
