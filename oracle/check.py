"""TEST INFRASTRUCTURE — the reference's GCN + logit-KD training step restated on the CPU in one call, used as the
checker by tests/, by __graft_entry__.smoke() and by bench.py's parity leg.  Nothing under efficient-gnns_b200/ imports it.

Follows arxiv_pyg/gnn.py:45-53 (GCN.forward), :102-195 (train(): forward, [train_idx] gather, criterion, backward) and
arxiv_pyg/criterion.py:8-21 (kd_criterion), through oracle.nn / oracle.criterion.

``act_masks`` lets a test impose the ACTIVATION PATTERN (ReLU-active AND kept-by-dropout, one bool matrix per hidden
layer) instead of deriving it from the sign of the pre-activation: an fp32 engine and an fp64 oracle legitimately
disagree on the sign of pre-activations that sit within rounding distance of zero, and a single such flip moves a few
gradient entries by far more than 1e-5.  With the pattern imposed the gradient comparison isolates arithmetic error; the
flips themselves are counted and bounded separately (`activation_flips`).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import criterion as oc, nn as onn


def gcn_kd_reference_step(x, rowptr, col, val, state: Dict[str, torch.Tensor], y, teacher, idx, alpha: float, T: float,
                          p: float, drop_masks: Optional[List[torch.Tensor]] = None,
                          act_masks: Optional[List[torch.Tensor]] = None, dtype=torch.float64, kd: bool = True):
    """One training step's forward + loss + backward.  ``state`` uses the reference module's state_dict keys
    (convs.i.weight [in,out], convs.i.bias, bns.i.weight, bns.i.bias).  Returns a dict with logits, the hidden
    pre-activations (BatchNorm outputs), the three losses and the gradients in parameter order
    (W0, b0, gamma0, beta0, W1, ..., W_last, b_last)."""
    L = sum(1 for k in state if k.startswith("convs.") and k.endswith(".weight"))
    cast = lambda t: t.detach().cpu().to(dtype)
    W = [cast(state[f"convs.{i}.weight"]).requires_grad_(True) for i in range(L)]
    B = [cast(state[f"convs.{i}.bias"]).requires_grad_(True) for i in range(L)]
    ga = [cast(state[f"bns.{i}.weight"]).requires_grad_(True) for i in range(L - 1)]
    be = [cast(state[f"bns.{i}.bias"]).requires_grad_(True) for i in range(L - 1)]
    v = val.to(dtype)
    h = x.to(dtype)
    pre = []
    for i in range(L - 1):
        h = onn.gcn_conv(h, rowptr, col, v, W[i], B[i])
        h = onn.batch_norm_train(h, ga[i], be[i])
        pre.append(h.detach())
        if act_masks is not None:
            h = h * act_masks[i].to(dtype)
        else:
            h = torch.relu(h)
            if drop_masks is not None:
                h = h * drop_masks[i].to(dtype)
        if p > 0 and (drop_masks is not None or act_masks is not None):
            h = h / (1.0 - p)
    hidden = h
    logits = onn.gcn_conv(h, rowptr, col, v, W[-1], B[-1])
    if kd:
        loss, lc, la = oc.kd_criterion(logits[idx], y[idx], teacher[idx].to(dtype), alpha, T)
    else:
        loss = lc = oc.cross_entropy(logits[idx], y[idx])
        la = loss * 0
    params = []
    for i in range(L):
        params += [W[i], B[i]]
        if i < L - 1:
            params += [ga[i], be[i]]
    grads = torch.autograd.grad(loss, params)
    return {"logits": logits.detach(), "hidden": hidden.detach(), "pre": pre,
            "loss": (float(loss.detach()), float(lc.detach()), float(la.detach())), "grads": [g.detach() for g in grads], "params": params}


def activation_flips(engine_act: torch.Tensor, pre: torch.Tensor, drop_mask: Optional[torch.Tensor]):
    """(#elements whose activation pattern differs between the engine and sign(pre)&drop_mask,
        largest |pre| among them relative to max|pre|)."""
    want = pre > 0
    if drop_mask is not None:
        want = want & drop_mask
    diff = engine_act != want
    n = int(diff.sum())
    worst = float(pre[diff].abs().max() / pre.abs().max()) if n else 0.0
    return n, worst


def rel_max(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    d = b.abs().max().item()
    return (a - b).abs().max().item() / (d if d > 0 else 1.0)


def rel_fro(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    d = b.norm().item()
    return (a - b).norm().item() / (d if d > 0 else 1.0)


def compare_engine_step(tr, x, y, teacher, idx, rowptr, col, val, drop_masks, state_before, kd: bool = True) -> dict:
    """Run AFTER ``tr.train_step`` on the same inputs: compares the engine's logits / losses / gradients with the fp64
    restatement (i) free-running and (ii) with the engine's activation pattern imposed.  Returns the numbers; the
    caller asserts."""
    L = tr.L
    acts = [tr.activation_pattern(l).cpu() for l in range(L - 1)]
    got = []
    for l in range(L):
        got += [tr.gW[l].cpu(), tr.gb[l].cpu()]
        if l < L - 1:
            got += [tr.ggamma[l].cpu(), tr.gbeta[l].cpu()]
    loss = tr.loss_out.cpu().tolist()
    out = {}
    for tag, am in (("free", None), ("pattern", acts)):
        ref = gcn_kd_reference_step(x, rowptr, col, val, state_before, y, teacher, idx, tr.alpha, tr.kd_T, tr.p,
                                    drop_masks, am, kd=kd)
        r = {"logits_max": rel_max(tr.Y[-1], ref["logits"]), "hidden_max": rel_max(tr.out_feat(), ref["hidden"]),
             "loss_rel": [abs(a - b) / max(abs(b), 1e-30) for a, b in zip(loss, ref["loss"])][: 3 if kd else 2]}
        scale = max(g.abs().max().item() for g in ref["grads"])
        gm, gf, names = [], [], []
        for i, (a, b) in enumerate(zip(got, ref["grads"])):
            hidden_bias = (i % 4 == 1) and i < 4 * (L - 1)   # a conv bias in front of BatchNorm: exact gradient is 0
            if hidden_bias:
                r.setdefault("hidden_bias_abs_over_scale", []).append(a.abs().max().item() / scale)
                continue
            gm.append(rel_max(a, b)); gf.append(rel_fro(a, b)); names.append(i)
        r["grad_max"], r["grad_fro"], r["grad_index"] = gm, gf, names
        if am is None:
            fl = [activation_flips(acts[l], ref["pre"][l], None if drop_masks is None else drop_masks[l])
                  for l in range(L - 1)]
            r["flips"], r["flip_worst_pre_rel"] = [f[0] for f in fl], [f[1] for f in fl]
            r["elements"] = [int(a.numel()) for a in acts]
        out[tag] = r
        del ref
    return out
