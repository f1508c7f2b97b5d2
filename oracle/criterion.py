"""The six distillation criteria, restated from first principles (explicit formulas, CPU PyTorch).

Each mirrors a function of the reference's arxiv_pyg/criterion.py (line ranges cited); the
restatement deliberately spells the math out instead of calling the same torch.nn.functional
helpers, so that agreeing with the reference file (tests/golden) is a real check.

All return (loss, loss_cls, loss_aux) like the reference.  ``sampled_inds`` replaces the
reference's ``np.random.choice`` draw (criterion.py:63,135) so tests can inject the sample.
"""
from __future__ import annotations

from typing import Optional

import torch

from .ops import segment_softmax


def _log_softmax(z):
    m = z.max(dim=1, keepdim=True).values
    s = z - m
    return s - s.exp().sum(dim=1, keepdim=True).log()


def cross_entropy(logits, labels):
    """mean over rows of -log_softmax(z)[y]."""
    ls = _log_softmax(logits)
    return -(ls.gather(1, labels.view(-1, 1)).squeeze(1)).mean()


def _kl_elementwise_mean(log_q, p):
    """F.kl_div(log_q, p, log_target=False) with the default 'mean' reduction = mean over ALL elements
    of p*log(p) - p*log_q, with 0*log(0) := 0."""
    plogp = torch.where(p > 0, p * p.clamp_min(1e-45).log(), torch.zeros_like(p))
    return (plogp - p * log_q).mean()


def _safe_sqrt(d2):
    """sqrt with the subgradient 0 at 0 — what torch's `.norm(p=2)` backward does (the reference's l2 kernels
    hit exact zeros: the diagonal of the pairwise matrix, criterion.py:80-81)."""
    return torch.where(d2 > 0, d2.clamp_min(1e-38).sqrt(), torch.zeros_like(d2))


def _l2_normalize(x, eps: float = 1e-12):
    return x / x.pow(2).sum(-1, keepdim=True).sqrt().clamp_min(eps)


def kd_criterion(logits, labels, teacher_logits, alpha: float = 0.9, T: float = 4.0):
    """criterion.py:8-21 — CE + KL(softmax(t/T) || softmax(z/T)), elementwise mean, scaled alpha*T^2."""
    loss_cls = cross_entropy(logits, labels)
    log_q = _log_softmax(logits / T)
    p = _log_softmax(teacher_logits / T).exp()
    loss_kd = _kl_elementwise_mean(log_q, p)
    return loss_kd * (alpha * T * T) + loss_cls * (1 - alpha), loss_cls, loss_kd


def fitnet_criterion(logits, labels, feat, teacher_feat, beta: float = 1000):
    """criterion.py:24-36 — MSE between L2-normalised features."""
    loss_cls = cross_entropy(logits, labels)
    d = _l2_normalize(feat) - _l2_normalize(teacher_feat)
    loss_aux = d.pow(2).mean()
    return loss_cls + beta * loss_aux, loss_cls, loss_aux


def at_criterion(logits, labels, feat, teacher_feat, beta: float = 1000):
    """criterion.py:39-54 — per-node squared norms, each [n] vector L2-normalised, MSE."""
    loss_cls = cross_entropy(logits, labels)
    a = _l2_normalize(feat.pow(2).sum(-1))
    b = _l2_normalize(teacher_feat.pow(2).sum(-1))
    loss_aux = (a - b).pow(2).mean()
    return loss_cls + beta * loss_aux, loss_cls, loss_aux


def _pairwise(feat, kernel: str):
    if kernel in ("cosine", "poly"):
        f = _l2_normalize(feat)
        s = f @ f.t()
        return s if kernel == "cosine" else s * s
    if kernel == "l2":
        # reference broadcasts (f_i - f_j) and takes the norm; do the same difference form to keep its rounding
        d = feat.unsqueeze(0) - feat.unsqueeze(1)
        return _safe_sqrt(d.pow(2).sum(-1))
    if kernel == "rbf":
        d = feat.unsqueeze(0) - feat.unsqueeze(1)
        return torch.exp(-0.5 * d.pow(2).sum(-1))
    raise NotImplementedError(kernel)


def gpw_criterion(logits, labels, feat, teacher_feat, kernel: str = "cosine", beta: float = 1,
                  max_samples: int = 8192, sampled_inds: Optional[torch.Tensor] = None):
    """criterion.py:57-92 — GSP: MSE between all-pairs similarity matrices of a row sample."""
    loss_cls = cross_entropy(logits, labels)
    if max_samples < feat.shape[0]:
        assert sampled_inds is not None, "inject the sample (reference draws it with np.random.choice)"
        feat, teacher_feat = feat[sampled_inds], teacher_feat[sampled_inds]
    loss_aux = (_pairwise(feat, kernel) - _pairwise(teacher_feat, kernel)).pow(2).mean()
    return loss_cls + beta * loss_aux, loss_cls, loss_aux


def _edge_similarity(feat, src, dst, kernel: str):
    a, b = feat.index_select(0, src), feat.index_select(0, dst)
    if kernel in ("cosine", "poly"):
        # F.cosine_similarity, eps=1e-8, torch>=1.12 form: each norm clamped separately
        na = a.pow(2).sum(-1).sqrt().clamp_min(1e-8)
        nb = b.pow(2).sum(-1).sqrt().clamp_min(1e-8)
        c = (a * b).sum(-1) / (na * nb)
        return c if kernel == "cosine" else c * c
    d2 = (a - b).pow(2).sum(-1)
    if kernel == "l2":
        return _safe_sqrt(d2)
    if kernel == "rbf":
        return torch.exp(-0.5 * d2)
    raise NotImplementedError(kernel)


def lpw_criterion(logits, labels, feat, teacher_feat, edge_index, kernel: str = "cosine", beta: float = 100,
                  criterion: str = "kld"):
    """criterion.py:95-126 — LSP: per-edge similarity, PyG softmax grouped by dst (the unsorted COO column),
    KL(teacher || student) with elementwise-mean reduction over the E edges (or MSE)."""
    loss_cls = cross_entropy(logits, labels)
    src, dst = edge_index[0], edge_index[1]
    ps = segment_softmax(_edge_similarity(feat, src, dst, kernel), dst)
    pt = segment_softmax(_edge_similarity(teacher_feat, src, dst, kernel), dst)
    if criterion == "mse":
        loss_aux = (ps - pt).pow(2).mean()
    elif criterion == "kld":
        loss_aux = _kl_elementwise_mean(ps.log(), pt)
    else:
        raise NotImplementedError(criterion)
    return loss_cls + beta * loss_aux, loss_cls, loss_aux


def nce_criterion(logits, labels, feat, teacher_feat, beta: float = 0.5, nce_T: float = 0.075,
                  max_samples: int = 8192, sampled_inds: Optional[torch.Tensor] = None):
    """criterion.py:129-149 — G-CRD: InfoNCE between normalised student rows and teacher rows, positives on the diagonal."""
    loss_cls = cross_entropy(logits, labels)
    if max_samples < feat.shape[0]:
        assert sampled_inds is not None
        feat, teacher_feat = feat[sampled_inds], teacher_feat[sampled_inds]
    z = (_l2_normalize(feat) @ _l2_normalize(teacher_feat).t()) / nce_T
    ls = _log_softmax(z)
    loss_aux = -ls.diagonal().mean()
    return loss_cls + beta * loss_aux, loss_cls, loss_aux


# ----------------------------------------------------------------------------------------------------------------------
# ppi_pyg/criterion.py — the multi-label variant: the classification term is binary cross-entropy with logits and
# kd_criterion distils through sigmoid(teacher_logits) (:8-19); the auxiliary terms are the ones above.
def bce_with_logits(logits, targets):
    """F.binary_cross_entropy_with_logits, mean over all elements: max(z,0) - z t + log(1 + exp(-|z|))."""
    z = logits
    return (z.clamp_min(0) - z * targets + torch.log1p(torch.exp(-z.abs()))).mean()


def kd_criterion_ppi(logits, labels, teacher_logits, alpha: float = 0.5, T: float = 1):
    """ppi_pyg/criterion.py:8-19."""
    loss_cls = bce_with_logits(logits, labels)
    loss_kd = bce_with_logits(logits, torch.sigmoid(teacher_logits))
    return loss_kd * (alpha * T * T) + loss_cls * (1 - alpha), loss_cls, loss_kd
