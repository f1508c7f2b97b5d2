"""Drop-in for the reference's ``ppi_pyg/criterion.py`` (multi-label PPI students): the same six function names,
argument order and defaults as that file (:8-146).  It differs from the arxiv/mag copy only in the classification term —
``F.binary_cross_entropy_with_logits`` with float multi-hot labels — and in ``kd_criterion``, which distils through
``sigmoid(teacher_logits)`` with alpha=0.5, T=1 (:8-19).  The auxiliary terms are the kernels of ``criterion.py``.
"""
from __future__ import annotations

from . import criterion as _c
from .criterion import bce_with_logits

__all__ = ["kd_criterion", "fitnet_criterion", "at_criterion", "gpw_criterion", "lpw_criterion", "nce_criterion"]


def kd_criterion(logits, labels, teacher_logits, alpha=0.5, T=1):
    """ppi_pyg/criterion.py:8-19."""
    loss_cls = bce_with_logits(logits, labels)
    loss_kd = _c._BCE.apply(logits, teacher_logits, True)
    return loss_kd * (alpha * T * T) + loss_cls * (1 - alpha), loss_cls, loss_kd


def fitnet_criterion(logits, labels, feat, teacher_feat, beta=1000):
    return _c.fitnet_criterion(logits, labels, feat, teacher_feat, beta, _cls=bce_with_logits)


def at_criterion(logits, labels, feat, teacher_feat, beta=1000):
    return _c.at_criterion(logits, labels, feat, teacher_feat, beta, _cls=bce_with_logits)


def gpw_criterion(logits, labels, feat, teacher_feat, kernel='cosine', beta=1, max_samples=8192, sampled_inds=None):
    return _c.gpw_criterion(logits, labels, feat, teacher_feat, kernel, beta, max_samples, sampled_inds, _cls=bce_with_logits)


def lpw_criterion(logits, labels, feat, teacher_feat, edge_index, kernel='cosine', beta=100, criterion='kld'):
    return _c.lpw_criterion(logits, labels, feat, teacher_feat, edge_index, kernel, beta, criterion, _cls=bce_with_logits)


def nce_criterion(logits, labels, feat, teacher_feat, beta=0.5, nce_T=0.075, max_samples=8192, sampled_inds=None):
    return _c.nce_criterion(logits, labels, feat, teacher_feat, beta, nce_T, max_samples, sampled_inds, _cls=bce_with_logits)
