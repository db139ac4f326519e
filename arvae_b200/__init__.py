"""arvae_b200 -- AR-VAE's attribute-regularization hot path, hand-written CUDA for NVIDIA B200.

Drop-in replacements for the reference's ``Trainer.compute_reg_loss`` / ``reg_loss_sign`` /
``compute_kld_loss`` and ``MnistVAE.reparametrize``, plus fused multi-dim and whole-head
forms, on top of the C ABI in ``include/arvae_b200.h`` (``csrc/libarvae_b200.so``).
"""
from __future__ import annotations

from . import _lib  # noqa: F401
from . import graphs  # noqa: F401
from . import music  # noqa: F401
from . import evaluation  # noqa: F401
from .ops import (ALGO_AUTO, ALGO_DENSE, ALGO_SORTED, ALGO_TRIANGLE, compute_kld_loss, compute_reg_loss, latent_head, latent_loss_head, mufu_per_pair,
                  reg_loss_fused, reg_loss_rows, reg_loss_sign, reparam_kld_reg, reparametrize, sign_matrix,
                  attr_argsort, pack_columns)

__version__ = "0.1.0"

__all__ = ["compute_reg_loss", "reg_loss_sign", "reg_loss_fused", "reg_loss_rows", "compute_kld_loss",
           "reparametrize", "latent_head", "latent_loss_head", "reparam_kld_reg", "sign_matrix", "install", "uninstall",
           "ALGO_AUTO", "ALGO_DENSE", "ALGO_SORTED", "ALGO_TRIANGLE", "mufu_per_pair", "attr_argsort", "pack_columns", "graphs", "music", "evaluation"]

_saved = {}


def _measure_vae_forward(self, measure_score_tensor, measure_metadata_tensor, train=True):
    """``MeasureVAE.forward`` (measurevae/measure_vae.py:97-131) with its inline ``rsample`` / prior / ``sample``
    block (lines 115-123) replaced by the fused :func:`reparametrize` -- same RNG draw order, same six results."""
    seq_len = measure_score_tensor.size(1)
    assert seq_len == self.num_ticks_per_measure
    z_dist = self.encoder(measure_score_tensor)
    z_tilde, z_prior, prior_dist = reparametrize(z_dist)
    weights, samples = self.decoder(z=z_tilde, score_tensor=measure_score_tensor, train=train)
    return weights, samples, z_dist, prior_dist, z_tilde, z_prior


def install(trainer_cls, vae_classes=(), wrap=None):
    """Swap the hot path into the reference's classes with zero edits to its trainers.

    ``trainer_cls`` is the reference's ``utils.trainer.Trainer``: its static methods
    ``compute_reg_loss`` / ``reg_loss_sign`` / ``compute_kld_loss`` (utils/trainer.py:354-403) are
    replaced, so ``ImageVAETrainer`` / ``MeasureVAETrainer.loss_and_acc_for_batch`` call the CUDA
    path unchanged.  Each class in ``vae_classes`` gets the fused reparametrize: a class with a
    ``reparametrize`` method (``MnistVAE``, ``DspritesVAE``; imagevae/mnist_vae.py:74-87) has that method
    replaced; a class without one (``MeasureVAE``, which samples inline in ``forward``,
    measurevae/measure_vae.py:115-123) has ``forward`` replaced by an equivalent that calls it.

    ``wrap(name, function)`` (tests, tracing): when given, what it returns is installed in slot ``name``
    (``"compute_reg_loss"``, ``"reg_loss_sign"``, ``"compute_kld_loss"``, ``"reparametrize"``, ``"forward"``)
    instead of ``function``; returning None keeps ``function``.
    """
    def pick(name, fn):
        other = wrap(name, fn) if wrap is not None else None
        return fn if other is None else other

    _saved[trainer_cls] = {n: trainer_cls.__dict__.get(n) for n in
                           ("compute_reg_loss", "reg_loss_sign", "compute_kld_loss")}
    trainer_cls.compute_reg_loss = staticmethod(pick("compute_reg_loss", compute_reg_loss))
    trainer_cls.reg_loss_sign = staticmethod(pick("reg_loss_sign", reg_loss_sign))
    trainer_cls.compute_kld_loss = staticmethod(pick("compute_kld_loss", compute_kld_loss))
    for cls in vae_classes:
        if hasattr(cls, "reparametrize"):
            _saved[cls] = {"reparametrize": cls.__dict__.get("reparametrize")}
            cls.reparametrize = pick("reparametrize", lambda self, z_dist: reparametrize(z_dist))
        else:
            _saved[cls] = {"forward": cls.__dict__.get("forward")}
            cls.forward = pick("forward", _measure_vae_forward)


def uninstall():
    """Undo :func:`install`."""
    for cls, attrs in _saved.items():
        for name, val in attrs.items():
            if val is None:
                if name in cls.__dict__:
                    delattr(cls, name)
            else:
                setattr(cls, name, val)
    _saved.clear()
