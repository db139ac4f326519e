"""The drop-in mirrors the reference's interface: same names, positional order and defaults as the static
methods it replaces, and ``install()`` / ``uninstall()`` swap them on the reference's own classes (when the
reference tree is present, i.e. in the dev container; the signature table is checked everywhere)."""
import inspect
import os
import sys
import types

import pytest

import arvae_b200

REF = "/root/reference"

# (name, positional parameters, defaults) of the reference's functions -- utils/trainer.py:354-403,
# imagevae/mnist_vae.py:74
REFERENCE_SIGNATURES = {
    "compute_reg_loss": (["z", "labels", "reg_dim", "gamma", "factor"], {"factor": 1.0}),
    "reg_loss_sign": (["latent_code", "attribute", "factor"], {"factor": 1.0}),
    "compute_kld_loss": (["z_dist", "prior_dist", "beta", "c"], {"c": 0.0}),
}


@pytest.mark.parametrize("name", sorted(REFERENCE_SIGNATURES))
def test_signatures_match_the_reference_table(name):
    params, defaults = REFERENCE_SIGNATURES[name]
    sig = inspect.signature(getattr(arvae_b200, name))
    assert list(sig.parameters) == params
    for k, v in defaults.items():
        assert sig.parameters[k].default == v
    for k in params:
        if k not in defaults:
            assert sig.parameters[k].default is inspect.Parameter.empty


def _import_reference_trainer():
    for name in ("tensorboardX", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from utils.trainer import Trainer
    from imagevae.mnist_vae import MnistVAE
    return Trainer, MnistVAE


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the dev container")
def test_install_swaps_the_reference_static_methods_and_back():
    Trainer, MnistVAE = _import_reference_trainer()
    for name, (params, defaults) in REFERENCE_SIGNATURES.items():  # the table above is the reference's truth
        sig = inspect.signature(getattr(Trainer, name))
        assert list(sig.parameters) == params, name
        for k, v in defaults.items():
            assert sig.parameters[k].default == v
    orig = {n: Trainer.__dict__[n] for n in REFERENCE_SIGNATURES}
    orig_rep = MnistVAE.__dict__["reparametrize"]
    arvae_b200.install(Trainer, vae_classes=[MnistVAE])
    try:
        assert Trainer.compute_reg_loss is arvae_b200.compute_reg_loss
        assert Trainer.reg_loss_sign is arvae_b200.reg_loss_sign
        assert Trainer.compute_kld_loss is arvae_b200.compute_kld_loss
        assert isinstance(Trainer.__dict__["compute_reg_loss"], staticmethod)  # callers use self.compute_reg_loss(...)
        assert MnistVAE.__dict__["reparametrize"] is not orig_rep
    finally:
        arvae_b200.uninstall()
    for n, v in orig.items():
        assert Trainer.__dict__[n] is v
    assert MnistVAE.__dict__["reparametrize"] is orig_rep
