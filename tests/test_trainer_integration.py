"""The REAL reference trainer driven through ``arvae_b200.install()`` (dev container only: needs /root/reference).

``ImageVAETrainer.loss_and_acc_for_batch`` (imagevae/image_vae_trainer.py:137-217) is imported unmodified (optional
plotting / music dependencies stubbed), given a real ``MnistVAE`` and one C1-shaped batch, and run twice on the CPU:
stock, and after ``install(..., wrap=recorder)``.  The recorder sits exactly where the CUDA functions are installed --
same names, same staticmethod / method slots -- logs every call's arguments and forwards to the reference's own
function, so the test proves

* the trainer reaches all three patch points with the arguments the drop-in is written for: ``labels[:, dim]`` as a
  stride-A view, the same ``dim`` as latent index, ``gamma`` / ``factor`` as keywords, a ``[1]``-shaped capacity
  tensor, ``reparametrize`` called once per forward with the encoder's ``Normal``;
* a whole train step (forward, backward, parameter gradients) is unchanged by the swap;
* the writer quirk (``self.writer`` is None but written to when the epoch number changes) is
  handled by keeping ``epoch_num`` constant, as a harness must.

The numerics of the CUDA functions themselves are covered by tests/test_gpu_parity.py and test_gpu_integration.py
(the GPU box has no /root/reference)."""
import os
import sys
import types

import pytest
import torch

import arvae_b200

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the dev container")


class _Stub(types.ModuleType):
    """Module whose every attribute is another stub (optional deps the trainer imports but this path never calls)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Stub(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return None


def _import_reference():
    for name in ("tensorboardX", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors",
                 "matplotlib.animation", "seaborn", "pypianoroll", "pretty_midi", "skimage", "skimage.morphology",
                 "skimage.measure", "skimage.filters", "skimage.transform", "skimage.draw", "music21"):
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    if not hasattr(sys.modules["tensorboardX"], "SummaryWriter"):
        sys.modules["tensorboardX"].SummaryWriter = object
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from imagevae.image_vae_trainer import ImageVAETrainer
    from imagevae.mnist_vae import MnistVAE
    from utils.trainer import Trainer
    return ImageVAETrainer, MnistVAE, Trainer


class MorphoMnistDataset:  # the trainer dispatches on the dataset's class NAME (image_vae_trainer.py:81-86)
    pass


def _make_trainer(ImageVAETrainer, MnistVAE, reg_dim):
    torch.manual_seed(0)
    model = MnistVAE()
    model.eval()  # no dropout noise: both runs see the same network function
    trainer = ImageVAETrainer(MorphoMnistDataset(), model, reg_type=("thickness", "slant", "width", "height"),
                              reg_dim=reg_dim, beta=4.0, gamma=10.0, delta=1.0, capacity=0.0, rand=0)
    trainer.writer = None
    return trainer, model


def _one_step(trainer, model, inputs, labels, seed):
    for p in model.parameters():
        p.grad = None
    torch.manual_seed(seed)  # same rsample noise in both runs
    loss, acc = trainer.loss_and_acc_for_batch((inputs, labels), epoch_num=trainer.cur_epoch_num, batch_num=1, train=True)
    loss.mean().backward()  # utils/trainer.py:140 (the image trainer's loss is [1]-shaped)
    return loss.detach().clone(), [p.grad.detach().clone() for p in model.parameters()]


def _recorder(Trainer, MnistVAE, calls):
    """install(wrap=...) hook: what install() would put in slot `name` is `ours`; the slot gets a function that records
    the call and runs the reference's own function (the CUDA one cannot run in the dev container)."""
    originals = {n: getattr(Trainer, n) for n in ("compute_reg_loss", "reg_loss_sign", "compute_kld_loss")}
    orig_rep = MnistVAE.__dict__["reparametrize"]

    def wrap(name, ours):
        def rec_static(*args, **kwargs):
            calls.append((name, args, kwargs))
            return originals[name](*args, **kwargs)

        def rec_reparametrize(self, z_dist):
            calls.append((name, (z_dist,), {}))
            return orig_rep(self, z_dist)

        assert callable(ours)
        return rec_reparametrize if name == "reparametrize" else rec_static

    return wrap


def test_real_trainer_step_through_install_with_recording_backend():
    ImageVAETrainer, MnistVAE, Trainer = _import_reference()
    reg_dim = (1, 2, 3, 4)
    B, A = 64, 7
    g = torch.Generator().manual_seed(1234)
    inputs = torch.rand(B, 1, 28, 28, generator=g)
    labels = torch.rand(B, A, generator=g) * 10.0
    trainer, model = _make_trainer(ImageVAETrainer, MnistVAE, reg_dim)
    ref_loss, ref_grads = _one_step(trainer, model, inputs, labels, seed=7)

    calls = []
    originals = {n: getattr(Trainer, n) for n in ("compute_reg_loss", "reg_loss_sign", "compute_kld_loss")}
    orig_rep = MnistVAE.__dict__["reparametrize"]
    recorder = _recorder(Trainer, MnistVAE, calls)
    arvae_b200.install(Trainer, vae_classes=[MnistVAE], wrap=recorder)
    try:
        assert isinstance(Trainer.__dict__["compute_reg_loss"], staticmethod)
        new_loss, new_grads = _one_step(trainer, model, inputs, labels, seed=7)
    finally:
        arvae_b200.uninstall()
    assert Trainer.__dict__["compute_reg_loss"].__func__ is originals["compute_reg_loss"]
    assert MnistVAE.__dict__["reparametrize"] is orig_rep

    # the whole step is unchanged
    assert torch.equal(new_loss, ref_loss)
    assert new_loss.shape == (1,)  # [1]-shaped capacity makes the image trainer's loss [1]-shaped
    for a, b in zip(new_grads, ref_grads):
        assert torch.equal(a, b)

    # the call pattern the drop-in is written for
    # (the reference's compute_reg_loss calls Trainer.reg_loss_sign itself, utils/trainer.py:375: those nested calls
    # hit the recorder too; the CUDA compute_reg_loss is one fused call and never goes through that slot)
    assert sum(c[0] == "reg_loss_sign" for c in calls) == len(reg_dim)
    calls = [c for c in calls if c[0] != "reg_loss_sign"]
    names = [c[0] for c in calls]
    assert names == ["reparametrize", "compute_kld_loss"] + ["compute_reg_loss"] * len(reg_dim)
    z_dist = calls[0][1][0]
    assert isinstance(z_dist, torch.distributions.Normal) and z_dist.loc.shape == (B, 16)
    _, kargs, kkw = calls[1]
    assert kargs[0] is z_dist and set(kkw) == {"beta", "c"}
    assert kkw["beta"] == 4.0 and isinstance(kkw["c"], torch.Tensor) and kkw["c"].shape == (1,)
    for (_, args, kw), dim in zip(calls[2:], reg_dim):
        z_tilde, lab, d = args
        assert d == dim and set(kw) == {"gamma", "factor"} and kw["gamma"] == 10.0 and kw["factor"] == 1.0
        assert z_tilde.shape == (B, 16) and z_tilde.requires_grad
        assert lab.shape == (B,) and lab.stride() == (A,)           # a strided view, not a copy
        assert lab.data_ptr() == labels[:, dim].data_ptr()


def test_non_tuple_reg_dim_still_raises_type_error():
    """image_vae_trainer.py:178-179: the trainer's own check, untouched by the swap."""
    ImageVAETrainer, MnistVAE, Trainer = _import_reference()
    trainer, model = _make_trainer(ImageVAETrainer, MnistVAE, [1, 2])
    arvae_b200.install(Trainer, vae_classes=[MnistVAE], wrap=_recorder(Trainer, MnistVAE, []))
    try:
        with pytest.raises(TypeError):
            trainer.loss_and_acc_for_batch((torch.rand(4, 1, 28, 28), torch.rand(4, 7)), epoch_num=0, batch_num=1)
    finally:
        arvae_b200.uninstall()
