#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference code.

Runs only in the dev container (needs /root/reference, which does not travel
to the GPU box).  The reference's ``utils/trainer.py`` imports two packages
that are not installed here (tensorboardX, matplotlib.pyplot, lines 10-11);
empty stub modules are placed in ``sys.modules`` so the file imports and its
static methods run exactly as written.  Every fixture stores its inputs as
well as the reference's outputs, so the tests never depend on RNG
reproducibility across machines.

    python tests/golden/make_golden.py            # rewrites every fixture

Reference entry points executed:
  utils.trainer.Trainer.compute_reg_loss / reg_loss_sign / compute_kld_loss
  imagevae.mnist_vae.MnistVAE.reparametrize (through an instance built without
  running the conv constructor's weights through anything)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference not present; golden vectors can only be made in the dev container")
    for name in ("tensorboardX", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from utils.trainer import Trainer  # noqa
    from imagevae.mnist_vae import MnistVAE  # noqa
    return Trainer, MnistVAE


def ref_dim_loop(Trainer, z, labels, reg_dims, gamma, delta):
    """imagevae/image_vae_trainer.py:171-180 verbatim in behaviour."""
    reg_loss = 0.0
    for dim in reg_dims:
        reg_loss += Trainer.compute_reg_loss(z, labels[:, dim], dim, gamma=gamma, factor=delta)
    return reg_loss


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.0f} KiB)")


def reg_case(Trainer, name, z, labels, reg_dims, gamma, delta, per_dim=True):
    z = z.clone().requires_grad_(True)
    loss = ref_dim_loop(Trainer, z, labels, reg_dims, gamma, delta)
    loss.backward()
    extra = {}
    if per_dim:
        per = []
        for dim in reg_dims:
            per.append(Trainer.compute_reg_loss(z.detach(), labels[:, dim], dim, gamma=gamma,
                                                factor=delta).item())
        extra["per_dim_loss"] = np.asarray(per, dtype=np.float32)
    save(name, z=z, labels=labels, reg_dims=np.asarray(reg_dims, dtype=np.int64),
         gamma=np.float64(gamma), delta=np.float64(delta), loss=loss, grad_z=z.grad, **extra)


def music_attribute_fixture():
    """Runs BarDataset.get_* of the unmodified reference on synthetic bars.  music21 is not installed: it is
    stubbed in sys.modules, with ``music21.pitch.Pitch(name).midi`` implemented by the standard pitch-name rule
    (arvae_b200.music.midi_from_pitch_name) -- the only music21 behaviour these four methods use."""
    from arvae_b200 import music, synth

    class Pitch:
        def __init__(self, name):
            self.midi = music.midi_from_pitch_name(name)

    m21 = types.ModuleType("music21")
    for sub in ("meter", "abcFormat", "note", "pitch", "interval", "stream", "corpus", "converter", "expressions",
                "duration", "tempo", "key", "metadata", "instrument"):
        mod = types.ModuleType("music21." + sub)
        setattr(m21, sub, mod)
        sys.modules["music21." + sub] = mod
    m21.abcFormat.ABCHandlerException = Exception
    m21.pitch.Pitch = Pitch
    sys.modules["music21"] = m21
    from data.dataloaders.bar_dataset import BarDataset  # the reference class, unmodified

    note2index, index2note = synth.music_vocabulary()
    fake_self = types.SimpleNamespace(note2index_dicts=note2index, index2note_dicts=index2note)
    measures = synth.make_measures(257, seed=2024)
    cols = [BarDataset.get_rhy_complexity(fake_self, measures), BarDataset.get_pitch_range_in_measure(fake_self, measures),
            BarDataset.get_note_density_in_measure(fake_self, measures), BarDataset.get_contour(fake_self, measures)]
    attrs = torch.stack([c.float().cpu() for c in cols], dim=1)  # MUSIC_REG_TYPE order
    save("music_attrs", measures=measures, attrs=attrs)


def main():
    torch.set_num_threads(8)
    Trainer, MnistVAE = import_reference()
    from arvae_b200 import synth

    # ---- the BASELINE.json config shapes that the reference can allocate -------------
    for name, B in (("c1_mnist_b64", None), ("c2_dsprites_b4096", 512), ("c2_dsprites_b4096", None),
                    ("c3_measure_b2048", None), ("c4_mnist_b65536", 1000)):
        c = synth.make_case(name, B)
        tag = name if B is None else f"{name.rsplit('_b', 1)[0]}_b{B}"
        reg_case(Trainer, "reg_" + tag, c["z"], c["labels"], c["reg_dims"], c["gamma"], c["delta"])

    # ---- single-call semantics of compute_reg_loss (one dim, strided label view, negative dim) ----
    c = synth.make_case("c1_mnist_b64")
    z = c["z"].clone().requires_grad_(True)
    l = Trainer.compute_reg_loss(z, c["labels"][:, 3], -2, gamma=2.5, factor=3.0)
    l.backward()
    save("reg_single_negdim", z=z, labels=c["labels"], label_col=np.int64(3), reg_dim=np.int64(-2),
         gamma=np.float64(2.5), delta=np.float64(3.0), loss=l, grad_z=z.grad)

    # ---- reg_loss_sign directly + the sign matrix it builds (trainer.py:394-400) ----
    g = torch.Generator().manual_seed(77)
    B = 37
    x = torch.randn(B, generator=g)
    a = torch.randint(0, 5, (B,), generator=g).float()
    a[3] = float("nan"); a[5] = float("inf"); a[6] = float("inf"); a[7] = float("-inf")
    a[8] = 0.0; a[9] = -0.0; a[10] = 1e-45; a[11] = 2e-45; a[12] = -1e-45
    A = a.view(-1, 1).repeat(1, B)
    sign_mat = torch.sign((A - A.transpose(1, 0))).to(torch.int8)
    xr = x.clone().requires_grad_(True)
    l = Trainer.reg_loss_sign(xr, a, factor=1.7)
    l.backward()
    save("sign_matrix_special", x=x, a=a, sign=sign_mat, delta=np.float64(1.7), loss=l, grad_x=xr.grad)

    # ---- edge cases ---------------------------------------------------------------
    def edge(name, x, a, delta):
        xr = x.clone().requires_grad_(True)
        l = Trainer.reg_loss_sign(xr, a, factor=delta)
        if l.requires_grad and x.numel() > 0:
            l.backward()
            gx = xr.grad
        else:
            gx = torch.zeros_like(x)
        save("edge_" + name, x=x, a=a, delta=np.float64(delta), loss=l, grad_x=gx)

    g = torch.Generator().manual_seed(99)
    edge("b1", torch.tensor([0.3]), torch.tensor([1.0]), 1.0)
    edge("b2", torch.tensor([0.3, -0.2]), torch.tensor([1.0, 2.0]), 1.0)
    edge("b3_ties", torch.tensor([0.3, -0.2, 0.9]), torch.tensor([1.0, 1.0, 1.0]), 2.0)
    for B in (127, 128, 129, 1000):
        edge(f"rand_b{B}", torch.randn(B, generator=g), torch.randn(B, generator=g), 1.0)
    B = 200
    xs = torch.sort(torch.randn(B, generator=g))[0]
    edge("sorted_agree", xs, torch.arange(B).float(), 10.0)          # loss -> ~0
    edge("sorted_oppose", xs, -torch.arange(B).float(), 10.0)        # loss -> ~2
    edge("all_ties", torch.randn(B, generator=g), torch.zeros(B), 1.0)
    edge("all_equal_z", torch.zeros(B), torch.randn(B, generator=g), 1.0)
    edge("all_equal_both", torch.full((B,), 0.25), torch.full((B,), 3.0), 5.0)
    edge("saturated", 40.0 * torch.randn(B, generator=g), torch.randn(B, generator=g), 10.0)
    edge("tiny_factor", torch.randn(B, generator=g), torch.randn(B, generator=g), 1e-3)
    edge("negative_factor", torch.randn(B, generator=g), torch.randn(B, generator=g), -1.5)
    edge("few_values", torch.randn(B, generator=g), torch.randint(0, 3, (B,), generator=g).float(), 1.0)
    dup = torch.randn(B // 2, generator=g)
    edge("dup_z_ties", torch.cat([dup, dup]), torch.randint(0, 2, (B,), generator=g).float(), 1.0)
    an = torch.randn(B, generator=g)
    an[::7] = float("nan"); an[1::11] = float("inf"); an[2::13] = float("-inf")
    edge("nan_inf_labels", torch.randn(B, generator=g), an, 1.0)
    sub = torch.randint(-3, 4, (B,), generator=g).float() * 1.4e-45
    edge("subnormal_labels", torch.randn(B, generator=g), sub, 1.0)

    # int64 / float64 labels: sign is taken in the label dtype first (App. A.3)
    xi = torch.randn(150, generator=g)
    for tag, lab in (("int64", torch.randint(-5, 6, (150,), generator=g)),
                     ("float64", torch.randn(150, generator=g, dtype=torch.float64))):
        xr = xi.clone().requires_grad_(True)
        l = Trainer.reg_loss_sign(xr, lab, factor=1.0)
        l.backward()
        save("edge_labels_" + tag, x=xi, a=lab, delta=np.float64(1.0), loss=l, grad_x=xr.grad)

    # ---- fused latent head (C3 shape): reparametrize + KLD + reg ---------------------
    c = synth.make_case("c3_measure_b2048")
    loc, log_std, _ = synth.make_latent_head(c["B"], c["Z"], c["seed"] + 100)
    beta, cap = 0.001, 0.0
    loc = loc.requires_grad_(True)
    log_std = log_std.requires_grad_(True)
    scale = torch.exp(log_std)                                   # measurevae/encoder.py:120-123
    scale.retain_grad()
    z_dist = torch.distributions.Normal(loc=loc, scale=scale)
    vae = MnistVAE.__new__(MnistVAE)                             # reparametrize uses no module state
    torch.manual_seed(4242)
    z_tilde, z_prior, prior = MnistVAE.reparametrize(vae, z_dist)
    torch.manual_seed(4242)
    eps = torch.distributions.utils._standard_normal(loc.shape, dtype=loc.dtype, device=loc.device)
    assert torch.equal((loc + eps * scale).detach(), z_tilde.detach())
    z_prior_again = torch.normal(torch.zeros_like(loc), torch.ones_like(loc))
    assert torch.equal(z_prior_again, z_prior)                   # RNG order: eps first, then z_prior
    z_tilde.retain_grad()
    kld = Trainer.compute_kld_loss(z_dist, prior, beta=beta, c=cap)
    reg = ref_dim_loop(Trainer, z_tilde, c["labels"], c["reg_dims"], c["gamma"], c["delta"])
    total = kld + reg
    total.backward()
    save("head_c3_measure_b2048", loc=loc, log_std=log_std, eps=eps, labels=c["labels"],
         reg_dims=np.asarray(c["reg_dims"], dtype=np.int64), beta=np.float64(beta), capacity=np.float64(cap),
         gamma=np.float64(c["gamma"]), delta=np.float64(c["delta"]), z_tilde=z_tilde, z_prior=z_prior,
         seed=np.int64(4242), kld_loss=kld, reg_loss=reg,
         grad_loc=loc.grad, grad_log_std=log_std.grad, grad_scale=scale.grad, grad_z=z_tilde.grad)

    # ---- KLD alone with the image trainer's 1-element capacity tensor -----------------
    c = synth.make_case("c1_mnist_b64")
    loc, log_std, _ = synth.make_latent_head(c["B"], c["Z"], 555)
    loc = loc.requires_grad_(True)
    scale = torch.exp(log_std).requires_grad_(True)
    z_dist = torch.distributions.Normal(loc=loc, scale=scale)
    prior = torch.distributions.Normal(torch.zeros_like(loc), torch.ones_like(scale))
    cap_t = torch.FloatTensor([1.5])                              # image_vae_trainer.py:94 passes a [1] tensor
    k = Trainer.compute_kld_loss(z_dist, prior, beta=4.0, c=cap_t)
    k.sum().backward()
    save("kld_c1_capacity_tensor", loc=loc, scale=scale, beta=np.float64(4.0), capacity=cap_t,
         kld_loss=k, grad_loc=loc.grad, grad_scale=scale.grad)

    # ---- musical attribute extractors (data/dataloaders/bar_dataset.py:338-500), reference method bodies ----
    music_attribute_fixture()

    with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as f:
        f.write("Golden vectors produced by tests/golden/make_golden.py from the unmodified reference\n"
                f"reference path: {REF} (ashispati/ar-vae, utils/trainer.py:354-403, imagevae/mnist_vae.py:74-87)\n"
                f"torch {torch.__version__}, numpy {np.__version__}, CPU, {torch.get_num_threads()} threads\n")


if __name__ == "__main__":
    main()
