"""Seeded synthetic latents and attribute labels shaped like the reference's datasets.

The reference's data (Morpho-MNIST, dSprites, Bach/Folk bars) is not available
offline, so benches and tests use stand-ins that mimic the VALUE SETS of the
real label columns -- ties and quantisation are what stress the sign matrix:

* Morpho-MNIST [B,7]: digit id + six continuous morphometrics
  (value ranges from imagevae/image_vae_trainer.py:30-38, rounded to 3 decimals).
* dSprites [B,6]: color, shape(3), scale(6), orientation(40), posX(32), posY(32)
  value grids (data/dataloaders/dsprites_dataset.py:38-53) -- massive ties.
* Music bars [B,4]: rhythmic complexity, pitch range, note density, contour
  (data/dataloaders/bar_dataset.py:338-500) -- heavily quantised.

Everything is generated with a CPU ``torch.Generator`` so that host and device
see identical bits.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

# (B, Z, reg_dims, label maker, gamma, delta) per BASELINE.json config
CONFIGS: Dict[str, dict] = {
    # C1: MnistVAE train-step shape (imagevae/mnist_vae.py z_dim 16; train_image_vae.py gamma 10, delta 1)
    "c1_mnist_b64": dict(B=64, Z=16, reg_dims=(1, 2, 3, 4), labels="morpho", gamma=10.0, delta=1.0, seed=1235),
    # C2: DspritesVAE z_dim 10, attrs shape/scale/orient/x/y
    "c2_dsprites_b4096": dict(B=4096, Z=10, reg_dims=(1, 2, 3, 4, 5), labels="dsprites", gamma=10.0, delta=1.0, seed=1236),
    # C3: MeasureVAE z_dim 32, 4 musical attrs, gamma 1 delta 10 (train_measure_vae.py:46-49)
    "c3_measure_b2048": dict(B=2048, Z=32, reg_dims=(0, 1, 2, 3), labels="music", gamma=1.0, delta=10.0, seed=1237),
    # C4: large batch, MNIST-shaped labels, six regularised attrs
    "c4_mnist_b65536": dict(B=65536, Z=16, reg_dims=(1, 2, 3, 4, 5, 6), labels="morpho", gamma=10.0, delta=1.0, seed=1238),
}


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def morpho_labels(B: int, g: torch.Generator) -> torch.Tensor:
    """[B,7] float32: digit, area, length, thickness, slant, width, height."""
    cols = [torch.randint(0, 10, (B,), generator=g).float()]
    for lo, hi in ((0.0, 350.0), (0.0, 100.0), (0.0, 15.0), (-1.2, 1.2), (0.0, 30.0), (0.0, 30.0)):
        u = torch.rand(B, generator=g, dtype=torch.float64) * (hi - lo) + lo
        cols.append((torch.round(u * 1000.0) / 1000.0).float())
    return torch.stack(cols, dim=1).contiguous()


def dsprites_labels(B: int, g: torch.Generator) -> torch.Tensor:
    """[B,6] float32 on the dSprites latent value grids."""
    def grid(lo, hi, n):
        vals = torch.linspace(lo, hi, n, dtype=torch.float64)
        return vals[torch.randint(0, n, (B,), generator=g)].float()
    cols = [torch.ones(B),
            torch.randint(1, 4, (B,), generator=g).float(),
            grid(0.5, 1.0, 6),
            grid(0.0, 2.0 * math.pi, 40),
            grid(0.0, 1.0, 32),
            grid(0.0, 1.0, 32)]
    return torch.stack(cols, dim=1).contiguous()


def music_labels(B: int, g: torch.Generator) -> torch.Tensor:
    """[B,4] float32: rhythmic complexity, pitch range, note density, contour."""
    # metrical weights in the spirit of bar_dataset_helpers.py:21-30 (24 ticks per bar)
    w = torch.tensor([5, 1, 1, 2, 1, 1, 3, 1, 1, 2, 1, 1, 4, 1, 1, 2, 1, 1, 3, 1, 1, 2, 1, 1],
                     dtype=torch.float64)
    onsets = (torch.rand(B, 24, generator=g, dtype=torch.float64) < 0.4).double()
    rhy = (onsets * w).sum(1) / w.sum()
    pr = torch.randint(0, 30, (B,), generator=g).double() / 26.0
    nd = torch.randint(0, 25, (B,), generator=g).double() / 24.0
    ct = torch.randint(-29, 30, (B,), generator=g).double() / 26.0
    return torch.stack([rhy, pr, nd, ct], dim=1).float().contiguous()


_LABELS = {"morpho": morpho_labels, "dsprites": dsprites_labels, "music": music_labels}


def make_labels(kind: str, B: int, seed: int) -> torch.Tensor:
    return _LABELS[kind](B, _gen(seed))


def make_case(name: str, B: int | None = None) -> dict:
    """Inputs for one named config (optionally at another batch size)."""
    cfg = dict(CONFIGS[name])
    if B is not None:
        cfg["B"] = int(B)
    g = _gen(cfg["seed"])
    cfg["z"] = torch.randn(cfg["B"], cfg["Z"], generator=g)
    cfg["labels"] = _LABELS[cfg["labels"]](cfg["B"], g)
    return cfg


def make_latent_head(B: int, Z: int, seed: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(loc, log_std, eps) for the fused reparametrize + KLD configs (C3)."""
    g = _gen(seed)
    loc = torch.randn(B, Z, generator=g)
    log_std = -1.0 + 0.25 * torch.randn(B, Z, generator=g)
    eps = torch.randn(B, Z, generator=g)
    return loc, log_std, eps


def music_vocabulary():
    """A note dictionary shaped like the Bach / folk bar datasets': five special symbols and pitch names."""
    names = []
    for octave in (3, 4, 5):
        for step in ("C", "C#", "D", "E-", "E", "F", "F#", "G", "A-", "A", "B-", "B"):
            names.append(f"{step}{octave}")
    symbols = ["__", "START", "END", "rest", None] + names
    note2index = {s: i for i, s in enumerate(symbols)}
    index2note = {i: s for s, i in note2index.items()}
    return note2index, index2note


def make_measures(B: int, seed: int, T: int = 24) -> torch.Tensor:
    """[B, T] int64 bars: note onsets followed by slurs, some rests, a few START / END / None paddings,
    plus empty and single-note bars (the reference's special cases)."""
    g = _gen(seed)
    note2index, _ = music_vocabulary()
    n_special = 5
    V = len(note2index)
    m = torch.full((B, T), note2index["__"], dtype=torch.int64)
    onset = torch.rand(B, T, generator=g) < 0.35
    pitch = torch.randint(n_special, V, (B, T), generator=g)
    m[onset] = pitch[onset]
    rest = (torch.rand(B, T, generator=g) < 0.08) & ~onset
    m[rest] = note2index["rest"]
    pad = torch.rand(B, generator=g) < 0.1
    m[pad, 0] = note2index["START"]
    m[pad, -1] = note2index["END"]
    nonebar = torch.rand(B, generator=g) < 0.05
    m[nonebar, 3] = note2index[None]
    if B > 3:
        m[0, :] = note2index["__"]                       # no notes at all
        m[1, :] = note2index["__"]; m[1, 5] = n_special + 7   # exactly one note
        m[2, :] = note2index["rest"]
    return m
