#!/usr/bin/env python
"""Generate tests/golden/eval_*.npz by running the UNMODIFIED reference utils/evaluation.py.

Dev container only (needs /root/reference).  Executes compute_correlation_score, _compute_correlation_matrix,
compute_sap_score, _compute_score_matrix (utils/evaluation.py:146-219) and scipy.stats.spearmanr exactly as the
reference calls it (:166), on seeded inputs shaped like compute_representations' output
(imagevae/image_vae_trainer.py:316-331: float32 latent codes [N, Z], float32 attributes [N, A]).

    python tests/golden/make_golden_eval.py
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference not present; golden vectors can only be made in the dev container")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import utils.evaluation as ev  # noqa
    return ev


def codes_for(attrs: np.ndarray, Z: int, reg_dims, seed: int, noise: float, weak=()):
    """Latent codes the way a trained AR-VAE leaves them: dim reg_dims[k] monotone in attribute k plus noise,
    the other dims independent; `weak` = (dim, attr, strength) adds faint dependence (p-values near the gate)."""
    rng = np.random.default_rng(seed)
    n = attrs.shape[0]
    z = rng.standard_normal((n, Z))
    for k, d in enumerate(reg_dims):
        a = attrs[:, k].astype(np.float64)
        s = a.std()
        a = (a - a.mean()) / (s if s > 0 else 1.0)
        z[:, d] = np.tanh(a) * 1.5 + noise * rng.standard_normal(n)
    for d, k, w in weak:
        a = attrs[:, k].astype(np.float64)
        z[:, d] += w * (a - a.mean()) / a.std()
    return z.astype(np.float32)


def run(ev, name, mus, ys):
    Z, A = mus.shape[1], ys.shape[1]
    from scipy.stats import spearmanr
    rho = np.zeros((Z, A))
    p = np.zeros((Z, A))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            for i in range(Z):
                for j in range(A):
                    rho[i, j], p[i, j] = spearmanr(mus[:, i], ys[:, j])     # evaluation.py:166
            corr_matrix = ev._compute_correlation_matrix(mus, ys)
            corr = ev.compute_correlation_score(mus, ys)["Corr_score"]
            sap_matrix = ev._compute_score_matrix(mus, ys)
            if Z >= 2:
                sap = ev.compute_sap_score(mus, ys)["SAP_score"]
            else:
                sap = np.nan
    np.savez_compressed(os.path.join(HERE, name + ".npz"), mus=mus, ys=ys, rho=rho, p=p, corr_matrix=corr_matrix,
                        corr_score=np.float64(corr), sap_matrix=sap_matrix, sap_score=np.float64(sap))
    print(f"{name}: N={mus.shape[0]} Z={Z} A={A} Corr_score={corr:.6f} SAP_score={sap:.6f} "
          f"gate-open={int((p <= 0.05).sum())}/{Z * A}")


def main():
    ev = import_reference()
    from arvae_b200 import synth

    # MNIST-shaped: 16 codes, six morphometric attributes on dims 1..6 (image_vae_trainer.py:30-38)
    lab = synth.make_labels("morpho", 3000, 11).numpy()[:, 1:]
    run(ev, "eval_mnist_n3000", codes_for(lab, 16, (1, 2, 3, 4, 5, 6), 21, 0.4, weak=((0, 2, 0.035), (9, 4, 0.04))), lab)

    # dSprites-shaped: 10 codes, shape/scale/orientation/x/y -- massive ties in every attribute
    lab = synth.make_labels("dsprites", 4096, 12).numpy()[:, 1:]
    run(ev, "eval_dsprites_n4096", codes_for(lab, 10, (1, 2, 3, 4, 5), 22, 0.7, weak=((7, 1, 0.03),)), lab)

    # ... including the constant 'color' column: spearmanr -> NaN (gate closed), SAP var_y = 0 -> NaN
    lab = synth.make_labels("dsprites", 1024, 13).numpy()
    run(ev, "eval_dsprites_const_attr", codes_for(lab[:, 1:], 10, (1, 2, 3, 4, 5), 23, 0.7), lab)

    # music-shaped: 32 codes, 4 quantised attributes on dims 0..3
    lab = synth.make_labels("music", 2048, 14).numpy()
    run(ev, "eval_music_n2048", codes_for(lab, 32, (0, 1, 2, 3), 24, 0.5, weak=((17, 0, 0.045),)), lab)

    # collapsed and duplicated code dims: var_mu <= 1e-12 gate, constant-input spearmanr, ties among codes
    lab = synth.make_labels("morpho", 500, 15).numpy()[:, 1:4]
    z = codes_for(lab, 6, (0, 1, 2), 25, 0.3)
    z[:, 3] = 0.25                                   # constant
    z[:, 4] = 1e-7 * np.sign(z[:, 0])                # variance below the gate, but rank-correlated
    z[:, 5] = np.round(z[:, 1], 1)                   # heavy ties in a code
    run(ev, "eval_collapsed_codes", z, lab)

    # tiny and odd sizes
    rng = np.random.default_rng(5)
    run(ev, "eval_n3", rng.standard_normal((3, 2)).astype(np.float32), rng.standard_normal((3, 2)).astype(np.float32))
    run(ev, "eval_n257_perfect", np.stack([np.arange(257.), -np.arange(257.) ** 3, rng.standard_normal(257)], 1).astype(np.float32),
        np.stack([np.arange(257.) * 2 + 1, rng.standard_normal(257)], 1).astype(np.float32))

    # non-finite values: NaN propagates through spearmanr and np.cov, infinities rank normally; -0.0 ties with 0.0
    z = rng.standard_normal((400, 5)).astype(np.float32)
    y = rng.integers(0, 4, (400, 3)).astype(np.float32)
    z[:, 1] = np.where(rng.random(400) < 0.5, 0.0, -0.0)
    z[::7, 2] = np.inf
    z[::11, 2] = -np.inf
    z[5, 3] = np.nan
    y[:, 2] = z[:, 0] * 2.0 + 1.0
    y[9, 1] = np.nan
    run(ev, "eval_nonfinite", z, y)

    # the reference's full evaluation size: 201 batches of 128 (image_vae_trainer.py:322-323)
    lab = synth.make_labels("morpho", 25728, 16).numpy()[:, 1:]
    run(ev, "eval_mnist_n25728", codes_for(lab, 16, (1, 2, 3, 4, 5, 6), 26, 0.4, weak=((0, 2, 0.012), (9, 4, 0.013))), lab)

    import scipy
    with open(os.path.join(HERE, "PROVENANCE_eval.txt"), "w") as f:
        f.write("eval_*.npz produced by tests/golden/make_golden_eval.py from the unmodified reference\n"
                f"reference path: {REF} (ashispati/ar-vae, utils/evaluation.py:146-219)\n"
                f"numpy {np.__version__}, scipy {scipy.__version__}, torch {torch.__version__}\n")


if __name__ == "__main__":
    main()
