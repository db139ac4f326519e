#!/usr/bin/env python
"""Every kernel family of the library on small shapes, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python bench_tools/sanitize_small.py
    compute-sanitizer --tool racecheck python bench_tools/sanitize_small.py
    compute-sanitizer --tool initcheck python bench_tools/sanitize_small.py

Dense pair kernel, cluster sort + rank merge + plan + attribute-sorted pair kernel (forced at a small batch), triangle
variant, one-launch latent-loss head and its backward, latent head, the NVLink-sharded step with 3 virtual ranks, the
evaluation metrics and the music attribute extractor.  Results are checked against each other so that a sanitizer run
is also a functional run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import arvae_b200
from arvae_b200 import distributed as adist, evaluation, music, synth

torch.manual_seed(0)
B, Z = 1500, 6
c = synth.make_case("c2_dsprites_b4096", B)
z = torch.randn(B, c["Z"]).cuda()
lab = c["labels"].cuda()
dims = (1, 2, 3)
z[7, 2] = 40.0  # one outlier: two-MUFU tiles in the sorted path
res = {}
for name, algo in (("dense", arvae_b200.ALGO_DENSE), ("sorted", arvae_b200.ALGO_SORTED), ("triangle", arvae_b200.ALGO_TRIANGLE)):
    loss, grad, _ = arvae_b200.reg_loss_rows(z, lab, dims, 10.0, 1.0, 0, B, algo=algo)
    res[name] = (loss.item(), grad.clone())
for name in ("sorted", "triangle"):
    assert abs(res[name][0] - res["dense"][0]) <= 1e-6 * abs(res["dense"][0]), (name, res[name][0], res["dense"][0])
    assert (res[name][1] - res["dense"][1]).abs().max() <= 1e-5 * res["dense"][1].abs().max()
# autograd wrappers + scatter
zz = z.clone().requires_grad_(True)
arvae_b200.reg_loss_fused(zz, lab, dims, 10.0, 1.0).backward()
# one-launch head (forward + backward), both input forms
loc0, log_std0, eps0 = synth.make_latent_head(B, c["Z"], 5)
loc, log_std = loc0.cuda().requires_grad_(True), log_std0.cuda().requires_grad_(True)
zt, scale, kld, reg = arvae_b200.latent_loss_head(loc, log_std, eps0.cuda(), lab, dims, 4.0, 0.0, 10.0, 1.0)
(kld + reg).backward()
sc = torch.exp(log_std0).cuda().requires_grad_(True)
z2, kld2, reg2 = arvae_b200.reparam_kld_reg(loc.detach().requires_grad_(True), sc, eps0.cuda(), lab, dims, 4.0, 0.0, 10.0, 1.0)
(kld2 + reg2).backward()
assert abs(reg2.item() - reg.item()) <= 1e-5 * abs(reg.item())
zh, kld_mean = arvae_b200.latent_head(loc.detach(), sc.detach(), eps0.cuda())
# NVLink-sharded step, three virtual ranks with unequal row counts
sizes = [400, 700, 400]
offs = [0, 400, 1100, 1500]
grp = adist.LocalShardGroup(3, max(sizes), len(dims))
outs = grp.step([z[offs[g]:offs[g + 1]] for g in range(3)], [lab[offs[g]:offs[g + 1]] for g in range(3)], dims, dims, 10.0, 1.0)
outs = grp.step([z[offs[g]:offs[g + 1]] for g in range(3)], [lab[offs[g]:offs[g + 1]] for g in range(3)], dims, dims, 10.0, 1.0)
torch.cuda.synchronize()
assert outs[0][0].item() == res["sorted"][0], (outs[0][0].item(), res["sorted"][0])
assert torch.equal(torch.cat([o[2] for o in outs]), res["sorted"][1])
grp.close()
# evaluation metrics and music attributes
codes = torch.randn(600, 5).cuda()
attrs = synth.make_labels("morpho", 600, 3).cuda()
evaluation.rank_metrics(codes, attrs)
n2i, _ = synth.music_vocabulary()
music.MeasureAttributeExtractor(n2i)(synth.make_measures(300, 2).cuda())
torch.cuda.synchronize()
print("sanitize_small: ok", res["dense"][0])
