"""The trainers' latent-loss block (imagevae/image_vae_trainer.py:157-180, measure_vae_trainer.py:111-142)
run twice on the GPU: once on stock PyTorch (the reference's op chain, oracle/torch_port.py) and once after
``arvae_b200.install()`` swapped compute_reg_loss / compute_kld_loss / reparametrize on the same classes --
zero edits to the calling code, same RNG stream, same loss and gradients."""
import pytest
import torch

from util import assert_grad_close, assert_loss_close

pytestmark = pytest.mark.gpu


def _make_classes():
    from oracle import torch_port

    class Trainer:  # stand-in for utils.trainer.Trainer: the three static methods the drop-in replaces
        compute_reg_loss = staticmethod(torch_port.compute_reg_loss)
        reg_loss_sign = staticmethod(torch_port.reg_loss_sign)
        compute_kld_loss = staticmethod(torch_port.compute_kld_loss)

    class VAE:  # stand-in for MnistVAE: only reparametrize matters (imagevae/mnist_vae.py:74-87)
        def reparametrize(self, z_dist):
            z_tilde, z_prior, _, prior = torch_port.reparametrize(z_dist.loc, z_dist.scale)
            return z_tilde, z_prior, prior

    return Trainer, VAE


def latent_block(trainer, vae, loc, log_std, labels, reg_dim, beta, capacity, gamma, delta):
    """Written the way the reference trainers write it."""
    z_dist = torch.distributions.Normal(loc=loc, scale=torch.exp(log_std))
    z_tilde, z_prior, prior_dist = vae.reparametrize(z_dist)
    dist_loss = trainer.compute_kld_loss(z_dist, prior_dist, beta=beta, c=capacity)
    loss = dist_loss
    reg_loss = 0.0
    if type(reg_dim) == tuple:
        for dim in reg_dim:
            reg_loss += trainer.compute_reg_loss(z_tilde, labels[:, dim], dim, gamma=gamma, factor=delta)
    else:
        raise TypeError("Regularization dimension must be a tuple of integers")
    loss = loss + reg_loss
    return loss, z_tilde, z_prior


@pytest.mark.parametrize("name,beta,gamma,delta,capacity", [
    ("c1_mnist_b64", 4.0, 10.0, 1.0, torch.FloatTensor([0.0])),       # image trainer: [1]-shaped capacity
    ("c3_measure_b2048", 0.001, 1.0, 10.0, 0.0),                      # measure trainer: python float
    ("c2_dsprites_b4096", 1.0, 10.0, 1.0, torch.FloatTensor([50.0])),
])
def test_trainer_block_is_unchanged_by_install(name, beta, gamma, delta, capacity):
    import arvae_b200
    from arvae_b200 import synth
    c = synth.make_case(name)
    loc0, log_std0, _ = synth.make_latent_head(c["B"], c["Z"], 31)
    labels = c["labels"].cuda()
    cap = capacity.cuda() if isinstance(capacity, torch.Tensor) else capacity
    Trainer, VAE = _make_classes()
    out = []
    for use_drop_in in (False, True):
        if use_drop_in:
            arvae_b200.install(Trainer, vae_classes=[VAE])
        try:
            loc = loc0.cuda().requires_grad_(True)
            log_std = log_std0.cuda().requires_grad_(True)
            torch.manual_seed(1234)
            loss, z_tilde, z_prior = latent_block(Trainer(), VAE(), loc, log_std, labels, c["reg_dims"], beta, cap,
                                                  gamma, delta)
            loss.sum().backward()
            after = torch.rand(1, device="cuda")  # the RNG stream must be in the same state afterwards
            out.append((loss.detach().clone(), z_tilde.detach().clone(), z_prior.clone(), loc.grad.clone(),
                        log_std.grad.clone(), after))
        finally:
            if use_drop_in:
                arvae_b200.uninstall()
    (l0, zt0, zp0, gl0, gs0, a0), (l1, zt1, zp1, gl1, gs1, a1) = out
    assert l0.shape == l1.shape
    assert torch.equal(zt0, zt1) and torch.equal(zp0, zp1) and torch.equal(a0, a1)
    assert_loss_close(l1.sum().item(), l0.sum().item())
    assert_grad_close(gl1.cpu().numpy(), gl0.cpu().numpy())
    assert_grad_close(gs1.cpu().numpy(), gs0.cpu().numpy())
    # the stand-in classes got their own methods back
    from oracle import torch_port
    assert Trainer.compute_reg_loss is torch_port.compute_reg_loss
