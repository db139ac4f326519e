"""Musical attribute extractors: the CPU oracle against the reference's own method bodies
(tests/golden/music_attrs.npz), and the CUDA kernel against both.  Integer-valued attributes divided by a
constant (pitch range, note density, contour) must be bit-exact; rhythmic complexity is a float32 sum whose
order the reference leaves to torch, so it is held to 1 ulp-level tolerance (3e-7 relative)."""
import numpy as np
import pytest
import torch

from conftest import golden


def _vocab():
    from arvae_b200 import synth
    return synth.music_vocabulary()


def test_oracle_matches_reference_extractors():
    from arvae_b200 import music
    from oracle import music_attrs
    g = golden("music_attrs")
    note2index, index2note = _vocab()
    got = music_attrs.all_attributes(g["measures"], note2index, index2note, music.midi_from_pitch_name)
    ref = g["attrs"]
    assert got.shape == ref.shape == (g["measures"].shape[0], 4)
    for col in (1, 2, 3):
        assert np.array_equal(got[:, col], ref[:, col]), col
    assert np.allclose(got[:, 0], ref[:, 0], rtol=3e-7, atol=0)
    # the reference's special cases are in the fixture: empty bar, single note, all rests
    assert ref[0, 1] == 0 and ref[0, 3] == 0 and ref[1, 1] == 0 and ref[1, 3] == 0 and ref[2, 2] == 0


def test_pitch_name_rule():
    from arvae_b200 import music
    assert music.midi_from_pitch_name("C4") == 60
    assert music.midi_from_pitch_name("A4") == 69
    assert music.midi_from_pitch_name("F#5") == 78
    assert music.midi_from_pitch_name("B-3") == 58
    assert music.midi_from_pitch_name("E--4") == 62
    with pytest.raises(ValueError):
        music.midi_from_pitch_name("rest")


def test_lut_layout():
    from arvae_b200 import music
    note2index, _ = _vocab()
    lut = music.build_lut(note2index)
    assert lut[note2index["__"]] == -1 and lut[note2index["rest"]] == -2 and lut[note2index[None]] == -3
    assert lut[note2index["START"]] == -4 and lut[note2index["END"]] == -5
    assert lut[note2index["C4"]] == 60 and lut.dtype == torch.int32


@pytest.mark.gpu
def test_kernel_matches_reference_and_oracle():
    from arvae_b200 import music, synth
    from oracle import music_attrs
    g = golden("music_attrs")
    note2index, index2note = _vocab()
    ex = music.MeasureAttributeExtractor(note2index)
    out = ex(torch.from_numpy(g["measures"]).cuda()).cpu().numpy()
    for col in (1, 2, 3):
        assert np.array_equal(out[:, col], g["attrs"][:, col]), col
    assert np.allclose(out[:, 0], g["attrs"][:, 0], rtol=3e-7, atol=0)
    # a larger batch against the oracle, and the reference's method names / shapes
    m = synth.make_measures(5000, seed=5)
    ref = music_attrs.all_attributes(m.numpy(), note2index, index2note, music.midi_from_pitch_name)
    mc = m.cuda()
    got = ex(mc).cpu().numpy()
    assert np.array_equal(got[:, 1:], ref[:, 1:])
    assert np.allclose(got[:, 0], ref[:, 0], rtol=3e-7, atol=0)
    assert tuple(ex.get_contour(mc).shape) == (5000,)
    labels = ex.compute_attribute_labels(mc, ["note_density", "rhy_complexity"])
    assert torch.equal(labels[:, 0], ex.get_note_density_in_measure(mc)) and torch.equal(labels[:, 1], ex.get_rhy_complexity(mc))
    with pytest.raises(ValueError):
        ex.compute_attribute_labels(mc, ["tempo"])
    # identical bars always get identical attribute values (what keeps the sign matrix stable)
    dup = torch.cat([mc[:100], mc[:100]], 0)
    o = ex(dup)
    assert torch.equal(o[:100], o[100:])
    # non-contiguous input (a slice of a longer score) is accepted
    wide = torch.cat([mc, mc], dim=1)
    assert torch.equal(ex(wide[:, :24]), ex(mc))


@pytest.mark.gpu
def test_attributes_feed_the_regularizer():
    """End of the MeasureVAE label path: extractor output -> reg loss, like measure_vae_trainer.py:131-142."""
    import arvae_b200
    from arvae_b200 import music, synth
    note2index, _ = _vocab()
    ex = music.MeasureAttributeExtractor(note2index)
    m = synth.make_measures(2048, seed=9).cuda()
    attr_labels = ex.compute_attribute_labels(m)
    z = torch.randn(2048, 32, device="cuda", requires_grad=True)
    reg_loss = 0.0
    for dim in (0, 1, 2, 3):
        reg_loss += arvae_b200.compute_reg_loss(z, attr_labels[:, dim], dim, gamma=1.0, factor=10.0)
    reg_loss.backward()
    fused = arvae_b200.reg_loss_fused(z.detach(), attr_labels, (0, 1, 2, 3), 1.0, 10.0)
    assert abs(reg_loss.item() - fused.item()) <= 1e-6 * fused.item()
    assert torch.isfinite(z.grad).all() and z.grad[:, 4:].abs().sum() == 0
