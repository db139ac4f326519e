"""CPU restatement of the reference's musical attribute extractors (TEST INFRASTRUCTURE ONLY).

Follows data/dataloaders/bar_dataset.py of the reference:
  :338-358 get_note_density_in_measure   :360-390 get_pitch_range_in_measure
  :442-468 get_rhy_complexity            :470-500 get_contour
with the note dictionaries passed in and ``midi_of(name)`` standing for ``music21.pitch.Pitch(name).midi``
(music21 is not installed here; the dataset dictionaries only hold plain pitch names).  Pure-Python loops,
like the reference's; pinned to tests/golden/music_attrs.npz, which tests/golden/make_golden.py produced by
running the reference's own method bodies with a stub ``music21.pitch.Pitch``.
"""
from __future__ import annotations

import numpy as np

SLUR_SYMBOL, START_SYMBOL, END_SYMBOL = "__", "START", "END"
RHY_COMPLEXITY_COEFFS = np.array([0.20, 1, 2, 0.5, 2, 1, 0.67, 1, 2, 0.5, 2, 1, 0.25, 1, 2, 0.5, 2, 1,
                                  0.67, 1, 2, 0.5, 2, 1])


def _special(note2index):
    return (note2index[SLUR_SYMBOL], note2index["rest"], note2index[None], note2index[START_SYMBOL],
            note2index[END_SYMBOL])


def note_density(measures: np.ndarray, note2index) -> np.ndarray:
    slur, rest, none, start, end = _special(note2index)
    T = measures.shape[1]
    excluded = ((measures == slur).sum(1) + (measures == rest).sum(1) + (measures == start).sum(1)
                + (measures == end).sum(1))  # None is NOT excluded, as in the reference
    return ((T - excluded).astype(np.float32) / np.float32(T)).astype(np.float32)


def _midi_notes(row, note2index, index2note, midi_of):
    skip = set(_special(note2index))
    return [int(midi_of(index2note[int(i)])) for i in row if int(i) not in skip]


def pitch_range(measures, note2index, index2note, midi_of) -> np.ndarray:
    out = np.zeros(measures.shape[0], dtype=np.float32)
    for b, row in enumerate(measures):
        notes = _midi_notes(row, note2index, index2note, midi_of)
        out[b] = 0 if len(notes) < 2 else max(notes) - min(notes)
    return (out / np.float32(26)).astype(np.float32)


def contour(measures, note2index, index2note, midi_of) -> np.ndarray:
    out = np.zeros(measures.shape[0], dtype=np.float32)
    for b, row in enumerate(measures):
        notes = _midi_notes(row, note2index, index2note, midi_of)
        out[b] = 0 if len(notes) < 2 else int(np.sum(np.diff(np.asarray(notes, dtype=np.float32))))
    return (out / np.float32(26)).astype(np.float32)


def rhy_complexity(measures, note2index) -> np.ndarray:
    skip = _special(note2index)
    beat = np.ones(measures.shape, dtype=np.float32)
    for s in skip:
        beat[measures == s] = 0
    w = RHY_COMPLEXITY_COEFFS.astype(np.float32)
    return ((w[None, :] * beat).sum(1, dtype=np.float32) / w.sum(dtype=np.float32)).astype(np.float32)


def all_attributes(measures, note2index, index2note, midi_of) -> np.ndarray:
    """[B, 4] in MUSIC_REG_TYPE order (measurevae/measure_vae_trainer.py:15-20)."""
    return np.stack([rhy_complexity(measures, note2index), pitch_range(measures, note2index, index2note, midi_of),
                     note_density(measures, note2index), contour(measures, note2index, index2note, midi_of)], axis=1)
