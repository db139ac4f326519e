"""Device-side attribute extractors for MeasureVAE.

The reference computes the labels its regularizer consumes on the fly from the integer score
(measurevae/measure_vae_trainer.py:133,167-186 -> data/dataloaders/bar_dataset.py:338-500); two of the four
extractors are per-sample Python loops with ``.item()`` and one music21 call per tick.  Here a note-index ->
MIDI lookup table is built once from the dataset's dictionaries and all four attributes come out of one kernel
launch, as a [B, 4] float tensor in ``MUSIC_REG_TYPE`` order -- exactly what ``compute_attribute_labels`` returns.

Same method names / shapes as the reference's ``BarDataset`` extractors, so an instance can stand in for
``trainer.dataset`` in ``MeasureVAETrainer.compute_attribute_labels``.
"""
from __future__ import annotations

import ctypes
import re
from typing import Callable, Mapping, Optional, Sequence

import torch

from . import _lib

SLUR_SYMBOL, START_SYMBOL, END_SYMBOL = "__", "START", "END"   # bar_dataset_helpers.py:8-10
MUSIC_REG_TYPE = {"rhy_complexity": 0, "pitch_range": 1, "note_density": 2, "contour": 3}  # measure_vae_trainer.py:15-20
# bar_dataset_helpers.py:21-30 (metrical weights of the 24 ticks of a 4/4 bar)
RHY_COMPLEXITY_COEFFS = (0.20, 1, 2, 0.5, 2, 1, 0.67, 1, 2, 0.5, 2, 1, 0.25, 1, 2, 0.5, 2, 1, 0.67, 1, 2, 0.5, 2, 1)

CODE_SLUR, CODE_REST, CODE_NONE, CODE_START, CODE_END = -1, -2, -3, -4, -5
_STEP = {"C": 0, "D": 2, "E": 4, "F": 5, "G": 7, "A": 9, "B": 11}


def midi_from_pitch_name(name: str) -> int:
    """MIDI number of a music21-style pitch name ('C4' = 60, 'F#5', 'B-3', 'E--4'): what
    ``music21.pitch.Pitch(name).midi`` returns for the names in the datasets' dictionaries."""
    m = re.fullmatch(r"([A-Ga-g])([#\-]*)(-?\d+)", name)
    if not m:
        raise ValueError(f"not a pitch name: {name!r}")
    step, acc, octave = m.group(1).upper(), m.group(2), int(m.group(3))
    # the octave's own minus sign is part of group 3 only when no accidental '-' precedes a digit ambiguity;
    # dataset names never use negative octaves
    return 12 * (octave + 1) + _STEP[step] + acc.count("#") - acc.count("-")


def build_lut(note2index: Mapping, midi_of: Callable[[str], int] = midi_from_pitch_name) -> torch.Tensor:
    """int32 [V] table: MIDI pitch for note symbols, negative codes for slur / rest / None / START / END."""
    V = max(int(i) for i in note2index.values()) + 1
    lut = torch.full((V,), CODE_NONE, dtype=torch.int32)
    special = {SLUR_SYMBOL: CODE_SLUR, "rest": CODE_REST, None: CODE_NONE, START_SYMBOL: CODE_START,
               END_SYMBOL: CODE_END}
    for sym, idx in note2index.items():
        lut[int(idx)] = special[sym] if sym in special else int(midi_of(sym))
    return lut


class MeasureAttributeExtractor:
    """``extractor(measure_tensor) -> [B, 4]`` = (rhy_complexity, pitch_range, note_density, contour)."""

    def __init__(self, note2index: Mapping, midi_of: Callable[[str], int] = midi_from_pitch_name, device="cuda",
                 weights: Sequence[float] = RHY_COMPLEXITY_COEFFS, strict: bool = True):
        """``strict``: indices outside the note dictionary fail (a device-side assertion, no host synchronisation) --
        the reference raises ``KeyError`` for them in pitch range / contour (bar_dataset.py:380, 489).  With
        ``strict=False`` they are treated like the ``None`` symbol."""
        self.strict = bool(strict)
        self.note2index_dicts = dict(note2index)
        self.lut = build_lut(note2index, midi_of).to(device)
        self.weights = torch.tensor(weights, dtype=torch.float64).float().to(device)  # .float() of the float64 table

    def __call__(self, measure_tensor: torch.Tensor) -> torch.Tensor:
        if not measure_tensor.is_cuda:
            raise RuntimeError("arvae_b200: measure_tensor must be a CUDA tensor (no CPU fallback exists for this path)")
        if measure_tensor.dim() != 2 or measure_tensor.dtype != torch.int64:
            raise RuntimeError("arvae_b200: measure_tensor must be [batch, ticks] int64")
        m = measure_tensor if measure_tensor.stride(1) == 1 else measure_tensor.contiguous()
        B, T = m.shape
        if self.strict and B > 0:
            torch._assert_async(((m >= 0) & (m < self.lut.numel())).all())
        if T != self.weights.numel():
            raise RuntimeError(f"arvae_b200: {T} ticks per measure but {self.weights.numel()} rhythmic weights")
        out = torch.empty((B, 4), dtype=torch.float32, device=m.device)
        with torch.cuda.device(m.device):
            rc = _lib.load().arvae_measure_attributes_i64(
                ctypes.c_void_p(m.data_ptr()), B, T, m.stride(0), ctypes.c_void_p(self.lut.data_ptr()),
                self.lut.numel(), ctypes.c_void_p(self.weights.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                ctypes.c_void_p(torch.cuda.current_stream(m.device).cuda_stream))
            _lib.check(rc, "arvae_measure_attributes_i64")
        return out

    # --- the reference's method names (bar_dataset.py:338,360,442,470), each returning [B] ---
    def get_rhy_complexity(self, measure_tensor):
        return self(measure_tensor)[:, 0]

    def get_pitch_range_in_measure(self, measure_tensor):
        return self(measure_tensor)[:, 1]

    def get_note_density_in_measure(self, measure_tensor):
        return self(measure_tensor)[:, 2]

    def get_contour(self, measure_tensor):
        return self(measure_tensor)[:, 3]

    def compute_attribute_labels(self, score: torch.Tensor, attr_list: Optional[Sequence[str]] = None) -> torch.Tensor:
        """Drop-in for ``MeasureVAETrainer.compute_attribute_labels`` (measure_vae_trainer.py:167-186): one launch."""
        all4 = self(score)
        if attr_list is None:
            return all4
        for name in attr_list:
            if name not in MUSIC_REG_TYPE:
                raise ValueError("Invalid regularization attribute")
        return all4[:, [MUSIC_REG_TYPE[n] for n in attr_list]]
