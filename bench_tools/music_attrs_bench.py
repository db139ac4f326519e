#!/usr/bin/env python
"""Attribute-label extraction for a MeasureVAE batch: CUDA kernel vs the reference's per-sample Python loops
(restated in oracle/music_attrs.py; the reference additionally pays a music21 Pitch() construction per tick)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arvae_b200 import music, synth
from oracle import music_attrs

note2index, index2note = synth.music_vocabulary()
ex = music.MeasureAttributeExtractor(note2index)
for B in (256, 4096, 65536):
    m = synth.make_measures(B, seed=B)
    mc = m.cuda()
    for _ in range(3):
        ex(mc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = ex(mc)
    e1.record()
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / 20
    n_cpu = min(B, 4096)
    t = time.perf_counter()
    ref = music_attrs.all_attributes(m[:n_cpu].numpy(), note2index, index2note, music.midi_from_pitch_name)
    cpu_ms = (time.perf_counter() - t) * 1e3 * (B / n_cpu)
    print(json.dumps({"B": B, "gpu_ms": gpu_ms, "cpu_python_loops_ms": cpu_ms, "bars_per_s_gpu": B / gpu_ms * 1e3,
                      "speedup": cpu_ms / gpu_ms, "algorithmic_bytes": B * 24 * 8 + B * 16}))
