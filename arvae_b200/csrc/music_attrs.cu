// music_attrs.cu -- the four musical attributes MeasureVAE is regularised on, computed on the device.
//
// Reference: data/dataloaders/bar_dataset.py:338-358 (note density), :360-390 (pitch range),
// :442-468 (rhythmic complexity), :470-500 (contour); assembled in MUSIC_REG_TYPE order by
// measurevae/measure_vae_trainer.py:15-20,167-186.  Two of the four are per-sample Python loops with
// `.item()` and a music21 call per tick in the reference -- the real bottleneck of its MeasureVAE step.
// Here: one thread per measure, one pass over its ticks, a note-index -> MIDI lookup table.
//
//   lut[v] >= 0 : MIDI pitch of note symbol v        lut[v] = -1 slur '__', -2 'rest', -3 None, -4 START, -5 END
//   out[b, 0] rhy_complexity = sum_t w_t [tick t holds a note onset] / sum_t w_t      (float; the numerator is
//                              accumulated exactly in double and rounded once)
//   out[b, 1] pitch_range    = (max - min MIDI over note onsets, 0 if fewer than 2) / 26
//   out[b, 2] note_density   = (T - #slur - #rest - #START - #END) / T      (None counts as a note, as in the reference)
//   out[b, 3] contour        = (last - first MIDI over note onsets, 0 if fewer than 2) / 26
#include "common.cuh"

namespace arvae {

__global__ void __launch_bounds__(256)
measure_attributes_kernel(const long long *__restrict__ measures, int64_t B, int64_t T, int64_t row_stride,
                          const int *__restrict__ lut, int64_t V, const float *__restrict__ weights,
                          float *__restrict__ out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long *m = measures + b * row_stride;
    double wsum = 0.0, wnorm = 0.0;
    int excluded = 0, n_notes = 0, lo = 0, hi = 0, first = 0, last = 0;
    for (int64_t t = 0; t < T; ++t) {
        const long long v = m[t];
        const int code = (v >= 0 && v < V) ? __ldg(lut + v) : -3;  // out-of-vocabulary behaves like None
        const float w = weights ? __ldg(weights + t) : 0.0f;
        wnorm += (double)w;
        if (code == -1 || code == -2 || code == -4 || code == -5) ++excluded;  // density ignores None on purpose
        if (code >= 0) {
            wsum += (double)w;
            if (n_notes == 0) { lo = hi = first = code; }
            lo = min(lo, code);
            hi = max(hi, code);
            last = code;
            ++n_notes;
        }
    }
    float *o = out + b * 4;
    o[0] = __fdiv_rn((float)wsum, (float)wnorm);
    o[1] = __fdiv_rn(n_notes >= 2 ? (float)(hi - lo) : 0.0f, 26.0f);
    o[2] = __fdiv_rn((float)(T - excluded), (float)T);
    o[3] = __fdiv_rn(n_notes >= 2 ? (float)(last - first) : 0.0f, 26.0f);
}

int run_measure_attributes(const long long *measures, int64_t B, int64_t T, int64_t row_stride, const int *lut,
                           int64_t V, const float *weights, float *out, cudaStream_t st) {
    if (B <= 0) return 0;
    measure_attributes_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, st>>>(measures, B, T, row_stride, lut, V,
                                                                         weights, out);
    ARVAE_LAUNCH_CHECK("measure_attributes_kernel");
    return 0;
}

}  // namespace arvae
