// gen_mel_tables.cpp — build-time generator of csrc/mel_tables_gen.inc.
// Emits the sparse slaney filterbank (first bin, width, weights) for 80 and 128 mels as `static __device__ const`
// tables; the log-mel kernel unrolls over them so every weight becomes an FFMA immediate.  Weights carry the 1/4 of
// the real-FFT unpack (|X|^2 = 0.25 * |2X|^2), an exact power-of-two scale.  Mels are split into 5 groups (one per
// warp of the kernel) balanced by the kernel's cost model (quads of taps + per-mel epilogue).
#include <cstdio>
#include <algorithm>
#include <vector>
#include "../audio_processor_b200/csrc/mel_design.h"

static const int kBins = 201, kWidth = 16, kGroups = 5;

static int emit(FILE* f, int nm) {
    std::vector<float> filt((size_t)nm * kBins);
    b2a_design::design_mel(nm, kBins, 16000.0, filt.data());
    std::vector<int> start(nm), len(nm);
    int total = 0;
    for (int m = 0; m < nm; m++) {
        int first = -1, last = -1;
        for (int k = 0; k < kBins; k++)
            if (filt[(size_t)m * kBins + k] != 0.0f) { if (first < 0) first = k; last = k; }
        if (first < 0) { first = 0; last = -1; }
        start[m] = first; len[m] = last - first + 1;
        if (len[m] > kWidth) { fprintf(stderr, "mel %d wider than %d bins\n", m, kWidth); return 1; }
        total += 9 * ((len[m] + 3) / 4) + 14;      // kernel cost model: 9 instructions per quad of taps + 14 per mel
    }
    fprintf(f, "// %d mels: sparse slaney filterbank, weights x 0.25\n", nm);
    fprintf(f, "template <> struct MelC<%d> {\n    static constexpr int start[%d] = {", nm, nm);
    for (int m = 0; m < nm; m++) fprintf(f, "%d,%s", start[m], (m % 20 == 19) ? "\n" : " ");
    fprintf(f, "};\n    static constexpr int len[%d] = {", nm);
    for (int m = 0; m < nm; m++) fprintf(f, "%d,%s", len[m], (m % 20 == 19) ? "\n" : " ");
    fprintf(f, "};\n");
    // group boundaries: greedy split at ~equal cumulative cost
    int bounds[kGroups + 1]; bounds[0] = 0; bounds[kGroups] = nm;
    int acc = 0, g = 1;
    for (int m = 0; m < nm && g < kGroups; m++) {
        acc += 9 * ((len[m] + 3) / 4) + 14;
        if (acc * kGroups >= total * g) bounds[g++] = m + 1;
    }
    for (; g < kGroups; g++) bounds[g] = nm;
    fprintf(f, "    static constexpr int group[%d] = {", kGroups + 1);
    for (int i = 0; i <= kGroups; i++) fprintf(f, "%d, ", bounds[i]);
    fprintf(f, "};\n    static constexpr int seg[%d][5] = {", kGroups);
    {
        std::vector<int> q2(nm);
        int prev = 1;
        for (int m = 0; m < nm; m++) { q2[m] = std::max(prev, (len[m] + 3) / 4); prev = q2[m]; }
        for (int g2 = 0; g2 < kGroups; g2++) {
            fprintf(f, "{");
            for (int q = 0; q <= 4; q++) {
                int m = bounds[g2];
                while (m < bounds[g2 + 1] && q2[m] <= q) m++;
                fprintf(f, "%d, ", q == 4 ? bounds[g2 + 1] : m);
            }
            fprintf(f, "}, ");
        }
    }
    fprintf(f, "};\n};\n");
    // table-driven form for the kernel's mel loop: weights flattened, every filter padded with zeros to whole quads of
    // taps (one 16-byte load per 4 taps); quads[m] is made non-decreasing in m (filters widen with frequency; a narrower
    // straggler is padded), so each warp's group is a few runs of constant quad count = loops without a per-mel branch.
    // desc[m] = {byte offset of the first power bin, byte offset of the first weight quad}; seg[g][q] = first mel of
    // group g with more than q quads (seg[g][0] = group start, seg[g][4] = group end).
    std::vector<int> quads(nm);
    {
        int prev = 1;
        for (int m = 0; m < nm; m++) { quads[m] = std::max(prev, (len[m] + 3) / 4); prev = quads[m]; }
    }
    {
        std::vector<float> flat;
        std::vector<unsigned> desc(2 * nm);
        for (int m = 0; m < nm; m++) {
            desc[2 * m] = (unsigned)start[m] * 4u;
            desc[2 * m + 1] = (unsigned)flat.size() * 4u;
            for (int j = 0; j < 4 * quads[m]; j++) flat.push_back(j < len[m] ? 0.25f * filt[(size_t)m * kBins + start[m] + j] : 0.0f);
        }
        fprintf(f, "static __device__ const unsigned kMelDesc%d[%d] = {", nm, 2 * nm);
        for (int m = 0; m < 2 * nm; m++) fprintf(f, "%uu,%s", desc[m], (m % 10 == 9) ? "\n" : " ");
        fprintf(f, "};\nconstexpr int kMelFlatN%d = %d;\nstatic __device__ const float kMelFlat%d[%d] = {\n", nm, (int)flat.size(), nm, (int)flat.size());
        for (size_t i = 0; i < flat.size(); i++) fprintf(f, "%.9ef,%s", flat[i], (i % 4 == 3) ? "\n" : " ");
        fprintf(f, "};\n");
    }
    fprintf(f, "static __device__ const float kMelW%d[%d] = {\n", nm, nm * kWidth);
    for (int m = 0; m < nm; m++) {
        for (int j = 0; j < kWidth; j++)
            fprintf(f, "%.9ef, ", j < len[m] ? 0.25f * filt[(size_t)m * kBins + start[m] + j] : 0.0f);
        fprintf(f, "\n");
    }
    fprintf(f, "};\n");
    return 0;
}

int main(int argc, char** argv) {
    FILE* f = argc > 1 ? fopen(argv[1], "w") : stdout;
    if (!f) return 1;
    fprintf(f, "// GENERATED by tools/gen_mel_tables.cpp — do not edit.\n");
    fprintf(f, "template <int NM> struct MelC;   // start[m], len[m]: non-zero bins of mel m; group[g]: mels of warp g\n");
    int rc = emit(f, 80) | emit(f, 128);
    if (f != stdout) fclose(f);
    return rc;
}
