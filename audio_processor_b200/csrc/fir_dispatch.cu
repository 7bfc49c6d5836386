// fir_dispatch.cu — picks the tensor-core FIR instantiation for a (rate, format, channels) triple.  One kernel per
// case, decided by the input alone: no environment switches, no alternative back ends.
#include "fir_mma.cuh"
#include "resample_generic.cuh"

namespace b2a {

int fir_mma_run_44100(int channels, const void*, i64, int16_t*, u64*, FirMmaPlan*, cudaStream_t, i64 first_tile);
int fir_mma_run_48000(int channels, const void*, i64, int16_t*, u64*, FirMmaPlan*, cudaStream_t, i64 first_tile);
int fir_tmem_run_44100(const void*, i64, int16_t*, u64*, FirMmaPlan*, const GenericParams*, cudaStream_t);
int fir_tmem_run_48000(const void*, i64, int16_t*, u64*, FirMmaPlan*, const GenericParams*, cudaStream_t);
i64 fir_tmem_plan_44100(i64, FirMmaPlan*);
i64 fir_tmem_plan_48000(i64, FirMmaPlan*);

// returns 1 if a tensor-core kernel was launched (plan filled), 0 if this input has no fast path, <0 on error.
// s16 input at the two named rates only; the pre-quantisation float output and every other case use the
// table-driven kernel in resample.cu.
//   stereo, at least one 512-run span: fir_tmem.cuh (tcgen05, fed by 2-D TMA, operand planes in TMEM) takes the spans and,
//           in its spare warps, the clip's head and tail outputs (table-driven, one thread per output) and, for clips of
//           300 spans (25 minutes) or more, the runs behind the last span; shorter clips send those to fir_mma.cuh;
//   mono, or stereo clips shorter than a span (< 5.1 s): fir_mma.cuh (mma.sync, 16-run tiles fed by a bulk-TMA ring).
int fir_fast_dispatch(int in_rate, int fmt, int channels, const void* d_in, i64 n_in, int16_t* d_out_s16, float* d_out_f32,
                      u64* d_energy, FirMmaPlan* plan, const GenericParams* edge, cudaStream_t stream) {
    plan->out_lo = plan->out_hi = 0;
    if (fmt != B2A_FMT_S16 || d_out_f32 || (channels != 1 && channels != 2)) return 0;
    if (in_rate != 44100 && in_rate != 48000) return 0;
    if (channels == 2 && edge) {
        FirMmaPlan head;
        const i64 spans = in_rate == 44100 ? fir_tmem_plan_44100(n_in, &head) : fir_tmem_plan_48000(n_in, &head);
        if (spans > 0) {
            // The (at most 511) runs behind the last span.  Long clips: the FIR kernel's spare warps take them together with
            // the clip's head and tail — 82 k outputs spread over 296 otherwise idle warps hide inside >= 60 us of streaming,
            // whereas a separate mma.sync launch was 5 us + a launch gap on the stream (cfg2 step 0.340 -> 0.331 ms).
            // Short clips keep the separate launch: their FIR kernel is too short to hide the spare warps' work, and in a
            // batch the small kernel overlaps other clips' kernels (cfg4, 10-minute clips: 4.11 ms with it, 4.23 ms without).
            const bool tail_in_spare_warps = spans >= 300;
            FirMmaPlan tail;
            tail.out_lo = tail.out_hi = 0;
            int rc = 0;
            if (!tail_in_spare_warps) {
                const i64 first = head.out_hi / (kFmRT * kFmNout);          // span ends are multiples of 16 runs
                rc = in_rate == 44100 ? fir_mma_run_44100(channels, d_in, n_in, d_out_s16, d_energy, &tail, stream, first)
                                      : fir_mma_run_48000(channels, d_in, n_in, d_out_s16, d_energy, &tail, stream, first);
                if (rc < 0) return rc;
            }
            GenericParams e = *edge;
            e.lo0 = 0; e.hi0 = head.out_lo;
            e.lo1 = rc > 0 ? tail.out_hi : head.out_hi; e.hi1 = e.n_out;
            rc = in_rate == 44100 ? fir_tmem_run_44100(d_in, n_in, d_out_s16, d_energy, &head, &e, stream)
                                  : fir_tmem_run_48000(d_in, n_in, d_out_s16, d_energy, &head, &e, stream);
            if (rc <= 0) return rc < 0 ? rc : B2A_ECUDA;
            plan->out_lo = 0; plan->out_hi = e.n_out;
            return 1;
        }
    }
    return in_rate == 44100 ? fir_mma_run_44100(channels, d_in, n_in, d_out_s16, d_energy, plan, stream, 1)
                            : fir_mma_run_48000(channels, d_in, n_in, d_out_s16, d_energy, plan, stream, 1);
}

}  // namespace b2a
