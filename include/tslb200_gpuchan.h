/*
 * tslb200_gpuchan.h -- C ABI of the B200 channel bank (CUDA layer).
 *
 * One gpuchan_t replaces the N per-channel `struct demod_thread` workers that the
 * reference's receiver fans every IQ buffer out to:
 *
 *   reference seam                                    replaced by
 *   ------------------------------------------------  --------------------------------
 *   multifm/demod.c:205-261  _demod_fir_prepare       gpuchan_prepare_taps / gpuchan_create
 *   filter/direct_fir.c:44-87 direct_fir_init         gpuchan_create (derotator increment)
 *   multifm/receiver.c:78-98 receiver_sample_buf_deliver
 *     -> multifm/demod.c:49-121 demod_thread_process  gpuchan_submit / gpuchan_submit_device
 *        filter/direct_fir.c:422 direct_fir_process   (fused kernel: mix+FIR+decimate+derotate)
 *        multifm/fm_demod.c:36 multifm_fm_demod_process (same kernel: discriminator)
 *   multifm/demod.c:93 write(fifo_fd, out_buf, ...)   gpuchan_collect (int16 PCM per channel)
 *   multifm/demod.c:75-81 signalDebugFile tap         gpuchan_collect_iq (GPUCHAN_F_KEEP_IQ)
 *
 * The stream contract is the reference's: output k of a channel is produced from input
 * samples [k*D, k*D+T) as soon as they have been submitted; submits may have any length
 * (4096-sample sample_bufs or megasample batches) and give identical output streams.
 *
 * Return convention is aresult_t compatible: 0 == A_OK, negative == error, FAILED(x) == (x != 0).
 * Plain pointers and sizes only; no CUDA or torch types cross this boundary
 * (a cudaStream_t travels as void *).
 */
#ifndef TSLB200_GPUCHAN_H
#define TSLB200_GPUCHAN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPUCHAN_OK            0
#define GPUCHAN_E_NOMEM     (-1)
#define GPUCHAN_E_BADARGS   (-2)
#define GPUCHAN_E_BUSY      (-4)
#define GPUCHAN_E_INVAL     (-5)
#define GPUCHAN_E_CUDA      (-64)   /* a CUDA runtime call failed; see gpuchan_last_error() */
#define GPUCHAN_E_NODEVICE  (-65)   /* no usable sm_100 device: there is NO CPU fallback */

/* flags */
#define GPUCHAN_F_ATAN_FMA   0x1u   /* fuse the one FMA gcc fuses in fast_atan2f.c:131 (reference Release build) */
#define GPUCHAN_F_KEEP_IQ    0x2u   /* also keep the post-FIR int16 IQ (signalDebugFile tap) */
#define GPUCHAN_F_DEFAULT    GPUCHAN_F_ATAN_FMA

/* kernel selection (engine) */
#define GPUCHAN_ENGINE_AUTO   0     /* tensor-core engine whenever its plan fits (measured faster even for 1 channel), else IMAD */
#define GPUCHAN_ENGINE_IMAD   1     /* exact int32 CUDA-core kernel */
#define GPUCHAN_ENGINE_TC     2     /* exact int8-limb tcgen05 kernel */

typedef struct gpuchan gpuchan_t;

typedef struct gpuchan_cfg {
    uint32_t struct_size;        /* = sizeof(gpuchan_cfg) */
    uint32_t sample_rate_hz;     /* sampleRateHz      (multifm/receiver.c:139) */
    uint32_t decimation;         /* decimationFactor  (multifm/receiver.c:160) */
    uint32_t nr_taps;            /* lpfTaps length    (multifm/receiver.c:166-184) */
    uint32_t nr_channels;        /* channels[] length (multifm/receiver.c:195) */
    int32_t  device;             /* CUDA ordinal */
    uint32_t max_batch_samples;  /* largest n_complex one submit may carry */
    uint32_t flags;              /* GPUCHAN_F_* */
    uint32_t engine;             /* GPUCHAN_ENGINE_* */
    uint32_t reserved;
    const double  *lpf_taps;     /* [nr_taps] real low-pass prototype, shared by all channels */
    const int32_t *offset_hz;    /* [nr_channels] chanCenterFreq - centerFreqHz (multifm/receiver.c:229) */
    const double  *gain;         /* [nr_channels] linear gain = 10^(dBGain/10) (receiver.c:220); NULL = 1.0 */
} gpuchan_cfg;

/* a1 on the host, in double with libm, exactly as the reference prepares taps. */
int gpuchan_prepare_taps(const double *lpf_taps, size_t nr_taps, int32_t offset_hz, uint32_t sample_rate_hz,
                         double gain, int16_t *c_re, int16_t *c_im);
int gpuchan_derot_increment(int32_t offset_hz, uint32_t sample_rate_hz, uint32_t decimation, int16_t incr[2]);
double gpuchan_db_to_gain(double db_gain);

int gpuchan_create(gpuchan_t **ph, const gpuchan_cfg *cfg);
int gpuchan_destroy(gpuchan_t **ph);

/* Append n_complex interleaved int16 I,Q samples from HOST memory (pinned or pageable) and launch the
 * kernels for every output that became computable.  Asynchronous; iq_host may be reused after return
 * only if it is pageable (pinned buffers must stay untouched until gpuchan_sync/collect). */
int gpuchan_submit(gpuchan_t *h, const int16_t *iq_host, size_t n_complex);

/* 8-bit captures: the bytes cross PCIe as they are (half the traffic of cs16) and are widened on the device with
 * the reference's own conversions:
 *   GPUCHAN_FMT_CS8     (int16_t)(int8_t)b                       multifm/file_if.c:67-110  (fileFormat "cs8")
 *   GPUCHAN_FMT_CU8     (int16_t)(int8_t)b - 127                 multifm/file_if.c:112-157 (fileFormat "cu8"; the
 *                       reference reads the unsigned bytes through an int8_t pointer -- reproduced as is)
 *   GPUCHAN_FMT_CU8_RTL ((int16_t)(uint8_t)b - 127) << 7         multifm/rtl_sdr_if.c:142-147 (live RTL-SDR contract)
 * iq8_host holds 2 * n_complex bytes (I, Q interleaved).  Otherwise identical to gpuchan_submit. */
#define GPUCHAN_FMT_CS8     1u
#define GPUCHAN_FMT_CU8     2u
#define GPUCHAN_FMT_CU8_RTL 3u
int gpuchan_submit_bytes(gpuchan_t *h, const uint8_t *iq8_host, size_t n_complex, uint32_t format);

/* Same, with the samples already resident on h's device (e.g. after an NCCL broadcast).  cuda_stream
 * (a cudaStream_t) only tells WHEN d_iq becomes readable: the bank waits for that stream's current position and
 * then works on its own streams, so the producer of batch i+1 overlaps the kernels of batch i.  NULL = readable
 * now.  d_iq must stay unmodified until the batch has been collected / discarded or gpuchan_sync returned. */
int gpuchan_submit_device(gpuchan_t *h, const int16_t *d_iq, size_t n_complex, void *cuda_stream);
/* Order cuda_stream after everything submitted so far (event timing, chaining consumers on a caller stream). */
int gpuchan_stream_wait(gpuchan_t *h, void *cuda_stream);

/* Wait for all submitted work. */
int gpuchan_sync(gpuchan_t *h);

/* Up to two submits may be in flight (the H2D copy of batch i overlaps the kernels of batch i-1 and the D2H
 * of batch i-2); a third submit returns GPUCHAN_E_BUSY until the oldest batch is collected or discarded.
 * Results come back in FIFO order. */
int gpuchan_in_flight(gpuchan_t *h);

/* Number of PCM samples per channel in the OLDEST uncollected batch (identical for all channels; 0 if none). */
int gpuchan_pending(gpuchan_t *h, size_t *n_per_channel);

/* Copy the oldest uncollected batch's PCM to host: channel c occupies pcm_host[c*cap_per_channel ... + n).
 * Blocks until the data has arrived, then retires the batch. */
int gpuchan_collect(gpuchan_t *h, int16_t *pcm_host, size_t cap_per_channel, size_t *n_per_channel);
/* The same in two halves: _begin enqueues the copy-out of the oldest batch (asynchronous), _end waits for it and retires the
 * batch.  Lets a caller with several banks (gpuchan_multi_collect) run all copy-outs side by side. */
int gpuchan_collect_begin(gpuchan_t *h, int16_t *pcm_host, size_t cap_per_channel);
int gpuchan_collect_end(gpuchan_t *h, size_t *n_per_channel);
/* Post-FIR IQ tap (2 int16 per output) of the batch most recently returned by gpuchan_collect;
 * needs GPUCHAN_F_KEEP_IQ and must be called before that batch's slot is reused (two submits later). */
int gpuchan_collect_iq(gpuchan_t *h, int16_t *iq_host, size_t cap_per_channel, size_t *n_per_channel);
/* Retire the oldest batch without copying it (results consumed on the device, or not wanted). */
int gpuchan_discard(gpuchan_t *h);

/* Device view of the MOST RECENT submit's PCM for chaining on-GPU consumers (pager bank): channel c starts at
 * d_pcm + c * pitch_samples. Valid until that slot is reused (two submits later). */
int gpuchan_device_pcm(gpuchan_t *h, const int16_t **d_pcm, size_t *pitch_samples, size_t *n_per_channel);

/* Introspection for parity tests */
int gpuchan_get_taps(gpuchan_t *h, uint32_t channel, int16_t *c_re, int16_t *c_im);
int gpuchan_get_rot_state(gpuchan_t *h, uint32_t channel, int16_t rot[2], int16_t incr[2],
                          uint64_t *outputs_so_far, uint32_t *cycle_mu, uint32_t *cycle_lambda);
int gpuchan_engine(gpuchan_t *h);           /* engine actually in use */
uint64_t gpuchan_kernel_launches(gpuchan_t *h);  /* kernels launched by this bank so far */
const char *gpuchan_last_error(void);

/* Page-locked host buffers (cudaHostAlloc) for callers that do not include CUDA headers: the receiver's
 * batch staging (replaces the frame_alloc pool as the H2D source). */
int gpuchan_host_alloc(void **pp, size_t bytes);
int gpuchan_host_free(void *p);

/* ---- one process, several GPUs ---------------------------------------------------------------------------------
 * The channels of ONE configuration sharded over N devices (contiguous channel ranges, sizes differing by at most
 * one), every batch delivered to all of them: receiver_sample_buf_deliver's fan-out (multifm/receiver.c:78-98) across
 * devices.  cfg describes ALL channels (cfg->device is ignored).  PCM comes back in the configuration's channel
 * order.  Same stream contract and the same two-batches-in-flight rule as a single bank. */
#define GPUCHAN_FANOUT_HOST   0     /* every GPU copies the (pinned) host batch over its own PCIe link */
#define GPUCHAN_FANOUT_RELAY  1     /* the batch enters devices[0]; NVLink relay chain on the copy engines (tslb200_gpurelay.h) */
typedef struct gpuchan_multi gpuchan_multi_t;
int gpuchan_multi_create(gpuchan_multi_t **ph, const gpuchan_cfg *cfg, const int32_t *devices, uint32_t nr_devices, uint32_t fanout);
int gpuchan_multi_destroy(gpuchan_multi_t **ph);
int gpuchan_multi_submit(gpuchan_multi_t *h, const int16_t *iq_host, size_t n_complex);
int gpuchan_multi_pending(gpuchan_multi_t *h, size_t *n_per_channel);
int gpuchan_multi_collect(gpuchan_multi_t *h, int16_t *pcm_host, size_t cap_per_channel, size_t *n_per_channel);
int gpuchan_multi_discard(gpuchan_multi_t *h);
int gpuchan_multi_sync(gpuchan_multi_t *h);
uint32_t gpuchan_multi_devices(gpuchan_multi_t *h);
/* the bank of device i and the channel range it owns (for chaining per-device pager banks on gpuchan_device_pcm) */
int gpuchan_multi_bank(gpuchan_multi_t *h, uint32_t i, gpuchan_t **bank, uint32_t *first_channel, uint32_t *nr_channels);

/* The FM discriminator as an object of its own, for callers written against multifm/fm_demod.h:22-34
 * (multifm_fm_demod_init / _process / _cleanup; include/compat/fm_demod.h wraps exactly this).  One real-valued int16
 * PCM sample per complex int16 input sample (fm_demod.c:53-77); the previous input sample is carried from call to
 * call and starts at (0, 0).  Inside a channel bank the discriminator is fused into the FIR kernel instead. */
typedef struct gpufm gpufm_t;
int gpufm_create(gpufm_t **ph, int32_t device, uint32_t max_samples_per_launch, uint32_t flags /* GPUCHAN_F_ATAN_FMA */);
int gpufm_process(gpufm_t *h, const int16_t *iq_host, size_t n_complex, int16_t *pcm_host);
int gpufm_destroy(gpufm_t **ph);

/* Instrumentation (bench.py roofline): CUDA events around each launch of the dominant FIR+FM kernel. */
int gpuchan_timing_enable(gpuchan_t *h, int on);
int gpuchan_timing_read(gpuchan_t *h, double *total_ms, uint64_t *nr_launches);

/* Tensor-core work of one launch, for the MAC/s roofline: out = { tcgen05.mma (M = 128, K = 32 int8) instructions per
 * tile, their N, PCM outputs per channel per tile, channel groups }; zeros on the IMAD engine. */
int gpuchan_tc_model(gpuchan_t *h, uint64_t out[4]);

/* Unit hook for tests: one tcgen05 kind::i8 tile, D[128][N] = A0*B0[shift0..]^T + A1*B1[shift1..]^T (int32, wraps).
 * All pointers are host pointers; A is 128 x Kp bytes, B is R x Kp bytes, row-major. */
int gpuchan_tc_selftest(const uint8_t *A0, const uint8_t *B0, const uint8_t *A1, const uint8_t *B1, int Kp, int R,
                        int N, int shift0, int shift1, int a0_signed, int b0_signed, int a1_signed, int b1_signed,
                        int32_t *out);

/* Unit hook for tests, host only (works without a device): the tensor-core engine's plan for a configuration.
 * info[16] = { ok, mode (0 sum / 1 radix), accumulators, sample stages, TMEM stages, tap-image chunks, bytes per group
 * image, bytes per sample stage, dynamic shared memory, arctangent table copies, MMAs per tile, first MMA of the second
 * issuing warp, Kp, Q, R, channel groups | channel groups per CTA << 16 }; tap_image (optional) receives
 * [G][chunk][2][128][16] int8 limbs; program
 * (optional) receives 4 words per MMA { a_lo, b_lo, d_acc, idesc } (csrc/tc_engine.cuh).  smem_max_bytes = 0: B200. */
int gpuchan_tc_plan_query(const gpuchan_cfg *cfg, uint32_t smem_max_bytes, uint32_t info[16],
                          uint8_t *tap_image, size_t tap_image_cap, uint32_t *program, size_t program_cap_entries);

/* Unit hook for tests: the kernels' re-formulated discriminator arithmetic (csrc/fm_math.cuh "v2") next to literal
 * transcriptions of multifm/fast_atan2f.c:101-174 and multifm/fm_demod.c:68-72, on the device.
 *   what = 0: `count` pseudo-random int32 operand pairs (seed) through fast_atan2f and the PCM scaling;
 *   what = 1: every float bit pattern in [seed_or_first, seed_or_first + count) with |phi| <= 3.2 through the PCM scaling.
 * out[0] = arctangent results differing in any bit, out[1] = PCM values differing, out[2] = uses of the exact FP64
 * path, out[3] = offenders recorded, out[4..7] = the first offender's operands and results. */
int gpuchan_math_selftest(uint32_t what, uint64_t seed_or_first, uint64_t count, uint32_t use_fma, uint64_t out[8]);

/* Unit hook for tests: the fused kernel's discriminator arithmetic (fm_math.cuh v3, packed pairs) on caller-provided int32
 * operand pairs: phi[i] = fast_atan2f((float)s_im[i], (float)s_re[i]) and pcm[i] = (int16)(float)((double)phi / M_PI * 16384)
 * (multifm/fm_demod.c:66-72), for comparison with fixtures recorded from the reference objects. */
int gpuchan_math_eval(const int32_t *s_im, const int32_t *s_re, size_t n, uint32_t use_fma, float *phi, int16_t *pcm);

#ifdef __cplusplus
}
#endif
#endif /* TSLB200_GPUCHAN_H */
