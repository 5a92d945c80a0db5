/*
 * wwfeat.h - C ABI of libwwfeat.so: the B200 (sm_100a) audio feature / augmentation path.
 *
 * Drop-in boundary for the hot path of sarpel/wakeword_trainer_home.  The reference has
 * no FFI of its own: the path is the Python call surface of its (uncommitted) src/data
 * package.  Each entry point below names the reference interface it stands behind
 * (paths relative to the reference checkout) so a maintainer can bind it with ctypes
 * (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain C types only; every buffer is a raw pointer + sizes; no torch types.
 *   - "dev" pointers are CUDA device pointers on the plan's device, owned by the caller
 *     (PyTorch's allocator); "host" pointers are ordinary host memory read during the call.
 *   - stream arguments are a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every function returns WWF_OK (0) or a negative wwf_status; the message for the
 *     calling thread's last failure is wwf_last_error().  No C++ exception crosses the ABI.
 *     Asynchronous CUDA faults surface at the caller's next synchronisation.
 *   - a plan is immutable after wwf_bank_register calls finish: wwf_featurize /
 *     wwf_augment / wwf_spec_augment are re-entrant across streams and host threads.
 *   - There is NO CPU implementation behind this ABI: without a CUDA device the library
 *     returns WWF_ERR_CUDA.
 */
#ifndef WWFEAT_H_
#define WWFEAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WWF_VERSION 130 /* 0.1.3: tensor-core DCT epilogue, mel lane schedule, launch cache, wwf_plan_set_option */

typedef enum wwf_status {
  WWF_OK = 0,
  WWF_ERR_INVALID = -1,     /* bad argument / config outside the supported envelope */
  WWF_ERR_UNSUPPORTED = -2, /* valid in the reference, not built here */
  WWF_ERR_CUDA = -3,        /* CUDA runtime error (message has cudaGetErrorString) */
  WWF_ERR_WORKSPACE = -4,   /* caller's workspace is too small / misaligned */
  WWF_ERR_NOMEM = -5
} wwf_status;

enum { WWF_FEAT_LOGMEL = 0, WWF_FEAT_MFCC = 1 };
enum { WWF_OUT_F32 = 0, WWF_OUT_F16 = 1 };
enum { WWF_BANK_NOISE = 0, WWF_BANK_RIR = 1 };
enum { WWF_BANK_F32 = 0, WWF_BANK_I16 = 1 }; /* element type of a device-resident clip bank */
enum { WWF_OPT_FEAT_PATH = 0, WWF_OPT_PDL = 1, WWF_OPT_EPILOGUE_WARP = 2, WWF_OPT_CONV_ORDER = 3 };   /* wwf_plan_set_option */
enum { WWF_PATH_AUTO = 0, WWF_PATH_FUSED = 1, WWF_PATH_FLAT = 2 };

/*
 * Feature configuration = the reference's FeatureExtractor constructor arguments
 * (src/evaluation/evaluator.py:86-94, src/evaluation/inference.py:94-102) and the
 * DataConfig fields that feed them (src/config/defaults.py:12-24), plus the
 * torchaudio defaults the missing module relies on.
 */
typedef struct wwf_config {
  int32_t sample_rate;  /* DataConfig.sample_rate (16000) */
  int32_t n_fft;        /* 256 | 400 | 512 | 1024 | 2048  (validator.py:129 set + BASELINE's 400) */
  int32_t hop_length;   /* 0 < hop < n_fft (validator.py:138) */
  int32_t n_mels;       /* 1..128 */
  int32_t n_mfcc;       /* 1..n_mels; used when feature_type == WWF_FEAT_MFCC */
  int32_t feature_type; /* WWF_FEAT_LOGMEL ('mel') | WWF_FEAT_MFCC ('mfcc') (validator.py:145) */
  int32_t out_dtype;    /* WWF_OUT_F32 | WWF_OUT_F16 */
  int32_t cmvn;         /* 0 = off (reference behaviour), 1 = per-utterance CMVN epilogue */
  float top_db;         /* AmplitudeToDB top_db (80); < 0 disables the floor */
  float f_min;          /* mel f_min (0) */
  float f_max;          /* mel f_max; <= 0 means sample_rate/2 */
  float cmvn_eps;       /* (x-mean)/(std+eps) */
  float mask_value;     /* SpecAugment fill value (0) */
  int32_t n_freq_masks; /* columns of wwf_aug.fmask_* (SpecAugment n_freq_masks), 0..8 */
  int32_t n_time_masks; /* columns of wwf_aug.tmask_* (SpecAugment n_time_masks), 0..8 */
  /* Optional host constants.  NULL = computed by the library (double precision, rounded).
   * The Python shim passes torch-computed float32 arrays so that the constants are
   * bit-identical to torchaudio's (their last ulp moves results by ~1e-4 dB). */
  const float* window;  /* host [n_fft]                       torch.hann_window(n_fft) */
  const float* mel_fb;  /* host [n_fft/2+1][n_mels] row-major melscale_fbanks(...)      */
  const float* dct;     /* host [n_mels][n_mfcc] row-major    create_dct(n_mfcc,n_mels,'ortho') */
} wwf_config;

/*
 * Per-batch augmentation draws, all explicit (BASELINE.json north_star: "RIR/noise
 * indices, SNRs and SpecAugment masks passed in explicitly").  Struct-of-arrays of DEVICE
 * pointers, each nullable (NULL = that augmentation is off for the whole batch).
 * Reference surface: AudioAugmentation(...)(wave) and SpecAugment(...)(spec),
 * tests/test_training_pipeline.py:230-262; kwargs src/ui/panel_training.py:309-318.
 */
typedef struct wwf_aug {
  const int32_t* rir_idx;     /* dev [B]  index into the RIR bank, < 0 = no reverb     */
  const int32_t* noise_idx;   /* dev [B]  index into the noise bank, < 0 = no noise    */
  const int64_t* noise_off;   /* dev [B]  first sample inside the noise clip (wraps)   */
  const float* snr_db;        /* dev [B]  target SNR in dB                             */
  const int32_t* fmask_start; /* dev [B][n_freq_masks]  first masked feature row       */
  const int32_t* fmask_len;   /* dev [B][n_freq_masks]  rows masked (0 = none)         */
  const int32_t* tmask_start; /* dev [B][n_time_masks]  first masked frame             */
  const int32_t* tmask_len;   /* dev [B][n_time_masks]  frames masked (0 = none)       */
  /* Draws of the two waveform-shape augmentations.  wwf_featurize / wwf_augment do NOT read them: the
   * caller runs wwf_time_stretch / wwf_pitch_shift on the batch first (the Python shim does).  They are
   * part of this struct so that wwf_draw_aug can fill every draw of a sample in one launch. */
  const double* stretch_rate; /* dev [B]  speed factor (> 1 = faster), exactly 1.0 = untouched */
  const int32_t* pitch_steps; /* dev [B]  semitones, 0 = untouched                            */
} wwf_aug;

typedef struct wwf_plan wwf_plan; /* opaque */

/* Shape facts of a plan (filled by wwf_plan_info). */
typedef struct wwf_info {
  int32_t n_freq;         /* n_fft/2 + 1 */
  int32_t n_feat;         /* rows of the output: n_mels or n_mfcc */
  int32_t device;         /* CUDA device ordinal */
  int32_t sm_count;       /* SMs of that device */
  int32_t rir_fft_size;   /* real FFT size P of the overlap-save reverb blocks (0 = no RIR bank) */
  int32_t rir_max_len;    /* longest registered RIR */
  int32_t n_rir, n_noise; /* registered bank sizes */
} wwf_info;

/* Library version (WWF_VERSION of the build). */
int wwf_version(void);

/* Message of the calling thread's most recent failure ("" if none). */
const char* wwf_last_error(void);

/*
 * Build the immutable device constants (window, FFT twiddles, sparse mel rows, DCT).
 * Replaces:  FeatureExtractor.__init__(sample_rate, feature_type, n_mels, n_mfcc, n_fft,
 *            hop_length, device)   src/evaluation/evaluator.py:86-94
 */
int wwf_plan_create(const wwf_config* cfg, int device, wwf_plan** out);
void wwf_plan_destroy(wwf_plan* plan);
int wwf_plan_info(const wwf_plan* plan, wwf_info* out);

/* Frames produced for n_samples: n_samples / hop + 1 (src/export/onnx_exporter.py:316-317). */
int wwf_num_frames(const wwf_plan* plan, int n_samples);

/*
 * Register a ragged bank of background-noise clips or room impulse responses that
 * wwf_aug.noise_idx / rir_idx index into.
 *   data    dev  float32, all clips back to back; BORROWED for noise (must outlive the plan's
 *                use), consumed during the call for RIRs (their spectra are plan-owned).
 *   offsets host int64 [count+1], clip i = data[offsets[i] .. offsets[i+1])
 * Replaces: the noise / RIR file lists AudioAugmentation loads from data/raw/{background,rirs}
 *           (README.md:57-58, src/ui/panel_training.py:329).  Synchronises `stream`.
 */
int wwf_bank_register(wwf_plan* plan, int kind, const float* data, const int64_t* offsets,
                      int count, void* stream);

/* Scratch bytes wwf_featurize / wwf_augment want for a (B, N) batch: the reverberated clips when an RIR bank is
 * registered, plus the dB tiles of the large-batch path (several rounds of the GPU's CTA slots; without that
 * part - or with a NULL workspace when no RIR bank is registered - wwf_featurize runs its single-kernel path).
 * May be 0. */
size_t wwf_workspace_bytes(const wwf_plan* plan, int B, int N);

/*
 * The hot path: raw clips -> [RIR reverb] -> [noise @ SNR] -> STFT -> mel -> dB/top_db
 * -> [DCT-II] -> [CMVN] -> [SpecAugment masks] -> features.
 *   wav         dev float32 [B][N], row stride wav_stride elements (>= N)
 *   aug         NULL = no augmentation (FeatureExtractor.__call__ only)
 *   out         dev [B][1][n_feat][T] (float32 or float16 per plan), clip stride out_stride
 *               elements (>= n_feat*T)
 *   workspace   dev, >= wwf_workspace_bytes(plan,B,N), 16-byte aligned.  The reverb part is mandatory when
 *               aug->rir_idx is used; the rest only enables the large-batch kernels (a smaller or NULL
 *               workspace selects the single-kernel path; results are bit-identical between the two, except
 *               that the energy of a dry (un-reverberated) clip that gets noise is summed in a different order)
 * Kernels: conv_kernel (reverb) -> feat_kernel, or for large batches feat_prep_kernel -> feat_frames_kernel ->
 * feat_epilogue_block_kernel; all enqueued on `stream`, no allocation, no synchronisation (graph-capturable).
 * Replaces: WakewordDataset.__getitem__'s  AudioAugmentation(wave) -> FeatureExtractor(wave)
 *           -> SpecAugment(feat) chain (src/evaluation/evaluator.py:125,204;
 *           src/evaluation/inference.py:197), batched.
 */
int wwf_featurize(wwf_plan* plan, const float* wav, int B, int N, int64_t wav_stride,
                  const wwf_aug* aug, void* out, int64_t out_stride, void* workspace,
                  size_t workspace_bytes, void* stream);

/*
 * Time-domain half only: [RIR reverb] -> [noise @ SNR], (B,N) -> (B,N) float32.
 * out_wav may alias wav only when aug->rir_idx is NULL.
 * Replaces: AudioAugmentation.__call__(waveform)  tests/test_training_pipeline.py:239-243.
 */
int wwf_augment(wwf_plan* plan, const float* wav, int B, int N, int64_t wav_stride,
                const wwf_aug* aug, float* out_wav, int64_t out_stride, void* workspace,
                size_t workspace_bytes, void* stream);

/*
 * Device-resident loader, part 1: assemble a batch from a clip bank that lives in HBM.
 *   bank  dev [n_clips][N] float32, or int16 PCM (converted as x / 32768, exact), row stride bank_stride
 *   idx   dev int64 [B] clip numbers;  out dev float32 [B][N], row stride out_stride
 * Replaces: WakewordDataset.__getitem__'s per-sample load for datasets held in memory
 *           (src/ui/panel_training.py:323-349: load_dataset_splits + DataLoader(num_workers=16)).
 */
int wwf_gather_clips(const void* bank, int dtype, int64_t n_clips, int N, int64_t bank_stride, const int64_t* idx,
                     int B, float* out, int64_t out_stride, int device, void* stream);

/*
 * Device-resident loader, part 2: make the augmentation draws of one batch ON the GPU.
 * Counter-based Philox4x32-10; sample i of the call is sample number first_index + i of the run and
 * its draws are a pure function of (seed, first_index + i) - the host can recompute them bit-exactly
 * (tests/helpers.py:philox_draws), so they remain "explicit" for the oracle.  Fills the arrays of
 * `out` (the same wwf_aug the plan's wwf_featurize then consumes; pointers are written through).
 *   apply probabilities and ranges = AugmentationConfig (src/config/defaults.py:73-95);
 *   mask arithmetic = torchaudio mask_along_axis; n_freq_masks / n_time_masks, n_feat come from the plan.
 * Replaces: the random.* / torch.rand calls inside AudioAugmentation.__call__ and SpecAugment.__call__.
 */
typedef struct wwf_draw_config {
  uint64_t seed;
  double rir_prob, noise_prob;             /* AugmentationConfig.rir_prob, background_noise_prob */
  double freq_mask_prob, time_mask_prob;   /* gate all freq / all time masks of a clip */
  float snr_lo, snr_hi;                    /* noise_snr_min / noise_snr_max, dB */
  int32_t freq_mask_param, time_mask_param;
  double stretch_prob, stretch_lo, stretch_hi; /* rate ~ U[time_stretch_min, time_stretch_max) with this probability */
  double pitch_prob;                           /* semitones ~ randint[pitch_shift_min, pitch_shift_max] (inclusive) */
  int32_t pitch_lo, pitch_hi;
} wwf_draw_config;
int wwf_draw_aug(wwf_plan* plan, const wwf_draw_config* cfg, uint64_t first_index, int B, int T,
                 const wwf_aug* out, void* stream);

/*
 * Time-stretch (pitch-preserving speed change), shape kept: out[b] = the clip played rates[b] times faster,
 * cropped / zero-padded back to N samples; rates[b] == 1.0 copies the clip.  Arithmetic: torchaudio's
 * STFT(512, hop 128) -> F.phase_vocoder(rate) -> iSTFT(length = round(N / rate)), i.e. F.pitch_shift's own
 * stretch stage (TA/functional/functional.py:1644-1693, 732-800) with the rate given directly.
 *   rates     dev float64 [B];   rate_lo  host lower bound of every rates[b] (sizes the workspace; >= 0.1)
 *   workspace dev, >= wwf_stretch_workspace_bytes(B, N, rate_lo), 16-byte aligned;  out may alias wav
 * Replaces: the time-stretch branch of AudioAugmentation.__call__ (kwarg time_stretch_range,
 *           tests/test_training_pipeline.py:233; AugmentationConfig.time_stretch_min/max, src/config/defaults.py:76-77).
 */
size_t wwf_stretch_workspace_bytes(int B, int N, double rate_lo);
/* Optional: replace the stretch stage's float32 analysis / synthesis window (host pointer, 512 values, copied)
 * by the caller's, normally torch.hann_window(512).  The library's default is the correctly rounded periodic
 * Hann window; torch's float32 one differs from it in the last 1-3 ulp, which moves a stretched pure tone by
 * 1e-4 relative - the Python shim passes torch's array. */
int wwf_set_stretch_window(wwf_plan* plan, const float* window);
int wwf_time_stretch(wwf_plan* plan, const float* wav, int B, int N, int64_t wav_stride, const double* rates,
                     double rate_lo, float* out, int64_t out_stride, void* workspace, size_t workspace_bytes,
                     void* stream);

/*
 * Pitch shift by an integer number of semitones per clip, shape kept: torchaudio F.pitch_shift
 * (TA/functional/functional.py:1596-1641) = stretch by 2^(-n/12), resample int(sr / rate) -> sr with the
 * windowed-sinc kernel, crop / zero-pad to N.  n_steps[b] == 0 copies the clip.  sr = the plan's sample_rate.
 *   n_steps   dev int32 [B], every value inside [step_lo, step_hi] (host bounds, within [-12, 12])
 *   workspace dev, >= wwf_pitch_workspace_bytes(B, N, step_lo, step_hi), 16-byte aligned;  out may alias wav
 * The first call for a new semitone value builds that ratio's coefficient table (synchronises `stream`).
 * Replaces: the pitch-shift branch of AudioAugmentation.__call__ (kwarg pitch_shift_range, integer semitones:
 *           tests/test_training_pipeline.py:234, src/config/validator.py:289-294, src/config/defaults.py:78-79).
 */
size_t wwf_pitch_workspace_bytes(int B, int N, int step_lo, int step_hi);
int wwf_pitch_shift(wwf_plan* plan, const float* wav, int B, int N, int64_t wav_stride, const int32_t* n_steps,
                    int step_lo, int step_hi, float* out, int64_t out_stride, void* workspace,
                    size_t workspace_bytes, void* stream);

/*
 * Sample-rate conversion of a batch, torchaudio F.resample defaults (sinc_interp_hann, lowpass_filter_width 6,
 * rolloff 0.99; TA/functional/functional.py:1305-1497): in [B][n_in] at orig_freq -> out [B][n_out] at new_freq.
 * wwf_resample_length(n_in, orig, new) = ceil(new * n_in / orig) is the length torchaudio returns; a smaller
 * n_out crops, a larger one is zero-padded.  The first call for a new ratio builds its coefficient table
 * (synchronises `stream`).  in and out must not overlap.
 * Replaces: the resampling step of AudioProcessor in front of the path (8-48 kHz files -> 16 kHz;
 *           src/evaluation/evaluator.py:76-79,119, src/ui/panel_docs.py:134-138).
 */
int wwf_resample_length(int n_in, int orig_freq, int new_freq);
int wwf_resample(wwf_plan* plan, const float* in, int B, int n_in, int64_t in_stride, int orig_freq, int new_freq,
                 float* out, int n_out, int64_t out_stride, void* stream);

/*
 * Per-clip peak normalisation y = x / max|x| (all-zero clips pass through), float32 [B][N] -> [B][N];
 * out may alias wav.  Needs no plan.
 * Replaces: the chunk normalisation in front of the feature extractor,
 *           src/evaluation/inference.py:189-191 (DataConfig.normalize_audio, src/config/defaults.py:24).
 */
int wwf_peak_normalize(const float* wav, int B, int N, int64_t wav_stride, float* out, int64_t out_stride,
                       int device, void* stream);

/*
 * In-place explicit-index SpecAugment on an existing feature tensor [B][F][T]
 * (float32 or float16 per `dtype`): rows [fstart, fstart+flen) and frames
 * [tstart, tstart+tlen) := mask_value.  Needs no plan.
 * Replaces: SpecAugment.__call__(spec)  tests/test_training_pipeline.py:252-262.
 */
int wwf_spec_augment(void* spec, int dtype, int B, int F, int T, int64_t clip_stride,
                     const int32_t* fmask_start, const int32_t* fmask_len, int n_freq_masks,
                     const int32_t* tmask_start, const int32_t* tmask_len, int n_time_masks,
                     float mask_value, int device, void* stream);

/*
 * Non-finite guard: wwf_featurize raises a plan-owned device flag when any feature it computed is
 * NaN/Inf (e.g. an all-zero noise clip makes F.add_noise's scale infinite, exactly like the oracle).
 * This call waits for `stream`, returns the flag (0/1) and clears it.
 * Serves: the reference trainer's "skip non-finite batch" rule, src/training/trainer.py:177-179.
 */
int wwf_check_finite(wwf_plan* plan, void* stream, int* nonfinite);

/*
 * Measurement hook (bench.py's roofline): when enabled, wwf_featurize brackets each of its
 * kernels with CUDA events on the launching stream.  wwf_profile_read waits for them, returns
 * the AVERAGE duration (ms) of the reverb kernel and of the feature kernel per call since the
 * last read plus the number of calls, and resets the counters.  Off by default.
 */
int wwf_profile_enable(wwf_plan* plan, int enable);
int wwf_profile_read(wwf_plan* plan, double* conv_ms, double* feat_ms, int* n_calls);
/* Same per kernel: kernel_ms[4] = { reverb kernel, feat_prep_kernel (per-clip mix records; only when noise is mixed),
 * feat_frames_kernel (flat path) or the fused feat_kernel (single-kernel path), the flat path's epilogue
 * (feat_epilogue_mma_kernel / feat_epilogue_block_kernel) }; n_split = calls that took the flat path (may be NULL).
 * wwf_profile_read's feat_ms is the sum of the last three. */
int wwf_profile_read_kernels(wwf_plan* plan, double* kernel_ms, int* n_calls, int* n_split);

/*
 * Launch options of a plan.  WWF_OPT_FEAT_PATH: WWF_PATH_AUTO (default: the library picks per batch shape),
 * WWF_PATH_FUSED (one kernel per call) or WWF_PATH_FLAT (flat frame queue + epilogue; needs the workspace);
 * WWF_OPT_PDL: 1 (default) chains the kernels of a call with programmatic dependent launch, 0 = plain launches;
 * WWF_OPT_EPILOGUE_WARP: 1 (default) lets MFCC calls without SpecAugment flags of the common shapes take the
 * warp-autonomous tensor-core epilogue (feat_epilogue_mma_warp_kernel), 0 = the block-wise one for every call;
 * WWF_OPT_CONV_ORDER: 1 (default) = when a call has more reverb work items than SMs, the reverb kernel deals the
 * reverberated clips round-robin over its persistent CTAs (it builds the order itself; a call that reverberates 30 %
 * of 1024 clips runs 1.8x faster), 0 = batch order - what a caller that KNOWS every clip of its batches is
 * reverberated may set to save the ~2 us the ordering costs (the Python shim does, per batch, from host-side draws).
 * This option does not drop the cached launch shapes.
 * They can be preset through the environment (WWF_FEAT_PATH=fused|split, WWF_NO_PDL, WWF_NO_EP_WARP,
 * WWF_NO_CONV_ORDER), which is read
 * once, at wwf_plan_create.  Changing an option drops the plan's cached launch shapes; call it between, not
 * concurrently with, wwf_featurize calls.  No reference counterpart: test / measurement control only.
 */
int wwf_plan_set_option(wwf_plan* plan, int option, int value);

/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t wwf_launch_count(void);

/*
 * Test control: fills the shared memory of every SM of `device` with the bit pattern `word` (0x7fc00000 = quiet NaN)
 * and synchronises.  Kernels must never depend on what a predecessor left in shared memory - not even through a
 * product with a zero weight; the GPU tests poison it (and freshly allocated global memory) before parity runs.
 * No reference counterpart.
 */
int wwf_debug_poison_smem(int device, uint32_t word);

#ifdef __cplusplus
}
#endif
#endif /* WWFEAT_H_ */
