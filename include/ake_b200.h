/* ake_b200.h -- C ABI of the B200-native Audio-Key-Estimation hot path.
 *
 * One shared library (libake_b200.so, built from audio_key_estimation_b200/csrc
 * with nvcc for sm_100a) exports exactly the entry points below.  They replace,
 * for the reference repository flo-stilz/Audio-Key-Estimation:
 *
 *   ake_cqt_*      librosa.cqt(...) + abs + log(1+x) + reshape at
 *                  KeyDataset.py:485-509 (same call at equivariance_test.py:155-170)
 *   ake_pcn_*      PitchClassNet.__init__ / forward, models.py:651-817
 *                  (layers models.py:22-399), state_dict contract eval.py:113-115
 *   ake_decode_*   the argmax key / tonic / genre rule, models.py:1083-1085,1096,923
 *   ake_estimate_* the two stages chained for host buffers (the loop eval.py:118-129
 *                  drives through KeyDataset + trainer.validate)
 *
 * Conventions: plain C types only; every pointer named *_dev is a CUDA device
 * pointer owned by the caller (PyTorch allocates them), *_host is host memory;
 * `stream` is a cudaStream_t passed as void*; calls are asynchronous on that
 * stream unless stated; the return value is 0 on success or a negative AKE_ERR_*
 * code, with a thread-local message available from ake_last_error().  Plans
 * hold no mutable global state: different plans may be driven from different
 * host threads / streams.  There is no CPU fallback: every compute entry point
 * launches CUDA kernels or fails.
 */
#ifndef AKE_B200_H
#define AKE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AKE_ABI_VERSION 2

#define AKE_OK 0
#define AKE_ERR_INVALID (-1)      /* bad argument / shape (the reference asserts: models.py:43-44,101,356-357) */
#define AKE_ERR_UNSUPPORTED (-2)  /* option combination outside the hot path */
#define AKE_ERR_CUDA (-3)         /* CUDA runtime error */
#define AKE_ERR_WORKSPACE (-4)    /* caller workspace too small */

typedef struct ake_pcn ake_pcn; /* PitchClassNet plan: architecture + device weights */
typedef struct ake_cqt ake_cqt; /* CQT plan: decimator taps + dense time-domain filter bank */

int ake_abi_version(void);
const char* ake_last_error(void);
/* Number of kernels this library launched from the calling thread since the last reset. */
int64_t ake_launch_count(int reset);

/* Optional device timing of the library's kernel sections (used by bench.py for the roofline line).
 * While enabled, each section ("cqt.decimate", "cqt.bank", "pcn.p2p", "pcn.equiv", "pcn.semitone", ...)
 * records a CUDA event pair on its launch stream.  ake_profile_collect synchronises on those events, sums
 * milliseconds and kernel launches per tag into the caller's arrays (tags_out: cap strings of tag_stride
 * bytes), clears the log and returns the number of tags (negative error). */
int ake_profile_enable(int on);
int ake_profile_collect(char* tags_out, int tag_stride, double* ms_out, int64_t* launches_out, int cap);

/* ---------------------------------------------------------------- PitchClassNet */

/* Mirrors the fields PitchClassNet.__init__/forward read from `opt`
 * (models.py:260-350, 662-742, 766-794) plus the positional ctor arguments
 * (models.py:653).  Booleans are 0/1. */
typedef struct ake_pcn_config {
  int32_t pitches;        /* ctor arg; 36*octaves: 288 (train_model.py:93-94) or 360 (equivariance_test.py:176) */
  int32_t pitch_classes;  /* ctor arg; must be 12 (models.py:171,208 hard-code it) */
  int32_t num_layers;     /* opt.num_layers, default 2 */
  int32_t kernel_size;    /* opt.kernel_size, default 7 */
  int32_t conv_layers;    /* opt.conv_layers, default 3 */
  int32_t n_filters;      /* opt.n_filters, default 4 */
  int32_t head_layers;    /* opt.head_layers, default 2 */
  int32_t time_pool_size; /* opt.time_pool_size, default 2 */
  int32_t genre;          /* opt.genre */
  int32_t max_pool;       /* opt.max_pool */
  /* non-default architectures, all on the generic fp32 CUDA-core path:
   *   resblock    Pitch2Pitch / PitchClass2PitchClass stacks = conv + conv_layers residual blocks (models.py:402-454)
   *   stay_sixth  layers >= 1 work on the semitone rows: no up_sixth / pool_semi (models.py:322-323, 366-367)
   *   p2pc_conv   Pitch2PitchClassConv (dilated conv + BN + LeakyReLU) instead of the octave max pool (models.py:108-133)
   *   pc2p_mem    PitchClass2Pitch_MemoryVariant: up-sampled features added to p, not concatenated (models.py:145-166)
   *   local       sliding-window heads: MaxPool2d((1, frames * loc_window_size - head_layers * (kernel_size - 1)), stride 1)
   *               behind the key / tonic heads, no time pooling, no temporal mean (models.py:349, 720-722, 804-810);
   *               outputs per clip are then 12 x T' (11 x T'' for genre) values, see ake_pcn_local_frames
   *   denseblock  DenseBlock / DenseBlockEquivariant stacks with DenseNet concatenation (models.py:456-648, 266-283); not
   *               together with resblock / pc2p_mem / stay_sixth (the reference's channel plan does not cover those)
   *   only_semitones: AKE_ERR_UNSUPPORTED (the reference cannot run it either: models.py:366-367 feeds 96 rows to pools built
   *               for 32) */
  int32_t resblock, denseblock, stay_sixth, only_semitones, p2pc_conv, pc2p_mem, local;
  int32_t frames, loc_window_size; /* opt.frames (default 5), opt.loc_window_size (default 10): read by `local` only */
} ake_pcn_config;

int ake_pcn_create(const ake_pcn_config* cfg, ake_pcn** out);
void ake_pcn_destroy(ake_pcn* plan);

/* The float tensors of the reference state_dict, in state_dict order, WITHOUT the int64
 * `num_batches_tracked` entries.  ake_pcn_set_params_f32 takes them concatenated in this order. */
int ake_pcn_num_tensors(const ake_pcn* plan);
const char* ake_pcn_tensor_name(const ake_pcn* plan, int i);
int ake_pcn_tensor_shape(const ake_pcn* plan, int i, int64_t shape4[4]); /* returns ndim */
int64_t ake_pcn_param_floats(const ake_pcn* plan);
int ake_pcn_set_params_f32(ake_pcn* plan, const float* flat_dev, int64_t n_floats, void* stream);

/* bn_mode: 0 = eval (running statistics, eval.py:116); 1 = train (batch statistics,
 * equivariance_test.py:178 leaves the model in train mode); 2 = train, and keep every activation in the
 * workspace for ake_pcn_backward_f32 (the training step of train_model.py:122 / models.py:952-961; built for
 * num_layers = 2, head_layers = 2, max_pool off -- AKE_ERR_UNSUPPORTED otherwise). */
size_t ake_pcn_workspace_bytes(const ake_pcn* plan, int B, int T, int bn_mode);

/* forward(mel, seq_length) of models.py:747-817.
 *   mel_dev      (B,1,pitches,T) fp32, contiguous
 *   seq_len_dev  int32[B] valid frames per clip, or NULL (= seq_length None: plain mean over all frames)
 *   key_out_dev  (B,12) sigmoid probabilities; tonic_out_dev (B,12) logits;
 *   genre_out_dev (B,11) logits, required iff cfg.genre
 *   bn_stats_out_dev  optional (train mode): for every BatchNorm site in state_dict order, C batch means
 *                     followed by C biased batch variances (so the caller can update running buffers) */
int ake_pcn_forward_f32(ake_pcn* plan, const float* mel_dev, int B, int T, const int32_t* seq_len_dev,
                        int bn_mode, float* key_out_dev, float* tonic_out_dev, float* genre_out_dev,
                        float* bn_stats_out_dev, void* ws_dev, size_t ws_bytes, void* stream);
int ake_pcn_bn_channels(const ake_pcn* plan); /* total channels over all BN sites */
/* Elements per channel every BatchNorm site normalised over in the LAST train-mode forward (state_dict order of the sites):
 * what the caller needs for the unbiased running-variance update.  Returns the number of sites (negative error). */
int ake_pcn_bn_counts(const ake_pcn* plan, int64_t* counts_out, int cap);
/* opt.local: frames per clip of the key / tonic outputs (T') and of the genre output (T'') for an input of T frames; the
 * forward then writes key_out (B, 12 * T'), tonic_out (B, 12 * T'), genre_out (B, 11 * T''), which the reference views as
 * (B, T', 12) / (B, T'', 11) by a plain reshape (models.py:806-810).  Negative error code if the plan is not `local`. */
int ake_pcn_local_frames(const ake_pcn* plan, int T, int* genre_frames_out);

/* The eval-mode forward with its three outputs written side by side as result rows, plus the argmax decode:
 *   rows_out_dev (B, AKE_ROW_FLOATS = 35) fp32 = [12 key probabilities | 12 tonic logits | 11 genre logits (zeros without a
 *   genre head)] -- the table a data-parallel job all-gathers (one collective) and the one D2H copy of a
 *   serving loop;  ids_out_dev (optional) int32 (3, B) = key signature ids, tonic ids, genre ids (-1 without genre head),
 *   the argmax rule of models.py:1083-1085, 1096, 923.  Same arithmetic as ake_pcn_forward_f32(bn_mode 0) + ake_decode_f32. */
#define AKE_ROW_FLOATS 35
int ake_pcn_forward_rows_f32(ake_pcn* plan, const float* mel_dev, int B, int T, const int32_t* seq_len_dev,
                             float* rows_out_dev, int32_t* ids_out_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* Backward pass of the last bn_mode = 2 forward (what loss.backward() runs through models.py:747-817 in the
 * reference).  d_*_out_dev: gradients of the loss w.r.t. key_out (B,12, AFTER the sigmoid), tonic_out (B,12) and
 * genre_out (B,11); NULL = zero.  grads_out_dev receives ake_pcn_param_floats() floats in the layout of
 * ake_pcn_set_params_f32 (entries of the running-statistics buffers are zero), so a data-parallel job all-reduces it
 * as one bucket.  ws_dev must be the workspace that forward used, untouched since (size: ake_pcn_workspace_bytes with
 * bn_mode 2).  One backward per kept forward. */
int ake_pcn_backward_f32(ake_pcn* plan, const float* d_key_out_dev, const float* d_tonic_out_dev,
                         const float* d_genre_out_dev, float* grads_out_dev, int64_t n_floats, void* ws_dev,
                         size_t ws_bytes, void* stream);

/* The training objective of models.py:855-896 (global key estimation) and its gradient w.r.t. the network outputs:
 *   loss = key_weight * BCELoss(key_out, key_labels) + tonic_weight * CrossEntropy(tonic_out, tonic_idx)
 *        + genre_weight * CrossEntropy(genre_out[mask], genre_idx[mask]),  mask = genre_idx >= 0 (models.py:836-840).
 * loss_out_dev[4] = {total, bce, tonic, genre}.  genre_out_dev / genre_idx_dev / d_genre_out_dev may be NULL. */
int ake_loss_f32(const float* key_out_dev, const float* tonic_out_dev, const float* genre_out_dev,
                 const float* key_labels_dev, const int32_t* tonic_idx_dev, const int32_t* genre_idx_dev, int B,
                 float key_weight, float tonic_weight, float genre_weight, float* loss_out_dev, float* d_key_out_dev,
                 float* d_tonic_out_dev, float* d_genre_out_dev, void* stream);
int ake_pcn_get_config(const ake_pcn* plan, ake_pcn_config* out);

/* Debug/parity taps: copy a named intermediate of the LAST forward out of the workspace.
 * Names follow oracle/pcn_port.py taps ("l0.semi", "l1.p2p2", "pc_final", "key_frames", ...).
 * Returns the number of floats (negative error); `out_dev` may be NULL to query the size. */
int64_t ake_pcn_get_tap(const ake_pcn* plan, const char* name, float* out_dev, int64_t cap, void* stream);

/* argmax key signature (cosine similarity against the 21x12 table of utils/key_signatures.py:19-42),
 * argmax tonic, argmax genre (genre_out_dev / genre_id_dev may be NULL). */
int ake_decode_f32(const float* key_out_dev, const float* tonic_out_dev, const float* genre_out_dev, int B,
                   int32_t* key_id_dev, int32_t* tonic_id_dev, int32_t* genre_id_dev, void* stream);

/* Fused optimizer step of the reference's training loop (models.py:1017-1027: torch.optim.Adam(betas, lr, weight_decay=reg),
 * the ExponentialLR factor folded into `lr` by the caller, accumulate_grad_batches folded into grad_scale): ONE launch over
 * the flat gradient bucket ake_pcn_backward_f32 fills.  m/v are flat moment buffers in the same layout (zero them once);
 * param_ptrs_dev[n_tensors] is a device table of the parameter storages in flat-buffer order (NULL: no parameter there,
 * e.g. BatchNorm running statistics), offsets_dev[n_tensors + 1] their offsets into the flat buffers; step counts from 1. */
int ake_adam_step_f32(const float* flat_grads_dev, float* m_flat_dev, float* v_flat_dev, float* const* param_ptrs_dev,
                      const int64_t* offsets_dev, int n_tensors, int64_t total, float lr, float beta1, float beta2, float eps,
                      float weight_decay, float grad_scale, int step, void* stream);

/* MIREX-weighted key score of models.py:1065-1116 (mirex_score) on the device, one thread per clip: the reference's
 * per-sample Python loop (cosine argmax against the 21-row table, 12-bit key comparison, tonic argmax, |signature id
 * difference| == 1 -> "fifth") with its per-sample .cuda() upload of the table.  ACCUMULATES into counters_dev[9] =
 * {samples, correct, fifths, relative, parallel, other, all-12-keys-right ("accuracy"), correct tonics, key bits right}
 * (zero them first; sum them over ranks with one all-reduce); mirex = (1.0 correct + 0.5 fifths + 0.3 relative +
 * 0.2 parallel) / samples.  key_signature_id_dev is (B, sig_width): the data layer delivers a 24-wide one-hot
 * (KeyDataset.py:366, 447) and the reference takes torch.argmax over whatever width arrives.  similarity_out_dev (B, cos(key_out, key_label), models.py:1094) and category_out_dev
 * (B, 0 correct / 1 fifth / 2 relative / 3 parallel / 4 other) may be NULL. */
int ake_mirex_f32(const float* key_out_dev, const float* tonic_out_dev, const float* key_labels_dev,
                  const float* tonic_labels_dev, const float* key_signature_id_dev, int sig_width, int B,
                  uint64_t* counters_dev, float* similarity_out_dev, int32_t* category_out_dev, void* stream);

/* ---------------------------------------------------------------- constant-Q front-end */

/* librosa.cqt(y, sr, hop_length, fmin, n_bins, bins_per_octave, filter_scale, sparsity) with the other
 * arguments at their librosa-0.9.2 defaults (tuning 0, norm 1, hann, scale True, pad_mode 'constant',
 * res_type None -> kaiser_fast).  fmin <= 0 selects C1.  Fails (AKE_ERR_INVALID) where librosa 0.9.2
 * raises ParameterError (hop not a multiple of 2^(n_octaves-1), top filter above Nyquist) and
 * (AKE_ERR_UNSUPPORTED) where its recursion would early-downsample or switch resampler. */
int ake_cqt_create(double sr, int hop_length, int n_bins, int bins_per_octave, double fmin, double filter_scale,
                   double sparsity, ake_cqt** out);
/* The same with the octave recursion selectable:
 *   AKE_CQT_RECURSION_092              librosa 0.9.2 (the release the reference pins, requirements.txt:250): every octave halves
 *                                      the rate, so hop_length must be a multiple of 2^(n_octaves-1) -- the reference's own
 *                                      hop = round(rate / 5) (KeyDataset.py:485) passes at 48 kHz (9600) and raises at 44.1 kHz
 *                                      (8820) and 22.05 kHz (4410);
 *   AKE_CQT_RECURSION_HALVE_WHILE_EVEN the rule later librosa releases use (vqt: `if my_hop % 2 == 0` halve): once the hop is
 *                                      odd the rate stays and the filters of the lower octaves double in length.  Built on the
 *                                      0.9.2 filter design and the kaiser_fast resampler, so it is a restated VARIANT (checked
 *                                      against oracle/cqt_port.py's restatement of the same rule), not a bit-level statement of
 *                                      any librosa release: those releases also changed the filter bandwidth and the default
 *                                      resampler (soxr_hq). */
#define AKE_CQT_RECURSION_092 0
#define AKE_CQT_RECURSION_HALVE_WHILE_EVEN 1
int ake_cqt_create_ex(double sr, int hop_length, int n_bins, int bins_per_octave, double fmin, double filter_scale,
                      double sparsity, int recursion, ake_cqt** out);
/* Amplitude contract of the plan.  The kernels carry samples as fp16 (hi, lo) pairs and pre-scale each clip by an exact power
 * of two so that they see |x| <= 1 (results carry no scaling error; librosa.cqt accepts any amplitude).
 *   peak > 0: the caller guarantees |sample| <= peak (default 1: torchaudio.load normalises, KeyDataset.py:478-481);
 *             samples beyond ~256 x peak overflow the fp16 operands (inf / NaN in the result);
 *   peak = 0: unknown -- one extra pass over the audio measures max |sample| per clip (costs ~25 % of the front-end's time). */
int ake_cqt_set_peak(ake_cqt* plan, float peak);
void ake_cqt_destroy(ake_cqt* plan);
int ake_cqt_n_fft(const ake_cqt* plan);
int ake_cqt_n_bins(const ake_cqt* plan);
int ake_cqt_frames(const ake_cqt* plan, int64_t n_samples); /* frames librosa returns for a clip of that length */
/* The dense real time-domain bank the kernels contract with: (2*bins_per_octave, n_fft) fp32, row 2k = Re, 2k+1 = Im. */
int ake_cqt_get_bank(const ake_cqt* plan, float* bank_host, int64_t cap_floats);
int ake_cqt_get_decimator(const ake_cqt* plan, float* taps_host, int cap); /* returns #taps of the half filter (32) */
size_t ake_cqt_workspace_bytes(const ake_cqt* plan, int B, int64_t n_max);

#define AKE_CQT_LOGMAG 0  /* log(1+|C|): what KeyDataset.py:497-499 feeds the network */
#define AKE_CQT_COMPLEX 1 /* raw complex64 C (interleaved re,im): what librosa.cqt returns */

/* audio_dev: B clips, clip b at audio_dev + b*stride, lengths_host[b] valid samples (NULL: all n_max).
 * out_dev: mode LOGMAG (B,1,n_bins,T_max) fp32 zero-padded beyond each clip's frames (KeyDataset.py:242-254);
 *          mode COMPLEX (B,n_bins,T_max,2) fp32.
 * seq_len_out_dev: optional int32[B] frames per clip (the `seq_length` batch field, KeyDataset.py:254). */
int ake_cqt_run_f32(ake_cqt* plan, const float* audio_dev, int64_t stride, const int64_t* lengths_host, int B,
                    int64_t n_max, int mode, float* out_dev, int T_max, int32_t* seq_len_out_dev, void* ws_dev,
                    size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- host-buffer pipeline */

size_t ake_estimate_workspace_bytes(const ake_cqt* cqt, const ake_pcn* pcn, int B, int64_t n_max);
/* Host audio in -> host predictions out: H2D copy, CQT, forward (eval mode), decode, D2H copy, stream sync.
 * audio_host should be page-locked for full PCIe speed.  Outputs: key (B,12), tonic (B,12), genre (B,11 or NULL),
 * ids int32 (3,B) = key signature ids, tonic ids, genre ids (-1 without genre head); any output may be NULL
 * (genre_out_host must be NULL when the plan has no genre head). */
int ake_estimate_host_f32(ake_cqt* cqt, ake_pcn* pcn, const float* audio_host, int64_t stride,
                          const int64_t* lengths_host, int B, int64_t n_max, float* key_out_host,
                          float* tonic_out_host, float* genre_out_host, int32_t* ids_host, void* ws_dev,
                          size_t ws_bytes, void* stream);

/* The same call for 16-bit PCM host audio (what the reference's .wav files hold): torchaudio.load normalises int16 samples
 * as int16 / 32768 (KeyDataset.py:478-481), which is exact in fp32, so the device-side conversion makes this entry point
 * bit-identical to ake_estimate_host_f32 on the normalised samples while moving half the bytes across PCIe (2.88 MB per
 * standard clip instead of 5.76 MB).  stride / n_max / lengths count samples. */
size_t ake_estimate_workspace_bytes_i16(const ake_cqt* cqt, const ake_pcn* pcn, int B, int64_t n_max);
int ake_estimate_host_i16(ake_cqt* cqt, ake_pcn* pcn, const int16_t* pcm_host, int64_t stride, const int64_t* lengths_host,
                          int B, int64_t n_max, float* key_out_host, float* tonic_out_host, float* genre_out_host,
                          int32_t* ids_host, void* ws_dev, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AKE_B200_H */
