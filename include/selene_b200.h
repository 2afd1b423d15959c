/* selene_b200 — C ABI of the B200-native (sm_100a) sequence -> prediction hot path of Selene.
 *
 * The reference (selene-sdk 0.6.0) has no FFI layer of its own: its only native code is two
 * Cython modules and its GPU work is whatever torch dispatches.  The entry points below are what
 * a maintainer would bind (ctypes) to replace, for this path only:
 *
 *   sb_encode_* / sb_genome_*   <- selene_sdk/sequences/_sequence.pyx:11-25
 *                                  (_fast_sequence_to_encoding) reached through
 *                                  selene_sdk/sequences/genome.py:96-166,436-542
 *   sb_intervals_* / sb_label_* <- selene_sdk/targets/_genomic_features.pyx:12-41
 *                                  (_fast_get_feature_data) reached through
 *                                  selene_sdk/targets/genomic_features.py:104-141,377-418
 *   sb_ism_*                    <- selene_sdk/predict/_in_silico_mutagenesis.py:8-143 and the batch
 *                                  assembly loop selene_sdk/predict/model_predict.py:599-660
 *   sb_snv_*                    <- selene_sdk/predict/_variant_effect_prediction.py:137-245 and the
 *                                  per-variant loop selene_sdk/predict/model_predict.py:1054-1124
 *   sb_net_*                    <- selene_sdk/predict/_common.py:65-97 (predict) running
 *                                  models/deepsea.py:21-58 (and the other conv stacks)
 *   score outputs of sb_net_*   <- predict_handlers/absolute_diff_score_handler.py:115,
 *                                  diff_score_handler.py:114, logit_score_handler.py:118-124
 *
 * Conventions
 *   - every function returns SB_OK (0) or an SB_ERR_* code and never throws; sb_last_error()
 *     returns the message of the calling thread's last failure;
 *   - "dev" pointers are device memory on the handle's GPU, "host" pointers are ordinary memory;
 *     the library never frees or retains caller buffers beyond the call (handles copy what they keep);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it asynchronously;
 *   - handles are thread-compatible (one thread at a time per handle), one per GPU;
 *   - base codes are canonical: A=0 C=1 G=2 T=3, 4 = unknown (any character outside ACGTacgt).
 *     `perm[b]` gives the one-hot channel of canonical base b (Genome.BASES_ARR is mutable class
 *     state in the reference, model_predict.py:176-183, so the order is a parameter, never baked in).
 *   - there is NO CPU fallback anywhere in this library.
 */
#ifndef SELENE_B200_H
#define SELENE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_ERR_ARG 1
#define SB_ERR_CUDA 2
#define SB_ERR_UNSUPPORTED 3
#define SB_ERR_NOMEM 4

#define SB_F32 0
#define SB_BF16 1
#define SB_U8 2
#define SB_I64 3
#define SB_F64 4

/* flags written per window by the encode entry points */
#define SB_FLAG_HAS_UPPER_N 1 /* 'N' in sequence  (genome.py:542: uppercase N only)        */
#define SB_FLAG_HAS_UNK 2     /* any character outside ACGTacgt (rows encoded as 0.25 x 4) */

typedef struct sb_genome sb_genome;
typedef struct sb_intervals sb_intervals;
typedef struct sb_net sb_net;

/* ------------------------------------------------------------------------------------------ */
/* misc                                                                                        */
/* ------------------------------------------------------------------------------------------ */
int sb_version(void);
int sb_last_error(char* buf, int n);
/* number of kernels this library has launched in this process so far */
unsigned long long sb_launch_count(void);
int sb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------ */
/* genome image: 2 bits/base + unknown mask + uppercase-N mask, all chromosomes concatenated   */
/* ------------------------------------------------------------------------------------------ */
/* host inputs: base i of the image has code (packed2bit[i>>2] >> 2*(i&3)) & 3, unknown bit
 * (unk_mask[i>>3] >> (i&7)) & 1 and uppercase-N bit upn_mask likewise.  chrom_off[c] is the image
 * offset (in bases) of chromosome c, chrom_len[c] its length; n_image_bases is the image size. */
int sb_genome_create(const uint8_t* packed2bit, const uint8_t* unk_mask, const uint8_t* upn_mask,
                     int64_t n_image_bases, const int64_t* chrom_off, const int64_t* chrom_len,
                     int n_chrom, int device, sb_genome** out);
int sb_genome_destroy(sb_genome* g);

/* windows [start, start+L) of chromosome chrom[i]; parts outside [0, chrom_len) read as 'N'
 * (genome.py:156-166, pad=True).  strand_rc[i] != 0 reverse-complements the in-bounds part only,
 * leaving the padding where it is — that is what genome.py:164-166 does.  Coordinate validity
 * (genome.py:51-92) is decided by the host before calling.  All array arguments are dev. */
int sb_genome_codes(const sb_genome* g, const int32_t* chrom, const int64_t* start,
                    const uint8_t* strand_rc /* may be NULL */, int n, int L,
                    uint8_t* codes_out /* (n,L) */, uint8_t* flags_out /* (n) or NULL */,
                    void* stream);

/* kernel (a): fused 2-bit decode -> one-hot (n, L, 4) of dtype SB_F32 or SB_BF16; unknown rows are
 * 0.25 x 4 (_sequence.pyx:17,21-24). */
int sb_encode_windows(const sb_genome* g, const int32_t* chrom, const int64_t* start,
                      const uint8_t* strand_rc, int n, int L, const uint8_t perm[4], int out_dtype,
                      void* out, uint8_t* flags_out, void* stream);

/* one-hot of code rows already on the device (sequence_to_encoding of arbitrary strings: the host
 * maps characters to codes). */
int sb_encode_codes(const uint8_t* codes /* dev (n,L) */, int64_t n, int L, const uint8_t perm[4],
                    int out_dtype, void* out /* dev (n,L,4) */, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* in-silico mutagenesis (single-base)                                                         */
/* ------------------------------------------------------------------------------------------ */
/* Enumerate the mutations of positions [pos0,pos1): ascending position, alternatives in channel
 * order (perm) skipping the reference base; an unknown position yields all four
 * (_in_silico_mutagenesis.py:91-107).  mut_alt holds canonical codes.  n_mut (dev, 1 int) receives
 * the count; at most `cap` entries are written. */
int sb_ism_enumerate(const uint8_t* codes /* dev (L) */, int L, int pos0, int pos1,
                     const uint8_t perm[4], int32_t* mut_pos /* dev */, uint8_t* mut_alt /* dev */,
                     int32_t* n_mut /* dev */, int cap, void* stream);

/* mutate_sequence (_in_silico_mutagenesis.py:138-143) for a whole batch: row m of `out` is the
 * one-hot of codes[src[m]] (src NULL: row 0) with position mut_pos[m] replaced by mut_alt[m]. */
int sb_ism_expand(const uint8_t* codes /* dev (n_src,L) */, int L, const int32_t* src /* dev/NULL */,
                  const int32_t* mut_pos, const uint8_t* mut_alt, int64_t n_mut,
                  const uint8_t perm[4], int out_dtype, void* out /* dev (n_mut,L,4) */,
                  void* stream);

/* ------------------------------------------------------------------------------------------ */
/* SNV ref/alt windows (variant effect prediction)                                             */
/* ------------------------------------------------------------------------------------------ */
/* For variant i with single-base ref/alt codes: `start[i]` is the 0-based window start the host
 * derives as model_predict.py:1057-1059 (center - start_radius), `var_row` the row of the variant,
 * _get_ref_idxs(L,1) (_variant_effect_prediction.py:137-143; 499 for L=1000).  ref_out has the VCF
 * ref base forced at that row (_handle_standard_ref, :226-245), alt_out the alt base (_process_alt
 * substitution, :195-200, built from the UNMODIFIED window); ref_match[i] = (genome row == VCF
 * ref row).  strand_rc applies get_reverse_complement_encoding (_common.py:37-62) to both
 * (model_predict.py:1102-1110).  ref_out / alt_out / codes_out / ref_match / flags_out may be NULL.
 * codes_out receives the (n,L) UNSUBSTITUTED window codes with the strand already applied (reverse
 * complemented rows for strand_rc != 0), ready for sb_net_forward_* with the substitution given as
 * (sub_pos = var_row, or L-1-var_row on the minus strand; sub_code = base, complemented likewise). */
int sb_snv_pairs(const sb_genome* g, const int32_t* chrom, const int64_t* start,
                 const uint8_t* ref_code, const uint8_t* alt_code, const uint8_t* strand_rc, int n,
                 int L, int var_row, const uint8_t perm[4], int out_dtype, void* ref_out,
                 void* alt_out, uint8_t* codes_out, uint8_t* ref_match, uint8_t* flags_out,
                 void* stream);

/* ------------------------------------------------------------------------------------------ */
/* kernel (b): interval-overlap labels                                                         */
/* ------------------------------------------------------------------------------------------ */
/* host inputs: n intervals (chrom index, start, end, feature index), any order; the handle sorts
 * them by (feature, chrom, start) and keeps a running max of `end` for exact overlap search. */
int sb_intervals_create(const int32_t* chrom, const int32_t* start, const int32_t* end,
                        const int32_t* feat, int64_t n, int n_features, int n_chrom, int device,
                        sb_intervals** out);
int sb_intervals_destroy(sb_intervals* iv);

/* labels for n query bins [bin_start, bin_end) on chrom: per feature the UNION coverage of the
 * clipped intervals (a zero-length clipped interval counts one base, _genomic_features.pyx:32-37)
 * compared with the integer threshold thr_i[f] (computed by the host exactly as :39-40):
 * label = coverage > thr_i[f].  thr_i < 0 selects the no-threshold mode of
 * genomic_features.py:406-415 (any overlapping row).  out (n, n_features) of out_dtype
 * SB_U8 / SB_F32 / SB_I64 / SB_F64.  Array arguments are dev. */
int sb_label_bins(const sb_intervals* iv, const int32_t* chrom, const int32_t* bin_start,
                  const int32_t* bin_end, int n, const int32_t* thr_i /* dev (F) */, int out_dtype,
                  void* out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* kernels (c)+(d): conv1d stacks and fused scoring                                            */
/* ------------------------------------------------------------------------------------------ */
#define SB_LAYER_CONV 1  /* Conv1d(cin,cout,k) valid, + bias, optional ReLU, optional MaxPool(pool) */
#define SB_LAYER_FC 2    /* Linear(cin,cout) + bias, optional ReLU                                  */
#define SB_LAYER_BILSTM 3 /* bidirectional single-layer LSTM(cin -> cout hidden); output 2*cout channels
                           * [forward | backward] per position, flattened position-major for a following
                           * FC.  Weights in torch layout, gate order i,f,g,o: weight = weight_ih_l0
                           * (4*cout, cin), bias = bias_ih_l0, aux[0] = weight_hh_l0 (4*cout, cout),
                           * aux[1] = bias_hh_l0, aux[2..5] = the same four tensors of the reverse
                           * direction.  THE RECURRENCE RUNS OVER THE ROWS OF A BATCH, independently for
                           * every position — that is what models/danQ.py:42-52 computes (it hands a
                           * (positions, batch, channels) tensor to a batch_first LSTM), so
                           * results depend on how rows are grouped: see sb_net_set_recurrence. */

typedef struct {
    int32_t type;          /* SB_LAYER_*                                                           */
    int32_t cin, cout, k;  /* k = 1 for FC                                                          */
    int32_t relu;          /* apply ReLU after bias                                                 */
    int32_t pool;          /* MaxPool1d(kernel=stride=pool) after ReLU; 0/1 = none                  */
    const float* weight;   /* host; torch layout: conv (cout,cin,k), fc (cout,cin)                  */
    const float* bias;     /* host (cout)                                                           */
    const float* aux[6];   /* host; SB_LAYER_BILSTM only (see above), else ignored                  */
} sb_layer;

/* A network = conv layers on a (B, 4, L) one-hot input, flatten (channel-major, torch .view order:
 * models/deepsea.py:56), FC layers, final sigmoid.  BatchNorm (eval) must be folded by the host.
 * perm: channel of canonical base b in the first conv's input. */
int sb_net_create(const sb_layer* layers, int n_layers, int L, const uint8_t perm[4], int device,
                  sb_net** out);
int sb_net_destroy(sb_net* net);
int sb_net_n_outputs(const sb_net* net);
/* FLOPs (2 x MACs) of one forward of one sequence */
double sb_net_flops_per_seq(const sb_net* net);

/* Row grouping for SB_LAYER_BILSTM networks: rows are taken in consecutive blocks of
 * group*interleave rows; inside a block, chain c (0 <= c < interleave) is rows c, c+interleave, ... and the
 * LSTM state runs along each chain (forward direction first row to last, backward the reverse), starting
 * from zero.  group = the reference caller's batch_size; interleave = 2 for interleaved ref/alt rows (the
 * reference predicts the ref batch and the alt batch separately), else 1.  Default: group 64, interleave 1. */
int sb_net_set_recurrence(sb_net* net, int group, int interleave);

/* NonStrandSpecific wrapper (utils/non_strand_specific_module.py:62-85; every manuscript config uses
 * `non_strand_specific: mean`): mode 1 (mean) / 2 (max) makes every forward entry point also run the rows with the
 * position axis and the channel axis reversed (channel c -> 3-c, which is the reverse complement for the ACGT
 * order) and combine the two sigmoid outputs before any score is computed; 0 (default) turns it off.  Set it
 * before asking for sb_net_workspace_bytes.  Doubles the FLOPs per prediction. */
int sb_net_set_strand_mode(sb_net* net, int mode);

/* Per-layer device timing for roofline reports: when enabled, every layer launch is bracketed by
 * CUDA events on the launching stream; sb_net_layer_time sums them (synchronising on the events)
 * since the last sb_net_set_timing call. */
int sb_net_set_timing(sb_net* net, int enable);
int sb_net_n_layers(const sb_net* net);
double sb_net_layer_flops_per_seq(const sb_net* net, int layer);
int sb_net_layer_time(sb_net* net, int layer, double* total_ms, int* launches);

#define SB_PREC_FP32 0 /* CUDA-core FFMA path, fp32 end to end (<= 1e-5 abs vs torch CPU)          */
#define SB_PREC_BF16 1 /* tcgen05 path: bf16 operands, fp32 accumulate in TMEM (<= 2e-2 abs)        */

size_t sb_net_workspace_bytes(const sb_net* net, int64_t n_rows, int precision);

/* Input rows are described without materialising them: row i is codes[src[i]] (src NULL: i) with
 * position sub_pos[i] replaced by canonical code sub_code[i] when sub_pos != NULL and >= 0.
 * pred (n, F) fp32 receives sigmoid outputs. */
int sb_net_forward_codes(sb_net* net, const uint8_t* codes /* dev (n_src,L) */,
                         const int32_t* src, const int32_t* sub_pos, const uint8_t* sub_code,
                         int64_t n, int precision, void* workspace, size_t workspace_bytes,
                         float* pred /* dev (n,F) */, void* stream);

/* Generic input: x (n, L, 4) fp32, any values (predict(), _common.py:65-97, after its transpose). */
int sb_net_forward_onehot(sb_net* net, const float* x, int64_t n, int precision, void* workspace,
                          size_t workspace_bytes, float* pred, void* stream);

/* Fused scoring epilogue.  Rows as in sb_net_forward_codes.  pair_mode:
 *   SB_PAIR_INTERLEAVED: n is even, row 2v is the reference and row 2v+1 the alternative of
 *       variant v; score outputs are (n/2, F).
 *   SB_PAIR_BASELINE: every row is an alternative; base_pred (dev, (n_base,F) fp32 sigmoid outputs,
 *       already clamped if logits were requested before — see below) holds baselines and
 *       base_index[i] (dev, NULL: 0) selects row i's baseline; score outputs are (n, F).
 * Any output pointer may be NULL.  abs_diff = |ref - alt| , diff = alt - ref, logit =
 * logit(alt') - logit(ref') with p' = (p==0 ? 1e-24 : p>=1 ? 0.999999 : p) evaluated in fp32 as
 * log(p/(1-p)) (logit_score_handler.py:118-124).  pred_ref / pred_alt receive the sigmoid outputs;
 * when `logit` is requested they receive the CLAMPED values, as the reference's in-place clamp
 * makes later handlers see. */
#define SB_PAIR_INTERLEAVED 0
#define SB_PAIR_BASELINE 1
int sb_net_forward_score(sb_net* net, const uint8_t* codes, const int32_t* src,
                         const int32_t* sub_pos, const uint8_t* sub_code, int64_t n, int precision,
                         int pair_mode, const float* base_pred, const int32_t* base_index,
                         void* workspace, size_t workspace_bytes, float* abs_diff, float* diff,
                         float* logit, float* pred_ref, float* pred_alt, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* training step (TrainModel.train, train_model.py:483-496): fp32                              */
/* ------------------------------------------------------------------------------------------ */
/* Allocates momentum buffers and gradient bookkeeping for the handle (conv / FC stacks). */
int sb_net_train_init(sb_net* net);
/* floats in the flat gradient vector (per layer: W as (K x ld) then bias (ld)) */
size_t sb_net_n_params(const sb_net* net);
/* offset (floats) of `layer`'s block (W then bias) in the flat gradient vector; layers are laid out in forward order */
size_t sb_net_param_offset(const sb_net* net, int layer);
/* cuda_event (a cudaEvent_t, or NULL to clear) is recorded on the launching stream inside
 * sb_net_forward_backward[_tc] as soon as the gradients of `layer` are complete.  Backward runs from the last layer
 * to the first, so at that point the tail of the flat vector from sb_net_param_offset(layer) on is final: a caller can
 * all-reduce that tail (fc1 = 89 % of DeepSEA's parameters) on another stream while the conv gradients are computed. */
int sb_net_set_grad_event(sb_net* net, int layer, void* cuda_event);
/* dropout probability applied to the output of `layer` (after ReLU / pool) in sb_net_forward_backward */
int sb_net_set_dropout(sb_net* net, int layer, float p);
size_t sb_net_train_workspace_bytes(const sb_net* net, int64_t batch);
/* forward in train mode + BCELoss(mean, logs clamped at -100) + backward.  x: dev (n, L, 4) fp32;
 * targets: dev (n, F) fp32; dropout_masks: host array of n_layers dev pointers (already scaled by
 * 1/(1-p)) or NULL entries / NULL to draw masks on the device from dropout_seed; grads: dev,
 * sb_net_n_params floats (overwritten) — all-reduce them across ranks before sb_net_sgd_step for
 * data-parallel training; loss: dev, 1 float. */
int sb_net_forward_backward(sb_net* net, const float* x, const float* targets, int64_t n,
                            const float* const* dropout_masks, unsigned long long dropout_seed,
                            void* workspace, size_t workspace_bytes, float* grads, float* loss, void* stream);
/* torch.optim.SGD(lr, momentum, weight_decay) update of the handle's fp32 weights (models/deepsea.py:66-74) */
int sb_net_sgd_step(sb_net* net, const float* grads, float lr, float momentum, float weight_decay, void* stream);
/* the same update for layers [layer_begin, layer_end) only; `finish` != 0 on the call that completes a step.  Lets a
 * data-parallel trainer update the classifier on its communication stream as soon as that part of the gradient is
 * reduced (after the event of sb_net_set_grad_event, which is recorded once the backward pass no longer reads those
 * layers' weights) while the convolution gradients are still being computed.  grads_bf16 != 0: `grads` is the flat
 * vector as bf16 (same element offsets), i.e. the buffer a bf16 all-reduce just filled, read without widening it. */
int sb_net_sgd_step_layers(sb_net* net, const void* grads, int grads_bf16, float lr, float momentum, float weight_decay,
                           int layer_begin, int layer_end, int finish, void* stream);
/* fp32 parameters (from_grads NULL) or gradients (from_grads = the flat vector) of one layer, to host,
 * in torch layout: conv (cout, cin, k), fc (cout, cin); bias (cout).  Synchronises the device. */
int sb_net_get_layer_params(const sb_net* net, int layer, const float* from_grads, float* w_host, float* b_host);
/* SGD momentum buffers of one layer, torch layout, to / from host: the ``momentum_buffer`` entries of the
 * ``optimizer`` state a Selene checkpoint holds (selene_sdk/train_model.py:430-437 writes it, :296-326 restores it). */
int sb_net_get_layer_momentum(const sb_net* net, int layer, float* w_host, float* b_host);
int sb_net_set_layer_momentum(sb_net* net, int layer, const float* w_host, const float* b_host);

/* The same step with the layer GEMMs (forward, dgrad, wgrad) on tcgen05: bf16 operands, fp32 accumulation,
 * fp32 master weights and gradients; the 4-channel first conv stays on CUDA cores.  Same arguments and
 * gradient layout as the fp32 functions above; use sb_net_sgd_step / sb_net_get_layer_params unchanged. */
int sb_net_train_init_tc(sb_net* net);
size_t sb_net_train_workspace_bytes_tc(const sb_net* net, int64_t batch);
int sb_net_forward_backward_tc(sb_net* net, const float* x, const float* targets, int64_t n,
                               const float* const* dropout_masks, unsigned long long dropout_seed,
                               void* workspace, size_t workspace_bytes, float* grads, float* loss, void* stream);

/* fp32 <-> bf16 casts of a flat device vector: data-parallel training sends its gradients as bf16 (half the bytes
 * of the all-reduce the reference's DataParallel / DDP would do in fp32, train_model.py:469-508). */
int sb_cast_f32_to_bf16(const float* in, int64_t n, void* out_bf16, void* stream);
int sb_cast_bf16_to_f32(const void* in_bf16, int64_t n, float* out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* packed training samples (samplers/dataloader.py:129-147, scripts/sampler_write_to_h5.py:63-89) */
/* ------------------------------------------------------------------------------------------ */
/* sequences: dev uint8 (n, ceil(L_stored/8), 4) = np.packbits(one_hot > 0, axis=1) — bit 7-(p%8) of byte (s, p/8, c);
 * rows with all four bits set (unknown bases) unpack to 1/4 (unpackbits_sequence).  Writes positions
 * [seq_start, seq_start + out_len) of every sample as (n, out_len, 4) fp32 / bf16 (the centre crop of
 * _H5Dataset.__getitem__, dataloader.py:236-246, is seq_start / out_len). */
int sb_unpack_sequences(const uint8_t* packed, int64_t n, int L_stored, int seq_start, int out_len,
                        int out_dtype, void* out, void* stream);
/* targets: dev uint8 (n, ceil(F/8)) = np.packbits(targets > 0, axis=1) -> (n, F) float32 or uint8 (unpackbits_targets) */
int sb_unpack_targets(const uint8_t* packed, int64_t n, int F, int out_dtype, void* out, void* stream);
/* the inverse (what sampler_write_to_h5.py --compress stores): dev fp32 (n, L, 4) and / or (n, F) -> packed bits;
 * either pair of pointers may be NULL */
int sb_pack_samples(const float* sequences, const float* targets, int64_t n, int L, int F,
                    uint8_t* packed_sequences, uint8_t* packed_targets, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* score writers (predict_handlers/handler.py:15-42): "{:.2e}" text of a score matrix           */
/* ------------------------------------------------------------------------------------------ */
/* Row r of x (dev, n_rows x F float32) as text: "{:.2e}".format(value) per value (correctly rounded, ties to
 * even, "nan" / "inf" / "-inf" as Python prints them), joined by tabs and ended by a newline — what
 * write_to_tsv_file puts after the id columns.  Two calls: sb_format_e2_row_bytes gives the bytes of every row;
 * the caller turns them into offsets (exclusive prefix sum, dev int64) and sb_format_e2_rows writes the text
 * at out + row_off[r]. */
int sb_format_e2_row_bytes(const float* x, int64_t n_rows, int F, int64_t* row_bytes, void* stream);
int sb_format_e2_rows(const float* x, int64_t n_rows, int F, const int64_t* row_off, uint8_t* out, void* stream);
/* The same with the id columns spliced in: row r = prefix[prefix_off[r] : prefix_off[r+1]] (dev bytes, e.g.
 * "pos\tref\talt\t") followed by the values; row_off must then be the prefix sum of (prefix length + value bytes),
 * and the whole file body is one buffer. */
int sb_format_e2_rows_prefixed(const float* x, int64_t n_rows, int F, const int64_t* row_off, const uint8_t* prefix,
                               const int64_t* prefix_off, uint8_t* out, void* stream);

/* Score two prediction matrices that already exist (handlers called on their own). */
int sb_score_pairs(const float* ref /* dev (n_ref,F) */, const float* alt /* dev (n,F) */,
                   const int32_t* base_index /* dev (n) or NULL: ref row i (n_ref==n) or 0 */,
                   int64_t n, int64_t n_ref, int F, float* abs_diff, float* diff, float* logit,
                   int clamp_inplace, void* stream);

/* Self-test hook for the tcgen05 GEMM core: C (M,N) fp32 = A (M,K) bf16 . B (N,K)^T bf16, both
 * K-major, arbitrary M, K % 8 == 0; used by tests to pin descriptor encodings on real hardware. */
int sb_tc_gemm_selftest(const void* a_bf16, const void* b_bf16, int M, int N, int K, float* c,
                        void* stream);

/* Self-test hook for the MN-major split-K weight-gradient GEMM: G (N, M) fp32 (zeroed by the caller) +=
 * dY (rows, M)^T . X-view, where row r of the view is the N elements starting at x + r*x_ld (overlapping
 * rows when N > x_ld, as for a multi-tap convolution).  M % 8 == 0, x_ld % 8 == 0. */
int sb_tc_wgrad_selftest(const void* dy_bf16, const void* x_bf16, long long rows, int M, int x_ld, int N,
                         float* g, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SELENE_B200_H */
