/* recur_b200.h — array-of-nets ("synchronic mini-batch") entry points and
 * device control of librecur_b200.so.
 *
 * recur trains on N cloned nets that share weights and gradient arrays and
 * are stepped one after another on one thread (rnn_new_training_set,
 * reference recur-nn-init.c:221-243).  The loops that do the stepping live
 * in the callers; these calls replace each such loop by one launch sequence
 * over all streams, keeping activations, history and errors in HBM:
 *
 *   reference loop                                 replaced by
 *   charmodel-predict.c:293-311 (text-predict)     rnn_batch_char_step / rnn_batch_text_train
 *   gstclassify.c:2201-2239 (classify train)       rnn_batch_opinion + rnn_batch_set_errors
 *                                                  + rnn_batch_calc_deltas + rnn_batch_advance
 *   gstrnnca.c:718-733 (rnnca trainers)            rnn_batch_set_inputs + rnn_batch_opinion
 *                                                  + rnn_batch_set_errors + rnn_batch_calc_deltas
 *   gstrnnca.c:805-830 (rnnca fill_frame)          rnn_cells_rnnca_frame / rnn_cells_rnnca_run
 *                                                  (or rnn_batch_rnnca_frame on cloned nets)
 *   charmodel-multi-predict.c:350-372              rnn_batch_opinion + rnn_batch_get_outputs
 *   mfcc.c:9-94 per channel (gstclassify.c:1984-1995)   rnn_mfcc_extract
 *   rescale.c:240-256 per plane (gstrnnca.c:619-637)    rnn_b200_adaptive_downscale
 *
 * Every function is plain C: pointers and sizes only.  Host arrays passed in
 * are read before the call returns unless stated; nothing here is reentrant
 * (the reference API is not either: SURVEY.md §8b "Threading").
 *
 * Errors follow the reference's conventions (SURVEY.md §8b "Errors"):
 * allocation or CUDA failure prints one line to stderr and abort()s; calls
 * that can fail for user-visible reasons return NULL / -1.
 */
#ifndef RECUR_B200_H
#define RECUR_B200_H 1

#include "recur-nn.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- device / library state ---------------------------------------------- */

/* Number of usable CUDA devices (0 when there is no driver or no GPU).
   Never aborts. */
int rnn_b200_device_count(void);

/* Select the CUDA device used for all subsequently created nets (default 0,
   or $RECUR_B200_DEVICE).  Returns 0, or -1 if the ordinal does not exist. */
int rnn_b200_set_device(int ordinal);

/* Block until all queued device work of this library is complete. */
void rnn_b200_synchronize(void);

/* The CUDA stream (cudaStream_t) the library launches on, as a void*; for
   callers that time with CUDA events or interleave their own work. */
void *rnn_b200_stream(void);

/* Library version string, e.g. "recur-b200 0.1 (sm_100a)". */
const char *rnn_b200_version(void);

/* Count of kernels launched by this library since load (monotonic). */
uint64_t rnn_b200_kernel_launches(void);

/* Name of the kernel that walked the history ring in the most recent BPTT
   call ("k_tc_chain_persistent", "k_walk_resident", "k_tc_nt<CHAIN>",
   "k_walk_single", "k_gemm<CHAIN>"; "" before the first).  Diagnostic: tests
   use it to prove which path they compared with the oracle. */
const char *rnn_b200_last_walk_kernel(void);

/* Choose the matrix engine for batches: 0 = automatic (tensor cores when the
   batch has >= 64 streams and the sizes allow, FP32 FMA otherwise),
   1 = force FP32 FMA kernels, 2 = force tensor-core kernels (aborts if the
   shapes do not allow them).  Returns the previous setting. */
int rnn_b200_set_engine(int engine);

/* Per-kernel-class device timing (for bench.py's roofline): when enabled,
   every launch of the big contractions and of the update kernel is bracketed
   by CUDA events on the library stream.  rnn_b200_profile_read synchronises
   and fills, per class, the summed milliseconds and the launch count since
   enabling; it returns the number of classes.  Classes: see
   rnn_b200_profile_class_name (0 "forward", 1 "bptt_chain", 2 "weight_grad",
   3 "update"). */
void rnn_b200_profile_enable(int on);
int rnn_b200_profile_read(double *ms, uint64_t *launches, int max_classes);
const char *rnn_b200_profile_class_name(int cls);

/* ---- mirrors --------------------------------------------------------------- */

/* Refresh the host mirrors of one net (input_layer/history, hidden_layer,
   output_layer, i/h/o_error, ih_scale, min_error_factor) from the device. */
void rnn_b200_pull(RecurNN *net);

/* Push the host mirrors of one net (history ring, hidden_layer, o_error,
   min_error_factor) to its device slot; needed only after host code has
   rewritten state that the per-net calls do not upload by themselves. */
void rnn_b200_push(RecurNN *net);

/* ---- array-of-nets calls --------------------------------------------------- */

typedef struct RnnBatch RnnBatch;

/* Sums the text-predict loop accumulates per character position
   (charmodel-predict.c:300-303): error += e; entropy += capped_log2f(1-e);
   correct += (winner == next). */
typedef struct RnnBatchCharStats {
  double error;
  double entropy;
  int64_t correct;
  int64_t count;
} RnnBatchCharStats;

/* Bind n nets that share weights (a training set, or any clones of one
   parent) into a batch.  All nets must come from the same rnn_new /
   rnn_clone family and have the same BPTT depth.  Returns NULL (with a line
   on stderr) if they do not.  The nets stay valid and usable one by one. */
RnnBatch *rnn_batch_new(RecurNN **nets, int n_nets);
void rnn_batch_delete(RnnBatch *batch);
int rnn_batch_size(const RnnBatch *batch);

/* rnn_bptt_advance (recur-nn.c:696-704) for every stream. */
void rnn_batch_advance(RnnBatch *batch);

/* Write every stream's real_inputs: `inputs` is n x input_size floats, row
   per stream. */
void rnn_batch_set_inputs(RnnBatch *batch, const float *inputs);

/* one_hot_opinion's input half (charmodel-helpers.h:16-33): zero the inputs
   and set inputs[hot[j]] = 1 for stream j. */
void rnn_batch_set_one_hot(RnnBatch *batch, const u8 *hot);

/* rnn_opinion(net, NULL, presynaptic_noise) (recur-nn.c:83-154) for every
   stream; the outputs stay on the device. */
void rnn_batch_opinion(RnnBatch *batch, float presynaptic_noise);

/* Copy out n x output_size floats (one row per stream); synchronises. */
void rnn_batch_get_outputs(RnnBatch *batch, float *outputs);

/* Copy out n x hidden_size+1 floats: each stream's hidden_layer[0..hidden_size]. */
void rnn_batch_get_hiddens(RnnBatch *batch, float *hiddens);

/* net_error_bptt's second half (charmodel-predict.c:21-26) for every stream:
   o_error = onehot(target) - softmax(output) with badmaths.h's fast_expf and
   clamp; err[j] = o_error[target[j]]; winner[j] = argmax.  err and winner
   may be NULL.  Synchronises only if one of them is given. */
void rnn_batch_softmax_error(RnnBatch *batch, const u8 *target,
    float *err, int32_t *winner);

/* Write every stream's bptt->o_error: n x output_size floats. */
void rnn_batch_set_errors(RnnBatch *batch, const float *o_error);

/* rnn_bptt_calc_deltas(net, j || accumulate, NULL) (recur-nn.c:707-772) for
   streams j = 0..n-1 in one go: with accumulate == 0 the shared ih/ho deltas
   are overwritten by the sum over all streams, otherwise added to. */
void rnn_batch_calc_deltas(RnnBatch *batch, int accumulate);

/* The same with some streams sitting the step out: active[j] == 0 means
   stream j's rnn_bptt_calc_deltas call is skipped altogether, as
   rnn_char_classify_epoch does for characters without a class
   (charmodel-classify.c:124-148); its generation, min_error_factor and
   contribution to the deltas stay untouched; its o_error row on the device
   is consumed (zeroed) by the call, so set the errors again before training
   that stream on them.  active == NULL: all train. */
void rnn_batch_calc_deltas_masked(RnnBatch *batch, int accumulate, const u8 *active);

/* rnn_apply_learning(nets[0], ...) (recur-nn.c:601-678). */
void rnn_batch_apply_learning(RnnBatch *batch, int learning_style, float momentum);

/* One character position of the synchronic text-predict loop
   (charmodel-predict.c:293-311): advance, one-hot forward on cur[j], softmax
   error against next[j], deltas summed over streams, one weight update.
   `stats` (may be NULL) is ADDED to; reading it back synchronises. */
void rnn_batch_char_step(RnnBatch *batch, const u8 *cur, const u8 *next,
    int learning_style, float momentum, RnnBatchCharStats *stats);

/* Keep an encoded text on the device for rnn_batch_text_train. */
void rnn_batch_text_upload(RnnBatch *batch, const u8 *text, int len);

/* `steps` consecutive positions of that loop over the uploaded text, stream j
   reading position i + j * ((len - 1) / n) as rnn_char_epoch does
   (charmodel-predict.c:273,295-298), momentum following
   rnn_calculate_momentum_soft_start.  Nothing crosses PCIe per step.
   Returns the position after the last step (wrapped into [0, len-1)). */
int rnn_batch_text_train(RnnBatch *batch, int start, int steps,
    int learning_style, float momentum, float momentum_soft_start,
    RnnBatchCharStats *stats);

/* Forward-only variant (rnn_opinion steps/sec): steps one-hot forwards per
   stream over the uploaded text, no training. */
int rnn_batch_text_forward(RnnBatch *batch, int start, int steps);

/* One frame of gstrnnca's cellular automaton (reference gstrnnca.c:805-830, fill_frame, with
   fill_net_inputs :670-691 and get_offset_point :644-667): `cells` holds one forward-only net
   per pixel, row-major, width * height of them.  frame_in / frame_out are three planes (Y, Cb,
   Cr) of width * height bytes in host memory.  Every cell reads len_y luma and len_c chroma
   neighbours of frame_in at the (dx, dy) pairs in offsets_y / offsets_c (clamped when `edges`,
   else wrapped once), scaled by 1/255, then x/width, y/height (and with len_pos 3 the radial
   term), runs rnn_opinion, and its three outputs go through fast_sigmoid and UNIT_TO_BYTE into
   frame_out.  Nothing but the two frames crosses PCIe. */
void rnn_batch_rnnca_frame(RnnBatch *cells, const unsigned char *frame_in,
    unsigned char *frame_out, int width, int height, const int *offsets_y, int len_y,
    const int *offsets_c, int len_c, int len_pos, int edges);

/* ---- a net per pixel without a struct per pixel ---------------------------- */

/* gstrnnca runs one forward-only net per pixel, all with the trainers' weights
   (gstrnnca.c:805-830); at 1080p that is two million RecurNN clones whose only
   content is 52 floats of hidden state.  RnnCells keeps that state on the
   device and nothing on the host: `prototype` lends its weights (they are read
   afresh on every call, so training the prototype in between is seen), the
   cells start with zeroed hidden layers like fresh clones.  Nets of up to 63
   hidden units, 40 inputs, at least 3 outputs, no bottom layer; NULL (with a
   line on stderr) otherwise. */
typedef struct RnnCells RnnCells;
RnnCells *rnn_cells_new(RecurNN *prototype, int width, int height);
/* The same frame shared by the GPUs of rnn_b200_comm_join: rank r of n keeps
   and computes rows [r height/n, (r+1) height/n); after every frame the bands
   are gathered, so the frame calls below - now collective - still take and
   return whole pictures on every rank.  height must be a multiple of n. */
RnnCells *rnn_cells_new_sharded(RecurNN *prototype, int width, int height);
void rnn_cells_delete(RnnCells *cells);
/* rnn_forget_history for every cell */
void rnn_cells_forget(RnnCells *cells);

/* rnn_batch_rnnca_frame on such cells: one frame in (host), one frame out. */
void rnn_cells_rnnca_frame(RnnCells *cells, const unsigned char *frame_in,
    unsigned char *frame_out, const int *offsets_y, int len_y, const int *offsets_c, int len_c,
    int len_pos, int edges);

/* The automaton running by itself: n_frames steps, each frame the input of the
   next, the pictures staying on the device.  frame_in may be NULL (go on from
   the last picture), frame_out may be NULL. */
void rnn_cells_rnnca_run(RnnCells *cells, const unsigned char *frame_in, int n_frames,
    unsigned char *frame_out, const int *offsets_y, int len_y, const int *offsets_c, int len_c,
    int len_pos, int edges);

/* One cell's hidden_layer: h_size floats (on a sharded object: a cell of this
   rank's band).  Synchronises. */
void rnn_cells_get_hidden(RnnCells *cells, int cell, float *hidden);

/* ---- gstrnnca's frame intake ------------------------------------------------ */

/* recur_adaptive_downscale (rescale.c:240-256; remember_frame,
   gstrnnca.c:619-637): box-filter a byte plane down to another size, bit for
   bit as the reference does it - every second row and column when shrinking by
   four or more both ways, a plain copy of width * height bytes for equal
   sizes, exact sums otherwise; what the reference's rounding leaves unwritten
   at the right and bottom of `dst` stays as it was.  Host planes (the call
   synchronises) or device planes (queued on the library's stream).  Returns 0,
   or -1 with a line on stderr for enlargements and for shrink factors beyond
   the reference's 16-bit column sums (more than 257 source rows per row). */
int rnn_b200_adaptive_downscale(const unsigned char *src, int s_width, int s_height, int s_stride,
    unsigned char *dst, int d_width, int d_height, int d_stride);
int rnn_b200_adaptive_downscale_device(const unsigned char *src_dev, int s_width, int s_height,
    int s_stride, unsigned char *dst_dev, int d_width, int d_height, int d_stride);

/* ---- the audio front end of gstclassify ------------------------------------- */

/* mfcc.c:9-94 for every channel's window in one launch: window function, real
   FFT, overlapping triangular mel bins, log(1 + power), and optionally the
   DCT (recur_extract_log_freq_bins / recur_extract_mfccs).  rnn_mfcc_new takes
   recur_audio_binner_new's arguments (mfcc.c:308-337; window_type as in mfcc.h:
   0 none, 1 Hann, 2 Vorbis, 3 MP3) and builds the same tables on the host.
   Windows are powers of two from 16 to 4096 samples; NULL (with a line on
   stderr) otherwise. */
typedef struct RnnMfcc RnnMfcc;
RnnMfcc *rnn_mfcc_new(int window_size, int window_type, int n_bins, float min_freq,
    float max_freq, float knee_freq, float focus_freq, float audio_rate, float scale,
    int value_size);
void rnn_mfcc_delete(RnnMfcc *mfcc);
/* n_windows windows of window_size samples in, n_windows rows of n_bins floats
   out (host memory; synchronises).  dct != 0: the bins' DCT. */
void rnn_mfcc_extract(RnnMfcc *mfcc, const float *pcm, int n_windows, float *out, int dct);
/* The same with both ends in device memory, queued on the library's stream. */
void rnn_mfcc_extract_device(RnnMfcc *mfcc, const float *pcm_dev, int n_windows, float *out_dev,
    int dct);
/* The tables rnn_mfcc_new built: window_size mask values and n_bins + 1 slopes. */
void rnn_mfcc_tables(RnnMfcc *mfcc, float *mask, int *left, int *right, float *left_fraction,
    float *right_fraction, float *slope);

/* Number of BPTT steps each stream executed in the most recent
   rnn_batch_calc_deltas / training step (the value the reference logs as
   "depth", plus one when the walk stopped early; recur-nn.c:387,416): n
   int32 values.  Synchronises. */
void rnn_batch_bptt_depths(RnnBatch *batch, int32_t *depths);

/* What bptt_and_accumulate_error and rnn_bptt_calc_deltas write to net->log
   for one stream (recur-nn.c:415-421,766-770), for every stream of the batch
   after its most recent walk: n records.  Synchronises. */
typedef struct RnnBatchBpttLog {
  int32_t depth;              /* "depth": bptt->depth - t at the exit */
  int32_t n_steps;            /* BPTT steps executed (as rnn_batch_bptt_depths) */
  float scaled_error;         /* "scaled_error": ih_scale * error_sum */
  float ih_scale;             /* "ih_scale" */
  float min_error_threshold;  /* "min_error_threshold" */
  float min_error_factor;     /* "min_error_factor" (after adaptation) */
  float cum_error;            /* "cum_error" */
  float error_sum;            /* error_sum of the last executed step */
  float top_error_scaled;     /* "top_error_scaled" */
  float top_error_raw;        /* "top_error_raw" */
} RnnBatchBpttLog;
void rnn_batch_bptt_log(RnnBatch *batch, RnnBatchBpttLog *log);

/* Refresh host mirrors and struct scalars (generation, ih_scale,
   min_error_factor, bptt->index) of every net in the batch. */
void rnn_batch_pull(RnnBatch *batch);

/* ---- multi-GPU: streams shard across processes, deltas are summed -------- */

/* One process per GPU.  The caller obtains a 128-byte NCCL unique id on rank
   0, distributes it by any means (bench.py uses torch.distributed), and every
   rank joins.  After rnn_b200_comm_join, rnn_batch_calc_deltas /
   rnn_batch_char_step / rnn_batch_text_train all-reduce (sum) the
   concatenated [ih_delta | ho_delta] over the ranks before the update, so
   every rank applies the identical update to its weight replica
   (SURVEY.md §8e).  With more than one rank the deltas of a step are
   exchanged whole: accumulate != 0 would add an already exchanged sum into
   the next exchange, so the library aborts on it.
   Returns 0 on success, -1 if NCCL cannot be loaded. */
int rnn_b200_comm_unique_id(void *id128);
int rnn_b200_comm_join(const void *id128, int rank, int n_ranks);
void rnn_b200_comm_leave(void);
int rnn_b200_comm_size(void);

/* Fused gradient exchange over NVLink peer memory (optional, after
   rnn_b200_comm_join).  Instead of "sum the split-K partials, then call
   ncclAllReduce", ONE kernel per rank finishes the weight-gradient GEMM's
   split-K reduction into a peer-visible staging buffer, reduces its 1/N slice
   over all ranks with peer loads, and scatters the result to every rank with
   peer stores (SURVEY.md §5 "natural fusion").  Each rank exports
   RNN_B200_P2P_HANDLE_BYTES of CUDA IPC handles for its batch; the launcher
   gathers them from all ranks (rank order) and every rank attaches.  Returns
   0 on success, -1 when peer access is not possible (NCCL keeps being used). */
#define RNN_B200_P2P_HANDLE_BYTES 192
int rnn_batch_p2p_export(RnnBatch *batch, void *handles_out);
int rnn_batch_p2p_attach(RnnBatch *batch, const void *all_handles, int rank, int n_ranks);
/* Measurement aid (collective: every rank calls it with the same `rounds`,
   after at least one training step on the tensor engine): `rounds` exchanges
   back to back with nothing else on the stream; returns the mean microseconds
   of one (exchange kernel + the consumer's wait and copy-out), -1 without an
   attached exchange.  ih_delta / ho_delta are overwritten with sums of stale
   split-K planes: call it between training steps, not between calc_deltas
   and apply_learning. */
float rnn_batch_p2p_probe(RnnBatch *batch, int rounds);

#ifdef __cplusplus
}
#endif
#endif
