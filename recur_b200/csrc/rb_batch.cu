/* rb_batch.cu — the array-of-nets entry points (include/recur_b200.h).
 *
 * These replace the callers' `for j in streams` loops (SURVEY.md §8b "New
 * batch entry points").  State stays in the device pool between calls; host
 * mirrors of the nets are refreshed only by rnn_batch_pull / rnn_b200_pull
 * (each net is flagged dev_ahead so that a later per-net call pulls first).
 */
#include "rb_kernels.h"
#include "rb_host.h"
#include "rb_comm.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CUDA_OR_DIE(call) do {                                          \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess)                                              \
      rb_die("recur-b200: %s failed at %s:%d: %s", #call, __FILE__,     \
          __LINE__, cudaGetErrorString(e_));                            \
  } while (0)

struct RnnBatch {
  RbNet **nets;
  int n;
  RbPool *pool;
  RbGroup *group;
  int contiguous, base;
  int *slots_dev;        /* only when the slots are not one ascending run */
  /* device scratch */
  u8 *cur_dev, *next_dev;
  float *err_dev;
  int *winner_dev;
  RbCharAccum *accum_dev;
  float *lr_dev, *mef_dev;
  float *io_dev;         /* n x max(input_size, o_size, hidden_size+1) floats */
  size_t io_floats;
  u8 *text_dev;
  int text_len;
  /* pinned staging */
  u8 *sym_host;          /* 2n */
  float *f_host;         /* like io_dev */
  int *i_host;           /* n */
  float *lr_host;        /* n, last uploaded learn rates */
  RbCharAccum *accum_host;
  u8 *rnnca_dev, *rnnca_host; /* rnn_batch_rnnca_frame: offsets, frame in, frame out */
  size_t rnnca_cap;
  int all_marked;       /* every net carries dev_ahead (see mark_ahead) ... */
  unsigned long long marked_epoch; /* ... as of this value of rb_ahead_epoch */
  int accum_reset;      /* the next accumulation starts from zero (sums were fetched) */
  int snapshot_valid;   /* accum_host holds the sums as of the last queued step */
  void *p2p;             /* fused gradient exchange (multi-GPU), or NULL */
  int masked;            /* the last calc_deltas left skip bits on the device */
};

static int g_engine = 0;

extern "C" int
rnn_b200_set_engine(int engine)
{
  int old = g_engine;
  if (engine >= 0 && engine <= 2)
    g_engine = engine;
  return old;
}

extern "C" int
rb_engine(void)
{
  return g_engine;
}

static void
batch_view(RnnBatch *b, RbView *v)
{
  rb_view_of_net(b->nets[0], v);
  v->cap = b->pool->cap; /* pools never grow while a batch exists on them... */
  v->n = b->n;
  v->contiguous = b->contiguous;
  v->base = b->base;
  v->slots = b->contiguous ? b->pool->iota + b->base : b->slots_dev;
  v->p2p = b->p2p;
  rb_bottom_attach(v, &b->nets[0]->pub);
}

static void
mark_ahead(RnnBatch *b)
{
  /* one pass per run of batch calls, not one per call: with a net per pixel
     (rnnca) the loop alone costs tens of milliseconds a frame.  Whoever clears
     any net's flag bumps rb_ahead_epoch (rb_net_pull / rb_net_push), and the
     next batch call flags again. */
  if (b->all_marked && b->marked_epoch == rb_ahead_epoch)
    return;
  for (int j = 0; j < b->n; j++)
    b->nets[j]->dev_ahead = 1;
  b->all_marked = 1;
  b->marked_epoch = rb_ahead_epoch;
}

extern "C" RnnBatch *
rnn_batch_new(RecurNN **nets, int n_nets)
{
  if (!nets || n_nets < 1) {
    fprintf(stderr, "rnn_batch_new: need at least one net\n");
    return NULL;
  }
  rb_require_device("rnn_batch_new");
  RnnBatch *b = (RnnBatch *)calloc(1, sizeof(RnnBatch));
  b->nets = (RbNet **)calloc(n_nets, sizeof(RbNet *));
  b->n = n_nets;
  for (int j = 0; j < n_nets; j++) {
    b->nets[j] = rb_net_of(nets[j]);
    if (b->nets[j]->pool != b->nets[0]->pool ||
        nets[j]->ih_weights != nets[0]->ih_weights) {
      fprintf(stderr, "rnn_batch_new: net %d does not share weights and BPTT depth "
          "with net 0\n", j);
      free(b->nets);
      free(b);
      return NULL;
    }
  }
  /* a previous batch may have left the structs behind the device */
  for (int j = 0; j < n_nets; j++) {
    if (b->nets[j]->dev_ahead)
      rb_net_pull(b->nets[j]);
  }
  b->pool = b->nets[0]->pool;
  b->group = b->nets[0]->group;
  b->base = b->nets[0]->slot;
  b->contiguous = 1;
  for (int j = 0; j < n_nets; j++)
    if (b->nets[j]->slot != b->base + j)
      b->contiguous = 0;
  const RbDims *d = &b->group->d;
  size_t n = (size_t)n_nets;
  if (!b->contiguous) {
    int *slots = (int *)malloc(n * sizeof(int));
    for (int j = 0; j < n_nets; j++)
      slots[j] = b->nets[j]->slot;
    CUDA_OR_DIE(cudaMalloc((void **)&b->slots_dev, n * sizeof(int)));
    CUDA_OR_DIE(cudaMemcpy(b->slots_dev, slots, n * sizeof(int), cudaMemcpyHostToDevice));
    free(slots);
  }
  size_t widest = d->input_size;
  if (nets[0]->bottom_layer && (size_t)nets[0]->bottom_layer->input_size > widest)
    widest = nets[0]->bottom_layer->input_size;
  if ((size_t)d->o_size > widest) widest = d->o_size;
  if ((size_t)d->hidden_size + 1 > widest) widest = d->hidden_size + 1;
  b->io_floats = n * widest;
  /* one allocation, [current symbols | next symbols], so that a step's symbols
     arrive in one copy */
  CUDA_OR_DIE(cudaMalloc((void **)&b->cur_dev, 2 * (size_t)n));
  b->next_dev = b->cur_dev + n;
  CUDA_OR_DIE(cudaMalloc((void **)&b->err_dev, n * sizeof(float)));
  CUDA_OR_DIE(cudaMalloc((void **)&b->winner_dev, n * sizeof(int)));
  CUDA_OR_DIE(cudaMalloc((void **)&b->accum_dev, sizeof(RbCharAccum)));
  CUDA_OR_DIE(cudaMemset(b->accum_dev, 0, sizeof(RbCharAccum)));
  CUDA_OR_DIE(cudaMalloc((void **)&b->lr_dev, n * sizeof(float)));
  CUDA_OR_DIE(cudaMalloc((void **)&b->mef_dev, n * sizeof(float)));
  CUDA_OR_DIE(cudaMalloc((void **)&b->io_dev, b->io_floats * sizeof(float)));
  CUDA_OR_DIE(cudaHostAlloc((void **)&b->sym_host, 2 * n, cudaHostAllocDefault));
  CUDA_OR_DIE(cudaHostAlloc((void **)&b->f_host, b->io_floats * sizeof(float), cudaHostAllocDefault));
  CUDA_OR_DIE(cudaHostAlloc((void **)&b->i_host, n * sizeof(int), cudaHostAllocDefault));
  CUDA_OR_DIE(cudaHostAlloc((void **)&b->lr_host, n * sizeof(float), cudaHostAllocDefault));
  CUDA_OR_DIE(cudaHostAlloc((void **)&b->accum_host, sizeof(RbCharAccum), cudaHostAllocDefault));
  /* per-stream training parameters as the structs have them now */
  if (b->pool->has_bptt) {
    RbView v;
    batch_view(b, &v);
    for (int j = 0; j < n_nets; j++) {
      b->lr_host[j] = nets[j]->bptt->learn_rate;
      b->f_host[j] = nets[j]->bptt->min_error_factor;
    }
    CUDA_OR_DIE(cudaMemcpyAsync(b->lr_dev, b->lr_host, n * sizeof(float),
            cudaMemcpyHostToDevice, rb_stream));
    CUDA_OR_DIE(cudaMemcpyAsync(b->mef_dev, b->f_host, n * sizeof(float),
            cudaMemcpyHostToDevice, rb_stream));
    rbk_set_params(&v, b->lr_dev, b->mef_dev,
        !!(nets[0]->flags & RNN_NET_FLAG_BPTT_ADAPTIVE_MIN_ERROR));
  }
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  return b;
}

extern "C" void
rnn_batch_delete(RnnBatch *b)
{
  if (!b)
    return;
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  cudaFree(b->slots_dev);
  cudaFree(b->cur_dev);
  cudaFree(b->err_dev);
  cudaFree(b->winner_dev);
  cudaFree(b->accum_dev);
  cudaFree(b->lr_dev);
  cudaFree(b->mef_dev);
  cudaFree(b->io_dev);
  cudaFree(b->text_dev);
  rb_p2p_delete(b->p2p);
  cudaFreeHost(b->sym_host);
  cudaFreeHost(b->f_host);
  cudaFreeHost(b->i_host);
  cudaFreeHost(b->lr_host);
  cudaFreeHost(b->accum_host);
  cudaFree(b->rnnca_dev);
  cudaFreeHost(b->rnnca_host);
  free(b->nets);
  free(b);
}

extern "C" int
rnn_batch_p2p_export(RnnBatch *b, void *handles_out)
{
  RecurNN *proto = &b->nets[0]->pub;
  if (!b->p2p)
    b->p2p = rb_p2p_new((size_t)proto->ih_size + proto->ho_size);
  /* the exchange treats [ih_delta | ho_delta] as one block */
  if (!proto->bptt || proto->bptt->ho_delta != proto->bptt->ih_delta + proto->ih_size)
    return -1;
  return rb_p2p_export(b->p2p, handles_out);
}

extern "C" int
rnn_batch_p2p_attach(RnnBatch *b, const void *all_handles, int rank, int n_ranks)
{
  if (!b->p2p)
    return -1;
  return rb_p2p_attach(b->p2p, all_handles, rank, n_ranks);
}

extern "C" float rb_tc_exchange_probe(RbPool *p, void *p2p, RecurNN *net, int rounds);

extern "C" float
rnn_batch_p2p_probe(RnnBatch *b, int rounds)
{
  if (!b->p2p)
    return -1.0f;
  return rb_tc_exchange_probe(b->pool, b->p2p, &b->nets[0]->pub, rounds);
}

extern "C" int
rnn_batch_size(const RnnBatch *b)
{
  return b->n;
}

static void
advance_host_side(RnnBatch *b)
{
  RbPool *p = b->pool;
  for (int j = 0; j < b->n; j++) {
    RecurNN *net = &b->nets[j]->pub;
    RecurNNBPTT *bp = net->bptt;
    bp->index++;
    if (bp->index == bp->depth)
      bp->index -= bp->depth;
    net->input_layer = bp->history + (size_t)bp->index * net->i_size;
    net->real_inputs = net->input_layer + net->hidden_size + 1;
    int s = b->nets[j]->slot;
    int pos = p->pos_shadow[s] + 1;
    p->pos_shadow[s] = (pos >= p->depth) ? pos - p->depth : pos;
  }
}

/* host bookkeeping of rnn_bptt_advance for every net, one kernel for the device */
extern "C" void
rnn_batch_advance(RnnBatch *b)
{
  advance_host_side(b);
  RbView v;
  batch_view(b, &v);
  rbk_advance(&v);
}

extern "C" void
rnn_batch_set_inputs(RnnBatch *b, const float *inputs)
{
  const RbDims *d = &b->group->d;
  RecurExtraLayer *bl = b->nets[0]->pub.bottom_layer;
  int width = bl ? bl->input_size : d->input_size; /* rows feed the bottom layer if there is one */
  size_t bytes = (size_t)b->n * width * sizeof(float);
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream)); /* staging buffer reuse */
  memcpy(b->f_host, inputs, bytes);
  CUDA_OR_DIE(cudaMemcpyAsync(b->io_dev, b->f_host, bytes, cudaMemcpyHostToDevice, rb_stream));
  RbView v;
  batch_view(b, &v);
  if (bl)
    rb_bottom_set_inputs(&v, b->io_dev, width);
  else {
    rbk_set_inputs(&v, b->io_dev);
    if (b->pool->x_planes_stale < 1)
      b->pool->x_planes_stale = 1;
  }
  mark_ahead(b);
}

extern "C" void rb_forward_dispatch(const RbView *v, float noise);

/* f4, gstrnnca.c:805-830 (fill_frame): one frame of the cellular automaton.
   The three planes go to the device once, every cell gathers its inputs
   there (fill_net_inputs, :670-691), runs forward, and the next frame comes
   back as bytes.  `cells` is one forward-only net per pixel, row-major. */
extern "C" void
rnn_batch_rnnca_frame(RnnBatch *b, const u8 *frame_in, u8 *frame_out, int width, int height,
    const int *offsets_y, int len_y, const int *offsets_c, int len_c, int len_pos, int edges)
{
  const RbDims *d = &b->group->d;
  if (width * height != b->n || d->input_size != len_y + 2 * len_c + len_pos ||
      d->output_size < 3 || b->nets[0]->pub.bottom_layer)
    rb_die("recur-b200: rnn_batch_rnnca_frame: %d cells of %d inputs do not match a %d x %d "
        "frame with %d + 2*%d + %d inputs", b->n, d->input_size, width, height, len_y, len_c,
        len_pos);
  const size_t frame_bytes = 3 * (size_t)b->n;
  const size_t off_ints = 2 * (size_t)(len_y + len_c);
  if (b->rnnca_cap < frame_bytes + off_ints * sizeof(int)) {
    CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
    cudaFree(b->rnnca_dev);
    cudaFreeHost(b->rnnca_host);
    b->rnnca_cap = frame_bytes + off_ints * sizeof(int);
    CUDA_OR_DIE(cudaMalloc((void **)&b->rnnca_dev, 2 * frame_bytes + off_ints * sizeof(int) + 64));
    CUDA_OR_DIE(cudaHostAlloc((void **)&b->rnnca_host, frame_bytes + off_ints * sizeof(int) + 64,
            cudaHostAllocDefault));
  }
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream)); /* staging buffer reuse */
  /* [offsets | frame] in one copy; the outgoing frame sits behind them */
  int *off_host = (int *)b->rnnca_host;
  memcpy(off_host, offsets_y, 2 * (size_t)len_y * sizeof(int));
  memcpy(off_host + 2 * len_y, offsets_c, 2 * (size_t)len_c * sizeof(int));
  memcpy(b->rnnca_host + off_ints * sizeof(int), frame_in, frame_bytes);
  CUDA_OR_DIE(cudaMemcpyAsync(b->rnnca_dev, b->rnnca_host, off_ints * sizeof(int) + frame_bytes,
          cudaMemcpyHostToDevice, rb_stream));
  const int *off_dev = (const int *)b->rnnca_dev;
  const u8 *frame_dev = b->rnnca_dev + off_ints * sizeof(int);
  u8 *out_dev = b->rnnca_dev + ((off_ints * sizeof(int) + frame_bytes + 15) & ~(size_t)15);
  RbView v;
  batch_view(b, &v);
  rb_matrices_to_device(&b->nets[0]->pub);
  if (rbk_rnnca_cells_usable(&v)) {
    /* tiny nets: a warp per cell, frame bytes in, frame bytes out */
    rbk_rnnca_cells(&v, frame_dev, out_dev, width, height, off_dev, len_y, off_dev + 2 * len_y,
        len_c, len_pos, edges);
  }
  else {
    rbk_rnnca_gather(&v, frame_dev, width, height, off_dev, len_y, off_dev + 2 * len_y, len_c,
        len_pos, edges);
    rb_forward_dispatch(&v, 0.0f);
    rbk_rnnca_emit(&v, out_dev, width, height);
  }
  CUDA_OR_DIE(cudaMemcpyAsync(b->rnnca_host, out_dev, frame_bytes, cudaMemcpyDeviceToHost,
          rb_stream));
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  memcpy(frame_out, b->rnnca_host, frame_bytes);
  mark_ahead(b);
}

extern "C" void
rnn_batch_set_one_hot(RnnBatch *b, const u8 *hot)
{
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  memcpy(b->sym_host, hot, b->n);
  CUDA_OR_DIE(cudaMemcpyAsync(b->cur_dev, b->sym_host, b->n, cudaMemcpyHostToDevice, rb_stream));
  RbView v;
  batch_view(b, &v);
  if (b->nets[0]->pub.bottom_layer)
    rb_bottom_one_hot(&v, b->cur_dev);
  else {
    rbk_set_one_hot(&v, b->cur_dev);
    if (b->pool->x_planes_stale < 1)
      b->pool->x_planes_stale = 1;
  }
  mark_ahead(b);
}

extern "C" void rb_forward_dispatch(const RbView *v, float noise);
extern "C" int rb_last_bptt_used_tensor_engine(void);
extern "C" void rb_top_and_bptt_dispatch(const RbView *v, float *ho_delta, float *ih_delta,
    int accumulate);

static void
upload_rng_if_noisy(RnnBatch *b, float noise)
{
  if (noise == 0.0f)
    return;
  /* each stream draws from its own generator: ship the states in, and back
     out afterwards so the structs stay the source of truth */
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  for (int j = 0; j < b->n; j++) {
    CUDA_OR_DIE(cudaMemcpyAsync(b->pool->rng + (size_t)b->nets[j]->slot * 4,
            &b->nets[j]->pub.rng, sizeof(rand_ctx), cudaMemcpyHostToDevice, rb_stream));
  }
}

static void
download_rng_if_noisy(RnnBatch *b, float noise)
{
  if (noise == 0.0f)
    return;
  for (int j = 0; j < b->n; j++) {
    CUDA_OR_DIE(cudaMemcpyAsync(&b->nets[j]->pub.rng,
            b->pool->rng + (size_t)b->nets[j]->slot * 4, sizeof(rand_ctx),
            cudaMemcpyDeviceToHost, rb_stream));
  }
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
}

extern "C" void
rnn_batch_opinion(RnnBatch *b, float presynaptic_noise)
{
  rb_matrices_to_device(&b->nets[0]->pub);
  RbView v;
  batch_view(b, &v);
  upload_rng_if_noisy(b, presynaptic_noise);
  if (b->nets[0]->pub.bottom_layer)
    rb_bottom_forward(&v, &b->nets[0]->pub, NULL, presynaptic_noise);
  rb_forward_dispatch(&v, presynaptic_noise);
  download_rng_if_noisy(b, presynaptic_noise);
  mark_ahead(b);
}

static void
gather_rows(RnnBatch *b, const float *pool_array, int row_stride, int n_cols, float *out)
{
  /* rows of the batch's slots -> packed host rows */
  if (b->contiguous) {
    CUDA_OR_DIE(cudaMemcpy2DAsync(out, (size_t)n_cols * sizeof(float),
            pool_array + (size_t)b->base * row_stride, (size_t)row_stride * sizeof(float),
            (size_t)n_cols * sizeof(float), b->n, cudaMemcpyDeviceToHost, rb_stream));
  }
  else {
    for (int j = 0; j < b->n; j++)
      CUDA_OR_DIE(cudaMemcpyAsync(out + (size_t)j * n_cols,
              pool_array + (size_t)b->nets[j]->slot * row_stride, n_cols * sizeof(float),
              cudaMemcpyDeviceToHost, rb_stream));
  }
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
}

extern "C" void
rnn_batch_get_outputs(RnnBatch *b, float *outputs)
{
  const RbDims *d = &b->group->d;
  gather_rows(b, b->pool->Y, d->o_size, d->output_size, outputs);
}

extern "C" void
rnn_batch_get_hiddens(RnnBatch *b, float *hiddens)
{
  const RbDims *d = &b->group->d;
  gather_rows(b, b->pool->Hd, d->h_size, d->hidden_size + 1, hiddens);
}

extern "C" void
rnn_batch_softmax_error(RnnBatch *b, const u8 *target, float *err, int32_t *winner)
{
  RbView v;
  batch_view(b, &v);
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  memcpy(b->sym_host + b->n, target, b->n);
  CUDA_OR_DIE(cudaMemcpyAsync(b->next_dev, b->sym_host + b->n, b->n, cudaMemcpyHostToDevice, rb_stream));
  rbk_softmax_error(&v, b->next_dev, b->err_dev, b->winner_dev, NULL);
  if (err)
    CUDA_OR_DIE(cudaMemcpyAsync(err, b->err_dev, b->n * sizeof(float), cudaMemcpyDeviceToHost, rb_stream));
  if (winner)
    CUDA_OR_DIE(cudaMemcpyAsync(winner, b->winner_dev, b->n * sizeof(int), cudaMemcpyDeviceToHost, rb_stream));
  if (err || winner)
    CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  mark_ahead(b);
}

extern "C" void
rnn_batch_set_errors(RnnBatch *b, const float *o_error)
{
  const RbDims *d = &b->group->d;
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  /* pad each row out to o_size with zeros, like the calloc'ed o_error */
  for (int j = 0; j < b->n; j++) {
    float *row = b->f_host + (size_t)j * d->o_size;
    memcpy(row, o_error + (size_t)j * d->output_size, d->output_size * sizeof(float));
    for (int i = d->output_size; i < d->o_size; i++)
      row[i] = 0.0f;
  }
  if (b->contiguous) {
    CUDA_OR_DIE(cudaMemcpyAsync(b->pool->OE + (size_t)b->base * d->o_size, b->f_host,
            (size_t)b->n * d->o_size * sizeof(float), cudaMemcpyHostToDevice, rb_stream));
  }
  else {
    for (int j = 0; j < b->n; j++)
      CUDA_OR_DIE(cudaMemcpyAsync(b->pool->OE + (size_t)b->nets[j]->slot * d->o_size,
              b->f_host + (size_t)j * d->o_size, d->o_size * sizeof(float),
              cudaMemcpyHostToDevice, rb_stream));
  }
  mark_ahead(b);
}

/* learn rates live in the structs: re-upload when a caller changed one */
static void
refresh_learn_rates(RnnBatch *b, const RbView *v)
{
  int changed = 0;
  for (int j = 0; j < b->n; j++) {
    float lr = b->nets[j]->pub.bptt->learn_rate;
    if (lr != b->lr_host[j])
      changed = 1;
  }
  if (!changed)
    return;
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  for (int j = 0; j < b->n; j++)
    b->lr_host[j] = b->nets[j]->pub.bptt->learn_rate;
  CUDA_OR_DIE(cudaMemcpyAsync(b->lr_dev, b->lr_host, b->n * sizeof(float),
          cudaMemcpyHostToDevice, rb_stream));
  rbk_set_params(v, b->lr_dev, NULL, -1);
}

static void
calc_deltas_async(RnnBatch *b, int accumulate, const u8 *active = NULL)
{
  RecurNN *proto = &b->nets[0]->pub;
  RecurNNBPTT *bp = proto->bptt;
  if (accumulate && rb_comm_size() > 1)
    rb_die("recur-b200: accumulate != 0 across GPUs would exchange the previous sum again "
        "(see recur_b200.h)");
  rb_matrices_to_device(proto);
  RbView v;
  batch_view(b, &v);
  refresh_learn_rates(b, &v);
  if (active || b->masked) {
    /* upload the mask (or clear the one a previous call left) */
    const u8 *mask_dev = NULL;
    if (active) {
      CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
      memcpy(b->sym_host, active, b->n);
      CUDA_OR_DIE(cudaMemcpyAsync(b->cur_dev, b->sym_host, b->n, cudaMemcpyHostToDevice, rb_stream));
      mask_dev = b->cur_dev;
    }
    rbk_mask_streams(&v, mask_dev);
    b->masked = active != NULL;
  }
  rb_top_and_bptt_dispatch(&v, bp->ho_delta, bp->ih_delta, accumulate);
  if (proto->bottom_layer) {
    rb_bottom_backward(&v, proto, accumulate);
    /* the bottom layer's weight gradient (recur-nn.c:395-401) is summed over
       every rank's streams like the other two; it is small (input_size x
       bottom size) and not on the headline path: NCCL */
    if (rb_comm_size() > 1) {
      RecurExtraLayer *bl = proto->bottom_layer;
      rb_comm_allreduce_sum(bl->delta, (size_t)bl->i_size * bl->o_size);
    }
  }
  /* [ih_delta | ho_delta] are adjacent in the prototype's delta block; when
     the peer-memory exchange is attached and the tensor engine ran, the sum
     over ranks already happened inside the weight-gradient reduction */
  if (rb_comm_size() > 1 && !(rb_p2p_ready(b->p2p) && rb_last_bptt_used_tensor_engine()))
    rb_comm_allreduce_sum(bp->ih_delta, (size_t)proto->ih_size + proto->ho_size);
  for (int j = 0; j < b->n; j++)
    if (!active || active[j])
      b->nets[j]->pub.generation++;
  mark_ahead(b);
}

extern "C" void
rnn_batch_calc_deltas_masked(RnnBatch *b, int accumulate, const u8 *active)
{
  calc_deltas_async(b, accumulate, active);
}

extern "C" void
rnn_batch_calc_deltas(RnnBatch *b, int accumulate)
{
  calc_deltas_async(b, accumulate);
  if (b->nets[0]->pub.log) {
    /* only the prototype logs (recur-nn-init.c:237): bring its scalars back */
    rb_net_pull(b->nets[0]);
    b->nets[0]->dev_ahead = 1;
  }
}

extern "C" void
rnn_batch_apply_learning(RnnBatch *b, int learning_style, float momentum)
{
  rb_apply_learning_async(&b->nets[0]->pub, learning_style, momentum);
}

static void
fetch_stats(RnnBatch *b, RnnBatchCharStats *stats)
{
  if (b->snapshot_valid) {
    /* the output kernel left the sums in pinned memory: nothing to copy, and
       the reset rides with the next step's kernel */
    CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
    b->snapshot_valid = 0;
    b->accum_reset = 1;
  }
  else {
    CUDA_OR_DIE(cudaMemcpyAsync(b->accum_host, b->accum_dev, sizeof(RbCharAccum),
            cudaMemcpyDeviceToHost, rb_stream));
    CUDA_OR_DIE(cudaMemsetAsync(b->accum_dev, 0, sizeof(RbCharAccum), rb_stream));
    CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  }
  stats->error += b->accum_host->error;
  stats->entropy += b->accum_host->entropy;
  stats->correct += b->accum_host->correct;
  stats->count += b->accum_host->count;
}

extern "C" void rb_char_forward_dispatch(const RbView *v, const u8 *text_dev, int len, int pos,
    int spacing, u8 *cur_dev, u8 *next_dev, float noise, int advance, int continues);

/* advance .. update for one character position; the symbols come from the
   uploaded text at position `pos` (text != 0) or are already in cur/next */
static void
char_step_device(RnnBatch *b, int learning_style, float momentum, int from_text, int pos,
    int continues = 0)
{
  RecurNN *proto = &b->nets[0]->pub;
  advance_host_side(b);
  RbView v;
  batch_view(b, &v);
  rb_matrices_to_device(proto);
  float noise = proto->presynaptic_noise;
  upload_rng_if_noisy(b, noise);
  if (proto->bottom_layer) {
    if (from_text)
      rbk_text_symbols(b->text_dev, b->text_len, pos, (b->text_len - 1) / b->n, b->n,
          b->cur_dev, b->next_dev);
    rbk_advance(&v);
    rb_bottom_one_hot(&v, b->cur_dev);
    rb_bottom_forward(&v, proto, NULL, noise);
    rb_forward_dispatch(&v, noise);
  }
  else {
    /* the output kernel may take the softmax error and its sums along */
    if (b->masked) { /* skip bits a masked call left behind: this step trains every stream */
      rbk_mask_streams(&v, NULL);
      b->masked = 0;
    }
    rbk_request_fused_loss(b->next_dev, b->err_dev, b->winner_dev, b->accum_dev, b->accum_host,
        b->accum_reset);
    rbk_request_fused_top(1);
    rb_char_forward_dispatch(&v, from_text ? b->text_dev : NULL, b->text_len, pos,
        from_text ? (b->text_len - 1) / b->n : 0, b->cur_dev, b->next_dev, noise, 1,
        continues);
  }
  download_rng_if_noisy(b, noise);
  if (rbk_fused_loss_done()) {
    b->snapshot_valid = 1;
    b->accum_reset = 0;
  }
  else {
    if (b->accum_reset) {
      CUDA_OR_DIE(cudaMemsetAsync(b->accum_dev, 0, sizeof(RbCharAccum), rb_stream));
      b->accum_reset = 0;
    }
    b->snapshot_valid = 0;
    rbk_softmax_error(&v, b->next_dev, b->err_dev, b->winner_dev, b->accum_dev);
  }
  /* the update follows at once: the weight gradient may stay in its split-K
     planes until then (single GPU; an exchange needs the finished sum) */
  rb_tc_defer_delta_reduce(rb_comm_size() <= 1 || rb_p2p_ready(b->p2p));
  calc_deltas_async(b, 0);
  rb_tc_defer_delta_reduce(0);
  rb_apply_learning_async(proto, learning_style, momentum);
}

extern "C" void
rnn_batch_char_step(RnnBatch *b, const u8 *cur, const u8 *next, int learning_style,
    float momentum, RnnBatchCharStats *stats)
{
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  memcpy(b->sym_host, cur, b->n);
  memcpy(b->sym_host + b->n, next, b->n);
  /* (letting the first kernel read the symbols from the pinned buffer itself
     was measured slower than these two small copies: 512 blocks each waiting
     on a PCIe read) */
  CUDA_OR_DIE(cudaMemcpyAsync(b->cur_dev, b->sym_host, 2 * (size_t)b->n, cudaMemcpyHostToDevice,
          rb_stream));
  char_step_device(b, learning_style, momentum, 0, 0);
  if (stats)
    fetch_stats(b, stats);
}

extern "C" void
rnn_batch_text_upload(RnnBatch *b, const u8 *text, int len)
{
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  cudaFree(b->text_dev);
  CUDA_OR_DIE(cudaMalloc((void **)&b->text_dev, len));
  CUDA_OR_DIE(cudaMemcpy(b->text_dev, text, len, cudaMemcpyHostToDevice));
  b->text_len = len;
}

extern "C" int
rnn_batch_text_train(RnnBatch *b, int start, int steps, int learning_style,
    float momentum, float momentum_soft_start, RnnBatchCharStats *stats)
{
  if (!b->text_dev)
    rb_die("recur-b200: rnn_batch_text_train before rnn_batch_text_upload");
  RecurNN *proto = &b->nets[0]->pub;
  int len = b->text_len;
  int i = start;
  for (int s = 0; s < steps; s++, i++) {
    if (i >= len - 1)
      i = 0;
    float m = rnn_calculate_momentum_soft_start(proto->generation, momentum,
        momentum_soft_start);
    /* from the second position on, the stream holds nothing behind the
       previous position's update */
    char_step_device(b, learning_style, m, 1, i, s > 0);
  }
  if (stats)
    fetch_stats(b, stats);
  return (i >= len - 1) ? 0 : i;
}

extern "C" int
rnn_batch_text_forward(RnnBatch *b, int start, int steps)
{
  if (!b->text_dev)
    rb_die("recur-b200: rnn_batch_text_forward before rnn_batch_text_upload");
  int len = b->text_len;
  int spacing = (len - 1) / b->n;
  int i = start;
  RbView v;
  batch_view(b, &v);
  rb_matrices_to_device(&b->nets[0]->pub);
  rbk_output_pipeline(1); /* a step's output layer beside the next step's forward GEMM */
  for (int s = 0; s < steps; s++, i++) {
    if (i >= len - 1)
      i = 0;
    /* rnn_opinion without rnn_bptt_advance: the current ring row is rewritten */
    rb_char_forward_dispatch(&v, b->text_dev, len, i, spacing, b->cur_dev, b->next_dev, 0.0f, 0,
        0);
  }
  rbk_output_pipeline(0);
  mark_ahead(b);
  return (i >= len - 1) ? 0 : i;
}

extern "C" void
rnn_batch_bptt_depths(RnnBatch *b, int32_t *depths)
{
  RbScalars *sc = (RbScalars *)malloc(sizeof(RbScalars) * b->pool->cap);
  CUDA_OR_DIE(cudaMemcpyAsync(sc, b->pool->sc, sizeof(RbScalars) * b->pool->cap,
          cudaMemcpyDeviceToHost, rb_stream));
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  for (int j = 0; j < b->n; j++)
    depths[j] = sc[b->nets[j]->slot].n_steps;
  free(sc);
}

extern "C" void
rnn_batch_bptt_log(RnnBatch *b, RnnBatchBpttLog *log)
{
  RbScalars *sc = (RbScalars *)malloc(sizeof(RbScalars) * b->pool->cap);
  CUDA_OR_DIE(cudaMemcpyAsync(sc, b->pool->sc, sizeof(RbScalars) * b->pool->cap,
          cudaMemcpyDeviceToHost, rb_stream));
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  for (int j = 0; j < b->n; j++) {
    const RbScalars *s = sc + b->nets[j]->slot;
    RnnBatchBpttLog *l = log + j;
    l->depth = b->pool->depth - s->t_left;
    l->n_steps = s->n_steps;
    l->scaled_error = s->ih_scale * s->err_sum;
    l->ih_scale = s->ih_scale;
    l->min_error_threshold = s->min_sum;
    l->min_error_factor = s->mef;
    l->cum_error = s->cum_error;
    l->error_sum = s->err_sum;
    l->top_error_scaled = s->top_scaled;
    l->top_error_raw = s->top_raw;
  }
  free(sc);
}

extern "C" void
rnn_batch_pull(RnnBatch *b)
{
  for (int j = 0; j < b->n; j++) {
    b->nets[j]->dev_ahead = 1;
    rb_net_pull(b->nets[j]);
  }
}
