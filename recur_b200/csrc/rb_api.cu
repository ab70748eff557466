/* rb_api.cu — the RecurNN C API (include/recur-nn.h) over the device pools.
 *
 * Construction mirrors what the reference's rnn_new / new_bptt / rnn_clone
 * set up (recur-nn-init.c:6-143, 296-350) — the same fields, defaults and
 * sharing rules — but the memory behind the pointers is arranged for the GPU
 * (see rb_internal.h).  The per-net compute calls copy the caller-visible
 * vectors in, run the kernels for a batch of one stream, and copy results
 * out before returning, so they behave like the reference's synchronous C
 * functions.  The array-of-nets calls live in rb_batch.cu.
 */
#include "rb_kernels.h"
#include "rb_host.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

unsigned long long rb_ahead_epoch = 0;

#define CUDA_OR_DIE(call) do {                                          \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess)                                              \
      rb_die("recur-b200: %s failed at %s:%d: %s", #call, __FILE__,     \
          __LINE__, cudaGetErrorString(e_));                            \
  } while (0)

static inline size_t
align4(size_t n)
{
  return (n + 3) & ~(size_t)3;
}

extern "C" RbNet *
rb_net_of(RecurNN *net)
{
  if (!net)
    rb_die("recur-b200: NULL net");
  RbNet *rn = (RbNet *)((char *)net - offsetof(RbNet, pub));
  if (rn->magic != RB_MAGIC)
    rb_die("recur-b200: RecurNN %p was not created by this library", (void *)net);
  return rn;
}

/* ---- views ----------------------------------------------------------------- */

extern "C" void
rb_view_of_net(RbNet *rn, RbView *v)
{
  RbPool *p = rn->pool;
  memset(v, 0, sizeof(*v));
  v->d = rn->group->d;
  v->cap = p->cap;
  v->depth = p->depth;
  v->n_part = p->n_part;
  v->X = p->X;
  v->Hd = p->Hd;
  v->Y = p->Y;
  v->OE = p->OE;
  v->E = p->E;
  v->partial = p->partial;
  v->noise = p->noise;
  v->pos = p->pos;
  v->sc = p->sc;
  v->rng = p->rng;
  v->slots = p->iota + rn->slot;
  v->n = 1;
  v->contiguous = 1;
  v->base = rn->slot;
  v->Wih = rn->pub.ih_weights;
  v->Who = rn->pub.ho_weights;
  v->activation = rn->pub.activation;
  v->pool = p;
  if (rb_have_device())
    rb_bottom_attach(v, &rn->pub);
}

static float *
dev_x_row(RbNet *rn, int back)
{
  RbPool *p = rn->pool;
  int pos = p->pos_shadow[rn->slot] - back;
  while (pos < 0)
    pos += p->depth;
  return p->X + ((size_t)pos * p->cap + rn->slot) * rn->group->d.i_size;
}

static void
h2d(void *dst, const void *src, size_t bytes)
{
  CUDA_OR_DIE(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, rb_stream));
}

static void
d2h(void *dst, const void *src, size_t bytes)
{
  CUDA_OR_DIE(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, rb_stream));
}

static void
sync_stream(void)
{
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
}

/* ---- mirrors --------------------------------------------------------------- */

/* device -> host for everything a caller can see of this stream */
extern "C" void
rb_net_pull(RbNet *rn)
{
  if (!rb_have_device())
    return;
  RecurNN *net = &rn->pub;
  RbPool *p = rn->pool;
  const RbDims *d = &rn->group->d;
  int s = rn->slot;
  d2h(net->hidden_layer, p->Hd + (size_t)s * d->h_size, d->h_size * sizeof(float));
  d2h(net->output_layer, p->Y + (size_t)s * d->o_size, d->o_size * sizeof(float));
  RbScalars sc;
  memset(&sc, 0, sizeof(sc));
  if (net->bptt && p->has_bptt) {
    RecurNNBPTT *b = net->bptt;
    /* ring: host slot (index - j) <- device slot (pos - j) */
    for (int j = 0; j < b->depth; j++) {
      int hslot = b->index - j;
      while (hslot < 0)
        hslot += b->depth;
      d2h(b->history + (size_t)hslot * d->i_size, dev_x_row(rn, j), d->i_size * sizeof(float));
    }
    d2h(b->o_error, p->OE + (size_t)s * d->o_size, d->o_size * sizeof(float));
    d2h(&sc, p->sc + s, sizeof(sc));
    sync_stream();
    /* the reference's two error buffers swap roles each BPTT step
       (recur-nn.c:384-386): after n steps the last two rows sit in them */
    int n = sc.n_steps;
    if (n >= 1) {
      float *last = (n & 1) ? b->i_error : b->h_error;
      float *prev = (n & 1) ? b->h_error : b->i_error;
      d2h(last, p->E + ((size_t)n * p->cap + s) * d->i_size, d->i_size * sizeof(float));
      d2h(prev, p->E + ((size_t)(n - 1) * p->cap + s) * d->i_size, d->i_size * sizeof(float));
    }
    else {
      d2h(b->h_error, p->E + (size_t)s * d->i_size, d->i_size * sizeof(float));
    }
    if (rn->dev_ahead) {
      b->ih_scale = sc.ih_scale;
      b->min_error_factor = sc.mef;
    }
  }
  else {
    d2h(net->input_layer, dev_x_row(rn, 0), d->i_size * sizeof(float));
  }
  sync_stream();
  rn->dev_ahead = 0;
  rb_ahead_epoch++; /* batches re-flag their nets on their next call */
}

/* host -> device */
extern "C" void
rb_net_push(RbNet *rn)
{
  if (!rb_have_device())
    return;
  RecurNN *net = &rn->pub;
  RbPool *p = rn->pool;
  const RbDims *d = &rn->group->d;
  int s = rn->slot;
  h2d(p->Hd + (size_t)s * d->h_size, net->hidden_layer, d->h_size * sizeof(float));
  if (net->bptt && p->has_bptt) {
    RecurNNBPTT *b = net->bptt;
    for (int j = 0; j < b->depth; j++) {
      int hslot = b->index - j;
      while (hslot < 0)
        hslot += b->depth;
      h2d(dev_x_row(rn, j), b->history + (size_t)hslot * d->i_size, d->i_size * sizeof(float));
    }
    h2d(p->OE + (size_t)s * d->o_size, b->o_error, d->o_size * sizeof(float));
  }
  else {
    h2d(dev_x_row(rn, 0), net->input_layer, d->i_size * sizeof(float));
  }
  p->x_planes_stale = 2;
  sync_stream();
  rn->dev_ahead = 0;
  rb_ahead_epoch++; /* batches re-flag their nets on their next call */
}

extern "C" void
rnn_b200_pull(RecurNN *net)
{
  rb_net_pull(rb_net_of(net));
}

extern "C" void
rnn_b200_push(RecurNN *net)
{
  rb_net_push(rb_net_of(net));
}

/* a batch call left newer state on the device than the mirrors hold */
static inline void
catch_up(RbNet *rn)
{
  if (rn->dev_ahead)
    rb_net_pull(rn);
}

/* ---- construction ---------------------------------------------------------- */

static void
set_ring_pointers(RecurNN *net)
{
  RecurNNBPTT *b = net->bptt;
  net->input_layer = b->history + (size_t)b->index * net->i_size;
  net->real_inputs = net->input_layer + net->hidden_size + 1;
}

static RecurNNBPTT *
new_bptt_state(RbNet *rn, int depth, float learn_rate, float momentum, u32 flags)
{
  RecurNN *net = &rn->pub;
  RecurNNBPTT *b = (RecurNNBPTT *)calloc(1, sizeof(RecurNNBPTT));
  if (!b)
    rb_die("recur-b200: out of memory");
  const size_t I = net->i_size, O = net->o_size;
  const size_t ih = net->ih_size, ho = net->ho_size;
  b->depth = depth;
  b->learn_rate = learn_rate;
  b->momentum = momentum;
  b->momentum_weight = RNN_MOMENTUM_WEIGHT;
  /* per-stream vectors: pinned mirror block */
  size_t n_mirror = O + 2 * I + (size_t)depth * I;
  float *m = (float *)rb_alloc_mirror(n_mirror * sizeof(float));
  rn->mirror_bptt = m;
  b->mem = m;
  b->o_error = m;
  m += O;
  b->i_error = m;
  m += I;
  b->h_error = m;
  m += I;
  b->history = m;
  /* shared matrices */
  if (!(flags & RNN_NET_FLAG_NO_MOMENTUMS)) {
    rn->n_momentums = ih + ho;
    rn->own_momentums = rb_alloc_matrix(rn->n_momentums);
    b->ih_momentum = rn->own_momentums;
    b->ho_momentum = rn->own_momentums + ih;
  }
  if (!(flags & RNN_NET_FLAG_NO_DELTAS)) {
    rn->n_deltas = ih + ho + ih;
    rn->own_deltas = rb_alloc_matrix(rn->n_deltas);
    b->ih_delta = rn->own_deltas;
    b->ho_delta = rn->own_deltas + ih;
    b->ih_delta_tmp = rn->own_deltas + ih + ho;
  }
  if ((flags & RNN_NET_FLAG_AUX_ARRAYS) && !(flags & RNN_NET_FLAG_NO_MOMENTUMS)) {
    rn->n_aux = ih + ho;
    rn->own_aux = rb_alloc_matrix(rn->n_aux);
    b->ih_aux = rn->own_aux;
    b->ho_aux = rn->own_aux + ih;
  }
  b->index = 0;
  b->ho_scale = 1.0f;
  b->ih_scale = 1.0f;
  b->min_error_factor = BASE_MIN_ERROR_FACTOR * net->h_size;
  return b;
}

/* rnn_new with an optional group to join (clones that borrow weights) */
static RecurNN *
net_create(uint input_size, uint hidden_size, uint output_size, u32 flags,
    u64 rng_seed, const char *log_file, int bptt_depth, float learn_rate,
    float momentum, float presynaptic_noise, rnn_activation activation,
    RbGroup *join)
{
  RbNet *rn = (RbNet *)calloc(1, sizeof(RbNet));
  if (!rn)
    rb_die("recur-b200: out of memory");
  rn->magic = RB_MAGIC;
  RecurNN *net = &rn->pub;
  size_t i_size = align4(hidden_size + input_size + 1);
  size_t h_size = align4(hidden_size + 1);
  size_t o_size = align4(output_size);
  net->i_size = (int)i_size;
  net->h_size = (int)h_size;
  net->o_size = (int)o_size;
  net->input_size = (int)input_size;
  net->hidden_size = (int)hidden_size;
  net->output_size = (int)output_size;
  net->ih_size = (int)(i_size * h_size);
  net->ho_size = (int)(h_size * o_size);
  net->generation = 0;
  net->flags = flags;
  net->presynaptic_noise = presynaptic_noise;
  if ((int)activation >= RNN_ACTIVATION_LAST)
    activation = RNN_RELU; /* recur-nn-init.c:104-106 */
  net->activation = activation;
  rb_init_rand64_maybe_randomly(&net->rng, rng_seed);

  float *m = (float *)rb_alloc_mirror((i_size + h_size + o_size) * sizeof(float));
  rn->mirror_net = m;
  net->mem = m;
  net->input_layer = m;
  m += i_size;
  net->hidden_layer = m;
  m += h_size;
  net->output_layer = m;

  if (flags & RNN_NET_FLAG_OWN_WEIGHTS) {
    rn->n_weights = (size_t)net->ih_size + net->ho_size;
    rn->own_weights = rb_alloc_matrix(rn->n_weights);
    net->ih_weights = rn->own_weights;
    net->ho_weights = rn->own_weights + net->ih_size;
  }
  if (join && !(flags & RNN_NET_FLAG_OWN_WEIGHTS)) {
    rn->group = join;
  }
  else {
    RbDims d = {net->i_size, net->h_size, net->o_size,
                net->input_size, net->hidden_size, net->output_size};
    rn->group = rb_group_new(&d);
  }
  rn->group->refs++;

  int has_bptt = (flags & RNN_NET_FLAG_OWN_BPTT) ? 1 : 0;
  rn->pool = rb_group_pool(rn->group, has_bptt ? bptt_depth : 1, has_bptt);
  rn->slot = rb_pool_take_slot(rn->pool);

  if (has_bptt) {
    net->bptt = new_bptt_state(rn, bptt_depth, learn_rate, momentum, flags);
    rnn_bptt_advance(net);
  }
  else {
    net->real_inputs = net->input_layer + net->hidden_size + 1;
  }
  if (log_file)
    rnn_set_log_file(net, log_file, flags & RNN_NET_FLAG_LOG_APPEND);
  return net;
}

/* reference recur-nn-init.c:80-143 */
extern "C" RecurNN *
rnn_new(uint input_size, uint hidden_size, uint output_size, u32 flags,
    u64 rng_seed, const char *log_file, int bptt_depth, float learn_rate,
    float momentum, float presynaptic_noise, rnn_activation activation)
{
  return net_create(input_size, hidden_size, output_size, flags, rng_seed, log_file,
      bptt_depth, learn_rate, momentum, presynaptic_noise, activation, NULL);
}

/* reference recur-nn-init.c:145-155: metadata and the bottom layer are not
   freed there either */
extern "C" void
rnn_delete_net(RecurNN *net)
{
  RbNet *rn = rb_net_of(net);
  if (rb_have_device())
    sync_stream();
  if (net->bptt && (net->flags & RNN_NET_FLAG_OWN_BPTT)) {
    rb_free_mirror(rn->mirror_bptt);
    free(net->bptt);
  }
  if (net->log)
    fclose(net->log);
  rb_free_matrix(rn->own_weights);
  rb_free_matrix(rn->own_momentums);
  rb_free_matrix(rn->own_deltas);
  rb_free_matrix(rn->own_aux);
  rb_free_mirror(rn->mirror_net);
  rb_pool_release_slot(rn->pool, rn->slot);
  rb_group_unref(rn->group);
  rn->magic = 0;
  free(rn);
}

/* reference recur-nn-init.c:296-350 */
extern "C" RecurNN *
rnn_clone(RecurNN *parent, u32 flags, u64 rng_seed, const char *log_file)
{
  RbNet *prn = rb_net_of(parent);
  if (rng_seed == RECUR_RNG_SUBSEED) {
    do {
      rng_seed = rb_rand64(&parent->rng);
    } while (rng_seed == RECUR_RNG_RANDOM_SEED);
  }
  float learn_rate = 0, momentum = 0;
  int depth = 0;
  int with_bptt = parent->bptt && (flags & RNN_NET_FLAG_OWN_BPTT);
  if (with_bptt) {
    learn_rate = parent->bptt->learn_rate;
    depth = parent->bptt->depth;
    momentum = parent->bptt->momentum;
  }
  if (flags & RNN_NET_FLAG_OWN_WEIGHTS)
    rb_host_will_touch_matrices(parent);
  RecurNN *net = net_create(parent->input_size, parent->hidden_size,
      parent->output_size, flags, rng_seed, log_file, depth, learn_rate, momentum,
      parent->presynaptic_noise, parent->activation, prn->group);
  if (with_bptt) {
    net->bptt->momentum_weight = parent->bptt->momentum_weight;
    if (flags & RNN_NET_FLAG_NO_MOMENTUMS) {
      net->bptt->ih_momentum = parent->bptt->ih_momentum;
      net->bptt->ho_momentum = parent->bptt->ho_momentum;
      net->bptt->ih_aux = parent->bptt->ih_aux;
      net->bptt->ho_aux = parent->bptt->ho_aux;
    }
    if (flags & RNN_NET_FLAG_NO_DELTAS) {
      net->bptt->ih_delta = parent->bptt->ih_delta;
      net->bptt->ho_delta = parent->bptt->ho_delta;
    }
  }
  if (flags & RNN_NET_FLAG_OWN_WEIGHTS) {
    memcpy(net->ih_weights, parent->ih_weights, net->ih_size * sizeof(float));
    memcpy(net->ho_weights, parent->ho_weights, net->ho_size * sizeof(float));
  }
  else {
    net->ih_weights = parent->ih_weights;
    net->ho_weights = parent->ho_weights;
  }
  net->bottom_layer = parent->bottom_layer;
  net->generation = parent->generation;
  net->presynaptic_noise = parent->presynaptic_noise;
  return net;
}

/* reference recur-nn-init.c:221-243 */
extern "C" RecurNN **
rnn_new_training_set(RecurNN *prototype, int n_nets)
{
  if (n_nets < 1) {
    fprintf(stderr, "A training set of size %d is not possible\n", n_nets);
    return NULL;
  }
  RbNet *prn = rb_net_of(prototype);
  RecurNN **nets = (RecurNN **)malloc(n_nets * sizeof(RecurNN *));
  if (!nets)
    rb_die("recur-b200: out of memory");
  nets[0] = prototype;
  u32 flags = prototype->flags;
  flags &= ~RNN_NET_FLAG_OWN_WEIGHTS;
  flags |= RNN_NET_FLAG_NO_MOMENTUMS;
  flags |= RNN_NET_FLAG_NO_DELTAS;
  /* the streams of a training set sit side by side in the pool */
  rb_pool_reserve(prn->pool, prn->pool->n_live + n_nets - 1);
  for (int i = 1; i < n_nets; i++) {
    nets[i] = rnn_clone(prototype, flags, RECUR_RNG_SUBSEED, NULL);
    nets[i]->bptt->ih_delta = prototype->bptt->ih_delta;
    nets[i]->bptt->ih_delta_tmp = prototype->bptt->ih_delta_tmp;
    nets[i]->bptt->ho_delta = prototype->bptt->ho_delta;
  }
  return nets;
}

/* reference recur-nn-init.c:245-257 */
extern "C" void
rnn_delete_training_set(RecurNN **nets, int n_nets, int leave_prototype)
{
  for (int i = !!leave_prototype; i < n_nets; i++) {
    if (nets[i])
      rnn_delete_net(nets[i]);
  }
  free(nets);
}

/* reference recur-nn-init.c:158-192 */
extern "C" RecurExtraLayer *
rnn_new_extra_layer(int input_size, int output_size, int overlap, u32 flags)
{
  RecurExtraLayer *layer = (RecurExtraLayer *)calloc(1, sizeof(RecurExtraLayer));
  if (!layer)
    rb_die("recur-b200: out of memory");
  layer->input_size = input_size;
  layer->output_size = output_size;
  layer->overlap = overlap;
  layer->learn_rate_scale = 1.0f;
  layer->i_size = (int)align4(input_size + 1);
  layer->o_size = (int)align4(output_size);
  size_t matrix = (size_t)layer->i_size * layer->o_size;
  int with_aux = !!(flags & RNN_NET_FLAG_AUX_ARRAYS);
  /* matrices: managed memory, like the recurrent layer's; the four small
     vectors: pinned host memory the kernels reach directly, because callers
     write inputs / read outputs around every call (one_hot_opinion writes
     bottom_layer->inputs, charmodel-helpers.h:19-31) and must not drag
     matrix pages back and forth.  (The reference never frees a bottom
     layer, recur-nn-init.c:145-155; neither does this library.) */
  float *m = rb_alloc_matrix(matrix * (3 + with_aux));
  layer->mem = m;
  layer->momentums = m;
  m += matrix;
  layer->weights = m;
  m += matrix;
  layer->delta = m;
  m += matrix;
  if (with_aux)
    layer->aux = m;
  float *vec = (float *)rb_alloc_mirror(2 * (size_t)(layer->i_size + layer->o_size) * sizeof(float));
  layer->inputs = vec;
  vec += layer->i_size;
  layer->outputs = vec;
  vec += layer->o_size;
  layer->i_error = vec;
  vec += layer->i_size;
  layer->o_error = vec;
  return layer;
}

/* reference recur-nn-init.c:194-219 */
extern "C" RecurNN *
rnn_new_with_bottom_layer(int n_inputs, int r_input_size, int hidden_size,
    int output_size, u32 flags, u64 rng_seed, const char *log_file,
    int bptt_depth, float learn_rate, float momentum, float presynaptic_noise,
    rnn_activation activation, int convolutional_overlap)
{
  if (r_input_size == 0) {
    flags &= ~RNN_NET_FLAG_BOTTOM_LAYER;
    return rnn_new(n_inputs, hidden_size, output_size, flags, rng_seed, log_file,
        bptt_depth, learn_rate, momentum, presynaptic_noise, activation);
  }
  flags |= RNN_NET_FLAG_BOTTOM_LAYER;
  RecurNN *net = rnn_new(r_input_size, hidden_size, output_size, flags, rng_seed,
      log_file, bptt_depth, learn_rate, momentum, presynaptic_noise, activation);
  net->bottom_layer = rnn_new_extra_layer(n_inputs, r_input_size,
      convolutional_overlap, net->flags);
  return net;
}

/* reference recur-nn-init.c:259-283 */
extern "C" void
rnn_set_log_file(RecurNN *net, const char *log_file, int append_dont_truncate)
{
  if (net->log)
    fclose(net->log);
  if (log_file) {
    net->log = fopen(log_file, append_dont_truncate ? "a" : "w");
    if (!append_dont_truncate)
      rnn_log_int(net, "generation", net->generation);
  }
  else {
    net->log = NULL;
  }
}

/* ---- a1 -------------------------------------------------------------------- */

/* reference recur-nn.c:696-704 */
extern "C" void
rnn_bptt_advance(RecurNN *net)
{
  RbNet *rn = rb_net_of(net);
  RecurNNBPTT *b = net->bptt;
  b->index++;
  if (b->index == b->depth)
    b->index -= b->depth;
  set_ring_pointers(net);
  RbPool *p = rn->pool;
  int pos = p->pos_shadow[rn->slot] + 1;
  p->pos_shadow[rn->slot] = (pos >= p->depth) ? pos - p->depth : pos;
  if (rb_have_device()) {
    RbView v;
    rb_view_of_net(rn, &v);
    rbk_advance(&v);
  }
}

/* ---- a3 -------------------------------------------------------------------- */

/* reference recur-nn.c:83-154 */
extern "C" float *
rnn_opinion(RecurNN *net, const float *inputs, float presynaptic_noise)
{
  rb_require_device("rnn_opinion");
  RbNet *rn = rb_net_of(net);
  catch_up(rn);
  rb_matrices_to_device(net);
  RbPool *p = rn->pool;
  const RbDims *d = &rn->group->d;
  int s = rn->slot;
  RecurExtraLayer *bl = net->bottom_layer;
  if (bl) {
    /* recur-nn.c:88-94: the layer's own (shared) input vector is the source */
    bl->inputs[0] = 1.0f;
    if (inputs)
      memcpy(bl->inputs + 1, inputs, bl->input_size * sizeof(float));
  }
  else if (inputs)
    memcpy(net->real_inputs, inputs, net->input_size * sizeof(float));
  RbView v;
  rb_view_of_net(rn, &v);
  p->x_planes_stale = 2; /* the per-net forward writes no operand planes */
  if (!bl && presynaptic_noise == 0.0f && rbk_opinion_single_usable(&v)) {
    /* one launch that reads and writes the pinned mirrors itself */
    rbk_opinion_single(&v, net->hidden_layer, net->real_inputs, net->input_layer,
        net->hidden_layer, net->output_layer);
    sync_stream();
    return net->output_layer;
  }
  float *xrow = dev_x_row(rn, 0);
  h2d(p->Hd + (size_t)s * d->h_size, net->hidden_layer, d->h_size * sizeof(float));
  if (!bl)
    h2d(xrow + d->hidden_size + 1, net->real_inputs, d->input_size * sizeof(float));
  if (presynaptic_noise != 0.0f)
    h2d(p->rng + (size_t)s * 4, &net->rng, sizeof(rand_ctx));
  if (bl)
    rb_bottom_forward(&v, net, bl->inputs, presynaptic_noise);
  rbk_forward(&v, presynaptic_noise);
  if (presynaptic_noise != 0.0f)
    d2h(&net->rng, p->rng + (size_t)s * 4, sizeof(rand_ctx));
  d2h(net->input_layer, xrow, d->i_size * sizeof(float));
  d2h(net->hidden_layer, p->Hd + (size_t)s * d->h_size, d->h_size * sizeof(float));
  d2h(net->output_layer, p->Y + (size_t)s * d->o_size, d->o_size * sizeof(float));
  sync_stream();
  return net->output_layer;
}

/* reference recur-nn.c:8-16 */
extern "C" void
rnn_forget_history(RecurNN *net, int bptt_too)
{
  RbNet *rn = rb_net_of(net);
  catch_up(rn);
  memset(net->hidden_layer, 0, net->h_size * sizeof(float));
  memset(net->input_layer, 0, (net->hidden_size + 1) * sizeof(float));
  if (bptt_too && net->bptt)
    memset(net->bptt->history, 0, (size_t)net->bptt->depth * net->i_size * sizeof(float));
  if (rb_have_device()) {
    RbPool *p = rn->pool;
    const RbDims *d = &rn->group->d;
    p->x_planes_stale = 2;
    CUDA_OR_DIE(cudaMemsetAsync(p->Hd + (size_t)rn->slot * d->h_size, 0,
            d->h_size * sizeof(float), rb_stream));
    CUDA_OR_DIE(cudaMemsetAsync(dev_x_row(rn, 0), 0,
            (d->hidden_size + 1) * sizeof(float), rb_stream));
    if (bptt_too && net->bptt) {
      for (int j = 0; j < p->depth; j++)
        CUDA_OR_DIE(cudaMemsetAsync(dev_x_row(rn, j), 0, d->i_size * sizeof(float), rb_stream));
    }
    sync_stream();
  }
}

/* ---- a7..a12 --------------------------------------------------------------- */

static RecurErrorRange *ranges_dev = NULL;
static int ranges_cap = 0;

static int
upload_ranges(const RecurErrorRange *ranges)
{
  if (!ranges)
    return 0;
  int n = 0;
  while (ranges[n].start >= 0)
    n++;
  if (n > ranges_cap) {
    if (ranges_dev)
      cudaFree(ranges_dev);
    ranges_cap = n + 16;
    CUDA_OR_DIE(cudaMalloc((void **)&ranges_dev, ranges_cap * sizeof(RecurErrorRange)));
  }
  if (n)
    CUDA_OR_DIE(cudaMemcpyAsync(ranges_dev, ranges, n * sizeof(RecurErrorRange),
            cudaMemcpyHostToDevice, rb_stream));
  return n;
}

/* the log block of bptt_and_accumulate_error, reference recur-nn.c:415-448 */
static void
log_bptt_block(RecurNN *net, const RbScalars *sc)
{
  if (!net->log)
    return;
  rnn_log_int(net, "depth", net->bptt->depth - sc->t_left);
  rnn_log_float(net, "scaled_error", sc->ih_scale * sc->err_sum);
  rnn_log_float(net, "ih_scale", sc->ih_scale);
  rnn_log_float(net, "min_error_threshold", sc->min_sum);
  rnn_log_float(net, "min_error_factor", sc->mef);
  rnn_log_float(net, "cum_error", sc->cum_error);
  if (net->bottom_layer) {
    float cie = 0;
    for (int y = 0; y < net->input_size; y++)
      cie += net->bottom_layer->o_error[y];
    rnn_log_float(net, "cum_input_error", cie);
  }
  if (net->flags & RNN_NET_FLAG_LOG_HIDDEN_SUM) {
    rnn_log_float(net, "hidden_sum", sc->hidden_sum);
    rnn_log_float(net, "hidden_magnitude", sc->hidden_mag);
    rnn_log_float(net, "hidden_zeros", sc->hidden_zeros / (float)net->hidden_size);
  }
  if (net->flags & RNN_NET_FLAG_LOG_WEIGHT_SUM) {
    static float *sum_dev = NULL;
    float sum = 0;
    if (!sum_dev)
      CUDA_OR_DIE(cudaMalloc((void **)&sum_dev, sizeof(float)));
    rbk_abs_sum(net->ih_weights, net->ih_size, sum_dev);
    d2h(&sum, sum_dev, sizeof(float));
    sync_stream();
    rnn_log_float(net, "weight_sum", sum);
  }
}

/* reference recur-nn.c:707-772 */
extern "C" void
rnn_bptt_calc_deltas(RecurNN *net, int accumulate_delta, RecurErrorRange *top_error_ranges)
{
  rb_require_device("rnn_bptt_calc_deltas");
  RbNet *rn = rb_net_of(net);
  catch_up(rn);
  rb_matrices_to_device(net);
  RecurNNBPTT *b = net->bptt;
  RbPool *p = rn->pool;
  const RbDims *d = &rn->group->d;
  int s = rn->slot;
  h2d(p->OE + (size_t)s * d->o_size, b->o_error, d->o_size * sizeof(float));
  RbView v;
  rb_view_of_net(rn, &v);
  rbk_set_params_scalar(&v, b->learn_rate, b->min_error_factor,
      !!(net->flags & RNN_NET_FLAG_BPTT_ADAPTIVE_MIN_ERROR));
  int n_ranges = upload_ranges(top_error_ranges);
  rbk_top_layer(&v, b->ho_delta, accumulate_delta, ranges_dev, n_ranges);
  rbk_bptt(&v, b->ih_delta, accumulate_delta);
  if (net->bottom_layer)
    rb_bottom_backward(&v, net, accumulate_delta);
  RbScalars sc;
  d2h(&sc, p->sc + s, sizeof(sc));
  sync_stream();
  b->ih_scale = sc.ih_scale;
  b->min_error_factor = sc.mef;
  log_bptt_block(net, &sc);
  if (net->bottom_layer && net->log) {
    RecurExtraLayer *bl = net->bottom_layer;
    float be = 0;
    for (int i = 0; i < bl->output_size; i++)
      be += fabsf(bl->o_error[i]);
    rnn_log_float(net, "bottom_error", be);
  }
  net->generation++;
  if (net->log) {
    rnn_log_float(net, "error_gain", sc.err_sum / (sc.top_scaled + 1e-6));
    rnn_log_float(net, "top_error_scaled", sc.top_scaled);
    rnn_log_float(net, "top_error_raw", sc.top_raw);
    rnn_log_int(net, "generation", net->generation);
  }
}

/* ---- a13 ------------------------------------------------------------------- */

/* reference recur-nn.c:595-599 */
extern "C" float
rnn_calculate_momentum_soft_start(float generation, float max_momentum, float x)
{
  float m = 1.0f - x / (1.0f + generation + 2.0f * x);
  return (max_momentum < m) ? max_momentum : m;
}

/* the dispatch of reference recur-nn.c:601-678 */
extern "C" void
rb_apply_learning_async(RecurNN *net, int method, float momentum)
{
  RecurNNBPTT *b = net->bptt;
  RecurExtraLayer *bl = net->bottom_layer;
  float mw = 0.0f;
  int kernel_method = method;
  switch (method) {
  case RNN_MOMENTUM_NESTEROV:
  case RNN_ADAGRAD:
  case RNN_ADADELTA:
  case RNN_RPROP:
    break;
  case RNN_MOMENTUM_SIMPLIFIED_NESTEROV:
    mw = (float)(momentum / (1.0 + momentum));
    kernel_method = RNN_MOMENTUM_WEIGHTED;
    break;
  case RNN_MOMENTUM_CLASSICAL:
    mw = 1.0f;
    kernel_method = RNN_MOMENTUM_WEIGHTED;
    break;
  default:
    mw = b->momentum_weight;
    kernel_method = RNN_MOMENTUM_WEIGHTED;
    break;
  }
  if ((method == RNN_ADADELTA || method == RNN_RPROP) && !b->ih_aux)
    rb_die("recur-b200: learning method %d needs RNN_NET_FLAG_AUX_ARRAYS", method);
  /* where the tensor engine holds operand planes of these weights, one kernel
     updates both matrices and rewrites the planes */
  RbPool *pool = rb_net_of(net)->pool;
  rb_mark_pre_update();
  int fused = rb_tc_fused_update(pool, net, kernel_method, momentum, mw);
  if (!fused) {
    rbk_apply_learning(kernel_method, net->ho_weights, b->ho_delta, b->ho_momentum,
        b->ho_aux, net->ho_size, b->learn_rate * b->ho_scale, momentum, mw, NULL);
    rbk_apply_learning(kernel_method, net->ih_weights, b->ih_delta, b->ih_momentum,
        b->ih_aux, net->ih_size, b->learn_rate, momentum, mw, NULL);
  }
  if (bl) {
    rbk_apply_learning(kernel_method, bl->weights, bl->delta, bl->momentums, bl->aux,
        bl->i_size * bl->o_size, b->learn_rate * bl->learn_rate_scale, momentum, mw, NULL);
  }
  rb_weights_changed(net);
  if (fused)
    rb_tc_planes_current(pool);
}

extern "C" void
rnn_apply_learning(RecurNN *net, int learning_method, float momentum)
{
  rb_require_device("rnn_apply_learning");
  rb_matrices_to_device(net);
  rb_apply_learning_async(net, learning_method, momentum);
  sync_stream();
}

/* reference recur-nn.c:681-693 */
extern "C" void
rnn_bptt_clear_deltas(RecurNN *net)
{
  rb_require_device("rnn_bptt_clear_deltas");
  RecurNNBPTT *b = net->bptt;
  rbk_fill(b->ih_delta, net->ih_size, 0.0f);
  rbk_fill(b->ho_delta, net->ho_size, 0.0f);
  if (net->bottom_layer) {
    RecurExtraLayer *bl = net->bottom_layer;
    rbk_fill(bl->o_error, bl->o_size, 0.0f);
    rbk_fill(bl->delta, (size_t)bl->i_size * bl->o_size, 0.0f);
  }
  sync_stream();
}

/* ---- a14: the single-net path, reference recur-nn.c:919-1019 -------------- */

extern "C" void
rnn_bptt_calculate(RecurNN *net, uint batch_size)
{
  rb_require_device("rnn_bptt_calculate");
  RbNet *rn = rb_net_of(net);
  catch_up(rn);
  rb_matrices_to_device(net);
  RecurNNBPTT *b = net->bptt;
  RbPool *p = rn->pool;
  const RbDims *d = &rn->group->d;
  int s = rn->slot;
  net->hidden_layer[0] = 1.0f;
  RbView v;
  rb_view_of_net(rn, &v);
  RbScalars sc;
  if (batch_size <= 1 && !net->bottom_layer && rbk_walk_single_usable(&v)) {
    /* top layer, Who update, the whole walk and the Wih update in one launch;
       o_error is read from its pinned mirror, the scalars come back the same way */
    static RbScalars *sc_pinned = NULL;
    if (!sc_pinned)
      CUDA_OR_DIE(cudaHostAlloc((void **)&sc_pinned, sizeof(RbScalars), cudaHostAllocDefault));
    rbk_calculate_single(&v, b->o_error, b->learn_rate, b->min_error_factor,
        !!(net->flags & RNN_NET_FLAG_BPTT_ADAPTIVE_MIN_ERROR), net->ho_weights, b->ho_momentum,
        net->ih_weights, b->ih_momentum, b->ih_delta, b->momentum, b->momentum_weight,
        sc_pinned);
    rb_weights_changed(net);
    sync_stream();
    sc = *sc_pinned;
    goto finish;
  }
  h2d(p->OE + (size_t)s * d->o_size, b->o_error, d->o_size * sizeof(float));
  rbk_set_params_scalar(&v, b->learn_rate, b->min_error_factor,
      !!(net->flags & RNN_NET_FLAG_BPTT_ADAPTIVE_MIN_ERROR));
  /* apply_sgd_top_layer: error back through the old weights, then the
     immediate update of ho (rate lr, not lr*ho_scale) */
  rbk_top_layer(&v, NULL, 0, NULL, 0);
  rbk_sgd_top_apply(&v, net->ho_weights, b->ho_momentum, b->learn_rate, b->momentum,
      b->momentum_weight);
  if (batch_size > 1) {
    /* apply_sgd_with_bptt_batch: ih_delta += ih_scale * G, applied every
       batch_size generations */
    rbk_bptt(&v, b->ih_delta, 1);
    if ((net->generation % batch_size) == 0) {
      rbk_apply_learning(RNN_MOMENTUM_WEIGHTED, net->ih_weights, b->ih_delta,
          b->ih_momentum, NULL, net->ih_size, b->learn_rate, b->momentum,
          b->momentum_weight, NULL);
      rbk_fill(b->ih_delta, net->ih_size, 0.0f);
    }
  }
  else {
    /* apply_sgd_with_bptt: ih_delta already carries ih_scale here, which the
       reference folds into the rate instead */
    rbk_bptt(&v, b->ih_delta, 0);
    rbk_apply_learning(RNN_MOMENTUM_WEIGHTED, net->ih_weights, b->ih_delta,
        b->ih_momentum, NULL, net->ih_size, b->learn_rate, b->momentum,
        b->momentum_weight, NULL);
  }
  rb_weights_changed(net);
  d2h(&sc, p->sc + s, sizeof(sc));
  sync_stream();
finish:
  b->ih_scale = sc.ih_scale;
  b->min_error_factor = sc.mef;
  log_bptt_block(net, &sc);
  net->generation++;
  if (net->log) {
    rnn_log_float(net, "top_error_scaled", sc.top_scaled);
    rnn_log_float(net, "top_error_raw", sc.top_raw);
    rnn_log_float(net, "error_sum", sc.err_sum);
    rnn_log_float(net, "error_gain", sc.err_sum / (sc.top_scaled + 1e-6));
    rnn_log_int(net, "generation", net->generation);
  }
  rnn_condition_net(net);
}

/* ---- a15: reference recur-nn.c:782-855 ------------------------------------- */

extern "C" void
rnn_condition_net(RecurNN *net)
{
  u32 mask = net->flags >> RNN_COND_USE_OFFSET;
  u32 m = net->generation % RNN_CONDITIONING_INTERVAL;
  if (((1u << m) & mask) == 0)
    return;
  rb_require_device("rnn_condition_net");
  rb_matrices_to_device(net);
  switch (m) {
  case RNN_COND_BIT_SCALE:
    rbk_scale(net->ih_weights, net->ih_size, WEIGHT_SCALE);
    rbk_scale(net->ho_weights, net->ho_size, WEIGHT_SCALE);
    break;
  case RNN_COND_BIT_ZERO:
    rbk_zero_small(net->ih_weights, net->ih_size);
    rbk_zero_small(net->ho_weights, net->ho_size);
    if (net->bptt) {
      rbk_zero_small(net->bptt->ih_momentum, net->ih_size);
      rbk_zero_small(net->bptt->ho_momentum, net->ho_size);
    }
    break;
  case RNN_COND_BIT_RAND: {
    int t = rb_rand_small_int(&net->rng, net->ih_size + net->ho_size);
    float damage = rb_cheap_gaussian_noise(&net->rng) * RANDOM_DAMAGE_FACTOR *
      net->h_size * net->bptt->learn_rate;
    if (t >= net->ih_size) {
      t -= net->ih_size;
      int col = t % net->o_size;
      if (col < net->output_size)
        rbk_add_at(net->ho_weights, t, damage);
    }
    else {
      int col = t % net->h_size;
      if (col >= 1 && col < net->hidden_size + 1)
        rbk_add_at(net->ih_weights, t, damage);
    }
  } break;
  case RNN_COND_BIT_TALL_POPPY:
    rbk_tall_poppy(net->ih_weights, net->ih_size, RNN_TALL_POPPY_THRESHOLD,
        RNN_TALL_POPPY_SCALE);
    break;
  case RNN_COND_BIT_LAWN_MOWER:
    rbk_clamp(net->ih_weights, net->ih_size, -RNN_LAWN_MOWER_THRESHOLD,
        RNN_LAWN_MOWER_THRESHOLD);
    break;
  }
  rb_weights_changed(net);
  sync_stream();
}

/* reference recur-nn-init.c:359-381 */
extern "C" void
rnn_set_momentum_values(RecurNN *net, float x)
{
  if (rb_have_device()) {
    rbk_fill(net->bptt->ho_momentum, net->ho_size, x);
    rbk_fill(net->bptt->ih_momentum, net->ih_size, x);
    if (net->bottom_layer)
      rbk_fill(net->bottom_layer->momentums,
          (size_t)net->bottom_layer->i_size * net->bottom_layer->o_size, x);
    sync_stream();
  }
  else {
    for (int i = 0; i < net->ho_size; i++) net->bptt->ho_momentum[i] = x;
    for (int i = 0; i < net->ih_size; i++) net->bptt->ih_momentum[i] = x;
  }
}

extern "C" void
rnn_set_aux_values(RecurNN *net, float x)
{
  if (!net->bptt->ih_aux)
    rb_die("recur-b200: rnn_set_aux_values needs RNN_NET_FLAG_AUX_ARRAYS");
  if (rb_have_device()) {
    rbk_fill(net->bptt->ho_aux, net->ho_size, x);
    rbk_fill(net->bptt->ih_aux, net->ih_size, x);
    if (net->bottom_layer && net->bottom_layer->aux)
      rbk_fill(net->bottom_layer->aux,
          (size_t)net->bottom_layer->i_size * net->bottom_layer->o_size, x);
    sync_stream();
  }
  else {
    for (int i = 0; i < net->ho_size; i++) net->bptt->ho_aux[i] = x;
    for (int i = 0; i < net->ih_size; i++) net->bptt->ih_aux[i] = x;
  }
}

/* reference recur-nn.c:885-904 */
extern "C" void
rnn_log_net(RecurNN *net)
{
  if (net->log == NULL)
    return;
  if (net->bptt) {
    if (rb_have_device())
      rb_net_pull(rb_net_of(net));
    float top_error = 0, hidden_error = 0;
    for (int i = 0; i < net->o_size; i++)
      top_error += fabsf(net->bptt->o_error[i]);
    for (int i = 0; i < net->h_size; i++)
      hidden_error += fabsf(net->bptt->h_error[i]);
    rnn_log_float(net, "output_error", top_error);
    rnn_log_float(net, "hidden_error", hidden_error);
  }
}
