/* rb_dispatch.cu — picks the matrix engine for the three big contractions.
 *
 * FMA engine (rb_kernels.cu): exact FP32 on CUDA cores, any batch.
 * Tensor engine (rb_tc.cu): tcgen05 3xTF32; needs >= 64 streams, a multiple
 * of 32, in one contiguous run of pool slots that advance in lockstep.
 */
#include "rb_kernels.h"
#include "rb_host.h"

extern "C" int rb_engine(void);

static cudaStream_t side_stream = NULL;
static cudaEvent_t ev_pre_update = NULL, ev_side_done = NULL;
static int pre_update_valid = 0;

/* called in front of a weight update: the point in the stream up to which a
   following step's input rows depend on earlier work */
extern "C" void
rb_mark_pre_update(void)
{
  if (!ev_pre_update)
    cudaEventCreateWithFlags(&ev_pre_update, cudaEventDisableTiming);
  cudaEventRecord(ev_pre_update, rb_stream);
  pre_update_valid = 1;
}

extern "C" void
rb_weights_changed(RecurNN *net)
{
  rb_net_of(net)->group->weights_version++;
}

static int
lockstep(const RbView *v)
{
  const int *pos = v->pool->pos_shadow;
  for (int j = 1; j < v->n; j++)
    if (pos[v->base + j] != pos[v->base])
      return 0;
  return 1;
}

static int
use_tensor_engine(const RbView *v)
{
  int engine = rb_engine();
  if (engine == 1)
    return 0;
  int ok = rb_tc_usable(v) && lockstep(v);
  if (engine == 2 && !ok)
    rb_die("recur-b200: the tensor engine was forced but this batch (%d streams%s) cannot "
        "use it", v->n, v->contiguous ? "" : ", scattered slots");
  return ok;
}

extern "C" void
rb_forward_dispatch(const RbView *v, float noise)
{
  if (use_tensor_engine(v))
    rb_tc_forward(v->pool, v, noise);
  else {
    v->pool->x_planes_stale = 2;
    rbk_forward(v, noise);
  }
}

/* fused start of a text step + forward (rb_batch.cu's character steps) */
extern "C" void
rb_char_forward_dispatch(const RbView *v, const u8 *text_dev, int len, int pos, int spacing,
    u8 *cur_dev, u8 *next_dev, float noise, int advance, int continues)
{
  int tensor = use_tensor_engine(v);
  if (!rbk_step_begin_usable(v)) {
    if (text_dev)
      rbk_text_symbols(text_dev, len, pos, spacing, v->n, cur_dev, next_dev);
    if (advance)
      rbk_advance(v);
    rbk_set_one_hot(v, cur_dev);
    rb_forward_dispatch(v, noise);
    return;
  }
  RbPlanes xp, *X = NULL;
  if (tensor) {
    rb_tc_x_planes(v->pool, &xp);
    X = &xp;
  }
  else
    v->pool->x_planes_stale = 2;
  if (text_dev && advance && continues && pre_update_valid && !rb_prof_active()) {
    /* The next position's input rows need the last forward pass and the text,
       not the weights: with the text resident the kernel runs on a side stream
       next to the update of the step before (everything queued up to the point
       marked in front of that update is waited for).  `continues`: the caller
       vouches that nothing but that update was queued since. */
    if (!side_stream) {
      cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&ev_side_done, cudaEventDisableTiming);
    }
    cudaStreamWaitEvent(side_stream, ev_pre_update, 0);
    rbk_step_begin_on(side_stream, v, text_dev, len, pos, spacing, cur_dev, next_dev, X,
        advance);
    cudaEventRecord(ev_side_done, side_stream);
    cudaStreamWaitEvent(rb_stream, ev_side_done, 0);
  }
  else
    rbk_step_begin(v, text_dev, len, pos, spacing, cur_dev, next_dev, X, advance);
  pre_update_valid = 0;
  if (tensor)
    rb_tc_forward_core(v->pool, v, noise);
  else
    rbk_forward_core(v, noise);
}

static int last_bptt_tensor = 0;
static const char *last_walk_kernel = "";

/* which kernel walked the ring in the most recent BPTT call (tests and
   smoke() assert that the path they mean to check is the one that ran) */
extern "C" void
rb_note_walk_kernel(const char *name)
{
  last_walk_kernel = name;
}

extern "C" const char *
rnn_b200_last_walk_kernel(void)
{
  return last_walk_kernel;
}

extern "C" int
rb_last_bptt_used_tensor_engine(void)
{
  return last_bptt_tensor;
}

/* a7..a11 for a batch without error ranges */
extern "C" void
rb_top_and_bptt_dispatch(const RbView *v, float *ho_delta, float *ih_delta, int accumulate)
{
  last_bptt_tensor = use_tensor_engine(v) && v->pool->has_bptt;
  if (last_bptt_tensor) {
    rb_tc_top_and_bptt(v->pool, v, ho_delta, ih_delta, accumulate);
  }
  else {
    rbk_top_layer(v, ho_delta, accumulate, NULL, 0);
    rbk_bptt(v, ih_delta, accumulate);
  }
}
