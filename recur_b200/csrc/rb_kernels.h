/* rb_kernels.h — launchers of the sm_100a kernels (rb_kernels.cu, rb_tc.cu).
 * All launch asynchronously on the library stream. */
#ifndef RB_KERNELS_H
#define RB_KERNELS_H

#include "rb_internal.h"
#include <cuda_runtime.h>

/* A batch of streams of one pool, as the kernels see it. */
typedef struct RbView {
  RbDims d;
  int cap, depth, n_part;
  float *X, *Hd, *Y, *OE, *E, *partial, *noise;
  int *pos;
  RbScalars *sc;
  uint64_t *rng;
  const int *slots; /* device: n pool slots */
  int n;
  int contiguous;   /* slots[j] == slots[0] + j */
  int base;         /* slots[0] when contiguous */
  const float *Wih; /* [i_size][h_size] */
  const float *Who; /* [h_size][o_size] */
  int activation;
  RbPool *pool;     /* host-side owner (not used by kernels) */
  void *p2p;        /* fused gradient exchange state of the batch, or NULL */
  /* bottom layer, or CIE == NULL */
  float *BI, *BO, *BN, *CIE, *BR;
  int bl_i, bl_o;
} RbView;

/* One operand of the tensor engine as FP16 hi/lo planes (rb_split.cuh): row r
   of the operand starts at r * pitch halves.  The values were multiplied by
   `scale` before the split; for the error chain that scale is chosen per walk
   and lives on the device (`scale_dev`, written by the kernel that makes the
   planes of E(0)), otherwise scale_dev is NULL and `scale` holds it. */
typedef struct RbPlanes {
  unsigned short *hi, *lo;
  int pitch;
  float scale;
  const float *scale_dev;
} RbPlanes;

/* device accumulators of the text-predict report sums */
typedef struct RbCharAccum {
  double error;
  double entropy;
  long long correct;
  long long count;
} RbCharAccum;

#ifdef __cplusplus
extern "C" {
#endif

extern cudaStream_t rb_stream;
void rb_view_of_net(RbNet *rn, RbView *v);
void rb_count_launch(int n);
void rb_note_walk_kernel(const char *name);
int rb_prof_active(void);
void rb_prof_begin(int cls);
void rb_prof_end(int cls);
enum { RB_PROF_FWD = 0, RB_PROF_CHAIN = 1, RB_PROF_DW = 2, RB_PROF_UPDATE = 3,
       RB_PROF_TOP = 4, RB_PROF_OUT = 5, RB_PROF_HO = 6, RB_PROF_SMALL = 7 };

void rbk_advance(const RbView *v);
void rbk_fill_iota(int *iota, int n);
void rbk_set_one_hot(const RbView *v, const u8 *hot_dev);
void rbk_set_inputs(const RbView *v, const float *inputs_dev);
void rbk_text_symbols(const u8 *text_dev, int len, int i, int spacing, int n,
    u8 *cur_dev, u8 *next_dev);
void rbk_forward(const RbView *v, float presynaptic_noise); /* a3..a5 */
void rbk_prepare_x(const RbView *v);
void rbk_forward_core(const RbView *v, float presynaptic_noise);
int rbk_step_begin_usable(const RbView *v);
void rbk_step_begin(const RbView *v, const u8 *text_dev, int len, int pos, int spacing,
    u8 *cur_dev, u8 *next_dev, const RbPlanes *X, int advance);
void rbk_step_begin_on(cudaStream_t stream, const RbView *v, const u8 *text_dev, int len, int pos,
    int spacing, u8 *cur_dev, u8 *next_dev, const RbPlanes *X, int advance);
void rb_mark_pre_update(void);
void rbk_output(const RbView *v);
void rbk_rnnca_gather(const RbView *v, const u8 *frame_dev, int width, int height,
    const int *off_y_dev, int len_y, const int *off_c_dev, int len_c, int len_pos, int edges);
void rbk_rnnca_emit(const RbView *v, u8 *frame_out_dev, int width, int height);
int rbk_rnnca_cells_usable(const RbView *v);
void rbk_rnnca_cells(const RbView *v, const u8 *frame_dev, u8 *frame_out_dev, int width,
    int height, const int *off_y_dev, int len_y, const int *off_c_dev, int len_c, int len_pos,
    int edges);
int rbk_walk_single_usable(const RbView *v);
int rbk_walk_resident_usable(const RbView *v);
void rbk_walk_resident(const RbView *v, const RbPlanes *E);
void rbk_dw_fma(const RbView *v, float *ih_delta, int accumulate);
int rbk_opinion_single_usable(const RbView *v);
void rbk_calculate_single(const RbView *v, const float *o_error_host, float lr, float mef,
    int adaptive, float *ho_w, float *ho_mom, float *ih_w, float *ih_mom, float *ih_delta,
    float momentum, float momentum_weight, RbScalars *sc_host);
void rbk_opinion_single(const RbView *v, const float *hidden_in, const float *inputs_in,
    float *input_layer_out, float *hidden_out, float *output_out);
/* split-K partial sums of a forward GEMM: [splits][rows of `pitch` floats] */
typedef struct RbFwdPartials {
  const float *part;
  size_t split_stride; /* floats between the planes of two splits */
  int pitch;
  int splits;
  int use_noise;
} RbFwdPartials;
int rbk_output_takes_partials(const RbView *v, int splits);
void rbk_output_from_partials(const RbView *v, const RbFwdPartials *fp);
void rbk_output_pipeline(int on);
void rbk_request_fused_loss(const u8 *target_dev, float *err_dev, int *winner_dev,
    RbCharAccum *accum_dev, RbCharAccum *snapshot_host, int reset);
int rbk_fused_loss_done(void);
void rbk_chain_decide(const RbView *v, int k);
void rbk_softmax_error(const RbView *v, const u8 *target_dev, float *err_dev,
    int *winner_dev, RbCharAccum *accum_dev);              /* a6 */
void rbk_top_layer(const RbView *v, float *ho_delta, int accumulate,
    const RecurErrorRange *ranges_dev, int n_ranges);       /* a7..a9 */
void rbk_top_layer_begin(const RbView *v, float *ho_delta, int accumulate,
    const RecurErrorRange *ranges_dev, int n_ranges);
void rbk_top_layer_join(void);
void rbk_top_layer_defer_ho_delta(int on);
void rbk_top_layer_ho_delta_now(void);
void rbk_top_layer_mark(void);
void rbk_request_fused_top(int on);
int rbk_top_was_fused(const RbView *v);
void rbk_bptt(const RbView *v, float *ih_delta, int accumulate); /* a10, a11 */
void rbk_set_params(const RbView *v, const float *lr_dev, const float *mef_dev, int adaptive);
void rbk_mask_streams(const RbView *v, const u8 *active_dev);
void rbk_set_params_scalar(const RbView *v, float lr, float mef, int adaptive);
void rbk_sgd_top_apply(const RbView *v, float *ho_weights, float *ho_momentum,
    float rate, float momentum, float momentum_weight);     /* a14 (recur-nn.c:941-964) */

/* a13: one optimiser step over `size` elements.  rate_scale_dev (may be NULL)
   points at a device float multiplied into the rate (ih_scale, a14). */
void rbk_apply_learning(int method, float *weights, const float *delta,
    float *momentums, float *aux, int size, float rate, float momentum,
    float momentum_weight, const float *rate_scale_dev);

/* a15 / a16 and friends */
void rbk_scale(float *a, int n, float s);
void rbk_zero_small(float *a, int n);
void rbk_clamp(float *a, int n, float lo, float hi);
void rbk_tall_poppy(float *a, int n, float threshold, float scale);
void rbk_add_at(float *a, int index, float value);
void rbk_axpy(float *dst, const float *src, int n, float s, const float *s_dev);
void rbk_fill(float *a, size_t n, float value);
void rbk_abs_sum(const float *a, int n, float *out_dev);
void rbk_gen_noise(const RbView *v, float deviation, int first_col, int n_cols);

/* bottom layer (rb_bottom.cu) */
void rb_bottom_attach(RbView *v, RecurNN *net);
void rb_bottom_forward(const RbView *v, RecurNN *net, const float *shared_inputs_or_null,
    float presynaptic_noise);
void rb_bottom_one_hot(const RbView *v, const u8 *hot_dev);
void rb_bottom_set_inputs(const RbView *v, const float *inputs_dev, int input_size);
void rb_bottom_backward(const RbView *v, RecurNN *net, int accumulate);
void rb_bottom_pool_release(RbPool *p);

/* tensor-core engine (rb_tc.cu) */
int rb_tc_usable(const RbView *v);
void rb_tc_forward(RbPool *p, const RbView *v, float presynaptic_noise);
void rb_tc_forward_core(RbPool *p, const RbView *v, float presynaptic_noise);
void rb_tc_x_planes(RbPool *p, RbPlanes *X);
void rb_tc_top_and_bptt(RbPool *p, const RbView *v, float *ho_delta, float *ih_delta,
    int accumulate);
void rb_tc_pool_release(RbPool *p);
void rb_tc_defer_delta_reduce(int on);
void rb_tc_materialise_delta(RbPool *p, float *ih_delta);
int rb_tc_fused_update(RbPool *p, RecurNN *net, int method, float momentum,
    float momentum_weight);
void rb_tc_planes_current(RbPool *p);

/* fused split-K reduction + all-reduce over peer memory (rb_p2p.cu) */
void *rb_p2p_new(size_t n_floats);
int rb_p2p_export(void *state, void *handles_out);
int rb_p2p_attach(void *state, const void *all_handles, int rank, int n_ranks);
int rb_p2p_ready(void *state);
void rb_p2p_exchange(void *state, const float *partial, int splits, int ih_size, int ho_size,
    const float *ho_delta);
void rb_p2p_result(void *state, const float **result, const unsigned int **flags,
    unsigned int *epoch, int *n);
void rb_p2p_copy_out(void *state, float *ih_delta);
void rb_p2p_delete(void *state);

#ifdef __cplusplus
}
#endif
#endif
