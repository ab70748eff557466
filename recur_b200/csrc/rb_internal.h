/* rb_internal.h — private structures of librecur_b200.so.
 *
 * Vocabulary (follows the reference): a *net* is one RecurNN struct; nets
 * made by rnn_clone / rnn_new_training_set without OWN_WEIGHTS borrow their
 * parent's weights (recur-nn-init.c:296-350) and are the parallel *streams*
 * of the synchronic mini-batch.  Here every family of nets that share weights
 * is a *group*, and the per-stream state of a group (history ring, hidden
 * and output activations, errors) lives in device *pools*, one per BPTT
 * depth, laid out stream-major so that a batch of streams is a matrix.
 */
#ifndef RB_INTERNAL_H
#define RB_INTERNAL_H

#include <stdint.h>
#include <stddef.h>
#include "../../include/recur_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RB_MAGIC 0x52423230u /* "RB20" */

/* per-stream scalars kept on the device (one per pool slot) */
typedef struct RbScalars {
  float top_raw;     /* sum |h_error| after the top layer (recur-nn.c:719) */
  float top_scaled;  /* after the soft clip (recur-nn.c:720) */
  float err_sum;     /* error_sum of the last executed BPTT step */
  float ih_scale;
  float mef;         /* bptt->min_error_factor */
  float lr;          /* bptt->learn_rate of this stream */
  float cum_error;
  float min_sum;     /* min_error_sum (recur-nn.c:321) */
  float max_sum;     /* max_error_sum (recur-nn.c:318) */
  float hidden_sum;
  float hidden_mag;
  int hidden_zeros;
  int live;          /* still walking back through the ring */
  int t_left;        /* value of the reference's loop counter t at exit */
  int n_steps;       /* BPTT steps executed: E[1..n_steps] are valid */
  int adaptive;      /* RNN_NET_FLAG_BPTT_ADAPTIVE_MIN_ERROR */
} RbScalars;

typedef struct RbDims {
  int i_size, h_size, o_size;
  int input_size, hidden_size, output_size;
} RbDims;

struct RbGroup;

/* Device-side state of all streams of one group with one BPTT depth. */
typedef struct RbPool {
  struct RbGroup *group;
  struct RbPool *next;
  int depth;      /* ring slots; nets without BPTT use a ring of one */
  int has_bptt;
  int cap;        /* slots allocated */
  int n_live;     /* slots in use */
  uint8_t *used;  /* host: cap flags; bit 0 in use, bit 1 has been used since the arrays were
                     (re)allocated zeroed */
  int free_hint;  /* no free slot below this index */
  int *pos_shadow;/* host copy of pos[] */
  /* device arrays (slot-major) */
  float *X;       /* [depth][cap][i_size]   history ring of input rows */
  float *Hd;      /* [cap][h_size]          hidden_layer */
  float *Y;       /* [cap][o_size]          output_layer */
  float *OE;      /* [cap][o_size]          bptt->o_error */
  float *E;       /* [depth+1][cap][i_size] E[0] top error, E[k+1] after step k */
  float *partial; /* [cap][n_part]          per column-block sums of e^2 */
  float *noise;   /* [cap][h_size]          presynaptic noise rows */
  int *pos;       /* [cap] ring position == bptt->index semantics */
  int *iota;      /* [cap] 0,1,2.. (slot lists for contiguous runs) */
  RbScalars *sc;  /* [cap] */
  uint64_t *rng;  /* [cap][4] per-stream PRNG state when noise runs on device */
  int n_part;
  void *tc;        /* tensor-core engine state (rb_tc.cu), made on first use */
  /* The tensor engine reads the ring through operand planes written next to X
     by its own forward pass.  Anything else that rewrites ring rows marks
     them stale: 1 = the current row only (inputs set, forward not yet run:
     rb_tc_forward re-splits that row anyway), 2 = any row (push, forget,
     FMA / per-net forward, recycled slot, regrown pool).  rb_tc_top_and_bptt
     re-splits the whole ring before the weight gradient when it finds 2. */
  int x_planes_stale;
  /* bottom layer (recur-nn.c:88-103,377-382,751-764), made on first use */
  float *BI;      /* [cap][bl_i] bottom inputs [1 | inputs] per stream */
  float *BO;      /* [cap][bl_o] bottom outputs (pre-ReLU) */
  float *BN;      /* [cap][bl_o] bottom presynaptic noise */
  float *CIE;     /* [cap][bl_o] cumulative input error of the current walk */
  float *BR;      /* [cap][bl_o] the shared accumulator as it stood after each stream */
  int bl_i, bl_o, bl_cap;
} RbPool;

typedef struct RbGroup {
  RbDims d;
  int refs;
  int device;
  RbPool *pools;
  int matrices_touched_by_host; /* prefetch to the device before next launch */
  uint64_t weights_version;     /* bumped whenever ih/ho weights may have changed */
} RbGroup;

/* Hidden header in front of every RecurNN this library hands out. */
typedef struct RbNet {
  uint32_t magic;
  RbGroup *group;
  RbPool *pool;
  int slot;
  int dev_ahead;        /* a batch call left newer state on the device than in the mirrors */
  /* allocations owned by this net */
  float *own_weights;   /* managed: ih | ho, or NULL when borrowed */
  float *own_momentums; /* managed */
  float *own_deltas;    /* managed: ih_delta | ho_delta | ih_delta_tmp */
  float *own_aux;       /* managed */
  size_t n_weights, n_momentums, n_deltas, n_aux;
  void *mirror_net;     /* pinned: input_layer(no bptt) | hidden | output */
  void *mirror_bptt;    /* pinned: o_error | i_error | h_error | history */
  RecurNN pub;          /* what the caller sees */
} RbNet;

RbNet *rb_net_of(RecurNN *net); /* aborts on a foreign pointer */

/* ---- memory (rb_device.cu) ------------------------------------------------ */
int rb_have_device(void);        /* 1 if CUDA works, never aborts */
void rb_require_device(const char *what); /* aborts with a message if not */
float *rb_alloc_matrix(size_t n_floats);   /* managed (or calloc w/o device), zeroed */
void rb_free_matrix(float *p);
void *rb_alloc_mirror(size_t bytes);       /* pinned host (or calloc), zeroed */
void rb_free_mirror(void *p);
void rb_matrices_to_device(RecurNN *net);  /* prefetch after host edits */
void rb_host_will_touch_matrices(RecurNN *net); /* sync + mark */

void rb_die(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));

/* ---- pools (rb_device.cu) ------------------------------------------------- */
RbGroup *rb_group_new(const RbDims *d);
void rb_group_unref(RbGroup *g);
RbPool *rb_group_pool(RbGroup *g, int depth, int has_bptt);
void rb_pool_reserve(RbPool *p, int n_slots);
int rb_pool_take_slot(RbPool *p);
void rb_pool_release_slot(RbPool *p, int slot);

/* ---- host-side pieces in C (rb_init.c, rb_io.c, rb_misc.c) ---------------- */
void rb_init_rand64_maybe_randomly(rand_ctx *ctx, u64 seed);
u64 rb_rand64(rand_ctx *x);
float rb_cheap_gaussian_noise(rand_ctx *ctx);
double rb_rand_double(rand_ctx *ctx);
int rb_rand_small_int(rand_ctx *ctx, int cap);

#ifdef __cplusplus
}
#endif
#endif
