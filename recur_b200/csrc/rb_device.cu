/* rb_device.cu — device context, memory kinds and the stream pools.
 *
 * Memory kinds (see include/recur-nn.h):
 *   matrix  weights / momentums / deltas / aux: CUDA managed memory, so the
 *           pointers in RecurNN/RecurNNBPTT stay valid for host code that
 *           pokes them while the kernels use the same addresses in HBM;
 *   mirror  pinned host copies of one stream's vectors;
 *   pool    plain device memory, stream-major (rb_internal.h).
 * Without a CUDA device the first two fall back to calloc so that the
 * host-side API (construction, initialisation, load/save) still works; any
 * compute call then aborts in rb_require_device().
 */
#include "rb_kernels.h"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

cudaStream_t rb_stream = 0;
static int rb_dev_state = -1; /* -1 unknown, 0 none, 1 ready */
static int rb_dev_ordinal = 0;
static uint64_t rb_launches = 0;

extern "C" void
rb_die(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
  fflush(stderr);
  abort();
}

#define CUDA_OR_DIE(call) do {                                          \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess)                                              \
      rb_die("recur-b200: %s failed at %s:%d: %s", #call, __FILE__,     \
          __LINE__, cudaGetErrorString(e_));                            \
  } while (0)

extern "C" void
rb_count_launch(int n)
{
  rb_launches += (uint64_t)n;
}

extern "C" uint64_t
rnn_b200_kernel_launches(void)
{
  return rb_launches;
}

static void
rb_probe_device(void)
{
  if (rb_dev_state >= 0)
    return;
  int n = 0;
  const char *env = getenv("RECUR_B200_DEVICE");
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    rb_dev_state = 0;
    return;
  }
  if (env)
    rb_dev_ordinal = atoi(env);
  if (rb_dev_ordinal < 0 || rb_dev_ordinal >= n)
    rb_dev_ordinal = 0;
  CUDA_OR_DIE(cudaSetDevice(rb_dev_ordinal));
  CUDA_OR_DIE(cudaStreamCreateWithFlags(&rb_stream, cudaStreamNonBlocking));
  rb_dev_state = 1;
}

extern "C" int
rb_have_device(void)
{
  rb_probe_device();
  return rb_dev_state == 1;
}

extern "C" void
rb_require_device(const char *what)
{
  if (!rb_have_device())
    rb_die("recur-b200: %s needs a CUDA device and none is usable "
        "(this library has no CPU compute path)", what);
  CUDA_OR_DIE(cudaSetDevice(rb_dev_ordinal));
}

extern "C" int
rnn_b200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int
rnn_b200_set_device(int ordinal)
{
  int n = rnn_b200_device_count();
  if (ordinal < 0 || ordinal >= n)
    return -1;
  if (rb_dev_state == 1 && ordinal != rb_dev_ordinal)
    rb_die("recur-b200: the device cannot change once nets exist");
  rb_dev_ordinal = ordinal;
  rb_probe_device();
  return 0;
}

extern "C" void
rnn_b200_synchronize(void)
{
  if (rb_have_device())
    CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
}

extern "C" void *
rnn_b200_stream(void)
{
  rb_probe_device();
  return (void *)rb_stream;
}

extern "C" const char *
rnn_b200_version(void)
{
  return "recur-b200 0.1 (sm_100a)";
}

/* ---- memory ---------------------------------------------------------------- */

extern "C" float *
rb_alloc_matrix(size_t n_floats)
{
  if (n_floats == 0)
    n_floats = 4;
  float *p = NULL;
  if (rb_have_device()) {
    CUDA_OR_DIE(cudaMallocManaged((void **)&p, n_floats * sizeof(float), cudaMemAttachGlobal));
    CUDA_OR_DIE(cudaMemsetAsync(p, 0, n_floats * sizeof(float), rb_stream));
    CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  }
  else {
    if (posix_memalign((void **)&p, 64, n_floats * sizeof(float)))
      rb_die("recur-b200: cannot allocate %zu floats", n_floats);
    memset(p, 0, n_floats * sizeof(float));
  }
  return p;
}

extern "C" void
rb_free_matrix(float *p)
{
  if (!p)
    return;
  if (rb_have_device()) {
    CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
    CUDA_OR_DIE(cudaFree(p));
  }
  else
    free(p);
}

/* Pinned host mirrors come out of slabs: one cudaHostAlloc per net costs tens
   of microseconds (and seconds once there are thousands of nets - an rnnca
   grid has one net per pixel); a slab is pinned once and carved up.  Each
   block carries its slab in a 64-byte header; a slab is returned to the driver
   when its last block is freed. */
#define MIRROR_SLAB_BYTES ((size_t)4 << 20)
#define MIRROR_HEADER 64

typedef struct RbSlab {
  char *base;
  size_t used, cap;
  long live;
  struct RbSlab *next, *prev;
} RbSlab;

static RbSlab *mirror_slabs = NULL; /* the head is the one being carved */

static RbSlab *
slab_new(size_t cap)
{
  RbSlab *sl = (RbSlab *)calloc(1, sizeof(RbSlab));
  if (!sl)
    rb_die("recur-b200: out of memory");
  CUDA_OR_DIE(cudaHostAlloc((void **)&sl->base, cap, cudaHostAllocDefault));
  sl->cap = cap;
  sl->next = mirror_slabs;
  if (mirror_slabs)
    mirror_slabs->prev = sl;
  mirror_slabs = sl;
  return sl;
}

extern "C" void *
rb_alloc_mirror(size_t bytes)
{
  if (bytes == 0)
    bytes = 16;
  void *p = NULL;
  if (rb_have_device()) {
    size_t need = ((bytes + 63) & ~(size_t)63) + MIRROR_HEADER;
    RbSlab *sl = mirror_slabs;
    if (need > MIRROR_SLAB_BYTES / 4) {
      /* big blocks (history rings of deep nets) get a slab of their own,
         linked behind the head so that carving goes on where it was */
      RbSlab *head = mirror_slabs;
      sl = slab_new(need);
      if (head) { /* move the new slab behind the old head */
        mirror_slabs = head;
        sl->next = head->next;
        if (head->next)
          head->next->prev = sl;
        head->next = sl;
        sl->prev = head;
        head->prev = NULL;
      }
    }
    else if (!sl || sl->cap - sl->used < need || sl->cap != MIRROR_SLAB_BYTES)
      sl = slab_new(MIRROR_SLAB_BYTES);
    char *blk = sl->base + sl->used;
    sl->used += need;
    sl->live++;
    *(RbSlab **)blk = sl;
    p = blk + MIRROR_HEADER;
    memset(p, 0, bytes);
  }
  else {
    if (posix_memalign(&p, 64, bytes))
      rb_die("recur-b200: cannot allocate %zu bytes", bytes);
    memset(p, 0, bytes);
  }
  return p;
}

extern "C" void
rb_free_mirror(void *p)
{
  if (!p)
    return;
  if (!rb_have_device()) {
    free(p);
    return;
  }
  RbSlab *sl = *(RbSlab **)((char *)p - MIRROR_HEADER);
  if (--sl->live > 0)
    return;
  if (sl == mirror_slabs && sl->cap == MIRROR_SLAB_BYTES) {
    sl->used = 0; /* the slab being carved: start over */
    return;
  }
  if (sl->prev)
    sl->prev->next = sl->next;
  else
    mirror_slabs = sl->next;
  if (sl->next)
    sl->next->prev = sl->prev;
  /* kernels may still be reading or writing the mirrors through the stream */
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  CUDA_OR_DIE(cudaFreeHost(sl->base));
  free(sl);
}

static void
prefetch(const float *p, size_t n_floats)
{
  if (p && n_floats)
    (void)cudaMemPrefetchAsync(p, n_floats * sizeof(float), rb_dev_ordinal, rb_stream);
}

/* Host code may have rewritten the matrices (initialisation, load, weight
   surgery): bring them back into HBM in one go rather than by page faults. */
extern "C" void
rb_matrices_to_device(RecurNN *net)
{
  if (!rb_have_device())
    return;
  RbNet *rn = rb_net_of(net);
  if (!rn->group->matrices_touched_by_host)
    return;
  rn->group->matrices_touched_by_host = 0;
  prefetch(net->ih_weights, net->ih_size);
  prefetch(net->ho_weights, net->ho_size);
  if (net->bptt) {
    RecurNNBPTT *b = net->bptt;
    prefetch(b->ih_momentum, net->ih_size);
    prefetch(b->ho_momentum, net->ho_size);
    prefetch(b->ih_delta, net->ih_size);
    prefetch(b->ho_delta, net->ho_size);
    if (net->flags & RNN_NET_FLAG_AUX_ARRAYS) {
      prefetch(b->ih_aux, net->ih_size);
      prefetch(b->ho_aux, net->ho_size);
    }
  }
  (void)cudaGetLastError();
}

/* Called by every host-side function about to read or write matrices. */
extern "C" void
rb_host_will_touch_matrices(RecurNN *net)
{
  if (!rb_have_device())
    return;
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  rb_net_of(net)->group->matrices_touched_by_host = 1;
}

/* ---- groups and pools ------------------------------------------------------ */

extern "C" RbGroup *
rb_group_new(const RbDims *d)
{
  RbGroup *g = (RbGroup *)calloc(1, sizeof(RbGroup));
  if (!g)
    rb_die("recur-b200: out of memory");
  g->d = *d;
  g->refs = 0;
  g->device = rb_dev_ordinal;
  g->matrices_touched_by_host = 1;
  return g;
}

static void
pool_free_device(RbPool *p)
{
  if (!rb_have_device())
    return;
  rb_tc_pool_release(p);
  rb_bottom_pool_release(p);
  cudaFree(p->X);
  cudaFree(p->Hd);
  cudaFree(p->Y);
  cudaFree(p->OE);
  cudaFree(p->E);
  cudaFree(p->partial);
  cudaFree(p->noise);
  cudaFree(p->pos);
  cudaFree(p->iota);
  cudaFree(p->sc);
  cudaFree(p->rng);
}

extern "C" void
rb_group_unref(RbGroup *g)
{
  if (--g->refs > 0)
    return;
  if (rb_have_device())
    CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  RbPool *p = g->pools;
  while (p) {
    RbPool *next = p->next;
    pool_free_device(p);
    free(p->used);
    free(p->pos_shadow);
    free(p);
    p = next;
  }
  free(g);
}

extern "C" RbPool *
rb_group_pool(RbGroup *g, int depth, int has_bptt)
{
  for (RbPool *p = g->pools; p; p = p->next) {
    if (p->depth == (depth > 0 ? depth : 1) && p->has_bptt == has_bptt)
      return p;
  }
  RbPool *p = (RbPool *)calloc(1, sizeof(RbPool));
  if (!p)
    rb_die("recur-b200: out of memory");
  p->group = g;
  p->depth = depth > 0 ? depth : 1;
  p->has_bptt = has_bptt;
  p->n_part = (g->d.i_size + 63) / 64;
  p->next = g->pools;
  g->pools = p;
  return p;
}

template <typename T>
static T *
dev_alloc_zero(size_t n)
{
  T *p = NULL;
  if (n == 0)
    n = 1;
  cudaError_t e = cudaMalloc((void **)&p, n * sizeof(T));
  if (e != cudaSuccess)
    rb_die("recur-b200: cudaMalloc of %zu bytes failed: %s", n * sizeof(T),
        cudaGetErrorString(e));
  CUDA_OR_DIE(cudaMemsetAsync(p, 0, n * sizeof(T), rb_stream));
  return p;
}

/* copy `blocks` blocks of old_cap rows into the front of blocks of new_cap rows */
template <typename T>
static void
regrow(T **arr, int blocks, int old_cap, int new_cap, size_t row)
{
  T *fresh = dev_alloc_zero<T>((size_t)blocks * new_cap * row);
  if (*arr && old_cap > 0) {
    for (int b = 0; b < blocks; b++) {
      CUDA_OR_DIE(cudaMemcpyAsync(fresh + (size_t)b * new_cap * row,
              *arr + (size_t)b * old_cap * row, (size_t)old_cap * row * sizeof(T),
              cudaMemcpyDeviceToDevice, rb_stream));
    }
  }
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  if (*arr)
    CUDA_OR_DIE(cudaFree(*arr));
  *arr = fresh;
}

extern "C" void
rb_pool_reserve(RbPool *p, int n_slots)
{
  if (n_slots <= p->cap)
    return;
  int new_cap = p->cap ? p->cap : 1;
  while (new_cap < n_slots)
    new_cap *= 2;
  if (new_cap > n_slots && n_slots > 64)
    new_cap = (n_slots + 63) & ~63; /* do not double big pools */
  p->used = (uint8_t *)realloc(p->used, new_cap);
  p->pos_shadow = (int *)realloc(p->pos_shadow, new_cap * sizeof(int));
  if (!p->used || !p->pos_shadow)
    rb_die("recur-b200: out of memory");
  memset(p->used + p->cap, 0, new_cap - p->cap);
  memset(p->pos_shadow + p->cap, 0, (new_cap - p->cap) * sizeof(int));
  if (rb_have_device()) {
    const RbDims *d = &p->group->d;
    int old = p->cap;
    regrow(&p->X, p->depth, old, new_cap, d->i_size);
    regrow(&p->Hd, 1, old, new_cap, d->h_size);
    regrow(&p->Y, 1, old, new_cap, d->o_size);
    regrow(&p->OE, 1, old, new_cap, d->o_size);
    if (p->has_bptt) {
      regrow(&p->E, p->depth + 1, old, new_cap, d->i_size);
      regrow(&p->partial, 1, old, new_cap, p->n_part);
    }
    regrow(&p->noise, 1, old, new_cap, d->h_size);
    regrow(&p->pos, 1, old, new_cap, 1);
    regrow(&p->sc, 1, old, new_cap, 1);
    regrow(&p->rng, 1, old, new_cap, 4);
    if (p->iota)
      CUDA_OR_DIE(cudaFree(p->iota));
    p->iota = dev_alloc_zero<int>(new_cap);
    rbk_fill_iota(p->iota, new_cap);
    CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  }
  p->cap = new_cap;
  p->x_planes_stale = 2;
}

extern "C" int
rb_pool_take_slot(RbPool *p)
{
  if (p->n_live == p->cap) {
    /* one more slot: grow by half, not by one block, or a grid of thousands
       of clones (one rnnca cell per pixel) copies the pool thousands of times */
    int want = p->cap + p->cap / 2;
    rb_pool_reserve(p, want > p->cap + 1 ? want : p->cap + 1);
  }
  int slot = -1;
  for (int i = p->free_hint; i < p->cap; i++) {
    if (!(p->used[i] & 1)) {
      slot = i;
      break;
    }
  }
  p->free_hint = slot + 1;
  const bool recycled = (p->used[slot] & 2) != 0;
  p->used[slot] = 3;
  p->n_live++;
  p->pos_shadow[slot] = 0;
  if (rb_have_device() && recycled) {
    p->x_planes_stale = 2;
    /* a recycled slot starts from zeroed state, like a calloc'ed net */
    const RbDims *d = &p->group->d;
    for (int r = 0; r < p->depth; r++)
      CUDA_OR_DIE(cudaMemsetAsync(p->X + ((size_t)r * p->cap + slot) * d->i_size, 0,
              d->i_size * sizeof(float), rb_stream));
    CUDA_OR_DIE(cudaMemsetAsync(p->Hd + (size_t)slot * d->h_size, 0,
            d->h_size * sizeof(float), rb_stream));
    CUDA_OR_DIE(cudaMemsetAsync(p->Y + (size_t)slot * d->o_size, 0,
            d->o_size * sizeof(float), rb_stream));
    CUDA_OR_DIE(cudaMemsetAsync(p->OE + (size_t)slot * d->o_size, 0,
            d->o_size * sizeof(float), rb_stream));
    CUDA_OR_DIE(cudaMemsetAsync(p->pos + slot, 0, sizeof(int), rb_stream));
    CUDA_OR_DIE(cudaMemsetAsync(p->sc + slot, 0, sizeof(RbScalars), rb_stream));
  }
  return slot;
}

extern "C" void
rb_pool_release_slot(RbPool *p, int slot)
{
  if (slot >= 0 && slot < p->cap && (p->used[slot] & 1)) {
    p->used[slot] &= ~1;
    p->n_live--;
    if (slot < p->free_hint)
      p->free_hint = slot;
  }
}

/* ---- per-class kernel timing ------------------------------------------------ */

#define RB_PROF_CLASSES 8
static int prof_on = 0;
static cudaEvent_t *prof_ev = NULL; /* pairs */
static int *prof_cls = NULL;
static size_t prof_n = 0, prof_cap = 0;

extern "C" void
rnn_b200_profile_enable(int on)
{
  if (rb_have_device())
    CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  prof_on = on && rb_have_device();
  prof_n = 0;
}

extern "C" const char *
rnn_b200_profile_class_name(int cls)
{
  static const char *names[RB_PROF_CLASSES] = {"forward", "bptt_chain", "weight_grad", "update",
    "top_layer", "output_layer", "ho_delta", "small_kernels"};
  return (cls >= 0 && cls < RB_PROF_CLASSES) ? names[cls] : NULL;
}

extern "C" int
rb_prof_active(void)
{
  return prof_on;
}

extern "C" void
rb_prof_begin(int cls)
{
  if (!prof_on)
    return;
  if (prof_n == prof_cap) {
    size_t cap = prof_cap ? prof_cap * 2 : 1024;
    prof_ev = (cudaEvent_t *)realloc(prof_ev, cap * 2 * sizeof(cudaEvent_t));
    prof_cls = (int *)realloc(prof_cls, cap * sizeof(int));
    for (size_t i = prof_cap * 2; i < cap * 2; i++)
      CUDA_OR_DIE(cudaEventCreate(&prof_ev[i]));
    prof_cap = cap;
  }
  prof_cls[prof_n] = cls;
  CUDA_OR_DIE(cudaEventRecord(prof_ev[prof_n * 2], rb_stream));
}

extern "C" void
rb_prof_end(int cls)
{
  (void)cls;
  if (!prof_on)
    return;
  CUDA_OR_DIE(cudaEventRecord(prof_ev[prof_n * 2 + 1], rb_stream));
  prof_n++;
}

extern "C" int
rnn_b200_profile_read(double *ms, uint64_t *launches, int max_classes)
{
  for (int c = 0; c < max_classes; c++) {
    ms[c] = 0;
    launches[c] = 0;
  }
  if (!rb_have_device())
    return RB_PROF_CLASSES;
  CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
  for (size_t i = 0; i < prof_n; i++) {
    float t = 0;
    CUDA_OR_DIE(cudaEventElapsedTime(&t, prof_ev[i * 2], prof_ev[i * 2 + 1]));
    int c = prof_cls[i];
    if (c < max_classes) {
      ms[c] += t;
      launches[c]++;
    }
  }
  return RB_PROF_CLASSES;
}
