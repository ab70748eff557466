/* rb_bottom.cu — the optional layer below the recurrent one.
 *
 * Forward (recur-nn.c:88-103): outputs = [1 | inputs] . W, presynaptic noise
 * on outputs[1..], real_inputs = relu(outputs).  Backward: the BPTT walk adds
 * every step's input-row errors into a cumulative vector (recur-nn.c:377-382,
 * done where the chain kernels produce those errors), which is scaled by
 * ih_scale^2 when the walk's gradient is clipped (395-401) and then turned
 * into weight deltas against the CURRENT bottom inputs (751-756).
 *
 * Reference quirks kept on purpose (SURVEY.md §8a notes): all clones share
 * one RecurExtraLayer, so the cumulative vector (bottom_layer->o_error) is one
 * accumulator that grows stream after stream and step after step until
 * rnn_bptt_clear_deltas zeroes it, and a clipped stream rescales the whole
 * accumulator, earlier streams' share included.  The batch kernels reproduce
 * that by walking the streams in order.
 */
#include "rb_kernels.h"
#include "rb_host.h"
#include "rb_comm.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CUDA_OR_DIE(call) do {                                          \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess)                                              \
      rb_die("recur-b200: %s failed at %s:%d: %s", #call, __FILE__,     \
          __LINE__, cudaGetErrorString(e_));                            \
  } while (0)

#define LAUNCH_CHECK(name) do {                                         \
    cudaError_t e_ = cudaGetLastError();                                \
    if (e_ != cudaSuccess)                                              \
      rb_die("recur-b200: launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    rb_count_launch(1);                                                 \
  } while (0)

static float *
dzero(size_t n)
{
  float *p = NULL;
  CUDA_OR_DIE(cudaMalloc((void **)&p, (n ? n : 1) * sizeof(float)));
  CUDA_OR_DIE(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(float), rb_stream));
  return p;
}

extern "C" void
rb_bottom_pool_release(RbPool *p)
{
  cudaFree(p->BI);
  cudaFree(p->BO);
  cudaFree(p->BN);
  cudaFree(p->CIE);
  cudaFree(p->BR);
  p->BI = p->BO = p->BN = p->CIE = p->BR = NULL;
}

/* per-stream bottom arrays live beside the pool; (re)made when the pool grew */
static void
ensure(RbPool *p, const RecurExtraLayer *bl)
{
  if (p->BI && p->bl_i == bl->i_size && p->bl_o == bl->o_size && p->bl_cap >= p->cap)
    return;
  if (p->BI) {
    CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
    rb_bottom_pool_release(p);
  }
  p->bl_i = bl->i_size;
  p->bl_o = bl->o_size;
  p->BI = dzero((size_t)p->cap * p->bl_i);
  p->BO = dzero((size_t)p->cap * p->bl_o);
  p->BN = dzero((size_t)p->cap * p->bl_o);
  p->CIE = dzero((size_t)p->cap * p->bl_o);
  p->BR = dzero((size_t)p->cap * p->bl_o);
  p->bl_cap = p->cap;
}

extern "C" void
rb_bottom_attach(RbView *v, RecurNN *net)
{
  v->BI = v->BO = v->BN = v->CIE = v->BR = NULL;
  v->bl_i = v->bl_o = 0;
  if (!net->bottom_layer)
    return;
  ensure(v->pool, net->bottom_layer);
  v->BI = v->pool->BI;
  v->BO = v->pool->BO;
  v->BN = v->pool->BN;
  v->CIE = v->pool->CIE;
  v->BR = v->pool->BR;
  v->bl_i = v->pool->bl_i;
  v->bl_o = v->pool->bl_o;
}

__device__ __forceinline__ int
slot_at(const RbView &v, int j)
{
  return v.contiguous ? v.base + j : v.slots[j];
}

/* one_hot_opinion on a net with a bottom layer writes inputs[hot] = 1 into
   the layer's input vector and rnn_opinion then sets inputs[0] = 1
   (charmodel-helpers.h:19-31, recur-nn.c:91) */
__global__ void
k_bottom_one_hot(RbView v, const u8 *hot)
{
  int s = slot_at(v, blockIdx.x);
  int h = hot[blockIdx.x];
  float *in = v.BI + (size_t)s * v.bl_i;
  for (int i = threadIdx.x; i < v.bl_i; i += blockDim.x)
    in[i] = (i == 0 || i == h) ? 1.0f : 0.0f;
}

__global__ void
k_bottom_set_inputs(RbView v, const float *inputs, int input_size)
{
  int s = slot_at(v, blockIdx.x);
  float *in = v.BI + (size_t)s * v.bl_i;
  const float *src = inputs + (size_t)blockIdx.x * input_size;
  for (int i = threadIdx.x; i < v.bl_i; i += blockDim.x)
    in[i] = (i == 0) ? 1.0f : (i <= input_size ? src[i - 1] : 0.0f);
}

/* shared_inputs != NULL: the per-net call, the caller's vector is the
   layer's own (shared) input array; the layer's output array is kept up to
   date too, as the reference leaves it */
__global__ void __launch_bounds__(256)
k_bottom_forward(RbView v, const float *__restrict__ W, const float *shared_inputs,
    float *shared_outputs, int use_noise)
{
  extern __shared__ float in_s[];
  const int s = slot_at(v, blockIdx.x);
  float *in = v.BI + (size_t)s * v.bl_i;
  for (int i = threadIdx.x; i < v.bl_i; i += blockDim.x) {
    float x = shared_inputs ? shared_inputs[i] : in[i];
    if (shared_inputs)
      in[i] = x;
    in_s[i] = x;
  }
  __syncthreads();
  int p = v.pos[s];
  float *real_inputs = v.X + ((size_t)p * v.cap + s) * v.d.i_size + v.d.hidden_size + 1;
  for (int x = threadIdx.x; x < v.bl_o; x += blockDim.x) {
    float acc = 0.0f;
    for (int y = 0; y < v.bl_i; y++) {
      float a = in_s[y];
      if (a != 0.0f)
        acc += a * W[(size_t)y * v.bl_o + x];
    }
    if (use_noise && x >= 1 && x < v.d.input_size)
      acc += v.BN[(size_t)s * v.bl_o + x];
    v.BO[(size_t)s * v.bl_o + x] = acc;
    if (shared_outputs)
      shared_outputs[x] = acc;
    if (x < v.d.input_size)
      real_inputs[x] = (acc > 0.0f) ? acc : 0.0f;
  }
}

/* noise rows for the bottom outputs: drawn before the hidden layer's, from
   the same per-stream generator (recur-nn.c:97 then :120) */
__global__ void
k_bottom_noise(RbView v, float deviation)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= v.n)
    return;
  int s = slot_at(v, j);
  uint64_t a = v.rng[s * 4 + 0], b = v.rng[s * 4 + 1], c = v.rng[s * 4 + 2], d = v.rng[s * 4 + 3];
  float *row = v.BN + (size_t)s * v.bl_o;
  for (int i = 1; i < v.d.input_size; i++) {
    long long acc = 0;
    for (int draw = 0; draw < 3; draw++) {
      uint64_t e = a - ((b << 7) | (b >> 57));
      a = b ^ ((c << 13) | (c >> 51));
      b = c + ((d << 37) | (d >> 27));
      c = d + e;
      d = e + a;
      acc += (long long)(d & 0xffff) + (long long)((d >> 16) & 0xffff) +
        (long long)((d >> 32) & 0xffff) + (long long)(d >> 48);
    }
    row[i] = ((float)(acc - 0xffff * 6) / (0xffff)) * deviation;
  }
  v.rng[s * 4 + 0] = a;
  v.rng[s * 4 + 1] = b;
  v.rng[s * 4 + 2] = c;
  v.rng[s * 4 + 3] = d;
}

/* the shared accumulator, stream after stream (one block, threads over the
   layer's outputs; the walk over streams is serial by construction) */
__global__ void
k_bottom_accumulate(RbView v, float *shared_o_error)
{
  for (int x = threadIdx.x; x < v.bl_o; x += blockDim.x) {
    float running = shared_o_error[x];
    for (int j = 0; j < v.n; j++) {
      int s = slot_at(v, j);
      const RbScalars sc = v.sc[s];
      if (sc.adaptive & 2) { /* stream sat this step out */
        v.BR[(size_t)s * v.bl_o + x] = 0.0f;
        continue;
      }
      if (x < v.d.input_size)
        running += v.CIE[(size_t)s * v.bl_o + x];
      if (sc.err_sum > ERROR_GAIN_CEILING * sc.top_scaled && x < v.d.input_size)
        running *= sc.ih_scale * sc.ih_scale;
      v.BR[(size_t)s * v.bl_o + x] = running;
    }
    shared_o_error[x] = running;
  }
}

/* Several GPUs: the accumulator's walk over the streams goes on through the
   ranks in rank order, as if the reference had all the streams in one array.
   What a rank's streams do to the accumulator is affine in the value they
   start from, running -> A * running + B per output; every rank works out its
   own (A, B), the ranks swap them, and each then knows the value its first
   stream starts from and the value the last rank's last stream ends with. */
__global__ void
k_bottom_affine(RbView v, float *ab /* [2][bl_o]: this rank's A, B */)
{
  for (int x = threadIdx.x; x < v.bl_o; x += blockDim.x) {
    float A = 1.0f, B = 0.0f;
    for (int j = 0; j < v.n; j++) {
      int s = slot_at(v, j);
      const RbScalars sc = v.sc[s];
      if (sc.adaptive & 2)
        continue;
      if (x < v.d.input_size)
        B += v.CIE[(size_t)s * v.bl_o + x];
      if (sc.err_sum > ERROR_GAIN_CEILING * sc.top_scaled && x < v.d.input_size) {
        const float m = sc.ih_scale * sc.ih_scale;
        A *= m;
        B *= m;
      }
    }
    ab[x] = A;
    ab[v.bl_o + x] = B;
  }
}

/* shared_o_error: in, the value before rank 0's first stream; out, the value
   this rank's first stream starts from.  after[x]: what to apply after this
   rank's streams to arrive where the last rank ends. */
__global__ void
k_bottom_compose(const float *all_ab /* [ranks][2][bl_o] */, int bl_o, int rank, int ranks,
    float *shared_o_error, float *after /* [2][bl_o] */)
{
  for (int x = threadIdx.x; x < bl_o; x += blockDim.x) {
    float running = shared_o_error[x];
    for (int q = 0; q < rank; q++)
      running = all_ab[(size_t)q * 2 * bl_o + x] * running + all_ab[(size_t)q * 2 * bl_o + bl_o + x];
    shared_o_error[x] = running;
    float A = 1.0f, B = 0.0f;
    for (int q = rank + 1; q < ranks; q++) {
      const float a = all_ab[(size_t)q * 2 * bl_o + x], b = all_ab[(size_t)q * 2 * bl_o + bl_o + x];
      A *= a;
      B = a * B + b;
    }
    after[x] = A;
    after[bl_o + x] = B;
  }
}

__global__ void
k_bottom_finish(float *shared_o_error, const float *after, int bl_o)
{
  for (int x = threadIdx.x; x < bl_o; x += blockDim.x)
    shared_o_error[x] = after[x] * shared_o_error[x] + after[bl_o + x];
}

/* delta[y, x] (+)= sum_j inputs_j[y] * accumulator_after_j[x] (recur-nn.c:755) */
__global__ void __launch_bounds__(256)
k_bottom_delta(RbView v, float *delta, int accumulate)
{
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= v.bl_i * v.bl_o)
    return;
  int y = idx / v.bl_o, x = idx - y * v.bl_o;
  float acc = accumulate ? delta[idx] : 0.0f;
  for (int j = 0; j < v.n; j++) {
    int s = slot_at(v, j);
    float a = v.BI[(size_t)s * v.bl_i + y];
    if (a != 0.0f)
      acc += a * v.BR[(size_t)s * v.bl_o + x];
  }
  delta[idx] = acc;
}

extern "C" void
rb_bottom_one_hot(const RbView *v, const u8 *hot_dev)
{
  k_bottom_one_hot<<<v->n, 64, 0, rb_stream>>>(*v, hot_dev);
  LAUNCH_CHECK("k_bottom_one_hot");
}

extern "C" void
rb_bottom_set_inputs(const RbView *v, const float *inputs_dev, int input_size)
{
  k_bottom_set_inputs<<<v->n, 64, 0, rb_stream>>>(*v, inputs_dev, input_size);
  LAUNCH_CHECK("k_bottom_set_inputs");
}

extern "C" void
rb_bottom_forward(const RbView *v, RecurNN *net, const float *shared_inputs,
    float presynaptic_noise)
{
  RecurExtraLayer *bl = net->bottom_layer;
  if (presynaptic_noise != 0.0f && net->input_size > 1) {
    k_bottom_noise<<<(v->n + 31) / 32, 32, 0, rb_stream>>>(*v, presynaptic_noise);
    LAUNCH_CHECK("k_bottom_noise");
  }
  k_bottom_forward<<<v->n, 256, v->bl_i * sizeof(float), rb_stream>>>(*v, bl->weights,
      shared_inputs, shared_inputs ? bl->outputs : NULL, presynaptic_noise != 0.0f);
  LAUNCH_CHECK("k_bottom_forward");
}

extern "C" void
rb_bottom_backward(const RbView *v, RecurNN *net, int accumulate)
{
  RecurExtraLayer *bl = net->bottom_layer;
  const int ranks = rb_comm_size(), rank = rb_comm_rank();
  static float *ab_dev = NULL;
  static size_t ab_floats = 0;
  float *after = NULL;
  if (ranks > 1) {
    /* [ranks][2][bl_o] gathered, then [2][bl_o] of "what comes after us" */
    const size_t need = ((size_t)ranks + 1) * 2 * v->bl_o;
    if (need > ab_floats) {
      cudaFree(ab_dev);
      CUDA_OR_DIE(cudaMalloc((void **)&ab_dev, need * sizeof(float)));
      ab_floats = need;
    }
    after = ab_dev + (size_t)ranks * 2 * v->bl_o;
    CUDA_OR_DIE(cudaMemsetAsync(ab_dev, 0, (size_t)ranks * 2 * v->bl_o * sizeof(float), rb_stream));
    k_bottom_affine<<<1, 128, 0, rb_stream>>>(*v, ab_dev + (size_t)rank * 2 * v->bl_o);
    LAUNCH_CHECK("k_bottom_affine");
    /* an all-gather spelt as a sum: every rank's rows are zero but its own */
    rb_comm_allreduce_sum(ab_dev, (size_t)ranks * 2 * v->bl_o);
    k_bottom_compose<<<1, 128, 0, rb_stream>>>(ab_dev, v->bl_o, rank, ranks, bl->o_error, after);
    LAUNCH_CHECK("k_bottom_compose");
  }
  k_bottom_accumulate<<<1, 128, 0, rb_stream>>>(*v, bl->o_error);
  LAUNCH_CHECK("k_bottom_accumulate");
  if (ranks > 1) {
    k_bottom_finish<<<1, 128, 0, rb_stream>>>(bl->o_error, after, v->bl_o);
    LAUNCH_CHECK("k_bottom_finish");
  }
  int total = v->bl_i * v->bl_o;
  k_bottom_delta<<<(total + 255) / 256, 256, 0, rb_stream>>>(*v, bl->delta, accumulate);
  LAUNCH_CHECK("k_bottom_delta");
}
