/* rb_kernels.cu — FP32 FMA kernels of the recur hot path for sm_100a.
 *
 * This file is the "FMA engine": every stage of the path as CUDA-core
 * kernels that are exact FP32 and work for any number of streams.  Batches
 * of >= 64 streams route the three big contractions (forward, BPTT chain,
 * weight gradient) to the tensor-core engine in rb_tc.cu instead; everything
 * else (gathers, soft clips, softmax, top layer, per-stream BPTT control,
 * optimisers, conditioning) is always done here.
 *
 * Reference stages -> kernels:
 *   a1  rnn_bptt_advance            k_advance
 *   a2  one_hot_opinion (inputs)    k_set_one_hot / k_set_inputs
 *   a3  rnn_opinion                 k_prepare_x, k_gemm<FWD>, k_out
 *   a4  calculate_interlayer        (zero rows are skipped in k_out; masked
 *                                    rows cost nothing in the GEMM tiles)
 *   a5  maybe_scale_inputs          k_prepare_x
 *   a6  softmax/net_error_bptt      k_softmax_error
 *   a7  backprop_top_layer          k_top
 *   a8  top soft clip               k_top
 *   a9  single_layer_sgd            k_ho_delta
 *   a10 bptt_and_accumulate_error   k_gemm<CHAIN> + k_chain_decide per step,
 *                                   then k_gemm<DW> (the outer products of all
 *                                   steps and streams as one contraction)
 *   a11 delta fold (ih_scale)       inside k_gemm<DW> (row scale)
 *   a13 rnn_apply_learning          k_apply_learning
 *   a14 rnn_bptt_calculate          k_sgd_top_apply + the above
 *   a15 rnn_condition_net           k_scale / k_zero_small / k_clamp / k_tall_poppy
 *   a16 rnn_bptt_clear_deltas       k_fill
 */
#include "rb_kernels.h"
#include "rb_split.cuh"
#include "rb_devmath.cuh"
#include "rb_rng.h"
#include "rb_optim.cuh"
#include <math.h>
#include <string.h>

#define RB_WARP 32

static inline int
cdiv(int a, int b)
{
  return (a + b - 1) / b;
}

/* ------------------------------------------------------------------------ */
/* small device helpers                                                       */

__device__ __forceinline__ float
warp_sum(float v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float
warp_max(float v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float
warp_min(float v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

/* block-wide sum for blockDim.x <= 1024; every thread gets the result */
__device__ __forceinline__ float
block_sum(float v, float *scratch /* >= 33 floats */)
{
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0)
    scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    float t = (lane < nw) ? scratch[lane] : 0.0f;
    t = warp_sum(t);
    if (lane == 0)
      scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

__device__ __forceinline__ float *
x_row(const RbView &v, int s, int back)
{
  int p = v.pos[s] - back;
  if (p < 0)
    p += v.depth;
  return v.X + ((size_t)p * v.cap + s) * v.d.i_size;
}

__device__ __forceinline__ float *
e_row(const RbView &v, int s, int k)
{
  return v.E + ((size_t)k * v.cap + s) * v.d.i_size;
}

/* ------------------------------------------------------------------------ */
/* a1, a2: ring position and inputs                                           */

__global__ void
k_advance(RbView v)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < v.n) {
    int s = v.slots[j];
    int p = v.pos[s] + 1;
    v.pos[s] = (p >= v.depth) ? p - v.depth : p;
  }
}

__global__ void
k_fill_iota(int *iota, int n)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n)
    iota[j] = j;
}

__global__ void
k_set_one_hot(RbView v, const u8 *hot)
{
  int s = v.slots[blockIdx.x];
  float *in = x_row(v, s, 0) + v.d.hidden_size + 1;
  int h = hot[blockIdx.x];
  for (int i = threadIdx.x; i < v.d.input_size; i += blockDim.x)
    in[i] = (i == h) ? 1.0f : 0.0f;
}

__global__ void
k_set_inputs(RbView v, const float *inputs)
{
  int s = v.slots[blockIdx.x];
  float *in = x_row(v, s, 0) + v.d.hidden_size + 1;
  const float *src = inputs + (size_t)blockIdx.x * v.d.input_size;
  for (int i = threadIdx.x; i < v.d.input_size; i += blockDim.x)
    in[i] = src[i];
}

/* stream j reads text position i + j*spacing, wrapped as charmodel-predict.c:295-298 */
__global__ void
k_text_symbols(const u8 *text, int len, int i, int spacing, int n, u8 *cur, u8 *next)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    long long off = (long long)i + (long long)j * spacing;
    if (off >= len - 1)
      off -= len - 1;
    cur[j] = text[off];
    next[j] = text[off + 1];
  }
}

/* ------------------------------------------------------------------------ */
/* a3 (first half), a5: build the input row [1 | hidden(t-1) | inputs | 0..]
   in the ring and soft-clip it in emergencies (recur-nn.c:68-81,108-115).   */

__global__ void __launch_bounds__(256)
k_prepare_x(RbView v)
{
  __shared__ float scratch[33];
  int s = v.slots[blockIdx.x];
  float *x = x_row(v, s, 0);
  const float *h = v.Hd + (size_t)s * v.d.h_size;
  int n_copy = v.d.hidden_size + 1;
  float sum = 0.0f;
  for (int i = threadIdx.x; i < v.d.i_size; i += blockDim.x) {
    float val;
    if (i < n_copy) {
      val = (i == 0) ? 1.0f : h[i];
      x[i] = val;
    }
    else {
      val = x[i];
    }
    sum += val;
  }
  sum = block_sum(sum, scratch);
  float softclip = v.d.i_size * INPUT_MEAN_SOFT_TOP;
  if (sum > softclip) {
    float scale = soft_clip_dev(sum, softclip);
    for (int i = threadIdx.x; i < v.d.i_size; i += blockDim.x)
      x[i] *= scale;
  }
}

/* ------------------------------------------------------------------------ */
/* The three big contractions as one tiled FP32 kernel.
 *
 *  FWD    hidden[b, h]  = act( sum_y x[b, y]   * Wih[y, h] )        K = i_size
 *  CHAIN  E(k+1)[b, y]  = mask * sum_x E(k)[b, x] * Wih[y, x]       K = h_size
 *  DW     delta[y, x]  += sum_r scale_r * x_r[y] * E_r[x]           K = n * depth
 *  HO     ho_delta[y,o] += sum_b hidden[b, y] * o_error[b, o]       K = n
 *  TOP    E(0)[b, y]    = mask * sum_o o_error[b, o] * Who[y, o]    K = o_size
 *
 * 64x64 output tile, 16-deep K slices, 256 threads, 4x4 per thread, global
 * loads of the next slice overlapped with the FMAs of the current one. */

#define TM 64
#define TN 64
#define TK 16
#define TPAD 4

enum { G_FWD = 0, G_CHAIN = 1, G_DW = 2, G_HO = 3, G_TOP = 4 };

struct GemmArgs {
  RbView v;
  int k;          /* CHAIN: BPTT step */
  float *delta;   /* DW: output [i_size][h_size] */
  int accumulate; /* DW: add to delta instead of overwriting */
  int use_noise;  /* FWD: add v.noise before the activation */
};

template <int MODE>
__global__ void __launch_bounds__(256)
k_gemm(GemmArgs g)
{
  const RbView &v = g.v;
  __shared__ __align__(16) float As[TK][TM + TPAD];
  __shared__ __align__(16) float Bs[TK][TN + TPAD];
  __shared__ int s_any_live;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int I = v.d.i_size, H = v.d.h_size;

  int M, N, K;
  if (MODE == G_FWD) { M = v.n; N = H; K = I; }
  else if (MODE == G_CHAIN) { M = v.n; N = I; K = H; }
  else if (MODE == G_DW) { M = I; N = H; K = v.n * v.depth; }
  else if (MODE == G_TOP) { M = v.n; N = H; K = v.d.o_size; }
  else { M = H; N = v.d.o_size; K = v.n; }

  /* --- per-thread load coordinates --- */
  /* "row" loaders: 64 rows x 16 k, thread -> (row = tid/4, kq = (tid%4)*4)
     "flat" loaders: 16 k x 64 cols, thread -> (k = tid/16, cq = (tid%16)*4) */
  const int lr = tid >> 2, lkq = (tid & 3) * 4;
  const int fk = tid >> 4, fcq = (tid & 15) * 4;

  const float *a_ptr = NULL; /* row loaders of A (FWD, CHAIN) */
  const float *b_ptr = NULL; /* row loader of B (CHAIN) / flat loader base (FWD) */
  bool a_ok = false, b_ok = false;

  if (MODE == G_CHAIN) {
    if (tid == 0)
      s_any_live = 0;
    __syncthreads();
  }
  if (MODE == G_TOP) {
    int m = m0 + lr;
    if (m < M) {
      a_ptr = v.OE + (size_t)v.slots[m] * K;
      a_ok = true;
    }
    int n = n0 + lr;
    b_ok = n < N;
    b_ptr = v.Who + (size_t)n * K;
  }
  if (MODE == G_FWD || MODE == G_CHAIN) {
    int m = m0 + lr;
    if (m < M) {
      int s = v.slots[m];
      if (MODE == G_FWD) {
        a_ptr = x_row(v, s, 0);
        a_ok = true;
      }
      else {
        a_ok = v.sc[s].live != 0;
        a_ptr = e_row(v, s, g.k);
        if (a_ok)
          s_any_live = 1;
      }
    }
  }
  if (MODE == G_CHAIN) {
    int n = n0 + lr;
    b_ok = n < N;
    b_ptr = v.Wih + (size_t)n * H;
    __syncthreads();
    if (!s_any_live)
      return; /* every stream of this tile has stopped */
  }

  float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;

  auto load_tile = [&](int k0) {
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = ra;
    if (MODE == G_FWD) {
      int k = k0 + lkq;
      if (a_ok && k < K)
        ra = *(const float4 *)(a_ptr + k);
      int kb = k0 + fk, c = n0 + fcq;
      if (kb < K && c < N)
        rb = *(const float4 *)(v.Wih + (size_t)kb * H + c);
    }
    else if (MODE == G_CHAIN || MODE == G_TOP) {
      int k = k0 + lkq;
      if (k < K) {
        if (a_ok)
          ra = *(const float4 *)(a_ptr + k);
        if (b_ok)
          rb = *(const float4 *)(b_ptr + k);
      }
    }
    else if (MODE == G_HO) {
      int r = k0 + fk;
      if (r < K) {
        int s = v.slots[r];
        int ca = m0 + fcq, cb = n0 + fcq;
        if (ca < M)
          ra = *(const float4 *)(v.Hd + (size_t)s * H + ca);
        if (cb < N)
          rb = *(const float4 *)(v.OE + (size_t)s * N + cb);
      }
    }
    else {
      int r = k0 + fk;
      if (r < K) {
        int step = r / v.n, b = r - step * v.n;
        int s = v.slots[b];
        if (step < v.sc[s].n_steps) {
          float scale = v.sc[s].ih_scale;
          int ca = m0 + fcq, cb = n0 + fcq;
          if (ca < M) {
            ra = *(const float4 *)(x_row(v, s, step) + ca);
            if (v.activation == RNN_RECLIP20) {
              /* recur-nn.c:347: a row whose input sits at the clip (>= 20)
                 is skipped altogether, its delta row included */
              if (!(ra.x < 20.0f)) ra.x = 0.0f;
              if (!(ra.y < 20.0f)) ra.y = 0.0f;
              if (!(ra.z < 20.0f)) ra.z = 0.0f;
              if (!(ra.w < 20.0f)) ra.w = 0.0f;
            }
            ra.x *= scale; ra.y *= scale; ra.z *= scale; ra.w *= scale;
          }
          if (cb < N)
            rb = *(const float4 *)(e_row(v, s, step) + cb);
        }
      }
    }
  };

  auto store_tile = [&]() {
    if (MODE == G_FWD) {
      As[lkq + 0][lr] = ra.x; As[lkq + 1][lr] = ra.y;
      As[lkq + 2][lr] = ra.z; As[lkq + 3][lr] = ra.w;
      *(float4 *)&Bs[fk][fcq] = rb;
    }
    else if (MODE == G_CHAIN || MODE == G_TOP) {
      As[lkq + 0][lr] = ra.x; As[lkq + 1][lr] = ra.y;
      As[lkq + 2][lr] = ra.z; As[lkq + 3][lr] = ra.w;
      Bs[lkq + 0][lr] = rb.x; Bs[lkq + 1][lr] = rb.y;
      Bs[lkq + 2][lr] = rb.z; Bs[lkq + 3][lr] = rb.w;
    }
    else {
      *(float4 *)&As[fk][fcq] = ra;
      *(float4 *)&Bs[fk][fcq] = rb;
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      acc[i][j] = 0.0f;

  const int n_tiles = (K + TK - 1) / TK;
  load_tile(0);
  store_tile();
  __syncthreads();
  for (int t = 0; t < n_tiles; t++) {
    if (t + 1 < n_tiles)
      load_tile((t + 1) * TK);
#pragma unroll
    for (int kk = 0; kk < TK; kk++) {
      float4 a = *(const float4 *)&As[kk][ty * 4];
      float4 b = *(const float4 *)&Bs[kk][tx * 4];
      acc[0][0] += a.x * b.x; acc[0][1] += a.x * b.y; acc[0][2] += a.x * b.z; acc[0][3] += a.x * b.w;
      acc[1][0] += a.y * b.x; acc[1][1] += a.y * b.y; acc[1][2] += a.y * b.z; acc[1][3] += a.y * b.w;
      acc[2][0] += a.z * b.x; acc[2][1] += a.z * b.y; acc[2][2] += a.z * b.z; acc[2][3] += a.z * b.w;
      acc[3][0] += a.w * b.x; acc[3][1] += a.w * b.y; acc[3][2] += a.w * b.z; acc[3][3] += a.w * b.w;
    }
    __syncthreads();
    if (t + 1 < n_tiles) {
      store_tile();
      __syncthreads();
    }
  }

  /* --- epilogues --- */
  if (MODE == G_FWD) {
    /* activation, recur-nn.c:120-148 */
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int m = m0 + ty * 4 + i;
      if (m >= M)
        continue;
      int s = v.slots[m];
      int c = n0 + tx * 4;
      if (c >= N)
        continue;
      float out[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float h = acc[i][j];
        int col = c + j;
        if (g.use_noise && col >= 1)
          h += v.noise[(size_t)s * H + col];
        if (v.activation == RNN_RESQRT) {
          h = (h > 0.0f) ? sqrtf(h + 1.0f) - 1.0f : 0.0f;
        }
        else if (v.activation == RNN_RECLIP20) {
          if (col >= 1) {
            h = h - RNN_HIDDEN_PENALTY;
            h = h < 20.0f ? h : 20.0f;
            h = (h > 0.0f) ? h : 0.0f;
          }
        }
        else {
          if (col >= 1) {
            h = h - RNN_HIDDEN_PENALTY;
            h = (h > 0.0f) ? h : 0.0f;
          }
        }
        if (col == 0)
          h = 1.0f;
        out[j] = h;
      }
      *(float4 *)(v.Hd + (size_t)s * H + c) = make_float4(out[0], out[1], out[2], out[3]);
    }
  }
  else if (MODE == G_TOP) {
    /* recur-nn.c:199-228: error reaches only hidden units that fired (and not
       the bias node); sum |e| per stream in column-block partials */
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int m = m0 + ty * 4 + i;
      float ab = 0.0f;
      int s = 0;
      int c = n0 + tx * 4;
      if (m < M) {
        s = v.slots[m];
        if (c < N) {
          const float4 hv = *(const float4 *)(v.Hd + (size_t)s * H + c);
          float h[4] = {hv.x, hv.y, hv.z, hv.w};
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            float e = (h[j] != 0.0f && c + j >= 1) ? acc[i][j] : 0.0f;
            ab += fabsf(e);
            /* pad units can fire when presynaptic noise is on; they count
               in the error sum (recur-nn.c:215-226) but the walk clears
               their error before using it (recur-nn.c:335-337) */
            if (c + j > v.d.hidden_size)
              e = 0.0f;
            o[j] = e;
          }
          *(float4 *)(e_row(v, s, 0) + c) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      ab += __shfl_xor_sync(0xffffffffu, ab, 8);
      ab += __shfl_xor_sync(0xffffffffu, ab, 4);
      ab += __shfl_xor_sync(0xffffffffu, ab, 2);
      ab += __shfl_xor_sync(0xffffffffu, ab, 1);
      if (tx == 0 && m < M)
        v.partial[(size_t)s * v.n_part + blockIdx.x] = ab;
    }
  }
  else if (MODE == G_CHAIN) {
    /* recur-nn.c:338-376 for one ring row: mask by the input that fed the
       row, ReSQRT derivative, squared-error partial sums per stream */
    const int hs1 = v.d.hidden_size + 1;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int m = m0 + ty * 4 + i;
      bool row_ok = false;
      int s = 0;
      if (m < M) {
        s = v.slots[m];
        row_ok = v.sc[s].live != 0;
      }
      float sq = 0.0f;
      int c = n0 + tx * 4;
      if (row_ok && c < N) {
        const float4 xin = *(const float4 *)(x_row(v, s, g.k) + c);
        float xi[4] = {xin.x, xin.y, xin.z, xin.w};
        float out[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          float e = 0.0f;
          float input = xi[j];
          if (input != 0.0f && (v.activation != RNN_RECLIP20 || input < 20.0f)) {
            e = acc[i][j];
            if (v.activation == RNN_RESQRT)
              e /= 2.0f * (input + 1.0f);
            sq += e * e;
          }
          int col = c + j;
          /* input-row errors feed the bottom layer (recur-nn.c:377-382) */
          if (v.CIE && col >= hs1 && col < hs1 + v.d.input_size)
            v.CIE[(size_t)s * v.bl_o + col - hs1] += e;
          /* the next step reads this row as h_error: bias and pad columns
             are cleared there (recur-nn.c:334-337) */
          if (col == 0 || (col >= hs1 && col < H))
            e = 0.0f;
          out[j] = e;
        }
        *(float4 *)(e_row(v, s, g.k + 1) + c) = make_float4(out[0], out[1], out[2], out[3]);
      }
      /* the 16 threads with equal ty hold one row of the tile */
      sq += __shfl_xor_sync(0xffffffffu, sq, 8);
      sq += __shfl_xor_sync(0xffffffffu, sq, 4);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      if (tx == 0 && row_ok)
        v.partial[(size_t)s * v.n_part + blockIdx.x] = sq;
    }
  }
  else {
    const int ldd = (MODE == G_HO) ? N : H;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int m = m0 + ty * 4 + i;
      int c = n0 + tx * 4;
      if (m >= M || c >= N)
        continue;
      float *dst = g.delta + (size_t)m * ldd + c;
      float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (g.accumulate) {
        float4 d = *(const float4 *)dst;
        o.x += d.x; o.y += d.y; o.z += d.z; o.w += d.w;
      }
      *(float4 *)dst = o;
    }
  }
}

/* ------------------------------------------------------------------------ */
/* a3 (second half): output = hidden . Who; hidden rows that are zero are
   skipped, which is the reference's row-skipping (recur-nn.c:39-48) — the
   branch is uniform across the block because the block is one stream.     */

__global__ void __launch_bounds__(256)
k_out(RbView v)
{
  extern __shared__ float sh[]; /* h_size hidden + reduction space */
  int s = v.slots[blockIdx.x];
  const int H = v.d.h_size, O = v.d.o_size;
  float *hid = sh;
  float *red = sh + H;
  for (int i = threadIdx.x; i < H; i += blockDim.x)
    hid[i] = v.Hd[(size_t)s * H + i];
  __syncthreads();
  const int CW = (O >= 256) ? 256 : O;
  const int G = 256 / CW;
  const int col = threadIdx.x % CW, grp = threadIdx.x / CW;
  float *y = v.Y + (size_t)s * O;
  /* wide output layers (config 4: 3650 outputs) spread their 256-column
     chunks over blockIdx.y */
  for (int c0 = blockIdx.y * CW; c0 < O; c0 += gridDim.y * CW) {
    int c = c0 + col;
    float acc = 0.0f;
    if (grp < G && c < O) {
      /* eight rows' loads in flight together; silent rows load nothing */
      for (int rb = grp; rb < H; rb += 8 * G) {
        float w[8], hv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          int r = rb + u * G;
          hv[u] = (r < H) ? hid[r] : 0.0f;
          w[u] = (hv[u] != 0.0f) ? __ldg(v.Who + (size_t)r * O + c) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
          acc = fmaf(hv[u], w[u], acc);
      }
    }
    if (G > 1) {
      if (grp < G)
        red[grp * CW + col] = acc;
      __syncthreads();
      if (grp == 0 && c < O) {
        float t = 0.0f;
        for (int q = 0; q < G; q++)
          t += red[q * CW + col];
        y[c] = t;
      }
      __syncthreads();
    }
    else if (c < O) {
      y[c] = acc;
    }
  }
}

/* ------------------------------------------------------------------------ */
/* a6: softmax with badmaths.h's clamp and fast_expf, error = onehot - p,
   argmax; one warp per stream, warp-shuffle reductions
   (badmaths.h:71-141, charmodel-predict.c:18-27).                          */

__device__ __forceinline__ void
softmax_error_warp(const RbView &v, int j, int lane, const u8 *target, float *err_out,
    int *winner_out, const float *y_row = NULL, float *work_row = NULL)
{
  /* y_row / work_row: the stream's outputs and a scratch row in shared memory,
     when the caller has them there (the error row is then left in work_row as
     well as in the pool) */
  int s = v.slots[j];
  const int len = v.d.output_size, O = v.d.o_size;
  const float *src = y_row ? y_row : v.Y + (size_t)s * O;
  float *out = v.OE + (size_t)s * O;
  float *err = work_row ? work_row : out;

  float mx = -INFINITY, mn = INFINITY;
  for (int i = lane; i < len; i += 32) {
    float y = src[i];
    mx = fmaxf(mx, y);
    mn = fminf(mn, y);
  }
  mx = warp_max(mx);
  mn = warp_min(mn);
  const float max_exp = 50.0f, min_exp = -60.0f;
  float adj = 0.0f;
  if (mx > max_exp)
    adj = max_exp - mx;
  else if (mn < min_exp)
    adj = fminf(min_exp - mn, max_exp - mx);

  float sum = 0.0f;
  for (int i = lane; i < len; i += 32) {
    float x = fast_expf_dev(src[i] + adj);
    err[i] = x;
    sum += x;
  }
  sum = warp_sum(sum);
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  int tgt = target ? (int)target[j] : -1;
  float e_t = 0.0f;
  for (int i = lane; i < len; i += 32) {
    float p = err[i] / sum;
    if (p > best) { /* first maximum wins within a lane: indices ascend */
      best = p;
      best_i = i;
    }
    float e = -p;
    if (i == tgt) {
      e += 1.0f;
      e_t = e;
    }
    err[i] = e;
    if (work_row)
      out[i] = e;
  }
  for (int i = len + lane; i < O; i += 32) {
    err[i] = 0.0f;
    if (work_row)
      out[i] = 0.0f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) {
      best = ob;
      best_i = oi;
    }
  }
  e_t = warp_sum(e_t);
  if (lane == 0) {
    if (err_out)
      err_out[j] = e_t;
    if (winner_out)
      winner_out[j] = best_i;
  }
}

__global__ void __launch_bounds__(128)
k_softmax_error(RbView v, const u8 *target, float *err_out, int *winner_out)
{
  int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j < v.n)
    softmax_error_warp(v, j, threadIdx.x & 31, target, err_out, winner_out);
}

/* sums of charmodel-predict.c:301-303 over the batch, in a fixed order */
__device__ __forceinline__ void
char_accum_block(const float *err, const int *winner, const u8 *target, int n,
    RbCharAccum *acc, RbCharAccum *snapshot = NULL, int reset = 0)
{
  /* the first 256 threads of the block do the sums (a fixed order whatever
     the block size); the others only keep the barriers company */
  __shared__ double s_err[256], s_ent[256];
  __shared__ int s_cor[256];
  double e = 0.0, h = 0.0;
  int c = 0;
  for (int j = threadIdx.x; j < n && threadIdx.x < 256; j += 256) {
    float ej = __ldcg(err + j); /* written by other blocks of the same kernel in the fused case */
    e += ej;
    float x = 1.0f - ej;
    h += (x < 1e-30f) ? -100.0f : log2f(x);
    c += (__ldcg(winner + j) == (int)target[j]);
  }
  if (threadIdx.x < 256) {
    s_err[threadIdx.x] = e;
    s_ent[threadIdx.x] = h;
    s_cor[threadIdx.x] = c;
  }
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_err[threadIdx.x] += s_err[threadIdx.x + o];
      s_ent[threadIdx.x] += s_ent[threadIdx.x + o];
      s_cor[threadIdx.x] += s_cor[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    RbCharAccum t = *acc;
    if (reset) /* the host has taken the sums so far */
      t.error = t.entropy = 0.0, t.correct = t.count = 0;
    t.error += s_err[0];
    t.entropy += s_ent[0];
    t.correct += s_cor[0];
    t.count += n;
    *acc = t;
    if (snapshot) /* pinned host memory: the caller reads it after its next synchronise */
      *snapshot = t;
  }
}

__global__ void __launch_bounds__(256)
k_char_accum(const float *err, const int *winner, const u8 *target, int n,
    RbCharAccum *acc)
{
  char_accum_block(err, winner, target, n, acc);
}

/* ------------------------------------------------------------------------ */
/* a7, a8: top layer back-propagation into E[0] with the soft clip, and the
   per-stream set-up of the BPTT walk (recur-nn.c:199-228, 318-322, 719-721) */

__global__ void __launch_bounds__(256)
k_top(RbView v, const RecurErrorRange *ranges, int n_ranges)
{
  extern __shared__ float sh[]; /* o_size errors + 33 scratch */
  int s = v.slots[blockIdx.x];
  const int H = v.d.h_size, O = v.d.o_size, I = v.d.i_size;
  float *oe = sh;
  float *scratch = sh + O;
  for (int i = threadIdx.x; i < O; i += blockDim.x)
    oe[i] = v.OE[(size_t)s * O + i];
  __syncthreads();
  const float *hid = v.Hd + (size_t)s * H;
  float *e0 = e_row(v, s, 0);
  /* With error ranges the reference leaves h_error[y] of silent hidden units
     untouched (no else branch, recur-nn.c:175-193), i.e. whatever the previous
     BPTT walk of this stream left in that buffer.  The two error buffers swap
     each step (recur-nn.c:384-386), so after n steps the one called h_error
     holds E(n) for even n and E(n-1) for odd n. */
  const int n_prev = v.sc[s].n_steps;
  const float *stale = e_row(v, s, (n_prev & 1) ? n_prev - 1 : n_prev);
  float abs_sum = 0.0f, hsum = 0.0f, hmag = 0.0f;
  int hzero = 0;
  for (int y = threadIdx.x; y < I; y += blockDim.x) {
    float e = 0.0f;
    if (y < H) {
      float h = hid[y];
      hsum += h;
      hmag += h * h;
      hzero += (h == 0.0f);
      if (n_ranges != 0 && !(y >= 1 && h != 0.0f))
        e = stale[y];
      if (y >= 1 && h != 0.0f) {
        const float *row = v.Who + (size_t)y * O;
        if (n_ranges == 0) {
          for (int x = 0; x < O; x += 4) {
            float4 w = *(const float4 *)(row + x);
            e += w.x * oe[x] + w.y * oe[x + 1] + w.z * oe[x + 2] + w.w * oe[x + 3];
          }
          abs_sum += fabsf(e);
        }
        else {
          /* recur-nn.c:178-191: the running dot product is added to the
             error sum once per range */
          for (int q = 0; q < n_ranges; q++) {
            int start = ranges[q].start & ~3;
            int len = (ranges[q].len + 3) & ~3;
            for (int x = start; x < start + len; x++)
              e += row[x] * oe[x];
            abs_sum += fabsf(e);
          }
        }
      }
    }
    if (y > v.d.hidden_size)
      e = 0.0f; /* pad units: counted above, cleared before the walk (recur-nn.c:335-337) */
    e0[y] = e;
  }
  abs_sum = block_sum(abs_sum, scratch);
  hsum = block_sum(hsum, scratch);
  hmag = block_sum(hmag, scratch);
  float hz = block_sum((float)hzero, scratch);
  float halfmax = H * MAX_TOP_ERROR_FACTOR;
  float top_scaled = abs_sum;
  if (abs_sum > halfmax) {
    float scale = soft_clip_dev(abs_sum, halfmax);
    for (int y = threadIdx.x; y < H; y += blockDim.x)
      e0[y] *= scale;
    top_scaled = scale * abs_sum;
  }
  if (v.CIE)
    for (int i = threadIdx.x; i < v.bl_o; i += blockDim.x)
      v.CIE[(size_t)s * v.bl_o + i] = 0.0f;
  if (threadIdx.x == 0) {
    RbScalars *sc = v.sc + s;
    sc->top_raw = abs_sum;
    sc->top_scaled = top_scaled;
    sc->hidden_sum = hsum;
    sc->hidden_mag = sqrtf(hmag);
    sc->hidden_zeros = (int)(hz + 0.5f);
    float min_gain = MIN_ERROR_GAIN * top_scaled;
    sc->min_sum = fminf(sc->mef / sc->lr, min_gain);
    sc->max_sum = MAX_ERROR_GAIN * top_scaled + 1.0f;
    sc->cum_error = 0.0f;
    sc->err_sum = 0.0f;
    sc->live = (v.depth > 0) && !(sc->adaptive & 2);
    sc->n_steps = 0;
    sc->t_left = v.depth;
    sc->ih_scale = 1.0f;
  }
}

/* a8 and the set-up of the BPTT walk after k_gemm<TOP>: per stream, total
   |e| from the column-block partials, the hidden statistics the log wants,
   the soft clip (recur-nn.c:720-721) and the walk's thresholds
   (recur-nn.c:318-322).                                                     */
__global__ void __launch_bounds__(256)
k_top_finish(RbView v, int n_col_blocks)
{
  __shared__ float scratch[33];
  const int s = v.slots[blockIdx.x];
  const int H = v.d.h_size;
  const float *hid = v.Hd + (size_t)s * H;
  float hsum = 0.0f, hmag = 0.0f, hzero = 0.0f;
  for (int y = threadIdx.x; y < H; y += blockDim.x) {
    float h = hid[y];
    hsum += h;
    hmag += h * h;
    hzero += (h == 0.0f);
  }
  float total = 0.0f;
  if (threadIdx.x == 0)
    for (int q = 0; q < n_col_blocks; q++)
      total += v.partial[(size_t)s * v.n_part + q];
  total = block_sum(total, scratch);
  hsum = block_sum(hsum, scratch);
  hmag = block_sum(hmag, scratch);
  hzero = block_sum(hzero, scratch);
  const float halfmax = H * MAX_TOP_ERROR_FACTOR;
  const float scale = (total > halfmax) ? soft_clip_dev(total, halfmax) : 1.0f;
  float *e0 = e_row(v, s, 0);
  if (scale != 1.0f) {
    for (int y = threadIdx.x; y < H; y += blockDim.x)
      e0[y] *= scale;
  }
  if (v.CIE)
    for (int i = threadIdx.x; i < v.bl_o; i += blockDim.x)
      v.CIE[(size_t)s * v.bl_o + i] = 0.0f;
  if (threadIdx.x == 0) {
    RbScalars *sc = v.sc + s;
    float top_scaled = (total > halfmax) ? scale * total : total;
    sc->top_raw = total;
    sc->top_scaled = top_scaled;
    sc->hidden_sum = hsum;
    sc->hidden_mag = sqrtf(hmag);
    sc->hidden_zeros = (int)(hzero + 0.5f);
    float min_gain = MIN_ERROR_GAIN * top_scaled;
    sc->min_sum = fminf(sc->mef / sc->lr, min_gain);
    sc->max_sum = MAX_ERROR_GAIN * top_scaled + 1.0f;
    sc->cum_error = 0.0f;
    sc->err_sum = 0.0f;
    sc->live = (v.depth > 0) && !(sc->adaptive & 2);
    sc->n_steps = 0;
    sc->t_left = v.depth;
    sc->ih_scale = 1.0f;
  }
}

/* a9: ho_delta[y, :] (+)= sum over streams hidden[y] * o_error[:]
   (recur-nn.c:256-301).  One thread per (y, x) element, streams in order. */
__global__ void __launch_bounds__(256)
k_ho_delta(RbView v, float *ho_delta, int accumulate,
    const RecurErrorRange *ranges, int n_ranges)
{
  const int H = v.d.h_size, O = v.d.o_size;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * O)
    return;
  int y = idx / O, x = idx - y * O;
  if (n_ranges) {
    bool inside = false;
    for (int q = 0; q < n_ranges; q++) {
      int start = ranges[q].start & ~3;
      int len = (ranges[q].len + 3) & ~3;
      inside |= (x >= start && x < start + len);
    }
    if (!inside) {
      if (!accumulate)
        ho_delta[idx] = 0.0f;
      return;
    }
  }
  float acc = accumulate ? ho_delta[idx] : 0.0f;
  for (int j = 0; j < v.n; j++) {
    int s = v.slots[j];
    float h = v.Hd[(size_t)s * H + y];
    if (h != 0.0f)
      acc += h * v.OE[(size_t)s * O + x];
  }
  ho_delta[idx] = acc;
}

/* ------------------------------------------------------------------------ */
/* Batch versions of the two top-layer kernels.  A block serves OS streams and
   first stages the whole of Who in shared memory (one round of coalesced
   loads), so Who is read once per OS streams instead of once per stream and
   the inner loops run out of shared memory.  Used when Who fits (h_size x
   o_size floats + the streams' vectors <= 200 KB); otherwise the per-stream
   kernels above do the job.                                                 */

#define OS 4 /* streams per block */

__device__ __forceinline__ int
slot_of(const RbView &v, int j)
{
  return v.contiguous ? v.base + j : v.slots[j];
}

/* A matrix into shared memory as bulk asynchronous copies (TMA, no tensor
   map): one thread issues them, the copy engine keeps the SM's L2 port full
   without a register or a warp being involved, everybody waits on the
   mbarrier when the data is needed.  n_floats * 4 must be a multiple of 16. */
__device__ __forceinline__ void
bulk_stage_begin(unsigned long long *bar, float *dst, const float *src, int n_floats)
{
  const unsigned int b = (unsigned int)__cvta_generic_to_shared(bar);
  unsigned int d = (unsigned int)__cvta_generic_to_shared(dst);
  const char *g = (const char *)src;
  unsigned int left = (unsigned int)n_floats * 4u;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(left)
      : "memory");
  while (left) {
    const unsigned int chunk = left < 32768u ? left : 32768u;
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(d), "l"(g), "r"(chunk), "r"(b) : "memory");
    d += chunk;
    g += chunk;
    left -= chunk;
  }
}

__device__ __forceinline__ void
bulk_stage_init(unsigned long long *bar)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(
          (unsigned int)__cvta_generic_to_shared(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void
bulk_stage_wait(unsigned long long *bar)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"((unsigned int)__cvta_generic_to_shared(bar)) : "memory");
}

__device__ __forceinline__ void
stage_matrix(float *dst, const float *__restrict__ src, int n_floats)
{
  const float4 *s4 = (const float4 *)src;
  float4 *d4 = (float4 *)dst;
  const int n4 = n_floats >> 2;
  int i = threadIdx.x;
  /* eight independent 16-byte loads in flight per thread */
  for (; i + 7 * (int)blockDim.x < n4; i += 8 * blockDim.x) {
    float4 t[8];
#pragma unroll
    for (int u = 0; u < 8; u++)
      t[u] = __ldg(s4 + i + u * blockDim.x);
#pragma unroll
    for (int u = 0; u < 8; u++)
      d4[i + u * blockDim.x] = t[u];
  }
  if (i < n4) { /* the last, partial batch: still all loads before all stores */
    float4 t[8];
#pragma unroll
    for (int u = 0; u < 8; u++)
      if (i + u * (int)blockDim.x < n4)
        t[u] = __ldg(s4 + i + u * blockDim.x);
#pragma unroll
    for (int u = 0; u < 8; u++)
      if (i + u * (int)blockDim.x < n4)
        d4[i + u * blockDim.x] = t[u];
  }
}

/* a5's activation (recur-nn.c:229-262) for four hidden sums starting at column col0 */
__device__ __forceinline__ float4
hidden_activation(const RbView &v, float4 sum, int col0, const float *noise_row)
{
  float h4[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
  for (int u = 0; u < 4; u++) {
    int col = col0 + u;
    float h = h4[u];
    if (noise_row && col >= 1)
      h += noise_row[col];
    if (v.activation == RNN_RESQRT) {
      h = (h > 0.0f) ? sqrtf(h + 1.0f) - 1.0f : 0.0f;
    }
    else if (v.activation == RNN_RECLIP20) {
      if (col >= 1) {
        h = h < 20.0f ? h : 20.0f;
        h = (h > 0.0f) ? h : 0.0f;
      }
    }
    else if (col >= 1) {
      h = (h > 0.0f) ? h : 0.0f;
    }
    if (col == 0)
      h = 1.0f;
    h4[u] = h;
  }
  return make_float4(h4[0], h4[1], h4[2], h4[3]);
}

/* FROM_PARTIALS: the hidden rows arrive as split-K partial sums of the tensor
   engine's forward GEMM; this kernel sums them (fixed order), applies the
   activation, writes the hidden rows and goes on to the output layer. */
/* optional tail of k_out_multi: the char model's softmax error against the
   next symbol for the block's streams, and - by whichever block finishes
   last - the sums over the batch, in the fixed order k_char_accum uses */
struct RbLossArgs {
  const u8 *target; /* NULL: no loss */
  float *err;
  int *winner;
  RbCharAccum *accum;
  RbCharAccum *snapshot; /* pinned host copy of the sums, or NULL */
  int reset;             /* start the sums from zero */
  int fuse_top;          /* go on to the top layer's back-propagation (a7, a8) */
  unsigned long long *dbg; /* RECUR_B200_OUT_TIMING: globaltimer stamps of block 0 */
};

#define OUT_STAMP(slot) do {                                            \
    if (loss.dbg && blockIdx.x == 0 && threadIdx.x == 0) {              \
      unsigned long long t_;                                            \
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_));             \
      loss.dbg[slot] = t_;                                              \
    }                                                                   \
  } while (0)

__device__ unsigned int rb_loss_ticket;

#define OUT_NT 512 /* threads of k_out_multi: sixteen warps to hide its many short dependent phases */

template <bool FROM_PARTIALS>
__global__ void __launch_bounds__(OUT_NT)
k_out_multi(RbView v, RbFwdPartials fp, RbLossArgs loss)
{
  extern __shared__ __align__(16) float sh[]; /* Who | OS hidden rows | reduction space */
  const int H = v.d.h_size, O = v.d.o_size;
  const int j0 = blockIdx.x * OS;
  const int ns = min(OS, v.n - j0);
  float *who = sh;
  float *hid = sh + (size_t)H * O;
  float *red = hid + (size_t)OS * H;
  float *ysm = red + OS * OUT_NT;  /* [OS][O] the block's output rows, for the loss */
  float *soe = ysm + OS * O;    /* [OS][O] ... and their error rows */
  __shared__ __align__(8) unsigned long long who_bar;
  OUT_STAMP(0);
  if (threadIdx.x == 0)
    bulk_stage_init(&who_bar);
  __syncthreads();
  if (threadIdx.x == 0)
    bulk_stage_begin(&who_bar, who, v.Who, H * O);
  if (FROM_PARTIALS) {
    /* every load of the block's rows is issued before the weights are staged */
    constexpr int MAXZ = 4;
    for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
      float4 pz[OS][MAXZ];
#pragma unroll
      for (int q = 0; q < OS; q++) {
        if (q < ns) {
          const float *row = fp.part + (size_t)slot_of(v, j0 + q) * fp.pitch + c;
#pragma unroll
          for (int z = 0; z < MAXZ; z++)
            if (z < fp.splits)
              pz[q][z] = __ldcg((const float4 *)(row + z * fp.split_stride));
        }
      }
#pragma unroll
      for (int q = 0; q < OS; q++) {
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < ns) {
          int slot = slot_of(v, j0 + q);
          h = pz[q][0];
#pragma unroll
          for (int z = 1; z < MAXZ; z++) {
            if (z < fp.splits) {
              h.x += pz[q][z].x; h.y += pz[q][z].y; h.z += pz[q][z].z; h.w += pz[q][z].w;
            }
          }
          h = hidden_activation(v, h, c, fp.use_noise ? v.noise + (size_t)slot * H : NULL);
          *(float4 *)(v.Hd + (size_t)slot * H + c) = h;
        }
        *(float4 *)(hid + (size_t)q * H + c) = h;
      }
    }
  }
  else {
    for (int q = 0; q < OS; q++) {
      if (q < ns)
        stage_matrix(hid + (size_t)q * H, v.Hd + (size_t)slot_of(v, j0 + q) * H, H);
      else
        for (int i = threadIdx.x; i < H; i += blockDim.x)
          hid[(size_t)q * H + i] = 0.0f;
    }
  }
  OUT_STAMP(1);
  bulk_stage_wait(&who_bar);
  __syncthreads();
  OUT_STAMP(2);
  const int CW = (O >= OUT_NT) ? OUT_NT : O;
  const int G = OUT_NT / CW;
  const int col = threadIdx.x % CW, grp = threadIdx.x / CW;
  for (int c0 = 0; c0 < O; c0 += CW) {
    int c = c0 + col;
    float acc[OS];
#pragma unroll
    for (int q = 0; q < OS; q++)
      acc[q] = 0.0f;
    if (grp < G && c < O) {
      /* each column group takes a contiguous run of hidden rows, four at a
         time: one 16-byte read of each stream's hidden values per four
         weights, sixteen multiply-adds per eight shared-memory reads */
      const int rows_per = ((H + G - 1) / G + 3) & ~3;
      const int r0 = grp * rows_per, r1 = min(H, r0 + rows_per);
#pragma unroll 2
      for (int r = r0; r < r1; r += 4) {
        const float *wp = who + (size_t)r * O + c;
        float w0 = wp[0], w1 = wp[O], w2 = wp[2 * O], w3 = wp[3 * O];
#pragma unroll
        for (int q = 0; q < OS; q++) {
          float4 h = *(const float4 *)(hid + q * H + r);
          acc[q] += h.x * w0;
          acc[q] += h.y * w1;
          acc[q] += h.z * w2;
          acc[q] += h.w * w3;
        }
      }
    }
    if (G > 1) {
      if (grp < G) {
#pragma unroll
        for (int q = 0; q < OS; q++)
          red[(grp * OS + q) * CW + col] = acc[q];
      }
      __syncthreads();
      if (grp == 0 && c < O) {
        for (int q = 0; q < ns; q++) {
          float t = 0.0f;
          for (int gq = 0; gq < G; gq++)
            t += red[(gq * OS + q) * CW + col];
          v.Y[(size_t)slot_of(v, j0 + q) * O + c] = t;
          if (loss.target)
            ysm[q * O + c] = t;
        }
      }
      __syncthreads();
    }
    else if (c < O) {
      for (int q = 0; q < ns; q++) {
        v.Y[(size_t)slot_of(v, j0 + q) * O + c] = acc[q];
        if (loss.target)
          ysm[q * O + c] = acc[q];
      }
    }
  }
  OUT_STAMP(3);
  if (loss.target) {
    __shared__ int s_last;
    __syncthreads(); /* the block's rows of Y are written */
    const int w = threadIdx.x >> 5;
    if (w < ns)
      softmax_error_warp(v, j0 + w, threadIdx.x & 31, loss.target, loss.err, loss.winner,
          ysm + w * O, soe + w * O);
    else if (w < OS)
      for (int i = threadIdx.x & 31; i < O; i += 32)
        soe[w * O + i] = 0.0f;
    OUT_STAMP(4);
    if (loss.fuse_top) {
      /* a7/a8 for the block's streams while Who, their hidden rows and now
         their error rows are in shared memory (recur-nn.c:199-228, 318-322,
         719-721): what k_gemm<TOP> + k_top_finish do for the general case */
      float *part = red;  /* [warps][4 * OS], then [4 * OS] totals */
      float *tot = red + (OUT_NT / 32) * (4 * OS);
      /* what the scalars at the end need from global memory, asked for now */
      float pre_mef = 0.0f, pre_lr = 1.0f;
      int pre_adaptive = 0;
      if ((int)threadIdx.x < ns) {
        const RbScalars *sc = v.sc + slot_of(v, j0 + threadIdx.x);
        pre_mef = sc->mef;
        pre_lr = sc->lr;
        pre_adaptive = sc->adaptive;
      }
      __syncthreads();    /* the error rows are in soe */
      OUT_STAMP(7);
      constexpr int YR = 2048 / OUT_NT; /* h_size <= 2048 on this path */
      float e[YR][OS];
      float st[4 * OS]; /* per stream: |e| sum, hidden sum, squares, zeros */
#pragma unroll
      for (int u = 0; u < 4 * OS; u++)
        st[u] = 0.0f;
#pragma unroll
      for (int r = 0; r < YR; r++) {
        const int y = threadIdx.x + OUT_NT * r;
#pragma unroll
        for (int q = 0; q < OS; q++)
          e[r][q] = 0.0f;
        if (y < H) {
          float h[OS];
          bool any = false;
#pragma unroll
          for (int q = 0; q < OS; q++) {
            h[q] = hid[q * H + y];
            st[4 * q + 1] += h[q];
            st[4 * q + 2] += h[q] * h[q];
            st[4 * q + 3] += (h[q] == 0.0f);
            any = any || h[q] != 0.0f;
          }
          if (y >= 1 && y <= v.d.hidden_size && any) {
            const float *wrow = who + (size_t)y * O;
            for (int x = 0; x < O; x += 4) {
              float4 wv = *(const float4 *)(wrow + x);
#pragma unroll
              for (int q = 0; q < OS; q++) {
                const float4 oq = *(const float4 *)(soe + q * O + x);
                e[r][q] += wv.x * oq.x + wv.y * oq.y + wv.z * oq.z + wv.w * oq.w;
              }
            }
#pragma unroll
            for (int q = 0; q < OS; q++) {
              if (h[q] == 0.0f)
                e[r][q] = 0.0f;
              st[4 * q] += fabsf(e[r][q]);
            }
          }
        }
      }
      OUT_STAMP(8);
#pragma unroll
      for (int u = 0; u < 4 * OS; u++) {
        float t = st[u];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0)
          part[(threadIdx.x >> 5) * (4 * OS) + u] = t;
      }
      __syncthreads();
      OUT_STAMP(9);
      /* the warps' parts, summed in warp order by one thread per value */
      if (threadIdx.x < 4 * OS) {
        float t = 0.0f;
        for (int wq = 0; wq < OUT_NT / 32; wq++)
          t += part[wq * (4 * OS) + threadIdx.x];
        tot[threadIdx.x] = t;
      }
      __syncthreads();
      float scale[OS];
#pragma unroll
      for (int q = 0; q < OS; q++) {
        const float total = tot[4 * q];
        const float halfmax = H * MAX_TOP_ERROR_FACTOR;
        scale[q] = (total > halfmax) ? soft_clip_dev(total, halfmax) : 1.0f;
        if ((int)threadIdx.x == q && q < ns) {
          const float hsum = tot[4 * q + 1], hmag = tot[4 * q + 2], hz = tot[4 * q + 3];
          RbScalars *sc = v.sc + slot_of(v, j0 + q);
          const float top_scaled = (total > halfmax) ? scale[q] * total : total;
          sc->top_raw = total;
          sc->top_scaled = top_scaled;
          sc->hidden_sum = hsum;
          sc->hidden_mag = sqrtf(hmag);
          sc->hidden_zeros = (int)(hz + 0.5f);
          sc->min_sum = fminf(pre_mef / pre_lr, MIN_ERROR_GAIN * top_scaled);
          sc->max_sum = MAX_ERROR_GAIN * top_scaled + 1.0f;
          sc->cum_error = 0.0f;
          sc->err_sum = 0.0f;
          sc->live = (v.depth > 0) && !(pre_adaptive & 2);
          sc->n_steps = 0;
          sc->t_left = v.depth;
          sc->ih_scale = 1.0f;
        }
      }
      OUT_STAMP(10);
#pragma unroll
      for (int r = 0; r < YR; r++) {
        const int y = threadIdx.x + OUT_NT * r;
        if (y < H) {
#pragma unroll
          for (int q = 0; q < OS; q++)
            if (q < ns)
              e_row(v, slot_of(v, j0 + q), 0)[y] = e[r][q] * scale[q];
        }
      }
    }
    OUT_STAMP(5);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
      s_last = (atomicAdd(&rb_loss_ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    OUT_STAMP(6);
    if (s_last) {
      __threadfence();
      char_accum_block((const float *)loss.err, (const int *)loss.winner, loss.target, v.n,
          loss.accum, loss.snapshot, loss.reset);
      if (threadIdx.x == 0)
        rb_loss_ticket = 0u;
    }
  }
}

/* a9 for a batch that fits in shared memory: a block owns 16 hidden rows of
   ho_delta, pulls its slab of the hidden activations and all output errors
   into shared memory in one round of loads, then sums over the streams. */
#define HO_ROWS 8

__global__ void __launch_bounds__(256)
k_ho_delta_slab(RbView v, float *ho_delta, int accumulate)
{
  extern __shared__ float sh[]; /* n x HO_ROWS hidden, n x o_size errors */
  const int H = v.d.h_size, O = v.d.o_size, n = v.n;
  const int y0 = blockIdx.x * HO_ROWS;
  float *sH = sh;
  float *sO = sh + (size_t)n * HO_ROWS;
  {
    /* every load of the slab is issued before the first use */
    const int per = HO_ROWS / 4;
    for (int i0 = threadIdx.x; i0 < n * per; i0 += 8 * blockDim.x) {
      float4 t[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int i = i0 + u * blockDim.x;
        t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n * per) {
          int b = i / per, q = (i - b * per) * 4;
          if (y0 + q < H)
            t[u] = __ldg((const float4 *)(v.Hd + (size_t)slot_of(v, b) * H + y0 + q));
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int i = i0 + u * blockDim.x;
        if (i < n * per)
          *(float4 *)(sH + (size_t)i * 4) = t[u];
      }
    }
    if (v.contiguous) {
      stage_matrix(sO, v.OE + (size_t)v.base * O, n * O);
    }
    else {
      for (int i = threadIdx.x; i < n * (O / 4); i += blockDim.x) {
        int b = i / (O / 4), q = (i - b * (O / 4)) * 4;
        *(float4 *)(sO + (size_t)b * O + q) =
          *(const float4 *)(v.OE + (size_t)v.slots[b] * O + q);
      }
    }
  }
  __syncthreads();
  /* thread = (hidden row of the slab, group of 12 output columns, slice of the
     streams): 12 accumulators per thread keep the FMA pipe busier than the
     shared-memory pipe; the slices are then summed in a fixed order */
  const int n_cg = (O + 11) / 12;
  const int S = 256 / (HO_ROWS * n_cg);
  const int yl = threadIdx.x % HO_ROWS;
  const int cg = (threadIdx.x / HO_ROWS) % n_cg;
  const int sl = threadIdx.x / (HO_ROWS * n_cg);
  float *red = sO + (size_t)n * O + 16;
  float acc[12];
#pragma unroll
  for (int u = 0; u < 12; u++)
    acc[u] = 0.0f;
  if (sl < S) {
    for (int b = sl; b < n; b += S) {
      float h = sH[b * HO_ROWS + yl];
      const float4 *e4 = (const float4 *)(sO + (size_t)b * O + cg * 12);
      float4 e0 = e4[0], e1 = e4[1], e2 = e4[2];
      acc[0] += h * e0.x; acc[1] += h * e0.y; acc[2] += h * e0.z; acc[3] += h * e0.w;
      acc[4] += h * e1.x; acc[5] += h * e1.y; acc[6] += h * e1.z; acc[7] += h * e1.w;
      acc[8] += h * e2.x; acc[9] += h * e2.y; acc[10] += h * e2.z; acc[11] += h * e2.w;
    }
#pragma unroll
    for (int u = 0; u < 12; u++)
      red[((size_t)sl * HO_ROWS + yl) * (n_cg * 12) + cg * 12 + u] = acc[u];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < HO_ROWS * O; i += blockDim.x) {
    int y = i / O, o = i - y * O;
    if (y0 + y >= H)
      continue;
    float t = 0.0f;
    for (int q = 0; q < S; q++)
      t += red[((size_t)q * HO_ROWS + y) * (n_cg * 12) + o];
    size_t idx = (size_t)(y0 + y) * O + o;
    ho_delta[idx] = (accumulate ? ho_delta[idx] : 0.0f) + t;
  }
}

/* The same with the streams split over blockIdx.y as well, for batches whose
   slab would keep a block to itself on an SM: every block leaves its part in
   `parts[z]`, the last block of a row group to finish (ticket) adds the parts
   in split order - a fixed order, whoever that block is. */
__global__ void __launch_bounds__(256)
k_ho_delta_split(RbView v, float *ho_delta, int accumulate, float *parts, unsigned int *tickets)
{
  extern __shared__ float sh[]; /* nz x HO_ROWS hidden, nz x o_size errors, reduction space */
  __shared__ int s_last;
  const int H = v.d.h_size, O = v.d.o_size;
  const int y0 = blockIdx.x * HO_ROWS;
  const int Z = gridDim.y, z = blockIdx.y;
  const int per_z = (v.n + Z - 1) / Z;
  const int b0 = z * per_z, nz = max(0, min(v.n, b0 + per_z) - b0);
  float *sH = sh;
  float *sO = sh + (size_t)per_z * HO_ROWS;
  float *red = sO + (size_t)per_z * O + 16;
  {
    const int per = HO_ROWS / 4;
    for (int i0 = threadIdx.x; i0 < nz * per; i0 += 4 * blockDim.x) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        int i = i0 + u * blockDim.x;
        t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < nz * per) {
          int b = i / per, q = (i - b * per) * 4;
          if (y0 + q < H)
            t[u] = __ldg((const float4 *)(v.Hd + (size_t)slot_of(v, b0 + b) * H + y0 + q));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        int i = i0 + u * blockDim.x;
        if (i < nz * per)
          *(float4 *)(sH + (size_t)i * 4) = t[u];
      }
    }
    if (v.contiguous) {
      stage_matrix(sO, v.OE + (size_t)(v.base + b0) * O, nz * O);
    }
    else {
      for (int i = threadIdx.x; i < nz * (O / 4); i += blockDim.x) {
        int b = i / (O / 4), q = (i - b * (O / 4)) * 4;
        *(float4 *)(sO + (size_t)b * O + q) =
          *(const float4 *)(v.OE + (size_t)v.slots[b0 + b] * O + q);
      }
    }
  }
  __syncthreads();
  const int n_cg = (O + 11) / 12;
  const int S = 256 / (HO_ROWS * n_cg);
  const int yl = threadIdx.x % HO_ROWS;
  const int cg = (threadIdx.x / HO_ROWS) % n_cg;
  const int sl = threadIdx.x / (HO_ROWS * n_cg);
  float acc[12];
#pragma unroll
  for (int u = 0; u < 12; u++)
    acc[u] = 0.0f;
  if (sl < S) {
    for (int b = sl; b < nz; b += S) {
      float h = sH[b * HO_ROWS + yl];
      const float4 *e4 = (const float4 *)(sO + (size_t)b * O + cg * 12);
      float4 e0 = e4[0], e1 = e4[1], e2 = e4[2];
      acc[0] += h * e0.x; acc[1] += h * e0.y; acc[2] += h * e0.z; acc[3] += h * e0.w;
      acc[4] += h * e1.x; acc[5] += h * e1.y; acc[6] += h * e1.z; acc[7] += h * e1.w;
      acc[8] += h * e2.x; acc[9] += h * e2.y; acc[10] += h * e2.z; acc[11] += h * e2.w;
    }
#pragma unroll
    for (int u = 0; u < 12; u++)
      red[((size_t)sl * HO_ROWS + yl) * (n_cg * 12) + cg * 12 + u] = acc[u];
  }
  __syncthreads();
  float *mine = parts + (size_t)z * H * O;
  for (int i = threadIdx.x; i < HO_ROWS * O; i += blockDim.x) {
    int y = i / O, o = i - y * O;
    if (y0 + y >= H)
      continue;
    float t = 0.0f;
    for (int q = 0; q < S; q++)
      t += red[((size_t)q * HO_ROWS + y) * (n_cg * 12) + o];
    __stcg(mine + (size_t)(y0 + y) * O + o, t);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    s_last = (atomicAdd(tickets + blockIdx.x, 1u) == (unsigned int)Z - 1);
    if (s_last)
      tickets[blockIdx.x] = 0u;
  }
  __syncthreads();
  if (!s_last)
    return;
  __threadfence();
  for (int i = threadIdx.x; i < HO_ROWS * O; i += blockDim.x) {
    int y = i / O, o = i - y * O;
    if (y0 + y >= H)
      continue;
    size_t idx = (size_t)(y0 + y) * O + o;
    float t = accumulate ? ho_delta[idx] : 0.0f;
    for (int q = 0; q < Z; q++)
      t += __ldcg(parts + (size_t)q * H * O + idx);
    ho_delta[idx] = t;
  }
}

/* a14: the single-net path updates the top layer straight away
   (recur-nn.c:941-964); rows of silent hidden units only decay momentum. */
__global__ void __launch_bounds__(256)
k_sgd_top_apply(RbView v, float *weights, float *momentums, float rate,
    float momentum, float momentum_weight)
{
  const int H = v.d.h_size, O = v.d.o_size;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * O)
    return;
  int s = v.slots[0];
  int y = idx / O, x = idx - y * O;
  float h = (y == 0) ? 1.0f : v.Hd[(size_t)s * H + y];
  float mm = momentums[idx];
  float w = weights[idx];
  if (h != 0.0f) {
    float m = h * rate;
    float d = v.OE[(size_t)s * O + x] * m;
    w += d + mm * momentum_weight;
    mm += d;
  }
  else {
    w += mm * momentum_weight;
  }
  weights[idx] = w;
  momentums[idx] = mm * momentum;
}

/* ------------------------------------------------------------------------ */
/* a10 control: after each chain step, per stream: accumulate cum_error,
   decide whether to walk further (recur-nn.c:383-389) and on the last step
   settle ih_scale / min_error_factor (recur-nn.c:393-413).                 */

__global__ void
k_chain_decide(RbView v, int k)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= v.n)
    return;
  int s = v.slots[j];
  RbScalars *sc = v.sc + s;
  if (!sc->live)
    return;
  float es = 0.0f;
  const float *p = v.partial + (size_t)s * v.n_part;
  for (int q = 0; q < v.n_part; q++)
    es += p[q];
  sc->err_sum = es;
  sc->cum_error += sqrtf(es);
  sc->n_steps = k + 1;
  int t = v.depth - k; /* the reference's loop counter during this step */
  bool stop = (es <= sc->min_sum || es > sc->max_sum);
  bool last = (k == v.depth - 1);
  if (!stop && !last)
    return;
  sc->live = 0;
  int t_left = stop ? t : 0;
  sc->t_left = t_left;
  float ceiling = ERROR_GAIN_CEILING * sc->top_scaled;
  if (es > ceiling) {
    sc->ih_scale = soft_clip_dev(es, sc->max_sum);
  }
  else {
    sc->ih_scale = 1.0f;
    if (sc->adaptive & 1) {
      int depth_error = v.depth / 4 - t_left;
      float min_gain = MIN_ERROR_GAIN * sc->top_scaled;
      float mef = sc->mef;
      if (mef < MAX_MIN_ERROR_FACTOR && (min_gain != sc->min_sum || depth_error < 0))
        mef = (float)((double)mef * (1.0 + depth_error * 1e-3));
      sc->mef = fmaxf(mef, ABS_MIN_ERROR_FACTOR);
    }
  }
}

/* per-stream training parameters the host owns (bptt->learn_rate,
   bptt->min_error_factor, the ADAPTIVE flag) */
__global__ void
k_set_params(RbView v, const float *lr, const float *mef, int adaptive)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= v.n)
    return;
  RbScalars *sc = v.sc + v.slots[j];
  if (lr)
    sc->lr = lr[j];
  if (mef)
    sc->mef = mef[j];
  if (adaptive >= 0)
    sc->adaptive = (sc->adaptive & 2) | (adaptive & 1);
}

/* streams that sit this step out (charmodel-classify.c:124-148: characters
   without a class are run forward but not trained on) */
__global__ void
k_mask_streams(RbView v, const u8 *active)
{
  int j = blockIdx.x;
  int s = v.slots[j];
  bool skip = active && !active[j];
  if (threadIdx.x == 0) {
    int a = v.sc[s].adaptive & 1;
    v.sc[s].adaptive = a | (skip ? 2 : 0);
  }
  if (skip)
    for (int i = threadIdx.x; i < v.d.o_size; i += blockDim.x)
      v.OE[(size_t)s * v.d.o_size + i] = 0.0f;
}

__global__ void
k_set_params_scalar(RbView v, float lr, float mef, int adaptive)
{
  RbScalars *sc = v.sc + v.slots[0];
  sc->lr = lr;
  sc->mef = mef;
  sc->adaptive = adaptive;
}

/* ------------------------------------------------------------------------ */
/* a13: the seven optimisers (recur-nn.c:454-593), elementwise.             */

__global__ void __launch_bounds__(256)
k_apply_learning(int method, float *__restrict__ weights,
    const float *__restrict__ delta, float *__restrict__ momentums,
    float *__restrict__ aux, int size, float rate, float momentum,
    float momentum_weight, const float *rate_scale)
{
  if (rate_scale)
    rate *= *rate_scale;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < size;
       i += gridDim.x * blockDim.x) {
    weights[i] = rb_optimiser_step(method, weights[i], delta[i], momentums, aux, i, rate,
        momentum, momentum_weight);
  }
}

/* ------------------------------------------------------------------------ */
/* a15, a16 and array helpers (recur-nn-helpers.h:22-168)                   */

__global__ void
k_scale(float *a, int n, float s)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    a[i] *= s;
}

__global__ void
k_zero_small(float *a, int n)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float x = a[i];
    a[i] = (fabsf(x) > 1e-34f) ? x : 0.0f;
  }
}

__global__ void
k_clamp(float *a, int n, float lo, float hi)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    a[i] = fminf(fmaxf(a[i], lo), hi);
}

__global__ void
k_fill(float *a, size_t n, float value)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    a[i] = value;
}

__global__ void
k_axpy(float *dst, const float *src, int n, float s, const float *s_dev)
{
  if (s_dev)
    s *= *s_dev;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    dst[i] += src[i] * s;
}

__global__ void
k_add_at(float *a, int index, float value)
{
  a[index] += value;
}

/* recur-nn.c:828-844: shrink the single largest |weight| (first one wins) */
__global__ void __launch_bounds__(1024)
k_tall_poppy(float *a, int n, float threshold, float scale)
{
  __shared__ float s_v[1024];
  __shared__ int s_i[1024];
  float bv = -1.0f;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float x = fabsf(a[i]);
    if (x > bv) {
      bv = x;
      bi = i;
    }
  }
  s_v[threadIdx.x] = bv;
  s_i[threadIdx.x] = bi;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      float ov = s_v[threadIdx.x + o];
      int oi = s_i[threadIdx.x + o];
      if (ov > s_v[threadIdx.x] || (ov == s_v[threadIdx.x] && oi < s_i[threadIdx.x])) {
        s_v[threadIdx.x] = ov;
        s_i[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && s_v[0] > threshold)
    a[s_i[0]] *= scale;
}

__global__ void __launch_bounds__(1024)
k_abs_sum(const float *a, int n, float *out)
{
  __shared__ float scratch[33];
  float t = 0.0f;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    t += fabsf(a[i]);
  t = block_sum(t, scratch);
  if (threadIdx.x == 0)
    *out = t;
}

/* presynaptic noise: each stream draws its row from its own generator, in
   the reference's order (recur-nn-helpers.h:170-181, recur-nn.c:120).       */
__global__ void
k_gen_noise(RbView v, float deviation, int first_col, int n_cols)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= v.n)
    return;
  int s = v.slots[j];
  rb_rng_state st;
  st.a = v.rng[s * 4 + 0];
  st.b = v.rng[s * 4 + 1];
  st.c = v.rng[s * 4 + 2];
  st.d = v.rng[s * 4 + 3];
  float *row = v.noise + (size_t)s * v.d.h_size;
  for (int i = 0; i < n_cols; i++)
    row[first_col + i] = rb_rng_cheap_gaussian(&st) * deviation;
  v.rng[s * 4 + 0] = st.a;
  v.rng[s * 4 + 1] = st.b;
  v.rng[s * 4 + 2] = st.c;
  v.rng[s * 4 + 3] = st.d;
}

/* ------------------------------------------------------------------------ */
/* launchers                                                                  */

#define LAUNCH_CHECK(name) do {                                         \
    cudaError_t e_ = cudaGetLastError();                                \
    if (e_ != cudaSuccess)                                              \
      rb_die("recur-b200: launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    rb_count_launch(1);                                                 \
  } while (0)

static inline int
grid1d(long long n, int block)
{
  long long g = (n + block - 1) / block;
  if (g > 148 * 16)
    g = 148 * 16;
  if (g < 1)
    g = 1;
  return (int)g;
}

extern "C" void
rbk_advance(const RbView *v)
{
  rb_prof_begin(RB_PROF_SMALL);
  k_advance<<<cdiv(v->n, 128), 128, 0, rb_stream>>>(*v);
  LAUNCH_CHECK("k_advance");
  rb_prof_end(RB_PROF_SMALL);
}

extern "C" void
rbk_fill_iota(int *iota, int n)
{
  k_fill_iota<<<cdiv(n, 256), 256, 0, rb_stream>>>(iota, n);
  LAUNCH_CHECK("k_fill_iota");
}

extern "C" void
rbk_set_one_hot(const RbView *v, const u8 *hot_dev)
{
  rb_prof_begin(RB_PROF_SMALL);
  k_set_one_hot<<<v->n, 64, 0, rb_stream>>>(*v, hot_dev);
  LAUNCH_CHECK("k_set_one_hot");
  rb_prof_end(RB_PROF_SMALL);
}

extern "C" void
rbk_set_inputs(const RbView *v, const float *inputs_dev)
{
  rb_prof_begin(RB_PROF_SMALL);
  k_set_inputs<<<v->n, 64, 0, rb_stream>>>(*v, inputs_dev);
  LAUNCH_CHECK("k_set_inputs");
  rb_prof_end(RB_PROF_SMALL);
}

extern "C" void
rbk_text_symbols(const u8 *text_dev, int len, int i, int spacing, int n,
    u8 *cur_dev, u8 *next_dev)
{
  rb_prof_begin(RB_PROF_SMALL);
  k_text_symbols<<<cdiv(n, 128), 128, 0, rb_stream>>>(text_dev, len, i, spacing, n,
      cur_dev, next_dev);
  LAUNCH_CHECK("k_text_symbols");
  rb_prof_end(RB_PROF_SMALL);
}

extern "C" void
rbk_gen_noise(const RbView *v, float deviation, int first_col, int n_cols)
{
  k_gen_noise<<<cdiv(v->n, 32), 32, 0, rb_stream>>>(*v, deviation, first_col, n_cols);
  LAUNCH_CHECK("k_gen_noise");
}

/* The start of a text-predict step for every stream in one launch: advance
   the ring (a1), pick this stream's symbol pair out of the text
   (charmodel-predict.c:295-298), write the one-hot input (a2), build the
   input row [1 | hidden(t-1) | inputs | 0..] with the emergency soft clip
   (a3/a5), and, for the tensor engine, the row's hi/lo planes.             */
struct StepBeginArgs {
  RbView v;
  const u8 *text; /* NULL: symbols are already in cur/next */
  int len, pos, spacing;
  u8 *cur, *next;
  RbPlanes X;     /* hi == NULL: no planes wanted */
  int advance;    /* 0: forward only, the row of the current ring position is rewritten */
};

__global__ void __launch_bounds__(256)
k_step_begin(StepBeginArgs a)
{
  __shared__ float scratch[33];
  __shared__ int s_pos, s_hot;
  const RbView &v = a.v;
  const int j = blockIdx.x;
  const int s = v.contiguous ? v.base + j : v.slots[j];
  if (threadIdx.x == 0) {
    int p = v.pos[s] + (a.advance ? 1 : 0);
    if (p >= v.depth)
      p -= v.depth;
    if (a.advance)
      v.pos[s] = p;
    s_pos = p;
    int hot;
    if (a.text) {
      long long off = (long long)a.pos + (long long)j * a.spacing;
      if (off >= a.len - 1)
        off -= a.len - 1;
      hot = a.text[off];
      a.cur[j] = (u8)hot;
      a.next[j] = a.text[off + 1];
    }
    else {
      hot = a.cur[j];
    }
    s_hot = hot;
  }
  __syncthreads();
  const int I = v.d.i_size, hs1 = v.d.hidden_size + 1;
  const size_t off = ((size_t)s_pos * v.cap + s) * I;
  const size_t prow = ((size_t)s_pos * v.cap + s) * a.X.pitch;
  float *x = v.X + off;
  const float *h = v.Hd + (size_t)s * v.d.h_size;
  const int hot_col = hs1 + s_hot;
  const int in_end = hs1 + v.d.input_size;
  float vals[8]; /* i_size <= 8 * 256 on this path */
  float sum = 0.0f;
#pragma unroll
  for (int u = 0; u < 8; u++) {
    int i = threadIdx.x + u * 256;
    float val = 0.0f;
    if (i < I) {
      if (i == 0)
        val = 1.0f;
      else if (i < hs1)
        val = h[i];
      else if (i < in_end)
        val = (i == hot_col) ? 1.0f : 0.0f;
    }
    vals[u] = val;
    sum += val;
  }
  sum = block_sum(sum, scratch);
  const float softclip = I * INPUT_MEAN_SOFT_TOP;
  const float scale = (sum > softclip) ? soft_clip_dev(sum, softclip) : 1.0f;
#pragma unroll
  for (int u = 0; u < 8; u++) {
    int i = threadIdx.x + u * 256;
    if (i < I) {
      float val = vals[u];
      if (scale != 1.0f)
        val *= scale;
      x[i] = val;
      if (a.X.hi) {
        rb_h16 hi, lo;
        rb_split_f16(val * a.X.scale, hi, lo);
        a.X.hi[prow + i] = hi;
        a.X.lo[prow + i] = lo;
      }
    }
  }
}

extern "C" int
rbk_step_begin_usable(const RbView *v)
{
  return v->d.i_size <= 8 * 256;
}

extern "C" void
rbk_step_begin(const RbView *v, const u8 *text_dev, int len, int pos, int spacing,
    u8 *cur_dev, u8 *next_dev, const RbPlanes *X, int advance)
{
  rbk_step_begin_on(rb_stream, v, text_dev, len, pos, spacing, cur_dev, next_dev, X, advance);
}

extern "C" void
rbk_step_begin_on(cudaStream_t stream, const RbView *v, const u8 *text_dev, int len, int pos,
    int spacing, u8 *cur_dev, u8 *next_dev, const RbPlanes *X, int advance)
{
  StepBeginArgs a;
  a.advance = advance;
  a.v = *v;
  a.text = text_dev;
  a.len = len;
  a.pos = pos;
  a.spacing = spacing;
  a.cur = cur_dev;
  a.next = next_dev;
  if (X)
    a.X = *X;
  else
    memset(&a.X, 0, sizeof(a.X));
  rb_prof_begin(RB_PROF_SMALL);
  k_step_begin<<<v->n, 256, 0, stream>>>(a);
  LAUNCH_CHECK("k_step_begin");
  rb_prof_end(RB_PROF_SMALL);
}

extern "C" void
rbk_prepare_x(const RbView *v)
{
  rb_prof_begin(RB_PROF_SMALL);
  k_prepare_x<<<v->n, 256, 0, rb_stream>>>(*v);
  LAUNCH_CHECK("k_prepare_x");
  rb_prof_end(RB_PROF_SMALL);
}

static size_t
out_multi_smem(const RbView *v)
{
  return ((size_t)v->d.h_size * v->d.o_size + (size_t)OS * v->d.h_size + (size_t)OS * OUT_NT +
      (size_t)2 * OS * v->d.o_size + 8) * sizeof(float);
}

static int
out_multi_usable(const RbView *v)
{
  return v->n >= 4 * OS && out_multi_smem(v) <= 220 * 1024;
}

static void
out_multi_attr(void)
{
  static int attr_done = 0;
  if (!attr_done) {
    /* 227 KB per block less the kernel's static shared memory */
    if (cudaFuncSetAttribute(k_out_multi<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            220 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_out_multi<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            220 * 1024) != cudaSuccess)
      rb_die("recur-b200: k_out_multi cannot have 220 KB of shared memory");
    attr_done = 1;
  }
}

extern "C" int
rbk_output_takes_partials(const RbView *v, int splits)
{
  return out_multi_usable(v) && splits <= 4 && (v->d.h_size % 4) == 0;
}

/* The first phase of k_out_multi<true> by itself: the forward GEMM's split-K
   partial sums -> activation -> hidden rows.  For forward-only runs, where the
   next step needs the hidden rows and nothing else (below). */
__global__ void __launch_bounds__(256)
k_hidden_from_partials(RbView v, RbFwdPartials fp)
{
  const int H = v.d.h_size;
  const int per_row = H / 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n * per_row)
    return;
  const int j = i / per_row, c = (i - j * per_row) * 4;
  const int slot = slot_of(v, j);
  const float *row = fp.part + (size_t)slot * fp.pitch + c;
  float4 h = __ldcg((const float4 *)row);
  for (int z = 1; z < fp.splits; z++) {
    const float4 p = __ldcg((const float4 *)(row + z * fp.split_stride));
    h.x += p.x; h.y += p.y; h.z += p.z; h.w += p.w;
  }
  h = hidden_activation(v, h, c, fp.use_noise ? v.noise + (size_t)slot * H : NULL);
  *(float4 *)(v.Hd + (size_t)slot * H + c) = h;
}

/* Forward-only runs (rnn_batch_text_forward: rnn_opinion step after step):
   step t + 1 needs step t's hidden rows, not its outputs.  With the pipeline
   on, the hidden rows are finished on the library stream and the output
   layer of step t runs on a side stream beside step t + 1's input rows and
   forward GEMM.  rbk_output_pipeline(0) joins. */
static int out_pipe = 0, out_pipe_pending = 0;
static cudaStream_t out_side = NULL;
static cudaEvent_t out_ev_hidden = NULL, out_ev_done = NULL;

extern "C" void
rbk_output_pipeline(int on)
{
  if (!on && out_pipe_pending) {
    cudaStreamWaitEvent(rb_stream, out_ev_done, 0);
    out_pipe_pending = 0;
  }
  out_pipe = on;
}

/* hidden activation + output layer from the split-K partial sums of the
   tensor engine's forward GEMM */
/* A caller about to run a forward pass whose outputs go straight into the
   char model's softmax error may ask for that to happen in the output kernel;
   rbk_fused_loss_done() afterwards says whether it did. */
static struct {
  RbLossArgs args;
  int armed, done;
  int want_top; /* the caller trains on this forward pass: fuse the top layer too */
} loss_request;

extern "C" void
rbk_request_fused_top(int on)
{
  loss_request.want_top = on;
}

extern "C" void
rbk_request_fused_loss(const u8 *target_dev, float *err_dev, int *winner_dev,
    RbCharAccum *accum_dev, RbCharAccum *snapshot_host, int reset)
{
  loss_request.args.snapshot = snapshot_host;
  loss_request.args.reset = reset;
  loss_request.args.fuse_top = 0;
  loss_request.want_top = 0;
  loss_request.args.target = target_dev;
  loss_request.args.err = err_dev;
  loss_request.args.winner = winner_dev;
  loss_request.args.accum = accum_dev;
  loss_request.armed = target_dev && err_dev && winner_dev && accum_dev;
  loss_request.done = 0;
}

extern "C" int
rbk_fused_loss_done(void)
{
  int done = loss_request.done;
  loss_request.armed = loss_request.done = 0;
  return done;
}

static int top_fused_pending = 0;
static const RbPool *top_fused_pool = NULL;
static int top_fused_base = 0, top_fused_n = 0;

/* did the last forward pass already run the top layer for this batch (E(0),
   its clip and the walk's thresholds)?  Asked once by whoever would otherwise
   launch it. */
extern "C" int
rbk_top_was_fused(const RbView *v)
{
  int f = top_fused_pending && top_fused_pool == v->pool && top_fused_base == v->base &&
      top_fused_n == v->n && v->contiguous;
  top_fused_pending = 0;
  return f;
}

extern "C" void
rbk_output_from_partials(const RbView *v, const RbFwdPartials *fp)
{
  RbLossArgs loss = {NULL, NULL, NULL, NULL, NULL, 0, 0, NULL};
  top_fused_pending = 0;
  static unsigned long long *dbg_dev = NULL;
  static int dbg_calls = 0;
  const bool timing = getenv("RECUR_B200_OUT_TIMING") != NULL;
  if (loss_request.armed && v->contiguous) {
    loss = loss_request.args;
    loss_request.done = 1;
    /* training step of the char model on a batch: the top layer rides along */
    if (loss_request.want_top && v->pool && v->pool->has_bptt && v->n >= 4 * OS &&
        v->d.h_size <= 2048 && !v->CIE && !getenv("RECUR_B200_NO_FUSED_TOP")) {
      loss.fuse_top = 1;
      top_fused_pending = 1;
      top_fused_pool = v->pool;
      top_fused_base = v->base;
      top_fused_n = v->n;
    }
  }
  if (timing) {
    if (!dbg_dev)
      cudaMalloc((void **)&dbg_dev, 16 * sizeof(unsigned long long));
    loss.dbg = dbg_dev;
  }
  out_multi_attr();
  if (out_pipe && !loss.target && !rb_prof_active() && (v->d.h_size % 4) == 0) {
    if (!out_side) {
      cudaStreamCreateWithFlags(&out_side, cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&out_ev_hidden, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&out_ev_done, cudaEventDisableTiming);
    }
    if (out_pipe_pending) /* the last step's output layer is still reading the hidden rows */
      cudaStreamWaitEvent(rb_stream, out_ev_done, 0);
    k_hidden_from_partials<<<cdiv(v->n * (v->d.h_size / 4), 256), 256, 0, rb_stream>>>(*v, *fp);
    LAUNCH_CHECK("k_hidden_from_partials");
    cudaEventRecord(out_ev_hidden, rb_stream);
    cudaStreamWaitEvent(out_side, out_ev_hidden, 0);
    RbFwdPartials none = {NULL, 0, 0, 0, 0};
    k_out_multi<false><<<cdiv(v->n, OS), OUT_NT, out_multi_smem(v), out_side>>>(*v, none, loss);
    LAUNCH_CHECK("k_out_multi");
    cudaEventRecord(out_ev_done, out_side);
    out_pipe_pending = 1;
    return;
  }
  rb_prof_begin(RB_PROF_OUT);
  k_out_multi<true><<<cdiv(v->n, OS), OUT_NT, out_multi_smem(v), rb_stream>>>(*v, *fp, loss);
  LAUNCH_CHECK("k_out_multi<partials>");
  rb_prof_end(RB_PROF_OUT);
  if (timing && loss.target && (++dbg_calls % 100) == 60) {
    unsigned long long h[16];
    cudaMemcpyAsync(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost, rb_stream);
    cudaStreamSynchronize(rb_stream);
    fprintf(stderr, "k_out_multi block 0, us after its start: rows summed %.2f, Who landed %.2f, "
        "outputs done %.2f, softmax done %.2f, (error rows in %.2f, dot products %.2f, "
        "sums %.2f, clip + scalars %.2f) top layer done %.2f, ticket taken %.2f\n",
        (h[1] - h[0]) * 1e-3, (h[2] - h[0]) * 1e-3, (h[3] - h[0]) * 1e-3, (h[4] - h[0]) * 1e-3,
        (h[7] - h[0]) * 1e-3, (h[8] - h[0]) * 1e-3, (h[9] - h[0]) * 1e-3, (h[10] - h[0]) * 1e-3,
        (h[5] - h[0]) * 1e-3, (h[6] - h[0]) * 1e-3);
  }
}

extern "C" void
rbk_output(const RbView *v)
{
  if (out_multi_usable(v)) {
    RbFwdPartials none = {NULL, 0, 0, 0, 0};
    RbLossArgs no_loss = {NULL, NULL, NULL, NULL, NULL, 0, 0, NULL};
    out_multi_attr();
    rb_prof_begin(RB_PROF_OUT);
    k_out_multi<false><<<cdiv(v->n, OS), OUT_NT, out_multi_smem(v), rb_stream>>>(*v, none, no_loss);
    LAUNCH_CHECK("k_out_multi");
    rb_prof_end(RB_PROF_OUT);
    return;
  }
  size_t sh = (size_t)(v->d.h_size + 256 + 8) * sizeof(float);
  int col_chunks = (v->d.o_size > 256) ? cdiv(v->d.o_size, 256) : 1;
  if (col_chunks > 64)
    col_chunks = 64;
  rb_prof_begin(RB_PROF_OUT);
  k_out<<<dim3(v->n, col_chunks), 256, sh, rb_stream>>>(*v);
  LAUNCH_CHECK("k_out");
  rb_prof_end(RB_PROF_OUT);
}

extern "C" void
rbk_chain_decide(const RbView *v, int k)
{
  k_chain_decide<<<cdiv(v->n, 128), 128, 0, rb_stream>>>(*v, k);
  LAUNCH_CHECK("k_chain_decide");
}

extern "C" void
rbk_forward(const RbView *v, float presynaptic_noise)
{
  rbk_prepare_x(v);
  rbk_forward_core(v, presynaptic_noise);
}

/* everything of a3 after the input row is in place */
extern "C" void
rbk_forward_core(const RbView *v, float presynaptic_noise)
{
  GemmArgs g;
  g.v = *v;
  g.k = 0;
  g.delta = NULL;
  g.accumulate = 0;
  g.use_noise = 0;
  if (presynaptic_noise != 0.0f) {
    rbk_gen_noise(v, presynaptic_noise, 1, v->d.h_size - 1);
    g.use_noise = 1;
  }
  dim3 grid(cdiv(v->d.h_size, TN), cdiv(v->n, TM));
  rb_prof_begin(RB_PROF_FWD);
  k_gemm<G_FWD><<<grid, 256, 0, rb_stream>>>(g);
  LAUNCH_CHECK("k_gemm<FWD>");
  rb_prof_end(RB_PROF_FWD);
  rbk_output(v);
}

static int *rb_winner_scratch = NULL;
static float *rb_err_scratch = NULL;
static int rb_scratch_cap = 0;

extern "C" void
rbk_softmax_error(const RbView *v, const u8 *target_dev, float *err_dev,
    int *winner_dev, RbCharAccum *accum_dev)
{
  if (accum_dev && (!err_dev || !winner_dev)) {
    if (rb_scratch_cap < v->n) {
      if (rb_winner_scratch) {
        cudaFree(rb_winner_scratch);
        cudaFree(rb_err_scratch);
      }
      rb_scratch_cap = v->n + 64;
      if (cudaMalloc(&rb_winner_scratch, rb_scratch_cap * sizeof(int)) != cudaSuccess ||
          cudaMalloc(&rb_err_scratch, rb_scratch_cap * sizeof(float)) != cudaSuccess)
        rb_die("recur-b200: out of device memory for softmax scratch");
    }
    if (!err_dev)
      err_dev = rb_err_scratch;
    if (!winner_dev)
      winner_dev = rb_winner_scratch;
  }
  rb_prof_begin(RB_PROF_SMALL);
  k_softmax_error<<<cdiv(v->n, 4), 128, 0, rb_stream>>>(*v, target_dev, err_dev, winner_dev);
  LAUNCH_CHECK("k_softmax_error");
  rb_prof_end(RB_PROF_SMALL);
  if (accum_dev) {
    rb_prof_begin(RB_PROF_SMALL);
    k_char_accum<<<1, 256, 0, rb_stream>>>(err_dev, winner_dev, target_dev, v->n, accum_dev);
    LAUNCH_CHECK("k_char_accum");
    rb_prof_end(RB_PROF_SMALL);
  }
}

static int ho_slab_attr_done = 0;
static cudaStream_t top_side = NULL;
static cudaEvent_t top_ev_fork = NULL, top_ev_join = NULL;
static int top_forked = 0;
static int top_defer = 0, top_deferred = 0, top_deferred_accumulate = 0;
static RbView top_deferred_v;
static float *top_deferred_delta = NULL;

/* ho_delta for a batch, on `stream` */
static void
launch_ho_delta_batch(const RbView *v, float *ho_delta, int accumulate, cudaStream_t stream)
{
  /* big batches: the streams are split over blocks too, so that several
     small blocks share an SM instead of one slab owning it */
  int Z = v->n / 128;
  if (Z > 4)
    Z = 4;
  if (Z >= 2) {
    static float *parts = NULL;
    static unsigned int *tickets = NULL;
    static size_t parts_cap = 0;
    const size_t need = (size_t)4 * v->d.h_size * v->d.o_size;
    if (parts_cap < need) {
      cudaStreamSynchronize(rb_stream);
      if (top_side)
        cudaStreamSynchronize(top_side);
      cudaFree(parts);
      cudaFree(tickets);
      if (cudaMalloc(&parts, need * sizeof(float)) != cudaSuccess ||
          cudaMalloc(&tickets, 4096 * sizeof(unsigned int)) != cudaSuccess)
        rb_die("recur-b200: out of device memory for the ho_delta parts");
      cudaMemset(tickets, 0, 4096 * sizeof(unsigned int));
      parts_cap = need;
    }
    const int per_z = (v->n + Z - 1) / Z;
    size_t smem = ((size_t)per_z * (HO_ROWS + v->d.o_size) + 16 + (size_t)256 * 12) * sizeof(float);
    static int attr_done = 0;
    if (!attr_done) {
      cudaFuncSetAttribute(k_ho_delta_split, cudaFuncAttributeMaxDynamicSharedMemorySize,
          100 * 1024);
      attr_done = 1;
    }
    if (smem <= 100 * 1024 && cdiv(v->d.h_size, HO_ROWS) <= 4096) {
      k_ho_delta_split<<<dim3(cdiv(v->d.h_size, HO_ROWS), Z), 256, smem, stream>>>(*v, ho_delta,
          accumulate, parts, tickets);
      LAUNCH_CHECK("k_ho_delta_split");
      return;
    }
  }
  size_t slab = ((size_t)v->n * (HO_ROWS + v->d.o_size) + 16 + (size_t)256 * 12) * sizeof(float);
  if (!ho_slab_attr_done) {
    cudaFuncSetAttribute(k_ho_delta_slab, cudaFuncAttributeMaxDynamicSharedMemorySize,
        200 * 1024);
    ho_slab_attr_done = 1;
  }
  k_ho_delta_slab<<<cdiv(v->d.h_size, HO_ROWS), 256, slab, stream>>>(*v, ho_delta, accumulate);
  LAUNCH_CHECK("k_ho_delta_slab");
}

/* The tensor engine's walk is one kernel on most of the SMs: ho_delta is
   launched after it (so that the walk's CTAs are placed first) and runs on the
   rest.  rbk_top_layer_defer_ho_delta(1) before rbk_top_layer_begin, then
   rbk_top_layer_ho_delta_now() once the walk is queued. */
extern "C" void
rbk_top_layer_defer_ho_delta(int on)
{
  top_defer = on;
}

/* move the point the deferred ho_delta waits for to "now" on the library
   stream: with it right in front of the walk's kernel, both become ready at
   the same moment and the stream priorities decide who gets the SMs first */
extern "C" void
rbk_top_layer_mark(void)
{
  if (top_deferred)
    cudaEventRecord(top_ev_fork, rb_stream);
}

extern "C" void
rbk_top_layer_ho_delta_now(void)
{
  if (!top_deferred)
    return;
  top_deferred = 0;
  cudaStreamWaitEvent(top_side, top_ev_fork, 0);
  launch_ho_delta_batch(&top_deferred_v, top_deferred_delta, top_deferred_accumulate, top_side);
  cudaEventRecord(top_ev_join, top_side);
  top_forked = 1;
}

/* wait (on the library stream) for what rbk_top_layer_begin left running
   beside it */
extern "C" void
rbk_top_layer_join(void)
{
  if (top_forked)
    cudaStreamWaitEvent(rb_stream, top_ev_join, 0);
  top_forked = 0;
}

extern "C" void
rbk_top_layer(const RbView *v, float *ho_delta, int accumulate,
    const RecurErrorRange *ranges_dev, int n_ranges)
{
  rbk_top_layer_begin(v, ho_delta, accumulate, ranges_dev, n_ranges);
  rbk_top_layer_join();
}

/* a7..a9.  Returns with the top-layer kernels queued on the library stream
   and, for a batch, ho_delta possibly still running on a side stream:
   rbk_top_layer_join() before anything reads ho_delta. */
extern "C" void
rbk_top_layer_begin(const RbView *v, float *ho_delta, int accumulate,
    const RecurErrorRange *ranges_dev, int n_ranges)
{
  /* a9 reads the hidden rows and the output errors, a7/a8 the same plus Who:
     neither needs the other, both are too small to fill the GPU.  In a batch
     the ho_delta kernel runs on a side stream next to the top-layer kernels
     (not while the per-class profiler is timing them). */
  const bool fused = rbk_top_was_fused(v) && n_ranges == 0; /* the forward pass did a7/a8 */
  size_t slab = ((size_t)v->n * (HO_ROWS + v->d.o_size) + 16 + (size_t)256 * 12) * sizeof(float);
  const bool slab_ok = ho_delta && n_ranges == 0 && v->n >= 8 && slab <= 200 * 1024 &&
      v->d.o_size <= 384;
  bool forked = false;
  if (slab_ok && v->n >= 4 * OS && !rb_prof_active()) {
    if (!top_side) {
      /* lowest priority: when the walk's kernel is waiting for SMs too, it goes first */
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      cudaStreamCreateWithPriority(&top_side, cudaStreamNonBlocking, lo);
      cudaEventCreateWithFlags(&top_ev_fork, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&top_ev_join, cudaEventDisableTiming);
    }
    cudaEventRecord(top_ev_fork, rb_stream);
    if (top_defer) {
      /* the caller launches its walk first and lets ho_delta follow it on to
         the SMs the walk leaves free: rbk_top_layer_ho_delta_now() */
      top_deferred_v = *v;
      top_deferred_delta = ho_delta;
      top_deferred_accumulate = accumulate;
      top_deferred = 1;
    }
    else {
      cudaStreamWaitEvent(top_side, top_ev_fork, 0);
      launch_ho_delta_batch(v, ho_delta, accumulate, top_side);
      cudaEventRecord(top_ev_join, top_side);
      top_forked = 1;
    }
    forked = true;
  }
  if (fused) {
    /* E(0), its clip and the thresholds are in place already */
  }
  else
  if (n_ranges == 0 && v->n >= 4 * OS && v->pool && v->pool->has_bptt) {
    /* a batch: E(0) = mask * (o_error . Who^T) as a tiled contraction */
    GemmArgs g;
    g.v = *v;
    g.k = 0;
    g.delta = NULL;
    g.accumulate = 0;
    g.use_noise = 0;
    dim3 grid(cdiv(v->d.h_size, TN), cdiv(v->n, TM));
    rb_prof_begin(RB_PROF_TOP);
    k_gemm<G_TOP><<<grid, 256, 0, rb_stream>>>(g);
    LAUNCH_CHECK("k_gemm<TOP>");
    k_top_finish<<<v->n, 256, 0, rb_stream>>>(*v, (int)grid.x);
    LAUNCH_CHECK("k_top_finish");
    rb_prof_end(RB_PROF_TOP);
  }
  else {
    size_t sh = (size_t)(v->d.o_size + 40) * sizeof(float);
    rb_prof_begin(RB_PROF_TOP);
    k_top<<<v->n, 256, sh, rb_stream>>>(*v, ranges_dev, n_ranges);
    LAUNCH_CHECK("k_top");
    rb_prof_end(RB_PROF_TOP);
  }
  if (forked)
    return;
  /* hidden slab + errors + slack + per-slice partial sums */
  if (slab_ok) {
    rb_prof_begin(RB_PROF_HO);
    launch_ho_delta_batch(v, ho_delta, accumulate, rb_stream);
    rb_prof_end(RB_PROF_HO);
  }
  else if (ho_delta && n_ranges == 0 && v->n >= 8) {
    /* the sum over streams as a tiled contraction */
    GemmArgs g;
    g.v = *v;
    g.k = 0;
    g.delta = ho_delta;
    g.accumulate = accumulate;
    g.use_noise = 0;
    dim3 grid(cdiv(v->d.o_size, TN), cdiv(v->d.h_size, TM));
    rb_prof_begin(RB_PROF_HO);
    k_gemm<G_HO><<<grid, 256, 0, rb_stream>>>(g);
    LAUNCH_CHECK("k_gemm<HO>");
    rb_prof_end(RB_PROF_HO);
  }
  else if (ho_delta) {
    int total = v->d.h_size * v->d.o_size;
    rb_prof_begin(RB_PROF_HO);
    k_ho_delta<<<cdiv(total, 256), 256, 0, rb_stream>>>(*v, ho_delta, accumulate,
        ranges_dev, n_ranges);
    LAUNCH_CHECK("k_ho_delta");
    rb_prof_end(RB_PROF_HO);
  }
}

extern "C" void
rbk_sgd_top_apply(const RbView *v, float *ho_weights, float *ho_momentum,
    float rate, float momentum, float momentum_weight)
{
  int total = v->d.h_size * v->d.o_size;
  k_sgd_top_apply<<<cdiv(total, 256), 256, 0, rb_stream>>>(*v, ho_weights, ho_momentum,
      rate, momentum, momentum_weight);
  LAUNCH_CHECK("k_sgd_top_apply");
}

/* ------------------------------------------------------------------------ */
/* a10 + a11 for ONE stream (the single-net path of config 1 and every
 * per-net rnn_bptt_calc_deltas call): the whole truncated BPTT walk of
 * recur-nn.c:303-450 in one launch.
 *
 * A single stream makes every step a matrix-vector product that depends on
 * the previous one; launched per step the walk is ~60 launches of ~5 us for
 * ~100 k multiply-adds each.  Here a cluster of eight CTAs stays resident for
 * the walk with the weights in shared memory: CTA c owns input rows
 * [c*R, (c+1)*R) of Wih (R*h_size floats) and the same rows of the gradient.
 * Per step, one warp per owned row y: if x_k[y] is zero the row is skipped
 * (the reference's skip of rows multiplied by zero, recur-nn.c:347; the test
 * is warp-uniform), else   G[y,:] += x_k[y] * E_k[:]   and
 * e = Wih[y,:] . E_k[:]   by warp shuffle.  The new error vector is pushed
 * into every CTA's shared memory through distributed shared memory, the sum
 * of squares likewise, one cluster barrier per step; every CTA then takes
 * the same stop decision (recur-nn.c:383-413).  At the end each CTA writes
 * its rows of ih_scale * G to the delta array.                              */

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define WALK_CTAS 8

/* rnn_bptt_calculate (a14, recur-nn.c:919-1019) folds more into the same
   launch: the top layer's error (a7, a8) with the immediate update of Who in
   front of the walk, the update of Wih behind it, o_error read from and the
   stream's scalars written to pinned host memory. */
struct WalkFuse {
  int enabled;
  const float *o_error_in;  /* pinned mirror */
  float lr, mef;
  int adaptive;
  float *ho_w, *ho_mom, *ih_w, *ih_mom;
  float momentum, momentum_weight;
  RbScalars *sc_out;        /* pinned */
};

struct WalkArgs {
  RbView v;
  float *delta;
  int accumulate;
  int rows_per; /* R */
  size_t plane_stride; /* floats between the output planes of two streams (0: one stream) */
  WalkFuse f;
};

/* REG: nets of up to 256 x 256 (config 1 is 244 x 200) keep the owned weight
   rows AND their gradient in registers - a warp owns four rows, a lane every
   32nd column - so a step reads nothing but the error vector from shared
   memory; larger nets keep both in shared memory. */
#define WALK_RW 4 /* rows per warp */
#define WALK_NC 8 /* columns per lane */

template <bool REG>
__global__ void __cluster_dims__(WALK_CTAS, 1, 1) __launch_bounds__(256, 1)
k_walk_single(WalkArgs a)
{
  extern __shared__ __align__(16) float wsh[];
  cg::cluster_group cluster = cg::this_cluster();
  const RbView &v = a.v;
  /* one cluster per stream; with several streams each writes its scaled
     gradient to its own plane and k_sum_stream_planes adds them in stream order */
  const int stream = blockIdx.x / WALK_CTAS;
  const int s = v.slots[stream];
  a.delta += (size_t)stream * a.plane_stride;
  const int I = v.d.i_size, H = v.d.h_size, hs1 = v.d.hidden_size + 1;
  const int R = a.rows_per;
  const int rank = (int)cluster.block_rank();
  const int y0 = min(I, rank * R), y1 = min(I, y0 + R), nr = y1 - y0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  float *W = wsh;                    /* [R][H] owned weight rows (not REG) */
  float *G = W + (REG ? 0 : (size_t)R * H); /* [R][H] their gradient (not REG) */
  float *E = G + (REG ? 0 : (size_t)R * H); /* [3][H] error vectors, rotating */
  float *ES = E + 3 * H;             /* [3][WALK_CTAS] partial sums of squares */
  float *xs = ES + 3 * WALK_CTAS;    /* [depth][R] this CTA's slice of every ring row of the walk */
  float *red = xs + (size_t)v.depth * R; /* [32] sums of squares per (warp, row), then [o_size]
                                            output errors */
  float *eh = red + 32 + v.d.o_size;      /* [depth][R] the owned entries of E(1..depth): written
                                            to the pool after the walk, so that no global store
                                            (and the barrier's wait for it) sits inside a step */

  RbScalars sc = v.sc[s];
  const int pos = v.pos[s];
  const int depth = v.depth;

  float Wr[WALK_RW][WALK_NC], Gr[WALK_RW][WALK_NC];
  if (REG) {
#pragma unroll
    for (int u = 0; u < WALK_RW; u++) {
      const int r = warp + 8 * u;
#pragma unroll
      for (int i = 0; i < WALK_NC; i++) {
        const int c = lane + 32 * i;
        Wr[u][i] = (r < nr && c < H) ? v.Wih[(size_t)(y0 + r) * H + c] : 0.0f;
        Gr[u][i] = 0.0f;
      }
    }
  }
  else {
    for (int i = threadIdx.x * 4; i < nr * H; i += blockDim.x * 4) {
      *(float4 *)(W + i) = *(const float4 *)(v.Wih + (size_t)y0 * H + i);
      *(float4 *)(G + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  /* the walk's inputs do not depend on the walk: fetch them all now, one
     round trip instead of one per step */
  for (int idx = threadIdx.x; idx < depth * nr; idx += blockDim.x) {
    int step = idx / nr, r = idx - step * nr;
    int p = pos - step;
    if (p < 0)
      p += depth;
    xs[step * R + r] = v.X[((size_t)p * v.cap + s) * I + y0 + r];
  }
  if (!a.f.enabled) {
    const float *e0 = e_row(v, s, 0);
    for (int i = threadIdx.x; i < H; i += blockDim.x)
      E[i] = e0[i];
  }
  else {
    /* a7/a8 as k_top does them (dense error), by every CTA for itself: the
       same operations in the same order give the same bits everywhere */
    __shared__ float scratch[33];
    const int O = v.d.o_size;
    float *oe = red + 32; /* [O] */
    for (int i = threadIdx.x; i < O; i += blockDim.x) {
      float e = a.f.o_error_in[i];
      oe[i] = e;
      if (rank == 0)
        v.OE[(size_t)s * O + i] = e;
    }
    __syncthreads();
    const float *hid = v.Hd + (size_t)s * H;
    float *e0g = e_row(v, s, 0);
    float abs_sum = 0.0f, hsum = 0.0f, hmag = 0.0f;
    int hzero = 0;
    for (int y = threadIdx.x; y < I; y += blockDim.x) {
      float e = 0.0f;
      if (y < H) {
        float hv = hid[y];
        hsum += hv;
        hmag += hv * hv;
        hzero += (hv == 0.0f);
        if (y >= 1 && hv != 0.0f) {
          const float *row = v.Who + (size_t)y * O;
          for (int x = 0; x < O; x += 4) {
            float4 w = *(const float4 *)(row + x);
            e += w.x * oe[x] + w.y * oe[x + 1] + w.z * oe[x + 2] + w.w * oe[x + 3];
          }
          abs_sum += fabsf(e);
        }
      }
      if (y > v.d.hidden_size)
        e = 0.0f;
      if (y < H)
        E[y] = e;
      if (rank == 0)
        e0g[y] = e;
    }
    abs_sum = block_sum(abs_sum, scratch);
    hsum = block_sum(hsum, scratch);
    hmag = block_sum(hmag, scratch);
    float hz = block_sum((float)hzero, scratch);
    const float halfmax = H * MAX_TOP_ERROR_FACTOR;
    float top_scaled = abs_sum;
    if (abs_sum > halfmax) {
      const float scale = soft_clip_dev(abs_sum, halfmax);
      for (int y = threadIdx.x; y < H; y += blockDim.x) {
        E[y] *= scale;
        if (rank == 0)
          e0g[y] = E[y];
      }
      top_scaled = scale * abs_sum;
    }
    sc.lr = a.f.lr;
    sc.mef = a.f.mef;
    sc.adaptive = (sc.adaptive & 2) | (a.f.adaptive & 1);
    sc.top_raw = abs_sum;
    sc.top_scaled = top_scaled;
    sc.hidden_sum = hsum;
    sc.hidden_mag = sqrtf(hmag);
    sc.hidden_zeros = (int)(hz + 0.5f);
    sc.min_sum = fminf(sc.mef / sc.lr, MIN_ERROR_GAIN * top_scaled);
    sc.max_sum = MAX_ERROR_GAIN * top_scaled + 1.0f;
    sc.cum_error = 0.0f;
    sc.err_sum = 0.0f;
    sc.live = (depth > 0) && !(sc.adaptive & 2);
    sc.n_steps = 0;
    sc.t_left = depth;
    sc.ih_scale = 1.0f;
  }
  const bool walk = sc.live != 0 && !(sc.adaptive & 2);
  cluster.sync(); /* everybody's shared memory is set up before anyone pushes into it
                     (and, fused: everybody has read Who) */
  if (a.f.enabled) {
    /* apply_sgd_top_layer's immediate update of Who (recur-nn.c:941-964);
       rows of silent hidden units only decay their momentum */
    const int O = v.d.o_size;
    const float *oe = red + 32;
    for (int idx = rank * blockDim.x + threadIdx.x; idx < H * O; idx += WALK_CTAS * blockDim.x) {
      int y = idx / O, xo = idx - y * O;
      float hv = (y == 0) ? 1.0f : v.Hd[(size_t)s * H + y];
      float mm = a.f.ho_mom[idx];
      float w = a.f.ho_w[idx];
      if (hv != 0.0f) {
        float d = oe[xo] * (hv * a.f.lr);
        w += d + mm * a.f.momentum_weight;
        mm += d;
      }
      else {
        w += mm * a.f.momentum_weight;
      }
      a.f.ho_w[idx] = w;
      a.f.ho_mom[idx] = mm * a.f.momentum;
    }
  }

  if (walk) {
    for (int k = 0; k < depth; k++) {
      const float *cur = E + (k % 3) * H;
      const int nb = (k + 1) % 3;
      const float *x_now = xs + k * R;
      float *e_hist = eh + (size_t)k * R;
      float sq4 = 0.0f; /* squares of the row this lane group finishes */
      /* four rows per warp at a time, so that their shared-memory reads and
         shuffle reductions overlap */
      for (int rb = warp; rb < nr; rb += 32) {
        float xv[4], dot[4];
        bool act[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          int r = rb + 8 * u;
          xv[u] = (r < nr) ? x_now[r] : 0.0f;
          act[u] = xv[u] != 0.0f && (v.activation != RNN_RECLIP20 || xv[u] < 20.0f);
          dot[u] = 0.0f;
        }
        if (REG) {
          /* rb == warp here (R <= 32): the rows are the ones in registers;
             silent rows ride along with a zero multiplier */
          float xa[4];
#pragma unroll
          for (int u = 0; u < 4; u++)
            xa[u] = act[u] ? xv[u] : 0.0f;
          float ec[WALK_NC]; /* all reads of the error vector first, then the arithmetic */
#pragma unroll
          for (int i = 0; i < WALK_NC; i++) {
            const int c = lane + 32 * i;
            ec[i] = (c < H) ? cur[c] : 0.0f;
          }
#pragma unroll
          for (int i = 0; i < WALK_NC; i++) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
              dot[u] = fmaf(Wr[u][i], ec[i], dot[u]);
              Gr[u][i] = fmaf(xa[u], ec[i], Gr[u][i]);
            }
          }
        }
        else {
          for (int c = lane; c < H; c += 32) {
            const float ec = cur[c];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              if (act[u]) { /* uniform over the warp */
                const size_t o = (size_t)(rb + 8 * u) * H + c;
                dot[u] = fmaf(W[o], ec, dot[u]);
                G[o] = fmaf(xv[u], ec, G[o]);
              }
            }
          }
        }
        /* four sums in six shuffles: after the first two exchanges the
           eight lanes of group g = lane / 8 hold the partial sums of row g only */
        float mine;
        {
          const bool hi = lane & 16;
          float k0 = hi ? dot[2] : dot[0], k1 = hi ? dot[3] : dot[1];
          float s0 = hi ? dot[0] : dot[2], s1 = hi ? dot[1] : dot[3];
          k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
          k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
          const bool mid = lane & 8;
          mine = mid ? k1 : k0;
          const float send = mid ? k0 : k1;
          mine += __shfl_xor_sync(0xffffffffu, send, 8);
          mine += __shfl_xor_sync(0xffffffffu, mine, 4);
          mine += __shfl_xor_sync(0xffffffffu, mine, 2);
          mine += __shfl_xor_sync(0xffffffffu, mine, 1);
        }
        /* the four rows finish side by side: lane = 8 * row + target CTA */
        {
          const int g = lane >> 3, q = lane & 7;
          const int r = rb + 8 * g, y = y0 + r;
          float e = 0.0f;
          if (r < nr) {
            const float x = x_now[r];
            if (x != 0.0f && (v.activation != RNN_RECLIP20 || x < 20.0f)) {
              e = mine;
              if (v.activation == RNN_RESQRT)
                e /= 2.0f * (x + 1.0f);
            }
          }
          const float e_store = (y == 0 || (y >= hs1 && y < H)) ? 0.0f : e;
          if (q == 0 && r < nr) {
            sq4 = fmaf(e, e, sq4);
            e_hist[r] = e_store;
            if (v.CIE && y >= hs1 && y < hs1 + v.d.input_size)
              v.CIE[(size_t)s * v.bl_o + y - hs1] += e;
          }
          /* ONE store instruction pushes the four new entries into all eight CTAs */
          if (r < nr && y < H)
            cluster.map_shared_rank(E, q)[nb * H + y] = e_store;
        }
      }
      if ((lane & 7) == 0)
        red[warp * 4 + (lane >> 3)] = sq4;
      __syncthreads();
      if (threadIdx.x < WALK_CTAS) {
        float t = 0.0f;
#pragma unroll
        for (int q = 0; q < 32; q++)
          t += red[q];
        cluster.map_shared_rank(ES, threadIdx.x)[nb * WALK_CTAS + rank] = t;
      }
      cluster.sync();
      float es = 0.0f;
#pragma unroll
      for (int q = 0; q < WALK_CTAS; q++)
        es += ES[nb * WALK_CTAS + q];
      /* recur-nn.c:383-413, identically in every thread of every CTA */
      sc.err_sum = es;
      sc.cum_error += sqrtf(es);
      sc.n_steps = k + 1;
      const int t_loop = depth - k;
      const bool stop = (es <= sc.min_sum || es > sc.max_sum);
      const bool last = (k == depth - 1);
      if (stop || last) {
        sc.live = 0;
        const int t_left = stop ? t_loop : 0;
        sc.t_left = t_left;
        const float ceiling = ERROR_GAIN_CEILING * sc.top_scaled;
        if (es > ceiling) {
          sc.ih_scale = soft_clip_dev(es, sc.max_sum);
        }
        else {
          sc.ih_scale = 1.0f;
          if (sc.adaptive & 1) {
            int depth_error = depth / 4 - t_left;
            float min_gain = MIN_ERROR_GAIN * sc.top_scaled;
            float mef = sc.mef;
            if (mef < MAX_MIN_ERROR_FACTOR && (min_gain != sc.min_sum || depth_error < 0))
              mef = (float)((double)mef * (1.0 + depth_error * 1e-3));
            sc.mef = fmaxf(mef, ABS_MIN_ERROR_FACTOR);
          }
        }
        break;
      }
    }
  }
  __syncthreads();
  /* the error rows of the executed steps, for whoever looks at the pool later
     (mirrors, the stale-row rule of sparse top layers) */
  for (int idx = threadIdx.x; idx < sc.n_steps * nr && walk; idx += blockDim.x) {
    int step = idx / nr, r = idx - step * nr;
    e_row(v, s, step + 1)[y0 + r] = eh[(size_t)step * R + r];
  }
  /* a11: ih_delta (+)= ih_scale * G on the owned rows */
  const float scale = walk ? sc.ih_scale : 0.0f;
  if (REG) {
#pragma unroll
    for (int u = 0; u < WALK_RW; u++) {
      const int r = warp + 8 * u;
#pragma unroll
      for (int i = 0; i < WALK_NC; i++) {
        const int c = lane + 32 * i;
        if (r < nr && c < H) {
          const size_t gi = (size_t)(y0 + r) * H + c;
          float d = fmaf(scale, Gr[u][i], a.accumulate ? a.delta[gi] : 0.0f);
          a.delta[gi] = d;
          if (a.f.enabled) {
            /* apply_sgd_with_bptt: the weighted-momentum step of a13, old
               weight still in its register */
            float t = d * a.f.lr;
            float m = a.f.ih_mom[gi];
            a.f.ih_w[gi] = Wr[u][i] + (t + m * a.f.momentum_weight);
            a.f.ih_mom[gi] = (m + t) * a.f.momentum;
          }
        }
      }
    }
  }
  else {
    for (int i = threadIdx.x * 4; i < nr * H; i += blockDim.x * 4) {
      float4 gv = *(const float4 *)(G + i);
      float4 o = a.accumulate ? *(const float4 *)(a.delta + (size_t)y0 * H + i)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      o.x = fmaf(scale, gv.x, o.x);
      o.y = fmaf(scale, gv.y, o.y);
      o.z = fmaf(scale, gv.z, o.z);
      o.w = fmaf(scale, gv.w, o.w);
      *(float4 *)(a.delta + (size_t)y0 * H + i) = o;
      if (a.f.enabled) {
        /* apply_sgd_with_bptt: the weighted-momentum step of a13 on the owned
           rows, old weights still in shared memory */
        const size_t gi = (size_t)y0 * H + i;
        float4 wv = *(const float4 *)(W + i);
        float4 mv = *(const float4 *)(a.f.ih_mom + gi);
        float dd[4] = {o.x, o.y, o.z, o.w};
        float ww[4] = {wv.x, wv.y, wv.z, wv.w};
        float mm[4] = {mv.x, mv.y, mv.z, mv.w};
  #pragma unroll
        for (int u = 0; u < 4; u++) {
          float t = dd[u] * a.f.lr;
          ww[u] += t + mm[u] * a.f.momentum_weight;
          mm[u] = (mm[u] + t) * a.f.momentum;
        }
        *(float4 *)(a.f.ih_w + gi) = make_float4(ww[0], ww[1], ww[2], ww[3]);
        *(float4 *)(a.f.ih_mom + gi) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      }
    }
  }
  if ((walk || a.f.enabled) && rank == 0 && threadIdx.x == 0) {
    v.sc[s] = sc;
    if (a.f.enabled)
      *a.f.sc_out = sc;
  }
  /* nobody leaves while a neighbour might still push into its shared memory */
  cluster.sync();
}

static int
walk_in_registers(const RbView *v)
{
  return (v->d.i_size + WALK_CTAS - 1) / WALK_CTAS <= 8 * WALK_RW && v->d.h_size <= 32 * WALK_NC;
}

static void
walk_launch(const WalkArgs &a, const RbView *v, size_t smem, const char *what)
{
  static int attr_done = 0;
  if (!attr_done) {
    cudaFuncSetAttribute(k_walk_single<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
        200 * 1024);
    cudaFuncSetAttribute(k_walk_single<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
        200 * 1024);
    attr_done = 1;
  }
  rb_prof_begin(RB_PROF_CHAIN);
  const int n_streams = a.plane_stride ? v->n : 1;
  if (walk_in_registers(v))
    k_walk_single<true><<<WALK_CTAS * n_streams, 256, smem, rb_stream>>>(a);
  else
    k_walk_single<false><<<WALK_CTAS * n_streams, 256, smem, rb_stream>>>(a);
  LAUNCH_CHECK(what);
  rb_prof_end(RB_PROF_CHAIN);
}

static size_t
walk_smem_bytes(const RbView *v, int *rows_per)
{
  int R = (v->d.i_size + WALK_CTAS - 1) / WALK_CTAS;
  *rows_per = R;
  size_t matrices = walk_in_registers(v) ? 0 : (size_t)2 * R * v->d.h_size;
  return (matrices + 3 * v->d.h_size + 3 * WALK_CTAS + (size_t)2 * v->depth * R + 32 + 8 +
      v->d.o_size) * sizeof(float);
}

extern "C" int
rbk_walk_single_usable(const RbView *v)
{
  int R;
  return v->n == 1 && (v->d.h_size % 4) == 0 && walk_smem_bytes(v, &R) <= 200 * 1024 &&
      !getenv("RECUR_B200_NO_WALK");
}

/* small batches (below the tensor engine's 64 streams): one cluster per stream */
#define WALK_MAX_STREAMS 64

static int
walk_streams_usable(const RbView *v)
{
  int R;
  return v->n >= 1 && v->n <= WALK_MAX_STREAMS && (v->d.h_size % 4) == 0 &&
      walk_smem_bytes(v, &R) <= 200 * 1024 && !getenv("RECUR_B200_NO_WALK");
}

/* delta (+)= plane_0 + plane_1 + ... in stream order (the order of the
   reference's fold, recur-nn.c:734-748) */
__global__ void __launch_bounds__(256)
k_sum_stream_planes(float *__restrict__ delta, const float *__restrict__ planes, int size,
    int n, int accumulate)
{
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < size;
       i += gridDim.x * blockDim.x * 4) {
    float4 t = accumulate ? *(const float4 *)(delta + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < n; j++) {
      float4 p = *(const float4 *)(planes + (size_t)j * size + i);
      t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
    }
    *(float4 *)(delta + i) = t;
  }
}

static float *walk_planes = NULL;
static size_t walk_planes_cap = 0;

extern "C" void
rbk_walk_single(const RbView *v, float *ih_delta, int accumulate)
{
  WalkArgs a;
  memset(&a, 0, sizeof(a));
  a.v = *v;
  a.delta = ih_delta;
  a.accumulate = accumulate;
  size_t smem = walk_smem_bytes(v, &a.rows_per);
  if (v->n > 1) {
    size_t size = (size_t)v->d.i_size * v->d.h_size;
    if (walk_planes_cap < size * v->n) {
      if (walk_planes) {
        cudaStreamSynchronize(rb_stream);
        cudaFree(walk_planes);
      }
      walk_planes_cap = size * v->n;
      if (cudaMalloc((void **)&walk_planes, walk_planes_cap * sizeof(float)) != cudaSuccess)
        rb_die("recur-b200: out of device memory for %d gradient planes", v->n);
    }
    a.delta = walk_planes;
    a.accumulate = 0;
    a.plane_stride = size;
    walk_launch(a, v, smem, "k_walk_single<streams>");
    rb_prof_begin(RB_PROF_DW);
    k_sum_stream_planes<<<grid1d((int)(size / 4), 256), 256, 0, rb_stream>>>(ih_delta,
        walk_planes, (int)size, v->n, accumulate);
    LAUNCH_CHECK("k_sum_stream_planes");
    rb_prof_end(RB_PROF_DW);
    return;
  }
  walk_launch(a, v, smem, "k_walk_single");
}

/* rnn_bptt_calculate with batch size 1 and no bottom layer, in one launch */
extern "C" void
rbk_calculate_single(const RbView *v, const float *o_error_host, float lr, float mef,
    int adaptive, float *ho_w, float *ho_mom, float *ih_w, float *ih_mom, float *ih_delta,
    float momentum, float momentum_weight, RbScalars *sc_host)
{
  WalkArgs a;
  memset(&a, 0, sizeof(a));
  a.v = *v;
  a.delta = ih_delta;
  a.accumulate = 0;
  a.f.enabled = 1;
  a.f.o_error_in = o_error_host;
  a.f.lr = lr;
  a.f.mef = mef;
  a.f.adaptive = adaptive;
  a.f.ho_w = ho_w;
  a.f.ho_mom = ho_mom;
  a.f.ih_w = ih_w;
  a.f.ih_mom = ih_mom;
  a.f.momentum = momentum;
  a.f.momentum_weight = momentum_weight;
  a.f.sc_out = sc_host;
  size_t smem = walk_smem_bytes(v, &a.rows_per);
  walk_launch(a, v, smem, "k_walk_single<calculate>");
}

/* ------------------------------------------------------------------------ */
/* a2..a5 for ONE stream through the per-net API: rnn_opinion as a single
 * launch.  The per-net protocol keeps the caller's vectors in pinned host
 * mirrors; done with copies and the batch kernels that
 * is eight dependent stream operations for 50 k multiply-adds.  Here the
 * same cluster of eight CTAs reads hidden(t-1) and the inputs straight from
 * the mirrors, builds the ring row (soft clip included), splits the rows of
 * Wih between the CTAs (rows whose x is zero are skipped, the test is uniform
 * over the block), sums the partial hidden vectors through distributed shared
 * memory in rank order, splits the rows of Who the same way, and writes the
 * ring row, hidden and output vectors both to the device pool and back into
 * the mirrors.                                                              */

struct OpinionArgs {
  RbView v;
  const float *hidden_in;  /* pinned mirrors, read and written by the kernel */
  const float *inputs_in;
  float *input_layer_out, *hidden_out, *output_out;
  int rows_per;            /* rows of Wih per CTA */
  int hrows_per;           /* rows of Who per CTA */
};

__global__ void __cluster_dims__(WALK_CTAS, 1, 1) __launch_bounds__(256, 1)
k_opinion_single(OpinionArgs a)
{
  extern __shared__ __align__(16) float osh[];
  __shared__ float scratch[33];
  cg::cluster_group cluster = cg::this_cluster();
  const RbView &v = a.v;
  const int s = v.slots[0];
  const int I = v.d.i_size, H = v.d.h_size, O = v.d.o_size, hs1 = v.d.hidden_size + 1;
  const int rank = (int)cluster.block_rank();
  float *x = osh;                 /* [I] the ring row */
  float *P = x + I;               /* [WALK_CTAS][H] partial hidden sums of every CTA */
  float *h = P + WALK_CTAS * H;   /* [H] */
  float *PY = h + H;              /* [WALK_CTAS][O] partial outputs (used in rank 0) */

  /* the row [1 | hidden(t-1) | inputs | 0..] and its emergency soft clip */
  float sum = 0.0f;
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    float val = 0.0f;
    if (i == 0)
      val = 1.0f;
    else if (i < hs1)
      val = a.hidden_in[i];
    else if (i < hs1 + v.d.input_size)
      val = a.inputs_in[i - hs1];
    x[i] = val;
    sum += val;
  }
  sum = block_sum(sum, scratch);
  const float softclip = I * INPUT_MEAN_SOFT_TOP;
  if (sum > softclip) {
    const float scale = soft_clip_dev(sum, softclip);
    for (int i = threadIdx.x; i < I; i += blockDim.x)
      x[i] *= scale;
  }
  __syncthreads();
  if (rank == 0) {
    float *xr = x_row(v, s, 0);
    for (int i = threadIdx.x; i < I; i += blockDim.x) {
      xr[i] = x[i];
      a.input_layer_out[i] = x[i];
    }
  }

  /* every CTA of the cluster is running before anyone writes into a
     neighbour's shared memory */
  cluster.sync();

  /* partial hidden sums over this CTA's rows of Wih */
  const int y0 = min(I, rank * a.rows_per), y1 = min(I, y0 + a.rows_per);
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float acc = 0.0f;
    for (int yb = y0; yb < y1; yb += 8) {
      float w[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int y = yb + u;
        w[u] = (y < y1 && x[y] != 0.0f) ? __ldg(v.Wih + (size_t)y * H + c) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int y = yb + u;
        if (y < y1)
          acc = fmaf(x[y], w[u], acc);
      }
    }
#pragma unroll
    for (int q = 0; q < WALK_CTAS; q++)
      cluster.map_shared_rank(P, q)[rank * H + c] = acc;
  }
  cluster.sync();

  /* every CTA finishes the hidden vector for itself (same order, same bits) */
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float t = 0.0f;
#pragma unroll
    for (int q = 0; q < WALK_CTAS; q++)
      t += P[q * H + c];
    if (v.activation == RNN_RESQRT) {
      t = (t > 0.0f) ? sqrtf(t + 1.0f) - 1.0f : 0.0f;
    }
    else if (v.activation == RNN_RECLIP20) {
      if (c >= 1) {
        t = t < 20.0f ? t : 20.0f;
        t = (t > 0.0f) ? t : 0.0f;
      }
    }
    else if (c >= 1) {
      t = (t > 0.0f) ? t : 0.0f;
    }
    if (c == 0)
      t = 1.0f;
    h[c] = t;
    if (rank == 0) {
      v.Hd[(size_t)s * H + c] = t;
      a.hidden_out[c] = t;
    }
  }
  __syncthreads();

  /* partial outputs over this CTA's rows of Who */
  const int c0 = min(H, rank * a.hrows_per), c1 = min(H, c0 + a.hrows_per);
  float *py0 = cluster.map_shared_rank(PY, 0) + (size_t)rank * O;
  for (int o = threadIdx.x; o < O; o += blockDim.x) {
    float acc = 0.0f;
    for (int cb = c0; cb < c1; cb += 8) {
      float w[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int c = cb + u;
        w[u] = (c < c1 && h[c] != 0.0f) ? __ldg(v.Who + (size_t)c * O + o) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int c = cb + u;
        if (c < c1)
          acc = fmaf(h[c], w[u], acc);
      }
    }
    py0[o] = acc;
  }
  cluster.sync();
  if (rank == 0) {
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
      float t = 0.0f;
#pragma unroll
      for (int q = 0; q < WALK_CTAS; q++)
        t += PY[(size_t)q * O + o];
      v.Y[(size_t)s * O + o] = t;
      a.output_out[o] = t;
    }
  }
}

static size_t
opinion_smem_bytes(const RbView *v)
{
  return ((size_t)v->d.i_size + (size_t)(WALK_CTAS + 1) * v->d.h_size +
      (size_t)WALK_CTAS * v->d.o_size + 16) * sizeof(float);
}

extern "C" int
rbk_opinion_single_usable(const RbView *v)
{
  return v->n == 1 && opinion_smem_bytes(v) <= 200 * 1024 &&
      (size_t)v->d.i_size * v->d.h_size <= 512 * 1024 && !getenv("RECUR_B200_NO_WALK");
}

extern "C" void
rbk_opinion_single(const RbView *v, const float *hidden_in, const float *inputs_in,
    float *input_layer_out, float *hidden_out, float *output_out)
{
  static int attr_done = 0;
  if (!attr_done) {
    cudaFuncSetAttribute(k_opinion_single, cudaFuncAttributeMaxDynamicSharedMemorySize,
        200 * 1024);
    attr_done = 1;
  }
  OpinionArgs a;
  a.v = *v;
  a.hidden_in = hidden_in;
  a.inputs_in = inputs_in;
  a.input_layer_out = input_layer_out;
  a.hidden_out = hidden_out;
  a.output_out = output_out;
  a.rows_per = (v->d.i_size + WALK_CTAS - 1) / WALK_CTAS;
  a.hrows_per = (v->d.h_size + WALK_CTAS - 1) / WALK_CTAS;
  rb_prof_begin(RB_PROF_FWD);
  k_opinion_single<<<WALK_CTAS, 256, opinion_smem_bytes(v), rb_stream>>>(a);
  LAUNCH_CHECK("k_opinion_single");
  rb_prof_end(RB_PROF_FWD);
}

/* ------------------------------------------------------------------------ */
/* a10 for a batch of SMALL nets: the whole of Wih resident in one SM's shared
 * memory (up to ~50 k weights: H199 nets are 46 k), one CTA per stream, no
 * synchronisation between CTAs at all.
 *
 * For such nets a BPTT step of the whole batch is a few MFLOP: the persistent
 * tensor kernel spends its 8-11 us per step on barriers and phases, not on
 * MMAs, and only a handful of its CTAs have work.  Streams are independent
 * given the weights, so here each CTA walks its stream alone at the speed of
 * shared memory: per step a warp takes four rows at a time, skips the rows
 * whose x_k is zero (recur-nn.c:347), reduces the four dot products in six
 * shuffles, and the block sums the squares for the stop rule
 * (recur-nn.c:383-413).  E(k+1) goes to the pool (and, for the tensor
 * engine's weight-gradient kernel, as hi/lo planes); the gradient itself is
 * left to that kernel.                                                      */

struct ResidentArgs {
  RbView v;
  RbPlanes E; /* optional planes of the error chain (hi == NULL: none) */
};

#define RES_THREADS 512
#define RES_WARPS (RES_THREADS / 32)

__global__ void __launch_bounds__(RES_THREADS, 1)
k_walk_resident(ResidentArgs a)
{
  extern __shared__ __align__(16) float rsh[];
  const RbView &v = a.v;
  const int s = slot_of(v, blockIdx.x);
  const int I = v.d.i_size, H = v.d.h_size, hs1 = v.d.hidden_size + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *W = rsh;                      /* [I][H] */
  float *E = W + (size_t)I * H;        /* [2][H] */
  float *xs = E + 2 * H;               /* [2][I] ring rows, this step's and the next */
  float *red = xs + 2 * I;             /* [4 * RES_WARPS] */
  int *rows = (int *)(red + 4 * RES_WARPS); /* [I] indices of the rows with a nonzero x_k */
  int *cnt = rows + I;                 /* [32 + 1] active rows per 32-row segment */

  RbScalars sc = v.sc[s];
  if (!sc.live || (sc.adaptive & 2))
    return; /* uniform over the block */
  const int pos = v.pos[s];
  const int depth = v.depth;
  const float e_scale = a.E.scale_dev ? *a.E.scale_dev : a.E.scale;
  for (int i = threadIdx.x * 4; i < I * H; i += RES_THREADS * 4)
    *(float4 *)(W + i) = __ldg((const float4 *)(v.Wih + i));
  {
    const float *e0 = e_row(v, s, 0);
    for (int i = threadIdx.x; i < H; i += RES_THREADS)
      E[i] = e0[i];
    const float *x0 = x_row(v, s, 0);
    for (int i = threadIdx.x; i < I; i += RES_THREADS)
      xs[i] = x0[i];
  }
  __syncthreads();
  const int n_seg = (I + 31) >> 5; /* <= 32 */

  for (int k = 0; k < depth; k++) {
    const float *cur = E + (k & 1) * H;
    float *nxt = E + ((k + 1) & 1) * H;
    const float *x_now = xs + (k & 1) * I;
    /* the next step's ring row travels while this one computes */
    float x_next[2] = {0.f, 0.f};
    if (k + 1 < depth) {
      int p = pos - (k + 1);
      if (p < 0)
        p += depth;
      const float *xr = v.X + ((size_t)p * v.cap + s) * I;
#pragma unroll
      for (int u = 0; u < 2; u++) {
        int i = threadIdx.x + u * RES_THREADS;
        if (i < I)
          x_next[u] = xr[i];
      }
    }
    float *e_next_row = e_row(v, s, k + 1);
    const size_t plane_off = ((size_t)(k + 1) * v.cap + s) * a.E.pitch;

    /* the rows that were multiplied by zero have no error and cost nothing
       (recur-nn.c:347): list the others, 32-row segments by ballot */
    unsigned int seg_mask[2] = {0u, 0u};
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int seg = warp + u * RES_WARPS;
      if (seg < n_seg) {
        const int y = seg * 32 + lane;
        const float x = (y < I) ? x_now[y] : 0.0f;
        const bool act = x != 0.0f && (v.activation != RNN_RECLIP20 || x < 20.0f);
        seg_mask[u] = __ballot_sync(0xffffffffu, act);
        if (lane == 0)
          cnt[seg] = __popc(seg_mask[u]);
        if (!act && y < I) {
          e_next_row[y] = 0.0f;
          if (y < H)
            nxt[y] = 0.0f;
          if (a.E.hi) {
            a.E.hi[plane_off + y] = 0;
            a.E.lo[plane_off + y] = 0;
          }
        }
      }
    }
    __syncthreads();
    int n_act = 0;
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int seg = warp + u * RES_WARPS;
      if (seg < n_seg) {
        int off = 0;
        for (int q = 0; q < seg; q++)
          off += cnt[q];
        if (seg_mask[u] & (1u << lane))
          rows[off + __popc(seg_mask[u] & ((1u << lane) - 1u))] = seg * 32 + lane;
      }
    }
    for (int q = 0; q < n_seg; q++)
      n_act += cnt[q];
    __syncthreads();

    float sq4 = 0.0f;
    for (int ib = warp * 4; ib < n_act; ib += RES_WARPS * 4) {
      int yr[4];
      float dot[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        yr[u] = (ib + u < n_act) ? rows[ib + u] : -1;
        dot[u] = 0.0f;
      }
      for (int c = lane; c < H; c += 32) {
        const float ec = cur[c];
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (yr[u] >= 0) /* uniform over the warp */
            dot[u] = fmaf(W[(size_t)yr[u] * H + c], ec, dot[u]);
      }
      float mine;
      {
        const bool hi = lane & 16;
        float k0 = hi ? dot[2] : dot[0], k1 = hi ? dot[3] : dot[1];
        float s0 = hi ? dot[0] : dot[2], s1 = hi ? dot[1] : dot[3];
        k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
        k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        const bool mid = lane & 8;
        mine = mid ? k1 : k0;
        const float send = mid ? k0 : k1;
        mine += __shfl_xor_sync(0xffffffffu, send, 8);
        mine += __shfl_xor_sync(0xffffffffu, mine, 4);
        mine += __shfl_xor_sync(0xffffffffu, mine, 2);
        mine += __shfl_xor_sync(0xffffffffu, mine, 1);
      }
      /* lanes 8g .. 8g+7 hold row g's sum; the first of them finishes the row */
      if ((lane & 7) == 0) {
        const int g = lane >> 3;
        const int y = (g == 0) ? yr[0] : (g == 1) ? yr[1] : (g == 2) ? yr[2] : yr[3];
        if (y >= 0) {
          float e = mine;
          if (v.activation == RNN_RESQRT)
            e /= 2.0f * (x_now[y] + 1.0f);
          sq4 = fmaf(e, e, sq4);
          if (v.CIE && y >= hs1 && y < hs1 + v.d.input_size)
            v.CIE[(size_t)s * v.bl_o + y - hs1] += e;
          const float e_store = (y == 0 || (y >= hs1 && y < H)) ? 0.0f : e;
          e_next_row[y] = e_store;
          if (y < H)
            nxt[y] = e_store;
          if (a.E.hi) {
            rb_h16 hiv, lov;
            rb_split_f16(e_store * e_scale, hiv, lov);
            a.E.hi[plane_off + y] = hiv;
            a.E.lo[plane_off + y] = lov;
          }
        }
      }
    }
    {
      /* the warp's four row groups first (fixed order), then one value per warp */
      float w = ((lane & 7) == 0) ? sq4 : 0.0f;
      w += __shfl_xor_sync(0xffffffffu, w, 8);
      w += __shfl_xor_sync(0xffffffffu, w, 16);
      if (lane == 0)
        red[warp] = w;
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      int i = threadIdx.x + u * RES_THREADS;
      if (i < I)
        xs[((k + 1) & 1) * I + i] = x_next[u];
    }
    __syncthreads();
    float es = 0.0f;
#pragma unroll
    for (int q = 0; q < RES_WARPS; q++)
      es += red[q];
    __syncthreads(); /* red, rows and cnt are rewritten next step */
    /* recur-nn.c:383-413, the same in every thread */
    sc.err_sum = es;
    sc.cum_error += sqrtf(es);
    sc.n_steps = k + 1;
    const int t_loop = depth - k;
    const bool stop = (es <= sc.min_sum || es > sc.max_sum);
    const bool last = (k == depth - 1);
    if (stop || last) {
      sc.live = 0;
      const int t_left = stop ? t_loop : 0;
      sc.t_left = t_left;
      const float ceiling = ERROR_GAIN_CEILING * sc.top_scaled;
      if (es > ceiling) {
        sc.ih_scale = soft_clip_dev(es, sc.max_sum);
      }
      else {
        sc.ih_scale = 1.0f;
        if (sc.adaptive & 1) {
          int depth_error = depth / 4 - t_left;
          float min_gain = MIN_ERROR_GAIN * sc.top_scaled;
          float mef = sc.mef;
          if (mef < MAX_MIN_ERROR_FACTOR && (min_gain != sc.min_sum || depth_error < 0))
            mef = (float)((double)mef * (1.0 + depth_error * 1e-3));
          sc.mef = fmaxf(mef, ABS_MIN_ERROR_FACTOR);
        }
      }
      break;
    }
  }
  if (threadIdx.x == 0)
    v.sc[s] = sc;
}

static size_t
resident_smem_bytes(const RbView *v)
{
  return ((size_t)v->d.i_size * v->d.h_size + 2 * v->d.h_size + 2 * v->d.i_size +
      4 * RES_WARPS + v->d.i_size + 40) * sizeof(float);
}

extern "C" int
rbk_walk_resident_usable(const RbView *v)
{
  return (v->d.h_size % 4) == 0 && v->d.i_size <= 1024 &&
      resident_smem_bytes(v) <= 224 * 1024 && !getenv("RECUR_B200_NO_RESIDENT");
}

/* the walk of every stream of the batch; E(1..) to the pool (and planes) */
extern "C" void
rbk_walk_resident(const RbView *v, const RbPlanes *E)
{
  static int attr_done = 0;
  if (!attr_done) {
    cudaFuncSetAttribute(k_walk_resident, cudaFuncAttributeMaxDynamicSharedMemorySize,
        224 * 1024);
    attr_done = 1;
  }
  ResidentArgs a;
  a.v = *v;
  if (E)
    a.E = *E;
  else
    memset(&a.E, 0, sizeof(a.E));
  rb_prof_begin(RB_PROF_CHAIN);
  k_walk_resident<<<v->n, RES_THREADS, resident_smem_bytes(v), rb_stream>>>(a);
  LAUNCH_CHECK("k_walk_resident");
  rb_prof_end(RB_PROF_CHAIN);
}

/* ------------------------------------------------------------------------ */
/* f4: the feature front end of gstrnnca (gstrnnca.c:644-691, 805-830), so
   that a frame of the cellular automaton never leaves the device between
   its planes coming in and going out.
   k_rnnca_gather: cell (cx, cy) reads len_Y luma and len_C chroma neighbours
   of the previous frame at the given offsets (clamped at the edges or wrapped
   once, get_offset_point), scaled to the unit interval, then its position
   (and optionally the radial term) - straight into the input part of its
   ring row.  k_rnnca_emit: fast_sigmoid of the three outputs, UNIT_TO_BYTE. */

struct RnncaArgs {
  RbView v;
  const u8 *frame;      /* [3][height][width]: Y, Cb, Cr */
  u8 *frame_out;
  int width, height;
  const int *off_y, *off_c; /* (dx, dy) pairs */
  int len_y, len_c, len_pos, edges;
};

__global__ void __launch_bounds__(128)
k_rnnca_gather(RnncaArgs a)
{
  const RbView &v = a.v;
  const int cell = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (cell >= v.n)
    return;
  const int s = slot_of(v, cell);
  const int cx = cell % a.width, cy = cell / a.width;
  const int plane = a.width * a.height;
  float *in = x_row(v, s, 0) + v.d.hidden_size + 1;
  const float unit = 1.0f / 255.0f;
  for (int j = lane; j < a.len_y; j += 32)
    in[j] = a.frame[rnnca_offset_point(a.off_y + 2 * j, cx, cy, a.width, a.height, a.edges)] * unit;
  for (int j = lane; j < a.len_c; j += 32) {
    int o = rnnca_offset_point(a.off_c + 2 * j, cx, cy, a.width, a.height, a.edges);
    in[a.len_y + 2 * j] = a.frame[plane + o] * unit;
    in[a.len_y + 2 * j + 1] = a.frame[2 * plane + o] * unit;
  }
  if (lane == 0) {
    const int i = a.len_y + 2 * a.len_c;
    const float xx = cx * 1.0f / a.width, yy = cy * 1.0f / a.height;
    in[i] = xx;
    in[i + 1] = yy;
    if (a.len_pos == 3)
      in[i + 2] = (float)(0.5 - (double)((yy - 0.5f) * (yy - 0.5f) + (xx - 0.5f) * (xx - 0.5f)));
  }
}

__global__ void
k_rnnca_emit(RnncaArgs a)
{
  const RbView &v = a.v;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= v.n)
    return;
  const int s = slot_of(v, cell);
  const int plane = a.width * a.height;
  const float *y = v.Y + (size_t)s * v.d.o_size;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float sg = 1.0f / (1.0f + fast_expf_dev(-y[i] * 1.0f)); /* badmaths.h:31-44, SIGMOID_SCALE 1 */
    a.frame_out[i * plane + cell] = (u8)(sg * 255.9f);       /* UNIT_TO_BYTE */
  }
}

/* The whole frame step for TINY nets (rnnca's cells are 35 / 51 / 3): the
   batch kernels give every cell a thread block for its input row and tile the
   forward pass as a GEMM - at two million cells of a 52-unit net that is
   eleven milliseconds of mostly empty blocks.  Here the weights (18 KB) sit in
   shared memory and ONE WARP does a cell from the frame bytes to the frame
   bytes: gather (gstrnnca.c:670-691), the row [1 | hidden(t-1) | inputs] with
   its soft clip, hidden = act(x . Wih) two units per lane, output = hidden .
   Who by shuffle, fast_sigmoid, UNIT_TO_BYTE; hidden state, ring row and
   outputs are also written to the pool, so every other call still sees the
   state it expects.  Warps walk the cells in a grid-stride loop. */
__global__ void __launch_bounds__(256)
k_rnnca_cells(RnncaArgs a)
{
  extern __shared__ __align__(16) float csh[];
  const RbView &v = a.v;
  const int I = v.d.i_size, H = v.d.h_size, O = v.d.o_size, hs1 = v.d.hidden_size + 1;
  float *W = csh;               /* [I][H] */
  float *Wo = W + I * H;        /* [H][O] */
  float *xw = Wo + H * O;       /* [8 warps][I] */
  for (int i = threadIdx.x; i < I * H; i += blockDim.x)
    W[i] = v.Wih[i];
  for (int i = threadIdx.x; i < H * O; i += blockDim.x)
    Wo[i] = v.Who[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *x = xw + warp * I;
  const int plane = a.width * a.height;
  const float unit = 1.0f / 255.0f;
  const int n_in = a.len_y + 2 * a.len_c;
  for (int cell = blockIdx.x * 8 + warp; cell < v.n; cell += gridDim.x * 8) {
    const int s = v.base + cell;
    const int cx = cell % a.width, cy = cell / a.width;
    const float *hprev = v.Hd + (size_t)s * H;
    float sum = 0.0f;
    for (int i = lane; i < I; i += 32) {
      float val = 0.0f;
      if (i == 0)
        val = 1.0f;
      else if (i < hs1)
        val = hprev[i];
      else if (i < hs1 + a.len_y)
        val = a.frame[rnnca_offset_point(a.off_y + 2 * (i - hs1), cx, cy, a.width, a.height,
                a.edges)] * unit;
      else if (i < hs1 + n_in) {
        const int j = (i - hs1 - a.len_y) >> 1, which = (i - hs1 - a.len_y) & 1;
        const int o = rnnca_offset_point(a.off_c + 2 * j, cx, cy, a.width, a.height, a.edges);
        val = a.frame[(1 + which) * plane + o] * unit;
      }
      else if (i < hs1 + n_in + a.len_pos) {
        const int k = i - hs1 - n_in;
        const float xx = cx * 1.0f / a.width, yy = cy * 1.0f / a.height;
        val = (k == 0) ? xx : (k == 1) ? yy
            : (float)(0.5 - (double)((yy - 0.5f) * (yy - 0.5f) + (xx - 0.5f) * (xx - 0.5f)));
      }
      x[i] = val;
      sum += val;
    }
    sum = warp_sum(sum);
    const float softclip = I * INPUT_MEAN_SOFT_TOP;
    const float scale = (sum > softclip) ? soft_clip_dev(sum, softclip) : 1.0f;
    __syncwarp();
    float *xr = x_row(v, s, 0);
    for (int i = lane; i < I; i += 32) {
      float val = x[i] * scale;
      if (scale != 1.0f)
        x[i] = val;
      xr[i] = val;
    }
    __syncwarp();
    /* hidden units lane and lane + 32 */
    float h0 = 0.0f, h1 = 0.0f;
    const int c0 = lane, c1 = lane + 32;
    for (int y = 0; y < I; y++) {
      const float xv = x[y];
      if (xv != 0.0f) { /* uniform over the warp: the reference's row skip */
        h0 = fmaf(xv, W[y * H + c0], h0);
        if (c1 < H)
          h1 = fmaf(xv, W[y * H + c1], h1);
      }
    }
    float hv[2] = {h0, h1};
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int c = lane + 32 * u;
      float t = hv[u];
      if (v.activation == RNN_RESQRT) {
        t = (t > 0.0f) ? sqrtf(t + 1.0f) - 1.0f : 0.0f;
      }
      else if (v.activation == RNN_RECLIP20) {
        if (c >= 1) {
          t = t < 20.0f ? t : 20.0f;
          t = (t > 0.0f) ? t : 0.0f;
        }
      }
      else if (c >= 1) {
        t = (t > 0.0f) ? t : 0.0f;
      }
      if (c == 0)
        t = 1.0f;
      if (c >= H)
        t = 0.0f;
      hv[u] = t;
      if (c < H)
        v.Hd[(size_t)s * H + c] = t;
    }
    /* outputs by shuffle; lane o keeps output o */
    float mine = 0.0f;
    for (int o = 0; o < O; o++) {
      float p = hv[0] * Wo[c0 * O + o] + ((c1 < H) ? hv[1] * Wo[c1 * O + o] : 0.0f);
      p = warp_sum(p);
      if (lane == o)
        mine = p;
    }
    if (lane < O)
      v.Y[(size_t)s * O + lane] = mine;
    if (lane < 3) {
      float sg = 1.0f / (1.0f + fast_expf_dev(-mine * 1.0f));
      a.frame_out[lane * plane + cell] = (u8)(sg * 255.9f);
    }
    __syncwarp();
  }
}

extern "C" int
rbk_rnnca_cells_usable(const RbView *v)
{
  return v->contiguous && v->d.h_size <= 64 && v->d.i_size <= 256 && v->d.o_size <= 32 &&
      v->d.output_size >= 3 && !getenv("RECUR_B200_NO_CELLS");
}

extern "C" void
rbk_rnnca_cells(const RbView *v, const u8 *frame_dev, u8 *frame_out_dev, int width, int height,
    const int *off_y_dev, int len_y, const int *off_c_dev, int len_c, int len_pos, int edges)
{
  RnncaArgs a;
  memset(&a, 0, sizeof(a));
  a.v = *v;
  a.frame = frame_dev;
  a.frame_out = frame_out_dev;
  a.width = width;
  a.height = height;
  a.off_y = off_y_dev;
  a.off_c = off_c_dev;
  a.len_y = len_y;
  a.len_c = len_c;
  a.len_pos = len_pos;
  a.edges = edges;
  size_t smem = ((size_t)v->d.i_size * v->d.h_size + (size_t)v->d.h_size * v->d.o_size +
      8 * (size_t)v->d.i_size) * sizeof(float);
  static int attr_done = 0;
  if (!attr_done) {
    cudaFuncSetAttribute(k_rnnca_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_done = 1;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int blocks = sms * 8;
  if (blocks > cdiv(v->n, 8))
    blocks = cdiv(v->n, 8);
  rb_prof_begin(RB_PROF_FWD);
  k_rnnca_cells<<<blocks, 256, smem, rb_stream>>>(a);
  LAUNCH_CHECK("k_rnnca_cells");
  rb_prof_end(RB_PROF_FWD);
}

extern "C" void
rbk_rnnca_gather(const RbView *v, const u8 *frame_dev, int width, int height,
    const int *off_y_dev, int len_y, const int *off_c_dev, int len_c, int len_pos, int edges)
{
  RnncaArgs a;
  memset(&a, 0, sizeof(a));
  a.v = *v;
  a.frame = frame_dev;
  a.width = width;
  a.height = height;
  a.off_y = off_y_dev;
  a.off_c = off_c_dev;
  a.len_y = len_y;
  a.len_c = len_c;
  a.len_pos = len_pos;
  a.edges = edges;
  rb_prof_begin(RB_PROF_SMALL);
  k_rnnca_gather<<<cdiv(v->n, 4), 128, 0, rb_stream>>>(a);
  LAUNCH_CHECK("k_rnnca_gather");
  rb_prof_end(RB_PROF_SMALL);
}

extern "C" void
rbk_rnnca_emit(const RbView *v, u8 *frame_out_dev, int width, int height)
{
  RnncaArgs a;
  memset(&a, 0, sizeof(a));
  a.v = *v;
  a.frame_out = frame_out_dev;
  a.width = width;
  a.height = height;
  rb_prof_begin(RB_PROF_SMALL);
  k_rnnca_emit<<<cdiv(v->n, 256), 256, 0, rb_stream>>>(a);
  LAUNCH_CHECK("k_rnnca_emit");
  rb_prof_end(RB_PROF_SMALL);
}

extern "C" void
rbk_bptt(const RbView *v, float *ih_delta, int accumulate)
{
  if (walk_streams_usable(v)) {
    rb_note_walk_kernel("k_walk_single");
    rbk_walk_single(v, ih_delta, accumulate);
    return;
  }
  rb_note_walk_kernel("k_gemm<CHAIN>");
  GemmArgs g;
  g.v = *v;
  g.delta = ih_delta;
  g.accumulate = accumulate;
  g.use_noise = 0;
  dim3 cgrid(cdiv(v->d.i_size, TN), cdiv(v->n, TM));
  for (int k = 0; k < v->depth; k++) {
    g.k = k;
    rb_prof_begin(RB_PROF_CHAIN);
    k_gemm<G_CHAIN><<<cgrid, 256, 0, rb_stream>>>(g);
    LAUNCH_CHECK("k_gemm<CHAIN>");
    rb_prof_end(RB_PROF_CHAIN);
    rbk_chain_decide(v, k);
  }
  dim3 dgrid(cdiv(v->d.h_size, TN), cdiv(v->d.i_size, TM));
  rb_prof_begin(RB_PROF_DW);
  k_gemm<G_DW><<<dgrid, 256, 0, rb_stream>>>(g);
  LAUNCH_CHECK("k_gemm<DW>");
  rb_prof_end(RB_PROF_DW);
}

/* the weight gradient alone on the FMA engine (the tensor engine's fallback
   for shapes its pair kernel has no grid for) */
extern "C" void
rbk_dw_fma(const RbView *v, float *ih_delta, int accumulate)
{
  GemmArgs g;
  g.v = *v;
  g.k = 0;
  g.delta = ih_delta;
  g.accumulate = accumulate;
  g.use_noise = 0;
  dim3 dgrid(cdiv(v->d.h_size, TN), cdiv(v->d.i_size, TM));
  k_gemm<G_DW><<<dgrid, 256, 0, rb_stream>>>(g);
  LAUNCH_CHECK("k_gemm<DW>");
}

extern "C" void
rbk_set_params(const RbView *v, const float *lr_dev, const float *mef_dev, int adaptive)
{
  rb_prof_begin(RB_PROF_SMALL);
  k_set_params<<<cdiv(v->n, 128), 128, 0, rb_stream>>>(*v, lr_dev, mef_dev, adaptive);
  LAUNCH_CHECK("k_set_params");
  rb_prof_end(RB_PROF_SMALL);
}

extern "C" void
rbk_mask_streams(const RbView *v, const u8 *active_dev)
{
  k_mask_streams<<<v->n, 64, 0, rb_stream>>>(*v, active_dev);
  LAUNCH_CHECK("k_mask_streams");
}

extern "C" void
rbk_set_params_scalar(const RbView *v, float lr, float mef, int adaptive)
{
  k_set_params_scalar<<<1, 1, 0, rb_stream>>>(*v, lr, mef, adaptive);
  LAUNCH_CHECK("k_set_params_scalar");
}

extern "C" void
rbk_apply_learning(int method, float *weights, const float *delta,
    float *momentums, float *aux, int size, float rate, float momentum,
    float momentum_weight, const float *rate_scale_dev)
{
  rb_prof_begin(RB_PROF_UPDATE);
  k_apply_learning<<<grid1d(size, 256), 256, 0, rb_stream>>>(method, weights, delta,
      momentums, aux, size, rate, momentum, momentum_weight, rate_scale_dev);
  LAUNCH_CHECK("k_apply_learning");
  rb_prof_end(RB_PROF_UPDATE);
}

extern "C" void
rbk_scale(float *a, int n, float s)
{
  k_scale<<<grid1d(n, 256), 256, 0, rb_stream>>>(a, n, s);
  LAUNCH_CHECK("k_scale");
}

extern "C" void
rbk_zero_small(float *a, int n)
{
  k_zero_small<<<grid1d(n, 256), 256, 0, rb_stream>>>(a, n);
  LAUNCH_CHECK("k_zero_small");
}

extern "C" void
rbk_clamp(float *a, int n, float lo, float hi)
{
  k_clamp<<<grid1d(n, 256), 256, 0, rb_stream>>>(a, n, lo, hi);
  LAUNCH_CHECK("k_clamp");
}

extern "C" void
rbk_tall_poppy(float *a, int n, float threshold, float scale)
{
  k_tall_poppy<<<1, 1024, 0, rb_stream>>>(a, n, threshold, scale);
  LAUNCH_CHECK("k_tall_poppy");
}

extern "C" void
rbk_add_at(float *a, int index, float value)
{
  k_add_at<<<1, 1, 0, rb_stream>>>(a, index, value);
  LAUNCH_CHECK("k_add_at");
}

extern "C" void
rbk_axpy(float *dst, const float *src, int n, float s, const float *s_dev)
{
  k_axpy<<<grid1d(n, 256), 256, 0, rb_stream>>>(dst, src, n, s, s_dev);
  LAUNCH_CHECK("k_axpy");
}

extern "C" void
rbk_fill(float *a, size_t n, float value)
{
  k_fill<<<grid1d((long long)n, 256), 256, 0, rb_stream>>>(a, n, value);
  LAUNCH_CHECK("k_fill");
}

extern "C" void
rbk_abs_sum(const float *a, int n, float *out_dev)
{
  k_abs_sum<<<1, 1024, 0, rb_stream>>>(a, n, out_dev);
  LAUNCH_CHECK("k_abs_sum");
}
