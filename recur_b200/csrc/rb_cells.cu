/* rb_cells.cu - rnnca at its own scale (BASELINE.json configs[4]): one tiny net
 * PER PIXEL of a video frame (reference gstrnnca.c:805-830, fill_frame), two
 * million of them at 1080p, all with the trainers' weights.  As host RecurNN
 * clones they cost seconds to create and exist only to hold 52 floats of hidden
 * state each; RnnCells keeps just that state, on the device.
 *
 * A frame is one kernel.  Two of them are here:
 *
 *   k_cells_frame_tc (the one that runs).  The cells' hidden state lives in HBM
 *     in the layout the tensor cores read: per tile of 128 cells the first K
 *     chunk of a K-major tcgen05 A operand, FP16 hi and lo planes (rb_split.cuh),
 *     SWIZZLE_128B rows.  One persistent CTA per SM, four kinds of warp, two
 *     tiles in flight: a thread bulk-copies a tile's state into shared memory;
 *     gather warps read the neighbourhood's bytes and write the second K chunk
 *     (the bytes as the FP16 integers they are, 1/255 folded into their
 *     weights); a thread issues the MMAs (the weights are the B operand, fetched
 *     once per CTA) into tensor memory; drain warps read the sums back, apply
 *     the input soft clip (linear in the inputs, so it multiplies the sums), the
 *     activation, the output layer and the sigmoid, write the three bytes, and
 *     send the new state back as planes through a staging tile and a bulk
 *     copy.  The frame is bounded by HBM (state in, state out), not by its
 *     19.6 GFLOP of multiply-adds.  Several GPUs share a frame by rows and swap
 *     the rows next to their bands between frames (cells_halo).
 *
 *   k_cells_frame (RnnCells created with RECUR_B200_CELLS_FMA=1; a second,
 *     independent reading the tests compare the first with): one thread does a
 *     cell in FP32 FMAs in the reference's order, weights in constant memory,
 *     state as plain floats, column-major ([hidden unit][cell]).
 */
#include "rb_internal.h"
#include "rb_kernels.h"
#include "rb_split.cuh"
#include "rb_devmath.cuh"
#include "rb_ptx.cuh"
#include "rb_comm.h"
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned char u8;

#define LAUNCH_CHECK(name) do {                                         \
    cudaError_t e_ = cudaGetLastError();                                \
    if (e_ != cudaSuccess)                                              \
      rb_die("recur-b200: launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    rb_count_launch(1);                                                 \
  } while (0)

static inline int
cdiv(int a, int b)
{
  return (a + b - 1) / b;
}

static float
rb_half_to_float(rb_h16 h)
{
  const int sign = h >> 15, e = (h >> 10) & 31, m = h & 1023;
  float v = e == 0 ? ldexpf((float)m, -24) : e == 31 ? (m ? NAN : INFINITY)
      : ldexpf((float)(m | 1024), e - 25);
  return sign ? -v : v;
}

/* ======================================================================== */
/* the FP32 frame (cross-check)                                               */

#define CELLS_XIN 40 /* most gathered inputs + position terms per cell */

struct CellsArgs {
  const float *Wih, *Who;
  int I, H, O, hs, activation;
  float *state; /* [H][n] */
  int n;        /* cells this GPU runs */
  int cell0;    /* the first of them in the frame (a band of rows, when several GPUs share it) */
  const u8 *frame;
  u8 *frame_out;
  int width, height;
  const int *off_y, *off_c;
  int len_y, len_c, len_pos, edges;
};

/* The weights as the kernel reads them: [i_size][HN] and [h_size][4], in
   CONSTANT memory.  Every thread of a warp wants the same weight at the same
   time; from shared memory that is an LDS.128 per four FMAs whose 512 bytes of
   register write-back cost four cycles of the SM's one load path (measured:
   the kernel ran at 17 % of the FMA rate).  From the constant bank the weight
   arrives through the uniform datapath and the FMA pipe is what is left. */
#define CELLS_MAX_I 104
__constant__ float4 cells_w[CELLS_MAX_I * 16];
__constant__ float4 cells_wo[64];

__global__ void
k_cells_pack(const float *__restrict__ Wih, const float *__restrict__ Who, int I, int H, int O,
    int HN, float *w, float *wo)
{
  /* column c of the packed rows is hidden unit c + 1: unit 0 is the bias, its
     sums are never used (recur-nn.c:144), and without it 51 units fill 13
     aligned float4s */
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < I * HN) {
    int y = t / HN, c = t - y * HN + 1;
    w[t] = (c < H) ? Wih[(size_t)y * H + c] : 0.0f;
  }
  if (t < H * 4) {
    int y = t >> 2, o = t & 3;
    wo[t] = (o < O) ? Who[(size_t)y * O + o] : 0.0f;
  }
}

#define CELLS_NT 128

template <int HN>
__device__ __forceinline__ void
cells_row_fma(float (&acc)[HN], float x, const float4 *w4)
{
#pragma unroll
  for (int c = 0; c < HN / 4; c++) {
    const float4 w = w4[c];
    acc[4 * c] = fmaf(x, w.x, acc[4 * c]);
    acc[4 * c + 1] = fmaf(x, w.y, acc[4 * c + 1]);
    acc[4 * c + 2] = fmaf(x, w.z, acc[4 * c + 2]);
    acc[4 * c + 3] = fmaf(x, w.w, acc[4 * c + 3]);
  }
}

template <int HN>
__global__ void __launch_bounds__(CELLS_NT, 4)
k_cells_frame(CellsArgs a)
{
  /* each thread's input vector, [row][thread]: written in the first pass (the
     sum for the soft clip needs every input before any product), read back in
     the second */
  extern __shared__ float xs[];
  float *mine = xs + threadIdx.x;
  const float4 *W = cells_w;
  const int hs1 = a.hs + 1;
  const int n_in = a.len_y + 2 * a.len_c, n_tot = n_in + a.len_pos;
  const int plane = a.width * a.height;
  const float unit = 1.0f / 255.0f;
  for (int cell = blockIdx.x * CELLS_NT + threadIdx.x; cell < a.n; cell += gridDim.x * CELLS_NT) {
    const int gcell = a.cell0 + cell;
    const int cx = gcell % a.width, cy = gcell / a.width;
    const float *st = a.state + cell;
    float sum = 1.0f; /* input 0, the bias */
#pragma unroll 10
    for (int i = 1; i < hs1; i++) {
      const float v = st[(size_t)i * a.n];
      mine[i * CELLS_NT] = v;
      sum += v;
    }
    /* fill_net_inputs (gstrnnca.c:670-691) */
#pragma unroll 1
    for (int j = 0; j < n_tot; j++) {
      float val;
      if (j < a.len_y)
        val = a.frame[rnnca_offset_point(a.off_y + 2 * j, cx, cy, a.width, a.height, a.edges)] *
            unit;
      else if (j < n_in) {
        const int q = (j - a.len_y) >> 1, which = (j - a.len_y) & 1;
        const int o = rnnca_offset_point(a.off_c + 2 * q, cx, cy, a.width, a.height, a.edges);
        val = a.frame[(1 + which) * plane + o] * unit;
      }
      else {
        const int k = j - n_in;
        const float xx = cx * 1.0f / a.width, yy = cy * 1.0f / a.height;
        val = (k == 0) ? xx : (k == 1) ? yy
            : (float)(0.5 - (((double)yy - 0.5) * ((double)yy - 0.5) +
                  ((double)xx - 0.5) * ((double)xx - 0.5)));
      }
      mine[(hs1 + j) * CELLS_NT] = val; /* real_inputs = input_layer + hidden_size + 1 */
      sum += val;
    }
    /* maybe_scale_inputs (recur-nn.c:68-81) */
    const float softclip = a.I * INPUT_MEAN_SOFT_TOP;
    const float scale = (sum > softclip) ? soft_clip_dev(sum, softclip) : 1.0f;

    float acc[HN];
#pragma unroll
    for (int c = 0; c < HN / 4; c++) {
      const float4 w = W[c];
      acc[4 * c] = scale * w.x;
      acc[4 * c + 1] = scale * w.y;
      acc[4 * c + 2] = scale * w.z;
      acc[4 * c + 3] = scale * w.w;
    }
#pragma unroll 2
    for (int i = 1; i < hs1; i++)
      cells_row_fma<HN>(acc, mine[i * CELLS_NT] * scale, W + i * (HN / 4));
#pragma unroll 2
    for (int j = 0; j < n_tot; j++)
      cells_row_fma<HN>(acc, mine[(hs1 + j) * CELLS_NT] * scale, W + (hs1 + j) * (HN / 4));

    /* activation (recur-nn.c:121-148), the new state, and the three outputs */
    const float4 wb = cells_wo[0];
    float y0 = wb.x, y1 = wb.y, y2 = wb.z; /* hidden unit 0 is 1 */
    float *sto = a.state + cell;
    sto[0] = 1.0f;
#pragma unroll
    for (int c = 0; c < HN; c++) {
      float t = acc[c];
      if (a.activation == RNN_RESQRT)
        t = (t > 0.0f) ? sqrtf(t + 1.0f) - 1.0f : 0.0f;
      else {
        if (a.activation == RNN_RECLIP20)
          t = t < 20.0f ? t : 20.0f;
        t = (t > 0.0f) ? t : 0.0f;
      }
      if (c + 1 < a.H) {
        sto[(size_t)(c + 1) * a.n] = t;
        const float4 wo = cells_wo[c + 1];
        y0 = fmaf(t, wo.x, y0);
        y1 = fmaf(t, wo.y, y1);
        y2 = fmaf(t, wo.z, y2);
      }
    }
    /* fast_sigmoid (badmaths.h:31-44) and UNIT_TO_BYTE */
    a.frame_out[gcell] = (u8)(1.0f / (1.0f + fast_expf_dev(-y0)) * 255.9f);
    a.frame_out[plane + gcell] = (u8)(1.0f / (1.0f + fast_expf_dev(-y1)) * 255.9f);
    a.frame_out[2 * plane + gcell] = (u8)(1.0f / (1.0f + fast_expf_dev(-y2)) * 255.9f);
  }
}


/* ======================================================================== */
/* the tensor-core frame                                                      */

#define CT_NT 128                    /* cells per tile = MMA rows */
#define CT_BN 64                     /* hidden units 1..64 = MMA columns */
#define CT_A_CHUNK (CT_NT * 128)     /* 64 halves of K for 128 rows: 16 KB */
#define CT_B_CHUNK (CT_BN * 128)     /* 64 halves of K for 64 units: 8 KB */
#define CT_X_TOP 8192.0f             /* rows with larger hidden values are stored scaled down */
#define CT_TILE_BYTES (2 * CT_A_CHUNK) /* a tile's hidden state in HBM: hi plane, lo plane */

/* K, the inner dimension of the product, is laid out for the kernel's
   convenience rather than in the order of the net's input vector:
     [0, 63)    hidden units 1..63 (unit u at u - 1)
     63         the bias (input 0, always 1)
     [64, 104)  the gathered bytes
     [104, 107) the position terms
   and zeros elsewhere.  Every section starts at a place known at compile time;
   the first 64 are exactly what a cell carries from frame to frame.  This maps
   a K index to the net's input row, -1 for padding. */
#define CT_K_BYTES 64
#define CT_K_POS 104
#define CT_K_END 112
__host__ __device__ __forceinline__ int
cells_k_row(int k, int hs, int n_in, int n_pos)
{
  /* the net's input vector: [1 | hidden 1..hs | inputs | pad]: the inputs start
     at hidden_size + 1, not at the aligned h_size (recur-nn-init.c:110-126) */
  if (k < CT_K_BYTES - 1)
    return k + 1 <= hs ? k + 1 : -1;
  if (k == CT_K_BYTES - 1)
    return 0;
  if (k < CT_K_POS)
    return k - CT_K_BYTES < n_in ? hs + 1 + k - CT_K_BYTES : -1;
  return k - CT_K_POS < n_pos ? hs + 1 + n_in + k - CT_K_POS : -1;
}

/* per gathered input j: its neighbour offset, its plane, and for cells whose
   whole neighbourhood is inside the frame the flat distance to the byte */
__constant__ int cells_dx[CT_K_END - CT_K_BYTES], cells_dy[CT_K_END - CT_K_BYTES],
    cells_plane[CT_K_END - CT_K_BYTES], cells_delta[CT_K_END - CT_K_BYTES];

/* The hidden state of the cells lives in HBM as the tensor cores read it: per
   tile of 128 cells the first K chunk of the A operand, FP16 hi plane then lo
   plane, 128-byte rows with their 16-byte pieces XORed with the row number
   (SWIZZLE_128B) - so a tile's state is two bulk copies into shared memory
   going in and two going out, and no thread touches it in between.  Beside it
   per cell: the sum of its hidden values (for the input soft clip) and their
   largest (rows that would not fit FP16 are stored divided by a power of two). */
struct CellsAux {
  float hsum, hmax;
};

struct CellsTcArgs {
  CellsArgs c;
  const rb_h16 *w_image; /* [hi | lo][chunk][64 units][64 halves], swizzled as in shared memory */
  unsigned char *planes; /* [tiles][hi 16 KB | lo 16 KB] */
  CellsAux *aux;         /* [tiles * 128] */
  int ksteps, reach, tiles; /* ksteps: bit s = K steps 16 s .. 16 s + 15 hold inputs */
};

/* the power of two a row with largest value m is stored divided by, and back */
__host__ __device__ __forceinline__ void
cells_row_scale(float m, float &pre, float &post)
{
  pre = post = 1.0f;
  if (m > CT_X_TOP) {
    union { float f; int i; } u;
    u.f = m;
    const int e = ((u.i >> 23) & 0xff) - 127;
    u.i = (127 + 13 - e) << 23;
    pre = u.f;
    u.i = (127 - 13 + e) << 23;
    post = u.f;
  }
}

/* fresh clones: hidden state zero, the bias one */
__global__ void
k_cells_reset(unsigned char *planes, CellsAux *aux, int tiles)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; /* one 16-byte piece each */
  const size_t pieces = (size_t)tiles * CT_TILE_BYTES / 16;
  if (i < pieces) {
    const int in_tile = (int)(i % (CT_TILE_BYTES / 16));
    const int plane = in_tile / (CT_A_CHUNK / 16), r = (in_tile % (CT_A_CHUNK / 16)) >> 3;
    const int piece = in_tile & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (plane == 0 && (piece ^ (r & 7)) == 7)
      v.w = 0x3C000000u; /* K 63 = 1.0 in the hi plane */
    ((uint4 *)planes)[i] = v;
  }
  if (i < (size_t)tiles * CT_NT) {
    aux[i].hsum = 0.0f;
    aux[i].hmax = 0.0f;
  }
}

/* The B operand as the kernel's shared memory holds it: a unit's weights
   along K, 128-byte rows, SWIZZLE_128B, as FP16 hi/lo planes of 64 w. */
__global__ void
k_cells_pack_tc(const float *__restrict__ Wih, const float *__restrict__ Who, int H, int hs,
    int O, int n_in, int n_pos, rb_h16 *image, float *wo)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < CT_BN * 2 * 8) {
    const int u = t % CT_BN, g = t / CT_BN; /* 8 inputs k = 8g .. 8g + 7 of unit u + 1 */
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const int k = cells_k_row(8 * g + e, hs, n_in, n_pos);
      v[e] = (k >= 0 && u + 1 < H) ? Wih[(size_t)k * H + u + 1] : 0.0f;
      /* the bytes go through the tensor cores as the integers they are (exact
         in FP16, no low plane): BYTE_TO_UNIT's 1/255 (gstrnnca.c:640) rides
         on their weights */
      if (8 * g + e >= CT_K_BYTES && 8 * g + e < CT_K_POS)
        v[e] *= 1.0f / 255.0f;
    }
    uint2 h0, l0, h1, l1;
    rb_split4(make_float4(v[0], v[1], v[2], v[3]), (float)RB_W_SCALE, h0, l0);
    rb_split4(make_float4(v[4], v[5], v[6], v[7]), (float)RB_W_SCALE, h1, l1);
    const size_t at = (size_t)(g >> 3) * CT_B_CHUNK + (size_t)u * 128 + (((g & 7) ^ (u & 7)) << 4);
    char *hi = (char *)image, *lo = (char *)image + 2 * CT_B_CHUNK;
    *(uint4 *)(hi + at) = make_uint4(h0.x, h0.y, h1.x, h1.y);
    *(uint4 *)(lo + at) = make_uint4(l0.x, l0.y, l1.x, l1.y);
  }
  if (t < H * 4) {
    int y = t >> 2, o = t & 3;
    wo[t] = (o < O) ? Who[(size_t)y * O + o] : 0.0f;
  }
}

/* eight consecutive inputs of a row -> 16 bytes of each plane, at K = 8 g of
   row r of an operand chunk (128-byte rows, SWIZZLE_128B) */
__device__ __forceinline__ void
cells_put8(unsigned char *hi, unsigned char *lo, int g, int r, const float *v, float pre)
{
  uint2 h0, l0, h1, l1;
  rb_split4(make_float4(v[0], v[1], v[2], v[3]), pre, h0, l0);
  rb_split4(make_float4(v[4], v[5], v[6], v[7]), pre, h1, l1);
  const uint32_t at = (uint32_t)r * 128 + (((g & 7) ^ (r & 7)) << 4);
  *(uint4 *)(hi + at) = make_uint4(h0.x, h0.y, h1.x, h1.y);
  *(uint4 *)(lo + at) = make_uint4(l0.x, l0.y, l1.x, l1.y);
}

__device__ __forceinline__ void
bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void
bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
      ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

/* The state streams through once per frame, a gigabyte of it; the frame's six
   megabytes are read 33 times per cell.  The streams are marked evict-first so
   that the frame stays in L2 under them. */
__device__ __forceinline__ uint64_t
l2_evict_first(void)
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

__device__ __forceinline__ void
bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

__device__ __forceinline__ void
bulk_s2g_hint(void *dst, const void *src, uint32_t bytes, uint64_t pol)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
      ::"l"(dst), "r"(smem_u32(src)), "r"(bytes), "l"(pol) : "memory");
}

/* A CTA per SM, four kinds of warp, two tiles in flight:
 *
 *   warp 29     one thread fetches the tile's hidden state, two bulk copies
 *               from HBM into the first K chunk of A[buf];
 *   warps 0-11  gather: three threads per cell read its neighbourhood's bytes and
 *               write them, with the position terms, as the second K chunk;
 *   warp 28     one thread issues the tile's MMAs into TMEM[buf];
 *   warps 12-27 drain: four threads per row, 16 hidden units each (warps w,
 *               w + 4, w + 8, w + 12 reach the same 32 TMEM lanes): soft clip,
 *               activation, output layer, sigmoid, bytes; the new state goes
 *               back to HBM as planes through a staging tile and bulk copies.
 *
 * While the drain warps finish tile i, the tensor core does tile i + 1 and
 * the copies and the gather for tile i + 2 are under way. */
#define CW_GATHER_ROLES 3
#define CW_GATHER (CW_GATHER_ROLES * CT_NT)
#define CW_GATHER_SLOTS ((CT_K_END - CT_K_BYTES) / CW_GATHER_ROLES) /* K slots per gather thread: 16 */
#define CW_DRAIN_ROLES 4
#define CW_DRAIN (CW_DRAIN_ROLES * CT_NT)
#define CW_DRAIN_COLS (CT_BN / CW_DRAIN_ROLES) /* hidden units a drain thread takes: 16 */
#define CW_MMA_THREAD (CW_GATHER + CW_DRAIN)
#define CW_TMA_THREAD (CW_GATHER + CW_DRAIN + 32)
#define CW_THREADS (CW_GATHER + CW_DRAIN + 64)
#define CW_A_BUF (4 * CT_A_CHUNK) /* one tile's operand: hi chunks 0, 1, lo chunks 0, 1 */
#define CW_SMEM (2 * CW_A_BUF + 4 * CT_B_CHUNK + CT_TILE_BYTES + 1024)

template <int ACT>
__global__ void __launch_bounds__(CW_THREADS, 1)
k_cells_frame_tc(CellsTcArgs t)
{
  extern __shared__ __align__(1024) unsigned char ct_smem[];
  __shared__ __align__(8) uint64_t w_bar, h_full[2], a_full[2], a_empty[2], t_full[2], t_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_sum[2][CW_GATHER_ROLES][CT_NT]; /* [buf][gather role][row]: parts of the soft clip's sum */
  __shared__ float s_post[2][CT_NT];  /* [buf][row]: what the stored row must be multiplied by */
  __shared__ float4 s_y[2][CW_DRAIN_ROLES][CT_NT]; /* [buf][drain role][row]: parts of the outputs, max */
  __shared__ float s_h[2][CW_DRAIN_ROLES][CT_NT];  /* [buf][drain role][row]: parts of the new hidden sum */
  const CellsArgs &a = t.c;
  /* A[buf]: hi chunk 0 (state), hi chunk 1 (gathered), lo chunk 0, lo chunk 1 */
  unsigned char *base = (unsigned char *)(((uintptr_t)ct_smem + 1023) & ~(uintptr_t)1023);
  unsigned char *b_hi = base + 2 * CW_A_BUF, *b_lo = b_hi + 2 * CT_B_CHUNK;
  unsigned char *stage = b_lo + 2 * CT_B_CHUNK; /* the new state of a tile on its way out */
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&w_bar, 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(&h_full[i], 1);
      mbar_init(&a_full[i], CW_GATHER);
      mbar_init(&a_empty[i], 1);
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], CW_DRAIN);
    }
    fence_barrier_init();
    mbar_expect_tx(&w_bar, 4 * CT_B_CHUNK);
    bulk_g2s(b_hi, t.w_image, 4 * CT_B_CHUNK, &w_bar);
  }
  if (warp == 0)
    tmem_alloc(&tmem_slot, 4 * CT_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int plane = a.width * a.height;

  if (tid < CW_GATHER) {
    /* ---- gather ------------------------------------------------------------ */
    /* three threads per cell, 16 K slots each: K 64..79, 80..95 (bytes) and
       96..111 (the last bytes and the position terms) */
    const int r = tid & (CT_NT - 1), role = tid >> 7;
    const int n_in = a.len_y + 2 * a.len_c;
    const int j0 = role * CW_GATHER_SLOTS;
    int it = 0;
    for (int tile = blockIdx.x; tile < t.tiles; tile += gridDim.x, it++) {
      const int buf = it & 1;
      const bool live = tile * CT_NT + r < a.n;
      const int cell = a.cell0 + (live ? tile * CT_NT + r : a.n - 1); /* the ragged end repeats a cell */
      unsigned char *a_hi = base + buf * CW_A_BUF + CT_A_CHUNK, *a_lo = a_hi + 2 * CT_A_CHUNK;
      const CellsAux ax = t.aux[tile * CT_NT + r];
      const int cx = cell % a.width, cy = cell / a.width;
      const bool interior = cx >= t.reach && cx < a.width - t.reach && cy >= t.reach &&
          cy < a.height - t.reach;
      const u8 *fr = a.frame + cell;
      unsigned int bb[CW_GATHER_SLOTS]; /* this thread's bytes; slots past the inputs are 0 */
      if (interior) {
#pragma unroll
        for (int q = 0; q < CW_GATHER_SLOTS; q++)
          bb[q] = (j0 + q < n_in) ? fr[cells_delta[j0 + q]] : 0u;
      }
      else {
        /* get_offset_point (gstrnnca.c:644-667) at the frame's border */
#pragma unroll
        for (int q = 0; q < CW_GATHER_SLOTS; q++) {
          const int j = j0 + q;
          int x = cx + cells_dx[j], y = cy + cells_dy[j];
          if (a.edges) {
            y = max(0, min(a.height - 1, y));
            x = max(0, min(a.width - 1, x));
          }
          else {
            y += (y < 0) ? a.height : (y >= a.height) ? -a.height : 0;
            x += (x < 0) ? a.width : (x >= a.width) ? -a.width : 0;
          }
          bb[q] = (j0 + q < n_in) ? a.frame[cells_plane[j] * plane + y * a.width + x] : 0u;
        }
      }
      float pre, post;
      cells_row_scale(ax.hmax, pre, post);
      /* the sum maybe_scale_inputs takes (recur-nn.c:72-75): bias, hidden, inputs */
      unsigned int total = 0;
#pragma unroll
      for (int q = 0; q < CW_GATHER_SLOTS; q++)
        total += bb[q];
      float sum = (float)total * (1.0f / 255.0f);
      if (role == 0)
        sum += 1.0f + ax.hsum;
      /* bytes as FP16 integers: 0x6400 | b is 1024 + b, minus 1024, times the
         row's power of two - all exact */
      const __half2 k1024 = __floats2half2_rn(1024.0f, 1024.0f);
      const __half2 pre2 = __floats2half2_rn(pre, pre);
      uint32_t hw[CW_GATHER_SLOTS / 2];
#pragma unroll
      for (int q = 0; q < CW_GATHER_SLOTS / 2; q++) {
        uint32_t u = 0x64006400u | bb[2 * q] | (bb[2 * q + 1] << 16);
        __half2 h = __hmul2(__hsub2(*(__half2 *)&u, k1024), pre2);
        hw[q] = *(uint32_t *)&h;
      }
      uint4 lo_last = make_uint4(0, 0, 0, 0);
      if (role == CW_GATHER_ROLES - 1 && a.len_pos > 0) {
        /* the position terms (gstrnnca.c:685-690), K 104..106: floats, both planes */
        const float xx = cx * 1.0f / a.width, yy = cy * 1.0f / a.height;
        float pv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        pv[0] = xx;
        if (a.len_pos > 1)
          pv[1] = yy;
        if (a.len_pos > 2)
          pv[2] = (float)(0.5 -
              (((double)yy - 0.5) * ((double)yy - 0.5) + ((double)xx - 0.5) * ((double)xx - 0.5)));
        sum += pv[0];
        sum += pv[1];
        sum += pv[2];
        uint2 h0, l0, h1, l1;
        rb_split4(make_float4(pv[0], pv[1], pv[2], pv[3]), pre, h0, l0);
        rb_split4(make_float4(pv[4], pv[5], pv[6], pv[7]), pre, h1, l1);
        hw[4] = h0.x;
        hw[5] = h0.y;
        hw[6] = h1.x;
        hw[7] = h1.y;
        lo_last = make_uint4(l0.x, l0.y, l1.x, l1.y);
      }
      if (it >= 2) { /* the MMAs of two tiles ago have read A[buf], its drain s_sum[buf] */
        mbar_wait(&a_empty[buf], ((it >> 1) - 1) & 1);
        mbar_wait(&t_empty[buf], ((it >> 1) - 1) & 1);
      }
#pragma unroll
      for (int g = 0; g < CW_GATHER_SLOTS / 8; g++) {
        const int gg = role * (CW_GATHER_SLOTS / 8) + g;
        const uint32_t at = (uint32_t)r * 128 + (((gg & 7) ^ (r & 7)) << 4);
        *(uint4 *)(a_hi + at) = make_uint4(hw[4 * g], hw[4 * g + 1], hw[4 * g + 2], hw[4 * g + 3]);
        /* the low plane is read only in the K step that holds the position
           terms (K 96..111): zeros under its bytes, the terms' low halves */
        if (gg >= 4)
          *(uint4 *)(a_lo + at) = (gg == 5) ? lo_last : make_uint4(0, 0, 0, 0);
      }
      s_sum[buf][role][r] = sum;
      if (role == 0)
        s_post[buf][r] = post;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&a_full[buf]);
    }
  }
  else if (tid < CW_GATHER + CW_DRAIN) {
    /* ---- drain ------------------------------------------------------------- */
    const int dt = tid - CW_GATHER, r = dt & (CT_NT - 1), role = dt >> 7;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const float softclip = a.I * INPUT_MEAN_SOFT_TOP;
    int it = 0;
    for (int tile = blockIdx.x; tile < t.tiles; tile += gridDim.x, it++) {
      const int buf = it & 1;
      const uint32_t par = (it >> 1) & 1;
      const bool live = tile * CT_NT + r < a.n;
      const int cell = a.cell0 + (live ? tile * CT_NT + r : a.n - 1);
      mbar_wait(&a_full[buf], par); /* the gather's sums are in s_sum */
      /* maybe_scale_inputs (recur-nn.c:68-81): every input times `scale`, so every sum too */
      const float sum = s_sum[buf][0][r] + s_sum[buf][1][r] + s_sum[buf][2][r];
      const float scale = (sum > softclip) ? soft_clip_dev(sum, softclip) : 1.0f;
      const float unscale = scale * s_post[buf][r] * (1.0f / RB_W_SCALE);
      float x[CW_DRAIN_COLS];
      mbar_wait(&t_full[buf], par);
      tc_fence_after();
      {
        float cr[CW_DRAIN_COLS];
        tmem_ld16_nowait(tmem_lane + buf * 2 * CT_BN + role * CW_DRAIN_COLS, x);
        tmem_ld16_nowait(tmem_lane + buf * 2 * CT_BN + CT_BN + role * CW_DRAIN_COLS, cr);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&t_empty[buf]); /* the sums are in registers: TMEM[buf] may be overwritten */
#pragma unroll
        for (int q = 0; q < CW_DRAIN_COLS; q++)
          x[q] = (x[q] + cr[q] * (1.0f / RB_LO_GAIN)) * unscale;
      }
      /* this thread's quarter of the row: activation (recur-nn.c:121-148), its
         share of the three outputs, of the new row's sum and maximum */
      float y0 = 0.0f, y1 = 0.0f, y2 = 0.0f, biggest = 0.0f, hsum = 0.0f;
      if (role == 0) {
        const float4 wb = cells_wo[0]; /* hidden unit 0 is 1 */
        y0 = wb.x;
        y1 = wb.y;
        y2 = wb.z;
      }
#pragma unroll
      for (int q = 0; q < CW_DRAIN_COLS; q++) {
        const int unit = role * CW_DRAIN_COLS + q + 1;
        float v = x[q];
        if (ACT == RNN_RESQRT)
          v = (v > 0.0f) ? sqrtf(v + 1.0f) - 1.0f : 0.0f;
        else {
          if (ACT == RNN_RECLIP20)
            v = fminf(v, 20.0f);
          v = fmaxf(v, 0.0f);
        }
        if (unit <= a.hs) {
          biggest = fmaxf(biggest, v);
          hsum += v;
          const float4 wo = cells_wo[unit];
          y0 = fmaf(v, wo.x, y0);
          y1 = fmaf(v, wo.y, y1);
          y2 = fmaf(v, wo.z, y2);
        }
        else
          v = 0.0f; /* pad units */
        x[q] = v;
      }
      if (role == CW_DRAIN_ROLES - 1)
        x[CW_DRAIN_COLS - 1] = 1.0f; /* K 63 is the bias */
      s_y[buf][role][r] = make_float4(y0, y1, y2, biggest);
      s_h[buf][role][r] = hsum;
      if (dt == 0) /* the staging tile's last journey to HBM has read it */
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(CW_DRAIN) : "memory");
      const float4 p0 = s_y[buf][0][r], p1 = s_y[buf][1][r], p2 = s_y[buf][2][r],
                   p3 = s_y[buf][3][r];
      const float hmax = fmaxf(fmaxf(p0.w, p1.w), fmaxf(p2.w, p3.w));
      float pre, post;
      cells_row_scale(hmax, pre, post);
#pragma unroll
      for (int g = 0; g < CW_DRAIN_COLS / 8; g++)
        cells_put8(stage, stage + CT_A_CHUNK, role * (CW_DRAIN_COLS / 8) + g, r, x + 8 * g, pre);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 2, %0;" ::"n"(CW_DRAIN) : "memory");
      if (dt == 0) {
        bulk_s2g_hint(t.planes + (size_t)tile * CT_TILE_BYTES, stage, CT_TILE_BYTES,
            l2_evict_first());
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      /* fast_sigmoid (badmaths.h:31-44) and UNIT_TO_BYTE (gstrnnca.c:642); the
         parts summed in the order of the hidden units */
      if (role == 3) {
        CellsAux ax;
        ax.hsum = ((s_h[buf][0][r] + s_h[buf][1][r]) + s_h[buf][2][r]) + s_h[buf][3][r];
        ax.hmax = hmax;
        t.aux[tile * CT_NT + r] = ax;
      }
      else if (live) {
        const float y = role == 0 ? ((p0.x + p1.x) + p2.x) + p3.x
            : role == 1 ? ((p0.y + p1.y) + p2.y) + p3.y : ((p0.z + p1.z) + p2.z) + p3.z;
        a.frame_out[role * plane + cell] = (u8)(1.0f / (1.0f + fast_expf_dev(-y)) * 255.9f);
      }
    }
    if (dt == 0)
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  else if (tid == CW_MMA_THREAD) {
    /* ---- the MMA thread ---------------------------------------------------- */
    const uint32_t idesc = umma_idesc_f16(CT_NT, CT_BN, 0, 0);
    mbar_wait(&w_bar, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < t.tiles; tile += gridDim.x, it++) {
      const int buf = it & 1;
      const uint32_t par = (it >> 1) & 1;
      mbar_wait(&h_full[buf], par);
      mbar_wait(&a_full[buf], par);
      if (it >= 2)
        mbar_wait(&t_empty[buf], par ^ 1);
      tc_fence_after();
      const uint32_t a_hi = smem_u32(base + buf * CW_A_BUF);
      const uint32_t a_lo = a_hi + 2 * CT_A_CHUNK;
      const uint32_t acc = tmem_base + buf * 2 * CT_BN;
      bool first = true;
      for (int ks = 0; ks < CT_K_END / 16; ks++) {
        if (!((t.ksteps >> ks) & 1))
          continue;
        const uint32_t ao = (uint32_t)(ks >> 2) * CT_A_CHUNK + (ks & 3) * 32;
        const uint32_t bo = (uint32_t)(ks >> 2) * CT_B_CHUNK + (ks & 3) * 32;
        const uint64_t dah = umma_desc(a_hi + ao, 16, 1024);
        const uint64_t dal = umma_desc(a_lo + ao, 16, 1024);
        const uint64_t dbh = umma_desc(smem_u32(b_hi) + bo, 16, 1024);
        const uint64_t dbl = umma_desc(smem_u32(b_lo) + bo, 16, 1024);
        umma_f16(acc, dah, dbh, idesc, !first);
        umma_f16(acc + CT_BN, dah, dbl, idesc, !first);
        if (ks < CT_K_BYTES / 16 || ks == CT_K_POS / 16) /* bytes have no low plane */
          umma_f16(acc + CT_BN, dal, dbh, idesc, 1u);
        first = false;
      }
      umma_commit(&a_empty[buf]);
      umma_commit(&t_full[buf]);
    }
  }
  else if (tid == CW_TMA_THREAD) {
    /* ---- the state's way in ------------------------------------------------ */
    const uint64_t pol = l2_evict_first();
    int it = 0;
    for (int tile = blockIdx.x; tile < t.tiles; tile += gridDim.x, it++) {
      const int buf = it & 1;
      if (it >= 2)
        mbar_wait(&a_empty[buf], ((it >> 1) - 1) & 1);
      unsigned char *a_hi = base + buf * CW_A_BUF;
      const unsigned char *src = t.planes + (size_t)tile * CT_TILE_BYTES;
      mbar_expect_tx(&h_full[buf], CT_TILE_BYTES);
      bulk_g2s_hint(a_hi, src, CT_A_CHUNK, &h_full[buf], pol);
      bulk_g2s_hint(a_hi + 2 * CT_A_CHUNK, src + CT_A_CHUNK, CT_A_CHUNK, &h_full[buf], pol);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 4 * CT_BN);
  }
}

struct RnnCells;
/* whose neighbourhood tables the constant memory holds, and for which pattern */
static const struct RnnCells *cells_tables_owner = NULL;
static size_t cells_tables_bytes = 0;
static int cells_tables_key[4 * CELLS_XIN + 2];

struct RnnCells {
  RecurNN *proto;
  int width, height;
  int n, tiles;               /* cells this GPU runs: all of them, or a band of rows */
  int cell0;                  /* the band's first cell */
  int frame_n;                /* width * height */
  int ranks;                  /* > 1: this rank computes a band; rows are exchanged between frames */
  int whole;                  /* the newest frame on this rank is a whole picture */
  unsigned char *halo;        /* device [ranks][3 planes][2 sides][reach rows][width] */
  int halo_each;              /* bytes per rank it has room for */
  int edges_now;
  unsigned char *planes;      /* device [tiles][32 KB]: the hidden state (see CellsTcArgs) */
  CellsAux *aux;              /* device [tiles * 128] */
  float *state;               /* device [h_size][n]: the FP32 cross-check kernel's state */
  u8 *frames;                 /* device: two frames of 3 n bytes, ping-pong */
  int *off_dev;
  u8 *host;                   /* pinned staging, 3 n bytes */
  int cur;                    /* which of the two frames holds the newest picture */
  rb_h16 *w_image;            /* device: the weights as k_cells_frame_tc's B operand */
  int reach;                  /* largest |dx|, |dy| of the neighbourhood */
  int fma;                    /* this object runs the FP32 kernel (RECUR_B200_CELLS_FMA at creation) */
  long frames_run;            /* since the last rnn_cells_forget */
};

static void
cells_reset(RnnCells *c)
{
  if (c->fma)
    cudaMemsetAsync(c->state, 0, (size_t)c->n * c->proto->h_size * sizeof(float), rb_stream);
  else {
    const size_t pieces = (size_t)c->tiles * CT_TILE_BYTES / 16;
    k_cells_reset<<<(unsigned)((pieces + 255) / 256), 256, 0, rb_stream>>>(c->planes, c->aux,
        c->tiles);
    LAUNCH_CHECK("k_cells_reset");
  }
  c->frames_run = 0;
}

static RnnCells *
cells_new(RecurNN *prototype, int width, int height, int rank, int ranks)
{
  rb_require_device("rnn_cells_new");
  if (!prototype || width < 1 || height < 1) {
    fprintf(stderr, "rnn_cells_new: need a prototype net and a frame size\n");
    return NULL;
  }
  if (prototype->h_size > 64 || prototype->output_size < 3 || prototype->bottom_layer ||
      prototype->input_size > CELLS_XIN || prototype->i_size > CELLS_MAX_I) {
    fprintf(stderr, "rnn_cells_new: the cell kernel takes nets of up to 63 hidden units, "
        "%d inputs, at least 3 outputs and no bottom layer\n", CELLS_XIN);
    return NULL;
  }
  rb_net_of(prototype); /* aborts on a foreign pointer */
  RnnCells *c = (RnnCells *)calloc(1, sizeof(RnnCells));
  c->proto = prototype;
  c->width = width;
  c->height = height;
  c->frame_n = width * height;
  c->ranks = ranks;
  c->n = (height / ranks) * width;
  c->cell0 = rank * c->n;
  c->tiles = cdiv(c->n, CT_NT);
  const char *env = getenv("RECUR_B200_CELLS_FMA");
  c->fma = (env && *env && *env != '0');
  size_t n = (size_t)c->frame_n; /* frames are whole on every rank */
  cudaError_t e = c->fma
      ? cudaMalloc((void **)&c->state, (size_t)c->n * prototype->h_size * sizeof(float))
      : cudaMalloc((void **)&c->planes, (size_t)c->tiles * CT_TILE_BYTES);
  if (e != cudaSuccess ||
      cudaMalloc((void **)&c->aux, (size_t)c->tiles * CT_NT * sizeof(CellsAux)) != cudaSuccess ||
      cudaMalloc((void **)&c->frames, 2 * 3 * n + 64) != cudaSuccess ||
      cudaMalloc((void **)&c->off_dev, 4 * CELLS_XIN * sizeof(int)) != cudaSuccess ||
      cudaMalloc((void **)&c->w_image, 4 * CT_B_CHUNK) != cudaSuccess ||
      cudaHostAlloc((void **)&c->host, 3 * n, cudaHostAllocDefault) != cudaSuccess)
    rb_die("recur-b200: out of memory for %d cells", c->n);
  if (ranks > 1) {
    c->halo_each = 3 * 2 * 8 * width; /* neighbourhoods up to 8 rows away */
    if (cudaMalloc((void **)&c->halo, (size_t)ranks * c->halo_each) != cudaSuccess)
      rb_die("recur-b200: out of memory for the halo rows");
  }
  c->whole = 1;
  cells_reset(c);
  cudaMemsetAsync(c->frames, 0, 2 * 3 * n, rb_stream);
  cudaStreamSynchronize(rb_stream);
  return c;
}

extern "C" RnnCells *
rnn_cells_new(RecurNN *prototype, int width, int height)
{
  return cells_new(prototype, width, height, 0, 1);
}

/* The same frame shared by the ranks of rnn_b200_comm_join:
   rank r of n keeps the hidden state of rows [r h/n, (r+1) h/n) and computes
   those; after every frame the bands are gathered (NCCL all-gather of the
   three planes), so every rank holds, and returns, the whole picture.  The
   frame calls become collective. */
extern "C" RnnCells *
rnn_cells_new_sharded(RecurNN *prototype, int width, int height)
{
  const int ranks = rb_comm_size();
  if (height % ranks) {
    fprintf(stderr, "rnn_cells_new_sharded: %d rows do not divide among %d GPUs\n", height, ranks);
    return NULL;
  }
  return cells_new(prototype, width, height, rb_comm_rank(), ranks);
}

extern "C" void
rnn_cells_delete(RnnCells *c)
{
  if (!c)
    return;
  if (cells_tables_owner == c)
    cells_tables_owner = NULL; /* a later object may get the same address */
  cudaStreamSynchronize(rb_stream);
  cudaFree(c->state);
  cudaFree(c->planes);
  cudaFree(c->aux);
  cudaFree(c->frames);
  cudaFree(c->off_dev);
  cudaFree(c->w_image);
  cudaFree(c->halo);
  cudaFreeHost(c->host);
  free(c);
}

extern "C" void
rnn_cells_forget(RnnCells *c)
{
  cells_reset(c);
}

static void
cells_launch(RnnCells *c, const u8 *in, u8 *out, int len_y, int len_c, int len_pos, int edges)
{
  RecurNN *p = c->proto;
  CellsArgs a;
  a.Wih = p->ih_weights;
  a.Who = p->ho_weights;
  a.I = p->i_size;
  a.H = p->h_size;
  a.O = p->o_size;
  a.hs = p->hidden_size;
  a.activation = p->activation;
  a.state = c->state;
  a.n = c->n;
  a.cell0 = c->cell0;
  a.frame = in;
  a.frame_out = out;
  a.width = c->width;
  a.height = c->height;
  a.off_y = c->off_dev;
  a.off_c = c->off_dev + 2 * len_y;
  a.len_y = len_y;
  a.len_c = len_c;
  a.len_pos = len_pos;
  a.edges = edges;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  static float *w_dev, *wo_dev;
  if (!w_dev) {
    cudaGetSymbolAddress((void **)&w_dev, cells_w);
    cudaGetSymbolAddress((void **)&wo_dev, cells_wo);
  }
  c->frames_run++;
  rb_prof_begin(RB_PROF_FWD);
  if (!c->fma) {
    CellsTcArgs t;
    t.c = a;
    t.w_image = c->w_image;
    t.planes = c->planes;
    t.aux = c->aux;
    const int n_in = len_y + 2 * len_c;
    t.ksteps = 0;
    for (int k = 0; k < CT_K_END; k++)
      if (cells_k_row(k, p->hidden_size, n_in, len_pos) >= 0)
        t.ksteps |= 1 << (k / 16);
    t.reach = c->reach;
    t.tiles = c->tiles;
    void (*kernel)(CellsTcArgs) = p->activation == RNN_RESQRT ? k_cells_frame_tc<RNN_RESQRT>
        : p->activation == RNN_RECLIP20 ? k_cells_frame_tc<RNN_RECLIP20>
        : k_cells_frame_tc<RNN_RELU>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM) !=
        cudaSuccess)
      rb_die("recur-b200: k_cells_frame_tc: cannot reserve shared memory");
    /* the prototype may have been trained since the last frame: repack every time (2 us) */
    k_cells_pack_tc<<<cdiv(CT_BN * 2 * 8, 256), 256, 0, rb_stream>>>(p->ih_weights,
        p->ho_weights, p->h_size, p->hidden_size, p->o_size, n_in, len_pos, c->w_image, wo_dev);
    /* a CTA per SM (194 KB of shared memory, 256 TMEM columns) */
    int blocks = t.tiles < sms ? t.tiles : sms;
    kernel<<<blocks, CW_THREADS, CW_SMEM, rb_stream>>>(t);
    LAUNCH_CHECK("k_cells_frame_tc");
    rb_prof_end(RB_PROF_FWD);
    return;
  }
  static int per_sm[2];
  const int big = p->h_size > 52;
  const int HN = big ? 64 : 52;
  const size_t sh = (size_t)p->i_size * CELLS_NT * sizeof(float);
  if (!per_sm[big]) {
    if (cudaFuncSetAttribute(big ? k_cells_frame<64> : k_cells_frame<52>,
            cudaFuncAttributeMaxDynamicSharedMemorySize, CELLS_MAX_I * CELLS_NT * sizeof(float)) !=
        cudaSuccess)
      rb_die("recur-b200: k_cells_frame: cannot reserve shared memory");
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[big],
        big ? k_cells_frame<64> : k_cells_frame<52>, CELLS_NT, sh);
    if (per_sm[big] < 1)
      per_sm[big] = 1;
  }
  int blocks = cdiv(c->n, CELLS_NT);
  if (blocks > sms * per_sm[big])
    blocks = sms * per_sm[big];
  k_cells_pack<<<cdiv(p->i_size * HN, 256), 256, 0, rb_stream>>>(p->ih_weights, p->ho_weights,
      p->i_size, p->h_size, p->o_size, HN, w_dev, wo_dev);
  if (!big)
    k_cells_frame<52><<<blocks, CELLS_NT, sh, rb_stream>>>(a);
  else
    k_cells_frame<64><<<blocks, CELLS_NT, sh, rb_stream>>>(a);
  LAUNCH_CHECK("k_cells_frame");
  rb_prof_end(RB_PROF_FWD);
}

/* every rank's band of the picture, to every rank */
static void
cells_gather(RnnCells *c, u8 *out)
{
  if (c->ranks <= 1 || c->whole)
    return;
  unsigned char *planes[3] = {out, out + c->frame_n, out + 2 * (size_t)c->frame_n};
  rb_comm_allgather_inplace(planes, 3, (size_t)c->n);
  c->whole = 1;
}

/* [plane][side: top rows, bottom rows][reach][width] of this rank's band */
__global__ void
k_cells_halo_pack(const u8 *__restrict__ frame, u8 *__restrict__ mine, int width, int frame_n,
    int row0, int rows, int reach)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * 2 * reach * width)
    return;
  const int x = i % width, row = (i / width) % reach, side = (i / (width * reach)) & 1;
  const int plane = i / (2 * width * reach);
  const int g = side ? row0 + rows - reach + row : row0 + row;
  mine[i] = frame[(size_t)plane * frame_n + (size_t)g * width + x];
}

/* the rows just above this rank's band (the bottom rows of the rank before) and
   just below it (the top rows of the rank after), around the frame's edge when
   it wraps (get_offset_point's edges == 0, gstrnnca.c:655-664) */
__global__ void
k_cells_halo_unpack(u8 *__restrict__ frame, const u8 *__restrict__ all, int width, int height,
    int frame_n, int row0, int rows, int reach, int rank, int ranks, int edges)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * 2 * reach * width)
    return;
  const int x = i % width, row = (i / width) % reach, below = (i / (width * reach)) & 1;
  const int plane = i / (2 * width * reach);
  int g = below ? row0 + rows + row : row0 - reach + row;
  if (g < 0 || g >= height) {
    if (edges)
      return; /* clamped edges never look past the frame */
    g += g < 0 ? height : -height;
  }
  const int from = below ? (rank + 1) % ranks : (rank + ranks - 1) % ranks;
  const size_t per_rank = (size_t)3 * 2 * reach * width;
  /* the neighbour's side facing us: its top rows if it is below us */
  const size_t at = (((size_t)plane * 2 + (below ? 0 : 1)) * reach + row) * width + x;
  frame[(size_t)plane * frame_n + (size_t)g * width + x] = all[from * per_rank + at];
}

/* Between frames nobody needs the whole picture: a cell reads `reach` rows
   past its band at most, so the ranks swap just those rows (SURVEY.md 8e's
   "2-row halo exchange") and gather whole pictures only when one is asked
   for. */
static void
cells_halo(RnnCells *c, u8 *out)
{
  if (c->ranks <= 1)
    return;
  const int rows = c->n / c->width;
  if (c->reach > rows || c->reach * c->width * 6 > c->halo_each) {
    c->whole = 0;
    cells_gather(c, out); /* neighbourhoods wider than a band: everything to everybody */
    return;
  }
  c->whole = 0;
  if (c->reach == 0)
    return;
  const int count = 3 * 2 * c->reach * c->width;
  const size_t each = (size_t)count;
  unsigned char *all = c->halo;
  k_cells_halo_pack<<<cdiv(count, 256), 256, 0, rb_stream>>>(out, all + (size_t)rb_comm_rank() * each,
      c->width, c->frame_n, c->cell0 / c->width, rows, c->reach);
  LAUNCH_CHECK("k_cells_halo_pack");
  rb_comm_allgather_inplace(&all, 1, each);
  k_cells_halo_unpack<<<cdiv(count, 256), 256, 0, rb_stream>>>(out, all, c->width, c->height,
      c->frame_n, c->cell0 / c->width, rows, c->reach, rb_comm_rank(), c->ranks, c->edges_now);
  LAUNCH_CHECK("k_cells_halo_unpack");
}

static void
cells_check(RnnCells *c, int len_y, int len_c, int len_pos)
{
  if (c->proto->input_size != len_y + 2 * len_c + len_pos || len_y < 0 || len_c < 0 ||
      len_pos < 0 || len_pos > 3)
    rb_die("recur-b200: rnn_cells: a net of %d inputs does not match %d + 2*%d + %d",
        c->proto->input_size, len_y, len_c, len_pos);
}

static void
cells_offsets(RnnCells *c, const int *offsets_y, int len_y, const int *offsets_c, int len_c)
{
  int off[4 * CELLS_XIN + 2];
  off[0] = len_y;
  off[1] = len_c;
  memcpy(off + 2, offsets_y, 2 * (size_t)len_y * sizeof(int));
  memcpy(off + 2 + 2 * len_y, offsets_c, 2 * (size_t)len_c * sizeof(int));
  /* the element passes the same pattern frame after frame: the tables on the
     device (constant memory: one set for the library) stay while nothing else
     has written them */
  const size_t off_bytes = (2 + 2 * (size_t)(len_y + len_c)) * sizeof(int);
  if (cells_tables_owner == c && cells_tables_bytes == off_bytes &&
      !memcmp(cells_tables_key, off, off_bytes))
    return;
  cells_tables_owner = c;
  cells_tables_bytes = off_bytes;
  memcpy(cells_tables_key, off, off_bytes);
  cudaMemcpyAsync(c->off_dev, off + 2, 2 * (size_t)(len_y + len_c) * sizeof(int),
      cudaMemcpyHostToDevice, rb_stream);
  /* the same per gathered input, in the order of fill_net_inputs (gstrnnca.c:672-684) */
  int dx[CELLS_XIN], dy[CELLS_XIN], pl[CELLS_XIN], delta[CELLS_XIN];
  const int plane = c->frame_n;
  int reach = 0, j = 0;
  for (int i = 0; i < len_y; i++, j++) {
    dx[j] = offsets_y[2 * i];
    dy[j] = offsets_y[2 * i + 1];
    pl[j] = 0;
  }
  for (int i = 0; i < len_c; i++) {
    for (int k = 1; k <= 2; k++, j++) {
      dx[j] = offsets_c[2 * i];
      dy[j] = offsets_c[2 * i + 1];
      pl[j] = k;
    }
  }
  for (int i = 0; i < j; i++) {
    delta[i] = pl[i] * plane + dy[i] * c->width + dx[i];
    reach = abs(dx[i]) > reach ? abs(dx[i]) : reach;
    reach = abs(dy[i]) > reach ? abs(dy[i]) : reach;
  }
  c->reach = reach;
  cudaMemcpyToSymbolAsync(cells_dx, dx, j * sizeof(int), 0, cudaMemcpyHostToDevice, rb_stream);
  cudaMemcpyToSymbolAsync(cells_dy, dy, j * sizeof(int), 0, cudaMemcpyHostToDevice, rb_stream);
  cudaMemcpyToSymbolAsync(cells_plane, pl, j * sizeof(int), 0, cudaMemcpyHostToDevice, rb_stream);
  cudaMemcpyToSymbolAsync(cells_delta, delta, j * sizeof(int), 0, cudaMemcpyHostToDevice,
      rb_stream);
  cudaStreamSynchronize(rb_stream); /* the tables are on this stack */
}

static bool
cells_is_pinned(const void *p)
{
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

/* One frame (reference gstrnnca.c:805-830, fill_frame): host frame in, host
   frame out; the hidden state of every cell stays on the device. */
extern "C" void
rnn_cells_rnnca_frame(RnnCells *c, const unsigned char *frame_in, unsigned char *frame_out,
    const int *offsets_y, int len_y, const int *offsets_c, int len_c, int len_pos, int edges)
{
  cells_check(c, len_y, len_c, len_pos);
  rb_matrices_to_device(c->proto);
  cells_offsets(c, offsets_y, len_y, offsets_c, len_c);
  const size_t fb = 3 * (size_t)c->frame_n;
  u8 *in = c->frames, *out = c->frames + fb;
  /* frames in page-locked memory (cudaHostAlloc / cudaHostRegister, a video
     pipeline's buffer pool) are copied from and to in place; pageable ones go
     through the staging buffer */
  const bool in_pinned = cells_is_pinned(frame_in), out_pinned = cells_is_pinned(frame_out);
  if (!in_pinned)
    memcpy(c->host, frame_in, fb);
  cudaMemcpyAsync(in, in_pinned ? frame_in : c->host, fb, cudaMemcpyHostToDevice, rb_stream);
  cells_launch(c, in, out, len_y, len_c, len_pos, edges);
  c->whole = 0;
  cells_gather(c, out); /* the caller wants the picture */
  cudaMemcpyAsync(out_pinned ? frame_out : c->host, out, fb, cudaMemcpyDeviceToHost, rb_stream);
  cudaStreamSynchronize(rb_stream);
  if (!out_pinned)
    memcpy(frame_out, c->host, fb);
  c->cur = 1;
}

/* The automaton running by itself, as the element does between key frames:
   n_frames steps, every frame the input of the next, the pictures never
   leaving the device.  frame_in (may be NULL: go on from the last picture) is
   uploaded first, frame_out (may be NULL) receives the last picture. */
extern "C" void
rnn_cells_rnnca_run(RnnCells *c, const unsigned char *frame_in, int n_frames,
    unsigned char *frame_out, const int *offsets_y, int len_y, const int *offsets_c, int len_c,
    int len_pos, int edges)
{
  cells_check(c, len_y, len_c, len_pos);
  rb_matrices_to_device(c->proto);
  cells_offsets(c, offsets_y, len_y, offsets_c, len_c);
  const size_t fb = 3 * (size_t)c->frame_n;
  if (frame_in) {
    memcpy(c->host, frame_in, fb);
    cudaMemcpyAsync(c->frames + (size_t)c->cur * fb, c->host, fb, cudaMemcpyHostToDevice,
        rb_stream);
    c->whole = 1;
  }
  c->edges_now = edges;
  for (int f = 0; f < n_frames; f++) {
    cells_launch(c, c->frames + (size_t)c->cur * fb, c->frames + (size_t)(c->cur ^ 1) * fb, len_y,
        len_c, len_pos, edges);
    cells_halo(c, c->frames + (size_t)(c->cur ^ 1) * fb);
    c->cur ^= 1;
  }
  if (frame_out) {
    cells_gather(c, c->frames + (size_t)c->cur * fb); /* collective: every rank asks or none */
    cudaMemcpyAsync(c->host, c->frames + (size_t)c->cur * fb, fb, cudaMemcpyDeviceToHost,
        rb_stream);
    cudaStreamSynchronize(rb_stream);
    memcpy(frame_out, c->host, fb);
  }
}

/* one cell's hidden_layer (h_size floats) */
extern "C" void
rnn_cells_get_hidden(RnnCells *c, int cell, float *hidden)
{
  const int H = c->proto->h_size;
  cell -= c->cell0; /* on a sharded object: cells of this rank's band only */
  if (cell < 0 || cell >= c->n)
    rb_die("recur-b200: rnn_cells_get_hidden: cell %d is not on this GPU", cell + c->cell0);
  if (c->fma) {
    cudaMemcpy2DAsync(hidden, sizeof(float), c->state + cell, (size_t)c->n * sizeof(float),
        sizeof(float), H, cudaMemcpyDeviceToHost, rb_stream);
    cudaStreamSynchronize(rb_stream);
    return;
  }
  /* the row of the cell's tile, out of both planes */
  const int tile = cell / CT_NT, r = cell % CT_NT;
  rb_h16 hi[64], lo[64];
  CellsAux ax;
  const unsigned char *src = c->planes + (size_t)tile * CT_TILE_BYTES + (size_t)r * 128;
  cudaMemcpyAsync(hi, src, 128, cudaMemcpyDeviceToHost, rb_stream);
  cudaMemcpyAsync(lo, src + CT_A_CHUNK, 128, cudaMemcpyDeviceToHost, rb_stream);
  cudaMemcpyAsync(&ax, c->aux + cell, sizeof(ax), cudaMemcpyDeviceToHost, rb_stream);
  cudaStreamSynchronize(rb_stream);
  float pre, post;
  cells_row_scale(ax.hmax, pre, post);
  hidden[0] = c->frames_run ? 1.0f : 0.0f; /* rnn_opinion sets it (recur-nn.c:144) */
  for (int u = 1; u < H; u++) {
    const int k = u - 1, at = (((k >> 3) ^ (r & 7)) << 3) + (k & 7);
    hidden[u] = (rb_half_to_float(hi[at]) + rb_half_to_float(lo[at]) * (1.0f / RB_LO_GAIN)) * post;
  }
}
