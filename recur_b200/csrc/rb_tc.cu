/* rb_tc.cu — the tensor-core engine: the three big contractions of the path
 * as tcgen05 (5th-gen tensor core) kernels for sm_100a, used when a batch has
 * >= 64 streams (a multiple of 32) in one contiguous run of pool slots.
 *
 *   FWD    hidden[b, h]   = act( sum_y x[b, y] * Wih[y, h] )       (recur-nn.c:117-148)
 *   CHAIN  E(k+1)[b, y]   = mask * sum_x E(k)[b, x] * Wih[y, x]    (recur-nn.c:338-376)
 *   DW     delta[y, x]   += sum_{k,b} x_k[b, y] * E(k)[b, x]       (recur-nn.c:353-356)
 *
 * Numerics: every FP32 operand is kept as two FP16 planes (rb_split.cuh); a
 * product is three kind::f16 MMAs into two FP32 accumulators in tensor
 * memory, 22 significant bits per operand - FP32-faithful within the 1e-4
 * parity tolerance at half the bytes and twice the MMA rate of TF32 planes.
 * The planes are materialised once where each operand is produced (weights:
 * in the update kernel; ring rows: when the row enters the ring; error rows:
 * by the kernel that computes them), so the GEMM main loops are pure
 * TMA -> shared memory -> tcgen05.mma pipelines.
 *
 * Kernel anatomy: warp 0 TMA producer (one lane), warp 1 TMEM allocator + MMA
 * issuer (one lane), warps 2..5 epilogue (one TMEM lane quarter each); the
 * persistent chain kernel adds two warps.  Operand tiles are swizzled; FWD and
 * CHAIN read both operands K-major, DW reads both MN-major (the contraction
 * runs over ring rows).
 */
#include "rb_kernels.h"
#include "rb_host.h"
#include "rb_optim.cuh"
#include "rb_split.cuh"
#include <cuda.h>
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

static inline int
cdiv(int a, int b)
{
  return (a + b - 1) / b;
}

#include "rb_ptx.cuh"

__device__ __forceinline__ float
block_sum_tc(float v, float *scratch /* >= 33 floats */)
{
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0)
    scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    float t = (lane < nw) ? scratch[lane] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0)
      scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

/* ======================================================================== */
/* the engine's extra state per pool                                          */

#define TC_BM 128      /* tile rows (TMEM lanes) */
#define TC_KB 64       /* K per stage: 64 halves = one 128-byte swizzle row */
#define TC_FWD_BN 64      /* FWD tile columns (unsplit variant) */
#define TC_FWD_STAGES 4
#define TC_NT_BN 128      /* split-K FWD and per-step CHAIN tile columns */
#define TC_NT_STAGES 3
#define TC_NT_SPLITS 4    /* split-K of those */
#define TC_DW_BK 32       /* DW ring rows per stage (two MMAs of K = 16) */
#define TC_DW2_BN 192     /* N of the CTA-pair weight-gradient tile (two halves of 96) */
#define TC_DW_SPLITS 4    /* most split-K planes the weight gradient writes */
#define CH_SQ_SLOTS 64    /* floats per stream in the chain's sum-of-squares exchange */
#define CH_SYNC_STRIDE 32 /* words between the barrier counters of two stream groups */

typedef struct RbTc {
  int cap, depth;
  int pitch;            /* halves per ring / chain plane row: i_size rounded up to 64 */
  int wpitch;           /* halves per Wih plane row: h_size rounded up to 64 */
  rb_h16 *Xhi, *Xlo;    /* [depth][cap][pitch]     planes of the ring */
  rb_h16 *Ehi, *Elo;    /* [depth+1][cap][pitch]   planes of the error chain */
  rb_h16 *Whi, *Wlo;    /* [i_size][wpitch]        Wih, K-major B of CHAIN */
  rb_h16 *WThi, *WTlo;  /* [h_size][pitch]         Wih^T, K-major B of FWD */
  float *escale;        /* device: the chain planes' scale of the current walk */
  float *partial;       /* [TC_DW_SPLITS][i_size][h_size] split-K planes of the weight gradient */
  float *cpartial;      /* [TC_NT_SPLITS][cap][i_size rounded up to 32] split-K partial sums */
  float *sqpart;        /* [2][m tiles][128][CH_SQ_SLOTS] per-CTA sums of squares of the chain */
  unsigned int *sync;   /* barrier counters of the persistent chain + kmax, two areas */
  size_t sync_words;
  int sync_flip;
  int delta_pending;    /* the last weight gradient is not in ih_delta yet: 1 = unsummed in
                           `partial`, 2 = summed over all ranks in the peer exchange's result
                           block (both deltas), behind its second flag round */
  int pending_accumulate;
  void *pending_p2p;
  const float *w_src;   /* weights the planes were made from */
  uint64_t w_version;
  /* tensor maps */
  CUtensorMap mXhi_k, mXlo_k;   /* ring rows as K-major A of FWD: box 64 x 128 */
  CUtensorMap mEhi_k, mElo_k;   /* error rows as K-major A of CHAIN (width h_size) */
  CUtensorMap mWhi_k, mWlo_k;   /* Wih rows as K-major B of CHAIN: box 64 x 128 */
  CUtensorMap mWThi_k, mWTlo_k; /* Wih^T rows as K-major B of FWD: box 64 x 64 */
  CUtensorMap mWThi_k128, mWTlo_k128; /* the same with 128-row boxes: split-K FWD */
  CUtensorMap mEhi_mn, mElo_mn; /* error rows as MN-major A of DW: 2 chunks of 64 columns */
  CUtensorMap mXhi_mn, mXlo_mn; /* ring rows as MN-major B half of DW: 3 chunks of 32 columns */
  int dw_splits;        /* split-K planes the last weight gradient wrote */
} RbTc;

typedef CUresult (*encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_fn_t
get_encode(void)
{
  static encode_fn_t fn = NULL;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void *p = NULL;
    CUDA_OR_DIE(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess)
      rb_die("recur-b200: the driver lacks cuTensorMapEncodeTiled");
    fn = (encode_fn_t)p;
  }
  return fn;
}

/* rows x width halves with a row pitch of pitch halves, boxes of 64 x box_rows
   (K-major operand tiles, 128-byte swizzle); reads past `width` or `rows`
   return zeros */
static void
make_map(CUtensorMap *m, rb_h16 *base, uint64_t width, uint64_t rows, uint64_t pitch,
    uint32_t box_rows)
{
  cuuint64_t dims[2] = {width, rows};
  cuuint64_t strides[1] = {pitch * sizeof(rb_h16)};
  cuuint32_t box[2] = {TC_KB, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box,
      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    rb_die("recur-b200: cuTensorMapEncodeTiled failed (%d) for %llu x %llu pitch %llu", (int)r,
        (unsigned long long)width, (unsigned long long)rows, (unsigned long long)pitch);
}

/* The same rows seen as [chunk of `chunk` columns][row][chunk halves], so
   that one TMA fetches n_chunks column chunks of box_rows rows into
   consecutive (chunk-major) shared memory: the MN-major operand layout of DW
   (chunk 64: 128-byte swizzle, chunk 32: 64-byte swizzle). */
static void
make_map_chunked(CUtensorMap *m, rb_h16 *base, uint64_t width, uint64_t rows, uint64_t pitch,
    uint32_t box_rows, uint32_t chunk, uint32_t n_chunks)
{
  cuuint64_t dims[3] = {chunk, rows, (width + chunk - 1) / chunk};
  cuuint64_t strides[2] = {pitch * sizeof(rb_h16), chunk * sizeof(rb_h16)};
  cuuint32_t box[3] = {chunk, box_rows, n_chunks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box,
      estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      chunk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    rb_die("recur-b200: cuTensorMapEncodeTiled (chunked) failed (%d)", (int)r);
}

template <typename T>
static T *
dmalloc0(size_t n)
{
  T *p = NULL;
  cudaError_t e = cudaMalloc((void **)&p, n * sizeof(T));
  if (e != cudaSuccess)
    rb_die("recur-b200: cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
  CUDA_OR_DIE(cudaMemsetAsync(p, 0, n * sizeof(T), rb_stream));
  return p;
}

static void
tc_free(RbTc *t)
{
  if (!t)
    return;
  cudaFree(t->Xhi); cudaFree(t->Xlo); cudaFree(t->Ehi); cudaFree(t->Elo);
  cudaFree(t->Whi); cudaFree(t->Wlo); cudaFree(t->WThi); cudaFree(t->WTlo);
  cudaFree(t->escale);
  cudaFree(t->partial);
  cudaFree(t->cpartial);
  cudaFree(t->sqpart);
  cudaFree(t->sync);
  free(t);
}

extern "C" void
rb_tc_pool_release(RbPool *p)
{
  tc_free((RbTc *)p->tc);
  p->tc = NULL;
}

static RbTc *
tc_state(RbPool *p)
{
  RbTc *t = (RbTc *)p->tc;
  if (t && t->cap == p->cap && t->depth == p->depth)
    return t;
  if (t) {
    CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
    tc_free(t);
  }
  const RbDims *d = &p->group->d;
  const size_t I = d->i_size, H = d->h_size;
  t = (RbTc *)calloc(1, sizeof(RbTc));
  t->cap = p->cap;
  t->depth = p->depth;
  t->pitch = (int)((I + 63) & ~(size_t)63);
  t->wpitch = (int)((H + 63) & ~(size_t)63);
  const size_t P = t->pitch, WP = t->wpitch;
  size_t ring_rows = (size_t)p->depth * p->cap, chain_rows = (size_t)(p->depth + 1) * p->cap;
  t->Xhi = dmalloc0<rb_h16>(ring_rows * P);
  t->Xlo = dmalloc0<rb_h16>(ring_rows * P);
  t->Ehi = dmalloc0<rb_h16>(chain_rows * P);
  t->Elo = dmalloc0<rb_h16>(chain_rows * P);
  t->Whi = dmalloc0<rb_h16>(I * WP);
  t->Wlo = dmalloc0<rb_h16>(I * WP);
  t->WThi = dmalloc0<rb_h16>(H * P);
  t->WTlo = dmalloc0<rb_h16>(H * P);
  t->escale = dmalloc0<float>(4);
  t->partial = dmalloc0<float>((size_t)TC_DW_SPLITS * I * H);
  t->cpartial = dmalloc0<float>((size_t)TC_NT_SPLITS * p->cap * ((I + 31) & ~(size_t)31));
  const size_t m_tiles = cdiv(p->cap, TC_BM) + 1;
  t->sqpart = dmalloc0<float>(2 * m_tiles * TC_BM * CH_SQ_SLOTS);
  /* two areas, used by alternate walks (see k_finalize_rows) */
  t->sync_words = m_tiles * CH_SYNC_STRIDE + 8;
  t->sync = dmalloc0<unsigned int>(2 * t->sync_words);
  t->sync_flip = 0;
  t->w_src = NULL;
  make_map(&t->mXhi_k, t->Xhi, I, ring_rows, P, TC_BM);
  make_map(&t->mXlo_k, t->Xlo, I, ring_rows, P, TC_BM);
  make_map(&t->mEhi_k, t->Ehi, H, chain_rows, P, TC_BM);
  make_map(&t->mElo_k, t->Elo, H, chain_rows, P, TC_BM);
  make_map(&t->mWhi_k, t->Whi, H, I, WP, TC_NT_BN);
  make_map(&t->mWlo_k, t->Wlo, H, I, WP, TC_NT_BN);
  make_map(&t->mWThi_k, t->WThi, I, H, P, TC_FWD_BN);
  make_map(&t->mWTlo_k, t->WTlo, I, H, P, TC_FWD_BN);
  make_map(&t->mWThi_k128, t->WThi, I, H, P, TC_NT_BN);
  make_map(&t->mWTlo_k128, t->WTlo, I, H, P, TC_NT_BN);
  make_map_chunked(&t->mEhi_mn, t->Ehi, H, chain_rows, P, TC_DW_BK, 64, TC_BM / 64);
  make_map_chunked(&t->mElo_mn, t->Elo, H, chain_rows, P, TC_DW_BK, 64, TC_BM / 64);
  make_map_chunked(&t->mXhi_mn, t->Xhi, I, ring_rows, P, TC_DW_BK, 32, TC_DW2_BN / 2 / 32);
  make_map_chunked(&t->mXlo_mn, t->Xlo, I, ring_rows, P, TC_DW_BK, 32, TC_DW2_BN / 2 / 32);
  t->dw_splits = TC_DW_SPLITS;
  p->tc = t;
  p->x_planes_stale = 2; /* the ring may hold rows from before the planes existed */
  return t;
}

static RbPlanes
planes_of(rb_h16 *hi, rb_h16 *lo, int pitch, float scale, const float *scale_dev)
{
  RbPlanes pl;
  pl.hi = hi;
  pl.lo = lo;
  pl.pitch = pitch;
  pl.scale = scale;
  pl.scale_dev = scale_dev;
  return pl;
}

/* ======================================================================== */
/* operand planes                                                             */

/* Wih -> hi/lo planes, plain and transposed (32x32 tiles through smem) */
__global__ void __launch_bounds__(256)
k_split_weights(const float *__restrict__ W, int I, int H, rb_h16 *Whi, rb_h16 *Wlo, int wpitch,
    rb_h16 *WThi, rb_h16 *WTlo, int tpitch)
{
  __shared__ rb_h16 th[32][34], tl[32][34];
  int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    int y = y0 + r, x = x0 + tx;
    rb_h16 hi = 0, lo = 0;
    if (y < I && x < H) {
      rb_split_f16(W[(size_t)y * H + x] * RB_W_SCALE, hi, lo);
      Whi[(size_t)y * wpitch + x] = hi;
      Wlo[(size_t)y * wpitch + x] = lo;
    }
    th[r][tx] = hi;
    tl[r][tx] = lo;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int x = x0 + r, y = y0 + tx;
    if (x < H && y < I) {
      WThi[(size_t)x * tpitch + y] = th[tx][r];
      WTlo[(size_t)x * tpitch + y] = tl[tx][r];
    }
  }
}

/* the current ring row of every stream of the batch -> its planes */
__global__ void __launch_bounds__(256)
k_split_rows(RbView v, RbPlanes X)
{
  int s = v.slots[blockIdx.x];
  size_t row = (size_t)v.pos[s] * v.cap + s;
  const float *src = v.X + row * v.d.i_size;
  for (int i = threadIdx.x; i < v.d.i_size; i += blockDim.x) {
    rb_h16 hi, lo;
    rb_split_f16(src[i] * X.scale, hi, lo);
    X.hi[row * X.pitch + i] = hi;
    X.lo[row * X.pitch + i] = lo;
  }
}

/* every row of the ring -> its planes (after something other than the tensor
   engine's own forward pass rewrote ring rows: RbPool.x_planes_stale) */
__global__ void __launch_bounds__(256)
k_split_ring(const float *__restrict__ X, int I, size_t n_rows, RbPlanes P)
{
  const int per_row = I / 4;
  const size_t n = n_rows * per_row;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n;
       q += (size_t)gridDim.x * blockDim.x) {
    size_t row = q / per_row;
    int c = (int)(q - row * per_row) * 4;
    float4 x = *(const float4 *)(X + row * I + c);
    uint2 h, l;
    rb_split4(x, P.scale, h, l);
    *(uint2 *)(P.hi + row * P.pitch + c) = h;
    *(uint2 *)(P.lo + row * P.pitch + c) = l;
  }
}

/* Planes of E(0) for the whole batch, and the scale of this walk's error
   planes.  The top layer's soft clip bounds what the walk can hold: elements
   of E(0) are at most top_scaled (the sum of their magnitudes), and a later
   row's elements at most sqrt(max_sum) = sqrt(2 top + 1) as long as the walk
   goes on (recur-nn.c:318,387).  The largest bound over the batch's streams,
   rounded up to a power of two, is put just inside FP16's range; every block
   computes it from the streams' scalars by itself (same loads, same order,
   same result), block 0 publishes it for the kernels that read the planes. */
__global__ void __launch_bounds__(256)
k_e0_planes(RbView v, RbPlanes E, float *scale_out)
{
  __shared__ float s_max[8];
  float m = 0.0f;
  for (int j = threadIdx.x; j < v.n; j += blockDim.x) {
    const RbScalars *sc = v.sc + v.slots[j];
    if (sc->live)
      m = fmaxf(m, sc->top_scaled);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0)
    s_max[threadIdx.x >> 5] = m;
  __syncthreads();
  m = s_max[0];
#pragma unroll
  for (int w = 1; w < 8; w++)
    m = fmaxf(m, s_max[w]);
  const float bound = fmaxf(m, sqrtf(2.0f * m + 1.0f));
  int e;
  frexpf(bound, &e);                       /* bound < 2^e */
  const float scale = (bound < 3.0e38f) ? ldexpf(1.0f, 15 - e) : 1.0f;   /* NaN or inf: anything */
  if (blockIdx.x == 0 && threadIdx.x == 0)
    *scale_out = scale;
  const int s = v.slots[blockIdx.x];
  const float *e0 = v.E + (size_t)s * v.d.i_size;
  rb_h16 *hi = E.hi + (size_t)s * E.pitch, *lo = E.lo + (size_t)s * E.pitch;
  for (int i = threadIdx.x * 4; i < v.d.h_size; i += blockDim.x * 4) {
    uint2 h, l;
    rb_split4(*(const float4 *)(e0 + i), scale, h, l);
    *(uint2 *)(hi + i) = h;
    *(uint2 *)(lo + i) = l;
  }
}

/* After the walk: rows of E beyond a stream's executed depth must not reach
   the weight gradient (zero them), and a stream whose gradient is clipped
   (ih_scale != 1, recur-nn.c:393-402) has its rows rescaled and re-split. */
__global__ void __launch_bounds__(256)
k_finalize_rows(RbView v, RbPlanes E, const unsigned int *kmax_dev, unsigned int *zero,
    int zero_words)
{
  /* the other of the two barrier/counter areas is cleared here for the next
     walk, which saves that walk a memset in front of its kernel */
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < zero_words;
       i += gridDim.x * blockDim.x)
    zero[i] = 0u;
  const int s = v.slots[blockIdx.x];
  const RbScalars sc = v.sc[s];
  const int kmax = min((int)*kmax_dev, v.depth);
  const float scale = *E.scale_dev;
  /* rows this stream never reached, up to the deepest step any stream took
     (the weight gradient stops there) */
  for (int step = sc.n_steps; step < kmax; step++) {
    size_t off = ((size_t)step * v.cap + s) * E.pitch;
    for (int i = threadIdx.x * 8; i < v.d.h_size; i += blockDim.x * 8) {
      *(uint4 *)(E.hi + off + i) = make_uint4(0u, 0u, 0u, 0u);
      *(uint4 *)(E.lo + off + i) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (sc.ih_scale != 1.0f) {
    const float f = sc.ih_scale * scale;
    for (int step = 0; step < sc.n_steps; step++) {
      size_t off = ((size_t)step * v.cap + s) * E.pitch;
      const float *src = v.E + ((size_t)step * v.cap + s) * v.d.i_size;
      for (int i = threadIdx.x * 4; i < v.d.h_size; i += blockDim.x * 4) {
        uint2 h, l;
        rb_split4(*(const float4 *)(src + i), f, h, l);
        *(uint2 *)(E.hi + off + i) = h;
        *(uint2 *)(E.lo + off + i) = l;
      }
    }
  }
}

/* deepest BPTT step any stream of the batch executed */
__global__ void __launch_bounds__(256)
k_compute_kmax(RbView v, unsigned int *kmax_dev)
{
  __shared__ int s_max;
  if (threadIdx.x == 0)
    s_max = 0;
  __syncthreads();
  int m = 0;
  for (int j = threadIdx.x; j < v.n; j += blockDim.x)
    m = max(m, v.sc[v.slots[j]].n_steps);
  atomicMax(&s_max, m);
  __syncthreads();
  if (threadIdx.x == 0)
    *kmax_dev = (unsigned int)s_max;
}

/* ======================================================================== */
/* FWD and per-step CHAIN: C[128 x BN] tiles, both operands K-major.
 *
 * FWD  (BN 64): activation epilogue straight from tensor memory.
 * FWD split-K / per-step CHAIN (BN 128, split-K over blockIdx.z): few tiles
 * and a long K, so K is split and each CTA writes its raw partial tile; the
 * output kernel (FWD) or k_chain_finish_step (CHAIN) sums the partials on its
 * way in.  The per-step CHAIN is the fallback for shapes the persistent
 * kernel below does not take.                                               */

struct NtArgs {
  RbView v;
  int mode;        /* 0 FWD, 1 CHAIN, 2 FWD split-K (raw partial sums) */
  int k;           /* CHAIN: step */
  int use_noise;
  float *cpartial; /* split-K: [splits][cap][i_size rounded up to 32] */
  float inv_scale; /* 1 / (scale of A's planes * scale of B's planes) ... */
  const float *a_scale_dev; /* ... CHAIN: A's scale lives on the device */
};

template <int BN, int STAGES>
struct NtCfg {
  static constexpr int A_BYTES = TC_BM * TC_KB * 2;   /* one plane of the A tile */
  static constexpr int B_BYTES = BN * TC_KB * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 2 * BN; /* main and correction accumulators */
};

/* the three MMAs of one 64-wide K block of split operands: `main` takes
   hi*hi, `corr` the two cross terms (rb_split.cuh) */
__device__ __forceinline__ void
issue_block_f16(uint32_t tmem_main, uint32_t tmem_corr, uint32_t a_hi, uint32_t a_lo,
    uint32_t b_hi, uint32_t b_lo, uint32_t idesc, bool first)
{
#pragma unroll
  for (int kk = 0; kk < TC_KB / 16; kk++) {
    /* K-major, 128B swizzle: 8-row groups 1024 B apart; a K step of 16
       halves moves the start address by 32 bytes inside the swizzle row */
    uint64_t dah = umma_desc(a_hi + kk * 32, 16, 1024);
    uint64_t dal = umma_desc(a_lo + kk * 32, 16, 1024);
    uint64_t dbh = umma_desc(b_hi + kk * 32, 16, 1024);
    uint64_t dbl = umma_desc(b_lo + kk * 32, 16, 1024);
    const uint32_t acc = (first && kk == 0) ? 0u : 1u;
    umma_f16(tmem_main, dah, dbh, idesc, acc);
    umma_f16(tmem_corr, dah, dbl, idesc, acc);
    umma_f16(tmem_corr, dal, dbh, idesc, 1u);
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
k_tc_nt(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
    const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo,
    NtArgs g)
{
  using Cfg = NtCfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const RbView &v = g.v;
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *full = (uint64_t *)(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_ready = empty + STAGES;
  uint32_t *tmem_slot = (uint32_t *)(acc_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = v.d.i_size, H = v.d.h_size;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int K = (g.mode == 1) ? H : I;
  const int n_kb_total = (K + TC_KB - 1) / TC_KB;
  const int kb_per_split = (n_kb_total + gridDim.z - 1) / gridDim.z;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(n_kb_total, kb_begin + kb_per_split);
  const int n_kb = max(0, kb_end - kb_begin);

  /* CHAIN: a tile whose streams have all stopped has nothing to do */
  if (g.mode == 1) {
    int alive = 0;
    for (int r = lane; r < TC_BM; r += 32) {
      int m = m0 + r;
      if (m < v.n && v.sc[v.base + m].live)
        alive = 1;
    }
    if (!__any_sync(0xffffffffu, alive))
      return;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_ready, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mAhi);
    tma_prefetch_desc(&mAlo);
    tma_prefetch_desc(&mBhi);
    tma_prefetch_desc(&mBlo);
  }
  if (warp == 1)
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      /* row of the A operand in its ring: FWD reads the newest x row, CHAIN E[k] */
      int ring_row;
      if (g.mode != 1)
        ring_row = v.pos[v.base] * v.cap + v.base + m0;
      else
        ring_row = g.k * v.cap + v.base + m0;
      for (int it = 0; it < n_kb; it++) {
        int kb = kb_begin + it;
        int s = it % STAGES;
        uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t *st = smem + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        tma_load_2d(&mAhi, &full[s], st, kb * TC_KB, ring_row);
        tma_load_2d(&mAlo, &full[s], st + Cfg::A_BYTES, kb * TC_KB, ring_row);
        tma_load_2d(&mBhi, &full[s], st + 2 * Cfg::A_BYTES, kb * TC_KB, n0);
        tma_load_2d(&mBlo, &full[s], st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, kb * TC_KB, n0);
      }
    }
  }
  else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(TC_BM, BN, 0, 0);
      for (int it = 0; it < n_kb; it++) {
        int s = it % STAGES;
        uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        uint32_t a_hi = smem_u32(smem + s * Cfg::STAGE_BYTES);
        uint32_t a_lo = a_hi + Cfg::A_BYTES;
        uint32_t b_hi = a_hi + 2 * Cfg::A_BYTES;
        uint32_t b_lo = b_hi + Cfg::B_BYTES;
        issue_block_f16(tmem_base, tmem_base + BN, a_hi, a_lo, b_hi, b_lo, idesc, it == 0);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_ready);
    }
  }
  else {
    /* epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 == tile rows */
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < v.n;
    const int sidx = v.base + (row_ok ? m : 0);
    if (n_kb > 0) {
      mbar_wait(acc_ready, 0);
      tc_fence_after();
    }
    const float inv = g.a_scale_dev ? g.inv_scale / *g.a_scale_dev : g.inv_scale;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float acc[32], cor[32];
    if (g.mode == 0) {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        tmem_ld32_nowait(taddr + c, acc);
        tmem_ld32(taddr + BN + c, cor);
        int col0 = n0 + c;
        if (!row_ok || col0 >= H)
          continue;
        float *dst = v.Hd + (size_t)sidx * H + col0;
        const float *nz = v.noise + (size_t)sidx * H + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float o[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            int col = col0 + j + u;
            float h = fmaf(cor[j + u], RB_LO_UNGAIN, acc[j + u]) * inv;
            if (g.use_noise && col >= 1 && col < H)
              h += nz[j + u];
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
            o[u] = h;
          }
          if (col0 + j < H)
            *(float4 *)(dst + j) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    else {
      /* raw partial sums of this K split; masks and the rest happen row-wise
         in the kernel that sums the splits */
      const bool live = row_ok && (g.mode == 2 || v.sc[sidx].live != 0);
      const int cpitch = (I + 31) & ~31;
      const int n_cols = (g.mode == 2) ? H : I;
      float *dst = g.cpartial + ((size_t)blockIdx.z * v.cap + sidx) * cpitch;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (n_kb > 0) {
          tmem_ld32_nowait(taddr + c, acc);
          tmem_ld32(taddr + BN + c, cor);
#pragma unroll
          for (int j = 0; j < 32; j++)
            acc[j] = fmaf(cor[j], RB_LO_UNGAIN, acc[j]) * inv;
        }
        else {
#pragma unroll
          for (int j = 0; j < 32; j++)
            acc[j] = 0.0f;
        }
        int col0 = n0 + c;
        if (!live || col0 >= n_cols)
          continue;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (col0 + j + 8 <= n_cols)
            st_global_v8(dst + col0 + j, acc + j);
          else if (col0 + j < n_cols)
            __stcg((float4 *)(dst + col0 + j),
                make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

/* What the reference does with one stream's error row once its sum of
   squares is known (recur-nn.c:383-413): count the step, stop the walk on
   either threshold, and on the way out clip the gradient or adapt
   min_error_factor.  k is the step just executed. */
__device__ __forceinline__ void
chain_decide(RbScalars &sc, float es, int k, int depth)
{
  sc.err_sum = es;
  sc.cum_error += sqrtf(es);
  sc.n_steps = k + 1;
  const int t = depth - k;
  const bool stop = (es <= sc.min_sum || es > sc.max_sum);
  const bool last = (k == depth - 1);
  if (!stop && !last)
    return;
  sc.live = 0;
  const int t_left = stop ? t : 0;
  sc.t_left = t_left;
  const float ceiling = ERROR_GAIN_CEILING * sc.top_scaled;
  if (es > ceiling) {
    float halfmax = sc.max_sum;
    float x = es / halfmax;
    float fudge = (float)(0.99 + (double)(x * x / 100.0f));
    sc.ih_scale = (halfmax == 0.0f) ? es : 2.0f * x / (1.0f + x * x * fudge);
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
}

/* Four columns of E(k+1) from the summed products `a` and the ring row `xin`
   they are masked with (recur-nn.c:338-376): the value that enters the sum
   of squares, the bottom layer's accumulator, and what is stored as the next
   step's operand (bias and padding columns carry no error).  PLAIN is the
   common case (ReLU, no bottom layer), kept free of branches. */
template <bool PLAIN>
__device__ __forceinline__ float4
chain_mask4(const RbView &v, float4 a, float4 xin, int c, int s, float &sq)
{
  const int H = v.d.h_size, hs1 = v.d.hidden_size + 1;
  float av[4] = {a.x, a.y, a.z, a.w};
  float xi[4] = {xin.x, xin.y, xin.z, xin.w};
  float o[4];
#pragma unroll
  for (int u = 0; u < 4; u++) {
    float e;
    if (PLAIN) {
      e = (xi[u] != 0.0f) ? av[u] : 0.0f;
      sq = fmaf(e, e, sq);
    }
    else {
      e = 0.0f;
      float input = xi[u];
      if (input != 0.0f && (v.activation != RNN_RECLIP20 || input < 20.0f)) {
        e = av[u];
        if (v.activation == RNN_RESQRT)
          e /= 2.0f * (input + 1.0f);
        sq += e * e;
      }
      int col = c + u;
      if (v.CIE && col >= hs1 && col < hs1 + v.d.input_size)
        v.CIE[(size_t)s * v.bl_o + col - hs1] += e;
    }
    o[u] = e;
  }
  if ((c == 0) || (c + 4 > hs1 && c < H)) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      int col = c + u;
      if (col == 0 || (col >= hs1 && col < H))
        o[u] = 0.0f;
    }
  }
  return make_float4(o[0], o[1], o[2], o[3]);
}

/* The row-wise half of a BPTT step (recur-nn.c:338-389, 393-413), one block
   per stream: sum the split-K partials in a fixed order, mask by the input
   that fed each row, ReSQRT derivative, write E(k+1) with its planes,
   sum the squares, and decide whether this stream walks further.           */
__global__ void __launch_bounds__(256)
k_chain_finish_step(RbView v, int k, const float *__restrict__ cpartial, int splits, RbPlanes E)
{
  __shared__ float scratch[33];
  const int s = v.slots[blockIdx.x];
  RbScalars *scp = v.sc + s;
  RbScalars sc = *scp; /* one read up front; thread 0 writes the changes back */
  if (!sc.live)
    return;
  const int I = v.d.i_size;
  const float scale = *E.scale_dev;
  int p = v.pos[s] - k;
  if (p < 0)
    p += v.depth;
  const float *xk = v.X + ((size_t)p * v.cap + s) * I;
  const size_t eoff = ((size_t)(k + 1) * v.cap + s) * I;
  const size_t poff = ((size_t)(k + 1) * v.cap + s) * E.pitch;
  const int cpitch = (I + 31) & ~31;
  const size_t split_stride = (size_t)v.cap * cpitch;
  const float *part = cpartial + (size_t)s * cpitch;
  float sq = 0.0f;
  for (int c = threadIdx.x * 4; c < I; c += blockDim.x * 4) {
    float4 a = *(const float4 *)(part + c);
    for (int z = 1; z < splits; z++) {
      float4 b = *(const float4 *)(part + z * split_stride + c);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    float4 o = chain_mask4<false>(v, a, *(const float4 *)(xk + c), c, s, sq);
    *(float4 *)(v.E + eoff + c) = o;
    uint2 h, l;
    rb_split4(o, scale, h, l);
    *(uint2 *)(E.hi + poff + c) = h;
    *(uint2 *)(E.lo + poff + c) = l;
  }
  float es = block_sum_tc(sq, scratch);
  if (threadIdx.x != 0)
    return;
  chain_decide(sc, es, k, v.depth);
  *scp = sc;
}

/* ======================================================================== */
/* The whole BPTT walk as ONE persistent kernel on thread-block clusters.
 *
 * A BPTT step is a small dependent GEMM, E(k+1) = mask * (E(k) . Wih^T): at
 * 512 streams of a 1023-unit net a [512 x 1024] x [1024 x 1068] product that
 * the next step cannot start without.  The grid is (column tiles x K splits)
 * x stream tiles of 128, one CTA per SM, alive for all `depth` steps:
 *
 *  * the K splits of one output tile form a CLUSTER.  Each CTA's slice of
 *    Wih (128 rows x its K range, both planes: 128 KB at H1023) is loaded
 *    into shared memory ONCE and stays there for the whole walk; only the
 *    error rows stream through the TMA ring each step;
 *  * the split-K reduction never leaves the cluster, and nobody waits for a
 *    load from another SM: each CTA finishes one column slice of the tile.
 *    From tensor memory it keeps its own slice in registers and parks the
 *    slices of its peers in shared memory; once every CTA of the cluster has
 *    signalled (remote mbarrier arrive) that its K loop is over - the ring is
 *    free to be written - one bulk copy per peer (cp.async.bulk shared::cta ->
 *    shared::cluster) drops each parked slice into the peer's ring memory and
 *    counts its bytes on the peer's mbarrier.  The receiver sums the slices
 *    in rank order out of its OWN shared memory, masks, writes E(k+1) with
 *    its planes, and the row's partial sum of squares;
 *  * the streams of a tile of 128 are an independent chain, so only the CTAs
 *    that share blockIdx.y synchronise, ONCE per step (monotonic counter,
 *    arrive = red.release.gpu, wait = ld.acquire.gpu by the TMA thread, which
 *    issues the next step's loads the moment the counter is there);
 *  * the stop rule (recur-nn.c:383-413) needs a row's sum of squares over
 *    ALL columns, which no CTA has: every CTA of the group sums the per-CTA
 *    parts of the PREVIOUS step for all 128 streams (same loads, same order,
 *    same result everywhere) while that step's successor is already in the
 *    tensor pipe, so the decision costs the critical path nothing; the group
 *    leaves one (speculative) K loop after its last stream stopped.  Rows of
 *    a stream beyond its own stop are computed but never stored;
 *  * columns past the last tile (the input rows of a one-hot text net: 44 of
 *    1068) would cost a whole extra tile per stream group and more clusters
 *    than fit the GPU's GPCs; their few nonzero rows are dot products on two
 *    spare warps instead (recur-nn.c:347 skips the zero rows just so).
 *
 * Shared memory: W slice 128 KB | TMA ring 3 x 32 KB, which between a step's
 * K loop and the next step's loads holds the outgoing (48 KB) and incoming
 * (48 KB) slices of the exchange | barriers.  Tensor memory: two 128-column
 * accumulators.                                                             */

#define CH_THREADS 256
#define CH_BN 128
#define CH_WKB 4      /* K blocks of Wih resident per CTA */
#define CH_STAGES 3
#define CH_PLANE_BYTES (TC_BM * TC_KB * 2)        /* 16 KB: 128 rows x 64 halves */
#define CH_STAGE_BYTES (2 * CH_PLANE_BYTES)       /* hi and lo planes of one K block */
#define CH_W_BYTES (CH_WKB * 2 * CH_PLANE_BYTES)  /* 128 KB */
#define CH_RING_BYTES (CH_STAGES * CH_STAGE_BYTES)
#define CH_SMEM_BYTES (CH_W_BYTES + CH_RING_BYTES + 1024 + 1024)

__device__ __forceinline__ unsigned int
ld_acquire_gpu(const unsigned int *p)
{
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

/* A poll is an L2 round trip (~0.7 us): 2^22 of them are seconds, against the
   microseconds a healthy barrier takes.  A grid that is not fully resident
   can never complete the barrier; trap (the host then aborts with the launch
   failure) rather than hang the device. */
__device__ __forceinline__ void
wait_counter(const unsigned int *counter, unsigned int target)
{
  unsigned int spins = 0;
  while (ld_acquire_gpu(counter) < target) {
    if (++spins > (1u << 22))
      __trap();
  }
}

/* arrive on the mbarrier at the same shared-memory offset in CTA `rank` of
   the cluster (a pure signal: relaxed, nothing is published with it) */
__device__ __forceinline__ void
mbar_arrive_cluster(uint64_t *bar, uint32_t rank)
{
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];"
      ::"r"(dsmem_addr(bar, rank)) : "memory");
}

__device__ __forceinline__ void
mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

/* `bytes` of this CTA's shared memory -> the same-layout buffer `dst` in CTA
   `rank`, completion counted on that CTA's mbarrier */
__device__ __forceinline__ void
bulk_copy_to_peer(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint32_t rank)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dsmem_addr(dst, rank)), "r"(smem_u32(src)), "r"(bytes), "r"(dsmem_addr(bar, rank))
      : "memory");
}

struct ChainArgs {
  RbView v;
  RbPlanes E;           /* planes of the chain; scale on the device */
  float *sqpart;        /* [2][m tiles][128][CH_SQ_SLOTS] */
  unsigned int *sync;   /* [m tiles][CH_SYNC_STRIDE] barrier counters */
  unsigned int *kmax;   /* out: deepest step any stream executed */
  int splits;           /* K splits == cluster size */
  int nkb_total, kb_per;/* 64-wide K blocks in all, per split */
  int tail0;            /* first column left to the CUDA-core tail (i_size: none) */
  unsigned long long *dbg; /* TIMING instantiation only: [step][16] globaltimer stamps of CTA 0 */
};

__device__ __forceinline__ unsigned long long
globaltimer_ns(void)
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

/* measurement build of the kernel only (RECUR_B200_CHAIN_TIMING): the
   production instantiation compiles these away */
#define CH_STAMP(slot) do {                                             \
    if (TIMING && blockIdx.x == 0 && blockIdx.y == 0 && k < 64)         \
      g.dbg[k * 32 + (slot)] = globaltimer_ns();                        \
  } while (0)

__device__ __forceinline__ void
named_bar_sync(int id, int count)
{
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

/* one stream's sum of squares of the step decided on: the per-CTA parts in
   slot order (the same in every CTA of the group) */
__device__ __forceinline__ float
sum_sq_parts(const float *sp, int n_slots)
{
  /* eight loads in flight at a time: summed one after the other they would
     each wait out an L2 round trip */
  float es = 0.0f;
#pragma unroll 1
  for (int q0 = 0; 4 * q0 < n_slots; q0 += 8) {
    float4 t[8];
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (4 * (q0 + q) < n_slots)
        t[q] = __ldcg((const float4 *)(sp + 4 * (q0 + q)));
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int c = 4 * (q0 + q);
      if (c < n_slots) es += t[q].x;
      if (c + 1 < n_slots) es += t[q].y;
      if (c + 2 < n_slots) es += t[q].z;
      if (c + 3 < n_slots) es += t[q].w;
    }
  }
  return es;
}


template <int SPLITS, bool TIMING>
__global__ void __launch_bounds__(CH_THREADS, 1)
k_tc_chain_persistent(const __grid_constant__ CUtensorMap mEhi,
    const __grid_constant__ CUtensorMap mElo, const __grid_constant__ CUtensorMap mWhi,
    const __grid_constant__ CUtensorMap mWlo, ChainArgs g)
{
  constexpr int CPR = CH_BN / SPLITS;     /* columns of the tile this CTA finishes: its slice */
  constexpr int SEG = CPR / 32;           /* 128-byte segments per slice row */
  constexpr int SLICE_BYTES = TC_BM * CPR * 4;
  constexpr int LPR = CPR / 4;            /* row phase: threads per row, one float4 each */
  constexpr int RPI = CH_THREADS / LPR;   /* rows per pass over the slice */
  constexpr int NG = TC_BM / RPI;         /* passes == float4 groups per thread */
  static_assert(2 * (SPLITS - 1) * SLICE_BYTES <= CH_RING_BYTES, "exchange buffers fit the ring");
  static_assert(SLICE_BYTES <= CH_STAGE_BYTES, "the own slice fits the last Wih slot");
  extern __shared__ uint8_t smem_raw[];
  const RbView &v = g.v;
  /* the dynamic shared window starts at the same offset in every CTA, so the
     rounded-up base does too: buffers and barriers have one address cluster-wide */
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *w_smem = smem;
  uint8_t *a_smem = smem + CH_W_BYTES;
  uint8_t *out_buf = a_smem;                                  /* [SPLITS-1] slices for the peers */
  uint8_t *in_buf = a_smem + (SPLITS - 1) * SLICE_BYTES;      /* [SPLITS-1] slices from the peers */
  /* This CTA's own slice is parked at the end of the Wih area.  When all
     CH_WKB slots hold weights, the planes under it (the last block's lo
     plane, at two splits the whole block) are fetched again at the start of
     the next step - 16 KB with no dependence on anything, long there before
     the last block's MMAs want them. */
  uint8_t *own_buf = w_smem + CH_W_BYTES - SLICE_BYTES;
  uint64_t *full = (uint64_t *)(a_smem + CH_RING_BYTES);
  uint64_t *empty = full + CH_STAGES;
  uint64_t *acc_ready = empty + CH_STAGES;
  uint64_t *w_ready = acc_ready + 1;
  uint64_t *step_go = w_ready + 1;
  uint64_t *k_done = step_go + 1;   /* every CTA of the cluster is through its K loop */
  uint64_t *data_in = k_done + 1;   /* the peers' slices have landed */
  uint64_t *w_again = data_in + 1;  /* the weights the own slice sat on are back */
  uint64_t *data_in2 = w_again + 1; /* ... and the second half of their rows */
  uint32_t *tmem_slot = (uint32_t *)(data_in2 + 1);
  int *s_any = (int *)(tmem_slot + 1);      /* [2] a stream walks on, by step parity */
  int *s_kmax = s_any + 2;
  uint8_t *s_live = (uint8_t *)(s_kmax + 1); /* [128] */

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = v.d.i_size, H = v.d.h_size;
  const uint32_t rank = cluster_ctarank();
  const int n0 = ((int)blockIdx.x / SPLITS) * CH_BN, m0 = blockIdx.y * TC_BM;
  const int grp = blockIdx.y;
  const int grp_ctas = gridDim.x, grp_cta = blockIdx.x;
  unsigned int *gsync = g.sync + (size_t)grp * CH_SYNC_STRIDE;
  const int kb_begin = (int)rank * g.kb_per;
  const int n_kb = min(g.kb_per, g.nkb_total - kb_begin);
  const bool has_tail = g.tail0 < I;
  const int n_slots = grp_ctas + (has_tail ? 1 : 0);
  const bool w_refetch = n_kb == CH_WKB; /* the own slice's parking place holds weights */
  float *sq_grp = g.sqpart + (size_t)grp * TC_BM * CH_SQ_SLOTS;
  const size_t sq_par = (size_t)gridDim.y * TC_BM * CH_SQ_SLOTS; /* floats between the two parities */
  const int depth = v.depth;
  const int pos0 = v.pos[v.base]; /* the batch advances in lockstep */
  const bool plain = (v.activation == RNN_RELU && v.CIE == NULL);

  if (threadIdx.x == 0) {
    for (int s = 0; s < CH_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_ready, 1);
    mbar_init(w_ready, 1);
    mbar_init(step_go, 1);
    mbar_init(k_done, SPLITS);
    mbar_init(data_in, 1);
    mbar_init(data_in2, 1);
    mbar_init(w_again, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mEhi);
    tma_prefetch_desc(&mElo);
    tma_prefetch_desc(&mWhi);
    tma_prefetch_desc(&mWlo);
    s_any[0] = 0;
    s_any[1] = 0;
    *s_kmax = 0;
  }
  if (warp == 1)
    tmem_alloc(tmem_slot, 2 * CH_BN);

  /* the decision threads (warps 2..5) keep one stream's scalars each */
  const int di = threadIdx.x - 64;
  const bool decider = di >= 0 && di < TC_BM;
  const bool d_valid = decider && m0 + di < v.n;
  RbScalars dsc;
  dsc.live = 0;
  dsc.n_steps = 0;
  if (d_valid)
    dsc = v.sc[v.base + m0 + di];
  if (decider)
    s_live[di] = d_valid && dsc.live;
  const bool owner = grp_cta == 0; /* this CTA writes the group's scalars back */

  tc_fence_before();
  __syncthreads();
  cluster_sync_all(); /* every CTA's barriers exist before a peer signals them */
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const float e_scale = *g.E.scale_dev;
  const float inv_scale = 1.0f / (e_scale * RB_W_SCALE);

  /* row phase: one float4 of this CTA's slice in every RPI-th row, a row's
     threads side by side (whole 128-byte runs to global memory) */
  const int r_row0 = threadIdx.x / LPR;
  const int r_j = threadIdx.x % LPR;               /* float4 index within the slice row */
  const int r_col = n0 + (int)rank * CPR + r_j * 4;

  unsigned int it = 0; /* ring iterations so far (producer and MMA keep equal counts) */
  bool finished = false;

  for (int k = 0; k < depth; k++) {
    const int par = k & 1;
    if (threadIdx.x == 0)
      CH_STAMP(0);
    /* the ring rows that mask this step's errors travel while the GEMM runs */
    float4 xq[NG];
    {
      int p = pos0 - k;
      if (p < 0)
        p += depth;
#pragma unroll
      for (int q = 0; q < NG; q++) {
        const int row = r_row0 + q * RPI;
        xq[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + row < v.n && r_col < I)
          xq[q] = __ldg((const float4 *)(v.X + ((size_t)p * v.cap + v.base + m0 + row) * I + r_col));
      }
    }

    if (warp == 0) {
      if (lane == 0) {
        if (k == 0) {
          mbar_expect_tx(w_ready, (uint32_t)n_kb * CH_STAGE_BYTES);
          for (int j = 0; j < n_kb; j++) {
            tma_load_2d(&mWhi, w_ready, w_smem + j * CH_STAGE_BYTES, (kb_begin + j) * TC_KB, n0);
            tma_load_2d(&mWlo, w_ready, w_smem + j * CH_STAGE_BYTES + CH_PLANE_BYTES,
                (kb_begin + j) * TC_KB, n0);
          }
        }
        else {
          if (w_refetch) {
            /* the weights under the own slice's parking place */
            const int j = CH_WKB - 1;
            if (SLICE_BYTES > CH_PLANE_BYTES) {
              mbar_expect_tx(w_again, CH_STAGE_BYTES);
              tma_load_2d(&mWhi, w_again, w_smem + j * CH_STAGE_BYTES, (kb_begin + j) * TC_KB, n0);
            }
            else
              mbar_expect_tx(w_again, CH_PLANE_BYTES);
            tma_load_2d(&mWlo, w_again, w_smem + j * CH_STAGE_BYTES + CH_PLANE_BYTES,
                (kb_begin + j) * TC_KB, n0);
          }
          /* every CTA of the group has written its part of E(k) - and, being
             there, has taken delivery of the slices this CTA sent it */
          wait_counter(gsync, (unsigned int)k * grp_ctas);
          mbar_arrive(step_go);
        }
        CH_STAMP(1);
        const int erow = k * v.cap + v.base + m0;
        for (int j = 0; j < n_kb; j++, it++) {
          int s = it % CH_STAGES;
          uint32_t ph = (it / CH_STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t *st = a_smem + s * CH_STAGE_BYTES;
          mbar_expect_tx(&full[s], CH_STAGE_BYTES);
          tma_load_2d(&mEhi, &full[s], st, (kb_begin + j) * TC_KB, erow);
          tma_load_2d(&mElo, &full[s], st + CH_PLANE_BYTES, (kb_begin + j) * TC_KB, erow);
        }
      }
    }
    else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_f16(TC_BM, CH_BN, 0, 0);
        if (k == 0)
          mbar_wait(w_ready, 0);
        for (int j = 0; j < n_kb; j++, it++) {
          int s = it % CH_STAGES;
          uint32_t ph = (it / CH_STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (j == 0)
            CH_STAMP(2);
          if (w_refetch && k > 0 && j == CH_WKB - 1) {
            mbar_wait(w_again, (k - 1) & 1);
            tc_fence_after();
          }
          uint32_t a_hi = smem_u32(a_smem + s * CH_STAGE_BYTES);
          uint32_t b_hi = smem_u32(w_smem + j * CH_STAGE_BYTES);
          issue_block_f16(tmem_base, tmem_base + CH_BN, a_hi, a_hi + CH_PLANE_BYTES, b_hi,
              b_hi + CH_PLANE_BYTES, idesc, j == 0);
          umma_commit(&empty[s]);
        }
        umma_commit(acc_ready);
        CH_STAMP(3);
        /* when the MMAs are through, this CTA's ring is free: tell the
           cluster (itself included).  Nothing of this thread's own needs
           publishing with it, hence relaxed. */
        mbar_wait(acc_ready, par);
#pragma unroll
        for (int z = 0; z < SPLITS; z++)
          mbar_arrive_cluster(k_done, (uint32_t)z);
      }
    }
    else if (warp < 6) {
      /* ---- the previous step's verdict, for every stream of the group ---- */
      if (k > 0) {
        mbar_wait(step_go, (k - 1) & 1);
        bool on = false;
        if (d_valid && dsc.live) {
          float es = sum_sq_parts(sq_grp + (size_t)((k - 1) & 1) * sq_par +
              (size_t)di * CH_SQ_SLOTS, n_slots);
          chain_decide(dsc, es, k - 1, depth);
          on = dsc.live != 0;
          if (!on) {
            s_live[di] = 0;
            if (owner)
              v.sc[v.base + m0 + di] = dsc;
          }
        }
        if (__any_sync(0xffffffffu, on) && lane == 0)
          s_any[par] = 1;
        if (threadIdx.x == 64)
          CH_STAMP(7);
      }
      /* ---- this step's partial tile: the peers' slices leave tensor memory ---- */
      const int q = warp & 3;
      const int row = q * 32 + lane;
      mbar_wait(acc_ready, par);
      tc_fence_after();
      named_bar_sync(1, 128); /* the four warps' verdicts are in s_any */
      const bool go_on = !(k > 0 && !s_any[par]);
      if (go_on) {
        if (threadIdx.x == 64)
          CH_STAMP(4);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        float acc[2][16], cor[2][16];
        tmem_ld16_nowait(taddr, acc[0]);
        tmem_ld16_nowait(taddr + CH_BN, cor[0]);
#pragma unroll
        for (int ci = 0; ci < CH_BN / 16; ci++) {
          tmem_wait_ld();
          if (ci + 1 < CH_BN / 16) {
            tmem_ld16_nowait(taddr + (ci + 1) * 16, acc[(ci + 1) & 1]);
            tmem_ld16_nowait(taddr + CH_BN + (ci + 1) * 16, cor[(ci + 1) & 1]);
          }
          const float *av = acc[ci & 1], *cv = cor[ci & 1];
          const int z = ci / (2 * SEG); /* the rank that finishes these 16 columns */
          /* a slice: rows of CPR floats, the eight float4 of a 128-byte
             segment XOR-swizzled by the row so that a warp's stores (one row
             per lane) spread over the banks.  Outgoing slots: peers in rank
             order, this CTA left out. */
          uint8_t *dst = (z == (int)rank ? own_buf
                  : out_buf + (z - (z > (int)rank ? 1 : 0)) * SLICE_BYTES) +
              (size_t)row * (CPR * 4);
          const int jb = (ci % (2 * SEG)) * 4; /* first float4 index within the slice row */
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int jj = jb + j;
            *(float4 *)(dst + (jj >> 3) * 128 + (((jj & 7) ^ (row & 7)) << 4)) = make_float4(
                fmaf(cv[4 * j], RB_LO_UNGAIN, av[4 * j]),
                fmaf(cv[4 * j + 1], RB_LO_UNGAIN, av[4 * j + 1]),
                fmaf(cv[4 * j + 2], RB_LO_UNGAIN, av[4 * j + 2]),
                fmaf(cv[4 * j + 3], RB_LO_UNGAIN, av[4 * j + 3]));
          }
        }
        tc_fence_before();
        if (threadIdx.x == 64)
          CH_STAMP(14);
        /* the outgoing slices were written through the generic proxy, the bulk
           copies read them through the async proxy */
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        named_bar_sync(1, 128);
        if (threadIdx.x == 64) {
          CH_STAMP(5);
          mbar_expect_tx(data_in, (SPLITS - 1) * (SLICE_BYTES / 2));
          mbar_expect_tx(data_in2, (SPLITS - 1) * (SLICE_BYTES / 2));
          /* no peer may still be reading its ring with the tensor core */
          mbar_wait_cluster(k_done, par);
          CH_STAMP(8);
          /* Rows 0..63 of every slice first, then rows 64..127: the receivers
             start on the first half while the second is on its way.  (All CTAs
             send to rank 0 first, so the higher ranks get theirs later - by a
             whole slice per rank when slices went out in one piece; starting
             every CTA with its right-hand neighbour instead makes EVERY rank
             wait for the last position and was measured 1 us slower.) */
#pragma unroll
          for (int half = 0; half < 2; half++) {
#pragma unroll
            for (int z = 0; z < SPLITS; z++) {
              if (z == (int)rank)
                continue;
              /* at the receiver: senders in rank order, the receiver left out */
              const int out_slot = z - (z > (int)rank ? 1 : 0);
              const int in_slot = (int)rank - ((int)rank > z ? 1 : 0);
              bulk_copy_to_peer(in_buf + in_slot * SLICE_BYTES + half * (SLICE_BYTES / 2),
                  out_buf + out_slot * SLICE_BYTES + half * (SLICE_BYTES / 2), SLICE_BYTES / 2,
                  half ? data_in2 : data_in, (uint32_t)z);
            }
          }
        }
      }
    }
    else if (has_tail) {
      /* ---- columns past the last tile: the nonzero rows among them as dot
         products over E(k) in FP32, two warps, this CTA's share of the
         group's streams ---- */
      if (k > 0)
        mbar_wait(step_go, (k - 1) & 1);
      /* half a warp per stream, so that this CTA's (usually four) streams
         wait for their loads together, not one after the other */
      const int spc = (TC_BM + grp_ctas - 1) / grp_ctas;
      const int hl = lane & 15;
      const unsigned int hmask = (lane & 16) ? 0xffff0000u : 0x0000ffffu;
      const int hs1 = v.d.hidden_size + 1;
      for (int t = (warp - 6) * 2 + (lane >> 4); t < spc; t += 4) {
        const int i = grp_cta * spc + t;
        if (i >= TC_BM || m0 + i >= v.n || !s_live[i])
          continue; /* uniform over the half warp */
        const int s = v.base + m0 + i;
        int p = pos0 - k;
        if (p < 0)
          p += depth;
        const float *xk = v.X + ((size_t)p * v.cap + s) * I;
        const float *ek = v.E + ((size_t)k * v.cap + s) * I;
        float *en = v.E + ((size_t)(k + 1) * v.cap + s) * I;
        float sq = 0.0f;
        for (int c00 = g.tail0; c00 < I; c00 += 64) {
          /* the ring row's entries of up to 64 columns first, in one go */
          float xs[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int y = c00 + 16 * u + hl;
            xs[u] = (y < I) ? __ldg(xk + y) : 0.0f;
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int c0 = c00 + 16 * u;
            if (c0 >= I)
              break;
            const int y = c0 + hl;
            const float x = xs[u];
            const bool act = x != 0.0f && (v.activation != RNN_RECLIP20 || x < 20.0f);
            unsigned int todo = (__ballot_sync(hmask, act) >> (lane & 16)) & 0xffffu;
            float mine = 0.0f;
            while (todo) {
              const int b = __ffs(todo) - 1;
              todo &= todo - 1;
              const float *wrow = v.Wih + (size_t)(c0 + b) * H;
              /* eight loads of each operand in flight before the first is
                 used (h_size <= 1024 here: two rounds) */
              float dot = 0.0f;
              for (int cb = hl * 4; cb < H; cb += 512) {
                float4 w4[8], e4[8];
#pragma unroll
                for (int u2 = 0; u2 < 8; u2++) {
                  const int c = cb + 64 * u2;
                  if (c < H) {
                    w4[u2] = __ldg((const float4 *)(wrow + c));
                    e4[u2] = __ldcg((const float4 *)(ek + c));
                  }
                }
#pragma unroll
                for (int u2 = 0; u2 < 8; u2++) {
                  if (cb + 64 * u2 < H) {
                    dot = fmaf(w4[u2].x, e4[u2].x, dot);
                    dot = fmaf(w4[u2].y, e4[u2].y, dot);
                    dot = fmaf(w4[u2].z, e4[u2].z, dot);
                    dot = fmaf(w4[u2].w, e4[u2].w, dot);
                  }
                }
              }
#pragma unroll
              for (int o = 8; o > 0; o >>= 1)
                dot += __shfl_xor_sync(hmask, dot, o);
              if (hl == b)
                mine = dot;
            }
            if (y < I) {
              float e = mine;
              if (act) {
                if (v.activation == RNN_RESQRT)
                  e /= 2.0f * (x + 1.0f);
                sq = fmaf(e, e, sq);
                if (v.CIE && y >= hs1 && y < hs1 + v.d.input_size)
                  v.CIE[(size_t)s * v.bl_o + y - hs1] += e;
              }
              __stcg(en + y, (y >= hs1 && y < H) ? 0.0f : e);
            }
          }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1)
          sq += __shfl_xor_sync(hmask, sq, o);
        if (hl == 0)
          __stcg(sq_grp + (size_t)par * sq_par + (size_t)i * CH_SQ_SLOTS + grp_ctas, sq);
      }
      if (threadIdx.x == 192)
        CH_STAMP(6);
    }
    __syncthreads(); /* the verdicts are in for every warp */
    if (k > 0 && !s_any[par]) {
      finished = true; /* no stream of the group walks on (the same in all its CTAs) */
      break;
    }

    /* ---- E(k+1): this CTA's slice of the tile, summed over the K splits:
       the peers' parts have landed in this CTA's own shared memory ---- */
    /* (the rows' first half arrives, and is worked on, while the second is
       still in flight) */
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      constexpr int NH = NG / 2;
      mbar_wait(hh ? data_in2 : data_in, par);
      if (threadIdx.x == 0 && hh == 1)
        CH_STAMP(10);
      float4 a[NH];
      {
        float4 pz[SPLITS][NH];
#pragma unroll
        for (int z = 0; z < SPLITS; z++) {
          const uint8_t *src = (z == (int)rank ? own_buf
                  : in_buf + (z - (z > (int)rank ? 1 : 0)) * SLICE_BYTES) + (r_j >> 3) * 128;
#pragma unroll
          for (int q = 0; q < NH; q++) {
            const int row = r_row0 + (hh * NH + q) * RPI;
            pz[z][q] = *(const float4 *)(src + (size_t)row * (CPR * 4) +
                (((r_j & 7) ^ (row & 7)) << 4));
          }
        }
#pragma unroll
        for (int q = 0; q < NH; q++) {
          a[q] = pz[0][q];
#pragma unroll
          for (int z = 1; z < SPLITS; z++) {
            a[q].x += pz[z][q].x; a[q].y += pz[z][q].y;
            a[q].z += pz[z][q].z; a[q].w += pz[z][q].w;
          }
          a[q].x *= inv_scale; a[q].y *= inv_scale; a[q].z *= inv_scale; a[q].w *= inv_scale;
        }
      }
      if (TIMING && threadIdx.x == 0 && hh == 1 && a[0].x != 123.456f && a[NH - 1].w != 123.456f)
        CH_STAMP(15);
#pragma unroll
      for (int q = 0; q < NH; q++) {
        const int qq = hh * NH + q;
        const int row = r_row0 + qq * RPI;
        const bool live = m0 + row < v.n && s_live[row];
        const int rs = v.base + (m0 + row < v.n ? m0 + row : 0);
        float sq = 0.0f;
        if (live && r_col < I) {
          float4 o = plain ? chain_mask4<true>(v, a[q], xq[qq], r_col, rs, sq)
                           : chain_mask4<false>(v, a[q], xq[qq], r_col, rs, sq);
          if (TIMING && threadIdx.x == 0 && qq == 0 && o.x != 123.456f && sq != 123.456f)
            CH_STAMP(16);
          uint2 h, l;
          rb_split4(o, e_scale, h, l);
          if (TIMING && threadIdx.x == 0 && qq == 0 && h.x != 12345u && l.y != 12345u)
            CH_STAMP(17);
          __stcg((float4 *)(v.E + ((size_t)(k + 1) * v.cap + rs) * I + r_col), o);
          const size_t poff = ((size_t)(k + 1) * v.cap + rs) * g.E.pitch + r_col;
          __stcg((uint2 *)(g.E.hi + poff), h);
          __stcg((uint2 *)(g.E.lo + poff), l);
          if (TIMING && threadIdx.x == 0 && qq == 0)
            CH_STAMP(18);
        }
        /* the row's LPR threads sit side by side in a warp */
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1)
          sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (live && r_j == 0)
          __stcg(sq_grp + (size_t)par * sq_par + (size_t)row * CH_SQ_SLOTS + grp_cta, sq);
        if (TIMING && threadIdx.x == 0 && qq == 0)
          CH_STAMP(19);
      }
    }
    if (threadIdx.x == 0) {
      s_any[par ^ 1] = 0; /* the next step's verdicts start from "nobody" */
      CH_STAMP(11);
    }
    /* E(k+1)'s planes were written through the generic proxy; the next step's
       TMA reads them through the async proxy */
    asm volatile("fence.proxy.async.global;" ::: "memory");
    if (threadIdx.x == 0)
      CH_STAMP(12);
    __syncthreads();
    if (threadIdx.x == 0) {
      CH_STAMP(13);
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(gsync), "r"(1u) : "memory");
      CH_STAMP(9);
      if (TIMING && k == 6) /* every CTA's arrival in one step: who is late */
        g.dbg[64 * 32 + blockIdx.y * gridDim.x + blockIdx.x] = globaltimer_ns();
    }
  }

  if (!finished) {
    /* all `depth` steps ran: the last step's verdicts are still out */
    if (threadIdx.x == 0)
      wait_counter(gsync, (unsigned int)depth * grp_ctas);
    __syncthreads();
    if (d_valid && dsc.live) {
      float es = sum_sq_parts(sq_grp + (size_t)((depth - 1) & 1) * sq_par +
          (size_t)di * CH_SQ_SLOTS, n_slots);
      chain_decide(dsc, es, depth - 1, depth);
      if (owner)
        v.sc[v.base + m0 + di] = dsc;
    }
  }
  /* (on the early exit the speculative K loop's accumulator was waited for
     like any other before the verdict was looked at: no MMA is in flight, and
     no slice was exchanged in that step) */
  if (owner && d_valid)
    atomicMax(s_kmax, dsc.n_steps);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all(); /* nobody leaves while a peer may still signal its barriers */
  if (owner && threadIdx.x == 0)
    atomicMax(g.kmax, (unsigned int)*s_kmax);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * CH_BN);
  }
}

/* ======================================================================== */
/* DW on CTA pairs: delta^T tile [256 h x 192 i] += E^T . X over (step, stream)
 * rows, both operands MN-major, 32 ring rows per stage, split-K across
 * blockIdx.z.
 *
 * A single SM takes in 64 B per clock from L2; a 128-row tile of this
 * contraction wants more than that per MMA cycle.  Two SMs of a TPC issuing
 * ONE MMA (tcgen05 cta_group::2, M = 256) each keep half of the N operand, so
 * per SM and 32 rows only 16 KB (its 128 M columns, both planes) + 12 KB (half
 * of 192 N columns) arrive and the tensor pipe, not the port, sets the pace.
 * M runs over the error columns h (1024 = 4 pairs of 128, no padding), N over
 * the input columns i (tiles of 192), D[h][i] = sum_rows E[row][h] X[row][i].
 * MN-major FP16 operands: the error planes come in 64-column chunks (128-byte
 * swizzle), the ring planes in 32-column chunks (64-byte swizzle: 96 = 3
 * chunks per CTA); one 3-D TMA per plane fetches all chunks of 32 ring rows.
 * The epilogue writes partial[z][i][h]: lanes are h, so every store is a full
 * line.
 *
 * Protocol (as CUTLASS's 2-SM pipelines): both CTAs' TMA loads count their
 * bytes on the leader's `full` barrier, on which only the leader's producer
 * arrives (expecting both halves); the leader's MMA thread releases a stage in
 * both CTAs with a multicast commit; the accumulator-ready commit is multicast
 * too, and each CTA's epilogue drains its own TMEM half.                     */

struct DwArgs {
  RbView v;
  float *partial; /* [splits][i_size][h_size] */
  const unsigned int *kmax; /* deepest step any stream executed: rows beyond contribute nothing */
  const float *e_scale_dev; /* scale of the error planes */
};

#define TC_DW2_STAGES 7
#define DW2_A_CHUNK (TC_DW_BK * 128)                         /* 64 columns x 32 rows: 4 KB */
#define DW2_B_CHUNK (TC_DW_BK * 64)                          /* 32 columns x 32 rows: 2 KB */
#define DW2_A_BYTES ((TC_BM / 64) * DW2_A_CHUNK)             /* 8 KB per plane */
#define DW2_B_BYTES ((TC_DW2_BN / 2 / 32) * DW2_B_CHUNK)     /* 6 KB per plane, this CTA's half */
#define DW2_STAGE_BYTES (2 * DW2_A_BYTES + 2 * DW2_B_BYTES)  /* 28 KB */
#define DW2_SMEM_BYTES (TC_DW2_STAGES * DW2_STAGE_BYTES + 1024 + 256)
#define DW2_TMEM_COLS 512
#define DW2_CORR_COL 256

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
k_tc_dw_pair(const __grid_constant__ CUtensorMap mEhi, const __grid_constant__ CUtensorMap mElo,
    const __grid_constant__ CUtensorMap mXhi, const __grid_constant__ CUtensorMap mXlo,
    DwArgs g)
{
  extern __shared__ uint8_t smem_raw[];
  const RbView &v = g.v;
  /* the dynamic shared window starts at the same offset in both CTAs, so the
     rounded-up base does too: descriptors and barrier offsets are valid for
     the pair */
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *full = (uint64_t *)(smem + TC_DW2_STAGES * DW2_STAGE_BYTES);
  uint64_t *empty = full + TC_DW2_STAGES;
  uint64_t *acc_ready = empty + TC_DW2_STAGES;
  uint32_t *tmem_slot = (uint32_t *)(acc_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = v.d.i_size, H = v.d.h_size;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int h0 = blockIdx.x * TC_BM;                      /* this CTA's M rows (blockIdx.x = 2 * pair + rank) */
  const int i0 = blockIdx.y * TC_DW2_BN;                  /* the pair's N columns */
  const int i_half = i0 + (int)rank * (TC_DW2_BN / 2);    /* the half this CTA loads */
  const int kb_per_step = v.n / TC_DW_BK;
  const int n_steps_max = min((int)*g.kmax, v.depth);
  const int n_kb_total = n_steps_max * kb_per_step;
  const int kb_per_split = (n_kb_total + gridDim.z - 1) / gridDim.z;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(n_kb_total, kb_begin + kb_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_DW2_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_ready, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mEhi);
    tma_prefetch_desc(&mElo);
    tma_prefetch_desc(&mXhi);
    tma_prefetch_desc(&mXlo);
  }
  if (warp == 1)
    tmem_alloc_pair(tmem_slot, DW2_TMEM_COLS);
  tc_fence_before();
  __syncthreads();    /* the allocation's address is in shared memory for every warp */
  cluster_sync_all(); /* both CTAs' barriers exist before anyone signals them */
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int pos = v.pos[v.base];
      for (int kb = kb_begin, it = 0; kb < kb_end; kb++, it++) {
        int s = it % TC_DW2_STAGES;
        uint32_t ph = (it / TC_DW2_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        int step = kb / kb_per_step;
        int b0 = (kb - step * kb_per_step) * TC_DW_BK;
        int slot = pos - step;
        if (slot < 0)
          slot += v.depth;
        int xrow = slot * v.cap + v.base + b0;
        int erow = step * v.cap + v.base + b0;
        uint8_t *st = smem + s * DW2_STAGE_BYTES;
        if (leader)
          mbar_expect_tx(&full[s], 2 * DW2_STAGE_BYTES);
        tma_load_3d_pair(&mEhi, &full[s], st, 0, erow, h0 / 64);
        tma_load_3d_pair(&mElo, &full[s], st + DW2_A_BYTES, 0, erow, h0 / 64);
        tma_load_3d_pair(&mXhi, &full[s], st + 2 * DW2_A_BYTES, 0, xrow, i_half / 32);
        tma_load_3d_pair(&mXlo, &full[s], st + 2 * DW2_A_BYTES + DW2_B_BYTES, 0, xrow,
            i_half / 32);
      }
    }
  }
  else if (warp == 1) {
    if (lane == 0 && leader) {
      const uint32_t idesc = umma_idesc_f16(2 * TC_BM, TC_DW2_BN, 1, 1);
      const uint32_t t_main = tmem_base, t_corr = tmem_base + DW2_CORR_COL;
      for (int kb = kb_begin, it = 0; kb < kb_end; kb++, it++) {
        int s = it % TC_DW2_STAGES;
        uint32_t ph = (it / TC_DW2_STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        uint32_t a_hi = smem_u32(smem + s * DW2_STAGE_BYTES);
        uint32_t a_lo = a_hi + DW2_A_BYTES;
        uint32_t b_hi = a_hi + 2 * DW2_A_BYTES;
        uint32_t b_lo = b_hi + DW2_B_BYTES;
#pragma unroll
        for (int kk = 0; kk < TC_DW_BK / 16; kk++) {
          /* MN-major: each K row is one swizzle row of 64 (A) or 32 (B) halves
             along M/N; chunks along M/N are one TMA box apart (leading
             offset), groups of 8 K rows 1024 B (A) / 512 B (B) apart (stride
             offset); one MMA consumes 16 K rows */
          uint64_t dah = umma_desc(a_hi + kk * 2048, DW2_A_CHUNK, 1024, UMMA_SW128);
          uint64_t dal = umma_desc(a_lo + kk * 2048, DW2_A_CHUNK, 1024, UMMA_SW128);
          uint64_t dbh = umma_desc(b_hi + kk * 1024, DW2_B_CHUNK, 512, UMMA_SW64);
          uint64_t dbl = umma_desc(b_lo + kk * 1024, DW2_B_CHUNK, 512, UMMA_SW64);
          const uint32_t acc = (it | kk) ? 1u : 0u;
          umma_f16_pair(t_main, dah, dbh, idesc, acc);
          umma_f16_pair(t_corr, dah, dbl, idesc, acc);
          umma_f16_pair(t_corr, dal, dbh, idesc, 1u);
        }
        umma_commit_pair(&empty[s]);
      }
      umma_commit_pair(acc_ready);
    }
  }
  else {
    const int q = warp & 3;
    const int h = h0 + q * 32 + lane;
    const bool any = kb_end > kb_begin;
    if (any) {
      mbar_wait(acc_ready, 0);
      tc_fence_after();
    }
    const float inv = 1.0f / *g.e_scale_dev; /* the ring planes carry no scale */
    float *dst = g.partial + (size_t)blockIdx.z * I * H + h;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float acc[32], cor[32];
#pragma unroll 1
    for (int c = 0; c < TC_DW2_BN; c += 32) {
      if (any) {
        tmem_ld32_nowait(taddr + c, acc);
        tmem_ld32(taddr + DW2_CORR_COL + c, cor);
      }
      else {
#pragma unroll
        for (int j = 0; j < 32; j++)
          acc[j] = cor[j] = 0.0f;
      }
      if (h < H) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          int i = i0 + c + j;
          if (i < I)
            dst[(size_t)i * H] = fmaf(cor[j], RB_LO_UNGAIN, acc[j]) * inv;
        }
      }
    }
    tc_fence_before();
  }
  /* nobody leaves (or frees tensor memory) while the other CTA may still be
     read by the pair's MMAs or signalled by their commits */
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, DW2_TMEM_COLS);
  }
}

/* delta (+)= sum of the split-K partials, in a fixed order */
__global__ void __launch_bounds__(256)
k_dw_reduce(float *__restrict__ delta, const float *__restrict__ partial, int size, int splits,
    int accumulate)
{
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < size;
       i += gridDim.x * blockDim.x * 4) {
    float4 a = accumulate ? *(const float4 *)(delta + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; s++) {
      float4 p = *(const float4 *)(partial + (size_t)s * size + i);
      a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    }
    *(float4 *)(delta + i) = a;
  }
}

/* The end of a training step in one pass over Wih: sum the split-K partials
   of the weight gradient (or take the finished delta), apply the optimiser
   (a13), and write the new weights together with the four operand planes the
   next step's GEMMs read - what k_dw_reduce, k_apply_learning and
   k_split_weights do in three passes.  32x32 tiles because of the transposed
   planes; the blocks past the last tile update Who elementwise.             */
struct UpdateArgs {
  float *W, *delta, *mom, *aux;
  const float *partial; /* NULL: delta is final */
  int splits, accumulate;
  int I, H;
  int method;
  float rate, momentum, momentum_weight;
  rb_h16 *Whi, *Wlo, *WThi, *WTlo;
  int wpitch, tpitch;
  /* the output matrix rides along */
  float *ho_W, *ho_mom, *ho_aux;
  const float *ho_delta;
  float *ho_delta_out;  /* where the API shows ho_delta, when ho_delta is read elsewhere */
  int ho_size;
  float ho_rate;
  int n_tiles_x, n_tiles;
  /* multi-GPU: `partial` and `ho_delta` are the peer exchange's result block;
     every rank's second epoch must be in before it is read */
  const unsigned int *wait_flags;
  unsigned int wait_epoch;
  int wait_n;
};

__device__ __forceinline__ unsigned int
ld_acquire_sys_u32(const unsigned int *p)
{
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
k_update_split(UpdateArgs a)
{
  if (a.wait_flags) {
    if ((int)threadIdx.x < a.wait_n) {
      const unsigned int *f = a.wait_flags + 8 /* second round */ + threadIdx.x;
      unsigned int spins = 0;
      while (ld_acquire_sys_u32(f) < a.wait_epoch) {
        if (++spins > (1u << 24))
          __trap();
      }
    }
    __syncthreads();
  }
  if ((int)blockIdx.x >= a.n_tiles) {
    int i = ((int)blockIdx.x - a.n_tiles) * 256 + threadIdx.x;
    if (i < a.ho_size) {
      float d = a.wait_flags ? __ldcv(a.ho_delta + i) : a.ho_delta[i];
      if (a.ho_delta_out)
        a.ho_delta_out[i] = d;
      a.ho_W[i] = rb_optimiser_step(a.method, a.ho_W[i], d, a.ho_mom, a.ho_aux, i,
          a.ho_rate, a.momentum, a.momentum_weight);
    }
    return;
  }
  __shared__ rb_h16 th[32][34], tl[32][34];
  const int I = a.I, H = a.H;
  const size_t size = (size_t)I * H;
  const int x0 = (blockIdx.x % a.n_tiles_x) * 32, y0 = (blockIdx.x / a.n_tiles_x) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = x0 + tx;
  float d[4], w[4];
  /* all loads of the thread's four elements first */
#pragma unroll
  for (int q = 0; q < 4; q++) {
    int y = y0 + ty + 8 * q;
    d[q] = 0.0f;
    w[q] = 0.0f;
    if (y < I && x < H) {
      size_t i = (size_t)y * H + x;
      w[q] = a.W[i];
      if (a.partial) {
        float t = a.accumulate ? a.delta[i] : 0.0f;
        if (a.wait_flags)
          t = __ldcv(a.partial + i); /* written by peers: past the caches */
        else {
#pragma unroll
          for (int z = 0; z < TC_DW_SPLITS; z++)
            if (z < a.splits)
              t += __ldcg(a.partial + (size_t)z * size + i);
        }
        d[q] = t;
      }
      else
        d[q] = a.delta[i];
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    int r = ty + 8 * q, y = y0 + r;
    rb_h16 hi = 0, lo = 0;
    if (y < I && x < H) {
      size_t i = (size_t)y * H + x;
      if (a.partial)
        a.delta[i] = d[q];
      float nw = rb_optimiser_step(a.method, w[q], d[q], a.mom, a.aux, i, a.rate, a.momentum,
          a.momentum_weight);
      a.W[i] = nw;
      rb_split_f16(nw * RB_W_SCALE, hi, lo);
      a.Whi[(size_t)y * a.wpitch + x] = hi;
      a.Wlo[(size_t)y * a.wpitch + x] = lo;
    }
    th[r][tx] = hi;
    tl[r][tx] = lo;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; q++) {
    int r = ty + 8 * q;
    int xx = x0 + r, y = y0 + tx;
    if (xx < H && y < I) {
      a.WThi[(size_t)xx * a.tpitch + y] = th[tx][r];
      a.WTlo[(size_t)xx * a.tpitch + y] = tl[tx][r];
    }
  }
}

/* ======================================================================== */
/* host side                                                                  */

extern "C" int
rb_tc_usable(const RbView *v)
{
  /* ReCLIP20 nets stay on the FMA engine: the reference leaves saturated rows
     (x >= 20) out of the weight gradient (recur-nn.c:347), and the operand
     planes of the ring, shared with the forward pass, hold them unmasked.
     i_size: a ring row's entries are bounded by the input soft clip at about
     27 * i_size (recur-nn.c:68-81), which must stay inside FP16. */
  return v->contiguous && v->n >= 64 && (v->n % TC_DW_BK) == 0 && v->d.h_size >= 64 &&
      v->d.i_size <= 2400 && v->activation != RNN_RECLIP20;
}

static void
refresh_weight_planes(RbTc *t, RbPool *p, const RbView *v)
{
  RbGroup *g = p->group;
  if (t->w_src == v->Wih && t->w_version == g->weights_version)
    return;
  const int I = v->d.i_size, H = v->d.h_size;
  dim3 grid(cdiv(H, 32), cdiv(I, 32));
  rb_prof_begin(RB_PROF_SMALL);
  k_split_weights<<<grid, 256, 0, rb_stream>>>(v->Wih, I, H, t->Whi, t->Wlo, t->wpitch, t->WThi,
      t->WTlo, t->pitch);
  LAUNCH_CHECK("k_split_weights");
  rb_prof_end(RB_PROF_SMALL);
  t->w_src = v->Wih;
  t->w_version = g->weights_version;
}

static int fwd_attr_done = 0, nt_attr_done = 0;
static int defer_delta_reduce = 0;
typedef NtCfg<TC_FWD_BN, TC_FWD_STAGES> FwdCfg;
typedef NtCfg<TC_NT_BN, TC_NT_STAGES> NtBigCfg;

static void
nt_big_attr(void)
{
  if (!nt_attr_done) {
    CUDA_OR_DIE(cudaFuncSetAttribute(k_tc_nt<TC_NT_BN, TC_NT_STAGES>,
            cudaFuncAttributeMaxDynamicSharedMemorySize, NtBigCfg::SMEM_BYTES));
    nt_attr_done = 1;
  }
}

extern "C" void
rb_tc_x_planes(RbPool *p, RbPlanes *X)
{
  RbTc *t = tc_state(p);
  *X = planes_of(t->Xhi, t->Xlo, t->pitch, 1.0f, NULL);
}

extern "C" void
rb_tc_forward(RbPool *p, const RbView *v, float presynaptic_noise)
{
  RbTc *t = tc_state(p);
  rbk_prepare_x(v);
  rb_prof_begin(RB_PROF_SMALL);
  k_split_rows<<<v->n, 256, 0, rb_stream>>>(*v, planes_of(t->Xhi, t->Xlo, t->pitch, 1.0f, NULL));
  LAUNCH_CHECK("k_split_rows");
  rb_prof_end(RB_PROF_SMALL);
  if (p->x_planes_stale == 1)
    p->x_planes_stale = 0; /* the row whose inputs were set is the one just split */
  rb_tc_forward_core(p, v, presynaptic_noise);
}

/* the forward contraction once the input row and its planes are in the ring */
extern "C" void
rb_tc_forward_core(RbPool *p, const RbView *v, float presynaptic_noise)
{
  RbTc *t = tc_state(p);
  refresh_weight_planes(t, p, v);
  NtArgs g;
  g.v = *v;
  g.mode = 0;
  g.k = 0;
  g.use_noise = 0;
  g.cpartial = NULL;
  g.inv_scale = 1.0f / RB_W_SCALE; /* ring planes: scale 1 */
  g.a_scale_dev = NULL;
  if (presynaptic_noise != 0.0f) {
    rbk_gen_noise(v, presynaptic_noise, 1, v->d.h_size - 1);
    g.use_noise = 1;
  }
  /* Few tiles and a long K: split K so the GEMM covers the SMs, and let the
     output-layer kernel sum the partials on its way in. */
  {
    int n_kb = cdiv(v->d.i_size, TC_KB);
    int tiles = cdiv(v->d.h_size, TC_NT_BN) * cdiv(v->n, TC_BM);
    int splits = TC_NT_SPLITS;
    while (splits > 1 && (n_kb / splits < 2 || tiles * splits > 148))
      splits /= 2;
    if (splits > 1 && rbk_output_takes_partials(v, splits)) {
      nt_big_attr();
      g.mode = 2;
      g.cpartial = t->cpartial;
      dim3 grid(cdiv(v->d.h_size, TC_NT_BN), cdiv(v->n, TC_BM), splits);
      rb_prof_begin(RB_PROF_FWD);
      k_tc_nt<TC_NT_BN, TC_NT_STAGES><<<grid, 192, NtBigCfg::SMEM_BYTES, rb_stream>>>(
          t->mXhi_k, t->mXlo_k, t->mWThi_k128, t->mWTlo_k128, g);
      LAUNCH_CHECK("k_tc_nt<FWD split-K>");
      rb_prof_end(RB_PROF_FWD);
      RbFwdPartials fp;
      fp.part = t->cpartial;
      fp.pitch = (v->d.i_size + 31) & ~31;
      fp.split_stride = (size_t)v->cap * fp.pitch;
      fp.splits = splits;
      fp.use_noise = g.use_noise;
      rbk_output_from_partials(v, &fp);
      return;
    }
  }
  if (!fwd_attr_done) {
    CUDA_OR_DIE(cudaFuncSetAttribute(k_tc_nt<TC_FWD_BN, TC_FWD_STAGES>,
            cudaFuncAttributeMaxDynamicSharedMemorySize, FwdCfg::SMEM_BYTES));
    fwd_attr_done = 1;
  }
  dim3 grid(cdiv(v->d.h_size, TC_FWD_BN), cdiv(v->n, TC_BM), 1);
  rb_prof_begin(RB_PROF_FWD);
  k_tc_nt<TC_FWD_BN, TC_FWD_STAGES><<<grid, 192, FwdCfg::SMEM_BYTES, rb_stream>>>(t->mXhi_k,
      t->mXlo_k, t->mWThi_k, t->mWTlo_k, g);
  LAUNCH_CHECK("k_tc_nt<FWD>");
  rb_prof_end(RB_PROF_FWD);
  rbk_output(v);
}

/* ---- the persistent chain's launch plan ---------------------------------- */

typedef struct ChainPlan {
  int ok;
  int splits, n_tiles, m_tiles, nkb_total, kb_per, tail0;
} ChainPlan;

template <int SPLITS>
static int
chain_clusters_fit(int n_clusters, dim3 grid)
{
  static int attr_done = 0, coop = -1;
  if (!attr_done) {
    if (cudaFuncSetAttribute(k_tc_chain_persistent<SPLITS, false>,
            cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_tc_chain_persistent<SPLITS, true>,
            cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    int dev = 0;
    CUDA_OR_DIE(cudaGetDevice(&dev));
    CUDA_OR_DIE(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    attr_done = 1;
  }
  if (!coop)
    return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(CH_THREADS);
  cfg.dynamicSmemBytes = CH_SMEM_BYTES;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = SPLITS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, k_tc_chain_persistent<SPLITS, false>, &cfg) !=
      cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return max_clusters >= n_clusters;
}

/* Column tiles x K splits x stream tiles, every cluster resident at once.
   All i_size columns on the tensor cores if that many clusters fit the GPU,
   else the h_size columns there and the rest on the CUDA-core tail. */
static ChainPlan
chain_plan(const RbView *v)
{
  ChainPlan pl;
  memset(&pl, 0, sizeof(pl));
  if (getenv("RECUR_B200_NO_PERSISTENT"))
    return pl;
  const int I = v->d.i_size, H = v->d.h_size;
  pl.nkb_total = cdiv(H, TC_KB);
  pl.m_tiles = cdiv(v->n, TC_BM);
  for (int splits = 4; splits >= 2 && !pl.ok; splits /= 2) {
    int kb_per = cdiv(pl.nkb_total, splits);
    if (kb_per > CH_WKB || (splits - 1) * kb_per >= pl.nkb_total)
      continue; /* the slice of Wih must fit, and every split must have work */
    int candidates[2] = {cdiv(I, CH_BN), cdiv(H, CH_BN)};
    for (int c = 0; c < 2 && !pl.ok; c++) {
      int n_tiles = candidates[c];
      if (c == 1 && n_tiles == candidates[0])
        break;
      if (n_tiles * splits + 1 > CH_SQ_SLOTS)
        continue;
      dim3 grid(n_tiles * splits, pl.m_tiles, 1);
      int fit = (splits == 4) ? chain_clusters_fit<4>(n_tiles * pl.m_tiles, grid)
                              : chain_clusters_fit<2>(n_tiles * pl.m_tiles, grid);
      if (fit) {
        pl.ok = 1;
        pl.splits = splits;
        pl.kb_per = kb_per;
        pl.n_tiles = n_tiles;
        pl.tail0 = (n_tiles * CH_BN < I) ? n_tiles * CH_BN : I;
      }
    }
  }
  return pl;
}

template <int SPLITS>
static void
launch_chain(RbTc *t, const ChainPlan *pl, ChainArgs *ca)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl->n_tiles * pl->splits, pl->m_tiles, 1);
  cfg.blockDim = dim3(CH_THREADS);
  cfg.dynamicSmemBytes = CH_SMEM_BYTES;
  cfg.stream = rb_stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = SPLITS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  /* cooperative: the driver guarantees that all CTAs are resident together
     or fails the launch, whatever else shares the device (the kernel spins
     on barriers between CTAs) */
  static int cooperative = -1;
  if (cooperative < 0)
    cooperative = getenv("RECUR_B200_PLAIN_LAUNCH") ? 0 : 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = cooperative;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  auto kernel = ca->dbg ? k_tc_chain_persistent<SPLITS, true> : k_tc_chain_persistent<SPLITS, false>;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, t->mEhi_k, t->mElo_k, t->mWhi_k, t->mWlo_k,
      *ca);
  if (e != cudaSuccess && cooperative) {
    /* a driver that does not combine cooperative launches with clusters:
       co-residency then rests on cudaOccupancyMaxActiveClusters (chain_plan)
       and on nothing else using the device; say so once */
    cudaGetLastError();
    fprintf(stderr, "recur-b200: cooperative cluster launch refused (%s); launching the "
        "persistent chain kernel plainly\n", cudaGetErrorString(e));
    cooperative = 0;
    at[1].val.cooperative = 0;
    e = cudaLaunchKernelEx(&cfg, kernel, t->mEhi_k, t->mElo_k, t->mWhi_k, t->mWlo_k, *ca);
  }
  if (e != cudaSuccess)
    rb_die("recur-b200: launch of k_tc_chain_persistent failed: %s", cudaGetErrorString(e));
}

extern "C" void
rb_tc_top_and_bptt(RbPool *p, const RbView *v, float *ho_delta, float *ih_delta, int accumulate)
{
  RbTc *t = tc_state(p);
  refresh_weight_planes(t, p, v);
  const RbPlanes Xp = planes_of(t->Xhi, t->Xlo, t->pitch, 1.0f, NULL);
  const RbPlanes Ep = planes_of(t->Ehi, t->Elo, t->pitch, 0.0f, t->escale);
  if (p->x_planes_stale) {
    /* ring rows were rewritten behind the planes' back (rnn_forget_history,
       rnn_b200_push, a per-net or FMA forward, a regrown pool): the weight
       gradient reads the ring only through the planes */
    rb_prof_begin(RB_PROF_SMALL);
    k_split_ring<<<148 * 4, 256, 0, rb_stream>>>(p->X, v->d.i_size, (size_t)p->depth * p->cap, Xp);
    LAUNCH_CHECK("k_split_ring");
    rb_prof_end(RB_PROF_SMALL);
    p->x_planes_stale = 0;
  }
  /* top layer: E(0), then its planes and the scale of this walk's planes */
  rbk_top_layer_defer_ho_delta(1);
  rbk_top_layer_begin(v, ho_delta, accumulate, NULL, 0);
  rbk_top_layer_defer_ho_delta(0);
  rb_prof_begin(RB_PROF_TOP);
  k_e0_planes<<<v->n, 256, 0, rb_stream>>>(*v, Ep, t->escale);
  LAUNCH_CHECK("k_e0_planes");
  rb_prof_end(RB_PROF_TOP);
  /* (ho_delta goes on beside the walk, on the SMs the chain kernel leaves
     free: nothing reads it before the exchange / the update, joined below) */
  rbk_top_layer_mark();

  /* small nets: every stream walks alone with the weights resident in its SM */
  const bool resident = rbk_walk_resident_usable(v);
  ChainPlan pl;
  memset(&pl, 0, sizeof(pl));
  if (!resident)
    pl = chain_plan(v);
  unsigned int *sync_area = t->sync + (size_t)t->sync_flip * t->sync_words;
  unsigned int *sync_next = t->sync + (size_t)(t->sync_flip ^ 1) * t->sync_words;
  t->sync_flip ^= 1;
  unsigned int *kmax_dev = sync_area + (size_t)cdiv(v->n, TC_BM) * CH_SYNC_STRIDE + 4;
  rb_note_walk_kernel(pl.ok ? "k_tc_chain_persistent"
      : resident ? "k_walk_resident" : "k_tc_nt<CHAIN>");
  if (pl.ok) {
    ChainArgs ca;
    ca.v = *v;
    ca.E = Ep;
    ca.sqpart = t->sqpart;
    ca.sync = sync_area;
    ca.kmax = kmax_dev;
    ca.splits = pl.splits;
    ca.nkb_total = pl.nkb_total;
    ca.kb_per = pl.kb_per;
    ca.tail0 = pl.tail0;
    ca.dbg = NULL;
    static unsigned long long *dbg_dev = NULL;
    static int dbg_calls = 0;
    const bool timing = getenv("RECUR_B200_CHAIN_TIMING") != NULL;
    if (timing) {
      if (!dbg_dev)
        CUDA_OR_DIE(cudaMalloc((void **)&dbg_dev, (64 * 32 + 256) * sizeof(unsigned long long)));
      CUDA_OR_DIE(cudaMemsetAsync(dbg_dev, 0, (64 * 32 + 256) * sizeof(unsigned long long), rb_stream));
      ca.dbg = dbg_dev;
    }
    rb_prof_begin(RB_PROF_CHAIN);
    if (pl.splits == 4)
      launch_chain<4>(t, &pl, &ca);
    else
      launch_chain<2>(t, &pl, &ca);
    LAUNCH_CHECK("k_tc_chain_persistent");
    rb_prof_end(RB_PROF_CHAIN);
    if (timing && (++dbg_calls % 100) == 60) {
      static unsigned long long h[64 * 32 + 256];
      CUDA_OR_DIE(cudaMemcpyAsync(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost, rb_stream));
      CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
      static const char *names[20] = {"step start", "barrier seen", "first block landed",
        "MMAs issued", "accumulator ready", "slices parked + fenced", "tail done", "verdicts in",
        "cluster's K loops over", "arrived", "peers' slices in", "rows stored", "proxy fence",
        "block barrier", "tensor memory drained", "slices summed", "row 0 masked", "row 0 split",
        "row 0 stores issued", "row 0 done"};
      static const int order[19] = {1, 7, 2, 3, 4, 14, 5, 6, 8, 10, 15, 16, 17, 18, 19, 11, 12, 13, 9};
      double sum[20] = {0};
      int steps = 0;
      for (int k = 2; k < 64 && h[k * 32 + 9]; k++, steps++)
        for (int q = 1; q < 20; q++)
          if (h[k * 32 + q])
            sum[q] += (double)(h[k * 32 + q] - h[k * 32]);
      fprintf(stderr, "chain timing, CTA 0, mean over %d steps, us after the step's start:", steps);
      for (int q = 0; q < 19 && steps; q++)
        fprintf(stderr, " %s %.2f;", names[order[q]], sum[order[q]] / steps * 1e-3);
      if (steps > 1)
        fprintf(stderr, " step period %.2f\n", (double)(h[(steps + 1) * 32] - h[2 * 32]) / (steps - 1) * 1e-3);
      else
        fprintf(stderr, "\n");
      {
        unsigned long long first = ~0ull;
        for (int i = 0; i < 256; i++)
          if (h[64 * 32 + i] && h[64 * 32 + i] < first)
            first = h[64 * 32 + i];
        fprintf(stderr, "  arrivals of step 6, ns after the first (rows: m tile; columns: CTA x):\n");
        const int gx = pl.n_tiles * pl.splits;
        for (int y = 0; y < pl.m_tiles; y++) {
          fprintf(stderr, "   ");
          for (int x = 0; x < gx; x++)
            fprintf(stderr, " %4llu", h[64 * 32 + y * gx + x] ? h[64 * 32 + y * gx + x] - first : 0ull);
          fprintf(stderr, "\n");
        }
      }
    }
  }
  else if (resident) {
    rbk_walk_resident(v, &Ep);
  }
  else {
    nt_big_attr();
    NtArgs g;
    g.v = *v;
    g.mode = 1;
    g.use_noise = 0;
    g.cpartial = t->cpartial;
    g.inv_scale = 1.0f / RB_W_SCALE;
    g.a_scale_dev = t->escale;
    /* split K only as far as there are K blocks to share out */
    int n_kb = cdiv(v->d.h_size, TC_KB);
    int splits = TC_NT_SPLITS;
    while (splits > 1 && n_kb / splits < 2)
      splits /= 2;
    dim3 cgrid(cdiv(v->d.i_size, TC_NT_BN), cdiv(v->n, TC_BM), splits);
    for (int k = 0; k < v->depth; k++) {
      g.k = k;
      rb_prof_begin(RB_PROF_CHAIN);
      k_tc_nt<TC_NT_BN, TC_NT_STAGES><<<cgrid, 192, NtBigCfg::SMEM_BYTES, rb_stream>>>(
          t->mEhi_k, t->mElo_k, t->mWhi_k, t->mWlo_k, g);
      LAUNCH_CHECK("k_tc_nt<CHAIN>");
      k_chain_finish_step<<<v->n, 256, 0, rb_stream>>>(*v, k, t->cpartial, splits, Ep);
      LAUNCH_CHECK("k_chain_finish_step");
      rb_prof_end(RB_PROF_CHAIN);
    }
  }
  if (!pl.ok) {
    k_compute_kmax<<<1, 256, 0, rb_stream>>>(*v, kmax_dev);
    LAUNCH_CHECK("k_compute_kmax");
  }
  rbk_top_layer_ho_delta_now(); /* behind the walk's kernel(s), beside them on the GPU */
  rb_prof_begin(RB_PROF_SMALL);
  k_finalize_rows<<<v->n, 256, 0, rb_stream>>>(*v, Ep, kmax_dev, sync_next, (int)t->sync_words);
  LAUNCH_CHECK("k_finalize_rows");
  rb_prof_end(RB_PROF_SMALL);

  DwArgs d;
  d.v = *v;
  d.partial = t->partial;
  d.kmax = kmax_dev;
  d.e_scale_dev = t->escale;
  static int dw_sms = 0, dw_pair_ok = 0;
  if (!dw_sms) {
    int dev = 0;
    CUDA_OR_DIE(cudaGetDevice(&dev));
    CUDA_OR_DIE(cudaDeviceGetAttribute(&dw_sms, cudaDevAttrMultiProcessorCount, dev));
    dw_pair_ok = !getenv("RECUR_B200_NO_PAIR_DW") &&
        cudaFuncSetAttribute(k_tc_dw_pair, cudaFuncAttributeMaxDynamicSharedMemorySize,
            DW2_SMEM_BYTES) == cudaSuccess;
  }
  rb_prof_begin(RB_PROF_DW);
  int pgx = 2 * cdiv(v->d.h_size, 2 * TC_BM), pgy = cdiv(v->d.i_size, TC_DW2_BN);
  int psplits = dw_sms / (pgx * pgy);
  if (psplits > TC_DW_SPLITS)
    psplits = TC_DW_SPLITS;
  int size = v->d.i_size * v->d.h_size;
  if (dw_pair_ok && psplits >= 1) {
    /* CTA pairs: one MMA across two SMs, each holding half of the N operand */
    dim3 dgrid(pgx, pgy, psplits);
    k_tc_dw_pair<<<dgrid, 192, DW2_SMEM_BYTES, rb_stream>>>(t->mEhi_mn, t->mElo_mn, t->mXhi_mn,
        t->mXlo_mn, d);
    LAUNCH_CHECK("k_tc_dw_pair");
    t->dw_splits = psplits;
  }
  else {
    /* no pair grid for this shape: the FMA engine's weight gradient, parked
       as a single "split" */
    rbk_dw_fma(v, t->partial, 0);
    t->dw_splits = 1;
  }
  rbk_top_layer_join(); /* ho_delta: whatever is queued from here on may read it */
  if (rb_p2p_ready(v->p2p)) {
    /* multi-GPU: the split-K sum is the first phase of the exchange kernel;
       whoever consumes the result waits for the peers' last stores */
    if (accumulate)
      rb_die("recur-b200: accumulate != 0 across GPUs (see recur_b200.h)");
    rb_p2p_exchange(v->p2p, t->partial, t->dw_splits, size, v->d.h_size * v->d.o_size, ho_delta);
    t->delta_pending = 2;
    t->pending_accumulate = 0;
    t->pending_p2p = v->p2p;
    if (!defer_delta_reduce)
      rb_tc_materialise_delta(p, ih_delta);
  }
  else if (defer_delta_reduce) {
    /* the caller applies the update next: rb_tc_fused_update sums the
       partials on its way through the weights */
    t->delta_pending = 1;
    t->pending_accumulate = accumulate;
  }
  else {
    k_dw_reduce<<<cdiv(size / 4, 256), 256, 0, rb_stream>>>(ih_delta, t->partial, size,
        t->dw_splits, accumulate);
    LAUNCH_CHECK("k_dw_reduce");
  }
  rb_prof_end(RB_PROF_DW);
}

/* Measurement aid: `rounds` gradient exchanges back to back on whatever the
   split-K planes hold, nothing else on the stream; mean microseconds of one.
   The exchanges synchronise the ranks among themselves (each waits for every
   peer's stores), so after the first the figure is the exchange's own latency
   without the ranks' skew from unequal BPTT depths.  A round = the exchange
   kernel + the consumer's wait and 4.4 MB local copy-out (the training step's
   consumer is the update kernel, which waits the same way). */
extern "C" float
rb_tc_exchange_probe(RbPool *p, void *p2p, RecurNN *net, int rounds)
{
  RbTc *t = (RbTc *)p->tc;
  if (!t || !rb_p2p_ready(p2p) || t->dw_splits < 1 || rounds < 1)
    return -1.0f;
  cudaEvent_t e0, e1;
  CUDA_OR_DIE(cudaEventCreate(&e0));
  CUDA_OR_DIE(cudaEventCreate(&e1));
  rb_p2p_exchange(p2p, t->partial, t->dw_splits, net->ih_size, net->ho_size,
      net->bptt->ho_delta); /* lines the ranks up */
  CUDA_OR_DIE(cudaEventRecord(e0, rb_stream));
  for (int i = 0; i < rounds; i++) {
    /* the consumer's wait for every rank's result stores is part of a round:
       it is what makes the inboxes free for the next one */
    rb_p2p_copy_out(p2p, net->bptt->ih_delta);
    rb_p2p_exchange(p2p, t->partial, t->dw_splits, net->ih_size, net->ho_size,
        net->bptt->ho_delta);
  }
  CUDA_OR_DIE(cudaEventRecord(e1, rb_stream));
  rb_p2p_copy_out(p2p, net->bptt->ih_delta);
  CUDA_OR_DIE(cudaEventSynchronize(e1));
  float ms = 0.0f;
  CUDA_OR_DIE(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return ms * 1e3f / rounds;
}

/* A caller that runs the update right after the deltas (the char step) lets
   the weight gradient stay in its split-K planes in between. */
extern "C" void
rb_tc_defer_delta_reduce(int on)
{
  defer_delta_reduce = on;
}

/* ih_delta as the API promises it, for anyone who looks before the update */
extern "C" void
rb_tc_materialise_delta(RbPool *p, float *ih_delta)
{
  RbTc *t = (RbTc *)p->tc;
  if (!t || !t->delta_pending)
    return;
  const RbDims *d = &p->group->d;
  int size = d->i_size * d->h_size;
  if (t->delta_pending == 2) {
    rb_p2p_copy_out(t->pending_p2p, ih_delta); /* [ih_delta | ho_delta] are adjacent */
    t->delta_pending = 0;
    return;
  }
  k_dw_reduce<<<cdiv(size / 4, 256), 256, 0, rb_stream>>>(ih_delta, t->partial, size,
      t->dw_splits, t->pending_accumulate);
  LAUNCH_CHECK("k_dw_reduce");
  t->delta_pending = 0;
}

/* The ih and ho updates of rnn_apply_learning fused with the weight-gradient
   reduction before and the operand split after.  Returns 0 (nothing done) when
   this pool's tensor engine holds no planes of these weights. */
extern "C" int
rb_tc_fused_update(RbPool *p, RecurNN *net, int method, float momentum, float momentum_weight)
{
  RbTc *t = (RbTc *)p->tc;
  if (!t)
    return 0;
  RecurNNBPTT *b = net->bptt;
  if (t->w_src != net->ih_weights) {
    rb_tc_materialise_delta(p, b->ih_delta);
    return 0;
  }
  const RbDims *d = &p->group->d;
  UpdateArgs a;
  a.W = net->ih_weights;
  a.delta = b->ih_delta;
  a.mom = b->ih_momentum;
  a.aux = b->ih_aux;
  a.partial = t->delta_pending ? t->partial : NULL;
  a.splits = t->dw_splits;
  a.accumulate = t->pending_accumulate;
  a.ho_delta = b->ho_delta;
  a.ho_delta_out = NULL;
  a.wait_flags = NULL;
  a.wait_epoch = 0;
  a.wait_n = 0;
  if (t->delta_pending == 2) {
    const float *result;
    rb_p2p_result(t->pending_p2p, &result, &a.wait_flags, &a.wait_epoch, &a.wait_n);
    a.partial = result;
    a.splits = 1;
    a.accumulate = 0;
    a.ho_delta = result + net->ih_size;
    a.ho_delta_out = b->ho_delta;
  }
  a.I = d->i_size;
  a.H = d->h_size;
  a.method = method;
  a.rate = b->learn_rate;
  a.momentum = momentum;
  a.momentum_weight = momentum_weight;
  a.Whi = t->Whi;
  a.Wlo = t->Wlo;
  a.WThi = t->WThi;
  a.WTlo = t->WTlo;
  a.wpitch = t->wpitch;
  a.tpitch = t->pitch;
  a.ho_W = net->ho_weights;
  a.ho_mom = b->ho_momentum;
  a.ho_aux = b->ho_aux;
  a.ho_size = net->ho_size;
  a.ho_rate = b->learn_rate * b->ho_scale;
  a.n_tiles_x = cdiv(a.H, 32);
  a.n_tiles = a.n_tiles_x * cdiv(a.I, 32);
  rb_prof_begin(RB_PROF_UPDATE);
  k_update_split<<<a.n_tiles + cdiv(a.ho_size, 256), 256, 0, rb_stream>>>(a);
  LAUNCH_CHECK("k_update_split");
  rb_prof_end(RB_PROF_UPDATE);
  t->delta_pending = 0;
  return 1;
}

/* after rb_weights_changed: the planes rb_tc_fused_update wrote are those of
   the new weights */
extern "C" void
rb_tc_planes_current(RbPool *p)
{
  RbTc *t = (RbTc *)p->tc;
  if (t)
    t->w_version = p->group->weights_version;
}
