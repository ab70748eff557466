/* rb_tc.cu — the tensor-core engine: the three big contractions of the path
 * as tcgen05 (5th-gen tensor core) kernels for sm_100a, used when a batch has
 * >= 64 streams (a multiple of 32) in one contiguous run of pool slots.
 *
 *   FWD    hidden[b, h]   = act( sum_y x[b, y] * Wih[y, h] )       (recur-nn.c:117-148)
 *   CHAIN  E(k+1)[b, y]   = mask * sum_x E(k)[b, x] * Wih[y, x]    (recur-nn.c:338-376)
 *   DW     delta[y, x]   += sum_{k,b} x_k[b, y] * E(k)[b, x]       (recur-nn.c:353-356)
 *
 * Numerics: "3xTF32".  Every FP32 operand a is split exactly into
 * hi = tf32(a) and lo = tf32(a - hi); a product a*b is issued as three
 * kind::tf32 MMAs hi*hi + hi*lo + lo*hi accumulated in FP32 in tensor
 * memory, which keeps ~21 mantissa bits per product (the dropped lo*lo term
 * is 2^-22 relative) — FP32-faithful within the 1e-4 parity tolerance.
 * The hi/lo planes of the operands are materialised once where each operand
 * is produced (weights: after an update; input rows: when the row enters the
 * ring; error rows: in the CHAIN epilogue), so the GEMM mainloops are pure
 * TMA -> shared memory -> tcgen05.mma pipelines.
 *
 * Kernel anatomy (both kernels): 192 threads = warp 0 TMA producer (one
 * elected lane), warp 1 TMEM allocator + MMA issuer (one elected lane),
 * warps 2..5 epilogue (one TMEM lane quarter each).  Operand tiles are
 * 128B-swizzled; FWD/CHAIN read both operands K-major, DW reads both
 * MN-major (the contraction runs over ring rows).
 */
#include "rb_kernels.h"
#include "rb_host.h"
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

/* ======================================================================== */
/* PTX wrappers                                                               */

__device__ __forceinline__ uint32_t
smem_u32(const void *p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void
mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void
mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
      ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void
mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void
fence_barrier_init(void)
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void
tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1)
{
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void
tma_prefetch_desc(const CUtensorMap *map)
{
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void
tmem_alloc(uint32_t *dst_smem, uint32_t cols)
{
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
      ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void
tmem_dealloc(uint32_t addr, uint32_t cols)
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
      : "memory");
}

__device__ __forceinline__ void
tc_fence_before(void)
{
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void
tc_fence_after(void)
{
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

/* D[tmem] (+)= A[smem] . B[smem], kind::tf32, issued by one thread */
__device__ __forceinline__ void
umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

/* arrive on an mbarrier when all MMAs issued so far have completed */
__device__ __forceinline__ void
umma_commit(uint64_t *bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
      ::"r"(smem_u32(bar)) : "memory");
}

/* 32 consecutive accumulator columns of this thread's TMEM lane */
__device__ __forceinline__ void
tmem_ld32(uint32_t taddr, float *v)
{
  uint32_t *r = (uint32_t *)v;
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31},"
      " [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

/* Shared-memory matrix descriptor for a 128B-swizzled operand tile
   (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address, leading and
   stride byte offsets in 16-byte units, version 1, layout SWIZZLE_128B. */
#define UMMA_SW128 2u        /* 16-byte chunks swizzled over 8 rows */
#define UMMA_SW128_BASE32 1u /* 32-byte chunks swizzled over 4 rows: the only
                                layout for MN-major 32-bit operands */

__device__ __forceinline__ uint64_t
umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
    uint32_t layout = UMMA_SW128)
{
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  /* descriptor version for sm_100 */
  d |= (uint64_t)layout << 61;
  return d;
}

/* Instruction descriptor (mma_sm100_desc.hpp InstrDescriptor) for
   kind::tf32, FP32 accumulate. */
__host__ __device__ constexpr uint32_t
umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major)
{
  return (1u << 4)      /* c_format  F32 */
    | (2u << 7)         /* a_format  TF32 */
    | (2u << 10)        /* b_format  TF32 */
    | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16)
    | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

/* exact split of an FP32 number into two TF32 numbers */
__device__ __forceinline__ void
split_tf32(float a, float &hi, float &lo)
{
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
  hi = __uint_as_float(h);
  float r = a - hi;
  uint32_t l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
  lo = __uint_as_float(l);
}

/* ======================================================================== */
/* the engine's extra state per pool                                          */

#define TC_BM 128      /* tile rows (TMEM lanes) */
#define TC_BK 32       /* K per stage: 32 floats = one 128-byte swizzle row */
#define TC_NT_BN 64    /* FWD / CHAIN tile columns */
#define TC_NT_STAGES 4
#define TC_DW_BN 128   /* DW tile columns */
#define TC_DW_STAGES 3
#define TC_DW_SPLITS 2

typedef struct RbTc {
  int cap, depth;
  float *Xhi, *Xlo;     /* [depth][cap][i_size]   planes of the ring */
  float *Ehi, *Elo;     /* [depth+1][cap][i_size] planes of the error chain */
  float *Whi, *Wlo;     /* [i_size][h_size] */
  float *WThi, *WTlo;   /* [h_size][i_size] */
  float *partial;       /* [TC_DW_SPLITS][i_size][h_size] */
  const float *w_src;   /* weights the planes were made from */
  uint64_t w_version;
  /* tensor maps */
  CUtensorMap mXhi_k, mXlo_k;   /* ring rows as K-major A of FWD: box 32 x 128 */
  CUtensorMap mEhi_k, mElo_k;   /* error rows as K-major A of CHAIN (width h_size) */
  CUtensorMap mWhi_k, mWlo_k;   /* Wih rows as K-major B of CHAIN: box 32 x 64 */
  CUtensorMap mWThi_k, mWTlo_k; /* Wih^T rows as K-major B of FWD: box 32 x 64 */
  CUtensorMap mXhi_mn, mXlo_mn; /* ring rows as MN-major A of DW: box 32 x 32 */
  CUtensorMap mEhi_mn, mElo_mn; /* error rows as MN-major B of DW (width h_size) */
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

/* rows x width floats with a row pitch of pitch floats, boxes of 32 x box_rows */
static void
make_map(CUtensorMap *m, float *base, uint64_t width, uint64_t rows, uint64_t pitch,
    uint32_t box_rows, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B)
{
  cuuint64_t dims[2] = {width, rows};
  cuuint64_t strides[1] = {pitch * sizeof(float)};
  cuuint32_t box[2] = {TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box,
      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    rb_die("recur-b200: cuTensorMapEncodeTiled failed (%d) for %llu x %llu pitch %llu", (int)r,
        (unsigned long long)width, (unsigned long long)rows, (unsigned long long)pitch);
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
  cudaFree(t->partial);
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
  size_t ring = (size_t)p->depth * p->cap * I, chain = (size_t)(p->depth + 1) * p->cap * I;
  t->Xhi = dmalloc0<float>(ring);
  t->Xlo = dmalloc0<float>(ring);
  t->Ehi = dmalloc0<float>(chain);
  t->Elo = dmalloc0<float>(chain);
  t->Whi = dmalloc0<float>(I * H);
  t->Wlo = dmalloc0<float>(I * H);
  t->WThi = dmalloc0<float>(I * H);
  t->WTlo = dmalloc0<float>(I * H);
  t->partial = dmalloc0<float>((size_t)TC_DW_SPLITS * I * H);
  t->w_src = NULL;
  uint64_t ring_rows = (uint64_t)p->depth * p->cap, chain_rows = (uint64_t)(p->depth + 1) * p->cap;
  make_map(&t->mXhi_k, t->Xhi, I, ring_rows, I, TC_BM);
  make_map(&t->mXlo_k, t->Xlo, I, ring_rows, I, TC_BM);
  make_map(&t->mEhi_k, t->Ehi, H, chain_rows, I, TC_BM);
  make_map(&t->mElo_k, t->Elo, H, chain_rows, I, TC_BM);
  make_map(&t->mWhi_k, t->Whi, H, I, H, TC_NT_BN);
  make_map(&t->mWlo_k, t->Wlo, H, I, H, TC_NT_BN);
  make_map(&t->mWThi_k, t->WThi, I, H, I, TC_NT_BN);
  make_map(&t->mWTlo_k, t->WTlo, I, H, I, TC_NT_BN);
  make_map(&t->mXhi_mn, t->Xhi, I, ring_rows, I, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  make_map(&t->mXlo_mn, t->Xlo, I, ring_rows, I, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  make_map(&t->mEhi_mn, t->Ehi, H, chain_rows, I, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  make_map(&t->mElo_mn, t->Elo, H, chain_rows, I, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  p->tc = t;
  return t;
}

/* ======================================================================== */
/* operand planes                                                             */

/* Wih -> hi/lo planes, plain and transposed (32x32 tiles through smem) */
__global__ void __launch_bounds__(256)
k_split_weights(const float *__restrict__ W, int I, int H, float *Whi, float *Wlo,
    float *WThi, float *WTlo)
{
  __shared__ float th[32][33], tl[32][33];
  int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    int y = y0 + r, x = x0 + tx;
    float hi = 0.f, lo = 0.f;
    if (y < I && x < H) {
      split_tf32(W[(size_t)y * H + x], hi, lo);
      Whi[(size_t)y * H + x] = hi;
      Wlo[(size_t)y * H + x] = lo;
    }
    th[r][tx] = hi;
    tl[r][tx] = lo;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int x = x0 + r, y = y0 + tx;
    if (x < H && y < I) {
      WThi[(size_t)x * I + y] = th[tx][r];
      WTlo[(size_t)x * I + y] = tl[tx][r];
    }
  }
}

/* one ring / chain row per block -> its hi/lo planes */
__global__ void __launch_bounds__(256)
k_split_rows(RbView v, int which /* 0: current x row, 1: E[0] */, float *hi_plane, float *lo_plane)
{
  int s = v.slots[blockIdx.x];
  size_t off;
  const float *src;
  if (which == 0) {
    off = ((size_t)v.pos[s] * v.cap + s) * v.d.i_size;
    src = v.X + off;
  }
  else {
    off = (size_t)s * v.d.i_size;
    src = v.E + off;
  }
  for (int i = threadIdx.x; i < v.d.i_size; i += blockDim.x) {
    float hi, lo;
    split_tf32(src[i], hi, lo);
    hi_plane[off + i] = hi;
    lo_plane[off + i] = lo;
  }
}

/* After the walk: rows of E beyond a stream's executed depth must not reach
   the weight gradient (zero them), and a stream whose gradient is clipped
   (ih_scale != 1, recur-nn.c:393-402) has its rows rescaled and re-split. */
__global__ void __launch_bounds__(256)
k_finalize_rows(RbView v, float *Ehi, float *Elo)
{
  int s = v.slots[blockIdx.x];
  int step = blockIdx.y;
  const RbScalars sc = v.sc[s];
  size_t off = ((size_t)step * v.cap + s) * v.d.i_size;
  if (step >= sc.n_steps) {
    for (int i = threadIdx.x; i < v.d.h_size; i += blockDim.x) {
      Ehi[off + i] = 0.0f;
      Elo[off + i] = 0.0f;
    }
  }
  else if (sc.ih_scale != 1.0f) {
    for (int i = threadIdx.x; i < v.d.h_size; i += blockDim.x) {
      float hi, lo;
      split_tf32(v.E[off + i] * sc.ih_scale, hi, lo);
      Ehi[off + i] = hi;
      Elo[off + i] = lo;
    }
  }
}

/* ======================================================================== */
/* FWD and CHAIN: C[128 x 64] tiles, both operands K-major                    */

struct NtArgs {
  RbView v;
  int mode;        /* 0 FWD, 1 CHAIN */
  int k;           /* CHAIN: step */
  int use_noise;
  float *Ehi, *Elo;
};

#define NT_A_BYTES (TC_BM * TC_BK * 4)        /* 16 KB */
#define NT_B_BYTES (TC_NT_BN * TC_BK * 4)     /* 8 KB */
#define NT_STAGE_BYTES (2 * NT_A_BYTES + 2 * NT_B_BYTES)
#define NT_SMEM_BYTES (TC_NT_STAGES * NT_STAGE_BYTES + 1024 + 256)

__global__ void __launch_bounds__(192, 1)
k_tc_nt(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
    const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo,
    NtArgs g)
{
  extern __shared__ uint8_t smem_raw[];
  const RbView &v = g.v;
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *full = (uint64_t *)(smem + TC_NT_STAGES * NT_STAGE_BYTES);
  uint64_t *empty = full + TC_NT_STAGES;
  uint64_t *acc_ready = empty + TC_NT_STAGES;
  uint32_t *tmem_slot = (uint32_t *)(acc_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = v.d.i_size, H = v.d.h_size;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_NT_BN;
  const int K = (g.mode == 0) ? I : H;
  const int n_kb = (K + TC_BK - 1) / TC_BK;

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
    for (int s = 0; s < TC_NT_STAGES; s++) {
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
    tmem_alloc(tmem_slot, TC_NT_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      /* row of the A operand in its ring: FWD reads the newest x row, CHAIN E[k] */
      int ring_row;
      if (g.mode == 0)
        ring_row = v.pos[v.base] * v.cap + v.base + m0;
      else
        ring_row = g.k * v.cap + v.base + m0;
      for (int kb = 0; kb < n_kb; kb++) {
        int s = kb % TC_NT_STAGES;
        uint32_t ph = (kb / TC_NT_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t *st = smem + s * NT_STAGE_BYTES;
        mbar_expect_tx(&full[s], NT_STAGE_BYTES);
        tma_load_2d(&mAhi, &full[s], st, kb * TC_BK, ring_row);
        tma_load_2d(&mAlo, &full[s], st + NT_A_BYTES, kb * TC_BK, ring_row);
        tma_load_2d(&mBhi, &full[s], st + 2 * NT_A_BYTES, kb * TC_BK, n0);
        tma_load_2d(&mBlo, &full[s], st + 2 * NT_A_BYTES + NT_B_BYTES, kb * TC_BK, n0);
      }
    }
  }
  else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(TC_BM, TC_NT_BN, 0, 0);
      for (int kb = 0; kb < n_kb; kb++) {
        int s = kb % TC_NT_STAGES;
        uint32_t ph = (kb / TC_NT_STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        uint32_t a_hi = smem_u32(smem + s * NT_STAGE_BYTES);
        uint32_t a_lo = a_hi + NT_A_BYTES;
        uint32_t b_hi = a_hi + 2 * NT_A_BYTES;
        uint32_t b_lo = b_hi + NT_B_BYTES;
#pragma unroll
        for (int kk = 0; kk < TC_BK / 8; kk++) {
          /* K-major, 128B swizzle: 8-row groups 1024 B apart; a K step of 8
             floats moves the start address by 32 bytes inside the swizzle row */
          uint64_t dah = umma_desc(a_hi + kk * 32, 16, 1024);
          uint64_t dal = umma_desc(a_lo + kk * 32, 16, 1024);
          uint64_t dbh = umma_desc(b_hi + kk * 32, 16, 1024);
          uint64_t dbl = umma_desc(b_lo + kk * 32, 16, 1024);
          umma_tf32(tmem_base, dal, dbh, idesc, (kb | kk) ? 1u : 0u);
          umma_tf32(tmem_base, dah, dbl, idesc, 1u);
          umma_tf32(tmem_base, dah, dbh, idesc, 1u);
        }
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
    mbar_wait(acc_ready, 0);
    tc_fence_after();
    float acc[32];
    if (g.mode == 0) {
#pragma unroll 1
      for (int c = 0; c < TC_NT_BN; c += 32) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, acc);
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
            float h = acc[j + u];
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
      const int hs1 = v.d.hidden_size + 1;
      const bool live = row_ok && v.sc[sidx].live != 0;
      const int p = v.pos[sidx] - g.k;
      const float *xk = v.X + ((size_t)(p < 0 ? p + v.depth : p) * v.cap + sidx) * I;
      const size_t eoff = ((size_t)(g.k + 1) * v.cap + sidx) * I;
      float sq = 0.0f;
#pragma unroll 1
      for (int c = 0; c < TC_NT_BN; c += 32) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, acc);
        int col0 = n0 + c;
        if (!live || col0 >= I)
          continue;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (col0 + j >= I)
            break;
          float4 xin = *(const float4 *)(xk + col0 + j);
          float xi[4] = {xin.x, xin.y, xin.z, xin.w};
          float o[4], ohi[4], olo[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            float e = 0.0f;
            float input = xi[u];
            if (input != 0.0f && (v.activation != RNN_RECLIP20 || input < 20.0f)) {
              e = acc[j + u];
              if (v.activation == RNN_RESQRT)
                e /= 2.0f * (input + 1.0f);
              sq += e * e;
            }
            int col = col0 + j + u;
            if (col == 0 || (col >= hs1 && col < H))
              e = 0.0f;
            o[u] = e;
            split_tf32(e, ohi[u], olo[u]);
          }
          *(float4 *)(v.E + eoff + col0 + j) = make_float4(o[0], o[1], o[2], o[3]);
          *(float4 *)(g.Ehi + eoff + col0 + j) = make_float4(ohi[0], ohi[1], ohi[2], ohi[3]);
          *(float4 *)(g.Elo + eoff + col0 + j) = make_float4(olo[0], olo[1], olo[2], olo[3]);
        }
      }
      if (live)
        v.partial[(size_t)sidx * v.n_part + blockIdx.x] = sq;
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_NT_BN);
  }
}

/* ======================================================================== */
/* DW: delta tile [128 y x 128 x] += X^T . E over (step, stream) rows;
   both operands MN-major, split-K across blockIdx.z                          */

struct DwArgs {
  RbView v;
  float *partial; /* [splits][i_size][h_size] */
};

#define DW_OP_BYTES (TC_BM * TC_BK * 4) /* 16 KB: 4 chunks of 32(MN) x 32(K) */
#define DW_STAGE_BYTES (4 * DW_OP_BYTES)
#define DW_SMEM_BYTES (TC_DW_STAGES * DW_STAGE_BYTES + 1024 + 256)

__global__ void __launch_bounds__(192, 1)
k_tc_dw(const __grid_constant__ CUtensorMap mXhi, const __grid_constant__ CUtensorMap mXlo,
    const __grid_constant__ CUtensorMap mEhi, const __grid_constant__ CUtensorMap mElo,
    DwArgs g)
{
  extern __shared__ uint8_t smem_raw[];
  const RbView &v = g.v;
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *full = (uint64_t *)(smem + TC_DW_STAGES * DW_STAGE_BYTES);
  uint64_t *empty = full + TC_DW_STAGES;
  uint64_t *acc_ready = empty + TC_DW_STAGES;
  uint32_t *tmem_slot = (uint32_t *)(acc_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = v.d.i_size, H = v.d.h_size;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_DW_BN;
  const int kb_per_step = v.n / TC_BK;
  const int n_kb_total = v.depth * kb_per_step;
  const int kb_per_split = (n_kb_total + gridDim.z - 1) / gridDim.z;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(n_kb_total, kb_begin + kb_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_DW_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_ready, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mXhi);
    tma_prefetch_desc(&mXlo);
    tma_prefetch_desc(&mEhi);
    tma_prefetch_desc(&mElo);
  }
  if (warp == 1)
    tmem_alloc(tmem_slot, TC_DW_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int pos = v.pos[v.base];
      for (int kb = kb_begin, it = 0; kb < kb_end; kb++, it++) {
        int s = it % TC_DW_STAGES;
        uint32_t ph = (it / TC_DW_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        int step = kb / kb_per_step;
        int b0 = (kb - step * kb_per_step) * TC_BK;
        int slot = pos - step;
        if (slot < 0)
          slot += v.depth;
        int xrow = slot * v.cap + v.base + b0;
        int erow = step * v.cap + v.base + b0;
        uint8_t *st = smem + s * DW_STAGE_BYTES;
        mbar_expect_tx(&full[s], DW_STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          tma_load_2d(&mXhi, &full[s], st + c * 4096, m0 + c * 32, xrow);
          tma_load_2d(&mXlo, &full[s], st + DW_OP_BYTES + c * 4096, m0 + c * 32, xrow);
          tma_load_2d(&mEhi, &full[s], st + 2 * DW_OP_BYTES + c * 4096, n0 + c * 32, erow);
          tma_load_2d(&mElo, &full[s], st + 3 * DW_OP_BYTES + c * 4096, n0 + c * 32, erow);
        }
      }
    }
  }
  else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(TC_BM, TC_DW_BN, 1, 1);
      for (int kb = kb_begin, it = 0; kb < kb_end; kb++, it++) {
        int s = it % TC_DW_STAGES;
        uint32_t ph = (it / TC_DW_STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        uint32_t a_hi = smem_u32(smem + s * DW_STAGE_BYTES);
        uint32_t a_lo = a_hi + DW_OP_BYTES;
        uint32_t b_hi = a_hi + 2 * DW_OP_BYTES;
        uint32_t b_lo = a_hi + 3 * DW_OP_BYTES;
#pragma unroll
        for (int kk = 0; kk < TC_BK / 8; kk++) {
          /* MN-major 32-bit operands (SWIZZLE_128B_BASE32B): each K row is one
             128-byte line of 32 floats along M/N; 32-float chunks along M/N
             are 4096 B apart (leading offset), groups of 4 K rows 512 B apart
             (stride offset); one MMA consumes 8 K rows = 1024 B */
          uint64_t dah = umma_desc(a_hi + kk * 1024, 4096, 512, UMMA_SW128_BASE32);
          uint64_t dal = umma_desc(a_lo + kk * 1024, 4096, 512, UMMA_SW128_BASE32);
          uint64_t dbh = umma_desc(b_hi + kk * 1024, 4096, 512, UMMA_SW128_BASE32);
          uint64_t dbl = umma_desc(b_lo + kk * 1024, 4096, 512, UMMA_SW128_BASE32);
          umma_tf32(tmem_base, dal, dbh, idesc, (it | kk) ? 1u : 0u);
          umma_tf32(tmem_base, dah, dbl, idesc, 1u);
          umma_tf32(tmem_base, dah, dbh, idesc, 1u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc_ready);
    }
  }
  else {
    const int q = warp & 3;
    const int y = m0 + q * 32 + lane;
    float acc[32];
    const bool any = kb_end > kb_begin;
    if (any) {
      mbar_wait(acc_ready, 0);
      tc_fence_after();
    }
    float *dst = g.partial + ((size_t)blockIdx.z * I + (y < I ? y : 0)) * H + n0;
#pragma unroll 1
    for (int c = 0; c < TC_DW_BN; c += 32) {
      if (any)
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, acc);
      else {
#pragma unroll
        for (int j = 0; j < 32; j++)
          acc[j] = 0.0f;
      }
      if (y < I) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (n0 + c + j < H)
            *(float4 *)(dst + c + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_DW_BN);
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

/* ======================================================================== */
/* host side                                                                  */

extern "C" int
rb_tc_usable(const RbView *v)
{
  return v->contiguous && v->n >= 64 && (v->n % TC_BK) == 0 && v->d.h_size >= 64;
}

static void
refresh_weight_planes(RbTc *t, RbPool *p, const RbView *v)
{
  RbGroup *g = p->group;
  if (t->w_src == v->Wih && t->w_version == g->weights_version)
    return;
  const int I = v->d.i_size, H = v->d.h_size;
  dim3 grid(cdiv(H, 32), cdiv(I, 32));
  k_split_weights<<<grid, 256, 0, rb_stream>>>(v->Wih, I, H, t->Whi, t->Wlo, t->WThi, t->WTlo);
  LAUNCH_CHECK("k_split_weights");
  t->w_src = v->Wih;
  t->w_version = g->weights_version;
}

static int nt_attr_done = 0, dw_attr_done = 0;

extern "C" void
rb_tc_forward(RbPool *p, const RbView *v, float presynaptic_noise)
{
  RbTc *t = tc_state(p);
  refresh_weight_planes(t, p, v);
  rbk_prepare_x(v);
  k_split_rows<<<v->n, 256, 0, rb_stream>>>(*v, 0, t->Xhi, t->Xlo);
  LAUNCH_CHECK("k_split_rows");
  NtArgs g;
  g.v = *v;
  g.mode = 0;
  g.k = 0;
  g.use_noise = 0;
  g.Ehi = g.Elo = NULL;
  if (presynaptic_noise != 0.0f) {
    rbk_gen_noise(v, presynaptic_noise, 1, v->d.h_size - 1);
    g.use_noise = 1;
  }
  if (!nt_attr_done) {
    CUDA_OR_DIE(cudaFuncSetAttribute(k_tc_nt, cudaFuncAttributeMaxDynamicSharedMemorySize,
            NT_SMEM_BYTES));
    nt_attr_done = 1;
  }
  dim3 grid(cdiv(v->d.h_size, TC_NT_BN), cdiv(v->n, TC_BM));
  rb_prof_begin(RB_PROF_FWD);
  k_tc_nt<<<grid, 192, NT_SMEM_BYTES, rb_stream>>>(t->mXhi_k, t->mXlo_k, t->mWThi_k, t->mWTlo_k, g);
  LAUNCH_CHECK("k_tc_nt<FWD>");
  rb_prof_end(RB_PROF_FWD);
  rbk_output(v);
}

extern "C" void
rb_tc_bptt(RbPool *p, const RbView *v, float *ih_delta, int accumulate)
{
  RbTc *t = tc_state(p);
  refresh_weight_planes(t, p, v);
  /* E[0] (written by k_top) -> planes */
  k_split_rows<<<v->n, 256, 0, rb_stream>>>(*v, 1, t->Ehi, t->Elo);
  LAUNCH_CHECK("k_split_rows");
  if (!nt_attr_done) {
    CUDA_OR_DIE(cudaFuncSetAttribute(k_tc_nt, cudaFuncAttributeMaxDynamicSharedMemorySize,
            NT_SMEM_BYTES));
    nt_attr_done = 1;
  }
  NtArgs g;
  g.v = *v;
  g.mode = 1;
  g.use_noise = 0;
  g.Ehi = t->Ehi;
  g.Elo = t->Elo;
  dim3 cgrid(cdiv(v->d.i_size, TC_NT_BN), cdiv(v->n, TC_BM));
  for (int k = 0; k < v->depth; k++) {
    g.k = k;
    rb_prof_begin(RB_PROF_CHAIN);
    k_tc_nt<<<cgrid, 192, NT_SMEM_BYTES, rb_stream>>>(t->mEhi_k, t->mElo_k, t->mWhi_k, t->mWlo_k, g);
    LAUNCH_CHECK("k_tc_nt<CHAIN>");
    rb_prof_end(RB_PROF_CHAIN);
    rbk_chain_decide(v, k);
  }
  dim3 fgrid(v->n, v->depth);
  k_finalize_rows<<<fgrid, 256, 0, rb_stream>>>(*v, t->Ehi, t->Elo);
  LAUNCH_CHECK("k_finalize_rows");
  if (!dw_attr_done) {
    CUDA_OR_DIE(cudaFuncSetAttribute(k_tc_dw, cudaFuncAttributeMaxDynamicSharedMemorySize,
            DW_SMEM_BYTES));
    dw_attr_done = 1;
  }
  DwArgs d;
  d.v = *v;
  d.partial = t->partial;
  dim3 dgrid(cdiv(v->d.h_size, TC_DW_BN), cdiv(v->d.i_size, TC_BM), TC_DW_SPLITS);
  rb_prof_begin(RB_PROF_DW);
  k_tc_dw<<<dgrid, 192, DW_SMEM_BYTES, rb_stream>>>(t->mXhi_mn, t->mXlo_mn, t->mEhi_mn, t->mElo_mn, d);
  LAUNCH_CHECK("k_tc_dw");
  int size = v->d.i_size * v->d.h_size;
  k_dw_reduce<<<cdiv(size / 4, 256), 256, 0, rb_stream>>>(ih_delta, t->partial, size,
      TC_DW_SPLITS, accumulate);
  LAUNCH_CHECK("k_dw_reduce");
  rb_prof_end(RB_PROF_DW);
}
