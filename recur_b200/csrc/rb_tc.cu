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
#include "rb_optim.cuh"
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
tma_load_3d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2)
{
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
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

/* ---- CTA-pair (cta_group::2) variants ---------------------------------- */

__device__ __forceinline__ uint32_t
cluster_ctarank(void)
{
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void
cluster_sync_all(void)
{
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

/* shared-window addresses of the two CTAs of a pair differ in this bit;
   clearing it names the even (leader) CTA's copy of an object */
#define PAIR_LEADER_MASK 0xFEFFFFFFu

/* TMA load into this CTA's shared memory, completion bytes counted on the
   LEADER CTA's mbarrier */
__device__ __forceinline__ void
tma_load_3d_pair(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2)
{
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & PAIR_LEADER_MASK), "r"(c0),
      "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void
tmem_alloc_pair(uint32_t *dst_smem, uint32_t cols)
{
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
      ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void
tmem_dealloc_pair(uint32_t addr, uint32_t cols)
{
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
      : "memory");
}

/* one MMA across both SMs of the pair: M = 256 (128 rows per CTA), each CTA
   holding half of B's N columns; issued by one thread of the leader CTA */
__device__ __forceinline__ void
umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
    uint32_t accumulate)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

/* arrive on the mbarrier at this shared-memory offset in BOTH CTAs of the pair
   when all MMAs issued so far have completed */
__device__ __forceinline__ void
umma_commit_pair(uint64_t *bar)
{
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
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

/* the same load without the wait, so that the next one can be in flight while
   this one's registers are being stored; tmem_wait_ld() before touching them */
__device__ __forceinline__ void
tmem_ld32_nowait(uint32_t taddr, float *v)
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
}

__device__ __forceinline__ void
tmem_wait_ld(void)
{
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

/* 32 bytes per lane per instruction: whole L2 sectors even when every lane
   writes its own row */
__device__ __forceinline__ void
st_global_v8(float *p, const float *a)
{
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"l"(p), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]),
      "f"(a[7]) : "memory");
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
#define TC_BK 32       /* K per stage: 32 floats = one 128-byte swizzle row */
#define TC_FWD_BN 64      /* FWD tile columns */
#define TC_FWD_STAGES 4
#define TC_CHAIN_BN 128   /* CHAIN tile columns */
#define TC_CHAIN_STAGES 3
#define TC_DW2_BN 192 /* N of the CTA-pair weight-gradient tile (two halves of 96) */
#define TC_CHAIN_SPLITS 4 /* split-K: 4 x 9 x 4 = 144 CTAs at 512 streams, H1023 */
#define TC_DW_BN 256      /* DW tile columns */
#define TC_DW_BK 16       /* DW ring rows per stage */
#define TC_DW_STAGES 4
#define TC_DW_SPLITS 4

typedef struct RbTc {
  int cap, depth;
  float *Xhi, *Xlo;     /* [depth][cap][i_size]   planes of the ring */
  float *Ehi, *Elo;     /* [depth+1][cap][i_size] planes of the error chain */
  float *Whi, *Wlo;     /* [i_size][h_size] */
  float *WThi, *WTlo;   /* [h_size][i_size] */
  float *partial;       /* [TC_DW_SPLITS][i_size][h_size] */
  float *cpartial;      /* [TC_CHAIN_SPLITS][cap][i_size rounded up to 32] split-K partial sums */
  unsigned int *sync;   /* grid barrier counter + per-step live counts of the persistent chain */
  size_t sync_words;
  int sync_flip;
  int persistent_ok;    /* decided per call */
  int delta_pending;    /* the last weight gradient still sits in `partial`, unsummed */
  int pending_accumulate;
  const float *w_src;   /* weights the planes were made from */
  uint64_t w_version;
  /* tensor maps */
  CUtensorMap mXhi_k, mXlo_k;   /* ring rows as K-major A of FWD: box 32 x 128 */
  CUtensorMap mEhi_k, mElo_k;   /* error rows as K-major A of CHAIN (width h_size) */
  CUtensorMap mWhi_k, mWlo_k;   /* Wih rows as K-major B of CHAIN: box 32 x 128 */
  CUtensorMap mWThi_k, mWTlo_k; /* Wih^T rows as K-major B of FWD: box 32 x 64 */
  CUtensorMap mWThi_k128, mWTlo_k128; /* the same with 128-row boxes: split-K FWD */
  CUtensorMap mXhi_mn, mXlo_mn; /* ring rows as MN-major A of DW: box 32 x 32 */
  CUtensorMap mEhi_mn, mElo_mn; /* error rows as MN-major B of DW (width h_size) */
  CUtensorMap mEhi_mn4, mElo_mn4; /* error rows as MN-major A of the pair DW: 4 chunks */
  CUtensorMap mXhi_mn3, mXlo_mn3; /* ring rows as MN-major B half of the pair DW: 3 chunks */
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

/* The same rows seen as [chunk of 32 columns][row][32 floats], so that one
   TMA fetches n_chunks column chunks of box_rows rows into consecutive
   (chunk-major) shared memory: the MN-major operand layout of DW.  A chunk
   that straddles the end of a row reads on into the next row; those columns
   only ever feed output rows/columns that are discarded. */
static void
make_map_chunked(CUtensorMap *m, float *base, uint64_t width, uint64_t rows, uint64_t pitch,
    uint32_t box_rows, uint32_t n_chunks)
{
  cuuint64_t dims[3] = {32, rows, (width + 31) / 32};
  cuuint64_t strides[2] = {pitch * sizeof(float), 32 * sizeof(float)};
  cuuint32_t box[3] = {32, box_rows, n_chunks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box,
      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
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
  cudaFree(t->partial);
  cudaFree(t->cpartial);
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
  size_t ring = (size_t)p->depth * p->cap * I, chain = (size_t)(p->depth + 1) * p->cap * I;
  t->Xhi = dmalloc0<float>(ring + 64);
  t->Xlo = dmalloc0<float>(ring + 64);
  t->Ehi = dmalloc0<float>(chain + 64);
  t->Elo = dmalloc0<float>(chain + 64);
  t->Whi = dmalloc0<float>(I * H);
  t->Wlo = dmalloc0<float>(I * H);
  t->WThi = dmalloc0<float>(I * H);
  t->WTlo = dmalloc0<float>(I * H);
  t->partial = dmalloc0<float>((size_t)TC_DW_SPLITS * I * H);
  t->cpartial = dmalloc0<float>((size_t)TC_CHAIN_SPLITS * p->cap * ((I + 31) & ~(size_t)31));
  /* two areas, used by alternate walks (see k_finalize_rows) */
  t->sync_words = (size_t)(cdiv(p->cap, TC_BM) + 1) * (p->depth + 8) + 8;
  t->sync = dmalloc0<unsigned int>(2 * t->sync_words);
  t->sync_flip = 0;
  t->persistent_ok = -1;
  t->w_src = NULL;
  uint64_t ring_rows = (uint64_t)p->depth * p->cap, chain_rows = (uint64_t)(p->depth + 1) * p->cap;
  make_map(&t->mXhi_k, t->Xhi, I, ring_rows, I, TC_BM);
  make_map(&t->mXlo_k, t->Xlo, I, ring_rows, I, TC_BM);
  make_map(&t->mEhi_k, t->Ehi, H, chain_rows, I, TC_BM);
  make_map(&t->mElo_k, t->Elo, H, chain_rows, I, TC_BM);
  make_map(&t->mWhi_k, t->Whi, H, I, H, TC_CHAIN_BN);
  make_map(&t->mWlo_k, t->Wlo, H, I, H, TC_CHAIN_BN);
  make_map(&t->mWThi_k, t->WThi, I, H, I, TC_FWD_BN);
  make_map(&t->mWTlo_k, t->WTlo, I, H, I, TC_FWD_BN);
  make_map(&t->mWThi_k128, t->WThi, I, H, I, TC_CHAIN_BN);
  make_map(&t->mWTlo_k128, t->WTlo, I, H, I, TC_CHAIN_BN);
  make_map_chunked(&t->mXhi_mn, t->Xhi, I, ring_rows, I, TC_DW_BK, TC_BM / 32);
  make_map_chunked(&t->mXlo_mn, t->Xlo, I, ring_rows, I, TC_DW_BK, TC_BM / 32);
  make_map_chunked(&t->mEhi_mn, t->Ehi, H, chain_rows, I, TC_DW_BK, TC_DW_BN / 32);
  make_map_chunked(&t->mElo_mn, t->Elo, H, chain_rows, I, TC_DW_BK, TC_DW_BN / 32);
  make_map_chunked(&t->mEhi_mn4, t->Ehi, H, chain_rows, I, TC_DW_BK, TC_BM / 32);
  make_map_chunked(&t->mElo_mn4, t->Elo, H, chain_rows, I, TC_DW_BK, TC_BM / 32);
  make_map_chunked(&t->mXhi_mn3, t->Xhi, I, ring_rows, I, TC_DW_BK, TC_DW2_BN / 2 / 32);
  make_map_chunked(&t->mXlo_mn3, t->Xlo, I, ring_rows, I, TC_DW_BK, TC_DW2_BN / 2 / 32);
  t->dw_splits = TC_DW_SPLITS;
  p->tc = t;
  p->x_planes_stale = 2; /* the ring may hold rows from before the planes existed */
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

/* every row of the ring -> its planes (after something other than the tensor
   engine's own forward pass rewrote ring rows: RbPool.x_planes_stale) */
__global__ void __launch_bounds__(256)
k_split_ring(const float *__restrict__ X, size_t n, float *hi_plane, float *lo_plane)
{
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n;
       i += (size_t)gridDim.x * blockDim.x * 4) {
    float4 x = *(const float4 *)(X + i);
    float4 h, l;
    split_tf32(x.x, h.x, l.x);
    split_tf32(x.y, h.y, l.y);
    split_tf32(x.z, h.z, l.z);
    split_tf32(x.w, h.w, l.w);
    *(float4 *)(hi_plane + i) = h;
    *(float4 *)(lo_plane + i) = l;
  }
}

/* After the walk: rows of E beyond a stream's executed depth must not reach
   the weight gradient (zero them), and a stream whose gradient is clipped
   (ih_scale != 1, recur-nn.c:393-402) has its rows rescaled and re-split. */
__global__ void __launch_bounds__(256)
k_finalize_rows(RbView v, float *Ehi, float *Elo, const unsigned int *kmax_dev,
    unsigned int *zero, int zero_words)
{
  /* the other of the two barrier/counter areas is cleared here for the next
     walk, which saves that walk a memset in front of its kernel */
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < zero_words;
       i += gridDim.x * blockDim.x)
    zero[i] = 0u;
  const int s = v.slots[blockIdx.x];
  const RbScalars sc = v.sc[s];
  const int kmax = min((int)*kmax_dev, v.depth);
  /* rows this stream never reached, up to the deepest step any stream took
     (the weight gradient stops there) */
  for (int step = sc.n_steps; step < kmax; step++) {
    size_t off = ((size_t)step * v.cap + s) * v.d.i_size;
    for (int i = threadIdx.x * 4; i < v.d.h_size; i += blockDim.x * 4) {
      *(float4 *)(Ehi + off + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      *(float4 *)(Elo + off + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (sc.ih_scale != 1.0f) {
    for (int step = 0; step < sc.n_steps; step++) {
      size_t off = ((size_t)step * v.cap + s) * v.d.i_size;
      for (int i = threadIdx.x; i < v.d.h_size; i += blockDim.x) {
        float hi, lo;
        split_tf32(v.E[off + i] * sc.ih_scale, hi, lo);
        Ehi[off + i] = hi;
        Elo[off + i] = lo;
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
/* FWD and CHAIN: C[128 x BN] tiles, both operands K-major.
 *
 * FWD  (BN 64): activation epilogue straight from tensor memory.
 * CHAIN (BN 128, split-K over blockIdx.z): one BPTT step is too small to
 * fill 148 SMs with tiles the tensor pipe likes, so K is split four ways and
 * each CTA writes its raw partial tile; k_chain_finish_step sums the
 * partials and does the row-wise part of the step.                          */

struct NtArgs {
  RbView v;
  int mode;        /* 0 FWD, 1 CHAIN, 2 FWD split-K (raw partial sums) */
  int k;           /* CHAIN: step */
  int use_noise;
  float *cpartial; /* CHAIN: [splits][cap][i_size] */
};

template <int BN, int STAGES>
struct NtCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;
  static constexpr int B_BYTES = BN * TC_BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

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
  const int n_kb_total = (K + TC_BK - 1) / TC_BK;
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
    tmem_alloc(tmem_slot, BN);
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
        tma_load_2d(&mAhi, &full[s], st, kb * TC_BK, ring_row);
        tma_load_2d(&mAlo, &full[s], st + Cfg::A_BYTES, kb * TC_BK, ring_row);
        tma_load_2d(&mBhi, &full[s], st + 2 * Cfg::A_BYTES, kb * TC_BK, n0);
        tma_load_2d(&mBlo, &full[s], st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, kb * TC_BK, n0);
      }
    }
  }
  else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(TC_BM, BN, 0, 0);
      for (int it = 0; it < n_kb; it++) {
        int s = it % STAGES;
        uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        uint32_t a_hi = smem_u32(smem + s * Cfg::STAGE_BYTES);
        uint32_t a_lo = a_hi + Cfg::A_BYTES;
        uint32_t b_hi = a_hi + 2 * Cfg::A_BYTES;
        uint32_t b_lo = b_hi + Cfg::B_BYTES;
#pragma unroll
        for (int kk = 0; kk < TC_BK / 8; kk++) {
          /* K-major, 128B swizzle: 8-row groups 1024 B apart; a K step of 8
             floats moves the start address by 32 bytes inside the swizzle row */
          uint64_t dah = umma_desc(a_hi + kk * 32, 16, 1024);
          uint64_t dal = umma_desc(a_lo + kk * 32, 16, 1024);
          uint64_t dbh = umma_desc(b_hi + kk * 32, 16, 1024);
          uint64_t dbl = umma_desc(b_lo + kk * 32, 16, 1024);
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
    float acc[32];
    if (g.mode == 0) {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
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
      /* raw partial sums of this K split; masks and the rest happen row-wise
         in k_chain_finish_step */
      const bool live = row_ok && (g.mode == 2 || v.sc[sidx].live != 0);
      const int cpitch = (I + 31) & ~31;
      const int n_cols = (g.mode == 2) ? H : I;
      float *dst = g.cpartial + ((size_t)blockIdx.z * v.cap + sidx) * cpitch;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (n_kb > 0)
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, acc);
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
    tmem_dealloc(tmem_base, BN);
  }
}

/* The row-wise half of a BPTT step (recur-nn.c:338-389, 393-413), one block
   per stream: sum the split-K partials in a fixed order, mask by the input
   that fed each row, ReSQRT derivative, write E(k+1) with its hi/lo planes,
   sum the squares, and decide whether this stream walks further.           */
__global__ void __launch_bounds__(256)
k_chain_finish_step(RbView v, int k, const float *__restrict__ cpartial, int splits,
    float *__restrict__ Ehi, float *__restrict__ Elo)
{
  __shared__ float scratch[33];
  const int s = v.slots[blockIdx.x];
  RbScalars *scp = v.sc + s;
  RbScalars scv = *scp; /* one read up front; thread 0 writes the changes back */
  RbScalars *sc = &scv;
  if (!sc->live)
    return;
  const int I = v.d.i_size, H = v.d.h_size, hs1 = v.d.hidden_size + 1;
  int p = v.pos[s] - k;
  if (p < 0)
    p += v.depth;
  const float *xk = v.X + ((size_t)p * v.cap + s) * I;
  const size_t eoff = ((size_t)(k + 1) * v.cap + s) * I;
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
    float4 xin = *(const float4 *)(xk + c);
    float acc[4] = {a.x, a.y, a.z, a.w};
    float xi[4] = {xin.x, xin.y, xin.z, xin.w};
    float o[4], ohi[4], olo[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      float e = 0.0f;
      float input = xi[u];
      if (input != 0.0f && (v.activation != RNN_RECLIP20 || input < 20.0f)) {
        e = acc[u];
        if (v.activation == RNN_RESQRT)
          e /= 2.0f * (input + 1.0f);
        sq += e * e;
      }
      int col = c + u;
      if (v.CIE && col >= hs1 && col < hs1 + v.d.input_size)
        v.CIE[(size_t)s * v.bl_o + col - hs1] += e;
      if (col == 0 || (col >= hs1 && col < H))
        e = 0.0f;
      o[u] = e;
      split_tf32(e, ohi[u], olo[u]);
    }
    *(float4 *)(v.E + eoff + c) = make_float4(o[0], o[1], o[2], o[3]);
    *(float4 *)(Ehi + eoff + c) = make_float4(ohi[0], ohi[1], ohi[2], ohi[3]);
    *(float4 *)(Elo + eoff + c) = make_float4(olo[0], olo[1], olo[2], olo[3]);
  }
  float es = block_sum_tc(sq, scratch);
  if (threadIdx.x != 0)
    return;
  sc->err_sum = es;
  sc->cum_error += sqrtf(es);
  sc->n_steps = k + 1;
  int t = v.depth - k;
  bool stop = (es <= sc->min_sum || es > sc->max_sum);
  bool last = (k == v.depth - 1);
  if (!stop && !last) {
    *scp = scv;
    return;
  }
  sc->live = 0;
  int t_left = stop ? t : 0;
  sc->t_left = t_left;
  float ceiling = ERROR_GAIN_CEILING * sc->top_scaled;
  if (es > ceiling) {
    float halfmax = sc->max_sum;
    float x = es / halfmax;
    float fudge = (float)(0.99 + (double)(x * x / 100.0f));
    sc->ih_scale = (halfmax == 0.0f) ? es : 2.0f * x / (1.0f + x * x * fudge);
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
  *scp = scv;
}

/* ======================================================================== */
/* The whole BPTT walk as ONE persistent, grid-synchronised kernel.
 *
 * A BPTT step is a small dependent GEMM; launched per step it spends more
 * time on launch gaps, pipeline prologues and cold tensor-map fetches than on
 * MMAs.  Here 144 CTAs (n-tiles x m-tiles x K-splits, one per SM, co-resident
 * through a cooperative launch) stay alive for all `depth` steps.  Per step:
 *   phase A  the split-K tcgen05 GEMM of k_tc_nt (TMEM allocated once, mbarrier
 *            pipeline state carried across steps), partial tiles to L2;
 *   grid barrier;
 *   phase B  the row-wise half (k_chain_finish_step's body), one row per
 *            epilogue warp across the grid, which also counts the streams
 *            that walk on;
 *   grid barrier; stop when no stream is left.
 * E(k+1)'s hi/lo planes are written with generic stores and read by the next
 * step's TMA, hence the generic->async proxy fence before the barrier.       */

__device__ __forceinline__ unsigned int
ld_acquire_gpu(const unsigned int *p)
{
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

/* per-stream scalars change between steps inside one launch: read them past L1 */
__device__ __forceinline__ RbScalars
load_scalars_cg(const RbScalars *p)
{
  static_assert(sizeof(RbScalars) == 64, "RbScalars is read as four 16-byte words");
  union { RbScalars s; int4 w[4]; } u;
  const int4 *src = (const int4 *)p;
#pragma unroll
  for (int i = 0; i < 4; i++)
    u.w[i] = __ldcg(src + i);
  return u.s;
}

struct ChainArgs {
  RbView v;
  float *cpartial;
  float *Ehi, *Elo;
  unsigned int *sync;   /* [0] barrier counter, [1 + k] streams alive at step k */
  unsigned long long *dbg; /* optional: 5 globaltimer stamps per step from CTA 0 */
  unsigned int *kmax;   /* out: steps executed before every stream had stopped */
};

__device__ __forceinline__ unsigned long long
globaltimer_ns(void)
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

#define ROLE_STAMP(slot) do {                                           \
    if (g.dbg && cta == 0 && k == 5)                                    \
      g.dbg[1200 + (slot)] = globaltimer_ns();                          \
  } while (0)

#define CHAIN_STAMP(slot) do {                                          \
    if (g.dbg && threadIdx.x == 0) {                                    \
      unsigned long long now_ = globaltimer_ns();                       \
      if (cta == 0)                                                     \
        g.dbg[k * 5 + (slot)] = now_;                                   \
      if (k == 5)                                                       \
        g.dbg[320 + cta * 5 + (slot)] = now_;                           \
    }                                                                   \
  } while (0)

__device__ __forceinline__ void
grid_barrier(unsigned int *counter, unsigned int target)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    /* arrive with a release reduction: nothing waits for the old value to
       come back before the polling starts */
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
    /* a poll is an L2 round trip (~0.7 us): 2^22 of them are seconds, against
       the microseconds a healthy barrier takes.  A grid that is not fully
       resident can never complete the barrier; trap (the host then aborts
       with the launch failure) rather than hang the device. */
    unsigned int spins = 0;
    while (ld_acquire_gpu(counter) < target) {
      if (++spins > (1u << 22))
        __trap();
    }
  }
  __syncthreads();
}

/* Four columns of E(k+1) from the summed partials `a` and the ring row `xin`
   they are masked with; PLAIN is the common case (ReLU, no bottom layer),
   kept free of branches.  Adds the squares to `sq`. */
template <bool PLAIN>
__device__ __forceinline__ void
chain_chunk(const RbView &v, float4 a, float4 xin, int c, int s, float &sq, float *e_out,
    float *hi_out, float *lo_out)
{
  const int H = v.d.h_size, hs1 = v.d.hidden_size + 1;
  float av[4] = {a.x, a.y, a.z, a.w};
  float xi[4] = {xin.x, xin.y, xin.z, xin.w};
  float o[4], ohi[4], olo[4];
  /* column 0 (the bias) and the padding between the hidden units and h_size
     carry no error */
  const bool edge = (c == 0) || (c + 4 > hs1 && c < H);
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
  if (edge) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      int col = c + u;
      if (col == 0 || (col >= hs1 && col < H))
        o[u] = 0.0f;
    }
  }
#pragma unroll
  for (int u = 0; u < 4; u++)
    split_tf32(o[u], ohi[u], olo[u]);
  __stcg((float4 *)e_out, make_float4(o[0], o[1], o[2], o[3]));
  __stcg((float4 *)hi_out, make_float4(ohi[0], ohi[1], ohi[2], ohi[3]));
  __stcg((float4 *)lo_out, make_float4(olo[0], olo[1], olo[2], olo[3]));
}

#define TC_CHAIN_THREADS 256 /* TMA warp, MMA warp, four epilogue warps, two more for the rows */

__device__ __forceinline__ void
named_bar_sync(int id, int count)
{
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_CHAIN_THREADS, 1)
k_tc_chain_persistent(const __grid_constant__ CUtensorMap mAhi,
    const __grid_constant__ CUtensorMap mAlo, const __grid_constant__ CUtensorMap mBhi,
    const __grid_constant__ CUtensorMap mBlo, ChainArgs g)
{
  using Cfg = NtCfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const RbView &v = g.v;
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *full = (uint64_t *)(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_ready = empty + STAGES;
  uint32_t *tmem_slot = (uint32_t *)(acc_ready + 1);
  float *pair_sq = (float *)(tmem_slot + 2); /* one sum of squares per warp */

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = v.d.i_size, H = v.d.h_size;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  /* The streams of one m-tile form an independent chain: only the CTAs that
     share blockIdx.y exchange data, so they synchronise among themselves and
     the four groups drift apart, one group's latency-bound row phase hiding
     under another group's loads. */
  const int grp = blockIdx.y;
  const int grp_ctas = gridDim.x * gridDim.z;
  const int grp_cta = blockIdx.z * gridDim.x + blockIdx.x;
  const int sync_stride = v.depth + 8;
  unsigned int *gsync = g.sync + (size_t)grp * sync_stride;
  const int n_kb_total = (H + TC_BK - 1) / TC_BK;
  const int kb_per_split = (n_kb_total + gridDim.z - 1) / gridDim.z;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(n_kb_total, kb_begin + kb_per_split);
  const int n_kb = max(0, kb_end - kb_begin);
  const int cpitch = (I + 31) & ~31; /* rows of the partial planes start on 128-byte lines */
  const size_t split_stride = (size_t)v.cap * cpitch;

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
    tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int pos0 = v.pos[v.base]; /* the batch advances in lockstep */
  const bool plain = (v.activation == RNN_RELU && v.CIE == NULL);
  unsigned int it = 0;       /* pipeline iterations so far (producer and MMA keep equal counts) */
  unsigned int n_gemms = 0;  /* accumulators completed by this CTA */
  unsigned int n_bar = 0;

  for (int k = 0; k < v.depth; k++) {
    /* ---- phase A: partial tile of E(k) . Wih^T over this CTA's K range ---- */
    CHAIN_STAMP(0);
    /* no per-tile liveness test here: it would put an L2 round trip in front
       of the first TMA; dead rows are simply not stored, and the walk ends
       for the whole grid once no stream is left */
    const bool tile_alive = n_kb > 0;
    if (threadIdx.x == 0)
      ROLE_STAMP(0);

    /* What the row phase needs and the GEMM does not produce - the stream's
       scalars and the ring row its errors are masked with - is fetched now,
       ahead of the barrier the partial sums have to wait for. */
    constexpr int GRP = 5; /* 5 x 256 columns: this warp's half of a row at H1023 in one
                              round of loads */
    /* wide rows (H1023: 1068 columns) take a pair of warps each; rows that one
       warp covers in a single round of loads (small nets: few CTAs per group,
       several rows per slot) take one warp each, twice as many at a time */
    const int wpr = (I <= 128 * GRP) ? 1 : 2;
    const int CSTEP = 128 * wpr;
    const int slot = warp / wpr, half = warp % wpr;
    const int slots_per_cta = (TC_CHAIN_THREADS / 32) / wpr;
    const int gw = grp_cta * slots_per_cta + slot;
    RbScalars sc_pre;
    float4 x_pre[GRP];
    sc_pre.live = 0;
    if (gw < TC_BM && m0 + gw < v.n) {
      const int s = v.base + m0 + gw;
      sc_pre = load_scalars_cg(v.sc + s);
      int p = pos0 - k;
      if (p < 0)
        p += v.depth;
      const float *xk = v.X + ((size_t)p * v.cap + s) * I;
#pragma unroll
      for (int i = 0; i < GRP; i++) {
        int c = half * 128 + lane * 4 + CSTEP * i;
        if (c < I)
          x_pre[i] = __ldg((const float4 *)(xk + c));
      }
    }

    if (tile_alive) {
      if (warp == 0) {
        if (lane == 0) {
          const int ring_row = k * v.cap + v.base + m0;
          for (int j = 0; j < n_kb; j++) {
            unsigned int i2 = it + j;
            int s = i2 % STAGES;
            uint32_t ph = (i2 / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t *st = smem + s * Cfg::STAGE_BYTES;
            int kb = kb_begin + j;
            mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
            tma_load_2d(&mAhi, &full[s], st, kb * TC_BK, ring_row);
            tma_load_2d(&mAlo, &full[s], st + Cfg::A_BYTES, kb * TC_BK, ring_row);
            tma_load_2d(&mBhi, &full[s], st + 2 * Cfg::A_BYTES, kb * TC_BK, n0);
            tma_load_2d(&mBlo, &full[s], st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, kb * TC_BK, n0);
          }
        }
      }
      else if (warp == 1) {
        if (lane == 0) {
          const uint32_t idesc = umma_idesc_tf32(TC_BM, BN, 0, 0);
          for (int j = 0; j < n_kb; j++) {
            unsigned int i2 = it + j;
            int s = i2 % STAGES;
            uint32_t ph = (i2 / STAGES) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (j == 0)
              ROLE_STAMP(1);
            if (j == n_kb - 1)
              ROLE_STAMP(2);
            uint32_t a_hi = smem_u32(smem + s * Cfg::STAGE_BYTES);
            uint32_t a_lo = a_hi + Cfg::A_BYTES;
            uint32_t b_hi = a_hi + 2 * Cfg::A_BYTES;
            uint32_t b_lo = b_hi + Cfg::B_BYTES;
#pragma unroll
            for (int kk = 0; kk < TC_BK / 8; kk++) {
              uint64_t dah = umma_desc(a_hi + kk * 32, 16, 1024);
              uint64_t dal = umma_desc(a_lo + kk * 32, 16, 1024);
              uint64_t dbh = umma_desc(b_hi + kk * 32, 16, 1024);
              uint64_t dbl = umma_desc(b_lo + kk * 32, 16, 1024);
              umma_tf32(tmem_base, dal, dbh, idesc, (j | kk) ? 1u : 0u);
              umma_tf32(tmem_base, dah, dbl, idesc, 1u);
              umma_tf32(tmem_base, dah, dbh, idesc, 1u);
            }
            umma_commit(&empty[s]);
          }
          umma_commit(acc_ready);
        }
      }
      else if (warp < 6) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int m = m0 + row;
        const bool row_ok = m < v.n;
        const int sidx = v.base + (row_ok ? m : 0);
        const bool live = row_ok && __ldcg(&v.sc[sidx].live) != 0;
        mbar_wait(acc_ready, n_gemms & 1);
        tc_fence_after();
        if (threadIdx.x == 64)
          ROLE_STAMP(3);
        float *dst = g.cpartial + (size_t)blockIdx.z * split_stride + (size_t)sidx * cpitch;
        float acc[2][32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        tmem_ld32_nowait(taddr, acc[0]);
#pragma unroll
        for (int ci = 0; ci < BN / 32; ci++) {
          tmem_wait_ld();
          if (ci + 1 < BN / 32)
            tmem_ld32_nowait(taddr + (ci + 1) * 32, acc[(ci + 1) & 1]);
          const float *av = acc[ci & 1];
          int col0 = n0 + ci * 32;
          if (!live || col0 >= I)
            continue;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j + 8 <= I)
              st_global_v8(dst + col0 + j, av + j);
            else if (col0 + j < I)
              __stcg((float4 *)(dst + col0 + j),
                  make_float4(av[j], av[j + 1], av[j + 2], av[j + 3]));
          }
        }
        tc_fence_before();
        if (threadIdx.x == 64)
          ROLE_STAMP(4);
      }
      it += n_kb;
      n_gemms++;
    }
    __syncthreads();
    CHAIN_STAMP(1);
    n_bar++;
    grid_barrier(gsync, n_bar * grp_ctas);
    CHAIN_STAMP(2);

    /* ---- phase B: rows of E(k+1), one per PAIR of warps across the group
       (all eight warps take part; each warp of a pair takes every other
       128-column chunk) ---- */
    {
      for (int r = gw; r < TC_BM; r += grp_ctas * slots_per_cta) {
        const int m = m0 + r;
        if (m >= v.n)
          break;
        const int s = v.base + m;
        RbScalars *scp = v.sc + s;
        /* the scalars travel with the row loads: a stopped stream's loads
           are wasted, nothing of it is stored */
        if (warp == 2 && lane == 0) ROLE_STAMP(10);
        const bool pre = (r == gw);
        RbScalars sc = pre ? sc_pre : load_scalars_cg(scp);
        int p = pos0 - k;
        if (p < 0)
          p += v.depth;
        const float *xk = v.X + ((size_t)p * v.cap + s) * I;
        const size_t eoff = ((size_t)(k + 1) * v.cap + s) * I;
        const float *part = g.cpartial + (size_t)s * cpitch;
        float sq = 0.0f;
        /* all loads of a group of column chunks are issued before any of
           their results is used or stored: the row costs a few L2 round
           trips instead of one per chunk */
        for (int c0 = half * 128 + lane * 4; c0 < I; c0 += CSTEP * GRP) {
          const bool x_here = pre && c0 < CSTEP; /* the first round of the first row */
          float4 a[GRP], xin[GRP], pz[TC_CHAIN_SPLITS - 1][GRP];
#pragma unroll
          for (int i = 0; i < GRP; i++) {
            int c = c0 + CSTEP * i;
            if (c < I) {
              a[i] = __ldcg((const float4 *)(part + c));
              xin[i] = x_here ? x_pre[i] : __ldg((const float4 *)(xk + c));
#pragma unroll
              for (int z = 1; z < TC_CHAIN_SPLITS; z++)
                if (z < (int)gridDim.z)
                  pz[z - 1][i] = __ldcg((const float4 *)(part + z * split_stride + c));
            }
          }
          if (!sc.live)
            break;
          if (warp == 2 && lane == 0) ROLE_STAMP(11);
#pragma unroll
          for (int z = 1; z < TC_CHAIN_SPLITS; z++) {
            if (z < (int)gridDim.z) {
#pragma unroll
              for (int i = 0; i < GRP; i++) {
                int c = c0 + CSTEP * i;
                if (c < I) {
                  a[i].x += pz[z - 1][i].x; a[i].y += pz[z - 1][i].y;
                  a[i].z += pz[z - 1][i].z; a[i].w += pz[z - 1][i].w;
                }
              }
            }
          }
          if (plain) {
#pragma unroll
            for (int i = 0; i < GRP; i++) {
              int c = c0 + CSTEP * i;
              if (c < I)
                chain_chunk<true>(v, a[i], xin[i], c, s, sq, v.E + eoff + c, g.Ehi + eoff + c,
                    g.Elo + eoff + c);
            }
          }
          else {
#pragma unroll
            for (int i = 0; i < GRP; i++) {
              int c = c0 + CSTEP * i;
              if (c < I)
                chain_chunk<false>(v, a[i], xin[i], c, s, sq, v.E + eoff + c, g.Ehi + eoff + c,
                    g.Elo + eoff + c);
            }
          }
        }
        if (!sc.live)
          continue;
        if (warp == 2 && lane == 0) ROLE_STAMP(12);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (wpr == 2) {
          if (lane == 0)
            pair_sq[warp] = sq;
          named_bar_sync(1 + slot, 64);
        }
        if (warp == 2 && lane == 0) ROLE_STAMP(13);
        if (half == 0 && lane == 0) {
          float es = (wpr == 2) ? pair_sq[warp] + pair_sq[warp + 1] : sq;
          sc.err_sum = es;
          sc.cum_error += sqrtf(es);
          sc.n_steps = k + 1;
          int t = v.depth - k;
          bool stop = (es <= sc.min_sum || es > sc.max_sum);
          bool last = (k == v.depth - 1);
          if (stop || last) {
            sc.live = 0;
            int t_left = stop ? t : 0;
            sc.t_left = t_left;
            float ceiling = ERROR_GAIN_CEILING * sc.top_scaled;
            if (es > ceiling) {
              float halfmax = sc.max_sum;
              float x = es / halfmax;
              float fudge = (float)(0.99 + (double)(x * x / 100.0f));
              sc.ih_scale = (halfmax == 0.0f) ? es : 2.0f * x / (1.0f + x * x * fudge);
            }
            else {
              sc.ih_scale = 1.0f;
              if (sc.adaptive & 1) {
                int depth_error = v.depth / 4 - t_left;
                float min_gain = MIN_ERROR_GAIN * sc.top_scaled;
                float mef = sc.mef;
                if (mef < MAX_MIN_ERROR_FACTOR && (min_gain != sc.min_sum || depth_error < 0))
                  mef = (float)((double)mef * (1.0 + depth_error * 1e-3));
                sc.mef = fmaxf(mef, ABS_MIN_ERROR_FACTOR);
              }
            }
          }
          else {
            atomicAdd(&gsync[1 + k], 1u);
          }
          *scp = sc;
        }
        if (wpr == 2)
          named_bar_sync(1 + slot, 64); /* pair_sq may be rewritten */
      }
      if (warp == 2 && lane == 0) ROLE_STAMP(14);
      /* E(k+1) planes were written through the generic proxy; the next step's
         TMA reads them through the async proxy */
      asm volatile("fence.proxy.async;" ::: "memory");
      if (warp == 2 && lane == 0) ROLE_STAMP(15);
    }
    __syncthreads();
    CHAIN_STAMP(3);
    n_bar++;
    grid_barrier(gsync, n_bar * grp_ctas);
    CHAIN_STAMP(4);
    if (__ldcg(&gsync[1 + k]) == 0) {
      if (grp_cta == 0 && threadIdx.x == 0)
        atomicMax(g.kmax, (unsigned int)(k + 1));
      break; /* every stream of this group has stopped (uniform across the group) */
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

/* ======================================================================== */
/* DW: delta tile [128 y x 256 x] += X^T . E over (step, stream) rows;
   both operands MN-major, 16 ring rows per stage, split-K across blockIdx.z  */

struct DwArgs {
  RbView v;
  float *partial; /* [splits][i_size][h_size] */
  const unsigned int *kmax; /* deepest step any stream executed: rows beyond contribute nothing */
};

#define DW_CHUNK_BYTES (32 * TC_DW_BK * 4)          /* one TMA box: 32 floats x 16 rows */
#define DW_A_BYTES ((TC_BM / 32) * DW_CHUNK_BYTES)    /* 8 KB */
#define DW_B_BYTES ((TC_DW_BN / 32) * DW_CHUNK_BYTES) /* 16 KB */
#define DW_STAGE_BYTES (2 * DW_A_BYTES + 2 * DW_B_BYTES)
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
  const int kb_per_step = v.n / TC_DW_BK;
  const int n_steps_max = min((int)*g.kmax, v.depth);
  const int n_kb_total = n_steps_max * kb_per_step;
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
        int b0 = (kb - step * kb_per_step) * TC_DW_BK;
        int slot = pos - step;
        if (slot < 0)
          slot += v.depth;
        int xrow = slot * v.cap + v.base + b0;
        int erow = step * v.cap + v.base + b0;
        uint8_t *st = smem + s * DW_STAGE_BYTES;
        mbar_expect_tx(&full[s], DW_STAGE_BYTES);
        tma_load_3d(&mXhi, &full[s], st, 0, xrow, m0 / 32);
        tma_load_3d(&mXlo, &full[s], st + DW_A_BYTES, 0, xrow, m0 / 32);
        tma_load_3d(&mEhi, &full[s], st + 2 * DW_A_BYTES, 0, erow, n0 / 32);
        tma_load_3d(&mElo, &full[s], st + 2 * DW_A_BYTES + DW_B_BYTES, 0, erow, n0 / 32);
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
        uint32_t a_lo = a_hi + DW_A_BYTES;
        uint32_t b_hi = a_hi + 2 * DW_A_BYTES;
        uint32_t b_lo = b_hi + DW_B_BYTES;
#pragma unroll
        for (int kk = 0; kk < TC_DW_BK / 8; kk++) {
          /* MN-major 32-bit operands (SWIZZLE_128B_BASE32B): each K row is one
             128-byte line of 32 floats along M/N; 32-float chunks along M/N
             are one TMA box apart (leading offset), groups of 4 K rows 512 B
             apart (stride offset); one MMA consumes 8 K rows = 1024 B */
          uint64_t dah = umma_desc(a_hi + kk * 1024, DW_CHUNK_BYTES, 512, UMMA_SW128_BASE32);
          uint64_t dal = umma_desc(a_lo + kk * 1024, DW_CHUNK_BYTES, 512, UMMA_SW128_BASE32);
          uint64_t dbh = umma_desc(b_hi + kk * 1024, DW_CHUNK_BYTES, 512, UMMA_SW128_BASE32);
          uint64_t dbl = umma_desc(b_lo + kk * 1024, DW_CHUNK_BYTES, 512, UMMA_SW128_BASE32);
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

/* ------------------------------------------------------------------------ */
/* DW on CTA pairs.
 *
 * k_tc_dw above is bound by the 64 B/clock a single SM can take in from L2:
 * a 128 x 256 tile needs 48 KB of operands per 16 rows of K.  Two SMs of a
 * TPC issuing ONE MMA (tcgen05 cta_group::2, M = 256) each keep half of the
 * N operand, so per SM and 16 rows only 16 KB (its 128 M columns) + 12 KB
 * (half of 192 N columns) arrive, and the tensor pipe, not the port, sets the
 * pace.  The output is transposed relative to k_tc_dw: M runs over the error
 * columns h (1024 = 4 pairs of 128, no padding), N over the input columns i
 * (6 tiles of 192), D[h][i] = sum_rows E[row][h] X[row][i]; both operands
 * stay MN-major.  The epilogue writes partial[z][i][h]: lanes are h, so every
 * store is a full line.
 *
 * Protocol (as CUTLASS's 2-SM pipelines): both CTAs' TMA loads count their
 * bytes on the leader's `full` barrier, on which only the leader's producer
 * arrives (expecting both halves); the leader's MMA thread releases a stage in
 * both CTAs with a multicast commit; the accumulator-ready commit is multicast
 * too, and each CTA's epilogue drains its own TMEM half.                     */

#define TC_DW2_STAGES 7
#define DW2_A_BYTES ((TC_BM / 32) * DW_CHUNK_BYTES)            /* 8 KB per plane */
#define DW2_B_BYTES ((TC_DW2_BN / 2 / 32) * DW_CHUNK_BYTES)    /* 6 KB per plane, this CTA's half */
#define DW2_STAGE_BYTES (2 * DW2_A_BYTES + 2 * DW2_B_BYTES)
#define DW2_SMEM_BYTES (TC_DW2_STAGES * DW2_STAGE_BYTES + 1024 + 256)
#define DW2_TMEM_COLS 256

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
        tma_load_3d_pair(&mEhi, &full[s], st, 0, erow, h0 / 32);
        tma_load_3d_pair(&mElo, &full[s], st + DW2_A_BYTES, 0, erow, h0 / 32);
        tma_load_3d_pair(&mXhi, &full[s], st + 2 * DW2_A_BYTES, 0, xrow, i_half / 32);
        tma_load_3d_pair(&mXlo, &full[s], st + 2 * DW2_A_BYTES + DW2_B_BYTES, 0, xrow,
            i_half / 32);
      }
    }
  }
  else if (warp == 1) {
    if (lane == 0 && leader) {
      const uint32_t idesc = umma_idesc_tf32(2 * TC_BM, TC_DW2_BN, 1, 1);
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
        for (int kk = 0; kk < TC_DW_BK / 8; kk++) {
          uint64_t dah = umma_desc(a_hi + kk * 1024, DW_CHUNK_BYTES, 512, UMMA_SW128_BASE32);
          uint64_t dal = umma_desc(a_lo + kk * 1024, DW_CHUNK_BYTES, 512, UMMA_SW128_BASE32);
          uint64_t dbh = umma_desc(b_hi + kk * 1024, DW_CHUNK_BYTES, 512, UMMA_SW128_BASE32);
          uint64_t dbl = umma_desc(b_lo + kk * 1024, DW_CHUNK_BYTES, 512, UMMA_SW128_BASE32);
          umma_tf32_pair(tmem_base, dal, dbh, idesc, (it | kk) ? 1u : 0u);
          umma_tf32_pair(tmem_base, dah, dbl, idesc, 1u);
          umma_tf32_pair(tmem_base, dah, dbh, idesc, 1u);
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
    float *dst = g.partial + (size_t)blockIdx.z * I * H + h;
    float acc[32];
#pragma unroll 1
    for (int c = 0; c < TC_DW2_BN; c += 32) {
      if (any)
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, acc);
      else {
#pragma unroll
        for (int j = 0; j < 32; j++)
          acc[j] = 0.0f;
      }
      if (h < H) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          int i = i0 + c + j;
          if (i < I)
            dst[(size_t)i * H] = acc[j];
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
  float *Whi, *Wlo, *WThi, *WTlo;
  /* the output matrix rides along */
  float *ho_W, *ho_mom, *ho_aux;
  const float *ho_delta;
  int ho_size;
  float ho_rate;
  int n_tiles_x, n_tiles;
};

__global__ void __launch_bounds__(256)
k_update_split(UpdateArgs a)
{
  if ((int)blockIdx.x >= a.n_tiles) {
    int i = ((int)blockIdx.x - a.n_tiles) * 256 + threadIdx.x;
    if (i < a.ho_size)
      a.ho_W[i] = rb_optimiser_step(a.method, a.ho_W[i], a.ho_delta[i], a.ho_mom, a.ho_aux, i,
          a.ho_rate, a.momentum, a.momentum_weight);
    return;
  }
  __shared__ float th[32][33], tl[32][33];
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
#pragma unroll
        for (int z = 0; z < TC_DW_SPLITS; z++)
          if (z < a.splits)
            t += __ldcg(a.partial + (size_t)z * size + i);
        d[q] = t;
      }
      else
        d[q] = a.delta[i];
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    int r = ty + 8 * q, y = y0 + r;
    float hi = 0.f, lo = 0.f;
    if (y < I && x < H) {
      size_t i = (size_t)y * H + x;
      if (a.partial)
        a.delta[i] = d[q];
      float nw = rb_optimiser_step(a.method, w[q], d[q], a.mom, a.aux, i, a.rate, a.momentum,
          a.momentum_weight);
      a.W[i] = nw;
      split_tf32(nw, hi, lo);
      a.Whi[i] = hi;
      a.Wlo[i] = lo;
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
      a.WThi[(size_t)xx * I + y] = th[tx][r];
      a.WTlo[(size_t)xx * I + y] = tl[tx][r];
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
     planes of the ring, shared with the forward pass, hold them unmasked */
  return v->contiguous && v->n >= 64 && (v->n % TC_BK) == 0 && v->d.h_size >= 64 &&
      v->activation != RNN_RECLIP20;
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
  k_split_weights<<<grid, 256, 0, rb_stream>>>(v->Wih, I, H, t->Whi, t->Wlo, t->WThi, t->WTlo);
  LAUNCH_CHECK("k_split_weights");
  rb_prof_end(RB_PROF_SMALL);
  t->w_src = v->Wih;
  t->w_version = g->weights_version;
}

static int fwd_attr_done = 0, chain_attr_done = 0, dw_attr_done = 0;
static int defer_delta_reduce = 0;
typedef NtCfg<TC_FWD_BN, TC_FWD_STAGES> FwdCfg;
typedef NtCfg<TC_CHAIN_BN, TC_CHAIN_STAGES> ChainCfg;

extern "C" void
rb_tc_x_planes(RbPool *p, float **Xhi, float **Xlo)
{
  RbTc *t = tc_state(p);
  *Xhi = t->Xhi;
  *Xlo = t->Xlo;
}

extern "C" void
rb_tc_forward(RbPool *p, const RbView *v, float presynaptic_noise)
{
  RbTc *t = tc_state(p);
  rbk_prepare_x(v);
  rb_prof_begin(RB_PROF_SMALL);
  k_split_rows<<<v->n, 256, 0, rb_stream>>>(*v, 0, t->Xhi, t->Xlo);
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
  if (presynaptic_noise != 0.0f) {
    rbk_gen_noise(v, presynaptic_noise, 1, v->d.h_size - 1);
    g.use_noise = 1;
  }
  /* Few tiles and a long K: split K so the GEMM covers the SMs, and let the
     output-layer kernel sum the partials on its way in. */
  {
    int n_kb = cdiv(v->d.i_size, TC_BK);
    int tiles = cdiv(v->d.h_size, TC_CHAIN_BN) * cdiv(v->n, TC_BM);
    int splits = TC_CHAIN_SPLITS;
    while (splits > 1 && (n_kb / splits < 2 || tiles * splits > 148))
      splits /= 2;
    if (splits > 1 && rbk_output_takes_partials(v, splits)) {
      if (!chain_attr_done) {
        CUDA_OR_DIE(cudaFuncSetAttribute(k_tc_nt<TC_CHAIN_BN, TC_CHAIN_STAGES>,
                cudaFuncAttributeMaxDynamicSharedMemorySize, ChainCfg::SMEM_BYTES));
        chain_attr_done = 1;
      }
      g.mode = 2;
      g.cpartial = t->cpartial;
      dim3 grid(cdiv(v->d.h_size, TC_CHAIN_BN), cdiv(v->n, TC_BM), splits);
      rb_prof_begin(RB_PROF_FWD);
      k_tc_nt<TC_CHAIN_BN, TC_CHAIN_STAGES><<<grid, 192, ChainCfg::SMEM_BYTES, rb_stream>>>(
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

extern "C" void
rb_tc_top_and_bptt(RbPool *p, const RbView *v, float *ho_delta, float *ih_delta, int accumulate)
{
  RbTc *t = tc_state(p);
  refresh_weight_planes(t, p, v);
  if (p->x_planes_stale) {
    /* ring rows were rewritten behind the planes' back (rnn_forget_history,
       rnn_b200_push, a per-net or FMA forward, a regrown pool): the weight
       gradient reads the ring only through the planes */
    size_t n = (size_t)p->depth * p->cap * v->d.i_size;
    rb_prof_begin(RB_PROF_SMALL);
    k_split_ring<<<148 * 4, 256, 0, rb_stream>>>(p->X, n, t->Xhi, t->Xlo);
    LAUNCH_CHECK("k_split_ring");
    rb_prof_end(RB_PROF_SMALL);
    p->x_planes_stale = 0;
  }
  /* top layer: E[0] and, where the batch kernel applies, its planes too */
  if (rbk_top_layer_can_write_planes(v)) {
    rbk_top_layer_planes(v, ho_delta, accumulate, NULL, 0, t->Ehi, t->Elo);
  }
  else {
    rbk_top_layer(v, ho_delta, accumulate, NULL, 0);
    rb_prof_begin(RB_PROF_SMALL);
    k_split_rows<<<v->n, 256, 0, rb_stream>>>(*v, 1, t->Ehi, t->Elo);
    LAUNCH_CHECK("k_split_rows");
    rb_prof_end(RB_PROF_SMALL);
  }
  if (!chain_attr_done) {
    CUDA_OR_DIE(cudaFuncSetAttribute(k_tc_nt<TC_CHAIN_BN, TC_CHAIN_STAGES>,
            cudaFuncAttributeMaxDynamicSharedMemorySize, ChainCfg::SMEM_BYTES));
    chain_attr_done = 1;
  }
  NtArgs g;
  g.v = *v;
  g.mode = 1;
  g.use_noise = 0;
  g.cpartial = t->cpartial;
  /* split K only as far as there are K blocks to share out */
  int n_kb = cdiv(v->d.h_size, TC_BK);
  int splits = TC_CHAIN_SPLITS;
  while (splits > 1 && n_kb / splits < 2)
    splits /= 2;
  dim3 cgrid(cdiv(v->d.i_size, TC_CHAIN_BN), cdiv(v->n, TC_BM), splits);
  int n_ctas = cgrid.x * cgrid.y * cgrid.z;
  /* device facts once; the choice of kernel per call (the batch may grow) */
  static int sms = 0, coop = 0, per_sm_single = 0;
  if (!sms) {
    int dev = 0;
    CUDA_OR_DIE(cudaGetDevice(&dev));
    CUDA_OR_DIE(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    CUDA_OR_DIE(cudaFuncSetAttribute(k_tc_chain_persistent<TC_CHAIN_BN, TC_CHAIN_STAGES>,
            cudaFuncAttributeMaxDynamicSharedMemorySize, ChainCfg::SMEM_BYTES));
    CUDA_OR_DIE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_single,
            k_tc_chain_persistent<TC_CHAIN_BN, TC_CHAIN_STAGES>, TC_CHAIN_THREADS, ChainCfg::SMEM_BYTES));
    CUDA_OR_DIE(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  t->persistent_ok = (coop && per_sm_single * sms >= n_ctas && !getenv("RECUR_B200_NO_PERSISTENT"));
  /* small nets: every stream walks alone with the weights resident in its SM */
  const bool resident = rbk_walk_resident_usable(v);
  if (resident)
    t->persistent_ok = 0;
  unsigned int *sync_area = t->sync + (size_t)t->sync_flip * t->sync_words;
  unsigned int *sync_next = t->sync + (size_t)(t->sync_flip ^ 1) * t->sync_words;
  t->sync_flip ^= 1;
  rb_note_walk_kernel(t->persistent_ok ? "k_tc_chain_persistent"
      : resident ? "k_walk_resident" : "k_tc_nt<CHAIN>");
  if (t->persistent_ok) {
    ChainArgs ca;
    ca.v = *v;
    ca.cpartial = t->cpartial;
    ca.Ehi = t->Ehi;
    ca.Elo = t->Elo;
    ca.sync = sync_area;
    ca.dbg = NULL;
    static unsigned long long *dbg_dev = NULL;
    static int dbg_calls = 0;
    const bool timing = getenv("RECUR_B200_CHAIN_TIMING") != NULL;
    if (timing) {
      if (!dbg_dev)
        CUDA_OR_DIE(cudaMalloc((void **)&dbg_dev, (1280) * sizeof(unsigned long long)));
      CUDA_OR_DIE(cudaMemsetAsync(dbg_dev, 0, (1280) * sizeof(unsigned long long), rb_stream));
      ca.dbg = dbg_dev;
    }
    ca.kmax = sync_area + (size_t)cgrid.y * (v->depth + 8) + 4;
    void *params[] = {(void *)&t->mEhi_k, (void *)&t->mElo_k, (void *)&t->mWhi_k,
                      (void *)&t->mWlo_k, (void *)&ca};
    rb_prof_begin(RB_PROF_CHAIN);
    if (!getenv("RECUR_B200_PLAIN_LAUNCH")) {
      /* cooperative by default: the driver guarantees all CTAs are resident
         together or fails the launch, whatever else shares the device */
      CUDA_OR_DIE(cudaLaunchCooperativeKernel(
              (void *)k_tc_chain_persistent<TC_CHAIN_BN, TC_CHAIN_STAGES>, cgrid, dim3(TC_CHAIN_THREADS),
              params, ChainCfg::SMEM_BYTES, rb_stream));
    }
    else {
      /* opt-out for measurements: co-residency then rests on the occupancy
         check above and on nothing else using the device */
      k_tc_chain_persistent<TC_CHAIN_BN, TC_CHAIN_STAGES><<<cgrid, TC_CHAIN_THREADS, ChainCfg::SMEM_BYTES,
        rb_stream>>>(t->mEhi_k, t->mElo_k, t->mWhi_k, t->mWlo_k, ca);
    }
    LAUNCH_CHECK("k_tc_chain_persistent");
    rb_prof_end(RB_PROF_CHAIN);
    if (timing && (++dbg_calls % 100) == 60) {
      static unsigned long long h[1280];
      CUDA_OR_DIE(cudaMemcpyAsync(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost, rb_stream));
      CUDA_OR_DIE(cudaStreamSynchronize(rb_stream));
      double a = 0, b1 = 0, b = 0, b2 = 0;
      int steps = 0;
      for (int k = 0; k < v->depth && k < 64 && h[k * 5 + 4]; k++, steps++) {
        a += (double)(h[k * 5 + 1] - h[k * 5 + 0]);
        b1 += (double)(h[k * 5 + 2] - h[k * 5 + 1]);
        b += (double)(h[k * 5 + 3] - h[k * 5 + 2]);
        b2 += (double)(h[k * 5 + 4] - h[k * 5 + 3]);
      }
      if (steps)
        fprintf(stderr, "chain timing (CTA 0, %d steps): phase A %.2f us, barrier %.2f us, "
            "phase B %.2f us, barrier %.2f us per step\n", steps, a / steps * 1e-3,
            b1 / steps * 1e-3, b / steps * 1e-3, b2 / steps * 1e-3);
      {
        const unsigned long long *q = h + 1200, t0 = h[320];
        fprintf(stderr, "step 5 CTA 0 phase A: alive check done +%.2f, first stage landed +%.2f, "
            "last stage landed +%.2f, accumulator ready +%.2f, epilogue done +%.2f, phase end +%.2f us\n",
            (double)(q[0] - t0) * 1e-3, (double)(q[1] - t0) * 1e-3, (double)(q[2] - t0) * 1e-3,
            (double)(q[3] - t0) * 1e-3, (double)(q[4] - t0) * 1e-3, (double)(h[321] - t0) * 1e-3);
        unsigned long long b0 = h[322];
        fprintf(stderr, "step 5 CTA 0 warp 2 phase B (after the barrier): start +%.2f, loads landed +%.2f, "
            "summed +%.2f, stored +%.2f, reduced +%.2f, tail done +%.2f, proxy fence done +%.2f, phase end +%.2f us\n",
            (double)(q[10] - b0) * 1e-3, (double)(q[11] - b0) * 1e-3, (double)(q[16] - b0) * 1e-3, (double)(q[12] - b0) * 1e-3,
            (double)(q[13] - b0) * 1e-3, (double)(q[14] - b0) * 1e-3, (double)(q[15] - b0) * 1e-3,
            (double)(h[323] - b0) * 1e-3);
      }
    }
  }
  else if (resident) {
    rbk_walk_resident(v, t->Ehi, t->Elo);
  }
  else
  for (int k = 0; k < v->depth; k++) {
    g.k = k;
    rb_prof_begin(RB_PROF_CHAIN);
    k_tc_nt<TC_CHAIN_BN, TC_CHAIN_STAGES><<<cgrid, 192, ChainCfg::SMEM_BYTES, rb_stream>>>(
        t->mEhi_k, t->mElo_k, t->mWhi_k, t->mWlo_k, g);
    LAUNCH_CHECK("k_tc_nt<CHAIN>");
    k_chain_finish_step<<<v->n, 256, 0, rb_stream>>>(*v, k, t->cpartial, splits, t->Ehi, t->Elo);
    LAUNCH_CHECK("k_chain_finish_step");
    rb_prof_end(RB_PROF_CHAIN);
  }
  unsigned int *kmax_dev = sync_area + (size_t)cdiv(v->n, TC_BM) * (v->depth + 8) + 4;
  if (!t->persistent_ok) {
    k_compute_kmax<<<1, 256, 0, rb_stream>>>(*v, kmax_dev);
    LAUNCH_CHECK("k_compute_kmax");
  }
  rb_prof_begin(RB_PROF_SMALL);
  k_finalize_rows<<<v->n, 256, 0, rb_stream>>>(*v, t->Ehi, t->Elo, kmax_dev, sync_next,
      (int)t->sync_words);
  LAUNCH_CHECK("k_finalize_rows");
  rb_prof_end(RB_PROF_SMALL);
  if (!dw_attr_done) {
    CUDA_OR_DIE(cudaFuncSetAttribute(k_tc_dw, cudaFuncAttributeMaxDynamicSharedMemorySize,
            DW_SMEM_BYTES));
    dw_attr_done = 1;
  }
  DwArgs d;
  d.v = *v;
  d.partial = t->partial;
  d.kmax = kmax_dev;
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
  if (dw_pair_ok && psplits >= 1) {
    /* CTA pairs: one MMA across two SMs, each holding half of the N operand */
    dim3 dgrid(pgx, pgy, psplits);
    k_tc_dw_pair<<<dgrid, 192, DW2_SMEM_BYTES, rb_stream>>>(t->mEhi_mn4, t->mElo_mn4,
        t->mXhi_mn3, t->mXlo_mn3, d);
    LAUNCH_CHECK("k_tc_dw_pair");
    t->dw_splits = psplits;
  }
  else {
    dim3 dgrid(cdiv(v->d.h_size, TC_DW_BN), cdiv(v->d.i_size, TC_BM), TC_DW_SPLITS);
    k_tc_dw<<<dgrid, 192, DW_SMEM_BYTES, rb_stream>>>(t->mXhi_mn, t->mXlo_mn, t->mEhi_mn,
        t->mElo_mn, d);
    LAUNCH_CHECK("k_tc_dw");
    t->dw_splits = TC_DW_SPLITS;
  }
  int size = v->d.i_size * v->d.h_size;
  if (rb_p2p_ready(v->p2p)) {
    /* multi-GPU: the split-K sum is the first phase of the exchange kernel */
    rb_p2p_reduce(v->p2p, t->partial, t->dw_splits, size, v->d.h_size * v->d.o_size, ih_delta,
        accumulate);
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
  a.ho_W = net->ho_weights;
  a.ho_mom = b->ho_momentum;
  a.ho_aux = b->ho_aux;
  a.ho_delta = b->ho_delta;
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
