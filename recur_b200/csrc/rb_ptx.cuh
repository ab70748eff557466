/* rb_ptx.cuh - thin wrappers over the sm_100a PTX the tensor-core kernels are
 * written in: mbarriers, TMA loads, TMEM allocation, tcgen05 MMAs / loads and
 * their shared-memory and instruction descriptors, cluster / DSMEM access.
 * Shared by rb_tc.cu (training walks) and rb_cells.cu (the cell automaton).
 */
#ifndef RB_PTX_CUH
#define RB_PTX_CUH

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

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
mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

/* D[tmem] (+)= A[smem] . B[smem], kind::f16, issued by one thread */
__device__ __forceinline__ void
umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
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

/* the address of a shared-memory object of this CTA as it appears in the
   shared window of the cluster's CTA `rank` (distributed shared memory) */
__device__ __forceinline__ uint32_t
dsmem_addr(const void *p, uint32_t rank)
{
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(p)), "r"(rank));
  return a;
}

__device__ __forceinline__ float4
ld_dsmem_v4(uint32_t addr)
{
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
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
umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
    uint32_t accumulate)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
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

/* 16 consecutive accumulator columns of this thread's TMEM lane, no wait */
__device__ __forceinline__ void
tmem_ld16_nowait(uint32_t taddr, float *v)
{
  uint32_t *r = (uint32_t *)v;
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
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
#define UMMA_SW128 2u /* 128-byte rows, 16-byte chunks swizzled over 8 rows */
#define UMMA_SW64 4u  /* 64-byte rows, swizzled over 4 rows */

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
   kind::f16 with FP16 operands (a_format = b_format = 0), FP32 accumulate. */
__host__ __device__ constexpr uint32_t
umma_idesc_f16(int M, int N, int a_mn_major, int b_mn_major)
{
  return (1u << 4)      /* c_format  F32 */
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

#endif
