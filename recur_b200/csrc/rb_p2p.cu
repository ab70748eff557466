/* rb_p2p.cu — the weight-gradient reduction fused with its all-reduce.
 *
 * Single GPU: k_dw_reduce sums the split-K partial planes of k_tc_dw into
 * ih_delta.  Multi GPU (one process per GPU): the same sum is the first phase
 * of ONE kernel that also exchanges the result over NVLink peer memory:
 *
 *   phase 1  local: stage[i] = (old delta) + sum_z partial[z][i]; ho_delta is
 *            appended, so stage is the whole [ih_delta | ho_delta] block
 *   flag     the last CTA to finish publishes this rank's epoch to every peer
 *   phase 2  this rank owns slice r of the block: peer loads of slice r from
 *            every rank's stage, summed in rank order (one rank computes each
 *            element, so all replicas receive bit-identical sums), peer stores
 *            of the result into every rank's result buffer
 *   flag     second epoch flag
 *   phase 3  result -> the (managed) delta arrays the API exposes
 *
 * Buffers are plain cudaMalloc memory shared through CUDA IPC handles that the
 * launcher gathers (include/recur_b200.h).  Peer data is read with ld.cv and
 * flags with acquire/release at system scope.  The grid is sized to be
 * co-resident (CTAs spin on flags).
 */
#include "rb_kernels.h"
#include "rb_host.h"
#include "rb_comm.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define RB_P2P_MAX 8

typedef struct RbP2P {
  int n, rank;
  size_t n_floats;          /* ih_size + ho_size */
  float *stage[RB_P2P_MAX];
  float *result[RB_P2P_MAX];
  unsigned int *flags[RB_P2P_MAX]; /* [2 * RB_P2P_MAX] epochs + [2] local CTA counters */
  unsigned int epoch;
  int attached;
} RbP2P;

struct P2PArgs {
  int n, rank;
  float *stage[RB_P2P_MAX];
  float *result[RB_P2P_MAX];
  unsigned int *flags[RB_P2P_MAX];
  unsigned int epoch;
  const float *partial;
  int splits;
  int ih_size, ho_size;
  float *ih_delta;  /* ho_delta follows */
  int accumulate;
};

__device__ __forceinline__ void
st_release_sys(unsigned int *p, unsigned int v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned int
ld_acquire_sys(const unsigned int *p)
{
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

/* every CTA has finished its part: the last one tells all ranks */
__device__ __forceinline__ void
publish_epoch(const P2PArgs &a, int which)
{
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int *counter = a.flags[a.rank] + 2 * RB_P2P_MAX + which;
    unsigned int done = atomicAdd(counter, 1u) + 1;
    if (done == gridDim.x) {
      *counter = 0;
      __threadfence_system();
      for (int q = 0; q < a.n; q++)
        st_release_sys(a.flags[q] + which * RB_P2P_MAX + a.rank, a.epoch);
    }
  }
}

__device__ __forceinline__ void
await_epoch(const P2PArgs &a, int which)
{
  if (threadIdx.x < a.n) {
    const unsigned int *f = a.flags[a.rank] + which * RB_P2P_MAX + threadIdx.x;
    while (ld_acquire_sys(f) < a.epoch)
      ;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512)
k_dw_reduce_allreduce(P2PArgs a)
{
  const int total = a.ih_size + a.ho_size;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  float *stage = a.stage[a.rank];
  /* phase 1 */
  for (int i = tid * 4; i < total; i += nthreads * 4) {
    float4 s;
    if (i < a.ih_size) {
      s = a.accumulate ? *(const float4 *)(a.ih_delta + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int z = 0; z < 4; z++) {
        if (z < a.splits) {
          float4 p = *(const float4 *)(a.partial + (size_t)z * a.ih_size + i);
          s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
        }
      }
    }
    else {
      s = *(const float4 *)(a.ih_delta + i); /* ho_delta, already summed over this rank's streams */
    }
    *(float4 *)(stage + i) = s;
  }
  publish_epoch(a, 0);
  await_epoch(a, 0);
  /* phase 2: slice `rank` of the block, in 4-float units */
  const int n4 = total / 4;
  const int per = (n4 + a.n - 1) / a.n;
  const int lo = a.rank * per, hi = min(n4, lo + per);
  for (int i = lo + tid; i < hi; i += nthreads) {
    /* all peer loads in flight together (an NVLink round trip each), then
       the sum in rank order */
    float4 p[RB_P2P_MAX];
#pragma unroll
    for (int q = 0; q < RB_P2P_MAX; q++)
      if (q < a.n)
        p[q] = __ldcv((const float4 *)a.stage[q] + i);
    float4 s = p[0];
#pragma unroll
    for (int q = 1; q < RB_P2P_MAX; q++) {
      if (q < a.n) {
        s.x += p[q].x; s.y += p[q].y; s.z += p[q].z; s.w += p[q].w;
      }
    }
#pragma unroll
    for (int q = 0; q < RB_P2P_MAX; q++)
      if (q < a.n)
        *((float4 *)a.result[q] + i) = s;
  }
  publish_epoch(a, 1);
  await_epoch(a, 1);
  /* phase 3 */
  const float *res = a.result[a.rank];
  for (int i = tid * 4; i < total; i += nthreads * 4)
    *(float4 *)(a.ih_delta + i) = __ldcv((const float4 *)(res + i));
}

#define CUDA_TRY(call) do {                                             \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess) {                                            \
      fprintf(stderr, "recur-b200: %s: %s\n", #call, cudaGetErrorString(e_)); \
      (void)cudaGetLastError();                                         \
      return -1;                                                        \
    }                                                                   \
  } while (0)

extern "C" void *
rb_p2p_new(size_t n_floats)
{
  RbP2P *p = (RbP2P *)calloc(1, sizeof(RbP2P));
  p->n_floats = n_floats;
  return p;
}

extern "C" int
rb_p2p_export(void *state, void *handles_out)
{
  RbP2P *p = (RbP2P *)state;
  float *stage = NULL, *result = NULL;
  unsigned int *flags = NULL;
  size_t bytes = (p->n_floats + 64) * sizeof(float);
  CUDA_TRY(cudaMalloc((void **)&stage, bytes));
  CUDA_TRY(cudaMalloc((void **)&result, bytes));
  CUDA_TRY(cudaMalloc((void **)&flags, 64 * sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(flags, 0, 64 * sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(stage, 0, bytes));
  CUDA_TRY(cudaMemset(result, 0, bytes));
  p->stage[0] = stage;   /* parked in slot 0 until attach knows our rank */
  p->result[0] = result;
  p->flags[0] = flags;
  cudaIpcMemHandle_t h[3];
  CUDA_TRY(cudaIpcGetMemHandle(&h[0], stage));
  CUDA_TRY(cudaIpcGetMemHandle(&h[1], result));
  CUDA_TRY(cudaIpcGetMemHandle(&h[2], flags));
  memcpy(handles_out, h, sizeof(h));
  return 0;
}

extern "C" int
rb_p2p_attach(void *state, const void *all_handles, int rank, int n_ranks)
{
  RbP2P *p = (RbP2P *)state;
  if (n_ranks > RB_P2P_MAX || n_ranks < 2)
    return -1;
  float *stage = p->stage[0], *result = p->result[0];
  unsigned int *flags = p->flags[0];
  const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)all_handles;
  for (int q = 0; q < n_ranks; q++) {
    if (q == rank) {
      p->stage[q] = stage;
      p->result[q] = result;
      p->flags[q] = flags;
      continue;
    }
    CUDA_TRY(cudaIpcOpenMemHandle((void **)&p->stage[q], h[q * 3 + 0], cudaIpcMemLazyEnablePeerAccess));
    CUDA_TRY(cudaIpcOpenMemHandle((void **)&p->result[q], h[q * 3 + 1], cudaIpcMemLazyEnablePeerAccess));
    CUDA_TRY(cudaIpcOpenMemHandle((void **)&p->flags[q], h[q * 3 + 2], cudaIpcMemLazyEnablePeerAccess));
  }
  p->n = n_ranks;
  p->rank = rank;
  p->epoch = 0;
  p->attached = 1;
  return 0;
}

extern "C" int
rb_p2p_ready(void *state)
{
  return state && ((RbP2P *)state)->attached;
}

/* sum of split-K partials + all-reduce of [ih_delta | ho_delta], one kernel */
extern "C" void
rb_p2p_reduce(void *state, const float *partial, int splits, int ih_size, int ho_size,
    float *ih_delta, int accumulate)
{
  RbP2P *p = (RbP2P *)state;
  P2PArgs a;
  memset(&a, 0, sizeof(a));
  a.n = p->n;
  a.rank = p->rank;
  for (int q = 0; q < p->n; q++) {
    a.stage[q] = p->stage[q];
    a.result[q] = p->result[q];
    a.flags[q] = p->flags[q];
  }
  a.epoch = ++p->epoch;
  a.partial = partial;
  a.splits = splits;
  a.ih_size = ih_size;
  a.ho_size = ho_size;
  a.ih_delta = ih_delta;
  a.accumulate = accumulate;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  k_dw_reduce_allreduce<<<sms, 512, 0, rb_stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    rb_die("recur-b200: launch of k_dw_reduce_allreduce failed: %s", cudaGetErrorString(e));
  rb_count_launch(1);
}

extern "C" void
rb_p2p_delete(void *state)
{
  RbP2P *p = (RbP2P *)state;
  if (!p)
    return;
  for (int q = 0; q < p->n; q++) {
    if (q == p->rank)
      continue;
    cudaIpcCloseMemHandle(p->stage[q]);
    cudaIpcCloseMemHandle(p->result[q]);
    cudaIpcCloseMemHandle(p->flags[q]);
  }
  int me = p->attached ? p->rank : 0;
  cudaFree(p->stage[me]);
  cudaFree(p->result[me]);
  cudaFree(p->flags[me]);
  free(p);
}
