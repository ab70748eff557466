/* rb_p2p.cu — the weight-gradient reduction fused with its exchange between
 * GPUs (one process per GPU, buffers shared through CUDA IPC, NVLink).
 *
 * Single GPU: the update kernel sums the split-K partial planes of the weight
 * gradient on its way through the weights.  Multi GPU: ONE kernel does that
 * sum and the reduce-scatter / all-gather of the [ih_delta | ho_delta] block
 * with PEER STORES only (nobody waits for a load from another GPU):
 *
 *   phase 1  every element's local sum (split-K planes; ho_delta as it is)
 *            goes straight into the inbox of the rank that owns the
 *            element's slice - slice q of the block belongs to rank q
 *   flag     the last CTA to finish publishes this rank's epoch to every peer
 *   phase 2  after all peers' epochs: this rank sums the N inbox copies of
 *            its slice in rank order (one rank computes each element, so all
 *            replicas receive bit-identical sums) and stores the result into
 *            every rank's result block
 *   flag     second epoch.  Nobody waits for it here: the consumer does - the
 *            update kernel, which reads the result block in place of the
 *            split-K planes (rb_tc_fused_update), or k_p2p_copy_out for a
 *            caller that wants the deltas in the arrays the API exposes.
 *
 * Per GPU and step that is 2 (N-1)/N of the block out over NVLink, two flag
 * rounds, and no pass over the block that does not also do arithmetic.
 * Buffers are plain cudaMalloc memory shared through CUDA IPC handles that the
 * launcher gathers (include/recur_b200.h); flags are written with release
 * and read with acquire at system scope.  The grid is sized to be co-resident
 * (CTAs spin on flags), and the spins are bounded: a peer that never arrives
 * traps the kernel instead of hanging the device.
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
  size_t slice;             /* floats per rank's slice (a multiple of 4) */
  float *inbox[RB_P2P_MAX]; /* rank q's inbox: [n][slice], row r written by rank r */
  float *result[RB_P2P_MAX];
  unsigned int *flags[RB_P2P_MAX]; /* [2 * RB_P2P_MAX] epochs + [2] local CTA counters */
  unsigned int epoch;
  int attached;
} RbP2P;

struct P2PArgs {
  int n, rank;
  int slice;
  float *inbox[RB_P2P_MAX];
  float *result[RB_P2P_MAX];
  unsigned int *flags[RB_P2P_MAX];
  unsigned int epoch;
  const float *partial;
  int splits;
  int ih_size, ho_size;
  const float *ho_delta; /* this rank's ho_delta */
};

__device__ __forceinline__ void
st_release_sys(unsigned int *p, unsigned int v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void
st_relaxed_sys(unsigned int *p, unsigned int v)
{
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned int
ld_acquire_sys(const unsigned int *p)
{
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

/* every CTA has finished its part: the last one tells all ranks.  One thread
   per CTA fences: what the others wrote it has observed through the block
   barrier, and release is cumulative. */
__device__ __forceinline__ void
publish_epoch(const P2PArgs &a, int which)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned int *counter = a.flags[a.rank] + 2 * RB_P2P_MAX + which;
    unsigned int done = atomicAdd(counter, 1u) + 1;
    if (done == gridDim.x) {
      *counter = 0;
      /* ONE fence orders every CTA's stores (observed through the counter)
         before the flags; the flags themselves go out relaxed, all n in
         flight at once - a release per flag would wait for the flag before it
         to cross NVLink, n round trips in a row */
      __threadfence_system();
      for (int q = 0; q < a.n; q++)
        st_relaxed_sys(a.flags[q] + which * RB_P2P_MAX + a.rank, a.epoch);
    }
  }
}

/* all ranks' epochs of round `which` have reached this rank's flags; bounded:
   2^24 polls of local memory are seconds */
__device__ __forceinline__ void
await_flags(const unsigned int *flags, int which, int n, unsigned int epoch)
{
  if ((int)threadIdx.x < n) {
    const unsigned int *f = flags + which * RB_P2P_MAX + threadIdx.x;
    unsigned int spins = 0;
    while (ld_acquire_sys(f) < epoch) {
      if (++spins > (1u << 24))
        __trap();
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512)
k_dw_reduce_exchange(P2PArgs a)
{
  const int total = a.ih_size + a.ho_size;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  /* phase 1: local sums, pushed to the owners of their slices */
  for (int i = tid * 4; i < total; i += nthreads * 4) {
    float4 s;
    if (i < a.ih_size) {
      s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int z = 0; z < 4; z++) {
        if (z < a.splits) {
          float4 p = __ldcg((const float4 *)(a.partial + (size_t)z * a.ih_size + i));
          s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
        }
      }
    }
    else {
      s = *(const float4 *)(a.ho_delta + (i - a.ih_size)); /* already summed over this rank's streams */
    }
    const int q = i / a.slice;
    *(float4 *)(a.inbox[q] + (size_t)a.rank * a.slice + (i - q * a.slice)) = s;
  }
  publish_epoch(a, 0);
  await_flags(a.flags[a.rank], 0, a.n, a.epoch);
  /* phase 2: this rank's slice, the copies in rank order, to everybody */
  const int lo = a.rank * a.slice, hi = min(total, lo + a.slice);
  const float *in = a.inbox[a.rank];
  for (int i = lo + tid * 4; i < hi; i += nthreads * 4) {
    float4 p[RB_P2P_MAX];
#pragma unroll
    for (int q = 0; q < RB_P2P_MAX; q++)
      if (q < a.n)
        p[q] = __ldcv((const float4 *)(in + (size_t)q * a.slice + (i - lo)));
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
        *(float4 *)(a.result[q] + i) = s;
  }
  publish_epoch(a, 1);
}

/* the exchanged block -> the delta arrays the API exposes */
__global__ void __launch_bounds__(512)
k_p2p_copy_out(const float *__restrict__ result, const unsigned int *flags, int n,
    unsigned int epoch, float *ih_delta, int total)
{
  await_flags(flags, 1, n, epoch);
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total;
       i += gridDim.x * blockDim.x * 4)
    *(float4 *)(ih_delta + i) = __ldcv((const float4 *)(result + i));
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
  /* the inbox holds n copies of one slice, n * ceil(total / n) floats: room
     for the block plus a 4-float round-up per rank */
  size_t bytes = (p->n_floats + 64) * sizeof(float);
  CUDA_TRY(cudaMalloc((void **)&stage, bytes));
  CUDA_TRY(cudaMalloc((void **)&result, bytes));
  CUDA_TRY(cudaMalloc((void **)&flags, 64 * sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(flags, 0, 64 * sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(stage, 0, bytes));
  CUDA_TRY(cudaMemset(result, 0, bytes));
  p->inbox[0] = stage;   /* parked in slot 0 until attach knows our rank */
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
  float *stage = p->inbox[0], *result = p->result[0];
  unsigned int *flags = p->flags[0];
  const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)all_handles;
  for (int q = 0; q < n_ranks; q++) {
    if (q == rank) {
      p->inbox[q] = stage;
      p->result[q] = result;
      p->flags[q] = flags;
      continue;
    }
    CUDA_TRY(cudaIpcOpenMemHandle((void **)&p->inbox[q], h[q * 3 + 0], cudaIpcMemLazyEnablePeerAccess));
    CUDA_TRY(cudaIpcOpenMemHandle((void **)&p->result[q], h[q * 3 + 1], cudaIpcMemLazyEnablePeerAccess));
    CUDA_TRY(cudaIpcOpenMemHandle((void **)&p->flags[q], h[q * 3 + 2], cudaIpcMemLazyEnablePeerAccess));
  }
  p->n = n_ranks;
  p->rank = rank;
  p->epoch = 0;
  p->slice = ((p->n_floats / 4 + n_ranks - 1) / n_ranks) * 4;
  if (p->slice * n_ranks > p->n_floats + 64)
    return -1;
  p->attached = 1;
  return 0;
}

extern "C" int
rb_p2p_ready(void *state)
{
  return state && ((RbP2P *)state)->attached;
}

/* sum of the split-K partials + exchange of [ih_delta | ho_delta], one kernel;
   the result waits in this rank's result block behind the second flag round */
extern "C" void
rb_p2p_exchange(void *state, const float *partial, int splits, int ih_size, int ho_size,
    const float *ho_delta)
{
  RbP2P *p = (RbP2P *)state;
  if ((size_t)ih_size + ho_size != p->n_floats || (ih_size & 3) || (ho_size & 3))
    rb_die("recur-b200: the peer exchange was set up for %zu floats, not %d + %d", p->n_floats,
        ih_size, ho_size);
  P2PArgs a;
  memset(&a, 0, sizeof(a));
  a.n = p->n;
  a.rank = p->rank;
  a.slice = (int)p->slice;
  for (int q = 0; q < p->n; q++) {
    a.inbox[q] = p->inbox[q];
    a.result[q] = p->result[q];
    a.flags[q] = p->flags[q];
  }
  a.epoch = ++p->epoch;
  a.partial = partial;
  a.splits = splits;
  a.ih_size = ih_size;
  a.ho_size = ho_size;
  a.ho_delta = ho_delta;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  k_dw_reduce_exchange<<<sms, 512, 0, rb_stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    rb_die("recur-b200: launch of k_dw_reduce_exchange failed: %s", cudaGetErrorString(e));
  rb_count_launch(1);
}

/* where the exchanged block is, and what its consumer has to wait for */
extern "C" void
rb_p2p_result(void *state, const float **result, const unsigned int **flags, unsigned int *epoch,
    int *n)
{
  RbP2P *p = (RbP2P *)state;
  *result = p->result[p->rank];
  *flags = p->flags[p->rank];
  *epoch = p->epoch;
  *n = p->n;
}

extern "C" void
rb_p2p_copy_out(void *state, float *ih_delta)
{
  RbP2P *p = (RbP2P *)state;
  k_p2p_copy_out<<<148, 512, 0, rb_stream>>>(p->result[p->rank], p->flags[p->rank], p->n,
      p->epoch, ih_delta, (int)p->n_floats);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    rb_die("recur-b200: launch of k_p2p_copy_out failed: %s", cudaGetErrorString(e));
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
    cudaIpcCloseMemHandle(p->inbox[q]);
    cudaIpcCloseMemHandle(p->result[q]);
    cudaIpcCloseMemHandle(p->flags[q]);
  }
  int me = p->attached ? p->rank : 0;
  cudaFree(p->inbox[me]);
  cudaFree(p->result[me]);
  cudaFree(p->flags[me]);
  free(p);
}
