/* rb_comm.cu — NCCL (dlopen'ed) behind rnn_b200_comm_* (include/recur_b200.h). */
#include "rb_comm.h"
#include "rb_kernels.h"
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* the handful of NCCL declarations used, as in nccl.h (ABI-stable since 2.x) */
typedef struct { char internal[128]; } rb_ncclUniqueId;
typedef void *rb_ncclComm_t;
enum { RB_NCCL_UINT8 = 1, RB_NCCL_FLOAT = 7, RB_NCCL_SUM = 0 };

static struct {
  void *lib;
  int (*GetUniqueId)(rb_ncclUniqueId *);
  int (*CommInitRank)(rb_ncclComm_t *, int, rb_ncclUniqueId, int);
  int (*AllReduce)(const void *, void *, size_t, int, int, rb_ncclComm_t, cudaStream_t);
  int (*AllGather)(const void *, void *, size_t, int, rb_ncclComm_t, cudaStream_t);
  int (*GroupStart)(void);
  int (*GroupEnd)(void);
  int (*CommDestroy)(rb_ncclComm_t);
  const char *(*GetErrorString)(int);
  rb_ncclComm_t comm;
  int rank, size;
} g_nccl = {0};

static int
load_nccl(void)
{
  if (g_nccl.lib)
    return 0;
  const char *names[] = {getenv("RECUR_B200_NCCL"), "libnccl.so.2", "libnccl.so", NULL};
  for (int i = 0; i < 4 && !g_nccl.lib; i++) {
    if (names[i])
      g_nccl.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  }
  if (!g_nccl.lib) {
    fprintf(stderr, "recur-b200: cannot load NCCL: %s\n", dlerror());
    return -1;
  }
#define SYM(field, name) do {                                           \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name);                \
    if (!g_nccl.field) {                                                \
      fprintf(stderr, "recur-b200: NCCL lacks %s\n", name);             \
      dlclose(g_nccl.lib);                                              \
      g_nccl.lib = NULL;                                                \
      return -1;                                                        \
    }} while (0)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return 0;
}

extern "C" int
rnn_b200_comm_unique_id(void *id128)
{
  if (load_nccl())
    return -1;
  rb_ncclUniqueId id;
  int r = g_nccl.GetUniqueId(&id);
  if (r) {
    fprintf(stderr, "recur-b200: ncclGetUniqueId: %s\n", g_nccl.GetErrorString(r));
    return -1;
  }
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int
rnn_b200_comm_join(const void *id128, int rank, int n_ranks)
{
  if (n_ranks <= 1) {
    g_nccl.size = 1;
    g_nccl.rank = 0;
    return 0;
  }
  rb_require_device("rnn_b200_comm_join");
  if (load_nccl())
    return -1;
  rb_ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  int r = g_nccl.CommInitRank(&g_nccl.comm, n_ranks, id, rank);
  if (r) {
    fprintf(stderr, "recur-b200: ncclCommInitRank: %s\n", g_nccl.GetErrorString(r));
    return -1;
  }
  g_nccl.rank = rank;
  g_nccl.size = n_ranks;
  return 0;
}

extern "C" void
rnn_b200_comm_leave(void)
{
  if (g_nccl.comm) {
    cudaStreamSynchronize(rb_stream);
    g_nccl.CommDestroy(g_nccl.comm);
    g_nccl.comm = NULL;
  }
  g_nccl.size = 1;
  g_nccl.rank = 0;
}

extern "C" int
rnn_b200_comm_size(void)
{
  return g_nccl.size > 1 ? g_nccl.size : 1;
}

extern "C" int
rb_comm_size(void)
{
  return rnn_b200_comm_size();
}

extern "C" int
rb_comm_rank(void)
{
  return g_nccl.rank;
}

extern "C" void
rb_comm_allreduce_sum(float *buf, size_t n)
{
  if (g_nccl.size <= 1 || !g_nccl.comm)
    return;
  int r = g_nccl.AllReduce(buf, buf, n, RB_NCCL_FLOAT, RB_NCCL_SUM, g_nccl.comm, rb_stream);
  if (r)
    rb_die("recur-b200: ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
}

/* `count` equal pieces of `bytes_each` per rank, piece k of rank r living at
   bufs[k] + r * bytes_each on every rank: afterwards every rank holds all of
   every buffer.  One NCCL group, queued on the library stream. */
extern "C" void
rb_comm_allgather_inplace(unsigned char **bufs, int count, size_t bytes_each)
{
  if (g_nccl.size <= 1 || !g_nccl.comm)
    return;
  int r = g_nccl.GroupStart();
  for (int k = 0; k < count && !r; k++)
    r = g_nccl.AllGather(bufs[k] + (size_t)g_nccl.rank * bytes_each, bufs[k], bytes_each,
        RB_NCCL_UINT8, g_nccl.comm, rb_stream);
  if (!r)
    r = g_nccl.GroupEnd();
  if (r)
    rb_die("recur-b200: ncclAllGather failed: %s", g_nccl.GetErrorString(r));
}
