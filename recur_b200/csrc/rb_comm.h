/* rb_comm.h — gradient exchange between the processes of a multi-GPU run.
 *
 * Streams shard across GPUs (one process per GPU); the only exchange the
 * path has is the sum of the shared delta arrays before the weight update
 * (SURVEY.md §8e).  NCCL is loaded at run time (dlopen) so the library has
 * no link-time dependency on it and single-GPU users never touch it. */
#ifndef RB_COMM_H
#define RB_COMM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
int rb_comm_size(void);
int rb_comm_rank(void);
/* in-place sum over ranks of n floats at a device-accessible address, queued
   on the library stream */
void rb_comm_allreduce_sum(float *buf, size_t n);
/* in-place all-gather of `count` buffers, each rank owning bytes_each bytes at
   offset rank * bytes_each of every buffer */
void rb_comm_allgather_inplace(unsigned char **bufs, int count, size_t bytes_each);
#ifdef __cplusplus
}
#endif
#endif
