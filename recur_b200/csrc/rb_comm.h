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
#ifdef __cplusplus
}
#endif
#endif
