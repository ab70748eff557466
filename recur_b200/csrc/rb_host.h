/* rb_host.h — functions shared between the API translation units. */
#ifndef RB_HOST_H
#define RB_HOST_H

#include "rb_internal.h"

#ifdef __cplusplus
extern "C" {
#endif

void rb_net_pull(RbNet *rn);
/* bumped whenever a net's dev_ahead flag is cleared (pull / push) */
extern unsigned long long rb_ahead_epoch;
void rb_net_push(RbNet *rn);
void rb_apply_learning_async(RecurNN *net, int method, float momentum);
void rb_weights_changed(RecurNN *net);

/* host restatements of badmaths.h / recur-nn-helpers.h scalars (rb_init.c) */
float rb_fast_expf(float x);

#ifdef __cplusplus
}
#endif
#endif
