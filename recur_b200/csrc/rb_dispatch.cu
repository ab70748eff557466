/* rb_dispatch.cu — picks the matrix engine for the three big contractions.
 *
 * FMA engine (rb_kernels.cu): exact FP32 on CUDA cores, any batch size.
 * Tensor engine (rb_tc.cu): tcgen05 3xTF32 for batches of >= 64 streams.
 */
#include "rb_kernels.h"
#include "rb_host.h"

extern "C" int rb_engine(void);

extern "C" void
rb_weights_changed(RecurNN *net)
{
  rb_net_of(net)->group->weights_version++;
}

extern "C" void
rb_forward_dispatch(const RbView *v, float noise)
{
  rbk_forward(v, noise);
}

extern "C" void
rb_bptt_dispatch(const RbView *v, float *ih_delta, int accumulate)
{
  rbk_bptt(v, ih_delta, accumulate);
}
