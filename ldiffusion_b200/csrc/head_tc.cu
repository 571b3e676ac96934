// tcgen05 / TMEM contraction for the classifier head (bf16 in, fp32 accumulate).
#include "common.cuh"

namespace ldiff {

int launch_head_logits_tc(const void*, const void*, const float*, float*, int, int, int, int,
                          cudaStream_t) {
  return LDIFF_EUNSUPPORTED;   // placeholder until the tensor-core kernel lands
}

}  // namespace ldiff
