// hmm_forward_w32c.cu — forward-kernel instances for full-warp jobs wider than one strip (chained strips).
#include "hmm_forward_kernel.cuh"
namespace nph_fwd {
NPH_DEFINE_LAUNCH_WIDTH(32, true)
}
