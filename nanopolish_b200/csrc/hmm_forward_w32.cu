// hmm_forward_w32.cu — forward-kernel instances for full-warp jobs that fit one strip (K <= 32*C).
#include "hmm_forward_kernel.cuh"
namespace nph_fwd {
NPH_DEFINE_LAUNCH_WIDTH(32, false)
}
