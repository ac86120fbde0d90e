// hmm_forward_w32.cu — forward-kernel instances for groups of 32 lanes per job (C = 1..10 columns per lane).
#include "hmm_forward_kernel.cuh"
namespace nph_fwd {
NPH_DEFINE_LAUNCH_WIDTH(32)
}
