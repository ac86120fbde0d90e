// hmm_forward_w16.cu — forward-kernel instances for groups of 16 lanes per job (C = 1..10 columns per lane).
#include "hmm_forward_kernel.cuh"
namespace nph_fwd {
NPH_DEFINE_LAUNCH_WIDTH(16, false)
}
