// hmm_forward_w8.cu — forward-kernel instances for groups of 8 lanes per job (C = 1..10 columns per lane).
#include "hmm_forward_kernel.cuh"
namespace nph_fwd {
NPH_DEFINE_LAUNCH_WIDTH(8, false)
}
