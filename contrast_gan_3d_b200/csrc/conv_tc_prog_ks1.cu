// Instantiations of the strided tap-program kernel for KSTEPS = 1 (Cin = 16 or the paired 8-channel slabs); see conv_tc_prog_kernel.cuh.
#include "conv_tc_prog_kernel.cuh"

namespace cg {
template int prog_launch_ks<1>(ProgPlan &, const ProgLaunchArgs &);
}  // namespace cg
