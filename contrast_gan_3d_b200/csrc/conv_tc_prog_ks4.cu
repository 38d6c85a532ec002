// Instantiations of the strided tap-program kernel for KSTEPS = 4 (Cin = 64); see conv_tc_prog_kernel.cuh.
#include "conv_tc_prog_kernel.cuh"

namespace cg {
template int prog_launch_ks<4>(ProgPlan &, const ProgLaunchArgs &);
}  // namespace cg
