// Instantiations of the strided tap-program kernel for KSTEPS = 2 (Cin = 32); see conv_tc_prog_kernel.cuh.
#include "conv_tc_prog_kernel.cuh"

namespace cg {
template int prog_launch_ks<2>(ProgPlan &, const ProgLaunchArgs &);
}  // namespace cg
