// Instantiations of the strided tap-program kernel for KSTEPS = 8 (Cin = 128); see conv_tc_prog_kernel.cuh.
#include "conv_tc_prog_kernel.cuh"

namespace cg {
template int prog_launch_ks<8>(ProgPlan &, const ProgLaunchArgs &);
}  // namespace cg
