// inst_dist.cu -- compiled once per distribution id (-DGENDR_DIST=k): instantiates the four render kernels
// (forward/backward x simple/parametric t-conorm) for that distribution and exports their launcher.
#include "render_kernels.cuh"
#ifndef GENDR_DIST
#error "compile with -DGENDR_DIST=<0..17>"
#endif
#define GENDR_CAT2(a, b) a##b
#define GENDR_CAT(a, b) GENDR_CAT2(a, b)
namespace gendr {
cudaError_t GENDR_CAT(launch_render_dist_, GENDR_DIST)(const RenderParams& P, const KernelIO& io, const LaunchCfg& cfg) {
    return launch_render_for_dist<GENDR_DIST>(P, io, cfg);
}
}  // namespace gendr
