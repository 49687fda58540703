// Process-wide slot for the latest UVs of the mesh under optimisation: written by CudaSymDirichletEnergy::bind, read by
// CudaLinSysSolver::set_pattern.  OptCuts::LinSysSolver's interface carries no geometry (LinSysSolver.hpp:37-135), and
// the CUDA solver's preconditioner hierarchy wants some (ocb_set_coordinate_hint); results do not depend on it.
#pragma once
#include <mutex>
#include <vector>
namespace OptCuts {
inline void cudaCoordinateHint(bool write, int& n, std::vector<double>& xy)
{
    static std::mutex mu; static std::vector<double> slot; static int slotN = 0;
    std::lock_guard<std::mutex> lock(mu);
    if (write) { slot = xy; slotN = n; } else { xy = slot; n = slotN; }
}
}  // namespace OptCuts
