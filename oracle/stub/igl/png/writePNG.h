#pragma once
// TEST INFRASTRUCTURE (oracle build only): no-op igl::png::writePNG (no libpng headless).
#include <Eigen/Core>
#include <string>
namespace igl { namespace png {
template <class M>
inline bool writePNG(const M&, const M&, const M&, const M&, const std::string&) { return true; }
}}
