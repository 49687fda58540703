// tbb::parallel_for(first, last, step, f) for the shim build of the host program (shim/Makefile puts this directory ahead of the
// oracle's stand-in on the include path).  TBB is the reference's one un-vendored dependency (cmake/OptCutsDownloadExternal.cmake:
// 33-40) and is not in this image; its 7 call sites are loops with disjoint writes.  This version runs them on the shim's
// persistent thread pool (OcbThreadPool.hpp) -- the oracle's stand-in creates its threads per call, which is fine for the
// reference arm (a few calls per 60 ms Newton iteration) and not for a host program whose Newton iteration takes 2 ms.
#pragma once
#include "../OcbThreadPool.hpp"
namespace tbb {
template <typename Index, typename F>
void parallel_for(Index first, Index last, Index step, const F& f)
{
    const long n = (static_cast<long>(last) - static_cast<long>(first) + static_cast<long>(step) - 1) / static_cast<long>(step);
    if (n <= 0) return;
    OptCuts::OcbThreadPool::get().run(static_cast<int>(n), [&](int k) { f(static_cast<Index>(first + static_cast<Index>(k) * step)); }, 2);
}
}  // namespace tbb
