#pragma once
// TEST INFRASTRUCTURE (oracle build only): minimal stand-in for tbb::parallel_for(first,last,step,f).
// TBB is the one un-vendored dependency of the reference (cmake/OptCutsDownloadExternal.cmake:33-40,
// wjakob/tbb@344fa84); it carries no arithmetic (7 call sites, disjoint writes), so this shim is
// result-identical.  Outermost call: std::thread workers pulling dynamic chunks; nested call: serial.
// ORACLE_THREADS=N overrides the worker count (default: hardware_concurrency).
#include <thread>
#include <vector>
#include <atomic>
#include <cstdlib>
#include <algorithm>
namespace tbb {
inline int shim_threads() {
    static int n = []() {
        const char* e = std::getenv("ORACLE_THREADS");
        int v = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
        return v < 1 ? 1 : v;
    }();
    return n;
}
inline bool& shim_in_parallel() { static thread_local bool b = false; return b; }
template <typename Index, typename F>
void parallel_for(Index first, Index last, Index step, const F& f) {
    const long n = (last - first + step - 1) / step;
    if (n <= 0) return;
    const int nt = (int)std::min<long>(shim_threads(), n);
    if (nt <= 1 || shim_in_parallel()) {
        for (Index i = first; i < last; i += step) f(i);
        return;
    }
    const long chunk = std::max<long>(1, n / (nt * 8L));
    std::atomic<long> next(0);
    auto work = [&]() {
        shim_in_parallel() = true;
        for (;;) {
            long b = next.fetch_add(chunk);
            if (b >= n) break;
            long e = std::min(n, b + chunk);
            for (long k = b; k < e; ++k) f((Index)(first + k * step));
        }
        shim_in_parallel() = false;
    };
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
}
}  // namespace tbb
