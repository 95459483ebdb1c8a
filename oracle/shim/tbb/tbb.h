// NOT reference code. Stand-in for <tbb/tbb.h> (TBB is not installed in this image).
// The reference only uses the 4-argument tbb::parallel_for(first, last, step, body) form
// (137 call sites). Here it splits the range statically over
// BMQ_SHIM_THREADS std::threads (default: hardware_concurrency; 1 = serial).
#pragma once
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <thread>
#include <vector>
#include <algorithm>
namespace tbb {
inline int shim_threads() {
    static int n = [] {
        const char *e = std::getenv("BMQ_SHIM_THREADS");
        int v = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
        return v > 0 ? v : 1;
    }();
    return n;
}
template <typename Index, typename Body>
void parallel_for(Index first, Index last, Index step, const Body &body) {
    if (last <= first) return;
    long long count = ((long long)last - (long long)first + (long long)step - 1) / (long long)step;
    int nt = (int)std::min<long long>(shim_threads(), count);
    if (nt <= 1) { for (Index i = first; i < last; i += step) body(i); return; }
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int t = 0; t < nt; ++t) {
        long long b = count * t / nt, e = count * (t + 1) / nt;
        pool.emplace_back([=, &body] {
            for (long long c = b; c < e; ++c) body((Index)(first + (Index)c * step));
        });
    }
    for (auto &th : pool) th.join();
}
template <typename I1, typename I2, typename I3, typename Body>
void parallel_for(I1 first, I2 last, I3 step, const Body &body) {
    parallel_for<int, Body>((int)first, (int)last, (int)step, body);
}
}  // namespace tbb
