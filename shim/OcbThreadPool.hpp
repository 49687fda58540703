// Persistent host thread pool of the shim build (header only).  Work items of the host program's parallel loops are short
// (a Triangle call on a 20-vertex air region: ~20 us; one candidate's local mesh: ~50 us) and the loops are entered hundreds
// of thousands of times per run, so the workers are created once and sleep on a condition variable between loops; the calling
// thread takes part in the work.  One loop at a time: a loop entered from inside a worker, or while another thread's loop is in
// flight, runs serially in its caller (the reference's loops write disjoint slots, so this is result-neutral).
#ifndef OcbThreadPool_hpp
#define OcbThreadPool_hpp

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace OptCuts {

class OcbThreadPool {
public:
    static OcbThreadPool& get(void) { static OcbThreadPool* p = new OcbThreadPool(); return *p; }     // never destroyed: the host program leaves through exit()
    int threads(void) const { return static_cast<int>(workers.size()) + 1; }
    void run(int n, const std::function<void(int)>& f, int serialBelow = 4)
    {
        if (n <= 0) return;
        if (n < serialBelow || workers.empty() || inWorker() || !busy.try_lock()) { for (int i = 0; i < n; ++i) f(i); return; }
        {
            std::lock_guard<std::mutex> lock(mu);
            fn = &f; total = n; chunk = n / (8 * threads()); if (chunk < 1) chunk = 1;
            next.store(0); pending.store(n); ++generation;
        }
        cv.notify_all();
        inWorker() = true;
        work();
        inWorker() = false;
        while (pending.load(std::memory_order_acquire) > 0) std::this_thread::yield();
        { std::lock_guard<std::mutex> lock(mu); fn = NULL; }                             // workers that wake up from now on skip this job
        while (inside.load(std::memory_order_acquire) > 0) std::this_thread::yield();    // ... and those already in have left before the next job is posted
        busy.unlock();
    }
private:
    OcbThreadPool(void)
    {
        const char* e = std::getenv("OCB_HOST_THREADS");
        const char* o = std::getenv("ORACLE_THREADS");                                   // the oracle build's knob is honoured too
        int nt = e ? std::atoi(e) : (o ? std::atoi(o) : static_cast<int>(std::thread::hardware_concurrency()));
        if (nt > 64) nt = 64;
        for (int t = 1; t < nt; ++t) workers.emplace_back([this]() { loop(); });
        for (auto& w : workers) w.detach();
    }
    static bool& inWorker(void) { static thread_local bool b = false; return b; }
    void work(void)
    {
        for (;;) {                                   // dynamic chunks: items of one loop can differ a lot in cost
            const int b = next.fetch_add(chunk);
            if (b >= total) break;
            const int e = b + chunk < total ? b + chunk : total;
            for (int i = b; i < e; ++i) (*fn)(i);
            pending.fetch_sub(e - b, std::memory_order_release);
        }
    }
    void loop(void)
    {
        inWorker() = true;
        unsigned long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lock(mu);
                cv.wait(lock, [&]() { return generation != seen; });
                seen = generation;
                if (!fn) continue;
                inside.fetch_add(1);
            }
            work();
            inside.fetch_sub(1, std::memory_order_release);
        }
    }
    std::vector<std::thread> workers;
    std::mutex mu, busy; std::condition_variable cv;
    const std::function<void(int)>* fn = NULL;
    int total = 0, chunk = 1; unsigned long generation = 0;
    std::atomic<int> next{0}, pending{0}, inside{0};
};

}  // namespace OptCuts
#endif
