// CPU exercise of csrc/sched.cuh (static shares + steal-half): one thread per "warp", every pair step of the launch must be
// claimed exactly once whatever the interleaving.  Built and run by tests/test_sched_cpu.py (g++, no GPU).
//   usage: sched_host n_warps total chunk min_steal slow_every
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include "../pyaudioanalysis_b200/csrc/sched.cuh"
using namespace b200aa;

int main(int argc, char **argv)
{
    const unsigned n_warps = argc > 1 ? atoi(argv[1]) : 64, total = argc > 2 ? atoi(argv[2]) : 100000;
    const unsigned chunk = argc > 3 ? atoi(argv[3]) : 4, min_steal = argc > 4 ? atoi(argv[4]) : 6, slow_every = argc > 5 ? atoi(argv[5]) : 5;
    std::vector<unsigned long long> ranges(n_warps, 0ull);
    std::vector<std::atomic<unsigned char>> hit(total ? total : 1);
    for (auto &h : hit) h.store(0);
    StealParams sp{ranges.data(), n_warps, total, 200u, chunk, min_steal};
    std::atomic<unsigned long long> steals{0}, chunks{0}, starts{0};
    auto worker = [&](unsigned w) {
        sched_begin(sp, w);
        unsigned last_end = ~0u, inc = chunk;
        for (;;) {
            unsigned g0 = 0, g1 = 0;
            if (sched_claim(sp, w, inc, g0, g1)) {
                if (g0 >= g1 || g1 > total || g1 - g0 > chunk) { fprintf(stderr, "bad chunk [%u, %u)\n", g0, g1); exit(2); }
                for (unsigned g = g0; g < g1; ++g) hit[g].fetch_add(1);
                chunks.fetch_add(1);
                if (g0 != last_end) starts.fetch_add(1);       // a run starts here (this is where the kernel pays a halo)
                last_end = g1;
                if (slow_every && (w % slow_every) == 0) std::this_thread::yield();      // uneven speeds force steals
                continue;
            }
            // the device scans 32 descriptors at a time; any victim choice is valid, the host takes the largest remainder
            bool got = false;
            for (int attempt = 0; attempt < 3 && !got; ++attempt) {
                unsigned best = 0, bv = 0;
                unsigned long long braw = 0;
                for (unsigned v = 0; v < n_warps; ++v) {
                    if (v == w) continue;
                    const unsigned long long raw = sched_load(ranges.data() + v);
                    const unsigned rem = sched_remainder(sp, raw);
                    if (rem > best) { best = rem; bv = v; braw = raw; }
                }
                if (!best) break;
                got = sched_try_steal(sp, w, bv, braw);
            }
            if (!got) return;
            inc = 1;
            steals.fetch_add(1);
        }
    };
    std::vector<std::thread> th;
    for (unsigned w = 0; w < n_warps; ++w) th.emplace_back(worker, w);
    for (auto &t : th) t.join();
    unsigned long long missed = 0, dup = 0;
    for (unsigned g = 0; g < total; ++g) { const int h = hit[g].load(); if (h == 0) ++missed; else if (h > 1) ++dup; }
    printf("{\"n_warps\": %u, \"total\": %u, \"missed\": %llu, \"duplicated\": %llu, \"steals\": %llu, \"chunks\": %llu, \"run_starts\": %llu}\n",
           n_warps, total, missed, dup, steals.load(), chunks.load(), starts.load());
    return (missed || dup) ? 1 : 0;
}
