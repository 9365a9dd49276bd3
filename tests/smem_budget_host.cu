// Host-side print of the fused kernel's shared-memory footprint per (window, hop) shape, default and lean layout
// (csrc/fast_kernel.cuh: fast_smem_bytes).  Built and read by tests/test_smem_budget_cpu.py; no GPU needed.
#include <cstdio>
#define B200AA_LAYOUT_ONLY 1      // skip the launchers: they would instantiate every kernel
#include "../pyaudioanalysis_b200/csrc/fast_kernel.cuh"
using namespace b200aa;

template <int R1, int R2>
static void row(int step, int blob_words)
{
    constexpr int N = 2 * R1 * R2;
    const bool runs = (N % 80 == 0) && (step % 8 == 0);
    printf("%d %d %d %zu %zu\n", N, step, int(runs), fast_smem_bytes<R1, R2, B200AA_FAST_G>(step, blob_words, runs, false),
           runs ? fast_smem_bytes<R1, R2, B200AA_FAST_G>(step, blob_words, runs, true) : size_t(0));
}

int main()
{
    const int words = 1300;     // mel + DCT + chroma blob, upper bound over the supported (fs, window) pairs
    row<20, 20>(400, words); row<20, 20>(800, words); row<20, 20>(160, words);
    row<21, 21>(441, words); row<21, 21>(882, words);
    row<20, 10>(160, words); row<20, 10>(200, words); row<20, 10>(400, words);
    row<20, 12>(240, words); row<20, 12>(160, words);
    row<20, 15>(300, words);
    row<16, 10>(160, words); row<16, 10>(320, words);
    row<20, 16>(320, words); row<20, 16>(160, words);
    return 0;
}
