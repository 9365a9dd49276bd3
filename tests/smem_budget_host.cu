// Host-side print of the shared-memory footprint of the CTA kernel per (window, hop) shape and of the pair kernel per window
// (csrc/fast_kernel.cuh: fast_smem_bytes).  Built and read by tests/test_smem_budget_cpu.py; no GPU needed.
#include <cstdio>
#define B200AA_LAYOUT_ONLY 1      // skip the launchers: they would instantiate every kernel
#include "../pyaudioanalysis_b200/csrc/fast_kernel.cuh"
#include "../pyaudioanalysis_b200/csrc/pair_kernel.cuh"
using namespace b200aa;

template <int R1, int R2>
static void row(int step, int blob_words)
{
    constexpr int N = 2 * R1 * R2;
    const bool runs = (N % 80 == 0) && (step % 8 == 0);
    printf("fast %d %d %d %zu\n", N, step, int(runs), fast_smem_bytes<R1, R2, B200AA_FAST_G>(step, blob_words, runs));
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
    // pair kernel: bytes per CTA (tables of <= 1 664 words = 6.5 KB) and warps per CTA
    const int pwords = 1664;
    printf("pair %d %d %zu\n", 320, pair_warps<10>(), pair_smem_bytes<10>(pwords));
    printf("pair %d %d %zu\n", 480, pair_warps<15>(), pair_smem_bytes<15>(pwords));
    printf("pair %d %d %zu\n", 512, pair_warps<16>(), pair_smem_bytes<16>(pwords));
    printf("pair %d %d %zu\n", 640, pair_warps<20>(), pair_smem_bytes<20>(pwords));
    printf("pair %d %d %zu\n", 800, pair_warps<25>(), pair_smem_bytes<25>(pwords));
    printf("pair %d %d %zu\n", 960, pair_warps<30>(), pair_smem_bytes<30>(pwords));
    printf("pair %d %d %zu\n", 1024, pair_warps<32>(), pair_smem_bytes<32>(pwords));
    return 0;
}
