// Host-side constant tables (double precision), shared by the C ABI and the plan builder.
// Each builder restates the table the reference derives per call / per frame:
//   mel filterbank   -- mfcc_filter_banks,      ShortTermFeatures.py:191-233
//   chroma operator  -- chroma_features_init +  the scatter in chroma_features, :257-302
//   DCT-II (ortho)   -- scipy.fftpack dct(type=2, norm='ortho') call at :253
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../include/b200aa.h"

namespace b200aa_host {

static const double kPi = 3.14159265358979323846264338327950288;

// [40 x K] dense.  Bin grid k*fs/K and edge indices floor(f*K/fs)+1: the reference hands
// K = window/2 to a routine that expects the FFT size; that 2x-wide axis is load-bearing.
inline int build_mel(int fs, int K, std::vector<double> &bank)
{
    const int n_lin = 13, n_log = 27, n_filt = n_lin + n_log;
    const double low = 133.33, lin_step = 200.0 / 3.0, log_ratio = 1.0711703;
    std::vector<double> edge(n_filt + 2);
    for (int i = 0; i < n_lin; ++i) edge[i] = low + i * lin_step;
    for (int i = n_lin; i < n_filt + 2; ++i) edge[i] = edge[n_lin - 1] * std::pow(log_ratio, double(i - n_lin + 1));
    bank.assign(size_t(n_filt) * K, 0.0);
    for (int i = 0; i < n_filt; ++i) {
        const double lo = edge[i], ce = edge[i + 1], hi = edge[i + 2];
        const double peak = 2.0 / (hi - lo);
        const long k_lo = long(std::floor(lo * K / fs)) + 1;
        const long k_ce = long(std::floor(ce * K / fs)) + 1;
        const long k_hi = long(std::floor(hi * K / fs)) + 1;
        if (k_hi - 1 >= K && k_hi > k_lo) return B200AA_ERR_MEL_RANGE;   // reference: IndexError
        const double up = peak / (ce - lo), down = peak / (hi - ce);
        for (long k = k_lo; k < k_ce; ++k) bank[size_t(i) * K + k] = up * (double(k) / K * fs - lo);
        for (long k = k_ce; k < k_hi; ++k) bank[size_t(i) * K + k] = down * (hi - double(k) / K * fs);
    }
    return B200AA_OK;
}

// semitone index per bin and how many bins share it (chroma_features_init)
inline void chroma_index(int fs, int K, std::vector<long> &semi, std::vector<double> &share)
{
    semi.resize(K);
    share.assign(K, 0.0);
    for (int k = 0; k < K; ++k) {
        const double f = (double(k + 1) * fs) / (2.0 * K);
        semi[k] = long(std::nearbyint(12.0 * std::log2(f / 27.50)));   // np.round: half to even
    }
    for (int k = 0; k < K; ++k) {
        int cnt = 0;
        for (int j = 0; j < K; ++j) cnt += (semi[j] == semi[k]);
        share[k] = cnt;
    }
}

// [12 x K] dense operator M with chroma = (M @ X^2) / sum(X^2).
// `C[semi] = X^2` is a NumPy fancy store: last source bin wins per target slot, negative slots
// wrap; `C /= share[semi]` divides slot j by share[semi[j]] (wrapping again); slots fold mod 12.
inline int build_chroma(int fs, int K, std::vector<double> &op)
{
    std::vector<long> semi;
    std::vector<double> share;
    chroma_index(fs, K, semi, share);
    long top = semi[0];
    for (int k = 1; k < K; ++k) top = semi[k] > top ? semi[k] : top;
    if (!(top < K)) return B200AA_ERR_CHROMA;
    auto wrap = [K](long j) -> long { return j < 0 ? j + K : j; };
    std::vector<long> winner(K, -1);
    for (int k = 0; k < K; ++k) {
        const long slot = wrap(semi[k]);
        if (slot < 0 || slot >= K) return B200AA_ERR_CHROMA;            // numpy: IndexError
        winner[slot] = k;
    }
    op.assign(size_t(12) * K, 0.0);
    for (int j = 0; j < K; ++j) {
        if (winner[j] < 0) continue;
        const long d = wrap(semi[j]);
        op[size_t(j % 12) * K + winner[j]] += 1.0 / share[d];
    }
    return B200AA_OK;
}

inline void build_dct(std::vector<double> &mat)
{
    const int n_in = B200AA_N_MEL, n_out = B200AA_N_MFCC;
    mat.assign(size_t(n_out) * n_in, 0.0);
    for (int k = 0; k < n_out; ++k)
        for (int n = 0; n < n_in; ++n)
            mat[size_t(k) * n_in + n] = (k == 0) ? std::sqrt(1.0 / n_in)
                                                 : std::sqrt(2.0 / n_in) * std::cos(kPi * k * (2 * n + 1) / (2.0 * n_in));
}

inline int64_t num_frames(int64_t n, int w, int s) { return n < w ? 0 : (n - w) / s + 1; }

// len(range(a, b, s)) for s > 0
inline int64_t range_len(int64_t a, int64_t b, int64_t s) { return b > a ? (b - a + s - 1) / s : 0; }

// prime-factor radix list for the generic Stockham transform (4s first, then primes ascending)
inline std::vector<int> radix_list(int n)
{
    std::vector<int> r;
    while (n % 4 == 0) { r.push_back(4); n /= 4; }
    for (int p = 2; p * p <= n; ++p)
        while (n % p == 0) { r.push_back(p); n /= p; }
    if (n > 1) r.push_back(n);
    return r;
}

}  // namespace b200aa_host
