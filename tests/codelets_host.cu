// Host-side check of the register FFT codelets (csrc/dft_codelets.cuh) and of the R1 x R2 packed-real transform the
// fused kernel builds from them: same templates, same index maps, executed on the CPU and compared with a naive
// float64 DFT.  Built and run by tests/test_codelets_cpu.py (nvcc host compile; no GPU needed).
#include <cmath>
#include <cstdio>
#include <vector>
#include "../pyaudioanalysis_b200/csrc/dft_codelets.cuh"

using namespace b200aa;

static double g_worst = 0.0;

static void report(const char *what, int a, int b, double err)
{
    printf("%s %d %d %.3e\n", what, a, b, err);
    if (err > g_worst) g_worst = err;
}

// complex product exactly as common.cuh's cmul (device-only there)
static float2 hmul(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }

static unsigned lcg(unsigned &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int R>
static void check_codelet()
{
    unsigned seed = 1234u + R;
    float2 v[R];
    std::vector<double> re(R), im(R);
    for (int n = 0; n < R; ++n) {
        re[n] = (double(lcg(seed) % 20001) - 10000.0) / 10000.0;
        im[n] = (double(lcg(seed) % 20001) - 10000.0) / 10000.0;
        v[n] = make_float2(float(re[n]), float(im[n]));
        re[n] = v[n].x; im[n] = v[n].y;
    }
    fft_r<R>(v);
    double err = 0.0, mag = 0.0;
    for (int k = 0; k < R; ++k) {
        double sr = 0.0, si = 0.0;
        for (int n = 0; n < R; ++n) {
            const double a = -2.0 * M_PI * double((k * n) % R) / double(R);
            sr += re[n] * std::cos(a) - im[n] * std::sin(a);
            si += re[n] * std::sin(a) + im[n] * std::cos(a);
        }
        err = std::fmax(err, std::fmax(std::fabs(sr - v[k].x), std::fabs(si - v[k].y)));
        mag = std::fmax(mag, std::hypot(sr, si));
    }
    report("codelet", R, 0, err / mag);
}

// the kernel's transform of one real frame of N = 2*R1*R2 samples: |X[k]|, k < Nc
template <int R1, int R2>
static void check_shape()
{
    constexpr int Nc = R1 * R2, N = 2 * Nc;
    unsigned seed = 99u + R1 * 100 + R2;
    std::vector<float> x(N);
    for (int n = 0; n < N; ++n) x[n] = float((int(lcg(seed) % 65536) - 32768));       // int16-valued samples
    const float d0 = x[0];
    std::vector<float2> E(size_t(R1) * R2), Z(Nc);
    // pass 1: column n2, R1-point transform over n1 of z[R2*n1 + n2], twiddle W_Nc^(k1*n2)
    for (int n2 = 0; n2 < R2; ++n2) {
        float2 v1[R1];
        for (int n1 = 0; n1 < R1; ++n1) {
            const int n = R2 * n1 + n2;
            v1[n1] = make_float2(x[2 * n] - d0, x[2 * n + 1] - d0);
        }
        fft_r<R1>(v1);
        for (int k1 = 0; k1 < R1; ++k1) {
            const double a = -2.0 * M_PI * double((k1 * n2) % Nc) / double(Nc);
            const float2 w = make_float2(float(std::cos(a)), float(std::sin(a)));
            E[size_t(k1) * R2 + n2] = k1 == 0 ? v1[0] : hmul(v1[k1], w);
        }
    }
    // pass 2: row k1, R2-point transform over n2 -> Z[k1 + R1*k2]
    for (int k1 = 0; k1 < R1; ++k1) {
        float2 v[R2];
        for (int n2 = 0; n2 < R2; ++n2) v[n2] = E[size_t(k1) * R2 + n2];
        fft_r<R2>(v);
        for (int k2 = 0; k2 < R2; ++k2) Z[k1 + R1 * k2] = v[k2];
    }
    // post-processing: X[k] = (ev + W_N^k od) / 2 from (Z[k], Z[Nc-k]); DC from the plain sums
    std::vector<double> got(Nc);
    got[0] = std::fabs(double(Z[0].x) + double(Z[0].y) + double(N) * d0);
    for (int k = 1; k < Nc; ++k) {
        const float2 zk = Z[k], zp = Z[Nc - k];
        const float2 ev = make_float2(zk.x + zp.x, zk.y - zp.y);
        const float2 od = make_float2(zk.y + zp.y, zp.x - zk.x);
        const double a = -2.0 * M_PI * double(k) / double(N);
        const float2 t = hmul(od, make_float2(float(std::cos(a)), float(std::sin(a))));
        got[k] = 0.5 * std::hypot(double(ev.x + t.x), double(ev.y + t.y));
    }
    double err = 0.0, mag = 0.0;
    for (int k = 0; k < Nc; ++k) {
        double sr = 0.0, si = 0.0;
        for (int n = 0; n < N; ++n) {
            const double a = -2.0 * M_PI * double((long(k) * n) % N) / double(N);
            sr += x[n] * std::cos(a);
            si += x[n] * std::sin(a);
        }
        const double ref = std::hypot(sr, si);
        err = std::fmax(err, std::fabs(ref - got[k]));
        mag = std::fmax(mag, ref);
    }
    report("shape", R1, R2, err / mag);
}

// the two-sequences-per-register 32-point codelet of the pair kernel's second pass
static void check_fft32_soa()
{
    unsigned seed = 777u;
    float2 re[16], im[16], out[32];
    double xr[32], xi[32];
    for (int n = 0; n < 32; ++n) {
        const float a = float((double(lcg(seed) % 20001) - 10000.0) / 10000.0), b = float((double(lcg(seed) % 20001) - 10000.0) / 10000.0);
        xr[n] = a; xi[n] = b;
        if (n & 1) { re[n / 2].y = a; im[n / 2].y = b; } else { re[n / 2].x = a; im[n / 2].x = b; }
    }
    fft32_soa(re, im, out);
    double err = 0.0, mag = 0.0;
    for (int k = 0; k < 32; ++k) {
        double sr = 0.0, si = 0.0;
        for (int n = 0; n < 32; ++n) {
            const double a = -2.0 * M_PI * double((k * n) % 32) / 32.0;
            sr += xr[n] * std::cos(a) - xi[n] * std::sin(a);
            si += xr[n] * std::sin(a) + xi[n] * std::cos(a);
        }
        err = std::fmax(err, std::fmax(std::fabs(sr - out[k].x), std::fabs(si - out[k].y)));
        mag = std::fmax(mag, std::hypot(sr, si));
    }
    report("soa32", 32, 0, err / mag);
}

// the pair kernel's transform (csrc/pair_kernel.cuh): two real frames of N = 32 * R samples through ONE complex FFT,
// n = 32 n1 + n2, k = k1 + R k2, |Xa[k]| = |Z[k] + conj Z[N-k]| / 2 sa, |Xb[k]| = |Z[k] - conj Z[N-k]| / 2 sb;
// "lanes" are loops here, the index maps and codelets are the kernel's
template <int R>
static void check_pair()
{
    constexpr int N = 32 * R, K = N / 2;
    unsigned seed = 4242u + R;
    std::vector<float> xa(N), xb(N);
    for (int n = 0; n < N; ++n) {
        xa[n] = float((int(lcg(seed) % 65536) - 32768));          // loud frame
        xb[n] = float((int(lcg(seed) % 61) - 30) + 7);              // quiet frame (60 dB below) with a DC offset
    }
    const float sa = 1.f / 16384.f, sb = 1.f / 16.f;                 // per-frame power-of-two scales
    std::vector<float2> T(size_t(R) * 33), Z(N + 1);
    for (int n2 = 0; n2 < 32; ++n2) {                                // pass 1: lane n2
        float2 z[R];
        for (int r = 0; r < R; ++r) z[r] = make_float2(sa * (xa[32 * r + n2] - xa[0]), sb * (xb[32 * r + n2] - xb[0]));
        fft_r<R>(z);
        for (int k1 = 0; k1 < R; ++k1) {
            const double a = -2.0 * M_PI * double((k1 * n2) % N) / double(N);
            T[size_t(k1) * 33 + n2] = k1 == 0 ? z[0] : hmul(z[k1], make_float2(float(std::cos(a)), float(std::sin(a))));
        }
    }
    for (int k1 = 0; k1 < R; ++k1) {                                 // pass 2: lane k1
        float2 re[16], im[16], v[32];
        for (int m = 0; m < 16; ++m) {
            re[m] = make_float2(T[size_t(k1) * 33 + 2 * m].x, T[size_t(k1) * 33 + 2 * m + 1].x);
            im[m] = make_float2(T[size_t(k1) * 33 + 2 * m].y, T[size_t(k1) * 33 + 2 * m + 1].y);
        }
        fft32_soa(re, im, v);
        for (int k2 = 0; k2 < 32; ++k2) Z[k1 + R * k2] = v[k2];
    }
    Z[N] = Z[0];
    double erra = 0.0, errb = 0.0, rmsa = 0.0, rmsb = 0.0;
    for (int k = 1; k < K; ++k) {
        const float2 zk = Z[k], pk = Z[N - k];
        const double ga = std::hypot(double(zk.x + pk.x), double(zk.y - pk.y)) / (2.0 * sa);
        const double gb = std::hypot(double(zk.x - pk.x), double(zk.y + pk.y)) / (2.0 * sb);
        double ar = 0, ai = 0, br = 0, bi = 0;
        for (int n = 0; n < N; ++n) {
            const double a = -2.0 * M_PI * double((long(k) * n) % N) / double(N);
            ar += xa[n] * std::cos(a); ai += xa[n] * std::sin(a);
            br += xb[n] * std::cos(a); bi += xb[n] * std::sin(a);
        }
        const double ra = std::hypot(ar, ai), rb = std::hypot(br, bi);
        erra = std::fmax(erra, std::fabs(ra - ga)); errb = std::fmax(errb, std::fabs(rb - gb));
        rmsa += ra * ra; rmsb += rb * rb;
    }
    // error of either spectrum relative to its OWN rms level (the quiet frame must not inherit the loud one's error)
    report("pair", R, 0, erra / std::sqrt(rmsa / K));
    report("pair", R, 1, errb / std::sqrt(rmsb / K));
}

int main()
{
    check_fft32_soa();
    check_pair<10>(); check_pair<15>(); check_pair<20>(); check_pair<25>(); check_pair<30>();
    check_codelet<25>(); check_codelet<30>(); check_codelet<32>();
    check_codelet<10>(); check_codelet<12>(); check_codelet<15>(); check_codelet<16>(); check_codelet<20>(); check_codelet<21>();
    check_shape<20, 20>(); check_shape<21, 21>(); check_shape<20, 10>(); check_shape<20, 12>();
    check_shape<20, 15>(); check_shape<16, 10>(); check_shape<20, 16>();
    check_shape<10, 20>(); check_shape<15, 20>();     // the solo kernel's 400 / 600 windows (R1 = points per lane, R2 = lanes)
    printf("worst %.3e\n", g_worst);
    return g_worst < 2e-6 ? 0 : 1;
}
