/* TEST INFRASTRUCTURE ONLY -- naive C restatement of the two SciPy transforms the
 * reference's path relies on (scipy.fftpack.fft at ShortTermFeatures.py:617 and
 * scipy.fftpack.dct(type=2, norm='ortho') at :253; SciPy is a third-party
 * dependency that is not vendored under /root/reference, requirements.txt pins
 * only scipy>=1.6.3).  Straight from the published definitions, O(N^2), double
 * precision; tests/test_oracle_golden.py pins numpy.fft and oracle.st_oracle's
 * DCT matrix against these.
 *
 * build:  gcc -O2 -shared -fPIC -o oracle/_build/libdftdef.so oracle/dft_def.c -lm
 */
#include <math.h>
#include <stddef.h>

/* X[k] = sum_n x[n] * exp(-2*pi*i*k*n/N), k = 0..N-1 (unnormalised forward DFT). */
void oracle_dft_real(const double *x, size_t n, double *out_re, double *out_im)
{
    const double two_pi = 6.283185307179586476925286766559;
    for (size_t k = 0; k < n; ++k) {
        double re = 0.0, im = 0.0;
        for (size_t j = 0; j < n; ++j) {
            /* reduce k*j modulo n before scaling so the angle stays accurate */
            double ang = two_pi * (double)((k * j) % n) / (double)n;
            re += x[j] * cos(ang);
            im -= x[j] * sin(ang);
        }
        out_re[k] = re;
        out_im[k] = im;
    }
}

/* Orthonormal DCT-II: y[0] = sqrt(1/N) sum x; y[k] = sqrt(2/N) sum x[n] cos(pi k (2n+1) / (2N)). */
void oracle_dct2_ortho(const double *x, size_t n, double *y, size_t n_out)
{
    const double pi = 3.14159265358979323846264338327950288;
    for (size_t k = 0; k < n_out; ++k) {
        double acc = 0.0;
        for (size_t j = 0; j < n; ++j)
            acc += x[j] * cos(pi * (double)k * (double)(2 * j + 1) / (2.0 * (double)n));
        y[k] = acc * (k == 0 ? sqrt(1.0 / (double)n) : sqrt(2.0 / (double)n));
    }
}
