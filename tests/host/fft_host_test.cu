// Host-side unit test of the FFT building blocks in joxsz_b200/csrc/jx_fft.cuh.
// The functions are __host__ __device__, so the index algebra (radix-16 codelet, twiddles, exchange
// layout, output distribution) is checked on the CPU against a long-double direct DFT.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../joxsz_b200/csrc/jx_fft.cuh"

static double frand() { return (double)rand() / RAND_MAX - 0.5; }

int main() {
    srand(1234);
    double worst16 = 0.0, worst256 = 0.0;
    // --- DFT-16 codelet
    for (int rep = 0; rep < 20; ++rep) {
        double re[16], im[16], xr[16], xi[16];
        for (int i = 0; i < 16; ++i) { xr[i] = re[i] = frand(); xi[i] = im[i] = frand(); }
        dft16(re, im);
        for (int p = 0; p < 16; ++p) {
            int k = rev16(p);
            long double sr = 0, si = 0;
            for (int n = 0; n < 16; ++n) {
                long double a = -2.0L * M_PIl * n * k / 16.0L;
                sr += xr[n] * cosl(a) - xi[n] * sinl(a);
                si += xr[n] * sinl(a) + xi[n] * cosl(a);
            }
            worst16 = fmax(worst16, fmax(fabs((double)(sr - re[p])), fabs((double)(si - im[p]))));
        }
    }
    // --- FFT-256 by 16 emulated threads
    std::vector<double2> tw(256), xbuf(JX_XB_ELEMS);
    for (int i = 0; i < 256; ++i) fft256_make_twiddle(i, tw[i]);
    for (int rep = 0; rep < 5; ++rep) {
        std::vector<double> xr(256), xi(256);
        for (int i = 0; i < 256; ++i) { xr[i] = frand(); xi[i] = frand(); }
        double re[16][16], im[16][16];
        for (int t = 0; t < 16; ++t)
            for (int j = 0; j < 16; ++j) { re[t][j] = xr[t + 16 * j]; im[t][j] = xi[t + 16 * j]; }
        for (int t = 0; t < 16; ++t) fft256_pass1(t, re[t], im[t], tw.data(), xbuf.data());
        for (int t = 0; t < 16; ++t) fft256_pass2(t, re[t], im[t], xbuf.data());
        for (int q = 0; q < 16; ++q)
            for (int p = 0; p < 16; ++p) {
                int k = q + 16 * rev16(p);
                long double sr = 0, si = 0;
                for (int n = 0; n < 256; ++n) {
                    long double a = -2.0L * M_PIl * ((n * k) % 256) / 256.0L;
                    sr += xr[n] * cosl(a) - xi[n] * sinl(a);
                    si += xr[n] * sinl(a) + xi[n] * cosl(a);
                }
                worst256 = fmax(worst256, fmax(fabs((double)(sr - re[q][p])), fabs((double)(si - im[q][p]))));
            }
    }
    // --- two real even sequences in one complex transform: real part / imaginary part separate
    double worst_even = 0.0;
    {
        std::vector<double> a(129), b(129);
        for (int i = 0; i < 129; ++i) { a[i] = i < 86 ? frand() : 0.0; b[i] = i < 86 ? frand() : 0.0; }
        double re[16][16], im[16][16];
        for (int t = 0; t < 16; ++t)
            for (int j = 0; j < 16; ++j) { int f = fold256(t + 16 * j); re[t][j] = a[f]; im[t][j] = b[f]; }
        for (int t = 0; t < 16; ++t) fft256_pass1(t, re[t], im[t], tw.data(), xbuf.data());
        for (int t = 0; t < 16; ++t) fft256_pass2(t, re[t], im[t], xbuf.data());
        for (int q = 0; q < 16; ++q)
            for (int p = 0; p < 16; ++p) {
                int k = q + 16 * rev16(p);
                long double sa = 0, sb = 0;
                for (int n = 0; n < 256; ++n) {
                    long double c = cosl(2.0L * M_PIl * ((n * k) % 256) / 256.0L);
                    sa += a[fold256(n)] * c;
                    sb += b[fold256(n)] * c;
                }
                worst_even = fmax(worst_even, fmax(fabs((double)(sa - re[q][p])), fabs((double)(sb - im[q][p]))));
            }
    }
    // --- even sequences by 9 emulated threads (fft256e_pass1 + fft256_pass2): every folded index is produced
    double worst_e9 = 0.0;
    {
        std::vector<double2> xe(JX_XE_ELEMS);
        for (int rep = 0; rep < 4; ++rep) {
            std::vector<double> a(129), b(129);
            for (int i = 0; i < 129; ++i) { a[i] = (rep & 1) || i < 86 ? frand() : 0.0; b[i] = (rep & 1) || i < 86 ? frand() : 0.0; }
            double re[9][16], im[9][16];
            for (int t = 0; t < 9; ++t)
                for (int j = 0; j < 16; ++j) { int f = fold256(t + 16 * j); re[t][j] = a[f]; im[t][j] = b[f]; }
            for (auto& e : xe) e = make_double2(1e300, 1e300);        // poison: unread entries must not matter
            for (int t = 0; t < 9; ++t) fft256e_pass1<16>(t, re[t], im[t], tw.data() + t, xe.data());
            for (int t = 0; t < 9; ++t) fft256_pass2(t, re[t], im[t], xe.data());
            std::vector<int> seen(129, 0);
            for (int q = 0; q < 9; ++q)
                for (int p = 0; p < 16; ++p) {
                    int k = q + 16 * rev16(p);
                    seen[fold256(k)]++;
                    long double sa = 0, sb = 0;
                    for (int n = 0; n < 256; ++n) {
                        long double c = cosl(2.0L * M_PIl * ((n * k) % 256) / 256.0L);
                        sa += a[fold256(n)] * c;
                        sb += b[fold256(n)] * c;
                    }
                    worst_e9 = fmax(worst_e9, fmax(fabs((double)(sa - re[q][p])), fabs((double)(sb - im[q][p]))));
                }
            for (int f = 0; f < 129; ++f) if (!seen[f]) worst_e9 = 1.0;
        }
    }
    printf("dft16 %.3e fft256 %.3e even_pair %.3e even9 %.3e\n", worst16, worst256, worst_even, worst_e9);
    return (worst16 < 1e-14 && worst256 < 2e-13 && worst_even < 2e-13 && worst_e9 < 2e-13) ? 0 : 1;
}
