// Register-resident FFT building blocks for the map stage (K3).
//
// A length-256 complex FFT is done by a group of 16 threads as two radix-16 passes with ONE
// exchange through shared memory (Cooley-Tukey, N = 16 x 16):
//
//   pass 1  thread t holds x[t + 16 j], j = 0..15    ->  A[t][k2] = sum_j x[t+16j] w16^(j k2)
//           twiddle A[t][k2] *= w256^(t k2), store to the exchange buffer
//   pass 2  thread q gathers A[.][q]                 ->  X[q + 16 k1] = sum_t A[t][q] w16^(t k1)
//
// so the output distribution (thread q holds k = q + 16 k1) has the same shape as the input
// distribution, and a second transform can follow without another exchange.  Every function is
// __host__ __device__ so the index algebra is unit-tested on the CPU (tests/host/fft_host_test.cu).
//
// Forward sign convention exp(-2 pi i n k / N).  All transforms of K3 act on even sequences, for
// which forward and inverse DFT coincide, so no inverse variant is needed.
#pragma once

#include "jx_common.cuh"

constexpr double JX_SQRT1_2 = 0.70710678118654752440;
constexpr double JX_COS_PI_8 = 0.92387953251128675613;
constexpr double JX_SIN_PI_8 = 0.38268343236508977173;

// position p of the in-place radix-4 x radix-4 DFT-16 holds output index rev16(p)
JX_HD constexpr int rev16(int p) { return (p >> 2) + 4 * (p & 3); }

// (r + i*im) *= exp(-2 pi i M / 16)
template <int M>
JX_HD void mul_w16(double& r, double& i) {
    constexpr int m = ((M % 16) + 16) % 16;
    if constexpr (m == 0) {
        return;
    } else if constexpr (m == 4) {          // -i
        double t = r; r = i; i = -t;
    } else if constexpr (m == 8) {          // -1
        r = -r; i = -i;
    } else if constexpr (m == 12) {         // +i
        double t = r; r = -i; i = t;
    } else {
        constexpr double c = (m == 1 || m == 15) ? JX_COS_PI_8 : (m == 2 || m == 14) ? JX_SQRT1_2
                           : (m == 3 || m == 13) ? JX_SIN_PI_8 : (m == 5 || m == 11) ? -JX_SIN_PI_8
                           : (m == 6 || m == 10) ? -JX_SQRT1_2 : -JX_COS_PI_8;   // 7, 9
        // exp(-i phi): imaginary part is -sin(phi)
        constexpr double s = (m == 1 || m == 7) ? -JX_SIN_PI_8 : (m == 2 || m == 6) ? -JX_SQRT1_2
                           : (m == 3 || m == 5) ? -JX_COS_PI_8 : (m == 9 || m == 15) ? JX_SIN_PI_8
                           : (m == 10 || m == 14) ? JX_SQRT1_2 : JX_COS_PI_8;    // 11, 13
        double t = r * c - i * s;
        i = r * s + i * c;
        r = t;
    }
}

// in-place forward DFT-4 of (x0, x1, x2, x3)
JX_HD void dft4(double& r0, double& i0, double& r1, double& i1, double& r2, double& i2, double& r3, double& i3) {
    double ar = r0 + r2, ai = i0 + i2;     // t0
    double br = r0 - r2, bi = i0 - i2;     // t1
    double cr = r1 + r3, ci = i1 + i3;     // t2
    double dr = r1 - r3, di = i1 - i3;     // t3
    r0 = ar + cr; i0 = ai + ci;
    r2 = ar - cr; i2 = ai - ci;
    r1 = br + di; i1 = bi - dr;            // t1 - i t3
    r3 = br - di; i3 = bi + dr;            // t1 + i t3
}

// in-place forward DFT-16: input natural order, output index rev16(p) at position p
JX_HD void dft16(double (&re)[16], double (&im)[16]) {
    // n = n1 + 4 n2: DFT-4 over n2 for each n1; position n1 + 4 k2 then holds B[n1][k2]
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1)
        dft4(re[n1], im[n1], re[n1 + 4], im[n1 + 4], re[n1 + 8], im[n1 + 8], re[n1 + 12], im[n1 + 12]);
    // twiddle by w16^(n1 k2)
    mul_w16<1>(re[5], im[5]);   mul_w16<2>(re[6], im[6]);   mul_w16<3>(re[7], im[7]);
    mul_w16<2>(re[9], im[9]);   mul_w16<4>(re[10], im[10]); mul_w16<6>(re[11], im[11]);
    mul_w16<3>(re[13], im[13]); mul_w16<6>(re[14], im[14]); mul_w16<9>(re[15], im[15]);
    // DFT-4 over n1 for each k2: position k1 + 4 k2 holds X[k2 + 4 k1]
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2)
        dft4(re[4 * k2], im[4 * k2], re[4 * k2 + 1], im[4 * k2 + 1], re[4 * k2 + 2], im[4 * k2 + 2],
             re[4 * k2 + 3], im[4 * k2 + 3]);
}

// exchange-buffer geometry of one 16-thread group: row pitch 17 complex keeps both the stores
// (threads consecutive in a row) and the loads (threads consecutive in a column) conflict free
constexpr int JX_XB_PITCH = 17;
constexpr int JX_XB_ELEMS = 16 * JX_XB_PITCH;     // double2 elements per group

// twiddle table w256^(t k2) laid out [k2][t]
JX_HD void fft256_make_twiddle(int idx, double2& w) {
    int k2 = idx >> 4, t = idx & 15;
    int m = (t * k2) & 255;
    // exact quadrant reduction keeps cos/sin arguments in [0, pi/4]
    double ang = -2.0 * 3.14159265358979323846 * (double)m / 256.0;
    w.x = cos(ang);
    w.y = sin(ang);
}

// pass 1 of thread t: re/im[j] = x[t + 16 j]
JX_HD void fft256_pass1(int t, double (&re)[16], double (&im)[16], const double2* __restrict__ tw,
                        double2* __restrict__ xbuf) {
    dft16(re, im);
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        const int k2 = rev16(p);
        double r = re[p], i = im[p];
        if (k2 != 0) {
            double2 w = tw[k2 * 16 + t];
            double tr = r * w.x - i * w.y;
            i = r * w.y + i * w.x;
            r = tr;
        }
        xbuf[k2 * JX_XB_PITCH + t] = make_double2(r, i);
    }
}

// pass 2 of thread q: afterwards position p holds X[q + 16 rev16(p)]
JX_HD void fft256_pass2(int q, double (&re)[16], double (&im)[16], const double2* __restrict__ xbuf, bool on = true) {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        double2 v = make_double2(0.0, 0.0);
        if (on) v = xbuf[q * JX_XB_PITCH + n1];
        re[n1] = v.x;
        im[n1] = v.y;
    }
    dft16(re, im);
}

// ---- even sequences: x[n] = x[256 - n] ---------------------------------------------------------------
// Every transform of the cyclic-length-256 map kernel acts on an even complex sequence (two real even
// sequences packed as real / imaginary part).  Pass 1 of thread 16 - t is then determined by thread t,
//   A_{16-t}[k2] w256^((16-t) k2) = conj(w256^(t k2)) A_t[-k2],
// and the spectrum is even as well, so a group of NINE threads (t = 0..8) does the whole transform: thread t
// writes its own exchange column and the mirrored one, thread q = 0..8 of pass 2 ends with X[q + 16 k1],
// k1 = 0..15, and those 9 x 16 values cover every index of the folded half-array (fold256).  Three groups
// share a warp (27 of 32 lanes).  Only exchange rows k2 = 0..8 are ever read.
constexpr int JX_XE_ROWS = 9;
constexpr int JX_XE_ELEMS = JX_XE_ROWS * JX_XB_PITCH;     // double2 elements per 9-thread group

// `tw_t` points at this thread's twiddles, w256^(t k2) at tw_t[k2 * TWS]: TWS = 16 with the [k2][t] table of
// fft256_make_twiddle (tw + t), or a per-lane copy [k2][lane] (TWS = 32) whose quarter-warps read 8 consecutive
// 16-byte entries -- in the [k2][t] table thread 8 and thread 0 of neighbouring groups share a bank.
// `on` = false (a lane that only shadows another one) skips the exchange stores.
// twiddles already in registers: w[k2] = w256^(t k2), k2 = 1..8 (w[0] unused)
JX_HD void fft256e_pass1_w(int t, double (&re)[16], double (&im)[16], const double2 (&w)[JX_XE_ROWS],
                           double2* __restrict__ xbuf, bool on = true) {
    dft16(re, im);
    const bool mirror = on && t >= 1 && t <= 7;
#pragma unroll
    for (int k2 = 0; k2 < JX_XE_ROWS; ++k2) {
        const int pd = rev16(k2), pm = rev16((16 - k2) & 15);
        double r = re[pd], i = im[pd], mr = re[pm], mi = im[pm];
        if (k2 != 0) {
            const double tr = r * w[k2].x - i * w[k2].y;
            i = r * w[k2].y + i * w[k2].x;
            r = tr;
            const double tm = mr * w[k2].x + mi * w[k2].y;        // times conj(w)
            mi = mi * w[k2].x - mr * w[k2].y;
            mr = tm;
        }
        if (on) xbuf[k2 * JX_XB_PITCH + t] = make_double2(r, i);
        if (mirror) xbuf[k2 * JX_XB_PITCH + 16 - t] = make_double2(mr, mi);
    }
}
template <int TWS>
JX_HD void fft256e_pass1(int t, double (&re)[16], double (&im)[16], const double2* __restrict__ tw_t,
                         double2* __restrict__ xbuf, bool on = true) {
    // the twiddles are loaded before the butterflies so that their shared-memory latency hides behind them
    double2 w[JX_XE_ROWS];
    w[0] = make_double2(1.0, 0.0);
#pragma unroll
    for (int k2 = 1; k2 < JX_XE_ROWS; ++k2) w[k2] = tw_t[k2 * TWS];
    fft256e_pass1_w(t, re, im, w, xbuf, on);
}
// pass 2 of thread q = 0..8 is fft256_pass2 (it reads exchange row q only)

// index of the even extension: sequence value at n (0 <= n < 256) is the half-array value at fold256(n)
JX_HD constexpr int fold256(int n) { return n <= 128 ? n : 256 - n; }
