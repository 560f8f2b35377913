// Device helpers shared by the two map kernels (k3_szmap.cu: maps whose cyclic length is 256, whole pipeline in
// shared memory; k3l_szmap.cu: larger maps, working set in an L2-resident global scratch).
#pragma once

#include "jx_fft.cuh"

constexpr int JX_D_MAXSPLIT = 8;

struct k3_args {
    jx_dev d;
    const double* coef;
    const uint32_t* flags;
    int W;
    double *convq, *g;    // g: large-map path (filter stage in the kernel); convq: parity tap
    double* tri;          // shared-memory path: packed u <= v triangle of the convolved map, [W][d.ktri]
    double* scratch;      // large-map path only: [gridDim.x][hp8][pitch] doubles
    double* scratch2;     // large-map path, direct y convolution: second map of the same shape (NULL: FFT form)
};

JX_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

JX_D void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
JX_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
JX_D void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
JX_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "JX_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra JX_DONE;\n"
        "bra JX_WAIT;\n"
        "JX_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

JX_D void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// value of the Compton-y spline on piece `s` at offset `t` from its left knot, Horner evaluation.  Coefficient
// layout: two planes of 16-byte entries, (c0, c1) of piece s at c[2 s] and (c2, c3) at c[2 nseg + 2 s], so that
// lanes on consecutive pieces read consecutive 16-byte words (a [piece][4] layout would put every load of a
// quarter-warp on four of the eight 16-byte bank groups)
JX_D double spline_eval(const double* __restrict__ c, int nseg, int s, double t) {
    const double2 c01 = *reinterpret_cast<const double2*>(c + 2 * s);
    const double2 c23 = *reinterpret_cast<const double2*>(c + 2 * nseg + 2 * s);
    return c01.x + t * (c01.y + t * (c23.x + t * c23.y));
}

// Phase D of the map kernel: G[kx] = sum_u hf[u, kx] sum_v conv_c[u, v] w_v cos(2 pi kx v / N) on the FP64
// tensor cores.  `xs` holds conv_c row-major with pitch PITCH (compile time; 0 = the runtime `pitch`), in
// shared memory (fast path) or in the CTA's global scratch (large-map path).  Work item = (kx tile, part of
// the u tiles); nsplit = 1 when there are at least as many warps as kx tiles (one item holds every u tile:
// the cosine fragments are loaded once).  Every item runs exactly NUT u-tiles so the DMMA loop carries no
// predicates: the last part starts early enough to end at the last tile and leaves the tiles an earlier
// part already covered out of the final fold.  The cosine fragments come from L2 (d.cfrag, fragment order)
// through a 4-deep register prefetch queue.  gpart_s: [nsplit][hp8], nsplit <= JX_D_MAXSPLIT.
template <int NUT, int PITCH>
JX_D void k3_phase_d(const jx_dev& d, const double* __restrict__ xs, int pitch, double* __restrict__ gpart_s, int warp,
                     int lane, int nwarps, int nsplit) {
    const int hp8 = d.hp8, ntile = hp8 >> 3, nks = hp8 >> 2;          // K = hp8 in steps of 4: nks is even
    const int ld = PITCH ? PITCH : pitch;
    const int frow = lane >> 2, fk = lane & 3;
    // k-permutation: step ks covers v = 8 (ks >> 1) + 2 (ks & 1) + {0, 1, 4, 5}; with a row pitch = 2 (mod 16)
    // doubles every 8-byte bank pair is hit by exactly two lanes: 2 wavefronts per 256-byte fragment load
    const int voff = (fk & 1) + 4 * (fk >> 1);
    for (int item = warp; item < nsplit * ntile; item += nwarps) {
        const int jt = item % ntile, part = item / ntile;
        const int ut_first = part * NUT;                                  // first tile this part is responsible for
        const int ut_lo = ut_first + NUT <= ntile ? ut_first : ntile - NUT;
        double acc[NUT][2];
#pragma unroll
        for (int i = 0; i < NUT; ++i) acc[i][0] = acc[i][1] = 0.0;
        const double* arow = xs + (size_t)(ut_lo * 8 + frow) * ld + voff;
        const double* bp = d.cfrag + (size_t)jt * nks * 32 + lane;     // B fragments, one coalesced load per k step
        double bq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) bq[q] = q < nks ? __ldg(bp + q * 32) : 0.0;
        auto kstep = [&](const double b, const double* ap) {
            double af[NUT];
#pragma unroll
            for (int i = 0; i < NUT; ++i) af[i] = ap[(size_t)i * 8 * ld];
#pragma unroll
            for (int i = 0; i < NUT; ++i) dmma884(acc[i][0], acc[i][1], af[i], b);
        };
        int ks0 = 0;
        for (; ks0 + 4 <= nks; ks0 += 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double b = bq[q];
                if (ks0 + 4 + q < nks) bq[q] = __ldg(bp + (ks0 + 4 + q) * 32);
                kstep(b, arow + 4 * ks0 + 8 * (q >> 1) + 2 * (q & 1));
            }
        }
        if (ks0 < nks) {                                                 // two steps left
            kstep(bq[0], arow + 4 * ks0);
            kstep(bq[1], arow + 4 * ks0 + 2);
        }
        // fold in hf[u, kx] and reduce over the 8 fragment rows
        double g0 = 0.0, g1 = 0.0;
        const int kc = jt * 8 + 2 * fk;
#pragma unroll
        for (int i = 0; i < NUT; ++i) {
            if (ut_lo + i < ut_first) continue;         // tile already covered by the previous part (warp-uniform)
            const double2 h = __ldg(reinterpret_cast<const double2*>(
                d.hf_pad + (size_t)((ut_lo + i) * 8 + frow) * hp8 + kc));
            g0 += acc[i][0] * h.x;
            g1 += acc[i][1] * h.y;
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            g0 += __shfl_xor_sync(0xffffffffu, g0, o);
            g1 += __shfl_xor_sync(0xffffffffu, g1, o);
        }
        if (lane < 4) {
            gpart_s[part * hp8 + kc] = g0;
            gpart_s[part * hp8 + kc + 1] = g1;
        }
    }
}

// Dispatch on the number of u-tiles per item.  After it (and a block barrier) G[kx] = sum over parts of gpart_s.
template <int PITCH>
JX_D int k3_run_phase_d(const jx_dev& d, const double* xs, int pitch, double* gpart_s, int warp, int lane, int nwarps) {
    const int ntile = d.hp8 >> 3;
    int nsplit = (ntile + 11) / 12;                    // at most 12 u-tiles (24 accumulators) per item
    if (nsplit == 1 && ntile > nwarps) nsplit = 2;     // more items than warps anyway: finer items balance better
    switch ((ntile + nsplit - 1) / nsplit) {
#define JX_D_CASE(n) case n: k3_phase_d<n, PITCH>(d, xs, pitch, gpart_s, warp, lane, nwarps, nsplit); break;
        JX_D_CASE(1) JX_D_CASE(2) JX_D_CASE(3) JX_D_CASE(4) JX_D_CASE(5) JX_D_CASE(6)
        JX_D_CASE(7) JX_D_CASE(8) JX_D_CASE(9) JX_D_CASE(10) JX_D_CASE(11)
        default: k3_phase_d<12, PITCH>(d, xs, pitch, gpart_s, warp, lane, nwarps, nsplit); break;
#undef JX_D_CASE
    }
    return nsplit;
}


