// Device helpers shared by the two map kernels (k3_szmap.cu: maps whose cyclic length is 256, whole pipeline in
// shared memory; k3l_szmap.cu: larger maps, working set in an L2-resident global scratch).
#pragma once

#include "jx_fft.cuh"

struct k3_args {
    jx_dev d;
    const double* coef;
    const uint32_t* flags;
    int W;
    double* convq;        // parity tap: quarter plane of the convolved map, [W][nh][nh]
    double* tri;          // packed u <= v triangle of the convolved map for the filter GEMM, [W][d.ktri]
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

// value of the Compton-y spline on piece `s` at offset `t` from its left knot, Horner evaluation.  Coefficient
// layout: two planes of 16-byte entries, (c0, c1) of piece s at c[2 s] and (c2, c3) at c[2 nseg + 2 s], so that
// lanes on consecutive pieces read consecutive 16-byte words (a [piece][4] layout would put every load of a
// quarter-warp on four of the eight 16-byte bank groups)
JX_D double spline_eval(const double* __restrict__ c, int nseg, int s, double t) {
    const double2 c01 = *reinterpret_cast<const double2*>(c + 2 * s);
    const double2 c23 = *reinterpret_cast<const double2*>(c + 2 * nseg + 2 * s);
    return c01.x + t * (c01.y + t * (c23.x + t * c23.y));
}
