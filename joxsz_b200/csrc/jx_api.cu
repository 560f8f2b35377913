// C ABI of libjoxsz_b200.so (see include/joxsz_b200.h): handle life cycle, the batched likelihood
// entry point, parity taps and measurement helpers.  No CPU fallback anywhere: every compute entry
// launches the CUDA kernels or fails with a status code.
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "jx_common.cuh"

static thread_local std::string g_create_error;

namespace {

int fail(jx_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

int cuda_fail(jx_handle* h, cudaError_t e, const char* where) {
    return fail(h, JX_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

#define JX_CUDA(h, call)                                         \
    do {                                                         \
        cudaError_t e_ = (call);                                 \
        if (e_ != cudaSuccess) return cuda_fail(h, e_, #call);   \
    } while (0)

template <class T>
int dev_alloc(jx_handle* h, T** out, size_t count) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc");
    if (h->nallocs >= (int)(sizeof(h->allocs) / sizeof(h->allocs[0]))) {
        cudaFree(p);
        return fail(h, JX_ERR_INVALID, "allocation table full");
    }
    h->allocs[h->nallocs++] = p;
    *out = static_cast<T*>(p);
    return JX_OK;
}

template <class T>
int upload(jx_handle* h, const T** out, const T* host, size_t count) {
    T* p = nullptr;
    int rc = dev_alloc(h, &p, count);
    if (rc) return rc;
    cudaError_t e = cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMemcpy H2D");
    *out = p;
    return JX_OK;
}

// [rows, cols] -> [rows, ld] zero padded
int upload_padded(jx_handle* h, const double** out, const double* host, int rows, int cols, int ld) {
    std::vector<double> tmp((size_t)rows * ld, 0.0);
    for (int r = 0; r < rows; ++r) memcpy(&tmp[(size_t)r * ld], host + (size_t)r * cols, sizeof(double) * cols);
    return upload(h, out, tmp.data(), tmp.size());
}

int check_ready(jx_handle* h, const double* theta, int W) {
    if (!h) return JX_ERR_INVALID;
    if (!theta || W < 0) return fail(h, JX_ERR_INVALID, "theta is NULL or W < 0");
    if (W > h->d.max_walkers) {
        char buf[128];
        snprintf(buf, sizeof buf, "W = %d exceeds max_walkers = %d", W, h->d.max_walkers);
        return fail(h, JX_ERR_CAPACITY, buf);
    }
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
    return JX_OK;
}

void flush_stage_events(jx_handle* h) {
    if (!h->pending) return;
    cudaEventSynchronize(h->ev[JX_NSTAGE]);
    cudaEventSynchronize(h->evx[1]);
    // events on the caller's stream: start, profiles, (unused), project, szmap, filter, tail; the X-ray kernel runs
    // on the side stream, concurrently with the projection GEMM, between evx[0] and evx[1]
    struct span { int stage; cudaEvent_t a, b; };
    const span spans[JX_NSTAGE] = {{JX_ST_PROFILES, h->ev[0], h->ev[1]}, {JX_ST_XRAY, h->evx[0], h->evx[1]},
                                   {JX_ST_PROJECT, h->ev[1], h->ev[3]},  {JX_ST_SZMAP, h->ev[3], h->ev[4]},
                                   {JX_ST_FILTER, h->ev[4], h->ev[5]},   {JX_ST_TAIL, h->ev[5], h->ev[6]}};
    for (const span& sp : spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            h->stage_ms[sp.stage] += ms;
            h->stage_launches[sp.stage] += 1;     // one kernel per stage
        }
    }
    h->pending = false;
}

}  // namespace

extern "C" const char* jx_build_info(void) {
    return "libjoxsz_b200 abi=" "7" " arch=sm_100a fp64 K1=profiles K2=dmma-project K3=fft256-szmap(smem)|fft512/1024-szmap(L2) K7=dmma-filter K4=xray K5=tail";
}

extern "C" const char* jx_last_error(const jx_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" void jx_destroy(jx_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (int i = 0; i < h->nallocs; ++i) cudaFree(h->allocs[i]);
    if (h->d.ws_convq) cudaFree(h->d.ws_convq);
    if (h->tap_scratch) cudaFree(h->tap_scratch);
    if (h->ev_ready) {
        for (auto& e : h->ev) cudaEventDestroy(e);
        for (auto& e : h->evx) cudaEventDestroy(e);
    }
    if (h->side) {
        cudaStreamDestroy(h->side);
        cudaEventDestroy(h->ev_fork);
        cudaEventDestroy(h->ev_join);
    }
    delete h;
}

extern "C" int jx_create(const jx_setup* s, jx_handle** out) {
    if (!s || !out) return fail(nullptr, JX_ERR_INVALID, "NULL setup or output pointer");
    *out = nullptr;
    if (s->abi_version != JX_ABI_VERSION) return fail(nullptr, JX_ERR_INVALID, "jx_setup.abi_version mismatch");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, JX_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU path)");
    if (s->device < 0 || s->device >= ndev) return fail(nullptr, JX_ERR_NO_DEVICE, "device ordinal out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, s->device) != cudaSuccess)
        return fail(nullptr, JX_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major < 10)
        return fail(nullptr, JX_ERR_NO_DEVICE, std::string("device is ") + prop.name + " (sm_" +
                    std::to_string(prop.major) + std::to_string(prop.minor) + "); sm_100 (B200) required");

    // ---- validate sizes
    auto bad = [&](const char* m) { return fail(nullptr, JX_ERR_INVALID, m); };
    if (s->ndim < 1 || s->ndim > JX_MAX_NDIM) return bad("ndim must be in 1..32");
    if (s->max_walkers < 1) return bad("max_walkers must be >= 1");
    if (s->nr < 4) return bad("nr must be >= 4");
    if (s->nmap < 3 || (s->nmap & 1) == 0) return bad("nmap must be odd (maps built by the reference have side 2m+1)");
    if (s->nh != s->nmap / 2 + 1) return bad("nh must equal nmap/2 + 1");
    if (s->nt != s->nh - 1) return bad("nt (sep) must equal nh - 1 (joxsz_funcs.py:469-473)");
    if (s->nt > s->nr) return bad("sep exceeds len(r_pp)");
    if (s->npad != 256 && s->npad != 512 && s->npad != 1024)
        return bad("cyclic length P must be 256 (shared-memory map kernel), 512 or 1024 (L2-staged map kernel)");
    if (s->nh > s->npad / 2 + 1) return bad("map quarter plane does not fit the cyclic length");
    if (s->nseg < 1 || s->nseg > 65535) return bad("nseg out of range");
    if (s->na < 1 || s->na > 32) return bad("1..32 annuli supported (one lane per annulus)");
    if (s->nb < 1 || s->ntab < 2 || s->nd < 1 || s->nconv < 2) return bad("empty X-ray / SZ data tables");
    for (int i = 0; i < JX_NPAR; ++i)
        if (s->slot_src[i] >= s->ndim) return bad("slot_src refers to a column beyond ndim");
    const void* required[] = {s->prior_kind, s->prior_a, s->prior_b, s->r_pp, s->proj_op, s->y_op, s->seg, s->dx,
                              s->bhat, s->cmat, s->hf, s->dinv, s->filt_q, s->w_t0, s->conv_T, s->conv_I, s->g_op,
                              s->flux, s->flux_err, s->midpt_kpc, s->projvols, s->tlog, s->lnrate0, s->lnrate1,
                              s->cts, s->srcscale, s->bkgterm};
    for (const void* p : required)
        if (!p) return bad("a required jx_setup array is NULL");
    if (s->calc_integ && (!s->w_integ || !(s->integ_sig > 0.0) || !std::isfinite(s->integ_mu)))
        return bad("calc_integ needs w_integ, a finite integ_mu and integ_sig > 0");
    if (!std::isfinite(s->prior_const)) return bad("prior_const is not finite (a frozen parameter is outside its prior)");
    for (int i = 0; i < s->nh * s->nh; ++i)
        if (s->seg[i] < 0 || s->seg[i] >= s->nseg) return bad("seg entry outside 0..nseg-1");

    JX_CUDA(nullptr, cudaSetDevice(s->device));
    jx_handle* h = new (std::nothrow) jx_handle();
    if (!h) return bad("out of host memory");
    h->device = s->device;
    h->sm_count = prop.multiProcessorCount;
    h->nallocs = 0;
    h->convq_capacity = 0;
    h->tap_scratch = nullptr;
    h->tap_scratch_walkers = 0;
    h->profiling = 0;
    h->ev_ready = false;
    h->pending = false;
    h->side = nullptr;
    h->collapsed_op = nullptr;
    memset(h->stage_ms, 0, sizeof h->stage_ms);
    memset(h->stage_launches, 0, sizeof h->stage_launches);
    jx_dev& d = h->d;
    memset(&d, 0, sizeof d);

    d.ndim = s->ndim; d.dens_mode = s->dens_mode; d.exclude_mass = s->exclude_unphy_mass;
    memcpy(d.slot_src, s->slot_src, sizeof d.slot_src);
    memcpy(d.slot_val, s->slot_val, sizeof d.slot_val);
    d.prior_const = s->prior_const;
    d.nr = s->nr; d.nrp = (s->nr + 7) & ~7; d.nt = s->nt; d.nmap = s->nmap; d.nh = s->nh; d.npad = s->npad;
    d.nq = s->npad / 2 + 1; d.nseg = s->nseg; d.ncoef = 4 * s->nseg;
    d.hp8 = (s->nh + 7) & ~7; d.hp16 = (s->nh + 15) & ~15;
    d.xs_pitch = ((d.nq > d.hp16 ? d.nq : d.hp16) + 3) & ~3;       // rows start on a 32-byte sector
    d.nconv = s->nconv; d.nd = s->nd; d.na = s->na; d.nb = s->nb; d.ntab = s->ntab;
    d.tmin = s->tmin; d.tmax = s->tmax; d.max_walkers = s->max_walkers;
    d.calc_integ = s->calc_integ ? 1 : 0; d.integ_mu = s->integ_mu; d.integ_sig = s->integ_sig;

    int rc = JX_OK;
#define UP(field, src, count) if (!rc) rc = upload(h, &d.field, src, (size_t)(count))
    const int H = s->nh, Q = d.nq;
    UP(prior_kind, s->prior_kind, s->ndim);
    UP(prior_a, s->prior_a, s->ndim);
    UP(prior_b, s->prior_b, s->ndim);
    UP(r_pp, s->r_pp, s->nr);
    if (!rc) rc = upload_padded(h, &d.proj_op_tap, s->proj_op, d.ncoef, s->nr, d.nrp);
    if (!rc) {   // production layout: two planes of 16-byte entries, row 2 g + p of plane 0 = coefficient p (0, 1) of
                 // piece g, row 2 nseg + 2 g + (p - 2) = coefficient p (2, 3): see spline_eval (k3_common.cuh)
        std::vector<double> il((size_t)d.ncoef * s->nr);
        for (int p = 0; p < 4; ++p)
            for (int g = 0; g < s->nseg; ++g)
                memcpy(&il[((size_t)(p >> 1) * 2 * s->nseg + 2 * g + (p & 1)) * s->nr],
                       s->proj_op + ((size_t)p * s->nseg + g) * s->nr, sizeof(double) * s->nr);
        rc = upload_padded(h, &d.proj_op, il.data(), d.ncoef, s->nr, d.nrp);
    }
    if (!rc) rc = upload_padded(h, &d.y_op, s->y_op, s->nr, s->nr, d.nrp);
    UP(seg, s->seg, H * H);
    UP(dx, s->dx, H * H);
    UP(bhat, s->bhat, Q * Q);
    d.nbeam = s->nbeam;
    if (!rc && s->bmix && s->nbeam >= 1 && s->nbeam <= JX_BMIX_ROWS) {
        d.bmix_pitch = s->npad == 256 ? JX_BMIX_PITCH : ((Q + 3) & ~3);
        std::vector<double> t((size_t)JX_BMIX_ROWS * d.bmix_pitch, 0.0);
        for (int j = 0; j < s->nbeam; ++j) memcpy(&t[(size_t)j * d.bmix_pitch], s->bmix + (size_t)j * Q, sizeof(double) * Q);
        rc = upload(h, &d.bmix, t.data(), t.size());
    }
    UP(hf, s->hf, H * H);
    UP(dinv, s->dinv, H * H);
    UP(filt_q, s->filt_q, H * H);
    UP(w_t0, s->w_t0, s->nt);
    UP(conv_T, s->conv_T, s->nconv);
    UP(conv_I, s->conv_I, s->nconv);
    UP(g_op, s->g_op, s->nd * H);
    UP(flux, s->flux, s->nd);
    UP(flux_err, s->flux_err, s->nd);
    UP(midpt_kpc, s->midpt_kpc, s->na);
    UP(projvols, s->projvols, s->na * s->na);
    UP(tlog, s->tlog, s->ntab);
    UP(lnrate0, s->lnrate0, s->nb * s->ntab);
    UP(lnrate1, s->lnrate1, s->nb * s->ntab);
    UP(cts, s->cts, s->nb * s->na);
    UP(srcscale, s->srcscale, s->nb * s->na);
    UP(bkgterm, s->bkgterm, s->nb * s->na);
#undef UP
    // derived tables
    if (!rc) {
        std::vector<double> t((size_t)H * H);
        for (int v = 0; v < H; ++v)
            for (int k = 0; k < H; ++k) t[(size_t)k * H + v] = s->cmat[(size_t)v * H + k];
        rc = upload(h, &d.cmat_t, t.data(), t.size());
    }
    if (!rc) {
        std::vector<uint16_t> s16((size_t)H * H);
        for (int i = 0; i < H * H; ++i) s16[i] = (uint16_t)s->seg[i];
        rc = upload(h, &d.seg16, s16.data(), s16.size());
    }
    if (!rc) {
        std::vector<double> c(s->nmap);
        for (int m = 0; m < s->nmap; ++m) c[m] = cos(2.0 * M_PI * (double)m / (double)s->nmap);
        rc = upload(h, &d.costab, c.data(), c.size());
    }
    if (!rc) {
        std::vector<double> t(s->nr, 0.0);
        if (s->w_integ) memcpy(t.data(), s->w_integ, sizeof(double) * s->nr);
        rc = upload(h, &d.w_integ, t.data(), t.size());
    }
    if (!rc) {
        std::vector<double> t(s->nr);
        for (int i = 0; i < s->nr; ++i) t[i] = log(s->r_pp[i]);
        rc = upload(h, &d.ln_r_pp, t.data(), t.size());
    }
    if (!rc) {
        std::vector<double> t(s->na);
        for (int i = 0; i < s->na; ++i) t[i] = log(s->midpt_kpc[i]);
        rc = upload(h, &d.ln_midpt, t.data(), t.size());
    }
    if (!rc) {
        std::vector<double> t((size_t)H * s->nd);
        for (int dp = 0; dp < s->nd; ++dp)
            for (int v = 0; v < H; ++v) t[(size_t)v * s->nd + dp] = s->g_op[(size_t)dp * H + v];
        rc = upload(h, &d.g_op_t, t.data(), t.size());
    }
    if (!rc) {   // synthesis table: pixels with u <= v, in thread order, padded to a multiple of the CTA size
        std::vector<jx_synth_px> t;
        for (int u = 0; u < H; ++u)
            for (int v = u; v < H; ++v) {
                jx_synth_px e;
                e.dx = s->dx[(size_t)u * H + v];
                e.seg = (uint16_t)s->seg[(size_t)u * H + v];
                e.u = (uint16_t)u; e.v = (uint16_t)v; e.pad = 0;
                t.push_back(e);
            }
        while (t.size() % 256) { jx_synth_px e; e.dx = 0.0; e.seg = 0; e.u = 0xffff; e.v = 0; e.pad = 0; t.push_back(e); }
        d.nsynth = (int)t.size();
        rc = upload(h, &d.synth, t.data(), t.size());
    }
    d.synth_tiles = nullptr; d.nsynth_tiles = 0;
    if (!rc && s->npad > 256) {   // the same pixels in 32 x 32 tiles (u block <= v block; diagonal tiles whole): the large-map
                                  // kernel stores a tile by rows and its mirror image through a shared-memory transpose
        std::vector<jx_synth_px> t;
        const int nt = (H + 31) / 32;
        for (int ub = 0; ub < nt; ++ub)
            for (int vb = ub; vb < nt; ++vb)
                for (int i = 0; i < 32; ++i)
                    for (int j = 0; j < 32; ++j) {
                        const int u = 32 * ub + i, v = 32 * vb + j;
                        jx_synth_px e;
                        if (u < H && v < H) {
                            e.dx = s->dx[(size_t)u * H + v];
                            e.seg = (uint16_t)s->seg[(size_t)u * H + v];
                            e.u = (uint16_t)u; e.v = (uint16_t)v; e.pad = 0;
                        } else {
                            e.dx = 0.0; e.seg = 0; e.u = 0xffff; e.v = 0; e.pad = 0;
                        }
                        t.push_back(e);
                    }
        d.nsynth_tiles = (int)(t.size() / 1024);
        rc = upload(h, &d.synth_tiles, t.data(), t.size());
    }
    if (!rc) {   // beam spectrum per column pair in the register order of the nine-thread column FFT:
                 // [cp][p][t] = bhat[fold(t + 16 rev16(p)), 2cp..2cp+1], t = 0..8
        const int ncp = (Q + 1) / 2, P = s->npad;
        std::vector<double> t((size_t)ncp * 16 * 9 * 2, 0.0);
        for (int cp = 0; cp < ncp; ++cp)
            for (int p = 0; p < 16; ++p)
                for (int th = 0; th < 9; ++th) {
                    const int n = th + 16 * ((p >> 2) + 4 * (p & 3));
                    const int f = n <= P / 2 ? n : P - n;
                    const size_t o = (((size_t)cp * 16 + p) * 9 + th) * 2;
                    if (f >= Q) continue;               // cyclic lengths other than 256 use the large-map kernel
                    t[o] = s->bhat[(size_t)f * Q + 2 * cp];
                    t[o + 1] = 2 * cp + 1 < Q ? s->bhat[(size_t)f * Q + 2 * cp + 1] : 0.0;
                }
        const double* dev = nullptr;
        rc = upload(h, &dev, t.data(), t.size());
        d.bhat_sw = reinterpret_cast<const double2*>(dev);
    }
    if (!rc) {
        // filter stage as one GEMM (k7_filter.cu): response of map_out[N//2, N//2 + x] to the convolved-map pixel
        // pair conv_c[u,v] = conv_c[v,u],
        //   R[(u,v), x] = sum_kx F[(u,v), kx] dinv[kx, x],   F = hf[u,kx] cmat[v,kx] + [u != v] hf[v,kx] cmat[u,kx],
        // stored as filt_op[x, (u,v)] with the output index padded to hpf rows and the pixel index to ktri columns
        d.ntri = H * (H + 1) / 2;
        d.ktri = (d.ntri + 31) & ~31;
        d.hpf = jx_filter_pitch(d.hp8);
        if (H <= 136) {
            // host, accumulated in long double so that the operator is correctly rounded to double
            std::vector<double> t((size_t)d.hpf * d.ktri, 0.0);
            std::vector<long double> f(H), dv((size_t)H * H);
            for (int i = 0; i < H * H; ++i) dv[i] = (long double)s->dinv[i];
            size_t idx = 0;
            for (int u = 0; u < H; ++u)
                for (int v = u; v < H; ++v, ++idx) {
                    for (int kx = 0; kx < H; ++kx) {
                        long double e = (long double)s->hf[(size_t)u * H + kx] * (long double)s->cmat[(size_t)v * H + kx];
                        if (u != v) e += (long double)s->hf[(size_t)v * H + kx] * (long double)s->cmat[(size_t)u * H + kx];
                        f[kx] = e;
                    }
                    for (int x = 0; x < H; ++x) {
                        long double acc = 0.0L;
                        for (int kx = 0; kx < H; ++kx) acc += f[kx] * dv[(size_t)kx * H + x];
                        t[(size_t)x * d.ktri + idx] = (double)acc;
                    }
                }
            rc = upload(h, &d.filt_op, t.data(), t.size());
        } else {
            // wide quarter planes (the 511-pixel maps: 32 896 pixels x 256 outputs): F on the host, the product with
            // dinv on the device by the DMMA GEMM, filt_op[x, :] = sum_kx dinv[kx, x] F[:, kx]
            double *fdev = nullptr, *ddev = nullptr, *op = nullptr;
            {
                std::vector<double> f((size_t)d.ktri * d.hp8, 0.0);
                size_t idx = 0;
                for (int u = 0; u < H; ++u)
                    for (int v = u; v < H; ++v, ++idx)
                        for (int kx = 0; kx < H; ++kx) {
                            double e = s->hf[(size_t)u * H + kx] * s->cmat[(size_t)v * H + kx];
                            if (u != v) e += s->hf[(size_t)v * H + kx] * s->cmat[(size_t)u * H + kx];
                            f[idx * d.hp8 + kx] = e;
                        }
                std::vector<double> dt((size_t)d.hpf * d.hp8, 0.0);
                for (int kx = 0; kx < H; ++kx)
                    for (int x = 0; x < H; ++x) dt[(size_t)x * d.hp8 + kx] = s->dinv[(size_t)kx * H + x];
                const double *fc = nullptr, *dc = nullptr;
                rc = upload(h, &fc, f.data(), f.size());
                if (!rc) rc = upload(h, &dc, dt.data(), dt.size());
                fdev = const_cast<double*>(fc); ddev = const_cast<double*>(dc);
            }
            if (!rc) rc = dev_alloc(h, &op, (size_t)d.hpf * d.ktri);
            if (!rc) {
                cudaError_t e = jx_gemm_configure();
                if (e == cudaSuccess) e = cudaMemset(op, 0, (size_t)d.hpf * d.ktri * sizeof(double));
                if (e == cudaSuccess) e = jx_launch_gemm_nt(ddev, d.hp8, fdev, d.hp8, op, d.ktri, d.hpf, d.ktri, d.hp8, 0);
                if (e == cudaSuccess) e = cudaDeviceSynchronize();
                if (e != cudaSuccess) rc = cuda_fail(h, e, "filter operator on the device");
            }
            d.filt_op = op;
        }
    }
    // workspace
    const size_t Wm = (size_t)s->max_walkers;
    if (!rc) rc = dev_alloc(h, &d.ws_pp, Wm * d.nrp);
    if (!rc) rc = dev_alloc(h, &d.ws_tsz, Wm * d.nt);
    if (!rc) rc = dev_alloc(h, &d.ws_ne, Wm * d.na);
    if (!rc) rc = dev_alloc(h, &d.ws_tx, Wm * d.na);
    if (!rc) rc = dev_alloc(h, &d.ws_prior, Wm);
    if (!rc) rc = dev_alloc(h, &d.ws_integ, Wm);
    if (!rc) rc = dev_alloc(h, &d.ws_xlike, Wm);
    if (!rc) rc = dev_alloc(h, &d.ws_flags, Wm);
    if (!rc) rc = dev_alloc(h, &d.ws_coef, Wm * d.ncoef);
    if (!rc) {
        rc = dev_alloc(h, &d.ws_tri, Wm * d.ktri);
        // the columns beyond ntri are never written by the map kernel and must not hold NaN patterns (they meet
        // zeros of filt_op); rows of walkers the map kernel skips are never read back
        if (!rc && cudaMemset(d.ws_tri, 0, Wm * d.ktri * sizeof(double)) != cudaSuccess)
            rc = fail(h, JX_ERR_CUDA, "cudaMemset(ws_tri)");
        if (!rc) rc = dev_alloc(h, &d.ws_rowp, (size_t)jx_filter_parts(d) * Wm * d.hpf);
    }
    if (!rc) {
        cudaError_t e = jx_profiles_configure(d);
        if (e == cudaSuccess) e = jx_gemm_configure();
        if (e != cudaSuccess) rc = cuda_fail(h, e, "configure profile / GEMM kernels");
    }
    if (!rc && d.npad == 256) {
        size_t smem = jx_szmap_smem_bytes(d);
        if (smem > (size_t)prop.sharedMemPerBlockOptin) {
            rc = fail(h, JX_ERR_INVALID, "map kernel needs more shared memory than the device offers");
        } else {
            d.k3_direct = jx_szmap_direct_ok(d) ? 1 : 0;
            d.k3_ws = d.k3_direct && jx_szmap_ws_ok(d) && jx_szmap_ws_smem_bytes(d) <= (size_t)prop.sharedMemPerBlockOptin ? 1 : 0;
            cudaError_t e = jx_szmap_configure(d);
            if (e == cudaSuccess && d.k3_ws) e = jx_szmap_ws_configure(d);
            if (e == cudaSuccess) e = jx_filter_configure(d);
            if (e != cudaSuccess) rc = cuda_fail(h, e, "configure map / filter kernels");
        }
    }
    if (!rc && d.npad != 256) {
        if (!jx_szmap_large_supported(d) || jx_szmap_large_smem_bytes(d) > (size_t)prop.sharedMemPerBlockOptin) {
            rc = fail(h, JX_ERR_INVALID, "large-map kernel: geometry does not fit the shared memory of an SM");
        } else {
            cudaError_t e = jx_szmap_large_configure(d);
            if (e == cudaSuccess) e = jx_filter_configure(d);
            d.k3l2 = jx_szmap_large2_ok(d) ? 1 : 0;
            if (e == cudaSuccess && d.k3l2) e = jx_szmap_large2_configure(d);
            if (e != cudaSuccess) rc = cuda_fail(h, e, "configure large-map kernel");
            const size_t nscr = (size_t)h->sm_count;                           // one scratch map pair per resident CTA
            if (!rc) rc = dev_alloc(h, &d.ws_scratch, nscr * d.hp8 * d.xs_pitch);
            if (!rc && d.bmix) rc = dev_alloc(h, &d.ws_scratch2, nscr * d.hp8 * d.xs_pitch);
        }
    }
    if (!rc) {
        cudaError_t e = cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "side stream");
    }
    if (rc) {
        g_create_error = h->err;
        jx_destroy(h);
        return rc;
    }
    *out = h;
    return JX_OK;
}

// Map stage + filter stage: spline coefficients -> map_out[N//2, N//2:] as `*nparts` partial rows at `*row`
// (leading dimension `*ld_row`) for the tail kernel.  Cyclic length 256: shared-memory map kernel writing the packed
// convolved map, then the filter GEMM over all walkers; 512 / 1024: L2-staged kernel, followed by the same GEMM.  `ev_mid`, when not NULL, is recorded between the two kernels (stage timers).
static cudaError_t launch_map_filter(jx_handle* h, const double* coef, const uint32_t* flags, int W, double* convq,
                                     const double** row, int* ld_row, int* nparts, cudaEvent_t ev_mid, cudaStream_t st) {
    const jx_dev& d = h->d;
    cudaError_t e;
    if (d.npad == 256) {
        e = d.k3_ws ? jx_launch_szmap_ws(d, coef, flags, W, h->sm_count, convq, d.ws_tri, st)
                    : jx_launch_szmap(d, coef, flags, W, h->sm_count, convq, d.ws_tri, st);
        if (e == cudaSuccess && ev_mid) e = cudaEventRecord(ev_mid, st);
        if (e != cudaSuccess) return e;
        *row = d.ws_rowp; *ld_row = d.hpf; *nparts = jx_filter_parts(d);
        return jx_launch_filter(d, d.ws_tri, W, d.ws_rowp, st);
    }
    e = d.k3l2 ? jx_launch_szmap_large2(d, coef, flags, W, h->sm_count, convq, d.ws_tri, d.ws_scratch, d.ws_scratch2, st)
               : jx_launch_szmap_large(d, coef, flags, W, h->sm_count, convq, d.ws_tri, d.ws_scratch, d.ws_scratch2, st);
    if (e == cudaSuccess && ev_mid) e = cudaEventRecord(ev_mid, st);
    if (e != cudaSuccess) return e;
    *row = d.ws_rowp; *ld_row = d.hpf; *nparts = jx_filter_parts(d);
    return jx_launch_filter(d, d.ws_tri, W, d.ws_rowp, st);
}

// ------------------------------------------------------------------------------------------------
// the hot path
// ------------------------------------------------------------------------------------------------
extern "C" int jx_loglike(jx_handle* h, const double* theta, int32_t W, double* ll, void* stream) {
    int rc = check_ready(h, theta, W);
    if (rc) return rc;
    if (!ll) return fail(h, JX_ERR_INVALID, "ll is NULL");
    if (W == 0) return JX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    jx_dev& d = h->d;
    const bool prof = h->profiling != 0;
    if (prof) {
        if (!h->ev_ready) {
            for (auto& e : h->ev) JX_CUDA(h, cudaEventCreate(&e));
            for (auto& e : h->evx) JX_CUDA(h, cudaEventCreate(&e));
            h->ev_ready = true;
        }
        flush_stage_events(h);
        JX_CUDA(h, cudaEventRecord(h->ev[0], st));
    }
    JX_CUDA(h, jx_launch_profiles(d, theta, W, d.ws_pp, d.nrp, d.ws_tsz, d.ws_ne, d.ws_tx, d.ws_flags, d.ws_prior,
                                  d.ws_integ, st));
    if (prof) JX_CUDA(h, cudaEventRecord(h->ev[1], st));
    // The X-ray kernel (latency bound, a few registers) only depends on the profiles: it runs on the handle's side
    // stream and shares the SMs with the projection GEMM.  It may set the "profile not > 0" bit of a walker while
    // the map kernel runs: the map kernel masks that bit out of its skip test (its decision only uses the bits the
    // profile kernel wrote, which are final), and the tail kernel, after the join, writes -inf for the walker.
    JX_CUDA(h, cudaEventRecord(h->ev_fork, st));
    JX_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    if (prof) JX_CUDA(h, cudaEventRecord(h->evx[0], h->side));
    JX_CUDA(h, jx_launch_xray(d, theta, d.ws_ne, d.ws_tx, W, nullptr, d.ws_xlike, d.ws_flags, h->side));
    if (prof) JX_CUDA(h, cudaEventRecord(h->evx[1], h->side));
    JX_CUDA(h, cudaEventRecord(h->ev_join, h->side));
    JX_CUDA(h, jx_launch_project(d, d.ws_pp, W, d.proj_op, d.ncoef, d.ws_coef, st));
    if (prof) JX_CUDA(h, cudaEventRecord(h->ev[3], st));
    const double* row = nullptr;
    int ld_row = 0, nparts = 1;
    JX_CUDA(h, launch_map_filter(h, d.ws_coef, d.ws_flags, W, nullptr, &row, &ld_row, &nparts, prof ? h->ev[4] : nullptr, st));
    if (prof) JX_CUDA(h, cudaEventRecord(h->ev[5], st));
    JX_CUDA(h, cudaStreamWaitEvent(st, h->ev_join, 0));
    JX_CUDA(h, jx_launch_tail(d, theta, row, ld_row, nparts, d.ws_tsz, d.ws_flags, d.ws_prior, d.ws_xlike, d.ws_integ, W,
                              nullptr, nullptr, nullptr, ll, nullptr, st));
    if (prof) {
        JX_CUDA(h, cudaEventRecord(h->ev[6], st));
        h->pending = true;
    }
    return JX_OK;
}

// ------------------------------------------------------------------------------------------------
// collapsed mode: the whole SZ chain between the pressure profile and the consumed row is linear (Abel projection,
// spline fit, map synthesis, beam convolution, filter: joxsz_funcs.py:457-467), so row = L pp with one constant
// operator L [nh, nr].  L is obtained by pushing the nr unit profiles through the STAGED kernels of this handle (same
// arithmetic, no second implementation) the first time the mode is used; afterwards a likelihood call is
// K1 -> one DMMA GEMM -> K5.  Offered for callers that only need `ll`; the staged path stays the reference-shaped one.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void collapse_unit_rows_kernel(double* pp, int n, int ld, int first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * ld) return;
    const int r = i / ld, c = i - r * ld;
    pp[i] = (c == first + r) ? 1.0 : 0.0;
}
// lop[x, first + r] = sum_p rowp[p][r][x]
__global__ void collapse_store_kernel(const double* rowp, int nparts, int W, int ld_row, int nh, double* lop, int ldl,
                                      int first, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * nh) return;
    const int r = i / nh, x = i - r * nh;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += rowp[((size_t)p * W + r) * ld_row + x];
    lop[(size_t)x * ldl + first + r] = s;
}
}  // namespace

static int build_collapsed(jx_handle* h, cudaStream_t st) {
    jx_dev& d = h->d;
    if (h->collapsed_op) return JX_OK;
    double* lop = nullptr;
    int rc = dev_alloc(h, &lop, (size_t)d.hp8 * d.nrp);
    if (rc) return rc;
    JX_CUDA(h, cudaMemsetAsync(lop, 0, sizeof(double) * (size_t)d.hp8 * d.nrp, st));
    const int chunk = d.max_walkers < d.nr ? d.max_walkers : d.nr;
    for (int first = 0; first < d.nr; first += chunk) {
        const int n = d.nr - first < chunk ? d.nr - first : chunk;
        collapse_unit_rows_kernel<<<(n * d.nrp + 255) / 256, 256, 0, st>>>(d.ws_pp, n, d.nrp, first);
        JX_CUDA(h, cudaGetLastError());
        JX_CUDA(h, jx_launch_project(d, d.ws_pp, n, d.proj_op, d.ncoef, d.ws_coef, st));
        const double* row = nullptr;
        int ld_row = 0, nparts = 1;
        JX_CUDA(h, launch_map_filter(h, d.ws_coef, nullptr, n, nullptr, &row, &ld_row, &nparts, nullptr, st));
        collapse_store_kernel<<<(n * d.nh + 255) / 256, 256, 0, st>>>(row, nparts, n, ld_row, d.nh, lop, d.nrp, first, n);
        JX_CUDA(h, cudaGetLastError());
    }
    JX_CUDA(h, cudaStreamSynchronize(st));
    h->collapsed_op = lop;
    return JX_OK;
}

extern "C" int jx_loglike_collapsed(jx_handle* h, const double* theta, int32_t W, double* ll, void* stream) {
    int rc = check_ready(h, theta, W);
    if (rc) return rc;
    if (!ll) return fail(h, JX_ERR_INVALID, "ll is NULL");
    if (W == 0) return JX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    jx_dev& d = h->d;
    if ((rc = build_collapsed(h, st))) return rc;
    JX_CUDA(h, jx_launch_profiles(d, theta, W, d.ws_pp, d.nrp, d.ws_tsz, d.ws_ne, d.ws_tx, d.ws_flags, d.ws_prior,
                                  d.ws_integ, st));
    JX_CUDA(h, cudaEventRecord(h->ev_fork, st));
    JX_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    JX_CUDA(h, jx_launch_xray(d, theta, d.ws_ne, d.ws_tx, W, nullptr, d.ws_xlike, d.ws_flags, h->side));
    JX_CUDA(h, cudaEventRecord(h->ev_join, h->side));
    // row [W, hpf] = pp [W, nrp] . L^T: the partial-row buffer of the staged path is free in this mode
    JX_CUDA(h, jx_launch_gemm_nt(d.ws_pp, d.nrp, h->collapsed_op, d.nrp, d.ws_rowp, d.hpf, W, d.hp8, d.nrp, st));
    JX_CUDA(h, cudaStreamWaitEvent(st, h->ev_join, 0));
    JX_CUDA(h, jx_launch_tail(d, theta, d.ws_rowp, d.hpf, 1, d.ws_tsz, d.ws_flags, d.ws_prior, d.ws_xlike, d.ws_integ, W,
                              nullptr, nullptr, nullptr, ll, nullptr, st));
    return JX_OK;
}

// ------------------------------------------------------------------------------------------------
// parity taps
// ------------------------------------------------------------------------------------------------
extern "C" int jx_profiles(jx_handle* h, const double* theta, int32_t W, double* pp, double* tsz, double* ne_ann,
                           double* tx_ann, uint32_t* flags, double* prior, void* stream) {
    int rc = check_ready(h, theta, W);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    jx_dev& d = h->d;
    // X-ray positivity is part of the status bits: run K1 into the workspace copies K4 needs
    JX_CUDA(h, jx_launch_profiles(d, theta, W, pp, d.nr, tsz, d.ws_ne, d.ws_tx, d.ws_flags, prior, nullptr, st));
    JX_CUDA(h, jx_launch_xray(d, theta, d.ws_ne, d.ws_tx, W, nullptr, nullptr, d.ws_flags, st));
    if (ne_ann) JX_CUDA(h, cudaMemcpyAsync(ne_ann, d.ws_ne, sizeof(double) * W * d.na, cudaMemcpyDeviceToDevice, st));
    if (tx_ann) JX_CUDA(h, cudaMemcpyAsync(tx_ann, d.ws_tx, sizeof(double) * W * d.na, cudaMemcpyDeviceToDevice, st));
    if (flags) JX_CUDA(h, cudaMemcpyAsync(flags, d.ws_flags, sizeof(uint32_t) * W, cudaMemcpyDeviceToDevice, st));
    return JX_OK;
}

extern "C" int jx_sz_project(jx_handle* h, const double* theta, int32_t W, double* y, double* coef, void* stream) {
    int rc = check_ready(h, theta, W);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    jx_dev& d = h->d;
    JX_CUDA(h, jx_launch_profiles(d, theta, W, d.ws_pp, d.nrp, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st));
    if (y) JX_CUDA(h, jx_launch_project(d, d.ws_pp, W, d.y_op, d.nr, y, st));
    if (coef) JX_CUDA(h, jx_launch_project(d, d.ws_pp, W, d.proj_op_tap, d.ncoef, coef, st));
    return JX_OK;
}

static int ensure_convq(jx_handle* h, int W) {
    if (h->convq_capacity >= (size_t)W) return JX_OK;
    if (h->d.ws_convq) cudaFree(h->d.ws_convq);
    h->d.ws_convq = nullptr;
    h->convq_capacity = 0;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sizeof(double) * (size_t)W * h->d.nh * h->d.nh);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc(convq tap)");
    h->d.ws_convq = (double*)p;
    h->convq_capacity = W;
    return JX_OK;
}

extern "C" int jx_sz_maps(jx_handle* h, const double* theta, int32_t W, double* y2d, double* conv2d, double* mapout,
                          void* stream) {
    int rc = check_ready(h, theta, W);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    jx_dev& d = h->d;
    JX_CUDA(h, jx_launch_profiles(d, theta, W, d.ws_pp, d.nrp, d.ws_tsz, nullptr, nullptr, nullptr, nullptr, d.ws_integ,
                                  st));
    JX_CUDA(h, jx_launch_project(d, d.ws_pp, W, d.proj_op, d.ncoef, d.ws_coef, st));
    if (y2d) JX_CUDA(h, jx_launch_tap_y2d(d, d.ws_coef, W, y2d, st));
    if (conv2d || mapout) {
        if ((rc = ensure_convq(h, W))) return rc;
        const double* row = nullptr;
        int ld_row = 0, nparts = 1;
        JX_CUDA(h, launch_map_filter(h, d.ws_coef, nullptr, W, d.ws_convq, &row, &ld_row, &nparts, nullptr, st));
        if (conv2d) JX_CUDA(h, jx_launch_tap_expand(d, d.ws_convq, W, conv2d, st));
        if (mapout) {
            if (h->tap_scratch_walkers < (size_t)W) {
                if (h->tap_scratch) cudaFree(h->tap_scratch);
                h->tap_scratch = nullptr;
                h->tap_scratch_walkers = 0;
                void* p = nullptr;
                JX_CUDA(h, cudaMalloc(&p, sizeof(double) * 2 * (size_t)W * d.nh * d.nh));
                h->tap_scratch = (double*)p;
                h->tap_scratch_walkers = W;
            }
            JX_CUDA(h, jx_launch_tap_mapout(d, d.ws_convq, W, mapout, h->tap_scratch, st));
        }
    }
    return JX_OK;
}

extern "C" int jx_sz_profile(jx_handle* h, const double* theta, int32_t W, double* row, double* bright,
                             double* model, double* chisq, double* cint, void* stream) {
    int rc = check_ready(h, theta, W);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    jx_dev& d = h->d;
    JX_CUDA(h, jx_launch_profiles(d, theta, W, d.ws_pp, d.nrp, d.ws_tsz, nullptr, nullptr, nullptr, nullptr, d.ws_integ,
                                  st));
    JX_CUDA(h, jx_launch_project(d, d.ws_pp, W, d.proj_op, d.ncoef, d.ws_coef, st));
    const double* rowp = nullptr;
    int ld_row = 0, nparts = 1;
    JX_CUDA(h, launch_map_filter(h, d.ws_coef, nullptr, W, nullptr, &rowp, &ld_row, &nparts, nullptr, st));
    JX_CUDA(h, jx_launch_tail(d, theta, rowp, ld_row, nparts, d.ws_tsz, nullptr, nullptr, nullptr, nullptr, W, bright,
                              model, chisq, nullptr, row, st));
    if (cint) JX_CUDA(h, cudaMemcpyAsync(cint, d.ws_integ, sizeof(double) * W, cudaMemcpyDeviceToDevice, st));
    return JX_OK;
}

extern "C" int jx_xray(jx_handle* h, const double* theta, int32_t W, double* pred, double* cash, void* stream) {
    int rc = check_ready(h, theta, W);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    jx_dev& d = h->d;
    JX_CUDA(h, jx_launch_profiles(d, theta, W, nullptr, d.nr, nullptr, d.ws_ne, d.ws_tx, nullptr, nullptr, nullptr, st));
    JX_CUDA(h, jx_launch_xray(d, theta, d.ws_ne, d.ws_tx, W, pred, cash, nullptr, st));
    return JX_OK;
}

extern "C" int jx_cash_from_profiles(jx_handle* h, const double* pred, int32_t W, double* cash, void* stream) {
    int rc = check_ready(h, pred, W);
    if (rc) return rc;
    if (!cash) return fail(h, JX_ERR_INVALID, "cash is NULL");
    JX_CUDA(h, jx_launch_cash(h->d, pred, W, cash, (cudaStream_t)stream));
    return JX_OK;
}

// ------------------------------------------------------------------------------------------------
// measurement helpers
// ------------------------------------------------------------------------------------------------
extern "C" int jx_set_profiling(jx_handle* h, int32_t on) {
    if (!h) return JX_ERR_INVALID;
    if (!on && h->pending) flush_stage_events(h);
    h->profiling = on;
    return JX_OK;
}

extern "C" int jx_stage_times(jx_handle* h, double* ms, int64_t* launches) {
    if (!h || !ms || !launches) return JX_ERR_INVALID;
    cudaSetDevice(h->device);
    flush_stage_events(h);
    for (int i = 0; i < JX_NSTAGE; ++i) {
        ms[i] = h->stage_ms[i];
        launches[i] = h->stage_launches[i];
        h->stage_ms[i] = 0.0;
        h->stage_launches[i] = 0;
    }
    return JX_OK;
}
