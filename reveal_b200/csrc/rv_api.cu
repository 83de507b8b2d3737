// rv_api.cu -- the C-ABI of libreveal_b200.so (see include/reveal_b200.h for the
// reference interface each entry point replaces).
#include "../../include/reveal_b200.h"
#include "rv_internal.h"
#include "rv_sweep.h"
#include <stdarg.h>
#include <stdlib.h>
#include <vector>

namespace rv {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list va;
    va_start(va, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, va);
    va_end(va);
}

int Arena::reserve(size_t bytes) {
    if (bytes <= cap) return RV_OK;
    release();
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return RV_ERR_NOMEM;
    }
    base = (unsigned char *)p;
    cap = bytes;
    off = 0;
    return RV_OK;
}
void Arena::release() {
    if (base) cudaFree(base);
    base = nullptr;
    cap = off = 0;
}

}  // namespace rv

using namespace rv;

struct rv_index {
    Stream st;
    bool own_stream = false;
    Arena arena;      // resident arrays + build workspace
    Arena sw;         // sweep tile scratch (grow-only)
    Arena res;        // last sweep result (grow-only)
    i64 n = 0;
    int nsamples = 0, rc = 0;
    std::vector<i64> nsep;
    unsigned char *dT = nullptr;
    int *dSA = nullptr, *dISA = nullptr, *dLCP = nullptr;
    unsigned short *dSO = nullptr;
    i64 *dNsep = nullptr;
    bool built = false;
    // last sweep
    int last_kind = 0;  // 1 pair, 2 multi
    i64 last_rec = 0, last_mem = 0;
    i64 *d_rows = nullptr, *d_members = nullptr;
    rv_times times;
    cudaEvent_t ev[6] = {0, 0, 0, 0, 0, 0};
    void *pool = nullptr;  // DevPool of rv_split.cu (children of the recursion)
    int device = 0;        // the GPU this handle lives on (the current device at rv_index_create)
};

extern "C" void rv_pool_destroy(void *pool);

static size_t pad256(size_t b) { return (b + 255) / 256 * 256; }

extern "C" {

const char *rv_last_error(void) { return g_err; }
const char *rv_version(void) {
#ifdef RV_EMU
    return "reveal_b200 0.1 (EMULATED kernels -- test build, not the product)";
#else
    return "reveal_b200 0.1 (sm_100a)";
#endif
}

int rv_device_count(int *count) {
    if (!count) return RV_ERR_ARG;
    RV_CUDA(cudaGetDeviceCount(count));
    return RV_OK;
}
int rv_set_device(int device) {
    RV_CUDA(cudaSetDevice(device));
    return RV_OK;
}

int rv_index_create(rv_index **out, void *stream) {
    if (!out) return RV_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    RV_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) {
        set_error("no CUDA device: libreveal_b200 has no CPU path");
        return RV_ERR_CUDA;
    }
    rv_index *h = new rv_index();
    memset(&h->times, 0, sizeof h->times);
    cudaGetDevice(&h->device);
    if (stream) {
        h->st.s = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&h->st.s, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
            delete h;
            return RV_ERR_CUDA;
        }
        h->own_stream = true;
    }
    {
        void *pp = nullptr;
        cudaError_t e = cudaMallocHost(&pp, 2048);
        if (e != cudaSuccess) {
            set_error("cudaMallocHost failed: %s", cudaGetErrorString(e));
            delete h;
            return RV_ERR_CUDA;
        }
        h->st.pinned = (u32 *)pp;
    }
    for (int i = 0; i < 6; i++) {
        cudaError_t e = cudaEventCreate(&h->ev[i]);
        if (e != cudaSuccess) {
            set_error("cudaEventCreate failed: %s", cudaGetErrorString(e));
            delete h;
            return RV_ERR_CUDA;
        }
    }
    *out = h;
    return RV_OK;
}

void rv_index_free(rv_index *h) {
    if (!h) return;
    cudaStreamSynchronize(h->st.s);
    rv_pool_destroy(h->pool);
    h->arena.release();
    h->sw.release();
    h->res.release();
    for (int i = 0; i < 6; i++)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    prof_collect(h->st);
    for (cudaEvent_t e : h->st.free_events) cudaEventDestroy(e);
    if (h->st.pinned) cudaFreeHost(h->st.pinned);
    if (h->own_stream) cudaStreamDestroy(h->st.s);
    delete h;
}

static int build_common(rv_index *h, const uint8_t *T, bool T_on_device, int64_t n, const int64_t *nsep, int32_t nsamples, int32_t rc,
                        const int32_t *hSA = nullptr, const int32_t *hLCP = nullptr) {
    if (!h || !T || n <= 0 || nsamples < 1 || (nsamples > 1 && !nsep)) {
        set_error(n <= 0 ? "No text to index." : "rv_build: bad argument");  // interface.c:177-180
        return RV_ERR_ARG;
    }
    if (n >= ((int64_t)1 << 30)) {
        set_error("rv_build: n=%lld not supported (limit 2^30-1 characters per index)", (long long)n);
        return RV_ERR_UNSUPPORTED;
    }
    if (rc && (nsamples < 2 || nsep[0] < 0 || nsep[0] >= n)) {
        // the reference would reverse-complement from T + nsep[0] = T - 1 (interface.c:170): undefined there, refused here
        set_error("rv_build: rc=1 needs two samples and a non-empty first sample");
        return RV_ERR_ARG;
    }
    RV_CUDA(cudaSetDevice(h->device));
    h->built = false;
    h->last_kind = 0;
    h->n = n;
    h->nsamples = nsamples;
    h->rc = rc ? 1 : 0;
    h->nsep.assign(nsep, nsep + (nsamples - 1));
    // The handle always works on its own copy of T: 256-byte aligned and zero padded past n, which the
    // word-wise suffix comparisons rely on (and rc rewrites the text in place, interface.c:168-172).
    size_t need = pad256((size_t)n + 64) + 3 * pad256((size_t)n * 4) + pad256((size_t)n * 2) + pad256((size_t)nsamples * 8) + sa_workspace_bytes(n) + 4096;
    RV_TRY(h->arena.reserve(need));
    h->arena.reset();
    unsigned char *textbuf = h->arena.take<unsigned char>((size_t)n + 64);
    h->dSA = h->arena.take<int>(n);
    h->dISA = h->arena.take<int>(n);
    h->dLCP = h->arena.take<int>(n);
    h->dSO = nsamples > 2 ? h->arena.take<unsigned short>(n) : nullptr;
    h->dNsep = h->arena.take<i64>(nsamples);
    if (!textbuf || !h->dSA || !h->dISA || !h->dLCP || !h->dNsep || (nsamples > 2 && !h->dSO)) {
        set_error("rv_build: arena too small");
        return RV_ERR_NOMEM;
    }
    Stream &st = h->st;
    st.launches_total += st.launches;
    st.launches = 0;
    PhaseTimes pt;
    RV_CUDA(cudaEventRecord(h->ev[0], st.s));
    RV_CUDA(cudaMemcpyAsync(textbuf, T, (size_t)n, T_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st.s));
    RV_CUDA(cudaMemsetAsync(textbuf + n, 0, 64, st.s));
    h->dT = textbuf;
    if (nsamples > 1) RV_CUDA(cudaMemcpyAsync(h->dNsep, h->nsep.data(), (size_t)(nsamples - 1) * 8, cudaMemcpyHostToDevice, st.s));
    RV_CUDA(cudaEventRecord(h->ev[1], st.s));
    if (rc) RV_TRY(revcomp_suffix(st, h->dT, h->nsep[0], n));
    if (nsamples > 2) RV_TRY(so_build(st, n, h->dNsep, nsamples, h->dSO));  // independent of the suffix array
    RV_CUDA(cudaEventRecord(h->ev[2], st.s));
    bool lcp_done = false;
    if (hSA) {  // suffix array (and maybe LCP) from a cache file (interface.c:224-231, 255-262)
        RV_CUDA(cudaMemcpyAsync(h->dSA, hSA, (size_t)n * 4, cudaMemcpyHostToDevice, st.s));
        u32 *d_bad = (u32 *)h->arena.take<u32>(64);
        if (!d_bad) { set_error("rv_build: arena too small"); return RV_ERR_NOMEM; }
        RV_CUDA(cudaMemsetAsync(d_bad, 0, 4, st.s));
        RV_TRY(isa_build(st, n, h->dSA, h->dISA, d_bad));
        u32 bad = 0;
        RV_CUDA(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st.s));
        RV_CUDA(cudaStreamSynchronize(st.s));
        if (bad) { set_error("the suffix array passed in is not a permutation of 0..n-1"); return RV_ERR_ARG; }
        if (hLCP) {
            RV_CUDA(cudaMemcpyAsync(h->dLCP, hLCP, (size_t)n * 4, cudaMemcpyHostToDevice, st.s));
            lcp_done = true;
        }
    } else {
        RV_TRY(sa_build(st, h->arena, h->dT, n, h->dSA, h->dISA, h->dLCP, &lcp_done, &pt));
    }
    RV_CUDA(cudaEventRecord(h->ev[3], st.s));
    if (!lcp_done) RV_TRY(lcp_build(st, h->dT, n, h->dSA, h->dISA, h->dLCP));
    RV_CUDA(cudaEventRecord(h->ev[4], st.s));
    RV_CUDA(cudaEventRecord(h->ev[5], st.s));
    RV_CUDA(cudaStreamSynchronize(st.s));
    RV_KCHECK();
    rv_times &t = h->times;
    RV_CUDA(cudaEventElapsedTime(&t.h2d_ms, h->ev[0], h->ev[1]));
    RV_CUDA(cudaEventElapsedTime(&t.pack_ms, h->ev[1], h->ev[2]));  // revcomp (+ SO fill)
    RV_CUDA(cudaEventElapsedTime(&t.sa_ms, h->ev[2], h->ev[3]));
    RV_CUDA(cudaEventElapsedTime(&t.lcp_ms, h->ev[3], h->ev[4]));
    RV_CUDA(cudaEventElapsedTime(&t.so_ms, h->ev[4], h->ev[5]));
    RV_CUDA(cudaEventElapsedTime(&t.total_ms, h->ev[0], h->ev[5]));
    t.sa_rounds = pt.sa_rounds;
    t.sa_sorted_items = pt.sa_sorted_items;
    t.launches = st.launches;
    h->built = true;
    return RV_OK;
}

int rv_build(rv_index *h, const uint8_t *T, int64_t n, const int64_t *nsep, int32_t nsamples, int32_t rc) {
    return build_common(h, T, false, n, nsep, nsamples, rc);
}
int rv_build_device(rv_index *h, const uint8_t *dT, int64_t n, const int64_t *nsep, int32_t nsamples, int32_t rc) {
    return build_common(h, dT, true, n, nsep, nsamples, rc);
}

int rv_build_cached(rv_index *h, const uint8_t *T, int64_t n, const int64_t *nsep, int32_t nsamples, int32_t rc, const int32_t *SA,
                    const int32_t *LCP) {
    if (!SA) { set_error("rv_build_cached: SA missing"); return RV_ERR_ARG; }
    return build_common(h, T, false, n, nsep, nsamples, rc, SA, LCP);
}

int rv_get_times(const rv_index *h, rv_times *out) {
    if (!h || !out) return RV_ERR_ARG;
    *out = h->times;
    return RV_OK;
}
int64_t rv_index_n(const rv_index *h) { return h ? h->n : 0; }

int rv_profile(rv_index *h, int32_t enable) {
    if (!h) return RV_ERR_ARG;
    RV_TRY(prof_collect(h->st));
    h->st.prof = enable != 0;
    for (int k = 0; k < 4; k++) {
        h->st.prof_ms[k] = 0;
        h->st.prof_launches[k] = 0;
        h->st.prof_bytes[k] = 0;
    }
    return RV_OK;
}
int rv_get_profile(rv_index *h, rv_kernel_profile *out) {
    if (!h || !out) return RV_ERR_ARG;
    RV_TRY(prof_collect(h->st));
    for (int k = 0; k < 4; k++) {
        out->ms[k] = h->st.prof_ms[k];
        out->launches[k] = h->st.prof_launches[k];
        out->bytes[k] = h->st.prof_bytes[k];
    }
    out->launches_total = h->st.launches_total + h->st.launches;
    return RV_OK;
}

static int need_built(const rv_index *h) {
    if (!h) { set_error("null index handle"); return RV_ERR_ARG; }
    if (!h->built) { set_error("index not constructed"); return RV_ERR_STATE; }
    RV_CUDA(cudaSetDevice(h->device));  // every entry point passes through here or build_common: stay on the handle's GPU
    return RV_OK;
}

static int fetch_i32(rv_index *h, const int *d, void *out, int32_t idx_bits, bool as_unsigned32) {
    RV_TRY(need_built(h));
    if (!out || (idx_bits != 32 && idx_bits != 64)) return RV_ERR_ARG;
    size_t n = (size_t)h->n;
    if (idx_bits == 32 || as_unsigned32) {
        RV_CUDA(cudaMemcpyAsync(out, d, n * 4, cudaMemcpyDeviceToHost, h->st.s));
        RV_CUDA(cudaStreamSynchronize(h->st.s));
        return RV_OK;
    }
    // widen on the host: the entries are the same numbers in wider integers
    int *tmp = (int *)malloc(n * 4);
    if (!tmp) { set_error("host malloc failed"); return RV_ERR_NOMEM; }
    cudaError_t e = cudaMemcpyAsync(tmp, d, n * 4, cudaMemcpyDeviceToHost, h->st.s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st.s);
    if (e != cudaSuccess) { free(tmp); set_error("D2H copy failed: %s", cudaGetErrorString(e)); return RV_ERR_CUDA; }
    int64_t *o = (int64_t *)out;
    for (size_t i = 0; i < n; i++) o[i] = tmp[i];
    free(tmp);
    return RV_OK;
}

int rv_get_sa(rv_index *h, void *out, int32_t idx_bits) { return fetch_i32(h, h ? h->dSA : nullptr, out, idx_bits, false); }
int rv_get_sai(rv_index *h, void *out, int32_t idx_bits) { return fetch_i32(h, h ? h->dISA : nullptr, out, idx_bits, false); }
int rv_get_lcp(rv_index *h, void *out, int32_t idx_bits) { return fetch_i32(h, h ? h->dLCP : nullptr, out, idx_bits, true); }
int rv_get_so(rv_index *h, uint16_t *out) {
    RV_TRY(need_built(h));
    if (!out) return RV_ERR_ARG;
    if (!h->dSO) { set_error("SO is only built for more than two samples (interface.c:265)"); return RV_ERR_STATE; }
    RV_CUDA(cudaMemcpyAsync(out, h->dSO, (size_t)h->n * 2, cudaMemcpyDeviceToHost, h->st.s));
    RV_CUDA(cudaStreamSynchronize(h->st.s));
    return RV_OK;
}
int rv_get_text(rv_index *h, uint8_t *out) {
    RV_TRY(need_built(h));
    if (!out) return RV_ERR_ARG;
    RV_CUDA(cudaMemcpyAsync(out, h->dT, (size_t)h->n, cudaMemcpyDeviceToHost, h->st.s));
    RV_CUDA(cudaStreamSynchronize(h->st.s));
    return RV_OK;
}
int rv_device_arrays(rv_index *h, const uint8_t **dT, const int32_t **dSA, const int32_t **dSAi, const int32_t **dLCP, const uint16_t **dSO) {
    RV_TRY(need_built(h));
    if (dT) *dT = h->dT;
    if (dSA) *dSA = h->dSA;
    if (dSAi) *dSAi = h->dISA;
    if (dLCP) *dLCP = h->dLCP;
    if (dSO) *dSO = h->dSO;
    return RV_OK;
}

// ---- sweeps ---------------------------------------------------------------------
static int run_pair(rv_index *h, const SweepArgs &a, int64_t *count) {
    if (!count) return RV_ERR_ARG;
    h->last_kind = 0;
    RV_TRY(h->sw.reserve(sweep_scratch_bytes(a.n)));
    if (h->res.cap == 0) RV_TRY(h->res.reserve((size_t)1 << 20));
    // speculative write into the buffer at hand; redone only if the count outgrows it
    i64 c = 0;
    const i64 cap = (i64)(h->res.cap / 24);
    RV_TRY(sweep_pair_count(h->st, a, h->sw.base, &c, (i64 *)h->res.base, cap));
    if (c > cap) {
        RV_TRY(h->res.reserve(pad256((size_t)c * 24 * 2)));
        RV_TRY(sweep_pair_write(h->st, a, h->sw.base, (i64 *)h->res.base, c));
    }
    h->d_rows = (i64 *)h->res.base;
    h->last_kind = 1;
    h->last_rec = c;
    h->last_mem = 0;
    *count = c;
    return RV_OK;
}

static int run_multi(rv_index *h, const SweepArgs &a, int64_t *nrec, int64_t *nmem) {
    if (!nrec || !nmem) return RV_ERR_ARG;
    h->last_kind = 0;
    if (a.main_nsamples > 2 && !a.SO) { set_error("multi sweep: SO missing for %d samples", a.main_nsamples); return RV_ERR_STATE; }
    RV_TRY(h->sw.reserve(sweep_scratch_bytes(a.n)));
    if (h->res.cap == 0) RV_TRY(h->res.reserve((size_t)1 << 20));
    // layout of the result buffer: first third header rows (24 B), the rest member rows (16 B)
    i64 r = 0, m = 0;
    size_t hdr_bytes = pad256(h->res.cap / 3);
    i64 hdr_cap = (i64)(hdr_bytes / 24), mem_cap = (i64)((h->res.cap - hdr_bytes) / 16);
    RV_TRY(sweep_multi_count(h->st, a, h->sw.base, &r, &m, (i64 *)h->res.base, hdr_cap, (i64 *)(h->res.base + hdr_bytes), mem_cap));
    if (r > hdr_cap || m > mem_cap) {
        size_t need_hdr = pad256((size_t)(r > 0 ? r : 1) * 24 * 2), need_mem = pad256((size_t)(m > 0 ? m : 1) * 16 * 2);
        size_t total = 3 * (need_hdr > need_mem / 2 ? need_hdr : need_mem / 2) + 1024;
        RV_TRY(h->res.reserve(total));
        hdr_bytes = pad256(h->res.cap / 3);
        hdr_cap = (i64)(hdr_bytes / 24);
        mem_cap = (i64)((h->res.cap - hdr_bytes) / 16);
        RV_TRY(sweep_multi_write(h->st, a, h->sw.base, (i64 *)h->res.base, hdr_cap, (i64 *)(h->res.base + hdr_bytes), mem_cap));
    }
    h->d_rows = (i64 *)h->res.base;
    h->d_members = (i64 *)(h->res.base + hdr_bytes);
    h->last_kind = 2;
    h->last_rec = r;
    h->last_mem = m;
    *nrec = r;
    *nmem = m;
    return RV_OK;
}

extern "C++" {
namespace rv {
struct MainView {
    Stream *st;
    unsigned char *T;
    int *SA, *ISA, *LCP;
    unsigned short *SO;
    i64 n, nsep0;
    int nsamples, rc;
    void **pool_slot;
};
int main_view(rv_index *h, MainView *out) {
    RV_TRY(need_built(h));
    out->st = &h->st;
    out->T = h->dT;
    out->SA = h->dSA;
    out->ISA = h->dISA;
    out->LCP = h->dLCP;
    out->SO = h->dSO;
    out->n = h->n;
    out->nsep0 = h->nsamples > 1 ? h->nsep[0] : -1;
    out->nsamples = h->nsamples;
    out->rc = h->rc;
    out->pool_slot = &h->pool;
    return RV_OK;
}
int sub_sweep_pair(rv_index *h, const SweepArgs &a, int64_t *count) { return run_pair(h, a, count); }
int sub_sweep_multi(rv_index *h, const SweepArgs &a, int64_t *nrec, int64_t *nmem) { return run_multi(h, a, nrec, nmem); }
}  // namespace rv
}  // extern "C++"

static SweepArgs root_args(const rv_index *h) {
    SweepArgs a;
    a.T = h->dT;
    a.SA = h->dSA;
    a.LCP = h->dLCP;
    a.SO = h->dSO;
    a.n = h->n;
    a.nT = h->n;  // construct sets nT = n (interface.c:195)
    a.nsep0 = h->nsamples > 1 ? h->nsep[0] : -1;
    a.rc = h->rc;
    a.flavour = 0;
    a.minl = 0;
    a.minn = 2;
    a.main_nsamples = h->nsamples;
    return a;
}

int rv_mums_pair_count(rv_index *h, int32_t minl, int32_t flavour, int64_t *count) {
    RV_TRY(need_built(h));
    if (h->nsamples < 2) { set_error("getmums needs two samples (nsep[0])"); return RV_ERR_STATE; }
    SweepArgs a = root_args(h);
    a.minl = minl;
    a.flavour = flavour ? 1 : 0;
    return run_pair(h, a, count);
}

int rv_mums_pair_fetch(rv_index *h, int64_t *rows, int64_t cap) {
    if (!h || h->last_kind != 1) { set_error("no pair sweep result to fetch"); return RV_ERR_STATE; }
    i64 k = h->last_rec < cap ? h->last_rec : cap;
    if (k > 0) {
        if (!rows) return RV_ERR_ARG;
        RV_CUDA(cudaMemcpyAsync(rows, h->d_rows, (size_t)k * 24, cudaMemcpyDeviceToHost, h->st.s));
    }
    RV_CUDA(cudaStreamSynchronize(h->st.s));
    return RV_OK;
}

int rv_mums_multi_count(rv_index *h, int32_t minl, int32_t minn, int64_t *nrec, int64_t *nmem) {
    RV_TRY(need_built(h));
    SweepArgs a = root_args(h);
    a.minl = minl;
    a.minn = minn;
    return run_multi(h, a, nrec, nmem);
}

int rv_mems_multi_count(rv_index *h, int32_t minl, int32_t minn, int64_t *nrec, int64_t *nmem) {
    RV_TRY(need_built(h));
    if (!nrec || !nmem) return RV_ERR_ARG;
    if (h->nsamples > 64) { set_error("getmultimems: more than 64 samples are not supported"); return RV_ERR_UNSUPPORTED; }
    if (h->nsamples < 2) { set_error("getmultimems needs at least two samples"); return RV_ERR_STATE; }
    SweepArgs a = root_args(h);
    a.minl = minl;
    a.minn = minn;
    h->last_kind = 0;
    RV_TRY(h->sw.reserve(mems_scratch_bytes(a.n)));
    i64 r = 0, m = 0;
    RV_TRY(sweep_mems_count(h->st, a, h->sw.base, &r, &m));
    size_t hdr_bytes = pad256((size_t)(r > 0 ? r : 1) * 24);
    RV_TRY(h->res.reserve(hdr_bytes + pad256((size_t)(m > 0 ? m : 1) * 16)));
    h->d_rows = (i64 *)h->res.base;
    h->d_members = (i64 *)(h->res.base + hdr_bytes);
    if (r > 0) RV_TRY(sweep_mems_write(h->st, a, h->sw.base, h->d_rows, r, h->d_members, m));
    h->last_kind = 2;
    h->last_rec = r;
    h->last_mem = m;
    *nrec = r;
    *nmem = m;
    return RV_OK;
}

int rv_mums_multi_fetch(rv_index *h, int64_t *hdr, int64_t hdr_cap, int64_t *members, int64_t mem_cap) {
    if (!h || h->last_kind != 2) { set_error("no multi sweep result to fetch"); return RV_ERR_STATE; }
    i64 r = h->last_rec < hdr_cap ? h->last_rec : hdr_cap;
    i64 m = h->last_mem < mem_cap ? h->last_mem : mem_cap;
    if (r > 0) {
        if (!hdr) return RV_ERR_ARG;
        RV_CUDA(cudaMemcpyAsync(hdr, h->d_rows, (size_t)r * 24, cudaMemcpyDeviceToHost, h->st.s));
    }
    if (m > 0) {
        if (!members) return RV_ERR_ARG;
        RV_CUDA(cudaMemcpyAsync(members, h->d_members, (size_t)m * 16, cudaMemcpyDeviceToHost, h->st.s));
    }
    RV_CUDA(cudaStreamSynchronize(h->st.s));
    return RV_OK;
}

int rv_result_device(rv_index *h, const int64_t **d_rows, int64_t *nrows, const int64_t **d_members, int64_t *nmembers) {
    if (!h || h->last_kind == 0) { set_error("no sweep result"); return RV_ERR_STATE; }
    if (d_rows) *d_rows = h->last_rec ? h->d_rows : nullptr;
    if (nrows) *nrows = h->last_rec;
    if (d_members) *d_members = (h->last_kind == 2 && h->last_mem) ? h->d_members : nullptr;
    if (nmembers) *nmembers = h->last_kind == 2 ? h->last_mem : 0;
    return RV_OK;
}

int rv_result_pack_device(rv_index *h, int64_t *d_dst, int64_t cap_rows) {
    if (!h || h->last_kind == 0 || !d_dst) { set_error("no sweep result"); return RV_ERR_STATE; }
    // the count travels through the pinned scratch word so that the copy below can stay asynchronous
    int64_t *cnt = (int64_t *)(h->st.pinned + 300);
    *cnt = h->last_rec;
    RV_CUDA(cudaMemcpyAsync(d_dst, cnt, 8, cudaMemcpyHostToDevice, h->st.s));
    i64 m = h->last_rec < cap_rows ? h->last_rec : cap_rows;
    if (m > 0) RV_CUDA(cudaMemcpyAsync(d_dst + 3, h->d_rows, (size_t)m * 24, cudaMemcpyDeviceToDevice, h->st.s));
    return RV_OK;
}

int rv_sweep_pair_device(rv_index *h, const uint8_t *dT, const int32_t *dSA, const int32_t *dLCP, int64_t n, int64_t nT, int64_t nsep0,
                         int32_t rc, int32_t flavour, int32_t minl, int64_t *count) {
    if (!h || !dT || !dSA || !dLCP || n < 0) return RV_ERR_ARG;
    SweepArgs a;
    a.T = dT; a.SA = dSA; a.LCP = dLCP; a.SO = nullptr;
    a.n = n; a.nT = nT; a.nsep0 = nsep0; a.rc = rc; a.flavour = flavour ? 1 : 0;
    a.minl = minl; a.minn = 2; a.main_nsamples = 2;
    return run_pair(h, a, count);
}

int rv_sweep_multi_device(rv_index *h, const uint8_t *dT, const int32_t *dSA, const int32_t *dLCP, const uint16_t *dSO, int64_t n,
                          int64_t nsep0, int32_t main_nsamples, int32_t minl, int32_t minn, int64_t *nrec, int64_t *nmem) {
    if (!h || !dT || !dSA || !dLCP || n < 0) return RV_ERR_ARG;
    SweepArgs a;
    a.T = dT; a.SA = dSA; a.LCP = dLCP; a.SO = dSO;
    a.n = n; a.nT = n; a.nsep0 = nsep0; a.rc = 0; a.flavour = 0;
    a.minl = minl; a.minn = minn; a.main_nsamples = main_nsamples;
    return run_multi(h, a, nrec, nmem);
}

}  // extern "C"
