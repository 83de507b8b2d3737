// rv_api.cu -- the C-ABI of libreveal_b200.so (see include/reveal_b200.h for the
// reference interface each entry point replaces).
#include "../../include/reveal_b200.h"
#ifdef RV_EMU  // peer blocks of the emulated build: POSIX shared memory
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <map>
#include <string>
#endif
#include "rv_internal.h"
#include "rv_sweep.h"
#include <stdarg.h>
#include <stdlib.h>
#include <mutex>
#include <vector>

namespace rv {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list va;
    va_start(va, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, va);
    va_end(va);
}

// Device slabs of released arenas and pinned host blocks wait here for the next user: a caller that creates one index object
// per alignment (the extension does) would otherwise pay cudaMalloc + cudaFree (both synchronise the device) and a
// cudaHostAlloc (page locking: milliseconds) per construct().
struct Cached { void *p; size_t bytes; int device; };
static std::mutex g_cache_mu;
static std::vector<Cached> g_dev_cache, g_host_cache, g_host_live;
static const size_t CACHE_SLOTS = 4;

static void *cache_take(std::vector<Cached> &c, size_t bytes, int device, size_t *got) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    int best = -1;
    for (int i = 0; i < (int)c.size(); i++)
        if (c[i].device == device && c[i].bytes >= bytes && (best < 0 || c[i].bytes < c[best].bytes)) best = i;
    if (best < 0) return nullptr;
    void *p = c[best].p;
    *got = c[best].bytes;
    c.erase(c.begin() + best);
    return p;
}
// returns the block that has to be released for real (the smallest one when the cache is full), or nullptr
static void *cache_put(std::vector<Cached> &c, void *p, size_t bytes, int device) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    c.push_back(Cached{p, bytes, device});
    if (c.size() <= CACHE_SLOTS) return nullptr;
    int worst = 0;
    for (int i = 1; i < (int)c.size(); i++)
        if (c[i].bytes < c[worst].bytes) worst = i;
    void *out = c[worst].p;
    c.erase(c.begin() + worst);
    return out;
}

int Arena::reserve(size_t bytes) {
    if (bytes <= cap) return RV_OK;
    release();
    int dev = 0;
    cudaGetDevice(&dev);
    size_t got = 0;
    if (void *c = cache_take(g_dev_cache, bytes, dev, &got)) {
        base = (unsigned char *)c;
        cap = got;
        off = 0;
        return RV_OK;
    }
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
    // callers synchronise the stream that used the slab first (rv_index_free; reserve() on a grown request runs between builds)
    if (base) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (void *drop = cache_put(g_dev_cache, base, cap, dev)) cudaFree(drop);
    }
    base = nullptr;
    cap = off = 0;
}

}  // namespace rv

using namespace rv;

// Header row (count, sequence number of this pack on the handle, 0) followed by the first m result rows.
__global__ void pack_rows_kernel(i64 *__restrict__ dst, const i64 *__restrict__ rows, i64 m, i64 count, i64 seq) {
    i64 w = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 3 * (m + 1)) return;
    dst[w] = w >= 3 ? rows[w - 3] : (w == 0 ? count : (w == 1 ? seq : 0));
}

struct rv_index {
    Stream st;
    bool own_stream = false;
    Arena arena;      // resident arrays + build workspace
    Arena sw;         // sweep tile scratch (grow-only)
    Arena res;        // last sweep result (grow-only)
    Arena chain;      // staging of rv_chain_batch (grow-only)
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
    i64 pack_seq = 0;      // number of rv_result_pack_device calls on this handle (header word 1 of a packed block)
    i64 *d_rows = nullptr, *d_members = nullptr;
    rv_times times;
    cudaEvent_t ev[6] = {0, 0, 0, 0, 0, 0};
    void *pool = nullptr;  // DevPool of rv_split.cu (children of the recursion)
    int device = 0;        // the GPU this handle lives on (the current device at rv_index_create)
};

extern "C" void rv_pool_destroy(void *pool);
extern "C" void rv_pool_trim(void);

// The CUDA objects of a handle that owns its stream -- stream, 2 KB of pinned memory, phase events, profile events -- and the
// alphabet of the last text it built outlive the handle in a small cache: a caller that creates one index object per alignment
// (the extension does) pays no cudaStreamCreate / cudaMallocHost / cudaFreeHost per construct(), and its first build starts
// speculatively with the previous text's code table instead of waiting for the byte histogram.
struct Shell {
    int device;
    cudaStream_t s;
    cudaStream_t side;
    cudaEvent_t ev_fork, ev_join;
    u32 *pinned;
    cudaEvent_t ev[6];
    std::vector<cudaEvent_t> free_events;
    AlphaCache alpha;
};
static std::vector<Shell> g_shells;
static const size_t SHELL_SLOTS = 4;

static void shell_destroy(Shell &sh) {
    for (int i = 0; i < 6; i++)
        if (sh.ev[i]) cudaEventDestroy(sh.ev[i]);
    for (cudaEvent_t e : sh.free_events) cudaEventDestroy(e);
    if (sh.pinned) cudaFreeHost(sh.pinned);
    if (sh.ev_fork) cudaEventDestroy(sh.ev_fork);
    if (sh.ev_join) cudaEventDestroy(sh.ev_join);
    if (sh.side) cudaStreamDestroy(sh.side);
    if (sh.s) cudaStreamDestroy(sh.s);
}


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

int rv_host_alloc(int64_t bytes, void **ptr) {
    if (bytes <= 0 || !ptr) return RV_ERR_ARG;
    size_t got = 0;
    if (void *c = cache_take(g_host_cache, (size_t)bytes, -1, &got)) {
        *ptr = c;
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_host_live.push_back(Cached{c, got, -1});
        return RV_OK;
    }
    void *p = nullptr;
#ifdef RV_EMU
    p = malloc((size_t)bytes);
    if (!p) { set_error("malloc(%lld) failed", (long long)bytes); return RV_ERR_NOMEM; }
#else
    cudaError_t e = cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaHostAlloc(%lld bytes) failed: %s", (long long)bytes, cudaGetErrorString(e));
        return RV_ERR_NOMEM;
    }
#endif
    *ptr = p;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_host_live.push_back(Cached{p, (size_t)bytes, -1});
    return RV_OK;
}
static void host_release(void *p) {
#ifdef RV_EMU
    free(p);
#else
    cudaFreeHost(p);
#endif
}
void rv_host_free(void *ptr) {
    if (!ptr) return;
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (size_t i = 0; i < g_host_live.size(); i++)
            if (g_host_live[i].p == ptr) {
                bytes = g_host_live[i].bytes;
                g_host_live.erase(g_host_live.begin() + (long)i);
                break;
            }
    }
    if (!bytes) return;  // not one of ours
    if (void *drop = cache_put(g_host_cache, ptr, bytes, -1)) host_release(drop);
}
int rv_trim(void) {
    std::vector<Cached> d, h;
    std::vector<Shell> shells;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        d.swap(g_dev_cache);
        h.swap(g_host_cache);
        shells.swap(g_shells);
    }
    for (Shell &sh : shells) shell_destroy(sh);
    rv_pool_trim();
    int cur = 0;
    cudaGetDevice(&cur);
    for (Cached &c : d) {
        cudaSetDevice(c.device);
        cudaFree(c.p);
    }
    cudaSetDevice(cur);
    for (Cached &c : h) host_release(c.p);
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
    if (!stream) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (size_t i = 0; i < g_shells.size(); i++)
            if (g_shells[i].device == h->device) {
                Shell &sh = g_shells[i];
                h->st.s = sh.s;
                h->st.side = sh.side;
                h->st.ev_fork = sh.ev_fork;
                h->st.ev_join = sh.ev_join;
                h->st.pinned = sh.pinned;
                h->st.alpha = sh.alpha;
                h->st.free_events.swap(sh.free_events);
                for (int k = 0; k < 6; k++) h->ev[k] = sh.ev[k];
                h->own_stream = true;
                g_shells.erase(g_shells.begin() + (long)i);
                *out = h;
                return RV_OK;
            }
    }
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
    if (h->st.side) cudaStreamSynchronize(h->st.side);
    rv_pool_destroy(h->pool);
    h->arena.release();
    h->sw.release();
    h->res.release();
    h->chain.release();
    prof_collect(h->st);
    Shell sh;
    sh.device = h->device;
    sh.s = h->own_stream ? h->st.s : (cudaStream_t)0;
    sh.side = h->st.side;
    sh.ev_fork = h->st.ev_fork;
    sh.ev_join = h->st.ev_join;
    sh.pinned = h->st.pinned;
    for (int i = 0; i < 6; i++) sh.ev[i] = h->ev[i];
    sh.free_events.swap(h->st.free_events);
    sh.alpha = h->st.alpha;
    bool kept = false;
    if (h->own_stream && cudaGetLastError() == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        if (g_shells.size() < SHELL_SLOTS) {
            g_shells.push_back(std::move(sh));
            kept = true;
        }
    }
    if (!kept) shell_destroy(sh);
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
    if (!lcp_done) RV_TRY(lcp_build(st, h->arena, h->dT, n, h->dSA, h->dISA, h->dLCP));
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
    h->st.prof_mask = enable > 1 ? (unsigned)enable >> 1 : 0xffffffffu;
    for (int k = 0; k < RV_PROF_SLOTS; k++) {
        h->st.prof_ms[k] = 0;
        h->st.prof_launches[k] = 0;
        h->st.prof_bytes[k] = 0;
    }
    return RV_OK;
}
int rv_get_profile(rv_index *h, rv_kernel_profile *out) {
    if (!h || !out) return RV_ERR_ARG;
    RV_TRY(prof_collect(h->st));
    for (int k = 0; k < RV_PROF_SLOTS; k++) {
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
int rv_put_text(rv_index *h, int64_t begin, const uint8_t *src, int64_t len) {
    RV_TRY(need_built(h));
    if (!src || begin < 0 || len < 0 || begin + len > h->n) { set_error("rv_put_text: range outside the text"); return RV_ERR_ARG; }
    if (len) RV_CUDA(cudaMemcpyAsync(h->dT + begin, src, (size_t)len, cudaMemcpyHostToDevice, h->st.s));
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
    const i64 *nsep_host;  // the separators (nsamples - 1 of them), host memory of the handle
    int nsep_count;
    int device;            // the GPU the handle lives on
    void **pool_slot;
};
int main_view(rv_index *h, MainView *out) {
    RV_TRY(need_built(h));
    RV_CUDA(cudaSetDevice(h->device));   // (a process may drive several GPUs: everything below works on the handle's)
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
    out->nsep_host = h->nsep.data();
    out->nsep_count = (int)h->nsep.size();
    out->device = h->device;
    out->pool_slot = &h->pool;
    return RV_OK;
}
struct ChainView {
    Stream *st;
    Arena *ws;
};
int chain_view(rv_index *h, ChainView *out) {  // rv_chain_batch works on any handle, built or not
    if (!h) { set_error("null index handle"); return RV_ERR_ARG; }
    RV_CUDA(cudaSetDevice(h->device));
    out->st = &h->st;
    out->ws = &h->chain;
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
    if (h->nsamples > 2 && (int)h->nsep.size() <= SW_NSEP_INLINE) {
        a.nsep_n = (int)h->nsep.size();
        for (int k = 0; k < a.nsep_n; k++) a.nsep_v[k] = h->nsep[(size_t)k];
    }
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
    if (!h || h->last_kind == 0 || !d_dst || cap_rows < 0) { set_error("no sweep result"); return RV_ERR_STATE; }
    RV_CUDA(cudaSetDevice(h->device));
    i64 m = h->last_rec < cap_rows ? h->last_rec : cap_rows;
    i64 words = 3 * (m + 1);
    h->pack_seq++;
    // one launch: header row + rows.  d_dst may be a peer block (rv_peer_open): plain stores over NVLink
    RV_LAUNCH(pack_rows_kernel, (unsigned)((words + 255) / 256), 256, 0, h->st.s, d_dst, (const i64 *)h->d_rows, m, (i64)h->last_rec,
              (i64)h->pack_seq);
    RV_KCHECK();
    h->st.launches++;
    h->st.launches_total++;
    return RV_OK;
}

// ---- peer blocks: memory of the collecting rank mapped by the other processes of the box ------------------
#ifdef RV_EMU
}  // extern "C"
namespace rv {
// The emulated build has no CUDA IPC: POSIX shared memory stands in, so that the host-side protocol of
// PeerGather (handle exchange, block offsets, reuse) runs between real processes in the CPU tests.
struct EmuPeer { std::string name; size_t bytes; bool owner; };
static std::map<void *, EmuPeer> g_emu_peers;
static int emu_map(const char *name, size_t bytes, bool create, void **out) {
    int fd = shm_open(name, create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) { set_error("shm_open(%s) failed", name); return RV_ERR_CUDA; }
    if (create && ftruncate(fd, (off_t)bytes) != 0) { close(fd); shm_unlink(name); set_error("ftruncate failed"); return RV_ERR_NOMEM; }
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { if (create) shm_unlink(name); set_error("mmap failed"); return RV_ERR_NOMEM; }
    g_emu_peers[p] = EmuPeer{name, bytes, create};
    *out = p;
    return RV_OK;
}
}  // namespace rv
extern "C" {
int rv_peer_alloc(int64_t bytes, void **d_ptr, uint8_t *handle) {
    if (bytes <= 0 || !d_ptr || !handle) return RV_ERR_ARG;
    static int serial = 0;
    memset(handle, 0, RV_PEER_HANDLE_BYTES);
    snprintf((char *)handle, 48, "/rvemu_%d_%d", (int)getpid(), serial++);
    memcpy(handle + 48, &bytes, 8);
    return emu_map((const char *)handle, (size_t)bytes, true, d_ptr);
}
int rv_peer_open(const uint8_t *handle, void **d_ptr) {
    if (!handle || !d_ptr) return RV_ERR_ARG;
    int64_t bytes = 0;
    memcpy(&bytes, handle + 48, 8);
    char name[49];
    memcpy(name, handle, 48);
    name[48] = 0;
    if (bytes <= 0 || name[0] != '/') { set_error("rv_peer_open: not a peer handle"); return RV_ERR_ARG; }
    return emu_map(name, (size_t)bytes, false, d_ptr);
}
static int emu_unmap(void *p, bool owner) {
    auto it = g_emu_peers.find(p);
    if (it == g_emu_peers.end() || it->second.owner != owner) { set_error("not a peer block of this process"); return RV_ERR_ARG; }
    munmap(p, it->second.bytes);
    if (owner) shm_unlink(it->second.name.c_str());
    g_emu_peers.erase(it);
    return RV_OK;
}
int rv_peer_close(void *d_ptr) { return emu_unmap(d_ptr, false); }
int rv_peer_free(void *d_ptr) { return emu_unmap(d_ptr, true); }
#else
int rv_peer_alloc(int64_t bytes, void **d_ptr, uint8_t *handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) <= RV_PEER_HANDLE_BYTES, "handle size");
    if (bytes <= 0 || !d_ptr || !handle) return RV_ERR_ARG;
    *d_ptr = nullptr;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);  // plain cudaMalloc: IPC handles cannot name pool allocations
    if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%lld bytes) failed: %s", (long long)bytes, cudaGetErrorString(e)); return RV_ERR_NOMEM; }
    cudaIpcMemHandle_t hd;
    e = cudaIpcGetMemHandle(&hd, p);
    if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
    if (e != cudaSuccess) { cudaGetLastError(); cudaFree(p); set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); return RV_ERR_CUDA; }
    memset(handle, 0, RV_PEER_HANDLE_BYTES);
    memcpy(handle, &hd, sizeof(hd));
    *d_ptr = p;
    return RV_OK;
}
int rv_peer_open(const uint8_t *handle, void **d_ptr) {
    if (!handle || !d_ptr) return RV_ERR_ARG;
    *d_ptr = nullptr;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, sizeof(hd));
    // maps the allocation into this process and enables peer access to its device (NVLink / NVSwitch) on first use
    cudaError_t e = cudaIpcOpenMemHandle(d_ptr, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); *d_ptr = nullptr; set_error("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e)); return RV_ERR_CUDA; }
    return RV_OK;
}
int rv_peer_close(void *d_ptr) {
    if (!d_ptr) return RV_ERR_ARG;
    RV_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return RV_OK;
}
int rv_peer_free(void *d_ptr) {
    if (!d_ptr) return RV_ERR_ARG;
    RV_CUDA(cudaFree(d_ptr));
    return RV_OK;
}
#endif
int rv_peer_read(const void *d_src, void *host_dst, int64_t bytes) {
    if (!d_src || !host_dst || bytes < 0) return RV_ERR_ARG;
    if (bytes) RV_CUDA(cudaMemcpy(host_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost));
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
