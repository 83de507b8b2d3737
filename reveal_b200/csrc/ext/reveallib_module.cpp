// reveallib_module.cpp -- the compiled CPython extension `reveallib` / `reveallib64`: the host side of the drop-in.
//
// Mirrors the reference's extension (reveallib/interface.c + the `aligner` of reveallib/reveal.c) symbol for symbol:
//   type `index`      interface.c:841-881     ctor kwargs sa, lcp, cache      interface.c:515-518
//   methods           interface.c:474-487     addsample addsequence construct align getmums getmultimums
//                                             getmultimems copy
//   getters           interface.c:731-785     n depth nsamples samples nodes leftnode rightnode nsep SA SAi SO LCP T
//   exception         interface.c:933-936     reveallib.error
// but holds no algorithm: every array operation goes through the C-ABI of libreveal_b200.so (include/reveal_b200.h),
// which is dlopen'ed when the module is imported.  There is no CPU path: without that library, or without a GPU,
// the calls raise.  Child indexes of the recursion are `index` objects too, as in the reference (reveal.c:1136-1207).
// Built twice: module name reveallib (default) and, with -DSA64, reveallib64 (reveal.h:7-13: same numbers in wider
// integers; the 32-bit total-length guard of addsequence is off).
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/reveal_b200.h"
#include "chain_dp.h"

#ifdef SA64
#define MODNAME "reveallib64"
#define MODINIT PyInit_reveallib64
#else
#define MODNAME "reveallib"
#define MODINIT PyInit_reveallib
#endif

// ---- the C-ABI, resolved at run time ---------------------------------------------------------------------------------
#define RV_API_LIST(X)                                                                                                  \
    X(rv_last_error) X(rv_version) X(rv_host_alloc) X(rv_host_free) X(rv_index_create) X(rv_index_free) X(rv_build) X(rv_build_cached) X(rv_get_times)   \
    X(rv_get_sa) X(rv_get_sai) X(rv_get_lcp) X(rv_get_so) X(rv_get_text) X(rv_put_text) X(rv_mums_pair_count) X(rv_mums_pair_fetch)    \
    X(rv_mums_multi_count) X(rv_mems_multi_count) X(rv_mums_multi_fetch) X(rv_sub_root) X(rv_sub_n) X(rv_sub_free) X(rv_sub_get)    \
    X(rv_sub_mums_pair) X(rv_sub_mums_multi) X(rv_sub_fetch) X(rv_sub_step) X(rv_sub_step_batch) X(rv_sub_step_batch_begin) X(rv_sub_step_batch_end) X(rv_sub_extract) X(rv_mums_tiny_batch)

struct Api {
#define X(name) decltype(&::name) name = nullptr;
    RV_API_LIST(X)
#undef X
    void *handle = nullptr;
    std::string path;
};
static Api g_api;
static PyObject *RevealError = nullptr;

static bool load_library(const char *path) {
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        PyErr_Format(PyExc_ImportError,
                     MODNAME ": cannot load %s (%s); build the CUDA library first "
                             "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.",
                     path, dlerror());
        return false;
    }
    Api a;
#define X(name)                                                                                    \
    a.name = (decltype(a.name))dlsym(h, #name);                                                    \
    if (!a.name) {                                                                                 \
        PyErr_Format(PyExc_ImportError, MODNAME ": %s lacks the symbol %s", path, #name);          \
        dlclose(h);                                                                                \
        return false;                                                                              \
    }
    RV_API_LIST(X)
#undef X
    a.handle = h;
    a.path = path;
    g_api = a;  // a previously loaded library stays mapped: live handles may still point into it
    return true;
}

static bool api_ready() {
    if (g_api.handle) return true;
    PyErr_SetString(PyExc_ImportError, MODNAME ": libreveal_b200.so is not loaded");
    return false;
}

// ---- host copy of the text: grows like the reference's realloc'ed T (interface.c:70-83), but in pinned memory from the
// library (rv_host_alloc) so that construct() is one DMA; plain malloc while the library is not loaded or refuses ----------
// A genome-sized sequence goes into the (pinned) text buffer on up to four threads: one core copies about 10 GB/s, and with
// the build itself under a millisecond the copies of addsequence() were a sixth of a step through the drop-in.
static void copy_big(char *dst, const char *src, size_t len) {
    static const size_t MIN_BIG = (size_t)1 << 21;
    static const unsigned cores = std::thread::hardware_concurrency();
    static const int forced = getenv("RV_TEXT_COPY_THREADS") ? atoi(getenv("RV_TEXT_COPY_THREADS")) : 0;  // 1..4 (measurements)
    // the processes of one box (one per GPU under torchrun) share its cores: a process takes its share, at most four threads
    static const int local_world = getenv("LOCAL_WORLD_SIZE") && atoi(getenv("LOCAL_WORLD_SIZE")) > 0 ? atoi(getenv("LOCAL_WORLD_SIZE")) : 1;
    static const unsigned share = cores / (2u * (unsigned)local_world);
    const size_t parts = len < MIN_BIG ? 1 : (forced >= 1 && forced <= 4 ? (size_t)forced : (share >= 4 ? 4 : (share >= 2 ? 2 : 1)));
    if (parts == 1) {
        memcpy(dst, src, len);
        return;
    }
    const size_t chunk = ((len / parts) + 4095) & ~(size_t)4095;
    std::thread helpers[3];
    size_t started = 0, done_upto = chunk < len ? chunk : len;
    try {
        for (size_t k = 1; k < parts && k * chunk < len; k++) {
            const size_t at = k * chunk, sz = at + chunk < len ? chunk : len - at;
            helpers[started] = std::thread([=] { memcpy(dst + at, src + at, sz); });
            started++;
            done_upto = at + sz;
        }
    } catch (...) {  // no thread to be had: the caller copies the rest itself
    }
    memcpy(dst, src, chunk < len ? chunk : len);
    for (size_t k = 0; k < started; k++) helpers[k].join();
    if (done_upto < len) memcpy(dst + done_upto, src + done_upto, len - done_upto);
}

struct HostText {
    char *p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned = false;
    HostText() {}
    HostText(const HostText &) = delete;
    HostText &operator=(const HostText &) = delete;
    ~HostText() { drop(); }
    void drop() {
        if (p) {
            if (pinned) { g_api.rv_host_free(p); last_cap() = cap; }
            else free(p);
        }
        p = nullptr;
        n = cap = 0;
    }
    // capacity of the pinned block released last: the library keeps released blocks for the next text, and a text of about the
    // same size as the previous one (one index per alignment job) then starts in a block that already holds all of it -- no
    // doubling copies while its sequences are appended
    static size_t &last_cap() { static size_t c = 0; return c; }
    bool reserve(size_t need) {
        if (need <= cap) return true;
        size_t c = cap ? cap * 2 : (size_t)1 << 12;
        while (c < need) c *= 2;
        static const bool hint = getenv("RV_TEXT_NO_HINT") == nullptr;
        if (hint && !p && need >= ((size_t)1 << 16) && last_cap() > c && last_cap() <= 4 * c) c = last_cap();
        char *q = nullptr;
        bool pin = false;
        void *vp = nullptr;
        if (c >= ((size_t)1 << 16) && g_api.rv_host_alloc && g_api.rv_host_alloc((int64_t)c, &vp) == 0) {
            q = (char *)vp;
            pin = true;
        } else {
            q = (char *)malloc(c);
        }
        if (!q) return false;
        if (n) memcpy(q, p, n);
        size_t keep = n;
        drop();
        p = q;
        n = keep;
        cap = c;
        pinned = pin;
        return true;
    }
    bool append(const char *src, size_t len) {
        if (!reserve(n + len + 1)) return false;
        copy_big(p + n, src, len);
        n += len;
        return true;
    }
    bool push_back(char c) { return append(&c, 1); }
    bool assign(const HostText &o) {
        n = 0;
        return append(o.p, o.n);
    }
    const char *data() const { return p ? p : ""; }
    size_t size() const { return n; }
};

// ---- the index type ----------------------------------------------------------------------------------------------------
struct Index {
    PyObject_HEAD
    // root state
    rv_index *h;
    HostText *T;                 // host copy of the text (root only)
    std::vector<int64_t> *nsep;
    int nsamples, rc, depth, cache, built, tdirty, extracted;
    int consumed;                // root: its own SA / LCP were given up at its first split inside align() (reveal.c:1279-1284)
    int64_t n, nT;
    std::string *safile, *lcpfile;
    // recursion state
    rv_sub *sub;                 // device view (children; the root gets one during align)
    Index *mainidx;              // borrowed-with-reference: the root this child belongs to (NULL for a root)
    PyObject *samples, *nodes, *left_node, *right_node, *skipmums;
    PyObject *shard_units;       // root, after a sharded align(): [(owner rank, nodes of the unit), ...]
};

static PyTypeObject IndexType = {PyVarObject_HEAD_INIT(nullptr, 0)};

static int fail_native(int status) {
    if (status == 0) return 0;
    PyErr_SetString(RevealError, g_api.rv_last_error ? g_api.rv_last_error() : "native call failed");
    return -1;
}

// Result lists hold hundreds of thousands of tuples of integers.  None of them can be part of a reference cycle, so they are taken
// out of the cyclic collector's lists as they are made (what the collector itself does with such tuples when it meets them,
// Objects/tupleobject.c _PyTuple_MaybeUntrack), and no collection is started while a list is being filled: more than half of a
// getmums() call was the collector traversing the young tuples over and over.
struct GcPause {
    int was;
    GcPause() : was(PyGC_Disable()) {}
    ~GcPause() { if (was) PyGC_Enable(); }
};
static inline PyObject *atomic_tuple(PyObject *t) {
    if (t) PyObject_GC_UnTrack(t);
    return t;
}

static Index *root_of(Index *self) { return self->mainidx ? self->mainidx : self; }

static int ensure_handle(Index *self) {
    if (!api_ready()) return -1;
    if (self->h) return 0;
    return fail_native(g_api.rv_index_create(&self->h, nullptr));
}

static PyObject *index_new(PyTypeObject *type, PyObject *, PyObject *) {
    Index *self = (Index *)type->tp_alloc(type, 0);
    if (!self) return nullptr;
    self->h = nullptr;
    self->T = new HostText();
    self->nsep = new std::vector<int64_t>();
    self->safile = new std::string();
    self->lcpfile = new std::string();
    self->nsamples = self->rc = self->depth = self->cache = self->built = self->tdirty = self->extracted = self->consumed = 0;
    self->n = self->nT = 0;
    self->sub = nullptr;
    self->mainidx = nullptr;
    self->samples = PyList_New(0);
    self->nodes = PySet_New(nullptr);
    self->skipmums = PyList_New(0);
    self->shard_units = PyList_New(0);
    Py_INCREF(Py_None);
    self->left_node = Py_None;
    Py_INCREF(Py_None);
    self->right_node = Py_None;
    return (PyObject *)self;
}

static int index_init(Index *self, PyObject *args, PyObject *kwds) {
    static const char *kwlist[] = {"sa", "lcp", "cache", nullptr};  // interface.c:515-518
    const char *sa = "", *lcp = "";
    int cache = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kwds, "|ssi", (char **)kwlist, &sa, &lcp, &cache)) return -1;
    *self->safile = sa;
    *self->lcpfile = lcp;
    self->cache = cache;
    return 0;
}

static void index_dealloc(Index *self) {
    if (self->sub && g_api.rv_sub_free) g_api.rv_sub_free(self->sub);
    if (self->h && g_api.rv_index_free) g_api.rv_index_free(self->h);
    delete self->T;
    delete self->nsep;
    delete self->safile;
    delete self->lcpfile;
    Py_XDECREF(self->samples);
    Py_XDECREF(self->nodes);
    Py_XDECREF(self->left_node);
    Py_XDECREF(self->right_node);
    Py_XDECREF(self->skipmums);
    Py_XDECREF(self->shard_units);
    Py_XDECREF((PyObject *)self->mainidx);
    Py_TYPE(self)->tp_free((PyObject *)self);
}

// ---- text assembly (interface.c:18-95) ------------------------------------------------------------------------------------
static PyObject *index_addsample(Index *self, PyObject *args) {
    PyObject *sample = PyTuple_Size(args) > 0 ? PyTuple_GetItem(args, 0) : nullptr;
    if (!sample) {
        PyErr_SetString(RevealError, "Specify name of sample as argument.");
        return nullptr;
    }
    if (!PyUnicode_Check(sample)) {
        PyErr_SetString(RevealError, "Sample name has to be a string.");
        return nullptr;
    }
    PyList_Append(self->samples, sample);
    if (self->nsamples > 0) self->nsep->push_back(self->n - 1);  // position of the last '$' of the previous sample
    self->nsamples++;
    Py_RETURN_NONE;
}

static PyObject *index_addsequence(Index *self, PyObject *args) {
    const char *seq;
    Py_ssize_t l;
    if (!PyArg_ParseTuple(args, "s#", &seq, &l)) return nullptr;
    // The reference refuses texts of 2^31 characters and more in its 32-bit build and sends the user to --64 (interface.c:61-68).
    // The device arrays of this build hold 30-bit positions in BOTH modules (reveallib64 widens the numbers on the way out,
    // reveal.h:7-13): the limit is 2^30 - 1 characters per index, said here -- where the text is assembled -- rather than after
    // gigabytes have been appended.
    if ((uint64_t)self->n + (uint64_t)(l + 1) + 1 > ((uint64_t)1 << 30)) {
        PyErr_SetString(RevealError, "Total amount of sequence too large: this build indexes at most 2^30 - 1 characters per index "
                                     "(32-bit positions on the device, also behind \"--64\" / reveallib64).");
        return nullptr;
    }
    int64_t s = self->n;
    if (!self->T->append(seq, (size_t)l) || !self->T->push_back('$')) return PyErr_NoMemory();
    self->n += l + 1;
    PyObject *intv = Py_BuildValue("(L,L)", (long long)s, (long long)(self->n - 1));
    if (!intv) return nullptr;
    PySet_Add(self->nodes, intv);
    return intv;
}

// ---- construct (interface.c:160-291) ----------------------------------------------------------------------------------------
static bool read_ints(const std::string &path, int64_t n, bool wide, std::vector<int32_t> &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    out.resize((size_t)n);
    bool ok = true;
    if (wide) {
        std::vector<int64_t> tmp((size_t)n);
        ok = fread(tmp.data(), 8, (size_t)n, f) == (size_t)n;
        for (int64_t i = 0; ok && i < n; i++) out[(size_t)i] = (int32_t)tmp[(size_t)i];
    } else {
        ok = fread(out.data(), 4, (size_t)n, f) == (size_t)n;
    }
    fclose(f);
    return ok;
}

static PyObject *index_construct(Index *self, PyObject *args, PyObject *kwds) {
    static const char *kwlist[] = {"rc", nullptr};
    int rc = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kwds, "|i", (char **)kwlist, &rc)) return nullptr;
    rc = rc == 1 ? 1 : 0;
    if (self->mainidx) {
        PyErr_SetString(RevealError, "construct() on a child index");
        return nullptr;
    }
    if (self->n == 0) {
        PyErr_SetString(RevealError, "No text to index.");  // interface.c:177-180
        return nullptr;
    }
    if (rc && self->nsamples < 2) {
        PyErr_SetString(RevealError, "rc=1 needs a second sample.");
        return nullptr;
    }
    if (ensure_handle(self) != 0) return nullptr;
    const int64_t *nsep = self->nsep->empty() ? nullptr : self->nsep->data();
    int status;
    if (!self->safile->empty()) {  // precomputed arrays (interface.c:224-231, 255-262): raw saidx_t / lcp_t
#ifdef SA64
        const bool wide = true;
#else
        const bool wide = false;
#endif
        std::vector<int32_t> sa, lcp;
        if (!read_ints(*self->safile, self->n, wide, sa)) {
            PyErr_Format(RevealError, "cannot read %lld suffix array entries from %s", (long long)self->n, self->safile->c_str());
            return nullptr;
        }
        bool have_lcp = !self->lcpfile->empty();
        if (have_lcp && !read_ints(*self->lcpfile, self->n, false, lcp)) {  // lcp_t is 32 bits in both builds
            PyErr_Format(RevealError, "cannot read %lld lcp entries from %s", (long long)self->n, self->lcpfile->c_str());
            return nullptr;
        }
        Py_BEGIN_ALLOW_THREADS;
        status = g_api.rv_build_cached(self->h, (const uint8_t *)self->T->data(), self->n, nsep, self->nsamples, rc, sa.data(),
                                       have_lcp ? lcp.data() : nullptr);
        Py_END_ALLOW_THREADS;
    } else if (!self->lcpfile->empty()) {
        PyErr_SetString(RevealError, "an lcp file needs its suffix array file (sa=...) as well");
        return nullptr;
    } else {
        Py_BEGIN_ALLOW_THREADS;
        status = g_api.rv_build(self->h, (const uint8_t *)self->T->data(), self->n, nsep, self->nsamples, rc);
        Py_END_ALLOW_THREADS;
    }
    if (fail_native(status) != 0) return nullptr;
    self->rc = rc;
    self->nT = self->n;
    self->built = 1;
    self->consumed = 0;
    self->tdirty = rc;  // the reference reverse-complements its T in place (interface.c:168-172)
    self->depth = 0;
    if (self->cache == 1) {  // interface.c:182-189, 273-285
        std::vector<int32_t> buf((size_t)self->n);
        FILE *f = fopen(".reveal.t", "wb");
        if (f) { fwrite(self->T->data(), 1, (size_t)self->n, f); fclose(f); }
        if (g_api.rv_get_sa(self->h, buf.data(), 32) == 0 && (f = fopen(".reveal.sa", "wb"))) {
#ifdef SA64
            for (int64_t i = 0; i < self->n; i++) { int64_t v = buf[(size_t)i]; fwrite(&v, 8, 1, f); }
#else
            fwrite(buf.data(), 4, (size_t)self->n, f);
#endif
            fclose(f);
        }
        if (g_api.rv_get_lcp(self->h, buf.data(), 32) == 0 && (f = fopen(".reveal.lcp", "wb"))) {
            fwrite(buf.data(), 4, (size_t)self->n, f);
            fclose(f);
        }
    }
    Py_RETURN_NONE;
}

static int need_built(Index *self, PyObject *exc, const char *msg) {
    Index *r = root_of(self);
    if (r->built) return 0;
    PyErr_SetString(exc, msg);
    return -1;
}
// for what reads the root's own SA / LCP: gone once align() has split the root (reveal.c:1279-1284; the reference would follow a
// NULL pointer there)
static int need_arrays(Index *self) {
    if (need_built(self, RevealError, "Index not yet constructed.") != 0) return -1;
    if (!self->mainidx && self->consumed) {
        PyErr_SetString(RevealError, "Index not yet constructed (its suffix array was given up by align()).");
        return -1;
    }
    return 0;
}

// ---- sweeps ------------------------------------------------------------------------------------------------------------------
static PyObject *multi_to_list(const std::vector<int64_t> &hdr, const std::vector<int64_t> &mem, int64_t nrec, int64_t nmem, bool counts_are_sizes) {
    GcPause gc_pause;
    PyObject *lst = PyList_New((Py_ssize_t)nrec);
    if (!lst) return nullptr;
    for (int64_t k = 0; k < nrec; k++) {
        int64_t l = hdr[3 * k], cnt = hdr[3 * k + 1], first = hdr[3 * k + 2];
        int64_t end = counts_are_sizes ? first + cnt : (k + 1 < nrec ? hdr[3 * (k + 1) + 2] : nmem);
        PyObject *members = atomic_tuple(PyTuple_New((Py_ssize_t)(end - first)));
        for (int64_t x = first; x < end; x++) {
            PyObject *sp = atomic_tuple(PyTuple_New(2));
            PyTuple_SET_ITEM(sp, 0, PyLong_FromLong((long)mem[2 * x]));
            PyTuple_SET_ITEM(sp, 1, PyLong_FromLongLong((long long)mem[2 * x + 1]));
            PyTuple_SET_ITEM(members, (Py_ssize_t)(x - first), sp);
        }
        PyObject *rec = atomic_tuple(PyTuple_New(3));  // (l, n, ((sample, position), ...)): reveal.c:497 / :353
        PyTuple_SET_ITEM(rec, 0, PyLong_FromLongLong((long long)l));
        PyTuple_SET_ITEM(rec, 1, PyLong_FromLong((long)cnt));
        PyTuple_SET_ITEM(rec, 2, members);
        PyList_SET_ITEM(lst, (Py_ssize_t)k, rec);
    }
    return lst;
}

static PyObject *index_getmums(Index *self, PyObject *args) {
    int minl = 0;
    if (!PyArg_ParseTuple(args, "i", &minl)) return nullptr;
    if (self->mainidx || need_arrays(self) != 0) {
        if (self->mainidx) PyErr_SetString(RevealError, "getmums() on a child index");
        return nullptr;
    }
    int64_t k = 0;
    int status;
    std::vector<int64_t> rows;
    if (self->extracted) {  // after extract() the index is its own (SA, LCP) pair over the main text
        if (self->rc) { PyErr_SetString(RevealError, "getmums() after extract() on a reverse-complement index is not supported"); return nullptr; }
        Py_BEGIN_ALLOW_THREADS;
        status = g_api.rv_sub_mums_pair(self->sub, minl, &k);
        Py_END_ALLOW_THREADS;
        if (fail_native(status) != 0) return nullptr;
        rows.resize((size_t)(3 * k + 3));
        if (fail_native(g_api.rv_sub_fetch(self->sub, rows.data(), k, nullptr, 0)) != 0) return nullptr;
    } else {
        Py_BEGIN_ALLOW_THREADS;
        status = g_api.rv_mums_pair_count(self->h, minl, 0, &k);
        Py_END_ALLOW_THREADS;
        if (fail_native(status) != 0) return nullptr;
        rows.resize((size_t)(3 * k + 3));
        if (fail_native(g_api.rv_mums_pair_fetch(self->h, rows.data(), k)) != 0) return nullptr;
    }
    GcPause gc_pause;
    PyObject *lst = PyList_New((Py_ssize_t)k);
    if (!lst) return nullptr;
    PyObject *rcobj = PyLong_FromLong(self->rc);
    for (int64_t i = 0; i < k; i++) {  // (l, (a, b), rc): reveal.c:102-106 -- built without format parsing, this list is the bulk of the call
        PyObject *ab = atomic_tuple(PyTuple_New(2)), *rec = atomic_tuple(PyTuple_New(3));
        PyTuple_SET_ITEM(ab, 0, PyLong_FromLongLong((long long)rows[3 * i + 1]));
        PyTuple_SET_ITEM(ab, 1, PyLong_FromLongLong((long long)rows[3 * i + 2]));
        PyTuple_SET_ITEM(rec, 0, PyLong_FromLongLong((long long)rows[3 * i]));
        PyTuple_SET_ITEM(rec, 1, ab);
        Py_INCREF(rcobj);
        PyTuple_SET_ITEM(rec, 2, rcobj);
        PyList_SET_ITEM(lst, (Py_ssize_t)i, rec);
    }
    Py_DECREF(rcobj);
    return lst;
}

static PyObject *multi_common(Index *self, PyObject *args, PyObject *kwds, bool mems) {
    static const char *kwlist[] = {"minlength", "minn", nullptr};
    int minl = 0, minn = 2;
    if (!PyArg_ParseTupleAndKeywords(args, kwds, "|ii", (char **)kwlist, &minl, &minn)) return nullptr;
    if (need_arrays(self) != 0) return nullptr;
    int64_t nr = 0, nm = 0;
    int status;
    std::vector<int64_t> hdr, mem;
    if (self->mainidx || self->extracted) {  // a child of the recursion, or a root after extract(): sweep its own arrays
        if (mems) {
            PyErr_SetString(RevealError, "getmultimems() on a child index is not supported");
            return nullptr;
        }
        Py_BEGIN_ALLOW_THREADS;
        status = g_api.rv_sub_mums_multi(self->sub, minl, minn, &nr, &nm);
        Py_END_ALLOW_THREADS;
        if (fail_native(status) != 0) return nullptr;
        hdr.resize((size_t)(3 * nr + 3));
        mem.resize((size_t)(2 * nm + 2));
        if (fail_native(g_api.rv_sub_fetch(self->sub, hdr.data(), nr, mem.data(), nm)) != 0) return nullptr;
    } else {
        Py_BEGIN_ALLOW_THREADS;
        status = mems ? g_api.rv_mems_multi_count(self->h, minl, minn, &nr, &nm) : g_api.rv_mums_multi_count(self->h, minl, minn, &nr, &nm);
        Py_END_ALLOW_THREADS;
        if (fail_native(status) != 0) return nullptr;
        hdr.resize((size_t)(3 * nr + 3));
        mem.resize((size_t)(2 * nm + 2));
        if (fail_native(g_api.rv_mums_multi_fetch(self->h, hdr.data(), nr, mem.data(), nm)) != 0) return nullptr;
    }
    return multi_to_list(hdr, mem, nr, nm, !mems);
}
static PyObject *index_getmultimums(Index *self, PyObject *args, PyObject *kwds) { return multi_common(self, args, kwds, false); }
static PyObject *index_getmultimems(Index *self, PyObject *args, PyObject *kwds) { return multi_common(self, args, kwds, true); }

// ---- the recursion: `align` (interface.c:293-415) and the single-threaded `aligner` loop (reveal.c:731-1338) ---------------------
static bool parse_intervals(PyObject *obj, std::vector<int64_t> &flat, int64_t &total, std::vector<int64_t> &begins) {
    flat.clear();
    begins.clear();
    total = 0;
    PyObject *iter = PyObject_GetIter(obj);
    if (!iter) return false;
    PyObject *tup;
    while ((tup = PyIter_Next(iter))) {
        long long b, e;
        if (!PyArg_ParseTuple(tup, "LL", &b, &e)) {
            Py_DECREF(tup);
            Py_DECREF(iter);
            return false;
        }
        flat.push_back(b);
        flat.push_back(e);
        begins.push_back(b);
        total += e - b;
        Py_DECREF(tup);
    }
    Py_DECREF(iter);
    return !PyErr_Occurred();
}

// number of distinct samples among the interval starts (reveal.c:1026-1041)
static int count_samples(Index *root, const std::vector<int64_t> &begins) {
    if (begins.empty()) return 0;
    if (root->nsamples > 2) {
        std::vector<char> seen((size_t)root->nsamples, 0);
        int c = 0;
        for (int64_t b : begins) {
            size_t lo = 0, hi = root->nsep->size();  // SO[begin] = number of separators before begin
            while (lo < hi) {
                size_t mid = (lo + hi) >> 1;
                if ((*root->nsep)[mid] < b) lo = mid + 1; else hi = mid;
            }
            if (lo < seen.size() && !seen[lo]) { seen[lo] = 1; c++; }
        }
        return c;
    }
    int64_t nsep0 = (*root->nsep)[0];
    int left = 0, right = 0;
    for (int64_t b : begins) {
        if (b < nsep0) left = 1;
        else if (b > nsep0) right = 1;
    }
    return left + right;
}

static Index *new_child(Index *root, rv_sub *sub, int64_t n, int depth, int nsamples, PyObject *nodes, PyObject *left, PyObject *right, PyObject *skip) {
    Index *c = (Index *)index_new(&IndexType, nullptr, nullptr);
    if (!c) return nullptr;
    c->sub = sub;
    c->n = n;
    c->nT = root->nT;
    c->depth = depth;
    c->nsamples = nsamples;
    c->rc = root->rc;
    c->built = 1;
    Py_INCREF((PyObject *)root);
    c->mainidx = root;
    Py_INCREF(nodes);
    Py_SETREF(c->nodes, nodes);
    Py_INCREF(left);
    Py_SETREF(c->left_node, left);
    Py_INCREF(right);
    Py_SETREF(c->right_node, right);
    Py_INCREF(skip);
    Py_SETREF(c->skipmums, skip);
    return c;
}

// ---- pair MUM rows as a sequence object --------------------------------------------------------------------------------------
// The recursion hands a MUM list to the mumpicker at every step; built as Python tuples (four objects per MUM, reveal.c:167-169)
// that is a fifth of the host time of an alignment, and a native picker reads the numbers straight back out of them.  `mumrows`
// keeps the (l, a, b) rows the device produced: len() and indexing give the reference's tuples (l, 2, ((0, a), (1, b))) on demand
// -- a callback written for the list works on it by iteration and indexing -- and the buffer protocol exposes the int64 rows
// (remcore.Graph.mumpicker* reads those).  Only handed out when align() is called with mums_as_rows=True.
struct MumRows {
    PyObject_HEAD
    std::vector<int64_t> *rows;
};
static PyTypeObject MumRowsType = {PyVarObject_HEAD_INIT(nullptr, 0)};
static void mumrows_dealloc(MumRows *self) {
    delete self->rows;
    Py_TYPE(self)->tp_free((PyObject *)self);
}
static Py_ssize_t mumrows_len(MumRows *self) { return (Py_ssize_t)(self->rows->size() / 3); }
static PyObject *mumrows_item(MumRows *self, Py_ssize_t i) {
    const Py_ssize_t n = mumrows_len(self);
    if (i < 0 || i >= n) {
        PyErr_SetString(PyExc_IndexError, "mumrows index out of range");
        return nullptr;
    }
    const int64_t *r = self->rows->data() + 3 * i;
    return Py_BuildValue("(L,i,((i,L),(i,L)))", (long long)r[0], 2, 0, (long long)r[1], 1, (long long)r[2]);  // reveal.c:167-169
}
static int mumrows_getbuffer(MumRows *self, Py_buffer *view, int flags) {
    return PyBuffer_FillInfo(view, (PyObject *)self, self->rows->data(), (Py_ssize_t)(self->rows->size() * 8), 1, flags);
}
static PySequenceMethods mumrows_as_sequence = {(lenfunc)mumrows_len, nullptr, nullptr, (ssizeargfunc)mumrows_item};
static PyBufferProcs mumrows_as_buffer = {(getbufferproc)mumrows_getbuffer, nullptr};

// The same for multi-MUMs (more than two samples): `multimumrows` keeps the record rows (l, n, first) and the member rows
// (sample, position) of the device sweep.  len() and indexing give the reference's tuples (l, n, ((sample, position), ...)),
// reveal.c:497; a native picker reaches the arrays through the capsule attribute `_rv_multi` (struct RvMultiView below; the
// capsule keeps the object alive).
struct RvMultiView {      // shared with remcore_module.cpp by layout
    const int64_t *hdr;   // nrec x (l, n, first member)
    int64_t nrec;
    const int64_t *mem;   // nmem x (sample, position)
    int64_t nmem;
};
struct MultiMumRows {
    PyObject_HEAD
    std::vector<int64_t> *hdr, *mem;
    RvMultiView view;
};
static PyTypeObject MultiMumRowsType = {PyVarObject_HEAD_INIT(nullptr, 0)};
static void multimumrows_dealloc(MultiMumRows *self) {
    delete self->hdr;
    delete self->mem;
    Py_TYPE(self)->tp_free((PyObject *)self);
}
static Py_ssize_t multimumrows_len(MultiMumRows *self) { return (Py_ssize_t)self->view.nrec; }
static PyObject *multimumrows_item(MultiMumRows *self, Py_ssize_t i) {
    if (i < 0 || i >= (Py_ssize_t)self->view.nrec) {
        PyErr_SetString(PyExc_IndexError, "multimumrows index out of range");
        return nullptr;
    }
    const int64_t *h = self->view.hdr + 3 * i;
    const int64_t cnt = h[1], first = h[2];
    PyObject *members = PyTuple_New((Py_ssize_t)cnt);
    if (!members) return nullptr;
    for (int64_t x = 0; x < cnt; x++)
        PyTuple_SET_ITEM(members, (Py_ssize_t)x, Py_BuildValue("(lL)", (long)self->view.mem[2 * (first + x)], (long long)self->view.mem[2 * (first + x) + 1]));
    return Py_BuildValue("(LlN)", (long long)h[0], (long)cnt, members);
}
static void multiview_capsule_free(PyObject *cap) { Py_XDECREF((PyObject *)PyCapsule_GetContext(cap)); }
static PyObject *multimumrows_view(MultiMumRows *self, void *) {
    PyObject *cap = PyCapsule_New(&self->view, "reveal_b200.RvMultiView", multiview_capsule_free);
    if (!cap) return nullptr;
    Py_INCREF((PyObject *)self);
    PyCapsule_SetContext(cap, self);
    return cap;
}
static PySequenceMethods multimumrows_as_sequence = {(lenfunc)multimumrows_len, nullptr, nullptr, (ssizeargfunc)multimumrows_item};
static PyGetSetDef multimumrows_getset[] = {{"_rv_multi", (getter)multimumrows_view, nullptr, "capsule over the record and member rows", nullptr},
                                            {nullptr, nullptr, nullptr, nullptr, nullptr}};

// MUMs of a (sub)index in the shape the reference hands to mumpicker (reveal.c:802-829)
static PyObject *extract_mums(Index *root, rv_sub *sub, int minl, int minn, bool as_rows = false) {
    int64_t nr = 0, nm = 0;
    int status;
    if (root->nsamples > 2) {
        Py_BEGIN_ALLOW_THREADS;
        status = g_api.rv_sub_mums_multi(sub, minl, minn, &nr, &nm);
        Py_END_ALLOW_THREADS;
        if (fail_native(status) != 0) return nullptr;
        std::vector<int64_t> hdr((size_t)(3 * nr + 3)), mem((size_t)(2 * nm + 2));
        if (fail_native(g_api.rv_sub_fetch(sub, hdr.data(), nr, mem.data(), nm)) != 0) return nullptr;
        if (as_rows) {
            MultiMumRows *mr = (MultiMumRows *)MultiMumRowsType.tp_alloc(&MultiMumRowsType, 0);
            if (!mr) return nullptr;
            mr->hdr = new std::vector<int64_t>(std::move(hdr));
            mr->mem = new std::vector<int64_t>(std::move(mem));
            mr->view = RvMultiView{mr->hdr->data(), nr, mr->mem->data(), nm};
            return (PyObject *)mr;
        }
        return multi_to_list(hdr, mem, nr, nm, true);
    }
    Py_BEGIN_ALLOW_THREADS;
    status = g_api.rv_sub_mums_pair(sub, minl, &nr);
    Py_END_ALLOW_THREADS;
    if (fail_native(status) != 0) return nullptr;
    std::vector<int64_t> rows((size_t)(3 * nr + 3));
    if (fail_native(g_api.rv_sub_fetch(sub, rows.data(), nr, nullptr, 0)) != 0) return nullptr;
    if (as_rows) {
        MumRows *mr = (MumRows *)MumRowsType.tp_alloc(&MumRowsType, 0);
        if (!mr) return nullptr;
        rows.resize((size_t)(3 * nr));
        mr->rows = new std::vector<int64_t>(std::move(rows));
        return (PyObject *)mr;
    }
    GcPause gc_pause;
    PyObject *lst = PyList_New((Py_ssize_t)nr);
    PyObject *two = PyLong_FromLong(2), *zero = PyLong_FromLong(0), *one = PyLong_FromLong(1);
    for (int64_t i = 0; i < nr; i++) {  // (l, 2, ((0, a), (1, b))): reveal.c:167-169
        PyObject *sa = atomic_tuple(PyTuple_New(2)), *sb = atomic_tuple(PyTuple_New(2)), *mem = atomic_tuple(PyTuple_New(2)), *rec = atomic_tuple(PyTuple_New(3));
        Py_INCREF(zero);
        PyTuple_SET_ITEM(sa, 0, zero);
        PyTuple_SET_ITEM(sa, 1, PyLong_FromLongLong((long long)rows[3 * i + 1]));
        Py_INCREF(one);
        PyTuple_SET_ITEM(sb, 0, one);
        PyTuple_SET_ITEM(sb, 1, PyLong_FromLongLong((long long)rows[3 * i + 2]));
        PyTuple_SET_ITEM(mem, 0, sa);
        PyTuple_SET_ITEM(mem, 1, sb);
        PyTuple_SET_ITEM(rec, 0, PyLong_FromLongLong((long long)rows[3 * i]));
        Py_INCREF(two);
        PyTuple_SET_ITEM(rec, 1, two);
        PyTuple_SET_ITEM(rec, 2, mem);
        PyList_SET_ITEM(lst, (Py_ssize_t)i, rec);
    }
    Py_DECREF(two);
    Py_DECREF(zero);
    Py_DECREF(one);
    return lst;
}

// wall time of the parts of the last align() (seconds): where a recursion spends its time
struct AlignStats {
    double extract = 0, pick = 0, galign = 0, parse = 0, step = 0, child = 0, total = 0, prefix = 0;
    long long steps = 0, picks = 0, batches = 0;
};
static AlignStats g_align_stats;
static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// One recursion step whose callbacks have run and whose device part is waiting for the batch launch.
struct PendingStep {
    Index *idx = nullptr;                    // owned reference
    PyObject *pick = nullptr, *result = nullptr;  // owned: keep the borrowed objects below alive
    PyObject *skipleft = nullptr, *skipright = nullptr, *leading = nullptr, *trailing = nullptr, *rest = nullptr, *newleft = nullptr, *newright = nullptr;
    std::vector<int64_t> lead, trail, par, match, lead_b, trail_b, par_b, match_b, mum_sp;
    int64_t leadn = 0, trailn = 0, parn = 0, matchn = 0;
    long long mum_l = 0;
    int mum_n = 0;
    int32_t sweep[3] = {0, 0, 0};
};

static void release_index_view(Index *idx) {
    // the device view of a processed sub-index is released right away, like the reference frees SA/LCP
    if (idx->sub) {
        g_api.rv_sub_free(idx->sub);
        idx->sub = nullptr;
    }
    Py_DECREF((PyObject *)idx);
}

static PyObject *index_align(Index *self, PyObject *args, PyObject *kwds) {
    static const char *kwlist[] = {"mumpicker", "align", "threads", "wpen", "wscore", "minl", "minn",  // interface.c:303
                                   "shard_rank", "shard_world", "shard_grain", "mumpicker_batch", "mums_as_rows", nullptr};
    PyObject *mumpicker, *graphalign, *mumpicker_batch = nullptr;
    int threads = 0, wpen = 0, wscore = 0, minl = 0, minn = 0, shard_rank = 0, shard_world = 1, shard_grain = 2, mums_as_rows = 0;
    // (the reference frees the root's SA and LCP at its first split, reveal.c:1279-1284: a second align() -- which would run on an
    // inverse array the first one rewrote -- stops here as well)
    if (self->mainidx || !self->built || self->consumed) {
        PyErr_SetString(RevealError, "Index not yet constructed, alignment stopped.");  // interface.c:295-298
        return nullptr;
    }
    if (!PyArg_ParseTupleAndKeywords(args, kwds, "OO|iiiiiiiiOp", (char **)kwlist, &mumpicker, &graphalign, &threads, &wpen, &wscore, &minl, &minn,
                                     &shard_rank, &shard_world, &shard_grain, &mumpicker_batch, &mums_as_rows))
        return nullptr;
    // mumpicker_batch (optional): callable([(mums, idx, precomputed), ...], minlength=) -> [pick, ...] -- the mumpicker for all
    // sub-indexes of a frontier batch at once (their picks do not depend on each other), so that e.g. their chaining recurrences
    // can share a device launch (remcore.Graph.mumpicker_batch).  Without it `mumpicker` is called once per sub-index.
    if (mumpicker_batch == Py_None) mumpicker_batch = nullptr;
    if (shard_world < 1 || shard_rank < 0 || shard_rank >= shard_world || shard_grain < 1) {
        PyErr_SetString(RevealError, "align: bad shard_rank / shard_world / shard_grain");
        return nullptr;
    }
    // Sharded recursion (one process per GPU, every process holding the same index and the same graph): the sub-indexes of the
    // recursion are independent, so the tree is cut where its sub-indexes get smaller than n / (shard_grain * shard_world).
    // Everything above the cut is processed by EVERY rank (same input, same callbacks: same state everywhere, nothing to
    // exchange); the sub-indexes below it are the units: dealt out by size (largest first to the least loaded rank), and a rank
    // only descends into its own.  `shard_units` then lists (owner, nodes) of every unit so that the caller can collect the
    // parts of the graph each rank refined (reveal_b200/rem.py).
    bool prefix = shard_world > 1;
    const int64_t unit_max = self->n / ((int64_t)shard_grain * shard_world);
    std::vector<Index *> units;  // owned references
    Py_XSETREF(self->shard_units, PyList_New(0));
    // The reference hands the queue of sub-indexes to `threads` workers (interface.c:338-399), which can only overlap their C
    // parts: the callback sections are serialised (reveal.c:779-780).  Here the queue's concurrency goes to the device instead:
    // the sub-indexes on the queue are independent, so their callbacks run one after the other and the device parts of up to
    // `batch` steps share one launch and one synchronisation (rv_sub_step_batch).  threads <= 1 keeps one step per launch in
    // the reference's LIFO order; the default (0) and larger values batch.  RV_ALIGN_BATCH overrides the batch size.
    size_t batch_max = threads == 1 ? 1 : 256;   // sub-indexes taken off the queue per round (the device takes up to 256 steps per launch)
    if (const char *e = getenv("RV_ALIGN_BATCH")) {
        long b = atol(e);
        if (b >= 1) batch_max = (size_t)b;
    }
    self->depth = 0;
    if (!self->extracted) {  // (after extract() the index already is a view of its own: the recursion starts from that)
        if (self->sub) {
            g_api.rv_sub_free(self->sub);
            self->sub = nullptr;
        }
        if (fail_native(g_api.rv_sub_root(self->h, &self->sub)) != 0) return nullptr;
    }
    self->extracted = 0;  // align() consumes the root's arrays either way
    std::vector<Index *> queue;  // owned references
    Py_INCREF((PyObject *)self);
    queue.push_back(self);
    bool ok = true;
    AlignStats &as = g_align_stats;
    as = AlignStats();
    const double t_begin = now_s();
    PyObject *kw_minl = PyLong_FromLong(minl);
    std::vector<PendingStep *> batch;
    // Device / host overlap: the device part of a batch is only ENQUEUED (rv_sub_step_batch_begin); while it runs, the callbacks
    // of the next sub-indexes on the queue are served, and its children are collected (rv_sub_step_batch_end) right before the
    // next batch goes out.  (Taking a frontier that fits one batch in two halves, so that there always is something to overlap,
    // was measured and dropped: the host side of a launch costs more than the wait it hides -- RV_ALIGN_SPLIT=1 brings it back.)
    struct InFlight {
        std::vector<PendingStep *> batch;
        std::vector<rv_step_desc> descs;
        rv_step_batch *ticket = nullptr;
        bool open = false;
    } fly;
    const bool overlap = batch_max > 1 && !getenv("RV_ALIGN_NO_OVERLAP");
    const bool split_frontier = getenv("RV_ALIGN_SPLIT") != nullptr;
    // waits for the batch in flight, creates its children (pushed on the queue) and releases its steps
    auto finish = [&]() {
        if (!fly.open) return;
        fly.open = false;
        double t0 = now_s();
        int status;
        Py_BEGIN_ALLOW_THREADS;
        status = g_api.rv_sub_step_batch_end(fly.ticket);
        Py_END_ALLOW_THREADS;
        fly.ticket = nullptr;
        double t1 = now_s(); as.step += t1 - t0; t0 = t1;
        if (status != 0 || !ok) {
            if (ok) { fail_native(status); ok = false; }
            for (rv_step_desc &d : fly.descs)   // children of the steps that did succeed
                for (int c = 0; c < 3; c++)
                    if (d.children[c]) g_api.rv_sub_free(d.children[c]);
        } else {
            PyObject *empty = PyList_New(0);
            for (size_t i = 0; i < fly.batch.size(); i++) {
                PendingStep &p = *fly.batch[i];
                rv_sub **kids = fly.descs[i].children;
                const int depth = p.idx->depth + 1;
                Index *i_par = kids[2] ? new_child(self, kids[2], p.parn, depth, count_samples(self, p.par_b), p.rest, p.idx->left_node, p.idx->right_node, empty) : nullptr;
                Index *i_lead = kids[0] ? new_child(self, kids[0], p.leadn, depth, count_samples(self, p.lead_b), p.leading, p.idx->left_node, p.newright, p.skipleft) : nullptr;
                Index *i_trail = kids[1] ? new_child(self, kids[1], p.trailn, depth, count_samples(self, p.trail_b), p.trailing, p.newleft, p.idx->right_node, p.skipright) : nullptr;
                if (i_par) queue.push_back(i_par);      // push order of reveal.c:1296-1324
                if (i_lead) queue.push_back(i_lead);
                if (i_trail) queue.push_back(i_trail);
            }
            Py_DECREF(empty);
        }
        as.child += now_s() - t0;
        for (PendingStep *p : fly.batch) {
            Py_XDECREF(p->pick);
            Py_XDECREF(p->result);
            release_index_view(p->idx);
            delete p;
        }
        fly.batch.clear();
    };
    for (;;) {
    while (ok && (!queue.empty() || fly.open)) {
        if (queue.empty()) {  // nothing to overlap with: the children of the batch in flight are the work
            finish();
            continue;
        }
        size_t take = batch_max;
        if (overlap && split_frontier && queue.size() > 2 && queue.size() < 2 * batch_max) take = (queue.size() + 1) / 2;
        // ---- the sub-indexes on top of the queue (LIFO, reveal.c:21-26) and their MUM lists ----
        std::vector<PendingStep *> cand;
        std::vector<PyObject *> cand_mums;   // owned
        std::vector<int> cand_pre;
        if (!PyCallable_Check(mumpicker)) {
            PyErr_SetString(PyExc_TypeError, "**** mumpicker isn't callable");
            ok = false;
        }
        while (ok && !queue.empty() && cand.size() < take) {
            Index *idx = queue.back();
            queue.pop_back();
            if (prefix && idx != self && idx->n <= unit_max) {  // below the cut: a unit, dealt out once the part above the cut is done
                units.push_back(idx);
                continue;
            }
            PendingStep *ps = new PendingStep();
            ps->idx = idx;
            cand.push_back(ps);
            int precomputed = PyList_Check(idx->skipmums) ? PyList_Size(idx->skipmums) > 0 : PyObject_Length(idx->skipmums) > 0;
            PyObject *mums;
            if (!precomputed) {
                const double t0 = now_s();
                mums = extract_mums(self, idx->sub, minl, minn, mums_as_rows != 0);  // (mumrows / multimumrows)
                as.extract += now_s() - t0;
                if (!mums) ok = false;
            } else {
                mums = idx->skipmums;
                Py_INCREF(mums);
            }
            cand_mums.push_back(mums);
            cand_pre.push_back(precomputed);
        }
        // ---- callback 1, mumpicker (reveal.c:851): per sub-index, or for the whole batch in one call ----
        if (ok && mumpicker_batch && !cand.empty()) {
            const double t0 = now_s();
            PyObject *entries = PyList_New((Py_ssize_t)cand.size());
            for (size_t i = 0; i < cand.size(); i++)
                PyList_SET_ITEM(entries, (Py_ssize_t)i, Py_BuildValue("(OOO)", cand_mums[i], (PyObject *)cand[i]->idx, cand_pre[i] ? Py_True : Py_False));
            PyObject *picks = PyObject_CallFunctionObjArgs(mumpicker_batch, entries, kw_minl, nullptr);
            Py_DECREF(entries);
            as.pick += now_s() - t0;
            as.picks += (long long)cand.size();
            if (!picks) ok = false;
            else if (!PyList_Check(picks) || PyList_Size(picks) != (Py_ssize_t)cand.size()) {
                PyErr_SetString(RevealError, "**** mumpicker_batch must return one pick per entry");
                ok = false;
            } else {
                for (size_t i = 0; i < cand.size(); i++) {
                    cand[i]->pick = PyList_GET_ITEM(picks, (Py_ssize_t)i);
                    Py_INCREF(cand[i]->pick);
                }
            }
            Py_XDECREF(picks);
        }
        // ---- per sub-index: the pick, then callback 2, graphalign (reveal.c:939), one after the other ----
        for (size_t ci = 0; ci < cand.size(); ci++) {
            PendingStep *ps = cand[ci];
            Index *idx = ps->idx;
            bool staged = false;
            if (ok) do {
                double t0 = now_s(), t1;
                if (!ps->pick) {
                    PyObject *cargs = Py_BuildValue("(OO)", cand_mums[ci], (PyObject *)idx);
                    PyObject *ckw = Py_BuildValue("{s:O,s:O}", "precomputed", cand_pre[ci] ? Py_True : Py_False, "minlength", kw_minl);
                    ps->pick = PyObject_Call(mumpicker, cargs, ckw);  // reveal.c:851
                    Py_DECREF(cargs);
                    Py_DECREF(ckw);
                    t1 = now_s(); as.pick += t1 - t0; t0 = t1;
                    as.picks++;
                }
                if (!ps->pick) { ok = false; break; }
                if (!PyTuple_Check(ps->pick)) {
                    PyErr_SetString(RevealError, "**** call to mumpicker failed");
                    ok = false;
                    break;
                }
                if (PyTuple_Size(ps->pick) == 0) break;  // no more MUMs in this sub-index
                PyObject *mumobject, *spd;
                if (!PyArg_ParseTuple(ps->pick, "OOO", &mumobject, &ps->skipleft, &ps->skipright)) { ok = false; break; }
                if (!PyArg_ParseTuple(mumobject, "LiO", &ps->mum_l, &ps->mum_n, &spd)) { ok = false; break; }
                ps->mum_sp.assign((size_t)(ps->mum_n > 0 ? ps->mum_n : 1), 0);
                for (int i = 0; i < ps->mum_n; i++) {
                    PyObject *tup = PySequence_GetItem(spd, i);
                    PyObject *pos = tup ? PySequence_GetItem(tup, 1) : nullptr;
                    ps->mum_sp[(size_t)i] = pos ? PyLong_AsLongLong(pos) : 0;
                    Py_XDECREF(pos);
                    Py_XDECREF(tup);
                }
                if (PyErr_Occurred()) { ok = false; break; }
                t0 = now_s();
                ps->result = PyObject_CallFunctionObjArgs(graphalign, (PyObject *)idx, mumobject, nullptr);  // reveal.c:939
                t1 = now_s(); as.galign += t1 - t0; t0 = t1;
                if (!ps->result) { ok = false; break; }
                if (ps->result == Py_None) break;
                if (!PyTuple_Check(ps->result)) {
                    PyErr_SetString(RevealError, "**** call to graphalign failed");
                    ok = false;
                    break;
                }
                PyObject *matching, *merged;
                if (!PyArg_ParseTuple(ps->result, "OOOOOOO", &ps->leading, &ps->trailing, &matching, &ps->rest, &merged, &ps->newleft, &ps->newright)) {
                    PyErr_Clear();  // the reference silently drops an unparsable result (reveal.c:987-999)
                    break;
                }
                if (!parse_intervals(ps->leading, ps->lead, ps->leadn, ps->lead_b) || !parse_intervals(ps->trailing, ps->trail, ps->trailn, ps->trail_b) ||
                    !parse_intervals(ps->rest, ps->par, ps->parn, ps->par_b) || !parse_intervals(matching, ps->match, ps->matchn, ps->match_b)) {
                    ok = false;
                    break;
                }
                ps->sweep[0] = PyObject_Length(ps->skipleft) == 0;
                ps->sweep[1] = PyObject_Length(ps->skipright) == 0;
                ps->sweep[2] = 1;
                t1 = now_s(); as.parse += t1 - t0;
                staged = true;
            } while (0);
            Py_XDECREF(cand_mums[ci]);
            if (staged) {
                batch.push_back(ps);
            } else {
                Py_XDECREF(ps->pick);
                Py_XDECREF(ps->result);
                release_index_view(idx);
                delete ps;
            }
        }
        // ---- the batch in flight has had its overlap: collect it (its children go on the queue) ----
        finish();
        // ---- the device part of every staged step: one launch, left in flight ----
        if (ok && !batch.empty()) {
            double t0 = now_s();
            fly.descs.assign(batch.size(), rv_step_desc());
            for (size_t i = 0; i < batch.size(); i++) {
                PendingStep &p = *batch[i];
                rv_step_desc &d = fly.descs[i];
                memset(&d, 0, sizeof d);
                d.parent = p.idx->sub;
                d.lead = p.lead.data(); d.nlead = (int32_t)p.lead_b.size();
                d.trail = p.trail.data(); d.ntrail = (int32_t)p.trail_b.size();
                d.par = p.par.data(); d.npar = (int32_t)p.par_b.size();
                d.mum_sp = p.mum_sp.data(); d.mum_n = p.mum_n; d.mum_l = p.mum_l;
                d.matching = p.match.data(); d.nmatch = (int32_t)p.match_b.size();
                for (int c = 0; c < 3; c++) d.sweep[c] = p.sweep[c];
            }
            int status = g_api.rv_sub_step_batch_begin(fly.descs.data(), (int32_t)fly.descs.size(), minl, minn, &fly.ticket);
            as.step += now_s() - t0;
            as.steps += (long long)batch.size();
            as.batches++;
            self->tdirty = 1;  // matched bases were lower-cased on the device (reveal.c:1230-1234)
            for (PendingStep *p : batch)
                if (p->idx == self) self->consumed = 1;  // the root was split: its SA / LCP are gone (reveal.c:1279-1284)
            if (fail_native(status) != 0) {
                ok = false;
            } else {
                fly.batch.swap(batch);
                fly.open = true;
                if (!overlap) finish();
            }
        }
        for (PendingStep *p : batch) {
            Py_XDECREF(p->pick);
            Py_XDECREF(p->result);
            release_index_view(p->idx);
            delete p;
        }
        batch.clear();
    }
    finish();  // (an error left a batch in flight)
    if (!ok || !prefix) break;
    // ---- the part above the cut is done on every rank: deal the units out ----
    prefix = false;
    as.prefix = now_s() - t_begin;
    {
        std::vector<size_t> order(units.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return units[a]->n > units[b]->n; });
        std::vector<int64_t> load((size_t)shard_world, 0);
        std::vector<int> owner(units.size(), 0);
        for (size_t i : order) {
            int best = 0;
            for (int r = 1; r < shard_world; r++)
                if (load[(size_t)r] < load[(size_t)best]) best = r;
            owner[i] = best;
            load[(size_t)best] += units[i]->n;
        }
        for (size_t i = 0; i < units.size(); i++) {
            PyObject *nodes = PySequence_Tuple(units[i]->nodes);
            PyObject *rec = nodes ? Py_BuildValue("(iN)", owner[i], nodes) : nullptr;
            if (!rec || PyList_Append(self->shard_units, rec) != 0) ok = false;
            Py_XDECREF(rec);
        }
        for (size_t i = units.size(); i-- > 0;) {  // mine go back on the queue (first unit on top), the others are dropped
            if (ok && owner[i] == shard_rank) queue.push_back(units[i]);
            else release_index_view(units[i]);
        }
        units.clear();
    }
    if (!ok) break;
    }
    for (Index *q : units) release_index_view(q);
    for (Index *q : queue) release_index_view(q);
    Py_DECREF(kw_minl);
    as.total = now_s() - t_begin;
    if (!ok) return nullptr;
    Py_RETURN_NONE;
}

static PyObject *mod_align_stats(PyObject *, PyObject *) {
    const AlignStats &a = g_align_stats;
    return Py_BuildValue("{s:d,s:d,s:d,s:d,s:d,s:d,s:d,s:d,s:L,s:L,s:L}", "total_s", a.total, "above_cut_s", a.prefix, "sweep_fetch_s", a.extract, "mumpicker_s", a.pick, "graphalign_s", a.galign,
                         "parse_s", a.parse, "device_step_s", a.step, "children_s", a.child, "steps", a.steps, "mumpicker_calls", a.picks, "device_batches", a.batches);
}

// ---- splitindex (reveal.c:1515-1748): one recursion step driven from Python -------------------------------------------------------
// splitindex(leading, trailing, matching, rest, merged, newleft, newright, skipleft, skipright) -> (lead | None, trail | None, par | None)
// The same device step as inside align(): label scatter, T lower-casing of the matching intervals, 3-way split, bubble_sort of the
// leading child.  The matching intervals must all have one length (they are the occurrences of one exact match).
static PyObject *index_splitindex(Index *idx, PyObject *args) {
    PyObject *leading, *trailing, *matching, *rest, *merged, *newleft, *newright, *skipleft, *skipright;
    if (!PyArg_ParseTuple(args, "OOOOOOOOO", &leading, &trailing, &matching, &rest, &merged, &newleft, &newright, &skipleft, &skipright)) return nullptr;
    Index *root = root_of(idx);
    if (need_arrays(idx) != 0) return nullptr;
    if (!idx->sub) {
        if (idx->mainidx) {
            PyErr_SetString(RevealError, "splitindex: the arrays of this sub-index were already released");
            return nullptr;
        }
        if (fail_native(g_api.rv_sub_root(root->h, &idx->sub)) != 0) return nullptr;
    }
    std::vector<int64_t> lead, trail, par, match, lead_b, trail_b, par_b, match_b;
    int64_t leadn, trailn, parn, matchn;
    if (!parse_intervals(leading, lead, leadn, lead_b) || !parse_intervals(trailing, trail, trailn, trail_b) ||
        !parse_intervals(rest, par, parn, par_b) || !parse_intervals(matching, match, matchn, match_b))
        return nullptr;
    int64_t mum_l = match_b.empty() ? 0 : match[1] - match[0];
    for (size_t k = 0; k < match_b.size(); k++)
        if (match[2 * k + 1] - match[2 * k] != mum_l) {
            PyErr_SetString(RevealError, "splitindex: matching intervals of different lengths are not supported");
            return nullptr;
        }
    if (match_b.empty()) match_b.push_back(0);
    int32_t sweep[3] = {0, 0, 0};
    rv_sub *kids[3] = {nullptr, nullptr, nullptr};
    const int32_t nmatch = (int32_t)(match.size() / 2);
    int status;
    Py_BEGIN_ALLOW_THREADS;
    status = g_api.rv_sub_step(idx->sub, lead.data(), (int32_t)lead_b.size(), trail.data(), (int32_t)trail_b.size(), par.data(), (int32_t)par_b.size(),
                               match_b.data(), nmatch, mum_l, match.data(), nmatch, sweep, 0, 2, kids);
    Py_END_ALLOW_THREADS;
    if (fail_native(status) != 0) return nullptr;
    root->tdirty = 1;
    const int depth = idx->depth + 1;
    PyObject *empty = PyList_New(0);
    PyObject *out[3];
    Index *c;
    c = kids[0] ? new_child(root, kids[0], leadn, depth, count_samples(root, lead_b), leading, idx->left_node, newright, skipleft) : nullptr;
    out[0] = c ? (PyObject *)c : (Py_INCREF(Py_None), Py_None);
    c = kids[1] ? new_child(root, kids[1], trailn, depth, count_samples(root, trail_b), trailing, newleft, idx->right_node, skipright) : nullptr;
    out[1] = c ? (PyObject *)c : (Py_INCREF(Py_None), Py_None);
    c = kids[2] ? new_child(root, kids[2], parn, depth, count_samples(root, par_b), rest, idx->left_node, idx->right_node, empty) : nullptr;
    out[2] = c ? (PyObject *)c : (Py_INCREF(Py_None), Py_None);
    Py_DECREF(empty);
    return Py_BuildValue("(NNN)", out[0], out[1], out[2]);
}

// extract(intervals) (reveal.c:1386-1505): takes the text positions of the (begin, end) intervals out of the index in place.
// With rc=1 the intervals of the second sample are mapped back first AND the caller's list is rewritten with the mapped
// tuples, like the reference does (:1411-1427).
static PyObject *index_extract(Index *self, PyObject *args) {
    PyObject *intervals;
    if (!PyArg_ParseTuple(args, "O", &intervals)) return nullptr;
    if (self->mainidx || need_arrays(self) != 0) {
        if (self->mainidx) PyErr_SetString(RevealError, "extract() on a child index is not supported");
        return nullptr;
    }
    PyObject *seq = PySequence_Fast(intervals, "intervals must be a sequence of (begin, end)");
    if (!seq) return nullptr;
    const Py_ssize_t m = PySequence_Fast_GET_SIZE(seq);
    std::vector<int64_t> flat;
    const int64_t nsep0 = self->nsep->empty() ? -1 : (*self->nsep)[0];
    for (Py_ssize_t x = 0; x < m; x++) {
        long long b, e;
        if (!PyArg_ParseTuple(PySequence_Fast_GET_ITEM(seq, x), "LL", &b, &e)) { Py_DECREF(seq); return nullptr; }
        if (self->rc == 1 && b > nsep0) {  // map qry coordinates back (reveal.c:1411-1416)
            const long long nb = nsep0 + (self->nT - b - (e - b)), ne = nsep0 + (self->nT - b);
            b = nb;
            e = ne;
            if (PyList_Check(intervals)) PyList_SetItem(intervals, x, Py_BuildValue("(LL)", b, e));
        }
        flat.push_back(b);
        flat.push_back(e);
    }
    Py_DECREF(seq);
    if (!self->sub && fail_native(g_api.rv_sub_root(self->h, &self->sub)) != 0) return nullptr;
    int status;
    Py_BEGIN_ALLOW_THREADS;
    status = g_api.rv_sub_extract(self->sub, flat.data(), (int32_t)(flat.size() / 2));
    Py_END_ALLOW_THREADS;
    if (fail_native(status) != 0) return nullptr;
    self->n = g_api.rv_sub_n ? g_api.rv_sub_n(self->sub) : self->n;
    self->extracted = 1;
    self->tdirty = 1;  // the extracted bases were lower-cased on the device (reveal.c:1432)
    Py_RETURN_NONE;
}

// puttext(begin, text): overwrites a stretch of the indexed text (host copy and device) -- used when the parts of a sharded
// recursion are collected: the owner of a unit sends the stretches whose matched bases it lower-cased.
static PyObject *index_puttext(Index *self, PyObject *args) {
    long long begin;
    const char *txt;
    Py_ssize_t len;
    if (!PyArg_ParseTuple(args, "Ls#", &begin, &txt, &len)) return nullptr;
    if (self->mainidx || need_built(self, RevealError, "Index not yet constructed.") != 0) return nullptr;
    if (begin < 0 || begin + (long long)len > self->n) {
        PyErr_SetString(RevealError, "puttext: range outside the text");
        return nullptr;
    }
    if (fail_native(g_api.rv_put_text(self->h, begin, (const uint8_t *)txt, (int64_t)len)) != 0) return nullptr;
    if (!self->tdirty) memcpy(self->T->p + begin, txt, (size_t)len);  // (a dirty host copy is refreshed from the device anyway)
    Py_RETURN_NONE;
}

// ---- copy (interface.c:432-470) -----------------------------------------------------------------------------------------------
static PyObject *index_copy(Index *self, PyObject *) {
    if (self->mainidx || need_arrays(self) != 0) return nullptr;
    Index *c = (Index *)index_new(&IndexType, nullptr, nullptr);
    if (!c) return nullptr;
    if (self->tdirty) {
        if (fail_native(g_api.rv_get_text(self->h, (uint8_t *)self->T->p)) != 0) { Py_DECREF(c); return nullptr; }
        self->tdirty = 0;
    }
    if (!c->T->assign(*self->T)) { Py_DECREF(c); return PyErr_NoMemory(); }
    *c->nsep = *self->nsep;
    c->n = self->n;
    c->nT = self->nT;
    c->nsamples = self->nsamples;
    c->rc = self->rc;
    Py_SETREF(c->samples, PySequence_List(self->samples));
    Py_SETREF(c->nodes, PySet_New(self->nodes));
    std::vector<int32_t> sa((size_t)self->n), lcp((size_t)self->n);
    if (ensure_handle(c) != 0 || fail_native(g_api.rv_get_sa(self->h, sa.data(), 32)) != 0 || fail_native(g_api.rv_get_lcp(self->h, lcp.data(), 32)) != 0 ||
        fail_native(g_api.rv_build_cached(c->h, (const uint8_t *)c->T->data(), c->n, c->nsep->empty() ? nullptr : c->nsep->data(), c->nsamples, 0, sa.data(),
                                          lcp.data())) != 0) {
        Py_DECREF(c);
        return nullptr;
    }
    c->built = 1;
    return (PyObject *)c;
}

// ---- getters (interface.c:539-785) ----------------------------------------------------------------------------------------------
static PyObject *ints_to_list(const std::vector<int32_t> &v, bool as_unsigned) {
    PyObject *lst = PyList_New((Py_ssize_t)v.size());
    if (!lst) return nullptr;
    for (size_t i = 0; i < v.size(); i++)
        PyList_SET_ITEM(lst, (Py_ssize_t)i, as_unsigned ? PyLong_FromUnsignedLong((uint32_t)v[i]) : PyLong_FromLong(v[i]));
    return lst;
}

static PyObject *get_array(Index *self, int which) {  // 0 SA, 1 SAi, 2 LCP
    if (need_built(self, PyExc_TypeError, "Index not yet constructed.") != 0) return nullptr;
    Index *r = root_of(self);
    if (!self->mainidx && self->consumed && which != 1) {  // interface.c:548-551 / :590-593 after reveal.c:1279-1284
        PyErr_SetString(PyExc_TypeError, "Index not yet constructed.");
        return nullptr;
    }
    std::vector<int32_t> v;
    int status;
    if ((self->mainidx || self->extracted) && which != 1) {
        if (!self->sub) {
            PyErr_SetString(PyExc_TypeError, "Index not yet constructed.");  // SA/LCP of a processed sub-index are freed
            return nullptr;
        }
        v.resize((size_t)self->n);
        status = g_api.rv_sub_get(self->sub, which == 0 ? 0 : 1, v.data());
    } else {
        v.resize((size_t)r->nT);   // the handle's arrays have one entry per character of the text (nT >= n after extract)
        status = which == 0 ? g_api.rv_get_sa(r->h, v.data(), 32) : (which == 1 ? g_api.rv_get_sai(r->h, v.data(), 32) : g_api.rv_get_lcp(r->h, v.data(), 32));
        v.resize((size_t)r->n);    // the reference lists n entries (interface.c:546-655)
    }
    if (fail_native(status) != 0) return nullptr;
#ifdef SA64
    return ints_to_list(v, which == 2);  // lcp_t is unsigned in the 64-bit build (reveal.h:8-9)
#else
    return ints_to_list(v, false);
#endif
}
static PyObject *get_SA(Index *self, void *) { return get_array(self, 0); }
static PyObject *get_SAi(Index *self, void *) { return get_array(self, 1); }
static PyObject *get_LCP(Index *self, void *) { return get_array(self, 2); }

static PyObject *get_SO(Index *self, void *) {
    Index *r = root_of(self);
    if (!r->built || r->nsamples <= 2) {
        PyErr_SetString(PyExc_TypeError, "SO not available.");  // interface.c:575-579
        return nullptr;
    }
    std::vector<uint16_t> v((size_t)r->n);
    if (fail_native(g_api.rv_get_so(r->h, v.data())) != 0) return nullptr;
    PyObject *lst = PyList_New((Py_ssize_t)v.size());
    for (size_t i = 0; i < v.size(); i++) PyList_SET_ITEM(lst, (Py_ssize_t)i, PyLong_FromLong(v[i]));
    return lst;
}

static PyObject *get_T(Index *self, void *) {
    Index *r = root_of(self);
    if (r->built && r->tdirty) {  // rc or align changed the device text: refresh the host copy
        if (fail_native(g_api.rv_get_text(r->h, (uint8_t *)r->T->p)) != 0) return nullptr;
        r->tdirty = 0;
    }
    return PyUnicode_DecodeLatin1(r->T->data(), (Py_ssize_t)r->T->size(), nullptr);
}

static PyObject *get_n(Index *self, void *) { return PyLong_FromLongLong(self->n); }
static PyObject *get_depth(Index *self, void *) { return PyLong_FromLong(self->depth); }
static PyObject *get_nsamples(Index *self, void *) { return PyLong_FromLong(self->nsamples); }
static PyObject *get_samples(Index *self, void *) { PyObject *o = root_of(self)->samples; Py_INCREF(o); return o; }
static PyObject *get_nodes(Index *self, void *) { Py_INCREF(self->nodes); return self->nodes; }
static PyObject *get_leftnode(Index *self, void *) { Py_INCREF(self->left_node); return self->left_node; }
static PyObject *get_rightnode(Index *self, void *) { Py_INCREF(self->right_node); return self->right_node; }
static PyObject *get_skipmums(Index *self, void *) { Py_INCREF(self->skipmums); return self->skipmums; }
static PyObject *get_shard_units(Index *self, void *) { Py_INCREF(self->shard_units); return self->shard_units; }
static PyObject *get_nsep(Index *self, void *) {
    Index *r = root_of(self);
    PyObject *lst = PyList_New((Py_ssize_t)r->nsep->size());
    for (size_t i = 0; i < r->nsep->size(); i++) PyList_SET_ITEM(lst, (Py_ssize_t)i, PyLong_FromLongLong((*r->nsep)[i]));
    return lst;
}
static PyObject *get_main(Index *self, void *) {
    PyObject *o = (PyObject *)root_of(self);
    Py_INCREF(o);
    return o;
}

static PyObject *index_times(Index *self, PyObject *) {
    Index *r = root_of(self);
    if (!r->h) Py_RETURN_NONE;
    rv_times t;
    if (fail_native(g_api.rv_get_times(r->h, &t)) != 0) return nullptr;
    return Py_BuildValue("{s:f,s:f,s:f,s:f,s:f,s:f,s:i,s:i,s:L}", "h2d_ms", t.h2d_ms, "pack_ms", t.pack_ms, "sa_ms", t.sa_ms, "lcp_ms", t.lcp_ms, "so_ms",
                         t.so_ms, "total_ms", t.total_ms, "sa_rounds", t.sa_rounds, "launches", t.launches, "sa_sorted_items", (long long)t.sa_sorted_items);
}

static PyObject *index_reduce(Index *, PyObject *) { Py_RETURN_NONE; }  // interface.c:417-422

static PyMethodDef index_methods[] = {
    {"align", (PyCFunction)index_align, METH_VARARGS | METH_KEYWORDS, nullptr},
    {"splitindex", (PyCFunction)index_splitindex, METH_VARARGS,
     "splitindex(leading, trailing, matching, rest, merged, newleft, newright, skipleft, skipright) -> (lead, trail, par): one recursion step."},
    {"copy", (PyCFunction)index_copy, METH_NOARGS, nullptr},
    {"extract", (PyCFunction)index_extract, METH_VARARGS, "extract(intervals): take the text positions of the (begin, end) intervals out of the index in place (reveal.c:1386-1505)."},
    {"puttext", (PyCFunction)index_puttext, METH_VARARGS, "puttext(begin, text): overwrite a stretch of the indexed text (sharded recursion: collecting the parts)."},
    {"addsample", (PyCFunction)index_addsample, METH_VARARGS, nullptr},
    {"addsequence", (PyCFunction)index_addsequence, METH_VARARGS, nullptr},
    {"construct", (PyCFunction)index_construct, METH_VARARGS | METH_KEYWORDS, nullptr},
    {"getmultimums", (PyCFunction)index_getmultimums, METH_VARARGS | METH_KEYWORDS, nullptr},
    {"getmultimems", (PyCFunction)index_getmultimems, METH_VARARGS | METH_KEYWORDS, nullptr},
    {"getmums", (PyCFunction)index_getmums, METH_VARARGS, nullptr},
    {"times", (PyCFunction)index_times, METH_NOARGS, "device milliseconds of the last construct() per phase"},
    {"__reduce__", (PyCFunction)index_reduce, METH_NOARGS, "For pickle"},
    {nullptr, nullptr, 0, nullptr}};

static PyGetSetDef index_getset[] = {
    {"n", (getter)get_n, nullptr, "Number of characters in the index.", nullptr},
    {"depth", (getter)get_depth, nullptr, "Depth of the index within the recursion tree.", nullptr},
    {"nsamples", (getter)get_nsamples, nullptr, "Number of samples in the index.", nullptr},
    {"samples", (getter)get_samples, nullptr, "Sample/file names used in the index.", nullptr},
    {"nodes", (getter)get_nodes, nullptr, "The set of intervals (nodes) associated with the index.", nullptr},
    {"leftnode", (getter)get_leftnode, nullptr, "Interval of the node bounding the index on the left.", nullptr},
    {"rightnode", (getter)get_rightnode, nullptr, "Interval of the node bounding the index on the right.", nullptr},
    {"skipmums", (getter)get_skipmums, nullptr, "Precomputed MUMs handed down by the mumpicker.", nullptr},
    {"nsep", (getter)get_nsep, nullptr, "Positions of the sentinels that separate the samples.", nullptr},
    {"SA", (getter)get_SA, nullptr, "The suffix array of the concatenation of input texts.", nullptr},
    {"SAi", (getter)get_SAi, nullptr, "The inverse of the suffix array.", nullptr},
    {"SO", (getter)get_SO, nullptr, "Sample id of every text position (more than two samples).", nullptr},
    {"LCP", (getter)get_LCP, nullptr, "Longest common prefix of consecutive suffixes.", nullptr},
    {"T", (getter)get_T, nullptr, "The concatenation of the input texts.", nullptr},
    {"main", (getter)get_main, nullptr, "The root index.", nullptr},
    {"shard_units", (getter)get_shard_units, nullptr, "After align(shard_world > 1): [(owner rank, nodes of the unit), ...] of the recursion's cut.", nullptr},
    {nullptr, nullptr, nullptr, nullptr, nullptr}};

// ---- module -------------------------------------------------------------------------------------------------------------------------
static bool g_test_hooks = false;  // REVEAL_B200_TEST_HOOKS=1 in the environment when the module was imported
static PyObject *mod_load(PyObject *, PyObject *args) {
    const char *path;
    if (!PyArg_ParseTuple(args, "s", &path)) return nullptr;
    if (!g_test_hooks) {
        PyErr_SetString(PyExc_RuntimeError, MODNAME "._load is a test hook: it only works when REVEAL_B200_TEST_HOOKS=1 was set before the import");
        return nullptr;
    }
    if (!load_library(path)) return nullptr;
    Py_RETURN_NONE;
}
static PyObject *mod_library(PyObject *, PyObject *) {
    if (!g_api.handle) Py_RETURN_NONE;
    return Py_BuildValue("(s,s)", g_api.path.c_str(), g_api.rv_version());
}

// chain_dp(start, length, gain, wpen, model, link, score) -- the O(m^2) chaining recurrence of the REM driver
// (reveal_b200/rem.py:chain; reference: reveal/schemes.py:20-104 with utils.gapcost) on int64 buffers:
//   start [m+1][k] coordinates (row 0 = the left bound, rows 1..m the anchors in processing order, row m = the right bound),
//   length/gain [m+1]; model 0 = sumofpairs, 1 = star-avg, 2 = star-med; writes link[m+1] and score[m+1].
// Row r may follow every earlier row that ends at or before it in every coordinate; equal totals go to the predecessor with
// the higher score, then to the one that became available earlier, then to the one processed earlier.
static PyObject *mod_chain_dp(PyObject *, PyObject *args) {
    Py_buffer bs, bl, bg, bk, bc;
    long long wpen;
    int model;
    if (!PyArg_ParseTuple(args, "y*y*y*Liw*w*", &bs, &bl, &bg, &wpen, &model, &bk, &bc)) return nullptr;
    const Py_ssize_t rows = bl.len / 8;
    const Py_ssize_t k = rows ? bs.len / 8 / rows : 0;
    PyObject *ret = nullptr;
    if (rows < 1 || k < 1 || bs.len != rows * k * 8 || bg.len != rows * 8 || bk.len != rows * 8 || bc.len != rows * 8 || k > 64) {
        PyErr_SetString(PyExc_ValueError, "chain_dp: inconsistent buffer sizes");
    } else {
        const int64_t *start = (const int64_t *)bs.buf, *length = (const int64_t *)bl.buf, *gain = (const int64_t *)bg.buf;
        int64_t *link = (int64_t *)bk.buf, *score = (int64_t *)bc.buf;
        Py_BEGIN_ALLOW_THREADS
        rv_chain_dp(start, length, gain, (long)rows, (long)k, (int64_t)wpen, model, link, score);
        Py_END_ALLOW_THREADS
        ret = Py_None;
        Py_INCREF(ret);
    }
    PyBuffer_Release(&bs); PyBuffer_Release(&bl); PyBuffer_Release(&bg); PyBuffer_Release(&bk); PyBuffer_Release(&bc);
    return ret;
}

// getmums_batch(pairs, minlength) -> [getmums list of pair 0, ...]
// Every (ref, qry) pair of short sequences is an index of its own -- index(); addsample; addsequence(ref); addsample;
// addsequence(qry); construct(); getmums(minlength) -- the way `extend` of finish / transform indexes the <= 200 bp flanks of
// every anchor (reveal/transformold.py:1170-1240); the whole list goes through ONE launch (rv_mums_tiny_batch, one thread block
// per pair).  A pair longer than the block path holds, or with more MUMs than a block reports, is built as a regular index.
static rv_index *g_batch_ws = nullptr;
static PyObject *pair_rows_to_list(const int64_t *rows, int64_t k) {
    GcPause gc_pause;
    PyObject *lst = PyList_New((Py_ssize_t)k);
    if (!lst) return nullptr;
    PyObject *zero = PyLong_FromLong(0);
    for (int64_t i = 0; i < k; i++) {  // (l, (a, b), rc): reveal.c:102-106
        PyObject *ab = atomic_tuple(PyTuple_New(2)), *rec = atomic_tuple(PyTuple_New(3));
        PyTuple_SET_ITEM(ab, 0, PyLong_FromLongLong((long long)rows[3 * i + 1]));
        PyTuple_SET_ITEM(ab, 1, PyLong_FromLongLong((long long)rows[3 * i + 2]));
        PyTuple_SET_ITEM(rec, 0, PyLong_FromLongLong((long long)rows[3 * i]));
        PyTuple_SET_ITEM(rec, 1, ab);
        Py_INCREF(zero);
        PyTuple_SET_ITEM(rec, 2, zero);
        PyList_SET_ITEM(lst, (Py_ssize_t)i, rec);
    }
    Py_DECREF(zero);
    return lst;
}
static PyObject *mod_getmums_batch(PyObject *, PyObject *args) {
    PyObject *pairs;
    int minl = 0;
    if (!PyArg_ParseTuple(args, "Oi", &pairs, &minl)) return nullptr;
    if (!api_ready()) return nullptr;
    PyObject *seq = PySequence_Fast(pairs, "pairs must be a sequence of (ref, qry)");
    if (!seq) return nullptr;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    std::string T;
    std::vector<int64_t> off(1, 0), sep;
    const int32_t cap = 64;
    const int64_t tiny_max = 1024;
    std::vector<char> tiny((size_t)n, 1);
    std::vector<int64_t> unit_of;   // batch unit -> pair
    std::string Tb;
    std::vector<int64_t> boff(1, 0), bsep;
    for (Py_ssize_t i = 0; i < n; i++) {
        const char *a, *b;
        Py_ssize_t la, lb;
        if (!PyArg_ParseTuple(PySequence_Fast_GET_ITEM(seq, i), "s#s#", &a, &la, &b, &lb)) { Py_DECREF(seq); return nullptr; }
        off.push_back(off.back() + la + lb + 2);
        sep.push_back(la);
        T.append(a, (size_t)la); T.push_back('$'); T.append(b, (size_t)lb); T.push_back('$');
        if (la + lb + 2 > tiny_max) { tiny[(size_t)i] = 0; continue; }
        unit_of.push_back(i);
        Tb.append(a, (size_t)la); Tb.push_back('$'); Tb.append(b, (size_t)lb); Tb.push_back('$');
        boff.push_back(boff.back() + la + lb + 2);
        bsep.push_back(la);
    }
    Py_DECREF(seq);
    if (!g_batch_ws && fail_native(g_api.rv_index_create(&g_batch_ws, nullptr)) != 0) return nullptr;
    const int32_t nu = (int32_t)unit_of.size();
    std::vector<int64_t> rows((size_t)nu * cap * 3 + 3);
    std::vector<int32_t> counts((size_t)nu + 1, 0);
    if (nu > 0) {
        int status;
        Py_BEGIN_ALLOW_THREADS;
        status = g_api.rv_mums_tiny_batch(g_batch_ws, (const uint8_t *)Tb.data(), boff.data(), bsep.data(), nu, minl, cap, rows.data(), counts.data());
        Py_END_ALLOW_THREADS;
        if (fail_native(status) != 0) return nullptr;
    }
    PyObject *out = PyList_New(n);
    if (!out) return nullptr;
    for (int32_t u = 0; u < nu; u++) {
        const int64_t i = unit_of[(size_t)u];
        if (counts[(size_t)u] < 0 || counts[(size_t)u] > cap) { tiny[(size_t)i] = 0; continue; }
        PyObject *lst = pair_rows_to_list(rows.data() + (size_t)u * cap * 3, counts[(size_t)u]);
        if (!lst) { Py_DECREF(out); return nullptr; }
        PyList_SET_ITEM(out, (Py_ssize_t)i, lst);
    }
    for (Py_ssize_t i = 0; i < n; i++) {  // the pairs the block path did not take: a regular index each
        if (tiny[(size_t)i]) continue;
        const int64_t len = off[(size_t)i + 1] - off[(size_t)i];
        int64_t nsep = sep[(size_t)i], k = 0;
        int status = g_api.rv_build(g_batch_ws, (const uint8_t *)T.data() + off[(size_t)i], len, &nsep, 2, 0);
        if (status == 0) status = g_api.rv_mums_pair_count(g_batch_ws, minl, 0, &k);
        std::vector<int64_t> r((size_t)(3 * k + 3));
        if (status == 0) status = g_api.rv_mums_pair_fetch(g_batch_ws, r.data(), k);
        if (fail_native(status) != 0) { Py_DECREF(out); return nullptr; }
        PyObject *lst = pair_rows_to_list(r.data(), k);
        if (!lst) { Py_DECREF(out); return nullptr; }
        PyList_SET_ITEM(out, i, lst);
    }
    return out;
}

static PyMethodDef module_methods[] = {
    {"getmums_batch", mod_getmums_batch, METH_VARARGS,
     "getmums_batch([(ref, qry), ...], minlength) -> [getmums list of every pair]: many tiny two-sample indexes in one launch (finish/transform extend)."},
    {"chain_dp", mod_chain_dp, METH_VARARGS, "Chaining recurrence of the REM driver on int64 buffers (see reveal_b200/rem.py:chain)."},
    {"_load", mod_load, METH_VARARGS, "Load a shared library exporting the C-ABI of include/reveal_b200.h (tests inject the emulated kernels)."},
    {"_library", mod_library, METH_NOARGS, "(path, version) of the loaded C-ABI library."},
    {"align_stats", mod_align_stats, METH_NOARGS, "Wall time of the parts of the last index.align() of this module (seconds), and its step count."},
    {nullptr, nullptr, 0, nullptr}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, MODNAME,
                                       "REVEAL index with the build, the MUM sweeps and the recursion steps on a B200 (libreveal_b200.so).", -1,
                                       module_methods};

PyMODINIT_FUNC MODINIT(void) {
    IndexType.tp_name = MODNAME ".index";
    IndexType.tp_basicsize = sizeof(Index);
    IndexType.tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_BASETYPE;
    IndexType.tp_doc = "index objects";
    IndexType.tp_new = index_new;
    IndexType.tp_init = (initproc)index_init;
    IndexType.tp_dealloc = (destructor)index_dealloc;
    IndexType.tp_methods = index_methods;
    IndexType.tp_getset = index_getset;
    if (PyType_Ready(&IndexType) < 0) return nullptr;
    MumRowsType.tp_name = MODNAME ".mumrows";
    MumRowsType.tp_basicsize = sizeof(MumRows);
    MumRowsType.tp_flags = Py_TPFLAGS_DEFAULT;
    MumRowsType.tp_doc = "pair MUM rows (l, a, b) of a sub-index: a sequence of the reference's MUM tuples, and a buffer of int64 rows";
    MumRowsType.tp_dealloc = (destructor)mumrows_dealloc;
    MumRowsType.tp_as_sequence = &mumrows_as_sequence;
    MumRowsType.tp_as_buffer = &mumrows_as_buffer;
    if (PyType_Ready(&MumRowsType) < 0) return nullptr;
    MultiMumRowsType.tp_name = MODNAME ".multimumrows";
    MultiMumRowsType.tp_basicsize = sizeof(MultiMumRows);
    MultiMumRowsType.tp_flags = Py_TPFLAGS_DEFAULT;
    MultiMumRowsType.tp_doc = "multi-MUM rows of a sub-index: a sequence of the reference's tuples (l, n, ((sample, position), ...))";
    MultiMumRowsType.tp_dealloc = (destructor)multimumrows_dealloc;
    MultiMumRowsType.tp_as_sequence = &multimumrows_as_sequence;
    MultiMumRowsType.tp_getset = multimumrows_getset;
    if (PyType_Ready(&MultiMumRowsType) < 0) return nullptr;
    PyObject *m = PyModule_Create(&moduledef);
    if (!m) return nullptr;
    Py_INCREF(&IndexType);
    PyModule_AddObject(m, "index", (PyObject *)&IndexType);
    RevealError = PyErr_NewException(MODNAME ".error", nullptr, nullptr);
    Py_INCREF(RevealError);
    PyModule_AddObject(m, "error", RevealError);
    {
        const char *e = getenv("REVEAL_B200_TEST_HOOKS");
        g_test_hooks = e && e[0] == '1';
    }
    // default library: libreveal_b200.so next to this module
    Dl_info info;
    if (dladdr((void *)&MODINIT, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t slash = p.rfind('/');
        p = (slash == std::string::npos ? std::string(".") : p.substr(0, slash)) + "/libreveal_b200.so";
        if (!load_library(p.c_str())) PyErr_Clear();  // import succeeds; every call raises until a library is loaded
    }
    return m;
}
