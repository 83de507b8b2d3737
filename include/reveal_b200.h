/*
 * reveal_b200.h -- C-ABI of libreveal_b200.so: the B200-native index build and
 * MUM sweeps behind the `reveallib` extension surface of jasperlinthorst/reveal.
 *
 * Every entry point is what the reference's extension glue would bind in place
 * of its CPU code for this path (paths relative to the reference tree):
 *
 *   rv_build / rv_build_device   <- construct():  reveallib/interface.c:160-291
 *                                   (revcomp :148-158,168-172; divsufsort :213-222;
 *                                    inverse SA :235-238; compute_lcp :97-114,253;
 *                                    build_SO :116-134,265-271)
 *   rv_get_*                     <- array getters of the index type, interface.c:539-655
 *   rv_mums_pair                 <- getmums  reveallib/reveal.c:55-116  (flavour 0)
 *                                   getmums_rem         reveal.c:119-180 (flavour 1)
 *   rv_mums_multi                <- getmultimums        reveal.c:436-580 + ismultimum :227-259
 *
 * Conventions: plain C types only; every function returns 0 on success or a
 * negative rv_status; rv_last_error() holds the message of the last failure on
 * the calling thread; outputs are caller-allocated host buffers unless the name
 * says "device".  Index entries are int32 on the device (n < 2^30); with
 * idx_bits == 64 the getters widen to int64 / uint32 like the reference's
 * reveallib64 build (reveallib/reveal.h:7-13) -- the numbers are the same.
 * There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef REVEAL_B200_H
#define REVEAL_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct rv_index rv_index; /* opaque: device-resident T, SA, SAi, LCP, SO + workspace */

enum rv_status {
    RV_OK = 0,
    RV_ERR_ARG = -1,         /* bad argument */
    RV_ERR_CUDA = -2,        /* CUDA runtime error (message in rv_last_error) */
    RV_ERR_NOMEM = -3,       /* out of host or device memory */
    RV_ERR_UNSUPPORTED = -4, /* e.g. n >= 2^30 */
    RV_ERR_STATE = -5        /* call order (e.g. sweep before build) */
};

/* per-phase device time of the last build, CUDA events on the build stream (ms) */
typedef struct rv_times {
    float h2d_ms, pack_ms, sa_ms, lcp_ms, so_ms, total_ms;
    int32_t sa_rounds;        /* prefix-doubling rounds */
    int32_t launches;         /* kernels this library launched for the build */
    int64_t sa_sorted_items;  /* sum over radix sorts of their item counts */
} rv_times;

const char *rv_last_error(void);
const char *rv_version(void);
int rv_device_count(int *count);
int rv_set_device(int device);

/* Handles.  A handle owns a grow-only device workspace that is reused by every
 * build on it, so steady-state builds never call cudaMalloc.
 * stream: a cudaStream_t passed as void* (NULL = a private non-blocking stream). */
int rv_index_create(rv_index **out, void *stream);
void rv_index_free(rv_index *idx);

/* Pinned host memory for the text handed to rv_build: the extension's addsequence (interface.c:45-95, which reallocs
 * its T) appends straight into such a block, so construct() is one DMA without a staging copy.  Blocks -- like the
 * device workspaces of freed handles -- are kept by the library for the next user; rv_trim() releases what is cached. */
int rv_host_alloc(int64_t bytes, void **ptr);
void rv_host_free(void *ptr);
int rv_trim(void);

/* construct(): T = concatenated text of n bytes ('$' after every sequence,
 * interface.c:71-85), nsep[k] = position of the last '$' of sample k
 * (nsamples-1 entries, interface.c:36-43), rc = 1 reverse-complements
 * T[nsep[0]..n) first (interface.c:168-172).  rv_build copies T from host
 * memory; rv_build_device takes T already resident in HBM (left untouched
 * unless rc, in which case the handle works on its own copy). */
int rv_build(rv_index *idx, const uint8_t *T, int64_t n, const int64_t *nsep, int32_t nsamples, int32_t rc);
int rv_build_device(rv_index *idx, const uint8_t *dT, int64_t n, const int64_t *nsep, int32_t nsamples, int32_t rc);
/* construct() with the suffix array (and optionally the LCP array) read from the reference's cache files
 * (index(sa=..., lcp=...), interface.c:224-231,255-262; raw little-endian int32): uploads them, derives the
 * inverse SA (interface.c:235-238), computes the LCP array when LCP == NULL, builds SO. */
int rv_build_cached(rv_index *idx, const uint8_t *T, int64_t n, const int64_t *nsep, int32_t nsamples, int32_t rc, const int32_t *SA,
                    const int32_t *LCP);
int rv_get_times(const rv_index *idx, rv_times *out);

/* Optional per-kernel profile: when enabled the library brackets the launches of its heavy kernels
 * with CUDA events on the build stream and accumulates, per slot, duration, launch count and
 * ALGORITHMIC bytes (stated per slot below and in DESIGN.md) until the next rv_profile(idx, 1). */
enum rv_prof_slot {
    RV_PROF_RADIX_PASS = 0, /* rs_pass_kernel: items x (key + 4 B suffix) x (read + write) */
    RV_PROF_PAIRS = 1,      /* sa_place_kernel: n x (key + 4 B suffix read; SA + SAi + LCP = 12 B written) */
    RV_PROF_SWEEP = 2,      /* pair / multi sweep kernels (count + write): n x 9 B (11 B with SO) per pass */
    RV_PROF_LCP = 3,        /* lcp_sparse_kernel (entries of the suffixes the doubling rounds placed): 13 B per marked position + n/8 */
    RV_PROF_LEAD = 4,       /* sa_lead_kernel: n x (key + 4 B suffix) read, sampled pairs compared and recorded */
    RV_PROF_TEXT = 5,       /* passes over the text: byte histogram, k-mer digit histograms, barrier bitmaps + packed text: n x 1 B each */
    RV_PROF_ROUNDS = 6,     /* stage 4 (refinement / doubling rounds): key gather, group scan and placement kernels, 24 B per active suffix */
    RV_PROF_SLOTS = 8
};
typedef struct rv_kernel_profile {
    double ms[RV_PROF_SLOTS];
    int64_t launches[RV_PROF_SLOTS];
    int64_t bytes[RV_PROF_SLOTS];
    int64_t launches_total; /* every kernel this handle launched since creation */
} rv_kernel_profile;
/* enable: 0 off, 1 every slot, else a mask -- bit k+1 set: slot k is bracketed (the event pairs cost a few microseconds of a
 * step themselves: a timed region brackets only the kernel it reports) */
int rv_profile(rv_index *idx, int32_t enable);
int rv_get_profile(rv_index *idx, rv_kernel_profile *out);
int64_t rv_index_n(const rv_index *idx);

/* Getters: copy an array to host memory. idx_bits 32 -> int32 entries (LCP int32),
 * 64 -> int64 entries (LCP uint32), like reveallib / reveallib64. */
int rv_get_sa(rv_index *idx, void *out, int32_t idx_bits);
int rv_get_sai(rv_index *idx, void *out, int32_t idx_bits);
int rv_get_lcp(rv_index *idx, void *out, int32_t idx_bits);
int rv_get_so(rv_index *idx, uint16_t *out);          /* RV_ERR_STATE when nsamples <= 2 (SO is NULL in the reference) */
int rv_get_text(rv_index *idx, uint8_t *out);         /* T as indexed (after rc) */
/* Overwrites T[begin .. begin+len) on the device: a rank of a sharded recursion hands the stretches of text it marked (the lower-cased
 * matched bases, reveal.c:1230-1234) to the rank that collects the alignment. */
int rv_put_text(rv_index *idx, int64_t begin, const uint8_t *src, int64_t len);
/* device pointers of the resident arrays (int32 / uint16 / uint8), for callers that stay on the GPU */
int rv_device_arrays(rv_index *idx, const uint8_t **dT, const int32_t **dSA, const int32_t **dSAi, const int32_t **dLCP, const uint16_t **dSO);

/* Pair sweep over the root index.  Rows (l, a, b) as int64 triples in ascending
 * SA rank, exactly the reference's list order.  Two-step: *_count runs the
 * sweep and leaves the result on the device, *_fetch copies up to cap rows. */
int rv_mums_pair_count(rv_index *idx, int32_t minl, int32_t flavour, int64_t *count);
int rv_mums_pair_fetch(rv_index *idx, int64_t *rows, int64_t cap);

/* Multi-genome sweep.  CSR result: hdr rows (l, n_members, first_member) and
 * member rows (sample, position), list order = the reference's pop order,
 * members in SA-rank order. */
int rv_mums_multi_count(rv_index *idx, int32_t minl, int32_t minn, int64_t *nrec, int64_t *nmem);
/* getmultimems (reveal.c:292-434 + ismultimem :261-290): lcp-intervals of any size whose members cover >= minn
 * samples; hdr rows (l, n_samples, first_member).  Fetched with rv_mums_multi_fetch.  At most 64 samples. */
int rv_mems_multi_count(rv_index *idx, int32_t minl, int32_t minn, int64_t *nrec, int64_t *nmem);
int rv_mums_multi_fetch(rv_index *idx, int64_t *hdr, int64_t hdr_cap, int64_t *members, int64_t mem_cap);

/* Device pointers of the last sweep result (rows / hdr as int64 triples, members
 * as int64 pairs; NULL when empty), for callers that gather over NCCL. */
int rv_result_device(rv_index *idx, const int64_t **d_rows, int64_t *nrows, const int64_t **d_members, int64_t *nmembers);

/* Packs the last sweep's rows for a fixed-capacity gather: d_dst[0] = row count, d_dst[1] = the number of packs
 * done on this handle so far (a sequence number the collector can check), d_dst[3..] = the first
 * min(count, cap_rows) rows (int64 triples); one kernel, asynchronous on the handle's stream.  d_dst may be local
 * memory or a mapped peer block (rv_peer_open). */
int rv_result_pack_device(rv_index *idx, int64_t *d_dst, int64_t cap_rows);

/* Peer blocks (multi-GPU, one box): device memory that the collecting rank allocates and the other ranks (one
 * process per GPU) map through CUDA IPC, so that rv_result_pack_device can write a rank's rows straight into the
 * collector's HBM over NVLink / NVSwitch -- no collective and no rendezvous on the data path
 * (reveal_b200/shard.py:PeerGather).  No counterpart in the reference, which is single-process (the
 * path shards as independent index builds, the only exchange is the MUM records travelling to the rank that runs
 * the callbacks).  `handle` is RV_PEER_HANDLE_BYTES opaque bytes to hand to the other
 * processes; rv_peer_open maps it (in another process than the allocating one), rv_peer_close unmaps,
 * rv_peer_free releases the allocation, rv_peer_read copies from a peer block to host memory (synchronous). */
#define RV_PEER_HANDLE_BYTES 64
int rv_peer_alloc(int64_t bytes, void **d_ptr, uint8_t *handle);
int rv_peer_open(const uint8_t *handle, void **d_ptr);
int rv_peer_close(void *d_ptr);
int rv_peer_free(void *d_ptr);
int rv_peer_read(const void *d_src, void *host_dst, int64_t bytes);

/* Sweeps over caller-supplied device arrays of a SUB-index that shares the
 * main text (the children of reveal.c:582-664 split; RevealIndex.main): the
 * same kernels, used by the recursion.  dSO may be NULL iff main_nsamples == 2. */
int rv_sweep_pair_device(rv_index *ws, const uint8_t *dT, const int32_t *dSA, const int32_t *dLCP, int64_t n, int64_t nT,
                         int64_t nsep0, int32_t rc, int32_t flavour, int32_t minl, int64_t *count);
int rv_sweep_multi_device(rv_index *ws, const uint8_t *dT, const int32_t *dSA, const int32_t *dLCP, const uint16_t *dSO, int64_t n,
                          int64_t nsep0, int32_t main_nsamples, int32_t minl, int32_t minn, int64_t *nrec, int64_t *nmem);

/* ---- recursion: sub-indexes (the children of reveal.c:582-664 split, RevealIndex.main) ----------
 * A sub-index is a device-resident (SA, LCP) pair of n entries that shares the main index's text,
 * inverse SA and sample array, exactly like the reference's child indexes (reveal.c:1136-1207).
 * rv_sub_root wraps the root arrays; rv_sub_split performs one step of the reference's aligner
 * after its two Python callbacks: label scatter (reveal.c:1005-1117), split (:582-664), lower-casing
 * of the matched bases in T (:1230-1234) and bubble_sort of the leading child (:666-727).
 * Intervals are (begin, end) pairs of text positions, end exclusive; children[0..2] receive the
 * leading, trailing and parallel child (NULL when empty).  The sweeps over a sub-index are
 * getmums_rem (reveal.c:119-180) / getmultimums (:436-580) as the aligner calls them (:802-829);
 * results are fetched with rv_mums_pair_fetch / rv_mums_multi_fetch on the main handle. */
typedef struct rv_sub rv_sub;
int rv_sub_root(rv_index *idx, rv_sub **out);
int64_t rv_sub_n(const rv_sub *sub);
void rv_sub_free(rv_sub *sub);
int rv_sub_get(rv_sub *sub, int32_t which /* 0 SA, 1 LCP */, int32_t *out);
int rv_sub_mums_pair(rv_sub *sub, int32_t minl, int64_t *count);
int rv_sub_mums_multi(rv_sub *sub, int32_t minl, int32_t minn, int64_t *nrec, int64_t *nmem);
/* rv_sub_step = rv_sub_split that may ALSO run the children's MUM sweep (sweep[c] != 0 for the children that
 * carry no precomputed skipmums) in the same launch: parents of at most 16384 suffixes are handled by one
 * thread block end to end (one launch, one synchronisation per recursion step).  A following
 * rv_sub_mums_pair / _multi on such a child answers from the stored result; rv_sub_fetch copies it. */
int rv_sub_step(rv_sub *parent, const int64_t *lead, int32_t nlead, const int64_t *trail, int32_t ntrail, const int64_t *par, int32_t npar,
                const int64_t *mum_sp, int32_t mum_n, int64_t mum_l, const int64_t *matching, int32_t nmatch, const int32_t *sweep, int32_t minl,
                int32_t minn, rv_sub **children);
/* extract (reveal.c:1386-1505): removes the text positions of nintervals (begin, end) pairs from the sub-index in place --
 * SA / LCP compacted (LCP of a survivor = minimum over the removed run before it), inverse rewritten, bases lower-cased,
 * bubble_sort with the intervals.  Slot 0 of the result holds the suffix that belongs there (the reference leaves it
 * uninitialised, reveal.c:1448). */
int rv_sub_extract(rv_sub *sub, const int64_t *intervals, int32_t nintervals);
/* Frontier batching: the sub-indexes waiting on the aligner's queue are independent (reveal.c:1296-1324 pushes up to three per
 * step; interface.c:316-385 hands them to a worker pool), so every step whose callbacks have run can go to the device together:
 * the steps whose parent fits one thread block share ONE launch and ONE synchronisation (one block per step), the others take
 * the general path one after the other.  All parents must belong to one main index; status / children are per step. */
typedef struct rv_step_desc {
    rv_sub *parent;
    const int64_t *lead; int32_t nlead;
    const int64_t *trail; int32_t ntrail;
    const int64_t *par; int32_t npar;
    const int64_t *mum_sp; int32_t mum_n; int64_t mum_l;
    const int64_t *matching; int32_t nmatch;
    int32_t sweep[3];      /* in: run the child's MUM sweep in the same launch (leading, trailing, parallel) */
    rv_sub *children[3];   /* out */
    int32_t status;        /* out: rv_status of this step */
} rv_step_desc;
int rv_sub_step_batch(rv_step_desc *steps, int32_t nsteps, int32_t minl, int32_t minn);
/* The same batch in two halves, so that the caller's host work for the NEXT batch (the two callbacks of every waiting
 * sub-index) overlaps the device part of this one: _begin stages the steps, enqueues the launch and returns a ticket, _end
 * waits for it (an event), fills children / status of every step and frees the ticket.  `steps` stays valid and untouched in
 * between; one open ticket per main index.  rv_sub_step_batch = _begin + _end. */
typedef struct rv_step_batch rv_step_batch;
int rv_sub_step_batch_begin(rv_step_desc *steps, int32_t nsteps, int32_t minl, int32_t minn, rv_step_batch **ticket);
int rv_sub_step_batch_end(rv_step_batch *ticket);
/* steps2[0]/seconds2[0]: recursion steps taken by the single-launch path and their host wall time; [1]: general path */
int rv_rec_stats(rv_index *idx, int64_t *steps2, double *seconds2);
int rv_rec_launches(rv_index *idx, int64_t *launches2); /* launches behind those steps: [0] one per batch, [1] general-path calls */
int rv_sub_fetch(rv_sub *sub, int64_t *rows, int64_t cap_rows, int64_t *members, int64_t cap_members);
int rv_sub_split(rv_sub *parent, const int64_t *lead, int32_t nlead, const int64_t *trail, int32_t ntrail, const int64_t *par, int32_t npar,
                 const int64_t *mum_sp, int32_t mum_n, int64_t mum_l, const int64_t *matching, int32_t nmatch, rv_sub **children);

/* ---- chaining on the device (the step right after MUM extraction) ------------------------------------
 * The O(m^2) chaining recurrence of the mumpicker (reveal/schemes.py:20-104 `chain`, gap cost reveal/utils.py:162-180) for
 * nlists anchor lists in ONE launch, one thread block per list.  List j has rows row_off[j] .. row_off[j+1] (row 0 = the left
 * bound, the last row = the right bound, the anchors in between in processing order), kk[j] coordinates per row stored at
 * start[start_off[j] ...] row-major; length / gain per row; model 0 = sumofpairs, 1 = star-avg, 2 = star-med.  link[r] = the
 * predecessor row (list-local), score[r] = the best total.  Host buffers; works on any handle (its stream and staging area). */
int rv_chain_batch(rv_index *idx, int32_t nlists, const int64_t *row_off, const int64_t *start_off, const int32_t *kk, const int64_t *start,
                   const int64_t *length, const int64_t *gain, int64_t wpen, int32_t model, int64_t *link, int64_t *score);

/* ---- many tiny indexes in one launch ----------------------------------------------------------------
 * `finish` / `transform` extend every anchor by indexing its two <= 200 bp flanks on their own (reveal/transformold.py:1170-1240
 * `extend`: index() / addsequence x 2 / construct() / getmums(minlocallength), four times per anchor).  Unit u is the text
 * T[off[u] .. off[u+1]) = "ref$qry$" exactly as addsequence assembles it (interface.c:71-85), at most 1024 characters, with
 * nsep0[u] = the position of its first '$' inside the unit.  One thread block per unit: suffix array, barrier-aware LCP and the
 * getmums scan (reveal.c:55-116) in shared memory.  rows[u * cap * 3 ...] receives up to cap rows (l, a, b) of unit u in
 * SA-rank order -- what index.getmums(minl) returns for that pair -- and counts[u] the number found (it may exceed cap). */
int rv_mums_tiny_batch(rv_index *idx, const uint8_t *T, const int64_t *off, const int64_t *nsep0, int32_t nunits, int32_t minl, int32_t cap,
                       int64_t *rows, int32_t *counts);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
