// rv_sweep.h -- arguments of the MUM sweeps (device pointers of one (sub)index).
#pragma once
#include "rv_internal.h"

namespace rv {

struct SweepArgs {
    const unsigned char *T;     // text of the MAIN index (positions in SA index into it)
    const int *SA;              // n entries of this (sub)index
    const int *LCP;             // n entries
    const unsigned short *SO;   // sample of every text position, or NULL (main nsamples == 2)
    i64 n;                      // entries in SA/LCP            (RevealIndex.n,  reveal.h:27)
    i64 nT;                     // length of the main text       (RevealIndex.nT, reveal.h:28)
    i64 nsep0;                  // position of the last '$' of sample 0 (nsep[0])
    int rc;                     // index built over a reverse-complemented second sample
    int flavour;                // 0: getmums (reveal.c:55)  1: getmums_rem (reveal.c:119)
    int minl;
    int minn;
    int main_nsamples;
    // Sample separators carried in the argument block (constant bank of the launch) when there are at most SW_NSEP_INLINE of
    // them: the sample of a position is then a handful of compares, sample(pos) = #{k : nsep[k] < pos} = SO[pos], instead of a
    // gather from the 2-bytes-per-character SO array that misses L2 at genome scale.  0: use SO.
    int nsep_n = 0;
    i64 nsep_v[15];
};
static const int SW_NSEP_INLINE = 15;

size_t sweep_scratch_bytes(i64 n);
int sweep_pair_count(Stream &st, const SweepArgs &p, void *scratch, i64 *count, i64 *d_spec, i64 cap_spec);
int sweep_pair_write(Stream &st, const SweepArgs &p, void *scratch, i64 *d_out, i64 cap);
int sweep_multi_count(Stream &st, const SweepArgs &p, void *scratch, i64 *nrec, i64 *nmem, i64 *d_hdr_spec, i64 hdr_cap_spec, i64 *d_mem_spec,
                      i64 mem_cap_spec);
int sweep_multi_write(Stream &st, const SweepArgs &p, void *scratch, i64 *d_hdr, i64 hdr_cap, i64 *d_mem, i64 mem_cap);

size_t mems_scratch_bytes(i64 n);
int sweep_mems_count(Stream &st, const SweepArgs &p, void *scratch, i64 *nrec, i64 *nmem);
int sweep_mems_write(Stream &st, const SweepArgs &p, void *scratch, i64 *d_hdr, i64 hdr_cap, i64 *d_mem, i64 mem_cap);

}  // namespace rv
