// chain_dp.h -- the O(m^2) chaining recurrence of the REM driver (reveal_b200/rem.py:chain; reference: reveal/schemes.py:20-104
// with utils.gapcost), shared by reveallib.chain_dp (called from Python) and remcore.Graph.pick.
//
//   start [rows][k]  coordinates; row 0 = the left bound, rows 1.. = the anchors in processing order (the last one is the
//                    right bound); length / gain [rows]; model 0 = sumofpairs, 1 = star-avg, 2 = star-med; k <= 64.
// Row r may follow every earlier row that ends at or before it in every coordinate.  Its score is the best of
// score[i] + gain[r] - wpen * gapcost(i, r); equal totals go to the predecessor with the higher score, then to the one that
// became available (fitted some row) earlier, then to the one processed earlier -- the order of the reference's `active` list.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

static inline void rv_chain_dp(const int64_t *start, const int64_t *length, const int64_t *gain, long rows, long k, int64_t wpen, int model,
                               int64_t *link, int64_t *score) {
    std::vector<int64_t> joined(rows, -1);
    joined[0] = 0;
    link[0] = 0;
    score[0] = 0;
    int64_t dist[64], tmp[64];
    for (long r = 1; r < rows; r++) {
        const int64_t *sr = start + r * k;
        long best = -1;
        int64_t best_total = 0;
        for (long i = 0; i < r; i++) {
            const int64_t *si = start + i * k;
            bool ok = true;
            for (long c = 0; c < k; c++) {
                int64_t d = sr[c] - (si[c] + length[i]);
                if (d < 0) { ok = false; break; }
                dist[c] = d;
            }
            if (!ok) continue;
            if (joined[i] < 0) joined[i] = r;
            int64_t pen = 0;
            if (model == 0) {
                for (long a = 0; a < k; a++)
                    for (long b = a + 1; b < k; b++) pen += dist[a] > dist[b] ? dist[a] - dist[b] : dist[b] - dist[a];
            } else if (model == 1) {
                int64_t sum = 0;
                for (long c = 0; c < k; c++) sum += dist[c];  // all distances are >= 0 here
                pen = sum / k;
            } else {
                memcpy(tmp, dist, sizeof(int64_t) * k);
                for (long a = 1; a < k; a++) {  // insertion sort, k is the number of samples
                    int64_t v = tmp[a];
                    long b = a;
                    while (b > 0 && tmp[b - 1] > v) { tmp[b] = tmp[b - 1]; b--; }
                    tmp[b] = v;
                }
                pen = tmp[k / 2];
            }
            const int64_t total = score[i] + gain[r] - wpen * pen;
            bool take = best < 0 || total > best_total;
            if (!take && total == best_total) {
                if (score[i] != score[best]) take = score[i] > score[best];
                else if (joined[i] != joined[best]) take = joined[i] < joined[best];
            }
            if (take) { best = i; best_total = total; }
        }
        if (best < 0) { best = 0; best_total = 0; }  // cannot happen with a proper left bound
        link[r] = best;
        score[r] = best_total;
    }
}
