// Hybrid-query policy: should a filtered top-k be answered ad hoc — score exactly the `subsetSize`
// labels that passed the filter (VecSimIndex_AdhocBfCtx_GetExactDistances: one batched gather on the device) — or in
// batches (VecSimBatchIterator, intersected with the filter until k results passed)?
//
// The reference answers with decision trees fitted to CPU timings (scripts/BF_batches_clf.py / HNSW_batches_clf.py →
// algorithms/brute_force/brute_force.h:380-451, algorithms/hnsw/hnsw.h:2340-2400). Their inputs are the same here
// (index size, dim, subset size, k) but the costs are not: on the device an ad-hoc pass gathers subset * row bytes at the
// re-rank kernel's rate, a flat batch pass streams the whole store once and then moves one score per row to the host,
// and an HNSW batch costs one latency-bound hop per result. So the rule is a cost comparison with the rates measured on
// B200 this round (profiles/r1_gather_v4_ncu.md, r1_scan_tma_b1_ncu.md, r1_hnsw_cfg5_1M.json).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>

namespace vsb {

struct HybridCosts {
    double launch_s = 30e-6;        // one query's fixed device + API cost
    double bw_gather = 2.9e12;      // B/s, exact_gather_kernel on scattered rows
    double bw_scan = 6.5e12;        // B/s, exact scan at one query per pass
    double bw_pcie = 25e9;          // B/s, pinned D2H
    double host_label_s = 30e-9;    // label -> id lookup + result write per ad-hoc label
    double host_select_s = 2e-9;    // per score per flat batch (nth_element + sort share)
    double hnsw_result_s = 8e-6;    // one traversal hop (~ one result) of the HNSW batch iterator, 6-11 us measured
    double expected_flat_batches = 2.0;
};

// flat index: ad hoc unless the subset is nearly the whole index AND few batches are expected
inline bool prefer_adhoc_flat(size_t index_size, size_t row_bytes, size_t subset, size_t k, const HybridCosts &c = HybridCosts()) {
    (void)k;
    subset = std::min(subset, index_size);
    if (index_size == 0) return true;
    const double n = (double)index_size, s = (double)subset, rb = (double)row_bytes;
    const double adhoc = c.launch_s + s * (rb / c.bw_gather + c.host_label_s + 8.0 / c.bw_pcie);
    const double batches = c.launch_s + n * rb / c.bw_scan + 8.0 * n / c.bw_pcie + c.expected_flat_batches * n * c.host_select_s;
    return adhoc <= batches;
}

// HNSW: a batch pass has to surface ~k / r results (r = subset / index size) before k of them pass the filter
inline bool prefer_adhoc_hnsw(size_t index_size, size_t row_bytes, size_t subset, size_t k, const HybridCosts &c = HybridCosts()) {
    subset = std::min(subset, index_size);
    if (index_size == 0 || subset == 0) return true;
    const double n = (double)index_size, s = (double)subset, rb = (double)row_bytes;
    const double adhoc = c.launch_s + s * (rb / c.bw_gather + c.host_label_s + 8.0 / c.bw_pcie);
    const double needed = std::min(n, (double)std::max<size_t>(k, 1) * n / s);
    const double batches = c.launch_s + needed * c.hnsw_result_s;
    return adhoc <= batches;
}

} // namespace vsb
