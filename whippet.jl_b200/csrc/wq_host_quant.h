// Host-side quantification state: the count stores downloaded from the device in canonical
// (sorted) order, the annotated isoform paths, and the host halves of the quant stage.
#pragma once
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>
#include "wq_types.h"
#include "wq_kernels.cuh"

// host threads this process may use: the machine's cores shared between the ranks of one node
static inline unsigned wq_host_threads() {
    if (const char* ev = getenv("WQ_HOST_THREADS")) return (unsigned)std::max(1, atoi(ev));
    unsigned hc = std::max(1u, std::thread::hardware_concurrency());
    unsigned ranks = 1;
    if (const char* lw = getenv("LOCAL_WORLD_SIZE")) ranks = (unsigned)std::max(1, atoi(lw));
    return std::max(1u, std::min(32u, hc / ranks));
}

struct WqKeyedStore {                                        // variable-length keyed counts (wq_build_key_* records)
    std::vector<uint32_t> words;                             // key words of all records, back to back
    std::vector<uint64_t> off;                               // [n+1]
    std::vector<uint64_t> count;                             // [n]
    std::vector<double> bias_p, bias_g;                      // [n] or empty: primer-adjusted count and GC factor (< 0: none) with --biascorrect
    WqKeyedStore() : off(1, 0) {}
    size_t size() const { return count.size(); }
    const uint32_t* key(size_t i) const { return words.data() + off[i]; }
    size_t len(size_t i) const { return (size_t)(off[i + 1] - off[i]); }
    void push(const uint32_t* w, size_t n, uint64_t c) {
        words.insert(words.end(), w, w + n);
        off.push_back(words.size());
        count.push_back(c);
    }
    static bool less(const uint32_t* a, size_t na, const uint32_t* b, size_t nb) {
        return std::lexicographical_compare(a, a + na, b, b + nb);
    }
    // sort lexicographically and add up equal keys (canonical order)
    void canonicalize() {
        const size_t n = size();
        std::vector<uint32_t> idx(n);
        for (size_t i = 0; i < n; i++) idx[i] = (uint32_t)i;
        std::sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return less(key(x), len(x), key(y), len(y)); });
        WqKeyedStore out;
        out.words.reserve(words.size());
        const bool fp = !bias_p.empty();                     // one device: every key sits in one slot, so equal keys never meet here
        for (size_t k = 0; k < n; k++) {
            uint32_t i = idx[k];
            if (out.size() && out.len(out.size() - 1) == len(i) &&
                std::equal(key(i), key(i) + len(i), out.key(out.size() - 1))) out.count.back() += count[i];
            else { out.push(key(i), len(i), count[i]); if (fp) { out.bias_p.push_back(bias_p[i]); out.bias_g.push_back(bias_g[i]); } }
        }
        *this = std::move(out);
    }
};

struct WqHostQuant {
    std::vector<uint64_t> node;                              // SpliceGraphQuant.node, by global node index
    std::vector<uint64_t> edge_key, edge_count;              // .edge and .circ, sorted by gene<<32 | l<<16 | r
    WqKeyedStore longs;                                      // .long (any order; order does not affect results)
    WqKeyedStore multis;                                     // MultiMapping.map in canonical (lexicographic) order
    WqCounters stats;
    // --biascorrect (JointBiasCounter, src/bias.jl:207-270): the counts the quant stage works with are Float64.  *_p = the count after
    // primer_adjust! (what build_equivalence_classes! and the EM see), *_g = the factor gc_adjust! multiplies it by afterwards (< 0: the
    // counter saw no GC window and is left alone); same order as node / edge_key.  The integer arrays still hold the raw read tallies.
    bool bias = false;
    std::vector<double> node_p, node_g, edge_p, edge_g;
    WqHostQuant() { memset(&stats, 0, sizeof stats); }
    double nodec(size_t i) const { return bias ? node_p[i] : (double)node[i]; }
    double edgec(size_t i) const { return bias ? edge_p[i] : (double)edge_count[i]; }
    double longc(size_t i) const { return bias ? longs.bias_p[i] : (double)longs.count[i]; }
    double multic(size_t i) const { return bias ? multis.bias_p[i] : (double)multis.count[i]; }
    static double gc_apply(double v, double g) { return g < 0.0 ? v : v * g; }       // cnt.count *= weight (src/bias.jl:264-268)
};

struct WqClassSet {                                         // IsoCompat / MultiCompat in canonical order (multi first)
    std::vector<uint32_t> ptr;                               // [C+1]
    std::vector<int32_t>  iso;                               // 0-based global isoform ids
    std::vector<double>   prop;
    std::vector<double>   count;                             // [C]
    std::vector<uint8_t>  state;                             // [C] 0 active, 1/2 finished (isdone)
    std::vector<uint8_t>  postassign;                        // [C]
    uint32_t n_multi = 0;
};

// State of the quant stage after build_equivalence_classes! / gene_em! / assign_ambig!.
struct WqHostQuantEx {
    bool built = false, em_done = false, ambig_done = false;
    std::vector<double> node_value;                          // node store as Float64 (fractional after assign_ambig!)
    std::vector<std::pair<uint64_t, double>> edge_adds;      // additions to the edge/circ stores, in the order made
    size_t adds_before_gc = 0;                               // --biascorrect: edge_adds[0 .. adds_before_gc) were made before gc_adjust! (the long-path re-count)
    std::vector<uint64_t> edge_key;                          // merged stores (valid after finalize_edges)
    std::vector<double>   edge_value;
    bool edges_final = false;
    std::vector<double> iso_count, tpm, genetpm;
    std::vector<uint8_t> gene_dirty;                         // [G+1] gene is named by some multi-map class: assign_ambig! will change its counts
    WqClassSet cls;
    bool local_built = false;                                // this rank's share of the classes is built (class building partitioned by gene)
    uint32_t own_g_lo = 1, own_g_hi = 0xFFFFFFFFu, own_m_lo = 0, nparts = 1;      // the share: gene range, first multi-map key
    std::vector<uint32_t> frag_cls;                          // class ranges of the class-building fragments (disjoint gene ranges)
    int64_t iterations = 0;
    float em_kernel_ms = 0.f;                                // device time of the transcript EM kernel (CUDA events)
};

// base counts + additions (stable by key, so every key sees its additions in the order they were made).  Keys are
// gene-major, so host threads take contiguous key ranges: each sorts the additions that fall into its range, merges them
// with its slice of the base edges, and the slices are concatenated.
static inline void wq_finalize_edges(const WqHostQuant& hq, WqHostQuantEx& X) {
    const size_t ne = hq.edge_key.size(), na = X.edge_adds.size();
    unsigned nt = wq_host_threads();
    if (ne + na < (1u << 16)) nt = 1;
    // split points: equal shares of the base edges, moved to a gene boundary so no key straddles two ranges
    std::vector<uint64_t> cut(nt + 1, 0);
    cut[nt] = ~0ull;
    for (unsigned t = 1; t < nt; t++) {
        const size_t i = ne * t / nt;
        cut[t] = i < ne ? (hq.edge_key[i] >> 32) << 32 : ~0ull;
        if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    }
    struct Part { std::vector<std::pair<uint64_t, double>> adds; std::vector<size_t> seq; std::vector<uint64_t> key; std::vector<double> value; };
    std::vector<Part> parts(nt);
    auto work = [&](unsigned t) {
        Part& P = parts[t];
        const uint64_t lo = cut[t], hi = cut[t + 1];
        if (!hq.bias) {
            for (const auto& a : X.edge_adds) if (a.first >= lo && a.first < hi) P.adds.push_back(a);
            std::stable_sort(P.adds.begin(), P.adds.end(), [](const std::pair<uint64_t, double>& a, const std::pair<uint64_t, double>& b) { return a.first < b.first; });
        } else {                                             // the same, remembering every addition's position in the order made
            std::vector<size_t> ord;
            for (size_t q = 0; q < X.edge_adds.size(); q++) if (X.edge_adds[q].first >= lo && X.edge_adds[q].first < hi) ord.push_back(q);
            std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return X.edge_adds[a].first < X.edge_adds[b].first; });
            for (size_t q : ord) { P.adds.push_back(X.edge_adds[q]); P.seq.push_back(q); }
        }
        size_t i = std::lower_bound(hq.edge_key.begin(), hq.edge_key.end(), lo) - hq.edge_key.begin();
        const size_t ie = hi == ~0ull ? ne : (size_t)(std::lower_bound(hq.edge_key.begin(), hq.edge_key.end(), hi) - hq.edge_key.begin());
        size_t j = 0;
        const size_t ja = P.adds.size();
        P.key.reserve(ie - i + ja); P.value.reserve(ie - i + ja);
        while (i < ie || j < ja) {
            uint64_t k;
            double v = 0.0;
            double g = -1.0;                                 // gc_adjust! acts on counters that exist when it runs: the base edges
            if (j >= ja || (i < ie && hq.edge_key[i] <= P.adds[j].first)) { k = hq.edge_key[i]; v = hq.edgec(i); if (hq.bias) g = hq.edge_g[i]; i++; }
            else k = P.adds[j].first;
            // with bias correction the additions made before gc_adjust! (long paths re-counted onto their edges by
            // build_equivalence_classes!) are inside the count it scales; the later ones (assign_ambig!) are not
            if (hq.bias) {
                while (j < ja && P.adds[j].first == k && P.seq[j] < X.adds_before_gc) { v += P.adds[j].second; j++; }
                v = WqHostQuant::gc_apply(v, g);
            }
            while (j < ja && P.adds[j].first == k) { v += P.adds[j].second; j++; }
            P.key.push_back(k);
            P.value.push_back(v);
        }
    };
    if (nt == 1) work(0);
    else { std::vector<std::thread> th; for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t); for (auto& x : th) x.join(); }
    std::vector<size_t> at(nt + 1, 0);
    for (unsigned t = 0; t < nt; t++) at[t + 1] = at[t] + parts[t].key.size();
    X.edge_key.resize(at[nt]); X.edge_value.resize(at[nt]);
    auto place = [&](unsigned t) {
        std::copy(parts[t].key.begin(), parts[t].key.end(), X.edge_key.begin() + at[t]);
        std::copy(parts[t].value.begin(), parts[t].value.end(), X.edge_value.begin() + at[t]);
    };
    if (nt == 1) place(0);
    else { std::vector<std::thread> th; for (unsigned t = 0; t < nt; t++) th.emplace_back(place, t); for (auto& x : th) x.join(); }
    X.edges_final = true;
}

static inline bool wq_edge_is_circ(uint64_t key) { return ((key >> 16) & 0xFFFF) >= (key & 0xFFFF); }
static inline bool wq_key_is_multi(const uint32_t* k) { return (k[0] >> 28) & 1; }

struct WqIsoIndex {                                          // SpliceGraph.annopath of every gene
    std::vector<uint32_t> iso_base;                          // [G+1]
    std::vector<uint32_t> node_ptr;                          // [I+1]
    std::vector<uint32_t> nodes;                             // ascending 1-based node ids
    std::vector<int64_t>  length;                            // GraphLibQuant.length (src/quant.jl:190-192)
    uint32_t I = 0;
};

static inline std::string wq_build_iso_index(const wq_index_desc* d, WqIsoIndex& S) {
    const uint32_t G = d->n_genes, I = d->n_isoforms;
    S.I = I;
    S.iso_base.assign(d->iso_base, d->iso_base + G + 1);
    S.node_ptr.assign(d->iso_node_ptr, d->iso_node_ptr + I + 1);
    S.nodes.assign(d->iso_nodes, d->iso_nodes + S.node_ptr[I]);
    S.length.assign(I, 0);
    if (S.iso_base[G] != I) return "iso_base does not cover n_isoforms";
    for (uint32_t g = 0; g < G; g++) {
        uint32_t nn = d->node_base[g + 1] - d->node_base[g];
        for (uint32_t i = S.iso_base[g]; i < S.iso_base[g + 1]; i++) {
            int64_t len = 0;
            for (uint32_t k = S.node_ptr[i]; k < S.node_ptr[i + 1]; k++) {
                uint32_t nd = S.nodes[k];
                if (nd < 1 || nd > nn) return "annotated path names a node outside its gene";
                if (k > S.node_ptr[i] && nd <= S.nodes[k - 1]) return "annotated path is not strictly ascending";
                len += d->nodelen[d->node_base[g] + nd - 1];
            }
            S.length[i] = len;
        }
    }
    return "";
}
