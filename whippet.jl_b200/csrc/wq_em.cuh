// Quantification stage: isoform compatibility classes (host), transcript-level EM (one persistent
// cooperative kernel), multi-map assignment (host), batched PSI EM (one event per block).
//
//   wq_host_build_classes   build_equivalence_classes!  src/quant.jl:327-389  + MultiCompat :394-435
//   wq_k_em_tpm             calculate_tpm! + gene_em!   src/quant.jl:655-667, 702-769
//   wq_host_assign_ambig    assign_ambig!               src/quant.jl:520-553
//   wq_k_em_psi             spliced_em! / rec_tandem_em!  src/events.jl:831-932
//
// Canonical orders (Julia Dict order is not reproducible): multi-map classes
// and long paths by key, isoform classes per gene by id vector; the per-isoform accumulation of a
// class sweep is done as a gather in exactly that order, so the floating-point adds happen in the
// same sequence a sequential sweep would perform them.  sum(tpm) is a fixed adjacent-pair tree.
#pragma once
#include <cooperative_groups.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <thread>
#include <vector>
#include "wq_host_quant.h"
#include "wq_host_index.h"

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------- host
struct WqGeneBits {                 // annopath bitsets of one gene
    uint32_t np = 0, words = 0;
    std::vector<uint64_t> bits;     // [np][words], bit n = node n (1-based)
    void build(const WqIsoIndex& S, uint32_t g, uint32_t nn) {
        np = S.iso_base[g + 1] - S.iso_base[g];
        words = (nn + 64) / 64;
        bits.assign((size_t)np * words, 0);
        for (uint32_t p = 0; p < np; p++) {
            uint32_t i = S.iso_base[g] + p;
            for (uint32_t k = S.node_ptr[i]; k < S.node_ptr[i + 1]; k++) bits[(size_t)p * words + (S.nodes[k] >> 6)] |= 1ull << (S.nodes[k] & 63);
        }
    }
    bool has(uint32_t p, uint32_t n) const { return (n >> 6) < words && ((bits[(size_t)p * words + (n >> 6)] >> (n & 63)) & 1); }
    // both ends annotated and no annotated node strictly between (src/quant.jl:68-79)
    bool adjacent(uint32_t p, uint32_t a, uint32_t b) const {
        if (!has(p, a) || !has(p, b)) return false;
        for (uint32_t i = a + 1; i < b; i++) if (has(p, i)) return false;
        return true;
    }
    bool path(uint32_t p, const uint32_t* nd, uint32_t n) const {
        if (n == 1) return has(p, nd[0]);
        for (uint32_t i = 0; i + 1 < n; i++) if (!adjacent(p, nd[i], nd[i + 1])) return false;
        return true;
    }
};

template <class F> static inline void wq_on_threads(unsigned nt, F&& f) {
    if (nt <= 1) { f(0u); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back([&f, t]() { f(t); });
    for (auto& x : th) x.join();
}

struct WqKeyPath { uint32_t gene, n, rn; const uint32_t* nd; const uint32_t* rnd; };
static inline size_t wq_key_next(const uint32_t* k, size_t pos, WqKeyPath& o) {
    uint32_t hdr = k[pos];
    o.gene = k[pos + 1];
    if ((hdr >> 28) & 2) { o.n = hdr & 0xFF; o.rn = (hdr >> 8) & 0xFF; o.nd = &k[pos + 2]; o.rnd = o.nd + o.n; return pos + 2 + o.n + o.rn; }
    o.n = hdr & 0xFFFF; o.rn = 0; o.nd = &k[pos + 2]; o.rnd = nullptr;
    return pos + 2 + o.n;
}
static inline bool wq_key_compatible(const WqGeneBits& B, uint32_t p, const WqKeyPath& kp) {
    return B.path(p, kp.nd, kp.n) && (!kp.rnd || B.path(p, kp.rnd, kp.rn));
}

// add `v` along a path: one node -> node store, otherwise every consecutive pair -> edge/circ store
// (assign_path!, src/quant.jl:263-324; the paired form skips mate-2 pieces already covered by mate 1)
// `stale`: bit i set = mate-2 node i counts as already seen on entry (assign_ambig! passes the paired method a set that still holds
// the global isoform ids of the class's last alignment, src/quant.jl:493-501, 541-551; a mate-2 node id equal to one of them is skipped)
static inline void wq_spread_path(const WqHostIndex& H, std::vector<double>& node_value, std::vector<std::pair<uint64_t, double>>& edge_adds,
                                  const WqKeyPath& kp, double v, uint32_t stale = 0) {
    const uint32_t nb = H.genes[kp.gene - 1].node_base;
    auto ekey = [&](uint32_t l, uint32_t r) { return ((uint64_t)kp.gene << 32) | ((uint64_t)(l & 0xFFFF) << 16) | (r & 0xFFFF); };
    uint32_t seen[2 * WQ_MAX_PATH + 2];
    int ns = 0;
    if (kp.n == 1) { node_value[nb + kp.nd[0] - 1] += v; seen[ns++] = kp.nd[0]; }
    else for (uint32_t i = 0; i + 1 < kp.n; i++) {
        if (ns + 2 <= (int)(sizeof seen / sizeof seen[0])) { seen[ns++] = kp.nd[i]; seen[ns++] = kp.nd[i + 1]; }
        edge_adds.emplace_back(ekey(kp.nd[i], kp.nd[i + 1]), v);
    }
    if (!kp.rnd) return;
    auto in_fwd = [&](uint32_t x) { for (int i = 0; i < ns; i++) if (seen[i] == x) return true; return false; };
    auto in_seen = [&](uint32_t i) { return ((stale >> i) & 1u) != 0 || in_fwd(kp.rnd[i]); };
    if (kp.rn == 1) { if (!in_seen(0)) node_value[nb + kp.rnd[0] - 1] += v; }
    else for (uint32_t i = 0; i + 1 < kp.rn; i++) {
        if (in_seen(i) && in_seen(i + 1)) continue;
        edge_adds.emplace_back(ekey(kp.rnd[i], kp.rnd[i + 1]), v);
    }
}

// Builds multi-map classes (first) and isoform classes, fills the unique isoform counts, and
// re-counts long paths onto their edges (assign_long=true).
// With nparts > 1 only the multi-map classes of key range `part` and the isoform classes of gene range `part` are built
// (the rank's share; wq_classes_pack / wq_classes_assemble exchange the shares), and long paths are re-counted onto the
// edges of that gene range only.
static inline std::string wq_host_build_classes(const WqHostIndex& H, const WqIsoIndex& S, WqHostQuant& hq, WqHostQuantEx& X,
                                                uint32_t part = 0, uint32_t nparts = 1) {
    const uint32_t G = H.G, I = S.I;
    static const bool prof2 = getenv("WQ_PROFILE") && atoi(getenv("WQ_PROFILE")) >= 2;
    const auto tc0 = std::chrono::steady_clock::now();
    auto lap2 = [&](const char* what) { if (prof2) fprintf(stderr, "[wq]     classes: %-24s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count()); };
    X = WqHostQuantEx();
    if (!hq.bias) X.node_value.assign(hq.node.begin(), hq.node.end());
    else {
        // --biascorrect: the store starts at the primer-adjusted counts, takes the long-path re-count below (a paired "long path" may
        // be two single-node mates, which add to nodes), and is scaled by gc_adjust! at the end of this function: nothing reads it
        // before assign_ambig! / the events.  Long paths must be visited in canonical (key) order now: their counts are no longer
        // integers, so the order of the additions shows in the result.
        X.node_value = hq.node_p;
        hq.longs.canonicalize();
    }
    X.iso_count.assign(I, 0.0);
    WqClassSet& C = X.cls;
    // ---- genes named by multi-map classes (assign_ambig! will change their counts) ----
    X.gene_dirty.assign(G + 1, 0);
    for (size_t m = 0; m < hq.multis.size(); m++) {
        const uint32_t* key = hq.multis.key(m);
        const uint32_t nh = key[0] & 0xFFFF;
        size_t pos = 1;
        for (uint32_t h = 0; h < nh; h++) {
            WqKeyPath kp;
            pos = wq_key_next(key, pos, kp);
            if (kp.gene < 1 || kp.gene > G) return "multi-map record names a gene out of range";
            X.gene_dirty[kp.gene] = 1;
        }
    }
    // ---- multi-map classes in key order: independent of each other, so host threads take contiguous ranges ----
    struct MFrag { std::vector<int32_t> iso; std::vector<uint32_t> end; std::vector<uint8_t> all; std::string err; };
    auto multi_work = [&](size_t m_lo, size_t m_hi, MFrag& F) {
        WqGeneBits B;
        uint32_t cached_gene = 0;
        std::vector<int32_t> ids;
        for (size_t m = m_lo; m < m_hi; m++) {
            const uint32_t* key = hq.multis.key(m);
            const uint32_t nh = key[0] & 0xFFFF;
            size_t pos = 1;
            ids.clear();
            bool all = true;
            for (uint32_t h = 0; h < nh && all; h++) {
                WqKeyPath kp;
                pos = wq_key_next(key, pos, kp);
                if (cached_gene != kp.gene) { B.build(S, kp.gene - 1, H.genes[kp.gene - 1].nn); cached_gene = kp.gene; }
                bool any = false;
                for (uint32_t p = 0; p < B.np; p++)
                    if (wq_key_compatible(B, p, kp)) { ids.push_back((int32_t)(S.iso_base[kp.gene - 1] + p)); any = true; }
                if (!any) all = false;
            }
            if (!all) ids.clear();
            std::sort(ids.begin(), ids.end());
            ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
            if (ids.size() > 500) { F.err = "multi-map class larger than 500 isoforms (prop_temp = zeros(500), src/quant.jl:725)"; return; }
            F.iso.insert(F.iso.end(), ids.begin(), ids.end());
            F.end.push_back((uint32_t)F.iso.size());
            F.all.push_back(all ? 1 : 0);
        }
    };
    // ---- long paths grouped by gene (the order inside a gene does not affect any result) ----
    std::vector<uint32_t> lptr(G + 2, 0), lidx(hq.longs.size());
    for (size_t i = 0; i < hq.longs.size(); i++) {
        uint32_t g = hq.longs.key(i)[1];
        if (g < 1 || g > G) return "long-path record names a gene out of range";
        lptr[g + 1]++;
    }
    for (uint32_t g = 0; g <= G; g++) lptr[g + 1] += lptr[g];
    {
        std::vector<uint32_t> cur(lptr.begin(), lptr.end() - 1);
        for (size_t i = 0; i < hq.longs.size(); i++) lidx[cur[hq.longs.key(i)[1]]++] = (uint32_t)i;
    }
    // ---- isoform classes, gene by gene; genes are independent, so host threads take contiguous gene ranges and
    //      the fragments are concatenated in gene order ----
    lap2("marks + long-path grouping");
    struct Frag { WqClassSet C; std::vector<std::pair<uint64_t, double>> adds; std::string err; };
    unsigned nthreads = wq_host_threads();
    if (G < 2048) nthreads = 1;
    std::vector<Frag> frags(nthreads);
    std::vector<MFrag> mfrags(nthreads);
    const size_t NMU_all = hq.multis.size();
    const size_t M_lo = NMU_all * part / nparts, NMU = NMU_all * (part + 1) / nparts - M_lo;          // this part's multi-map keys
    const uint32_t Gp_lo = (uint32_t)((uint64_t)G * part / nparts), Gp_n = (uint32_t)((uint64_t)G * (part + 1) / nparts) - Gp_lo;   // and genes
    X.own_g_lo = Gp_lo + 1; X.own_g_hi = Gp_lo + Gp_n; X.own_m_lo = (uint32_t)M_lo; X.nparts = nparts;
    auto work = [&](unsigned t) {
        multi_work(M_lo + NMU * t / nthreads, M_lo + NMU * (t + 1) / nthreads, mfrags[t]);
        Frag& F = frags[t];
        const uint32_t g_lo = Gp_lo + 1 + (uint32_t)((uint64_t)Gp_n * t / nthreads), g_hi = Gp_lo + (uint32_t)((uint64_t)Gp_n * (t + 1) / nthreads);
        WqGeneBits Bt;
        WqHostQuantEx local;                     // only edge_adds / node_value targets are used through `sink`
        std::vector<int32_t> setbuf;             // id lists of this gene's count entries, back to back
        struct Rec { uint32_t off, len; double value; };
        std::vector<Rec> recs;
        std::vector<uint32_t> order;
        const size_t ne = hq.edge_key.size();
        size_t e = std::lower_bound(hq.edge_key.begin(), hq.edge_key.end(), (uint64_t)g_lo << 32) - hq.edge_key.begin();
        for (uint32_t g = g_lo; g <= g_hi; g++) {
            Bt.build(S, g - 1, H.genes[g - 1].nn);
            const WqGeneBits& b = Bt;
            const uint32_t nn = H.genes[g - 1].nn, nb = H.genes[g - 1].node_base, ib = S.iso_base[g - 1];
            setbuf.clear();
            recs.clear();
            auto account = [&](uint32_t off, double value) {
                uint32_t len = (uint32_t)setbuf.size() - off;
                if (len == 1) { X.iso_count[ib + setbuf[off] - 1] += value; setbuf.resize(off); }
                else if (len > 1) recs.push_back(Rec{off, len, value});
            };
            for (uint32_t n = 1; n <= nn; n++) {
                uint32_t off = (uint32_t)setbuf.size();
                for (uint32_t p = 0; p < b.np; p++) if (b.has(p, n)) setbuf.push_back((int32_t)p + 1);
                account(off, hq.nodec(nb + n - 1));
            }
            while (e < ne && (hq.edge_key[e] >> 32) < g) e++;
            for (; e < ne && (hq.edge_key[e] >> 32) == g; e++) {
                if (wq_edge_is_circ(hq.edge_key[e])) continue;
                uint32_t l = (hq.edge_key[e] >> 16) & 0xFFFF, r = hq.edge_key[e] & 0xFFFF;
                uint32_t off = (uint32_t)setbuf.size();
                for (uint32_t p = 0; p < b.np; p++) if (b.adjacent(p, l, r)) setbuf.push_back((int32_t)p + 1);
                account(off, hq.edgec(e));
            }
            for (uint32_t k = lptr[g]; k < lptr[g + 1]; k++) {
                WqKeyPath kp;
                wq_key_next(hq.longs.key(lidx[k]), 0, kp);
                uint32_t off = (uint32_t)setbuf.size();
                for (uint32_t p = 0; p < b.np; p++) if (wq_key_compatible(b, p, kp)) setbuf.push_back((int32_t)p + 1);
                account(off, hq.longc(lidx[k]));
                wq_spread_path(H, X.node_value, F.adds, kp, hq.longc(lidx[k]));
            }
            // distinct id sets in lexicographic order, values added up in the order met (nodes, edges, long paths: the order the
            // reference's temp[arr] += value sees them in, src/quant.jl:333-343; it only shows with --biascorrect, where they are not integers)
            order.resize(recs.size());
            for (uint32_t i = 0; i < recs.size(); i++) order[i] = i;
            auto lt = [&](uint32_t x, uint32_t y) {
                return std::lexicographical_compare(setbuf.begin() + recs[x].off, setbuf.begin() + recs[x].off + recs[x].len,
                                                    setbuf.begin() + recs[y].off, setbuf.begin() + recs[y].off + recs[y].len);
            };
            std::stable_sort(order.begin(), order.end(), lt);
            for (size_t i = 0; i < order.size();) {
                size_t j = i;
                double total = 0.0;
                while (j < order.size() && !lt(order[i], order[j]) && !lt(order[j], order[i])) { total += recs[order[j]].value; j++; }
                const Rec& r0 = recs[order[i]];
                if (total > 0.0) {
                    if (r0.len > 500) { F.err = "isoform class larger than 500 (prop_temp = zeros(500), src/quant.jl:725)"; return; }
                    for (uint32_t k = 0; k < r0.len; k++) { F.C.iso.push_back((int32_t)(ib + setbuf[r0.off + k] - 1)); F.C.prop.push_back(1.0 / (double)r0.len); }
                    F.C.ptr.push_back((uint32_t)F.C.iso.size());
                    F.C.count.push_back(total);
                }
                i = j;
            }
        }
    };
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthreads; t++) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    lap2("threads done");
    // concatenate: multi-map fragments first (key order), then the isoform-class fragments (gene order); sizes -> prefix
    // sums -> every host thread places its own fragment
    for (MFrag& F : mfrags) if (!F.err.empty()) return F.err;
    for (Frag& F : frags) if (!F.err.empty()) return F.err;
    std::vector<size_t> mc_at(nthreads + 1, 0), mm_at(nthreads + 1, 0), fc_at(nthreads + 1, 0), fm_at(nthreads + 1, 0), fa_at(nthreads + 1, 0);
    for (unsigned t = 0; t < nthreads; t++) { mc_at[t + 1] = mc_at[t] + mfrags[t].end.size(); mm_at[t + 1] = mm_at[t] + mfrags[t].iso.size(); }
    const size_t nmc = mc_at[nthreads], nmm = mm_at[nthreads];
    for (unsigned t = 0; t < nthreads; t++) {
        fc_at[t + 1] = fc_at[t] + frags[t].C.count.size(); fm_at[t + 1] = fm_at[t] + frags[t].C.iso.size(); fa_at[t + 1] = fa_at[t] + frags[t].adds.size();
    }
    const size_t ncls = nmc + fc_at[nthreads], nmem = nmm + fm_at[nthreads];
    C.iso.resize(nmem); C.prop.resize(nmem); C.ptr.resize(ncls + 1); C.count.resize(ncls); C.state.resize(ncls); C.postassign.resize(ncls);
    C.ptr[0] = 0;
    C.n_multi = (uint32_t)nmc;
    X.edge_adds.resize(fa_at[nthreads]);
    X.frag_cls.assign(nthreads + 1, 0);
    for (unsigned t = 0; t <= nthreads; t++) X.frag_cls[t] = (uint32_t)(nmc + fc_at[t]);
    wq_on_threads(nthreads, [&](unsigned t) {
        {   // multi-map classes of this thread's key range
            const MFrag& F = mfrags[t];
            uint32_t prev = 0;
            size_t c = mc_at[t], q0 = mm_at[t];
            for (size_t j = 0; j < F.end.size(); j++, c++) {
                const uint32_t n = F.end[j] - prev;
                for (uint32_t q = prev; q < F.end[j]; q++) { C.iso[q0 + q] = F.iso[q]; C.prop[q0 + q] = 1.0 / (double)n; }
                prev = F.end[j];
                C.ptr[c + 1] = (uint32_t)(q0 + F.end[j]);
                C.count[c] = hq.multic(M_lo + c);
                C.state[c] = F.all[j] ? 0 : 2;
                C.postassign[c] = F.all[j] ? 0 : 1;
            }
        }
        {   // isoform classes of this thread's gene range
            const Frag& F = frags[t];
            const size_t c0 = nmc + fc_at[t], q0 = nmm + fm_at[t];
            std::copy(F.C.iso.begin(), F.C.iso.end(), C.iso.begin() + q0);
            std::copy(F.C.prop.begin(), F.C.prop.end(), C.prop.begin() + q0);
            for (size_t j = 0; j < F.C.ptr.size(); j++) C.ptr[c0 + j + 1] = (uint32_t)(q0 + F.C.ptr[j]);
            std::copy(F.C.count.begin(), F.C.count.end(), C.count.begin() + c0);
            std::fill(C.state.begin() + c0, C.state.begin() + c0 + F.C.count.size(), (uint8_t)0);
            std::fill(C.postassign.begin() + c0, C.postassign.begin() + c0 + F.C.count.size(), (uint8_t)0);
            std::copy(F.adds.begin(), F.adds.end(), X.edge_adds.begin() + fa_at[t]);
        }
    });
    lap2("fragments concatenated");
    if (nparts > 1) {
        // assign_long for the genes whose classes another rank builds: their edge values must be complete here too (see assign_ambig);
        // a key belongs to one gene, so the order of the additions to any one edge is the order inside that gene's long paths
        for (uint32_t g = 1; g <= G; g++) {
            if (g >= X.own_g_lo && g <= X.own_g_hi) continue;
            for (uint32_t k = lptr[g]; k < lptr[g + 1]; k++) {
                WqKeyPath kp;
                wq_key_next(hq.longs.key(lidx[k]), 0, kp);
                wq_spread_path(H, X.node_value, X.edge_adds, kp, hq.longc(lidx[k]));
            }
        }
    }
    X.adds_before_gc = X.edge_adds.size();
    if (hq.bias)                                             // gc_adjust!(quant, mod) on the node counters (the edge counters: wq_finalize_edges)
        for (size_t i = 0; i < X.node_value.size(); i++) X.node_value[i] = WqHostQuant::gc_apply(X.node_value[i], hq.node_g[i]);
    X.built = nparts == 1;                   // a share is not a class set: wq_classes_assemble completes it
    X.local_built = true;
    return "";
}

// ---- exchange of class shares between ranks (class building partitioned by gene) ----
// Buffer: u64 hdr[8] = {n_multi, multi_members, n_iso_classes, iso_members, iso_lo, iso_hi, own_m_lo, 0}, then
// u32 ptr[n_cls + 1] (padded to 8 bytes), i32 iso[n_mem] (padded), f64 count[n_cls], u8 state[n_cls], u8 postassign[n_cls]
// (padded), f64 iso_count[iso_hi - iso_lo].
static inline void wq_classes_pack_impl(const WqIsoIndex& S, const WqHostQuantEx& X, std::vector<uint8_t>& out) {
    const WqClassSet& C = X.cls;
    const uint64_t ncls = C.count.size(), nmem = C.iso.size();
    const uint64_t iso_lo = S.iso_base[X.own_g_lo - 1], iso_hi = S.iso_base[X.own_g_hi];
    auto pad8 = [](uint64_t b) { return (b + 7) & ~7ull; };
    const uint64_t bytes = 64 + pad8(4 * (ncls + 1)) + pad8(4 * nmem) + 8 * ncls + pad8(2 * ncls) + 8 * (iso_hi - iso_lo);
    out.assign(bytes, 0);
    uint8_t* p = out.data();
    const uint64_t hdr[8] = {C.n_multi, C.n_multi ? C.ptr[C.n_multi] : 0, ncls - C.n_multi, nmem - (C.n_multi ? C.ptr[C.n_multi] : 0), iso_lo, iso_hi, X.own_m_lo, 0};
    memcpy(p, hdr, 64); p += 64;
    memcpy(p, C.ptr.data(), 4 * (ncls + 1)); p += pad8(4 * (ncls + 1));
    memcpy(p, C.iso.data(), 4 * nmem); p += pad8(4 * nmem);
    memcpy(p, C.count.data(), 8 * ncls); p += 8 * ncls;
    memcpy(p, C.state.data(), ncls); memcpy(p + ncls, C.postassign.data(), ncls); p += pad8(2 * ncls);
    memcpy(p, X.iso_count.data() + iso_lo, 8 * (iso_hi - iso_lo));
}

// All ranks' shares, in rank order -> the complete class set (multi-map classes of every rank first, then the isoform
// classes: the canonical order) and the unique isoform counts.  The node / edge values stay those of this rank's genes.
static inline std::string wq_classes_assemble_impl(const WqIsoIndex& S, WqHostQuantEx& X, const void* const* bufs, const uint64_t* lens, uint32_t n) {
    if (!X.local_built) return "wq_classes_assemble needs wq_classes_build_local first";
    auto pad8 = [](uint64_t b) { return (b + 7) & ~7ull; };
    struct View { uint64_t nm, nmm, nc, ncm, iso_lo, iso_hi; const uint32_t* ptr; const int32_t* iso; const double* count; const uint8_t *state, *post; const double* ic; };
    std::vector<View> V(n);
    uint64_t tm = 0, tmm = 0, tc = 0, tcm = 0;
    for (uint32_t r = 0; r < n; r++) {
        if (!bufs[r] || lens[r] < 64) return "class share buffer too short";
        const uint8_t* p = (const uint8_t*)bufs[r];
        uint64_t hdr[8];
        memcpy(hdr, p, 64); p += 64;
        View& v = V[r];
        v.nm = hdr[0]; v.nmm = hdr[1]; v.nc = hdr[2]; v.ncm = hdr[3]; v.iso_lo = hdr[4]; v.iso_hi = hdr[5];
        const uint64_t ncls = v.nm + v.nc, nmem = v.nmm + v.ncm;
        if (v.iso_hi < v.iso_lo || v.iso_hi > S.I) return "class share names isoforms out of range";
        const uint64_t need = 64 + pad8(4 * (ncls + 1)) + pad8(4 * nmem) + 8 * ncls + pad8(2 * ncls) + 8 * (v.iso_hi - v.iso_lo);
        if (lens[r] < need) return "class share buffer truncated";
        v.ptr = (const uint32_t*)p; p += pad8(4 * (ncls + 1));
        v.iso = (const int32_t*)p; p += pad8(4 * nmem);
        v.count = (const double*)p; p += 8 * ncls;
        v.state = p; v.post = p + ncls; p += pad8(2 * ncls);
        v.ic = (const double*)p;
        tm += v.nm; tmm += v.nmm; tc += v.nc; tcm += v.ncm;
    }
    WqClassSet C;
    C.ptr.resize(tm + tc + 1); C.iso.resize(tmm + tcm); C.prop.resize(tmm + tcm); C.count.resize(tm + tc); C.state.resize(tm + tc); C.postassign.resize(tm + tc);
    C.ptr[0] = 0;
    C.n_multi = (uint32_t)tm;
    X.frag_cls.assign(n + 1, 0);
    X.iso_count.assign(S.I, 0.0);
    // class / member offsets of every (pass, share): multi-map classes of all shares first, then the isoform classes;
    // the shares are then placed by the host threads (one task per pass and share)
    std::vector<uint64_t> c_at(2 * (size_t)n + 1, 0), q_at(2 * (size_t)n + 1, 0);
    for (int pass = 0; pass < 2; pass++)
        for (uint32_t r = 0; r < n; r++) {
            const size_t t = (size_t)pass * n + r;
            c_at[t + 1] = c_at[t] + (pass == 0 ? V[r].nm : V[r].nc);
            q_at[t + 1] = q_at[t] + (pass == 0 ? V[r].nmm : V[r].ncm);
        }
    for (uint32_t r = 0; r <= n; r++) X.frag_cls[r] = (uint32_t)c_at[(size_t)n + r];
    auto place = [&](size_t t) {
        const int pass = t < n ? 0 : 1;
        const View& v = V[t - (size_t)pass * n];
        const uint64_t c0 = pass == 0 ? 0 : v.nm, c1 = pass == 0 ? v.nm : v.nm + v.nc;
        uint64_t c = c_at[t], q = q_at[t];
        for (uint64_t k = c0; k < c1; k++, c++) {
            const uint32_t b = v.ptr[k], e = v.ptr[k + 1], len = e - b;
            for (uint32_t j = b; j < e; j++, q++) { C.iso[q] = v.iso[j]; C.prop[q] = 1.0 / (double)len; }
            C.ptr[c + 1] = (uint32_t)q;
            C.count[c] = v.count[k]; C.state[c] = v.state[k]; C.postassign[c] = v.post[k];
        }
        if (pass == 1) memcpy(X.iso_count.data() + v.iso_lo, v.ic, 8 * (v.iso_hi - v.iso_lo));
    };
    {
        const unsigned nt = std::max(1u, std::min<unsigned>(wq_host_threads(), 2 * n));
        wq_on_threads(nt, [&](unsigned th) { for (size_t t = th; t < 2 * (size_t)n; t += nt) place(t); });
    }
    X.cls = std::move(C);
    X.built = true;
    return "";
}

// assign_ambig!, src/quant.jl:520-553 (host: one pass over the multi-map classes)
static inline std::string wq_host_assign_ambig_impl(const WqHostIndex& H, const WqIsoIndex& S, const WqHostQuant& hq, WqHostQuantEx& X) {
    if (!X.built || !X.em_done) return "assign_ambig needs wq_em_tpm first (it uses the EM class proportions and gene TPM)";
    if (X.ambig_done) return "";
    const WqClassSet& C = X.cls;
    // the share of every alignment of every class (independent: host threads), then the additions in class order
    // (sequential: floating-point sums onto one node / edge must see their terms in the reference's order)
    std::vector<double> share((size_t)C.n_multi * WQ_MAX_TOLERATE, 0.0);
    std::vector<uint8_t> skip(C.n_multi, 0);
    std::vector<uint32_t> stale(C.n_multi, 0);      // paired classes: mate-2 nodes of the first alignment that the stale isoform-id set covers
    unsigned nthreads = wq_host_threads();
    if (C.n_multi < 1024) nthreads = 1;
    std::vector<std::string> errs(nthreads);
    auto work = [&](unsigned t) {
        WqGeneBits B;
        uint32_t cached = 0;
        WqKeyPath vp[WQ_MAX_TOLERATE];
        double w[WQ_MAX_TOLERATE];
        const uint32_t c_lo = (uint32_t)((uint64_t)C.n_multi * t / nthreads), c_hi = (uint32_t)((uint64_t)C.n_multi * (t + 1) / nthreads);
        for (uint32_t c = c_lo; c < c_hi; c++) {
            const uint32_t* key = hq.multis.key(c);
            const uint32_t nh = key[0] & 0xFFFF;
            if (nh > WQ_MAX_TOLERATE) { errs[t] = "multi-map class with more alignments than WQ_MAX_TOLERATE"; return; }
            size_t pos = 1;
            for (uint32_t h = 0; h < nh; h++) pos = wq_key_next(key, pos, vp[h]);
            // (with the quant stage partitioned over ranks every rank still applies every class: the event stage is partitioned by
            // pieces of work, not by gene range, so a rank may build events of any gene and needs complete node / edge values)
            for (uint32_t h = 0; h < nh; h++) {
                if (C.postassign[c]) { w[h] = X.genetpm[vp[h].gene - 1]; continue; }
                if (cached != vp[h].gene) { B.build(S, vp[h].gene - 1, H.genes[vp[h].gene - 1].nn); cached = vp[h].gene; }
                const uint32_t ib = S.iso_base[vp[h].gene - 1];
                double acc = 0.0;
                for (uint32_t k = C.ptr[c]; k < C.ptr[c + 1]; k++) {
                    int64_t p = (int64_t)C.iso[k] - ib;
                    if (p >= 0 && p < B.np && wq_key_compatible(B, (uint32_t)p, vp[h])) acc += C.prop[k];
                }
                w[h] = acc;
            }
            double tot = 0.0;
            for (uint32_t h = 0; h < nh; h++) tot += w[h];
            // mc.count: with --biascorrect gc_adjust!(multi, mod) has scaled it since the EM (src/quant.jl:486-495, bin/whippet-quant.jl:169)
            const double mcount = hq.bias ? WqHostQuant::gc_apply(C.count[c], hq.multis.bias_g[c]) : C.count[c];
            for (uint32_t h = 0; h < nh; h++) share[(size_t)c * WQ_MAX_TOLERATE + h] = (w[h] / tot) * mcount;
            if (!C.postassign[c] && vp[0].rnd) {
                // ambig.iset still holds {p + geneidx[g]} of the LAST alignment when the first one is spread (see wq_spread_path)
                const WqKeyPath& last = vp[nh - 1];
                if (cached != last.gene) { B.build(S, last.gene - 1, H.genes[last.gene - 1].nn); cached = last.gene; }
                const int64_t ib = S.iso_base[last.gene - 1];
                uint32_t m = 0;
                for (uint32_t i = 0; i < vp[0].rn && i < 32; i++) {
                    const int64_t p = (int64_t)vp[0].rnd[i] - ib - 1;          // node id == p + 1 + geneidx  <=>  0-based isoform p
                    if (p >= 0 && p < B.np && wq_key_compatible(B, (uint32_t)p, last)) m |= 1u << i;
                }
                stale[c] = m;
            }
        }
    };
    if (nthreads == 1) work(0);
    else { std::vector<std::thread> th; for (unsigned t = 0; t < nthreads; t++) th.emplace_back(work, t); for (auto& x : th) x.join(); }
    for (auto& e : errs) if (!e.empty()) return e;
    for (uint32_t c = 0; c < C.n_multi; c++) {
        if (skip[c]) continue;
        const uint32_t* key = hq.multis.key(c);
        const uint32_t nh = key[0] & 0xFFFF;
        size_t pos = 1;
        for (uint32_t h = 0; h < nh; h++) {
            WqKeyPath kp;
            pos = wq_key_next(key, pos, kp);
            wq_spread_path(H, X.node_value, X.edge_adds, kp, share[(size_t)c * WQ_MAX_TOLERATE + h], h == 0 ? stale[c] : 0u);
        }
    }
    X.ambig_done = true;
    X.edges_final = false;
    return "";
}

// ------------------------------------------------------------------------------------------ device
// Device layout of the class set: classes are PERMUTED so that every block of the cooperative grid owns one contiguous
// range of classes and therefore one contiguous range of class members (multi-map classes, the slow ones to converge, are
// dealt round-robin over the blocks; isoform classes follow in gene order, balanced by member count).  The canonical class
// order only matters for the per-isoform accumulation, and that order lives in the contribution lists (con_*).
struct WqEmDev {
    uint32_t I, C, ntiles;
    const double* leff;          // [I] max(length - readlen, 1)
    double* count;               // [I] unique + finished-class counts (in/out)
    double* tpm;                 // [I] out: rounded TPM
    double* tpm_raw;             // [I]
    const uint32_t* cls_ptr;     // [C+1] member ranges (device class order)
    const int32_t*  cls_iso;     // [M] member isoforms
    const uint32_t* mem_cls;     // [M] class of every member
    double*         prop;        // [M]
    const double*   cls_count;   // [C]
    uint8_t*        state;       // [C] 0 active, 1 finished in this sweep, 2 finished earlier (or never active)
    const uint32_t* blk_cls;     // [grid+1] class range of every block
    const uint32_t* blk_mem;     // [grid+1] member range of every block
    double* vals;                // [M] scratch: TPM of every member of an active class
    double* psum;                // [C] scratch
    int*    cdiff;               // [C] scratch: some proportion of the class changed
    const uint32_t* con_ptr;     // [I+1] contributions of each isoform, in canonical class order
    const uint2*    con_rec;     // [M] (device class, member position)
    double* w_all;               // [M] scratch: prop * count of unfinished classes, else +0
    double* w_fin;               // [M] scratch: prop * count of classes finished in this sweep, else +0
    double* partial;             // [ntiles]
    int*    flags;               // [4] rotating "tpm changed" flags + [3] "tpm != ones" of the first calculate_tpm!
    int64_t* it_out;
    int64_t maxit;
    double sc_tpm, sc_prop;      // 10^sig, 10^4
    int    sig;
    long long* dbg;              // optional [grid][4] clock totals (gather, barrier 1, normalise + sweep, barrier 2) when WQ_EM_TIMING is set
};

__device__ __forceinline__ double wq_round_sc(double x, double sc) {   // Base.round(x, digits=d)
    double r = __ddiv_rn(rint(__dmul_rn(x, sc)), sc);
    return isfinite(r) ? r : x;
}

// adjacent-pair tree over the 1024 values of a block (one per thread), every thread gets the sum.  The pairing is the
// declared canonical one (v[t] + v[t + stride] for t a multiple of 2*stride): the first five levels are warp shuffles
// (lane l adds lane l + stride; only lanes that are multiples of 2*stride carry tree nodes), then the 32 warp sums
// go through shared memory and one more warp-level pass.
__device__ __forceinline__ double wq_tile_tree(double v, double* buf, int t) {
    #pragma unroll
    for (int stride = 1; stride < 32; stride <<= 1) v = __dadd_rn(v, __shfl_down_sync(0xFFFFFFFFu, v, stride));
    __syncthreads();                                   // buf may still be read from a previous call
    if ((t & 31) == 0) buf[t >> 5] = v;
    __syncthreads();
    double w = buf[t & 31];
    #pragma unroll
    for (int stride = 1; stride < 32; stride <<= 1) w = __dadd_rn(w, __shfl_down_sync(0xFFFFFFFFu, w, stride));
    return __shfl_sync(0xFFFFFFFFu, w, 0);
}

// calculate_tpm!, src/quant.jl:655-667, for one isoform given the raw value count/leff and the folded sum
__device__ __forceinline__ double wq_tpm_value(const WqEmDev& d, double raw, double rpk) {
    double v = __ddiv_rn(__dmul_rn(raw, 1000000.0), rpk);
    if (d.sig > 0) v = wq_round_sc(v, d.sc_tpm);
    return v;
}

// the call before gene_em! is calculate_tpm!(quant, readlen = readlen) with calculate_tpm!'s OWN default sig = 1 (bin/whippet-quant.jl:163,
// src/quant.jl:655), whatever sig gene_em! is given
__device__ __forceinline__ double wq_tpm_value_init(double raw, double rpk) {
    return wq_round_sc(__ddiv_rn(__dmul_rn(raw, 1000000.0), rpk), 10.0);
}

// Persistent cooperative kernel, TWO grid barriers per EM iteration, every pass parallel over class MEMBERS (the
// divisions and roundings), with only the order-defining additions left sequential per class / per isoform:
//   phase B(it)  per tile of 1024 isoforms: (1) every contribution computes prop * count of its class; (2) every isoform
//                adds its contributions up in canonical class order (temp counts; classes finished in this sweep are
//                folded into `count`), raw TPM, tile sum                                                    | grid.sync
//   phase A(it)  every block folds the tile sums, normalises + rounds its isoforms and raises the "changed" flag, then
//                sweeps its classes FOR ITERATION it+1: (1) TPM of every member, (2) per-class sum in member order,
//                (3) new rounded proportion of every member (copy_isdone!), (4) class state                 | grid.sync
// The sweep reads TPM values recomputed from tpm_raw and the folded sum (the same arithmetic as the normalise step, so the
// same bits) because other blocks are still writing d.tpm in that phase.  The sweep of iteration it+1 is issued before the
// loop condition of iteration `it` is known; when the loop then ends because the rounded TPM vector did not change, that
// extra sweep recomputed every proportion from an unchanged TPM vector (identical values) and only moved class states,
// which nothing reads afterwards.  It is skipped when the loop ends at maxit.
__global__ void __launch_bounds__(1024) wq_k_em_tpm(WqEmDev d) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double buf[32];
    const int t = threadIdx.x;
    const uint32_t gtid = blockIdx.x * blockDim.x + t, gsz = gridDim.x * blockDim.x;
    const uint32_t c_lo = d.blk_cls[blockIdx.x], c_hi = d.blk_cls[blockIdx.x + 1];
    const uint32_t m_lo = d.blk_mem[blockIdx.x], m_hi = d.blk_mem[blockIdx.x + 1];
    // ---- calculate_tpm!(quant, readlen) on the unique counts (bin/whippet-quant.jl:163); differs = tpm != ones(I) ----
    for (uint32_t tile = blockIdx.x; tile < d.ntiles; tile += gridDim.x) {
        const uint32_t i = tile * 1024 + t;
        double raw = 0.0;
        if (i < d.I) { raw = __ddiv_rn(d.count[i], d.leff[i]); d.tpm_raw[i] = raw; }
        double s = wq_tile_tree(raw, buf, t);
        if (t == 0) d.partial[tile] = s;
    }
    grid.sync();
    {
        double rpk = fmax(wq_tile_tree((uint32_t)t < d.ntiles ? d.partial[t] : 0.0, buf, t), 1.0);
        int mine = 0;
        for (uint32_t i = gtid; i < d.I; i += gsz) {
            double v = wq_tpm_value_init(d.tpm_raw[i], rpk);
            if (v != 1.0) mine = 1;
            d.tpm[i] = v;
        }
        if (__syncthreads_or(mine) && t == 0) atomicOr(&d.flags[3], 1);
    }
    grid.sync();
    int64_t it = 1;
    bool differs = d.flags[3] != 0;
    while (differs && it < d.maxit) {
        // ---- B: contributions, then the per-isoform sums in class order, raw TPM, tile sums ----
        long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
        if (d.dbg && t == 0) tk0 = clock64();
        if (gtid == 0) d.flags[(it + 1) % 3] = 0;
        for (uint32_t tile = blockIdx.x; tile < d.ntiles; tile += gridDim.x) {
            const uint32_t i = tile * 1024 + t, i_end = min(d.I, tile * 1024 + 1024);
            const uint32_t j_lo = d.con_ptr[tile * 1024], j_hi = d.con_ptr[i_end];
            for (uint32_t j = j_lo + t; j < j_hi; j += 1024) {
                const uint2 rec = d.con_rec[j];
                const uint8_t st = d.state[rec.x];
                double v = 0.0;
                if (st != 2) v = __dmul_rn(d.prop[rec.y], d.cls_count[rec.x]);
                d.w_all[j] = v;
                d.w_fin[j] = st == 1 ? v : 0.0;
            }
            __syncthreads();
            double raw = 0.0;
            if (i < d.I) {
                const double c0 = d.count[i];
                double ct = c0, cn = c0;
                // adding +0.0 for a finished class leaves the (non-negative) sums bit-identical to skipping it
                for (uint32_t j = d.con_ptr[i], je = d.con_ptr[i + 1]; j < je; j++) {
                    ct = __dadd_rn(ct, d.w_all[j]);
                    cn = __dadd_rn(cn, d.w_fin[j]);
                }
                if (cn != c0) d.count[i] = cn;
                raw = __ddiv_rn(ct, d.leff[i]);
                d.tpm_raw[i] = raw;
            }
            double s = wq_tile_tree(raw, buf, t);
            if (t == 0) d.partial[tile] = s;
        }
        if (d.dbg && t == 0) tk1 = clock64();
        grid.sync();
        if (d.dbg && t == 0) tk2 = clock64();
        // ---- A: normalise + round (calculate_tpm!); every block folds the tile sums itself ----
        const double rpk = fmax(wq_tile_tree((uint32_t)t < d.ntiles ? d.partial[t] : 0.0, buf, t), 1.0);
        int mine = 0;
        for (uint32_t i = gtid; i < d.I; i += gsz) {
            double v = wq_tpm_value(d, d.tpm_raw[i], rpk);
            if (v != d.tpm[i]) mine = 1;
            d.tpm[i] = v;
        }
        // ---- sweep over this block's classes for iteration it+1 (maximize step, src/quant.jl:726-749) ----
        if (it + 1 < d.maxit) {
            for (uint32_t k = m_lo + t; k < m_hi; k += 1024)
                if (d.state[d.mem_cls[k]] == 0) d.vals[k] = wq_tpm_value(d, d.tpm_raw[d.cls_iso[k]], rpk);
            __syncthreads();
            for (uint32_t c = c_lo + t; c < c_hi; c += 1024) {
                if (d.state[c] != 0) continue;
                double ps = 0.0;
                for (uint32_t k = d.cls_ptr[c], ke = d.cls_ptr[c + 1]; k < ke; k++) ps = __dadd_rn(ps, d.vals[k]);
                d.psum[c] = ps;
                d.cdiff[c] = 0;
            }
            __syncthreads();
            for (uint32_t k = m_lo + t; k < m_hi; k += 1024) {          // copy_isdone!, src/quant.jl:702-712
                const uint32_t c = d.mem_cls[k];
                if (d.state[c] != 0) continue;
                const double val = wq_round_sc(__ddiv_rn(d.vals[k], d.psum[c]), d.sc_prop);
                if (val != d.prop[k]) d.cdiff[c] = 1;
                d.prop[k] = isnan(val) ? 0.0 : val;
            }
            __syncthreads();
            for (uint32_t c = c_lo + t; c < c_hi; c += 1024) {
                const uint8_t st = d.state[c];
                if (st == 1) d.state[c] = 2;
                else if (st == 0 && (!d.cdiff[c] || d.psum[c] == 0.0)) d.state[c] = 1;
            }
        }
        if (__syncthreads_or(mine) && t == 0) atomicOr(&d.flags[it % 3], 1);
        if (d.dbg && t == 0) tk3 = clock64();
        grid.sync();
        if (d.dbg && t == 0) {
            long long* o = d.dbg + 4 * blockIdx.x;
            o[0] += tk1 - tk0; o[1] += tk2 - tk1; o[2] += tk3 - tk2; o[3] += clock64() - tk3;
        }
        differs = d.flags[it % 3] != 0;
        it += 1;
    }
    if (gtid == 0) *d.it_out = it;
}

// The same algorithm with the block's working set in SHARED MEMORY.  grid.sync() fences invalidate L1, so the kernel
// above re-fetches its (constant) index arrays and its block-local scratch from L2 in every iteration, several dependent
// round trips per phase.  Here every block copies, once, the member / class index of its class range and the
// contribution records of its isoform tile into dynamic shared memory, keeps the sweep's scratch (member TPM, class sums,
// proportions, class states) and the gather's products there, and holds each isoform's count, raw and rounded TPM in the
// registers of the thread that owns it.  Per iteration only three kinds of global traffic remain: state / prop / class
// count of every contribution (one round trip), the tile sums, and tpm_raw of the class members (one round trip);
// proportions and states are written through to global memory only when they change.  Same arithmetic in the same order
// as wq_k_em_tpm, so the same bits.  Requires one isoform tile per block (ntiles <= grid) and a working set within the
// opt-in shared-memory limit; wq_run_em_tpm falls back to wq_k_em_tpm otherwise.
struct WqEmSmem { uint32_t max_mem, max_cls, max_con; };
__global__ void __launch_bounds__(1024) wq_k_em_tpm_smem(WqEmDev d, const WqEmSmem z) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_vals = reinterpret_cast<double*>(smem_raw);            // [max_mem] TPM of every member of an active class
    double* s_prop = s_vals + z.max_mem;                             // [max_mem]
    double* s_psum = s_prop + z.max_mem;                             // [max_cls]
    double* s_w    = s_psum + z.max_cls;                             // [max_con] prop * count of unfinished classes, else +0
    uint2*  s_rec  = reinterpret_cast<uint2*>(s_w + z.max_con);      // [max_con] (device class, member position)
    double* buf    = reinterpret_cast<double*>(s_rec + z.max_con);   // [32]
    uint32_t* s_iso  = reinterpret_cast<uint32_t*>(buf + 32);        // [max_mem]
    uint32_t* s_mcls = s_iso + z.max_mem;                            // [max_mem] class of the member, local index
    uint32_t* s_cptr = s_mcls + z.max_mem;                           // [max_cls + 1] member ranges, local offsets
    uint8_t* s_state = reinterpret_cast<uint8_t*>(s_cptr + z.max_cls + 1);   // [max_cls]
    uint8_t* s_cdiff = s_state + z.max_cls;                          // [max_cls]
    uint8_t* s_wst   = s_cdiff + z.max_cls;                          // [max_con] class state seen by the contribution
    const int t = threadIdx.x;
    const uint32_t gtid = blockIdx.x * blockDim.x + t;
    const uint32_t c_lo = d.blk_cls[blockIdx.x], ncls = d.blk_cls[blockIdx.x + 1] - c_lo;
    const uint32_t m_lo = d.blk_mem[blockIdx.x], nmem = d.blk_mem[blockIdx.x + 1] - m_lo;
    const bool has_tile = blockIdx.x < d.ntiles;
    const uint32_t i = blockIdx.x * 1024 + t;                        // the isoform this thread owns
    const bool own = has_tile && i < d.I;
    uint32_t j_lo = 0, ncon = 0, my_j = 0, my_n = 0;
    if (has_tile) { j_lo = d.con_ptr[blockIdx.x * 1024]; ncon = d.con_ptr[min(d.I, blockIdx.x * 1024 + 1024)] - j_lo; }
    if (own) { my_j = d.con_ptr[i] - j_lo; my_n = d.con_ptr[i + 1] - d.con_ptr[i]; }
    // ---- one-time staging ----
    for (uint32_t k = t; k < nmem; k += 1024) { s_iso[k] = (uint32_t)d.cls_iso[m_lo + k]; s_mcls[k] = d.mem_cls[m_lo + k] - c_lo; s_prop[k] = d.prop[m_lo + k]; }
    for (uint32_t c = t; c <= ncls; c += 1024) s_cptr[c] = d.cls_ptr[c_lo + c] - m_lo;
    for (uint32_t c = t; c < ncls; c += 1024) s_state[c] = d.state[c_lo + c];
    for (uint32_t j = t; j < ncon; j += 1024) s_rec[j] = d.con_rec[j_lo + j];
    const double leff = own ? d.leff[i] : 1.0;
    double cnt = own ? d.count[i] : 0.0;                             // unique + finished-class counts of this isoform
    double raw = 0.0, tpm = 0.0;
    __syncthreads();
    // ---- calculate_tpm!(quant, readlen) on the unique counts (bin/whippet-quant.jl:163); differs = tpm != ones(I) ----
    if (has_tile) {
        if (own) { raw = __ddiv_rn(cnt, leff); d.tpm_raw[i] = raw; }
        double sm = wq_tile_tree(raw, buf, t);
        if (t == 0) d.partial[blockIdx.x] = sm;
    }
    grid.sync();
    {
        const double rpk = fmax(wq_tile_tree((uint32_t)t < d.ntiles ? d.partial[t] : 0.0, buf, t), 1.0);
        int mine = 0;
        if (own) { tpm = wq_tpm_value_init(raw, rpk); if (tpm != 1.0) mine = 1; }
        if (__syncthreads_or(mine) && t == 0) atomicOr(&d.flags[3], 1);
    }
    grid.sync();
    int64_t it = 1;
    bool differs = d.flags[3] != 0;
    while (differs && it < d.maxit) {
        long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
        if (d.dbg && t == 0) tk0 = clock64();
        if (gtid == 0) d.flags[(it + 1) % 3] = 0;
        // ---- B: contributions, then the per-isoform sums in class order, raw TPM, tile sum ----
        if (has_tile) {
            for (uint32_t j = t; j < ncon; j += 1024) {
                const uint2 rec = s_rec[j];
                const uint8_t st = d.state[rec.x];
                double v = 0.0;
                if (st != 2) v = __dmul_rn(d.prop[rec.y], d.cls_count[rec.x]);
                s_w[j] = v;
                s_wst[j] = st;
            }
            __syncthreads();
            raw = 0.0;
            if (own) {
                double ct = cnt, cn = cnt;
                // adding +0.0 for a finished class leaves the (non-negative) sums bit-identical to skipping it
                for (uint32_t j = my_j, je = my_j + my_n; j < je; j++) {
                    const double v = s_w[j];
                    ct = __dadd_rn(ct, v);
                    cn = __dadd_rn(cn, s_wst[j] == 1 ? v : 0.0);
                }
                cnt = cn;
                raw = __ddiv_rn(ct, leff);
                d.tpm_raw[i] = raw;
            }
            double sm = wq_tile_tree(raw, buf, t);
            if (t == 0) d.partial[blockIdx.x] = sm;
        }
        if (d.dbg && t == 0) tk1 = clock64();
        grid.sync();
        if (d.dbg && t == 0) tk2 = clock64();
        // ---- A: normalise + round (calculate_tpm!); every block folds the tile sums itself ----
        const double rpk = fmax(wq_tile_tree((uint32_t)t < d.ntiles ? d.partial[t] : 0.0, buf, t), 1.0);
        int mine = 0;
        if (own) { const double v = wq_tpm_value(d, raw, rpk); if (v != tpm) mine = 1; tpm = v; }
        // ---- sweep over this block's classes for iteration it+1 (maximize step, src/quant.jl:726-749) ----
        if (it + 1 < d.maxit) {
            for (uint32_t k = t; k < nmem; k += 1024)
                if (s_state[s_mcls[k]] == 0) s_vals[k] = wq_tpm_value(d, d.tpm_raw[s_iso[k]], rpk);
            __syncthreads();
            for (uint32_t c = t; c < ncls; c += 1024) {
                if (s_state[c] != 0) continue;
                double ps = 0.0;
                for (uint32_t k = s_cptr[c], ke = s_cptr[c + 1]; k < ke; k++) ps = __dadd_rn(ps, s_vals[k]);
                s_psum[c] = ps;
                s_cdiff[c] = 0;
            }
            __syncthreads();
            for (uint32_t k = t; k < nmem; k += 1024) {               // copy_isdone!, src/quant.jl:702-712
                const uint32_t c = s_mcls[k];
                if (s_state[c] != 0) continue;
                const double val = wq_round_sc(__ddiv_rn(s_vals[k], s_psum[c]), d.sc_prop);
                const double old = s_prop[k];
                if (val != old) s_cdiff[c] = 1;
                const double nv = isnan(val) ? 0.0 : val;
                if (nv != old || isnan(old)) { s_prop[k] = nv; d.prop[m_lo + k] = nv; }
            }
            __syncthreads();
            for (uint32_t c = t; c < ncls; c += 1024) {
                const uint8_t st = s_state[c];
                if (st == 1) { s_state[c] = 2; d.state[c_lo + c] = 2; }
                else if (st == 0 && (!s_cdiff[c] || s_psum[c] == 0.0)) { s_state[c] = 1; d.state[c_lo + c] = 1; }
            }
        }
        if (__syncthreads_or(mine) && t == 0) atomicOr(&d.flags[it % 3], 1);
        if (d.dbg && t == 0) tk3 = clock64();
        grid.sync();
        if (d.dbg && t == 0) {
            long long* o = d.dbg + 4 * blockIdx.x;
            o[0] += tk1 - tk0; o[1] += tk2 - tk1; o[2] += tk3 - tk2; o[3] += clock64() - tk3;
        }
        differs = d.flags[it % 3] != 0;
        it += 1;
    }
    if (own) { d.count[i] = cnt; d.tpm[i] = tpm; }
    if (gtid == 0) *d.it_out = it;
}

// grow-only device scratch kept in the handle (cudaMalloc/cudaFree per call would serialise the device)
struct WqEmScratch {
    void* p[26] = {nullptr}; size_t cap[26] = {0};
    void* hp = nullptr; size_t hcap = 0;               // page-locked staging for results (a copy into pageable memory is staged by the driver, copy by copy)
    cudaError_t pinned(size_t bytes, void** out) {
        if (bytes > hcap) {
            if (hp) cudaFreeHost(hp);
            hp = nullptr; hcap = 0;
            cudaError_t e = cudaMallocHost(&hp, bytes + bytes / 4 + 4096);
            if (e != cudaSuccess) return e;
            hcap = bytes + bytes / 4 + 4096;
        }
        *out = hp;
        return cudaSuccess;
    }
    template <class T> cudaError_t get(int slot, size_t n, T** out) {
        size_t bytes = std::max<size_t>(1, n) * sizeof(T);
        if (bytes > cap[slot]) {
            if (p[slot]) cudaFree(p[slot]);
            p[slot] = nullptr; cap[slot] = 0;
            cudaError_t e = cudaMalloc(&p[slot], bytes + bytes / 4);
            if (e != cudaSuccess) return e;
            cap[slot] = bytes + bytes / 4;
        }
        *out = (T*)p[slot];
        return cudaSuccess;
    }
    template <class T> cudaError_t put(int slot, const std::vector<T>& v, T** out, cudaStream_t s) {
        cudaError_t e = get(slot, v.size(), out);
        if (e == cudaSuccess && !v.empty()) e = cudaMemcpyAsync(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
        return e;
    }
    // side streams for kernels that may run next to each other (created on first use, on the current device)
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t tk0 = nullptr, tk1 = nullptr;      // device time of the last kernel (group) launched through this scratch
    bool timed = false;
    cudaError_t timing_events() {
        if (tk0) return cudaSuccess;
        cudaError_t e = cudaEventCreate(&tk0);
        return e != cudaSuccess ? e : cudaEventCreate(&tk1);
    }
    float last_ms() { float ms = 0.f; if (timed && cudaEventElapsedTime(&ms, tk0, tk1) != cudaSuccess) ms = 0.f; return ms; }
    cudaError_t side_streams() {
        if (fork) return cudaSuccess;
        for (int i = 0; i < 3; i++) {
            cudaError_t e = cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking);
            if (e != cudaSuccess) return e;
            e = cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        return cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
    }
    void release() {
        for (int i = 0; i < 26; i++) { if (p[i]) cudaFree(p[i]); p[i] = nullptr; cap[i] = 0; }
        if (hp) cudaFreeHost(hp);
        hp = nullptr; hcap = 0;
        for (int i = 0; i < 3; i++) { if (side[i]) cudaStreamDestroy(side[i]); if (join[i]) cudaEventDestroy(join[i]); side[i] = nullptr; join[i] = nullptr; }
        if (fork) cudaEventDestroy(fork);
        if (tk0) cudaEventDestroy(tk0);
        if (tk1) cudaEventDestroy(tk1);
        fork = tk0 = tk1 = nullptr; timed = false;
    }
};

// powers of ten 10^0 .. 10^308 as the host's libm computes them, uploaded into slot 23 of a scratch
static inline cudaError_t wq_p10_table(struct WqEmScratch& M, cudaStream_t stream, const double** out);
#define WQ_EMCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return std::string(#call) + ": " + cudaGetErrorString(e_); } while (0)

static inline cudaError_t wq_p10_table(WqEmScratch& M, cudaStream_t stream, const double** out) {
    static const std::vector<double> tab = [] { std::vector<double> t(309); for (int k = 0; k <= 308; k++) t[k] = std::pow(10.0, (double)k); return t; }();
    double* dp = nullptr;
    const bool fresh = M.cap[23] == 0;
    cudaError_t e = M.get(23, tab.size(), &dp);
    if (e == cudaSuccess && fresh) e = cudaMemcpyAsync(dp, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, stream);
    *out = dp;
    return e;
}

static inline std::string wq_run_em_tpm(const WqHostIndex& H, const WqIsoIndex& S, WqHostQuant& hq, WqHostQuantEx& X, WqEmScratch& M,
                                        const wq_em_param& p, cudaStream_t stream, uint64_t* launches,
                                        double* tpm_out, double* count_out, double* genetpm_out, int64_t* iterations,
                                        const std::function<void()>* while_running = nullptr) {
    if (p.sig < 0 || p.sig > 12) return "sig out of range";
    const bool prof = getenv("WQ_PROFILE") != nullptr;
    if (!X.built) {
        auto t0 = std::chrono::steady_clock::now();
        std::string e = wq_host_build_classes(H, S, hq, X);
        if (!e.empty()) return e;
        if (prof) fprintf(stderr, "[wq]   build_classes              %9.3f ms (%zu classes, %u multi)\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), X.cls.count.size(), X.cls.n_multi);
    }
    if (X.em_done) return "wq_em_tpm was already run on this count state (gene_em! mutates it); call wq_quant_reset first";
    static const bool prof2 = getenv("WQ_PROFILE") && atoi(getenv("WQ_PROFILE")) >= 2;
    const auto te0 = std::chrono::steady_clock::now();
    auto lap2 = [&](const char* what) { if (prof2) fprintf(stderr, "[wq]     em host: %-24s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - te0).count()); };
    const uint32_t I = S.I;
    WqClassSet& C = X.cls;
    const uint32_t NC = (uint32_t)C.count.size();
    const uint32_t ntiles = (I + 1023) / 1024;
    if (ntiles > 1024) return "more than 1048576 isoforms";
    std::vector<double> leff(I), tpm(I), count = X.iso_count;
    unsigned hnt = wq_host_threads();
    if (NC < 4096) hnt = 1;
    wq_on_threads(hnt, [&](unsigned t) {
        for (uint32_t i = (uint32_t)((uint64_t)I * t / hnt), ie = (uint32_t)((uint64_t)I * (t + 1) / hnt); i < ie; i++)
            leff[i] = std::max((double)(S.length[i] - p.readlen), 1.0);
    });
    int dev = 0, sms = 0, per_sm = 0;
    WQ_EMCK(cudaGetDevice(&dev));
    WQ_EMCK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    WQ_EMCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wq_k_em_tpm, 1024, 0));
    if (per_sm < 1) return "EM kernel does not fit on an SM";
    int grid = sms;                              // one block per SM: a second resident block only lengthens the grid barrier
    if (const char* ev = getenv("WQ_EM_BLOCKS")) grid = std::max(1, std::min(sms * per_sm, atoi(ev)));
    // ---- device class order: block b owns dorder[blk_cls[b] .. blk_cls[b+1]) ----
    const uint32_t NM = (uint32_t)C.iso.size(), nb = (uint32_t)grid, nmulti = C.n_multi;
    std::vector<uint32_t> dorder(NC), dinv(NC), blk_cls(nb + 1, 0), blk_mem(nb + 1, 0);
    {
        // isoform classes: contiguous canonical ranges with ~equal member counts
        const uint64_t iso_members = NM - C.ptr[nmulti];
        std::vector<uint32_t> iso_cut(nb + 1, NC);
        iso_cut[0] = nmulti;
        uint32_t c = nmulti;
        for (uint32_t b = 1; b < nb; b++) {
            const uint64_t want = C.ptr[nmulti] + iso_members * b / nb;
            while (c < NC && C.ptr[c] < want) c++;
            iso_cut[b] = c;
        }
        uint32_t k = 0;
        for (uint32_t b = 0; b < nb; b++) {
            blk_cls[b] = k;
            for (uint32_t m = b; m < nmulti; m += nb) dorder[k++] = m;
            for (uint32_t q = iso_cut[b]; q < iso_cut[b + 1]; q++) dorder[k++] = q;
        }
        blk_cls[nb] = k;
    }
    std::vector<uint32_t> d_cptr_h(NC + 1, 0), mem_cls(NM), mpos(NM);       // mpos: canonical member position -> device position
    std::vector<int32_t> d_iso_h(NM);
    std::vector<double> d_prop_h(NM), d_ccount_h(NC);
    std::vector<uint8_t> d_state_h(NC);
    for (uint32_t k = 0; k < NC; k++) {                      // serial prefix over the device order (cheap), members in parallel below
        const uint32_t c = dorder[k];
        dinv[c] = k;
        d_cptr_h[k + 1] = d_cptr_h[k] + (C.ptr[c + 1] - C.ptr[c]);
    }
    wq_on_threads(hnt, [&](unsigned t) {
        for (uint32_t k = (uint32_t)((uint64_t)NC * t / hnt), ke = (uint32_t)((uint64_t)NC * (t + 1) / hnt); k < ke; k++) {
            const uint32_t c = dorder[k];
            d_ccount_h[k] = C.count[c];
            d_state_h[k] = C.state[c];
            for (uint32_t q = C.ptr[c], o = d_cptr_h[k]; q < C.ptr[c + 1]; q++, o++) { d_iso_h[o] = C.iso[q]; d_prop_h[o] = C.prop[q]; mem_cls[o] = k; mpos[q] = o; }
        }
    });
    for (uint32_t b = 0; b <= nb; b++) blk_mem[b] = d_cptr_h[blk_cls[b]];
    // contributions (isoform -> class members) in canonical class order.  Multi-map classes come first and may name any
    // isoform (serial, few members); the isoform classes of one class-building fragment only name isoforms of that
    // fragment's gene range, so the fragments fill disjoint isoform ranges in parallel.
    std::vector<uint32_t> con_ptr(I + 1, 0);
    std::vector<uint2> con_rec(NM);
    const std::vector<uint32_t>& fcl = X.frag_cls;           // [nfrag + 1] class ranges of the fragments
    const unsigned nfrag = fcl.size() > 1 ? (unsigned)fcl.size() - 1 : 0;
    const bool frag_ok = nfrag > 0 && fcl[0] == nmulti && fcl[nfrag] == NC;
    if (frag_ok) {
        for (uint32_t q = 0; q < C.ptr[nmulti]; q++) con_ptr[C.iso[q] + 1]++;
        wq_on_threads(nfrag, [&](unsigned t) { for (uint32_t q = C.ptr[fcl[t]]; q < C.ptr[fcl[t + 1]]; q++) con_ptr[C.iso[q] + 1]++; });
    } else for (int32_t i : C.iso) con_ptr[i + 1]++;
    for (uint32_t i = 0; i < I; i++) con_ptr[i + 1] += con_ptr[i];
    {
        std::vector<uint32_t> cur(con_ptr.begin(), con_ptr.end() - 1);
        auto fill = [&](uint32_t c_lo, uint32_t c_hi) {
            for (uint32_t c = c_lo; c < c_hi; c++)
                for (uint32_t q = C.ptr[c]; q < C.ptr[c + 1]; q++) con_rec[cur[C.iso[q]]++] = make_uint2(dinv[c], mpos[q]);
        };
        if (frag_ok) { fill(0, nmulti); wq_on_threads(nfrag, [&](unsigned t) { fill(fcl[t], fcl[t + 1]); }); }
        else fill(0, NC);
    }
    lap2("layout built");
    WqEmDev d;
    double *d_leff, *d_count, *d_tpm, *d_raw, *d_prop, *d_ccount, *d_partial, *d_vals, *d_psum, *d_wall, *d_wfin;
    uint32_t *d_cptr, *d_conptr, *d_memcls, *d_blkcls, *d_blkmem;
    uint2* d_conrec;
    int32_t* d_ciso;
    uint8_t* d_state;
    int *d_flags, *d_cdiff;
    int64_t* d_it;
    WQ_EMCK(M.put(0, leff, &d_leff, stream)); WQ_EMCK(M.put(1, count, &d_count, stream));
    WQ_EMCK(M.get(2, I, &d_tpm)); WQ_EMCK(M.get(3, I, &d_raw));
    WQ_EMCK(M.put(4, d_prop_h, &d_prop, stream)); WQ_EMCK(M.put(5, d_ccount_h, &d_ccount, stream)); WQ_EMCK(M.get(6, 1024, &d_partial));
    WQ_EMCK(M.put(7, d_cptr_h, &d_cptr, stream)); WQ_EMCK(M.put(8, con_ptr, &d_conptr, stream)); WQ_EMCK(M.put(9, con_rec, &d_conrec, stream));
    WQ_EMCK(M.put(10, mem_cls, &d_memcls, stream)); WQ_EMCK(M.put(11, d_iso_h, &d_ciso, stream)); WQ_EMCK(M.put(12, d_state_h, &d_state, stream));
    WQ_EMCK(M.get(13, 4, &d_flags)); WQ_EMCK(M.get(14, 1, &d_it));
    WQ_EMCK(M.put(15, blk_cls, &d_blkcls, stream)); WQ_EMCK(M.put(16, blk_mem, &d_blkmem, stream));
    WQ_EMCK(M.get(17, NM, &d_vals)); WQ_EMCK(M.get(18, NC, &d_psum)); WQ_EMCK(M.get(19, NC, &d_cdiff));
    WQ_EMCK(M.get(20, NM, &d_wall)); WQ_EMCK(M.get(21, NM, &d_wfin));
    WQ_EMCK(cudaMemsetAsync(d_tpm, 0, std::max<size_t>(1, I) * 8, stream));
    WQ_EMCK(cudaMemsetAsync(d_flags, 0, 4 * sizeof(int), stream));
    d.I = I; d.C = NC; d.ntiles = ntiles; d.leff = d_leff; d.count = d_count; d.tpm = d_tpm; d.tpm_raw = d_raw;
    d.cls_ptr = d_cptr; d.cls_iso = d_ciso; d.mem_cls = d_memcls; d.prop = d_prop; d.cls_count = d_ccount; d.state = d_state;
    d.blk_cls = d_blkcls; d.blk_mem = d_blkmem; d.vals = d_vals; d.psum = d_psum; d.cdiff = d_cdiff;
    d.con_ptr = d_conptr; d.con_rec = d_conrec; d.w_all = d_wall; d.w_fin = d_wfin; d.partial = d_partial; d.flags = d_flags;
    d.it_out = d_it; d.sig = (int)p.sig; d.sc_tpm = std::pow(10.0, (double)p.sig); d.sc_prop = 10000.0;
    d.maxit = p.maxit;
    d.dbg = nullptr;
    const bool timing = getenv("WQ_EM_TIMING") != nullptr;
    if (timing) { WQ_EMCK(M.get(22, (size_t)grid * 4, &d.dbg)); WQ_EMCK(cudaMemsetAsync(d.dbg, 0, (size_t)grid * 32, stream)); }
    void* args[] = {&d};
    auto tk0 = std::chrono::steady_clock::now();
    lap2("uploads enqueued");
    WQ_EMCK(M.timing_events());
    WQ_EMCK(cudaEventRecord(M.tk0, stream));
    // working set per block for the shared-memory kernel
    WqEmSmem zs;
    zs.max_mem = 1; zs.max_cls = 1; zs.max_con = 1;
    for (uint32_t b = 0; b < nb; b++) { zs.max_mem = std::max(zs.max_mem, blk_mem[b + 1] - blk_mem[b]); zs.max_cls = std::max(zs.max_cls, blk_cls[b + 1] - blk_cls[b]); }
    for (uint32_t tl = 0; tl < ntiles; tl++) zs.max_con = std::max(zs.max_con, con_ptr[std::min<uint32_t>(I, tl * 1024 + 1024)] - con_ptr[tl * 1024]);
    zs.max_mem = (zs.max_mem + 1) & ~1u; zs.max_cls = (zs.max_cls + 3) & ~3u; zs.max_con = (zs.max_con + 3) & ~3u;      // keep every array aligned
    const size_t smem_bytes = 8ull * (2 * (size_t)zs.max_mem + zs.max_cls + 2 * (size_t)zs.max_con + 32) + 4ull * (2 * (size_t)zs.max_mem + zs.max_cls + 1)
                            + 2ull * zs.max_cls + zs.max_con + 16;
    int smem_optin = 0;
    WQ_EMCK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    static const bool smem_off = getenv("WQ_EM_SMEM") && atoi(getenv("WQ_EM_SMEM")) == 0;
    const bool use_smem = !smem_off && ntiles <= nb && smem_bytes <= (size_t)smem_optin;
    if (use_smem) {
        WQ_EMCK(cudaFuncSetAttribute(wq_k_em_tpm_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        void* args2[] = {&d, &zs};
        WQ_EMCK(cudaLaunchCooperativeKernel((void*)wq_k_em_tpm_smem, dim3(grid), dim3(1024), args2, smem_bytes, stream));
    } else {
        WQ_EMCK(cudaLaunchCooperativeKernel((void*)wq_k_em_tpm, dim3(grid), dim3(1024), args, 0, stream));
    }
    if (prof) fprintf(stderr, "[wq]   em kernel: %s (%zu B shared memory per block; max %u members, %u classes, %u contributions)\n",
                      use_smem ? "shared-memory working set" : "global working set", smem_bytes, zs.max_mem, zs.max_cls, zs.max_con);
    WQ_EMCK(cudaEventRecord(M.tk1, stream));
    M.timed = true;
    if (launches) *launches += 1;
    if (while_running) (*while_running)();       // host work that does not depend on the EM result overlaps the kernel
    int64_t it = 0;
    {   // results through one page-locked staging block: five copies in flight, one synchronisation
        const size_t o_tpm = 8, o_cnt = o_tpm + (size_t)I * 8, o_prop = o_cnt + (size_t)I * 8, o_state = o_prop + (size_t)NM * 8, total = o_state + NC;
        void* stg;
        WQ_EMCK(M.pinned(total, &stg));
        uint8_t* sb = (uint8_t*)stg;
        WQ_EMCK(cudaMemcpyAsync(sb, d_it, 8, cudaMemcpyDeviceToHost, stream));
        WQ_EMCK(cudaMemcpyAsync(sb + o_tpm, d_tpm, (size_t)I * 8, cudaMemcpyDeviceToHost, stream));
        WQ_EMCK(cudaMemcpyAsync(sb + o_cnt, d_count, (size_t)I * 8, cudaMemcpyDeviceToHost, stream));
        WQ_EMCK(cudaMemcpyAsync(sb + o_prop, d_prop, (size_t)NM * 8, cudaMemcpyDeviceToHost, stream));
        WQ_EMCK(cudaMemcpyAsync(sb + o_state, d_state, NC, cudaMemcpyDeviceToHost, stream));
        WQ_EMCK(cudaStreamSynchronize(stream));
        memcpy(&it, sb, 8);
        memcpy(tpm.data(), sb + o_tpm, (size_t)I * 8);
        memcpy(count.data(), sb + o_cnt, (size_t)I * 8);
        memcpy(d_prop_h.data(), sb + o_prop, (size_t)NM * 8);
        memcpy(d_state_h.data(), sb + o_state, NC);
    }
    if (timing) {
        std::vector<long long> dbg((size_t)grid * 4);
        WQ_EMCK(cudaMemcpy(dbg.data(), d.dbg, dbg.size() * 8, cudaMemcpyDeviceToHost));
        const char* nm[4] = {"gather", "barrier 1", "normalise+sweep", "barrier 2"};
        for (int k = 0; k < 4; k++) {
            long long mn = dbg[k], mx = dbg[k], sum = 0;
            for (int b = 0; b < grid; b++) { mn = std::min(mn, dbg[4 * b + k]); mx = std::max(mx, dbg[4 * b + k]); sum += dbg[4 * b + k]; }
            fprintf(stderr, "[wq]     em %-16s cycles/iteration per block: min %8.0f avg %8.0f max %8.0f\n", nm[k],
                    (double)mn / std::max<int64_t>(1, it - 1), (double)sum / grid / std::max<int64_t>(1, it - 1), (double)mx / std::max<int64_t>(1, it - 1));
        }
    }
    for (uint32_t q = 0; q < NM; q++) C.prop[q] = d_prop_h[mpos[q]];
    for (uint32_t c = 0; c < NC; c++) C.state[c] = d_state_h[dinv[c]];
    if (prof) fprintf(stderr, "[wq]   em kernel (grid %d, %lld it) %9.3f ms\n", grid, (long long)it,
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tk0).count());
    X.tpm = tpm; X.iso_count = count;
    X.genetpm.assign(H.G, 0.0);                  // set_gene_tpm!, src/quant.jl:248-257
    for (uint32_t g = 0; g < H.G; g++) {
        double s = 0.0;
        for (uint32_t i = S.iso_base[g]; i < S.iso_base[g + 1]; i++) s += tpm[i];
        X.genetpm[g] = s;
    }
    lap2("results back + gene TPM");
    X.em_done = true;
    X.iterations = it;
    X.em_kernel_ms = M.last_ms();
    if (tpm_out) memcpy(tpm_out, tpm.data(), I * 8);
    if (count_out) memcpy(count_out, count.data(), I * 8);
    if (genetpm_out) memcpy(genetpm_out, X.genetpm.data(), H.G * 8);
    if (iterations) *iterations = it;
    return "";
}

// ------------------------------------------------------------------------------------------ PSI EM
// One event per thread: spliced_em! (src/events.jl:853-893) and rec_tandem_em! (:895-932) with
// calculate_psi! (:831-850) and divsignif! (:802-812).  Events are tiny (a handful of paths and ambiguous
// sets) but run up to 1500 strictly sequential sweeps, so the batch dimension is the parallel one.
struct WqPsiDev {
    uint32_t n_events;
    const uint32_t* path_ptr; const uint32_t* n_inc; const double* count; const double* length;
    const uint32_t* amb_ptr; const uint32_t* amb_path_ptr; const uint32_t* amb_paths; const double* amb_mult;
    const uint8_t* is_tandem;
    double readlen; int sig; int64_t maxit;
    double* psi; double* prev; double* ctemp; double* amb_prop; int32_t* iterations;
    double p10[23];               // exact powers of ten
    double ip10[23];              // 1 / 10^k, correctly rounded (the same bits as an IEEE division on the device)
    const double* p10big;         // [309] 10^k by the host's libm pow (global memory)
};

// IEEE operations that the compiler must not contract or reassociate: intrinsics on the device, plain operators in the host build of
// the same functions (tests/emu drives the cluster kernel's phases on the CPU; host code is compiled without FMA contraction).
#ifdef __CUDA_ARCH__
#define WQ_EM_DADD(a, b) __dadd_rn((a), (b))
#define WQ_EM_DSUB(a, b) __dsub_rn((a), (b))
#define WQ_EM_DMUL(a, b) __dmul_rn((a), (b))
#define WQ_EM_DDIV(a, b) __ddiv_rn((a), (b))
#define WQ_HIINT(x) __double2hiint(x)
#else
#define WQ_EM_DADD(a, b) ((a) + (b))
#define WQ_EM_DSUB(a, b) ((a) - (b))
#define WQ_EM_DMUL(a, b) ((a) * (b))
#define WQ_EM_DDIV(a, b) ((a) / (b))
static inline int wq_hiint_host(double x) { unsigned long long u; memcpy(&u, &x, 8); return (int)(u >> 32); }
#define WQ_HIINT(x) wq_hiint_host(x)
#endif

// 10^k as the host's libm rounds it (Base.round scales by 10.0^digits computed through libm pow; the device pow may be an
// ulp or two off, which changes the rounded result of a value as small as 1e-188 — a psi that decayed for 1500 sweeps).
// Exact for k <= 22 (kernel parameter table), host-computed table up to 10^308, +Inf beyond as in IEEE arithmetic.
__host__ __device__ __forceinline__ double wq_pow10(const WqPsiDev& d, int k) {
    if (k >= 0 && k <= 22) return d.p10[k];
    if (k > 22 && k <= 308 && d.p10big) return d.p10big[k];
    return pow(10.0, (double)k);
}
// hidigit(x, 10) = 1 + floor(log10|x|) (Base.round(x, sigdigits=n)).  The decade is taken from the binary exponent
// and fixed up against exact powers of ten for |x| >= 1 and correctly rounded ones below; values within an ulp of a
// power of ten may land in the neighbouring decade compared with a libm log10, which cannot change the rounded
// result (both decades round such a value to that same power of ten).
__host__ __device__ __forceinline__ int wq_hidigit(const WqPsiDev& d, double ax) {
    if (ax >= 1e-22 && ax < 1e22) {
        const int e = ((WQ_HIINT(ax) >> 20) & 0x7FF) - 1023;  // 2^e <= ax < 2^(e+1) (ax is a normal number here)
        int h = (int)floor((double)e * 0.30102999566398120) + 1;   // candidate: 10^(h-1) <= ax < 10^h, at most one off
        // pw(k) = 10^k
        auto pw = [&](int k) { return k >= 0 ? d.p10[k] : d.ip10[-k]; };
        if (ax >= pw(h)) h += 1;
        else if (ax < pw(h - 1)) h -= 1;
        return h;
    }
    if (d.p10big && ax >= 1e-300 && ax < 1e300) {               // same scheme over the big table (10^-k as the IEEE quotient 1 / 10^k)
        const int e = ((WQ_HIINT(ax) >> 20) & 0x7FF) - 1023;
        int h = (int)floor((double)e * 0.30102999566398120) + 1;
        auto pw = [&](int k) { return k >= 0 ? d.p10big[k] : WQ_EM_DDIV(1.0, d.p10big[-k]); };
        if (ax >= pw(h)) h += 1;
        else if (ax < pw(h - 1)) h -= 1;
        return h;
    }
    return 1 + (int)floor(log10(ax));
}
// Base.round(x, sigdigits = sig)
__host__ __device__ __forceinline__ double wq_round_sig(const WqPsiDev& d, double x, int sig) {
    if (!isfinite(x) || x == 0.0) return x;
    int h = wq_hidigit(d, fabs(x));
    int digits = sig - h;
    double r;
    if (digits >= 0) { double sc = wq_pow10(d, digits); r = WQ_EM_DDIV(rint(WQ_EM_DMUL(x, sc)), sc); }
    else { double sc = wq_pow10(d, -digits); r = WQ_EM_DMUL(rint(WQ_EM_DDIV(x, sc)), sc); }
    if (!isfinite(r)) return digits > 0 ? x : r;
    return r;
}

// One event per block of T = 32 * W threads (events with more than 4 paths), written for the LATENCY of one sweep:
// a fifth of the events never converge under the 4-significant-digit rounding and run all 1500 sweeps, so the launch lasts
// 1500 x (the dependent chain of one sweep of its slowest event).  Per sweep:
//   (1) prop_sum of every set: sequential in member order (the reference's `ac.prop_sum += prev_psi`), the sets side by side on the threads;
//   (2) every (set, path) share `prop / prop_sum * multiplier`: independent IEEE divisions, spread evenly over ALL threads (a path that
//       belongs to many sets does not serialise its divisions on one lane) and stored in the order of the inverted index path -> sets;
//   (3) every path by its owner thread: unambiguous count + its shares added in set order (the order the reference's loop over `ambig`
//       adds them in), then count / length;
//   (4) the sum over all paths as ONE sequential chain in path order (every thread forms it from shared memory), division + rounding.
// Only additions are chained; loads feeding them do not depend on the running sum and are unrolled so that they overlap.
// Shared memory per event: 40 * paths + 16 * sets + 4 * (sets + 1) + 4 * (paths + 1) + 14 * members bytes.  Events too large for that
// (hundreds of paths, tens of thousands of set members: the stress genes of BASELINE config 5, deeply covered paralog families) keep
// the per-path arrays in shared memory, the per-set arrays (prop_sum, multiplier, member offsets) there as well when they fit, and the
// member lists (set -> paths, path -> sets: 2 bytes per member each) in a global scratch (`big`, L2-resident).  The block has 1024
// threads and every path's owner divides its own shares (a path of such an event belongs to hundreds of sets, so the owners are
// as evenly loaded as a pair-parallel split would be, and no share array has to go through memory: 4 bytes of traffic per member
// and sweep instead of 38 -- with 168 such events in flight the sweep is bound by L2 bandwidth, not by the divisions).
struct WqPsiBigCfg {                  // mem == nullptr: everything in shared memory
    double* psum; double* mult; uint32_t* moff; uint16_t* mem; uint16_t* pset;
    const uint64_t* mbase; const uint64_t* abase;      // per launch slot: first member / first set of the event inside the scratch
    uint32_t sets_smem;               // the per-set arrays sit in shared memory behind the per-path ones
};
__device__ __forceinline__ void wq_psi_sync(const bool one_warp) { if (one_warp) __syncwarp(); else __syncthreads(); }
__global__ void __launch_bounds__(1024) wq_k_em_psi(const WqPsiDev d, const uint32_t* __restrict__ list, const uint32_t n_list, const WqPsiBigCfg big) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x, T = blockDim.x;
    const bool one_warp = T == 32;
    const uint32_t slot = blockIdx.x;
    if (slot >= n_list) return;
    const uint32_t e = list[slot];
    const uint32_t p0 = d.path_ptr[e], np = d.path_ptr[e + 1] - p0, ni = d.n_inc[e];
    const uint32_t a0 = d.amb_ptr[e], na = d.amb_ptr[e + 1] - a0;
    const uint32_t m0 = d.amb_path_ptr[a0], nm = d.amb_path_ptr[a0 + na] - m0;
    const bool tandem = d.is_tandem[e] != 0;
    double* psi = reinterpret_cast<double*>(smem_raw);
    double* prev = psi + np;
    double* ct = prev + np;
    double* den = ct + np;
    double* cnt = den + np;
    double* psum; double* mult; double* quo; uint32_t* moff; uint32_t* poff; uint16_t* mem; uint16_t* pset; uint16_t* ppath;
    const bool is_big = big.mem != nullptr;
    if (!is_big) {
        psum = cnt + np;                                          // [na]
        mult = psum + na;                                         // [na]
        quo = mult + na;                                          // [nm] shares, in inverted-index order
        moff = reinterpret_cast<uint32_t*>(quo + nm);             // [na + 1] member offsets of the sets, event-local
        poff = moff + na + 1;                                     // [np + 1] offsets of the inverted index (path -> its sets, in set order)
        mem = reinterpret_cast<uint16_t*>(poff + np + 1);         // [nm] event-local path index of every member
        pset = mem + nm;                                          // [nm] inverted index: the set of every slot
        ppath = pset + nm;                                        // [nm] inverted index: the path of every slot
    } else {
        const uint64_t mb = big.mbase[slot], ab = big.abase[slot];
        if (big.sets_smem) {
            psum = cnt + np; mult = psum + na;
            moff = reinterpret_cast<uint32_t*>(mult + na);
            poff = moff + na + 1;
        } else {
            poff = reinterpret_cast<uint32_t*>(cnt + np);
            psum = big.psum + ab; mult = big.mult + ab; moff = big.moff + ab + slot;
        }
        quo = nullptr; ppath = nullptr;
        mem = big.mem + mb; pset = big.pset + mb;
    }
    const uint32_t* gmem = d.amb_paths + m0;
    for (uint32_t i = tid; i < np; i += T) {
        const double len = d.length[p0 + i];
        cnt[i] = d.count[p0 + i];
        den[i] = tandem ? fmax(__dsub_rn(len, d.readlen), 1.0) : len;
        prev[i] = 0.0; psi[i] = 0.0; ct[i] = cnt[i];
    }
    for (uint32_t a = tid; a <= na; a += T) moff[a] = d.amb_path_ptr[a0 + a] - m0;
    for (uint32_t a = tid; a < na; a += T) mult[a] = d.amb_mult[a0 + a];
    for (uint32_t k = tid; k < nm; k += T) mem[k] = (uint16_t)gmem[k];
    for (uint32_t i = tid; i <= np; i += T) poff[i] = 0;
    wq_psi_sync(one_warp);
    if (tid == 0) {                 // counting sort of the (set, path) pairs by path, stable in the set order; once per event
        for (uint32_t k = 0; k < nm; k++) poff[mem[k] + 1]++;
        for (uint32_t i = 0; i < np; i++) poff[i + 1] += poff[i];
        for (uint32_t a = 0; a < na; a++)
            for (uint32_t k = moff[a]; k < moff[a + 1]; k++) { const uint32_t at = poff[mem[k]]++; pset[at] = (uint16_t)a; if (ppath) ppath[at] = mem[k]; }
        for (uint32_t i = np; i > 0; i--) poff[i] = poff[i - 1];
        poff[0] = 0;
    }
    wq_psi_sync(one_warp);
    // sum of psi over the paths in the reference's order (exclusion paths, then inclusion paths, then their sum; tandem: all paths)
    auto total_of = [&]() -> double {
        if (tandem) {
            double t = 0.0;
            #pragma unroll 8
            for (uint32_t i = 0; i < np; i++) t = __dadd_rn(t, psi[i]);
            return t;
        }
        double se = 0.0, si = 0.0;
        #pragma unroll 8
        for (uint32_t i = ni; i < np; i++) se = __dadd_rn(se, psi[i]);
        #pragma unroll 4
        for (uint32_t i = 0; i < ni; i++) si = __dadd_rn(si, psi[i]);
        return __dadd_rn(se, si);
    };
    __shared__ double s_total;
    auto divide = [&](int sig) {            // psi[] holds count / length on entry
        double total;
        if (one_warp) { total = total_of(); __syncwarp(); }
        else {                              // one warp forms the chain (a broadcast read per path and warp), the others take the result
            if (tid < 32) { const double t = total_of(); if (tid == 0) s_total = t; }
            __syncthreads();
            total = s_total;
        }
        for (uint32_t i = tid; i < np; i += T) {
            const double q = __ddiv_rn(psi[i], total);
            psi[i] = sig > 0 ? wq_round_sig(d, q, sig) : q;
        }
        wq_psi_sync(one_warp);
    };
    auto differs = [&]() -> bool {
        bool diff = false;
        for (uint32_t i = tid; i < np; i += T) if (prev[i] != psi[i]) diff = true;
        return one_warp ? __any_sync(0xffffffffu, diff) != 0 : __syncthreads_or(diff) != 0;
    };
    if (!tandem) {                          // calculate_psi!(inc, exc, counts) before spliced_em!
        for (uint32_t i = tid; i < np; i += T) psi[i] = __ddiv_rn(ct[i], den[i]);
        wq_psi_sync(one_warp);
        divide(0);
    }
    int64_t it = 1;
    for (;;) {
        if (!tandem && !(differs() && it < d.maxit)) break;
        // (1) prop_sum of every set (psi is stable here: the last write was followed by a barrier)
        if (it > 1) {
            for (uint32_t a = tid; a < na; a += T) {
                double sm = 0.0;
                const uint32_t kb = moff[a], ke = moff[a + 1];
                #pragma unroll 4
                for (uint32_t k = kb; k < ke; k++) sm = __dadd_rn(sm, psi[mem[k]]);
                psum[a] = sm;
            }
            wq_psi_sync(one_warp);
        }
        // (2) the shares: first sweep ones(n)/n over prop_sum = 1.0, later sweeps the current psi over their sum (src/events.jl:869-886)
        if (is_big) {
            // (2 + 3) large events: every owner divides its own shares and adds them in set order
            for (uint32_t i = tid; i < np; i += T) {
                double c = cnt[i];
                const double pi = psi[i];
                const uint32_t kb = poff[i], ke = poff[i + 1];
                if (it > 1) {
                    #pragma unroll 4
                    for (uint32_t k = kb; k < ke; k++) { const uint32_t a = pset[k]; c = __dadd_rn(c, __dmul_rn(__ddiv_rn(pi, psum[a]), mult[a])); }
                } else {
                    #pragma unroll 4
                    for (uint32_t k = kb; k < ke; k++) { const uint32_t a = pset[k]; c = __dadd_rn(c, __dmul_rn(__ddiv_rn(1.0, (double)(moff[a + 1] - moff[a])), mult[a])); }
                }
                ct[i] = c;
                prev[i] = pi;
                psi[i] = __ddiv_rn(c, den[i]);            // no other thread reads psi[i] in this step (prop_sum is done, shares are the owner's)
            }
            wq_psi_sync(one_warp);
            divide(d.sig);
            if (tandem) { if (differs() && it < d.maxit) it += 1; else break; }
            else it += 1;
            continue;
        }
        if (it > 1) {
            #pragma unroll 2
            for (uint32_t k = tid; k < nm; k += T) {
                const uint32_t a = pset[k];
                quo[k] = __dmul_rn(__ddiv_rn(psi[ppath[k]], psum[a]), mult[a]);
            }
        } else {
            for (uint32_t k = tid; k < nm; k += T) {
                const uint32_t a = pset[k];
                quo[k] = __dmul_rn(__ddiv_rn(1.0, (double)(moff[a + 1] - moff[a])), mult[a]);      // (1 / n) / 1.0 = 1 / n
            }
        }
        wq_psi_sync(one_warp);
        // (3) every path by its owner
        for (uint32_t i = tid; i < np; i += T) {
            double c = cnt[i];
            const uint32_t kb = poff[i], ke = poff[i + 1];
            #pragma unroll 4
            for (uint32_t k = kb; k < ke; k++) c = __dadd_rn(c, quo[k]);
            ct[i] = c;
            prev[i] = psi[i];
            psi[i] = __ddiv_rn(c, den[i]);
        }
        wq_psi_sync(one_warp);
        // (4)
        divide(d.sig);
        if (tandem) { if (differs() && it < d.maxit) it += 1; else break; }
        else it += 1;
    }
    for (uint32_t i = tid; i < np; i += T) { d.psi[p0 + i] = psi[i]; d.ctemp[p0 + i] = ct[i]; }
    if (tid == 0) d.iterations[e] = (int32_t)it;
}
static inline size_t wq_psi_block_smem(uint32_t np, uint32_t na, uint32_t nm) {
    size_t b = 40ull * np + 16ull * na + 4ull * (na + 1) + 4ull * (np + 1) + 14ull * nm;
    return (b + 15) & ~15ull;
}
static inline size_t wq_psi_big_smem(uint32_t np, uint32_t na, bool sets) {      // per-path arrays (+ the per-set ones)
    return (40ull * np + 4ull * (np + 1) + (sets ? 16ull * na + 4ull * (na + 1) : 0ull) + 15) & ~15ull;
}


// ---- large events on a thread-block cluster ----
// An event of ~1000-1900 paths, ~150 ambiguous sets and ~60-140 k set members runs all 1500 sweeps, and a deep human job has ~170 of
// them (the nodes of one titin-like gene under one long exon-skipping junction).  One block per event leaves such a sweep at ~100 us:
// 0.6 MB of member lists streamed from L2, 140 k IEEE divisions on one SM's Float64 pipe, and most SMs idle when the events are shared
// between the ranks of a multi-GPU job.  Here C = 2 / 4 / 8 blocks of a cluster share one event.  Paths are dealt to the blocks in
// groups of 32 (block r owns the paths i with (i / 32) % C == r), sets round-robin (a % C == r); a block keeps ONLY the members of
// its own sets (as bit masks over the paths) and the inverted index of its own paths -- divided by C the whole working set sits in
// shared memory and a sweep touches no global memory at all.  Per sweep, three cluster barriers:
//   (1) prop_sum of the block's sets (sequential in member order = ascending path order, a lane per set walking the set's bit mask
//       over broadcast reads of psi), written into every block's psum[] through distributed shared memory;                                                                                  -- barrier
//   (2) the block's paths: unambiguous count + shares in set order (the owner divides its own shares), count / length written into
//       every block's raw[];                                                                                        -- barrier
//   (3) every block forms the same total from raw[] (the reference's order: exclusion paths, inclusion paths, their sum -- two
//       chains on two warps), then normalises + rounds its own paths, writes them into every block's psi[] and its "changed" flag
//       into every block's flag word.                                                                               -- barrier
// The arithmetic of every path and set is the block kernel's, operation for operation; only who does it changed.  The phases are
// __host__ __device__ so that tests/emu can run the same code for all blocks and threads on the CPU against the oracle.
struct WqPsiClCfg {
    uint32_t C;                                   // blocks per cluster
    uint32_t np_cap, na_cap, npl_cap, nal_cap;    // over the launch's events: paths, sets, path slots of one block, sets of one block
    uint32_t inv_cap;                             // entries of one block's inverted index
    uint32_t split;                               // 1: several lanes share a path's divisions in phase (2) when the block has threads to spare (WQ_PSI_CL_SPLIT=1; slower as measured)
    long long* dbg;                               // optional [8] clock totals of cluster 0, block 0 (WQ_PSI_CL_TIMING): build, prop_sum, barrier, shares, barrier, total, normalise + barrier, sweeps
};
struct WqPsiClSmem {                              // one block's arrays; the layout is identical in every block of the launch
    double *psi, *raw, *psum, *mult;              // complete copies: [np] rounded psi, [np] count / length, [na], [na]
    double *prev, *ct, *den, *cnt, *tot;          // the block's own path slots; tot[2] = the two partial sums of the total
    uint32_t *lmask, *poff, *cur, *flags;         // [nal][ceil(np / 32)] member bit masks of the own sets, [npl + 1] inverted index, build cursor, [8] changed flags
    uint16_t *lpset;                              // sets of the own paths, in set order
};
WQ_HDN inline size_t wq_psi_cl_layout(const WqPsiClCfg& z, unsigned char* base, WqPsiClSmem* o) {
    size_t at = 0;
    auto take = [&](size_t bytes) { const size_t p = at; at += (bytes + 15) & ~(size_t)15; return p; };
    const uint32_t nwords = (z.np_cap + 31u) >> 5;               // psi[] is padded to whole mask words: phase (1) reads 32 values per word
    const size_t o_psi = take(8ull * 32u * nwords), o_raw = take(8ull * z.np_cap), o_psum = take(8ull * z.na_cap), o_mult = take(8ull * z.na_cap);
    const size_t o_prev = take(8ull * z.npl_cap), o_ct = take(8ull * z.npl_cap), o_den = take(8ull * z.npl_cap), o_cnt = take(8ull * z.npl_cap), o_tot = take(16);
    const size_t o_lmask = take(4ull * z.nal_cap * nwords), o_poff = take(4ull * (z.npl_cap + 1)), o_cur = take(4ull * z.npl_cap), o_flags = take(32);
    const size_t o_lpset = take(2ull * z.inv_cap);
    if (o) {
        o->psi = reinterpret_cast<double*>(base + o_psi); o->raw = reinterpret_cast<double*>(base + o_raw);
        o->psum = reinterpret_cast<double*>(base + o_psum); o->mult = reinterpret_cast<double*>(base + o_mult);
        o->prev = reinterpret_cast<double*>(base + o_prev); o->ct = reinterpret_cast<double*>(base + o_ct);
        o->den = reinterpret_cast<double*>(base + o_den); o->cnt = reinterpret_cast<double*>(base + o_cnt); o->tot = reinterpret_cast<double*>(base + o_tot);
        o->lmask = reinterpret_cast<uint32_t*>(base + o_lmask); o->poff = reinterpret_cast<uint32_t*>(base + o_poff);
        o->cur = reinterpret_cast<uint32_t*>(base + o_cur); o->flags = reinterpret_cast<uint32_t*>(base + o_flags);
        o->lpset = reinterpret_cast<uint16_t*>(base + o_lpset);
    }
    return at;
}
// ownership: paths in groups of 32 round-robin over the blocks, sets round-robin
WQ_HDN inline uint32_t wq_cl_owner(uint32_t i, uint32_t C) { return (i >> 5) % C; }
WQ_HDN inline uint32_t wq_cl_slot(uint32_t i, uint32_t C) { return ((i >> 5) / C) * 32u + (i & 31u); }
WQ_HDN inline uint32_t wq_cl_path(uint32_t s, uint32_t r, uint32_t C) { return ((((s >> 5) * C) + r) << 5) + (s & 31u); }
WQ_HDN inline uint32_t wq_cl_nslots(uint32_t np, uint32_t r, uint32_t C) { const uint32_t w = (np + 31u) >> 5; return w > r ? ((w - r + C - 1) / C) * 32u : 0u; }
WQ_HDN inline uint32_t wq_cl_nsets(uint32_t na, uint32_t r, uint32_t C) { return na > r ? (na - r + C - 1) / C : 0u; }

struct WqPsiClEv { uint32_t e, p0, np, ni, a0, na, m0, nm; bool tandem; };
WQ_HDN inline WqPsiClEv wq_psi_cl_event(const WqPsiDev& d, uint32_t e) {
    WqPsiClEv v;
    v.e = e; v.p0 = d.path_ptr[e]; v.np = d.path_ptr[e + 1] - v.p0; v.ni = d.n_inc[e];
    v.a0 = d.amb_ptr[e]; v.na = d.amb_ptr[e + 1] - v.a0;
    v.m0 = d.amb_path_ptr[v.a0]; v.nm = d.amb_path_ptr[v.a0 + v.na] - v.m0;
    v.tandem = d.is_tandem[e] != 0;
    return v;
}
struct WqClCtx {
    uint32_t r, C, T;                             // this block's rank in the cluster, blocks per cluster, threads per block
    WqPsiClSmem S;
    unsigned char* self;                          // host emulation only: this block's buffer and every block's
    unsigned char* const* all;
};
// the same array element in block rr's shared memory
template <class Tp> WQ_HDN inline Tp* wq_cl_remote(const WqClCtx& c, Tp* p, uint32_t rr) {
#ifdef __CUDA_ARCH__
    return cg::this_cluster().map_shared_rank(p, rr);
#else
    return reinterpret_cast<Tp*>(c.all[rr] + (reinterpret_cast<unsigned char*>(p) - c.self));
#endif
}
// -- build, once per event.  a: own path slots, complete arrays, own set offsets.
WQ_HDN inline void wq_cl_init_a(const WqPsiDev& d, const WqPsiClEv& ev, const WqClCtx& c, uint32_t tid) {
    const WqPsiClSmem& S = c.S;
    const uint32_t nsl = wq_cl_nslots(ev.np, c.r, c.C);
    for (uint32_t s = tid; s < nsl; s += c.T) {
        const uint32_t i = wq_cl_path(s, c.r, c.C);
        if (i < ev.np) {
            const double len = d.length[ev.p0 + i];
            S.cnt[s] = d.count[ev.p0 + i];
            S.den[s] = ev.tandem ? fmax(WQ_EM_DSUB(len, d.readlen), 1.0) : len;
            S.prev[s] = 0.0; S.ct[s] = S.cnt[s];
        }
    }
    for (uint32_t s = tid; s <= nsl; s += c.T) S.poff[s] = 0;
    const uint32_t nw = (ev.np + 31u) >> 5;
    for (uint32_t i = tid; i < 32u * nw; i += c.T) S.psi[i] = 0.0;                  // including the padding of the last mask word
    for (uint32_t i = tid; i < ev.np; i += c.T) S.raw[i] = 0.0;
    for (uint32_t a = tid; a < ev.na; a += c.T) { S.mult[a] = d.amb_mult[ev.a0 + a]; S.psum[a] = 1.0; }
    const uint32_t nal = wq_cl_nsets(ev.na, c.r, c.C);
    for (uint32_t k = tid; k < nal * nw; k += c.T) S.lmask[k] = 0u;
    if (tid < 8) S.flags[tid] = 0;
}
// b: the members of the own sets as bit masks; how many sets every own path belongs to
WQ_HDN inline void wq_cl_init_b(const WqPsiDev& d, const WqPsiClEv& ev, const WqClCtx& c, uint32_t tid) {
    const WqPsiClSmem& S = c.S;
    const uint32_t nal = wq_cl_nsets(ev.na, c.r, c.C), nw = (ev.np + 31u) >> 5;
    for (uint32_t al = 0; al < nal; al++) {
        const uint32_t a = c.r + c.C * al;
        const uint32_t gb = d.amb_path_ptr[ev.a0 + a], n = d.amb_path_ptr[ev.a0 + a + 1] - gb;
        for (uint32_t k = tid; k < n; k += c.T) {
            const uint32_t m = d.amb_paths[gb + k];
#ifdef __CUDA_ARCH__
            atomicOr(&S.lmask[al * nw + (m >> 5)], 1u << (m & 31u));
#else
            S.lmask[al * nw + (m >> 5)] |= 1u << (m & 31u);
#endif
        }
    }
    for (uint32_t k = tid; k < ev.nm; k += c.T) {
        const uint32_t m = d.amb_paths[ev.m0 + k];
        if (wq_cl_owner(m, c.C) != c.r) continue;
#ifdef __CUDA_ARCH__
        atomicAdd(&S.poff[wq_cl_slot(m, c.C) + 1], 1u);
#else
        S.poff[wq_cl_slot(m, c.C) + 1] += 1u;
#endif
    }
}
// c (one thread): offsets of the inverted index
WQ_HDN inline void wq_cl_init_c(const WqPsiClEv& ev, const WqClCtx& c) {
    const WqPsiClSmem& S = c.S;
    const uint32_t nsl = wq_cl_nslots(ev.np, c.r, c.C);
    for (uint32_t s = 0; s < nsl; s++) { S.poff[s + 1] += S.poff[s]; S.cur[s] = S.poff[s]; }
}
// d, set by set with a block barrier in between: the own paths' entries in set order.  A path occurs at most once in a set
// (checked on the host before this kernel is chosen), so no two threads of one step share a cursor.
WQ_HDN inline void wq_cl_init_d(const WqPsiDev& d, const WqPsiClEv& ev, const WqClCtx& c, uint32_t tid, uint32_t a) {
    const WqPsiClSmem& S = c.S;
    const uint32_t gb = d.amb_path_ptr[ev.a0 + a], ge = d.amb_path_ptr[ev.a0 + a + 1];
    for (uint32_t k = gb + tid; k < ge; k += c.T) {
        const uint32_t m = d.amb_paths[k];
        if (wq_cl_owner(m, c.C) != c.r) continue;
        const uint32_t at = S.cur[wq_cl_slot(m, c.C)]++;
        S.lpset[at] = (uint16_t)a;
    }
}
// -- sweep.  (1) prop_sum of the own sets, to every block.  The reference adds a set's members in list order, and the lists are in
// ascending path order (checked on the host before this kernel is chosen), so a lane walks ITS set's bit mask over the paths in
// order and adds psi[i] where the bit is set (+0.0 elsewhere: an exact identity on these non-negative sums).  Every lane reads the same psi[i] at the same time -- a broadcast, one shared-memory
// wavefront -- where a gather through member lists cost a bank-conflicted access per member (measured: 26 clocks per member of the
// longest set, against ~11 for the dependent addition itself; the longest set of a large event names all paths but one, so the
// chain of ~np additions is the floor of this phase either way).
WQ_HDN inline void wq_cl_phase_psum(const WqPsiClEv& ev, const WqClCtx& c, uint32_t tid) {
    const WqPsiClSmem& S = c.S;
    const uint32_t nal = wq_cl_nsets(ev.na, c.r, c.C), nw = (ev.np + 31u) >> 5;
    for (uint32_t al = tid; al < nal; al += c.T) {
        double sm = 0.0;
        const uint32_t* mk = S.lmask + al * nw;
        for (uint32_t w = 0; w < nw; w++) {
            const uint32_t bits = mk[w];
            const double2* v2 = reinterpret_cast<const double2*>(S.psi + 32u * w);
            // a path outside the set adds +0.0, which leaves the sum as it is bit for bit (psi >= +0, the sum starts at +0.0, NaN stays NaN);
            // written as a select so that the 16 loads of a word are unconditional and issued ahead of the chain of additions
            #pragma unroll
            for (int b2 = 0; b2 < 16; b2++) {
                const double2 v = v2[b2];
                sm = WQ_EM_DADD(sm, (bits >> (2 * b2)) & 1u ? v.x : 0.0);
                sm = WQ_EM_DADD(sm, (bits >> (2 * b2 + 1)) & 1u ? v.y : 0.0);
            }
        }
        const uint32_t a = c.r + c.C * al;
        for (uint32_t rr = 0; rr < c.C; rr++) *wq_cl_remote(c, S.psum + a, rr) = sm;
    }
}
// calculate_psi!(inc, exc, counts) before spliced_em!: count / length of the own paths, to every block
WQ_HDN inline void wq_cl_phase_pre(const WqPsiClEv& ev, const WqClCtx& c, uint32_t tid) {
    const WqPsiClSmem& S = c.S;
    const uint32_t nsl = wq_cl_nslots(ev.np, c.r, c.C);
    for (uint32_t s = tid; s < nsl; s += c.T) {
        const uint32_t i = wq_cl_path(s, c.r, c.C);
        if (i >= ev.np) continue;
        const double rv = WQ_EM_DDIV(S.ct[s], S.den[s]);
        for (uint32_t rr = 0; rr < c.C; rr++) *wq_cl_remote(c, S.raw + i, rr) = rv;
    }
}
// (2) the own paths: count + shares in set order (first sweep: ones(n) / n over prop_sum = 1.0), count / length to every block
WQ_HDN inline void wq_cl_phase_shares(const WqPsiDev& d, const WqPsiClEv& ev, const WqClCtx& c, uint32_t tid, bool first) {
    const WqPsiClSmem& S = c.S;
    const uint32_t nsl = wq_cl_nslots(ev.np, c.r, c.C);
    for (uint32_t s = tid; s < nsl; s += c.T) {
        const uint32_t i = wq_cl_path(s, c.r, c.C);
        if (i >= ev.np) continue;
        double cc = S.cnt[s];
        const double pi = S.psi[i];
        const uint32_t kb = S.poff[s], ke = S.poff[s + 1];
        if (!first) {
            #pragma unroll 4
            for (uint32_t k = kb; k < ke; k++) { const uint32_t a = S.lpset[k]; cc = WQ_EM_DADD(cc, WQ_EM_DMUL(WQ_EM_DDIV(pi, S.psum[a]), S.mult[a])); }
        } else {
            for (uint32_t k = kb; k < ke; k++) {
                const uint32_t a = S.lpset[k];
                const uint32_t n = d.amb_path_ptr[ev.a0 + a + 1] - d.amb_path_ptr[ev.a0 + a];
                cc = WQ_EM_DADD(cc, WQ_EM_DMUL(WQ_EM_DDIV(1.0, (double)n), S.mult[a]));
            }
        }
        S.ct[s] = cc;
        S.prev[s] = pi;
        const double rv = WQ_EM_DDIV(cc, S.den[s]);
        for (uint32_t rr = 0; rr < c.C; rr++) *wq_cl_remote(c, S.raw + i, rr) = rv;
    }
}
// (3a) the total in the reference's order, by every block for itself: warp 0 the exclusion paths (tandem: all paths), warp 1 the inclusion paths
WQ_HDN inline void wq_cl_phase_total(const WqPsiClEv& ev, const WqClCtx& c, uint32_t tid) {
    const WqPsiClSmem& S = c.S;
    const uint32_t warp = tid >> 5;
    if (warp > 1) return;
    uint32_t lo, hi;
    if (ev.tandem) { if (warp == 1) return; lo = 0; hi = ev.np; }
    else if (warp == 0) { lo = ev.ni; hi = ev.np; }
    else { lo = 0; hi = ev.ni; }
    double t = 0.0;
    uint32_t i = lo;
    if (hi - i >= 8) {                                               // pipelined like prop_sum: the next eight values are on their way while these are added
        double va[8], vb[8];
        #pragma unroll
        for (int j = 0; j < 8; j++) va[j] = S.raw[i + j];
        i += 8;
        for (; hi - i >= 8; i += 8) {
            #pragma unroll
            for (int j = 0; j < 8; j++) vb[j] = S.raw[i + j];
            #pragma unroll
            for (int j = 0; j < 8; j++) t = WQ_EM_DADD(t, va[j]);
            #pragma unroll
            for (int j = 0; j < 8; j++) va[j] = vb[j];
        }
        #pragma unroll
        for (int j = 0; j < 8; j++) t = WQ_EM_DADD(t, va[j]);
    }
    for (; i < hi; i++) t = WQ_EM_DADD(t, S.raw[i]);
    if ((tid & 31u) == 0) S.tot[warp] = t;
}
// (3b) divsignif! of the own paths, to every block; returns whether one of this thread's paths changed
WQ_HDN inline bool wq_cl_phase_norm(const WqPsiDev& d, const WqPsiClEv& ev, const WqClCtx& c, uint32_t tid, int sig) {
    const WqPsiClSmem& S = c.S;
    const double total = ev.tandem ? S.tot[0] : WQ_EM_DADD(S.tot[0], S.tot[1]);
    const uint32_t nsl = wq_cl_nslots(ev.np, c.r, c.C);
    bool diff = false;
    for (uint32_t s = tid; s < nsl; s += c.T) {
        const uint32_t i = wq_cl_path(s, c.r, c.C);
        if (i >= ev.np) continue;
        double q = WQ_EM_DDIV(S.raw[i], total);
        if (sig > 0) q = wq_round_sig(d, q, sig);
        if (S.prev[s] != q) diff = true;
        for (uint32_t rr = 0; rr < c.C; rr++) *wq_cl_remote(c, S.psi + i, rr) = q;
    }
    return diff;
}
WQ_HDN inline void wq_cl_write_out(const WqPsiDev& d, const WqPsiClEv& ev, const WqClCtx& c, uint32_t tid, int64_t it) {
    const WqPsiClSmem& S = c.S;
    const uint32_t nsl = wq_cl_nslots(ev.np, c.r, c.C);
    for (uint32_t s = tid; s < nsl; s += c.T) {
        const uint32_t i = wq_cl_path(s, c.r, c.C);
        if (i >= ev.np) continue;
        d.psi[ev.p0 + i] = S.psi[i];
        d.ctemp[ev.p0 + i] = S.ct[s];
    }
    if (c.r == 0 && tid == 0) d.iterations[ev.e] = (int32_t)it;
}

// Phase (2) with L = 2 / 4 / 8 lanes per path (device only; the arithmetic and its order are those of wq_cl_phase_shares).  A block of a
// cluster owns a quarter or an eighth of the paths, fewer than it has threads, and what bounds the phase is the path that belongs to
// the most sets: its IEEE divisions run one after the other on one lane (the division's special-case branch keeps the compiler from
// overlapping them: ~125 clocks per set measured).  Here lane h of a group divides the shares k0 + h of every round of L sets, and
// every lane of the group adds the L quotients in set order (shuffles that name the group's lanes only, so the groups of a warp
// diverge freely); lane 0 writes.  MEASURED SLOWER than one lane per path (L = 2: 45.9 k against 18.9 k clocks per sweep, L = 4: 17.6 k
// against 16.2 k: the partial-mask shuffles and the groups' divergence cost more than the halved chain saves), so it is OFF unless
// WQ_PSI_CL_SPLIT=1; kept as a tested variant.
template <int L>
__device__ __forceinline__ void wq_cl_phase_shares_split(const WqPsiDev& d, const WqPsiClEv& ev, const WqClCtx& c, uint32_t tid, bool first) {
    const WqPsiClSmem& S = c.S;
    const uint32_t nsl = wq_cl_nslots(ev.np, c.r, c.C);
    const uint32_t lane = tid & 31u, h = lane & (uint32_t)(L - 1), base = lane & ~(uint32_t)(L - 1);
    const unsigned gmask = (L == 32 ? 0xFFFFFFFFu : ((1u << L) - 1u)) << base;
    for (uint32_t s = tid / L; s < nsl; s += c.T / L) {              // s is the same for the L lanes of a group
        const uint32_t i = wq_cl_path(s, c.r, c.C);
        if (i >= ev.np) continue;
        double cc = S.cnt[s];
        const double pi = S.psi[i];
        const uint32_t kb = S.poff[s], ke = S.poff[s + 1];
        for (uint32_t k0 = kb; k0 < ke; k0 += L) {
            const uint32_t k = k0 + h;
            double q = 0.0;
            if (k < ke) {
                const uint32_t a = S.lpset[k];
                if (!first) q = WQ_EM_DMUL(WQ_EM_DDIV(pi, S.psum[a]), S.mult[a]);
                else {
                    const uint32_t n = d.amb_path_ptr[ev.a0 + a + 1] - d.amb_path_ptr[ev.a0 + a];
                    q = WQ_EM_DMUL(WQ_EM_DDIV(1.0, (double)n), S.mult[a]);
                }
            }
            #pragma unroll
            for (int j = 0; j < L; j++) {
                const double qj = __shfl_sync(gmask, q, (int)base + j);
                if (k0 + (uint32_t)j < ke) cc = WQ_EM_DADD(cc, qj);
            }
        }
        if (h == 0) {
            S.ct[s] = cc;
            S.prev[s] = pi;
            const double rv = WQ_EM_DDIV(cc, S.den[s]);
            for (uint32_t rr = 0; rr < c.C; rr++) *wq_cl_remote(c, S.raw + i, rr) = rv;
        }
    }
}
__device__ __forceinline__ void wq_cl_phase_shares_dev(const WqPsiDev& d, const WqPsiClEv& ev, const WqClCtx& c, uint32_t tid, bool first, bool split) {
    const uint32_t nsl = wq_cl_nslots(ev.np, c.r, c.C);
    if (split && nsl * 8u <= c.T) wq_cl_phase_shares_split<8>(d, ev, c, tid, first);
    else if (split && nsl * 4u <= c.T) wq_cl_phase_shares_split<4>(d, ev, c, tid, first);
    else if (split && nsl * 2u <= c.T) wq_cl_phase_shares_split<2>(d, ev, c, tid, first);
    else wq_cl_phase_shares(d, ev, c, tid, first);
}

__global__ void __launch_bounds__(1024) wq_k_em_psi_cl(const WqPsiDev d, const uint32_t* __restrict__ list, const uint32_t n_list, const WqPsiClCfg z) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cl = cg::this_cluster();
    WqClCtx c;
    c.r = cl.block_rank(); c.C = z.C; c.T = blockDim.x; c.self = nullptr; c.all = nullptr;
    wq_psi_cl_layout(z, smem_raw, &c.S);
    const uint32_t tid = threadIdx.x;
    const uint32_t ncl = gridDim.x / z.C, cid = blockIdx.x / z.C;
    const bool timing = z.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0;
    auto lap = [&](int k) { if (timing) { const long long t1 = clock64(); tk[k] += t1 - t0; t0 = t1; } };
    for (uint32_t slot = cid; slot < n_list; slot += ncl) {          // every block of a cluster walks the same events
        const WqPsiClEv ev = wq_psi_cl_event(d, list[slot]);
        if (timing) t0 = clock64();
        wq_cl_init_a(d, ev, c, tid);
        __syncthreads();
        wq_cl_init_b(d, ev, c, tid);
        __syncthreads();
        if (tid == 0) wq_cl_init_c(ev, c);
        __syncthreads();
        for (uint32_t a = 0; a < ev.na; a++) { wq_cl_init_d(d, ev, c, tid, a); __syncthreads(); }
        cl.sync();
        lap(0);
        bool differs = false;
        auto norm = [&](int sig) {
            const bool df = wq_cl_phase_norm(d, ev, c, tid, sig);
            const int any = __syncthreads_or(df ? 1 : 0);
            if (tid == 0) for (uint32_t rr = 0; rr < c.C; rr++) *wq_cl_remote(c, c.S.flags + c.r, rr) = (uint32_t)(any != 0);
            cl.sync();
            uint32_t f = 0;
            for (uint32_t rr = 0; rr < c.C; rr++) f |= c.S.flags[rr];
            differs = f != 0;
        };
        if (!ev.tandem) {
            wq_cl_phase_pre(ev, c, tid);
            cl.sync();
            wq_cl_phase_total(ev, c, tid);
            __syncthreads();
            norm(0);
        }
        int64_t it = 1;
        for (;;) {
            if (!ev.tandem && !(differs && it < d.maxit)) break;
            if (timing) t0 = clock64();
            if (it > 1) { wq_cl_phase_psum(ev, c, tid); lap(1); cl.sync(); lap(2); }
            wq_cl_phase_shares_dev(d, ev, c, tid, it == 1, z.split != 0);
            lap(3);
            cl.sync();
            lap(4);
            wq_cl_phase_total(ev, c, tid);
            __syncthreads();
            lap(5);
            norm(d.sig);
            lap(6);
            if (timing) tk[7] += 1;                                  // slot 7 counts sweeps
            if (ev.tandem) { if (differs && it < d.maxit) it += 1; else break; }
            else it += 1;
        }
        wq_cl_write_out(d, ev, c, tid, it);
        cl.sync();                                                   // no block rebuilds its arrays while another still reads them
    }
    if (timing) for (int k = 0; k < 8; k++) z.dbg[k] = tk[k];
}

// Events with at most 4 paths (more than 90 % of them), at most WQ_PSI4_SETS ambiguous sets of at most 4 members: 4 lanes per
// event, 8 events per warp.  Lane i of the group OWNS path i (psi / previous psi / temp count in registers); once per sweep the
// four psi values are exchanged with four shuffles, after which every lane forms every set's prop_sum from registers (the same
// sequential additions, redundantly) and adds its own shares; the set descriptors (member list as 4-bit path indices, count,
// multiplier) sit in shared memory, written once per event.  The divisions of up to four sets are issued together (they do not
// depend on each other; only the additions into the temp count are chained, in set order).  Every loop has a group-uniform trip
// count and every shuffle / vote names the group's lanes only, so the groups of a warp diverge freely.
#define WQ_PSI4_SETS 12
__global__ void __launch_bounds__(128) wq_k_em_psi4(const WqPsiDev d, const uint32_t* __restrict__ list, const uint32_t n_list) {
    __shared__ double s_mult[32][WQ_PSI4_SETS];
    __shared__ uint32_t s_desc[32][WQ_PSI4_SETS];      // bits 0..15: four 4-bit path indices in member order; bits 16..18: member count
    const uint32_t lane = threadIdx.x & 31, sub = lane & 3, base = lane - sub, grp = threadIdx.x >> 2;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    if (slot >= n_list) return;
    const unsigned gmask = 0xFu << base;
    const uint32_t e = list[slot];
    const uint32_t p0 = d.path_ptr[e], np = d.path_ptr[e + 1] - p0, ni = d.n_inc[e];
    const uint32_t a0 = d.amb_ptr[e], na = d.amb_ptr[e + 1] - a0;
    const bool tandem = d.is_tandem[e] != 0;
    const bool own = sub < np;
    const double cnt = own ? d.count[p0 + sub] : 0.0;
    const double len = own ? d.length[p0 + sub] : 1.0;
    const double den = tandem ? fmax(__dsub_rn(len, d.readlen), 1.0) : len;
    for (uint32_t a = sub; a < na; a += 4) {
        const uint32_t b = d.amb_path_ptr[a0 + a], n = d.amb_path_ptr[a0 + a + 1] - b;
        uint32_t desc = n << 16;
        for (uint32_t k = 0; k < n; k++) desc |= (d.amb_paths[b + k] & 15u) << (4 * k);
        s_desc[grp][a] = desc;
        s_mult[grp][a] = d.amb_mult[a0 + a];
    }
    __syncwarp(gmask);
    double psi = 0.0, prev = 0.0, ct = cnt;
    auto normalise = [&](int sig) {
        const double q = __ddiv_rn(ct, den);
        const double q0 = __shfl_sync(gmask, q, base), q1 = __shfl_sync(gmask, q, base + 1), q2 = __shfl_sync(gmask, q, base + 2), q3 = __shfl_sync(gmask, q, base + 3);
        double total;
        if (tandem) {
            total = __dadd_rn(0.0, q0);
            if (np > 1) total = __dadd_rn(total, q1);
            if (np > 2) total = __dadd_rn(total, q2);
            if (np > 3) total = __dadd_rn(total, q3);
        } else {
            double se = 0.0, si = 0.0;
            if (ni <= 0 && np > 0) se = __dadd_rn(se, q0);
            if (ni <= 1 && np > 1) se = __dadd_rn(se, q1);
            if (ni <= 2 && np > 2) se = __dadd_rn(se, q2);
            if (ni <= 3 && np > 3) se = __dadd_rn(se, q3);
            if (ni > 0) si = __dadd_rn(si, q0);
            if (ni > 1) si = __dadd_rn(si, q1);
            if (ni > 2) si = __dadd_rn(si, q2);
            if (ni > 3) si = __dadd_rn(si, q3);
            total = __dadd_rn(se, si);
        }
        const double r = __ddiv_rn(q, total);
        psi = sig > 0 ? wq_round_sig(d, r, sig) : r;
    };
    auto differs = [&]() { return __any_sync(gmask, own && prev != psi) != 0; };
    if (!tandem) normalise(0);                      // calculate_psi!(inc, exc, counts) before spliced_em!
    int64_t it = 1;
    for (;;) {
        if (!tandem && !(differs() && it < d.maxit)) break;
        ct = cnt; prev = psi;
        const double v0 = __shfl_sync(gmask, psi, base), v1 = __shfl_sync(gmask, psi, base + 1), v2 = __shfl_sync(gmask, psi, base + 2), v3 = __shfl_sync(gmask, psi, base + 3);
        for (uint32_t ab = 0; ab < na; ab += 4) {
            double share[4];
            uint32_t member_of = 0;                                                              // bit u: this lane's path is in set ab + u
            #pragma unroll
            for (int u = 0; u < 4; u++) {
                share[u] = 0.0;
                const uint32_t a = ab + u;
                if (a < na) {
                    const uint32_t desc = s_desc[grp][a], n = desc >> 16;
                    double psum = 1.0;
                    bool mine = false;
                    if (it > 1) psum = 0.0;
                    #pragma unroll
                    for (uint32_t k = 0; k < 4; k++) {
                        if (k < n) {
                            const uint32_t pth = (desc >> (4 * k)) & 15u;
                            if (it > 1) psum = __dadd_rn(psum, pth == 0 ? v0 : pth == 1 ? v1 : pth == 2 ? v2 : v3);
                            if (pth == sub) mine = true;
                        }
                    }
                    // first sweep: ones(n)/n over prop_sum = 1.0; later sweeps: current psi over their sum (src/events.jl:869-886)
                    if (mine) { share[u] = __dmul_rn(__ddiv_rn(it > 1 ? psi : __ddiv_rn(1.0, (double)n), psum), s_mult[grp][a]); member_of |= 1u << u; }
                }
            }
            #pragma unroll
            for (int u = 0; u < 4; u++) if (member_of & (1u << u)) ct = __dadd_rn(ct, share[u]);
        }
        normalise(d.sig);
        if (tandem) { if (differs() && it < d.maxit) it += 1; else break; }
        else it += 1;
    }
    if (own) { d.psi[p0 + sub] = psi; d.ctemp[p0 + sub] = ct; }
    if (sub == 0) d.iterations[e] = (int32_t)it;
}

// `wait` = false leaves the launch and its result copies in flight on `stream` (the result buffers must then be pinned
// host memory and stay alive until the stream has been synchronised)
// One pass over the member lists of the given events for the cluster kernel: are they all in ascending path order (the kernel walks a
// set's members as a bit mask in path order, which is the reference's order of additions only then; the event construction writes
// them that way: inclusion paths, then exclusion paths, each in order; ascending also means no path twice -- the inverted-index build
// relies on that -- and every member inside the event), and how many inverted-index entries the fullest block holds with 8 / 4 / 2
// blocks per event (a path's owner with C blocks is its owner with 8 blocks modulo C).
struct WqPsiClScan { bool ascending = true; uint32_t inv_cap[3] = {0, 0, 0}; };
static inline WqPsiClScan wq_psi_cl_scan(const wq_psi_batch& b, const uint32_t* events, uint32_t n) {
    WqPsiClScan sc;
    for (uint32_t k = 0; k < n; k++) {
        const uint32_t e = events[k], np = b.path_ptr[e + 1] - b.path_ptr[e];
        uint32_t own8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t a = b.amb_ptr[e]; a < b.amb_ptr[e + 1]; a++) {
            const uint32_t jb = b.amb_path_ptr[a], je = b.amb_path_ptr[a + 1];
            for (uint32_t j = jb; j < je; j++) {
                const uint32_t m = b.amb_paths[j];
                if (m >= np || (j > jb && m <= b.amb_paths[j - 1])) sc.ascending = false;
                own8[(m >> 5) & 7u]++;
            }
        }
        for (uint32_t r = 0; r < 8; r++) sc.inv_cap[0] = std::max(sc.inv_cap[0], own8[r]);
        for (uint32_t r = 0; r < 4; r++) sc.inv_cap[1] = std::max(sc.inv_cap[1], own8[r] + own8[r + 4]);
        for (uint32_t r = 0; r < 2; r++) sc.inv_cap[2] = std::max(sc.inv_cap[2], own8[r] + own8[r + 2] + own8[r + 4] + own8[r + 6]);
    }
    return sc;
}
// capacities of the cluster kernel's shared-memory layout over the given events; returns the bytes one block needs
static inline size_t wq_psi_cl_config(const wq_psi_batch& b, const uint32_t* events, uint32_t n, uint32_t C, WqPsiClCfg& z, const WqPsiClScan* scan = nullptr) {
    WqPsiClScan local;
    if (!scan) { local = wq_psi_cl_scan(b, events, n); scan = &local; }
    memset(&z, 0, sizeof z);
    z.C = C;
    z.inv_cap = scan->inv_cap[C == 8 ? 0 : C == 4 ? 1 : 2];
    for (uint32_t k = 0; k < n; k++) {
        const uint32_t e = events[k];
        const uint32_t np = b.path_ptr[e + 1] - b.path_ptr[e], na = b.amb_ptr[e + 1] - b.amb_ptr[e];
        z.np_cap = std::max(z.np_cap, np); z.na_cap = std::max(z.na_cap, na);
        z.npl_cap = std::max(z.npl_cap, wq_cl_nslots(np, 0, C)); z.nal_cap = std::max(z.nal_cap, wq_cl_nsets(na, 0, C));
    }
    return wq_psi_cl_layout(z, nullptr, nullptr);
}
static inline bool wq_psi_cl_unique(const wq_psi_batch& b, const uint32_t* events, uint32_t n) { return wq_psi_cl_scan(b, events, n).ascending; }

static inline std::string wq_run_em_psi(wq_psi_batch& b, WqEmScratch& M, cudaStream_t stream, uint64_t* launches, bool wait = true) {
    const uint32_t E = b.n_events;
    if (E == 0) return "";
    if (!b.path_ptr || !b.n_inc || !b.amb_ptr || !b.amb_path_ptr || !b.is_tandem || !b.psi || !b.iterations) return "wq_psi_batch has null arrays";
    const uint32_t P = b.path_ptr[E], A = b.amb_ptr[E], AP = b.amb_path_ptr[A];
    WqPsiDev d;
    d.n_events = E; d.readlen = (double)b.readlen; d.sig = (int)b.sig; d.maxit = b.maxit;
    for (int k = 0; k <= 22; k++) { d.p10[k] = std::pow(10.0, (double)k); d.ip10[k] = 1.0 / d.p10[k]; }
    WQ_EMCK(wq_p10_table(M, stream, &d.p10big));
    uint32_t *d_pp, *d_ni, *d_ap, *d_app, *d_apaths;
    double *d_cnt, *d_len, *d_mult, *d_psi, *d_prev, *d_ct, *d_prop;
    uint8_t* d_tan;
    int32_t* d_it;
    auto up = [&](int slot, const void* src, size_t bytes, void** out) -> cudaError_t {
        uint8_t* p;
        cudaError_t e = M.get<uint8_t>(slot, bytes, &p);
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, stream);
        *out = p;
        return e;
    };
    WQ_EMCK(up(0, b.path_ptr, (E + 1) * 4ull, (void**)&d_pp)); WQ_EMCK(up(1, b.n_inc, E * 4ull, (void**)&d_ni));
    WQ_EMCK(up(2, b.amb_ptr, (E + 1) * 4ull, (void**)&d_ap)); WQ_EMCK(up(3, b.amb_path_ptr, (A + 1) * 4ull, (void**)&d_app));
    WQ_EMCK(up(4, b.amb_paths, AP * 4ull, (void**)&d_apaths)); WQ_EMCK(up(5, b.path_count, P * 8ull, (void**)&d_cnt));
    WQ_EMCK(up(6, b.path_length, P * 8ull, (void**)&d_len)); WQ_EMCK(up(7, b.amb_mult, A * 8ull, (void**)&d_mult));
    WQ_EMCK(up(8, b.is_tandem, E, (void**)&d_tan));
    WQ_EMCK(M.get(9, P, &d_psi)); WQ_EMCK(M.get(10, P, &d_prev)); WQ_EMCK(M.get(11, P, &d_ct)); WQ_EMCK(M.get(12, AP, &d_prop));
    WQ_EMCK(M.get(13, E, &d_it));
    d.path_ptr = d_pp; d.n_inc = d_ni; d.count = d_cnt; d.length = d_len; d.amb_ptr = d_ap; d.amb_path_ptr = d_app;
    d.amb_paths = d_apaths; d.amb_mult = d_mult; d.is_tandem = d_tan; d.psi = d_psi; d.prev = d_prev; d.ctemp = d_ct;
    d.amb_prop = d_prop; d.iterations = d_it;
    // Launch classes: 0 = at most 4 paths with small sets (4 lanes per event); 1..4 = one block per event, grouped by the shared memory the
    // event needs (4 / 16 / 64 / 200 KB, with 32 / 64 / 128 / 256 threads: more threads shorten the share divisions of a large event);
    // 5 / 6 = events too large for that: per-path arrays (5: and per-set arrays) in shared memory, member lists in a global scratch, 1024
    // threads.  The classes run concurrently on side streams (a handful of large events that need all 1500 sweeps would otherwise add
    // their whole latency to the batch).
    static const size_t kBlockSmem[4] = {4 * 1024, 16 * 1024, 64 * 1024, 200 * 1024};
    static unsigned kBlockThreads[4] = {32, 64, 128, 256};
    static bool threads_env = false;
    if (!threads_env) {                                       // experiment knob: WQ_PSI_THREADS="32,64,128,256"
        threads_env = true;
        if (const char* ev = getenv("WQ_PSI_THREADS")) {
            unsigned v[4];
            if (sscanf(ev, "%u,%u,%u,%u", &v[0], &v[1], &v[2], &v[3]) == 4)
                for (int k = 0; k < 4; k++) if (v[k] >= 32 && v[k] <= 1024 && v[k] % 32 == 0) kBlockThreads[k] = v[k];
        }
    }
    const int NC = 7;
    std::vector<uint32_t> order(E);
    std::vector<uint8_t> cls(E);
    uint32_t nclass[NC] = {0}, cbase[NC + 1] = {0};
    auto cls_of = [&](uint32_t e) -> int {
        const uint32_t np = b.path_ptr[e + 1] - b.path_ptr[e];
        const uint32_t a_lo = b.amb_ptr[e], a_hi = b.amb_ptr[e + 1], na = a_hi - a_lo, nm = b.amb_path_ptr[a_hi] - b.amb_path_ptr[a_lo];
        if (np <= 4 && na <= WQ_PSI4_SETS) {
            bool small = true;
            for (uint32_t a = a_lo; a < a_hi && small; a++) if (b.amb_path_ptr[a + 1] - b.amb_path_ptr[a] > 4) small = false;
            if (small) return 0;
        }
        if (np > 65535 || na > 65535) return -1;
        const size_t need = wq_psi_block_smem(np, na, nm);
        for (int k = 0; k < 4; k++) if (need <= kBlockSmem[k]) return 1 + k;
        if (wq_psi_big_smem(np, na, true) <= kBlockSmem[3]) return 5;
        if (wq_psi_big_smem(np, na, false) <= kBlockSmem[3]) return 6;
        return -1;
    };
    size_t big_smem[2] = {0, 0};
    for (uint32_t e = 0; e < E; e++) {
        const int c = cls_of(e);
        if (c < 0) return "PSI event too large for the device kernel (more than ~4600 paths or 65535 ambiguous sets)";
        if (c >= 5) big_smem[c - 5] = std::max(big_smem[c - 5], wq_psi_big_smem(b.path_ptr[e + 1] - b.path_ptr[e], b.amb_ptr[e + 1] - b.amb_ptr[e], c == 5));
        cls[e] = (uint8_t)c;
        nclass[c]++;
    }
    for (int c = 0; c < NC; c++) cbase[c + 1] = cbase[c] + nclass[c];
    { uint32_t cur[NC]; for (int c = 0; c < NC; c++) cur[c] = cbase[c]; for (uint32_t e = 0; e < E; e++) order[cur[cls[e]]++] = e; }
    uint32_t* d_list;
    WQ_EMCK(up(14, order.data(), E * 4ull, (void**)&d_list));
    WqPsiBigCfg bigc[2];
    memset(bigc, 0, sizeof bigc);
    WqPsiBigCfg staged;
    memset(&staged, 0, sizeof staged);
    const uint32_t nbig = nclass[5] + nclass[6];
    // Large events on clusters of C = 8 / 4 / 2 blocks (wq_k_em_psi_cl) or on the block kernel, whichever needs the least time for
    // this many events (see below).  WQ_PSI_CLUSTER = 0 / 2 / 4 / 8 overrides (0: never).
    WqPsiClCfg clz;
    memset(&clz, 0, sizeof clz);
    size_t cl_smem = 0;
    uint32_t cl_clusters = 0;
    if (nbig) {
        const int cl_env = getenv("WQ_PSI_CLUSTER") ? atoi(getenv("WQ_PSI_CLUSTER")) : -1;      // read per call: tests switch it
        static bool cl_attr_set = false;
        const size_t smem_max = 227 * 1024;
        const WqPsiClScan scan = cl_env != 0 ? wq_psi_cl_scan(b, order.data() + cbase[5], nbig) : WqPsiClScan();
        const bool unique = scan.ascending;
        auto config = [&](uint32_t C, WqPsiClCfg& z) -> size_t { return wq_psi_cl_config(b, order.data() + cbase[5], nbig, C, z, &scan); };
        auto resident = [&](uint32_t C, size_t smem) -> uint32_t {       // clusters of C blocks the device holds at once
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3(C * nbig); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, wq_k_em_psi_cl, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
            return n > 0 ? (uint32_t)n : 0u;
        };
        if (unique && cl_env != 0) {
            if (!cl_attr_set) {
                if (cudaFuncSetAttribute(wq_k_em_psi_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess) cl_attr_set = true;
                else cudaGetLastError();
            }
            // the cheapest of: C = 8 / 4 / 2 blocks per event in rounds of the clusters the device holds at once, or the block kernel in
            // rounds of one event per SM.  Relative cost of one round, measured on 24 events of 1000-1900 paths (profiles/r02_psi_large_events.txt):
            // block kernel 272 ms, C = 2: 55 ms, C = 4: 46 ms, C = 8: ~35 ms.
            int sms = 148;
            { int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; } }
            const uint32_t tryC[3] = {8, 4, 2};
            const double round_cost[3] = {0.75, 1.0, 1.2};
            // (that ratio holds for events of ~100 k set members; for the smaller ones of these classes -- a few hundred paths, ~15 k members --
            // nothing was measured, and the block kernel keeps them unless the clusters need no more rounds than it does)
            uint64_t members = 0;
            for (uint32_t k = 0; k < nbig; k++) { const uint32_t e = order[cbase[5] + k]; members += b.amb_path_ptr[b.amb_ptr[e + 1]] - b.amb_path_ptr[b.amb_ptr[e]]; }
            const double block_round = members / nbig >= 30000 ? 5.9 : 1.5;
            double best = cl_env > 0 ? 1e300 : block_round * (double)((nbig + (uint32_t)sms - 1) / (uint32_t)sms);
            for (int t = 0; t < 3 && cl_attr_set; t++) {
                const uint32_t C = tryC[t];
                if (cl_env > 0 && C != (uint32_t)cl_env) continue;
                WqPsiClCfg z;
                const size_t need = config(C, z);
                if (need > smem_max) continue;
                const uint32_t res = resident(C, need);
                if (!res) continue;
                const double cost = round_cost[t] * (double)((nbig + res - 1) / res);
                if (cost < best) {
                    best = cost; clz = z; cl_smem = need; cl_clusters = std::min(nbig, res);
                    clz.split = getenv("WQ_PSI_CL_SPLIT") ? (uint32_t)(atoi(getenv("WQ_PSI_CL_SPLIT")) != 0) : 0u;
                }
            }
        }
    }
    if (nbig && !cl_clusters) {                                               // global scratch of the large events (classes 5 and 6 are adjacent in `order`): members and sets back to back
        std::vector<uint64_t> base(2ull * nbig);
        uint64_t mtot = 0, atot = 0;
        for (uint32_t k = 0; k < nbig; k++) {
            const uint32_t e = order[cbase[5] + k];
            const uint32_t a_lo = b.amb_ptr[e], a_hi = b.amb_ptr[e + 1];
            base[k] = mtot; base[nbig + k] = atot;
            mtot += b.amb_path_ptr[a_hi] - b.amb_path_ptr[a_lo];
            atot += a_hi - a_lo;
        }
        uint64_t* d_base;
        WQ_EMCK(up(24, base.data(), base.size() * 8ull, (void**)&d_base));
        uint8_t* scratch;
        const size_t m8 = (mtot + 3) & ~3ull, a8 = (atot + nbig + 1) & ~1ull;
        WQ_EMCK(M.get<uint8_t>(25, 2 * 2 * m8 + 16 * atot + 4 * a8 + 64, &scratch));
        WqPsiBigCfg g;
        memset(&g, 0, sizeof g);
        g.psum = reinterpret_cast<double*>(scratch);
        g.mult = g.psum + atot;
        g.moff = reinterpret_cast<uint32_t*>(g.mult + atot);
        g.mem = reinterpret_cast<uint16_t*>(g.moff + a8);
        g.pset = g.mem + m8;
        for (int k = 0; k < 2; k++) {                         // class 5 uses slots [0, nclass[5]), class 6 the ones behind
            bigc[k] = g;
            bigc[k].mbase = d_base + (k ? nclass[5] : 0);
            bigc[k].abase = d_base + nbig + (k ? nclass[5] : 0);
            bigc[k].sets_smem = k == 0 ? 1u : 0u;
        }
        // class 6's moff base: big.moff + ab + slot with slot counted inside its own launch; shift so the regions do not overlap with class 5's
        bigc[1].moff = g.moff + nclass[5];
    }
    WQ_EMCK(M.side_streams());
    WQ_EMCK(M.timing_events());
    static bool smem_attr_set = false;
    if (!smem_attr_set) {
        WQ_EMCK(cudaFuncSetAttribute(wq_k_em_psi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlockSmem[3]));
        smem_attr_set = true;
    }
    WQ_EMCK(cudaEventRecord(M.tk0, stream));
    WQ_EMCK(cudaEventRecord(M.fork, stream));
    int used = 0;
    for (int c = NC - 1; c >= 0; c--) {             // the large events first: their sweeps are the long pole
        if (!nclass[c]) continue;
        if (c == 5 && cl_clusters && nclass[6]) continue;         // went out with class 6's cluster launch
        cudaStream_t ss = used == 0 ? stream : M.side[(used - 1) % 3];
        if (used > 0) WQ_EMCK(cudaStreamWaitEvent(ss, M.fork, 0));
        const uint32_t n = nclass[c];
        const uint32_t* lst = d_list + cbase[c];
        if (c == 0) wq_k_em_psi4<<<(n + 31) / 32, 128, 0, ss>>>(d, lst, n);
        else if (c >= 5 && cl_clusters) {                     // classes 5 and 6 are adjacent in `order`: one cluster launch covers both
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3(clz.C * cl_clusters); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = cl_smem; cfg.stream = ss;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = clz.C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            const bool cl_timing = getenv("WQ_PSI_CL_TIMING") != nullptr;      // diagnostic: per-phase clocks of cluster 0, block 0; waits for the launch
            if (cl_timing) { WQ_EMCK(M.get<long long>(24, 8, &clz.dbg)); WQ_EMCK(cudaMemsetAsync(clz.dbg, 0, 64, ss)); }
            WQ_EMCK(cudaLaunchKernelEx(&cfg, wq_k_em_psi_cl, d, (const uint32_t*)(d_list + cbase[5]), nbig, clz));
            if (cl_timing) {
                long long tk[8];
                WQ_EMCK(cudaStreamSynchronize(ss));
                WQ_EMCK(cudaMemcpy(tk, clz.dbg, 64, cudaMemcpyDeviceToHost));
                const double sw = tk[7] > 0 ? (double)tk[7] : 1.0;
                fprintf(stderr, "[wq]       psi cluster kernel (%u events, %u clusters of %u): build %.0f clocks per event-walk; per sweep of cluster 0 (%lld sweeps): prop_sum %.0f, barrier %.0f, shares %.0f, "
                        "barrier %.0f, total %.0f, normalise + barrier %.0f clocks\n", nbig, cl_clusters, clz.C, (double)tk[0], tk[7], tk[1] / sw, tk[2] / sw, tk[3] / sw, tk[4] / sw, tk[5] / sw, tk[6] / sw);
            }
        }
        else if (c >= 5) wq_k_em_psi<<<n, 1024, big_smem[c - 5], ss>>>(d, lst, n, bigc[c - 5]);
        else wq_k_em_psi<<<n, kBlockThreads[c - 1], kBlockSmem[c - 1], ss>>>(d, lst, n, staged);
        if (launches) *launches += 1;
        WQ_EMCK(cudaGetLastError());
        static const bool class_timing = getenv("WQ_PSI_CLASS_TIMING") != nullptr;     // diagnostic: serialises the classes
        if (class_timing) {
            const auto t0 = std::chrono::steady_clock::now();
            WQ_EMCK(cudaStreamSynchronize(ss));
            uint64_t np_max = 0, sweeps = 0, capped = 0;
            std::vector<int32_t> its(E);
            WQ_EMCK(cudaMemcpy(its.data(), d_it, E * 4ull, cudaMemcpyDeviceToHost));
            for (uint32_t k = 0; k < n; k++) {
                const uint32_t e = order[cbase[c] + k];
                np_max = std::max<uint64_t>(np_max, b.path_ptr[e + 1] - b.path_ptr[e]);
                sweeps += (uint64_t)its[e]; if (its[e] >= b.maxit) capped++;
            }
            if (c >= 5 && cl_clusters) fprintf(stderr, "[wq]       psi classes 5+6 on %u clusters of %u blocks (%zu B shared memory each), %u events; next line: class %d's share\n", cl_clusters, clz.C, cl_smem, nbig, c);
            fprintf(stderr, "[wq]       psi class %d: %u events (max %llu paths, %llu sweeps, %llu at the cap) %9.3f ms\n", c, n, (unsigned long long)np_max,
                    (unsigned long long)sweeps, (unsigned long long)capped, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        }
        if (used > 0) { WQ_EMCK(cudaEventRecord(M.join[(used - 1) % 3], ss)); WQ_EMCK(cudaStreamWaitEvent(stream, M.join[(used - 1) % 3], 0)); }
        used++;
    }
    WQ_EMCK(cudaEventRecord(M.tk1, stream));
    M.timed = true;
    WQ_EMCK(cudaMemcpyAsync(b.psi, d_psi, P * 8ull, cudaMemcpyDeviceToHost, stream));
    if (b.path_total) WQ_EMCK(cudaMemcpyAsync(b.path_total, d_ct, P * 8ull, cudaMemcpyDeviceToHost, stream));
    WQ_EMCK(cudaMemcpyAsync(b.iterations, d_it, E * 4ull, cudaMemcpyDeviceToHost, stream));
    if (wait) WQ_EMCK(cudaStreamSynchronize(stream));
    return "";
}


// ------------------------------------------------------------------------------------------ Beta posterior CI
// beta_posterior_ci(psi, total; ci = 0.9, sig = 3), src/events.jl:816-828 (call site :790): the 5 % and 95 % quantiles of
// Beta(psi * n + 1, (1 - psi) * n + 1), rounded to 3 significant digits.  The reference gets the quantile from
// Distributions 0.24.15 -> Rmath qbeta, which is not available here (parity unpinned); this is the regularised incomplete
// beta function by Lentz's continued fraction and a bracketed Newton iteration on it, carried to ~1e-14 relative, so the
// 3-digit rounded values agree with any accurate quantile except within ~1e-11 of a rounding boundary.
// One thread per quantile (2 per event): independent Float64 work, a few thousand flops each.
#define WQ_CI_GL 32
struct WqCiDev { uint32_t n; const double* psi; const double* total; double* lo; double* hi; WqPsiDev tab;
                 double gl_y[WQ_CI_GL], gl_w[WQ_CI_GL]; };      // Gauss-Legendre nodes / weights on [0, 1]

__device__ inline double wq_betacf(double a, double b, double x) {          // continued fraction of I_x(a, b), modified Lentz
    const double tiny = 1e-300, eps = 1e-14;     // (machine epsilon is 2.2e-16: a smaller bound may never be met)
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 10000; m++) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return h;
}
// I_x(a, b) for a, b > 3000 by Gauss-Legendre quadrature of the density between x and a point ~10 standard deviations
// beyond the mean (the continued fraction needs O(sqrt(max(a, b))) terms there); relative error of the resulting
// quantile below 2e-10 (checked against scipy.special.betaincinv over n = 1e4 .. 1e7)
__device__ inline double wq_betainc_quad(const WqCiDev& d, double a, double b, double x, double lbeta) {
    const double a1 = a - 1.0, b1 = b - 1.0, mu = a / (a + b), lnmu = log(mu), lnmuc = log1p(-mu);
    const double t = sqrt(a * b / ((a + b) * (a + b) * (a + b + 1.0)));
    double xu;
    if (x > mu) { if (x >= 1.0) return 1.0; xu = fmin(1.0, fmax(mu + 10.0 * t, x + 5.0 * t)); }
    else { if (x <= 0.0) return 0.0; xu = fmax(0.0, fmin(mu - 10.0 * t, x - 5.0 * t)); }
    double sum = 0.0;
    for (int j = 0; j < WQ_CI_GL; j++) {
        const double tt = x + (xu - x) * d.gl_y[j];
        sum += d.gl_w[j] * exp(a1 * (log(tt) - lnmu) + b1 * (log1p(-tt) - lnmuc));
    }
    const double ans = sum * (xu - x) * exp(a1 * lnmu + b1 * lnmuc - lbeta);
    return ans > 0.0 ? 1.0 - ans : -ans;
}
// I_x(a, b) and the density at x
__device__ inline double wq_betainc(const WqCiDev& d, double a, double b, double x, double lbeta, double* pdf) {
    if (x <= 0.0) { *pdf = 0.0; return 0.0; }
    if (x >= 1.0) { *pdf = 0.0; return 1.0; }
    const double lfront = a * log(x) + b * log1p(-x) - lbeta;
    *pdf = exp(lfront - log(x) - log1p(-x));
    if (a > 3000.0 && b > 3000.0) return wq_betainc_quad(d, a, b, x, lbeta);
    const double front = exp(lfront);
    if (x < (a + 1.0) / (a + b + 2.0)) return front * wq_betacf(a, b, x) / a;
    return 1.0 - front * wq_betacf(b, a, 1.0 - x) / b;
}
__device__ inline double wq_beta_quantile(const WqCiDev& d, double a, double b, double q, double z) {
    const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
    double lo = 0.0, hi = 1.0;
    // start at the normal approximation of the quantile (the continued fraction is slowest at the mean)
    double x = a / (a + b) + z * sqrt(a * b / ((a + b) * (a + b) * (a + b + 1.0)));
    if (!(x > 0.0)) x = 0.5 * (a / (a + b));
    if (!(x < 1.0)) x = 0.5 * (1.0 + a / (a + b));
    for (int it = 0; it < 200; it++) {
        double pdf;
        const double f = wq_betainc(d, a, b, x, lbeta, &pdf) - q;
        if (f > 0.0) hi = x; else lo = x;
        double nx = (pdf > 0.0 && isfinite(pdf)) ? x - f / pdf : 0.5 * (lo + hi);
        if (!(nx >= lo && nx <= hi)) nx = 0.5 * (lo + hi);          // inclusive: a step too small to move x is convergence
        if (fabs(nx - x) <= 1e-13 * fabs(x) || hi - lo <= 1e-15 * hi) { x = nx; break; }
        x = nx;
    }
    return x;
}
__global__ void __launch_bounds__(128) wq_k_beta_ci(const WqCiDev d) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * d.n) return;
    const uint32_t e = t >> 1;
    const double p = d.psi[e], n = d.total[e];
    const double a = __dadd_rn(__dmul_rn(p, n), 1.0), b = __dadd_rn(__dmul_rn(__dsub_rn(1.0, p), n), 1.0);
    const double q = (t & 1) ? 1.0 - (1.0 - 0.9) / 2 : (1.0 - 0.9) / 2;       // lo_q = (1 - ci) / 2, hi_q = 1 - lo_q
    const double v = wq_round_sig(d.tab, wq_beta_quantile(d, a, b, q, (t & 1) ? 1.6448536269514722 : -1.6448536269514722), 3);
    if (t & 1) d.hi[e] = v; else d.lo[e] = v;
}

// `wait` = false leaves the launch and its result copies in flight (lo / hi must then be pinned host memory)
// CI of the events that needed EM, on the stream of their PSI launch: psi = sum of the inclusion paths' psi in path order
// (what the host computes for the record, src/events.jl:734-741), then the same quantiles.  total < 0 marks problems
// without a CI (tandem UTR events).
__global__ void __launch_bounds__(128) wq_k_psi_ci(const WqCiDev d, const uint32_t* __restrict__ path_ptr, const uint32_t* __restrict__ n_inc,
                                                   const double* __restrict__ path_psi) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * d.n) return;
    const uint32_t e = t >> 1;
    const double n = d.total[e];
    double p = 0.0;
    for (uint32_t i = path_ptr[e], ie = path_ptr[e] + n_inc[e]; i < ie; i++) p = __dadd_rn(p, path_psi[i]);
    double v = 0.0;
    if (n > 0.0 && p >= 0.0 && p <= 1.0) {
        const double a = __dadd_rn(__dmul_rn(p, n), 1.0), b = __dadd_rn(__dmul_rn(__dsub_rn(1.0, p), n), 1.0);
        const double q = (t & 1) ? 1.0 - (1.0 - 0.9) / 2 : (1.0 - 0.9) / 2;
        v = wq_round_sig(d.tab, wq_beta_quantile(d, a, b, q, (t & 1) ? 1.6448536269514722 : -1.6448536269514722), 3);
    }
    if (t & 1) d.hi[e] = v; else d.lo[e] = v;
}

static inline cudaError_t wq_ci_tables(WqCiDev& d, WqEmScratch& M, cudaStream_t stream);
// The totals go up BEFORE the PSI launch (a copy from pageable memory enqueued behind the kernels would block the host
// until they finish); the kernel itself goes behind wq_run_em_psi(b, M, stream, ..., wait = false) on the same scratch and
// stream: out = lo[E] then hi[E] (pinned).
static inline std::string wq_psi_ci_upload(uint32_t E, const double* total, WqEmScratch& M, cudaStream_t stream) {
    if (E == 0) return "";
    double* d_tot;
    WQ_EMCK(M.get(19, E, &d_tot));
    WQ_EMCK(cudaMemcpyAsync(d_tot, total, E * 8ull, cudaMemcpyHostToDevice, stream));
    return "";
}
static inline std::string wq_run_psi_ci(uint32_t E, double* out, WqEmScratch& M, cudaStream_t stream, uint64_t* launches) {
    if (E == 0) return "";
    WqCiDev d;
    WQ_EMCK(wq_ci_tables(d, M, stream));
    d.n = E;
    double *d_tot, *d_lo, *d_hi;
    WQ_EMCK(M.get(19, E, &d_tot)); WQ_EMCK(M.get(20, E, &d_lo)); WQ_EMCK(M.get(21, E, &d_hi));
    d.psi = nullptr; d.total = d_tot; d.lo = d_lo; d.hi = d_hi;
    wq_k_psi_ci<<<(2 * E + 127) / 128, 128, 0, stream>>>(d, (const uint32_t*)M.p[0], (const uint32_t*)M.p[1], (const double*)M.p[9]);
    if (launches) *launches += 1;
    WQ_EMCK(cudaGetLastError());
    WQ_EMCK(cudaMemcpyAsync(out, d_lo, E * 8ull, cudaMemcpyDeviceToHost, stream));
    WQ_EMCK(cudaMemcpyAsync(out + E, d_hi, E * 8ull, cudaMemcpyDeviceToHost, stream));
    return "";
}

static inline std::string wq_run_beta_ci(uint32_t n, const double* psi, const double* total, double* lo, double* hi,
                                         WqEmScratch& M, cudaStream_t stream, uint64_t* launches, bool wait = true) {
    if (n == 0) return "";
    WqCiDev d;
    WQ_EMCK(wq_ci_tables(d, M, stream));
    d.n = n;
    double *d_psi, *d_tot, *d_lo, *d_hi;
    WQ_EMCK(M.get(15, n, &d_psi)); WQ_EMCK(M.get(16, n, &d_tot)); WQ_EMCK(M.get(17, n, &d_lo)); WQ_EMCK(M.get(18, n, &d_hi));
    WQ_EMCK(cudaMemcpyAsync(d_psi, psi, n * 8ull, cudaMemcpyHostToDevice, stream));
    WQ_EMCK(cudaMemcpyAsync(d_tot, total, n * 8ull, cudaMemcpyHostToDevice, stream));
    d.psi = d_psi; d.total = d_tot; d.lo = d_lo; d.hi = d_hi;
    wq_k_beta_ci<<<(2 * n + 127) / 128, 128, 0, stream>>>(d);
    if (launches) *launches += 1;
    WQ_EMCK(cudaGetLastError());
    WQ_EMCK(cudaMemcpyAsync(lo, d_lo, n * 8ull, cudaMemcpyDeviceToHost, stream));
    WQ_EMCK(cudaMemcpyAsync(hi, d_hi, n * 8ull, cudaMemcpyDeviceToHost, stream));
    if (wait) WQ_EMCK(cudaStreamSynchronize(stream));
    return "";
}

static inline cudaError_t wq_ci_tables(WqCiDev& d, WqEmScratch& M, cudaStream_t stream) {
    for (int k = 0; k <= 22; k++) { d.tab.p10[k] = std::pow(10.0, (double)k); d.tab.ip10[k] = 1.0 / d.tab.p10[k]; }
    cudaError_t e0 = wq_p10_table(M, stream, &d.tab.p10big);
    if (e0 != cudaSuccess) return e0;
    {   // Gauss-Legendre nodes on [-1, 1] by Newton iteration on P_n, mapped to [0, 1]
        const int n = WQ_CI_GL;
        for (int i = 0; i < (n + 1) / 2; i++) {
            double zz = std::cos(3.14159265358979323846 * (i + 0.75) / (n + 0.5)), pp = 1.0;
            for (int it = 0; it < 100; it++) {
                double p1 = 1.0, p2 = 0.0;
                for (int j = 0; j < n; j++) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * zz * p2 - j * p3) / (j + 1.0); }
                pp = n * (zz * p1 - p2) / (zz * zz - 1.0);
                const double z1 = zz;
                zz = z1 - p1 / pp;
                if (std::fabs(zz - z1) < 1e-16) break;
            }
            const double w = 2.0 / ((1.0 - zz * zz) * pp * pp);
            d.gl_y[i] = 0.5 * (1.0 - zz); d.gl_y[n - 1 - i] = 0.5 * (1.0 + zz);
            d.gl_w[i] = 0.5 * w; d.gl_w[n - 1 - i] = 0.5 * w;
        }
    }
    return cudaSuccess;
}
