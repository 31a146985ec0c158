// Host side of the event stage: turns the node/edge stores of every gene into PSI-EM problems
// (process_events / _process_events, /root/reference/src/events.jl:744-800 with isnodeok=false as the CLI sets it,
// bin/whippet-quant.jl:181), which wq_k_em_psi then solves in one batched launch.  Node sets are bit windows
// over the event's node range.  The writers of src/io.jl are out of scope: every value they would print is
// returned instead.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>
#include "wq_host_quant.h"
#include "wq_host_index.h"

enum : uint8_t { WQM_TXST = 0, WQM_TXEN = 1, WQM_ALTF = 2, WQM_ALTL = 3, WQM_RETI = 4, WQM_SKIP = 5, WQM_ALTD = 6, WQM_ALTA = 7, WQM_NONE = 8 };

// (edgetype[i], edgetype[i+1]) -> motif, src/events.jl:41-106
static inline uint8_t wq_motif(uint8_t cur, uint8_t next) {
    // rows: current edge type SL SR LS RS LR LL RR; columns: next edge type
    static const uint8_t T[7][7] = {
        /* SL */ {WQM_TXST, WQM_NONE, WQM_NONE, WQM_NONE, WQM_NONE, WQM_NONE, WQM_NONE},
        /* SR */ {WQM_NONE, WQM_ALTL, WQM_SKIP, WQM_NONE, WQM_SKIP, WQM_SKIP, WQM_ALTA},
        /* LS */ {WQM_TXST, WQM_NONE, WQM_ALTF, WQM_NONE, WQM_ALTF, WQM_ALTF, WQM_NONE},
        /* RS */ {WQM_NONE, WQM_TXEN, WQM_NONE, WQM_TXEN, WQM_NONE, WQM_NONE, WQM_NONE},
        /* LR */ {WQM_NONE, WQM_ALTL, WQM_SKIP, WQM_NONE, WQM_SKIP, WQM_SKIP, WQM_ALTA},
        /* LL */ {WQM_NONE, WQM_NONE, WQM_ALTD, WQM_NONE, WQM_ALTD, WQM_ALTD, WQM_RETI},
        /* RR */ {WQM_NONE, WQM_ALTL, WQM_SKIP, WQM_NONE, WQM_SKIP, WQM_SKIP, WQM_ALTA}};
    if (cur > 6 || next > 6) return WQM_NONE;
    return T[cur][next];
}

#define WQ_WIN_INLINE 6                 // 384 nodes: the widest gene of the RefSeq annotation (TTN, 365 nodes) needs no heap per set
struct WqBitWin {                       // node set as bits over [lo, lo + 64*nw); windows up to 64 * WQ_WIN_INLINE nodes need no heap
    uint32_t lo = 0, nw = 0;
    uint64_t in[WQ_WIN_INLINE] = {0};
    std::vector<uint64_t> big;
    uint64_t* p() { return nw <= WQ_WIN_INLINE ? in : big.data(); }
    const uint64_t* p() const { return nw <= WQ_WIN_INLINE ? in : big.data(); }
    void init(uint32_t lo_, uint32_t hi_) {
        lo = lo_; nw = (hi_ - lo_) / 64 + 1;
        for (auto& w : in) w = 0;
        if (nw > WQ_WIN_INLINE) big.assign(nw, 0); else big.clear();
    }
    void set(uint32_t n) { p()[(n - lo) >> 6] |= 1ull << ((n - lo) & 63); }
    bool has(uint32_t n) const { uint32_t k = n - lo; return n >= lo && (k >> 6) < nw && ((p()[k >> 6] >> (k & 63)) & 1); }
    uint32_t first() const { const uint64_t* w = p(); for (uint32_t i = 0; i < nw; i++) if (w[i]) return lo + 64 * i + (uint32_t)__builtin_ctzll(w[i]); return 0; }
    uint32_t last() const { const uint64_t* w = p(); for (uint32_t i = nw; i-- > 0;) if (w[i]) return lo + 64 * i + 63 - (uint32_t)__builtin_clzll(w[i]); return 0; }
    bool operator==(const WqBitWin& o) const {
        if (nw != o.nw) return false;
        const uint64_t *a = p(), *b = o.p();
        for (uint32_t i = 0; i < nw; i++) if (a[i] != b[i]) return false;
        return true;
    }
    void unite(const WqBitWin& o) { uint64_t* a = p(); const uint64_t* b = o.p(); for (uint32_t i = 0; i < nw; i++) a[i] |= b[i]; }
    // both ends present and nothing strictly between (src/quant.jl:68-79)
    bool adjacent(uint32_t a, uint32_t b) const {
        if (!has(a) || !has(b)) return false;
        if (b <= a + 1) return true;
        // any bit strictly between a and b?  (word masks instead of a per-node probe)
        const uint32_t x = a + 1 - lo, y = b - 1 - lo;                  // inclusive bit range, inside the window (a, b are set)
        const uint64_t* w = p();
        const uint32_t wx = x >> 6, wy = y >> 6;
        const uint64_t mx = ~0ull << (x & 63), my = ~0ull >> (63 - (y & 63));
        if (wx == wy) return (w[wx] & mx & my) == 0;
        if (w[wx] & mx) return false;
        for (uint32_t i = wx + 1; i < wy; i++) if (w[i]) return false;
        return (w[wy] & my) == 0;
    }
    void list(std::vector<uint32_t>& out) const {
        const uint64_t* w = p();
        for (uint32_t i = 0; i < nw; i++) { uint64_t x = w[i]; while (x) { out.push_back(lo + 64 * i + (uint32_t)__builtin_ctzll(x)); x &= x - 1; } }
    }
};

struct WqPathSet {                      // PsiGraph without psi (src/events.jl:134-141)
    std::vector<WqBitWin> nodes;
    std::vector<double> count, length;
    uint32_t mn = 0, mx = 0;
    size_t size() const { return nodes.size(); }
    bool covers(uint32_t n) const { for (auto& s : nodes) if (s.has(n)) return true; return false; }
    long find(const WqBitWin& s) const { for (size_t i = 0; i < nodes.size(); i++) if (nodes[i] == s) return (long)i; return -1; }
};
struct WqGeneEdge { uint32_t a, b; double v; };

struct WqAmbSet { std::vector<uint32_t> paths; double mult; };      // AmbigCounts (paths 0-based here)
// the ambiguous sets of one event, flat (no allocation per set): set a = paths[off[a] .. off[a+1]), multiplier mult[a]
struct WqAmbFlat {
    std::vector<uint32_t> paths, off{0};
    std::vector<double> mult;
    void clear() { paths.clear(); off.resize(1); mult.clear(); }
    size_t size() const { return mult.size(); }
    // ambig_add!: a set seen before takes the value, a new one is appended when the value is positive (src/events.jl:299-326)
    void add(const std::vector<uint32_t>& set, double v) {
        for (size_t a = 0; a < mult.size(); a++)
            if (off[a + 1] - off[a] == set.size() && std::equal(set.begin(), set.end(), paths.begin() + off[a])) { mult[a] += v; return; }
        if (v > 0) { paths.insert(paths.end(), set.begin(), set.end()); off.push_back((uint32_t)paths.size()); mult.push_back(v); }
    }
};

struct WqEventRec {                     // one record per visited node (include/whippet_b200.h::wq_event)
    uint32_t gene, node; uint8_t motif, kind, flags;
    uint32_t n_utr = 0, n_inc = 0, n_exc = 0;
    double psi = 0.0, total = 0.0, ci_lo = 0.0, ci_hi = 0.0;
    long problem = -1;                  // index into the EM batch, -1 when no EM is needed
    int32_t iterations_ = 0;
    uint32_t path_first = 0, n_paths = 0;           // into the owning WqEventPool: utr paths, or inclusion then exclusion paths
    uint32_t value_start = 0, n_values = 0;         // into the finished value array
    int32_t ci_slot = -1;                           // spliced event solved without EM: its slot in the batch's Beta-CI launch
    uint8_t batch = 0;                              // which gene group built it (0: genes no multi-map read touches, 1: the others)
};

struct WqEventPool {                    // node ids of all paths of a batch of records, CSR
    std::vector<uint64_t> path_ptr{0};
    std::vector<uint32_t> path_nodes;
    uint32_t add(const WqBitWin& s) { s.list(path_nodes); path_ptr.push_back(path_nodes.size()); return (uint32_t)path_ptr.size() - 2; }
};

struct WqPsiProblems {                  // flattened wq_psi_batch inputs
    std::vector<uint32_t> path_ptr{0}, n_inc, amb_ptr{0}, amb_path_ptr{0}, amb_paths;
    std::vector<double> count, length, amb_mult;
    std::vector<uint8_t> is_tandem;
    size_t size() const { return n_inc.size(); }
    long add(const WqPathSet* inc, const WqPathSet* exc, const std::vector<WqAmbSet>& amb, bool tandem) {
        for (const WqPathSet* g : {inc, exc}) if (g) { count.insert(count.end(), g->count.begin(), g->count.end()); length.insert(length.end(), g->length.begin(), g->length.end()); }
        path_ptr.push_back((uint32_t)count.size());
        n_inc.push_back((uint32_t)(inc ? inc->size() : 0));
        for (auto& a : amb) { amb_paths.insert(amb_paths.end(), a.paths.begin(), a.paths.end()); amb_path_ptr.push_back((uint32_t)amb_paths.size()); amb_mult.push_back(a.mult); }
        amb_ptr.push_back((uint32_t)amb_mult.size());
        is_tandem.push_back(tandem ? 1 : 0);
        return (long)size() - 1;
    }
    // parallel concatenation: sizes() of every part -> prefix sums -> resize_total() once -> place() per part (any thread)
    struct Sizes { size_t ev = 0, paths = 0, amb = 0, amb_paths = 0; };
    Sizes sizes() const { Sizes z; z.ev = size(); z.paths = count.size(); z.amb = amb_mult.size(); z.amb_paths = amb_paths.size(); return z; }
    void resize_total(const Sizes& z) {
        path_ptr.resize(z.ev + 1); n_inc.resize(z.ev); amb_ptr.resize(z.ev + 1); amb_path_ptr.resize(z.amb + 1); amb_paths.resize(z.amb_paths);
        count.resize(z.paths); length.resize(z.paths); amb_mult.resize(z.amb); is_tandem.resize(z.ev);
        path_ptr[0] = 0; amb_ptr[0] = 0; amb_path_ptr[0] = 0;
    }
    void place(const WqPsiProblems& o, const Sizes& at) {
        const uint32_t pb = (uint32_t)at.paths, ab = (uint32_t)at.amb, apb = (uint32_t)at.amb_paths;
        std::copy(o.count.begin(), o.count.end(), count.begin() + at.paths);
        std::copy(o.length.begin(), o.length.end(), length.begin() + at.paths);
        for (size_t i = 1; i < o.path_ptr.size(); i++) path_ptr[at.ev + i] = pb + o.path_ptr[i];
        std::copy(o.n_inc.begin(), o.n_inc.end(), n_inc.begin() + at.ev);
        for (size_t i = 1; i < o.amb_ptr.size(); i++) amb_ptr[at.ev + i] = ab + o.amb_ptr[i];
        for (size_t i = 1; i < o.amb_path_ptr.size(); i++) amb_path_ptr[at.amb + i] = apb + o.amb_path_ptr[i];
        std::copy(o.amb_paths.begin(), o.amb_paths.end(), amb_paths.begin() + at.amb_paths);
        std::copy(o.amb_mult.begin(), o.amb_mult.end(), amb_mult.begin() + at.amb);
        std::copy(o.is_tandem.begin(), o.is_tandem.end(), is_tandem.begin() + at.ev);
    }
    long add(const WqPathSet* inc, const WqPathSet* exc, const WqAmbFlat& amb, bool tandem) {
        for (const WqPathSet* g : {inc, exc}) if (g) { count.insert(count.end(), g->count.begin(), g->count.end()); length.insert(length.end(), g->length.begin(), g->length.end()); }
        path_ptr.push_back((uint32_t)count.size());
        n_inc.push_back((uint32_t)(inc ? inc->size() : 0));
        const uint32_t base = (uint32_t)amb_paths.size();
        amb_paths.insert(amb_paths.end(), amb.paths.begin(), amb.paths.end());
        for (size_t a = 0; a < amb.size(); a++) { amb_path_ptr.push_back(base + amb.off[a + 1]); amb_mult.push_back(amb.mult[a]); }
        amb_ptr.push_back((uint32_t)amb_mult.size());
        is_tandem.push_back(tandem ? 1 : 0);
        return (long)size() - 1;
    }
    void append(const WqPsiProblems& o, long& shift) {
        shift = (long)size();
        const uint32_t pb = (uint32_t)count.size(), ab = (uint32_t)amb_mult.size(), apb = (uint32_t)amb_paths.size();
        count.insert(count.end(), o.count.begin(), o.count.end()); length.insert(length.end(), o.length.begin(), o.length.end());
        for (size_t i = 1; i < o.path_ptr.size(); i++) path_ptr.push_back(pb + o.path_ptr[i]);
        n_inc.insert(n_inc.end(), o.n_inc.begin(), o.n_inc.end());
        for (size_t i = 1; i < o.amb_ptr.size(); i++) amb_ptr.push_back(ab + o.amb_ptr[i]);
        for (size_t i = 1; i < o.amb_path_ptr.size(); i++) amb_path_ptr.push_back(apb + o.amb_path_ptr[i]);
        amb_paths.insert(amb_paths.end(), o.amb_paths.begin(), o.amb_paths.end());
        amb_mult.insert(amb_mult.end(), o.amb_mult.begin(), o.amb_mult.end());
        is_tandem.insert(is_tandem.end(), o.is_tandem.begin(), o.is_tandem.end());
    }
};

static inline void wq_amb_add(std::vector<WqAmbSet>& amb, const std::vector<uint32_t>& set, double v) {
    for (auto& a : amb) if (a.paths == set) { a.mult += v; return; }
    if (v > 0) amb.push_back(WqAmbSet{set, v});
}

#ifdef WQ_EV_PROF
#include <x86intrin.h>
static unsigned long long wq_evp[8];
static unsigned long long wq_evc[8];     // [0] reductions, [1] rounds, [2] unions, [3] largest path set before truncation, [4] paths visited
#define WQ_EVC(k, v) (wq_evc[k] += (v))
#define WQ_EVC_MAX(k, v) do { if ((unsigned long long)(v) > wq_evc[k]) wq_evc[k] = (v); } while (0)
#define WQ_EVP(k) do { unsigned long long t_ = __rdtsc(); wq_evp[k] += t_ - evp_t; evp_t = t_; } while (0)
#define WQ_EVP_START unsigned long long evp_t = __rdtsc()
#else
#define WQ_EVC(k, v) do { } while (0)
#define WQ_EVC_MAX(k, v) do { } while (0)
#define WQ_EVP(k) do { } while (0)
#define WQ_EVP_START do { } while (0)
#endif

// Build minimal node-set paths from edges sharing terminal nodes (reduce_graph, src/events.jl:223-266).
//
// The rounds run on FLAT storage: a node set is `nw` consecutive 64-bit words of one array (nw = words of the event's node window,
// 1 for nearly every gene), so a union is nw ORs into the tail of the array, no set owns heap memory, and membership (`jointset in
// newgraph.nodes`, a linear scan in the reference, src/events.jl:246,254, i.e. quadratic in the number of paths) is an open-addressing
// table of set indices keyed by a hash of the words.  Only membership is asked, so the table changes no result: paths keep their
// order of first appearance.  An edge joins a path when its last node is the path's first or its first node is the path's last
// (hasintersect_terminal); the reference scans every edge for every path (:242-243), here the edges are grouped by first and by last
// node once (the edge set never changes during the reduction) and a path visits its two groups merged in ascending edge order, which
// is the order the scan would have met them in.
static inline WqPathSet wq_reduce(const WqPathSet& edges, bool& subsampled) {
#ifdef WQ_EV_PROF
    const unsigned long long t_enter = __rdtsc();
#endif
    WqPathSet nxt;
    nxt.mn = edges.mn; nxt.mx = edges.mx;
    const size_t E = edges.size();
    if (E == 0) return nxt;
    struct Scratch {
        std::vector<uint32_t> ef, el, by_first, by_last, off_first, off_last, cf, cl;
        std::vector<uint64_t> ew, cs, ns;            // edge sets, current paths, next paths: nw words each
        std::vector<double> clen, ccnt, nlen, ncnt;
        std::vector<uint32_t> tab;                   // 0 = empty, else 1 + index of a set in ns
    };
    static thread_local Scratch* scratch = nullptr;  // one TLS lookup per call (every access to a thread_local of a shared library is a call)
    if (!scratch) scratch = new Scratch();
    Scratch& S = *scratch;
    const uint32_t nw = edges.nodes[0].nw, wlo = edges.nodes[0].lo;
    for (size_t j = 0; j < E; j++)
        if (edges.nodes[j].nw != nw || edges.nodes[j].lo != wlo) { nxt = edges; return nxt; }       // cannot happen: one window per event
    S.ef.resize(E); S.el.resize(E); S.ew.resize(E * nw);
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    for (size_t j = 0; j < E; j++) {
        const uint64_t* w = edges.nodes[j].p();
        for (uint32_t k = 0; k < nw; k++) S.ew[j * nw + k] = w[k];
        S.ef[j] = edges.nodes[j].first(); S.el[j] = edges.nodes[j].last();
        lo = std::min(lo, S.ef[j]); hi = std::max(hi, S.el[j]);
    }
    const uint32_t span = hi - lo + 1;
    S.off_first.assign(span + 1, 0); S.off_last.assign(span + 1, 0);
    for (size_t j = 0; j < E; j++) { S.off_first[S.ef[j] - lo + 1]++; S.off_last[S.el[j] - lo + 1]++; }
    for (uint32_t k = 0; k < span; k++) { S.off_first[k + 1] += S.off_first[k]; S.off_last[k + 1] += S.off_last[k]; }
    S.by_first.resize(E); S.by_last.resize(E);
    S.cf.assign(S.off_first.begin(), S.off_first.end() - 1); S.cl.assign(S.off_last.begin(), S.off_last.end() - 1);
    for (size_t j = 0; j < E; j++) { S.by_first[S.cf[S.ef[j] - lo]++] = (uint32_t)j; S.by_last[S.cl[S.el[j] - lo]++] = (uint32_t)j; }     // ascending j inside a group
    WQ_EVC(0, 1);
#ifdef WQ_EV_PROF
    const unsigned long long t_pre = __rdtsc();
    wq_evc[6] += t_pre - t_enter;
#endif
    S.cs.assign(S.ew.begin(), S.ew.end());
    S.clen.assign(edges.length.begin(), edges.length.end()); S.ccnt.assign(edges.count.begin(), edges.count.end());
    S.ns.clear(); S.nlen.clear(); S.ncnt.clear();
    uint32_t nmn = edges.mn, nmx = edges.mx;
    size_t used = 0;
    auto hash = [nw](const uint64_t* w) -> uint64_t {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (uint32_t k = 0; k < nw; k++) { h ^= w[k]; h *= 0xFF51AFD7ED558CCDull; h ^= h >> 31; }
        return h;
    };
    auto place = [&](uint32_t idx1) { const size_t mask = S.tab.size() - 1; size_t h = (size_t)hash(&S.ns[(size_t)(idx1 - 1) * nw]) & mask; while (S.tab[h]) h = (h + 1) & mask; S.tab[h] = idx1; };
    // the candidate set sits at the tail of ns (index `at`): true when an equal set is already there, else it stays and is indexed
    auto seen_or_keep = [&](size_t at) -> bool {
        if (2 * (used + 1) > S.tab.size()) {            // grow, re-inserting what is there
            S.tab.assign(S.tab.size() * 2, 0);
            for (size_t i = 0; i < at; i++) place((uint32_t)i + 1);
        }
        const uint64_t* w = &S.ns[at * nw];
        const size_t mask = S.tab.size() - 1;
        size_t h = (size_t)hash(w) & mask;
        while (S.tab[h]) {
            const uint64_t* o = &S.ns[(size_t)(S.tab[h] - 1) * nw];
            bool eq = true;
            for (uint32_t k = 0; k < nw && eq; k++) eq = o[k] == w[k];
            if (eq) return true;
            h = (h + 1) & mask;
        }
        S.tab[h] = (uint32_t)at + 1;
        used++;
        return false;
    };
    auto first_of = [&](const uint64_t* w) -> uint32_t { for (uint32_t k = 0; k < nw; k++) if (w[k]) return wlo + 64 * k + (uint32_t)__builtin_ctzll(w[k]); return 0; };
    auto last_of = [&](const uint64_t* w) -> uint32_t { for (uint32_t k = nw; k-- > 0;) if (w[k]) return wlo + 64 * k + 63 - (uint32_t)__builtin_clzll(w[k]); return 0; };
    for (int it = 1; it <= 100; it++) {
        WQ_EVC(1, 1);
        if (it > 1) {
            if (S.ns == S.cs) break;
            S.cs.swap(S.ns); S.clen.swap(S.nlen); S.ccnt.swap(S.ncnt);
            S.ns.clear(); S.nlen.clear(); S.ncnt.clear();
            WQ_EVC_MAX(3, S.clen.size());
            if (S.clen.size() > 1024) { subsampled = true; S.cs.resize((size_t)1024 * nw); S.clen.resize(1024); S.ccnt.resize(1024); }
        }
        const size_t ncur = S.clen.size();
        {
            size_t cap = 64;
            while (cap < 8 * ncur + 32) cap <<= 1;
            S.tab.assign(cap, 0);
            used = 0;
        }
        for (size_t i = 0; i < ncur; i++) {
            const uint64_t* w = &S.cs[i * nw];
            const uint32_t fi = first_of(w), li = last_of(w);
            // edges with last == fi, merged with edges with first == li, ascending, without duplicates
            const uint32_t *a = nullptr, *ae = nullptr, *b = nullptr, *be = nullptr;
            if (fi >= lo && fi <= hi) { a = S.by_last.data() + S.off_last[fi - lo]; ae = S.by_last.data() + S.off_last[fi - lo + 1]; }
            if (li >= lo && li <= hi) { b = S.by_first.data() + S.off_first[li - lo]; be = S.by_first.data() + S.off_first[li - lo + 1]; }
            const bool none = a == ae && b == be;
            WQ_EVC(4, 1);
            while (a != ae || b != be) {
                uint32_t j;
                if (b == be || (a != ae && *a <= *b)) { j = *a++; if (b != be && *b == j) b++; }
                else j = *b++;
                WQ_EVC(2, 1);
                const size_t at = S.nlen.size();
                S.ns.resize((at + 1) * nw);
                w = &S.cs[i * nw];
                for (uint32_t k = 0; k < nw; k++) S.ns[at * nw + k] = w[k] | S.ew[(size_t)j * nw + k];
                if (seen_or_keep(at)) { S.ns.resize(at * nw); continue; }
                S.nlen.push_back(S.clen[i] + edges.length[j]);
                S.ncnt.push_back(S.ccnt[i] + edges.count[j]);
                nmn = std::min(std::min(fi, S.ef[j]), nmn);
                nmx = std::max(std::max(li, S.el[j]), nmx);
            }
            if (none) {
                const size_t at = S.nlen.size();
                S.ns.resize((at + 1) * nw);
                w = &S.cs[i * nw];
                for (uint32_t k = 0; k < nw; k++) S.ns[at * nw + k] = w[k];
                if (seen_or_keep(at)) { S.ns.resize(at * nw); continue; }
                S.nlen.push_back(S.clen[i]);
                S.ncnt.push_back(S.ccnt[i]);
                nmn = std::min(fi, nmn);
                nmx = std::max(li, nmx);
            }
        }
    }
#ifdef WQ_EV_PROF
    wq_evc[7] += __rdtsc() - t_pre;
#endif
    nxt.mn = nmn; nxt.mx = nmx;
    const size_t nout = S.nlen.size();
    nxt.nodes.resize(nout);
    for (size_t i = 0; i < nout; i++) {
        WqBitWin& s = nxt.nodes[i];
        s.lo = wlo; s.nw = nw;
        if (nw > WQ_WIN_INLINE) s.big.assign(S.ns.begin() + i * nw, S.ns.begin() + (i + 1) * nw);
        else for (uint32_t k = 0; k < nw; k++) s.in[k] = S.ns[i * nw + k];
    }
    nxt.count.assign(S.ncnt.begin(), S.ncnt.end());
    nxt.length.assign(S.nlen.begin(), S.nlen.end());
    return nxt;
}
struct WqGeneEvents {
    const WqHostIndex& H; const std::vector<double>& node_value; uint32_t g;      // g 1-based
    std::vector<WqGeneEdge> edges;                                                // (first,last) order
    uint32_t nn() const { return H.genes[g - 1].nn; }
    uint8_t et(uint32_t j) const { const WqNodeRec* r = &H.nodes[H.genes[g - 1].node_base]; return j <= nn() ? r[j - 1].et_left : r[nn() - 1].et_right; }
    double nodev(uint32_t n) const { return node_value[H.genes[g - 1].node_base + n - 1]; }
    double nodelen(uint32_t n) const { return (double)H.nodes[H.genes[g - 1].node_base + n - 1].len; }
    double edgev(uint32_t a, uint32_t b) const { for (auto& e : edges) if (e.a == a && e.b == b) return e.v; return 0.0; }
};

static inline void wq_push_edge(WqPathSet& g, const WqGeneEdge& e, uint32_t lo, uint32_t hi) {
    WqBitWin s;
    s.init(lo, hi);
    s.set(e.a); s.set(e.b);
    g.nodes.push_back(s); g.count.push_back(0.0); g.length.push_back(0.0);
    g.mn = std::min(g.mn, e.a); g.mx = std::max(g.mx, e.b);
}

// The nodes the loop of _process_events visits (src/events.jl:760-800): every node, except that a tandem-UTR event consumes the
// run of nodes it prints (src/io.jl:279-316).  The runs depend on the edge types only, so the visit list is known without any
// counts; a gene with hundreds of nodes is cut at visited nodes into pieces that different host threads build.
static inline void wq_gene_visits(const WqGeneEvents& V, std::vector<uint32_t>& visits) {
    const uint32_t ne = V.nn() + 1;
    visits.clear();
    if (ne <= 2) return;
    for (uint32_t i = 1; i < ne; i++) {
        visits.push_back(i);
        const uint8_t motif = wq_motif(V.et(i), V.et(i + 1));
        if (motif > 1) continue;
        uint32_t lo = i, hi = i;
        if (motif == WQM_TXEN) lo = i - 1;
        uint32_t k = i;
        for (;;) { hi = k; k++; if (!(k < ne && wq_motif(V.et(k), V.et(k + 1)) == motif)) break; }
        if (motif == WQM_TXST && k <= V.nn()) hi = k;
        const uint32_t n_utr = hi - lo + 1;                  // the used nodes are the contiguous run lo..hi
        const uint32_t st = motif == WQM_TXST ? i : i - 1;
        i = st + n_utr - 1;
    }
}

// Estimated cost of the events of nodes node_lo..node_hi of one gene, from its expressed edges (el[k], er[k]): the paths of node n's
// event are drawn from the edges inside the hull of the edges that span or touch n and carry at least 2 % of the gene's reference
// edge value (src/events.jl:647-742 collects exactly those before reduce_graph enumerates; a node without both an inclusion and an
// exclusion edge has no enumeration at all), and the work grows roughly with the cube of their number e_n up to the 1024-path cap
// (measured on the counts of an 80 M-read job: paths x ambiguous sets = 18 at e ~ 12, 1.1 k at e ~ 48, 135 k at e ~ 170 and beyond).
// Used only to place the cuts of the event stage between ranks; any estimate gives the same records.
static inline double wq_piece_cost(const std::vector<uint32_t>& el, const std::vector<uint32_t>& er, const std::vector<double>& ev, uint32_t node_lo, uint32_t node_hi) {
    double c = 0.0;
    const size_t E = el.size();
    const double maxvalue = E ? ev[E - 1] : 0.0;              // the same reading of maximum(sgquant.edge) as the event build
    for (uint32_t n = node_lo; n <= node_hi; n++) {
        uint32_t lo = 0xFFFFFFFFu, hi = 0;
        bool has_inc = false, has_exc = false;
        for (size_t k = 0; k < E; k++) {
            if (!(el[k] <= n && n <= er[k]) || ev[k] / maxvalue < 0.02) continue;
            lo = std::min(lo, el[k]); hi = std::max(hi, er[k]);
            if (el[k] == n || er[k] == n) has_inc = true; else has_exc = true;
        }
        double e = 0.0;
        if (has_inc && has_exc) for (size_t k = 0; k < E; k++) if (el[k] >= lo && er[k] <= hi) e += 1.0;
        e = std::min(e, 128.0);                               // beyond ~128 edges the enumeration stops at the 1024-path cap
        c += 1.0 + e * e * e;
        if (n == 0xFFFFFFFFu) break;
    }
    return c;
}

// Work list of the event stage.  The gene range is cut into many more contiguous pieces than there are host threads and the threads
// pull pieces as they finish: the cost of a gene goes from nothing to hundreds of path enumerations (TTN: 365 nodes), so a gene with
// many nodes is itself cut into pieces at visited nodes (every node's event is independent of the others).  The pieces depend on the
// index, the thread count and the partition only, so both gene groups and every call see the same list.
struct WqEvPiece { uint32_t g_lo, g_hi, node_lo, node_hi; };      // a contiguous gene range, or a range of visited nodes of one gene with many nodes
static inline void wq_event_make_pieces(const WqHostIndex& H, const std::vector<double>& node_value, unsigned nt, uint32_t nparts, std::vector<WqEvPiece>& mk) {
    const uint32_t G = H.G;
    mk.clear();
    const uint32_t per_piece = nt == 1 && nparts == 1 ? G : std::max<uint32_t>(8, G / (16 * std::max(nt, 4u)) + 1), big_gene = 32;
    std::vector<uint32_t> visits;
    WqGeneEvents Vv{H, node_value, 0, {}};
    uint32_t start = 1;
    auto flush = [&](uint32_t upto) { if (upto >= start) mk.push_back(WqEvPiece{start, upto, 1, 0xFFFFFFFFu}); start = upto + 1; };
    for (uint32_t g = 1; g <= G; g++) {
        if ((nt > 1 || nparts > 1) && H.genes[g - 1].nn >= big_gene) {
            flush(g - 1);
            Vv.g = g;
            wq_gene_visits(Vv, visits);
            const size_t visits_per_piece = H.genes[g - 1].nn >= 128 ? 2 : 8;      // the widest genes have the costliest nodes
            for (size_t v = 0; v < visits.size(); v += visits_per_piece) {
                const size_t ve = std::min(visits.size(), v + visits_per_piece);
                mk.push_back(WqEvPiece{g, g, visits[v], visits[ve - 1]});
            }
            start = g + 1;
        } else if (g - start + 1 >= per_piece) flush(g);
    }
    flush(G);
    if (mk.empty()) mk.push_back(WqEvPiece{1, 0, 1, 0});
}
// The ranks' ranges of pieces (wq_set_gene_partition): rank r takes pieces [cuts[r], cuts[r + 1]), cut where the estimated cost reaches
// r / nparts of the total.  Whole-gene pieces cost nodes + E^3 (E = the gene's expressed edges); a gene cut into node pieces costs
// wq_piece_cost of each piece, so the ~170 events of one titin-like gene that sit under one long exon-skipping junction (each
// ~1000-1900 paths: nearly all of the stage once a human job has tens of millions of reads) are spread over nearly all ranks instead
// of landing on the one or two that hold the middle of the gene.  edge_key must be sorted (gene, first, last).
static inline void wq_event_cuts(const WqHostIndex& H, const std::vector<uint64_t>& edge_key, const std::vector<double>& edge_value, const std::vector<WqEvPiece>& pieces, uint32_t nparts,
                                 std::vector<unsigned>& cuts, std::vector<double>* cum_out = nullptr) {
    const uint32_t G = H.G;
    const unsigned npieces = (unsigned)pieces.size();
    std::vector<double> gcost(G + 2, 0.0);
    for (uint64_t k : edge_key) { const uint64_t g = k >> 32; if (g >= 1 && g <= G && !wq_edge_is_circ(k)) gcost[g] += 1.0; }
    for (uint32_t g = 1; g <= G; g++) gcost[g] = (double)H.genes[g - 1].nn + gcost[g] * gcost[g] * gcost[g];
    std::vector<double> cum(npieces + 1, 0.0);
    std::vector<uint32_t> el, er;
    std::vector<double> ev;
    uint32_t el_gene = 0;
    for (unsigned t = 0; t < npieces; t++) {
        const WqEvPiece& pc = pieces[t];
        double c = 0.0;
        if (pc.g_lo == pc.g_hi && pc.node_hi != 0xFFFFFFFFu) {
            const uint32_t g = pc.g_lo;
            if (el_gene != g) {
                el.clear(); er.clear(); ev.clear(); el_gene = g;
                for (size_t e = std::lower_bound(edge_key.begin(), edge_key.end(), (uint64_t)g << 32) - edge_key.begin(); e < edge_key.size() && (edge_key[e] >> 32) == g; e++) {
                    if (wq_edge_is_circ(edge_key[e])) continue;
                    el.push_back((uint32_t)((edge_key[e] >> 16) & 0xFFFF)); er.push_back((uint32_t)(edge_key[e] & 0xFFFF)); ev.push_back(edge_value[e]);
                }
            }
            c = wq_piece_cost(el, er, ev, pc.node_lo, pc.node_hi);
        } else for (uint32_t g = pc.g_lo; g <= pc.g_hi && g <= G; g++) c += gcost[g];
        cum[t + 1] = cum[t] + c;
    }
    cuts.assign(nparts + 1, npieces);
    cuts[0] = 0;
    for (uint32_t r = 1; r < nparts; r++) {
        const double target = cum[npieces] * (double)r / (double)nparts;
        cuts[r] = std::max(cuts[r - 1], std::min(npieces, (unsigned)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin())));
    }
    if (cum_out) *cum_out = cum;
}

// all events of one gene (src/events.jl:760-800), or of its visited nodes in [node_lo, node_hi] (node_lo must be a visited node);
// problems needing EM are appended to P
static inline void wq_gene_events(const WqGeneEvents& V, int64_t readlen, std::vector<WqEventRec>& out, WqPsiProblems& P, WqEventPool& pool,
                                  uint32_t node_lo = 1, uint32_t node_hi = 0xFFFFFFFFu) {
    const uint32_t ne = V.nn() + 1;
    if (ne <= 2) return;
    // thread-local scratch, looked up once per gene (every access to a thread_local of a shared library is a call)
    struct Scratch { std::vector<const WqGeneEdge*> touch, amb_edges; WqPathSet inc, exc; WqAmbFlat amb; std::vector<uint32_t> set; };
    static thread_local Scratch tl_scratch;
    Scratch& sc = tl_scratch;
    std::vector<const WqGeneEdge*>&touch = sc.touch, &amb_edges = sc.amb_edges;
    WqPathSet &inc = sc.inc, &exc = sc.exc;
    WqAmbFlat& amb = sc.amb;
    std::vector<uint32_t>& set = sc.set;
    for (uint32_t i = node_lo; i < ne && i <= node_hi; i++) {
        WQ_EVP_START;
        const uint8_t motif = wq_motif(V.et(i), V.et(i + 1));
        const bool next_tandem = (i + 1 < ne) && wq_motif(V.et(i + 1), V.et(i + 2)) <= 1;
        WqEventRec R;
        R.gene = V.g; R.node = i; R.motif = motif; R.kind = 0; R.flags = 0;
        if (motif == WQM_NONE) { R.kind = next_tandem ? 3 : 0; out.push_back(R); WQ_EVP(0); continue; }
        if (motif <= 1) {
            // ---- tandem UTR (src/events.jl:593-645) ----
            std::vector<uint32_t> used;
            double total = 0.0;
            if (motif == WQM_TXEN) { used.push_back(i - 1); total += V.nodev(i - 1); }
            uint32_t k = i;
            for (;;) {
                used.push_back(k);
                total += V.nodev(k);
                if (k > 1) total += V.edgev(k - 1, k);
                k++;
                if (!(k < ne && wq_motif(V.et(k), V.et(k + 1)) == motif)) break;
            }
            if (motif == WQM_TXST && k <= V.nn()) { used.push_back(k); total += V.nodev(k); total += V.edgev(k - 1, k); }
            std::sort(used.begin(), used.end());
            used.erase(std::unique(used.begin(), used.end()), used.end());
            R.kind = 2;
            R.n_utr = (uint32_t)used.size();
            if (total > 2) {
                const uint32_t lo = used.front(), hi = used.back();
                WqPathSet g;
                g.mn = lo; g.mx = hi;
                WqBitWin acc;
                acc.init(lo, hi);
                for (size_t t = 0; t < used.size(); t++) {      // nested paths growing from the transcript end inwards
                    acc.set(motif == WQM_TXST ? used[used.size() - 1 - t] : used[t]);
                    g.nodes.push_back(acc); g.count.push_back(0.0); g.length.push_back(0.0);
                }
                std::vector<WqAmbSet> amb;
                std::vector<uint32_t> set;
                for (uint32_t n = lo; n <= hi; n++) {
                    set.clear();
                    for (size_t p = 0; p < g.size(); p++) if (g.nodes[p].has(n)) { set.push_back((uint32_t)p); g.length[p] += V.nodelen(n); }
                    if (set.empty()) continue;
                    if (set.size() == 1) g.count[set[0]] += V.nodev(n); else wq_amb_add(amb, set, V.nodev(n));
                }
                for (auto& e : V.edges) {
                    set.clear();
                    for (size_t p = 0; p < g.size(); p++) if (e.a >= lo && e.b <= hi && g.nodes[p].adjacent(e.a, e.b)) set.push_back((uint32_t)p);
                    if (set.empty()) continue;
                    if (set.size() == 1) g.count[set[0]] += e.v; else wq_amb_add(amb, set, e.v);
                }
                double amb_total = 0.0, cnt_total = 0.0;
                for (auto& a : amb) amb_total += a.mult;
                for (double c : g.count) cnt_total += c;
                R.total = cnt_total + amb_total;
                R.problem = P.add(&g, nullptr, amb, true);
                R.path_first = (uint32_t)pool.path_ptr.size() - 1; R.n_paths = (uint32_t)g.size();
                for (auto& s : g.nodes) pool.add(s);
            } else {
                R.flags |= 1;
            }
            out.push_back(R);
            WQ_EVP(1);
            const uint32_t st = motif == WQM_TXST ? i : i - 1;
            i = st + R.n_utr - 1;         // output_utr returns the last node it printed (src/io.jl:279-316)
            continue;
        }
        // ---- spliced node (src/events.jl:647-742) ----
        uint32_t lo = 0xFFFFFFFFu, hi = 0;
        const double maxvalue = V.edges.empty() ? 0.0 : V.edges.back().v;     // maximum(sgquant.edge): greatest (first,last) entry
        // per-thread scratch, cleared per node (one heap allocation per container and thread instead of per node)
        touch.clear(); amb_edges.clear(); amb.clear();
        inc.nodes.clear(); inc.count.clear(); inc.length.clear(); inc.mx = 0;
        exc.nodes.clear(); exc.count.clear(); exc.length.clear(); exc.mx = 0;
        WQ_EVP(7);
        for (auto& e : V.edges) {
            if (!(e.a <= i && i <= e.b)) continue;
            if (e.v / maxvalue < 0.02) continue;
            touch.push_back(&e);
            lo = std::min(lo, e.a); hi = std::max(hi, e.b);
        }
        inc.mn = exc.mn = 0xFFFFFFFFu;
        bool has_inc = false, has_exc = false;
        double connecting = 0.0, spanning = 0.0;
        for (const WqGeneEdge* e : touch) {
            if (e->a == i || e->b == i) { wq_push_edge(inc, *e, lo, hi); has_inc = true; connecting += e->v; }
            else { wq_push_edge(exc, *e, lo, hi); has_exc = true; spanning += e->v; }
            amb_edges.push_back(e);
        }
        bool sub = false;
        WQ_EVP(2);
        if (has_inc && has_exc) {
            // bridge with neighbouring edges inside [lo, hi] (extend_edges!, src/events.jl:329-394)
            for (uint32_t idx = i - 1; idx > lo && idx >= 1; idx--)
                for (auto& e : V.edges) {
                    if (e.b != idx || e.a < lo) continue;
                    bool pushed = false;
                    if (exc.covers(e.b)) { wq_push_edge(exc, e, lo, hi); pushed = true; }
                    if (inc.covers(e.b)) { wq_push_edge(inc, e, lo, hi); pushed = true; }
                    if (pushed) amb_edges.push_back(&e);
                }
            for (uint32_t idx = i + 1; idx < hi; idx++)
                for (auto& e : V.edges) {
                    if (e.a != idx || e.b > hi) continue;
                    bool pushed = false;
                    if (exc.covers(e.a)) { wq_push_edge(exc, e, lo, hi); pushed = true; }
                    if (inc.covers(e.a)) { wq_push_edge(inc, e, lo, hi); pushed = true; }
                    if (pushed) amb_edges.push_back(&e);
                }
            WQ_EVP(3);
            inc = wq_reduce(inc, sub);
            exc = wq_reduce(exc, sub);
            WQ_EVP(4);
            for (const WqGeneEdge* e : amb_edges) {
                set.clear();
                for (size_t p = 0; p < inc.size(); p++) if (inc.nodes[p].adjacent(e->a, e->b)) { set.push_back((uint32_t)p); inc.length[p] += 1; }
                for (size_t p = 0; p < exc.size(); p++) if (exc.nodes[p].adjacent(e->a, e->b)) { set.push_back((uint32_t)(p + inc.size())); exc.length[p] += 1; }
                if (set.empty()) continue;
                if (set.size() == 1) { if (set[0] < inc.size()) inc.count[set[0]] += e->v; else exc.count[set[0] - inc.size()] += e->v; }
                else amb.add(set, e->v);
            }
        } else {
            // in place: one edge is already its own path, two edges join when they share a terminal node
            // (reduce_graph_simple, src/events.jl:268-282); only three or more go through the general reduction
            for (WqPathSet* g : {&inc, &exc}) {
                if (g->size() <= 1) continue;
                if (g->size() > 2) { *g = wq_reduce(*g, sub); continue; }
                const uint32_t f0 = g->nodes[0].first(), l0 = g->nodes[0].last(), f1 = g->nodes[1].first(), l1 = g->nodes[1].last();
                if (f0 == l1 || f1 == l0) {
                    g->nodes[0].unite(g->nodes[1]);
                    g->length[0] += g->length[1];
                    g->count[0] += g->count[1];
                    g->nodes.pop_back(); g->length.pop_back(); g->count.pop_back();
                }
            }
        }
        WQ_EVP(5);
        bool has_psi = false;
        if (!has_exc) {
            if (has_inc && (connecting >= 2 || (connecting >= 1 && (motif & 6) == 6))) { R.psi = 1.0; has_psi = true; }
        } else if (!has_inc) {
            if (spanning >= 1) { R.psi = 0.0; has_psi = true; }
        } else {
            R.problem = P.add(&inc, &exc, amb, false);
            has_psi = true;
        }
        R.total = connecting + spanning;
        R.n_inc = has_inc ? (uint32_t)inc.size() : 0;
        R.n_exc = has_exc ? (uint32_t)exc.size() : 0;
        R.path_first = (uint32_t)pool.path_ptr.size() - 1; R.n_paths = R.n_inc + R.n_exc;
        for (const WqPathSet* gph : {&inc, &exc}) for (auto& s : gph->nodes) pool.add(s);
        if (sub) R.flags |= 2;
        R.kind = has_psi ? 1 : 0;         // for EM problems the 0 <= psi <= 1 && total > 0 test is applied after the kernel
        out.push_back(R);
        WQ_EVP(6);
    }
}
