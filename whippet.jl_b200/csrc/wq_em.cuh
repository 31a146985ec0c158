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
static inline void wq_spread_path(const WqHostIndex& H, std::vector<double>& node_value, std::vector<std::pair<uint64_t, double>>& edge_adds,
                                  const WqKeyPath& kp, double v) {
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
    auto in_seen = [&](uint32_t x) { for (int i = 0; i < ns; i++) if (seen[i] == x) return true; return false; };
    if (kp.rn == 1) { if (!in_seen(kp.rnd[0])) node_value[nb + kp.rnd[0] - 1] += v; }
    else for (uint32_t i = 0; i + 1 < kp.rn; i++) {
        if (in_seen(kp.rnd[i]) && in_seen(kp.rnd[i + 1])) continue;
        edge_adds.emplace_back(ekey(kp.rnd[i], kp.rnd[i + 1]), v);
    }
}

// Builds multi-map classes (first) and isoform classes, fills the unique isoform counts, and
// re-counts long paths onto their edges (assign_long=true).
static inline std::string wq_host_build_classes(const WqHostIndex& H, const WqIsoIndex& S, WqHostQuant& hq, WqHostQuantEx& X) {
    const uint32_t G = H.G, I = S.I;
    X = WqHostQuantEx();
    X.node_value.assign(hq.node.begin(), hq.node.end());
    X.iso_count.assign(I, 0.0);
    WqClassSet& C = X.cls;
    C.ptr.push_back(0);
    WqGeneBits B;
    uint32_t cached_gene = 0;
    auto bits_of = [&](uint32_t g1) -> const WqGeneBits& {
        if (cached_gene != g1) { B.build(S, g1 - 1, H.genes[g1 - 1].nn); cached_gene = g1; }
        return B;
    };
    // ---- multi-map classes, key order ----
    std::vector<int32_t> ids;
    for (size_t m = 0; m < hq.multis.size(); m++) {
        const uint32_t* key = hq.multis.key(m);
        const uint32_t nh = key[0] & 0xFFFF;
        size_t pos = 1;
        ids.clear();
        bool all = true;
        for (uint32_t h = 0; h < nh && all; h++) {
            WqKeyPath kp;
            pos = wq_key_next(key, pos, kp);
            const WqGeneBits& b = bits_of(kp.gene);
            bool any = false;
            for (uint32_t p = 0; p < b.np; p++)
                if (wq_key_compatible(b, p, kp)) { ids.push_back((int32_t)(S.iso_base[kp.gene - 1] + p)); any = true; }
            if (!any) all = false;
        }
        if (!all) ids.clear();
        std::sort(ids.begin(), ids.end());
        ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        if (ids.size() > 500) return "multi-map class larger than 500 isoforms (prop_temp = zeros(500), src/quant.jl:725)";
        for (int32_t i : ids) { C.iso.push_back(i); C.prop.push_back(1.0 / (double)ids.size()); }
        C.ptr.push_back((uint32_t)C.iso.size());
        C.count.push_back((double)hq.multis.count[m]);
        C.state.push_back(all ? 0 : 2);
        C.postassign.push_back(all ? 0 : 1);
    }
    C.n_multi = (uint32_t)C.count.size();
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
    struct Frag { WqClassSet C; std::vector<std::pair<uint64_t, double>> adds; std::string err; };
    unsigned nthreads = wq_host_threads();
    if (G < 2048) nthreads = 1;
    std::vector<Frag> frags(nthreads);
    auto work = [&](unsigned t) {
        Frag& F = frags[t];
        const uint32_t g_lo = 1 + (uint32_t)((uint64_t)G * t / nthreads), g_hi = (uint32_t)((uint64_t)G * (t + 1) / nthreads);
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
                account(off, (double)hq.node[nb + n - 1]);
            }
            while (e < ne && (hq.edge_key[e] >> 32) < g) e++;
            for (; e < ne && (hq.edge_key[e] >> 32) == g; e++) {
                if (wq_edge_is_circ(hq.edge_key[e])) continue;
                uint32_t l = (hq.edge_key[e] >> 16) & 0xFFFF, r = hq.edge_key[e] & 0xFFFF;
                uint32_t off = (uint32_t)setbuf.size();
                for (uint32_t p = 0; p < b.np; p++) if (b.adjacent(p, l, r)) setbuf.push_back((int32_t)p + 1);
                account(off, (double)hq.edge_count[e]);
            }
            for (uint32_t k = lptr[g]; k < lptr[g + 1]; k++) {
                WqKeyPath kp;
                wq_key_next(hq.longs.key(lidx[k]), 0, kp);
                uint32_t off = (uint32_t)setbuf.size();
                for (uint32_t p = 0; p < b.np; p++) if (wq_key_compatible(b, p, kp)) setbuf.push_back((int32_t)p + 1);
                account(off, (double)hq.longs.count[lidx[k]]);
                wq_spread_path(H, X.node_value, F.adds, kp, (double)hq.longs.count[lidx[k]]);
            }
            // distinct id sets in lexicographic order, values added up (they are integers: order-free)
            order.resize(recs.size());
            for (uint32_t i = 0; i < recs.size(); i++) order[i] = i;
            auto lt = [&](uint32_t x, uint32_t y) {
                return std::lexicographical_compare(setbuf.begin() + recs[x].off, setbuf.begin() + recs[x].off + recs[x].len,
                                                    setbuf.begin() + recs[y].off, setbuf.begin() + recs[y].off + recs[y].len);
            };
            std::sort(order.begin(), order.end(), lt);
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
    for (Frag& F : frags) {
        if (!F.err.empty()) return F.err;
        const uint32_t base = (uint32_t)C.iso.size();
        C.iso.insert(C.iso.end(), F.C.iso.begin(), F.C.iso.end());
        C.prop.insert(C.prop.end(), F.C.prop.begin(), F.C.prop.end());
        for (uint32_t p : F.C.ptr) C.ptr.push_back(base + p);
        C.count.insert(C.count.end(), F.C.count.begin(), F.C.count.end());
        C.state.insert(C.state.end(), F.C.count.size(), 0);
        C.postassign.insert(C.postassign.end(), F.C.count.size(), 0);
        X.edge_adds.insert(X.edge_adds.end(), F.adds.begin(), F.adds.end());
    }
    X.built = true;
    return "";
}

// assign_ambig!, src/quant.jl:520-553 (host: one pass over the multi-map classes)
static inline std::string wq_host_assign_ambig_impl(const WqHostIndex& H, const WqIsoIndex& S, const WqHostQuant& hq, WqHostQuantEx& X) {
    if (!X.built || !X.em_done) return "assign_ambig needs wq_em_tpm first (it uses the EM class proportions and gene TPM)";
    if (X.ambig_done) return "";
    const WqClassSet& C = X.cls;
    WqGeneBits B;
    uint32_t cached = 0;
    WqKeyPath vp[WQ_MAX_TOLERATE];
    double w[WQ_MAX_TOLERATE];
    for (uint32_t c = 0; c < C.n_multi; c++) {
        const uint32_t* key = hq.multis.key(c);
        const uint32_t nh = key[0] & 0xFFFF;
        if (nh > WQ_MAX_TOLERATE) return "multi-map class with more alignments than WQ_MAX_TOLERATE";
        size_t pos = 1;
        for (uint32_t h = 0; h < nh; h++) pos = wq_key_next(key, pos, vp[h]);
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
        for (uint32_t h = 0; h < nh; h++) wq_spread_path(H, X.node_value, X.edge_adds, vp[h], (w[h] / tot) * C.count[c]);
    }
    X.ambig_done = true;
    X.edges_final = false;
    return "";
}

// ------------------------------------------------------------------------------------------ device
struct WqEmDev {
    uint32_t I, C, ntiles;
    const double* leff;          // [I] max(length - readlen, 1)
    double* count;               // [I] unique + finished-class counts (in/out)
    double* tpm;                 // [I] in: calculate_tpm!(quant) of the unique counts; out: final
    double* tpm_raw;             // [I]
    const uint32_t* cls_ptr;     // [C+1]
    const int32_t*  cls_iso;
    double*         prop;
    const double*   cls_count;   // [C]
    uint8_t*        state;       // [C] 0 active, 1 finished in this sweep, 2 finished earlier (or never active)
    const uint32_t* con_ptr;     // [I+1] contributions of each isoform, in class order
    const uint32_t* con_cls;
    const uint32_t* con_pos;
    double* partial;             // [ntiles]
    int*    flags;               // [3] rotating "tpm changed" flags
    int64_t* it_out;
    int64_t maxit;
    double sc_tpm, sc_prop;      // 10^sig, 10^4
    int    sig;
    int    first_differs;        // tpm != ones(I) before the first sweep
};

__device__ __forceinline__ double wq_round_sc(double x, double sc) {   // Base.round(x, digits=d)
    double r = __ddiv_rn(rint(__dmul_rn(x, sc)), sc);
    return isfinite(r) ? r : x;
}

// adjacent-pair tree over a 1024-slot shared tile: buf[0] = sum
__device__ __forceinline__ double wq_tile_tree(double* buf, int t) {
    for (int stride = 1; stride < 1024; stride <<= 1) {
        __syncthreads();
        if ((t & (2 * stride - 1)) == 0) buf[t] = __dadd_rn(buf[t], buf[t + stride]);
    }
    __syncthreads();
    double r = buf[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(1024) wq_k_em_tpm(WqEmDev d) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double buf[1024];
    __shared__ int s_changed;
    const int t = threadIdx.x;
    const uint32_t gtid = blockIdx.x * blockDim.x + t, gsz = gridDim.x * blockDim.x;
    int64_t it = 1;
    bool differs = d.first_differs != 0;
    while (differs && it < d.maxit) {
        // ---- sweep over classes: new proportions from the current TPM (maximize step) ----
        if (gtid == 0) d.flags[(it + 1) % 3] = 0;
        for (uint32_t c = gtid; c < d.C; c += gsz) {
            uint8_t st = d.state[c];
            if (st == 1) { d.state[c] = 2; continue; }
            if (st != 0 || it == 1) continue;
            const uint32_t b = d.cls_ptr[c], e = d.cls_ptr[c + 1];
            double psum = 0.0;
            for (uint32_t k = b; k < e; k++) psum = __dadd_rn(psum, d.tpm[d.cls_iso[k]]);
            bool different = false;
            for (uint32_t k = b; k < e; k++) {          // copy_isdone!, src/quant.jl:702-712
                double val = wq_round_sc(__ddiv_rn(d.tpm[d.cls_iso[k]], psum), d.sc_prop);
                if (val != d.prop[k]) different = true;
                d.prop[k] = isnan(val) ? 0.0 : val;
            }
            if (!different || psum == 0.0) d.state[c] = 1;
        }
        grid.sync();
        // ---- gather per isoform in class order, raw TPM, tile sums ----
        for (uint32_t tile = blockIdx.x; tile < d.ntiles; tile += gridDim.x) {
            const uint32_t i = tile * 1024 + t;
            double raw = 0.0;
            if (i < d.I) {
                double ct = d.count[i], cn = ct;
                for (uint32_t k = d.con_ptr[i]; k < d.con_ptr[i + 1]; k++) {
                    const uint32_t c = d.con_cls[k];
                    const uint8_t st = d.state[c];
                    if (st == 2) continue;
                    const double v = __dmul_rn(d.prop[d.con_pos[k]], d.cls_count[c]);
                    ct = __dadd_rn(ct, v);
                    if (st == 1) cn = __dadd_rn(cn, v);
                }
                d.count[i] = cn;
                raw = __ddiv_rn(ct, d.leff[i]);
                d.tpm_raw[i] = raw;
            }
            buf[t] = raw;
            double s = wq_tile_tree(buf, t);
            if (t == 0) d.partial[tile] = s;
        }
        grid.sync();
        // ---- normalise + round (calculate_tpm!, src/quant.jl:655-667); every block folds the tile sums itself ----
        buf[t] = (uint32_t)t < d.ntiles ? d.partial[t] : 0.0;
        double rpk = wq_tile_tree(buf, t);
        rpk = fmax(rpk, 1.0);
        if (t == 0) s_changed = 0;
        __syncthreads();
        int mine = 0;
        for (uint32_t i = gtid; i < d.I; i += gsz) {
            double v = __ddiv_rn(__dmul_rn(d.tpm_raw[i], 1000000.0), rpk);
            if (d.sig > 0) v = wq_round_sc(v, d.sc_tpm);
            if (v != d.tpm[i]) mine = 1;
            d.tpm[i] = v;
        }
        if (mine) s_changed = 1;
        __syncthreads();
        if (t == 0 && s_changed) atomicOr(&d.flags[it % 3], 1);
        grid.sync();
        differs = d.flags[it % 3] != 0;
        it += 1;
    }
    if (gtid == 0) *d.it_out = it;
}

// grow-only device scratch kept in the handle (cudaMalloc/cudaFree per call would serialise the device)
struct WqEmScratch {
    void* p[16] = {nullptr}; size_t cap[16] = {0};
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
    void release() { for (int i = 0; i < 16; i++) { if (p[i]) cudaFree(p[i]); p[i] = nullptr; cap[i] = 0; } }
};

#define WQ_EMCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return std::string(#call) + ": " + cudaGetErrorString(e_); } while (0)

static inline std::string wq_run_em_tpm(const WqHostIndex& H, const WqIsoIndex& S, WqHostQuant& hq, WqHostQuantEx& X, WqEmScratch& M,
                                        const wq_em_param& p, cudaStream_t stream, uint64_t* launches,
                                        double* tpm_out, double* count_out, double* genetpm_out, int64_t* iterations) {
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
    const uint32_t I = S.I;
    WqClassSet& C = X.cls;
    const uint32_t NC = (uint32_t)C.count.size();
    const uint32_t ntiles = (I + 1023) / 1024;
    if (ntiles > 1024) return "more than 1048576 isoforms";
    std::vector<double> leff(I), tpm(I), count = X.iso_count;
    for (uint32_t i = 0; i < I; i++) leff[i] = std::max((double)(S.length[i] - p.readlen), 1.0);
    // contributions (isoform -> class positions) in class order
    std::vector<uint32_t> con_ptr(I + 1, 0), con_cls(C.iso.size()), con_pos(C.iso.size());
    for (int32_t i : C.iso) con_ptr[i + 1]++;
    for (uint32_t i = 0; i < I; i++) con_ptr[i + 1] += con_ptr[i];
    {
        std::vector<uint32_t> cur(con_ptr.begin(), con_ptr.end() - 1);
        for (uint32_t c = 0; c < NC; c++)
            for (uint32_t k = C.ptr[c]; k < C.ptr[c + 1]; k++) { uint32_t s = cur[C.iso[k]]++; con_cls[s] = c; con_pos[s] = k; }
    }
    WqEmDev d;
    double *d_leff, *d_count, *d_tpm, *d_raw, *d_prop, *d_ccount, *d_partial;
    uint32_t *d_cptr, *d_conptr, *d_concls, *d_conpos;
    int32_t* d_ciso;
    uint8_t *d_state, *d_parked;
    int* d_flags;
    int64_t* d_it;
    WQ_EMCK(M.put(0, leff, &d_leff, stream)); WQ_EMCK(M.put(1, count, &d_count, stream));
    WQ_EMCK(M.get(2, I, &d_tpm)); WQ_EMCK(M.get(3, I, &d_raw));
    WQ_EMCK(M.put(4, C.prop, &d_prop, stream)); WQ_EMCK(M.put(5, C.count, &d_ccount, stream)); WQ_EMCK(M.get(6, 1024, &d_partial));
    WQ_EMCK(M.put(7, C.ptr, &d_cptr, stream)); WQ_EMCK(M.put(8, con_ptr, &d_conptr, stream)); WQ_EMCK(M.put(9, con_cls, &d_concls, stream));
    WQ_EMCK(M.put(10, con_pos, &d_conpos, stream)); WQ_EMCK(M.put(11, C.iso, &d_ciso, stream)); WQ_EMCK(M.put(12, C.state, &d_state, stream));
    WQ_EMCK(M.get(13, 4, &d_flags)); WQ_EMCK(M.get(14, 1, &d_it));
    std::vector<uint8_t> parked(NC, 2);
    WQ_EMCK(M.put(15, parked, &d_parked, stream));
    WQ_EMCK(cudaMemsetAsync(d_tpm, 0, std::max<size_t>(1, I) * 8, stream));
    WQ_EMCK(cudaMemsetAsync(d_flags, 0, 4 * sizeof(int), stream));
    d.I = I; d.C = NC; d.ntiles = ntiles; d.leff = d_leff; d.count = d_count; d.tpm = d_tpm; d.tpm_raw = d_raw;
    d.cls_ptr = d_cptr; d.cls_iso = d_ciso; d.prop = d_prop; d.cls_count = d_ccount; d.state = d_state;
    d.con_ptr = d_conptr; d.con_cls = d_concls; d.con_pos = d_conpos; d.partial = d_partial; d.flags = d_flags;
    d.it_out = d_it; d.sig = (int)p.sig; d.sc_tpm = std::pow(10.0, (double)p.sig); d.sc_prop = 10000.0;
    int dev = 0, sms = 0, per_sm = 0;
    WQ_EMCK(cudaGetDevice(&dev));
    WQ_EMCK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    WQ_EMCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wq_k_em_tpm, 1024, 0));
    if (per_sm < 1) return "EM kernel does not fit on an SM";
    int grid = sms * per_sm;
    if (const char* ev = getenv("WQ_EM_BLOCKS")) grid = std::max(1, std::min(grid, atoi(ev)));
    // calculate_tpm!(quant, readlen) before gene_em! (bin/whippet-quant.jl:163): one sweep with every class parked
    {
        WqEmDev d0 = d;
        d0.state = d_parked; d0.maxit = 2; d0.first_differs = 1;
        void* args0[] = {&d0};
        WQ_EMCK(cudaLaunchCooperativeKernel((void*)wq_k_em_tpm, dim3(grid), dim3(1024), args0, 0, stream));
        if (launches) *launches += 1;
    }
    WQ_EMCK(cudaMemcpyAsync(tpm.data(), d_tpm, I * 8, cudaMemcpyDeviceToHost, stream));
    WQ_EMCK(cudaStreamSynchronize(stream));
    bool differs = false;                        // `tpm_temp != quant.tpm` with tpm_temp = ones (src/quant.jl:724,752)
    for (uint32_t i = 0; i < I; i++) if (tpm[i] != 1.0) { differs = true; break; }
    d.maxit = p.maxit; d.first_differs = differs ? 1 : 0;
    WQ_EMCK(cudaMemsetAsync(d_flags, 0, 4 * sizeof(int), stream));
    void* args[] = {&d};
    auto tk0 = std::chrono::steady_clock::now();
    WQ_EMCK(cudaLaunchCooperativeKernel((void*)wq_k_em_tpm, dim3(grid), dim3(1024), args, 0, stream));
    if (launches) *launches += 1;
    int64_t it = 0;
    WQ_EMCK(cudaMemcpyAsync(&it, d_it, 8, cudaMemcpyDeviceToHost, stream));
    WQ_EMCK(cudaMemcpyAsync(tpm.data(), d_tpm, I * 8, cudaMemcpyDeviceToHost, stream));
    WQ_EMCK(cudaMemcpyAsync(count.data(), d_count, I * 8, cudaMemcpyDeviceToHost, stream));
    WQ_EMCK(cudaMemcpyAsync(C.prop.data(), d_prop, C.prop.size() * 8, cudaMemcpyDeviceToHost, stream));
    WQ_EMCK(cudaMemcpyAsync(C.state.data(), d_state, C.state.size(), cudaMemcpyDeviceToHost, stream));
    WQ_EMCK(cudaStreamSynchronize(stream));
    if (prof) fprintf(stderr, "[wq]   em kernel (grid %d, %lld it) %9.3f ms\n", grid, (long long)it,
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tk0).count());
    X.tpm = tpm; X.iso_count = count;
    X.genetpm.assign(H.G, 0.0);                  // set_gene_tpm!, src/quant.jl:248-257
    for (uint32_t g = 0; g < H.G; g++) {
        double s = 0.0;
        for (uint32_t i = S.iso_base[g]; i < S.iso_base[g + 1]; i++) s += tpm[i];
        X.genetpm[g] = s;
    }
    X.em_done = true;
    X.iterations = it;
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
};

__device__ __forceinline__ double wq_pow10(const WqPsiDev& d, int k) { return (k >= 0 && k <= 22) ? d.p10[k] : pow(10.0, (double)k); }
// Base.round(x, sigdigits = sig)
__device__ __forceinline__ double wq_round_sig(const WqPsiDev& d, double x, int sig) {
    if (!isfinite(x) || x == 0.0) return x;
    int h = 1 + (int)floor(log10(fabs(x)));
    int digits = sig - h;
    double r;
    if (digits >= 0) { double sc = wq_pow10(d, digits); r = __ddiv_rn(rint(__dmul_rn(x, sc)), sc); }
    else { double sc = wq_pow10(d, -digits); r = __dmul_rn(rint(__ddiv_rn(x, sc)), sc); }
    if (!isfinite(r)) return digits > 0 ? x : r;
    return r;
}

// One warp per event.  Element-wise work (copies, the division + significant-digit rounding of every path,
// the additions of one ambiguous set, whose paths are distinct) is spread over the lanes; every sum is formed
// sequentially in path order by all lanes redundantly, so the arithmetic is the per-event sequential one.
// Small events keep psi / previous psi / temp counts in a shared-memory slab, larger ones in global scratch.
#define WQ_PSI_SLAB 48            // paths per warp slab
__global__ void __launch_bounds__(128) wq_k_em_psi(const WqPsiDev d) {
    __shared__ double slab[4][3 * WQ_PSI_SLAB];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t e = blockIdx.x * 4 + warp;
    if (e >= d.n_events) return;
    const uint32_t p0 = d.path_ptr[e], p1 = d.path_ptr[e + 1], ni = d.n_inc[e];
    const uint32_t a0 = d.amb_ptr[e], a1 = d.amb_ptr[e + 1];
    const bool tandem = d.is_tandem[e] != 0;
    const uint32_t np = p1 - p0;
    const bool small = np <= WQ_PSI_SLAB;
    double* psi = small ? &slab[warp][0] : d.psi + p0;
    double* prev = small ? &slab[warp][WQ_PSI_SLAB] : d.prev + p0;
    double* ct = small ? &slab[warp][2 * WQ_PSI_SLAB] : d.ctemp + p0;
    const double* cnt = d.count + p0;
    const double* len = d.length + p0;
    auto normalise = [&](int sig) {
        for (uint32_t i = lane; i < np; i += 32)
            psi[i] = tandem ? __ddiv_rn(ct[i], fmax(__dsub_rn(len[i], d.readlen), 1.0)) : __ddiv_rn(ct[i], len[i]);
        __syncwarp();
        double total;
        if (tandem) { total = 0.0; for (uint32_t i = 0; i < np; i++) total = __dadd_rn(total, psi[i]); }
        else {
            double se = 0.0, si = 0.0;
            for (uint32_t i = ni; i < np; i++) se = __dadd_rn(se, psi[i]);
            for (uint32_t i = 0; i < ni; i++) si = __dadd_rn(si, psi[i]);
            total = __dadd_rn(se, si);
        }
        __syncwarp();
        for (uint32_t i = lane; i < np; i += 32) {
            double q = __ddiv_rn(psi[i], total);
            psi[i] = sig > 0 ? wq_round_sig(d, q, sig) : q;
        }
        __syncwarp();
    };
    auto differs = [&]() {
        bool diff = false;
        for (uint32_t i = lane; i < np; i += 32) if (prev[i] != psi[i]) diff = true;
        return __any_sync(0xffffffffu, diff) != 0;
    };
    for (uint32_t i = lane; i < np; i += 32) { prev[i] = 0.0; psi[i] = 0.0; ct[i] = cnt[i]; }
    __syncwarp();
    if (!tandem) normalise(0);                      // calculate_psi!(inc, exc, counts) before spliced_em!
    int64_t it = 1;
    for (;;) {
        if (!tandem && !(differs() && it < d.maxit)) break;
        for (uint32_t i = lane; i < np; i += 32) { ct[i] = cnt[i]; prev[i] = psi[i]; }
        __syncwarp();
        for (uint32_t a = a0; a < a1; a++) {
            const uint32_t b = d.amb_path_ptr[a], n = d.amb_path_ptr[a + 1] - b;
            double psum = 1.0;
            if (it > 1) {
                psum = 0.0;
                for (uint32_t k = 0; k < n; k++) psum = __dadd_rn(psum, psi[d.amb_paths[b + k]]);
            }
            const double mult = d.amb_mult[a];
            for (uint32_t k = lane; k < n; k += 32) {
                const uint32_t pth = d.amb_paths[b + k];
                // first sweep: ones(n)/n over prop_sum = 1.0; later sweeps: current psi over their sum (src/events.jl:869-886)
                const double base = it > 1 ? psi[pth] : __ddiv_rn(1.0, (double)n);
                const double prop = __ddiv_rn(base, psum);
                ct[pth] = __dadd_rn(ct[pth], __dmul_rn(prop, mult));
            }
            __syncwarp();
        }
        normalise(d.sig);
        if (tandem) { if (differs() && it < d.maxit) it += 1; else break; }
        else it += 1;
    }
    if (small) for (uint32_t i = lane; i < np; i += 32) { d.psi[p0 + i] = psi[i]; d.ctemp[p0 + i] = ct[i]; }
    if (lane == 0) d.iterations[e] = (int32_t)it;
}

static inline std::string wq_run_em_psi(wq_psi_batch& b, WqEmScratch& M, cudaStream_t stream, uint64_t* launches) {
    const uint32_t E = b.n_events;
    if (E == 0) return "";
    if (!b.path_ptr || !b.n_inc || !b.amb_ptr || !b.amb_path_ptr || !b.is_tandem || !b.psi || !b.iterations) return "wq_psi_batch has null arrays";
    const uint32_t P = b.path_ptr[E], A = b.amb_ptr[E], AP = b.amb_path_ptr[A];
    WqPsiDev d;
    d.n_events = E; d.readlen = (double)b.readlen; d.sig = (int)b.sig; d.maxit = b.maxit;
    for (int k = 0; k <= 22; k++) d.p10[k] = std::pow(10.0, (double)k);
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
    wq_k_em_psi<<<(E + 3) / 4, 128, 0, stream>>>(d);
    if (launches) *launches += 1;
    WQ_EMCK(cudaGetLastError());
    WQ_EMCK(cudaMemcpyAsync(b.psi, d_psi, P * 8ull, cudaMemcpyDeviceToHost, stream));
    if (b.path_total) WQ_EMCK(cudaMemcpyAsync(b.path_total, d_ct, P * 8ull, cudaMemcpyDeviceToHost, stream));
    WQ_EMCK(cudaMemcpyAsync(b.iterations, d_it, E * 4ull, cudaMemcpyDeviceToHost, stream));
    WQ_EMCK(cudaStreamSynchronize(stream));
    return "";
}
