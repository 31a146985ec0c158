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
#include <cmath>
#include <map>
#include <string>
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
static inline size_t wq_key_next(const std::vector<uint32_t>& k, size_t pos, WqKeyPath& o) {
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
static inline void wq_spread_path(const WqHostIndex& H, WqHostQuantEx& Q, const WqKeyPath& kp, double v) {
    const uint32_t nb = H.genes[kp.gene - 1].node_base;
    auto ekey = [&](uint32_t l, uint32_t r) { return ((uint64_t)kp.gene << 32) | ((uint64_t)(l & 0xFFFF) << 16) | (r & 0xFFFF); };
    std::vector<uint32_t> seen;
    if (kp.n == 1) { Q.node_value[nb + kp.nd[0] - 1] += v; seen.push_back(kp.nd[0]); }
    else for (uint32_t i = 0; i + 1 < kp.n; i++) { seen.push_back(kp.nd[i]); seen.push_back(kp.nd[i + 1]); Q.edge_value[ekey(kp.nd[i], kp.nd[i + 1])] += v; }
    if (!kp.rnd) return;
    auto in_seen = [&](uint32_t x) { for (uint32_t s : seen) if (s == x) return true; return false; };
    if (kp.rn == 1) { if (!in_seen(kp.rnd[0])) Q.node_value[nb + kp.rnd[0] - 1] += v; }
    else for (uint32_t i = 0; i + 1 < kp.rn; i++) {
        if (in_seen(kp.rnd[i]) && in_seen(kp.rnd[i + 1])) continue;
        Q.edge_value[ekey(kp.rnd[i], kp.rnd[i + 1])] += v;
    }
}

// Builds multi-map classes (first) and isoform classes, fills the unique isoform counts, and
// re-counts long paths onto their edges (assign_long=true).
static inline std::string wq_host_build_classes(const WqHostIndex& H, const WqIsoIndex& S, WqHostQuant& hq, WqHostQuantEx& X) {
    const uint32_t G = H.G, I = S.I;
    X = WqHostQuantEx();
    X.node_value.assign(hq.node.begin(), hq.node.end());
    X.edge_value = hq.edge;
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
    for (auto& kv : hq.keyed) {
        if (!wq_key_is_multi(kv.first)) continue;
        const uint32_t nh = kv.first[0] & 0xFFFF;
        size_t pos = 1;
        std::vector<int32_t> ids;
        bool all = true;
        for (uint32_t h = 0; h < nh && all; h++) {
            WqKeyPath kp;
            pos = wq_key_next(kv.first, pos, kp);
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
        C.count.push_back((double)kv.second);
        C.state.push_back(all ? 0 : 2);
        C.postassign.push_back(all ? 0 : 1);
        X.multi_keys.push_back(&kv.first);
    }
    C.n_multi = (uint32_t)C.count.size();
    // ---- isoform classes, gene by gene ----
    auto lit = hq.keyed.begin();
    std::vector<std::pair<const std::vector<uint32_t>*, uint64_t>> longs_by_gene_sorted;
    // long keys are sorted by (npath, gene, ...) — group them by gene first, keeping key order inside a gene
    std::vector<std::vector<std::pair<const std::vector<uint32_t>*, uint64_t>>> longs(G + 1);
    for (auto& kv : hq.keyed) if (!wq_key_is_multi(kv.first)) longs[kv.first[1]].push_back({&kv.first, kv.second});
    (void)lit;
    for (uint32_t g = 1; g <= G; g++) {
        const WqGeneBits& b = bits_of(g);
        const uint32_t nn = H.genes[g - 1].nn, nb = H.genes[g - 1].node_base, ib = S.iso_base[g - 1];
        std::map<std::vector<int32_t>, double> temp;
        std::vector<int32_t> set;
        auto account = [&](double value) {
            if (set.size() == 1) X.iso_count[ib + set[0] - 1] += value;
            else if (set.size() > 1) temp[set] += value;
        };
        for (uint32_t n = 1; n <= nn; n++) {
            set.clear();
            for (uint32_t p = 0; p < b.np; p++) if (b.has(p, n)) set.push_back((int32_t)p + 1);
            account((double)hq.node[nb + n - 1]);
        }
        for (auto it = hq.edge.lower_bound((uint64_t)g << 32); it != hq.edge.end() && (it->first >> 32) == g; ++it) {
            if (wq_edge_is_circ(it->first)) continue;
            uint32_t l = (it->first >> 16) & 0xFFFF, r = it->first & 0xFFFF;
            set.clear();
            for (uint32_t p = 0; p < b.np; p++) if (b.adjacent(p, l, r)) set.push_back((int32_t)p + 1);
            account(it->second);
        }
        for (auto& lk : longs[g]) {
            WqKeyPath kp;
            wq_key_next(*lk.first, 0, kp);
            set.clear();
            for (uint32_t p = 0; p < b.np; p++) if (wq_key_compatible(b, p, kp)) set.push_back((int32_t)p + 1);
            account((double)lk.second);
            wq_spread_path(H, X, kp, (double)lk.second);
        }
        for (auto& kv : temp) {
            if (!(kv.second > 0.0)) continue;
            if (kv.first.size() > 500) return "isoform class larger than 500 (prop_temp = zeros(500), src/quant.jl:725)";
            for (int32_t p : kv.first) { C.iso.push_back((int32_t)(ib + p - 1)); C.prop.push_back(1.0 / (double)kv.first.size()); }
            C.ptr.push_back((uint32_t)C.iso.size());
            C.count.push_back(kv.second);
            C.state.push_back(0);
            C.postassign.push_back(0);
        }
    }
    X.built = true;
    return "";
}

// assign_ambig!, src/quant.jl:520-553 (host: one pass over the multi-map classes)
static inline std::string wq_host_assign_ambig_impl(const WqHostIndex& H, const WqIsoIndex& S, WqHostQuantEx& X) {
    if (!X.built || !X.em_done) return "assign_ambig needs wq_em_tpm first (it uses the EM class proportions and gene TPM)";
    if (X.ambig_done) return "";
    const WqClassSet& C = X.cls;
    WqGeneBits B;
    uint32_t cached = 0;
    for (uint32_t c = 0; c < C.n_multi; c++) {
        const std::vector<uint32_t>& key = *X.multi_keys[c];
        const uint32_t nh = key[0] & 0xFFFF;
        std::vector<WqKeyPath> vp(nh);
        size_t pos = 1;
        for (uint32_t h = 0; h < nh; h++) pos = wq_key_next(key, pos, vp[h]);
        std::vector<double> w(nh, 0.0);
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
        for (double x : w) tot += x;
        for (uint32_t h = 0; h < nh; h++) wq_spread_path(H, X, vp[h], (w[h] / tot) * C.count[c]);
    }
    X.ambig_done = true;
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

template <class T>
struct WqTmpDev {
    T* p = nullptr;
    cudaError_t put(const std::vector<T>& v, cudaStream_t s) {
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(1, v.size()) * sizeof(T));
        if (e != cudaSuccess) return e;
        if (!v.empty()) e = cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
        return e;
    }
    cudaError_t zero(size_t n) {
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(1, n) * sizeof(T));
        if (e == cudaSuccess) e = cudaMemset(p, 0, std::max<size_t>(1, n) * sizeof(T));
        return e;
    }
    ~WqTmpDev() { if (p) cudaFree(p); }
};

#define WQ_EMCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return std::string(#call) + ": " + cudaGetErrorString(e_); } while (0)

static inline std::string wq_run_em_tpm(const WqHostIndex& H, const WqIsoIndex& S, WqHostQuant& hq, WqHostQuantEx& X,
                                        const wq_em_param& p, cudaStream_t stream, uint64_t* launches,
                                        double* tpm_out, double* count_out, double* genetpm_out, int64_t* iterations) {
    if (p.sig < 0 || p.sig > 12) return "sig out of range";
    if (!X.built) { std::string e = wq_host_build_classes(H, S, hq, X); if (!e.empty()) return e; }
    if (X.em_done) return "wq_em_tpm was already run on this count state (gene_em! mutates it); call wq_quant_reset first";
    const uint32_t I = S.I;
    WqClassSet& C = X.cls;
    const uint32_t NC = (uint32_t)C.count.size();
    const uint32_t ntiles = (I + 1023) / 1024;
    if (ntiles > 1024) return "more than 1048576 isoforms";
    // calculate_tpm!(quant, readlen) on the unique counts (host: one pass; same arithmetic as the kernel)
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
    WqTmpDev<double> d_leff, d_count, d_tpm, d_raw, d_prop, d_ccount, d_partial;
    WqTmpDev<uint32_t> d_cptr, d_conptr, d_concls, d_conpos;
    WqTmpDev<int32_t> d_ciso;
    WqTmpDev<uint8_t> d_state;
    WqTmpDev<int> d_flags;
    WqTmpDev<int64_t> d_it;
    WQ_EMCK(d_leff.put(leff, stream)); WQ_EMCK(d_count.put(count, stream)); WQ_EMCK(d_tpm.zero(I)); WQ_EMCK(d_raw.zero(I));
    WQ_EMCK(d_prop.put(C.prop, stream)); WQ_EMCK(d_ccount.put(C.count, stream)); WQ_EMCK(d_partial.zero(1024));
    WQ_EMCK(d_cptr.put(C.ptr, stream)); WQ_EMCK(d_conptr.put(con_ptr, stream)); WQ_EMCK(d_concls.put(con_cls, stream));
    WQ_EMCK(d_conpos.put(con_pos, stream)); WQ_EMCK(d_ciso.put(C.iso, stream)); WQ_EMCK(d_state.put(C.state, stream));
    WQ_EMCK(d_flags.zero(4)); WQ_EMCK(d_it.zero(1));
    WqEmDev d;
    d.I = I; d.C = NC; d.ntiles = ntiles; d.leff = d_leff.p; d.count = d_count.p; d.tpm = d_tpm.p; d.tpm_raw = d_raw.p;
    d.cls_ptr = d_cptr.p; d.cls_iso = d_ciso.p; d.prop = d_prop.p; d.cls_count = d_ccount.p; d.state = d_state.p;
    d.con_ptr = d_conptr.p; d.con_cls = d_concls.p; d.con_pos = d_conpos.p; d.partial = d_partial.p; d.flags = d_flags.p;
    d.it_out = d_it.p; d.sig = (int)p.sig; d.sc_tpm = std::pow(10.0, (double)p.sig); d.sc_prop = 10000.0;
    int dev = 0, sms = 0, per_sm = 0;
    WQ_EMCK(cudaGetDevice(&dev));
    WQ_EMCK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    WQ_EMCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wq_k_em_tpm, 1024, 0));
    if (per_sm < 1) return "EM kernel does not fit on an SM";
    int grid = sms * per_sm;
    // pass 0: calculate_tpm!(quant, readlen=...) before gene_em! (bin/whippet-quant.jl:163) = one sweep with no
    // active class; run it as a 2-iteration call whose classes are all parked, then the real call.
    {
        std::vector<uint8_t> parked(NC, 2);
        WqTmpDev<uint8_t> d_parked;
        WQ_EMCK(d_parked.put(parked, stream));
        WqEmDev d0 = d;
        d0.state = d_parked.p; d0.maxit = 2; d0.first_differs = 1;
        void* args0[] = {&d0};
        WQ_EMCK(cudaLaunchCooperativeKernel((void*)wq_k_em_tpm, dim3(grid), dim3(1024), args0, 0, stream));
        WQ_EMCK(cudaStreamSynchronize(stream));
        if (launches) *launches += 1;
    }
    WQ_EMCK(cudaMemcpy(tpm.data(), d_tpm.p, I * 8, cudaMemcpyDeviceToHost));
    bool differs = false;                        // `tpm_temp != quant.tpm` with tpm_temp = ones (src/quant.jl:724,752)
    for (uint32_t i = 0; i < I; i++) if (tpm[i] != 1.0) { differs = true; break; }
    d.maxit = p.maxit; d.first_differs = differs ? 1 : 0;
    WQ_EMCK(cudaMemset(d_flags.p, 0, 4 * sizeof(int)));
    void* args[] = {&d};
    WQ_EMCK(cudaLaunchCooperativeKernel((void*)wq_k_em_tpm, dim3(grid), dim3(1024), args, 0, stream));
    WQ_EMCK(cudaStreamSynchronize(stream));
    if (launches) *launches += 1;
    int64_t it = 0;
    WQ_EMCK(cudaMemcpy(&it, d_it.p, 8, cudaMemcpyDeviceToHost));
    WQ_EMCK(cudaMemcpy(tpm.data(), d_tpm.p, I * 8, cudaMemcpyDeviceToHost));
    WQ_EMCK(cudaMemcpy(count.data(), d_count.p, I * 8, cudaMemcpyDeviceToHost));
    WQ_EMCK(cudaMemcpy(C.prop.data(), d_prop.p, C.prop.size() * 8, cudaMemcpyDeviceToHost));
    WQ_EMCK(cudaMemcpy(C.state.data(), d_state.p, C.state.size(), cudaMemcpyDeviceToHost));
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

static inline std::string wq_run_em_psi(wq_psi_batch&, cudaStream_t) { return "PSI EM is not built yet"; }
