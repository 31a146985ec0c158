// TEST INFRASTRUCTURE ONLY: drives the HOST stages of the library (class building, class shares, event construction:
// whippet.jl_b200/csrc/wq_em.cuh host part, wq_events.h) on given count stores, so they can be checked against the oracle
// in a container without a GPU.  No kernel is launched here and nothing of this file is part of the product library.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "../../whippet.jl_b200/csrc/wq_kernels.cuh"
#include "../../whippet.jl_b200/csrc/wq_host_index.h"
#include "../../whippet.jl_b200/csrc/wq_host_quant.h"
#include "../../whippet.jl_b200/csrc/wq_em.cuh"
#include "../../whippet.jl_b200/csrc/wq_events.h"
#include "../../whippet.jl_b200/csrc/wq_sam.h"
#include "../../whippet.jl_b200/csrc/wq_inflate.h"

struct HostEmu {
    WqHostIndex H; WqIsoIndex S; WqHostQuant hq; WqHostQuantEx X;
    std::vector<uint8_t> pack;
    std::vector<double> node_value;
    std::vector<WqEventRec> recs; WqPsiProblems P; WqEventPool pool;
    std::string err;
};

extern "C" {

static int g_event_pieces = 0;
static double* g_gene_times = nullptr;      // optional [G + 1] milliseconds per gene of the last hemu_events (profiling aid)
void hemu_set_gene_times(double* p) { g_gene_times = p; }
void hemu_set_event_pieces(int k) { g_event_pieces = k; }
HostEmu* hemu_create(const wq_index_desc* d) {
    HostEmu* e = new HostEmu();
    std::string err = wq_build_host_index(d, e->H);
    if (err.empty()) err = wq_build_iso_index(d, e->S);
    if (!err.empty()) { fprintf(stderr, "hemu: %s\n", err.c_str()); delete e; return nullptr; }
    return e;
}
void hemu_destroy(HostEmu* e) { delete e; }
#ifdef WQ_EV_PROF
void hemu_prof(unsigned long long* out, int reset) { for (int i = 0; i < 8; i++) { out[i] = wq_evp[i]; if (reset) wq_evp[i] = 0; } }
void hemu_prof_counts(unsigned long long* out, int reset) { for (int i = 0; i < 8; i++) { out[i] = wq_evc[i]; if (reset) wq_evc[i] = 0; } }
#endif
const char* hemu_error(HostEmu* e) { return e->err.c_str(); }

// count stores in the layout of wq_counts / wq_keyed (edges sorted by key; keyed records in any order)
void hemu_set_counts(HostEmu* e, const uint64_t* node, uint64_t ne, const uint64_t* ekey, const uint64_t* ecount,
                     uint64_t nl, const uint64_t* lptr, const uint32_t* lwords, const uint64_t* lcount,
                     uint64_t nm, const uint64_t* mptr, const uint32_t* mwords, const uint64_t* mcount) {
    e->hq = WqHostQuant();
    e->hq.node.assign(node, node + e->H.N);
    e->hq.edge_key.assign(ekey, ekey + ne); e->hq.edge_count.assign(ecount, ecount + ne);
    for (uint64_t i = 0; i < nl; i++) e->hq.longs.push(lwords + lptr[i], (size_t)(lptr[i + 1] - lptr[i]), lcount[i]);
    for (uint64_t i = 0; i < nm; i++) e->hq.multis.push(mwords + mptr[i], (size_t)(mptr[i + 1] - mptr[i]), mcount[i]);
    e->hq.multis.canonicalize();
}

// --biascorrect: the Float64 side of the stores just set (same orders; the multi-map records must have been given in canonical order)
void hemu_set_bias(HostEmu* e, const double* node_p, const double* node_g, const double* edge_p, const double* edge_g,
                   const double* long_p, const double* long_g, const double* multi_p, const double* multi_g) {
    WqHostQuant& Q = e->hq;
    Q.bias = true;
    Q.node_p.assign(node_p, node_p + Q.node.size()); Q.node_g.assign(node_g, node_g + Q.node.size());
    Q.edge_p.assign(edge_p, edge_p + Q.edge_key.size()); Q.edge_g.assign(edge_g, edge_g + Q.edge_key.size());
    Q.longs.bias_p.assign(long_p, long_p + Q.longs.size()); Q.longs.bias_g.assign(long_g, long_g + Q.longs.size());
    Q.multis.bias_p.assign(multi_p, multi_p + Q.multis.size()); Q.multis.bias_g.assign(multi_g, multi_g + Q.multis.size());
}

int hemu_build_classes(HostEmu* e, uint32_t part, uint32_t nparts) {
    e->err = wq_host_build_classes(e->H, e->S, e->hq, e->X, part, nparts);
    if (!e->err.empty()) return -1;
    if (nparts > 1) wq_classes_pack_impl(e->S, e->X, e->pack);
    return 0;
}
uint64_t hemu_pack(HostEmu* e, const void** p) { *p = e->pack.data(); return e->pack.size(); }
int hemu_assemble(HostEmu* e, const void* const* bufs, const uint64_t* lens, uint32_t n) {
    e->err = wq_classes_assemble_impl(e->S, e->X, bufs, lens, n);
    return e->err.empty() ? 0 : -1;
}
uint64_t hemu_n_classes(HostEmu* e, uint32_t* n_multi) { *n_multi = e->X.cls.n_multi; return e->X.cls.count.size(); }
uint64_t hemu_class(HostEmu* e, uint64_t c, int32_t* ids, uint64_t cap, double* count, int32_t* flags) {
    const WqClassSet& C = e->X.cls;
    const uint64_t n = C.ptr[c + 1] - C.ptr[c];
    for (uint64_t k = 0; k < n && k < cap; k++) ids[k] = C.iso[C.ptr[c] + k];
    *count = C.count[c];
    *flags = (C.state[c] ? 1 : 0) | (C.postassign[c] ? 2 : 0);
    return n;
}
void hemu_iso_count(HostEmu* e, double* out) { memcpy(out, e->X.iso_count.data(), e->X.iso_count.size() * 8); }
uint64_t hemu_edge_adds(HostEmu* e, uint64_t* keys, double* vals, uint64_t cap) {
    const uint64_t n = e->X.edge_adds.size();
    for (uint64_t i = 0; i < n && i < cap; i++) { keys[i] = e->X.edge_adds[i].first; vals[i] = e->X.edge_adds[i].second; }
    return n;
}

// event construction of every gene (src/events.jl:760-800) from the given node / edge values (sorted keys)
uint64_t hemu_events(HostEmu* e, const double* node_value, uint64_t ne, const uint64_t* ekey, const double* evalue, int64_t readlen) {
    e->node_value.assign(node_value, node_value + e->H.N);
    e->recs.clear(); e->P = WqPsiProblems(); e->pool = WqEventPool();
    WqGeneEvents V{e->H, e->node_value, 0, {}};
    uint64_t k = 0;
    for (uint32_t g = 1; g <= e->H.G; g++) {
        V.g = g;
        V.edges.clear();
        while (k < ne && (ekey[k] >> 32) < g) k++;
        for (; k < ne && (ekey[k] >> 32) == g; k++) {
            if (wq_edge_is_circ(ekey[k])) continue;
            V.edges.push_back(WqGeneEdge{(uint32_t)((ekey[k] >> 16) & 0xFFFF), (uint32_t)(ekey[k] & 0xFFFF), evalue[k]});
        }
        // hemu_set_event_pieces(k): build every gene in pieces of k visited nodes, the way the library cuts genes with many nodes
        // over its host threads (wq_gene_visits); the records must not depend on the cut
        const int pieces = g_event_pieces;
        const auto tg0 = std::chrono::steady_clock::now();
        if (pieces > 0) {
            std::vector<uint32_t> visits;
            wq_gene_visits(V, visits);
            for (size_t v = 0; v < visits.size(); v += (size_t)pieces)
                wq_gene_events(V, readlen, e->recs, e->P, e->pool, visits[v], visits[std::min(visits.size(), v + (size_t)pieces) - 1]);
        } else wq_gene_events(V, readlen, e->recs, e->P, e->pool);
        if (g_gene_times) g_gene_times[g] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tg0).count();
    }
    return e->recs.size();
}
// the event stage's work list and the ranks' cuts for the edges given (sorted keys): pieces as rows (g_lo, g_hi, node_lo, node_hi),
// cum[npieces + 1] the running cost estimate, cuts[nparts + 1].  Returns the number of pieces (call with out == NULL first).
uint64_t hemu_event_cuts(HostEmu* e, uint64_t ne, const uint64_t* ekey, const double* evalue, unsigned nt, uint32_t nparts, uint32_t* pieces_out, double* cum_out, uint32_t* cuts_out) {
    std::vector<WqEvPiece> pieces;
    if (e->node_value.size() != e->H.N) e->node_value.assign(e->H.N, 0.0);
    wq_event_make_pieces(e->H, e->node_value, nt, nparts, pieces);
    if (!pieces_out) return pieces.size();
    std::vector<uint64_t> keys(ekey, ekey + ne);
    std::vector<double> vals(evalue, evalue + ne);
    std::vector<unsigned> cuts; std::vector<double> cum;
    wq_event_cuts(e->H, keys, vals, pieces, nparts, cuts, &cum);
    for (size_t t = 0; t < pieces.size(); t++) { pieces_out[4 * t] = pieces[t].g_lo; pieces_out[4 * t + 1] = pieces[t].g_hi; pieces_out[4 * t + 2] = pieces[t].node_lo; pieces_out[4 * t + 3] = pieces[t].node_hi; }
    for (size_t t = 0; t < cum.size(); t++) cum_out[t] = cum[t];
    for (size_t r = 0; r < cuts.size(); r++) cuts_out[r] = cuts[r];
    return pieces.size();
}
// event construction of ONE rank's range of the work list, the way wq_build_event_batch walks it (pieces from wq_event_make_pieces, cuts
// from wq_event_cuts on the given edges, gene ranges and node ranges of genes with many nodes); replaces the records of the last call
uint64_t hemu_events_part(HostEmu* e, const double* node_value, uint64_t ne, const uint64_t* ekey, const double* evalue, int64_t readlen,
                          unsigned nt, uint32_t nparts, uint32_t part) {
    e->node_value.assign(node_value, node_value + e->H.N);
    e->recs.clear(); e->P = WqPsiProblems(); e->pool = WqEventPool();
    std::vector<WqEvPiece> pieces;
    wq_event_make_pieces(e->H, e->node_value, nt, nparts, pieces);
    std::vector<uint64_t> keys(ekey, ekey + ne);
    std::vector<double> vals(evalue, evalue + ne);
    std::vector<unsigned> cuts;
    wq_event_cuts(e->H, keys, vals, pieces, nparts, cuts);
    WqGeneEvents V{e->H, e->node_value, 0, {}};
    for (unsigned t = cuts[part]; t < cuts[part + 1]; t++) {
        const WqEvPiece pc = pieces[t];
        size_t k = std::lower_bound(keys.begin(), keys.end(), (uint64_t)pc.g_lo << 32) - keys.begin();
        for (uint32_t g = pc.g_lo; g <= pc.g_hi && g <= e->H.G; g++) {
            V.g = g;
            V.edges.clear();
            while (k < ne && (keys[k] >> 32) < g) k++;
            for (; k < ne && (keys[k] >> 32) == g; k++) {
                if (wq_edge_is_circ(keys[k])) continue;
                V.edges.push_back(WqGeneEdge{(uint32_t)((keys[k] >> 16) & 0xFFFF), (uint32_t)(keys[k] & 0xFFFF), vals[k]});
            }
            wq_gene_events(V, readlen, e->recs, e->P, e->pool, pc.node_lo, pc.node_hi);
        }
    }
    return e->recs.size();
}
// header[9] = gene, node, motif, kind, flags, n_utr, n_inc, n_exc, n_paths; vals[2] = psi, total; returns the EM problem index or -1
int64_t hemu_event(HostEmu* e, uint64_t i, uint32_t* header, double* vals) {
    const WqEventRec& r = e->recs[i];
    header[0] = r.gene; header[1] = r.node; header[2] = r.motif; header[3] = r.kind; header[4] = r.flags;
    header[5] = r.n_utr; header[6] = r.n_inc; header[7] = r.n_exc; header[8] = r.n_paths;
    vals[0] = r.psi; vals[1] = r.total;
    return (int64_t)r.problem;
}
uint64_t hemu_event_path(HostEmu* e, uint64_t i, uint64_t k, uint32_t* out, uint64_t cap) {
    const WqEventRec& r = e->recs[i];
    const uint64_t b = e->pool.path_ptr[r.path_first + k], en = e->pool.path_ptr[r.path_first + k + 1];
    for (uint64_t j = b; j < en && j - b < cap; j++) out[j - b] = e->pool.path_nodes[j];
    return en - b;
}
// EM problem p: sizes[4] = n_paths, n_inc, n_amb, tandem; then the arrays (two-call: pass NULLs for sizes only)
void hemu_problem(HostEmu* e, uint64_t p, uint32_t* sizes, double* count, double* length, uint32_t* amb_ptr, uint32_t* amb_paths, double* amb_mult) {
    const WqPsiProblems& P = e->P;
    const uint32_t p0 = P.path_ptr[p], p1 = P.path_ptr[p + 1], a0 = P.amb_ptr[p], a1 = P.amb_ptr[p + 1];
    sizes[0] = p1 - p0; sizes[1] = P.n_inc[p]; sizes[2] = a1 - a0; sizes[3] = P.is_tandem[p];
    sizes[4] = P.amb_path_ptr[a1] - P.amb_path_ptr[a0];
    if (!count) return;
    for (uint32_t i = p0; i < p1; i++) { count[i - p0] = P.count[i]; length[i - p0] = P.length[i]; }
    for (uint32_t a = a0; a <= a1; a++) amb_ptr[a - a0] = P.amb_path_ptr[a] - P.amb_path_ptr[a0];
    for (uint32_t a = a0; a < a1; a++) amb_mult[a - a0] = P.amb_mult[a];
    for (uint32_t j = P.amb_path_ptr[a0]; j < P.amb_path_ptr[a1]; j++) amb_paths[j - P.amb_path_ptr[a0]] = P.amb_paths[j];
}

// all EM problems of the last hemu_events call as the flat arrays of wq_psi_batch (sizes first: out == NULL)
void hemu_problems_flat(HostEmu* e, uint64_t* sizes4, uint32_t* path_ptr, uint32_t* n_inc, double* count, double* length, uint32_t* amb_ptr,
                        uint32_t* amb_path_ptr, uint32_t* amb_paths, double* amb_mult, uint8_t* is_tandem) {
    const WqPsiProblems& P = e->P;
    sizes4[0] = P.size(); sizes4[1] = P.count.size(); sizes4[2] = P.amb_mult.size(); sizes4[3] = P.amb_paths.size();
    if (!path_ptr) return;
    memcpy(path_ptr, P.path_ptr.data(), P.path_ptr.size() * 4); memcpy(n_inc, P.n_inc.data(), P.n_inc.size() * 4);
    memcpy(count, P.count.data(), P.count.size() * 8); memcpy(length, P.length.data(), P.length.size() * 8);
    memcpy(amb_ptr, P.amb_ptr.data(), P.amb_ptr.size() * 4); memcpy(amb_path_ptr, P.amb_path_ptr.data(), P.amb_path_ptr.size() * 4);
    memcpy(amb_paths, P.amb_paths.data(), P.amb_paths.size() * 4); memcpy(amb_mult, P.amb_mult.data(), P.amb_mult.size() * 8);
    memcpy(is_tandem, P.is_tandem.data(), P.is_tandem.size());
}

// The cluster PSI kernel (wq_k_em_psi_cl) on the CPU: the same phase functions the kernel calls, run for every block of the cluster
// and every thread in turn, a phase at a time (a cluster barrier separates the phases on the device; within a phase no thread reads
// what another writes -- that part only the GPU run can show).  Remote shared-memory writes go to the other blocks' buffers.
// Returns 0, or -1 when the layout does not fit the device's shared memory / a set names a path twice (the library would then keep
// these events on the block kernel).
int hemu_psi_cluster(uint32_t E, const uint32_t* path_ptr, const uint32_t* n_inc, const double* count, const double* length, const uint32_t* amb_ptr,
                     const uint32_t* amb_path_ptr, const uint32_t* amb_paths, const double* amb_mult, const uint8_t* is_tandem, int64_t readlen,
                     int sig, int64_t maxit, uint32_t C, uint32_t T, double* psi, double* ctemp, int32_t* iterations, uint64_t* smem_bytes) {
    wq_psi_batch b;
    memset(&b, 0, sizeof b);
    b.n_events = E; b.path_ptr = path_ptr; b.n_inc = n_inc; b.path_count = count; b.path_length = length; b.amb_ptr = amb_ptr;
    b.amb_path_ptr = amb_path_ptr; b.amb_paths = amb_paths; b.amb_mult = amb_mult; b.is_tandem = is_tandem;
    std::vector<uint32_t> events(E);
    for (uint32_t e = 0; e < E; e++) events[e] = e;
    if (!wq_psi_cl_unique(b, events.data(), E)) return -1;
    WqPsiClCfg z;
    const size_t need = wq_psi_cl_config(b, events.data(), E, C, z);
    if (smem_bytes) *smem_bytes = need;
    if (need > 227 * 1024) return -1;
    static const std::vector<double> big = [] { std::vector<double> t(309); for (int k = 0; k <= 308; k++) t[k] = std::pow(10.0, (double)k); return t; }();
    WqPsiDev d;
    memset(&d, 0, sizeof d);
    d.n_events = E; d.path_ptr = path_ptr; d.n_inc = n_inc; d.count = count; d.length = length; d.amb_ptr = amb_ptr; d.amb_path_ptr = amb_path_ptr;
    d.amb_paths = amb_paths; d.amb_mult = amb_mult; d.is_tandem = is_tandem; d.readlen = (double)readlen; d.sig = sig; d.maxit = maxit;
    d.psi = psi; d.ctemp = ctemp; d.iterations = iterations; d.p10big = big.data();
    for (int k = 0; k <= 22; k++) { d.p10[k] = std::pow(10.0, (double)k); d.ip10[k] = 1.0 / d.p10[k]; }
    std::vector<std::vector<unsigned char>> buf(C, std::vector<unsigned char>(need + 16, 0xAB));      // poisoned: nothing may rely on zeroed shared memory
    unsigned char* all[8];
    for (uint32_t r = 0; r < C; r++) all[r] = buf[r].data();
    std::vector<WqClCtx> ctx(C);
    for (uint32_t r = 0; r < C; r++) { ctx[r].r = r; ctx[r].C = C; ctx[r].T = T; ctx[r].self = all[r]; ctx[r].all = all; wq_psi_cl_layout(z, all[r], &ctx[r].S); }
    auto each = [&](const std::function<void(const WqClCtx&, uint32_t)>& f) { for (uint32_t r = 0; r < C; r++) for (uint32_t t = 0; t < T; t++) f(ctx[r], t); };
    for (uint32_t e = 0; e < E; e++) {
        const WqPsiClEv ev = wq_psi_cl_event(d, e);
        each([&](const WqClCtx& c, uint32_t t) { wq_cl_init_a(d, ev, c, t); });
        each([&](const WqClCtx& c, uint32_t t) { wq_cl_init_b(d, ev, c, t); });
        for (uint32_t r = 0; r < C; r++) wq_cl_init_c(ev, ctx[r]);
        for (uint32_t r = 0; r < C; r++)                                 // the layout's capacities must cover what the build produced
            if (ctx[r].S.poff[wq_cl_nslots(ev.np, r, C)] > z.inv_cap || wq_cl_nslots(ev.np, r, C) > z.npl_cap || wq_cl_nsets(ev.na, r, C) > z.nal_cap || ev.np > z.np_cap || ev.na > z.na_cap) return -2;
        for (uint32_t a = 0; a < ev.na; a++) each([&](const WqClCtx& c, uint32_t t) { wq_cl_init_d(d, ev, c, t, a); });
        bool differs = false;
        auto norm = [&](int sg) {
            for (uint32_t r = 0; r < C; r++) {
                bool any = false;
                for (uint32_t t = 0; t < T; t++) any |= wq_cl_phase_norm(d, ev, ctx[r], t, sg);
                for (uint32_t rr = 0; rr < C; rr++) *wq_cl_remote(ctx[r], ctx[r].S.flags + r, rr) = any ? 1u : 0u;
            }
            differs = false;
            for (uint32_t r = 0; r < C; r++) {                       // every block must see the same answer
                uint32_t f = 0;
                for (uint32_t rr = 0; rr < C; rr++) f |= ctx[r].S.flags[rr];
                if (r == 0) differs = f != 0; else if (differs != (f != 0)) abort();
            }
        };
        if (!ev.tandem) {
            each([&](const WqClCtx& c, uint32_t t) { wq_cl_phase_pre(ev, c, t); });
            each([&](const WqClCtx& c, uint32_t t) { wq_cl_phase_total(ev, c, t); });
            norm(0);
        }
        int64_t it = 1;
        for (;;) {
            if (!ev.tandem && !(differs && it < d.maxit)) break;
            if (it > 1) each([&](const WqClCtx& c, uint32_t t) { wq_cl_phase_psum(ev, c, t); });
            each([&](const WqClCtx& c, uint32_t t) { wq_cl_phase_shares(d, ev, c, t, it == 1); });
            each([&](const WqClCtx& c, uint32_t t) { wq_cl_phase_total(ev, c, t); });
            norm(d.sig);
            if (ev.tandem) { if (differs && it < d.maxit) it += 1; else break; }
            else it += 1;
        }
        each([&](const WqClCtx& c, uint32_t t) { wq_cl_write_out(d, ev, c, t, it); });
    }
    return 0;
}

// WqInflate (csrc/wq_inflate.h) on a gzip file image: the text is read in pieces of `piece` bytes into out[cap].  Returns the
// number of bytes produced, or -1 - (bytes produced) when the stream ended in an error (message in errbuf).
int64_t hemu_inflate(const uint8_t* data, uint64_t n, uint8_t* out, uint64_t cap, uint64_t piece, char* errbuf, uint64_t errcap, uint64_t* members) {
    WqInflate z;
    z.open(data, (size_t)n);
    uint64_t got = 0;
    for (;;) {
        const size_t want = (size_t)std::min<uint64_t>(piece, cap - got);
        if (!want) break;
        const size_t r = z.read(out + got, want);
        got += r;
        if (r < want) break;
    }
    if (members) *members = z.members;
    if (errbuf && errcap) { strncpy(errbuf, z.err.c_str(), errcap - 1); errbuf[errcap - 1] = 0; }
    return z.failed() ? -1 - (int64_t)got : (int64_t)got;
}

// wq_crc32 (csrc/wq_inflate.h: carry-less-multiplication CRC-32 with zlib for the tail), zlib's calling convention
uint32_t hemu_crc32(uint32_t crc, const uint8_t* p, uint64_t n) { return wq_crc32(crc, p, (size_t)n); }

// WqInflate::raw_into: one raw deflate stream into exactly out[0 .. out_len) (the way the blocked-gzip reader inflates a block next to
// its neighbours).  Returns 0, or -1 with the message in errbuf.
int hemu_inflate_raw(const uint8_t* data, uint64_t n, uint8_t* out, uint64_t out_len, char* errbuf, uint64_t errcap) {
    static thread_local WqInflate* z = nullptr;
    if (!z) z = new WqInflate();
    const bool ok = z->raw_into(data, (size_t)n, out, (size_t)out_len);
    if (errbuf && errcap) { strncpy(errbuf, z->err.c_str(), errcap - 1); errbuf[errcap - 1] = 0; }
    return ok ? 0 : -1;
}

// WqBitWin self-test: adjacent() (word masks) against the per-node definition (src/quant.jl:68-79), first() / last() /
// list() against a plain scan, on random node sets in windows of up to `max_span` nodes.  Returns the number of mismatches.
uint64_t hemu_bitwin_selftest(uint32_t max_span, uint32_t rounds, uint64_t seed) {
    uint64_t bad = 0, x = seed | 1;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t lo = 1 + (uint32_t)(rnd() % 50), span = 1 + (uint32_t)(rnd() % max_span), hi = lo + span - 1;
        WqBitWin w;
        w.init(lo, hi);
        std::vector<uint8_t> ref(span, 0);
        const uint32_t dens = 1 + (uint32_t)(rnd() % 8);
        for (uint32_t n = lo; n <= hi; n++) if (rnd() % 8 < dens) { w.set(n); ref[n - lo] = 1; }
        std::vector<uint32_t> lst, want;
        w.list(lst);
        for (uint32_t n = lo; n <= hi; n++) if (ref[n - lo]) want.push_back(n);
        if (lst != want) bad++;
        if (!want.empty() && (w.first() != want.front() || w.last() != want.back())) bad++;
        for (int t = 0; t < 64; t++) {
            uint32_t a = lo + (uint32_t)(rnd() % span), b = lo + (uint32_t)(rnd() % span);
            if (a > b) std::swap(a, b);
            bool naive = ref[a - lo] && ref[b - lo];
            for (uint32_t i = a + 1; naive && i < b; i++) if (ref[i - lo]) naive = false;
            if (w.adjacent(a, b) != naive) bad++;
        }
    }
    return bad;
}

// assign_ambig! (host stage) given the EM's results: class proportions in canonical order and gene TPM
int hemu_assign_ambig(HostEmu* e, const double* prop, uint64_t nprop, const double* genetpm) {
    if (nprop != e->X.cls.prop.size()) { e->err = "prop array has the wrong length"; return -1; }
    e->X.cls.prop.assign(prop, prop + nprop);
    e->X.genetpm.assign(genetpm, genetpm + e->H.G);
    e->X.em_done = true;
    e->err = wq_host_assign_ambig_impl(e->H, e->S, e->hq, e->X);
    if (!e->err.empty()) return -1;
    wq_finalize_edges(e->hq, e->X);
    return 0;
}
uint64_t hemu_values(HostEmu* e, double* node_value, uint64_t* ekey, double* evalue) {
    if (node_value) memcpy(node_value, e->X.node_value.data(), e->X.node_value.size() * 8);
    if (ekey) { memcpy(ekey, e->X.edge_key.data(), e->X.edge_key.size() * 8); memcpy(evalue, e->X.edge_value.data(), e->X.edge_value.size() * 8); }
    return e->X.edge_key.size();
}

// SAM text of a batch (wq_sam.h: write_sam, src/io.jl:69-271) from given alignments
struct HostSam { WqSamIndex X; std::string text; };
HostSam* hsam_create(const wq_index_desc* d, const char* const* seqname, const uint8_t* strand) {
    HostSam* s = new HostSam();
    s->X.node_base.assign(d->node_base, d->node_base + d->n_genes + 1);
    s->X.nodeoffset.assign(d->nodeoffset, d->nodeoffset + d->n_nodes);
    s->X.nodecoord.assign(d->nodecoord, d->nodecoord + d->n_nodes);
    s->X.nodelen.assign(d->nodelen, d->nodelen + d->n_nodes);
    for (uint32_t g = 0; g < d->n_genes; g++) { s->X.seqname.emplace_back(seqname[g]); s->X.strand.push_back(strand[g] ? 1 : 0); }
    s->X.has_info = true;
    return s;
}
void hsam_destroy(HostSam* s) { delete s; }
uint64_t hsam_batch(HostSam* s, const wq_read_batch* fwd, const wq_read_batch* rev, const uint64_t* name_off, const uint32_t* name_len, const char* names,
                    const uint64_t* rname_off, const uint32_t* rname_len, const char* rnames, const wq_align_out* out, int qualoffset) {
    s->text.clear();
    wq_sam_range(s->X, s->text, fwd, rev, name_off, name_len, names, rname_off, rname_len, rnames, out, qualoffset, 0, fwd->n_reads);
    return s->text.size();
}
const char* hsam_text(HostSam* s) { return s->text.data(); }
uint64_t hsam_header(HostSam* s) { s->text = wq_sam_header_text(s->X); return s->text.size(); }

}  // extern "C"
