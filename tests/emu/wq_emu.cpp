// TEST INFRASTRUCTURE ONLY: drives the per-thread kernel bodies of whippet.jl_b200/csrc/wq_kernels.cuh
// single-threaded on the CPU so their logic can be checked against the oracle in a container that
// has no GPU.  Never built into, nor loadable by, the product library (which has no CPU path).
#define WQ_EMU_STATS 1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../whippet.jl_b200/csrc/wq_kernels.cuh"
#include "../../whippet.jl_b200/csrc/wq_host_index.h"
#include "../../whippet.jl_b200/csrc/wq_bias.cuh"
#include <algorithm>

struct Emu {
    WqHostIndex H;
    WqDevIndex ix;
    std::vector<uint64_t> node_count;
    WqCounters counters;
    std::vector<uint64_t> ekey, ecount, khash, kcount;
    std::vector<uint32_t> koff, kpool;
    uint64_t pool_cursor = 0;
    WqQuantDev q;
    // last batch
    std::vector<WqSeed> seeds;
    std::vector<uint32_t> hit_read, hit_s, hit_s2, hit_gene, pending;
    bool paired = false;
    std::vector<WqHit> hits;
    std::vector<wq_align_node> pool;
    uint32_t cursors[8];
    std::vector<uint32_t> pending2, defer;
    std::vector<uint2> ftab;
    std::vector<WqPNode> path, side;     // first pass: WQ_FAST_PATH nodes, no side buffers; second pass: WQ_MAX_PATH + side buffers
    uint64_t deferred = 0;               // extension tasks that went to the second pass (so tests can see both passes ran)
    // --biascorrect (csrc/wq_bias.cuh)
    bool bias = false;
    std::vector<uint64_t> bias_tab, bias_rec;      // fore[4096] back[4096] cursor[2]; records
    std::vector<double> node_p, node_g, edge_p, edge_g, keyed_p, keyed_g, fore_n, back_n;
    WqBiasDev bias_dev(uint32_t n) {
        const uint64_t used = bias_tab[2 * WQ_BIAS_NHEX];
        if (bias_rec.size() < used + (uint64_t)n * WQ_BIAS_MAX_REC) bias_rec.resize(used + (uint64_t)n * WQ_BIAS_MAX_REC);
        WqBiasDev B;
        B.fore = bias_tab.data(); B.back = B.fore + WQ_BIAS_NHEX; B.cursor = B.back + WQ_BIAS_NHEX; B.rec = bias_rec.data(); B.cap = bias_rec.size();
        wq_bias_gc_table(B.gcbin);
        return B;
    }
};
static WqPathStore emu_store(Emu* e, bool first) {
    e->path.assign(WQ_MAX_PATH, WqPNode());
    e->side.assign((size_t)WQ_MAX_DEPTH * WQ_MAX_PATH, WqPNode());
    return first ? WqPathStore{e->path.data(), 1, WQ_FAST_PATH, nullptr, 0} : WqPathStore{e->path.data(), 1, WQ_MAX_PATH, e->side.data(), 1};
}

extern "C" {

Emu* emu_create(const wq_index_desc* d) {
    Emu* e = new Emu();
    std::string err = wq_build_host_index(d, e->H);
    if (!err.empty()) { fprintf(stderr, "emu: %s\n", err.c_str()); delete e; return nullptr; }
    WqHostIndex& H = e->H;
    WqDevIndex& ix = e->ix;
    ix.occ = H.occ.data(); ix.sa = d->sa; ix.text2 = H.text2.data(); ix.amb = H.has_amb ? H.amb.data() : nullptr;
    ix.nodes = H.nodes.data(); ix.genes = H.genes.data(); ix.node_bucket = H.node_bucket.data();
    ix.left_ptr = d->left_ptr; ix.left_ent = H.left_ent.data(); ix.right_ptr = d->right_ptr; ix.right_ent = H.right_ent.data();
    ix.n = H.n; ix.sentinel = H.sentinel; for (int i = 0; i < 4; i++) ix.C[i] = H.C[i];
    ix.N = H.N; ix.G = H.G; ix.K = H.K;
    // k-mer table of the seeding kernel, built by the same rule the library's wq_k_ftab kernel uses (smaller k: CPU time)
    ix.ftab = nullptr; ix.ftab_k = 0;
    const char* fk = getenv("WQ_EMU_FTAB_K");
    const int k = fk ? atoi(fk) : 7;
    if (k > 0) {
        e->ftab.resize((size_t)1 << (2 * k));
        for (uint32_t i = 0; i < e->ftab.size(); i++) e->ftab[i] = wq_ftab_entry(ix, i, k);
        ix.ftab = e->ftab.data(); ix.ftab_k = k;
    }
    e->node_count.assign(H.N, 0);
    memset(&e->counters, 0, sizeof e->counters);
    size_t es = 1 << 16, ks = 1 << 16;
    e->ekey.assign(es, WQ_EMPTY_KEY); e->ecount.assign(es, 0);
    e->khash.assign(ks, 0); e->kcount.assign(ks, 0); e->koff.assign(ks, WQ_KEYOFF_UNSET); e->kpool.assign(1 << 22, 0);
    e->q.node_count = e->node_count.data(); e->q.counters = &e->counters;
    e->q.edges = WqEdgeTable{e->ekey.data(), e->ecount.data(), es - 1};
    e->q.keyed = WqKeyedTable{e->khash.data(), e->koff.data(), e->kcount.data(), e->kpool.data(), &e->pool_cursor, e->kpool.size(), ks - 1};
    return e;
}
void emu_destroy(Emu* e) { delete e; }

int emu_align_se(Emu* e, const wq_align_param* ap, const wq_read_batch* rb) {
    WqParams p;
    std::string err = wq_make_params(ap, e->H.K, p);
    if (!err.empty()) { fprintf(stderr, "emu: %s\n", err.c_str()); return -1; }
    WqBatchDev b{rb->n_reads, rb->len, rb->seq_off, rb->seq2, rb->qual_off, rb->qual};
    uint32_t n = rb->n_reads;
    e->seeds.assign(n, WqSeed());
    uint32_t hit_cap = n * p.seed_tolerate + 1;
    e->hit_read.assign(hit_cap, 0); e->hit_s.assign(hit_cap, 0); e->hits.assign(hit_cap, WqHit());
    e->pool.assign((size_t)hit_cap * 4 + 64, wq_align_node());
    e->pending.assign(n + 1, 0);
    e->pending2.assign(n + 1, 0);
    for (auto& x : e->cursors) x = 0;
    e->paired = false;
    e->defer.assign(hit_cap, 0);
    WqWork wk{e->seeds.data(), e->hit_read.data(), e->hit_s.data(), nullptr, nullptr, e->hits.data(), e->pool.data(), e->cursors,
              e->pending.data(), e->pending2.data(), hit_cap, (uint32_t)e->pool.size(), e->defer.data(), nullptr};
    uint64_t wbuf[WQ_MAX_READ_WORDS + 1];
    for (uint32_t r = 0; r < n; r++) wq_seed_thread(e->ix, p, b, wk, r, wbuf);
    uint32_t nf = e->cursors[0], nr = e->cursors[5];
    const WqPathStore st1 = emu_store(e, true), st2 = emu_store(e, false);
    for (uint32_t t = 0; t < nf + nr; t++) wq_extend_thread(e->ix, p, b, wk, t < nf ? t : hit_cap - 1 - (t - nf), wbuf, st1, true);
    for (uint32_t t = 0; t < e->cursors[6]; t++) wq_extend_thread(e->ix, p, b, wk, e->defer[t], wbuf, st2, false);
    e->deferred += e->cursors[6];
    uint64_t stats[5] = {0, 0, 0, 0, 0};
    for (uint32_t r = 0; r < n; r++) wq_resolve_thread(e->ix, p, b, wk, e->q, r, stats);
    if (e->bias) { const WqBiasDev B = e->bias_dev(n); for (uint32_t r = 0; r < n; r++) wq_bias_thread(e->ix, b, wk, e->q, B, r); }
    e->counters.total += n; e->counters.mapped += stats[0]; e->counters.multi += stats[1];
    e->counters.sum_readlen += stats[2]; e->counters.path_overflow += stats[3]; e->counters.pool_overflow += stats[4];
    return (int)e->cursors[2];   // pending (always 0 single-threaded)
}

int emu_align_pe(Emu* e, const wq_align_param* ap, const wq_read_batch* rf, const wq_read_batch* rr) {
    WqParams p;
    std::string err = wq_make_params(ap, e->H.K, p);
    if (!err.empty()) { fprintf(stderr, "emu: %s\n", err.c_str()); return -1; }
    WqBatchDev bf{rf->n_reads, rf->len, rf->seq_off, rf->seq2, rf->qual_off, rf->qual};
    WqBatchDev br{rr->n_reads, rr->len, rr->seq_off, rr->seq2, rr->qual_off, rr->qual};
    uint32_t n = rf->n_reads;
    e->seeds.assign(n, WqSeed());
    uint32_t hit_cap = n * p.seed_tolerate + 1;
    e->hit_read.assign(hit_cap, 0); e->hit_s.assign(hit_cap, 0); e->hit_s2.assign(hit_cap, 0); e->hit_gene.assign(hit_cap, 0);
    e->hits.assign((size_t)2 * hit_cap, WqHit());
    e->pool.assign((size_t)hit_cap * 8 + 64, wq_align_node());
    e->pending.assign(n + 1, 0);
    for (auto& x : e->cursors) x = 0;
    e->paired = true;
    e->defer.assign((size_t)2 * hit_cap, 0);
    WqWork wk{e->seeds.data(), e->hit_read.data(), e->hit_s.data(), e->hit_s2.data(), e->hit_gene.data(), e->hits.data(),
              e->pool.data(), e->cursors, e->pending.data(), e->pending.data(), hit_cap, (uint32_t)e->pool.size(), e->defer.data(), nullptr};
    uint64_t wbuf[WQ_MAX_READ_WORDS + 1];
    for (uint32_t r = 0; r < n; r++) wq_seed_pe_thread(e->ix, p, bf, br, wk, r, wbuf);
    uint32_t nh = e->cursors[0];
    const WqPathStore st1 = emu_store(e, true), st2 = emu_store(e, false);
    for (uint32_t s = 0; s < nh; s++) { wq_extend_pe_thread(e->ix, p, bf, br, wk, s, 0, wbuf, st1, true); wq_extend_pe_thread(e->ix, p, bf, br, wk, s, 1, wbuf, st1, true); }
    for (uint32_t t = 0; t < e->cursors[6]; t++) wq_extend_pe_thread(e->ix, p, bf, br, wk, e->defer[t] >> 1, (int)(e->defer[t] & 1), wbuf, st2, false);
    e->deferred += e->cursors[6];
    uint64_t stats[5] = {0, 0, 0, 0, 0};
    for (uint32_t r = 0; r < n; r++) wq_resolve_pe_thread(e->ix, p, bf, br, wk, e->q, r, stats);
    if (e->bias) { const WqBiasDev B = e->bias_dev(n); for (uint32_t r = 0; r < n; r++) wq_bias_pe_thread(e->ix, bf, br, wk, e->q, B, r); }
    e->counters.total += n; e->counters.mapped += stats[0]; e->counters.multi += stats[1];
    e->counters.sum_readlen += stats[2]; e->counters.path_overflow += stats[3]; e->counters.pool_overflow += stats[4];
    return (int)e->cursors[2];
}

// results of the last batch in the oracle's CSR layout (kept alignments only, reference order)
uint64_t emu_result_sizes(Emu* e, uint64_t* n_nodes) {
    uint64_t na = 0, nn = 0;
    for (auto& s : e->seeds)
        for (int k = 0; k < s.cnt; k++) {
            if (e->paired) {
                const WqHit& hf = e->hits[2 * (s.hit_base + k)];
                const WqHit& hr = e->hits[2 * (s.hit_base + k) + 1];
                if (hf.flags & 4) { na++; nn += hf.npath + hr.npath; }
                continue;
            }
            const WqHit& h = e->hits[s.hit_base + k];
            if (h.flags & 4) { na++; nn += h.npath; }
        }
    *n_nodes = nn;
    return na;
}
static wq_alignment emu_mk(Emu* e, const WqHit& h, wq_align_node* nodes, uint64_t& nn) {
    wq_alignment a;
    a.offset = h.offset; a.path_start = (uint32_t)nn; a.offsetread = h.offsetread; a.strand = h.strand;
    a.isvalid = h.flags & 1; a.npath = h.npath;
    for (int i = 0; i < h.npath; i++) nodes[nn++] = e->pool[h.path_start + i];
    return a;
}
void emu_result_copy_pe(Emu* e, uint32_t* hit_start, uint8_t* nhits, uint8_t* flipped, wq_alignment* aln, wq_alignment* alm, wq_align_node* nodes) {
    uint64_t na = 0, nn = 0;
    for (size_t r = 0; r < e->seeds.size(); r++) {
        const WqSeed& s = e->seeds[r];
        hit_start[r] = (uint32_t)na;
        uint8_t c = 0;
        for (int k = 0; k < s.cnt; k++) {
            const WqHit& hf = e->hits[2 * (s.hit_base + k)];
            const WqHit& hr = e->hits[2 * (s.hit_base + k) + 1];
            if (!(hf.flags & 4)) continue;
            aln[na] = emu_mk(e, hf, nodes, nn);
            alm[na] = emu_mk(e, hr, nodes, nn);
            na++; c++;
        }
        nhits[r] = c;
        flipped[r] = (uint8_t)((!s.strand ? 1 : 0) | ((s.sp & 0x100u) ? 2 : 0));
    }
}
void emu_result_copy(Emu* e, uint32_t* hit_start, uint8_t* nhits, uint8_t* flipped, wq_alignment* aln, wq_align_node* nodes) {
    uint64_t na = 0, nn = 0;
    for (size_t r = 0; r < e->seeds.size(); r++) {
        const WqSeed& s = e->seeds[r];
        hit_start[r] = (uint32_t)na;
        uint8_t c = 0;
        for (int k = 0; k < s.cnt; k++) {
            const WqHit& h = e->hits[s.hit_base + k];
            if (!(h.flags & 4)) continue;
            wq_alignment a;
            a.offset = h.offset; a.path_start = (uint32_t)nn; a.offsetread = h.offsetread; a.strand = h.strand;
            a.isvalid = h.flags & 1; a.npath = h.npath;
            aln[na++] = a;
            for (int i = 0; i < h.npath; i++) nodes[nn++] = e->pool[h.path_start + i];
            c++;
        }
        nhits[r] = c;
        flipped[r] = (s.cnt > 0 && !s.strand) ? 1 : 0;
    }
}
void emu_seeds(Emu* e, int64_t* out) {
    for (size_t r = 0; r < e->seeds.size(); r++) {
        const WqSeed& s = e->seeds[r];
        out[4 * r] = s.sp; out[4 * r + 1] = (int64_t)s.sp + s.cnt - 1; out[4 * r + 2] = s.readloc; out[4 * r + 3] = s.strand;
    }
}
const uint64_t* emu_node_counts(Emu* e) { return e->node_count.data(); }
uint64_t emu_edges(Emu* e, uint64_t* keys, uint64_t* vals) {
    uint64_t n = 0;
    for (size_t i = 0; i < e->ekey.size(); i++)
        if (e->ekey[i] != WQ_EMPTY_KEY) { if (keys) { keys[n] = e->ekey[i]; vals[n] = e->ecount[i]; } n++; }
    return n;
}
// keyed entries as [nw, words..., count_lo, count_hi] records back to back; returns words written/needed
uint64_t emu_keyed(Emu* e, uint32_t* out) {
    uint64_t w = 0;
    for (size_t i = 0; i < e->khash.size(); i++) {
        if (!e->khash[i]) continue;
        uint32_t off = e->koff[i], nw = e->kpool[off];
        if (out) {
            out[w] = nw;
            memcpy(out + w + 1, &e->kpool[off + 1], nw * 4);
            out[w + 1 + nw] = (uint32_t)e->kcount[i];
            out[w + 2 + nw] = (uint32_t)(e->kcount[i] >> 32);
        }
        w += nw + 3;
    }
    return w;
}
// ---- --biascorrect: the per-read tallies (same bodies as wq_k_bias / wq_k_bias_pe), then sort + run lengths + the adjust arithmetic
// of wq_k_bias_model / wq_k_bias_adjust on the host
void emu_set_bias(Emu* e, int on) { e->bias = on != 0; e->bias_tab.assign(2 * WQ_BIAS_NHEX + 2, 0); e->bias_rec.clear(); }
uint64_t emu_bias_finish(Emu* e) {
    const uint64_t nrec = e->bias_tab[2 * WQ_BIAS_NHEX];
    std::vector<uint64_t> rec(e->bias_rec.begin(), e->bias_rec.begin() + nrec), ukey;
    std::vector<uint32_t> ucnt;
    std::sort(rec.begin(), rec.end());
    for (uint64_t i = 0; i < nrec; i++) { if (ukey.empty() || ukey.back() != rec[i]) { ukey.push_back(rec[i]); ucnt.push_back(1); } else ucnt.back()++; }
    const uint64_t* fore = e->bias_tab.data(); const uint64_t* back = fore + WQ_BIAS_NHEX;
    uint64_t fs = 0, bs = 0;
    for (int k = 0; k < WQ_BIAS_NHEX; k++) { fs += fore[k]; bs += back[k]; }
    std::vector<double> ratio(WQ_BIAS_NHEX), gcval(32);
    e->fore_n.assign(WQ_BIAS_NHEX, 0.0); e->back_n.assign(WQ_BIAS_NHEX, 0.0);
    for (int k = 0; k < WQ_BIAS_NHEX; k++) wq_bias_model_entry(fore, back, (double)fs, (double)bs, k, ratio.data(), gcval.data(), e->fore_n.data(), e->back_n.data());
    e->node_p.assign(e->H.N, 0.0); e->node_g.assign(e->H.N, -1.0);
    e->edge_p.assign(e->ekey.size(), 0.0); e->edge_g.assign(e->ekey.size(), -1.0);
    e->keyed_p.assign(e->khash.size(), 0.0); e->keyed_g.assign(e->khash.size(), -1.0);
    const uint32_t n = (uint32_t)ukey.size();
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t id = ukey[i] >> 13;
        if (i > 0 && (ukey[i - 1] >> 13) == id) continue;
        double c, g;
        wq_bias_adjust_run(ukey.data(), ucnt.data(), n, i, ratio.data(), gcval.data(), c, g);
        const uint32_t kind = (uint32_t)(id >> 32), slot = (uint32_t)id;
        (kind == 0 ? e->node_p : kind == 1 ? e->edge_p : e->keyed_p)[slot] = c;
        (kind == 0 ? e->node_g : kind == 1 ? e->edge_g : e->keyed_g)[slot] = g;
    }
    return e->bias_tab[2 * WQ_BIAS_NHEX + 1];      // mapped reads that were too short
}
void emu_bias_nodes(Emu* e, double* p, double* g) { memcpy(p, e->node_p.data(), e->node_p.size() * 8); memcpy(g, e->node_g.data(), e->node_g.size() * 8); }
void emu_bias_edges(Emu* e, double* p, double* g) {       // same order as emu_edges
    uint64_t n = 0;
    for (size_t i = 0; i < e->ekey.size(); i++) if (e->ekey[i] != WQ_EMPTY_KEY) { p[n] = e->edge_p[i]; g[n] = e->edge_g[i]; n++; }
}
void emu_bias_keyed(Emu* e, double* p, double* g) {       // same order as emu_keyed
    uint64_t n = 0;
    for (size_t i = 0; i < e->khash.size(); i++) if (e->khash[i]) { p[n] = e->keyed_p[i]; g[n] = e->keyed_g[i]; n++; }
}
void emu_bias_model(Emu* e, double* fore, double* back) { memcpy(fore, e->fore_n.data(), WQ_BIAS_NHEX * 8); memcpy(back, e->back_n.data(), WQ_BIAS_NHEX * 8); }
const WqCounters* emu_counters(Emu* e) { return &e->counters; }
uint64_t emu_deferred(Emu* e) { return e->deferred; }
const uint64_t* emu_ext_stats() { return g_wq_emu_stats; }
}
