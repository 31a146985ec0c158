/*
 * whippet_oracle.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Single-threaded CPU restatement of the read -> splice-graph alignment path of
 * timbitz/Whippet.jl v1.6.2, used as the parity checker for the CUDA library and as the
 * timed CPU baseline.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product library never does.
 *
 * Each function follows the cited Julia lines statement by statement (1-based indices are
 * kept, through the at1() helpers).  The FM-index arithmetic lives in the un-vendored
 * packages FMIndexes 1.0.0 / WaveletMatrices 0.2.0 (Manifest.toml:126-132,400-406); it is
 * restated from its published semantics and pinned byte-for-byte by test/test_index.jls
 * (see tests/test_oracle_fixture.py).  Alignment is pinned by the ten reads of
 * test/runtests.jl:290-330.  EM numerics are "parity unpinned" (no golden TPM/PSI values
 * exist in the reference); see DESIGN.md.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <map>
#include <string>
#include <vector>

#include "../include/whippet_b200.h"

namespace {

typedef int64_t Int;

// instrumentation for the byte model of SURVEY.md section 8d (not part of the algorithm)
struct Counters { uint64_t fm_steps = 0, hits = 0, nodes_touched = 0, bases = 0, out_nodes = 0, out_alns = 0, count_updates = 0;
                  uint64_t steps_hist[34] = {0}; };      // steps_hist[s] = backward searches that executed s steps (byte model of a k-mer table start)
thread_local Counters g_ctr;


// EdgeType codes, src/graph.jl:39-46
enum : uint8_t { ET_SL = 0, ET_SR = 1, ET_LS = 2, ET_RS = 3, ET_LR = 4, ET_LL = 5, ET_RR = 6 };

struct Score {            // SGAlignScore, src/align.jl:48-52
    uint8_t matches = 0;
    uint8_t mismatches = 0;
    float mistolerance = 0.f;
};
struct ANode {            // SGAlignNode, src/align.jl:58-62
    uint32_t gene, node;
    Score score;
};
struct Align {            // SGAlignment, src/align.jl:76-82
    uint32_t offset = 0;
    uint8_t offsetread = 0;
    std::vector<ANode> path;
    bool strand = true;
    bool isvalid = false;
};

struct Read {             // FASTQRecord after fill!, src/record.jl:2-19
    std::vector<uint8_t> seq;    // 2-bit codes
    std::vector<uint8_t> qual;   // phred
};

struct Lib {
    const wq_index_desc* d;
    std::vector<uint32_t> occ;   // occurrence checkpoints every 64 symbols, 4 per block
    Int n;
    Int G;
    // --- 1-based views ---
    Int gene_offset(Int g) const { return d->gene_offset[g - 1]; }
    Int nnodes(Int g) const { return d->node_base[g] - d->node_base[g - 1]; }
    Int seqlen(Int g) const { return (g < G ? (Int)d->gene_offset[g] : n) - (Int)d->gene_offset[g - 1]; }
    Int nodeoffset(Int g, Int i) const { return d->nodeoffset[d->node_base[g - 1] + i - 1]; }
    Int nodelen(Int g, Int i) const { return d->nodelen[d->node_base[g - 1] + i - 1]; }
    size_t eidx(Int g, Int j) const { return (size_t)d->node_base[g - 1] + (g - 1) + (j - 1); }
    uint8_t edgetype(Int g, Int j) const { return d->edgetype[eidx(g, j)]; }
    uint64_t edgeleft(Int g, Int j) const { return d->edgeleft[eidx(g, j)]; }
    uint64_t edgeright(Int g, Int j) const { return d->edgeright[eidx(g, j)]; }
    uint8_t seq(Int g, Int i) const { return d->seq4[(size_t)d->gene_offset[g - 1] + i - 1]; }
};

void build_occ(Lib& L) {
    Int nb = L.n / 64 + 1;
    L.occ.assign((size_t)nb * 4, 0);
    uint32_t c[4] = {0, 0, 0, 0};
    for (Int i = 0; i < L.n; i++) {
        if (i % 64 == 0) memcpy(&L.occ[(size_t)(i / 64) * 4], c, sizeof c);
        c[L.d->bwt[i] & 3]++;
    }
    if (L.n % 64 == 0) memcpy(&L.occ[(size_t)(L.n / 64) * 4], c, sizeof c);
}

// WaveletMatrices.rank(char, bwt, i): occurrences of `char` in bwt[1..i]; call site src/fmindex_patch.jl:26-27
Int rank(const Lib& L, uint8_t ch, Int i) {
    if (i <= 0) return 0;
    if (i > L.n) i = L.n;
    Int b = i / 64;
    Int r = L.occ[(size_t)b * 4 + ch];
    for (Int k = b * 64; k < i; k++) r += (L.d->bwt[k] == ch);
    return r;
}

// FMIndexes.sa_range override, src/fmindex_patch.jl:19-31 (initial range 1:(n+1))
void sa_range(const Lib& L, const uint8_t* q, Int len, Int& sp, Int& ep) {
    sp = 1;
    ep = L.n + 1;
    Int sentinel = (Int)L.d->sentinel;
    Int i = len;
    Int steps = 0;
    while (sp <= ep && i >= 1) {
        uint8_t ch = q[i - 1];
        g_ctr.fm_steps++;
        steps++;
        Int c = L.d->count[ch];
        sp = c + rank(L, ch, (sp > sentinel ? sp - 1 : sp) - 1) + 1;
        ep = c + rank(L, ch, (ep > sentinel ? ep - 1 : ep));
        i -= 1;
    }
    g_ctr.steps_hist[steps > 33 ? 33 : steps]++;
}

// FMIndexes.sa_value at r=1 (call sites src/align.jl:240, src/paired.jl:9): row 1 is `$`
Int sa_value(const Lib& L, Int i) { return i == 1 ? L.n : (Int)L.d->sa[i - 2]; }

template <class T>
Int searchsortedlast_u32(const T* a, Int n, Int x) {   // last 1-based index with a[i] <= x, 0 if none
    Int lo = 0, hi = n + 1;
    while (lo < hi - 1) {
        Int m = (lo + hi) >> 1;
        if ((Int)a[m - 1] <= x) lo = m; else hi = m;
    }
    return lo;
}

// src/align.jl:90-114
uint8_t matches(const std::vector<ANode>& v, size_t from = 0, size_t to = (size_t)-1) {
    uint8_t m = 0;
    if (to > v.size()) to = v.size();
    for (size_t i = from; i < to; i++) m = (uint8_t)(m + v[i].score.matches);
    return m;
}
float mistolerance(const std::vector<ANode>& v) {
    float m = 0.f;
    for (auto& x : v) m += x.score.mistolerance;
    return m;
}
float score(const std::vector<ANode>& v) {
    float s = 0.f;
    for (auto& x : v) s += (float)x.score.matches - x.score.mistolerance;
    return s;
}
float score(const Align& a) { return score(a.path); }
float identity(const Align& a, Int readlen) { return score(a) / (float)readlen; }      // src/align.jl:114
bool isvalid(const Align& a) {                                                         // src/align.jl:115
    return a.isvalid && a.path.size() >= 1 && (float)matches(a.path) > mistolerance(a.path);
}

float PHRED_PROB[256];
// phred_to_prob, src/align.jl:124: Float32(1 - 10^(-Int8(phred)/10))
void init_phred() {
    for (int i = 0; i < 256; i++) PHRED_PROB[i] = (float)(1.0 - std::pow(10.0, -(double)(int8_t)i / 10.0));
}

// kmer_index(read.sequence[a:b]) for a 2-bit read, src/sgkmer.jl:12-26 (never 0)
Int read_kmer_index(const Read& r, Int a, Int k) {
    uint64_t x = 0;
    for (Int i = a; i < a + k; i++) x = (x << 2) | r.seq[i - 1];
    return (Int)x + 1;
}

struct Ctx {
    const Lib& L;
    const wq_align_param& p;
    const Read& read;
    Int readlen;
};

const uint32_t* table_slot(const uint32_t* ptr, const uint32_t* nodes, Int kmer_ind, Int& cnt) {
    // isassigned(lib.edges.X, ind): an #undef slot is an empty CSR range
    uint32_t a = ptr[kmer_ind - 1], b = ptr[kmer_ind];
    cnt = (Int)b - (Int)a;
    return nodes + 2 * (size_t)a;
}

// intersect_sorted with the gene-only ordering, src/edges.jl:15-18,40-55: returns arrB entries
void intersect_sorted(const uint32_t* A, Int na, const uint32_t* B, Int nb, std::vector<std::pair<uint32_t, uint32_t>>& res) {
    Int i = 1, j = 1;
    while (i <= na && j <= nb) {
        uint32_t ga = A[2 * (i - 1)], gb = B[2 * (j - 1)];
        if (ga > gb) j += 1;
        else if (ga < gb) i += 1;
        else {
            res.emplace_back(B[2 * (j - 1)], B[2 * (j - 1) + 1]);
            i += 1;
            j += 1;
        }
    }
}

Align ungapped_fwd_extend(const Ctx& c, uint32_t geneind, Int sgidx, Int ridx, Align align, uint32_t nodeidx);
Align ungapped_rev_extend(const Ctx& c, uint32_t geneind, Int sgidx, Int ridx, Align align, uint32_t nodeidx);

// closure spliced_fwd_extend, src/align.jl:287-311
Align spliced_fwd_extend(const Ctx& c, uint32_t geneind, uint32_t edgeind, Int ridx, Int right_kmer_ind, const Align& align) {
    const Lib& L = c.L;
    Int nr, nl;
    const uint32_t* R = table_slot(L.d->right_ptr, L.d->right_nodes, right_kmer_ind, nr);
    if (nr == 0) return align;
    Int left_kmer_ind = (Int)L.edgeleft(geneind, edgeind) + 1;
    const uint32_t* Lt = table_slot(L.d->left_ptr, L.d->left_nodes, left_kmer_ind, nl);
    if (nl == 0) return align;
    std::vector<std::pair<uint32_t, uint32_t>> right_nodes;
    intersect_sorted(Lt, nl, R, nr, right_nodes);

    Align best = align;
    for (auto& rn : right_nodes) {
        if (rn.first != align.path[0].gene) continue;
        Int rn_offset = L.nodeoffset(rn.first, rn.second);
        Align res_align = ungapped_fwd_extend(c, rn.first, rn_offset, ridx, align /*deepcopy*/, rn.second);
        if (score(res_align) > score(best)) best = res_align;
    }
    if (score(best) >= score(align) + (float)c.p.kmer_size) best.isvalid = true;
    return best;
}

// ungapped_fwd_extend, src/align.jl:275-418
Align ungapped_fwd_extend(const Ctx& c, uint32_t geneind, Int sgidx, Int ridx, Align align, uint32_t nodeidx) {
    const Lib& L = c.L;
    const wq_align_param& p = c.p;
    const Int readlen = c.readlen;
    const Int seqlen = L.seqlen(geneind);
    const Int nn = L.nnodes(geneind);

    std::vector<uint8_t> passed_edges;
    bool has_passed = false;
    float mistol = mistolerance(align.path);

    align.path.push_back(ANode{geneind, nodeidx, Score()});
    g_ctr.nodes_touched++;

    while ((double)mistol <= (double)p.mismatches && ridx <= readlen && sgidx <= seqlen) {
        if ((Int)nodeidx <= nn && sgidx == L.nodeoffset(geneind, nodeidx) + L.nodelen(geneind, nodeidx)) {   // hit next edge
            uint32_t curedge = nodeidx + 1;
            uint8_t et = L.edgetype(geneind, curedge);
            if (et == ET_LR && L.nodelen(geneind, curedge) >= p.kmer_size && readlen - ridx + 1 >= p.kmer_size) {
                Int rkmer_ind = read_kmer_index(c.read, ridx, p.kmer_size);
                if (rkmer_ind == 0) break;
                if ((Int)L.edgeright(geneind, curedge) + 1 == rkmer_ind) {
                    ridx += p.kmer_size - 1;
                    sgidx += p.kmer_size - 1;
                    nodeidx += 1;
                    Score s;
                    s.matches = (uint8_t)(p.kmer_size - 1);
                    align.path.push_back(ANode{geneind, nodeidx, s});
                    align.isvalid = true;
                } else {
                    align = spliced_fwd_extend(c, geneind, curedge, ridx, rkmer_ind, align);
                    break;
                }
            } else if (et == ET_LR || et == ET_LL) {
                has_passed = true;
                passed_edges.push_back((uint8_t)align.path.size());
                nodeidx += 1;
                align.path.push_back(ANode{geneind, nodeidx, Score()});
            } else if (et == ET_LS) {
                if (readlen - ridx + 1 >= p.kmer_size) {
                    Int rkmer_ind = read_kmer_index(c.read, ridx, p.kmer_size);
                    if (rkmer_ind == 0) break;
                    align = spliced_fwd_extend(c, geneind, curedge, ridx, rkmer_ind, align);
                }
                break;
            } else if (et == ET_SR) {
                break;
            } else {   // 'RR' || 'SL' || 'RS'
                nodeidx += 1;
                if ((Int)nodeidx <= nn) align.path.push_back(ANode{geneind, nodeidx, Score()});
            }
        }
        Score& sc = align.path.back().score;
        g_ctr.bases++;
        if ((uint8_t)(1u << c.read.seq[ridx - 1]) == L.seq(geneind, sgidx)) {
            sc.matches += 1;
        } else {
            float prob = PHRED_PROB[c.read.qual[ridx - 1]];
            sc.mistolerance += prob;
            mistol += prob;
            sc.mismatches += 1;
        }
        ridx += 1;
        sgidx += 1;
    }

    // passed-edge backtracking, src/align.jl:382-405
    if (has_passed) {
        for (Int ci = (Int)passed_edges.size(); ci >= 1; ci--) {
            Int cval = passed_edges[ci - 1];
            if ((Int)matches(align.path, (size_t)cval) >= p.kmer_size + p.mismatches) {
                align.isvalid = true;
                break;
            } else {
                uint32_t cur_edge = align.path[cval - 1].node + 1;
                align.path.resize((size_t)cval);   // splice!( path, cval+1:end )
                Int cur_ridx = ridx - (sgidx - L.nodeoffset(geneind, cur_edge));
                if (!((cur_ridx + p.kmer_size - 1) <= readlen)) continue;
                Int rkmer_ind = read_kmer_index(c.read, cur_ridx, p.kmer_size);
                if (rkmer_ind == 0) continue;
                Align res_align = spliced_fwd_extend(c, geneind, cur_edge, cur_ridx, rkmer_ind, align);
                if (res_align.path.size() > align.path.size()) {
                    if (score(res_align) > score(align)) align = res_align;
                }
            }
        }
    }

    // clean up bad trailing nodes, src/align.jl:408-412
    while (align.path.size() >= 1 &&
           (align.path.back().score.mismatches >= align.path.back().score.matches ||
            (Int)align.path.back().score.matches - (Int)align.path.back().score.mismatches <= 1)) {
        align.path.pop_back();
        if (align.path.size() == 0) align.isvalid = false;
    }

    if ((double)identity(align, readlen) >= p.score_min) align.isvalid = true;
    return align;
}

// closure spliced_rev_extend, src/align.jl:437-461
Align spliced_rev_extend(const Ctx& c, uint32_t geneind, uint32_t edgeind, Int ridx, Int left_kmer_index, const Align& align) {
    const Lib& L = c.L;
    Int nr, nl;
    const uint32_t* Lt = table_slot(L.d->left_ptr, L.d->left_nodes, left_kmer_index, nl);
    if (nl == 0) return align;
    Int right_kmer_ind = (Int)L.edgeright(geneind, edgeind) + 1;
    const uint32_t* R = table_slot(L.d->right_ptr, L.d->right_nodes, right_kmer_ind, nr);
    if (nr == 0) return align;
    std::vector<std::pair<uint32_t, uint32_t>> left_nodes;
    intersect_sorted(R, nr, Lt, nl, left_nodes);

    Align best = align;
    for (auto& rn : left_nodes) {
        if (rn.first != align.path[0].gene) continue;
        Int rn_offset = L.nodeoffset(rn.first, rn.second - 1) + L.nodelen(rn.first, rn.second - 1) - 1;
        Align res_align = ungapped_rev_extend(c, rn.first, rn_offset, ridx, align /*deepcopy*/, rn.second - 1);
        if (score(res_align) > score(best)) best = res_align;
    }
    if (score(best) >= score(align) + (float)c.p.kmer_size) best.isvalid = true;
    return best;
}

// ungapped_rev_extend, src/align.jl:423-575
Align ungapped_rev_extend(const Ctx& c, uint32_t geneind, Int sgidx, Int ridx, Align align, uint32_t nodeidx) {
    const Lib& L = c.L;
    const wq_align_param& p = c.p;
    const Int readlen = c.readlen;

    std::vector<uint8_t> passed_edges;
    bool has_passed = false;
    float mistol = mistolerance(align.path);

    if (align.path.size() <= 0 || align.path[0].gene != geneind || align.path[0].node != nodeidx)
        align.path.insert(align.path.begin(), ANode{geneind, nodeidx, Score()});

    while ((double)mistol <= (double)p.mismatches && ridx > 0 && sgidx > 0) {
        if (sgidx == L.nodeoffset(geneind, nodeidx) - 1) {   // hit right edge
            uint32_t leftnode = nodeidx - 1;
            uint8_t et = L.edgetype(geneind, nodeidx);
            if (et == ET_LR && L.nodelen(geneind, leftnode) >= p.kmer_size && ridx >= p.kmer_size) {
                Int lkmer_ind = read_kmer_index(c.read, ridx - p.kmer_size + 1, p.kmer_size);
                if (lkmer_ind == 0) break;
                if ((Int)L.edgeleft(geneind, nodeidx) + 1 == lkmer_ind) {
                    ridx -= p.kmer_size - 1;
                    sgidx -= p.kmer_size - 1;
                    nodeidx -= 1;
                    Score s;
                    s.matches = (uint8_t)(p.kmer_size - 1);
                    align.path.insert(align.path.begin(), ANode{geneind, nodeidx, s});
                    align.isvalid = true;
                } else {
                    align = spliced_rev_extend(c, geneind, nodeidx, ridx, lkmer_ind, align);
                    break;
                }
            } else if (et == ET_LR || et == ET_RR) {
                has_passed = true;
                passed_edges.push_back((uint8_t)(align.path.size() - 1));
                nodeidx -= 1;
                align.path.insert(align.path.begin(), ANode{geneind, nodeidx, Score()});
            } else if (et == ET_SR) {
                if (ridx >= p.kmer_size) {
                    Int lkmer_ind = read_kmer_index(c.read, ridx - p.kmer_size + 1, p.kmer_size);
                    if (lkmer_ind == 0) break;
                    align = spliced_rev_extend(c, geneind, nodeidx, ridx, lkmer_ind, align);
                }
                break;
            } else if (et == ET_LS) {
                break;
            } else {   // 'LL' || 'RS' || 'SL'
                nodeidx -= 1;
                if (nodeidx > 0) align.path.insert(align.path.begin(), ANode{geneind, nodeidx, Score()});
            }
        }
        Score& sc = align.path.front().score;
        g_ctr.bases++;
        if ((uint8_t)(1u << c.read.seq[ridx - 1]) == L.seq(geneind, sgidx)) {
            sc.matches += 1;
        } else {
            float prob = PHRED_PROB[c.read.qual[ridx - 1]];
            sc.mistolerance += prob;
            mistol += prob;
            sc.mismatches += 1;
        }
        ridx -= 1;
        sgidx -= 1;
    }

    Int cur_ridx = ridx + 1;
    Int cur_sgidx = sgidx + 1;
    // passed-edge backtracking, src/align.jl:526-549
    if (has_passed) {
        for (Int ci = (Int)passed_edges.size(); ci >= 1; ci--) {
            Int cval = passed_edges[ci - 1];
            Int len = (Int)align.path.size();
            Int upto = len - (cval + 1);   // path[1:(end-(cval+1))]
            if ((Int)matches(align.path, 0, (size_t)std::max<Int>(upto, 0)) >= p.kmer_size + p.mismatches) {
                align.isvalid = true;
                break;
            } else {
                uint32_t cur_edge = align.path[len - cval - 1].node;       // path[end-cval]
                if (upto > 0) align.path.erase(align.path.begin(), align.path.begin() + upto);
                cur_ridx = ridx + (L.nodeoffset(geneind, cur_edge) - sgidx);
                cur_sgidx = L.nodeoffset(geneind, cur_edge);
                if (!((cur_ridx - p.kmer_size) > 0)) continue;
                Int lkmer_ind = read_kmer_index(c.read, cur_ridx - p.kmer_size, p.kmer_size);
                if (lkmer_ind == 0) continue;
                Align res_align = spliced_rev_extend(c, geneind, cur_edge, cur_ridx - 1, lkmer_ind, align);
                if (res_align.path.size() > align.path.size()) {
                    if (score(res_align) > score(align)) align = res_align;
                }
            }
        }
    }

    if ((double)identity(align, readlen) >= p.score_min) align.isvalid = true;
    if (sgidx < (Int)align.offset) align.offset = (uint32_t)cur_sgidx;
    if (cur_ridx < (Int)align.offsetread) align.offsetread = (uint8_t)cur_ridx;

    // clean up bad leading nodes, src/align.jl:562-572
    while (align.path.size() >= 1 &&
           (align.path.front().score.mismatches >= align.path.front().score.matches ||
            (Int)align.path.front().score.matches - (Int)align.path.front().score.mismatches <= 1)) {
        uint8_t offset_change = (uint8_t)(align.path.front().score.matches + align.path.front().score.mismatches);
        align.path.erase(align.path.begin());
        if (align.path.size() >= 1) {
            align.offset = (uint32_t)L.nodeoffset(geneind, align.path.front().node);
            align.offsetread = (uint8_t)(align.offsetread + offset_change);
        } else {
            align.isvalid = false;
        }
    }
    return align;
}

// _ungapped_align, src/align.jl:209-223
Align ungapped_align_at(const Ctx& c, Int indx, Int readloc, bool ispos, Int geneind = -1) {
    const Lib& L = c.L;
    if (geneind < 0) geneind = searchsortedlast_u32(L.d->gene_offset, L.G, indx);
    Int sgidx = indx - L.gene_offset(geneind);
    uint32_t offset_node = (uint32_t)searchsortedlast_u32(L.d->nodeoffset + L.d->node_base[geneind - 1], L.nnodes(geneind), sgidx + 1);
    Align a;
    a.offset = (uint32_t)(sgidx + 1);
    a.offsetread = (uint8_t)(readloc + 1);
    a.strand = ispos;
    a.isvalid = false;
    Align align = ungapped_fwd_extend(c, (uint32_t)geneind, sgidx + 1, readloc + 1, a, offset_node);
    align = ungapped_rev_extend(c, (uint32_t)geneind, sgidx, readloc, align, offset_node);
    return align;
}

struct Seed {
    Int sp, ep;    // sa range (sp > ep: empty)
    Int readloc;
    bool strand;
};

// seed_locate, src/align.jl:145-185
Seed seed_locate(const Lib& L, const wq_align_param& p, const Read& read, bool offset_left, bool both_strands) {
    Int readlen = (Int)read.seq.size();
    Int ctry = 1;
    Int increment, increment_sm, curpos;
    if (offset_left) {
        increment = p.seed_inc;
        increment_sm = 1;
        curpos = p.seed_buffer;
    } else {
        increment = -p.seed_inc;
        increment_sm = -1;
        curpos = readlen - p.seed_length - p.seed_buffer;
    }
    Int maxright = readlen - p.seed_length - p.seed_buffer;
    std::vector<uint8_t> curseq((size_t)p.seed_length);
    while (ctry <= p.seed_try && p.seed_buffer <= curpos && curpos <= maxright) {
        Int minq = 256;
        for (Int i = curpos; i <= curpos + p.seed_length - 1; i++) minq = std::min<Int>(minq, read.qual[i - 1]);
        if (minq <= p.seed_min_qual) {
            curpos += increment_sm;
            continue;
        }
        for (Int i = 0; i < p.seed_length; i++) curseq[i] = read.seq[curpos - 1 + i];
        Int sp, ep;
        sa_range(L, curseq.data(), p.seed_length, sp, ep);
        Int cnt = ep >= sp ? ep - sp + 1 : 0;
        ctry += 1;
        if (cnt == 0 || cnt > p.seed_tolerate) {
            if (both_strands) {
                std::reverse(curseq.begin(), curseq.end());
                for (auto& b : curseq) b = 3 - b;
                Int rsp, rep;
                sa_range(L, curseq.data(), p.seed_length, rsp, rep);
                Int rev_cnt = rep >= rsp ? rep - rsp + 1 : 0;
                if (0 < rev_cnt && rev_cnt <= p.seed_tolerate) {
                    Int rev_pos = readlen - (curpos + p.seed_length - 1) + 1;
                    return Seed{rsp, rep, rev_pos, false};
                }
            }
        } else {
            return Seed{sp, ep, curpos, true};
        }
        curpos += increment;
    }
    return Seed{2, 1, curpos, true};
}

// reverse_complement!(rec), src/record.jl:23-26
void reverse_complement(Read& r) {
    std::reverse(r.seq.begin(), r.seq.end());
    for (auto& b : r.seq) b = 3 - b;
    std::reverse(r.qual.begin(), r.qual.end());
}

// splice_by_identity!, src/align.jl:198-207
void splice_by_identity(std::vector<Align>& arr, float threshold, double buffer, Int readlen) {
    size_t i = 0;
    while (i < arr.size()) {
        if ((double)(threshold - identity(arr[i], readlen)) > buffer) arr.erase(arr.begin() + i);
        else i += 1;
    }
}

// ungapped_align (single end), src/align.jl:225-270.  Returns false for a null result.
bool ungapped_align_se(const Lib& L, const wq_align_param& p, Read& read, std::vector<Align>& res, bool& flipped) {
    Seed seed = seed_locate(L, p, read, true, !p.is_stranded);
    Int readlen = (Int)read.seq.size();
    bool ispos = true;
    flipped = false;
    if (!seed.strand) {
        reverse_complement(read);
        ispos = false;
        flipped = true;
    }
    bool isnull = true;
    float maxscore = 0.f;
    res.clear();
    Ctx c{L, p, read, readlen};
    for (Int row = seed.sp; row <= seed.ep; row++) {
        Int s = sa_value(L, row) + 1;
        g_ctr.hits++;
        Align align = ungapped_align_at(c, s, seed.readloc, ispos);
        float scvar = identity(align, readlen);
        if (isvalid(align)) {
            if (isnull) {
                isnull = false;
                res.push_back(align);
                maxscore = scvar;
            } else {
                if (scvar > maxscore) {
                    if ((double)(scvar - maxscore) > p.score_range) {
                        res.clear();
                    } else {
                        splice_by_identity(res, scvar, p.score_range, readlen);
                    }
                    maxscore = scvar;
                    res.push_back(align);
                } else {
                    if ((double)(maxscore - scvar) <= p.score_range) res.push_back(align);
                }
            }
        }
    }
    return !isnull;
}

// sorted_offsets, src/paired.jl:5-14
std::vector<Int> sorted_offsets(const Lib& L, const Seed& s) {
    std::vector<Int> res;
    for (Int row = s.sp; row <= s.ep; row++) res.push_back(sa_value(L, row) + 1);
    std::sort(res.begin(), res.end());
    return res;
}

// paired identity, src/paired.jl:2-3 (Float32 + Float32, divided by an Int sum)
float identity_pair(const Align& one, const Align& two, Int onelen, Int twolen) {
    return (score(one) + score(two)) / (float)(onelen + twolen);
}

// paired ungapped_align, src/paired.jl:42-125
bool ungapped_align_pe(const Lib& L, const wq_align_param& p, Read& fwd, Read& rev,
                       std::vector<Align>& fwd_res, std::vector<Align>& rev_res, uint8_t& flipped) {
    Seed fs = seed_locate(L, p, fwd, true, !p.is_stranded);
    bool strand = fs.strand;
    Seed rs;
    flipped = 0;
    if ((p.is_pair_rc && strand) || (!p.is_pair_rc && !strand)) {
        reverse_complement(rev);
        flipped |= 2;
        rs = seed_locate(L, p, rev, false, false);
    } else {
        rs = seed_locate(L, p, rev, true, false);
    }
    Int fwd_len = (Int)fwd.seq.size();
    Int rev_len = (Int)rev.seq.size();
    bool ispos = true;
    if (!strand) {
        reverse_complement(fwd);
        flipped |= 1;
        ispos = false;
    }
    bool isnull = true;
    float maxscore = 0.f;
    fwd_res.clear();
    rev_res.clear();
    std::vector<Int> fwd_sorted = sorted_offsets(L, fs);
    std::vector<Int> rev_sorted = sorted_offsets(L, rs);
    Ctx cf{L, p, fwd, fwd_len};
    Ctx cr{L, p, rev, rev_len};
    size_t fidx = 1, ridx = 1;
    while (fidx <= fwd_sorted.size() && ridx <= rev_sorted.size()) {
        Int dist = std::llabs(fwd_sorted[fidx - 1] - rev_sorted[ridx - 1]);
        if (dist < p.seed_pair_range) {
            Int geneind = searchsortedlast_u32(L.d->gene_offset, L.G, fwd_sorted[fidx - 1]);
            Int next_offset = geneind < L.G ? L.gene_offset(geneind + 1) : L.gene_offset(geneind) + L.seqlen(geneind);
            if (L.gene_offset(geneind) <= rev_sorted[ridx - 1] && rev_sorted[ridx - 1] < next_offset) {
                Align fwd_aln = ungapped_align_at(cf, fwd_sorted[fidx - 1], fs.readloc, ispos, geneind);
                Align rev_aln = ungapped_align_at(cr, rev_sorted[ridx - 1], rs.readloc, !ispos, geneind);
                float scvar = identity_pair(fwd_aln, rev_aln, fwd_len, rev_len);
                if (isvalid(fwd_aln) && isvalid(rev_aln)) {
                    if (isnull) {
                        isnull = false;
                        fwd_res.assign(1, fwd_aln);
                        rev_res.assign(1, rev_aln);
                        maxscore = scvar;
                    } else {
                        if (scvar > maxscore) {
                            if ((double)(scvar - maxscore) > p.score_range) {
                                fwd_res.clear();
                                rev_res.clear();
                            } else {
                                size_t i = 0;   // paired splice_by_identity!, src/paired.jl:29-39
                                while (i < fwd_res.size()) {
                                    if ((double)(scvar - identity_pair(fwd_res[i], rev_res[i], fwd_len, rev_len)) > p.score_range) {
                                        fwd_res.erase(fwd_res.begin() + i);
                                        rev_res.erase(rev_res.begin() + i);
                                    } else i += 1;
                                }
                            }
                            maxscore = scvar;
                            fwd_res.push_back(fwd_aln);
                            rev_res.push_back(rev_aln);
                        } else {
                            if ((double)(maxscore - scvar) <= p.score_range) {
                                fwd_res.push_back(fwd_aln);
                                rev_res.push_back(rev_aln);
                            }
                        }
                    }
                }
                ridx += 1;
                fidx += 1;
                continue;
            }
        }
        if (fwd_sorted[fidx - 1] > rev_sorted[ridx - 1]) ridx += 1;
        else fidx += 1;
    }
    return !isnull;
}

Read load_read(const wq_read_batch* b, uint32_t i) {
    Read r;
    Int len = b->len[i];
    r.seq.resize(len);
    r.qual.resize(len);
    const uint64_t* w = b->seq2 + b->seq_off[i];
    for (Int j = 0; j < len; j++) r.seq[j] = (uint8_t)((w[j >> 5] >> (2 * (j & 31))) & 3);
    memcpy(r.qual.data(), b->qual + b->qual_off[i], (size_t)len);
    return r;
}

struct Result {
    std::vector<uint32_t> hit_start;
    std::vector<uint8_t> nhits, flipped;
    std::vector<wq_alignment> aln, aln_mate;
    std::vector<wq_align_node> nodes;
};

void emit(Result& R, const Align& a, std::vector<wq_alignment>& dst) {
    wq_alignment o;
    o.offset = a.offset;
    o.path_start = (uint32_t)R.nodes.size();
    o.offsetread = a.offsetread;
    o.strand = a.strand;
    o.isvalid = a.isvalid;
    o.npath = (uint8_t)std::min<size_t>(a.path.size(), 255);
    dst.push_back(o);
    for (auto& n : a.path) {
        wq_align_node x;
        x.gene = n.gene;
        x.node = n.node;
        x.mistolerance = n.score.mistolerance;
        x.matches = n.score.matches;
        x.mismatches = n.score.mismatches;
        x._pad = 0;
        R.nodes.push_back(x);
    }
}

// ---------------------------------------------------------------------------------------------
// Counting: SpliceGraphQuant / MultiMapping with DefaultCounter (src/quant.jl:129-144, 439-446,
// src/bias.jl:66-85: every read adds 1.0).  Stores are kept in ordered maps; the sorted key order is
// the declared canonical order (Julia Dict order is not reproducible).
// Keys use the same record format as the library's download (include/whippet_b200.h::wq_keyed).
// --biascorrect (bin/whippet-quant.jl:116-122): JointBiasMod / JointBiasCounter, src/bias.jl:103-270.
// Julia's div for floats is exact floor-toward-zero division, round((x - rem(x, y)) / y) (Base; the reference's own test
// `length(ExpectedGC(seq)) == 20`, test/runtests.jl:101, holds only under it: div(1.0, 0.05) == 19.0).
double jl_div(double x, double y) { return std::nearbyint((x - std::fmod(x, y)) / y); }
struct BiasTemp {               // JointBiasTemp, src/bias.jl:103-106
    uint16_t primer[2];
    int32_t gc[100];
};
struct BiasMod {                // JointBiasMod(), src/bias.jl:108-143 with the defaults of the constructor
    std::vector<double> fore, back;                                    // 4^6
    Int size = 6, backoffset = 23, backnpos = 10, foreoffset = 1, forenpos = 2;
    std::vector<double> gcfore, gcback;                                // Int(ceil(1.0 / 0.05) + 1) = 21
    Int gcsize = 50; double gcwidth = 0.05; Int gcoffset = 5, gcincrement = 10;
    BiasTemp temp;
    BiasMod() : fore(4096, 0.0), back(4096, 0.0), gcfore(21, 0.0), gcback(21, 0.0) { memset(&temp, 0, sizeof temp); }
};
struct BiasCounter {            // JointBiasCounter without `count` (that is the Float64 of the Quant store), src/bias.jl:207-218
    std::map<uint16_t, int32_t> map;       // Dict{UInt16,Int32}: iterated in ascending key order (canonical; Julia Dict order is unspecified)
    int32_t gc[21];
    BiasCounter() { memset(gc, 0, sizeof gc); }
};
// kmer_index_trailing(UInt16, seq[i:i+size-1]), src/sgkmer.jl:17-28 (reads hold no ambiguous symbols after fill_ambiguous)
uint16_t kmer_trailing(const std::vector<uint8_t>& seq, Int i1, Int size) {
    uint16_t x = 0;
    for (Int k = 0; k < size; k++) x = (uint16_t)((x << 2) | seq[(size_t)(i1 - 1 + k)]);
    return x;
}
// primer_count!, src/bias.jl:154-167.  false = the read is too short for the background window (the reference throws BoundsError)
bool primer_count(BiasMod& m, const std::vector<uint8_t>& seq) {
    if ((Int)seq.size() < m.backoffset + m.backnpos - 1 + m.size - 1) return false;
    Int idx = 1;
    for (Int i = m.foreoffset; i <= m.foreoffset + m.forenpos - 1; i++) {
        uint16_t hept = kmer_trailing(seq, i, m.size);
        m.fore[hept] += 1.0;
        m.temp.primer[idx - 1] = hept;
        idx += 1;
    }
    for (Int i = m.backoffset; i <= m.backoffset + m.backnpos - 1; i++) m.back[kmer_trailing(seq, i, m.size)] += 1.0;
    return true;
}
// gc_count!, src/bias.jl:169-181 (idx restarts at 1 on every call: in the paired form the mate's bins overwrite the first entries)
void gc_count(BiasMod& m, const std::vector<uint8_t>& seq) {
    Int idx = 1;
    for (Int i = m.gcoffset; i <= (Int)seq.size() - m.gcsize; i += m.gcincrement) {
        Int ngc = 0;
        for (Int k = 0; k < m.gcsize; k++) { uint8_t c = seq[(size_t)(i - 1 + k)]; ngc += (c == 1 || c == 2); }
        double gc = (double)ngc / (double)m.gcsize;                    // gc_content(seq[i:i+gcsize-1])
        Int bin = (Int)(jl_div(gc, m.gcwidth) + 1);
        m.gcfore[(size_t)bin - 1] += 1.0;
        m.temp.gc[idx - 1] = (int32_t)bin;
        idx += 1;
    }
}
// count!(mod, seq) / count!(mod, fwd, rev), src/bias.jl:183-203
bool bias_count(BiasMod& m, const std::vector<uint8_t>& seq, const std::vector<uint8_t>* rev = nullptr) {
    if (!primer_count(m, seq)) return false;
    memset(m.temp.gc, 0, sizeof m.temp.gc);
    gc_count(m, seq);
    if (rev) gc_count(m, *rev);
    return true;
}
// push!(cnt::JointBiasCounter, temp), src/bias.jl:222-235
void bias_push(BiasCounter& c, const BiasTemp& t) {
    for (int i = 0; i < 2; i++) c.map[t.primer[i]] += 1;
    for (int i = 0; i < 100; i++) {
        if (t.gc[i] > 0) c.gc[t.gc[i] - 1] += 1;
        else return;
    }
}

struct Quant {
    bool bias = false;                                      // CounterType = JointBiasCounter, BiasModelType = JointBiasMod
    BiasMod mod;
    std::vector<BiasCounter> bnode;
    std::map<uint64_t, BiasCounter> bedge;
    std::map<std::vector<uint32_t>, BiasCounter> blongs, bmulti;
    uint64_t bias_short = 0;
    std::vector<double> node;                               // by global node index
    std::map<uint64_t, double> edge;                        // gene<<32 | l<<16 | r  (edge when l<r, circ otherwise)
    std::map<std::vector<uint32_t>, double> longs;          // [npath, gene, nodes...]  /  PE: [npf | npr<<8 | 2<<28, gene, fwd.., rev..]
    std::map<std::vector<uint32_t>, double> multi;          // [1<<28 | nhits, per hit: npath, gene, nodes...]
    uint64_t total = 0, mapped = 0, nmulti = 0, sum_readlen = 0;
    double mean_readlen = 0.0;                              // running mean in read order, src/reads.jl:69
};

// count!(quant, align, value), src/quant.jl:556-594
// With a JointBiasTemp as `value` push! files the read's hexamers and GC bins in the counter and leaves its count alone
// (src/bias.jl:222-235); the stores' Float64 stay 0.0 until primer_adjust!.
void count_se(const Lib& L, Quant& Q, const Align& a, double value, const BiasTemp* bt = nullptr) {
    if (!a.isvalid) return;
    uint32_t init_gene = a.path[0].gene;
    size_t nb = L.d->node_base[init_gene - 1];
    if (bt) value = 0.0;
    if (a.path.size() == 1) {
        Q.node[nb + a.path[0].node - 1] += value;
        if (bt) bias_push(Q.bnode[nb + a.path[0].node - 1], *bt);
    } else {
        for (auto& n : a.path) if (n.gene != init_gene) return;
        if (a.path.size() == 2) {
            uint32_t l = a.path[0].node, r = a.path[1].node;
            uint64_t key = ((uint64_t)init_gene << 32) | ((uint64_t)(l & 0xFFFF) << 16) | (r & 0xFFFF);
            Q.edge[key] += value;     // l < r: edge interval; l >= r: circ tuple (same packed key space)
            if (bt) bias_push(Q.bedge[key], *bt);
        } else {
            std::vector<uint32_t> key{(uint32_t)a.path.size(), init_gene};
            for (auto& n : a.path) key.push_back(n.node);
            Q.longs[key] += value;
            if (bt) bias_push(Q.blongs[key], *bt);
        }
    }
}

// Base.in(first::SGAlignPath, last::SGAlignPath), src/quant.jl:47-62
bool path_in_path(const std::vector<ANode>& first, const std::vector<ANode>& last) {
    for (Int i = 1; i <= (Int)last.size() - (Int)first.size() + 1; i++) {
        Int offset = i - 1;
        bool match = true;
        for (Int j = 1; j <= (Int)first.size(); j++) {
            const ANode& a = first[j - 1];
            const ANode& b = last[offset + j - 1];
            if (!(a.gene == b.gene && a.node == b.node)) { match = false; break; }
        }
        if (match) return true;
    }
    return false;
}

std::vector<uint32_t> paired_key(const Align& f, const Align& r) {
    std::vector<uint32_t> key{(2u << 28) | (uint32_t)f.path.size() | ((uint32_t)r.path.size() << 8), f.path[0].gene};
    for (auto& n : f.path) key.push_back(n.node);
    for (auto& n : r.path) key.push_back(n.node);
    return key;
}

// count!(quant, fwd, rev, value), src/quant.jl:597-653
void count_pe(const Lib& L, Quant& Q, const Align& fwd, const Align& rev, double value, const BiasTemp* bt = nullptr) {
    if (!(fwd.isvalid && rev.isvalid)) return;
    if (bt) value = 0.0;
    uint32_t init_gene = fwd.path[0].gene, rev_gene = rev.path[0].gene;
    if (init_gene != rev_gene) return;
    size_t nb = L.d->node_base[init_gene - 1];
    bool singlebool = false;
    const std::vector<ANode>* single = nullptr;
    if (fwd.path.size() <= rev.path.size()) {
        if (path_in_path(fwd.path, rev.path)) { singlebool = true; single = &rev.path; }
    } else {
        if (path_in_path(rev.path, fwd.path)) { singlebool = true; single = &fwd.path; }
    }
    if (singlebool) {
        if (single->size() == 1) {
            Q.node[nb + (*single)[0].node - 1] += value;
            if (bt) bias_push(Q.bnode[nb + (*single)[0].node - 1], *bt);
        } else {
            for (auto& n : *single) if (n.gene != init_gene) return;
            if (single->size() == 2) {
                uint32_t l = (*single)[0].node, r = (*single)[1].node;
                const uint64_t key = ((uint64_t)init_gene << 32) | ((uint64_t)(l & 0xFFFF) << 16) | (r & 0xFFFF);
                Q.edge[key] += value;
                if (bt) bias_push(Q.bedge[key], *bt);
            }
        }
    } else {
        const std::vector<uint32_t> key = paired_key(fwd, rev);
        Q.longs[key] += value;
        if (bt) bias_push(Q.blongs[key], *bt);
    }
}

// push!(multi, fwd, rev, value, ...), src/quant.jl:471-484
void push_multi_pe(Quant& Q, const std::vector<Align>& f, const std::vector<Align>& r, double value, const BiasTemp* bt = nullptr) {
    std::vector<uint32_t> key{(3u << 28) | (uint32_t)f.size()};
    for (size_t i = 0; i < f.size(); i++) {
        std::vector<uint32_t> k = paired_key(f[i], r[i]);
        key.insert(key.end(), k.begin(), k.end());
    }
    Q.multi[key] += bt ? 0.0 : value;
    if (bt) bias_push(Q.bmulti[key], *bt);
}

// push!(multi, alns, value, ...), src/quant.jl:456-469: the key is the ordered vector of node paths
void push_multi_se(Quant& Q, const std::vector<Align>& alns, double value, const BiasTemp* bt = nullptr) {
    std::vector<uint32_t> key{(1u << 28) | (uint32_t)alns.size()};
    for (auto& a : alns) {
        key.push_back((uint32_t)a.path.size());
        key.push_back(a.path[0].gene);
        for (auto& n : a.path) key.push_back(n.node);
    }
    Q.multi[key] += bt ? 0.0 : value;
    if (bt) bias_push(Q.bmulti[key], *bt);
}

// primer_normalize!(mod) (src/bias.jl:145-152) + primer_adjust!(quant, mod) + primer_adjust!(multi, mod) (src/quant.jl:200-221, 486-495;
// bin/whippet-quant.jl:156-158).  value!(mod, hept) = back[hept+1] / fore[hept+1] (src/bias.jl:205).
double primer_value(const BiasCounter& c, const BiasMod& m) {
    double count = 0.0;                                     // JointBiasCounter() starts at 0.0
    for (auto& kv : c.map) count += (m.back[kv.first] / m.fore[kv.first]) * (double)kv.second;
    count /= (double)m.forenpos;
    return count;
}
void bias_primer_adjust(Quant& Q) {
    BiasMod& m = Q.mod;
    double foresum = 0.0, backsum = 0.0;                    // integer-valued addends: exact in any order
    for (size_t i = 0; i < m.fore.size(); i++) { foresum += m.fore[i]; backsum += m.back[i]; }
    for (size_t i = 0; i < m.fore.size(); i++) { m.fore[i] = m.fore[i] / foresum; m.back[i] = m.back[i] / backsum; }
    static const BiasCounter none;
    for (size_t i = 0; i < Q.node.size(); i++) Q.node[i] = primer_value(Q.bnode[i], m);
    auto get = [&](auto& store, const auto& key) -> const BiasCounter& { auto it = store.find(key); return it == store.end() ? none : it->second; };
    for (auto& kv : Q.edge) kv.second = primer_value(get(Q.bedge, kv.first), m);
    for (auto& kv : Q.longs) kv.second = primer_value(get(Q.blongs, kv.first), m);
    for (auto& kv : Q.multi) kv.second = primer_value(get(Q.bmulti, kv.first), m);       // mc.count = get(mc.raw)
}
// gc_adjust!(quant, mod) + gc_adjust!(multi, mod) (src/bias.jl:258-269, bin/whippet-quant.jl:167-169).  `value!(mod, i)` with i::Int
// dispatches to the method that reads mod.fore / mod.back -- the normalised HEXAMER tables -- at index `bin` (src/bias.jl:206), not
// gcfore / gcback; gc_normalize! (src/quant.jl:223-238) therefore changes nothing that is read later and is not restated.
double gc_weight(const BiasCounter& c, const BiasMod& m, bool& has) {
    double weight = 0.0;
    int64_t total = 0;
    for (int i = 1; i <= 21; i++) {
        const double v = m.fore[i - 1] > 0.0 ? m.back[i - 1] / m.fore[i - 1] : 1.0;
        weight += v * (double)c.gc[i - 1];
        total += c.gc[i - 1];
    }
    has = total > 0;
    if (has) weight /= (double)total;
    return weight;
}
void bias_gc_adjust(Quant& Q) {
    const BiasMod& m = Q.mod;
    bool has;
    for (size_t i = 0; i < Q.node.size(); i++) { double w = gc_weight(Q.bnode[i], m, has); if (has) Q.node[i] *= w; }
    for (auto& kv : Q.edge) { auto it = Q.bedge.find(kv.first); if (it != Q.bedge.end()) { double w = gc_weight(it->second, m, has); if (has) kv.second *= w; } }
    for (auto& kv : Q.longs) { auto it = Q.blongs.find(kv.first); if (it != Q.blongs.end()) { double w = gc_weight(it->second, m, has); if (has) kv.second *= w; } }
    for (auto& kv : Q.multi) { auto it = Q.bmulti.find(kv.first); if (it != Q.bmulti.end()) { double w = gc_weight(it->second, m, has); if (has) kv.second *= w; } }
}


// ---------------------------------------------------------------------------------------------
// Quantification: isoform compatibility classes, transcript EM, multi-map assignment.
// Canonical orders (declared, because Julia Dict iteration order cannot be reproduced without
// Julia): multi-map classes and long paths in lexicographic key order, isoform
// classes per gene in lexicographic order of their id vectors, edges in (first,last) order (this one
// IS the reference order: IntervalMap iteration).  sum(quant.tpm) uses the fixed adjacent-pair tree
// of tree_sum() (Julia's pairwise mapreduce is not reproducible either).  EM numerics: parity unpinned.

struct EqClass {               // IsoCompat / MultiCompat, src/quant.jl:152-162, 422-435
    std::vector<int32_t> cls;  // 1-based global isoform ids
    std::vector<double> prop;
    double prop_sum = 1.0;
    double count = 0.0;
    bool isdone = false;
    bool postassign = false;
};

struct IsoView {
    const Lib& L;
    Int niso(Int g) const { return L.d->iso_base[g] - L.d->iso_base[g - 1]; }
    Int geneidx(Int g) const { return L.d->iso_base[g - 1]; }
    // node in annopath[p] of gene g (p 1-based inside the gene)
    bool has(Int g, Int p, uint32_t node) const {
        uint32_t i = L.d->iso_base[g - 1] + (uint32_t)p - 1;
        const uint32_t* b = L.d->iso_nodes + L.d->iso_node_ptr[i];
        const uint32_t* e = L.d->iso_nodes + L.d->iso_node_ptr[i + 1];
        return std::binary_search(b, e, node);
    }
    // Base.in(first, last, is::BitSet), src/quant.jl:68-79
    bool edge_in(Int g, Int p, uint32_t first, uint32_t last) const {
        if (has(g, p, first) && has(g, p, last)) {
            for (uint32_t i = first + 1; i + 1 <= last && i < last; i++)
                if (has(g, p, i)) return false;
            return true;
        }
        return false;
    }
    // Base.in(path::SGAlignPath, is::BitSet), src/quant.jl:108-121
    bool path_in(Int g, Int p, const uint32_t* nodes, size_t n) const {
        if (n == 1) return has(g, p, nodes[0]);
        for (size_t i = 0; i + 1 < n; i++)
            if (!edge_in(g, p, nodes[i], nodes[i + 1])) return false;
        return true;
    }
};

struct QuantState {            // GraphLibQuant, src/quant.jl:165-198
    std::vector<double> tpm, count, genetpm;
    std::vector<int64_t> length;
    std::vector<EqClass> classes;                                   // isoform classes, canonical order
    std::vector<std::pair<std::vector<uint32_t>, EqClass>> multi;   // multi-map classes, canonical (key) order
};

// a long / multi record holds one or more paths: [npath(|flags), gene, nodes...]
struct PathRef { uint32_t gene; const uint32_t* nodes; size_t n; const uint32_t* rnodes; size_t rn; };
size_t parse_path(const std::vector<uint32_t>& k, size_t pos, PathRef& out) {
    uint32_t hdr = k[pos];
    bool paired = ((hdr >> 28) & 2) != 0;
    out.gene = k[pos + 1];
    if (!paired) {
        out.n = hdr & 0xFFFF; out.nodes = &k[pos + 2]; out.rnodes = nullptr; out.rn = 0;
        return pos + 2 + out.n;
    }
    out.n = hdr & 0xFF; out.rn = (hdr >> 8) & 0xFF;
    out.nodes = &k[pos + 2]; out.rnodes = &k[pos + 2 + out.n];
    return pos + 2 + out.n + out.rn;
}
bool pathref_in(const IsoView& V, Int g, Int p, const PathRef& r) {   // src/quant.jl:123-124
    return V.path_in(g, p, r.nodes, r.n) && (r.rnodes == nullptr || V.path_in(g, p, r.rnodes, r.rn));
}

void edge_add(Quant& Q, uint32_t gene, uint32_t l, uint32_t r, double v) {
    Q.edge[((uint64_t)gene << 32) | ((uint64_t)(l & 0xFFFF) << 16) | (r & 0xFFFF)] += v;
}

// assign_path!, src/quant.jl:263-324 (single and paired)
// `stale`: what temp_iset holds on entry.  assign_ambig! hands the paired method ambig.iset, which set_equivalence_class! left
// filled with the GLOBAL ISOFORM ids compatible with the class's last alignment (src/quant.jl:493-501, 541-551); the method tests
// mate-2 NODE ids against it and empties it at its end (:283-323), so only the first alignment of a class sees those ids.
void assign_path(const Lib& L, Quant& Q, const PathRef& r, double value, const std::vector<uint32_t>* stale = nullptr) {
    size_t nb = L.d->node_base[r.gene - 1];
    std::vector<uint32_t> temp_iset;
    if (stale && r.rnodes != nullptr) temp_iset = *stale;
    auto in_set = [&](uint32_t x) { return std::find(temp_iset.begin(), temp_iset.end(), x) != temp_iset.end(); };
    bool paired = r.rnodes != nullptr;
    if (r.n == 1) {
        Q.node[nb + r.nodes[0] - 1] += value;
        temp_iset.push_back(r.nodes[0]);
    } else {
        for (size_t i = 0; i + 1 < r.n; i++) {
            uint32_t l = r.nodes[i], rr = r.nodes[i + 1];
            temp_iset.push_back(l);
            temp_iset.push_back(rr);
            edge_add(Q, r.gene, l, rr, value);
        }
    }
    if (!paired) return;
    if (r.rn == 1) {
        if (!in_set(r.rnodes[0])) Q.node[nb + r.rnodes[0] - 1] += value;
    } else {
        for (size_t i = 0; i + 1 < r.rn; i++) {
            uint32_t l = r.rnodes[i], rr = r.rnodes[i + 1];
            if (in_set(l) && in_set(rr)) continue;
            edge_add(Q, r.gene, l, rr, value);
        }
    }
}

// build_equivalence_classes!, src/quant.jl:327-389
void build_equivalence_classes(const Lib& L, Quant& Q, QuantState& S) {
    IsoView V{L};
    const Int G = L.G;
    auto li = Q.longs.begin();
    auto ei = Q.edge.begin();
    std::vector<std::pair<std::vector<uint32_t>, double>> long_snapshot;
    for (Int g = 1; g <= G; g++) {
        std::map<std::vector<int32_t>, double> temp;
        const Int np = V.niso(g), idx = V.geneidx(g), nn = L.nnodes(g);
        std::vector<int32_t> iset;
        auto handle = [&](double value) {
            if (iset.size() == 1) S.count[iset[0] - 1 + idx] += value;
            else if (iset.size() > 1) temp[iset] += value;
        };
        for (Int n = 1; n <= nn; n++) {
            iset.clear();
            for (Int p = 1; p <= np; p++) if (V.has(g, p, (uint32_t)n)) iset.push_back((int32_t)p);
            handle(Q.node[L.d->node_base[g - 1] + n - 1]);
        }
        // edges of this gene in (first,last) order; circ entries (l >= r) are not part of .edge
        ei = Q.edge.lower_bound((uint64_t)g << 32);
        for (; ei != Q.edge.end() && (ei->first >> 32) == (uint64_t)g; ++ei) {
            uint32_t l = (ei->first >> 16) & 0xFFFF, r = ei->first & 0xFFFF;
            if (l >= r) continue;
            iset.clear();
            for (Int p = 1; p <= np; p++) if (V.edge_in(g, p, l, r)) iset.push_back((int32_t)p);
            handle(ei->second);
        }
        // long paths of this gene (first path's gene)
        long_snapshot.clear();
        for (auto& kv : Q.longs) if (kv.first[1] == (uint32_t)g) long_snapshot.push_back(kv);
        for (auto& kv : long_snapshot) {
            PathRef r;
            parse_path(kv.first, 0, r);
            iset.clear();
            for (Int p = 1; p <= np; p++) if (pathref_in(V, g, p, r)) iset.push_back((int32_t)p);
            handle(kv.second);
            assign_path(L, Q, r, kv.second);     // assign_long = true
        }
        for (auto& kv : temp) {
            if (kv.second > 0.0) {
                EqClass c;
                for (auto p : kv.first) c.cls.push_back((int32_t)(p + idx));
                c.prop.assign(c.cls.size(), 1.0 / (double)c.cls.size());
                c.count = kv.second;
                S.classes.push_back(c);
            }
        }
    }
    (void)li;
}

// MultiCompat(graphq, lib, cont), src/quant.jl:431-434 with equivalence_class, :394-417
EqClass make_multicompat(const Lib& L, const std::vector<uint32_t>& key, double rawcount) {
    IsoView V{L};
    EqClass c;
    uint32_t nh = key[0] & 0xFFFF;
    size_t pos = 1;
    std::vector<int32_t> set;
    bool matches_all = true;
    for (uint32_t h = 0; h < nh; h++) {
        PathRef r;
        pos = parse_path(key, pos, r);
        Int g = r.gene;
        bool has_match = false;
        for (Int p = 1; p <= V.niso(g); p++)
            if (pathref_in(V, g, p, r)) { set.push_back((int32_t)(p + V.geneidx(g))); has_match = true; }
        if (!has_match) { matches_all = false; set.clear(); break; }
    }
    std::sort(set.begin(), set.end());
    set.erase(std::unique(set.begin(), set.end()), set.end());
    c.cls = set;
    c.prop.assign(set.size(), 1.0 / (double)set.size());
    c.count = rawcount;              // mc.count = get(mc.raw) after primer_adjust!, src/quant.jl:487-493
    c.isdone = !matches_all;
    c.postassign = !matches_all;
    return c;
}

// fixed adjacent-pair tree over tiles of 1024 (declared canonical order of sum(quant.tpm))
double tree_sum(const double* a, size_t n) {
    if (n == 0) return 0.0;
    std::vector<double> partial;
    double buf[1024];
    for (size_t t = 0; t < n; t += 1024) {
        size_t m = std::min<size_t>(1024, n - t);
        for (size_t i = 0; i < 1024; i++) buf[i] = i < m ? a[t + i] : 0.0;
        for (size_t stride = 1; stride < 1024; stride <<= 1)
            for (size_t i = 0; i + stride < 1024; i += 2 * stride) buf[i] += buf[i + stride];
        partial.push_back(buf[0]);
    }
    if (partial.size() == 1) return partial[0];
    return tree_sum(partial.data(), partial.size());
}

double round_digits(double x, int digits) {   // Base.round(x, digits=d): rint(x*10^d)/10^d, non-finite result -> x
    double sc = std::pow(10.0, digits);
    double r = std::nearbyint(x * sc) / sc;
    if (!std::isfinite(r)) return x;
    return r;
}

// calculate_tpm!, src/quant.jl:655-667
void calculate_tpm(QuantState& S, const std::vector<double>& counts, int64_t readlen, int sig) {
    const size_t I = S.tpm.size();
    for (size_t i = 0; i < I; i++) S.tpm[i] = counts[i] / std::max((double)(S.length[i] - readlen), 1.0);
    double rpk_sum = std::max(tree_sum(S.tpm.data(), I), 1.0);
    for (size_t i = 0; i < I; i++) {
        double v = S.tpm[i] * 1000000.0 / rpk_sum;
        S.tpm[i] = sig > 0 ? round_digits(v, sig) : v;
    }
}

// copy_isdone!, src/quant.jl:702-712
bool copy_isdone(std::vector<double>& dest, const std::vector<double>& src, double denominator, int sig = 4) {
    bool isdifferent = false;
    for (size_t i = 0; i < dest.size(); i++) {
        double val = round_digits(src[i] / denominator, sig);
        if (val != dest[i]) isdifferent = true;
        dest[i] = std::isnan(val) ? 0.0 : val;
    }
    return !isdifferent || denominator == 0;
}

// gene_em!, src/quant.jl:719-769
int64_t gene_em(QuantState& S, int64_t maxit, int sig, int64_t readlen) {
    int64_t it = 1;
    const size_t I = S.count.size();
    std::vector<double> count_temp(I, 1.0), tpm_temp(I, 1.0), prop_temp;
    auto maximize_and_assign = [&](EqClass& eq, bool maximize) {
        if (eq.isdone) return;
        if (maximize) {
            eq.prop_sum = 0.0;
            prop_temp.assign(eq.cls.size(), 0.0);
            for (size_t c = 0; c < eq.cls.size(); c++) {
                double cur = S.tpm[eq.cls[c] - 1];
                prop_temp[c] = cur;
                eq.prop_sum += cur;
            }
            eq.isdone = copy_isdone(eq.prop, prop_temp, eq.prop_sum);
        }
        for (size_t c = 0; c < eq.cls.size(); c++) {
            size_t i = eq.cls[c] - 1;
            if (eq.isdone) S.count[i] += eq.prop[c] * eq.count;
            count_temp[i] += eq.prop[c] * eq.count;
        }
    };
    while (tpm_temp != S.tpm && it < maxit) {
        count_temp = S.count;
        tpm_temp = S.tpm;
        for (auto& kv : S.multi) maximize_and_assign(kv.second, it > 1);
        for (auto& eq : S.classes) maximize_and_assign(eq, it > 1);
        calculate_tpm(S, count_temp, readlen, sig);
        it += 1;
    }
    return it;
}

// set_gene_tpm!, src/quant.jl:248-257
void set_gene_tpm(const Lib& L, QuantState& S) {
    for (Int g = 1; g <= L.G; g++) {
        double t = 0.0;
        for (uint32_t i = L.d->iso_base[g - 1]; i < L.d->iso_base[g]; i++) t += S.tpm[i];
        S.genetpm[g - 1] = t;
    }
}

// assign_ambig!, src/quant.jl:520-553
void assign_ambig(const Lib& L, Quant& Q, QuantState& S) {
    IsoView V{L};
    for (auto& kv : S.multi) {
        const std::vector<uint32_t>& key = kv.first;
        EqClass& mc = kv.second;
        uint32_t nh = key[0] & 0xFFFF;
        std::vector<PathRef> vp(nh);
        size_t pos = 1;
        for (uint32_t h = 0; h < nh; h++) pos = parse_path(key, pos, vp[h]);
        std::vector<double> vals(nh, 0.0);
        if (mc.postassign) {
            for (uint32_t i = 0; i < nh; i++) vals[i] = S.genetpm[vp[i].gene - 1];
        } else {
            for (uint32_t i = 0; i < nh; i++) {
                Int g = vp[i].gene;
                double value = 0.0;      // get_class_value, src/quant.jl:510-518
                for (size_t c = 0; c < mc.cls.size(); c++) {
                    Int p = mc.cls[c] - V.geneidx(g);
                    if (p >= 1 && p <= V.niso(g) && pathref_in(V, g, p, vp[i])) value += mc.prop[c];
                }
                vals[i] = value;
            }
        }
        double s = 0.0;
        for (double v : vals) s += v;
        for (double& v : vals) v /= s;
        std::vector<uint32_t> stale;
        if (!mc.postassign && vp[nh - 1].rnodes != nullptr) {       // ambig.iset after the last set_equivalence_class! of the loop above
            const Int g = vp[nh - 1].gene;
            for (Int p = 1; p <= V.niso(g); p++)
                if (pathref_in(V, g, p, vp[nh - 1])) stale.push_back((uint32_t)(p + V.geneidx(g)));
        }
        for (uint32_t i = 0; i < nh; i++) assign_path(L, Q, vp[i], vals[i] * mc.count, i == 0 ? &stale : nullptr);
    }
}


// ---------------------------------------------------------------------------------------------
// Events: process_events / _process_events, src/events.jl:744-800 with the CLI settings
// isnodeok=false (bin/whippet-quant.jl:181).  Writers (src/io.jl) are out of scope: the computed
// quantities of every output line are returned instead.  Unpinned (no Julia here): IntervalTrees
// iteration order (taken as (first,last) order), `maximum(sgquant.edge)` (taken as the entry with the
// greatest (first,last) key, the IntervalValue isless order), sums over more than a handful of paths
// (sequential here), the random subsampling beyond 1024 paths (flagged, not reproduced).
enum : uint8_t { M_TXST = 0, M_TXEN = 1, M_ALTF = 2, M_ALTL = 3, M_RETI = 4, M_SKIP = 5, M_ALTD = 6, M_ALTA = 7, M_NONE = 8 };

uint8_t motif_of(uint8_t cur, uint8_t next) {     // MOTIF_TABLE, src/events.jl:41-106
    static uint8_t table[64];
    static bool init = false;
    if (!init) {
        for (auto& t : table) t = M_NONE;
        auto set = [&](int idx, uint8_t m) { table[idx] = m; };
        set(0b000000, M_TXST); set(0b010000, M_TXST);
        set(0b011011, M_TXEN); set(0b011001, M_TXEN);
        set(0b010101, M_ALTF); set(0b010100, M_ALTF); set(0b010010, M_ALTF);
        set(0b110001, M_ALTL); set(0b100001, M_ALTL); set(0b001001, M_ALTL);
        set(0b101110, M_RETI);
        set(0b110101, M_SKIP); set(0b110100, M_SKIP); set(0b110010, M_SKIP);
        set(0b100101, M_SKIP); set(0b100100, M_SKIP); set(0b100010, M_SKIP);
        set(0b001101, M_SKIP); set(0b001100, M_SKIP); set(0b001010, M_SKIP);
        set(0b101101, M_ALTD); set(0b101100, M_ALTD); set(0b101010, M_ALTD);
        set(0b110110, M_ALTA); set(0b100110, M_ALTA); set(0b001110, M_ALTA);
        init = true;
    }
    return table[((cur & 7) << 3) | (next & 7)];
}
bool isobligate(uint8_t m) { return m <= 1; }
bool isaltsplice(uint8_t m) { return (m & 0b110) == 0b110 && m != M_NONE; }

typedef std::vector<uint32_t> NodeSet;          // BitSet as an ascending vector
struct EdgeIV { uint32_t first, last; double value; };
struct PsiGraph {                               // src/events.jl:134-141
    std::vector<double> psi, count, length;
    std::vector<NodeSet> nodes;
    uint32_t min = 0, max = 0;
};
struct AmbigCounts { std::vector<uint32_t> paths; std::vector<double> prop; double prop_sum = 1.0, multiplier = 0.0; };

bool set_has(const NodeSet& s, uint32_t x) { return std::binary_search(s.begin(), s.end(), x); }
bool graph_has(const PsiGraph& g, uint32_t x) { for (auto& s : g.nodes) if (set_has(s, x)) return true; return false; }
bool edge_in_set(const EdgeIV& e, const NodeSet& s) {          // in(edge, is::BitSet), src/quant.jl:68-81
    if (!set_has(s, e.first) || !set_has(s, e.last)) return false;
    for (uint32_t i = e.first + 1; i < e.last; i++) if (set_has(s, i)) return false;
    return true;
}
NodeSet set_union(const NodeSet& a, const NodeSet& b) {
    NodeSet r;
    std::set_union(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(r));
    return r;
}
bool hasintersect_terminal(const NodeSet& a, const NodeSet& b) { return a.front() == b.back() || b.front() == a.back(); }

void graph_push_edge(PsiGraph& g, const EdgeIV& e) {           // push!(pgraph, edg; value_bool=false), src/events.jl:284-297
    g.count.push_back(0.0);
    g.length.push_back(0.0);
    g.nodes.push_back(NodeSet{e.first, e.last});
    if (e.first < g.min) g.min = e.first;
    if (e.last > g.max) g.max = e.last;
}

// reduce_graph, src/events.jl:223-266
PsiGraph reduce_graph(const PsiGraph& edges, bool& subsampled, int maxit = 100, size_t maxiso = 1024) {
    PsiGraph graph = edges;
    PsiGraph newgraph;
    newgraph.min = graph.min; newgraph.max = graph.max;
    int it = 1;
    while ((newgraph.nodes != graph.nodes || it == 1) && it <= maxit) {
        if (it > 1) {
            std::swap(graph, newgraph);
            newgraph.nodes.clear(); newgraph.length.clear(); newgraph.count.clear();
            if (graph.nodes.size() > maxiso) {       // reference: random subsample (unseeded shuffle)
                subsampled = true;
                graph.nodes.resize(maxiso); graph.length.resize(maxiso); graph.count.resize(maxiso);
            }
        }
        for (size_t i = 0; i < graph.nodes.size(); i++) {
            bool push_i = false;
            for (size_t j = 0; j < edges.nodes.size(); j++) {
                if (hasintersect_terminal(graph.nodes[i], edges.nodes[j])) {
                    push_i = true;
                    NodeSet joint = set_union(graph.nodes[i], edges.nodes[j]);
                    if (std::find(newgraph.nodes.begin(), newgraph.nodes.end(), joint) != newgraph.nodes.end()) continue;
                    newgraph.nodes.push_back(joint);
                    newgraph.length.push_back(graph.length[i] + edges.length[j]);
                    newgraph.count.push_back(graph.count[i] + edges.count[j]);
                    newgraph.min = std::min(std::min(graph.nodes[i].front(), edges.nodes[j].front()), newgraph.min);
                    newgraph.max = std::max(std::max(graph.nodes[i].back(), edges.nodes[j].back()), newgraph.max);
                }
            }
            if (!push_i && std::find(newgraph.nodes.begin(), newgraph.nodes.end(), graph.nodes[i]) == newgraph.nodes.end()) {
                newgraph.nodes.push_back(graph.nodes[i]);
                newgraph.length.push_back(graph.length[i]);
                newgraph.count.push_back(graph.count[i]);
                newgraph.min = std::min(graph.nodes[i].front(), newgraph.min);
                newgraph.max = std::max(graph.nodes[i].back(), newgraph.max);
            }
        }
        it += 1;
    }
    return newgraph;
}
// reduce_graph_simple, src/events.jl:268-282
PsiGraph reduce_graph_simple(PsiGraph edges, bool& subsampled) {
    if (edges.nodes.size() == 2) {
        if (hasintersect_terminal(edges.nodes[0], edges.nodes[1])) {
            edges.nodes[0] = set_union(edges.nodes[0], edges.nodes[1]);
            edges.length[0] += edges.length[1];
            edges.count[0] += edges.count[1];
            edges.nodes.pop_back(); edges.length.pop_back(); edges.count.pop_back();
        }
        return edges;
    }
    return reduce_graph(edges, subsampled);
}

double round_sigdigits(double x, int sig) {    // Base.round(x, sigdigits=n): hidigit via log10, then round to digits
    if (!std::isfinite(x)) return x;
    if (x == 0.0) return x;     // hidigit(0) = 0 -> round(0 * 10^sig) / 10^sig = 0
    int h = 1 + (int)std::floor(std::log10(std::fabs(x)));
    int digits = sig - h;
    double r;
    if (digits >= 0) { double sc = std::pow(10.0, digits); r = std::nearbyint(x * sc) / sc; }
    else { double sc = std::pow(10.0, -digits); r = std::nearbyint(x / sc) * sc; }
    if (!std::isfinite(r)) return digits > 0 ? x : r;
    return r;
}
void divsignif(std::vector<double>& arr, double divisor, int sig) {     // src/events.jl:802-812
    for (auto& a : arr) a = sig > 0 ? round_sigdigits(a / divisor, sig) : a / divisor;
}
double seqsum(const std::vector<double>& v) { double s = 0.0; for (double x : v) s += x; return s; }

// calculate_psi! (spliced), src/events.jl:831-841
void calculate_psi2(PsiGraph& ig, PsiGraph& eg, const std::vector<double>& counts, int sig) {
    for (size_t i = 0; i < ig.psi.size(); i++) ig.psi[i] = counts[i] / ig.length[i];
    for (size_t i = 0; i < eg.psi.size(); i++) eg.psi[i] = counts[i + ig.psi.size()] / eg.length[i];
    double cnt_sum = seqsum(eg.psi) + seqsum(ig.psi);
    divsignif(ig.psi, cnt_sum, sig);
    divsignif(eg.psi, cnt_sum, sig);
}
// calculate_psi! (tandem), src/events.jl:843-850
void calculate_psi1(PsiGraph& pg, const std::vector<double>& counts, int sig, int64_t readlen) {
    for (size_t i = 0; i < pg.psi.size(); i++) pg.psi[i] = counts[i] / std::max(pg.length[i] - (double)readlen, 1.0);
    divsignif(pg.psi, seqsum(pg.psi), sig);
}
void ambig_step(std::vector<AmbigCounts>& ambig, int64_t it, const std::vector<double>& psi_all, std::vector<double>& count_temp) {
    for (auto& ac : ambig) {
        if (it > 1) {
            ac.prop_sum = 0.0;
            for (size_t k = 0; k < ac.paths.size(); k++) {
                double prev = psi_all[ac.paths[k] - 1];
                ac.prop[k] = prev;
                ac.prop_sum += prev;
            }
        }
        for (size_t k = 0; k < ac.paths.size(); k++) {
            double prop = ac.prop[k] / ac.prop_sum;
            ac.prop[k] = prop;
            count_temp[ac.paths[k] - 1] += prop * ac.multiplier;
        }
    }
}
// spliced_em!, src/events.jl:853-893
int64_t spliced_em(PsiGraph& ig, PsiGraph& eg, std::vector<AmbigCounts>& ambig, int sig, int64_t maxit = 1500) {
    int64_t it = 1;
    std::vector<double> inc_temp(ig.count.size(), 0.0), exc_temp(eg.count.size(), 0.0), count_temp(ig.count.size() + eg.count.size(), 1.0);
    while ((inc_temp != ig.psi || eg.psi != exc_temp) && it < maxit) {
        std::copy(ig.count.begin(), ig.count.end(), count_temp.begin());
        std::copy(eg.count.begin(), eg.count.end(), count_temp.begin() + ig.count.size());
        inc_temp = ig.psi;
        exc_temp = eg.psi;
        std::vector<double> psi_all = ig.psi;
        psi_all.insert(psi_all.end(), eg.psi.begin(), eg.psi.end());
        ambig_step(ambig, it, psi_all, count_temp);
        calculate_psi2(ig, eg, count_temp, sig);
        it += 1;
    }
    return it;
}
// rec_tandem_em!, src/events.jl:895-932 (tail recursion written as a loop)
int64_t rec_tandem_em(PsiGraph& pg, std::vector<AmbigCounts>& ambig, int sig, int64_t readlen, int64_t maxit = 1500) {
    int64_t it = 1;
    std::vector<double> utr_temp(pg.count.size(), 0.0), count_temp(pg.count.size(), 1.0);
    for (;;) {
        count_temp = pg.count;
        utr_temp = pg.psi;
        ambig_step(ambig, it, pg.psi, count_temp);
        calculate_psi1(pg, count_temp, sig, readlen);
        if (utr_temp != pg.psi && it < maxit) it += 1; else break;
    }
    return it;
}

void ambig_account(std::vector<AmbigCounts>& ambig, const std::vector<uint32_t>& iset, double value) {
    for (auto& am : ambig)
        if (am.paths == iset) { am.multiplier += value; return; }
    if (value > 0) {
        AmbigCounts a;
        a.paths = iset;
        a.prop.assign(iset.size(), 1.0 / (double)iset.size());
        a.prop_sum = 1.0;
        a.multiplier = value;
        ambig.push_back(a);
    }
}

struct EventOut {
    uint32_t gene, node; uint8_t motif, kind, flags;     // kind 0 empty line, 1 psi line, 2 utr lines, 3 no line; flags bit0 utr empty, bit1 subsampled
    double psi = 0.0, total = 0.0;
    std::vector<double> utr_psi;                         // kind 2: round.(psi, digits=4) per path (zeros when empty)
    PsiGraph inc, exc;                                   // kind 1 (either may be absent: has_inc / has_exc)
    bool has_inc = false, has_exc = false;
    PsiGraph utr; bool has_utr = false;
    int64_t iterations = 0;
};

struct GeneView {
    const Lib& L; const Quant& Q; Int g;
    std::vector<EdgeIV> edges;                           // sgquant.edge of this gene in (first,last) order
    double node(Int n) const { return Q.node[L.d->node_base[g - 1] + n - 1]; }
    double leng(Int n) const { return (double)L.nodelen(g, n); }      // effective_lengths!, src/quant.jl:684-694
    double edge_value(uint32_t f, uint32_t l) const { for (auto& e : edges) if (e.first == f && e.last == l) return e.value; return 0.0; }
    Int nedgetype() const { return L.nnodes(g) + 1; }
    uint8_t et(Int j) const { return L.edgetype(g, j); }
};

// _process_tandem_utr, src/events.jl:593-645
void process_tandem_utr(const GeneView& V, uint32_t node, uint8_t motif, int64_t readlen, EventOut& out, size_t& len) {
    NodeSet used;
    double total_cnt = 0.0;
    auto push_used = [&](uint32_t x) { if (!set_has(used, x)) { used.push_back(x); std::sort(used.begin(), used.end()); } };
    if (motif == M_TXEN) { push_used(node - 1); total_cnt += V.node(node - 1); }
    Int i = node;
    uint8_t curmotif = motif;
    while (curmotif == motif) {
        push_used((uint32_t)i);
        total_cnt += V.node(i);
        if (i > 1) total_cnt += V.edge_value((uint32_t)i - 1, (uint32_t)i);
        i += 1;
        if (i < V.nedgetype()) curmotif = motif_of(V.et(i), V.et(i + 1)); else break;
    }
    if (motif == M_TXST && i <= V.L.nnodes(V.g)) {     // (the reference would throw a BoundsError past the last node)
        push_used((uint32_t)i);
        total_cnt += V.node(i);
        total_cnt += V.edge_value((uint32_t)i - 1, (uint32_t)i);
    }
    len = used.size();
    if (total_cnt > 2) {
        PsiGraph g;
        g.min = used.front(); g.max = used.back();
        NodeSet nodes = used, curset;
        while (!nodes.empty()) {                      // build_utr_graph, src/events.jl:570-591
            uint32_t cur;
            if (motif == M_TXST) { cur = nodes.back(); nodes.pop_back(); } else { cur = nodes.front(); nodes.erase(nodes.begin()); }
            curset.push_back(cur);
            std::sort(curset.begin(), curset.end());
            g.nodes.push_back(curset); g.psi.push_back(0.0); g.length.push_back(0.0); g.count.push_back(0.0);
        }
        std::vector<AmbigCounts> ambig;
        std::vector<uint32_t> iset;
        for (uint32_t n = g.min; n <= g.max; n++) {   // add_node_counts! (one graph), src/events.jl:397-435
            iset.clear();
            for (size_t k = 0; k < g.nodes.size(); k++) if (set_has(g.nodes[k], n)) { iset.push_back((uint32_t)k + 1); g.length[k] += V.leng(n); }
            if (iset.empty()) continue;
            if (iset.size() == 1) g.count[iset[0] - 1] += V.node(n);
            else ambig_account(ambig, iset, V.node(n));
        }
        for (auto& edg : V.edges) {                   // add_edge_counts! (one graph), src/events.jl:490-523
            iset.clear();
            for (size_t k = 0; k < g.nodes.size(); k++) if (edge_in_set(edg, g.nodes[k])) iset.push_back((uint32_t)k + 1);
            if (iset.empty()) continue;
            if (iset.size() == 1) g.count[iset[0] - 1] += edg.value;
            else ambig_account(ambig, iset, edg.value);
        }
        out.iterations = rec_tandem_em(g, ambig, 4, readlen);
        out.has_utr = true;
        double amb = 0.0;
        for (auto& a : ambig) amb += a.multiplier;
        out.total = seqsum(g.count) + amb;
        out.utr = g;
    }
}

// _process_spliced, src/events.jl:647-742
void process_spliced(const GeneView& V, uint32_t node, uint8_t motif, bool isnodeok, EventOut& out, bool& has_psi, double minedgeweight = 0.02) {
    PsiGraph inc, exc;
    bool has_inc = false, has_exc = false, has_ambig_edge = false;
    std::vector<EdgeIV> ambig_edge;
    std::vector<AmbigCounts> ambig_cnt;
    double connecting_val = 0.0, spanning_val = 0.0;
    double maxvalue = V.edges.empty() ? 0.0 : V.edges.back().value;     // maximum(sgquant.edge): greatest (first,last) key
    for (auto& edg : V.edges) {
        if (!(edg.first <= node && node <= edg.last)) continue;         // intersect(sgquant.edge, (node,node))
        if (edg.value / maxvalue < minedgeweight) continue;
        if (edg.first == node || edg.last == node) {
            if (!has_inc) { inc = PsiGraph(); inc.min = edg.first; inc.max = edg.last; has_inc = true; }
            has_ambig_edge = true;
            graph_push_edge(inc, edg);
            ambig_edge.push_back(edg);
            connecting_val += edg.value;
        } else {
            if (!has_exc) { exc = PsiGraph(); exc.min = edg.first; exc.max = edg.last; has_exc = true; }
            has_ambig_edge = true;
            graph_push_edge(exc, edg);
            ambig_edge.push_back(edg);
            spanning_val += edg.value;
        }
    }
    bool sub = false;
    if (has_exc && has_inc) {
        // extend_edges!(sgquant.edge, exc, inc, ambig_edge, node), src/events.jl:329-394
        uint32_t minv = std::min(exc.min, inc.min), maxv = std::max(exc.max, inc.max);
        Int idx = (Int)node - 1;
        while (idx > (Int)minv) {
            for (auto& edg : V.edges) {
                if (!(edg.first <= idx && idx <= edg.last)) continue;
                if ((edg.first == idx || edg.last == idx) && edg.last == idx && edg.first >= minv) {
                    bool shouldpush = false;
                    if (graph_has(exc, edg.last)) { graph_push_edge(exc, edg); shouldpush = true; }
                    if (graph_has(inc, edg.last)) { graph_push_edge(inc, edg); shouldpush = true; }
                    if (shouldpush) { if (edg.first < minv) minv = edg.first; ambig_edge.push_back(edg); }
                }
            }
            idx -= 1;
        }
        idx = (Int)node + 1;
        while (idx < (Int)maxv) {
            for (auto& edg : V.edges) {
                if (!(edg.first <= idx && idx <= edg.last)) continue;
                if ((edg.first == idx || edg.last == idx) && edg.first == idx && edg.last <= maxv) {
                    bool shouldpush = false;
                    if (graph_has(exc, edg.first)) { graph_push_edge(exc, edg); shouldpush = true; }
                    if (graph_has(inc, edg.first)) { graph_push_edge(inc, edg); shouldpush = true; }
                    if (shouldpush) { if (edg.last > maxv) maxv = edg.last; ambig_edge.push_back(edg); }
                }
            }
            idx += 1;
        }
        inc = reduce_graph(inc, sub);
        exc = reduce_graph(exc, sub);
        // add_edge_counts! (two graphs), src/events.jl:525-568
        std::vector<uint32_t> iset;
        for (auto& edg : ambig_edge) {
            iset.clear();
            for (size_t k = 0; k < inc.nodes.size(); k++) if (edge_in_set(edg, inc.nodes[k])) { iset.push_back((uint32_t)k + 1); inc.length[k] += 1; }
            for (size_t k = 0; k < exc.nodes.size(); k++) if (edge_in_set(edg, exc.nodes[k])) { iset.push_back((uint32_t)(k + 1 + inc.nodes.size())); exc.length[k] += 1; }
            if (iset.empty()) continue;
            if (iset.size() == 1) {
                if (iset[0] <= inc.nodes.size()) inc.count[iset[0] - 1] += edg.value;
                else exc.count[iset[0] - inc.nodes.size() - 1] += edg.value;
            } else ambig_account(ambig_cnt, iset, edg.value);
        }
        (void)isnodeok;     // CLI: false (bin/whippet-quant.jl:181); node counts are not used for spliced events
    } else {
        if (has_inc) inc = reduce_graph_simple(inc, sub);
        if (has_exc) exc = reduce_graph_simple(exc, sub);
    }
    (void)has_ambig_edge;
    has_psi = false;
    double psi = 0.0;
    if (!has_exc) {
        if (has_inc && (connecting_val >= 2 || (connecting_val >= 1 && isaltsplice(motif)))) { psi = 1.0; has_psi = true; }
    } else {
        if (!has_inc) {
            if (spanning_val >= 1) { psi = 0.0; has_psi = true; }
        } else {
            exc.psi.assign(exc.count.size(), 0.0);
            inc.psi.assign(inc.count.size(), 0.0);
            std::vector<double> counts = inc.count;
            counts.insert(counts.end(), exc.count.begin(), exc.count.end());
            calculate_psi2(inc, exc, counts, 0);
            out.iterations = spliced_em(inc, exc, ambig_cnt, 4);
            psi = seqsum(inc.psi);
            has_psi = true;
        }
    }
    out.psi = psi;
    out.total = connecting_val + spanning_val;
    out.inc = inc; out.exc = exc; out.has_inc = has_inc; out.has_exc = has_exc;
    if (sub) out.flags |= 2;
}

// _process_events, src/events.jl:760-800
void process_events_gene(const Lib& L, const Quant& Q, Int g, int64_t readlen, std::vector<EventOut>& outs) {
    GeneView V{L, Q, g, {}};
    if (V.nedgetype() <= 2) return;
    auto lo = Q.edge.lower_bound((uint64_t)g << 32);
    for (auto it = lo; it != Q.edge.end() && (it->first >> 32) == (uint64_t)g; ++it) {
        uint32_t l = (it->first >> 16) & 0xFFFF, r = it->first & 0xFFFF;
        if (l < r) V.edges.push_back(EdgeIV{l, r, it->second});
    }
    Int i = 1;
    const Int ne = V.nedgetype();
    while (i < ne) {
        uint8_t motif = motif_of(V.et(i), V.et(i + 1));
        bool next_istandem = (i + 1 < ne) ? isobligate(motif_of(V.et(i + 1), V.et(i + 2))) : false;
        EventOut o;
        o.gene = (uint32_t)g; o.node = (uint32_t)i; o.motif = motif; o.kind = 0; o.flags = 0;
        if (motif == M_NONE) {
            o.kind = next_istandem ? 3 : 0;
            outs.push_back(o);
        } else if (isobligate(motif)) {
            size_t len = 0;
            process_tandem_utr(V, (uint32_t)i, motif, readlen, o, len);
            o.kind = 2;
            bool anynan = false;
            if (o.has_utr) for (double x : o.utr.psi) if (std::isnan(x)) anynan = true;
            if (o.has_utr && !anynan) {
                for (double x : o.utr.psi) o.utr_psi.push_back(round_digits(x, 4));
            } else {
                o.utr_psi.assign(len, 0.0);
                o.flags |= 1;
                o.total = 0.0;
            }
            Int st = motif == M_TXST ? i : i - 1;
            Int en = st + (Int)o.utr_psi.size() - 1;
            outs.push_back(o);
            i = en;
        } else {
            bool has_psi = false;
            process_spliced(V, (uint32_t)i, motif, false, o, has_psi);
            o.kind = (has_psi && 0 <= o.psi && o.psi <= 1 && o.total > 0) ? 1 : 0;
            outs.push_back(o);
        }
        i += 1;
    }
}

}  // namespace

struct wqo_handle {
    Lib L;
    Result R;
    std::string err;
    Quant Q;
    QuantState S;
    std::vector<EventOut> events;
};

extern "C" {

wqo_handle* wqo_create(const wq_index_desc* d) {
    init_phred();
    wqo_handle* h = new wqo_handle{Lib{d, {}, (Int)d->text_len, (Int)d->n_genes}, {}, {}};
    build_occ(h->L);
    h->Q.node.assign(d->n_nodes, 0.0);
    return h;
}
void wqo_quant_reset(wqo_handle* h) {
    const bool bias = h->Q.bias;
    h->Q = Quant();
    h->Q.node.assign(h->L.d->n_nodes, 0.0);
    h->Q.bias = bias;
    if (bias) h->Q.bnode.assign(h->L.d->n_nodes, BiasCounter());
}
// CounterType / BiasModelType of bin/whippet-quant.jl:116-122; takes effect at the next wqo_quant_reset
void wqo_set_bias(wqo_handle* h, int on) { h->Q.bias = on != 0; }
void wqo_bias_primer_adjust(wqo_handle* h) { if (h->Q.bias) bias_primer_adjust(h->Q); }
void wqo_bias_gc_adjust(wqo_handle* h) {
    if (!h->Q.bias) return;
    bias_gc_adjust(h->Q);
    for (auto& kv : h->S.multi) {                           // mc.count = get(mc.raw), src/quant.jl:486-495: the classes of the EM see the new count
        auto it = h->Q.multi.find(kv.first);
        if (it != h->Q.multi.end()) kv.second.count = it->second;
    }
}
void wqo_bias_model(wqo_handle* h, double* fore4096, double* back4096, uint64_t* short_reads) {
    memcpy(fore4096, h->Q.mod.fore.data(), 4096 * 8); memcpy(back4096, h->Q.mod.back.data(), 4096 * 8);
    *short_reads = h->Q.bias_short;
}
// known-answer helpers (test/runtests.jl:62-103): the GC bin of a window with k G/C bases out of n, a counter holding `hept` twice
int64_t wqo_bias_gc_bin(int64_t k, int64_t n, double width) { return (int64_t)(jl_div((double)k / (double)n, width) + 1); }
int64_t wqo_bias_expected_gc_bins(double width) { return (int64_t)(jl_div(1.0, width) + 1); }

// process_reads!, src/reads.jl:41-80 (DefaultBiasMod: biasval = 1.0)
int wqo_process_reads_se(wqo_handle* h, const wq_align_param* p, const wq_read_batch* b) {
    Quant& Q = h->Q;
    std::vector<Align> res;
    for (uint32_t i = 0; i < b->n_reads; i++) {
        Read r = load_read(b, i);
        bool fl;
        bool ok = ungapped_align_se(h->L, *p, r, res, fl);
        Q.total += 1;
        if (ok) {
            g_ctr.out_alns += res.size();
            for (auto& a : res) g_ctr.out_nodes += a.path.size();
            g_ctr.count_updates += 1;
            const BiasTemp* bt = nullptr;
            if (Q.bias) {                                    // biasval = count!(mod, reads[i].sequence): the read as ungapped_align left it
                if (!bias_count(Q.mod, r.seq)) { Q.bias_short += 1; return -2; }
                bt = &Q.mod.temp;
            }
            if (res.size() > 1) { push_multi_se(Q, res, 1.0, bt); Q.nmulti += 1; }
            else count_se(h->L, Q, res[0], 1.0, bt);
            Q.mapped += 1;
            Q.sum_readlen += r.seq.size();
            Q.mean_readlen += ((double)r.seq.size() - Q.mean_readlen) / (double)Q.mapped;
        }
    }
    return 0;
}
// process_paired_reads!, src/reads.jl:83-129
int wqo_process_reads_pe(wqo_handle* h, const wq_align_param* p, const wq_read_batch* f, const wq_read_batch* rv) {
    Quant& Q = h->Q;
    std::vector<Align> fr, rr;
    for (uint32_t i = 0; i < f->n_reads; i++) {
        Read a = load_read(f, i), b = load_read(rv, i);
        uint8_t fl;
        bool ok = ungapped_align_pe(h->L, *p, a, b, fr, rr, fl);
        Q.total += 1;
        if (ok) {
            const BiasTemp* bt = nullptr;
            if (Q.bias) {                                    // biasval = count!(mod, fwd.sequence, rev.sequence)
                if (!bias_count(Q.mod, a.seq, &b.seq)) { Q.bias_short += 1; return -2; }
                bt = &Q.mod.temp;
            }
            if (fr.size() > 1) { push_multi_pe(Q, fr, rr, 1.0, bt); Q.nmulti += 1; }
            else count_pe(h->L, Q, fr[0], rr[0], 1.0, bt);
            g_ctr.out_alns += 2 * fr.size();
            for (auto& a2 : fr) g_ctr.out_nodes += a2.path.size();
            for (auto& a2 : rr) g_ctr.out_nodes += a2.path.size();
            g_ctr.count_updates += 1;
            Q.mapped += 1;
            Q.sum_readlen += a.seq.size();
            Q.mean_readlen += ((double)a.seq.size() - Q.mean_readlen) / (double)Q.mapped;
        }
    }
    return 0;
}
// add the count stores of `src` into `dst` (reads sharded over host threads for the timed CPU baseline;
// every read contributes the integer 1, so the merged stores equal a single sequential pass)
void wqo_merge(wqo_handle* dst, const wqo_handle* src) {
    Quant& A = dst->Q;
    const Quant& B = src->Q;
    for (size_t i = 0; i < A.node.size(); i++) A.node[i] += B.node[i];
    for (auto& kv : B.edge) A.edge[kv.first] += kv.second;
    for (auto& kv : B.longs) A.longs[kv.first] += kv.second;
    for (auto& kv : B.multi) A.multi[kv.first] += kv.second;
    A.total += B.total; A.mapped += B.mapped; A.nmulti += B.nmulti; A.sum_readlen += B.sum_readlen;
}
void wqo_stats(wqo_handle* h, uint64_t* out5, double* mean_readlen) {
    out5[0] = h->Q.total; out5[1] = h->Q.mapped; out5[2] = h->Q.nmulti; out5[3] = h->Q.sum_readlen; out5[4] = 0;
    *mean_readlen = h->Q.mean_readlen;
}
// direct injection of counts (used by known-answer tests that start from given alignments)
void wqo_add_node(wqo_handle* h, uint32_t gene, uint32_t node, double v) { h->Q.node[h->L.d->node_base[gene - 1] + node - 1] += v; }
void wqo_add_edge(wqo_handle* h, uint32_t gene, uint32_t l, uint32_t r, double v) { edge_add(h->Q, gene, l, r, v); }
void wqo_add_keyed(wqo_handle* h, int kind, const uint32_t* words, uint64_t nw, double v) {
    std::vector<uint32_t> k(words, words + nw);
    (kind == 0 ? h->Q.longs : h->Q.multi)[k] += v;
}

// bin/whippet-quant.jl:161-166: build_equivalence_classes!, calculate_tpm!, gene_em!, set_gene_tpm!
int64_t wqo_em_tpm(wqo_handle* h, int64_t readlen, int64_t sig, int64_t maxit) {
    const Lib& L = h->L;
    QuantState& S = h->S;
    S = QuantState();
    const uint32_t I = L.d->n_isoforms;
    S.tpm.assign(I, 0.0); S.count.assign(I, 0.0); S.genetpm.assign(L.G, 0.0); S.length.assign(I, 0);
    for (Int g = 1; g <= L.G; g++)
        for (uint32_t i = L.d->iso_base[g - 1]; i < L.d->iso_base[g]; i++) {
            int64_t len = 0;
            for (uint32_t k = L.d->iso_node_ptr[i]; k < L.d->iso_node_ptr[i + 1]; k++) len += L.nodelen(g, L.d->iso_nodes[k]);
            S.length[i] = len;
        }
    for (auto& kv : h->Q.multi) S.multi.emplace_back(kv.first, make_multicompat(L, kv.first, kv.second));
    build_equivalence_classes(L, h->Q, S);
    calculate_tpm(S, S.count, readlen, 1);       // calculate_tpm!(quant, readlen = readlen): its own default sig = 1 (bin/whippet-quant.jl:163)
    int64_t it = gene_em(S, maxit, (int)sig, readlen);
    set_gene_tpm(L, S);
    return it;
}
// The transcript EM on a problem given directly (known-answer tests): I isoforms with lengths and unique counts, n_multi multi-map
// classes followed by the isoform classes (cls_ptr / cls_ids 1-based / cls_count), visited in that order.  The same
// calculate_tpm / gene_em as wqo_em_tpm runs (initial TPM with calculate_tpm!'s own default sig = 1, bin/whippet-quant.jl:163).
// trace (optional) receives the TPM vector after each of the first trace_n iterations.
int64_t wqo_gene_em_direct(uint32_t I, const int64_t* length, const double* count, uint32_t n_multi, uint32_t n_cls, const uint32_t* cls_ptr,
                           const int32_t* cls_ids, const double* cls_count, int64_t readlen, int64_t sig, int64_t maxit,
                           double* tpm_out, double* count_out, double* trace, uint32_t trace_n) {
    QuantState S;
    S.tpm.assign(I, 0.0); S.count.assign(count, count + I); S.genetpm.clear(); S.length.assign(length, length + I);
    for (uint32_t c = 0; c < n_cls; c++) {
        EqClass q;
        q.cls.assign(cls_ids + cls_ptr[c], cls_ids + cls_ptr[c + 1]);
        q.prop.assign(q.cls.size(), 1.0 / (double)q.cls.size());
        q.count = cls_count[c];
        if (c < n_multi) S.multi.emplace_back(std::vector<uint32_t>{c}, q); else S.classes.push_back(q);
    }
    calculate_tpm(S, S.count, readlen, 1);
    int64_t it = 1;
    if (trace && trace_n) {            // iteration by iteration: gene_em with maxit = it + 1 runs exactly one more sweep
        for (uint32_t t = 0; t < trace_n; t++) {
            QuantState T = S;
            // re-run from the start with a bounded iteration count (the class state lives inside S)
            T = QuantState();
            T.tpm.assign(I, 0.0); T.count.assign(count, count + I); T.length = S.length;
            T.multi = S.multi; T.classes = S.classes;
            calculate_tpm(T, T.count, readlen, 1);
            gene_em(T, (int64_t)t + 2, (int)sig, readlen);
            memcpy(trace + (size_t)t * I, T.tpm.data(), 8 * (size_t)I);
        }
    }
    it = gene_em(S, maxit, (int)sig, readlen);
    memcpy(tpm_out, S.tpm.data(), 8 * (size_t)I);
    memcpy(count_out, S.count.data(), 8 * (size_t)I);
    return it;
}
void wqo_em_results(wqo_handle* h, double* tpm, double* count, double* genetpm) {
    memcpy(tpm, h->S.tpm.data(), h->S.tpm.size() * 8);
    memcpy(count, h->S.count.data(), h->S.count.size() * 8);
    memcpy(genetpm, h->S.genetpm.data(), h->S.genetpm.size() * 8);
}
uint64_t wqo_n_classes(wqo_handle* h, int multi) { return multi ? h->S.multi.size() : h->S.classes.size(); }
// class i: ids into out_ids (cap), returns size; props/count/flags through pointers
uint64_t wqo_class(wqo_handle* h, int multi, uint64_t i, int32_t* ids, double* props, double* count, int32_t* flags) {
    EqClass& c = multi ? h->S.multi[i].second : h->S.classes[i];
    for (size_t k = 0; k < c.cls.size(); k++) { ids[k] = c.cls[k]; props[k] = c.prop[k]; }
    *count = c.count;
    *flags = (c.isdone ? 1 : 0) | (c.postassign ? 2 : 0);
    return c.cls.size();
}
void wqo_assign_ambig(wqo_handle* h) { assign_ambig(h->L, h->Q, h->S); }

// process_events (src/events.jl:744-758) over all genes, on the stores as they stand (after assign_ambig!)
uint64_t wqo_process_events(wqo_handle* h, int64_t readlen) {
    h->events.clear();
    for (Int g = 1; g <= h->L.G; g++) process_events_gene(h->L, h->Q, g, readlen, h->events);
    return h->events.size();
}
// event i: header[8] = gene, node, motif, kind, flags, n_utr, n_inc, n_exc; vals[3] = psi, total, iterations
void wqo_event(wqo_handle* h, uint64_t i, uint32_t* header, double* vals) {
    const EventOut& e = h->events[i];
    header[0] = e.gene; header[1] = e.node; header[2] = e.motif; header[3] = e.kind; header[4] = e.flags;
    header[5] = (uint32_t)e.utr_psi.size(); header[6] = e.has_inc ? (uint32_t)e.inc.nodes.size() : 0; header[7] = e.has_exc ? (uint32_t)e.exc.nodes.size() : 0;
    vals[0] = e.psi; vals[1] = e.total; vals[2] = (double)e.iterations;
}
// All event records at once in a canonical flat form (what the writers of src/io.jl would print from):
//   hdr[8 * i ..]  gene, node, motif, kind, flags, n_utr, n_inc, n_exc      dbl[3 * i ..]  psi, total, iterations
//   values: kind 2 -> the n_utr rounded psi values; kind 1 with both graphs -> inclusion then exclusion path psi values
//   paths (kind != 2): inclusion then exclusion node sets: path_len[] per path, path_nodes[] concatenated
// Two calls: with NULL arrays the sizes come back in n[0..2] = {values, paths, path nodes}.
void wqo_events_flat(wqo_handle* h, uint64_t* n, uint32_t* hdr, double* dbl, double* values, uint32_t* path_len, uint32_t* path_nodes) {
    uint64_t nv = 0, np = 0, nn = 0;
    for (size_t i = 0; i < h->events.size(); i++) {
        const EventOut& e = h->events[i];
        const uint32_t n_utr = (uint32_t)e.utr_psi.size(), n_inc = e.has_inc ? (uint32_t)e.inc.nodes.size() : 0, n_exc = e.has_exc ? (uint32_t)e.exc.nodes.size() : 0;
        if (hdr) {
            uint32_t* x = hdr + 8 * i;
            x[0] = e.gene; x[1] = e.node; x[2] = e.motif; x[3] = e.kind; x[4] = e.flags; x[5] = n_utr; x[6] = n_inc; x[7] = n_exc;
            dbl[3 * i] = e.psi; dbl[3 * i + 1] = e.total; dbl[3 * i + 2] = (double)e.iterations;
        }
        if (e.kind == 2) {
            if (values) memcpy(values + nv, e.utr_psi.data(), 8 * (size_t)n_utr);
            nv += n_utr;
            continue;
        }
        if (e.kind == 1 && n_inc && n_exc) {
            if (values) { memcpy(values + nv, e.inc.psi.data(), 8 * (size_t)n_inc); memcpy(values + nv + n_inc, e.exc.psi.data(), 8 * (size_t)n_exc); }
            nv += n_inc + n_exc;
        }
        for (int w = 0; w < 2; w++) {
            if (!(w == 0 ? e.has_inc : e.has_exc)) continue;
            const PsiGraph& g = w == 0 ? e.inc : e.exc;
            for (auto& st : g.nodes) {
                if (path_len) { path_len[np] = (uint32_t)st.size(); memcpy(path_nodes + nn, st.data(), 4 * st.size()); }
                np++; nn += st.size();
            }
        }
    }
    n[0] = nv; n[1] = np; n[2] = nn;
}
// path psi values: which 0 = utr (rounded psi), 1 = inclusion paths, 2 = exclusion paths
void wqo_event_psi(wqo_handle* h, uint64_t i, int which, double* out) {
    const EventOut& e = h->events[i];
    const std::vector<double>& v = which == 0 ? e.utr_psi : which == 1 ? e.inc.psi : e.exc.psi;
    memcpy(out, v.data(), v.size() * 8);
}
// node set of path k of graph `which` (0 utr, 1 inc, 2 exc): returns size, writes ids
uint64_t wqo_event_path(wqo_handle* h, uint64_t i, int which, uint64_t k, uint32_t* out) {
    const EventOut& e = h->events[i];
    const PsiGraph& g = which == 0 ? e.utr : which == 1 ? e.inc : e.exc;
    if (k >= g.nodes.size()) return 0;
    memcpy(out, g.nodes[k].data(), g.nodes[k].size() * 4);
    return g.nodes[k].size();
}
// One EM problem given directly (checker for wq_em_psi_batch): the first n_inc paths are inclusion paths, the rest
// exclusion paths (spliced_em!, src/events.jl:853-893), or all paths of a tandem-UTR event (rec_tandem_em!, :895-932).
// AmbigCounts.paths are 1-based here as in the reference.  Returns the iteration count; psi[np] receives the psi values.
int64_t wqo_em_psi(uint32_t np, uint32_t n_inc, const double* count, const double* length, uint32_t n_amb, const uint32_t* amb_ptr,
                   const uint32_t* amb_paths, const double* amb_mult, int tandem, int64_t readlen, int sig, int64_t maxit, double* psi) {
    std::vector<AmbigCounts> ambig(n_amb);
    for (uint32_t a = 0; a < n_amb; a++) {
        const uint32_t n = amb_ptr[a + 1] - amb_ptr[a];
        for (uint32_t k = 0; k < n; k++) ambig[a].paths.push_back(amb_paths[amb_ptr[a] + k] + 1);
        ambig[a].prop.assign(n, 1.0 / (double)n);           // AmbigCounts(paths, ones(n) / n, 1.0, multiplier), src/events.jl:309-326
        ambig[a].prop_sum = 1.0;
        ambig[a].multiplier = amb_mult[a];
    }
    int64_t it;
    if (tandem) {
        PsiGraph pg;
        pg.count.assign(count, count + np); pg.length.assign(length, length + np); pg.psi.assign(np, 0.0);
        it = rec_tandem_em(pg, ambig, sig, readlen, maxit);
        for (uint32_t i = 0; i < np; i++) psi[i] = pg.psi[i];
    } else {
        PsiGraph ig, eg;
        ig.count.assign(count, count + n_inc); ig.length.assign(length, length + n_inc); ig.psi.assign(n_inc, 0.0);
        eg.count.assign(count + n_inc, count + np); eg.length.assign(length + n_inc, length + np); eg.psi.assign(np - n_inc, 0.0);
        std::vector<double> counts(count, count + np);
        calculate_psi2(ig, eg, counts, 0);                  // calculate_psi!(inc, exc, counts) before spliced_em! (src/events.jl:733)
        it = spliced_em(ig, eg, ambig, sig, maxit);
        for (uint32_t i = 0; i < n_inc; i++) psi[i] = ig.psi[i];
        for (uint32_t i = n_inc; i < np; i++) psi[i] = eg.psi[i - n_inc];
    }
    return it;
}

// known-answer hook: Base.in(first::SGAlignPath, last::SGAlignPath) (test/runtests.jl:421-438); paths as (gene, node) pairs
int wqo_path_in_path(const uint32_t* first, uint64_t nf, const uint32_t* last, uint64_t nl) {
    std::vector<ANode> a(nf), b(nl);
    for (uint64_t i = 0; i < nf; i++) { a[i].gene = first[2 * i]; a[i].node = first[2 * i + 1]; }
    for (uint64_t i = 0; i < nl; i++) { b[i].gene = last[2 * i]; b[i].node = last[2 * i + 1]; }
    return path_in_path(a, b) ? 1 : 0;
}

// known-answer hook: reduce_graph on a list of weighted edges (test/runtests.jl:535-580)
uint64_t wqo_reduce_graph(const uint32_t* first, const uint32_t* last, const double* value, uint64_t n, uint32_t* out_sets, uint64_t cap) {
    PsiGraph g;
    g.min = 0xFFFFFFFFu; g.max = 0;
    for (uint64_t i = 0; i < n; i++) {
        g.psi.push_back(0.0); g.length.push_back(1.0); g.count.push_back(value[i]);
        g.nodes.push_back(NodeSet{first[i], last[i]});
        g.min = std::min(g.min, first[i]); g.max = std::max(g.max, last[i]);
    }
    bool sub = false;
    PsiGraph r = reduce_graph(g, sub);
    uint64_t w = 0;
    for (auto& s : r.nodes) {
        if (w + 1 + s.size() > cap) return 0;
        out_sets[w++] = (uint32_t)s.size();
        for (uint32_t x : s) out_sets[w++] = x;
    }
    return r.nodes.size();
}

void wqo_counters(uint64_t* out7, int reset) {
    out7[0] = g_ctr.fm_steps; out7[1] = g_ctr.hits; out7[2] = g_ctr.nodes_touched; out7[3] = g_ctr.bases;
    out7[4] = g_ctr.out_nodes; out7[5] = g_ctr.out_alns; out7[6] = g_ctr.count_updates;
    if (reset) g_ctr = Counters();
}
void wqo_steps_hist(uint64_t* out34) { memcpy(out34, g_ctr.steps_hist, sizeof g_ctr.steps_hist); }
void wqo_node_counts(wqo_handle* h, double* out) { memcpy(out, h->Q.node.data(), h->Q.node.size() * 8); }
uint64_t wqo_edge_count(wqo_handle* h) { return h->Q.edge.size(); }
void wqo_edges(wqo_handle* h, uint64_t* keys, double* vals) {
    size_t i = 0;
    for (auto& kv : h->Q.edge) { keys[i] = kv.first; vals[i] = kv.second; i++; }
}
// keyed stores: kind 0 = long, 1 = multi.  First call with NULLs returns sizes.
void wqo_keyed(wqo_handle* h, int kind, uint64_t* n_rec, uint64_t* n_words, uint64_t* rec_ptr, uint32_t* words, double* vals) {
    auto& M = kind == 0 ? h->Q.longs : h->Q.multi;
    uint64_t nr = 0, nw = 0;
    for (auto& kv : M) {
        if (rec_ptr) {
            rec_ptr[nr] = nw;
            memcpy(words + nw, kv.first.data(), kv.first.size() * 4);
            vals[nr] = kv.second;
        }
        nr++;
        nw += kv.first.size();
    }
    if (rec_ptr) rec_ptr[nr] = nw;
    *n_rec = nr;
    *n_words = nw;
}
void wqo_destroy(wqo_handle* h) { delete h; }

// FM search of one 2-bit query: returns sp, ep (src/fmindex_patch.jl:19-31)
void wqo_sa_range(wqo_handle* h, const uint8_t* q, int64_t len, int64_t* sp, int64_t* ep) {
    Int a, b;
    sa_range(h->L, q, len, a, b);
    *sp = a;
    *ep = b;
}
int64_t wqo_sa_value(wqo_handle* h, int64_t row) { return sa_value(h->L, row); }

// seed_locate for every read of a batch: out[4*i..] = sp, ep, readloc, strand
void wqo_seed_locate(wqo_handle* h, const wq_align_param* p, const wq_read_batch* b, int64_t* out) {
    for (uint32_t i = 0; i < b->n_reads; i++) {
        Read r = load_read(b, i);
        Seed s = seed_locate(h->L, *p, r, true, !p->is_stranded);
        out[4 * i] = s.sp;
        out[4 * i + 1] = s.ep;
        out[4 * i + 2] = s.readloc;
        out[4 * i + 3] = s.strand;
    }
}

int wqo_align_se(wqo_handle* h, const wq_align_param* p, const wq_read_batch* b) {
    Result& R = h->R;
    R = Result();
    std::vector<Align> res;
    for (uint32_t i = 0; i < b->n_reads; i++) {
        Read r = load_read(b, i);
        bool fl;
        bool ok = ungapped_align_se(h->L, *p, r, res, fl);
        R.hit_start.push_back((uint32_t)R.aln.size());
        R.nhits.push_back(ok ? (uint8_t)res.size() : 0);
        R.flipped.push_back(fl);
        if (ok) for (auto& a : res) emit(R, a, R.aln);
    }
    return 0;
}

int wqo_align_pe(wqo_handle* h, const wq_align_param* p, const wq_read_batch* f, const wq_read_batch* rv) {
    Result& R = h->R;
    R = Result();
    std::vector<Align> fr, rr;
    for (uint32_t i = 0; i < f->n_reads; i++) {
        Read a = load_read(f, i), b = load_read(rv, i);
        uint8_t fl;
        bool ok = ungapped_align_pe(h->L, *p, a, b, fr, rr, fl);
        R.hit_start.push_back((uint32_t)R.aln.size());
        R.nhits.push_back(ok ? (uint8_t)fr.size() : 0);
        R.flipped.push_back(fl);
        if (ok) for (size_t k = 0; k < fr.size(); k++) {
            emit(R, fr[k], R.aln);
            emit(R, rr[k], R.aln_mate);
        }
    }
    return 0;
}

uint64_t wqo_result_n_aln(wqo_handle* h) { return h->R.aln.size(); }
uint64_t wqo_result_n_nodes(wqo_handle* h) { return h->R.nodes.size(); }
void wqo_result_copy(wqo_handle* h, uint32_t* hit_start, uint8_t* nhits, uint8_t* flipped,
                     wq_alignment* aln, wq_alignment* aln_mate, wq_align_node* nodes) {
    Result& R = h->R;
    if (hit_start) memcpy(hit_start, R.hit_start.data(), R.hit_start.size() * 4);
    if (nhits) memcpy(nhits, R.nhits.data(), R.nhits.size());
    if (flipped) memcpy(flipped, R.flipped.data(), R.flipped.size());
    if (aln) memcpy(aln, R.aln.data(), R.aln.size() * sizeof(wq_alignment));
    if (aln_mate && !R.aln_mate.empty()) memcpy(aln_mate, R.aln_mate.data(), R.aln_mate.size() * sizeof(wq_alignment));
    if (nodes) memcpy(nodes, R.nodes.data(), R.nodes.size() * sizeof(wq_align_node));
}

}  // extern "C"
