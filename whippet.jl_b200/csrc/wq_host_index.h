// Host-side re-layout of a flat GraphLib (wq_index_desc) into the sector-sized records the kernels
// read (WqOccBlock / WqNodeRec / WqGeneRec / bucket table / 2-bit text).  Plain C++: the CUDA
// library uploads the result, tests/emu uses it in place.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "wq_types.h"
#include "wq_align.cuh"

struct WqHostIndex {
    std::vector<WqOccBlock> occ;
    std::vector<uint64_t>   text2;
    std::vector<uint32_t>   amb;
    bool                    has_amb = false;
    std::vector<WqNodeRec>  nodes;
    std::vector<WqGeneRec>  genes;
    std::vector<uint32_t>   node_bucket;
    std::vector<uint2>      left_ent, right_ent;
    uint64_t n = 0, sentinel = 0, nslots = 0;
    uint32_t C[4] = {0, 0, 0, 0};
    uint32_t N = 0, G = 0;
    int32_t  K = 0;
};

static inline std::string wq_build_host_index(const wq_index_desc* d, WqHostIndex& H) {
    const uint64_t n = d->text_len;
    const uint32_t G = d->n_genes, N = d->n_nodes;
    if (d->kmer < 1 || d->kmer > 13) return "kmer size K must be in 1..13 (edge tables are dense 4^K arrays)";
    if (n == 0 || G == 0 || N == 0) return "empty index";
    if (n >= 0xFFFFFFF0ull) return "text longer than CoordInt (UInt32) allows";
    H.n = n; H.G = G; H.N = N; H.K = d->kmer; H.sentinel = d->sentinel;
    for (int i = 0; i < 4; i++) H.C[i] = (uint32_t)d->count[i];
    // --- occurrence blocks: 64 symbols per 32-byte sector ---
    const uint64_t nb = n / 64 + 1;
    H.occ.assign(nb, WqOccBlock{{0, 0, 0, 0}, 0, 0});
    uint32_t c[4] = {0, 0, 0, 0};
    for (uint64_t i = 0; i < n; i++) {
        uint64_t b = i >> 6;
        if ((i & 63) == 0) memcpy(H.occ[b].cnt, c, sizeof c);
        uint8_t s = d->bwt[i] & 3;
        if (s & 1) H.occ[b].lo |= 1ull << (i & 63);
        if (s & 2) H.occ[b].hi |= 1ull << (i & 63);
        c[s]++;
    }
    if ((n & 63) == 0) memcpy(H.occ[n >> 6].cnt, c, sizeof c);
    // --- 2-bit text + ambiguity mask ---
    H.text2.assign(n / 32 + 3, 0);
    H.amb.assign(n / 32 + 3, 0);
    for (uint64_t i = 0; i < n; i++) {
        uint8_t v = d->seq4[i];
        uint64_t code;
        switch (v) {
            case 1: code = 0; break;
            case 2: code = 1; break;
            case 4: code = 2; break;
            case 8: code = 3; break;
            default: code = 0; H.amb[i >> 5] |= 1u << (i & 31); H.has_amb = true;
        }
        H.text2[i >> 5] |= code << (2 * (i & 31));
    }
    // --- genes + nodes ---
    H.genes.resize(G);
    H.nodes.resize(N);
    for (uint32_t g = 0; g < G; g++) {
        WqGeneRec& gr = H.genes[g];
        gr.text_start = d->gene_offset[g];
        gr.text_end = g + 1 < G ? d->gene_offset[g + 1] : (uint32_t)n;
        gr.node_base = d->node_base[g];
        gr.nn = d->node_base[g + 1] - d->node_base[g];
        if (gr.nn == 0) return "gene without nodes";
        if (gr.nn > 65535) return "more than 65535 nodes in one gene (ExonInt = UInt16, src/types.jl:20)";
        for (uint32_t j = 0; j < gr.nn; j++) {
            uint32_t gi = gr.node_base + j;
            size_t e = (size_t)gr.node_base + g + j;          // edge j+1 (left of node j+1)
            WqNodeRec& nr = H.nodes[gi];
            nr.abs_start = gr.text_start + d->nodeoffset[gi] - 1;
            nr.len = d->nodelen[gi];
            nr.gene = g + 1;
            nr.node = j + 1;
            nr.kleft = (uint32_t)d->edgeleft[e];
            nr.kright = (uint32_t)d->edgeright[e];
            nr.et_left = d->edgetype[e];
            nr.et_right = d->edgetype[e + 1];
            nr._pad = 0; nr._pad2 = 0;
            if (gi > 0 && nr.abs_start < H.nodes[gi - 1].abs_start) return "node offsets are not non-decreasing";
        }
    }
    // --- bucket table: #nodes with abs_start <= b<<6 ---
    const uint64_t nbk = (n >> WQ_NODE_BUCKET_SHIFT) + 2;
    H.node_bucket.assign(nbk, 0);
    uint32_t k = 0;
    for (uint64_t b = 0; b < nbk; b++) {
        uint64_t pos = b << WQ_NODE_BUCKET_SHIFT;
        while (k < N && (uint64_t)H.nodes[k].abs_start <= pos) k++;
        H.node_bucket[b] = k;
    }
    // --- k-mer edge tables: (gene, global node index) ---
    H.nslots = 1ull << (2 * d->kmer);
    auto conv = [&](const uint32_t* ptr, const uint32_t* ent, std::vector<uint2>& out) -> bool {
        uint32_t tot = ptr[H.nslots];
        out.resize(tot ? tot : 1);
        for (uint32_t i = 0; i < tot; i++) {
            uint32_t g = ent[2 * i], nd = ent[2 * i + 1];
            if (g < 1 || g > G || nd < 1 || nd > H.genes[g - 1].nn) return false;
            out[i].x = g;
            out[i].y = H.genes[g - 1].node_base + nd - 1;
        }
        return true;
    };
    if (!conv(d->left_ptr, d->left_nodes, H.left_ent)) return "edge table entry out of range (left)";
    if (!conv(d->right_ptr, d->right_nodes, H.right_ent)) return "edge table entry out of range (right)";
    return "";
}

// AlignParam -> kernel parameter block.  phred_to_prob (src/align.jl:124) = Float32(1 - 10^(-phred/10)).
static inline std::string wq_make_params(const wq_align_param* a, int index_kmer, WqParams& p) {
    if (a->seed_length < 1 || a->seed_length > 32) return "seed_length must be in 1..32";
    if (a->seed_buffer < 1) return "seed_buffer must be >= 1 (the reference indexes quality[seed_buffer:...], src/align.jl:160)";
    if (a->seed_tolerate < 1 || a->seed_tolerate > WQ_MAX_TOLERATE) return "seed_tolerate must be in 1..8";
    if (a->seed_inc < 1) return "seed_inc must be >= 1";
    if (a->kmer_size != index_kmer) return "AlignParam.kmer_size must equal GraphLib.kmer (bin/whippet-quant.jl:124)";
    if (a->mismatches < 0) return "mismatches must be >= 0";
    p.mismatches = a->mismatches; p.K = a->kmer_size; p.seed_try = a->seed_try; p.seed_tolerate = a->seed_tolerate;
    p.seed_min_qual = a->seed_min_qual; p.seed_length = a->seed_length; p.seed_buffer = a->seed_buffer;
    p.seed_inc = a->seed_inc; p.seed_pair_range = a->seed_pair_range;
    p.mmlimit = (float)a->mismatches;
    p.kf = (float)a->kmer_size;
    p.score_range = a->score_range; p.score_min = a->score_min;
    p.is_stranded = a->is_stranded; p.is_pair_rc = a->is_pair_rc;
    for (int i = 0; i < 128; i++) p.phred_prob[i] = (float)(1.0 - std::pow(10.0, -(double)i / 10.0));
    return "";
}
