// Host-side quantification state: the count stores downloaded from the device in canonical
// (sorted) order, the annotated isoform paths, and the host halves of the quant stage.
#pragma once
#include <map>
#include <string>
#include <vector>
#include "wq_types.h"
#include "wq_kernels.cuh"

struct WqHostQuant {
    std::vector<uint64_t> node;                              // SpliceGraphQuant.node, by global node index
    std::map<uint64_t, double> edge;                         // .edge and .circ: gene<<32 | l<<16 | r
    std::map<std::vector<uint32_t>, uint64_t> keyed;         // .long and MultiMapping.map records (wq_build_key_*)
    WqCounters stats;
    WqHostQuant() { memset(&stats, 0, sizeof stats); }
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
    std::map<uint64_t, double> edge_value;                   // edge + circ stores
    std::vector<double> iso_count, tpm, genetpm;
    WqClassSet cls;
    std::vector<const std::vector<uint32_t>*> multi_keys;    // key of each multi-map class (points into WqHostQuant.keyed)
    int64_t iterations = 0;
};

static inline bool wq_edge_is_circ(uint64_t key) { return ((key >> 16) & 0xFFFF) >= (key & 0xFFFF); }
static inline bool wq_key_is_multi(const std::vector<uint32_t>& k) { return !k.empty() && ((k[0] >> 28) & 1); }

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
