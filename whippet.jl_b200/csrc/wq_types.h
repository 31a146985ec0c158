// Device-side data layout of the splice-graph index and of the per-batch work buffers.
// See DESIGN.md "Data layout in HBM".
#pragma once
#include <stdint.h>
#include "../../include/whippet_b200.h"

#if defined(__CUDACC__)
#define WQ_HD __host__ __device__ __forceinline__
#define WQ_HDN __host__ __device__
#else
#define WQ_HD inline
#define WQ_HDN
struct uint2 { uint32_t x, y; };
#endif

// EdgeType codes (reference: src/graph.jl:39-46)
enum : uint8_t { WQ_SL = 0, WQ_SR = 1, WQ_LS = 2, WQ_RS = 3, WQ_LR = 4, WQ_LL = 5, WQ_RR = 6 };

#define WQ_MAX_READ_WORDS 8        // 255 bases / 32 per u64
#define WQ_MAX_DEPTH 8             // nesting bound of spliced extensions (junction jumps found by table lookup)
#define WQ_NODE_BUCKET_SHIFT 6     // node lookup table: one entry per 64 text positions

// One occurrence block = 64 BWT symbols = one 32-byte sector:
//   cnt[c]  occurrences of c before the block, lo/hi = bit planes of the 64 two-bit symbols.
struct __attribute__((aligned(32))) WqOccBlock {
    uint32_t cnt[4];
    uint64_t lo;
    uint64_t hi;
};

// One splice-graph node = one 32-byte sector.  Index in the array = node_base[gene-1] + node-1,
// so the neighbours inside a gene are at +-1.
struct __attribute__((aligned(32))) WqNodeRec {
    uint32_t abs_start;    // 0-based text position of the first base (gene_offset + nodeoffset - 1)
    uint32_t len;          // nodelen
    uint32_t gene;         // 1-based gene id
    uint32_t node;         // 1-based node id inside the gene
    uint32_t kleft;        // edgeleft[node]  (SGKmer bits; meaningful where is_edge(edgetype[node], true))
    uint32_t kright;       // edgeright[node]
    uint8_t  et_left;      // edgetype[node]
    uint8_t  et_right;     // edgetype[node+1]
    uint16_t _pad;
    uint32_t _pad2;
};

struct __attribute__((aligned(16))) WqGeneRec {
    uint32_t text_start;   // GraphLib.offset[gene] (0-based)
    uint32_t text_end;     // exclusive
    uint32_t node_base;    // global index of node 1
    uint32_t nn;           // number of nodes
};

struct WqDevIndex {
    const WqOccBlock* occ;
    const uint32_t*   sa;
    const uint64_t*   text2;       // 32 bases per word, base j at bits 2*(j%32)
    const uint32_t*   amb;         // 1 bit per base, 32 per word; NULL when the text is pure ACGT
    const WqNodeRec*  nodes;
    const WqGeneRec*  genes;
    const uint32_t*   node_bucket; // [n>>6 + 2] number of nodes with abs_start <= (b<<6)
    const uint32_t*   left_ptr;    // [4^K+1]
    const uint2*      left_ent;    // (gene, global node index)
    const uint32_t*   right_ptr;
    const uint2*      right_ent;
    const uint2*      ftab;        // [4^ftab_k] rows (sp, ep) of every ftab_k-mer (sp > ep: absent); NULL = no table
    int32_t           ftab_k;
    uint64_t n;
    uint64_t sentinel;
    uint32_t C[4];
    uint32_t N, G;
    int32_t  K;
};

// Alignment path node of the working alignment (8 bytes).  A gene has at most 65535 nodes (ExonInt = UInt16,
// src/types.jl:20; enforced when the index is created), so the node id fits 16 bits.
struct __attribute__((aligned(8))) WqPNode {
    uint16_t node;     // 1-based node id inside the gene
    uint8_t  m;        // matches
    uint8_t  mm;       // mismatches
    float    tol;      // SGAlignScore.mistolerance
};

#define WQ_FAST_PATH 12            // path capacity of the first-pass extension kernel (shared memory); longer paths and
                                   // extensions that must park a candidate go to the second-pass kernel (WQ_MAX_PATH)

// Per-read seed result (seed_locate, src/align.jl:145-185)
struct WqSeed {
    uint32_t sp;        // first SA row (1-based, (n+1)-row space)
    uint32_t hit_base;  // first slot in the hit arrays
    uint8_t  cnt;       // rows (0 = no seed)
    uint8_t  readloc;
    uint8_t  strand;    // 1 = forward
    uint8_t  _pad;
};

// Per-hit extension result header
struct WqHit {
    float    identity;    // identity(align, readlen), Float32
    uint32_t offset;
    uint32_t path_start;  // into the path pool
    uint8_t  offsetread;
    uint8_t  npath;
    uint8_t  flags;       // bit0 align.isvalid, bit1 isvalid(align) (src/align.jl:115), bit2 kept by the score filter, bit3 overflow
    uint8_t  strand;
};
