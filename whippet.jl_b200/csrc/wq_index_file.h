// Native on-disk form of the index as the kernels read it.
//
// The reference stores `GraphLib` with Julia's Serialization (`serialize(io, graphome)`, bin/whippet-index.jl:97-101) and
// pays a full deserialisation on every start (`@timer lib = open(deserialize, indexname)`, bin/whippet-quant.jl:105; the
// FM index is an `FMIndex{2,UInt32}` whose wavelet-matrix ranks the kernels do not use, src/index.jl:57,97).  This file
// holds the RE-LAID index instead -- 32-byte occurrence blocks, the full suffix array, 2-bit text + ambiguity mask,
// 32-byte node records, gene records, the node bucket table, the k-mer edge tables as CSR -- plus the flat annotation
// arrays the host stages need (isoform paths, node coordinates for SAM, lib.info), each section page-aligned, so loading
// is: map the file, copy the sections to the device.  wq_index_write builds the sections with the very function
// wq_index_create uses (wq_build_host_index), so both ways of making a handle upload the same bytes.
//
// Layout (little endian): WqIdxHeader, section table (WqIdxSection x n_sections), then the sections at 4096-byte aligned
// offsets.  Every section carries the sum of its 64-bit words (tail bytes zero-extended) as a checksum.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "wq_host_index.h"
#include "wq_host_quant.h"

#define WQ_IDX_MAGIC "WQXIDX\x01\x00"
#define WQ_IDX_VERSION 1u

enum WqIdxSec : uint32_t {
    WQ_SEC_OCC = 1, WQ_SEC_SA, WQ_SEC_TEXT2, WQ_SEC_AMB, WQ_SEC_NODES, WQ_SEC_GENES, WQ_SEC_BUCKET,
    WQ_SEC_LPTR, WQ_SEC_RPTR, WQ_SEC_LENT, WQ_SEC_RENT,
    WQ_SEC_NODE_BASE, WQ_SEC_NODEOFFSET, WQ_SEC_NODECOORD, WQ_SEC_NODELEN,
    WQ_SEC_ISO_BASE, WQ_SEC_ISO_NODE_PTR, WQ_SEC_ISO_NODES,
    WQ_SEC_NAME_OFF, WQ_SEC_NAMES, WQ_SEC_STRAND,
    WQ_SEC_END_
};

struct WqIdxHeader {
    char     magic[8];
    uint32_t version, header_bytes;      // header_bytes = sizeof(WqIdxHeader) + n_sections * sizeof(WqIdxSection)
    uint64_t file_bytes;
    int32_t  K; uint32_t G, N, I;
    uint64_t n, sentinel, nslots;
    uint32_t C[4];
    uint32_t has_amb, has_info, n_sections, rec_sizes;   // rec_sizes: sizeof(WqOccBlock) | sizeof(WqNodeRec) << 8 | sizeof(WqGeneRec) << 16
    uint64_t reserved[4];
};
struct WqIdxSection { uint32_t id, elem_bytes; uint64_t offset, count, checksum; };

static inline uint64_t wq_idx_checksum(const void* p, uint64_t bytes) {
    const unsigned nt = bytes < (8u << 20) ? 1u : std::min(16u, wq_host_threads());
    std::vector<uint64_t> part(nt, 0);
    const uint64_t words = bytes / 8;
    auto work = [&](unsigned t) {
        const uint64_t lo = words * t / nt, hi = words * (t + 1) / nt;
        uint64_t s = 0, w;
        const unsigned char* b = (const unsigned char*)p;
        for (uint64_t i = lo; i < hi; i++) { memcpy(&w, b + i * 8, 8); s += w; }
        part[t] = s;
    };
    if (nt == 1) work(0);
    else { std::vector<std::thread> th; for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t); for (auto& x : th) x.join(); }
    uint64_t s = 0;
    for (uint64_t v : part) s += v;
    if (bytes & 7) { uint64_t w = 0; memcpy(&w, (const unsigned char*)p + words * 8, bytes & 7); s += w; }
    return s;
}

// ---- writer (host only: no device is touched) ----
static inline std::string wq_index_write_impl(const wq_index_desc* d, const char* const* seqname, const uint8_t* strand, const char* path) {
    WqHostIndex H;
    std::string err = wq_build_host_index(d, H);
    if (!err.empty()) return err;
    {   // the annotation must be consistent before it is frozen into a file
        WqIsoIndex S;
        err = wq_build_iso_index(d, S);
        if (!err.empty()) return err;
    }
    struct Sec { uint32_t id, elem; const void* p; uint64_t count; };
    std::vector<Sec> secs;
    auto add = [&](uint32_t id, uint32_t elem, const void* p, uint64_t count) { secs.push_back(Sec{id, elem, p, count}); };
    const uint32_t G = d->n_genes, N = d->n_nodes, I = d->n_isoforms;
    add(WQ_SEC_OCC, sizeof(WqOccBlock), H.occ.data(), H.occ.size());
    add(WQ_SEC_SA, 4, d->sa, H.n);
    add(WQ_SEC_TEXT2, 8, H.text2.data(), H.text2.size());
    if (H.has_amb) add(WQ_SEC_AMB, 4, H.amb.data(), H.amb.size());
    add(WQ_SEC_NODES, sizeof(WqNodeRec), H.nodes.data(), H.nodes.size());
    add(WQ_SEC_GENES, sizeof(WqGeneRec), H.genes.data(), H.genes.size());
    add(WQ_SEC_BUCKET, 4, H.node_bucket.data(), H.node_bucket.size());
    add(WQ_SEC_LPTR, 4, d->left_ptr, H.nslots + 1);
    add(WQ_SEC_RPTR, 4, d->right_ptr, H.nslots + 1);
    add(WQ_SEC_LENT, sizeof(uint2), H.left_ent.data(), H.left_ent.size());
    add(WQ_SEC_RENT, sizeof(uint2), H.right_ent.data(), H.right_ent.size());
    add(WQ_SEC_NODE_BASE, 4, d->node_base, (uint64_t)G + 1);
    add(WQ_SEC_NODEOFFSET, 4, d->nodeoffset, N);
    add(WQ_SEC_NODECOORD, 4, d->nodecoord, N);
    add(WQ_SEC_NODELEN, 4, d->nodelen, N);
    add(WQ_SEC_ISO_BASE, 4, d->iso_base, (uint64_t)G + 1);
    add(WQ_SEC_ISO_NODE_PTR, 4, d->iso_node_ptr, (uint64_t)I + 1);
    add(WQ_SEC_ISO_NODES, 4, d->iso_nodes, d->iso_node_ptr[I]);
    std::vector<uint64_t> name_off;
    std::string names;
    std::vector<uint8_t> strands;
    const bool info = seqname && strand;
    if (info) {
        name_off.resize((size_t)G + 1);
        for (uint32_t g = 0; g < G; g++) {
            if (!seqname[g]) return "null sequence name";
            name_off[g] = names.size();
            names += seqname[g];
        }
        name_off[G] = names.size();
        strands.assign(strand, strand + G);
        for (auto& x : strands) x = x ? 1 : 0;
        add(WQ_SEC_NAME_OFF, 8, name_off.data(), name_off.size());
        add(WQ_SEC_NAMES, 1, names.data(), names.size());
        add(WQ_SEC_STRAND, 1, strands.data(), strands.size());
    }
    WqIdxHeader hd;
    memset(&hd, 0, sizeof hd);
    memcpy(hd.magic, WQ_IDX_MAGIC, 8);
    hd.version = WQ_IDX_VERSION;
    hd.header_bytes = (uint32_t)(sizeof(WqIdxHeader) + secs.size() * sizeof(WqIdxSection));
    hd.K = H.K; hd.G = G; hd.N = N; hd.I = I; hd.n = H.n; hd.sentinel = H.sentinel; hd.nslots = H.nslots;
    for (int i = 0; i < 4; i++) hd.C[i] = H.C[i];
    hd.has_amb = H.has_amb ? 1 : 0; hd.has_info = info ? 1 : 0; hd.n_sections = (uint32_t)secs.size();
    hd.rec_sizes = (uint32_t)sizeof(WqOccBlock) | ((uint32_t)sizeof(WqNodeRec) << 8) | ((uint32_t)sizeof(WqGeneRec) << 16);
    std::vector<WqIdxSection> table(secs.size());
    uint64_t off = (hd.header_bytes + 4095ull) & ~4095ull;
    for (size_t i = 0; i < secs.size(); i++) {
        const uint64_t bytes = secs[i].count * secs[i].elem;
        table[i] = WqIdxSection{secs[i].id, secs[i].elem, off, secs[i].count, wq_idx_checksum(secs[i].p, bytes)};
        off = (off + bytes + 4095ull) & ~4095ull;
    }
    hd.file_bytes = off;
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return std::string("cannot create ") + tmp;
    bool ok = fwrite(&hd, sizeof hd, 1, f) == 1 && (table.empty() || fwrite(table.data(), sizeof(WqIdxSection), table.size(), f) == table.size());
    static const char zeros[4096] = {0};
    uint64_t pos = hd.header_bytes;
    for (size_t i = 0; i < secs.size() && ok; i++) {
        if (table[i].offset > pos) { ok = fwrite(zeros, 1, (size_t)(table[i].offset - pos), f) == (size_t)(table[i].offset - pos); pos = table[i].offset; }
        const uint64_t bytes = secs[i].count * secs[i].elem;
        if (ok && bytes) ok = fwrite(secs[i].p, 1, (size_t)bytes, f) == (size_t)bytes;
        pos += bytes;
    }
    if (ok && hd.file_bytes > pos) ok = fwrite(zeros, 1, (size_t)(hd.file_bytes - pos), f) == (size_t)(hd.file_bytes - pos);
    ok = (fclose(f) == 0) && ok;
    if (!ok) { remove(tmp.c_str()); return std::string("write error on ") + tmp; }
    if (rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return std::string("cannot rename to ") + path; }
    return "";
}

// ---- reader: the file mapped, every section located and bounds-checked ----
struct WqIdxMapped {
    const unsigned char* base = nullptr; size_t len = 0; int fd = -1;
    WqIdxHeader hd;
    const WqIdxSection* sec[WQ_SEC_END_] = {nullptr};
    ~WqIdxMapped() { if (base) munmap((void*)base, len); if (fd >= 0) close(fd); }
    template <class T> const T* ptr(uint32_t id) const { return sec[id] ? reinterpret_cast<const T*>(base + sec[id]->offset) : nullptr; }
    uint64_t count(uint32_t id) const { return sec[id] ? sec[id]->count : 0; }
};

static inline std::string wq_index_map(const char* path, WqIdxMapped& M, bool verify) {
    M.fd = open(path, O_RDONLY);
    if (M.fd < 0) return std::string("cannot open ") + path;
    struct stat st;
    if (fstat(M.fd, &st) != 0) return std::string("cannot stat ") + path;
    if ((size_t)st.st_size < sizeof(WqIdxHeader)) return std::string(path) + ": not a whippet_b200 index file (too short)";
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, M.fd, 0);
    if (m == MAP_FAILED) return std::string("cannot map ") + path;
    M.base = (const unsigned char*)m; M.len = (size_t)st.st_size;
    madvise(m, M.len, MADV_SEQUENTIAL);
    memcpy(&M.hd, M.base, sizeof M.hd);
    const WqIdxHeader& hd = M.hd;
    if (memcmp(hd.magic, WQ_IDX_MAGIC, 8) != 0) return std::string(path) + ": not a whippet_b200 index file (bad magic)";
    if (hd.version != WQ_IDX_VERSION) return std::string(path) + ": index file version " + std::to_string(hd.version) + ", this library reads version " + std::to_string(WQ_IDX_VERSION);
    if (hd.rec_sizes != ((uint32_t)sizeof(WqOccBlock) | ((uint32_t)sizeof(WqNodeRec) << 8) | ((uint32_t)sizeof(WqGeneRec) << 16)))
        return std::string(path) + ": record sizes differ from this build";
    if (hd.file_bytes != (uint64_t)M.len) return std::string(path) + ": truncated index file (" + std::to_string(M.len) + " of " + std::to_string(hd.file_bytes) + " bytes)";
    if (hd.n_sections > 64 || hd.header_bytes != sizeof(WqIdxHeader) + hd.n_sections * sizeof(WqIdxSection) || hd.header_bytes > M.len)
        return std::string(path) + ": malformed section table";
    const WqIdxSection* tab = reinterpret_cast<const WqIdxSection*>(M.base + sizeof(WqIdxHeader));
    for (uint32_t i = 0; i < hd.n_sections; i++) {
        const WqIdxSection& s = tab[i];
        if (s.id == 0 || s.id >= WQ_SEC_END_ || M.sec[s.id]) return std::string(path) + ": unknown or repeated section";
        if ((s.offset & 4095) || s.offset > M.len || s.elem_bytes == 0 || s.count > (M.len - s.offset) / s.elem_bytes) return std::string(path) + ": section outside the file";
        M.sec[s.id] = &s;
    }
    auto need = [&](uint32_t id, uint32_t elem, uint64_t count) -> bool { return M.sec[id] && M.sec[id]->elem_bytes == elem && M.sec[id]->count == count; };
    const uint64_t n = hd.n, G = hd.G, N = hd.N, I = hd.I;
    if (hd.K < 1 || hd.K > 13 || hd.nslots != (1ull << (2 * hd.K)) || n == 0 || G == 0 || N == 0) return std::string(path) + ": bad header values";
    bool ok = need(WQ_SEC_OCC, sizeof(WqOccBlock), n / 64 + 1) && need(WQ_SEC_SA, 4, n) && need(WQ_SEC_TEXT2, 8, n / 32 + 3) &&
              (!hd.has_amb || need(WQ_SEC_AMB, 4, n / 32 + 3)) && need(WQ_SEC_NODES, sizeof(WqNodeRec), N) && need(WQ_SEC_GENES, sizeof(WqGeneRec), G) &&
              need(WQ_SEC_BUCKET, 4, (n >> WQ_NODE_BUCKET_SHIFT) + 2) && need(WQ_SEC_LPTR, 4, hd.nslots + 1) && need(WQ_SEC_RPTR, 4, hd.nslots + 1) &&
              M.sec[WQ_SEC_LENT] && M.sec[WQ_SEC_RENT] && need(WQ_SEC_NODE_BASE, 4, G + 1) && need(WQ_SEC_NODEOFFSET, 4, N) &&
              need(WQ_SEC_NODECOORD, 4, N) && need(WQ_SEC_NODELEN, 4, N) && need(WQ_SEC_ISO_BASE, 4, G + 1) && need(WQ_SEC_ISO_NODE_PTR, 4, I + 1) &&
              M.sec[WQ_SEC_ISO_NODES] && (!hd.has_info || (need(WQ_SEC_NAME_OFF, 8, G + 1) && M.sec[WQ_SEC_NAMES] && need(WQ_SEC_STRAND, 1, G)));
    if (!ok) return std::string(path) + ": a section is missing or has the wrong size";
    {   // cross-section consistency of what indexes into what
        const uint32_t* lp = M.ptr<uint32_t>(WQ_SEC_LPTR); const uint32_t* rp = M.ptr<uint32_t>(WQ_SEC_RPTR);
        const uint64_t lt = lp[hd.nslots], rt = rp[hd.nslots];
        if (M.count(WQ_SEC_LENT) != (lt ? lt : 1) || M.count(WQ_SEC_RENT) != (rt ? rt : 1)) return std::string(path) + ": edge table sizes disagree";
        if (M.ptr<uint32_t>(WQ_SEC_ISO_NODE_PTR)[I] != M.count(WQ_SEC_ISO_NODES)) return std::string(path) + ": isoform path sizes disagree";
        if (hd.has_info && M.ptr<uint64_t>(WQ_SEC_NAME_OFF)[G] != M.count(WQ_SEC_NAMES)) return std::string(path) + ": name table sizes disagree";
    }
    for (uint32_t i = 0; i < hd.n_sections; i++) {
        const WqIdxSection& s = tab[i];
        const uint64_t bytes = s.count * s.elem_bytes;
        if (!verify && bytes > (1u << 20)) continue;             // the large sections only on request (WQ_INDEX_VERIFY=1)
        if (wq_idx_checksum(M.base + s.offset, bytes) != s.checksum) return std::string(path) + ": checksum mismatch in section " + std::to_string(s.id);
    }
    return "";
}
