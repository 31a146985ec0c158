// SAM emission from the alignments the device produced.
//
// Follows /root/reference/src/io.jl:69-271: seq_write / qual_write (:69-84), sam_flag (:86-105), soft_pad (:107-110),
// cigar_string (:111-147), write_sam for one alignment (:151-190), indmax_score (:193-217), write_sam for a vector of
// alignments (:221-236), for paired vectors (:239-258) and write_sam_header (:261-276); the callers are
// process_reads! / process_paired_reads! (src/reads.jl:63-67, 110-121).  Text formatting is host work: the reads of a
// batch are split over the host threads and the pieces are concatenated in read order.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>
#include "wq_host_quant.h"
#include "../../include/whippet_b200.h"

struct WqSamIndex {                       // the parts of GraphLib that only SAM output needs
    std::vector<uint32_t> node_base, nodeoffset, nodecoord, nodelen;      // desc arrays (1-based node i of gene g at node_base[g-1] + i-1)
    std::vector<std::string> seqname;     // lib.info[g].name
    std::vector<uint8_t> strand;          // lib.info[g].strand (1 = '+')
    bool has_info = false;
};

struct WqSamRead {                        // FASTQRecord as write_sam sees it (after ungapped_align's in-place reverse complement)
    const wq_read_batch* b; uint32_t i; bool flipped;
    uint32_t len() const { return b->len[i]; }
    // base j (0-based) of the current orientation, as a 2-bit code
    uint8_t code(uint32_t j, bool rc) const {
        const uint32_t L = len(), k = rc ? L - 1 - j : j;
        const uint64_t so = b->fixed_len ? (uint64_t)i * ((b->fixed_len + 31) / 32) : b->seq_off[i];
        const uint8_t c = (uint8_t)((b->seq2[so + (k >> 5)] >> (2 * (k & 31))) & 3);
        return rc ? (uint8_t)(3 - c) : c;
    }
    uint8_t qual(uint32_t j, bool rc) const {
        const uint32_t L = len();
        const uint32_t k = rc ? L - 1 - j : j;
        if (b->qual_bits == 4) {                               // 4-bit dictionary indices (fixed_len batches)
            const uint8_t v = b->qual[(uint64_t)i * ((((uint64_t)b->fixed_len + 1) / 2 + 3) & ~3ull) + (k >> 1)];
            return b->qual_dict[(k & 1) ? (v >> 4) : (v & 15)];
        }
        const uint64_t qo = b->fixed_len ? (uint64_t)i * ((b->fixed_len + 3) & ~3u) : b->qual_off[i];
        return b->qual[qo + k];
    }
};

static inline void wq_sam_uint(std::string& o, uint64_t v) { char t[24]; int n = snprintf(t, sizeof t, "%llu", (unsigned long long)v); o.append(t, (size_t)n); }

// score(align) = matches - mistolerance summed over the path in Float32 (src/align.jl:90-112)
static inline float wq_sam_score(const wq_alignment& a, const wq_align_node* nodes) {
    float s = 0.0f;
    for (uint32_t k = 0; k < a.npath; k++) { const wq_align_node& n = nodes[a.path_start + k]; s += (float)n.matches - n.mistolerance; }
    return s;
}

// cigar_string (src/io.jl:111-147); returns curpos
static inline uint32_t wq_sam_cigar(const WqSamIndex& X, const wq_alignment& a, const wq_align_node* nodes, uint32_t gene, bool strand,
                                    uint32_t readlen, std::string& cigar) {
    const uint32_t nb = X.node_base[gene - 1];
    auto coord = [&](uint32_t n) { return (int64_t)X.nodecoord[nb + n - 1]; };
    auto nlen = [&](uint32_t n) { return (int64_t)X.nodelen[nb + n - 1]; };
    uint32_t curpos = a.offset;
    uint32_t running = 0;
    int64_t total = (int64_t)a.offsetread - 1;
    std::vector<std::string> parts;
    { const int64_t pad = (int64_t)a.offsetread - 1; parts.push_back(pad > 0 ? std::to_string(pad) + "S" : std::string()); }
    for (uint32_t idx = 0; idx < a.npath; idx++) {
        const wq_align_node& nd = nodes[a.path_start + idx];
        const uint32_t matches = (uint8_t)(nd.matches + nd.mismatches);            // UInt8 + UInt8
        total += matches;
        curpos += matches;
        if (idx + 1 < a.npath) {
            const uint32_t curi = nd.node, nexti = nodes[a.path_start + idx + 1].node;
            const int64_t intron = strand ? coord(nexti) - (coord(curi) + nlen(curi) - 1) - 1
                                          : coord(curi) - (coord(nexti) + nlen(nexti) - 1) - 1;
            if (intron >= 1) {
                parts.push_back(std::to_string(running + matches) + "M");
                parts.push_back(std::to_string(intron) + "N");
                running = 0;
                curpos = X.nodeoffset[nb + nexti - 1];
            } else running += matches;
        } else parts.push_back(std::to_string(running + matches) + "M");
    }
    if (total < (int64_t)readlen) parts.push_back(std::to_string((int64_t)readlen - total) + "S");
    cigar.clear();
    if (strand) for (auto& p : parts) cigar += p;
    else for (size_t k = parts.size(); k-- > 0;) cigar += parts[k];
    return curpos;
}

// write_sam for one alignment (src/io.jl:151-190).  `rc_state` is the read's orientation relative to the FASTQ text
// (true after an odd number of reverse_complement! calls); a regular (non-supplementary) entry of a minus-strand gene
// reverse-complements the read in place, as the reference does.
static inline void wq_sam_entry(const WqSamIndex& X, std::string& o, const char* name, uint32_t name_len, const WqSamRead& rd, bool& rc_state,
                                const wq_alignment& a, const wq_align_node* nodes, int mapq, bool paired, bool fwd_mate, bool supplemental,
                                int qualoffset) {
    if (!(a.isvalid && a.npath >= 1)) return;
    const uint32_t gene = nodes[a.path_start].gene;
    const bool strand = X.strand[gene - 1] != 0;
    const uint32_t first = nodes[a.path_start].node, last = nodes[a.path_start + a.npath - 1].node;
    const uint32_t nodeind = strand ? first : last;
    if (last < nodeind || last < first) return;                                  // circular paths are not written
    std::string cigar;
    const uint32_t endpos = wq_sam_cigar(X, a, nodes, gene, strand, rd.len(), cigar);
    const uint32_t nb = X.node_base[gene - 1];
    const uint32_t noff = X.nodeoffset[nb + nodeind - 1];
    const uint32_t offset = strand ? (uint32_t)(a.offset - noff) : (uint32_t)(X.nodelen[nb + nodeind - 1] - (uint32_t)(endpos - noff));   // UInt32 arithmetic
    if (!strand && !supplemental) rc_state = !rc_state;
    uint32_t idlen = 0;                                                         // string(read): the identifier stops at the first blank
    while (idlen < name_len && name[idlen] != ' ' && name[idlen] != '\t') idlen++;
    o.append(name, idlen); o += '\t';
    uint32_t flag = 0;                                                          // sam_flag (src/io.jl:86-105)
    if (paired) { flag |= 0x01 | 0x02; flag |= fwd_mate ? 0x40 : 0x80; if (!strand) flag |= 0x10; }
    else if (!strand) flag ^= 0x10;
    if (supplemental) flag |= 0x100;
    wq_sam_uint(o, flag); o += '\t';
    o += X.seqname[gene - 1]; o += '\t';
    wq_sam_uint(o, (uint32_t)(X.nodecoord[nb + nodeind - 1] + offset)); o += '\t';
    wq_sam_uint(o, (uint64_t)mapq); o += '\t';
    o += cigar; o += "\t*\t0\t0\t";
    if (!supplemental) {
        const uint32_t L = rd.len();
        for (uint32_t j = 0; j < L; j++) o += "ACGT"[rd.code(j, rc_state)];
        o += '\t';
        for (uint32_t j = 0; j < L; j++) o += (char)(uint8_t)(rd.qual(j, rc_state) + qualoffset);
    } else o += "*\t*";
    o += '\n';
}

// process_reads! / process_paired_reads! with sam=true (src/reads.jl:59-67, 106-121) for reads lo..hi-1 of a batch
static inline void wq_sam_range(const WqSamIndex& X, std::string& o, const wq_read_batch* fwd, const wq_read_batch* rev,
                                const uint64_t* name_off, const uint32_t* name_len, const char* names,
                                const uint64_t* rname_off, const uint32_t* rname_len, const char* rnames,
                                const wq_align_out* out, int qualoffset, uint32_t lo, uint32_t hi) {
    const bool paired = rev != nullptr;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t nh = out->nhits[i];
        if (nh == 0) continue;
        const wq_alignment* av = out->aln + out->hit_start[i];
        const wq_alignment* mv = paired ? out->aln_mate + out->hit_start[i] : nullptr;
        const WqSamRead r1{fwd, i, (out->flipped[i] & 1) != 0};
        bool rc1 = r1.flipped;
        const char* n1 = names + name_off[i];
        if (!paired) {
            if (nh == 1) { wq_sam_entry(X, o, n1, name_len[i], r1, rc1, av[0], out->nodes, 0, false, true, false, qualoffset); continue; }
            uint32_t best = 0;                                                  // indmax_score (src/io.jl:193-204)
            float mx = wq_sam_score(av[0], out->nodes);
            for (uint32_t k = 1; k < nh; k++) { float c = wq_sam_score(av[k], out->nodes); if (c > mx) { mx = c; best = k; } }
            const int mapq = 10 * (int)nh;
            wq_sam_entry(X, o, n1, name_len[i], r1, rc1, av[best], out->nodes, mapq, false, true, false, qualoffset);
            for (uint32_t k = 0; k < nh; k++) if (k != best) wq_sam_entry(X, o, n1, name_len[i], r1, rc1, av[k], out->nodes, mapq, false, true, true, qualoffset);
            continue;
        }
        const WqSamRead r2{rev, i, (out->flipped[i] & 2) != 0};
        bool rc2 = r2.flipped;
        const char* n2 = rnames + rname_off[i];
        if (nh == 1) {
            wq_sam_entry(X, o, n1, name_len[i], r1, rc1, av[0], out->nodes, 0, true, true, false, qualoffset);
            wq_sam_entry(X, o, n2, rname_len[i], r2, rc2, mv[0], out->nodes, 0, true, false, false, qualoffset);
            continue;
        }
        uint32_t best = 0;                                                      // indmax_score for pairs (src/io.jl:206-217)
        float mx = wq_sam_score(av[0], out->nodes) + wq_sam_score(mv[0], out->nodes);
        for (uint32_t k = 1; k < nh; k++) { float c = wq_sam_score(av[k], out->nodes) + wq_sam_score(mv[k], out->nodes); if (c > mx) { mx = c; best = k; } }
        const int mapq = 10 * (int)nh;
        wq_sam_entry(X, o, n1, name_len[i], r1, rc1, av[best], out->nodes, mapq, true, true, false, qualoffset);
        wq_sam_entry(X, o, n2, rname_len[i], r2, rc2, mv[best], out->nodes, mapq, true, false, false, qualoffset);
        for (uint32_t k = 0; k < nh; k++) if (k != best) {
            // both supplementary entries are written with fwd_mate = true, as the reference does (src/io.jl:252-256)
            wq_sam_entry(X, o, n1, name_len[i], r1, rc1, av[k], out->nodes, mapq, true, true, true, qualoffset);
            wq_sam_entry(X, o, n2, rname_len[i], r2, rc2, mv[k], out->nodes, mapq, true, true, true, qualoffset);
        }
    }
}

// write_sam_header (src/io.jl:261-276).  The reference iterates a Dict (order unspecified); sequence names are written in
// ascending order here.
static inline std::string wq_sam_header_text(const WqSamIndex& X) {
    std::map<std::string, uint64_t> refs;
    const uint32_t G = (uint32_t)X.seqname.size();
    for (uint32_t g = 0; g < G; g++) {
        uint64_t len = 0;
        for (uint32_t k = X.node_base[g]; k < X.node_base[g + 1]; k++) len = std::max<uint64_t>(len, X.nodecoord[k]);
        auto it = refs.find(X.seqname[g]);
        if (it == refs.end() || it->second < len) refs[X.seqname[g]] = len + 10000;
    }
    std::string o = "@HD\tVN:1.0\tSO:unsorted\n";
    for (auto& kv : refs) { o += "@SQ\tSN:"; o += kv.first; o += "\tLN:"; wq_sam_uint(o, kv.second); o += '\n'; }
    return o;
}
