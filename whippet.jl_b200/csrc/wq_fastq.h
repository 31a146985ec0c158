// FASTQ ingest: inflate (zlib), 4-line record split, and packing into wq_read_batch on host threads.
//
// Follows /root/reference/src/reads.jl:2-23 and src/record.jl:13-19:
//   make_fqparser            FASTQ.Reader(stream, fill_ambiguous = DNA_A): every sequence byte that is not one of
//                            ACGTacgt becomes 'A' before the 2-bit conversion (FASTX 1.1.3 generate_fill_ambiguous)
//   fill!(rec, offset)       sequence -> 2-bit BioSequence, quality = raw quality bytes .- offset (UInt8 arithmetic)
// The reference parses one record at a time on one thread through Libz; here the serial part is only the inflate and
// the newline scan, the conversion of the records of a batch is spread over the host threads, and the batch is written
// into page-locked buffers so the H2D copies of wq_align_se / wq_align_pe run at full PCIe rate.
// Host-only code (no kernel): it is the stage in front of the device path, SURVEY.md section 8f row 1.
#pragma once
#include <zlib.h>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "wq_host_quant.h"          // wq_host_threads
#include "../../include/whippet_b200.h"

struct WqPinned {                    // grow-only page-locked buffer (plain memory when no CUDA device is usable)
    void* p = nullptr; size_t cap = 0; bool pinned = false;
    void* ensure(size_t bytes) {
        if (bytes <= cap) return p;
        release();
        size_t want = bytes + bytes / 4 + 4096;
        void* q = nullptr;
        if (cudaHostAlloc(&q, want, cudaHostAllocDefault) == cudaSuccess) { p = q; pinned = true; }
        else { (void)cudaGetLastError(); p = aligned_alloc(4096, (want + 4095) & ~(size_t)4095); pinned = false; }
        cap = p ? want : 0;
        return p;
    }
    void release() { if (p) { if (pinned) cudaFreeHost(p); else free(p); } p = nullptr; cap = 0; }
};

struct WqFastqSlot {                 // one packed batch + the record names (for SAM output)
    WqPinned len, seq_off, seq2, qual_off, qual;
    std::vector<uint64_t> name_off; std::vector<uint32_t> name_len; std::vector<char> names;
    uint32_t n = 0; uint64_t seq2_words = 0, qual_bytes = 0;
    void release() { len.release(); seq_off.release(); seq2.release(); qual_off.release(); qual.release(); }
};

struct wq_fastq {
    gzFile gz = nullptr;                          // file input (gzread is transparent for uncompressed files)
    const char* mem = nullptr; uint64_t mem_len = 0, mem_pos = 0;     // or an in-memory text
    int qualoffset = 33;
    std::vector<char> text;                       // inflated text not yet consumed
    size_t text_pos = 0;
    bool eof = false;
    WqFastqSlot slot[2];
    int cur = 0;
    uint64_t records = 0;                         // records returned so far (for error messages)
    std::string err;
};

static inline const uint8_t* wq_fastq_code_table() {
    static uint8_t T[256];
    static bool init = false;
    if (!init) {
        memset(T, 0, sizeof T);                   // everything that is not ACGTacgt -> 'A' = 0 (fill_ambiguous = DNA_A)
        T[(uint8_t)'C'] = T[(uint8_t)'c'] = 1; T[(uint8_t)'G'] = T[(uint8_t)'g'] = 2; T[(uint8_t)'T'] = T[(uint8_t)'t'] = 3;
        init = true;
    }
    return T;
}

// pull more text until at least `want` bytes are buffered after text_pos or the input ends
static inline bool wq_fastq_fill(wq_fastq* f, size_t want, bool compact) {
    if (compact && f->text_pos > 0) {                                    // drop the consumed prefix (only between batches:
        f->text.erase(f->text.begin(), f->text.begin() + f->text_pos);   //  records of the current batch are offsets into text)
        f->text_pos = 0;
    }
    while (!f->eof && f->text.size() - f->text_pos < want) {
        const size_t chunk = 8u << 20;
        const size_t old = f->text.size();
        f->text.resize(old + chunk);
        size_t got = 0;
        if (f->gz) {
            int r = gzread(f->gz, f->text.data() + old, (unsigned)chunk);
            if (r < 0) { int en = 0; f->err = std::string("gzread: ") + gzerror(f->gz, &en); f->text.resize(old); return false; }
            got = (size_t)r;
        } else {
            got = (size_t)std::min<uint64_t>(chunk, f->mem_len - f->mem_pos);
            memcpy(f->text.data() + old, f->mem + f->mem_pos, got);
            f->mem_pos += got;
        }
        f->text.resize(old + got);
        if (got == 0) f->eof = true;
    }
    return true;
}

// Next batch of at most max_reads records.  out->n_reads == 0 at end of input.  The buffers belong to the reader and
// stay valid until the call after the next one (two slots alternate, so a batch can be aligned while the next is parsed).
static inline int wq_fastq_next_impl(wq_fastq* f, uint32_t max_reads, wq_read_batch* out) {
    struct Rec { size_t name, seq, qual; uint32_t name_len, len; };          // offsets into f->text (it may be reallocated)
    std::vector<Rec> recs;
    recs.reserve(max_reads);
    // ---- serial: split lines; more text is appended as needed ----
    if (!wq_fastq_fill(f, 1u << 20, true)) return -20;
    while (recs.size() < max_reads) {
        const char* base = f->text.data();
        const char* end = base + f->text.size();
        const char* p = base + f->text_pos;
        // skip blank lines between records (a trailing newline at the end of the file)
        while (p < end && (*p == '\n' || *p == '\r')) p++;
        f->text_pos = (size_t)(p - base);
        if (p >= end) {
            if (f->eof) break;
            if (!wq_fastq_fill(f, 1u << 20, false)) return -20;
            continue;
        }
        const char* ln[5];
        ln[0] = p;
        bool complete = true, open_end = false;      // open_end: the last line of the file has no newline (ln[4] = end, nothing to strip)
        for (int k = 1; k <= 4; k++) {
            const char* nl = (const char*)memchr(ln[k - 1], '\n', (size_t)(end - ln[k - 1]));
            if (!nl) {
                if (f->eof && k == 4 && ln[3] < end) { ln[4] = end; open_end = true; break; }
                complete = false; break;
            }
            ln[k] = nl + 1;
        }
        if (!complete) {
            if (f->eof) { f->err = "truncated FASTQ record after record " + std::to_string(f->records + recs.size()); return -21; }
            if (!wq_fastq_fill(f, (f->text.size() - f->text_pos) + (1u << 20), false)) return -20;
            continue;
        }
        auto line_len = [&](int k) { size_t n = (size_t)(ln[k + 1] - ln[k]) - ((k == 3 && open_end) ? 0 : 1); if (n && ln[k][n - 1] == '\r') n--; return n; };
        const size_t nl = line_len(0), sl = line_len(1), ql = line_len(3);
        if (ln[0][0] != '@' || ln[2][0] != '+') { f->err = "malformed FASTQ record " + std::to_string(f->records + recs.size() + 1) + " (expected '@' / '+' lines)"; return -21; }
        if (sl != ql) { f->err = "FASTQ record " + std::to_string(f->records + recs.size() + 1) + ": sequence and quality lengths differ"; return -21; }
        if (sl > 255) { f->err = "FASTQ record " + std::to_string(f->records + recs.size() + 1) + ": read longer than 255 (ReadLengthInt = UInt8, src/align.jl:74)"; return -21; }
        recs.push_back(Rec{(size_t)(ln[0] + 1 - base), (size_t)(ln[1] - base), (size_t)(ln[3] - base), (uint32_t)(nl - 1), (uint32_t)sl});
        f->text_pos = (size_t)(ln[4] - base);
    }
    const char* text = f->text.data();
    f->cur ^= 1;
    WqFastqSlot& S = f->slot[f->cur];
    const uint32_t n = (uint32_t)recs.size();
    S.n = n;
    memset(out, 0, sizeof *out);
    if (n == 0) return 0;
    // ---- offsets (serial prefix), then conversion on the host threads ----
    uint8_t* len = (uint8_t*)S.len.ensure(n);
    uint64_t* seq_off = (uint64_t*)S.seq_off.ensure((size_t)n * 8);
    uint64_t* qual_off = (uint64_t*)S.qual_off.ensure((size_t)n * 8);
    if (!len || !seq_off || !qual_off) { f->err = "out of memory"; return -22; }
    S.name_off.resize(n); S.name_len.resize(n);
    uint64_t w = 0, q = 0, nm = 0;
    for (uint32_t i = 0; i < n; i++) {
        len[i] = (uint8_t)recs[i].len;
        seq_off[i] = w; qual_off[i] = q; S.name_off[i] = nm; S.name_len[i] = recs[i].name_len;
        w += (recs[i].len + 31) / 32 + (recs[i].len == 0 ? 1 : 0);
        q += (recs[i].len + 3) & ~3u;
        nm += recs[i].name_len;
    }
    S.seq2_words = w; S.qual_bytes = q;
    uint64_t* seq2 = (uint64_t*)S.seq2.ensure((size_t)std::max<uint64_t>(w, 1) * 8);
    uint8_t* qual = (uint8_t*)S.qual.ensure((size_t)std::max<uint64_t>(q, 4));
    if (!seq2 || !qual) { f->err = "out of memory"; return -22; }
    S.names.resize(nm);
    const uint8_t* T = wq_fastq_code_table();
    const uint8_t qo = (uint8_t)f->qualoffset;
    unsigned nt = wq_host_threads();
    if (n < 4096) nt = 1;
    auto work = [&](unsigned t) {
        const uint32_t lo = (uint32_t)((uint64_t)n * t / nt), hi = (uint32_t)((uint64_t)n * (t + 1) / nt);
        for (uint32_t i = lo; i < hi; i++) {
            const Rec& r = recs[i];
            uint64_t* wd = seq2 + seq_off[i];
            const uint32_t nw = (r.len + 31) / 32;
            for (uint32_t k = 0; k < nw; k++) {
                uint64_t acc = 0;
                const uint32_t b0 = k * 32, b1 = std::min(r.len, b0 + 32);
                for (uint32_t j = b0; j < b1; j++) acc |= (uint64_t)T[(uint8_t)text[r.seq + j]] << (2 * (j - b0));
                wd[k] = acc;
            }
            if (r.len == 0) wd[0] = 0;
            uint8_t* qd = qual + qual_off[i];
            for (uint32_t j = 0; j < r.len; j++) qd[j] = (uint8_t)((uint8_t)text[r.qual + j] - qo);       // UInt8 arithmetic (src/record.jl:17)
            for (uint32_t j = r.len; j < ((r.len + 3) & ~3u); j++) qd[j] = 0;
            memcpy(S.names.data() + S.name_off[i], text + r.name, r.name_len);
        }
    };
    if (nt == 1) work(0);
    else { std::vector<std::thread> th; for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t); for (auto& x : th) x.join(); }
    f->records += n;
    out->n_reads = n; out->len = len; out->seq_off = seq_off; out->seq2 = seq2; out->seq2_words = w;
    out->qual_off = qual_off; out->qual = qual; out->qual_bytes = q;
    return 0;
}
