// FASTQ ingest: inflate (zlib), 4-line record split, and packing into wq_read_batch on host threads.
//
// Follows /root/reference/src/reads.jl:2-23 and src/record.jl:13-19:
//   make_fqparser            FASTQ.Reader(stream, fill_ambiguous = DNA_A): every sequence byte that is not one of
//                            ACGTacgt becomes 'A' before the 2-bit conversion (FASTX 1.1.3 generate_fill_ambiguous)
//   fill!(rec, offset)       sequence -> 2-bit BioSequence, quality = raw quality bytes .- offset (UInt8 arithmetic)
// The reference parses one record at a time on one thread through Libz.  Here every stage that can be is spread over the
// host threads, because at the device's alignment rate (> 100 M reads/s) the text side is what bounds file -> result:
//   * plain files are mapped, not read: the newline scan and the conversion work on the mapping directly;
//   * blocked gzip (BGZF: bgzip, bcl2fastq output -- independent deflate blocks that announce their size) is inflated block
//     by block on all threads straight into the text window (wq_inflate.h, no intermediate copy); an ordinary single-member gzip stream has no entry points and
//     is inflated serially (wq_inflate.h, about twice zlib's rate on FASTQ text), overlapped with the alignment of the previous batch;
//   * the newline scan runs in parallel over pieces of the window (records are strictly four lines, so a record boundary is
//     a line number, not a guess about '@'), then lengths / validation / offsets by a two-pass prefix sum, then the 2-bit and
//     phred conversion of the records;
//   * the batch is written into page-locked buffers so the H2D copies of wq_align_se / wq_align_pe run at full PCIe rate.
// Host-only code (no kernel): it is the stage in front of the device path, SURVEY.md section 8f row 1.
#pragma once
#include <chrono>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include <atomic>
#include <algorithm>
#include <condition_variable>
#include <emmintrin.h>
#include <mutex>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "wq_host_quant.h"          // wq_host_threads
#include "../../include/whippet_b200.h"
#include "wq_inflate.h"

// Page-locked blocks released by a reader are kept for the next one (process lifetime, bounded): cudaHostAlloc costs about
// as much as parsing the batch that goes into the block, and wq_process_fastq opens a reader per call.
struct WqPinnedCache {
    struct Blk { void* p; size_t cap; };
    std::mutex mu; std::vector<Blk> free_; size_t bytes = 0;
    static WqPinnedCache& get() { static WqPinnedCache* c = new WqPinnedCache(); return *c; }     // never destroyed (no CUDA calls at exit)
    void* take(size_t want, size_t* cap) {
        std::lock_guard<std::mutex> g(mu);
        size_t best = free_.size();
        for (size_t i = 0; i < free_.size(); i++)
            if (free_[i].cap >= want && (best == free_.size() || free_[i].cap < free_[best].cap)) best = i;
        if (best == free_.size() || free_[best].cap > 4 * want + (1u << 20)) return nullptr;
        void* p = free_[best].p; *cap = free_[best].cap; bytes -= *cap;
        free_.erase(free_.begin() + best);
        return p;
    }
    bool give(void* p, size_t cap) {
        std::lock_guard<std::mutex> g(mu);
        if (free_.size() >= 32 || bytes + cap > (size_t)4 << 30) return false;
        free_.push_back(Blk{p, cap}); bytes += cap;
        return true;
    }
};

struct WqPinned {                    // grow-only page-locked buffer (plain memory when no CUDA device is usable)
    void* p = nullptr; size_t cap = 0; bool pinned = false;
    void* ensure(size_t bytes) {
        if (bytes <= cap) return p;
        release();
        size_t want = bytes + bytes / 4 + 4096;
        void* q = WqPinnedCache::get().take(want, &cap);
        if (q) { p = q; pinned = true; return p; }
        if (cudaHostAlloc(&q, want, cudaHostAllocDefault) == cudaSuccess) { p = q; pinned = true; }
        else { (void)cudaGetLastError(); p = aligned_alloc(4096, (want + 4095) & ~(size_t)4095); pinned = false; }
        cap = p ? want : 0;
        return p;
    }
    void release() {
        if (p) {
            if (pinned) { if (!WqPinnedCache::get().give(p, cap)) cudaFreeHost(p); }
            else free(p);
        }
        p = nullptr; cap = 0;
    }
};

// text window of the gzip modes: raw storage that is never zero-filled (std::vector::resize would touch every byte first)
struct WqTextBuf {
    char* p = nullptr; size_t len = 0, cap = 0;
    ~WqTextBuf() { free(p); }
    size_t size() const { return len; }
    char* data() { return p; }
    bool reserve_more(size_t add) {
        if (len + add <= cap) return true;
        size_t nc = std::max(len + add, cap + cap / 2);
        char* q = (char*)realloc(p, nc);
        if (!q) return false;
        p = q; cap = nc;
        return true;
    }
    void drop_front(size_t n) { if (n >= len) { len = 0; return; } memmove(p, p + n, len - n); len -= n; }
};

struct WqFastqSlot {                 // one packed batch + the record names (for SAM output)
    WqPinned len, seq_off, seq2, qual_off, qual;
    std::vector<uint64_t> name_off; std::vector<uint32_t> name_len; std::vector<char> names;
    uint32_t n = 0; uint64_t seq2_words = 0, qual_bytes = 0;
    void release() { len.release(); seq_off.release(); seq2.release(); qual_off.release(); qual.release(); }
};

struct wq_fastq {
    // ---- input: exactly one of these ----
    const char* map = nullptr; size_t map_len = 0; int fd = -1; bool map_owned = false;     // plain text, mapped (or caller's memory)
    gzFile gz = nullptr;                                                                   // ordinary gzip through zlib's gzread (WQ_GZ_ZLIB=1, or the file cannot be mapped)
    WqInflate* inf = nullptr;                                                              // ordinary gzip: `map` holds the COMPRESSED file, inflated serially by wq_inflate.h
    bool bgzf = false; size_t bgzf_pos = 0;                                                 // blocked gzip: `map` holds the COMPRESSED file
    int qualoffset = 33;
    // ---- text window: stream offsets [woff, woff + wlen) are in memory at wbase ----
    WqTextBuf text;                    // owns the window for the gzip modes
    const char* wbase = nullptr; uint64_t woff = 0, wlen = 0;
    bool eof = false;                  // no more text will be appended to the window
    // ---- line index: stream offsets of the newlines found so far, nl_pos = first one not yet consumed by a batch ----
    std::vector<uint64_t> nl; size_t nl_pos = 0;
    uint64_t scanned = 0;              // the window has been scanned up to this stream offset
    uint64_t line_start = 0;           // stream offset of the first line not yet consumed
    bool closed_tail = false;          // the virtual newline after an unterminated last line has been added
    WqFastqSlot slot[2];
    int cur = 0;
    uint64_t records = 0;              // records returned so far (for error messages)
    double avg_rec = 300.0;            // bytes per record seen so far (sizes the next scan)
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

// 16 sequence bytes -> 16 two-bit codes in 32 bits (base j in bits 2j..2j+1).  For ACGTacgt the code is ((c >> 1) ^ (c >> 2)) & 3
// (A 0, C 1, G 2, T 3 in either case); every other byte becomes 0 = A (fill_ambiguous = DNA_A).
static inline uint32_t wq_fastq_pack16(const char* p) {
    const __m128i v = _mm_loadu_si128((const __m128i*)p);
    const __m128i u = _mm_and_si128(v, _mm_set1_epi8((char)0xDF));                    // fold the case bit
    const __m128i ok = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(u, _mm_set1_epi8('A')), _mm_cmpeq_epi8(u, _mm_set1_epi8('C'))),
                                    _mm_or_si128(_mm_cmpeq_epi8(u, _mm_set1_epi8('G')), _mm_cmpeq_epi8(u, _mm_set1_epi8('T'))));
    __m128i c = _mm_xor_si128(_mm_srli_epi16(v, 1), _mm_srli_epi16(v, 2));            // bits leaking across bytes are masked next
    c = _mm_and_si128(_mm_and_si128(c, _mm_set1_epi8(3)), ok);
    // fold pairs: 16 x 2 bits (one per byte) -> 8 x 4 bits -> 4 x 8 bits -> 2 x 16 bits per 64-bit half
    uint64_t lo = (uint64_t)_mm_cvtsi128_si64(c), hi = (uint64_t)_mm_cvtsi128_si64(_mm_unpackhi_epi64(c, c));
    auto fold = [](uint64_t y) -> uint32_t {
        y = (y | (y >> 6)) & 0x000F000F000F000Full;
        y = (y | (y >> 12)) & 0x000000FF000000FFull;
        y = (y | (y >> 24)) & 0xFFFFull;
        return (uint32_t)y;
    };
    return fold(lo) | (fold(hi) << 16);
}

template <class F> static inline void wq_fq_parallel(unsigned nt, F&& f) {
    if (nt <= 1) { f(0u); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back([&f, t]() { f(t); });
    for (auto& x : th) x.join();
}

// BGZF block at `p` (RFC 1952 member with FEXTRA subfield 'B','C'): total block size, or 0 when it is not one
static inline size_t wq_bgzf_block_size(const unsigned char* p, size_t avail) {
    if (avail < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
    const size_t xlen = p[10] | ((size_t)p[11] << 8);
    if (avail < 12 + xlen) return 0;
    for (size_t o = 12; o + 4 <= 12 + xlen;) {
        const size_t slen = p[o + 2] | ((size_t)p[o + 3] << 8);
        if (p[o] == 'B' && p[o + 1] == 'C' && slen == 2) return (size_t)(p[o + 4] | ((size_t)p[o + 5] << 8)) + 1;
        o += 4 + slen;
    }
    return 0;
}

static inline int wq_fastq_open_impl(wq_fastq* f, const char* path) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) { f->err = std::string("cannot open ") + path; return -20; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); f->err = std::string("cannot stat ") + path; return -20; }
    unsigned char head[32];
    const ssize_t got = pread(fd, head, sizeof head, 0);
    const bool is_gz = got >= 2 && head[0] == 0x1f && head[1] == 0x8b;
    const bool is_bgzf = is_gz && wq_bgzf_block_size(head, (size_t)got) != 0;
    const bool use_zlib = getenv("WQ_GZ_ZLIB") && atoi(getenv("WQ_GZ_ZLIB")) != 0;
    if (is_gz && !is_bgzf && !use_zlib && st.st_size > 0) {
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) {
            madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
            f->map = (const char*)m; f->map_len = (size_t)st.st_size; f->map_owned = true; f->fd = fd;
            f->inf = new WqInflate();
            f->inf->open(m, (size_t)st.st_size);
            return 0;
        }
    }
    if (is_gz && !is_bgzf) {
        close(fd);
        f->gz = gzopen(path, "rb");
        if (!f->gz) { f->err = std::string("cannot open ") + path; return -20; }
        gzbuffer(f->gz, 1u << 20);
        return 0;
    }
    if (st.st_size > 0) {
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); f->err = std::string("cannot map ") + path; return -20; }
        madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
        f->map = (const char*)m; f->map_len = (size_t)st.st_size; f->map_owned = true;
    }
    f->fd = fd;
    f->bgzf = is_bgzf;
    if (!is_bgzf) { f->wbase = f->map; f->woff = 0; f->wlen = f->map_len; f->eof = true; }      // the whole text is the window
    return 0;
}

static inline void wq_fastq_release(wq_fastq* f) {
    if (f->gz) gzclose(f->gz);
    delete f->inf; f->inf = nullptr;
    if (f->map_owned && f->map) munmap((void*)f->map, f->map_len);
    if (f->fd >= 0) close(f->fd);
    f->slot[0].release(); f->slot[1].release();
}

// gzip modes: drop the consumed prefix of the window (between batches only) and append more inflated text
static inline bool wq_fastq_pull(wq_fastq* f, size_t want) {
    if (f->eof) return true;
    const uint64_t keep_from = f->line_start;                       // everything before the first unconsumed line can go
    if (keep_from > f->woff) {
        f->text.drop_front((size_t)(keep_from - f->woff));
        f->woff = keep_from;
    }
    if (f->inf) {
        const size_t old = f->text.size();
        if (!f->text.reserve_more(want)) { f->err = "out of memory"; return false; }
        const size_t got = f->inf->read((uint8_t*)f->text.data() + old, want);
        if (f->inf->failed()) { f->err = f->inf->err; return false; }
        if (got < want) f->eof = true;
        f->text.len = old + got;
    } else if (f->gz) {
        const size_t old = f->text.size();
        if (!f->text.reserve_more(want)) { f->err = "out of memory"; return false; }
        size_t got = 0;
        while (got < want) {
            int r = gzread(f->gz, f->text.data() + old + got, (unsigned)std::min<size_t>(want - got, 1u << 30));
            if (r < 0) { int en = 0; f->err = std::string("gzread: ") + gzerror(f->gz, &en); return false; }
            if (r == 0) { f->eof = true; break; }
            got += (size_t)r;
        }
        f->text.len = old + got;
    } else if (f->bgzf) {
        // blocks up to ~want bytes of text: sizes from the block trailers (ISIZE), then every block inflated by some thread
        struct Blk { size_t off, csize, hdr; uint32_t isize; size_t out; };
        std::vector<Blk> blks;
        size_t total = 0;
        const unsigned char* base = (const unsigned char*)f->map;
        while (f->bgzf_pos < f->map_len && total < want) {
            const size_t bs = wq_bgzf_block_size(base + f->bgzf_pos, f->map_len - f->bgzf_pos);
            if (bs < 26 || f->bgzf_pos + bs > f->map_len) { f->err = "malformed BGZF block"; return false; }
            const unsigned char* p = base + f->bgzf_pos;
            const size_t xlen = p[10] | ((size_t)p[11] << 8);
            if (12 + xlen + 8 > bs) { f->err = "malformed BGZF block"; return false; }
            uint32_t isize;
            memcpy(&isize, p + bs - 4, 4);
            blks.push_back(Blk{f->bgzf_pos, bs, 12 + xlen, isize, total});
            total += isize;
            f->bgzf_pos += bs;
        }
        if (f->bgzf_pos >= f->map_len) f->eof = true;
        const size_t old = f->text.size();
        if (!f->text.reserve_more(total)) { f->err = "out of memory"; return false; }
        f->text.len = old + total;
        char* dst = f->text.data() + old;
        std::atomic<size_t> next{0};
        std::atomic<int> bad{0};
        const unsigned nt = blks.size() < 8 ? 1u : wq_host_threads();
        static const bool bgzf_zlib = getenv("WQ_GZ_ZLIB") && atoi(getenv("WQ_GZ_ZLIB")) != 0;
        wq_fq_parallel(nt, [&](unsigned) {
            if (!bgzf_zlib) {                                // the library's own inflate, straight into the block's place in the window
                WqInflate* z = new WqInflate();
                for (size_t b; (b = next.fetch_add(1)) < blks.size();) {
                    const Blk& k = blks[b];
                    if (k.isize == 0) continue;
                    if (!z->raw_into(base + k.off + k.hdr, k.csize - k.hdr - 8, (uint8_t*)dst + k.out, k.isize)) bad = 1;
                }
                delete z;
                return;
            }
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
            for (size_t b; (b = next.fetch_add(1)) < blks.size();) {
                const Blk& k = blks[b];
                if (k.isize == 0) continue;
                inflateReset(&zs);
                zs.next_in = (Bytef*)(base + k.off + k.hdr);
                zs.avail_in = (uInt)(k.csize - k.hdr - 8);
                zs.next_out = (Bytef*)(dst + k.out);
                zs.avail_out = k.isize;
                const int r = inflate(&zs, Z_FINISH);
                if (r != Z_STREAM_END || zs.avail_out != 0) bad = 1;
            }
            inflateEnd(&zs);
        });
        if (bad) { f->err = "BGZF block failed to inflate"; return false; }
    }
    f->wbase = f->text.data();
    f->wlen = f->text.size();
    return true;
}

// newline offsets of the window part [from, to) appended to f->nl, scanned in parallel
static inline void wq_fastq_scan(wq_fastq* f, uint64_t from, uint64_t to) {
    if (to <= from) return;
    const uint64_t span = to - from;
    unsigned nt = span < (4u << 20) ? 1u : wq_host_threads();
    std::vector<std::vector<uint64_t>> found(nt);
    const char* base = f->wbase - f->woff;                         // base + stream offset = address
    wq_fq_parallel(nt, [&](unsigned t) {
        const uint64_t a = from + span * t / nt, b = from + span * (t + 1) / nt;
        std::vector<uint64_t>& v = found[t];
        v.reserve((size_t)((b - a) / 60 + 16));
        const char* p = base + a;
        const char* e = base + b;
        // 64 bytes per step: four 16-byte compares folded into one 64-bit mask (lines are short, a memchr call per line costs more)
        const __m128i nlv = _mm_set1_epi8('\n');
        while (p + 64 <= e) {
            const uint64_t m0 = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)p), nlv));
            const uint64_t m1 = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + 16)), nlv));
            const uint64_t m2 = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + 32)), nlv));
            const uint64_t m3 = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + 48)), nlv));
            uint64_t m = m0 | (m1 << 16) | (m2 << 32) | (m3 << 48);
            const uint64_t at = (uint64_t)(p - base);
            while (m) { v.push_back(at + (uint64_t)__builtin_ctzll(m)); m &= m - 1; }
            p += 64;
        }
        for (; p < e; p++) if (*p == '\n') v.push_back((uint64_t)(p - base));
    });
    size_t add = 0;
    for (auto& v : found) add += v.size();
    const size_t old = f->nl.size();
    f->nl.resize(old + add);
    size_t at = old;
    for (auto& v : found) { memcpy(f->nl.data() + at, v.data(), v.size() * 8); at += v.size(); }
}

// Next batch of at most max_reads records.  out->n_reads == 0 at end of input.  The buffers belong to the reader and
// stay valid until the call after the next one (two slots alternate, so a batch can be aligned while the next is parsed).
static inline int wq_fastq_next_impl(wq_fastq* f, uint32_t max_reads, wq_read_batch* out) {
    memset(out, 0, sizeof *out);
    static const bool prof = getenv("WQ_FASTQ_PROF") != nullptr;
    auto now = []() { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t_begin = now();
    // ---- forget the consumed part of the line index ----
    if (f->nl_pos > 0) { f->nl.erase(f->nl.begin(), f->nl.begin() + f->nl_pos); f->nl_pos = 0; }
    // ---- enough lines for the batch, or the end of the input ----
    const size_t need = (size_t)max_reads * 4;
    for (;;) {
        if (f->nl.size() >= need) break;
        const uint64_t wend = f->woff + f->wlen;
        if (f->scanned < wend) {
            const uint64_t want = (uint64_t)((double)(need - f->nl.size()) / 4.0 * f->avg_rec * 1.05) + (1u << 16);
            const uint64_t to = std::min(wend, f->scanned + want);
            wq_fastq_scan(f, f->scanned, to);
            f->scanned = to;
            continue;
        }
        if (!f->eof) {
            const size_t want = (size_t)std::max<double>(32.0 * (1 << 20), (double)(need - f->nl.size()) / 4.0 * f->avg_rec * 1.05);
            if (!wq_fastq_pull(f, want)) return -20;
            continue;
        }
        // end of input: an unterminated last line counts as a line
        if (!f->closed_tail) {
            f->closed_tail = true;
            const uint64_t last_start = f->nl.empty() ? f->line_start : f->nl.back() + 1;
            if (wend > last_start) f->nl.push_back(wend);
        }
        break;
    }
    const auto t_lines = now();
    // trailing blank lines at the end of the file are not records
    size_t nlines = std::min(f->nl.size(), need);
    const char* base = f->wbase - f->woff;
    if (f->nl.size() < need) {
        auto blank = [&](size_t k) {
            const uint64_t s = k == 0 ? f->line_start : f->nl[k - 1] + 1, e = f->nl[k];
            return e == s || (e == s + 1 && base[s] == '\r');
        };
        while (nlines > 0 && blank(nlines - 1)) nlines--;
        if (nlines % 4 != 0) { f->err = "truncated FASTQ record after record " + std::to_string(f->records + nlines / 4); return -21; }
    }
    const uint32_t n = (uint32_t)(nlines / 4);
    f->cur ^= 1;
    WqFastqSlot& S = f->slot[f->cur];
    S.n = n;
    if (n == 0) { f->nl_pos = f->nl.size(); return 0; }
    uint8_t* len = (uint8_t*)S.len.ensure(n);
    uint64_t* seq_off = (uint64_t*)S.seq_off.ensure((size_t)n * 8);
    uint64_t* qual_off = (uint64_t*)S.qual_off.ensure((size_t)n * 8);
    if (!len || !seq_off || !qual_off) { f->err = "out of memory"; return -22; }
    S.name_off.resize(n); S.name_len.resize(n);
    const uint64_t first = f->line_start;
    const uint64_t* nl = f->nl.data();
    auto lstart = [&](size_t k) -> uint64_t { return k == 0 ? first : nl[k - 1] + 1; };
    auto llen = [&](size_t k) -> size_t {              // without the newline and a carriage return before it
        const uint64_t s = lstart(k), e = nl[k];
        size_t m = (size_t)(e - s);
        if (m && base[e - 1] == '\r') m--;
        return m;
    };
    unsigned nt = n < 4096 ? 1u : wq_host_threads();
    // ---- pass 1: lengths + validation + per-thread totals ----
    std::vector<uint64_t> tw(nt + 1, 0), tq(nt + 1, 0), tn(nt + 1, 0);
    std::vector<int64_t> bad(nt, -1);
    std::vector<int> why(nt, 0);
    wq_fq_parallel(nt, [&](unsigned t) {
        const uint32_t lo = (uint32_t)((uint64_t)n * t / nt), hi = (uint32_t)((uint64_t)n * (t + 1) / nt);
        uint64_t w = 0, q = 0, nm = 0;
        for (uint32_t i = lo; i < hi; i++) {
            const size_t k = (size_t)i * 4;
            const size_t l0 = llen(k), sl = llen(k + 1), l2 = llen(k + 2), ql = llen(k + 3);
            if (l0 == 0 || base[lstart(k)] != '@' || l2 == 0 || base[lstart(k + 2)] != '+') { if (bad[t] < 0) { bad[t] = i; why[t] = 1; } continue; }
            if (sl != ql) { if (bad[t] < 0) { bad[t] = i; why[t] = 2; } continue; }
            if (sl > 255) { if (bad[t] < 0) { bad[t] = i; why[t] = 3; } continue; }
            len[i] = (uint8_t)sl;
            S.name_len[i] = (uint32_t)(l0 - 1);
            w += (sl + 31) / 32 + (sl == 0 ? 1 : 0);
            q += (sl + 3) & ~(size_t)3;
            nm += l0 - 1;
        }
        tw[t + 1] = w; tq[t + 1] = q; tn[t + 1] = nm;
    });
    const auto t_pass1 = now();
    for (unsigned t = 0; t < nt; t++)
        if (bad[t] >= 0) {
            const std::string r = std::to_string(f->records + (uint64_t)bad[t] + 1);
            f->err = why[t] == 1 ? "malformed FASTQ record " + r + " (expected '@' / '+' lines)"
                   : why[t] == 2 ? "FASTQ record " + r + ": sequence and quality lengths differ"
                                 : "FASTQ record " + r + ": read longer than 255 (ReadLengthInt = UInt8, src/align.jl:74)";
            return -21;
        }
    for (unsigned t = 0; t < nt; t++) { tw[t + 1] += tw[t]; tq[t + 1] += tq[t]; tn[t + 1] += tn[t]; }
    const uint64_t w = tw[nt], q = tq[nt], nm = tn[nt];
    S.seq2_words = w; S.qual_bytes = q;
    uint64_t* seq2 = (uint64_t*)S.seq2.ensure((size_t)std::max<uint64_t>(w, 1) * 8);
    uint8_t* qual = (uint8_t*)S.qual.ensure((size_t)std::max<uint64_t>(q, 4));
    if (!seq2 || !qual) { f->err = "out of memory"; return -22; }
    S.names.resize(nm);
    const uint8_t* T = wq_fastq_code_table();
    const uint8_t qo = (uint8_t)f->qualoffset;
    const auto t_alloc = now();
    // ---- pass 2: offsets of the thread's range, then the conversion of its records ----
    wq_fq_parallel(nt, [&](unsigned t) {
        const uint32_t lo = (uint32_t)((uint64_t)n * t / nt), hi = (uint32_t)((uint64_t)n * (t + 1) / nt);
        uint64_t wo = tw[t], qo_ = tq[t], no = tn[t];
        for (uint32_t i = lo; i < hi; i++) {
            const size_t k = (size_t)i * 4;
            const uint32_t L = len[i];
            seq_off[i] = wo; qual_off[i] = qo_; S.name_off[i] = no;
            const char* sp = base + lstart(k + 1);
            const char* qp = base + lstart(k + 3);
            uint64_t* wd = seq2 + wo;
            const uint32_t nw = (L + 31) / 32;
            for (uint32_t kk = 0; kk < nw; kk++) {
                const uint32_t b0 = kk * 32, b1 = std::min(L, b0 + 32);
                if (b1 - b0 == 32) { wd[kk] = (uint64_t)wq_fastq_pack16(sp + b0) | ((uint64_t)wq_fastq_pack16(sp + b0 + 16) << 32); continue; }
                uint64_t acc = 0;
                for (uint32_t j = b0; j < b1; j++) acc |= (uint64_t)T[(uint8_t)sp[j]] << (2 * (j - b0));
                wd[kk] = acc;
            }
            if (L == 0) wd[0] = 0;
            uint8_t* qd = qual + qo_;
            {   // UInt8 arithmetic (src/record.jl:17), 16 bytes at a time
                const __m128i qov = _mm_set1_epi8((char)qo);
                uint32_t j = 0;
                for (; j + 16 <= L; j += 16) _mm_storeu_si128((__m128i*)(qd + j), _mm_sub_epi8(_mm_loadu_si128((const __m128i*)(qp + j)), qov));
                for (; j < L; j++) qd[j] = (uint8_t)((uint8_t)qp[j] - qo);
            }
            for (uint32_t j = L; j < ((L + 3) & ~3u); j++) qd[j] = 0;
            memcpy(S.names.data() + no, base + lstart(k) + 1, S.name_len[i]);
            wo += nw + (L == 0 ? 1 : 0);
            qo_ += (L + 3) & ~3u;
            no += S.name_len[i];
        }
    });
    if (prof)
        fprintf(stderr, "[wq] fastq batch %u reads: lines (inflate + scan) %.2f ms, pass 1 %.2f ms, buffers %.2f ms, pass 2 %.2f ms\n", n,
                ms(t_begin, t_lines), ms(t_lines, t_pass1), ms(t_pass1, t_alloc), ms(t_alloc, now()));
    // ---- consume ----
    const uint64_t consumed_to = nl[nlines - 1] + 1;
    f->avg_rec = std::max(16.0, (double)(consumed_to - f->line_start) / (double)n);
    f->line_start = std::min<uint64_t>(consumed_to, f->woff + f->wlen);
    f->nl_pos = f->nl.size() < need ? f->nl.size() : nlines;      // at the end of the input the blank tail is dropped too
    f->records += n;
    out->n_reads = n; out->len = len; out->seq_off = seq_off; out->seq2 = seq2; out->seq2_words = w;
    out->qual_off = qual_off; out->qual = qual; out->qual_bytes = q;
    return 0;
}
