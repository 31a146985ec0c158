// Streaming gzip inflate for FASTQ ingest (RFC 1951 / 1952), written for the one thing zlib is slow at here: long runs of
// Huffman-coded literals (sequence lines are ~2 bits per base of literals; zlib's inflate moves ~0.2-0.4 GB/s of such text, and an
// ordinary -- single-member -- .fastq.gz cannot be inflated in parallel, so this loop IS the file -> result rate of such a file).
//
//   * the whole compressed file is in memory (mapped), so the bit reader refills 7-8 bytes at a time from an unaligned 64-bit load
//     and one refill covers a complete length / distance pair (at most 48 bits);
//   * literal / length codes are looked up in an 11-bit table (sub-tables for longer codes), distances in an 8-bit table; an entry
//     carries the bits to consume, the kind, the number of extra bits and the literal / base value;
//   * several literals are decoded per refill; matches are copied 8 bytes at a time;
//   * output goes to an internal window (32 KB of history + a few MB) from which read() hands text to the caller, so the caller's
//     buffer may drop consumed text at any time; the CRC-32 of every member is checked (carry-less multiplication, ~13 GB/s) and so is ISIZE.
//
// Members may follow each other (cat a.gz b.gz); bytes after the last member that do not start a member are ignored, as gzread does.
// Any malformed input ends in an error message, never in an out-of-bounds access: every table index is masked, every distance is
// checked against the history of the current member, and input / output positions are checked before every step of the slow path
// (the fast loop runs only while 8 input bytes and 274 output bytes are in reach).
// Checked against zlib on the CPU (tests/test_inflate_cpu.py: every compression level and strategy, stored / fixed / dynamic blocks,
// multi-member files, truncation and corruption).
#pragma once
#include <zlib.h>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

// CRC-32 of the gzip trailer (reflected polynomial 0xEDB88320) by carry-less multiplication: 64 bytes per step folded into four 128-bit
// accumulators, then reduced (V. Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ", Intel 2009; the folding
// constants are the published ones for this polynomial).  About three times zlib's table-driven crc32, which was a sixth of the inflate
// time.  len >= 64 and a multiple of 16; crc is the running register (pre-conditioned), not zlib's public value.
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("pclmul,sse4.1")))
static inline uint32_t wq_crc32_clmul(uint32_t crc, const uint8_t* buf, size_t len) {
    alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};
    alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};
    alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};
    alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
    x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
    x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = _mm_load_si128((const __m128i*)k1k2);
    buf += 64; len -= 64;
    while (len >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00)); y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
        y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20)); y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5); x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7); x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64; len -= 64;
    }
    x0 = _mm_load_si128((const __m128i*)k3k4);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128((const __m128i*)buf);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16; len -= 16;
    }
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64((const __m128i*)k5k0);
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_load_si128((const __m128i*)poly);
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
#endif
// zlib's convention (crc32(0, ...) starts a checksum); the tail and machines without the instruction go through zlib
static inline uint32_t wq_crc32(uint32_t crc, const uint8_t* p, size_t n) {
#if defined(__x86_64__)
    static const bool have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    if (have && n >= 64) {
        const size_t m = n & ~(size_t)15;
        crc = ~wq_crc32_clmul(~crc, p, m);
        p += m; n -= m;
    }
#endif
    while (n) { const size_t k = std::min<size_t>(n, 1u << 30); crc = (uint32_t)crc32(crc, p, (uInt)k); p += k; n -= k; }
    return crc;
}

struct WqInflate {
    static constexpr uint32_t LIT_BITS = 11, DIST_BITS = 8, HIST = 32768;
    static constexpr size_t CHUNK = 4u << 20, SLACK = 320;
    enum { K_LIT = 0, K_BASE = 1, K_EOB = 2, K_SUB = 3, K_BAD = 4, K_LIT2 = 5 };      // K_LIT2: two literals in one entry (value = first | second << 8)
    enum { S_HEADER, S_BLOCK, S_STORED, S_HUFF, S_TRAILER, S_DONE, S_ERROR };

    const uint8_t* in = nullptr; const uint8_t* in_end = nullptr; const uint8_t* p = nullptr;
    uint64_t bitbuf = 0; uint32_t bitcnt = 0;
    std::vector<uint8_t> buf;              // [history | fresh output | slack]
    uint8_t* obase = nullptr;              // where the output goes: buf, or the caller's memory in raw_into()
    size_t out_pos = 0, read_pos = 0, crc_pos = 0, win_lo = 0;      // offsets into buf; win_lo = first byte of the current member's history
    int state = S_HEADER;
    bool last_block = false;
    uint32_t stored_left = 0;
    uint32_t crc = 0; uint64_t member_out = 0;
    uint64_t members = 0;
    uint32_t lit[4096], dist[1024];
    std::string err;

    void open(const void* data, size_t n) {
        in = p = (const uint8_t*)data; in_end = in + n;
        bitbuf = 0; bitcnt = 0;
        buf.assign(HIST + CHUNK + SLACK, 0);
        obase = buf.data();
        out_pos = read_pos = crc_pos = win_lo = 0;
        state = S_HEADER; last_block = false; stored_left = 0; crc = 0; member_out = 0; members = 0; err.clear();
    }
    bool failed() const { return state == S_ERROR; }

    // hands out up to `want` bytes of inflated text; 0 = end of the stream (or an error: see failed())
    size_t read(uint8_t* dst, size_t want) {
        size_t got = 0;
        while (got < want) {
            if (read_pos < out_pos) {
                const size_t n = std::min(want - got, out_pos - read_pos);
                memcpy(dst + got, buf.data() + read_pos, n);
                read_pos += n; got += n;
                continue;
            }
            if (state == S_DONE || state == S_ERROR) break;
            if (out_pos > HIST + CHUNK / 2) slide();
            decode_some();
        }
        return got;
    }

    // One raw deflate stream (RFC 1951, e.g. the payload of a blocked-gzip block) straight into dst[0 .. out_len): no window, no copy.
    // The fast loop may write a little past the symbol it is decoding, so it stops 300 bytes before the end and an exact loop
    // finishes -- the bytes behind dst + out_len belong to somebody else (the next block, inflated by another thread).
    bool raw_into(const void* data, size_t n, uint8_t* dst, size_t out_len) {
        in = p = (const uint8_t*)data; in_end = in + n;
        bitbuf = 0; bitcnt = 0; obase = dst; out_pos = read_pos = crc_pos = win_lo = 0;
        state = S_BLOCK; last_block = false; stored_left = 0; err.clear();
        while (state != S_TRAILER && state != S_ERROR) {
            if (state == S_BLOCK) block_header();
            else if (state == S_STORED) {
                if (stored_left > out_len - out_pos) { fail("more output than announced"); break; }
                if (stored_left > (size_t)(in_end - p)) { fail("truncated stored block"); break; }
                memcpy(dst + out_pos, p, stored_left);
                p += stored_left; out_pos += stored_left; stored_left = 0;
                state = last_block ? S_TRAILER : S_BLOCK;
            } else if (state == S_HUFF) {
                if (out_len - out_pos > 300) huffman<false>(out_len - 300);
                if (state == S_HUFF) huffman<true>(out_len);
            } else break;
        }
        if (state == S_TRAILER && out_pos != out_len) fail("less output than announced");
        return state == S_TRAILER;
    }

private:
    static uint64_t load64(const uint8_t* q) { uint64_t v; memcpy(&v, q, 8); return v; }
    static void store64(uint8_t* q, uint64_t v) { memcpy(q, &v, 8); }
    bool fail(const char* what) { err = std::string("gzip: ") + what; state = S_ERROR; return false; }

    void refill() {
        if (in_end - p >= 8) { bitbuf |= load64(p) << bitcnt; p += 7 - (bitcnt >> 3); bitcnt |= 56; }      // bits above bitcnt are the stream's next bits (re-ORed identically later)
        else while (bitcnt <= 56 && p < in_end) { bitbuf |= (uint64_t)*p++ << bitcnt; bitcnt += 8; }
    }
    bool need(uint32_t n) { if (bitcnt < n) { refill(); if (bitcnt < n) return fail("truncated input"); } return true; }
    uint32_t take(uint32_t n) { const uint32_t v = (uint32_t)(bitbuf & ((1ull << n) - 1)); bitbuf >>= n; bitcnt -= n; return v; }
    void align_to_byte() {                  // drop the rest of the current byte and give the whole unread bytes back to the input
        const uint32_t r = bitcnt & 7; bitbuf >>= r; bitcnt -= r;
        p -= bitcnt >> 3; bitbuf = 0; bitcnt = 0;
    }
    void crc_upto(size_t pos) { if (pos > crc_pos) { crc = wq_crc32(crc, buf.data() + crc_pos, pos - crc_pos); member_out += pos - crc_pos; crc_pos = pos; } }
    void slide() {                          // everything has been handed out: keep the last 32 KB as history at the front
        crc_upto(out_pos);
        const size_t keep = std::min<size_t>(out_pos, HIST), shift = out_pos - keep;
        if (!shift) return;
        memmove(buf.data(), buf.data() + shift, keep);
        out_pos = read_pos = crc_pos = keep;
        win_lo = win_lo > shift ? win_lo - shift : 0;
    }

    // table entry: bits 0-4 = bits to consume (code + extra bits, so one shift moves past the whole symbol and the next lookup does not wait
    // for the extra bits to be extracted), 5-7 = kind, 8-11 = length of the code alone (K_SUB: index bits of the sub-table), 12-31 = literal /
    // base value / sub-table start
    static uint32_t entry(uint32_t codelen, uint32_t kind, uint32_t extra, uint32_t value) { return (kind == K_SUB ? codelen : codelen + extra) | kind << 5 | (kind == K_SUB ? extra : codelen) << 8 | value << 12; }
    static uint32_t rev_bits(uint32_t code, uint32_t len) { uint32_t r = 0; for (uint32_t i = 0; i < len; i++) { r = r << 1 | (code & 1); code >>= 1; } return r; }

    // canonical Huffman decoding table from code lengths; what = 0 literal / length alphabet, 1 distance alphabet, 2 code-length alphabet
    bool build(const uint8_t* lens, uint32_t nsyms, uint32_t tbits, uint32_t* table, uint32_t cap, int what) {
        static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        uint32_t count[16] = {0};
        for (uint32_t s = 0; s < nsyms; s++) { if (lens[s] > 15) return fail("bad code length"); count[lens[s]]++; }
        count[0] = 0;
        int64_t left = 1;
        for (uint32_t l = 1; l <= 15; l++) { left <<= 1; left -= count[l]; if (left < 0) return fail("over-subscribed code"); }
        uint32_t next[16]; next[0] = 0; next[1] = 0;
        for (uint32_t l = 1; l < 15; l++) next[l + 1] = (next[l] + count[l]) << 1;
        const uint32_t tsize = 1u << tbits;
        for (uint32_t i = 0; i < tsize; i++) table[i] = entry(0, K_BAD, 0, 0);      // an incomplete code leaves unused patterns: hitting one is an error
        auto make = [&](uint32_t s, uint32_t nbits) -> uint32_t {
            if (what == 2) return entry(nbits, K_LIT, 0, s);
            if (what == 1) return s < 30 ? entry(nbits, K_BASE, dext[s], dbase[s]) : entry(nbits, K_BAD, 0, 0);
            if (s < 256) return entry(nbits, K_LIT, 0, s);
            if (s == 256) return entry(nbits, K_EOB, 0, 0);
            return s < 286 ? entry(nbits, K_BASE, lext[s - 257], lbase[s - 257]) : entry(nbits, K_BAD, 0, 0);
        };
        // codes longer than the table: one sub-table per table prefix, as wide as the longest code under that prefix
        uint8_t submax[1u << LIT_BITS];
        bool any_long = false;
        uint32_t code_of[320];
        {
            uint32_t nx[16];
            memcpy(nx, next, sizeof nx);
            for (uint32_t s = 0; s < nsyms; s++) { const uint32_t l = lens[s]; if (l) code_of[s] = rev_bits(nx[l]++, l); }
        }
        for (uint32_t s = 0; s < nsyms; s++) if (lens[s] > tbits) { if (!any_long) { memset(submax, 0, tsize); any_long = true; } uint8_t& m = submax[code_of[s] & (tsize - 1)]; if (lens[s] > m) m = lens[s]; }
        uint32_t used = tsize;
        if (any_long)
            for (uint32_t i = 0; i < tsize; i++) if (submax[i]) {
                const uint32_t sb = submax[i] - tbits, n = 1u << sb;
                if (used + n > cap) return fail("code table overflow");
                table[i] = entry(tbits, K_SUB, sb, used);
                for (uint32_t k = 0; k < n; k++) table[used + k] = entry(0, K_BAD, 0, 0);
                used += n;
            }
        for (uint32_t s = 0; s < nsyms; s++) {
            const uint32_t l = lens[s];
            if (!l) continue;
            const uint32_t c = code_of[s];
            if (l <= tbits) { const uint32_t e = make(s, l); for (uint32_t i = c; i < tsize; i += 1u << l) table[i] = e; }
            else {
                const uint32_t sub = table[c & (tsize - 1)], sb = (sub >> 8) & 15, start = sub >> 12, e = make(s, l - tbits);
                for (uint32_t i = c >> tbits; i < (1u << sb); i += 1u << (l - tbits)) table[start + i] = e;
            }
        }
        if (what == 0) {
            // Two literals per lookup where both codes fit into the table's index: the decoding chain is one table load + one shift per
            // lookup, and FASTQ text is almost only literals with 2-4 bit codes.  The second symbol is looked up with the unknown high
            // bits as zeros, which is right whenever its code fits into the bits that are known (entries repeat over the higher bits).
            uint32_t single[1u << LIT_BITS];
            memcpy(single, table, sizeof(uint32_t) << tbits);
            for (uint32_t i = 0; i < tsize; i++) {
                const uint32_t e1 = single[i], n1 = e1 & 31;
                if ((e1 & 0xE0) != 0 || n1 == 0 || n1 >= tbits) continue;
                const uint32_t e2 = single[i >> n1], n2 = e2 & 31;
                if ((e2 & 0xE0) != 0 || n2 == 0 || n1 + n2 > tbits) continue;
                table[i] = entry(n1 + n2, K_LIT2, 0, (e1 >> 12) | (e2 >> 12) << 8);
            }
        }
        return true;
    }

    bool fixed_tables() {
        uint8_t l[288];
        for (int i = 0; i < 144; i++) l[i] = 8;
        for (int i = 144; i < 256; i++) l[i] = 9;
        for (int i = 256; i < 280; i++) l[i] = 7;
        for (int i = 280; i < 288; i++) l[i] = 8;
        if (!build(l, 288, LIT_BITS, lit, 4096, 0)) return false;
        uint8_t d[32];
        for (int i = 0; i < 32; i++) d[i] = 5;
        return build(d, 32, DIST_BITS, dist, 1024, 1);
    }

    bool dynamic_tables() {
        if (!need(14)) return false;
        const uint32_t hlit = take(5) + 257, hdist = take(5) + 1, hclen = take(4) + 4;
        if (hlit > 286 || hdist > 30) return fail("too many length or distance symbols");
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t pl[19] = {0};
        for (uint32_t i = 0; i < hclen; i++) { if (!need(3)) return false; pl[order[i]] = (uint8_t)take(3); }
        uint32_t pre[128];
        if (!build(pl, 19, 7, pre, 128, 2)) return false;
        uint8_t l[320];
        uint32_t n = 0;
        while (n < hlit + hdist) {
            refill();
            const uint32_t e = pre[bitbuf & 127], nb = e & 31;
            if (((e >> 5) & 7) != K_LIT || nb > bitcnt) return fail("bad code-length code");
            take(nb);
            const uint32_t sym = e >> 12;
            if (sym < 16) { l[n++] = (uint8_t)sym; continue; }
            uint32_t rep, val = 0;
            if (sym == 16) { if (!n) return fail("repeat without a previous length"); if (bitcnt < 2) return fail("truncated input"); val = l[n - 1]; rep = 3 + take(2); }
            else if (sym == 17) { if (bitcnt < 3) return fail("truncated input"); rep = 3 + take(3); }
            else { if (bitcnt < 7) return fail("truncated input"); rep = 11 + take(7); }
            if (n + rep > hlit + hdist) return fail("code lengths overrun");
            while (rep--) l[n++] = (uint8_t)val;
        }
        if (!l[256]) return fail("no end-of-block code");
        if (!build(l, hlit, LIT_BITS, lit, 4096, 0)) return false;
        return build(l + hlit, hdist, DIST_BITS, dist, 1024, 1);
    }

    bool header() {
        const uint8_t* q = p;
        if (in_end - q < 10) return fail("truncated header");
        if (q[0] != 0x1f || q[1] != 0x8b) return fail("not a gzip member");
        if (q[2] != 8) return fail("unknown compression method");
        const uint8_t flg = q[3];
        if (flg & 0xE0) return fail("reserved header flags set");
        q += 10;
        if (flg & 4) { if (in_end - q < 2) return fail("truncated header"); const size_t xl = q[0] | (size_t)q[1] << 8; q += 2; if ((size_t)(in_end - q) < xl) return fail("truncated header"); q += xl; }
        for (int k = 0; k < 2; k++) if (flg & (k ? 16 : 8)) { while (q < in_end && *q) q++; if (q >= in_end) return fail("truncated header"); q++; }
        if (flg & 2) { if (in_end - q < 2) return fail("truncated header"); q += 2; }
        p = q; bitbuf = 0; bitcnt = 0;
        crc_upto(out_pos);
        crc = 0; member_out = 0; win_lo = out_pos; members++;
        return true;
    }

    bool trailer() {
        align_to_byte();
        if (in_end - p < 8) return fail("truncated trailer");
        crc_upto(out_pos);
        const uint32_t want_crc = p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
        const uint32_t want_len = p[4] | (uint32_t)p[5] << 8 | (uint32_t)p[6] << 16 | (uint32_t)p[7] << 24;
        p += 8;
        if (want_crc != crc) return fail("CRC mismatch");
        if (want_len != (uint32_t)member_out) return fail("length mismatch");
        state = (in_end - p >= 2 && p[0] == 0x1f && p[1] == 0x8b) ? S_HEADER : S_DONE;      // another member, or the end (trailing bytes are ignored)
        return true;
    }

    void block_header() {
        if (!need(3)) return;
        last_block = take(1) != 0;
        const uint32_t type = take(2);
        if (type == 0) {
            align_to_byte();
            if (in_end - p < 4) { fail("truncated stored block"); return; }
            const uint32_t len = p[0] | (uint32_t)p[1] << 8, nlen = p[2] | (uint32_t)p[3] << 8;
            if ((len ^ 0xFFFF) != nlen) { fail("stored block length check"); return; }
            p += 4; stored_left = len; state = S_STORED;
        } else if (type == 1) { if (fixed_tables()) state = S_HUFF; }
        else if (type == 2) { if (dynamic_tables()) state = S_HUFF; }
        else fail("reserved block type");
    }

    // one step of the state machine; returns when the output window is full, a member ended, or the state changed
    void decode_some() {
        const size_t out_limit = HIST + CHUNK;            // decoding stops once out_pos passes this (SLACK bytes remain behind it)
        switch (state) {
        case S_HEADER: if (header()) state = S_BLOCK; return;
        case S_BLOCK: block_header(); return;
        case S_STORED: {
            if (out_pos >= out_limit) return;
            const size_t n = std::min<size_t>(stored_left, std::min<size_t>(out_limit + SLACK / 2 - out_pos, (size_t)(in_end - p)));
            if (n == 0 && stored_left) { fail("truncated stored block"); return; }
            memcpy(buf.data() + out_pos, p, n);
            p += n; out_pos += n; stored_left -= (uint32_t)n;
            if (!stored_left) state = last_block ? S_TRAILER : S_BLOCK;
            return;
        }
        case S_HUFF: huffman<false>(out_limit); return;
        case S_TRAILER: trailer(); return;
        default: return;
        }
    }

    // The bit reader's state lives in locals here: the output is written through a byte pointer, which may alias anything, so members
    // would be reloaded after every literal.
    // EXACT: nothing is written behind base + out_limit (one byte at a time, every write checked); the loop then ends only at the end of
    // the block, and output beyond the limit is an error.
    template <bool EXACT> void huffman(size_t out_limit) {
        uint8_t* const base = obase;
        uint8_t* out = base + out_pos;
        uint8_t* const out_stop = base + out_limit;
        const uint8_t* const lo = base + win_lo;
        const uint32_t* const LT = lit; const uint32_t* const DT = dist;
        const uint32_t lmask = (1u << LIT_BITS) - 1, dmask = (1u << DIST_BITS) - 1;
        uint64_t bb = bitbuf; uint32_t bc = bitcnt; const uint8_t* q = p; const uint8_t* const qe = in_end;
        const char* bad = nullptr;
        #define WQ_INF_REFILL() do { if (qe - q >= 8) { bb |= load64(q) << bc; q += 7 - (bc >> 3); bc |= 56; } \
                                     else while (bc <= 56 && q < qe) { bb |= (uint64_t)*q++ << bc; bc += 8; } } while (0)
        for (;;) {
            if (!EXACT && out >= out_stop) break;                              // window full: the caller reads, slides, and comes back
            WQ_INF_REFILL();
            uint32_t e = LT[bb & lmask];
            if (!EXACT && __builtin_expect((e & 0xE0) == (K_BASE << 5) && bc >= 56, 1)) {
                // a match with its length and distance codes in the main tables and a distance of 8 or more: the bulk of FASTQ text (sequence
                // lines come out as matches of 4-10 bases against the last 32 KB).  Nothing is committed until all of that is known; anything
                // else falls through to the general path below with the reader untouched.
                const uint32_t nb = e & 31, lcl = (e >> 8) & 15;
                const uint64_t b2 = bb >> nb;
                const uint32_t d = DT[b2 & dmask];
                const uint32_t dn = d & 31, dcl = (d >> 8) & 15;
                const uint32_t dd = (d >> 12) + (uint32_t)((b2 >> dcl) & ((1u << (dn - dcl)) - 1));
                if (__builtin_expect((d & 0xE0) == (K_BASE << 5) && dd >= 8 && dd <= (size_t)(out - lo), 1)) {
                    const uint32_t len = (e >> 12) + (uint32_t)((bb >> lcl) & ((1u << (nb - lcl)) - 1));
                    bb = b2 >> dn; bc -= nb + dn;                              // at most 20 + 28 of the 56 bits
                    const uint8_t* src = out - dd;
                    store64(out, load64(src)); store64(out + 8, load64(src + 8));
                    if (__builtin_expect(len > 16, 0)) { uint8_t* dst = out + 16; src += 16; do { store64(dst, load64(src)); dst += 8; src += 8; } while (dst < out + len); }
                    out += len;
                    continue;
                }
            }
            if (!EXACT && ((e & 0xE0) == 0 || (e & 0xE0) == (K_LIT2 << 5)) && (e & 31) <= bc) {      // one or two literals of the main table, the common case
                // up to four lookups from the same refill (56 bits were there when input remains; each needs at most 11)
                for (int k = 0;; k++) {
                    bb >>= (e & 31); bc -= (e & 31);
                    out[0] = (uint8_t)(e >> 12); out[1] = (uint8_t)(e >> 20);  // the second byte is overwritten by the next symbol when there was one literal
                    out += 1 + ((e >> 5) & 7) / K_LIT2;
                    if (k == 3) break;
                    e = LT[bb & lmask];
                    if (((e & 0xE0) != 0 && (e & 0xE0) != (K_LIT2 << 5)) || (e & 31) > bc) break;      // anything else goes through the main path after a refill
                }
                continue;
            }
            if (((e >> 5) & 7) == K_SUB) {
                if (bc < LIT_BITS) { bad = "truncated input"; break; }
                const uint32_t sb = (e >> 8) & 15;
                e = LT[(e >> 12) + ((bb >> LIT_BITS) & ((1u << sb) - 1))];
                bb >>= LIT_BITS; bc -= LIT_BITS;
            }
            const uint32_t nb = e & 31, kind = (e >> 5) & 7;
            if (nb > bc) { bad = "truncated input"; break; }
            const uint64_t lsaved = bb;
            bb >>= nb; bc -= nb;                                               // past the code and its extra bits in one step
            if (kind == K_LIT || kind == K_LIT2) {
                const uint32_t nlit = kind == K_LIT2 ? 2u : 1u;
                if (EXACT && nlit > (size_t)(out_stop - out)) { bad = "more output than announced"; break; }
                *out++ = (uint8_t)(e >> 12);
                if (nlit == 2) *out++ = (uint8_t)(e >> 20);
                continue;
            }
            if (kind == K_EOB) { state = last_block ? S_TRAILER : S_BLOCK; break; }
            if (kind != K_BASE) { bad = "invalid literal / length code"; break; }
            const uint32_t lcl = (e >> 8) & 15;
            const uint32_t len = (e >> 12) + (uint32_t)((lsaved >> lcl) & ((1u << (nb - lcl)) - 1));
            if (bc < 15 + 13) WQ_INF_REFILL();                                 // only near the end of the input (a full refill covers the pair)
            uint32_t d = DT[bb & dmask];
            if (((d >> 5) & 7) == K_SUB) {
                if (bc < DIST_BITS) { bad = "truncated input"; break; }
                const uint32_t sb = (d >> 8) & 15;
                d = DT[(d >> 12) + ((bb >> DIST_BITS) & ((1u << sb) - 1))];
                bb >>= DIST_BITS; bc -= DIST_BITS;
            }
            const uint32_t dn = d & 31, dcl = (d >> 8) & 15;
            if (((d >> 5) & 7) != K_BASE) { bad = "invalid distance code"; break; }
            if (dn > bc) { bad = "truncated input"; break; }
            const uint32_t dd = (d >> 12) + (uint32_t)((bb >> dcl) & ((1u << (dn - dcl)) - 1));
            bb >>= dn; bc -= dn;
            if (dd > (size_t)(out - lo)) { bad = "distance beyond the start of the member"; break; }
            if (EXACT) {
                if (len > (size_t)(out_stop - out)) { bad = "more output than announced"; break; }
                const uint8_t* sx = out - dd;
                for (uint32_t k = 0; k < len; k++) out[k] = sx[k];
                out += len;
                continue;
            }
            const uint8_t* src = out - dd;
            uint8_t* dst = out;
            out += len;                                                        // len <= 258 and out < out_stop before: inside the slack
            if (dd >= 8) {
                // 8 bytes at a time is right for any distance >= 8 (a chunk only reads what earlier chunks or earlier output wrote); the first two
                // chunks go unconditionally -- most matches of FASTQ text are shorter than 16 and a loop exit per match is a mispredicted branch
                store64(dst, load64(src)); store64(dst + 8, load64(src + 8));
                if (len > 16) { dst += 16; src += 16; do { store64(dst, load64(src)); dst += 8; src += 8; } while (dst < out); }
            }
            else if (dd == 1) { memset(dst, *src, len); }
            else { do { *dst++ = *src++; } while (dst < out); }
        }
        #undef WQ_INF_REFILL
        bitbuf = bb; bitcnt = bc; p = q;
        out_pos = (size_t)(out - base);
        if (bad) fail(bad);
    }
};
