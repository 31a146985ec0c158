// Host-side suffix-array builder used only to fabricate synthetic GraphLib bundles
// (the reference builds its FM index with FMIndexes/SuffixArrays at src/index.jl:97; index
// construction is outside the accelerated path — this file exists so benchmarks and tests can
// make GENCODE-scale inputs without Julia).
//
// Order: plain suffix order with the implicit terminator smaller than every symbol
// ("shorter is smaller"), which is what the fixture test/test_index.jls was verified to hold.
//
// Method: 64-bit key of the first 32 symbols, bucket by the top 16 bits, std::sort inside
// buckets (parallel over buckets), equal-key groups resolved by word-wise suffix comparison.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>

namespace {

struct Text {
    const uint8_t* c;
    uint64_t n;
    std::vector<uint64_t> packed;   // 32 symbols per word, first symbol in the top bits
    void pack() {
        uint64_t nw = n / 32 + 2;
        packed.assign(nw, 0);
        for (uint64_t i = 0; i < n; i++) packed[i >> 5] |= (uint64_t)(c[i] & 3) << (62 - 2 * (i & 31));
    }
    // 32 symbols starting at i (zero padded past the end)
    inline uint64_t key(uint64_t i) const {
        uint64_t w = i >> 5, s = 2 * (i & 31);
        uint64_t k = packed[w] << s;
        if (s) k |= packed[w + 1] >> (64 - s);
        return k;
    }
    // true if suffix a < suffix b, both known equal on their first `skip` symbols
    bool less(uint64_t a, uint64_t b, uint64_t skip) const {
        uint64_t la = n - a, lb = n - b, m = std::min(la, lb);
        uint64_t off = skip;
        while (off < m) {
            uint64_t ka = key(a + off), kb = key(b + off);
            if (ka != kb) {
                uint64_t rem = m - off;
                if (rem >= 32) return ka < kb;
                uint64_t sh = 64 - 2 * rem;
                ka >>= sh; kb >>= sh;
                if (ka != kb) return ka < kb;
                break;
            }
            off += 32;
        }
        return la < lb;
    }
};

struct KP { uint64_t key; uint32_t pos; };

}  // namespace

extern "C" int wq_build_sa(const uint8_t* codes, uint64_t n, uint32_t* sa, int threads) {
    if (n == 0) return 0;
    if (n >= 0xFFFFFFFFull) return -1;
    Text T{codes, n, {}};
    T.pack();
    const int B = 16;
    const uint64_t NB = 1ull << B;
    std::vector<uint64_t> start(NB + 1, 0);
    for (uint64_t i = 0; i < n; i++) start[(T.key(i) >> (64 - B)) + 1]++;
    for (uint64_t b = 0; b < NB; b++) start[b + 1] += start[b];
    {
        std::vector<uint64_t> cur(start.begin(), start.end() - 1);
        for (uint64_t i = 0; i < n; i++) sa[cur[T.key(i) >> (64 - B)]++] = (uint32_t)i;
    }
    if (threads < 1) threads = 1;
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        std::vector<KP> buf;
        for (;;) {
            uint64_t b = next.fetch_add(64);
            if (b >= NB) break;
            for (uint64_t bb = b; bb < std::min(b + 64, NB); bb++) {
                uint64_t lo = start[bb], hi = start[bb + 1];
                if (hi - lo < 2) continue;
                buf.resize(hi - lo);
                for (uint64_t i = lo; i < hi; i++) buf[i - lo] = KP{T.key(sa[i]), sa[i]};
                std::sort(buf.begin(), buf.end(), [](const KP& x, const KP& y) { return x.key < y.key; });
                size_t i = 0;
                while (i < buf.size()) {
                    size_t j = i + 1;
                    while (j < buf.size() && buf[j].key == buf[i].key) j++;
                    if (j - i > 1)
                        std::sort(buf.begin() + i, buf.begin() + j,
                                  [&](const KP& x, const KP& y) { return T.less(x.pos, y.pos, 0); });
                    i = j;
                }
                for (uint64_t i2 = lo; i2 < hi; i2++) sa[i2] = buf[i2 - lo].pos;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return 0;
}

// BWT without the `$` row + sentinel row (1-based, in the (n+1)-row matrix), FMIndex field layout.
extern "C" int wq_build_bwt(const uint8_t* codes, uint64_t n, const uint32_t* sa, uint8_t* bwt, uint64_t* sentinel) {
    if (n == 0) return 0;
    uint64_t w = 0;
    bwt[w++] = codes[n - 1];           // row 1 is the `$` suffix
    for (uint64_t i = 0; i < n; i++) {
        if (sa[i] == 0) { *sentinel = i + 2; continue; }
        bwt[w++] = codes[sa[i] - 1];
    }
    return w == n ? 0 : -2;
}
