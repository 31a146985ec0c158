// Mutation fuzzing of csrc/wq_inflate.h under AddressSanitizer + UndefinedBehaviorSanitizer (TEST INFRASTRUCTURE): valid gzip files given on the
// command line are truncated, bit-flipped, overwritten, spliced and cut, then inflated through read() in pieces of random size and through
// raw_into() with a random announced length and a canary behind it.  Build and run (from scripts/):
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=undefined -o /tmp/fuzz_inflate fuzz_inflate.cpp -lz
//   ASAN_OPTIONS=detect_leaks=0 /tmp/fuzz_inflate a.gz b.gz ... SECONDS
// Last run of round 2: 286 574 mutated streams (286 038 ended in an error message) + 286 540 raw runs in 150 s, no report, no canary touched.
#include "../whippet.jl_b200/csrc/wq_inflate.h"
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iterator>
#include <random>
int main(int argc, char** argv) {
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 1; i < argc - 1; i++) { std::ifstream f(argv[i], std::ios::binary); seeds.emplace_back((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()); if (seeds.back().size() > 300000) seeds.back().resize(300000); }
    const double secs = atof(argv[argc - 1]);
    std::mt19937_64 rng(12345);
    std::vector<uint8_t> out(8u << 20);
    unsigned long long runs = 0, errs = 0, raw_runs = 0;
    WqInflate* z = new WqInflate();      // (not freed: the process ends)
    auto t0 = std::chrono::steady_clock::now();
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < secs) {
        std::vector<uint8_t> d = seeds[rng() % seeds.size()];
        const int kind = rng() % 6;
        if (kind == 0) d.resize(rng() % (d.size() + 1));
        else if (kind == 1) for (int k = 0, n = 1 + rng() % 8; k < n; k++) d[rng() % d.size()] ^= (uint8_t)(1u << (rng() % 8));
        else if (kind == 2) for (int k = 0, n = 1 + rng() % 64; k < n; k++) d[rng() % d.size()] = (uint8_t)rng();
        else if (kind == 3) { size_t a = rng() % d.size(), b = rng() % d.size(), n = rng() % 200; for (size_t k = 0; k < n && a + k < d.size() && b + k < d.size(); k++) d[a + k] = d[b + k]; }
        else if (kind == 4) { size_t a = 10 + rng() % 64; for (size_t k = a; k < d.size() && k < a + 40; k++) d[k] = (uint8_t)rng(); }      // damage the first block header
        else { size_t a = rng() % d.size(); d.erase(d.begin() + a, d.begin() + std::min(d.size(), a + 1 + rng() % 50)); }
        z->open(d.data(), d.size());
        size_t got = 0;
        for (;;) { const size_t want = 1 + rng() % (1u << 20); if (got + want > out.size()) break; const size_t r = z->read(out.data() + got, want); got += r; if (r < want) break; }
        runs++; errs += z->failed();
        // raw form on the deflate payload behind the 10-byte header, with a random announced length and canary
        if (d.size() > 30) {
            const size_t len = rng() % 70000;
            memset(out.data() + len, 0x5A, 64);
            z->raw_into(d.data() + 10, d.size() - 10, out.data(), len);
            for (int k = 0; k < 64; k++) if (out[len + k] != 0x5A) { printf("CANARY OVERWRITTEN\n"); return 1; }
            raw_runs++;
        }
    }
    printf("runs %llu (errors %llu), raw runs %llu: no crash\n", runs, errs, raw_runs);
}
