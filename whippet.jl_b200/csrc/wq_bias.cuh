// --biascorrect on the device: JointBiasMod / JointBiasCounter (src/bias.jl:103-270) as the reference's CLI uses them
// (bin/whippet-quant.jl:116-122 selects the types, :156-158 primer_normalize! / primer_adjust!, :167-169 gc_adjust!;
// src/reads.jl:59-66, 108-118 `biasval = count!(mod, read.sequence)` then push! into the counter the read is counted in).
//
// The reference keeps, per counter (every node, edge, circular edge, long path and multi-map class), a Dict hexamer -> Int32 and
// 21 GC-bin tallies, and turns them into a Float64 count after all reads are in:
//     primer_adjust!   count = sum over the counter's hexamers k of (back[k] / fore[k]) * n_k, divided by forenpos = 2
//     gc_adjust!       count *= (sum over bins i of value!(mod, i) * gc_i) / sum(gc_i)        (when the counter saw any window)
// with fore / back the model's normalised hexamer frequencies at read positions 1-2 / 23-32.  Everything that enters those sums is
// an INTEGER tally until the very end, so the device keeps integers and no float atomics are needed:
//   * model tables: u64[4096] x 2, one atomicAdd per hexamer occurrence;
//   * per read, one 64-bit record per (counter, hexamer) and per (counter, GC bin): [kind:2 | slot:32 | symbol:13], appended to a
//     record buffer by wq_k_bias / wq_k_bias_pe, which run after the chunk's counting kernels and look the read's counter up in the
//     count tables (its slot is the record's identity);
//   * at the end of the job the records are radix-sorted and run-length encoded (CUB), which yields every counter's tallies in
//     ascending symbol order: the declared canonical order of the reference's Dict iteration;
//   * wq_k_bias_adjust walks each counter's run once, sequentially (the order of the floating-point additions is the order of
//     the run), and writes the primer-adjusted count and the GC factor of every slot.
// Reference behaviours reproduced on purpose: in the paired form the mate's GC bins OVERWRITE the first entries of the read's
// (gc_count! restarts idx at 1, src/bias.jl:169-181, 194-203); value!(mod, bin::Int) reads the HEXAMER tables at index `bin`
// (src/bias.jl:206), so gcfore / gcback and gc_normalize! (src/quant.jl:223-238) never influence an output and are not computed;
// Julia's float `div` is exact floor division (div(1.0, 0.05) == 19.0; the reference's own test length(ExpectedGC(seq)) == 20,
// test/runtests.jl:101, relies on it), so the GC bin of k G/C bases in a 50-base window comes from a host-computed table.
// A mapped read shorter than 37 bases makes the reference throw (BoundsError in primer_count!); here the job fails with an error.
#pragma once
#include <cmath>
#ifdef __CUDACC__
#include <cub/device/device_run_length_encode.cuh>
#endif
// IEEE double operations that must not be fused or reassociated: intrinsics in device code, plain operators on the host (the test
// harness builds with -ffp-contract=off, the library with -fmad=false)
#ifdef __CUDA_ARCH__
#define WQ_DADD(a, b) __dadd_rn(a, b)
#define WQ_DMUL(a, b) __dmul_rn(a, b)
#define WQ_DDIV(a, b) __ddiv_rn(a, b)
#else
#define WQ_DADD(a, b) ((a) + (b))
#define WQ_DMUL(a, b) ((a) * (b))
#define WQ_DDIV(a, b) ((a) / (b))
#endif

#define WQ_BIAS_KSIZE 6
#define WQ_BIAS_NHEX 4096
#define WQ_BIAS_FORE_OFF 1
#define WQ_BIAS_FORE_N 2
#define WQ_BIAS_BACK_OFF 23
#define WQ_BIAS_BACK_N 10
#define WQ_BIAS_GC_SIZE 50
#define WQ_BIAS_GC_OFF 5
#define WQ_BIAS_GC_INC 10
#define WQ_BIAS_GC_BINS 21
#define WQ_BIAS_MIN_LEN (WQ_BIAS_BACK_OFF + WQ_BIAS_BACK_N - 1 + WQ_BIAS_KSIZE - 1)      // 37
#define WQ_BIAS_MAX_WIN 21                                                              // windows of a 255-base read
#define WQ_BIAS_MAX_REC (WQ_BIAS_FORE_N + WQ_BIAS_MAX_WIN)
#define WQ_BIAS_SYM_GC 4096u                                                            // symbol of GC bin b (1-based): 4096 + b

struct WqBiasDev {
    uint64_t* fore;            // [4096] tallies (every occurrence adds 1.0 in the reference: exact as integers)
    uint64_t* back;            // [4096]
    uint64_t* rec;             // records
    uint64_t* cursor;          // [0] records written, [1] mapped reads shorter than WQ_BIAS_MIN_LEN
    uint64_t  cap;
    uint8_t   gcbin[WQ_BIAS_GC_SIZE + 1];      // bin (1-based) of a window with k G/C bases
};

WQ_HD uint64_t wq_bias_rec(uint32_t kind, uint32_t slot, uint32_t sym) { return ((uint64_t)kind << 45) | ((uint64_t)slot << 13) | (uint64_t)sym; }

// hexamers / GC bins of one read in the orientation ungapped_align left it in (w = its 2-bit words, already reverse-complemented when
// the seed was on the other strand).  kmer_index_trailing packs the first base into the most significant bits (src/sgkmer.jl:17-28).
WQ_HD uint32_t wq_bias_hex(const uint64_t* w, int i1) {           // i1: 1-based start
    uint32_t x = 0;
    for (int k = 0; k < WQ_BIAS_KSIZE; k++) { const int j = i1 - 1 + k; x = (x << 2) | ((uint32_t)(w[j >> 5] >> (2 * (j & 31))) & 3u); }
    return x;
}
WQ_HD int wq_bias_gc(const WqBiasDev& B, const uint64_t* w, int len, uint8_t* bins) {      // gc_count!: returns the number of windows
    int n = 0;
    for (int i = WQ_BIAS_GC_OFF; i <= len - WQ_BIAS_GC_SIZE; i += WQ_BIAS_GC_INC) {
        int gc = 0;
        for (int k = 0; k < WQ_BIAS_GC_SIZE; k++) { const int j = i - 1 + k; const uint32_t c = (uint32_t)(w[j >> 5] >> (2 * (j & 31))) & 3u; gc += (c == 1u || c == 2u); }
        bins[n++] = B.gcbin[gc];
    }
    return n;
}

WQ_HD void wq_bias_load(const WqBatchDev& b, uint32_t r, bool flip, uint64_t* w, int& len) {
    len = WQ_LDG(b.len + r);
    const uint64_t* src = b.seq2 + WQ_LDG(b.seq_off + r);
    const int nw = (len + 31) >> 5;
    for (int i = 0; i <= WQ_MAX_READ_WORDS; i++) {
        uint64_t v = i < nw ? WQ_LDG(src + i) : 0ull;
        if (i == nw - 1 && (len & 31)) v &= (~0ull >> (64 - 2 * (len & 31)));
        w[i] = v;
    }
    if (flip) wq_revcomp(w, len);
}

// one mapped read (pair): model tallies, then the records of the counter it was counted in
WQ_HD void wq_bias_read(const WqDevIndex& ix, const WqWork& wk, const WqQuantDev& q, const WqBiasDev& B, const WqTarget& t, const uint32_t* kw,
                        const uint64_t* w, int len, const uint64_t* w2, int len2) {
    if (len < WQ_BIAS_MIN_LEN) { WQ_ATOMIC_ADD_U64(&B.cursor[1], 1); return; }
    uint32_t hex[WQ_BIAS_FORE_N];
    for (int i = 0; i < WQ_BIAS_FORE_N; i++) { hex[i] = wq_bias_hex(w, WQ_BIAS_FORE_OFF + i); WQ_ATOMIC_ADD_U64(&B.fore[hex[i]], 1); }
    for (int i = 0; i < WQ_BIAS_BACK_N; i++) WQ_ATOMIC_ADD_U64(&B.back[wq_bias_hex(w, WQ_BIAS_BACK_OFF + i)], 1);
    uint8_t bins[WQ_BIAS_MAX_WIN], bins2[WQ_BIAS_MAX_WIN];
    int nb = wq_bias_gc(B, w, len, bins);
    if (w2) {                                       // the mate's bins overwrite the first entries (idx restarts at 1)
        const int nb2 = wq_bias_gc(B, w2, len2, bins2);
        for (int i = 0; i < nb2; i++) bins[i] = bins2[i];
        if (nb2 > nb) nb = nb2;
    }
    if (t.kind < 0) return;                         // the model saw the read; no counter did (count! returned early)
    int64_t slot = -1;
    if (t.kind == 0) slot = (int64_t)t.node_idx;
    else if (t.kind == 1) slot = wq_edge_find(q.edges, t.ekey);
    else slot = wq_keyed_find(q.keyed, kw, t.nw);
    if (slot < 0) return;                           // the store overflowed: reported by the count tables' own counters
    const uint64_t at = WQ_ATOMIC_ADD_U64(&B.cursor[0], (uint64_t)(WQ_BIAS_FORE_N + nb));
    if (at + WQ_BIAS_FORE_N + nb > B.cap) return;   // cannot happen: the host sizes the buffer for the worst case of the chunk
    for (int i = 0; i < WQ_BIAS_FORE_N; i++) B.rec[at + i] = wq_bias_rec((uint32_t)t.kind, (uint32_t)slot, hex[i]);
    for (int i = 0; i < nb; i++) B.rec[at + WQ_BIAS_FORE_N + i] = wq_bias_rec((uint32_t)t.kind, (uint32_t)slot, WQ_BIAS_SYM_GC + bins[i]);
}

WQ_HDN void wq_bias_thread(const WqDevIndex& ix, const WqBatchDev& b, const WqWork& wk, const WqQuantDev& q, const WqBiasDev& B, uint32_t r) {
    const WqSeed s = wk.seeds[r];
    int nkept = 0;
    for (int k = 0; k < s.cnt; k++) nkept += (wk.hits[s.hit_base + k].flags & 4) ? 1 : 0;
    if (nkept == 0) return;                         // isnull(align)
    uint64_t w[WQ_MAX_READ_WORDS + 1];
    int len;
    wq_bias_load(b, r, !s.strand, w, len);
    uint32_t kw[WQ_KEY_WORDS_MAX];
    const WqTarget t = wq_target_se(ix, wk, s.hit_base, s.cnt, nkept, kw);
    wq_bias_read(ix, wk, q, B, t, kw, w, len, nullptr, 0);
}
WQ_HDN void wq_bias_pe_thread(const WqDevIndex& ix, const WqBatchDev& bf, const WqBatchDev& br, const WqWork& wk, const WqQuantDev& q,
                              const WqBiasDev& B, uint32_t r) {
    const WqSeed s = wk.seeds[r];
    int nkept = 0;
    for (int k = 0; k < s.cnt; k++) nkept += (wk.hits[2 * (s.hit_base + k)].flags & 4) ? 1 : 0;
    if (nkept == 0) return;
    uint64_t w[WQ_MAX_READ_WORDS + 1], w2[WQ_MAX_READ_WORDS + 1];
    int len, len2;
    wq_bias_load(bf, r, !s.strand, w, len);
    wq_bias_load(br, r, (s.sp & 0x100u) != 0, w2, len2);
    uint32_t kw[WQ_KEY_WORDS_MAX_PE];
    const WqTarget t = wq_target_pe(ix, wk, s.hit_base, s.cnt, nkept, kw);
    wq_bias_read(ix, wk, q, B, t, kw, w, len, w2, len2);
}

// One counter's run of (symbol, tally) pairs, ascending, starting at unique record i: the count after primer_adjust! and the factor
// of gc_adjust! (-1: the counter saw no window).  The additions happen in the order of the run.
WQ_HD void wq_bias_adjust_run(const uint64_t* ukey, const uint32_t* ucnt, uint32_t n, uint32_t i, const double* ratio, const double* gcval,
                              double& count_out, double& g_out) {
    const uint64_t id = ukey[i] >> 13;
    double count = 0.0, weight = 0.0;                      // JointBiasCounter().count == 0.0
    long long gctotal = 0;
    for (uint32_t j = i; j < n && (ukey[j] >> 13) == id; j++) {
        const uint32_t sym = (uint32_t)(ukey[j] & 0x1FFFu);
        const double v = (double)ucnt[j];
        if (sym < WQ_BIAS_SYM_GC) count = WQ_DADD(count, WQ_DMUL(ratio[sym], v));                            // cnt.count += value!(mod, k) * v
        else { weight = WQ_DADD(weight, WQ_DMUL(gcval[sym - WQ_BIAS_SYM_GC - 1], v)); gctotal += ucnt[j]; }     // weight += value!(mod, i) * cnt.gc[i]
    }
    count_out = WQ_DDIV(count, (double)WQ_BIAS_FORE_N);                                                       // cnt.count /= mod.forenpos
    g_out = gctotal > 0 ? WQ_DDIV(weight, (double)gctotal) : -1.0;
}
// ratio[k] = back[k] / fore[k] over the NORMALISED tables (primer_normalize!, src/bias.jl:145-152, then value!(mod, hept), :205);
// gcval[i - 1] = value!(mod, i::Int) for i = 1..21: the same tables read at index i (src/bias.jl:206)
WQ_HD void wq_bias_model_entry(const uint64_t* fore, const uint64_t* back, double foresum, double backsum, int k, double* ratio, double* gcval,
                               double* fore_n, double* back_n) {
    const double fn = WQ_DDIV((double)fore[k], foresum), bn = WQ_DDIV((double)back[k], backsum);
    fore_n[k] = fn; back_n[k] = bn;
    ratio[k] = WQ_DDIV(bn, fn);
    if (k < WQ_BIAS_GC_BINS) gcval[k] = fn > 0.0 ? WQ_DDIV(bn, fn) : 1.0;
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(128) wq_k_bias(const __grid_constant__ WqDevIndex ix, const WqBatchDev b, const WqWork wk, const WqQuantDev q,
                                                 const __grid_constant__ WqBiasDev B) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < b.n) wq_bias_thread(ix, b, wk, q, B, r);
}
__global__ void __launch_bounds__(128) wq_k_bias_pe(const __grid_constant__ WqDevIndex ix, const WqBatchDev bf, const WqBatchDev br, const WqWork wk,
                                                    const WqQuantDev q, const __grid_constant__ WqBiasDev B) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < bf.n) wq_bias_pe_thread(ix, bf, br, wk, q, B, r);
}

// foresum / backsum are sums of integers, exact in any order
__global__ void wq_k_bias_model(const uint64_t* fore, const uint64_t* back, double* ratio, double* gcval, double* fore_n, double* back_n) {
    __shared__ unsigned long long s_f, s_b;
    if (threadIdx.x == 0) { s_f = 0; s_b = 0; }
    __syncthreads();
    unsigned long long f = 0, bk = 0;
    for (int k = threadIdx.x; k < WQ_BIAS_NHEX; k += blockDim.x) { f += fore[k]; bk += back[k]; }
    atomicAdd(&s_f, f); atomicAdd(&s_b, bk);
    __syncthreads();
    const double foresum = (double)s_f, backsum = (double)s_b;
    for (int k = threadIdx.x; k < WQ_BIAS_NHEX; k += blockDim.x) wq_bias_model_entry(fore, back, foresum, backsum, k, ratio, gcval, fore_n, back_n);
}

// every counter's run of (symbol, tally) pairs, ascending: thread i takes the run that STARTS at unique record i
__global__ void __launch_bounds__(256) wq_k_bias_adjust(const uint64_t* __restrict__ ukey, const uint32_t* __restrict__ ucnt, const uint32_t* __restrict__ n_unique,
                                                        const double* __restrict__ ratio, const double* __restrict__ gcval,
                                                        double* node_p, double* node_g, double* edge_p, double* edge_g, double* keyed_p, double* keyed_g) {
    const uint32_t n = *n_unique;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t id = ukey[i] >> 13;
        if (i > 0 && (ukey[i - 1] >> 13) == id) continue;
        double count, g;
        wq_bias_adjust_run(ukey, ucnt, n, i, ratio, gcval, count, g);
        const uint32_t kind = (uint32_t)(id >> 32), slot = (uint32_t)id;
        double* P = kind == 0 ? node_p : kind == 1 ? edge_p : keyed_p;
        double* Gf = kind == 0 ? node_g : kind == 1 ? edge_g : keyed_g;
        P[slot] = count;
        Gf[slot] = g;                                      // -1 marks "no window seen": gc_adjust! leaves such a counter alone
    }
}
// compacted (slot) lists -> values in that order
__global__ void wq_k_bias_gather(const uint32_t* __restrict__ slots, uint32_t n, const double* __restrict__ P, const double* __restrict__ Gf, double* op, double* og) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { op[i] = P[slots[i]]; og[i] = Gf[slots[i]]; }
}
__global__ void wq_k_fill_f64(double* p, double v, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
#endif

// the GC-bin table with Julia's float division: Int(div(k / 50, 0.05) + 1), div(x, y) = round((x - rem(x, y)) / y)
static inline void wq_bias_gc_table(uint8_t* out) {
    for (int k = 0; k <= WQ_BIAS_GC_SIZE; k++) {
        const double x = (double)k / (double)WQ_BIAS_GC_SIZE, y = 0.05;
        out[k] = (uint8_t)((long)std::nearbyint((x - std::fmod(x, y)) / y) + 1);
    }
}
