// Per-thread bodies of the alignment kernels (one grid thread = one call).  They are written as
// __host__ __device__ functions so that tests/emu can drive the very same code single-threaded on
// a CPU while no GPU is attached; the shipped library only ever launches them as CUDA kernels.
//
//   wq_seed_thread     kernel 1: one read  -> seed_locate + locate of every SA row of the seed
//   wq_extend_thread   kernel 2: one hit   -> _ungapped_align (fwd + rev extension)
//   wq_resolve_thread  kernel 3: one read  -> score filter of ungapped_align (src/align.jl:239-268),
//                                count! (src/quant.jl:556-594) / push!(multi) (src/quant.jl:456-469)
#pragma once
#include "wq_align.cuh"

#if defined(__CUDA_ARCH__)
#define WQ_ATOMIC_ADD_U32(p, v) atomicAdd((unsigned int*)(p), (unsigned int)(v))
#define WQ_ATOMIC_ADD_U64(p, v) atomicAdd((unsigned long long*)(p), (unsigned long long)(v))
#define WQ_ATOMIC_CAS_U64(p, c, v) atomicCAS((unsigned long long*)(p), (unsigned long long)(c), (unsigned long long)(v))
#define WQ_ATOMIC_EXCH_U32(p, v) atomicExch((unsigned int*)(p), (unsigned int)(v))
#define WQ_FENCE() __threadfence()
#define WQ_VLOAD_U32(p) (*(volatile const uint32_t*)(p))
#define WQ_VLOAD_U64(p) (*(volatile const uint64_t*)(p))
// Reserve `cnt` (< 32) consecutive slots behind a global cursor with ONE atomic per warp: the lanes that are here
// together sum their requests with ballots (one per bit of cnt), the first of them moves the cursor.  A cursor that
// every thread of a 10-million-read batch bumps on its own would serialise the whole kernel on one L2 atomic unit.
__device__ __forceinline__ uint32_t wq_warp_alloc(uint32_t* cursor, uint32_t cnt) {
    const unsigned m = __activemask();
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t prefix = 0, total = 0;
#pragma unroll
    for (int b = 0; b < 5; b++) {
        const unsigned bal = __ballot_sync(m, (cnt >> b) & 1u);
        prefix += (uint32_t)__popc(bal & lt) << b;
        total += (uint32_t)__popc(bal) << b;
    }
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader && total) base = atomicAdd((unsigned int*)cursor, total);
    base = __shfl_sync(m, base, leader);
    return base + prefix;
}
#define WQ_ALLOC(p, v) wq_warp_alloc((uint32_t*)(p), (uint32_t)(v))
#else
template <class T, class V> static inline T wq_emu_add(T* p, V v) { T o = *p; *p = (T)(o + (T)v); return o; }
static inline uint64_t wq_emu_cas(uint64_t* p, uint64_t c, uint64_t v) { uint64_t o = *p; if (o == c) *p = v; return o; }
static inline uint32_t wq_emu_exch(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = v; return o; }
#define WQ_ATOMIC_ADD_U32(p, v) wq_emu_add((uint32_t*)(p), (v))
#define WQ_ATOMIC_ADD_U64(p, v) wq_emu_add((uint64_t*)(p), (v))
#define WQ_ATOMIC_CAS_U64(p, c, v) wq_emu_cas((uint64_t*)(p), (c), (v))
#define WQ_ATOMIC_EXCH_U32(p, v) wq_emu_exch((uint32_t*)(p), (v))
#define WQ_ALLOC(p, v) wq_emu_add((uint32_t*)(p), (v))
#define WQ_FENCE() ((void)0)
#define WQ_VLOAD_U32(p) (*(const uint32_t*)(p))
#define WQ_VLOAD_U64(p) (*(const uint64_t*)(p))
#endif

struct WqBatchDev {                // a read batch resident in device memory (wq_read_batch with device pointers)
    uint32_t n;
    const uint8_t*  len;
    const uint64_t* seq_off;
    const uint64_t* seq2;
    const uint64_t* qual_off;
    const uint8_t*  qual;
};

// ---- count stores -------------------------------------------------------------------------------
#define WQ_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define WQ_KEYOFF_UNSET 0xFFFFFFFFu
#define WQ_KEYOFF_FULL 0xFFFFFFFEu

struct WqEdgeTable {               // SpliceGraphQuant.edge + .circ: key = gene<<32 | lnode<<16 | rnode
    uint64_t* key;
    uint64_t* count;
    uint64_t  mask;                // slots - 1
};
struct WqKeyedTable {              // SpliceGraphQuant.long + MultiMapping.map: exact variable-length keys
    uint64_t* hash;                // 0 = empty
    uint32_t* keyoff;              // word offset of [nwords, words...] in pool
    uint64_t* count;
    uint32_t* pool;
    uint64_t* pool_cursor;
    uint64_t  pool_cap;
    uint64_t  mask;
};
struct WqCounters {                // scalar stats, all u64 (kept right behind node_count so one allreduce covers both)
    uint64_t total, mapped, multi, sum_readlen, path_overflow, edge_full, keyed_full, pool_overflow;
};

struct WqWork {                    // per-batch scratch
    WqSeed*   seeds;               // [n]
    uint32_t* hit_read;            // [hit_cap]
    uint32_t* hit_s;               // [hit_cap] 1-based text position of the located seed
    uint32_t* hit_s2;              // [hit_cap] paired: text position of the mate-2 seed
    uint32_t* hit_gene;            // [hit_cap] paired: gene both mates are aligned against
    WqHit*    hits;                // [hit_cap]  (paired: [2*hit_cap], mate m of pair slot s at 2*s+m)
    wq_align_node* pool;           // [pool_cap]
    uint32_t* cursors;             // [0] hit cursor, [1] path-pool cursor, [2] pending keyed inserts, [5] hits of reverse-strand seeds (allocated from the back of the hit arrays)
    uint32_t* pending;             // [n] reads whose keyed insert must be retried (also: seed queue A)
    uint32_t* pending2;            // [n] seed queue B
    uint32_t  hit_cap, pool_cap;
    uint32_t* defer;               // [hit_cap] extension tasks the first-pass kernel hands to the second pass (count in cursors[6])
    uint32_t* mapped_bits;         // [(n + 31) / 32] bit r = read r was mapped (running mean_readlen of src/reads.jl:69); cursors[3] = max(255 - len),
                                   // cursors[4] = max(len), cursors[7] = mapped reads of the chunk
};

// Where a thread keeps its working alignment while extending (WqCtx.path / .side)
struct WqPathStore {
    WqPNode* path; int32_t pstride, pcap;
    WqPNode* side; int32_t sstride;          // NULL in the first-pass kernel
};
WQ_HD void wq_bind_store(WqCtx& c, const WqPathStore& st) {
    c.path = st.path; c.pstride = st.pstride; c.pcap = st.pcap; c.side = st.side; c.sstride = st.sstride;
}

// result of one extension -> hit header + path nodes in the pool
WQ_HD void wq_emit_hit(const WqCtx& c, const WqWork& wk, uint32_t hslot, float identity) {
    WqHit h;
    h.identity = identity;
    h.offset = c.offset;
    h.offsetread = c.offsetread;
    const int np = c.overflow ? 0 : wq_npath(c);
    h.npath = (uint8_t)np;
    h.strand = c.strand;
    h.flags = (uint8_t)((c.isvalid ? 1 : 0) | (wq_isvalid(c) ? 2 : 0) | (c.overflow ? 8 : 0));
    if (c.overflow) h.flags &= (uint8_t)~3;
    const uint32_t ps = WQ_ALLOC(&wk.cursors[1], np);
    if (np) {
        if (ps + np <= wk.pool_cap) {
            for (int i = 0; i < np; i++) {
                const WqPNode n = wq_pn(c, i);
                wq_align_node o;
                o.gene = c.gene;
                o.node = n.node;
                o.mistolerance = n.tol;
                o.matches = n.m;
                o.mismatches = n.mm;
                o._pad = 0;
                wk.pool[ps + i] = o;
            }
        } else {
            h.flags |= 16;           // path pool exhausted: reported through WqCounters.pool_overflow
        }
    }
    h.path_start = ps;
    wk.hits[hslot] = h;
}

WQ_HD void wq_load_read(WqCtx& c, const WqBatchDev& b, uint32_t r) {
    int len = WQ_LDG(b.len + r);
    c.readlen = len;
    const uint64_t* src = b.seq2 + WQ_LDG(b.seq_off + r);
    int nw = (len + 31) >> 5;
#pragma unroll
    for (int i = 0; i <= WQ_MAX_READ_WORDS; i++) {
        uint64_t v = i < nw ? WQ_LDG(src + i) : 0ull;
        if (i == nw - 1 && (len & 31)) v &= (~0ull >> (64 - 2 * (len & 31)));
        c.w[i] = v;
    }
    c.qual = b.qual + WQ_LDG(b.qual_off + r);
    c.flipped = 0;
    c.overflow = 0;
    c.depth = 0;
}

// bit j set when quality[j] <= seed_min_qual (orientation: as stored, or reversed when `rev`)
WQ_HD uint32_t wq_le4(uint32_t v, uint32_t lim) {     // per byte: v <= lim -> 0x01 in that byte
#if defined(__CUDA_ARCH__)
    return __vcmpleu4(v, lim) & 0x01010101u;
#else
    uint32_t m = 0;
    for (int i = 0; i < 4; i++) if (((v >> (8 * i)) & 0xFF) <= ((lim >> (8 * i)) & 0xFF)) m |= 1u << (8 * i);
    return m;
#endif
}
WQ_HD void wq_badq(const WqCtx& c, int min_qual, bool rev, uint64_t* badq) {
    uint64_t fw[4] = {0, 0, 0, 0};
    const int len = c.readlen;
    if (min_qual >= 0) {
        const uint32_t lim = (uint32_t)(min_qual > 255 ? 255 : min_qual) * 0x01010101u;
        int j = 0;
        if ((((uintptr_t)c.qual) & 3) == 0) {           // 4 phred bytes per load
            const uint32_t* q4 = reinterpret_cast<const uint32_t*>(c.qual);
            for (; j + 4 <= len; j += 4) {
                uint32_t m = wq_le4(WQ_LDG(q4 + (j >> 2)), lim);
                uint64_t bits = (uint64_t)((m * 0x01020408u) >> 24) & 0xF;
                fw[j >> 6] |= bits << (j & 63);
            }
        }
        for (; j < len; j++)
            if ((int)WQ_LDG(c.qual + j) <= min_qual) fw[j >> 6] |= 1ull << (j & 63);
    }
    if (!rev) { for (int i = 0; i < 4; i++) badq[i] = fw[i]; return; }
    // reversed orientation: bit k = fw bit (len-1-k)
    for (int i = 0; i < 4; i++) badq[i] = 0;
    for (int k = 0; k < len; k++) {
        int j = len - 1 - k;
        if ((fw[j >> 6] >> (j & 63)) & 1) badq[k >> 6] |= 1ull << (k & 63);
    }
}

// ---- kernel 1 -----------------------------------------------------------------------------------
WQ_HD void wq_seed_finish(const WqDevIndex& ix, const WqWork& wk, uint32_t r, int cnt, int64_t sp, int readloc, bool strand) {
    WqSeed s;
    s.sp = (uint32_t)sp;
    s.cnt = (uint8_t)cnt;
    s.readloc = (uint8_t)readloc;
    s.strand = strand ? 1 : 0;
    s._pad = 0;
    s.hit_base = 0;
    // hits of forward-strand seeds fill the hit arrays from the front, those of reverse-strand seeds from the back,
    // so the extension kernel's warps see one kind only (the read is reverse-complemented for the second kind)
    const uint32_t bf = WQ_ALLOC(&wk.cursors[0], strand ? cnt : 0);
    const uint32_t bb = WQ_ALLOC(&wk.cursors[5], strand ? 0 : cnt);
    if (cnt > 0) {
        const uint32_t base = strand ? bf : wk.hit_cap - bb - (uint32_t)cnt;
        s.hit_base = base;
        for (int k = 0; k < cnt; k++)
            if (base + k < wk.hit_cap) {    // hit_cap = reads x seed_tolerate, so this always holds; a guard against a mis-sized launch
                wk.hit_read[base + k] = r;
                wk.hit_s[base + k] = (uint32_t)(wq_sa_value(ix, sp + k) + 1);
            }
    }
    wk.seeds[r] = s;
}

WQ_HDN void wq_seed_thread(const WqDevIndex& ix, const WqParams& p, const WqBatchDev& b, const WqWork& wk, uint32_t r, uint64_t* wbuf) {
    WqCtx c;
    c.ix = &ix;
    c.p = &p;
    c.w = wbuf;
    wq_load_read(c, b, r);
    uint64_t badq[4];
    wq_badq(c, p.seed_min_qual, false, badq);
    int64_t sp = 2;
    int readloc = 0;
    bool strand = true;
    int cnt = wq_seed_locate(c, badq, true, !p.is_stranded, sp, readloc, strand);
    wq_seed_finish(ix, wk, r, cnt, sp, readloc, strand);
}

// ---- kernel 2 -----------------------------------------------------------------------------------
// `first_pass`: a task that overflows the small path store (or would have to park a candidate) is queued for the second pass,
// which runs the same code with the full WQ_MAX_PATH store and side buffers.
WQ_HDN void wq_extend_thread(const WqDevIndex& ix, const WqParams& p, const WqBatchDev& b, const WqWork& wk, uint32_t slot, uint64_t* wbuf,
                             const WqPathStore& st, bool first_pass) {
    uint32_t r = wk.hit_read[slot];
    WqSeed s = wk.seeds[r];
    WqCtx c;
    c.ix = &ix;
    c.p = &p;
    c.w = wbuf;
    wq_bind_store(c, st);
    wq_load_read(c, b, r);
    if (!s.strand) {                 // reverse_complement!(read), src/align.jl:231-234
        wq_revcomp(c.w, c.readlen);
        c.flipped = 1;
    }
    wq_align_at(c, (int64_t)wk.hit_s[slot], s.readloc, s.strand != 0, 0);
    if (c.overflow && first_pass) {
        wk.defer[WQ_ATOMIC_ADD_U32(&wk.cursors[6], 1)] = slot;
        return;
    }
    wq_emit_hit(c, wk, slot, wq_identity(c));
}

// ---- keyed store --------------------------------------------------------------------------------
WQ_HD uint64_t wq_hash_words(const uint32_t* w, int n) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)n;
    for (int i = 0; i < n; i++) {
        h ^= w[i];
        h *= 0xFF51AFD7ED558CCDull;
        h ^= h >> 29;
    }
    h ^= h >> 32;
    return h ? h : 1;
}

// returns 0 = counted, 1 = owner has not published its key yet (retry), 2 = table or pool full
WQ_HD int wq_keyed_insert(const WqKeyedTable& t, const uint32_t* kw, int nw, uint64_t amount) {
    uint64_t h = wq_hash_words(kw, nw);
    uint64_t slot = h & t.mask;
    for (uint64_t probe = 0; probe <= t.mask; probe++, slot = (slot + 1) & t.mask) {
        uint64_t cur = WQ_VLOAD_U64(&t.hash[slot]);
        if (cur == 0) {
            uint64_t prev = WQ_ATOMIC_CAS_U64(&t.hash[slot], 0ull, h);
            if (prev == 0) {          // this thread owns the slot: store the key, then publish it
                uint64_t off = WQ_ATOMIC_ADD_U64(t.pool_cursor, (uint64_t)nw + 1);
                if (off + nw + 1 > t.pool_cap || off + nw + 1 >= 0xFFFFFFF0ull) {
                    WQ_ATOMIC_EXCH_U32(&t.keyoff[slot], WQ_KEYOFF_FULL);
                    return 2;
                }
                t.pool[off] = (uint32_t)nw;
                for (int i = 0; i < nw; i++) t.pool[off + 1 + i] = kw[i];
                WQ_FENCE();
                WQ_ATOMIC_EXCH_U32(&t.keyoff[slot], (uint32_t)off);
                WQ_ATOMIC_ADD_U64(&t.count[slot], amount);
                return 0;
            }
            cur = prev;
        }
        if (cur == h) {
            uint32_t off = WQ_VLOAD_U32(&t.keyoff[slot]);
            if (off == WQ_KEYOFF_UNSET) return 1;
            if (off == WQ_KEYOFF_FULL) return 2;
            WQ_FENCE();
            bool same = WQ_VLOAD_U32(&t.pool[off]) == (uint32_t)nw;
            for (int i = 0; same && i < nw; i++) same = WQ_VLOAD_U32(&t.pool[off + 1 + i]) == kw[i];
            if (same) {
                WQ_ATOMIC_ADD_U64(&t.count[slot], amount);
                return 0;
            }
        }
    }
    return 2;
}

WQ_HD int wq_edge_insert(const WqEdgeTable& t, uint64_t key, uint64_t amount) {
    uint64_t h = key * 0x9E3779B97F4A7C15ull;
    uint64_t slot = (h >> 20) & t.mask;
    for (uint64_t probe = 0; probe <= t.mask; probe++, slot = (slot + 1) & t.mask) {
        uint64_t cur = WQ_VLOAD_U64(&t.key[slot]);
        if (cur == WQ_EMPTY_KEY) {
            uint64_t prev = WQ_ATOMIC_CAS_U64(&t.key[slot], WQ_EMPTY_KEY, key);
            cur = prev == WQ_EMPTY_KEY ? key : prev;
        }
        if (cur == key) {
            WQ_ATOMIC_ADD_U64(&t.count[slot], amount);
            return 0;
        }
    }
    return 2;
}

// read-only lookups of a counter's slot (after every insertion of the chunk has completed): -1 = not present
WQ_HD int64_t wq_edge_find(const WqEdgeTable& t, uint64_t key) {
    uint64_t h = key * 0x9E3779B97F4A7C15ull;
    uint64_t slot = (h >> 20) & t.mask;
    for (uint64_t probe = 0; probe <= t.mask; probe++, slot = (slot + 1) & t.mask) {
        const uint64_t cur = WQ_VLOAD_U64(&t.key[slot]);
        if (cur == key) return (int64_t)slot;
        if (cur == WQ_EMPTY_KEY) return -1;
    }
    return -1;
}
WQ_HD int64_t wq_keyed_find(const WqKeyedTable& t, const uint32_t* kw, int nw) {
    const uint64_t h = wq_hash_words(kw, nw);
    uint64_t slot = h & t.mask;
    for (uint64_t probe = 0; probe <= t.mask; probe++, slot = (slot + 1) & t.mask) {
        const uint64_t cur = WQ_VLOAD_U64(&t.hash[slot]);
        if (cur == 0) return -1;
        if (cur == h) {
            const uint32_t off = WQ_VLOAD_U32(&t.keyoff[slot]);
            if (off >= WQ_KEYOFF_FULL) continue;
            bool same = WQ_VLOAD_U32(&t.pool[off]) == (uint32_t)nw;
            for (int i = 0; same && i < nw; i++) same = WQ_VLOAD_U32(&t.pool[off + 1 + i]) == kw[i];
            if (same) return (int64_t)slot;
        }
    }
    return -1;
}

// Key words of the alignments a read kept.  SE:
//   one hit,  >= 3 nodes : [0<<28 | npath, gene, node...]                      (SpliceGraphQuant.long key)
//   several hits         : [1<<28 | nhits, then per hit: npath, gene, node...] (MultiMapping.map key)
#define WQ_KEY_WORDS_MAX (1 + WQ_MAX_TOLERATE * (2 + 2 * WQ_MAX_PATH))
WQ_HD int wq_build_key_se(const WqWork& wk, uint32_t base, int cnt, int nkept, uint32_t* kw) {
    int n = 0;
    if (nkept > 1) kw[n++] = (1u << 28) | (uint32_t)nkept;
    for (int k = 0; k < cnt; k++) {
        const WqHit h = wk.hits[base + k];
        if (!(h.flags & 4)) continue;
        kw[n++] = nkept > 1 ? (uint32_t)h.npath : ((0u << 28) | (uint32_t)h.npath);
        kw[n++] = wk.pool[h.path_start].gene;
        for (int i = 0; i < h.npath; i++) kw[n++] = wk.pool[h.path_start + i].node;
    }
    return n;
}

struct WqQuantDev {
    uint64_t*    node_count;       // [N] SpliceGraphQuant.node, by global node index
    WqCounters*  counters;
    WqEdgeTable  edges;
    WqKeyedTable keyed;
};

// Which counter a read's kept alignments go to (the branch structure of count!, src/quant.jl:556-594, and of the `length(align) > 1`
// test of process_reads!, src/reads.jl:60-66): kind -1 = counted nowhere, 0 = node store (node_idx), 1 = edge / circ table (ekey),
// 2 = keyed store (long path or multi-map class; kw[0..nw) holds the key).
struct WqTarget { int kind; uint32_t node_idx; uint64_t ekey; int nw; };
WQ_HD WqTarget wq_target_se(const WqDevIndex& ix, const WqWork& wk, uint32_t base, int cnt, int nkept, uint32_t* kw) {
    WqTarget t{-1, 0u, 0ull, 0};
    if (nkept == 1) {
        int k = 0;
        while (!(wk.hits[base + k].flags & 4)) k++;
        const WqHit h = wk.hits[base + k];
        if (!(h.flags & 1)) return t;               // align.isvalid == true || return
        const wq_align_node n0 = wk.pool[h.path_start];
        if (h.npath == 1) { t.kind = 0; t.node_idx = WQ_LDG(&ix.genes[n0.gene - 1].node_base) + n0.node - 1; return t; }
        if (h.npath == 2) {
            const wq_align_node n1 = wk.pool[h.path_start + 1];
            t.kind = 1;
            t.ekey = ((uint64_t)n0.gene << 32) | ((uint64_t)(n0.node & 0xFFFF) << 16) | (uint64_t)(n1.node & 0xFFFF);
            return t;
        }
    }
    t.nw = wq_build_key_se(wk, base, cnt, nkept, kw);
    t.kind = 2;
    return t;
}

// count!(quant, align, 1.0) for a single kept alignment, or push!(multi, ...) for several.
// `retry` = only redo the keyed insertion (dense counts were already applied).
WQ_HD void wq_count_read(const WqDevIndex& ix, const WqWork& wk, const WqQuantDev& q, uint32_t r,
                         uint32_t base, int cnt, int nkept, bool retry) {
    uint32_t kw[WQ_KEY_WORDS_MAX];
    const WqTarget t = wq_target_se(ix, wk, base, cnt, nkept, kw);
    if (t.kind < 0) return;
    if (t.kind == 0) { if (!retry) WQ_ATOMIC_ADD_U64(&q.node_count[t.node_idx], 1); return; }
    if (t.kind == 1) { if (!retry && wq_edge_insert(q.edges, t.ekey, 1)) WQ_ATOMIC_ADD_U64(&q.counters->edge_full, 1); return; }
    int rc = wq_keyed_insert(q.keyed, kw, t.nw, 1);
    if (rc == 1) {
        uint32_t slot = WQ_ATOMIC_ADD_U32(&wk.cursors[2], 1);
        wk.pending[slot] = r;
    } else if (rc == 2) {
        WQ_ATOMIC_ADD_U64(&q.counters->keyed_full, 1);
    }
}

// ---- kernel 3 -----------------------------------------------------------------------------------
WQ_HDN void wq_resolve_thread(const WqDevIndex& ix, const WqParams& p, const WqBatchDev& b, const WqWork& wk,
                              const WqQuantDev& q, uint32_t r, uint64_t* stats /* [mapped, multi, sum_readlen, overflow, pool_overflow] */) {
    const WqSeed s = wk.seeds[r];
    const int cnt = s.cnt;
    if (cnt == 0) return;
    const uint32_t base = s.hit_base;
    const int readlen = WQ_LDG(b.len + r);
    (void)readlen;
    // score filter of ungapped_align (src/align.jl:239-268), hits visited in suffix-array row order
    bool isnull = true;
    float maxscore = 0.f;
    uint32_t keep = 0;                  // bit k = hit k is in res.value
    bool overflow = false, pool_over = false;
    for (int k = 0; k < cnt; k++) {
        const WqHit h = wk.hits[base + k];
        if (h.flags & 8) overflow = true;
        if (h.flags & 16) pool_over = true;
        if (!(h.flags & 2)) continue;
        const float scvar = h.identity;
        if (isnull) {
            isnull = false;
            keep = 1u << k;
            maxscore = scvar;
        } else if (scvar > maxscore) {
            if ((double)(scvar - maxscore) > p.score_range) {
                keep = 0;
            } else {                    // splice_by_identity!(res, scvar, score_range, readlen)
                for (int j = 0; j < k; j++)
                    if ((keep >> j) & 1)
                        if ((double)(scvar - wk.hits[base + j].identity) > p.score_range) keep &= ~(1u << j);
            }
            maxscore = scvar;
            keep |= 1u << k;
        } else if ((double)(maxscore - scvar) <= p.score_range) {
            keep |= 1u << k;
        }
    }
    if (overflow) stats[3] += 1;
    if (pool_over) stats[4] += 1;
    if (isnull || overflow || pool_over) return;
    int nkept = 0;
    for (int k = 0; k < cnt; k++)
        if ((keep >> k) & 1) {
            wk.hits[base + k].flags |= 4;
            nkept++;
        }
    stats[0] += 1;
    stats[2] += (uint64_t)WQ_LDG(b.len + r);
    if (nkept > 1) stats[1] += 1;
    wq_count_read(ix, wk, q, r, base, cnt, nkept, false);
}

WQ_HDN void wq_retry_thread(const WqDevIndex& ix, const WqWork& wk, const WqQuantDev& q, uint32_t r) {
    const WqSeed s = wk.seeds[r];
    int nkept = 0;
    for (int k = 0; k < s.cnt; k++) nkept += (wk.hits[s.hit_base + k].flags & 4) ? 1 : 0;
    wq_count_read(ix, wk, q, r, s.hit_base, s.cnt, nkept, true);
}


// =================================================================================================
// Paired end (src/paired.jl:42-125, src/quant.jl:597-653, src/reads.jl:83-129)
//   WqSeed reuse: sp = rev_readloc | (mate-2 flipped) << 8, readloc = fwd_readloc, strand = mate-1 strand,
//   cnt = candidate pairs found by the sorted-offset walk.
// =================================================================================================
WQ_HD void wq_sort_small(int64_t* a, int n) {
    for (int i = 1; i < n; i++) {
        int64_t v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; }
        a[j + 1] = v;
    }
}

// kernel P1: one read pair -> both seeds, sorted offsets, two-pointer walk (src/paired.jl:44-51, 64-77, 112-121)
WQ_HDN void wq_seed_pe_thread(const WqDevIndex& ix, const WqParams& p, const WqBatchDev& bf, const WqBatchDev& br,
                              const WqWork& wk, uint32_t r, uint64_t* wbuf) {
    WqCtx c;
    c.ix = &ix;
    c.p = &p;
    c.w = wbuf;
    uint64_t badq[4];
    // mate 1: both strands unless stranded
    wq_load_read(c, bf, r);
    wq_badq(c, p.seed_min_qual, false, badq);
    int64_t fsp = 2, rsp = 2;
    int floc = 0, rloc = 0;
    bool strand = true, dummy = true;
    int fcnt = wq_seed_locate(c, badq, true, !p.is_stranded, fsp, floc, strand);
    // mate 2: one strand only; reverse-complemented first when the pair is expected head to head
    wq_load_read(c, br, r);
    bool flip2 = (p.is_pair_rc && strand) || (!p.is_pair_rc && !strand);
    int rcnt;
    if (flip2) {
        wq_revcomp(c.w, c.readlen);
        c.flipped = 1;
        wq_badq(c, p.seed_min_qual, true, badq);
        rcnt = wq_seed_locate(c, badq, false, false, rsp, rloc, dummy);
    } else {
        wq_badq(c, p.seed_min_qual, false, badq);
        rcnt = wq_seed_locate(c, badq, true, false, rsp, rloc, dummy);
    }
    int64_t fpos[WQ_MAX_TOLERATE], rpos[WQ_MAX_TOLERATE];
    for (int k = 0; k < fcnt; k++) fpos[k] = wq_sa_value(ix, fsp + k) + 1;
    for (int k = 0; k < rcnt; k++) rpos[k] = wq_sa_value(ix, rsp + k) + 1;
    wq_sort_small(fpos, fcnt);
    wq_sort_small(rpos, rcnt);
    uint32_t pf[WQ_MAX_TOLERATE], pr[WQ_MAX_TOLERATE], pg[WQ_MAX_TOLERATE];
    int np = 0;
    int fi = 0, ri = 0;
    while (fi < fcnt && ri < rcnt) {
        int64_t dist = fpos[fi] - rpos[ri];
        if (dist < 0) dist = -dist;
        if (dist < p.seed_pair_range) {
            // geneind = searchsortedlast(lib.offset, fwd_sorted[fidx]) and its text span
            uint64_t s = (uint64_t)fpos[fi] > ix.n ? ix.n : (uint64_t)fpos[fi];
            uint32_t gi = wq_locate_node(ix, s);
            uint32_t gene = WQ_LDG(&ix.nodes[gi].gene);
            const WqGeneRec* g = ix.genes + (gene - 1);
            int64_t lo = WQ_LDG(&g->text_start), hi = WQ_LDG(&g->text_end);
            if (lo <= rpos[ri] && rpos[ri] < hi) {
                pf[np] = (uint32_t)fpos[fi]; pr[np] = (uint32_t)rpos[ri]; pg[np] = gene; np++;
                ri++; fi++;
                continue;
            }
        }
        if (fpos[fi] > rpos[ri]) ri++; else fi++;
    }
    WqSeed sd;
    sd.sp = (uint32_t)rloc | (flip2 ? 0x100u : 0u);
    sd.cnt = (uint8_t)np;
    sd.readloc = (uint8_t)floc;
    sd.strand = strand ? 1 : 0;
    sd._pad = 0;
    sd.hit_base = 0;
    const uint32_t base = WQ_ALLOC(&wk.cursors[0], np);
    if (np > 0) {
        sd.hit_base = base;
        for (int k = 0; k < np; k++)
            if (base + k < wk.hit_cap) {
                wk.hit_read[base + k] = r;
                wk.hit_s[base + k] = pf[k];
                wk.hit_s2[base + k] = pr[k];
                wk.hit_gene[base + k] = pg[k];
            }
    }
    wk.seeds[r] = sd;
}

// kernel P2: one (pair slot, mate) -> _ungapped_align with the gene pinned (src/paired.jl:79-80); task = 2*slot + mate
WQ_HDN void wq_extend_pe_thread(const WqDevIndex& ix, const WqParams& p, const WqBatchDev& bf, const WqBatchDev& br,
                                const WqWork& wk, uint32_t slot, int mate, uint64_t* wbuf, const WqPathStore& st, bool first_pass) {
    uint32_t r = wk.hit_read[slot];
    WqSeed s = wk.seeds[r];
    WqCtx c;
    c.ix = &ix;
    c.p = &p;
    c.w = wbuf;
    wq_bind_store(c, st);
    bool ispos = s.strand != 0;
    int readloc;
    int64_t pos;
    if (mate == 0) {
        wq_load_read(c, bf, r);
        if (!ispos) { wq_revcomp(c.w, c.readlen); c.flipped = 1; }      // reverse_complement!(fwd), src/paired.jl:55-58
        readloc = s.readloc;
        pos = wk.hit_s[slot];
    } else {
        wq_load_read(c, br, r);
        if (s.sp & 0x100u) { wq_revcomp(c.w, c.readlen); c.flipped = 1; }
        readloc = (int)(s.sp & 0xFF);
        pos = wk.hit_s2[slot];
        ispos = !ispos;
    }
    wq_align_at(c, pos, readloc, ispos, wk.hit_gene[slot]);
    if (c.overflow && first_pass) {
        wk.defer[WQ_ATOMIC_ADD_U32(&wk.cursors[6], 1)] = 2 * slot + (uint32_t)mate;
        return;
    }
    wq_emit_hit(c, wk, 2 * slot + (uint32_t)mate, wq_score(c));     // paired scoring combines the two raw scores (src/paired.jl:2-3)
}

// Base.in(first::SGAlignPath, last::SGAlignPath): contiguous sub-path (src/quant.jl:47-62)
WQ_HD bool wq_subpath(const wq_align_node* a, int na, const wq_align_node* b, int nb) {
    for (int i = 0; i + na <= nb; i++) {
        bool match = true;
        for (int j = 0; j < na; j++)
            if (a[j].gene != b[i + j].gene || a[j].node != b[i + j].node) { match = false; break; }
        if (match) return true;
    }
    return false;
}

// key words of the kept pairs:
//   one pair, not nested : [2<<28 | npf | npr<<8, gene, fwd nodes.., rev nodes..]               (long, SGAlignPaired)
//   several pairs        : [3<<28 | npairs, then per pair the record above]                      (MultiMapping)
#define WQ_KEY_WORDS_MAX_PE (1 + WQ_MAX_TOLERATE * (2 + 2 * WQ_MAX_PATH))
WQ_HD int wq_build_key_pe(const WqWork& wk, uint32_t base, int cnt, int nkept, uint32_t* kw) {
    int n = 0;
    if (nkept > 1) kw[n++] = (3u << 28) | (uint32_t)nkept;
    for (int k = 0; k < cnt; k++) {
        const WqHit hf = wk.hits[2 * (base + k)], hr = wk.hits[2 * (base + k) + 1];
        if (!(hf.flags & 4)) continue;
        kw[n++] = (2u << 28) | (uint32_t)hf.npath | ((uint32_t)hr.npath << 8);
        kw[n++] = wk.pool[hf.path_start].gene;
        for (int i = 0; i < hf.npath; i++) kw[n++] = wk.pool[hf.path_start + i].node;
        for (int i = 0; i < hr.npath; i++) kw[n++] = wk.pool[hr.path_start + i].node;
    }
    return n;
}

// the paired form of the same decision: count!(quant, fwd, rev, value), src/quant.jl:597-653
WQ_HD WqTarget wq_target_pe(const WqDevIndex& ix, const WqWork& wk, uint32_t base, int cnt, int nkept, uint32_t* kw) {
    WqTarget t{-1, 0u, 0ull, 0};
    if (nkept == 1) {
        int k = 0;
        while (!(wk.hits[2 * (base + k)].flags & 4)) k++;
        const WqHit hf = wk.hits[2 * (base + k)], hr = wk.hits[2 * (base + k) + 1];
        if (!((hf.flags & 1) && (hr.flags & 1))) return t;
        const wq_align_node* pf = wk.pool + hf.path_start;
        const wq_align_node* pr = wk.pool + hr.path_start;
        if (pf[0].gene != pr[0].gene) return t;
        const wq_align_node* single = nullptr;
        int ns = 0;
        if (hf.npath <= hr.npath) { if (wq_subpath(pf, hf.npath, pr, hr.npath)) { single = pr; ns = hr.npath; } }
        else { if (wq_subpath(pr, hr.npath, pf, hf.npath)) { single = pf; ns = hf.npath; } }
        if (single) {
            if (ns == 1) { t.kind = 0; t.node_idx = WQ_LDG(&ix.genes[single[0].gene - 1].node_base) + single[0].node - 1; }
            else if (ns == 2) {
                t.kind = 1;
                t.ekey = ((uint64_t)single[0].gene << 32) | ((uint64_t)(single[0].node & 0xFFFF) << 16) | (uint64_t)(single[1].node & 0xFFFF);
            }                               // a nested path of >= 3 nodes is not counted (no else branch at src/quant.jl:631-640)
            return t;
        }
    }
    t.nw = wq_build_key_pe(wk, base, cnt, nkept, kw);
    t.kind = 2;
    return t;
}

// count!(quant, fwd, rev, 1.0) (src/quant.jl:597-653) or push!(multi, fwd, rev, ...) (:471-484)
WQ_HD void wq_count_pair(const WqDevIndex& ix, const WqWork& wk, const WqQuantDev& q, uint32_t r,
                         uint32_t base, int cnt, int nkept, bool retry) {
    uint32_t kw[WQ_KEY_WORDS_MAX_PE];
    const WqTarget t = wq_target_pe(ix, wk, base, cnt, nkept, kw);
    if (t.kind < 0) return;
    if (t.kind == 0) { if (!retry) WQ_ATOMIC_ADD_U64(&q.node_count[t.node_idx], 1); return; }
    if (t.kind == 1) { if (!retry && wq_edge_insert(q.edges, t.ekey, 1)) WQ_ATOMIC_ADD_U64(&q.counters->edge_full, 1); return; }
    int rc = wq_keyed_insert(q.keyed, kw, t.nw, 1);
    if (rc == 1) {
        uint32_t slot = WQ_ATOMIC_ADD_U32(&wk.cursors[2], 1);
        wk.pending[slot] = r;
    } else if (rc == 2) {
        WQ_ATOMIC_ADD_U64(&q.counters->keyed_full, 1);
    }
}

// kernel P3: one read pair -> pair score filter (src/paired.jl:82-111) + counting
WQ_HDN void wq_resolve_pe_thread(const WqDevIndex& ix, const WqParams& p, const WqBatchDev& bf, const WqBatchDev& br,
                                 const WqWork& wk, const WqQuantDev& q, uint32_t r, uint64_t* stats) {
    const WqSeed s = wk.seeds[r];
    const int cnt = s.cnt;
    if (cnt == 0) return;
    const uint32_t base = s.hit_base;
    const float lensum = (float)((int)WQ_LDG(bf.len + r) + (int)WQ_LDG(br.len + r));
    bool isnull = true;
    float maxscore = 0.f;
    uint32_t keep = 0;
    bool overflow = false, pool_over = false;
    for (int k = 0; k < cnt; k++) {
        const WqHit hf = wk.hits[2 * (base + k)], hr = wk.hits[2 * (base + k) + 1];
        if ((hf.flags | hr.flags) & 8) overflow = true;
        if ((hf.flags | hr.flags) & 16) pool_over = true;
        if (!((hf.flags & 2) && (hr.flags & 2))) continue;
        const float scvar = (hf.identity + hr.identity) / lensum;
        if (isnull) {
            isnull = false;
            keep = 1u << k;
            maxscore = scvar;
        } else if (scvar > maxscore) {
            if ((double)(scvar - maxscore) > p.score_range) {
                keep = 0;
            } else {
                for (int j = 0; j < k; j++)
                    if ((keep >> j) & 1) {
                        float sj = (wk.hits[2 * (base + j)].identity + wk.hits[2 * (base + j) + 1].identity) / lensum;
                        if ((double)(scvar - sj) > p.score_range) keep &= ~(1u << j);
                    }
            }
            maxscore = scvar;
            keep |= 1u << k;
        } else if ((double)(maxscore - scvar) <= p.score_range) {
            keep |= 1u << k;
        }
    }
    if (overflow) stats[3] += 1;
    if (pool_over) stats[4] += 1;
    if (isnull || overflow || pool_over) return;
    int nkept = 0;
    for (int k = 0; k < cnt; k++)
        if ((keep >> k) & 1) {
            wk.hits[2 * (base + k)].flags |= 4;
            wk.hits[2 * (base + k) + 1].flags |= 4;
            nkept++;
        }
    stats[0] += 1;
    stats[2] += (uint64_t)WQ_LDG(bf.len + r);
    if (nkept > 1) stats[1] += 1;
    wq_count_pair(ix, wk, q, r, base, cnt, nkept, false);
}

WQ_HDN void wq_retry_pe_thread(const WqDevIndex& ix, const WqWork& wk, const WqQuantDev& q, uint32_t r) {
    const WqSeed s = wk.seeds[r];
    int nkept = 0;
    for (int k = 0; k < s.cnt; k++) nkept += (wk.hits[2 * (s.hit_base + k)].flags & 4) ? 1 : 0;
    wq_count_pair(ix, wk, q, r, s.hit_base, s.cnt, nkept, true);
}
