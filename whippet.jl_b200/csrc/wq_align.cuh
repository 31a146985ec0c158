// Device-side seeding and ungapped splice-graph extension.
//
// What it computes (bit-exact with the reference, which it cites, but organised for a GPU thread:
// 2-bit words compared 32 bases at a time with XOR/popc, occurrence blocks of one sector each,
// node records of one sector each, an explicit per-thread path array instead of heap vectors):
//   fm_search      FMIndexes.sa_range override        /root/reference/src/fmindex_patch.jl:19-31
//   seed_locate    seed scheduling                     src/align.jl:145-185
//   fwd_extend     ungapped_fwd_extend + closure       src/align.jl:275-418
//   rev_extend     ungapped_rev_extend + closure       src/align.jl:423-575
//   align_at       _ungapped_align                     src/align.jl:209-223
#pragma once
#include "wq_types.h"

#if defined(__CUDA_ARCH__)
#define WQ_LDG(p) __ldg(p)
#define WQ_POPCLL(x) __popcll(x)
#define WQ_CTZLL(x) (__ffsll((long long)(x)) - 1)
#define WQ_CLZLL(x) __clzll((long long)(x))
#define WQ_CLZ32(x) __clz((int)(x))
#else
#define WQ_LDG(p) (*(p))
#define WQ_POPCLL(x) __builtin_popcountll(x)
#define WQ_CTZLL(x) __builtin_ctzll(x)
#define WQ_CLZLL(x) __builtin_clzll(x)
#define WQ_CLZ32(x) __builtin_clz(x)
#endif

struct WqParams {                  // AlignParam in kernel-friendly types
    int32_t mismatches, K, seed_try, seed_tolerate, seed_min_qual, seed_length, seed_buffer, seed_inc, seed_pair_range;
    float   mmlimit;               // Float32(mismatches): `mistol <= p.mismatches`
    float   kf;                    // Float32(kmer_size)
    double  score_range, score_min;
    uint8_t is_stranded, is_pair_rc;
    float   phred_prob[128];       // phred_to_prob, src/align.jl:124 (host-computed table)
};

struct WqCtx {
    const WqDevIndex* ix;
    const WqParams*   p;
    uint64_t* w;                         // [WQ_MAX_READ_WORDS+1] read in alignment orientation, zero padded (shared memory in the kernels)
    const uint8_t* qual;                 // global, original orientation
    int32_t  readlen;
    uint8_t  flipped;
    uint8_t  overflow;                   // bit0 path longer than the capacity, bit1 nesting deeper than WQ_MAX_DEPTH, bit2 a candidate had to be
                                         // parked and there is no side buffer (first-pass kernel), bit3 undefined in the reference
    uint8_t  depth;
    // gene being aligned against
    uint32_t gene, gstart, seqlen, gbase, nn;
    // The ONE working alignment of this thread (SGAlignment, src/align.jl:76-82).  Spliced extensions do not deep-copy it per
    // candidate as the reference does: they extend it in place and roll back to a checkpoint (node count + the few header
    // fields), which gives the same result because an extension only ever appends nodes beyond the checkpoint or drops
    // nodes from the end it grows at (never rewrites the kept ones; the one exception, the front node of a reverse
    // extension, is part of the checkpoint).  path[(lo..hi-1) * pstride]: the forward phase grows hi, the reverse phase
    // (after the path was moved to the end of the buffer) grows downwards at lo.
    WqPNode* path;
    int32_t  pstride, pcap;
    int32_t  lo, hi;
    uint32_t offset;
    uint8_t  offsetread, isvalid, strand;
    WqPNode* side;                       // [(WQ_MAX_DEPTH * pcap) * sstride] parked best candidates, or NULL
    int32_t  sstride;
};

// ---------------------------------------------------------------- read access
WQ_HD uint32_t wq_rbase(const WqCtx& c, int j) { return (uint32_t)(c.w[j >> 5] >> (2 * (j & 31))) & 3u; }
WQ_HD float wq_rprob(const WqCtx& c, int j) {   // j 0-based in alignment orientation
    uint8_t q = WQ_LDG(c.qual + (c.flipped ? c.readlen - 1 - j : j));
    return c.p->phred_prob[q & 127];
}
// 32 bases of the read starting at 0-based j (bases past the end are zero)
WQ_HD uint64_t wq_rext(const WqCtx& c, int j) {
    int wi = j >> 5, s = 2 * (j & 31);
    uint64_t v = c.w[wi] >> s;
    if (s) v |= c.w[wi + 1] << (64 - s);
    return v;
}
// kmer_index(read.sequence[a:a+K-1]) - 1, first base most significant (src/sgkmer.jl:12-26); a 1-based
WQ_HD uint32_t wq_read_kmer(const WqCtx& c, int a, int K) {
    uint64_t v = wq_rext(c, a - 1);          // base a at bits 0..1, a+1 at bits 2..3, ...
    uint32_t x = 0;
    for (int i = 0; i < K; i++) { x = (x << 2) | (uint32_t)(v & 3); v >>= 2; }
    return x;
}
// reverse complement of `len` bases held in w[] (reverse_complement!, src/record.jl:23-26)
WQ_HD uint64_t wq_rev2(uint64_t x) {      // reverse the order of the 32 two-bit groups of a word
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    return (x >> 32) | (x << 32);
}
WQ_HD void wq_revcomp(uint64_t* w, int len) {
    // complement = bitwise NOT of the 2-bit codes; reverse = reverse words and the groups inside them,
    // then shift the whole string down by the unused tail of the last word
    const int nw = (len + 31) >> 5;
    const int sh = 2 * (32 * nw - len);                 // bits of padding that end up at the bottom
    uint64_t r[WQ_MAX_READ_WORDS + 1];
#pragma unroll
    for (int i = 0; i <= WQ_MAX_READ_WORDS; i++) r[i] = i < nw ? wq_rev2(~w[nw - 1 - i]) : 0ull;
#pragma unroll
    for (int i = 0; i <= WQ_MAX_READ_WORDS; i++) {
        uint64_t lo = i < nw ? r[i] : 0ull, hi = (i + 1 < nw) ? r[i + 1] : 0ull;
        uint64_t v = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
        w[i] = i < nw ? v : 0ull;
    }
    if (len & 31) w[nw - 1] &= (~0ull >> (64 - 2 * (len & 31)));
}

// ---------------------------------------------------------------- text access
// mismatch mask (bit 2k set = base k differs) between 32 read bases and the text at 0-based pos
WQ_HD uint64_t wq_text_mism(const WqDevIndex& ix, uint64_t pos, uint64_t rword) {
    uint64_t wi = pos >> 5;
    int s = 2 * (int)(pos & 31);
    uint64_t t = WQ_LDG(ix.text2 + wi) >> s;
    if (s) t |= WQ_LDG(ix.text2 + wi + 1) << (64 - s);
    uint64_t x = t ^ rword;
    uint64_t m = (x | (x >> 1)) & 0x5555555555555555ull;
    if (ix.amb) {   // an ambiguous reference base never equals a read base (DNA `==`, src/align.jl:367)
        int s1 = (int)(pos & 31);
        uint64_t a = (uint64_t)WQ_LDG(ix.amb + wi) >> s1;
        if (s1) a |= (uint64_t)WQ_LDG(ix.amb + wi + 1) << (32 - s1);
        a &= 0xffffffffull;
        // spread 32 bits to even bit positions
        a = (a | (a << 16)) & 0x0000ffff0000ffffull;
        a = (a | (a << 8)) & 0x00ff00ff00ff00ffull;
        a = (a | (a << 4)) & 0x0f0f0f0f0f0f0f0full;
        a = (a | (a << 2)) & 0x3333333333333333ull;
        a = (a | (a << 1)) & 0x5555555555555555ull;
        m |= a;
    }
    return m;
}

// ---------------------------------------------------------------- node access (1-based node ids)
WQ_HD const WqNodeRec* wq_node(const WqCtx& c, uint32_t nodeidx) { return c.ix->nodes + (c.gbase + nodeidx - 1); }
WQ_HD uint32_t wq_nodeoffset(const WqCtx& c, uint32_t nodeidx) { return WQ_LDG(&wq_node(c, nodeidx)->abs_start) - c.gstart + 1; }
WQ_HD uint32_t wq_nodelen(const WqCtx& c, uint32_t nodeidx) { return WQ_LDG(&wq_node(c, nodeidx)->len); }
// edgetype[j] for 1 <= j <= nn+1
WQ_HD uint8_t wq_edgetype(const WqCtx& c, uint32_t j) {
    return j <= c.nn ? WQ_LDG(&wq_node(c, j)->et_left) : WQ_LDG(&wq_node(c, c.nn)->et_right);
}

// ---------------------------------------------------------------- path helpers (src/align.jl:90-115)
// logical node i of the working alignment (0-based)
WQ_HD WqPNode& wq_pn(const WqCtx& c, int i) { return c.path[(c.lo + i) * c.pstride]; }
WQ_HD int wq_npath(const WqCtx& c) { return c.hi - c.lo; }
WQ_HD float wq_score(const WqCtx& c) {
    float s = 0.f;
    for (int i = c.lo; i < c.hi; i++) { const WqPNode n = c.path[i * c.pstride]; s += (float)n.m - n.tol; }
    return s;
}
WQ_HD float wq_mistol(const WqCtx& c) {
    float s = 0.f;
    for (int i = c.lo; i < c.hi; i++) s += c.path[i * c.pstride].tol;
    return s;
}
WQ_HD uint8_t wq_matches(const WqCtx& c, int from, int to) {     // logical [from, to)
    uint8_t m = 0;
    for (int i = from; i < to; i++) m = (uint8_t)(m + wq_pn(c, i).m);
    return m;
}
WQ_HD float wq_identity(const WqCtx& c) { return wq_score(c) / (float)c.readlen; }
WQ_HD bool wq_isvalid(const WqCtx& c) {
    return c.isvalid && wq_npath(c) >= 1 && (float)wq_matches(c, 0, wq_npath(c)) > wq_mistol(c);
}
WQ_HD void wq_push_back(WqCtx& c, uint32_t node, uint8_t m) {
    if (c.hi >= c.pcap) { c.overflow |= 1; return; }
    WqPNode n;
    n.node = (uint16_t)node; n.tol = 0.f; n.m = m; n.mm = 0;
    c.path[c.hi * c.pstride] = n;
    c.hi++;
}
WQ_HD void wq_push_front(WqCtx& c, uint32_t node, uint8_t m) {
    if (c.lo <= 0) { c.overflow |= 1; return; }
    c.lo--;
    WqPNode n;
    n.node = (uint16_t)node; n.tol = 0.f; n.m = m; n.mm = 0;
    c.path[c.lo * c.pstride] = n;
}
WQ_HD bool wq_bad_node(const WqPNode& n) { return n.mm >= n.m || (int)n.m - (int)n.mm <= 1; }

// ---------------------------------------------------------------- FM index
// occurrences of ch in bwt[0..i)
WQ_HD uint32_t wq_occ(const WqDevIndex& ix, uint32_t ch, uint64_t i) {
    const WqOccBlock* b = ix.occ + (i >> 6);
#if defined(__CUDA_ARCH__)
    const uint4 planes = __ldg(reinterpret_cast<const uint4*>(&b->lo));
    uint64_t lo = ((uint64_t)planes.y << 32) | planes.x, hi = ((uint64_t)planes.w << 32) | planes.z;
    uint32_t base = __ldg(&b->cnt[ch]);
#else
    uint64_t lo = b->lo, hi = b->hi;
    uint32_t base = b->cnt[ch];
#endif
    uint64_t m = ((ch & 1) ? lo : ~lo) & ((ch & 2) ? hi : ~hi);
    uint32_t r = (uint32_t)(i & 63);
    m &= (r ? (~0ull >> (64 - r)) : 0ull);
    return base + (uint32_t)WQ_POPCLL(m);
}

// One backward-search step (src/fmindex_patch.jl:23-29) on rows sp..ep of the (n+1)-row matrix.
WQ_HD void wq_fm_step(const WqDevIndex& ix, uint32_t ch, uint32_t& sp, uint32_t& ep) {
    const uint32_t sent = (uint32_t)ix.sentinel;
    const uint32_t a = (sp > sent ? sp - 1 : sp) - 1;
    const uint32_t b = (ep > sent ? ep - 1 : ep);
    const uint32_t oa = wq_occ(ix, ch, a);
    const uint32_t ob = wq_occ(ix, ch, b);
    sp = ix.C[ch] + oa + 1;
    ep = ix.C[ch] + ob;
}
// Rows of the k-mer whose symbols (text order) are the 2-bit groups of idx, lowest group first: what k backward steps starting
// from 1:(n+1) give.  The same rule fills the table (wq_k_ftab) and serves as the reference for it.
WQ_HD uint2 wq_ftab_entry(const WqDevIndex& ix, uint32_t idx, int k) {
    uint32_t sp = 1, ep = (uint32_t)ix.n + 1;
    for (int i = k - 1; i >= 0 && sp <= ep; i--) wq_fm_step(ix, (idx >> (2 * i)) & 3u, sp, ep);
    uint2 r;
    if (sp <= ep) { r.x = sp; r.y = ep; } else { r.x = 1; r.y = 0; }
    return r;
}

WQ_HD int64_t wq_sa_value(const WqDevIndex& ix, int64_t row) {
    return row == 1 ? (int64_t)ix.n : (int64_t)WQ_LDG(ix.sa + (row - 2));
}

// seed_locate (src/align.jl:145-185) with sa_range (src/fmindex_patch.jl:19-31) folded in.  `badq` has bit j set when
// quality[j] <= seed_min_qual (0-based, orientation of c.w).  Returns cnt (0 when no seed) and fills sp/readloc/strand.
//
// One loop, ONE place where the occurrence table is read: the lanes of a warp are in different tries, on different strands and
// at different depths of their searches, and they all meet at that step instead of serialising over separate copies of the
// search loop.  A search starts from the k-mer table when the seed is long enough (the first ftab_k of its steps in one
// lookup: the rows after i symbols depend on those symbols only), and ends as sa_range does: all symbols consumed, or sp > ep.
WQ_HD int wq_seed_locate(const WqCtx& c, const uint64_t* badq, bool offset_left, bool both_strands,
                         int64_t& out_sp, int& out_readloc, bool& out_strand) {
    const WqParams& p = *c.p;
    const WqDevIndex& ix = *c.ix;
    const int readlen = c.readlen;
    const int L = p.seed_length;
    int ctry = 1;
    int increment, increment_sm, curpos;
    if (offset_left) { increment = p.seed_inc; increment_sm = 1; curpos = p.seed_buffer; }
    else { increment = -p.seed_inc; increment_sm = -1; curpos = readlen - L - p.seed_buffer; }
    const int maxright = readlen - L - p.seed_buffer;
    const int kt = (ix.ftab != nullptr && L >= ix.ftab_k) ? ix.ftab_k : 0;
    const uint64_t lmask = L >= 32 ? ~0ull : ((1ull << (2 * L)) - 1);
    out_strand = true;
    bool searching = false, rc = false;
    int i = 0;
    uint32_t sp = 1, ep = 0;
    uint64_t seedw = 0;                 // the L bases of the window, base j at bits 2j
    for (;;) {
        if (!searching) {
            if (!(ctry <= p.seed_try && p.seed_buffer <= curpos && curpos <= maxright)) break;
            // minimum(quality[curpos : curpos+L-1]) <= seed_min_qual  <=>  any bad bit in the window: skipped without spending a try
            const int a = curpos - 1;
            bool bad = false;
            for (int j = a; j < a + L;) {
                int wi = j >> 6, bo = j & 63;
                int take = 64 - bo;
                if (take > a + L - j) take = a + L - j;
                uint64_t mask = (take == 64 ? ~0ull : ((1ull << take) - 1)) << bo;
                if (badq[wi] & mask) { bad = true; break; }
                j += take;
            }
            if (bad) { curpos += increment_sm; continue; }
            seedw = wq_rext(c, a) & lmask;
            ctry += 1;
            searching = true;
            rc = false;
        }
        if (i == 0 && sp > ep) {        // start the search of this strand (state after the previous one: i = 0, sp > ep)
            if (kt) {
                // forward: the last kt bases of the window, in text order; reverse complement: its last kt symbols are the complements
                // of the window's first kt bases in reverse order
                uint32_t idx;
                if (!rc) idx = (uint32_t)(seedw >> (2 * (L - kt)));
                else idx = (uint32_t)(wq_rev2(~seedw) >> (2 * (32 - kt)));
                idx &= (1u << (2 * kt)) - 1u;
                const uint2 e = WQ_LDG(ix.ftab + idx);
                sp = e.x; ep = e.y; i = kt;
            } else { sp = 1; ep = (uint32_t)ix.n + 1; }
        }
        if (i < L && sp <= ep) {
            // forward query is consumed last symbol first; the reverse complement consumes the window first symbol first, complemented
            const uint32_t ch = rc ? 3u - (uint32_t)((seedw >> (2 * i)) & 3) : (uint32_t)((seedw >> (2 * (L - 1 - i))) & 3);
            wq_fm_step(ix, ch, sp, ep);
            i++;
        }
        if (!(i < L && sp <= ep)) {     // this search is over
            const int64_t cnt = ep >= sp ? (int64_t)ep - (int64_t)sp + 1 : 0;
            const bool ok = rc ? (0 < cnt && cnt <= p.seed_tolerate) : !(cnt == 0 || cnt > p.seed_tolerate);
            if (ok) {
                out_sp = sp;
                out_readloc = rc ? readlen - (curpos + L - 1) + 1 : curpos;
                out_strand = !rc;
                return (int)cnt;
            }
            i = 0; sp = 1; ep = 0;
            if (!rc && both_strands) rc = true;
            else { searching = false; curpos += increment; }
        }
    }
    out_sp = 2;
    out_readloc = curpos;
    return 0;
}

// ---------------------------------------------------------------- base-by-base walk, 32 at a time
// Forward: compare read[ridx..] with gene position sgidx.. (both 1-based) for at most `lim` bases
// with the per-base rules of src/align.jl:367-378; stops after the base that lifts mistol above the
// limit.  Returns bases consumed.
WQ_HD int wq_walk_fwd(const WqCtx& c, WqPNode& ndref, float& mistol, int ridx, int64_t sgidx, int lim) {
    int done = 0;
    WqPNode nd = ndref;
    uint64_t tpos = (uint64_t)c.gstart + (uint64_t)sgidx - 1;
    while (done < lim) {
        int chunk = lim - done < 32 ? lim - done : 32;
        uint64_t m = wq_text_mism(*c.ix, tpos + done, wq_rext(c, ridx - 1 + done));
        if (chunk < 32) m &= (1ull << (2 * chunk)) - 1;
        int prev = 0;
        while (m) {
            int b = WQ_CTZLL(m) >> 1;
            m &= m - 1;
            nd.m = (uint8_t)(nd.m + (b - prev));
            float prob = wq_rprob(c, ridx - 1 + done + b);
            nd.tol += prob;
            mistol += prob;
            nd.mm = (uint8_t)(nd.mm + 1);
            prev = b + 1;
            if (!(mistol <= c.p->mmlimit)) { ndref = nd; return done + prev; }
        }
        nd.m = (uint8_t)(nd.m + (chunk - prev));
        done += chunk;
    }
    ndref = nd;
    return done;
}
// Reverse: compare read[ridx], read[ridx-1].. with gene positions sgidx, sgidx-1.. for at most `lim` bases.
WQ_HD int wq_walk_rev(const WqCtx& c, WqPNode& ndref, float& mistol, int ridx, int64_t sgidx, int lim) {
    int done = 0;
    WqPNode nd = ndref;
    while (done < lim) {
        int chunk = lim - done < 32 ? lim - done : 32;
        int r0 = ridx - done - chunk;                              // 0-based first read base of the chunk
        uint64_t t0 = (uint64_t)c.gstart + (uint64_t)(sgidx - done - chunk);   // 0-based text pos of the chunk start
        uint64_t m = wq_text_mism(*c.ix, t0, wq_rext(c, r0));
        if (chunk < 32) m &= (1ull << (2 * chunk)) - 1;
        int prev = chunk;                                          // bases [prev, chunk) already accounted
        while (m) {
            int b = (63 - WQ_CLZLL(m)) >> 1;
            m &= ~(1ull << (2 * b));
            nd.m = (uint8_t)(nd.m + (prev - 1 - b));
            float prob = wq_rprob(c, r0 + b);
            nd.tol += prob;
            mistol += prob;
            nd.mm = (uint8_t)(nd.mm + 1);
            prev = b;
            if (!(mistol <= c.p->mmlimit)) { ndref = nd; return done + (chunk - prev); }
        }
        nd.m = (uint8_t)(nd.m + prev);
        done += chunk;
    }
    ndref = nd;
    return done;
}

// ---------------------------------------------------------------- extension
// The nested (spliced) calls go through the out-of-line instantiation <false>; the outermost call of a hit is
// <true> and is inlined into the kernel.  The outermost call hands a COPY of the context to the out-of-line
// functions, so its own context never has its address taken and stays in registers; the copy lives in the
// thread's stack frame and is only touched when a spliced extension actually happens.
#if defined(__CUDACC__)
#define WQ_NOINLINE __noinline__
#else
#define WQ_NOINLINE
#endif
template <bool TOP> WQ_HDN void wq_fwd_extend(WqCtx& c, int64_t sgidx, int ridx, uint32_t nodeidx);
template <bool TOP> WQ_HDN void wq_rev_extend(WqCtx& c, int64_t sgidx, int ridx, uint32_t nodeidx);
WQ_HDN WQ_NOINLINE void wq_fwd_extend_nested(WqCtx& c, int64_t sgidx, int ridx, uint32_t nodeidx);
WQ_HDN WQ_NOINLINE void wq_rev_extend_nested(WqCtx& c, int64_t sgidx, int ridx, uint32_t nodeidx);

#if defined(WQ_EMU_STATS)
static uint64_t g_wq_emu_stats[8];      // tests/emu only: [0] parks, [1] unparks, [2] rollbacks, [3] nested extensions, [4] rollbacks below the checkpoint
#define WQ_STAT(i) (g_wq_emu_stats[i]++)
#else
#define WQ_STAT(i) ((void)0)
#endif
struct WqCheckpoint {                  // what a rollback restores
    int32_t  lo, hi;
    uint32_t offset;
    uint8_t  offsetread, isvalid;
    WqPNode  front;                    // reverse extensions may add to the front node they start from
};
WQ_HD WqCheckpoint wq_checkpoint(const WqCtx& c, bool rev) {
    WqCheckpoint k;
    k.lo = c.lo; k.hi = c.hi; k.offset = c.offset; k.offsetread = c.offsetread; k.isvalid = c.isvalid;
    if (rev && c.hi > c.lo) k.front = c.path[c.lo * c.pstride];
    else { k.front.node = 0; k.front.m = 0; k.front.mm = 0; k.front.tol = 0.f; }
    return k;
}
WQ_HD void wq_rollback(WqCtx& c, const WqCheckpoint& k, bool rev) {
    WQ_STAT(2);
    if (rev ? c.lo > k.lo : c.hi < k.hi) WQ_STAT(4);
    c.lo = k.lo; c.hi = k.hi; c.offset = k.offset; c.offsetread = k.offsetread; c.isvalid = k.isvalid;
    if (rev && k.hi > k.lo) c.path[k.lo * c.pstride] = k.front;
}
// Park the alignment now in place (the best candidate so far) in the side buffer of this nesting level: the nodes beyond
// the checkpoint (forward: [k.hi, c.hi); reverse: [c.lo, k.lo] including the front node) and the header.
struct WqParked { int32_t lo, hi; uint32_t offset; uint8_t offsetread, isvalid; };
WQ_HD bool wq_park(WqCtx& c, const WqCheckpoint& k, bool rev, WqParked& b) {
    if (!c.side) { c.overflow |= 4; return false; }
    WQ_STAT(0);
    WqPNode* sd = c.side + (size_t)c.depth * c.pcap * c.sstride;
    b.lo = c.lo; b.hi = c.hi; b.offset = c.offset; b.offsetread = c.offsetread; b.isvalid = c.isvalid;
    if (!rev) { for (int i = k.hi; i < c.hi; i++) sd[(i - k.hi) * c.sstride] = c.path[i * c.pstride]; }
    else { for (int i = c.lo; i <= k.lo && i < c.hi; i++) sd[(i - c.lo) * c.sstride] = c.path[i * c.pstride]; }
    return true;
}
WQ_HD void wq_unpark(WqCtx& c, const WqCheckpoint& k, bool rev, const WqParked& b) {
    WQ_STAT(1);
    const WqPNode* sd = c.side + (size_t)c.depth * c.pcap * c.sstride;
    c.lo = b.lo; c.hi = b.hi; c.offset = b.offset; c.offsetread = b.offsetread; c.isvalid = b.isvalid;
    if (!rev) { for (int i = k.hi; i < b.hi; i++) c.path[i * c.pstride] = sd[(i - k.hi) * c.sstride]; }
    else { for (int i = b.lo; i <= k.lo && i < b.hi; i++) c.path[i * c.pstride] = sd[(i - b.lo) * c.sstride]; }
}

// spliced_fwd_extend (src/align.jl:287-311): candidates = edges.left[kmer(edgeleft[edge])] ∩ edges.right[read kmer]
// under the gene-only ordering of src/edges.jl:15-18,40-55 (both cursors advance on a gene match,
// the right-table entry is returned), same gene as the alignment only.  [lb,le) / [rb,re) are the two non-empty lists.
WQ_HDN WQ_NOINLINE void wq_spliced_fwd(WqCtx& c, int ridx, uint32_t lb, uint32_t le, uint32_t rb, uint32_t re) {
    const WqDevIndex& ix = *c.ix;
    if (c.depth >= WQ_MAX_DEPTH) { c.overflow |= 2; return; }
    const WqCheckpoint ck = wq_checkpoint(c, false);
    float best_score = wq_score(c);
    const float base_score = best_score;
    bool in_place = false, parked = false;      // where the best candidate so far is (neither: the checkpoint is the best)
    WqParked pk = {0, 0, 0, 0, 0};
    uint32_t i = lb, j = rb;
    while (i < le && j < re && !c.overflow) {
        uint32_t ga = WQ_LDG(&ix.left_ent[i].x);
        uint2 eb = WQ_LDG(&ix.right_ent[j]);
        if (ga > eb.x) j++;
        else if (ga < eb.x) i++;
        else {
            i++; j++;
            if (eb.x != c.gene) continue;          // rn.gene == align.path[1].gene || continue
            uint32_t rn_node = eb.y - c.gbase + 1;
            int64_t rn_offset = (int64_t)wq_nodeoffset(c, rn_node);
            if (in_place) {                        // the next candidate starts from the checkpoint again (deepcopy(align))
                if (!wq_park(c, ck, false, pk)) break;
                parked = true; in_place = false;
                wq_rollback(c, ck, false);
            }
            c.depth++;
            WQ_STAT(3);
            wq_fwd_extend_nested(c, rn_offset, ridx, rn_node);
            c.depth--;
            float s = wq_score(c);
            if (s > best_score) { best_score = s; in_place = true; parked = false; }
            else wq_rollback(c, ck, false);
        }
    }
    if (c.overflow) return;
    if (!in_place && parked) wq_unpark(c, ck, false, pk);
    if (best_score >= base_score + c.p->kf) c.isvalid = 1;
}

// copies back what an out-of-line extension may have changed
WQ_HD void wq_ctx_merge(WqCtx& c, const WqCtx& cc) {
    c.overflow = cc.overflow; c.lo = cc.lo; c.hi = cc.hi; c.offset = cc.offset; c.offsetread = cc.offsetread; c.isvalid = cc.isvalid;
}
template <bool TOP> WQ_HD void wq_try_spliced_fwd(WqCtx& c, uint32_t edgeind, int ridx, uint32_t rkmer) {
    const WqDevIndex& ix = *c.ix;
    uint32_t rb = WQ_LDG(ix.right_ptr + rkmer), re = WQ_LDG(ix.right_ptr + rkmer + 1);
    if (rb == re) return;
    uint32_t lk = WQ_LDG(&wq_node(c, edgeind)->kleft);
    uint32_t lb = WQ_LDG(ix.left_ptr + lk), le = WQ_LDG(ix.left_ptr + lk + 1);
    if (lb == le) return;
    if (TOP) { WqCtx cc = c; wq_spliced_fwd(cc, ridx, lb, le, rb, re); wq_ctx_merge(c, cc); }
    else wq_spliced_fwd(c, ridx, lb, le, rb, re);
}

// ungapped_fwd_extend (src/align.jl:275-418).  sgidx/ridx are 1-based as in the reference.
template <bool TOP> WQ_HDN void wq_fwd_extend(WqCtx& c, int64_t sgidx, int ridx, uint32_t nodeidx) {
    const WqParams& p = *c.p;
    const int readlen = c.readlen;
    const int K = p.K;
    uint32_t passed = 0;                // bit v: an edge was passed when the path held v nodes (strictly increasing values)
    float mistol = wq_mistol(c);

    wq_push_back(c, nodeidx, 0);

    while (mistol <= p.mmlimit && ridx <= readlen && sgidx <= (int64_t)c.seqlen && !c.overflow) {
        int64_t node_end = 0;
        bool have_node = nodeidx <= c.nn;
        if (have_node) {
            const WqNodeRec* nr = wq_node(c, nodeidx);
            node_end = (int64_t)(WQ_LDG(&nr->abs_start) - c.gstart + 1) + (int64_t)WQ_LDG(&nr->len);
        }
        if (have_node && sgidx == node_end) {   // hit next edge
            uint32_t curedge = nodeidx + 1;
            uint8_t et = wq_edgetype(c, curedge);
            if (et == WQ_LR && (int64_t)wq_nodelen(c, curedge) >= K && readlen - ridx + 1 >= K) {
                uint32_t rkmer = wq_read_kmer(c, ridx, K);
                if (WQ_LDG(&wq_node(c, curedge)->kright) == rkmer) {
                    ridx += K - 1;
                    sgidx += K - 1;
                    nodeidx += 1;
                    wq_push_back(c, nodeidx, (uint8_t)(K - 1));
                    c.isvalid = 1;
                } else {
                    wq_try_spliced_fwd<TOP>(c, curedge, ridx, rkmer);
                    break;
                }
            } else if (et == WQ_LR || et == WQ_LL) {
                passed |= 1u << wq_npath(c);
                nodeidx += 1;
                wq_push_back(c, nodeidx, 0);
            } else if (et == WQ_LS) {
                if (readlen - ridx + 1 >= K) {
                    uint32_t rkmer = wq_read_kmer(c, ridx, K);
                    wq_try_spliced_fwd<TOP>(c, curedge, ridx, rkmer);
                }
                break;
            } else if (et == WQ_SR) {
                break;
            } else {   // 'RR' || 'SL' || 'RS'
                nodeidx += 1;
                if (nodeidx <= c.nn) wq_push_back(c, nodeidx, 0);
            }
            if (c.overflow) break;
            // the base at (ridx, sgidx) is compared in this same iteration; the next edge test happens
            // only after it (an `if`, not a `while`: src/align.jl:314)
            have_node = nodeidx <= c.nn;
            if (have_node) {
                const WqNodeRec* nr = wq_node(c, nodeidx);
                node_end = (int64_t)(WQ_LDG(&nr->abs_start) - c.gstart + 1) + (int64_t)WQ_LDG(&nr->len);
            }
        }
        int lim = readlen - ridx + 1;
        int64_t l2 = (int64_t)c.seqlen - sgidx + 1;
        if (l2 < lim) lim = (int)l2;
        if (have_node && sgidx < node_end && node_end - sgidx < lim) lim = (int)(node_end - sgidx);
        int used = wq_walk_fwd(c, c.path[(c.hi - 1) * c.pstride], mistol, ridx, sgidx, lim);
        ridx += used;
        sgidx += used;
    }

    // passed-edge backtracking (src/align.jl:382-405), most recent edge first
    while (passed && !c.overflow) {
        const int cval = 31 - WQ_CLZ32(passed);
        passed &= ~(1u << cval);
        if ((int)wq_matches(c, cval, wq_npath(c)) >= K + p.mismatches) {
            c.isvalid = 1;
            break;
        } else {
            if (cval > wq_npath(c)) { c.overflow |= 8; break; }      // undefined in the reference (out-of-bounds under @inbounds)
            uint32_t cur_edge = (uint32_t)wq_pn(c, cval - 1).node + 1;
            c.hi = c.lo + cval;
            int64_t cur_ridx = (int64_t)ridx - (sgidx - (int64_t)wq_nodeoffset(c, cur_edge));
            if (!((cur_ridx + K - 1) <= readlen)) continue;
            uint32_t rkmer = wq_read_kmer(c, (int)cur_ridx, K);
            // res = deepcopy(align); spliced_fwd_extend(res); kept only when it is longer AND scores higher
            const WqCheckpoint ck = wq_checkpoint(c, false);
            const float s0 = wq_score(c);
            wq_try_spliced_fwd<TOP>(c, cur_edge, (int)cur_ridx, rkmer);
            if (c.overflow) break;
            if (!(wq_npath(c) > ck.hi - ck.lo && wq_score(c) > s0)) wq_rollback(c, ck, false);
        }
    }

    // clean up bad trailing nodes (src/align.jl:408-412)
    while (wq_npath(c) >= 1 && wq_bad_node(c.path[(c.hi - 1) * c.pstride])) {
        c.hi--;
        if (wq_npath(c) == 0) c.isvalid = 0;
    }
    if ((double)wq_identity(c) >= p.score_min) c.isvalid = 1;
}

// spliced_rev_extend (src/align.jl:437-461): edges.right[kmer(edgeright[edge])] ∩ edges.left[read kmer],
// returning the left-table entries.
WQ_HDN WQ_NOINLINE void wq_spliced_rev(WqCtx& c, int ridx, uint32_t lb, uint32_t le, uint32_t rb, uint32_t re) {
    const WqDevIndex& ix = *c.ix;
    if (c.depth >= WQ_MAX_DEPTH) { c.overflow |= 2; return; }
    const WqCheckpoint ck = wq_checkpoint(c, true);
    float best_score = wq_score(c);
    const float base_score = best_score;
    bool in_place = false, parked = false;
    WqParked pk = {0, 0, 0, 0, 0};
    uint32_t i = rb, j = lb;     // A = right list, B = left list
    while (i < re && j < le && !c.overflow) {
        uint32_t ga = WQ_LDG(&ix.right_ent[i].x);
        uint2 eb = WQ_LDG(&ix.left_ent[j]);
        if (ga > eb.x) j++;
        else if (ga < eb.x) i++;
        else {
            i++; j++;
            if (eb.x != c.gene) continue;
            uint32_t rn_node = eb.y - c.gbase + 1;
            if (rn_node < 2) { c.overflow |= 8; continue; }        // nodeoffset[0] in the reference: undefined
            int64_t rn_offset = (int64_t)wq_nodeoffset(c, rn_node - 1) + (int64_t)wq_nodelen(c, rn_node - 1) - 1;
            if (in_place) {
                if (!wq_park(c, ck, true, pk)) break;
                parked = true; in_place = false;
                wq_rollback(c, ck, true);
            }
            c.depth++;
            WQ_STAT(3);
            wq_rev_extend_nested(c, rn_offset, ridx, rn_node - 1);
            c.depth--;
            float s = wq_score(c);
            if (s > best_score) { best_score = s; in_place = true; parked = false; }
            else wq_rollback(c, ck, true);
        }
    }
    if (c.overflow) return;
    if (!in_place && parked) { wq_rollback(c, ck, true); wq_unpark(c, ck, true, pk); }
    if (best_score >= base_score + c.p->kf) c.isvalid = 1;
}

template <bool TOP> WQ_HD void wq_try_spliced_rev(WqCtx& c, uint32_t edgeind, int ridx, uint32_t lkmer) {
    const WqDevIndex& ix = *c.ix;
    uint32_t lb = WQ_LDG(ix.left_ptr + lkmer), le = WQ_LDG(ix.left_ptr + lkmer + 1);
    if (lb == le) return;
    uint32_t rk = WQ_LDG(&wq_node(c, edgeind)->kright);
    uint32_t rb = WQ_LDG(ix.right_ptr + rk), re = WQ_LDG(ix.right_ptr + rk + 1);
    if (rb == re) return;
    if (TOP) { WqCtx cc = c; wq_spliced_rev(cc, ridx, lb, le, rb, re); wq_ctx_merge(c, cc); }
    else wq_spliced_rev(c, ridx, lb, le, rb, re);
}

// ungapped_rev_extend (src/align.jl:423-575)
template <bool TOP> WQ_HDN void wq_rev_extend(WqCtx& c, int64_t sgidx, int ridx, uint32_t nodeidx) {
    const WqParams& p = *c.p;
    const int K = p.K;
    uint32_t passed = 0;                // bit v: an edge was passed when the path held v + 1 nodes
    float mistol = wq_mistol(c);

    if (wq_npath(c) == 0 || (uint32_t)c.path[c.lo * c.pstride].node != nodeidx) wq_push_front(c, nodeidx, 0);

    while (mistol <= p.mmlimit && ridx > 0 && sgidx > 0 && !c.overflow) {
        int64_t node_start = nodeidx >= 1 ? (int64_t)wq_nodeoffset(c, nodeidx) : -1;
        if (nodeidx >= 1 && sgidx == node_start - 1) {   // hit right edge
            uint32_t leftnode = nodeidx - 1;
            uint8_t et = wq_edgetype(c, nodeidx);
            if (et == WQ_LR && (int64_t)wq_nodelen(c, leftnode) >= K && ridx >= K) {
                uint32_t lkmer = wq_read_kmer(c, ridx - K + 1, K);
                if (WQ_LDG(&wq_node(c, nodeidx)->kleft) == lkmer) {
                    ridx -= K - 1;
                    sgidx -= K - 1;
                    nodeidx -= 1;
                    wq_push_front(c, nodeidx, (uint8_t)(K - 1));
                    c.isvalid = 1;
                } else {
                    wq_try_spliced_rev<TOP>(c, nodeidx, ridx, lkmer);
                    break;
                }
            } else if (et == WQ_LR || et == WQ_RR) {
                passed |= 1u << (wq_npath(c) - 1);
                nodeidx -= 1;
                wq_push_front(c, nodeidx, 0);
            } else if (et == WQ_SR) {
                if (ridx >= K) {
                    uint32_t lkmer = wq_read_kmer(c, ridx - K + 1, K);
                    wq_try_spliced_rev<TOP>(c, nodeidx, ridx, lkmer);
                }
                break;
            } else if (et == WQ_LS) {
                break;
            } else {   // 'LL' || 'RS' || 'SL'
                nodeidx -= 1;
                if (nodeidx > 0) wq_push_front(c, nodeidx, 0);
            }
            if (c.overflow) break;
            node_start = nodeidx >= 1 ? (int64_t)wq_nodeoffset(c, nodeidx) : -1;
        }
        int lim = ridx;
        if (sgidx < lim) lim = (int)sgidx;
        // next edge test fires when sgidx == node_start-1, i.e. after sgidx-(node_start-1) more bases
        if (nodeidx >= 1 && sgidx > node_start - 1 && sgidx - (node_start - 1) < lim) lim = (int)(sgidx - (node_start - 1));
        int used = wq_walk_rev(c, c.path[c.lo * c.pstride], mistol, ridx, sgidx, lim);
        ridx -= used;
        sgidx -= used;
    }

    int64_t cur_ridx = ridx + 1;
    int64_t cur_sgidx = sgidx + 1;
    while (passed && !c.overflow) {
        const int cval = 31 - WQ_CLZ32(passed);
        passed &= ~(1u << cval);
        int len = wq_npath(c);
        int upto = len - (cval + 1);
        if ((int)wq_matches(c, 0, upto > 0 ? upto : 0) >= K + p.mismatches) {
            c.isvalid = 1;
            break;
        } else {
            if (len - cval - 1 < 0) { c.overflow |= 8; break; }      // undefined in the reference
            uint32_t cur_edge = wq_pn(c, len - cval - 1).node;
            if (upto > 0) c.lo += upto;                               // deleteat!(path, 1:upto)
            int64_t no = (int64_t)wq_nodeoffset(c, cur_edge);
            cur_ridx = (int64_t)ridx + (no - sgidx);
            cur_sgidx = no;
            if (!((cur_ridx - K) > 0)) continue;
            uint32_t lkmer = wq_read_kmer(c, (int)(cur_ridx - K), K);
            const WqCheckpoint ck = wq_checkpoint(c, true);
            const float s0 = wq_score(c);
            wq_try_spliced_rev<TOP>(c, cur_edge, (int)(cur_ridx - 1), lkmer);
            if (c.overflow) break;
            if (!(wq_npath(c) > ck.hi - ck.lo && wq_score(c) > s0)) wq_rollback(c, ck, true);
        }
    }

    if ((double)wq_identity(c) >= p.score_min) c.isvalid = 1;
    if (sgidx < (int64_t)c.offset) c.offset = (uint32_t)cur_sgidx;
    if (cur_ridx < (int64_t)c.offsetread) c.offsetread = (uint8_t)cur_ridx;

    // clean up bad leading nodes (src/align.jl:562-572)
    while (wq_npath(c) >= 1 && wq_bad_node(c.path[c.lo * c.pstride])) {
        const WqPNode f = c.path[c.lo * c.pstride];
        uint8_t offset_change = (uint8_t)(f.m + f.mm);
        c.lo++;
        if (wq_npath(c) >= 1) {
            c.offset = wq_nodeoffset(c, c.path[c.lo * c.pstride].node);
            c.offsetread = (uint8_t)(c.offsetread + offset_change);
        } else {
            c.isvalid = 0;
        }
    }
}

WQ_HDN WQ_NOINLINE void wq_fwd_extend_nested(WqCtx& c, int64_t sgidx, int ridx, uint32_t nodeidx) { wq_fwd_extend<false>(c, sgidx, ridx, nodeidx); }
WQ_HDN WQ_NOINLINE void wq_rev_extend_nested(WqCtx& c, int64_t sgidx, int ridx, uint32_t nodeidx) { wq_rev_extend<false>(c, sgidx, ridx, nodeidx); }

// Locate gene + node of 1-based text position s (searchsortedlast over GraphLib.offset and
// SpliceGraph.nodeoffset, src/align.jl:210-213) with one bucket lookup over the globally sorted
// node starts.  forced_gene > 0 pins the gene (paired call, src/paired.jl:79-80).
WQ_HD void wq_set_gene(WqCtx& c, uint32_t gene) {
    const WqGeneRec* g = c.ix->genes + (gene - 1);
    c.gene = gene;
    c.gstart = WQ_LDG(&g->text_start);
    c.seqlen = WQ_LDG(&g->text_end) - c.gstart;
    c.gbase = WQ_LDG(&g->node_base);
    c.nn = WQ_LDG(&g->nn);
}
WQ_HD uint32_t wq_locate_node(const WqDevIndex& ix, uint64_t s) {
    // global index of the last node whose abs_start <= s (s is used as a 0-based position: the
    // reference compares a 1-based position with 0-based gene offsets, src/align.jl:210-212)
    uint64_t b = s >> WQ_NODE_BUCKET_SHIFT;
    uint32_t k = WQ_LDG(ix.node_bucket + b);          // #nodes with abs_start <= b<<6  (>= 1)
    while (k < ix.N && (uint64_t)WQ_LDG(&ix.nodes[k].abs_start) <= s) k++;
    return k - 1;
}

// _ungapped_align (src/align.jl:209-223).  The result is the context's working alignment (path[lo..hi), offset, ...).
WQ_HD void wq_align_at(WqCtx& c, int64_t indx, int readloc, bool ispos, uint32_t forced_gene) {
    const WqDevIndex& ix = *c.ix;
    uint32_t offset_node;
    if (forced_gene == 0) {
        uint64_t s = (uint64_t)indx > ix.n ? ix.n : (uint64_t)indx;
        uint32_t gi = wq_locate_node(ix, s);
        wq_set_gene(c, WQ_LDG(&ix.nodes[gi].gene));
        offset_node = gi - c.gbase + 1;
    } else {
        wq_set_gene(c, forced_gene);
        // searchsortedlast(nodeoffset, sgidx+1) inside the pinned gene
        int64_t target = indx - (int64_t)c.gstart;       // sgidx
        uint32_t lo = 0, hi = c.nn + 1;
        while (lo + 1 < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if ((int64_t)wq_nodeoffset(c, mid) <= target + 1) lo = mid; else hi = mid;
        }
        offset_node = lo;
    }
    int64_t sgidx = indx - (int64_t)c.gstart;
    c.offset = (uint32_t)(sgidx + 1);
    c.offsetread = (uint8_t)(readloc + 1);
    c.lo = c.hi = 0;
    c.isvalid = 0;
    c.strand = ispos ? 1 : 0;
    c.depth = 0;
#ifdef WQ_TOP_NESTED
    wq_fwd_extend_nested(c, sgidx + 1, readloc + 1, offset_node);
#else
    wq_fwd_extend<true>(c, sgidx + 1, readloc + 1, offset_node);
#endif
    if (c.overflow) return;
    // the reverse extension grows the path at its front: move it to the end of the buffer once, so that a push at the
    // front is a store below `lo` instead of a shift of every node
    const int k = c.hi;
    for (int i = k - 1; i >= 0; i--) c.path[(c.pcap - k + i) * c.pstride] = c.path[i * c.pstride];
    c.lo = c.pcap - k; c.hi = c.pcap;
#ifdef WQ_TOP_NESTED
    wq_rev_extend_nested(c, sgidx, readloc, offset_node);
#else
    wq_rev_extend<true>(c, sgidx, readloc, offset_node);
#endif
}
