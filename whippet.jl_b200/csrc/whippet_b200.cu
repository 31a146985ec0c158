// libwhippet_b200.so — C ABI (include/whippet_b200.h) over hand-written sm_100a kernels.
//
// Kernels (bodies in wq_kernels.cuh / wq_align.cuh, EM in wq_em.cuh):
//   wq_k_seed      batched FM-index backward search + locate   (fmindex_patch.jl:19-31, align.jl:145-185)
//   wq_k_extend    ungapped fwd/rev extension across nodes     (align.jl:209-223, 275-575)
//   wq_k_resolve   multi-hit resolution + count!/push!(multi)  (align.jl:239-268, quant.jl:456-469, 556-594)
//   wq_k_retry     second chance for keyed inserts that met an unpublished slot
// There is no CPU implementation behind this ABI: without an sm_100 device wq_index_create fails.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "wq_kernels.cuh"
#include "wq_host_index.h"
#include "wq_host_quant.h"
#include "wq_em.cuh"
#include "wq_events.h"
#include "wq_fastq.h"
#include "wq_sam.h"
#include "wq_index_file.h"
#include "wq_bias.cuh"
#include <future>
#include <atomic>
#include <dlfcn.h>

#include <chrono>
struct WqTimer {
    const char* name; std::chrono::steady_clock::time_point t0; bool on;
    explicit WqTimer(const char* n) : name(n), t0(std::chrono::steady_clock::now()) { static int e = getenv("WQ_PROFILE") ? atoi(getenv("WQ_PROFILE")) : 0; on = e != 0; }
    ~WqTimer() { if (on) fprintf(stderr, "[wq] %-28s %9.3f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); }
};
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(-10, std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 16;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Events of one gene group.  Group 0 = genes that no multi-map class names: assign_ambig! cannot change their counts, so
// their events are built on the host threads while the transcript EM kernel runs and their PSI EM launch overlaps
// assign_ambig!.  Group 1 = the remaining genes, built after assign_ambig!.
struct WqEvPart { std::vector<WqEventRec> recs; WqPsiProblems P; WqEventPool pool; std::vector<double> ci_psi, ci_tot, p_total; };
struct WqEvBatch {
    std::vector<WqEvPart> parts;         // one per host thread, contiguous gene ranges in gene order
    WqPsiProblems P;                     // problems of all parts, concatenated
    std::vector<long> pshift;            // first problem of every part in P
    WqEmScratch dev;                     // device arrays of this group's PSI launch (the two launches overlap in time)
    double* psi = nullptr; int32_t* its = nullptr; size_t psi_cap = 0, its_cap = 0;     // pinned results
    // Beta posterior CI of the events whose psi is known without EM (most of them): launched with the batch
    std::vector<double> ci_psi, ci_tot;
    double* ci_out = nullptr; size_t ci_cap = 0;                                         // pinned: lo[n] then hi[n]
    // ... and of the events that need EM: a kernel behind their PSI launch on the same stream (total < 0: no CI)
    std::vector<double> p_total; double* ci_em = nullptr; size_t ci_em_cap = 0;          // pinned: lo[E] then hi[E]
    cudaStream_t stream = nullptr;       // own stream: the group's uploads, kernels and result copies are ordered against nothing else
    bool built = false, launched = false;
    int64_t readlen = -1;
    cudaError_t ensure_stream() { return stream ? cudaSuccess : cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking); }
    cudaError_t ensure_pinned(size_t npsi, size_t nits) {
        if (npsi > psi_cap) { if (psi) cudaFreeHost(psi); psi = nullptr; psi_cap = 0; cudaError_t e = cudaMallocHost(&psi, (npsi + npsi / 4 + 16) * 8); if (e != cudaSuccess) return e; psi_cap = npsi + npsi / 4 + 16; }
        if (nits > its_cap) { if (its) cudaFreeHost(its); its = nullptr; its_cap = 0; cudaError_t e = cudaMallocHost(&its, (nits + nits / 4 + 16) * 4); if (e != cudaSuccess) return e; its_cap = nits + nits / 4 + 16; }
        return cudaSuccess;
    }
    void release() { dev.release(); if (stream) cudaStreamDestroy(stream); stream = nullptr; if (ci_out) cudaFreeHost(ci_out); ci_out = nullptr; ci_cap = 0; if (ci_em) cudaFreeHost(ci_em); ci_em = nullptr; ci_em_cap = 0; if (psi) cudaFreeHost(psi); if (its) cudaFreeHost(its); psi = nullptr; its = nullptr; psi_cap = its_cap = 0; }
};

struct wq_handle {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, own_stream = nullptr;
    cudaStreamAttrValue l2attr; bool has_l2attr = false;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    WqDevIndex dix;
    WqHostIndex H;               // host mirror (kept for node/gene lookups while building classes)
    WqIsoIndex iso;              // annotated isoform paths (host)
    WqSamIndex sam;              // node coordinates + lib.info (SAM output only)
    // device index storage
    DevBuf<WqOccBlock> d_occ; DevBuf<uint32_t> d_sa; DevBuf<uint64_t> d_text2; DevBuf<uint32_t> d_amb;
    DevBuf<WqNodeRec> d_nodes; DevBuf<WqGeneRec> d_genes; DevBuf<uint32_t> d_bucket;
    DevBuf<uint32_t> d_lptr, d_rptr; DevBuf<uint2> d_lent, d_rent, d_ftab;
    // quant stores
    DevBuf<uint64_t> d_dense;    // [N] node counts followed by WqCounters
    DevBuf<uint64_t> d_ekey, d_ecount, d_khash, d_kcount; DevBuf<uint32_t> d_koff, d_kpool; DevBuf<uint64_t> d_kcursor;
    uint64_t edge_slots = 0, keyed_slots = 0, kpool_cap = 0;
    WqQuantDev q;
    // per-batch buffers
    DevBuf<uint8_t> d_len[2], d_qual[2], d_qualpk[2]; DevBuf<uint64_t> d_seqoff[2], d_qualoff[2], d_seq2[2];
    DevBuf<WqSeed> d_seeds; DevBuf<uint32_t> d_hit_read, d_hit_s, d_hit_s2, d_hit_gene, d_pending, d_pending2, d_cursors;
    DevBuf<WqHit> d_hits; DevBuf<wq_align_node> d_pool; DevBuf<uint32_t> d_defer; DevBuf<WqPNode> d_slab;
    int n_sm = 148; uint64_t pool_nodes_per_read = 8; uint64_t last_deferred = 0;
    DevBuf<uint32_t> d_mapped_bits;
    double rl_mean = 0.0; uint64_t rl_mapped = 0;      // running mean_readlen over the mapped reads in read order (src/reads.jl:69)
    uint32_t* h_cursors = nullptr;   // pinned
    // host-side quant state (filled by wq_counts_sync / merge; input of EM)
    WqHostQuant hq;
    WqHostQuantEx hx;
    WqEmScratch em_scratch;
    std::vector<WqEventRec> ev_recs; std::vector<double> ev_values; std::vector<uint64_t> ev_path_ptr{0}; std::vector<uint32_t> ev_path_nodes;
    bool events_done = false;
    std::vector<uint8_t> class_pack;     // this rank's share of the classes (wq_classes_build_local)
    std::vector<wq_event> ev_out;        // the finished records in ABI form (wq_process_events copies them, _view lends them)
    float psi_kernel_ms = 0.f;   // device time of the PSI EM launches of the last wq_process_events (CUDA events on the stream)
    uint64_t psi_sweep_bytes = 0;
    std::vector<WqEvPiece> ev_pieces; uint32_t ev_pieces_key[3] = {0xFFFFFFFFu, 0, 0};      // work list of the event build (index + partition only)
    std::vector<unsigned> ev_cuts;                              // the ranks' ranges of that list for the current counts (empty: not computed yet)
    WqEvBatch ev_batch[2];       // event problems of the genes no multi-map read touches (built while the transcript EM runs) / of the others
    DevBuf<uint64_t> d_ck, d_cv, d_ck2, d_cv2; DevBuf<uint32_t> d_ckoff; DevBuf<uint8_t> d_cubtmp;
    DevBuf<uint64_t> d_xpack; DevBuf<uint32_t> d_xpending;      // multi-GPU exchange: this rank's packed sparse stores / merge queue
    uint32_t part = 0, nparts = 1;                              // gene partition of the event stage (wq_set_gene_partition)
    // multi-GPU exchange inside the library (wq_comm_init_rank / wq_counts_sync / wq_classes_sync)
    void* nccl_comm = nullptr; int comm_rank = 0, comm_size = 1;
    DevBuf<uint64_t> d_xgather; uint64_t xcap = 1ull << 20;    // words per rank of the sparse all-gather (same on every rank: derived from shared headers)
    uint64_t* h_xhdr = nullptr;                                // pinned: the ranks' pack headers
    DevBuf<uint8_t> d_cshare, d_cgather; uint64_t ccap = 1ull << 20; uint8_t* h_cgather = nullptr; size_t h_cgather_cap = 0;
    // --biascorrect (csrc/wq_bias.cuh): model tallies, record buffer (grows, contents kept), per-slot results
    bool bias_on = false;
    DevBuf<uint64_t> d_bias_tab;                                // fore[4096], back[4096], cursor[2]
    uint64_t* d_bias_rec = nullptr; uint64_t bias_rec_cap = 0, bias_rec_used = 0;
    DevBuf<double> d_bias_model;                                // ratio[4096], gcval[32], fore_n[4096], back_n[4096]
    DevBuf<double> d_bias_slot;                                 // node_p[N] node_g[N] edge_p[S] edge_g[S] keyed_p[K] keyed_g[K]
    uint8_t bias_gcbin[WQ_BIAS_GC_SIZE + 1];
    bool hq_valid = false;
    uint32_t chunk_reads = 1u << 22;
    float last_kernel_ms[3] = {0, 0, 0};
    uint64_t launches = 0;       // kernels launched by this handle so far
    uint32_t last_hit_cap = 0;
};

// ------------------------------------------------------------------------------------------ kernels
#ifndef WQ_EXT_MINBLOCKS
#define WQ_EXT_MINBLOCKS 8
#endif
#ifndef WQ_SEED_MINBLOCKS
#define WQ_SEED_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(128, WQ_SEED_MINBLOCKS) wq_k_seed(const __grid_constant__ WqDevIndex ix, const __grid_constant__ WqParams p,
                                                 const WqBatchDev b, const WqWork wk) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ uint64_t s_w[128 * (WQ_MAX_READ_WORDS + 1)];
    if (r >= b.n) return;
    wq_seed_thread(ix, p, b, wk, r, s_w + threadIdx.x * (WQ_MAX_READ_WORDS + 1));
}

// First pass: one thread per located hit, working alignment in shared memory (WQ_FAST_PATH nodes per thread, node i of
// thread t at s_path[i * 128 + t]: conflict-free).  Tasks that need more (a longer path, or a parked candidate) are queued.
__global__ void __launch_bounds__(128, WQ_EXT_MINBLOCKS) wq_k_extend(const __grid_constant__ WqDevIndex ix, const __grid_constant__ WqParams p,
                                                                    const WqBatchDev b, const WqWork wk) {
    __shared__ uint64_t s_w[128 * (WQ_MAX_READ_WORDS + 1)];
    __shared__ WqPNode s_path[WQ_FAST_PATH * 128];
    const WqPathStore st{s_path + threadIdx.x, 128, WQ_FAST_PATH, nullptr, 0};
    const uint32_t nf = wk.cursors[0], nr = wk.cursors[5];        // forward-strand hits from the front, reverse-strand from the back
#pragma unroll 1
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nf + nr; t += gridDim.x * blockDim.x) {
        const uint32_t slot = t < nf ? t : wk.hit_cap - 1 - (t - nf);
        wq_extend_thread(ix, p, b, wk, slot, s_w + threadIdx.x * (WQ_MAX_READ_WORDS + 1), st, true);
    }
}
// Second pass over the queued tasks: full-size path store and one side buffer per nesting level in a global slab
// (thread t: path node i at slab[i * T + t], side buffers behind the T * WQ_MAX_PATH path nodes).
#define WQ_SLAB_NODES (WQ_MAX_PATH * (1 + WQ_MAX_DEPTH))
__global__ void __launch_bounds__(128) wq_k_extend2(const __grid_constant__ WqDevIndex ix, const __grid_constant__ WqParams p,
                                                    const WqBatchDev b, const WqWork wk, WqPNode* slab) {
    __shared__ uint64_t s_w[128 * (WQ_MAX_READ_WORDS + 1)];
    const uint32_t T = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const WqPathStore st{slab + tid, (int32_t)T, WQ_MAX_PATH, slab + (size_t)WQ_MAX_PATH * T + tid, (int32_t)T};
    const uint32_t nd = wk.cursors[6];
    for (uint32_t t = tid; t < nd; t += T)
        wq_extend_thread(ix, p, b, wk, wk.defer[t], s_w + threadIdx.x * (WQ_MAX_READ_WORDS + 1), st, false);
}

// what the running mean_readlen needs from a chunk: a mapped bit per read, and whether all mapped reads have one length
__device__ __forceinline__ void wq_note_mapped(const WqWork& wk, uint32_t r, bool mapped, uint32_t len) {
    const unsigned bal = __ballot_sync(0xffffffffu, mapped);
    const unsigned lmax = __reduce_max_sync(0xffffffffu, mapped ? len : 0u);
    const unsigned lmin = __reduce_max_sync(0xffffffffu, mapped ? 255u - len : 0u);
    if ((threadIdx.x & 31) == 0) {
        wk.mapped_bits[r >> 5] = bal;
        if (bal) { atomicMax(&wk.cursors[3], lmin); atomicMax(&wk.cursors[4], lmax); atomicAdd(&wk.cursors[7], (unsigned)__popc(bal)); }
    }
}

__global__ void __launch_bounds__(128) wq_k_seed_pe(const __grid_constant__ WqDevIndex ix, const __grid_constant__ WqParams p,
                                                    const WqBatchDev bf, const WqBatchDev br, const WqWork wk) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ uint64_t s_w[128 * (WQ_MAX_READ_WORDS + 1)];
    if (r >= bf.n) return;
    wq_seed_pe_thread(ix, p, bf, br, wk, r, s_w + threadIdx.x * (WQ_MAX_READ_WORDS + 1));
}

__global__ void __launch_bounds__(128, WQ_EXT_MINBLOCKS) wq_k_extend_pe(const __grid_constant__ WqDevIndex ix, const __grid_constant__ WqParams p,
                                                      const WqBatchDev bf, const WqBatchDev br, const WqWork wk) {
    uint32_t nh = wk.cursors[0];
    if (nh > wk.hit_cap) nh = wk.hit_cap;
    __shared__ uint64_t s_w[128 * (WQ_MAX_READ_WORDS + 1)];
    __shared__ WqPNode s_path[WQ_FAST_PATH * 128];
    const WqPathStore st{s_path + threadIdx.x, 128, WQ_FAST_PATH, nullptr, 0};
#pragma unroll 1
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < 2 * nh; t += gridDim.x * blockDim.x)
        wq_extend_pe_thread(ix, p, bf, br, wk, t >> 1, (int)(t & 1), s_w + threadIdx.x * (WQ_MAX_READ_WORDS + 1), st, true);
}
__global__ void __launch_bounds__(128) wq_k_extend2_pe(const __grid_constant__ WqDevIndex ix, const __grid_constant__ WqParams p,
                                                       const WqBatchDev bf, const WqBatchDev br, const WqWork wk, WqPNode* slab) {
    __shared__ uint64_t s_w[128 * (WQ_MAX_READ_WORDS + 1)];
    const uint32_t T = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const WqPathStore st{slab + tid, (int32_t)T, WQ_MAX_PATH, slab + (size_t)WQ_MAX_PATH * T + tid, (int32_t)T};
    const uint32_t nd = wk.cursors[6];
    for (uint32_t t = tid; t < nd; t += T)
        wq_extend_pe_thread(ix, p, bf, br, wk, wk.defer[t] >> 1, (int)(wk.defer[t] & 1), s_w + threadIdx.x * (WQ_MAX_READ_WORDS + 1), st, false);
}

__global__ void __launch_bounds__(128) wq_k_resolve_pe(const __grid_constant__ WqDevIndex ix, const __grid_constant__ WqParams p,
                                                       const WqBatchDev bf, const WqBatchDev br, const WqWork wk, const WqQuantDev q) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t st[5] = {0, 0, 0, 0, 0};
    if (wk.cursors[1] > wk.pool_cap) return;      // path pool exhausted: nothing of this chunk is counted, the host re-runs it with a larger pool
    if (r < bf.n) wq_resolve_pe_thread(ix, p, bf, br, wk, q, r, st);
    wq_note_mapped(wk, r, st[0] != 0, (uint32_t)st[2]);
#pragma unroll
    for (int i = 0; i < 5; i++) {
        unsigned long long v = st[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        st[i] = v;
    }
    if ((threadIdx.x & 31) == 0) {
        if (st[0]) atomicAdd((unsigned long long*)&q.counters->mapped, (unsigned long long)st[0]);
        if (st[1]) atomicAdd((unsigned long long*)&q.counters->multi, (unsigned long long)st[1]);
        if (st[2]) atomicAdd((unsigned long long*)&q.counters->sum_readlen, (unsigned long long)st[2]);
        if (st[3]) atomicAdd((unsigned long long*)&q.counters->path_overflow, (unsigned long long)st[3]);
        if (st[4]) atomicAdd((unsigned long long*)&q.counters->pool_overflow, (unsigned long long)st[4]);
    }
    if (r == 0) atomicAdd((unsigned long long*)&q.counters->total, (unsigned long long)bf.n);
}

__global__ void __launch_bounds__(128) wq_k_retry_pe(const __grid_constant__ WqDevIndex ix, const WqWork wk, const WqQuantDev q,
                                                     const uint32_t* list, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    wq_retry_pe_thread(ix, wk, q, list[i]);
}

__global__ void __launch_bounds__(128) wq_k_resolve(const __grid_constant__ WqDevIndex ix, const __grid_constant__ WqParams p,
                                                    const WqBatchDev b, const WqWork wk, const WqQuantDev q) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t st[5] = {0, 0, 0, 0, 0};
    if (wk.cursors[1] > wk.pool_cap) return;      // path pool exhausted: nothing of this chunk is counted, the host re-runs it with a larger pool
    if (r < b.n) wq_resolve_thread(ix, p, b, wk, q, r, st);
    wq_note_mapped(wk, r, st[0] != 0, (uint32_t)st[2]);
    // warp-level reduction of the scalar stats, one atomic per warp and counter
#pragma unroll
    for (int i = 0; i < 5; i++) {
        unsigned long long v = st[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        st[i] = v;
    }
    if ((threadIdx.x & 31) == 0) {
        if (st[0]) atomicAdd((unsigned long long*)&q.counters->mapped, (unsigned long long)st[0]);
        if (st[1]) atomicAdd((unsigned long long*)&q.counters->multi, (unsigned long long)st[1]);
        if (st[2]) atomicAdd((unsigned long long*)&q.counters->sum_readlen, (unsigned long long)st[2]);
        if (st[3]) atomicAdd((unsigned long long*)&q.counters->path_overflow, (unsigned long long)st[3]);
        if (st[4]) atomicAdd((unsigned long long*)&q.counters->pool_overflow, (unsigned long long)st[4]);
    }
    if (r == 0) atomicAdd((unsigned long long*)&q.counters->total, (unsigned long long)b.n);
}

__global__ void __launch_bounds__(128) wq_k_retry(const __grid_constant__ WqDevIndex ix, const WqWork wk, const WqQuantDev q,
                                                  const uint32_t* list, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    wq_retry_thread(ix, wk, q, list[i]);
}

// compaction of the open-addressing count tables before download: occupied slots -> dense arrays
// (oslot, when given, receives the table slot of every compacted record: the identity the bias tallies are filed under)
__global__ void wq_k_compact_edges(const uint64_t* key, const uint64_t* count, uint64_t slots, uint64_t* okey, uint64_t* ocount, uint32_t* cursor,
                                   uint32_t* oslot = nullptr) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < slots; i += stride) {
        uint64_t k = key[i];
        if (k == WQ_EMPTY_KEY) continue;
        uint32_t o = atomicAdd(cursor, 1u);
        okey[o] = k;
        ocount[o] = count[i];
        if (oslot) oslot[o] = (uint32_t)i;
    }
}
__global__ void wq_k_compact_keyed(const uint64_t* hash, const uint32_t* keyoff, const uint64_t* count, uint64_t slots,
                                   uint32_t* ooff, uint64_t* ocount, uint32_t* cursor, uint32_t* oslot = nullptr) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < slots; i += stride) {
        if (hash[i] == 0) continue;
        uint32_t o = atomicAdd(cursor, 1u);
        ooff[o] = keyoff[i];
        ocount[o] = count[i];
        if (oslot) oslot[o] = (uint32_t)i;
    }
}

// Multi-GPU exchange: another rank's compacted sparse stores (wq_sparse_device_pack layout) are added into this rank's
// device tables.  Pass 0 queues keyed records whose slot owner has not published its key yet; pass 1 (a later launch, so
// every owner has published) adds the queued ones.
__global__ void __launch_bounds__(256) wq_k_merge_sparse(const WqQuantDev q, const uint64_t* __restrict__ buf, uint32_t* pending,
                                                         uint32_t* cursor, const int pass) {
    const uint64_t ne = buf[0], nk = buf[1];
    const uint64_t* ekey = buf + 4;
    const uint64_t* ecount = ekey + ne;
    const uint64_t* kcount = ecount + ne;
    const uint32_t* koff = (const uint32_t*)(kcount + nk);
    const uint32_t* pool = koff + ((nk + 1) & ~1ull);
    const uint64_t total = pass == 0 ? ne + nk : (uint64_t)*cursor;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t j;
        if (pass == 0) {
            if (i < ne) { if (wq_edge_insert(q.edges, ekey[i], ecount[i])) atomicAdd((unsigned long long*)&q.counters->edge_full, 1ull); continue; }
            j = i - ne;
        } else j = pending[i];
        const uint32_t off = koff[j];
        const int rc = wq_keyed_insert(q.keyed, pool + off + 1, (int)pool[off], kcount[j]);
        if (rc == 1 && pass == 0) pending[atomicAdd(cursor, 1u)] = (uint32_t)j;
        else if (rc != 0) atomicAdd((unsigned long long*)&q.counters->keyed_full, 1ull);
    }
}

// k-mer table of the seeding kernel: rows of every ftab_k-mer (the first ftab_k steps of any backward search in one lookup)
__global__ void __launch_bounds__(256) wq_k_ftab(const __grid_constant__ WqDevIndex ix, uint2* out, const int k) {
    const uint32_t n = 1u << (2 * k);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = wq_ftab_entry(ix, i, k);
}

// Sparse stores of this rank, compacted (wq_k_compact_*), laid out as ONE record for the all-gather without the host knowing
// the sizes: {n_edge, n_keyed, pool_words, total_words}, edge_key[ne], edge_count[ne], keyed_count[nk], then u32 keyed_off[nk]
// (padded to even) and pool[pool_words] (padded to even).  A record longer than `cap` words carries its header only.
__device__ __forceinline__ uint64_t wq_pack_words(uint64_t ne, uint64_t nk, uint64_t kw) {
    return 4 + 2 * ne + nk + ((nk + 1) & ~1ull) / 2 + ((kw + 1) & ~1ull) / 2;
}
__global__ void __launch_bounds__(256) wq_k_pack_sparse(uint64_t* __restrict__ buf, const uint64_t cap, const uint32_t* __restrict__ counts,
                                                        const uint64_t* __restrict__ kcursor, const uint64_t* __restrict__ ck,
                                                        const uint64_t* __restrict__ cv, const uint64_t* __restrict__ kcnt,
                                                        const uint32_t* __restrict__ koff, const uint32_t* __restrict__ pool) {
    const uint64_t ne = counts[0], nk = counts[1], kw = *kcursor;
    const uint64_t total = wq_pack_words(ne, nk, kw);
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, T = (uint64_t)gridDim.x * blockDim.x;
    if (tid == 0) { buf[0] = ne; buf[1] = nk; buf[2] = kw; buf[3] = total; }
    if (total > cap) return;
    uint64_t* ekey = buf + 4;
    uint64_t* ecount = ekey + ne;
    uint64_t* kcount = ecount + ne;
    uint32_t* ko = (uint32_t*)(kcount + nk);
    uint32_t* pl = ko + ((nk + 1) & ~1ull);
    for (uint64_t i = tid; i < ne; i += T) { ekey[i] = ck[i]; ecount[i] = cv[i]; }
    for (uint64_t i = tid; i < nk; i += T) { kcount[i] = kcnt[i]; ko[i] = koff[i]; }
    for (uint64_t i = tid; i < kw; i += T) pl[i] = pool[i];
}
// every rank's header says its record fits: the merge of all peers happens, or none of it (all ranks see the same headers)
__device__ __forceinline__ bool wq_gather_ok(const uint64_t* g, uint64_t cap, int world) {
    for (int r = 0; r < world; r++) if (g[(uint64_t)r * cap + 3] > cap) return false;
    return true;
}
__global__ void __launch_bounds__(256) wq_k_merge_gathered(const WqQuantDev q, const uint64_t* __restrict__ g, const uint64_t cap, const int world,
                                                           const int peer, uint32_t* pending, uint32_t* cursor, const int pass) {
    if (!wq_gather_ok(g, cap, world)) return;
    const uint64_t* buf = g + (uint64_t)peer * cap;
    const uint64_t ne = buf[0], nk = buf[1];
    const uint64_t* ekey = buf + 4;
    const uint64_t* ecount = ekey + ne;
    const uint64_t* kcount = ecount + ne;
    const uint32_t* koff = (const uint32_t*)(kcount + nk);
    const uint32_t* pool = koff + ((nk + 1) & ~1ull);
    const uint64_t total = pass == 0 ? ne + nk : (uint64_t)*cursor;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t j;
        if (pass == 0) {
            if (i < ne) { if (wq_edge_insert(q.edges, ekey[i], ecount[i])) atomicAdd((unsigned long long*)&q.counters->edge_full, 1ull); continue; }
            j = i - ne;
        } else j = pending[i];
        const uint32_t off = koff[j];
        const int rc = wq_keyed_insert(q.keyed, pool + off + 1, (int)pool[off], kcount[j]);
        if (rc == 1 && pass == 0) pending[atomicAdd(cursor, 1u)] = (uint32_t)j;
        else if (rc != 0) atomicAdd((unsigned long long*)&q.counters->keyed_full, 1ull);
    }
}

__global__ void wq_k_fill_u64(uint64_t* p, uint64_t v, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// ------------------------------------------------------------------------------------------ helpers
static uint64_t env_u64(const char* name, uint64_t def) {
    const char* v = getenv(name);
    if (!v || !*v) return def;
    return strtoull(v, nullptr, 10);
}
static uint64_t pow2_at_least(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

template <class T>
static cudaError_t upload(DevBuf<T>& d, const T* src, size_t n, cudaStream_t s) {
    cudaError_t e = d.ensure(n ? n : 1);
    if (e != cudaSuccess) return e;
    if (n) e = cudaMemcpyAsync(d.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s);
    return e;
}

extern "C" {

int wq_comm_destroy(wq_handle* h);
const char* wq_last_error(void) { return g_err.c_str(); }
int wq_version(void) { return 100; }

int wq_pinned_alloc(void** p, uint64_t bytes) {
    CK(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}
int wq_pinned_free(void* p) {
    CK(cudaFreeHost(p));
    return 0;
}

void wq_destroy(wq_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    h->d_occ.release(); h->d_sa.release(); h->d_text2.release(); h->d_amb.release(); h->d_nodes.release();
    h->d_genes.release(); h->d_bucket.release(); h->d_lptr.release(); h->d_rptr.release(); h->d_lent.release();
    h->d_rent.release(); h->d_ftab.release(); h->d_dense.release(); h->d_ekey.release(); h->d_ecount.release(); h->d_khash.release();
    h->d_kcount.release(); h->d_koff.release(); h->d_kpool.release(); h->d_kcursor.release();
    for (int i = 0; i < 2; i++) {
        h->d_len[i].release(); h->d_qual[i].release(); h->d_qualpk[i].release(); h->d_seqoff[i].release(); h->d_qualoff[i].release(); h->d_seq2[i].release();
        if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    }
    h->d_seeds.release(); h->d_hit_read.release(); h->d_hit_s.release(); h->d_hit_s2.release(); h->d_hit_gene.release(); h->d_pending.release(); h->d_pending2.release();
    h->d_cursors.release(); h->d_hits.release(); h->d_pool.release(); h->d_defer.release(); h->d_slab.release(); h->d_mapped_bits.release();
    h->em_scratch.release(); h->ev_batch[0].release(); h->ev_batch[1].release(); h->d_ck.release(); h->d_cv.release(); h->d_ck2.release(); h->d_cv2.release(); h->d_ckoff.release(); h->d_cubtmp.release(); h->d_xpack.release(); h->d_xpending.release();
    wq_comm_destroy(h);
    h->d_bias_tab.release(); h->d_bias_model.release(); h->d_bias_slot.release();
    if (h->d_bias_rec) cudaFree(h->d_bias_rec);
    h->d_xgather.release(); h->d_cshare.release(); h->d_cgather.release();
    if (h->h_xhdr) cudaFreeHost(h->h_xhdr);
    if (h->h_cgather) cudaFreeHost(h->h_cgather);
    if (h->h_cursors) cudaFreeHost(h->h_cursors);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
}


// the derived quant state (classes, EM results, event batches) no longer matches the count stores
static void reset_quant_ex(wq_handle* h) {
    for (WqEvBatch& B : h->ev_batch) {
        if (B.launched) cudaStreamSynchronize(B.stream);        // its result copies target B's pinned buffers
        B.built = B.launched = false;
    }
    h->hx = WqHostQuantEx();
    h->ev_cuts.clear();
    h->events_done = false;
}

// Event construction (src/events.jl:760-800) for the genes of one group on the host threads, then the parts' EM problems
// are concatenated (prefix sums + parallel copies).  Needs finalized edges.
static void wq_build_event_batch(wq_handle* h, int64_t readlen, int group, WqEvBatch& B) {
    const WqHostQuantEx& X = h->hx;
    static const bool prof2 = getenv("WQ_PROFILE") && atoi(getenv("WQ_PROFILE")) >= 2;
    const auto tb0 = std::chrono::steady_clock::now();
    auto lap2 = [&](const char* what) { if (prof2) fprintf(stderr, "[wq]       batch %d %-20s %9.3f ms\n", group, what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count()); };
    const uint32_t G = h->H.G;
    unsigned nt = wq_host_threads();
    if (G < 2048) nt = 1;
    // work list: see wq_event_make_pieces
    typedef WqEvPiece Piece;
    if (h->ev_pieces_key[2] != nt || h->ev_pieces_key[1] != h->nparts || h->ev_pieces.empty()) {
        wq_event_make_pieces(h->H, X.node_value, nt, h->nparts, h->ev_pieces);      // cached in the handle: depends on the index, the host thread count and the partition only
        h->ev_pieces_key[0] = 0; h->ev_pieces_key[1] = h->nparts; h->ev_pieces_key[2] = nt;
    }
    const std::vector<Piece>& pieces = h->ev_pieces;
    const unsigned npieces = (unsigned)pieces.size();
    // This rank's range of pieces (wq_set_gene_partition): contiguous, so the ranks' records concatenate to the whole output; every rank
    // holds the same counts after wq_counts_sync and the same piece list, so every rank computes the same cuts (wq_event_cuts).
    unsigned pc_lo = 0, pc_hi = npieces;
    if (h->nparts > 1) {
        // from the RAW count store (identical on every rank after the exchange, and the same for both gene groups of a job: the merged
        // values in X change when assign_ambig! runs between the two builds), once per job
        if (h->ev_cuts.size() != (size_t)h->nparts + 1) {
            std::vector<double> raw(h->hq.edge_count.size());
            for (size_t i = 0; i < raw.size(); i++) raw[i] = (double)h->hq.edge_count[i];
            wq_event_cuts(h->H, h->hq.edge_key, raw, pieces, h->nparts, h->ev_cuts);
        }
        pc_lo = h->ev_cuts[h->part]; pc_hi = h->ev_cuts[h->part + 1];
    }
    // the group of genes that multi-map classes name is often empty (no multi-mapping reads in the job): nothing to build or launch then
    bool any = group == 0;
    if (group == 1) for (unsigned t = pc_lo; t < pc_hi && !any; t++) for (uint32_t g = pieces[t].g_lo; g <= pieces[t].g_hi && g <= G && !any; g++) any = X.gene_dirty[g] != 0;
    if (!any) {
        B.parts.clear();
        B.parts.resize(npieces);
        B.P = WqPsiProblems();
        B.pshift.assign(npieces, 0);
        B.ci_psi.clear(); B.ci_tot.clear(); B.p_total.clear();
        B.built = true; B.launched = false; B.readlen = readlen;
        return;
    }
    B.parts.clear();
    B.parts.resize(npieces);
    auto run = [&](const std::function<void(unsigned)>& f) {
        if (nt == 1) { for (unsigned t = 0; t < npieces; t++) f(t); return; }
        std::atomic<unsigned> next{0};
        std::vector<std::thread> th;
        for (unsigned k = 0; k < nt; k++) th.emplace_back([&]() { for (unsigned t; (t = next.fetch_add(1)) < npieces;) f(t); });
        for (auto& x : th) x.join();
    };
    run([&](unsigned t) {
        if (t < pc_lo || t >= pc_hi) return;                 // another rank's piece
        const Piece pc = pieces[t];
        const uint32_t g_lo = pc.g_lo, g_hi = pc.g_hi;
        size_t e = std::lower_bound(X.edge_key.begin(), X.edge_key.end(), (uint64_t)g_lo << 32) - X.edge_key.begin();
        WqGeneEvents V{h->H, X.node_value, 0, {}};
        WqEvPart& part = B.parts[t];
        for (uint32_t g = g_lo; g <= g_hi; g++) {
            if ((int)(X.gene_dirty[g] != 0) != group) continue;
            V.g = g;
            V.edges.clear();
            while (e < X.edge_key.size() && (X.edge_key[e] >> 32) < g) e++;
            for (; e < X.edge_key.size() && (X.edge_key[e] >> 32) == g; e++) {
                if (wq_edge_is_circ(X.edge_key[e])) continue;
                V.edges.push_back(WqGeneEdge{(uint32_t)((X.edge_key[e] >> 16) & 0xFFFF), (uint32_t)(X.edge_key[e] & 0xFFFF), X.edge_value[e]});
            }
            wq_gene_events(V, readlen, part.recs, part.P, part.pool, pc.node_lo, pc.node_hi);
        }
        // Beta-CI inputs of this part: spliced events whose psi is known without EM get a launch slot (part-local for now);
        // EM problems of spliced events carry their total (tandem UTR problems: -1, no CI)
        part.ci_psi.clear(); part.ci_tot.clear();
        part.p_total.assign(part.P.size(), -1.0);
        for (auto& r : part.recs) {
            r.batch = (uint8_t)group;
            if (r.problem >= 0 && r.kind != 2) part.p_total[(size_t)r.problem] = r.total;
            if (r.kind == 1 && r.problem < 0 && r.total > 0 && r.psi >= 0.0 && r.psi <= 1.0) {
                r.ci_slot = (int32_t)part.ci_psi.size();
                part.ci_psi.push_back(r.psi); part.ci_tot.push_back(r.total);
            }
        }
    });
    lap2("genes built");
    std::vector<WqPsiProblems::Sizes> at(npieces + 1);
    B.pshift.assign(npieces, 0);
    for (unsigned t = 0; t < npieces; t++) {
        const WqPsiProblems::Sizes z = B.parts[t].P.sizes();
        at[t + 1].ev = at[t].ev + z.ev; at[t + 1].paths = at[t].paths + z.paths; at[t + 1].amb = at[t].amb + z.amb; at[t + 1].amb_paths = at[t].amb_paths + z.amb_paths;
        B.pshift[t] = (long)at[t].ev;
    }
    std::vector<size_t> ci_at(npieces + 1, 0);
    for (unsigned t = 0; t < npieces; t++) ci_at[t + 1] = ci_at[t] + B.parts[t].ci_psi.size();
    B.P.resize_total(at[npieces]);
    B.ci_psi.resize(ci_at[npieces]); B.ci_tot.resize(ci_at[npieces]);
    B.p_total.resize(at[npieces].ev);
    run([&](unsigned t) {
        WqEvPart& part = B.parts[t];
        B.P.place(part.P, at[t]);
        std::copy(part.ci_psi.begin(), part.ci_psi.end(), B.ci_psi.begin() + ci_at[t]);
        std::copy(part.ci_tot.begin(), part.ci_tot.end(), B.ci_tot.begin() + ci_at[t]);
        std::copy(part.p_total.begin(), part.p_total.end(), B.p_total.begin() + at[t].ev);
        if (ci_at[t]) for (auto& r : part.recs) if (r.ci_slot >= 0) r.ci_slot += (int32_t)ci_at[t];
    });
    lap2("problems concatenated");
    lap2("CI inputs collected");
    B.built = true;
    B.launched = false;
    B.readlen = readlen;
}

// one batched PSI EM launch for every problem of the group; results land in the group's pinned buffers when the stream
// reaches them (no host wait here)
static int wq_launch_event_batch(wq_handle* h, WqEvBatch& B) {
    WqPsiProblems& P = B.P;
    B.launched = false;
    CK(B.ensure_stream());
    if (const size_t nci = B.ci_psi.size()) {
        if (2 * nci > B.ci_cap) { if (B.ci_out) cudaFreeHost(B.ci_out); B.ci_out = nullptr; B.ci_cap = 0; CK(cudaMallocHost(&B.ci_out, (2 * nci + 64) * 8)); B.ci_cap = 2 * nci + 64; }
        std::string err = wq_run_beta_ci((uint32_t)nci, B.ci_psi.data(), B.ci_tot.data(), B.ci_out, B.ci_out + nci, B.dev, B.stream, &h->launches, /*wait=*/false);
        if (!err.empty()) return fail(-13, err);
        B.launched = true;
    }
    if (!P.size()) return 0;
    CK(B.ensure_pinned(P.count.size(), P.size()));
    wq_psi_batch b;
    memset(&b, 0, sizeof b);
    b.n_events = (uint32_t)P.size();
    b.path_ptr = P.path_ptr.data(); b.n_inc = P.n_inc.data(); b.path_count = P.count.data(); b.path_length = P.length.data();
    b.amb_ptr = P.amb_ptr.data(); b.amb_path_ptr = P.amb_path_ptr.data(); b.amb_paths = P.amb_paths.data(); b.amb_mult = P.amb_mult.data();
    b.is_tandem = P.is_tandem.data(); b.readlen = B.readlen; b.sig = 4; b.maxit = 1500;
    b.psi = B.psi; b.iterations = B.its;
    const size_t E = P.size();
    if (2 * E > B.ci_em_cap) { if (B.ci_em) cudaFreeHost(B.ci_em); B.ci_em = nullptr; B.ci_em_cap = 0; CK(cudaMallocHost(&B.ci_em, (2 * E + 64) * 8)); B.ci_em_cap = 2 * E + 64; }
    std::string err = wq_psi_ci_upload((uint32_t)E, B.p_total.data(), B.dev, B.stream);
    if (err.empty()) err = wq_run_em_psi(b, B.dev, B.stream, &h->launches, /*wait=*/false);
    if (err.empty()) err = wq_run_psi_ci((uint32_t)E, B.ci_em, B.dev, B.stream, &h->launches);
    if (!err.empty()) return fail(-13, err);
    B.launched = true;
    return 0;
}

int wq_quant_reset(wq_handle* h) {
    if (!h) return fail(-1, "null handle");
    WqTimer tm_("quant_reset");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    CK(cudaMemsetAsync(h->d_dense.p, 0, (h->H.N + sizeof(WqCounters) / 8) * sizeof(uint64_t), s));
    wq_k_fill_u64<<<592, 256, 0, s>>>(h->d_ekey.p, WQ_EMPTY_KEY, h->edge_slots);
    h->launches += 1;
    CK(cudaMemsetAsync(h->d_ecount.p, 0, h->edge_slots * 8, s));
    CK(cudaMemsetAsync(h->d_khash.p, 0, h->keyed_slots * 8, s));
    CK(cudaMemsetAsync(h->d_kcount.p, 0, h->keyed_slots * 8, s));
    CK(cudaMemsetAsync(h->d_koff.p, 0xFF, h->keyed_slots * 4, s));
    CK(cudaMemsetAsync(h->d_kcursor.p, 0, 8, s));
    if (h->bias_on) { CK(cudaMemsetAsync(h->d_bias_tab.p, 0, (2 * WQ_BIAS_NHEX + 2) * 8, s)); h->bias_rec_used = 0; }
    CK(cudaStreamSynchronize(s));
    h->hq = WqHostQuant();
    reset_quant_ex(h);
    h->hq_valid = false;
    h->rl_mean = 0.0; h->rl_mapped = 0;
    return 0;
}

// CounterType = JointBiasCounter, BiasModelType = JointBiasMod (bin/whippet-quant.jl:116-122): chosen before the first read
int wq_set_biascorrect(wq_handle* h, int on) {
    if (!h) return fail(-1, "null handle");
    CK(cudaSetDevice(h->device));
    if (on && h->nccl_comm) return fail(-16, "bias correction is single-GPU in this version: the per-counter tallies are keyed by table slot, which differs between ranks");
    h->bias_on = on != 0;
    if (h->bias_on) {
        CK(h->d_bias_tab.ensure(2 * WQ_BIAS_NHEX + 2));
        wq_bias_gc_table(h->bias_gcbin);
    }
    return wq_quant_reset(h);
}

// the large arrays of the index as they are uploaded: views into WqHostIndex + the caller's wq_index_desc (wq_index_create), or into
// a mapped index file (wq_index_load)
struct WqBigArrays {
    const WqOccBlock* occ; size_t n_occ; const uint32_t* sa; const uint64_t* text2; size_t n_text2; const uint32_t* amb; size_t n_amb;
    const uint32_t* bucket; size_t n_bucket; const uint32_t* lptr; const uint32_t* rptr;
};

static int index_device_check(int device, cudaDeviceProp& prop) {
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev == 0)
        return fail(-2, std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e0));
    if (device < 0 || device >= ndev) return fail(-2, "device ordinal out of range");
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(-2, std::string("device '") + prop.name + "' is not sm_100-class; the kernels are built for sm_100a only");
    CK(cudaSetDevice(device));
    return 0;
}

// h->H (nodes, genes, edge-table entries, scalars), h->iso and h->sam are filled; upload everything and set up the count stores
static int index_finish(wq_handle* h, const cudaDeviceProp& prop, const WqBigArrays& A, wq_handle** out) {
#define CKH(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_); wq_destroy(h); return fail(-10, m_); } } while (0)
    CKH(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    CKH(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CKH(cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
        CKH(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
    CKH(cudaHostAlloc((void**)&h->h_cursors, 64, cudaHostAllocDefault));
    h->n_sm = prop.multiProcessorCount;
    h->pool_nodes_per_read = env_u64("WQ_POOL_NODES_PER_READ", 8);
    // spliced extensions nest (one level per junction found through the k-mer tables, at most WQ_MAX_DEPTH); the working alignment
    // lives in shared memory / a global slab, so a nested frame holds scalars only
    CKH(cudaDeviceSetLimit(cudaLimitStackSize, env_u64("WQ_STACK_BYTES", 4 * 1024)));
    WqHostIndex& H = h->H;
    cudaStream_t s = h->stream;
    CKH(upload(h->d_occ, A.occ, A.n_occ, s));
    CKH(upload(h->d_sa, A.sa, (size_t)H.n, s));
    CKH(upload(h->d_text2, A.text2, A.n_text2, s));
    if (H.has_amb) CKH(upload(h->d_amb, A.amb, A.n_amb, s));
    CKH(upload(h->d_nodes, H.nodes.data(), H.nodes.size(), s));
    CKH(upload(h->d_genes, H.genes.data(), H.genes.size(), s));
    CKH(upload(h->d_bucket, A.bucket, A.n_bucket, s));
    CKH(upload(h->d_lptr, A.lptr, (size_t)H.nslots + 1, s));
    CKH(upload(h->d_rptr, A.rptr, (size_t)H.nslots + 1, s));
    CKH(upload(h->d_lent, H.left_ent.data(), H.left_ent.size(), s));
    CKH(upload(h->d_rent, H.right_ent.data(), H.right_ent.size(), s));
    WqDevIndex& ix = h->dix;
    ix.occ = h->d_occ.p; ix.sa = h->d_sa.p; ix.text2 = h->d_text2.p; ix.amb = H.has_amb ? h->d_amb.p : nullptr;
    ix.nodes = h->d_nodes.p; ix.genes = h->d_genes.p; ix.node_bucket = h->d_bucket.p;
    ix.left_ptr = h->d_lptr.p; ix.left_ent = h->d_lent.p; ix.right_ptr = h->d_rptr.p; ix.right_ent = h->d_rent.p;
    ix.n = H.n; ix.sentinel = H.sentinel;
    for (int i = 0; i < 4; i++) ix.C[i] = H.C[i];
    ix.N = H.N; ix.G = H.G; ix.K = H.K;
    ix.ftab = nullptr; ix.ftab_k = 0;
    {   // k-mer table (default k = 10: 8 MB, resident in L2 beside the occurrence blocks)
        const int k = (int)env_u64("WQ_FTAB_K", 10);
        if (k > 0) {
            if (k > 13) { wq_destroy(h); return fail(-1, "WQ_FTAB_K must be at most 13"); }
            CKH(h->d_ftab.ensure((size_t)1 << (2 * k)));
            wq_k_ftab<<<4 * prop.multiProcessorCount, 256, 0, s>>>(ix, h->d_ftab.p, k);
            CKH(cudaGetLastError());
            h->launches += 1;
            ix.ftab = h->d_ftab.p; ix.ftab_k = k;
        }
    }
    // free the bulky host mirrors the quant stage does not need
    std::vector<WqOccBlock>().swap(H.occ);
    std::vector<uint64_t>().swap(H.text2);
    std::vector<uint32_t>().swap(H.amb);
    std::vector<uint32_t>().swap(H.node_bucket);
    // L2 persisting window over the occurrence blocks (the hot random-access structure of kernel 1)
    {
        size_t occ_bytes = h->d_occ.cap * sizeof(WqOccBlock);
        size_t maxp = (size_t)prop.persistingL2CacheMaxSize;
        if (maxp > 0 && env_u64("WQ_L2_PERSIST", 1)) {
            size_t want = std::min(occ_bytes, maxp);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof attr);
            attr.accessPolicyWindow.base_ptr = (void*)h->d_occ.p;
            attr.accessPolicyWindow.num_bytes = std::min(occ_bytes, (size_t)prop.accessPolicyMaxWindowSize);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)std::max<size_t>(1, attr.accessPolicyWindow.num_bytes));
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
            cudaGetLastError();
            h->l2attr = attr; h->has_l2attr = true;
        }
    }
    // count stores
    h->edge_slots = pow2_at_least(env_u64("WQ_EDGE_SLOTS", std::max<uint64_t>(1u << 16, 8ull * H.N)));
    h->keyed_slots = pow2_at_least(env_u64("WQ_KEYED_SLOTS", std::max<uint64_t>(1u << 16, 16ull * H.N)));
    h->kpool_cap = env_u64("WQ_KEYED_POOL_WORDS", std::max<uint64_t>(1u << 20, 128ull * H.N));
    CKH(h->d_dense.ensure(H.N + sizeof(WqCounters) / 8));
    CKH(h->d_ekey.ensure(h->edge_slots)); CKH(h->d_ecount.ensure(h->edge_slots));
    CKH(h->d_khash.ensure(h->keyed_slots)); CKH(h->d_kcount.ensure(h->keyed_slots)); CKH(h->d_koff.ensure(h->keyed_slots));
    CKH(h->d_kpool.ensure(h->kpool_cap)); CKH(h->d_kcursor.ensure(1));
    CKH(h->d_cursors.ensure(16));
    h->q.node_count = h->d_dense.p;
    h->q.counters = reinterpret_cast<WqCounters*>(h->d_dense.p + H.N);
    h->q.edges = WqEdgeTable{h->d_ekey.p, h->d_ecount.p, h->edge_slots - 1};
    h->q.keyed = WqKeyedTable{h->d_khash.p, h->d_koff.p, h->d_kcount.p, h->d_kpool.p, h->d_kcursor.p, h->kpool_cap, h->keyed_slots - 1};
    h->chunk_reads = (uint32_t)env_u64("WQ_CHUNK_READS", 1u << 21);
    CKH(cudaStreamSynchronize(s));                           // the uploads read caller / mapped memory: done before it goes away
    int rc = wq_quant_reset(h);
    if (rc) { wq_destroy(h); return rc; }
#undef CKH
    *out = h;
    return 0;
}


int wq_index_create(const wq_index_desc* d, int device, wq_handle** out) {
    if (!d || !out) return fail(-1, "null argument");
    *out = nullptr;
    cudaDeviceProp prop;
    if (int rc = index_device_check(device, prop)) return rc;
    wq_handle* h = new wq_handle();
    h->device = device;
    std::string err = wq_build_host_index(d, h->H);
    if (err.empty()) err = wq_build_iso_index(d, h->iso);
    if (err.empty()) {                                       // SAM output only
        h->sam.node_base.assign(d->node_base, d->node_base + d->n_genes + 1);
        h->sam.nodeoffset.assign(d->nodeoffset, d->nodeoffset + d->n_nodes);
        h->sam.nodecoord.assign(d->nodecoord, d->nodecoord + d->n_nodes);
        h->sam.nodelen.assign(d->nodelen, d->nodelen + d->n_nodes);
    }
    if (!err.empty()) { delete h; return fail(-3, err); }
    const WqHostIndex& H = h->H;
    const WqBigArrays A{H.occ.data(), H.occ.size(), d->sa, H.text2.data(), H.text2.size(), H.amb.data(), H.amb.size(),
                        H.node_bucket.data(), H.node_bucket.size(), d->left_ptr, d->right_ptr};
    return index_finish(h, prop, A, out);
}

// The index in its native on-disk form (csrc/wq_index_file.h): wq_index_write is host-only; wq_index_load maps the file and copies
// the sections to the device (no re-layout, no deserialisation).
int wq_index_write(const wq_index_desc* d, const char* const* seqname, const uint8_t* strand, const char* path) {
    if (!d || !path) return fail(-1, "null argument");
    const std::string err = wq_index_write_impl(d, seqname, strand, path);
    if (!err.empty()) return fail(-3, err);
    return 0;
}

int wq_index_load(const char* path, int device, wq_handle** out) {
    if (!path || !out) return fail(-1, "null argument");
    *out = nullptr;
    WqIdxMapped M;
    {
        const std::string err = wq_index_map(path, M, env_u64("WQ_INDEX_VERIFY", 0) != 0);
        if (!err.empty()) return fail(-3, err);
    }
    cudaDeviceProp prop;
    if (int rc = index_device_check(device, prop)) return rc;
    wq_handle* h = new wq_handle();
    h->device = device;
    const WqIdxHeader& hd = M.hd;
    WqHostIndex& H = h->H;
    H.n = hd.n; H.sentinel = hd.sentinel; H.nslots = hd.nslots; H.N = hd.N; H.G = hd.G; H.K = hd.K; H.has_amb = hd.has_amb != 0;
    for (int i = 0; i < 4; i++) H.C[i] = hd.C[i];
    H.nodes.assign(M.ptr<WqNodeRec>(WQ_SEC_NODES), M.ptr<WqNodeRec>(WQ_SEC_NODES) + hd.N);
    H.genes.assign(M.ptr<WqGeneRec>(WQ_SEC_GENES), M.ptr<WqGeneRec>(WQ_SEC_GENES) + hd.G);
    H.left_ent.assign(M.ptr<uint2>(WQ_SEC_LENT), M.ptr<uint2>(WQ_SEC_LENT) + M.count(WQ_SEC_LENT));
    H.right_ent.assign(M.ptr<uint2>(WQ_SEC_RENT), M.ptr<uint2>(WQ_SEC_RENT) + M.count(WQ_SEC_RENT));
    wq_index_desc d;                                         // the fields the annotation index reads
    memset(&d, 0, sizeof d);
    d.n_genes = hd.G; d.n_nodes = hd.N; d.n_isoforms = hd.I;
    d.node_base = M.ptr<uint32_t>(WQ_SEC_NODE_BASE); d.nodelen = M.ptr<uint32_t>(WQ_SEC_NODELEN);
    d.iso_base = M.ptr<uint32_t>(WQ_SEC_ISO_BASE); d.iso_node_ptr = M.ptr<uint32_t>(WQ_SEC_ISO_NODE_PTR); d.iso_nodes = M.ptr<uint32_t>(WQ_SEC_ISO_NODES);
    for (uint32_t g = 0; g < hd.G; g++)
        if (d.node_base[g + 1] < d.node_base[g] || d.node_base[g + 1] > hd.N) { delete h; return fail(-3, std::string(path) + ": node_base is not a CSR over the nodes"); }
    const std::string err = wq_build_iso_index(&d, h->iso);
    if (!err.empty()) { delete h; return fail(-3, err); }
    h->sam.node_base.assign(d.node_base, d.node_base + hd.G + 1);
    h->sam.nodeoffset.assign(M.ptr<uint32_t>(WQ_SEC_NODEOFFSET), M.ptr<uint32_t>(WQ_SEC_NODEOFFSET) + hd.N);
    h->sam.nodecoord.assign(M.ptr<uint32_t>(WQ_SEC_NODECOORD), M.ptr<uint32_t>(WQ_SEC_NODECOORD) + hd.N);
    h->sam.nodelen.assign(d.nodelen, d.nodelen + hd.N);
    if (hd.has_info) {
        const uint64_t* no = M.ptr<uint64_t>(WQ_SEC_NAME_OFF); const char* nm = M.ptr<char>(WQ_SEC_NAMES); const uint8_t* st = M.ptr<uint8_t>(WQ_SEC_STRAND);
        for (uint32_t g = 0; g < hd.G; g++) {
            if (no[g + 1] < no[g] || no[g + 1] > M.count(WQ_SEC_NAMES)) { delete h; return fail(-3, std::string(path) + ": malformed name table"); }
            h->sam.seqname.emplace_back(nm + no[g], (size_t)(no[g + 1] - no[g]));
            h->sam.strand.push_back(st[g] ? 1 : 0);
        }
        h->sam.has_info = true;
    }
    const WqBigArrays A{M.ptr<WqOccBlock>(WQ_SEC_OCC), (size_t)M.count(WQ_SEC_OCC), M.ptr<uint32_t>(WQ_SEC_SA), M.ptr<uint64_t>(WQ_SEC_TEXT2),
                        (size_t)M.count(WQ_SEC_TEXT2), M.ptr<uint32_t>(WQ_SEC_AMB), (size_t)M.count(WQ_SEC_AMB), M.ptr<uint32_t>(WQ_SEC_BUCKET),
                        (size_t)M.count(WQ_SEC_BUCKET), M.ptr<uint32_t>(WQ_SEC_LPTR), M.ptr<uint32_t>(WQ_SEC_RPTR)};
    return index_finish(h, prop, A, out);                    // synchronises before M unmaps the file
}

// Run the alignment kernels over reads [0, b.n) already resident on the device (br != NULL: read pairs).
static int run_chunk(wq_handle* h, const WqParams& p, const WqBatchDev& b, const WqBatchDev* br, bool time_kernels) {
    const uint32_t n = b.n;
    if (n == 0) return 0;
    const bool pe = br != nullptr;
    cudaStream_t s = h->stream;
    const uint64_t hit_cap64 = (uint64_t)n * (uint64_t)p.seed_tolerate;
    if (hit_cap64 >= 0x7FFFFFF0ull) return fail(-1, "chunk too large: reads per chunk x seed_tolerate must stay below 2^31 (lower WQ_CHUNK_READS)");
    const uint32_t hit_cap = (uint32_t)hit_cap64;
    const uint32_t ntask = hit_cap * (pe ? 2u : 1u);
    CK(h->d_seeds.ensure(n)); CK(h->d_hit_read.ensure(hit_cap)); CK(h->d_hit_s.ensure(hit_cap));
    CK(h->d_hits.ensure((size_t)ntask)); CK(h->d_defer.ensure((size_t)ntask));
    if (pe) { CK(h->d_hit_s2.ensure(hit_cap)); CK(h->d_hit_gene.ensure(hit_cap)); }
    CK(h->d_pending.ensure(n)); CK(h->d_pending2.ensure(n)); CK(h->d_mapped_bits.ensure(((size_t)n + 127) / 128 * 4));
    const unsigned grid2 = 2u * (unsigned)h->n_sm;                    // second-pass extension kernel: small persistent grid
    CK(h->d_slab.ensure((size_t)grid2 * 128 * WQ_SLAB_NODES));
    // the path pool is an average budget (a read may legally need seed_tolerate x WQ_MAX_PATH nodes); when a chunk exceeds it nothing of
    // the chunk is counted (the resolve kernel checks the cursor first) and the chunk runs again with the pool it asked for
    uint64_t pool_want = std::max<uint64_t>(h->pool_nodes_per_read * (uint64_t)n * (pe ? 2 : 1) + 1024, h->d_pool.cap);
    cudaEvent_t ev[4];
    if (time_kernels) for (auto& e : ev) CK(cudaEventCreate(&e));
    for (int attempt = 0;; attempt++) {
        const uint32_t pool_cap = (uint32_t)std::min<uint64_t>(pool_want, 0xFFFFFFF0ull);
        CK(h->d_pool.ensure(pool_cap));
        WqWork wk{h->d_seeds.p, h->d_hit_read.p, h->d_hit_s.p, h->d_hit_s2.p, h->d_hit_gene.p, h->d_hits.p, h->d_pool.p,
                  h->d_cursors.p, h->d_pending.p, h->d_pending2.p, hit_cap, pool_cap, h->d_defer.p, h->d_mapped_bits.p};
        CK(cudaMemsetAsync(h->d_cursors.p, 0, 16 * sizeof(uint32_t), s));
        const int T = 128;
        const uint32_t gread = (n + T - 1) / T;
        if (time_kernels) CK(cudaEventRecord(ev[0], s));
        if (!pe) wq_k_seed<<<gread, T, 0, s>>>(h->dix, p, b, wk);
        else wq_k_seed_pe<<<gread, T, 0, s>>>(h->dix, p, b, *br, wk);
        h->launches += 4;
        if (time_kernels) CK(cudaEventRecord(ev[1], s));
        // one thread per located hit; the hit count lives on the device, so the grid is sized for the usual
        // ~1 hit per read and strides over the rest
        if (!pe) {
            wq_k_extend<<<gread + gread / 8 + 1, T, 0, s>>>(h->dix, p, b, wk);
            wq_k_extend2<<<grid2, T, 0, s>>>(h->dix, p, b, wk, h->d_slab.p);
        } else {
            wq_k_extend_pe<<<2 * gread + gread / 4 + 1, T, 0, s>>>(h->dix, p, b, *br, wk);
            wq_k_extend2_pe<<<grid2, T, 0, s>>>(h->dix, p, b, *br, wk, h->d_slab.p);
        }
        if (time_kernels) CK(cudaEventRecord(ev[2], s));
        if (!pe) wq_k_resolve<<<gread, T, 0, s>>>(h->dix, p, b, wk, h->q);
        else wq_k_resolve_pe<<<gread, T, 0, s>>>(h->dix, p, b, *br, wk, h->q);
        if (time_kernels) CK(cudaEventRecord(ev[3], s));
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h->h_cursors, h->d_cursors.p, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        h->last_hit_cap = hit_cap;
        h->last_deferred += h->h_cursors[6];
        if (time_kernels)
            for (int i = 0; i < 3; i++) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); h->last_kernel_ms[i] += ms; }
        if (h->h_cursors[1] <= pool_cap) break;
        if (attempt >= 2 || pool_cap >= 0xFFFFFFF0u) {
            if (time_kernels) for (auto& e : ev) cudaEventDestroy(e);
            return fail(-5, "alignment path pool exhausted for this chunk (nothing of the chunk was counted; lower WQ_CHUNK_READS)");
        }
        pool_want = (uint64_t)h->h_cursors[1] + 1024;
    }
    if (time_kernels) for (auto& e : ev) cudaEventDestroy(e);
    // mean_readlen (src/reads.jl:69): `mean += (len - mean) / mapped` over the mapped reads in read order.  When every mapped read
    // of the chunk has the length the mean already has (fixed-length input), the recurrence adds exact zeros; otherwise it is
    // replayed here from the chunk's mapped bits.
    if (const uint32_t mc = h->h_cursors[7]) {
        const uint32_t lmin = 255u - h->h_cursors[3], lmax = h->h_cursors[4];
        if (lmin == lmax && (h->rl_mapped == 0 || h->rl_mean == (double)lmin)) {
            h->rl_mean = (double)lmin;
            h->rl_mapped += mc;
        } else {
            std::vector<uint32_t> bits(((size_t)n + 31) / 32);
            std::vector<uint8_t> lens(n);
            CK(cudaMemcpyAsync(bits.data(), h->d_mapped_bits.p, bits.size() * 4, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(lens.data(), b.len, n, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            double mean = h->rl_mean;
            uint64_t mapped = h->rl_mapped;
            for (uint32_t r = 0; r < n; r++)
                if ((bits[r >> 5] >> (r & 31)) & 1u) { mapped += 1; mean += ((double)lens[r] - mean) / (double)mapped; }
            h->rl_mean = mean; h->rl_mapped = mapped;
        }
    }
    WqWork wk{h->d_seeds.p, h->d_hit_read.p, h->d_hit_s.p, h->d_hit_s2.p, h->d_hit_gene.p, h->d_hits.p, h->d_pool.p,
              h->d_cursors.p, h->d_pending.p, h->d_pending2.p, hit_cap, (uint32_t)std::min<uint64_t>(h->d_pool.cap, 0xFFFFFFF0ull), h->d_defer.p, h->d_mapped_bits.p};
    const int T = 128;
    // keyed inserts that raced with the slot owner: retry until none is left
    uint32_t pending = h->h_cursors[2];
    int guard = 0;
    while (pending) {
        if (++guard > 64) return fail(-6, "keyed store retry did not converge");
        CK(cudaMemcpyAsync(h->d_pending2.p, h->d_pending.p, pending * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        const uint32_t* list = h->d_pending2.p;
        CK(cudaMemsetAsync(h->d_cursors.p + 2, 0, sizeof(uint32_t), s));
        if (!pe) wq_k_retry<<<(pending + T - 1) / T, T, 0, s>>>(h->dix, wk, h->q, list, pending);
        else wq_k_retry_pe<<<(pending + T - 1) / T, T, 0, s>>>(h->dix, wk, h->q, list, pending);
        h->launches += 1;
        CK(cudaMemcpyAsync(h->h_cursors + 8, h->d_cursors.p, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        pending = h->h_cursors[10];
    }
    if (h->bias_on) {
        // biasval = count!(mod, read.sequence) + push! into the read's counter (src/reads.jl:59-66, 108-118), after every insertion of the
        // chunk is in: the kernel looks each mapped read's counter up and appends its (counter, hexamer) / (counter, GC bin) records
        const uint64_t need = h->bias_rec_used + (uint64_t)n * WQ_BIAS_MAX_REC;
        if (need > h->bias_rec_cap) {                          // grow, keeping the records of earlier chunks
            const uint64_t want = need + need / 2;
            uint64_t* nb = nullptr;
            CK(cudaMalloc(&nb, want * 8));
            if (h->bias_rec_used) CK(cudaMemcpyAsync(nb, h->d_bias_rec, h->bias_rec_used * 8, cudaMemcpyDeviceToDevice, s));
            CK(cudaStreamSynchronize(s));
            if (h->d_bias_rec) cudaFree(h->d_bias_rec);
            h->d_bias_rec = nb; h->bias_rec_cap = want;
        }
        WqBiasDev B;
        B.fore = h->d_bias_tab.p; B.back = B.fore + WQ_BIAS_NHEX; B.cursor = B.back + WQ_BIAS_NHEX; B.rec = h->d_bias_rec; B.cap = h->bias_rec_cap;
        memcpy(B.gcbin, h->bias_gcbin, sizeof B.gcbin);
        const unsigned gread = (n + T - 1) / T;
        if (!pe) wq_k_bias<<<gread, T, 0, s>>>(h->dix, b, wk, h->q, B);
        else wq_k_bias_pe<<<gread, T, 0, s>>>(h->dix, b, *br, wk, h->q, B);
        h->launches += 1;
        CK(cudaGetLastError());
        uint64_t cur[2];
        CK(cudaMemcpyAsync(cur, B.cursor, 16, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        h->bias_rec_used = cur[0];
        if (cur[1]) return fail(-17, "bias correction: a mapped read is shorter than 37 bases (the reference throws BoundsError in primer_count!, src/bias.jl:163-165)");
    }
    return 0;
}

// bytes of one read's row in wq_read_batch.qual of a fixed_len batch (phred bytes, or 4-bit dictionary indices)
static inline uint64_t batch_qrow(const wq_read_batch* r) {
    return r->qual_bits == 4 ? ((((uint64_t)r->fixed_len + 1) / 2 + 3) & ~3ull) : (((uint64_t)r->fixed_len + 3) & ~3ull);
}
static int check_batch(const wq_read_batch* r) {
    if (!r) return fail(-1, "null read batch");
    if (r->n_reads && (!r->len || !r->seq2 || !r->qual)) return fail(-1, "read batch has null arrays");
    if (r->n_reads && r->fixed_len == 0 && (!r->seq_off || !r->qual_off)) return fail(-1, "read batch has null offset arrays (allowed only with fixed_len)");
    if (r->fixed_len > 255) return fail(-1, "wq_read_batch.fixed_len must be at most 255 (ReadLengthInt = UInt8)");
    if (r->qual_bits != 0 && r->qual_bits != 8 && r->qual_bits != 4) return fail(-1, "wq_read_batch.qual_bits must be 0, 8 or 4");
    if (r->qual_bits == 4 && r->fixed_len == 0) return fail(-1, "4-bit quality indices need a fixed_len batch");
    if (r->n_reads && r->fixed_len) {
        const uint64_t ws = (r->fixed_len + 31) / 32, qs = batch_qrow(r);
        if ((uint64_t)r->n_reads * ws > r->seq2_words || (uint64_t)r->n_reads * qs > r->qual_bytes) return fail(-1, "fixed_len batch: seq2_words / qual_bytes smaller than n_reads strides");
    }
    return 0;
}

// wq_align_out whose arrays the library owns and may grow (streaming SAM output of wq_process_fastq_sam)
struct WqOwnedOut {
    std::vector<uint32_t> hit_start; std::vector<uint8_t> nhits, flipped;
    std::vector<wq_alignment> aln, aln_mate; std::vector<wq_align_node> nodes;
    wq_align_out v;
    void bind(bool pe) {
        v.hit_start = hit_start.data(); v.nhits = nhits.data(); v.flipped = flipped.data();
        v.aln = aln.data(); v.aln_mate = pe ? aln_mate.data() : nullptr; v.nodes = nodes.data();
        v.aln_cap = aln.size(); v.nodes_cap = nodes.size();
    }
    void start(uint32_t n, bool pe) {
        hit_start.resize(n); nhits.resize(n); flipped.resize(n);
        if (aln.size() < (size_t)n * 2) aln.resize((size_t)n * 2);
        if (pe && aln_mate.size() < aln.size()) aln_mate.resize(aln.size());
        if (nodes.size() < (size_t)n * 4) nodes.resize((size_t)n * 4);
        bind(pe); v.aln_used = 0; v.nodes_used = 0;
    }
};
static thread_local WqOwnedOut* g_owned_out = nullptr;     // set by wq_process_fastq_sam around its wq_align_* calls

// Download the alignments of the chunk just processed (reads [first, first+n) of the caller's batch).
static int collect_chunk(wq_handle* h, uint32_t first, uint32_t n, bool pe, wq_align_out* out) {
    cudaStream_t s = h->stream;
    uint32_t nh = h->h_cursors[0], np = h->h_cursors[1];
    const uint32_t nback = pe ? 0 : h->h_cursors[5];          // single-end: reverse-strand hits sit at the back of the array
    const uint32_t hit_cap = h->last_hit_cap;
    if (nback) nh = hit_cap;                                     // index space of hit_base
    const uint32_t mult = pe ? 2 : 1;
    std::vector<WqSeed> seeds(n);
    std::vector<WqHit> hits((size_t)nh * mult);
    std::vector<wq_align_node> pool(np);
    CK(cudaMemcpyAsync(seeds.data(), h->d_seeds.p, n * sizeof(WqSeed), cudaMemcpyDeviceToHost, s));
    if (nback) {
        const uint32_t nfront = h->h_cursors[0];
        if (nfront) CK(cudaMemcpyAsync(hits.data(), h->d_hits.p, nfront * sizeof(WqHit), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(hits.data() + (hit_cap - nback), h->d_hits.p + (hit_cap - nback), nback * sizeof(WqHit), cudaMemcpyDeviceToHost, s));
    } else if (nh) CK(cudaMemcpyAsync(hits.data(), h->d_hits.p, hits.size() * sizeof(WqHit), cudaMemcpyDeviceToHost, s));
    if (np) CK(cudaMemcpyAsync(pool.data(), h->d_pool.p, np * sizeof(wq_align_node), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (pe && !out->aln_mate) return fail(-1, "wq_align_out.aln_mate is required for paired batches");
    if (g_owned_out && out == &g_owned_out->v) {               // library-owned output: make room for this chunk's worst case
        WqOwnedOut& O = *g_owned_out;
        const size_t need_nodes = (size_t)out->nodes_used + (size_t)np;       // every path of the chunk lies in the pool
        size_t need_aln = (size_t)out->aln_used;
        for (uint32_t i = 0; i < n; i++) need_aln += seeds[i].cnt;
        if (need_aln > O.aln.size()) O.aln.resize(need_aln + need_aln / 2);
        if (pe && O.aln_mate.size() < O.aln.size()) O.aln_mate.resize(O.aln.size());
        if (need_nodes > O.nodes.size()) O.nodes.resize(need_nodes + need_nodes / 2);
        O.bind(pe);
    }
    auto put = [&](const WqHit& hh, wq_alignment* dst) -> bool {
        if (out->nodes_used + hh.npath > out->nodes_cap) return false;
        wq_alignment a;
        a.offset = hh.offset; a.path_start = (uint32_t)out->nodes_used; a.offsetread = hh.offsetread;
        a.strand = hh.strand; a.isvalid = hh.flags & 1; a.npath = hh.npath;
        *dst = a;
        memcpy(out->nodes + out->nodes_used, pool.data() + hh.path_start, hh.npath * sizeof(wq_align_node));
        out->nodes_used += hh.npath;
        return true;
    };
    for (uint32_t i = 0; i < n; i++) {
        const WqSeed& sd = seeds[i];
        uint32_t r = first + i;
        out->hit_start[r] = (uint32_t)out->aln_used;
        uint8_t c = 0;
        for (int k = 0; k < sd.cnt; k++) {
            const WqHit& hh = hits[(size_t)(sd.hit_base + k) * mult];
            if (!(hh.flags & 4)) continue;
            if (out->aln_used >= out->aln_cap) return fail(-7, "wq_align_out capacity too small");
            if (!put(hh, &out->aln[out->aln_used])) return fail(-7, "wq_align_out capacity too small");
            if (pe && !put(hits[(size_t)(sd.hit_base + k) * 2 + 1], &out->aln_mate[out->aln_used])) return fail(-7, "wq_align_out capacity too small");
            out->aln_used++;
            c++;
        }
        out->nhits[r] = c;
        if (!pe) out->flipped[r] = (sd.cnt > 0 && !sd.strand) ? 1 : 0;
        else out->flipped[r] = (uint8_t)((!sd.strand ? 1 : 0) | ((sd.sp & 0x100u) ? 2 : 0));
    }
    return 0;
}

struct BatchOnDev { DevBuf<uint8_t>* len; DevBuf<uint64_t>* seqoff; DevBuf<uint64_t>* seq2; DevBuf<uint64_t>* qualoff; DevBuf<uint8_t>* qual; };

// position of read i of a batch (explicit offsets, or the strides of a fixed_len batch)
static inline uint64_t batch_soff(const wq_read_batch* r, uint64_t i) { return r->fixed_len ? i * ((r->fixed_len + 31) / 32) : r->seq_off[i]; }
static inline uint64_t batch_qoff(const wq_read_batch* r, uint64_t i) { return r->fixed_len ? i * batch_qrow(r) : r->qual_off[i]; }      // in the caller's `qual`
struct WqQualDict { uint8_t v[16]; };
// 4-bit quality indices of reads [lo, hi) -> phred bytes in the layout the kernels read (rows of (L + 3) & ~3 bytes, zero beyond the read)
__global__ void wq_k_unpack_qual(const uint8_t* __restrict__ packed, uint8_t* __restrict__ out, const uint8_t* __restrict__ len, uint32_t lo, uint32_t hi,
                                 uint32_t prow, uint32_t qrow, const WqQualDict d) {
    const uint32_t groups = qrow / 4;
    const uint64_t total = (uint64_t)(hi - lo) * groups;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t i = lo + (uint32_t)(t / groups), g = (uint32_t)(t % groups);
        const uint32_t L = len[i];
        const uint8_t* src = packed + (uint64_t)i * prow + 2 * g;
        uint32_t w = 0;
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t j = 4 * g + k;
            if (j < L) { const uint8_t b = src[k >> 1]; w |= (uint32_t)d.v[(k & 1) ? (b >> 4) : (b & 15)] << (8 * k); }
        }
        *reinterpret_cast<uint32_t*>(out + (uint64_t)i * qrow + 4 * g) = w;
    }
}
// quality bytes of reads [lo, hi) of batch `which` onto the device: a plain copy, or (4-bit indices) the packed rows + the expansion kernel
static int put_quals(wq_handle* h, const wq_read_batch* r, int which, uint32_t lo, uint32_t hi, uint64_t q0, uint64_t q1, cudaStream_t s) {
    if (r->qual_bits != 4) { CK(cudaMemcpyAsync(h->d_qual[which].p + q0, r->qual + q0, q1 - q0, cudaMemcpyHostToDevice, s)); return 0; }
    CK(cudaMemcpyAsync(h->d_qualpk[which].p + q0, r->qual + q0, q1 - q0, cudaMemcpyHostToDevice, s));
    WqQualDict d;
    memcpy(d.v, r->qual_dict, 16);
    wq_k_unpack_qual<<<592, 256, 0, s>>>(h->d_qualpk[which].p, h->d_qual[which].p, h->d_len[which].p, lo, hi, (uint32_t)batch_qrow(r), (r->fixed_len + 3) & ~3u, d);
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}
static int ensure_quals(wq_handle* h, const wq_read_batch* r, int which) {
    if (r->qual_bits == 4) {
        CK(h->d_qualpk[which].ensure(r->qual_bytes + 16));
        CK(h->d_qual[which].ensure((uint64_t)r->n_reads * ((r->fixed_len + 3) & ~3u) + 16));
    } else CK(h->d_qual[which].ensure(r->qual_bytes + 16));
    return 0;
}
// offsets of reads [lo, hi) of a fixed_len batch, written on the device instead of copied from the host
__global__ void wq_k_fill_offsets(uint64_t* seq_off, uint64_t* qual_off, uint32_t lo, uint32_t hi, uint32_t wstride, uint32_t qstride) {
    for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * blockDim.x) {
        seq_off[i] = i * wstride;
        qual_off[i] = i * qstride;
    }
}
static int put_offsets(wq_handle* h, const wq_read_batch* r, int which, uint32_t lo, uint32_t hi, cudaStream_t s) {
    if (r->fixed_len) {
        wq_k_fill_offsets<<<296, 256, 0, s>>>(h->d_seqoff[which].p, h->d_qualoff[which].p, lo, hi, (r->fixed_len + 31) / 32, (r->fixed_len + 3) & ~3u);
        h->launches += 1;
        CK(cudaGetLastError());
    } else {
        CK(cudaMemcpyAsync(h->d_seqoff[which].p + lo, r->seq_off + lo, (hi - lo) * 8ull, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(h->d_qualoff[which].p + lo, r->qual_off + lo, (hi - lo) * 8ull, cudaMemcpyHostToDevice, s));
    }
    return 0;
}

static int upload_batch(wq_handle* h, const wq_read_batch* reads, int which, cudaStream_t s) {
    const uint32_t n = reads->n_reads;
    CK(h->d_len[which].ensure(n)); CK(h->d_seqoff[which].ensure(n)); CK(h->d_qualoff[which].ensure(n));
    CK(h->d_seq2[which].ensure(reads->seq2_words + 2));
    if (int rc = ensure_quals(h, reads, which)) return rc;
    CK(cudaMemcpyAsync(h->d_len[which].p, reads->len, n, cudaMemcpyHostToDevice, s));
    if (int rc = put_offsets(h, reads, which, 0, n, s)) return rc;
    CK(cudaMemcpyAsync(h->d_seq2[which].p, reads->seq2, reads->seq2_words * 8ull, cudaMemcpyHostToDevice, s));
    if (int rc = put_quals(h, reads, which, 0, n, 0, reads->qual_bytes, s)) return rc;
    return 0;
}

static int align_impl(wq_handle* h, const wq_align_param* ap, const wq_read_batch* reads, const wq_read_batch* mates,
                      wq_align_out* out, bool on_device) {
    if (!h || !ap) return fail(-1, "null argument");
    if (int rc = check_batch(reads)) return rc;
    const bool pe = mates != nullptr;
    if (pe) {
        if (int rc = check_batch(mates)) return rc;
        if (mates->n_reads != reads->n_reads) return fail(-1, "mate batches differ in length");
    }
    CK(cudaSetDevice(h->device));
    WqParams p;
    std::string err = wq_make_params(ap, h->H.K, p);
    if (!err.empty()) return fail(-4, err);
    if (out) { out->aln_used = 0; out->nodes_used = 0; }
    const uint32_t n = reads->n_reads;
    if (n == 0) return 0;
    h->hq_valid = false;
    reset_quant_ex(h);
    for (auto& m : h->last_kernel_ms) m = 0;
    h->last_deferred = 0;
    const wq_read_batch* src[2] = {reads, mates};
    const int nb = pe ? 2 : 1;
    WqBatchDev base[2];
    const uint32_t nchunks = (n + h->chunk_reads - 1) / h->chunk_reads;
    std::vector<cudaEvent_t> copied;
    if (on_device) {
        for (int m = 0; m < nb; m++) if (src[m]->qual_bits == 4) return fail(-1, "device-resident batches take phred bytes (qual_bits 4 is a transfer format for host batches)");
        for (int m = 0; m < nb; m++) {
            if (src[m]->fixed_len && (!src[m]->seq_off || !src[m]->qual_off)) {       // device-resident fixed_len batch without offsets: derive them
                CK(h->d_seqoff[m].ensure(n)); CK(h->d_qualoff[m].ensure(n));
                if (int rc = put_offsets(h, src[m], m, 0, n, h->stream)) return rc;
                base[m] = WqBatchDev{n, src[m]->len, h->d_seqoff[m].p, src[m]->seq2, h->d_qualoff[m].p, src[m]->qual};
            } else base[m] = WqBatchDev{n, src[m]->len, src[m]->seq_off, src[m]->seq2, src[m]->qual_off, src[m]->qual};
        }
    } else {
        // Host buffers.  The packed arrays of chunk c+1 cross PCIe on the copy stream while chunk c is being
        // aligned: every chunk's slices are enqueued up front (async when the caller's buffers are page-locked,
        // e.g. from wq_pinned_alloc) and the compute stream waits on one event per chunk.  Needs ascending
        // offsets (what any packer produces); otherwise the whole batch is copied first.
        bool ascending = true;
        for (int m = 0; m < nb && ascending; m++)
            for (uint32_t lo = 0; lo < n; lo += h->chunk_reads) {
                uint32_t hi = std::min(n, lo + h->chunk_reads);
                uint64_t s0 = batch_soff(src[m], lo), s1 = hi < n ? batch_soff(src[m], hi) : src[m]->seq2_words;
                uint64_t q0 = batch_qoff(src[m], lo), q1 = hi < n ? batch_qoff(src[m], hi) : src[m]->qual_bytes;
                if (s1 < s0 || q1 < q0 || s1 > src[m]->seq2_words || q1 > src[m]->qual_bytes) { ascending = false; break; }
            }
        for (int m = 0; m < nb; m++) {
            CK(h->d_len[m].ensure(n)); CK(h->d_seqoff[m].ensure(n)); CK(h->d_qualoff[m].ensure(n));
            CK(h->d_seq2[m].ensure(src[m]->seq2_words + 2));
            if (int rc = ensure_quals(h, src[m], m)) return rc;
            base[m] = WqBatchDev{n, h->d_len[m].p, h->d_seqoff[m].p, h->d_seq2[m].p, h->d_qualoff[m].p, h->d_qual[m].p};
        }
        if (!ascending || nchunks == 1) {
            for (int m = 0; m < nb; m++) if (int rc = upload_batch(h, src[m], m, h->stream)) return rc;
        } else {
            cudaStream_t cs = h->copy_stream;
            CK(cudaStreamSynchronize(h->stream));       // buffers may still be read by a previous call's kernels
            copied.resize(nchunks);
            for (uint32_t c = 0; c < nchunks; c++) {
                uint32_t lo = c * h->chunk_reads, hi = std::min(n, lo + h->chunk_reads);
                for (int m = 0; m < nb; m++) {
                    const wq_read_batch* r = src[m];
                    uint64_t s0 = batch_soff(r, lo), s1 = hi < n ? batch_soff(r, hi) : r->seq2_words;
                    uint64_t q0 = batch_qoff(r, lo), q1 = hi < n ? batch_qoff(r, hi) : r->qual_bytes;
                    CK(cudaMemcpyAsync(h->d_len[m].p + lo, r->len + lo, hi - lo, cudaMemcpyHostToDevice, cs));
                    if (int rc = put_offsets(h, r, m, lo, hi, cs)) return rc;
                    CK(cudaMemcpyAsync(h->d_seq2[m].p + s0, r->seq2 + s0, (s1 - s0) * 8ull, cudaMemcpyHostToDevice, cs));
                    if (int rc = put_quals(h, r, m, lo, hi, q0, q1, cs)) return rc;
                }
                CK(cudaEventCreateWithFlags(&copied[c], cudaEventDisableTiming));
                CK(cudaEventRecord(copied[c], cs));
            }
        }
    }
    int rc_all = 0;
    for (uint32_t c = 0; c < nchunks && !rc_all; c++) {
        uint32_t lo = c * h->chunk_reads;
        uint32_t m = std::min(h->chunk_reads, n - lo);
        WqBatchDev b[2];
        for (int k = 0; k < nb; k++)
            b[k] = WqBatchDev{m, base[k].len + lo, base[k].seq_off + lo, base[k].seq2, base[k].qual_off + lo, base[k].qual};
        if (!copied.empty()) { cudaError_t e = cudaStreamWaitEvent(h->stream, copied[c], 0); if (e != cudaSuccess) { rc_all = fail(-10, cudaGetErrorString(e)); break; } }
        rc_all = run_chunk(h, p, b[0], pe ? &b[1] : nullptr, on_device);
        if (!rc_all && out) rc_all = collect_chunk(h, lo, m, pe, out);
    }
    if (!copied.empty()) {
        cudaStreamSynchronize(h->copy_stream);
        for (auto& e : copied) if (e) cudaEventDestroy(e);
    }
    return rc_all;
}

int wq_align_se(wq_handle* h, const wq_align_param* p, const wq_read_batch* reads, wq_align_out* out) {
    return align_impl(h, p, reads, nullptr, out, false);
}
int wq_align_se_device(wq_handle* h, const wq_align_param* p, const wq_read_batch* reads) {
    return align_impl(h, p, reads, nullptr, nullptr, true);
}
int wq_align_pe(wq_handle* h, const wq_align_param* p, const wq_read_batch* fwd, const wq_read_batch* rev, wq_align_out* out) {
    if (!rev) return fail(-1, "null mate batch");
    return align_impl(h, p, fwd, rev, out, false);
}
int wq_align_pe_device(wq_handle* h, const wq_align_param* p, const wq_read_batch* fwd, const wq_read_batch* rev) {
    if (!rev) return fail(-1, "null mate batch");
    return align_impl(h, p, fwd, rev, nullptr, true);
}

// ---- FASTQ ingest (src/reads.jl:2-23, src/record.jl:13-19) -----------------------------------------------
int wq_fastq_open(const char* path, int qualoffset, wq_fastq** out) {
    if (!path || !out) return fail(-1, "null argument");
    *out = nullptr;
    wq_fastq* f = new wq_fastq();
    f->qualoffset = qualoffset;
    if (int rc = wq_fastq_open_impl(f, path)) { std::string e = f->err; delete f; return fail(rc, e); }
    *out = f;
    return 0;
}
int wq_fastq_from_memory(const void* text, uint64_t len, int qualoffset, wq_fastq** out) {
    if ((!text && len) || !out) return fail(-1, "null argument");
    wq_fastq* f = new wq_fastq();
    f->map = (const char*)text; f->map_len = len; f->map_owned = false; f->qualoffset = qualoffset;
    f->wbase = f->map; f->woff = 0; f->wlen = len; f->eof = true;
    *out = f;
    return 0;
}
int wq_fastq_next(wq_fastq* f, uint32_t max_reads, wq_read_batch* out) {
    if (!f || !out) return fail(-1, "null argument");
    if (max_reads == 0) return fail(-1, "max_reads must be positive");
    int rc = wq_fastq_next_impl(f, max_reads, out);
    if (rc) return fail(rc, f->err);
    return 0;
}
// identifiers (the text after '@' up to the end of the line) of the batch returned by the last wq_fastq_next
int wq_fastq_names(wq_fastq* f, const uint64_t** name_off, const uint32_t** name_len, const char** names) {
    if (!f || !name_off || !name_len || !names) return fail(-1, "null argument");
    const WqFastqSlot& S = f->slot[f->cur];
    *name_off = S.name_off.data(); *name_len = S.name_len.data(); *names = S.names.data();
    return 0;
}
void wq_fastq_close(wq_fastq* f) {
    if (!f) return;
    wq_fastq_release(f);
    delete f;
}

// process_reads! / process_paired_reads! (src/reads.jl:41-129) over FASTQ files: the next batch is inflated, split and
// packed on the host threads while the device aligns and counts the current one.  With sam_fd >= 0 the SAM header and the
// records of every batch are written to that descriptor in read order (sam = true: src/reads.jl:49-52, 63-67, 110-121).
static int wq_sam_format(wq_handle* h, const wq_read_batch* fwd, const wq_read_batch* rev,
                         const uint64_t* name_off, const uint32_t* name_len, const char* names,
                         const uint64_t* rname_off, const uint32_t* rname_len, const char* rnames,
                         const wq_align_out* out, int qualoffset, std::vector<std::string>& parts);
static int wq_write_all(int fd, const char* p, size_t n) {
    while (n) {
        const ssize_t w = write(fd, p, n);
        if (w < 0) { if (errno == EINTR) continue; return -1; }
        p += w; n -= (size_t)w;
    }
    return 0;
}
int wq_process_fastq_sam(wq_handle* h, const wq_align_param* p, const char* fwd_path, const char* rev_path, int qualoffset,
                         uint32_t batch_reads, int sam_fd, wq_map_stats* st) {
    if (!h || !p || !fwd_path) return fail(-1, "null argument");
    const bool sam = sam_fd >= 0;
    if (sam && !h->sam.has_info) return fail(-15, "SAM output needs wq_index_set_gene_info (lib.info names and strands)");
    if (batch_reads == 0) batch_reads = 1u << 19;      // small enough that parsing batch k+1 hides behind aligning batch k from the first batch on
    wq_fastq* rd[2] = {nullptr, nullptr};
    if (int rc = wq_fastq_open(fwd_path, qualoffset, &rd[0])) return rc;
    if (rev_path) { if (int rc = wq_fastq_open(rev_path, qualoffset, &rd[1])) { wq_fastq_close(rd[0]); return rc; } }
    struct Pair { wq_read_batch b[2]; const WqFastqSlot* names[2] = {nullptr, nullptr}; int rc = 0; std::string err; };
    (void)wq_fastq_code_table();                               // initialised before two readers run side by side
    auto parse = [&]() {
        Pair P;
        if (!rev_path) {
            P.rc = wq_fastq_next_impl(rd[0], batch_reads, &P.b[0]);
            if (P.rc) P.err = rd[0]->err;
        } else {
            // the two files side by side: an ordinary .fastq.gz is inflated serially, so the mates' files one after the other would halve
            // the rate of the usual paired input (readers share nothing but the page-locked block cache, which has its own lock)
            int rc1 = 0;
            std::thread mate([&]() { rc1 = wq_fastq_next_impl(rd[1], batch_reads, &P.b[1]); });
            const int rc0 = wq_fastq_next_impl(rd[0], batch_reads, &P.b[0]);
            mate.join();
            if (rc0) { P.rc = rc0; P.err = rd[0]->err; }
            else if (rc1) { P.rc = rc1; P.err = rd[1]->err; }
        }
        for (int k = 0; k < (rev_path ? 2 : 1); k++) P.names[k] = &rd[k]->slot[rd[k]->cur];      // these slots are not touched again until the parse after the next

        if (!P.rc && rev_path && P.b[0].n_reads != P.b[1].n_reads) { P.rc = -21; P.err = "paired FASTQ files hold different numbers of records"; }
        return P;
    };
    int rc = 0;
    std::string err;
    if (sam) {
        const std::string hd = wq_sam_header_text(h->sam);
        if (wq_write_all(sam_fd, hd.data(), hd.size())) { wq_fastq_close(rd[0]); wq_fastq_close(rd[1]); return fail(-22, "cannot write SAM output"); }
    }
    WqOwnedOut own;
    std::vector<std::string> parts;
    Pair cur = parse();
    while (!cur.rc && cur.b[0].n_reads) {
        std::future<Pair> nxt = std::async(std::launch::async, parse);
        wq_align_out* out = nullptr;
        if (sam) { own.start(cur.b[0].n_reads, rev_path != nullptr); out = &own.v; g_owned_out = &own; }
        rc = rev_path ? wq_align_pe(h, p, &cur.b[0], &cur.b[1], out) : wq_align_se(h, p, &cur.b[0], out);
        g_owned_out = nullptr;
        if (rc) err = g_err;
        if (!rc && sam) {
            const WqFastqSlot* N0 = cur.names[0]; const WqFastqSlot* N1 = cur.names[1];
            rc = wq_sam_format(h, &cur.b[0], rev_path ? &cur.b[1] : nullptr, N0->name_off.data(), N0->name_len.data(), N0->names.data(),
                               N1 ? N1->name_off.data() : nullptr, N1 ? N1->name_len.data() : nullptr, N1 ? N1->names.data() : nullptr,
                               &own.v, qualoffset, parts);
            if (rc) err = g_err;
            for (size_t t = 0; t < parts.size() && !rc; t++)
                if (wq_write_all(sam_fd, parts[t].data(), parts[t].size())) { rc = -22; err = "cannot write SAM output"; }
        }
        cur = nxt.get();
        if (rc) break;
    }
    if (!rc && cur.rc) { rc = cur.rc; err = cur.err; }
    wq_fastq_close(rd[0]); wq_fastq_close(rd[1]);
    if (rc) return fail(rc, err);
    return st ? wq_map_stats_get(h, st) : 0;
}
int wq_process_fastq(wq_handle* h, const wq_align_param* p, const char* fwd_path, const char* rev_path, int qualoffset,
                     uint32_t batch_reads, wq_map_stats* st) {
    return wq_process_fastq_sam(h, p, fwd_path, rev_path, qualoffset, batch_reads, -1, st);
}

// ---- SAM output (src/io.jl:69-276) ----------------------------------------------------------------------
// lib.info (src/index.jl:8, src/types.jl:33-37): reference sequence name and strand of every gene
int wq_index_set_gene_info(wq_handle* h, const char* const* seqname, const uint8_t* strand) {
    if (!h || !seqname || !strand) return fail(-1, "null argument");
    h->sam.seqname.clear(); h->sam.strand.clear();
    for (uint32_t g = 0; g < h->H.G; g++) {
        if (!seqname[g]) return fail(-1, "null sequence name");
        h->sam.seqname.emplace_back(seqname[g]);
        h->sam.strand.push_back(strand[g] ? 1 : 0);
    }
    h->sam.has_info = true;
    return 0;
}

int wq_sam_header(wq_handle* h, char* buf, uint64_t cap, uint64_t* used) {
    if (!h || !used) return fail(-1, "null argument");
    if (!h->sam.has_info) return fail(-15, "SAM output needs wq_index_set_gene_info (lib.info names and strands)");
    const std::string t = wq_sam_header_text(h->sam);
    *used = t.size();
    if (!buf) return 0;
    if (cap < t.size()) return fail(-7, "SAM buffer too small");
    memcpy(buf, t.data(), t.size());
    return 0;
}

// write_sam for every read of an aligned batch, in read order (src/reads.jl:63-67, 110-121).  `out` is the wq_align_out the
// batch was aligned into; names are the FASTQ identifiers (wq_fastq_names layout).  Two-call protocol: buf == NULL
// returns the size (the text is kept for the second call when the arguments are the same batch).
int wq_sam_batch(wq_handle* h, const wq_read_batch* fwd, const wq_read_batch* rev,
                 const uint64_t* name_off, const uint32_t* name_len, const char* names,
                 const uint64_t* rname_off, const uint32_t* rname_len, const char* rnames,
                 const wq_align_out* out, int qualoffset, char* buf, uint64_t cap, uint64_t* used) {
    if (!h || !fwd || !out || !used || !name_off || !name_len || !names) return fail(-1, "null argument");
    if (rev && (!rname_off || !rname_len || !rnames || !out->aln_mate)) return fail(-1, "paired SAM output needs the mate's names and alignments");
    if (!h->sam.has_info) return fail(-15, "SAM output needs wq_index_set_gene_info (lib.info names and strands)");
    if (!out->hit_start || !out->nhits || !out->flipped || !out->aln || !out->nodes) return fail(-1, "wq_align_out has null arrays");
    std::vector<std::string> parts;
    if (int rc = wq_sam_format(h, fwd, rev, name_off, name_len, names, rname_off, rname_len, rnames, out, qualoffset, parts)) return rc;
    uint64_t total = 0;
    for (auto& s2 : parts) total += s2.size();
    *used = total;
    if (!buf) return 0;
    if (cap < total) return fail(-7, "SAM buffer too small");
    uint64_t o = 0;
    for (auto& s2 : parts) { memcpy(buf + o, s2.data(), s2.size()); o += s2.size(); }
    return 0;
}

// the reads of a batch split over the host threads; parts concatenate to the text in read order
static int wq_sam_format(wq_handle* h, const wq_read_batch* fwd, const wq_read_batch* rev,
                         const uint64_t* name_off, const uint32_t* name_len, const char* names,
                         const uint64_t* rname_off, const uint32_t* rname_len, const char* rnames,
                         const wq_align_out* out, int qualoffset, std::vector<std::string>& parts) {
    const uint32_t n = fwd->n_reads;
    unsigned nt = wq_host_threads();
    if (n < 4096) nt = 1;
    parts.assign(nt, std::string());
    auto work = [&](unsigned t) {
        wq_sam_range(h->sam, parts[t], fwd, rev, name_off, name_len, names, rname_off, rname_len, rnames, out, qualoffset,
                     (uint32_t)((uint64_t)n * t / nt), (uint32_t)((uint64_t)n * (t + 1) / nt));
    };
    if (nt == 1) work(0);
    else { std::vector<std::thread> th; for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t); for (auto& x : th) x.join(); }
    return 0;
}

// per-kernel device milliseconds of the last wq_align_*_device call: [seed, extend, resolve]
int wq_last_kernel_ms(wq_handle* h, float* ms3) {
    if (!h || !ms3) return fail(-1, "null argument");
    for (int i = 0; i < 3; i++) ms3[i] = h->last_kernel_ms[i];
    return 0;
}

// Run all work of this handle on a caller-owned stream (NULL restores the handle's own stream), so a
// host that times with its own events, or orders work against its own copies, sees the kernels.
int wq_set_stream(wq_handle* h, void* stream) {
    if (!h) return fail(-1, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    if (h->has_l2attr) { cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &h->l2attr); cudaGetLastError(); }
    return 0;
}

uint64_t wq_launch_count(wq_handle* h) { return h ? h->launches : 0; }
uint64_t wq_last_deferred(wq_handle* h) { return h ? h->last_deferred : 0; }

int wq_quant_kernel_stats(wq_handle* h, double* out8) {
    if (!h || !out8) return fail(-1, "null argument");
    const WqClassSet& C = h->hx.cls;
    out8[0] = h->hx.em_kernel_ms; out8[1] = h->psi_kernel_ms; out8[2] = (double)h->hx.iterations;
    out8[3] = (double)C.count.size(); out8[4] = (double)C.iso.size(); out8[5] = (double)h->iso.I;
    out8[6] = (double)(h->ev_batch[0].P.size() + h->ev_batch[1].P.size()); out8[7] = (double)h->psi_sweep_bytes;
    return 0;
}

int wq_map_stats_get(wq_handle* h, wq_map_stats* out) {
    if (!h || !out) return fail(-1, "null argument");
    CK(cudaSetDevice(h->device));
    WqCounters c;
    CK(cudaMemcpy(&c, h->q.counters, sizeof c, cudaMemcpyDeviceToHost));
    out->total = c.total; out->mapped = c.mapped; out->multi = c.multi; out->sum_readlen = c.sum_readlen;
    out->path_overflow = c.path_overflow;
    out->mean_readlen = h->rl_mean;
    if (c.edge_full) return fail(-8, "edge count table is full (raise WQ_EDGE_SLOTS)");
    if (c.keyed_full) return fail(-8, "keyed count store is full (raise WQ_KEYED_SLOTS / WQ_KEYED_POOL_WORDS)");
    if (c.pool_overflow) return fail(-8, "alignment path pool overflowed");
    return 0;
}

int wq_counts_device_ptr(wq_handle* h, void** dptr, uint64_t* n_u64) {
    if (!h || !dptr || !n_u64) return fail(-1, "null argument");
    *dptr = h->d_dense.p;
    *n_u64 = h->H.N + 4;     // node counts + total, mapped, multi, sum_readlen
    return 0;
}

// Pull the device count stores into the handle's host-side quant state: occupied table slots are
// compacted on the device, edges are radix-sorted there (canonical order), only the dense results
// cross PCIe.
static int sync_host_quant(wq_handle* h) {
    if (h->hq_valid) return 0;
    WqTimer tm_("sync_host_quant");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    WqHostQuant& Q = h->hq;
    Q = WqHostQuant();
    const uint32_t N = h->H.N;
    std::vector<uint64_t> dense(N + sizeof(WqCounters) / 8);
    CK(cudaMemcpyAsync(dense.data(), h->d_dense.p, dense.size() * 8, cudaMemcpyDeviceToHost, s));
    // cursors[8] = #edges, cursors[9] = #keyed entries
    CK(cudaMemsetAsync(h->d_cursors.p + 12, 0, 2 * sizeof(uint32_t), s));
    CK(h->d_ck.ensure(h->edge_slots)); CK(h->d_cv.ensure(h->edge_slots));
    DevBuf<uint32_t> eslot, kslot, eslot_sorted;                 // --biascorrect: table slots of the compacted records
    if (h->bias_on) { CK(eslot.ensure(h->edge_slots)); CK(kslot.ensure(h->keyed_slots)); }
    wq_k_compact_edges<<<1184, 256, 0, s>>>(h->d_ekey.p, h->d_ecount.p, h->edge_slots, h->d_ck.p, h->d_cv.p, h->d_cursors.p + 12, eslot.p);
    CK(h->d_ckoff.ensure(h->keyed_slots)); CK(h->d_cv2.ensure(h->keyed_slots));
    wq_k_compact_keyed<<<1184, 256, 0, s>>>(h->d_khash.p, h->d_koff.p, h->d_kcount.p, h->keyed_slots, h->d_ckoff.p, h->d_cv2.p, h->d_cursors.p + 13, kslot.p);
    h->launches += 2;
    uint64_t kcur = 0;
    CK(cudaMemcpyAsync(h->h_cursors + 12, h->d_cursors.p + 12, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&kcur, h->d_kcursor.p, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint32_t ne = h->h_cursors[12], nk = h->h_cursors[13];
    // --biascorrect: primer_normalize! / primer_adjust! / the factors of gc_adjust! for every counter, on the device (csrc/wq_bias.cuh)
    const uint64_t ES = h->edge_slots, KS = h->keyed_slots;
    double *d_node_p = nullptr, *d_node_g = nullptr, *d_edge_p = nullptr, *d_edge_g = nullptr, *d_keyed_p = nullptr, *d_keyed_g = nullptr;
    if (h->bias_on) {
        CK(h->d_bias_slot.ensure(2ull * N + 2 * ES + 2 * KS));
        d_node_p = h->d_bias_slot.p; d_node_g = d_node_p + N; d_edge_p = d_node_g + N; d_edge_g = d_edge_p + ES; d_keyed_p = d_edge_g + ES; d_keyed_g = d_keyed_p + KS;
        CK(cudaMemsetAsync(d_node_p, 0, N * 8ull, s)); CK(cudaMemsetAsync(d_edge_p, 0, ES * 8ull, s)); CK(cudaMemsetAsync(d_keyed_p, 0, KS * 8ull, s));
        wq_k_fill_f64<<<592, 256, 0, s>>>(d_node_g, -1.0, N);
        wq_k_fill_f64<<<592, 256, 0, s>>>(d_edge_g, -1.0, ES);
        wq_k_fill_f64<<<592, 256, 0, s>>>(d_keyed_g, -1.0, KS);
        CK(h->d_bias_model.ensure(3 * WQ_BIAS_NHEX + 32));
        double* ratio = h->d_bias_model.p; double* gcval = ratio + WQ_BIAS_NHEX; double* fore_n = gcval + 32; double* back_n = fore_n + WQ_BIAS_NHEX;
        wq_k_bias_model<<<1, 1024, 0, s>>>(h->d_bias_tab.p, h->d_bias_tab.p + WQ_BIAS_NHEX, ratio, gcval, fore_n, back_n);
        h->launches += 4;
        const uint64_t nrec = h->bias_rec_used;
        if (nrec >= (1ull << 31)) return fail(-18, "bias correction: more than 2^31 tally records in one job (about 250 M mapped reads); split the input");
        if (nrec) {
            DevBuf<uint64_t> sorted, ukey; DevBuf<uint32_t> ucnt, nuniq;
            CK(sorted.ensure(nrec)); CK(ukey.ensure(nrec)); CK(ucnt.ensure(nrec)); CK(nuniq.ensure(1));
            size_t t1 = 0, t2 = 0;
            CK(cub::DeviceRadixSort::SortKeys(nullptr, t1, h->d_bias_rec, sorted.p, (int)nrec, 0, 47, s));
            CK(cub::DeviceRunLengthEncode::Encode(nullptr, t2, sorted.p, ukey.p, ucnt.p, nuniq.p, (int)nrec, s));
            CK(h->d_cubtmp.ensure(std::max(t1, t2)));
            size_t tb = h->d_cubtmp.cap;
            CK(cub::DeviceRadixSort::SortKeys(h->d_cubtmp.p, tb, h->d_bias_rec, sorted.p, (int)nrec, 0, 47, s));
            tb = h->d_cubtmp.cap;
            CK(cub::DeviceRunLengthEncode::Encode(h->d_cubtmp.p, tb, sorted.p, ukey.p, ucnt.p, nuniq.p, (int)nrec, s));
            wq_k_bias_adjust<<<1184, 256, 0, s>>>(ukey.p, ucnt.p, nuniq.p, ratio, gcval, d_node_p, d_node_g, d_edge_p, d_edge_g, d_keyed_p, d_keyed_g);
            h->launches += 3;
            cudaError_t e1 = cudaGetLastError();
            cudaError_t e2 = cudaStreamSynchronize(s);
            sorted.release(); ukey.release(); ucnt.release(); nuniq.release();
            CK(e1); CK(e2);
        }
        Q.bias = true;
        Q.node_p.resize(N); Q.node_g.resize(N);
        CK(cudaMemcpyAsync(Q.node_p.data(), d_node_p, N * 8ull, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(Q.node_g.data(), d_node_g, N * 8ull, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    Q.node.assign(dense.begin(), dense.begin() + N);
    memcpy(&Q.stats, dense.data() + N, sizeof(WqCounters));
    if (Q.stats.keyed_full || kcur > h->kpool_cap) return fail(-8, "keyed count store overflowed (raise WQ_KEYED_SLOTS / WQ_KEYED_POOL_WORDS)");
    if (Q.stats.edge_full) return fail(-8, "edge count table is full (raise WQ_EDGE_SLOTS)");
    // edges: device radix sort by key
    Q.edge_key.resize(ne); Q.edge_count.resize(ne);
    if (ne) {
        CK(h->d_ck2.ensure(ne));
        uint64_t* vals_out = h->d_cv2.p + nk;          // tail of the keyed-count buffer is free
        if ((size_t)nk + ne > h->d_cv2.cap) { CK(h->d_cubtmp.ensure(1)); }
        DevBuf<uint64_t> tmpvals;
        if ((size_t)nk + ne > h->d_cv2.cap) { CK(tmpvals.ensure(ne)); vals_out = tmpvals.p; }
        size_t tmp_bytes = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->d_ck.p, h->d_ck2.p, h->d_cv.p, vals_out, (int)ne, 0, 64, s));
        CK(h->d_cubtmp.ensure(tmp_bytes));
        CK(cub::DeviceRadixSort::SortPairs(h->d_cubtmp.p, tmp_bytes, h->d_ck.p, h->d_ck2.p, h->d_cv.p, vals_out, (int)ne, 0, 64, s));
        h->launches += 1;
        CK(cudaMemcpyAsync(Q.edge_key.data(), h->d_ck2.p, ne * 8ull, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(Q.edge_count.data(), vals_out, ne * 8ull, cudaMemcpyDeviceToHost, s));
        if (h->bias_on) {                                   // the same sort carrying the slots, then the slots' results in key order
            DevBuf<uint64_t> keys2; DevBuf<double> pg;
            CK(eslot_sorted.ensure(ne)); CK(keys2.ensure(ne)); CK(pg.ensure(2ull * ne));
            size_t tb2 = 0;
            CK(cub::DeviceRadixSort::SortPairs(nullptr, tb2, h->d_ck.p, keys2.p, eslot.p, eslot_sorted.p, (int)ne, 0, 64, s));
            CK(h->d_cubtmp.ensure(tb2));
            CK(cub::DeviceRadixSort::SortPairs(h->d_cubtmp.p, tb2, h->d_ck.p, keys2.p, eslot.p, eslot_sorted.p, (int)ne, 0, 64, s));
            wq_k_bias_gather<<<592, 256, 0, s>>>(eslot_sorted.p, ne, d_edge_p, d_edge_g, pg.p, pg.p + ne);
            h->launches += 2;
            Q.edge_p.resize(ne); Q.edge_g.resize(ne);
            CK(cudaMemcpyAsync(Q.edge_p.data(), pg.p, ne * 8ull, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(Q.edge_g.data(), pg.p + ne, ne * 8ull, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            keys2.release(); pg.release();
        }
        CK(cudaStreamSynchronize(s));
        tmpvals.release();
    }
    // keyed entries + the used part of the key pool
    std::vector<uint32_t> koff(nk), kpool(kcur);
    std::vector<uint64_t> kcount(nk);
    std::vector<double> kpg;
    if (nk) {
        CK(cudaMemcpyAsync(koff.data(), h->d_ckoff.p, nk * 4ull, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(kcount.data(), h->d_cv2.p, nk * 8ull, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(kpool.data(), h->d_kpool.p, kcur * 4ull, cudaMemcpyDeviceToHost, s));
        if (h->bias_on) {
            DevBuf<double> pg;
            CK(pg.ensure(2ull * nk));
            wq_k_bias_gather<<<592, 256, 0, s>>>(kslot.p, nk, d_keyed_p, d_keyed_g, pg.p, pg.p + nk);
            h->launches += 1;
            kpg.resize(2ull * nk);
            CK(cudaMemcpyAsync(kpg.data(), pg.p, 2ull * nk * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            pg.release();
        }
        CK(cudaStreamSynchronize(s));
    }
    eslot.release(); kslot.release(); eslot_sorted.release();
    Q.longs.words.reserve(kcur);
    for (uint32_t i = 0; i < nk; i++) {
        if (koff[i] >= WQ_KEYOFF_FULL) return fail(-8, "keyed count store overflowed");
        const uint32_t nw = kpool[koff[i]];
        const uint32_t* w = &kpool[koff[i] + 1];
        WqKeyedStore& K = wq_key_is_multi(w) ? Q.multis : Q.longs;
        K.push(w, nw, kcount[i]);
        if (h->bias_on) { K.bias_p.push_back(kpg[i]); K.bias_g.push_back(kpg[nk + i]); }
    }
    Q.multis.canonicalize();
    h->hq_valid = true;
    return 0;
}

int wq_counts_download(wq_handle* h, wq_counts* c) {
    if (!h || !c) return fail(-1, "null argument");
    if (int rc = sync_host_quant(h)) return rc;
    WqHostQuant& Q = h->hq;
    uint64_t ne = 0, nc = 0;
    for (uint64_t k : Q.edge_key) { if (wq_edge_is_circ(k)) nc++; else ne++; }
    bool sizing = !c->node_count && !c->edge_key && !c->circ_key;
    if (!sizing && (c->n_nodes < Q.node.size() || c->n_edge < ne || c->n_circ < nc)) return fail(-7, "wq_counts buffers too small");
    c->n_nodes = Q.node.size(); c->n_edge = ne; c->n_circ = nc;
    if (sizing) return 0;
    if (c->node_count) memcpy(c->node_count, Q.node.data(), Q.node.size() * 8);
    uint64_t ie = 0, ic = 0;
    for (size_t i = 0; i < Q.edge_key.size(); i++) {
        if (wq_edge_is_circ(Q.edge_key[i])) { if (c->circ_key) { c->circ_key[ic] = Q.edge_key[i]; c->circ_count[ic] = Q.edge_count[i]; } ic++; }
        else { if (c->edge_key) { c->edge_key[ie] = Q.edge_key[i]; c->edge_count[ie] = Q.edge_count[i]; } ie++; }
    }
    return 0;
}

int wq_keyed_download(wq_handle* h, int kind, wq_keyed* k) {
    if (!h || !k) return fail(-1, "null argument");
    if (kind != 0 && kind != 1) return fail(-1, "kind must be 0 (long paths) or 1 (multi-map classes)");
    if (int rc = sync_host_quant(h)) return rc;
    if (kind == 0) h->hq.longs.canonicalize();        // sorted on demand; the quant stage does not need the order
    const WqKeyedStore& K = kind == 0 ? h->hq.longs : h->hq.multis;
    bool sizing = !k->rec_ptr && !k->words && !k->count;
    if (!sizing && (k->n_rec < K.size() || k->n_words < K.words.size())) return fail(-7, "wq_keyed buffers too small");
    k->n_rec = K.size(); k->n_words = K.words.size();
    if (sizing) return 0;
    memcpy(k->rec_ptr, K.off.data(), K.off.size() * 8);
    if (!K.words.empty()) memcpy(k->words, K.words.data(), K.words.size() * 4);
    if (K.size()) memcpy(k->count, K.count.data(), K.size() * 8);
    return 0;
}

// --biascorrect: the Float64 side of every counter, aligned with wq_counts_download / wq_keyed_download
int wq_bias_download(wq_handle* h, wq_bias_counts* o) {
    if (!h || !o) return fail(-1, "null argument");
    if (!h->bias_on) return fail(-16, "bias correction is off (wq_set_biascorrect)");
    if (int rc = sync_host_quant(h)) return rc;
    WqHostQuant& Q = h->hq;
    Q.longs.canonicalize();
    uint64_t ne = 0, nc = 0;
    for (uint64_t k : Q.edge_key) { if (wq_edge_is_circ(k)) nc++; else ne++; }
    const bool sizing = !o->node_primer && !o->edge_primer && !o->circ_primer && !o->long_primer && !o->multi_primer;
    if (!sizing && (o->n_nodes < Q.node_p.size() || o->n_edge < ne || o->n_circ < nc || o->n_long < Q.longs.size() || o->n_multi < Q.multis.size()))
        return fail(-7, "wq_bias_counts buffers too small");
    o->n_nodes = Q.node_p.size(); o->n_edge = ne; o->n_circ = nc; o->n_long = Q.longs.size(); o->n_multi = Q.multis.size();
    if (sizing) return 0;
    auto put = [](double* dst, const std::vector<double>& v) { if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * 8); };
    put(o->node_primer, Q.node_p); put(o->node_gc, Q.node_g);
    uint64_t ie = 0, ic = 0;
    for (size_t i = 0; i < Q.edge_key.size(); i++) {
        if (wq_edge_is_circ(Q.edge_key[i])) { if (o->circ_primer) o->circ_primer[ic] = Q.edge_p[i]; if (o->circ_gc) o->circ_gc[ic] = Q.edge_g[i]; ic++; }
        else { if (o->edge_primer) o->edge_primer[ie] = Q.edge_p[i]; if (o->edge_gc) o->edge_gc[ie] = Q.edge_g[i]; ie++; }
    }
    put(o->long_primer, Q.longs.bias_p); put(o->long_gc, Q.longs.bias_g);
    put(o->multi_primer, Q.multis.bias_p); put(o->multi_gc, Q.multis.bias_g);
    if (o->fore || o->back) {
        CK(cudaSetDevice(h->device));
        const double* fore_n = h->d_bias_model.p + WQ_BIAS_NHEX + 32;
        if (o->fore) CK(cudaMemcpyAsync(o->fore, fore_n, WQ_BIAS_NHEX * 8, cudaMemcpyDeviceToHost, h->stream));
        if (o->back) CK(cudaMemcpyAsync(o->back, fore_n + WQ_BIAS_NHEX, WQ_BIAS_NHEX * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

int wq_counts_merge(wq_handle* h, const wq_counts* c, const wq_keyed* longs, const wq_keyed* multis) {
    if (!h) return fail(-1, "null handle");
    if (int rc = sync_host_quant(h)) return rc;
    WqHostQuant& Q = h->hq;
    reset_quant_ex(h);
    if (c) {
        if (c->node_count) {
            if (c->n_nodes != Q.node.size()) return fail(-1, "node count vector has the wrong length");
            for (size_t i = 0; i < Q.node.size(); i++) Q.node[i] += c->node_count[i];
        }
        std::vector<std::pair<uint64_t, uint64_t>> all;
        all.reserve(Q.edge_key.size() + c->n_edge + c->n_circ);
        for (size_t i = 0; i < Q.edge_key.size(); i++) all.emplace_back(Q.edge_key[i], Q.edge_count[i]);
        for (uint64_t i = 0; i < c->n_edge; i++) all.emplace_back(c->edge_key[i], c->edge_count[i]);
        for (uint64_t i = 0; i < c->n_circ; i++) all.emplace_back(c->circ_key[i], c->circ_count[i]);
        std::sort(all.begin(), all.end());
        Q.edge_key.clear(); Q.edge_count.clear();
        for (auto& kv : all) {
            if (!Q.edge_key.empty() && Q.edge_key.back() == kv.first) Q.edge_count.back() += kv.second;
            else { Q.edge_key.push_back(kv.first); Q.edge_count.push_back(kv.second); }
        }
    }
    const wq_keyed* in[2] = {longs, multis};
    WqKeyedStore* dst[2] = {&Q.longs, &Q.multis};
    for (int t = 0; t < 2; t++) {
        if (!in[t] || !in[t]->n_rec) continue;
        for (uint64_t i = 0; i < in[t]->n_rec; i++)
            dst[t]->push(in[t]->words + in[t]->rec_ptr[i], (size_t)(in[t]->rec_ptr[i + 1] - in[t]->rec_ptr[i]), in[t]->count[i]);
        dst[t]->canonicalize();
    }
    return 0;
}

// Device-side exchange of the sparse stores: every rank compacts its edge / circ table and its keyed
// store (long paths, multi-map classes) into ONE device buffer, the buffers are all-gathered over NCCL by the caller, and
// every rank adds the other ranks' records into its own device tables with a kernel.  The host-side stores are then pulled
// once from the merged tables.  Buffer layout (u64 words): {n_edge, n_keyed, pool_words, 0}, edge_key[n_edge],
// edge_count[n_edge], keyed_count[n_keyed], then u32: keyed_off[n_keyed] (padded to even), pool[pool_words] (padded to even),
// where pool[off] = number of key words and the words follow.
int wq_sparse_device_pack(wq_handle* h, void** dptr, uint64_t* bytes) {
    if (!h || !dptr || !bytes) return fail(-1, "null argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    CK(cudaMemsetAsync(h->d_cursors.p + 12, 0, 2 * sizeof(uint32_t), s));
    CK(h->d_ck.ensure(h->edge_slots)); CK(h->d_cv.ensure(h->edge_slots));
    wq_k_compact_edges<<<1184, 256, 0, s>>>(h->d_ekey.p, h->d_ecount.p, h->edge_slots, h->d_ck.p, h->d_cv.p, h->d_cursors.p + 12);
    CK(h->d_ckoff.ensure(h->keyed_slots)); CK(h->d_cv2.ensure(h->keyed_slots));
    wq_k_compact_keyed<<<1184, 256, 0, s>>>(h->d_khash.p, h->d_koff.p, h->d_kcount.p, h->keyed_slots, h->d_ckoff.p, h->d_cv2.p, h->d_cursors.p + 13);
    h->launches += 2;
    uint64_t kcur = 0;
    CK(cudaMemcpyAsync(h->h_cursors + 12, h->d_cursors.p + 12, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&kcur, h->d_kcursor.p, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint64_t ne = h->h_cursors[12], nk = h->h_cursors[13];
    if (kcur > h->kpool_cap) return fail(-8, "keyed count store overflowed (raise WQ_KEYED_POOL_WORDS)");
    const uint64_t koff_words = (nk + 1) & ~1ull, pool_words = (kcur + 1) & ~1ull;
    const uint64_t total = 4 + 2 * ne + nk + koff_words / 2 + pool_words / 2;
    CK(h->d_xpack.ensure(total));
    const uint64_t hdr[4] = {ne, nk, kcur, 0};
    uint64_t* b = h->d_xpack.p;
    CK(cudaMemcpyAsync(b, hdr, sizeof hdr, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b + 4, h->d_ck.p, ne * 8, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(b + 4 + ne, h->d_cv.p, ne * 8, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(b + 4 + 2 * ne, h->d_cv2.p, nk * 8, cudaMemcpyDeviceToDevice, s));
    uint32_t* w = (uint32_t*)(b + 4 + 2 * ne + nk);
    CK(cudaMemcpyAsync(w, h->d_ckoff.p, nk * 4, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(w + koff_words, h->d_kpool.p, kcur * 4, cudaMemcpyDeviceToDevice, s));
    CK(cudaStreamSynchronize(s));
    *dptr = b;
    *bytes = total * 8;
    return 0;
}

int wq_sparse_device_merge(wq_handle* h, const void* dbuf, uint64_t bytes) {
    if (!h || !dbuf) return fail(-1, "null argument");
    if (bytes < 32) return fail(-1, "sparse device buffer too short");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    uint64_t hdr[4];
    CK(cudaMemcpyAsync(hdr, dbuf, sizeof hdr, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint64_t ne = hdr[0], nk = hdr[1], kw = hdr[2];
    const uint64_t need = (4 + 2 * ne + nk + ((nk + 1) & ~1ull) / 2 + ((kw + 1) & ~1ull) / 2) * 8;
    if (bytes < need) return fail(-1, "sparse device buffer truncated");
    CK(h->d_xpending.ensure(nk + 1));
    CK(cudaMemsetAsync(h->d_cursors.p + 14, 0, sizeof(uint32_t), s));
    const uint64_t n = ne + nk;
    if (n) {
        const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 16);
        wq_k_merge_sparse<<<grid, 256, 0, s>>>(h->q, (const uint64_t*)dbuf, h->d_xpending.p, h->d_cursors.p + 14, 0);
        wq_k_merge_sparse<<<148, 256, 0, s>>>(h->q, (const uint64_t*)dbuf, h->d_xpending.p, h->d_cursors.p + 14, 1);
        h->launches += 2;
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(s));
    h->hq = WqHostQuant();
    h->hq_valid = false;             // the host-side stores are pulled again from the merged tables
    reset_quant_ex(h);
    return 0;
}

// ---- multi-GPU exchange inside the library --------------------------------------------------------------------------------
// libnccl.so.2 is opened at run time: the few entry points used have had the same signatures since NCCL 2.0.
struct WqNcclId { char internal[128]; };
struct WqNccl {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /* ncclUniqueId by value: 128 bytes */ WqNcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static WqNccl g_nccl;
static std::string wq_nccl_load() {
    if (g_nccl.lib) return "";
    const char* names[] = {getenv("WQ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) if (n && *n && (lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!lib) return std::string("cannot open libnccl.so.2 (set WQ_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
    auto sym = [&](const char* n) { return dlsym(lib, n); };
    g_nccl.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, WqNcclId, int))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
    g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.AllGather) return "libnccl.so.2 lacks an expected entry point";
    g_nccl.lib = lib;
    return "";
}
#define NK(call)                                                                                                    \
    do {                                                                                                            \
        int r_ = (call);                                                                                            \
        if (r_ != 0) return fail(-30, std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "NCCL error")); \
    } while (0)
enum { WQ_NCCL_UINT8 = 1, WQ_NCCL_UINT64 = 5, WQ_NCCL_SUM = 0 };

int wq_comm_unique_id(void* id128) {
    if (!id128) return fail(-1, "null argument");
    std::string e = wq_nccl_load();
    if (!e.empty()) return fail(-30, e);
    NK(g_nccl.GetUniqueId(id128));
    return 0;
}
int wq_comm_init_rank(wq_handle* h, int nranks, int rank, const void* id128) {
    if (h && h->bias_on && nranks > 1) return fail(-16, "bias correction is single-GPU in this version: the per-counter tallies are keyed by table slot, which differs between ranks");
    if (!h || !id128) return fail(-1, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(-1, "rank out of range");
    std::string e = wq_nccl_load();
    if (!e.empty()) return fail(-30, e);
    CK(cudaSetDevice(h->device));
    wq_comm_destroy(h);
    WqNcclId id;
    memcpy(&id, id128, sizeof id);
    NK(g_nccl.CommInitRank(&h->nccl_comm, nranks, id, rank));
    h->comm_rank = rank; h->comm_size = nranks;
    if (!h->h_xhdr) CK(cudaHostAlloc((void**)&h->h_xhdr, 1024 * 4 * sizeof(uint64_t), cudaHostAllocDefault));
    return wq_set_gene_partition(h, (uint32_t)rank, (uint32_t)nranks);
}
int wq_comm_destroy(wq_handle* h) {
    if (!h || !h->nccl_comm) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (g_nccl.CommDestroy) g_nccl.CommDestroy(h->nccl_comm);
    h->nccl_comm = nullptr; h->comm_rank = 0; h->comm_size = 1;
    return 0;
}

int wq_counts_sync(wq_handle* h) {
    if (!h) return fail(-1, "null handle");
    if (!h->nccl_comm) return fail(-31, "wq_counts_sync needs wq_comm_init_rank first");
    if (h->comm_size > 1024) return fail(-31, "more than 1024 ranks");
    WqTimer tm_("counts_sync");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int W = h->comm_size;
    CK(h->d_ck.ensure(h->edge_slots)); CK(h->d_cv.ensure(h->edge_slots));
    CK(h->d_ckoff.ensure(h->keyed_slots)); CK(h->d_cv2.ensure(h->keyed_slots));
    for (int attempt = 0;; attempt++) {
        const uint64_t cap = h->xcap;
        CK(h->d_xpack.ensure(cap)); CK(h->d_xgather.ensure(cap * (uint64_t)W)); CK(h->d_xpending.ensure(cap));
        CK(cudaMemsetAsync(h->d_cursors.p + 12, 0, 2 * sizeof(uint32_t), s));
        wq_k_compact_edges<<<1184, 256, 0, s>>>(h->d_ekey.p, h->d_ecount.p, h->edge_slots, h->d_ck.p, h->d_cv.p, h->d_cursors.p + 12);
        wq_k_compact_keyed<<<1184, 256, 0, s>>>(h->d_khash.p, h->d_koff.p, h->d_kcount.p, h->keyed_slots, h->d_ckoff.p, h->d_cv2.p, h->d_cursors.p + 13);
        wq_k_pack_sparse<<<592, 256, 0, s>>>(h->d_xpack.p, cap, h->d_cursors.p + 12, h->d_kcursor.p, h->d_ck.p, h->d_cv.p, h->d_cv2.p, h->d_ckoff.p, h->d_kpool.p);
        h->launches += 3;
        CK(cudaGetLastError());
        NK(g_nccl.AllGather(h->d_xpack.p, h->d_xgather.p, cap, WQ_NCCL_UINT64, h->nccl_comm, s));
        for (int r = 0; r < W; r++) {
            if (r == h->comm_rank) continue;
            CK(cudaMemsetAsync(h->d_cursors.p + 14, 0, sizeof(uint32_t), s));
            wq_k_merge_gathered<<<592, 256, 0, s>>>(h->q, h->d_xgather.p, cap, W, r, h->d_xpending.p, h->d_cursors.p + 14, 0);
            wq_k_merge_gathered<<<148, 256, 0, s>>>(h->q, h->d_xgather.p, cap, W, r, h->d_xpending.p, h->d_cursors.p + 14, 1);
            h->launches += 2;
        }
        CK(cudaGetLastError());
        // the ranks' headers (4 words each, `cap` words apart) in one strided copy, then the only synchronisation of the exchange
        CK(cudaMemcpy2DAsync(h->h_xhdr, 4 * sizeof(uint64_t), h->d_xgather.p, cap * sizeof(uint64_t), 4 * sizeof(uint64_t), (size_t)W, cudaMemcpyDeviceToHost, s));
        if (attempt == 0)       // dense vector: node counts + the scalar statistics (exact integer sums), once
            NK(g_nccl.AllReduce(h->d_dense.p, h->d_dense.p, h->H.N + sizeof(WqCounters) / 8, WQ_NCCL_UINT64, WQ_NCCL_SUM, h->nccl_comm, s));
        CK(cudaStreamSynchronize(s));
        uint64_t need = 0;
        for (int r = 0; r < W; r++) need = std::max(need, h->h_xhdr[4 * r + 3]);
        const bool fits = need <= cap;
        // next capacity: a quarter above the largest record seen (identical on every rank: all ranks read the same headers)
        h->xcap = std::max<uint64_t>(1ull << 16, need + need / 4 + 1024);
        if (fits) break;
        if (attempt >= 2) return fail(-31, "sparse exchange did not fit after growing the buffers");
    }
    h->hq = WqHostQuant();
    h->hq_valid = false;             // the host-side stores are pulled from the merged tables
    reset_quant_ex(h);
    return 0;
}

int wq_classes_sync(wq_handle* h) {
    if (!h) return fail(-1, "null handle");
    if (!h->nccl_comm) return fail(-31, "wq_classes_sync needs wq_comm_init_rank first");
    if (int rc = wq_classes_build_local(h)) return rc;
    WqTimer tm_("classes_sync (exchange + assemble)");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int W = h->comm_size;
    const uint64_t mine = h->class_pack.size();
    for (int attempt = 0;; attempt++) {
        const uint64_t cap = h->ccap;                    // bytes per rank, multiple of 16: [u64 length][bytes]
        CK(h->d_cshare.ensure(cap)); CK(h->d_cgather.ensure(cap * (uint64_t)W));
        if (cap * (uint64_t)W > h->h_cgather_cap) {
            if (h->h_cgather) cudaFreeHost(h->h_cgather);
            h->h_cgather = nullptr; h->h_cgather_cap = 0;
            CK(cudaHostAlloc((void**)&h->h_cgather, cap * (uint64_t)W, cudaHostAllocDefault));
            h->h_cgather_cap = cap * (uint64_t)W;
        }
        CK(cudaMemcpyAsync(h->d_cshare.p, &mine, 8, cudaMemcpyHostToDevice, s));
        if (mine + 8 <= cap && mine) CK(cudaMemcpyAsync(h->d_cshare.p + 8, h->class_pack.data(), mine, cudaMemcpyHostToDevice, s));
        NK(g_nccl.AllGather(h->d_cshare.p, h->d_cgather.p, cap, WQ_NCCL_UINT8, h->nccl_comm, s));
        CK(cudaMemcpyAsync(h->h_cgather, h->d_cgather.p, cap * (uint64_t)W, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        uint64_t need = 0;
        for (int r = 0; r < W; r++) { uint64_t len; memcpy(&len, h->h_cgather + (uint64_t)r * cap, 8); need = std::max(need, len + 8); }
        const bool fits = need <= cap;
        h->ccap = (std::max<uint64_t>(1ull << 16, need + need / 4 + 1024) + 15) & ~15ull;
        if (fits) {
            std::vector<const void*> bufs(W);
            std::vector<uint64_t> lens(W);
            for (int r = 0; r < W; r++) { memcpy(&lens[r], h->h_cgather + (uint64_t)r * cap, 8); bufs[r] = h->h_cgather + (uint64_t)r * cap + 8; }
            // wq_classes_assemble updates ccap-independent state only; the next call's capacity was fixed above from the shared lengths
            return wq_classes_assemble(h, bufs.data(), lens.data(), (uint32_t)W);
        }
        if (attempt >= 2) return fail(-31, "class exchange did not fit after growing the buffers");
    }
}

// Event stage partitioned by gene (north_star: "EM is then partitioned by gene"): this handle builds and solves the events of
// the contiguous gene range `part` of `nparts` only; wq_process_events then returns that range's records.  Concatenating
// the outputs of parts 0..nparts-1 gives the unpartitioned output.
int wq_set_gene_partition(wq_handle* h, uint32_t part, uint32_t nparts) {
    if (!h) return fail(-1, "null handle");
    if (nparts == 0 || part >= nparts) return fail(-1, "gene partition out of range");
    if (h->part != part || h->nparts != nparts) {
        for (WqEvBatch& B : h->ev_batch) { if (B.launched) cudaStreamSynchronize(B.stream); B.built = B.launched = false; }
        h->events_done = false;
        h->ev_cuts.clear();
    }
    h->part = part; h->nparts = nparts;
    return 0;
}

// Sparse stores of this handle as one flat byte buffer for an all-gather between ranks:
//   u64 header[5] = {n_edge, n_long, n_multi, long_words, multi_words}
//   u64 edge_key[n_edge], u64 edge_count[n_edge]
//   u64 long_off[n_long+1], u64 long_count[n_long], u32 long_words[] (padded to 8 bytes)
//   u64 multi_off[n_multi+1], u64 multi_count[n_multi], u32 multi_words[] (padded)
// Two-call protocol: buf == NULL returns the size in *used.
int wq_sparse_pack(wq_handle* h, void* buf, uint64_t cap, uint64_t* used) {
    if (!h || !used) return fail(-1, "null argument");
    if (int rc = sync_host_quant(h)) return rc;
    const WqHostQuant& Q = h->hq;
    auto pad8 = [](uint64_t b) { return (b + 7) & ~7ull; };
    const uint64_t ne = Q.edge_key.size(), nl = Q.longs.size(), nm = Q.multis.size();
    uint64_t need = 40 + 16 * ne + 8 * (nl + 1) + 8 * nl + pad8(4 * Q.longs.words.size()) + 8 * (nm + 1) + 8 * nm + pad8(4 * Q.multis.words.size());
    *used = need;
    if (!buf) return 0;
    if (cap < need) return fail(-7, "sparse pack buffer too small");
    uint8_t* p = (uint8_t*)buf;
    uint64_t hdr[5] = {ne, nl, nm, Q.longs.words.size(), Q.multis.words.size()};
    memcpy(p, hdr, 40); p += 40;
    memcpy(p, Q.edge_key.data(), 8 * ne); p += 8 * ne;
    memcpy(p, Q.edge_count.data(), 8 * ne); p += 8 * ne;
    for (const WqKeyedStore* K : {&Q.longs, &Q.multis}) {
        memcpy(p, K->off.data(), 8 * (K->size() + 1)); p += 8 * (K->size() + 1);
        memcpy(p, K->count.data(), 8 * K->size()); p += 8 * K->size();
        memcpy(p, K->words.data(), 4 * K->words.size()); p += pad8(4 * K->words.size());
    }
    return 0;
}

// Merge another rank's wq_sparse_pack buffer into this handle's host-side stores.
int wq_sparse_merge(wq_handle* h, const void* buf, uint64_t len) {
    if (!h || !buf) return fail(-1, "null argument");
    if (len < 40) return fail(-1, "sparse buffer too short");
    if (int rc = sync_host_quant(h)) return rc;
    WqHostQuant& Q = h->hq;
    reset_quant_ex(h);
    auto pad8 = [](uint64_t b) { return (b + 7) & ~7ull; };
    const uint8_t* p = (const uint8_t*)buf;
    uint64_t hdr[5];
    memcpy(hdr, p, 40); p += 40;
    const uint64_t ne = hdr[0];
    uint64_t need = 40 + 16 * ne + 8 * (hdr[1] + 1) + 8 * hdr[1] + pad8(4 * hdr[3]) + 8 * (hdr[2] + 1) + 8 * hdr[2] + pad8(4 * hdr[4]);
    if (len < need) return fail(-1, "sparse buffer truncated");
    const uint64_t* ek = (const uint64_t*)p; p += 8 * ne;
    const uint64_t* ec = (const uint64_t*)p; p += 8 * ne;
    {   // both edge lists are sorted: linear merge
        std::vector<uint64_t> k2, c2;
        k2.reserve(Q.edge_key.size() + ne); c2.reserve(Q.edge_key.size() + ne);
        size_t i = 0, j = 0;
        while (i < Q.edge_key.size() || j < ne) {
            if (j >= ne || (i < Q.edge_key.size() && Q.edge_key[i] < ek[j])) { k2.push_back(Q.edge_key[i]); c2.push_back(Q.edge_count[i]); i++; }
            else if (i >= Q.edge_key.size() || ek[j] < Q.edge_key[i]) { k2.push_back(ek[j]); c2.push_back(ec[j]); j++; }
            else { k2.push_back(ek[j]); c2.push_back(Q.edge_count[i] + ec[j]); i++; j++; }
        }
        Q.edge_key.swap(k2); Q.edge_count.swap(c2);
    }
    WqKeyedStore* dst[2] = {&Q.longs, &Q.multis};
    for (int t = 0; t < 2; t++) {
        const uint64_t n = hdr[1 + t], nw = hdr[3 + t];
        const uint64_t* off = (const uint64_t*)p; p += 8 * (n + 1);
        const uint64_t* cnt = (const uint64_t*)p; p += 8 * n;
        const uint32_t* words = (const uint32_t*)p; p += pad8(4 * nw);
        for (uint64_t i = 0; i < n; i++) dst[t]->push(words + off[i], (size_t)(off[i + 1] - off[i]), cnt[i]);
    }
    // multi-map keys are classes: equal keys must become one record, in canonical order.  Long paths may stay
    // duplicated and unordered: every use of them is an order-free integer accumulation.
    Q.multis.canonicalize();
    return 0;
}

// Replace the dense part of the host quant state after an NCCL all-reduce of wq_counts_device_ptr.
int wq_counts_refresh_dense(wq_handle* h) {
    if (!h) return fail(-1, "null handle");
    if (int rc = sync_host_quant(h)) return rc;
    const uint32_t N = h->H.N;
    std::vector<uint64_t> dense(N + sizeof(WqCounters) / 8);
    CK(cudaMemcpy(dense.data(), h->d_dense.p, dense.size() * 8, cudaMemcpyDeviceToHost));
    h->hq.node.assign(dense.begin(), dense.begin() + N);
    memcpy(&h->hq.stats, dense.data() + N, sizeof(WqCounters));
    return 0;
}

// Class building partitioned by gene across ranks (multi-GPU: the host threads of a box are shared by its ranks).  After the
// count exchange every rank holds the same stores; rank `part` (wq_set_gene_partition) builds the multi-map classes of its
// key range and the isoform classes of its gene range, the shares are all-gathered by the caller, and every rank assembles
// the complete class set for wq_em_tpm.
int wq_classes_build_local(wq_handle* h) {
    if (!h) return fail(-1, "null handle");
    if (int rc = sync_host_quant(h)) return rc;
    WqTimer tm_("classes_build_local");
    std::string err = wq_host_build_classes(h->H, h->iso, h->hq, h->hx, h->part, h->nparts);
    if (!err.empty()) return fail(-11, err);
    wq_classes_pack_impl(h->iso, h->hx, h->class_pack);
    return 0;
}
int wq_classes_pack(wq_handle* h, const void** ptr, uint64_t* bytes) {
    if (!h || !ptr || !bytes) return fail(-1, "null argument");
    if (!h->hx.local_built) return fail(-11, "wq_classes_pack needs wq_classes_build_local first");
    *ptr = h->class_pack.data(); *bytes = h->class_pack.size();
    return 0;
}
int wq_classes_assemble(wq_handle* h, const void* const* bufs, const uint64_t* lens, uint32_t n) {
    if (!h || !bufs || !lens) return fail(-1, "null argument");
    if (n != h->nparts) return fail(-11, "wq_classes_assemble needs one share per part of wq_set_gene_partition");
    WqTimer tm_("classes_assemble");
    std::string err = wq_classes_assemble_impl(h->iso, h->hx, bufs, lens, n);
    if (!err.empty()) return fail(-11, err);
    return 0;
}

int wq_em_tpm(wq_handle* h, const wq_em_param* p, double* tpm, double* count, double* genetpm, int64_t* iterations) {
    if (!h || !p) return fail(-1, "null argument");
    if (int rc = sync_host_quant(h)) return rc;
    CK(cudaSetDevice(h->device));
    WqTimer tm_("em_tpm (classes + kernel)");
    // While the EM kernel runs, the host threads build the events of every gene that assign_ambig! cannot touch
    // (process_events, bin/whippet-quant.jl:181, uses the same read length as the EM, :163-165).
    static const bool overlap_on = !(getenv("WQ_EVENTS_OVERLAP") && atoi(getenv("WQ_EVENTS_OVERLAP")) == 0);
    int overlap_rc = 0;
    const std::function<void()> overlap = [&]() {
        WqTimer t2("  events of untouched genes (under the EM kernel)");
        { WqTimer t3("    finalize edges"); wq_finalize_edges(h->hq, h->hx); }
        wq_build_event_batch(h, p->readlen, 0, h->ev_batch[0]);
        // their PSI EM goes to its own stream right away: its blocks fill in beside the EM kernel's and the few events that need
        // all 1500 sweeps (a pure latency chain) start early
        { WqTimer t4("    batch 0 launch (uploads + kernels enqueued)"); overlap_rc = wq_launch_event_batch(h, h->ev_batch[0]); }
    };
    std::string err = wq_run_em_tpm(h->H, h->iso, h->hq, h->hx, h->em_scratch, *p, h->stream, &h->launches, tpm, count, genetpm, iterations,
                                    overlap_on ? &overlap : nullptr);
    if (!err.empty()) return fail(-11, err);
    return overlap_rc;
}

int wq_assign_ambig(wq_handle* h, wq_values* v) {
    if (!h || !v) return fail(-1, "null argument");
    if (int rc = sync_host_quant(h)) return rc;
    WqTimer tm_("assign_ambig");
    std::string err = wq_host_assign_ambig_impl(h->H, h->iso, h->hq, h->hx);
    if (!err.empty()) return fail(-12, err);
    WqHostQuantEx& X = h->hx;
    if (!X.edges_final) wq_finalize_edges(h->hq, X);
    uint64_t ne = 0, nc = 0;
    for (uint64_t k : X.edge_key) { if (wq_edge_is_circ(k)) nc++; else ne++; }
    bool sizing = !v->node_value && !v->edge_key && !v->circ_key;
    if (!sizing && (v->n_nodes < X.node_value.size() || v->n_edge < ne || v->n_circ < nc)) return fail(-7, "wq_values buffers too small");
    v->n_nodes = X.node_value.size(); v->n_edge = ne; v->n_circ = nc;
    if (sizing) return 0;
    if (v->node_value) memcpy(v->node_value, X.node_value.data(), X.node_value.size() * 8);
    uint64_t ie = 0, ic = 0;
    for (size_t i = 0; i < X.edge_key.size(); i++) {
        if (wq_edge_is_circ(X.edge_key[i])) { if (v->circ_key) { v->circ_key[ic] = X.edge_key[i]; v->circ_value[ic] = X.edge_value[i]; } ic++; }
        else { if (v->edge_key) { v->edge_key[ie] = X.edge_key[i]; v->edge_value[ie] = X.edge_value[i]; } ie++; }
    }
    return 0;
}

int wq_em_psi_batch(wq_handle* h, wq_psi_batch* b) {
    if (!h || !b) return fail(-1, "null argument");
    CK(cudaSetDevice(h->device));
    std::string err = wq_run_em_psi(*b, h->em_scratch, h->stream, &h->launches);
    if (!err.empty()) return fail(-13, err);
    return 0;
}

static int process_events_impl(wq_handle* h, int64_t readlen, wq_events_out* out, bool view);
int wq_process_events(wq_handle* h, int64_t readlen, wq_events_out* out) { return process_events_impl(h, readlen, out, false); }
int wq_process_events_view(wq_handle* h, int64_t readlen, wq_events_out* out) { return process_events_impl(h, readlen, out, true); }
static int process_events_impl(wq_handle* h, int64_t readlen, wq_events_out* out, bool view) {
    if (!h || !out) return fail(-1, "null argument");
    if (!h->hx.built || !h->hx.em_done || !h->hx.ambig_done) return fail(-14, "wq_process_events needs wq_em_tpm and wq_assign_ambig first (bin/whippet-quant.jl:177-181)");
    CK(cudaSetDevice(h->device));
    WqHostQuantEx& X = h->hx;
    if (!X.edges_final) wq_finalize_edges(h->hq, X);
    if (!h->events_done) {
        WqTimer tm_("process_events");
        const bool prof = getenv("WQ_PROFILE") != nullptr;
        auto lap = [&](const char* what) { if (prof) fprintf(stderr, "[wq]   events: %-22s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tm_.t0).count()); };
        WqEvBatch& B0 = h->ev_batch[0];
        WqEvBatch& B1 = h->ev_batch[1];
        if (!B0.built || B0.readlen != readlen) {            // not built under the EM kernel (or for another read length)
            if (B0.launched) CK(cudaStreamSynchronize(B0.stream));
            wq_build_event_batch(h, readlen, 0, B0);
            if (int rc = wq_launch_event_batch(h, B0)) return rc;
        }
        wq_build_event_batch(h, readlen, 1, B1);
        if (int rc = wq_launch_event_batch(h, B1)) return rc;
        lap("+ touched genes built");
        // records in gene order: per host-thread gene range, merge the two groups' records by gene
        const size_t nparts = B0.parts.size();
        if (B1.parts.size() != nparts) return fail(-14, "event batches were built with different host thread counts");

        std::vector<size_t> rec_at(nparts + 1, 0);
        std::vector<uint64_t> pool_paths(2 * nparts + 1, 0), pool_nodes(2 * nparts + 1, 0);      // group-major: (group, part)
        for (size_t t = 0; t < nparts; t++) rec_at[t + 1] = rec_at[t] + B0.parts[t].recs.size() + B1.parts[t].recs.size();
        for (int g = 0; g < 2; g++) for (size_t t = 0; t < nparts; t++) {
            const WqEventPool& pl = h->ev_batch[g].parts[t].pool;
            const size_t k = g * nparts + t;
            pool_paths[k + 1] = pool_paths[k] + (pl.path_ptr.size() - 1);
            pool_nodes[k + 1] = pool_nodes[k] + pl.path_nodes.size();
        }
        h->ev_recs.resize(rec_at[nparts]);
        h->ev_path_ptr.resize(pool_paths[2 * nparts] + 1);
        h->ev_path_nodes.resize(pool_nodes[2 * nparts]);
        h->ev_path_ptr[0] = 0;
        auto place = [&](size_t t) {
            const std::vector<WqEventRec>*rs[2] = {&B0.parts[t].recs, &B1.parts[t].recs};
            size_t i[2] = {0, 0}, o = rec_at[t];
            while (i[0] < rs[0]->size() || i[1] < rs[1]->size()) {
                const int g = (i[1] >= rs[1]->size() || (i[0] < rs[0]->size() && (*rs[0])[i[0]].gene < (*rs[1])[i[1]].gene)) ? 0 : 1;
                WqEventRec r = (*rs[g])[i[g]++];
                if (r.problem >= 0) r.problem += h->ev_batch[g].pshift[t];
                r.path_first += (uint32_t)pool_paths[g * nparts + t];
                h->ev_recs[o++] = r;
            }
            for (int g = 0; g < 2; g++) {
                const WqEventPool& pl = h->ev_batch[g].parts[t].pool;
                const size_t k = g * nparts + t;
                for (size_t q = 1; q < pl.path_ptr.size(); q++) h->ev_path_ptr[pool_paths[k] + q] = pool_nodes[k] + pl.path_ptr[q];
                std::copy(pl.path_nodes.begin(), pl.path_nodes.end(), h->ev_path_nodes.begin() + pool_nodes[k]);
            }
        };
        auto par = [&](const std::function<void(size_t)>& f) {          // pieces pulled by the host threads
            const unsigned nthr = (unsigned)std::min<size_t>(nparts, wq_host_threads());
            if (nthr <= 1) { for (size_t t = 0; t < nparts; t++) f(t); return; }
            std::atomic<size_t> next{0};
            std::vector<std::thread> th;
            for (unsigned k = 0; k < nthr; k++) th.emplace_back([&]() { for (size_t t; (t = next.fetch_add(1)) < nparts;) f(t); });
            for (auto& x : th) x.join();
        };
        par(place);
        lap("+ records merged");
        if (B0.launched) CK(cudaStreamSynchronize(B0.stream));           // PSI launches and their result copies
        if (B1.launched) CK(cudaStreamSynchronize(B1.stream));
        h->psi_kernel_ms = (B0.launched ? B0.dev.last_ms() : 0.f) + (B1.launched ? B1.dev.last_ms() : 0.f);
        h->psi_sweep_bytes = 0;
        for (WqEvBatch* B : {&B0, &B1}) {
            if (!B->launched) continue;
            const WqPsiProblems& P = B->P;
            for (size_t e = 0; e < P.size(); e++) {
                const uint64_t np = P.path_ptr[e + 1] - P.path_ptr[e], na = P.amb_ptr[e + 1] - P.amb_ptr[e];
                const uint64_t ap = P.amb_path_ptr[P.amb_ptr[e + 1]] - P.amb_path_ptr[P.amb_ptr[e]];
                h->psi_sweep_bytes += (uint64_t)B->its[e] * (24 * np + 4 * ap + 8 * na);
            }
        }
        B0.launched = B1.launched = false;
        lap("+ PSI EM kernels done");
        if (prof) fprintf(stderr, "[wq]   events: %zu records, %zu + %zu EM problems (untouched + touched genes)\n", h->ev_recs.size(), B0.P.size(), B1.P.size());
        // finish the records (src/events.jl:776-796) on the host threads: record ranges as merged above; the value array is
        // laid out by a counting pass + prefix sums.  Spliced events solved without EM pick up the Beta posterior CI that was
        // launched with their batch; the EM events' psi is only known now, so theirs is one more (small) launch.
        auto n_values_of = [&](const WqEventRec& r) -> uint32_t {
            if (r.kind == 2) {
                if (r.problem < 0) return r.n_utr;
                const WqEvBatch& B = h->ev_batch[r.batch];
                const uint32_t p0 = B.P.path_ptr[r.problem], p1 = B.P.path_ptr[r.problem + 1];
                for (uint32_t i = p0; i < p1; i++) if (std::isnan(B.psi[i])) return r.n_utr;
                return p1 - p0;
            }
            if (r.problem >= 0) { const WqPsiProblems& P = h->ev_batch[r.batch].P; return P.path_ptr[r.problem + 1] - P.path_ptr[r.problem]; }
            return 0;
        };
        std::vector<uint64_t> val_at(nparts + 1, 0);
        par([&](size_t t) { uint64_t c = 0; for (size_t i = rec_at[t]; i < rec_at[t + 1]; i++) c += n_values_of(h->ev_recs[i]); val_at[t + 1] = c; });
        for (size_t t = 0; t < nparts; t++) val_at[t + 1] += val_at[t];
        h->ev_values.resize(val_at[nparts]);
        par([&](size_t t) {
            uint64_t vo = val_at[t];
            for (size_t ri = rec_at[t]; ri < rec_at[t + 1]; ri++) {
                WqEventRec& r = h->ev_recs[ri];
                const WqEvBatch& B = h->ev_batch[r.batch];
                const WqPsiProblems& P = B.P;
                const double* psi = B.psi;
                const int32_t* its = B.its;
                r.value_start = (uint32_t)vo;
                if (r.kind == 2) {
                    bool ok = r.problem >= 0;
                    if (ok) {
                        const uint32_t p0 = P.path_ptr[r.problem], p1 = P.path_ptr[r.problem + 1];
                        for (uint32_t i = p0; i < p1; i++) if (std::isnan(psi[i])) ok = false;
                        if (ok) for (uint32_t i = p0; i < p1; i++) { double v = std::nearbyint(psi[i] * 10000.0) / 10000.0; h->ev_values[vo++] = std::isfinite(v) ? v : psi[i]; }
                        r.iterations_ = its[r.problem];
                    }
                    if (!ok) { for (uint32_t k = 0; k < r.n_utr; k++) h->ev_values[vo++] = 0.0; r.flags |= 1; r.total = 0.0; }
                } else if (r.problem >= 0) {
                    const uint32_t p0 = P.path_ptr[r.problem], p1 = P.path_ptr[r.problem + 1], ni = P.n_inc[r.problem];
                    double sm = 0.0;
                    for (uint32_t i = p0; i < p0 + ni; i++) sm += psi[i];
                    r.psi = sm;
                    for (uint32_t i = p0; i < p1; i++) h->ev_values[vo++] = psi[i];
                    r.iterations_ = its[r.problem];
                    r.kind = (0.0 <= r.psi && r.psi <= 1.0 && r.total > 0) ? 1 : 0;
                    if (r.kind == 1) { const size_t E = P.size(); r.ci_lo = B.ci_em[r.problem]; r.ci_hi = B.ci_em[E + r.problem]; }
                } else if (r.kind == 1) {
                    r.kind = (r.total > 0) ? 1 : 0;
                    if (r.kind == 1 && r.ci_slot >= 0) { const size_t nci = B.ci_psi.size(); r.ci_lo = B.ci_out[r.ci_slot]; r.ci_hi = B.ci_out[nci + r.ci_slot]; }
                }
                r.n_values = (uint32_t)(vo - r.value_start);
            }
        });
        lap("+ records finished");
        h->ev_out.resize(h->ev_recs.size());
        par([&](size_t t) {
            for (size_t i = rec_at[t]; i < rec_at[t + 1]; i++) {
                const WqEventRec& r = h->ev_recs[i];
                wq_event e;
                memset(&e, 0, sizeof e);
                e.gene = r.gene; e.node = r.node; e.motif = r.motif; e.kind = r.kind; e.flags = r.flags;
                e.n_utr = r.n_utr; e.n_inc = r.n_inc; e.n_exc = r.n_exc; e.value_start = r.value_start; e.path_start = r.path_first;
                e.iterations = r.iterations_; e.psi = r.psi; e.total = r.total; e.ci_lo = r.ci_lo; e.ci_hi = r.ci_hi;
                h->ev_out[i] = e;
            }
        });
        h->events_done = true;
    }
    const uint64_t nv = h->ev_values.size(), np = h->ev_path_ptr.size() - 1, npn = h->ev_path_nodes.size();
    if (view) {                          // lend the library's own arrays (valid until the count state changes or the handle dies)
        out->n_events = h->ev_out.size(); out->n_values = nv; out->n_paths = np; out->n_path_nodes = npn;
        out->events = h->ev_out.data(); out->values = h->ev_values.data(); out->path_ptr = h->ev_path_ptr.data(); out->path_nodes = h->ev_path_nodes.data();
        return 0;
    }
    bool sizing = !out->events && !out->values && !out->path_ptr && !out->path_nodes;
    if (!sizing && (out->n_events < h->ev_recs.size() || out->n_values < nv || out->n_paths < np || out->n_path_nodes < npn))
        return fail(-7, "wq_events_out buffers too small");
    out->n_events = h->ev_recs.size(); out->n_values = nv; out->n_paths = np; out->n_path_nodes = npn;
    if (sizing) return 0;
    if (!h->ev_out.empty()) memcpy(out->events, h->ev_out.data(), h->ev_out.size() * sizeof(wq_event));
    if (nv) memcpy(out->values, h->ev_values.data(), nv * 8);
    memcpy(out->path_ptr, h->ev_path_ptr.data(), (np + 1) * 8);
    if (npn) memcpy(out->path_nodes, h->ev_path_nodes.data(), npn * 4);
    return 0;
}

}  // extern "C"
