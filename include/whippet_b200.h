/*
 * whippet_b200.h — C ABI of the B200-native read -> splice-graph alignment + EM library.
 *
 * The reference (timbitz/Whippet.jl v1.6.2) has no FFI: the seams this library replaces are
 * the ordinary Julia calls made by bin/whippet-quant.jl:143-181.  Every entry point below
 * cites the reference function whose work it performs; a Julia host binds them with `ccall`
 * (see INTEGRATION.md).  Plain pointers and sizes only; all host buffers are caller-owned and
 * borrowed for the duration of a call; every function returns 0 on success and a negative
 * code on error, with a thread-local message available from wq_last_error().
 * There is no CPU fallback: wq_index_create fails when no sm_100 device is present.
 */
#ifndef WHIPPET_B200_H
#define WHIPPET_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Flat view of `GraphLib` (src/index.jl:5-15), `SpliceGraph{K}` (src/graph.jl:104-115),
 * `Edges{K}` (src/edges.jl:10-13) and `FMIndex{2,UInt32}` (FMIndexes 1.0.0; layout pinned by
 * test/test_index.jls).  Genes/nodes/isoforms are numbered from 1 as in Julia;
 * arrays are 0-based C arrays, i.e. gene g lives at index g-1.
 * ---------------------------------------------------------------------------------------- */
typedef struct wq_index_desc {
    int32_t   kmer;          /* GraphLib.kmer (K) */
    uint32_t  n_genes;       /* G */
    uint32_t  n_nodes;       /* N = sum of nodes over genes */
    uint32_t  n_isoforms;    /* I = sum of annotated isoforms over genes */
    uint64_t  text_len;      /* n = sum of graph sequence lengths = FM text length */
    const uint32_t* gene_offset;  /* [G]   GraphLib.offset: 0-based running text offset */
    const uint32_t* node_base;    /* [G+1] first node slot of each gene (CSR) */
    const uint32_t* nodeoffset;   /* [N]   SpliceGraph.nodeoffset, 1-based inside the gene sequence */
    const uint32_t* nodecoord;    /* [N]   SpliceGraph.nodecoord (genome coordinate; SAM only) */
    const uint32_t* nodelen;      /* [N]   SpliceGraph.nodelen */
    const uint8_t*  edgetype;     /* [N+G] SpliceGraph.edgetype; gene g (1-based), edge j at node_base[g-1]+(g-1)+(j-1) */
    const uint64_t* edgeleft;     /* [N+G] SGKmer bits; meaningful only where is_edge(edgetype,true)  (src/edges.jl:104-107) */
    const uint64_t* edgeright;    /* [N+G] SGKmer bits; meaningful only where is_edge(edgetype,false) (src/edges.jl:108-111) */
    const uint8_t*  seq4;         /* [n]   graph sequences back to back, one 4-bit one-hot code per byte (A=1 C=2 G=4 T=8, may be ambiguous) */
    const uint8_t*  bwt;          /* [n]   stored BWT symbols 0..3 (no `$` row) */
    uint64_t        sentinel;     /* FMIndex.sentinel: 1-based row of `$` in the (n+1)-row matrix */
    int64_t         count[4];     /* FMIndex.count */
    const uint32_t* sa;           /* [n]   FMIndex.samples at r=1 (src/index.jl:97): row i>=2 -> sa[i-2] */
    const uint32_t* left_ptr;     /* [4^K+1] CSR heads of Edges.left  (an #undef slot is an empty range) */
    const uint32_t* left_nodes;   /* [2*left_ptr[4^K]]  (gene,node) pairs, sorted as in src/edges.jl:20,91 */
    const uint32_t* right_ptr;    /* [4^K+1] CSR heads of Edges.right */
    const uint32_t* right_nodes;  /* [2*right_ptr[4^K]] */
    const uint32_t* iso_base;     /* [G+1] first isoform slot of each gene (GraphLibQuant.geneidx, src/quant.jl:183-189) */
    const uint32_t* iso_node_ptr; /* [I+1] CSR heads of SpliceGraph.annopath bitsets */
    const uint32_t* iso_nodes;    /* ascending 1-based node ids of each annotated path */
} wq_index_desc;

/* AlignParam, field for field (src/align.jl:2-19). */
typedef struct wq_align_param {
    int32_t mismatches;
    int32_t kmer_size;
    int32_t seed_try;
    int32_t seed_tolerate;
    int32_t seed_min_qual;
    int32_t seed_length;
    int32_t seed_buffer;
    int32_t seed_inc;
    int32_t seed_pair_range;
    int32_t _pad;
    double  score_range;
    double  score_min;
    uint8_t is_stranded;
    uint8_t is_paired;
    uint8_t is_pair_rc;
    uint8_t is_trans_ok;
    uint8_t is_circ_ok;
    uint8_t _pad2[3];
} wq_align_param;

/* A batch of FASTQRecords after `fill!` (src/record.jl:13-19): 2-bit sequence + phred bytes.
 * fixed_len = 0: seq_off / qual_off give the position of every read.  fixed_len = L > 0 (a batch of equal-capacity records, the usual
 * case for Illumina runs): read i starts at word i * ceil(L / 32) of seq2 and at byte i * ((L + 3) & ~3) of qual, len[i] <= L, and
 * seq_off / qual_off may be NULL -- the 16 bytes per read of offsets then never cross PCIe (the device derives them).
 * qual_bits = 0 or 8: one phred byte per base.  qual_bits = 4 (fixed_len batches only): `qual` holds 4-bit indices into qual_dict, two per
 * byte (base j of read i in byte i * ((ceil(L / 2) + 3) & ~3) + j / 2, even j in the low nibble) -- instruments that bin their qualities
 * (4 or 8 distinct values per run) then move half the quality bytes; the device expands them to phred bytes before the kernels run, so
 * every result is that of the expanded batch. */
typedef struct wq_read_batch {
    uint32_t        n_reads;
    uint32_t        fixed_len;
    const uint8_t*  len;        /* [n]   read length (ReadLengthInt = UInt8, src/align.jl:74) */
    const uint64_t* seq_off;    /* [n]   first 64-bit word of read i inside seq2 */
    const uint64_t* seq2;       /* 2-bit codes A=0 C=1 G=2 T=3, 32 per word, base j at bits 2*(j%32) */
    uint64_t        seq2_words;
    const uint64_t* qual_off;   /* [n]   first byte of read i inside qual */
    const uint8_t*  qual;       /* phred scores with the FASTQ offset already removed */
    uint64_t        qual_bytes;
    uint8_t         qual_bits;  /* 0 / 8: bytes; 4: dictionary indices (see above) */
    uint8_t         qual_dict[16];
    uint8_t         _pad2[7];
} wq_read_batch;

/* SGAlignNode + SGAlignScore (src/align.jl:48-62). */
typedef struct wq_align_node {
    uint32_t gene;
    uint32_t node;
    float    mistolerance;
    uint8_t  matches;
    uint8_t  mismatches;
    uint16_t _pad;
} wq_align_node;

/* SGAlignment (src/align.jl:76-82); the path is nodes[path_start .. path_start+npath). */
typedef struct wq_alignment {
    uint32_t offset;
    uint32_t path_start;
    uint8_t  offsetread;
    uint8_t  strand;
    uint8_t  isvalid;
    uint8_t  npath;
} wq_alignment;

/* Result of `ungapped_align` for a batch (src/align.jl:225-270 / src/paired.jl:42-125):
 * read i has nhits[i] alignments aln[hit_start[i] ..] in the reference's own hit order.
 * For paired batches `aln` holds mate-1 alignments and `aln_mate` the index-matched mate-2 ones. */
typedef struct wq_align_out {
    uint32_t*      hit_start;   /* [n] */
    uint8_t*       nhits;       /* [n]  0 = Nullable() */
    uint8_t*       flipped;     /* [n]  bit0: read 1 was reverse-complemented in place (src/align.jl:231-234); bit1: mate 2 was */
    wq_alignment*  aln;         /* [aln_cap] */
    wq_alignment*  aln_mate;    /* [aln_cap] or NULL for single-end */
    wq_align_node* nodes;       /* [nodes_cap] */
    uint64_t       aln_cap;
    uint64_t       nodes_cap;
    uint64_t       aln_used;    /* out */
    uint64_t       nodes_used;  /* out */
} wq_align_out;

/* (mapped, total, mean_readlen) of process_reads! (src/reads.jl:41-80) plus the multi-map total.
 * mean_readlen is the reference's RUNNING mean, `mean += (len - mean) / mapped` over the mapped reads in read order
 * (src/reads.jl:69, :121; floored by the caller, bin/whippet-quant.jl:145,151), not sum / mapped: the two can floor
 * differently at an integer boundary.  Fixed-length input keeps it equal to that length at no cost; for ragged input the
 * library replays the recurrence on the host from a per-read mapped bit (reads of this handle only: with reads sharded
 * over several handles the caller chains the shards in read order, or uses fixed-length input). */
typedef struct wq_map_stats {
    uint64_t total;
    uint64_t mapped;
    uint64_t multi;           /* reads routed to MultiMapping (length(align) > 1) */
    uint64_t sum_readlen;     /* sum of mapped read lengths */
    uint64_t path_overflow;   /* reads whose path exceeded WQ_MAX_PATH (reported, never silently dropped) */
    double   mean_readlen;    /* running mean as defined above (0 when nothing mapped) */
} wq_map_stats;

#define WQ_MAX_PATH 24        /* device-side bound on nodes per alignment path */
#define WQ_MAX_TOLERATE 8     /* device-side bound on AlignParam.seed_tolerate */

typedef struct wq_handle wq_handle;

const char* wq_last_error(void);
int  wq_version(void);

/* Upload + re-layout the index on `device` (GraphLib is the input of bin/whippet-quant.jl:105). */
int  wq_index_create(const wq_index_desc* desc, int device, wq_handle** out);
void wq_destroy(wq_handle* h);

/* The index in the library's native on-disk form.  The reference serialises GraphLib with Julia's
 * Serialization (bin/whippet-index.jl:97-101) and deserialises all of it on every start (bin/whippet-quant.jl:105, its first
 * @timer line).  wq_index_write stores the RE-LAID index instead (occurrence blocks, suffix array, 2-bit text, node / gene
 * records, bucket table, k-mer edge tables, annotation paths, and lib.info when seqname / strand are given) in one file of
 * page-aligned, checksummed sections; it is host-only and builds the sections with the same code wq_index_create uploads from.
 * wq_index_load maps such a file and copies the sections to the device: a handle equal to the one wq_index_create makes from
 * the same wq_index_desc (+ wq_index_set_gene_info).  WQ_INDEX_VERIFY=1 also checks the checksums of the large sections. */
int  wq_index_write(const wq_index_desc* desc, const char* const* seqname, const uint8_t* strand, const char* path);
int  wq_index_load(const char* path, int device, wq_handle** out);

/* GraphLibQuant{C,DefaultCounter}(lib) + MultiMapping{C,DefaultCounter}() (bin/whippet-quant.jl:125-126):
 * zero all device-side count stores. */
int  wq_quant_reset(wq_handle* h);

/* `--biascorrect` (bin/whippet-quant.jl:116-122: CounterType = JointBiasCounter, BiasModelType = JointBiasMod; src/bias.jl:103-270).
 * wq_set_biascorrect(h, 1) before the first batch (it resets the count stores).  From then on every mapped read also feeds the 5' hexamer /
 * GC-window model (count!(mod, read.sequence), src/reads.jl:59, 108) and files its hexamers and GC bins under the counter it is counted
 * in (push!, src/bias.jl:222-235); the quant stage then works with the Float64 counts of primer_adjust! (classes, EM) and, from
 * assign_ambig! on, of gc_adjust! (bin/whippet-quant.jl:156-169), inside wq_em_tpm / wq_assign_ambig / wq_process_events as before.
 * Everything is tallied in integers on the device and turned into Float64 in one fixed order per counter (ascending hexamer, then
 * ascending bin: the declared canonical order of the reference's Dict iteration), so results do not depend on thread timing.
 * Mapped reads must be at least 37 bases long (the reference throws BoundsError otherwise); one GPU per job in this version.
 * wq_bias_download returns, aligned with wq_counts_download and wq_keyed_download (kinds 0 and 1), every counter's count after
 * primer_adjust! and the factor gc_adjust! multiplies it by (< 0: the counter saw no GC window and is left alone), plus the
 * normalised model tables mod.fore / mod.back (4096 each; NULL to skip).  NULL arrays = sizes only. */
typedef struct wq_bias_counts {
    uint64_t n_nodes, n_edge, n_circ, n_long, n_multi;
    double *node_primer, *node_gc;
    double *edge_primer, *edge_gc;
    double *circ_primer, *circ_gc;
    double *long_primer, *long_gc;
    double *multi_primer, *multi_gc;
    double *fore, *back;
} wq_bias_counts;
int  wq_set_biascorrect(wq_handle* h, int on);
int  wq_bias_download(wq_handle* h, wq_bias_counts* out);

/* process_reads! body (src/reads.jl:56-71): ungapped_align (src/align.jl:225-270) for every read of
 * the batch, then count!/push!(multi) (src/quant.jl:556-594, 456-469) accumulated on the device.
 * `out` may be NULL (quantification only); when given the alignments are downloaded as well. */
int  wq_align_se(wq_handle* h, const wq_align_param* p, const wq_read_batch* reads, wq_align_out* out);

/* process_paired_reads! body (src/reads.jl:101-123): paired ungapped_align (src/paired.jl:42-125)
 * and the paired count! (src/quant.jl:597-653). */
int  wq_align_pe(wq_handle* h, const wq_align_param* p, const wq_read_batch* fwd, const wq_read_batch* rev,
                 wq_align_out* out);

/* Same as wq_align_se, but the batch already lives in device memory (pointers inside `reads`
 * are device pointers).  Used to time the kernels with inputs resident in HBM. */
int  wq_align_se_device(wq_handle* h, const wq_align_param* p, const wq_read_batch* reads);
int  wq_align_pe_device(wq_handle* h, const wq_align_param* p, const wq_read_batch* fwd, const wq_read_batch* rev);

int  wq_map_stats_get(wq_handle* h, wq_map_stats* out);

/* SAM output from the alignments of a batch: sam_flag, cigar_string, write_sam, write_sam_header (src/io.jl:69-276) as
 * process_reads! / process_paired_reads! call them with sam=true (src/reads.jl:63-67, 110-121).
 * wq_index_set_gene_info supplies lib.info (reference sequence name and strand per gene; src/types.jl:33-37), which
 * only SAM output needs.  wq_sam_batch formats every read of a batch aligned into `out` (in read order, on the host
 * threads); rev / rname_* are NULL for single-end batches.  buf == NULL returns the size in *used.  The @SQ lines of
 * the header are written in ascending name order (the reference iterates a Dict). */
int  wq_index_set_gene_info(wq_handle* h, const char* const* seqname, const uint8_t* strand);
int  wq_sam_header(wq_handle* h, char* buf, uint64_t cap, uint64_t* used);
int  wq_sam_batch(wq_handle* h, const wq_read_batch* fwd, const wq_read_batch* rev,
                  const uint64_t* name_off, const uint32_t* name_len, const char* names,
                  const uint64_t* rname_off, const uint32_t* rname_len, const char* rnames,
                  const wq_align_out* out, int qualoffset, char* buf, uint64_t cap, uint64_t* used);

/* FASTQ ingest: make_fqparser + read_chunk! + fill! (src/reads.jl:2-23, src/record.jl:13-19).  The reader inflates
 * (gzip or plain), splits 4-line records and packs them into a wq_read_batch on the host threads: every sequence byte
 * that is not one of ACGTacgt becomes A (FASTQ.Reader(..., fill_ambiguous = DNA_A)), quality = byte - qualoffset in
 * UInt8 arithmetic.  wq_fastq_next returns at most max_reads records (n_reads == 0 at the end of the input) in
 * page-locked buffers owned by the reader; they stay valid until the call after the next one.  wq_fastq_names gives the
 * identifiers of the batch returned last (name i = names[name_off[i] .. name_off[i] + name_len[i])).
 * wq_process_fastq is process_reads! / process_paired_reads! (src/reads.jl:41-129) over whole files (rev_path NULL =
 * single end): the next batch is parsed while the device aligns and counts the current one. */
typedef struct wq_fastq wq_fastq;
int  wq_fastq_open(const char* path, int qualoffset, wq_fastq** out);
int  wq_fastq_from_memory(const void* text, uint64_t len, int qualoffset, wq_fastq** out);   /* text must outlive the reader */
int  wq_fastq_next(wq_fastq* f, uint32_t max_reads, wq_read_batch* out);
int  wq_fastq_names(wq_fastq* f, const uint64_t** name_off, const uint32_t** name_len, const char** names);
void wq_fastq_close(wq_fastq* f);
int  wq_process_fastq(wq_handle* h, const wq_align_param* p, const char* fwd_path, const char* rev_path, int qualoffset,
                      uint32_t batch_reads, wq_map_stats* st);
/* The same with sam = true (src/reads.jl:49-52, 63-67, 110-121; `whippet-quant.jl --sam` writes to stdout): the header
 * (write_sam_header) and then the records of every batch, in read order, are written to the file descriptor sam_fd while
 * the next batch is being parsed (sam_fd < 0: no SAM output, i.e. wq_process_fastq).  Needs wq_index_set_gene_info. */
int  wq_process_fastq_sam(wq_handle* h, const wq_align_param* p, const char* fwd_path, const char* rev_path, int qualoffset,
                          uint32_t batch_reads, int sam_fd, wq_map_stats* st);

/* Plumbing with no reference counterpart: run the handle's work on a caller-owned CUDA stream
 * (NULL = the handle's own), per-kernel device milliseconds [seed, extend, resolve] of the last
 * wq_align_*_device call, page-locked host buffers for read batches, and a refresh of the host-side
 * dense counts after the caller all-reduced wq_counts_device_ptr over NCCL. */
int  wq_set_stream(wq_handle* h, void* cuda_stream);
int  wq_last_kernel_ms(wq_handle* h, float* ms3);
int  wq_pinned_alloc(void** p, uint64_t bytes);
int  wq_pinned_free(void* p);
int  wq_counts_refresh_dense(wq_handle* h);
/* sparse stores (edges, circ, long paths, multi-map classes) as one flat buffer for an all-gather between ranks;
 * buf == NULL returns the size.  wq_sparse_merge adds another rank's buffer into this handle. */
int  wq_sparse_pack(wq_handle* h, void* buf, uint64_t cap, uint64_t* used);
int  wq_sparse_merge(wq_handle* h, const void* buf, uint64_t len);
/* Device-side form of the same exchange (no reference counterpart): wq_sparse_device_pack compacts
 * this rank's edge/circ table and keyed stores into one DEVICE buffer owned by the handle (valid until the next pack);
 * the caller all-gathers the buffers (NCCL) and hands every other rank's buffer (a device pointer) to
 * wq_sparse_device_merge, which adds its records into this rank's device tables with a kernel.  The host-side stores are
 * rebuilt from the merged tables on the next download / wq_em_tpm. */
int  wq_sparse_device_pack(wq_handle* h, void** dptr, uint64_t* bytes);
int  wq_sparse_device_merge(wq_handle* h, const void* dbuf, uint64_t bytes);
/* The exchange step of the path on several GPUs (`*_counts_sync (NCCL)`), inside the library so that a
 * host without torch (the Julia CLI, one process per GPU) can run it.  libnccl.so.2 is opened at run time (no link-time
 * dependency; single-GPU use never touches it).
 *   wq_comm_unique_id   rank 0 fills 128 bytes (ncclUniqueId) and hands them to the other ranks by any means
 *   wq_comm_init_rank   every rank: communicator of this handle (its device, its stream)
 *   wq_counts_sync      after the last batch: ONE ncclAllReduce over the dense vector (node counts + statistics: every read
 *                       adds the integer 1, so the sum is exact) and ONE ncclAllGather of the compacted sparse stores (edge /
 *                       circ table, long paths, multi-map classes), merged into the device tables by kernels; one stream
 *                       synchronisation, nothing staged through the host.  Afterwards every rank holds the whole job's counts.
 *   wq_classes_sync     wq_classes_build_local + all-gather of the shares + wq_classes_assemble (see below)
 * Every rank must call these in the same order.  wq_set_gene_partition(rank, nranks) is implied by wq_comm_init_rank. */
int  wq_comm_unique_id(void* id128);
int  wq_comm_init_rank(wq_handle* h, int nranks, int rank, const void* id128);
int  wq_counts_sync(wq_handle* h);
int  wq_classes_sync(wq_handle* h);
int  wq_comm_destroy(wq_handle* h);

/* Event stage partitioned across ranks: this handle builds and solves the events of part `part` of `nparts` only and wq_process_events
 * returns that part's records; the parts concatenate to the full output.  A part is a contiguous range of the event work list (gene
 * ranges, and ranges of nodes inside a gene with many nodes), cut where the estimated cost of the path enumeration -- per node, from the
 * expressed edges inside the hull of the edges that span or touch it -- reaches part / nparts of the total, so that one gene whose nodes
 * sit under a long exon-skipping junction (the bulk of the stage in a deep human job) is shared by all ranks.  Every rank
 * holds the same counts after wq_counts_sync and therefore computes the same cuts; node / edge values (wq_assign_ambig) are complete
 * for every gene on every rank. */
int  wq_set_gene_partition(wq_handle* h, uint32_t part, uint32_t nparts);
/* Class building (build_equivalence_classes!, src/quant.jl:327-389) partitioned the same way: after the count exchange
 * every rank holds the same stores; wq_classes_build_local builds this part's share (the multi-map classes of its key
 * range, the isoform classes of its gene range), wq_classes_pack lends it as a host buffer owned by the handle, the caller
 * all-gathers the buffers, and wq_classes_assemble (shares of all parts, in part order) completes the class set that
 * wq_em_tpm then uses.
 * Without these calls wq_em_tpm builds every class itself. */
int  wq_classes_build_local(wq_handle* h);
int  wq_classes_pack(wq_handle* h, const void** ptr, uint64_t* bytes);
int  wq_classes_assemble(wq_handle* h, const void* const* bufs, const uint64_t* lens, uint32_t n);
uint64_t wq_launch_count(wq_handle* h);   /* kernels launched by this handle so far */
uint64_t wq_last_deferred(wq_handle* h);  /* extension tasks of the last wq_align_* call that went to the second-pass kernel (paths
                                           * longer than the first pass's shared-memory store, or a parked candidate) */
/* Measurement aid (no reference counterpart): device times of the quant-stage kernels of the last wq_em_tpm /
 * wq_process_events, taken with CUDA events on the launching stream, and the sizes their algorithmic-byte models use
 * (EM bytes/iteration = 12 * members + 8 * classes + 24 * isoforms).
 * out8 = {em_tpm kernel ms, PSI EM kernels ms, EM iterations, classes, class members, isoforms, PSI problems,
 *         PSI bytes = sum over events of sweeps * (24 * paths + 4 * ambiguous-set members + 8 * ambiguous sets)}. */
int  wq_quant_kernel_stats(wq_handle* h, double* out8);

/* Device pointer + element count of the dense count vector (u64 per node, then the scalar
 * stats) so a host can all-reduce it over NCCL between ranks. */
int  wq_counts_device_ptr(wq_handle* h, void** dptr, uint64_t* n_u64);

/* Download the count stores filled by count! (src/quant.jl:129-135).  Two-call protocol: call with
 * NULL buffers to get sizes, then with buffers.  Keys are sorted, which is also the canonical
 * class order of the EM.
 *   node_count  [N]              SpliceGraphQuant.node
 *   edge_key    [n_edge]         (gene<<32 | lnode<<16 | rnode), lnode<rnode  -> SpliceGraphQuant.edge
 *   circ_key    [n_circ]         same packing, lnode>=rnode                   -> SpliceGraphQuant.circ
 *   keyed paths: long paths (SpliceGraphQuant.long) and multi-map classes (MultiMapping.map) as
 *   records of u32 words, see wq_keyed_download. */
typedef struct wq_counts {
    uint64_t  n_nodes;   uint64_t* node_count;
    uint64_t  n_edge;    uint64_t* edge_key;  uint64_t* edge_count;
    uint64_t  n_circ;    uint64_t* circ_key;  uint64_t* circ_count;
} wq_counts;
int  wq_counts_download(wq_handle* h, wq_counts* c);

/* Keyed stores.  kind 0 = long paths, kind 1 = multi-map classes.  A record is
 *   long  SE: [npath, gene, node_1 .. node_npath]
 *   long  PE: [npath_fwd | npath_rev<<16, gene, fwd nodes.., rev nodes..]
 *   multi   : [nhits, then per hit a long-style record]
 * rec_ptr[i]..rec_ptr[i+1] delimit record i inside `words`; count[i] is its read count.
 * Records are sorted lexicographically by their words. */
typedef struct wq_keyed {
    uint64_t  n_rec;     uint64_t n_words;
    uint64_t* rec_ptr;   uint32_t* words;  uint64_t* count;
} wq_keyed;
int  wq_keyed_download(wq_handle* h, int kind, wq_keyed* k);

/* Merge counts produced by another rank/handle into this handle (host-side exchange of the sparse
 * stores; the dense vector goes through NCCL). */
int  wq_counts_merge(wq_handle* h, const wq_counts* c, const wq_keyed* longs, const wq_keyed* multis);

/* build_equivalence_classes! (src/quant.jl:327-389) + MultiCompat construction (src/quant.jl:394-435)
 * + calculate_tpm! (src/quant.jl:655-667) + gene_em! (src/quant.jl:719-769) + set_gene_tpm!
 * (src/quant.jl:248-257) on the device.  tpm/count are [I], genetpm is [G]; returns iterations. */
typedef struct wq_em_param {
    int64_t readlen;
    int64_t sig;       /* digits of gene_em!, CLI: 1 (the calculate_tpm! call before the loop rounds with its own default, 1: bin/whippet-quant.jl:163) */
    int64_t maxit;     /* CLI: 10000 */
} wq_em_param;
int  wq_em_tpm(wq_handle* h, const wq_em_param* p, double* tpm, double* count, double* genetpm, int64_t* iterations);

/* assign_ambig! (src/quant.jl:520-553): after it, node/edge/circ stores hold fractional values.
 *   node_value [N]; edge/circ as sorted (key,value) lists (two-call protocol). */
typedef struct wq_values {
    uint64_t  n_nodes;   double* node_value;
    uint64_t  n_edge;    uint64_t* edge_key;  double* edge_value;
    uint64_t  n_circ;    uint64_t* circ_key;  double* circ_value;
} wq_values;
int  wq_assign_ambig(wq_handle* h, wq_values* v);

/* Batched PSI EM: spliced_em! (src/events.jl:853-893) and rec_tandem_em! (src/events.jl:895-932),
 * one event per thread block.  Event e owns paths [path_ptr[e], path_ptr[e+1]) and ambiguous sets
 * [amb_ptr[e], amb_ptr[e+1]); ambiguous set a lists its member paths (event-local indices) in
 * amb_paths[amb_path_ptr[a] .. amb_path_ptr[a+1]). */
typedef struct wq_psi_batch {
    uint32_t n_events;   uint32_t _pad;
    const uint32_t* path_ptr;      /* [E+1] */
    const uint32_t* n_inc;         /* [E] the first n_inc paths of an event are inclusion paths (sums run exclusion first, src/events.jl:838) */
    const double*   path_count;    /* [P] unambiguous counts */
    const double*   path_length;   /* [P] */
    const uint32_t* amb_ptr;       /* [E+1] */
    const uint32_t* amb_path_ptr;  /* [A+1] */
    const uint32_t* amb_paths;     /* event-local path indices */
    const double*   amb_mult;      /* [A] AmbigCounts.multiplier */
    const uint8_t*  is_tandem;     /* [E] 1 = tandem-UTR variant (effective length max(len-readlen,1)) */
    int64_t readlen;
    int64_t sig;                   /* significant digits, reference: 4 */
    int64_t maxit;                 /* reference: 1500 */
    double*   psi;                 /* out [P] */
    double*   path_total;          /* out [P] counts after EM */
    int32_t*  iterations;          /* out [E] */
} wq_psi_batch;
int  wq_em_psi_batch(wq_handle* h, wq_psi_batch* b);

/* process_events / _process_events (src/events.jl:744-800, isnodeok=false): event construction on the host side
 * of the library from the stores left by wq_assign_ambig, all EM problems solved by one wq_em_psi_batch launch.
 * The writers of src/io.jl are not reproduced: every record carries the values they print.
 *   kind 0 = `output_empty` line, 1 = PSI line, 2 = tandem-UTR lines (n_utr nodes), 3 = no line
 *   flags bit0 = tandem event without enough reads (NA columns), bit1 = more than 1024 paths (the reference
 *   subsamples at random there; not reproducible)
 * values[value_start ..]: kind 2 -> n_utr rounded PSI values; kind 1 -> n_inc + n_exc path PSI values (when both exist)
 * paths: path_ptr[path_start + k] .. path_ptr[path_start + k + 1] delimit the node ids of path k in path_nodes
 * (utr paths, or inclusion paths followed by exclusion paths). */
typedef struct wq_event {
    uint32_t gene, node;
    uint8_t  motif, kind, flags, _pad;
    uint32_t n_utr, n_inc, n_exc;
    uint32_t value_start, path_start;
    int32_t  iterations;
    uint32_t _pad2;
    double   psi, total;
    double   ci_lo, ci_hi;     /* beta_posterior_ci(psi, total, sig = 3) (src/events.jl:790, 816-828); kind == 1 only */
} wq_event;
typedef struct wq_events_out {
    uint64_t n_events;     wq_event* events;
    uint64_t n_values;     double*   values;
    uint64_t n_paths;      uint64_t* path_ptr;      /* [n_paths+1] */
    uint64_t n_path_nodes; uint32_t* path_nodes;
} wq_events_out;
int  wq_process_events(wq_handle* h, int64_t readlen, wq_events_out* out);   /* two-call protocol (NULL buffers -> sizes) */
/* Same result without the copy: the pointers of `out` are set to arrays owned by the handle, valid until the count state
 * changes (wq_quant_reset, wq_align_*, merges, wq_set_gene_partition) or the handle is destroyed. */
int  wq_process_events_view(wq_handle* h, int64_t readlen, wq_events_out* out);

#ifdef __cplusplus
}
#endif
#endif /* WHIPPET_B200_H */
