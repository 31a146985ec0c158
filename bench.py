#!/usr/bin/env python
"""Headline benchmark: aligned+quantified reads/sec of the read -> splice-graph hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5]

--config selects one of BASELINE.json's configs (default 2, the one the metric is quoted on):
  1  test/test_index.jls + 100k simulated SE reads (lengths 11-21 from the annotated isoforms; the literal
     101-bp reads cannot come from a 65-nt gene and are run as the reject-path parity case)
  2  hg19-scale splice graphs + 10 M x 101 bp SE reads per GPU.  The GENCODE GTF is not available offline; the graphs are
     built from the transcript structures of the reference's own anno/refseq_hg19.flat.gz (wq_inputs/graphbuild.py) over
     a random-sequence genome (--index synth40k selects round 1's fully synthetic 40 000-gene index instead)
  3  the same index with 200 paralog families + 20 M 2x101 bp read pairs per GPU (--pair-range 5000)
  4  mm10-scale synthetic index (50 000 genes) + 50 M x 150 bp SE reads in total, sharded over the GPUs (STRONG scaling)
  5  read-length sweep {50,75,101,125,150,200,255} x 1 M reads on the config-2 index + 2 000 stress genes whose cassette
     events have 64-1024 paths, driven through process_events

A step = one pass of the hot path over every part of the config: reset, seed + extend + resolve/count [+ exchange],
classes + transcript EM + multi-map assignment + events / PSI EM.  `value` has the batch resident in HBM; `e2e` goes
through wq_align_se / wq_align_pe with pinned HOST buffers (H2D inside the timed region) and reads the results back.

Every run ends with a parity check: the GPU path (through the C ABI) and the oracle process the same sample of the
workload and every output is compared bit for bit (statistics, node / edge / keyed counts, TPM, EM iterations,
multi-map assignment, every event record).  A mismatch fails the run.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402

FIXTURE = os.path.join(ROOT, "tests", "golden", "test_index_flat.npz")
SWEEP = (50, 75, 101, 125, 150, 200, 255)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=int(os.environ.get("WQ_BENCH_CONFIG", 2)), choices=[1, 2, 3, 4, 5])
    ap.add_argument("--index", default=os.environ.get("WQ_BENCH_INDEX", "refseq"), choices=["refseq", "synth40k"])
    ap.add_argument("--reads", type=int, default=int(os.environ.get("WQ_BENCH_READS", 0)), help="override the config's read count")
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("WQ_BENCH_CPU_SAMPLE", 1_500_000)))
    ap.add_argument("--no-parity", action="store_true", help="skip the parity check (profiling runs only; the line says so)")
    ap.add_argument("--no-fastq", action="store_true", help="skip the file -> result measurement (e2e_fastq)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ workloads
class Part:
    """One job of a config: an index, a read batch (or pair of batches), the alignment parameters."""

    def __init__(self, name, idx, idx_key, param, reads, mate=None, parity_reads=None, parity_idx=None, note=""):
        self.name, self.idx, self.idx_key, self.param, self.reads, self.mate = name, idx, idx_key, param, reads, mate
        self.parity_reads = parity_reads          # (reads, mate) the parity check runs, default: a prefix of `reads`
        self.parity_idx = parity_idx              # a smaller instance of the same generator when the oracle cannot do the full index
        self.note = note
        lens = np.asarray(reads.len)
        self.readlen = int(lens[0]) if len(lens) and (lens == lens[0]).all() else None      # None: ragged (running mean decides)

    @property
    def n(self):
        return self.reads.n


_index_cache = {}


def get_index(key, a):
    import whippet_jl_b200 as wq
    from wq_inputs import synth, graphbuild
    if key in _index_cache:
        return _index_cache[key]
    t0 = time.time()
    if key == "fixture":
        idx = wq.FlatIndex.load(FIXTURE)
    elif key == "refseq":
        idx = graphbuild.index_from_structure(graphbuild.load_structure(), K=9, seed=1)
    elif key == "refseq+paralogs":
        idx = graphbuild.index_from_structure(graphbuild.load_structure(), K=9, seed=1, paralog_families=200)
    elif key == "synth40k":
        idx = synth.make_index(n_genes=40000, K=9, seed=1)
    elif key == "synth40k+paralogs":
        idx = synth.make_index(n_genes=40000, K=9, seed=1, paralog_frac=0.0125)
    elif key == "mm10scale":
        idx = synth.make_index(n_genes=50000, K=9, seed=4, max_isoforms=3)
    elif key == "stress":
        idx = synth.make_stress_index(n_genes=2000, K=9, seed=5)
    elif key == "stress-small":
        idx = synth.make_stress_index(n_genes=120, K=9, seed=5)
    else:
        raise KeyError(key)
    print(f"[bench] index {key}: G={idx.n_genes} n={idx.text_len} N={idx.n_nodes} I={idx.n_isoforms} in {time.time() - t0:.1f}s",
          file=sys.stderr, flush=True)
    _index_cache[key] = idx
    return idx


def build_parts(a, rank, world):
    import whippet_jl_b200 as wq
    from wq_inputs import synth
    cfg = a.config
    t0 = time.time()
    parts = []
    if cfg == 1:
        idx = get_index("fixture", a)
        p = wq.AlignParam(1, 2, 4, 4, 4, 5, 1, 2, 1000, 0.05, 0.7, False, False, True, False, True)   # test/runtests.jl:283-284
        n = a.reads or 100_000
        parts.append(Part("fixture 11-21 bp", idx, "fixture", p, synth.fixture_reads(idx, n, seed=20261019 + rank)))
        rng = np.random.default_rng(7 + rank)
        lit = wq.ReadBatch.from_codes(rng.integers(0, 4, (n, 101), dtype=np.uint8), np.full((n, 101), 33, np.uint8))
        parts.append(Part("fixture literal 101 bp (reject path)", idx, "fixture", p, lit))
    elif cfg == 2:
        key = a.index
        idx = get_index(key, a)
        n = a.reads or 10_000_000
        parts.append(Part(f"{n} x 101 bp SE", idx, key, wq.AlignParam(), synth.simulate_reads(idx, n, 101, seed=2 + rank)))
    elif cfg == 3:
        key = a.index + "+paralogs"
        idx = get_index(key, a)
        n = a.reads or 20_000_000
        p = wq.AlignParam(seed_pair_range=5000, is_paired=True, is_pair_rc=True)
        f, r = synth.simulate_reads(idx, n, 101, seed=3 + rank, paired=True)
        parts.append(Part(f"{n} x 2x101 bp PE", idx, key, p, f, r))
    elif cfg == 4:
        idx = get_index("mm10scale", a)
        total = a.reads or 50_000_000
        lo, hi = total * rank // world, total * (rank + 1) // world          # contiguous shard of the global read range
        parts.append(Part(f"{total} x 150 bp SE in total, shard [{lo}, {hi})", idx, "mm10scale", wq.AlignParam(),
                          synth.simulate_reads(idx, hi - lo, 150, seed=4 + 1000 * rank)))
    elif cfg == 5:
        key = a.index
        idx = get_index(key, a)
        n = a.reads or 1_000_000
        tx = synth.Transcriptome(idx)
        for R in SWEEP:
            parts.append(Part(f"sweep {R} bp", idx, key, wq.AlignParam(), synth.simulate_reads(idx, n, R, seed=50 + R + 1000 * rank, tx=tx)))
        sidx, small = get_index("stress", a), get_index("stress-small", a)
        stx = synth.Transcriptome(sidx, synth.stress_paths(sidx, seed=6))
        smalltx = synth.Transcriptome(small, synth.stress_paths(small, seed=6))
        parts.append(Part("stress genes (64-1024-path events)", sidx, "stress", wq.AlignParam(),
                          synth.simulate_reads(sidx, 2 * n, 101, seed=60 + 1000 * rank, tx=stx, random_frac=0.0),
                          parity_reads=(synth.simulate_reads(small, 200_000, 101, seed=61, tx=smalltx, random_frac=0.0), None),
                          parity_idx=("stress-small", small),
                          note="parity on a 120-gene instance of the same generator (the oracle's event enumeration needs ~0.2 s per stress gene)"))
    print(f"[bench r{rank}] config {cfg}: {len(parts)} part(s), {sum(p.n for p in parts)} reads in {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
    return parts


def workload_name(a, parts):
    names = {1: "config 1: test/test_index.jls + 100k simulated SE reads (lengths 11-21 from the annotated isoforms, plus the literal 101-bp reject case)",
             2: "config 2: hg19-scale splice graphs (%s) + %d x 101 bp SE reads per GPU",
             3: "config 3: hg19-scale splice graphs (%s) with 200 paralog families + %d x 2x101 bp PE read pairs per GPU",
             4: "config 4: mm10-scale synthetic splice graphs (50 000 genes) + %d x 150 bp SE reads in total, sharded over the GPUs",
             5: "config 5: read-length sweep {50..255} x %d reads on the config-2 index (%s) + 2000 stress genes with 64-1024-path events"}
    src = ("built from the transcript structures of the reference's anno/refseq_hg19.flat.gz over a random-sequence genome; stands in for "
           "gencode_hg19.v25.tsl1, which is not available offline" if a.index == "refseq" else "40 000 synthetic genes, random-sequence genome")
    if a.config == 1:
        return names[1]
    if a.config in (2, 3):
        return names[a.config] % (src, parts[0].n)
    if a.config == 4:
        return names[4] % (a.reads or 50_000_000)
    return names[5] % (parts[0].n, src)


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every few ms DURING the timed region
    (same fields as the nvidia-smi clocks line of B200_PROFILING.md)."""

    def __init__(self, gpu):
        self.gpu, self.sm, self.reasons, self.stop_flag, self.ok = gpu, [], set(), False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(gpu).uuid)
            except Exception:
                pass
            self.h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        self.h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        try:
                            self.h = pynvml.nvmlDeviceGetHandleByUUID(cand)
                            break
                        except Exception:
                            self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def start(self):
        if not self.ok:
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self.stop_flag = True
        self.t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------ CPU side (the oracle)
def byte_model(ctr, nreads, R, paired, ftab_k, seed_len=18):
    """Algorithmic bytes per read (per pair), SURVEY.md section 8d, from oracle-instrumented counters.  A backward search
    that starts from the k-mer table reads one 8-byte table entry and then one 64-byte step (two 32-byte occurrence
    blocks) per symbol beyond the table; searches that die inside the table's depth read the entry only."""
    hist = ctr["steps_hist"]
    searches = sum(hist)
    k = ftab_k if (ftab_k and seed_len >= ftab_k) else 0
    steps = sum(h * max(0, s - k) for s, h in enumerate(hist))
    S = steps / nreads
    H = ctr["hits"] / nreads
    nodes = ctr["nodes_touched"] / max(1, ctr["hits"])
    r_in = ((R + 3) // 4 + R + 8) * (2 if paired else 1)
    a_out = (8 * ctr["out_alns"] + 16 * ctr["out_nodes"]) / nreads
    C = ctr["count_updates"] / nreads
    per_kernel = {"seed": r_in + 64 * S + (8 * searches / nreads if k else 0) + 4 * H,
                  "extend": H * ((R + 1) // 2 + 25 * nodes) + a_out, "resolve": 8 * C}
    return per_kernel, dict(S_steps_beyond_table=S, S_steps_plain=ctr["fm_steps"] / nreads, searches=searches / nreads, ftab_k=k,
                            H=H, nodes_per_hit=nodes, R_in=r_in, A_out=a_out, C=C)


def oracle_process(o, part_param, reads, mate):
    if mate is None:
        o.process_reads(part_param, reads)
    else:
        o.process_paired_reads(part_param, reads, mate)


def cpu_oracle_run(part, nsample, threads):
    """Times the oracle (CPU restatement) on the first `nsample` reads of a part: alignment + counting sharded over `threads`
    host threads, then the merged stores go through classes + transcript EM + multi-map assignment + events once.
    Returns the full-workload rate extrapolated from the sample, and the oracle itself (the parity check reads it)."""
    from oracle import Oracle, counters
    rb, mate = part.reads, part.mate
    nsample = min(nsample, rb.n)
    per = (nsample + threads - 1) // threads
    oracles = [Oracle(part.idx) for _ in range(threads)]
    sl = [(rb.slice(i * per, min(nsample, (i + 1) * per)), mate.slice(i * per, min(nsample, (i + 1) * per)) if mate is not None else None)
          for i in range(threads)]
    counters(reset=True)
    ctr = {}

    def work(i):
        oracle_process(oracles[i], part.param, sl[i][0], sl[i][1])
        if i == 0:
            ctr.update(counters(reset=True))

    t0 = time.perf_counter()
    if threads == 1:
        work(0)
    else:
        ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    t_align = time.perf_counter() - t0
    for o in oracles[1:]:
        oracles[0].merge(o)
    o = oracles[0]
    st = o.stats()
    # readlen = Int(floor(mean_readlen)) (bin/whippet-quant.jl:145,151); the running mean needs one sequential pass
    rl = part.readlen if part.readlen is not None else (int(np.floor(st["mean_readlen"])) if threads == 1 else int(st["sum_readlen"] // max(1, st["mapped"])))
    res = dict(readlen=rl)
    res["em"] = o.em_tpm(rl, 1, 10000)
    o.assign_ambig()
    res["events"] = o.events_flat(rl)
    dt = time.perf_counter() - t0
    # whole-job rate: alignment scales with the reads, the quant stage is a per-job cost -> extrapolate the sample to the
    # full workload (this favours the CPU arm compared with dividing the sample by its own wall time)
    full = rb.n
    rate = full / (full * t_align / nsample + (dt - t_align))
    return dict(rate=rate, dt=dt, t_align=t_align, ctr=ctr, n0=sl[0][0].n, nsample=nsample, oracle=o, res=res, stats=st)


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    parts = build_parts(a, 0, 1)
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    have_julia = subprocess.call("julia --version", shell=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 0
    vals = []
    total_reads = sum(p.n for p in parts)
    budget = a.cpu_sample
    for i in range(a.warmup + a.steps):
        # bounded sample: the per-job quant stage of the restatement is single-threaded (like the reference) and costs ~20 s per
        # step whatever the sample, so the alignment sample is kept small enough for W + K steps to end within a few minutes
        t_full, dt_all, ns_all = 0.0, 0.0, 0
        for p in parts:
            nsample = min(p.n, max(budget * 2 // 3 // len(parts), 125_000 * threads // len(parts)))
            r = cpu_oracle_run(p, nsample, threads)
            t_full += p.n / r["rate"]
            dt_all += r["dt"]
            ns_all += r["nsample"]
        if i >= a.warmup:
            vals.append((total_reads / t_full, dt_all, ns_all))
        if i == 0 and dt_all * (a.warmup + a.steps) > 240:   # keep the whole run within a few minutes
            budget = max(100000, int(budget * 240 / (dt_all * (a.warmup + a.steps))))
    v = float(np.mean([x[0] for x in vals]))
    sample = (f"first {vals[-1][2]} reads of the workload per step (over {len(parts)} part(s)) aligned+counted on {threads} host threads, then classes + "
              f"EM + multi-map assignment + events once on the merged counts; value = full-workload rate extrapolated from the sample "
              f"(alignment time scaled to {total_reads} reads + the per-job quant time)")
    print(json.dumps({
        "impl": "reference", "metric": "aligned+quantified reads/sec", "value": v, "unit": "reads/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000 * float(np.mean([x[1] for x in vals])),
        "higher_is_better": True, "scaling": "strong" if a.config == 4 else "weak", "vs_baseline": None, "dtype": "u8/u32 integer + f32 scores",
        "data": "synthetic", "config": {"workload": workload_name(a, parts)},
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "oracle C++ restatement of the Julia path (julia %s on this box)" % ("present but unused" if have_julia else "absent")},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------ GPU side
class PartRun:
    """Device-resident and pinned-host copies of one part's reads, and the calls that run it."""

    def __init__(self, part, q, dev, torch):
        import ctypes as C
        from whippet_jl_b200 import abi
        self.part, self.q = part, q
        self.keep = []

        def batch(rb, to, fixed):
            # fixed > 0: an equal-length batch at the strides of wq_read_batch.fixed_len; its offsets stay on the host (16 B per read less H2D)
            q4 = getattr(rb, "qual4", None) if fixed else None         # 4-bit quality indices (binned qualities): half the quality bytes
            t = [to(rb.len), None if fixed else to(rb.seq_off.view(np.int64)), to(rb.seq2.view(np.int64)),
                 None if fixed else to(rb.qual_off.view(np.int64)), to(rb.qual if q4 is None else q4)]
            self.keep.append(t)
            b = abi.ReadBatchC()
            b.n_reads = rb.n
            b.fixed_len = fixed
            b.len = C.cast(t[0].data_ptr(), abi.u8p)
            b.seq2 = C.cast(t[2].data_ptr(), abi.u64p)
            b.seq2_words = len(rb.seq2)
            b.qual = C.cast(t[4].data_ptr(), abi.u8p)
            b.qual_bytes = len(rb.qual) if q4 is None else len(q4)
            if q4 is not None:
                b.qual_bits = 4
                for k in range(16):
                    b.qual_dict[k] = int(rb.qual_dict[k])
            if not fixed:
                b.seq_off = C.cast(t[1].data_ptr(), abi.u64p)
                b.qual_off = C.cast(t[3].data_ptr(), abi.u64p)
            return b
        on_dev = lambda x: torch.from_numpy(x).to(dev)
        pinned = lambda x: torch.from_numpy(x).pin_memory()
        rbs = [part.reads] + ([part.mate] if part.mate is not None else [])
        self.dev = [batch(rb, on_dev, 0) for rb in rbs]
        # the host-side (e2e) copies: equal-capacity batches go in the fixed_len layout of the ABI (no offset arrays, quality rows padded to 4)
        hrbs = [rb.to_fixed(pack_qual=True) or rb for rb in rbs]
        fixed = [rb.fixed_stride() for rb in hrbs]
        self.fixed_len = fixed
        self.host = [batch(rb, pinned, f) for rb, f in zip(hrbs, fixed)]
        qbytes = lambda rb, f: rb.qual4.nbytes if f and getattr(rb, "qual4", None) is not None else rb.qual.nbytes
        self.h2d = int(sum(rb.len.nbytes + rb.seq2.nbytes + qbytes(rb, f) + (0 if f else rb.seq_off.nbytes + rb.qual_off.nbytes) for rb, f in zip(hrbs, fixed)))
        self.layout = [("fixed_len %d rows without offset arrays" % f + (", qualities as 4-bit dictionary indices" if getattr(rb, "qual4", None) is not None else ""))
                       if f else "per-read offset arrays, phred bytes" for rb, f in zip(hrbs, fixed)]
        self.pc = part.param.c()

    def align(self, host):
        import ctypes as C
        from whippet_jl_b200 import lib
        L, q, b = self.q.L, self.q, (self.host if host else self.dev)
        lib.check(L.wq_quant_reset(q.h))
        if len(b) == 1:
            rc = L.wq_align_se(q.h, C.byref(self.pc), C.byref(b[0]), None) if host else L.wq_align_se_device(q.h, C.byref(self.pc), C.byref(b[0]))
        else:
            rc = (L.wq_align_pe(q.h, C.byref(self.pc), C.byref(b[0]), C.byref(b[1]), None) if host
                  else L.wq_align_pe_device(q.h, C.byref(self.pc), C.byref(b[0]), C.byref(b[1])))
        lib.check(rc)


def quantify(q, part, world, phase=None):
    """Exchange + build_equivalence_classes! + calculate_tpm! + gene_em! + set_gene_tpm! + assign_ambig! + process_events"""
    t0 = time.perf_counter()
    if world > 1:
        q.counts_sync()                 # ONE NCCL all-reduce of the dense counts + all-gather / merge of the sparse stores (inside the library)
    t1 = time.perf_counter()
    if world > 1:
        q.build_classes_distributed()   # each rank builds the classes of its gene range; shares all-gathered over NCCL
    rl = part.readlen if part.readlen is not None else int(np.floor(q.stats().mean_readlen))
    it, tpm, cnt, gt = q.gene_em(rl, 1, 10000)
    q.assign_ambig_inplace()
    ev = q.process_events_view(rl)      # event construction + batched PSI EM (records lent by the library, no copy)
    t2 = time.perf_counter()
    if phase is not None:
        phase["exchange"] += t1 - t0
        phase["quant"] += t2 - t1
        phase["em_iterations"] = it
        phase["events"] = len(ev[0])
    return it, tpm, cnt, gt, ev, rl


def fastq_e2e(run, world):
    """File -> result through wq_process_fastq (inflate / map, newline scan, 2-bit + phred packing on the host threads, pipelined with
    the alignment of the previous batch) + the same quant stage, on FASTQ renderings of a prefix of the bench reads: a plain
    file, a blocked-gzip (BGZF) file and an ordinary single-member gzip file (serial inflate).  Files live in /dev/shm or the
    temp directory and are in the page cache (just written)."""
    import gzip
    import tempfile
    from wq_inputs import synth
    part, q = run.part, run.q
    rb = part.reads
    d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    res = {"unit": "reads/s", "note": "reads / (wq_quant_reset + wq_process_fastq + classes + EM + multi-map assignment + events), one GPU; "
                                      "the text is a rendering of the first reads of the bench batch"}
    plan = [("plain", min(rb.n, 4_000_000)), ("bgzf", min(rb.n, 2_000_000)), ("gzip", min(rb.n, 1_000_000))]
    text = synth.fastq_bytes(rb, 0, plan[0][1])
    rec = len(text) // plan[0][1]
    for kind, n in plan:
        path = os.path.join(d, f"wq_bench_{os.getpid()}_{kind}.fastq" + ("" if kind == "plain" else ".gz"))
        t = text[:n * rec]
        try:
            if kind == "plain":
                open(path, "wb").write(t)
            elif kind == "bgzf":
                open(path, "wb").write(synth.bgzf_compress(t, threads=min(32, os.cpu_count() or 8)))
            else:
                with gzip.open(path, "wb", compresslevel=1) as fh:
                    fh.write(t)
            best = None
            for rep in range(2):
                t0 = time.perf_counter()
                q.reset()
                st = q.process_fastq(part.param, path)
                quantify(q, part, world)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            assert st.total == n, (st.total, n)
            res[kind] = {"value": n / best, "reads": n, "file_bytes": os.path.getsize(path), "text_bytes": len(t), "seconds": best}
        finally:
            if os.path.exists(path):
                os.remove(path)
    return res


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype.itemsize == b.dtype.itemsize and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def parity_check(a, parts, handles, world, rank, dist, oracle_runs):
    """GPU (through the C ABI) vs oracle on a sample of every part, everything bit for bit.  With several ranks the sample is
    sharded over the ranks, exchanged like the real job, and rank 0 compares (the event stage unpartitioned there)."""
    import whippet_jl_b200 as wq
    from oracle import Oracle, canonical_events
    out, ok = [], True
    per_part = max(20_000, a.cpu_sample // max(1, len(parts))) if len(parts) > 1 else a.cpu_sample
    for pi, part in enumerate(parts):
        idx, q = part.idx, handles[part.idx_key]
        if part.parity_idx is not None:
            key, idx = part.parity_idx
            if key not in handles:
                handles[key] = wq.GraphLibQuant(idx, device=q.device)
                if world > 1:
                    handles[key].comm_init(dist)
            q = handles[key]
        if part.parity_reads is not None:
            rb, mate = part.parity_reads
        else:
            ns = min(per_part, part.n)
            rb, mate = part.reads.slice(0, ns), (part.mate.slice(0, ns) if part.mate is not None else None)
        # ---- GPU: every rank aligns its slice of the sample (rank 0's sample, broadcast when there are several ranks)
        if world > 1:
            obj = [(rb, mate)] if rank == 0 else [None]
            dist.broadcast_object_list(obj, src=0)
            rb, mate = obj[0]
        n = rb.n
        lo, hi = n * rank // world, n * (rank + 1) // world
        q.set_gene_partition(0, 1)
        q.reset()
        if mate is None:
            q.process_reads(part.param, rb.slice(lo, hi))
        else:
            q.process_paired_reads(part.param, rb.slice(lo, hi), mate.slice(lo, hi))
        if world > 1:
            q.counts_sync()
        res = {"part": part.name, "reads": int(n)}
        if rank == 0:
            st = q.stats()
            # ---- oracle (the cpu_baseline run is reused when it covered exactly this sample)
            pre = oracle_runs.get(pi)
            if pre is not None and pre["nsample"] == n and part.parity_reads is None:
                o, ores, ost = pre["oracle"], pre["res"], pre["stats"]
                # that run already took the oracle through assign_ambig!; the count stores of the alignment stage are compared
                # against a second, alignment-only oracle pass
                o_align = Oracle(idx)
                oracle_process(o_align, part.param, rb, mate)
            else:
                o = Oracle(idx)
                oracle_process(o, part.param, rb, mate)
                o_align, ost, ores = o, o.stats(), None
            rl = part.readlen if part.readlen is not None else int(np.floor(ost["mean_readlen"]))
            chk = {}
            same_tot = (st.total, st.mapped, st.multi, st.sum_readlen) == (ost["total"], ost["mapped"], ost["multi"], ost["sum_readlen"])
            # the running mean is per handle (per rank); with several ranks only the exchanged totals are comparable
            chk["stats"] = same_tot and st.path_overflow == 0 and (world > 1 or st.mean_readlen == ost["mean_readlen"])
            node, ek, ec, ck, cc = q.counts()
            oek, oev = o_align.edges()
            gk, gv = np.concatenate([ek, ck]), np.concatenate([ec, cc]).astype(float)
            order = np.argsort(gk, kind="stable")
            chk["counts"] = (np.array_equal(node.astype(float), o_align.node_counts()) and np.array_equal(gk[order], oek) and np.array_equal(gv[order], oev)
                             and q.keyed(0) == {k: int(v) for k, v in o_align.keyed(0).items()}
                             and q.keyed(1) == {k: int(v) for k, v in o_align.keyed(1).items()})
            if ores is None:
                ores = dict(readlen=rl, em=o.em_tpm(rl, 1, 10000))
                o.assign_ambig()
                ores["events"] = o.events_flat(rl)
            it, tpm, cnt, gt = q.gene_em(rl, 1, 10000)
            oit, otpm, ocnt, ogt = ores["em"]
            chk["tpm"] = it == oit and bits_equal(tpm, otpm) and bits_equal(cnt, ocnt) and bits_equal(gt, ogt)
            gnode, gek, gev, gck, gcv = q.assign_ambig()
            oek2, oev2 = o.edges()
            gk2, gv2 = np.concatenate([gek, gck]), np.concatenate([gev, gcv])
            order = np.argsort(gk2, kind="stable")
            chk["assign_ambig"] = bits_equal(gnode, o.node_counts()) and np.array_equal(gk2[order], oek2) and bits_equal(gv2[order], oev2)
            got = canonical_events(*q.process_events_raw(rl))
            want = ores["events"]
            chk["events"] = all(bits_equal(x, y) for x, y in zip(got, want))
            res.update({k: ("equal" if v else "DIFFERENT") for k, v in chk.items()})
            res.update(em_iterations=int(it), event_records=int(len(want[0])), mapped=int(st.mapped), readlen=int(rl))
            if part.note:
                res["note"] = part.note
            ok = ok and all(chk.values())
        out.append(res)
    flag = [ok]
    if world > 1:
        dist.broadcast_object_list(flag, src=0)
    return out, flag[0]


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    import torch
    import torch.distributed as dist
    import ctypes as C
    import whippet_jl_b200 as wq
    from whippet_jl_b200 import abi, lib

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    parts = build_parts(a, rank, world)
    stream = torch.cuda.current_stream()
    handles, runs = {}, []
    # The product handles come from the library's native on-disk index (wq_index_write once, by rank 0, host only; wq_index_load on every
    # rank: map + copy, no re-layout), so the in-run parity check also covers the file path.  index_create_s (rank 0, N = 1) is the same
    # handle made from the flat GraphLib arrays (wq_index_create: re-layout on the host), for comparison.
    import tempfile
    idx_dir = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    idx_files, index_write_s, index_create_s = {}, 0.0, None
    for p in parts:
        if p.idx_key not in idx_files:
            idx_files[p.idx_key] = os.path.join(idx_dir, f"wq_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}_{p.idx_key.replace('+', '_')}.wqx")
            if rank == 0:
                t0 = time.perf_counter()
                wq.write_index(p.idx, idx_files[p.idx_key])
                index_write_s += time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    if world == 1 and not a.no_parity:
        t0 = time.perf_counter()
        for key in idx_files:
            qq = wq.GraphLibQuant(next(p.idx for p in parts if p.idx_key == key), device=local)
            qq.close()
        index_create_s = time.perf_counter() - t0
    index_load_s = 0.0
    for p in parts:
        if p.idx_key not in handles:
            t0 = time.perf_counter()
            q = wq.GraphLibQuant(p.idx, device=local, index_file=idx_files[p.idx_key])
            index_load_s += time.perf_counter() - t0
            lib.check(q.L.wq_set_stream(q.h, C.c_void_p(stream.cuda_stream)))
            if world > 1:
                q.comm_init(dist)               # NCCL communicator inside the library (unique id broadcast over the process group)
            handles[p.idx_key] = q
        runs.append(PartRun(p, handles[p.idx_key], dev, torch))
    index_file_bytes = sum(os.path.getsize(f) for f in idx_files.values())
    if world > 1:
        dist.barrier()
    if rank == 0:
        for f in idx_files.values():
            if os.path.exists(f):
                os.remove(f)
    phase = {"align": 0.0, "exchange": 0.0, "quant": 0.0, "em_iterations": 0}

    def set_partition():
        if world > 1:
            for q in handles.values():
                q.set_gene_partition(rank, world)   # classes / events of gene range `rank` of `world`; the ranks' outputs concatenate

    def step(host=False):
        last = None
        for r in runs:
            t0 = time.perf_counter()
            r.align(host)
            phase["align"] += time.perf_counter() - t0
            last = quantify(r.q, r.part, world, phase)
        return last

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    set_partition()
    for _ in range(a.warmup):
        step()
    for k in ("align", "exchange", "quant"):
        phase[k] = 0.0
    kms = np.zeros(3)
    qms = np.zeros(2)
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = sum(q.L.wq_launch_count(q.h) for q in handles.values())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    main_run = max(runs, key=lambda r: r.part.n)
    qs = {}
    for _ in range(a.steps):
        step()
        qs = main_run.q.quant_kernel_stats()
        if len(runs) == 1:
            kms += np.array(main_run.q.last_kernel_ms())
            qms += np.array([qs["em_tpm_ms"], qs["em_psi_ms"]])
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1) / a.steps
    launches = int(sum(q.L.wq_launch_count(q.h) for q in handles.values()) - l0)
    phase_ms = {"align": 1000 * phase["align"] / a.steps, "exchange": 1000 * phase["exchange"] / a.steps,
                "quant": 1000 * phase["quant"] / a.steps, "em_iterations": int(phase["em_iterations"]),
                "events": int(phase.get("events", 0))}
    clocks = sampler.stop()
    kms /= a.steps
    qms /= a.steps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    reads_step = sum(p.n for p in parts)
    tot = torch.tensor([float(reads_step)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)
    total_reads = float(tot.item())
    value = total_reads / (ms / 1000.0)
    stats = main_run.q.stats()

    # ---- end to end: pinned host buffers through wq_align_se / wq_align_pe, results read back ---------------
    h2d = sum(r.h2d for r in runs)
    d2h = sum(C.sizeof(abi.MapStatsC) + 8 * (2 * r.part.idx.n_isoforms + r.part.idx.n_genes) for r in runs)    # stats + tpm/count/genetpm
    step(host=True)
    sync_all()
    t0 = time.perf_counter()
    nrep = max(1, min(a.steps, 3))
    for _ in range(nrep):
        step(host=True)
        for r in runs:
            r.q.stats()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1000 / nrep
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_reads / (float(t.item()) / 1000.0)

    # ---- file -> result: the closest equal of the reference's process_reads!, which parses (src/reads.jl:54-57) ----------
    e2e_fastq = None
    if world == 1 and a.config == 2 and not a.no_fastq:
        e2e_fastq = fastq_e2e(main_run, world)

    # ---- CPU baseline (N = 1 only) and the parity check (every run) -------------------------------------------
    oracle_runs = {}
    cpu = None
    if world == 1 and rank == 0:
        per_part = max(20_000, a.cpu_sample // len(parts)) if len(parts) > 1 else a.cpu_sample
        t_full = 0.0
        for pi, p in enumerate(parts):
            r = cpu_oracle_run(p, per_part, 1)
            oracle_runs[pi] = r
            t_full += p.n / r["rate"]
        cpu = {"value": reads_step / t_full, "unit": "reads/s", "cores": 1, "kind": "port",
               "sample": "first %s reads of every part (%d part(s)), oracle restatement on 1 host thread: align+count %.1f s, then EM + multi-map "
                         "assignment + events %.1f s; value = full-workload rate extrapolated (alignment scaled to the part's reads + per-job quant time)"
                         % (per_part, len(parts), sum(r["t_align"] for r in oracle_runs.values()), sum(r["dt"] - r["t_align"] for r in oracle_runs.values()))}
    parity, parity_ok = (None, True)
    if not a.no_parity:
        parity, parity_ok = parity_check(a, parts, handles, world, rank, dist, oracle_runs)
        set_partition()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if not parity_ok:
            raise SystemExit(3)
        return
    mp = main_run.part
    out = {
        "metric": "aligned+quantified reads/sec", "value": value, "unit": "reads/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if a.config == 4 else "weak",
        "vs_baseline": None, "dtype": "u8/u32 integer + f32 scores", "data": "synthetic",
        "config": {"workload": workload_name(a, parts), "config": a.config, "parts": [p.name for p in parts],
                   "l2": "inputs (%.2f GB per step) are larger than the 126 MB L2" % (h2d / 1e9) if h2d > 3e8 else
                         "inputs are %.1f MB per step (smaller than L2); every step rewrites the count tables, hit lists and path pool (> L2) between passes" % (h2d / 1e6),
                   "text_len": int(mp.idx.text_len), "nodes": int(mp.idx.n_nodes), "isoforms": int(mp.idx.n_isoforms), "genes": int(mp.idx.n_genes),
                   "mapped_frac": stats.mapped / max(1, stats.total), "multi_frac": stats.multi / max(1, stats.total),
                   "index_load_s": index_load_s, "index_load": "wq_index_load of the native index file (written by wq_index_write in "
                   f"{index_write_s:.2f} s, {index_file_bytes} bytes)", "index_create_s": index_create_s,
                   "parallelism": f"reads sharded over {world} GPU(s), index replicated" + (", dense counts all-reduced and sparse stores all-gathered over NCCL inside the library (wq_counts_sync), class building partitioned by gene (shares all-gathered), event stage by cost-balanced ranges of the event work list (genes, and node ranges of genes with many nodes), transcript EM replicated" if world > 1 else "")},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "host_batch_layout": sorted(set(x for r in runs for x in r.layout))},
        "kernel_ms": {"seed": float(kms[0]), "extend": float(kms[1]), "resolve": float(kms[2]),
                      "em_tpm": float(qms[0]), "em_psi_span": float(qms[1])},
        "phase_ms": phase_ms,
        "parity_check": parity if parity is not None else "skipped (--no-parity)",
    }
    if e2e_fastq is not None:
        out["e2e_fastq"] = e2e_fastq

    def roofline(ctr, n0):
        R = mp.readlen or 101
        per_kernel, model = byte_model(ctr, n0, R, mp.mate is not None, int(os.environ.get("WQ_FTAB_K", 10)), mp.param.seed_length)
        # per-launch algorithmic bytes of every kernel of the step: alignment kernels from the
        # oracle-instrumented per-read counters, EM kernels from SURVEY.md section 8d's per-iteration figure
        alg = {"wq_k_seed": per_kernel["seed"] * mp.n, "wq_k_extend": per_kernel["extend"] * mp.n,
               "wq_k_resolve": per_kernel["resolve"] * mp.n,
               "wq_k_em_tpm": qs["em_iterations"] * (12 * qs["class_members"] + 8 * qs["classes"] + 24 * qs["isoforms"])}
        kt = {"wq_k_seed": float(kms[0]), "wq_k_extend": float(kms[1]), "wq_k_resolve": float(kms[2]), "wq_k_em_tpm": float(qms[0])}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        chunks = max(1, -(-mp.n // int(os.environ.get("WQ_CHUNK_READS", 1 << 21))))      # the alignment kernels run once per chunk of reads
        per = {k: {"ms": kt[k], "launches_per_step": (1 if k == "wq_k_em_tpm" else chunks), "algorithmic_bytes": float(alg[k]),
                   "achieved": (alg[k] / (kt[k] / 1000.0) / 1e9 if kt[k] > 0 else 0.0)} for k in kt}
        for k in per:
            per[k]["frac"] = per[k]["achieved"] / peak
        dom = max(kt, key=lambda k: kt[k])           # the longest SINGLE kernel of the step
        traffic = None
        try:    # dram__bytes_read + write per launch of that kernel from the committed ncu --set full capture of this round, if any
            # (captured at `reads_per_alignment_launch` reads; the alignment kernels' traffic grows with the reads of the launch)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
            k = tr["kernels"].get(dom) or tr["kernels"].get(dom + "_smem")
            if k and a.config == 2 and mp.mate is None:
                scale = mp.n / float(tr["reads_per_alignment_launch"]) if dom in ("wq_k_seed", "wq_k_extend", "wq_k_resolve") else 1.0
                traffic = (k["dram_bytes_read"] + k["dram_bytes_write"]) * scale
        except Exception:
            pass
        align_ms = float(kms.sum())
        align_bytes = alg["wq_k_seed"] + alg["wq_k_extend"] + alg["wq_k_resolve"]
        return {"bound": "hbm", "kernel": dom, "achieved": per[dom]["achieved"], "peak": peak, "unit": "GB/s",
                "frac": per[dom]["frac"], "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s",
                "kernels": per,
                "psi_em_span": {"ms": float(qms[1]), "algorithmic_bytes": float(qs["psi_bytes"]),
                                "note": "span of the concurrent PSI EM / CI launches on their side streams (several kernels, overlapped with the transcript EM): not a single kernel"},
                "alignment_path": {"ms": align_ms, "achieved": align_bytes / (align_ms / 1000.0) / 1e9 if align_ms > 0 else 0.0,
                                   "frac": (align_bytes / (align_ms / 1000.0) / 1e9 / peak) if align_ms > 0 else 0.0},
                "algorithmic_bytes_per_read": per_kernel, "model": model,
                "note": "dominant kernel = the kernel with the longest device time inside one step (rank 0; CUDA events on the launching stream; ms and "
                        "algorithmic_bytes are per step, i.e. summed over the kernel's launches_per_step launches, so achieved = bytes per launch / average "
                        "launch duration); the PSI EM launches run concurrently on side streams and are reported as a span; every kernel here is latency-bound (random access into L2-resident tables, or a chain of grid barriers), so fractions "
                        "of the HBM roofline are small by construction"}

    if len(runs) == 1:
        if world == 1:
            r = oracle_runs[0]
            out["roofline"] = roofline(r["ctr"], r["n0"])
        else:
            # N > 1: no CPU baseline (reported at N = 1 only); the byte model of the alignment kernels comes from the oracle's
            # counters over a small alignment-only sample of rank 0's reads
            try:
                from oracle import Oracle, counters
                n0 = min(200_000, mp.n)
                counters(reset=True)
                oracle_process(Oracle(mp.idx), mp.param, mp.reads.slice(0, n0), mp.mate.slice(0, n0) if mp.mate is not None else None)
                out["roofline"] = roofline(counters(reset=True), n0)
            except Exception as e:  # noqa: BLE001  (the bench line must not be lost over the explanatory block)
                out["roofline_error"] = repr(e)
    else:
        out["roofline"] = {"bound": "hbm", "kernel": None, "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                           "note": "multi-part config: the per-kernel roofline is reported on the single-part configs (2, 3, 4)"}
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not parity_ok:
        print("[bench] PARITY CHECK FAILED", file=sys.stderr)
        raise SystemExit(3)


if __name__ == "__main__":
    main()
