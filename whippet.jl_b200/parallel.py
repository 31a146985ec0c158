"""Multi-GPU exchange step: reads are sharded over ranks with the index
replicated; after alignment the dense per-node counts and the scalar statistics are combined with ONE
all-reduce over the device-resident vector (NCCL over NVLink on GPUs), and the sparse keyed stores
(edges, circular edges, long paths, multi-map classes) with one all-gather of packed records that each
rank merges in canonical key order.  The reference has no counterpart (it is single-threaded,
bin/whippet-quant.jl:137-138)."""
from __future__ import annotations

import numpy as np


def pack_sparse(edge_key, edge_count, circ_key, circ_count, longs: dict, multis: dict) -> np.ndarray:
    """One u64 buffer: [n_edge, n_circ, n_long, n_multi, long_words, multi_words,
    edge keys.., edge counts.., circ keys.., circ counts.., per long: nwords,count,words.., per multi: same]."""
    parts = [np.array([len(edge_key), len(circ_key), len(longs), len(multis),
                       sum(len(k) for k in longs), sum(len(k) for k in multis)], np.uint64),
             np.asarray(edge_key, np.uint64), np.asarray(edge_count, np.uint64),
             np.asarray(circ_key, np.uint64), np.asarray(circ_count, np.uint64)]
    for d in (longs, multis):
        for k in sorted(d):
            parts.append(np.array([len(k), d[k]], np.uint64))
            parts.append(np.asarray(k, np.uint64))
    return np.concatenate(parts) if parts else np.zeros(6, np.uint64)


def unpack_sparse(buf: np.ndarray):
    ne, nc, nl, nm = (int(x) for x in buf[:4])
    p = 6
    ek, ec = buf[p:p + ne], buf[p + ne:p + 2 * ne]
    p += 2 * ne
    ck, cc = buf[p:p + nc], buf[p + nc:p + 2 * nc]
    p += 2 * nc
    out = []
    for n in (nl, nm):
        d = {}
        for _ in range(n):
            nw, cnt = int(buf[p]), int(buf[p + 1])
            d[tuple(int(x) for x in buf[p + 2:p + 2 + nw])] = cnt
            p += 2 + nw
        out.append(d)
    return ek, ec, ck, cc, out[0], out[1]


def allgather_buffers(buf: np.ndarray, device=None):
    """All-gather variable-length u64 buffers (as int64 tensors on `device`, or CPU for gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([len(buf)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    mine = torch.zeros(m, dtype=torch.int64, device=dev)
    mine[:len(buf)] = torch.from_numpy(buf.view(np.int64)).to(dev)
    outs = [torch.zeros(m, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(outs, mine)
    return [o[:s].cpu().numpy().view(np.uint64) for o, s in zip(outs, sizes)]


def merge_sparse(parts):
    """Sum a list of unpacked sparse stores (used by the CPU test and as the reference for the merge)."""
    edges, circs, longs, multis = {}, {}, {}, {}
    for ek, ec, ck, cc, lo, mu in parts:
        for k, v in zip(ek.tolist(), ec.tolist()):
            edges[k] = edges.get(k, 0) + v
        for k, v in zip(ck.tolist(), cc.tolist()):
            circs[k] = circs.get(k, 0) + v
        for k, v in lo.items():
            longs[k] = longs.get(k, 0) + v
        for k, v in mu.items():
            multis[k] = multis.get(k, 0) + v
    return edges, circs, longs, multis
