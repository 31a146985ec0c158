"""N>1 exchange path on CPU: world_size-2 gloo processes exchange packed sparse count stores and a dense
vector, and every rank ends with the union-sum."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _stores(rank):
    rng = np.random.default_rng(100 + rank)
    ek = np.unique(rng.integers(1, 2000, 300).astype(np.uint64) << np.uint64(16) | rng.integers(1, 50, 300).astype(np.uint64))
    ec = rng.integers(1, 100, len(ek)).astype(np.uint64)
    ck = np.unique(rng.integers(1, 50, 20).astype(np.uint64))
    cc = rng.integers(1, 10, len(ck)).astype(np.uint64)
    longs = {(3, int(g), 1, 2, 3 + int(g) % 3): int(rng.integers(1, 9)) for g in rng.integers(1, 60, 40)}
    multis = {((1 << 28) | 2, 1, int(g), 1, 1, int(g) + 1, 2): int(rng.integers(1, 9)) for g in rng.integers(1, 60, 30)}
    dense = rng.integers(0, 1000, 5000).astype(np.int64)
    return ek, ec, ck, cc, longs, multis, dense


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT]
    import torch
    import torch.distributed as dist
    from whippet_jl_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ek, ec, ck, cc, longs, multis, dense = _stores(rank)
    t = torch.from_numpy(dense.copy())
    dist.all_reduce(t)
    bufs = parallel.allgather_buffers(parallel.pack_sparse(ek, ec, ck, cc, longs, multis))
    merged = parallel.merge_sparse([parallel.unpack_sparse(b) for b in bufs])
    q.put((rank, t.numpy().copy(), merged))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    from whippet_jl_b200 import parallel
    parts = [_stores(r) for r in range(world)]
    want_dense = sum(p[6] for p in parts)
    want = parallel.merge_sparse([p[:6] for p in parts])
    for rank, dense, merged in res:
        assert np.array_equal(dense, want_dense)
        assert merged == want


def test_pack_roundtrip():
    from whippet_jl_b200 import parallel
    ek, ec, ck, cc, longs, multis, _ = _stores(0)
    out = parallel.unpack_sparse(parallel.pack_sparse(ek, ec, ck, cc, longs, multis))
    assert np.array_equal(out[0], ek) and np.array_equal(out[1], ec) and np.array_equal(out[2], ck) and np.array_equal(out[3], cc)
    assert out[4] == longs and out[5] == multis
    empty = parallel.unpack_sparse(parallel.pack_sparse([], [], [], [], {}, {}))
    assert len(empty[0]) == 0 and empty[4] == {} and empty[5] == {}


# ---- class building partitioned by gene: the shares of two gloo ranks assemble to the unpartitioned class set ----
def _class_counts():
    """Deterministic count stores from the oracle (every rank computes the same ones: after the count exchange every rank
    holds identical stores)."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import whippet_jl_b200 as wq
    from wq_inputs import synth
    from oracle import Oracle
    idx = synth.make_index(n_genes=400, K=9, seed=81, paralog_frac=0.1)
    rb = synth.simulate_reads(idx, 60000, 101, seed=82, sub_rate=0.01)
    o = Oracle(idx)
    o.process_reads(wq.AlignParam(), rb)
    ek, ev = o.edges()
    return idx, dict(node=o.node_counts().astype("<u8"), ekey=ek, ecount=ev.astype("<u8"), longs=o.keyed(0), multis=o.keyed(1))


def _class_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx, counts = _class_counts()
    from emu.host_emu import HostEmu
    e = HostEmu(idx)
    e.set_counts(**counts)
    e.build_classes(rank, world)                          # this rank's share (multi-map key range + gene range)
    mine = np.frombuffer(e.share(), dtype=np.uint8)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(mine)], dtype=torch.int64))
    m = max(int(s.item()) for s in sizes)
    padded = torch.zeros(m, dtype=torch.uint8)
    padded[:len(mine)] = torch.from_numpy(mine.copy())
    outs = [torch.zeros(m, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(outs, padded)
    e.assemble([outs[r][:int(sizes[r].item())].numpy().tobytes() for r in range(world)])
    classes, nm = e.classes()
    q.put((rank, classes, nm, e.iso_count()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_class_shares():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_class_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=300) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    idx, counts = _class_counts()
    from emu.host_emu import HostEmu
    full = HostEmu(idx)
    full.set_counts(**counts)
    full.build_classes()
    want, nm = full.classes()
    assert nm > 10 and len(want) > 500
    for rank, classes, gm, iso_count in res:
        assert gm == nm and classes == want, rank
        assert np.array_equal(iso_count, full.iso_count()), rank


# ---- event stage partitioned by cost-balanced ranges of the work list: two gloo ranks, each builds its range, the ranges concatenate ----
def _event_case():
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import whippet_jl_b200 as wq
    from wq_inputs import synth
    from oracle import Oracle
    idx = synth.make_index(n_genes=2200, K=9, seed=91, mean_exons=14.0)          # some genes with 32 nodes or more: cut into node pieces
    rb = synth.simulate_reads(idx, 150000, 101, seed=92, sub_rate=0.01)
    o = Oracle(idx)
    o.process_reads(wq.AlignParam(), rb)
    ek, ev = o.edges()
    return idx, o.node_counts().astype("<f8"), ek, ev.astype("<f8")


def _event_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx, node, ek, ev = _event_case()
    # the ranks agree on the counts the cuts are computed from (on the GPUs: after wq_counts_sync); checked with a checksum
    chk = torch.tensor([float(node.sum()), float(ev.sum()), float(len(ek))], dtype=torch.float64)
    both = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(both, chk)
    assert all(torch.equal(b, chk) for b in both)
    from emu.host_emu import HostEmu
    e = HostEmu(idx)
    recs = e.events_part(node, ek, ev, 101, rank, world)
    q.put((rank, recs))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_event_ranges_concatenate():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_event_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    idx, node, ek, ev = _event_case()
    from emu.host_emu import HostEmu
    full = HostEmu(idx)
    want = full.events_part(node, ek, ev, 101, 0, 1)
    assert int((np.diff(idx.node_base) >= 32).sum()) >= 3 and len(want) > 10000
    assert min(len(res[0]), len(res[1])) > len(want) // 20                       # both ranks have work
    assert res[0] + res[1] == want
