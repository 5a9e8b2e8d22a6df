"""N > 1 host logic on CPU: two gloo ranks shard the chain range, run their shard through the CPU
oracle and all-reduce the film -- the result must equal the single-process job (chains are
independent units; the only collective is the film sum)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, SCENES


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def shard(total, world, rank):
    """contiguous chain-id ranges: GPU g of G gets [g N/G, (g+1) N/G)"""
    base = rank * (total // world)
    n = total // world if rank < world - 1 else total - base
    return base, n


def worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle()
    h = o.load(os.path.join(SCENES, "torus", "lmc.xml"))
    o.set_option(h, "maxdepth", 4)
    total, steps = 192, 20
    payload = torch.zeros(total + 1)
    if rank == 0:
        norm, init_ls = o.mlt_init(h, 30000, total, 32)
        payload[0] = norm
        payload[1:] = torch.from_numpy(init_ls)
    dist.broadcast(payload, 0)
    norm, init_ls = float(payload[0]), payload[1:].numpy().copy()
    base, n = shard(total, world, rank)
    film, trace, a, stats = o.run_chains(h, n, steps, norm, init_ls, chain_base=base, total_chains=total,
                                         samples_per_chain=steps, threads=2)
    film_t = torch.from_numpy(film)
    dist.all_reduce(film_t)
    st = torch.from_numpy(stats.astype(np.int64))
    dist.all_reduce(st)
    if rank == 0:
        np.save(os.path.join(out_dir, "film.npy"), film_t.numpy())
        np.save(os.path.join(out_dir, "stats.npy"), st.numpy())
        np.save(os.path.join(out_dir, "init.npy"), np.concatenate([[norm], init_ls]).astype(np.float32))
    np.save(os.path.join(out_dir, "trace%d.npy" % rank), trace)
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(oracle, tmp_path):
    world = 2
    mp.spawn(worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    film = np.load(tmp_path / "film.npy")
    stats = np.load(tmp_path / "stats.npy")
    init = np.load(tmp_path / "init.npy")
    h = oracle.load(os.path.join(SCENES, "torus", "lmc.xml"))
    oracle.set_option(h, "maxdepth", 4)
    f1, t1, a1, s1 = oracle.run_chains(h, 192, 20, float(init[0]), init[1:].copy(), samples_per_chain=20, threads=2)
    t = np.concatenate([np.load(tmp_path / "trace0.npy"), np.load(tmp_path / "trace1.npy")], axis=0)
    assert np.array_equal(t, t1)
    assert np.array_equal(stats, s1.astype(np.int64))
    assert np.allclose(film, f1, rtol=1e-5, atol=1e-6)


def test_shard_ranges_cover_exactly():
    for total in (1, 7, 1 << 20, (1 << 23) + 3):
        for world in (1, 2, 4, 8):
            seen = 0
            for r in range(world):
                base, n = shard(total, world, r)
                assert base == seen
                seen += n
            assert seen == total
