"""Multi-process (world size 2, gloo, CPU) check of the km-field sharding used by `bench.py --gpus N`.

Each rank interpolates its slice of the field batch with the CPU oracle (this is a test, so the oracle may be
executed) and the slices are all-gathered; the result must equal the single-process call over the whole batch
bit for bit, including ibo — i.e. fields really are independent given the grid pair and no collective is needed
on the data path.
"""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ipcases as ic

pkg = importlib.import_module("nceplibs-ip_b200")


def test_field_slice_partitions():
    for km in (0, 1, 5, 8, 1024, 8192):
        for world in (1, 2, 3, 4, 8):
            parts = [pkg.field_slice(km, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == km
            assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert pkg.weak_slice(1024, 8, 3) == (3072, 4096)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(vector):
    kind = "4"
    gin, gout = ic.g2_input(kind, vector), ic.g2_templates(kind, vector)["218"]
    mi, mo, km = ic.npts_of(gin), ic.npts_of(gout), 5
    rng = np.random.default_rng(3)
    base = ic.inputs()["u" if vector else "snoalb"]
    g = np.stack([base * (1 + 0.05 * k) + rng.standard_normal(mi) for k in range(km)]).astype(np.float32)
    v = np.stack([rng.standard_normal(mi) + k for k in range(km)]).astype(np.float32)
    li = (rng.random((km, mi)) < 0.7).astype(np.uint8)
    ibi = [k % 2 for k in range(km)]
    return kind, gin, gout, mi, mo, km, g, v, li, ibi


def _worker(rank, world, port, vector, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_lib import oracle
    kind, gin, gout, mi, mo, km, g, v, li, ibi = _case(vector)
    k0, k1 = pkg.field_slice(km, world, rank)
    o = oracle(kind)
    if vector:
        r = o.ipolatev(0, [0] * 20, gin, gout, mi, mo, k1 - k0, ibi[k0:k1], li[k0:k1], g[k0:k1], v[k0:k1])
        mine = np.concatenate([r["uo"], r["vo"], r["lo"].astype(np.float32), r["ibo"][:, None].astype(np.float32).repeat(mo, 1)])
    else:
        r = o.ipolates(3, [-1] * 20, gin, gout, mi, mo, k1 - k0, ibi[k0:k1], li[k0:k1], g[k0:k1])
        mine = np.concatenate([r["go"], r["lo"].astype(np.float32), r["ibo"][:, None].astype(np.float32).repeat(mo, 1)])
    parts = [None] * world
    dist.all_gather_object(parts, (k0, k1, mine))
    if rank == 0:
        np.save(out_path, np.array([p[2] for p in parts], dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("vector", [False, True])
def test_two_rank_sharding_equals_single_call(tmp_path, vector):
    world = 2
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), vector, out), nprocs=world, join=True)
    parts = np.load(out, allow_pickle=True)
    from oracle_lib import oracle
    kind, gin, gout, mi, mo, km, g, v, li, ibi = _case(vector)
    o = oracle(kind)
    ncomp = 2 if vector else 1
    if vector:
        r = o.ipolatev(0, [0] * 20, gin, gout, mi, mo, km, ibi, li, g, v)
        comps = [r["uo"], r["vo"]]
    else:
        r = o.ipolates(3, [-1] * 20, gin, gout, mi, mo, km, ibi, li, g)
        comps = [r["go"]]
    for c in range(ncomp + 2):
        got = np.concatenate([p.reshape(ncomp + 2, -1, mo)[c] for p in parts])
        if c < ncomp:
            assert np.array_equal(got.view(np.uint32), comps[c].view(np.uint32))
        elif c == ncomp:
            assert np.array_equal(got.astype(np.uint8), r["lo"])
        else:
            assert np.array_equal(got[:, 0].astype(np.int64), r["ibo"].astype(np.int64))
