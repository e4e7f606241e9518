"""N>1 path on CPU: world_size-2 gloo processes exercise the shard planner and the gather that the multi-GPU
drivers use (moleculekit_b200/sharding.py).  The compute inside each shard is the CPU ORACLE here -- test
infrastructure standing in for the CUDA kernels, which cannot run without a GPU; the product wrappers
(voxelize_sharded / project_sharded) call exactly these helpers around the CUDA path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_properties():
    from moleculekit_b200.sharding import balanced_partition, partition

    for n in (0, 1, 7, 256, 1000):
        for w in (1, 2, 3, 8):
            off = partition(n, w)
            assert off[0] == 0 and off[-1] == n and np.all(np.diff(off) >= 0)
            assert np.diff(off).max() - np.diff(off).min() <= 1
    costs = np.array([1, 1, 1, 1, 10, 1, 1, 1, 1, 10], dtype=float)
    off = balanced_partition(costs, 2)
    assert off.tolist()[0] == 0 and off[-1] == 10 and abs(costs[:off[1]].sum() - costs[off[1]:].sum()) <= 10
    assert balanced_partition([], 4).tolist() == [0, 0, 0, 0, 0]
    assert balanced_partition([0, 0, 0], 2).tolist() == [0, 2, 3]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from moleculekit_b200 import workloads
        from moleculekit_b200.molecule_lite import MolLite
        from moleculekit_b200.sharding import gather_rows, project_sharded, run_sharded
        from moleculekit_b200.tools.voxeldescriptors import getCenters
        from oracle import cpu_oracle

        # ragged gather
        local = torch.full((3 + 2 * rank, 4), float(rank))
        full = gather_rows(local, [3, 5])
        assert full.shape == (8, 4) and torch.all(full[:3] == 0) and torch.all(full[3:] == 1)

        # occupancy batch sharded over the two ranks, oracle compute per shard, gathered == single process
        w = workloads.protein_pockets(B=5, n_atoms=60, box=8.0, radius=3.0, seed=3)

        def voxelise(b, e):
            outs = []
            for i in range(b, e):
                c, _ = getCenters(boxsize=w["boxsize"], center=w["centers"][i], voxelsize=1.0)
                o = np.zeros((c.shape[0], 8)); cpu_oracle.calculate_occupancy(c, w["coords"][i], w["sigmas"][i], o)
                outs.append(o.astype(np.float32))
            return torch.from_numpy(np.concatenate(outs)) if outs else torch.zeros((0, 8))

        loc, (b, e) = run_sharded(5, voxelise)
        assert (b, e) == ((0, 3) if rank == 0 else (3, 5)) and loc.shape == ((e - b) * 512, 8)
        full, _ = run_sharded(5, voxelise, rows_per_item=512, gather=True)
        assert torch.equal(full, voxelise(0, 5))

        # frame-sharded projection: a CPU stand-in projection built on the oracle, through project_sharded
        rng = np.random.default_rng(0)
        mol = MolLite((rng.normal(size=(12, 3, 7)) * 5).astype(np.float32), box=np.full((3, 7), 9.0, np.float32))

        class OracleDistance:
            def project(self, m):
                s1 = np.arange(0, 5, dtype=np.uint32); s2 = np.arange(5, 12, dtype=np.uint32)
                ch = np.zeros(12, np.uint32); ch[s2] = 1
                r = np.zeros((m.numFrames, 35), np.float32)
                cpu_oracle.dist_trajectory(m.coords, m.box, s1, s2, ch, False, True, r)
                return r

        res = project_sharded(OracleDistance(), mol, gather=True)
        assert res.shape == (7, 35) and np.array_equal(res, OracleDistance().project(mol))
        part, (f0, f1) = project_sharded(OracleDistance(), mol, gather=False)
        assert (f0, f1) == ((0, 4) if rank == 0 else (4, 7)) and np.array_equal(part, res[f0:f1])
        # frame-sharded wrapping (orthorhombic and triclinic stand-ins on the oracle) and hydrogen bonds
        from moleculekit_b200.sharding import hbonds_sharded, wrap_sharded

        groups = np.arange(0, 13, 3, dtype=np.uint32)
        L = 9.0
        bv = np.repeat(np.array([[L, 0, 0], [0, L, 0], [L / 2, L / 2, L * 2 ** 0.5 / 2]])[:, :, None], 7, axis=2)

        def oracle_wrap(m, mode=None):
            if mode is None:
                cpu_oracle.wrap_box(groups, m.coords, m.box, np.arange(3, dtype=np.uint32), np.zeros(3, np.float32))
            else:
                cpu_oracle.wrap_compact_unitcell(groups, m.coords, m.boxvectors, np.arange(3, dtype=np.uint32),
                                                 np.zeros(3, np.float32), mode)

        for mode in (None, 1):
            want = MolLite(mol.coords.copy(), box=mol.box.copy()); want.boxvectors = bv
            oracle_wrap(want, mode)
            got = MolLite(mol.coords.copy(), box=mol.box.copy()); got.boxvectors = bv
            assert wrap_sharded(got, wrap_fn=oracle_wrap, mode=mode) == (0, 7)
            assert np.array_equal(got.coords, want.coords), mode
            part = MolLite(mol.coords.copy(), box=mol.box.copy()); part.boxvectors = bv
            f0, f1 = wrap_sharded(part, gather=False, wrap_fn=oracle_wrap, mode=mode)
            assert (f0, f1) == ((0, 4) if rank == 0 else (4, 7))
            assert np.array_equal(part.coords[:, :, f0:f1], want.coords[:, :, f0:f1])
            other = np.ones(7, bool); other[f0:f1] = False
            assert np.array_equal(part.coords[:, :, other], mol.coords[:, :, other])

        don = np.array([[0, 1], [3, 4], [6, 7]], dtype=np.uint32); acc = np.array([2, 5, 8, 11], dtype=np.uint32)
        ones = np.ones(12, np.uint32)

        def oracle_hb(m, donors, acceptors, sel1, sel2, **kw):
            r = cpu_oracle.hbonds_calculate(donors, acceptors, m.coords, m.box, ones, ones, 6.0, 20.0, True, False)
            return [np.array(x, dtype=np.int64).reshape(-1, 3) for x in r]

        want = oracle_hb(mol, don, acc, None, None)
        got = hbonds_sharded(mol, don, acc, hbonds_fn=oracle_hb)
        assert len(got) == 7 and all(np.array_equal(a, b) for a, b in zip(got, want)) and sum(len(x) for x in want) > 0
        loc, (f0, f1) = hbonds_sharded(mol, don, acc, gather=False, hbonds_fn=oracle_hb)
        assert len(loc) == f1 - f0 and all(np.array_equal(a, b) for a, b in zip(loc, want[f0:f1]))
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gloo_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
