"""CPU, world_size 2, gloo: the host-side logic of the scene-sharded chamfer (sharded.py) --
key packing, the MIN all-reduce, ownership masking in backward, the parameter-gradient all-reduce --
with the CUDA search replaced by the CPU oracle through the `_search` injection point."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _oracle_search(a, b):
    from oracle import chamfer_oracle as co
    d1, d2, i1, i2 = co.dist_chamfer(a.detach().numpy(), b.detach().numpy())
    return torch.tensor(d2), torch.tensor(i2), torch.tensor(d1), torch.tensor(i1)


def _worker(rank, world, init_file, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_pkg
    fpv = load_pkg()
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    T, N, M = 3, 37, 101
    a0 = torch.randint(-3, 4, (T, N, 3), generator=g).float()      # lattice: exact ties across shards
    a0 += 0.25 * torch.rand(T, N, 3, generator=g)
    scene = torch.randint(-3, 4, (M, 3), generator=g).float()
    w1 = torch.rand(T, M, generator=g)
    w2 = torch.rand(T, N, generator=g)
    lo, hi = fpv.shard_range(M, world, rank)
    a = a0.clone().requires_grad_(True)
    d_b2a, d_a2b, i_b2a, i_a2b = fpv.distChamferSharded(a, scene[lo:hi], lo, None, _oracle_search)
    # local loss: replicated term scaled by 1/world, shard-local term over the local columns
    loss = (d_a2b * w2).sum() / world + (d_b2a * w1[:, lo:hi]).sum()
    loss.backward()
    fpv.allreduce_grads([a])
    tot = loss.detach().clone()
    dist.all_reduce(tot)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), d_a2b=d_a2b.detach().numpy(), i_a2b=i_a2b.numpy(),
             d_b2a=d_b2a.detach().numpy(), i_b2a=i_b2a.numpy(), grad=a.grad.numpy(), loss=tot.numpy(),
             a=a0.numpy(), scene=scene.numpy(), w1=w1.numpy(), w2=w2.numpy(), lo=lo, hi=hi)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_chamfer_world2_matches_unsharded():
    from oracle import chamfer_oracle as co
    world = 2
    with tempfile.TemporaryDirectory() as td:
        init_file = os.path.join(td, "init")
        mp.spawn(_worker, args=(world, init_file, td), nprocs=world, join=True)
        r = [dict(np.load(os.path.join(td, f"r{k}.npz"))) for k in range(world)]
    a, scene, w1, w2 = r[0]["a"], r[0]["scene"], r[0]["w1"], r[0]["w2"]
    d1, d2, i1, i2 = co.dist_chamfer(a, scene)
    for k in range(world):
        assert np.array_equal(r[k]["i_a2b"], i2) and np.array_equal(r[k]["d_a2b"], d2)      # combined, global idx
        lo, hi = int(r[k]["lo"]), int(r[k]["hi"])
        assert np.array_equal(r[k]["d_b2a"], d1[:, lo:hi]) and np.array_equal(r[k]["i_b2a"], i1[:, lo:hi])
    ga, _ = co.dist_chamfer_bwd(a, scene, w1, w2, i1, i2)
    np.testing.assert_allclose(r[0]["grad"], ga, rtol=1e-5, atol=1e-5)
    assert np.array_equal(r[0]["grad"], r[1]["grad"])
    np.testing.assert_allclose(r[0]["loss"], (d2 * w2).sum() + (d1 * w1).sum(), rtol=1e-5)


def test_key_pack_unpack_roundtrip_and_order():
    sys.path.insert(0, ROOT)
    from conftest import load_pkg
    fpv = load_pkg()
    from oracle import chamfer_oracle as co
    d = torch.tensor([0.0, 1.5, 1.5, 3.0e-39, float("inf")])
    i = torch.tensor([7, 2, 9, 0, 4294967295 - 1])
    k = fpv.sharded.pack_keys_torch(d, i)
    for j in range(5):
        assert int(k[j]) == co.pack_key(float(d[j]), int(i[j]))
    dd, ii = fpv.sharded.unpack_keys_torch(k)
    assert torch.equal(dd, d) and torch.equal(ii, i)
    assert int(torch.argmin(k)) == 0 and k[1] < k[2] and k[3] < k[1] and k[4] == k.max()
