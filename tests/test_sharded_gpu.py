"""GPU, >= 2 devices, one rank per GPU: the scene-sharded chamfer over NCCL and the sharded fit step (peer-memory
mailbox over NVLink) vs the single-GPU path.  Skipped on a 1-GPU box, where tests/test_p2p_gpu.py runs the same
multi-rank path with two ranks on one device and tests/test_sharded_gloo.py covers the host logic on CPU."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, init_file, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_pkg
    fpv = load_pkg()
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world, device_id=dev)
    g = torch.Generator().manual_seed(0)
    T, N, M = 4, 3000, 50_001
    a0 = torch.randn(T, N, 3, generator=g)
    scene = torch.randn(M, 3, generator=g)
    scene[M // 2: M // 2 + 500] = scene[:500]                     # exact ties across the shard boundary
    w1, w2 = torch.rand(T, M, generator=g), torch.rand(T, N, generator=g)
    lo, hi = fpv.shard_range(M, world, rank)
    a = a0.clone().to(dev).requires_grad_(True)
    d_b2a, d_a2b, i_b2a, i_a2b = fpv.distChamferSharded(a, scene[lo:hi].to(dev), lo)
    loss = (d_a2b * w2.to(dev)).sum() / world + (d_b2a * w1[:, lo:hi].to(dev)).sum()
    loss.backward()
    fpv.allreduce_grads([a])
    prob = fpv.FitProblem(T=4, M=30_000, device=dev, seed=1236, rank=rank, world_size=world)
    fl = prob.step().clone()                                      # already the GLOBAL loss (summed in the gradient exchange)
    fit_grad, fit_scale = prob.params.grad.clone(), prob.scale.grad.clone()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), d_a2b=d_a2b.detach().cpu().numpy(), i_a2b=i_a2b.cpu().numpy(),
             d_b2a=d_b2a.detach().cpu().numpy(), i_b2a=i_b2a.cpu().numpy(), grad=a.grad.cpu().numpy(),
             a=a0.numpy(), scene=scene.numpy(), w1=w1.numpy(), w2=w2.numpy(), lo=lo, hi=hi,
             fit_loss=fl.cpu().numpy(), fit_grad=fit_grad.cpu().numpy(), fit_scale=fit_scale.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_nccl_matches_single_gpu(fpv, cuda_dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import chamfer_oracle as co
    world = 2
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(world, os.path.join(td, "init"), td), nprocs=world, join=True)
        r = [dict(np.load(os.path.join(td, f"r{k}.npz"))) for k in range(world)]
    a, scene, w1, w2 = r[0]["a"], r[0]["scene"], r[0]["w1"], r[0]["w2"]
    d1, d2, i1, i2 = co.dist_chamfer(a, scene)
    for k in range(world):
        lo, hi = int(r[k]["lo"]), int(r[k]["hi"])
        assert np.array_equal(r[k]["i_a2b"], i2) and np.array_equal(r[k]["d_a2b"], d2)
        assert np.array_equal(r[k]["d_b2a"], d1[:, lo:hi]) and np.array_equal(r[k]["i_b2a"], i1[:, lo:hi])
    ga, _ = co.dist_chamfer_bwd(a, scene, w1, w2, i1, i2)
    np.testing.assert_allclose(r[0]["grad"], ga, rtol=1e-5, atol=1e-5 * np.abs(ga).max())
    assert np.array_equal(r[0]["grad"], r[1]["grad"])
    # the sharded fit step equals the single-GPU fit step
    prob = fpv.FitProblem(T=4, M=30_000, device=cuda_dev, seed=1236)
    loss = prob.step()
    assert float(r[0]["fit_loss"]) == pytest.approx(loss.item(), rel=1e-5)
    g1 = prob.params.grad.cpu().numpy()
    np.testing.assert_allclose(r[0]["fit_grad"], g1, rtol=1e-5, atol=1e-5 * np.abs(g1).max())
    np.testing.assert_allclose(r[0]["fit_scale"], prob.scale.grad.cpu().numpy(), rtol=1e-5)
