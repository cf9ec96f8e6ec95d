"""Host-side orderings (CPU): the k-d partition that freezes the body clusters / orders the scene (spatial.kd_order) and
the dealing of scene blocks to the ranks of a sharded fit (fit._deal_blocks).  No reference counterpart (the reference
is a dense single-GPU bmm); these pin the properties the search kernels rely on."""
import importlib

import numpy as np
import pytest
import torch

spatial = importlib.import_module("4dcapture-fpv_b200.spatial")
fit = importlib.import_module("4dcapture-fpv_b200.fit")
sharded = importlib.import_module("4dcapture-fpv_b200.sharded")


def _cloud(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, generator=g) * torch.tensor([8.0, 8.0, 3.0]) - torch.tensor([4.0, 4.0, 0.0])


@pytest.mark.parametrize("n,leaf,align", [(1000, 16, "pow4"), (4096, 16, "pow4"), (5000, 64, "pow2"), (777, 16, "pow2")])
def test_kd_order_is_a_deterministic_permutation(n, leaf, align):
    pts = _cloud(n, n)
    perm = spatial.kd_order(pts, leaf=leaf, align=align)
    assert perm.dtype == torch.int64 and perm.shape == (n,)
    assert torch.equal(torch.sort(perm).values, torch.arange(n))
    assert torch.equal(perm, spatial.kd_order(pts.clone(), leaf=leaf, align=align))


def test_kd_order_pow2_aligned_blocks_are_cells():
    """align='pow2': every aligned block of leaf * 2^k positions is a k-d cell -- two sibling blocks are separated by a
    plane along one axis (what makes the scene's tiles of 64 / groups of 128 / super-tiles of 2048 compact)."""
    n, leaf = 4096, 64
    pts = _cloud(n, 3)
    s = pts[spatial.kd_order(pts, leaf=leaf, align="pow2")].numpy()
    blk = n // 2
    while blk >= leaf:
        for lo in range(0, n, 2 * blk):
            left, right = s[lo:lo + blk], s[lo + blk:lo + 2 * blk]
            assert any(left[:, ax].max() <= right[:, ax].min() for ax in range(3)), (blk, lo)
        blk //= 2


def test_kd_order_clusters_are_compact():
    """The reason the ordering exists: 16-point clusters of a k-d partition are compact in all three axes (the comparison
    with a Morton curve -- less than half the radius on the body, DESIGN 4.0 -- needs the device and lives in
    profiles/r02_sphere_order_tune.txt)."""
    pts = _cloud(8192, 5)

    def mean_radius(order):
        c = pts[order].reshape(-1, 16, 3)
        return float((c - c.mean(1, keepdim=True)).norm(dim=2).max(1).values.mean())

    kd = spatial.kd_order(pts, leaf=16, align="pow4")
    r_kd = mean_radius(kd)
    r_rand = mean_radius(torch.randperm(8192, generator=torch.Generator().manual_seed(1)))
    assert r_kd < 0.25 * r_rand
    # a cell of 16 of 8192 uniform points in an 8 x 8 x 3 m box has a volume of 0.375 m^3: a radius of about half a metre
    assert r_kd < 0.75


def test_kd_order_puts_non_finite_points_last():
    pts = _cloud(300, 7)
    pts[17, 1] = float("nan")
    pts[250, 0] = float("inf")
    perm = spatial.kd_order(pts, leaf=16)
    assert set(perm[-2:].tolist()) == {17, 250}
    assert torch.equal(torch.sort(perm).values, torch.arange(300))


@pytest.mark.parametrize("M,world,blk", [(1000, 2, 128), (100000, 8, 128), (4097, 4, 2048), (130, 8, 128), (64, 3, 128)])
def test_deal_blocks_is_an_exact_block_aligned_partition(M, world, blk):
    perm = fit._deal_blocks(M, world, blk)
    assert torch.equal(torch.sort(perm).values, torch.arange(M))
    whole = 0
    for r in range(world):
        b, e = sharded.shard_range(M, world, r)
        share = perm[b:e]
        # the share starts with whole, aligned blocks dealt round-robin (block k goes to rank k % world) ...
        k = 0
        while (k + 1) * blk <= share.numel():
            first = int(share[k * blk])
            if first % blk != 0 or not torch.equal(share[k * blk:(k + 1) * blk], torch.arange(first, first + blk)):
                break
            assert (first // blk) % world == r
            k += 1
        whole += k
        # ... and ends with at most a tail handed over from longer shares
        assert share.numel() - k * blk < 2 * blk + world
    assert whole >= M // blk - world
