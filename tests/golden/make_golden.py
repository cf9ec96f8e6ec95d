"""Generate the golden vectors under tests/golden/ from the REAL reference.

Run in the authoring container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports /root/reference/chamfer_python.py UNMODIFIED and calls its distChamfer.  The only
concession is the one-line CPU shim SURVEY.md section 8c documents: chamfer_python.py:24 builds
its diagonal index with torch.cuda.LongTensor, which needs a driver; the shim aliases that name
to torch.LongTensor so the very same code runs on CPU.  The literal function only accepts
N == M (chamfer_python.py:24-27), so every golden case is square.

Cases
  probe4        the hand-computed 4-point case of SURVEY.md section 8c
  lattice_*     small-integer coordinates with many duplicates: every fp32 form of the squared
                distance is exact, so distances AND indices (ties included) must match bit for bit
  dyadic_*      coordinates k/8: still exact in fp32
  random_*      unit-cube floats: indices must match wherever the NN margin exceeds the fp32 error
                of the reference's expanded form; distances within 1e-5 relative
Each case also stores the reference's autograd gradients for loss = sum(g1*d1) + sum(g2*d2).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference/chamfer_python.py"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    torch.cuda.LongTensor = torch.LongTensor  # CPU shim for chamfer_python.py:24
    spec = importlib.util.spec_from_file_location("chamfer_python_ref", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_case(ref, name, a, b, g1, g2):
    ta = torch.tensor(a, dtype=torch.float32, requires_grad=True)
    tb = torch.tensor(b, dtype=torch.float32, requires_grad=True)
    d1, d2, i1, i2 = ref.distChamfer(ta, tb)
    loss = (d1 * torch.tensor(g1)).sum() + (d2 * torch.tensor(g2)).sum()
    loss.backward()
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        a=a, b=b, g_b2a=g1, g_a2b=g2,
        d_b2a=d1.detach().numpy(), d_a2b=d2.detach().numpy(),
        i_b2a=i1.numpy(), i_a2b=i2.numpy(),
        grad_a=ta.grad.numpy(), grad_b=tb.grad.numpy(),
    )
    print(f"{name}: bs={a.shape[0]} N=M={a.shape[1]}  d_b2a[0,:4]={d1[0,:4].tolist()}")


def main():
    if not os.path.exists(REF):
        sys.exit("reference not present; golden vectors can only be regenerated in the authoring container")
    ref = load_reference()
    rng = np.random.default_rng(20261017)

    a = np.array([[[0, 0, 0], [5, 5, 5], [1, 0, 0], [9, 9, 9]]], np.float32)
    b = np.array([[[1, 0, 0], [1, 0, 0], [-1, 0, 0], [1, 0, 0]]], np.float32)
    run_case(ref, "probe4", a, b, np.ones((1, 4), np.float32), np.ones((1, 4), np.float32))

    def dy(shape):  # dyadic upstream weights keep the reference's gradient arithmetic exact
        return (rng.integers(-8, 9, size=shape) / 16.0).astype(np.float32)

    for bs, n, lim in [(2, 96, 3), (1, 1024, 8), (3, 257, 2)]:
        a = rng.integers(-lim, lim + 1, size=(bs, n, 3)).astype(np.float32)
        b = rng.integers(-lim, lim + 1, size=(bs, n, 3)).astype(np.float32)
        run_case(ref, f"lattice_bs{bs}_n{n}", a, b, dy((bs, n)), dy((bs, n)))

    a = (rng.integers(-32, 33, size=(2, 512, 3)) / 8.0).astype(np.float32)
    b = (rng.integers(-32, 33, size=(2, 512, 3)) / 8.0).astype(np.float32)
    run_case(ref, "dyadic_bs2_n512", a, b, dy((2, 512)), dy((2, 512)))

    for bs, n in [(3, 300), (1, 1500)]:
        a = rng.random((bs, n, 3), dtype=np.float32)
        b = rng.random((bs, n, 3), dtype=np.float32)
        g1 = rng.standard_normal((bs, n)).astype(np.float32)
        g2 = rng.standard_normal((bs, n)).astype(np.float32)
        run_case(ref, f"random_bs{bs}_n{n}", a, b, g1, g2)


if __name__ == "__main__":
    main()
