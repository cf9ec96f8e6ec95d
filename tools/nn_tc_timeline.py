"""Dumps the clock64 pipeline timeline of CTA 0 of nn_tc_kernel (MMA issue / epilogue wait / ready / done)."""
import importlib, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fpv = importlib.import_module("4dcapture-fpv_b200")
L = fpv._lib.lib()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
M = 1_000_000
scene = (torch.rand(M, 3, generator=g) * torch.tensor([8.0, 8.0, 3.0]) - torch.tensor([4.0, 4.0, 0.0])).to(dev)
body = (torch.rand(8, 10475, 3, generator=g) * torch.tensor([0.6, 0.6, 1.8]) + torch.tensor([0.5, -1.0, 0.0])).to(dev)
pl = fpv.pack_planes(scene)
for ns in (128, 64):
    L.fpv_nn_set_engine(2, ns << 8)
    fpv.nn_search(body, pl, M)
    dbg = torch.zeros(1024, dtype=torch.int64, device=dev)
    L.fpv_nn_tc_debug(fpv._lib.ptr(dbg))
    fpv.nn_search(body, pl, M)
    torch.cuda.synchronize()
    L.fpv_nn_tc_debug(None)
    d = dbg.cpu().view(4, 256)
    t0 = int(d[0, 0])
    print(f"--- sub-tile {ns}: seq | mma_issue | epi_wait_start | epi_ready | epi_done   (cycles since first issue; epilogue stamps of warp 4 (set 0) and warp 8 (set 1))")
    for s in list(range(192, 232)):
        row = [int(d[r, s]) - t0 if int(d[r, s]) else -1 for r in range(4)]
        print(f"{s:4d} {row[0]:9d} {row[1]:9d} {row[2]:9d} {row[3]:9d}   wait={row[2]-row[1]:6d} proc={row[3]-row[2]:6d}")
    iss = d[0, 100:250]
    print("steady-state cycles per MMA visit:", float((iss[-1] - iss[0])) / (len(iss) - 1))
