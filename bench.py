#!/usr/bin/env python
"""bench.py -- fit steps/sec of the global-optimisation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--T 300] [--M 1000000]

A "step" is one pass of the hot path over one synthetic clip (fit.FitProblem.step): SMPL-X forward,
scale + world transform, chamfer both directions against the scene, robust contact + smoothness
residuals, full backward to the per-frame parameters.  Workload at every N: BASELINE.json configs[1]
(T=300 frames, V=10,475, 1M-point scene); with N>1 the scene is sharded over the ranks, so the same job
gets faster ("scaling": "strong").  Prints ONE JSON line on rank 0.

--impl reference times the reference's CPU path (the torch restatement of chamfer_python.py plus the
torch SMPL-X forward/backward; kind "port": the reference is Python and /root/reference is absent on the
GPU box) on the host cores, on a bounded sample of the same workload, and reports the same metric.
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fit steps/sec (T=300, 1M-pt scene)"
V = 10475


def load_pkg():
    pkg = importlib.import_module("4dcapture-fpv_b200")
    sys.modules["fpv_b200"] = pkg
    return pkg


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# CPU reference path (kind "port"): the only place besides tests/ and smoke() that executes oracle/.
# ------------------------------------------------------------------------------------------------------
def cpu_reference_step(T: int, M: int, budget_s: float = 12.0, seed: int = 1235):
    """Times the reference's CPU path on a bounded sample and extrapolates to one full step.

    chamfer: chamfer_ref_port.distChamfer (chamfer_python.py arithmetic) for ONE frame of V body vertices
    against a scene sample sized from a probe so the run takes ~budget_s, scaled linearly to T frames x M
    points (brute force is exactly linear in both).  SMPL-X: the float32 torch restatement, forward +
    backward with the vertex-smoothness loss, at the full T.
    """
    from oracle import chamfer_ref_port as port
    from oracle import smplx_oracle as so
    pkg = load_pkg()
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    consts = pkg.synthetic.make_body_constants(seed)
    clip = pkg.synthetic.make_clip_params(T, seed)
    keys = ["betas", "global_orient", "body_pose", "transl", "left_hand_pose", "right_hand_pose"]
    p = {k: clip[k].clone().requires_grad_(True) for k in keys}
    t0 = time.perf_counter()
    verts, joints = so.smplx_forward(consts, **p, dtype=torch.float32)
    diff = verts[0:-1] - verts[1:]
    loss = torch.mean(torch.abs(diff[0:-1] - diff[1:])) + torch.mean(torch.abs(joints[0:-1] - joints[1:]))
    loss.backward()
    t_smplx = time.perf_counter() - t0
    body = verts[0].detach().contiguous().unsqueeze(0)
    scene = pkg.synthetic.make_scene(M, "uniform", seed)
    probe_m = min(M, 65536)
    t0 = time.perf_counter()
    port.distChamfer(body, scene[:probe_m].unsqueeze(0))
    t_probe = time.perf_counter() - t0
    m_s = int(min(M, max(probe_m, probe_m * budget_s / max(t_probe, 1e-3))))
    t0 = time.perf_counter()
    port.distChamfer(body, scene[:m_s].unsqueeze(0))
    t_ch = time.perf_counter() - t0
    t_full = t_ch * (M / m_s) * T + t_smplx
    sample = (f"chamfer: 1 of {T} frames x {m_s} of {M} scene points in {t_ch:.2f}s, scaled x{T} x{M / m_s:.2f}; "
              f"SMPL-X fwd+bwd full T={T} in {t_smplx:.2f}s; torch CPU, {cores} threads")
    return 1.0 / t_full, cores, sample, t_ch + t_smplx + t_probe


def ncu_traffic(kernel_name, args, world):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None when
    the capture does not describe this workload."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_dram_traffic.json")) as f:
            rec = json.load(f)
    except (OSError, ValueError):
        return None
    w = rec.get("workload", {})
    if (w.get("frames"), w.get("scene_points"), w.get("n_gpus"), w.get("scene")) != (args.T, args.M, world, args.scene):
        return None
    for key, val in rec.items():
        if isinstance(val, (int, float)) and kernel_name.startswith(key):
            return float(val)
    return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, sample, cores, spent = [], "", 1, 0.0
    n = max(1, min(args.steps, 3))
    for _ in range(n):
        v, cores, sample, dt = cpu_reference_step(args.T, args.M, budget_s=10.0)
        vals.append(v)
        spent += dt
    value = statistics.median(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": n, "warmup": 0, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: T={args.T} frames, V={V}, M={args.M}-point {args.scene} scene, both chamfer directions, exact",
                       "frames": args.T, "scene_points": args.M,
                       "reference_arm": "the reference's CPU arithmetic (oracle/chamfer_ref_port.py + oracle/smplx_oracle.py) on all "
                                        "host threads; each step times a bounded sample and extrapolates to the full workload"},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def collect_profile(L):
    recs = []
    name = ctypes.create_string_buffer(48)
    ms, b, w = ctypes.c_float(), ctypes.c_double(), ctypes.c_double()
    for i in range(L.fpv_profile_count()):
        if L.fpv_profile_get(i, name, ctypes.byref(ms), ctypes.byref(b), ctypes.byref(w)) == 0:
            recs.append((name.value.decode(), ms.value, b.value, w.value))
    return recs


def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = load_pkg()
    L = pkg._lib.lib()
    W = max(3, args.warmup)
    K = max(1, args.steps)
    idx_dtype = torch.int64 if args.idx64 else torch.int32
    prob = pkg.FitProblem(T=args.T, M=args.M, device=dev, seed=1235, rank=rank, world_size=world, idx_dtype=idx_dtype,
                          front_end=not args.no_front_end, scene_kind=args.scene)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    for _ in range(W):
        prob.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: device-resident steps, per-kernel events on for the roofline ----
    L.fpv_profile_enable(1)
    launches0 = L.fpv_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host = time.perf_counter()
    for _ in range(K):
        if args.jitter > 0:
            with torch.no_grad():
                prob.params.add_(torch.randn_like(prob.params) * args.jitter)
        prob.step()
    host_ms_per_step = (time.perf_counter() - t_host) * 1e3 / K     # host time to ENQUEUE a step (no sync inside)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = int(L.fpv_launch_count() - launches0)
    recs = collect_profile(L)
    L.fpv_profile_enable(0)
    # ---- timed region 2: end to end from host buffers (H2D of every input, D2H of loss + gradients) ----
    for _ in range(2):
        prob.step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        prob.step_e2e()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), 0.0))
    wall_e2e = time.perf_counter() - t0
    # ---- timed region 3: the same step captured once and replayed as ONE CUDA graph (FitProblem.capture, SURVEY 8f
    # row f1): identical kernels and collectives, no per-launch host work.  When it is available it is the headline
    # (`value`, `e2e`); the eager numbers of regions 1-2 stay in the line under `eager`. ----
    graph_info = None
    if world == 1 and not args.no_graph:   # the sharded step stays eager (see FitProblem.capture)
        ok = 1
        try:
            prob.capture()
            prob.step_graph()
        except Exception as ex:  # capture is a host-side optimisation; the eager path above is always measured
            ok, graph_info = 0, {"error": str(ex)[:200]}
        if world > 1:
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if ok:
            barrier()
            e0.record()
            for _ in range(K):
                prob.step_graph()
            e1.record()
            barrier()
            ms_graph = max_over_ranks(e0.elapsed_time(e1))
            for _ in range(2):
                prob.step_e2e_graph()
            barrier()
            e0.record()
            for _ in range(K):
                prob.step_e2e_graph()
            e1.record()
            barrier()
            ms_graph_e2e = max_over_ranks(e0.elapsed_time(e1))
            graph_info = {"steps_per_s": K / (ms_graph * 1e-3), "ms_per_step": ms_graph / K,
                          "e2e_steps_per_s": K / (ms_graph_e2e * 1e-3), "e2e_h2d_bytes_per_step": prob.h2d_bytes_graph()}
        elif graph_info is None:
            graph_info = {"error": "capture failed on another rank"}
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel ----
    hbm_peak, peak_src = measured_peaks()
    by = {}
    for name, ms, b, w in recs:
        by.setdefault(name, []).append((ms, b, w))
    dom = max(by.items(), key=lambda kv: sum(x[0] for x in kv[1]))
    dms = statistics.mean(x[0] for x in dom[1])
    dbytes, dwork = dom[1][0][1], dom[1][0][2]
    fma = ctypes.c_double()
    pkg._lib.check(L.fpv_fp32_probe(ctypes.byref(fma), pkg._lib.stream_ptr()), "fpv_fp32_probe")
    kernel_ms = sum(x[0] for v in by.values() for x in v) / K
    value = K / (ms_total * 1e-3)
    hbm_gbs = dbytes / (dms * 1e-3) / 1e9
    extra = {}
    if dom[0].startswith("nn_tc"):
        # tensor-core filter: one K=16 TF32 contraction (32 flop) per query-candidate pair
        tf = dwork * 32.0 / (dms * 1e-3) / 1e12
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        tpeak = float(peaks.get("bf16_tflops", 1590.0))
        roofline = {"bound": "tensor", "kernel": dom[0], "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                    "traffic": None, "peak_source": ("measured bf16 (MEASURED_PEAKS.json)" if peaks else "fallback bf16 (B200_PROFILING.md)") +
                    "; the kernel issues kind::tf32 MMAs whose nominal dense rate is half the bf16 rate",
                    "flop_per_pair": 32, "pairs_per_launch": dwork, "pairs_per_s": dwork / (dms * 1e-3), "ms_per_launch": dms,
                    "hbm_algorithmic_GBps": hbm_gbs,
                    "note": "co-limited by the fp32 min-reduction of the accumulators on the ALU pipe (profiles/r01_nn_tc_ncu.md)"}
    else:
        roofline = {"bound": "hbm", "kernel": dom[0], "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_gbs / hbm_peak, "traffic": ncu_traffic(dom[0], args, world), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": dbytes, "ms_per_launch": dms}
        if dom[0].startswith("nn_culled") or dom[0].startswith("nn_sphere"):
            ch = importlib.import_module("4dcapture-fpv_b200.chamfer")
            b2a = "rep" in dom[0] or dom[0].startswith("nn_sphere")
            st = ch.LAST_STATS.get("tiles_searched_b2a" if b2a else "tiles_searched")
            if st is not None:
                tile = ch.SPHERE_TILE if dom[0].startswith("nn_sphere") else (32 if "rep" in dom[0] else 64)
                pairs = float(st.reshape(-1)[0].item()) * tile * 128.0          # 128 queries of a warp meet every point of a searched tile
                lane_ops = pairs * 6.0 / (dms * 1e-3)
                roofline["note"] = ("exact search with culling: the binding resource is FP32 issue on the surviving tiles "
                                    "plus the per-query cluster tests, not HBM: see `simt`")
                extra["simt"] = {"pairs_evaluated_per_launch": pairs, "fraction_of_all_pairs": pairs / dwork,
                                 "fp32_lane_ops_per_pair": 6, "achieved_lane_ops_per_s": lane_ops,
                                 "peak_lane_fma_per_s_measured": fma.value, "frac": lane_ops / fma.value if fma.value else None}
        if dom[0].startswith("nn_search"):
            lane_ops = dwork * 6.0 / (dms * 1e-3)                  # 3 sub + 1 mul + 2 fma per pair
            roofline["note"] = "exact brute force is FP32-issue-bound, not HBM-bound: see `simt`"
            extra["simt"] = {"pairs_per_launch": dwork, "pairs_per_s": dwork / (dms * 1e-3), "fp32_lane_ops_per_pair": 6,
                             "achieved_lane_ops_per_s": lane_ops, "peak_lane_fma_per_s_measured": fma.value,
                             "frac": lane_ops / fma.value if fma.value else None}
    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: T={args.T} frames, V={V}, M={args.M}-point {args.scene} scene, both chamfer directions, exact",
                   "frames": args.T, "scene_points": args.M, "scene_sharding": f"{world} shards (contiguous ranges of the stored scene)" if world > 1 else "none",
                   "index_dtype": "int64" if args.idx64 else "int32",
                   "step": ("6D row -> convert_to_3D_rot -> VPoser decode -> " if not args.no_front_end else "") +
                           "SMPL-X -> scale/world transform -> chamfer (both directions) -> contact / smoothness" +
                           (" / VPoser / DCT" if not args.no_front_end else "") + " residuals -> full backward",
                   "param_jitter_per_step": args.jitter,
                   "scene_order": "Morton-sorted once on the host" + (", dealt to ranks in blocks of 2048" if world > 1 else ""),
                   "search": "body->scene: Morton-tiled box-culled exact search; scene->body: per-query bounding-sphere hierarchy over the Morton-sorted body; both seeded with the previous step's winners (hints; results exact)",
                   "l2": "per-step working set (>=2.4 GB of [T,M] outputs) exceeds the 126 MB L2; no explicit flush"},
        "roofline": roofline,
        "e2e": {"value": K / (ms_e2e * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": prob.h2d_bytes(),
                "d2h_bytes_per_step": prob.d2h_bytes(), "wall_s": wall_e2e,
                "inputs": "every step: observed data, parameters, scale, camera_ext, DCT coefficients AND the scene shard "
                          "from pinned host memory; loss + all gradients back"},
        "launch": "eager", "host_enqueue_ms_per_step_eager": host_ms_per_step,
        "gpu_launches": launches, "cuda_graph_replay": graph_info, "clocks": clocks, "fp32_lane_fma_per_s_measured": fma.value,
        "nn_kernels_share_of_step": kernel_ms / (ms_total / K),
        "kernels": {k: {"launches_per_step": len(v) / K, "ms_mean": statistics.mean(x[0] for x in v)} for k, v in by.items()},
    }
    if graph_info and "steps_per_s" in graph_info:
        # headline = the captured step; keep the eager measurements alongside
        line["eager"] = {"value": line["value"], "ms_per_step": line["ms_per_step"], "e2e": line["e2e"]}
        line["value"], line["ms_per_step"], line["launch"] = graph_info["steps_per_s"], graph_info["ms_per_step"], "cuda graph replay"
        line["e2e"] = {"value": graph_info["e2e_steps_per_s"], "unit": "steps/s",
                       "h2d_bytes_per_step": graph_info["e2e_h2d_bytes_per_step"], "d2h_bytes_per_step": prob.d2h_bytes(),
                       "inputs": "every step: observed data, parameters, scale, camera_ext, DCT coefficients from pinned host "
                                 "memory into the captured step's static buffers; loss + all gradients back; the scene is "
                                 "resident, uploaded once before the loop like the reference (global_optimization.py:173-176)"}
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_step(args.T, args.M, budget_s=12.0)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--T", type=int, default=300)
    ap.add_argument("--M", type=int, default=1_000_000)
    ap.add_argument("--idx64", action="store_true", help="reference-faithful int64 index outputs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="skip the informational CUDA-graph replay leg")
    ap.add_argument("--scene", default="uniform", choices=["uniform", "surface"],
                    help="synthetic scene: uniform in the room volume (configs 1, 2, 4) or points on surfaces (configs 3, 5)")
    ap.add_argument("--jitter", type=float, default=0.0,
                    help="add N(0, jitter^2) to the optimised parameters before every step (an optimiser-like drift; "
                         "shows that the carried seeds and the scene cache do not depend on identical inputs)")
    ap.add_argument("--no-front-end", action="store_true",
                    help="optimise the axis-angle row directly (skip the 6D codec, VPoser decode and DCT prior)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
