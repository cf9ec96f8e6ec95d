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


def ncu_traffic(kernel_name, args, world, section=None):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None when
    the capture does not describe this workload."""
    for fn in ("r02_dram_traffic.json", "r01_dram_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                rec = json.load(f)
        except (OSError, ValueError):
            continue
        w = rec.get("workload", {})
        if (w.get("frames"), w.get("scene_points"), w.get("n_gpus"), w.get("scene")) != (args.T, args.M, world, args.scene):
            continue
        if section:
            rec = rec.get(section, {})
        for key, val in rec.items():
            if isinstance(val, (int, float)) and kernel_name.startswith(key):
                return float(val)
    return None


def workload_name(args, world):
    if args.clips > 1:
        tag = "configs[4]"
    elif args.T == 300 and args.M == 1_000_000:
        tag = "configs[1]"
    elif args.T == 1800 and args.M == 5_000_000:
        tag = "configs[2]"
    elif args.T == 30 and args.M == 100_000:
        tag = "configs[0]"
    else:
        tag = "custom"
    clips = f"{args.clips} clips x {args.T // args.clips} frames" if args.clips > 1 else f"T={args.T} frames"
    return f"{tag}: {clips}, V={V}, M={args.M}-point {args.scene} scene, both chamfer directions, exact"


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
            "config": {"workload": workload_name(args, 1), "frames": args.T, "scene_points": args.M,
                       "reference_arm": "the reference's CPU arithmetic (oracle/chamfer_ref_port.py + oracle/smplx_oracle.py) on all "
                                        "host threads; each step times a bounded sample and EXTRAPOLATES linearly to the full workload"},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample,
                             "extrapolated": True},
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
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the ONE JSON line (NCCL prints its banner there)
        dist.init_process_group("nccl", device_id=dev)
    pkg = load_pkg()
    L = pkg._lib.lib()
    W = max(3, args.warmup)
    K = max(1, args.steps)
    idx_dtype = torch.int64 if args.idx64 else torch.int32
    prob = pkg.FitProblem(T=args.T, M=args.M, device=dev, seed=1235, rank=rank, world_size=world, idx_dtype=idx_dtype,
                          front_end=not args.no_front_end, scene_kind=args.scene, fused=not args.no_fused,
                          comm=args.comm, clips=args.clips, scene_order=args.scene_order)

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

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n):
        """n calls of fn between two events, barrier + synchronize on both sides, max over ranks (ms total)."""
        barrier()
        e0.record()
        t_host = time.perf_counter()
        for _ in range(n):
            fn()
        host = (time.perf_counter() - t_host) * 1e3 / n
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), host

    # ---- cold: the very first step builds the scene index, orders the body and searches unseeded ----
    cold_ms, _ = timed(lambda: prob.step(update=True), 1)
    for _ in range(W - 1):
        prob.step(update=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- region 1: eager steps WITH the Adam update (the body moves every step), per-kernel events on ----
    L.fpv_profile_enable(1)
    launches0 = L.fpv_launch_count()
    ms_drift, host_ms_per_step = timed(lambda: prob.step(update=True), K)
    launches = int(L.fpv_launch_count() - launches0)
    recs = collect_profile(L)
    L.fpv_profile_enable(0)
    # ---- region 2: eager, static inputs (no update: every carried seed is already the answer -- the best case) ----
    ms_static, _ = timed(lambda: prob.step(update=False), K)
    # ---- region 3 (headline): the drifting step captured once and replayed as ONE CUDA graph:
    # forward, backward, the cross-rank exchange and the Adam update; then the same end to end from host buffers ----
    graph_info = None
    if not args.no_graph:
        ok = 1
        try:
            prob.capture(update=True)
            prob.step_graph()
        except Exception as ex:  # capture is a host-side optimisation; the eager path above is always measured
            ok, graph_info = 0, {"error": str(ex)[:300]}
        if world > 1:
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if ok:
            ms_graph, _ = timed(prob.step_graph, K)
            for _ in range(2):
                prob.step_e2e_graph()
            t0 = time.perf_counter()
            ms_graph_e2e, _ = timed(prob.step_e2e_graph, K)
            wall_e2e = time.perf_counter() - t0
            graph_info = {"steps_per_s": K / (ms_graph * 1e-3), "ms_per_step": ms_graph / K,
                          "e2e_steps_per_s": K / (ms_graph_e2e * 1e-3), "e2e_h2d_bytes_per_step": prob.h2d_bytes_graph(),
                          "e2e_wall_s": wall_e2e}
        elif graph_info is None:
            graph_info = {"error": "capture failed on another rank"}
    # ---- eager end to end (re-uploads every input incl. the scene shard; no update) ----
    for _ in range(2):
        prob.step_e2e()
    ms_e2e, _ = timed(prob.step_e2e, K)
    clocks = sampler.stop() if rank == 0 else None
    if prob.comm is not None:
        prob.comm.check()
    # ---- cal_loss2 step (mode 'local', the second stage of fitting(mode='local')): the HBM-bound residual stage ----
    local_info = None
    if world == 1 and not args.no_local and args.clips == 1:
        lp = pkg.FitProblem(T=args.T, M=0, device=dev, seed=1235, front_end=not args.no_front_end, mode="local")
        for _ in range(W):
            lp.step(update=True)
        try:
            lp.capture(update=True)
            ms_local, _ = timed(lp.step_graph, K)
            launch = "cuda graph replay"
        except Exception:
            ms_local, _ = timed(lambda: lp.step(update=True), K)
            launch = "eager"
        # SMPL-X fwd 144 MB + bwd 144 MB + vertex smoothness fwd+bwd 75.4 MB at T=300, scaled by T
        bytes_local = (144e6 + 144e6 + 75.4e6) * args.T / 300.0
        local_info = {"step": "FittingOP.cal_loss2 (global_optimization.py:368-447) + backward + Adam", "launch": launch,
                      "ms_per_step": ms_local / K, "steps_per_s": K / (ms_local * 1e-3),
                      "algorithmic_bytes_per_step": bytes_local,
                      "achieved_GBps": bytes_local / (ms_local / K * 1e-3) / 1e9}
    # ---- the reference's literal call sequence (dist1 only on the contact vertices, .repeat(T,1,1) scene) ----
    literal_info = None
    if world == 1 and not args.no_local and args.clips == 1:
        rp = pkg.FitProblem(T=args.T, M=args.M, device=dev, seed=1235, front_end=not args.no_front_end, mode="reference",
                            scene_kind=args.scene, scene_order=args.scene_order)
        for _ in range(W):
            rp.step(update=True)
        try:
            rp.capture(update=True)
            ms_lit, _ = timed(rp.step_graph, K)
            launch = "cuda graph replay"
        except Exception:
            ms_lit, _ = timed(lambda: rp.step(update=True), K)
            launch = "eager"
        literal_info = {"step": "FittingOP.cal_loss as written (global_optimization.py:249-312): ext.chamferDist()(contact_verts, "
                                "s_verts_batch.repeat(T,1,1)) with dist2 discarded, loss = 0.1 contact + smoothing + rec (:570), "
                                "backward, Adam", "launch": launch, "contact_vertices": int(rp.contact_ids.numel()),
                        "ms_per_step": ms_lit / K, "steps_per_s": K / (ms_lit * 1e-3)}
        del rp
    if rank != 0:
        prob.close()
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (eager drifting region) ----
    hbm_peak, peak_src = measured_peaks()
    by = {}
    for name, ms, b, w in recs:
        by.setdefault(name, []).append((ms, b, w))
    dom = max(by.items(), key=lambda kv: sum(x[0] for x in kv[1]))
    dms = statistics.mean(x[0] for x in dom[1])
    dbytes, dwork = dom[1][0][1], dom[1][0][2]
    kernel_ms = sum(x[0] for v in by.values() for x in v) / K
    hbm_gbs = dbytes / (dms * 1e-3) / 1e9
    local_info and local_info.update(frac_of_hbm_peak=local_info["achieved_GBps"] / hbm_peak)
    roofline = {"bound": "hbm", "kernel": dom[0], "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_gbs / hbm_peak, "traffic": ncu_traffic(dom[0], args, world), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dbytes, "ms_per_launch": dms,
                "algorithmic_bytes": "SURVEY 8(d): queries once + candidates once per frame + outputs (the fused kernel's outputs are "
                                     "the per-frame sums and the per-vertex accumulators); the carried seed buffer is NOT counted",
                "note": "exact search through a bounding-sphere hierarchy: the binding resource is FP32 issue on the per-query "
                        "sphere tests and the surviving clusters, not HBM (profiles/r02_nn_sphere_ncu.md)"}
    winst = ncu_traffic(dom[0], args, world, section="warp_instructions")
    if winst:
        # what actually binds the kernel: warp instructions per launch (ncu count of the committed capture; the search is
        # deterministic, so the count belongs to the workload) over the LIVE launch time, against 4 schedulers x SMs x clock
        sm_mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
        peak_issue = 4.0 * torch.cuda.get_device_properties(dev).multi_processor_count * sm_mhz * 1e6
        roofline["issue"] = {"warp_instructions_per_launch": winst, "achieved_Ginst_per_s": winst / (dms * 1e-3) / 1e9,
                             "peak_Ginst_per_s": peak_issue / 1e9, "frac": winst / (dms * 1e-3) / peak_issue,
                             "source": "instruction count: ncu smsp__inst_executed.sum of the committed capture of this build (profiles/r02_nn_sphere_ncu.md); time: live"}
    extra = {}
    st = prob.search_state.stats.get("tiles_searched_b2a")
    if st is not None and dom[0].startswith("nn_sphere"):
        pairs = float(st.reshape(-1)[0].item()) * prob.options.sphere_tile * 128.0
        extra["search_stats"] = {"pairs_evaluated_per_launch": pairs, "fraction_of_all_pairs": pairs / dwork}
    value_eager = K / (ms_drift * 1e-3)
    line = {
        "metric": METRIC, "value": value_eager, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_drift / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, world),
                   "frames": args.T, "scene_points": args.M, "clips": args.clips,
                   "scene_sharding": f"{world} shards (contiguous ranges of the stored scene)" if world > 1 else "none",
                   "index_dtype": "int64" if args.idx64 else "int32",
                   "step": ("6D row -> convert_to_3D_rot -> VPoser decode -> " if not args.no_front_end else "") +
                           "SMPL-X -> scale/world transform -> chamfer (both directions) -> contact / smoothness" +
                           (" / VPoser / DCT" if not args.no_front_end else "") + " residuals -> full backward -> Adam update",
                   "optimizer": "adam lr=0.005 on (body_rotation_rec, scale) inside the timed step, as in the first stage of "
                                "fitting() (global_optimization.py:565-568, :592): the body moves every step",
                   "scene_to_body": "reduced in the search kernel (sum + per-vertex accumulators; no [T,M] output)" if not args.no_fused
                                    else "materialised [T,M] distances + indices",
                   "exchange": ("peer-memory mailbox (keys pushed from the search epilogue, flag barrier)" if prob.comm is not None
                                else "NCCL all_reduce") if world > 1 else "none",
                   "scene_order": ("k-d partitioned" if args.scene_order == "kd" else "Morton-sorted") + " once on the host" +
                                  (", dealt to ranks in blocks of 2048" if world > 1 else ""),
                   "search": "body->scene: Morton-tiled box-culled exact search; scene->body: per-query bounding-sphere hierarchy over "
                             "the body in a frozen Morton order; both seeded with the previous step's winners (hints; results exact)",
                   "l2": "per-step working set (seed buffer + tables > 1.2 GB) exceeds the 126 MB L2; no explicit flush"},
        "roofline": roofline,
        "e2e": {"value": K / (ms_e2e * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": prob.h2d_bytes(),
                "d2h_bytes_per_step": prob.d2h_bytes(),
                "inputs": "every step: observed data, parameters, scale, camera_ext, DCT coefficients AND the scene shard "
                          "from pinned host memory; loss + all gradients back (eager, no update)"},
        "launch": "eager", "host_enqueue_ms_per_step_eager": host_ms_per_step,
        "gpu_launches": launches, "cuda_graph_replay": graph_info, "clocks": clocks,
        "nn_kernels_share_of_step": kernel_ms / (ms_drift / K),
        "kernels": {k: {"launches_per_step": len(v) / K, "ms_mean": statistics.mean(x[0] for x in v)} for k, v in by.items()},
        "ms_per_step_by_condition": {"cold_first_step": cold_ms, "eager_static_inputs": ms_static / K,
                                     "eager_drifting_adam": ms_drift / K,
                                     "graph_drifting_adam": graph_info.get("ms_per_step") if graph_info else None},
        "local_mode": local_info,
        "reference_literal_step": literal_info,
    }
    if graph_info and "steps_per_s" in graph_info:
        # headline = the captured drifting step; keep the eager measurements alongside
        line["eager"] = {"value": line["value"], "ms_per_step": line["ms_per_step"], "e2e": line["e2e"]}
        line["value"], line["ms_per_step"], line["launch"] = graph_info["steps_per_s"], graph_info["ms_per_step"], "cuda graph replay"
        line["e2e"] = {"value": graph_info["e2e_steps_per_s"], "unit": "steps/s",
                       "h2d_bytes_per_step": graph_info["e2e_h2d_bytes_per_step"], "d2h_bytes_per_step": prob.d2h_bytes(),
                       "inputs": "every step: the observed data from pinned host memory into the captured step's static buffer; "
                                 "loss + the updated leaves back; parameters and optimiser state live on the device and the "
                                 "scene is resident, as in the reference (global_optimization.py:173-188)"}
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_step(args.T, args.M, budget_s=12.0)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample,
                                "extrapolated": True}
    print(json.dumps(line), flush=True)
    prob.close()
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
    ap.add_argument("--clips", type=int, default=1, help="independent clips batched into one step (configs[4]: 16); T is the total")
    ap.add_argument("--no-fused", action="store_true", help="materialise the [T,M] scene->body outputs instead of the fused sum")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"], help="transport of the sharded key / gradient exchange")
    ap.add_argument("--no-local", action="store_true", help="skip the cal_loss2 (mode 'local') leg")
    ap.add_argument("--scene-order", default="kd", choices=["kd", "morton"],
                    help="one-time host-side ordering of the scene: left-balanced k-d partition (default) or Morton curve")
    ap.add_argument("--no-front-end", action="store_true",
                    help="optimise the axis-angle row directly (skip the 6D codec, VPoser decode and DCT prior)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
