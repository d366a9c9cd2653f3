#!/usr/bin/env python
"""Benchmark of the VANeRF novel-view render path on B200 (BASELINE.json metric: rays/s and ms per 334x512 view).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|bf16] [--workload B|C] [--impl ours|reference]

A step = one full 334x512 novel view (171 008 rays, V=3 source views, 64 coarse + 128 fine network evaluations per
ray, `fine=True, uniform=True`; BASELINE.json configs[1]) on synthetic inputs (vanerf_b200.synthetic) with random-init
weights of the configs/vanerf.json architecture.  N > 1 (torchrun): the view is partitioned over the ranks with the
reference's own interleaved decomposition (src/model.py:1050-1085) and the output tiles are all-gathered over NCCL;
total work is fixed ("strong" scaling).

One JSON line on rank 0:  value = rays/s with inputs resident in HBM (CUDA events, max over ranks);  e2e = the same
through the public API (vanerf_b200.model.VANeRF.render_pifu_nerf) with pinned HOST buffers: per step H2D of the
source images / masks / feature maps, per-frame setup, render, D2H of the image;  roofline = dominant kernel (fused
MLP) from CUDA events recorded around its launches inside the timed region;  cpu_baseline = the oracle (CPU port of the
reference's algorithm) on a bounded sample (N=1, rank 0 only).
`--impl reference` times that CPU port on the host cores (the reference itself needs /root/reference, kaolin and
pytorch3d, none of which exist on the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, V = 512, 334, 3
S_C, S_F = 64, 64
FLOP_PER_SAMPLE = lambda v: 2.0 * (v * 142580 + 15488)
GATHER_BYTES_PER_SAMPLE = lambda v, e: v * (4 * (64 + 8 + 8 + 3 + 1) + 2 * (64 + 8 + 29)) * e + 12


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(precision):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu capture
    (profiles/r01_<precision>_traffic.json, written by tools/ncu_key.py); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", f"r01_{precision}_traffic.json")
    try:
        return json.load(open(p))["dram_bytes_per_launch"]
    except Exception:
        return None


def partition(rank, world):
    """Interleaved pixel partition: G=2 -> 1x2, 4 -> 2x2, 8 -> 4x2 (y-period x x-period)."""
    gy, gx = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (4, 2)}.get(world, (world, 1))
    ry, rx = divmod(rank, gx)
    ys, xs = np.meshgrid(np.arange(ry, H, gy), np.arange(rx, W, gx), indexing="ij")
    return np.stack([xs, ys], -1).reshape(-1, 2).astype(np.int32), (gy, gx, ry, rx)


def cpu_port_rays_per_s(n_side, fine, steps=1, warmup=0, layout="narrow"):
    """The oracle (CPU restatement of the reference) on an n_side x n_side lattice of target pixels."""
    import torch
    from oracle import oracle_torch as OT
    from vanerf_b200 import synthetic, weights
    torch.set_num_threads(os.cpu_count() or 1)
    sc = synthetic.make_scene(H, W, V, layout=layout)
    inp = synthetic.to_torch(sc)
    sd = weights.init_state_dict(H, W, mode="ref")
    ii, jj = np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij")
    pix = np.stack([(5 + (W - 10) * ii // n_side).ravel(), (8 + (H - 16) * jj // n_side).ravel()], 1).astype(np.int64)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc = OT.Oracle(sd, inp)                    # per-frame setup is part of a view
        orc.render(fine=fine, pixels=pix, S_c=S_C, S_f=S_F)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return pix.shape[0] / float(np.mean(times)), float(np.mean(times)), pix.shape[0]


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    n_side = 12
    rps, sec, n = cpu_port_rays_per_s(n_side, True, steps=max(1, args.steps), warmup=min(args.warmup, 1),
                                      layout="bvv" if args.workload == "C" else "narrow")
    sample = f"{n} rays ({n_side}x{n_side} lattice of the 334x512 view), 64 coarse + 128 fine evaluations/ray, V=3, per-frame setup included"
    line = {"impl": "reference", "metric": "rays_per_s", "value": rps, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(args, "cpu"),
            "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config_dict(args, where):
    return {"workload": f"{'C: vanerf_bvv big-view-variation' if args.workload == 'C' else 'B: vanerf.json'} full 334x512 novel view, "
                        f"171008 rays, V=3 source views, 64 coarse + 128 fine evaluations/ray (fine=True, uniform=True)",
            "precision": args.precision, "rays_per_view": H * W, "source_views": V, "samples": [S_C, S_C + S_F],
            "ray_partition": "interleaved pixels across ranks + NCCL all_gather of output tiles" if args.gpus > 1 else "single GPU",
            "l2": "L2 flushed (256 MiB write) between timed iterations; per-step working set (gather records) exceeds L2",
            "where": where}


def run_dynamic(args):
    """Workload D (BASELINE.json configs[3], render_dynamic): a sequence of frames, each with its own mesh, source images
    and feature maps and its own 334x512 target camera on a 360-degree path.  A step = `--frames` frames end to end
    from pinned host buffers: H2D of the frame's maps, per-frame setup (BVH, vertex visibility, vertex tables, bf16 maps),
    render, D2H of the image.  Frames are dealt round-robin to the ranks (vanerf_b200.dynamic), no collective inside the
    timed path; per-GPU work shrinks with N ("strong" scaling over a fixed sequence)."""
    import torch
    import torch.distributed as dist
    from vanerf_b200 import dynamic, synthetic, weights
    from vanerf_b200.model import VANeRF

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F = max(args.frames, world)
    ids = dynamic.frames_for_rank(F, rank, world)
    net = VANeRF(device=dev, precision=args.precision).eval()
    net.load_state_dict(weights.init_state_dict(H, W, mode="ref"))
    pin = lambda t: t.contiguous().pin_memory()
    frames = {}
    for f in ids:                                               # synthetic frames of this rank, pinned host memory
        fr = synthetic.to_torch(synthetic.make_scene(H, W, V, frame=f))
        fr["img"], fr["feat_tex"] = pin(fr["img"]), pin(fr["feat_tex"])
        fr["feat_geo"] = [pin(t) for t in fr["feat_geo"]]
        fr["src_foreground_mask"] = pin(fr["src_foreground_mask"].to(torch.uint8))
        frames[f] = fr
    K = frames[ids[0]]["cam_tar"]["K"][0, :3, :3]
    cams = dynamic.orbit_cameras(F, K, width=W, height=H)
    out_host = torch.empty(len(ids), 1, 8, H, W).pin_memory()

    def step():
        dynamic.render_sequence(net, lambda f: frames[f], F, lambda f: [cams[f]], rank, world, out_host=out_host,
                                fine=True, sample_per_ray_c=S_C, sample_per_ray_f=S_F)
        torch.cuda.current_stream().synchronize()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = net.renderer.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        a.record()
        step()
        b.record()
    sync_all()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    if rank == 0:
        clk = clocks.stop()
        rps = F * H * W / (ms_step * 1e-3)
        h2d = dynamic.h2d_bytes(frames[ids[0]]) * F
        line = {"metric": "rays_per_s", "value": rps, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "ms_per_frame": ms_step / F, "frames_per_s": F / (ms_step * 1e-3), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": f"D: render_dynamic, {F} frames x one 334x512 view (171008 rays, V=3, 64 coarse + 128 fine "
                                       f"evaluations/ray), per-frame mesh / maps / camera, per-frame setup included",
                           "precision": args.precision, "frame_partition": "round robin over ranks, no collective in the path",
                           "l2": "every frame brings new maps and records: the working set exceeds L2", "where": "gpu"},
                "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": F * 8 * H * W * 4,
                        "note": "the timed step IS the end-to-end path (host buffers in, host images out)"},
                "gpu_launches": int(net.renderer.launches - l0), "clocks": clk}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # headline = the bf16-MLP tensor-core path (north star kernel 2); the fp32 FFMA path is measured next to it
    ap.add_argument("--precision", default=os.environ.get("VANERF_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--no-fp32-path", action="store_true", help="skip the secondary fp32-path measurement")
    ap.add_argument("--workload", default="B", choices=["B", "C", "D"])
    ap.add_argument("--frames", type=int, default=8, help="workload D: frames per step (BASELINE.json configs[3] uses 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reuse-variant", action="store_true", help="skip the secondary coarse-reuse measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "D":
        return run_dynamic(args)

    import torch
    import torch.distributed as dist
    from vanerf_b200 import _lib as L
    from vanerf_b200 import synthetic, weights
    from vanerf_b200.model import VANeRF

    # stdout carries exactly one JSON line: NCCL's own banner / debug output goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W_steps = max(3, args.warmup)
    prec = L.FP32 if args.precision == "fp32" else L.BF16

    sc = synthetic.make_scene(H, W, V, layout="bvv" if args.workload == "C" else "narrow")
    inp_host = synthetic.to_torch(sc)                                   # CPU tensors, reference layouts
    sd = weights.init_state_dict(H, W, mode="ref")
    net = VANeRF(device=dev, precision=args.precision).eval()
    net.load_state_dict(sd)
    r = net.renderer

    # ---------------- device-resident inputs
    mv = lambda t: t.to(dev)
    inp = dict(inp_host)
    inp["img"], inp["feat_tex"], inp["src_foreground_mask"] = mv(inp_host["img"]), mv(inp_host["feat_tex"]), mv(inp_host["src_foreground_mask"])
    inp["feat_geo"] = [mv(t) for t in inp_host["feat_geo"]]
    r.set_frame(inp["img"], inp["cam_in"], inp["targets"], inp["sp_data"], inp["feat_geo"], inp["feat_tex"], inp["src_foreground_mask"])
    tar = r.make_target(inp["cam_tar"], inp["bounds"])
    pix_np, (gy, gx, ry, rx) = partition(rank, world)
    pix = torch.from_numpy(pix_np).to(dev)
    R_local = pix.shape[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tiles = [torch.empty(R_local, 8, device=dev) for _ in range(world)] if world > 1 else None

    def step_resident():
        oc, of = r.render_rays(tar, pix, S_C, S_F, True, prec)
        if world > 1:
            dist.all_gather(tiles, of)
        return of

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W_steps):
        step_resident()
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    r.timing(True)
    r.timing_read(reset=True)
    launches0 = r.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    for a, b in ev:
        flush.fill_(1)
        a.record()
        step_resident()
        b.record()
    sync_all()
    ms_local = sum(a.elapsed_time(b) for a, b in ev)
    ktimes = r.timing_read(reset=True)
    r.timing(False)
    launches = r.launches - launches0
    t = torch.tensor([ms_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = H * W / (ms_step * 1e-3)

    # ---------------- end to end through the public API, host buffers
    pin = lambda x: x.contiguous().pin_memory()
    h_img, h_tex, h_fg = pin(inp_host["img"]), pin(inp_host["feat_tex"]), pin(inp_host["src_foreground_mask"].to(torch.uint8))
    h_g0, h_g1 = pin(inp_host["feat_geo"][0]), pin(inp_host["feat_geo"][1])
    h_out = torch.empty(H * W, 8).pin_memory() if rank == 0 else None
    h2d = sum(x.numel() * x.element_size() for x in (h_img, h_tex, h_fg, h_g0, h_g1))
    d2h = H * W * 8 * 4

    def step_e2e():
        d_img, d_tex = h_img.to(dev, non_blocking=True), h_tex.to(dev, non_blocking=True)
        d_fg = h_fg.to(dev, non_blocking=True).bool()
        d_g = [h_g0.to(dev, non_blocking=True), h_g1.to(dev, non_blocking=True)]
        out = VANeRF.batch_render_pifu_nerf(net, d_img, inp["cam_in"], inp["hand_type"], inp["targets"], V, inp["cam_tar"], 1, 0, None,
                                            d_g, d_tex, None, inp["sp_data"], inp["objcenter"], fine=True, uniform=True,
                                            sample_per_ray_c=S_C, sample_per_ray_f=S_F, src_foreground_mask=d_fg, bounds=inp["bounds"],
                                            pixel_override=pix[None])
        rows = torch.cat([out["tex_fg_fine"][0].reshape(3, -1).T, out["depth_fine"].reshape(-1, 1), out["alpha_fine"].reshape(-1, 1),
                          out["sdf"].reshape(-1, 1), out["tex_fg"][0].reshape(3, -1).T[:, :2]], 1).contiguous()
        if world > 1:
            dist.all_gather(tiles, rows)
            full = torch.cat(tiles, 0)
        else:
            full = rows
        if rank == 0:
            h_out.copy_(full, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        step_e2e()
    sync_all()
    n_e2e = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        step_e2e()
    sync_all()
    te = torch.tensor([(time.perf_counter() - t0) / n_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = H * W / float(te.item())

    # ---------------- the other precision path of BASELINE.json configs[1], one timed view (N = 1 only)
    fp32_extra = None
    if world == 1 and args.precision == "bf16" and not args.no_fp32_path:
        r.render_rays(tar, pix, S_C, S_F, True, L.FP32)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1)
        a.record()
        r.render_rays(tar, pix, S_C, S_F, True, L.FP32)
        b.record()
        torch.cuda.synchronize()
        fp32_extra = {"ms_per_view": a.elapsed_time(b), "value": H * W / (a.elapsed_time(b) * 1e-3), "unit": "rays/s",
                      "note": "fp32 FFMA path (k_gather + k_mlp_simt), 1 warm-up + 1 timed view, inputs resident"}

    # ---------------- coarse reuse (vanerf_set_reuse_coarse): same output bits, 64 + 64 instead of 64 + 128 evaluations per ray
    reuse_extra = None
    if world == 1 and not args.no_reuse_variant:
        r.set_reuse_coarse(True)
        r.render_rays(tar, pix, S_C, S_F, True, prec)
        torch.cuda.synchronize()
        evr = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(1, args.steps))]
        for a, b in evr:
            flush.fill_(1)
            a.record()
            r.render_rays(tar, pix, S_C, S_F, True, prec)
            b.record()
        torch.cuda.synchronize()
        r.set_reuse_coarse(False)
        ms_r = sum(a.elapsed_time(b) for a, b in evr) / len(evr)
        reuse_extra = {"ms_per_view": ms_r, "value": H * W / (ms_r * 1e-3), "unit": "rays/s", "evaluations_per_ray": S_C + S_F,
                       "note": "NOT the headline: fine pass evaluates only the 64 new depths and reuses the coarse pass for the 64 coarse "
                               "depths of the merged set (bit-identical output, tests: *_coarse_reuse_is_bit_identical); inputs resident"}

    if rank == 0:
        pk = peaks()
        clk = clocks.stop()
        n_samples_step = R_local * (S_C + S_C + S_F)
        mlp_ms, mlp_n = ktimes["mlp"]
        gat_ms, gat_n = ktimes["gather"]
        flops = FLOP_PER_SAMPLE(V) * n_samples_step * args.steps
        ach_tf = flops / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
        gather_elem = 4 if args.precision == "fp32" else 2          # fp32 maps vs bf16 maps / vertex tables
        gbytes = GATHER_BYTES_PER_SAMPLE(V, gather_elem) * n_samples_step * args.steps
        ach_gb = gbytes / (gat_ms * 1e-3) / 1e9 if gat_ms > 0 else 0.0
        line = {
            "metric": "rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": W_steps,
            "ms_per_step": ms_step, "ms_per_view": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic", "config": config_dict(args, "gpu"),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_view": 1e3 * float(te.item())},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_mlp_simt (fused PE + fusion + MLP, fp32 FFMA)" if args.precision == "fp32" else "k_mlp_tc (tcgen05)",
                         "bound": "tensor", "achieved": ach_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                         "frac": ach_tf / pk["tf_sust"], "traffic": ncu_traffic(args.precision),
                         "peak_source": pk["src"] + " bf16 sustained",
                         "launches": int(mlp_n), "avg_launch_ms": mlp_ms / max(1, mlp_n),
                         "algorithmic_flop_per_sample": FLOP_PER_SAMPLE(V)},
            "roofline_gather": {"kernel": "k_gather" if args.precision == "fp32" else "k_gather_tc", "bound": "hbm", "achieved": ach_gb, "peak": pk["hbm"], "unit": "GB/s",
                                "frac": ach_gb / pk["hbm"], "launches": int(gat_n), "avg_launch_ms": gat_ms / max(1, gat_n),
                                "algorithmic_bytes_per_sample": GATHER_BYTES_PER_SAMPLE(V, gather_elem)},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in ktimes.items()},
            "clocks": clk,
        }
        if fp32_extra is not None:
            line["fp32_path"] = fp32_extra
        if reuse_extra is not None:
            line["coarse_reuse"] = reuse_extra
        if world == 1 and not args.no_cpu_baseline:
            rps, sec, n = cpu_port_rays_per_s(32, False)
            line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{n} rays (32x32 lattice), 64 coarse samples/ray, V=3, fine=False (BASELINE.json configs[0]); {sec:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
