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
    (profiles/r02_<precision>_traffic.json, else r01; written by tools/ncu_key.py); None when no capture is committed."""
    for rnd in ("r02", "r01"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", f"{rnd}_{precision}_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            continue
    return None


def cpu_port(n_side, fine, steps=1, warmup=0, layout="narrow"):
    """The oracle (CPU restatement of the reference: oracle/oracle_torch.py + oracle/geom_oracle.c) on an n_side x n_side lattice
    of target pixels of the 334x512 view.  Returns rays/s of the render alone, its seconds, the ray count and the seconds
    of the per-frame setup (vertex visibility raster, vertex tables, global feature), which is timed separately."""
    import torch
    from oracle import oracle_torch as OT
    from vanerf_b200 import synthetic, weights
    torch.set_num_threads(os.cpu_count() or 1)
    sc = synthetic.make_scene(H, W, V, layout=layout)
    inp = synthetic.to_torch(sc)
    sd = weights.init_state_dict(H, W, mode="ref")
    ii, jj = np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij")
    pix = np.stack([(5 + (W - 10) * ii // n_side).ravel(), (8 + (H - 16) * jj // n_side).ravel()], 1).astype(np.int64)
    t_render, t_setup = [], []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc = OT.Oracle(sd, inp)
        orc.frame_setup()
        t1 = time.perf_counter()
        orc.render(fine=fine, pixels=pix, S_c=S_C, S_f=S_F)
        t2 = time.perf_counter()
        if i >= warmup:
            t_setup.append(t1 - t0)
            t_render.append(t2 - t1)
    sec = float(np.mean(t_render))
    return pix.shape[0] / sec, sec, pix.shape[0], float(np.mean(t_setup))


def run_reference(args):
    """Reference arm: the reference's algorithm on the host cores (the oracle port: the reference itself needs /root/reference,
    kaolin and pytorch3d, none of which exist on the GPU box).  A step = a bounded sample of the workload: 1 024 rays (32 x 32
    lattice of the 334x512 view) x (64 coarse + 128 fine) evaluations, all host threads; the per-frame setup is timed
    separately and NOT charged to the rays (a full view would amortise it over 171 008 rays)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    n_side = 32
    rps, sec, n, setup = cpu_port(n_side, True, steps=max(1, args.steps), warmup=min(args.warmup, 1),
                                  layout="bvv" if args.workload == "C" else "narrow")
    sample = (f"{n} rays ({n_side}x{n_side} lattice of the 334x512 view), 64 coarse + 128 fine evaluations/ray, V=3; per-frame setup "
              f"({setup:.2f} s) timed separately and not included")
    line = {"impl": "reference", "metric": "rays_per_s", "value": rps, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(args, "cpu"), "setup_s": setup,
            "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config_dict(args, where, workload=None):
    wl = workload or args.workload
    return {"workload": f"{'C: vanerf_bvv big-view-variation' if wl == 'C' else 'B: vanerf.json'} full 334x512 novel view, "
                        f"171008 rays, V=3 source views, 64 coarse + 128 fine evaluations/ray (fine=True, uniform=True)",
            "precision": args.precision, "rays_per_view": H * W, "source_views": V, "samples": [S_C, S_C + S_F],
            "ray_partition": "interleaved pixels across ranks (vanerf_b200.dist.render_view) + NCCL all_gather of output tiles" if args.gpus > 1 else "single GPU",
            "l2": "L2 flushed (256 MiB write) between timed iterations; per-step working set (gather records) exceeds L2",
            "where": where}


def dist_env():
    return int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))


def measure_dynamic(args, net, F, steps, warmup):
    """Workload D (BASELINE.json configs[3], render_dynamic): a sequence of F frames, each with its own mesh, source images
    and feature maps and its own 334x512 target camera on a 360-degree path.  A step = all F frames end to end from pinned
    host buffers: H2D of the frame's maps, per-frame setup (BVHs, vertex visibility, vertex tables, bf16 maps), render, D2H
    of the image.  Frames are dealt round-robin to the ranks (vanerf_b200.dynamic), no collective inside the timed path;
    per-GPU work shrinks with N ("strong" scaling over a fixed sequence).  Returns the result dict on rank 0."""
    import torch
    import torch.distributed as dist
    from vanerf_b200 import dynamic, synthetic
    world, rank, local = dist_env()
    dev = net.device
    F = max(F, world)
    ids = dynamic.frames_for_rank(F, rank, world)
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

    for _ in range(warmup):
        step()
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = net.renderer.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record()
        step()
        b.record()
    sync_all()
    net.renderer.finish()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    if rank != 0:
        return None
    clk = clocks.stop()
    rps = F * H * W / (ms_step * 1e-3)
    return {"metric": "rays_per_s", "value": rps, "unit": "rays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "ms_per_frame": ms_step / F, "frames": F, "frames_per_s": F / (ms_step * 1e-3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"D: render_dynamic, {F} frames x one 334x512 view (171008 rays, V=3, 64 coarse + 128 fine "
                                   f"evaluations/ray), per-frame mesh / maps / camera, per-frame setup included",
                       "precision": args.precision, "frame_partition": "round robin over ranks, no collective in the path",
                       "l2": "every frame brings new maps and records: the working set exceeds L2", "where": "gpu"},
            "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": dynamic.h2d_bytes(frames[ids[0]]) * F, "d2h_bytes_per_step": F * 8 * H * W * 4,
                    "note": "the timed step IS the end-to-end path (host buffers in, host images out)"},
            "gpu_launches": int(net.renderer.launches - l0), "clocks": clk}


def measure_train(args, steps, warmup, patch=64):
    """Workload E (BASELINE.json configs[4]): training step on a 4096-ray batch per GPU (one random 64x64 patch, `fine=True`,
    `uniform=False`, view dropout, density noise 0.01): forward + backward through the render path (vanerf_b200/train.py: our
    kernels for sampling / geometry / projection / gathers / compositing incl. their backward, cuBLAS for the dense layers of the
    unfused graph), L1 coarse (x1) + L1 fine (x10) against a random target, ONE flat-bucket gradient all-reduce over NCCL, Adam
    (lr 1e-3) on the render-path parameters.  Weak scaling: every rank trains on its own patch.  `value` = rays/s with the frame
    resident; e2e = the same step from pinned host buffers (H2D of the maps, per-frame setup, step, D2H of the loss)."""
    import torch
    import torch.distributed as dist
    from vanerf_b200 import synthetic, weights
    from vanerf_b200 import train as T
    world, rank, local = dist_env()
    dev = torch.device("cuda", local)
    inp_host = synthetic.to_torch(synthetic.make_scene(H, W, V))
    path = T.TrainableRenderPath(weights.init_state_dict(H, W, mode="ref"), dev, rand_noise_std=0.01)
    opt = torch.optim.Adam(path.parameters(), lr=1e-3)
    n_params = sum(p.numel() for p in path.parameters())
    pin = lambda x: x.contiguous().pin_memory()
    h = dict(img=pin(inp_host["img"]), feat_tex=pin(inp_host["feat_tex"]), g0=pin(inp_host["feat_geo"][0]), g1=pin(inp_host["feat_geo"][1]),
             fg=pin(inp_host["src_foreground_mask"].to(torch.uint8)))
    msk = inp_host["src_foreground_mask"][0, 0, 0].bool()              # patch centres: foreground of source view 0 (stand-in for the target mask)
    rand = T.TrainRandom(1000 + rank, 2000 + rank)
    target = torch.rand(patch * patch, 3, generator=torch.Generator().manual_seed(rank)).to(dev)
    cfg = dict(training=True, uniform=False, fine=True, S_c=S_C, S_f=S_F)

    def frame_dev():
        f = dict(inp_host)
        f["img"], f["feat_tex"] = h["img"].to(dev, non_blocking=True), h["feat_tex"].to(dev, non_blocking=True)
        f["feat_geo"] = [h["g0"].to(dev, non_blocking=True), h["g1"].to(dev, non_blocking=True)]
        f["src_foreground_mask"] = h["fg"].to(dev, non_blocking=True).bool()
        return f

    frame = frame_dev()
    path.set_frame(frame)
    loss_host = torch.zeros(1).pin_memory()

    def step(e2e):
        nonlocal frame
        if e2e:
            frame = frame_dev()
            path.set_frame(frame)
        pix = T.patch_pixels(msk, W, H, patch, patch, rand)
        loss, _ = T.training_step(path, frame, pix, target, opt, rand, world, **cfg)
        if e2e:
            loss_host.copy_(loss.reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return loss

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = {}
    for mode in ("resident", "e2e", "tf32"):
        # "tf32": the same resident step with the dense layers of the training graph on TF32 tensor cores (opt-in, T.matmul_precision)
        ctx = T.matmul_precision("tf32" if mode == "tf32" else "fp32")
        ctx.__enter__()
        for _ in range(warmup):
            step(mode == "e2e")
        sync_all()
        l0 = path.renderer.launches
        torch.cuda.reset_peak_memory_stats(dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            a.record()
            loss = step(mode == "e2e")
            b.record()
        sync_all()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = (float(t.item()) / steps, path.renderer.launches - l0, float(loss))
        ctx.__exit__(None, None, None)
    if rank != 0:
        return None
    ms, launches, loss = res["resident"]
    ms_e, _, _ = res["e2e"]
    rays = world * patch * patch
    n_samples = patch * patch * (S_C + S_C + S_F)
    flops = 3.0 * FLOP_PER_SAMPLE(V) * n_samples                       # forward + backward (dX and dW) of the dense layers, per rank
    pk = peaks()
    h2d = sum(x.numel() * x.element_size() for x in h.values())
    return {"metric": "rays_per_s", "value": rays / (ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"E: training step, {patch}x{patch} random patch = {patch * patch} rays per GPU, V=3, 64 coarse + 128 fine evaluations/ray, "
                                   "uniform=False (stratified jitter), view dropout, density noise 0.01, L1 coarse + 10 L1 fine, backward through the "
                                   "render path, one flat-bucket NCCL all-reduce of the gradients, Adam lr 1e-3",
                       "parameters": n_params, "allreduce_bytes": 4 * n_params, "where": "gpu"},
            "e2e": {"value": rays / (ms_e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "loss": loss, "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
            "tf32_variant": {"ms_per_step": res["tf32"][0], "value": rays / (res["tf32"][0] * 1e-3), "unit": "rays/s",
                             "note": "NOT the headline: dense layers of the training graph on TF32 tensor cores (vanerf_b200.train.matmul_precision('tf32')); "
                                     "the headline and the gradient-parity tests run exact fp32"},
            "roofline": {"kernel": "dense layers of the unfused training graph (cuBLAS fp32 GEMMs, library)", "bound": "tensor",
                         "achieved": flops / (ms * 1e-3) / 1e12, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / pk["tf_sust"],
                         "traffic": None, "note": "whole step time; fp32 SGEMM cannot reach the bf16 tensor peak it is divided by"}}


def run_train(args):
    import torch
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure_train(args, args.steps, max(1, min(args.warmup, 2)))
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_dynamic(args):
    import torch
    import torch.distributed as dist
    from vanerf_b200 import weights
    from vanerf_b200.model import VANeRF
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = VANeRF(device=dev, precision=args.precision).eval()
    net.load_state_dict(weights.init_state_dict(H, W, mode="ref"))
    line = measure_dynamic(args, net, args.frames, args.steps, max(1, min(args.warmup, 2)))
    if rank == 0:
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
    # headline = the bf16-MLP tensor-core path (north star kernel 2); the fp32 path is measured next to it
    ap.add_argument("--precision", default=os.environ.get("VANERF_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--workload", default="B", choices=["B", "C", "D", "E"], help="headline workload of the JSON line (default B; C and D are "
                    "also measured as secondary blocks of the default run)")
    ap.add_argument("--frames", type=int, default=64, help="workload D: frames per step (BASELINE.json configs[3]: 64)")
    ap.add_argument("--no-fp32-path", action="store_true", help="skip the secondary fp32-path measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reuse-variant", action="store_true", help="skip the secondary coarse-reuse measurement")
    ap.add_argument("--no-secondary", action="store_true", help="skip the workload C / D blocks (developer runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "D":
        return run_dynamic(args)
    if args.workload == "E":
        return run_train(args)

    import torch
    import torch.distributed as dist
    from vanerf_b200 import _lib as L
    from vanerf_b200 import dist as D
    from vanerf_b200 import synthetic, weights
    from vanerf_b200.model import VANeRF

    # stdout carries exactly one JSON line: NCCL's own banner / debug output goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W_steps = max(3, args.warmup)
    prec = L.FP32 if args.precision == "fp32" else L.BF16
    sd = weights.init_state_dict(H, W, mode="ref")
    net = VANeRF(device=dev, precision=args.precision).eval()
    net.load_state_dict(sd)
    r = net.renderer
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    R_local = D.partition_pixels(H, W, rank, world).shape[0]
    pk = peaks()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure_view(layout, precision, steps, warmup, e2e=True):
        """One workload (camera layout) on one precision path: (1) timed views with inputs resident, NO per-kernel events in the
        loop; (2) a second, instrumented pass of the same views for the per-kernel-class times (roofline); (3) end to end through
        VANeRF.batch_render_pifu_nerf with pinned host buffers.  Multi-GPU: vanerf_b200.dist.render_view (interleaved pixel
        partition, all_gather of tiles, image assembled in the reference's pixel order on every rank)."""
        inp_host = synthetic.to_torch(synthetic.make_scene(H, W, V, layout=layout))
        mv = lambda t: t.to(dev)
        inp = dict(inp_host)
        inp["img"], inp["feat_tex"], inp["src_foreground_mask"] = mv(inp_host["img"]), mv(inp_host["feat_tex"]), mv(inp_host["src_foreground_mask"])
        inp["feat_geo"] = [mv(t) for t in inp_host["feat_geo"]]
        r.set_frame(inp["img"], inp["cam_in"], inp["targets"], inp["sp_data"], inp["feat_geo"], inp["feat_tex"], inp["src_foreground_mask"])
        tar = r.make_target(inp["cam_tar"], inp["bounds"])
        step = lambda: D.render_view(r, tar, H, W, rank, world, S_C, S_F, True, precision)
        for _ in range(warmup):
            step()
        sync_all()
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        launches0 = r.launches
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            flush.fill_(1)
            a.record()
            step()
            b.record()
        sync_all()
        r.finish()
        launches = r.launches - launches0
        ms_step = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev)) / steps
        clk = clocks.stop() if rank == 0 else None
        need = torch.tensor([1 if (rank == 0 and not clk["samples"]) else 0], device=dev)
        if world > 1:
            dist.broadcast(need, 0)
        if int(need.item()):
            # the timed region was shorter than nvidia-smi's sampling period (many GPUs, short views): sample the clocks over an extra,
            # untimed repetition of the very same steps
            if rank == 0:
                clocks = ClockSampler(local)
                clocks.start()
            t_end = time.perf_counter() + 1.0
            go = torch.ones(1, device=dev)
            while int(go.item()):
                for _ in range(4):
                    step()
                torch.cuda.synchronize()
                go.fill_(1 if time.perf_counter() < t_end else 0)
                if world > 1:
                    dist.broadcast(go, 0)
            if rank == 0:
                clk = clocks.stop()
                clk["sampled"] = "extra untimed repetition of the same steps (timed region shorter than the sampling period)"
        # ---- instrumented pass (per-kernel-class CUDA events on the launching stream); not the headline
        r.timing(True)
        r.timing_read(reset=True)
        n_prof = min(2, steps)
        for _ in range(n_prof):
            flush.fill_(1)
            step()
        sync_all()
        ktimes = r.timing_read(reset=True)
        r.timing(False)
        res = {"ms_per_view": ms_step, "value": H * W / (ms_step * 1e-3), "unit": "rays/s", "steps": steps, "warmup": warmup,
               "gpu_launches": int(launches), "clocks": clk, "kernel_ms_per_view": {k: v[0] / n_prof for k, v in ktimes.items()},
               "_ktimes": ktimes, "_n_prof": n_prof}
        if not e2e:
            return res
        # ---- end to end through the public API, host buffers
        pin = lambda x: x.contiguous().pin_memory()
        h_img, h_tex, h_fg = pin(inp_host["img"]), pin(inp_host["feat_tex"]), pin(inp_host["src_foreground_mask"].to(torch.uint8))
        h_g0, h_g1 = pin(inp_host["feat_geo"][0]), pin(inp_host["feat_geo"][1])
        h_out = torch.empty(H, W, 8).pin_memory() if rank == 0 else None
        pix = torch.from_numpy(D.partition_pixels(H, W, rank, world)).to(dev)
        rows_pad = D.padded_tile_rows(H, W, world)
        tiles = [torch.empty(rows_pad, 8, device=dev) for _ in range(world)] if world > 1 else None

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
                tile = rows.new_zeros((rows_pad, 8))
                tile[: rows.shape[0]] = rows
                dist.all_gather(tiles, tile)
                full = D.assemble(tiles, H, W, world)            # the image a user gets: reference pixel order
            else:
                full = rows.view(H, W, 8)
            if rank == 0:
                h_out.copy_(full, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(2):
            step_e2e()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        sync_all()
        r.finish()
        sec = max_over_ranks((time.perf_counter() - t0) / steps)
        h2d = sum(x.numel() * x.element_size() for x in (h_img, h_tex, h_fg, h_g0, h_g1))
        res["e2e"] = {"value": H * W / sec, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": H * W * 8 * 4, "ms_per_view": 1e3 * sec}
        res["_frame"] = (inp, tar)
        return res

    def rooflines(res, precision):
        ktimes, n_prof = res["_ktimes"], res["_n_prof"]
        n_samples = R_local * (S_C + S_C + S_F) * n_prof
        mlp_ms, mlp_n = ktimes["mlp"]
        gat_ms, gat_n = ktimes["gather"]
        geo_ms, geo_n = ktimes["geom"]
        ach_tf = FLOP_PER_SAMPLE(V) * n_samples / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
        e = 4 if precision == L.FP32 else 2                  # fp32 maps vs bf16 maps / vertex tables
        ach_gb = GATHER_BYTES_PER_SAMPLE(V, e) * n_samples / (gat_ms * 1e-3) / 1e9 if gat_ms > 0 else 0.0
        n_queries = R_local * (S_C + S_F) * n_prof           # mesh queries actually made (geometry reuse: coarse depths once)
        name = "fp32" if precision == L.FP32 else "bf16"
        return {
            "roofline": {"kernel": "k_mlp_tc<SPLIT> (tcgen05, split precision: bf16 hi/lo operands, 3 MMAs per K step, fp32 accumulate)" if precision == L.FP32 else "k_mlp_tc (tcgen05)",
                         "bound": "tensor", "achieved": ach_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach_tf / pk["tf_sust"],
                         "traffic": ncu_traffic(name), "peak_source": pk["src"] + " bf16 sustained (cuBLAS, seconds-long loop)",
                         "launches": int(mlp_n), "avg_launch_ms": mlp_ms / max(1, mlp_n), "algorithmic_flop_per_sample": FLOP_PER_SAMPLE(V),
                         "timed_by": "CUDA events around the kernel's launches in a separate instrumented pass of the same views"},
            "roofline_gather": {"kernel": "k_gather + k_rec_split" if precision == L.FP32 else "k_gather_tc", "bound": "hbm", "achieved": ach_gb, "peak": pk["hbm"],
                                "unit": "GB/s", "frac": ach_gb / pk["hbm"], "launches": int(gat_n), "avg_launch_ms": gat_ms / max(1, gat_n),
                                "algorithmic_bytes_per_sample": GATHER_BYTES_PER_SAMPLE(V, e)},
            "geometry": {"kernel": "k_geom_query", "ms_per_view": geo_ms / n_prof, "queries_per_view": n_queries // n_prof,
                         "queries_per_s": n_queries / (geo_ms * 1e-3) if geo_ms > 0 else 0.0,
                         "note": "exact closest face + inside parity + nearest vertex per query; no roofline in the north star"},
        }

    pub = lambda d: {k: v for k, v in d.items() if not k.startswith("_")}
    # ================= headline: workload B (or C with --workload C) on the selected precision path
    main_res = measure_view("bvv" if args.workload == "C" else "narrow", prec, args.steps, W_steps)
    line = None
    if rank == 0:
        line = {"metric": "rays_per_s", "value": main_res["value"], "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": W_steps,
                "ms_per_step": main_res["ms_per_view"], "ms_per_view": main_res["ms_per_view"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic", "config": config_dict(args, "gpu"),
                "e2e": main_res["e2e"], "gpu_launches": main_res["gpu_launches"], "kernel_ms_per_step": main_res["kernel_ms_per_view"],
                "clocks": main_res["clocks"]}
        line.update(rooflines(main_res, prec))

    # ================= the other precision path of BASELINE.json configs[1] (N = 1): >= 3 timed views, own roofline block
    if world == 1 and args.precision == "bf16" and not args.no_fp32_path:
        fres = measure_view("narrow", L.FP32, max(3, min(args.steps, 3)), 1, e2e=False)
        blk = pub(fres)
        blk.update(rooflines(fres, L.FP32))
        blk["note"] = ("fp32 path = split-precision tensor-core kernel (operands as bf16 hi + lo, 3 MMAs per K step, fp32 epilogues; 1e-3 bar), "
                       "inputs resident; the roofline counts the ALGORITHMIC FLOPs once (the 3x MMA work is overhead) against the same measured peaks")
        line["fp32_path"] = blk

    # ================= coarse reuse (vanerf_set_reuse_coarse): same output bits, 64 + 64 instead of 64 + 128 evaluations per ray
    if world == 1 and not args.no_reuse_variant:
        inp, tar = main_res["_frame"]
        r.set_frame(inp["img"], inp["cam_in"], inp["targets"], inp["sp_data"], inp["feat_geo"], inp["feat_tex"], inp["src_foreground_mask"])
        pix = torch.from_numpy(D.partition_pixels(H, W, 0, 1)).to(dev)
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
        line["coarse_reuse"] = {"ms_per_view": ms_r, "value": H * W / (ms_r * 1e-3), "unit": "rays/s", "evaluations_per_ray": S_C + S_F,
                                "note": "NOT the headline: fine pass evaluates only the 64 new depths and reuses the coarse pass for the 64 coarse "
                                        "depths of the merged set (bit-identical output, tests: *_coarse_reuse_is_bit_identical); inputs resident"}

    # ================= secondary workloads of BASELINE.json: C (vanerf_bvv layout) and D (render_dynamic, 64 frames)
    if not args.no_secondary:
        if args.workload == "B":
            cres = measure_view("bvv", prec, max(2, min(args.steps, 3)), 2)
            if rank == 0:
                blk = pub(cres)
                blk.update(rooflines(cres, prec))
                blk["config"] = config_dict(args, "gpu", "C")
                line["workload_C"] = blk
        dres = measure_dynamic(args, net, args.frames, 1, 1)
        if rank == 0:
            line["workload_D"] = dres
        if world == 1:
            # the step in front of the path: source images -> feature maps, cuDNN through vanerf_b200/encoders.py
            enc = net.build_encoders(seed=1)
            ims = torch.rand(V, 3, H, W, device=dev)
            blk = {}
            for tag, dt in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
                enc.autocast_dtype = dt
                for _ in range(2):
                    enc.encode_geo(ims); enc.encode_tex(ims)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    g = enc.encode_geo(ims); t = enc.encode_tex(ims)
                e1.record()
                torch.cuda.synchronize()
                blk["ms_per_frame_" + tag] = e0.elapsed_time(e1) / 3
            blk["maps"] = {"geo0": list(g[0].shape), "geo1": list(g[1].shape), "tex": list(t.shape)}
            blk["note"] = "HGFilterV2 + ResBlkEncoder (28.3 M parameters) on V=3 source images of 334x512, channels_last, cuDNN: library code, outside the hot path"
            line["encoders"] = blk
            net.encoders = None
        del net, r
        torch.cuda.empty_cache()
        eres = measure_train(args, 2, 1)
        if rank == 0:
            line["workload_E"] = eres

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            rps, sec, n, setup = cpu_port(32, True)
            line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{n} rays (32x32 lattice of the 334x512 view), 64 coarse + 128 fine evaluations/ray, V=3: {sec:.1f} s; "
                                              f"per-frame setup {setup:.2f} s timed separately, not included"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
