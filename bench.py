#!/usr/bin/env python
"""Benchmark of the Defense-GAN projection loop (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision fp16|fp32]
                    [--config mnist|fmnist|celeba] [--batch B --rec_rr R --rec_iters L]

A "step" is one pass of the hot path over one batch of synthetic images: one
`gan.reconstruct` call = R restarts x L momentum-GD steps of generator forward + MSE +
backward-to-z, then arg-min select.  At N=1 the workload is BASELINE.json configs[1]
(MNIST 28x28, R=10, L=200, batch=256 on one B200); for N>1 every rank gets the same per-GPU
batch (weak scaling), the image axis is sharded with no data-path collective and one NCCL
all-gather of the reconstructions ends each step.

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM, timed with CUDA
events; `e2e` = the same through the public Python API with pinned HOST buffers (H2D of the
images and D2H of the reconstructions inside the timed region); `roofline` = the dominant
kernel's algorithmic FLOP/s (CUDA events per launch, separate untimed-for-throughput pass)
against the measured bf16 tensor peak in MEASURED_PEAKS.json; `cpu_baseline` = the oracle port
of the reference's TF1 CPU path on this box's host cores (bounded sample).
`--impl reference` times only that CPU port (the reference itself cannot run: no TF1/py2).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

CONFIGS = {
    # name: (dataset, default batch per GPU, R, L)
    "mnist": ("mnist", 256, 10, 200),     # BASELINE.json configs[1]: the configuration the metric is quoted on
    "fmnist": ("f-mnist", 256, 10, 200),  # configs[2]
    "celeba": ("celeba", 128, 10, 200),   # configs[3]
}
FALLBACK_PEAKS = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


class ClockSampler:
    """nvidia-smi sampler running DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [c for c, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    """Threads the CPU port actually uses: the per-step tensors are small (tens of rows), so the
    torch CPU kernels stop scaling (and then slow down) well before a 100+-core box is full."""
    return int(os.environ.get("DGAN_CPU_THREADS", min(os.cpu_count() or 1, 16)))


def cpu_port_images_per_sec(dataset, R, L, sample_images, threads, repeats=1):
    """The oracle port of the reference's TF1 CPU path (oracle/defensegan_oracle.py), fp32, all
    host threads, on a bounded sample of the same workload: `sample_images` images at the full
    R and L.  Returns (images/s, seconds per call)."""
    from oracle import defensegan_oracle as O
    torch.set_num_threads(threads)
    arch = O.canonical_arch(dataset)
    w = O.init_generator_weights(arch)
    imgs = O.synthetic_images(arch, w, sample_images)
    z0 = O.sample_z0(sample_images * R, 128)
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.reconstruct(arch, w, imgs, R, L, z_init_val=z0)
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    return sample_images / t, t


def run_reference_arm(args, rank, world, out):
    """--impl reference: the reference's own CPU implementation of the path cannot run here
    (Python 2 + TensorFlow 1.7, neither present nor installable offline) => the oracle port is
    timed on the host cores, rank 0 only."""
    if rank != 0:
        return
    dataset, B, R, L = resolve_workload(args)
    threads = host_threads()
    sample = max(1, args.ref_sample)
    for _ in range(args.warmup):
        cpu_port_images_per_sec(dataset, R, L, sample, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_images_per_sec(dataset, R, L, sample, threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "reconstructed images/sec at R=%d,L=%d" % (R, L), "value": value,
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(dataset, B, R, L), "dataset": dataset, "rec_rr": R, "rec_iters": L,
                   "per_step_sample_images": sample},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": "%d images per step at full R=%d, L=%d (oracle restatement of the TF1 CPU path; "
                                   "TF1/py2 reference is not runnable offline)" % (sample, R, L)},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.emit(json.dumps(line))


def resolve_workload(args):
    dataset, B, R, L = CONFIGS[args.config]
    if args.batch:
        B = args.batch
    if args.rec_rr:
        R = args.rec_rr
    if args.rec_iters:
        L = args.rec_iters
    return dataset, B, R, L


def workload_name(dataset, B, R, L):
    return "%s %s generator projection, batch=%d/GPU, R=%d, L=%d" % (
        dataset, "64x64x3" if dataset == "celeba" else "28x28x1", B, R, L)


class _OnlyJsonOnStdout:
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner to fd 1) must not
    leak into it: while active, fd 1 points at stderr; emit() writes to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self._real, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._real, 1)
        os.close(self._real)
        return False


def main():
    with _OnlyJsonOnStdout() as out:
        _main(out)


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--config", default="mnist", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--rec_rr", type=int, default=0)
    ap.add_argument("--rec_iters", type=int, default=0)
    ap.add_argument("--ref_sample", type=int, default=4, help="images per step of the CPU reference arm")
    ap.add_argument("--cpu_sample", type=int, default=4, help="images of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--no_profile", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world, out)
        return
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    if args.steps < 1 or args.warmup < 3:
        print("note: the timing rules ask for >= 3 warm-up steps", file=sys.stderr)

    import torch.distributed as dist
    from defensegan_b200.models.gan import dataset_gan_dict
    from defensegan_b200.parallel import reconstruct_sharded

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    dataset, B, R, L = resolve_workload(args)
    gan = dataset_gan_dict[dataset](test_mode=True, verbose=False, precision=args.precision, batch_size=R * 5)
    gan.rec_rr, gan.rec_iters, gan.rec_lr = R, L, 10.0
    hwc = int(np.prod(gan.image_dim))
    B_global = B * world

    # synthetic inputs (on-manifold + noise), generated ON DEVICE by the native generator
    g = torch.Generator(device="cpu").manual_seed(1990)
    zstar = torch.randn(B_global, gan.latent_dim, generator=g) * (1.0 / gan.latent_dim) ** 0.5
    eps = torch.randn(B_global, *gan.image_dim, generator=g)
    lo = -1.0 if dataset == "celeba" else 0.0
    x_full = (gan.generator_fn(zstar.to(dev)) + 0.1 * eps.to(dev)).clamp_(lo, 1.0).contiguous()
    z0_full = (torch.randn(B_global * R, gan.latent_dim, generator=g) * (1.0 / gan.latent_dim) ** 0.5).to(dev)
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def one_step():
        if distributed:
            return reconstruct_sharded(gan, x_full, z_init_val=z0_full)
        return gan.reconstruct(x_full, z_init_val=z0_full)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-timed throughput (`value`) ----------------------------------------------------
    for _ in range(args.warmup):
        flush.zero_()
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (inside the bracket)
        rec = one_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = gan._native.last_launch_count
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = B_global * args.steps / (ms_max / 1000.0)

    # ---- end-to-end through the public API with HOST buffers (`e2e`) -----------------------------
    lo_i, hi_i = rank * B, (rank + 1) * B
    x_host = x_full[lo_i:hi_i].cpu().pin_memory()
    z0_loc = z0_full[lo_i * R:hi_i * R].contiguous()
    out_host = torch.empty_like(x_host).pin_memory()

    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)          # H2D of this step's inputs (pinned)
        r = gan.reconstruct(xd, z_init_val=z0_loc)      # the call a user makes
        out_host.copy_(r, non_blocking=True)            # D2H of the step's result
        return r

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        flush.zero_()
        e2e_step()
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = B_global * args.steps / (float(te.item()) / 1000.0)

    # ---- per-kernel timing pass for the roofline (rank 0, not part of `value`) ---------------------
    peaks = load_peaks()
    roofline, kernels = None, None
    if rank == 0 and not args.no_profile:
        nat = gan._native
        x_loc = x_full[:B].contiguous()
        nat.profile_enable(True)
        gan.reconstruct(x_loc, z_init_val=z0_full[:B * R].contiguous())
        torch.cuda.synchronize(dev)
        prof = nat.profile_read()
        nat.profile_enable(False)
        tot_ms = sum(k["ms"] for k in prof) or 1.0
        kernels = []
        for k in prof:
            if k["launches"] == 0:
                continue
            avg_ms = k["ms"] / k["launches"]
            tf = k["flops_per_launch"] / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
            kernels.append({"kernel": k["name"], "launches": k["launches"], "avg_us": round(avg_ms * 1e3, 2),
                            "share": round(k["ms"] / tot_ms, 4), "tflops": round(tf, 2)})
        dom = max(kernels, key=lambda k: k["share"])
        peak = peaks["bf16_tflops"]        # kernel timed alone -> burst figure (fp16 and bf16 share kind::f16 rate)
        if args.precision == "fp32":
            peak = None
        traffic = None   # DRAM read+write bytes per launch of that kernel from the committed ncu --set full capture
        tp = os.path.join(ROOT, "profiles", "r1d_dram_traffic.json")
        if os.path.exists(tp) and args.precision == "fp16" and dataset == "mnist" and B * R == 2560:
            with open(tp) as f:
                traffic = json.load(f)["kernels"].get(dom["kernel"], {}).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": dom["kernel"], "achieved": dom["tflops"], "peak": peak,
                    "unit": "TFLOP/s", "frac": (dom["tflops"] / peak) if peak else None, "traffic": traffic,
                    "traffic_unit": "bytes/launch (dram__bytes_read+write, profiles/r1d_dram_traffic.json)",
                    "peak_source": "%s cuBLAS bf16 burst (MEASURED_PEAKS.json)" % peaks["_source"],
                    "operand_format": args.precision}

    # ---- CPU baseline (rank 0, N=1 only, bounded sample) ----------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and args.cpu_sample > 0:
        threads = host_threads()
        v, secs = cpu_port_images_per_sec(dataset, R, L, args.cpu_sample, threads)
        cpu_baseline = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                        "sample": "%d images at full R=%d, L=%d in %.1f s (oracle restatement of the TF1 CPU path)"
                                  % (args.cpu_sample, R, L, secs)}

    if rank == 0:
        macs = gan._native.macs_per_row
        gflop_per_image = 4.0 * macs * R * L / 1e9        # 2 FLOP/MAC x (fwd + bwd-to-z)
        step_tflops = value * gflop_per_image / 1e3
        line = {
            "metric": "reconstructed images/sec at R=%d,L=%d" % (R, L), "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if args.precision == "fp16" else "f32", "data": "synthetic",
            "config": {"workload": workload_name(dataset, B, R, L), "dataset": dataset, "global_batch": B_global,
                       "rec_rr": R, "rec_iters": L, "rec_lr": 10.0, "precision": args.precision,
                       "accumulate": "f32", "parallelism": "image-shard x%d + 1 all-gather" % world,
                       "l2": "192 MiB memset between steps (inside the timed bracket)"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B_global * hwc * 4,
                    "d2h_bytes_per_step": B_global * hwc * 4},
            "gpu_launches": int(launches_per_step) * args.steps * world,
            "clocks": clocks,
            "algorithmic": {"gflop_per_image": gflop_per_image, "tflops_whole_step": step_tflops,
                            "frac_of_sustained_bf16_peak": step_tflops / (world * peaks["bf16_tflops_sustained"])},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "kernels": kernels,
        }
        out.emit(json.dumps(line))
    if distributed:
        dist.destroy_process_group()
    gan.close()


if __name__ == "__main__":
    main()
