#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native NeRF hot path.

Workload (BASELINE.json configs[1]): BARF — coarse-to-fine positional-encoding mask + SE(3)
camera-pose refinement on a synthetic 400x400 Blender-shaped scene, 4096 rays x 128 samples per
GPU and step (weak scaling: rays shard across GPUs, one NCCL gradient all-reduce per step).
A "step" = pose transform -> uniform sampling -> fused PE+MLP forward -> compositing -> MSE ->
compositing backward -> fused MLP backward (data + weight gradients, pose gradients) ->
all-reduce -> fused Adam.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                       the reference's CPU PyTorch
                                                               arithmetic (oracle port) on the
                                                               host cores, bounded sample

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 4096
SAMPLES = 128
IMAGE_SIZE = 400
N_IMAGES = 20
NEAR, FAR = 2.0, 8.0
METRIC = "train rays/s (fwd+bwd)"
WORKLOAD = ("BARF c2f PE mask + SE(3) pose refinement, synthetic 400x400 SDF scene, "
            "4096 rays x 128 samples per GPU, NerfModel 4x256x2seg, PE 10/4 + identity")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.02)     # nvidia-smi itself takes ~30 ms: ~20 samples per second of load

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# model construction (run_barf.py:151-196 shape)
# --------------------------------------------------------------------------------------------
def build_model(n_images: int):
    import torch as th
    from nerf_experiments_b200 import model_interpolation as mi
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import positional_encodings as pe
    from nerf_experiments_b200.model_camera_extrinsics import CameraExtrinsics
    th.manual_seed(1337)
    ep = pe.BarfPositionalEncoding(10, 0.0, 0.5, 2.5, True, 1.0)
    ed = pe.BarfPositionalEncoding(4, 0.0, 0.5, 2.5, True, 1.0)
    net = arch.NerfModel(4, 256, True, False, 2, ep, ed, 5e-4, 1e-5, 200000)
    model = mi.NerfInterpolation(NEAR, FAR, net, SAMPLES, "equidistant", -1.0, "middle", None, 0)
    cam = CameraExtrinsics(n_images, 1e-3, 1e-5, 200000)
    model.camera_extrinsics = cam
    model.param_groups = model.param_groups + cam.param_groups
    # mid-schedule coarse-to-fine mask: partly open, so the masked path is what is timed
    ep.update_alpha(1.25)
    ed.update_alpha(1.25)
    return model


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch as th
    import torch.distributed as dist
    from nerf_experiments_b200 import _lib, scene
    from nerf_experiments_b200.engine import TrainEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=th.device("cuda", local))
    dev = th.device("cuda", local)
    th.cuda.set_device(dev)
    _lib.lib()   # fail loudly if the CUDA library is missing

    sc = scene.make_scene(N_IMAGES, IMAGE_SIZE, IMAGE_SIZE, dev, rotation_noise=0.15, translation_noise=0.15)
    model = build_model(N_IMAGES)
    eng = TrainEngine(model, dev)
    field = model.model_radiance.fused_field()

    K, W = args.steps, args.warmup
    g = th.Generator(device=dev).manual_seed(1000 + rank)
    n_batches = K + W
    idx = th.randint(0, sc.n_rays, (n_batches, RAYS_PER_GPU), device=dev, generator=g)
    batches = [sc.batch(idx[i]) for i in range(n_batches)]          # resident in HBM
    host_batches = [tuple(t.cpu().pin_memory() for t in b) for b in batches[:K]]

    def barrier():
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()

    for i in range(W):
        eng.step(*batches[i])
    barrier()

    # ---- device-resident timed region -----------------------------------------------------
    field.timers = {}
    launches0 = _lib.launch_count()
    # clocks / throttle reasons are sampled from here to the end of the end-to-end region below: both
    # timed regions (and nothing but the few milliseconds between them) run under the sampler
    clocks = ClockSampler(local)
    clocks.__enter__()
    barrier()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = eng.step(*batches[W + i])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    timers = field.timers
    field.timers = None
    t_ms = th.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_total = float(t_ms.item())

    # ---- end-to-end: host buffers in, loss out, every step ------------------------------------
    # through the public host-batch API (engine.HostStepper): every step copies its inputs from
    # pinned host memory (on a copy stream, overlapping the previous step) and copies its loss
    # back; the host reads each loss one step late so that it never waits on the step in flight.
    from nerf_experiments_b200.engine import HostStepper
    stepper = HostStepper(eng)
    barrier()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    losses = []
    for i in range(K):
        prev = stepper.submit(host_batches[i])
        if prev is not None:
            losses.append(prev)
    losses.append(stepper.flush())                                  # D2H read of the last step's loss
    e1.record()
    barrier()
    clocks.__exit__(None, None, None)
    last = losses[-1]
    assert len(losses) == K
    t_e2e = th.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_e2e.item())
    h2d = sum(t.numel() * t.element_size() for t in host_batches[0])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    n_samples = RAYS_PER_GPU * SAMPLES
    macs = field.macs_per_sample()
    kernels = {}
    for name, evs in timers.items():
        tms = sum(a.elapsed_time(b) for a, b in evs) / max(len(evs), 1)
        key = {"mlp_fwd_train": "fwd", "mlp_bwd_inputs": "bwd_inputs", "mlp_bwd": "bwd", "mlp_wgrad": "wgrad"}[name]
        flops = 2.0 * macs[key] * n_samples
        kernels[name] = {"ms": tms, "tflops": flops / (tms * 1e-3) / 1e12, "flops_per_launch": flops}
    top = max(kernels, key=lambda k: kernels[k]["ms"])
    peak = peaks["bf16_tflops_sustained"]
    peak_hbm = peaks["hbm_gbs"]
    # algorithmic HBM bytes per launch: the forward writes the activation stash
    # and the sign bits, the backward reads the sign bits and writes the dY stash, the weight-
    # gradient kernel reads both stashes once
    n_tiles = (n_samples + 127) // 128
    slab = 128 * 128
    x_bytes = n_tiles * field.compiled.stash_slabs_per_tile * slab
    m_bytes = n_tiles * field.compiled.mask_words_per_tile * 128 * 4
    dy_bytes = n_tiles * field.bwd[True].dy_slabs_per_tile * slab
    io_bytes = n_samples * 16
    alg_bytes = {"mlp_fwd_train": x_bytes + m_bytes + io_bytes, "mlp_bwd_inputs": dy_bytes + m_bytes + 2 * io_bytes,
                 "mlp_bwd": dy_bytes + m_bytes + 2 * io_bytes, "mlp_wgrad": x_bytes + dy_bytes}
    for k, v in kernels.items():
        v["gbs"] = alg_bytes[k] / (v["ms"] * 1e-3) / 1e9
        v["bytes_per_launch"] = alg_bytes[k]
    # DRAM traffic of the same kernel per launch, from the committed `ncu --set full` capture of
    # this workload (newest profiles/r*_traffic.json; dram__bytes_read.sum + dram__bytes_write.sum)
    traffic = None
    try:
        import glob
        newest = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*_traffic.json")))[-1]
        prof = json.load(open(newest))
        kname = {"mlp_fwd_train": "mlp_fwd_kernel", "mlp_bwd_inputs": "mlp_bwd_kernel", "mlp_bwd": "mlp_bwd_kernel",
                 "mlp_wgrad": "mlp_wgrad_kernel"}[top]
        rec = prof["kernels"][kname]
        traffic = rec["dram_bytes_read"] + rec["dram_bytes_write"]
    except Exception:  # noqa: BLE001 - the capture is optional
        traffic = None
    # the dominant kernel is reported against the roof it sits closer to: the fused forward /
    # backward are tensor-bound by design, the weight-gradient kernel streams 128 FLOP per byte
    tf_frac, hbm_frac = kernels[top]["tflops"] / peak, kernels[top]["gbs"] / peak_hbm
    if hbm_frac > tf_frac:
        roofline = {"kernel": top, "bound": "hbm", "achieved": round(kernels[top]["gbs"], 1), "peak": peak_hbm,
                    "unit": "GB/s", "frac": round(hbm_frac, 4), "traffic": traffic,
                    "peak_source": f"{peaks['source']} HBM copy bandwidth"}
    else:
        roofline = {"kernel": top, "bound": "tensor", "achieved": round(kernels[top]["tflops"], 2), "peak": peak,
                    "unit": "TFLOP/s", "frac": round(tf_frac, 4), "traffic": traffic,
                    "peak_source": f"{peaks['source']} bf16 sustained (kernel timed inside a long step)"}
    roofline["kernels"] = {k: {"ms": round(v["ms"], 4), "tflops": round(v["tflops"], 2),
                               "tensor_frac": round(v["tflops"] / peak, 4), "gbs": round(v["gbs"], 1),
                               "hbm_frac": round(v["gbs"] / peak_hbm, 4)} for k, v in kernels.items()}
    roofline["mlp_share_of_step"] = round(sum(v["ms"] for v in kernels.values()) / (ms_total / K), 4)

    cpu = cpu_baseline_sample(steps=2, rays=256) if world == 1 else None   # N=1 only (tier rule)
    rays_total = world * RAYS_PER_GPU * K
    out = {
        "metric": METRIC, "value": rays_total / (ms_total * 1e-3), "unit": "rays/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_gpu": RAYS_PER_GPU, "samples_per_ray": SAMPLES,
                   "global_rays_per_step": world * RAYS_PER_GPU, "parallelism": f"dp{world}",
                   "l2": "per-step working set (activation stash ~2.9 GB) exceeds the 126 MB L2",
                   "accumulate": "bf16 operands, fp32 accumulate, fp32 master weights"},
        "e2e": {"value": rays_total / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K},
        "gpu_launches": int(launches), "loss": float(loss.item()), "loss_e2e": last,
        "clocks": clocks.summary(), "roofline": roofline, "cpu_baseline": cpu,
    }
    emit_json(out)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's arithmetic (oracle port) on the host cores
# --------------------------------------------------------------------------------------------
def cpu_baseline_sample(steps: int, rays: int, warmup: int = 1):
    """Times the oracle restatement of the same training step (fwd + loss + bwd, fp32, torch CPU)
    on `rays` rays x 128 samples per step with every host thread."""
    import torch as th
    from oracle import ref_step
    threads = os.cpu_count() or 1
    th.set_num_threads(threads)
    th.manual_seed(1337)
    sd, cfg, pe_cfg = oracle_net()
    g = th.Generator().manual_seed(5)
    rot = (th.randn((N_IMAGES, 3), generator=g) * 0.01).requires_grad_()
    tr = (th.randn((N_IMAGES, 3), generator=g) * 0.01).requires_grad_()
    times = []
    for s in range(warmup + steps):
        o = th.nn.functional.normalize(th.randn((rays, 3), generator=g), dim=1) * 4.0
        d = th.nn.functional.normalize(-o + 0.3 * th.randn((rays, 3), generator=g), dim=1)
        target = th.rand((rays, 3), generator=g)
        idx = th.randint(0, N_IMAGES, (rays,), generator=g)
        u = {"offset": th.rand((rays, 1), generator=g)}
        t0 = time.perf_counter()
        loss, _ = ref_step.barf_step(sd, cfg, pe_cfg, rot, tr, idx, o, d, target, NEAR, FAR, SAMPLES, "middle", u)
        loss.backward()
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        for v in list(sd.values()) + [rot, tr]:
            v.grad = None
    best = min(times)
    return {"value": rays / best, "unit": "rays/s", "cores": threads, "kind": "port",
            "sample": f"{rays} rays x {SAMPLES} samples per step, best of {steps} after {warmup} warm-up, "
                      f"fwd+loss+bwd (no optimizer), fp32 torch CPU", "s_per_step": best}


def oracle_net():
    import torch as th
    import torch.nn as nn

    def lin(i, o):
        m = nn.Linear(i, o)
        return m.weight.detach().clone().requires_grad_(), m.bias.detach().clone().requires_grad_()

    sd = {}
    for seg, d_in, d_out in ((0, 63, 256), (1, 319, 257)):
        dims = [(d_in, 256), (256, 256), (256, 256), (256, 256), (256, d_out)]
        for k, (i, o) in enumerate(dims):
            sd[f"model_segments.{seg}.{2 * k}.weight"], sd[f"model_segments.{seg}.{2 * k}.bias"] = lin(i, o)
    sd["model_color.0.weight"], sd["model_color.0.bias"] = lin(283, 128)
    sd["model_color.2.weight"], sd["model_color.2.bias"] = lin(128, 3)
    cfg = dict(n_hidden=4, n_segments=2, delayed_direction=True, delayed_density=False)
    pe_cfg = dict(pos_levels=10, dir_levels=4, scale=1.0, identity=True, alpha_pos=th.tensor(3.75),
                  alpha_dir=th.tensor(1.5))
    return sd, cfg, pe_cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    rays = 512
    cpu = cpu_baseline_sample(steps=max(args.steps, 1), rays=rays, warmup=max(args.warmup, 1))
    out = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "rays/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["s_per_step"] * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "note": "reference arithmetic (oracle port of the PyTorch path) on the "
                      "host cores; each step is a bounded sample of the workload"},
           "cpu_baseline": cpu,
           "e2e": {"value": cpu["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_json(out)


_STDOUT_FD = None


def capture_stdout():
    """stdout carries the one JSON line and nothing else: whatever libraries print there (NCCL's
    version banner under NCCL_DEBUG, for one) is diverted to stderr at file-descriptor level."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _STDOUT_FD is None:
        os.write(1, line)
    else:
        os.write(_STDOUT_FD, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    # defaults: 100 timed steps of ~3.3 ms for our arm (enough for ~10 clock samples), 3 steps of the
    # bounded CPU sample for the reference arm
    if args.steps is None:
        args.steps = 100 if args.impl == "ours" else 3
    if args.warmup is None:
        args.warmup = 10 if args.impl == "ours" else 1
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched directly: re-launch one rank per GPU (the driver uses the same torchrun line)
        import socket
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--impl", args.impl]
        sys.exit(subprocess.call(cmd))
    capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
