#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native NeRF hot path.

Workload (BASELINE.json configs[1]): BARF — coarse-to-fine positional-encoding mask + SE(3)
camera-pose refinement on a synthetic 400x400 Blender-shaped scene, 4096 rays x 128 samples per
GPU and step (weak scaling: rays shard across GPUs, one NCCL gradient all-reduce per step).
A "step" = the reference's BarfModel.training_step + optimiser (barf/model_barf.py:29-92): per-step
coarse-to-fine alpha update, blurred targets from the image pyramid, pose transform -> uniform
sampling -> fused PE+MLP forward -> compositing -> MSE / PSNR / per-step Kabsch pose error ->
compositing backward -> fused MLP backward (data + weight gradients, pose gradients) ->
all-reduce -> fused Adam with the device-side schedule — replayed as ONE CUDA graph per step.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                       the reference's own CPU implementation
                                                               (unmodified barf/*.py from oracle/_ref)
                                                               on the host cores, same configuration

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
BLUR_SIGMAS = [4.0, 1.41, 0.5, 0.0]          # barf/run_barf.py:48-53 with start_blur_sigma = 4, n_blur_sigmas = 4
ALPHA_EPOCHS = (-0.5, 1.5)                   # coarse-to-fine ramp: alpha ~ 2.5 .. 3.3 over the timed steps
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


def build_barf_model(sc, n_batches: int):
    """The reference's BarfModel as barf/run_barf.py:151-196 builds it (PE 10/4 + identity under the
    coarse-to-fine mask, 4x256x2-segment NerfModel, equidistant sampling with offset -1, blur schedule
    from alpha, pose refinement), wired to the GPU-resident ray batcher of the synthetic scene. The alpha
    ramp is placed so that the mask is partly open and MOVING over the timed steps (the schedule update
    of every step is part of what is timed)."""
    import torch as th
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import positional_encodings as pe
    from nerf_experiments_b200.model_camera_calibration import BarfModel, LoopState
    th.manual_seed(1337)
    ep = pe.BarfPositionalEncoding(10, 0.0, ALPHA_EPOCHS[0], ALPHA_EPOCHS[1], True, 1.0)
    ed = pe.BarfPositionalEncoding(4, 0.0, ALPHA_EPOCHS[0], ALPHA_EPOCHS[1], True, 1.0)
    net = arch.NerfModel(4, 256, True, False, 2, ep, ed, 5e-4, 1e-5, 200000)
    model = BarfModel(n_training_images=sc.n_images, camera_learning_rate_start=1e-3, camera_learning_rate_stop=1e-5,
                      camera_learning_rate_decay_end=200000, near_sphere_normalized=NEAR, far_sphere_normalized=FAR,
                      model_radiance=net, samples_per_ray_radiance=SAMPLES, samples_per_ray_proposal=0,
                      max_gaussian_sigma=BLUR_SIGMAS[0], uniform_sampling_strategy="equidistant",
                      uniform_sampling_offset_size=-1.0)
    model.loop = LoopState(sc.batcher, n_batches)
    return model


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch as th
    import torch.distributed as dist
    from nerf_experiments_b200 import _lib, scene
    from nerf_experiments_b200.engine import HostStepper, TrainEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=th.device("cuda", local))
    dev = th.device("cuda", local)
    th.cuda.set_device(dev)
    _lib.lib()   # fail loudly if the CUDA library is missing

    sc = scene.make_scene(N_IMAGES, IMAGE_SIZE, IMAGE_SIZE, dev, rotation_noise=0.15, translation_noise=0.15,
                          blur_sigmas=BLUR_SIGMAS)
    n_epoch_batches = len(sc.batcher) // (RAYS_PER_GPU * world)
    model = build_barf_model(sc, n_epoch_batches)
    eng = TrainEngine(model, dev, loss_fn=model.training_loss)
    field = model.model_radiance.fused_field()

    K, W = args.steps, args.warmup
    N_PROFILE = 5                                   # eager steps with per-kernel CUDA events (roofline)
    g = th.Generator(device=dev).manual_seed(1000 + rank)
    n_batches = K + W + N_PROFILE
    idx = th.randint(0, len(sc.batcher), (n_batches, RAYS_PER_GPU), device=dev, generator=g)
    # the reference's 7-tuple (o_raw, o_noisy, d_raw, d_noisy, blur pyramid (B, n_sigmas, 3), img_idx, pixel_width)
    batches = [sc.batcher.batch(idx[i]) for i in range(n_batches)]          # resident in HBM
    host_batches = [tuple(t.cpu().pin_memory() for t in b) for b in batches[:K]]

    def barrier():
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()

    # ---- warm-up (eager), per-kernel timing (eager, CUDA events around each MLP kernel), capture ----
    for i in range(W):
        eng.step(*batches[i])
    barrier()
    field.timers = {}
    for i in range(N_PROFILE):
        eng.step(*batches[W + i])
    barrier()
    timers = field.timers
    field.timers = None
    eng.capture(*batches[0])                         # the whole step incl. the NCCL all-reduce as ONE graph
    for i in range(3):
        eng.replay(*batches[i])
    # device-resident batches in the layout of the graph's inputs: a step is one device copy + one graph launch
    packed = [eng.pack_batch(batches[W + N_PROFILE + i]) for i in range(K)]
    barrier()
    # warm-up of the measured path itself, right in front of the timed region: W replays of the captured step
    # (the GPUs idled while the host packed the batches; at N = 8 the first replays after that gap ran ~7 %
    # slower than the same replays later in the run)
    for i in range(max(W, 3)):
        eng.replay_packed(packed[i % K])
    barrier()

    # ---- device-resident timed region: K graph replays -------------------------------------------
    # clocks / throttle reasons are sampled from here to the end of the end-to-end region below
    clocks = ClockSampler(local)
    clocks.__enter__()
    barrier()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = eng.replay_packed(packed[i])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    loss_value = float(loss.item())
    logs = {k: float(v) for k, v in eng.last_logs.items()}
    launches = int(eng.launches_per_replay) * K
    t_ms = th.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_total = float(t_ms.item())

    # ---- end-to-end: host buffers in, loss out, every step -----------------------------------------
    # through the public host-batch API (engine.HostStepper): every step copies its 7-tuple from pinned
    # host memory (copy stream, overlapping earlier steps), replays the captured step and copies its
    # loss back; the host reads each loss `depth` steps late so that it never waits on a step in flight.
    stepper = HostStepper(eng, depth=2, use_graph=True)
    barrier()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    losses = []
    for i in range(K):
        prev = stepper.submit(host_batches[i])
        if prev is not None:
            losses.append(prev)
    losses += stepper.drain()                                       # D2H read of the last steps' losses
    e1.record()
    barrier()
    clocks.__exit__(None, None, None)
    assert len(losses) == K
    t_e2e = th.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_e2e.item())
    h2d = sum(t.numel() * t.element_size() for t in host_batches[0])

    eng.release_graph()          # before the process group goes: the graph holds the NCCL communicator
    if rank != 0:
        shutdown(world)
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
    # DRAM traffic of the same kernel per launch: NOT measured in this run (that needs ncu) — taken from
    # the committed `ncu --set full` capture of this workload (newest profiles/r*_traffic.json), and labelled so
    traffic, traffic_source = None, None
    try:
        import glob
        kname = {"mlp_fwd_train": "mlp_fwd_kernel", "mlp_bwd_inputs": "mlp_bwd_kernel", "mlp_bwd": "mlp_bwd_kernel",
                 "mlp_wgrad": "mlp_wgrad_kernel"}[top]
        # the newest capture of THIS workload (the GARF captures under the same naming hold a list of launches)
        cands = [f for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
                 if isinstance(json.load(open(f)).get("kernels"), dict)]
        newest = cands[-1]
        prof = json.load(open(newest))
        rec = prof["kernels"][kname]
        traffic = rec["dram_bytes_read"] + rec["dram_bytes_write"]
        traffic_source = "profiles/" + os.path.basename(newest) + " (ncu --set full capture, not this run)"
    except Exception:  # noqa: BLE001 - the capture is optional
        traffic = None
    # the dominant kernel is reported against the roof it sits closer to: the fused forward /
    # backward are tensor-bound by design, the weight-gradient kernel streams 128 FLOP per byte
    tf_frac, hbm_frac = kernels[top]["tflops"] / peak, kernels[top]["gbs"] / peak_hbm
    if hbm_frac > tf_frac:
        roofline = {"kernel": top, "bound": "hbm", "achieved": round(kernels[top]["gbs"], 1), "peak": peak_hbm,
                    "unit": "GB/s", "frac": round(hbm_frac, 4), "traffic": traffic, "traffic_source": traffic_source,
                    "peak_source": f"{peaks['source']} HBM copy bandwidth"}
    else:
        roofline = {"kernel": top, "bound": "tensor", "achieved": round(kernels[top]["tflops"], 2), "peak": peak,
                    "unit": "TFLOP/s", "frac": round(tf_frac, 4), "traffic": traffic, "traffic_source": traffic_source,
                    "peak_source": f"{peaks['source']} bf16 sustained (kernel timed inside a long step)"}
    roofline["kernels"] = {k: {"ms": round(v["ms"], 4), "tflops": round(v["tflops"], 2),
                               "tensor_frac": round(v["tflops"] / peak, 4), "gbs": round(v["gbs"], 1),
                               "hbm_frac": round(v["gbs"] / peak_hbm, 4)} for k, v in kernels.items()}
    roofline["kernel_times_from"] = f"{N_PROFILE} eager steps of the same workload before the timed region (CUDA events per launch)"
    roofline["mlp_share_of_step"] = round(sum(v["ms"] for v in kernels.values()) / (ms_total / K), 4)

    extras = None
    if world == 1 and not args.no_extras:
        try:
            extras = run_extras(dev, peaks)
        except Exception as exc:  # noqa: BLE001 - the extras never take the headline down with them
            extras = {"error": repr(exc)[:300]}
    cpu = cpu_baseline_sample(steps=1, rays=RAYS_PER_GPU) if world == 1 else None   # N=1 only (tier rule)
    rays_total = world * RAYS_PER_GPU * K
    out = {
        "metric": METRIC, "value": rays_total / (ms_total * 1e-3), "unit": "rays/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_gpu": RAYS_PER_GPU, "samples_per_ray": SAMPLES,
                   "global_rays_per_step": world * RAYS_PER_GPU, "parallelism": f"dp{world}",
                   "step": "BarfModel.training_step semantics: per-step alpha update, blurred targets from the "
                           "pyramid, pose transform, render, MSE, PSNR, per-step Kabsch pose error, backward, "
                           "all-reduce, fused Adam — ONE CUDA graph per step (engine.capture / replay)",
                   "l2": "per-step working set (activation stash ~2.9 GB) exceeds the 126 MB L2",
                   "accumulate": "bf16 operands, fp32 accumulate, fp32 master weights"},
        "e2e": {"value": rays_total / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K,
                "path": "engine.HostStepper(depth=2, use_graph=True): pinned host 7-tuples in, loss out"},
        "gpu_launches": launches, "gpu_launches_per_step": int(eng.launches_per_replay),
        "loss": loss_value, "loss_e2e": losses[-1], "logs": logs,
        "clocks": clocks.summary(), "roofline": roofline, "cpu_baseline": cpu, "extras": extras,
    }
    emit_json(out)
    shutdown(world)


def shutdown(world: int):
    """Tears the process group down; never lets a stuck NCCL teardown hold the job (the result is out)."""
    if world <= 1:
        return
    import torch.distributed as dist
    timer = threading.Timer(20.0, lambda: os._exit(0))
    timer.daemon = True
    timer.start()
    try:
        dist.destroy_process_group()
    finally:
        timer.cancel()


# --------------------------------------------------------------------------------------------
# the rest of BASELINE.json's metric on the same box (N = 1): the other configurations and the
# memory-bound kernels at render scale, as extra keys of the same JSON line
# --------------------------------------------------------------------------------------------
def _timeit(fn, iters=10, warm=3):
    import torch as th
    for _ in range(warm):
        fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _rays(B, dev, n_images=20, seed=0):
    import torch as th
    g = th.Generator().manual_seed(seed)
    o = th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0
    d = th.nn.functional.normalize(-o + 0.3 * th.randn((B, 3), generator=g), dim=1)
    return (o.to(dev), d.to(dev), th.rand((B, 3), generator=g).to(dev),
            th.randint(0, n_images, (B,), generator=g).int().to(dev), th.full((B, 1), 1 / 555.0, device=dev))


def run_extras(dev, peaks):
    import torch as th
    from nerf_experiments_b200 import model_interpolation as mi
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import ops
    from nerf_experiments_b200 import positional_encodings as pe
    from nerf_experiments_b200.engine import TrainEngine
    out = {"note": "CUDA events, 3 warm-up + 10 timed iterations each, one B200, synthetic rays"}

    # C1 vanilla NeRF 1024 x 64 (run_vanilla_as_barf.py shape), one CUDA graph per step
    th.manual_seed(1337)
    net = arch.NerfModel(4, 256, True, False, 2, pe.BarfPositionalEncoding(10, 10.0, 0.0, 0.0, False, 1.0),
                         pe.BarfPositionalEncoding(4, 4.0, 0.0, 0.0, False, 1.0), 5e-4, 1e-5, 200000)
    model = mi.NerfInterpolation(2.0, 8.0, net, 64, "stratified_uniform", -1.0, "middle", None, 0)
    eng = TrainEngine(model, dev)
    o, d, tgt, idx, pw = _rays(1024, dev)
    eng.step(o, d, tgt, None, pw)
    eng.capture(o, d, tgt, None, pw)
    ms = _timeit(lambda: eng.replay(o, d, tgt, None, pw))
    out["c1_vanilla_1024x64"] = {"ms_per_step": round(ms, 4), "rays_per_s": round(1024 / ms * 1e3), "path": "CUDA graph"}
    del eng, model, net

    # C3 Mip-BARF 8192 rays x (64 + 256): integrated PE, one shared network, pose refinement
    from nerf_experiments_b200.model_mip import MipBarf
    th.manual_seed(1337)
    ep = pe.IntegratedFourierFeatures(levels=10, include_identity=True, scale=1., distribute_variance=False)
    ed = pe.BarfPositionalEncoding(0, 1, 0, 1, True)
    net = arch.NerfModel(4, 256, True, False, 2, ep, ed, 5e-4, 1e-5, 200000)
    model = MipBarf(model_radiance=net, samples_per_ray_radiance=256, n_training_images=20,
                    camera_learning_rate_start=1e-3, camera_learning_rate_stop=1e-5, camera_learning_rate_decay_end=200000,
                    uniform_sampling_strategy="equidistant", uniform_sampling_offset_size=-1., samples_per_ray_proposal=64,
                    sigma_decay_start_step=0, sigma_decay_end_step=100000, start_blur_sigma=8., start_pixel_width_sigma=1.5)
    eng = TrainEngine(model, dev)
    o, d, tgt, idx, pw = _rays(8192, dev)
    ms = _timeit(lambda: eng.step(o, d, tgt, idx, pw, coarse_weight=0.1), iters=5)
    out["c3_mipbarf_8192x(64+256)"] = {"ms_per_step": round(ms, 4), "rays_per_s": round(8192 / ms * 1e3)}
    del eng, model, net

    # C4 GARF (garf/main.py shape): Gaussian-activation radiance + proposal network, inverse-CDF resampling
    try:
        from nerf_experiments_b200.model_garf import GarfModel, garf_engine
        for B in (1024, 4096):
            th.manual_seed(1337)
            m = GarfModel(2.0, 7.0, 64, 192, 0.5, 1.5, 1.0, 1e-3, 1e-4, 100000, 0.0, 1e-3, 1e-4, 100000, 0.0).to(dev)
            m.train()
            eng = garf_engine(m, dev)
            o, d, tgt, idx, pw = _rays(B, dev)
            ms = _timeit(lambda: eng.step(o, d, tgt), iters=5)
            out[f"c4_garf_{B}x(64+192)"] = {"ms_per_step": round(ms, 4), "rays_per_s": round(B / ms * 1e3)}
            del eng, m
    except Exception as exc:  # noqa: BLE001
        out["c4_garf"] = {"error": repr(exc)[:200]}

    # C5 800 x 800 render in 16384-ray chunks (C2 network, no gradient)
    from nerf_experiments_b200.ray_batcher import render_image
    model = build_model(20).to(dev)
    o, d, _, _, _ = _rays(640000, dev)
    ms = _timeit(lambda: render_image(model, o, d, 800, 800, 1 / 555.0, chunk=16384), iters=3, warm=1)
    out["c5_render_800x800_128spp"] = {"ms_per_image": round(ms, 3), "rays_per_s": round(640000 / ms * 1e3)}
    del model, o, d

    # compositing / resampling at render scale: achieved HBM GB/s on algorithmic bytes
    hbm = peaks["hbm_gbs"]
    Br, Sr = 262144, 128
    sigma = th.nn.functional.softplus(th.randn((Br, Sr), device=dev))
    delta = th.full((Br, Sr), 6.0 / Sr, device=dev)
    rgb = th.rand((Br, Sr, 3), device=dev)
    g_rgb, g_w = th.randn((Br, 3), device=dev), th.randn((Br, Sr), device=dev)
    ms_f = _timeit(lambda: ops.composite_fwd(sigma, delta, rgb))
    ms_b = _timeit(lambda: ops.composite_bwd(sigma, delta, rgb, g_rgb, g_w))
    bf, bb = Br * Sr * 24 + Br * 12, Br * Sr * 40 + Br * 12
    out["composite_fwd_262144x128"] = {"ms": round(ms_f, 4), "GBps": round(bf / ms_f / 1e6, 1), "hbm_frac": round(bf / ms_f / 1e6 / hbm, 4)}
    out["composite_bwd_262144x128"] = {"ms": round(ms_b, 4), "GBps": round(bb / ms_b / 1e6, 1), "hbm_frac": round(bb / ms_b / 1e6 / hbm, 4)}
    del sigma, delta, rgb, g_rgb, g_w
    Br = 524288
    tc0, tc1 = ops.sample_uniform(2.0, 8.0, Br, 64, dev, None, th.rand((Br, 1), device=dev), -1.0)
    w = th.rand((Br, 64), device=dev) ** 4
    ms_a = _timeit(lambda: ops.resample_alloc(tc0, w, tc1 - tc0, 256, 2.0, 8.0))
    ba = Br * (3 * 4 * 64 + 2 * 4 * 256)
    out["resample_alloc_64to256"] = {"ms": round(ms_a, 4), "GBps": round(ba / ms_a / 1e6, 1), "hbm_frac": round(ba / ms_a / 1e6 / hbm, 4)}
    edges = th.linspace(0, 1, 65, device=dev).repeat(Br, 1)
    cdf = th.cat((th.zeros(Br, 1, device=dev), th.cumsum(w, 1)), 1)
    cdf = cdf / cdf[:, -1:]
    u = th.rand((Br,), device=dev)
    ms_i = _timeit(lambda: ops.resample_icdf(edges, cdf, 192, u))
    bi = Br * (4 * 65 * 2 + 4 + 4 * 193)
    out["resample_icdf_64to192"] = {"ms": round(ms_i, 4), "GBps": round(bi / ms_i / 1e6, 1), "hbm_frac": round(bi / ms_i / 1e6 / hbm, 4)}
    del tc0, tc1, w, edges, cdf, u
    th.cuda.empty_cache()
    try:       # the unmodified reference on the same GPU (checker-side; absent when oracle/_ref was not vendored)
        ref_gpu = reference_on_gpu(dev, RAYS_PER_GPU)
        if ref_gpu is not None:
            out["c2_reference_modules_on_this_gpu"] = ref_gpu
    except Exception as e:                                   # never let the optional line break the bench
        out["c2_reference_modules_on_this_gpu"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    return out


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU implementation on the host cores
# --------------------------------------------------------------------------------------------
def cpu_baseline_sample(steps: int, rays: int, warmup: int = 1):
    """Times one training step (BarfModel.training_step + backward, fp32, torch CPU, every host
    thread) of the UNMODIFIED reference modules (oracle/_ref, kind "reference") on `rays` rays x 128
    samples of the bench workload; falls back to the oracle restatement (kind "port") when the
    vendored reference is absent."""
    import torch as th
    from oracle import ref_runner
    threads = os.cpu_count() or 1
    th.set_num_threads(threads)
    if ref_runner.reference_dir() is None:
        return cpu_baseline_port(steps, rays, warmup)
    ref = ref_runner.load_reference()
    g = th.Generator().manual_seed(5)
    cam_o = th.nn.functional.normalize(th.randn((N_IMAGES, 3), generator=g), dim=1) * 4.0
    cam_on = cam_o + 0.15 * th.randn((N_IMAGES, 3), generator=g)
    n_epoch_batches = N_IMAGES * IMAGE_SIZE * IMAGE_SIZE // rays
    model = ref_runner.build_barf(ref, N_IMAGES, SAMPLES, NEAR, FAR, n_epoch_batches, cam_o, cam_on, BLUR_SIGMAS,
                                  BLUR_SIGMAS[0], alpha_epochs=ALPHA_EPOCHS)
    times = []
    for s in range(warmup + steps):
        o = th.nn.functional.normalize(th.randn((rays, 3), generator=g), dim=1) * 4.0
        d = th.nn.functional.normalize(-o + 0.3 * th.randn((rays, 3), generator=g), dim=1)
        colors = th.rand((rays, len(BLUR_SIGMAS), 3), generator=g)
        idx = th.randint(0, N_IMAGES, (rays,), generator=g)
        pw = th.full((rays,), 1 / 555.0)
        t0 = time.perf_counter()
        ref_runner.training_step(model, (o, o, d, d, colors, idx, pw), 100 + s)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    best = min(times)
    return {"value": rays / best, "unit": "rays/s", "cores": threads, "kind": "reference",
            "sample": f"{rays} rays x {SAMPLES} samples per step, best of {steps} after {warmup} warm-up: the unmodified "
                      f"reference BarfModel.training_step + backward (no optimizer step), fp32 torch CPU, oracle/_ref",
            "s_per_step": best}


def reference_on_gpu(dev, rays: int, steps: int = 5, warmup: int = 2):
    """SURVEY.md section 8(d), optional line: the UNMODIFIED reference modules (oracle/_ref) moved to the same
    B200 — eager PyTorch with TF32 matmuls as barf/run_barf.py:101 sets them — timed on the bench workload's
    step (BarfModel.training_step + backward + torch Adam, eps 1e-5). A checker-side number: the product
    path never touches it."""
    import torch as th
    from oracle import ref_runner
    if ref_runner.reference_dir() is None:
        return None
    th.set_float32_matmul_precision("high")
    ref = ref_runner.load_reference()
    g = th.Generator().manual_seed(5)
    cam_o = th.nn.functional.normalize(th.randn((N_IMAGES, 3), generator=g), dim=1) * 4.0
    cam_on = cam_o + 0.15 * th.randn((N_IMAGES, 3), generator=g)
    n_epoch_batches = N_IMAGES * IMAGE_SIZE * IMAGE_SIZE // rays
    model = ref_runner.build_barf(ref, N_IMAGES, SAMPLES, NEAR, FAR, n_epoch_batches, cam_o.to(dev), cam_on.to(dev),
                                  BLUR_SIGMAS, BLUR_SIGMAS[0], alpha_epochs=ALPHA_EPOCHS).to(dev)
    opt = th.optim.Adam(model.parameters(), lr=5e-4, eps=1e-5)
    o = (th.nn.functional.normalize(th.randn((rays, 3), generator=g), dim=1) * 4.0).to(dev)
    d = th.nn.functional.normalize(-o.cpu() + 0.3 * th.randn((rays, 3), generator=g), dim=1).to(dev)
    colors = th.rand((rays, len(BLUR_SIGMAS), 3), generator=g).to(dev)
    idx = th.randint(0, N_IMAGES, (rays,), generator=g).to(dev)
    pw = th.full((rays,), 1 / 555.0, device=dev)

    def step(s):
        opt.zero_grad(set_to_none=True)
        loss = model.training_step((o, o, d, d, colors, idx, pw), 100 + s)
        loss.backward()
        opt.step()
    for s in range(warmup):
        step(s)
    th.cuda.synchronize(dev)
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        step(warmup + s)
    e1.record()
    th.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": round(ms, 3), "rays_per_s": round(rays / ms * 1e3),
            "what": f"unmodified reference BarfModel.training_step + backward + torch Adam on this GPU, eager PyTorch, "
                    f"TF32 matmuls, {rays} rays x {SAMPLES} samples (oracle/_ref; CUDA events, {steps} steps after {warmup})"}


def cpu_baseline_port(steps: int, rays: int, warmup: int = 1):
    """The oracle restatement of the same training step (fwd + loss + bwd, fp32, torch CPU)."""
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
    # the SAME configuration as our arm: 4096 rays x 128 samples per step (a step takes ~2.5 s on 16 cores)
    cpu = cpu_baseline_sample(steps=max(min(args.steps, 3), 1), rays=RAYS_PER_GPU, warmup=max(min(args.warmup, 1), 1))
    out = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "rays/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["s_per_step"] * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "rays_per_gpu": RAYS_PER_GPU, "samples_per_ray": SAMPLES,
                      "note": "the reference's own CPU implementation (unmodified barf/*.py from oracle/_ref, "
                              "BarfModel.training_step + backward) on the host cores of rank 0; one process "
                              "whatever N is (the reference is single-process)"},
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
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configurations / micro-benchmarks")
    args = ap.parse_args()
    # defaults: 400 timed steps of ~3.5 ms for our arm in each of the two timed regions (~3 s under the
    # clock sampler, whose nvidia-smi calls take ~0.3 s each), 3 steps for the reference arm
    if args.steps is None:
        args.steps = 400 if args.impl == "ours" else 3
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
