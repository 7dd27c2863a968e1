"""Out-of-bounds / uninitialised-read check of the kernels without a sanitizer (compute-sanitizer is
closed on the GPU pool): every device buffer the Python wrappers allocate for a kernel (outputs,
stashes, workspaces, gradient buffers) is carved out of a larger arena with 64 KB guard bands on both
sides, filled with a byte pattern; buffers the wrappers allocate with `empty` are filled with 0xFF
(NaN in fp32 / bf16, -1 in the integer types) so that a kernel consuming memory it never wrote
poisons its result. After whole training steps through every kernel family the guard bands must be
untouched and the losses / parameters must equal those of the unguarded run."""
import contextlib

import pytest
import torch as th

pytestmark = pytest.mark.gpu

GUARD = 65536
PATTERN = 0xA5


class GuardedAllocations:
    def __init__(self):
        self.records = []
        self._orig = {}

    def _carve(self, shape, dtype, device, zero):
        shape = tuple(int(s) for s in shape)
        n = 1
        for s in shape:
            n *= s
        nbytes = n * th.empty((), dtype=dtype).element_size()
        padded = (nbytes + 255) // 256 * 256
        arena = self._orig["empty"](2 * GUARD + padded, dtype=th.uint8, device=device)
        arena.fill_(PATTERN)
        body = arena[GUARD:GUARD + nbytes]
        body.fill_(0 if zero else 0xFF)
        self.records.append((arena, nbytes))
        return body.view(dtype).view(shape)

    @staticmethod
    def _is_cuda(device):
        return device is not None and th.device(device).type == "cuda"

    def _factory(self, name, zero):
        orig = self._orig[name]

        def f(*size, dtype=None, device=None, **kw):
            if not self._is_cuda(device) or kw.get("out") is not None or kw.get("pin_memory"):
                return orig(*size, dtype=dtype, device=device, **kw)
            if len(size) == 1 and not isinstance(size[0], int):
                size = tuple(size[0])
            out = self._carve(size, dtype or th.get_default_dtype(), device, zero)
            if kw.get("requires_grad"):
                out.requires_grad_()
            return out
        return f

    def _like(self, name, zero):
        orig = self._orig[name]

        def f(t, dtype=None, device=None, **kw):
            dev = device if device is not None else t.device
            if not self._is_cuda(dev) or not t.is_contiguous():
                return orig(t, dtype=dtype, device=device, **kw)
            return self._carve(t.shape, dtype or t.dtype, dev, zero)
        return f

    def __enter__(self):
        for name in ("empty", "zeros", "empty_like", "zeros_like"):
            self._orig[name] = getattr(th, name)
        th.empty = self._factory("empty", False)
        th.zeros = self._factory("zeros", True)
        th.empty_like = self._like("empty_like", False)
        th.zeros_like = self._like("zeros_like", True)
        return self

    def __exit__(self, *exc):
        for name, f in self._orig.items():
            setattr(th, name, f)

    def check(self):
        th.cuda.synchronize()
        bad = []
        for k, (arena, nbytes) in enumerate(self.records):
            head_ok = (arena[:GUARD] == PATTERN).all()
            tail_ok = (arena[GUARD + nbytes:] == PATTERN).all()
            if not (bool(head_ok) and bool(tail_ok)):
                bad.append((k, nbytes, bool(head_ok), bool(tail_ok)))
        assert not bad, f"guard bands overwritten (record, bytes, head intact, tail intact): {bad[:8]}"
        return len(self.records)


def _nerf_case(cuda, guarded: bool):
    from nerf_experiments_b200 import model_interpolation as mi
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import positional_encodings as pe
    from nerf_experiments_b200.engine import TrainEngine
    from nerf_experiments_b200.model_camera_extrinsics import CameraExtrinsics
    g = th.Generator().manual_seed(1)
    B, n_img = 200, 4                                    # 200 rays x 24 / 48 samples: ragged last tile
    o = th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0
    d = th.nn.functional.normalize(-o + 0.3 * th.randn((B, 3), generator=g), dim=1)
    target = th.rand((B, 3), generator=g)
    idx = th.randint(0, n_img, (B,), generator=g).int()
    pw = th.full((B, 1), 1 / 555.0)
    ctx = GuardedAllocations() if guarded else contextlib.nullcontext()
    with ctx as ga:
        th.manual_seed(0)

        def net():
            ep = pe.BarfPositionalEncoding(10, 0.0, 1.0, 2.0, True, 1.0)
            ed = pe.BarfPositionalEncoding(4, 0.0, 1.0, 2.0, True, 1.0)
            m = arch.NerfModel(4, 256, True, False, 2, ep, ed, 5e-4, 1e-5, 1000)
            ep.alpha.fill_(7.25); ed.alpha.fill_(4.0)
            return m
        model = mi.NerfInterpolation(2.0, 8.0, net(), 48, "equidistant", 0.0, "middle", net(), 24)
        cam = CameraExtrinsics(n_img, 1e-3, 1e-5, 1000)
        model.camera_extrinsics = cam
        model.param_groups = model.param_groups + cam.param_groups
        model = model.to(cuda)
        eng = TrainEngine(model, cuda)
        losses = [float(eng.step(*(t.to(cuda) for t in (o, d, target, idx, pw)))) for _ in range(3)]
        rgb, depth, w = model.render(o.to(cuda), d.to(cuda), pw.to(cuda), 2.0, 8.0)
        n = ga.check() if guarded else 0
        return losses, eng.flat.flat.detach().clone(), rgb.clone(), n


def _garf_case(cuda, guarded: bool):
    from nerf_experiments_b200.model_garf import garf_engine
    from nerf_experiments_b200.model_garf_camera_calibration import CameraCalibrationModel
    g = th.Generator().manual_seed(2)
    B, n_img = 200, 4
    o = th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0
    d = th.nn.functional.normalize(-o + 0.3 * th.randn((B, 3), generator=g), dim=1)
    o_n = o + 0.05 * th.randn((B, 3), generator=g)
    d_n = th.nn.functional.normalize(d + 0.05 * th.randn((B, 3), generator=g), dim=1)
    target = th.rand((B, 3), generator=g)
    idx = th.randint(0, n_img, (B,), generator=g)
    u0, u1 = th.rand(B, generator=g), th.rand(B, generator=g)
    batch = tuple(t.to(cuda) for t in (o, o_n, d, d_n, target, idx, u0, u1))
    ctx = GuardedAllocations() if guarded else contextlib.nullcontext()
    with ctx as ga:
        th.manual_seed(5)
        m = CameraCalibrationModel(n_img, 1e-3, 1e-5, 40, 10, 2.0, 7.0, 16, 24, 0.5, 1.5, 2.0,
                                   1e-3, 1e-4, 50, 0.0, 2e-3, 1e-4, 60, 0.0).to(cuda)
        m.train()
        eng = garf_engine(m, cuda)
        losses = []
        for _ in range(3):
            eng.step(*batch)
            losses.append(float(eng.last_logs["loss_fine"] + eng.last_logs["train_proposal_loss"]))
        n = ga.check() if guarded else 0
        return losses, eng.flat.flat.detach().clone(), n


def test_nerf_step_kernels_stay_inside_their_buffers(cuda):
    plain = _nerf_case(cuda, False)
    guarded = _nerf_case(cuda, True)
    assert guarded[3] > 20                               # the wrappers' allocations did go through the arena
    assert all(l == l for l in guarded[0])               # no NaN: nothing read a buffer it had not written
    assert guarded[0] == pytest.approx(plain[0], rel=1e-4)
    assert th.isfinite(guarded[2]).all() and th.allclose(guarded[2], plain[2], atol=2e-3)
    diff = (guarded[1] - plain[1]).abs()
    assert float((diff > 2e-5).float().mean()) < 1e-3 and float(diff.max()) < 5e-3


def test_garf_step_kernels_stay_inside_their_buffers(cuda):
    plain = _garf_case(cuda, False)
    guarded = _garf_case(cuda, True)
    assert guarded[2] > 20
    assert all(l == l for l in guarded[0])
    assert guarded[0] == pytest.approx(plain[0], rel=2e-3)
    diff = (guarded[1] - plain[1]).abs()
    assert float((diff > 5e-5).float().mean()) < 2e-3 and float(diff.max()) < 2e-2
