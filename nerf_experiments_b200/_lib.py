"""ctypes binding of libnerfb200.so (the C ABI declared in include/nerfb200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``.  There is no CPU
fallback: if the shared object is missing, loading raises, and every wrapper raises
``RuntimeError`` with the library's error text on a non-zero status.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# NERFB200_LIB: another build of the same library (A/B timing of kernel experiments, scripts/ only)
LIB_PATH = os.environ.get("NERFB200_LIB") or os.path.join(_HERE, "libnerfb200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

# ---- mirrors of csrc/mlp.h -----------------------------------------------------------------
NB_MAX_OPS = 32
NB_MAX_CHUNKS = 12
NB_MAX_BLOCKS = 3
NB_TILE_ROWS = 128
NB_SLAB_BYTES = 128 * 128
NB_N_SLABS = 6
NB_RING_STAGE_BYTES = 256 * 128

EPI_RELU, EPI_LINEAR, EPI_LINEAR_SIGMA, EPI_RGB, EPI_RGB_SIGMA, EPI_RELU_SIGMA = range(6)
BEPI_MASK, BEPI_PLAIN, BEPI_PLAIN_SIGMA, BEPI_MASK_SIGMA, BEPI_NONE, BEPI_PEGRAD_POS, BEPI_PEGRAD_DIR = range(7)
PE_CANON_LEVELS, PE_CANON_COLS, PE_CANON_IDENTITY = 10, 64, 60
PE_IDENTITY, PE_FOURIER, PE_INTEGRATED = range(3)
COMPOSITE_BARF, COMPOSITE_NERFACC = 0, 1
ACT_GAUSS, ACT_SARF, ACT_GABOR = 0, 1, 2


class NbBlock(C.Structure):
    _fields_ = [("tmem_col", C.c_int16), ("n", C.c_int16), ("row0", C.c_int16), ("accum_in", C.c_int16)]


class NbOp(C.Structure):
    _fields_ = [
        ("n_chunks", C.c_int8), ("n_blocks", C.c_int8), ("epi", C.c_int8), ("out_chunks", C.c_int8),
        ("a_src", C.c_int8 * NB_MAX_CHUNKS), ("k16", C.c_int8 * NB_MAX_CHUNKS),
        ("blk_mask", C.c_int8 * NB_MAX_CHUNKS), ("n_sub", C.c_int8 * NB_MAX_CHUNKS),
        ("w_rows", C.c_int16 * NB_MAX_CHUNKS), ("w_off", C.c_int32 * NB_MAX_CHUNKS),
        ("blocks", NbBlock * NB_MAX_BLOCKS),
        ("bias_off", C.c_int32), ("stash_slab", C.c_int32), ("mask_word", C.c_int32),
        ("out_width", C.c_int32),
    ]


class NbProgram(C.Structure):
    _fields_ = [("n_ops", C.c_int32), ("stash_slabs_per_tile", C.c_int32),
                ("mask_words_per_tile", C.c_int32), ("n_slabs", C.c_int16), ("n_stages", C.c_int16),
                ("ops", NbOp * NB_MAX_OPS)]


class NbPeCfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("levels", C.c_int32), ("include_identity", C.c_int32),
                ("use_mask", C.c_int32), ("distribute_variance", C.c_int32), ("scale", C.c_float),
                ("pixel_width_sigma", C.c_float), ("slab", C.c_int32), ("stash_slab", C.c_int32),
                ("encode_before_op", C.c_int32)]


class NbMlpInputs(C.Structure):
    _fields_ = [("N", C.c_int64), ("S", C.c_int32), ("t_mode", C.c_int32),
                ("ray_o", C.c_void_p), ("ray_d", C.c_void_p), ("t_start", C.c_void_p),
                ("t_end", C.c_void_p), ("pixel_width", C.c_void_p), ("pos", C.c_void_p),
                ("dir", C.c_void_p), ("pixel_width_per_sample", C.c_int32), ("reserved", C.c_int32)]


class NbPackChunk(C.Structure):
    _fields_ = [("base", C.c_int64), ("row_stride", C.c_int32), ("col_stride", C.c_int32),
                ("n_rows", C.c_int32), ("n_cols", C.c_int32), ("rows_padded", C.c_int32),
                ("dst_off", C.c_int32), ("dst_row0", C.c_int32), ("dst_row_step", C.c_int32),
                ("img_rows", C.c_int32), ("reserved", C.c_int32)]

    def __init__(self, *args, **kw):
        kw.setdefault("dst_row0", 0)
        kw.setdefault("dst_row_step", 1)
        kw.setdefault("img_rows", 0)
        kw.setdefault("reserved", 0)
        super().__init__(*args, **kw)


class NbPackBias(C.Structure):
    _fields_ = [("base", C.c_int64), ("n", C.c_int32), ("n_padded", C.c_int32),
                ("dst_off", C.c_int32), ("kind", C.c_int32), ("stride", C.c_int32), ("reserved", C.c_int32)]

    def __init__(self, *args, **kw):
        kw.setdefault("kind", 0)
        kw.setdefault("stride", 1)
        kw.setdefault("reserved", 0)
        super().__init__(*args, **kw)


PACK_COPY, PACK_GAUSS = 0, 1


class NbWgradItem(C.Structure):
    _fields_ = [("tile_begin", C.c_int32), ("tile_end", C.c_int32), ("n_dy_slabs", C.c_int32),
                ("n_x_slabs", C.c_int32), ("dy_slab", C.c_int32), ("x_slab", C.c_int32),
                ("m_real", C.c_int32), ("n_real", C.c_int32), ("dst", C.c_int64), ("ld", C.c_int32),
                ("bias_dst", C.c_int32), ("mode", C.c_int32), ("coef_dst", C.c_int32),
                ("z_slab", C.c_int32), ("n_z_slabs", C.c_int32), ("z_first", C.c_int32), ("zbias_dst", C.c_int32),
                ("x2_slab", C.c_int32)]

    def __init__(self, *args, **kw):
        kw.setdefault("mode", 0)
        kw.setdefault("coef_dst", -1)
        kw.setdefault("z_slab", -1)
        kw.setdefault("n_z_slabs", 0)
        kw.setdefault("z_first", 0)
        kw.setdefault("zbias_dst", -1)
        kw.setdefault("x2_slab", -1)
        super().__init__(*args, **kw)


WGRAD_MMA, WGRAD_COLSUM = 0, 1


class NbGaussLayer(C.Structure):
    _fields_ = [("w_off", C.c_int64), ("b_off", C.c_int64), ("g_off", C.c_int64), ("in_f", C.c_int32), ("out_f", C.c_int32)]

# ---- mirrors of include/nerfb200_garf.h ------------------------------------------------------
NG_MAX_OPS, NG_MAX_CHUNKS, NG_N_SLABS, NG_GEN_COLS, NG_MAX_FLOATS = 32, 6, 6, 128, 7680
NG_MAX_PROGRAM_FLOATS = 6900   # NG_MAX_FLOATS minus the step / op tables the kernels keep in the same region
NG_STEP_NONE, NG_STEP_GEN, NG_STEP_ACT, NG_STEP_LINEAR, NG_STEP_RGB, NG_STEP_SIGMA = range(6)
NG_BSTEP_HEAD, NG_BSTEP_ACT, NG_BSTEP_PLAIN = 8, 9, 10
NG_F_SIGMA, NG_F_HOLD_SAVE, NG_F_HOLD_ADD, NG_F_DIRECT, NG_F_FIRST_LAYER = 1, 2, 4, 8, 16


class NgStep(C.Structure):
    _fields_ = [("kind", C.c_int8), ("wait_lag", C.c_int8), ("n_slabs", C.c_int8), ("out_slab", C.c_int8),
                ("res_slab", C.c_int8), ("skip_src", C.c_int8), ("flags", C.c_int8), ("reserved", C.c_int8),
                ("src_col", C.c_int16), ("sigma_col", C.c_int16), ("bias_off", C.c_int32), ("coef_off", C.c_int32),
                ("skip_off", C.c_int32), ("y_stash", C.c_int32), ("z_stash", C.c_int32), ("gen_col0", C.c_int32)]


class NgBlock(C.Structure):
    _fields_ = [("tmem_col", C.c_int16), ("n", C.c_int16), ("row0", C.c_int16), ("reserved", C.c_int16)]


class NgOp(C.Structure):
    _fields_ = [("n_chunks", C.c_int8), ("n_blocks", C.c_int8), ("accumulate", C.c_int8), ("early", C.c_int8),
                ("a_slab", C.c_int8 * NG_MAX_CHUNKS), ("k16", C.c_int8 * NG_MAX_CHUNKS), ("w_rows", C.c_int16),
                ("reserved2", C.c_int16), ("w_off", C.c_int32 * NG_MAX_CHUNKS), ("blocks", NgBlock * 2)]


class NgProgram(C.Structure):
    _fields_ = [("n_ops", C.c_int32), ("y_slabs_per_tile", C.c_int32), ("z_slabs_per_tile", C.c_int32),
                ("n_floats", C.c_int32), ("w1_off", C.c_int64), ("b1_off", C.c_int64), ("g1_off", C.c_int64),
                ("n1", C.c_int32), ("aux_pos_stash", C.c_int32), ("aux_dir_stash", C.c_int32),
                ("sigma_bias", C.c_float), ("ops", NgOp * NG_MAX_OPS), ("steps", NgStep * (NG_MAX_OPS + 1))]


class NbAdamGroup(C.Structure):
    _fields_ = [("begin", C.c_longlong), ("end", C.c_longlong), ("n_steps", C.c_longlong), ("lr0", C.c_float),
                ("log_factor", C.c_float), ("weight_decay", C.c_float), ("mode", C.c_int32)]


LR_LE_NICE, LR_EXPONENTIAL = 0, 1

_lib = None
ABI_VERSION = 6      # include/nerfb200.h: NERFB200_ABI_VERSION


def build(verbose: bool = False) -> str:
    """Compiles csrc/*.cu for sm_100a into libnerfb200.so (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libnerfb200.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        if _lib.nerfb200_abi_version() != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH} implements ABI version {_lib.nerfb200_abi_version()}, this package binds "
                               f"version {ABI_VERSION}: rebuild it (python -c 'import __graft_entry__ as g; g.build()')")
        _lib.nerfb200_last_error.restype = C.c_char_p
        _lib.nerfb200_launch_count.restype = C.c_longlong
        _declare(_lib)
    return _lib


def _declare(L):
    vp, i32, f64, f32 = C.c_void_p, C.c_int, C.c_double, C.c_float
    L.nerfb200_sample_uniform.argtypes = [f64, f64, i32, i32, vp, vp, f64, vp, vp, vp]
    L.nerfb200_composite_fwd.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]
    L.nerfb200_composite_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp]
    L.nerfb200_resample_alloc.argtypes = [vp, vp, vp, i32, i32, i32, f64, vp, vp, vp, vp, vp]
    L.nerfb200_resample_fallback.argtypes = [vp, f64, f64, i32, i32, vp, vp, vp, vp]
    L.nerfb200_resample_icdf.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp]
    L.nerfb200_pose_fwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
    L.nerfb200_pose_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp, vp]
    L.nerfb200_so3_to_SO3.argtypes = [vp, i32, vp, vp]
    L.nerfb200_mlp_pack.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp]
    L.nerfb200_mlp_fwd.argtypes = [vp, vp, vp, C.POINTER(NbMlpInputs), C.POINTER(NbPeCfg),
                                   C.POINTER(NbPeCfg), vp, vp, f32, vp, vp, vp, vp, i32, vp]
    L.nerfb200_mlp_fwd2.argtypes = [vp, vp, vp, C.POINTER(NbMlpInputs), C.POINTER(NbPeCfg),
                                    C.POINTER(NbPeCfg), vp, vp, f32, vp, vp, vp, vp, i32, i32, vp]
    L.nerfb200_mlp_bwd.argtypes = [vp, vp, C.POINTER(NbMlpInputs), C.POINTER(NbPeCfg), C.POINTER(NbPeCfg),
                                   vp, vp, vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, vp, vp, vp, vp, vp]
    L.nerfb200_mlp_wgrad.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, vp]
    L.nerfb200_mlp_workspace_bytes.argtypes = [vp, C.c_longlong, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.nerfb200_garf_workspace_bytes.argtypes = [vp, C.c_longlong, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.nerfb200_garf_fwd.argtypes = [vp, vp, vp, vp, C.POINTER(NbMlpInputs), vp, vp, vp, vp, vp]
    L.nerfb200_garf_bwd.argtypes = [vp, vp, vp, vp, C.POINTER(NbMlpInputs), vp, vp, vp, vp, vp, vp, i32,
                                    vp, vp, vp, vp, vp]
    L.nerfb200_adam_step.argtypes = [vp, vp, vp, vp, C.c_longlong, i32, vp, vp, vp, vp, f32, f32, f32,
                                     C.c_longlong, f32, vp]
    L.nerfb200_adam_step_dev.argtypes = [vp, vp, vp, vp, C.c_longlong, C.POINTER(NbAdamGroup), i32, f32, f32, f32,
                                         f32, vp, vp, vp]
    L.nerfb200_pe_fwd.argtypes = [C.POINTER(NbPeCfg), vp, vp, vp, vp, vp, vp, C.c_longlong, vp, vp]
    L.nerfb200_pe_bwd.argtypes = [C.POINTER(NbPeCfg), vp, vp, vp, vp, vp, vp, vp, C.c_longlong, vp, vp, vp]
    L.nerfb200_act_fwd.argtypes = [i32, vp, vp, vp, C.c_longlong, i32, vp, i32, vp]
    L.nerfb200_act_bwd.argtypes = [i32, vp, vp, vp, vp, C.c_longlong, i32, vp, vp, vp, vp, i32, vp]
    L.nerfb200_render_rays_workspace_bytes.argtypes = [i32, i32, C.POINTER(C.c_longlong)]
    L.nerfb200_render_rays.argtypes = [vp, vp, vp, i32, C.POINTER(NbPeCfg), C.POINTER(NbPeCfg), vp, vp, f32, vp, vp, vp,
                                       i32, i32, f32, f32, vp, vp, vp, vp, f32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.nerfb200_lindisp_intervals.argtypes = [vp, f32, f32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.nerfb200_trans_cdf_fwd.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp]
    L.nerfb200_trans_cdf_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp]
    L.nerfb200_prop_loss.argtypes = [vp, vp, vp, vp, i32, i32, i32, f32, f32, vp, vp, vp]
    L.nerfb200_gauss_width_grad.argtypes = [vp, i32, C.c_longlong, vp, vp, f32, vp]
    L.nerfb200_ray_batch.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, f32, f32, i32, i32, f32,
                                     vp, vp, vp, vp, vp, vp, vp, vp]
    L.nerfb200_kabsch.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("nerfb200_last_error", "nerfb200_launch_count"):
            fn.restype = C.c_int


# every symbol include/nerfb200.h declares (tests check that the library exports them all)
EXPORTS = [
    "nerfb200_last_error", "nerfb200_abi_version", "nerfb200_launch_count",
    "nerfb200_sample_uniform", "nerfb200_composite_fwd", "nerfb200_composite_bwd",
    "nerfb200_resample_alloc", "nerfb200_resample_fallback", "nerfb200_resample_icdf",
    "nerfb200_pose_fwd", "nerfb200_pose_bwd", "nerfb200_so3_to_SO3",
    "nerfb200_mlp_pack", "nerfb200_mlp_fwd", "nerfb200_mlp_fwd2", "nerfb200_mlp_bwd", "nerfb200_mlp_wgrad", "nerfb200_pe_fwd", "nerfb200_pe_bwd",
    "nerfb200_mlp_workspace_bytes", "nerfb200_garf_workspace_bytes", "nerfb200_garf_fwd", "nerfb200_garf_bwd",
    "nerfb200_render_rays_workspace_bytes", "nerfb200_render_rays", "nerfb200_lindisp_intervals", "nerfb200_trans_cdf_fwd", "nerfb200_trans_cdf_bwd", "nerfb200_prop_loss", "nerfb200_gauss_width_grad",
    "nerfb200_adam_step", "nerfb200_adam_step_dev", "nerfb200_act_fwd", "nerfb200_act_bwd", "nerfb200_ray_batch", "nerfb200_kabsch",
]


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().nerfb200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")


def launch_count() -> int:
    return int(lib().nerfb200_launch_count())
