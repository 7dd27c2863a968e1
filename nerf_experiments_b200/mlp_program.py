"""Compiles a network description into the tile programs the fused MLP kernels execute
(csrc/mlp.h): forward ops, weight-image / bias packing descriptors, stash and mask layout.

The network-specific knowledge lives here, on the host: the kernels only interpret programs.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

from . import _lib
from ._lib import (NB_MAX_CHUNKS, NB_MAX_OPS, NbBlock, NbOp, NbPackBias, NbPackChunk, NbProgram)

SLAB_PE_POS = 4
SLAB_PE_DIR = 5
TMEM_EXTRA_COL = 256      # forward: density block = column 0 of the OTHER accumulator buffer


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b


@dataclass
class Linear:
    """One nn.Linear inside the flat fp32 parameter buffer."""
    w_off: int          # float offset of weight[0,0]; weight is (out_f, in_f) row-major
    b_off: int          # float offset of bias[0]
    out_f: int
    in_f: int


@dataclass
class KSource:
    """One run of input columns of a Linear and the shared-memory slabs that feed it."""
    kind: str           # "act" | "pos" | "dir"
    width: int          # real number of columns


@dataclass
class LayerSpec:
    lin: Linear
    sources: List[KSource]
    act: str                    # "relu" | "linear" | "rgb"
    sigma: Optional[str] = None  # None | "extra" (column out_main of this layer) | "col3"
    out_main: int = 0           # features that feed the next layer (excludes the density column)


@dataclass
class CompiledMlp:
    program: NbProgram
    pack_chunks: List[NbPackChunk]
    pack_biases: List[NbPackBias]
    wpack_bytes: int
    bias_floats: int
    layers: List[LayerSpec]
    # per fwd op: for every K chunk, (stash slab index or None, k16) ; output stash slab
    op_inputs: List[List[Tuple[int, int]]] = field(default_factory=list)
    stash_slabs_per_tile: int = 0
    mask_words_per_tile: int = 0
    dir_slab: int = SLAB_PE_DIR
    dir_encode_before_op: int = 0
    # two-tile forward kernel (csrc/mlp_fwd.cu, mlp_fwd2): the main weight chunks once more as
    # per-K-step images, the density row of the "extra" layer as fp32 weights behind the biases
    # (it is evaluated by the epilogue of the layer in front), and whether the program qualifies
    pack_chunks_k16: List[NbPackChunk] = field(default_factory=list)
    density_w_off: int = -1
    two_tile_ok: bool = False


def nerf_model_layers(lins: dict, n_hidden: int, hidden_dim: int, n_segments: int,
                      delayed_direction: bool, delayed_density: bool, pos_dim: int,
                      dir_dim: int) -> List[LayerSpec]:
    """Layer list of NerfModel (reference barf/model_interpolation_architecture.py:72-87,104-138).
    lins: {"model_segments.{i}.{2k}" | "model_segments.{i}" | "model_color.{0,2}": Linear}."""
    layers: List[LayerSpec] = []
    for i in range(n_segments):
        first_sources = []
        if i > 0:
            first_sources.append(KSource("act", hidden_dim))
        if not delayed_direction:
            first_sources.append(KSource("dir", dir_dim))
        first_sources.append(KSource("pos", pos_dim))
        n_lin = 1 if n_hidden == 0 else n_hidden + 1
        for k in range(n_lin):
            name = f"model_segments.{i}" if n_hidden == 0 else f"model_segments.{i}.{2 * k}"
            lin = lins[name]
            last_of_net = (i == n_segments - 1) and (k == n_lin - 1)
            sources = first_sources if k == 0 else [KSource("act", hidden_dim)]
            if last_of_net:
                if delayed_density:
                    layers.append(LayerSpec(lin, sources, "linear", None, hidden_dim))
                else:
                    layers.append(LayerSpec(lin, sources, "linear", "extra", hidden_dim))
            else:
                layers.append(LayerSpec(lin, sources, "relu", None, lin.out_f))
    color_sources = [KSource("act", hidden_dim)]
    if delayed_direction:
        color_sources.append(KSource("dir", dir_dim))
    layers.append(LayerSpec(lins["model_color.0"], color_sources, "relu", None, hidden_dim // 2))
    layers.append(LayerSpec(lins["model_color.2"], [KSource("act", hidden_dim // 2)], "rgb",
                            "col3" if delayed_density else None, 3))
    return layers


def compile_forward(layers: List[LayerSpec]) -> CompiledMlp:
    if len(layers) > NB_MAX_OPS:
        raise RuntimeError(f"network too deep for one tile program ({len(layers)} > {NB_MAX_OPS} layers)")
    prog = NbProgram()
    prog.n_ops = len(layers)
    # The direction encoding can take over slab 4 once the position encoding is dead (delayed
    # direction): 5 slabs leave room for a 4-stage weight ring. Otherwise 6 slabs + 3 stages.
    pos_uses = [i for i, L in enumerate(layers) if any(s.kind == "pos" for s in L.sources)]
    dir_uses = [i for i, L in enumerate(layers) if any(s.kind == "dir" for s in L.sources)]
    share = bool(dir_uses) and (not pos_uses or min(dir_uses) >= max(pos_uses) + 2)
    dir_slab = SLAB_PE_POS if (share or not dir_uses) else SLAB_PE_DIR
    dir_before = (max(pos_uses) + 1) if share and pos_uses else 0   # encoded while that op's MMAs run
    prog.n_slabs, prog.n_stages = (6, 3) if (dir_uses and not share) else (5, 4)
    chunks: List[NbPackChunk] = []
    chunks_k16: List[NbPackChunk] = []
    two_tile_ok = (prog.n_slabs == 5)
    density_w_off = -1
    biases: List[NbPackBias] = []
    w_units = 0          # 1024 B units used in the packed weight buffer
    bias_floats = 0
    stash = 2            # slab 0: pos encoding, slab 1: dir encoding
    mask_words = 0
    op_inputs = []
    prev_stash = None    # stash slab of the previous op's output
    for li, L in enumerate(layers):
        op = prog.ops[li]
        lin = L.lin
        if L.act == "rgb":
            n_main, out_chunks = 16, 0
            if lin.out_f > 16:
                raise RuntimeError("output layer wider than 16")
        else:
            if L.out_main > 256:
                raise RuntimeError(f"hidden width {L.out_main} > 256 is not supported by the fused kernel")
            out_chunks = _ceil(L.out_main, 64)
            n_main = out_chunks * 64
        extra = (L.sigma == "extra")
        rows_img = n_main + (16 if extra else 0)
        # K chunks
        col = 0
        ci = 0
        inputs = []
        for src in L.sources:
            if src.kind == "act":
                n_ch = _ceil(src.width, 64)
                for c in range(n_ch):
                    w = min(64, src.width - c * 64)
                    inputs.append((c, col + c * 64, w, (prev_stash + c) if prev_stash is not None else None))
                col += src.width
            else:
                if src.width > 64:
                    raise RuntimeError(f"{src.kind} encoding wider than 64 columns is not supported")
                slab = SLAB_PE_POS if src.kind == "pos" else dir_slab
                inputs.append((slab, col, src.width, 0 if src.kind == "pos" else 1))
                col += src.width
        if col != lin.in_f:
            raise RuntimeError(f"layer {li}: sources cover {col} columns, weight has {lin.in_f}")
        # the density row rides in the same image when it fits a ring stage, else in images of
        # its own (one 16-row image per K chunk, feeding only the extra block)
        split_extra = extra and rows_img * 128 > _lib.NB_RING_STAGE_BYTES
        n_entries = len(inputs) * (2 if split_extra else 1)
        for ci in range(NB_MAX_CHUNKS):
            op.n_sub[ci] = 1
        if n_entries > NB_MAX_CHUNKS:
            raise RuntimeError(f"layer {li}: too many K chunks")
        op.n_chunks = n_entries
        rec = []
        for ci, (slab, c0, w, st) in enumerate(inputs):
            op.a_src[ci] = slab
            op.k16[ci] = _ceil(w, 16)
            op.w_rows[ci] = n_main if split_extra else rows_img
            op.w_off[ci] = w_units
            op.blk_mask[ci] = 1 if (split_extra or not extra) else 3
            # main rows
            chunks.append(NbPackChunk(base=lin.w_off + c0, row_stride=lin.in_f, col_stride=1,
                                      n_rows=min(L.out_main, lin.out_f) if L.act != "rgb" else lin.out_f,
                                      n_cols=w, rows_padded=n_main, dst_off=w_units))
            chunks_k16.append(NbPackChunk(base=lin.w_off + c0, row_stride=lin.in_f, col_stride=1,
                                          n_rows=min(L.out_main, lin.out_f) if L.act != "rgb" else lin.out_f,
                                          n_cols=w, rows_padded=n_main, dst_off=w_units, img_rows=int(op.w_rows[ci])))
            if extra and not split_extra:
                two_tile_ok = False          # the density row shares the image: old kernel only
            if extra and not split_extra:
                chunks.append(NbPackChunk(base=lin.w_off + L.out_main * lin.in_f + c0, row_stride=lin.in_f,
                                          col_stride=1, n_rows=1, n_cols=w, rows_padded=16,
                                          dst_off=w_units + n_main // 8))
            w_units += op.w_rows[ci] // 8
            rec.append((st, _ceil(w, 16)))
        if split_extra:
            # runs of consecutive full-width act slabs share one image (one ring slot)
            ci = len(inputs)
            cj = 0
            while cj < len(inputs):
                slab, c0, w, st = inputs[cj]
                run = 1
                while (cj + run < len(inputs) and inputs[cj + run][0] == slab + run and w == 64
                       and inputs[cj + run][2] == 64):
                    run += 1
                op.a_src[ci] = slab
                op.k16[ci] = _ceil(w, 16)
                op.w_rows[ci] = 16
                op.w_off[ci] = w_units
                op.blk_mask[ci] = 2
                op.n_sub[ci] = run
                for r in range(run):
                    _, c0r, wr, _ = inputs[cj + r]
                    chunks.append(NbPackChunk(base=lin.w_off + L.out_main * lin.in_f + c0r, row_stride=lin.in_f,
                                              col_stride=1, n_rows=1, n_cols=wr, rows_padded=16, dst_off=w_units))
                    w_units += 2
                ci += 1
                cj += run
            op.n_chunks = ci
        op_inputs.append(rec)
        # blocks
        op.n_blocks = 2 if extra else 1
        op.blocks[0] = NbBlock(0, n_main, 0, 0)
        if extra:
            op.blocks[1] = NbBlock(TMEM_EXTRA_COL, 16, 0 if split_extra else n_main, 0)
        # bias
        op.bias_off = bias_floats
        biases.append(NbPackBias(base=lin.b_off, n=(L.out_main if L.act != "rgb" else lin.out_f),
                                 n_padded=n_main, dst_off=bias_floats, reserved=0))
        bias_floats += n_main
        if extra:
            biases.append(NbPackBias(base=lin.b_off + L.out_main, n=1, n_padded=16, dst_off=bias_floats, reserved=0))
            bias_floats += 16
            # two-tile kernel: density = <row out_main of W, input activations> in the epilogue of the
            # layer in front; needs that input to be exactly the 256 hidden columns of slabs 0..3
            plain = (li > 0 and len(L.sources) == 1 and L.sources[0].kind == "act" and L.sources[0].width == 256
                     and lin.in_f == 256 and layers[li - 1].act == "relu" and layers[li - 1].out_main == 256
                     and density_w_off < 0)
            if plain:
                density_w_off = bias_floats
                biases.append(NbPackBias(base=lin.w_off + L.out_main * lin.in_f, n=256, n_padded=256,
                                         dst_off=bias_floats, reserved=0))
                bias_floats += 256
            else:
                two_tile_ok = False
        # epilogue
        if L.act == "relu":
            op.epi = _lib.EPI_RELU_SIGMA if extra else _lib.EPI_RELU
        elif L.act == "linear":
            op.epi = _lib.EPI_LINEAR_SIGMA if extra else _lib.EPI_LINEAR
        else:
            op.epi = _lib.EPI_RGB_SIGMA if L.sigma == "col3" else _lib.EPI_RGB
        op.out_chunks = out_chunks
        op.out_width = L.out_main
        if out_chunks > 0:
            op.stash_slab = stash
            prev_stash = stash
            stash += out_chunks
            if L.act == "relu":
                op.mask_word = mask_words
                mask_words += out_chunks * 2
            else:
                op.mask_word = -1
        else:
            op.stash_slab = -1
            op.mask_word = -1
    prog.stash_slabs_per_tile = stash
    prog.mask_words_per_tile = mask_words
    return CompiledMlp(program=prog, pack_chunks=chunks, pack_biases=biases,
                       wpack_bytes=w_units * 1024, bias_floats=bias_floats, layers=layers,
                       op_inputs=op_inputs, stash_slabs_per_tile=stash,
                       mask_words_per_tile=mask_words, dir_slab=dir_slab, dir_encode_before_op=dir_before,
                       pack_chunks_k16=chunks_k16, density_w_off=density_w_off,
                       two_tile_ok=bool(two_tile_ok and bias_floats <= 3072
                                        and all(L.out_main in (0, 3, 64, 128, 192, 256) or L.act == "rgb" for L in layers)))


def to_device_array(items, ctype, device):
    """ctypes struct list -> uint8 CUDA tensor holding the array."""
    import torch as th
    n = len(items)
    arr = (ctype * max(n, 1))(*items)
    buf = bytes(arr)[: n * C.sizeof(ctype)] if n else b""
    t = th.frombuffer(bytearray(buf if buf else b"\0"), dtype=th.uint8).clone()
    return t.to(device)


# ---------------------------------------------------------------------------------------------
# backward (data gradients) and weight-gradient schedules
# ---------------------------------------------------------------------------------------------
@dataclass
class WgradUnit:
    dy_slab: int
    n_dy_slabs: int
    x_slab: int
    n_x_slabs: int
    m_real: int
    n_real: int
    dst: int
    ld: int
    bias_dst: int = -1     # float index of the bias gradient of the unit's first output feature
    mode: int = 0          # _lib.WGRAD_MMA | _lib.WGRAD_COLSUM (no weight block: n_x_slabs == 0, z duty only)
    coef_dst: int = -1     # z duty: float index of the Gaussian inverse-std parameter of the first feature
    z_slab: int = -1       # z duty (include/nerfb200_mlp.h): first z stash slab, count, first dY slab it belongs to,
    n_z_slabs: int = 0     # float index of the bias gradient of the first feature
    z_first: int = 0
    zbias_dst: int = -1
    x2_slab: int = -1      # >= 0: the LAST X slab comes from this stash slab (two input sources, one unit)


@dataclass
class CompiledBackward:
    program: NbProgram
    pack_chunks: List[NbPackChunk]
    wpack_units: int                  # 1024 B units appended to the packed weight buffer
    dy_slabs_per_tile: int
    head_bias_off: int
    head_sigma_col3: int
    pos_grad_cols: int
    dir_grad_cols: int
    bias_map: List[int]
    units: List[WgradUnit]


def _out_chunks(L: LayerSpec) -> int:
    return 0 if L.act == "rgb" else _ceil(L.out_main, 64)


def compile_backward(cm: CompiledMlp, want_input_grads: bool, encoders=None) -> CompiledBackward:
    """encoders: {"pos": (levels, has_identity_columns), "dir": ...} — needed for the canonical
    column order of the encoding-gradient accumulators (include/nerfb200_mlp.h)."""
    layers = cm.layers
    n_layers = len(layers)
    fprog = cm.program
    w_units = cm.wpack_bytes // 1024
    w_units0 = w_units
    chunks: List[NbPackChunk] = []

    # dY stash: slab 0 = head; then dY_{l-1} in the order the backward ops create them
    dy_slab = {n_layers - 1: 0}
    nxt = 1
    for l in range(n_layers - 1, 0, -1):
        prev = layers[l - 1]
        dy_slab[l - 1] = nxt
        nxt += _out_chunks(prev) + (1 if prev.sigma == "extra" else 0)
    dy_slabs_per_tile = nxt

    prog = NbProgram()
    ops = []
    grad_cols = {"pos": 0, "dir": 0}
    for l in range(n_layers - 1, -1, -1):
        L = layers[l]
        lin = L.lin
        # A operand: dY_l
        if L.act == "rgb":
            a_chunks = [(0, 1, 0, lin.out_f)]            # (slab, k16, first out row, n out rows)
        else:
            a_chunks = []
            for c in range(_out_chunks(L)):
                w = min(64, L.out_main - 64 * c)
                a_chunks.append((c, _ceil(w, 16), 64 * c, w))
            if L.sigma == "extra":
                a_chunks.append((4, 1, L.out_main, 1))
        col0 = 0
        main_src = None
        aux_srcs = []
        for src in L.sources:
            if src.kind == "act":
                main_src = (col0, src.width)
            else:
                aux_srcs.append((src.kind, col0, src.width))
            col0 += src.width

        def emit(op: NbOp, col_first: int, width: int, n_block: int, tmem_col: int, accum: int):
            nonlocal w_units
            op.n_chunks = len(a_chunks)
            for ci, (slab, k16, krow0, kcols) in enumerate(a_chunks):
                op.a_src[ci] = slab
                op.k16[ci] = k16
                op.w_rows[ci] = n_block
                op.w_off[ci] = w_units
                op.blk_mask[ci] = 1
                op.n_sub[ci] = 1
                chunks.append(NbPackChunk(base=lin.w_off + krow0 * lin.in_f + col_first, row_stride=1,
                                          col_stride=lin.in_f, n_rows=width, n_cols=kcols,
                                          rows_padded=n_block, dst_off=w_units))
                w_units += n_block // 8
            op.n_blocks = 1
            op.blocks[0] = NbBlock(tmem_col, n_block, 0, accum)

        if want_input_grads:
            for kind, c0, width in aux_srcs:
                levels, has_id = encoders[kind]
                if levels > _lib.PE_CANON_LEVELS or width != 3 * has_id + 6 * levels:
                    raise RuntimeError(f"{kind} encoding ({levels} levels, width {width}) does not fit the "
                                       "canonical gradient layout")
                op = NbOp()
                n_block = _lib.PE_CANON_COLS
                op.n_chunks = len(a_chunks)
                for ci, (slab, k16, krow0, kcols) in enumerate(a_chunks):
                    op.a_src[ci] = slab
                    op.k16[ci] = k16
                    op.w_rows[ci] = n_block
                    op.w_off[ci] = w_units
                    op.blk_mask[ci] = 1
                    op.n_sub[ci] = 1
                    base = lin.w_off + krow0 * lin.in_f + c0
                    # image row = canonical column; rows no descriptor writes stay zero
                    groups = []
                    if has_id:
                        groups.append((0, 3, _lib.PE_CANON_IDENTITY, 1))
                    off = 3 * has_id
                    for cc in range(3):
                        groups.append((off + cc * levels, levels, 2 * _lib.PE_CANON_LEVELS * cc, 2))
                        groups.append((off + 3 * levels + cc * levels, levels, 2 * _lib.PE_CANON_LEVELS * cc + 1, 2))
                    for e0, n_e, row0, step in groups:
                        if n_e > 0:
                            chunks.append(NbPackChunk(base=base + e0, row_stride=1, col_stride=lin.in_f,
                                                      n_rows=n_e, n_cols=kcols, rows_padded=n_e,
                                                      dst_off=w_units, dst_row0=row0, dst_row_step=step))
                    w_units += n_block // 8
                op.n_blocks = 1
                op.blocks[0] = NbBlock(0, n_block, 0, 0)
                grad_cols[kind] = n_block
                op.epi = _lib.BEPI_PEGRAD_POS if kind == "pos" else _lib.BEPI_PEGRAD_DIR
                op.out_chunks = 0
                op.bias_off = -1
                op.stash_slab = -1
                op.mask_word = -1
                ops.append(op)
        if main_src is not None:
            if l == 0:
                raise RuntimeError("first layer cannot take activations")
            prev = layers[l - 1]
            c0, width = main_src
            op = NbOp()
            emit(op, c0, width, _ceil(width, 64) * 64, 0, 0)
            relu = (prev.act == "relu")
            sig = (prev.sigma == "extra")
            op.epi = {(True, False): _lib.BEPI_MASK, (False, False): _lib.BEPI_PLAIN,
                      (False, True): _lib.BEPI_PLAIN_SIGMA, (True, True): _lib.BEPI_MASK_SIGMA}[(relu, sig)]
            op.out_chunks = _ceil(width, 64)
            op.bias_off = fprog.ops[l - 1].bias_off
            op.stash_slab = dy_slab[l - 1]
            op.mask_word = fprog.ops[l - 1].mask_word if relu else -1
            op.out_width = width
            ops.append(op)
    if len(ops) > NB_MAX_OPS:
        raise RuntimeError("backward program too long")
    prog.n_ops = len(ops)
    for i, op in enumerate(ops):
        prog.ops[i] = op
    prog.stash_slabs_per_tile = dy_slabs_per_tile
    prog.mask_words_per_tile = cm.mask_words_per_tile
    prog.n_slabs, prog.n_stages = 5, 4

    bias_map = [-1] * cm.bias_floats
    for pb in cm.pack_biases:
        for i in range(pb.n):
            bias_map[pb.dst_off + i] = pb.base + i

    # weight-gradient units
    units: List[WgradUnit] = []
    for l, L in enumerate(layers):
        lin = L.lin
        n_main_slabs = 1 if L.act == "rgb" else _out_chunks(L)
        m_main = lin.out_f if L.act == "rgb" else L.out_main
        col0 = 0
        for src in L.sources:
            if src.kind == "act":
                x_slab, n_x = fprog.ops[l - 1].stash_slab, _ceil(src.width, 64)
            else:
                x_slab, n_x = (0 if src.kind == "pos" else 1), 1
            # the unit of a layer's first input source also carries the bias gradient
            first = (col0 == 0)
            units.append(WgradUnit(dy_slab[l], n_main_slabs, x_slab, n_x, m_main, src.width,
                                   lin.w_off + col0, lin.in_f, lin.b_off if first else -1))
            if L.sigma == "extra":
                units.append(WgradUnit(dy_slab[l] + n_main_slabs, 1, x_slab, n_x, 1, src.width,
                                       lin.w_off + L.out_main * lin.in_f + col0, lin.in_f,
                                       lin.b_off + L.out_main if first else -1))
            col0 += src.width

    head = layers[-1]
    return CompiledBackward(program=prog, pack_chunks=chunks, wpack_units=w_units - w_units0,
                            dy_slabs_per_tile=dy_slabs_per_tile, head_bias_off=fprog.ops[n_layers - 1].bias_off,
                            head_sigma_col3=1 if head.sigma == "col3" else 0,
                            pos_grad_cols=grad_cols["pos"], dir_grad_cols=grad_cols["dir"],
                            bias_map=bias_map, units=units)


def schedule_wgrad(units: List[WgradUnit], n_tiles: int, n_workers: int, per_worker: Optional[float] = None):
    """Splits every unit over tile ranges so that ~3 items per worker of similar cost result;
    returns NbWgradItem structs sorted by decreasing cost (static round-robin in the kernel).
    The kernel is HBM-bound, so an SM idling at the end of the launch is lost bandwidth, while
    every item costs a pipeline fill and a TMEM flush: measured on the bench workload (4096
    tiles, 148 SMs) 1 / 2 / 3 / 4 / 6 / 8 items per worker take 1.55 / 1.29 / 1.13 / 1.19 / 1.19 /
    1.37 ms (a longest-first assignment of the same items is no better). Round 2, GARF radiance network
    (6144 tiles, 22 units of cost 5 .. 8 slabs per tile): 2 / 3 / 4 / 6 / 8 / 10 / 16 items per worker take
    2.50 / 2.60 / 2.18 / 1.98 / 2.03 / 2.12 / 2.21 ms, and an equal-bytes partition with one or two LARGE items
    per worker is the slowest of all (2.77 ms; 1.49 ms on the ReLU network): what matters is that the
    accumulator flushes of the workers are spread over the launch instead of meeting at its end. The jagged
    curve came from the static round-robin handing worker 0 the largest item of every round; with every other
    round reversed (below) it is smooth — 2 / 3 / 4 / 5 / 6 / 8 items: 2.90 / 2.02 / 1.85 / 1.87 / 1.89 / 1.93 ms
    (after the concatenating layers' units were merged) — so the GARF fields pass `per_worker` ~ streamed bytes /
    (workers x 20 MB), at least 3; the ReLU network stays at 3 (1.08 ms either way)."""
    import os
    from ._lib import NbWgradItem
    if per_worker is None:
        per_worker = 3.0
    per_worker = float(os.environ.get("NB_WGRAD_ITEMS_PER_WORKER", per_worker))   # the env knob is for experiments
    cost = [(u.n_dy_slabs + u.n_x_slabs + u.n_z_slabs) for u in units]
    total = sum(cost) * n_tiles
    target = max(total / max(per_worker * n_workers, 1), 1.0)
    items = []
    for u in units:     # one pipeline stage of csrc/mlp_wgrad.cu holds nine half slabs: dY (pairs) | X | Z
        if 2 * ((u.n_dy_slabs + 1) // 2) + u.n_x_slabs + u.n_z_slabs > 9 or not (1 <= u.n_dy_slabs <= 4) or u.n_x_slabs > 4:
            raise RuntimeError(f"weight-gradient unit does not fit a pipeline stage: {u}")
    for u, c in zip(units, cost):
        splits = int(min(n_tiles, max(1, round(c * n_tiles / target))))
        for s in range(splits):
            t0 = (n_tiles * s) // splits
            t1 = (n_tiles * (s + 1)) // splits
            if t1 > t0:
                items.append((c * (t1 - t0), NbWgradItem(tile_begin=t0, tile_end=t1, n_dy_slabs=u.n_dy_slabs,
                                                         n_x_slabs=u.n_x_slabs, dy_slab=u.dy_slab, x_slab=u.x_slab,
                                                         m_real=u.m_real, n_real=u.n_real, dst=u.dst, ld=u.ld,
                                                         bias_dst=u.bias_dst, mode=u.mode, coef_dst=u.coef_dst,
                                                         z_slab=u.z_slab, n_z_slabs=u.n_z_slabs, z_first=u.z_first,
                                                         zbias_dst=u.zbias_dst, x2_slab=u.x2_slab)))
    items.sort(key=lambda t: -t[0])
    if os.environ.get("NB_WGRAD_SNAKE", "1") == "1":
        # The kernel hands item i to worker i mod n_workers: with the items sorted by decreasing cost, worker 0
        # would get the largest item of EVERY round. Reversing every other round (boustrophedon) pairs a
        # worker's large item of one round with a small one of the next.
        for r in range(1, (len(items) + n_workers - 1) // n_workers, 2):
            items[r * n_workers: (r + 1) * n_workers] = items[r * n_workers: (r + 1) * n_workers][::-1]
    return [it for _, it in items]
