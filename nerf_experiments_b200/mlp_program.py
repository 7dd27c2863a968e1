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
TMEM_EXTRA_COL = 256      # forward: density block; backward: position-encoding gradients
TMEM_DIR_COL = 320        # backward: direction-encoding gradients


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
    chunks: List[NbPackChunk] = []
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
                slab = SLAB_PE_POS if src.kind == "pos" else SLAB_PE_DIR
                inputs.append((slab, col, src.width, 0 if src.kind == "pos" else 1))
                col += src.width
        if col != lin.in_f:
            raise RuntimeError(f"layer {li}: sources cover {col} columns, weight has {lin.in_f}")
        if len(inputs) > NB_MAX_CHUNKS:
            raise RuntimeError(f"layer {li}: too many K chunks")
        op.n_chunks = len(inputs)
        rec = []
        for ci, (slab, c0, w, st) in enumerate(inputs):
            op.a_src[ci] = slab
            op.k16[ci] = _ceil(w, 16)
            op.w_rows[ci] = rows_img
            op.w_off[ci] = w_units
            # main rows
            chunks.append(NbPackChunk(base=lin.w_off + c0, row_stride=lin.in_f, col_stride=1,
                                      n_rows=min(L.out_main, lin.out_f) if L.act != "rgb" else lin.out_f,
                                      n_cols=w, rows_padded=n_main, dst_off=w_units))
            if extra:
                chunks.append(NbPackChunk(base=lin.w_off + L.out_main * lin.in_f + c0, row_stride=lin.in_f,
                                          col_stride=1, n_rows=1, n_cols=w, rows_padded=16,
                                          dst_off=w_units + n_main // 8))
            w_units += rows_img // 8
            rec.append((st, _ceil(w, 16)))
        op_inputs.append(rec)
        # blocks
        op.n_blocks = 2 if extra else 1
        op.blocks[0] = NbBlock(0, n_main, 0, 0)
        if extra:
            op.blocks[1] = NbBlock(TMEM_EXTRA_COL, 16, n_main, 0)
        # bias
        op.bias_off = bias_floats
        biases.append(NbPackBias(base=lin.b_off, n=(L.out_main if L.act != "rgb" else lin.out_f),
                                 n_padded=n_main, dst_off=bias_floats, reserved=0))
        bias_floats += n_main
        if extra:
            biases.append(NbPackBias(base=lin.b_off + L.out_main, n=1, n_padded=16, dst_off=bias_floats, reserved=0))
            bias_floats += 16
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
                       mask_words_per_tile=mask_words)


def to_device_array(items, ctype, device):
    """ctypes struct list -> uint8 CUDA tensor holding the array."""
    import torch as th
    n = len(items)
    arr = (ctype * max(n, 1))(*items)
    buf = bytes(arr)[: n * C.sizeof(ctype)] if n else b""
    t = th.frombuffer(bytearray(buf if buf else b"\0"), dtype=th.uint8).clone()
    return t.to(device)
