"""Compiles the GARF radiance / proposal networks (reference garf/model_radiance.py:23-96,
garf/model_proposal.py:22-56) into the tile programs of the fused GARF kernels
(include/nerfb200_garf.h): forward and backward op / step lists, weight-image and fp32 packing
descriptors, stash layouts and the work units of the weight-gradient kernel.

Shared-memory slabs: 0..3 "work" (the 256 columns an op reads / an epilogue writes), 4..5 "hold"
(the 128-column output z1 of the first sub-network, which feeds two column blocks of the next layer
and the residual; in the backward pass the 128-column gradient blocks of a 512-wide layer).
Accumulators: R0 = TMEM columns 0..255, R1 = 256..511; a column-blocked layer pair alternates
them (the block goes to R0, the layer behind it accumulates in R1).
"""
from dataclasses import dataclass, field
from typing import List, Optional

from . import _lib
from ._lib import (NG_BSTEP_ACT, NG_BSTEP_HEAD, NG_BSTEP_PLAIN, NG_F_DIRECT, NG_F_FIRST_LAYER, NG_F_HOLD_ADD,
                   NG_F_HOLD_SAVE, NG_F_SIGMA, NG_GEN_COLS, NG_MAX_CHUNKS, NG_MAX_FLOATS, NG_MAX_OPS, NG_STEP_ACT,
                   NG_STEP_GEN, NG_STEP_LINEAR, NG_STEP_NONE, NG_STEP_RGB, NG_STEP_SIGMA, NbPackBias, NbPackChunk,
                   NgBlock, NgOp, NgProgram, NgStep)
from .mlp_program import Linear, WgradUnit

R0, R1 = 0, 256
HOLD = 4


@dataclass
class GaussLinear:
    """nn.Linear (+ GaussAct) inside the flat fp32 parameter buffer."""
    lin: Linear
    g_off: int = -1          # float offset of GaussAct.inv_standard_deviation (out_f values), -1: no activation


@dataclass
class CompiledGarf:
    fwd: NgProgram
    bwd: NgProgram
    pack_chunks: List[NbPackChunk]
    fwd_floats: List[NbPackBias]
    bwd_floats: List[NbPackBias]
    wpack_units: int
    units: List[WgradUnit]
    has_rgb: bool
    macs_per_sample: int
    gauss_per_sample: int
    gauss_layers: List = field(default_factory=list)   # GaussLinear layers whose width gradient follows from dW and db


class _Builder:
    def __init__(self):
        self.ops: List[NgOp] = []
        self.steps: List[NgStep] = []
        self.floats: List[NbPackBias] = []
        self.n_floats = 0
        self.chunks: List[NbPackChunk] = []
        self.w_units = 0

    # ---- packed fp32 values ----
    def pack(self, base: int, n: int, n_padded: Optional[int] = None, kind: int = _lib.PACK_COPY, stride: int = 1) -> int:
        n_padded = n if n_padded is None else n_padded
        off = self.n_floats
        self.floats.append(NbPackBias(base=base, n=n, n_padded=n_padded, dst_off=off, kind=kind, stride=stride))
        self.n_floats += n_padded
        return off

    def pack_skip(self, lin: Linear, row0: int, n_rows: int, col0: int) -> int:
        """[3][n_rows] block of lin.weight[row0:row0+n_rows, col0:col0+3] (one segment per input column)."""
        off = self.n_floats
        for c in range(3):
            self.pack(lin.w_off + row0 * lin.in_f + col0 + c, n_rows, stride=lin.in_f)
        return off

    # ---- weight images ----
    def image(self, lin: Linear, out0: int, n_out: int, in0: int, n_in: int, rows_padded: int, transposed: bool,
              dst_row0: int = 0) -> int:
        """One [rows][64] bf16 image. Forward: rows = output features out0.., columns = inputs in0..
        Transposed (data gradients): rows = input features in0.., columns = output features out0.."""
        off = self.w_units
        if not transposed:
            self.chunks.append(NbPackChunk(base=lin.w_off + out0 * lin.in_f + in0, row_stride=lin.in_f, col_stride=1,
                                           n_rows=n_out, n_cols=n_in, rows_padded=rows_padded, dst_off=off,
                                           dst_row0=dst_row0))
        else:
            self.chunks.append(NbPackChunk(base=lin.w_off + out0 * lin.in_f + in0, row_stride=1, col_stride=lin.in_f,
                                           n_rows=n_in, n_cols=n_out, rows_padded=rows_padded, dst_off=off,
                                           dst_row0=dst_row0))
        return off

    def alloc_image(self, rows: int) -> None:
        self.w_units += rows // 8

    # ---- program entries ----
    def step(self, kind=NG_STEP_NONE, lag=0, n_slabs=0, out_slab=-1, res_slab=-1, skip_src=0, flags=0, src_col=0,
             sigma_col=0, bias_off=-1, coef_off=-1, skip_off=-1, y_stash=-1, z_stash=-1, gen_col0=0):
        self.steps.append(NgStep(kind=kind, wait_lag=lag, n_slabs=n_slabs, out_slab=out_slab, res_slab=res_slab,
                                 skip_src=skip_src, flags=flags, reserved=0, src_col=src_col, sigma_col=sigma_col,
                                 bias_off=bias_off, coef_off=coef_off, skip_off=skip_off, y_stash=y_stash,
                                 z_stash=z_stash, gen_col0=gen_col0))

    def op(self, a_slabs, k16s, w_offs, w_rows, blocks, accumulate=False):
        if len(a_slabs) > NG_MAX_CHUNKS:
            raise RuntimeError("GARF program: too many K chunks in one op")
        o = NgOp()
        o.n_chunks = len(a_slabs)
        o.n_blocks = max(len(blocks), 1)
        o.accumulate = int(accumulate)
        o.w_rows = w_rows
        for c, (a, k, w) in enumerate(zip(a_slabs, k16s, w_offs)):
            o.a_slab[c], o.k16[c], o.w_off[c] = a, k, w
        for b, (col, n, row0) in enumerate(blocks):
            o.blocks[b] = NgBlock(tmem_col=col, n=n, row0=row0, reserved=0)
        self.ops.append(o)

    @staticmethod
    def _early_ok(st, o) -> bool:
        """May the MMAs of op k start slab by slab while step k is still running (include/nerfb200_garf.h,
        NgOp.early)? Only behind a Gaussian step that publishes slabs, and only if the op writes no accumulator
        column that step still has to read."""
        if o.n_chunks == 0 or st.out_slab < 0 or st.n_slabs <= 0 or (st.flags & NG_F_DIRECT):
            return False
        if st.kind not in (NG_STEP_ACT, NG_BSTEP_ACT):
            return False
        lo, hi = st.src_col, st.src_col + 64 * st.n_slabs
        for b in range(o.n_blocks):
            c0, c1 = o.blocks[b].tmem_col, o.blocks[b].tmem_col + o.blocks[b].n
            if c0 < hi and lo < c1:
                return False
        return True

    def nop(self):
        self.ops.append(NgOp())

    def finish(self, prog: NgProgram):
        if len(self.ops) > NG_MAX_OPS or len(self.steps) != len(self.ops) + 1:
            raise RuntimeError(f"GARF program: {len(self.ops)} ops / {len(self.steps)} steps")
        if self.n_floats > _lib.NG_MAX_PROGRAM_FLOATS:
            raise RuntimeError(f"GARF program: {self.n_floats} packed floats exceed {_lib.NG_MAX_PROGRAM_FLOATS}")
        if self.steps[-1].wait_lag != 0:
            raise RuntimeError("GARF program: the last step must wait for the last op")
        # the residual-path gradient parked in TMEM (backward: NG_F_HOLD_SAVE .. NG_F_HOLD_ADD, 64 columns at the steps'
        # sigma_col): no op issued in between may write those columns
        save = [k for k, st in enumerate(self.steps) if st.kind == NG_BSTEP_PLAIN and (st.flags & NG_F_HOLD_SAVE)]
        add = [k for k, st in enumerate(self.steps) if st.kind == NG_BSTEP_ACT and (st.flags & NG_F_HOLD_ADD)]
        if save or add:
            if len(save) != 1 or len(add) != 1 or self.steps[save[0]].sigma_col != self.steps[add[0]].sigma_col:
                raise RuntimeError("GARF program: the hold save / add steps do not pair up")
            c0 = self.steps[save[0]].sigma_col
            for k in range(save[0], add[0]):          # op k is issued after step k; the add step reads before op add[0]
                for b in range(self.ops[k].n_blocks if self.ops[k].n_chunks else 0):
                    blk = self.ops[k].blocks[b]
                    if blk.tmem_col < c0 + 64 and c0 < blk.tmem_col + blk.n:
                        raise RuntimeError(f"GARF program: op {k} overwrites the TMEM columns that hold the residual gradient")
        prog.n_ops = len(self.ops)
        prog.n_floats = self.n_floats
        for i, o in enumerate(self.ops):
            o.early = int(self._early_ok(self.steps[i], o))
            prog.ops[i] = o
        for i, s in enumerate(self.steps):
            prog.steps[i] = s


def _ceil(a, b):
    return (a + b - 1) // b


class _Stash:
    """Sequential allocation of per-tile slab indices."""

    def __init__(self):
        self.n = 0

    def take(self, n_slabs: int) -> int:
        first = self.n
        self.n += n_slabs
        return first


def _act_layer_fwd(b: _Builder, layer: GaussLinear, src_col, out_slab, y_stash, z_stash, lag=0, out0=0, n_out=None,
                   skip=None):
    """Epilogue step of a Gaussian layer (or of a column block of it): bias, coefficient, optional skip."""
    lin = layer.lin
    n_out = lin.out_f if n_out is None else n_out
    nsl = _ceil(n_out, 64)
    bias = b.pack(lin.b_off + out0, n_out, 64 * nsl)
    coef = b.pack(layer.g_off + out0, n_out, 64 * nsl, kind=_lib.PACK_GAUSS)
    skip_off, skip_src = -1, 0
    if skip is not None:
        skip_src, col0 = skip
        if n_out != 64 * nsl:
            raise RuntimeError("skip blocks must be multiples of 64 columns")
        skip_off = b.pack_skip(lin, out0, n_out, col0)
    b.step(NG_STEP_ACT, lag=lag, n_slabs=nsl, out_slab=out_slab, src_col=src_col, bias_off=bias, coef_off=coef,
           skip_off=skip_off, skip_src=skip_src, y_stash=y_stash, z_stash=z_stash)


def _mma_fwd(b: _Builder, lin: Linear, a_slabs, in0, out0, n_out, tmem_col, accumulate=False, extra_row=None):
    """Forward op: D[:, out0:out0+n_out] (+)= A[slabs] W[out0:.., in0: in0 + 64 * len(a_slabs)]^T.
    extra_row: (row index of W, tmem column) of a single extra output (the density column)."""
    n_pad = _ceil(n_out, 16) * 16
    rows = n_pad + (16 if extra_row is not None else 0)
    offs = []
    for c, _ in enumerate(a_slabs):
        n_in = min(64, lin.in_f - (in0 + 64 * c))
        offs.append(b.image(lin, out0, n_out, in0 + 64 * c, n_in, n_pad, False))
        if extra_row is not None:
            b.image(lin, extra_row[0], 1, in0 + 64 * c, n_in, 16, False, dst_row0=n_pad)
        b.alloc_image(rows)
    blocks = [(tmem_col, n_pad, 0)]
    if extra_row is not None:
        blocks.append((extra_row[1], 16, n_pad))
    b.op(a_slabs, [4] * len(a_slabs), offs, rows, blocks, accumulate)


def _mma_bwd(b: _Builder, lin: Linear, a_slabs, k_outs, in0, n_in, tmem_col, accumulate=False):
    """Data-gradient op: D[:, in0:in0+n_in] (+)= dZ[slabs] W[outs, in0:in0+n_in]; k_outs = per A slab the
    (first output feature, number of output features) it holds."""
    n_pad = _ceil(n_in, 16) * 16
    offs, k16s = [], []
    for (o0, n_o) in k_outs:
        offs.append(b.image(lin, o0, n_o, in0, n_in, n_pad, True))
        b.alloc_image(n_pad)
        k16s.append(_ceil(n_o, 16))
    b.op(a_slabs, k16s, offs, n_pad, [(tmem_col, n_pad, 0)], accumulate)


def _colsum_units(units, dy0, z0, layer: GaussLinear):
    """Bias gradient of a Gaussian layer (column sums of its dz slabs): rides on ONE weight unit per group of dz
    slabs, as for the ReLU network. The width gradient needs no pass over the samples at all: sum z dz =
    W . dW + b db per output feature (include/nerfb200.h, nerfb200_gauss_width_grad), so the weight-gradient
    kernel never reads the z stash. Must be called after the layer's weight units have been appended."""
    n_slabs = _ceil(layer.lin.out_f, 64)
    todo = [True] * n_slabs
    for u in units:
        if u.mode != _lib.WGRAD_MMA or u.bias_dst >= 0:
            continue
        lo, hi = u.dy_slab - dy0, u.dy_slab - dy0 + u.n_dy_slabs
        if lo < 0 or hi > n_slabs or not all(todo[lo:hi]):
            continue
        u.bias_dst = layer.lin.b_off + 64 * lo
        for s in range(lo, hi):
            todo[s] = False
    if any(todo):
        raise RuntimeError("GARF program: a dz slab without a weight unit to carry its bias gradient")


def _weight_units(units, lin: Linear, dy0, n_out, x0, in0, n_in, bias=False, out0=0):
    """dW[out0:out0+n_out, in0:in0+n_in] from dY slabs dy0.. (n_out features) and X slabs x0.. (n_in columns)."""
    for u in range(_ceil(n_out, 256)):
        m = min(256, n_out - 256 * u)
        for v in range(_ceil(n_in, 256)):
            n = min(256, n_in - 256 * v)
            first = bias and v == 0
            units.append(WgradUnit(dy0 + 4 * u, _ceil(m, 64), x0 + 4 * v, _ceil(n, 64), m, n,
                                   lin.w_off + (out0 + 256 * u) * lin.in_f + in0 + 256 * v, lin.in_f,
                                   bias_dst=(lin.b_off + out0 + 256 * u) if first else -1))


def _weight_units_concat(units, lin: Linear, dy0, n_out, x0, n_main, x_extra, n_extra):
    """A concatenating layer [main (n_main columns, X slabs x0..) | extra (n_extra <= 64 columns, slab x_extra)] as ONE
    unit per 256 output features: the extra slab rides as the unit's last X slab (NbWgradItem.x2_slab), so the dY
    slabs are streamed once instead of once per source."""
    if n_main % 64 or n_main + 64 > 256 or n_extra > 64 or lin.in_f != n_main + n_extra:
        raise RuntimeError("concatenating layer does not fit one weight-gradient unit")
    for u in range(_ceil(n_out, 256)):
        m = min(256, n_out - 256 * u)
        units.append(WgradUnit(dy0 + 4 * u, _ceil(m, 64), x0, n_main // 64 + 1, m, n_main + n_extra,
                               lin.w_off + 256 * u * lin.in_f, lin.in_f, x2_slab=x_extra))


def compile_radiance(L: List[GaussLinear], Lc: List[GaussLinear]) -> CompiledGarf:
    """L = [L1 (3->1024 G), L2 (1024->256 G), L3 (256->128 G), L4 (128->128 G), L5 (131->512 G),
    L6 (512->256 G), L7 (256->128 G), L8 (128->129)], Lc = [C1 (131->256 G), C2 (256->3)]
    (garf/model_radiance.py:23-60); forward :84-96."""
    shapes = [(l.lin.in_f, l.lin.out_f) for l in L + Lc]
    if shapes != [(3, 1024), (1024, 256), (256, 128), (128, 128), (131, 512), (512, 256), (256, 128), (128, 129),
                  (131, 256), (256, 3)]:
        raise RuntimeError(f"the fused GARF radiance program is laid out for the reference's layer sizes, got {shapes}")
    L1, L2, L3, L4, L5, L6, L7, L8 = L
    C1, C2 = Lc
    ys, zs = _Stash(), _Stash()
    aux_pos, aux_dir = ys.take(1), ys.take(1)
    y1, z1 = ys.take(16), zs.take(16)
    y2, z2 = ys.take(4), zs.take(4)
    y3, z3 = ys.take(2), zs.take(2)
    y4, z4 = ys.take(2), zs.take(2)
    y5, z5 = ys.take(8), zs.take(8)
    y6, z6 = ys.take(4), zs.take(4)
    y7, z7 = ys.take(2), zs.take(2)
    ysum = ys.take(2)
    yc1, zc1 = ys.take(4), zs.take(4)

    # ------------------------------------------------------------------ forward
    f = _Builder()
    n_gen = L1.lin.out_f // NG_GEN_COLS
    for blk in range(n_gen):                 # first layer in registers, second layer accumulating in R1
        pair = 2 * (blk & 1)
        f.step(NG_STEP_GEN, lag=1, n_slabs=2, out_slab=pair, gen_col0=blk * NG_GEN_COLS, y_stash=y1 + 2 * blk,
               z_stash=z1 + 2 * blk)
        _mma_fwd(f, L2.lin, [pair, pair + 1], blk * NG_GEN_COLS, 0, 256, R1, accumulate=blk > 0)
    _act_layer_fwd(f, L2, R1, 0, y2, z2)
    _mma_fwd(f, L3.lin, [0, 1, 2, 3], 0, 0, 128, R0)
    _act_layer_fwd(f, L3, R0, 0, y3, z3)
    _mma_fwd(f, L4.lin, [0, 1], 0, 0, 128, R1)      # regions alternate: an op never overwrites the columns the step in front of it reads,
    _act_layer_fwd(f, L4, R1, HOLD, y4, z4)         # so its MMAs may start slab by slab (NgOp.early); z1 stays in the hold slabs
    for blk in range(2):                     # 131 -> 512 in two column blocks, 512 -> 256 accumulating in R1
        if blk > 0:
            f.step(NG_STEP_NONE, lag=1)
        _mma_fwd(f, L5.lin, [HOLD, HOLD + 1], 0, 256 * blk, 256, R0)
        _act_layer_fwd(f, L5, R0, 0, y5 + 4 * blk, z5 + 4 * blk, out0=256 * blk, n_out=256, skip=(1, 128))
        _mma_fwd(f, L6.lin, [0, 1, 2, 3], 256 * blk, 0, 256, R1, accumulate=blk > 0)
    _act_layer_fwd(f, L6, R1, 0, y6, z6)
    _mma_fwd(f, L7.lin, [0, 1, 2, 3], 0, 0, 128, R0)
    _act_layer_fwd(f, L7, R0, 0, y7, z7)
    _mma_fwd(f, L8.lin, [0, 1], 0, 0, 128, R1, extra_row=(128, R1 + 128))  # density = column 128 (garf/model_radiance.py:91)
    bias8 = f.pack(L8.lin.b_off, 129, 144)
    f.step(NG_STEP_LINEAR, n_slabs=2, out_slab=0, src_col=R1, bias_off=bias8, res_slab=HOLD, flags=NG_F_SIGMA,
           sigma_col=R1 + 128, y_stash=ysum)                              # z1 + z2[:, :128]  (:93)
    _mma_fwd(f, C1.lin, [0, 1], 0, 0, 256, R0)
    _act_layer_fwd(f, C1, R0, 0, yc1, zc1, skip=(2, 128))
    _mma_fwd(f, C2.lin, [0, 1, 2, 3], 0, 0, 3, R1)
    f.step(NG_STEP_RGB, src_col=R1, bias_off=f.pack(C2.lin.b_off, 3, 16))
    fwd = NgProgram()
    f.finish(fwd)
    fwd.y_slabs_per_tile, fwd.z_slabs_per_tile = ys.n, zs.n
    fwd.w1_off, fwd.b1_off, fwd.g1_off, fwd.n1 = L1.lin.w_off, L1.lin.b_off, L1.g_off, L1.lin.out_f
    fwd.aux_pos_stash, fwd.aux_dir_stash = aux_pos, aux_dir
    fwd.sigma_bias = -1.0                                                 # softplus(z2[:, 128] - 1)  (:91)

    # ------------------------------------------------------------------ backward
    ds = _Stash()
    d_head = ds.take(1)
    d_c1 = ds.take(4)
    d_8 = ds.take(3)            # 128 columns + the density column in a slab of its own
    d_7, d_6, d_5, d_4, d_3, d_2, d_1 = ds.take(2), ds.take(4), ds.take(8), ds.take(2), ds.take(2), ds.take(4), ds.take(16)
    g = _Builder()
    g.w_units = f.w_units
    coef = lambda layer, o0, n: g.pack(layer.g_off + o0, n, _ceil(n, 64) * 64, kind=_lib.PACK_GAUSS)

    g.step(NG_BSTEP_HEAD, lag=0, n_slabs=1, out_slab=0, y_stash=d_head)
    _mma_bwd(g, C2.lin, [0], [(0, 3)], 0, 256, R0)
    g.step(NG_BSTEP_ACT, n_slabs=4, out_slab=0, src_col=R0, coef_off=coef(C1, 0, 256), z_stash=zc1, y_stash=d_c1,
           skip_src=2, skip_off=g.pack_skip(C1.lin, 0, 256, 128))
    _mma_bwd(g, C1.lin, [0, 1, 2, 3], [(64 * c, 64) for c in range(4)], 0, 128, R1)
    HOLD_COL = R0 + 128        # 64 TMEM columns no op touches between the save and the add (checked in finish())
    g.step(NG_BSTEP_PLAIN, n_slabs=2, out_slab=0, src_col=R1, flags=NG_F_HOLD_SAVE | NG_F_SIGMA, y_stash=d_8,
           sigma_col=HOLD_COL)
    _mma_bwd(g, L8.lin, [0, 1, 2], [(0, 64), (64, 64), (128, 1)], 0, 128, R0)
    g.step(NG_BSTEP_ACT, n_slabs=2, out_slab=0, src_col=R0, coef_off=coef(L7, 0, 128), z_stash=z7, y_stash=d_7)
    _mma_bwd(g, L7.lin, [0, 1], [(0, 64), (64, 64)], 0, 256, R1)
    g.step(NG_BSTEP_ACT, n_slabs=4, out_slab=0, src_col=R1, coef_off=coef(L6, 0, 256), z_stash=z6, y_stash=d_6)
    for blk in range(4):        # d(y5) in four 128-column blocks (R0), d(z1) accumulating in R1
        if blk > 0:
            g.step(NG_STEP_NONE, lag=1)
        _mma_bwd(g, L6.lin, [0, 1, 2, 3], [(64 * c, 64) for c in range(4)], 128 * blk, 128, R0)
        g.step(NG_BSTEP_ACT, n_slabs=2, out_slab=HOLD, src_col=R0, coef_off=coef(L5, 128 * blk, 128),
               z_stash=z5 + 2 * blk, y_stash=d_5 + 2 * blk, skip_src=1, skip_off=g.pack_skip(L5.lin, 128 * blk, 128, 128))
        _mma_bwd(g, L5.lin, [HOLD, HOLD + 1], [(128 * blk, 64), (128 * blk + 64, 64)], 0, 128, R1, accumulate=blk > 0)
    g.step(NG_BSTEP_ACT, n_slabs=2, out_slab=0, src_col=R1, coef_off=coef(L4, 0, 128), z_stash=z4, y_stash=d_4,
           flags=NG_F_HOLD_ADD, sigma_col=HOLD_COL)                       # + the residual path (z1 + z2)
    _mma_bwd(g, L4.lin, [0, 1], [(0, 64), (64, 64)], 0, 128, R0)
    g.step(NG_BSTEP_ACT, n_slabs=2, out_slab=0, src_col=R0, coef_off=coef(L3, 0, 128), z_stash=z3, y_stash=d_3)
    _mma_bwd(g, L3.lin, [0, 1], [(0, 64), (64, 64)], 0, 256, R1)
    g.step(NG_BSTEP_ACT, n_slabs=4, out_slab=0, src_col=R1, coef_off=coef(L2, 0, 256), z_stash=z2, y_stash=d_2)
    _first_layer_bwd(g, L1, L2, z1, d_1)
    bwd = NgProgram()
    g.finish(bwd)
    bwd.y_slabs_per_tile, bwd.z_slabs_per_tile = ds.n, zs.n
    bwd.w1_off, bwd.b1_off, bwd.g1_off, bwd.n1 = fwd.w1_off, fwd.b1_off, fwd.g1_off, fwd.n1
    bwd.aux_pos_stash = bwd.aux_dir_stash = -1
    bwd.sigma_bias = fwd.sigma_bias

    # ------------------------------------------------------------------ weight-gradient units
    units: List[WgradUnit] = []
    _weight_units(units, L1.lin, d_1, 1024, aux_pos, 0, 3)
    _weight_units(units, L2.lin, d_2, 256, y1, 0, 1024)
    _weight_units(units, L3.lin, d_3, 128, y2, 0, 256)
    _weight_units(units, L4.lin, d_4, 128, y3, 0, 128)
    _weight_units_concat(units, L5.lin, d_5, 512, y4, 128, aux_pos, 3)
    _weight_units(units, L6.lin, d_6, 256, y5, 0, 512)
    _weight_units(units, L7.lin, d_7, 128, y6, 0, 256)
    _weight_units(units, L8.lin, d_8, 128, y7, 0, 128, bias=True)
    _weight_units(units, L8.lin, d_8 + 2, 1, y7, 0, 128, bias=True, out0=128)
    _weight_units_concat(units, C1.lin, d_c1, 256, ysum, 128, aux_dir, 3)
    _weight_units(units, C2.lin, d_head, 3, yc1, 0, 256, bias=True)
    for layer, dy0, zz in ((L1, d_1, z1), (L2, d_2, z2), (L3, d_3, z3), (L4, d_4, z4), (L5, d_5, z5), (L6, d_6, z6),
                           (L7, d_7, z7), (C1, d_c1, zc1)):
        _colsum_units(units, dy0, zz, layer)
    macs = sum(l.lin.in_f * l.lin.out_f for l in L + Lc)
    gauss = sum(l.lin.out_f for l in L + Lc if l.g_off >= 0)
    return CompiledGarf(fwd=fwd, bwd=bwd, pack_chunks=f.chunks + g.chunks, fwd_floats=f.floats, bwd_floats=g.floats,
                        wpack_units=g.w_units, units=units, has_rgb=True, macs_per_sample=macs, gauss_per_sample=gauss,
                        gauss_layers=[l for l in L + Lc if l.g_off >= 0])


def _first_layer_bwd(g: _Builder, L1: GaussLinear, L2: GaussLinear, z1: int, d_1: int):
    """d(y1) = dz2 W2 in 256-column blocks alternating R0 / R1 (the MMAs of block b + 1 run under the
    epilogue of block b), each multiplied by the Gaussian derivative of the first layer and written
    straight to the HBM stash (no later MMA reads it: d(position) is a rank-3 fp32 sum in the epilogue)."""
    n_blk = L1.lin.out_f // 256
    k_outs = [(64 * c, 64) for c in range(4)]

    def epilogue(blk, lag):
        g.step(NG_BSTEP_ACT, lag=lag, n_slabs=4, src_col=(R0, R1)[blk & 1], flags=NG_F_DIRECT | NG_F_FIRST_LAYER,
               coef_off=g.pack(L1.g_off + 256 * blk, 256, kind=_lib.PACK_GAUSS), z_stash=z1 + 4 * blk,
               y_stash=d_1 + 4 * blk, skip_src=1, gen_col0=256 * blk)

    for blk in range(n_blk):
        if blk == 1:
            g.step(NG_STEP_NONE, lag=1)
        elif blk >= 2:
            epilogue(blk - 2, 1)
        _mma_bwd(g, L2.lin, [0, 1, 2, 3], k_outs, 256 * blk, 256, (R0, R1)[blk & 1])
    if n_blk >= 2:
        epilogue(n_blk - 2, 1)
        g.nop()
    epilogue(n_blk - 1, 0)


def compile_proposal(L: List[GaussLinear]) -> CompiledGarf:
    """L = [L1 (3->512 G), L2 (512->256 G), L3 (256->128 G), L4 (128->1)] (garf/model_proposal.py:22-31);
    forward :55-56: softplus_8 of the last layer."""
    shapes = [(l.lin.in_f, l.lin.out_f) for l in L]
    if shapes != [(3, 512), (512, 256), (256, 128), (128, 1)]:
        raise RuntimeError(f"the fused GARF proposal program is laid out for the reference's layer sizes, got {shapes}")
    L1, L2, L3, L4 = L
    ys, zs = _Stash(), _Stash()
    aux_pos = ys.take(1)
    y1, z1 = ys.take(8), zs.take(8)
    y2, z2 = ys.take(4), zs.take(4)
    y3, z3 = ys.take(2), zs.take(2)
    f = _Builder()
    for blk in range(L1.lin.out_f // NG_GEN_COLS):
        pair = 2 * (blk & 1)
        f.step(NG_STEP_GEN, lag=1, n_slabs=2, out_slab=pair, gen_col0=blk * NG_GEN_COLS, y_stash=y1 + 2 * blk,
               z_stash=z1 + 2 * blk)
        _mma_fwd(f, L2.lin, [pair, pair + 1], blk * NG_GEN_COLS, 0, 256, R1, accumulate=blk > 0)
    _act_layer_fwd(f, L2, R1, 0, y2, z2)
    _mma_fwd(f, L3.lin, [0, 1, 2, 3], 0, 0, 128, R0)
    _act_layer_fwd(f, L3, R0, 0, y3, z3)
    _mma_fwd(f, L4.lin, [0, 1], 0, 0, 1, R1)
    f.step(NG_STEP_SIGMA, src_col=R1, bias_off=f.pack(L4.lin.b_off, 1, 16))
    fwd = NgProgram()
    f.finish(fwd)
    fwd.y_slabs_per_tile, fwd.z_slabs_per_tile = ys.n, zs.n
    fwd.w1_off, fwd.b1_off, fwd.g1_off, fwd.n1 = L1.lin.w_off, L1.lin.b_off, L1.g_off, L1.lin.out_f
    fwd.aux_pos_stash, fwd.aux_dir_stash = aux_pos, -1
    fwd.sigma_bias = 0.0

    ds = _Stash()
    d_head, d_3, d_2, d_1 = ds.take(1), ds.take(2), ds.take(4), ds.take(8)
    g = _Builder()
    g.w_units = f.w_units
    coef = lambda layer, o0, n: g.pack(layer.g_off + o0, n, _ceil(n, 64) * 64, kind=_lib.PACK_GAUSS)
    g.step(NG_BSTEP_HEAD, n_slabs=1, out_slab=0, flags=NG_F_SIGMA, y_stash=d_head)
    _mma_bwd(g, L4.lin, [0], [(0, 1)], 0, 128, R0)
    g.step(NG_BSTEP_ACT, n_slabs=2, out_slab=0, src_col=R0, coef_off=coef(L3, 0, 128), z_stash=z3, y_stash=d_3)
    _mma_bwd(g, L3.lin, [0, 1], [(0, 64), (64, 64)], 0, 256, R1)
    g.step(NG_BSTEP_ACT, n_slabs=4, out_slab=0, src_col=R1, coef_off=coef(L2, 0, 256), z_stash=z2, y_stash=d_2)
    _first_layer_bwd(g, L1, L2, z1, d_1)
    bwd = NgProgram()
    g.finish(bwd)
    bwd.y_slabs_per_tile, bwd.z_slabs_per_tile = ds.n, zs.n
    bwd.w1_off, bwd.b1_off, bwd.g1_off, bwd.n1 = fwd.w1_off, fwd.b1_off, fwd.g1_off, fwd.n1
    bwd.aux_pos_stash = bwd.aux_dir_stash = -1
    bwd.sigma_bias = 0.0

    units: List[WgradUnit] = []
    _weight_units(units, L1.lin, d_1, 512, aux_pos, 0, 3)
    _weight_units(units, L2.lin, d_2, 256, y1, 0, 512)
    _weight_units(units, L3.lin, d_3, 128, y2, 0, 256)
    _weight_units(units, L4.lin, d_head, 1, y3, 0, 128, bias=True)
    for layer, dy0, zz in ((L1, d_1, z1), (L2, d_2, z2), (L3, d_3, z3)):
        _colsum_units(units, dy0, zz, layer)
    macs = sum(l.lin.in_f * l.lin.out_f for l in L)
    gauss = sum(l.lin.out_f for l in L if l.g_off >= 0)
    return CompiledGarf(fwd=fwd, bwd=bwd, pack_chunks=f.chunks + g.chunks, fwd_floats=f.floats, bwd_floats=g.floats,
                        wpack_units=g.w_units, units=units, has_rgb=False, macs_per_sample=macs, gauss_per_sample=gauss,
                        gauss_layers=[l for l in L if l.g_off >= 0])
