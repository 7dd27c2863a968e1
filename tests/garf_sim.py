"""CPU interpreter of the GARF tile programs (include/nerfb200_garf.h) — TEST INFRASTRUCTURE.

Executes what csrc/garf_fwd.cu / garf_bwd.cu / mlp_pack.cu / mlp_wgrad.cu do with a compiled program,
in torch on the CPU with the kernels' roundings (bf16 operands and stashes, fp32 accumulation), one
128-sample tile at a time. It validates the host-side compiler (nerf_experiments_b200/garf_program.py:
op / step order, weight-image and float packing descriptors, stash layout, weight-gradient units)
without a GPU; the GPU tests then compare the real kernels with the oracle.
"""
import math

import torch as th

from nerf_experiments_b200 import _lib

LOG2E = 1.4426950408889634
TWO_LN2 = 1.3862943611198906


def bf(x: th.Tensor) -> th.Tensor:
    return x.to(th.bfloat16).to(th.float32)


def pack_weights(params: th.Tensor, chunks, n_units: int) -> th.Tensor:
    """(n_units * 8, 64) bf16-rounded image rows (unit = 1024 B = 8 rows of 64 bf16)."""
    w = th.zeros((n_units * 8, 64))
    for ch in chunks:
        r = th.arange(ch.rows_padded).view(-1, 1)
        c = th.arange(64).view(1, -1)
        ok = (r < ch.n_rows) & (c < ch.n_cols)
        idx = (ch.base + r * ch.row_stride + c * ch.col_stride).clamp(0, params.numel() - 1)
        vals = th.where(ok, params[idx], th.zeros(()))
        rows = ch.dst_off * 8 + ch.dst_row0 + th.arange(ch.rows_padded) * ch.dst_row_step
        w[rows] = bf(vals)
    return w


def pack_floats(params: th.Tensor, descs, n_floats: int) -> th.Tensor:
    out = th.zeros(max(n_floats, 1))
    for d in descs:
        stride = d.stride if d.stride > 0 else 1
        v = params[d.base + th.arange(d.n) * stride]
        if d.kind == _lib.PACK_GAUSS:
            v = -(v * v + 1e-6) * LOG2E
        out[d.dst_off: d.dst_off + d.n] = v
        out[d.dst_off + d.n: d.dst_off + d.n_padded] = 0.0
    return out


def _run_op(op, slabs, tmem, wpack):
    if op.n_chunks == 0:
        return
    for b in range(op.n_blocks):
        blk = op.blocks[b]
        acc = tmem[:, blk.tmem_col: blk.tmem_col + blk.n].clone() if op.accumulate else th.zeros((128, blk.n))
        for c in range(op.n_chunks):
            k = 16 * op.k16[c]
            A = slabs[op.a_slab[c]][:, :k]
            rows = op.w_off[c] * 8 + blk.row0
            B = wpack[rows: rows + blk.n, :k]
            acc = acc + A @ B.T
        tmem[:, blk.tmem_col: blk.tmem_col + blk.n] = acc


def forward_tile(prog, params, wpack, floats, pos, dirs, training=True):
    """pos, dirs: (128, 3) fp32. Returns sigma (128,), rgb (128,3) or None, y_stash / z_stash dicts
    {slab index: (128, 64) tensor}."""
    slabs = [th.zeros((128, 64)) for _ in range(_lib.NG_N_SLABS)]
    tmem = th.zeros((128, 512))
    y_stash, z_stash = {}, {}
    sigma, rgb = th.zeros(128), None
    if training:
        aux = th.zeros((128, 64)); aux[:, :3] = bf(pos)
        y_stash[prog.aux_pos_stash] = aux
        if prog.aux_dir_stash >= 0:
            aux = th.zeros((128, 64)); aux[:, :3] = bf(dirs)
            y_stash[prog.aux_dir_stash] = aux
    W1 = params[prog.w1_off: prog.w1_off + prog.n1 * 3].view(prog.n1, 3)
    b1 = params[prog.b1_off: prog.b1_off + prog.n1]
    s1 = params[prog.g1_off: prog.g1_off + prog.n1]
    for k in range(prog.n_ops + 1):
        st = prog.steps[k]
        n = 64 * st.n_slabs
        if st.kind == _lib.NG_STEP_GEN:
            cols = slice(st.gen_col0, st.gen_col0 + 128)
            z = pos @ W1[cols].T + b1[cols]
            y = th.exp2(z * z * (-(s1[cols] ** 2 + 1e-6) * LOG2E))
            for j in range(2):
                slabs[st.out_slab + j] = bf(y[:, 64 * j: 64 * j + 64])
                if training:
                    y_stash[st.y_stash + j] = slabs[st.out_slab + j].clone()
                    z_stash[st.z_stash + j] = bf(z[:, 64 * j: 64 * j + 64])
        elif st.kind == _lib.NG_STEP_ACT:
            z = tmem[:, st.src_col: st.src_col + n] + floats[st.bias_off: st.bias_off + n]
            if st.skip_src:
                x3 = pos if st.skip_src == 1 else dirs
                z = z + x3 @ floats[st.skip_off: st.skip_off + 3 * n].view(3, n)
            y = th.exp2(z * z * floats[st.coef_off: st.coef_off + n])
            for j in range(st.n_slabs):
                slabs[st.out_slab + j] = bf(y[:, 64 * j: 64 * j + 64])
                if training:
                    if st.y_stash >= 0:
                        y_stash[st.y_stash + j] = slabs[st.out_slab + j].clone()
                    if st.z_stash >= 0:
                        z_stash[st.z_stash + j] = bf(z[:, 64 * j: 64 * j + 64])
        elif st.kind == _lib.NG_STEP_LINEAR:
            a = tmem[:, st.src_col: st.src_col + n] + floats[st.bias_off: st.bias_off + n]
            if st.res_slab >= 0:
                a = a + th.cat([slabs[st.res_slab + j] for j in range(st.n_slabs)], dim=1)
            if st.flags & _lib.NG_F_SIGMA:
                pre = tmem[:, st.sigma_col] + floats[st.bias_off + n]
                sigma = th.nn.functional.softplus(pre + prog.sigma_bias, beta=1.0, threshold=8.0)
            for j in range(st.n_slabs):
                slabs[st.out_slab + j] = bf(a[:, 64 * j: 64 * j + 64])
                if training and st.y_stash >= 0:
                    y_stash[st.y_stash + j] = slabs[st.out_slab + j].clone()
        elif st.kind == _lib.NG_STEP_RGB:
            rgb = th.sigmoid(tmem[:, st.src_col: st.src_col + 3] + floats[st.bias_off: st.bias_off + 3])
        elif st.kind == _lib.NG_STEP_SIGMA:
            sigma = th.nn.functional.softplus(tmem[:, st.src_col] + floats[st.bias_off] + prog.sigma_bias,
                                              beta=1.0, threshold=8.0)
        if k < prog.n_ops:
            _run_op(prog.ops[k], slabs, tmem, wpack)
    return sigma, rgb, y_stash, z_stash


def backward_tile(prog, params, wpack, floats, pos, dirs, sigma, rgb, g_sigma, g_rgb, z_stash, want=True):
    """Returns dy_stash dict, d_pos (128,3), d_dir (128,3)."""
    slabs = [th.zeros((128, 64)) for _ in range(_lib.NG_N_SLABS)]
    tmem = th.zeros((128, 512))
    dy = {}
    hold = [th.zeros((128, 64)), th.zeros((128, 64))]
    dpos, ddir = th.zeros((128, 3)), th.zeros((128, 3))
    dsp = g_sigma * th.where(sigma > 8.0, th.ones_like(sigma), -th.expm1(-sigma))
    W1 = params[prog.w1_off: prog.w1_off + prog.n1 * 3].view(prog.n1, 3)
    for k in range(prog.n_ops + 1):
        st = prog.steps[k]
        n = 64 * st.n_slabs
        if st.kind == _lib.NG_BSTEP_HEAD:
            h = th.zeros((128, 64))
            if st.flags & _lib.NG_F_SIGMA:
                h[:, 0] = bf(dsp)
            else:
                h[:, :3] = bf(g_rgb * rgb * (1 - rgb))
            slabs[st.out_slab] = h
            dy[st.y_stash] = h.clone()
        elif st.kind == _lib.NG_BSTEP_ACT:
            g = tmem[:, st.src_col: st.src_col + n].clone()
            if st.flags & _lib.NG_F_HOLD_ADD:
                g = g + th.cat(hold, dim=1)[:, :n]
            z = th.cat([z_stash[st.z_stash + j] for j in range(st.n_slabs)], dim=1)
            t = z * floats[st.coef_off: st.coef_off + n]
            dz = g * th.exp2(z * t) * (t * TWO_LN2)
            dzb = bf(dz)
            if want and st.skip_src:
                if st.flags & _lib.NG_F_FIRST_LAYER:
                    wk = W1[st.gen_col0: st.gen_col0 + n]                 # (n, 3)
                    acc3 = dz @ wk
                else:
                    acc3 = dz @ floats[st.skip_off: st.skip_off + 3 * n].view(3, n).T
                if st.skip_src == 2:
                    ddir += acc3
                else:
                    dpos += acc3
            for j in range(st.n_slabs):
                piece = dzb[:, 64 * j: 64 * j + 64]
                if not (st.flags & _lib.NG_F_DIRECT):
                    slabs[st.out_slab + j] = piece.clone()
                dy[st.y_stash + j] = piece.clone()
        elif st.kind == _lib.NG_BSTEP_PLAIN:
            g = bf(tmem[:, st.src_col: st.src_col + n])
            for j in range(st.n_slabs):
                piece = g[:, 64 * j: 64 * j + 64]
                slabs[st.out_slab + j] = piece.clone()
                dy[st.y_stash + j] = piece.clone()
                if (st.flags & _lib.NG_F_HOLD_SAVE) and j < 2:
                    hold[j] = piece.clone()
            if st.flags & _lib.NG_F_SIGMA:
                s = th.zeros((128, 64)); s[:, 0] = bf(dsp)
                slabs[st.out_slab + st.n_slabs] = s
                dy[st.y_stash + st.n_slabs] = s.clone()
        if k < prog.n_ops:
            _run_op(prog.ops[k], slabs, tmem, wpack)
    return dy, dpos, ddir


def weight_gradients(units, params, grad, y_stash, dy_stash, z_stash):
    """Accumulates one tile's contribution of every weight-gradient unit into `grad` (flat fp32)."""
    for u in units:
        dyv = th.cat([dy_stash[u.dy_slab + j] for j in range(u.n_dy_slabs)], dim=1)[:, :u.m_real]
        if u.n_z_slabs > 0:      # z duty: bias and Gaussian-width gradients of dY slabs z_first .. z_first + n_z - 1
            c0, c1 = 64 * u.z_first, min(64 * (u.z_first + u.n_z_slabs), u.m_real)
            zv = th.cat([z_stash[u.z_slab + j] for j in range(u.n_z_slabs)], dim=1)[:, :c1 - c0]
            if u.zbias_dst >= 0:
                grad[u.zbias_dst + c0: u.zbias_dst + c1] += dyv[:, c0:c1].sum(0)
            if u.coef_dst >= 0:
                s = params[u.coef_dst + c0: u.coef_dst + c1]
                grad[u.coef_dst + c0: u.coef_dst + c1] += (zv * dyv[:, c0:c1]).sum(0) * s / (s * s + 1e-6)
        if u.mode == _lib.WGRAD_COLSUM:
            continue
        xs = [u.x_slab + j for j in range(u.n_x_slabs)]
        if u.x2_slab >= 0:
            xs[-1] = u.x2_slab
        xv = th.cat([y_stash[j] for j in xs], dim=1)[:, :u.n_real]
        dw = dyv.T @ xv                                                       # (m_real, n_real)
        for m in range(u.m_real):
            grad[u.dst + m * u.ld: u.dst + m * u.ld + u.n_real] += dw[m]
        if u.bias_dst >= 0:
            grad[u.bias_dst: u.bias_dst + u.m_real] += dyv.sum(0)


def run_network(cg, params, pos, dirs, g_sigma, g_rgb):
    """Whole batch (N multiple of 128 not required): forward, backward and parameter gradients."""
    N = pos.shape[0]
    wpack = pack_weights(params, cg.pack_chunks, cg.wpack_units)
    f_fwd = pack_floats(params, cg.fwd_floats, cg.fwd.n_floats)
    f_bwd = pack_floats(params, cg.bwd_floats, cg.bwd.n_floats)
    sig_out, rgb_out = th.zeros(N), th.zeros((N, 3))
    grad = th.zeros_like(params)
    dpos_out, ddir_out = th.zeros((N, 3)), th.zeros((N, 3))
    for t0 in range(0, N, 128):
        idx = th.arange(t0, t0 + 128).clamp(max=N - 1)
        valid = (th.arange(t0, t0 + 128) < N)
        p, d = pos[idx], (dirs[idx] if dirs is not None else pos[idx])
        sigma, rgb, ys, zs = forward_tile(cg.fwd, params, wpack, f_fwd, p, d)
        gs = th.where(valid, g_sigma[idx], th.zeros(()))
        gr = None
        if rgb is not None:
            gr = th.where(valid[:, None], g_rgb[idx], th.zeros(()))
        dy, dp, dd = backward_tile(cg.bwd, params, wpack, f_bwd, p, d, sigma, rgb, gs, gr, zs)
        weight_gradients(cg.units, params, grad, ys, dy, zs)
        nv = int(valid.sum())
        sig_out[t0: t0 + nv] = sigma[:nv]
        if rgb is not None:
            rgb_out[t0: t0 + nv] = rgb[:nv]
        dpos_out[t0: t0 + nv] = dp[:nv]
        ddir_out[t0: t0 + nv] = dd[:nv]
    # Gaussian widths from the weight / bias gradients (nerfb200_gauss_width_grad): sum z dz = W . dW + b db
    for layer in cg.gauss_layers:
        lin = layer.lin
        W = params[lin.w_off: lin.w_off + lin.out_f * lin.in_f].view(lin.out_f, lin.in_f)
        dW = grad[lin.w_off: lin.w_off + lin.out_f * lin.in_f].view(lin.out_f, lin.in_f)
        b, db = params[lin.b_off: lin.b_off + lin.out_f], grad[lin.b_off: lin.b_off + lin.out_f]
        sdev = params[layer.g_off: layer.g_off + lin.out_f]
        grad[layer.g_off: layer.g_off + lin.out_f] += ((W * dW).sum(1) + b * db) * sdev / (sdev * sdev + 1e-6)
    return sig_out, (rgb_out if cg.has_rgb else None), grad, dpos_out, ddir_out
