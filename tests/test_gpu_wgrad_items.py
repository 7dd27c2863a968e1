"""The weight-gradient kernel driven directly through the C ABI with hand-made stashes and work items: the MMA
block (dW = dY^T X over a tile range), the bias column sums that ride on it, and the "z duty" / column-sum-only
items (bias + Gaussian-width sums from a Z stash: include/nerfb200_mlp.h) — the capability activations whose
parameter gradients do not reduce to dW and db need; the shipped GARF programs get their widths from
nerfb200_gauss_width_grad instead, which is checked here against the same direct sums."""
import ctypes as C

import numpy as np
import pytest
import torch as th

pytestmark = pytest.mark.gpu

SLAB = 16384


def _bf16_round(x: th.Tensor) -> th.Tensor:
    return x.to(th.bfloat16).to(th.float32)


def _pack_slabs(vals: th.Tensor, z_layout: bool = False) -> th.Tensor:
    """vals: (n_tiles, n_slabs, 128, 64) fp32 (already bf16-representable) -> uint8 stash in the slab layout
    (row * 128 + ((col / 8) ^ (row & 7)) * 16 + (col % 8) * 2), or the z stash's variant (sector halves in
    natural order: csrc/garf_kernels.cuh zstash_offset)."""
    T, S = vals.shape[:2]
    out = np.zeros((T, S, SLAB), dtype=np.uint8)
    bits = (vals.to(th.bfloat16).view(th.int16).numpy().astype(np.uint16))
    rows, cols = np.meshgrid(np.arange(128), np.arange(64), indexing="ij")
    if z_layout:
        off = rows * 128 + ((((cols >> 4) ^ ((rows & 7) >> 1)) & 3) << 5) + ((cols & 15) << 1)
    else:
        off = rows * 128 + (((cols >> 3) ^ (rows & 7)) << 4) + ((cols & 7) << 1)
    flat = out.reshape(T, S, SLAB)
    flat[:, :, off] = (bits & 0xff).astype(np.uint8)
    flat[:, :, off + 1] = (bits >> 8).astype(np.uint8)
    return th.from_numpy(out.reshape(-1))


def test_wgrad_items_mma_bias_zduty_and_width_identity(cuda):
    from nerf_experiments_b200 import _lib
    from nerf_experiments_b200._lib import NbGaussLayer, NbWgradItem, check, lib
    from nerf_experiments_b200.mlp_program import to_device_array
    g = th.Generator().manual_seed(0)
    T, n_dy, n_x, n_z = 5, 4, 2, 4
    dy = _bf16_round(th.randn((T, n_dy, 128, 64), generator=g) * 0.1)
    x = _bf16_round(th.randn((T, n_x, 128, 64), generator=g))
    z = _bf16_round(th.randn((T, n_z, 128, 64), generator=g))
    M, K = 256, 128
    # parameter buffer: W (M, K) | b (M) | s (M);  gradient buffer of the same layout
    w_off, b_off, g_off = 0, M * K, M * K + M
    params = th.randn(M * K + 2 * M, generator=g)
    params[g_off:] = params[g_off:].abs() + 0.5
    d_params = th.zeros_like(params)
    items = [
        # the whole layer as one MMA item over tiles 1..4, bias riding on it, z duty for dY slabs 1..3
        NbWgradItem(tile_begin=1, tile_end=5, n_dy_slabs=4, n_x_slabs=2, dy_slab=0, x_slab=0, m_real=M, n_real=K,
                    dst=w_off, ld=K, bias_dst=-1, mode=_lib.WGRAD_MMA, coef_dst=g_off, z_slab=1, n_z_slabs=3, z_first=1,
                    zbias_dst=b_off),
        # tile 0 as a second item (no z duty, bias of all four slabs)
        NbWgradItem(tile_begin=0, tile_end=1, n_dy_slabs=4, n_x_slabs=2, dy_slab=0, x_slab=0, m_real=M, n_real=K,
                    dst=w_off, ld=K, bias_dst=b_off, mode=_lib.WGRAD_MMA),
        # what is left of the column sums as column-sum-only items: slab 0 over tiles 1..4, slabs 1..3 z part of tile 0
        NbWgradItem(tile_begin=1, tile_end=5, n_dy_slabs=1, n_x_slabs=0, dy_slab=0, x_slab=0, m_real=64, n_real=0,
                    dst=0, ld=0, bias_dst=-1, mode=_lib.WGRAD_COLSUM, coef_dst=g_off, z_slab=0, n_z_slabs=1, z_first=0,
                    zbias_dst=b_off),
    ]
    dev_items = to_device_array(items, NbWgradItem, cuda)
    dy_d, x_d, z_d = _pack_slabs(dy).to(cuda), _pack_slabs(x).to(cuda), _pack_slabs(z, z_layout=True).to(cuda)
    p_d, g_d = params.to(cuda), d_params.to(cuda)
    stream = th.cuda.current_stream().cuda_stream
    check(lib().nerfb200_mlp_wgrad(dev_items.data_ptr(), len(items), x_d.data_ptr(), n_x, dy_d.data_ptr(), n_dy,
                                   z_d.data_ptr(), n_z, p_d.data_ptr(), g_d.data_ptr(), stream), "mlp_wgrad")
    th.cuda.synchronize()
    got = g_d.cpu()
    dyf = dy.permute(0, 2, 1, 3).reshape(T * 128, n_dy * 64)          # (samples, M)
    xf = x.permute(0, 2, 1, 3).reshape(T * 128, n_x * 64)
    zf = z.permute(0, 2, 1, 3).reshape(T * 128, n_z * 64)
    assert th.allclose(got[w_off: w_off + M * K].view(M, K), dyf.T @ xf, rtol=1e-4, atol=1e-4)
    # bias: item 2 covers every slab on tile 0; item 1's z duty slabs 1..3 and item 3 slab 0 cover tiles 1..4
    assert th.allclose(got[b_off: b_off + M], dyf.sum(0), rtol=1e-4, atol=1e-4)
    s = params[g_off:]
    zdz = (zf[128:] * dyf[128:]).sum(0)                               # tiles 1..4 only: tile 0 has no z duty here
    assert th.allclose(got[g_off: g_off + M], zdz * s / (s * s + 1e-6), rtol=1e-4, atol=1e-4)

    # the width identity: with z = W x + b exactly, sum z dz = W . dW + b db
    zlin = xf @ params[w_off: w_off + M * K].view(M, K).T + params[b_off: b_off + M]
    layers = to_device_array([NbGaussLayer(w_off=w_off, b_off=b_off, g_off=g_off, in_f=K, out_f=M)], NbGaussLayer, cuda)
    g2 = th.zeros_like(params)
    g2[w_off: w_off + M * K] = (dyf.T @ xf).reshape(-1)
    g2[b_off: b_off + M] = dyf.sum(0)
    g2[g_off:] = 7.0                                                   # what the buffer held before must survive
    g2_d = g2.to(cuda)
    check(lib().nerfb200_gauss_width_grad(layers.data_ptr(), 1, M, p_d.data_ptr(), g2_d.data_ptr(), 1.0, stream), "width")
    th.cuda.synchronize()
    direct = (zlin * dyf).sum(0) * s / (s * s + 1e-6)
    assert th.allclose(g2_d.cpu()[g_off:] - 7.0, direct, rtol=2e-3, atol=2e-3)
    check(lib().nerfb200_gauss_width_grad(layers.data_ptr(), 1, M, p_d.data_ptr(), g2_d.data_ptr(), -1.0, stream), "width")
    th.cuda.synchronize()
    assert th.allclose(g2_d.cpu()[g_off:], th.full((M,), 7.0), atol=1e-3)
