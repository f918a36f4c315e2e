"""GPU test of the tcgen05 building blocks the fused kernel rests on (srcnn_debug_tc_selftest):
  1. one SS MMA (M128 N256 K16) whose A operand is the FP16 Y tile in the chunked no-swizzle layout,
     started at a kernel-row offset, against a Toeplitz-style B tile;
  2. ReLU + FP16 pack + tcgen05.st, then TS MMAs (A from TMEM) for a 64->32 GEMM.
A wrong descriptor / TMEM layout assumption shows up here, isolated from the pipeline logic."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TILE_ROWS = 136


def _pack_b(mat):
    """[N][16] float -> SWIZZLE_NONE K-major image: (n,k) at (k//8)*(N*16) + n*16 + (k%8)*2 bytes."""
    n = mat.shape[0]
    img = np.zeros((2, n, 8), np.float16)
    for kc in range(2):
        img[kc] = mat[:, kc * 8:(kc + 1) * 8].astype(np.float16)
    return img.reshape(-1).view(np.uint8)


@pytest.mark.parametrize("row_off", [0, 3, 8])
def test_ss_and_ts_mma(engine, row_off):
    import torch
    rng = np.random.default_rng(row_off)
    ytile = rng.integers(0, 256, (TILE_ROWS, 16)).astype(np.float32)          # rows x 16 pixels
    a_img = np.zeros((2, TILE_ROWS, 8), np.float16)                            # [chunk][row][8 px]
    a_img[0], a_img[1] = ytile[:, :8], ytile[:, 8:]
    b1 = (rng.standard_normal((256, 16)) * 0.25).astype(np.float16).astype(np.float32)
    b2 = (rng.standard_normal((32, 64)) * 0.25).astype(np.float16).astype(np.float32)
    b2_img = np.concatenate([_pack_b(b2[:, ks * 16:(ks + 1) * 16]) for ks in range(4)])

    dev = torch.device("cuda:0")
    t_a = torch.from_numpy(a_img.reshape(-1).view(np.uint8).copy()).to(dev)
    t_b1 = torch.from_numpy(_pack_b(b1).copy()).to(dev)
    t_b2 = torch.from_numpy(b2_img.copy()).to(dev)
    d1 = torch.zeros((128, 256), dtype=torch.float32, device=dev)
    d2 = torch.zeros((128, 32), dtype=torch.float32, device=dev)
    fn = engine.L.srcnn_debug_tc_selftest
    fn.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p]
    torch.cuda.synchronize()
    rc = fn(engine.ctx, t_a.data_ptr(), t_b1.data_ptr(), t_b2.data_ptr(), row_off, d1.data_ptr(), d2.data_ptr())
    assert rc == 0
    engine.sync()

    a = ytile[row_off:row_off + 128]                      # lane r reads tile row r + row_off
    ref1 = a.astype(np.float64) @ b1.T.astype(np.float64)
    got1 = d1.cpu().numpy()
    assert np.allclose(got1, ref1, rtol=1e-5, atol=1e-2), np.abs(got1 - ref1).max()

    act = np.maximum(got1[:, :64], 0).astype(np.float16).astype(np.float64)   # what the kernel packed
    ref2 = act @ b2.T.astype(np.float64)
    got2 = d2.cpu().numpy()
    assert np.allclose(got2, ref2, rtol=1e-4, atol=0.5), np.abs(got2 - ref2).max()
