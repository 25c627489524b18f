"""Host-side logic of the brake-model stem path (no GPU): the K order produced by ops.pack_stem_weights and the
staged-row addressing of csrc/stem.cu, restated in numpy, reproduce Normalize + conv 7x7/s2/p3 of the reference
(team_code_v2/models/rgb.py:66-70, lav/models/resnet.py:178,235-238) on the side-by-side camera image."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lav_b200 import ops

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
QW, ROWS, IN_ROWS = 792, 8, 21          # kStemQW, kStemRows, kStemInRows of csrc/stem.cu


def emulate(img, wk):
    """numpy restatement of stem7x7_u8_kernel's indexing: per block stage IN_ROWS x QW normalised elements, then
    A[pixel p][k] = S[2*rr + min(k // 22, 6)][6*p + k - 22*min(k // 22, 6)] and out = A @ wk.T."""
    b, ncam, h, cam_w, _ = img.shape
    W = ncam * cam_w
    ho, wo = (h - 1) // 2 + 1, (W - 1) // 2 + 1
    na = [1 / (255 * s) for s in STD]
    nb = [-m / s for m, s in zip(MEAN, STD)]
    out = np.zeros((b, ho, wo, 64), np.float32)
    flat = img.reshape(-1)
    cam_bytes, row_bytes = cam_w * 3, W * 3
    ks = np.arange(160)
    ky = np.minimum(ks // 22, 6)
    kp = ks - 22 * ky
    for bi in range(b):
        for oy0 in range(0, ho, ROWS):
            for ox0 in range(0, wo, 128):
                S = np.full((IN_ROWS, QW), np.nan, np.float32)
                base = 6 * ox0 - 9
                w_first = (base - 3) >> 2
                for r in range(IN_ROWS):
                    iy = 2 * oy0 - 3 + r
                    for wi in range((QW + 6) // 4 + 1):
                        byte0 = 4 * (w_first + wi)
                        ok = 0 <= iy < h and 0 <= byte0 < row_bytes
                        if ok:
                            cam = int(byte0 >= cam_bytes) + int(byte0 >= 2 * cam_bytes) + int(byte0 >= 3 * cam_bytes)
                            addr = bi * ncam * h * cam_bytes + (cam * h + iy) * cam_bytes + (byte0 - cam * cam_bytes)
                            assert addr % 4 == 0
                        for j in range(4):
                            q = byte0 + j - base
                            if 0 <= q < QW:
                                c = (byte0 + j) % 3
                                S[r, q] = flat[addr + j] * na[c] + nb[c] if ok else 0.0
                assert not np.isnan(S).any()            # every staged element the MMA loop can touch is initialised
                for rr in range(ROWS):
                    oy = oy0 + rr
                    if oy >= ho:
                        break
                    for p in range(min(128, wo - ox0)):
                        out[bi, oy, ox0 + p] = wk @ S[2 * rr + ky, 6 * p + kp]
    return out


@pytest.mark.parametrize("shape", [(1, 3, 18, 12), (1, 1, 31, 52), (1, 3, 9, 88)])
def test_stem_k_order_and_staging(shape):
    b, ncam, h, cam_w = shape
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (b, ncam, h, cam_w, 3), dtype=np.uint8)
    w = torch.from_numpy(rng.standard_normal((64, 3, 7, 7)).astype(np.float32) * 0.1)
    wk = ops.pack_stem_weights(w)
    assert wk.shape == (64, 160) and wk.dtype == ops.h16()
    wq = wk.float().numpy()
    assert np.all(wq.reshape(64, -1)[:, 154:] == 0) and np.all(wq[:, :154].reshape(64, 7, 22)[:, :, 21] == 0)
    got = emulate(img, wq)
    wide = torch.from_numpy(img).permute(0, 2, 1, 3, 4).reshape(b, h, ncam * cam_w, 3).permute(0, 3, 1, 2).float()
    x = (wide / 255. - torch.tensor(MEAN)[None, :, None, None]) / torch.tensor(STD)[None, :, None, None]
    ref = F.conv2d(x, w.to(ops.h16()).float(), None, 2, 3).permute(0, 2, 3, 1).numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
