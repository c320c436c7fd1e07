"""-m gpu: every layer kind of the dense model path, through the kernel the execution planner picks for the
shape (w2x_run_conv_layer, include/w2x_dev.h), against torch.nn.functional in fp32 on the CPU -- an implementation that shares
nothing with the repo.  Inputs and weights are fp16-representable, the kernels accumulate in fp32 and store fp16, so the result
must agree with the fp32 reference to fp16 rounding of the output (|err| <= 2^-11 relative + accumulation-order noise).
The packing below restates csrc/model_pack.h's operand layouts independently of the C++ packer."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand16(rng, shape, scale):
    return (rng.uniform(-1, 1, size=shape) * scale).astype(np.float16)


def _ceil16(n):
    return (n + 15) // 16 * 16


def _nhwc(t):  # torch NCHW f32 -> numpy NHWC
    return t.permute(0, 2, 3, 1).contiguous().numpy()


def _lrelu(t):
    return F.leaky_relu(t, 0.1)


def _case(kind, n, h, w, cin, cout, seed=7):
    """-> (x NHWC fp16, packed weights, bias[npad], skip or None, reference NHWC f32 [n][oh][ow][cout or 3])"""
    rng = np.random.default_rng(seed)
    x = _rand16(rng, (n, h, w, cin), 1.0)
    xt = torch.from_numpy(x.astype(np.float32)).permute(0, 3, 1, 2)
    if kind == 0:      # conv3x3 valid + bias + LeakyReLU(0.1)
        wt = _rand16(rng, (cout, cin, 3, 3), 1 / np.sqrt(9 * cin))
        b = rng.uniform(-0.1, 0.1, size=cout).astype(np.float32)
        ref = _lrelu(F.conv2d(xt, torch.from_numpy(wt.astype(np.float32)), torch.from_numpy(b)))
        wp = np.zeros((_ceil16(cout), 9 * cin), np.float16)
        wp[:cout] = wt.transpose(0, 2, 3, 1).reshape(cout, 9 * cin)           # k = (ky*3+kx)*cin + ci
        bp = np.zeros(_ceil16(cout), np.float32); bp[:cout] = b
        return x, wp, bp, None, _nhwc(ref)
    if kind == 1:      # conv2x2 stride 2 + bias + LeakyReLU
        wt = _rand16(rng, (cout, cin, 2, 2), 1 / np.sqrt(4 * cin))
        b = rng.uniform(-0.1, 0.1, size=cout).astype(np.float32)
        ref = _lrelu(F.conv2d(xt, torch.from_numpy(wt.astype(np.float32)), torch.from_numpy(b), stride=2))
        wp = np.zeros((_ceil16(cout), 4 * cin), np.float16)
        wp[:cout] = wt.transpose(0, 2, 3, 1).reshape(cout, 4 * cin)           # k = (dy*2+dx)*cin + ci
        bp = np.zeros(_ceil16(cout), np.float32); bp[:cout] = b
        return x, wp, bp, None, _nhwc(ref)
    if kind == 2:      # ConvTranspose 2x2 stride 2 + bias + LeakyReLU, then + skip cropped by 4
        wt = _rand16(rng, (cin, cout, 2, 2), 1 / np.sqrt(cin))                # torch ConvTranspose2d weight [cin][cout][kh][kw]
        b = rng.uniform(-0.1, 0.1, size=cout).astype(np.float32)
        skip = _rand16(rng, (n, 2 * h + 8, 2 * w + 8, cout), 0.5)
        up = _lrelu(F.conv_transpose2d(xt, torch.from_numpy(wt.astype(np.float32)), torch.from_numpy(b), stride=2))
        st = torch.from_numpy(skip.astype(np.float32)).permute(0, 3, 1, 2)
        ref = up + st[:, :, 4:-4, 4:-4]
        wp = wt.transpose(2, 3, 1, 0).reshape(4 * cout, cin)                  # row = (dy*2+dx)*cout + co, k = ci
        bp = np.tile(b, 4)
        return x, wp, bp, skip, _nhwc(ref)
    if kind == 3:      # ConvTranspose 4x4 stride 2 padding 3 -> 3 channels, no activation
        wt = _rand16(rng, (cin, 3, 4, 4), 1 / np.sqrt(4 * cin))
        b = rng.uniform(-0.1, 0.1, size=3).astype(np.float32)
        ref = F.conv_transpose2d(xt, torch.from_numpy(wt.astype(np.float32)), torch.from_numpy(b), stride=2, padding=3)
        wp = np.zeros((16, 4 * cin), np.float16)
        bp = np.zeros(16, np.float32)
        for py in range(2):
            for px in range(2):
                for co in range(3):
                    row = (py * 2 + px) * 4 + co
                    bp[row] = b[co]
                    for wy in range(2):
                        for wx in range(2):
                            ky, kx = 2 + py - 2 * wy, 2 + px - 2 * wx
                            wp[row, (wy * 2 + wx) * cin:(wy * 2 + wx + 1) * cin] = wt[:, co, ky, kx]
        return x, wp, bp, None, _nhwc(ref)
    if kind == 4:      # conv3x3 -> 3 channels + residual cropped by 20 + clamp [0, 1]
        wt = _rand16(rng, (3, cin, 3, 3), 1 / np.sqrt(9 * cin))
        b = rng.uniform(0.2, 0.6, size=3).astype(np.float32)
        skip = np.zeros((n, h - 2 + 40, w - 2 + 40, 4), np.float16)
        skip[..., :3] = _rand16(rng, (n, h - 2 + 40, w - 2 + 40, 3), 0.5)
        st = torch.from_numpy(skip[..., :3].astype(np.float32)).permute(0, 3, 1, 2)
        ref = torch.clamp(F.conv2d(xt, torch.from_numpy(wt.astype(np.float32)), torch.from_numpy(b)) + st[:, :, 20:-20, 20:-20], 0, 1)
        wp = np.zeros((16, 9 * cin), np.float16)
        wp[:3] = wt.transpose(0, 2, 3, 1).reshape(3, 9 * cin)
        bp = np.zeros(16, np.float32); bp[:3] = b
        return x, wp, bp, skip, _nhwc(ref)
    raise ValueError(kind)


# (kind, head kernel?, n, h, w, cin, cout) -- the real layer shapes of CUNet / UpCUNet at tile 256 are marked
CASES = [
    (0, 0, 1, 20, 24, 64, 64),     # one pixel tile with edges (patch kernel)
    (0, 0, 8, 100, 100, 64, 64),   # > 148 tiles: persistent loop, ring wrap, double-buffered staging
    (0, 0, 1, 60, 52, 128, 64),    # two 64-channel K chunks
    (0, 0, 3, 45, 77, 64, 128),    # two output-channel slices
    (0, 0, 2, 61, 45, 32, 64),     # cin 32: 64-byte swizzle path, odd sizes
    (0, 0, 1, 126, 126, 64, 128),  # unet1.conv2.conv.0 at tile 256
    (0, 0, 1, 117, 117, 128, 256),  # unet2.conv3.conv.0 at tile 256 (odd extent, nSplit = 4)
    (0, 0, 1, 115, 115, 256, 128),  # unet2.conv3.conv.2 at tile 256 (K = 2304)
    (0, 0, 1, 226, 226, 128, 64),   # unet2.conv4.conv.0 at tile 256
    (1, 0, 2, 60, 60, 64, 64),     # down 2x2 s2
    (1, 0, 1, 234, 234, 128, 128),  # unet2.conv2_down at tile 256
    (2, 0, 2, 26, 26, 64, 64),     # convT 2x2 + skip, N = 256
    (2, 0, 1, 17, 17, 128, 128),   # N = 512 -> two N tiles
    (2, 0, 1, 113, 113, 128, 128),  # unet2.conv3_up at tile 256
    (3, 0, 2, 50, 50, 64, 3),      # convT 4x4 s2 p3 head (generic kernel)
    (4, 0, 2, 58, 58, 64, 3),      # image head + z1 crop + clamp (generic kernel)
    (4, 1, 2, 58, 58, 64, 3),      # image head kernel: partial tiles on both edges
    (4, 1, 3, 26, 146, 64, 3),     # image head kernel: exact multiples of the 8 x 16 tile, three images
    (3, 1, 2, 50, 50, 64, 3),      # convT 4x4 head kernel: partial tiles
    (3, 1, 3, 33, 17, 64, 3),      # convT 4x4 head kernel: exactly 32 x 16 GEMM pixels, three images
]


@pytest.mark.parametrize("kind,head,n,h,w,cin,cout", CASES)
def test_layer_matches_torch_fp32(kind, head, n, h, w, cin, cout, built_lib):
    import w2x
    x, wp, bp, skip, ref = _case(kind, n, h, w, cin, cout)
    got = w2x.run_conv_layer(kind, x, wp, bp, cout, skip=skip, head=bool(head)).astype(np.float32)
    c = ref.shape[-1]
    assert got.shape[:3] == ref.shape[:3]
    err = np.abs(got[..., :c] - ref)
    # fp16 storage of the result: half an ulp = 2^-11 relative; fp32 accumulation order adds ~1e-6 * sqrt(K)
    tol = np.abs(ref) * 2.0 ** -10 + 2e-4
    assert (err <= tol).all(), (float(err.max()), float((err / tol).max()))
    if got.shape[-1] > c:
        assert (got[..., c:] == 0).all()   # the padding channel of NHWC4 heads stays zero


@pytest.mark.parametrize("kind,n,h,w,cin,cout", [(0, 2, 40, 40, 64, 64), (2, 1, 20, 20, 64, 64), (4, 1, 58, 58, 64, 3)])
def test_scalar_reference_kernel_agrees_too(kind, n, h, w, cin, cout, built_lib):
    """the CUDA-core kernel kept as an on-device cross-check (w2x_selftest_conv) sees the same numbers"""
    import w2x
    d = w2x.selftest_conv(kind, n, h, w, cin, cout, seed=7)
    assert 0 <= d < 4e-3, f"max |diff| = {d}"
