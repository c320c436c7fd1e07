"""-m gpu: the tcgen05 implicit-GEMM convolution vs the scalar CUDA-core reference kernel on the device, per layer
kind and at the real layer shapes (both accumulate in fp32; outputs are fp16, so agreement is to ~1 fp16 ulp)."""
import pytest

pytestmark = pytest.mark.gpu

# kind: 0 conv3x3, 1 conv2x2s2, 2 convT2x2s2(+skip), 3 convT4x4s2p3 head, 4 conv3x3 image head (+skip, clamp),
# 5 the same image head through the dedicated taps-in-N kernel (conv_head.cu), 6 the convT4x4 head through its taps-in-N kernel
CASES = [
    (0, 1, 20, 24, 64, 64),     # one M tile with edges (patch kernel)
    (0, 8, 100, 100, 64, 64),   # > 148 tiles: persistent loop, stage ring wrap, double-buffered staging
    (0, 1, 60, 52, 128, 64),    # patch kernel, two 64-channel chunks
    (0, 3, 45, 77, 64, 128),    # patch kernel, two output-channel slices
    (0, 2, 61, 45, 32, 64),     # cin 32 -> 64-byte swizzle path, odd sizes
    (0, 1, 126, 126, 64, 128),  # u1.conv2.0 at T=256
    (0, 1, 40, 40, 128, 256),   # N = 256
    (0, 1, 33, 37, 256, 128),   # K = 2304
    (1, 2, 60, 60, 64, 64),     # down 2x2 s2
    (1, 1, 42, 42, 128, 128),
    (2, 2, 26, 26, 64, 64),     # convT 2x2 + skip, N = 256
    (2, 1, 17, 17, 128, 128),   # N = 512 -> two N tiles
    (3, 2, 50, 50, 64, 3),      # convT 4x4 s2 p3 head
    (4, 2, 58, 58, 64, 3),      # image head + z1 crop + clamp
    (5, 2, 58, 58, 64, 3),      # image head kernel: partial tiles on both edges
    (5, 3, 26, 146, 64, 3),     # image head kernel: exact multiples of the 8 x 16 tile, three images
    (6, 2, 50, 50, 64, 3),      # convT 4x4 head kernel: partial tiles
    (6, 3, 33, 17, 64, 3),      # convT 4x4 head kernel: exactly 32 x 16 GEMM pixels, three images
]


@pytest.mark.parametrize("kind,n,h,w,cin,cout", CASES)
def test_igemm_matches_reference_kernel(kind, n, h, w, cin, cout, built_lib):
    import w2x
    d = w2x.selftest_conv(kind, n, h, w, cin, cout, seed=7)
    assert 0 <= d < 4e-3, f"max |diff| = {d}"
