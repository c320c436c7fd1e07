"""Fused SwinUNet attention kernel (LayerNorm + QKV + W-MSA / SW-MSA [+ proj + residual at c = 96], kernels/swin_attn_sm100.cu) against
torchvision's own fp32 `shifted_window_attention` (the function SwinTransformerBlock.attn calls: x = x + attn(norm1(x))), through the C ABI."""
import numpy as np
import pytest
import torch

import w2x

pytestmark = pytest.mark.gpu


def relative_position_bias(table, window=6):
    """[heads][36][36] from a [(2w-1)^2][heads] table, indexed as torchvision's ShiftedWindowAttention does."""
    coords = torch.stack(torch.meshgrid(torch.arange(window), torch.arange(window), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += window - 1
    rel[:, :, 1] += window - 1
    rel[:, :, 0] *= 2 * window - 1
    index = rel.sum(-1).flatten()
    n = window * window
    return table[index].view(n, n, -1).permute(2, 0, 1).contiguous()


def reference(x16, gamma, beta, eps, wqkv16, bqkv, wproj16, bproj, relpos, heads, shift):
    """c = 96: x + attn(norm1(x)); c = 192: the attention output before the projection (what the kernel hands to the Linear kernel)."""
    from torchvision.models.swin_transformer import shifted_window_attention
    x = torch.from_numpy(x16.astype(np.float32))
    c = x.shape[-1]
    ln = torch.nn.functional.layer_norm(x, (c,), torch.from_numpy(gamma), torch.from_numpy(beta), eps)
    ln = ln.half().float()  # the kernel feeds fp16 rows to the tensor cores
    if c == 192:
        return shifted_window_attention(ln, torch.from_numpy(wqkv16.astype(np.float32)), torch.eye(c), torch.from_numpy(relpos).unsqueeze(0), [6, 6], heads,
                                        [shift, shift], qkv_bias=torch.from_numpy(bqkv), proj_bias=None).numpy()
    att = shifted_window_attention(ln, torch.from_numpy(wqkv16.astype(np.float32)), torch.from_numpy(wproj16.astype(np.float32)),
                                   torch.from_numpy(relpos).unsqueeze(0), [6, 6], heads, [shift, shift],
                                   qkv_bias=torch.from_numpy(bqkv), proj_bias=torch.from_numpy(bproj))
    return (x + att).numpy()


def make_case(n, h, w, seed, c=96, heads=6):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((n, h, w, c)) * 1.5 + rng.standard_normal((n, h, w, 1))).astype(np.float16)
    gamma = (1.0 + 0.2 * rng.standard_normal(c)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(c)).astype(np.float32)
    wqkv = (rng.standard_normal((3 * c, c)) * 1.5 / np.sqrt(c)).astype(np.float16)   # scores with a spread of a few units: a peaked softmax
    bqkv = (0.1 * rng.standard_normal(3 * c)).astype(np.float32)
    wproj = (rng.standard_normal((c, c)) / np.sqrt(c)).astype(np.float16)
    bproj = (0.1 * rng.standard_normal(c)).astype(np.float32)
    table = torch.from_numpy((0.5 * rng.standard_normal((121, heads))).astype(np.float32))
    relpos = relative_position_bias(table).numpy().astype(np.float32)
    return x, gamma, beta, 1e-5, wqkv, bqkv, wproj, bproj, relpos


# geometries: one window; three windows = one full tile; a ragged last tile; several tiles per CTA with images whose windows straddle tiles;
# the level-1 token grid of a batch of four 256-pixel tiles.  shift 3 exercises the roll and the region mask on the last window row / column.
@pytest.mark.parametrize("shift", [0, 3])
@pytest.mark.parametrize("c,n,h,w", [(96, 1, 6, 6), (96, 1, 6, 18), (96, 2, 12, 30), (96, 3, 60, 66), (96, 5, 120, 126), (96, 4, 240, 240),
                                     (192, 1, 6, 6), (192, 1, 18, 6), (192, 2, 12, 30), (192, 4, 60, 60), (192, 4, 120, 120)])
def test_fused_attention_matches_torchvision_fp32(c, n, h, w, shift):
    case = make_case(n, h, w, 7 * n + h + w + shift, c)
    out, _ = w2x.run_swin_attn(*case, heads=6, shift=shift)
    ref = reference(*case, 6, shift)
    err = np.abs(out.astype(np.float32) - ref)
    assert np.isfinite(out.astype(np.float32)).all()
    # fp16 output (half an ulp at |x| <= 8 is 0.004) and fp16 Q / K / V / P operands
    assert err.max() <= 0.03, f"max |diff| {err.max()}"
    assert err.mean() <= 2e-3, f"mean |diff| {err.mean()}"


@pytest.mark.parametrize("c", [96, 192])
def test_fused_attention_windows_are_independent(c):
    """A window's result must not depend on the tile it lands in or on its neighbours in the tile (row-band sharding relies on it)."""
    case = make_case(3, 12, 18, 11, c)
    full, _ = w2x.run_swin_attn(*case, shift=0)
    one, _ = w2x.run_swin_attn(case[0][1:2], *case[1:], shift=0)
    assert np.array_equal(full[1:2].view(np.uint16), one.view(np.uint16))
