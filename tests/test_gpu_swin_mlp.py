"""Fused SwinUNet MLP kernel (LayerNorm + fc1 + GELU + fc2 + residual, kernels/swin_mlp_sm100.cu) against a plain PyTorch fp32
reference of the same op (torchvision SwinTransformerBlock: x = x + mlp(norm2(x)), nn.GELU() erf form), through the C ABI."""
import numpy as np
import pytest
import torch

import w2x

pytestmark = pytest.mark.gpu


def reference(x16, gamma, beta, eps, w1_16, b1, w2_16, b2):
    x = torch.from_numpy(x16.astype(np.float32))
    ln = torch.nn.functional.layer_norm(x, (x.shape[1],), torch.from_numpy(gamma), torch.from_numpy(beta), eps)
    ln = ln.half().float()  # the kernel feeds fp16 rows to the tensor cores
    h = torch.nn.functional.gelu(ln @ torch.from_numpy(w1_16.astype(np.float32)).T + torch.from_numpy(b1))
    h = h.half().float()    # hidden tile is stored as fp16 (as in the unfused path)
    return (x + h @ torch.from_numpy(w2_16.astype(np.float32)).T + torch.from_numpy(b2)).numpy()


def make_case(tokens, seed, c=96):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((tokens, c)) * 1.5 + rng.standard_normal((tokens, 1))).astype(np.float16)
    gamma = (1.0 + 0.2 * rng.standard_normal(c)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(c)).astype(np.float32)
    w1 = (rng.standard_normal((2 * c, c)) / np.sqrt(c)).astype(np.float16)
    b1 = (0.1 * rng.standard_normal(2 * c)).astype(np.float32)
    w2 = (rng.standard_normal((c, 2 * c)) / np.sqrt(2 * c)).astype(np.float16)
    b2 = (0.1 * rng.standard_normal(c)).astype(np.float32)
    return x, gamma, beta, 1e-5, w1, b1, w2, b2


# c / variant: 96 with resident weights (level-1 blocks), 96 and 192 with the weights streamed per hidden chunk (level-2 blocks).
# tokens: 1 (a single ragged tile), exactly one tile, one SM's worth + ragged tail, several tiles per CTA (persistent loop, both
# buffer parities), and the level-1 token count of a batch of four 256-pixel tiles
@pytest.mark.parametrize("c,variant", [(96, 0), (96, 1), (192, 0)])
@pytest.mark.parametrize("tokens", [1, 128, 129, 148 * 128 + 77, 5 * 148 * 128 + 1, 4 * 240 * 240])
def test_fused_mlp_matches_torch_fp32(tokens, c, variant):
    case = make_case(tokens, 100 + tokens % 97, c)
    out, _ = w2x.run_swin_mlp(*case, variant=variant)
    ref = reference(*case)
    err = np.abs(out.astype(np.float32) - ref)
    # fp16 output: half an ulp at |x| <= 8 is 0.004; accumulation order differences stay far below that
    assert np.isfinite(out.astype(np.float32)).all()
    assert err.max() <= 0.02, f"max |diff| {err.max()}"
    assert err.mean() <= 1.5e-3, f"mean |diff| {err.mean()}"


@pytest.mark.parametrize("c", [96, 192])
def test_fused_mlp_rows_are_independent(c):
    """A token's result must not depend on its tile neighbours or on the tile it lands in (row-band sharding relies on it)."""
    case = make_case(1000, 7, c)
    full, _ = w2x.run_swin_mlp(*case)
    part, _ = w2x.run_swin_mlp(case[0][300:517], *case[1:])
    assert np.array_equal(full[300:517].view(np.uint16), part.view(np.uint16))
