"""Host-side operand preparation of the fused SwinUNet kernels (no GPU): the regrouped / pre-scaled QKV operands of swin_attn_kernel and
the PatchUp + ToImage composition, each against a plain NumPy / PyTorch restatement of the layers they replace."""
import numpy as np
import torch
import torch.nn.functional as F

import w2x

LOG2E = 1.4426950408889634


def test_attention_operands_are_regrouped_and_prescaled(built_lib):
    rng = np.random.default_rng(5)
    for c in (96, 192):
        heads, d = 6, c // 6
        wqkv = (rng.standard_normal((3 * c, c)) / np.sqrt(c)).astype(np.float16)
        bqkv = rng.standard_normal(3 * c).astype(np.float32)
        relpos = rng.standard_normal((heads, 36, 36)).astype(np.float32)
        w, b, rel = w2x.swin_attn_prepare(wqkv, bqkv, relpos, heads)
        scale = np.float32(LOG2E / np.sqrt(d))
        for ch in range(c // 32):
            for t in range(3):   # q | k | v
                src = slice(t * c + 32 * ch, t * c + 32 * ch + 32)
                dst = slice(ch * 96 + t * 32, ch * 96 + t * 32 + 32)
                s = scale if t == 0 else np.float32(1)
                assert np.array_equal(w[dst], (wqkv[src].astype(np.float32) * s).astype(np.float16))
                assert np.allclose(b[dst], bqkv[src] * s, rtol=1e-6, atol=0)
        assert rel.shape[2] >= 36 and np.allclose(rel[:, :, :36], relpos * np.float32(LOG2E), rtol=1e-6, atol=0)
        # the softmax the kernel evaluates (exp2 of the pre-scaled scores) equals softmax(q k^T / sqrt(d) + bias)
        x = rng.standard_normal((36, c)).astype(np.float32)
        q = x @ wqkv[:d].astype(np.float32).T + bqkv[:d]
        k = x @ wqkv[c:c + d].astype(np.float32).T + bqkv[c:c + d]
        ref = torch.softmax(torch.from_numpy(q @ k.T / np.sqrt(d) + relpos[0]), -1).numpy()
        q2 = x @ w[:d].astype(np.float32).T + b[:d]          # head 0 = first d rows of chunk 0's q block
        k2 = x @ w[32:32 + d].astype(np.float32).T + b[32:32 + d]
        s2 = q2 @ k2.T + rel[0, :, :36]
        mine = np.exp2(s2 - s2.max(1, keepdims=True))
        mine /= mine.sum(1, keepdims=True)
        assert np.abs(mine - ref).max() < 5e-3   # fp16 rounding of the pre-scaled q rows


def test_patchup_toimage_composition_matches_the_two_layers(built_lib):
    rng = np.random.default_rng(9)
    c, k = 96, 96
    up = torch.nn.Linear(k, 4 * c)
    img = torch.nn.Linear(c, 12)
    with torch.no_grad():
        for m in (up, img):   # weights representable in fp16: the composition then only differs by its final rounding
            m.weight.copy_(m.weight.half().float())
            m.bias.copy_(m.bias.half().float())
    # packed layouts (model_pack.cpp packLinear): PatchUp row q*cout + c <- torch channel c*4 + q; ToImage row q*4 + c3 <- torch channel c3*4 + q
    wu = up.weight.detach().numpy().reshape(c, 4, k).transpose(1, 0, 2).reshape(4 * c, k)
    bu = up.bias.detach().numpy().reshape(c, 4).T.reshape(4 * c)
    wi = np.zeros((16, c), np.float32)
    bi = np.zeros(16, np.float32)
    wi.reshape(4, 4, c)[:, :3] = img.weight.detach().numpy().reshape(3, 4, c).transpose(1, 0, 2)
    bi.reshape(4, 4)[:, :3] = img.bias.detach().numpy().reshape(3, 4).T
    w, b = w2x.compose_up_to_image(wu, bu, wi, bi)
    x = torch.from_numpy(rng.standard_normal((1, 5, 7, k)).astype(np.float32))
    with torch.no_grad():
        y = F.pixel_shuffle(up(x).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)        # PatchUp: BHWC
        ref = F.pixel_shuffle(img(y).permute(0, 3, 1, 2), 2)[0].permute(1, 2, 0).numpy()   # ToImage: [4h][4w][3]
    out = x[0].numpy() @ w.astype(np.float32).T + b                                  # [h][w][64], row (oy*4 + ox)*4 + c3
    got = out.reshape(5, 7, 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(20, 28, 4)
    assert np.abs(got[..., 3]).max() == 0
    assert np.abs(got[..., :3] - ref).max() < 2e-2   # fp16 rounding of the composed weights (K = 96 terms of magnitude ~0.1)
    assert np.abs(got[..., :3] - ref).mean() < 2e-3
