"""-m gpu: SwinUNet (swin_unet/*) through the C ABI vs the fp32 PyTorch oracle (torchvision SwinTransformerBlock).
Same tolerance as the CUNet family: u8 within +-1 LSB on >= 99.9% of pixels, PSNR >= 50 dB."""
import numpy as np
import pytest
import torch

from oracle import tiling

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def swin_models(tmp_path_factory):
    import __graft_entry__
    d = str(tmp_path_factory.mktemp("swin_models"))
    return {s: __graft_entry__.make_synthetic_model(d, scale=s, noise=3, model="swin_unet/art") for s in (1, 2, 4)}


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def _engine(swin_models, scale, tile, batch, blend=1 / 16):
    import w2x
    model_t, path = swin_models[scale]
    e = w2x.Img2Img()
    msgs = []
    e.setMessageCallback(lambda s, m: msgs.append((s, m)))
    assert e.build(path, w2x.BuildConfig.fixed(batch, tile)), msgs
    assert e.load(path, w2x.RenderConfig(batchSize=batch, height=tile, width=tile, scaling=scale, overlap=(blend, blend))), msgs
    return e, model_t, msgs


@pytest.mark.parametrize("scale", [1, 2, 4])
def test_swin_infer_matches_fp32_oracle(scale, built_lib, swin_models):
    e, model_t, msgs = _engine(swin_models, scale, 64, 2)
    assert e.output_tile_size == 48 * scale
    x = np.random.default_rng(0).random((2, 3, 64, 64), dtype=np.float32)
    x = (np.rint(x * 255) / 255).astype(np.float32)
    y = e.infer(x)
    assert y is not None, msgs
    with torch.no_grad():
        ref = model_t(torch.from_numpy(x)).numpy()
    assert y.shape == ref.shape
    err = np.abs(y - ref)
    assert (err * 255 <= 1.0).mean() >= 0.999, ((err * 255 <= 1.0).mean(), err.max())
    assert _psnr(y * 255, ref * 255) > 50
    e.close()


def test_swin_tile_112_two_window_rows(built_lib, swin_models):
    """T = 112: 96 tokens -> 16 / 8 / 4 windows per side at the three levels (the shift mask matters on every level)."""
    e, model_t, msgs = _engine(swin_models, 2, 112, 1)
    x = np.random.default_rng(1).random((1, 3, 112, 112), dtype=np.float32)
    x = (np.rint(x * 255) / 255).astype(np.float32)
    y = e.infer(x)
    assert y is not None, msgs
    with torch.no_grad():
        ref = model_t(torch.from_numpy(x)).numpy()
    err = np.abs(y - ref)
    assert (err * 255 <= 1.0).mean() >= 0.999 and _psnr(y * 255, ref * 255) > 50, ((err * 255 <= 1.0).mean(), err.max())
    e.close()


def test_swin_infer_tile256_scale4(built_lib, swin_models):
    """BASELINE configs[3] tile: swin_unet/art scale 4, tileSize 256 -> 960 (40 / 20 / 10 windows per side)."""
    e, model_t, msgs = _engine(swin_models, 4, 256, 1)
    assert e.output_tile_size == 960
    x = np.random.default_rng(2).random((1, 3, 256, 256), dtype=np.float32)
    x = (np.rint(x * 255) / 255).astype(np.float32)
    y = e.infer(x)
    assert y is not None, msgs
    with torch.no_grad():
        ref = model_t(torch.from_numpy(x)).numpy()
    err = np.abs(y - ref)
    assert (err * 255 <= 1.0).mean() >= 0.999 and _psnr(y * 255, ref * 255) > 50, ((err * 255 <= 1.0).mean(), err.max())
    e.close()


def test_swin_render_matches_oracle(built_lib, swin_models):
    """cfg4 in miniature: swin_unet/art scale 4, batch 4 with padding slots, blend 1/16."""
    e, model_t, msgs = _engine(swin_models, 4, 64, 4)
    src = tiling.synthetic_frame(100, 70, 9)
    dst = e.render(src)
    assert dst is not None, msgs

    def f(x):
        with torch.no_grad():
            return model_t(torch.from_numpy(np.ascontiguousarray(x))).numpy()

    ref = tiling.render(src, f, 64, e.output_tile_size, 4, 1 / 16, 4)
    assert dst.shape == ref.shape == (280, 400, 3)
    diff = np.abs(dst.astype(np.int32) - ref.astype(np.int32))
    assert (diff <= 1).mean() >= 0.999 and _psnr(dst, ref) >= 50, ((diff <= 1).mean(), diff.max())
    e.close()


def test_swin_rejects_bad_tile(built_lib, swin_models):
    """(T - 16) % 48 != 0 (e.g. the CLI's 128) must fail loudly, not crash."""
    import w2x
    _, path = swin_models[2]
    e = w2x.Img2Img()
    msgs = []
    e.setMessageCallback(lambda s, m: msgs.append(m))
    assert e.build(path, w2x.BuildConfig.fixed(1, 128))
    assert not e.load(path, w2x.RenderConfig(batchSize=1, height=128, width=128, scaling=2))
    assert "not supported by swin_unet" in msgs[-1]
    e.close()


def test_swin_photo_stream_cfg5_order_and_parity(built_lib, tmp_path):
    """BASELINE configs[4] in miniature: swin_unet/photo scale 4, tileSize 256, an ordered stream of 8 distinct frames through
    the pipelined submit/wait path (pinned buffers, three frames in flight).  Every output must be the render of ITS frame
    (order check) and match the fp32 oracle of that frame."""
    import __graft_entry__
    import w2x
    model_t, path = __graft_entry__.make_synthetic_model(str(tmp_path), scale=4, noise=3, model="swin_unet/photo")
    assert "swin_unet" in path and "photo" in path
    e = w2x.Img2Img()
    msgs = []
    e.setMessageCallback(lambda s, m: msgs.append((s, m)))
    assert e.build(path, w2x.BuildConfig.fixed(4, 256)), msgs
    assert e.load(path, w2x.RenderConfig(batchSize=4, height=256, width=256, scaling=4)), msgs
    W, H, n = 250, 200, 8   # 2 x 1 tiles of 256 per frame
    frames = [tiling.synthetic_frame(W, H, 100 + s) for s in range(n)]
    pin_in = [w2x.PinnedArray((H, W, 3)) for _ in range(n)]
    pin_out = [w2x.PinnedArray((H * 4, W * 4, 3)) for _ in range(n)]
    tickets = []
    for f, pi, po in zip(frames, pin_in, pin_out):
        pi.array[...] = f
        t = e.submit(pi.ptr, W, H, po.ptr)
        assert t >= 0, msgs
        tickets.append(t)
    assert tickets == sorted(tickets)
    outs = []
    for t, po in zip(tickets, pin_out):
        assert e.wait(t)
        outs.append(po.array.copy())
    for i in range(n):
        assert np.array_equal(outs[i], e.render(frames[i])), f"frame {i} came back out of order or differs from the synchronous render"
    for i in range(n - 1):
        assert not np.array_equal(outs[i], outs[i + 1])

    def f(x):
        with torch.no_grad():
            return model_t(torch.from_numpy(np.ascontiguousarray(x))).numpy()

    for i in (0, n - 1):
        ref = tiling.render(frames[i], f, 256, 960, 4, 1 / 16, 4)
        diff = np.abs(outs[i].astype(np.int32) - ref.astype(np.int32))
        assert (diff <= 1).mean() >= 0.999 and _psnr(outs[i], ref) >= 50, (i, (diff <= 1).mean(), diff.max())
    e.close()
    for p in pin_in + pin_out:
        p.free()
