"""-m gpu: whole-model and whole-render parity against the fp32 PyTorch/NumPy oracle through the C ABI.
Tolerance (BASELINE.json north_star): uint8 output within +-1 LSB on >= 99.9% of pixels at fp16, PSNR >= 50 dB vs fp32."""
import numpy as np
import pytest
import torch

from oracle import tiling

pytestmark = pytest.mark.gpu


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def _engine(models_dir, scale, tile, batch, blend=1 / 16, tta=False):
    import w2x
    _, per = models_dir
    model_t, path = per[scale]
    e = w2x.Img2Img()
    msgs = []
    e.setMessageCallback(lambda s, m: msgs.append((s, m)))
    assert e.build(path, w2x.BuildConfig.fixed(batch, tile)), msgs
    assert e.load(path, w2x.RenderConfig(batchSize=batch, height=tile, width=tile, scaling=scale, overlap=(blend, blend), tta=tta)), msgs
    return e, model_t, msgs


def _model_fn(model_t):
    def f(x):
        with torch.no_grad():
            return model_t(torch.from_numpy(np.ascontiguousarray(x))).numpy()
    return f


@pytest.mark.parametrize("scale,tile", [(2, 64), (1, 128), (2, 128)])
def test_infer_matches_fp32_oracle(scale, tile, built_lib, models_dir):
    e, model_t, msgs = _engine(models_dir, scale, tile, 2)
    x = np.random.default_rng(0).random((2, 3, tile, tile), dtype=np.float32)
    x = (np.rint(x * 255) / 255).astype(np.float32)
    y = e.infer(x)
    assert y is not None, msgs
    ref = _model_fn(model_t)(x)
    assert y.shape == ref.shape
    err = np.abs(y - ref)
    assert err.max() < 2.5e-3, err.max()          # < 0.64 LSB worst case
    assert _psnr(y * 255, ref * 255) > 55
    e.close()


def test_build_writes_reference_named_artifacts(built_lib, models_dir):
    import hashlib, json, os, ctypes as C
    import w2x
    e, _, _ = _engine(models_dir, 2, 64, 3)
    _, per = models_dir
    path = per[2][1]
    stem = os.path.splitext(path)[0]
    sidecars = [f for f in os.listdir(os.path.dirname(path)) if f.endswith(".json") and f.startswith(os.path.basename(stem) + "_")]
    assert sidecars
    found = False
    for s in sidecars:
        j = json.load(open(os.path.join(os.path.dirname(path), s)))
        assert list(j) == ["deviceName", "precision", "minBatchSize", "optBatchSize", "maxBatchSize", "minChannels", "optChannels",
                           "maxChannels", "minWidth", "optWidth", "maxWidth", "minHeight", "optHeight", "maxHeight"]
        if j["optBatchSize"] == 3:
            hs = f'{j["deviceName"].replace(" ", "")}.FP16.3.3.3.3.3.3.64.64.64.64.64.64'
            assert s == os.path.basename(stem) + "_" + hashlib.sha256(hs.encode()).hexdigest()[:16] + ".json"
            assert os.path.exists(os.path.join(os.path.dirname(path), s[:-5] + ".w2x"))
            found = True
    assert found
    # an incompatible render config must fail like the reference: "could not satisfy render configuration"
    msgs = []
    e.setMessageCallback(lambda s, m: msgs.append(m))
    assert not e.load(path, w2x.RenderConfig(batchSize=5, height=64, width=64, scaling=2))
    assert "could not satisfy render configuration" in msgs[-1]
    e.close()


@pytest.mark.parametrize("w,h,scale,tile,batch,blend", [
    (256, 256, 2, 64, 1, 1 / 16),     # cfg1 (BASELINE configs[0])
    (150, 97, 2, 64, 4, 1 / 8),       # padding slots in the last batch, ragged edges
    (200, 120, 1, 128, 2, 1 / 32),    # cunet 1x
    (90, 70, 2, 64, 3, 0.0),          # no blending
])
def test_render_matches_oracle(w, h, scale, tile, batch, blend, built_lib, models_dir):
    e, model_t, msgs = _engine(models_dir, scale, tile, batch, blend)
    src = tiling.synthetic_frame(w, h, 5)
    dst = e.render(src)
    assert dst is not None, msgs
    ref = tiling.render(src, _model_fn(model_t), tile, e.output_tile_size, scale, blend, batch)
    assert dst.shape == ref.shape == (h * scale, w * scale, 3)
    diff = np.abs(dst.astype(np.int32) - ref.astype(np.int32))
    assert (diff <= 1).mean() >= 0.999, ((diff <= 1).mean(), diff.max())
    assert _psnr(dst, ref) >= 50
    # idempotence / determinism: rendering the same frame again is byte-identical
    assert np.array_equal(e.render(src), dst)
    e.close()


def test_render_tile400_cunet1x(built_lib, models_dir):
    """BASELINE configs[2] tile size: cunet/art scale 1, tileSize 400 (out tile 344, iov = oov = 25), batch 2."""
    e, model_t, msgs = _engine(models_dir, 1, 400, 2, 1 / 16)
    assert e.output_tile_size == 344
    src = tiling.synthetic_frame(420, 400, 11)
    dst = e.render(src)
    assert dst is not None, msgs
    ref = tiling.render(src, _model_fn(model_t), 400, 344, 1, 1 / 16, 2)
    diff = np.abs(dst.astype(np.int32) - ref.astype(np.int32))
    assert (diff <= 1).mean() >= 0.999 and _psnr(dst, ref) >= 50, ((diff <= 1).mean(), diff.max())
    e.close()


def test_render_tta_matches_oracle_mean(built_lib, models_dir):
    """cfg3 shape in miniature: cunet 1x + 8-way TTA (mean)."""
    e, model_t, msgs = _engine(models_dir, 1, 128, 4, 1 / 16, tta=True)
    src = tiling.synthetic_frame(150, 100, 6)
    dst = e.render(src)
    assert dst is not None, msgs
    ref = tiling.render(src, _model_fn(model_t), 128, e.output_tile_size, 1, 1 / 16, 4, tta=True)
    diff = np.abs(dst.astype(np.int32) - ref.astype(np.int32))
    assert (diff <= 1).mean() >= 0.999 and _psnr(dst, ref) >= 50
    e.close()


def test_pipelined_submit_equals_sync_render(built_lib, models_dir):
    import w2x
    e, _, msgs = _engine(models_dir, 2, 64, 4)
    frames = [tiling.synthetic_frame(120, 80, s) for s in range(5)]
    sync = [e.render(f).copy() for f in frames]
    pin_in = [w2x.PinnedArray((80, 120, 3)) for _ in frames]
    pin_out = [w2x.PinnedArray((160, 240, 3)) for _ in frames]
    tickets = []
    for f, pi, po in zip(frames, pin_in, pin_out):
        pi.array[...] = f
        t = e.submit(pi.ptr, 120, 80, po.ptr)
        assert t >= 0, msgs
        tickets.append(t)
    for t, po, s in zip(tickets, pin_out, sync):
        assert e.wait(t)
        assert np.array_equal(po.array, s)
    e.close()
    for p in pin_in + pin_out:
        p.free()


@pytest.mark.parametrize("w,h", [(300, 260), (64, 200), (257, 65)])
def test_render_bandwise_download_strided_destination(built_lib, models_dir, w, h):
    """render() stitches and downloads the output band by band (rows that are final after a tile row) while later batches compute
    (engine.cpp: renderOnStream / render).  Several tile rows, a batch size that does not divide a tile row, a destination with a row
    stride larger than the row and guard bytes around it: the bytes inside must equal the single-stitch path (submit / wait), the guard
    bytes must stay untouched."""
    import w2x
    e, _, msgs = _engine(models_dir, 2, 64, 3)
    f = tiling.synthetic_frame(w, h, 17)
    ow, oh = 2 * w, 2 * h
    stride = ow * 3 + 40
    buf = np.full((oh + 2) * stride, 0xA5, np.uint8)
    dst = np.lib.stride_tricks.as_strided(buf[stride:], shape=(oh, ow, 3), strides=(stride, 3, 1))
    assert e.render_into(f, dst), msgs
    pin_in, pin_out = w2x.PinnedArray((h, w, 3)), w2x.PinnedArray((oh, ow, 3))
    pin_in.array[...] = f
    t = e.submit(pin_in.ptr, w, h, pin_out.ptr)
    assert t >= 0 and e.wait(t), msgs
    assert np.array_equal(dst, pin_out.array)
    guard = buf.reshape(oh + 2, stride)
    assert (guard[0] == 0xA5).all() and (guard[-1] == 0xA5).all() and (guard[1:-1, ow * 3:] == 0xA5).all()
    e.close()
    pin_in.free()
    pin_out.free()


def test_reload_and_changing_frame_sizes(built_lib, models_dir):
    """load may be called repeatedly (img2img_load.cpp:149-163,209-222 tears the engine down and re-allocates); frames of
    different sizes through one engine must not disturb each other."""
    import w2x
    e, model_t, msgs = _engine(models_dir, 2, 64, 2)
    a = tiling.synthetic_frame(90, 70, 1)
    b = tiling.synthetic_frame(140, 64, 2)
    ra, rb = e.render(a).copy(), e.render(b).copy()
    assert np.array_equal(e.render(a), ra) and np.array_equal(e.render(b), rb)
    _, per = models_dir
    path1 = per[1][1]
    assert e.build(path1, w2x.BuildConfig.fixed(3, 128)), msgs
    assert e.load(path1, w2x.RenderConfig(batchSize=3, height=128, width=128, scaling=1)), msgs   # different model, tile, batch
    assert e.output_tile_size == 72
    r1 = e.render(a)
    assert r1 is not None and r1.shape == a.shape
    path2 = per[2][1]
    assert e.load(path2, w2x.RenderConfig(batchSize=2, height=64, width=64, scaling=2)), msgs      # back to the first config
    assert np.array_equal(e.render(a), ra)
    e.close()


def test_flops_per_tile_matches_survey(built_lib, models_dir):
    """UpCUNet T=256 = 81.90 GFLOP / tile; T=64 = 2.39 GFLOP."""
    e, _, _ = _engine(models_dir, 2, 64, 1)
    assert abs(e.flops_per_tile / 1e9 - 2.39) < 0.01
    e.close()
    e, _, _ = _engine(models_dir, 2, 256, 1)
    assert abs(e.flops_per_tile / 1e9 - 81.90) < 0.05
    e.close()


def test_full_size_frame_properties(built_lib, models_dir):
    """BASELINE configs[1] at full size (1920x1080 -> 3840x2160, tile 256): too large for the CPU oracle, so the render is
    checked through size-independent properties: (a) the output does not depend on the batch size (tiles are independent and
    the SE sums are exact integers, so a tile's result is independent of its batch slot and of the padding slots),
    (b) repeated renders and the pipelined submit/wait path are byte-identical, (c) a frame that is constant per channel
    gives a 4-periodic output (every tile sees the same values; the stride-2 layers introduce the period), and (d) the top-left 256-pixel
    block equals a stand-alone render of the crop that covers the same first tile."""
    import w2x
    frame = tiling.synthetic_frame(1920, 1080, 3)
    e8, _, msgs = _engine(models_dir, 2, 256, 8)
    a = e8.render(frame)
    assert a is not None and a.shape == (2160, 3840, 3), msgs
    a = a.copy()
    assert np.array_equal(e8.render(frame), a)                                   # (b) deterministic
    pin_in, pin_out = w2x.PinnedArray((1080, 1920, 3)), w2x.PinnedArray((2160, 3840, 3))
    pin_in.array[...] = frame
    t = e8.submit(pin_in.ptr, 1920, 1080, pin_out.ptr)
    assert t >= 0 and e8.wait(t)
    assert np.array_equal(pin_out.array, a)                                      # (b) pipelined path
    flat = np.empty_like(frame)
    flat[...] = np.array([40, 128, 200], np.uint8)
    f = e8.render(flat)
    assert f is not None
    # (c) every tile sees the same flat input (replicate padding), so every tile produces the same pattern; the stride-2
    # (de)convolutions make it periodic with period 4 rather than constant, and tile origins are multiples of 408 = 4 * 102
    tiled = np.tile(f[:4, :4], (540, 960, 1)).astype(np.int32)
    assert np.abs(f.astype(np.int32) - tiled).max() <= 1                          # blend bands may round once
    e8.close()
    e5, _, msgs = _engine(models_dir, 2, 256, 5)
    b = e5.render(frame)
    assert b is not None, msgs
    assert np.array_equal(a, b)                                                  # (a) batch 8 == batch 5 (60 tiles: 4 vs 0 padding slots)
    # (d) the first tile covers input [0, 256) minus the model's context; pixels whose blend weights come from that tile alone
    # must equal a render of just that 256 x 256 crop (same tile, same padding on the top/left edges)
    crop = e5.render(np.ascontiguousarray(frame[:256, :256]))
    e5.close()
    own = 2 * 256 - 72 - 32                                                      # output pixels owned by tile 0 alone: out tile 440 minus the 32-pixel blend band
    assert np.array_equal(a[:own, :own], crop[:own, :own])
    pin_in.free(); pin_out.free()


def test_infer_cfg2_tile256_batch8_matches_fp32_oracle(built_lib, models_dir):
    """BASELINE configs[1] exactly at the model boundary: cunet/art scale 2, tileSize 256, batchSize 8 -- the only size that
    exercises nSplit = 4, the 256-wide layers and the odd extents 117 / 113.  Eight real image tiles through w2x_infer vs the
    fp32 PyTorch graph: u8 within +-1 LSB on >= 99.9 % of values, PSNR >= 50 dB (BASELINE.json north_star)."""
    e, model_t, msgs = _engine(models_dir, 2, 256, 8)
    assert e.output_tile_size == 440
    frame = tiling.synthetic_frame(1920, 1080, 7)[..., ::-1]
    g = tiling.calculate_tiles(1920, 1080, 3840, 2160, 256, 256, 440, 440, 2, 1 / 16, 1 / 16)
    picks = [0, 5, 6, 17, 29, 41, 54, 59]  # corners (replicate padding), edges and interior tiles of the 10 x 6 grid
    x = np.stack([tiling.normalize_u8(tiling.pad_roi(frame, g.in_rects[i])).transpose(2, 0, 1) for i in picks]).astype(np.float32)
    y = e.infer(x)
    assert y is not None, msgs
    ref = _model_fn(model_t)(x)
    assert y.shape == ref.shape == (8, 3, 440, 440)
    a, b = np.clip(np.rint(y * 255), 0, 255), np.clip(np.rint(ref * 255), 0, 255)
    frac = (np.abs(a - b) <= 1).mean()
    assert frac >= 0.999, (frac, np.abs(a - b).max())
    assert _psnr(a, b) >= 50
    for i in range(8):  # per tile as well: no slot of the batch may be worse than the bar
        assert (np.abs(a[i] - b[i]) <= 1).mean() >= 0.999, i
    e.close()


def test_render_cfg3_tile400_tta(built_lib, models_dir):
    """BASELINE configs[2] exactly: cunet/art scale 1 (denoise only), --tta, tileSize 400 (out tile 344, overlap 25 / 25).
    Two tiles x 8 augmentations; the product averages the eight de-augmented outputs."""
    e, model_t, msgs = _engine(models_dir, 1, 400, 4, 1 / 16, tta=True)
    assert e.output_tile_size == 344
    src = tiling.synthetic_frame(400, 330, 12)
    dst = e.render(src)
    assert dst is not None, msgs
    ref = tiling.render(src, _model_fn(model_t), 400, 344, 1, 1 / 16, 4, tta=True)
    diff = np.abs(dst.astype(np.int32) - ref.astype(np.int32))
    assert (diff <= 1).mean() >= 0.999 and _psnr(dst, ref) >= 50, ((diff <= 1).mean(), diff.max())
    e.close()


def test_pool_round_robin_keeps_frame_order(built_lib, models_dir):
    """w2x_pool_*: frames sharded round-robin over the pool's engines (here the same GPU listed twice, so it runs on a one-GPU box;
    tests/test_gpu_banded.py-style multi-GPU runs use distinct ids) and retired by ticket == frame number."""
    import w2x
    _, per = models_dir
    _, path = per[2]
    pool = w2x.Img2ImgPool([0, 0])
    msgs = []
    pool.setMessageCallback(lambda s, m: msgs.append((s, m)))
    assert pool.build(path, w2x.BuildConfig.fixed(4, 64)), msgs
    assert pool.load(path, w2x.RenderConfig(batchSize=4, height=64, width=64, scaling=2)), msgs
    single, _, _ = _engine(models_dir, 2, 64, 4)
    n, W, H = 11, 120, 80
    frames = [tiling.synthetic_frame(W, H, 50 + s) for s in range(n)]
    want = [single.render(f).copy() for f in frames]
    pin_in = [w2x.PinnedArray((H, W, 3)) for _ in range(n)]
    pin_out = [w2x.PinnedArray((2 * H, 2 * W, 3)) for _ in range(n)]
    tickets = []
    for f, pi, po in zip(frames, pin_in, pin_out):
        pi.array[...] = f
        tickets.append(pool.submit(pi.ptr, W, H, po.ptr))
    assert tickets == list(range(n)), (tickets, msgs)
    for t in reversed(tickets[:3]):      # any order is allowed ...
        assert pool.wait(t)
    for t in tickets[3:]:                # ... a writer asks in frame order
        assert pool.wait(t)
    for i in range(n):
        assert np.array_equal(pin_out[i].array, want[i]), i
    counts = pool.launch_counts()
    assert len(counts) == 2 and all(c > 0 for c in counts)     # both engines took part: frames 0,2,4,.. and 1,3,5,..
    assert not pool.wait(0)                                      # a ticket can be retired once
    pool.close()
    single.close()
    for p in pin_in + pin_out:
        p.free()
