"""-m gpu: one image sharded over GPUs by bands of tile rows with a peer-to-peer seam exchange must be
byte-identical to the single-GPU render (same global tile grid, same fp32 add order)."""
import numpy as np
import pytest
import torch

from oracle import tiling

pytestmark = pytest.mark.gpu


def _engine(models_dir, device, scale=2, tile=64, batch=4, tta=False, blend=1 / 16):
    import w2x
    _, per = models_dir
    _, path = per[scale]
    e = w2x.Img2Img()
    msgs = []
    e.setMessageCallback(lambda s, m: msgs.append((s, m)))
    assert e.build(path, w2x.BuildConfig.fixed(batch, tile, device=device)), msgs
    assert e.load(path, w2x.RenderConfig(deviceId=device, batchSize=batch, height=tile, width=tile, scaling=scale, tta=tta, overlap=(blend, blend))), msgs
    return e


@pytest.mark.parametrize("nbands,tta,blend,scale", [(2, False, 1 / 16, 2), (3, False, 1 / 8, 2), (2, True, 1 / 16, 1), (3, False, 0.0, 2), (5, False, 1 / 16, 2)])
def test_banded_on_one_gpu_is_byte_identical(nbands, tta, blend, scale, built_lib, models_dir):
    """The band logic (band rows + halo upload, per-band slot tables, seam strip exchange, TTA mean per band, partial stitch) with
    every band's engine on the SAME device, so it runs on a one-GPU box; the multi-GPU test below differs only in the device ids."""
    import w2x
    tile = 128 if scale == 1 else 64
    engines = [_engine(models_dir, 0, scale=scale, tile=tile, batch=3, tta=tta, blend=blend) for _ in range(nbands)]
    for (w, h) in [(150, 230), (70, 331)]:
        src = tiling.synthetic_frame(w, h, 7)
        ref = engines[0].render(src)
        got = w2x.render_banded(engines, src)
        assert got is not None, engines[0].last_error
        assert np.array_equal(got, ref), (w, h, int(np.abs(got.astype(int) - ref.astype(int)).max()))
    for e in engines:
        e.close()


def test_banded_single_engine_equals_render(built_lib, models_dir):
    import w2x
    e = _engine(models_dir, 0)
    src = tiling.synthetic_frame(150, 130, 3)
    ref = e.render(src)
    got = w2x.render_banded([e], src)
    assert got is not None and np.array_equal(got, ref)
    e.close()


@pytest.mark.parametrize("ngpu", [2, 4])
def test_banded_multi_gpu_is_byte_identical(ngpu, built_lib, models_dir):
    import w2x
    if torch.cuda.device_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    engines = [_engine(models_dir, d) for d in range(ngpu)]
    for (w, h) in [(150, 130), (97, 260)]:
        src = tiling.synthetic_frame(w, h, 4)
        ref = engines[0].render(src)
        got = w2x.render_banded(engines, src)
        assert got is not None, engines[0].last_error
        assert np.array_equal(got, ref), (w, h, int(np.abs(got.astype(int) - ref.astype(int)).max()))
    for e in engines:
        e.close()
