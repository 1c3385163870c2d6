"""Host feeding: nvJPEG batch decode into the device batch and the generic overlapped
host-batch streamer.  JPEG decoders differ in their IDCT and colour-conversion rounding (measured here: nvJPEG vs
libjpeg-turbo max 4, mean 0.45 on 4:4:4 streams), so the decode is checked against OpenCV's decode of the same streams
within that band and for being as close to the ORIGINAL pixels as OpenCV's decode is -- not bit for bit."""
import numpy as np
import pytest
import torch

from stainlib_b200.synth import synth_batch, synth_tile

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def sb(lib_built):
    import stainlib_b200
    return stainlib_b200


def _encode(tile, quality=95):
    ok, buf = cv2.imencode(".jpg", cv2.cvtColor(tile, cv2.COLOR_RGB2BGR),
                           [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])
    assert ok
    return buf.tobytes()


def test_decode_jpeg_batch_vs_opencv(sb):
    from stainlib_b200.io import decode_jpeg_batch
    tiles = synth_batch(700, 6, 256, 320)
    jpegs = [_encode(t) for t in tiles]
    out = decode_jpeg_batch(jpegs, 256, 320)
    assert out.is_cuda and out.shape == (6, 256, 320, 3)
    got = out.cpu().numpy().astype(np.int32)
    for i, j in enumerate(jpegs):
        ref = cv2.cvtColor(cv2.imdecode(np.frombuffer(j, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB).astype(np.int32)
        d = np.abs(got[i] - ref)
        assert d.max() <= 5 and d.mean() < 0.6, (i, d.max(), d.mean())
        orig = tiles[i].astype(np.int32)
        assert np.abs(got[i] - orig).mean() <= np.abs(ref - orig).mean() + 0.1
    # decoded tiles feed the normaliser like any other device batch
    n = sb.ExtractiveStainNormalizer("macenko")
    n.fit(synth_tile(1, 256, kind="target"))
    assert n.transform(out).shape == out.shape
    with pytest.raises(Exception):
        decode_jpeg_batch(jpegs, 128, 128)                 # wrong tile size
    with pytest.raises(Exception):
        decode_jpeg_batch([b"not a jpeg"], 256, 320)


def test_stream_host_batches_equals_device_path(sb):
    from stainlib_b200.augmentation.augmenter import HedLightColorAugmenter
    from stainlib_b200.io import stream_host_batches
    B, H, W = 37, 128, 160                                  # not a multiple of the chunk size
    host = torch.from_numpy(synth_batch(800, B, H, W)).pin_memory()
    rein = sb.ReinhardStainNormalizer()
    rein.fit(synth_tile(2, H, W, kind="target"))
    rng = np.random.default_rng(0)
    sig, bia = rng.uniform(-0.1, 0.1, (B, 3)), rng.uniform(-0.1, 0.1, (B, 3))
    hed = HedLightColorAugmenter()
    want = rein.transform(hed.transform(host.cuda(), sigmas=sig, biases=bia)).cpu()
    state = {"t0": 0}

    def op(x):                                              # per-tile parameters follow the chunk
        t0 = state["t0"]
        state["t0"] += x.shape[0]
        return rein.transform(hed.transform(x, sigmas=sig[t0:t0 + x.shape[0]], biases=bia[t0:t0 + x.shape[0]]))
    got = stream_host_batches(op, host, chunk_tiles=8)
    assert torch.equal(got, want)


def test_stream_jpeg_batches_equals_chunkwise_decode(sb):
    """The overlapped decode -> transform -> copy-back pipeline (decodes in flight back to back on one stream, no host
    synchronisation between chunks) gives the bytes of decoding everything first and transforming it in one call."""
    from stainlib_b200.io import decode_jpeg_batch, stream_jpeg_batches
    tiles = synth_batch(710, 37, 128, 160)
    jpegs = [_encode(t) for t in tiles]
    n = sb.ExtractiveStainNormalizer("macenko")
    n.fit(synth_tile(1, 128, kind="target"))
    dec = decode_jpeg_batch(jpegs, 128, 160)
    want = n.transform(dec)
    for chunk in (5, 16, 64):
        got = stream_jpeg_batches(n.transform, jpegs, 128, 160, chunk_tiles=chunk)
        assert torch.equal(got, want.cpu())
    kept = stream_jpeg_batches(n.transform, jpegs, 128, 160, chunk_tiles=8, keep_on_device=True)
    assert kept.is_cuda and torch.equal(kept, want)
