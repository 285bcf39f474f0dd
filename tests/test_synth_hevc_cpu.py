"""CPU: the synthetic HEVC clips (tools/synth_hevc.py) are conforming streams - libavcodec's native HEVC decoder returns every
picture sample-exact - and the product's demuxer handles the hvc1 / hvcC side (index, metadata, stream copy) the way it handles
avc1.  The NVDEC half (cb_decoder_decode on these clips) needs a GPU and is not part of this file."""

from __future__ import annotations

import os

import numpy as np
import pytest

from cosmos_curate_b200.data_model import Video
from cosmos_curate_b200.runtime import mp4_index
from cosmos_curate_b200.sampling import video_metadata_from_index
from cosmos_curate_b200.stages.clip_stream_copy import mp4_cut
from oracle import color
from tools import synth_hevc

os.environ.setdefault("OPENCV_LOG_LEVEL", "ERROR")
cv2 = pytest.importorskip("cv2")


def _decode(data, tmp_path, name, rgb: bool):
    p = tmp_path / name
    p.write_bytes(bytes(data))
    cap = cv2.VideoCapture(str(p))
    assert cap.isOpened()
    if not rgb:
        cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    out = []
    while True:
        ok, f = cap.read()
        if not ok:
            return out
        out.append(np.asarray(f).copy())


@pytest.mark.parametrize(("w", "h"), [(320, 192), (854, 480), (642, 362)])
def test_libavcodec_decodes_every_picture_sample_exact(tmp_path, w, h):
    """IDR pictures of PCM coding units come back exactly (luma, and chroma through the bit-exact swscale oracle); the skipped P
    pictures repeat them; sizes that are not multiples of the 16-pixel coding tree block use the conformance window."""
    gop = 4
    data, src = synth_hevc.make_clip(w, h, 30, 0.3, seed=5, gop=gop, return_sources=True)
    luma = _decode(data, tmp_path, "a.mp4", rgb=False)
    bgr = _decode(data, tmp_path, "a.mp4", rgb=True)
    assert len(luma) == len(bgr) == 9 and sorted(src) == [0, 4, 8]
    for i in range(9):
        y, u, v = src[(i // gop) * gop]
        assert np.array_equal(luma[i].reshape(-1)[: w * h].reshape(h, w), y), i
        nv12 = np.concatenate([y, np.stack([u, v], axis=-1).reshape(h // 2, w)], axis=0)
        assert np.array_equal(bgr[i][..., ::-1], color.nv12_to_rgb_swscale(nv12, h, w)), i  # chroma planes exact too
    assert not np.array_equal(src[0][0], src[4][0])  # the GOPs really differ


def test_index_metadata_and_stream_copy_of_hvc1(tmp_path):
    w, h, gop = 640, 368, 5
    data, src = synth_hevc.make_clip(w, h, 25, 0.6, seed=9, gop=gop, return_sources=True)
    idx = mp4_index(np.frombuffer(data, dtype=np.uint8))
    assert (idx["codec"], idx["width"], idx["height"], idx["n_samples"]) == (8, w, h, 15)  # cudaVideoCodec_HEVC
    assert idx["sync"].tolist() == [1, 0, 0, 0, 0] * 3 and np.all(np.diff(idx["pts"]) == idx["pts"][1] - idx["pts"][0])
    m = video_metadata_from_index(idx)
    assert m.video_codec == "hevc" and m.fps == 25.0 and m.num_frames == 15 and (m.width, m.height) == (w, h)
    v = Video(input_video="x.mp4", encoded_data=data)
    v.populate_metadata()
    v.populate_timestamps()
    assert v.metadata.video_codec == "hevc" and v.nvdec_support() and len(v.timestamps) == 15 and v.timestamps[5] == np.float32(0.2)
    cut = mp4_cut(np.frombuffer(data, dtype=np.uint8), 5, 7)  # second GOP + two pictures of the third
    ci = mp4_index(cut)
    assert ci["codec"] == 8 and ci["n_samples"] == 7 and ci["sync"].tolist() == [1, 0, 0, 0, 0, 1, 0]
    frames = _decode(cut, tmp_path, "cut.mp4", rgb=False)
    assert len(frames) == 7
    for i, f in enumerate(frames):
        assert np.array_equal(f.reshape(-1)[: w * h].reshape(h, w), src[5 if i < 5 else 10][0]), i


def test_cabac_pieces():
    """Context initialisation (9.3.2.2 at SliceQpY = 26) and the engine's bookkeeping on hand-checked cases."""
    for init, state, mps in ((184, 0, 1), (197, 15, 0), (185, 8, 1), (201, 16, 1), (154, 0, 1), (139, 0, 0)):
        c = synth_hevc.Context(init)
        assert (c.state, c.mps) == (state, mps), init
    c = synth_hevc.Cabac()
    c.terminate(1)
    assert c.finish() == bytes([0b11111110, 0b10000000])  # low = 508 flushed: ten bits with the first one dropped, then the stop bit and padding
    ctx = synth_hevc.Context(184)
    e = synth_hevc.Cabac()
    for _ in range(200):
        e.decision(ctx, 1)
    assert (ctx.state, ctx.mps) == (62, 1)  # saturates on the most probable symbol
    e.decision(ctx, 0)
    assert ctx.state == synth_hevc.TRANS_LPS[62] == 38
    assert len(synth_hevc.RANGE_TAB_LPS) == 64 == len(synth_hevc.TRANS_LPS) and all(a >= b for a, b in zip(synth_hevc.RANGE_TAB_LPS, synth_hevc.RANGE_TAB_LPS[1:63]))
    vps, sps, pps = synth_hevc.parameter_sets(1920, 1080)
    assert (vps[0] >> 1, sps[0] >> 1, pps[0] >> 1) == (32, 33, 34) and vps[1] == sps[1] == pps[1] == 1
    assert synth_hevc.p_skip_slice_data(68, 120) == synth_hevc.p_skip_slice_data(68, 120) and len(synth_hevc.p_skip_slice_data(68, 120)) < 200  # ~0.1 bit per skipped CTU
