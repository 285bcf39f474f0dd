"""CPU tests of the host side: C-ABI export, product index math vs reference vectors, MP4 index."""

from __future__ import annotations

import json
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_json, load_golden

F = np.float32


# ---- C ABI -----------------------------------------------------------------------------------
def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    from cosmos_curate_b200 import _lib

    declared = _lib.header_symbols()
    assert len(declared) >= 25
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    lib = _lib.load()  # binds every signature; AttributeError if the .so lacks one
    assert lib.cb_abi_version() == 1
    nm = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in nm.splitlines() if " T " in line}
    assert set(declared) <= exported


def test_no_cpu_fallback_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cosmos_curate_b200 import _lib
    from cosmos_curate_b200.runtime import Context

    with pytest.raises(_lib.CurateB200Error):
        Context(0)
    import ctypes as C

    h = C.c_void_p()
    assert _lib.load().cb_init(0, C.byref(h)) == -1  # CB_ERR_CUDA, loudly
    assert b"no CPU fallback" in _lib.load().cb_last_error(None)


def test_product_never_imports_the_oracle():
    pkg = ROOT / "cosmos_curate_b200"
    for p in pkg.rglob("*"):
        if p.suffix in (".py", ".cu", ".cpp", ".h", ".cuh"):
            text = p.read_text()
            assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text, p


# ---- index math (product module) vs the reference ----------------------------------------------
def test_product_sampling_matches_reference_vectors():
    from cosmos_curate_b200 import sampling

    g = load_golden("sampling_ref.npz")
    n = 0
    for m in golden_json(g, "meta"):
        if "raises" in m:
            continue
        c = m["case"]
        ts = g[f"c{c}_ts"]
        ids, counts, _ = sampling.sample_closest(ts, m["rate"], start=ts[0], stop=ts[-1], endpoint=m["endpoint"], dedup=True)
        np.testing.assert_array_equal(ids, g[f"c{c}_ids"], err_msg=json.dumps(m))
        np.testing.assert_array_equal(counts, g[f"c{c}_counts"], err_msg=json.dumps(m))
        assert ids.dtype == np.int32 and counts.dtype == np.int32
        n += 1
    assert n >= 100
    for k in range(8):
        np.testing.assert_array_equal(sampling.find_closest_indices(g[f"f{k}_src"], g[f"f{k}_dst"]), g[f"f{k}_idx"])


def test_product_sampling_reference_kats():
    from cosmos_curate_b200 import sampling
    from test_oracle_cpu import FCI_KATS, SC_KATS

    for src, dst, want in FCI_KATS:
        assert sampling.find_closest_indices(np.array(src, dtype=F), np.array(dst, dtype=F)).tolist() == want
    for src, rate, start, stop, endpoint, ids, counts, dedup in SC_KATS:
        gi, gc, _ = sampling.sample_closest(np.array(src, dtype=F), rate, start, stop, endpoint, dedup)
        assert gi.tolist() == ids and gc.tolist() == counts
    with pytest.raises(ValueError):
        sampling.sample_closest(np.arange(3, dtype=F), -1.0)


def test_signature_and_plan():
    from cosmos_curate_b200 import sampling as s

    sig = s.FrameExtractionSignature(s.FrameExtractionPolicy.sequence, 1.0).to_str()
    assert sig == "FrameExtractionPolicy.sequence-1000"  # key format of clip.extracted_frames
    ts = np.arange(300, dtype=F) / F(30)
    plan = s.plan_extraction(ts, (s.FrameExtractionPolicy.sequence,), [1, 2])
    assert plan["FrameExtractionPolicy.sequence-2000"].tolist() == [*range(0, 300, 15), 299]
    np.testing.assert_array_equal(plan["FrameExtractionPolicy.sequence-1000"], plan["FrameExtractionPolicy.sequence-2000"][::2])
    plan = s.plan_extraction(ts, (s.FrameExtractionPolicy.sequence,), [1.5])
    assert len(plan["FrameExtractionPolicy.sequence-1500"]) == 16  # endpoint=True adds the last frame
    ids, counts = s.frame_ids(ts, s.FrameExtractionPolicy.middle, 1.0)
    assert ids.tolist() == [0] and counts.tolist() == [1]  # reference quirk
    with pytest.raises(ValueError):
        s.frame_ids(np.array([], dtype=F), s.FrameExtractionPolicy.sequence, 1.0)
    with pytest.raises(NotImplementedError):
        s.frame_ids(ts, s.FrameExtractionPolicy.first, 1.0)


# ---- MP4 index (host-only C code) ---------------------------------------------------------------
def test_mp4_index_of_reference_fixture():
    from cosmos_curate_b200 import sampling
    from cosmos_curate_b200.runtime import mp4_index

    data = np.fromfile(GOLDEN / "sintel_clip_10s.mp4", dtype=np.uint8)
    idx = mp4_index(data)
    assert (idx["codec"], idx["width"], idx["height"], idx["timescale"]) == (4, 854, 480, 12288)
    assert idx["n_samples"] == 240 and idx["n_sync"] == 1 and not idx["has_ctts"]
    np.testing.assert_array_equal(idx["pts"], np.arange(240) * 512)
    ts = sampling.timestamps_from_index(idx["pts"], idx["timescale"])
    assert ts.dtype == np.float32
    ids, _ = sampling.frame_ids(ts, sampling.FrameExtractionPolicy.sequence, 1.0)
    assert ids.tolist() == [0, 24, 48, 72, 96, 120, 144, 168, 192, 216, 239]
    cv2 = pytest.importorskip("cv2")
    cap = cv2.VideoCapture(str(GOLDEN / "sintel_clip_10s.mp4"))
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == idx["n_samples"]


def test_video_metadata_from_index_matches_container_facts():
    from cosmos_curate_b200 import sampling
    from cosmos_curate_b200.runtime import mp4_index
    from tools import synth_h264

    m = sampling.video_metadata_from_index(mp4_index(np.fromfile(GOLDEN / "sintel_clip_10s.mp4", dtype=np.uint8)))
    assert (m.width, m.height, m.fps, m.num_frames, m.video_codec, m.pixel_format, m.video_duration) == (854, 480, 24.0, 240, "h264", "yuv420p", 10.0)
    assert 1300 < m.bit_rate_k < 1500  # 1.76 MB over 10 s
    clip = synth_h264.make_clip(320, 192, 30, 1.0, seed=1)
    m = sampling.video_metadata_from_index(mp4_index(np.frombuffer(clip, dtype=np.uint8)))
    assert (m.width, m.height, m.fps, m.num_frames) == (320, 192, 30.0, 30)  # reference KATs: avg rate 30.0 (test_decoder_utils.py:358-385)
    cv2 = pytest.importorskip("cv2")
    cap = cv2.VideoCapture(str(GOLDEN / "sintel_clip_10s.mp4"))
    assert cap.get(cv2.CAP_PROP_FPS) == 24.0 and int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)) == 854


def test_mp4_index_rejects_garbage():
    from cosmos_curate_b200._lib import CurateB200Error
    from cosmos_curate_b200.runtime import mp4_index

    data = np.fromfile(GOLDEN / "sintel_clip_10s.mp4", dtype=np.uint8)
    for bad in (np.zeros(4, np.uint8), np.zeros(4096, np.uint8), data[: len(data) // 2], np.frombuffer(b"\x00\x00\x00\x18ftypisom" + b"\x00" * 64, np.uint8)):
        with pytest.raises(CurateB200Error) as e:
            mp4_index(bad)
        assert e.value.code == -5  # CB_ERR_DEMUX
    rng = np.random.default_rng(0)
    for _ in range(200):  # bit-flip fuzz of the moov box: must never crash
        d = data.copy()
        pos = rng.integers(len(d) - 20000, len(d), size=8)
        d[pos] = rng.integers(0, 256, size=8)
        try:
            mp4_index(d)
        except CurateB200Error:
            pass


def test_span_frame_ids_matches_a_standalone_clip():
    """A span of the source sampled in place == cutting that span into its own clip and sampling the clip."""
    from cosmos_curate_b200 import sampling

    ts = (np.arange(240) / 24).astype(np.float32)
    assert sampling.span_frame_ids(ts, (0.0, 10.0), 1.0).tolist() == [0, 24, 48, 72, 96, 120, 144, 168, 192, 216, 239]
    assert sampling.span_frame_ids(ts, (2.5, 7.5), 1.0).tolist() == [60, 84, 108, 132, 156, 179]
    for a, b, fps in ((12, 228, 1.0), (100, 103, 2.0), (0, 1, 1.0), (37, 200, 4.0)):
        span = (a / 24.0, b / 24.0)  # TransNetV2 spans are frame / fps
        ids = sampling.span_frame_ids(ts, span, fps)
        clip_ts = (ts[a:b] - ts[a]).astype(np.float32)
        want_ids, want_counts = sampling.frame_ids(clip_ts, sampling.FrameExtractionPolicy.sequence, fps)
        assert ids.tolist() == (a + np.repeat(want_ids, want_counts)).tolist()
    with pytest.raises(ValueError, match="selects no frame"):
        sampling.span_frame_ids(ts, (11.0, 12.0), 1.0)


def test_pynvc_target_size_rule():
    from cosmos_curate_b200 import sampling
    from oracle import preprocess

    for w, h in ((1920, 1080), (854, 480), (3840, 2160), (255, 144), (256, 144), (640, 360), (1280, 720), (511, 300), (770, 431)):
        assert sampling.pynvc_target_size(w, h) == preprocess.pynvc_target_size(w, h)
    assert sampling.pynvc_target_size(1920, 1080) == (274, 154)  # 1920 // 256 = 7: round(274.29), round(154.29)
    assert sampling.pynvc_target_size(854, 480) == (285, 160)    # factor 3: round(284.67), 160
    assert sampling.pynvc_target_size(200, 100) == (200, 100)
    assert sampling.pynvc_target_size(1920, 1080, 48, 27) == (48, 27)


def test_internvideo2_frame_selection_and_resampling_rule():
    """Host logic of InternVideo2FrameCreationStage: `frames[::len // fnum][:fnum]` (internvideo2_mm.py:399-400) against the
    oracle's restatement, the rate-doubling rule for short clips (internvideo2_stages.py:157-176) expressed on id lists,
    and the error keys that need no GPU."""
    import uuid

    from cosmos_curate_b200.data_model import Clip, SplitPipeTask, Video
    from cosmos_curate_b200.models.internvideo2_frames import select_frame_ids
    from cosmos_curate_b200.stages.internvideo2_frames import InternVideo2FrameCreationStage, sampled_ids_with_regen
    from oracle import sampling as O
    from oracle import video_tube as T

    for n in range(0, 70):
        assert select_frame_ids(n, 8) == T.select_frames(n, 8) == (list(range(n))[:: max(1, n // 8)][:8] if n >= 8 else [])
    ts = (np.arange(300) / 30.0).astype(np.float32)
    ids, used = sampled_ids_with_regen(ts, 2.0, 8)
    assert used == 2.0 and len(ids) == 21
    ts = (np.arange(45) / 30.0).astype(np.float32)  # 1.5 s: 2 fps -> 4 frames, 4 fps -> 7, 8 fps -> 13
    ids, used = sampled_ids_with_regen(ts, 2.0, 8)
    want_ids, want_counts, _ = O.sample_closest(ts, 8.0, start=ts[0], stop=ts[-1], endpoint=True, dedup=True)
    assert used == 8.0 and np.array_equal(ids, np.repeat(want_ids, want_counts)) and len(ids) >= 8
    ts = (np.arange(6) / 30.0).astype(np.float32)  # 0.2 s: still < 8 frames at 16 fps, 32 fps is beyond max_fps = 20
    ids, used = sampled_ids_with_regen(ts, 2.0, 8)
    assert used == 16.0 and len(ids) < 8
    stage = InternVideo2FrameCreationStage(target_fps=2.0)
    assert stage._frame_extraction_signature == "FrameExtractionPolicy.sequence-2000"
    with pytest.raises(ValueError, match="source"):
        InternVideo2FrameCreationStage(source="cpu")
    clips = [Clip(uuid=uuid.uuid4(), source_video="v.mp4", span=(0.0, 1.0)), Clip(uuid=uuid.uuid4(), source_video="v.mp4", span=(0.0, 1.0), encoded_data=b"xx")]
    task = SplitPipeTask(session_id="s", video=Video(input_video="v.mp4", clips=clips))
    assert stage.process_data([task]) == [task]  # neither clip reaches the GPU
    assert clips[0].errors == {"encoded_data": "empty"}
    assert clips[1].errors == {"frames-FrameExtractionPolicy.sequence-2000": "missing"}


def test_mp4_index_names_fragmented_files():
    """A fragmented MP4 (ffmpeg -movflags frag_keyframe / CMAF: empty moov sample tables + mvex, samples in moof boxes) is refused
    with a message that says so; a plain track without samples likewise - not reported as a zero-frame video."""
    import struct

    from cosmos_curate_b200._lib import CurateB200Error
    from cosmos_curate_b200.runtime import mp4_index

    def box(kind: bytes, body: bytes) -> bytes:
        return struct.pack(">I", 8 + len(body)) + kind + body

    def full(kind: bytes, body: bytes, version=0, flags=0) -> bytes:
        return box(kind, struct.pack(">I", (version << 24) | flags) + body)

    avcc = box(b"avcC", bytes([1, 66, 0, 30, 0xFF, 0xE1]) + struct.pack(">H", 4) + bytes([0x67, 66, 0, 30]) + bytes([1]) + struct.pack(">H", 2) + bytes([0x68, 0xCE]))
    entry = struct.pack(">I", 0)[:0] + bytes(6) + struct.pack(">H", 1) + bytes(16) + struct.pack(">HH", 320, 192) + struct.pack(">II", 0x480000, 0x480000) + bytes(4) + struct.pack(">H", 1) + bytes(32) + struct.pack(">Hh", 24, -1)
    avc1 = box(b"avc1", entry + avcc)
    stbl = box(b"stbl", full(b"stsd", struct.pack(">I", 1) + avc1) + full(b"stsz", struct.pack(">II", 0, 0)) + full(b"stco", struct.pack(">I", 0)) +
               full(b"stsc", struct.pack(">I", 0)) + full(b"stts", struct.pack(">I", 0)))
    mdia = box(b"mdia", full(b"mdhd", struct.pack(">IIIIHH", 0, 0, 30000, 0, 0x55C4, 0)) + full(b"hdlr", struct.pack(">I", 0) + b"vide" + bytes(12) + b"v\0") +
               box(b"minf", stbl))
    trak = box(b"trak", mdia)
    mvhd = full(b"mvhd", struct.pack(">IIII", 0, 0, 1000, 0) + bytes(80))
    ftyp = box(b"ftyp", b"isom" + struct.pack(">I", 512) + b"isomiso5")
    for with_mvex, text in ((True, "fragmented MP4"), (False, "no samples")):
        moov = box(b"moov", mvhd + trak + (box(b"mvex", full(b"trex", bytes(20))) if with_mvex else b""))
        data = np.frombuffer(ftyp + moov + box(b"moof", bytes(16)) + box(b"mdat", bytes(64)), dtype=np.uint8)
        with pytest.raises(CurateB200Error, match=text) as e:
            mp4_index(data)
        assert e.value.code == -5  # CB_ERR_DEMUX: the stages record video_decode_failed
