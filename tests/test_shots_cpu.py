"""CPU: the shot-detection oracle against the reference's golden outputs, the product's host-side shot logic against
both, and the stage contract with a stub model (mirrors tests/cosmos_curate/pipelines/video/clipping/
test_transnetv2_extraction.py:26-107 of the reference)."""

from __future__ import annotations

import uuid

import numpy as np
import pytest

from conftest import load_golden

from cosmos_curate_b200 import shots
from cosmos_curate_b200.data_model import SplitPipeTask, Video
from cosmos_curate_b200.interfaces import SequentialRunner, run_pipeline
from oracle import transnetv2 as tn


@pytest.fixture(scope="module")
def golden():
    return load_golden("transnetv2_ref.npz")


def _cfg(row):
    mn, mx, mode, crop = (int(v) for v in row)
    return None if mn < 0 else mn, None if mx < 0 else mx, "stride" if mode else "truncate", None if crop < 0 else crop


def test_oracle_model_matches_reference_outputs(golden):
    sd = tn.random_state_dict(int(golden["seed"]))
    for k in ("full", "short95", "short45", "tiny7"):
        p = tn.forward(sd, golden[f"win_{k}"][None]).numpy()[0, :, 0]
        np.testing.assert_allclose(p, golden[f"prob_{k}"], rtol=0, atol=1e-6)


def test_oracle_window_stitching_matches_reference(golden):
    sd = tn.random_state_dict(int(golden["seed"]))
    thr = float(golden["pred_threshold"])
    for n in (120, 51, 10):  # 170 / 100 are covered on the GPU side; keep the CPU suite short
        p = tn.probabilities(sd, golden["video"][:n])
        np.testing.assert_allclose(p, golden[f"probs_{n}"], rtol=0, atol=1e-6)
        assert np.array_equal(tn.predictions(sd, golden["video"][:n], thr), golden[f"pred_{n}"])
    assert 0 < golden["pred_120"].sum() < 120  # the seeded weights straddle the threshold: the comparison is not vacuous


def test_window_plan_never_pads_the_tail():
    assert tn.window_plan(120) == [(0, 75, 25), (25, 95, 0), (75, 45, 0)]
    assert tn.window_plan(100) == [(0, 75, 25), (25, 75, 0)]
    assert tn.window_plan(10) == [(0, 10, 25)]
    for n in (1, 49, 50, 51, 99, 101, 1234):
        kept = sum(max(0, min(75, real + pad) - 25) for _, real, pad in tn.window_plan(n))
        assert kept >= n  # frames 25..74 of every window cover the video


def test_shot_logic_oracle_and_product_match_reference(golden):
    cfgs = [_cfg(r) for r in golden["filter_cfgs"]]
    for ti in range(int(golden["n_tracks"])):
        track = golden[f"track_{ti}"]
        for entire in (0, 1):
            want = golden[f"scenes_{ti}_{entire}"]
            a = tn.scenes_from_predictions(track, bool(entire))
            b = shots.scenes_from_predictions(track.reshape(-1, 1), entire_scene_as_clip=bool(entire))
            assert np.array_equal(a, want) and np.array_equal(b, want) and b.dtype == np.int32
            for ci, (mn, mx, mode, crop) in enumerate(cfgs):
                wf = golden[f"scenes_{ti}_{entire}_f{ci}"]
                assert np.array_equal(tn.filter_scenes(want, mn, mx, mode, crop), wf)
                assert np.array_equal(shots.filter_scenes(want, mn, mx, mode, crop), wf)


def test_product_shot_logic_equals_oracle_on_random_tracks():
    rng = np.random.default_rng(99)
    for _ in range(300):
        n = int(rng.integers(1, 400))
        track = (rng.random(n) < rng.choice([0.01, 0.1, 0.5, 0.9])).astype(np.uint8)
        for entire in (False, True):
            a = tn.scenes_from_predictions(track, entire)
            assert np.array_equal(a, shots.scenes_from_predictions(track, entire_scene_as_clip=entire))
            mn = rng.choice([None, 0, 5, 30])
            mx = rng.choice([None, 7, 20, 100])
            crop = rng.choice([None, 0, 2, 9])
            for mode in ("stride", "truncate"):
                assert np.array_equal(tn.filter_scenes(a, mn, mx, mode, crop), shots.filter_scenes(a, mn, mx, mode, crop)), (track, mn, mx, mode, crop)


def test_threshold_is_compared_in_float32():
    p = np.array([np.float32(0.4), np.nextafter(np.float32(0.4), np.float32(1))], dtype=np.float32)
    assert shots.predictions_from_probabilities(p, 0.4).ravel().tolist() == [0, 1]  # float32(0.4) > 0.4 in double, but not in float32
    import torch

    assert (torch.from_numpy(p) > 0.4).tolist() == [False, True]


def test_stage_lengths_and_clip_ids():
    assert shots.stage_lengths(29.97, 2.0, 48, 60.0, 0.5) == (60, 1799, 14)
    assert shots.stage_lengths(24.0, 2.0, 48, 60.0, 0.5) == (48, 1440, 12)
    assert shots.stage_lengths(10.0, None, 48, None, None) == (48, None, None)
    assert shots.stage_lengths(24.0, 2.0, None, 60.0, 0.0) == (48, 1440, None)
    sc = np.array([[12, 228]], dtype=np.int32)
    (uid, span), = shots.clips_from_scenes("s3://b/v.mp4", sc, 24.0)
    assert uid == uuid.uuid5(uuid.NAMESPACE_URL, "s3://b/v.mp4_12_228") and span == (0.5, 9.5)
    assert shots.clips_from_scenes("v", np.array([[0, 5], [5, 9], [9, 12]], np.int32), 1.0, limit_clips=2)[-1][1] == (5.0, 9.0)
    assert tn.clips_for_video("s3://b/v.mp4", sc, 24.0) == [(uid, span)]


def test_seeded_weights_agree_between_product_and_oracle():
    from cosmos_curate_b200.models.transnetv2 import seeded_state_dict

    a, b = seeded_state_dict(5), tn.random_state_dict(5)
    assert a.keys() == b.keys()
    assert all(np.array_equal(a[k], b[k]) for k in a)


# ---- stage contract with a stub model ---------------------------------------------------------------------------------
class _StubModel:
    conda_env_name = "unified"
    model_id_names = ["Sn4kehead/TransNetV2"]

    def __init__(self, prob):
        self.prob = np.asarray(prob, dtype=np.float32)

    def setup(self):
        self.ready = True

    def predict_video(self, frames):
        import torch

        assert tuple(frames.shape[1:]) == (27, 48, 3)
        return torch.from_numpy(self.prob[: len(frames)])


def _video_task(n=240, fps=24.0, with_frames=True, with_metadata=True):
    v = Video(input_video="clips/sintel.mp4")
    if with_metadata:
        m = v.metadata
        m.height, m.width, m.framerate, m.num_frames, m.duration, m.video_codec = 480, 854, fps, n, n / fps, "h264"
    if with_frames:
        v.frame_array = np.zeros((n, 27, 48, 3), dtype=np.uint8)
    return SplitPipeTask(video=v)


def _stage(prob, **kw):
    from cosmos_curate_b200.stages import TransNetV2ClipExtractionStage

    return TransNetV2ClipExtractionStage(model=_StubModel(prob), **kw)


def test_stage_requires_frame_extraction():
    out = run_pipeline([_video_task(with_frames=False)], [_stage(np.zeros(240))], runner=SequentialRunner())
    assert len(out) == 1 and len(out[0].video.clips) == 0


def test_stage_skips_incomplete_metadata():
    out = run_pipeline([_video_task(with_metadata=False)], [_stage(np.zeros(240))], runner=SequentialRunner())
    assert len(out[0].video.clips) == 0 and out[0].video.frame_array  # untouched


def test_stage_rejects_wrong_frame_shape():
    t = _video_task()
    t.video.frame_array = np.zeros((10, 28, 48, 3), dtype=np.uint8)
    with pytest.raises(ValueError, match="27x48x3"):
        run_pipeline([t], [_stage(np.zeros(240))], runner=SequentialRunner())


def test_stage_default_extraction_and_stats():
    prob = np.zeros(240, dtype=np.float32)
    prob[100:103] = 0.9  # one transition
    out = run_pipeline([_video_task()], [_stage(prob, log_stats=True)], runner=SequentialRunner())
    v = out[0].video
    # shots (0,100) and (103,239) -> crop 12 frames each side -> both >= 48 frames
    assert [c.span for c in v.clips] == [(12 / 24.0, 88 / 24.0), (115 / 24.0, 227 / 24.0)]
    assert v.clips[0].uuid == uuid.uuid5(uuid.NAMESPACE_URL, "clips/sintel.mp4_12_88") and v.clips[0].source_video == "clips/sintel.mp4"
    assert not v.frame_array  # dropped
    assert "TransNetV2ClipExtractionStage" in out[0].stage_perf
    for c in v.clips:
        assert 0.0 <= c.span[0] < c.span[1] <= v.metadata.duration


def test_stage_no_transitions():
    prob = np.full(240, 0.99, dtype=np.float32)  # threshold 1.0: nothing is a transition
    out = run_pipeline([_video_task()], [_stage(prob, threshold=1.0, entire_scene_as_clip=False)], runner=SequentialRunner())
    assert len(out[0].video.clips) == 0
    out = run_pipeline([_video_task()], [_stage(prob, threshold=1.0, entire_scene_as_clip=True, crop_s=0.0)], runner=SequentialRunner())
    (clip,) = out[0].video.clips
    assert clip.span == (0.0, 10.0)


def test_stage_limit_clips_and_bad_lengths():
    prob = np.zeros(240, dtype=np.float32)
    prob[60] = prob[120] = prob[180] = 0.9
    out = run_pipeline([_video_task()], [_stage(prob, limit_clips=1, min_length_s=1.0, min_length_frames=None)], runner=SequentialRunner())
    assert len(out[0].video.clips) == 1
    from cosmos_curate_b200.stages import TransNetV2ClipExtractionStage

    with pytest.raises(ValueError, match="Max length is smaller"):
        TransNetV2ClipExtractionStage(min_length_s=5.0, max_length_s=2.0, model=_StubModel([]))


def test_model_interface_surface():
    from cosmos_curate_b200.models import TransNetV2

    m = TransNetV2()
    assert m.conda_env_name == "unified" and m.model_id_names == ["Sn4kehead/TransNetV2"]
    with pytest.raises(FileNotFoundError):
        m.load()  # no checkpoint and synthetic weights not requested: never a silent substitute
    assert set(TransNetV2(seed=1).load()) == set(tn.random_state_dict(1))
