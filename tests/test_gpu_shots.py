"""GPU: the fp32 shot-transition network (cb_transnet_*) against the oracle and the reference's golden outputs, the
window stitching, and the shot-detection stages end to end."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from gpu_helpers import ctx  # noqa: F401

pytestmark = pytest.mark.gpu

# fp32 everywhere; what differs from the CPU reference is summation order inside convolutions / dot products and
# expf vs the vectorised CPU exp.  Probabilities agree to ~1e-6; the bound leaves a margin.
PROB_ATOL = 2e-5


@pytest.fixture(scope="module")
def golden():
    return load_golden("transnetv2_ref.npz")


@pytest.fixture(scope="module")
def net(ctx, golden):  # noqa: F811
    from cosmos_curate_b200.runtime import ShotNet
    from oracle import transnetv2 as tn

    n = ShotNet(ctx, tn.random_state_dict(int(golden["seed"])), max_windows=4)
    yield n
    n.close()


def _decisive(prob, thr, margin=1e-4):
    return np.abs(prob - np.float32(thr)) > margin


@pytest.mark.parametrize("key", ["full", "short95", "short45", "tiny7"])
def test_forward_matches_reference_outputs(net, golden, key):
    w = torch.from_numpy(golden[f"win_{key}"]).cuda()[None]
    got = net.forward(w)
    assert got.shape == (1, w.shape[1], 1) and got.dtype == torch.float32
    np.testing.assert_allclose(got.cpu().numpy()[0, :, 0], golden[f"prob_{key}"], rtol=0, atol=PROB_ATOL)


def test_forward_batches_are_independent_and_chunked(net, golden):
    """7 windows through a workspace of 4: two library passes; every window equals its single-window result."""
    from oracle import transnetv2 as tn

    video = tn.synthetic_frames(400, seed=21, cuts=(90, 200, 333))
    wins = np.stack([video[i : i + 100] for i in range(0, 350, 50)])  # [7,100,27,48,3]
    got = net.forward(torch.from_numpy(wins).cuda()).cpu().numpy()[..., 0]
    sd = tn.random_state_dict(int(golden["seed"]))
    for i in (0, 3, 4, 6):
        want = tn.forward(sd, wins[i][None]).numpy()[0, :, 0]
        np.testing.assert_allclose(got[i], want, rtol=0, atol=PROB_ATOL)
    single = net.forward(torch.from_numpy(wins[5:6]).cuda()).cpu().numpy()[0, :, 0]
    assert np.array_equal(single, got[5])  # bit-identical regardless of batch position


@pytest.mark.parametrize("n", [170, 120, 100, 51, 10])
def test_predict_stitches_windows_like_the_reference(net, golden, n):
    from cosmos_curate_b200 import shots

    frames = torch.from_numpy(golden["video"][:n]).cuda()
    prob = net.predict(frames).cpu().numpy()
    want = golden[f"probs_{n}"]
    assert prob.shape == (n,)
    np.testing.assert_allclose(prob, want, rtol=0, atol=PROB_ATOL)
    thr = float(golden["pred_threshold"])
    pred = shots.predictions_from_probabilities(prob, thr)
    ok = _decisive(want, thr)
    assert ok.mean() > 0.95
    assert np.array_equal(pred.ravel()[ok], golden[f"pred_{n}"].ravel()[ok])


def test_predict_long_video_boundaries_equal_oracle(net, golden):
    """1,030 frames = 21 windows (19 full in batches of 4, two short tails): identical shot boundaries."""
    from cosmos_curate_b200 import shots
    from oracle import transnetv2 as tn

    video = tn.synthetic_frames(1030, seed=5, cuts=(100, 260, 275, 600, 601, 880))
    sd = tn.random_state_dict(int(golden["seed"]))
    want = tn.probabilities(sd, video)
    got = net.predict(torch.from_numpy(video).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=PROB_ATOL)
    for thr in (0.3, 0.4, 0.5, 0.6):
        if not _decisive(want, thr).all():
            continue
        a = shots.scenes_from_predictions(shots.predictions_from_probabilities(got, thr), entire_scene_as_clip=True)
        b = tn.scenes_from_predictions((torch.from_numpy(want) > thr).numpy(), True)
        assert np.array_equal(a, b)


def test_histogram_branch_is_exact(ctx, golden):  # noqa: F811
    """Colour-histogram similarities are integer counts / exact sqrt: zeroing every other input of fc1 isolates them."""
    from cosmos_curate_b200.runtime import ShotNet
    from oracle import transnetv2 as tn

    sd = tn.random_state_dict(9)
    w = sd["fc1.weight"].copy()
    w[:, 128:] = 0  # keep only the colour-histogram features (concat columns 0..127)
    sd["fc1.weight"] = w
    n = ShotNet(ctx, sd, max_windows=1)
    win = golden["win_full"]
    got = n.forward(torch.from_numpy(win).cuda()[None]).cpu().numpy()[0, :, 0]
    want = tn.forward(sd, win[None]).numpy()[0, :, 0]
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
    n.close()


def test_api_errors(ctx):  # noqa: F811
    import ctypes as C

    from cosmos_curate_b200 import _lib
    from cosmos_curate_b200._lib import CurateB200Error
    from cosmos_curate_b200.runtime import ShotNet
    from oracle import transnetv2 as tn

    lib = _lib.load()
    h = C.c_void_p()
    assert lib.cb_transnet_create(ctx.h, C.byref(h)) == 0
    buf = torch.zeros((1, 10, 27, 48, 3), dtype=torch.uint8, device="cuda")
    out = torch.zeros(10, device="cuda")
    assert lib.cb_transnet_forward(h, buf.data_ptr(), 1, 10, out.data_ptr(), None) != 0  # not finalized
    assert lib.cb_transnet_finalize(h, 1) != 0  # tensors missing
    assert b"never set" in lib.cb_last_error(ctx.h)
    a = np.zeros(5, dtype=np.float32)
    assert lib.cb_transnet_set_tensor(h, b"fc1.bias", a.ctypes.data, a.size) != 0  # wrong size
    assert lib.cb_transnet_set_tensor(h, b"cls_layer2.weight", a.ctypes.data, a.size) != 0  # unused head is not accepted
    lib.cb_transnet_destroy(h)
    sd = tn.random_state_dict(1)
    sd.pop("fc1.bias")
    with pytest.raises(CurateB200Error):
        ShotNet(ctx, sd)
    net = ShotNet(ctx, tn.random_state_dict(1), max_windows=1)
    with pytest.raises(CurateB200Error):
        net.forward(torch.zeros((1, 101, 27, 48, 3), dtype=torch.uint8, device="cuda"))
    net.close()


def _seeded_model(seed=3):
    from cosmos_curate_b200.models import TransNetV2

    return TransNetV2(seed=seed, max_windows=4)


def test_model_interface_call_signature(golden):
    from oracle import transnetv2 as tn

    m = _seeded_model(int(golden["seed"]))
    m.setup()
    x = torch.from_numpy(golden["win_short45"])[None]  # host tensor is accepted and moved
    y = m(x)
    assert y.is_cuda and y.shape == (1, 45, 1)
    np.testing.assert_allclose(y.cpu().numpy()[0, :, 0], golden["prob_short45"], rtol=0, atol=PROB_ATOL)
    with pytest.raises(AssertionError):
        m(torch.zeros((1, 10, 27, 48, 3)))  # not uint8
    assert tn.random_state_dict(3).keys() == m.load().keys()


def test_stages_end_to_end_on_the_fixture(golden):
    """Fixture mp4 -> (a) VideoFrameExtractionStage -> TransNetV2ClipExtractionStage and (b) the fused NVDEC stage:
    same clips, and equal to the oracle run on the very thumbnails the GPU decoded."""
    from cosmos_curate_b200 import shots
    from cosmos_curate_b200.data_model import SplitPipeTask, Video
    from cosmos_curate_b200.interfaces import SequentialRunner, run_pipeline
    from cosmos_curate_b200.stages import NvdecShotDetectionStage, TransNetV2ClipExtractionStage, VideoFrameExtractionStage
    from oracle import transnetv2 as tn

    data = (GOLDEN / "sintel_clip_10s.mp4").read_bytes()
    seed = int(golden["seed"])
    kw = dict(threshold=0.5, min_length_s=0.25, min_length_frames=4, crop_s=0.1, log_stats=True)

    def task():
        v = Video(input_video="fixtures/sintel_clip_10s.mp4", encoded_data=data)
        v.populate_metadata()
        return SplitPipeTask(video=v)

    grabbed = {}

    class _Spy(TransNetV2ClipExtractionStage):
        def _assign_clips(self, video, frames):
            grabbed["frames"] = np.array(frames)
            super()._assign_clips(video, frames)

    two = run_pipeline([task()], [VideoFrameExtractionStage(), _Spy(model=_seeded_model(seed), **kw)], runner=SequentialRunner())
    fused = run_pipeline([task()], [NvdecShotDetectionStage(model=_seeded_model(seed), **kw)], runner=SequentialRunner())
    v2, vf = two[0].video, fused[0].video
    assert v2.metadata.framerate == 24.0 and v2.metadata.num_frames == 240
    assert grabbed["frames"].shape == (240, 27, 48, 3)
    assert len(v2.clips) > 0
    assert [(c.uuid, c.span) for c in v2.clips] == [(c.uuid, c.span) for c in vf.clips]
    assert "_Spy" in two[0].stage_perf and "NvdecShotDetectionStage" in fused[0].stage_perf
    assert not v2.frame_array and not vf.frame_array
    # oracle on the same thumbnails
    sd = tn.random_state_dict(seed)
    want_p = tn.probabilities(sd, grabbed["frames"])
    if _decisive(want_p, 0.5).all():
        sc = tn.scenes_from_predictions((torch.from_numpy(want_p) > 0.5).numpy(), True)
        mn, mx, crop = tn.stage_lengths(24.0, 0.25, 4, 60.0, 0.1)
        want = tn.clips_for_video("fixtures/sintel_clip_10s.mp4", tn.filter_scenes(sc, mn, mx, "stride", crop), 24.0)
        assert [(c.uuid, c.span) for c in v2.clips] == want
    for c in v2.clips:
        assert 0.0 <= c.span[0] < c.span[1] <= v2.metadata.duration
