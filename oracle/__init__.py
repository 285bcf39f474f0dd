"""CPU oracle for the decode -> sample -> preprocess -> embed/classify hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cosmos_curate_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker (or as the
timed CPU baseline), never as the product.

Each function restates one piece of the reference algorithm and cites the reference
``file:line`` (relative to the nvidia-cosmos/cosmos-curate checkout) it follows.

Pinning status (see DESIGN.md "Oracle"):
  * sampling.py   - pinned: every exact vector of the reference's own
                    tests/cosmos_curate/pipelines/video/utils/test_decoder_utils.py:40-201
                    plus vectors produced by importing the reference's decoder_utils here
                    (tests/golden/sampling_*.json, generator oracle/make_golden.py).
  * preprocess.py - pinned against the reference's own ``_CLIPImageEmbeddings.transforms``
                    (cosmos_curate/models/clip.py:48-62) executed on torch-CPU here
                    (tests/golden/clip_preprocess_*.npz).
  * color.py      - NV12->RGB follows OpenCV's COLOR_YUV2RGB_NV12 (the CPU stand-in for
                    cvcuda.cvtcolor_into, nvcodec_utils.py:178).  PARITY UNPINNED against
                    CV-CUDA itself: the reference has no test for nvcodec_utils and CV-CUDA is
                    not installable here; pinned bit-exactly against cv2 instead.
  * vit.py        - pinned against transformers' CLIPModel / SiglipVisionModel (the library the
                    reference calls, clip.py:41,71) with seeded random weights
                    (tests/golden/vit_*.npz).  The reference's aesthetic goldens
                    (4.8575 / 3.7989) need the real checkpoints, which are not available
                    offline: noted as unreproducible here.
"""
