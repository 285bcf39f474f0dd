"""Synthetic HEVC (H.265)-in-MP4 clip writer (test / bench tooling, not on the product path): the HEVC sibling of
tools/synth_h264.make_clip, so that the hvc1 / hvcC side of the demuxer and NVDEC's HEVC decoder can be exercised - BASELINE.json's
configs[3] names 4K HEVC and no HEVC encoder exists in this image.

A minimal, standards-conforming Main-profile encoder written from the published syntax (ITU-T H.265 7.3, 9.3):

  * every GOP starts with an IDR picture whose coding units are all PCM (`pcm_flag` = 1: raw samples, lossless - the decoded
    frame is known exactly); 16 x 16 coding tree blocks = minimum coding blocks, so the quadtree carries no split flags and a
    CTU is: part_mode (one context-coded bin), pcm_flag (terminate bin, arithmetic coder flushed), 384 raw bytes, coder
    re-initialised, end_of_slice_segment_flag;
  * the other pictures are P pictures of skipped coding units (cu_skip_flag = 1, MaxNumMergeCand = 1: zero motion from the
    previous picture), one short-term reference picture set in the SPS; their slice data is the same bytes for every picture;
  * CABAC is implemented as specified (9.3.4: EncodeDecision / EncodeTerminate / EncodeFlush, the 64 x 4 LPS range table, the
    state transition tables, context initialisation from the initValue tables at SliceQpY = 26); deblocking, SAO, tiles, WPP
    off; VPS / SPS / PPS carried in hvcC (sample entry hvc1), 4-byte NAL lengths.

Conformance is tested, not assumed (tests/test_synth_hevc_cpu.py): libavcodec's native HEVC decoder returns every picture
sample-exact.  Bit rate is that of raw video on the IDRs; stated wherever a number is measured on these clips.
"""

from __future__ import annotations

import struct

import numpy as np

from tools.synth_h264 import Bits, _escape_fast, box, mux_mp4, source_picture

CTB = 16

RANGE_TAB_LPS = [
    (128, 176, 208, 240), (128, 167, 197, 227), (128, 158, 187, 216), (123, 150, 178, 205), (116, 142, 169, 195), (111, 135, 160, 185),
    (105, 128, 152, 175), (100, 122, 144, 166), (95, 116, 137, 158), (90, 110, 130, 150), (85, 104, 123, 142), (81, 99, 117, 135),
    (77, 94, 111, 128), (73, 89, 105, 122), (69, 85, 100, 116), (66, 80, 95, 110), (62, 76, 90, 104), (59, 72, 86, 99), (56, 69, 81, 94),
    (53, 65, 77, 89), (51, 62, 73, 85), (48, 59, 69, 80), (46, 56, 66, 76), (43, 53, 63, 72), (41, 50, 59, 69), (39, 48, 56, 65),
    (37, 45, 54, 62), (35, 43, 51, 59), (33, 41, 48, 56), (32, 39, 46, 53), (30, 37, 43, 50), (29, 35, 41, 48), (27, 33, 39, 45),
    (26, 31, 37, 43), (24, 30, 35, 41), (23, 28, 33, 39), (22, 27, 32, 37), (21, 26, 30, 35), (20, 24, 29, 33), (19, 23, 27, 31),
    (18, 22, 26, 30), (17, 21, 25, 28), (16, 20, 23, 27), (15, 19, 22, 25), (14, 18, 21, 24), (14, 17, 20, 23), (13, 16, 19, 22),
    (12, 15, 18, 21), (12, 14, 17, 20), (11, 14, 16, 19), (11, 13, 15, 18), (10, 12, 15, 17), (10, 12, 14, 16), (9, 11, 13, 15),
    (9, 11, 12, 14), (8, 10, 12, 14), (8, 9, 11, 13), (7, 9, 11, 12), (7, 9, 10, 12), (7, 8, 10, 11), (6, 8, 9, 11), (6, 7, 9, 10),
    (6, 7, 8, 9), (2, 2, 2, 2),
]  # fmt: skip
TRANS_LPS = [0, 0, 1, 2, 2, 4, 4, 5, 6, 7, 8, 9, 9, 11, 11, 12, 13, 13, 15, 15, 16, 16, 18, 18, 19, 19, 21, 21, 22, 22, 23, 24, 24, 25, 26, 26, 27,
             27, 28, 29, 29, 30, 30, 30, 31, 32, 32, 33, 33, 33, 34, 34, 35, 35, 35, 36, 36, 36, 37, 37, 37, 38, 38, 63]  # fmt: skip


class Context:
    """One CABAC context variable, initialised from its initValue at SliceQpY (9.3.2.2)."""

    def __init__(self, init_value: int, qp: int = 26):
        m, n = (init_value >> 4) * 5 - 45, ((init_value & 15) << 3) - 16
        pre = min(126, max(1, ((m * min(51, max(0, qp))) >> 4) + n))
        self.mps = 0 if pre <= 63 else 1
        self.state = pre - 64 if self.mps else 63 - pre


class Cabac:
    """The arithmetic encoder of 9.3.4.2-9.3.4.5, emitting into a byte buffer; `pcm()` splices raw bytes the way pcm_sample does."""

    def __init__(self):
        self.out = bytearray()
        self.bits: list[int] = []
        self._init_engine()

    def _init_engine(self):
        self.low, self.range, self.first, self.outstanding = 0, 510, True, 0

    def _write(self, b: int):
        self.bits.append(b)

    def _put(self, b: int):
        if self.first:
            self.first = False
        else:
            self._write(b)
        while self.outstanding > 0:
            self._write(1 - b)
            self.outstanding -= 1

    def _renorm(self):
        while self.range < 256:
            if self.low < 256:
                self._put(0)
            elif self.low >= 512:
                self.low -= 512
                self._put(1)
            else:
                self.low -= 256
                self.outstanding += 1
            self.range <<= 1
            self.low <<= 1

    def decision(self, ctx: Context, b: int):
        lps = RANGE_TAB_LPS[ctx.state][(self.range >> 6) & 3]
        self.range -= lps
        if b != ctx.mps:
            self.low += self.range
            self.range = lps
            if ctx.state == 0:
                ctx.mps = 1 - ctx.mps
            ctx.state = TRANS_LPS[ctx.state]
        elif ctx.state < 62:
            ctx.state += 1
        self._renorm()

    def terminate(self, b: int):
        self.range -= 2
        if b:
            self.low += self.range
            self._flush()
        else:
            self._renorm()

    def _flush(self):
        self.range = 2
        self._renorm()
        self._put((self.low >> 9) & 1)
        self._write((self.low >> 8) & 1)
        self._write(1)  # ((low >> 7) & 3) | 1, two bits: the second one is the stop / alignment-one bit

    def _drain_aligned(self):
        while len(self.bits) % 8:
            self.bits.append(0)
        if self.bits:
            a = np.packbits(np.array(self.bits, dtype=np.uint8))
            self.out += a.tobytes()
            self.bits = []

    def pcm(self, raw: bytes):
        """after pcm_flag = 1 (terminate(1) was just called): pcm_alignment_zero_bits, the samples, engine re-initialised (9.3.2.5)"""
        self._drain_aligned()
        self.out += raw
        self._init_engine()

    def finish(self) -> bytes:
        """after end_of_slice_segment_flag = 1: the flush's last bit is the rbsp_stop_one_bit; pad to the byte boundary"""
        self._drain_aligned()
        return bytes(self.out)


def nal(nal_type: int, rbsp: bytes) -> bytes:
    return bytes([nal_type << 1, 1]) + _escape_fast(np.frombuffer(rbsp, dtype=np.uint8))


def _ptl(b: Bits, level_idc: int) -> None:
    b.u(2, 0).u(1, 0).u(5, 1)  # profile space, tier, Main
    b.u(32, 0x60000000)  # compatible with profiles 1 and 2
    b.u(1, 1).u(1, 0).u(1, 0).u(1, 1)  # progressive, interlaced, non-packed, frame-only
    b.u(44, 0)
    b.u(8, level_idc)


def parameter_sets(width: int, height: int, level_idc: int = 153) -> tuple[bytes, bytes, bytes]:
    w16, h16 = (width + CTB - 1) // CTB * CTB, (height + CTB - 1) // CTB * CTB
    v = Bits()
    v.u(4, 0).u(1, 1).u(1, 1).u(6, 0).u(3, 0).u(1, 1).u(16, 0xFFFF)
    _ptl(v, level_idc)
    v.u(1, 1).ue(1).ue(0).ue(0)  # sub-layer ordering info: dpb 2 pictures, no reordering
    v.u(6, 0).ue(0).u(1, 0).u(1, 0)  # max layer id, one layer set, no timing, no extension
    s = Bits()
    s.u(4, 0).u(3, 0).u(1, 1)
    _ptl(s, level_idc)
    s.ue(0).ue(1).ue(w16).ue(h16)  # sps id, 4:2:0, coded size
    if (w16, h16) != (width, height):
        s.u(1, 1).ue(0).ue((w16 - width) // 2).ue(0).ue((h16 - height) // 2)
    else:
        s.u(1, 0)
    s.ue(0).ue(0).ue(4)  # 8-bit luma / chroma, 8-bit POC lsb
    s.u(1, 1).ue(1).ue(0).ue(0)
    s.ue(1).ue(0)  # minimum coding block 16, coding tree block 16
    s.ue(0).ue(2)  # transform blocks 4..16
    s.ue(0).ue(0)  # transform hierarchy depths
    s.u(1, 0).u(1, 0).u(1, 0)  # scaling lists, AMP, SAO
    s.u(1, 1).u(4, 7).u(4, 7).ue(1).ue(0).u(1, 1)  # PCM: 8-bit samples, 16 x 16 only, loop filter off on PCM blocks
    s.ue(1)  # one short-term reference picture set:
    s.ue(1).ue(0).ue(0).u(1, 1)  # the previous picture, used by the current one
    s.u(1, 0).u(1, 0).u(1, 0)  # long-term refs, temporal MVP, strong intra smoothing
    s.u(1, 0).u(1, 0)  # VUI, extension
    p = Bits()
    p.ue(0).ue(0)
    p.u(1, 0).u(1, 0).u(3, 0).u(1, 0).u(1, 0)  # dependent slices, output flag, extra header bits, sign hiding, cabac_init_present
    p.ue(0).ue(0).se(0)  # default active refs, init_qp_minus26
    p.u(1, 0).u(1, 0).u(1, 0)  # constrained intra, transform skip, cu_qp_delta
    p.se(0).se(0)
    p.u(1, 0).u(1, 0).u(1, 0).u(1, 0).u(1, 0).u(1, 0)  # slice chroma offsets, weighted pred x2, transquant bypass, tiles, WPP
    p.u(1, 0)  # loop filter across slices
    p.u(1, 1).u(1, 0).u(1, 1)  # deblocking control present: no override, deblocking disabled
    p.u(1, 0).u(1, 0).ue(0).u(1, 0).u(1, 0)  # scaling list data, lists modification, parallel merge level, header extension, extension
    return nal(32, v.trailing()), nal(33, s.trailing()), nal(34, p.trailing())


def _slice_header(idr: bool, poc_lsb: int = 0) -> bytes:
    h = Bits()
    h.u(1, 1)  # first_slice_segment_in_pic_flag
    if idr:
        h.u(1, 0)  # no_output_of_prior_pics_flag
    h.ue(0)  # pps id
    h.ue(2 if idr else 1)  # slice_type: I / P
    if not idr:
        h.u(8, poc_lsb & 255).u(1, 1)  # slice_pic_order_cnt_lsb, short_term_ref_pic_set_sps_flag (set 0: no index bits)
        h.u(1, 0)  # num_ref_idx_active_override_flag
        h.ue(4)  # five_minus_max_num_merge_cand: one merge candidate, merge_idx is never coded
    h.se(0)  # slice_qp_delta
    return h.trailing()  # byte_alignment(): a one bit, then zeros


def idr_picture(y: np.ndarray, u: np.ndarray, v: np.ndarray) -> bytes:
    """y [H16, W16], u / v [H16/2, W16/2] uint8 -> one IDR_W_RADL NAL unit."""
    h, w = y.shape
    rows, cols = h // CTB, w // CTB
    yb = y.reshape(rows, CTB, cols, CTB).transpose(0, 2, 1, 3).reshape(rows * cols, 256)
    ub = u.reshape(rows, 8, cols, 8).transpose(0, 2, 1, 3).reshape(rows * cols, 64)
    vb = v.reshape(rows, 8, cols, 8).transpose(0, 2, 1, 3).reshape(rows * cols, 64)
    raw = np.concatenate([yb, ub, vb], axis=1)
    # The arithmetic coder is re-initialised after every PCM block, so the bytes in front of a CTU's samples depend only on the
    # part_mode context state (it saturates after ~60 CTUs) and on whether an end_of_slice_segment_flag = 0 precedes it: simulate
    # each distinct case once instead of 32 k times per 4K picture.
    part_mode = Context(184)  # part_mode, initType 0
    cache: dict[tuple[int, int, bool], tuple[bytes, int, int]] = {}
    n = rows * cols
    parts: list[bytes] = []
    for k in range(n):
        key = (part_mode.state, part_mode.mps, k == 0)
        hit = cache.get(key)
        if hit is None:
            c, ctx = Cabac(), Context(184)
            ctx.state, ctx.mps = part_mode.state, part_mode.mps
            if k > 0:
                c.terminate(0)  # the previous CTU's end_of_slice_segment_flag, coded right after the re-initialisation
            c.decision(ctx, 1)  # PART_2Nx2N
            c.terminate(1)  # pcm_flag
            c.pcm(b"")
            hit = cache[key] = (bytes(c.out), ctx.state, ctx.mps)
        parts.append(hit[0])
        part_mode.state, part_mode.mps = hit[1], hit[2]
        parts.append(raw[k].tobytes())
    tail = Cabac()
    tail.terminate(1)  # end_of_slice_segment_flag of the last CTU
    return nal(19, _slice_header(True) + b"".join(parts) + tail.finish())


def p_skip_slice_data(rows: int, cols: int) -> bytes:
    """Slice data of a P picture whose coding units are all skipped (the same bytes for every such picture of the sequence)."""
    c = Cabac()
    skip = [Context(197), Context(185), Context(201)]  # cu_skip_flag, initType 1
    for r in range(rows):
        for q in range(cols):
            c.decision(skip[(q > 0) + (r > 0)], 1)  # ctxInc = skipped-and-available left + above neighbours
            c.terminate(1 if (r, q) == (rows - 1, cols - 1) else 0)
    return c.finish()


def p_picture(slice_data: bytes, poc_lsb: int) -> bytes:
    return nal(1, _slice_header(False, poc_lsb) + slice_data)  # TRAIL_R


def hvc1_entry(width: int, height: int, vps: bytes, sps: bytes, pps: bytes, level_idc: int = 153) -> bytes:
    arrays = b"".join(bytes([0x80 | t]) + struct.pack(">HH", 1, len(x)) + x for t, x in ((32, vps), (33, sps), (34, pps)))
    hvcc = box(b"hvcC", bytes([1, 1]), struct.pack(">I", 0x60000000), bytes([0x90, 0, 0, 0, 0, 0]), bytes([level_idc]), struct.pack(">H", 0xF000),
               bytes([0xFC, 0xFD, 0xF8, 0xF8]), struct.pack(">H", 0), bytes([0x0F, 3]), arrays)  # fmt: skip
    return box(b"hvc1", b"\x00" * 6, struct.pack(">H", 1), b"\x00" * 16, struct.pack(">HH", width, height), struct.pack(">II", 0x00480000, 0x00480000),
               b"\x00" * 4, struct.pack(">H", 1), b"\x00" * 32, struct.pack(">Hh", 0x18, -1), hvcc)  # fmt: skip


def make_clip(width: int, height: int, fps: int, seconds: float, seed: int = 0, gop: int | None = None, return_sources: bool = False):
    """Returns mp4 bytes (and, optionally, {frame_index: (y, u, v)} of the IDR source pictures, cropped)."""
    assert width % 2 == 0 and height % 2 == 0
    gop = fps if gop is None else gop
    n_frames = int(round(fps * seconds))
    w16, h16 = (width + CTB - 1) // CTB * CTB, (height + CTB - 1) // CTB * CTB
    vps, sps, pps = parameter_sets(width, height)
    skip_data = p_skip_slice_data(h16 // CTB, w16 // CTB)
    samples, sync, sources = [], [], {}
    for i in range(n_frames):
        if i % gop == 0:
            y, u, v = source_picture(width, height, seed, i // gop)
            unit = idr_picture(y[:h16, :w16], u[: h16 // 2, : w16 // 2], v[: h16 // 2, : w16 // 2])
            if return_sources:
                sources[i] = (y[:height, :width].copy(), u[: height // 2, : width // 2].copy(), v[: height // 2, : width // 2].copy())
        else:
            unit = p_picture(skip_data, i % gop)
        samples.append(struct.pack(">I", len(unit)) + unit)
        sync.append(i % gop == 0)
    data = mux_mp4(samples, sync, b"", b"", width, height, fps, sample_entry=hvc1_entry(width, height, vps[0:], sps, pps))
    return (data, sources) if return_sources else data
