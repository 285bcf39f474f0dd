"""Synthetic H.264-in-MP4 clip writer (test / bench tooling, not on the product path).

There is no H.264 encoder in this image (no ffmpeg CLI, no PyAV, cv2.VideoWriter only opens mp4v) and
B200 has no NVENC, so BASELINE.json's "synthetic 1080p30 H.264 clips" are produced by this minimal,
standards-conforming encoder:

  * every GOP starts with an IDR picture whose macroblocks are all I_PCM (raw samples: lossless, so the
    decoded frame is known exactly - NVDEC and libavcodec output can be checked against the source);
  * the other pictures are P pictures made of P_L0_16x16 macroblocks with one global motion vector and no
    residual (5 bits per macroblock): the picture pans by (dx, dy) whole pixels per frame, exercising
    reference fetch + motion compensation in the decoder;
  * Baseline profile, CAVLC, pic_order_cnt_type 2 (display order == decode order, no reordering),
    deblocking disabled, one slice per picture; ISO-BMFF with avcC, moov before mdat.

Bitrate is unrealistic (an I_PCM 1080p picture is 3.1 MB; with GOP 30 a 10 s clip is ~33 MB instead of the
~5 MB of a 4 Mb/s encode), and there is no residual/entropy-decoding load on the P pictures: decode-rate
numbers measured on these clips are stated as such (DESIGN.md "Synthetic clips").
"""

from __future__ import annotations

import struct

import numpy as np


# ------------------------------------------------------------------------------------------ bit writing
class Bits:
    def __init__(self):
        self.s = []

    def u(self, n: int, v: int):
        self.s.append(format(v, f"0{n}b") if n else "")
        return self

    def ue(self, v: int):
        x = v + 1
        n = x.bit_length()
        self.s.append("0" * (n - 1) + format(x, "b"))
        return self

    def se(self, v: int):
        return self.ue(2 * v - 1 if v > 0 else -2 * v)

    def raw(self, bits: str):
        self.s.append(bits)
        return self

    def trailing(self):
        b = "".join(self.s) + "1"
        b += "0" * (-len(b) % 8)
        return int(b, 2).to_bytes(len(b) // 8, "big")

    def align_zero(self):
        b = "".join(self.s)
        b += "0" * (-len(b) % 8)
        self.s = [b]
        return self

    def tobytes_aligned(self):
        b = "".join(self.s)
        assert len(b) % 8 == 0
        return int(b, 2).to_bytes(len(b) // 8, "big") if b else b""


def escape(rbsp: bytes) -> bytes:
    """Emulation prevention: insert 0x03 after any 00 00 followed by a byte <= 3."""
    out = bytearray()
    zeros = 0
    for b in rbsp:
        if zeros >= 2 and b <= 3:
            out.append(3)
            zeros = 0
        out.append(b)
        zeros = zeros + 1 if b == 0 else 0
    return bytes(out)


def nal(ref_idc: int, typ: int, payload: bytes) -> bytes:
    return bytes([(ref_idc << 5) | typ]) + payload


# ------------------------------------------------------------------------------------------ parameter sets
def level_for(mbs: int, fps: float) -> int:
    rate = mbs * fps
    for lvl, max_fs, max_rate in ((30, 1620, 40500), (31, 3600, 108000), (40, 8192, 245760), (42, 8704, 522240), (50, 22080, 589824),
                                  (51, 36864, 983040), (52, 36864, 2073600)):  # fmt: skip
        if mbs <= max_fs and rate <= max_rate:
            return lvl
    return 52


def sps(width: int, height: int, fps: float) -> bytes:
    mbw, mbh = (width + 15) // 16, (height + 15) // 16
    b = Bits()
    b.u(8, 66).u(8, 0b11000000).u(8, level_for(mbw * mbh, fps))  # Baseline, constraint_set0/1
    b.ue(0)  # sps id
    b.ue(0)  # log2_max_frame_num_minus4 -> 4 bits
    b.ue(2)  # pic_order_cnt_type 2: output order == decoding order
    b.ue(1)  # max_num_ref_frames
    b.u(1, 0)  # gaps_in_frame_num_value_allowed_flag
    b.ue(mbw - 1).ue(mbh - 1)
    b.u(1, 1)  # frame_mbs_only_flag
    b.u(1, 1)  # direct_8x8_inference_flag
    crop_r, crop_b = mbw * 16 - width, mbh * 16 - height
    if crop_r or crop_b:
        b.u(1, 1).ue(0).ue(crop_r // 2).ue(0).ue(crop_b // 2)
    else:
        b.u(1, 0)
    b.u(1, 0)  # vui_parameters_present_flag
    return nal(3, 7, escape(b.trailing()))


def pps() -> bytes:
    b = Bits()
    b.ue(0).ue(0)  # pps id, sps id
    b.u(1, 0)  # entropy_coding_mode_flag: CAVLC
    b.u(1, 0)  # bottom_field_pic_order_in_frame_present_flag
    b.ue(0)  # num_slice_groups_minus1
    b.ue(0).ue(0)  # num_ref_idx_l0/l1_default_active_minus1
    b.u(1, 0).u(2, 0)  # weighted_pred_flag, weighted_bipred_idc
    b.se(0).se(0).se(0)  # pic_init_qp/qs, chroma_qp_index_offset
    b.u(1, 1)  # deblocking_filter_control_present_flag
    b.u(1, 0).u(1, 0)  # constrained_intra_pred_flag, redundant_pic_cnt_present_flag
    return nal(3, 8, escape(b.trailing()))


# ------------------------------------------------------------------------------------------ pictures
def idr_picture(y: np.ndarray, u: np.ndarray, v: np.ndarray, idr_id: int) -> bytes:
    """y [H16, W16], u/v [H16/2, W16/2] uint8 with every sample >= 1 (keeps the payload free of 00 00)."""
    h, w = y.shape
    mbh, mbw = h // 16, w // 16
    hdr = Bits()
    hdr.ue(0).ue(7).ue(0).u(4, 0).ue(idr_id & 0xFFFF)  # first_mb, slice_type I, pps, frame_num, idr_pic_id
    hdr.u(1, 0).u(1, 0)  # no_output_of_prior_pics_flag, long_term_reference_flag
    hdr.se(0)  # slice_qp_delta
    hdr.ue(1)  # disable_deblocking_filter_idc = 1
    hdr.ue(25)  # mb_type I_PCM of the first macroblock
    head = hdr.align_zero().tobytes_aligned()
    ymb = y.reshape(mbh, 16, mbw, 16).transpose(0, 2, 1, 3).reshape(mbh * mbw, 256)
    umb = u.reshape(mbh, 8, mbw, 8).transpose(0, 2, 1, 3).reshape(mbh * mbw, 64)
    vmb = v.reshape(mbh, 8, mbw, 8).transpose(0, 2, 1, 3).reshape(mbh * mbw, 64)
    body = np.empty((mbh * mbw, 386), dtype=np.uint8)
    body[:, 0], body[:, 1] = 0x0D, 0x00  # ue(25) = 0000 1101 0, then pcm alignment zero bits
    body[:, 2:258], body[:, 258:322], body[:, 322:386] = ymb, umb, vmb
    payload = escape(head) + body.reshape(-1)[2:].tobytes() + b"\x80"  # first MB's mb_type is in `head`
    return nal(3, 5, payload)


def p_picture(n_mbs: int, frame_num: int, mv_x_px: int, mv_y_px: int) -> bytes:
    b = Bits()
    b.ue(0).ue(5).ue(0).u(4, frame_num & 15)  # first_mb, slice_type P, pps, frame_num
    b.u(1, 0)  # num_ref_idx_active_override_flag
    b.u(1, 0)  # ref_pic_list_modification_flag_l0
    b.u(1, 0)  # adaptive_ref_pic_marking_mode_flag
    b.se(0)  # slice_qp_delta
    b.ue(1)  # disable_deblocking_filter_idc
    # first macroblock carries the global vector (quarter-pel), the rest predict it exactly (mvd = 0)
    b.ue(0).ue(0).se(4 * mv_x_px).se(4 * mv_y_px).ue(0)  # mb_skip_run, P_L0_16x16, mvd_x, mvd_y, cbp = 0
    b.raw("11111" * (n_mbs - 1))
    return nal(2, 1, escape(b.trailing()))


# ------------------------------------------------------------------------------------------ MP4
def box(typ: bytes, *payload: bytes) -> bytes:
    body = b"".join(payload)
    return struct.pack(">I4s", 8 + len(body), typ) + body


def full(typ: bytes, version: int, flags: int, *payload: bytes) -> bytes:
    return box(typ, struct.pack(">I", (version << 24) | flags), *payload)


def mux_mp4(samples: list[bytes], sync: list[bool], sps_nal: bytes, pps_nal: bytes, width: int, height: int, fps_num: int, fps_den: int = 1,
            sample_entry: bytes | None = None) -> bytes:
    """`sample_entry`: a complete visual sample entry box (e.g. hvc1 from tools/synth_hevc) instead of the avc1 entry built here."""
    timescale, delta = fps_num * 512 if fps_den == 1 else fps_num, 512 if fps_den == 1 else fps_den
    n = len(samples)
    duration = n * delta
    if sample_entry is None:
        avcc = box(b"avcC", bytes([1, sps_nal[1], sps_nal[2], sps_nal[3], 0xFF, 0xE1]), struct.pack(">H", len(sps_nal)), sps_nal, bytes([1]),
                   struct.pack(">H", len(pps_nal)), pps_nal)  # fmt: skip
        sample_entry = box(b"avc1", b"\x00" * 6, struct.pack(">H", 1), b"\x00" * 16, struct.pack(">HH", width, height), struct.pack(">II", 0x00480000, 0x00480000),
                           b"\x00" * 4, struct.pack(">H", 1), b"\x00" * 32, struct.pack(">Hh", 0x18, -1), avcc)  # fmt: skip
    stsd = full(b"stsd", 0, 0, struct.pack(">I", 1), sample_entry)
    stts = full(b"stts", 0, 0, struct.pack(">III", 1, n, delta))
    sync_ids = [i + 1 for i, s in enumerate(sync) if s]
    stss = full(b"stss", 0, 0, struct.pack(">I", len(sync_ids)), b"".join(struct.pack(">I", i) for i in sync_ids))
    stsc = full(b"stsc", 0, 0, struct.pack(">IIII", 1, 1, n, 1))  # one chunk holding every sample
    stsz = full(b"stsz", 0, 0, struct.pack(">II", 0, n), np.array([len(s) for s in samples], dtype=">u4").tobytes())

    def moov_with(chunk_offset: int) -> bytes:
        co64 = full(b"co64", 0, 0, struct.pack(">IQ", 1, chunk_offset))
        stbl = box(b"stbl", stsd, stts, stss, stsc, stsz, co64)
        dinf = box(b"dinf", full(b"dref", 0, 0, struct.pack(">I", 1), full(b"url ", 0, 1)))
        minf = box(b"minf", full(b"vmhd", 0, 1, b"\x00" * 8), dinf, stbl)
        mdhd = full(b"mdhd", 0, 0, struct.pack(">IIIIHH", 0, 0, timescale, duration, 0x55C4, 0))
        hdlr = full(b"hdlr", 0, 0, struct.pack(">I4s", 0, b"vide"), b"\x00" * 12, b"VideoHandler\x00")
        mdia = box(b"mdia", mdhd, hdlr, minf)
        matrix = struct.pack(">9I", 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000)
        tkhd = full(b"tkhd", 0, 3, struct.pack(">IIIII", 0, 0, 1, 0, duration * 1000 // timescale), b"\x00" * 8, struct.pack(">HHHH", 0, 0, 0, 0), matrix,
                    struct.pack(">II", width << 16, height << 16))  # fmt: skip
        trak = box(b"trak", tkhd, mdia)
        mvhd = full(b"mvhd", 0, 0, struct.pack(">IIII", 0, 0, 1000, duration * 1000 // timescale), struct.pack(">IH", 0x10000, 0x100), b"\x00" * 10, matrix,
                    b"\x00" * 24, struct.pack(">I", 2))  # fmt: skip
        return box(b"moov", mvhd, trak)

    ftyp = box(b"ftyp", b"isom", struct.pack(">I", 0x200), b"isomiso2avc1mp41")
    moov = moov_with(0)
    mdat_hdr_len = 16
    offset = len(ftyp) + len(moov) + mdat_hdr_len
    moov = moov_with(offset)
    payload_len = sum(len(s) for s in samples)
    mdat_hdr = struct.pack(">I4sQ", 1, b"mdat", mdat_hdr_len + payload_len)
    return b"".join([ftyp, moov, mdat_hdr, *samples])


# ------------------------------------------------------------------------------------------ clip generator
def source_picture(width: int, height: int, seed: int, gop: int) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Seeded moving-gradient + noise-block picture (SURVEY.md 8d C2 content), samples in 16..235/240, padded to MBs."""
    w16, h16 = (width + 15) // 16 * 16, (height + 15) // 16 * 16
    rng = np.random.default_rng(seed * 1000 + gop)
    yy, xx = np.mgrid[0:h16, 0:w16]
    base = 16 + ((xx * 2 + yy * 3 + gop * 41 + seed * 17) % 200)
    blocks = rng.integers(-20, 21, size=(h16 // 8, w16 // 8)).repeat(8, 0).repeat(8, 1)
    fine = rng.integers(-6, 7, size=(h16, w16))
    y = np.clip(base + blocks + fine, 16, 235).astype(np.uint8)
    u = np.clip(128 + rng.integers(-60, 61, size=(h16 // 16, w16 // 16)).repeat(8, 0).repeat(8, 1) + rng.integers(-4, 5, size=(h16 // 2, w16 // 2)), 16, 240).astype(np.uint8)
    v = np.clip(128 + rng.integers(-60, 61, size=(h16 // 16, w16 // 16)).repeat(8, 0).repeat(8, 1) + rng.integers(-4, 5, size=(h16 // 2, w16 // 2)), 16, 240).astype(np.uint8)
    return y, u, v


def make_clip(width: int, height: int, fps: int, seconds: float, seed: int = 0, gop: int | None = None, pan: tuple[int, int] = (2, 0),
              return_sources: bool = False):
    """Returns mp4 bytes (and, optionally, {frame_index: (y, u, v)} of the IDR source pictures, cropped)."""
    gop = fps if gop is None else gop
    n_frames = int(round(fps * seconds))
    w16, h16 = (width + 15) // 16 * 16, (height + 15) // 16 * 16
    n_mbs = (w16 // 16) * (h16 // 16)
    s, p = sps(width, height, fps), pps()
    samples, sync, sources = [], [], {}
    for i in range(n_frames):
        if i % gop == 0:
            y, u, v = source_picture(width, height, seed, i // gop)
            nalu = idr_picture(y, u, v, i // gop)
            sync.append(True)
            if return_sources:
                sources[i] = (y[:height, :width].copy(), u[: height // 2, : width // 2].copy(), v[: height // 2, : width // 2].copy())
        else:
            nalu = p_picture(n_mbs, i % gop, pan[0], pan[1])
            sync.append(False)
        samples.append(struct.pack(">I", len(nalu)) + nalu)
    mp4 = mux_mp4(samples, sync, s, p, width, height, fps)
    return (mp4, sources) if return_sources else mp4


# =========================================================================================== residual-coded clips
# mode="coded": pictures whose macroblocks carry CAVLC-coded transform coefficients at a real-world bit rate (default
# 4 Mb/s, the reference's transcode default, decoder_utils.py:43) instead of raw I_PCM samples:
#
#   * IDR pictures: Intra16x16 macroblocks, DC prediction, one luma DC level per macroblock chosen in closed loop so the
#     picture reproduces a seeded 16x16 mosaic of the source picture exactly before deblocking (at QP 28 a DC level c adds
#     exactly c to the 256 luma samples), optional sparse luma AC coefficients (texture) and chroma DC levels;
#   * P pictures: runs of P_Skip, P_L0_16x16 macroblocks with quarter-pel motion-vector differences and sparse 4x4 residual
#     blocks (coded_block_pattern over the four 8x8 quadrants + chroma DC), a few Intra16x16 macroblocks; the share of coded
#     macroblocks is solved for the requested bit rate;
#   * in-loop deblocking ON (disable_deblocking_filter_idc = 0), Baseline / CAVLC, one slice per picture, closed GOP.
#
# The generator is OPEN LOOP for the P pictures (it emits legal syntax and does not track the reconstruction - any decoder
# reconstructs the same pictures, which is what the NVDEC-vs-libavcodec bit-exactness tests check), so every block is
# limited to at most one coefficient: then nC (the neighbour-count context of coeff_token) is always 0 or 1, one VLC table
# serves every block, and a macroblock's bits do not depend on its neighbours - the whole picture is assembled with numpy.
# A parse error anywhere in a slice would desynchronise everything after it (CAVLC has no resync inside a slice): the LAST
# macroblock of every picture is an I_PCM sentinel with known samples, which tests read back from the decoded frames.

_TZ4x4 = [(1, 1), (3, 3), (2, 3), (3, 4), (2, 4), (3, 5), (2, 5), (3, 6), (2, 6), (3, 7), (2, 7), (3, 8), (2, 8), (3, 9), (2, 9), (1, 9)]
_TZ_CDC = [(1, 1), (1, 2), (1, 3), (0, 3)]
# coded_block_pattern me(v) code numbers for Inter macroblocks (subset of Table 9-4 used here)
_CBP_INTER_CODE = {0: 0, 16: 1, 1: 2, 2: 3, 4: 4, 8: 5, 32: 6, 3: 7, 5: 8, 10: 9, 12: 10, 15: 11, 47: 12, 31: 19}
QP_CODED = 28  # slice QP: a DC-only Intra16x16 level c is +c on every luma sample, a chroma DC level c is +2c


def _ue_tok(v):
    """(value, nbits) arrays of Exp-Golomb ue(v)."""
    v = np.asarray(v, dtype=np.int64) + 1
    n = np.floor(np.log2(v)).astype(np.int64) + 1
    return v, 2 * n - 1


def _se_tok(v):
    v = np.asarray(v, dtype=np.int64)
    return _ue_tok(np.where(v > 0, 2 * v - 1, -2 * v))


def _level_tok(level):
    """First (and only) non-trailing-one level of a block with suffixLength 0: |level| >= 2."""
    level = np.asarray(level, dtype=np.int64)
    code = 2 * np.abs(level) - 4 + (level < 0)
    val = np.ones_like(code)
    nb = code + 1
    mid = (code >= 14) & (code < 30)
    val = np.where(mid, (1 << 4) | (code - 14), val)
    nb = np.where(mid, 19, nb)
    big = code >= 30
    val = np.where(big, (1 << 12) | (code - 30), val)
    nb = np.where(big, 28, nb)
    return val, nb


def _block_toks(level, pos, tz_table, tok0, tok_t1, tok_lv):
    """Token columns [n, 3] (value, nbits) of residual_block_cavlc for blocks with <= 1 coefficient.
    level: coefficient value (0 = empty block), pos: its scan index (= total_zeros)."""
    level = np.asarray(level, dtype=np.int64)
    n = level.shape[0]
    tzv = np.array([t[0] for t in tz_table], dtype=np.int64)[pos]
    tzn = np.array([t[1] for t in tz_table], dtype=np.int64)[pos]
    val = np.zeros((n, 3), dtype=np.int64)
    nb = np.zeros((n, 3), dtype=np.int64)
    empty, one, big = level == 0, np.abs(level) == 1, np.abs(level) >= 2
    val[empty, 0], nb[empty, 0] = tok0
    # trailing one: coeff_token, sign flag, total_zeros
    val[one, 0], nb[one, 0] = tok_t1
    val[one, 1], nb[one, 1] = (level[one] < 0).astype(np.int64), 1
    val[one, 2], nb[one, 2] = tzv[one], tzn[one]
    lv, ln = _level_tok(np.where(big, level, 2))
    val[big, 0], nb[big, 0] = tok_lv
    val[big, 1], nb[big, 1] = lv[big], ln[big]
    val[big, 2], nb[big, 2] = tzv[big], tzn[big]
    return val, nb


def _luma_block(level, pos):
    return _block_toks(level, pos, _TZ4x4, (1, 1), (1, 2), (0b000101, 6))


def _chroma_dc_block(level, pos):
    return _block_toks(level, pos, _TZ_CDC, (0b01, 2), (1, 1), (0b000111, 6))


def _pack_tokens(val: np.ndarray, nb: np.ndarray) -> np.ndarray:
    """Concatenate variable-length codes (value, nbits <= 32) into a bit array (uint8 0/1)."""
    val, nb = val.reshape(-1), nb.reshape(-1)
    keep = nb > 0
    val, nb = val[keep], nb[keep]
    shifts = np.arange(31, -1, -1, dtype=np.int64)
    bits = ((val[:, None] >> shifts[None, :]) & 1).astype(np.uint8)
    mask = shifts[None, :] < nb[:, None]
    return bits[mask]


def _escape_fast(rbsp: np.ndarray) -> bytes:
    """Emulation prevention on a uint8 array (vectorised candidate search, sequential rule on the few candidates)."""
    z = rbsp == 0
    cand = np.flatnonzero(z[:-2] & z[1:-1] & (rbsp[2:] <= 3)) + 2
    if cand.size == 0:
        return rbsp.tobytes()
    out, last, floor = [], 0, 0
    for i in cand.tolist():
        if i - 2 < floor:  # one of the two zeros was already consumed by the previous insertion
            continue
        out.append(rbsp[last:i].tobytes())
        out.append(b"\x03")
        last, floor = i, i
    out.append(rbsp[last:].tobytes())
    return b"".join(out)


def sentinel_samples(seed: int) -> np.ndarray:
    """The 384 raw samples (256 Y, 64 Cb, 64 Cr) of the I_PCM macroblock that ends every coded picture."""
    return np.random.default_rng(0xC0DED + seed).integers(16, 236, size=384).astype(np.uint8)


def _finish_picture(tok_val, tok_nb, pcm: np.ndarray) -> np.ndarray:
    """slice bits (header + macroblocks up to and including the sentinel's mb_type) -> RBSP bytes with the PCM payload."""
    bits = _pack_tokens(tok_val, tok_nb)
    pad = (-bits.size) % 8  # pcm_alignment_zero_bit
    if pad:
        bits = np.concatenate([bits, np.zeros(pad, dtype=np.uint8)])
    return np.concatenate([np.packbits(bits), pcm, np.array([0x80], dtype=np.uint8)])  # rbsp_slice_trailing_bits


def _hdr_tokens(fields):
    v = np.array([f[0] for f in fields], dtype=np.int64)
    n = np.array([f[1] for f in fields], dtype=np.int64)
    return v, n


def _tok(fn, x):
    v, n = fn(np.array([x]))
    return int(v[0]), int(n[0])


def _deblock_fields(deblock: bool):
    if deblock:
        return [_tok(_ue_tok, 0), _tok(_se_tok, 0), _tok(_se_tok, 0)]  # idc 0 + alpha/beta offsets
    return [_tok(_ue_tok, 1)]


def coded_idr_picture(y_mosaic: np.ndarray, u_mosaic: np.ndarray, v_mosaic: np.ndarray, idr_id: int, rng, pcm: np.ndarray,
                      ac_density: float = 0.25, deblock: bool = True) -> bytes:
    """y/u/v_mosaic: int arrays [mbh, mbw] - the target value of every macroblock (luma reproduced exactly before deblocking
    when ac_density == 0).  Returns the IDR slice NAL unit."""
    mbh, mbw = y_mosaic.shape
    n = mbh * mbw
    T = y_mosaic.astype(np.int64)
    pred = np.full((mbh, mbw), 128, dtype=np.int64)
    pred[0, 1:] = T[0, :-1]
    pred[1:, 0] = T[:-1, 0]
    pred[1:, 1:] = (T[:-1, 1:] + T[1:, :-1] + 1) >> 1  # flat neighbours: (16 top + 16 left + 16) >> 5
    c_y = (T - pred).reshape(-1)

    def chroma_levels(m):
        m = m.astype(np.int64)
        p = np.full((mbh, mbw), 128, dtype=np.int64)
        p[0, 1:] = m[0, :-1]
        p[1:, 0] = m[:-1, 0]
        p[1:, 1:] = (m[:-1, 1:] + m[1:, :-1] + 1) >> 1
        return np.clip(np.round((m - p) / 2.0), -60, 60).astype(np.int64).reshape(-1)  # a chroma DC level adds 2c

    c_u, c_v = chroma_levels(u_mosaic), chroma_levels(v_mosaic)
    has_ac = rng.random(n) < ac_density
    has_ac[-1] = False
    cbp_chroma = ((c_u != 0) | (c_v != 0)).astype(np.int64)
    mb_type = 1 + 2 + 4 * cbp_chroma + 12 * has_ac  # Intra16x16, DC prediction
    cols_v, cols_n = [], []

    def add(v, nbits):
        cols_v.append(np.asarray(v, dtype=np.int64).reshape(n, -1))
        cols_n.append(np.asarray(nbits, dtype=np.int64).reshape(n, -1))

    add(*_ue_tok(mb_type))
    add(np.ones(n), np.ones(n))  # intra_chroma_pred_mode ue(0)
    add(np.ones(n), np.ones(n))  # mb_qp_delta se(0)
    add(*_luma_block(c_y, np.zeros(n, dtype=np.int64)))  # Intra16x16DCLevel: the DC-of-DC coefficient only
    for _ in range(16):  # Intra16x16ACLevel blocks (15 coefficients each) of macroblocks with luma cbp 15
        lvl = np.where(rng.random(n) < 0.35, rng.choice(np.array([-2, -1, -1, 1, 1, 2]), size=n), 0) * has_ac
        v3, n3 = _luma_block(lvl, rng.integers(0, 6, size=n))
        n3 *= has_ac[:, None]
        add(v3, n3)
    for c in (c_u, c_v):
        v3, n3 = _chroma_dc_block(c, np.zeros(n, dtype=np.int64))
        n3 *= cbp_chroma[:, None]
        add(v3, n3)
    mb_v, mb_n = np.concatenate(cols_v, axis=1), np.concatenate(cols_n, axis=1)
    mb_v, mb_n = mb_v[:-1], mb_n[:-1]  # the last macroblock is the I_PCM sentinel
    hdr = [_tok(_ue_tok, 0), _tok(_ue_tok, 7), _tok(_ue_tok, 0), (0, 4), _tok(_ue_tok, idr_id & 0xFFFF), (0, 1), (0, 1),
           _tok(_se_tok, QP_CODED - 26), *_deblock_fields(deblock)]  # fmt: skip
    hv, hn = _hdr_tokens(hdr)
    tv, tn = _ue_tok(np.array([25]))  # mb_type I_PCM
    rbsp = _finish_picture(np.concatenate([hv, mb_v.reshape(-1), tv]), np.concatenate([hn, mb_n.reshape(-1), tn]), pcm)
    return nal(3, 5, _escape_fast(rbsp))


def coded_p_body(mbw: int, mbh: int, rng, coded_share: float, pcm: np.ndarray, pan_qpel: tuple[int, int] = (0, 0), deblock: bool = True,
                 intra_share: float = 0.04, cbp_set=None, mvd_share: float = 0.3):
    """One P picture (frame_num 0 in the header; p_picture_from_body patches it).  Returns (NAL payload bytes, stats)."""
    n = mbw * mbh
    coded = rng.random(n) < coded_share
    coded[0] = True  # carries the global motion vector difference
    coded[-1] = False  # sentinel handled separately
    idx = np.flatnonzero(coded)
    k = idx.size
    run = np.diff(np.concatenate([[-1], idx])) - 1  # skipped macroblocks in front of every coded one
    intra = rng.random(k) < intra_share
    intra[0] = False
    # coded_block_pattern: mostly one or two 8x8 quadrants, sometimes everything, sometimes chroma only
    cbp_choices = np.array([1, 2, 4, 8, 3, 12, 5, 10, 15, 16, 31, 47, 0])
    cbp_p = np.array([0.11, 0.11, 0.11, 0.11, 0.06, 0.06, 0.06, 0.06, 0.10, 0.04, 0.04, 0.04, 0.10])
    if cbp_set is not None:
        cbp_p = np.array([pr if c in cbp_set else 0.0 for c, pr in zip(cbp_choices, cbp_p)])
    cbp = rng.choice(cbp_choices, size=k, p=cbp_p / cbp_p.sum())
    mvd = np.where(rng.random((k, 2)) < mvd_share, rng.integers(-6, 7, size=(k, 2)), 0)
    mvd[0] = pan_qpel
    cols_v, cols_n = [], []

    def add(v, nbits, on=None):
        v = np.asarray(v, dtype=np.int64).reshape(k, -1)
        nbits = np.asarray(nbits, dtype=np.int64).reshape(k, -1)
        if on is not None:
            nbits = nbits * on[:, None]
        cols_v.append(v)
        cols_n.append(nbits)

    inter = ~intra
    add(*_ue_tok(run))
    i16_chroma = (rng.random(k) < 0.5).astype(np.int64)
    add(*_ue_tok(np.where(intra, 5 + 1 + 2 + 4 * i16_chroma, 0)))  # P_L0_16x16 or Intra16x16 (DC prediction, no luma AC)
    add(*_se_tok(mvd[:, 0]), on=inter)
    add(*_se_tok(mvd[:, 1]), on=inter)
    add(np.ones(k), np.ones(k), on=intra)  # intra_chroma_pred_mode ue(0)
    add(*_ue_tok(np.vectorize(_CBP_INTER_CODE.get)(cbp)), on=inter)
    add(np.ones(k), np.ones(k), on=(intra | (cbp > 0)))  # mb_qp_delta se(0)
    add(*_luma_block(rng.integers(-12, 13, size=k), np.zeros(k, dtype=np.int64)), on=intra)  # Intra16x16DCLevel
    for q in range(4):
        on_q = inter & ((cbp >> q) & 1).astype(bool)
        first = True
        for _ in range(4):
            lvl = np.where(rng.random(k) < (0.9 if first else 0.45), rng.choice(np.array([-3, -2, -1, -1, -1, 1, 1, 1, 2, 3]), size=k), 0)
            add(*_luma_block(lvl, rng.integers(0, 10, size=k)), on=on_q)
            first = False
    chroma_on = (inter & (cbp >= 16)) | (intra & (i16_chroma > 0))
    for _ in range(2):
        lvl = np.where(rng.random(k) < 0.7, rng.choice(np.array([-2, -1, -1, 1, 1, 2]), size=k), 0)
        add(*_chroma_dc_block(lvl, rng.integers(0, 4, size=k)), on=chroma_on)
    add(np.full(k, 0xFF), np.full(k, 8), on=inter & (cbp >= 32))  # chroma cbp 2: eight chroma AC blocks, all empty ('1' each)
    mb_v, mb_n = np.concatenate(cols_v, axis=1), np.concatenate(cols_n, axis=1)
    tail_run = n - 1 - (int(idx[-1]) + 1)  # skipped macroblocks between the last coded one and the sentinel
    tv, tn = _hdr_tokens([_tok(_ue_tok, tail_run), _tok(_ue_tok, 30)])  # mb_skip_run, mb_type I_PCM in a P slice (5 + 25)
    hdr = [_tok(_ue_tok, 0), _tok(_ue_tok, 5), _tok(_ue_tok, 0), (0, 4), (0, 1), (0, 1), (0, 1), _tok(_se_tok, QP_CODED - 26), *_deblock_fields(deblock)]
    hv, hn = _hdr_tokens(hdr)
    rbsp = _finish_picture(np.concatenate([hv, mb_v.reshape(-1), tv]), np.concatenate([hn, mb_n.reshape(-1), tn]), pcm)
    body = bytearray(_escape_fast(rbsp))
    return body, {"coded_mbs": int(k), "bytes": len(body)}


def p_picture_from_body(body: bytearray, frame_num: int) -> bytes:
    """frame_num u(4) sits at slice-header bits 7..10: first_mb '1', slice_type ue(5) '00110', pps '1', then the 4 bits."""
    f = frame_num & 15
    b = bytearray(body)
    b[0] = (b[0] & 0xFE) | (f >> 3)
    b[1] = (b[1] & 0x1F) | ((f & 7) << 5)
    return nal(2, 1, bytes(b))


def mosaic_of(width: int, height: int, seed: int, gop_index: int):
    """Macroblock means of the seeded source picture (the mosaic the coded IDR pictures reproduce)."""
    y, u, v = source_picture(width, height, seed, gop_index)
    h16, w16 = y.shape
    ym = y.reshape(h16 // 16, 16, w16 // 16, 16).astype(np.int64).mean(axis=(1, 3)).round().astype(np.int64)
    um = u.reshape(h16 // 16, 8, w16 // 16, 8).astype(np.int64).mean(axis=(1, 3)).round().astype(np.int64)
    vm = v.reshape(h16 // 16, 8, w16 // 16, 8).astype(np.int64).mean(axis=(1, 3)).round().astype(np.int64)
    return np.clip(ym, 20, 230), np.clip(um, 20, 235), np.clip(vm, 20, 235)


def make_coded_clip(width: int, height: int, fps: int, seconds: float, seed: int = 0, gop: int | None = None, bitrate: float = 4e6,
                    p_variants: int = 6, deblock: bool = True, ac_density: float = 0.25, return_info: bool = False):
    """Residual-coded clip at ~`bitrate` b/s (see the section comment).  Returns mp4 bytes (and an info dict)."""
    gop = fps if gop is None else gop
    n_frames = int(round(fps * seconds))
    mbw, mbh = (width + 15) // 16, (height + 15) // 16
    rng = np.random.default_rng(0x5EED0000 + seed)
    pcm = sentinel_samples(seed)
    s, p = sps(width, height, fps), pps()
    n_idr = (n_frames + gop - 1) // gop
    idrs = [coded_idr_picture(*mosaic_of(width, height, seed, g), g, rng, pcm, ac_density=ac_density, deblock=deblock) for g in range(n_idr)]
    idr_bytes = sum(len(x) for x in idrs)
    target_p = max(2000.0, (bitrate / 8.0 * seconds - idr_bytes) / max(1, n_frames - n_idr))  # bytes per P picture
    # solve the coded-macroblock share for the byte target with two probes (bytes are linear in the share)
    probe = [coded_p_body(mbw, mbh, np.random.default_rng(seed + 77 + i), sh, pcm, deblock=deblock)[1]["bytes"] for i, sh in enumerate((0.05, 0.25))]
    slope = (probe[1] - probe[0]) / 0.20
    share = float(np.clip(0.05 + (target_p - probe[0]) / max(slope, 1.0), 0.002, 0.9))
    bodies = [coded_p_body(mbw, mbh, rng, share, pcm, pan_qpel=(4 * (i % 3), 0), deblock=deblock)[0] for i in range(p_variants)]
    samples, sync = [], []
    for i in range(n_frames):
        if i % gop == 0:
            nalu = idrs[i // gop]
            sync.append(True)
        else:
            nalu = p_picture_from_body(bodies[i % p_variants], i % gop)
            sync.append(False)
        samples.append(struct.pack(">I", len(nalu)) + nalu)
    mp4 = mux_mp4(samples, sync, s, p, width, height, fps)
    if return_info:
        total = sum(len(x) for x in samples)
        return mp4, {"bitrate": 8.0 * total / seconds, "idr_bytes": idr_bytes / n_idr, "p_bytes": (total - idr_bytes) / max(1, n_frames - n_idr),
                     "coded_share": share, "mosaics": [mosaic_of(width, height, seed, g)[0] for g in range(n_idr)], "pcm": pcm}  # fmt: skip
    return mp4
