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


def mux_mp4(samples: list[bytes], sync: list[bool], sps_nal: bytes, pps_nal: bytes, width: int, height: int, fps_num: int, fps_den: int = 1) -> bytes:
    timescale, delta = fps_num * 512 if fps_den == 1 else fps_num, 512 if fps_den == 1 else fps_den
    n = len(samples)
    duration = n * delta
    avcc = box(b"avcC", bytes([1, sps_nal[1], sps_nal[2], sps_nal[3], 0xFF, 0xE1]), struct.pack(">H", len(sps_nal)), sps_nal, bytes([1]),
               struct.pack(">H", len(pps_nal)), pps_nal)  # fmt: skip
    avc1 = box(b"avc1", b"\x00" * 6, struct.pack(">H", 1), b"\x00" * 16, struct.pack(">HH", width, height), struct.pack(">II", 0x00480000, 0x00480000),
               b"\x00" * 4, struct.pack(">H", 1), b"\x00" * 32, struct.pack(">Hh", 0x18, -1), avcc)  # fmt: skip
    stsd = full(b"stsd", 0, 0, struct.pack(">I", 1), avc1)
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
