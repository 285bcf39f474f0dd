// Minimal ISO-BMFF (MP4) video-track index + AVCC/HVCC -> Annex-B conversion.
// Replaces the libavformat demux the reference reaches through PyAV (decoder_utils.py:260-276) and
// PyNvDemuxer (nvcodec_utils.py:224-232) for the container layout the pipeline itself writes
// (clip_extraction_stages.py transcodes every clip to H.264/HEVC MP4).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace cb {

struct Mp4Track {
  int codec = -1;  // 4 = H.264, 8 = HEVC (cudaVideoCodec numbering)
  int width = 0, height = 0;
  uint32_t timescale = 0;
  uint64_t duration = 0;
  int nal_length_size = 4;
  bool has_ctts = false;
  std::vector<uint8_t> param_sets_annexb;  // SPS/PPS (and VPS) with start codes
  std::vector<uint64_t> offset;            // per sample (decode order)
  std::vector<uint32_t> size;
  std::vector<int64_t> pts;                // composition time in `timescale` ticks, edit list applied
  std::vector<int64_t> dts;
  std::vector<uint8_t> sync;               // 1 = sync sample
  size_t stsd_off = 0, stsd_size = 0;      // the whole stsd box inside the source buffer (copied verbatim by mp4_cut)
};

// Returns empty string on success, else a reason.
std::string mp4_parse(const uint8_t* data, size_t size, Mp4Track* out);
// Appends sample `i` as Annex-B (start codes instead of length prefixes) to `dst`; false if malformed.
bool mp4_sample_annexb(const uint8_t* data, size_t size, const Mp4Track& t, size_t i, std::vector<uint8_t>* dst);

// Stream copy of samples [first, first + count) (decode order; `first` must be a sync sample) into a standalone MP4: the
// source's stsd (avcC / hvcC) verbatim, timestamps re-based so the first sample decodes at 0, an edit list when composition
// offsets delay the first presented frame.  No bit of the coded pictures is touched.  Returns "" or a reason.
std::string mp4_cut(const uint8_t* data, size_t size, const Mp4Track& t, size_t first, size_t count, std::vector<uint8_t>* out);

}  // namespace cb
