// ISO-BMFF video track index (see mp4_demux.h).  Bounds-checked: malformed input yields an error string,
// never a crash - the stage records it in clip.errors like the reference does for undecodable clips
// (clip_frame_extraction_stages.py:160-165).
#include "mp4_demux.h"

#include <climits>
#include <cstdint>

#include <string.h>

#include <algorithm>

namespace cb {
namespace {

struct Reader {
  const uint8_t* p;
  size_t n;
  bool ok(size_t off, size_t len) const { return off <= n && len <= n - off; }
  uint32_t u32(size_t off) const { return (uint32_t)p[off] << 24 | (uint32_t)p[off + 1] << 16 | (uint32_t)p[off + 2] << 8 | p[off + 3]; }
  uint64_t u64(size_t off) const { return (uint64_t)u32(off) << 32 | u32(off + 4); }
  uint16_t u16(size_t off) const { return (uint16_t)(p[off] << 8 | p[off + 1]); }
};

struct Box {
  uint32_t type;
  size_t body, end;  // body offset, end offset (absolute)
};

// Iterate child boxes of [begin, end).  Returns false on a malformed header.
template <typename F>
bool for_boxes(const Reader& r, size_t begin, size_t end, F&& fn) {
  size_t off = begin;
  while (off + 8 <= end) {
    uint64_t sz = r.u32(off);
    const uint32_t type = r.u32(off + 4);
    size_t hdr = 8;
    if (sz == 1) {
      if (off + 16 > end) return false;
      sz = r.u64(off + 8);
      hdr = 16;
    } else if (sz == 0) {
      sz = end - off;
    }
    if (sz < hdr || sz > end - off) return false;
    Box b{type, off + hdr, off + (size_t)sz};
    if (!fn(b)) return true;  // callback asked to stop
    off += (size_t)sz;
  }
  return true;
}

constexpr uint32_t fourcc(const char (&s)[5]) { return (uint32_t)s[0] << 24 | (uint32_t)s[1] << 16 | (uint32_t)s[2] << 8 | (uint32_t)s[3]; }

bool find_box(const Reader& r, size_t begin, size_t end, uint32_t type, Box* out) {
  bool found = false;
  for_boxes(r, begin, end, [&](const Box& b) {
    if (b.type == type) {
      *out = b;
      found = true;
      return false;
    }
    return true;
  });
  return found;
}

void append_start_code(std::vector<uint8_t>* v) {
  static const uint8_t sc[4] = {0, 0, 0, 1};
  v->insert(v->end(), sc, sc + 4);
}

std::string parse_avcc(const Reader& r, const Box& b, Mp4Track* t) {
  size_t o = b.body;
  if (b.end - o < 7) return "avcC too short";
  t->nal_length_size = (r.p[o + 4] & 3) + 1;
  int nsps = r.p[o + 5] & 31;
  o += 6;
  for (int pass = 0; pass < 2; ++pass) {
    int cnt = pass == 0 ? nsps : (o < b.end ? r.p[o++] : 0);
    for (int i = 0; i < cnt; ++i) {
      if (o + 2 > b.end) return "avcC truncated";
      const size_t len = r.u16(o);
      o += 2;
      if (len > b.end - o) return "avcC parameter set overruns box";
      append_start_code(&t->param_sets_annexb);
      t->param_sets_annexb.insert(t->param_sets_annexb.end(), r.p + o, r.p + o + len);
      o += len;
    }
  }
  return "";
}

std::string parse_hvcc(const Reader& r, const Box& b, Mp4Track* t) {
  size_t o = b.body;
  if (b.end - o < 23) return "hvcC too short";
  t->nal_length_size = (r.p[o + 21] & 3) + 1;
  const int arrays = r.p[o + 22];
  o += 23;
  for (int a = 0; a < arrays; ++a) {
    if (o + 3 > b.end) return "hvcC truncated";
    const int cnt = r.u16(o + 1);
    o += 3;
    for (int i = 0; i < cnt; ++i) {
      if (o + 2 > b.end) return "hvcC truncated";
      const size_t len = r.u16(o);
      o += 2;
      if (len > b.end - o) return "hvcC parameter set overruns box";
      append_start_code(&t->param_sets_annexb);
      t->param_sets_annexb.insert(t->param_sets_annexb.end(), r.p + o, r.p + o + len);
      o += len;
    }
  }
  return "";
}

std::string parse_trak(const Reader& r, const Box& trak, uint32_t movie_timescale, Mp4Track* t, bool* is_video) {
  *is_video = false;
  Box mdia, hdlr, mdhd, minf, stbl;
  if (!find_box(r, trak.body, trak.end, fourcc("mdia"), &mdia)) return "no mdia";
  if (!find_box(r, mdia.body, mdia.end, fourcc("hdlr"), &hdlr) || hdlr.end - hdlr.body < 12) return "no hdlr";
  if (r.u32(hdlr.body + 8) != fourcc("vide")) return "";
  *is_video = true;
  if (!find_box(r, mdia.body, mdia.end, fourcc("mdhd"), &mdhd) || mdhd.end - mdhd.body < 24) return "no mdhd";
  const int mdhd_ver = r.p[mdhd.body];
  if (mdhd_ver == 1) {
    if (mdhd.end - mdhd.body < 36) return "mdhd truncated";
    t->timescale = r.u32(mdhd.body + 20);
    t->duration = r.u64(mdhd.body + 24);
  } else {
    t->timescale = r.u32(mdhd.body + 12);
    t->duration = r.u32(mdhd.body + 16);
  }
  if (t->timescale == 0) return "zero timescale";
  if (!find_box(r, mdia.body, mdia.end, fourcc("minf"), &minf) || !find_box(r, minf.body, minf.end, fourcc("stbl"), &stbl)) return "no stbl";

  // ---- stsd: first sample entry
  Box stsd;
  if (!find_box(r, stbl.body, stbl.end, fourcc("stsd"), &stsd) || stsd.end - stsd.body < 16) return "no stsd";
  t->stsd_off = stsd.body - 8, t->stsd_size = stsd.end - (stsd.body - 8);
  if (r.u32(stsd.body - 8) == 1) return "64-bit stsd box";  // never written by any muxer; keeps the verbatim copy simple
  {
    const size_t e = stsd.body + 8;  // first entry: size, format
    const uint32_t esz = r.u32(e), fmt = r.u32(e + 4);
    if (esz < 86 || esz > stsd.end - e) return "bad sample entry";
    t->width = r.u16(e + 32);
    t->height = r.u16(e + 34);
    const size_t child = e + 86, eend = e + esz;
    Box cfg;
    if (fmt == fourcc("avc1") || fmt == fourcc("avc3")) {
      t->codec = 4;
      if (!find_box(r, child, eend, fourcc("avcC"), &cfg)) return "no avcC";
      std::string err = parse_avcc(r, cfg, t);
      if (!err.empty()) return err;
    } else if (fmt == fourcc("hvc1") || fmt == fourcc("hev1")) {
      t->codec = 8;
      if (!find_box(r, child, eend, fourcc("hvcC"), &cfg)) return "no hvcC";
      std::string err = parse_hvcc(r, cfg, t);
      if (!err.empty()) return err;
    } else {
      char name[5] = {(char)(fmt >> 24), (char)(fmt >> 16), (char)(fmt >> 8), (char)fmt, 0};
      return std::string("unsupported sample entry '") + name + "'";
    }
  }

  // ---- stsz
  Box b;
  if (!find_box(r, stbl.body, stbl.end, fourcc("stsz"), &b) || b.end - b.body < 12) return "no stsz";
  const uint32_t uniform = r.u32(b.body + 4), count = r.u32(b.body + 8);
  if (count > (1u << 24)) return "implausible sample count";
  if (uniform == 0 && (size_t)count * 4 > b.end - (b.body + 12)) return "stsz truncated";
  t->size.resize(count);
  for (uint32_t i = 0; i < count; ++i) t->size[i] = uniform ? uniform : r.u32(b.body + 12 + (size_t)i * 4);

  // ---- chunk offsets
  std::vector<uint64_t> chunk_off;
  if (find_box(r, stbl.body, stbl.end, fourcc("stco"), &b)) {
    if (b.end - b.body < 8) return "stco truncated";
    const uint32_t n = r.u32(b.body + 4);
    if ((size_t)n * 4 > b.end - (b.body + 8)) return "stco truncated";
    for (uint32_t i = 0; i < n; ++i) chunk_off.push_back(r.u32(b.body + 8 + (size_t)i * 4));
  } else if (find_box(r, stbl.body, stbl.end, fourcc("co64"), &b)) {
    if (b.end - b.body < 8) return "co64 truncated";
    const uint32_t n = r.u32(b.body + 4);
    if ((size_t)n * 8 > b.end - (b.body + 8)) return "co64 truncated";
    for (uint32_t i = 0; i < n; ++i) chunk_off.push_back(r.u64(b.body + 8 + (size_t)i * 8));
  } else {
    return "no chunk offsets";
  }

  // ---- stsc -> per-sample offsets
  if (!find_box(r, stbl.body, stbl.end, fourcc("stsc"), &b) || b.end - b.body < 8) return "no stsc";
  {
    const uint32_t n = r.u32(b.body + 4);
    if ((size_t)n * 12 > b.end - (b.body + 8)) return "stsc truncated";
    t->offset.resize(count);
    uint32_t s = 0;
    for (uint32_t e = 0; e < n && s < count; ++e) {
      const uint32_t first = r.u32(b.body + 8 + (size_t)e * 12), per = r.u32(b.body + 12 + (size_t)e * 12);
      const uint32_t next_first = (e + 1 < n) ? r.u32(b.body + 8 + (size_t)(e + 1) * 12) : (uint32_t)chunk_off.size() + 1;
      if (first == 0 || next_first < first) return "bad stsc";
      for (uint32_t c = first; c < next_first && s < count; ++c) {
        if (c > chunk_off.size()) return "stsc references a missing chunk";
        uint64_t off = chunk_off[c - 1];
        for (uint32_t k = 0; k < per && s < count; ++k) {
          t->offset[s] = off;
          off += t->size[s];
          ++s;
        }
      }
    }
    if (s != count) return "stsc does not cover all samples";
  }

  // ---- stts (+ ctts) -> dts / pts
  if (!find_box(r, stbl.body, stbl.end, fourcc("stts"), &b) || b.end - b.body < 8) return "no stts";
  {
    const uint32_t n = r.u32(b.body + 4);
    if ((size_t)n * 8 > b.end - (b.body + 8)) return "stts truncated";
    t->dts.resize(count);
    int64_t cur = 0;
    uint32_t s = 0;
    for (uint32_t e = 0; e < n && s < count; ++e) {
      const uint32_t cnt = r.u32(b.body + 8 + (size_t)e * 8), delta = r.u32(b.body + 12 + (size_t)e * 8);
      for (uint32_t k = 0; k < cnt && s < count; ++k) {
        t->dts[s++] = cur;
        cur += delta;
      }
    }
    if (s != count) return "stts does not cover all samples";
  }
  t->pts = t->dts;
  if (find_box(r, stbl.body, stbl.end, fourcc("ctts"), &b) && b.end - b.body >= 8) {
    const int ver = r.p[b.body];
    const uint32_t n = r.u32(b.body + 4);
    if ((size_t)n * 8 > b.end - (b.body + 8)) return "ctts truncated";
    uint32_t s = 0;
    for (uint32_t e = 0; e < n && s < count; ++e) {
      const uint32_t cnt = r.u32(b.body + 8 + (size_t)e * 8);
      const uint32_t raw = r.u32(b.body + 12 + (size_t)e * 8);
      const int64_t off = ver == 0 ? (int64_t)raw : (int64_t)(int32_t)raw;
      if (off != 0) t->has_ctts = true;
      for (uint32_t k = 0; k < cnt && s < count; ++k) t->pts[s++] += off;
    }
  }
  // ---- edit list: leading empty edits delay the track, the first real edit's media_time is the origin
  Box edts, elst;
  if (find_box(r, trak.body, trak.end, fourcc("edts"), &edts) && find_box(r, edts.body, edts.end, fourcc("elst"), &elst) &&
      elst.end - elst.body >= 8) {
    const int ver = r.p[elst.body];
    const uint32_t n = r.u32(elst.body + 4);
    const size_t esz = ver == 1 ? 20 : 12;
    if ((size_t)n * esz <= elst.end - (elst.body + 8)) {
      int64_t empty = 0, start = 0;
      for (uint32_t e = 0; e < n; ++e) {
        const size_t o = elst.body + 8 + (size_t)e * esz;
        const int64_t dur = ver == 1 ? (int64_t)r.u64(o) : (int64_t)r.u32(o);
        const int64_t mt = ver == 1 ? (int64_t)r.u64(o + 8) : (int64_t)(int32_t)r.u32(o + 4);
        if (mt == -1) {
          empty += dur;
        } else {
          start = mt;
          break;
        }
      }
      const int64_t shift = (movie_timescale ? empty * (int64_t)t->timescale / (int64_t)movie_timescale : 0) - start;
      for (auto& p : t->pts) p += shift;
      for (auto& d : t->dts) d += shift;
    }
  }
  // ---- stss
  t->sync.assign(count, 0);
  if (find_box(r, stbl.body, stbl.end, fourcc("stss"), &b) && b.end - b.body >= 8) {
    const uint32_t n = r.u32(b.body + 4);
    if ((size_t)n * 4 > b.end - (b.body + 8)) return "stss truncated";
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t s = r.u32(b.body + 8 + (size_t)i * 4);
      if (s >= 1 && s <= count) t->sync[s - 1] = 1;
    }
  } else {
    std::fill(t->sync.begin(), t->sync.end(), 1);  // no stss: every sample is a sync sample
  }
  for (uint32_t i = 0; i < count; ++i)
    if (t->offset[i] > r.n || t->size[i] > r.n - t->offset[i]) return "sample data outside the file";
  return "";
}

}  // namespace

std::string mp4_parse(const uint8_t* data, size_t size, Mp4Track* out) {
  if (!data || size < 16) return "buffer too small";
  Reader r{data, size};
  Box moov;
  bool ok_hdr = true, have = false;
  ok_hdr = for_boxes(r, 0, size, [&](const Box& b) {
    if (b.type == fourcc("moov")) {
      moov = b;
      have = true;
      return false;
    }
    return true;
  });
  if (!ok_hdr) return "malformed top-level box";
  if (!have) return "no moov box";
  uint32_t movie_ts = 0;
  Box mvhd;
  if (find_box(r, moov.body, moov.end, fourcc("mvhd"), &mvhd) && mvhd.end - mvhd.body >= 24)
    movie_ts = r.p[mvhd.body] == 1 ? (mvhd.end - mvhd.body >= 32 ? r.u32(mvhd.body + 20) : 0) : r.u32(mvhd.body + 12);
  std::string err = "no video track";
  bool done = false;
  for_boxes(r, moov.body, moov.end, [&](const Box& b) {
    if (b.type != fourcc("trak")) return true;
    Mp4Track t;
    bool is_video = false;
    std::string e = parse_trak(r, b, movie_ts, &t, &is_video);
    if (!is_video) return true;
    if (e.empty()) {
      *out = std::move(t);
      done = true;
    } else {
      err = e;
    }
    return false;  // first video track only (stream_idx 0 in the reference)
  });
  if (done && out->size.empty()) {
    // fragmented files (ffmpeg -movflags frag_keyframe, DASH / CMAF segments) keep their samples in moof / trun boxes behind an
    // mvex declaration; the reference's clips are plain ffmpeg MP4s (clip_extraction_stages.py:318-442) - say what this is instead
    // of reporting a zero-frame video
    Box mvex;
    if (find_box(r, moov.body, moov.end, fourcc("mvex"), &mvex)) return "fragmented MP4 (mvex / moof sample tables) is not supported";
    return "video track has no samples";
  }
  return done ? "" : err;
}

namespace {

struct Out {
  std::vector<uint8_t>* v;
  void u8(uint32_t x) { v->push_back((uint8_t)x); }
  void u16(uint32_t x) { u8(x >> 8), u8(x); }
  void u32(uint32_t x) { u16(x >> 16), u16(x); }
  void u64(uint64_t x) { u32((uint32_t)(x >> 32)), u32((uint32_t)x); }
  void zeros(size_t n) { v->insert(v->end(), n, 0); }
  void bytes(const uint8_t* p, size_t n) { v->insert(v->end(), p, p + n); }
  size_t begin(const char* type, int full_version = -1, uint32_t flags = 0) {  // returns the box start for end()
    const size_t at = v->size();
    u32(0);
    bytes((const uint8_t*)type, 4);
    if (full_version >= 0) u32(((uint32_t)full_version << 24) | flags);
    return at;
  }
  void end(size_t at) {
    const uint32_t n = (uint32_t)(v->size() - at);
    (*v)[at] = n >> 24, (*v)[at + 1] = n >> 16, (*v)[at + 2] = n >> 8, (*v)[at + 3] = n;
  }
};

}  // namespace

std::string mp4_cut(const uint8_t* data, size_t size, const Mp4Track& t, size_t first, size_t count, std::vector<uint8_t>* out) {
  const size_t n = t.size.size();
  if (count == 0 || first >= n || count > n - first) return "cut range outside the track";
  if (!t.sync[first]) return "a stream-copied clip must start on a sync sample";
  if (t.stsd_size < 16 || t.stsd_off + t.stsd_size > size) return "no sample description to copy";
  std::vector<int64_t> delta(count), cts(count);
  bool any_cts = false, neg_cts = false;
  int64_t min_pts = INT64_MAX, max_end = 0;
  for (size_t i = 0; i < count; ++i) {
    const size_t s = first + i;
    delta[i] = s + 1 < n ? t.dts[s + 1] - t.dts[s] : (i > 0 ? delta[i - 1] : 1);
    if (delta[i] <= 0) return "non-increasing decode timestamps";
    cts[i] = t.pts[s] - t.dts[s];
    any_cts |= cts[i] != 0, neg_cts |= cts[i] < 0;
    const int64_t p = (t.dts[s] - t.dts[first]) + cts[i];
    min_pts = std::min(min_pts, p), max_end = std::max(max_end, p + delta[i]);
  }
  int64_t media_dur = 0;
  uint64_t payload = 0;
  for (size_t i = 0; i < count; ++i) media_dur += delta[i], payload += t.size[first + i];
  const uint32_t movie_ts = 1000;
  const uint64_t pres_dur = (uint64_t)(max_end - min_pts);
  const uint64_t movie_dur = pres_dur * movie_ts / t.timescale;

  out->clear();
  out->reserve((size_t)payload + 1024 + 16 * count);
  Out o{out};
  size_t b = o.begin("ftyp");
  o.bytes((const uint8_t*)"isom", 4), o.u32(0x200), o.bytes((const uint8_t*)"isomiso2avc1mp41", 16);
  o.end(b);
  const size_t moov = o.begin("moov");
  static const uint32_t kMatrix[9] = {0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000};
  b = o.begin("mvhd", 0);
  o.u32(0), o.u32(0), o.u32(movie_ts), o.u32((uint32_t)movie_dur), o.u32(0x10000), o.u16(0x100), o.zeros(10);
  for (uint32_t m : kMatrix) o.u32(m);
  o.zeros(24), o.u32(2);
  o.end(b);
  const size_t trak = o.begin("trak");
  b = o.begin("tkhd", 0, 3);
  o.u32(0), o.u32(0), o.u32(1), o.u32(0), o.u32((uint32_t)movie_dur), o.zeros(8), o.u16(0), o.u16(0), o.u16(0), o.u16(0);
  for (uint32_t m : kMatrix) o.u32(m);
  o.u32((uint32_t)t.width << 16), o.u32((uint32_t)t.height << 16);
  o.end(b);
  if (min_pts != 0) {  // composition offsets: presentation starts at the earliest composition time of the range
    const size_t edts = o.begin("edts");
    b = o.begin("elst", 1);
    o.u32(1), o.u64(movie_dur), o.u64((uint64_t)min_pts), o.u16(1), o.u16(0);
    o.end(b), o.end(edts);
  }
  const size_t mdia = o.begin("mdia");
  b = o.begin("mdhd", 1);
  o.u64(0), o.u64(0), o.u32(t.timescale), o.u64((uint64_t)media_dur), o.u16(0x55C4), o.u16(0);
  o.end(b);
  b = o.begin("hdlr", 0);
  o.u32(0), o.bytes((const uint8_t*)"vide", 4), o.zeros(12), o.bytes((const uint8_t*)"VideoHandler", 13);
  o.end(b);
  const size_t minf = o.begin("minf");
  b = o.begin("vmhd", 0, 1);
  o.zeros(8);
  o.end(b);
  const size_t dinf = o.begin("dinf");
  b = o.begin("dref", 0);
  o.u32(1);
  const size_t url = o.begin("url ", 0, 1);
  o.end(url), o.end(b), o.end(dinf);
  const size_t stbl = o.begin("stbl");
  o.bytes(data + t.stsd_off, t.stsd_size);
  b = o.begin("stts", 0);
  {
    const size_t cnt_at = out->size();
    o.u32(0);
    uint32_t entries = 0;
    for (size_t i = 0; i < count;) {
      size_t j = i;
      while (j < count && delta[j] == delta[i]) ++j;
      o.u32((uint32_t)(j - i)), o.u32((uint32_t)delta[i]);
      ++entries, i = j;
    }
    (*out)[cnt_at] = entries >> 24, (*out)[cnt_at + 1] = entries >> 16, (*out)[cnt_at + 2] = entries >> 8, (*out)[cnt_at + 3] = entries;
  }
  o.end(b);
  if (any_cts) {
    b = o.begin("ctts", neg_cts ? 1 : 0);
    const size_t cnt_at = out->size();
    o.u32(0);
    uint32_t entries = 0;
    for (size_t i = 0; i < count;) {
      size_t j = i;
      while (j < count && cts[j] == cts[i]) ++j;
      o.u32((uint32_t)(j - i)), o.u32((uint32_t)(int32_t)cts[i]);
      ++entries, i = j;
    }
    (*out)[cnt_at] = entries >> 24, (*out)[cnt_at + 1] = entries >> 16, (*out)[cnt_at + 2] = entries >> 8, (*out)[cnt_at + 3] = entries;
    o.end(b);
  }
  b = o.begin("stss", 0);
  {
    uint32_t ns = 0;
    for (size_t i = 0; i < count; ++i) ns += t.sync[first + i];
    o.u32(ns);
    for (size_t i = 0; i < count; ++i)
      if (t.sync[first + i]) o.u32((uint32_t)i + 1);
  }
  o.end(b);
  b = o.begin("stsc", 0);
  o.u32(1), o.u32(1), o.u32((uint32_t)count), o.u32(1);
  o.end(b);
  b = o.begin("stsz", 0);
  o.u32(0), o.u32((uint32_t)count);
  for (size_t i = 0; i < count; ++i) o.u32(t.size[first + i]);
  o.end(b);
  b = o.begin("co64", 0);
  o.u32(1);
  const size_t co_at = out->size();
  o.u64(0);
  o.end(b);
  o.end(stbl), o.end(minf), o.end(mdia), o.end(trak), o.end(moov);
  const uint64_t mdat_at = out->size();
  o.u32(1), o.bytes((const uint8_t*)"mdat", 4), o.u64(16 + payload);
  const uint64_t chunk = mdat_at + 16;
  for (int k = 0; k < 8; ++k) (*out)[co_at + k] = (uint8_t)(chunk >> (56 - 8 * k));
  for (size_t i = 0; i < count; ++i) o.bytes(data + t.offset[first + i], t.size[first + i]);
  return "";
}

bool mp4_sample_annexb(const uint8_t* data, size_t size, const Mp4Track& t, size_t i, std::vector<uint8_t>* dst) {
  if (i >= t.size.size()) return false;
  size_t o = (size_t)t.offset[i];
  const size_t end = o + t.size[i];
  if (end > size) return false;
  const int L = t.nal_length_size;
  while (o + L <= end) {
    size_t len = 0;
    for (int k = 0; k < L; ++k) len = len << 8 | data[o + k];
    o += L;
    if (len > end - o) return false;
    append_start_code(dst);
    dst->insert(dst->end(), data + o, data + o + len);
    o += len;
  }
  return o == end;
}

}  // namespace cb
