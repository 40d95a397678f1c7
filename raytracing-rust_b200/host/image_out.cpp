// Host-side image output for ptb200: the consumer of the read-back accumulator.
// Follows crates/output/src/lib.rs:74-113 (save_data_to_image): the filename must be `<stem>.<ext>`;
// 8-bit formats store (v^(1/gamma) * 255.999) as u8 with Rust's saturating float->int cast; the float format
// stores linear radiance and ignores gamma. Every extension the reference accepts is written here without third-party
// code: png (stored deflate), jpg / jpeg (baseline DCT, quality 75 = the `image` crate's default, 4:4:4), tiff (baseline,
// uncompressed RGB strip), ppm, bmp, exr (OpenEXR 2 scanline file, three FLOAT channels, no compression — linear f32 like
// the reference's Rgb32FImage), plus pfm as a second linear format.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ptb200.h"

namespace {

inline uint8_t to_u8(float val, float gamma) {
  float v = std::pow(val, 1.0f / gamma) * 255.999f;
  if (!(v > 0.0f)) return 0;  // NaN and negatives saturate to 0 (Rust `as u8`)
  if (v >= 255.0f) return 255;
  return (uint8_t)v;
}

uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
  return crc;
}

void put_be32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

void png_chunk(std::vector<uint8_t>& out, const char type[4], const std::vector<uint8_t>& data) {
  put_be32(out, (uint32_t)data.size());
  size_t start = out.size();
  out.insert(out.end(), type, type + 4);
  out.insert(out.end(), data.begin(), data.end());
  uint32_t crc = crc32_update(0xFFFFFFFFu, out.data() + start, out.size() - start) ^ 0xFFFFFFFFu;
  put_be32(out, crc);
}

bool write_file(const char* name, const std::vector<uint8_t>& bytes) {
  FILE* f = std::fopen(name, "wb");
  if (!f) return false;
  bool ok = std::fwrite(bytes.data(), 1, bytes.size(), f) == bytes.size();
  return std::fclose(f) == 0 && ok;
}

// ---------------------------------------------------------------------------------------------- exr
// OpenEXR 2.0 single-part scanline file ("OpenEXR File Layout"): magic, version, attribute list, line offset table, one
// chunk per scanline holding the channels in alphabetical order (B, G, R), each `width` little-endian floats.
void put_le32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }
void put_le64(std::vector<uint8_t>& v, uint64_t x) { for (int i = 0; i < 8; ++i) v.push_back((uint8_t)(x >> (8 * i))); }
void put_f32(std::vector<uint8_t>& v, float f) { uint32_t u; std::memcpy(&u, &f, 4); put_le32(v, u); }
void put_str(std::vector<uint8_t>& v, const char* s) { v.insert(v.end(), s, s + std::strlen(s) + 1); }
void exr_attr(std::vector<uint8_t>& v, const char* name, const char* type, const std::vector<uint8_t>& value) {
  put_str(v, name); put_str(v, type); put_le32(v, (uint32_t)value.size());
  v.insert(v.end(), value.begin(), value.end());
}
std::vector<uint8_t> encode_exr(uint32_t w, uint32_t h, const float* rgb) {
  std::vector<uint8_t> out;
  put_le32(out, 20000630u);  // 0x76 0x2f 0x31 0x01
  put_le32(out, 2u);         // version 2, no flags: single-part scanline
  std::vector<uint8_t> a;
  for (const char* ch : {"B", "G", "R"}) {
    put_str(a, ch);
    put_le32(a, 2u);  // FLOAT
    a.push_back(0); a.push_back(0); a.push_back(0); a.push_back(0);  // pLinear + reserved
    put_le32(a, 1u); put_le32(a, 1u);                                // x / y sampling
  }
  a.push_back(0);
  exr_attr(out, "channels", "chlist", a);
  exr_attr(out, "compression", "compression", std::vector<uint8_t>{0});
  a.clear(); put_le32(a, 0); put_le32(a, 0); put_le32(a, w - 1); put_le32(a, h - 1);
  exr_attr(out, "dataWindow", "box2i", a);
  exr_attr(out, "displayWindow", "box2i", a);
  exr_attr(out, "lineOrder", "lineOrder", std::vector<uint8_t>{0});  // increasing y: top row first, like the image buffer
  a.clear(); put_f32(a, 1.0f);
  exr_attr(out, "pixelAspectRatio", "float", a);
  a.clear(); put_f32(a, 0.0f); put_f32(a, 0.0f);
  exr_attr(out, "screenWindowCenter", "v2f", a);
  a.clear(); put_f32(a, 1.0f);
  exr_attr(out, "screenWindowWidth", "float", a);
  out.push_back(0);  // end of header
  const uint64_t chunk = 8ull + 12ull * w, table = out.size() + 8ull * h;
  for (uint32_t y = 0; y < h; ++y) put_le64(out, table + chunk * y);
  out.reserve(out.size() + chunk * h);
  for (uint32_t y = 0; y < h; ++y) {
    put_le32(out, y);
    put_le32(out, 12u * w);
    for (int c = 2; c >= 0; --c)  // B, G, R
      for (uint32_t x = 0; x < w; ++x) put_f32(out, rgb[((size_t)y * w + x) * 3 + c]);
  }
  return out;
}

// ---------------------------------------------------------------------------------------------- tiff
// TIFF 6.0 baseline RGB, one uncompressed strip, little endian.
std::vector<uint8_t> encode_tiff(uint32_t w, uint32_t h, const std::vector<uint8_t>& px) {
  std::vector<uint8_t> out;
  auto put16 = [&](uint16_t x) { out.push_back((uint8_t)x); out.push_back((uint8_t)(x >> 8)); };
  out.push_back('I'); out.push_back('I'); put16(42);
  const uint32_t data_off = 8, ifd_off = data_off + (uint32_t)px.size() + ((uint32_t)px.size() & 1u);
  put_le32(out, ifd_off);
  out.insert(out.end(), px.begin(), px.end());
  if (px.size() & 1u) out.push_back(0);
  const uint16_t n_tags = 11;
  const uint32_t bits_off = ifd_off + 2 + 12 * n_tags + 4;
  auto tag = [&](uint16_t id, uint16_t type, uint32_t count, uint32_t value) {
    put16(id); put16(type); put_le32(out, count); put_le32(out, value);
  };
  put16(n_tags);
  tag(256, 4, 1, w);                       // ImageWidth
  tag(257, 4, 1, h);                       // ImageLength
  tag(258, 3, 3, bits_off);                // BitsPerSample -> 8,8,8
  tag(259, 3, 1, 1);                       // Compression: none
  tag(262, 3, 1, 2);                       // PhotometricInterpretation: RGB
  tag(273, 4, 1, data_off);                // StripOffsets
  tag(277, 3, 1, 3);                       // SamplesPerPixel
  tag(278, 4, 1, h);                       // RowsPerStrip
  tag(279, 4, 1, (uint32_t)px.size());     // StripByteCounts
  tag(284, 3, 1, 1);                       // PlanarConfiguration: chunky
  tag(296, 3, 1, 1);                       // ResolutionUnit: none
  put_le32(out, 0);                        // no further IFD
  put16(8); put16(8); put16(8);
  return out;
}

// ---------------------------------------------------------------------------------------------- jpeg
// Baseline sequential DCT (ITU-T T.81), YCbCr 4:4:4, the Annex K quantisation and Huffman tables, libjpeg quality scaling.
struct BitWriter {
  std::vector<uint8_t>& out;
  uint32_t acc = 0;
  int nbits = 0;
  void put(uint32_t code, int len) {
    acc = (acc << len) | (code & ((1u << len) - 1u));
    nbits += len;
    while (nbits >= 8) {
      const uint8_t b = (uint8_t)(acc >> (nbits - 8));
      out.push_back(b);
      if (b == 0xFF) out.push_back(0);  // byte stuffing
      nbits -= 8;
    }
  }
  void flush() { if (nbits) put(0x7Fu, 8 - nbits); }
};
struct Huff { uint16_t code[256]; uint8_t len[256]; };
void build_huff(const uint8_t bits[16], const uint8_t* vals, Huff& h) {
  std::memset(&h, 0, sizeof h);
  uint32_t code = 0;
  int k = 0;
  for (int l = 1; l <= 16; ++l) {
    for (int i = 0; i < bits[l - 1]; ++i, ++k) { h.code[vals[k]] = (uint16_t)code++; h.len[vals[k]] = (uint8_t)l; }
    code <<= 1;
  }
}
const uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t kQLum[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                           18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kQChr[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                           99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
const uint8_t kDcLumBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChrBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
    0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChrBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChrVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
    0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

void fdct8x8(float* b) {  // separable orthonormal DCT-II
  static float c[8][8];
  static bool init = false;
  if (!init) {
    for (int u = 0; u < 8; ++u)
      for (int x = 0; x < 8; ++x) c[u][x] = (u == 0 ? std::sqrt(0.125f) : 0.5f) * std::cos((2 * x + 1) * u * 3.14159265358979323846f / 16.0f);
    init = true;
  }
  float t[64];
  for (int y = 0; y < 8; ++y)
    for (int u = 0; u < 8; ++u) {
      float s = 0.0f;
      for (int x = 0; x < 8; ++x) s += c[u][x] * b[y * 8 + x];
      t[y * 8 + u] = s;
    }
  for (int u = 0; u < 8; ++u)
    for (int v = 0; v < 8; ++v) {
      float s = 0.0f;
      for (int y = 0; y < 8; ++y) s += c[v][y] * t[y * 8 + u];
      b[v * 8 + u] = s;
    }
}
std::vector<uint8_t> encode_jpeg(uint32_t w, uint32_t h, const std::vector<uint8_t>& px, int quality) {
  std::vector<uint8_t> out;
  auto put16 = [&](uint16_t x) { out.push_back((uint8_t)(x >> 8)); out.push_back((uint8_t)x); };
  auto marker = [&](uint8_t m, uint16_t len) { out.push_back(0xFF); out.push_back(m); put16(len); };
  const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
  uint8_t q[2][64];
  for (int i = 0; i < 64; ++i) {
    int a = (kQLum[i] * scale + 50) / 100, b = (kQChr[i] * scale + 50) / 100;
    q[0][i] = (uint8_t)(a < 1 ? 1 : a > 255 ? 255 : a);
    q[1][i] = (uint8_t)(b < 1 ? 1 : b > 255 ? 255 : b);
  }
  out.push_back(0xFF); out.push_back(0xD8);  // SOI
  marker(0xE0, 16);
  const uint8_t jfif[14] = {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
  out.insert(out.end(), jfif, jfif + 14);
  for (int t = 0; t < 2; ++t) {  // DQT, zigzag order
    marker(0xDB, 67);
    out.push_back((uint8_t)t);
    for (int i = 0; i < 64; ++i) out.push_back(q[t][kZigzag[i]]);
  }
  marker(0xC0, 17);  // SOF0
  out.push_back(8); put16((uint16_t)h); put16((uint16_t)w); out.push_back(3);
  for (int cidx = 0; cidx < 3; ++cidx) { out.push_back((uint8_t)(cidx + 1)); out.push_back(0x11); out.push_back(cidx ? 1 : 0); }
  struct Tab { uint8_t id; const uint8_t* bits; const uint8_t* vals; int n; };
  const Tab tabs[4] = {{0x00, kDcLumBits, kDcVals, 12}, {0x10, kAcLumBits, kAcLumVals, 162}, {0x01, kDcChrBits, kDcVals, 12}, {0x11, kAcChrBits, kAcChrVals, 162}};
  for (const Tab& t : tabs) {  // DHT
    marker(0xC4, (uint16_t)(19 + t.n));
    out.push_back(t.id);
    out.insert(out.end(), t.bits, t.bits + 16);
    out.insert(out.end(), t.vals, t.vals + t.n);
  }
  marker(0xDA, 12);  // SOS
  out.push_back(3);
  out.push_back(1); out.push_back(0x00); out.push_back(2); out.push_back(0x11); out.push_back(3); out.push_back(0x11);
  out.push_back(0); out.push_back(63); out.push_back(0);
  Huff dc[2], ac[2];
  build_huff(kDcLumBits, kDcVals, dc[0]); build_huff(kAcLumBits, kAcLumVals, ac[0]);
  build_huff(kDcChrBits, kDcVals, dc[1]); build_huff(kAcChrBits, kAcChrVals, ac[1]);
  BitWriter bw{out};
  int pred[3] = {0, 0, 0};
  auto category = [](int v) { int a = v < 0 ? -v : v, n = 0; while (a) { ++n; a >>= 1; } return n; };
  for (uint32_t by = 0; by < h; by += 8)
    for (uint32_t bx = 0; bx < w; bx += 8) {
      float blk[3][64];
      for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) {
          const uint32_t sx = bx + x < w ? bx + x : w - 1, sy = by + y < h ? by + y : h - 1;  // edge replication
          const uint8_t* p = &px[((size_t)sy * w + sx) * 3];
          const float r = p[0], g = p[1], b = p[2];
          blk[0][y * 8 + x] = 0.299f * r + 0.587f * g + 0.114f * b - 128.0f;
          blk[1][y * 8 + x] = -0.168736f * r - 0.331264f * g + 0.5f * b;
          blk[2][y * 8 + x] = 0.5f * r - 0.418688f * g - 0.081312f * b;
        }
      for (int cidx = 0; cidx < 3; ++cidx) {
        const int t = cidx ? 1 : 0;
        fdct8x8(blk[cidx]);
        int zz[64];
        for (int i = 0; i < 64; ++i) zz[i] = (int)std::lrintf(blk[cidx][kZigzag[i]] / (float)q[t][kZigzag[i]]);
        const int diff = zz[0] - pred[cidx];
        pred[cidx] = zz[0];
        int cat = category(diff);
        bw.put(dc[t].code[cat], dc[t].len[cat]);
        if (cat) bw.put((uint32_t)(diff < 0 ? diff - 1 : diff), cat);
        int run = 0;
        for (int i = 1; i < 64; ++i) {
          if (zz[i] == 0) { ++run; continue; }
          while (run > 15) { bw.put(ac[t].code[0xF0], ac[t].len[0xF0]); run -= 16; }
          cat = category(zz[i]);
          const int sym = (run << 4) | cat;
          bw.put(ac[t].code[sym], ac[t].len[sym]);
          bw.put((uint32_t)(zz[i] < 0 ? zz[i] - 1 : zz[i]), cat);
          run = 0;
        }
        if (run) bw.put(ac[t].code[0x00], ac[t].len[0x00]);  // EOB
      }
    }
  bw.flush();
  out.push_back(0xFF); out.push_back(0xD9);  // EOI
  return out;
}

}  // namespace

extern "C" int32_t ptb_image_save(const char* filename, uint32_t width, uint32_t height, const float* rgb, float gamma) {
  if (!filename || !rgb || width == 0 || height == 0) return PTB_ERR_INVALID;
  std::string name(filename);
  // output/lib.rs:81-85: exactly one '.' in the whole filename
  size_t parts = 1;
  for (char c : name) parts += c == '.';
  if (parts != 2) return PTB_ERR_INVALID;
  std::string ext = name.substr(name.find('.') + 1);
  const size_t n = (size_t)width * height * 3;

  if (ext == "exr") return write_file(filename, encode_exr(width, height, rgb)) ? PTB_OK : PTB_ERR_IO;
  if (ext == "pfm") {
    std::vector<uint8_t> out;
    char hdr[64];
    int hl = std::snprintf(hdr, sizeof hdr, "PF\n%u %u\n-1.0\n", width, height);
    out.insert(out.end(), hdr, hdr + hl);
    for (uint32_t y = height; y-- > 0;) {  // PFM rows go bottom to top
      const uint8_t* row = reinterpret_cast<const uint8_t*>(rgb + (size_t)y * width * 3);
      out.insert(out.end(), row, row + (size_t)width * 12);
    }
    return write_file(filename, out) ? PTB_OK : PTB_ERR_IO;
  }

  std::vector<uint8_t> px(n);
  for (size_t i = 0; i < n; ++i) px[i] = to_u8(rgb[i], gamma);

  if (ext == "ppm") {
    std::vector<uint8_t> out;
    char hdr[64];
    int hl = std::snprintf(hdr, sizeof hdr, "P6\n%u %u\n255\n", width, height);
    out.insert(out.end(), hdr, hdr + hl);
    out.insert(out.end(), px.begin(), px.end());
    return write_file(filename, out) ? PTB_OK : PTB_ERR_IO;
  }
  if (ext == "bmp") {
    const uint32_t row = (width * 3 + 3) & ~3u;
    const uint32_t size = 54 + row * height;
    std::vector<uint8_t> out(size, 0);
    out[0] = 'B'; out[1] = 'M';
    std::memcpy(&out[2], &size, 4);
    uint32_t off = 54, ih = 40, planes_bpp = 1 | (24u << 16), img = row * height;
    std::memcpy(&out[10], &off, 4);
    std::memcpy(&out[14], &ih, 4);
    std::memcpy(&out[18], &width, 4);
    std::memcpy(&out[22], &height, 4);
    std::memcpy(&out[26], &planes_bpp, 4);
    std::memcpy(&out[34], &img, 4);
    for (uint32_t y = 0; y < height; ++y) {
      uint8_t* dst = &out[54 + (size_t)(height - 1 - y) * row];
      const uint8_t* src = &px[(size_t)y * width * 3];
      for (uint32_t x = 0; x < width; ++x) { dst[3 * x] = src[3 * x + 2]; dst[3 * x + 1] = src[3 * x + 1]; dst[3 * x + 2] = src[3 * x]; }
    }
    return write_file(filename, out) ? PTB_OK : PTB_ERR_IO;
  }
  if (ext == "png") {
    std::vector<uint8_t> raw;
    raw.reserve((size_t)height * (width * 3 + 1));
    for (uint32_t y = 0; y < height; ++y) {
      raw.push_back(0);  // filter: none
      raw.insert(raw.end(), px.begin() + (size_t)y * width * 3, px.begin() + (size_t)(y + 1) * width * 3);
    }
    std::vector<uint8_t> z;
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    size_t pos = 0;
    while (pos < raw.size() || raw.empty()) {
      size_t len = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
      bool last = pos + len >= raw.size();
      z.push_back(last ? 1 : 0);
      z.push_back((uint8_t)(len & 0xFF)); z.push_back((uint8_t)(len >> 8));
      z.push_back((uint8_t)(~len & 0xFF)); z.push_back((uint8_t)((~len >> 8) & 0xFF));
      z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + len);
      for (size_t i = pos; i < pos + len; ++i) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
      pos += len;
      if (last) break;
    }
    put_be32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width); put_be32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", z);
    png_chunk(out, "IEND", std::vector<uint8_t>());
    return write_file(filename, out) ? PTB_OK : PTB_ERR_IO;
  }
  if (ext == "tiff") return write_file(filename, encode_tiff(width, height, px)) ? PTB_OK : PTB_ERR_IO;
  if (ext == "jpg" || ext == "jpeg") {
    if (width > 65535u || height > 65535u) return PTB_ERR_INVALID;  // JPEG's 16-bit frame header
    return write_file(filename, encode_jpeg(width, height, px, 75)) ? PTB_OK : PTB_ERR_IO;
  }
  return PTB_ERR_UNSUPPORTED;  // output/lib.rs:106-109: "unknown filetype"
}
