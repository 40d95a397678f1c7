// Host-side image output for ptb200: the consumer of the read-back accumulator.
// Follows crates/output/src/lib.rs:74-113 (save_data_to_image): the filename must be `<stem>.<ext>`;
// 8-bit formats store (v^(1/gamma) * 255.999) as u8 with Rust's saturating float->int cast; the float format
// stores linear radiance and ignores gamma. Encoders available here without third-party crates: ppm, bmp,
// png (stored/uncompressed deflate) and pfm (linear f32, standing in for the reference's exr).
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

  if (ext == "pfm" || ext == "exr") {
    if (ext == "exr") return PTB_ERR_UNSUPPORTED;  // no OpenEXR encoder in this image; use .pfm for linear f32
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
  // jpg / jpeg / tiff are accepted by the reference through the `image` crate; no encoder here
  return PTB_ERR_UNSUPPORTED;
}
