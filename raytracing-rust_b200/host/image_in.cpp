// ptb200 host side: image decoding for ImageTexture and the Perlin tables.
//   ptb_image_load    <- ImageTexture::new (implementations/src/textures/mod.rs:208-245): `image::open(path)` then
//                        `to_rgb32f()` (8-bit -> v/255, 16-bit -> v/65535, grey replicated, alpha dropped).
//                        Formats decoded here without external libraries: ppm/pgm (P2 P3 P5 P6), pfm (PF Pf), bmp
//                        (uncompressed 24/32 bit), png (non-interlaced, 8/16 bit, all colour types; own inflate).
//   ptb_perlin_tables <- Perlin::new / generate_perm / permute (textures/mod.rs:90-112, 141-159).
// No GPU code; nothing here is on the hot path.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ptb200.h"

namespace {

bool read_file(const char* path, std::vector<uint8_t>& out) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (n < 0) { std::fclose(f); return false; }
  out.resize((size_t)n);
  size_t got = n ? std::fread(out.data(), 1, (size_t)n, f) : 0;
  std::fclose(f);
  return got == (size_t)n;
}

// ---------------------------------------------------------------------------------------------- inflate (RFC 1951)
struct BitReader {
  const uint8_t* p;
  size_t n, pos = 0;
  uint32_t bitbuf = 0;
  int bitcnt = 0;
  bool bad = false;
  uint32_t bits(int need) {
    while (bitcnt < need) {
      if (pos >= n) { bad = true; return 0; }
      bitbuf |= (uint32_t)p[pos++] << bitcnt;
      bitcnt += 8;
    }
    uint32_t v = bitbuf & ((need == 32) ? 0xFFFFFFFFu : ((1u << need) - 1u));
    bitbuf >>= need;
    bitcnt -= need;
    return v;
  }
};
struct Huffman {
  uint16_t count[16];
  uint16_t symbol[288];
  void build(const uint8_t* lengths, int n) {
    std::memset(count, 0, sizeof(count));
    for (int i = 0; i < n; ++i) count[lengths[i]]++;
    count[0] = 0;
    uint16_t offs[16];
    offs[1] = 0;
    for (int len = 1; len < 15; ++len) offs[len + 1] = offs[len] + count[len];
    for (int i = 0; i < n; ++i)
      if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
  }
  int decode(BitReader& br) const {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
      code |= (int)br.bits(1);
      if (br.bad) return -1;
      int c = count[len];
      if (code - c < first) return symbol[index + (code - first)];
      index += c;
      first += c;
      first <<= 1;
      code <<= 1;
    }
    return -1;
  }
};
bool inflate_raw(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
  static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint16_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint16_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  BitReader br{src, n};
  for (;;) {
    uint32_t last = br.bits(1), type = br.bits(2);
    if (br.bad) return false;
    if (type == 0) {
      br.bitbuf = 0;
      br.bitcnt = 0;
      if (br.pos + 4 > n) return false;
      uint32_t len = src[br.pos] | (src[br.pos + 1] << 8);
      br.pos += 4;
      if (br.pos + len > n) return false;
      out.insert(out.end(), src + br.pos, src + br.pos + len);
      br.pos += len;
    } else if (type == 1 || type == 2) {
      Huffman lit, dist;
      uint8_t lengths[320];
      if (type == 1) {
        int i = 0;
        for (; i < 144; ++i) lengths[i] = 8;
        for (; i < 256; ++i) lengths[i] = 9;
        for (; i < 280; ++i) lengths[i] = 7;
        for (; i < 288; ++i) lengths[i] = 8;
        lit.build(lengths, 288);
        for (i = 0; i < 30; ++i) lengths[i] = 5;
        dist.build(lengths, 30);
      } else {
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
        if (br.bad || nlen > 286 || ndist > 30) return false;
        uint8_t cl[19] = {0};
        for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)br.bits(3);
        Huffman lencode;
        lencode.build(cl, 19);
        int idx = 0;
        while (idx < nlen + ndist) {
          int sym = lencode.decode(br);
          if (sym < 0) return false;
          if (sym < 16) lengths[idx++] = (uint8_t)sym;
          else {
            int rep, val = 0;
            if (sym == 16) { if (idx == 0) return false; val = lengths[idx - 1]; rep = 3 + (int)br.bits(2); }
            else if (sym == 17) rep = 3 + (int)br.bits(3);
            else rep = 11 + (int)br.bits(7);
            if (idx + rep > nlen + ndist) return false;
            while (rep--) lengths[idx++] = (uint8_t)val;
          }
        }
        lit.build(lengths, nlen);
        dist.build(lengths + nlen, ndist);
      }
      for (;;) {
        int sym = lit.decode(br);
        if (sym < 0 || br.bad) return false;
        if (sym < 256) out.push_back((uint8_t)sym);
        else if (sym == 256) break;
        else {
          sym -= 257;
          if (sym >= 29) return false;
          size_t len = lbase[sym] + br.bits(lext[sym]);
          int ds = dist.decode(br);
          if (ds < 0 || ds >= 30) return false;
          size_t d = dbase[ds] + br.bits(dext[ds]);
          if (d > out.size()) return false;
          size_t from = out.size() - d;
          for (size_t i = 0; i < len; ++i) out.push_back(out[from + i]);
        }
      }
    } else {
      return false;
    }
    if (last) break;
  }
  return !br.bad;
}

// Header dimensions are untrusted: 65535 per side keeps every size product below 2^32 * 8 (no size_t wrap on 64-bit hosts)
// and bounds the allocation a hostile header can request.
bool dims_ok(uint32_t w, uint32_t h) { return w <= 65535u && h <= 65535u; }

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// ---------------------------------------------------------------------------------------------- png
bool decode_png(const std::vector<uint8_t>& f, uint32_t& w, uint32_t& h, std::vector<float>& rgb, std::string& err) {
  size_t pos = 8;
  uint32_t depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte;
  bool have_ihdr = false;
  while (pos + 12 <= f.size()) {
    uint32_t len = be32(&f[pos]);
    const uint8_t* type = &f[pos + 4];
    if (pos + 12 + (size_t)len > f.size()) { err = "png: truncated chunk"; return false; }
    const uint8_t* data = &f[pos + 8];
    if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
      w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
      have_ihdr = true;
    } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
    else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
    else if (!std::memcmp(type, "IEND", 4)) break;
    pos += 12 + (size_t)len;
  }
  if (!have_ihdr || w == 0 || h == 0) { err = "png: missing IHDR"; return false; }
  if (!dims_ok(w, h)) { err = "png: image dimensions above the supported 65535 x 65535"; return false; }
  if (interlace) { err = "png: interlaced files are not supported"; return false; }
  int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  // PNG specification table 11.1: greyscale 1/2/4/8/16, palette 1/2/4/8, every other colour type 8/16
  const bool depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                        : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                     : (depth == 8 || depth == 16);
  if (!channels || !depth_ok) { err = "png: bad colour type / bit depth combination"; return false; }
  if (idat.size() < 6) { err = "png: no image data"; return false; }
  std::vector<uint8_t> raw;
  if (!inflate_raw(idat.data() + 2, idat.size() - 2, raw)) { err = "png: corrupt deflate stream"; return false; }
  const size_t bpp_bits = (size_t)channels * depth, stride = ((size_t)w * bpp_bits + 7) / 8, bpp = (bpp_bits + 7) / 8;
  if (raw.size() < (stride + 1) * h) { err = "png: image data too short"; return false; }
  std::vector<uint8_t> img(stride * h);
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t* in = &raw[(stride + 1) * y];
    uint8_t* cur = &img[stride * y];
    const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
    const uint8_t ft = in[0];
    ++in;
    for (size_t i = 0; i < stride; ++i) {
      int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0, pr = 0;
      switch (ft) {
        case 0: pr = 0; break;
        case 1: pr = a; break;
        case 2: pr = b; break;
        case 3: pr = (a + b) >> 1; break;
        case 4: { int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                  pr = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
        default: err = "png: bad filter"; return false;
      }
      cur[i] = (uint8_t)(in[i] + pr);
    }
  }
  rgb.resize((size_t)w * h * 3);
  const float maxv = (float)((1u << depth) - 1u);
  auto sample = [&](const uint8_t* row, size_t idx) -> uint32_t {  // idx-th sample of `depth` bits in the row
    if (depth == 8) return row[idx];
    if (depth == 16) return ((uint32_t)row[2 * idx] << 8) | row[2 * idx + 1];
    size_t bit = idx * depth;
    return (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
  };
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t* row = &img[stride * y];
    for (uint32_t x = 0; x < w; ++x) {
      float* o = &rgb[((size_t)y * w + x) * 3];
      if (ctype == 3) {
        uint32_t pi = sample(row, x);
        if ((size_t)pi * 3 + 2 >= plte.size()) { o[0] = o[1] = o[2] = 0.0f; continue; }
        for (int k = 0; k < 3; ++k) o[k] = (float)plte[pi * 3 + k] / 255.0f;
      } else if (ctype == 0 || ctype == 4) {
        float v = (float)sample(row, (size_t)x * channels) / maxv;
        o[0] = o[1] = o[2] = v;
      } else {
        for (int k = 0; k < 3; ++k) o[k] = (float)sample(row, (size_t)x * channels + k) / maxv;
      }
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------- pnm / pfm
struct Tok {
  const std::vector<uint8_t>& f;
  size_t pos;
  bool next(std::string& t) {
    for (;;) {
      while (pos < f.size() && std::isspace(f[pos])) ++pos;
      if (pos < f.size() && f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') ++pos; continue; }
      break;
    }
    t.clear();
    while (pos < f.size() && !std::isspace(f[pos])) t.push_back((char)f[pos++]);
    return !t.empty();
  }
};
bool decode_pnm(const std::vector<uint8_t>& f, uint32_t& w, uint32_t& h, std::vector<float>& rgb, std::string& err) {
  Tok tk{f, 0};
  std::string magic, t;
  if (!tk.next(magic)) { err = "pnm: empty"; return false; }
  const bool pfm = magic == "PF" || magic == "Pf";
  const int ch = (magic == "P3" || magic == "P6" || magic == "PF") ? 3 : 1;
  if (!tk.next(t)) { err = "pnm: truncated header"; return false; }
  w = (uint32_t)std::strtoul(t.c_str(), nullptr, 10);
  if (!tk.next(t)) { err = "pnm: truncated header"; return false; }
  h = (uint32_t)std::strtoul(t.c_str(), nullptr, 10);
  if (!tk.next(t)) { err = "pnm: truncated header"; return false; }
  if (w == 0 || h == 0) { err = "pnm: zero dimension"; return false; }
  if (!dims_ok(w, h)) { err = "pnm: image dimensions above the supported 65535 x 65535"; return false; }
  rgb.resize((size_t)w * h * 3);
  if (pfm) {
    const double scale = std::strtod(t.c_str(), nullptr);
    const bool little = scale < 0.0;
    size_t pos = tk.pos + 1;
    if (pos + (size_t)w * h * ch * 4 > f.size()) { err = "pfm: truncated data"; return false; }
    for (uint32_t y = 0; y < h; ++y)      // pfm rows run bottom to top
      for (uint32_t x = 0; x < w; ++x)
        for (int k = 0; k < 3; ++k) {
          const uint8_t* p = &f[pos + (((size_t)(h - 1 - y) * w + x) * ch + (ch == 3 ? k : 0)) * 4];
          uint8_t b[4];
          for (int i = 0; i < 4; ++i) b[i] = little ? p[i] : p[3 - i];
          float v;
          std::memcpy(&v, b, 4);
          rgb[((size_t)y * w + x) * 3 + k] = v;
        }
    return true;
  }
  const uint32_t maxv = (uint32_t)std::strtoul(t.c_str(), nullptr, 10);
  if (maxv == 0 || maxv > 65535) { err = "pnm: bad maxval"; return false; }
  const size_t n = (size_t)w * h * ch;
  std::vector<uint32_t> v(n);
  if (magic == "P2" || magic == "P3") {
    for (size_t i = 0; i < n; ++i) {
      if (!tk.next(t)) { err = "pnm: truncated data"; return false; }
      v[i] = (uint32_t)std::strtoul(t.c_str(), nullptr, 10);
    }
  } else if (magic == "P5" || magic == "P6") {
    size_t pos = tk.pos + 1;
    const size_t bytes = maxv > 255 ? 2 : 1;
    if (pos + n * bytes > f.size()) { err = "pnm: truncated data"; return false; }
    for (size_t i = 0; i < n; ++i) v[i] = bytes == 2 ? ((uint32_t)f[pos + 2 * i] << 8) | f[pos + 2 * i + 1] : f[pos + i];
  } else {
    err = "pnm: unsupported magic " + magic;
    return false;
  }
  // the `image` crate decodes pnm samples to 8 or 16 bit and to_rgb32f divides by the type's maximum
  const float denom = maxv > 255 ? 65535.0f : 255.0f;
  const float rescale = maxv > 255 ? 65535.0f / (float)maxv : 255.0f / (float)maxv;
  for (size_t i = 0; i < (size_t)w * h; ++i)
    for (int k = 0; k < 3; ++k) {
      const uint32_t s = v[i * ch + (ch == 3 ? k : 0)];
      const float q = (maxv == 255 || maxv == 65535) ? (float)s : std::floor((float)s * rescale + 0.5f);
      rgb[i * 3 + k] = q / denom;
    }
  return true;
}

// ---------------------------------------------------------------------------------------------- bmp
bool decode_bmp(const std::vector<uint8_t>& f, uint32_t& w, uint32_t& h, std::vector<float>& rgb, std::string& err) {
  if (f.size() < 54) { err = "bmp: truncated header"; return false; }
  auto le32 = [&](size_t o) { return (uint32_t)f[o] | ((uint32_t)f[o + 1] << 8) | ((uint32_t)f[o + 2] << 16) | ((uint32_t)f[o + 3] << 24); };
  const uint32_t off = le32(10);
  const int32_t wi = (int32_t)le32(18), hi = (int32_t)le32(22);
  const uint32_t bpp = f[28] | (f[29] << 8), comp = le32(30);
  if (wi <= 0 || hi == 0 || (bpp != 24 && bpp != 32) || (comp != 0 && comp != 3)) { err = "bmp: only uncompressed 24/32-bit files are supported"; return false; }
  w = (uint32_t)wi;
  h = hi < 0 ? (uint32_t)(-(int64_t)hi) : (uint32_t)hi;
  if (!dims_ok(w, h)) { err = "bmp: image dimensions above the supported 65535 x 65535"; return false; }
  const size_t stride = (((size_t)w * bpp + 31) / 32) * 4;
  if ((size_t)off + stride * h > f.size()) { err = "bmp: truncated data"; return false; }
  rgb.resize((size_t)w * h * 3);
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t* row = &f[off + stride * (hi < 0 ? y : h - 1 - y)];
    for (uint32_t x = 0; x < w; ++x) {
      const uint8_t* p = row + (size_t)x * (bpp / 8);
      float* o = &rgb[((size_t)y * w + x) * 3];
      o[0] = (float)p[2] / 255.0f; o[1] = (float)p[1] / 255.0f; o[2] = (float)p[0] / 255.0f;
    }
  }
  return true;
}

thread_local std::string g_image_error;

// Philox4x32-10 (same generator as the render path; Random123 KAT-pinned in the tests)
void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

}  // namespace

extern "C" {

const char* ptb_image_last_error(void) { return g_image_error.c_str(); }

int32_t ptb_image_load(const char* filename, uint32_t* width, uint32_t* height, float** rgb) {
  if (!filename || !width || !height || !rgb) return PTB_ERR_INVALID;
  try {  // no C++ exception may cross the C boundary (std::bad_alloc from the decoders' buffers)
  std::vector<uint8_t> f;
  if (!read_file(filename, f)) { g_image_error = std::string("cannot read ") + filename; return PTB_ERR_IO; }
  std::vector<float> px;
  uint32_t w = 0, h = 0;
  bool ok;
  static const uint8_t png_sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (f.size() >= 8 && !std::memcmp(f.data(), png_sig, 8)) ok = decode_png(f, w, h, px, g_image_error);
  else if (f.size() >= 2 && f[0] == 'B' && f[1] == 'M') ok = decode_bmp(f, w, h, px, g_image_error);
  else if (f.size() >= 2 && f[0] == 'P') ok = decode_pnm(f, w, h, px, g_image_error);
  else { g_image_error = std::string("unsupported image format: ") + filename; return PTB_ERR_UNSUPPORTED; }
  if (!ok) return PTB_ERR_PARSE;
  float* out = (float*)std::malloc(px.size() * sizeof(float));
  if (!out) return PTB_ERR_OOM;
  std::memcpy(out, px.data(), px.size() * sizeof(float));
  *width = w; *height = h; *rgb = out;
  return PTB_OK;
  } catch (const std::bad_alloc&) {
    g_image_error = std::string("out of memory decoding ") + filename;
    return PTB_ERR_OOM;
  } catch (const std::exception& e) {
    g_image_error = std::string("decoder failure: ") + e.what();
    return PTB_ERR_PARSE;
  }
}
void ptb_image_free(float* rgb) { std::free(rgb); }

int32_t ptb_perlin_tables(uint64_t seed, float* out) {
  if (!out) return PTB_ERR_INVALID;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t blk[4];
  // ran_vecs: rng.gen_range(-1.0..1.0) * Vec3::one()   (textures/mod.rs:96-99)
  for (uint32_t i = 0; i < 256; ++i) {
    if ((i & 3u) == 0u) philox(i >> 2, 0u, 0u, 0u, k0, k1, blk);
    out[i] = 2.0f * ((float)(blk[i & 3u] >> 8) * (1.0f / 16777216.0f)) - 1.0f;
  }
  // generate_perm + permute (textures/mod.rs:141-159): identity, then for i in (1..256).rev() swap(i, gen_range(0..i))
  uint32_t* perm = reinterpret_cast<uint32_t*>(out) + 256;
  for (uint32_t t = 0; t < 3; ++t) {
    uint32_t* p = perm + 256u * t;
    for (uint32_t i = 0; i < 256; ++i) p[i] = i;
    uint32_t draw = 0;
    for (uint32_t i = 255; i >= 1; --i, ++draw) {
      if ((draw & 3u) == 0u) philox(draw >> 2, 1u + t, 0u, 0u, k0, k1, blk);
      const uint32_t target = (uint32_t)(((uint64_t)blk[draw & 3u] * i) >> 32);
      const uint32_t tmp = p[i]; p[i] = p[target]; p[target] = tmp;
    }
  }
  return PTB_OK;
}

}  // extern "C"
