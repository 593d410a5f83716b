// jpeg.cpp — baseline JPEG writer (JFIF, 8-bit, YCbCr 4:4:4, the standard Annex K tables).
//
// The reference's F11 screenshot ends in stbi_write_jpg(name, W, H, 4, rgba, W*4) (glfw_events.cpp:92-94): four
// components in, alpha ignored, and — the stride landing in the quality parameter — quality clamped to 100.
// This is the headless equivalent behind `OptixHello --out file.jpg` (quality 100 by default for the same
// result); written from the JPEG specification, no third-party code.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/rdc_b200.h"

namespace rdc {
void set_error(const char* fmt, ...);  // capi.cu
}

namespace {

const uint8_t kZigZag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

const uint8_t kQuantLuma[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kQuantChroma[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                  99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                  99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

// Annex K.3: number of codes of each length 1..16, then the symbols in code order
const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81,
    0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18,
    0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75,
    0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
    0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5,
    0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08,
    0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25,
    0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47,
    0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74,
    0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97,
    0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4,
    0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

struct HuffTable {
  uint16_t code[256];
  uint8_t length[256];  // 0: the symbol has no code
};

// canonical code assignment (Annex C): codes of one length are consecutive, the first of a length is twice the
// successor of the last of the previous length
bool build_table(const uint8_t bits[16], const uint8_t* vals, int n_vals, HuffTable& t) {
  std::memset(&t, 0, sizeof t);
  uint32_t code = 0;
  int k = 0;
  for (int len = 1; len <= 16; ++len) {
    for (int i = 0; i < bits[len - 1]; ++i) {
      if (k >= n_vals || t.length[vals[k]] != 0 || code >= (1u << len)) return false;
      t.code[vals[k]] = (uint16_t)code++;
      t.length[vals[k]] = (uint8_t)len;
      ++k;
    }
    code <<= 1;
  }
  return k == n_vals;
}

struct BitWriter {
  std::vector<uint8_t>& out;
  uint32_t acc = 0;
  int n = 0;
  void put(uint32_t value, int bits) {
    acc = (acc << bits) | (value & ((1u << bits) - 1u));
    n += bits;
    while (n >= 8) {
      const uint8_t byte = (uint8_t)(acc >> (n - 8));
      out.push_back(byte);
      if (byte == 0xFF) out.push_back(0x00);  // byte stuffing
      n -= 8;
    }
  }
  void flush() {
    if (n > 0) put((1u << (8 - n)) - 1u, 8 - n);  // pad with ones
  }
};

void put16(std::vector<uint8_t>& o, unsigned v) {
  o.push_back((uint8_t)(v >> 8));
  o.push_back((uint8_t)(v & 0xFF));
}

// 8x8 forward DCT-II, separable, straight from the definition (a screenshot writer, not a hot path)
void fdct8x8(const float in[64], float out[64]) {
  static float basis[8][8];
  static bool ready = false;
  if (!ready) {
    for (int u = 0; u < 8; ++u)
      for (int x = 0; x < 8; ++x)
        basis[u][x] = (u == 0 ? std::sqrt(0.125f) : 0.5f) * std::cos((2 * x + 1) * u * 3.14159265358979323846f / 16.0f);
    ready = true;
  }
  float tmp[64];
  for (int y = 0; y < 8; ++y)
    for (int u = 0; u < 8; ++u) {
      float s = 0.0f;
      for (int x = 0; x < 8; ++x) s += in[8 * y + x] * basis[u][x];
      tmp[8 * y + u] = s;
    }
  for (int v = 0; v < 8; ++v)
    for (int u = 0; u < 8; ++u) {
      float s = 0.0f;
      for (int y = 0; y < 8; ++y) s += tmp[8 * y + u] * basis[v][y];
      out[8 * v + u] = s;
    }
}

int bit_size(int v) {  // number of bits of |v| (the JPEG "category")
  int a = v < 0 ? -v : v, n = 0;
  while (a) {
    ++n;
    a >>= 1;
  }
  return n;
}

void encode_block(BitWriter& bw, const float pixels[64], const uint8_t quant[64], const HuffTable& dc, const HuffTable& ac,
                  int& prev_dc) {
  float coef[64];
  fdct8x8(pixels, coef);
  int q[64];
  for (int i = 0; i < 64; ++i) {
    const float v = coef[kZigZag[i]] / (float)quant[kZigZag[i]];
    q[i] = (int)std::lrintf(v);
  }
  const int diff = q[0] - prev_dc;
  prev_dc = q[0];
  int size = bit_size(diff);
  bw.put(dc.code[size], dc.length[size]);
  if (size) bw.put((uint32_t)(diff < 0 ? diff - 1 : diff), size);  // negative values: one's complement of |v|
  int last = 63;
  while (last > 0 && q[last] == 0) --last;
  int run = 0;
  for (int i = 1; i <= last; ++i) {
    if (q[i] == 0) {
      ++run;
      continue;
    }
    while (run > 15) {
      bw.put(ac.code[0xF0], ac.length[0xF0]);  // ZRL: sixteen zeros
      run -= 16;
    }
    size = bit_size(q[i]);
    const int symbol = (run << 4) | size;
    bw.put(ac.code[symbol], ac.length[symbol]);
    bw.put((uint32_t)(q[i] < 0 ? q[i] - 1 : q[i]), size);
    run = 0;
  }
  if (last < 63) bw.put(ac.code[0x00], ac.length[0x00]);  // EOB
}

}  // namespace

extern "C" int rdc_write_jpg(const char* path, const uint8_t* rgba, int width, int height, int quality) {
  if (!path || !rgba || width <= 0 || height <= 0 || width > 65535 || height > 65535) {
    rdc::set_error("write_jpg: bad argument (sizes up to 65535)");
    return RDC_E_INVALID;
  }
  if (quality < 1) quality = 1;
  if (quality > 100) quality = 100;  // what stb does with the reference's out-of-range value
  HuffTable dc_l, dc_c, ac_l, ac_c;
  if (!build_table(kDcLumaBits, kDcVals, 12, dc_l) || !build_table(kDcChromaBits, kDcVals, 12, dc_c) ||
      !build_table(kAcLumaBits, kAcLumaVals, 162, ac_l) || !build_table(kAcChromaBits, kAcChromaVals, 162, ac_c)) {
    rdc::set_error("write_jpg: Huffman tables are inconsistent");
    return RDC_E_INVALID;
  }
  // quality -> scale of the Annex K quantisation tables (the IJG convention stb follows)
  const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
  uint8_t ql[64], qc[64];
  for (int i = 0; i < 64; ++i) {
    int a = (kQuantLuma[i] * scale + 50) / 100, b = (kQuantChroma[i] * scale + 50) / 100;
    ql[i] = (uint8_t)(a < 1 ? 1 : a > 255 ? 255 : a);
    qc[i] = (uint8_t)(b < 1 ? 1 : b > 255 ? 255 : b);
  }

  std::vector<uint8_t> o;
  o.reserve((size_t)width * height / 2 + 1024);
  put16(o, 0xFFD8);  // SOI
  put16(o, 0xFFE0);  // APP0 / JFIF 1.01, no density, no thumbnail
  put16(o, 16);
  const uint8_t jfif[] = {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
  o.insert(o.end(), jfif, jfif + sizeof jfif);
  for (int t = 0; t < 2; ++t) {  // DQT, zig-zag order
    put16(o, 0xFFDB);
    put16(o, 67);
    o.push_back((uint8_t)t);
    for (int i = 0; i < 64; ++i) o.push_back((t ? qc : ql)[kZigZag[i]]);
  }
  put16(o, 0xFFC0);  // SOF0: baseline, 8 bits, three components sampled 1x1
  put16(o, 17);
  o.push_back(8);
  put16(o, (unsigned)height);
  put16(o, (unsigned)width);
  o.push_back(3);
  for (int c = 0; c < 3; ++c) {
    o.push_back((uint8_t)(c + 1));
    o.push_back(0x11);
    o.push_back((uint8_t)(c ? 1 : 0));
  }
  auto dht = [&](int cls_id, const uint8_t bits[16], const uint8_t* vals, int n) {
    put16(o, 0xFFC4);
    put16(o, (unsigned)(19 + n));
    o.push_back((uint8_t)cls_id);
    o.insert(o.end(), bits, bits + 16);
    o.insert(o.end(), vals, vals + n);
  };
  dht(0x00, kDcLumaBits, kDcVals, 12);
  dht(0x10, kAcLumaBits, kAcLumaVals, 162);
  dht(0x01, kDcChromaBits, kDcVals, 12);
  dht(0x11, kAcChromaBits, kAcChromaVals, 162);
  put16(o, 0xFFDA);  // SOS
  put16(o, 12);
  o.push_back(3);
  o.push_back(1); o.push_back(0x00);
  o.push_back(2); o.push_back(0x11);
  o.push_back(3); o.push_back(0x11);
  o.push_back(0); o.push_back(63); o.push_back(0);

  BitWriter bw{o};
  int prev[3] = {0, 0, 0};
  float block[3][64];
  for (int by = 0; by < height; by += 8)
    for (int bx = 0; bx < width; bx += 8) {
      for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) {
          const int sy = by + y < height ? by + y : height - 1, sx = bx + x < width ? bx + x : width - 1;  // edge repeats
          const uint8_t* p = rgba + ((size_t)sy * width + sx) * 4;
          const float r = p[0], g = p[1], b = p[2];
          block[0][8 * y + x] = 0.299f * r + 0.587f * g + 0.114f * b - 128.0f;
          block[1][8 * y + x] = -0.168736f * r - 0.331264f * g + 0.5f * b;
          block[2][8 * y + x] = 0.5f * r - 0.418688f * g - 0.081312f * b;
        }
      encode_block(bw, block[0], ql, dc_l, ac_l, prev[0]);
      encode_block(bw, block[1], qc, dc_c, ac_c, prev[1]);
      encode_block(bw, block[2], qc, dc_c, ac_c, prev[2]);
    }
  bw.flush();
  put16(o, 0xFFD9);  // EOI

  FILE* f = std::fopen(path, "wb");
  if (!f) {
    rdc::set_error("write_jpg: cannot open %s", path);
    return RDC_E_IO;
  }
  const bool ok = std::fwrite(o.data(), 1, o.size(), f) == o.size();
  if (std::fclose(f) != 0 || !ok) {
    rdc::set_error("write_jpg: short write to %s", path);
    return RDC_E_IO;
  }
  return 0;
}
