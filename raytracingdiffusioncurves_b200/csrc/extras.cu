// extras.cu — the steps either side of the hot path in the reference's frame loop:
// view changes between frames, a running mean over frames (stand-in for the closed OptiX temporal
// denoiser's role: less noise at low ray counts), and a PNG writer for headless output.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "device_scene.h"

namespace {

// accum <- accum + (image - accum) / (frames_so_far + 1), all four channels
__global__ void k_accumulate(float4* accum, const float4* image, size_t n, float inv_count) {
  const size_t stride = (size_t)blockDim.x * gridDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float4 a = accum[i], v = image[i];
    accum[i] = make_float4(a.x + (v.x - a.x) * inv_count, a.y + (v.y - a.y) * inv_count, a.z + (v.z - a.z) * inv_count,
                           a.w + (v.w - a.w) * inv_count);
  }
}

uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
  static uint32_t table[256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    ready = true;
  }
  for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
  return crc;
}

void put_be32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

void png_chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& data) {
  put_be32(out, (uint32_t)data.size());
  const size_t start = out.size();
  out.insert(out.end(), type, type + 4);
  out.insert(out.end(), data.begin(), data.end());
  put_be32(out, crc32_update(0xFFFFFFFFu, out.data() + start, out.size() - start) ^ 0xFFFFFFFFu);
}

}  // namespace

extern "C" {

// scroll_callback (glfw_events.cpp:105-112): zoom_factor *= 1.5^-yoffset
void rdc_view_scroll(rdc_frame_params* p, double yoffset) {
  if (p) p->zoom_factor *= powf(1.5f, (float)-yoffset);
}

// mouse_cursor_callback while dragging (glfw_events.cpp:115-130): offset -= cursor delta * zoom_factor
void rdc_view_drag(rdc_frame_params* p, double dx, double dy) {
  if (!p) return;
  p->offset_x -= dx * p->zoom_factor;
  p->offset_y -= dy * p->zoom_factor;
}

// Image-level comparison of two float4 images in host memory (the diff tool). RGB only; pixels
// that are NaN (all rays missed, DeviceCode.cu:176-181) in BOTH images are skipped, NaN in one only counts as a
// full-scale error. psnr = 10 log10(1 / mse) on the [0,1] scale (+inf when identical), max_abs = largest |difference|.
int rdc_psnr(const float* a, const float* b, size_t n_pixels, double* psnr, double* max_abs) {
  if (!a || !b || n_pixels == 0 || !psnr || !max_abs) {
    rdc::set_error("psnr: bad argument");
    return RDC_E_INVALID;
  }
  double sum = 0.0, worst = 0.0;
  size_t counted = 0;
  for (size_t i = 0; i < n_pixels; ++i)
    for (int c = 0; c < 3; ++c) {
      const float x = a[4 * i + c], y = b[4 * i + c];
      const bool nx = x != x, ny = y != y;
      if (nx && ny) continue;
      const double d = (nx || ny) ? 1.0 : (double)x - (double)y;
      sum += d * d;
      if (std::fabs(d) > worst) worst = std::fabs(d);
      counted++;
    }
  const double mse = counted ? sum / (double)counted : 0.0;
  *psnr = mse == 0.0 ? INFINITY : 10.0 * std::log10(1.0 / mse);
  *max_abs = worst;
  return 0;
}

int rdc_accumulate(float* accum, const float* image, size_t n_pixels, uint32_t frames_so_far, rdc_stream stream) {
  if (!accum || !image || n_pixels == 0) {
    rdc::set_error("accumulate: bad argument");
    return RDC_E_INVALID;
  }
  if (frames_so_far == 0) {
    RDC_CUDA(cudaMemcpyAsync(accum, image, n_pixels * sizeof(float4), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
  }
  const unsigned blocks = (unsigned)((n_pixels + 255) / 256 < 148 * 16 ? (n_pixels + 255) / 256 : 148 * 16);
  k_accumulate<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(accum), reinterpret_cast<const float4*>(image),
                                                          n_pixels, 1.0f / (float)(frames_so_far + 1));
  RDC_CUDA(cudaGetLastError());
  return 0;
}

// 8-bit RGBA PNG, deflate "stored" blocks (no compression library in the image).
int rdc_write_png(const char* path, const uint8_t* rgba, int width, int height) {
  if (!path || !rgba || width <= 0 || height <= 0) {
    rdc::set_error("write_png: bad argument");
    return RDC_E_INVALID;
  }
  const size_t row = (size_t)width * 4 + 1;
  std::vector<uint8_t> raw(row * height);
  for (int y = 0; y < height; ++y) {
    raw[row * y] = 0;  // filter: none
    std::memcpy(&raw[row * y + 1], rgba + (size_t)4 * width * y, (size_t)width * 4);
  }
  std::vector<uint8_t> z;
  z.push_back(0x78); z.push_back(0x01);
  uint32_t a = 1, b = 0;
  for (size_t pos = 0; pos < raw.size();) {
    const size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
    z.push_back(pos + n == raw.size() ? 1 : 0);
    z.push_back((uint8_t)(n & 0xFF)); z.push_back((uint8_t)(n >> 8));
    z.push_back((uint8_t)(~n & 0xFF)); z.push_back((uint8_t)((~n >> 8) & 0xFF));
    z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
    for (size_t i = 0; i < n; ++i) {
      a = (a + raw[pos + i]) % 65521u;
      b = (b + a) % 65521u;
    }
    pos += n;
  }
  put_be32(z, (b << 16) | a);
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  std::vector<uint8_t> ihdr;
  put_be32(ihdr, (uint32_t)width); put_be32(ihdr, (uint32_t)height);
  ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  png_chunk(out, "IHDR", ihdr);
  png_chunk(out, "IDAT", z);
  png_chunk(out, "IEND", {});
  std::FILE* f = std::fopen(path, "wb");
  if (!f) {
    rdc::set_error("write_png: cannot open %s", path);
    return RDC_E_IO;
  }
  const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
  if (std::fclose(f) != 0 || !ok) {
    rdc::set_error("write_png: write failed for %s", path);
    return RDC_E_IO;
  }
  return 0;
}

}  // extern "C"
