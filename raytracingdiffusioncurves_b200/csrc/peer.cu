// peer.cu — one frame split over the GPUs of a box, host side in C++ behind the C ABI.
//
// The reference's frame loop (optixHello.cpp:1163-1259) drives one GPU. Here every GPU ("rank") renders the 8-row
// strips t with t % world == rank of the same frame (scene and tree replicated, random numbers keyed by the global
// pixel index, so the assembled frame equals the single-GPU frame bit for bit) and the pieces meet in one of two ways:
//
//   device consumer (rdc_peer_render_frame): the render kernel stores every finished pixel straight into its place
//     in the consumers' frames over NVLink (rdc::render with target frames) — rank 0's finished frame when the scene
//     has no blur; everybody's copy of the rendered frame when it has, after which every rank blurs one contiguous
//     band and the blur's vertical pass stores into rank 0's finished frame. What is left of the collectives is one
//     barrier per exchange step: k_peer_barrier, flags in peer memory (release/acquire at system scope).
//   host consumer (rdc_peer_frame_to_host): nothing is gathered on a GPU at all. Every rank copies ITS OWN rows of
//     the finished frame over ITS OWN PCIe link into one host frame all ranks share (pinned memory; a POSIX shared
//     memory segment registered with CUDA when the ranks are processes, rdc_host_frame_*): packed strips in one strided
//     copy when there is no blur, the rank's blurred band otherwise. Round 1 gathered to GPU 0 and pushed the whole
//     frame through one link; eight links carry an eighth each.
//
// Ranks may be threads-of-nothing in one process (OptixHello --gpus N: one host thread enqueues on every device;
// rdc_peer_frames_connect_local, cudaDeviceEnablePeerAccess) or one process per GPU (torchrun: rdc_peer_frames_export
// / _connect_ipc over cudaIpcMemHandle; the 64-byte handles travel over whatever the launcher offers).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "device_scene.h"

struct rdc_peer_frames {
  int rank = 0, world = 1, device = 0;
  uint32_t width = 0, height = 0;
  // what this rank owns (device memory of `device`)
  float4* full_image = nullptr;   // the whole rendered frame (blur scenes: every rank's copy is filled)
  float* full_sigma = nullptr;    // + one float: the max-sigma flag of this rank's launches
  float4* frames[2] = {nullptr, nullptr};  // finished frames, used in turn (consumed on rank 0; the band blur's target elsewhere)
  float4* scratch = nullptr;      // blur scratch
  float4* packed[2] = {nullptr, nullptr};  // this rank's strips, packed (host path without blur), used in turn
  float* packed_sigma = nullptr;
  unsigned int* pads = nullptr;   // [world] arrival flags written by the peers, [world] = error flag
  // every rank's buffers as seen from here (own entries = the pointers above)
  float4* peer_full_image[RDC_MAX_FRAME_TARGETS] = {};
  float* peer_full_sigma[RDC_MAX_FRAME_TARGETS] = {};
  float4* peer_frames[2][RDC_MAX_FRAME_TARGETS] = {};
  unsigned int* peer_pads[RDC_MAX_FRAME_TARGETS] = {};
  bool opened_ipc = false, connected = false;
  unsigned int epoch = 0;
  int turn = 0;
  // host path: copies run on their own stream so that frame f's copy overlaps frame f+1's rendering
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t rendered[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
  int host_slot = 0;
};

namespace rdc {
namespace {

constexpr int kBuffers = 5;  // full_image, full_sigma, frames[0], frames[1], pads — the order of the export blob

// Every rank tells every other "I am at barrier number `epoch`" and waits until all have said so: thread t handles
// peer t. Writes that precede the barrier on a rank's stream (the render kernel's stores into peers' frames) are
// ordered before the flag by the release at system scope; the waiter's acquire orders its later reads after it.
// A peer that never arrives (a rank died) must not wedge the GPU: after ~4 s of spinning the error flag is raised and
// the kernel returns.
struct PadTable {
  unsigned int* p[RDC_MAX_FRAME_TARGETS];
};

__global__ void k_peer_barrier(const PadTable peers, unsigned int* my_pads, int rank, int world, unsigned int epoch) {
  const int t = threadIdx.x;
  if (t >= world) return;
  if (t != rank) {
    unsigned int* flag = peers.p[t] + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    const long long start = clock64();
    for (;;) {
      unsigned int seen;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(my_pads + t) : "memory");
      if ((int)(seen - epoch) >= 0) break;
      if (clock64() - start > 8000000000ll) {  // ~4 s at 2 GHz
        atomicExch(my_pads + world, 1u);
        break;
      }
      __nanosleep(64);
    }
  }
}

int free_local(rdc_peer_frames* f) {
  cudaFree(f->full_image);
  cudaFree(f->full_sigma);
  cudaFree(f->frames[0]);
  cudaFree(f->frames[1]);
  cudaFree(f->scratch);
  cudaFree(f->packed[0]);
  cudaFree(f->packed[1]);
  cudaFree(f->packed_sigma);
  cudaFree(f->pads);
  return 0;
}

}  // namespace
}  // namespace rdc

extern "C" {

void rdc_peer_frames_destroy(rdc_peer_frames* f);

int rdc_peer_frames_create(uint32_t width, uint32_t height, int rank, int world, rdc_peer_frames** out) {
  if (!out || width == 0 || height == 0 || world < 1 || world > RDC_MAX_FRAME_TARGETS || rank < 0 || rank >= world) {
    rdc::set_error("peer frames: bad argument (1 <= world <= %d, 0 <= rank < world)", RDC_MAX_FRAME_TARGETS);
    return RDC_E_INVALID;
  }
  rdc_peer_frames* f = new (std::nothrow) rdc_peer_frames();
  if (!f) return RDC_E_LIMIT;
  f->rank = rank;
  f->world = world;
  f->width = width;
  f->height = height;
  cudaGetDevice(&f->device);
  const size_t n = (size_t)width * height;
  const uint32_t strips = (height + RDC_STRIP_ROWS - 1) / RDC_STRIP_ROWS;
  const size_t packed_rows = (size_t)((strips + world - 1) / world) * RDC_STRIP_ROWS;
  cudaError_t e = cudaMalloc((void**)&f->full_image, n * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc((void**)&f->full_sigma, (n + 1) * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&f->frames[0], n * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc((void**)&f->frames[1], n * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc((void**)&f->scratch, n * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc((void**)&f->packed[0], packed_rows * width * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc((void**)&f->packed[1], packed_rows * width * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc((void**)&f->packed_sigma, (packed_rows * width + 1) * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&f->pads, (RDC_MAX_FRAME_TARGETS + 1) * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(f->pads, 0, (RDC_MAX_FRAME_TARGETS + 1) * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->copy_stream, cudaStreamNonBlocking);
  for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
    e = cudaEventCreateWithFlags(&f->rendered[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->copied[k], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    rdc::free_local(f);
    delete f;
    return rdc::cuda_fail(e, "peer frames: allocation");
  }
  f->peer_full_image[rank] = f->full_image;
  f->peer_full_sigma[rank] = f->full_sigma;
  f->peer_frames[0][rank] = f->frames[0];
  f->peer_frames[1][rank] = f->frames[1];
  f->peer_pads[rank] = f->pads;
  f->connected = world == 1;
  // load every kernel the frame loop launches now: a first launch in the middle of a frame would load its module then,
  // which may wait for the device to go idle — while a barrier kernel of this very frame is spinning on it
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, rdc::k_peer_barrier);
  if (e != cudaSuccess || rdc::blur_preload() != 0) {
    rdc_peer_frames_destroy(f);
    return e != cudaSuccess ? rdc::cuda_fail(e, "peer frames: kernel load") : RDC_E_INVALID;
  }
  *out = f;
  return 0;
}

int rdc_peer_frames_export(const rdc_peer_frames* f, void* handle_bytes) {
  if (!f || !handle_bytes) {
    rdc::set_error("peer export: null argument");
    return RDC_E_INVALID;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "RDC_PEER_HANDLE_BYTES assumes 64-byte IPC handles");
  void* bufs[rdc::kBuffers] = {f->full_image, f->full_sigma, f->frames[0], f->frames[1], f->pads};
  for (int k = 0; k < rdc::kBuffers; ++k) {
    cudaIpcMemHandle_t h;
    RDC_CUDA(cudaIpcGetMemHandle(&h, bufs[k]));
    std::memcpy(static_cast<char*>(handle_bytes) + 64 * k, &h, 64);
  }
  return 0;
}

int rdc_peer_frames_connect_ipc(rdc_peer_frames* f, const void* all_handles) {
  if (!f || !all_handles) {
    rdc::set_error("peer connect: null argument");
    return RDC_E_INVALID;
  }
  if (f->connected) return 0;
  for (int r = 0; r < f->world; ++r) {
    if (r == f->rank) continue;
    void* ptr[rdc::kBuffers];
    for (int k = 0; k < rdc::kBuffers; ++k) {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, static_cast<const char*>(all_handles) + (size_t)RDC_PEER_HANDLE_BYTES * r + 64 * k, 64);
      RDC_CUDA(cudaIpcOpenMemHandle(&ptr[k], h, cudaIpcMemLazyEnablePeerAccess));
    }
    f->peer_full_image[r] = static_cast<float4*>(ptr[0]);
    f->peer_full_sigma[r] = static_cast<float*>(ptr[1]);
    f->peer_frames[0][r] = static_cast<float4*>(ptr[2]);
    f->peer_frames[1][r] = static_cast<float4*>(ptr[3]);
    f->peer_pads[r] = static_cast<unsigned int*>(ptr[4]);
  }
  f->opened_ipc = true;
  f->connected = true;
  return 0;
}

int rdc_peer_frames_connect_local(rdc_peer_frames* const* all, int world) {
  if (!all || world < 1 || world > RDC_MAX_FRAME_TARGETS) {
    rdc::set_error("peer connect: bad argument");
    return RDC_E_INVALID;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  for (int r = 0; r < world; ++r) {
    rdc_peer_frames* f = all[r];
    if (!f || f->world != world || f->rank != r) {
      rdc::set_error("peer connect: entry %d is not rank %d of %d", r, r, world);
      return RDC_E_INVALID;
    }
    RDC_CUDA(cudaSetDevice(f->device));
    for (int q = 0; q < world; ++q) {
      if (q == r) continue;
      if (all[q]->device != f->device) {
        int can = 0;
        RDC_CUDA(cudaDeviceCanAccessPeer(&can, f->device, all[q]->device));
        if (!can) {
          rdc::set_error("peer connect: device %d cannot address device %d", f->device, all[q]->device);
          cudaSetDevice(prev);
          return RDC_E_LIMIT;
        }
        cudaError_t e = cudaDeviceEnablePeerAccess(all[q]->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) {
          cudaSetDevice(prev);
          return rdc::cuda_fail(e, "cudaDeviceEnablePeerAccess");
        }
      }
      f->peer_full_image[q] = all[q]->full_image;
      f->peer_full_sigma[q] = all[q]->full_sigma;
      f->peer_frames[0][q] = all[q]->frames[0];
      f->peer_frames[1][q] = all[q]->frames[1];
      f->peer_pads[q] = all[q]->pads;
    }
    f->connected = true;
  }
  cudaSetDevice(prev);
  return 0;
}

void rdc_peer_frames_destroy(rdc_peer_frames* f) {
  if (!f) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(f->device);
  cudaDeviceSynchronize();
  if (f->opened_ipc)
    for (int r = 0; r < f->world; ++r) {
      if (r == f->rank) continue;
      cudaIpcCloseMemHandle(f->peer_full_image[r]);
      cudaIpcCloseMemHandle(f->peer_full_sigma[r]);
      cudaIpcCloseMemHandle(f->peer_frames[0][r]);
      cudaIpcCloseMemHandle(f->peer_frames[1][r]);
      cudaIpcCloseMemHandle(f->peer_pads[r]);
    }
  if (f->copy_stream) cudaStreamDestroy(f->copy_stream);
  for (int k = 0; k < 2; ++k) {
    if (f->rendered[k]) cudaEventDestroy(f->rendered[k]);
    if (f->copied[k]) cudaEventDestroy(f->copied[k]);
  }
  rdc::free_local(f);
  cudaSetDevice(prev);
  delete f;
}

int rdc_peer_barrier(rdc_peer_frames* f, rdc_stream stream) {
  if (!f || !f->connected) {
    rdc::set_error("peer barrier: frames are not connected");
    return RDC_E_INVALID;
  }
  if (f->world == 1) return 0;
  rdc::PadTable peers;
  for (int r = 0; r < RDC_MAX_FRAME_TARGETS; ++r) peers.p[r] = f->peer_pads[r];
  f->epoch += 1;  // every rank calls the barrier the same number of times: the counters agree without talking
  rdc::k_peer_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(peers, f->pads, f->rank, f->world, f->epoch);
  RDC_CUDA(cudaGetLastError());
  return 0;
}

int rdc_peer_status(rdc_peer_frames* f) {
  if (!f) return RDC_E_INVALID;
  unsigned int flag = 0;
  RDC_CUDA(cudaMemcpy(&flag, f->pads + f->world, sizeof flag, cudaMemcpyDeviceToHost));
  if (flag) {
    rdc::set_error("peer barrier: a rank did not arrive within the time limit");
    return RDC_E_LIMIT;
  }
  return 0;
}

int rdc_peer_frame_ptr(const rdc_peer_frames* f, int turn, float** image) {
  if (!f || !image || turn < 0 || turn > 1) return RDC_E_INVALID;
  *image = reinterpret_cast<float*>(f->frames[turn]);
  return 0;
}

// One frame over all ranks, device consumer: the finished frame ends up in rank 0's frames[turn]; *frame_out gets
// that pointer on rank 0 and NULL elsewhere. Enqueue-only. `wait_event` (cudaEvent_t or NULL): the stream waits for it
// right before the frame's first barrier — rank 0 passes "whatever still reads the buffer the NEXT frame will be
// written into has finished" (the frame returned two calls ago), because the peers only start that next frame after
// this barrier. halo_rows = ceil(3 * largest sigma the scene can produce), 0 = the scene has no blur.
int rdc_peer_render_frame(rdc_scene* scene, rdc_peer_frames* f, const rdc_frame_params* params, int use_blur, int halo_rows,
                          void* wait_event, rdc_stream stream, float** frame_out) {
  if (!scene || !f || !params || !f->connected) {
    rdc::set_error("peer render: null argument or frames not connected");
    return RDC_E_INVALID;
  }
  if (params->image_width != f->width || params->image_height != f->height || params->row_begin != 0 ||
      params->row_end != params->image_height) {
    rdc::set_error("peer render: the parameters must describe the whole %ux%u frame", f->width, f->height);
    return RDC_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool blur = use_blur && halo_rows > 0;
  const int turn = f->turn;
  f->turn ^= 1;
  rdc_frame_params p = *params;
  p.strip_stride = f->world > 1 ? (uint32_t)f->world : 0;
  p.strip_offset = f->world > 1 ? (uint32_t)f->rank : 0;
  if (frame_out) *frame_out = f->rank == 0 ? reinterpret_cast<float*>(f->frames[turn]) : nullptr;
  if (!blur) {
    float* image0 = reinterpret_cast<float*>(f->peer_frames[turn][0]);
    float* sigma0 = f->peer_full_sigma[0];
    if (int rc = rdc::render(scene, p, nullptr, nullptr, st, 1, &image0, &sigma0)) return rc;  // the gather
    if (wait_event) RDC_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)wait_event, 0));
    return rdc_peer_barrier(f, stream);
  }
  float* images[RDC_MAX_FRAME_TARGETS];
  float* sigmas[RDC_MAX_FRAME_TARGETS];
  for (int r = 0; r < f->world; ++r) {
    images[r] = reinterpret_cast<float*>(f->peer_full_image[r]);
    sigmas[r] = f->peer_full_sigma[r];
  }
  if (int rc = rdc::render(scene, p, nullptr, nullptr, st, (uint32_t)f->world, images, sigmas)) return rc;  // the all-gather
  if (wait_event) RDC_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)wait_event, 0));
  if (int rc = rdc_peer_barrier(f, stream)) return rc;
  // contiguous band of this rank, remainder rows spread over the first ranks
  const int base = (int)f->height / f->world, rem = (int)f->height % f->world;
  const int b = f->rank * base + (f->rank < rem ? f->rank : rem), e = b + base + (f->rank < rem ? 1 : 0);
  if (e > b)
    if (int rc = rdc::gaussian_blur(f->peer_frames[turn][0], f->full_image, f->full_sigma, f->scratch, (int)f->width, (int)f->height, b, e,
                                    nullptr, st, halo_rows))
      return rc;  // the vertical pass stores into rank 0's finished frame: the gather
  return rdc_peer_barrier(f, stream);
}

// One frame over all ranks, host consumer: this rank's rows of the finished frame are copied, over this rank's own
// PCIe link, to their place in `host_frame` — the base of ONE full frame (float4[height*width]) in pinned host memory
// that all ranks address (rdc_host_frame_open for ranks in different processes). Enqueue-only: render (+ exchange and
// band blur when the scene has blur) on `stream`, the copy on the handle's own copy stream so that it overlaps the
// next frame's rendering; rdc_peer_frames_wait blocks until this rank's copies have landed. Give consecutive frames
// different host buffers. A rank's rows: without blur the 8-row strips it rendered, with blur its contiguous band.
int rdc_peer_frame_to_host(rdc_scene* scene, rdc_peer_frames* f, const rdc_frame_params* params, int use_blur, int halo_rows,
                           float* host_frame, rdc_stream stream) {
  if (!scene || !f || !params || !host_frame || !f->connected) {
    rdc::set_error("peer frame to host: null argument or frames not connected");
    return RDC_E_INVALID;
  }
  if (params->image_width != f->width || params->image_height != f->height || params->row_begin != 0 ||
      params->row_end != params->image_height) {
    rdc::set_error("peer frame to host: the parameters must describe the whole %ux%u frame", f->width, f->height);
    return RDC_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool blur = use_blur && halo_rows > 0;
  const int slot = f->host_slot ^= 1;
  const size_t row_bytes = (size_t)f->width * sizeof(float4);
  rdc_frame_params p = *params;
  p.strip_stride = f->world > 1 ? (uint32_t)f->world : 0;
  p.strip_offset = f->world > 1 ? (uint32_t)f->rank : 0;
  if (!blur) {
    // nothing crosses NVLink: render my strips packed, one strided copy puts them in place
    float4* image = f->world > 1 ? f->packed[slot] : f->frames[slot];
    float* sigma = f->world > 1 ? f->packed_sigma : f->full_sigma;
    RDC_CUDA(cudaStreamWaitEvent(st, f->copied[slot], 0));  // the copy that last read this buffer
    if (int rc = rdc::render(scene, p, image, sigma, st)) return rc;
    RDC_CUDA(cudaEventRecord(f->rendered[slot], st));
    RDC_CUDA(cudaStreamWaitEvent(f->copy_stream, f->rendered[slot], 0));
    const uint32_t strips = (f->height + RDC_STRIP_ROWS - 1) / RDC_STRIP_ROWS;
    const uint32_t mine = (uint32_t)f->rank < strips ? (strips - f->rank + f->world - 1) / f->world : 0;
    if (mine > 0) {
      // strip k of mine is strip rank + k*world of the frame; the last strip of the frame may be short
      const uint32_t last_strip = f->rank + (mine - 1) * f->world;
      const uint32_t last_rows = last_strip == strips - 1 ? f->height - last_strip * RDC_STRIP_ROWS : RDC_STRIP_ROWS;
      const uint32_t full = last_rows == RDC_STRIP_ROWS ? mine : mine - 1;
      char* dst = reinterpret_cast<char*>(host_frame) + (size_t)f->rank * RDC_STRIP_ROWS * row_bytes;
      if (full > 0)
        RDC_CUDA(cudaMemcpy2DAsync(dst, (size_t)f->world * RDC_STRIP_ROWS * row_bytes, image, RDC_STRIP_ROWS * row_bytes,
                                   RDC_STRIP_ROWS * row_bytes, full, cudaMemcpyDeviceToHost, f->copy_stream));
      if (full < mine)
        RDC_CUDA(cudaMemcpyAsync(dst + (size_t)full * f->world * RDC_STRIP_ROWS * row_bytes,
                                 reinterpret_cast<char*>(image) + (size_t)full * RDC_STRIP_ROWS * row_bytes, last_rows * row_bytes,
                                 cudaMemcpyDeviceToHost, f->copy_stream));
    }
    RDC_CUDA(cudaEventRecord(f->copied[slot], f->copy_stream));
    return 0;
  }
  // blur: every rank needs the whole rendered frame (the taps reach across strips) — all-gather through the render
  // kernel's stores, barrier, then this rank blurs its band into its own frames[slot] and copies the band out
  float* images[RDC_MAX_FRAME_TARGETS];
  float* sigmas[RDC_MAX_FRAME_TARGETS];
  for (int r = 0; r < f->world; ++r) {
    images[r] = reinterpret_cast<float*>(f->peer_full_image[r]);
    sigmas[r] = f->peer_full_sigma[r];
  }
  // full_image of a rank is read by its band blur and overwritten by every rank's next render: a closing barrier per frame
  if (int rc = rdc::render(scene, p, nullptr, nullptr, st, (uint32_t)f->world, images, sigmas)) return rc;
  if (int rc = rdc_peer_barrier(f, stream)) return rc;
  const int base = (int)f->height / f->world, rem = (int)f->height % f->world;
  const int b = f->rank * base + (f->rank < rem ? f->rank : rem), e = b + base + (f->rank < rem ? 1 : 0);
  RDC_CUDA(cudaStreamWaitEvent(st, f->copied[slot], 0));
  if (e > b)
    if (int rc = rdc::gaussian_blur(f->frames[slot], f->full_image, f->full_sigma, f->scratch, (int)f->width, (int)f->height, b, e, nullptr,
                                    st, halo_rows))
      return rc;
  RDC_CUDA(cudaEventRecord(f->rendered[slot], st));
  if (int rc = rdc_peer_barrier(f, stream)) return rc;
  RDC_CUDA(cudaStreamWaitEvent(f->copy_stream, f->rendered[slot], 0));
  if (e > b)
    RDC_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(host_frame) + (size_t)b * row_bytes,
                             reinterpret_cast<char*>(f->frames[slot]) + (size_t)b * row_bytes, (size_t)(e - b) * row_bytes,
                             cudaMemcpyDeviceToHost, f->copy_stream));
  RDC_CUDA(cudaEventRecord(f->copied[slot], f->copy_stream));
  return 0;
}

int rdc_peer_frames_wait(rdc_peer_frames* f) {
  if (!f) {
    rdc::set_error("peer wait: null argument");
    return RDC_E_INVALID;
  }
  RDC_CUDA(cudaStreamSynchronize(f->copy_stream));
  return rdc_peer_status(f);
}

// ---- one host frame for all ranks: POSIX shared memory, page-locked and registered with CUDA in every process ----
int rdc_host_frame_open(const char* name, size_t bytes, int create, float** out) {
  if (!name || !out || bytes == 0) {
    rdc::set_error("host frame: bad argument");
    return RDC_E_INVALID;
  }
  const int fd = shm_open(name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
  if (fd < 0) {
    rdc::set_error("host frame: shm_open(%s) failed", name);
    return RDC_E_IO;
  }
  if (create && ftruncate(fd, (off_t)bytes) != 0) {
    close(fd);
    rdc::set_error("host frame: cannot size %s to %zu bytes", name, bytes);
    return RDC_E_IO;
  }
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) {
    rdc::set_error("host frame: mmap(%s) failed", name);
    return RDC_E_IO;
  }
  if (create) std::memset(p, 0, bytes);  // touch every page before it is pinned
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    munmap(p, bytes);
    return rdc::cuda_fail(e, "cudaHostRegister(host frame)");
  }
  *out = static_cast<float*>(p);
  return 0;
}

int rdc_host_frame_close(const char* name, float* frame, size_t bytes, int unlink_it) {
  if (frame) {
    cudaHostUnregister(frame);
    munmap(frame, bytes);
  }
  if (unlink_it && name) shm_unlink(name);
  return 0;
}

}  // extern "C"
