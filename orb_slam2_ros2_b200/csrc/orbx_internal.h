// orbx_internal.h -- host-side state shared by the translation units behind the C ABI (orbx_api.cu, orbx_sequence.cu).
#pragma once

#include <string>
#include <vector>

#include "orbx_device.cuh"

struct orbx_ctx
{
  orbx_config cfg;
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  int n_img_max = 0;
  std::vector<orbx::Level> levels;
  std::vector<orbx::Tile> tiles;
  int n_tiles0 = 0; // level-0 tiles (first in `tiles`)
  std::vector<orbx::Cell> cells;
  orbx::Params p;            // template: geometry + base pointers
  size_t qt_smem = 0;
  std::string last_error;
  int64_t launches = 0;
  int64_t alg_bytes_image = 0, alg_bytes_stereo = 0;
  // device allocations
  std::vector<void *> allocs;
  uint8_t *d_in = nullptr;      // staging for host-side calls: [n_img_max][H][in_pitch]
  size_t in_pitch = 0;
  uint8_t *d_depth_in = nullptr; // [max_batch][H][W] float/uint16 (sized for float)
  int last_images = 0;          // images processed by the most recent call (for orbx_get_pyramid)
  int last_stereo = 0;
  int last_frames = 0;
  orbx::LevelMaps maps;      // TMA descriptors of the pyramid levels, FAST patch boxes (kernel parameter, __grid_constant__)
  orbx::LevelMaps maps_blur; // TMA descriptors of the blurred levels, BRIEF patch boxes
  orbx::LevelMaps maps_src;  // TMA descriptors over level 0 of `pyr`, one box size per resized level (its tiles' source rectangles)
  float min_u = 0, min_v = 0, max_u = 0, max_v = 0; // undistorted image bounds (VirtualFrame ctor, Frame.h:33-43)
  // host-batch pipeline: chunks of frames round-robin over kPipe streams so that H2D, kernels and D2H overlap
  static constexpr int kPipeMax = 8;
  int kPipe = 8;  // streams (tunable for experiments: ORBX_PIPE); 8 x 8 frames measured best on B200
  int kChunk = 8; // frames per chunk (ORBX_CHUNK)
  cudaStream_t pipe[kPipeMax] = {};
  cudaEvent_t fork_ev = nullptr;
  cudaEvent_t join_ev[kPipeMax] = {};
  // staging for the host-side matcher calls (orbx_search_in_area / orbx_verify_angle), grown on demand
  uint8_t *match_scratch = nullptr;
  size_t match_scratch_bytes = 0;
  // bag-of-words transform: per-feature and per-frame result buffers, allocated on first use
  orbx::BowArgs bow{};
  bool bow_ready = false;
  // sequences (orbx_sequence.cu): per-slot record staging for host outputs, gathered arrays of single-rank calls
  uint8_t *rec_staging = nullptr;
  size_t rec_staging_bytes = 0;
  struct orbx_comm *self_comm = nullptr;
  uint8_t *h_rec1 = nullptr; // pinned host record of the single-frame calls (one D2H copy instead of nine)
  // single-pair latency path: the launch sequence of one stereo frame captured once as a CUDA graph (orbx_set_graph)
  int use_graph = 1;
  cudaGraph_t graph1 = nullptr;
  cudaGraphExec_t graph1_exec = nullptr;
  cudaGraphNode_t graph1_pyr = nullptr; // the only node whose parameters carry the caller's input pointers
  const uint8_t *graph1_left = nullptr, *graph1_right = nullptr;
  size_t graph1_stride = 0, graph1_fs = 0;
  int graph1_kernels = 0;
  uint64_t frame_epoch = 0;   // bumped by every call that produces frames
  uint64_t bow_epoch = ~0ull; // frame_epoch at the last bag-of-words call
  int bow_frames = 0;         // frames covered by that call
};

namespace orbx
{

// records `msg` as the context's last error and returns `code`
int fail(orbx_ctx *c, int code, const std::string &msg);
// Params whose per-image buffers start at image `img0` (and per-frame buffers at frame `frame0`)
Params params_at(const orbx_ctx *c, int img0, int frame0);
// the kernels of a stereo batch for device slots [frame0, frame0 + nf) on stream s (inputs: frame 0 of the range at d_left / d_right)
int run_stereo_range(orbx_ctx *c, cudaStream_t s, int frame0, int nf, const uint8_t *d_left, const uint8_t *d_right, size_t stride, size_t frame_stride);

// one stereo frame in device slot 0: a single graph launch when the context's graph is enabled and `s` can be captured
int run_stereo_single(orbx_ctx *c, cudaStream_t s, const uint8_t *d_left, const uint8_t *d_right, size_t stride, size_t frame_stride);
// per-frame records (orbx_record_layout) of device slots [d0, d0 + nf) assembled at d_rec; device staging for host outputs
int pack_frame_records(orbx_ctx *c, cudaStream_t s, int d0, int nf, uint8_t *d_rec, size_t rec_stride);
int ensure_record_staging(orbx_ctx *c, size_t rec_stride);
// frees the gathered arrays a context keeps for single-rank sequence calls
void destroy_self_comm(orbx_ctx *c);

} // namespace orbx

#define ORBX_CUDA(ctx, expr)                                                                                               \
  do                                                                                                                       \
  {                                                                                                                        \
    cudaError_t e__ = (expr);                                                                                              \
    if (e__ != cudaSuccess) return orbx::fail((ctx), ORBX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)
