// Internal (non-ABI) definitions shared by the translation units of libposeengine.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: the ranges cost nothing unless a profiler (nsys / ncu --nvtx) is attached

// NVTX range for one host-side stage of an entry point ("pe_topdown/forward", ...): shows up on the timeline next to the
// per-stage CUDA-event timers (pe_model_profile)
struct PeRange {
  explicit PeRange(const char* name) { nvtxRangePushA(name); }
  ~PeRange() { nvtxRangePop(); }
  PeRange(const PeRange&) = delete;
  PeRange& operator=(const PeRange&) = delete;
};

#include "../../include/poseengine.h"

struct pe_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  uint8_t* d_frames = nullptr;       // owned frame store
  const uint8_t* frames = nullptr;   // current frames (owned store or caller's device memory)
  size_t frames_cap = 0;
  int n_frames = 0, fh = 0, fw = 0;
  // double-buffered upload slots (pe_frames_upload / pe_frames_select): a decode thread copies block k+1 on copy_stream while
  // the engine stream computes on block k
  cudaStream_t copy_stream = nullptr;
  uint8_t* d_slot[2] = {nullptr, nullptr};       // current block of each slot: engine-owned buffer or caller memory
  uint8_t* d_slot_own[2] = {nullptr, nullptr};   // the engine-owned buffers
  size_t slot_cap[2] = {0, 0};
  int slot_n[2] = {0, 0}, slot_h[2] = {0, 0}, slot_w[2] = {0, 0};
  cudaEvent_t slot_ready[2] = {nullptr, nullptr};
  // handles created on this engine and still alive: pe_engine_destroy destroys them first, so a model / lifter handle
  // released after its engine (garbage-collection order is arbitrary on the Python side) is a no-op, never a dangling `e`
  std::vector<pe_model*> models;
  std::vector<pe_lifter*> lifters;
  std::vector<pe_detector*> detectors;
};

// Live-handle registry (engine.cu).  Every pe_*_destroy first asks pe_handle_release(): false = the handle is not (or no
// longer) alive -> the destroy call returns PE_OK without touching it.
enum { PE_H_ENGINE = 0, PE_H_MODEL = 1, PE_H_LIFTER = 2, PE_H_DETECTOR = 3 };
void pe_handle_register(int kind, void* h);
bool pe_handle_release(int kind, void* h);
bool pe_handle_alive(int kind, void* h);
// true when the CUDA runtime can still be used on `device` (false during process teardown: cudaErrorCudartUnloading,
// destroyed primary context); destroy paths then only release host memory
bool pe_cuda_usable(int device);

int pe_set_error(int code, const char* msg);
int pe_fail(int code, const char* fmt, ...);
