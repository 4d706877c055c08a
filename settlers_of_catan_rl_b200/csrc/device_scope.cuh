// Scoped current-device switch for the C ABI: every entry point runs on the device of its handle / its buffers and puts the
// caller's current device back when it returns (the library must not change the process-wide current device under PyTorch).
#pragma once
#include <cuda_runtime.h>

namespace catanb {

struct DeviceScope {
  int prev = -1;
  bool switched = false;
  // make `device` current; 0 on success
  int enter(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) return -1;
    if (device >= 0 && prev != device) {
      if (cudaSetDevice(device) != cudaSuccess) return -1;
      switched = true;
    }
    return 0;
  }
  // make the device that holds `ptr` current (an entry point without a handle: the buffer says where the kernel must run)
  // (skipped while `stream` is capturing: a capture already runs under the device of its stream, and no query may disturb it)
  int enter_for(const void* ptr, cudaStream_t stream = nullptr) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (stream != nullptr && cudaStreamIsCapturing(stream, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) return 0;
    cudaPointerAttributes a;
    if (ptr == nullptr || cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return enter(-1); }
    return enter(a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged ? a.device : -1);
  }
  ~DeviceScope() { if (switched) cudaSetDevice(prev); }
  DeviceScope() = default;
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
};

}  // namespace catanb
