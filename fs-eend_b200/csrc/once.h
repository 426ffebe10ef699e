// One-time per-DEVICE launch setup: cudaFuncSetAttribute (dynamic shared-memory opt-in) and device buffers are per device,
// so a process that drives several GPUs must repeat them on each one (a plain `static bool` did them on the first only).
#pragma once
#include <cuda_runtime.h>

namespace fseend {

struct PerDeviceOnce {
  bool done[64] = {};
  int device = 0;
  bool first() {      // true the first time this call site runs on the current device
    cudaGetDevice(&device);
    device &= 63;
    if (done[device]) return false;
    done[device] = true;
    return true;
  }
};

}  // namespace fseend
