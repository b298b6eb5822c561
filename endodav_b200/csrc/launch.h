// Launch bookkeeping shared by all host-side launchers.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/endodav_b200.h"

namespace edv {

struct Launch {
  cudaStream_t stream = nullptr;
  int count = 0;         // kernels launched
  int status = EDV_OK;   // first error
  std::string err;

  void fail(int code, const std::string& what) {
    if (status == EDV_OK) {
      status = code;
      err = what;
    }
  }
  void check(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fail(EDV_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    ++count;
  }
  bool ok() const { return status == EDV_OK; }
};

}  // namespace edv
