// clearcam_b200 — shared host-side helpers: error reporting (never abort: the reference's caller
// supervises failures itself, clearcam.py:543-546), CUDA checks, driver entry points.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace cc {

enum : int {
  CC_OK = 0,
  CC_ERR_INVALID = -1,   // bad argument / unsupported shape
  CC_ERR_CUDA = -2,      // CUDA runtime / driver error
  CC_ERR_NOGPU = -3,     // no sm_100 device
  CC_ERR_STATE = -4,     // handle used in the wrong state
};

void set_error(const char* fmt, ...);
const char* last_error();

#define CC_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::cc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return ::cc::CC_ERR_CUDA;                                                              \
    }                                                                                        \
  } while (0)

#define CC_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      ::cc::set_error(__VA_ARGS__);    \
      return ::cc::CC_ERR_INVALID;     \
    }                                  \
  } while (0)

// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency, so the
// library loads on a box without a driver and fails only when a GPU op is actually requested).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

int device_sm_count();  // cached; <=0 if no usable device

}  // namespace cc
