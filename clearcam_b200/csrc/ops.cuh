// clearcam_b200 — parameter blocks + launchers of the non-GEMM kernels on the hot path.
// All activations are NHWC; a tensor is addressed as (base pointer, channel stride `cs` = channels per
// pixel of the underlying buffer, channel offset `co`, channel count) so concat/chunk are free.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace cc {

struct DirectConvParams {
  const __nv_bfloat16* in; int in_cs, in_co, Cin;
  int N, Hin, Win;        // logical input extent used for bounds (may be smaller than the stored buffer)
  int Hbuf, Wbuf;         // stored buffer extent (row/col pitch)
  int k, stride, pad, groups;
  const __nv_bfloat16* w; // [Cout][k][k][Cin/groups]
  const float* bias;
  void* out; int out_cs, out_co, Cout, out_f32;
  int Hout, Wout;
  int act;
  const void* res; int res_cs, res_co;
};
int conv_direct_launch(const DirectConvParams& p, cudaStream_t stream);

}  // namespace cc
