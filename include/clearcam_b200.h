/* clearcam_b200 — C-ABI of the B200-native per-frame vision hot path of roryclear/clearcam.
 *
 * The reference has no FFI: its boundary is Python call signatures.  This header is the
 * C-ABI that sits UNDER those signatures; clearcam_b200/detection/yolov9.py and clearcam_b200/models/objects.py
 * bind it with ctypes and keep the reference's names/arguments.  Conventions:
 *   - every function returns int: 0 = ok, <0 = error (cc_last_error() gives the message); nothing aborts/throws
 *     (the reference's caller supervises failures itself: clearcam.py:543-546);
 *   - all pointers named d_* are DEVICE pointers owned by the caller (torch tensors on the Python side);
 *     h_* are host pointers; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - calls are stream-ordered and never synchronise; one handle per GPU, not thread-safe per handle
 *     (the reference funnels all model calls to one thread: clearcam.py:1214-1226);
 *   - activations are NHWC, bf16 unless stated; a tensor slice is (ptr, cs = channels per pixel of the
 *     underlying buffer, co = first channel, C = channel count).
 */
#ifndef CLEARCAM_B200_H
#define CLEARCAM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CC_ABI_VERSION 1

int cc_version(void);
const char* cc_last_error(void);
/* number of SMs of the current device if it is sm_100 (B200), else <0 */
int cc_device_check(void);

/* activation codes */
#define CC_ACT_NONE 0
#define CC_ACT_SILU 1      /* x*sigmoid(x): detection/yolov9.py:38 */
#define CC_ACT_GELU_TANH 2 /* tinygrad Tensor.gelu(): models/objects.py:125,179 */

/* ---- kernel-level ops (used by the graphs below and exposed for the parity tests) ---- */

/* Conv2d(bias) [+act] [+residual], k in {1,3}, stride in {1,2}, pad = k/2, NHWC bf16 in, bf16|fp32 out.
 * Replaces nn.Conv2d + .silu() of detection/yolov9.py:33-38 (and the bare nn.Conv2d of :173,:186,:224).
 * d_w: bf16 [Cout][k][k][Cin/groups]; d_bias: fp32 [Cout] or NULL; d_res: same dtype/shape class as out or NULL.
 * impl: 0 = auto (tcgen05 implicit GEMM when the shape allows, else direct), 1 = force tcgen05, 2 = force direct.
 * bn: tcgen05 N-tile override (0 = heuristic). */
int cc_conv2d(const void* d_in, int N, int Hin, int Win, int in_cs, int in_co, int Cin,
              const void* d_w, const float* d_bias, int Cout, int k, int stride, int groups,
              void* d_out, int out_cs, int out_co, int out_f32, int act,
              const void* d_res, int res_cs, int res_co, int impl, int bn, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLEARCAM_B200_H */
