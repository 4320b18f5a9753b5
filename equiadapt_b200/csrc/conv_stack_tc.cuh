// tcgen05 implicit-GEMM layers of the a7 conv stack (conv_stack_tc.cu), called from conv_stack.cu.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace eqb {

// per-layer device record (8 floats): operand scales and value bounds, chained layer to layer on the device
enum { LAY_INBOUND = 0, LAY_SIN = 1, LAY_SW = 2, LAY_CINV = 3, LAY_OUTBOUND = 4, LAY_SOUT = 5, LAY_FLOATS = 8 };

bool ctc_eligible(int Npad, int num_layers);
int ctc_atom_channels(int Cpad);       // channels per K atom for an operand padded to Cpad channels: 16 or 32
size_t ctc_pack_bytes(int Npad, int Cpad, int k);
// w (N, K) fp32 filter, vecs = [bias | scale | shift] (3 x Npad); tensor_in: this layer consumes fp16 hi/lo operands;
// rowstat: 2 * N floats of device scratch
int ctc_layer_stats(const float *w, const float *vecs, int N, int Npad, int K, const float *in_bound_ptr, int in_stride, int B,
                    int tensor_in, float *lay, float *nxt, float *rowstat, cudaStream_t st);
// x (B, C, H, W) fp32 -> fp16 hi / lo NHWC with C padded to Cpad, scaled by pow2(*absmax)
int ctc_input_split(const float *x, const float *absmax, int B, int C, int H, int W, int Cpad, __half *hi, __half *lo,
                    cudaStream_t st);
int ctc_pack(const float *w, const float *lay, int N, int C, int Cpad, int k, int Npad, unsigned char *out, cudaStream_t st);
// in_hi / in_lo: (B, H, W, Cpad) fp16; out: fp32 (B, N, Ho, Wo) or the fp16 pair (B, Ho, Wo, Cpad_out)
int ctc_conv_layer(const __half *in_hi, const __half *in_lo, int B, int Cpad, int H, int W, int k, const unsigned char *wpack,
                   const float *vecs, const float *lay, int N, int Npad, int relu, float *out_nchw, __half *out_hi,
                   __half *out_lo, int Cpad_out, cudaStream_t st);

}  // namespace eqb
