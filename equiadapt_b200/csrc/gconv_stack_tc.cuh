// Interface between gconv_stack.cu (planning, packing, finish kernel) and the tcgen05 stack kernel (gconv_stack_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace eqb {

struct TcArgs {
    const float *x;              // (B, cin, H, W) pre-network image
    int B, cin, H, W, ksz, Wo, P;
    int K0, K0pad, N, Npad;      // lift K = cin*k*k (padded to 32), N = Cout*|G|, Npad = stride of bias / S_part rows
    const float *bias1, *bias2;  // expanded biases (Npad)
    const unsigned char *wpack;  // header + fp16 UMMA weight images (tc_pack)
    const float *absmax;         // (B) device floats: max |x| of every image (tc_absmax or eqb_crop_resize_aa_absmax)
    double *S_part;              // [B][chunks][Npad] channel sums of the last executed layer
    int tiles, chunks, tiles_per_chunk;  // 128-pixel tiles per image, grouped into chunks (= work items)
    int tiles2, chunks2, tiles_per_chunk2;  // the same in 256-pixel pair-tiles (CTA-pair kernel)
    int epi1_groups, lift_early;         // pipeline shape (set by tc_launch)
    long long *trace;            // debug timeline (eqb_debug_stack_trace): [tile][64 slots] of clock64() from cluster 0 / CTA 0, or null
    int trace_tiles;
    int cta_waits;               // pair2 kernel: issuer polls with CTA-scope acquire (knob)
    int lift_at;                 // pair2 kernel: K atom of the 1x1 GEMM behind which pass 0 of the next tile's lift is queued
    int tail_n;                          // pair2: 1x1 MMAs of an image's last tile shrunk to its valid pixels
    int epi1_split;                      // pair2: both epilogue-1 groups convert every atom (16 channels each)
    int lift_after_gemm, epi2_pipelined; // CTA-pair kernel: lift of tile t+1 queued behind the last 1x1 MMA of tile t; LDTM of chunk c+1 in flight while chunk c is summed
};

bool tc_eligible(int N, int K0, int n_gemm);
bool tc_use_pair(int N, int K0);       // the CTA-pair kernel (cta_group::2, stationary weights) covers this shape
size_t tc_pack_bytes(int N, int K0);
// Wt0 [K0pad][Npad], Wt1 [Npad][Npad]: K-major fp32 operands built by the filter-orbit kernels; bias1: expanded (Npad)
int tc_pack(const float *Wt0, int K0, const float *Wt1, const float *bias1, int Npad, int N, unsigned char *out,
            cudaStream_t st);
int tc_absmax(const float *x, int B, size_t n_per_image, float *absmax, cudaStream_t st);   // absmax[b] = max |x[b]|
int tc_launch(TcArgs a, cudaStream_t st);
// valid k x k convolution with N = 256 and cin * k * k <= 256 on the tensor pipe (training path); *handled = 0: shape not
// taken, run the SIMT kernel.  x_absmax / y_absmax: per-image maxima, (B) floats or null
int tc_pw_conv(const float *x, const float *w, const float *bias, const float *mask, float *y, int B, int cin, int H, int W,
               int N, int k, int relu, const float *x_absmax, float *y_absmax, cudaStream_t st, int *handled);
// ... and its weight gradient dw[n, c] = sum_{b, p} dy[b, n, p] x[b, c, p] (dw zeroed by the caller; split-K fp32 atomics)
int tc_pw_wgrad(const float *dy, const float *x, float *dw, float *dy_rowsum, int B, int cin, int H, int W, int N, int k,
                const float *dy_absmax, const float *x_absmax, cudaStream_t st, int *handled);   // dy_rowsum: (N), zeroed, or null
int tc_set_trace(long long *device_buffer, int tiles);   // next CTA-pair launches record a timeline (null: off)
int tc_last_stall(int *out5);  // {flag, block, warp, barrier id, parity} of the first pipeline stall that trapped

}  // namespace eqb
