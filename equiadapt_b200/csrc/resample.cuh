// Shared declarations of the group-action resampling kernels (resample.cu: generic path, resample_tma.cu: TMA path).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace eqb {

constexpr int TILE = 32;
constexpr int BB = 48;       // max footprint side
constexpr int PITCH = 49;    // odd pitch
constexpr int THREADS = 256;
constexpr int PIX = TILE * TILE / THREADS;  // 4 pixels per thread

enum { MODE_CANON = 0, MODE_INV_SCALAR = 1, MODE_INV_REGULAR = 2, MODE_ORBIT = 3, MODE_AFFINE = 4 };

struct TileGeom;   // per-(group element, tile) geometry of the TMA kernel (resample_tma.cu)

struct ResampleArgs {
    const float *src;
    float *dst;
    const int32_t *idx;  // per source sample group element (unused for MODE_ORBIT)
    int B;               // source samples
    int C, Hs, Ws, Hd, Wd;
    int N, reflect, G;
    signed char roll[64];  // regular-rep channel shift per rotation index (see regular_roll_shift)
    int mode;
    int pad;        // >=0: taps inside [-pad, size-1+pad] are replicate-clamped, zero beyond
    double ox, oy;  // dst pixel -> coordinate relative to the rotation centre: u = xd + ox
    int tiles_x, tiles_y;
    int sample0;    // first destination sample of this launch (TMA path: the batch is chunked over gridDim.z)
    // (cos, sin) of 2*pi*r/N for r < N <= 16, filled on the host (exact 0 / +-1 at quarter turns); has_cs = 0 ->
    // the kernels evaluate sincospi themselves
    int has_cs;
    double cs[32];
    // MODE_AFFINE (continuous groups): per-sample 2x2 matrices instead of a group index.  mats_forward = 1: the
    // matrix maps source to destination about the centre (warp_affine convention) and is inverted here;
    // 0: it maps destination to source directly (affine_grid convention).  refl (may be null): per-sample 0/1,
    // the source is mirrored first.  (scx, scy): rotation centre in source pixel coordinates.
    const float *mats;
    const float *refl;
    int mats_forward;
    double scx, scy;
    // TMA path, discrete modes: [G][tiles_y * tiles_x] table built once per call shape (null: every CTA computes its own)
    const TileGeom *geom;
};

// fills tiles_x / tiles_y / cs
void finish_args(ResampleArgs &a);

__device__ __forceinline__ void group_cs(const ResampleArgs &a, int r, double sign, double &c, double &s) {
    if (a.has_cs) {
        c = a.cs[2 * r];
        s = sign * a.cs[2 * r + 1];
    } else {
        rot_cs(r, a.N, sign, c, s);
    }
}

// TMA-staged implementation (resample_tma.cu).  Returns 0 when it launched, <0 / >0 on error, and sets
// *handled = 0 (launching nothing) when the tensor does not meet the TMA layout rules (16-byte aligned base,
// row pitch a multiple of 16 bytes, at least one 52x48 box) so that the caller takes the generic kernel.
int launch_resample_tma(const ResampleArgs &a, int n_dst_samples, cudaStream_t st, const char *what, int *handled);

// 3-D tensor map over a contiguous fp32 (planes, H, W) tensor with a (box_w, box_h, 1) box (resample_tma.cu)
int make_plane_map(CUtensorMap *m, const float *src, int W, int H, long long planes, int box_w, int box_h,
                   CUtensorMapSwizzle swz);

}  // namespace eqb
