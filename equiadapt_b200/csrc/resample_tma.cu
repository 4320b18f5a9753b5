// TMA-staged group-action resampling (the HBM-bound kernel behind eqb_warp_canonicalize / eqb_warp_invert /
// eqb_orbit_expand when the source tensor meets the TMA layout rules; resample.cu holds the generic kernel).
//
// Same arithmetic contract as resample.cu (see its header): each output pixel is a 4-tap bilinear sample at
//   src = centre + A (dst - centre),
// taps replicate-clamped into the image inside the padded extent and zero beyond it
// (discrete_group.py:207-215, images/utils.py:57-64, discrete_group.py:401-409 of the reference).
//
// B200 design.  One CTA = one 32x32 output tile of one sample, all channels (CG planes per pass):
//   * the source footprint is fetched by the TMA unit (cp.async.bulk.tensor.3d, one instruction per plane,
//     completion on an mbarrier) while all 256 threads compute their per-pixel taps - no LDG/STS staging
//     instructions, no registers held for loads in flight;
//   * quarter turns / mirrors with integral source coordinates ("exact" tiles) are pure permutations: the box
//     is the 32x32 source tile itself, landed with the 128-byte swizzle so that both the row-wise and the
//     column-wise (transposing) reads are at most 4-way bank conflicted, one LDS per output element;
//   * every other tile uses a 52x48 box (bounding box of the rotated tile + 1 tap, start rounded down to the
//     16-byte boundary the TMA unit requires), coordinates are evaluated
//     in fp32 RELATIVE to the box origin (|coord| < 64 -> 4e-6 px resolution; the box origin itself comes from
//     fp64), 4 LDS + 4 FMA per output element with offsets/weights shared by all channels;
//   * each warp stores full 128-byte rows (st.global.L1::no_allocate).
// Algorithmic traffic per sample: 1 read + 1 write of the image (SURVEY.md 8d "W"); the 2x footprint overlap of
// rotated tiles is served by L2.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "resample.cuh"

namespace eqb {

// TMA rule measured on B200 (tools/tma_probe.cu): with INTERLEAVE_NONE the box must START on a 16-byte boundary
// of global memory, i.e. its x coordinate must be a multiple of 4 floats (a misaligned start raises "illegal
// instruction" at the UTMALDG).  The bilinear box is therefore 47 (+1 tap) + 3 (alignment slack) -> 52 wide.
constexpr int BOXW = 52, BOXH = 48;            // bilinear source box (floats)
constexpr int BOX_BYTES = BOXW * BOXH * 4;     // 9984 bytes landed per plane
constexpr int PLANE_BYTES = 10 * 1024;         // plane stride in shared memory: keeps every plane 1024-byte aligned
constexpr int TILE_BYTES = TILE * TILE * 4;    // exact path: 32 x 32 box, 128-byte rows, SWIZZLE_128B

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(x), "r"(y), "r"(z)
        : "memory");
}

// Geometry of one tile, computed once per CTA by warp 0 and broadcast through shared memory
// (it costs ~400 instructions incl. fp64: done by all 8 warps it was 45 % of the kernel's issue slots).
struct TileGeom {
    int exact, interior, box_x, box_y, y_lo;
    int cx_lo, cx_hi, bhm1;             // bilinear: clamp range of the taps inside the box
    float bx, by, f00, f01, f10, f11;   // bilinear: box-relative source coordinate of pixel (tx0,ty0) and the matrix
    int i00, i01, i10, i11, sx0, sy0;   // exact: integer matrix and box-relative source of pixel (tx0,ty0)
    int r, plane0;                      // rotation index (regular-rep roll) and first source plane
};

__device__ __forceinline__ int regular_src_channel(const ResampleArgs &a, int cs, int r) {
    // out[:, f*G+g] = in[:, f*G + src_g(g)]   (roll_by_gather, images/utils.py:8-29,66-77)
    const int f = cs / a.G, g = cs - f * a.G, sh = a.roll[r];
    int sg;
    if (g < a.N) sg = (g - sh + a.N) % a.N;
    else sg = a.N + (g - a.N + sh) % a.N;
    return f * a.G + sg;
}

// Geometry of the 32x32 destination tile at (tx0, ty0) under group element g (discrete modes) or the matrix of sample
// `sample_s` (MODE_AFFINE): fp64 throughout (~400 instructions), evaluated once per (g, tile) into a table by
// tile_geometry_kernel instead of by every CTA (it was ~1.2 us of every CTA's ~4.5 us life).
__device__ TileGeom tile_geometry(const ResampleArgs &a, int g, int sample_s, int tx0, int ty0) {
    TileGeom geo;
    int r = 0, mirror_src = 0, mirror_dst = 0;
    double sign = 1.0;
    if (a.mode == MODE_ORBIT) {
        r = g % a.N;
        mirror_dst = g >= a.N;  // rotate, THEN hflip (discrete_group.py:404-406)
        sign = -1.0;
    } else if (a.mode != MODE_AFFINE) {
        r = g % a.N;
        const int refl = g >= a.N;
        if (a.mode == MODE_CANON) {
            mirror_src = refl;  // hflip, THEN rotate(-theta) (discrete_group.py:209-213)
            sign = -1.0;
        } else {
            mirror_dst = a.reflect && !refl;  // images/utils.py:59-64 (reference quirk A.4-2)
            sign = 1.0;
        }
    }
    double a00, a01, a10, a11;
    double cx = 0.5 * (a.Ws - 1), cy = 0.5 * (a.Hs - 1);
    if (a.mode == MODE_AFFINE) {   // per-sample 2x2 matrix (continuous groups), see resample.cu
        const float *m = a.mats + 4 * (size_t)sample_s;
        const double m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
        if (a.mats_forward) {
            const double det = m00 * m11 - m01 * m10;
            a00 = m11 / det; a01 = -m01 / det; a10 = -m10 / det; a11 = m00 / det;
        } else {
            a00 = m00; a01 = m01; a10 = m10; a11 = m11;
        }
        cx = a.scx; cy = a.scy;
        if (a.refl && a.refl[sample_s] > 0.5f) {
            a00 = -a00; a01 = -a01;
            cx = (double)(a.Ws - 1) - cx;
        }
    } else {
        double c, s;
        group_cs(a, r, sign, c, s);
        a00 = c; a01 = -s; a10 = s; a11 = c;
        if (mirror_dst) { a00 = -a00; a10 = -a10; }
        if (mirror_src) { a00 = -a00; a01 = -a01; }
    }
    // a signed permutation matrix (quarter turns / mirrors): candidates for the exact path
    const bool unit = (fabs(a00) == 1.0 && a01 == 0.0 && a10 == 0.0 && fabs(a11) == 1.0) ||
                      (a00 == 0.0 && fabs(a01) == 1.0 && fabs(a10) == 1.0 && a11 == 0.0);
    // ---- source footprint of the tile ----------------------------------------------------------
    const int tw = min(TILE, a.Wd - tx0), th = min(TILE, a.Hd - ty0);
    const double u0 = (double)tx0 + a.ox, v0 = (double)ty0 + a.oy;
    const double xs_org = cx + a00 * u0 + a01 * v0, ys_org = cy + a10 * u0 + a11 * v0;  // source of (tx0,ty0)
    const double dw = (double)(tw - 1), dh = (double)(th - 1);
    const double xmin = xs_org + fmin(a00 * dw, 0.0) + fmin(a01 * dh, 0.0);
    const double xmax = xs_org + fmax(a00 * dw, 0.0) + fmax(a01 * dh, 0.0);
    const double ymin = ys_org + fmin(a10 * dw, 0.0) + fmin(a11 * dh, 0.0);
    const double ymax = ys_org + fmax(a10 * dw, 0.0) + fmax(a11 * dh, 0.0);
    const int fxmin = (int)floor(xmin), fxmax = (int)floor(xmax), fymin = (int)floor(ymin), fymax = (int)floor(ymax);
    // exact tile: quarter turn (cos / sin are exact 0 / +-1 there), integral source coordinates, no clamping,
    // source tile starting on a 16-byte boundary (always true for square images whose side is a multiple of 4)
    // footprint of any rotation fits the 52 x 48 box; a matrix that expands the tile (not a group element) cannot
    const bool fits = (xmax - xmin) <= 46.0 && (ymax - ymin) <= 46.0 && xmin > -1e9 && xmax < 1e9 && ymin > -1e9 && ymax < 1e9;
    const bool exact = unit && xs_org == floor(xs_org) && ys_org == floor(ys_org) && fxmin >= 0 &&
                       fxmax <= a.Ws - 1 && fymin >= 0 && fymax <= a.Hs - 1 && (fxmin & 3) == 0;
    // interior tile: every tap, even one pixel beyond the fp64 footprint (fp32 rounding of the per-pixel
    // coordinates may floor() to the neighbour), lies inside the image -> no clamping in the pixel loop
    const bool interior = !exact && fits && fxmin >= 1 && fxmax + 2 <= a.Ws - 1 && fymin >= 1 && fymax + 2 <= a.Hs - 1;
    int box_x, box_y, y_lo = 0, cx_lo = 0, cx_hi = 0, bhm1 = 0;
    if (exact) {
        box_x = fxmin; box_y = fymin;
    } else if (interior) {
        box_x = (fxmin - 1) & ~3; box_y = fymin - 1; y_lo = box_y;     // <= 1 + 46 + 1 + 3 = 51 < 52 wide
    } else {
        const int x_lo = min(max(fxmin, 0), a.Ws - 1);
        y_lo = min(max(fymin, 0), a.Hs - 1);
        box_x = x_lo & ~3; box_y = y_lo;                               // 16-byte aligned start
        cx_lo = x_lo - box_x;                                          // taps are clamped into [cx_lo, cx_hi]
        cx_hi = min(min(max(fxmax + 1, 0), a.Ws - 1) - box_x, BOXW - 1);   // <= 46 + 3 for a group element
        bhm1 = min(min(max(fymax + 1, 0), a.Hs - 1) - y_lo, BOXH - 1);     // <= 46
        cx_lo = min(cx_lo, cx_hi);
    }
    geo.exact = exact; geo.interior = interior; geo.box_x = box_x; geo.box_y = box_y; geo.y_lo = y_lo;
    geo.cx_lo = cx_lo; geo.cx_hi = cx_hi; geo.bhm1 = bhm1;
    geo.bx = (float)(xs_org - (double)box_x); geo.by = (float)(ys_org - (double)box_y);
    geo.f00 = (float)a00; geo.f01 = (float)a01; geo.f10 = (float)a10; geo.f11 = (float)a11;
    geo.i00 = (int)a00; geo.i01 = (int)a01; geo.i10 = (int)a10; geo.i11 = (int)a11;
    geo.sx0 = (int)xs_org - box_x; geo.sy0 = (int)ys_org - box_y;
    geo.r = r; geo.plane0 = 0;
    return geo;
}

__global__ void tile_geometry_kernel(const ResampleArgs a, TileGeom *__restrict__ table) {
    const int tiles = a.tiles_x * a.tiles_y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.G * tiles) return;
    const int g = t / tiles, tile = t - g * tiles;
    table[t] = tile_geometry(a, g, 0, (tile % a.tiles_x) * TILE, (tile / a.tiles_x) * TILE);
}

// One channel pass (nc <= CG planes starting at channel c0) of one 32x32 destination tile whose source box(es) sit in shared
// memory at sm0: shared by the one-tile-per-CTA kernel and the persistent kernel.  8 warps: lane = x, warp = first row.
template <int CG, bool ZERO>
__device__ __forceinline__ void resample_tile_pass(const ResampleArgs &a, const TileGeom &geo, uint32_t sm0, int sample_d,
                                                   int tx0, int ty0, int c0, int nc, int warp, int lane) {
    const bool exact = geo.exact != 0;
    const int xd = tx0 + lane;
    if (xd >= a.Wd) return;
    const int plane_d = a.Hd * a.Wd;                       // host guarantees C * Hd * Wd < 2^31
    float *dst_px = a.dst + (size_t)sample_d * a.C * plane_d + (size_t)(ty0 + warp) * a.Wd + xd;
    const int row_step = (THREADS / 32) * a.Wd;            // this thread's pixels are 8 rows apart
    const int rows_left = a.Hd - ty0 - warp;               // pixel p is inside the image iff p*8 < rows_left
    {
        float *dp = dst_px + (size_t)c0 * plane_d;
        if (exact) {
            const int sxl = geo.sx0 + geo.i00 * lane + geo.i01 * warp, syl = geo.sy0 + geo.i10 * lane + geo.i11 * warp;
            const int dsx = geo.i01 * (THREADS / 32), dsy = geo.i11 * (THREADS / 32);
#pragma unroll
            for (int p = 0; p < PIX; ++p) {
                if (p * (THREADS / 32) < rows_left) {
                    const int sx = (sxl + p * dsx) & (TILE - 1), sy = (syl + p * dsy) & (TILE - 1);
                    // SWIZZLE_128B: 16-byte chunk index XOR (row & 7); rows are 128 bytes, planes 1024-byte aligned
                    const uint32_t o = sm0 + 4u * (uint32_t)(sy * TILE + ((((sx >> 2) ^ (sy & 7)) << 2) | (sx & 3)));
#pragma unroll
                    for (int cc = 0; cc < CG; ++cc)
                        if (cc < nc) st_stream(dp + p * row_step + cc * plane_d, lds_f32(o + cc * PLANE_BYTES));
                }
            }
        } else {
            // taps of one pixel at a time (offsets / weights shared by the CG channels); the 4 pixels of a thread
            // are 8 rows apart, so their source coordinates advance by 8 * (f01, f11)
            float xr = fmaf(geo.f00, (float)lane, fmaf(geo.f01, (float)warp, geo.bx));
            float yr = fmaf(geo.f10, (float)lane, fmaf(geo.f11, (float)warp, geo.by));
            const float dxr = geo.f01 * (float)(THREADS / 32), dyr = geo.f11 * (float)(THREADS / 32);
            const bool interior = geo.interior != 0;
            const int cx_lo = geo.cx_lo, cx_hi = geo.cx_hi, bhm1 = geo.bhm1;
#pragma unroll
            for (int p = 0; p < PIX; ++p, xr += dxr, yr += dyr) {
                if (p * (THREADS / 32) >= rows_left) break;
                const float xf = floorf(xr), yf = floorf(yr);
                const float fx = xr - xf, fy = yr - yf;
                const int x0 = (int)xf, y0 = (int)yf;
                float wx0 = 1.f - fx, wx1 = fx, wy0 = 1.f - fy, wy1 = fy;
                if (ZERO && !interior) {   // (every tap of an interior tile lies inside the image: nothing to zero)
                    // absolute tap coordinates against the padded extent [-pad, size-1+pad]
                    const int ax0 = x0 + geo.box_x, ay0 = y0 + geo.y_lo;
                    const int lo = -a.pad, hx = a.Ws - 1 + a.pad, hy = a.Hs - 1 + a.pad;
                    if (ax0 < lo || ax0 > hx) wx0 = 0.f;
                    if (ax0 + 1 < lo || ax0 + 1 > hx) wx1 = 0.f;
                    if (ay0 < lo || ay0 > hy) wy0 = 0.f;
                    if (ay0 + 1 < lo || ay0 + 1 > hy) wy1 = 0.f;
                }
                const float w00 = wy0 * wx0, w01 = wy0 * wx1, w10 = wy1 * wx0, w11 = wy1 * wx1;
                float *dpp = dp + p * row_step;
                if (interior) {
                    const uint32_t o = sm0 + 4u * (uint32_t)(y0 * BOXW + x0);
#pragma unroll
                    for (int cc = 0; cc < CG; ++cc) {
                        if (cc < nc) {
                            const uint32_t oc = o + cc * PLANE_BYTES;
                            // same tap order as ATen grid_sample: nw, ne, sw, se
                            float v = lds_f32(oc) * w00;
                            v = fmaf(lds_f32(oc + 4), w01, v);
                            v = fmaf(lds_f32(oc + 4 * BOXW), w10, v);
                            v = fmaf(lds_f32(oc + 4 * BOXW + 4), w11, v);
                            st_stream(dpp + cc * plane_d, v);
                        }
                    }
                } else {
                    const int cx0 = min(max(x0, cx_lo), cx_hi) * 4, cx1 = min(max(x0 + 1, cx_lo), cx_hi) * 4;
                    const uint32_t cy0 = sm0 + (uint32_t)(min(max(y0, 0), bhm1) * (4 * BOXW));
                    const uint32_t cy1 = sm0 + (uint32_t)(min(max(y0 + 1, 0), bhm1) * (4 * BOXW));
                    const uint32_t o00 = cy0 + cx0, o01 = cy0 + cx1, o10 = cy1 + cx0, o11 = cy1 + cx1;
#pragma unroll
                    for (int cc = 0; cc < CG; ++cc) {
                        if (cc < nc) {
                            float v = lds_f32(o00 + cc * PLANE_BYTES) * w00;
                            v = fmaf(lds_f32(o01 + cc * PLANE_BYTES), w01, v);
                            v = fmaf(lds_f32(o10 + cc * PLANE_BYTES), w10, v);
                            v = fmaf(lds_f32(o11 + cc * PLANE_BYTES), w11, v);
                            st_stream(dpp + cc * plane_d, v);
                        }
                    }
                }
            }
        }
    }
}

template <int CG, bool ZERO>
__global__ void __launch_bounds__(THREADS, 6) resample_tma_kernel(const __grid_constant__ CUtensorMap map_box,
                                                                  const __grid_constant__ CUtensorMap map_tile,
                                                                  const __grid_constant__ ResampleArgs a) {
    extern __shared__ unsigned char smem_raw[];
    // [CG planes, 10 KB apart, 1024-aligned][mbarrier][TileGeom]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *aligned = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar = base + CG * PLANE_BYTES;
    TileGeom &geo = *reinterpret_cast<TileGeom *>(aligned + CG * PLANE_BYTES + 16);

    const int sample_d = blockIdx.z + a.sample0;  // destination sample
    const int ty0 = blockIdx.y * TILE, tx0 = blockIdx.x * TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (warp == 0) {
        if (lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // group element of this sample -> geometry of the tile: from the per-call-shape table (one 80-byte read) for the
            // discrete modes, computed here for per-sample matrices
            int sample_s = sample_d, g = 0;
            if (a.mode == MODE_ORBIT) {
                g = sample_d / a.B;
                sample_s = sample_d - g * a.B;
            } else if (a.mode != MODE_AFFINE) {
                g = min(max(a.idx[sample_s], 0), a.G - 1);
            }
            if (a.geom) geo = a.geom[(size_t)g * (a.tiles_x * a.tiles_y) + blockIdx.y * a.tiles_x + blockIdx.x];
            else geo = tile_geometry(a, g, sample_s, tx0, ty0);
            geo.plane0 = sample_s * a.C;   // first source plane of this sample in the (W,H,B*C) tensor map
            // first pass of the TMA loads goes out before anybody else needs the geometry
            const int nc = min(CG, a.C);
            mbar_expect_tx(bar, (uint32_t)nc * (geo.exact ? TILE_BYTES : BOX_BYTES));
            for (int cc = 0; cc < nc; ++cc) {
                const int cs = a.mode == MODE_INV_REGULAR ? regular_src_channel(a, cc, geo.r) : cc;
                tma_load_3d(base + cc * PLANE_BYTES, geo.exact ? &map_tile : &map_box, bar, geo.box_x, geo.box_y, geo.plane0 + cs);
            }
        }
    }
    __syncthreads();

    const uint32_t sm0 = base;
    uint32_t parity = 0;
    for (int c0 = 0; c0 < a.C; c0 += CG) {
        const int nc = min(CG, a.C - c0);
        if (c0) {
            __syncthreads();  // every thread is done reading the planes of the previous pass
            if (threadIdx.x == 0) {
                mbar_expect_tx(bar, (uint32_t)nc * (geo.exact ? TILE_BYTES : BOX_BYTES));
                for (int cc = 0; cc < nc; ++cc) {
                    const int cs = a.mode == MODE_INV_REGULAR ? regular_src_channel(a, c0 + cc, geo.r) : c0 + cc;
                    tma_load_3d(base + cc * PLANE_BYTES, geo.exact ? &map_tile : &map_box, bar, geo.box_x, geo.box_y,
                                geo.plane0 + cs);
                }
            }
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        resample_tile_pass<CG, ZERO>(a, geo, sm0, sample_d, tx0, ty0, c0, nc, warp, lane);
    }
}

// Persistent variant: 3 CTAs per SM, each with a producer warp (warp 8: geometry of the next work unit from the table, TMA
// loads into the other buffer) and 8 consumer warps (the pass above).  A work unit = (destination tile, channel pass); units
// are dealt round-robin, tile x fastest, so the CTAs of a wave work on neighbouring tiles and their overlapping footprints
// meet in L2.  The one-tile-per-CTA kernel above leaves the load/store pipe idle while a CTA computes its geometry, waits
// for its box and winds down (6 CTAs per SM overlap only statistically: LSU 63 % busy, 0.70 of the HBM peak on all-bilinear
// tiles); here the box of unit k+1 lands while unit k is computed.
constexpr int P_THREADS = THREADS + 32;
struct UnitInfo {
    TileGeom geo;
    int sample_d, tx0, ty0, c0, nc, pad_;
};

template <int CG, bool ZERO>
__global__ void __launch_bounds__(P_THREADS, 3) resample_tma_persistent_kernel(const __grid_constant__ CUtensorMap map_box,
                                                                               const __grid_constant__ CUtensorMap map_tile,
                                                                               const __grid_constant__ ResampleArgs a,
                                                                               const int total_units, const int passes) {
    extern __shared__ unsigned char smem_raw[];
    // [2 buffers x CG planes, 10 KB apart, 1024-aligned][full[2], empty[2]][UnitInfo[2]]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *aligned = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + 2 * CG * PLANE_BYTES;
    UnitInfo *info = reinterpret_cast<UnitInfo *>(aligned + 2 * CG * PLANE_BYTES + 32);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(bars, 1);          // full[0]
        mbar_init(bars + 8, 1);      // full[1]
        mbar_init(bars + 16, 8);     // empty[0]: one arrival per consumer warp
        mbar_init(bars + 24, 8);     // empty[1]
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tiles_per_sample = a.tiles_x * a.tiles_y;
    if (warp == 8) {
        if (lane == 0) {
            int k = 0;
            for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++k) {
                const int s = k & 1;
                mbar_wait(bars + 16 + 8 * s, (((uint32_t)k >> 1) & 1u) ^ 1u);     // consumers are done with this buffer
                const int pass = unit % passes, tile = unit / passes;
                const int sample_d = tile / tiles_per_sample, t2 = tile - sample_d * tiles_per_sample;
                const int tyi = t2 / a.tiles_x, txi = t2 - tyi * a.tiles_x;
                int sample_s = sample_d, g = 0;
                if (a.mode == MODE_ORBIT) {
                    g = sample_d / a.B;
                    sample_s = sample_d - g * a.B;
                } else if (a.mode != MODE_AFFINE) {
                    g = min(max(a.idx[sample_s], 0), a.G - 1);
                }
                UnitInfo &u = info[s];
                if (a.geom) u.geo = a.geom[(size_t)g * tiles_per_sample + t2];
                else u.geo = tile_geometry(a, g, sample_s, txi * TILE, tyi * TILE);
                u.geo.plane0 = sample_s * a.C;
                u.sample_d = sample_d; u.tx0 = txi * TILE; u.ty0 = tyi * TILE;
                u.c0 = pass * CG; u.nc = min(CG, a.C - pass * CG);
                const uint32_t full = bars + 8 * s, dst = base + s * CG * PLANE_BYTES;
                mbar_expect_tx(full, (uint32_t)u.nc * (u.geo.exact ? TILE_BYTES : BOX_BYTES));
                for (int cc = 0; cc < u.nc; ++cc) {
                    const int cs = a.mode == MODE_INV_REGULAR ? regular_src_channel(a, u.c0 + cc, u.geo.r) : u.c0 + cc;
                    tma_load_3d(dst + cc * PLANE_BYTES, u.geo.exact ? &map_tile : &map_box, full, u.geo.box_x, u.geo.box_y,
                                u.geo.plane0 + cs);
                }
            }
        }
    } else {
        int k = 0;
        for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++k) {
            const int s = k & 1;
            mbar_wait(bars + 8 * s, ((uint32_t)k >> 1) & 1u);
            const UnitInfo &u = info[s];
            resample_tile_pass<CG, ZERO>(a, u.geo, base + s * CG * PLANE_BYTES, u.sample_d, u.tx0, u.ty0, u.c0, u.nc, warp, lane);
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 16 + 8 * s) : "memory");
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        (void)cudaGetLastError();
        return (EncodeTiledFn)p;
    }();
    return fn;
}

int make_plane_map(CUtensorMap *m, const float *src, int W, int H, long long planes, int box_w, int box_h,
                          CUtensorMapSwizzle swz) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        return (int)cudaErrorNotSupported;
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    const cuuint64_t gstride[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * (cuuint64_t)H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)src, gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (W=%d H=%d planes=%lld box=%dx%d)", (int)r, W, H, planes,
                  box_w, box_h);
        return (int)cudaErrorInvalidValue;
    }
    return 0;
}

template <int CG, bool ZERO>
static int launch_cfg(const CUtensorMap &mb, const CUtensorMap &mt, ResampleArgs a, int n_dst_samples,
                      cudaStream_t st) {
    const char *pe = getenv("EQB_WARP_PERSISTENT");
    const int passes = (a.C + CG - 1) / CG;
    const long long units = (long long)a.tiles_x * a.tiles_y * n_dst_samples * passes;
    // Persistent producer / consumer kernel where it measured faster: 3-channel images of >= 128 x 128 (canonicalize 132.9 ->
    // 122.8 us, invert 133.4 -> 125.5 us on all-bilinear C8 elements at 512 x 3 x 224 x 224).  With four channels per pass its two
    // buffers leave room for two CTAs per SM only (regular-representation invert: 315 -> 354 us) and on small images (orbit
    // expand at 96 x 96: 50 -> 58 us) the 24 instead of 48 warps per SM cost more than the prefetch gains.
    const bool want_persistent = pe ? pe[0] != '0' : (CG == 3 && a.Hs >= 128 && a.Ws >= 128);
    if (want_persistent && units < (1LL << 31)) {
        // the box of work unit k+1 lands while unit k is computed
        const size_t smem = (size_t)2 * CG * PLANE_BYTES + 1024 + 32 + 2 * sizeof(UnitInfo);
        static PerDeviceOnce configured_p;
        if (configured_p.first()) {
            EQB_CUDA(cudaFuncSetAttribute(resample_tma_persistent_kernel<CG, ZERO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
        }
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(3, (size_t)(227 * 1024) / (smem + 1024)));
        const long long grid = std::min<long long>(units, (long long)per_sm * num_sms());
        resample_tma_persistent_kernel<CG, ZERO><<<(unsigned)grid, P_THREADS, smem, st>>>(mb, mt, a, (int)units, passes);
        return 0;
    }
    const size_t smem = (size_t)CG * PLANE_BYTES + 1024 + 16 + sizeof(TileGeom);
    static PerDeviceOnce configured;
    if (configured.first()) {
        EQB_CUDA(cudaFuncSetAttribute(resample_tma_kernel<CG, ZERO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    }
    // grid = (tile x, tile y, sample): no integer division in the kernel; gridDim.z <= 65535 -> chunk the batch
    for (int z0 = 0; z0 < n_dst_samples; z0 += 32768) {
        a.sample0 = z0;
        const dim3 grid(a.tiles_x, a.tiles_y, std::min(32768, n_dst_samples - z0));
        resample_tma_kernel<CG, ZERO><<<grid, THREADS, smem, st>>>(mb, mt, a);
    }
    return 0;
}

// Geometry tables, one per (device, call shape), built on first use by tile_geometry_kernel and kept for the life of the
// process (a few KB each).  Building one allocates device memory, which a stream capture forbids: capture after a warm-up
// call (graphed.CapturedStep does), as for every other lazily created resource of this library.
struct GeomKey {
    int dev, mode, N, reflect, G, Hs, Ws, Hd, Wd, pad, has_cs;
    double ox, oy;
    bool operator==(const GeomKey &o) const {
        return dev == o.dev && mode == o.mode && N == o.N && reflect == o.reflect && G == o.G && Hs == o.Hs && Ws == o.Ws &&
               Hd == o.Hd && Wd == o.Wd && pad == o.pad && has_cs == o.has_cs && ox == o.ox && oy == o.oy;
    }
};
static int geometry_table(const ResampleArgs &a, cudaStream_t st, const TileGeom **out) {
    constexpr int MAX_TABLES = 64;
    static GeomKey keys[MAX_TABLES];
    static TileGeom *tables[MAX_TABLES];
    static int count = 0;
    *out = nullptr;
    if (a.mode == MODE_AFFINE || a.G <= 0 || a.G > 64 || getenv("EQB_WARP_NO_TABLE")) return 0;
    const GeomKey k{current_device(), a.mode, a.N, a.reflect, a.G, a.Hs, a.Ws, a.Hd, a.Wd, a.pad, a.has_cs, a.ox, a.oy};
    for (int i = 0; i < count; ++i)
        if (keys[i] == k) {
            *out = tables[i];
            return 0;
        }
    if (count == MAX_TABLES) return 0;           // beyond the cache: CTAs compute their own geometry
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) return 0;   // (no allocation in a capture)
    const int n = a.G * a.tiles_x * a.tiles_y;
    TileGeom *t = nullptr;
    EQB_CUDA(cudaMalloc((void **)&t, (size_t)n * sizeof(TileGeom)));
    ResampleArgs b = a;
    b.geom = nullptr;
    tile_geometry_kernel<<<(n + 127) / 128, 128, 0, st>>>(b, t);
    EQB_CUDA(cudaGetLastError());
    EQB_CUDA(cudaStreamSynchronize(st));         // other streams may use the table from now on
    keys[count] = k;
    tables[count] = t;
    ++count;
    *out = t;
    return 0;
}

int launch_resample_tma(const ResampleArgs &a_in, int n_dst_samples, cudaStream_t st, const char *what, int *handled) {
    *handled = 0;
    ResampleArgs a = a_in;
    a.geom = nullptr;
    const long long planes = (long long)a.B * a.C;
    if (((uintptr_t)a.src & 15) != 0 || (a.Ws & 3) != 0 || a.Ws < BOXW || a.Hs < BOXH || planes <= 0 ||
        planes >= (1LL << 31) || (long long)a.C * a.Hd * a.Wd >= (1LL << 31) || a.tiles_y > 65535)
        return 0;
    const long long blocks = (long long)a.tiles_x * a.tiles_y * n_dst_samples;
    if (blocks == 0) {
        *handled = 1;
        return 0;
    }
    CUtensorMap mb, mt;
    int e = geometry_table(a, st, &a.geom);
    if (e) return e;
    e = make_plane_map(&mb, a.src, a.Ws, a.Hs, planes, BOXW, BOXH, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (e) return e;
    e = make_plane_map(&mt, a.src, a.Ws, a.Hs, planes, TILE, TILE, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
    // can a tap leave the padded extent (zero region)?  radius of the destination rectangle about the centre
    const double ux = fmax(fabs(a.ox), fabs(a.Wd - 1 + a.ox)), uy = fmax(fabs(a.oy), fabs(a.Hd - 1 + a.oy));
    const double rad = sqrt(ux * ux + uy * uy) + 1.0;
    const bool zero = rad > 0.5 * (a.Ws - 1) + a.pad || rad > 0.5 * (a.Hs - 1) + a.pad;
    const int cg = (a.C % 3 == 0 && a.C % 4 != 0) ? 3 : (a.C >= 4 ? 4 : a.C);
    const int nb = n_dst_samples;
    switch (cg * 2 + (zero ? 1 : 0)) {
        case 2: e = launch_cfg<1, false>(mb, mt, a, nb, st); break;
        case 3: e = launch_cfg<1, true>(mb, mt, a, nb, st); break;
        case 4: e = launch_cfg<2, false>(mb, mt, a, nb, st); break;
        case 5: e = launch_cfg<2, true>(mb, mt, a, nb, st); break;
        case 6: e = launch_cfg<3, false>(mb, mt, a, nb, st); break;
        case 7: e = launch_cfg<3, true>(mb, mt, a, nb, st); break;
        case 8: e = launch_cfg<4, false>(mb, mt, a, nb, st); break;
        default: e = launch_cfg<4, true>(mb, mt, a, nb, st); break;
    }
    if (e) return e;
    *handled = 1;
    return finish_launch(what);
}

}  // namespace eqb
