// Group-action image resampling: one kernel family behind eqb_warp_canonicalize (a10),
// eqb_warp_invert (a11) and eqb_orbit_expand (a12).
//
// Every discrete group element acts on pixel coordinates as  src = centre + A (dst - centre)  with a
// 2x2 matrix A built from (cos, sin) of a multiple of 360/N degrees and optional mirror signs.
// What the reference does with pad(replicate) -> hflip blend -> kornia rotate -> crop
// (discrete_group.py:207-215; images/utils.py:57-64; discrete_group.py:401-409) collapses to ONE
// pass over the data: each output pixel is a 4-tap bilinear sample of the source image, taps clamped
// into the image (replicate) while inside the padded extent and zero beyond it.  Quarter turns give
// integral source coordinates, i.e. a pure permutation (the reference's fp32 matrix is +-4e-8 off
// that, SURVEY.md section 7, hard part 2).
//
// HBM-bound design (B200): a CTA owns a 32x32 output tile of one image.  The source footprint of the
// tile (<= 46x46 for any angle) is staged row-wise with coalesced loads into shared memory for up to
// CG channels at once (odd pitch -> the diagonal / column gathers are bank-conflict free or 2-way),
// the bilinear weights are computed once per pixel in fp64 and reused for every channel, and each
// warp writes full 128-byte rows.  Algorithmic traffic: 1 read + 1 write of the image; the 2x
// footprint overlap between neighbouring tiles is absorbed by L2.
#include <math.h>
#include <stdlib.h>

#include "resample.cuh"

namespace eqb {

template <int CG>
__global__ void __launch_bounds__(THREADS) resample_kernel(const __grid_constant__ ResampleArgs a) {
    extern __shared__ float smem[];  // [CG][BB][PITCH]
    const int tiles = a.tiles_x * a.tiles_y;
    const int sample_d = blockIdx.x / tiles;  // destination sample
    const int t = blockIdx.x - sample_d * tiles;
    const int ty0 = (t / a.tiles_x) * TILE, tx0 = (t % a.tiles_x) * TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- group element of this sample -> A ------------------------------------------------
    int sample_s, r = 0, mirror_src = 0, mirror_dst = 0;
    double sign = 1.0;
    if (a.mode == MODE_AFFINE) {
        sample_s = sample_d;
    } else if (a.mode == MODE_ORBIT) {
        const int g = sample_d / a.B;
        sample_s = sample_d - g * a.B;
        r = g % a.N;
        mirror_dst = g >= a.N;  // rotate, THEN hflip (discrete_group.py:404-406)
        sign = -1.0;
    } else {
        sample_s = sample_d;
        const int g = min(max(a.idx[sample_s], 0), a.G - 1);
        r = g % a.N;
        const int refl = g >= a.N;
        if (a.mode == MODE_CANON) {
            mirror_src = refl;  // hflip, THEN rotate(-theta) (discrete_group.py:209-213)
            sign = -1.0;
        } else {
            // rotate(+theta), then x*r + hflip(x)*(1-r): un-reflected samples of a reflection
            // group come back mirrored (images/utils.py:59-64, reference quirk A.4-2)
            mirror_dst = a.reflect && !refl;
            sign = 1.0;
        }
    }
    double a00, a01, a10, a11;
    double cx = 0.5 * (a.Ws - 1), cy = 0.5 * (a.Hs - 1);
    if (a.mode == MODE_AFFINE) {
        const float *m = a.mats + 4 * (size_t)sample_s;
        const double m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
        if (a.mats_forward) {   // dst - c = M (src - c)  ->  src - c = M^-1 (dst - c)
            const double det = m00 * m11 - m01 * m10;
            a00 = m11 / det; a01 = -m01 / det; a10 = -m10 / det; a11 = m00 / det;
        } else {
            a00 = m00; a01 = m01; a10 = m10; a11 = m11;
        }
        cx = a.scx; cy = a.scy;
        if (a.refl && a.refl[sample_s] > 0.5f) {   // the source was mirrored first: xs -> (Ws-1) - xs
            a00 = -a00; a01 = -a01;
            cx = (double)(a.Ws - 1) - cx;
        }
    } else {
        double c, s;
        group_cs(a, r, sign, c, s);
        // src = centre + [[c,-s],[s,c]] (u,v);  mirror_dst: u -> -u;  mirror_src: xs -> (Ws-1) - xs
        a00 = c; a01 = -s; a10 = s; a11 = c;
        if (mirror_dst) { a00 = -a00; a10 = -a10; }
        if (mirror_src) { a00 = -a00; a01 = -a01; }
    }

    // ---- source footprint of the tile --------------------------------------------------------
    const int tw = min(TILE, a.Wd - tx0), th = min(TILE, a.Hd - ty0);
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double u = (double)(tx0 + ((k & 1) ? tw - 1 : 0)) + a.ox;
        const double v = (double)(ty0 + ((k & 2) ? th - 1 : 0)) + a.oy;
        const double xs = cx + a00 * u + a01 * v, ys = cy + a10 * u + a11 * v;
        xmin = fmin(xmin, xs); xmax = fmax(xmax, xs);
        ymin = fmin(ymin, ys); ymax = fmax(ymax, ys);
    }
    // (a non-finite or expanding matrix cannot be a group element; its footprint is cut to the staging box)
    if (!(xmin > -1e9 && xmax < 1e9 && ymin > -1e9 && ymax < 1e9)) { xmin = xmax = ymin = ymax = 0.0; }
    const int x_lo = min(max((int)floor(xmin), 0), a.Ws - 1);
    const int x_hi = min(min(max((int)floor(xmax) + 1, 0), a.Ws - 1), x_lo + BB - 1);
    const int y_lo = min(max((int)floor(ymin), 0), a.Hs - 1);
    const int y_hi = min(min(max((int)floor(ymax) + 1, 0), a.Hs - 1), y_lo + BB - 1);
    const int bw = x_hi - x_lo + 1, bh = y_hi - y_lo + 1;

    // ---- per-pixel taps (shared by all channels) --------------------------------------------
    float w00[PIX], w01[PIX], w10[PIX], w11[PIX];
    int o00[PIX], o01[PIX], o10[PIX], o11[PIX];
    const int xd = tx0 + lane;
#pragma unroll
    for (int p = 0; p < PIX; ++p) {
        const int yd = ty0 + warp + p * (THREADS / 32);
        const double u = (double)xd + a.ox, v = (double)yd + a.oy;
        const double xs = cx + a00 * u + a01 * v, ys = cy + a10 * u + a11 * v;
        const double xf = floor(xs), yf = floor(ys);
        const float fx = (float)(xs - xf), fy = (float)(ys - yf);
        const int x0 = (int)xf, y0 = (int)yf, x1 = x0 + 1, y1 = y0 + 1;
        const int lo_x = -a.pad, hi_x = a.Ws - 1 + a.pad, lo_y = -a.pad, hi_y = a.Hs - 1 + a.pad;
        const float vx0 = (x0 >= lo_x && x0 <= hi_x) ? 1.f : 0.f, vx1 = (x1 >= lo_x && x1 <= hi_x) ? 1.f : 0.f;
        const float vy0 = (y0 >= lo_y && y0 <= hi_y) ? 1.f : 0.f, vy1 = (y1 >= lo_y && y1 <= hi_y) ? 1.f : 0.f;
        // taps are clamped into the staged footprint (which itself is clamped into the image)
        const int cx0 = min(max(x0, x_lo), x_hi) - x_lo, cx1 = min(max(x1, x_lo), x_hi) - x_lo;
        const int cy0 = (min(max(y0, y_lo), y_hi) - y_lo) * PITCH, cy1 = (min(max(y1, y_lo), y_hi) - y_lo) * PITCH;
        w00[p] = (1.f - fy) * (1.f - fx) * vy0 * vx0;
        w01[p] = (1.f - fy) * fx * vy0 * vx1;
        w10[p] = fy * (1.f - fx) * vy1 * vx0;
        w11[p] = fy * fx * vy1 * vx1;
        o00[p] = cy0 + cx0; o01[p] = cy0 + cx1; o10[p] = cy1 + cx0; o11[p] = cy1 + cx1;
    }

    const size_t plane_s = (size_t)a.Hs * a.Ws, plane_d = (size_t)a.Hd * a.Wd;
    const float *src_n = a.src + (size_t)sample_s * a.C * plane_s;
    float *dst_n = a.dst + (size_t)sample_d * a.C * plane_d;

    for (int c0 = 0; c0 < a.C; c0 += CG) {
        const int nc = min(CG, a.C - c0);
        if (c0) __syncthreads();
        // ---- stage the footprint, coalesced along rows ---------------------------------------
        for (int cc = 0; cc < nc; ++cc) {
            int cs = c0 + cc;
            if (a.mode == MODE_INV_REGULAR) {
                // out[:, f*G+g] = in[:, f*G + src_g(g)]   (roll_by_gather, images/utils.py:8-29,66-77)
                const int f = cs / a.G, g = cs - f * a.G, sh = a.roll[r];
                int sg;
                if (g < a.N) sg = (g - sh + a.N) % a.N;
                else sg = a.N + (g - a.N + sh) % a.N;
                cs = f * a.G + sg;
            }
            const float *sp = src_n + (size_t)cs * plane_s + (size_t)y_lo * a.Ws + x_lo;
            float *sm = smem + cc * (BB * PITCH);
            for (int row = warp; row < bh; row += THREADS / 32) {
                const float *rp = sp + (size_t)row * a.Ws;
                if (lane < bw) sm[row * PITCH + lane] = __ldg(rp + lane);
                if (lane + 32 < bw) sm[row * PITCH + lane + 32] = __ldg(rp + lane + 32);
            }
        }
        __syncthreads();
        // ---- gather + store ------------------------------------------------------------------
        if (xd < a.Wd) {
            for (int cc = 0; cc < nc; ++cc) {
                const float *sm = smem + cc * (BB * PITCH);
                float *dp = dst_n + (size_t)(c0 + cc) * plane_d + xd;
#pragma unroll
                for (int p = 0; p < PIX; ++p) {
                    const int yd = ty0 + warp + p * (THREADS / 32);
                    if (yd < a.Hd) {
                        // same tap order as ATen grid_sample: nw, ne, sw, se
                        float v = sm[o00[p]] * w00[p];
                        v = fmaf(sm[o01[p]], w01[p], v);
                        v = fmaf(sm[o10[p]], w10[p], v);
                        v = fmaf(sm[o11[p]], w11[p], v);
                        st_stream(dp + (size_t)yd * a.Wd, v);
                    }
                }
            }
        }
    }
}

// The reference turns the angle back into a channel shift in fp32 and TRUNCATES it:
//   shift = (angle / 360.0 * num_rotations).long()          (images/utils.py:67, :28)
// with angle = torch.linspace(0, 360, N+1)[r] (discrete_group.py:110-112).  For N a power of two
// this is exactly r; for other N the fp32 round trip can land just below r and truncate to r-1
// (reference quirk).  Reproduced here op for op in IEEE float.
static int regular_roll_shift(int r, int N) {
    const volatile float step = (360.0f - 0.0f) / (float)N;  // at::linspace, step_t = float
    const int steps = N + 1, halfway = steps / 2;
    volatile float angle = (r < halfway) ? 0.0f + step * (float)r : 360.0f - step * (float)(steps - r - 1);
    volatile float q = angle / 360.0f;
    volatile float sh = q * (float)N;
    return (int)sh;  // .long() truncates toward zero
}

void finish_args(ResampleArgs &a) {
    a.tiles_x = (a.Wd + TILE - 1) / TILE;
    a.tiles_y = (a.Hd + TILE - 1) / TILE;
    a.has_cs = a.N <= 16;
    if (a.has_cs) {
        static const double quarter[4][2] = {{1.0, 0.0}, {0.0, 1.0}, {-1.0, 0.0}, {0.0, -1.0}};
        for (int r = 0; r < a.N; ++r) {
            if ((4 * r) % a.N == 0) {
                a.cs[2 * r] = quarter[4 * r / a.N][0];
                a.cs[2 * r + 1] = quarter[4 * r / a.N][1];
            } else {
                const double ang = 2.0 * 3.14159265358979323846 * (double)r / (double)a.N;
                a.cs[2 * r] = cos(ang);
                a.cs[2 * r + 1] = sin(ang);
            }
        }
    }
}


// ---- adjoint of the warp (N3, partial): gradient with respect to the warped image -----------------------------
// out = W(g) in is linear in `in`: grad_in = W(g)^T grad_out.  One thread per destination pixel recomputes the four
// taps of the forward pass (same geometry as resample_kernel, fp64 coordinates) and scatters w * grad_out into the
// source gradient with atomicAdd (the summation order of taps that collide is not fixed: results agree to fp32
// rounding, not bit for bit).  Quarter turns / mirrors have one unit tap per pixel: an exact inverse permutation.
__global__ void __launch_bounds__(256) resample_adjoint_kernel(const __grid_constant__ ResampleArgs a) {
    const long long npix = (long long)a.B * a.Hd * a.Wd;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < npix; t += (long long)gridDim.x * blockDim.x) {
        const int xd = (int)(t % a.Wd), yd = (int)((t / a.Wd) % a.Hd), sample = (int)(t / ((long long)a.Wd * a.Hd));
        const int g = min(max(a.idx[sample], 0), a.G - 1), r = g % a.N, refl = g >= a.N;
        int mirror_src = 0, mirror_dst = 0;
        double sign;
        if (a.mode == MODE_CANON) { mirror_src = refl; sign = -1.0; }
        else { mirror_dst = a.reflect && !refl; sign = 1.0; }
        double c, s;
        group_cs(a, r, sign, c, s);
        double a00 = c, a01 = -s, a10 = s, a11 = c;
        if (mirror_dst) { a00 = -a00; a10 = -a10; }
        if (mirror_src) { a00 = -a00; a01 = -a01; }
        const double cx = 0.5 * (a.Ws - 1), cy = 0.5 * (a.Hs - 1);
        const double u = (double)xd + a.ox, v = (double)yd + a.oy;
        const double xs = cx + a00 * u + a01 * v, ys = cy + a10 * u + a11 * v;
        const double xf = floor(xs), yf = floor(ys);
        const float fx = (float)(xs - xf), fy = (float)(ys - yf);
        const int x0 = (int)xf, y0 = (int)yf;
        const int lo_x = -a.pad, hi_x = a.Ws - 1 + a.pad, lo_y = -a.pad, hi_y = a.Hs - 1 + a.pad;
        float w[4];
        int off[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xt = x0 + (k & 1), yt = y0 + (k >> 1);
            const bool in = xt >= lo_x && xt <= hi_x && yt >= lo_y && yt <= hi_y;
            w[k] = in ? ((k >> 1) ? fy : 1.f - fy) * ((k & 1) ? fx : 1.f - fx) : 0.f;
            off[k] = min(max(yt, 0), a.Hs - 1) * a.Ws + min(max(xt, 0), a.Ws - 1);
        }
        const size_t plane_s = (size_t)a.Hs * a.Ws, plane_d = (size_t)a.Hd * a.Wd;
        const float *gy = a.src + (size_t)sample * a.C * plane_d + (size_t)yd * a.Wd + xd;   // src = grad of the warp's output
        float *gx = a.dst + (size_t)sample * a.C * plane_s;                                    // dst = grad of its input
        for (int ch = 0; ch < a.C; ++ch) {
            int cs = ch;
            if (a.mode == MODE_INV_REGULAR) {
                const int f = ch / a.G, gg = ch - f * a.G, sh = a.roll[r];
                const int sg = gg < a.N ? (gg - sh + a.N) % a.N : a.N + (gg - a.N + sh) % a.N;
                cs = f * a.G + sg;
            }
            const float gval = gy[(size_t)ch * plane_d];
            float *gp = gx + (size_t)cs * plane_s;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (w[k] != 0.f) atomicAdd(gp + off[k], w[k] * gval);
        }
    }
}

static int launch_resample(ResampleArgs &a, int n_dst_samples, cudaStream_t st, const char *what) {
    finish_args(a);
    if (!getenv("EQB_NO_TMA")) {
        int handled = 0;
        const int e = launch_resample_tma(a, n_dst_samples, st, what, &handled);
        if (e || handled) return e;
    }
    const long long blocks = (long long)a.tiles_x * a.tiles_y * n_dst_samples;
    if (blocks == 0) return 0;
    EQB_REQUIRE(blocks < (1LL << 31), "%s: grid too large (%lld tiles)", what, blocks);
    const int cg = (a.C % 3 == 0 && a.C % 4 != 0) ? 3 : (a.C >= 4 ? 4 : a.C);
    const size_t smem = (size_t)cg * BB * PITCH * sizeof(float);
    switch (cg) {
        case 1: resample_kernel<1><<<(unsigned)blocks, THREADS, smem, st>>>(a); break;
        case 2: resample_kernel<2><<<(unsigned)blocks, THREADS, smem, st>>>(a); break;
        case 3: resample_kernel<3><<<(unsigned)blocks, THREADS, smem, st>>>(a); break;
        default: resample_kernel<4><<<(unsigned)blocks, THREADS, smem, st>>>(a); break;
    }
    return finish_launch(what);
}

}  // namespace eqb

using namespace eqb;

// host-only helper exported for tests: the truncated channel shift the reference derives from angle r*360/N
extern "C" int eqb_regular_roll_shift(int r, int num_rotations) { return regular_roll_shift(r, num_rotations); }

extern "C" int eqb_warp_canonicalize(const float *x, float *y, const int32_t *idx, int B, int C, int H, int W,
                                     int num_rotations, int reflect, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0, "eqb_warp_canonicalize: bad shape (%d,%d,%d,%d)", B, C, H, W);
    EQB_REQUIRE(num_rotations > 0, "eqb_warp_canonicalize: num_rotations must be positive");
    EQB_REQUIRE(B == 0 || (x && y && idx), "eqb_warp_canonicalize: null pointer");
    ResampleArgs a{};
    a.src = x; a.dst = y; a.idx = idx; a.B = B; a.C = C;
    a.Hs = a.Hd = H; a.Ws = a.Wd = W;
    a.N = num_rotations; a.reflect = reflect != 0; a.G = num_rotations * (reflect ? 2 : 1);
    a.mode = MODE_CANON;
    // Pad(ceil(W/2), edge) on all four sides for C != 1, Identity for grayscale (discrete_group.py:60-66)
    a.pad = (C == 1) ? 0 : (W + 1) / 2;
    a.ox = -0.5 * (W - 1); a.oy = -0.5 * (H - 1);
    return launch_resample(a, B, (cudaStream_t)stream, "eqb_warp_canonicalize");
}

extern "C" int eqb_warp_invert(const float *f, float *out, const int32_t *idx, int B, int C, int H, int W,
                               int num_rotations, int reflect, int rep, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0, "eqb_warp_invert: bad shape (%d,%d,%d,%d)", B, C, H, W);
    EQB_REQUIRE(num_rotations > 0, "eqb_warp_invert: num_rotations must be positive");
    EQB_REQUIRE(rep == EQB_REP_SCALAR || rep == EQB_REP_REGULAR, "eqb_warp_invert: rep must be scalar or regular");
    const int G = num_rotations * (reflect ? 2 : 1);
    EQB_REQUIRE(rep != EQB_REP_REGULAR || C % G == 0,
                "eqb_warp_invert: regular representation needs C %% |G| == 0 (C=%d, |G|=%d)", C, G);
    EQB_REQUIRE(B == 0 || (f && out && idx), "eqb_warp_invert: null pointer");
    ResampleArgs a{};
    a.src = f; a.dst = out; a.idx = idx; a.B = B; a.C = C;
    a.Hs = a.Hd = H; a.Ws = a.Wd = W;
    a.N = num_rotations; a.reflect = reflect != 0; a.G = G;
    a.mode = rep == EQB_REP_REGULAR ? MODE_INV_REGULAR : MODE_INV_SCALAR;
    if (rep == EQB_REP_REGULAR) {
        EQB_UNSUPPORTED(num_rotations > 64, "eqb_warp_invert: regular representation supports num_rotations <= 64");
        for (int r = 0; r < num_rotations; ++r) a.roll[r] = (signed char)regular_roll_shift(r, num_rotations);
    }
    a.pad = 0;  // bare kornia rotate: zero fill (images/utils.py:57)
    a.ox = -0.5 * (W - 1); a.oy = -0.5 * (H - 1);
    return launch_resample(a, B, (cudaStream_t)stream, "eqb_warp_invert");
}

extern "C" int eqb_orbit_expand(const float *x, float *out, int B, int C, int h, int w, int pad, int out_size,
                                int num_rotations, int reflect, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && C > 0 && h > 0 && w > 0, "eqb_orbit_expand: bad shape (%d,%d,%d,%d)", B, C, h, w);
    EQB_REQUIRE(num_rotations > 0 && pad >= 0, "eqb_orbit_expand: bad group / pad");
    ResampleArgs a{};
    a.src = x; a.dst = out; a.idx = nullptr; a.B = B; a.C = C;
    a.Hs = h; a.Ws = w;
    a.N = num_rotations; a.reflect = reflect != 0; a.G = num_rotations * (reflect ? 2 : 1);
    a.mode = MODE_ORBIT;
    if (C == 1) {  // grayscale: pad_group_augment / crop_group_augment are Identity (discrete_group.py:365-376)
        a.pad = 0; a.Hd = h; a.Wd = w;
        a.ox = -0.5 * (w - 1); a.oy = -0.5 * (h - 1);
    } else {
        EQB_REQUIRE(out_size > 0 && out_size <= h + 2 * pad && out_size <= w + 2 * pad,
                    "eqb_orbit_expand: crop %d does not fit the padded image", out_size);
        a.pad = pad; a.Hd = a.Wd = out_size;
        // CenterCrop offsets in the padded frame (python round == rint, half to even)
        const double left = rint((w + 2 * pad - out_size) / 2.0), top = rint((h + 2 * pad - out_size) / 2.0);
        a.ox = left - 0.5 * (w + 2 * pad - 1);
        a.oy = top - 0.5 * (h + 2 * pad - 1);
    }
    EQB_REQUIRE(B == 0 || (x && out), "eqb_orbit_expand: null pointer");
    return launch_resample(a, B * a.G, (cudaStream_t)stream, "eqb_orbit_expand");
}

// ---- N2: continuous rotations / roto-reflections of images ---------------------------------------
extern "C" int eqb_warp_affine(const float *x, float *y, const float *mats, const float *refl, int mats_forward, int B,
                               int C, int H, int W, int pad, double cx, double cy, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && pad >= 0, "eqb_warp_affine: bad shape (%d,%d,%d,%d) / pad %d", B, C, H, W, pad);
    EQB_REQUIRE(B == 0 || (x && y && mats), "eqb_warp_affine: null pointer");
    ResampleArgs a{};
    a.src = x; a.dst = y; a.idx = nullptr; a.B = B; a.C = C;
    a.Hs = a.Hd = H; a.Ws = a.Wd = W;
    a.N = 1; a.G = 1; a.reflect = 0;
    a.mode = MODE_AFFINE;
    a.mats = mats; a.refl = refl; a.mats_forward = mats_forward != 0;
    a.pad = pad;
    a.scx = cx; a.scy = cy;
    a.ox = -cx; a.oy = -cy;     // destination pixel relative to the same centre
    return launch_resample(a, B, (cudaStream_t)stream, "eqb_warp_affine");
}

// ---- N3 (partial): gradient of the discrete warps with respect to their image argument --------------------
// mode 0: adjoint of eqb_warp_canonicalize, 1: of eqb_warp_invert (scalar), 2: of eqb_warp_invert (regular).
extern "C" int eqb_warp_adjoint(const float *grad_out, float *grad_in, const int32_t *idx, int B, int C, int H, int W,
                                int num_rotations, int reflect, int mode, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && num_rotations > 0, "eqb_warp_adjoint: bad shape");
    EQB_REQUIRE(mode >= 0 && mode <= 2, "eqb_warp_adjoint: mode must be 0 (canonicalize), 1 (invert scalar) or 2 (invert regular)");
    const int G = num_rotations * (reflect ? 2 : 1);
    EQB_REQUIRE(mode != 2 || C % G == 0, "eqb_warp_adjoint: regular representation needs C %% |G| == 0");
    EQB_REQUIRE(B == 0 || (grad_out && grad_in && idx), "eqb_warp_adjoint: null pointer");
    if (B == 0) return 0;
    ResampleArgs a{};
    a.src = grad_out; a.dst = grad_in; a.idx = idx; a.B = B; a.C = C;
    a.Hs = a.Hd = H; a.Ws = a.Wd = W;
    a.N = num_rotations; a.reflect = reflect != 0; a.G = G;
    a.mode = mode == 0 ? MODE_CANON : mode == 1 ? MODE_INV_SCALAR : MODE_INV_REGULAR;
    if (mode == 2) {
        EQB_UNSUPPORTED(num_rotations > 64, "eqb_warp_adjoint: regular representation supports num_rotations <= 64");
        for (int r = 0; r < num_rotations; ++r) a.roll[r] = (signed char)regular_roll_shift(r, num_rotations);
    }
    a.pad = (mode == 0 && C != 1) ? (W + 1) / 2 : 0;
    a.ox = -0.5 * (W - 1); a.oy = -0.5 * (H - 1);
    finish_args(a);
    cudaStream_t st = (cudaStream_t)stream;
    EQB_CUDA(cudaMemsetAsync(grad_in, 0, (size_t)B * C * H * W * sizeof(float), st));
    const long long npix = (long long)B * H * W;
    long long blocks = (npix + 255) / 256;
    if (blocks > (long long)num_sms() * 32) blocks = (long long)num_sms() * 32;
    resample_adjoint_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    return finish_launch("eqb_warp_adjoint");
}

// ---- N4: evaluation-time group orbit (examples/images/classification/inference_utils.py:97-122) -----------------
// out[g, b, c] = CenterCrop(H, W)( torchvision.rotate_nearest( [hflip]( Pad(pad, edge)(x[b, c]) ), degree_g ) ).
// The source pixel of an output pixel depends on (g, y, x) only, so a thread resolves it once and then streams every
// (b, c) plane of its chunk through it.  The float32 operation order is torchvision's (affine base grid times the
// rescaled theta, then grid_sample's unnormalise + nearbyint) so that rounding ties fall the same way.
struct OrbitNearestArgs {
    const float *src;
    float *dst;
    int planes, H, W, pad, G, N;
    int top, left;
    float r[64][6];   // per rotation: theta^T / (0.5 wp, 0.5 hp) as [r00 r10 r20 r01 r11 r21]
};

// source offset (sy * W + sx) of output pixel (y, x) under element g, or -1 when torchvision's zero fill applies
__device__ __forceinline__ int orbit_nearest_source(const OrbitNearestArgs &a, int g, int y, int x) {
    const int hp = a.H + 2 * a.pad, wp = a.W + 2 * a.pad;
    const float *r = a.r[g % a.N];
    const float bx = __fadd_rn((float)(x + a.left), (float)(-wp * 0.5 + 0.5));
    const float by = __fadd_rn((float)(y + a.top), (float)(-hp * 0.5 + 0.5));
    const float gx = __fadd_rn(__fadd_rn(__fmul_rn(bx, r[0]), __fmul_rn(by, r[1])), r[2]);
    const float gy = __fadd_rn(__fadd_rn(__fmul_rn(bx, r[3]), __fmul_rn(by, r[4])), r[5]);
    const float ix = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)wp), -1.f), 0.5f);
    const float iy = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)hp), -1.f), 0.5f);
    const float fx = rintf(ix), fy = rintf(iy);
    if (!(fx >= 0.f && fx < (float)wp && fy >= 0.f && fy < (float)hp)) return -1;
    int jx = (int)fx;
    const int jy = (int)fy;
    if (g >= a.N) jx = wp - 1 - jx;
    return min(max(jy - a.pad, 0), a.H - 1) * a.W + min(max(jx - a.pad, 0), a.W - 1);
}

// grid (32-pixel columns, 8-row bands, |G| x plane chunks).  A warp owns an 8 x 4 pixel patch, not a 32 x 1 row: under a
// quarter turn a row of 32 outputs reads 32 different source rows (32 L1 wavefronts per load instruction), while the
// patch reads 4..9 rows and still stores four complete 32-byte sectors.  Each thread resolves its source pixel once
// and walks the planes of its chunk with UNROLL loads in flight.
__global__ void __launch_bounds__(256) orbit_nearest_kernel(const __grid_constant__ OrbitNearestArgs a, int chunks) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int y = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (x >= a.W || y >= a.H) return;
    const int g = blockIdx.z / chunks, chunk = blockIdx.z % chunks;
    const int so = orbit_nearest_source(a, g, y, x);
    const size_t plane = (size_t)a.H * a.W;
    const int per = (a.planes + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(p0 + per, a.planes);
    float *d = a.dst + ((size_t)g * a.planes + p0) * plane + (size_t)y * a.W + x;
    if (so < 0) {                                    // torchvision's zero fill outside the padded image
        for (int p = p0; p < p1; ++p, d += plane) __stcs(d, 0.f);
        return;
    }
    const float *s = a.src + (size_t)p0 * plane + so;
    constexpr int UNROLL = 8;
    int p = p0;
    for (; p + UNROLL <= p1; p += UNROLL) {
        float val[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) val[u] = __ldg(s + (size_t)u * plane);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) __stcs(d + (size_t)u * plane, val[u]);
        s += (size_t)UNROLL * plane; d += (size_t)UNROLL * plane;
    }
    for (; p < p1; ++p, s += plane, d += plane) __stcs(d, __ldg(s));
}

// torch.linspace(0, 360, n + 1)[i] in float32: first half start + i*step, second half end - (steps-1-i)*step
static float linspace_degree(int i, int n) {
    const int steps = n + 1;
    const float step = 360.0f / (float)(steps - 1);
    return i < steps / 2 ? 0.0f + step * (float)i : 360.0f - step * (float)(steps - 1 - i);
}

extern "C" int eqb_orbit_rotate_nearest(const float *x, float *out, int B, int C, int H, int W, int num_rotations,
                                        int reflect, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0, "eqb_orbit_rotate_nearest: bad shape (%d,%d,%d,%d)", B, C, H, W);
    EQB_REQUIRE(num_rotations > 0, "eqb_orbit_rotate_nearest: num_rotations must be positive");
    EQB_UNSUPPORTED(num_rotations > 64, "eqb_orbit_rotate_nearest: num_rotations <= 64 supported");
    EQB_REQUIRE(B == 0 || (x && out), "eqb_orbit_rotate_nearest: null pointer");
    if (B == 0) return 0;
    OrbitNearestArgs a{};
    a.src = x; a.dst = out; a.planes = B * C; a.H = H; a.W = W;
    a.pad = (int)ceil(H * 0.4);                     // transforms.Pad(math.ceil(in_shape[-2] * 0.4), "edge")
    a.N = num_rotations; a.G = num_rotations * (reflect ? 2 : 1);
    const int hp = H + 2 * a.pad, wp = W + 2 * a.pad;
    a.top = (int)rint((hp - H) / 2.0); a.left = (int)rint((wp - W) / 2.0);
    const float sw = 0.5f * (float)wp, sh = 0.5f * (float)hp;
    for (int i = 0; i < num_rotations; ++i) {
        // torchvision F.rotate: _get_inverse_affine_matrix([0,0], -angle, [0,0], 1, [0,0]) in doubles -> float32 theta
        const double rot = -(double)linspace_degree(i, num_rotations) * (M_PI / 180.0);   // math.radians
        const double ca = cos(rot), sa = sin(rot);
        const float m0 = (float)ca, m1 = (float)sa, m3 = (float)(-sa), m4 = (float)ca;
        a.r[i][0] = m0 / sw; a.r[i][1] = m1 / sw; a.r[i][2] = 0.f / sw;
        a.r[i][3] = m3 / sh; a.r[i][4] = m4 / sh; a.r[i][5] = 0.f / sh;
    }
    // plane chunks: enough CTAs for many waves of 8 resident CTAs per SM, at least 8 planes each so the index
    // arithmetic stays amortised
    const int tiles_x = (W + 31) / 32, tiles_y = (H + 7) / 8;
    EQB_REQUIRE(tiles_y <= 65535, "eqb_orbit_rotate_nearest: image too tall for one launch");
    const long long base = (long long)tiles_x * tiles_y * a.G;
    static const int ctas_per_sm = getenv("EQB_ORBIT_CTAS_PER_SM") ? atoi(getenv("EQB_ORBIT_CTAS_PER_SM")) : 64;   // tuning knob
    int chunks = (int)(((long long)ctas_per_sm * num_sms() + base - 1) / base);
    chunks = std::max(1, std::min(std::min(chunks, (a.planes + 7) / 8), 65535 / a.G));
    dim3 grid(tiles_x, tiles_y, a.G * chunks);
    orbit_nearest_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, chunks);
    return finish_launch("eqb_orbit_rotate_nearest");
}

// ---- N3: gradient of the discrete warps with respect to the GROUP ELEMENT ----------------------------------------
// In training the reference feeds the warps a straight-through element: rotation (degrees) = sum_g onehot_g * angle_g,
// reflection = sum_g onehot_g * [g >= N] (discrete_group.py:110-133), so the task loss reaches the canonicalization
// network through d warp / d rotation (kornia rotate -> grid_sample's grid gradient) and through the flip blend
// (1 - r) x + r hflip(x) (discrete_group.py:209-210; images/utils.py:59-64).  One thread per destination pixel:
//   d out / d theta = (d out / d xs) (d xs / d theta) + (d out / d ys) (d ys / d theta), bilinear cell derivatives
//                     with the forward's validity masks (zero beyond the padded extent, replicate inside it);
//   d out / d r     = value under the mirrored blend minus value under the plain one.
// Per-sample sums over (c, y, x) are reduced per block in fp64 and added atomically.
namespace eqb {

struct ElementGradArgs {
    ResampleArgs r;          // src = the warp's input image, dst unused
    const float *grad_out;
    float *grad_rotation;    // (B), d loss / d rotation in DEGREES
    float *grad_reflection;  // (B) or null
};

struct BilinearCell {
    float wx[2], wy[2], valid[4];
    int xt[4], yt[4];
    // taps one step before the cell, used for the symmetric derivative when the sample sits exactly on a lattice line:
    // (x0-1, y0), (x0-1, y1) when on_x; (x0, y0-1), (x1, y0-1) when on_y
    bool on_x, on_y;
    float valid_m[4];
    int xm[4], ym[4];
};

__device__ __forceinline__ void bilinear_cell(const ResampleArgs &a, double xs, double ys, BilinearCell &q) {
    const double xf = floor(xs), yf = floor(ys);
    const float fx = (float)(xs - xf), fy = (float)(ys - yf);
    const int x0 = (int)xf, y0 = (int)yf;
    const int lo_x = -a.pad, hi_x = a.Ws - 1 + a.pad, lo_y = -a.pad, hi_y = a.Hs - 1 + a.pad;
    q.wx[0] = 1.f - fx; q.wx[1] = fx; q.wy[0] = 1.f - fy; q.wy[1] = fy;
    q.on_x = fx == 0.f; q.on_y = fy == 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xk = x0 + (k & 1), yk = y0 + (k >> 1);
        q.valid[k] = (xk >= lo_x && xk <= hi_x && yk >= lo_y && yk <= hi_y) ? 1.f : 0.f;
        q.xt[k] = min(max(xk, 0), a.Ws - 1);
        q.yt[k] = min(max(yk, 0), a.Hs - 1);
        const int xj = k < 2 ? x0 - 1 : x0 + (k & 1), yj = k < 2 ? y0 + (k & 1) : y0 - 1;
        q.valid_m[k] = (xj >= lo_x && xj <= hi_x && yj >= lo_y && yj <= hi_y) ? 1.f : 0.f;
        q.xm[k] = min(max(xj, 0), a.Ws - 1);
        q.ym[k] = min(max(yj, 0), a.Hs - 1);
    }
}

__global__ void __launch_bounds__(256) warp_element_grad_kernel(const __grid_constant__ ElementGradArgs e) {
    const ResampleArgs &a = e.r;
    const int sample = blockIdx.y;
    const int g = min(max(a.idx[sample], 0), a.G - 1), r = g % a.N, refl = g >= a.N;
    const bool canon = a.mode == MODE_CANON;
    double c, s;
    group_cs(a, r, canon ? -1.0 : 1.0, c, s);       // rotate(x, -theta) for canonicalize, rotate(f, +theta) for invert
    const double dsign = canon ? -1.0 : 1.0;        // d(angle of the matrix) / d(theta)
    const double cx = 0.5 * (a.Ws - 1), cy = 0.5 * (a.Hs - 1);
    const size_t plane = (size_t)a.Hs * a.Ws;
    const float *img = a.src + (size_t)sample * a.C * plane;
    const float *go = e.grad_out + (size_t)sample * a.C * plane;
    double acc_rot = 0.0, acc_ref = 0.0;
    const int npix = a.Hd * a.Wd;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < npix; t += gridDim.x * blockDim.x) {
        const int xd = t % a.Wd, yd = t / a.Wd;
        // canonicalize: the blended image (mirrored when refl) is rotated; sample position from (xd, yd).
        // invert: out(xd) = r R(xd) + (1 - r) R(W-1-xd) with R = rotate(f): the active sample sits at xd when refl = 1
        // (or there is no reflection at all), at the mirrored column otherwise.
        const int xa = (!canon && a.reflect && !refl) ? a.Wd - 1 - xd : xd;
        const double u = (double)xa + a.ox, v = (double)yd + a.oy;
        const double xs = cx + c * u - s * v, ys = cy + s * u + c * v;
        const double dxs = dsign * (-s * u - c * v), dys = dsign * (c * u - s * v);     // per radian of theta
        BilinearCell q;
        bilinear_cell(a, xs, ys, q);
        // the other sample of the blend: canonicalize -> same taps in the mirrored image; invert -> R at the mirrored column
        BilinearCell q2;
        if (a.reflect && !canon) {
            const double u2 = (double)(a.Wd - 1 - xa) + a.ox;
            bilinear_cell(a, cx + c * u2 - s * v, cy + s * u2 + c * v, q2);
        }
        const bool flip_src = canon && refl;
        float sum_rot = 0.f, sum_ref = 0.f;
        for (int ch = 0; ch < a.C; ++ch) {
            int cs = ch;
            if (a.mode == MODE_INV_REGULAR) {
                const int f = ch / a.G, gg = ch - f * a.G, sh = a.roll[r];
                const int sg = gg < a.N ? (gg - sh + a.N) % a.N : a.N + (gg - a.N + sh) % a.N;
                cs = f * a.G + sg;
            }
            const float *ip = img + (size_t)cs * plane;
            const float gval = go[(size_t)ch * plane + (size_t)yd * a.Wd + xd];
            float val[4], alt = 0.f, cur = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xk = flip_src ? a.Ws - 1 - q.xt[k] : q.xt[k];
                val[k] = q.valid[k] * __ldg(ip + q.yt[k] * a.Ws + xk);
                cur += val[k] * q.wx[k & 1] * q.wy[k >> 1];
                if (canon && a.reflect) alt += q.valid[k] * __ldg(ip + q.yt[k] * a.Ws + (a.Ws - 1 - xk)) * q.wx[k & 1] * q.wy[k >> 1];
            }
            // one-sided cell derivatives; exactly on a lattice line (quarter turns: every pixel) the bilinear interpolant
            // has a kink and the SYMMETRIC derivative is returned -- the mean of the two one-sided values between which
            // the reference's fp32 coordinates (+-4e-8 off the lattice) choose at random
            float gx = q.wy[0] * (val[1] - val[0]) + q.wy[1] * (val[3] - val[2]);
            float gy = q.wx[0] * (val[2] - val[0]) + q.wx[1] * (val[3] - val[1]);
            if (q.on_x) {
                const float m0 = q.valid_m[0] * __ldg(ip + q.ym[0] * a.Ws + (flip_src ? a.Ws - 1 - q.xm[0] : q.xm[0]));
                const float m1 = q.valid_m[1] * __ldg(ip + q.ym[1] * a.Ws + (flip_src ? a.Ws - 1 - q.xm[1] : q.xm[1]));
                gx = 0.5f * (gx + q.wy[0] * (val[0] - m0) + q.wy[1] * (val[2] - m1));
            }
            if (q.on_y) {
                const float m0 = q.valid_m[2] * __ldg(ip + q.ym[2] * a.Ws + (flip_src ? a.Ws - 1 - q.xm[2] : q.xm[2]));
                const float m1 = q.valid_m[3] * __ldg(ip + q.ym[3] * a.Ws + (flip_src ? a.Ws - 1 - q.xm[3] : q.xm[3]));
                gy = 0.5f * (gy + q.wx[0] * (val[0] - m0) + q.wx[1] * (val[1] - m1));
            }
            sum_rot += gval * (gx * (float)dxs + gy * (float)dys);
            if (a.reflect) {
                if (!canon) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) alt += q2.valid[k] * __ldg(ip + q2.yt[k] * a.Ws + q2.xt[k]) * q2.wx[k & 1] * q2.wy[k >> 1];
                }
                // d/dr of r*A + (1-r)*B = A - B, where `cur` is the active member (A when refl = 1, B when refl = 0)
                sum_ref += gval * (refl ? cur - alt : alt - cur);
            }
        }
        acc_rot += (double)sum_rot;
        acc_ref += (double)sum_ref;
    }
    // block reduction (fp64), one atomic per block and output
    __shared__ double red[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        acc_rot += __shfl_down_sync(0xffffffffu, acc_rot, o);
        acc_ref += __shfl_down_sync(0xffffffffu, acc_ref, o);
    }
    if (lane == 0) { red[0][warp] = acc_rot; red[1][warp] = acc_ref; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tr = 0.0, tf = 0.0;
        for (int w = 0; w < 8; ++w) { tr += red[0][w]; tf += red[1][w]; }
        atomicAdd(e.grad_rotation + sample, (float)(tr * (3.14159265358979323846 / 180.0)));
        if (e.grad_reflection) atomicAdd(e.grad_reflection + sample, (float)tf);
    }
}

}  // namespace eqb

extern "C" int eqb_warp_element_grad(const float *in, const float *grad_out, const int32_t *idx, int B, int C, int H, int W,
                                     int num_rotations, int reflect, int mode, float *grad_rotation,
                                     float *grad_reflection, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && num_rotations > 0, "eqb_warp_element_grad: bad shape");
    EQB_REQUIRE(mode >= 0 && mode <= 2, "eqb_warp_element_grad: mode must be 0 (canonicalize), 1 (invert scalar) or 2 (invert regular)");
    const int G = num_rotations * (reflect ? 2 : 1);
    EQB_REQUIRE(mode != 2 || C % G == 0, "eqb_warp_element_grad: regular representation needs C %% |G| == 0");
    EQB_REQUIRE(B == 0 || (in && grad_out && idx && grad_rotation), "eqb_warp_element_grad: null pointer");
    EQB_REQUIRE(B <= 65535, "eqb_warp_element_grad: at most 65535 samples per call");
    if (B == 0) return 0;
    ElementGradArgs e{};
    ResampleArgs &a = e.r;
    a.src = in; a.dst = nullptr; a.idx = idx; a.B = B; a.C = C;
    a.Hs = a.Hd = H; a.Ws = a.Wd = W;
    a.N = num_rotations; a.reflect = reflect != 0; a.G = G;
    a.mode = mode == 0 ? MODE_CANON : mode == 1 ? MODE_INV_SCALAR : MODE_INV_REGULAR;
    if (mode == 2) {
        EQB_UNSUPPORTED(num_rotations > 64, "eqb_warp_element_grad: regular representation supports num_rotations <= 64");
        for (int r = 0; r < num_rotations; ++r) a.roll[r] = (signed char)regular_roll_shift(r, num_rotations);
    }
    a.pad = (mode == 0 && C != 1) ? (W + 1) / 2 : 0;
    a.ox = -0.5 * (W - 1); a.oy = -0.5 * (H - 1);
    finish_args(a);
    e.grad_out = grad_out; e.grad_rotation = grad_rotation; e.grad_reflection = reflect ? grad_reflection : nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    EQB_CUDA(cudaMemsetAsync(grad_rotation, 0, (size_t)B * sizeof(float), st));
    if (e.grad_reflection) EQB_CUDA(cudaMemsetAsync(grad_reflection, 0, (size_t)B * sizeof(float), st));
    int bx = (H * W + 255) / 256;
    const int want = std::max(1, (8 * num_sms() + B - 1) / B);
    bx = std::max(1, std::min(bx, want));
    warp_element_grad_kernel<<<dim3(bx, B), 256, 0, st>>>(e);
    return finish_launch("eqb_warp_element_grad");
}

// ---- N3: gradient of the continuous warp (eqb_warp_affine, canonicalize direction) with respect to the affine map ----
// The reference differentiates K.geometry.warp_affine with respect to its 2x3 matrix (continuous_group.py:183-208); the
// matrix enters through its inverse, the translation column through alpha / beta only, and the reflection through the
// flip blend.  All of that is tiny (B,3,3) algebra that stays in torch autograd; what needs a kernel is the reduction
//     d loss / d theta[b] (2x3),  theta = destination -> source map in un-padded pixel coordinates:
//         xs = t0 xd + t1 yd + t2,  ys = t3 xd + t4 yd + t5,
//     d loss / d reflection[b]   (value under the mirrored blend minus value under the plain one),
// with the forward's sampling rules (bilinear, replicate inside the padded extent, zero beyond, symmetric derivative on
// lattice lines).
namespace eqb {

struct AffineGradArgs {
    const float *in, *grad_out, *theta, *refl;
    float *gtheta, *grefl;
    int C, H, W, pad;
};

__global__ void __launch_bounds__(256) warp_affine_grad_kernel(const __grid_constant__ AffineGradArgs e) {
    const int sample = blockIdx.y;
    ResampleArgs a{};                 // only the fields bilinear_cell reads
    a.Hs = e.H; a.Ws = e.W; a.pad = e.pad;
    const float *th = e.theta + (size_t)sample * 6;
    const double t0 = th[0], t1 = th[1], t2 = th[2], t3 = th[3], t4 = th[4], t5 = th[5];
    const bool refl = e.refl && e.refl[sample] > 0.5f;
    const size_t plane = (size_t)e.H * e.W;
    const float *img = e.in + (size_t)sample * e.C * plane;
    const float *go = e.grad_out + (size_t)sample * e.C * plane;
    double acc[7] = {};
    const int npix = e.H * e.W;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < npix; t += gridDim.x * blockDim.x) {
        const int xd = t % e.W, yd = t / e.W;
        BilinearCell q;
        bilinear_cell(a, t0 * xd + t1 * yd + t2, t3 * xd + t4 * yd + t5, q);
        float sgx = 0.f, sgy = 0.f, sref = 0.f;
        for (int ch = 0; ch < e.C; ++ch) {
            const float *ip = img + (size_t)ch * plane;
            const float gval = go[(size_t)ch * plane + t];
            float val[4], cur = 0.f, alt = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xk = refl ? e.W - 1 - q.xt[k] : q.xt[k];
                val[k] = q.valid[k] * __ldg(ip + q.yt[k] * e.W + xk);
                cur += val[k] * q.wx[k & 1] * q.wy[k >> 1];
                if (e.grefl) alt += q.valid[k] * __ldg(ip + q.yt[k] * e.W + (e.W - 1 - xk)) * q.wx[k & 1] * q.wy[k >> 1];
            }
            float gx = q.wy[0] * (val[1] - val[0]) + q.wy[1] * (val[3] - val[2]);
            float gy = q.wx[0] * (val[2] - val[0]) + q.wx[1] * (val[3] - val[1]);
            if (q.on_x) {
                const float m0 = q.valid_m[0] * __ldg(ip + q.ym[0] * e.W + (refl ? e.W - 1 - q.xm[0] : q.xm[0]));
                const float m1 = q.valid_m[1] * __ldg(ip + q.ym[1] * e.W + (refl ? e.W - 1 - q.xm[1] : q.xm[1]));
                gx = 0.5f * (gx + q.wy[0] * (val[0] - m0) + q.wy[1] * (val[2] - m1));
            }
            if (q.on_y) {
                const float m0 = q.valid_m[2] * __ldg(ip + q.ym[2] * e.W + (refl ? e.W - 1 - q.xm[2] : q.xm[2]));
                const float m1 = q.valid_m[3] * __ldg(ip + q.ym[3] * e.W + (refl ? e.W - 1 - q.xm[3] : q.xm[3]));
                gy = 0.5f * (gy + q.wx[0] * (val[0] - m0) + q.wx[1] * (val[1] - m1));
            }
            sgx += gval * gx;
            sgy += gval * gy;
            sref += gval * (refl ? cur - alt : alt - cur);
        }
        acc[0] += (double)sgx * xd; acc[1] += (double)sgx * yd; acc[2] += (double)sgx;
        acc[3] += (double)sgy * xd; acc[4] += (double)sgy * yd; acc[5] += (double)sgy;
        acc[6] += (double)sref;
    }
    __shared__ double red[8][7];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double tsum = 0.0;
        for (int w = 0; w < 8; ++w) tsum += red[w][threadIdx.x];
        if (threadIdx.x < 6) atomicAdd(e.gtheta + (size_t)sample * 6 + threadIdx.x, (float)tsum);
        else if (e.grefl) atomicAdd(e.grefl + sample, (float)tsum);
    }
}

}  // namespace eqb

extern "C" int eqb_warp_affine_grad(const float *in, const float *grad_out, const float *theta, const float *refl, int B,
                                    int C, int H, int W, int pad, float *grad_theta, float *grad_refl, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && pad >= 0, "eqb_warp_affine_grad: bad shape");
    EQB_REQUIRE(B == 0 || (in && grad_out && theta && grad_theta), "eqb_warp_affine_grad: null pointer");
    EQB_REQUIRE(!grad_refl || refl, "eqb_warp_affine_grad: grad_refl needs refl");
    EQB_REQUIRE(B <= 65535, "eqb_warp_affine_grad: at most 65535 samples per call");
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    EQB_CUDA(cudaMemsetAsync(grad_theta, 0, (size_t)B * 6 * sizeof(float), st));
    if (grad_refl) EQB_CUDA(cudaMemsetAsync(grad_refl, 0, (size_t)B * sizeof(float), st));
    AffineGradArgs e{in, grad_out, theta, refl, grad_theta, grad_refl, C, H, W, pad};
    int bx = (H * W + 255) / 256;
    bx = std::max(1, std::min(bx, std::max(1, (8 * num_sms() + B - 1) / B)));
    warp_affine_grad_kernel<<<dim3(bx, B), 256, 0, st>>>(e);
    return finish_launch("eqb_warp_affine_grad");
}
