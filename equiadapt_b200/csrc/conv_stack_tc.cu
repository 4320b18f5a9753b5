// a7 on the tensor cores: the INNER layers of the e2cnn-style conv stack (conv_stack.cu) as a tcgen05 implicit GEMM.
//
//   out[b, n, oy, ox] = relu( scale_n * ( sum_{c,ky,kx} W[n,c,ky,kx] * in[b, c, oy+ky, ox+kx] + bias_n ) + shift_n )
//   M = 128 output pixels (an 8 x 16 patch of one image), N = Cout*|G| (padded to 32..256), K = Cin * k * k.
//
// Activations travel between layers as two fp16 tensors in NHWC (hi / lo halves of the power-of-two scaled fp32 value,
// the same 22-bit split as gconv_stack_tc.cu).  A tile is a 16 x 8 patch of output pixels; for every 32-channel group
// ONE TMA box fetches the patch WITH ITS HALO,
//   cp.async.bulk.tensor.4d over (C, W, H, B), box 32 x (8 + k - 1) x (16 + k - 1) x 1, SWIZZLE_64B,
// and all k*k (ky, kx) K atoms read their A operand from it: rows (py + ky) * HW + kx + 0..7 of the box are exactly the
// 8-row core groups of a K-major SWIZZLE_64B operand with group stride HW * 64 bytes, so the (ky, kx) shift is nothing
// but the descriptor's start address (the UMMA swizzle is a function of the absolute shared-memory address: any row
// shift and any group pitch work with base_offset 0 - verified bit-exact with tools/umma_shift_probe.cu).  Each input
// element crosses L2 -> SM once instead of k*k times; image borders are TMA zero fill.  Weights are pre-packed per
// atom as hi / lo UMMA images and streamed by cp.async.bulk through their own ring.  Each product is a_hi*w_hi + a_lo*w_hi +
// a_hi*w_lo (three kind::f16 MMAs, fp32 accumulation in TMEM).  Two accumulators (2 x Npad columns) let the epilogue
// of tile t overlap the main loop of tile t+1.
//   warp 0  producer (one lane): 2 TMA boxes + 1 bulk weight copy per stage, mbarrier tx-count
//   warp 1  MMA issuer (elected lane)      warp 2  TMEM allocator      warps 4-7  epilogue
// The epilogue undoes the operand scales, applies bias / folded batch norm / ReLU and writes either the next layer's
// NHWC fp16 hi / lo pair or (for the layer that feeds the folded last layer) fp32 NCHW.
// Operand scales (exact powers of two) are chained on the device from max|x| and the weights' absolute row sums
// (ctc_layer_stats_kernel), no host synchronisation.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv_stack_tc.cuh"

namespace eqb {
namespace ctc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait (~2 s): a pipeline bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > 4000000000LL) __trap();
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
constexpr uint32_t DESC_HI_64B = (512u >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "mov.b64 da, {%1, %5};\n"
        "mov.b64 db, {%2, %5};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(DESC_HI_64B)
        : "memory");
}
__device__ __forceinline__ void tc_mma_hw(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hiw, uint32_t b_lo, uint32_t b_hiw,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hiw), "r"(b_lo), "r"(b_hiw), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, %1;\n"
        "@px mov.s32 %0, 1;\n"
        "}\n"
        : "+r"(pred)
        : "r"(0xFFFFFFFFu));
    return pred != 0;
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
__host__ __device__ __forceinline__ float pow2_scale(float m) {
    if (!(m > 0.f) || !(m < 3.0e38f)) return 1.f;
    int e;
    frexpf(m, &e);
    return ldexpf(1.f, 14 - e);
}

constexpr int THREADS = 256;
constexpr int MAX_W_STAGES = 10, A_STAGES = 2;

struct Args {
    const unsigned char *wpack;            // [c-atom][ky*k+kx][hi: Npad x 64 B | lo: Npad x 64 B]
    const float *bias, *scale, *shift;     // [Npad] (scale 1 / shift 0 when the layer has no affine)
    const float *lay;                      // device: layer record (see LAY_*)
    float *out_nchw;                       // (B, N, Ho, Wo) fp32, or null
    __half *out_hi, *out_lo;               // (B, Ho, Wo, Cpad_out) fp16 pair, or null
    int B, k, Ho, Wo, N, Npad, catoms, tiles_x, tiles_y, tiles, relu, Cpad_out;
    int HW, a_half, w_stages;              // halo width (8 + k - 1), bytes of one halo box (1024-rounded), weight ring depth
    int dual;                              // two tiles per weight pass, one MMA issuer warp each (Npad <= 128)
    int cw;                                // channels per K atom: 32 (64-byte rows, SWIZZLE_64B, two K-steps) or 16 (32-byte
                                           // rows, SWIZZLE_32B, one K-step; used when the layer has <= 16 input channels)
};

__global__ void __launch_bounds__(THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap map_hi,
                                                             const __grid_constant__ CUtensorMap map_lo, const Args a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    if ((base & 1023u) != 0) __trap();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rowb = 2u * (uint32_t)a.cw;                               // bytes of one operand row (cw fp16 channels)
    const uint32_t a_stage = 2u * (uint32_t)a.a_half, w_stage = (uint32_t)a.Npad * 2u * rowb;
    const int nq = a.dual ? 2 : 1;                                           // MMA issuers = tiles in flight per weight pass
    const uint32_t a_ring = base, w_ring = base + (uint32_t)nq * A_STAGES * a_stage;   // A ring of issuer q at a_ring + q * A_STAGES * a_stage
    const uint32_t misc = w_ring + (uint32_t)a.w_stages * w_stage;          // barriers, TMEM slot, per-channel vectors
    // barrier table: A full / empty per (issuer, stage), W full / empty per stage, accumulator full / empty per (issuer, buffer)
    enum { B_AFULL = 0, B_AEMPTY = 2 * A_STAGES, B_WFULL = 4 * A_STAGES, B_WEMPTY = B_WFULL + MAX_W_STAGES,
           B_ACCFULL = B_WEMPTY + MAX_W_STAGES, B_ACCEMPTY = B_ACCFULL + 4, B_COUNT = B_ACCEMPTY + 4 };
    auto bar = [&](int i) { return misc + 8u * (uint32_t)i; };
    const uint32_t tmem_slot = misc + 8u * B_COUNT;
    float *vec = reinterpret_cast<float *>(smem_raw + (misc - base) + 8 * B_COUNT + 16);   // [3][Npad]
    const int kk2 = a.k * a.k;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * A_STAGES; ++s) {
            mbar_init(bar(B_AFULL + s), 1);
            mbar_init(bar(B_AEMPTY + s), 1);
        }
        for (int s = 0; s < a.w_stages; ++s) {
            mbar_init(bar(B_WFULL + s), 1);
            mbar_init(bar(B_WEMPTY + s), (uint32_t)nq);      // every issuer releases a weight stage
        }
        for (int b = 0; b < 4; ++b) {
            mbar_init(bar(B_ACCFULL + b), 1);
            mbar_init(bar(B_ACCEMPTY + b), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int n = threadIdx.x; n < a.Npad; n += THREADS) {
        vec[n] = a.bias[n];
        vec[a.Npad + n] = a.scale[n];
        vec[2 * a.Npad + n] = a.shift[n];
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - base));
    const int per_img = a.tiles_x * a.tiles_y;

    if (warp == 0) {
        // ===== producer: one halo box pair per channel group, one weight stage per (channel group, ky, kx) =========
        if (lane == 0) {
            int as = 0, ws = 0;
            uint32_t aph = 0, wph = 0;
            const uint32_t box_bytes = 2u * (uint32_t)(a.HW * (16 + a.k - 1)) * rowb;
            // tiles are taken nq at a time: issuer q works on tile0 + q * gridDim.x with the SAME weight stages
            for (int tile0 = blockIdx.x; tile0 < a.tiles; tile0 += nq * gridDim.x) {
                for (int ca = 0; ca < a.catoms; ++ca) {
                    for (int q = 0; q < nq; ++q) {
                        const int tile = tile0 + q * (int)gridDim.x;
                        if (tile >= a.tiles) break;
                        const int b = tile / per_img, r = tile - b * per_img, ty = r / a.tiles_x, tx = r - ty * a.tiles_x;
                        const int oy0 = ty * 16, ox0 = tx * 8;
                        const int bi = q * A_STAGES + as;
                        mbar_wait(bar(B_AEMPTY + bi), aph ^ 1u);
                        mbar_expect_tx(bar(B_AFULL + bi), box_bytes);
                        const uint32_t dst = a_ring + (uint32_t)bi * a_stage;
                        tma_load_4d(dst, &map_hi, bar(B_AFULL + bi), ca * a.cw, ox0, oy0, b);
                        tma_load_4d(dst + (uint32_t)a.a_half, &map_lo, bar(B_AFULL + bi), ca * a.cw, ox0, oy0, b);
                    }
                    if (++as == A_STAGES) { as = 0; aph ^= 1u; }
                    for (int kk = 0; kk < kk2; ++kk) {
                        mbar_wait(bar(B_WEMPTY + ws), wph ^ 1u);
                        mbar_expect_tx(bar(B_WFULL + ws), w_stage);
                        bulk_load(w_ring + (uint32_t)ws * w_stage, a.wpack + (size_t)(ca * kk2 + kk) * w_stage, w_stage,
                                  bar(B_WFULL + ws));
                        if (++ws == a.w_stages) { ws = 0; wph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1 || (warp == 3 && a.dual)) {
        // ===== MMA issuer q (warp 1: q = 0, warp 3: q = 1): its own A ring and accumulators, shared weight stages =====
        const int q = warp == 3 ? 1 : 0;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(a.Npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        // A: group stride = one halo row (HW pixels of rowb bytes); B: 8 rows of rowb bytes; layout type 4 = SWIZZLE_64B,
        // 6 = SWIZZLE_32B
        const uint32_t lt = a.cw == 32 ? 4u : 6u;
        const uint32_t a_hiw = (((uint32_t)a.HW * rowb) >> 4) | (1u << 14) | (lt << 29), b_hiw = ((8u * rowb) >> 4) | (1u << 14) | (lt << 29);
        const uint32_t w_lo0 = desc_lo(w_ring), w_step = w_stage >> 4, w_lo_off = ((uint32_t)a.Npad * rowb) >> 4;
        const uint32_t px16 = rowb >> 4;           // descriptor units (16 bytes) per pixel row of the operand
        const bool two_steps = a.cw == 32;
        int as = 0, ws = 0;
        uint32_t aph = 0, wph = 0, accph[2] = {0, 0};
        int buf = 0;
        for (int tile0 = blockIdx.x; tile0 < a.tiles; tile0 += nq * gridDim.x) {
            const bool have = tile0 + q * (int)gridDim.x < a.tiles;   // the last group may have no tile for issuer 1
            if (have) {
                mbar_wait(bar(B_ACCEMPTY + 2 * q + buf), accph[buf] ^ 1u);      // the epilogue has drained this accumulator
                tc_fence_after();
            }
            const uint32_t d = tmem + (uint32_t)((2 * q + buf) * a.Npad);
            for (int ca = 0; ca < a.catoms; ++ca) {
                if (!have) {   // keep the shared weight ring moving: release the stages without using them
                    for (int kk = 0; kk < kk2; ++kk) {
                        mbar_wait(bar(B_WFULL + ws), wph);
                        if (lane == 0) mbar_arrive(bar(B_WEMPTY + ws));
                        __syncwarp();
                        if (++ws == a.w_stages) { ws = 0; wph ^= 1u; }
                    }
                    continue;
                }
                mbar_wait(bar(B_AFULL + q * A_STAGES + as), aph);
                // descriptor low words advance by (bytes >> 4): 4 per pixel column (kx), 4 * HW per pixel row (ky);
                // everything per K atom is adds on running values (an issue thread that rebuilds descriptors paces
                // the tensor pipe: r1f finding on the stack kernel)
                const uint32_t ah0 = desc_lo(a_ring + (uint32_t)(q * A_STAGES + as) * a_stage), al0 = ah0 + ((uint32_t)a.a_half >> 4);
                uint32_t row = 0;
                int kk = 0;
                for (int ky = 0; ky < a.k; ++ky, row += px16 * (uint32_t)a.HW) {
                    for (int kx = 0; kx < a.k; ++kx, ++kk) {
                        mbar_wait(bar(B_WFULL + ws), wph);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t a_hi = ah0 + row + px16 * (uint32_t)kx, a_lo = al0 + row + px16 * (uint32_t)kx;
                            const uint32_t w_hi = w_lo0 + (uint32_t)ws * w_step, w_lo = w_hi + w_lo_off;
                            tc_mma_hw(d, a_hi, a_hiw, w_hi, b_hiw, idesc, (ca | kk) != 0);
                            if (two_steps) tc_mma_hw(d, a_hi + 2, a_hiw, w_hi + 2, b_hiw, idesc, 1);
                            tc_mma_hw(d, a_lo, a_hiw, w_hi, b_hiw, idesc, 1);
                            if (two_steps) tc_mma_hw(d, a_lo + 2, a_hiw, w_hi + 2, b_hiw, idesc, 1);
                            tc_mma_hw(d, a_hi, a_hiw, w_lo, b_hiw, idesc, 1);
                            if (two_steps) tc_mma_hw(d, a_hi + 2, a_hiw, w_lo + 2, b_hiw, idesc, 1);
                            tc_commit(bar(B_WEMPTY + ws));
                            if (kk == kk2 - 1) {
                                tc_commit(bar(B_AEMPTY + q * A_STAGES + as));
                                if (ca == a.catoms - 1) tc_commit(bar(B_ACCFULL + 2 * q + buf));
                            }
                        }
                        __syncwarp();
                        if (++ws == a.w_stages) { ws = 0; wph ^= 1u; }
                    }
                }
                if (++as == A_STAGES) { as = 0; aph ^= 1u; }
            }
            if (have) {
                accph[buf] ^= 1u;
                buf ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue ==========================================================================================
        const int q = warp & 3, row = q * 32 + lane, py = row >> 3, px = row & 7;   // TMEM lane = py * 8 + px
        uint32_t accph[2] = {0, 0};
        int buf = 0;
        for (int tile0 = blockIdx.x; tile0 < a.tiles; tile0 += nq * gridDim.x) {
            for (int qq = 0; qq < nq; ++qq) {
                const int tile = tile0 + qq * (int)gridDim.x;
                if (tile >= a.tiles) break;
                const int acc = 2 * qq + buf;                                  // accumulator of issuer qq, buffer buf
                const int b = tile / per_img, r = tile - b * per_img, ty = r / a.tiles_x, tx = r - ty * a.tiles_x;
                const int oy = ty * 16 + py, ox = tx * 8 + px;
                const bool valid = oy < a.Ho && ox < a.Wo;
                const float cinv = a.lay[(size_t)b * LAY_FLOATS + LAY_CINV], s_out = a.lay[(size_t)b * LAY_FLOATS + LAY_SOUT];   // this image's scales
                mbar_wait(bar(B_ACCFULL + acc), accph[buf]);
                tc_fence_after();
                for (int c = 0; c < a.Npad / 32; ++c) {
                    float v[32];
                    tc_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * a.Npad + c * 32), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int n = c * 32 + i;
                        float y = fmaf(fmaf(v[i], cinv, vec[n]), vec[a.Npad + n], vec[2 * a.Npad + n]);
                        v[i] = a.relu ? fmaxf(y, 0.f) : y;
                    }
                    if (valid) {
                        if (a.out_nchw) {
                            float *o = a.out_nchw + (((size_t)b * a.N + c * 32) * a.Ho + oy) * a.Wo + ox;
                            const size_t plane = (size_t)a.Ho * a.Wo;
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (c * 32 + i < a.N) o[(size_t)i * plane] = v[i];
                        } else {
                            const size_t off = (((size_t)b * a.Ho + oy) * a.Wo + ox) * a.Cpad_out + c * 32;
                            uint32_t hi[16], lo[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) split2(v[2 * i] * s_out, v[2 * i + 1] * s_out, hi[i], lo[i]);
                            uint4 *ph4 = reinterpret_cast<uint4 *>(a.out_hi + off), *pl4 = reinterpret_cast<uint4 *>(a.out_lo + off);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                ph4[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                                pl4[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(bar(B_ACCEMPTY + acc));
            }
            accph[buf] ^= 1u;
            buf ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// One layer's operand scales and output bound, chained on the device (lay = this layer's record, nxt = the next one's).
//   in_bound = *in_bound_ptr;  s_in = pow2(in_bound);  sw = pow2(max |W|);  cinv = 1 / (s_in * sw)
//   out_bound = max_n ( |scale_n| (in_bound * sum_k |W[n,k]| + |bias_n|) + |shift_n| );  s_out = pow2(out_bound)
// (two kernels: one block per output channel reduces its filter row, one block combines the rows)
__global__ void __launch_bounds__(128) ctc_row_stats_kernel(const float *__restrict__ w, int K, float *__restrict__ rowstat) {
    __shared__ float red[2][4];
    const int n = blockIdx.x;
    float wmax = 0.f, rs = 0.f;
    for (int kk = threadIdx.x; kk < K; kk += blockDim.x) {
        const float v = fabsf(w[(size_t)n * K + kk]);
        wmax = fmaxf(wmax, v);
        rs += v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        rs += __shfl_xor_sync(0xffffffffu, rs, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = wmax; red[1][threadIdx.x >> 5] = rs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        rowstat[2 * n] = fmaxf(fmaxf(red[0][0], red[0][1]), fmaxf(red[0][2], red[0][3]));
        rowstat[2 * n + 1] = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
    }
}
// One block per IMAGE: the operand scales are chained per image from ITS max |x| (a sample's activations never depend on
// its batch-mates; the weight scale sw is the same in every record).  in_bound of image b = in_bound_ptr[b * in_stride].
__global__ void __launch_bounds__(256) ctc_layer_stats_kernel(const float *__restrict__ rowstat, const float *__restrict__ vecs,
                                                              int N, int Npad, const float *__restrict__ in_bound_ptr,
                                                              int in_stride, int tensor_in, float *__restrict__ lay_all,
                                                              float *__restrict__ nxt_all) {
    __shared__ float red[2][256];
    const int b = blockIdx.x;
    float *lay = lay_all + (size_t)b * LAY_FLOATS, *nxt = nxt_all + (size_t)b * LAY_FLOATS;
    const float in_bound = in_bound_ptr[(size_t)b * in_stride];
    float wmax = 0.f, ob = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        wmax = fmaxf(wmax, rowstat[2 * n]);
        const float rs = rowstat[2 * n + 1];
        const float o = fabsf(vecs[Npad + n]) * (in_bound * rs * 1.0001f + fabsf(vecs[n])) + fabsf(vecs[2 * Npad + n]);
        ob = fmaxf(ob, o);
    }
    red[0][threadIdx.x] = wmax; red[1][threadIdx.x] = ob;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            red[0][threadIdx.x] = fmaxf(red[0][threadIdx.x], red[0][threadIdx.x + o]);
            red[1][threadIdx.x] = fmaxf(red[1][threadIdx.x], red[1][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float s_in = tensor_in ? pow2_scale(in_bound) : 1.f, sw = tensor_in ? pow2_scale(red[0][0]) : 1.f;
        const float out_bound = red[1][0] * 1.0001f, s_out = pow2_scale(out_bound);
        lay[LAY_INBOUND] = in_bound; lay[LAY_SIN] = s_in; lay[LAY_SW] = sw; lay[LAY_CINV] = 1.f / (s_in * sw);
        lay[LAY_OUTBOUND] = out_bound; lay[LAY_SOUT] = s_out;
        nxt[LAY_INBOUND] = out_bound;
    }
}

// filter (N, C, k, k) fp32 -> per atom (cw-channel group, then ky, kx) an Npad x (2 cw)-byte hi image followed by the lo
// image, swizzled for the UMMA descriptor, scaled by sw; rows >= N and channels >= C are zero
__global__ void ctc_pack_kernel(const float *__restrict__ w, const float *__restrict__ lay, int N, int C, int k, int Npad,
                                int catoms, int cw, unsigned char *__restrict__ out) {
    const float sw = lay[LAY_SW];
    const long long total = (long long)k * k * catoms * Npad * cw;
    const size_t rowb = 2 * (size_t)cw;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ks = (int)(t % cw), n = (int)((t / cw) % Npad);
        const int atom = (int)(t / ((long long)cw * Npad));
        const int ca = atom / (k * k), kk = atom - ca * (k * k), ky = kk / k, kx = kk - ky * k;
        const int c = ca * cw + ks;
        const float v = (n < N && c < C) ? w[(((size_t)n * C + c) * k + ky) * k + kx] * sw : 0.f;
        const __half hi = __float2half_rn(v), lo = __float2half_rn(v - __half2float(hi));
        const size_t stage = (size_t)Npad * 2 * rowb;
        // 16-byte chunk index XOR: SWIZZLE_64B (cw = 32): (n >> 1) & 3;  SWIZZLE_32B (cw = 16): (n >> 2) & 1
        const int x = cw == 32 ? ((n >> 1) & 3) : ((n >> 2) & 1);
        const size_t off = (size_t)n * rowb + (size_t)((((ks >> 3) ^ x) << 4) | ((ks & 7) << 1));
        *reinterpret_cast<__half *>(out + (size_t)atom * stage + off) = hi;
        *reinterpret_cast<__half *>(out + (size_t)atom * stage + (size_t)Npad * rowb + off) = lo;
    }
}

// network input (B, C, H, W) fp32 -> the fp16 hi / lo NHWC pair with C padded to Cpad (zeros), scaled by s_in = pow2(max |x|)
// (writes the first layer's record fields it needs: LAY_INBOUND is set by the caller's stats chain)
__global__ void __launch_bounds__(256) ctc_input_split_kernel(const float *__restrict__ x, const float *__restrict__ absmax,
                                                              int B, int C, int H, int W, int Cpad, __half *__restrict__ hi,
                                                              __half *__restrict__ lo) {
    const size_t npix = (size_t)B * H * W, plane = (size_t)H * W;
    const int pairs = Cpad / 2;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < npix * pairs; t += (size_t)gridDim.x * blockDim.x) {
        const size_t pix = t / pairs;
        const int c = 2 * (int)(t - pix * pairs);
        const size_t b = pix / plane, r = pix - b * plane;
        const float s = pow2_scale(absmax[b]);          // per image
        const float v0 = c < C ? x[(b * C + c) * plane + r] * s : 0.f;
        const float v1 = c + 1 < C ? x[(b * C + c + 1) * plane + r] * s : 0.f;
        uint32_t h, l;
        split2(v0, v1, h, l);
        reinterpret_cast<uint32_t *>(hi)[pix * pairs + (c >> 1)] = h;
        reinterpret_cast<uint32_t *>(lo)[pix * pairs + (c >> 1)] = l;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
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
// (C, W, H, B) fp16 NHWC tensor, box 32 x HW x HH x 1 (the 8 x 16 output patch with its halo), SWIZZLE_64B
static int make_act_map(CUtensorMap *m, const void *ptr, int Cpad, int W, int H, int B, int HW, int HH, int cw) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        return (int)cudaErrorNotSupported;
    }
    const cuuint64_t gdim[4] = {(cuuint64_t)Cpad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t gstride[3] = {(cuuint64_t)Cpad * 2, (cuuint64_t)W * Cpad * 2, (cuuint64_t)H * W * Cpad * 2};
    const cuuint32_t box[4] = {(cuuint32_t)cw, (cuuint32_t)HW, (cuuint32_t)HH, 1}, estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptr), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, cw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (C=%d W=%d H=%d B=%d)", (int)r, Cpad, W, H, B);
        return (int)cudaErrorInvalidValue;
    }
    return 0;
}

}  // namespace ctc

bool ctc_eligible(int Npad, int num_layers) {
    if (getenv("EQB_CONV_NO_TC")) return false;
    return num_layers >= 2 && Npad >= 32 && Npad <= 256;
}

// channels per K atom for an operand whose channel count is padded to Cpad: 32 (64-byte rows) or 16 (32-byte rows)
int ctc_atom_channels(int Cpad) { return Cpad % 32 == 0 ? 32 : 16; }

size_t ctc_pack_bytes(int Npad, int Cpad, int k) { return (size_t)k * k * Cpad * Npad * 4; }

int ctc_layer_stats(const float *w, const float *vecs, int N, int Npad, int K, const float *in_bound_ptr, int in_stride, int B,
                    int tensor_in, float *lay, float *nxt, float *rowstat, cudaStream_t st) {
    ctc::ctc_row_stats_kernel<<<N, 128, 0, st>>>(w, K, rowstat);
    ctc::ctc_layer_stats_kernel<<<B, 256, 0, st>>>(rowstat, vecs, N, Npad, in_bound_ptr, in_stride, tensor_in, lay, nxt);
    return finish_launch("ctc_layer_stats_kernel");
}

int ctc_input_split(const float *x, const float *absmax, int B, int C, int H, int W, int Cpad, __half *hi, __half *lo,
                    cudaStream_t st) {
    const size_t work = (size_t)B * H * W * (Cpad / 2);
    size_t blocks = (work + 255) / 256;
    if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
    if (blocks == 0) return 0;
    ctc::ctc_input_split_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, absmax, B, C, H, W, Cpad, hi, lo);
    return finish_launch("ctc_input_split_kernel");
}

int ctc_pack(const float *w, const float *lay, int N, int C, int Cpad, int k, int Npad, unsigned char *out, cudaStream_t st) {
    const int cw = ctc_atom_channels(Cpad);
    ctc::ctc_pack_kernel<<<256, 256, 0, st>>>(w, lay, N, C, k, Npad, Cpad / cw, cw, out);
    return finish_launch("ctc_pack_kernel");
}

int ctc_conv_layer(const __half *in_hi, const __half *in_lo, int B, int Cpad, int H, int W, int k, const unsigned char *wpack,
                   const float *vecs, const float *lay, int N, int Npad, int relu, float *out_nchw, __half *out_hi,
                   __half *out_lo, int Cpad_out, cudaStream_t st) {
    const int HW = 8 + k - 1, HH = 16 + k - 1;
    EQB_UNSUPPORTED(HW > 256 || HH > 256, "eqb_conv_stack (tcgen05): kernel size %d too large for one TMA box", k);
    CUtensorMap mh, ml;
    const int cw = ctc_atom_channels(Cpad);
    int e = ctc::make_act_map(&mh, in_hi, Cpad, W, H, B, HW, HH, cw);
    if (e) return e;
    e = ctc::make_act_map(&ml, in_lo, Cpad, W, H, B, HW, HH, cw);
    if (e) return e;
    ctc::Args a{};
    a.wpack = wpack; a.bias = vecs; a.scale = vecs + Npad; a.shift = vecs + 2 * Npad; a.lay = lay;
    a.out_nchw = out_nchw; a.out_hi = out_hi; a.out_lo = out_lo;
    a.B = B; a.k = k; a.Ho = H - k + 1; a.Wo = W - k + 1; a.N = N; a.Npad = Npad;
    a.catoms = Cpad / cw; a.cw = cw;
    a.tiles_x = (a.Wo + 7) / 8; a.tiles_y = (a.Ho + 15) / 16; a.tiles = B * a.tiles_x * a.tiles_y;
    a.relu = relu; a.Cpad_out = Cpad_out;
    a.HW = HW;
    a.a_half = (HW * HH * 2 * cw + 1023) & ~1023;
    // two tiles per weight pass (one issuer warp each) when four accumulators fit in TMEM: halves the weight stream per
    // pixel and doubles the MMA issue rate, which paces the N = 128 case
    a.dual = Npad <= 128 && !getenv("EQB_CONV_TC_SINGLE");
    const size_t a_bytes = (size_t)(a.dual ? 2 : 1) * ctc::A_STAGES * 2 * a.a_half, w_stage = (size_t)Npad * 4 * cw;
    const size_t misc = 8 * (4 * ctc::A_STAGES + 2 * ctc::MAX_W_STAGES + 8) + 16 + 3 * (size_t)Npad * sizeof(float);
    int w_stages = (int)((227 * 1024 - misc - a_bytes) / w_stage);
    if (w_stages > ctc::MAX_W_STAGES) w_stages = ctc::MAX_W_STAGES;
    EQB_UNSUPPORTED(w_stages < 2, "eqb_conv_stack (tcgen05): operand rings do not fit in shared memory (k = %d, N = %d)", k, Npad);
    a.w_stages = w_stages;
    const size_t smem = a_bytes + (size_t)w_stages * w_stage + misc;
    static PerDeviceOnce configured;
    if (configured.first()) {
        EQB_CUDA(cudaFuncSetAttribute(ctc::conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    if (a.tiles == 0) return 0;
    const int grid = a.tiles < num_sms() ? a.tiles : num_sms();
    ctc::conv_tc_kernel<<<grid, ctc::THREADS, smem, st>>>(mh, ml, a);
    return finish_launch("conv_tc_kernel");
}

}  // namespace eqb
