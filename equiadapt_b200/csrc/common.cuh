// Shared helpers for the equiadapt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <nvtx3/nvToolsExt.h>

#include "../../include/equiadapt_b200.h"

namespace eqb {

void set_error(const char *fmt, ...);

inline int finish_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

#define EQB_REQUIRE(cond, ...)              \
    do {                                    \
        if (!(cond)) {                      \
            eqb::set_error(__VA_ARGS__);    \
            return EQB_ERR_INVALID;         \
        }                                   \
    } while (0)

#define EQB_UNSUPPORTED(cond, ...)          \
    do {                                    \
        if (cond) {                         \
            eqb::set_error(__VA_ARGS__);    \
            return EQB_ERR_UNSUPPORTED;     \
        }                                   \
    } while (0)

#define EQB_CUDA(call)                                                     \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) {                                          \
            eqb::set_error("%s: %s", #call, cudaGetErrorString(e__));      \
            return (int)e__;                                               \
        }                                                                  \
    } while (0)

// NVTX range named after the C-ABI entry point: one range per call of the boundary in an nsys / ncu --nvtx timeline
// (header-only NVTX3: a no-op costing one predictable branch when no tool is attached).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
#define EQB_NVTX_RANGE() eqb::NvtxRange eqb_nvtx_range__(__func__)

inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev;
}

constexpr int EQB_MAX_DEVICES = 64;

// Per-DEVICE one-time setup (function attributes, device symbols): a process may drive several GPUs, and
// cudaFuncSetAttribute / cudaMemcpyToSymbol act on the current device only.
struct PerDeviceOnce {
    bool done[EQB_MAX_DEVICES] = {};
    bool first() {
        const int dev = current_device();
        if (dev >= EQB_MAX_DEVICES) return true;   // beyond the table: configure every time (cheap)
        if (done[dev]) return false;
        done[dev] = true;
        return true;
    }
};

inline int num_sms() {
    static int n[EQB_MAX_DEVICES] = {};
    const int dev = current_device();
    int v = dev < EQB_MAX_DEVICES ? n[dev] : 0;
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        if (dev < EQB_MAX_DEVICES) n[dev] = v;
    }
    return v;
}

// 2x2 matrix + flags describing one discrete group element's action on pixel coordinates:
//   src = centre + A * (dst - centre)
struct Affine2 {
    double a00, a01, a10, a11;
    int exact;  // all entries in {0,+-1}: a pure index permutation when the offsets are integral
};

// cos/sin of (sign * 2*pi*r/N), exact at quarter turns (sincospi is exact at multiples of 1/2).
__device__ __forceinline__ void rot_cs(int r, int N, double sign, double &c, double &s) {
    sincospi(sign * 2.0 * (double)r / (double)N, &s, &c);
}

__device__ __forceinline__ float ld_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void st_stream(float *p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v));
}

}  // namespace eqb
