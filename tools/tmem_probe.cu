// Probe: TMEM -> register read throughput (tcgen05.ld 32x32b.x32) on one SM as a function of the number of reading warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_probe tools/tmem_probe.cu && tools/tmem_probe
// Each warp reads its own 32 lanes x 32 columns (4 KB) `iters` times back to back (wait::ld after each, or after every 4);
// reports cycles per load and bytes / cycle for 1 .. 16 warps (warps w and w+4 share a scheduler and a lane quadrant).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(512, 1) probe(long long *out, int nwarps, int iters, int batch) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
        __syncwarp();
        t0 = clock64();
        for (int it = 0; it < iters; it += batch) {
            for (int b = 0; b < batch; ++b) {
                uint32_t r[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(addr)
                    : "memory");
                if (b == batch - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= r[0] ^ r[13] ^ r[31];
            }
        }
        t1 = clock64();
    }
    __syncthreads();
    if (lane == 0 && warp < nwarps) out[warp] = t1 - t0;
    if (acc == 0x12345678u) out[63] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    long long *d, h[64];
    cudaMalloc(&d, 64 * sizeof(long long));
    const int iters = 4096;
    for (int batch = 1; batch <= 4; batch *= 4)
        for (int nw = 1; nw <= 16; nw *= 2) {
            cudaMemset(d, 0, 64 * sizeof(long long));
            probe<<<1, 512>>>(d, nw, iters, batch);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
            printf("wait every %d loads, %2d warps: %.1f cycles per 4 KB load per warp, %.1f B/clk per SM\n", batch, nw, (double)mx / iters,
                   (double)nw * iters * 4096.0 / mx);
        }
    return 0;
}
