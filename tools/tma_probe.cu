// Probe which TMA box configurations run on this GPU: tma_probe W H boxw boxh swizzle(0|3) x y
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, float *out, int n, int x, int y, int z) {
    extern __shared__ unsigned char raw[];
    uint32_t base = ((uint32_t)__cvta_generic_to_shared(raw) + 1023u) & ~1023u;
    float *sm = (float *)(raw + (base - (uint32_t)__cvta_generic_to_shared(raw)));
    uint32_t bar = base + 64 * 1024;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n * 4));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(base), "l"((uint64_t)&map), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(0));
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char **argv) {
    int W = atoi(argv[1]), H = atoi(argv[2]), bw = atoi(argv[3]), bh = atoi(argv[4]), sw = atoi(argv[5]), x = atoi(argv[6]), y = atoi(argv[7]);
    int planes = 6;
    float *src, *out;
    cudaMalloc(&src, (size_t)W * H * planes * 4);
    cudaMalloc(&out, 64 * 1024);
    float *h = (float *)malloc((size_t)W * H * planes * 4);
    for (int i = 0; i < W * H * planes; ++i) h[i] = (float)i;
    cudaMemcpy(src, h, (size_t)W * H * planes * 4, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, src, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    (CUtensorMapSwizzle)sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("W=%d H=%d box=%dx%d swz=%d at (%d,%d): encode=%d ", W, H, bw, bh, sw, x, y, (int)r);
    if (r) { printf("\n"); return 1; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    probe<<<1, 128, 66 * 1024>>>(m, out, bw * bh, x, y, 1);
    cudaError_t e = cudaDeviceSynchronize();
    float o[4] = {0, 0, 0, 0};
    if (!e) cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost);
    printf("run=%s first=%g (expect %g) second=%g\n", cudaGetErrorString(e), o[0], (x >= 0 && y >= 0 && x < W && y < H) ? (float)(W * H + y * W + x) : 0.f, o[1]);
    return 0;
}
