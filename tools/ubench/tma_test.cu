// standalone check of the 4-D TMA tile load used by csrc/dwt.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ CUtensorMap tmap, float *out, int bw, int bh, int x, int y, int z, int w)
{
    extern __shared__ __align__(128) float smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 32768);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bw * bh * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(smem_u32(smem)), "l"(&tmap), "r"(x), "r"(y), "r"(z), "r"(w), "r"(smem_u32(bar)) : "memory");
    }
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
    } while (!done);
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = smem[i];
}
int main(int argc, char **argv)
{
    const int cols = 348, rows = 308, planes = 2, subs = 4, pitch = 348;
    const int bw = argc > 1 ? atoi(argv[1]) : 132, bh = argc > 2 ? atoi(argv[2]) : 66;
    size_t plane_stride = (size_t)pitch * rows;
    std::vector<float> h(plane_stride * planes * subs);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Fn fn = (Fn)p;
    CUtensorMap map;
    cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes, (cuuint64_t)subs};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, plane_stride * 4, plane_stride * 4 * planes};
    cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    float *out; cudaMalloc(&out, bw * bh * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
    k<<<1, 256, 140000>>>(map, out, bw, bh, 10, 20, 1, 2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s\n", cudaGetErrorString(e));
    std::vector<float> o(bw * bh); cudaMemcpy(o.data(), out, bw * bh * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < bh; ++yy) for (int xx = 0; xx < bw; ++xx) {
        size_t src = (size_t)2 * plane_stride * planes + 1 * plane_stride + (size_t)(20 + yy) * pitch + 10 + xx;
        if (o[yy * bw + xx] != h[src]) ++bad;
    }
    printf("mismatches: %d\n", bad);
    return 0;
}
