// TMA tile-load probe: tma_test2 <rank 2|4> <box_w> <box_h> <x> <y>
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_test2 tma_test2.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tmap, float *out, int n, int x, int y, int z, int w)
{
    extern __shared__ __align__(1024) unsigned char raw[];
    float *smem = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(raw) + 127) & ~(uintptr_t)127);
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(n * 4) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(smem)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(smem_u32(smem)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = smem[i];
}

int main(int argc, char **argv)
{
    const int rank = argc > 1 ? atoi(argv[1]) : 2;
    const int bw = argc > 2 ? atoi(argv[2]) : 64, bh = argc > 3 ? atoi(argv[3]) : 32;
    const int x = argc > 4 ? atoi(argv[4]) : 8, y = argc > 5 ? atoi(argv[5]) : 20;
    const int cols = 348, rows = 308, planes = 2, subs = 4, pitch = 348;
    const int z = 1, w = 2;
    size_t plane_stride = (size_t)pitch * rows;
    std::vector<float> h(plane_stride * planes * subs);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(ge), (int)q, p);
    Fn fn = (Fn)p;
    alignas(64) CUtensorMap map;
    cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes, (cuuint64_t)subs};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, plane_stride * 4, plane_stride * 4 * planes};
    cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank=%d box=%dx%d at (%d,%d): encode rc=%d\n", rank, bw, bh, x, y, (int)r);
    const int n = bw * bh;
    float *out; cudaMalloc(&out, n * 4);
    const size_t smem = (size_t)n * 4 + 1024;
    if (rank == 2) {
        cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<2><<<1, 128, smem>>>(map, out, n, x, y, z, w);
    } else {
        cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<4><<<1, 128, smem>>>(map, out, n, x, y, z, w);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> o(n); cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < bh; ++yy) for (int xx = 0; xx < bw; ++xx) {
        const int gy = y + yy, gx = x + xx;
        float ref = 0.f;
        if (gy >= 0 && gy < rows && gx >= 0 && gx < cols) {
            size_t src = (rank == 4 ? (size_t)w * plane_stride * planes + (size_t)z * plane_stride : 0) + (size_t)gy * pitch + gx;
            ref = h[src];
        }
        if (o[yy * bw + xx] != ref) ++bad;
    }
    printf("mismatches: %d of %d\n", bad, n);
    return bad != 0;
}
