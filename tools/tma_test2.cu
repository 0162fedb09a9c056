// TMA sanity test with the libcu++ wrappers (CUDA programming guide pattern)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int EX = 36, EY = 8, EZ = 6, VOL = EX * EY * EZ;
__global__ void k(const __grid_constant__ CUtensorMap tm, double *out, int x, int y, int z) {
    __shared__ alignas(128) double sm[VOL];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_3d_global_to_shared(&sm, &tm, x, y, z, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(sm));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < VOL; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char **argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int n0 = 258, n1 = 257, n2 = 258;
    std::vector<double> h((size_t)n0 * n1 * n2);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)i;
    double *d, *o;
    cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, VOL * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)n0, (cuuint64_t)n1, (cuuint64_t)n2}, strides[2] = {(cuuint64_t)n0 * 8, (cuuint64_t)n0 * n1 * 8};
    cuuint32_t box[3] = {EX, EY, EZ}, es[3] = {1, 1, 1};
    CUresult r = ((enc_fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    int x0 = (variant & 1) ? 0 : -1;
    if (variant & 2) {  // describe the same memory as uint64 elements
        r = ((enc_fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode u64 %d\n", (int)r);
    }
    if (variant & 4) {  // float32 view: inner dim doubled
        cuuint64_t d4[3] = {(cuuint64_t)n0 * 2, (cuuint64_t)n1, (cuuint64_t)n2}; cuuint32_t b4[3] = {EX * 2, EY, EZ};
        r = ((enc_fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, d4, strides, b4, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode f32 %d\n", (int)r); x0 *= 2;
    }
    k<<<1, 256>>>(tm, o, x0, 3, 5);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync %s\n", cudaGetErrorString(e));
    std::vector<double> ho(VOL);
    cudaMemcpy(ho.data(), o, VOL * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int iz = 0; iz < EZ; iz++) for (int iy = 0; iy < EY; iy++) for (int ix = 0; ix < EX; ix++) {
        int gx = ((variant & 1) ? 0 : -1) + ix, gy = 3 + iy, gz = 5 + iz;
        double want = (gx < 0 || gx >= n0) ? 0.0 : (double)((size_t)gx + (size_t)n0 * (gy + (size_t)n1 * gz));
        if (ho[ix + EX * (iy + EY * iz)] != want) bad++;
    }
    printf("mismatches %d\n", bad);
    return 0;
}
