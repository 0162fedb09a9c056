// minimal TMA sanity test: 3-D box of doubles, negative start coordinate, OOB zero fill
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#ifndef EXV
#define EXV 36
#endif
constexpr int EX = EXV, EY = 8, EZ = 6, VOL = EX * EY * EZ;
__global__ void k(const __grid_constant__ CUtensorMap tm, double *out, int x, int y, int z) {
    extern __shared__ __align__(128) unsigned char raw[];
    double *sm = (double *)raw;
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"((unsigned)(VOL * 8)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"((unsigned)__cvta_generic_to_shared(sm)), "l"(&tm), "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WL;\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < VOL; i += blockDim.x) out[i] = sm[i];
}
int main() {
    const int n0 = 258, n1 = 257, n2 = 258;
    std::vector<double> h((size_t)n0 * n1 * n2);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)i;
    double *d, *o;
    cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, VOL * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry %d %d %p\n", (int)e, (int)q, p);
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)n0, (cuuint64_t)n1, (cuuint64_t)n2}, strides[2] = {(cuuint64_t)n0 * 8, (cuuint64_t)n0 * n1 * 8};
    cuuint32_t box[3] = {EX, EY, EZ}, es[3] = {1, 1, 1};
    CUresult r = ((enc_fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, VOL * 8);
    k<<<1, 256, VOL * 8>>>(tm, o, -1, 3, 5);
    e = cudaDeviceSynchronize();
    printf("sync %s\n", cudaGetErrorString(e));
    std::vector<double> ho(VOL);
    cudaMemcpy(ho.data(), o, VOL * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int iz = 0; iz < EZ; iz++) for (int iy = 0; iy < EY; iy++) for (int ix = 0; ix < EX; ix++) {
        int gx = -1 + ix, gy = 3 + iy, gz = 5 + iz;
        double want = (gx < 0 || gx >= n0) ? 0.0 : (double)((size_t)gx + (size_t)n0 * (gy + (size_t)n1 * gz));
        if (ho[ix + EX * (iy + EY * iz)] != want) bad++;
    }
    printf("mismatches %d\n", bad);
    return 0;
}
