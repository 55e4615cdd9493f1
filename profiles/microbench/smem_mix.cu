// smem_mix.cu — do warp shuffles / redux share the shared-memory wavefront pipe with LDS?
// One CTA per SM, 16 warps.  Each variant issues N ops per iteration; prints cycles per warp-op per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
template <int OP>
__global__ void k(int* out, long long* cyc) {
    __shared__ int sm[32 * 64];
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) sm[i] = i;
    __syncthreads();
    int lane = threadIdx.x & 31;
    int a = lane, b = lane * 3, c = 0, d = 1;
    unsigned addr = (unsigned)__cvta_generic_to_shared(sm) + lane * 4;
    long long t0 = clock64();
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == 0 || OP == 3 || OP == 4) { int v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr + j * 128)); c += v; }
            if (OP == 1 || OP == 3) a += __shfl_xor_sync(0xffffffffu, a, 1 + (j & 3));
            if (OP == 2 || OP == 4) b += __reduce_add_sync(0xffffffffu, b + j);
            if (OP == 5) { d = __vadd2(d, a); a ^= d; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, int nops) {
    int* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    k<OP><<<148, 512>>>(out, cyc); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%-28s %7.2f SM-cycles per warp-op (16 warps/SM, %d ops/iter)\n", name, avg / (16.0 * ITER * 8 * nops) , nops);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("LDS.32 (1 wavefront)", 1);
    run<1>("SHFL.BFLY", 1);
    run<2>("REDUX.SUM", 1);
    run<3>("LDS.32 + SHFL", 2);
    run<4>("LDS.32 + REDUX", 2);
    return 0;
}
