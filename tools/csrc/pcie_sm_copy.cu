// Micro-benchmark: host<->device transfers driven by SMs (kernels that read / write mapped pinned host memory)
// against the copy engines (cudaMemcpyAsync), alone and in both directions at once, whole and in 8 chunks.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/pcie_sm_copy tools/csrc/pcie_sm_copy.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>

__global__ void k_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) dst[i] = src[i];
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

int main(int argc, char** argv) {
    const size_t NI = 100u << 20, NO = 112u << 20;
    int blocks = argc > 1 ? atoi(argv[1]) : 64, threads = argc > 2 ? atoi(argv[2]) : 256;
    uint8_t *hi, *ho, *di, *dout;
    CK(cudaHostAlloc(&hi, NI, cudaHostAllocPortable | cudaHostAllocMapped));
    CK(cudaHostAlloc(&ho, NO, cudaHostAllocPortable | cudaHostAllocMapped));
    memset(hi, 1, NI); memset(ho, 2, NO);
    CK(cudaMalloc(&di, NI)); CK(cudaMalloc(&dout, NO));
    CK(cudaMemset(dout, 3, NO));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    auto run = [&](const char* name, bool sm, bool in, bool out, int chunks) {
        double best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaDeviceSynchronize());
            auto t0 = std::chrono::steady_clock::now();
            for (int k = 0; k < chunks; ++k) {
                size_t ci = NI / chunks, co = NO / chunks;
                if (in) {
                    if (sm) k_copy<<<blocks, threads, 0, s1>>>((const uint4*)(hi + k * ci), (uint4*)(di + k * ci), ci / 16);
                    else CK(cudaMemcpyAsync(di + k * ci, hi + k * ci, ci, cudaMemcpyHostToDevice, s1));
                }
                if (out) {
                    if (sm) k_copy<<<blocks, threads, 0, s2>>>((const uint4*)(dout + k * co), (uint4*)(ho + k * co), co / 16);
                    else CK(cudaMemcpyAsync(ho + k * co, dout + k * co, co, cudaMemcpyDeviceToHost, s2));
                }
            }
            CK(cudaDeviceSynchronize());
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (ms < best) best = ms;
        }
        double bytes = (in ? NI : 0) + (out ? NO : 0);
        printf("%-28s %s chunks %d: %.3f ms  %.1f GB/s\n", name, sm ? "SM " : "DMA", chunks, best, bytes / best / 1e6);
    };
    printf("blocks %d threads %d\n", blocks, threads);
    for (int sm = 0; sm < 2; ++sm) {
        run("h2d alone", sm, true, false, 1);
        run("d2h alone", sm, false, true, 1);
        run("both", sm, true, true, 1);
        run("both", sm, true, true, 8);
        run("both", sm, true, true, 32);
    }
    if (ho[5] != 3 || ho[NO - 1] != 3) printf("BAD COPY\n");
    return 0;
}
