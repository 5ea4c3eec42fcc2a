// Mock of spl_encode_batch's three-stream pipeline, to find what keeps the copy engines below the 95 GB/s duplex
// rate this box sustains: 8 chunks of 12.5 MB in / 14 MB out; flags switch the ingredients on one by one.
//   bit0 timing events on s_in after every chunk      bit1 small (100 KB) copies next to the big ones
//   bit2 a kernel per chunk between in and out (wait on ev_in, D2H enqueued by the host after ev_done)
//   bit3 events created with cudaEventDisableTiming    bit4 memset (3 MB) on the kernel stream per chunk
//   bit5 kernel writes one word to mapped host memory
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_work(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16, int passes, unsigned long long* host_word) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            uint4 a = src[i]; a.x += p; dst[i] = a;
        }
    if (host_word && blockIdx.x == 0 && threadIdx.x == 0) *host_word = n16;
}

int main(int argc, char** argv) {
    const int C = 8;
    const size_t CI = 12500000 / 16 * 16, CO = 14000000 / 16 * 16, SMALL = 100000;
    uint8_t *hi, *ho, *hs, *di, *dout, *dz;
    unsigned long long* hmeta;
    CK(cudaHostAlloc(&hi, CI * C, cudaHostAllocPortable | cudaHostAllocMapped));
    CK(cudaHostAlloc(&ho, CO * C, cudaHostAllocPortable | cudaHostAllocMapped));
    CK(cudaHostAlloc(&hs, SMALL * 2 * C, cudaHostAllocPortable | cudaHostAllocMapped));
    CK(cudaHostAlloc(&hmeta, 4096, cudaHostAllocPortable | cudaHostAllocMapped));
    memset(hi, 1, CI * C); memset(ho, 2, CO * C);
    CK(cudaMalloc(&di, CI * C + SMALL * C)); CK(cudaMalloc(&dout, CO * C + SMALL * C)); CK(cudaMalloc(&dz, 4 << 20));
    cudaStream_t s_in, s_k, s_out;
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (int flags : {0, 1, 1 | 8, 2, 4, 4 | 1, 4 | 1 | 2, 4 | 1 | 2 | 16, 4 | 1 | 2 | 16 | 32, 4 | 8 | 2 | 16 | 32, 4 | 8}) {
        cudaEvent_t ev_in[C], ev_done[C];
        for (int k = 0; k < C; ++k) {
            CK(cudaEventCreateWithFlags(&ev_in[k], (flags & 8) ? cudaEventDisableTiming : cudaEventDefault));
            CK(cudaEventCreateWithFlags(&ev_done[k], (flags & 8) ? cudaEventDisableTiming : cudaEventDefault));
        }
        double best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaDeviceSynchronize());
            auto t0 = std::chrono::steady_clock::now();
            if (!(flags & 4)) {
                for (int k = 0; k < C; ++k) {
                    CK(cudaMemcpyAsync(di + k * CI, hi + k * CI, CI, cudaMemcpyHostToDevice, s_in));
                    if (flags & 2) CK(cudaMemcpyAsync(di + C * CI + k * SMALL, hs + k * SMALL, SMALL, cudaMemcpyHostToDevice, s_in));
                    if (flags & 1) CK(cudaEventRecord(ev_in[k], s_in));
                    if (flags & 2) CK(cudaMemcpyAsync(hs + (C + k) * SMALL, dout + C * CO + k * SMALL, SMALL, cudaMemcpyDeviceToHost, s_out));
                    CK(cudaMemcpyAsync(ho + k * CO, dout + k * CO, CO, cudaMemcpyDeviceToHost, s_out));
                }
            } else {
                int next_out = 0;
                auto drain = [&](int k) {
                    if (flags & 2) CK(cudaMemcpyAsync(hs + (C + k) * SMALL, dout + C * CO + k * SMALL, SMALL, cudaMemcpyDeviceToHost, s_out));
                    CK(cudaMemcpyAsync(ho + k * CO, dout + k * CO, CO, cudaMemcpyDeviceToHost, s_out));
                };
                for (int k = 0; k < C; ++k) {
                    CK(cudaMemcpyAsync(di + k * CI, hi + k * CI, CI, cudaMemcpyHostToDevice, s_in));
                    if (flags & 2) CK(cudaMemcpyAsync(di + C * CI + k * SMALL, hs + k * SMALL, SMALL, cudaMemcpyHostToDevice, s_in));
                    CK(cudaEventRecord(ev_in[k], s_in));
                    CK(cudaStreamWaitEvent(s_k, ev_in[k], 0));
                    if (flags & 16) CK(cudaMemsetAsync(dz, 0, 3 << 20, s_k));
                    k_work<<<148 * 4, 256, 0, s_k>>>((const uint4*)(di + k * CI), (uint4*)(dout + k * CO), CI / 16, 6, (flags & 32) ? hmeta + k : nullptr);
                    CK(cudaEventRecord(ev_done[k], s_k));
                    while (next_out <= k && cudaEventQuery(ev_done[next_out]) == cudaSuccess) drain(next_out++);
                }
                while (next_out < C) { CK(cudaEventSynchronize(ev_done[next_out])); drain(next_out++); }
            }
            CK(cudaDeviceSynchronize());
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (ms < best) best = ms;
        }
        printf("flags %2d [%s%s%s%s%s%s]: %.3f ms  %.1f GB/s total\n", flags, (flags & 1) ? "ev " : "", (flags & 8) ? "notiming " : "",
               (flags & 2) ? "small " : "", (flags & 4) ? "kernel-dep " : "", (flags & 16) ? "memset " : "", (flags & 32) ? "hostword " : "",
               best, (double)(CI + CO) * C / best / 1e6);
        for (int k = 0; k < C; ++k) { cudaEventDestroy(ev_in[k]); cudaEventDestroy(ev_done[k]); }
    }
    return 0;
}
