// pinned H2D of 635 MB in 40 MB pieces on one stream vs alternating on two / three streams, with a 568 MB D2H running
// concurrently on its own stream (the traffic of one C2 e2e step)
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
int main() {
    const size_t in_b = 635040000, out_b = 567832203, piece = 40u << 20;
    char *h_in, *h_out, *d_in, *d_out;
    cudaMallocHost(&h_in, in_b); cudaMallocHost(&h_out, out_b);
    cudaMalloc(&d_in, in_b); cudaMalloc(&d_out, out_b);
    cudaStream_t s[4], sd;
    for (auto &x : s) cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ns = 1; ns <= 3; ns++)
        for (int with_d2h = 0; with_d2h <= 1; with_d2h++) {
            float best = 1e9f;
            for (int rep = 0; rep < 6; rep++) {
                cudaDeviceSynchronize();
                cudaEventRecord(e0, s[0]);
                for (int i = 1; i < ns; i++) cudaStreamWaitEvent(s[i], e0, 0);
                cudaStreamWaitEvent(sd, e0, 0);
                int k = 0;
                for (size_t o = 0; o < in_b; o += piece, k++)
                    cudaMemcpyAsync(d_in + o, h_in + o, std::min(piece, in_b - o), cudaMemcpyHostToDevice, s[k % ns]);
                if (with_d2h)
                    for (size_t o = 0; o < out_b; o += piece)
                        cudaMemcpyAsync(h_out + o, d_out + o, std::min(piece, out_b - o), cudaMemcpyDeviceToHost, sd);
                cudaEvent_t ej[4];
                for (int i = 1; i < ns; i++) { cudaEventCreate(&ej[i]); cudaEventRecord(ej[i], s[i]); cudaStreamWaitEvent(s[0], ej[i], 0); }
                cudaEventCreate(&ej[0]); cudaEventRecord(ej[0], sd); cudaStreamWaitEvent(s[0], ej[0], 0);
                cudaEventRecord(e1, s[0]);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("H2D on %d stream(s)%s: %.2f ms  (H2D %.1f GB/s)\n", ns, with_d2h ? " + D2H" : "", best, in_b / best / 1e6);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
