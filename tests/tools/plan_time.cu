// Not a test: measures cuFFT plan-creation and execution time for candidate FFT lengths (2-D R2C batch 8 + C2R batch 3).
#include <cufft.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    std::vector<int> sizes;
    for (int i = 1; i < argc; i++) sizes.push_back(atoi(argv[i]));
    cudaFree(0);
    float *in; cufftComplex *spec;
    size_t maxM = 0; for (int m : sizes) if ((size_t) m > maxM) maxM = m;
    cudaMalloc(&in, maxM * maxM * 8 * 4); cudaMalloc(&spec, maxM * (maxM / 2 + 1) * 8 * 8);
    cudaMemset(in, 0, maxM * maxM * 8 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int M : sizes) {
        int n[2] = {M, M};
        cufftHandle f, b;
        double t0 = now();
        cufftResult r1 = cufftPlanMany(&f, 2, n, nullptr, 1, M * M, nullptr, 1, M * (M / 2 + 1), CUFFT_R2C, 8);
        double t1 = now();
        cufftResult r2 = cufftPlanMany(&b, 2, n, nullptr, 1, M * (M / 2 + 1), nullptr, 1, M * M, CUFFT_C2R, 3);
        double t2 = now();
        size_t ws1 = 0, ws2 = 0; cufftGetSize(f, &ws1); cufftGetSize(b, &ws2);
        for (int w = 0; w < 3; w++) { cufftExecR2C(f, in, spec); cufftExecC2R(b, spec, in); }
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int w = 0; w < 20; w++) cufftExecR2C(f, in, spec);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float msf; cudaEventElapsedTime(&msf, e0, e1);
        cudaEventRecord(e0);
        for (int w = 0; w < 20; w++) cufftExecC2R(b, spec, in);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float msb; cudaEventElapsedTime(&msb, e0, e1);
        printf("M=%5d plan R2Cx8 %8.1f ms (rc %d)  C2Rx3 %8.1f ms (rc %d)  exec R2Cx8 %7.1f us  C2Rx3 %7.1f us  ws %zu/%zu KB\n", M, t1 - t0, r1, t2 - t1, r2,
               msf / 20 * 1e3, msb / 20 * 1e3, ws1 >> 10, ws2 >> 10);
        fflush(stdout);
        cufftDestroy(f); cufftDestroy(b);
    }
    return 0;
}
