// bw_probe.cu -- practical HBM ceilings on this GPU for the access mixes the fused stage kernel has:
// pure streaming read, 90/10 read/write, and 50/50 copy.  Prints GB/s (best of 10, CUDA events).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_read(const double2 *__restrict__ a, size_t n, double *out)
{
    double s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double2 v = a[i];
        s += v.x + v.y;
    }
    if (s == 123.456) out[0] = s;
}
// reads nr arrays, writes 1 (nr = 9 -> 90/10 mix; nr = 1 -> copy)
template <int NR>
__global__ void k_mix(const double2 *__restrict__ a, size_t n, double2 *__restrict__ w)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double2 s = make_double2(0, 0);
#pragma unroll
        for (int r = 0; r < NR; ++r) { double2 v = a[(size_t)r * n + i]; s.x += v.x; s.y += v.y; }
        w[i] = s;
    }
}
template <class F> float best(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float b = 1e30f;
    for (int it = 0; it < 12; ++it) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2 && ms < b) b = ms;
    }
    return b;
}
int main()
{
    const size_t n = (size_t)16 << 20;  // double2 elements per array: 256 MB
    double2 *a, *w; double *o;
    CK(cudaMalloc(&a, 10 * n * sizeof(double2))); CK(cudaMalloc(&w, n * sizeof(double2))); CK(cudaMalloc(&o, 8));
    CK(cudaMemset(a, 0, 10 * n * sizeof(double2)));
    for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        float t0 = best([&] { k_read<<<blocks, 256>>>(a, 10 * n, o); });
        float t1 = best([&] { k_mix<9><<<blocks, 256>>>(a, n, w); });
        float t2 = best([&] { k_mix<1><<<blocks, 256>>>(a, n, w); });
        printf("blocks %5d  read %.0f GB/s   90/10 mix %.0f GB/s   copy %.0f GB/s\n", blocks, 10.0 * n * 16 / t0 / 1e6,
               10.0 * n * 16 / t1 / 1e6, 2.0 * n * 16 / t2 / 1e6);
    }
    CK(cudaGetLastError());
    return 0;
}
