// accuracy of dlsm::fast_sqrt against IEEE sqrt on the device
#include "../../dynetlsm_b200/csrc/dlsm_device.cuh"
#include <cstdio>
__global__ void k(const double *x, double *out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = dlsm::fast_sqrt(x[i]);
}
int main() {
    const int n = 1 << 22;
    double *hx = new double[n], *ho = new double[n], *dx, *dout;
    unsigned long long s = 88172645463325252ULL;
    for (int i = 0; i < n; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        double u = (s >> 11) * (1.0 / 9007199254740992.0);
        hx[i] = exp((u - 0.5) * 80.0); }   // 1e-17 .. 1e17
    hx[0] = 0.0; hx[1] = 1e-290; hx[2] = 4.0; hx[3] = 2.0;
    cudaMalloc(&dx, n * 8); cudaMalloc(&dout, n * 8);
    cudaMemcpy(dx, hx, n * 8, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(dx, dout, n);
    cudaMemcpy(ho, dout, n * 8, cudaMemcpyDeviceToHost);
    double max_ulp = 0; int exact = 0;
    for (int i = 4; i < n; i++) {
        double ref = sqrt(hx[i]);
        double ulp = fabs(ho[i] - ref) / (nextafter(ref, INFINITY) - ref);
        if (ulp > max_ulp) max_ulp = ulp;
        exact += (ho[i] == ref);
    }
    printf("max ulp error %.2f, exactly rounded %.4f%%, sqrt(0)=%g sqrt(1e-290)=%g sqrt(4)=%.17g sqrt(2)=%.17g\n",
           max_ulp, 100.0 * exact / (n - 4), ho[0], ho[1], ho[2], ho[3]);
    return 0;
}
