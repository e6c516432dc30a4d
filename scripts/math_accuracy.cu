// Accuracy check of the replay-tier math helpers (frcp/fdiv/fsqrt/frsqrt/flog/fsincospi in csrc/ptl_physics.cuh) against the
// correctly-rounded / libdevice results, in ulps, over 2^24 random arguments per function.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I particulator.jl_b200/csrc -I include scripts/math_accuracy.cu -o /tmp/math_accuracy
#include <cstdio>
#include <cstdint>
#include <cmath>
#include "ptl_physics.cuh"

using ptl::philox4x32_10; using ptl::bits_to_u01;

__device__ double ulps(double a, double ref) {
    if (a == ref) return 0;
    if (!(fabs(ref) > 0) || isinf(ref) || isnan(ref)) return isnan(a) == isnan(ref) && isinf(a) == isinf(ref) ? 0 : 1e9;
    int e;
    frexp(ref, &e);
    return fabs(a - ref) / ldexp(1.0, e - 53);
}

__device__ double rnd_arg(uint32_t i, int mode) {
    uint32_t o[4];
    philox4x32_10(i, (uint32_t)mode, 0x1234u, 0x5678u, 0xdeadbeefu, 0xcafef00du, o);
    double u = bits_to_u01(o[0], o[1]), v = bits_to_u01(o[2], o[3]);
    switch (mode) {
    case 0: return u;                                   // uniform deviates: -log(u)
    case 1: return exp((v - 0.5) * 200.0) * (1 + u);    // 1e-43 .. 1e43 (momenta^2, energies, rates)
    case 2: return 1.0 + (u - 0.5) * 1e-3 * v;          // around 1 (cancellation in log)
    default: return ldexp(1 + u, (int)(v * 40) - 20);
    }
}

__global__ void k_check(int mode, double* maxerr) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    double x = rnd_arg(i, mode);
    double e[7];
    e[0] = ulps(ptl::flog(x), log(x));
    e[1] = ulps(ptl::frsqrt(x), rsqrt(x));
    e[2] = ulps(ptl::fsqrt(x), sqrt(x));
    e[3] = ulps(ptl::frcp(x), 1.0 / x);
    e[4] = ulps(ptl::fdiv(0.7310585786300049, x), 0.7310585786300049 / x);
    {
        double sr, cr, sf, cf, xx = mode == 0 ? 2 * x : (mode == 2 ? x : 2 * rnd_arg(i, 0) * (i & 1 ? 1.0 : 1e-3));
        sincospi(xx, &sr, &cr);
        ptl::fsincospi(xx, sf, cf);
        e[5] = ulps(sf, sr); e[6] = ulps(cf, cr);
    }
    for (int k = 0; k < 7; k++) {
        double m = e[k];
        for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
        if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&maxerr[k], (unsigned long long)__double_as_longlong(m));
    }
}

int main() {
    double* d;
    cudaMalloc(&d, 7 * sizeof(double));
    const char* names[7] = {"flog", "frsqrt", "fsqrt", "frcp", "fdiv", "sinpi", "cospi"};
    const char* modes[4] = {"u in (0,1)", "1e-43..1e43", "1 +- 5e-4", "2^-20..2^20"};
    int bad = 0;
    for (int mode = 0; mode < 4; mode++) {
        cudaMemset(d, 0, 7 * sizeof(double));
        k_check<<<(1 << 24) / 256, 256>>>(mode, d);
        double h[7];
        if (cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 2; }
        printf("%-12s", modes[mode]);
        for (int k = 0; k < 7; k++) { printf("  %s %.3f ulp", names[k], h[k]); if (!(h[k] <= 2.0)) bad = 1; }
        printf("\n");
    }
    printf(bad ? "FAIL (> 2 ulp)\n" : "OK (all <= 2 ulp)\n");
    return bad;
}
