#include <cstdio>
#include <cstring>
#include "../polars-strsim_b200/csrc/row_ascii_reg.cuh"
using namespace strsim;
__global__ void k(const uint32_t* in, int na, int nb, double* out, int* ints) {
    uint32_t a[REG_WORDS], b[REG_WORDS];
    for (int i = 0; i < 8; i++) { a[i] = in[i]; b[i] = in[8 + i]; }
    PairInts pi;
    out[0] = row_ascii_reg<0, 5>(a, b, na, nb, pi);
    ints[0] = pi.x0;
    out[1] = row_ascii_reg<1, 5>(a, b, na, nb, pi);
    ints[1] = pi.x0; ints[2] = pi.x1;
    out[2] = row_ascii_reg<3, 7>(a, b, na, nb, pi);
    ints[3] = pi.x0;
}
int main() {
    uint32_t h[16] = {0}; memcpy(h, "phillips", 8); memcpy(h + 8, "philips", 7);
    uint32_t* d; cudaMalloc(&d, 64); cudaMemcpy(d, h, 64, cudaMemcpyHostToDevice);
    double* o; cudaMalloc(&o, 64); int* it; cudaMalloc(&it, 64);
    k<<<1, 1>>>(d, 8, 7, o, it);
    double ho[3]; int hi[4]; cudaMemcpy(ho, o, 24, cudaMemcpyDeviceToHost); cudaMemcpy(hi, it, 16, cudaMemcpyDeviceToHost);
    printf("dev lev %g d=%d jaro %g m=%d t=%d jac %g inter=%d\n", ho[0], hi[0], ho[1], hi[1], hi[2], ho[2], hi[3]);
    uint32_t a[8], b[8]; memcpy(a, h, 32); memcpy(b, h + 8, 32); PairInts pi;
    printf("host lev %g\n", row_ascii_reg<0, 5>(a, b, 8, 7, pi));
    return 0;
}
