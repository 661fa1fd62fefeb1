// selftest.cu — device-side self checks reachable through the C-ABI (tests/test_gpu_parity.py).
//
// rt_selftest_div3: div3() (device_common.cuh: three divisions sharing one refined reciprocal) against the plain
// IEEE `/` operator on pseudo-random operand quadruples, bit for bit.  Three operand families per index: raw random
// bit patterns (zeros, denormals, infinities, NaNs included), magnitudes of scene scale (the fast path), and
// exponents straddling the 2^-60 / 2^60 guard.
#include <cstdint>
#include <string>

#include <cuda_runtime.h>

#include "device_common.cuh"

namespace rtb {

namespace {

__device__ __forceinline__ uint32_t mix32(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return (uint32_t) x;
}

__device__ __forceinline__ float operand(uint64_t i, int which, int family, uint32_t seed) {
    const uint32_t r = mix32(i * 4 + which + ((uint64_t) seed << 40));
    if (family == 0) return __uint_as_float(r);
    const uint32_t sign = r & 0x80000000u, mant = r & 0x007fffffu;
    const uint32_t e = family == 1 ? 117u + ((r >> 23) & 15u)              // 2^-10 .. 2^5
                                   : (((r >> 23) & 1u) ? 187u : 67u) - 4u + ((r >> 24) & 7u);  // around 2^60 / 2^-60
    return __uint_as_float(sign | (e << 23) | mant);
}

__device__ __forceinline__ bool same(float x, float y) {
    return __float_as_uint(x) == __float_as_uint(y) || (x != x && y != y);
}

__global__ void div3_check_kernel(uint64_t n, uint32_t seed, unsigned long long *out) {
    unsigned long long bad = 0, fast = 0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const int family = (int) (i % 3);
        const float a0 = operand(i, 0, family, seed), a1 = operand(i, 1, family, seed), a2 = operand(i, 2, family, seed);
        const float b = operand(i, 3, family, seed);
        float q0, q1, q2;
        div3(a0, a1, a2, b, q0, q1, q2);
        if (!same(q0, a0 / b) || !same(q1, a1 / b) || !same(q2, a2 / b)) bad++;
        if (div_in_range(a0) && div_in_range(a1) && div_in_range(a2) && div_in_range(b)) {
            fast++;
            // the two-plus-one form of the triangle test
            const float r = rcp_refined(b);
            if (!same(div_quot(a0, b, r), a0 / b)) bad++;
        }
    }
    bad = __reduce_add_sync(0xffffffffu, (unsigned) bad);
    fast = __reduce_add_sync(0xffffffffu, (unsigned) fast);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], bad);
        atomicAdd(&out[1], fast);
    }
}

}  // namespace

}  // namespace rtb

extern "C" int rt_selftest_div3(uint64_t n, uint32_t seed, uint64_t *mismatches, uint64_t *fast_path) {
    unsigned long long *d = nullptr, h[2] = {0, 0};
    if (cudaMalloc(&d, sizeof h) != cudaSuccess) return RT_ERR_CUDA;
    cudaMemset(d, 0, sizeof h);
    rtb::div3_check_kernel<<<148 * 8, 256>>>(n, seed, d);
    const cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return RT_ERR_CUDA;
    if (mismatches) *mismatches = h[0];
    if (fast_path) *fast_path = h[1];
    return RT_OK;
}
