#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k(const float2* a, const float2* b, const float2* c, int n, unsigned long long* bad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 x = a[i], y = b[i], z = c[i];
    float2 r = __ffma2_rn(x, y, z);
    float sx = __fmaf_rn(x.x, y.x, z.x), sy = __fmaf_rn(x.y, y.y, z.y);
    if (__float_as_uint(r.x) != __float_as_uint(sx) || __float_as_uint(r.y) != __float_as_uint(sy)) atomicAdd(&bad[0], 1ull);
    float2 m = __fmul2_rn(x, y);
    if (__float_as_uint(m.x) != __float_as_uint(__fmul_rn(x.x, y.x)) || __float_as_uint(m.y) != __float_as_uint(__fmul_rn(x.y, y.y))) atomicAdd(&bad[1], 1ull);
    float2 s = __fadd2_rn(x, z);
    if (__float_as_uint(s.x) != __float_as_uint(__fadd_rn(x.x, z.x)) || __float_as_uint(s.y) != __float_as_uint(__fadd_rn(x.y, z.y))) atomicAdd(&bad[2], 1ull);
}
int main()
{
    const int n = 1 << 24;
    float2 *a, *b, *c; unsigned long long* bad;
    cudaMallocManaged(&a, n * 8); cudaMallocManaged(&b, n * 8); cudaMallocManaged(&c, n * 8); cudaMallocManaged(&bad, 24);
    for (int mode = 0; mode < 4; ++mode)
    {
        srand(1234 + mode);
        for (int i = 0; i < n; ++i)
        {
            auto rnd = [&](int m) -> float {
                if (m == 0) return (float)(rand() % 512 - 256);                                   // small integers
                if (m == 1) return (float)(rand() % 65536) / 64.0f - 300.0f;                      // multiples of 1/64
                if (m == 2) { uint32_t u = ((uint32_t)rand() << 16) ^ (uint32_t)rand(); float f; u &= 0xBFFFFFFFu; memcpy(&f, &u, 4); return f; }   // any bits (incl. denormals, no huge)
                return ((float)rand() / RAND_MAX - 0.5f) * 1e-19f;                                // tiny: products denormal
            };
            a[i] = make_float2(rnd(mode), rnd(mode)); b[i] = make_float2(rnd(mode), rnd(mode)); c[i] = make_float2(rnd(mode), rnd(mode));
        }
        bad[0] = bad[1] = bad[2] = 0;
        k<<<n / 256, 256>>>(a, b, c, n, bad);
        cudaDeviceSynchronize();
        printf("mode %d: fma2 mismatches %llu  mul2 %llu  add2 %llu of %d\n", mode, bad[0], bad[1], bad[2], n);
    }
    return 0;
}
