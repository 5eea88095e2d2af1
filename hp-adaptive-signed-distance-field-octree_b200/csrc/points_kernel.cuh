// points_kernel.cuh — the synthetic query workload of BASELINE config 5: points uniform in a box from
// Philox4x32-10 (Salmon et al., SC'11) with key = seed and counter = GLOBAL point index, 53-bit mantissas. A point depends
// only on (seed, index), so any number of ranks, in any chunking, evaluate the same 10^9 points; tests/philox.py is the
// numpy statement of the same generator (known-answer vectors of Random123 included).
#pragma once
#include "hp_common.h"

namespace hpsdf
{
    struct Philox4 { uint32_t x, y, z, w; };

    __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
    {
        #pragma unroll
        for (int r = 0; r < 10; ++r)
        {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
        return o;
    }

    // 53 random bits -> [0, 1)
    __device__ __forceinline__ double unitDouble(uint32_t hi, uint32_t lo)
    {
        return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
    }

    struct Box3d { double lo[3], ext[3]; };

    // xyz[3 i .. 3 i + 2] = lo + u * ext for point first + i: x, y from Philox block (index, 0), z from block (index, 1)
    __global__ void __launch_bounds__(256) uniformPointsKernel(uint64_t seed, uint64_t first, size_t n, const Box3d box, double* __restrict__ xyz)
    {
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        {
            const uint64_t idx = first + i;
            const Philox4 a = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u, k0, k1);
            const Philox4 b = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), 1u, 0u, k0, k1);
            xyz[3 * i]     = __dadd_rn(box.lo[0], __dmul_rn(unitDouble(a.x, a.y), box.ext[0]));
            xyz[3 * i + 1] = __dadd_rn(box.lo[1], __dmul_rn(unitDouble(a.z, a.w), box.ext[1]));
            xyz[3 * i + 2] = __dadd_rn(box.lo[2], __dmul_rn(unitDouble(b.x, b.y), box.ext[2]));
        }
    }

    cudaError_t launchUniformPoints(uint64_t seed, uint64_t first, size_t n, const double lo[3], const double hi[3], double* dXyz, cudaStream_t stream)
    {
        if (!n) return cudaSuccess;
        Box3d box;
        for (int a = 0; a < 3; ++a) { box.lo[a] = lo[a]; box.ext[a] = hi[a] - lo[a]; }
        const size_t want = (n + 255) / 256;
        uniformPointsKernel<<<(unsigned)(want < 148 * 32 ? want : 148 * 32), 256, 0, stream>>>(seed, first, n, box, dXyz);
        return cudaGetLastError();
    }
}
