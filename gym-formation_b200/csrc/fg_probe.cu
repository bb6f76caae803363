// fg_probe.cu -- FP32 pipe probes (sm_100a).  The north star grades the large-N step+reward kernel
// against the FP32 peak, which MEASURED_PEAKS.json does not hold; bench.py measures it live with
// these kernels (CUDA events around the launch) and reports the fraction against the measured number.
//   variant 0: scalar FFMA, 3 register operands      (2 flop / lane / instruction)
//   variant 1: packed FFMA2 (fma.rn.f32x2, Blackwell)  (4 flop / lane / instruction)
//   variant 2: FMNMX (alu pipe)                        (1 op  / lane / instruction)
//   variant 3: the pair-loop mix: FADD2,FADD2,FMUL2,FFMA2 + FMNMX3 per two pairs
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/formation_gym_b200.h"

namespace {

__device__ __forceinline__ uint64_t pk(float a, float b) {
    return ((uint64_t)__float_as_uint(b) << 32) | (uint64_t)__float_as_uint(a);
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

template <int V>
__global__ void __launch_bounds__(256) k_probe(int iters, float seed, float* out) {
    const float s = seed + (float)threadIdx.x * 1e-6f;
    float r = 0.f;
    if (V == 0) {
        float a0 = s, a1 = s + 1, a2 = s + 2, a3 = s + 3, a4 = s + 4, a5 = s + 5, a6 = s + 6, a7 = s + 7;
        const float b = 0.999f + s * 1e-9f, c = 1e-3f + s;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
                a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
            }
        }
        r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    } else if (V == 1) {
        uint64_t a0 = pk(s, s + 1), a1 = pk(s + 2, s + 3), a2 = pk(s + 4, s + 5), a3 = pk(s + 6, s + 7);
        uint64_t a4 = pk(s + 8, s + 9), a5 = pk(s + 10, s + 11), a6 = pk(s + 12, s + 13), a7 = pk(s + 14, s + 15);
        const uint64_t b = pk(0.999f + s * 1e-9f, 0.998f), c = pk(1e-3f + s, 2e-3f);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a0 = ffma2(a0, b, c); a1 = ffma2(a1, b, c); a2 = ffma2(a2, b, c); a3 = ffma2(a3, b, c);
                a4 = ffma2(a4, b, c); a5 = ffma2(a5, b, c); a6 = ffma2(a6, b, c); a7 = ffma2(a7, b, c);
            }
        }
        uint64_t x = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
        r = __uint_as_float((uint32_t)x) + __uint_as_float((uint32_t)(x >> 32));
    } else if (V == 2) {
        float a0 = s, a1 = s + 1, a2 = s + 2, a3 = s + 3, a4 = s + 4, a5 = s + 5, a6 = s + 6, a7 = s + 7;
        float b = s * 3.f;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a0 = fminf(a0, b); a1 = fmaxf(a1, a0); a2 = fminf(a2, a1); a3 = fmaxf(a3, a2);
                a4 = fminf(a4, a3); a5 = fmaxf(a5, a4); a6 = fminf(a6, a5); a7 = fmaxf(a7, a6);
                b = __uint_as_float(__float_as_uint(b) ^ (uint32_t)i);       // keep the chain data-dependent
            }
        }
        r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    } else {
        // two partner pairs per iteration of the unrolled body: dx,dy (FADD2 x2), dx*dx (FMUL2),
        // dy*dy+ (FFMA2), running minimum (3-input FMNMX)
        uint64_t xj = pk(s, s + 0.5f), yj = pk(s + 0.25f, s + 0.75f);
        const uint64_t nxi = pk(-0.3f - s, -0.3f - s), nyi = pk(0.2f + s, 0.2f + s);
        const uint64_t step = pk(1e-3f, 2e-3f);
        float m0 = 1e30f, m1 = 1e30f;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                uint64_t dx = fadd2(xj, nxi), dy = fadd2(yj, nyi);
                uint64_t d2 = ffma2(dy, dy, fmul2(dx, dx));
                float lo = __uint_as_float((uint32_t)d2), hi = __uint_as_float((uint32_t)(d2 >> 32));
                if (u & 1) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m1) : "f"(lo), "f"(hi));
                else       asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m0) : "f"(lo), "f"(hi));
                xj = fadd2(xj, step); yj = fadd2(yj, step);
            }
        }
        r = m0 + m1;
    }
    if (r == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = r;     // never true; defeats DCE
}

// Write-only HBM streams (what an observation writer can reach at best): plain 16-byte stores and TMA
// bulk stores of `chunk` bytes from shared memory, grid-stride over `bytes`.
__global__ void __launch_bounds__(256) k_wr_stg(float4* dst, size_t n16) {
    const float4 v = make_float4(1.f, 2.f, 3.f, (float)threadIdx.x);
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n16; q += (size_t)gridDim.x * blockDim.x)
        __stcs(dst + q, v);
}

__global__ void __launch_bounds__(256) k_wr_bulk(unsigned char* dst, size_t bytes, unsigned chunk, int evict_first,
                                                 int blocked) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (unsigned q = threadIdx.x * 16; q < chunk; q += blockDim.x * 16)
        *reinterpret_cast<float4*>(sm + q) = make_float4(1.f, 2.f, 3.f, (float)q);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        const size_t nchunks = bytes / chunk;
        // cyclic: CTA b writes chunks b, b + grid, ... (all CTAs advance through one moving window of grid x chunk bytes);
        // blocked: CTA b writes its own contiguous range of chunks (grid independent sequential streams)
        const size_t per = (nchunks + gridDim.x - 1) / gridDim.x;
        const size_t c0 = blocked ? blockIdx.x * per : blockIdx.x, c1 = blocked ? min(nchunks, c0 + per) : nchunks;
        const size_t cs = blocked ? 1 : gridDim.x;
        for (size_t c = c0; c < c1; c += cs) {
            if (evict_first)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                             :: "l"(dst + c * chunk), "r"((uint32_t)__cvta_generic_to_shared(sm)), "r"(chunk), "l"(pol) : "memory");
            else
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(dst + c * chunk), "r"((uint32_t)__cvta_generic_to_shared(sm)), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// variant 3: observation rows written with PLAIN 8-byte streaming stores from registers, the way a warp-per-env
// kernel would write hd rows without a shared-memory image: `chunk` = N agents; a warp owns one env (N rows of 3N
// float2 items, 8-byte aligned only) and writes, per row, the N dynamic items with one store instruction and the 2N
// static items with ceil(2N / 32) more.
__global__ void __launch_bounds__(128) k_wr_rows(float2* dst, size_t n_env, unsigned N) {
    const unsigned lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((size_t)gridDim.x * blockDim.x) >> 5;
    const float2 v = make_float2(1.f + (float)lane, 2.f);
    for (size_t e = warp; e < n_env; e += nwarp) {
        float2* out = dst + e * (size_t)N * 3 * N;
        for (unsigned r = 0; r < N; ++r, out += 3 * N) {
            for (unsigned k = lane; k < N; k += 32) __stcs(out + k, v);
            for (unsigned k = N + lane; k < 3 * N; k += 32) __stcs(out + k, v);
        }
    }
}

}  // namespace

extern "C" int fg_write_probe(int variant, void* dst, unsigned long long bytes, unsigned chunk, int ctas, void* stream) {
    if (!dst || bytes < 16 || ctas < 1) return FG_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (variant == 0) {
        k_wr_stg<<<ctas, 256, 0, st>>>((float4*)dst, (size_t)(bytes / 16));
    } else if (variant == 3) {
        if (chunk < 1 || chunk > 256) return FG_ERR_ARG;
        k_wr_rows<<<ctas, 128, 0, st>>>((float2*)dst, (size_t)(bytes / ((size_t)24 * chunk * chunk)), chunk);
    } else {
        if (chunk < 16 || (chunk & 15) || chunk > 200 * 1024) return FG_ERR_ARG;
        if (chunk > 48 * 1024 &&
            cudaFuncSetAttribute(k_wr_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chunk) != cudaSuccess)
            return FG_ERR_CUDA;
        k_wr_bulk<<<ctas, 256, chunk, st>>>((unsigned char*)dst, (size_t)bytes, chunk, variant >= 2, variant == 4);
    }
    return cudaGetLastError() == cudaSuccess ? FG_OK : FG_ERR_CUDA;
}

extern "C" int fg_fp32_probe(int variant, int iters, int ctas, float* scratch, void* stream) {
    if (variant < 0 || variant > 3 || iters < 1 || ctas < 1 || !scratch) return FG_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (variant) {
        case 0: k_probe<0><<<ctas, 256, 0, st>>>(iters, 1.0f, scratch); break;
        case 1: k_probe<1><<<ctas, 256, 0, st>>>(iters, 1.0f, scratch); break;
        case 2: k_probe<2><<<ctas, 256, 0, st>>>(iters, 1.0f, scratch); break;
        default: k_probe<3><<<ctas, 256, 0, st>>>(iters, 1.0f, scratch); break;
    }
    return cudaGetLastError() == cudaSuccess ? FG_OK : FG_ERR_CUDA;
}
