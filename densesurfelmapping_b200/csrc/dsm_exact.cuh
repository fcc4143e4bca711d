// Exactness helpers shared by the kernel files: the float-vs-double-literal compare constants, warp helpers and the
// expression-by-expression restatement of calculate_cost (fusion_functions.cpp:364-387).
#pragma once
#include "dsm_device.cuh"
#include <climits>
#include <cuda/barrier>
#include <cuda/ptx>

#define HUBER_RANGE 0.4       // fusion_functions.h:13
#define MAX_ANGLE_COS 0.1     // fusion_functions.h:11
#define BASELINE 0.5          // fusion_functions.h:14
#define DISPARITY_ERROR 4.0   // fusion_functions.h:15
#define MIN_TOLERATE_DIFF 0.1 // fusion_functions.h:16

#define FULL 0xffffffffu

// ---- programmatic dependent launch (PDL) ----
// The per-frame schedule is a chain of kernels on one stream, each consuming its predecessor's output.  Every kernel
// of the chain starts with pdl_enter(): griddepcontrol.wait returns once the preceding grid has completed and its
// writes are visible (a no-op for a launch without the attribute); griddepcontrol.launch_dependents then lets the NEXT
// grid of the chain be set up -- its CTAs become resident as SM resources free up in this grid's last wave and wait at
// their own griddepcontrol.wait -- so the launch latency and the CTA ramp of the next kernel overlap this kernel's tail.
// Safe by construction: no kernel touches global memory before the wait, and a kernel is only launched with the
// attribute (pdl_launch) if it begins with pdl_enter().
#ifndef DSM_PDL
#define DSM_PDL 1
#endif
__device__ __forceinline__ void pdl_enter()
{
#if DSM_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
template <typename... KArgs, typename... Args>
static inline void pdl_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
#if DSM_PDL
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, args...);
#else
    kern<<<grid, block, smem, s>>>(args...);
#endif
}

// Comparisons of a float against a double literal (the reference promotes the float): for a
// literal c that is not a float, with c_lo/c_hi the neighbouring floats,
//   (double)x <  c  <=>  x <  c_hi        (double)x >  c  <=>  x >  c_lo
//   (double)x >= c  <=>  x >= c_hi        (double)x <= c  <=>  x <= c_lo
// and for -c by symmetry (x > -c <=> x > -c_hi).  Used in the hot loops to keep the FP64 pipe for
// the arithmetic that really needs it; tests/test_abi.py re-derives every constant.
#define F_0p4_HI __uint_as_float(0x3ecccccdu)
#define F_0p4_LO __uint_as_float(0x3eccccccu)
#define F_0p1_HI __uint_as_float(0x3dcccccdu)
#define F_0p1_LO __uint_as_float(0x3dccccccu)
#define F_0p01_HI __uint_as_float(0x3c23d70bu)
#define F_0p01_LO __uint_as_float(0x3c23d70au)
#define F_0p05_LO __uint_as_float(0x3d4cccccu)
#define F_0p2_HI __uint_as_float(0x3e4ccccdu)
#define F_0p8_HI __uint_as_float(0x3f4ccccdu)

// -------------------------------------------------------------------------------------------
// small helpers
// -------------------------------------------------------------------------------------------
using dsm_barrier = cuda::barrier<cuda::thread_scope_block>; // mbarrier for the TMA (cp.async.bulk) stagings
// Wait for phase 0 of a tile barrier.  Experimental kernels only: a byte-count mistake would otherwise spin forever and
// take the GPU with it, so the wait is bounded (each try_wait already blocks for a hardware time slice) and traps.
__device__ __forceinline__ void tile_wait(dsm_barrier &bar)
{
    for (unsigned spin = 0; !cuda::ptx::mbarrier_try_wait_parity(cuda::device::barrier_native_handle(bar), 0); spin++)
        if (spin > (1u << 22)) __trap();
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v)
{ // st.shared with a precomputed 32-bit shared-window address (keeps the address math out of the hot loops)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v));
}
__device__ __forceinline__ float warp_sum_f(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_excl_scan(int v, int lane, int &total)
{
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int n = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += n;
    }
    total = __shfl_sync(FULL, incl, 31);
    return incl - v;
}

// (hi, lo) split of 1.0 / (double)mean_depth for the fp32 cost filter.  mean_depth <= 0 (or NaN): the candidate has no
// depth term (:378), the pair is unused.  mean_depth < 2^-10: hi = 1e18 (its square is still finite, so a zero depth weight cannot produce a NaN) -- the depth term of such a seed exceeds 1e6 for
// every valid pixel depth (> 0.01 m, i.e. inverse < 100), so it can never be the minimum (:408, :427) and the filter
// may ignore it; if NO candidate stays below 1e6 the pixel goes to the exact path anyway.
__device__ __forceinline__ float2 split_inverse(float md)
{
    if (!(md > 0.f)) return make_float2(0.f, 0.f);
    if (md < 0.0009765625f) return make_float2(1e18f, 0.f);
    const double inv = 1.0 / (double)md;
    const float hi = (float)inv;
    return make_float2(hi, (float)(inv - (double)hi));
}

struct SeedC
{
    float x, y, I, md;
    double inv;
};

// Branch-free: both costs and the has-depth predicate are always computed; the caller selects.
// (When mean_depth <= 0 the hoisted 1/mean_depth is inf/NaN; the result is discarded by the select.)
__device__ __forceinline__ bool calc_cost(const SeedC &sd, float pix_i, float pix_inv, double pix_inv_d, float fx, float fy,
                                          float &nodepth, float &withdepth)
{
    const float ax = sd.x - fx, ay = sd.y - fy;
    const float dist = ax * ax + ay * ay;
    float n = dist * 0.0625f; // / (SP_SIZE/2)^2, exact power of two (:374)
    const float idf = sd.I - pix_i;
    // (double)(idf*idf) / 100.0 (:376), correctly rounded without the division subroutine:
    // q0 = RN(a*y), r = a - 100*q0 (exact in one FMA), q = RN(q0 + r*y) with y = RN(1/100) is the
    // correctly rounded quotient (Markstein); checked against x/100.0 on 3.7e8 inputs.
    const double a = (double)(idf * idf);
    const double q0 = a * 0.01;
    const double q = __fma_rn(__fma_rn(-q0, 100.0, a), 0.01, q0);
    const double nd = (double)n + q;
    n = (float)nd;
    nodepth = n;
    const bool has = sd.md > 0 && pix_inv > 0; // (:378)
    const float idd = (float)(sd.inv - pix_inv_d);                  // (:380)
    const float wd = (float)((double)n + (double)(idd * idd) * 400.0); // (:381)
    withdepth = has ? wd : n;
    return has;
}

