// common.cuh -- shared device/host definitions for libbldfm_b200 (sm_100a).
//
// This translation unit family is compiled with --fmad=false: a plain `a*b+c` is NEVER contracted.
// Wherever a fused multiply-add is wanted it is written as an explicit fma().  That keeps the
// rounding of every operation under our control, which the linear-shooting combine needs
// (round-off is amplified by e^{2*kappa}).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <unordered_map>

namespace bldfm {

// cudaFuncAttributeMaxDynamicSharedMemorySize once per (kernel, device, size) instead of in front of every launch
// (a microsecond of driver time each, three times per solve, on the host path in front of the kernels).
template <class K>
inline cudaError_t set_max_dyn_smem(K kernel, int bytes)
{
    static std::mutex mu;
    static std::unordered_map<uint64_t, int> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t key = (uint64_t)reinterpret_cast<uintptr_t>(reinterpret_cast<const void*>(kernel)) * 64u + (uint64_t)(dev & 63);
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = done.find(key);
        if (it != done.end() && it->second == bytes) return cudaSuccess;
    }
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(mu);
        done[key] = bytes;
    }
    return e;
}

// Per-level scalars of one march step i (solver.py:357-364), precomputed on the host in the
// reference's operation order and staged in shared memory by the march kernels.
// 16 doubles = 128 B so that a level is four 32 B sectors and every field is LDS.128-friendly.
struct __align__(16) LevelCoef {
    double Kx, Ky;     // Kx[i], Ky[i]
    double u, v;       // u[i], v[i]
    double s, h;       // 0.5*Kzinv ; dz[i]
    double h2, h3;     // dz*dz ; (dz*dz)*dz
    double s6, s61;    // (1/6)*(Kzinv*Kzinv) ; (1/6)*Kzinv
    double c0, w;      // (-Kzinv)*dz ; trapezoid weight 0.5/Kz[i] + 0.5/Kz[i+1]   (solver.py:248)
    double sh2, s6h3;  // FMA mode only: s*h2 ; s6*h3
    double s61h3, pad; // FMA mode only: s61*h3
};
static_assert(sizeof(LevelCoef) == 128, "LevelCoef must be 128 bytes");

// One march group: a unique (z, profiles, srf_bg_conc).  Towers that share it differ only by the
// phase shift.
struct GroupDesc {
    int32_t S;            // number of march steps = nz - 1
    int32_t tow_begin;    // first entry in the tower list
    int32_t tow_count;    // number of towers / output slots fed by this march
    int32_t pad;
    double kz_top;        // Kz[nz-1]                          solver.py:228
    double kinv_top;      // 1.0/Kz[nz-1]                      solver.py:164
    double kxk, kyk;      // Kx[nz-1]*Kzinv, Ky[nz-1]*Kzinv    solver.py:165-166
    double c1, c2;        // u[nz-1]*Kzinv, v[nz-1]*Kzinv      solver.py:172-173
    double p000;          // srf_bg_conc                       solver.py:83
    double h_analytic;    // z[level]-z[0] (analytic branch)   solver.py:197
};

struct TowerDesc {
    double sx, sy;        // phase = lx*sx + ly*sy             solver.py:255,260
    int32_t shift;        // 0: no phase shift applied         solver.py:254-262
    int32_t slot;         // output problem index
};

// ---- mbarrier / bulk-copy primitives (PTX; sm_90+) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// global -> shared bulk copy, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace bldfm
