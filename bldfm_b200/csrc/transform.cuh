// transform.cuh -- data-movement kernels around the horizontal transforms (K1, K3, K9, K11).
//
// Replaces np.pad (solver.py:116), fftshift/slice/ifftshift (:139-145), fftshift/pad/ifftshift
// (:265-278) and the crop + .real (:282-290) of the reference, each of which is a separate full
// array pass there.  Here truncation is an index map folded into the march kernel's loads, and
// un-truncation / crop are single gather/scatter kernels around the library transform
// (BLDFM_FFT_LIBRARY path) or folded into the pruned in-house transform (fft.cuh).
#pragma once

#include "common.cuh"

namespace bldfm {

// signed-frequency wrap: truncated index i (fftfreq order, length nl) -> index in a length-nf axis
__device__ __forceinline__ int wrap_freq(int i, int nl, int nf)
{
    return i < (nl + 1) / 2 ? i : i - nl + nf;
}

// K1: embed q0[ny][nx] (real) into a zero halo as complex [nye][nxe]                solver.py:116
__global__ void __launch_bounds__(256)
k_pad_source(const double* __restrict__ q0, double2* __restrict__ dst, int nx, int ny, int px,
             int py, int nxe, int nye)
{
    const int64_t n = (int64_t)nxe * nye;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / nxe), x = (int)(i - (int64_t)y * nxe);
        const int sy = y - py, sx = x - px;
        double v = 0.0;
        if (sy >= 0 && sy < ny && sx >= 0 && sx < nx) v = q0[(size_t)sy * nx + sx];
        dst[i] = make_double2(v, 0.0);
    }
}

// K9: scatter compact spectra [nfields][nly][nlx] into [nfields][nfy][nfx] at the wrapped
// positions.  Everything else in `dst` is zero and stays zero between solves (static halo).
template <typename C>
__global__ void __launch_bounds__(256)
k_scatter_spectrum(const C* __restrict__ src, C* __restrict__ dst, int nlx, int nly, int nfx,
                   int nfy, int nfields)
{
    const int64_t per = (int64_t)nlx * nly;
    const int64_t n = per * nfields;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(i / per);
        const int64_t r = i - (int64_t)f * per;
        const int ky = (int)(r / nlx), kx = (int)(r - (int64_t)ky * nlx);
        const int wy = wrap_freq(ky, nly, nfy), wx = wrap_freq(kx, nlx, nfx);
        dst[((size_t)f * nfy + wy) * nfx + wx] = src[i];
    }
}

// K11: crop [py:py+ny, px:px+nx] and keep the real part                          solver.py:282-290
template <typename C, typename R>
__global__ void __launch_bounds__(256)
k_crop_real(const C* __restrict__ src, R* __restrict__ dst, int nx, int ny, int px, int py,
            int nfx, int nfy, int nfields)
{
    const int64_t per = (int64_t)nx * ny;
    const int64_t n = per * nfields;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(i / per);
        const int64_t r = i - (int64_t)f * per;
        const int y = (int)(r / nx), x = (int)(r - (int64_t)y * nx);
        dst[i] = src[((size_t)f * nfy + (py + y)) * nfx + (px + x)].x;
    }
}

// f-4: footprint-weighted sums sum_{y,x} field[f][y][x] * w[y][x]  (point_measurement, utils.py:80-92)
// stage 1: kReduceBlocks partial sums per field (fixed assignment -> deterministic order)
constexpr int kReduceBlocks = 64;

template <typename R>
__global__ void __launch_bounds__(256)
k_weighted_partial(const R* __restrict__ fields, const double* __restrict__ w, double* __restrict__ partial,
                   int64_t per_field)
{
    const int f = blockIdx.y;
    const R* src = fields + (size_t)f * per_field;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_field;
         i += (int64_t)gridDim.x * blockDim.x)
        acc = fma((double)src[i], w[i], acc);
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)f * gridDim.x + blockIdx.x] = red[0];
}

// stage 2: one thread per field adds its partial sums in index order
__global__ void __launch_bounds__(128)
k_weighted_final(const double* __restrict__ partial, double* __restrict__ out, int nfields, int nblocks)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfields) return;
    double acc = 0.0;
    for (int b = 0; b < nblocks; ++b) acc += partial[(size_t)f * nblocks + b];
    out[f] = acc;
}

// f-4: time aggregation (examples/timeseries_example.py:46, np.mean over the timesteps of one tower):
// acc[slot_of[b]] += field[b] for the problems b of one batch, in problem order (deterministic; with the
// timesteps batched in order this is the summation order of np.add.reduce along axis 0).
// fields [nprob][per] (per = nlv*ny*nx), acc [nslots][per] float64, slot_of[b] < 0: problem not accumulated.
template <typename R>
__global__ void __launch_bounds__(256)
k_accumulate(const R* __restrict__ fields, double* __restrict__ acc, const int32_t* __restrict__ slot_of,
             int nprob, int64_t per)
{
    const int slot = blockIdx.y;
    double* dst = acc + (size_t)slot * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (int64_t)gridDim.x * blockDim.x) {
        double a = dst[i];
        for (int b = 0; b < nprob; ++b)
            if (slot_of[b] == slot) a += (double)fields[(size_t)b * per + i];
        dst[i] = a;
    }
}

// opt-in float32 delivery (BLDFM_DELIVER_F32): round the float64 fields to float32 on the device so that half
// the bytes cross PCIe; two fields per launch
__global__ void __launch_bounds__(256)
k_downcast2(const double* __restrict__ a, const double* __restrict__ b, float* __restrict__ oa, float* __restrict__ ob,
            int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        oa[i] = (float)a[i];
        ob[i] = (float)b[i];
    }
}

// The per-solve parameter block of a single solve (13.8 KB at config 2) fetched from the page-locked staging buffer
// by the device itself instead of a copy-engine H2D copy: the march behind it is launched with programmatic
// dependent launch, so its CTAs are resident and past their table-independent prologue set-up when the block
// lands (cudaGridDependencySynchronize in k_march) -- a copy followed by a kernel costs ~8 us on the stream,
// this pair ~3.
__global__ void __launch_bounds__(256)
k_fetch_params(const uint4* __restrict__ src_mapped_host, uint4* __restrict__ dst, int n16)
{
    cudaTriggerProgrammaticLaunchCompletion();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x)
        dst[i] = src_mapped_host[i];
}

}  // namespace bldfm
