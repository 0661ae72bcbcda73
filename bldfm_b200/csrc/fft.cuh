// fft.cuh -- pruned in-house horizontal back-transform (K9 + K10 + K11 fused) for sm_100a.
//
// The reference un-truncates the retained modes into the full padded spectrum, runs two complete
// complex 2-D FFTs of size nfy x nfx per output level and then crops the real part
// (src/bldfm/solver.py:265-290).  Only nly x nlx inputs are non-zero and only ny x nx outputs are
// kept, so the work is done here as two batched 1-D passes that never materialise a padded array:
//
//   pass X  for each retained ky row : A[ky][x']  = sum_kx S[ky][kx] w_x^{fx (px + x')},  x' in [0,nx)
//   pass Y  for each kept column x'  : out[y'][x'] = Re sum_ky A[ky][x'] w_y^{fy (py + y')}, y' in [0,ny)
//
// Each 1-D transform is a full length-N mixed-radix (2,3,4,5,8) decimation-in-time FFT held in
// shared memory (in place, digit-reversed scatter on load, natural order on store) with zero-filled
// inputs; the output window is the only thing written.  Several transforms share a CTA so that the
// strided side of each pass still moves whole 32/64-byte segments.  Twiddles come from a per-plan
// table exp(-2*pi*i*k/N) built on the host in extended precision.
#pragma once

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/bldfm_b200.h"
#include "common.cuh"

namespace bldfm {

constexpr int kFftMaxStages = 12;
constexpr int kFftMaxThreads = 384;
constexpr int kFftPad = 2;   // elements of padding between the transforms of one CTA

// shared-memory elements reserved per transform: the XOR swizzle permutes within aligned groups of
// eight, so the last group must exist in full even when N is not a multiple of 8
__host__ __device__ __forceinline__ int fft_stride(int N) { return ((N + 7) & ~7) + kFftPad; }

struct FftPassArgs {
    int32_t N;                 // transform length
    int32_t nstages;
    int32_t radix[kFftMaxStages];
    int32_t lshift[kFftMaxStages];   // log2 of the sub-transform length L entering stage s, or -1
    // inputs : in_freq ? n_in entries in fftfreq order (0..(n-1)/2, -(n/2)..-1)
    //                    : a window [in_off, in_off + n_in) ; everything else is zero
    // outputs: out_freq ? n_out entries in fftfreq order : the window [out_off, out_off + n_out)
    int32_t in_freq, n_in, in_off;
    int32_t out_freq, n_out, out_off;
    // with out_freq: output j is entry (out_foff + j) of an fftfreq-ordered set of out_ftotal entries
    // (out_ftotal = 0 means the n_out entries are the whole set)
    int32_t out_ftotal, out_foff;
    int32_t cw;                // transforms per CTA
    int32_t ntrans;            // transforms per field
    int32_t t_fast;            // 1: consecutive threads walk the transform index first
    int32_t conj_io;           // 1: inverse transform through conj(FFT(conj(x)))
    int64_t in_field_stride, in_tstride, in_kstride;      // in elements
    int64_t out_field_stride, out_tstride, out_jstride;   // in elements
    // optional blocking of the output index j (ky-slab sharding: one block per destination rank):
    // offset = (j / out_block) * out_block_stride + (j % out_block) * out_jstride ; out_block = 0: off
    int32_t out_block;
    int64_t out_block_stride;
    void* const* out_peer;     // optional [n blocks] base pointers (peer GPUs); overrides `out`
    int32_t nfields_first;     // fields [0, nfields_first) use in/out, the rest in2/out2
    const void* in;
    const void* in2;
    void* out;
    void* out2;
    const void* twiddle;       // [N] complex of the compute type: exp(-2*pi*i*k/N)
    const int32_t* rev;        // [N] natural index -> slot in the digit-reversed DIT layout
};

// XOR swizzle of the 16-byte element index inside one transform: keeps every access pattern of the
// DIT stages (stride-1 in j for L >= 8, stride-8 blocks for L = 1) free of shared-memory bank
// conflicts; it permutes elements only within aligned groups of eight.
__device__ __forceinline__ int fft_swz(int i)
{
    return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7);
}

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

template <typename T> __device__ __forceinline__ typename Vec2<T>::type mk2(T x, T y);
template <> __device__ __forceinline__ double2 mk2<double>(double x, double y) { return make_double2(x, y); }
template <> __device__ __forceinline__ float2 mk2<float>(float x, float y) { return make_float2(x, y); }

template <typename T> __device__ __forceinline__ T xfma(T a, T b, T c);
template <> __device__ __forceinline__ double xfma<double>(double a, double b, double c) { return fma(a, b, c); }
template <> __device__ __forceinline__ float xfma<float>(float a, float b, float c) { return fmaf(a, b, c); }

template <typename T>
struct Cplx {
    T r, i;
};

template <typename T> __device__ __forceinline__ Cplx<T> cadd(Cplx<T> a, Cplx<T> b) { return {a.r + b.r, a.i + b.i}; }
template <typename T> __device__ __forceinline__ Cplx<T> csub(Cplx<T> a, Cplx<T> b) { return {a.r - b.r, a.i - b.i}; }
// -i * a
template <typename T> __device__ __forceinline__ Cplx<T> cmuli_neg(Cplx<T> a) { return {a.i, -a.r}; }
template <typename T> __device__ __forceinline__ Cplx<T> cmul(Cplx<T> a, Cplx<T> b)
{
    return {xfma<T>(a.r, b.r, -(a.i * b.i)), xfma<T>(a.r, b.i, a.i * b.r)};
}

// ---- forward butterflies (w_r = exp(-2*pi*i/r)), in place on v[0..r-1]
template <typename T> __device__ __forceinline__ void bfly2(Cplx<T>* v)
{
    const Cplx<T> a = v[0], b = v[1];
    v[0] = cadd(a, b); v[1] = csub(a, b);
}

template <typename T> __device__ __forceinline__ void bfly3(Cplx<T>* v)
{
    const T s = (T)0.86602540378443864676;   // sin(2*pi/3)
    const Cplx<T> t1 = cadd(v[1], v[2]);
    const Cplx<T> m = {xfma<T>((T)-0.5, t1.r, v[0].r), xfma<T>((T)-0.5, t1.i, v[0].i)};
    const Cplx<T> d = {s * (v[1].r - v[2].r), s * (v[1].i - v[2].i)};
    v[0] = cadd(v[0], t1);
    v[1] = {m.r + d.i, m.i - d.r};
    v[2] = {m.r - d.i, m.i + d.r};
}

template <typename T> __device__ __forceinline__ void bfly4(Cplx<T>* v)
{
    const Cplx<T> a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
    const Cplx<T> a2 = cadd(v[1], v[3]), a3 = cmuli_neg(csub(v[1], v[3]));
    v[0] = cadd(a0, a2); v[2] = csub(a0, a2);
    v[1] = cadd(a1, a3); v[3] = csub(a1, a3);
}

template <typename T> __device__ __forceinline__ void bfly5(Cplx<T>* v)
{
    const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;
    const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;
    const Cplx<T> t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
    const Cplx<T> t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    const Cplx<T> m1 = {xfma<T>(c1, t1.r, xfma<T>(c2, t2.r, v[0].r)), xfma<T>(c1, t1.i, xfma<T>(c2, t2.i, v[0].i))};
    const Cplx<T> m2 = {xfma<T>(c2, t1.r, xfma<T>(c1, t2.r, v[0].r)), xfma<T>(c2, t1.i, xfma<T>(c1, t2.i, v[0].i))};
    const Cplx<T> n1 = {xfma<T>(s1, t3.r, s2 * t4.r), xfma<T>(s1, t3.i, s2 * t4.i)};
    const Cplx<T> n2 = {xfma<T>(s2, t3.r, -(s1 * t4.r)), xfma<T>(s2, t3.i, -(s1 * t4.i))};
    v[0] = cadd(v[0], cadd(t1, t2));
    v[1] = {m1.r + n1.i, m1.i - n1.r};     // m1 - i*n1
    v[4] = {m1.r - n1.i, m1.i + n1.r};
    v[2] = {m2.r + n2.i, m2.i - n2.r};
    v[3] = {m2.r - n2.i, m2.i + n2.r};
}

template <typename T> __device__ __forceinline__ void bfly8(Cplx<T>* v)
{
    const T h = (T)0.70710678118654752440;
    Cplx<T> e[4] = {v[0], v[2], v[4], v[6]};
    Cplx<T> o[4] = {v[1], v[3], v[5], v[7]};
    bfly4<T>(e);
    bfly4<T>(o);
    // o[k] *= w_8^k : w^1 = (1-i)/sqrt2, w^2 = -i, w^3 = (-1-i)/sqrt2
    const Cplx<T> o1 = {h * (o[1].r + o[1].i), h * (o[1].i - o[1].r)};
    const Cplx<T> o2 = cmuli_neg(o[2]);
    const Cplx<T> o3 = {h * (o[3].i - o[3].r), -h * (o[3].r + o[3].i)};
    v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);   v[5] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);   v[6] = csub(e[2], o2);
    v[3] = cadd(e[3], o3);   v[7] = csub(e[3], o3);
}

template <typename T, int R>
__device__ __forceinline__ void fft_stage(typename Vec2<T>::type* buf, const typename Vec2<T>::type* __restrict__ tw,
                                          int N, int L, int lshift, int cw, int keep_lo, int keep_hi)
{
    using V = typename Vec2<T>::type;
    const int nbf = N / R;           // butterflies per transform
    const int step = N / (L * R);    // twiddle stride
    int t = threadIdx.x / nbf;
    int jj = threadIdx.x - t * nbf;
    while (t < cw) {
        int blk, j;
        if (lshift >= 0) { blk = jj >> lshift; j = jj & (L - 1); }
        else { blk = (int)((unsigned)jj / (unsigned)L); j = jj - blk * L; }
        V* p = buf + (size_t)t * fft_stride(N);
        const int i0 = blk * L * R + j;
        Cplx<T> v[R];
#pragma unroll
        for (int u = 0; u < R; ++u) { const V x = p[fft_swz(i0 + u * L)]; v[u] = {x.x, x.y}; }
        if (L > 1) {
            // w_u = w^u from one table load: w = exp(-2*pi*i*j/(L*R)); squares/products instead of
            // R-1 dependent global loads (the FP64 pipe has slack here, the load path does not)
            const V w1v = tw[(size_t)j * step];
            const Cplx<T> w1 = {w1v.x, w1v.y};
            if (R == 2) {
                v[1] = cmul<T>(v[1], w1);
            } else if (R == 3) {
                const Cplx<T> w2 = cmul<T>(w1, w1);
                v[1] = cmul<T>(v[1], w1); v[2] = cmul<T>(v[2], w2);
            } else if (R == 4) {
                const Cplx<T> w2 = cmul<T>(w1, w1), w3 = cmul<T>(w2, w1);
                v[1] = cmul<T>(v[1], w1); v[2] = cmul<T>(v[2], w2); v[3] = cmul<T>(v[3], w3);
            } else if (R == 5) {
                const Cplx<T> w2 = cmul<T>(w1, w1), w3 = cmul<T>(w2, w1), w4 = cmul<T>(w2, w2);
                v[1] = cmul<T>(v[1], w1); v[2] = cmul<T>(v[2], w2); v[3] = cmul<T>(v[3], w3);
                v[4] = cmul<T>(v[4], w4);
            } else {
                const Cplx<T> w2 = cmul<T>(w1, w1), w3 = cmul<T>(w2, w1), w4 = cmul<T>(w2, w2);
                const Cplx<T> w5 = cmul<T>(w4, w1), w6 = cmul<T>(w4, w2), w7 = cmul<T>(w4, w3);
                v[1] = cmul<T>(v[1], w1); v[2] = cmul<T>(v[2], w2); v[3] = cmul<T>(v[3], w3);
                v[4] = cmul<T>(v[4], w4); v[5] = cmul<T>(v[5], w5); v[6] = cmul<T>(v[6], w6);
                v[7] = cmul<T>(v[7], w7);
            }
        }
        if (R == 2) bfly2<T>(v);
        else if (R == 3) bfly3<T>(v);
        else if (R == 4) bfly4<T>(v);
        else if (R == 5) bfly5<T>(v);
        else bfly8<T>(v);
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const int i = i0 + u * L;
            if (i >= keep_lo && i < keep_hi) p[fft_swz(i)] = mk2<T>(v[u].r, v[u].i);
        }
        jj += (int)blockDim.x;
        while (jj >= nbf) { jj -= nbf; ++t; }
    }
}

// Generic prime radix (7, 11, 13) as a MIDDLE stage: each thread produces one output
//   y[i0 + q*L] = sum_u x[i0 + u*L] * w^(u*(j*step + q*N/R)),   w = exp(-2*pi*i/N)
// with O(R) table twiddles, out of place (src -> dst halves of the shared buffer).  O(N*R) work for
// this one stage -- it exists so that padded sizes with a factor 7/11/13 (e.g. halo 500 m on a
// 0.39 m grid: 2816 = 2^8*11, 5376 = 2^8*3*7) stay on the in-house pruned path.
template <typename T>
__device__ __forceinline__ void fft_stage_generic(const typename Vec2<T>::type* src, typename Vec2<T>::type* dst,
                                                  const typename Vec2<T>::type* __restrict__ tw, int N, int L,
                                                  int R, int cw)
{
    using V = typename Vec2<T>::type;
    const int LR = L * R;
    const int step = N / LR;
    const int nr = N / R;
    const int total = cw * N;
    for (int e = threadIdx.x; e < total; e += (int)blockDim.x) {
        const int t = (int)((unsigned)e / (unsigned)N);
        const int i = e - t * N;
        const int blk = (int)((unsigned)i / (unsigned)LR);
        const int rem = i - blk * LR;
        const int q = (int)((unsigned)rem / (unsigned)L);
        const int j = rem - q * L;
        const int i0 = blk * LR + j;
        int base = j * step + q * nr;
        if (base >= N) base -= N * (base / N);
        const V* p = src + (size_t)t * fft_stride(N);
        Cplx<T> acc = {(T)0, (T)0};
        int idx = 0;
        for (int u = 0; u < R; ++u) {
            const V x = p[fft_swz(i0 + u * L)];
            const V w = tw[idx];
            acc.r = xfma<T>(x.x, w.x, xfma<T>(-x.y, w.y, acc.r));
            acc.i = xfma<T>(x.x, w.y, xfma<T>(x.y, w.x, acc.i));
            idx += base;
            if (idx >= N) idx -= N;
        }
        dst[(size_t)t * fft_stride(N) + fft_swz(i)] = mk2<T>(acc.r, acc.i);
    }
}

__device__ __forceinline__ bool fft_is_generic(int r) { return r == 7 || r > 8; }

// grid = (ceil(ntrans/cw), nfields) ; block = fft_pick_threads() ; dynamic smem = cw*(N+pad)*sizeof(complex)
template <typename T, bool REAL_IN, bool REAL_OUT>
__global__ void __launch_bounds__(kFftMaxThreads, 2)
k_fft_pass(const FftPassArgs a)
{
    using V = typename Vec2<T>::type;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    V* buf = reinterpret_cast<V*>(fft_smem);
    const int N = a.N;
    const int t0 = blockIdx.x * a.cw;
    const int cw = min(a.cw, a.ntrans - t0);
    const bool second = (int)blockIdx.y >= a.nfields_first;
    const size_t field = second ? blockIdx.y - a.nfields_first : blockIdx.y;
    const void* inp = second ? a.in2 : a.in;
    void* outp = second ? a.out2 : a.out;
    const T sgn = a.conj_io ? (T)-1 : (T)1;

    // zero fill, then scatter the non-zero inputs to their digit-reversed slots
    for (int e = threadIdx.x; e < cw * fft_stride(N); e += (int)blockDim.x) buf[e] = mk2<T>((T)0, (T)0);
    __syncthreads();
    const int npos_in = (a.n_in + 1) / 2;
    for (int e = threadIdx.x; e < cw * a.n_in; e += (int)blockDim.x) {
        int t, k;
        if (a.t_fast) { k = e / cw; t = e - k * cw; }
        else          { t = e / a.n_in; k = e - t * a.n_in; }
        const int i = a.in_freq ? (k < npos_in ? k : k - a.n_in + N) : a.in_off + k;
        const size_t g = field * a.in_field_stride + (size_t)(t0 + t) * a.in_tstride + (size_t)k * a.in_kstride;
        T xr, xi;
        if (REAL_IN) { xr = reinterpret_cast<const T*>(inp)[g]; xi = (T)0; }
        else { const V x = reinterpret_cast<const V*>(inp)[g]; xr = x.x; xi = x.y; }
        buf[(size_t)t * fft_stride(N) + fft_swz(a.rev[i])] = mk2<T>(xr, sgn * xi);
    }
    __syncthreads();

    const V* tw = reinterpret_cast<const V*>(a.twiddle);
    V* other = buf + (size_t)a.cw * fft_stride(N);     // second half, only present for generic radices
    int L = 1;
    for (int s = 0; s < a.nstages; ++s) {
        const int r = a.radix[s];
        // the last stage only needs to keep what the output window will read
        const bool last = (s == a.nstages - 1) && !a.out_freq;
        const int lo = last ? a.out_off : 0, hi = last ? a.out_off + a.n_out : N;
        if (fft_is_generic(r)) {
            fft_stage_generic<T>(buf, other, tw, N, L, r, cw);
            V* tmp = buf; buf = other; other = tmp;
        } else {
            switch (r) {
                case 2: fft_stage<T, 2>(buf, tw, N, L, a.lshift[s], cw, lo, hi); break;
                case 3: fft_stage<T, 3>(buf, tw, N, L, a.lshift[s], cw, lo, hi); break;
                case 4: fft_stage<T, 4>(buf, tw, N, L, a.lshift[s], cw, lo, hi); break;
                case 5: fft_stage<T, 5>(buf, tw, N, L, a.lshift[s], cw, lo, hi); break;
                default: fft_stage<T, 8>(buf, tw, N, L, a.lshift[s], cw, lo, hi); break;
            }
        }
        L *= r;
        __syncthreads();
    }

    // store the requested outputs
    const int ftotal = a.out_ftotal > 0 ? a.out_ftotal : a.n_out;
    const int npos_out = (ftotal + 1) / 2;
    for (int e = threadIdx.x; e < cw * a.n_out; e += (int)blockDim.x) {
        int t, j;
        if (a.t_fast) { j = e / cw; t = e - j * cw; }
        else          { t = e / a.n_out; j = e - t * a.n_out; }
        const int kk = j + a.out_foff;
        const int i = a.out_freq ? (kk < npos_out ? kk : kk - ftotal + N) : a.out_off + j;
        const V x = buf[(size_t)t * fft_stride(N) + fft_swz(i)];
        size_t o = field * a.out_field_stride + (size_t)(t0 + t) * a.out_tstride;
        void* dst = outp;
        if (a.out_block > 0) {
            const int blk = j / a.out_block;
            const int jb = j - blk * a.out_block;
            if (a.out_peer) dst = a.out_peer[blk];
            else o += (size_t)blk * a.out_block_stride;
            o += (size_t)jb * a.out_jstride;
        } else {
            o += (size_t)j * a.out_jstride;
        }
        if (REAL_OUT) reinterpret_cast<T*>(dst)[o] = x.x;
        else reinterpret_cast<V*>(dst)[o] = mk2<T>(x.x, sgn * x.y);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
inline bool fft_factorize(int n, std::vector<int>& radix)
{
    radix.clear();
    if (n < 2) return false;
    // powers of two first: the sub-transform length L stays a power of two (shift/mask indexing)
    // until the 3s and 5s come in at the end; generic primes (7, 11, 13) go in the middle because
    // the first and the last stage are specialised (register-resident) for radices 2,3,4,5,8.
    std::vector<int> pool, generic, odd;
    while (n % 8 == 0) { pool.push_back(8); n /= 8; }
    while (n % 4 == 0) { pool.push_back(4); n /= 4; }
    while (n % 2 == 0) { pool.push_back(2); n /= 2; }
    while (n % 3 == 0) { odd.push_back(3); n /= 3; }
    while (n % 5 == 0) { odd.push_back(5); n /= 5; }
    for (int p : {7, 11, 13})
        while (n % p == 0) { generic.push_back(p); n /= p; }
    if (n != 1) return false;
    pool.insert(pool.end(), odd.begin(), odd.end());
    if (generic.empty()) {
        radix = pool;
    } else {
        if (pool.size() < 2) return false;
        radix.assign(pool.begin(), pool.end() - 1);
        radix.insert(radix.end(), generic.begin(), generic.end());
        radix.push_back(pool.back());
    }
    return !radix.empty() && (int)radix.size() <= kFftMaxStages;
}

inline bool fft_has_generic(int n)
{
    return n % 7 == 0 || n % 11 == 0 || n % 13 == 0;
}

// fills radix[], lshift[] of a pass
inline void fft_set_stages(FftPassArgs& a, const std::vector<int>& radix)
{
    a.nstages = (int)radix.size();
    int L = 1;
    for (size_t i = 0; i < radix.size(); ++i) {
        a.radix[i] = radix[i];
        int sh = -1;
        if ((L & (L - 1)) == 0) { sh = 0; while ((1 << sh) < L) ++sh; }
        a.lshift[i] = sh;
        L *= radix[i];
    }
}

// rev[i] = slot of natural input index i in the digit-reversed layout of the DIT stages
inline void fft_rev_table(int N, const std::vector<int>& radix, std::vector<int32_t>& rev)
{
    rev.resize((size_t)N);
    for (int i0 = 0; i0 < N; ++i0) {
        int i = i0, pos = 0, ncur = N;
        for (int s = (int)radix.size() - 1; s >= 0; --s) {
            const int r = radix[(size_t)s];
            ncur /= r;
            pos += (i % r) * ncur;
            i /= r;
        }
        rev[(size_t)i0] = pos;
    }
}

inline size_t fft_smem_bytes(int N, int cw, bool f32)
{
    // sizes with a generic-radix stage need a second (out-of-place) half
    return (size_t)(fft_has_generic(N) ? 2 : 1) * (size_t)cw * (size_t)fft_stride(N) *
           (f32 ? sizeof(float2) : sizeof(double2));
}

// Tuning switches.  Each is read from the environment ONCE (first use) and cached -- the launch path never
// calls getenv() again -- and can be overridden at run time with bldfm_set_option() (tests, sweeps).
struct OptionTable {
    std::mutex mu;
    std::unordered_map<std::string, int> val;
};
inline OptionTable& fft_options()
{
    static OptionTable t;
    return t;
}
inline int fft_env_int(const char* name, int dflt)
{
    OptionTable& t = fft_options();
    std::lock_guard<std::mutex> lock(t.mu);
    auto it = t.val.find(name);
    if (it != t.val.end()) return it->second;
    const char* v = std::getenv(name);
    const int r = (v && *v) ? std::atoi(v) : dflt;
    t.val.emplace(name, r);
    return r;
}
// value = INT_MIN forgets the cached value (the environment / default is consulted again)
inline void fft_set_option(const char* name, int value)
{
    OptionTable& t = fft_options();
    std::lock_guard<std::mutex> lock(t.mu);
    if (value == INT_MIN) t.val.erase(name);
    else t.val[name] = value;
}

// transforms per CTA for the two passes given the shared-memory budget
// `ntrans_total` (transforms of the whole launch) shrinks cw for small launches so that the grid
// still covers every SM a few times -- a 512x512 single-level solve has only ~500 transforms per pass.
inline int fft_pick_cw(int N, bool f32, size_t smem_optin, int want, int64_t ntrans_total = -1, int num_sms = 148)
{
    int cw = fft_env_int("BLDFM_FFT_CW", 0);
    if (cw <= 0) {
        cw = want;
        if (ntrans_total > 0)
            while (cw > 1 && ntrans_total / cw < (int64_t)3 * num_sms) cw /= 2;
    }
    while (cw > 1 && fft_smem_bytes(N, cw, f32) > smem_optin / 2) cw /= 2;   // keep >= 2 CTAs / SM
    return cw;
}

// threads per CTA: the widest stage (largest radix first) has cw*N/r butterflies; give each thread
// a whole number of them where possible
inline int fft_pick_threads(int N, int cw, int r0)
{
    const int forced = fft_env_int("BLDFM_FFT_THREADS", 0);
    if (forced > 0) return std::min(kFftMaxThreads, std::max(32, forced / 32 * 32));
    const int nb = cw * (N / std::max(r0, 1));
    int best = 256;
    for (int t : {384, 256, 320, 192, 128}) {
        if (t <= nb && nb % t == 0) { best = t; break; }
    }
    if (nb < 128) best = std::max(32, (nb + 31) / 32 * 32);
    return best;
}

inline bool pruned_fft_supported(const bldfm_geometry& g, bool f32, size_t smem_optin)
{
    std::vector<int> r;
    if (g.nfx < 2 || g.nfy < 2) return false;
    if (!fft_factorize(g.nfx, r) || !fft_factorize(g.nfy, r)) return false;
    if (g.px + g.nx > g.nfx || g.py + g.ny > g.nfy) return false;
    return fft_smem_bytes(g.nfx, 1, f32) <= smem_optin && fft_smem_bytes(g.nfy, 1, f32) <= smem_optin;
}

inline size_t pruned_fft_work_bytes(const bldfm_geometry& g, bool f32, int64_t nfields)
{
    const int64_t chunk = std::min<int64_t>(nfields, 16384);
    return (size_t)2 * (size_t)chunk * (size_t)g.nly * (size_t)g.nx * (f32 ? sizeof(float2) : sizeof(double2));
}

// exp(-2*pi*i*k/N) evaluated in x87 extended precision and rounded to double (host)
inline void fft_twiddles(int N, std::vector<double>& out)
{
    out.resize((size_t)2 * N);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int k = 0; k < N; ++k) {
        const long double x = two_pi * (long double)k / (long double)N;
        out[(size_t)2 * k] = (double)cosl(x);
        out[(size_t)2 * k + 1] = (double)(-sinl(x));
    }
}

struct PrunedFftTables {
    const void* tw_x = nullptr;   // device, compute type
    const void* tw_y = nullptr;
    const int32_t* rev_x = nullptr;
    const int32_t* rev_y = nullptr;
    const void* t24_x = nullptr;  // lane-contiguous twiddle tables of the N = 3P passes (fft24.cuh), or NULL
    const void* t24_y = nullptr;
    const void* t48_x = nullptr;  // stage-1 table of the two-stage variant (fft48.cuh), or NULL
    const void* t48_y = nullptr;
};

// Runs both passes for the `nfields` compact spectra of spec_p (-> out_p) and of spec_q (-> out_q)
// in two launches.  `work` holds 2*nfields intermediate fields [nly][nx] complex.
template <typename T>
inline cudaError_t pruned_fft_launch(cudaStream_t stream, size_t smem_optin, const bldfm_geometry& g,
                                     bool forward_dir, const void* spec_p, const void* spec_q,
                                     int64_t nfields, void* work, void* out_p, void* out_q,
                                     const PrunedFftTables& tab, int* nlaunch)
{
    using V = typename Vec2<T>::type;
    const bool f32 = sizeof(T) == 4;
    std::vector<int> rx, ry;
    fft_factorize(g.nfx, rx);
    fft_factorize(g.nfy, ry);

    FftPassArgs ax{};
    ax.N = g.nfx; fft_set_stages(ax, rx); ax.rev = tab.rev_x;
    ax.in_freq = 1; ax.n_in = g.nlx; ax.in_off = 0; ax.out_freq = 0; ax.n_out = g.nx; ax.out_off = g.px;
    ax.cw = fft_pick_cw(g.nfx, f32, smem_optin, 4);
    ax.ntrans = g.nly; ax.t_fast = 0; ax.conj_io = forward_dir ? 0 : 1;
    ax.in_field_stride = (int64_t)g.nly * g.nlx; ax.in_tstride = g.nlx; ax.in_kstride = 1;
    ax.out_field_stride = (int64_t)g.nly * g.nx; ax.out_tstride = g.nx; ax.out_jstride = 1;
    ax.twiddle = tab.tw_x;

    FftPassArgs ay{};
    ay.N = g.nfy; fft_set_stages(ay, ry); ay.rev = tab.rev_y;
    ay.in_freq = 1; ay.n_in = g.nly; ay.in_off = 0; ay.out_freq = 0; ay.n_out = g.ny; ay.out_off = g.py;
    ay.cw = fft_pick_cw(g.nfy, f32, smem_optin, 4);
    ay.ntrans = g.nx; ay.t_fast = 1; ay.conj_io = forward_dir ? 0 : 1;
    ay.in_field_stride = (int64_t)g.nly * g.nx; ay.in_tstride = 1; ay.in_kstride = g.nx;
    ay.out_field_stride = (int64_t)g.ny * g.nx; ay.out_tstride = 1; ay.out_jstride = g.nx;
    ay.twiddle = tab.tw_y;

    const size_t sx = fft_smem_bytes(ax.N, ax.cw, f32), sy = fft_smem_bytes(ay.N, ay.cw, f32);
    cudaError_t e;
    e = set_max_dyn_smem(k_fft_pass<T, false, false>, (int)smem_optin);
    if (e != cudaSuccess) return e;
    e = set_max_dyn_smem(k_fft_pass<T, false, true>, (int)smem_optin);
    if (e != cudaSuccess) return e;

    // gridDim.y is limited to 65535: split the field batch if needed
    const int64_t chunk = 16384;
    for (int64_t f0 = 0; f0 < nfields; f0 += chunk) {
        const int nf = (int)std::min<int64_t>(chunk, nfields - f0);
        FftPassArgs bx = ax, by = ay;
        bx.nfields_first = nf; by.nfields_first = nf;
        bx.in = reinterpret_cast<const V*>(spec_p) + (size_t)f0 * ax.in_field_stride;
        bx.in2 = reinterpret_cast<const V*>(spec_q) + (size_t)f0 * ax.in_field_stride;
        bx.out = reinterpret_cast<V*>(work);
        bx.out2 = reinterpret_cast<V*>(work) + (size_t)nf * ax.out_field_stride;
        by.in = bx.out; by.in2 = bx.out2;
        by.out = reinterpret_cast<T*>(out_p) + (size_t)f0 * ay.out_field_stride;
        by.out2 = reinterpret_cast<T*>(out_q) + (size_t)f0 * ay.out_field_stride;
        k_fft_pass<T, false, false><<<dim3((unsigned)((ax.ntrans + ax.cw - 1) / ax.cw), (unsigned)(2 * nf)),
                                      fft_pick_threads(ax.N, ax.cw, ax.radix[0]), sx, stream>>>(bx);
        k_fft_pass<T, false, true><<<dim3((unsigned)((ay.ntrans + ay.cw - 1) / ay.cw), (unsigned)(2 * nf)),
                                     fft_pick_threads(ay.N, ay.cw, ay.radix[0]), sy, stream>>>(by);
        *nlaunch += 2;
    }
    return cudaGetLastError();
}

// ---- ky-slab sharded back-transform ------------------------------------------------
// stage 1 (every rank): x-transform of the local rows [ky0, ky0+rows) of each field; the output index
// x' is blocked by destination rank: send[field][dst][rows][nx/G]  (or written straight into the
// peers' receive buffers when `peer_p/peer_q` are given -- the transpose then rides on the stores).
inline cudaError_t sharded_xpass(cudaStream_t stream, size_t smem_optin, const bldfm_geometry& g,
                                 bool forward_dir, int rows, int nranks, const void* spec_p, const void* spec_q,
                                 int nfields, void* send_p, void* send_q, void* const* peer_p,
                                 void* const* peer_q, int64_t peer_field_stride, const PrunedFftTables& tab,
                                 int* nlaunch)
{
    std::vector<int> rx;
    fft_factorize(g.nfx, rx);
    const int nxl = g.nx / nranks;
    FftPassArgs ax{};
    ax.N = g.nfx; fft_set_stages(ax, rx); ax.rev = tab.rev_x;
    ax.in_freq = 1; ax.n_in = g.nlx; ax.in_off = 0; ax.out_freq = 0; ax.n_out = g.nx; ax.out_off = g.px;
    ax.cw = fft_pick_cw(g.nfx, false, smem_optin, 4);
    ax.ntrans = rows; ax.t_fast = 0; ax.conj_io = forward_dir ? 0 : 1;
    ax.in_field_stride = (int64_t)rows * g.nlx; ax.in_tstride = g.nlx; ax.in_kstride = 1;
    ax.out_tstride = nxl; ax.out_jstride = 1; ax.out_block = nxl;
    if (peer_p) {
        ax.out_field_stride = peer_field_stride;      // receiver layout [field][nly][nx/G]
        ax.out_block_stride = 0;
        ax.out_peer = peer_p;
    } else {
        ax.out_field_stride = (int64_t)rows * g.nx;   // [field][dst][rows][nx/G]
        ax.out_block_stride = (int64_t)rows * nxl;
    }
    ax.twiddle = tab.tw_x;
    ax.nfields_first = nfields;
    ax.in = spec_p; ax.in2 = spec_q; ax.out = send_p; ax.out2 = send_q;
    cudaError_t e = set_max_dyn_smem(k_fft_pass<double, false, false>, (int)smem_optin);
    if (e != cudaSuccess) return e;
    const dim3 grid((unsigned)((ax.ntrans + ax.cw - 1) / ax.cw), (unsigned)nfields);
    const int th = fft_pick_threads(ax.N, ax.cw, ax.radix[0]);
    const size_t sm = fft_smem_bytes(ax.N, ax.cw, false);
    if (peer_p) {
        // two launches: the p fields then the q fields (each has its own peer pointer table)
        FftPassArgs a1 = ax; a1.in2 = spec_p; a1.out2 = send_p;
        k_fft_pass<double, false, false><<<grid, th, sm, stream>>>(a1);
        FftPassArgs a2 = ax; a2.in = spec_q; a2.in2 = spec_q; a2.out_peer = peer_q;
        k_fft_pass<double, false, false><<<grid, th, sm, stream>>>(a2);
        *nlaunch += 2;
    } else {
        const dim3 grid2(grid.x, (unsigned)(2 * nfields));
        k_fft_pass<double, false, false><<<grid2, th, sm, stream>>>(ax);
        *nlaunch += 1;
    }
    return cudaGetLastError();
}

// stage 2 (every rank): y-transform of the received [field][nly][nx/G] -> real [field][ny][nx/G]
inline cudaError_t sharded_ypass(cudaStream_t stream, size_t smem_optin, const bldfm_geometry& g,
                                 bool forward_dir, int nranks, const void* recv_p, const void* recv_q,
                                 int nfields, void* out_p, void* out_q, const PrunedFftTables& tab, int* nlaunch)
{
    std::vector<int> ry;
    fft_factorize(g.nfy, ry);
    const int nxl = g.nx / nranks;
    FftPassArgs ay{};
    ay.N = g.nfy; fft_set_stages(ay, ry); ay.rev = tab.rev_y;
    ay.in_freq = 1; ay.n_in = g.nly; ay.in_off = 0; ay.out_freq = 0; ay.n_out = g.ny; ay.out_off = g.py;
    ay.cw = fft_pick_cw(g.nfy, false, smem_optin, 4);
    ay.ntrans = nxl; ay.t_fast = 1; ay.conj_io = forward_dir ? 0 : 1;
    ay.in_field_stride = (int64_t)g.nly * nxl; ay.in_tstride = 1; ay.in_kstride = nxl;
    ay.out_field_stride = (int64_t)g.ny * nxl; ay.out_tstride = 1; ay.out_jstride = nxl;
    ay.twiddle = tab.tw_y;
    ay.nfields_first = nfields;
    ay.in = recv_p; ay.in2 = recv_q; ay.out = out_p; ay.out2 = out_q;
    cudaError_t e = set_max_dyn_smem(k_fft_pass<double, false, true>, (int)smem_optin);
    if (e != cudaSuccess) return e;
    k_fft_pass<double, false, true><<<dim3((unsigned)((ay.ntrans + ay.cw - 1) / ay.cw), (unsigned)(2 * nfields)),
                                      fft_pick_threads(ay.N, ay.cw, ay.radix[0]),
                                      fft_smem_bytes(ay.N, ay.cw, false), stream>>>(ay);
    *nlaunch += 1;
    return cudaGetLastError();
}

// Pruned forward transform of the padded source (K1+K2+K3 fused): q0[ny][nx] real (its zero halo is
// implicit) -> unnormalised spectrum on the retained modes, compact [nly][nlx] c128.
// `work` holds [ny][nlx] c128.  Two launches.
inline cudaError_t pruned_fft_forward(cudaStream_t stream, size_t smem_optin, const bldfm_geometry& g,
                                      const double* q0, void* work, void* spec, const PrunedFftTables& tab,
                                      int* nlaunch, int ky0 = 0, int rows = -1, bool skip_x = false)
{
    if (rows < 0) rows = g.nly;      // ky-slab sharding: only rows [ky0, ky0+rows) of the spectrum
    std::vector<int> rx, ry;
    fft_factorize(g.nxe, rx);
    fft_factorize(g.nye, ry);
    FftPassArgs ax{};
    ax.N = g.nxe; fft_set_stages(ax, rx); ax.rev = tab.rev_x;
    ax.in_freq = 0; ax.n_in = g.nx; ax.in_off = g.px; ax.out_freq = 1; ax.n_out = g.nlx; ax.out_off = 0;
    ax.cw = fft_pick_cw(g.nxe, false, smem_optin, 4);
    ax.ntrans = g.ny; ax.t_fast = 0; ax.conj_io = 0;
    ax.in_field_stride = 0; ax.in_tstride = g.nx; ax.in_kstride = 1;
    ax.out_field_stride = 0; ax.out_tstride = g.nlx; ax.out_jstride = 1;
    ax.nfields_first = 1; ax.in = q0; ax.in2 = q0; ax.out = work; ax.out2 = work; ax.twiddle = tab.tw_x;

    FftPassArgs ay{};
    ay.N = g.nye; fft_set_stages(ay, ry); ay.rev = tab.rev_y;
    ay.in_freq = 0; ay.n_in = g.ny; ay.in_off = g.py; ay.out_freq = 1; ay.n_out = rows; ay.out_off = 0;
    ay.out_ftotal = g.nly; ay.out_foff = ky0;
    ay.cw = fft_pick_cw(g.nye, false, smem_optin, 4);
    ay.ntrans = g.nlx; ay.t_fast = 1; ay.conj_io = 0;
    ay.in_field_stride = 0; ay.in_tstride = 1; ay.in_kstride = g.nlx;
    ay.out_field_stride = 0; ay.out_tstride = 1; ay.out_jstride = g.nlx;
    ay.nfields_first = 1; ay.in = work; ay.in2 = work; ay.out = spec; ay.out2 = spec; ay.twiddle = tab.tw_y;

    cudaError_t e;
    e = set_max_dyn_smem(k_fft_pass<double, true, false>, (int)smem_optin);
    if (e != cudaSuccess) return e;
    e = set_max_dyn_smem(k_fft_pass<double, false, false>, (int)smem_optin);
    if (e != cudaSuccess) return e;
    if (!skip_x) {     // `work` already holds the x-pass when a second row range of the same source is asked for
        k_fft_pass<double, true, false><<<dim3((unsigned)((ax.ntrans + ax.cw - 1) / ax.cw), 1),
                                          fft_pick_threads(ax.N, ax.cw, ax.radix[0]),
                                          fft_smem_bytes(ax.N, ax.cw, false), stream>>>(ax);
        *nlaunch += 1;
    }
    k_fft_pass<double, false, false><<<dim3((unsigned)((ay.ntrans + ay.cw - 1) / ay.cw), 1),
                                       fft_pick_threads(ay.N, ay.cw, ay.radix[0]),
                                       fft_smem_bytes(ay.N, ay.cw, false), stream>>>(ay);
    *nlaunch += 1;
    return cudaGetLastError();
}

}  // namespace bldfm
