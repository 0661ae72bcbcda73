// fft_herm.cuh -- real-output back-transform on half the work (K9 + K10 + K11 fused).
//
// The fields the solver returns are the REAL part of a 2-D transform of the retained spectrum S
// (src/bldfm/solver.py:282-287).  Re F[S] = F[H] with H = (S + S~)/2, S~[fy][fx] = conj(S[-fy][-fx]),
// and F[H] is exactly real because H is Hermitian.  That symmetry halves both passes:
//
//   pass X  rows fy = 0 .. nly/2 only (A[-fy][x] = conj(A[fy][x]) is implied); the Hermitian combine
//           of the two source rows happens in the load of the first butterfly stage;
//   pass Y  two columns per complex transform: c[f] = A1[f] + i*A2[f] for f >= 0 and
//           conj(A1[-f]) + i*conj(A2[-f]) for f < 0, so that F[c] = out1 + i*out2.
//
// Each transform is the same in-place shared-memory mixed-radix DIT FFT as fft.cuh, but the first
// stage takes its operands straight from global memory (no zero-fill / scatter pass) and the last
// stage stores the output window straight from registers (no final shared-memory round trip).
#pragma once

#include "fft.cuh"

namespace bldfm {

struct FftHArgs {
    int32_t N, nstages;
    int32_t radix[kFftMaxStages];
    int32_t lshift[kFftMaxStages];
    int32_t cw;                // transforms per CTA (rows in pass X, column pairs in pass Y)
    int32_t ntrans;            // transforms per field
    int32_t conj_io;           // 1: inverse transform through conj(FFT(conj(x)))
    int32_t nlx, nly;          // retained modes of S
    int32_t nrow;              // rows of A = nly/2 + 1
    int32_t nx;                // kept columns
    int32_t out_off, n_out;    // output window of this pass
    int32_t nfields_first;
    int32_t hs;                // pass X: S is conjugate-symmetric (half-plane march); used by fft24.cuh
    int32_t tfast;             // pass X: transform index fastest across lanes (BLDFM_B200_FFT_TFAST)
    // ky-slab sharding (pass X): this launch transforms the rows row0 .. row0+ntrans-1 of A and blocks
    // the kept columns by destination rank: out[field][blk][row-row0][out_block] (out_block = nx/G), or
    // straight into the peers' receive buffers out_peer[blk][field][row-row0][out_block]
    int32_t row0, out_block;
    int64_t out_field_stride, out_block_stride;
    void* const* out_peer;
    const void* in;            // pass X: S [field][nly][nlx]      pass Y: A [field][nrow][nx]
    const void* in2;
    void* out;                 // pass X: A [field][nrow][nx]      pass Y: real [field][ny][nx]
    void* out2;
    const void* twiddle;
    const int32_t* rev;
    const void* tw24;          // fft24.cuh: [24][Q] stage-1 table | [Q] | [Q/r0] in-place stage tables
    const void* tw48;          // fft48.cuh: [48][Q] stage-1 table
};

// S[fy][fx] with zero outside the retained set (signed frequencies)
template <typename T>
__device__ __forceinline__ Cplx<T> herm_s(const typename Vec2<T>::type* __restrict__ S, int nlx, int nly, int fy, int fx)
{
    if (fy < -(nly / 2) || fy > (nly - 1) / 2 || fx < -(nlx / 2) || fx > (nlx - 1) / 2) return {(T)0, (T)0};
    const int ky = fy >= 0 ? fy : fy + nly;
    const int kx = fx >= 0 ? fx : fx + nlx;
    const typename Vec2<T>::type v = S[(size_t)ky * nlx + kx];
    return {v.x, v.y};
}

// pass X operand: sum over the signed frequencies f == i (mod N), |f| <= nlx/2, of H[fy][f]
template <typename T>
__device__ __forceinline__ Cplx<T> herm_load_x(const FftHArgs& a, const typename Vec2<T>::type* __restrict__ S, int fy, int i)
{
    const int hmax = a.nlx / 2;
    Cplx<T> acc = {(T)0, (T)0};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int f = k == 0 ? i : i - a.N;     // the signed frequencies congruent to i
        if (f > hmax || f < -hmax) continue;
        const Cplx<T> s1 = herm_s<T>(S, a.nlx, a.nly, fy, f);
        const Cplx<T> s2 = herm_s<T>(S, a.nlx, a.nly, -fy, -f);
        acc.r += (T)0.5 * (s1.r + s2.r);
        acc.i += (T)0.5 * (s1.i - s2.i);
    }
    return acc;
}

// pass Y operand for the column pair (x1, x1+1)
template <typename T>
__device__ __forceinline__ Cplx<T> herm_load_y(const FftHArgs& a, const typename Vec2<T>::type* __restrict__ A, int x1, int i)
{
    const int hmax = a.nly / 2;
    Cplx<T> acc = {(T)0, (T)0};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int f = k == 0 ? i : i - a.N;
        if (f > hmax || f < -hmax) continue;
        const int af = f >= 0 ? f : -f;
        const typename Vec2<T>::type v1 = A[(size_t)af * a.nx + x1];
        typename Vec2<T>::type v2 = mk2<T>((T)0, (T)0);
        if (x1 + 1 < a.nx) v2 = A[(size_t)af * a.nx + x1 + 1];
        if (f == 0) { acc.r += v1.x; acc.i += v2.x; }                    // A[0] is real
        else if (f > 0) { acc.r += v1.x - v2.y; acc.i += v1.y + v2.x; }  // A1 + i*A2
        else { acc.r += v1.x + v2.y; acc.i += v2.x - v1.y; }             // conj(A1) + i*conj(A2)
    }
    return acc;
}

// Two adjacent complex elements in ONE load: pass Y packs the column pair (x, x+1) of the intermediate, i.e. every
// thread reads 32 contiguous bytes (complex128).  As two LDG.128 each warp instruction touches every 128-byte
// line it needs but uses only half of it -- ncu: 7.8 data-pipe wavefronts per request in pass Y against 4.4 in
// pass X -- and the LSU data pipe is what bounds these kernels.  sm_100 has 256-bit global loads (LDG.E.256).
// `p` must be aligned to the size of the pair.
__device__ __forceinline__ void ld_pair(const double2* p, double2& x1, double2& x2)
{
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(x1.x), "=d"(x1.y), "=d"(x2.x), "=d"(x2.y) : "l"(p));
}
__device__ __forceinline__ void ld_pair(const float2* p, float2& x1, float2& x2)
{
    const float4 v = *reinterpret_cast<const float4*>(p);
    x1 = make_float2(v.x, v.y);
    x2 = make_float2(v.z, v.w);
}

// Operand loads are split in two steps so that all the global loads of a butterfly are in flight before
// the first one is consumed: `fetch` is branch-free (out-of-range elements read a valid
// dummy address and are zeroed by a select), `combine` is pure arithmetic.
template <typename T> struct Fft24Raw { typename Vec2<T>::type a, b; int flag; };

// pass X operand H[fy][f] = (S[fy][f] + conj(S[-fy][-f]))/2 for a signed frequency f, fy >= 0
template <typename T>
__device__ __forceinline__ Fft24Raw<T> fft24_fetch_x(const FftHArgs& a, const typename Vec2<T>::type* __restrict__ S, int fy, int f)
{
    const int hx = a.nlx / 2, hy = a.nly / 2, px = (a.nlx - 1) / 2, py = (a.nly - 1) / 2;
    // conjugate-symmetric spectrum (half-plane march): in the interior the two terms are equal bit for bit
    const bool one = a.hs && f > -hx && f < hx && fy > 0 && fy < hy;
    const bool ok1 = fy <= py && f >= -hx && f <= px;                   // S[fy][f] is a retained mode
    const bool ok2 = !one && fy <= hy && -f >= -hx && -f <= px;         // S[-fy][-f] is one
    const int i1 = ok1 ? fy * a.nlx + (f >= 0 ? f : f + a.nlx) : 0;
    const int i2 = ok2 ? (fy > 0 ? a.nly - fy : 0) * a.nlx + (f > 0 ? a.nlx - f : -f) : i1;
    Fft24Raw<T> r;
    r.a = S[i1];
    r.b = S[i2];
    r.flag = (one ? 4 : 0) | (ok1 ? 1 : 0) | (ok2 ? 2 : 0);
    return r;
}

template <typename T>
__device__ __forceinline__ Cplx<T> fft24_combine_x(const Fft24Raw<T>& r)
{
    const T ax = (r.flag & 1) ? r.a.x : (T)0, ay = (r.flag & 1) ? r.a.y : (T)0;
    const T bx = (r.flag & 2) ? r.b.x : (T)0, by = (r.flag & 2) ? r.b.y : (T)0;
    const Cplx<T> h = {(T)0.5 * (ax + bx), (T)0.5 * (ay - by)};
    return (r.flag & 4) ? Cplx<T>{r.a.x, r.a.y} : h;
}

// pass Y operand of the column pair (x1, x1+1) for a signed frequency f, |f| <= nly/2
template <typename T>
__device__ __forceinline__ Fft24Raw<T> fft24_fetch_y(const FftHArgs& a, const typename Vec2<T>::type* __restrict__ A, int x1, int f)
{
    const int af0 = f >= 0 ? f : -f;
    const bool ok = af0 <= a.nly / 2;                         // callers with arbitrary N may ask beyond the band
    const int af = ok ? af0 : 0;
    const bool pair = x1 + 1 < a.nx;
    Fft24Raw<T> r;
    r.a = A[(size_t)af * a.nx + x1];
    r.b = A[(size_t)af * a.nx + x1 + (pair ? 1 : 0)];
    r.flag = (pair ? 1 : 0) | (f == 0 ? 2 : 0) | (f < 0 ? 4 : 0) | (ok ? 0 : 8);
    return r;
}

template <typename T>
__device__ __forceinline__ Cplx<T> fft24_combine_y(const Fft24Raw<T>& r)
{
    if (r.flag & 8) return {(T)0, (T)0};
    const T bx = (r.flag & 1) ? r.b.x : (T)0, by = (r.flag & 1) ? r.b.y : (T)0;
    if (r.flag & 2) return {r.a.x, bx};                      // A[0] is real
    const T sg = (r.flag & 4) ? (T)-1 : (T)1;                // f > 0: A1 + i*A2 ; f < 0: conj(A1) + i*conj(A2)
    return {r.a.x - sg * by, sg * r.a.y + bx};
}

template <typename T, int PASS>
__device__ __forceinline__ void herm_emit(const FftHArgs& a, void* outp, size_t field, int tg, int i, Cplx<T> v, T sgn)
{
    using V = typename Vec2<T>::type;
    const int o = i - a.out_off;
    if (o < 0 || o >= a.n_out) return;
    if (PASS == 0) {
        if (a.out_block > 0) {
            const int blk = o / a.out_block;
            const int jb = o - blk * a.out_block;
            V* dst = reinterpret_cast<V*>(a.out_peer ? a.out_peer[blk] : outp);
            const size_t off = field * (size_t)a.out_field_stride + (a.out_peer ? 0 : blk * (size_t)a.out_block_stride) +
                               (size_t)(tg - a.row0) * a.out_block + jb;
            dst[off] = mk2<T>(v.r, sgn * v.i);
        } else {
            reinterpret_cast<V*>(outp)[(field * a.nrow + tg) * (size_t)a.nx + o] = mk2<T>(v.r, sgn * v.i);
        }
    } else {
        T* dst = reinterpret_cast<T*>(outp) + (field * a.n_out + o) * (size_t)a.nx + 2 * tg;
        const T im = sgn * v.i;
        if (2 * tg + 1 < a.nx) {
            if ((reinterpret_cast<uintptr_t>(dst) & (sizeof(V) - 1)) == 0) *reinterpret_cast<V*>(dst) = mk2<T>(v.r, im);
            else { dst[0] = v.r; dst[1] = im; }
        } else {
            dst[0] = v.r;
        }
    }
}

template <typename T, int PASS, int R>
__device__ __forceinline__ void herm_first_stage(const FftHArgs& a, typename Vec2<T>::type* buf,
                                                 const typename Vec2<T>::type* __restrict__ src, void* outp,
                                                 size_t field, int t0, int cw, T sgn)
{
    const int N = a.N;
    const int nb = N / R;
    const int total = nb * cw;
    for (int b = threadIdx.x; b < total; b += (int)blockDim.x) {
        int t, il;
        if (PASS == 1) { il = b / cw; t = b - il * cw; }     // lanes walk the column pairs first
        else { t = b / nb; il = b - t * nb; }                // lanes walk the frequency index
        // (fetching the operands as a branch-free batch, as fft24.cuh does, was measured here too: 7 % slower --
        // this kernel is bound by instruction issue, not by the latency of its loads)
        Cplx<T> v[R];
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const int i = il + u * nb;
            v[u] = PASS == 0 ? herm_load_x<T>(a, src, t0 + t, i) : herm_load_y<T>(a, src, 2 * (t0 + t), i);
            v[u].i *= sgn;
        }
        if (R == 2) bfly2<T>(v);
        else if (R == 3) bfly3<T>(v);
        else if (R == 4) bfly4<T>(v);
        else if (R == 5) bfly5<T>(v);
        else bfly8<T>(v);
        if (a.nstages == 1) {
#pragma unroll
            for (int q = 0; q < R; ++q) herm_emit<T, PASS>(a, outp, field, t0 + t, q, v[q], sgn);
        } else {
            const int base = a.rev[il];
            typename Vec2<T>::type* p = buf + (size_t)t * fft_stride(N);
#pragma unroll
            for (int q = 0; q < R; ++q) p[fft_swz(base + q)] = mk2<T>(v[q].r, v[q].i);
        }
    }
}

template <typename T, int PASS, int R>
__device__ __forceinline__ void herm_last_stage(const FftHArgs& a, const typename Vec2<T>::type* buf,
                                                const typename Vec2<T>::type* __restrict__ tw, void* outp,
                                                size_t field, int t0, int cw, T sgn)
{
    using V = typename Vec2<T>::type;
    const int N = a.N;
    const int L = N / R;
    const int total = L * cw;
    for (int b = threadIdx.x; b < total; b += (int)blockDim.x) {
        int t, j;
        if (PASS == 1) { j = b / cw; t = b - j * cw; }
        else { t = b / L; j = b - t * L; }
        // skip butterflies none of whose outputs lie in the window
        bool any = false;
#pragma unroll
        for (int q = 0; q < R; ++q) { const int o = j + q * L - a.out_off; any |= (o >= 0 && o < a.n_out); }
        if (!any) continue;
        const V* p = buf + (size_t)t * fft_stride(N);
        Cplx<T> v[R];
#pragma unroll
        for (int u = 0; u < R; ++u) { const V x = p[fft_swz(j + u * L)]; v[u] = {x.x, x.y}; }
        {
            const V w1v = tw[j];
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
        for (int q = 0; q < R; ++q) herm_emit<T, PASS>(a, outp, field, t0 + t, j + q * L, v[q], sgn);
    }
}

// grid = (ceil(ntrans/cw), fields) ; dynamic smem = cw*(N+pad)*sizeof(complex)
template <typename T, int PASS>
__global__ void __launch_bounds__(kFftMaxThreads, 2)
k_fft_h(const FftHArgs a)
{
    using V = typename Vec2<T>::type;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    V* buf = reinterpret_cast<V*>(fft_smem);
    const int N = a.N;
    const int cw = min(a.cw, a.ntrans - (int)blockIdx.x * a.cw);
    const int t0 = blockIdx.x * a.cw + (PASS == 0 ? a.row0 : 0);
    const bool second = (int)blockIdx.y >= a.nfields_first;
    const size_t field = second ? blockIdx.y - a.nfields_first : blockIdx.y;
    const size_t in_stride = PASS == 0 ? (size_t)a.nly * a.nlx : (size_t)a.nrow * a.nx;
    const V* src = reinterpret_cast<const V*>(second ? a.in2 : a.in) + field * in_stride;
    void* outp = second ? a.out2 : a.out;
    const T sgn = a.conj_io ? (T)-1 : (T)1;
    const V* tw = reinterpret_cast<const V*>(a.twiddle);

    switch (a.radix[0]) {
        case 2: herm_first_stage<T, PASS, 2>(a, buf, src, outp, field, t0, cw, sgn); break;
        case 3: herm_first_stage<T, PASS, 3>(a, buf, src, outp, field, t0, cw, sgn); break;
        case 4: herm_first_stage<T, PASS, 4>(a, buf, src, outp, field, t0, cw, sgn); break;
        case 5: herm_first_stage<T, PASS, 5>(a, buf, src, outp, field, t0, cw, sgn); break;
        default: herm_first_stage<T, PASS, 8>(a, buf, src, outp, field, t0, cw, sgn); break;
    }
    if (a.nstages == 1) return;
    __syncthreads();
    V* other = buf + (size_t)a.cw * fft_stride(N);     // second half, only present for generic radices
    int L = a.radix[0];
    for (int s = 1; s < a.nstages - 1; ++s) {
        const int r = a.radix[s];
        if (fft_is_generic(r)) {
            fft_stage_generic<T>(buf, other, tw, N, L, r, cw);
            V* tmp = buf; buf = other; other = tmp;
        } else {
            switch (r) {
                case 2: fft_stage<T, 2>(buf, tw, N, L, a.lshift[s], cw, 0, N); break;
                case 3: fft_stage<T, 3>(buf, tw, N, L, a.lshift[s], cw, 0, N); break;
                case 4: fft_stage<T, 4>(buf, tw, N, L, a.lshift[s], cw, 0, N); break;
                case 5: fft_stage<T, 5>(buf, tw, N, L, a.lshift[s], cw, 0, N); break;
                default: fft_stage<T, 8>(buf, tw, N, L, a.lshift[s], cw, 0, N); break;
            }
        }
        L *= r;
        __syncthreads();
    }
    switch (a.radix[a.nstages - 1]) {
        case 2: herm_last_stage<T, PASS, 2>(a, buf, tw, outp, field, t0, cw, sgn); break;
        case 3: herm_last_stage<T, PASS, 3>(a, buf, tw, outp, field, t0, cw, sgn); break;
        case 4: herm_last_stage<T, PASS, 4>(a, buf, tw, outp, field, t0, cw, sgn); break;
        case 5: herm_last_stage<T, PASS, 5>(a, buf, tw, outp, field, t0, cw, sgn); break;
        default: herm_last_stage<T, PASS, 8>(a, buf, tw, outp, field, t0, cw, sgn); break;
    }
}

inline size_t herm_work_bytes(const bldfm_geometry& g, bool f32, int64_t nfields)
{
    const int64_t chunk = std::min<int64_t>(nfields, 16384);
    return (size_t)2 * (size_t)chunk * (size_t)(g.nly / 2 + 1) * (size_t)g.nx * (f32 ? sizeof(float2) : sizeof(double2));
}

// specialised passes for N = 3P (fft24.cuh)
inline int fft24_lq(int N, int nl, int n, int p);
inline int fft24_pick_cw(int lq, bool f32, size_t smem_optin, int want, int64_t ntrans_total, int num_sms);
template <typename T, int PASS>
inline cudaError_t fft24_launch_pass(cudaStream_t stream, size_t smem_optin, int lq, const FftHArgs& a, dim3 grid);
// two-stage variant for P = 256, 512 (fft48.cuh)
inline int fft48_lq(int N, int nl, int n, int p);
inline int fft48_pick_cw(int lq, bool f32, size_t smem_optin, int want, int64_t ntrans_total, int num_sms);
template <typename T, int PASS>
inline cudaError_t fft48_launch_pass(cudaStream_t stream, size_t smem_optin, int lq, const FftHArgs& a, dim3 grid);

// persistent, bulk-copy pipelined variant for launches that fill the GPU many times over (fft24p.cuh)
struct Fft24pLayout;
inline bool fft24p_usable(int lq, int pass, const FftHArgs& a, int nfields, bool f32, size_t smem_optin, int num_sms,
                          Fft24pLayout* lay, int* grid);
template <typename T, int PASS>
inline cudaError_t fft24p_launch_pass(cudaStream_t stream, size_t smem_optin, int lq, const FftHArgs& a, int nfields,
                                      const Fft24pLayout& lay, int grid);

template <typename T, int PASS>
inline bool herm_try_fft24p(cudaStream_t stream, size_t smem_optin, int lq, const FftHArgs& a, int nfields);

// launch one pass with the specialised kernel when its geometry allows (lq >= 0), else with k_fft_h
template <typename T, int PASS>
inline cudaError_t herm_launch_pass(cudaStream_t stream, size_t smem_optin, int lq, FftHArgs a, unsigned nfields_y,
                                    int64_t ntrans_total)
{
    const bool f32 = sizeof(T) == 4;
    if (lq >= 0 && a.tw24) {
        // many work items per resident CTA: the persistent kernel that prefetches its operands with bulk copies
        if (herm_try_fft24p<T, PASS>(stream, smem_optin, lq, a, (int)nfields_y)) return cudaGetLastError();
    }
    // lq >= 0 implies N = 3P with P = nlx = n_out = out_off: try the two-stage variant first
    // (measured: 5 % faster than fft24.cuh when the launch fills the GPU many times over, 2 % slower for the
    // ~500 transforms of a single solve -> used for large launches only; BLDFM_B200_FFT48 = 0 never, 2 always)
    const int mode48 = fft_env_int("BLDFM_B200_FFT48", 1);
    const bool want48 = mode48 == 2 || (mode48 == 1 && ntrans_total >= 4096);
    const int lq48 = (lq >= 0 && a.tw48 && want48) ? fft48_lq(a.N, a.n_out, a.n_out, a.out_off) : -1;
    if (lq48 >= 0) {
        a.cw = fft48_pick_cw(lq48, f32, smem_optin, PASS == 1 ? 4 : 2, ntrans_total, 148);
        return fft48_launch_pass<T, PASS>(stream, smem_optin, lq48, a, dim3((unsigned)((a.ntrans + a.cw - 1) / a.cw), nfields_y));
    }
    if (lq >= 0 && a.tw24) {
        a.cw = fft24_pick_cw(lq, f32, smem_optin, PASS == 1 ? 4 : 2, ntrans_total, 148);
        return fft24_launch_pass<T, PASS>(stream, smem_optin, lq, a, dim3((unsigned)((a.ntrans + a.cw - 1) / a.cw), nfields_y));
    }
    k_fft_h<T, PASS><<<dim3((unsigned)((a.ntrans + a.cw - 1) / a.cw), nfields_y), fft_pick_threads(a.N, a.cw, a.radix[0]),
                       fft_smem_bytes(a.N, a.cw, f32), stream>>>(a);
    return cudaGetLastError();
}

// both passes for the nfields spectra of spec_p (-> out_p) and spec_q (-> out_q): two launches
template <typename T>
inline cudaError_t herm_fft_launch(cudaStream_t stream, size_t smem_optin, const bldfm_geometry& g,
                                   bool forward_dir, const void* spec_p, const void* spec_q, int64_t nfields,
                                   void* work, void* out_p, void* out_q, const PrunedFftTables& tab, int* nlaunch,
                                   bool herm_spec = false)
{
    using V = typename Vec2<T>::type;
    const bool f32 = sizeof(T) == 4;
    std::vector<int> rx, ry;
    fft_factorize(g.nfx, rx);
    fft_factorize(g.nfy, ry);
    const int nrow = g.nly / 2 + 1;

    FftHArgs ax{};
    ax.N = g.nfx; ax.nstages = (int)rx.size();
    { FftPassArgs tmp{}; fft_set_stages(tmp, rx); for (int i = 0; i < tmp.nstages; ++i) { ax.radix[i] = tmp.radix[i]; ax.lshift[i] = tmp.lshift[i]; } }
    ax.cw = fft_pick_cw(g.nfx, f32, smem_optin, 4, (int64_t)nrow * 2 * nfields);
    ax.ntrans = nrow; ax.conj_io = forward_dir ? 0 : 1;
    ax.nlx = g.nlx; ax.nly = g.nly; ax.nrow = nrow; ax.nx = g.nx;
    ax.out_off = g.px; ax.n_out = g.nx;
    ax.twiddle = tab.tw_x; ax.rev = tab.rev_x; ax.tw24 = tab.t24_x; ax.tw48 = tab.t48_x;
    ax.hs = herm_spec ? 1 : 0;
    ax.tfast = fft_env_int("BLDFM_B200_FFT_TFAST", 0);
    const bool use24 = fft_env_int("BLDFM_B200_FFT24", 1) != 0;
    const int lqx = use24 ? fft24_lq(g.nfx, g.nlx, g.nx, g.px) : -1;
    const int lqy = use24 ? fft24_lq(g.nfy, g.nly, g.ny, g.py) : -1;

    FftHArgs ay = ax;
    ay.N = g.nfy; ay.nstages = (int)ry.size();
    { FftPassArgs tmp{}; fft_set_stages(tmp, ry); for (int i = 0; i < tmp.nstages; ++i) { ay.radix[i] = tmp.radix[i]; ay.lshift[i] = tmp.lshift[i]; } }
    ay.cw = fft_pick_cw(g.nfy, f32, smem_optin, 4, (int64_t)((g.nx + 1) / 2) * 2 * nfields);
    ay.ntrans = (g.nx + 1) / 2;
    ay.out_off = g.py; ay.n_out = g.ny;
    ay.twiddle = tab.tw_y; ay.rev = tab.rev_y; ay.tw24 = tab.t24_y; ay.tw48 = tab.t48_y;

    cudaError_t e;
    e = set_max_dyn_smem(k_fft_h<T, 0>, (int)smem_optin);
    if (e != cudaSuccess) return e;
    e = set_max_dyn_smem(k_fft_h<T, 1>, (int)smem_optin);
    if (e != cudaSuccess) return e;
    const int64_t chunk = 16384;
    for (int64_t f0 = 0; f0 < nfields; f0 += chunk) {
        const int nf = (int)std::min<int64_t>(chunk, nfields - f0);
        FftHArgs bx = ax, by = ay;
        bx.nfields_first = nf; by.nfields_first = nf;
        bx.in = reinterpret_cast<const V*>(spec_p) + (size_t)f0 * g.nly * g.nlx;
        bx.in2 = reinterpret_cast<const V*>(spec_q) + (size_t)f0 * g.nly * g.nlx;
        bx.out = work;
        bx.out2 = reinterpret_cast<V*>(work) + (size_t)nf * nrow * g.nx;
        by.in = bx.out; by.in2 = bx.out2;
        by.out = reinterpret_cast<T*>(out_p) + (size_t)f0 * g.ny * g.nx;
        by.out2 = reinterpret_cast<T*>(out_q) + (size_t)f0 * g.ny * g.nx;
        e = herm_launch_pass<T, 0>(stream, smem_optin, lqx, bx, (unsigned)(2 * nf), (int64_t)ax.ntrans * 2 * nf);
        if (e != cudaSuccess) return e;
        e = herm_launch_pass<T, 1>(stream, smem_optin, lqy, by, (unsigned)(2 * nf), (int64_t)ay.ntrans * 2 * nf);
        if (e != cudaSuccess) return e;
        *nlaunch += 2;
    }
    return cudaGetLastError();
}

// ---- ky-slab sharded real-output back-transform -------------------------------------
// The nly/2+1 rows of A are split in blocks of Rp = ceil((nly/2+1)/G) rows: rank r owns rows
// [r*Rp, min((r+1)*Rp, nly/2+1)).  Receiver layout per field: [G*Rp][nx/G] (row = global row of A).
inline int herm_shard_rows(const bldfm_geometry& g, int nranks) { return (g.nly / 2 + 1 + nranks - 1) / nranks; }

// stage 1: pass X over this rank's rows of A, output blocked by destination rank
inline cudaError_t herm_sharded_xpass(cudaStream_t stream, size_t smem_optin, const bldfm_geometry& g,
                                      bool forward_dir, int row0, int rows, int nranks, const void* spec_p,
                                      const void* spec_q, int nfields, void* send_p, void* send_q,
                                      void* const* peer_p, void* const* peer_q, const PrunedFftTables& tab,
                                      int* nlaunch)
{
    std::vector<int> rx;
    fft_factorize(g.nfx, rx);
    const int nxl = g.nx / nranks;
    const int Rp = herm_shard_rows(g, nranks);
    FftHArgs ax{};
    ax.N = g.nfx; ax.nstages = (int)rx.size();
    { FftPassArgs tmp{}; fft_set_stages(tmp, rx); for (int i = 0; i < tmp.nstages; ++i) { ax.radix[i] = tmp.radix[i]; ax.lshift[i] = tmp.lshift[i]; } }
    ax.cw = fft_pick_cw(g.nfx, false, smem_optin, 4, (int64_t)rows * 2 * nfields);
    ax.ntrans = rows; ax.conj_io = forward_dir ? 0 : 1;
    ax.nlx = g.nlx; ax.nly = g.nly; ax.nrow = g.nly / 2 + 1; ax.nx = g.nx;
    ax.out_off = g.px; ax.n_out = g.nx;
    ax.twiddle = tab.tw_x; ax.rev = tab.rev_x; ax.tw24 = tab.t24_x; ax.tw48 = tab.t48_x;
    ax.row0 = row0; ax.out_block = nxl;
    ax.out_field_stride = (int64_t)nranks * Rp * nxl;
    ax.out_block_stride = (int64_t)Rp * nxl;
    ax.nfields_first = nfields;
    ax.in = spec_p; ax.in2 = spec_q; ax.out = send_p; ax.out2 = send_q;
    cudaError_t e = set_max_dyn_smem(k_fft_h<double, 0>, (int)smem_optin);
    if (e != cudaSuccess) return e;
    ax.hs = 1;                                             // the sharded march is always the half-plane one
    const int lqx = fft_env_int("BLDFM_B200_FFT24", 1) ? fft24_lq(g.nfx, g.nlx, g.nx, g.px) : -1;
    if (peer_p) {
        // the p fields then the q fields: each has its own table of peer pointers
        FftHArgs a1 = ax; a1.in2 = spec_p; a1.out_peer = peer_p;
        e = herm_launch_pass<double, 0>(stream, smem_optin, lqx, a1, (unsigned)nfields, (int64_t)rows * nfields);
        if (e != cudaSuccess) return e;
        FftHArgs a2 = ax; a2.in = spec_q; a2.in2 = spec_q; a2.out_peer = peer_q;
        e = herm_launch_pass<double, 0>(stream, smem_optin, lqx, a2, (unsigned)nfields, (int64_t)rows * nfields);
        if (e != cudaSuccess) return e;
        *nlaunch += 2;
    } else {
        e = herm_launch_pass<double, 0>(stream, smem_optin, lqx, ax, (unsigned)(2 * nfields), (int64_t)rows * 2 * nfields);
        if (e != cudaSuccess) return e;
        *nlaunch += 1;
    }
    return cudaGetLastError();
}

// stage 2: pass Y (column pairs) over the received [field][G*Rp][nx/G] -> real [field][ny][nx/G]
inline cudaError_t herm_sharded_ypass(cudaStream_t stream, size_t smem_optin, const bldfm_geometry& g,
                                      bool forward_dir, int nranks, const void* recv_p, const void* recv_q,
                                      int nfields, void* out_p, void* out_q, const PrunedFftTables& tab, int* nlaunch)
{
    std::vector<int> ry;
    fft_factorize(g.nfy, ry);
    const int nxl = g.nx / nranks;
    FftHArgs ay{};
    ay.N = g.nfy; ay.nstages = (int)ry.size();
    { FftPassArgs tmp{}; fft_set_stages(tmp, ry); for (int i = 0; i < tmp.nstages; ++i) { ay.radix[i] = tmp.radix[i]; ay.lshift[i] = tmp.lshift[i]; } }
    ay.ntrans = (nxl + 1) / 2;
    ay.cw = fft_pick_cw(g.nfy, false, smem_optin, 4, (int64_t)ay.ntrans * 2 * nfields);
    ay.conj_io = forward_dir ? 0 : 1;
    ay.nlx = g.nlx; ay.nly = g.nly; ay.nrow = nranks * herm_shard_rows(g, nranks); ay.nx = nxl;
    ay.out_off = g.py; ay.n_out = g.ny;
    ay.twiddle = tab.tw_y; ay.rev = tab.rev_y; ay.tw24 = tab.t24_y; ay.tw48 = tab.t48_y;
    ay.nfields_first = nfields;
    ay.in = recv_p; ay.in2 = recv_q; ay.out = out_p; ay.out2 = out_q;
    cudaError_t e = set_max_dyn_smem(k_fft_h<double, 1>, (int)smem_optin);
    if (e != cudaSuccess) return e;
    const int lqy = fft_env_int("BLDFM_B200_FFT24", 1) ? fft24_lq(g.nfy, g.nly, g.ny, g.py) : -1;
    e = herm_launch_pass<double, 1>(stream, smem_optin, lqy, ay, (unsigned)(2 * nfields), (int64_t)ay.ntrans * 2 * nfields);
    if (e != cudaSuccess) return e;
    *nlaunch += 1;
    return cudaGetLastError();
}

}  // namespace bldfm
