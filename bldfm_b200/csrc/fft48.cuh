// fft48.cuh -- two-stage variant of the default-halo back-transform passes for P = 256 and 512.
//
// Same decomposition as fft24.cuh one radix up: N = 3P = 48*Q (Q = P/16), n = n1 + Q*n2, k = 48*k1 + k2.
// Only n2 in {0..7, 40..47} is non-zero, so the radix-48 first stage is three radix-16 butterflies
// (k2 = r + 3q) of the sixteen inputs pre-rotated by the constants w_48^{n2 r}, fed from global memory.
// The 48 sequences have length Q = 32 (16), which ONE thread transforms in registers (radix-32 / 16) and
// stores straight to global memory: a single shared-memory exchange per transform instead of the two of
// fft24.cuh, whose cost is the LDS/STS/LDG path, not the FP64 pipe.
#pragma once

#include "fft24.cuh"

namespace bldfm {

constexpr int kFft48Threads = 384;

// cos(2*pi*m/96) for any integer m, folded at compile time (96 = lcm(24, 32, 48))
__host__ __device__ constexpr double fft96_quarter(int k)
{
    constexpr double t[25] = {
        1.0,
        0.9978589232386035067381,
        0.9914448613738104111446,
        0.9807852804032304491262,
        0.9659258262890682867497,
        0.9469301294951056642558,
        0.9238795325112867561282,
        0.8968727415326883038941,
        0.8660254037844386467637,
        0.8314696123025452370788,
        0.7933533402912351645798,
        0.7518398074789773964075,
        0.7071067811865475244008,
        0.6593458151000688684251,
        0.6087614290087206394161,
        0.5555702330196022247428,
        0.5,
        0.4422886902190012819952,
        0.3826834323650897717285,
        0.3214394653031615807011,
        0.2588190451025207623489,
        0.1950903220161282678483,
        0.1305261922200515915484,
        0.06540312923014306681532,
        0.0};
    return t[k];
}
__host__ __device__ constexpr double fft96_cos(int m)
{
    m = ((m % 96) + 96) % 96;
    return m <= 24 ? fft96_quarter(m) : m <= 48 ? -fft96_quarter(48 - m) : m <= 72 ? -fft96_quarter(m - 48)
                                                                                   : fft96_quarter(96 - m);
}
__host__ __device__ constexpr double fft96_sin(int m) { return fft96_cos(m - 24); }

// v * exp(-2*pi*i*M/D), D in {32, 48}
template <typename T, int M, int D>
__device__ __forceinline__ Cplx<T> fft96_rot(Cplx<T> v)
{
    constexpr int m = (((M * (96 / D)) % 96) + 96) % 96;
    if (m == 0) return v;
    if (m == 24) return {v.i, -v.r};
    if (m == 48) return {-v.r, -v.i};
    if (m == 72) return {-v.i, v.r};
    constexpr T c = (T)fft96_cos(m), s = (T)(-fft96_sin(m));
    return {xfma<T>(v.r, c, -(v.i * s)), xfma<T>(v.r, s, v.i * c)};
}

template <typename T, int K> struct Fft48Unroll {
    // v[u] *= w_48^{n2(u)*R}, n2(u) = u for u < 8, u - 16 for u >= 8   (u = K-1 down to 1)
    template <int R> static __device__ __forceinline__ void prerotate(Cplx<T>* v)
    {
        constexpr int u = K - 1;
        v[u] = fft96_rot<T, (u < 8 ? u : u - 16) * R, 48>(v[u]);
        Fft48Unroll<T, K - 1>::template prerotate<R>(v);
    }
    // o[k] *= w_32^k   (k = K-1 down to 1)
    static __device__ __forceinline__ void twiddle32(Cplx<T>* o)
    {
        o[K - 1] = fft96_rot<T, K - 1, 32>(o[K - 1]);
        Fft48Unroll<T, K - 1>::twiddle32(o);
    }
};
template <typename T> struct Fft48Unroll<T, 1> {
    template <int R> static __device__ __forceinline__ void prerotate(Cplx<T>*) {}
    static __device__ __forceinline__ void twiddle32(Cplx<T>*) {}
};

template <typename T> __device__ __forceinline__ void bfly32(Cplx<T>* v)
{
    Cplx<T> e[16], o[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
    bfly16<T>(e);
    bfly16<T>(o);
    Fft48Unroll<T, 16>::twiddle32(o);
#pragma unroll
    for (int k = 0; k < 16; ++k) { v[k] = cadd(e[k], o[k]); v[k + 16] = csub(e[k], o[k]); }
}

// grid = (ceil(ntrans/cw), fields) ; dynamic smem = cw*TS*sizeof(complex), TS = 48*(Q+1) + 8/cw
template <typename T, int PASS, int LQ>
__global__ void __launch_bounds__(kFft48Threads, 1)
k_fft48(const FftHArgs a)
{
    using V = typename Vec2<T>::type;
    constexpr int Q = 1 << LQ, LD = Q + 1;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    V* buf = reinterpret_cast<V*>(fft_smem);
    const int cw = min(a.cw, a.ntrans - (int)blockIdx.x * a.cw);
    const int lcw_full = 31 - __clz(a.cw);
    const int TS = 48 * LD + (8 >> lcw_full);
    // pass Y always walks the column pairs first; pass X does so only on request (a.tfast): measured 3-5 % slower
    // in the throughput regime (lanes of one row only 16 apart), a partial last CTA splits by division
    const int lcw = (cw == a.cw && (PASS == 1 || a.tfast)) ? lcw_full : -1;
    const int t0 = blockIdx.x * a.cw + (PASS == 0 ? a.row0 : 0);
    const bool second = (int)blockIdx.y >= a.nfields_first;
    const size_t field = second ? blockIdx.y - a.nfields_first : blockIdx.y;
    const size_t in_stride = PASS == 0 ? (size_t)a.nly * a.nlx : (size_t)a.nrow * a.nx;
    const V* src = reinterpret_cast<const V*>(second ? a.in2 : a.in) + field * in_stride;
    void* outp = second ? a.out2 : a.out;
    const T sgn = a.conj_io ? (T)-1 : (T)1;
    const V* tw = reinterpret_cast<const V*>(a.tw48);    // [48][Q]: w_N^{n1 k2}

    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();

    // ---- stage 1: sparse radix-48 from global memory, item = (n1, r) -> outputs k2 = r + 3q, q = 0..15
    for (int idx = threadIdx.x; idx < cw * 3 * Q; idx += (int)blockDim.x) {
        int t, it;
        fft24_split<PASS>(idx, lcw, 3 * Q, t, it);
        const int r = it >> LQ, n1 = it & (Q - 1);
        const int tg = t0 + t;
        Cplx<T> v[16], e = {(T)0, (T)0};
        if (PASS == 0 && a.hs && tg > 0 && tg < a.nly / 2) {
            // interior row of a conjugate-symmetric spectrum: H = S; f = n1 + Q*n2 sits at column n1 + Q*u
            const V* row = src + (size_t)tg * a.nlx + n1;
            const V xe = src[(size_t)(a.nly - tg) * a.nlx + 8 * Q];
#pragma unroll
            for (int u = 0; u < 16; ++u) { const V x = row[u * Q]; v[u] = {x.x, sgn * x.y}; }
            if (n1 == 0) {
                // Nyquist column: H[fy][-P/2] = S[fy][-P/2]/2 and H[fy][+P/2] = conj(S[-fy][-P/2])/2
                v[8] = {(T)0.5 * v[8].r, (T)0.5 * v[8].i};
                e = {(T)0.5 * xe.x, sgn * ((T)-0.5 * xe.y)};
            }
        } else if (PASS == 1 && (a.nx & 1) == 0) {
            // column pair (2tg, 2tg+1): f > 0 for u < 8 (rows n1 + Q*u of A), f < 0 for u >= 8 (rows
            // Q*(16-u) - n1 = (Q - n1) + Q*(15-u)); only f = 0 (n1 = 0, u = 0) packs differently
            const V* up = src + (size_t)n1 * a.nx + 2 * tg;
            const V* dn = src + (size_t)(Q - n1) * a.nx + 2 * tg;
            const size_t step = (size_t)Q * a.nx;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                V x1, x2;
                ld_pair(up + u * step, x1, x2);                                              // one 256-bit load
                v[u] = {x1.x - x2.y, sgn * (x1.y + x2.x)};                                   // A1 + i*A2
            }
#pragma unroll
            for (int u = 8; u < 16; ++u) {
                V x1, x2;
                ld_pair(dn + (15 - u) * step, x1, x2);
                v[u] = {x1.x + x2.y, sgn * (x2.x - x1.y)};                                   // conj(A1) + i*conj(A2)
            }
            if (n1 == 0) {
                V x1, x2, y1, y2;
                ld_pair(src + 2 * tg, x1, x2);
                v[0] = {x1.x, sgn * x2.x};                                                   // A[0] is real
                ld_pair(src + (size_t)(8 * Q) * a.nx + 2 * tg, y1, y2);
                e = {y1.x - y2.y, sgn * (y1.y + y2.x)};
            }
        } else {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int f = n1 + Q * (u < 8 ? u : u - 16);
                const Fft24Raw<T> raw = PASS == 0 ? fft24_fetch_x<T>(a, src, tg, f) : fft24_fetch_y<T>(a, src, 2 * tg, f);
                v[u] = PASS == 0 ? fft24_combine_x<T>(raw) : fft24_combine_y<T>(raw);
                v[u].i *= sgn;
            }
            if (n1 == 0) {
                const Fft24Raw<T> raw = PASS == 0 ? fft24_fetch_x<T>(a, src, tg, 8 * Q) : fft24_fetch_y<T>(a, src, 2 * tg, 8 * Q);
                e = PASS == 0 ? fft24_combine_x<T>(raw) : fft24_combine_y<T>(raw);
                e.i *= sgn;
            }
        }
        if (r == 1) Fft48Unroll<T, 16>::template prerotate<1>(v);
        else if (r == 2) Fft48Unroll<T, 16>::template prerotate<2>(v);
        bfly16<T>(v);
        if (n1 == 0) {
            // the input f = +P/2 (n2 = 8): w_48^{8(r+3q)} = w_6^r (-1)^q
            if (r == 1) e = fft96_rot<T, 8, 48>(e);
            else if (r == 2) e = fft96_rot<T, 16, 48>(e);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = (q & 1) ? csub(v[q], e) : cadd(v[q], e);
        }
        V* p = buf + (t * TS + n1);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int k2 = r + 3 * q;
            const V wv = tw[k2 * Q + n1];                   // w_N^{n1 k2}; lanes walk n1: contiguous
            const Cplx<T> y = cmul<T>(v[q], {wv.x, wv.y});
            p[k2 * LD] = mk2<T>(y.r, y.i);
        }
    }
    __syncthreads();

    // ---- stage 2: one thread transforms one sequence (radix Q in registers): output c of sequence k2 is
    // X[48 c + k2]; the window goes straight to global memory
    for (int idx = threadIdx.x; idx < cw * 48; idx += (int)blockDim.x) {
        int t, k2;
        fft24_split<PASS>(idx, lcw, 48, t, k2);
        const V* p = buf + (t * TS + k2 * LD);
        Cplx<T> v[Q];
#pragma unroll
        for (int u = 0; u < Q; ++u) { const V x = p[u]; v[u] = {x.x, x.y}; }
        if (Q == 32) bfly32<T>(v);
        else bfly16<T>(v);
        const int tg = t0 + t;
        const int o0 = k2 - a.out_off;                      // output c lands at o0 + 48*c of the window
        if (PASS == 0 && a.out_block == 0) {
            V* dst = reinterpret_cast<V*>(outp) + (field * a.nrow + tg) * (size_t)a.nx;
#pragma unroll
            for (int c = 0; c < Q; ++c) {
                const int o = o0 + 48 * c;
                if ((unsigned)o < (unsigned)a.n_out) dst[o] = mk2<T>(v[c].r, sgn * v[c].i);
            }
        } else if (PASS == 1 && (a.nx & 1) == 0) {
            V* dst = reinterpret_cast<V*>(reinterpret_cast<T*>(outp) + field * a.n_out * (size_t)a.nx + 2 * tg);
            const size_t pitch = (size_t)(a.nx >> 1);
#pragma unroll
            for (int c = 0; c < Q; ++c) {
                const int o = o0 + 48 * c;
                if ((unsigned)o < (unsigned)a.n_out) dst[o * pitch] = mk2<T>(v[c].r, sgn * v[c].i);
            }
        } else {
#pragma unroll
            for (int c = 0; c < Q; ++c) herm_emit<T, PASS>(a, outp, field, tg, 48 * c + k2, v[c], sgn);
        }
    }
}

// log2 Q of the two-stage plan for this pass, or -1
inline int fft48_lq(int N, int nl, int n, int p)
{
    if (N != 3 * nl || nl != n || p != n) return -1;
    if (nl == 512) return 5;
    if (nl == 256) return 4;
    return -1;
}

// host: stage-1 twiddle table [48][Q] w_N^{n1 k2}, interleaved (re, im) doubles
inline void fft48_tables(int lq, std::vector<double>& out)
{
    const int Q = 1 << lq, N = 48 * Q;
    std::vector<double> w;
    fft_twiddles(N, w);
    out.assign((size_t)2 * 48 * Q, 0.0);
    for (int k2 = 0; k2 < 48; ++k2)
        for (int n1 = 0; n1 < Q; ++n1) {
            const size_t k = (size_t)n1 * k2;               // < N
            out[2 * ((size_t)k2 * Q + n1)] = w[2 * k];
            out[2 * ((size_t)k2 * Q + n1) + 1] = w[2 * k + 1];
        }
}

inline size_t fft48_smem_bytes(int lq, int cw, bool f32)
{
    return (size_t)cw * (48 * ((1 << lq) + 1) + 8 / cw) * (f32 ? sizeof(float2) : sizeof(double2));
}

inline int fft48_pick_cw(int lq, bool f32, size_t smem_optin, int want, int64_t ntrans_total, int num_sms = 148)
{
    int cw = want;
    const int forced = fft_env_int(want == 4 ? "BLDFM_FFT24_CW_Y" : "BLDFM_FFT24_CW_X", 0);   // tuning sweeps
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) cw = forced;
    while (cw > 1 && fft48_smem_bytes(lq, cw, f32) > smem_optin) cw >>= 1;
    while (cw > 1 && ntrans_total / cw < 2 * (int64_t)num_sms) cw >>= 1;
    return cw;
}

template <typename T, int PASS>
inline cudaError_t fft48_launch_pass(cudaStream_t stream, size_t smem_optin, int lq, const FftHArgs& a, dim3 grid)
{
    const bool f32 = sizeof(T) == 4;
    const bool pdl = fft_env_int("BLDFM_B200_PDL", 1) != 0;
    const size_t sm = fft48_smem_bytes(lq, a.cw, f32);
    const int items = a.cw * 3 * (1 << lq);
    const int tmax = std::min(kFft48Threads, std::max(32, fft_env_int(PASS == 1 ? "BLDFM_FFT24_THREADS_Y" : "BLDFM_FFT24_THREADS_X", kFft48Threads) / 32 * 32));
    const int threads = std::min(tmax, std::max(32, (items + 31) / 32 * 32));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = sm; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e;
    if (lq == 5) {
        e = set_max_dyn_smem(k_fft48<T, PASS, 5>, (int)smem_optin);
        if (e != cudaSuccess) return e;
        e = cudaLaunchKernelEx(&cfg, k_fft48<T, PASS, 5>, a);
    } else if (lq == 4) {
        e = set_max_dyn_smem(k_fft48<T, PASS, 4>, (int)smem_optin);
        if (e != cudaSuccess) return e;
        e = cudaLaunchKernelEx(&cfg, k_fft48<T, PASS, 4>, a);
    } else {
        return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace bldfm
