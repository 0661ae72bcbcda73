// fft24.cuh -- real-output back-transform passes specialised for the default-halo geometry.
//
// With halo = max(domain) (src/bldfm/solver.py:108-109) and modes == grid size, every 1-D transform of
// the back-transform (solver.py:265-290) has length N = 3P with P = nlx = nx = px: P (+1) non-zero
// inputs, the frequencies |f| <= P/2, and P kept outputs, the window [P, 2P).  BASELINE configs 2-5
// are all of this shape (N = 1536, 3072, 12288).  Writing N = 24*Q (Q = P/8) and n = n1 + Q*n2,
// k = 24*k1 + k2:
//
//   X[24 k1 + k2] = sum_{n1} w_Q^{n1 k1} * [ w_N^{n1 k2} * sum_{n2} x[n1 + Q n2] w_24^{n2 k2} ]
//
// Only n2 in {0..3, 20..23} is non-zero, so the radix-24 first stage is three radix-8 butterflies
// (k2 = r + 3q, r = 0,1,2) of the eight inputs pre-rotated by the CONSTANTS w_24^{n2 r}: no zero is ever
// loaded, multiplied or stored, and the stage is fed straight from global memory.  The 24 length-Q
// sequences are then transformed in place in shared memory by 1-2 more radix-8/16 DIF stages (two
// exchanges in total for N = 1536 and 3072, where fft_herm.cuh needs four or five) with every index
// computed from compile-time constants; the last stage stores the window straight from registers.
// Operand loads (Hermitian combine / column-pair packing) and stores are those of fft_herm.cuh.
#pragma once

#include "fft_herm.cuh"

namespace bldfm {

constexpr int kFft24Threads = 384;

// exp(-2*pi*i*m/24) for any integer m, folded at compile time
__host__ __device__ constexpr double fft24_quarter(int k)   // cos(2*pi*k/24), k = 0..6
{
    return k == 0 ? 1.0 : k == 1 ? 0.96592582628906828675 : k == 2 ? 0.86602540378443864676
         : k == 3 ? 0.70710678118654752440 : k == 4 ? 0.5 : k == 5 ? 0.25881904510252076235 : 0.0;
}
__host__ __device__ constexpr double fft24_cos(int m)
{
    m = ((m % 24) + 24) % 24;
    return m <= 6 ? fft24_quarter(m) : m <= 12 ? -fft24_quarter(12 - m) : m <= 18 ? -fft24_quarter(m - 12)
                                                                                  : fft24_quarter(24 - m);
}
__host__ __device__ constexpr double fft24_sin(int m) { return fft24_cos(m - 6); }

template <typename T, int M>
__device__ __forceinline__ Cplx<T> fft24_rot(Cplx<T> v)     // v * exp(-2*pi*i*M/24)
{
    constexpr int m = ((M % 24) + 24) % 24;
    if (m == 0) return v;
    if (m == 6) return {v.i, -v.r};
    if (m == 12) return {-v.r, -v.i};
    if (m == 18) return {-v.i, v.r};
    constexpr T c = (T)fft24_cos(m), s = (T)(-fft24_sin(m));
    return {xfma<T>(v.r, c, -(v.i * s)), xfma<T>(v.r, s, v.i * c)};
}

// v[u] *= w_24^{n2(u)*R}, n2(u) = u for u < 4, u - 8 for u >= 4
template <typename T, int R>
__device__ __forceinline__ void fft24_prerotate(Cplx<T>* v)
{
    v[1] = fft24_rot<T, 1 * R>(v[1]); v[2] = fft24_rot<T, 2 * R>(v[2]); v[3] = fft24_rot<T, 3 * R>(v[3]);
    v[4] = fft24_rot<T, -4 * R>(v[4]); v[5] = fft24_rot<T, -3 * R>(v[5]); v[6] = fft24_rot<T, -2 * R>(v[6]);
    v[7] = fft24_rot<T, -1 * R>(v[7]);
}

template <typename T> __device__ __forceinline__ void bfly16(Cplx<T>* v)
{
    const T h = (T)0.70710678118654752440, c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173;
    Cplx<T> e[8] = {v[0], v[2], v[4], v[6], v[8], v[10], v[12], v[14]};
    Cplx<T> o[8] = {v[1], v[3], v[5], v[7], v[9], v[11], v[13], v[15]};
    bfly8<T>(e);
    bfly8<T>(o);
    // o[k] *= w_16^k = cos(k*pi/8) - i*sin(k*pi/8)
    o[1] = cmul<T>(o[1], {c1, -s1});
    o[2] = {h * (o[2].r + o[2].i), h * (o[2].i - o[2].r)};
    o[3] = cmul<T>(o[3], {s1, -c1});
    o[4] = cmuli_neg(o[4]);
    o[5] = cmul<T>(o[5], {-s1, -c1});
    o[6] = {h * (o[6].i - o[6].r), -h * (o[6].r + o[6].i)};
    o[7] = cmul<T>(o[7], {-c1, -s1});
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = cadd(e[k], o[k]); v[k + 8] = csub(e[k], o[k]); }
}

template <typename T, int R> __device__ __forceinline__ void bfly_pow2(Cplx<T>* v)
{
    if (R == 2) bfly2<T>(v);
    else if (R == 4) bfly4<T>(v);
    else if (R == 8) bfly8<T>(v);
    else bfly16<T>(v);
}

// radices of the in-place stages over the 24 sequences of length Q = 2^LQ
template <int LQ> struct Fft24Plan;
template <> struct Fft24Plan<5> { static constexpr int n = 2, r0 = 8, r1 = 4, r2 = 1; };
template <> struct Fft24Plan<6> { static constexpr int n = 2, r0 = 8, r1 = 8, r2 = 1; };
template <> struct Fft24Plan<7> { static constexpr int n = 2, r0 = 16, r1 = 8, r2 = 1; };
template <> struct Fft24Plan<8> { static constexpr int n = 2, r0 = 16, r1 = 16, r2 = 1; };
template <> struct Fft24Plan<9> { static constexpr int n = 3, r0 = 8, r1 = 8, r2 = 8; };

// split a flat work index into (transform, item).  lcw >= 0: the TRANSFORM index is fastest -- the 2^lcw transforms
// of a CTA walk the same items on neighbouring lanes (pass Y: the column pairs form contiguous 32 .. 64-byte pieces
// and share every twiddle load); lcw < 0: the items of one transform stay on consecutive lanes (pass X: contiguous
// global rows), any number of transforms
template <int PASS>
__device__ __forceinline__ void fft24_split(int idx, int lcw, int nitems, int& t, int& it)
{
    if (lcw >= 0) { it = idx >> lcw; t = idx & ((1 << lcw) - 1); }
    else { t = idx / nitems; it = idx - t * nitems; }       // partial last CTA: any number of transforms
}

// one in-place DIF stage of radix R over blocks of length M inside every sequence (not the last stage)
// `tw` is the stage's own table: tw[d*SUB + s] = w_M^{s d} (lanes walk s: contiguous, broadcast across sequences)
template <typename T, int PASS, int LQ, int M, int R>
__device__ __forceinline__ void fft24_stage(typename Vec2<T>::type* buf, const typename Vec2<T>::type* __restrict__ tw,
                                            int cw, int lcw, int TS)
{
    using V = typename Vec2<T>::type;
    constexpr int Q = 1 << LQ, LD = Q + 1, SUB = M / R, NIT = 24 * Q / R;
    for (int idx = threadIdx.x; idx < cw * NIT; idx += (int)blockDim.x) {
        int t, it;
        fft24_split<PASS>(idx, lcw, NIT, t, it);
        const int s = it % SUB;
        const int blk = it / SUB;                   // (sequence, block) pair: Q/M blocks per sequence
        const int k2 = blk / (Q / M), b = blk - k2 * (Q / M);
        V* p = buf + (t * TS + k2 * LD + b * M + s);
        Cplx<T> v[R];
#pragma unroll
        for (int u = 0; u < R; ++u) { const V x = p[u * SUB]; v[u] = {x.x, x.y}; }
        bfly_pow2<T, R>(v);
        p[0] = mk2<T>(v[0].r, v[0].i);
        // twiddles w_M^{s d}, d = 1 .. R-1: the powers of two come from the table, the others are ONE product of
        // two of them (w^d = w^{hi(d)} * w^{d - hi(d)}, at most two roundings beyond the table's).  These
        // passes are bound by the L1 / shared-memory data path, not by the FP64 pipe (ncu: l1tex 75-86 %,
        // FP64 30-40 %): 4 complex multiplies replace 4 of 7 loads for R = 8, 11 replace 11 of 15 for R = 16.
        Cplx<T> w[R];
#pragma unroll
        for (int d = 1; d < R; d <<= 1) { const V wv = tw[d * SUB + s]; w[d] = {wv.x, wv.y}; }
#pragma unroll
        for (int d = 3; d < R; ++d) {
            const int hi = d >= 8 ? 8 : d >= 4 ? 4 : 2;
            if (d != hi) w[d] = cmul<T>(w[hi], w[d - hi]);
        }
#pragma unroll
        for (int d = 1; d < R; ++d) {
            const Cplx<T> y = cmul<T>(v[d], w[d]);
            p[d * SUB] = mk2<T>(y.r, y.i);
        }
    }
}

// grid = (ceil(ntrans/cw), fields) ; dynamic smem = cw*TS*sizeof(complex), TS = 24*(Q+1) + 8/cw
template <typename T, int PASS, int LQ>
__global__ void __launch_bounds__(kFft24Threads, (LQ <= 6 ? 2 : 1))
k_fft24(const FftHArgs a)
{
    using V = typename Vec2<T>::type;
    using PL = Fft24Plan<LQ>;
    constexpr int Q = 1 << LQ, LD = Q + 1;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    V* buf = reinterpret_cast<V*>(fft_smem);
    const int cw = min(a.cw, a.ntrans - (int)blockIdx.x * a.cw);
    const int lcw_full = 31 - __clz(a.cw);
    const int TS = 24 * LD + (8 >> lcw_full);
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
    const V* tw = reinterpret_cast<const V*>(a.tw24);    // [24][Q] | [Q] | [Q/r0]

    // programmatic dependent launch: this grid may have been started while its producer (the march, or
    // pass X) was still draining; let the next pass do the same, then wait for the producer's results
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();

    // ---- stage 1: sparse radix-24 from global memory, item = (n1, r) -> outputs k2 = r + 3q
    for (int idx = threadIdx.x; idx < cw * 3 * Q; idx += (int)blockDim.x) {
        int t, it;
        fft24_split<PASS>(idx, lcw, 3 * Q, t, it);
        const int r = it >> LQ, n1 = it & (Q - 1);
        const int tg = t0 + t;
        Cplx<T> v[8], e = {(T)0, (T)0};
        if (PASS == 0 && a.hs && tg > 0 && tg < a.nly / 2) {
            // interior row of a conjugate-symmetric spectrum: H = S, and f = n1 + Q*n2 sits at column
            // n1 + Q*u.  The Nyquist column (items n1 = 0) has no partner in the retained set:
            // H[fy][-P/2] = S[fy][-P/2]/2 and H[fy][+P/2] = conj(S[-fy][-P/2])/2.
            const V* row = src + (size_t)tg * a.nlx + n1;
            const V xe = src[(size_t)(a.nly - tg) * a.nlx + 4 * Q];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const V x = row[u * Q]; v[u] = {x.x, sgn * x.y}; }
            if (n1 == 0) {
                v[4] = {(T)0.5 * v[4].r, (T)0.5 * v[4].i};
                e = {(T)0.5 * xe.x, sgn * ((T)-0.5 * xe.y)};
            }
        } else if (PASS == 1 && (a.nx & 1) == 0) {
            // column pair (2tg, 2tg+1): f > 0 for u < 4 (rows n1 + Q*u of A), f < 0 for u >= 4 (rows
            // Q*(8-u) - n1): the packing signs are compile-time; only f = 0 (n1 = 0, u = 0) differs
            const V* up = src + (size_t)n1 * a.nx + 2 * tg;
            const V* dn = src + (size_t)(Q - n1) * a.nx + 2 * tg;          // row Q*(8-u) - n1 = (Q - n1) + Q*(7-u)
            const size_t step = (size_t)Q * a.nx;
            V x1[9], x2[9];
#pragma unroll
            for (int u = 0; u < 4; ++u) ld_pair(up + u * step, x1[u], x2[u]);
#pragma unroll
            for (int u = 4; u < 8; ++u) ld_pair(dn + (7 - u) * step, x1[u], x2[u]);
            ld_pair(src + (size_t)(n1 == 0 ? 4 * Q : n1) * a.nx + 2 * tg, x1[8], x2[8]);
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = {x1[u].x - x2[u].y, sgn * (x1[u].y + x2[u].x)};        // A1 + i*A2
#pragma unroll
            for (int u = 4; u < 8; ++u) v[u] = {x1[u].x + x2[u].y, sgn * (x2[u].x - x1[u].y)};        // conj(A1) + i*conj(A2)
            if (n1 == 0) {
                v[0] = {x1[0].x, sgn * x2[0].x};                                                     // A[0] is real
                e = {x1[8].x - x2[8].y, sgn * (x1[8].y + x2[8].x)};
            }
        } else {
            Fft24Raw<T> raw[9];
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                // u = 8: the one input outside the eight, f = +P/2 (n2 = 4), which only the items n1 = 0 hold
                const int f = u < 8 ? n1 + Q * (u < 4 ? u : u - 8) : (n1 == 0 ? 4 * Q : n1);
                raw[u] = PASS == 0 ? fft24_fetch_x<T>(a, src, tg, f) : fft24_fetch_y<T>(a, src, 2 * tg, f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                v[u] = PASS == 0 ? fft24_combine_x<T>(raw[u]) : fft24_combine_y<T>(raw[u]);
                v[u].i *= sgn;
            }
            e = PASS == 0 ? fft24_combine_x<T>(raw[8]) : fft24_combine_y<T>(raw[8]);
            e.i *= sgn;
        }
        if (r == 1) fft24_prerotate<T, 1>(v);
        else if (r == 2) fft24_prerotate<T, 2>(v);
        bfly8<T>(v);
        if (n1 == 0) {
            // w_24^{4(r+3q)} = w_6^r (-1)^q
            if (r == 1) e = fft24_rot<T, 4>(e);
            else if (r == 2) e = fft24_rot<T, 8>(e);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = (q & 1) ? csub(v[q], e) : cadd(v[q], e);
        }
        V* p = buf + (t * TS + n1);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k2 = r + 3 * q;
            const V wv = tw[k2 * Q + n1];                   // w_N^{n1 k2}; lanes walk n1: contiguous
            const Cplx<T> y = cmul<T>(v[q], {wv.x, wv.y});
            p[k2 * LD] = mk2<T>(y.r, y.i);
        }
    }
    __syncthreads();

    // ---- in-place stages over the 24 sequences
    fft24_stage<T, PASS, LQ, Q, PL::r0>(buf, tw + 24 * Q, cw, lcw, TS);
    __syncthreads();
    if (PL::n == 3) {
        fft24_stage<T, PASS, LQ, Q / PL::r0, PL::r1>(buf, tw + 25 * Q, cw, lcw, TS);
        __syncthreads();
    }

    // ---- last stage: radix RL over contiguous blocks, outputs go straight to global memory.
    // Block b of sequence k2 holds the sub-transform with k1 = rev(b) (mod Q/RL); output c adds (Q/RL)*c.
    constexpr int RL = PL::n == 3 ? PL::r2 : PL::r1;
    constexpr int NB = Q / RL;                              // blocks per sequence
    for (int idx = threadIdx.x; idx < cw * 24 * NB; idx += (int)blockDim.x) {
        int t, it;
        fft24_split<PASS>(idx, lcw, 24 * NB, t, it);
        const int b = it / 24, k2 = it - b * 24;            // k2 fastest: consecutive outputs k
        const V* p = buf + (t * TS + k2 * LD + b * RL);
        Cplx<T> v[RL];
#pragma unroll
        for (int u = 0; u < RL; ++u) { const V x = p[u]; v[u] = {x.x, x.y}; }
        bfly_pow2<T, RL>(v);
        // b = d1*(NB/r0) + d2 (three stages) or b = d1 (two stages);  k1 = d1 + r0*d2 + (Q/RL)*c
        const int k1lo = PL::n == 3 ? (b / (NB / PL::r0)) + PL::r0 * (b % (NB / PL::r0)) : b;
        const int tg = t0 + t;
        const int o0 = 24 * k1lo + k2 - a.out_off;          // output c lands at o0 + 24*NB*c of the window
        if (PASS == 0 && a.out_block == 0) {
            V* dst = reinterpret_cast<V*>(outp) + (field * a.nrow + tg) * (size_t)a.nx;
#pragma unroll
            for (int c = 0; c < RL; ++c) {
                const int o = o0 + 24 * NB * c;
                if ((unsigned)o < (unsigned)a.n_out) dst[o] = mk2<T>(v[c].r, sgn * v[c].i);
            }
        } else if (PASS == 1 && (a.nx & 1) == 0) {
            // two adjacent real outputs per store; nx even keeps every pair 2*sizeof(T)-aligned
            V* dst = reinterpret_cast<V*>(reinterpret_cast<T*>(outp) + field * a.n_out * (size_t)a.nx + 2 * tg);
            const size_t pitch = (size_t)(a.nx >> 1);
#pragma unroll
            for (int c = 0; c < RL; ++c) {
                const int o = o0 + 24 * NB * c;
                if ((unsigned)o < (unsigned)a.n_out) dst[o * pitch] = mk2<T>(v[c].r, sgn * v[c].i);
            }
        } else {
#pragma unroll
            for (int c = 0; c < RL; ++c)
                herm_emit<T, PASS>(a, outp, field, tg, 24 * (k1lo + NB * c) + k2, v[c], sgn);
        }
    }
}

// can this 1-D pass (length N, nl retained modes, n kept outputs after an offset p) use k_fft24 ?
inline int fft24_lq(int N, int nl, int n, int p)
{
    if (N != 3 * nl || nl != n || p != n) return -1;
    for (int lq = 5; lq <= 9; ++lq)
        if (nl == (8 << lq)) return lq;
    return -1;
}

// host: the kernel's twiddle tables, interleaved (re, im) doubles: [24][Q] w_N^{n1 k2} | [Q] w_Q^{s d} at
// d*(Q/r0)+s | [Q/r0] w_{Q/r0}^{s d} at d*(Q/r0/r1)+s (three-stage plans only)
inline void fft24_tables(int lq, std::vector<double>& out)
{
    const int Q = 1 << lq, N = 24 * Q;
    const int r0 = lq == 7 || lq == 8 ? 16 : 8, r1 = lq == 5 ? 4 : lq == 8 ? 16 : 8;
    std::vector<double> w;
    fft_twiddles(N, w);
    out.assign((size_t)2 * (24 * Q + Q + Q / r0), 0.0);
    auto put = [&](size_t pos, int64_t k) { out[2 * pos] = w[2 * (size_t)(k % N)]; out[2 * pos + 1] = w[2 * (size_t)(k % N) + 1]; };
    for (int k2 = 0; k2 < 24; ++k2)
        for (int n1 = 0; n1 < Q; ++n1) put((size_t)k2 * Q + n1, (int64_t)n1 * k2);
    const int sub0 = Q / r0;
    for (int d = 0; d < r0; ++d)
        for (int s = 0; s < sub0; ++s) put((size_t)24 * Q + d * sub0 + s, (int64_t)24 * s * d);
    if (lq == 9) {
        const int M = Q / r0, sub1 = M / r1;
        for (int d = 0; d < r1; ++d)
            for (int s = 0; s < sub1; ++s) put((size_t)25 * Q + d * sub1 + s, (int64_t)(N / M) * s * d);
    }
}

inline size_t fft24_smem_bytes(int lq, int cw, bool f32)
{
    return (size_t)cw * (24 * ((1 << lq) + 1) + 8 / cw) * (f32 ? sizeof(float2) : sizeof(double2));
}

// transforms per CTA: as many as fit (<= want), fewer when the launch would not fill the SMs
inline int fft24_pick_cw(int lq, bool f32, size_t smem_optin, int want, int64_t ntrans_total, int num_sms = 148)
{
    int cw = want;
    const int forced = fft_env_int(want == 4 ? "BLDFM_FFT24_CW_Y" : "BLDFM_FFT24_CW_X", 0);   // tuning sweeps
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) cw = forced;
    const size_t ctas = lq <= 6 ? 2 : 1;                    // resident CTAs per SM the launch bounds aim at
    while (cw > 1 && fft24_smem_bytes(lq, cw, f32) * ctas > smem_optin) cw >>= 1;
    if (forced > 0) return cw;
    // small launches: keep enough CTAs to cover the SMs -- twice over for pass X; pass Y (want == 4) gains
    // more from two column pairs per CTA (full 32-byte store sectors, 64-byte loads) than from the extra CTAs
    // (measured at config 2, 512 transforms: cw = 2 -> 19.2 us for both passes, cw = 1 -> 23.4 us)
    const int64_t min_ctas = want == 4 ? (3 * (int64_t)num_sms) / 2 : 2 * (int64_t)num_sms;
    while (cw > 1 && ntrans_total / cw < min_ctas) cw >>= 1;
    return cw;
}

template <typename T, int PASS>
inline cudaError_t fft24_launch_pass(cudaStream_t stream, size_t smem_optin, int lq, const FftHArgs& a, dim3 grid)
{
    const bool f32 = sizeof(T) == 4;
    // overlap this grid's launch with the tail of the kernel before it in the stream (see k_fft24)
    const bool pdl = fft_env_int("BLDFM_B200_PDL", 1) != 0;
    const size_t sm = fft24_smem_bytes(lq, a.cw, f32);
    const int items = a.cw * 3 * (1 << lq);
    // pass X of the radix-16 plans (150 registers) runs best as two resident CTAs of 192 threads (measured)
    const int tdef = (PASS == 0 && lq >= 7 && lq <= 8) ? 192 : kFft24Threads;
    const int tmax = std::min(kFft24Threads, std::max(64, fft_env_int(PASS == 1 ? "BLDFM_FFT24_THREADS_Y" : "BLDFM_FFT24_THREADS_X", tdef) / 32 * 32));
    const int threads = std::min(tmax, std::max(96, (items + 31) / 32 * 32));
#define BLDFM_FFT24_CASE(LQ)                                                                                  \
    case LQ: {                                                                                                \
        cudaError_t e = set_max_dyn_smem(k_fft24<T, PASS, LQ>, (int)smem_optin);                                                \
        if (e != cudaSuccess) return e;                                                                       \
        cudaLaunchConfig_t cfg = {};                                                                          \
        cfg.gridDim = grid; cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = sm; cfg.stream = stream; \
        cudaLaunchAttribute at[1];                                                                            \
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                        \
        at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;                                       \
        cfg.attrs = at; cfg.numAttrs = 1;                                                                     \
        e = cudaLaunchKernelEx(&cfg, k_fft24<T, PASS, LQ>, a);                                                \
        if (e != cudaSuccess) return e;                                                                       \
        break;                                                                                                \
    }
    switch (lq) {
        BLDFM_FFT24_CASE(5)
        BLDFM_FFT24_CASE(6)
        BLDFM_FFT24_CASE(7)
        BLDFM_FFT24_CASE(8)
        BLDFM_FFT24_CASE(9)
        default: return cudaErrorInvalidValue;
    }
#undef BLDFM_FFT24_CASE
    return cudaGetLastError();
}

}  // namespace bldfm
