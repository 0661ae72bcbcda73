// fft24p.cuh -- persistent, software-pipelined form of the default-halo back-transform passes (fft24.cuh)
// for launches that fill the GPU many times over (batched solves, all-levels outputs).
//
// Same mathematics, same tables and the same operation order per element as k_fft24 -- the results are
// bit-identical -- but the data movement is rebuilt around the Blackwell async-copy path:
//
//   * one CTA per SM slot, looping over its share of (field, transform group) work items;
//   * the operands of work item i+1 are brought into shared memory with bulk async copies
//     (cp.async.bulk, completion counted on an mbarrier) while item i is in its butterflies: the global
//     load latency that k_fft24 exposes at the start of every CTA (top stall: long_scoreboard) leaves the
//     critical path, and no register is held for a load in flight;
//       pass X  a work item is `cw` rows of the half-plane spectrum: ONE contiguous bulk copy per row
//               (nlx complex = 8 KB at nlx = 512); for a conjugate-symmetric spectrum that row IS the
//               operand set of the transform (H = S), only the Nyquist column needs one extra element;
//       pass Y  a work item is `cw` column pairs of the intermediate A: one bulk copy of 2*cw complex per
//               row of A, issued by as many threads as there are rows;
//   * the twiddle tables ([24][Q] stage-1 table, in-place stage tables) are staged in shared memory once
//     per CTA instead of being re-read through L1 by every transform.
//
// Rows that need the general Hermitian combine (fy = 0, the Nyquist row, or a spectrum that is not
// conjugate-symmetric) take their operands straight from global memory like k_fft24 does.
#pragma once

#include "fft24.cuh"

namespace bldfm {

// ---- launch geometry ---------------------------------------------------------------------------------
// transforms per work item: 2 rows (pass X) / 2 column pairs (pass Y) up to N = 3072; one beyond
__host__ __device__ constexpr int fft24p_cw(int lq) { return lq <= 7 ? 2 : 1; }

template <typename T> __host__ __device__ constexpr size_t fft24p_csize() { return 2 * sizeof(T); }

struct Fft24pLayout {
    size_t off_stage[2], off_buf, off_tw, off_bar, total;
    int stage_elems;   // complex elements of one operand buffer
    int tw_elems;      // complex elements of the twiddle tables kept in shared memory (0: read through L1)
};

// shared-memory map: [operand buffer 0][operand buffer 1][FFT work buffer][twiddle tables][2 mbarriers]
inline Fft24pLayout fft24p_layout(int lq, int pass, int nlx, int nrow, bool f32, size_t smem_optin, int ctas_per_sm)
{
    const int Q = 1 << lq, cw = fft24p_cw(lq);
    const size_t cs = f32 ? sizeof(float2) : sizeof(double2);
    Fft24pLayout L{};
    // pass X: cw rows of nlx elements + one partner element per row (padded to 2 elements per row for alignment)
    // pass Y: nrow rows of 2*cw elements
    L.stage_elems = pass == 0 ? cw * (nlx + 2) : nrow * 2 * cw;
    const size_t stage_b = ((size_t)L.stage_elems * cs + 127) / 128 * 128;
    const size_t buf_b = ((size_t)cw * (24 * (Q + 1) + 8 / cw) * cs + 127) / 128 * 128;
    const int r0 = lq == 7 || lq == 8 ? 16 : 8;
    const int tw_all = 24 * Q + Q + (lq == 9 ? Q / r0 : 0);
    L.off_stage[0] = 0;
    L.off_stage[1] = stage_b;
    L.off_buf = 2 * stage_b;
    L.off_tw = L.off_buf + buf_b;
    const size_t budget = smem_optin / (size_t)ctas_per_sm - 1024;     // 1 KB per resident CTA is reserved
    const size_t with_tw = L.off_tw + (size_t)tw_all * cs + 64;
    L.tw_elems = with_tw <= budget ? tw_all : 0;
    L.off_bar = L.off_tw + (size_t)L.tw_elems * cs;
    L.total = L.off_bar + 64;
    return L;
}

struct Fft24pArgs {
    FftHArgs h;
    int32_t nwork;          // work items of the launch = groups per field * fields
    int32_t ngroups;        // transform groups per field
    int32_t nfields;        // fields of the launch (p fields then q fields, h.nfields_first of the former)
    int32_t stage_elems, tw_elems;
    uint32_t off_stage0, off_stage1, off_buf, off_tw, off_bar;
};

// is row tg of pass X an interior row of a conjugate-symmetric spectrum (operands = the row itself)?
__device__ __forceinline__ bool fft24p_interior(const FftHArgs& a, int tg) { return a.hs && tg > 0 && tg < a.nly / 2; }

// grid = min(nwork, resident CTA slots); block = 384 (192 for the radix-16 plans of pass X)
template <typename T, int PASS, int LQ>
__global__ void __launch_bounds__(kFft24Threads, (LQ <= 6 ? 2 : 1))
k_fft24p(const Fft24pArgs pa)
{
    using V = typename Vec2<T>::type;
    using PL = Fft24Plan<LQ>;
    constexpr int Q = 1 << LQ, LD = Q + 1;
    constexpr int CW = fft24p_cw(LQ), LCW = CW == 2 ? 1 : 0;
    constexpr int TS = 24 * LD + (8 >> LCW);
    const FftHArgs& a = pa.h;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    V* stage[2] = {reinterpret_cast<V*>(fft_smem + pa.off_stage0), reinterpret_cast<V*>(fft_smem + pa.off_stage1)};
    V* buf = reinterpret_cast<V*>(fft_smem + pa.off_buf);
    V* stw = reinterpret_cast<V*>(fft_smem + pa.off_tw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(fft_smem + pa.off_bar);
    const T sgn = a.conj_io ? (T)-1 : (T)1;
    const int tid = threadIdx.x, nthr = blockDim.x;

    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    // the twiddle tables do not depend on the producer kernel: stage them while it may still be draining
    const V* gtw = reinterpret_cast<const V*>(a.tw24);
    for (int i = tid; i < pa.tw_elems; i += nthr) stw[i] = gtw[i];
    const V* tw = pa.tw_elems ? stw : gtw;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();

    const size_t in_stride = PASS == 0 ? (size_t)a.nly * a.nlx : (size_t)a.nrow * a.nx;

    // where work item w lives
    auto locate = [&](int w, size_t& field, bool& second, int& t0, int& cw) {
        const int f = w / pa.ngroups, g = w - f * pa.ngroups;
        second = f >= a.nfields_first;
        field = second ? f - a.nfields_first : f;
        t0 = g * CW + (PASS == 0 ? a.row0 : 0);
        cw = min(CW, a.ntrans - g * CW);
    };

    // enqueue the bulk copies of work item w into operand buffer b (all threads call it)
    auto prefetch = [&](int w, int b) {
        size_t field; bool second; int t0, cw;
        locate(w, field, second, t0, cw);
        const V* src = reinterpret_cast<const V*>(second ? a.in2 : a.in) + field * in_stride;
        if (PASS == 0) {
            if (tid == 0) {
                uint32_t bytes = 0;
                for (int t = 0; t < cw; ++t)
                    if (fft24p_interior(a, t0 + t)) bytes += (uint32_t)(a.nlx * sizeof(V)) + 16u;
                fence_proxy_async();
                mbar_expect_tx(&bar[b], bytes);
                for (int t = 0; t < cw; ++t) {
                    const int tg = t0 + t;
                    if (!fft24p_interior(a, tg)) continue;
                    V* dst = stage[b] + (size_t)t * (a.nlx + 2);
                    bulk_g2s(dst, src + (size_t)tg * a.nlx, (uint32_t)(a.nlx * sizeof(V)), &bar[b]);
                    // Nyquist-column partner S[-fy][-P/2], needed by the items n1 = 0 (16 bytes: one complex128,
                    // or the complex64 pair starting there)
                    bulk_g2s(dst + a.nlx, src + (size_t)(a.nly - tg) * a.nlx + 4 * Q, 16u, &bar[b]);
                }
            }
        } else {
            // rows 0 .. nrow-1 of A, 2*cw consecutive columns each
            const uint32_t row_bytes = (uint32_t)(2 * cw * sizeof(V));
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(&bar[b], row_bytes * (uint32_t)(a.nly / 2 + 1));
            }
            __syncwarp();
            for (int r = tid; r <= a.nly / 2; r += nthr)
                bulk_g2s(stage[b] + (size_t)r * 2 * CW, src + (size_t)r * a.nx + 2 * t0, row_bytes, &bar[b]);
        }
    };

    int w = blockIdx.x;
    if (w < pa.nwork) prefetch(w, 0);
    uint32_t phase[2] = {0, 0};
    for (int it = 0; w < pa.nwork; w += gridDim.x, ++it) {
        const int b = it & 1;
        const int wn = w + (int)gridDim.x;
        if (wn < pa.nwork) prefetch(wn, b ^ 1);
        size_t field; bool second; int t0, cw;
        locate(w, field, second, t0, cw);
        const V* src = reinterpret_cast<const V*>(second ? a.in2 : a.in) + field * in_stride;
        void* outp = second ? a.out2 : a.out;
        mbar_wait(&bar[b], phase[b]);
        phase[b] ^= 1;
        const V* ops = stage[b];

        // ---- stage 1: sparse radix-24, item = (n1, r) -> outputs k2 = r + 3q (operands from shared memory)
        for (int idx = tid; idx < cw * 3 * Q; idx += nthr) {
            int t, itx;
            if (PASS == 1) { itx = idx >> LCW; t = idx & (CW - 1); if (cw < CW) { itx = idx; t = 0; } }
            else { t = idx / (3 * Q); itx = idx - t * 3 * Q; }
            const int r = itx >> LQ, n1 = itx & (Q - 1);
            const int tg = t0 + t;
            Cplx<T> v[8], e = {(T)0, (T)0};
            if (PASS == 0 && fft24p_interior(a, tg)) {
                const V* row = ops + (size_t)t * (a.nlx + 2) + n1;
#pragma unroll
                for (int u = 0; u < 8; ++u) { const V x = row[u * Q]; v[u] = {x.x, sgn * x.y}; }
                if (n1 == 0) {
                    const V xe = ops[(size_t)t * (a.nlx + 2) + a.nlx];
                    v[4] = {(T)0.5 * v[4].r, (T)0.5 * v[4].i};
                    e = {(T)0.5 * xe.x, sgn * ((T)-0.5 * xe.y)};
                }
            } else if (PASS == 1) {
                // column pair t of the staged rows: f > 0 for u < 4 (rows n1 + Q*u), f < 0 for u >= 4 (rows
                // Q*(8-u) - n1).  The pair's two elements are adjacent (32 bytes): half of each quarter-warp
                // reads them in swapped order so that one LDS.128 of eight lanes covers eight different
                // 16-byte bank groups.
                const bool sw = (tid >> 2) & 1;
                const int c0 = 2 * t + (sw ? 1 : 0), c1 = 2 * t + (sw ? 0 : 1);
                V x1[9], x2[9];
#pragma unroll
                for (int u = 0; u < 9; ++u) {
                    const int rowi = u < 4 ? n1 + Q * u : u < 8 ? (Q - n1) + Q * (7 - u) : (n1 == 0 ? 4 * Q : n1);
                    const V p0 = ops[(size_t)rowi * 2 * CW + c0];
                    const V p1 = ops[(size_t)rowi * 2 * CW + c1];
                    x1[u] = sw ? p1 : p0;
                    x2[u] = sw ? p0 : p1;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = {x1[u].x - x2[u].y, sgn * (x1[u].y + x2[u].x)};        // A1 + i*A2
#pragma unroll
                for (int u = 4; u < 8; ++u) v[u] = {x1[u].x + x2[u].y, sgn * (x2[u].x - x1[u].y)};        // conj(A1) + i*conj(A2)
                if (n1 == 0) {
                    v[0] = {x1[0].x, sgn * x2[0].x};                                                     // A[0] is real
                    e = {x1[8].x - x2[8].y, sgn * (x1[8].y + x2[8].x)};
                }
            } else {
                // edge rows of pass X: general Hermitian combine straight from global memory (as k_fft24)
                Fft24Raw<T> raw[9];
#pragma unroll
                for (int u = 0; u < 9; ++u) {
                    const int f = u < 8 ? n1 + Q * (u < 4 ? u : u - 8) : (n1 == 0 ? 4 * Q : n1);
                    raw[u] = fft24_fetch_x<T>(a, src, tg, f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) { v[u] = fft24_combine_x<T>(raw[u]); v[u].i *= sgn; }
                e = fft24_combine_x<T>(raw[8]);
                e.i *= sgn;
            }
            if (r == 1) fft24_prerotate<T, 1>(v);
            else if (r == 2) fft24_prerotate<T, 2>(v);
            bfly8<T>(v);
            if (n1 == 0) {
                if (r == 1) e = fft24_rot<T, 4>(e);
                else if (r == 2) e = fft24_rot<T, 8>(e);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = (q & 1) ? csub(v[q], e) : cadd(v[q], e);
            }
            V* p = buf + (t * TS + n1);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k2 = r + 3 * q;
                const V wv = tw[k2 * Q + n1];
                const Cplx<T> y = cmul<T>(v[q], {wv.x, wv.y});
                p[k2 * LD] = mk2<T>(y.r, y.i);
            }
        }
        __syncthreads();

        // ---- in-place stages over the 24 sequences (fft24.cuh)
        fft24_stage<T, PASS, LQ, Q, PL::r0>(buf, tw + 24 * Q, cw, cw == CW ? LCW : -1, TS);
        __syncthreads();
        if (PL::n == 3) {
            fft24_stage<T, PASS, LQ, Q / PL::r0, PL::r1>(buf, tw + 25 * Q, cw, cw == CW ? LCW : -1, TS);
            __syncthreads();
        }

        // ---- last stage: radix RL over contiguous blocks, outputs go straight to global memory
        constexpr int RL = PL::n == 3 ? PL::r2 : PL::r1;
        constexpr int NB = Q / RL;
        for (int idx = tid; idx < cw * 24 * NB; idx += nthr) {
            int t, itx;
            if (PASS == 1) { if (cw == CW) { itx = idx >> LCW; t = idx & (CW - 1); } else { itx = idx; t = 0; } }
            else { t = idx / (24 * NB); itx = idx - t * 24 * NB; }
            const int bb = itx / 24, k2 = itx - bb * 24;
            const V* p = buf + (t * TS + k2 * LD + bb * RL);
            Cplx<T> v[RL];
#pragma unroll
            for (int u = 0; u < RL; ++u) { const V x = p[u]; v[u] = {x.x, x.y}; }
            bfly_pow2<T, RL>(v);
            const int k1lo = PL::n == 3 ? (bb / (NB / PL::r0)) + PL::r0 * (bb % (NB / PL::r0)) : bb;
            const int tg = t0 + t;
            const int o0 = 24 * k1lo + k2 - a.out_off;
            if (PASS == 0 && a.out_block == 0) {
                V* dst = reinterpret_cast<V*>(outp) + (field * a.nrow + tg) * (size_t)a.nx;
#pragma unroll
                for (int c = 0; c < RL; ++c) {
                    const int o = o0 + 24 * NB * c;
                    if ((unsigned)o < (unsigned)a.n_out) dst[o] = mk2<T>(v[c].r, sgn * v[c].i);
                }
            } else if (PASS == 1) {
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
        __syncthreads();      // the work buffer and this operand buffer are free for the next items
    }
}

// can a launch use the pipelined kernel?  Default-halo geometry (lq >= 0), conjugate-symmetric spectrum for
// pass X, even nx for pass Y, enough work to keep every resident CTA busy for several items, and the shared
// memory map must fit.
inline bool fft24p_usable(int lq, int pass, const FftHArgs& a, int nfields, bool f32, size_t smem_optin, int num_sms,
                          Fft24pLayout* lay, int* grid)
{
    // 0 (default): never, 1: launches with >= 4 work items per resident CTA, 2: whenever possible.
    // Measured on B200 (profiles/r2_fft24p_throughput.jsonl): 128 fields of 1536^2 -> 512^2: 4.47 us per field,
    // the same as k_fft24 (k_fft48: 4.20); 3072^2 -> 1024^2: 24.8 us per field against 18.3.  Taking the operand
    // latency off the critical path buys nothing here: the passes are bound by the shared-memory / LSU path
    // (operands staged through shared memory cross it once more) and the FP64 pipe, not by exposed latency.
    const int mode = fft_env_int("BLDFM_B200_FFT24P", 0);
    if (mode == 0 || lq < 5 || lq > 9) return false;
    if (pass == 0 && (!a.hs || a.out_block != 0 || a.row0 != 0)) return false;
    if (pass == 1 && ((a.nx & 1) || a.nrow != a.nly / 2 + 1)) return false;
    if (((size_t)a.nlx * (f32 ? 8 : 16)) % 16) return false;
    const int ctas = lq <= 6 ? 2 : 1;
    const int cw = fft24p_cw(lq);
    const int ngroups = (a.ntrans + cw - 1) / cw;
    const int64_t nwork = (int64_t)ngroups * nfields;
    const int slots = num_sms * ctas;
    if (mode == 1 && nwork < 4 * (int64_t)slots) return false;
    *lay = fft24p_layout(lq, pass, a.nlx, a.nly / 2 + 1, f32, smem_optin, ctas);
    if (lay->total * ctas > smem_optin) return false;
    *grid = (int)std::min<int64_t>(nwork, slots);
    return true;
}

template <typename T, int PASS>
inline cudaError_t fft24p_launch_pass(cudaStream_t stream, size_t smem_optin, int lq, const FftHArgs& a, int nfields,
                                      const Fft24pLayout& lay, int grid)
{
    const bool pdl = fft_env_int("BLDFM_B200_PDL", 1) != 0;
    Fft24pArgs pa{};
    pa.h = a;
    const int cw = fft24p_cw(lq);
    pa.ngroups = (a.ntrans + cw - 1) / cw;
    pa.nfields = nfields;
    pa.nwork = pa.ngroups * nfields;
    pa.stage_elems = lay.stage_elems; pa.tw_elems = lay.tw_elems;
    pa.off_stage0 = (uint32_t)lay.off_stage[0]; pa.off_stage1 = (uint32_t)lay.off_stage[1];
    pa.off_buf = (uint32_t)lay.off_buf; pa.off_tw = (uint32_t)lay.off_tw; pa.off_bar = (uint32_t)lay.off_bar;
    const int threads = (PASS == 0 && lq >= 7 && lq <= 8) ? 192 : kFft24Threads;
#define BLDFM_FFT24P_CASE(LQ)                                                                                  \
    case LQ: {                                                                                                 \
        cudaError_t e = set_max_dyn_smem(k_fft24p<T, PASS, LQ>, (int)smem_optin);                                                 \
        if (e != cudaSuccess) return e;                                                                        \
        cudaLaunchConfig_t cfg = {};                                                                           \
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads);                            \
        cfg.dynamicSmemBytes = lay.total; cfg.stream = stream;                                                 \
        cudaLaunchAttribute at[1];                                                                             \
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                         \
        at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;                                        \
        cfg.attrs = at; cfg.numAttrs = 1;                                                                      \
        e = cudaLaunchKernelEx(&cfg, k_fft24p<T, PASS, LQ>, pa);                                               \
        if (e != cudaSuccess) return e;                                                                        \
        break;                                                                                                 \
    }
    switch (lq) {
        BLDFM_FFT24P_CASE(5)
        BLDFM_FFT24P_CASE(6)
        BLDFM_FFT24P_CASE(7)
        BLDFM_FFT24P_CASE(8)
        BLDFM_FFT24P_CASE(9)
        default: return cudaErrorInvalidValue;
    }
#undef BLDFM_FFT24P_CASE
    return cudaGetLastError();
}

template <typename T, int PASS>
inline bool herm_try_fft24p(cudaStream_t stream, size_t smem_optin, int lq, const FftHArgs& a, int nfields)
{
    Fft24pLayout lay;
    int grid = 0;
    if (!fft24p_usable(lq, PASS, a, nfields, sizeof(T) == 4, smem_optin, 148, &lay, &grid)) return false;
    return fft24p_launch_pass<T, PASS>(stream, smem_optin, lq, a, nfields, lay, grid) == cudaSuccess;
}

}  // namespace bldfm
