// march.cuh -- the per-(kx,ky)-mode vertical march of BLDFM on sm_100a.
//
// Replaces ivp_solver (src/bldfm/solver.py:307-374) AND the numpy stages around it that the
// reference runs as separate array passes: top eigenvalue (solver.py:164-174), shooting
// coefficient + combination (:228-235), the degenerate (0,0) mode (:239-251) and the phase
// shift (:254-262).  One thread owns one Fourier mode and carries BOTH initial-value problems in
// registers; per-level coefficients are staged once per CTA in shared memory; only the combined,
// shifted spectra of the requested levels are written (coalesced, 16 B per thread).
//
// Arithmetic modes
//   FMA=false (default): every operation of the march is an individually rounded IEEE binary64
//       op in the reference's order -> equal to the reference's numba code bit for
//       bit (up to the sign of exact zeros).
//   FMA=true  (opt-in) : algebraically identical, contracted into FMAs (46 instead of 85 FP64
//       instructions per mode-step).  Differs from the reference at the self-noise level.
//   SWEEP: the same discrete two-point problem solved by ONE downward sweep instead of two upward
//       initial-value problems -- see sweep_body (one output level) and sweep_multi_body (several).
#pragma once

#include "common.cuh"

namespace bldfm {

// ------------------------------------------------------------------------------------------------
// complex helpers mirroring the numpy kernels that surround the march
// ------------------------------------------------------------------------------------------------

// numpy contiguous complex128 multiply: re = fma(ar,br,-(ai*bi)), im = fma(ar,bi,ai*br)
__device__ __forceinline__ void cmul_np(double ar, double ai, double br, double bi, double& cr,
                                        double& ci)
{
    cr = fma(ar, br, -(ai * bi));
    ci = fma(ar, bi, ai * br);
}

// numpy complex128 divide: Smith's algorithm, no FMA
__device__ __forceinline__ void cdiv_np(double ar, double ai, double br, double bi, double& cr,
                                        double& ci)
{
    if (fabs(br) >= fabs(bi)) {
        const double rat = bi / br;
        const double scl = 1.0 / (br + bi * rat);
        cr = (ar + ai * rat) * scl;
        ci = (ai - ar * rat) * scl;
    } else {
        const double rat = br / bi;
        const double scl = 1.0 / (bi + br * rat);
        cr = (ar * rat + ai) * scl;
        ci = (ai * rat - ar) * scl;
    }
}

// hypot to (almost always) correct rounding: x*x + y*y in double-double, one Newton correction.
__device__ __forceinline__ double hypot_cr(double x, double y)
{
    const double x2 = x * x, ex = fma(x, x, -x2);
    const double y2 = y * y, ey = fma(y, y, -y2);
    const double s = x2 + y2;
    const double bb = s - x2;
    const double es = (x2 - (s - bb)) + (y2 - bb);   // two-sum error
    const double h = sqrt(s);
    if (h == 0.0) return 0.0;
    const double r = fma(-h, h, s) + (es + ex + ey); // s_exact - h*h
    return h + r / (2.0 * h);
}

// csqrt as glibc computes it for finite, non-tiny arguments (numpy's np.sqrt on complex128 calls it)
__device__ __forceinline__ void csqrt_np(double re, double im, double& sr, double& si)
{
    if (re == 0.0 && im == 0.0) { sr = 0.0; si = im; return; }
    const double d = hypot_cr(re, im);
    double r, s;
    if (re > 0.0) {
        r = sqrt(0.5 * (d + re));
        s = 0.5 * (im / r);
    } else {
        s = sqrt(0.5 * (d - re));
        r = fabs(0.5 * (im / s));
    }
    sr = r;
    si = copysign(s, im);
}

// ------------------------------------------------------------------------------------------------
// one march step
// ------------------------------------------------------------------------------------------------
struct Prop { double ar, ai, br, bi, cr, ci; };   // a (= d), b, c of solver.py:361-364

template <bool FMA>
__device__ __forceinline__ void propagator(const LevelCoef& c, double lx, double ly, double lx2,
                                           double ly2, Prop& P)
{
    if (!FMA) {
        // exact operation order of SURVEY.md A.2 ; --fmad=false keeps each op individually rounded
        const double tr = -(c.Kx * lx2 + c.Ky * ly2);
        const double ti = -(c.u * lx) - (c.v * ly);
        // (0.0 - x) of the reference is written -x: identical except for the sign of an exact zero,
        // and the negation folds into the consumers' operand modifiers (2 FP64 instructions saved)
        P.ar = 1.0 - (c.s * tr) * c.h2;
        P.ai = -((c.s * ti) * c.h2);
        P.br = c.c0 - (c.s6 * tr) * c.h3;
        P.bi = -((c.s6 * ti) * c.h3);
        const double t2r = tr * tr - ti * ti;
        const double m = tr * ti;                  // t2i = tr*ti + ti*tr == 2m exactly
        P.cr = tr * c.h - (c.s61 * t2r) * c.h3;
        // (s61*(2m))*h3 == 2*((s61*m)*h3) exactly (scaling by 2 commutes with rounding), and
        // fma(-2, x, y) rounds y - 2x once, like the reference's subtraction: one DADD saved
        P.ci = fma(-2.0, (c.s61 * m) * c.h3, ti * c.h);
    } else {
        const double tr = -fma(c.Kx, lx2, c.Ky * ly2);
        const double ti = -fma(c.u, lx, c.v * ly);
        P.ar = fma(-c.sh2, tr, 1.0);
        P.ai = -(c.sh2 * ti);
        P.br = fma(-c.s6h3, tr, c.c0);
        P.bi = -(c.s6h3 * ti);
        // c = T*(h - s61h3*T)
        const double gr = fma(-c.s61h3, tr, c.h);
        const double gi = -(c.s61h3 * ti);
        P.cr = fma(tr, gr, -(ti * gi));
        P.ci = fma(tr, gi, ti * gr);
    }
}

template <bool FMA>
__device__ __forceinline__ void apply(const Prop& P, double& pr, double& pi, double& qr, double& qi)
{
    double npr, npi, nqr, nqi;
    if (!FMA) {
        npr = (P.ar * pr - P.ai * pi) + (P.br * qr - P.bi * qi);
        npi = (P.ar * pi + P.ai * pr) + (P.br * qi + P.bi * qr);
        nqr = (P.cr * pr - P.ci * pi) + (P.ar * qr - P.ai * qi);
        nqi = (P.cr * pi + P.ci * pr) + (P.ar * qi + P.ai * qr);
    } else {
        npr = fma(P.ar, pr, fma(-P.ai, pi, fma(P.br, qr, -(P.bi * qi))));
        npi = fma(P.ar, pi, fma(P.ai, pr, fma(P.br, qi, P.bi * qr)));
        nqr = fma(P.cr, pr, fma(-P.ci, pi, fma(P.ar, qr, -(P.ai * qi))));
        nqi = fma(P.cr, pi, fma(P.ci, pr, fma(P.ar, qi, P.ai * qr)));
    }
    pr = npr; pi = npi; qr = nqr; qi = nqi;
}

// ------------------------------------------------------------------------------------------------
// fused march kernel
// ------------------------------------------------------------------------------------------------
struct MarchArgs {
    int32_t nlx, nly;          // retained modes
    int32_t ky0, nrows;        // this launch marches rows ky0 .. ky0+nrows-1 (ky-slab sharding)
    int32_t nly_loc;           // rows per field of the output spectra (nrows; nly in half-plane mode)
    int32_t nlv;               // output rows
    int32_t coef_stride;       // LevelCoef entries per group (= max steps)
    int32_t nrow_of;           // entries in row_of (= max nz)
    int32_t snap_level;        // single-row path: level to snapshot, -1 = none visited
    int32_t last_level;        // multi-row path: highest requested level
    int32_t single;            // precision == "single": round tfftp/tfftq to complex64 values
    int32_t out_f32;           // spectra stored as float2 (no shift applied, single)
    int32_t footprint;
    int32_t herm;              // 1: rows lie in the half-plane ky <= nly/2; conjugates are stored at (-ky,-kx)
    int32_t skip_mirror;       // 1: ... except that nobody will read them (see below): the mirror stores are skipped
    int32_t src_pitch;         // row pitch (complex elements) of src_spec
    int32_t src_nfx, src_nfy;  // size of the forward spectrum for index wrapping
    int32_t src_ky0;           // first ky row held by src_spec (ky-slab sharding), else 0
    double  q0_const;          // (1/nxe)/nye                                   solver.py:134
    double  src_scale;         // 1/(nxe*nye)   norm="forward"                  solver.py:136
    const double2*   src_spec; // forward spectrum of the padded source (non-footprint)
    const LevelCoef* coef;     // [ngroups][coef_stride]
    const int32_t*   row_of;   // [nrow_of]  level -> output row or -1
    const GroupDesc* groups;
    const TowerDesc* towers;
    const double*    lx;       // [nlx]
    const double*    ly;       // [nly]
    void* outp;                // [slot][row][nly_loc][nlx] complex (double2 | float2)
    void* outq;
    int64_t slot_stride;       // complex elements between output slots (= nlv*nly_loc*nlx)
    unsigned long long* trace; // diagnostics (BLDFM_B200_MARCH_TRACE): per CTA %globaltimer at start, after
                               // staging, after the march loop, at the end; NULL = off
};

__device__ __forceinline__ unsigned long long march_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ double round_f32(double x) { return (double)(float)x; }

struct Emit {
    const MarchArgs& a;
    const GroupDesc& gd;
    const TowerDesc* tw;
    int64_t mode;
    int64_t mirror;      // index of the mode (-ky,-kx) that receives the conjugate, or -1
    double lx, ly;
    double cs0, sn0;     // phase factor of the group's first tower, hoisted out of the row loop

    __device__ __forceinline__ Emit(const MarchArgs& a_, const GroupDesc& gd_, const TowerDesc* tw_, int64_t mode_,
                                    int64_t mirror_, double lx_, double ly_)
        // skip_mirror: the consumer is pass X of the sparse radix-24/48 back-transform, which for a
        // conjugate-symmetric spectrum of even size reads the rows ky <= nly/2 and, of the other rows, only the
        // Nyquist column -- written by the extra threads as their OWN mode (fft24.cuh, "interior row"); the
        // mirror stores would be half of this kernel's store traffic for nothing
        : a(a_), gd(gd_), tw(tw_), mode(mode_), mirror(a_.skip_mirror ? -1 : mirror_), lx(lx_), ly(ly_),
          cs0(1.0), sn0(0.0)
    {
        if (gd.tow_count > 0 && tw[0].shift) phase(tw[0], cs0, sn0);
    }

    // exp(1j*(Lx*sx + Ly*sy))                                              solver.py:255 / :260
    __device__ __forceinline__ void phase(const TowerDesc& td, double& cs, double& sn) const
    {
        const double th = lx * td.sx + ly * td.sy;
        sincos(th, &sn, &cs);
    }

    // store one combined (p,q) pair into row `row` of every tower slot of the group
    __device__ __forceinline__ void operator()(int row, double pr, double pi, double qr,
                                               double qi) const
    {
        if (a.single) {
            pr = round_f32(pr); pi = round_f32(pi); qr = round_f32(qr); qi = round_f32(qi);
        }
        const int64_t rowbase = (int64_t)row * a.nly_loc * a.nlx;
        const int64_t rowoff = rowbase + mode;
        for (int t = 0; t < gd.tow_count; ++t) {
            const TowerDesc td = tw[t];
            double opr = pr, opi = pi, oqr = qr, oqi = qi;
            if (td.shift) {
                double sn = sn0, cs = cs0;
                if (t > 0) phase(td, cs, sn);
                cmul_np(pr, pi, cs, sn, opr, opi);
                cmul_np(qr, qi, cs, sn, oqr, oqi);
            }
            const int64_t o = (int64_t)td.slot * a.slot_stride + rowoff;
            if (a.out_f32) {
                reinterpret_cast<float2*>(a.outp)[o] = make_float2((float)opr, (float)opi);
                reinterpret_cast<float2*>(a.outq)[o] = make_float2((float)oqr, (float)oqi);
            } else {
                reinterpret_cast<double2*>(a.outp)[o] = make_double2(opr, opi);
                reinterpret_cast<double2*>(a.outq)[o] = make_double2(oqr, oqi);
            }
            if (mirror >= 0) {
                // mode (-ky,-kx): every operation of the march, the radiation condition and the
                // phase factor is sign-symmetric in the imaginary parts, so the reference's value
                // there is the exact complex conjugate (bit for bit)
                const int64_t om = (int64_t)td.slot * a.slot_stride + rowbase + mirror;
                if (a.out_f32) {
                    reinterpret_cast<float2*>(a.outp)[om] = make_float2((float)opr, -(float)opi);
                    reinterpret_cast<float2*>(a.outq)[om] = make_float2((float)oqr, -(float)oqi);
                } else {
                    reinterpret_cast<double2*>(a.outp)[om] = make_double2(opr, -opi);
                    reinterpret_cast<double2*>(a.outq)[om] = make_double2(oqr, -oqi);
                }
            }
        }
    }
};

// Thread -> mode map.  Full plane: one thread per retained mode of rows ky0 .. ky0+nly_loc-1.
// Half plane (herm): T(-k) = conj(T(k)), both initial states are conjugate-symmetric for a real source,
// hence state(-k) = conj(state(k)): only rows ky = 0 .. nly/2 are marched and each thread also stores
// the conjugate at (-ky,-kx).  Modes without a partner in the retained set are marched on their own:
// row 0 (its partner is in the same row; marched in full), the Nyquist row ky = nly/2 and the Nyquist
// column kx = nlx/2 (even sizes; fftfreq keeps only -n/2) -- the latter's lower half by extra threads
// appended after the rows of the launch.
struct ModeMap { int ky, kx; int64_t mode, mirror; };

// extra threads of a half-plane launch over rows [row0, row0+rows): the modes (-ky, Nyquist kx) of the
// rows ky in [e_lo, e_lo+n_extra) that have a partner row
__host__ __device__ __forceinline__ void march_extras(int nlx, int nly, int row0, int rows, int& e_lo, int& n_extra)
{
    const int nmir = (nly - 1) / 2;                        // rows 1..nmir have a partner row nly-ky
    e_lo = row0 > 1 ? row0 : 1;
    const int e_hi = (row0 + rows) < (nmir + 1) ? (row0 + rows) : (nmir + 1);
    n_extra = (nlx % 2 == 0 && e_hi > e_lo) ? e_hi - e_lo : 0;
}

__host__ __device__ __forceinline__ int64_t march_thread_count(int nlx, int nly, int row0, int rows, bool herm)
{
    int e_lo, n_extra = 0;
    if (herm) march_extras(nlx, nly, row0, rows, e_lo, n_extra);
    return (int64_t)nlx * rows + n_extra;
}

__host__ __device__ __forceinline__ bool march_map(const MarchArgs& a, int64_t tid, ModeMap& m)
{
    const int64_t nmain = (int64_t)a.nlx * a.nrows;
    if (!a.herm) {
        // spectra hold the rows ky0 .. ky0+nrows-1 only
        if (tid >= nmain) return false;
        const int kyl = (int)(tid / a.nlx);
        m.kx = (int)(tid - (int64_t)kyl * a.nlx);
        m.ky = a.ky0 + kyl;
        m.mode = tid;
        m.mirror = -1;
        return true;
    }
    // spectra hold all nly rows; this launch fills rows ky0 .. ky0+nrows-1 (<= nly/2) and their mirrors
    const int nmir = (a.nly - 1) / 2;
    const bool even_x = (a.nlx % 2) == 0;
    if (tid < nmain) {
        const int kyl = (int)(tid / a.nlx);
        m.kx = (int)(tid - (int64_t)kyl * a.nlx);
        m.ky = a.ky0 + kyl;
        m.mode = (int64_t)m.ky * a.nlx + m.kx;
        const bool has = m.ky >= 1 && m.ky <= nmir && !(even_x && m.kx == a.nlx / 2);
        m.mirror = has ? (int64_t)(a.nly - m.ky) * a.nlx + (m.kx == 0 ? 0 : a.nlx - m.kx) : -1;
        return true;
    }
    int e_lo, n_extra;
    march_extras(a.nlx, a.nly, a.ky0, a.nrows, e_lo, n_extra);
    const int64_t e = tid - nmain;
    if (e >= n_extra) return false;
    m.ky = a.nly - (e_lo + (int)e);
    m.kx = a.nlx / 2;
    m.mode = (int64_t)m.ky * a.nlx + m.kx;
    m.mirror = -1;
    return true;
}

constexpr int kMarchThreads = 128;

// ------------------------------------------------------------------------------------------------
// Downward sweep (arithmetic mode 2, one output level)
//
// The reference solves the discrete two-point problem  v_{i+1} = M_i v_i  (M_i = [[a,b],[c,a]] of solver.py:361-364),
// q_0 = q0 at the ground, q_S = Kz*eig*p_S at the top, by linear shooting: two initial-value problems marched
// UPWARD (solver.py:220-226) and combined with alpha (:228-235).  Upward, the wanted solution decays like
// e^{-kappa} while both auxiliary solutions grow like e^{+kappa}: their combination cancels e^{2 kappa} of
// round-off.  The same discrete solution follows from ONE vector swept DOWNWARD from the
// radiation condition:  w_S = (1, Kz*eig),  w_i = adj(M_i) w_{i+1} = det(M_i) M_i^{-1} w_{i+1}  with
// adj(M) = [[a,-b],[-c,a]] -- no division, the same a, b, c -- and
//     v_L = w_L * q0 * prod_{i<L} det(M_i) / (w_0).q ,      det(M_i) = a*a - b*c .
// Downward the wanted solution is the dominant one, so nothing cancels: in binary64 the sweep reproduces the
// extended-precision discrete solution to 1e-15 for every kappa, where the reference's own binary64 shooting is
// off by 7e-13 (kappa = 7) ... 2e-8 (kappa = 15) (tests/tools/sweep_accuracy.py).  Its deviation from the
// reference IS the reference's round-off, hence the same kappa gate as for the FMA march.
// Cost per mode-step: 14.5 (a, b, c) + 16 (one matrix-vector product) FP64 instructions above the output level,
// + 12 (determinant and its running product) below it, against 46.5 for the two upward problems.
// The host admits the sweep only where it can neither overflow nor meet a singular M_i (sweep_admissible).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void apply_adj(const Prop& P, double& pr, double& pi, double& qr, double& qi)
{
    const double npr = fma(P.ar, pr, fma(-P.ai, pi, fma(-P.br, qr, P.bi * qi)));
    const double npi = fma(P.ar, pi, fma(P.ai, pr, fma(-P.br, qi, -(P.bi * qr))));
    const double nqr = fma(P.ar, qr, fma(-P.ai, qi, fma(-P.cr, pr, P.ci * pi)));
    const double nqi = fma(P.ar, qi, fma(P.ai, qr, fma(-P.cr, pi, -(P.ci * pr))));
    pr = npr; pi = npi; qr = nqr; qi = nqi;
}

template <class Coef>
__device__ __forceinline__ void sweep_body(const MarchArgs& a, const GroupDesc& gd, const Emit& emit,
                                           const Coef& sc, double lx, double ly, double q0r, double q0i)
{
    const int S = gd.S;
    const double lx2 = lx * lx, ly2 = ly * ly;
    const int snap = a.snap_level;
    if (snap < 0 || snap > S) {                    // the requested level is never visited: rows stay zero
        sc.ready();                                // (no CTA may retire while its bulk copy is in flight)
        for (int r = 0; r < a.nlv; ++r) emit(r, 0.0, 0.0, 0.0, 0.0);
        return;
    }
    // radiation condition at the top (solver.py:164-174): w_S = (1, Kz_top*eig)
    double pr = 1.0, pi = 0.0, qr, qi;
    {
        const double are = gd.kxk * lx2 + gd.kyk * ly2;
        const double aim = gd.c1 * lx + gd.c2 * ly;
        double er, ei;
        csqrt_np(are, aim, er, ei);
        qr = gd.kz_top * er; qi = gd.kz_top * ei;
    }
    sc.ready();
    if (a.trace && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 4 + 1] = march_now();
    Prop P;
#pragma unroll 4
    for (int i = S - 1; i >= snap; --i) {
        propagator<true>(sc[i], lx, ly, lx2, ly2, P);
        apply_adj(P, pr, pi, qr, qi);
    }
    const double spr = pr, spi = pi, sqr = qr, sqi = qi;      // w at the output level
    double dr = 1.0, di = 0.0;                                // prod_{i<L} det(M_i)
#pragma unroll 4
    for (int i = snap - 1; i >= 0; --i) {
        propagator<true>(sc[i], lx, ly, lx2, ly2, P);
        apply_adj(P, pr, pi, qr, qi);
        const double er = fma(P.ar, P.ar, fma(-P.ai, P.ai, fma(-P.br, P.cr, P.bi * P.ci)));
        const double ei = fma(P.ar + P.ar, P.ai, -fma(P.br, P.ci, P.bi * P.cr));
        const double ndr = fma(dr, er, -(di * ei));
        di = fma(dr, ei, di * er);
        dr = ndr;
    }
    if (a.trace && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 4 + 2] = march_now();
    // s = q0 * D / (w_0).q  (Smith's division: |w_0| may be as large as e^{kappa_top})
    double nr, ni, sr, si;
    cmul_np(q0r, q0i, dr, di, nr, ni);
    cdiv_np(nr, ni, qr, qi, sr, si);
    double opr, opi, oqr, oqi;
    cmul_np(sr, si, spr, spi, opr, opi);
    cmul_np(sr, si, sqr, sqi, oqr, oqi);
    if (a.nlv > 0) emit(0, opr, opi, oqr, oqi);
    for (int r = 1; r < a.nlv; ++r) emit(r, 0.0, 0.0, 0.0, 0.0);
}

// Several output levels in sweep mode: the downward sweep only has to deliver alpha -- at the ground
// v_0 = (alpha, q0) = w_0 * q0 / (w_0).q, no determinant involved -- and the solution itself is then marched
// upward as ONE vector from (alpha, q0), emitting the requested rows on the way: 2 x 30.5 FP64 instructions per
// mode-step against 2 x 46.5 for the FMA shooting march (which carries both auxiliary problems twice).  The
// upward leg amplifies round-off like the reference's own combination alpha*P1 + P2 does (e^{2 kappa}), hence the
// same kappa gate.
template <class Coef>
__device__ __forceinline__ void sweep_multi_body(const MarchArgs& a, const GroupDesc& gd, const Emit& emit,
                                                 const Coef& sc, double lx, double ly, double q0r, double q0i)
{
    const int S = gd.S;
    const double lx2 = lx * lx, ly2 = ly * ly;
    double pr = 1.0, pi = 0.0, qr, qi;
    {
        const double are = gd.kxk * lx2 + gd.kyk * ly2;
        const double aim = gd.c1 * lx + gd.c2 * ly;
        double er, ei;
        csqrt_np(are, aim, er, ei);
        qr = gd.kz_top * er; qi = gd.kz_top * ei;
    }
    sc.ready();
    if (a.trace && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 4 + 1] = march_now();
    Prop P;
#pragma unroll 4
    for (int i = S - 1; i >= 0; --i) {
        propagator<true>(sc[i], lx, ly, lx2, ly2, P);
        apply_adj(P, pr, pi, qr, qi);
    }
    // alpha = q0 * (w_0).p / (w_0).q
    double rr, ri, alr, ali;
    cdiv_np(pr, pi, qr, qi, rr, ri);
    cmul_np(q0r, q0i, rr, ri, alr, ali);
    if (a.trace && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 4 + 2] = march_now();
    pr = alr; pi = ali; qr = q0r; qi = q0i;
    const int last = a.last_level < S ? a.last_level : S;
    int rows = 0;
    for (int i = 0; i <= last; ++i) {
        const int row = sc.row(i, a);
        if (row >= 0) { emit(row, pr, pi, qr, qi); ++rows; }
        if (i < last) {
            propagator<true>(sc[i], lx, ly, lx2, ly2, P);
            apply<true>(P, pr, pi, qr, qi);
        }
    }
    for (int r = rows; r < a.nlv; ++r) emit(r, 0.0, 0.0, 0.0, 0.0);
}

// The per-level table of the march, staged once per CTA in shared memory from the per-solve parameter buffer.
// (Measured alternative, round 2: the table carried in the kernel's parameter space and read with
// warp-uniform LDC inside the loop -- no H2D copy, no staging, no barrier -- is 14 % (exact) to 36 % (fma)
// SLOWER at config 2: the constant-bank loads do not keep up with 8 broadcast LDS.128 per step.
// Also measured: software-pipelining the propagator of step i+1 next to the state update of step i (more
// independent DFMA chains for the FMA mode, whose top stall is "wait") needs 12 more live registers than
// the 72 that 7 CTAs per SM allow; the spills land inside the loop: 151 us instead of 61 us.)
struct SmemCoef {
    const LevelCoef* sc;
    const int32_t* srow;
    uint64_t* bar;       // mbarrier the bulk copy of the table completes on
    __device__ __forceinline__ const LevelCoef& operator[](int i) const { return sc[i]; }
    __device__ __forceinline__ int row(int i, const MarchArgs&) const { return srow[i]; }
    // called once by every thread, after the table-independent part of its prologue and before the first
    // table access: both bulk copies (table and row_of) have completed on the mbarrier.  Deliberately no
    // CTA barrier here: the threads arrive from different branches (mode (0,0), sweep, shooting).
    __device__ __forceinline__ void ready() const { mbar_wait(bar, 0); }
};

template <int ARITH, bool MULTI, class Coef>
__device__ __forceinline__ void march_body(const MarchArgs& a, const GroupDesc& gd, const TowerDesc* towers,
                                           const Coef& sc, int64_t tid)
{
    constexpr bool FMA = ARITH != 0;
    const int S = gd.S;
    ModeMap mm;
    if (!march_map(a, tid, mm)) return;      // (a thread that never reads the table need not wait for it)
    const int kx = mm.kx, ky = mm.ky;
    const double lx = a.lx[kx], ly = a.ly[ky];

    // source spectrum of this mode (solver.py:134 / :136-145)
    double q0r, q0i;
    if (a.footprint) {
        q0r = a.q0_const; q0i = 0.0;
    } else {
        const int pk = (a.nlx + 1) / 2, pl = (a.nly + 1) / 2;
        const int wx = kx < pk ? kx : kx - a.nlx + a.src_nfx;
        const int wy = ky < pl ? ky : ky - a.nly + a.src_nfy;
        const double2 sv = a.src_spec[(size_t)(wy - a.src_ky0) * a.src_pitch + wx];
        q0r = sv.x * a.src_scale; q0i = sv.y * a.src_scale;
    }

    const Emit emit(a, gd, towers, mm.mode, mm.mirror, lx, ly);

    const bool is00 = ky == 0 && kx == 0;
    auto mode00 = [&]() {
        // degenerate mode: flux constant, concentration by the trapezoid rule (solver.py:190-191,239-251)
        double pr = gd.p000, pi = 0.0;
        bool any = false;
        for (int i = 0; i < S; ++i) {
            const int r = sc.row(i, a);
            if (r >= 0) { emit(r, pr, pi, q0r, q0i); any = true; }
            pr = pr - (q0r * sc[i].h) * sc[i].w;
            pi = pi - (q0i * sc[i].h) * sc[i].w;
        }
        if (sc.row(S, a) >= 0) { emit(sc.row(S, a), pr, pi, q0r, q0i); any = true; }
        // rows never visited keep tfftp[0,0,0]=p000 (row 0) / 0 and tfftq[:,0,0]=tfftq0[0,0]
        int visited = 0;
        for (int i = 0; i <= S; ++i) visited += (sc.row(i, a) >= 0);
        for (int r = visited; r < a.nlv; ++r)
            emit(r, (r == 0 && !any) ? gd.p000 : 0.0, 0.0, q0r, q0i);
    };
    if (is00) { sc.ready(); mode00(); return; }
    if (ARITH == 2 && !MULTI) { sweep_body(a, gd, emit, sc, lx, ly, q0r, q0i); return; }
    if (ARITH == 2 && MULTI) { sweep_multi_body(a, gd, emit, sc, lx, ly, q0r, q0i); return; }

    const double lx2 = lx * lx, ly2 = ly * ly;
    sc.ready();
    if (a.trace && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 4 + 1] = march_now();

    // IVP1 from (1,0), IVP2 from (0,q0)   (solver.py:220-226)
    double p1r = 1.0, p1i = 0.0, q1r = 0.0, q1i = 0.0;
    double p2r = 0.0, p2i = 0.0, q2r = q0r, q2i = q0i;
    // single-row snapshot registers
    double s1pr = 0.0, s1pi = 0.0, s1qr = 0.0, s1qi = 0.0;
    double s2pr = 0.0, s2pi = 0.0, s2qr = 0.0, s2qi = 0.0;

    Prop P;
    if (!MULTI) {
        const int snap = a.snap_level;
        const int n1 = (snap >= 0 && snap <= S) ? snap : S;
#pragma unroll 4
        for (int i = 0; i < n1; ++i) {
            propagator<FMA>(sc[i], lx, ly, lx2, ly2, P);
            apply<FMA>(P, p1r, p1i, q1r, q1i);
            apply<FMA>(P, p2r, p2i, q2r, q2i);
        }
        if (snap >= 0 && snap <= S) {
            s1pr = p1r; s1pi = p1i; s1qr = q1r; s1qi = q1i;
            s2pr = p2r; s2pi = p2i; s2qr = q2r; s2qi = q2i;
        }
#pragma unroll 4
        for (int i = n1; i < S; ++i) {
            propagator<FMA>(sc[i], lx, ly, lx2, ly2, P);
            apply<FMA>(P, p1r, p1i, q1r, q1i);
            apply<FMA>(P, p2r, p2i, q2r, q2i);
        }
    } else {
#pragma unroll 4
        for (int i = 0; i < S; ++i) {
            propagator<FMA>(sc[i], lx, ly, lx2, ly2, P);
            apply<FMA>(P, p1r, p1i, q1r, q1i);
            apply<FMA>(P, p2r, p2i, q2r, q2i);
        }
    }
    if (a.trace && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 4 + 2] = march_now();

    // radiation condition at the top: eig, alpha (solver.py:164-174, 228-230)
    double alr, ali;
    {
        const double are = gd.kxk * lx2 + gd.kyk * ly2;
        const double aim = gd.c1 * lx + gd.c2 * ly;
        double er, ei;
        csqrt_np(are, aim, er, ei);
        const double ker = gd.kz_top * er, kei = gd.kz_top * ei;
        double t2r, t2i, t1r, t1i;
        cmul_np(ker, kei, p2r, p2i, t2r, t2i);
        cmul_np(ker, kei, p1r, p1i, t1r, t1i);
        const double nr = -(q2r - t2r), ni = -(q2i - t2i);
        const double dr = q1r - t1r, di = q1i - t1i;
        cdiv_np(nr, ni, dr, di, alr, ali);
    }

    if (!MULTI) {
        // p = alpha*P1 + P2 ; q = alpha*Q1 + Q2   (solver.py:234-235)
        double mr, mi, pr, pi, qr, qi;
        cmul_np(alr, ali, s1pr, s1pi, mr, mi); pr = mr + s2pr; pi = mi + s2pi;
        cmul_np(alr, ali, s1qr, s1qi, mr, mi); qr = mr + s2qr; qi = mi + s2qi;
        if (a.nlv > 0) emit(0, pr, pi, qr, qi);
        for (int r = 1; r < a.nlv; ++r) emit(r, 0.0, 0.0, 0.0, 0.0);
    } else {
        // re-march (deterministic, so bitwise the same states) and emit each requested row
        p1r = 1.0; p1i = 0.0; q1r = 0.0; q1i = 0.0;
        p2r = 0.0; p2i = 0.0; q2r = q0r; q2i = q0i;
        const int last = a.last_level < S ? a.last_level : S;
        int rows = 0;
        for (int i = 0; i <= last; ++i) {
            const int row = sc.row(i, a);
            if (row >= 0) {
                double mr, mi, pr, pi, qr, qi;
                cmul_np(alr, ali, p1r, p1i, mr, mi); pr = mr + p2r; pi = mi + p2i;
                cmul_np(alr, ali, q1r, q1i, mr, mi); qr = mr + q2r; qi = mi + q2i;
                emit(row, pr, pi, qr, qi);
                ++rows;
            }
            if (i < last) {
                propagator<FMA>(sc[i], lx, ly, lx2, ly2, P);
                apply<FMA>(P, p1r, p1i, q1r, q1i);
                apply<FMA>(P, p2r, p2i, q2r, q2i);
            }
        }
        for (int r = rows; r < a.nlv; ++r) emit(r, 0.0, 0.0, 0.0, 0.0);
    }
}

// grid = (ceil(nthreads / kMarchThreads), ngroups) ; dynamic smem = coef_stride*128 + nrow_of*4
//
// (Measured alternative, round 2: ONE CTA of 896 threads per SM whose 28 warps meet at a barrier every 8 steps.
// Motivation: per-CTA timestamps show the warp schedulers sharing the FP64 pipe unevenly between the 7 resident
// CTAs -- the median CTA is done at 35 us (FMA) while the kernel runs until 52 us.  In lock-step every CTA takes
// 48 us, but the kernel is no faster (60.9 vs 59.2 us FMA, 83.9 vs 82.0 us exact; 413 steps: 205 vs 199 us): the
// uneven progress never cost throughput, and a full-SM CTA has to wait for its SM to drain completely.)
template <int ARITH, bool MULTI>
__global__ void __launch_bounds__(kMarchThreads, 7)
k_march(const MarchArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LevelCoef* sc = reinterpret_cast<LevelCoef*>(smem_raw);
    int32_t* srow = reinterpret_cast<int32_t*>(sc + a.coef_stride);
    const uint32_t row_bytes = ((uint32_t)a.nrow_of * 4u + 15u) & ~15u;
    uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(srow) + row_bytes);

    // the back-transform that follows may be launched as soon as every CTA of this grid is resident; its
    // CTAs start when ours retire and wait for the spectra in cudaGridDependencySynchronize()
    cudaTriggerProgrammaticLaunchCompletion();
    if (a.trace && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 4 + 0] = march_now();
    // the parameter block may still be on its way (k_fetch_params in front of a programmatic launch); a no-op
    // for an ordinary launch
    cudaGridDependencySynchronize();
    const GroupDesc gd = a.groups[blockIdx.y];
    // The per-level table (S x 128 B) and row_of arrive by two bulk copies (cp.async.bulk -> mbarrier) issued by
    // one thread; every thread runs the table-independent part of its prologue (mode map, wavenumbers, source
    // spectrum, the tower's sincos, the top eigenvalue) meanwhile and waits in SmemCoef::ready().
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t coef_bytes = (uint32_t)gd.S * (uint32_t)sizeof(LevelCoef);
        mbar_expect_tx(bar, coef_bytes + row_bytes);
        bulk_g2s(sc, a.coef + (size_t)blockIdx.y * a.coef_stride, coef_bytes, bar);
        bulk_g2s(srow, a.row_of, row_bytes, bar);
    }
    march_body<ARITH, MULTI>(a, gd, a.towers + gd.tow_begin, SmemCoef{sc, srow, bar},
                             (int64_t)blockIdx.x * kMarchThreads + threadIdx.x);
    if (a.trace && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 4 + 3] = march_now();
}

// ------------------------------------------------------------------------------------------------
// analytic branch for constant profiles (solver.py:193-202), one level
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMarchThreads)
k_analytic(const MarchArgs a)
{
    const GroupDesc gd = a.groups[blockIdx.y];
    ModeMap mm;
    if (!march_map(a, (int64_t)blockIdx.x * kMarchThreads + threadIdx.x, mm)) return;
    const int kx = mm.kx, ky = mm.ky;
    const double lx = a.lx[kx], ly = a.ly[ky];
    double q0r, q0i;
    if (a.footprint) {
        q0r = a.q0_const; q0i = 0.0;
    } else {
        const int pk = (a.nlx + 1) / 2, pl = (a.nly + 1) / 2;
        const int wx = kx < pk ? kx : kx - a.nlx + a.src_nfx;
        const int wy = ky < pl ? ky : ky - a.nly + a.src_nfy;
        const double2 sv = a.src_spec[(size_t)(wy - a.src_ky0) * a.src_pitch + wx];
        q0r = sv.x * a.src_scale; q0i = sv.y * a.src_scale;
    }
    const Emit emit(a, gd, a.towers + gd.tow_begin, mm.mode, mm.mirror, lx, ly);
    const double h = gd.h_analytic;
    if (ky == 0 && kx == 0) {
        // tfftp[:,0,0] = p000 - tfftq0[0,0]*Kzinv*h ; tfftq[:,0,0] = tfftq0[0,0]
        const double pr = gd.p000 - (q0r * gd.kinv_top) * h;
        const double pi = 0.0 - (q0i * gd.kinv_top) * h;
        emit(0, pr, pi, q0r, q0i);
        return;
    }
    const double lx2 = lx * lx, ly2 = ly * ly;
    const double are = gd.kxk * lx2 + gd.kyk * ly2;
    const double aim = gd.c1 * lx + gd.c2 * ly;
    double er, ei;
    csqrt_np(are, aim, er, ei);
    // tfftq = tfftq0 * exp(-eig*h)
    const double xr = (-er) * h, xi = (-ei) * h;
    const double mag = exp(xr);
    double sn, cs;
    sincos(xi, &sn, &cs);
    double qr, qi;
    cmul_np(q0r, q0i, mag * cs, mag * sn, qr, qi);
    if (a.single) { qr = round_f32(qr); qi = round_f32(qi); }
    // tfftp = tfftq * Kzinv / eig
    double pr, pi;
    cdiv_np(qr * gd.kinv_top, qi * gd.kinv_top, er, ei, pr, pi);
    emit(0, pr, pi, qr, qi);
}

// ------------------------------------------------------------------------------------------------
// plain ivp_solver (solver.py:307-374) for isolated parity tests: arbitrary (p0,q0,Lx,Ly) per mode,
// raw snapshots P,Q [nlv][M].  Same propagator/apply code as the fused kernel.
// ------------------------------------------------------------------------------------------------
template <bool FMA>
__global__ void __launch_bounds__(kMarchThreads)
k_ivp(int64_t M, int S, const LevelCoef* __restrict__ coef, const int32_t* __restrict__ row_of,
      const double2* __restrict__ p0, const double2* __restrict__ q0,
      const double* __restrict__ Lx, const double* __restrict__ Ly,
      double2* __restrict__ p_top, double2* __restrict__ q_top,
      double2* __restrict__ Pout, double2* __restrict__ Qout)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LevelCoef* sc = reinterpret_cast<LevelCoef*>(smem_raw);
    int32_t* srow = reinterpret_cast<int32_t*>(sc + S);
    {
        const double2* src = reinterpret_cast<const double2*>(coef);
        double2* dst = reinterpret_cast<double2*>(sc);
        for (int i = threadIdx.x; i < S * 8; i += kMarchThreads) dst[i] = src[i];
        for (int i = threadIdx.x; i <= S; i += kMarchThreads) srow[i] = row_of[i];
    }
    __syncthreads();
    const int64_t m = (int64_t)blockIdx.x * kMarchThreads + threadIdx.x;
    if (m >= M) return;
    const double lx = Lx[m], ly = Ly[m];
    const double lx2 = lx * lx, ly2 = ly * ly;
    double pr = p0[m].x, pi = p0[m].y, qr = q0[m].x, qi = q0[m].y;
    Prop P;
    for (int i = 0; i < S; ++i) {
        if (srow[i] >= 0) {
            Pout[(size_t)srow[i] * M + m] = make_double2(pr, pi);
            Qout[(size_t)srow[i] * M + m] = make_double2(qr, qi);
        }
        propagator<FMA>(sc[i], lx, ly, lx2, ly2, P);
        apply<FMA>(P, pr, pi, qr, qi);
    }
    if (srow[S] >= 0) {
        Pout[(size_t)srow[S] * M + m] = make_double2(pr, pi);
        Qout[(size_t)srow[S] * M + m] = make_double2(qr, qi);
    }
    p_top[m] = make_double2(pr, pi);
    q_top[m] = make_double2(qr, qi);
}

}  // namespace bldfm
