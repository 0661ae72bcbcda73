/*
 * bldfm_b200.h -- C ABI of libbldfm_b200.so, the B200 (sm_100a) implementation of BLDFM's
 * steady-state spectral advection-diffusion hot path.
 *
 * The reference (pure Python) has no FFI; the "binding" a maintainer adds is a ctypes stub that
 * replaces the body of bldfm.solver.steady_state_transport_solver.  Each entry
 * point below cites the reference interface it replaces (paths relative to /root/reference/).
 *
 * Conventions
 *   - plain C types only; every function returns an int status (0 = BLDFM_OK, <0 = error) and never
 *     throws; bldfm_last_error_string() gives the message for the calling thread's last error.
 *   - complex arrays are interleaved (re,im) doubles == numpy complex128.
 *   - "host" pointers are ordinary process memory; "device" pointers are CUDA device memory on the
 *     plan's device.  Which one a buffer is, is stated per argument / selected by BLDFM_*_ON_DEVICE.
 *   - a plan is not thread-safe; use one plan per host thread (the analogue of the reference's
 *     per-process FFTManager singleton, src/bldfm/fft_manager.py:117-139).
 */
#ifndef BLDFM_B200_H
#define BLDFM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLDFM_OK                    0
#define BLDFM_ERR_ODD_MODES        -1  /* "modes must consist of even numbers."          src/bldfm/solver.py:90-91   */
#define BLDFM_ERR_PRECISION        -2  /* "precision must be single (default) or double." src/bldfm/solver.py:187-188 */
#define BLDFM_ERR_INVALID          -3  /* bad argument (message says which)                                          */
#define BLDFM_ERR_CUDA             -4  /* CUDA runtime error (no device, launch failure, ...)                        */
#define BLDFM_ERR_CUFFT            -5
#define BLDFM_ERR_ALLOC            -6
#define BLDFM_ERR_ODD_PAD          -7  /* (nxe-nlx) or (nye-nly) odd: src/bldfm/solver.py:130,142 mis-slices there    */
#define BLDFM_ERR_ANALYTIC_LEVELS  -8  /* analytic=True broadcasts only for one level    src/bldfm/solver.py:197-202 */
#define BLDFM_ERR_LEVEL_RANGE      -9  /* z[levels] would raise IndexError               src/bldfm/solver.py:296     */

/* flags for bldfm_solve* (bit-or) */
#define BLDFM_FOOTPRINT        0x001  /* footprint=True                                  src/bldfm/solver.py:25  */
#define BLDFM_ANALYTIC         0x002  /* analytic=True                                   src/bldfm/solver.py:26  */
#define BLDFM_DOUBLE           0x004  /* precision="double" (absent: "single")           src/bldfm/solver.py:28  */
#define BLDFM_MARCH_FMA        0x008  /* opt-in: FMA-contracted march (faster, not bit-mirrored)                 */
#define BLDFM_SRC_ON_DEVICE    0x010  /* srf_flx is a device pointer                                              */
#define BLDFM_OUT_ON_DEVICE    0x020  /* conc/flx are device pointers                                             */
#define BLDFM_ASYNC            0x040  /* enqueue only, do not synchronise.  Device outputs: ordered on the plan's
                                         stream.  Host outputs: they must be PINNED and may be read only after
                                         bldfm_plan_synchronize(); the D2H of one solve then overlaps the next
                                         solve's kernels (double-buffered device results, separate copy stream) */
#define BLDFM_FFT_LIBRARY      0x080  /* force the cuFFT transform path instead of the pruned in-house kernels    */
#define BLDFM_FFT_FULL         0x100  /* in-house back-transform without the real-output (Hermitian) halving      */
#define BLDFM_DELIVER_F32      0x800  /* opt-in, host outputs only: conc/flx are delivered as float32 even where the
                                         reference returns float64 (rounded on the device; half the PCIe bytes)  */
#define BLDFM_OUT_MAPPED       0x1000 /* host outputs only: conc and flx lie in ONE page-locked allocation that is mapped
                                         into the device address space (cudaHostAlloc / bldfm_host_alloc / one
                                         bldfm_host_register range).  If they are adjacent, one D2H copy delivers
                                         both.  Opt-in on top (option BLDFM_B200_DIRECT_HOST = result bytes up to
                                         which): the last kernel stores straight into them, no D2H copy -- measured
                                         slower for the transform's 64-byte row segments, see DESIGN.md 6         */
#define BLDFM_MARCH_SWEEP     0x2000 /* opt-in: the two-point problem of every mode is solved by a single downward sweep
                                         from the radiation condition instead of two upward initial-value problems
                                         (same discrete solution, fewer flops; differs from the reference by the
                                         reference's own round-off).  One output level: the sweep alone; several:
                                         sweep for alpha, then one vector upward.  Falls back to BLDFM_MARCH_FMA
                                         where the sweep could overflow (bldfm_sweep_admissible)                */
#define BLDFM_MARCH_AUTO       0x400  /* fast march (sweep for one output level, FMA-contracted for several or where the sweep is not admissible)
                                         where linear shooting is well conditioned (kappa at the
                                         highest output level <= bldfm_auto_kappa_limit() for every march of the
                                         call: predicted deviation from the reference <= 1e-11 rel-L2, SURVEY.md
                                         Appendix C), the bit-mirrored march otherwise                          */
#define BLDFM_MARCH_FULL       0x200  /* march every retained mode instead of the half-plane ky <= nly/2 whose
                                         conjugates fill the other half (cross-check; same values bit for bit)  */

typedef struct bldfm_plan bldfm_plan;

/* Grid bookkeeping of steady_state_transport_solver (src/bldfm/solver.py:93-130). */
typedef struct bldfm_geometry {
    int32_t nx, ny;      /* srf_flx.shape == (ny, nx)                                   :94      */
    int32_t px, py;      /* int(halo/dx), int(halo/dy)                                  :112-113 */
    int32_t nxe, nye;    /* padded grid nx+2px, ny+2py                                  :119-120 */
    int32_t nlx, nly;    /* retained modes AFTER the clamp (both set to nxe,nye)        :122-127 */
    int32_t nfx, nfy;    /* size of the back-transform: nl + 2*((ne-nl)//2)             :130,269-278 */
    int32_t clamped;     /* 1 if the clamp of :122-127 fired (caller logs the warning)           */
    int32_t reserved;
    double  dx, dy;      /* xmax/nx, ymax/ny                                            :98      */
    double  halo;        /* resolved halo [m] (max(xmax,ymax) when None)                :108-109 */
    double  xmax, ymax;
} bldfm_geometry;

/* One met/tower condition: the (z, profiles, meas_pt, srf_bg_conc) arguments of
 * steady_state_transport_solver (src/bldfm/solver.py:16-30).  All arrays are HOST, length nz. */
typedef struct bldfm_problem {
    const double *z;
    const double *u, *v, *Kx, *Ky, *Kz;
    int32_t nz;
    int32_t reserved;
    double  xm, ym;          /* meas_pt     */
    double  srf_bg_conc;     /* srf_bg_conc */
} bldfm_problem;

/* Timings of the most recent solve on a plan, CUDA-event measured on the plan's stream [ms]. */
typedef struct bldfm_timings {
    double forward_ms;   /* K1-K3  pad + forward transform (0 in footprint mode) */
    double march_ms;     /* K4-K8  fused vertical march kernel                   */
    double inverse_ms;   /* K9-K11 untruncate + back-transform + crop            */
    double total_ms;     /* first kernel -> last kernel / copy                   */
} bldfm_timings;

/* ---- library ---------------------------------------------------------------------------------- */
const char *bldfm_version(void);
const char *bldfm_last_error_string(void);
int  bldfm_device_count(int *count);
/* Tuning switches (kernel selection, launch shapes; the BLDFM_B200_* / BLDFM_FFT* environment variables of
 * DESIGN.md).  The environment is read once per switch and cached; bldfm_set_option overrides a switch at
 * run time (value INT32_MIN: forget the override and the cached value), bldfm_get_option reads it. */
int  bldfm_set_option(const char *name, int32_t value);
int  bldfm_get_option(const char *name, int32_t dflt);

/* Geometry (pure host arithmetic, usable without a GPU).  halo_is_none != 0 means halo=None.
 * Returns BLDFM_ERR_ODD_MODES for odd modes.                              src/bldfm/solver.py:85-130 */
int  bldfm_geometry_init(int32_t nx, int32_t ny, double xmax, double ymax, int32_t nlx, int32_t nly,
                         int32_t halo_is_none, double halo, bldfm_geometry *out);

/* Truncated wavenumber tables lx[nlx], ly[nly] (host arithmetic, bitwise equal to the numpy
 * expressions of src/bldfm/solver.py:148-153). */
int  bldfm_wavenumbers(const bldfm_geometry *g, double *lx, double *ly);

/* 1 if the reference would return float32 fields for these arguments (precision="single", no phase
 * shift applied: src/bldfm/solver.py:177-180,254-262 and SURVEY.md A.4), else 0 (float64). */
int  bldfm_output_is_f32(int flags, double xm, double ym);

/* ---- plan: device state reused across solves (cuFFT plans, workspaces, staged tables).
 * Replaces FFTManager / get_fft_manager / reset_fft_manager (src/bldfm/fft_manager.py:12-145) and the
 * numba JIT cache behind @parallelize (src/bldfm/utils.py:95-106). */
int  bldfm_plan_create(const bldfm_geometry *g, int device, bldfm_plan **out);
int  bldfm_plan_destroy(bldfm_plan *plan);
/* cudaStream_t of the plan as an opaque pointer (for event timing / stream interop by the caller) */
void *bldfm_plan_stream(bldfm_plan *plan);
int  bldfm_plan_synchronize(bldfm_plan *plan);
/* after several BLDFM_ASYNC solves with host outputs: wait until the results of the solve BEFORE the most
 * recent one have reached the host (the most recent one may still be running) -- lets a driver post-process
 * batch k while batch k+1 computes */
int  bldfm_plan_synchronize_previous(bldfm_plan *plan);
/* number of this library's own kernels launched on the plan so far */
int64_t bldfm_plan_launch_count(const bldfm_plan *plan);
/* enable (1) / disable (0) per-stage event timing; read the last solve's numbers (synchronises) */
int  bldfm_plan_set_profiling(bldfm_plan *plan, int enabled);
int  bldfm_plan_last_timings(bldfm_plan *plan, bldfm_timings *out);
/* diagnostics: with BLDFM_B200_MARCH_TRACE=1 in the environment the march kernel records four %globaltimer
 * stamps per CTA (start, tables staged, march loop done, end); copies up to max_ctas x 4 of the last march */
int  bldfm_plan_march_trace(bldfm_plan *plan, uint64_t *host, int64_t max_ctas, int64_t *nctas);
/* bytes of device workspace currently held by the plan */
int64_t bldfm_plan_workspace_bytes(const bldfm_plan *plan);

/* ---- the hot path.  Replaces the body of steady_state_transport_solver (src/bldfm/solver.py:82-290):
 * pad/FFT/truncate, two ivp_solver marches, shooting combine, (0,0) mode, phase shift, untruncate,
 * back-transform and crop.  Grid construction (:293-298) and the disk cache (:77-80,301-302) stay on
 * the Python side.
 *   levels[nlv]  level indices; rows are filled in visit order like the reference (:352-355).
 *   srf_flx      [ny][nx] float64 (ignored and may be NULL with BLDFM_FOOTPRINT).
 *   conc, flx    [nlv][ny][nx]; float64, or float32 when bldfm_output_is_f32(). */
int  bldfm_solve(bldfm_plan *plan, const bldfm_problem *prob, const int64_t *levels, int32_t nlv,
                 const double *srf_flx, int flags, void *conc, void *flx);

/* Many (tower, met) conditions in one launch.  Replaces the loops / process pool of
 * run_bldfm_timeseries/_multitower/_parallel (src/bldfm/interface.py:141-326).  Problems whose
 * (z, profiles, srf_bg_conc) are byte-identical share one march (towers differ only by the phase
 * shift).  All problems share geometry, levels, flags and srf_flx.
 *   conc, flx    [nprob][nlv][ny][nx]. */
int  bldfm_solve_batched(bldfm_plan *plan, int32_t nprob, const bldfm_problem *probs,
                         const int64_t *levels, int32_t nlv, const double *srf_flx, int flags,
                         void *conc, void *flx);

/* Footprint-weighted measurements without moving the fields to the host: for every problem b and
 * level row r returns sum_{y,x} conc[b][r][y][x]*weight[y][x] and the same for flx -- what
 * point_measurement(f, g) (src/bldfm/utils.py:80-92) computes per footprint on the host.
 *   weight        host [ny][nx] float64 (e.g. the surface flux map)
 *   conc_w, flx_w host [nprob][nlv] float64.  With BLDFM_ASYNC the call only enqueues: conc_w/flx_w must
 *                 then be PINNED and are valid after bldfm_plan_synchronize(). */
int  bldfm_solve_batched_measure(bldfm_plan *plan, int32_t nprob, const bldfm_problem *probs,
                                 const int64_t *levels, int32_t nlv, const double *srf_flx, int flags,
                                 const double *weight, double *conc_w, double *flx_w);

/* Time aggregation without moving the fields to the host (examples/timeseries_example.py:46,
 * np.mean([r["flx"] for r in results], axis=0)): solves the batch and adds every problem's fields to
 * accumulator slot slot_of[b] (problem order; slot_of[b] < 0 skips the problem).
 *   slot_of       host [nprob] int32
 *   acc_conc/flx  DEVICE [nslots][nlv][ny][nx] float64, caller-owned (bldfm_device_alloc / bldfm_device_memset);
 *                 ordered on the plan's stream, BLDFM_ASYNC skips the final synchronisation. */
int  bldfm_solve_batched_accumulate(bldfm_plan *plan, int32_t nprob, const bldfm_problem *probs,
                                    const int64_t *levels, int32_t nlv, const double *srf_flx, int flags,
                                    const int32_t *slot_of, int32_t nslots, double *acc_conc, double *acc_flx);

/* Conditioning number kappa(z[level]) of linear shooting for one problem (pure host arithmetic, usable without
 * a GPU): what BLDFM_MARCH_AUTO compares with bldfm_auto_kappa_limit() (default 8.5, env BLDFM_B200_AUTO_KAPPA).
 * bldfm_plan_last_march_mode: arithmetic of the plan's most recent march: 0 bit-mirrored, 1 FMA-contracted,
 * 2 downward sweep. */
int    bldfm_kappa(const bldfm_geometry *g, const bldfm_problem *prob, int32_t level, double *kappa);
/* May the downward sweep (BLDFM_MARCH_SWEEP) serve this problem at output level `level`?  *ok = 1 where neither the
 * swept vector nor the determinant product can leave the binary64 range and no step can be singular (bounds at the
 * largest retained wavenumbers; pure host arithmetic).  Where it is 0 the FMA-contracted shooting march runs.
 * (A solve with several output levels asks with level 0: its sweep needs no determinant product.) */
int    bldfm_sweep_admissible(const bldfm_geometry *g, const bldfm_problem *prob, int32_t level, int32_t *ok);
double bldfm_auto_kappa_limit(void);
int    bldfm_plan_last_march_mode(const bldfm_plan *plan);

/* One OVERSIZED problem sharded by ky-slab over `nranks` GPUs (one process per GPU); replaces nothing
 * in the reference -- it scales a single steady_state_transport_solver call (src/bldfm/solver.py:16)
 * beyond one device.  float64.  In non-footprint mode every rank holds the whole source srf_flx and computes only
 * its own rows of the source spectrum (the x-pass over the ny source rows is replicated -- ~1 % of the march --
 * so the forward side needs no exchange).  Every rank calls
 *   stage1: march + x-transform of this rank's block of the half-plane rows ky in [0, nly/2] (the
 *           spectra of a real source are conjugate-symmetric; blocks of Rp = ceil((nly/2+1)/G) rows,
 *           rank r owns [r*Rp, min((r+1)*Rp, nly/2+1))); results go to the device buffers send_p/send_q
 *           laid out [nlv][dst][Rp][nx/G] complex128, ready for an all-to-all (done by the caller, e.g.
 *           torch.distributed over NCCL) -- or, when peer_p/peer_q (DEVICE arrays of G pointers into
 *           every rank's receive buffer, offset to this rank's row block) are given, straight into the
 *           peers' memory over NVLink (fused transpose);
 *   stage2: real-output y-transform of the received [nlv][G*Rp][nx/G] complex128 into the real column
 *           slabs conc_slab/flx_slab [nlv][ny][nx/G] float64 (device).
 * With BLDFM_MARCH_FULL (cross-check) every retained row is marched: blocks of nly/G rows, buffers
 * [nlv][dst][nly/G][nx/G] and [nlv][nly][nx/G], full complex transforms.
 * Both stages are enqueued on the plan's stream; BLDFM_ASYNC skips the final synchronisation. */
int  bldfm_sharded_stage1(bldfm_plan *plan, const bldfm_problem *prob, const int64_t *levels, int32_t nlv,
                          const double *srf_flx, int flags, int32_t rank, int32_t nranks, void *send_p, void *send_q,
                          void *const *peer_p, void *const *peer_q);
int  bldfm_sharded_stage2(bldfm_plan *plan, int32_t nlv, int flags, int32_t rank, int32_t nranks,
                          const void *recv_p, const void *recv_q, void *conc_slab, void *flx_slab);
/* Device-side synchronisation of the fused transpose, enqueued on the plan's stream (no host thread, no
 * collective): after stage 1, bldfm_peer_signal stores `value` (the solve's sequence number) into
 * *peer_slots[d] for every rank d -- peer_slots is a DEVICE array of nranks pointers to this rank's slot in
 * each peer's flag array (uint64, IPC-mapped); before stage 2, bldfm_peer_wait spins until all nranks slots
 * of this rank's own flag array are >= value.  A peer that never arrives within timeout_s (<= 0: 20 s) ends
 * the wait and is reported by bldfm_peer_status (1 + slot index, 0 = fine; synchronises the stream). */
int  bldfm_peer_signal(bldfm_plan *plan, void *const *peer_slots, int32_t nranks, uint64_t value);
int  bldfm_peer_wait(bldfm_plan *plan, const void *flags, int32_t nranks, uint64_t value, double timeout_s);
int  bldfm_peer_status(bldfm_plan *plan, int32_t *status);
/* CUDA IPC helpers for the fused transpose (device memory from bldfm_device_alloc only). */
int  bldfm_ipc_export(void *dev_ptr, unsigned char *handle64);
int  bldfm_ipc_open(int device, const unsigned char *handle64, void **out);
int  bldfm_ipc_close(int device, void *p);

/* Spectral-stage export for parity tests: the combined, phase-shifted spectra tfftp/tfftq
 * [nlv][nly][nlx] complex128 (host) as they stand before src/bldfm/solver.py:265. */
int  bldfm_solve_spectral(bldfm_plan *plan, const bldfm_problem *prob, const int64_t *levels,
                          int32_t nlv, const double *srf_flx, int flags, double *tfftp, double *tfftq);

/* ivp_solver itself (src/bldfm/solver.py:307-374) for isolated parity tests: host arrays,
 * p0,q0,Lx,Ly [M]; outputs p_top,q_top [M] and P,Q [nlv][M] (complex128). */
int  bldfm_march(int device, int64_t M, const double *p0, const double *q0, int32_t nz,
                 const double *z, const double *u, const double *v, const double *Kx,
                 const double *Ky, const double *Kz, int32_t nlv, const int64_t *levels,
                 const double *Lx, const double *Ly, int flags,
                 double *p_top, double *q_top, double *P, double *Q);

/* Test hook for the thread -> Fourier-mode map of the march kernel (bldfm_b200/csrc/march.cuh), evaluated on
 * the host: for a launch over the rows [row0, row0+rows) -- of the half-plane ky <= nly/2 when half_plane != 0,
 * of all nly rows otherwise -- adds 1 to count[ky*nlx + kx] (host, [nly][nlx] int32, caller-zeroed) for every mode
 * the launch writes, directly or as the conjugate mirror; *nthreads = threads of the launch.  The modes of
 * ivp_solver (src/bldfm/solver.py:158-162,221) must each be written exactly once. */
int  bldfm_march_coverage(const bldfm_geometry *g, int32_t row0, int32_t rows, int32_t half_plane, int32_t *count,
                          int64_t *nthreads);

/* ---- memory helpers (so that callers need no other CUDA binding) */
int  bldfm_host_alloc(int64_t bytes, void **out);      /* pinned host memory */
int  bldfm_host_free(void *p);
int  bldfm_device_alloc(int device, int64_t bytes, void **out);
int  bldfm_device_free(int device, void *p);
int  bldfm_device_memset(int device, void *p, int value, int64_t bytes);
/* page-lock caller memory (e.g. a shared-memory segment the ranks of a node deliver their results into) */
int  bldfm_host_register(void *p, int64_t bytes);
int  bldfm_host_unregister(void *p);
int  bldfm_memcpy_d2h(int device, void *dst_host, const void *src_dev, int64_t bytes);
int  bldfm_memcpy_h2d(int device, void *dst_dev, const void *src_host, int64_t bytes);

/* FP64 pipe micro-benchmark used for the roofline denominator: runs `iters` dependent-chain
 * DFMA (fma != 0) or DMUL/DADD (fma == 0) operations per thread on a full grid and returns the
 * achieved rate in G(FMA-or-op)/s.  One FMA counts as ONE here; callers multiply by 2 for flops. */
int  bldfm_fp64_peak(int device, int fma, int iters, double *gops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* BLDFM_B200_H */
