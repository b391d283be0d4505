/*
 * b200fdtd.h — C-ABI of the B200-native FDTD time-stepping engine (libb200fdtd.so).
 *
 * This is the drop-in boundary underneath the CSXCAD/openEMS-style Python surface that
 * the reference's solver backends call.  The reference has no FFI of its own: the seam
 * is `FDTD.Run(sim_path, ...)` / `nf2ff.CalcNF2FF(...)` / `port.CalcPort(...)`
 *   antenna_sim/solver_fdtd_openems_microstrip_3d.py:82-93,176-179,214,225
 *   antenna_sim/solver_fdtd_openems_microstrip_multi_3d.py:287-292,541,564,610,621
 *   antenna_sim/solver_fdtd_openems_microstrip.py:406-416 (CalcPort / S11)
 * Every entry point below names the openEMS engine facility that `Run` used to provide
 * and that the shim (fdtd-solver-antennas_b200/openEMS) now binds through ctypes.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch types.  "dev" pointers are CUDA device
 *    pointers (torch tensor .data_ptr()); "host" pointers are ordinary host memory and
 *    are copied during the call.
 *  - every function returns 0 on success, non-zero on failure; the message is
 *    available from b200fdtd_last_error() (thread local).  Nothing aborts.
 *  - volumetric arrays are fp32, laid out [3][nz+2][ny][px]: component, z-plane,
 *    y-row, x (fastest).  Plane 0 and plane nz+1 are ghost planes (z-slab halos; zero
 *    on a global boundary), px >= nx is the row pitch in floats (multiple of 4).
 *    lin(c,k,j,i) = ((c*(nz+2) + k+1)*ny + j)*px + i   with k in [-1, nz].
 *  - state: volt = edge voltages E·dl, curr = edge currents H·dl~ (openEMS convention);
 *    update equations: SURVEY.md App. A1.
 */
#ifndef B200FDTD_H
#define B200FDTD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200fdtd_ctx b200fdtd_ctx;

/* last error message of the calling thread ("" if none) */
const char* b200fdtd_last_error(void);
/* library/ABI version (bumped when a signature changes) */
int b200fdtd_version(void);
/* number of kernels this library has launched since load (all contexts) */
int64_t b200fdtd_launch_count(void);

/* ---- context ------------------------------------------------------------------ */
/* Engine instance for one z-slab on one GPU.  `stream` is a cudaStream_t; every kernel and
 * copy of this context is issued on it (NULL = the default stream, which is also
 * torch's default stream).  Replaces openEMS::SetupFDTD's engine
 * allocation inside FDTD.Run (…microstrip_3d.py:214). */
int b200fdtd_create(b200fdtd_ctx** out, int device, int nx, int ny, int nz, int px, void* stream);
int b200fdtd_destroy(b200fdtd_ctx* ctx);

/* volt/curr [3][nz+2][ny][px] (dev, read-write) */
int b200fdtd_bind_fields(b200fdtd_ctx* ctx, float* volt, float* curr);
/* vv,vi,ii,iv [3][nz+2][ny][px] (dev, read-only): the openEMS operator (App. A2) */
int b200fdtd_bind_coeffs(b200fdtd_ctx* ctx, const float* vv, const float* vi,
                         const float* ii, const float* iv);
/* Row compression of the operator (optional, lossless).  For the E pass (which = 0: vv, vi) or the H pass
 * (which = 1: ii, iv): xvecs = dev [nvec][px] table of x-vectors; meta = dev, writable, [(nz+2)*ny] records of 32 bytes
 *   { float scale[6]; uint8 vec_id[6]; uint8 pad[2]; }   slots: ca_x, ca_y, ca_z, cb_x, cb_y, cb_z  (ca = vv|ii, cb = vi|iv)
 * vec_id 255 = row is streamed from the full array.  A row with vec_id < nvec claims full[i] == fl32(scale*xvec[i]);
 * the library verifies every claim on the device and demotes rows that do not match bit for bit, so results never
 * depend on the compression.  The full arrays must stay bound.  nvec = 0 switches compression off. */
int b200fdtd_set_row_compression(b200fdtd_ctx* ctx, int which, int nvec, const float* xvecs, void* meta,
                                 int64_t* n_compressed /*host out, may be NULL*/, int64_t* n_demoted /*host out, may be NULL*/);
/* Expansion of a compressed operator into the bound (writable) full arrays of one pass: full[i] = fl32(scale*xvec[i]) for
 * every compressed row slot, 0 for slots with vec_id 255 (the caller writes those rows itself afterwards, then calls
 * b200fdtd_set_row_compression, which verifies the result).  Lets a host ship the operator in its compressed form
 * (~0.1 % of the bytes) and build the full arrays at HBM speed. */
int b200fdtd_expand_rows(b200fdtd_ctx* ctx, int which, int nvec, const float* xvecs /*dev*/, const void* meta /*dev*/);
/* tuning knobs of the volume kernels: planes marched per CTA (kz), rows per CTA (ty in {2,4,8}),
 * variant bits: 1 = PML slabs by the separate pre/post kernel instead of fused rows, 2 = no side stream,
 * 4 = ignore the row compression, 8 = narrow x-slabs by the separate kernel, 16 = ignore the PML slab compression,
 * 32 = slab side stream at highest priority, 64 = one side stream for all slab launches, 128 = no fused H->E launches,
 * bits 8-12 = rows per CTA of the slab launches (0 = ty), bits 16-17 = L2 prefetch distance of the plain fusion
 * (0 = default 1, 3 = off), 512 = plain fusion (update_he_kernel) instead of the TMA-staged update_he6_kernel,
 * bit 20 = high-end H / low-end E slab launches beside the fused launch instead of before / after it,
 * bit 23 = whole-row PML slabs inside the fused H->E launch (double-buffered current flux; bit-exact, measured slower than
 * their own launches on B200, hence off by default), bit 22 = never (overrides bit 23),
 * bit 24 = Mur edges by index lists even where they form long arithmetic runs, bit 25 = per-row 1-D TMA copies instead of
 * tiled copies in the fused launch, bit 26 = one launch per PML slab instead of one for all, bit 27 = one fused span per
 * sampling interval instead of one per b200fdtd_run call, bit 28 = one launch per boundary plane of a z-slab rank */
int b200fdtd_set_tuning(b200fdtd_ctx* ctx, int kz, int ty, int variant);

/* ---- excitation (openEMS Engine_Ext_Excitation::Apply2Voltages; AddLumpedPort's
 *      soft E-field source, …microstrip_3d.py:176; SetGaussExcite :83) ------------- */
/* n edges: volt[idx[e]] += amp[e] * signal[ts - delay[e]]  (0 outside [0,siglen)).
 * idx/amp/delay/signal are host arrays. */
int b200fdtd_set_excitation(b200fdtd_ctx* ctx, int64_t n, const int64_t* idx, const float* amp,
                            const int32_t* delay, const float* signal, int32_t siglen);

/* ---- Mur 1st-order ABC (SetBoundaryCond(['MUR']*6), …microstrip_3d.py:84-90) ---- */
/* n boundary edges (host arrays): dst = linear index of the boundary edge, src = its
 * inward neighbour, coeff = (c dt - d)/(c dt + d). */
int b200fdtd_set_mur(b200fdtd_ctx* ctx, int64_t n, const int64_t* dst, const int64_t* src,
                     const float* coeff);

/* ---- PML_8 split-flux UPML slabs (SetBoundaryCond(['PML_8']*6)) ------------------ */
typedef struct {
    int32_t x0, y0, z0;            /* start of the box; z0 in local plane coordinates   */
    int32_t bx, by, bz;            /* extent of the box in cells                         */
    float* flux_v;                 /* dev [3][bz][by][bx] voltage flux (zero-initialised) */
    float* flux_i;                 /* dev [3][bz][by][bx] current flux                   */
    const float* vv;               /* dev [3][bz][by][bx] 2nd-stage self coefficient     */
    const float* vvfo;             /* dev … old-flux coefficient                          */
    const float* vvfn;             /* dev … new-flux coefficient                          */
    const float* ii;               /* dev same three for the currents                    */
    const float* iifo;
    const float* iifn;
    /* optional lossless row compression of the slab coefficients (like b200fdtd_set_row_compression): x-vector
     * tables dev [nvec][bx] and writable records dev [bz*by][48 bytes] = { float scale[9]; uint8 vec_id[9]; pad[3] },
     * slots a_x,a_y,a_z, fo_xyz, fn_xyz (a = vv|ii ...); vec_id 255 = stream the row from the full array.  Claims are
     * verified on the device; needs bx % 4 == 0.  Leave NULL / 0 to stream everything. */
    const float* xvecs_v; void* meta_v; int32_t nvec_v;
    const float* xvecs_i; void* meta_i; int32_t nvec_i;
} b200fdtd_pml_box;
int b200fdtd_set_pml(b200fdtd_ctx* ctx, int nboxes, const b200fdtd_pml_box* boxes /*host*/);

/* slab rows (row x slot) whose compression was accepted / demoted by the device-side check in the last set_pml */
int b200fdtd_pml_compression_info(b200fdtd_ctx* ctx, int64_t* rows_compressed, int64_t* rows_demoted);

/* ---- probes: voltage line integrals / current loop integrals, their time series
 *      and a running DFT (AddLumpedPort's port_ut/port_it probes; CalcPort,
 *      …microstrip.py:408-413) ------------------------------------------------------ */
/* nprobes probes; probe p sums weight[e]*field[idx[e]] for e in [offset[p],offset[p+1])
 * where field = volt (kind 0, time stamp ts*dt) or curr (kind 1, time stamp (ts+1/2)*dt).
 * Sampled every `interval` steps.  series: dev [nprobes][max_samples] f32.
 * dft: dev [nprobes][nfreq][2] f32 (re,im) running sum of x(t) exp(-j 2 pi f t)
 * (may be NULL with nfreq 0).  kind/offset/idx/weight/freqs are host arrays. */
int b200fdtd_set_probes(b200fdtd_ctx* ctx, int nprobes, const int32_t* kind, const int64_t* offset,
                        const int64_t* idx, const float* weight, int interval, int max_samples,
                        float* series, int nfreq, const double* freqs, float* dft, double dt);

/* ---- NF2FF Huygens box: running DFT of node-interpolated tangential E and H
 *      (CreateNF2FFBox, …microstrip_3d.py:179; CalcNF2FF :225) ---------------------- */
typedef struct {
    int32_t normal;                /* 0,1,2 = x,y,z                                       */
    int32_t plane;                 /* line index of the face along `normal` (local for z) */
    int32_t a0, a1;                /* inclusive node range along axis (normal+1)%3        */
    int32_t b0, b1;                /* inclusive node range along axis (normal+2)%3        */
    float* acc;                    /* dev [4][nfreq][nb][na][2] f32: Ea,Eb,Ha,Hb          */
} b200fdtd_nf2ff_face;
/* inv_len[axis] (host, n_axis floats each): 1/primal edge length; inv_dual: 1/dual
 * edge length.  For z both arrays are local to the slab and carry nz+2 entries with
 * entry 0 describing plane -1. */
int b200fdtd_set_nf2ff(b200fdtd_ctx* ctx, int nfaces, const b200fdtd_nf2ff_face* faces /*host*/,
                       int nfreq, const double* freqs /*host*/, int interval, double dt,
                       const float* inv_len_x, const float* inv_len_y, const float* inv_len_z,
                       const float* inv_dual_x, const float* inv_dual_y, const float* inv_dual_z);

/* Time-domain store of the face samples (optional; what openEMS dumps to nf2ff_E_n.h5 / nf2ff_H_n.h5 and CalcNF2FF reads
 * back, so that the far field can be asked for at ANY frequency after the run: …microstrip_3d.py:225 passes the caller's
 * frequency_hz).  td[q] = dev [max_samples][4][nb][na] f32 for face q of the last b200fdtd_set_nf2ff (same order); sample
 * s is taken at step (s+1)*interval.  The running DFT at the registered frequencies continues unchanged. */
int b200fdtd_set_nf2ff_td(b200fdtd_ctx* ctx, int nfaces, float* const* td /*host array of dev pointers*/, int max_samples);
/* DFT of the first `nsamples` stored samples of one face at nfreq frequencies (host array):
 * out = dev [4][nfreq][nb][na][2] f32, the layout and scaling of the running-DFT accumulators (sum x(t) exp(-j 2 pi f t),
 * E stamped ts*dt, H stamped (ts+1/2)*dt).  Synchronises the stream. */
int b200fdtd_nf2ff_td_dft(b200fdtd_ctx* ctx, int face, int nfreq, const double* freqs /*host*/, int nsamples, float* out);

/* ---- time stepping (FDTD.Run, …microstrip_3d.py:214) ---------------------------- */
/* current time-step counter (number of completed steps) */
int b200fdtd_get_timestep(b200fdtd_ctx* ctx, int64_t* ts);
int b200fdtd_set_timestep(b200fdtd_ctx* ctx, int64_t ts);
/* run n full steps (pre/post extensions, E update, excitation, H update, sampling).
 * use_graph != 0 replays a captured CUDA graph of one sampling interval. */
int b200fdtd_run(b200fdtd_ctx* ctx, int64_t nsteps, int use_graph);
/* split phases for z-slab halo exchange (host drives the exchange between them):
 *   phase 0: everything up to and including Apply2Voltages (E half step)
 *   phase 1: current half step; increments the step counter
 *   phase 2: probe / NF2FF sampling if the new step count is a multiple of the interval
 *            (after the H halo so slab-boundary nodes see their neighbour's new values)
 *   phase 3: the same between two fused steps (b200fdtd_fused_step_part), where the E update of the next step has
 *            already been done: the voltages are taken from the field copy that is NOT current (the fused launch only
 *            read it), the currents from the current one */
int b200fdtd_half_step(b200fdtd_ctx* ctx, int phase);
/* the same half steps cut in two so the halo exchange hides behind the interior launch:
 *   phase 0: part 0 = pre passes + planes [1,nz),   part 1 = plane 0 (reads the lower ghost H) + post passes
 *   phase 1: part 0 = pre passes + planes [0,nz-1), part 1 = plane nz-1 (reads the upper ghost E) + post, ++ts */
int b200fdtd_half_step_part(b200fdtd_ctx* ctx, int phase, int part);
/* Fused H->E steps on a z-slab rank (see b200fdtd_run for the single-slab case).  The caller owns the second copy of the
 * fields ([3][nz+2][ny][px] each, zero-initialised) and does the halo exchange on whichever copy is current (ghost planes
 * that are never exchanged must hold the same values in both copies).  One step = parts 0..3 in order:
 *   0: Mur pre, H of plane 0 and of the interior PML slabs   [4: optional, fused launch over the lower half of the
 *   interior planes, hides the wait for the upper ghost E]   1: H of the top plane (after the upper ghost E has arrived;
 *   then send H_new(top) up from the NOT yet current H copy)   2: fused launch over planes [1,nz-1), H copy flips, ++ts
 *   3: E of the interior PML slabs, plane 0 (after the lower ghost H_new has arrived in the current H copy) and top
 *      plane, E copy flips, Mur post / excitation / Mur apply (then send E_new(0) down from the current E copy).
 * The step starts from (E(n), H(n-1)) and ends at (E(n+1), H(n)): bracket a span of steps with half steps
 * b200fdtd_half_step_part(0, .) and (1, .), which work on the current copies. */
int b200fdtd_bind_alt_fields(b200fdtd_ctx* ctx, float* volt2, float* curr2);
int b200fdtd_fused_step_part(b200fdtd_ctx* ctx, int part);
/* which copy holds E / H now: 0 = the arrays of b200fdtd_bind_fields, 1 = the second copy */
int b200fdtd_current_copy(b200fdtd_ctx* ctx, int* vcur, int* ccur);
/* the caller has copied the state back into the bound arrays: copy 0 is current again */
int b200fdtd_reset_current_copy(b200fdtd_ctx* ctx);
/* only the volume kernels (bench / roofline): which = 0 E update, 1 H update (plain + fused PML slab launches);
 * 2 / 3 = only the plain launch of the E / H update (rows outside the fused PML slabs);
 * 4 = only the fused H->E launch over the plain region (writes the second field copy: the state is untouched) */
int b200fdtd_update_only(b200fdtd_ctx* ctx, int which);
/* how the volume is split: cells (incl. pad columns) swept by the plain launch, by the fused PML slab launches, and
 * by the separate PML pre/post kernel */
int b200fdtd_plan_info(b200fdtd_ctx* ctx, int64_t* plain_cells, int64_t* fused_cells, int64_t* separate_cells);
/* Fused H->E launches (b200fdtd_run on one slab, graph or eager): the H update of step n and the E update of step n+1 of the plain region
 * are done in one sweep that writes a second copy of the fields (allocated by the library, 24 B/cell; the run falls
 * back to the separate E and H launches if that allocation fails, if a PML box is not slab-shaped, or with
 * variant bit 128).  Results are identical either way.  rows in {3,7,15} = rows per CTA, planes >= 1 = planes marched
 * per CTA (0 keeps the current value); bits 24-25 of `planes` = E planes landing ahead of the two in use (1 | 2; 0 = as many
 * as shared memory allows without costing a resident CTA). */
int b200fdtd_set_he_tuning(b200fdtd_ctx* ctx, int rows, int planes);
/* *active = 1 if the last b200fdtd_run used fused H->E launches */
int b200fdtd_he_info(b200fdtd_ctx* ctx, int* active);
/* openEMS CalcFastEnergy: 0.5*eps0*sum(volt^2) + 0.5*mu0*sum(curr^2) over owned planes
 * (synchronises the stream) */
int b200fdtd_energy(b200fdtd_ctx* ctx, double* energy);
/* wait for the context's stream */
int b200fdtd_sync(b200fdtd_ctx* ctx);
/* number of probe samples taken so far */
int b200fdtd_num_samples(b200fdtd_ctx* ctx, int* n);

/* ---- equivalent currents of the Huygens box (first half of nf2ff.CalcNF2FF, …microstrip_3d.py:225) ------------------ */
/* One face of the box with its spectra at ONE frequency resident on the device: component c (Ea, Eb, Ha, Hb: the two
 * tangential axes a = (normal+1)%3, b = (normal+2)%3) at acc + c*comp_stride floats, each [nb][na][2] f32 (re, im).
 * xa/xb = node coordinates along a / b (host, metres), wa/wb = trapezoid weights (host), coord = position of the face along
 * its normal, side = 0 lower / 1 upper face (outward normal -n / +n). */
typedef struct {
    int32_t normal, side, na, nb;
    double coord;
    const float* acc; int64_t comp_stride;
    const double *xa, *xb, *wa, *wb;
} b200fdtd_nf2ff_src_face;
/* Fills pos[3][npts], J[3][npts][2], M[3][npts][2] (dev f32, npts = sum na*nb, faces in order) with the node positions
 * relative to `center` and J = n x H, M = -n x E times dA times `scale` (the DFT normalisation 2*dt_sample), and returns
 * Prad = 1/2 Re sum (E x H*) . n dA * scale^2 in *prad (host).  Synchronises the stream. */
int b200fdtd_nf2ff_sources(int device, void* stream, int nfaces, const b200fdtd_nf2ff_src_face* faces /*host*/, double scale,
                           const double* center /*host [3]*/, float* pos, float* J, float* M, double* prad /*host out*/);

/* ---- far field (nf2ff.CalcNF2FF radiation integral, …microstrip_3d.py:225) ------ */
/* npts surface points with equivalent currents (dev, f32 SoA arrays of npts):
 *   pos[3][npts], J[3][npts][2], M[3][npts][2] (already multiplied by dA).
 * For each of ndir directions (theta,phi host arrays, radians) computes
 *   N = sum J e^{+j k r^.r'},  L = sum M e^{+j k r^.r'}  projected on theta^,phi^
 * out: dev [ndir][4][2] f32 = (N_theta, N_phi, L_theta, L_phi). */
int b200fdtd_farfield(int device, void* stream, int64_t npts, const float* pos, const float* J,
                      const float* M, double k, int ndir, const double* theta, const double* phi,
                      float* out);

#ifdef __cplusplus
}
#endif
#endif /* B200FDTD_H */
