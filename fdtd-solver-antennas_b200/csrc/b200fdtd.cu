// b200fdtd.cu — hand-written sm_100a kernels + C-ABI of the B200-native FDTD engine.
//
// Replaces the time-stepping core that the reference hands to openEMS through
// FDTD.Run / CalcNF2FF / CalcPort (antenna_sim/solver_fdtd_openems_microstrip_3d.py:214,225;
// antenna_sim/solver_fdtd_openems_microstrip.py:408-413).  Equations: SURVEY.md App. A.
//
// Data layout (see include/b200fdtd.h): fp32 [3][nz+2][ny][px], x fastest, z slowest, one
// ghost plane below and above the owned z-slab.  All volume kernels stream whole x-rows
// with 16-byte accesses, march along z keeping the z-neighbour plane in registers, take
// the x-neighbour from the adjacent lane with a warp shuffle and the y-neighbour row from
// L1/L2 (it is the row another warp of the same CTA streams at the same time).
//
// Arithmetic contract (bit-exact with oracle/fdtd_ref.c):
//   curl = ((a - b) - c) + d          (three rounded fp32 adds, this order)
//   f    = fmaf(ca, f, cb * curl)     (one rounded multiply, one fused multiply-add)
#include <cuda_runtime.h>
#include <cuda.h>               // CUtensorMap (the encoder itself comes from the driver through cudaGetDriverEntryPoint: no -lcuda)
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <stdlib.h>
#include <vector>
#include <atomic>

#include "b200fdtd.h"

// ------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

static int fail(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return 1;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define CKL() do { g_launches.fetch_add(1, std::memory_order_relaxed); cudaError_t e_ = cudaGetLastError(); \
    if (e_ != cudaSuccess) return fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" const char* b200fdtd_last_error(void) { return g_err; }
extern "C" int b200fdtd_version(void) { return 3; }
extern "C" int64_t b200fdtd_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------
struct PmlBoxDev {
    int x0, y0, z0, bx, by, bz;
    long long start;               // first flat element of this box in the concatenated space
    float *flux_v, *flux_i;
    const float *vv, *vvfo, *vvfn, *ii, *iifo, *iifn;
    const float *xv_v, *xv_i;                  // row compression of the slab coefficients (optional)
    const unsigned char *meta_v, *meta_i;
};
#define MAX_PML_BOXES 8
struct PmlTable { int n; long long total; PmlBoxDev b[MAX_PML_BOXES]; };

// launch plan of the volume kernels: PML boxes spanning whole x-rows are fused into the volume launches
struct FusedBox { int y0, by, z0, bz; float *flux_v, *flux_i; const float *vv, *vvfo, *vvfn, *ii, *iifo, *iifn;
                  const float *xv_v, *xv_i; const unsigned char *meta_v, *meta_i;
                  float* flux_i_alt; };                      // second copy of the current flux (library-owned, see flux_pp)
struct VolumePlan {
    bool valid = false;
    int nseg = 0; int seg0[3], seg1[3];          // plane ranges of the plain launches (complement of fused z-slabs)
    int nskip = 0; int sj0[2], sj1[2];           // rows of fused y-slabs
    int nfused = 0; FusedBox fb[4];
    // narrow x-slabs folded into "x-edge" launches of the plain rows
    bool xedge = false; int has_lo = 0, has_hi = 0; FusedBox xlo{}, xhi{}; int xw0 = 0, xx1 = 0, xw1 = 0;
    int ym0 = 0, ym1 = 0;                        // rows of the plain launches (complement of the fused y-slabs)
};

struct FaceDev {
    int normal, plane, a0, a1, b0, b1;
    float* acc;
    float* td;                     // optional time-domain store [td_max][4][nb][na] (b200fdtd_set_nf2ff_td)
};
#define MAX_FACES 8
struct FaceTable { int n; FaceDev f[MAX_FACES]; };

struct b200fdtd_ctx {
    int device = 0;
    int nx = 0, ny = 0, nz = 0, px = 0;
    long long sz = 0, cs = 0;      // plane stride, component stride (floats)
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side = nullptr;           // runs the fused PML slab launches concurrently with the plain launch
    cudaStream_t side_lo = nullptr, side_hi = nullptr;   // the same at default / highest priority (variant bit 32 picks)
    cudaStream_t side2 = nullptr;          // whole-row slab launches (z/y) while `side` runs the narrow x-slab launch
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
    cudaStream_t slab_s[3] = {nullptr, nullptr, nullptr};   // one stream per further whole-row slab: the slab launches are short and
    cudaEvent_t slab_ev[3] = {nullptr, nullptr, nullptr};   // latency-bound, so they run side by side rather than one after the other
    float *volt = nullptr, *curr = nullptr;
    // second copy of the fields for the fused H->E launches (they cannot update in place: a CTA recomputes the halo of its
    // tile from the old values its neighbours are overwriting).  vcur/ccur say which copy holds the current E / H; both are
    // 0 whenever control returns to the caller, so the bound arrays always hold the state.
    float *alt_volt = nullptr, *alt_curr = nullptr;
    int he_mid = 0;                        // z-slab fused step: plane where the upper half of the fused launch starts (0 = whole)
    bool alt_owned = false;                // allocated by the library (else bound by the caller, b200fdtd_bind_alt_fields)
    int vcur = 0, ccur = 0;
    bool flip = false;                     // volume launches write the other copy instead of updating in place
    int he_ty = 7, he_kz = 32;             // fused launch: rows per CTA (+1 halo row), planes marched per CTA
    int he_de = 0;                         // E planes landing ahead of the two in use (0 = as many as shared memory allows, max 2)
    // The fused launch sweeps the whole-row PML slabs too and recomputes H_new on tile halos from the OLD current flux, so
    // the current flux of those slabs is double buffered: every H pass reads copy fcur and writes the other one.
    bool flux_pp = false; int fcur = 0;
    float* flux_alt[4] = {nullptr, nullptr, nullptr, nullptr};
    const float *vv = nullptr, *vi = nullptr, *ii = nullptr, *iv = nullptr;
    int kz = 16, ty = 4, variant = 0;
    const float* cmp_xv[2] = {nullptr, nullptr};          // row compression tables of the E and H pass (caller-owned)
    const unsigned char* cmp_meta[2] = {nullptr, nullptr};
    int cmp_nvec[2] = {0, 0};
    // step counter
    int64_t ts = 0;
    int* d_ts = nullptr;
    int ts_lag = 0;                        // steps counted on the host but not yet added to the device counter (host-driven
                                           // z-slab steps pass the lag as an offset instead of launching ts_add every step)
    // excitation
    int64_t n_exc = 0; int64_t* exc_idx = nullptr; float* exc_amp = nullptr; int* exc_delay = nullptr;
    float* exc_sig = nullptr; int exc_siglen = 0;
    // mur
    int64_t n_mur = 0; int64_t *mur_dst = nullptr, *mur_src = nullptr; float *mur_coeff = nullptr, *mur_tmp = nullptr;
    struct MurSeg* mur_segs = nullptr; int n_mur_segs = 0;   // the same lists as arithmetic runs (used when the runs are long)
    // pml: all boxes as given, the boxes left to the separate pre/post passes, and the fused launch plan
    int64_t pml_rows_compressed = 0, pml_rows_demoted = 0;
    PmlTable pml_all{};
    PmlTable pml{};
    VolumePlan plan{};
    // probes
    int n_probes = 0; int* pr_kind = nullptr; int64_t* pr_off = nullptr; int64_t* pr_idx = nullptr; float* pr_w = nullptr;
    int interval = 0; int max_samples = 0; float* pr_series = nullptr; int pr_nfreq = 0; double* pr_freqs = nullptr;
    float* pr_dft = nullptr; double dt = 0.0;
    // nf2ff
    FaceTable faces{}; int nf_nfreq = 0; double* nf_freqs = nullptr; int nf_interval = 0; double nf_dt = 0.0;
    float* inv_len[3] = {nullptr, nullptr, nullptr}; float* inv_dual[3] = {nullptr, nullptr, nullptr};
    int nf_max_nodes = 0; int nf_td_max = 0;
    // energy
    double* d_partials = nullptr; int n_partials = 0; double* d_energy = nullptr;
    // graph
    cudaGraphExec_t graph = nullptr; int graph_steps = 0; int64_t graph_kernels = 0;
    // pipelined stepping (run_pipelined): one graph of `pgraph_steps` fused steps per parity of the field copies at its entry
    cudaGraphExec_t pgraph[4] = {nullptr, nullptr, nullptr, nullptr}; int pgraph_steps = 0; int64_t pgraph_kernels[4] = {0, 0, 0, 0};
    bool he_fused = false;                 // the last b200fdtd_run used fused H->E launches
    bool graph_fused = false;              // ... and so does the captured chunk
    // device copies of the slab / face tables
    PmlTable* d_pml = nullptr; FaceTable* d_faces = nullptr;
};

static float* cur_volt(const b200fdtd_ctx* c) { return c->vcur ? c->alt_volt : c->volt; }
static float* oth_volt(const b200fdtd_ctx* c) { return c->vcur ? c->volt : c->alt_volt; }
static float* cur_curr(const b200fdtd_ctx* c) { return c->ccur ? c->alt_curr : c->curr; }
static float* oth_curr(const b200fdtd_ctx* c) { return c->ccur ? c->curr : c->alt_curr; }

static int sample_interval(const b200fdtd_ctx* c) {
    if (c->n_probes > 0 && c->interval > 0) return c->interval;
    if (c->faces.n > 0 && c->nf_interval > 0) return c->nf_interval;
    return 0;
}

#include "kernels_volume.cuh"      // K1/K2 volume updates
#include "kernels_fused.cuh"       // fused H->E launch (all generations)

static void fill_vol_params(const b200fdtd_ctx* c, int which, VolParams& p)
{
    p.fin = which == 0 ? cur_volt(c) : cur_curr(c);
    p.f = c->flip ? (which == 0 ? oth_volt(c) : oth_curr(c)) : const_cast<float*>(p.fin);
    p.g = which == 0 ? cur_curr(c) : cur_volt(c);
    p.ca = which == 0 ? c->vv : c->ii;
    p.cb = which == 0 ? c->vi : c->iv;
    p.nx = c->nx; p.ny = c->ny; p.nz = c->nz; p.px = c->px; p.sz = c->sz; p.cs = c->cs;
    p.kz = 1; p.kspan = 0; p.k0 = 0; p.k1 = 0;
    p.xv = c->cmp_xv[which]; p.meta = c->cmp_meta[which];
}

template <int MODE>
static int launch_volume_one(b200fdtd_ctx* c, int which, int k0, int k1, const RowParams& r, cudaStream_t stream, int kz, int nchunks, int grid_y = 0,
                             SlabSet* collect = nullptr)
{
    if (k1 <= k0 || r.j1 <= r.j0 || nchunks <= 0) return 0;
    if (collect != nullptr && MODE != 0) {                  // merged slab launch: only note what this launch would have been
        if (collect->n >= MAX_SLABS) return fail("too many PML slabs for one merged launch");
        const int ty_s = ((c->variant >> 8) & 31) ? ((c->variant >> 8) & 31) : c->ty;
        SlabEntry& E = collect->e[collect->n++];
        E.r = r; E.mode = MODE; E.gx = nchunks; E.gy = grid_y > 0 ? grid_y : (r.j1 - r.j0 + ty_s - 1) / ty_s; E.gz = (k1 - k0 + kz - 1) / kz;
        E.kz = kz; E.k0 = k0; E.k1 = k1; E.cta0 = 0;
        return 0;
    }
    VolParams p; fill_vol_params(c, which, p);
    p.kz = kz; p.k0 = k0; p.k1 = k1;
    // slab launches may use their own CTA height (variant bits 8-12): a 2-row slab CTA has the register footprint of one
    // plain CTA, so it fits the slot a retiring plain CTA frees when both run concurrently
    const int ty_slab = (c->variant >> 8) & 31;
    const int ty = (MODE != 0 && ty_slab) ? ty_slab : c->ty;
    dim3 block(32, ty);
    dim3 grid(nchunks, grid_y > 0 ? grid_y : (r.j1 - r.j0 + ty - 1) / ty, (k1 - k0 + kz - 1) / kz);
    if (grid.y > 65535 || grid.z > 65535) return fail("grid too large for launch (ny/ty=%u, nz/kz=%u)", grid.y, grid.z);
    const bool cmp = c->cmp_meta[which] != nullptr && (c->variant & 4) == 0;
#define LAUNCH(TYV) do { \
        if (which == 0) { if (cmp) update_e_kernel<TYV, MODE, true><<<grid, block, 0, stream>>>(p, r); \
                          else update_e_kernel<TYV, MODE, false><<<grid, block, 0, stream>>>(p, r); } \
        else { if (cmp) update_h_kernel<TYV, MODE, true><<<grid, block, 0, stream>>>(p, r); \
               else update_h_kernel<TYV, MODE, false><<<grid, block, 0, stream>>>(p, r); } } while (0)
    switch (ty) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        case 8: LAUNCH(8); break;
        case 16: LAUNCH(16); break;
        default: return fail("unsupported ty=%d", ty);
    }
#undef LAUNCH
    CKL();
    return 0;
}

// Split the PML boxes into fused slabs (whole x-rows; handled inside the volume launches) and boxes for the
// separate pre/post kernel.  z-slabs (all rows of some planes) and y-slabs (some rows of all remaining planes) qualify.
static int normalize_flux(b200fdtd_ctx* c);
static int build_plan(b200fdtd_ctx* c)
{
    if (c->plan.valid && c->flux_pp && c->fcur) if (normalize_flux(c)) return 1;   // the old plan's second flux copy holds the state
    VolumePlan P;
    PmlTable rest; memset(&rest, 0, sizeof(rest));
    const PmlTable& A = c->pml_all;
    const bool allow = (c->variant & 1) == 0;
    int kind[MAX_PML_BOXES];                      // 0 separate, 1 z-slab, 2 y-slab candidate
    int zlo = 0, zhi = c->nz;                     // complement of the fused z-slabs must stay one contiguous range
    for (int b = 0; b < A.n; ++b) {
        const PmlBoxDev& B = A.b[b];
        kind[b] = 0;
        if (!allow || B.x0 != 0 || B.bx != c->px) continue;
        if (((uintptr_t)B.flux_v | (uintptr_t)B.flux_i | (uintptr_t)B.vv | (uintptr_t)B.vvfo | (uintptr_t)B.vvfn |
             (uintptr_t)B.ii | (uintptr_t)B.iifo | (uintptr_t)B.iifn) & 15) continue;
        if (B.y0 == 0 && B.by == c->ny) {
            if (B.z0 == 0 && B.bz < c->nz && B.bz > zlo) { kind[b] = 1; }
            else if (B.z0 + B.bz == c->nz && B.z0 > 0) { kind[b] = 1; }
            else if (B.z0 == 0 && B.bz == c->nz) { kind[b] = 1; }      // the whole slab is PML
        } else if (B.y0 == 0 || B.y0 + B.by == c->ny) kind[b] = 2;     // y-slabs sit at the low or the high end of y
    }
    {   // at most four whole-row slabs are folded into the volume launches (two z, two y): decided BEFORE their rows and
        // planes are carved out of the plain launch, so a slab beyond that keeps the separate pre/post passes over cells
        // the plain launch still updates
        int nlo = 0, nhi = 0;                                // one z-slab per end; the y-slabs are limited to two below
        for (int b = 0; b < A.n; ++b) if (kind[b] == 1) {
            const bool lo = A.b[b].z0 == 0;
            if ((lo ? nlo++ : nhi++) > 0) kind[b] = 0;
        }
    }
    for (int b = 0; b < A.n; ++b) if (kind[b] == 1) {
        const PmlBoxDev& B = A.b[b];
        if (B.z0 == 0) zlo = zlo > B.bz ? zlo : B.bz;
        else zhi = zhi < B.z0 ? zhi : B.z0;
    }
    if (zhi < zlo) zhi = zlo;
    for (int b = 0; b < A.n; ++b) if (kind[b] == 2) {
        const PmlBoxDev& B = A.b[b];
        if (!(B.z0 == zlo && B.z0 + B.bz == zhi) || P.nskip >= 2) kind[b] = 0;
        else { P.sj0[P.nskip] = B.y0; P.sj1[P.nskip] = B.y0 + B.by; P.nskip++; }
    }
    // narrow x-slabs: float4-aligned column ranges [0,w) and [x0,x0+w), w <= 32, spanning exactly the plain rows and planes
    if (allow && (c->variant & 8) == 0) {
        int ym0 = 0, ym1 = c->ny;
        for (int q = 0; q < P.nskip; ++q) { if (P.sj0[q] == 0) ym0 = P.sj1[q]; else ym1 = P.sj0[q]; }
        for (int b = 0; b < A.n; ++b) {
            const PmlBoxDev& B = A.b[b];
            if (kind[b] != 0) continue;
            const bool shape = (B.x0 % 4 == 0) && (B.bx % 4 == 0) && B.bx <= 32 && B.x0 + B.bx <= c->px &&
                               B.y0 == ym0 && B.y0 + B.by == ym1 && B.z0 == zlo && B.z0 + B.bz == zhi;
            const bool aligned = !(((uintptr_t)B.flux_v | (uintptr_t)B.flux_i | (uintptr_t)B.vv | (uintptr_t)B.vvfo | (uintptr_t)B.vvfn |
                                    (uintptr_t)B.ii | (uintptr_t)B.iifo | (uintptr_t)B.iifn) & 15);
            if (!shape || !aligned) continue;
            FusedBox Fb; Fb.flux_i_alt = nullptr; Fb.y0 = B.y0; Fb.by = B.by; Fb.z0 = B.z0; Fb.bz = B.bz;
            Fb.flux_v = B.flux_v; Fb.flux_i = B.flux_i; Fb.vv = B.vv; Fb.vvfo = B.vvfo; Fb.vvfn = B.vvfn; Fb.ii = B.ii; Fb.iifo = B.iifo; Fb.iifn = B.iifn;
            Fb.xv_v = B.xv_v; Fb.xv_i = B.xv_i; Fb.meta_v = B.meta_v; Fb.meta_i = B.meta_i;
            if (B.x0 == 0 && !P.has_lo && (!P.has_hi || B.bx <= P.xx1)) { P.has_lo = 1; P.xlo = Fb; P.xw0 = B.bx; kind[b] = 3; }
            else if (B.x0 > 0 && !P.has_hi && (!P.has_lo || B.x0 >= P.xw0)) { P.has_hi = 1; P.xhi = Fb; P.xx1 = B.x0; P.xw1 = B.bx; kind[b] = 3; }
        }
        P.xedge = P.has_lo || P.has_hi;
    }
    for (int b = 0; b < A.n; ++b) {
        const PmlBoxDev& B = A.b[b];
        if (kind[b] == 3) continue;
        if (kind[b] == 0 || P.nfused >= 4) {
            PmlBoxDev D = B; D.start = rest.total; rest.b[rest.n++] = D; rest.total += 3LL * B.bx * B.by * B.bz;
        } else {
            FusedBox& Fb = P.fb[P.nfused++]; Fb.flux_i_alt = nullptr;
            Fb.y0 = B.y0; Fb.by = B.by; Fb.z0 = B.z0; Fb.bz = B.bz;
            Fb.flux_v = B.flux_v; Fb.flux_i = B.flux_i; Fb.vv = B.vv; Fb.vvfo = B.vvfo; Fb.vvfn = B.vvfn;
            Fb.ii = B.ii; Fb.iifo = B.iifo; Fb.iifn = B.iifn;
            Fb.xv_v = B.xv_v; Fb.xv_i = B.xv_i; Fb.meta_v = B.meta_v; Fb.meta_i = B.meta_i;
        }
    }
    P.nseg = 0;
    if (zhi > zlo) { P.seg0[0] = zlo; P.seg1[0] = zhi; P.nseg = 1; }
    P.ym0 = 0; P.ym1 = c->ny;
    for (int q = 0; q < P.nskip; ++q) { if (P.sj0[q] == 0) P.ym0 = P.sj1[q]; else P.ym1 = P.sj0[q]; }
    P.valid = true;
    // a new plan: the second flux copies (if any) belong to the old one
    for (int q = 0; q < 4; ++q) { if (c->flux_alt[q]) { cudaFree(c->flux_alt[q]); c->flux_alt[q] = nullptr; } }
    c->flux_pp = false; c->fcur = 0;
    c->plan = P;
    c->pml = rest;
    return 0;
}

// plain (non-PML) volume launches of one half step restricted to planes [k0,k1)
static int launch_volume_plain(b200fdtd_ctx* c, int which, int k0, int k1, cudaStream_t stream)
{
    const VolumePlan& P = c->plan;
    RowParams r; memset(&r, 0, sizeof(r));
    r.j0 = 0; r.j1 = c->ny;
    if (P.nskip > 0) { r.sj0a = P.sj0[0]; r.sj1a = P.sj1[0]; }
    if (P.nskip > 1) { r.sj0b = P.sj0[1]; r.sj1b = P.sj1[1]; }
    // columns owned by the narrow-slab launches are not stored by the plain launch
    if (P.has_lo) r.xw0 = P.xw0;
    if (P.has_hi) { r.xx1 = P.xx1; r.xw1 = P.xw1; } else { r.xx1 = 0; r.xw1 = 0; }
    for (int s = 0; s < P.nseg; ++s) {
        const int a = k0 > P.seg0[s] ? k0 : P.seg0[s], b = k1 < P.seg1[s] ? k1 : P.seg1[s];
        if (launch_volume_one<0>(c, which, a, b, r, stream, c->kz, (c->px + 127) / 128)) return 1;
    }
    return 0;
}

static int slab_ty(const b200fdtd_ctx* c) { const int t = (c->variant >> 8) & 31; return t ? t : c->ty; }
static int slots_for(int w) { return (w + 3) / 4; }

// planes marched per CTA of a thin slab launch: enough CTAs to fill the machine several times over (the march is a
// serial chain of DRAM round trips, so small launches need their parallelism from the grid), chunks of equal length
static int slab_kz(int kz, int planes, long long ctas_per_chunk)
{
    auto chunks = [&](int q) { return (planes + q - 1) / q; };
    static const long long target = [] { const char* e = getenv("B200FDTD_SLAB_CTAS"); return e ? atoll(e) : 148LL * 4; }();
    while (kz > 2 && ctas_per_chunk * chunks(kz) < target) kz = (kz + 1) / 2;
    if (kz > planes) kz = planes;
    const int n = chunks(kz);
    return (planes + n - 1) / n;
}

// narrow x-slab launches (MODE 2): PML pre/update/post on the slab columns of the plain rows
// sel: 0 = both slabs, 1 = only the slab at the low end of x, 2 = only the one at the high end
static int launch_volume_xslabs(b200fdtd_ctx* c, int which, int k0, int k1, cudaStream_t stream, int sel = 0, SlabSet* collect = nullptr)
{
    VolumePlan P = c->plan;
    if (!P.xedge) return 0;
    if (sel == 1) P.has_hi = 0;
    if (sel == 2) P.has_lo = 0;
    if (!P.has_lo && !P.has_hi) return 0;
    RowParams e; memset(&e, 0, sizeof(e));
    const FusedBox& L = P.xlo; const FusedBox& H = P.xhi;
    const FusedBox& any = P.has_lo ? L : H;
    e.j0 = any.y0; e.j1 = any.y0 + any.by;
    e.y0 = any.y0; e.z0 = any.z0; e.by = any.by; e.bz = any.bz;
    if (P.has_lo) { e.xflux0 = which == 0 ? L.flux_v : L.flux_i; e.xa0 = which == 0 ? L.vv : L.ii;
                    e.xfo0 = which == 0 ? L.vvfo : L.iifo; e.xfn0 = which == 0 ? L.vvfn : L.iifn; e.xw0 = P.xw0; e.xs0 = slots_for(P.xw0);
                    if ((c->variant & 16) == 0) { e.pxv0 = which == 0 ? L.xv_v : L.xv_i; e.pmeta0 = which == 0 ? L.meta_v : L.meta_i; if (!e.pxv0) e.pmeta0 = nullptr; } }
    if (P.has_hi) { e.xflux1 = which == 0 ? H.flux_v : H.flux_i; e.xa1 = which == 0 ? H.vv : H.ii;
                    e.xfo1 = which == 0 ? H.vvfo : H.iifo; e.xfn1 = which == 0 ? H.vvfn : H.iifn; e.xx1 = P.xx1; e.xw1 = P.xw1; e.xs1 = slots_for(P.xw1);
                    if ((c->variant & 16) == 0) { e.pxv1 = which == 0 ? H.xv_v : H.xv_i; e.pmeta1 = which == 0 ? H.meta_v : H.meta_i; if (!e.pxv1) e.pmeta1 = nullptr; } }
    const int a = k0 > any.z0 ? k0 : any.z0, b = k1 < any.z0 + any.bz ? k1 : any.z0 + any.bz;
    if (b <= a) return 0;
    // rows per CTA = ty * 32/xs: size the row grid for the slab with the fewest rows per warp
    const int xs = (P.has_lo && P.has_hi) ? (e.xs0 > e.xs1 ? e.xs0 : e.xs1) : (P.has_lo ? e.xs0 : e.xs1);
    RowParams g = e;
    g.j1 = e.j1;                                             // kernel bounds rows by j1; grid.y from the widest slab
    const int rows_per_cta = slab_ty(c) * (32 / xs);
    const int gy = (any.by + rows_per_cta - 1) / rows_per_cta;
    // the narrow-slab launch is latency-bound (one row per lane): it wants four times the CTAs of the whole-row slabs
    const int kz = slab_kz(c->kz, b - a, ((long long)gy * (P.has_lo + P.has_hi) + 3) / 4);
    // launch_volume_one derives grid.y from (j1-j0)/ty: pass an equivalent row count
    g.j0 = e.j0; 
    return launch_volume_one<2>(c, which, a, b, g, stream, kz, P.has_lo + P.has_hi, gy, collect);
}

// volume launches over the fused PML slabs (pre -> update -> post in registers)
// sel: 0 = all slabs, 1 = only the slabs at the low end of their axis (z0 = 0 / y0 = 0), 2 = only those at the high end
static int launch_volume_fused(b200fdtd_ctx* c, int which, int k0, int k1, cudaStream_t stream, int sel = 0, SlabSet* collect = nullptr)
{
    const VolumePlan& P = c->plan;
    for (int q = 0; q < P.nfused; ++q) {
        const FusedBox& B = P.fb[q];
        if (sel) {
            const bool zslab = B.y0 == 0 && B.by == c->ny;
            const bool lo = zslab ? B.z0 == 0 : B.y0 == 0;
            if ((sel == 1) != lo) continue;
        }
        RowParams f; memset(&f, 0, sizeof(f));
        f.j0 = B.y0; f.j1 = B.y0 + B.by; f.y0 = B.y0; f.z0 = B.z0; f.by = B.by; f.bz = B.bz;
        if (which == 0) { f.flux = B.flux_v; f.flux_out = B.flux_v; }
        else if (c->flux_pp) { f.flux = c->fcur ? B.flux_i_alt : B.flux_i; f.flux_out = c->fcur ? B.flux_i : B.flux_i_alt; }
        else { f.flux = B.flux_i; f.flux_out = B.flux_i; }
        f.a = which == 0 ? B.vv : B.ii; f.fo = which == 0 ? B.vvfo : B.iifo; f.fn = which == 0 ? B.vvfn : B.iifn;
        if ((c->variant & 16) == 0) { f.pxv = which == 0 ? B.xv_v : B.xv_i; f.pmeta = which == 0 ? B.meta_v : B.meta_i; if (!f.pxv) f.pmeta = nullptr; }
        const int a = k0 > B.z0 ? k0 : B.z0, b = k1 < B.z0 + B.bz ? k1 : B.z0 + B.bz;
        if (b <= a) continue;
        // thin slabs: march fewer planes per CTA so the launch still fills the machine (>= ~8 CTAs per SM)
        const long long per_chunk = (long long)((c->px + 127) / 128) * ((B.by + slab_ty(c) - 1) / slab_ty(c));
        const int kz = slab_kz(c->kz, b - a, per_chunk);
        cudaStream_t st = stream;
        if (stream == c->side2 && q > 0 && q <= 3 && (c->variant & 1024) == 0) st = c->slab_s[q - 1];
        if (launch_volume_one<1>(c, which, a, b, f, st, kz, (c->px + 127) / 128, 0, collect)) return 1;
    }
    return 0;
}

// every PML slab of one half step (planes [k0,k1)) in one launch (update_slabs_kernel); variant bit 26: one launch per slab
static bool slabs_merged(const b200fdtd_ctx* c) { return (c->variant & (1 << 26)) == 0; }
static int launch_slabs_merged(b200fdtd_ctx* c, int which, int k0, int k1, cudaStream_t stream, bool with_whole_rows = true, int sel = 0,
                               int k2 = 0, int k3 = 0 /* optional second plane range [k2,k3) */)
{
    SlabSet T; memset(&T, 0, sizeof(T));
    if (launch_volume_xslabs(c, which, k0, k1, stream, sel, &T)) return 1;
    if (with_whole_rows) if (launch_volume_fused(c, which, k0, k1, stream, sel, &T)) return 1;
    if (k3 > k2) {
        if (launch_volume_xslabs(c, which, k2, k3, stream, sel, &T)) return 1;
        if (with_whole_rows) if (launch_volume_fused(c, which, k2, k3, stream, sel, &T)) return 1;
    }
    if (T.n == 0) return 0;
    long long total = 0;
    for (int q = 0; q < T.n; ++q) { T.e[q].cta0 = (int)total; total += (long long)T.e[q].gx * T.e[q].gy * T.e[q].gz; }
    if (total > 0x7fffffffLL) return fail("merged slab launch too large");
    VolParams p; fill_vol_params(c, which, p);
    const bool cmp = c->cmp_meta[which] != nullptr && (c->variant & 4) == 0;
    const int ty = slab_ty(c);
    dim3 block(32, ty);
#define LAUNCH_SLABS(TYV) do { \
        if (which == 0) { if (cmp) update_slabs_kernel<0, TYV, true><<<(unsigned)total, block, 0, stream>>>(p, T); \
                          else update_slabs_kernel<0, TYV, false><<<(unsigned)total, block, 0, stream>>>(p, T); } \
        else { if (cmp) update_slabs_kernel<1, TYV, true><<<(unsigned)total, block, 0, stream>>>(p, T); \
               else update_slabs_kernel<1, TYV, false><<<(unsigned)total, block, 0, stream>>>(p, T); } } while (0)
    switch (ty) {
        case 1: LAUNCH_SLABS(1); break;
        case 2: LAUNCH_SLABS(2); break;
        case 4: LAUNCH_SLABS(4); break;
        case 8: LAUNCH_SLABS(8); break;
        case 16: LAUNCH_SLABS(16); break;
        default: return fail("unsupported slab ty=%d", ty);
    }
#undef LAUNCH_SLABS
    CKL();
    return 0;
}
// the slab launches of one half step on `stream` (merged) or on the side streams the caller has forked (one launch per slab)
static int launch_slabs(b200fdtd_ctx* c, int which, int k0, int k1, cudaStream_t merged_stream, cudaStream_t x_stream, cudaStream_t row_stream,
                        bool with_whole_rows = true, int sel = 0)
{
    if (slabs_merged(c)) return launch_slabs_merged(c, which, k0, k1, merged_stream, with_whole_rows, sel);
    if (launch_volume_xslabs(c, which, k0, k1, x_stream, sel)) return 1;
    if (with_whole_rows) return launch_volume_fused(c, which, k0, k1, row_stream, sel);
    return 0;
}

// the two boundary planes ka < kb of a z-slab rank in one plain launch (two z-chunks of one plane each) and one slab launch
static int launch_volume_two_planes(b200fdtd_ctx* c, int which, int ka, int kb);

// all volume launches of one half step restricted to planes [k0,k1), on the main stream
static int launch_volume(b200fdtd_ctx* c, int which, int k0, int k1)
{
    if (!c->volt || !c->vv) return fail("fields/coefficients not bound");
    if (!c->plan.valid) if (build_plan(c)) return 1;
    if (launch_volume_plain(c, which, k0, k1, c->stream)) return 1;
    return launch_slabs(c, which, k0, k1, c->stream, c->stream, c->stream);
}


// ---- the fused H->E launch ------------------------------------------------------------------------------------------
// update_he6_kernel needs the row-compressed operator (its x-vector slices live in shared memory); other operators take
// update_he_kernel, the plain fusion.  he6 also sweeps the whole-row PML slabs (variant bit 22 keeps them as separate launches).
static size_t he6_smem(int ty, int de, size_t xs_bytes)
{
    switch (ty) {
        case 3: return (de == 2 ? sizeof(He6Smem<3, 2>) : sizeof(He6Smem<3, 1>)) + xs_bytes;
        case 7: return (de == 2 ? sizeof(He6Smem<7, 2>) : sizeof(He6Smem<7, 1>)) + xs_bytes;
        case 15: return (de == 2 ? sizeof(He6Smem<15, 2>) : sizeof(He6Smem<15, 1>)) + xs_bytes;
    }
    return (size_t)1 << 30;
}
// 4-D tensor map of one field array [3][nz+2][ny][px] with a box of [3][rows][cols] (one plane): the whole tile of a CTA
// comes in with one cp.async.bulk.tensor instruction per field; elements outside the grid read as zero
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = [] {
        void* f = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); f = nullptr; }
        return (EncodeTiledFn)f;
    }();
    return fn;
}
static bool encode_field_map(CUtensorMap* m, const b200fdtd_ctx* c, const float* field, int box_cols, int box_rows)
{
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)c->px, (cuuint64_t)c->ny, (cuuint64_t)(c->nz + 2), 3};
    const cuuint64_t strides[3] = {(cuuint64_t)c->px * 4, (cuuint64_t)c->sz * 4, (cuuint64_t)c->cs * 4};
    const cuuint32_t box[4] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1, 3};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(field), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct HeChoice { bool he6; int de; bool pml; };
static HeChoice he_choice(const b200fdtd_ctx* c)
{
    HeChoice h{false, 1, false};
    const bool cmp = c->cmp_meta[0] != nullptr && c->cmp_meta[1] != nullptr && (c->variant & 4) == 0;
    const size_t xs_bytes = (size_t)(c->cmp_nvec[0] + c->cmp_nvec[1]) * 32 * sizeof(float4);
    if (!cmp || (c->variant & 512) != 0 || !(c->he_ty == 3 || c->he_ty == 7 || c->he_ty == 15)) return h;
    const size_t limit = 227 * 1024, per_sm = 228 * 1024;
    if (he6_smem(c->he_ty, 1, xs_bytes) > limit) return h;
    h.he6 = true;
    // two E planes landing if that does not cost a resident CTA (each CTA also reserves 1 KB)
    const int ctas1 = (int)(per_sm / (he6_smem(c->he_ty, 1, xs_bytes) + 1024));
    const int ctas2 = he6_smem(c->he_ty, 2, xs_bytes) <= limit ? (int)(per_sm / (he6_smem(c->he_ty, 2, xs_bytes) + 1024)) : 0;
    const int reg_ctas = 16 / (c->he_ty + 1) > 0 ? 16 / (c->he_ty + 1) : 1;       // __launch_bounds__ of the kernel
    const int r1 = ctas1 < reg_ctas ? ctas1 : reg_ctas, r2 = ctas2 < reg_ctas ? ctas2 : reg_ctas;
    h.de = c->he_de ? c->he_de : ((r2 >= r1 && r2 > 0) ? 2 : 1);
    if (h.de == 2 && ctas2 == 0) h.de = 1;
    // Whole-row PML slabs inside the fused launch (variant bit 23): one launch instead of 2 x (1 + slabs) per step and 96
    // instead of 120 B per PML cell, bit-exact like the separate launches.  Measured on B200 it loses on every grid size:
    // a slab row costs 3-4 x a plain row inside the fused launch (0.32 M cells: 58 vs 31 us per step; 106 M cells: 1.73 vs
    // 1.40 ms), and a CTA advances at the pace of its slowest row, so the default keeps the slabs in their own launches.
    h.pml = c->plan.nfused > 0 && (c->variant & (1 << 22)) == 0 && (c->variant & (1 << 23)) != 0;
    return h;
}

// second copies of the current flux of the whole-row slabs (the fused launch reads the old flux on tile halos)
static int ensure_flux_pp(b200fdtd_ctx* c)
{
    if (c->flux_pp || !he_choice(c).pml) return 0;
    VolumePlan& P = c->plan;
    for (int q = 0; q < P.nfused; ++q) {
        const size_t bytes = sizeof(float) * 3 * (size_t)P.fb[q].bz * P.fb[q].by * c->px;
        if (cudaMalloc((void**)&c->flux_alt[q], bytes) != cudaSuccess) {
            cudaGetLastError();
            for (int u = 0; u < q; ++u) { cudaFree(c->flux_alt[u]); c->flux_alt[u] = nullptr; }
            return 0;                                           // no room: the slabs keep their separate launches
        }
        CK(cudaMemsetAsync(c->flux_alt[q], 0, bytes, c->stream));
    }
    for (int q = 0; q < P.nfused; ++q) P.fb[q].flux_i_alt = c->flux_alt[q];
    c->flux_pp = true; c->fcur = 0;
    return 0;
}
// does the fused launch sweep the whole-row PML slabs of this run?
static bool he_pml(const b200fdtd_ctx* c) { return c->flux_pp && he_choice(c).pml; }

// the state of the current flux goes back to the caller's arrays (copy 0)
static int normalize_flux(b200fdtd_ctx* c)
{
    if (!c->flux_pp || c->fcur == 0) return 0;
    const VolumePlan& P = c->plan;
    for (int q = 0; q < P.nfused; ++q) {
        const size_t bytes = sizeof(float) * 3 * (size_t)P.fb[q].bz * P.fb[q].by * c->px;
        CK(cudaMemcpyAsync(P.fb[q].flux_i, P.fb[q].flux_i_alt, bytes, cudaMemcpyDeviceToDevice, c->stream));
    }
    c->fcur = 0;
    return 0;
}

static int launch_volume_two_planes(b200fdtd_ctx* c, int which, int ka, int kb)
{
    if (!c->volt || !c->vv) return fail("fields/coefficients not bound");
    if (!c->plan.valid) if (build_plan(c)) return 1;
    const VolumePlan& P = c->plan;
    const bool plain2 = P.nseg == 1 && ka >= P.seg0[0] && kb < P.seg1[0] && kb > ka && (c->variant & (1 << 28)) == 0;
    if (!plain2 || !slabs_merged(c)) {                       // a boundary plane inside a z-slab (end ranks), or the per-slab launches
        if (launch_volume(c, which, ka, ka + 1)) return 1;
        return launch_volume(c, which, kb, kb + 1);
    }
    RowParams r; memset(&r, 0, sizeof(r));
    r.j0 = 0; r.j1 = c->ny;
    if (P.nskip > 0) { r.sj0a = P.sj0[0]; r.sj1a = P.sj1[0]; }
    if (P.nskip > 1) { r.sj0b = P.sj0[1]; r.sj1b = P.sj1[1]; }
    if (P.has_lo) r.xw0 = P.xw0;
    if (P.has_hi) { r.xx1 = P.xx1; r.xw1 = P.xw1; }
    VolParams p; fill_vol_params(c, which, p);
    p.k0 = ka; p.k1 = kb + 1; p.kz = kb - ka; p.kspan = 1;
    const int ty = c->ty;
    dim3 block(32, ty), grid((c->px + 127) / 128, (c->ny + ty - 1) / ty, 2);
    const bool cmp = c->cmp_meta[which] != nullptr && (c->variant & 4) == 0;
#define LAUNCH2(TYV) do { \
        if (which == 0) { if (cmp) update_e_kernel<TYV, 0, true><<<grid, block, 0, c->stream>>>(p, r); \
                          else update_e_kernel<TYV, 0, false><<<grid, block, 0, c->stream>>>(p, r); } \
        else { if (cmp) update_h_kernel<TYV, 0, true><<<grid, block, 0, c->stream>>>(p, r); \
               else update_h_kernel<TYV, 0, false><<<grid, block, 0, c->stream>>>(p, r); } } while (0)
    switch (ty) {
        case 1: LAUNCH2(1); break;
        case 2: LAUNCH2(2); break;
        case 4: LAUNCH2(4); break;
        case 8: LAUNCH2(8); break;
        case 16: LAUNCH2(16); break;
        default: return fail("unsupported ty=%d", ty);
    }
#undef LAUNCH2
    CKL();
    return launch_slabs_merged(c, which, ka, ka + 1, c->stream, true, 0, kb, kb + 1);
}

// fused H->E launch: reads the current copies, writes the other copies (the caller flips)
static int launch_he(b200fdtd_ctx* c, cudaStream_t stream, int zc0 = -1, int zc1 = -1)
{
    const VolumePlan& P = c->plan;
    const bool clipped = zc0 >= 0;                           // z-slab ranks: interior planes [zc0, zc1) only
    const HeChoice hc = he_choice(c);
    const bool pml = hc.he6 && hc.pml && c->flux_pp;
    if (P.nseg != 1 && !pml) return clipped ? 0 : fail("fused H->E launch without a plain region");
    HeParams p;
    p.ein = cur_volt(c); p.hin = cur_curr(c); p.eout = oth_volt(c); p.hout = oth_curr(c);
    p.vv = c->vv; p.vi = c->vi; p.ii = c->ii; p.iv = c->iv;
    const bool cmp = c->cmp_meta[0] != nullptr && c->cmp_meta[1] != nullptr && (c->variant & 4) == 0;
    p.xv_e = c->cmp_xv[0]; p.meta_e = c->cmp_meta[0]; p.xv_h = c->cmp_xv[1]; p.meta_h = c->cmp_meta[1];
    p.ny = c->ny; p.px = c->px; p.sz = c->sz; p.cs = c->cs;
    p.X0 = P.has_lo ? P.xw0 : 0; p.X1 = P.has_hi ? P.xx1 : c->px; p.XT0 = P.has_hi ? P.xx1 + P.xw1 : c->px;
    p.X0s = p.X0;
    p.Y0 = P.ym0; p.Y1 = P.ym1; p.Z0 = P.nseg ? P.seg0[0] : 0; p.Z1 = P.nseg ? P.seg1[0] : 0;
    HePml Q; memset(&Q, 0, sizeof(Q));
    Q.b_zlo = Q.b_zhi = Q.b_ylo = Q.b_yhi = -1;
    if (pml) {
        // the launch covers every row and plane; rows of the whole-row slabs take the PML path
        Q.zlo = 0; Q.zhi = c->nz; Q.ym0 = 0; Q.ym1 = c->ny;
        for (int q = 0; q < P.nfused; ++q) {
            const FusedBox& B = P.fb[q];
            HePmlBox& D = Q.b[q];
            D.y0 = B.y0; D.by = B.by; D.z0 = B.z0; D.bz = B.bz;
            D.gin = c->fcur ? B.flux_i_alt : B.flux_i; D.gout = c->fcur ? B.flux_i : B.flux_i_alt; D.fv = B.flux_v;
            D.ah = B.ii; D.foh = B.iifo; D.fnh = B.iifn; D.ae = B.vv; D.foe = B.vvfo; D.fne = B.vvfn;
            const bool pc = (c->variant & 16) == 0;
            D.xvh = pc ? B.xv_i : nullptr; D.mh = pc && B.xv_i ? B.meta_i : nullptr;
            D.xve = pc ? B.xv_v : nullptr; D.me = pc && B.xv_v ? B.meta_v : nullptr;
            if (B.y0 == 0 && B.by == c->ny) { if (B.z0 == 0) { Q.b_zlo = q; Q.zlo = B.bz; } else { Q.b_zhi = q; Q.zhi = B.z0; } }
            else if (B.y0 == 0) { Q.b_ylo = q; Q.ym0 = B.by; }
            else { Q.b_yhi = q; Q.ym1 = B.y0; }
        }
        p.X0s = 0; p.Y0 = 0; p.Y1 = c->ny; p.Z0 = 0; p.Z1 = c->nz;
    }
    if (clipped) { if (p.Z0 < zc0) p.Z0 = zc0; if (p.Z1 > zc1) p.Z1 = zc1; if (p.Z1 <= p.Z0) return 0; }
    if (p.Y1 <= p.Y0 || p.Z1 <= p.Z0 || (!pml && p.X1 <= p.X0)) return fail("fused H->E launch over an empty region");
    const int ty = c->he_ty;
    int kz = c->he_kz; if (kz > p.Z1 - p.Z0) kz = p.Z1 - p.Z0;
    {   // small grids: the z-march is a serial chain of planes per CTA, so the launch gets its parallelism from more, shorter
        // chunks (each pays one extra plane of halo recompute) until the machine is filled a few times over
        const long long per_chunk = (long long)((c->px - p.X0s + HE_SEG - 1) / HE_SEG) * ((p.Y1 - p.Y0 + ty - 1) / ty);
        while (kz > 4 && per_chunk * ((p.Z1 - p.Z0 + kz - 1) / kz) < 148LL * 4) kz = (kz + 1) / 2;
    }
    { const int n = (p.Z1 - p.Z0 + kz - 1) / kz; kz = (p.Z1 - p.Z0 + n - 1) / n; }      // chunks of equal length
    p.kz = kz;
    p.pf = (c->variant >> 16) & 3;                   // L2 prefetch distance in planes (plain fusion only): 0 = default (1), 3 = off
    p.pf = p.pf == 0 ? 1 : (p.pf == 3 ? 0 : p.pf);
    dim3 block(32, ty + 1);
    dim3 grid((c->px - p.X0s + HE_SEG - 1) / HE_SEG, (p.Y1 - p.Y0 + ty - 1) / ty, (p.Z1 - p.Z0 + kz - 1) / kz);
    if (grid.y > 65535 || grid.z > 65535) return fail("grid too large for the fused launch");
    p.b_sz = 4 * p.sz; p.b_cs = 4 * p.cs; p.b_2cs = 8 * p.cs;
    p.b_row = 4LL * p.px; p.b_row_2cs = 4 * (p.px + 2 * p.cs);
    p.meta_step = 32 * p.ny;
    p.nv_e = c->cmp_nvec[0]; p.nv_h = c->cmp_nvec[1];
    const size_t xs_bytes = (size_t)(p.nv_e + p.nv_h) * 32 * sizeof(float4);
    // whole-tile staging by tiled TMA copies (one instruction per field and plane instead of eight row copies per warp);
    // variant bit 25 keeps the per-row 1-D bulk copies
    CUtensorMap tmE, tmH; memset(&tmE, 0, sizeof(tmE)); memset(&tmH, 0, sizeof(tmH));
    const bool tm = hc.he6 && (c->variant & (1 << 25)) == 0 && p.px * 4LL < (1LL << 32) &&
                    encode_field_map(&tmE, c, p.ein, 132, ty + 2) && encode_field_map(&tmH, c, p.hin, 128, ty + 1);
#define LAUNCH_HE6(TYV, DEV, PMLV) do { const size_t sm6 = sizeof(He6Smem<TYV, DEV>) + xs_bytes; \
        if (tm) { CK(cudaFuncSetAttribute(update_he6_kernel<TYV, DEV, PMLV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm6)); \
                  update_he6_kernel<TYV, DEV, PMLV, true><<<grid, block, sm6, stream>>>(p, Q, tmE, tmH); } \
        else { CK(cudaFuncSetAttribute(update_he6_kernel<TYV, DEV, PMLV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm6)); \
               update_he6_kernel<TYV, DEV, PMLV, false><<<grid, block, sm6, stream>>>(p, Q, tmE, tmH); } } while (0)
#define LAUNCH_HE(TYV) do { \
        if (hc.he6) { if (hc.de == 2) { if (pml) LAUNCH_HE6(TYV, 2, true); else LAUNCH_HE6(TYV, 2, false); } \
                      else { if (pml) LAUNCH_HE6(TYV, 1, true); else LAUNCH_HE6(TYV, 1, false); } } \
        else if (cmp) update_he_kernel<TYV, true><<<grid, block, 0, stream>>>(p); \
        else update_he_kernel<TYV, false><<<grid, block, 0, stream>>>(p); } while (0)
    switch (ty) {
        case 3: LAUNCH_HE(3); break;
        case 7: LAUNCH_HE(7); break;
        case 15: LAUNCH_HE(15); break;
        default: return fail("unsupported fused-launch tile height %d", ty);
    }
#undef LAUNCH_HE
#undef LAUNCH_HE6
    CKL();
    return 0;
}

// can one graph chunk of `steps` steps use the fused H->E launches?  (single slab, every PML box fused into the volume
// launches, a plain region, memory for the second copy of the fields)
static bool he_ready(b200fdtd_ctx* c, int steps)
{
    if ((c->variant & 128) || steps < 2 || !c->plan.valid || c->pml.n > 0) return false;
    const VolumePlan& P = c->plan;
    const bool pml_ok = he_choice(c).pml;
    if (P.nseg != 1 && !pml_ok) return false;
    if (P.nseg == 1 && !pml_ok && (P.ym1 <= P.ym0 || (P.has_hi ? P.xx1 : c->px) <= (P.has_lo ? P.xw0 : 0))) return false;
    if (!c->alt_volt) {
        const size_t bytes = sizeof(float) * 3 * (size_t)c->cs;
        if (cudaMalloc((void**)&c->alt_volt, bytes) != cudaSuccess) { cudaGetLastError(); c->alt_volt = nullptr; return false; }
        if (cudaMalloc((void**)&c->alt_curr, bytes) != cudaSuccess) { cudaGetLastError(); cudaFree(c->alt_volt); c->alt_volt = c->alt_curr = nullptr; return false; }
        cudaMemsetAsync(c->alt_volt, 0, bytes, c->stream);
        cudaMemsetAsync(c->alt_curr, 0, bytes, c->stream);
        c->alt_owned = true;
    }
    if (ensure_flux_pp(c)) return false;
    if (P.nseg != 1 && !he_pml(c)) return false;
    return true;
}

// ghost planes of the second copy follow the bound arrays (the caller may rewrite its ghost planes between runs)
static int sync_alt_ghosts(b200fdtd_ctx* c)
{
    if (!c->alt_volt) return 0;
    const size_t pitch = sizeof(float) * (size_t)c->cs, w = sizeof(float) * (size_t)c->sz;
    const long long top = (long long)(c->nz + 1) * c->sz;
    CK(cudaMemcpy2DAsync(c->alt_volt, pitch, c->volt, pitch, w, 3, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpy2DAsync(c->alt_volt + top, pitch, c->volt + top, pitch, w, 3, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpy2DAsync(c->alt_curr, pitch, c->curr, pitch, w, 3, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpy2DAsync(c->alt_curr + top, pitch, c->curr + top, pitch, w, 3, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

// fork/join of the side stream that runs the fused PML slab launches next to the plain launch
static int fork_side(b200fdtd_ctx* c)
{
    CK(cudaEventRecord(c->ev_fork, c->stream));
    CK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    CK(cudaStreamWaitEvent(c->side2, c->ev_fork, 0));
    for (int q = 0; q < 3; ++q) CK(cudaStreamWaitEvent(c->slab_s[q], c->ev_fork, 0));
    return 0;
}
static int join_side(b200fdtd_ctx* c)
{
    CK(cudaEventRecord(c->ev_join, c->side));
    CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    CK(cudaEventRecord(c->ev_join2, c->side2));
    CK(cudaStreamWaitEvent(c->stream, c->ev_join2, 0));
    for (int q = 0; q < 3; ++q) { CK(cudaEventRecord(c->slab_ev[q], c->slab_s[q])); CK(cudaStreamWaitEvent(c->stream, c->slab_ev[q], 0)); }
    return 0;
}
// the stream of the whole-row slab launches: their own side stream, or the x-slab one (variant bit 64)
static cudaStream_t slab_stream(b200fdtd_ctx* c) { return (c->variant & 64) ? c->side : c->side2; }

#include "kernels_narrow.cuh"      // K3..K11 narrow-band kernels

// ------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------
template <typename T>
static int upload(T** dst, const T* src, int64_t n, cudaStream_t s)
{
    if (*dst) { cudaFree(*dst); *dst = nullptr; }
    if (n <= 0) return 0;
    CK(cudaMalloc((void**)dst, sizeof(T) * n));
    CK(cudaMemcpyAsync(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}
static void drop_graph(b200fdtd_ctx* c)
{
    if (c->graph) { cudaGraphExecDestroy(c->graph); c->graph = nullptr; c->graph_steps = 0; }
    for (int q = 0; q < 4; ++q) if (c->pgraph[q]) { cudaGraphExecDestroy(c->pgraph[q]); c->pgraph[q] = nullptr; }
    c->pgraph_steps = 0;
}


// ------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------
// experiment switches for a whole test run: B200FDTD_VARIANT_OR is OR-ed into every variant word
static int env_variant_or() { static const int v = [] { const char* e = getenv("B200FDTD_VARIANT_OR"); return e ? atoi(e) : 0; }(); return v; }

extern "C" int b200fdtd_create(b200fdtd_ctx** out, int device, int nx, int ny, int nz, int px, void* stream)
{
    if (!out) return fail("out is NULL");
    if (nx < 2 || ny < 2 || nz < 1) return fail("grid too small: %d x %d x %d", nx, ny, nz);
    if (px < nx || (px % 4) != 0) return fail("px=%d must be >= nx=%d and a multiple of 4", px, nx);
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail("device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    b200fdtd_ctx* c = new b200fdtd_ctx();
    c->device = device; c->nx = nx; c->ny = ny; c->nz = nz; c->px = px;
    c->sz = (long long)ny * px; c->cs = (long long)(nz + 2) * c->sz;
    c->stream = (cudaStream_t)stream;                       // NULL = the default stream (torch's default stream)
    c->own_stream = false;
    CK(cudaStreamCreateWithFlags(&c->side_lo, cudaStreamNonBlocking));
    { int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&c->side_hi, cudaStreamNonBlocking, hi)); }
    c->side = c->side_lo;
    CK(cudaStreamCreateWithFlags(&c->side2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming));
    for (int q = 0; q < 3; ++q) { CK(cudaStreamCreateWithFlags(&c->slab_s[q], cudaStreamNonBlocking));
                                  CK(cudaEventCreateWithFlags(&c->slab_ev[q], cudaEventDisableTiming)); }
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CK(cudaMalloc((void**)&c->d_ts, sizeof(int)));
    CK(cudaMemsetAsync(c->d_ts, 0, sizeof(int), c->stream));
    c->variant = env_variant_or();
    if (const char* e = getenv("B200FDTD_HE_TY")) { const int t = atoi(e); if (t == 3 || t == 7 || t == 15) c->he_ty = t; }
    if (const char* e = getenv("B200FDTD_HE_KZ")) { const int t = atoi(e); if (t >= 1) c->he_kz = t; }
    if (const char* e = getenv("B200FDTD_HE_DE")) { const int t = atoi(e); if (t == 1 || t == 2) c->he_de = t; }
    c->n_partials = 148 * 8;
    CK(cudaMalloc((void**)&c->d_partials, sizeof(double) * 2 * c->n_partials));
    CK(cudaMalloc((void**)&c->d_energy, sizeof(double) * 2));
    *out = c;
    return 0;
}

extern "C" int b200fdtd_destroy(b200fdtd_ctx* c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    drop_graph(c);
    cudaFree(c->d_ts); cudaFree(c->d_partials); cudaFree(c->d_energy);
    if (c->alt_owned) { cudaFree(c->alt_volt); cudaFree(c->alt_curr); }
    for (int q = 0; q < 4; ++q) cudaFree(c->flux_alt[q]);
    cudaFree(c->exc_idx); cudaFree(c->exc_amp); cudaFree(c->exc_delay); cudaFree(c->exc_sig);
    cudaFree(c->mur_dst); cudaFree(c->mur_src); cudaFree(c->mur_coeff); cudaFree(c->mur_tmp); cudaFree(c->mur_segs);
    cudaFree(c->pr_kind); cudaFree(c->pr_off); cudaFree(c->pr_idx); cudaFree(c->pr_w); cudaFree(c->pr_freqs);
    cudaFree(c->nf_freqs);
    for (int a = 0; a < 3; ++a) { cudaFree(c->inv_len[a]); cudaFree(c->inv_dual[a]); }
    cudaFree(c->d_pml); cudaFree(c->d_faces);
    if (c->side_lo) { cudaStreamSynchronize(c->side_lo); cudaStreamDestroy(c->side_lo); }
    if (c->side_hi) { cudaStreamSynchronize(c->side_hi); cudaStreamDestroy(c->side_hi); }
    if (c->side2) { cudaStreamSynchronize(c->side2); cudaStreamDestroy(c->side2); }
    if (c->ev_join2) cudaEventDestroy(c->ev_join2);
    for (int q = 0; q < 3; ++q) { if (c->slab_s[q]) { cudaStreamSynchronize(c->slab_s[q]); cudaStreamDestroy(c->slab_s[q]); }
                                  if (c->slab_ev[q]) cudaEventDestroy(c->slab_ev[q]); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

extern "C" int b200fdtd_bind_fields(b200fdtd_ctx* c, float* volt, float* curr)
{
    if (!c || !volt || !curr) return fail("NULL argument");
    if (((uintptr_t)volt | (uintptr_t)curr) & 15) return fail("field pointers must be 16-byte aligned");
    c->volt = volt; c->curr = curr; drop_graph(c);
    return 0;
}

extern "C" int b200fdtd_bind_coeffs(b200fdtd_ctx* c, const float* vv, const float* vi, const float* ii, const float* iv)
{
    if (!c || !vv || !vi || !ii || !iv) return fail("NULL argument");
    if (((uintptr_t)vv | (uintptr_t)vi | (uintptr_t)ii | (uintptr_t)iv) & 15) return fail("coefficient pointers must be 16-byte aligned");
    c->vv = vv; c->vi = vi; c->ii = ii; c->iv = iv; drop_graph(c);
    c->cmp_xv[0] = c->cmp_xv[1] = nullptr; c->cmp_meta[0] = c->cmp_meta[1] = nullptr;   // new arrays: compression must be set again
    return 0;
}

extern "C" int b200fdtd_set_tuning(b200fdtd_ctx* c, int kz, int ty, int variant)
{
    if (!c) return fail("NULL ctx");
    if (kz < 1) return fail("kz must be >= 1");
    if (!(ty == 1 || ty == 2 || ty == 4 || ty == 8 || ty == 16)) return fail("ty must be 1,2,4,8 or 16");
    { const int t = (variant >> 8) & 31; if (!(t == 0 || t == 1 || t == 2 || t == 4 || t == 8 || t == 16)) return fail("slab ty (variant bits 8-12) must be 0,1,2,4,8 or 16"); }
    c->kz = kz; c->ty = ty; c->variant = variant | env_variant_or(); drop_graph(c);
    c->side = (variant & 32) ? c->side_hi : c->side_lo;
    c->plan.valid = false;
    return 0;
}

// one warp per (row, slot): a compressed row must reproduce the full array bit for bit, else it is demoted to ROW_FULL
__global__ void __launch_bounds__(256) verify_rows_kernel(unsigned char* __restrict__ meta, const float* __restrict__ xv, int nvec,
        const float* __restrict__ ca, const float* __restrict__ cb, int ny, int nz, int px, long long sz, long long cs,
        unsigned long long* __restrict__ counts /* [0] demoted, [1] compressed row-slots */)
{
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long nrows = (long long)nz * ny;
    if (w >= nrows * 6) return;
    const int slot = (int)(w % 6);
    const long long row = w / 6;
    const int k = (int)(row / ny), j = (int)(row % ny);
    RowMeta* M = reinterpret_cast<RowMeta*>(meta + ((long long)(k + 1) * ny + j) * 32);
    const unsigned id = M->id[slot];
    if (id == ROW_FULL) return;
    bool ok = id < (unsigned)nvec;
    if (ok) {
        const float sc = M->sc[slot];
        const float* full = (slot < 3 ? ca : cb) + (long long)(slot % 3) * cs + (long long)(k + 1) * sz + (long long)j * px;
        const float* v = xv + (size_t)id * px;
        for (int i = lane; i < px; i += 32)
            if (__float_as_uint(__fmul_rn(sc, v[i])) != __float_as_uint(full[i])) ok = false;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        if (!ok) { M->id[slot] = (unsigned char)ROW_FULL; atomicAdd(&counts[0], 1ULL); }
        else atomicAdd(&counts[1], 1ULL);
    }
}

// expansion of a compressed operator into the bound full arrays: one warp per (row, slot), full[i] = fl32(scale*xvec[i])
// for a compressed slot, 0 for a slot that is streamed in full (the caller patches those rows afterwards)
__global__ void __launch_bounds__(256) expand_rows_kernel(const unsigned char* __restrict__ meta, const float* __restrict__ xv, int nvec,
        float* __restrict__ ca, float* __restrict__ cb, long long nrows, int px, long long cs)
{
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nrows * 6) return;
    const int slot = (int)(w % 6);
    const long long row = w / 6;
    const RowMeta* M = reinterpret_cast<const RowMeta*>(meta + row * 32);
    const unsigned id = M->id[slot];
    const float sc = M->sc[slot];
    float4* dst = reinterpret_cast<float4*>((slot < 3 ? ca : cb) + (long long)(slot % 3) * cs + row * px);
    const float4* v = reinterpret_cast<const float4*>(xv + (size_t)(id < (unsigned)nvec ? id : 0) * px);
    const bool cmp = id < (unsigned)nvec;
    for (int i = lane; i < px / 4; i += 32) {
        float4 o = zero4();
        if (cmp) { const float4 x = __ldg(v + i); o = make_float4(__fmul_rn(sc, x.x), __fmul_rn(sc, x.y), __fmul_rn(sc, x.z), __fmul_rn(sc, x.w)); }
        __stcs(dst + i, o);
    }
}

extern "C" int b200fdtd_expand_rows(b200fdtd_ctx* c, int which, int nvec, const float* xvecs, const void* meta)
{
    if (!c) return fail("NULL ctx");
    if (which != 0 && which != 1) return fail("which must be 0 (E pass) or 1 (H pass)");
    if (!c->vv) return fail("bind the coefficient arrays before expanding into them");
    if (nvec < 1 || nvec > 255 || !xvecs || !meta) return fail("bad compression tables");
    if (((uintptr_t)xvecs | (uintptr_t)meta) & 15) return fail("compression tables must be 16-byte aligned");
    CK(cudaSetDevice(c->device));
    drop_graph(c);
    const long long nrows = (long long)(c->nz + 2) * c->ny;
    const long long blocks = (nrows * 6 * 32 + 255) / 256;
    expand_rows_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>((const unsigned char*)meta, xvecs, nvec,
        const_cast<float*>(which == 0 ? c->vv : c->ii), const_cast<float*>(which == 0 ? c->vi : c->iv), nrows, c->px, c->cs);
    CKL();
    return 0;
}

// pad[0] of a row record = 1 if any of its six slots is streamed in full (the volume kernels branch on it once per row)
__global__ void __launch_bounds__(256) flag_rows_kernel(unsigned char* __restrict__ meta, long long nrows)
{
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= nrows) return;
    RowMeta* M = reinterpret_cast<RowMeta*>(meta + q * 32);
    bool any = false;
    for (int sl = 0; sl < 6; ++sl) any |= M->id[sl] == ROW_FULL;
    M->pad[0] = any ? 1 : 0;
}

extern "C" int b200fdtd_set_row_compression(b200fdtd_ctx* c, int which, int nvec, const float* xvecs, void* meta,
                                             int64_t* n_compressed, int64_t* n_demoted)
{
    if (!c) return fail("NULL ctx");
    if (which != 0 && which != 1) return fail("which must be 0 (E pass) or 1 (H pass)");
    CK(cudaSetDevice(c->device));
    drop_graph(c);
    if (nvec == 0 || !xvecs || !meta) { c->cmp_xv[which] = nullptr; c->cmp_meta[which] = nullptr; return 0; }
    if (!c->vv) return fail("bind the coefficients before setting their row compression");
    if (nvec < 0 || nvec > 255) return fail("nvec=%d out of range (1..255)", nvec);
    if (((uintptr_t)xvecs | (uintptr_t)meta) & 15) return fail("compression tables must be 16-byte aligned");
    unsigned long long* d_counts = nullptr;
    CK(cudaMalloc((void**)&d_counts, 2 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(d_counts, 0, 2 * sizeof(unsigned long long), c->stream));
    const long long warps = (long long)c->nz * c->ny * 6;
    const long long blocks = (warps * 32 + 255) / 256;
    verify_rows_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>((unsigned char*)meta, xvecs, nvec,
        which == 0 ? c->vv : c->ii, which == 0 ? c->vi : c->iv, c->ny, c->nz, c->px, c->sz, c->cs, d_counts);
    g_launches.fetch_add(1);
    { const long long nrows = (long long)(c->nz + 2) * c->ny;
      flag_rows_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, c->stream>>>((unsigned char*)meta, nrows);
      g_launches.fetch_add(1); }
    cudaError_t e = cudaGetLastError();
    unsigned long long h[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_counts, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_counts);
    if (e != cudaSuccess) return fail("row compression verification failed: %s", cudaGetErrorString(e));
    if (n_demoted) *n_demoted = (int64_t)h[0];
    if (n_compressed) *n_compressed = (int64_t)h[1];
    c->cmp_xv[which] = xvecs; c->cmp_meta[which] = (const unsigned char*)meta; c->cmp_nvec[which] = nvec;
    return 0;
}

extern "C" int b200fdtd_set_excitation(b200fdtd_ctx* c, int64_t n, const int64_t* idx, const float* amp,
                                        const int32_t* delay, const float* signal, int32_t siglen)
{
    if (!c) return fail("NULL ctx");
    if (n < 0 || (n > 0 && (!idx || !amp || !delay || !signal || siglen <= 0))) return fail("bad excitation arguments");
    CK(cudaSetDevice(c->device));
    const long long total = 3 * c->cs;
    for (int64_t e = 0; e < n; ++e) if (idx[e] < 0 || idx[e] >= total) return fail("excitation index %lld out of range", (long long)idx[e]);
    drop_graph(c);
    c->n_exc = n; c->exc_siglen = siglen;
    if (upload(&c->exc_idx, idx, n, c->stream)) return 1;
    if (upload(&c->exc_amp, amp, n, c->stream)) return 1;
    if (upload(&c->exc_delay, (const int*)delay, n, c->stream)) return 1;
    if (upload(&c->exc_sig, signal, n > 0 ? siglen : 0, c->stream)) return 1;
    return 0;
}

extern "C" int b200fdtd_set_mur(b200fdtd_ctx* c, int64_t n, const int64_t* dst, const int64_t* src, const float* coeff)
{
    if (!c) return fail("NULL ctx");
    if (n < 0 || (n > 0 && (!dst || !src || !coeff))) return fail("bad Mur arguments");
    CK(cudaSetDevice(c->device));
    const long long total = 3 * c->cs;
    for (int64_t e = 0; e < n; ++e)
        if (dst[e] < 0 || dst[e] >= total || src[e] < 0 || src[e] >= total) return fail("Mur index out of range at entry %lld", (long long)e);
    drop_graph(c);
    c->n_mur = n;
    if (upload(&c->mur_dst, dst, n, c->stream)) return 1;
    if (upload(&c->mur_src, src, n, c->stream)) return 1;
    if (upload(&c->mur_coeff, coeff, n, c->stream)) return 1;
    if (c->mur_tmp) { cudaFree(c->mur_tmp); c->mur_tmp = nullptr; }
    if (n > 0) { CK(cudaMalloc((void**)&c->mur_tmp, sizeof(float) * n)); CK(cudaMemsetAsync(c->mur_tmp, 0, sizeof(float) * n, c->stream)); }
    // arithmetic runs of (dst, src): rows / columns of the boundary faces (at most MUR_RUN_MAX edges each: one warp sweeps a
    // run, so short runs keep the sweep a few loads deep and the faces spread over the whole machine)
    std::vector<MurSeg> segs;
    for (int64_t e = 0; e < n;) {
        MurSeg S; S.dst0 = dst[e]; S.src0 = src[e]; S.e0 = e; S.pad = 0; S.sd = 0; S.ss = 0; S.count = 1;
        if (e + 1 < n) {
            const long long sd = dst[e + 1] - dst[e], ss = src[e + 1] - src[e];
            if (sd > -(1LL << 30) && sd < (1LL << 30) && ss > -(1LL << 30) && ss < (1LL << 30)) {
                S.sd = (int)sd; S.ss = (int)ss;
                while (e + S.count < n && S.count < MUR_RUN_MAX && dst[e + S.count] - dst[e + S.count - 1] == sd &&
                       src[e + S.count] - src[e + S.count - 1] == ss) S.count++;
            }
        }
        segs.push_back(S);
        e += S.count;
    }
    if (c->mur_segs) { cudaFree(c->mur_segs); c->mur_segs = nullptr; }
    c->n_mur_segs = 0;
    if (n > 0 && (int64_t)segs.size() * 16 <= n && segs.size() < (1u << 30) && (c->variant & (1 << 24)) == 0) {   // long runs: drop the index lists
        if (upload(&c->mur_segs, segs.data(), (int64_t)segs.size(), c->stream)) return 1;
        c->n_mur_segs = (int)segs.size();
    }
    return 0;
}

// one warp per (slab row, slot): compressed slab rows must reproduce the full coefficient arrays bit for bit
__global__ void __launch_bounds__(256) verify_pml_rows_kernel(unsigned char* __restrict__ meta, const float* __restrict__ xv, int nvec,
        const float* __restrict__ a, const float* __restrict__ fo, const float* __restrict__ fn, int bx, int by, int bz,
        unsigned long long* __restrict__ counts)
{
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long nrows = (long long)bz * by;
    if (w >= nrows * 9) return;
    const int slot = (int)(w % 9);
    const long long row = w / 9;
    PmlRowMeta* M = reinterpret_cast<PmlRowMeta*>(meta + row * 48);
    const unsigned id = M->id[slot];
    if (id == ROW_FULL) return;
    bool ok = id < (unsigned)nvec;
    if (ok) {
        const float sc = M->sc[slot];
        const float* full = (slot < 3 ? a : (slot < 6 ? fo : fn)) + ((long long)(slot % 3) * nrows + row) * bx;
        const float* v = xv + (size_t)id * bx;
        for (int i = lane; i < bx; i += 32)
            if (__float_as_uint(__fmul_rn(sc, v[i])) != __float_as_uint(full[i])) ok = false;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        if (!ok) { M->id[slot] = (unsigned char)ROW_FULL; atomicAdd(&counts[0], 1ULL); }
        else atomicAdd(&counts[1], 1ULL);
    }
}

// pad[0] of a slab row record = 1 if any of its nine slots is streamed in full
__global__ void __launch_bounds__(256) flag_pml_rows_kernel(unsigned char* __restrict__ meta, long long nrows)
{
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= nrows) return;
    PmlRowMeta* M = reinterpret_cast<PmlRowMeta*>(meta + q * 48);
    bool any = false;
    for (int sl = 0; sl < 9; ++sl) any |= M->id[sl] == ROW_FULL;
    M->pad[0] = any ? 1 : 0;
}

static int verify_pml_rows(b200fdtd_ctx* c, const b200fdtd_pml_box& B, int which)
{
    unsigned long long* d_counts = nullptr;
    CK(cudaMalloc((void**)&d_counts, 2 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(d_counts, 0, 2 * sizeof(unsigned long long), c->stream));
    const long long warps = (long long)B.bz * B.by * 9;
    const long long blocks = (warps * 32 + 255) / 256;
    verify_pml_rows_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>((unsigned char*)(which == 0 ? B.meta_v : B.meta_i),
        which == 0 ? B.xvecs_v : B.xvecs_i, which == 0 ? B.nvec_v : B.nvec_i,
        which == 0 ? B.vv : B.ii, which == 0 ? B.vvfo : B.iifo, which == 0 ? B.vvfn : B.iifn, B.bx, B.by, B.bz, d_counts);
    g_launches.fetch_add(1);
    { const long long nrows = (long long)B.bz * B.by;
      flag_pml_rows_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, c->stream>>>((unsigned char*)(which == 0 ? B.meta_v : B.meta_i), nrows);
      g_launches.fetch_add(1); }
    cudaError_t e = cudaGetLastError();
    unsigned long long h[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_counts, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_counts);
    if (e != cudaSuccess) return fail("PML row compression verification failed: %s", cudaGetErrorString(e));
    c->pml_rows_compressed += (int64_t)h[1]; c->pml_rows_demoted += (int64_t)h[0];
    return 0;
}

extern "C" int b200fdtd_set_pml(b200fdtd_ctx* c, int nboxes, const b200fdtd_pml_box* boxes)
{
    if (!c) return fail("NULL ctx");
    if (nboxes < 0 || nboxes > MAX_PML_BOXES) return fail("nboxes=%d out of range (max %d)", nboxes, MAX_PML_BOXES);
    CK(cudaSetDevice(c->device));
    drop_graph(c);
    PmlTable t; memset(&t, 0, sizeof(t));
    c->pml_rows_compressed = c->pml_rows_demoted = 0;
    t.n = nboxes; long long start = 0;
    for (int b = 0; b < nboxes; ++b) {
        const b200fdtd_pml_box& B = boxes[b];
        if (B.bx <= 0 || B.by <= 0 || B.bz <= 0 || B.x0 < 0 || B.y0 < 0 || B.z0 < 0 ||
            B.x0 + B.bx > c->px || B.y0 + B.by > c->ny || B.z0 + B.bz > c->nz)
            return fail("PML box %d outside the grid", b);
        if (!B.flux_v || !B.flux_i || !B.vv || !B.vvfo || !B.vvfn || !B.ii || !B.iifo || !B.iifn) return fail("PML box %d has NULL arrays", b);
        for (int q = 0; q < b; ++q) {
            const b200fdtd_pml_box& O = boxes[q];
            const bool apart = B.x0 >= O.x0 + O.bx || O.x0 >= B.x0 + B.bx || B.y0 >= O.y0 + O.by || O.y0 >= B.y0 + B.by ||
                               B.z0 >= O.z0 + O.bz || O.z0 >= B.z0 + B.bz;
            if (!apart) return fail("PML boxes %d and %d overlap (every cell belongs to at most one box)", q, b);
        }
        PmlBoxDev& D = t.b[b];
        D.x0 = B.x0; D.y0 = B.y0; D.z0 = B.z0; D.bx = B.bx; D.by = B.by; D.bz = B.bz; D.start = start;
        D.flux_v = B.flux_v; D.flux_i = B.flux_i; D.vv = B.vv; D.vvfo = B.vvfo; D.vvfn = B.vvfn; D.ii = B.ii; D.iifo = B.iifo; D.iifn = B.iifn;
        D.xv_v = D.xv_i = nullptr; D.meta_v = D.meta_i = nullptr;
        if (B.nvec_v > 0 && B.xvecs_v && B.meta_v && B.bx % 4 == 0 && !(((uintptr_t)B.xvecs_v | (uintptr_t)B.meta_v) & 15)) {
            if (verify_pml_rows(c, B, 0)) return 1;
            D.xv_v = B.xvecs_v; D.meta_v = (const unsigned char*)B.meta_v;
        }
        if (B.nvec_i > 0 && B.xvecs_i && B.meta_i && B.bx % 4 == 0 && !(((uintptr_t)B.xvecs_i | (uintptr_t)B.meta_i) & 15)) {
            if (verify_pml_rows(c, B, 1)) return 1;
            D.xv_i = B.xvecs_i; D.meta_i = (const unsigned char*)B.meta_i;
        }
        start += 3LL * B.bx * B.by * B.bz;
    }
    t.total = start;
    c->pml_all = t;
    c->plan.valid = false;
    return build_plan(c);
}

extern "C" int b200fdtd_set_probes(b200fdtd_ctx* c, int nprobes, const int32_t* kind, const int64_t* offset,
                                    const int64_t* idx, const float* weight, int interval, int max_samples,
                                    float* series, int nfreq, const double* freqs, float* dft, double dt)
{
    if (!c) return fail("NULL ctx");
    if (nprobes < 0) return fail("nprobes < 0");
    CK(cudaSetDevice(c->device));
    drop_graph(c);
    c->n_probes = nprobes;
    if (nprobes == 0) return 0;
    if (!kind || !offset || !idx || !weight || !series) return fail("NULL probe argument");
    if (interval < 1 || max_samples < 1) return fail("interval and max_samples must be >= 1");
    if (nfreq < 0 || (nfreq > 0 && (!freqs || !dft))) return fail("bad probe DFT arguments");
    if (c->faces.n > 0 && c->nf_interval != interval) return fail("probe interval %d differs from NF2FF interval %d", interval, c->nf_interval);
    const long long total = 3 * c->cs;
    const int64_t ne = offset[nprobes];
    for (int64_t e = 0; e < ne; ++e) if (idx[e] < 0 || idx[e] >= total) return fail("probe index out of range at entry %lld", (long long)e);
    c->interval = interval; c->max_samples = max_samples; c->pr_series = series; c->pr_nfreq = nfreq; c->pr_dft = dft; c->dt = dt;
    if (upload(&c->pr_kind, (const int*)kind, nprobes, c->stream)) return 1;
    if (upload(&c->pr_off, offset, nprobes + 1, c->stream)) return 1;
    if (upload(&c->pr_idx, idx, ne, c->stream)) return 1;
    if (upload(&c->pr_w, weight, ne, c->stream)) return 1;
    if (upload(&c->pr_freqs, freqs, nfreq, c->stream)) return 1;
    return 0;
}

extern "C" int b200fdtd_set_nf2ff(b200fdtd_ctx* c, int nfaces, const b200fdtd_nf2ff_face* faces, int nfreq,
                                   const double* freqs, int interval, double dt,
                                   const float* ilx, const float* ily, const float* ilz,
                                   const float* idx_, const float* idy, const float* idz)
{
    if (!c) return fail("NULL ctx");
    if (nfaces < 0 || nfaces > MAX_FACES) return fail("nfaces=%d out of range (max %d)", nfaces, MAX_FACES);
    CK(cudaSetDevice(c->device));
    drop_graph(c);
    c->faces.n = nfaces;
    if (nfaces == 0) return 0;
    if (nfreq < 1 || !freqs || interval < 1) return fail("bad NF2FF frequency/interval arguments");
    if (c->n_probes > 0 && c->interval != interval) return fail("NF2FF interval %d differs from probe interval %d", interval, c->interval);
    if (!ilx || !ily || !ilz || !idx_ || !idy || !idz) return fail("NULL mesh spacing array");
    const int dims[3] = {c->nx, c->ny, c->nz};
    int max_nodes = 0;
    for (int q = 0; q < nfaces; ++q) {
        const b200fdtd_nf2ff_face& F = faces[q];
        if (F.normal < 0 || F.normal > 2 || !F.acc) return fail("bad NF2FF face %d", q);
        const int a = (F.normal + 1) % 3, b = (F.normal + 2) % 3;
        // z indices may touch plane 0 of the slab (needs the ghost plane below); x/y must be interior
        const int lo_n = F.normal == 2 ? 0 : 1, lo_a = a == 2 ? 0 : 1, lo_b = b == 2 ? 0 : 1;
        if (F.plane < lo_n || F.plane >= dims[F.normal] || F.a0 < lo_a || F.a1 >= dims[a] || F.a0 > F.a1 ||
            F.b0 < lo_b || F.b1 >= dims[b] || F.b0 > F.b1)
            return fail("NF2FF face %d outside the grid", q);
        FaceDev& D = c->faces.f[q];
        D.normal = F.normal; D.plane = F.plane; D.a0 = F.a0; D.a1 = F.a1; D.b0 = F.b0; D.b1 = F.b1; D.acc = F.acc; D.td = nullptr;
        const int nn = (F.a1 - F.a0 + 1) * (F.b1 - F.b0 + 1);
        if (nn > max_nodes) max_nodes = nn;
    }
    c->nf_max_nodes = max_nodes; c->nf_nfreq = nfreq; c->nf_interval = interval; c->nf_dt = dt; c->nf_td_max = 0;
    if (upload(&c->nf_freqs, freqs, nfreq, c->stream)) return 1;
    const float* il[3] = {ilx, ily, ilz}; const float* id[3] = {idx_, idy, idz};
    for (int a = 0; a < 3; ++a) {
        const int n = a == 2 ? c->nz + 2 : dims[a];
        if (upload(&c->inv_len[a], il[a], n, c->stream)) return 1;
        if (upload(&c->inv_dual[a], id[a], n, c->stream)) return 1;
    }
    if (!c->d_faces) CK(cudaMalloc((void**)&c->d_faces, sizeof(FaceTable)));
    CK(cudaMemcpyAsync(c->d_faces, &c->faces, sizeof(FaceTable), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int b200fdtd_set_nf2ff_td(b200fdtd_ctx* c, int nfaces, float* const* td, int max_samples)
{
    if (!c) return fail("NULL ctx");
    if (nfaces != c->faces.n) return fail("set_nf2ff_td: %d faces given, %d registered (call b200fdtd_set_nf2ff first)", nfaces, c->faces.n);
    if (nfaces > 0 && (!td || max_samples < 1)) return fail("bad time-domain store arguments");
    CK(cudaSetDevice(c->device));
    drop_graph(c);
    for (int q = 0; q < nfaces; ++q) {
        if (!td[q] || ((uintptr_t)td[q] & 3)) return fail("time-domain store of face %d is NULL or misaligned", q);
        c->faces.f[q].td = td[q];
    }
    c->nf_td_max = max_samples;
    if (nfaces > 0) {
        CK(cudaMemcpyAsync(c->d_faces, &c->faces, sizeof(FaceTable), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

extern "C" int b200fdtd_nf2ff_td_dft(b200fdtd_ctx* c, int face, int nfreq, const double* freqs, int nsamples, float* out)
{
    if (!c) return fail("NULL ctx");
    if (face < 0 || face >= c->faces.n) return fail("face %d out of range", face);
    const FaceDev& F = c->faces.f[face];
    if (!F.td) return fail("no time-domain store bound for face %d (b200fdtd_set_nf2ff_td)", face);
    if (nfreq < 1 || !freqs || !out) return fail("bad arguments");
    if (nsamples < 0 || nsamples > c->nf_td_max) return fail("nsamples=%d outside the store (max %d)", nsamples, c->nf_td_max);
    CK(cudaSetDevice(c->device));
    double* d_f = nullptr;
    CK(cudaMalloc((void**)&d_f, sizeof(double) * nfreq));
    CK(cudaMemcpyAsync(d_f, freqs, sizeof(double) * nfreq, cudaMemcpyHostToDevice, c->stream));
    const long long nn = (long long)(F.a1 - F.a0 + 1) * (F.b1 - F.b0 + 1);
    nf2ff_td_dft_kernel<<<(unsigned)((4 * nn + 255) / 256), 256, 0, c->stream>>>(F.td, nn, nsamples, c->nf_interval, c->nf_dt, nfreq, d_f, out);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_f);
    if (e != cudaSuccess) return fail("nf2ff_td_dft failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int b200fdtd_get_timestep(b200fdtd_ctx* c, int64_t* ts) { if (!c || !ts) return fail("NULL argument"); *ts = c->ts; return 0; }

extern "C" int b200fdtd_set_timestep(b200fdtd_ctx* c, int64_t ts)
{
    if (!c) return fail("NULL ctx");
    if (ts < 0 || ts > 0x7fffffff) return fail("timestep out of range");
    CK(cudaSetDevice(c->device));
    int v = (int)ts;
    c->ts_lag = 0;
    CK(cudaMemcpyAsync(c->d_ts, &v, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->ts = ts;
    return 0;
}

// ---- one step, expressed as launches with a device-side step offset -----------------
static int launch_pml(b200fdtd_ctx* c, int which, int post)
{
    for (int b = 0; b < c->pml.n; ++b) {
        const PmlBoxDev& B = c->pml.b[b];
        const int tx = B.bx <= 8 ? 8 : (B.bx <= 16 ? 16 : 32);
        const int ty = 256 / tx;
        if (B.bz > 65535) return fail("PML box too tall for one launch");
        dim3 block(tx, ty), grid((B.by + ty - 1) / ty, B.bz, 3);
        pml_kernel<<<grid, block, 0, c->stream>>>(which == 0 ? cur_volt(c) : cur_curr(c), B, which, post, c->px, c->sz, c->cs);
        CKL();
    }
    return 0;
}
static int launch_mur(b200fdtd_ctx* c, int phase)
{
    if (c->n_mur == 0) return 0;
    if (c->n_mur_segs > 0) {
        mur_seg_kernel<<<(unsigned)(((long long)c->n_mur_segs * 32 + 127) / 128), 128, 0, c->stream>>>(cur_volt(c), c->mur_segs, c->n_mur_segs,
                                                                                                       c->mur_coeff, c->mur_tmp, phase);
        CKL();
        return 0;
    }
    const int threads = 256;
    mur_kernel<<<(unsigned)((c->n_mur + threads - 1) / threads), threads, 0, c->stream>>>(cur_volt(c), c->mur_dst, c->mur_src,
                                                                                       c->mur_coeff, c->mur_tmp, c->n_mur, phase);
    CKL();
    return 0;
}
static int launch_excite(b200fdtd_ctx* c, int ts_off)
{
    if (c->n_exc == 0) return 0;
    const int threads = 128;
    excite_kernel<<<(unsigned)((c->n_exc + threads - 1) / threads), threads, 0, c->stream>>>(cur_volt(c), c->exc_idx, c->exc_amp,
        c->exc_delay, c->exc_sig, c->exc_siglen, c->n_exc, c->d_ts, ts_off + c->ts_lag);
    CKL();
    return 0;
}
// pipelined = true: the E update of the NEXT step has already been done (fused H->E launches); the voltages the sample
// wants are still intact in the other copy, which that launch only read
static int launch_sampling(b200fdtd_ctx* c, int ts_off, bool pipelined = false)
{
    ts_off += c->ts_lag;
    const float* sv = pipelined ? oth_volt(c) : cur_volt(c);
    if (c->n_probes > 0) {
        probe_kernel<<<c->n_probes, 128, 0, c->stream>>>(sv, cur_curr(c), c->pr_kind, c->pr_off, c->pr_idx, c->pr_w,
            c->interval, c->max_samples, c->pr_series, c->pr_nfreq, c->pr_freqs, c->pr_dft, c->dt, c->d_ts, ts_off);
        CKL();
    }
    if (c->faces.n > 0) {
        Nf2ffParams P;
        P.volt = sv; P.curr = cur_curr(c); P.ny = c->ny; P.px = c->px; P.sz = c->sz; P.cs = c->cs;
        for (int a = 0; a < 3; ++a) { P.il[a] = c->inv_len[a]; P.idl[a] = c->inv_dual[a]; }
        P.nfreq = c->nf_nfreq; P.freqs = c->nf_freqs; P.dt = c->nf_dt; P.d_ts = c->d_ts; P.ts_off = ts_off;
        P.interval = c->nf_interval; P.td_max = c->nf_td_max;
        dim3 grid((c->nf_max_nodes + 127) / 128, c->faces.n);
        nf2ff_kernel<<<grid, 128, sizeof(float) * 4 * c->nf_nfreq, c->stream>>>(c->d_faces, P);
        CKL();
    }
    return 0;
}
static int launch_ts_add(b200fdtd_ctx* c, int n)
{
    ts_add_kernel<<<1, 32, 0, c->stream>>>(c->d_ts, n);
    CKL();
    return 0;
}
// the device counter catches up with the host's (before anything that relies on their being equal: graph replays, resets)
static int flush_ts_lag(b200fdtd_ctx* c)
{
    if (c->ts_lag == 0) return 0;
    const int n = c->ts_lag; c->ts_lag = 0;
    return launch_ts_add(c, n);
}

// E half step with device step offset `off` (host knows ts + off)
static int e_half(b200fdtd_ctx* c, int off)
{
    if (!c->plan.valid) if (build_plan(c)) return 1;
    const bool side = (c->plan.nfused > 0 || c->plan.xedge) && (c->variant & 2) == 0;
    if (launch_mur(c, 0)) return 1;              // Mur sees the true field, before any PML pass swaps in the flux
    if (side) { if (fork_side(c)) return 1; if (launch_slabs(c, 0, 0, c->nz, c->side, c->side, slab_stream(c))) return 1; }
    if (launch_pml(c, 0, 0)) return 1;
    if (launch_volume_plain(c, 0, 0, c->nz, c->stream)) return 1;
    if (!side) { if (launch_slabs(c, 0, 0, c->nz, c->stream, c->stream, c->stream)) return 1; }
    if (launch_pml(c, 0, 1)) return 1;
    if (side) if (join_side(c)) return 1;
    if (c->flip) c->vcur ^= 1;                    // the new E lives in the other copy
    if (launch_mur(c, 1)) return 1;
    if (launch_excite(c, off)) return 1;
    if (launch_mur(c, 2)) return 1;
    return 0;
}
static int h_half(b200fdtd_ctx* c)
{
    if (!c->plan.valid) if (build_plan(c)) return 1;
    const bool side = (c->plan.nfused > 0 || c->plan.xedge) && (c->variant & 2) == 0;
    if (side) { if (fork_side(c)) return 1; if (launch_slabs(c, 1, 0, c->nz, c->side, c->side, slab_stream(c))) return 1; }
    if (launch_pml(c, 1, 0)) return 1;
    if (launch_volume_plain(c, 1, 0, c->nz, c->stream)) return 1;
    if (!side) { if (launch_slabs(c, 1, 0, c->nz, c->stream, c->stream, c->stream)) return 1; }
    if (launch_pml(c, 1, 1)) return 1;
    if (side) if (join_side(c)) return 1;
    if (c->flip) c->ccur ^= 1;
    if (c->flux_pp) c->fcur ^= 1;                 // the whole-row slab launches read one copy of the current flux and wrote the other
    return 0;
}

// H update of step n and E update of step n+1 in one sweep (graph chunks only; see update_he_kernel):
//   PML slabs: H update into the other copy  ->  fused H->E launch over the plain region  ->  PML slabs: E update
static int he_step(b200fdtd_ctx* c, int off)
{
    const bool inhe = he_pml(c);                 // whole-row slabs are swept by the fused launch itself
    const bool slabs = (c->plan.nfused > 0 && !inhe) || c->plan.xedge;
    const bool side = slabs && (c->variant & 2) == 0;
    if (launch_mur(c, 0)) return 1;              // Mur reads the old E
    c->flip = true;
    int rc = 0;
    const bool overlap = side && !inhe && (c->variant & (1 << 20)) != 0;      // experiment (measured 1.5 % slower: the fused launch fills every SM)
    do {
        if (overlap) {
            // Only the slabs at the LOW ends feed the fused launch (its halo cells outside the region are at i-1, j-1, k-1), and
            // their E update needs nothing from it (their own low-side neighbours are low-end slabs again).  So: low-end H,
            // then the fused launch with the high-end H and the low-end E launches beside it, then the high-end E.
            if ((rc = fork_side(c))) break;
            if ((rc = launch_volume_xslabs(c, 1, 0, c->nz, c->side, 1))) break;
            if ((rc = launch_volume_fused(c, 1, 0, c->nz, slab_stream(c), 1))) break;
            if ((rc = join_side(c))) break;
            if ((rc = fork_side(c))) break;                                  // side streams continue from here, not from the fused launch
            if ((rc = launch_volume_xslabs(c, 1, 0, c->nz, c->side, 2))) break;
            if ((rc = launch_volume_fused(c, 1, 0, c->nz, slab_stream(c), 2))) break;
            if ((rc = launch_he(c, c->stream))) break;
            c->ccur ^= 1;                                                    // pointers of the launches below: H is new
            if (c->flux_pp) c->fcur ^= 1;
            if ((rc = launch_volume_xslabs(c, 0, 0, c->nz, c->side, 1))) break;
            if ((rc = launch_volume_fused(c, 0, 0, c->nz, slab_stream(c), 1))) break;
            if ((rc = join_side(c))) break;
            if ((rc = fork_side(c))) break;
            if ((rc = launch_volume_xslabs(c, 0, 0, c->nz, c->side, 2))) break;
            if ((rc = launch_volume_fused(c, 0, 0, c->nz, slab_stream(c), 2))) break;
            if ((rc = join_side(c))) break;
            c->vcur ^= 1;
            break;
        }
        if (inhe || slabs_merged(c)) {
            // H of the slabs the fused launch does not sweep (one merged launch), the fused launch (it reads their H_new on
            // its halo), E of those slabs: three launches on one stream
            if ((rc = launch_slabs(c, 1, 0, c->nz, c->stream, c->stream, c->stream, !inhe))) break;
            if ((rc = launch_he(c, c->stream))) break;
            c->ccur ^= 1; if (c->flux_pp) c->fcur ^= 1;
            if ((rc = launch_slabs(c, 0, 0, c->nz, c->stream, c->stream, c->stream, !inhe))) break;
            c->vcur ^= 1;
            break;
        }
        if (side) { if ((rc = fork_side(c))) break; }
        if ((rc = launch_volume_xslabs(c, 1, 0, c->nz, side ? c->side : c->stream))) break;
        if ((rc = launch_volume_fused(c, 1, 0, c->nz, side ? slab_stream(c) : c->stream))) break;
        if (side) { if ((rc = join_side(c))) break; }
        if ((rc = launch_he(c, c->stream))) break;
        c->ccur ^= 1;                            // H is new from here on
        if (c->flux_pp) c->fcur ^= 1;
        if (side) { if ((rc = fork_side(c))) break; }
        if ((rc = launch_volume_xslabs(c, 0, 0, c->nz, side ? c->side : c->stream))) break;
        if ((rc = launch_volume_fused(c, 0, 0, c->nz, side ? slab_stream(c) : c->stream))) break;
        if (side) { if ((rc = join_side(c))) break; }
        c->vcur ^= 1;
    } while (0);
    c->flip = false;
    if (rc) return 1;
    if (launch_mur(c, 1)) return 1;
    if (launch_excite(c, off)) return 1;
    return launch_mur(c, 2);
}

// n consecutive steps with no sampling point inside (device step counter untouched: launches use offsets 0..n-1).
//   unfused: E(0) H(0) E(1) H(1) ...
//   fused  : E(0) | H(0)+E(1) | ... | H(n-2)+E(n-1) | H(n-1): each fused launch flips both field copies; the two unfused
//            half steps at the ends flip too when the number of fused launches is odd, so the span ends where it began
static int run_span(b200fdtd_ctx* c, int n, bool fuse)
{
    int rc = 0;
    if (!fuse || n < 2) {
        for (int s = 0; s < n && !rc; ++s) {
            rc = e_half(c, s);
            if (!rc) rc = h_half(c);
        }
        if (!rc) rc = normalize_flux(c);
        return rc;
    }
    const bool odd = ((n - 1) & 1) != 0;
    c->vcur = c->ccur = 0;
    c->flip = odd; rc = e_half(c, 0); c->flip = false;
    for (int s = 1; s < n && !rc; ++s) rc = he_step(c, s);
    if (!rc) { c->flip = odd; rc = h_half(c); c->flip = false; }
    if (!rc && (c->vcur || c->ccur)) rc = fail("fused span did not return to the bound field arrays");
    c->vcur = c->ccur = 0; c->flip = false;
    if (!rc) rc = normalize_flux(c);            // an odd number of H passes: the current flux goes back to the caller's arrays
    return rc;
}

static int run_eager(b200fdtd_ctx* c, int64_t n)
{
    const int iv = sample_interval(c);
    bool ghosts = false;
    while (n > 0) {
        int64_t span = n < 64 ? n : 64;
        if (iv > 0) { const int64_t to_sample = iv - (c->ts % iv); if (span > to_sample) span = to_sample; }
        const bool fuse = he_ready(c, (int)span);
        if (fuse) { c->he_fused = true; if (!ghosts) { if (sync_alt_ghosts(c)) return 1; ghosts = true; } }
        if (run_span(c, (int)span, fuse)) return 1;
        if (launch_ts_add(c, (int)span)) return 1;
        c->ts += span; n -= span;
        if (iv > 0 && (c->ts % iv) == 0) if (launch_sampling(c, 0)) return 1;
    }
    return 0;
}

static int build_graph(b200fdtd_ctx* c, int steps)
{
    drop_graph(c);
    const int iv = sample_interval(c);
    cudaGraph_t g = nullptr;
    const int64_t before = g_launches.load();
    const bool fuse = he_ready(c, steps);           // allocates the second field copy: before the capture starts
    if (fuse) c->he_fused = true;
    c->graph_fused = fuse;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = run_span(c, steps, fuse);
    if (!rc && iv > 0) rc = launch_sampling(c, steps);      // graph starts at ts % iv == 0 and spans iv steps
    if (!rc) rc = launch_ts_add(c, steps);
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return 1; }
    if (e != cudaSuccess) return fail("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&c->graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { c->graph = nullptr; return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
    c->graph_steps = steps;
    c->graph_kernels = g_launches.load() - before;      // kernel nodes in the graph
    g_launches.fetch_sub(c->graph_kernels);             // capturing is not launching
    return 0;
}

// ---- pipelined stepping ---------------------------------------------------------------------------------------------
// A whole run call is ONE fused span:  E(0) | H(0)+E(1) | H(1)+E(2) | ... | H(n-1).  A sampling point inside it falls
// between the H update of step s-1 and the E update of step s, which a fused launch has done in one sweep: the sample
// takes H from the new copy and E from the copy that launch only read (launch_sampling(pipelined)).  So the two unfused
// half steps are paid once per call (~1 % of a run), not once per sampling interval (4 steps on the 10 GHz vacuum cube,
// 13 on the reference scene).  Whole sampling intervals are replayed from a CUDA graph of g fused steps (g = interval, or
// twice that if it is odd: the field copies swap roles every step, so a graph must span an even number; one graph per
// state of the copies at its entry).  When the number of fused steps of a call is odd, the two unfused half steps at its
// ends write the other copy as well, so the state always ends in the bound arrays.
static int build_pgraph(b200fdtd_ctx* c, int g, int iv, int parity)
{
    cudaGraph_t gr = nullptr;
    const int64_t before = g_launches.load();
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    const int v0 = c->vcur, c0 = c->ccur, f0 = c->fcur;
    for (int q = 1; q <= g && !rc; ++q) {
        rc = he_step(c, q);
        if (!rc && iv > 0 && (q % iv) == 0) rc = launch_sampling(c, q, true);      // the chunk starts on an interval boundary
    }
    if (!rc) rc = launch_ts_add(c, g);
    cudaError_t e = cudaStreamEndCapture(c->stream, &gr);
    c->vcur = v0; c->ccur = c0; c->fcur = f0;               // g is even: the copies are back where they were
    if (rc) { if (gr) cudaGraphDestroy(gr); return 1; }
    if (e != cudaSuccess) return fail("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&c->pgraph[parity], gr, 0);
    cudaGraphDestroy(gr);
    if (e != cudaSuccess) { c->pgraph[parity] = nullptr; return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
    c->pgraph_kernels[parity] = g_launches.load() - before;
    g_launches.fetch_sub(c->pgraph_kernels[parity]);        // capturing is not launching
    return 0;
}

static int run_pipelined(b200fdtd_ctx* c, int64_t n, bool use_graph)
{
    const int iv = sample_interval(c);
    const int64_t T0 = c->ts;
    int64_t dts = T0;                                        // value of the device step counter
    const int64_t F = n - 1;                                 // fused steps
    const bool odd = (F & 1) != 0;
    int g = 0;                                               // graph chunk (fused steps), 0 = no graph
    if (use_graph && c->stream != nullptr && iv <= 4096) { g = iv > 0 ? iv : 16; if (g & 1) g *= 2; }
    if (g != c->pgraph_steps) { for (int q = 0; q < 4; ++q) if (c->pgraph[q]) { cudaGraphExecDestroy(c->pgraph[q]); c->pgraph[q] = nullptr; } c->pgraph_steps = g; }
    c->he_fused = F > 0;
    if (sync_alt_ghosts(c)) return 1;
    c->vcur = c->ccur = 0;
    c->flip = odd; { const int rc = e_half(c, 0); c->flip = false; if (rc) return 1; }     // E(T0): into the pipelined state
    int64_t q = 0;
    while (q < F) {
        if (g > 0 && ((T0 + q) % g) == 0 && F - q >= g) {
            const int par = 2 * c->vcur + c->ccur;
            if (dts != T0 + q) { if (launch_ts_add(c, (int)(T0 + q - dts))) return 1; dts = T0 + q; }
            if (!c->pgraph[par]) if (build_pgraph(c, g, iv, par)) return 1;
            CK(cudaGraphLaunch(c->pgraph[par], c->stream));
            g_launches.fetch_add(c->pgraph_kernels[par], std::memory_order_relaxed);
            dts += g; q += g;
            continue;
        }
        if (he_step(c, (int)(T0 + q + 1 - dts))) return 1;  // H(T0+q) + E(T0+q+1)
        ++q;
        if (iv > 0 && ((T0 + q) % iv) == 0) if (launch_sampling(c, (int)(T0 + q - dts), true)) return 1;
    }
    c->flip = odd; { const int rc = h_half(c); c->flip = false; if (rc) return 1; }          // H(T0+F): out of the pipelined state
    if (c->vcur || c->ccur) return fail("pipelined run did not return to the bound field arrays");
    if (iv > 0 && ((T0 + n) % iv) == 0) if (launch_sampling(c, (int)(T0 + n - dts))) return 1;
    if (normalize_flux(c)) return 1;
    if (T0 + n != dts) if (launch_ts_add(c, (int)(T0 + n - dts))) return 1;
    c->ts = T0 + n;
    return 0;
}

extern "C" int b200fdtd_run(b200fdtd_ctx* c, int64_t nsteps, int use_graph)
{
    if (!c) return fail("NULL ctx");
    if (nsteps < 0) return fail("nsteps < 0");
    if (!c->volt || !c->vv) return fail("fields/coefficients not bound");
    CK(cudaSetDevice(c->device));
    if (!c->plan.valid) if (build_plan(c)) return 1;        // never inside a stream capture
    if (flush_ts_lag(c)) return 1;
    c->he_fused = false;
    const int iv = sample_interval(c);
    if (nsteps == 0) return 0;
    // fused H->E launches available (second field copy, every PML box in the volume launches): one pipelined span per call
    // (variant bit 27: one span per sampling interval, the round-1 schedule)
    if (nsteps >= 3 && (c->variant & (1 << 27)) == 0 && he_ready(c, (int)(nsteps > 1000000 ? 1000000 : nsteps)))
        return run_pipelined(c, nsteps, use_graph != 0);
    // the NULL stream cannot be captured; a sampling interval of thousands of steps (tiny cells -> tiny time step) is not
    // worth a graph of that many nodes
    if (!use_graph || c->stream == nullptr || iv > 4096) return run_eager(c, nsteps);
    const int chunk = iv > 0 ? iv : 16;
    int64_t left = nsteps;
    // align to a chunk boundary
    const int64_t mis = c->ts % chunk;
    if (mis != 0) {
        const int64_t n = (chunk - mis) < left ? (chunk - mis) : left;
        if (run_eager(c, n)) return 1;
        left -= n;
    }
    if (left >= chunk) {
        if (!c->graph || c->graph_steps != chunk) if (build_graph(c, chunk)) return 1;
        if (c->graph_fused) { c->he_fused = true; if (sync_alt_ghosts(c)) return 1; }
        while (left >= chunk) {
            CK(cudaGraphLaunch(c->graph, c->stream));
            g_launches.fetch_add(c->graph_kernels, std::memory_order_relaxed);
            c->ts += chunk; left -= chunk;
        }
    }
    if (left > 0) if (run_eager(c, left)) return 1;
    return 0;
}

extern "C" int b200fdtd_half_step(b200fdtd_ctx* c, int phase)
{
    if (!c) return fail("NULL ctx");
    if (!c->volt || !c->vv) return fail("fields/coefficients not bound");
    CK(cudaSetDevice(c->device));
    if (phase == 0) return e_half(c, 0);
    if (phase == 1) {
        if (h_half(c)) return 1;
        c->ts += 1; c->ts_lag += 1;
        return 0;
    }
    if (phase == 2 || phase == 3) {                          // 3: in the pipelined state of the fused steps (E is one update ahead)
        const int iv = sample_interval(c);
        if (phase == 3 && !c->alt_volt) return fail("pipelined sampling needs the second field copy");
        if (iv > 0 && (c->ts % iv) == 0) return launch_sampling(c, 0, phase == 3);
        return 0;
    }
    return fail("phase must be 0, 1, 2 or 3");
}

// z-slab overlap: each half step in two parts so a halo exchange can hide behind the interior launch.
//   phase 0 (E): part 0 = pre passes + planes [1,nz)   | part 1 = plane 0 (needs the lower ghost H) + post passes
//   phase 1 (H): part 0 = pre passes + planes [0,nz-1) | part 1 = plane nz-1 (needs the upper ghost E) + post, ++ts
extern "C" int b200fdtd_half_step_part(b200fdtd_ctx* c, int phase, int part)
{
    if (!c) return fail("NULL ctx");
    if (!c->volt || !c->vv) return fail("fields/coefficients not bound");
    if ((phase != 0 && phase != 1) || (part != 0 && part != 1)) return fail("phase and part must be 0 or 1");
    CK(cudaSetDevice(c->device));
    if (!c->plan.valid) if (build_plan(c)) return 1;
    const int nz = c->nz;
    if (phase == 0) {
        if (part == 0) {
            if (launch_mur(c, 0)) return 1;
            if (launch_pml(c, 0, 0)) return 1;
            return launch_volume(c, 0, 1, nz);
        }
        if (launch_volume(c, 0, 0, 1)) return 1;
        if (launch_pml(c, 0, 1)) return 1;
        if (launch_mur(c, 1)) return 1;
        if (launch_excite(c, 0)) return 1;
        return launch_mur(c, 2);
    }
    if (part == 0) {
        if (launch_pml(c, 1, 0)) return 1;
        return launch_volume(c, 1, 0, nz - 1);
    }
    if (launch_volume(c, 1, nz - 1, nz)) return 1;
    if (launch_pml(c, 1, 1)) return 1;
    if (c->flux_pp) c->fcur ^= 1;                 // both parts read one copy of the slabs' current flux and wrote the other
    c->ts += 1; c->ts_lag += 1;
    return 0;
}

// ---- fused H->E step on a z-slab rank -------------------------------------------------------------------------------
// The fused launch cannot cover a rank's boundary planes: E_new of plane 0 needs the lower neighbour's H_new (computed in
// the same sweep over there), H_new of the top plane needs the upper neighbour's E.  So the two boundary planes keep the
// separate H and E launches (they write the other field copy like the fused launch does) and the fused launch covers the
// interior planes [1, nz-1); it reads H_new of plane 0 from the output copy like any other halo cell outside its region.
//   part 0: Mur pre; H of plane 0 and of the PML slabs of the interior planes            (needs no ghost plane)
//   part 1: H of the top plane                          (needs the upper ghost E; caller then sends H_new(top) up)
//   part 4: (optional, between 0 and 1) fused H->E launch over the lower half of the interior planes, so that the wait for
//           the upper ghost E hides behind it
//   part 2: fused H->E launch over the (rest of the) interior planes; H is new from here on; ++ts   (overlaps that exchange)
//   part 3: E of the PML slabs of the interior planes, of plane 0 (needs the lower ghost H_new) and of the top plane; E is
//           new from here on; Mur post, excitation, Mur apply                            (caller then sends E_new(0) down)
// b200fdtd_current_copy tells which copy (0 = bound arrays, 1 = second copy) holds E and H afterwards.
extern "C" int b200fdtd_fused_step_part(b200fdtd_ctx* c, int part)
{
    if (!c) return fail("NULL ctx");
    if (!c->volt || !c->vv) return fail("fields/coefficients not bound");
    if (!c->alt_volt || !c->alt_curr) return fail("bind the second field copy first (b200fdtd_bind_alt_fields)");
    if (part < 0 || part > 4) return fail("part must be 0..4");
    CK(cudaSetDevice(c->device));
    if (!c->plan.valid) if (build_plan(c)) return 1;
    if (c->pml.n > 0) return fail("fused steps need every PML box fused into the volume launches");
    const int nz = c->nz;
    if (nz < 3) return fail("fused steps need at least 3 planes per slab");
    int rc = 0;
    if (ensure_flux_pp(c)) return 1;
    const bool inhe = he_pml(c);                 // the fused launch sweeps the whole-row slabs of the interior planes itself
    const bool side = ((c->plan.nfused > 0 && !inhe) || c->plan.xedge) && (c->variant & 2) == 0;    // interior slab launches side by side
    if (part == 0) {
        if (launch_mur(c, 0)) return 1;
        c->flip = true;
        if (side) rc = fork_side(c);
        if (!rc) rc = launch_slabs(c, 1, 1, nz - 1, side ? c->side : c->stream, side ? c->side : c->stream, side ? slab_stream(c) : c->stream, !inhe);
        if (!rc) rc = launch_volume(c, 1, 0, 1);
        if (!rc && side) rc = join_side(c);
        c->flip = false;
        return rc;
    }
    if (part == 1) {
        c->flip = true;
        rc = launch_volume(c, 1, nz - 1, nz);
        c->flip = false;
        return rc;
    }
    if (part == 4) {                                         // optional: lower half of the interior planes first
        c->he_mid = 1 + (nz - 2) / 2;
        return launch_he(c, c->stream, 1, c->he_mid);
    }
    if (part == 2) {
        if (launch_he(c, c->stream, c->he_mid > 0 ? c->he_mid : 1, nz - 1)) return 1;
        c->he_mid = 0;
        c->ccur ^= 1;
        if (c->flux_pp) c->fcur ^= 1;
        c->ts += 1; c->ts_lag += 1;
        return 0;
    }
    c->flip = true;
    if (side) rc = fork_side(c);
    if (!rc) rc = launch_slabs(c, 0, 1, nz - 1, side ? c->side : c->stream, side ? c->side : c->stream, side ? slab_stream(c) : c->stream, !inhe);
    if (!rc) rc = launch_volume_two_planes(c, 0, 0, nz - 1);
    if (!rc && side) rc = join_side(c);
    c->flip = false;
    if (rc) return 1;
    c->vcur ^= 1;
    if (launch_mur(c, 1)) return 1;
    if (launch_excite(c, 0)) return 1;
    return launch_mur(c, 2);
}

extern "C" int b200fdtd_bind_alt_fields(b200fdtd_ctx* c, float* volt2, float* curr2)
{
    if (!c) return fail("NULL ctx");
    if ((volt2 == nullptr) != (curr2 == nullptr)) return fail("bind both copies or none");
    if (((uintptr_t)volt2 | (uintptr_t)curr2) & 15) return fail("field pointers must be 16-byte aligned");
    if (c->vcur || c->ccur) return fail("the state lives in the second copy: normalise first (b200fdtd_reset_current_copy)");
    CK(cudaSetDevice(c->device));
    drop_graph(c);
    if (c->alt_owned) { cudaFree(c->alt_volt); cudaFree(c->alt_curr); c->alt_owned = false; }
    c->alt_volt = volt2; c->alt_curr = curr2;
    return 0;
}

extern "C" int b200fdtd_current_copy(b200fdtd_ctx* c, int* vcur, int* ccur)
{
    if (!c || !vcur || !ccur) return fail("NULL argument");
    *vcur = c->vcur; *ccur = c->ccur;
    return 0;
}

extern "C" int b200fdtd_reset_current_copy(b200fdtd_ctx* c)
{
    if (!c) return fail("NULL ctx");
    c->vcur = c->ccur = 0;                       // the caller has copied the state back into the bound arrays
    CK(cudaSetDevice(c->device));
    return normalize_flux(c);                    // ... and the library does the same for the PML flux copy it owns
}

extern "C" int b200fdtd_update_only(b200fdtd_ctx* c, int which)
{
    if (!c) return fail("NULL ctx");
    if (which < 0 || which > 4) return fail("which must be 0..4");
    CK(cudaSetDevice(c->device));
    if (which == 4) {
        // only the fused H->E launch over the plain region, for timing: it reads the bound arrays and writes the second
        // copy, so the state of the run is untouched
        if (!c->volt || !c->vv) return fail("fields/coefficients not bound");
        if (!c->plan.valid) if (build_plan(c)) return 1;
        if (!he_ready(c, 2)) return fail("fused H->E launch not available for this set-up");
        return launch_he(c, c->stream);
    }
    if (which < 2) {
        if (launch_volume(c, which, 0, c->nz)) return 1;
        if (which == 1 && c->flux_pp) c->fcur ^= 1;
        return 0;
    }
    // 2/3: only the plain (non-PML) launch of the E/H update — the kernel the roofline is quoted on
    if (!c->volt || !c->vv) return fail("fields/coefficients not bound");
    if (!c->plan.valid) if (build_plan(c)) return 1;
    return launch_volume_plain(c, which - 2, 0, c->nz, c->stream);
}

extern "C" int b200fdtd_plan_info(b200fdtd_ctx* c, int64_t* plain_cells, int64_t* fused_cells, int64_t* separate_cells)
{
    if (!c || !plain_cells || !fused_cells || !separate_cells) return fail("NULL argument");
    if (!c->plan.valid) if (build_plan(c)) return 1;
    const VolumePlan& P = c->plan;
    int64_t skip = 0, fused = 0, sep = 0, planes = 0;
    for (int s = 0; s < P.nskip; ++s) skip += P.sj1[s] - P.sj0[s];
    for (int s = 0; s < P.nseg; ++s) planes += P.seg1[s] - P.seg0[s];
    for (int q = 0; q < P.nfused; ++q) fused += (int64_t)c->px * P.fb[q].by * P.fb[q].bz;
    if (P.has_lo) fused += (int64_t)P.xw0 * P.xlo.by * P.xlo.bz;
    if (P.has_hi) fused += (int64_t)P.xw1 * P.xhi.by * P.xhi.bz;
    for (int b = 0; b < c->pml.n; ++b) sep += (int64_t)c->pml.b[b].bx * c->pml.b[b].by * c->pml.b[b].bz;
    *plain_cells = (int64_t)c->px * (c->ny - skip) * planes;       // cells (incl. pad columns) swept by the plain launch
    *fused_cells = fused; *separate_cells = sep;
    return 0;
}

extern "C" int b200fdtd_set_he_tuning(b200fdtd_ctx* c, int rows, int planes)
{
    if (!c) return fail("NULL ctx");
    if (!(rows == 0 || rows == 3 || rows == 7 || rows == 15)) return fail("rows must be 0, 3, 7 or 15");
    if (planes < 0) return fail("planes must be >= 0");
    const int de = (planes >> 24) & 3; planes &= 0xffffff;     // bits 24-25 of `planes`: E planes landing ahead (1 | 2; 0 = automatic)
    if (de == 3) return fail("the E lookahead must be 0 (automatic), 1 or 2");
    if (rows) c->he_ty = rows;
    if (planes) c->he_kz = planes;
    c->he_de = de;
    drop_graph(c);
    return 0;
}

extern "C" int b200fdtd_he_info(b200fdtd_ctx* c, int* active)
{
    if (!c || !active) return fail("NULL argument");
    *active = c->he_fused ? 1 : 0;
    return 0;
}

extern "C" int b200fdtd_pml_compression_info(b200fdtd_ctx* c, int64_t* rows_compressed, int64_t* rows_demoted)
{
    if (!c || !rows_compressed || !rows_demoted) return fail("NULL argument");
    *rows_compressed = c->pml_rows_compressed; *rows_demoted = c->pml_rows_demoted;
    return 0;
}

extern "C" int b200fdtd_energy(b200fdtd_ctx* c, double* energy)
{
    if (!c || !energy) return fail("NULL argument");
    if (!c->volt) return fail("fields not bound");
    CK(cudaSetDevice(c->device));
    const long long n_owned = (long long)c->nz * c->sz;
    energy_partial_kernel<<<c->n_partials, 256, 0, c->stream>>>(cur_volt(c), cur_curr(c), c->sz, c->cs, n_owned, c->d_partials);
    CKL();
    energy_final_kernel<<<1, 32, 0, c->stream>>>(c->d_partials, c->n_partials, c->d_energy);
    CKL();
    double h[2];
    CK(cudaMemcpyAsync(h, c->d_energy, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const double EPS0 = 8.85418781762e-12, MUE0 = 1.256637062e-6;
    *energy = 0.5 * EPS0 * h[0] + 0.5 * MUE0 * h[1];
    return 0;
}

extern "C" int b200fdtd_sync(b200fdtd_ctx* c)
{
    if (!c) return fail("NULL ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int b200fdtd_num_samples(b200fdtd_ctx* c, int* n)
{
    if (!c || !n) return fail("NULL argument");
    const int iv = sample_interval(c);
    int s = iv > 0 ? (int)(c->ts / iv) : 0;
    if (c->n_probes > 0 && s > c->max_samples) s = c->max_samples;
    *n = s;
    return 0;
}

extern "C" int b200fdtd_nf2ff_sources(int device, void* stream, int nfaces, const b200fdtd_nf2ff_src_face* faces, double scale,
                                       const double* center, float* pos, float* J, float* M, double* prad)
{
    if (nfaces < 0 || nfaces > 16 || (nfaces > 0 && !faces) || !center || !prad) return fail("bad nf2ff_sources arguments");
    CK(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    long long npts = 0, nblocks = 0; size_t nd = 0;
    for (int q = 0; q < nfaces; ++q) {
        const b200fdtd_nf2ff_src_face& F = faces[q];
        if (F.na < 1 || F.nb < 1 || F.normal < 0 || F.normal > 2 || !F.acc || !F.xa || !F.xb || !F.wa || !F.wb) return fail("bad source face %d", q);
        npts += (long long)F.na * F.nb; nblocks += ((long long)F.na * F.nb + 255) / 256; nd += 2 * ((size_t)F.na + F.nb);
    }
    *prad = 0.0;
    if (npts == 0) return 0;
    if (!pos || !J || !M) return fail("NULL output array");
    // one staging buffer: the faces' line coordinates and weights, then the per-block Prad partials
    std::vector<double> h(nd);
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, sizeof(double) * (nd + (size_t)nblocks)));
    size_t o = 0; long long off = 0, b0 = 0;
    std::vector<SrcFace> L(nfaces);
    for (int q = 0; q < nfaces; ++q) {
        const b200fdtd_nf2ff_src_face& F = faces[q];
        SrcFace& S = L[q];
        S.normal = F.normal; S.side = F.side; S.na = F.na; S.nb = F.nb; S.coord = F.coord; S.acc = F.acc; S.cstride = F.comp_stride;
        memcpy(&h[o], F.xa, sizeof(double) * F.na); S.xa = d + o; o += F.na;
        memcpy(&h[o], F.xb, sizeof(double) * F.nb); S.xb = d + o; o += F.nb;
        memcpy(&h[o], F.wa, sizeof(double) * F.na); S.wa = d + o; o += F.na;
        memcpy(&h[o], F.wb, sizeof(double) * F.nb); S.wb = d + o; o += F.nb;
        S.off = off; off += (long long)F.na * F.nb;
    }
    cudaError_t e = cudaMemcpyAsync(d, h.data(), sizeof(double) * nd, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);                    // h is a pageable temporary
    for (int q = 0; q < nfaces && e == cudaSuccess; ++q) {
        const long long nb = ((long long)L[q].na * L[q].nb + 255) / 256;
        nf2ff_sources_kernel<<<(unsigned)nb, 256, 0, s>>>(L[q], scale, center[0], center[1], center[2], npts, pos, J, M, d + nd + b0);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
        b0 += nb;
    }
    std::vector<double> part((size_t)nblocks);
    if (e == cudaSuccess) e = cudaMemcpyAsync(part.data(), d + nd, sizeof(double) * nblocks, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    if (e != cudaSuccess) return fail("nf2ff_sources failed: %s", cudaGetErrorString(e));
    double t = 0; for (double v : part) t += v;
    *prad = t;
    return 0;
}

extern "C" int b200fdtd_farfield(int device, void* stream, int64_t npts, const float* pos, const float* J,
                                  const float* M, double k, int ndir, const double* theta, const double* phi, float* out)
{
    if (npts <= 0 || ndir <= 0 || !pos || !J || !M || !theta || !phi || !out) return fail("bad far-field arguments");
    CK(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    // direction + partial-sum buffer of the calling thread, grown on demand and kept (the reference calls CalcNF2FF once per phi: 73 calls)
    static thread_local double* d_dir = nullptr; static thread_local int d_cap = 0, d_dev = -1;
    const int target = 148 * 4, ngroups = (ndir + 7) / 8;                            // a block = 8 directions (one per warp)
    int nsplit = (target + ngroups - 1) / ngroups;
    if ((long long)nsplit * 2048 > npts) nsplit = (int)((npts + 2047) / 2048);       // at least ~2 k points per block
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 64) nsplit = 64;                                                    // (partial-sum buffer: 64 shares per direction)
    if (d_dev != device || d_cap < ndir) {
        if (d_dir && d_dev >= 0) { cudaSetDevice(d_dev); cudaFree(d_dir); cudaSetDevice(device); }
        d_dir = nullptr; d_cap = 0; d_dev = device;
        const int cap = ndir < 8192 ? 8192 : ndir;
        CK(cudaMalloc((void**)&d_dir, sizeof(double) * (2 * (size_t)cap + 12ULL * 64 * cap)));
        d_cap = cap;
    }
    double* d_part = d_dir + 2 * (size_t)d_cap;
    CK(cudaMemcpyAsync(d_dir, theta, sizeof(double) * ndir, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_dir + d_cap, phi, sizeof(double) * ndir, cudaMemcpyHostToDevice, s));
    farfield_kernel<<<dim3(ngroups, nsplit), 256, 0, s>>>(npts, pos, J, M, k, ndir, d_dir, d_dir + d_cap, d_part);
    CKL();
    farfield_final_kernel<<<(ndir + 127) / 128, 128, 0, s>>>(ndir, nsplit, d_part, d_dir, d_dir + d_cap, out);
    CKL();
    CK(cudaStreamSynchronize(s));                 // theta/phi are pageable host arrays of the caller
    return 0;
}
