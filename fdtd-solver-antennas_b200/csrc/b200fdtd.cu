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
extern "C" int b200fdtd_version(void) { return 2; }
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
                  const float *xv_v, *xv_i; const unsigned char *meta_v, *meta_i; };
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
    const float *vv = nullptr, *vi = nullptr, *ii = nullptr, *iv = nullptr;
    int kz = 16, ty = 4, variant = 0;
    const float* cmp_xv[2] = {nullptr, nullptr};          // row compression tables of the E and H pass (caller-owned)
    const unsigned char* cmp_meta[2] = {nullptr, nullptr};
    int cmp_nvec[2] = {0, 0};
    // step counter
    int64_t ts = 0;
    int* d_ts = nullptr;
    // excitation
    int64_t n_exc = 0; int64_t* exc_idx = nullptr; float* exc_amp = nullptr; int* exc_delay = nullptr;
    float* exc_sig = nullptr; int exc_siglen = 0;
    // mur
    int64_t n_mur = 0; int64_t *mur_dst = nullptr, *mur_src = nullptr; float *mur_coeff = nullptr, *mur_tmp = nullptr;
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
    int nf_max_nodes = 0;
    // energy
    double* d_partials = nullptr; int n_partials = 0; double* d_energy = nullptr;
    // graph
    cudaGraphExec_t graph = nullptr; int graph_steps = 0; int64_t graph_kernels = 0;
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

// ------------------------------------------------------------------------------------
// volume kernels (K1 E update, K2 H update)
// ------------------------------------------------------------------------------------
struct VolParams {
    float* __restrict__ f;          // updated field is written here (volt for E, curr for H)
    const float* fin;               // ... and read from here (== f: in place; the other copy in a ping-pong step)
    const float* __restrict__ g;    // the other field (read only in this pass)
    const float* __restrict__ ca;   // vv / ii
    const float* __restrict__ cb;   // vi / iv
    int nx, ny, nz, px;
    long long sz, cs;
    int kz;                         // planes marched per CTA
    int k0, k1;                     // plane range [k0,k1) handled by this launch
    const float* __restrict__ xv;   // row compression: table of x-vectors [nvec][px]
    const unsigned char* __restrict__ meta;   // per row (k,j): 6 scales + 6 vector ids (32 B), see RowMeta
};

// Row compression of the operator (the openEMS "compressed operator" idea, applied per x-row): on a rectilinear mesh
// a coefficient row is very often  scale(j,k) * xvec[i]  with one of a handful of x-vectors (all vacuum rows, PML rows,
// boundary rows).  Such rows are not streamed from HBM: the kernel reads the 32-byte row record and the (L1-resident)
// x-vector and multiplies.  The full arrays stay bound and hold exactly fl32(scale*xvec) for every compressed row
// (checked on the device by verify_rows_kernel, which demotes any row that does not match bit for bit), so results are
// identical with and without compression and identical to the oracle, which reads the full arrays.
struct RowMeta { float sc[6]; unsigned char id[6]; unsigned char pad[2]; };   // slots: ca_x, ca_y, ca_z, cb_x, cb_y, cb_z
#define ROW_FULL 255u

__device__ __forceinline__ float4 coef4(unsigned id, float sc, const float* full, const float* __restrict__ xv, int i0, int px)
{
    if (id != ROW_FULL) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xv + (size_t)id * px + i0));
        return make_float4(__fmul_rn(sc, v.x), __fmul_rn(sc, v.y), __fmul_rn(sc, v.z), __fmul_rn(sc, v.w));
    }
    return __ldcs(reinterpret_cast<const float4*>(full));
}

// xv_lane_ = this lane's column of the x-vector table (hoisted out of the march), xv_pitch_ = bytes per x-vector
__device__ __forceinline__ float4 xvg4(const char* xv_lane, unsigned id, unsigned pitch, float sc)
{
    const float4 v = __ldg(reinterpret_cast<const float4*>(xv_lane + (unsigned long long)id * pitch));
    return make_float4(__fmul_rn(sc, v.x), __fmul_rn(sc, v.y), __fmul_rn(sc, v.z), __fmul_rn(sc, v.w));
}
#define LOAD_COEFFS_CMP()                                                                                   \
    do {                                                                                                    \
        const float4* m_ = reinterpret_cast<const float4*>(p.meta + ((long long)(k + 1) * p.ny + j) * 32);  \
        prefetch_l1(p.meta + ((long long)(k + 1 + KSTEP) * p.ny + j) * 32);   /* next plane's record: ghost planes exist */ \
        const float4 m0_ = __ldg(m_), m1_ = __ldg(m_ + 1);                                                  \
        const unsigned w0_ = __float_as_uint(m1_.z), w1_ = __float_as_uint(m1_.w);                          \
        if (((w1_ >> 16) & 255u) == 0) {          /* pad[0]: no slot of this row is streamed in full (row-uniform) */ \
            ax = xvg4(xv_lane_, w0_ & 255u, xv_pitch_, m0_.x);                                              \
            ay = xvg4(xv_lane_, (w0_ >> 8) & 255u, xv_pitch_, m0_.y);                                       \
            az = xvg4(xv_lane_, (w0_ >> 16) & 255u, xv_pitch_, m0_.z);                                      \
            bx = xvg4(xv_lane_, w0_ >> 24, xv_pitch_, m0_.w);                                               \
            by = xvg4(xv_lane_, w1_ & 255u, xv_pitch_, m1_.x);                                              \
            bz = xvg4(xv_lane_, (w1_ >> 8) & 255u, xv_pitch_, m1_.y);                                       \
        } else {                                                                                            \
            ax = coef4(w0_ & 255u, m0_.x, p.ca + base, p.xv, i0, p.px);                                     \
            ay = coef4((w0_ >> 8) & 255u, m0_.y, p.ca + cs + base, p.xv, i0, p.px);                         \
            az = coef4((w0_ >> 16) & 255u, m0_.z, p.ca + 2 * cs + base, p.xv, i0, p.px);                    \
            bx = coef4(w0_ >> 24, m0_.w, p.cb + base, p.xv, i0, p.px);                                      \
            by = coef4(w1_ & 255u, m1_.x, p.cb + cs + base, p.xv, i0, p.px);                                \
            bz = coef4((w1_ >> 8) & 255u, m1_.y, p.cb + 2 * cs + base, p.xv, i0, p.px);                     \
        }                                                                                                   \
    } while (0)


// the row records steer dependent loads: pulling the next plane's record into L1 one iteration ahead keeps the march
// at one DRAM round trip per plane
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4_ro(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4_nc(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// f = fmaf(ca, f, cb*(((a-b)-c)+d)) per lane of a float4
__device__ __forceinline__ float upd1(float ca, float f, float cb, float a, float b, float c, float d) {
    float curl = __fadd_rn(__fsub_rn(__fsub_rn(a, b), c), d);
    return __fmaf_rn(ca, f, __fmul_rn(cb, curl));
}
__device__ __forceinline__ float4 upd4(float4 ca, float4 f, float4 cb, float4 a, float4 b, float4 c, float4 d) {
    float4 r;
    r.x = upd1(ca.x, f.x, cb.x, a.x, b.x, c.x, d.x);
    r.y = upd1(ca.y, f.y, cb.y, a.y, b.y, c.y, d.y);
    r.z = upd1(ca.z, f.z, cb.z, a.z, b.z, c.z, d.z);
    r.w = upd1(ca.w, f.w, cb.w, a.w, b.w, c.w, d.w);
    return r;
}

// Row selection of one volume launch and, for launches over a fused PML slab, the slab arrays.
// A PML box that spans whole x-rows (x0 = 0, bx = px) is not swept by the separate pre/post passes: the
// volume kernel does  pre -> update -> post  on the values it already holds in registers (same arithmetic,
// same order per cell as the separate passes; App. A4), so volt/curr are read and written once.
struct RowParams {
    int j0, j1;                     // rows handled by this launch
    int sj0a, sj1a, sj0b, sj1b;     // rows that belong to fused y-slabs (skipped by the plain launch; empty ranges if none)
    float* flux;                    // fused slab arrays [3][bz][by][px] (PML launches only)
    const float* a; const float* fo; const float* fn;
    const float* pxv; const unsigned char* pmeta;   // row compression of a/fo/fn (48-byte records per slab row) or NULL
    int y0, z0, by, bz;
    // narrow x-slabs (columns [0,xw0) and [xx1,xx1+xw1), multiples of 4): the plain launch (MODE 0) does not store
    // these columns; a narrow launch (MODE 2, blockIdx.x = slab) owns them: a warp covers xs float4 columns of 32/xs
    // rows and does the PML pre/update/post like the fused row launch
    float* xflux0; const float* xa0; const float* xfo0; const float* xfn0; int xw0, xs0;
    float* xflux1; const float* xa1; const float* xfo1; const float* xfn1; int xx1, xw1, xs1;
    const float* pxv0; const unsigned char* pmeta0; const float* pxv1; const unsigned char* pmeta1;
};

__device__ __forceinline__ float4 pml_pre4(float4 a, float4 fo, float4 fl, float4 e) {
    // h = a*e - fo*flux   (the field itself becomes the old flux)
    float4 h;
    h.x = __fmaf_rn(a.x, e.x, -__fmul_rn(fo.x, fl.x));
    h.y = __fmaf_rn(a.y, e.y, -__fmul_rn(fo.y, fl.y));
    h.z = __fmaf_rn(a.z, e.z, -__fmul_rn(fo.z, fl.z));
    h.w = __fmaf_rn(a.w, e.w, -__fmul_rn(fo.w, fl.w));
    return h;
}
__device__ __forceinline__ float4 pml_post4(float4 fn, float4 F, float4 h) {
    return make_float4(__fmaf_rn(fn.x, F.x, h.x), __fmaf_rn(fn.y, F.y, h.y), __fmaf_rn(fn.z, F.z, h.z), __fmaf_rn(fn.w, F.w, h.w));
}

// Slab coefficient rows are compressible exactly like the operator rows (a, fo, fn are products of 1-D PML profiles):
// 48-byte record per slab row = 9 scales (a_xyz, fo_xyz, fn_xyz) + 9 vector ids; the x-vectors have the slab's row width.
struct PmlRowMeta { float sc[9]; unsigned char id[9]; unsigned char pad[3]; };

__device__ __forceinline__ float4 pcoef4(unsigned id, float sc, const float* full, const float* __restrict__ xv, int col, int w)
{
    if (id != ROW_FULL) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xv + (size_t)id * w + col));
        return make_float4(__fmul_rn(sc, v.x), __fmul_rn(sc, v.y), __fmul_rn(sc, v.z), __fmul_rn(sc, v.w));
    }
    return __ldg(reinterpret_cast<const float4*>(full));
}

// the three PML coefficients of component c4 (0..2) of the current slab row; PM_ = row record or NULL
#define PML_COEFFS(c4, PA, PFO, PFN, PXV, COL, W, lofs)                                                         \
    float4 a_, fo_, fn_;                                                                                        \
    if (pm_ != nullptr && pfull_ == 0) {          /* every slot of the slab row is compressed */                 \
        const char* pl_ = reinterpret_cast<const char*>((PXV) + (COL)); const unsigned pp_ = 4u * (unsigned)(W); \
        a_ = xvg4(pl_, pid_[c4], pp_, psc_[c4]);                                                                \
        fo_ = xvg4(pl_, pid_[3 + c4], pp_, psc_[3 + c4]);                                                       \
        fn_ = xvg4(pl_, pid_[6 + c4], pp_, psc_[6 + c4]);                                                       \
    } else if (pm_ != nullptr) {                                                                                \
        a_ = pcoef4(pid_[c4], psc_[c4], (PA) + (lofs), PXV, COL, W);                                            \
        fo_ = pcoef4(pid_[3 + c4], psc_[3 + c4], (PFO) + (lofs), PXV, COL, W);                                  \
        fn_ = pcoef4(pid_[6 + c4], psc_[6 + c4], (PFN) + (lofs), PXV, COL, W);                                  \
    } else { a_ = ld4_nc((PA) + (lofs)); fo_ = ld4_nc((PFO) + (lofs)); fn_ = ld4_nc((PFN) + (lofs)); }

// load the row record into registers (row-uniform in MODE 1, per lane in MODE 2)
#define PML_ROW_META(PMETA, ROW)                                                                                \
    const unsigned char* pm_ = (PMETA) ? (PMETA) + (long long)(ROW) * 48 : nullptr;                             \
    float psc_[9]; unsigned pid_[9]; unsigned pfull_ = 1;                                                       \
    if (pm_ != nullptr) {                                                                                       \
        if (k + KSTEP >= kbeg && k + KSTEP < kend) prefetch_l1(pm_ + (long long)KSTEP * r.by * 48);             \
        const float4 q0_ = __ldg(reinterpret_cast<const float4*>(pm_)), q1_ = __ldg(reinterpret_cast<const float4*>(pm_) + 1), \
                     q2_ = __ldg(reinterpret_cast<const float4*>(pm_) + 2);                                     \
        psc_[0] = q0_.x; psc_[1] = q0_.y; psc_[2] = q0_.z; psc_[3] = q0_.w;                                     \
        psc_[4] = q1_.x; psc_[5] = q1_.y; psc_[6] = q1_.z; psc_[7] = q1_.w; psc_[8] = q2_.x;                    \
        const unsigned w0_ = __float_as_uint(q2_.y), w1_ = __float_as_uint(q2_.z), w2_ = __float_as_uint(q2_.w); \
        pid_[0] = w0_ & 255u; pid_[1] = (w0_ >> 8) & 255u; pid_[2] = (w0_ >> 16) & 255u; pid_[3] = w0_ >> 24;   \
        pid_[4] = w1_ & 255u; pid_[5] = (w1_ >> 8) & 255u; pid_[6] = (w1_ >> 16) & 255u; pid_[7] = w1_ >> 24;   \
        pid_[8] = w2_ & 255u; pfull_ = (w2_ >> 8) & 255u;      /* pad[0]: a slot of this slab row is streamed in full */ \
    }

// one component of a fused PML row: pre, update, post.  f4 holds the field on entry and the new field on exit; fl4 holds
// the old flux on entry (loaded by the caller together with every other load of the plane, so one plane costs one
// round of DRAM latency, not one per component) and the new flux on exit (stored by the caller after all three).
#define PML_COMP(c4, f4, fl4, ca4, cb4, A, B, C, D, lofs)                                     \
    do {                                                                                      \
        float4 h_ = zero4(), fnn_ = zero4();                                                  \
        if (act) {                                                                            \
            PML_COEFFS(c4, r.a, r.fo, r.fn, r.pxv, i0, p.px, lofs)                            \
            h_ = pml_pre4(a_, fo_, fl4, f4); fnn_ = fn_;                                      \
        }                                                                                     \
        fl4 = upd4(ca4, fl4, cb4, A, B, C, D);                                                \
        f4 = pml_post4(fnn_, fl4, h_);                                                        \
    } while (0)

// the same with explicit slab pointers (narrow x-slab launches: every lane is inside its slab)
#define PML_COMP_X(c4, f4, fl4, ca4, cb4, A, B, C, D, lofs)                                   \
    do {                                                                                      \
        PML_COEFFS(c4, xa, xfo, xfn, xpxv, i0 - xx0, xw, lofs)                                \
        const float4 h_ = pml_pre4(a_, fo_, fl4, f4);                                         \
        fl4 = upd4(ca4, fl4, cb4, A, B, C, D);                                                \
        f4 = pml_post4(fn_, fl4, h_);                                                         \
    } while (0)

// E update: volt_n = vv_n volt_n + vi_n curl_n(curr)   (App. A1)
//   x: ((Hz - Hz[j-1]) - Hy) + Hy[k-1]
//   y: ((Hx - Hx[k-1]) - Hz) + Hz[i-1]
//   z: ((Hy - Hy[i-1]) - Hx) + Hx[j-1]
template <int TY, int MODE, bool CMP>      // MODE 0 plain rows, 1 fused PML rows, 2 plain rows whose x-edge lanes are PML
__global__ void __launch_bounds__(32 * TY, MODE ? (16 / TY > 0 ? 16 / TY : 1) : (24 / TY > 0 ? 24 / TY : 1)) update_e_kernel(const VolParams p, const RowParams r)
{
    constexpr bool PML = MODE == 1;
    constexpr int KSTEP = 1;                                // the E march goes up in z
    const int lane = threadIdx.x;
    int i0, j;
    bool act;
    // narrow-slab launch state (MODE 2)
    const bool hi_slab = MODE == 2 && (blockIdx.x == 1 || r.xw0 == 0);
    const int xw = hi_slab ? r.xw1 : r.xw0, xx0 = hi_slab ? r.xx1 : 0;
    if (MODE == 2) {
        const int xs = hi_slab ? r.xs1 : r.xs0;             // float4 slots per row, 32/xs rows per warp (12 columns: 10 rows, 2 idle lanes)
        const int c4 = lane % xs, rw = lane / xs;
        j = r.j0 + (blockIdx.y * TY + threadIdx.y) * (32 / xs) + rw;
        i0 = xx0 + c4 * 4;
        if (rw >= 32 / xs || j >= r.j1 || c4 * 4 >= xw) return;   // per lane (no warp collectives in this mode)
        act = true;
    } else {
        i0 = (blockIdx.x * 32 + lane) * 4;
        j = r.j0 + blockIdx.y * TY + threadIdx.y;
        if (j >= r.j1) return;                              // warp-uniform
        if (MODE == 0) {
            if ((j >= r.sj0a && j < r.sj1a) || (j >= r.sj0b && j < r.sj1b)) return;
        }
        act = i0 < p.px;
    }
    // plain launch: columns owned by a narrow-slab launch are computed but not stored
    const bool own = MODE != 0 || !(i0 < r.xw0 || (i0 >= r.xx1 && i0 < r.xx1 + r.xw1));
    const int kbeg = p.k0 + blockIdx.z * p.kz;
    const int kend = min(kbeg + p.kz, p.k1);
    const long long cs = p.cs, sz = p.sz;
    const float* __restrict__ g = p.g;
    float* __restrict__ f = p.f;
    const float* fin = p.fin;
    const char* xv_lane_ = reinterpret_cast<const char*>(p.xv + i0); const unsigned xv_pitch_ = 4u * (unsigned)p.px;
    (void)xv_lane_; (void)xv_pitch_;

    long long base = (long long)kbeg * sz + (long long)j * p.px + i0;   // plane kbeg-1 (ghost offset +1 applied below)
    float4 hx_km = zero4(), hy_km = zero4();
    if (act) { hx_km = ld4(g + base); hy_km = ld4(g + cs + base); }
    base += sz;                                             // plane kbeg
    const bool has_jm = j > 0;
    const bool edge_load = act && lane == 0 && i0 > 0;
    long long lb = 0, lsz = 0, lcs = 0;
    if (PML) {
        lsz = (long long)r.by * p.px; lcs = lsz * r.bz;
        lb = (long long)(kbeg - r.z0) * lsz + (long long)(j - r.y0) * p.px + i0;
    }
    float* xflux = nullptr; const float* xa = nullptr; const float* xfo = nullptr; const float* xfn = nullptr;
    const float* xpxv = nullptr; const unsigned char* xpmeta = nullptr;
    if (MODE == 2) {
        xflux = hi_slab ? r.xflux1 : r.xflux0; xa = hi_slab ? r.xa1 : r.xa0; xfo = hi_slab ? r.xfo1 : r.xfo0; xfn = hi_slab ? r.xfn1 : r.xfn0;
        xpxv = hi_slab ? r.pxv1 : r.pxv0; xpmeta = hi_slab ? r.pmeta1 : r.pmeta0;
        lsz = (long long)r.by * xw; lcs = lsz * r.bz;
        lb = (long long)(kbeg - r.z0) * lsz + (long long)(j - r.y0) * xw + (i0 - xx0);
    }

    for (int k = kbeg; k < kend; ++k, base += sz, lb += lsz) {
        float4 hx = zero4(), hy = zero4(), hz = zero4(), hz_jm = zero4(), hx_jm = zero4();
        float4 ex = zero4(), ey = zero4(), ez = zero4();
        float4 ax = zero4(), ay = zero4(), az = zero4(), bx = zero4(), by = zero4(), bz = zero4();
        float hz_e = 0.f, hy_e = 0.f;
        // every load of the plane is issued up front (slab row record and old flux included): one DRAM round trip per plane
        PML_ROW_META(MODE == 1 ? r.pmeta : (MODE == 2 ? xpmeta : nullptr), (long long)(k - r.z0) * r.by + (j - r.y0))
        float4 fl0 = zero4(), fl1 = zero4(), fl2 = zero4();
        if (MODE == 1 && act) { fl0 = ld4_stream(r.flux + lb); fl1 = ld4_stream(r.flux + lcs + lb); fl2 = ld4_stream(r.flux + 2 * lcs + lb); }
        if (MODE == 2) { fl0 = ld4_stream(xflux + lb); fl1 = ld4_stream(xflux + lcs + lb); fl2 = ld4_stream(xflux + 2 * lcs + lb); }
        if (MODE != 0 && act && k + 1 < kend) {             // slab launches are latency-bound: pull the next plane into L2
            prefetch_l2(g + base + sz); prefetch_l2(g + cs + base + sz); prefetch_l2(g + 2 * cs + base + sz);
            prefetch_l2(fin + base + sz); prefetch_l2(fin + cs + base + sz); prefetch_l2(fin + 2 * cs + base + sz);
            const float* fx = MODE == 1 ? r.flux : xflux;
            prefetch_l2(fx + lb + lsz); prefetch_l2(fx + lcs + lb + lsz); prefetch_l2(fx + 2 * lcs + lb + lsz);
        }
        if (act) {
            hx = ld4(g + base); hy = ld4(g + cs + base); hz = ld4(g + 2 * cs + base);
            if (has_jm) { hz_jm = ld4(g + 2 * cs + base - p.px); hx_jm = ld4(g + base - p.px); }
            ex = ld4_stream(fin + base); ey = ld4_stream(fin + cs + base); ez = ld4_stream(fin + 2 * cs + base);
            if (CMP) LOAD_COEFFS_CMP();
            else {
                ax = ld4_ro(p.ca + base); ay = ld4_ro(p.ca + cs + base); az = ld4_ro(p.ca + 2 * cs + base);
                bx = ld4_ro(p.cb + base); by = ld4_ro(p.cb + cs + base); bz = ld4_ro(p.cb + 2 * cs + base);
            }
        }
        float hz_l, hy_l;
        if (MODE == 2) {
            hz_l = i0 > 0 ? g[2 * cs + base - 1] : 0.f; hy_l = i0 > 0 ? g[cs + base - 1] : 0.f;
        } else {
            if (edge_load) { hz_e = g[2 * cs + base - 1]; hy_e = g[cs + base - 1]; }
            hz_l = __shfl_up_sync(0xffffffffu, hz.w, 1);
            hy_l = __shfl_up_sync(0xffffffffu, hy.w, 1);
            if (lane == 0) { hz_l = hz_e; hy_l = hy_e; }
        }
        const float4 hz_im = make_float4(hz_l, hz.x, hz.y, hz.z);
        const float4 hy_im = make_float4(hy_l, hy.x, hy.y, hy.z);

        if (PML) {
            PML_COMP(0, ex, fl0, ax, bx, hz, hz_jm, hy, hy_km, lb);
            PML_COMP(1, ey, fl1, ay, by, hx, hx_km, hz, hz_im, lcs + lb);
            PML_COMP(2, ez, fl2, az, bz, hy, hy_im, hx, hx_jm, 2 * lcs + lb);
            if (act) { st4(r.flux + lb, fl0); st4(r.flux + lcs + lb, fl1); st4(r.flux + 2 * lcs + lb, fl2); }
        } else if (MODE == 2) {
            PML_COMP_X(0, ex, fl0, ax, bx, hz, hz_jm, hy, hy_km, lb);
            PML_COMP_X(1, ey, fl1, ay, by, hx, hx_km, hz, hz_im, lcs + lb);
            PML_COMP_X(2, ez, fl2, az, bz, hy, hy_im, hx, hx_jm, 2 * lcs + lb);
            st4(xflux + lb, fl0); st4(xflux + lcs + lb, fl1); st4(xflux + 2 * lcs + lb, fl2);
        } else {
            ex = upd4(ax, ex, bx, hz, hz_jm, hy, hy_km);
            ey = upd4(ay, ey, by, hx, hx_km, hz, hz_im);
            ez = upd4(az, ez, bz, hy, hy_im, hx, hx_jm);
        }
        if (act && own) {
            st4(f + base, ex); st4(f + cs + base, ey); st4(f + 2 * cs + base, ez);
        }
        hx_km = hx; hy_km = hy;
    }
}

// H update: curr_n = ii_n curr_n + iv_n curl_n(volt)   (App. A1), marching downwards in z
//   x: ((Ez - Ez[j+1]) - Ey) + Ey[k+1]
//   y: ((Ex - Ex[k+1]) - Ez) + Ez[i+1]
//   z: ((Ey - Ey[i+1]) - Ex) + Ex[j+1]
template <int TY, int MODE, bool CMP>
__global__ void __launch_bounds__(32 * TY, MODE ? (16 / TY > 0 ? 16 / TY : 1) : (24 / TY > 0 ? 24 / TY : 1)) update_h_kernel(const VolParams p, const RowParams r)
{
    constexpr bool PML = MODE == 1;
    constexpr int KSTEP = -1;                               // the H march goes down in z
    const int lane = threadIdx.x;
    int i0, j;
    bool act;
    // narrow-slab launch state (MODE 2)
    const bool hi_slab = MODE == 2 && (blockIdx.x == 1 || r.xw0 == 0);
    const int xw = hi_slab ? r.xw1 : r.xw0, xx0 = hi_slab ? r.xx1 : 0;
    if (MODE == 2) {
        const int xs = hi_slab ? r.xs1 : r.xs0;             // float4 slots per row, 32/xs rows per warp (12 columns: 10 rows, 2 idle lanes)
        const int c4 = lane % xs, rw = lane / xs;
        j = r.j0 + (blockIdx.y * TY + threadIdx.y) * (32 / xs) + rw;
        i0 = xx0 + c4 * 4;
        if (rw >= 32 / xs || j >= r.j1 || c4 * 4 >= xw) return;   // per lane (no warp collectives in this mode)
        act = true;
    } else {
        i0 = (blockIdx.x * 32 + lane) * 4;
        j = r.j0 + blockIdx.y * TY + threadIdx.y;
        if (j >= r.j1) return;                              // warp-uniform
        if (MODE == 0) {
            if ((j >= r.sj0a && j < r.sj1a) || (j >= r.sj0b && j < r.sj1b)) return;
        }
        act = i0 < p.px;
    }
    // plain launch: columns owned by a narrow-slab launch are computed but not stored
    const bool own = MODE != 0 || !(i0 < r.xw0 || (i0 >= r.xx1 && i0 < r.xx1 + r.xw1));
    const int kbeg = p.k0 + blockIdx.z * p.kz;
    const int kend = min(kbeg + p.kz, p.k1);
    const long long cs = p.cs, sz = p.sz;
    const float* __restrict__ g = p.g;
    float* __restrict__ f = p.f;
    const float* fin = p.fin;
    const char* xv_lane_ = reinterpret_cast<const char*>(p.xv + i0); const unsigned xv_pitch_ = 4u * (unsigned)p.px;
    (void)xv_lane_; (void)xv_pitch_;

    long long base = (long long)(kend + 1) * sz + (long long)j * p.px + i0;   // plane kend (k+1 of the first plane)
    float4 ex_kp = zero4(), ey_kp = zero4();
    if (act) { ex_kp = ld4(g + base); ey_kp = ld4(g + cs + base); }
    base -= sz;
    const bool has_jp = j + 1 < p.ny;
    const bool last = act && (lane == 31 || i0 + 4 >= p.px);
    const bool edge_load = last && (i0 + 4 < p.px);
    long long lb = 0, lsz = 0, lcs = 0;
    if (PML) {
        lsz = (long long)r.by * p.px; lcs = lsz * r.bz;
        lb = (long long)(kend - 1 - r.z0) * lsz + (long long)(j - r.y0) * p.px + i0;
    }
    float* xflux = nullptr; const float* xa = nullptr; const float* xfo = nullptr; const float* xfn = nullptr;
    const float* xpxv = nullptr; const unsigned char* xpmeta = nullptr;
    if (MODE == 2) {
        xflux = hi_slab ? r.xflux1 : r.xflux0; xa = hi_slab ? r.xa1 : r.xa0; xfo = hi_slab ? r.xfo1 : r.xfo0; xfn = hi_slab ? r.xfn1 : r.xfn0;
        xpxv = hi_slab ? r.pxv1 : r.pxv0; xpmeta = hi_slab ? r.pmeta1 : r.pmeta0;
        lsz = (long long)r.by * xw; lcs = lsz * r.bz;
        lb = (long long)(kend - 1 - r.z0) * lsz + (long long)(j - r.y0) * xw + (i0 - xx0);
    }

    for (int k = kend - 1; k >= kbeg; --k, base -= sz, lb -= lsz) {
        float4 ex = zero4(), ey = zero4(), ez = zero4(), ez_jp = zero4(), ex_jp = zero4();
        float4 hx = zero4(), hy = zero4(), hz = zero4();
        float4 ax = zero4(), ay = zero4(), az = zero4(), bx = zero4(), by = zero4(), bz = zero4();
        float ez_e = 0.f, ey_e = 0.f;
        PML_ROW_META(MODE == 1 ? r.pmeta : (MODE == 2 ? xpmeta : nullptr), (long long)(k - r.z0) * r.by + (j - r.y0))
        float4 fl0 = zero4(), fl1 = zero4(), fl2 = zero4();
        if (MODE == 1 && act) { fl0 = ld4_stream(r.flux + lb); fl1 = ld4_stream(r.flux + lcs + lb); fl2 = ld4_stream(r.flux + 2 * lcs + lb); }
        if (MODE == 2) { fl0 = ld4_stream(xflux + lb); fl1 = ld4_stream(xflux + lcs + lb); fl2 = ld4_stream(xflux + 2 * lcs + lb); }
        if (MODE != 0 && act && k - 1 >= kbeg) {
            prefetch_l2(g + base - sz); prefetch_l2(g + cs + base - sz); prefetch_l2(g + 2 * cs + base - sz);
            prefetch_l2(fin + base - sz); prefetch_l2(fin + cs + base - sz); prefetch_l2(fin + 2 * cs + base - sz);
            const float* fx = MODE == 1 ? r.flux : xflux;
            prefetch_l2(fx + lb - lsz); prefetch_l2(fx + lcs + lb - lsz); prefetch_l2(fx + 2 * lcs + lb - lsz);
        }
        if (act) {
            ex = ld4(g + base); ey = ld4(g + cs + base); ez = ld4(g + 2 * cs + base);
            if (has_jp) { ez_jp = ld4(g + 2 * cs + base + p.px); ex_jp = ld4(g + base + p.px); }
            hx = ld4_stream(fin + base); hy = ld4_stream(fin + cs + base); hz = ld4_stream(fin + 2 * cs + base);
            if (CMP) LOAD_COEFFS_CMP();
            else {
                ax = ld4_ro(p.ca + base); ay = ld4_ro(p.ca + cs + base); az = ld4_ro(p.ca + 2 * cs + base);
                bx = ld4_ro(p.cb + base); by = ld4_ro(p.cb + cs + base); bz = ld4_ro(p.cb + 2 * cs + base);
            }
        }
        float ez_r, ey_r;
        if (MODE == 2) {
            ez_r = i0 + 4 < p.px ? g[2 * cs + base + 4] : 0.f; ey_r = i0 + 4 < p.px ? g[cs + base + 4] : 0.f;
        } else {
            if (edge_load) { ez_e = g[2 * cs + base + 4]; ey_e = g[cs + base + 4]; }
            ez_r = __shfl_down_sync(0xffffffffu, ez.x, 1);
            ey_r = __shfl_down_sync(0xffffffffu, ey.x, 1);
            if (last || !act) { ez_r = ez_e; ey_r = ey_e; }
        }
        const float4 ez_ip = make_float4(ez.y, ez.z, ez.w, ez_r);
        const float4 ey_ip = make_float4(ey.y, ey.z, ey.w, ey_r);

        if (PML) {
            PML_COMP(0, hx, fl0, ax, bx, ez, ez_jp, ey, ey_kp, lb);
            PML_COMP(1, hy, fl1, ay, by, ex, ex_kp, ez, ez_ip, lcs + lb);
            PML_COMP(2, hz, fl2, az, bz, ey, ey_ip, ex, ex_jp, 2 * lcs + lb);
            if (act) { st4(r.flux + lb, fl0); st4(r.flux + lcs + lb, fl1); st4(r.flux + 2 * lcs + lb, fl2); }
        } else if (MODE == 2) {
            PML_COMP_X(0, hx, fl0, ax, bx, ez, ez_jp, ey, ey_kp, lb);
            PML_COMP_X(1, hy, fl1, ay, by, ex, ex_kp, ez, ez_ip, lcs + lb);
            PML_COMP_X(2, hz, fl2, az, bz, ey, ey_ip, ex, ex_jp, 2 * lcs + lb);
            st4(xflux + lb, fl0); st4(xflux + lcs + lb, fl1); st4(xflux + 2 * lcs + lb, fl2);
        } else {
            hx = upd4(ax, hx, bx, ez, ez_jp, ey, ey_kp);
            hy = upd4(ay, hy, by, ex, ex_kp, ez, ez_ip);
            hz = upd4(az, hz, bz, ey, ey_ip, ex, ex_jp);
        }
        if (act && own) {
            st4(f + base, hx); st4(f + cs + base, hy); st4(f + 2 * cs + base, hz);
        }
        ex_kp = ex; ey_kp = ey;
    }
}


// ------------------------------------------------------------------------------------
// fused H->E launch (temporal blocking over the pair  H update of step n, E update of step n+1)
// ------------------------------------------------------------------------------------
// Between the H update of one step and the E update of the next nothing else touches the fields, so both can be done
// in one sweep: 48 B/cell of field traffic (E, H read + written once) instead of 72 B (each pass re-reads the other
// field).  A CTA owns TY rows x 124 columns and marches up in z.  E_new(i,j,k) needs H_new at (i-1), (j-1), (k-1): the
// CTA recomputes H_new on a one-cell halo at its low sides (row 0 of the CTA, lane 0 of every warp, one extra plane below
// the chunk) from the OLD fields, which is why the launch writes a second copy of the fields instead of updating in place.
// Halo cells that lie outside the launch region (PML slabs, whose H update ran just before this launch into the same
// output copy) are read from the output copy instead.  Same arithmetic per cell as update_h_kernel / update_e_kernel.
struct HeParams {
    const float* __restrict__ ein; const float* __restrict__ hin;
    float* __restrict__ eout; float* hout;          // hout is also read (halo cells outside the region)
    const float* __restrict__ vv; const float* __restrict__ vi; const float* __restrict__ ii; const float* __restrict__ iv;
    const float* __restrict__ xv_e; const unsigned char* __restrict__ meta_e;
    const float* __restrict__ xv_h; const unsigned char* __restrict__ meta_h;
    int ny, px; long long sz, cs;
    int X0, X1, XT0;                // owned columns [X0,X1) and [XT0,px), multiples of 4 (the gap is a narrow PML x-slab)
    int Y0, Y1, Z0, Z1;             // owned rows and planes
    int kz;
    int pf;                         // planes of L2 prefetch distance (0 = off)
    // byte offsets as launch constants (update_he2_kernel adds them to per-thread plane pointers: two integer
    // instructions per address instead of a 64-bit index computation)
    long long b_sz, b_cs, b_2cs, b_sz_cs, b_sz_2cs, b_row, b_row_2cs, b_pfe[3], b_pfh[3];
    unsigned xv_pitch; int meta_step;
    int nv_e, nv_h;                 // x-vectors of the E / H pass (update_he2_kernel keeps its 128-column slice of them in smem)
};

template <bool CMP>
__device__ __forceinline__ void load_coefs6(const float* __restrict__ ca, const float* __restrict__ cb, const float* __restrict__ xv,
        const unsigned char* __restrict__ meta, long long base, long long cs, long long row, int ny, int i0, int px,
        float4& ax, float4& ay, float4& az, float4& bx, float4& by, float4& bz)
{
    if (CMP) {
        const float4* m_ = reinterpret_cast<const float4*>(meta + row * 32);
        prefetch_l1(meta + (row + ny) * 32);                 // the march goes up: next plane's record
        const float4 m0_ = __ldg(m_), m1_ = __ldg(m_ + 1);
        const unsigned w0_ = __float_as_uint(m1_.z), w1_ = __float_as_uint(m1_.w);
        ax = coef4(w0_ & 255u, m0_.x, ca + base, xv, i0, px);
        ay = coef4((w0_ >> 8) & 255u, m0_.y, ca + cs + base, xv, i0, px);
        az = coef4((w0_ >> 16) & 255u, m0_.z, ca + 2 * cs + base, xv, i0, px);
        bx = coef4(w0_ >> 24, m0_.w, cb + base, xv, i0, px);
        by = coef4(w1_ & 255u, m1_.x, cb + cs + base, xv, i0, px);
        bz = coef4((w1_ >> 8) & 255u, m1_.y, cb + 2 * cs + base, xv, i0, px);
    } else {
        ax = ld4_ro(ca + base); ay = ld4_ro(ca + cs + base); az = ld4_ro(ca + 2 * cs + base);
        bx = ld4_ro(cb + base); by = ld4_ro(cb + cs + base); bz = ld4_ro(cb + 2 * cs + base);
    }
}

#define HE_SEG 124          // columns owned by a warp: lanes 1..31; lane 0 is the x-halo

template <int TY, bool CMP>
__global__ void __launch_bounds__(32 * (TY + 1), 16 / (TY + 1)) update_he_kernel(const HeParams p)
{
    __shared__ float4 xb[2][TY + 1][2][32];                  // H_new (hz, hx) of every row, for the row above; double buffered
    const int lane = threadIdx.x, r = threadIdx.y;
    const int i0 = p.X0 - 4 + HE_SEG * (int)blockIdx.x + 4 * lane;
    const int j = p.Y0 - 1 + TY * (int)blockIdx.y + r;
    const int kbeg = p.Z0 + (int)blockIdx.z * p.kz;
    const int kend = min(kbeg + p.kz, p.Z1);
    const bool in_grid = i0 >= 0 && i0 < p.px && j >= 0 && j < p.Y1;        // rows >= Y1 are needed by nobody here
    const bool reg_x = (i0 >= p.X0 && i0 < p.X1) || i0 >= p.XT0;
    const bool ext = in_grid && (!reg_x || j < p.Y0);       // H_new was written by a slab launch: read it
    const bool calc = in_grid && !ext;                       // H_new is computed here (owned cells and halo cells)
    const bool own = calc && lane >= 1 && r >= 1;            // ... and stored, together with E_new
    const bool has_jp = j + 1 < p.ny;
    const bool edge_load = in_grid && lane == 31 && i0 + 4 < p.px;
    const long long cs = p.cs, sz = p.sz;
    const float* __restrict__ ein = p.ein; const float* __restrict__ hin = p.hin;
    float* __restrict__ eout = p.eout; float* hout = p.hout;

    const int kfirst = kbeg > p.Z0 ? kbeg - 1 : kbeg;        // one plane below the chunk: H_new(kbeg-1) is recomputed
    long long base = (long long)(kfirst + 1) * sz + (long long)j * p.px + i0;   // plane kfirst (ghost offset +1)
    float4 ex = zero4(), ey = zero4(), ez = zero4();         // E_old(k)
    float4 hx_km = zero4(), hy_km = zero4();                 // H_new(k-1)
    if (in_grid) { ex = ld4(ein + base); ey = ld4(ein + cs + base); ez = ld4(ein + 2 * cs + base); }
    if (kfirst == kbeg && own) { hx_km = ld4(hout + base - sz); hy_km = ld4(hout + cs + base - sz); }

    for (int k = kfirst; k < kend; ++k, base += sz) {
        const bool pro = k < kbeg;                           // prologue plane: H_new only, nothing stored
        float4 ex1 = zero4(), ey1 = zero4(), ez1 = zero4(), ez_jp = zero4(), ex_jp = zero4();
        float4 hx = zero4(), hy = zero4(), hz = zero4();
        float4 ax = zero4(), ay = zero4(), az = zero4(), bx = zero4(), by = zero4(), bz = zero4();
        float ez_e = 0.f, ey_e = 0.f;
        if (p.pf > 0 && in_grid && k + p.pf < kend) {
            // the march is a chain of DRAM round trips with few warps per SM: pull the planes of a later iteration into L2
            const long long d = (long long)p.pf * sz;
            prefetch_l2(ein + base + sz + d); prefetch_l2(ein + cs + base + sz + d); prefetch_l2(ein + 2 * cs + base + sz + d);
            if (calc) { prefetch_l2(hin + base + d); prefetch_l2(hin + cs + base + d); prefetch_l2(hin + 2 * cs + base + d); }
        }
        if (in_grid) {
            ex1 = ld4(ein + base + sz); ey1 = ld4(ein + cs + base + sz); ez1 = ld4(ein + 2 * cs + base + sz);
            if (has_jp) { ez_jp = ld4(ein + 2 * cs + base + p.px); ex_jp = ld4(ein + base + p.px); }
        }
        if (calc) {
            hx = ld4_stream(hin + base); hy = ld4_stream(hin + cs + base); hz = ld4_stream(hin + 2 * cs + base);
            load_coefs6<CMP>(p.ii, p.iv, p.xv_h, p.meta_h, base, cs, (long long)(k + 1) * p.ny + j, p.ny, i0, p.px, ax, ay, az, bx, by, bz);
        } else if (ext) {
            hx = ld4(hout + base); hy = ld4(hout + cs + base); hz = ld4(hout + 2 * cs + base);
        }
        if (edge_load) { ez_e = ein[2 * cs + base + 4]; ey_e = ein[cs + base + 4]; }
        float ez_r = __shfl_down_sync(0xffffffffu, ez.x, 1);
        float ey_r = __shfl_down_sync(0xffffffffu, ey.x, 1);
        if (lane == 31) { ez_r = ez_e; ey_r = ey_e; }
        if (calc) {
            const float4 ez_ip = make_float4(ez.y, ez.z, ez.w, ez_r);
            const float4 ey_ip = make_float4(ey.y, ey.z, ey.w, ey_r);
            hx = upd4(ax, hx, bx, ez, ez_jp, ey, ey1);
            hy = upd4(ay, hy, by, ex, ex1, ez, ez_ip);
            hz = upd4(az, hz, bz, ey, ey_ip, ex, ex_jp);
        }
        // hx, hy, hz now hold H_new(k) (zero outside the grid)
        if (!pro) {
            if (own) { st4(hout + base, hx); st4(hout + cs + base, hy); st4(hout + 2 * cs + base, hz); }
            xb[k & 1][r][0][lane] = hz; xb[k & 1][r][1][lane] = hx;
            __syncthreads();
            const float hz_l = __shfl_up_sync(0xffffffffu, hz.w, 1);
            const float hy_l = __shfl_up_sync(0xffffffffu, hy.w, 1);
            if (own) {
                const float4 hz_jm = xb[k & 1][r - 1][0][lane], hx_jm = xb[k & 1][r - 1][1][lane];
                const float4 hz_im = make_float4(hz_l, hz.x, hz.y, hz.z);
                const float4 hy_im = make_float4(hy_l, hy.x, hy.y, hy.z);
                load_coefs6<CMP>(p.vv, p.vi, p.xv_e, p.meta_e, base, cs, (long long)(k + 1) * p.ny + j, p.ny, i0, p.px, ax, ay, az, bx, by, bz);
                ex = upd4(ax, ex, bx, hz, hz_jm, hy, hy_km);
                ey = upd4(ay, ey, by, hx, hx_km, hz, hz_im);
                ez = upd4(az, ez, bz, hy, hy_im, hx, hx_jm);
                st4(eout + base, ex); st4(eout + cs + base, ey); st4(eout + 2 * cs + base, ez);
            }
        }
        hx_km = hx; hy_km = hy;
        ex = ex1; ey = ey1; ez = ez1;
    }
}


// ---- the same sweep with the planes staged through shared memory by cp.async (LDGSTS) ----
// The register version above is bound by DRAM latency: 16 warps per SM, each waiting on the loads of its own plane.
// Here every thread copies the float4s it will need one plane ahead straight into shared memory (no registers held while
// the copy is in flight), so a CTA always has a whole plane of loads outstanding while it computes the previous one, and
// the y-neighbour rows come from shared memory instead of a second global load.
//   E ring: HE_DIST+2 planes (k and k+1 in use, the rest landing)   [3 comps][TY+2 rows][33 float4]   (row TY+1 / column 32 = +1 halo)
//   H ring: HE_DIST+1 planes (k in use, the rest landing)           [3 comps][TY+1 rows][32 float4]
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    const int n = valid ? 16 : 0;                            // 0 source bytes: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

#define HE_DIST 2           // planes in flight ahead of the one being computed
template <int TY>
struct HeSmem {
    float4 e[HE_DIST + 2][3][TY + 2][33];
    float4 h[HE_DIST + 1][3][TY + 1][32];
    float4 xb[2][TY + 1][2][32];
};

template <int TY, bool CMP>
__global__ void __launch_bounds__(32 * (TY + 1), 16 / (TY + 1)) update_he_staged_kernel(const HeParams p)
{
    extern __shared__ __align__(16) unsigned char he_smem_raw[];
    HeSmem<TY>& S = *reinterpret_cast<HeSmem<TY>*>(he_smem_raw);
    const int lane = threadIdx.x, r = threadIdx.y;
    const int i0 = p.X0 - 4 + HE_SEG * (int)blockIdx.x + 4 * lane;
    const int j = p.Y0 - 1 + TY * (int)blockIdx.y + r;
    const int kbeg = p.Z0 + (int)blockIdx.z * p.kz;
    const int kend = min(kbeg + p.kz, p.Z1);
    const bool col_ok = i0 >= 0 && i0 < p.px;
    const bool in_grid = col_ok && j >= 0 && j < p.Y1;      // rows >= Y1 are needed only as the +1 neighbour row
    const bool reg_x = (i0 >= p.X0 && i0 < p.X1) || i0 >= p.XT0;
    const bool ext = in_grid && (!reg_x || j < p.Y0);       // H_new was written by a slab launch: read it
    const bool calc = in_grid && !ext;
    const bool own = calc && lane >= 1 && r >= 1;
    // what this thread stages: its own float4 of E (any row of the grid up to Y1, which is the +1 row of the last owned
    // row), the +1 row for the top warp, the +1 column for lane 31, and its own float4 of H_old where H_new is computed
    const bool e_ok = col_ok && j >= 0 && j < p.ny && j <= p.Y1;
    const bool top = r == TY;
    const bool e_top_ok = top && col_ok && j + 1 < p.ny && j + 1 <= p.Y1;
    const bool e_col_ok = lane == 31 && i0 + 4 < p.px && j >= 0 && j < p.Y1;
    const long long cs = p.cs, sz = p.sz;
    const float* __restrict__ ein = p.ein; const float* __restrict__ hin = p.hin;
    float* __restrict__ eout = p.eout; float* hout = p.hout;
    const long long rowoff = (long long)j * p.px + i0;       // may be "negative" for halo threads: only used when valid

    auto stage_e = [&](int k) {                              // E_old(k) -> ring slot k % (HE_DIST+2)
        const int s = k % (HE_DIST + 2);
        const long long b = (long long)(k + 1) * sz + rowoff;
        const float* src = e_ok ? ein + b : ein;
        cp_async16(&S.e[s][0][r][lane], src, e_ok);
        cp_async16(&S.e[s][1][r][lane], src + (e_ok ? cs : 0), e_ok);
        cp_async16(&S.e[s][2][r][lane], src + (e_ok ? 2 * cs : 0), e_ok);
        if (top) {                                           // row TY+1: (ex, ez) of row j+1
            const float* st = e_top_ok ? ein + b + p.px : ein;
            cp_async16(&S.e[s][0][TY + 1][lane], st, e_top_ok);
            cp_async16(&S.e[s][2][TY + 1][lane], st + (e_top_ok ? 2 * cs : 0), e_top_ok);
        }
        if (lane == 31) {                                    // column 32: (ey, ez) of the float4 right of the segment
            const float* sc = e_col_ok ? ein + b + 4 : ein;
            cp_async16(&S.e[s][1][r][32], sc + (e_col_ok ? cs : 0), e_col_ok);
            cp_async16(&S.e[s][2][r][32], sc + (e_col_ok ? 2 * cs : 0), e_col_ok);
        }
    };
    auto stage_h = [&](int k) {                              // H_old(k) -> ring slot k % (HE_DIST+1)
        const int s = k % (HE_DIST + 1);
        const float* src = calc ? hin + (long long)(k + 1) * sz + rowoff : hin;
        cp_async16(&S.h[s][0][r][lane], src, calc);
        cp_async16(&S.h[s][1][r][lane], src + (calc ? cs : 0), calc);
        cp_async16(&S.h[s][2][r][lane], src + (calc ? 2 * cs : 0), calc);
    };

    const int kfirst = kbeg > p.Z0 ? kbeg - 1 : kbeg;        // one plane below the chunk: H_new(kbeg-1) is recomputed
    // one commit group per plane of the march: group d holds what iteration kfirst+d needs on top of the groups before it
    stage_e(kfirst); stage_e(kfirst + 1); stage_h(kfirst);
    cp_async_commit();
#pragma unroll
    for (int d = 1; d < HE_DIST; ++d) {
        if (kfirst + d < kend) { stage_e(kfirst + d + 1); stage_h(kfirst + d); }
        cp_async_commit();
    }
    long long base = (long long)(kfirst + 1) * sz + rowoff;  // plane kfirst (ghost offset +1)
    float4 hx_km = zero4(), hy_km = zero4();                 // H_new(k-1)
    if (kfirst == kbeg && own) { hx_km = ld4(hout + base - sz); hy_km = ld4(hout + cs + base - sz); }
    cp_async_wait<HE_DIST - 1>();
    __syncthreads();

    for (int k = kfirst; k < kend; ++k, base += sz) {
        const bool pro = k < kbeg;                           // prologue plane: H_new only, nothing stored
        if (k + HE_DIST < kend) { stage_e(k + HE_DIST + 1); stage_h(k + HE_DIST); }
        cp_async_commit();
        const int se = k % (HE_DIST + 2), se1 = (k + 1) % (HE_DIST + 2), sh = k % (HE_DIST + 1);
        float4 hx = zero4(), hy = zero4(), hz = zero4();
        float4 ax, ay, az, bx, by, bz;
        const float4 ex = S.e[se][0][r][lane], ey = S.e[se][1][r][lane], ez = S.e[se][2][r][lane];
        if (calc) {
            load_coefs6<CMP>(p.ii, p.iv, p.xv_h, p.meta_h, base, cs, (long long)(k + 1) * p.ny + j, p.ny, i0, p.px, ax, ay, az, bx, by, bz);
            hx = S.h[sh][0][r][lane]; hy = S.h[sh][1][r][lane]; hz = S.h[sh][2][r][lane];
        } else if (ext) {
            hx = ld4(hout + base); hy = ld4(hout + cs + base); hz = ld4(hout + 2 * cs + base);
        }
        float ez_r = __shfl_down_sync(0xffffffffu, ez.x, 1);
        float ey_r = __shfl_down_sync(0xffffffffu, ey.x, 1);
        if (lane == 31) { ez_r = S.e[se][2][r][32].x; ey_r = S.e[se][1][r][32].x; }
        if (calc) {
            const float4 ex1 = S.e[se1][0][r][lane], ey1 = S.e[se1][1][r][lane];
            const float4 ex_jp = S.e[se][0][r + 1][lane], ez_jp = S.e[se][2][r + 1][lane];
            const float4 ez_ip = make_float4(ez.y, ez.z, ez.w, ez_r);
            const float4 ey_ip = make_float4(ey.y, ey.z, ey.w, ey_r);
            hx = upd4(ax, hx, bx, ez, ez_jp, ey, ey1);
            hy = upd4(ay, hy, by, ex, ex1, ez, ez_ip);
            hz = upd4(az, hz, bz, ey, ey_ip, ex, ex_jp);
        }
        // hx, hy, hz now hold H_new(k) (zero outside the grid)
        if (own && !pro) { st4(hout + base, hx); st4(hout + cs + base, hy); st4(hout + 2 * cs + base, hz); }
        S.xb[k & 1][r][0][lane] = hz; S.xb[k & 1][r][1][lane] = hx;
        cp_async_wait<HE_DIST - 1>();                        // the next plane has landed (this thread's copies) ...
        __syncthreads();                                     // ... and everybody's, together with this plane's H_new rows
        const float hz_l = __shfl_up_sync(0xffffffffu, hz.w, 1);
        const float hy_l = __shfl_up_sync(0xffffffffu, hy.w, 1);
        if (own && !pro) {
            const float4 hz_jm = S.xb[k & 1][r - 1][0][lane], hx_jm = S.xb[k & 1][r - 1][1][lane];
            const float4 hz_im = make_float4(hz_l, hz.x, hz.y, hz.z);
            const float4 hy_im = make_float4(hy_l, hy.x, hy.y, hy.z);
            load_coefs6<CMP>(p.vv, p.vi, p.xv_e, p.meta_e, base, cs, (long long)(k + 1) * p.ny + j, p.ny, i0, p.px, ax, ay, az, bx, by, bz);
            const float4 exn = upd4(ax, ex, bx, hz, hz_jm, hy, hy_km);
            const float4 eyn = upd4(ay, ey, by, hx, hx_km, hz, hz_im);
            const float4 ezn = upd4(az, ez, bz, hy, hy_im, hx, hx_jm);
            st4(eout + base, exn); st4(eout + cs + base, eyn); st4(eout + 2 * cs + base, ezn);
        }
        hx_km = hx; hy_km = hy;
    }
}


// ---- the register version again, with the address arithmetic written out ----
// update_he_kernel spends ~60 % of its instructions on 64-bit index arithmetic and on the "row streamed in full"
// alternative of every coefficient; with 16 warps per SM that, not DRAM, bounds it.  Here every array has one per-thread
// byte pointer that advances by a plane per iteration, all other offsets are launch constants, the x-vector of a
// compressed row is one mad.wide away, and rows with a slot streamed in full take a (warp-uniform) side path.
__device__ __forceinline__ float4 ldb4(const char* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldb4_cs(const char* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldb4_nc(const char* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stb4(char* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 xv4(const float4* xs, unsigned id, float sc)      // xs = this lane's column of the smem copy
{
    const float4 v = xs[id * 32];
    return make_float4(__fmul_rn(sc, v.x), __fmul_rn(sc, v.y), __fmul_rn(sc, v.z), __fmul_rn(sc, v.w));
}
#define ROW_ANY_FULL(w1) (((w1) >> 16) & 255u)       // pad[0] of the record: set on the device when a slot is streamed in full

// coefficients of one row from its record (m0, m1 = the two float4 halves of the 32-byte record, already in registers)
__device__ __forceinline__ void row_coefs(const float4 m0, const float4 m1, const float4* xs,
        const float* __restrict__ ca, const float* __restrict__ cb, const float* __restrict__ xv, long long base, long long cs, int i0, int px,
        float4& ax, float4& ay, float4& az, float4& bx, float4& by, float4& bz)
{
    const unsigned w0 = __float_as_uint(m1.z), w1 = __float_as_uint(m1.w);
    if (ROW_ANY_FULL(w1) == 0) {
        ax = xv4(xs, w0 & 255u, m0.x);
        ay = xv4(xs, (w0 >> 8) & 255u, m0.y);
        az = xv4(xs, (w0 >> 16) & 255u, m0.z);
        bx = xv4(xs, w0 >> 24, m0.w);
        by = xv4(xs, w1 & 255u, m1.x);
        bz = xv4(xs, (w1 >> 8) & 255u, m1.y);
    } else {
        ax = coef4(w0 & 255u, m0.x, ca + base, xv, i0, px);
        ay = coef4((w0 >> 8) & 255u, m0.y, ca + cs + base, xv, i0, px);
        az = coef4((w0 >> 16) & 255u, m0.z, ca + 2 * cs + base, xv, i0, px);
        bx = coef4(w0 >> 24, m0.w, cb + base, xv, i0, px);
        by = coef4(w1 & 255u, m1.x, cb + cs + base, xv, i0, px);
        bz = coef4((w1 >> 8) & 255u, m1.y, cb + 2 * cs + base, xv, i0, px);
    }
}

template <int TY>
__global__ void __launch_bounds__(32 * (TY + 1)) __maxnreg__(TY == 5 ? 112 : (TY == 9 ? 96 : 128)) update_he2_kernel(const HeParams p)
{
    __shared__ float4 xb[2][TY + 1][2][32];
    // row records of the H and E pass, staged one plane ahead by the warp that uses them (lanes 0-3, cp.async): the
    // records steer dependent loads, so they must not cost a cache miss on the critical path of the march
    __shared__ float4 ms[2][TY + 1][4];
    // the CTA's 128-column slice of every x-vector ([nv_h + nv_e][32] float4, loaded once): a coefficient of a compressed
    // row is one LDS and one multiply
    extern __shared__ float4 xs_all[];
    const int lane = threadIdx.x, r = threadIdx.y;
    const int i0 = p.X0 - 4 + HE_SEG * (int)blockIdx.x + 4 * lane;
    {
        const bool col_ok = i0 >= 0 && i0 < p.px;
        for (int v = r; v < p.nv_h + p.nv_e; v += TY + 1) {
            const float* src = v < p.nv_h ? p.xv_h + (size_t)v * p.px : p.xv_e + (size_t)(v - p.nv_h) * p.px;
            xs_all[v * 32 + lane] = col_ok ? __ldg(reinterpret_cast<const float4*>(src + i0)) : zero4();
        }
    }
    const float4* xsh = xs_all + lane;
    const float4* xse = xs_all + p.nv_h * 32 + lane;
    const int j = p.Y0 - 1 + TY * (int)blockIdx.y + r;
    const int kbeg = p.Z0 + (int)blockIdx.z * p.kz;
    const int kend = min(kbeg + p.kz, p.Z1);
    const bool in_grid = i0 >= 0 && i0 < p.px && j >= 0 && j < p.Y1;
    const bool reg_x = (i0 >= p.X0 && i0 < p.X1) || i0 >= p.XT0;
    const bool ext = in_grid && (!reg_x || j < p.Y0);
    const bool calc = in_grid && !ext;
    const bool own = calc && lane >= 1 && r >= 1;
    const bool has_jp = in_grid && j + 1 < p.ny;
    const bool edge_load = in_grid && lane == 31 && i0 + 4 < p.px;
    const bool row_ok = j >= 0 && j < p.Y1;                  // the row has records (warp-uniform)
    const int kfirst = kbeg > p.Z0 ? kbeg - 1 : kbeg;
    const long long base0 = (long long)(kfirst + 1) * p.sz + (long long)j * p.px + i0;
    const long long mrow0 = ((long long)(kfirst + 1) * p.ny + j) * 32;
    // per-thread plane pointers (never dereferenced where the thread is outside the grid)
    const char* pe = reinterpret_cast<const char*>(p.ein + base0);
    const char* ph = reinterpret_cast<const char*>(p.hin + base0);
    char* qe = reinterpret_cast<char*>(p.eout + base0);
    char* qh = reinterpret_cast<char*>(p.hout + base0);
    // lanes 0,1 stage the H record, lanes 2,3 the E record of this warp's row
    const char* mrec = (lane < 2 ? reinterpret_cast<const char*>(p.meta_h) : reinterpret_cast<const char*>(p.meta_e)) + mrow0 + (lane & 1) * 16;

    float4 ex = zero4(), ey = zero4(), ez = zero4();         // E_old(k)
    float4 hx_km = zero4(), hy_km = zero4();                 // H_new(k-1)
    if (lane < 4) cp_async16(&ms[kfirst & 1][r][lane], row_ok ? mrec : reinterpret_cast<const char*>(p.meta_h), row_ok);
    cp_async_commit();
    if (in_grid) { ex = ldb4(pe); ey = ldb4(pe + p.b_cs); ez = ldb4(pe + p.b_2cs); }
    if (kfirst == kbeg && own) { hx_km = ldb4(qh - p.b_sz); hy_km = ldb4(qh - p.b_sz + p.b_cs); }
    cp_async_wait<0>();
    __syncthreads();                                         // x-vector slices and the first records are in place

    for (int k = kfirst; k < kend; ++k, pe += p.b_sz, ph += p.b_sz, qe += p.b_sz, qh += p.b_sz, mrec += p.meta_step) {
        const bool pro = k < kbeg;
        float4 ex1 = zero4(), ey1 = zero4(), ez1 = zero4(), ez_jp = zero4(), ex_jp = zero4();
        float4 hx = zero4(), hy = zero4(), hz = zero4();
        float4 ax = zero4(), ay = zero4(), az = zero4(), bx = zero4(), by = zero4(), bz = zero4();
        float ez_e = 0.f, ey_e = 0.f;
        // next plane's records (ghost planes have records too)
        if (lane < 4) cp_async16(&ms[(k + 1) & 1][r][lane], row_ok ? mrec + p.meta_step : reinterpret_cast<const char*>(p.meta_h), row_ok);
        cp_async_commit();
        if (in_grid) {
            if (p.pf > 0 && k + 1 < kend) {
                if (p.pf == 2) {
                    prefetch_l1(pe + p.b_pfe[0]); prefetch_l1(pe + p.b_pfe[1]); prefetch_l1(pe + p.b_pfe[2]);
                    if (calc) { prefetch_l1(ph + p.b_pfh[0]); prefetch_l1(ph + p.b_pfh[1]); prefetch_l1(ph + p.b_pfh[2]); }
                } else {
                    prefetch_l2(pe + p.b_pfe[0]); prefetch_l2(pe + p.b_pfe[1]); prefetch_l2(pe + p.b_pfe[2]);
                    if (calc) { prefetch_l2(ph + p.b_pfh[0]); prefetch_l2(ph + p.b_pfh[1]); prefetch_l2(ph + p.b_pfh[2]); }
                }
            }
            ex1 = ldb4(pe + p.b_sz); ey1 = ldb4(pe + p.b_sz_cs); ez1 = ldb4(pe + p.b_sz_2cs);
            if (has_jp) { ex_jp = ldb4(pe + p.b_row); ez_jp = ldb4(pe + p.b_row_2cs); }
        }
        if (calc) {
            hx = ldb4_cs(ph); hy = ldb4_cs(ph + p.b_cs); hz = ldb4_cs(ph + p.b_2cs);
            row_coefs(ms[k & 1][r][0], ms[k & 1][r][1], xsh, p.ii, p.iv, p.xv_h,
                      (long long)((ph - reinterpret_cast<const char*>(p.hin)) >> 2), p.cs, i0, p.px, ax, ay, az, bx, by, bz);
        } else if (ext) {
            hx = ldb4(qh); hy = ldb4(qh + p.b_cs); hz = ldb4(qh + p.b_2cs);
        }
        if (edge_load) { ey_e = *reinterpret_cast<const float*>(pe + p.b_cs + 16); ez_e = *reinterpret_cast<const float*>(pe + p.b_2cs + 16); }
        float ez_r = __shfl_down_sync(0xffffffffu, ez.x, 1);
        float ey_r = __shfl_down_sync(0xffffffffu, ey.x, 1);
        if (lane == 31) { ez_r = ez_e; ey_r = ey_e; }
        if (calc) {
            const float4 ez_ip = make_float4(ez.y, ez.z, ez.w, ez_r);
            const float4 ey_ip = make_float4(ey.y, ey.z, ey.w, ey_r);
            hx = upd4(ax, hx, bx, ez, ez_jp, ey, ey1);
            hy = upd4(ay, hy, by, ex, ex1, ez, ez_ip);
            hz = upd4(az, hz, bz, ey, ey_ip, ex, ex_jp);
        }
        if (own && !pro) {
            stb4(qh, hx); stb4(qh + p.b_cs, hy); stb4(qh + p.b_2cs, hz);
            // the E coefficients do not depend on H_new: fetch them before the barrier, into the registers the H pass freed
            row_coefs(ms[k & 1][r][2], ms[k & 1][r][3], xse, p.vv, p.vi, p.xv_e,
                      (long long)((pe - reinterpret_cast<const char*>(p.ein)) >> 2), p.cs, i0, p.px, ax, ay, az, bx, by, bz);
        }
        xb[k & 1][r][0][lane] = hz; xb[k & 1][r][1][lane] = hx;
        cp_async_wait<0>();                                  // next plane's records (this warp's own copies)
        __syncthreads();
        const float hz_l = __shfl_up_sync(0xffffffffu, hz.w, 1);
        const float hy_l = __shfl_up_sync(0xffffffffu, hy.w, 1);
        if (own && !pro) {
            const float4 hz_jm = xb[k & 1][r - 1][0][lane], hx_jm = xb[k & 1][r - 1][1][lane];
            const float4 hz_im = make_float4(hz_l, hz.x, hz.y, hz.z);
            const float4 hy_im = make_float4(hy_l, hy.x, hy.y, hy.z);
            ex = upd4(ax, ex, bx, hz, hz_jm, hy, hy_km);
            ey = upd4(ay, ey, by, hx, hx_km, hz, hz_im);
            ez = upd4(az, ez, bz, hy, hy_im, hx, hx_jm);
            stb4(qe, ex); stb4(qe + p.b_cs, ey); stb4(qe + p.b_2cs, ez);
        }
        hx_km = hx; hy_km = hy;
        ex = ex1; ey = ey1; ez = ez1;
    }
}


// ---- update_he2_kernel with the field planes staged one plane ahead by cp.async ----
// ncu on update_he2_kernel: a third of all stall samples sit on the first use of the plane's global loads (L2 latency
// under load, 16 warps per SM to hide it) and a fifth on the barrier.  Here every thread copies the float4s it needs for
// the NEXT plane straight into shared memory at the top of the iteration (LDGSTS: no registers held, a whole plane of the
// CTA in flight while the current one is computed); the plane being computed is read from shared memory, including the
// +1 row (y-neighbour) and the +1 column of lane 31, so no second global load and no carried E registers.  With the
// x-vector slices already in shared memory the smaller L1 no longer matters (it did for update_he_staged_kernel).
//   es: E ring, 3 planes (k, k+1 in use, k+2 landing)  [3][3 comps][TY+2 rows][33 float4]
//   hs: H ring, 2 planes (k in use, k+1 landing)       [2][3 comps][TY+1 rows][32 float4]
template <int TY>
struct He3Smem {
    float4 es[3][3][TY + 2][33];
    float4 hs[2][3][TY + 1][32];
    float4 xb[2][TY + 1][2][32];
    float4 ms[2][TY + 1][4];
    float4 xs[1];                                            // [nv_h + nv_e][32], sized at launch
};

template <int TY>
__global__ void __launch_bounds__(32 * (TY + 1), 16 / (TY + 1)) update_he3_kernel(const HeParams p)
{
    extern __shared__ __align__(16) unsigned char he3_raw[];
    He3Smem<TY>& S = *reinterpret_cast<He3Smem<TY>*>(he3_raw);
    const int lane = threadIdx.x, r = threadIdx.y;
    const int i0 = p.X0 - 4 + HE_SEG * (int)blockIdx.x + 4 * lane;
    const int j = p.Y0 - 1 + TY * (int)blockIdx.y + r;
    const int kbeg = p.Z0 + (int)blockIdx.z * p.kz;
    const int kend = min(kbeg + p.kz, p.Z1);
    const bool col_ok = i0 >= 0 && i0 < p.px;
    {
        for (int v = r; v < p.nv_h + p.nv_e; v += TY + 1) {
            const float* src = v < p.nv_h ? p.xv_h + (size_t)v * p.px : p.xv_e + (size_t)(v - p.nv_h) * p.px;
            S.xs[v * 32 + lane] = col_ok ? __ldg(reinterpret_cast<const float4*>(src + i0)) : zero4();
        }
    }
    const float4* xsh = S.xs + lane;
    const float4* xse = S.xs + p.nv_h * 32 + lane;
    const bool in_grid = col_ok && j >= 0 && j < p.Y1;
    const bool reg_x = (i0 >= p.X0 && i0 < p.X1) || i0 >= p.XT0;
    const bool ext = in_grid && (!reg_x || j < p.Y0);
    const bool calc = in_grid && !ext;
    const bool own = calc && lane >= 1 && r >= 1;
    const bool row_ok = j >= 0 && j < p.Y1;
    // what this thread stages per plane: its own float4 of E (rows up to Y1, the +1 row of the last owned row), the +1 row
    // for the top warp, the +1 column for lane 31, its own float4 of H_old where H_new is computed, 16 bytes of a record
    const bool e_ok = col_ok && j >= 0 && j < p.ny && j <= p.Y1;
    const bool top = r == TY;
    const bool e_top_ok = top && col_ok && j + 1 >= 0 && j + 1 < p.ny && j + 1 <= p.Y1;
    const bool e_col_ok = lane == 31 && i0 + 4 < p.px && j >= 0 && j < p.Y1;
    const int kfirst = kbeg > p.Z0 ? kbeg - 1 : kbeg;
    const long long base0 = (long long)(kfirst + 1) * p.sz + (long long)j * p.px + i0;
    const long long mrow0 = ((long long)(kfirst + 1) * p.ny + j) * 32;
    const char* pe = reinterpret_cast<const char*>(p.ein + base0);          // plane k of E_old / H_old / outputs
    const char* ph = reinterpret_cast<const char*>(p.hin + base0);
    char* qe = reinterpret_cast<char*>(p.eout + base0);
    char* qh = reinterpret_cast<char*>(p.hout + base0);
    const char* mrec = (lane < 2 ? reinterpret_cast<const char*>(p.meta_h) : reinterpret_cast<const char*>(p.meta_e)) + mrow0 + (lane & 1) * 16;
    const char* safe = reinterpret_cast<const char*>(p.ein);               // any valid address for the zero-fill copies

    // stage E_old of the plane d planes above the current one into ring slot se
    auto stage_e = [&](int se, long long d) {
        const char* b = pe + d;
        cp_async16(&S.es[se][0][r][lane], e_ok ? b : safe, e_ok);
        cp_async16(&S.es[se][1][r][lane], e_ok ? b + p.b_cs : safe, e_ok);
        cp_async16(&S.es[se][2][r][lane], e_ok ? b + p.b_2cs : safe, e_ok);
        if (top) {
            cp_async16(&S.es[se][0][TY + 1][lane], e_top_ok ? b + p.b_row : safe, e_top_ok);
            cp_async16(&S.es[se][2][TY + 1][lane], e_top_ok ? b + p.b_row_2cs : safe, e_top_ok);
        }
        if (lane == 31) {
            cp_async16(&S.es[se][1][r][32], e_col_ok ? b + p.b_cs + 16 : safe, e_col_ok);
            cp_async16(&S.es[se][2][r][32], e_col_ok ? b + p.b_2cs + 16 : safe, e_col_ok);
        }
    };
    auto stage_h = [&](int sh, long long d) {
        const char* b = ph + d;
        cp_async16(&S.hs[sh][0][r][lane], calc ? b : safe, calc);
        cp_async16(&S.hs[sh][1][r][lane], calc ? b + p.b_cs : safe, calc);
        cp_async16(&S.hs[sh][2][r][lane], calc ? b + p.b_2cs : safe, calc);
    };

    int se = 0, sh = 0;                                      // ring slots of plane k
    stage_e(0, 0); stage_e(1, p.b_sz); stage_h(0, 0);
    if (lane < 4) cp_async16(&S.ms[0][r][lane], row_ok ? mrec : safe, row_ok);
    cp_async_commit();
    float4 hx_km = zero4(), hy_km = zero4();                 // H_new(k-1)
    if (kfirst == kbeg && own) { hx_km = ldb4(qh - p.b_sz); hy_km = ldb4(qh - p.b_sz + p.b_cs); }
    cp_async_wait<0>();
    __syncthreads();

    for (int k = kfirst; k < kend; ++k, pe += p.b_sz, ph += p.b_sz, qe += p.b_sz, qh += p.b_sz, mrec += p.meta_step) {
        const bool pro = k < kbeg;
        const int se1 = se == 2 ? 0 : se + 1, se2 = se1 == 2 ? 0 : se1 + 1, mb = (k - kfirst) & 1;
        if (k + 1 < kend) {                                  // next iteration's new data: E(k+2), H_old(k+1), records of k+1
            stage_e(se2, 2 * p.b_sz); stage_h(sh ^ 1, p.b_sz);
            if (lane < 4) cp_async16(&S.ms[mb ^ 1][r][lane], row_ok ? mrec + p.meta_step : safe, row_ok);
        }
        cp_async_commit();
        float4 hx = zero4(), hy = zero4(), hz = zero4();
        float4 ax = zero4(), ay = zero4(), az = zero4(), bx = zero4(), by = zero4(), bz = zero4();
        const float4 ex = S.es[se][0][r][lane], ey = S.es[se][1][r][lane], ez = S.es[se][2][r][lane];
        if (calc) {
            hx = S.hs[sh][0][r][lane]; hy = S.hs[sh][1][r][lane]; hz = S.hs[sh][2][r][lane];
            row_coefs(S.ms[mb][r][0], S.ms[mb][r][1], xsh, p.ii, p.iv, p.xv_h,
                      (long long)((ph - reinterpret_cast<const char*>(p.hin)) >> 2), p.cs, i0, p.px, ax, ay, az, bx, by, bz);
        } else if (ext) {
            hx = ldb4(qh); hy = ldb4(qh + p.b_cs); hz = ldb4(qh + p.b_2cs);
        }
        float ez_r = __shfl_down_sync(0xffffffffu, ez.x, 1);
        float ey_r = __shfl_down_sync(0xffffffffu, ey.x, 1);
        if (lane == 31) { ez_r = S.es[se][2][r][32].x; ey_r = S.es[se][1][r][32].x; }
        if (calc) {
            const float4 ex1 = S.es[se1][0][r][lane], ey1 = S.es[se1][1][r][lane];
            const float4 ex_jp = S.es[se][0][r + 1][lane], ez_jp = S.es[se][2][r + 1][lane];
            const float4 ez_ip = make_float4(ez.y, ez.z, ez.w, ez_r);
            const float4 ey_ip = make_float4(ey.y, ey.z, ey.w, ey_r);
            hx = upd4(ax, hx, bx, ez, ez_jp, ey, ey1);
            hy = upd4(ay, hy, by, ex, ex1, ez, ez_ip);
            hz = upd4(az, hz, bz, ey, ey_ip, ex, ex_jp);
        }
        if (own && !pro) {
            stb4(qh, hx); stb4(qh + p.b_cs, hy); stb4(qh + p.b_2cs, hz);
            row_coefs(S.ms[mb][r][2], S.ms[mb][r][3], xse, p.vv, p.vi, p.xv_e,
                      (long long)((pe - reinterpret_cast<const char*>(p.ein)) >> 2), p.cs, i0, p.px, ax, ay, az, bx, by, bz);
        }
        S.xb[mb][r][0][lane] = hz; S.xb[mb][r][1][lane] = hx;
        cp_async_wait<0>();                                  // next plane has landed (this thread's copies) ...
        __syncthreads();                                     // ... and everybody's, together with this plane's H_new rows
        const float hz_l = __shfl_up_sync(0xffffffffu, hz.w, 1);
        const float hy_l = __shfl_up_sync(0xffffffffu, hy.w, 1);
        if (own && !pro) {
            const float4 hz_jm = S.xb[mb][r - 1][0][lane], hx_jm = S.xb[mb][r - 1][1][lane];
            const float4 hz_im = make_float4(hz_l, hz.x, hz.y, hz.z);
            const float4 hy_im = make_float4(hy_l, hy.x, hy.y, hy.z);
            const float4 exn = upd4(ax, ex, bx, hz, hz_jm, hy, hy_km);
            const float4 eyn = upd4(ay, ey, by, hx, hx_km, hz, hz_im);
            const float4 ezn = upd4(az, ez, bz, hy, hy_im, hx, hx_jm);
            stb4(qe, exn); stb4(qe + p.b_cs, eyn); stb4(qe + p.b_2cs, ezn);
        }
        hx_km = hx; hy_km = hy;
        se = se1; sh ^= 1;
    }
}


// ---- update_he3_kernel without the CTA barrier ----
// ncu on update_he3_kernel: 30 % of the stall samples sit on the per-plane __syncthreads (8 warps in lock step, the
// slowest warp's memory latency is everybody's).  A warp only needs its two neighbours: the H_new row of the warp below,
// the staged +1 row of the warp above.  Three monotonic per-warp counters in shared memory replace the barrier:
//   prod[w] = planes whose H_new row warp w has published (xb is 2 deep: w waits for rd[w+1] >= t-1 before reuse)
//   stg[w]  = planes whose staged copies of warp w have landed (+1: plane t+1 is in place when stg[w] >= t+2)
//   rd[w]   = planes for which warp w is done reading other warps' data (w+1 may then reuse the E ring slot)
// Every wait is for a warp at an earlier or equal plane, so the slowest warp can always proceed (no cycle).
__device__ __forceinline__ void spin_ge(const volatile int* f, int v)
{
    while (*f < v) { }
    __threadfence_block();
}
__device__ __forceinline__ void publish1(volatile int* f, int v)
{
    __syncwarp();
    __threadfence_block();
    if (threadIdx.x == 0) *f = v;
}
__device__ __forceinline__ void publish2(volatile int* f, int v, volatile int* g, int w)
{
    __syncwarp();
    __threadfence_block();
    if (threadIdx.x == 0) { *f = v; *g = w; }
}
template <int TY>
struct He4Smem {
    float4 es[3][3][TY + 2][33];
    float4 hs[2][3][TY + 1][32];
    float4 xb[2][TY + 1][2][32];
    float4 ms[2][TY + 1][4];
    int prod[TY + 2], rd[TY + 2], stg[TY + 2];               // per-warp progress counters (see update_he4_kernel)
    int pad_[(4 - (3 * (TY + 2)) % 4) % 4];
    float4 xs[1];                                            // [nv_h + nv_e][32], sized at launch
};

template <int TY>
__global__ void __launch_bounds__(32 * (TY + 1), 16 / (TY + 1)) update_he4_kernel(const HeParams p)
{
    extern __shared__ __align__(16) unsigned char he4_raw[];
    He4Smem<TY>& S = *reinterpret_cast<He4Smem<TY>*>(he4_raw);
    const int lane = threadIdx.x, r = threadIdx.y;
    const int i0 = p.X0 - 4 + HE_SEG * (int)blockIdx.x + 4 * lane;
    const int j = p.Y0 - 1 + TY * (int)blockIdx.y + r;
    const int kbeg = p.Z0 + (int)blockIdx.z * p.kz;
    const int kend = min(kbeg + p.kz, p.Z1);
    const bool col_ok = i0 >= 0 && i0 < p.px;
    {
        for (int v = r; v < p.nv_h + p.nv_e; v += TY + 1) {
            const float* src = v < p.nv_h ? p.xv_h + (size_t)v * p.px : p.xv_e + (size_t)(v - p.nv_h) * p.px;
            S.xs[v * 32 + lane] = col_ok ? __ldg(reinterpret_cast<const float4*>(src + i0)) : zero4();
        }
    }
    const float4* xsh = S.xs + lane;
    const float4* xse = S.xs + p.nv_h * 32 + lane;
    const bool in_grid = col_ok && j >= 0 && j < p.Y1;
    const bool reg_x = (i0 >= p.X0 && i0 < p.X1) || i0 >= p.XT0;
    const bool ext = in_grid && (!reg_x || j < p.Y0);
    const bool calc = in_grid && !ext;
    const bool own = calc && lane >= 1 && r >= 1;
    const bool row_ok = j >= 0 && j < p.Y1;
    // what this thread stages per plane: its own float4 of E (rows up to Y1, the +1 row of the last owned row), the +1 row
    // for the top warp, the +1 column for lane 31, its own float4 of H_old where H_new is computed, 16 bytes of a record
    const bool e_ok = col_ok && j >= 0 && j < p.ny && j <= p.Y1;
    const bool top = r == TY;
    const bool e_top_ok = top && col_ok && j + 1 >= 0 && j + 1 < p.ny && j + 1 <= p.Y1;
    const bool e_col_ok = lane == 31 && i0 + 4 < p.px && j >= 0 && j < p.Y1;
    const int kfirst = kbeg > p.Z0 ? kbeg - 1 : kbeg;
    const long long base0 = (long long)(kfirst + 1) * p.sz + (long long)j * p.px + i0;
    const long long mrow0 = ((long long)(kfirst + 1) * p.ny + j) * 32;
    const char* pe = reinterpret_cast<const char*>(p.ein + base0);          // plane k of E_old / H_old / outputs
    const char* ph = reinterpret_cast<const char*>(p.hin + base0);
    char* qe = reinterpret_cast<char*>(p.eout + base0);
    char* qh = reinterpret_cast<char*>(p.hout + base0);
    const char* mrec = (lane < 2 ? reinterpret_cast<const char*>(p.meta_h) : reinterpret_cast<const char*>(p.meta_e)) + mrow0 + (lane & 1) * 16;
    const char* safe = reinterpret_cast<const char*>(p.ein);               // any valid address for the zero-fill copies

    // stage E_old of the plane d planes above the current one into ring slot se
    auto stage_e = [&](int se, long long d) {
        const char* b = pe + d;
        cp_async16(&S.es[se][0][r][lane], e_ok ? b : safe, e_ok);
        cp_async16(&S.es[se][1][r][lane], e_ok ? b + p.b_cs : safe, e_ok);
        cp_async16(&S.es[se][2][r][lane], e_ok ? b + p.b_2cs : safe, e_ok);
        if (top) {
            cp_async16(&S.es[se][0][TY + 1][lane], e_top_ok ? b + p.b_row : safe, e_top_ok);
            cp_async16(&S.es[se][2][TY + 1][lane], e_top_ok ? b + p.b_row_2cs : safe, e_top_ok);
        }
        if (lane == 31) {
            cp_async16(&S.es[se][1][r][32], e_col_ok ? b + p.b_cs + 16 : safe, e_col_ok);
            cp_async16(&S.es[se][2][r][32], e_col_ok ? b + p.b_2cs + 16 : safe, e_col_ok);
        }
    };
    auto stage_h = [&](int sh, long long d) {
        const char* b = ph + d;
        cp_async16(&S.hs[sh][0][r][lane], calc ? b : safe, calc);
        cp_async16(&S.hs[sh][1][r][lane], calc ? b + p.b_cs : safe, calc);
        cp_async16(&S.hs[sh][2][r][lane], calc ? b + p.b_2cs : safe, calc);
    };

    int se = 0, sh = 0;                                      // ring slots of plane k
    stage_e(0, 0); stage_e(1, p.b_sz); stage_h(0, 0);
    if (lane < 4) cp_async16(&S.ms[0][r][lane], row_ok ? mrec : safe, row_ok);
    cp_async_commit();
    float4 hx_km = zero4(), hy_km = zero4();                 // H_new(k-1)
    if (kfirst == kbeg && own) { hx_km = ldb4(qh - p.b_sz); hy_km = ldb4(qh - p.b_sz + p.b_cs); }
    if (lane == 0) { S.prod[r] = 0; S.rd[r] = 0; S.stg[r] = 1; }
    cp_async_wait<0>();
    __syncthreads();
    volatile int* const prod = S.prod; volatile int* const rd = S.rd; volatile int* const stg = S.stg;

    for (int k = kfirst; k < kend; ++k, pe += p.b_sz, ph += p.b_sz, qe += p.b_sz, qh += p.b_sz, mrec += p.meta_step) {
        const bool pro = k < kbeg;
        const int t = k - kfirst;
        const int se1 = se == 2 ? 0 : se + 1, se2 = se1 == 2 ? 0 : se1 + 1, mb = t & 1;
        // ring slot se2 held plane k-1, whose row r the warp below read during its iteration t-1
        if (r >= 1 && t >= 1) spin_ge(&rd[r - 1], t);
        if (k + 1 < kend) {                                  // next iteration's new data: E(k+2), H_old(k+1), records of k+1
            stage_e(se2, 2 * p.b_sz); stage_h(sh ^ 1, p.b_sz);
            if (lane < 4) cp_async16(&S.ms[mb ^ 1][r][lane], row_ok ? mrec + p.meta_step : safe, row_ok);
        }
        cp_async_commit();
        float4 hx = zero4(), hy = zero4(), hz = zero4();
        float4 ax = zero4(), ay = zero4(), az = zero4(), bx = zero4(), by = zero4(), bz = zero4();
        if (r < TY) spin_ge(&stg[r + 1], t + 1);             // row r+1 of plane k is staged by the warp above
        const float4 ex = S.es[se][0][r][lane], ey = S.es[se][1][r][lane], ez = S.es[se][2][r][lane];
        if (calc) {
            hx = S.hs[sh][0][r][lane]; hy = S.hs[sh][1][r][lane]; hz = S.hs[sh][2][r][lane];
            row_coefs(S.ms[mb][r][0], S.ms[mb][r][1], xsh, p.ii, p.iv, p.xv_h,
                      (long long)((ph - reinterpret_cast<const char*>(p.hin)) >> 2), p.cs, i0, p.px, ax, ay, az, bx, by, bz);
        } else if (ext) {
            hx = ldb4(qh); hy = ldb4(qh + p.b_cs); hz = ldb4(qh + p.b_2cs);
        }
        float ez_r = __shfl_down_sync(0xffffffffu, ez.x, 1);
        float ey_r = __shfl_down_sync(0xffffffffu, ey.x, 1);
        if (lane == 31) { ez_r = S.es[se][2][r][32].x; ey_r = S.es[se][1][r][32].x; }
        if (calc) {
            const float4 ex1 = S.es[se1][0][r][lane], ey1 = S.es[se1][1][r][lane];
            const float4 ex_jp = S.es[se][0][r + 1][lane], ez_jp = S.es[se][2][r + 1][lane];
            const float4 ez_ip = make_float4(ez.y, ez.z, ez.w, ez_r);
            const float4 ey_ip = make_float4(ey.y, ey.z, ey.w, ey_r);
            hx = upd4(ax, hx, bx, ez, ez_jp, ey, ey1);
            hy = upd4(ay, hy, by, ex, ex1, ez, ez_ip);
            hz = upd4(az, hz, bz, ey, ey_ip, ex, ex_jp);
        }
        if (own && !pro) {
            stb4(qh, hx); stb4(qh + p.b_cs, hy); stb4(qh + p.b_2cs, hz);
            row_coefs(S.ms[mb][r][2], S.ms[mb][r][3], xse, p.vv, p.vi, p.xv_e,
                      (long long)((pe - reinterpret_cast<const char*>(p.ein)) >> 2), p.cs, i0, p.px, ax, ay, az, bx, by, bz);
        }
        // publish this plane's H_new row for the warp above (it must have consumed the row of two planes ago) ...
        if (r < TY && t >= 2) spin_ge(&rd[r + 1], t - 1);
        S.xb[mb][r][0][lane] = hz; S.xb[mb][r][1][lane] = hx;
        cp_async_wait<0>();                                  // ... and the staged data of the next plane (this warp's copies)
        publish2(&prod[r], t + 1, &stg[r], t + 2);
        if (r >= 1) spin_ge(&prod[r - 1], t + 1);            // the row below has published H_new(k)
        const float hz_l = __shfl_up_sync(0xffffffffu, hz.w, 1);
        const float hy_l = __shfl_up_sync(0xffffffffu, hy.w, 1);
        float4 hz_jm = zero4(), hx_jm = zero4();
        if (r >= 1) { hz_jm = S.xb[mb][r - 1][0][lane]; hx_jm = S.xb[mb][r - 1][1][lane]; }
        publish1(&rd[r], t + 1);                             // done with every other warp's data of this plane
        if (own && !pro) {
            const float4 hz_im = make_float4(hz_l, hz.x, hz.y, hz.z);
            const float4 hy_im = make_float4(hy_l, hy.x, hy.y, hy.z);
            const float4 exn = upd4(ax, ex, bx, hz, hz_jm, hy, hy_km);
            const float4 eyn = upd4(ay, ey, by, hx, hx_km, hz, hz_im);
            const float4 ezn = upd4(az, ez, bz, hy, hy_im, hx, hx_jm);
            stb4(qe, exn); stb4(qe + p.b_cs, eyn); stb4(qe + p.b_2cs, ezn);
        }
        hx_km = hx; hy_km = hy;
        se = se1; sh ^= 1;
    }
}



// ---- update_he3_kernel with the planes staged by the TMA engine (cp.async.bulk + mbarrier) ----
// The row segments a CTA stages are contiguous in global memory (33 float4 of E, 32 of H per row and component), so one
// elected lane per warp hands them to the TMA engine as 1-D bulk copies that complete on an mbarrier; the 256 threads no
// longer spend ~25 instructions each per plane on LDGSTS and their addresses.  Out-of-grid parts of the ring are zeroed
// once at the start and never written again (a bulk copy only covers the in-grid part of its row).
// full[b]: completion of the copies issued during iteration t (consumed in iteration t+1), b = (t+1) & 1.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int TY>
struct He5Smem {
    float4 es[3][3][TY + 2][33];
    float4 hs[2][3][TY + 1][32];
    float4 xb[2][TY + 1][2][32];
    float4 ms[2][TY + 1][4];
    unsigned long long full[2];
    float4 xs[1];                                            // [nv_h + nv_e][32], sized at launch
};

template <int TY>
__global__ void __launch_bounds__(32 * (TY + 1), 16 / (TY + 1)) update_he5_kernel(const HeParams p)
{
    extern __shared__ __align__(16) unsigned char he5_raw[];
    He5Smem<TY>& S = *reinterpret_cast<He5Smem<TY>*>(he5_raw);
    const int lane = threadIdx.x, r = threadIdx.y;
    const int i_seg = p.X0 - 4 + HE_SEG * (int)blockIdx.x;  // column of lane 0
    const int i0 = i_seg + 4 * lane;
    const int j = p.Y0 - 1 + TY * (int)blockIdx.y + r;
    const int kbeg = p.Z0 + (int)blockIdx.z * p.kz;
    const int kend = min(kbeg + p.kz, p.Z1);
    const bool col_ok = i0 >= 0 && i0 < p.px;
    {   // x-vector slices; zero the rings (the out-of-grid parts stay zero for the whole launch)
        for (int v = r; v < p.nv_h + p.nv_e; v += TY + 1) {
            const float* src = v < p.nv_h ? p.xv_h + (size_t)v * p.px : p.xv_e + (size_t)(v - p.nv_h) * p.px;
            S.xs[v * 32 + lane] = col_ok ? __ldg(reinterpret_cast<const float4*>(src + i0)) : zero4();
        }
        float4* z = &S.es[0][0][0][0];
        const int nz4 = (int)((sizeof(S.es) + sizeof(S.hs)) / sizeof(float4));
        for (int q = r * 32 + lane; q < nz4; q += 32 * (TY + 1)) z[q] = zero4();
        if (r == 0 && lane == 0) { mbar_init(&S.full[0], TY + 1); mbar_init(&S.full[1], TY + 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zeros are ordered before the TMA writes
    }
    __syncthreads();
    const float4* xsh = S.xs + lane;
    const float4* xse = S.xs + p.nv_h * 32 + lane;
    const bool in_grid = col_ok && j >= 0 && j < p.Y1;
    const bool reg_x = (i0 >= p.X0 && i0 < p.X1) || i0 >= p.XT0;
    const bool ext = in_grid && (!reg_x || j < p.Y0);
    const bool calc = in_grid && !ext;
    const bool own = calc && lane >= 1 && r >= 1;
    const bool row_ok = j >= 0 && j < p.Y1;
    // warp-uniform staging plan: in-grid float4 range [c_lo, c_hi) of the 33-wide E row segment (32-wide for H)
    const int c_lo = i_seg < 0 ? (-i_seg + 3) / 4 : 0;
    const int c_hi_e = min(33, (p.px - i_seg) / 4), c_hi_h = min(32, (p.px - i_seg) / 4);
    const bool e_row = j >= 0 && j < p.ny && j <= p.Y1 && c_hi_e > c_lo;         // this warp's own row of E
    const bool e_top = r == TY && j + 1 >= 0 && j + 1 < p.ny && j + 1 <= p.Y1 && c_hi_e > c_lo;   // +1 row (top warp)
    const bool h_row = j >= p.Y0 && j < p.Y1 && c_hi_h > c_lo;                    // H_old where H_new is computed (j >= Y0 >= 0)
    const unsigned nb_e = (unsigned)(c_hi_e - c_lo) * 16u, nb_h = (unsigned)(c_hi_h - c_lo) * 16u;
    const int kfirst = kbeg > p.Z0 ? kbeg - 1 : kbeg;
    const long long base0 = (long long)(kfirst + 1) * p.sz + (long long)j * p.px + i0;
    const long long mrow0 = ((long long)(kfirst + 1) * p.ny + j) * 32;
    const char* pe = reinterpret_cast<const char*>(p.ein + base0);
    const char* ph = reinterpret_cast<const char*>(p.hin + base0);
    char* qe = reinterpret_cast<char*>(p.eout + base0);
    char* qh = reinterpret_cast<char*>(p.hout + base0);
    // lane 0's view: start of the in-grid part of this warp's row segment in plane k
    const long long seg0 = (long long)(kfirst + 1) * p.sz + (long long)j * p.px + i_seg + 4 * c_lo;
    const char* ge = reinterpret_cast<const char*>(p.ein + seg0);
    const char* gh = reinterpret_cast<const char*>(p.hin + seg0);
    const char* mrec = (lane < 2 ? reinterpret_cast<const char*>(p.meta_h) : reinterpret_cast<const char*>(p.meta_e)) + mrow0 + (lane & 1) * 16;
    const char* safe = reinterpret_cast<const char*>(p.ein);

    // lane 0 of every warp: hand the warp's rows of E_old(plane k + de) and H_old(plane k + dh) to the TMA engine
    auto stage = [&](unsigned long long* bar, int se, bool with_e, long long de, int se_b, bool with_e2, long long de2,
                     int shs, bool with_h, long long dh) {
        if (lane != 0) return;
        unsigned bytes = 0;
        if (with_e) bytes += (e_row ? 3u * nb_e : 0u) + (e_top ? 2u * nb_e : 0u);
        if (with_e2) bytes += (e_row ? 3u * nb_e : 0u) + (e_top ? 2u * nb_e : 0u);
        if (with_h && h_row) bytes += 3u * nb_h;
        mbar_arrive_expect(bar, bytes);
        auto rows_e = [&](int slot, long long d) {
            if (e_row) {
                bulk_g2s(&S.es[slot][0][r][c_lo], ge + d, nb_e, bar);
                bulk_g2s(&S.es[slot][1][r][c_lo], ge + d + p.b_cs, nb_e, bar);
                bulk_g2s(&S.es[slot][2][r][c_lo], ge + d + p.b_2cs, nb_e, bar);
            }
            if (e_top) {
                bulk_g2s(&S.es[slot][0][TY + 1][c_lo], ge + d + p.b_row, nb_e, bar);
                bulk_g2s(&S.es[slot][2][TY + 1][c_lo], ge + d + p.b_row_2cs, nb_e, bar);
            }
        };
        if (with_e) rows_e(se, de);
        if (with_e2) rows_e(se_b, de2);
        if (with_h && h_row) {
            bulk_g2s(&S.hs[shs][0][r][c_lo], gh + dh, nb_h, bar);
            bulk_g2s(&S.hs[shs][1][r][c_lo], gh + dh + p.b_cs, nb_h, bar);
            bulk_g2s(&S.hs[shs][2][r][c_lo], gh + dh + p.b_2cs, nb_h, bar);
        }
    };

    int se = 0, sh = 0;
    stage(&S.full[0], 0, true, 0, 1, true, p.b_sz, 0, true, 0);              // planes kfirst, kfirst+1 of E, kfirst of H
    if (lane < 4) cp_async16(&S.ms[0][r][lane], row_ok ? mrec : safe, row_ok);
    cp_async_commit();
    float4 hx_km = zero4(), hy_km = zero4();
    if (kfirst == kbeg && own) { hx_km = ldb4(qh - p.b_sz); hy_km = ldb4(qh - p.b_sz + p.b_cs); }
    cp_async_wait<0>();
    __syncwarp();

    for (int k = kfirst; k < kend; ++k, pe += p.b_sz, ph += p.b_sz, qe += p.b_sz, qh += p.b_sz, ge += p.b_sz, gh += p.b_sz, mrec += p.meta_step) {
        const bool pro = k < kbeg;
        const int t = k - kfirst;
        const int se1 = se == 2 ? 0 : se + 1, se2 = se1 == 2 ? 0 : se1 + 1, mb = t & 1;
        // next iteration's new data: E(k+2) -> slot se2, H_old(k+1) -> slot sh^1 (every warp arrives, with or without bytes)
        stage(&S.full[(t + 1) & 1], se2, k + 1 < kend, 2 * p.b_sz, 0, false, 0, sh ^ 1, k + 1 < kend, p.b_sz);
        if (k + 1 < kend && lane < 4) cp_async16(&S.ms[mb ^ 1][r][lane], row_ok ? mrec + p.meta_step : safe, row_ok);
        cp_async_commit();
        mbar_wait(&S.full[t & 1], (unsigned)(t >> 1) & 1u);  // this plane's staged rows have landed (all warps' copies)
        float4 hx = zero4(), hy = zero4(), hz = zero4();
        float4 ax = zero4(), ay = zero4(), az = zero4(), bx = zero4(), by = zero4(), bz = zero4();
        const float4 ex = S.es[se][0][r][lane], ey = S.es[se][1][r][lane], ez = S.es[se][2][r][lane];
        if (calc) {
            hx = S.hs[sh][0][r][lane]; hy = S.hs[sh][1][r][lane]; hz = S.hs[sh][2][r][lane];
            row_coefs(S.ms[mb][r][0], S.ms[mb][r][1], xsh, p.ii, p.iv, p.xv_h,
                      (long long)((ph - reinterpret_cast<const char*>(p.hin)) >> 2), p.cs, i0, p.px, ax, ay, az, bx, by, bz);
        } else if (ext) {
            hx = ldb4(qh); hy = ldb4(qh + p.b_cs); hz = ldb4(qh + p.b_2cs);
        }
        float ez_r = __shfl_down_sync(0xffffffffu, ez.x, 1);
        float ey_r = __shfl_down_sync(0xffffffffu, ey.x, 1);
        if (lane == 31) { ez_r = S.es[se][2][r][32].x; ey_r = S.es[se][1][r][32].x; }
        if (calc) {
            const float4 ex1 = S.es[se1][0][r][lane], ey1 = S.es[se1][1][r][lane];
            const float4 ex_jp = S.es[se][0][r + 1][lane], ez_jp = S.es[se][2][r + 1][lane];
            const float4 ez_ip = make_float4(ez.y, ez.z, ez.w, ez_r);
            const float4 ey_ip = make_float4(ey.y, ey.z, ey.w, ey_r);
            hx = upd4(ax, hx, bx, ez, ez_jp, ey, ey1);
            hy = upd4(ay, hy, by, ex, ex1, ez, ez_ip);
            hz = upd4(az, hz, bz, ey, ey_ip, ex, ex_jp);
        }
        if (own && !pro) {
            stb4(qh, hx); stb4(qh + p.b_cs, hy); stb4(qh + p.b_2cs, hz);
            row_coefs(S.ms[mb][r][2], S.ms[mb][r][3], xse, p.vv, p.vi, p.xv_e,
                      (long long)((pe - reinterpret_cast<const char*>(p.ein)) >> 2), p.cs, i0, p.px, ax, ay, az, bx, by, bz);
        }
        S.xb[mb][r][0][lane] = hz; S.xb[mb][r][1][lane] = hx;
        cp_async_wait<0>();                                  // next plane's records (this warp's own copies)
        __syncthreads();                                     // H_new rows visible; everybody is done with slots se2 / sh^1's old planes
        const float hz_l = __shfl_up_sync(0xffffffffu, hz.w, 1);
        const float hy_l = __shfl_up_sync(0xffffffffu, hy.w, 1);
        if (own && !pro) {
            const float4 hz_jm = S.xb[mb][r - 1][0][lane], hx_jm = S.xb[mb][r - 1][1][lane];
            const float4 hz_im = make_float4(hz_l, hz.x, hz.y, hz.z);
            const float4 hy_im = make_float4(hy_l, hy.x, hy.y, hy.z);
            const float4 exn = upd4(ax, ex, bx, hz, hz_jm, hy, hy_km);
            const float4 eyn = upd4(ay, ey, by, hx, hx_km, hz, hz_im);
            const float4 ezn = upd4(az, ez, bz, hy, hy_im, hx, hx_jm);
            stb4(qe, exn); stb4(qe + p.b_cs, eyn); stb4(qe + p.b_2cs, ezn);
        }
        hx_km = hx; hy_km = hy;
        se = se1; sh ^= 1;
    }
}

template <int MODE>
static int launch_volume_one(b200fdtd_ctx* c, int which, int k0, int k1, const RowParams& r, cudaStream_t stream, int kz, int nchunks, int grid_y = 0)
{
    if (k1 <= k0 || r.j1 <= r.j0 || nchunks <= 0) return 0;
    VolParams p;
    p.fin = which == 0 ? cur_volt(c) : cur_curr(c);
    p.f = c->flip ? (which == 0 ? oth_volt(c) : oth_curr(c)) : const_cast<float*>(p.fin);
    p.g = which == 0 ? cur_curr(c) : cur_volt(c);
    p.ca = which == 0 ? c->vv : c->ii;
    p.cb = which == 0 ? c->vi : c->iv;
    p.nx = c->nx; p.ny = c->ny; p.nz = c->nz; p.px = c->px; p.sz = c->sz; p.cs = c->cs;
    p.kz = kz; p.k0 = k0; p.k1 = k1;
    // slab launches may use their own CTA height (variant bits 8-12): a 2-row slab CTA has the register footprint of one
    // plain CTA, so it fits the slot a retiring plain CTA frees when both run concurrently
    const int ty_slab = (c->variant >> 8) & 31;
    const int ty = (MODE != 0 && ty_slab) ? ty_slab : c->ty;
    dim3 block(32, ty);
    dim3 grid(nchunks, grid_y > 0 ? grid_y : (r.j1 - r.j0 + ty - 1) / ty, (k1 - k0 + kz - 1) / kz);
    if (grid.y > 65535 || grid.z > 65535) return fail("grid too large for launch (ny/ty=%u, nz/kz=%u)", grid.y, grid.z);
    const bool cmp = c->cmp_meta[which] != nullptr && (c->variant & 4) == 0;
    p.xv = c->cmp_xv[which]; p.meta = c->cmp_meta[which];
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
static int build_plan(b200fdtd_ctx* c)
{
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
        } else kind[b] = 2;
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
            FusedBox Fb; Fb.y0 = B.y0; Fb.by = B.by; Fb.z0 = B.z0; Fb.bz = B.bz;
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
            FusedBox& Fb = P.fb[P.nfused++];
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
static int launch_volume_xslabs(b200fdtd_ctx* c, int which, int k0, int k1, cudaStream_t stream, int sel = 0)
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
    return launch_volume_one<2>(c, which, a, b, g, stream, kz, P.has_lo + P.has_hi, gy);
}

// volume launches over the fused PML slabs (pre -> update -> post in registers)
// sel: 0 = all slabs, 1 = only the slabs at the low end of their axis (z0 = 0 / y0 = 0), 2 = only those at the high end
static int launch_volume_fused(b200fdtd_ctx* c, int which, int k0, int k1, cudaStream_t stream, int sel = 0)
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
        f.flux = which == 0 ? B.flux_v : B.flux_i;
        f.a = which == 0 ? B.vv : B.ii; f.fo = which == 0 ? B.vvfo : B.iifo; f.fn = which == 0 ? B.vvfn : B.iifn;
        if ((c->variant & 16) == 0) { f.pxv = which == 0 ? B.xv_v : B.xv_i; f.pmeta = which == 0 ? B.meta_v : B.meta_i; if (!f.pxv) f.pmeta = nullptr; }
        const int a = k0 > B.z0 ? k0 : B.z0, b = k1 < B.z0 + B.bz ? k1 : B.z0 + B.bz;
        if (b <= a) continue;
        // thin slabs: march fewer planes per CTA so the launch still fills the machine (>= ~8 CTAs per SM)
        const long long per_chunk = (long long)((c->px + 127) / 128) * ((B.by + slab_ty(c) - 1) / slab_ty(c));
        const int kz = slab_kz(c->kz, b - a, per_chunk);
        cudaStream_t st = stream;
        if (stream == c->side2 && q > 0 && q <= 3 && (c->variant & 1024) == 0) st = c->slab_s[q - 1];
        if (launch_volume_one<1>(c, which, a, b, f, st, kz, (c->px + 127) / 128)) return 1;
    }
    return 0;
}

// all volume launches of one half step restricted to planes [k0,k1), on the main stream
static int launch_volume(b200fdtd_ctx* c, int which, int k0, int k1)
{
    if (!c->volt || !c->vv) return fail("fields/coefficients not bound");
    if (!c->plan.valid) if (build_plan(c)) return 1;
    if (launch_volume_plain(c, which, k0, k1, c->stream)) return 1;
    if (launch_volume_xslabs(c, which, k0, k1, c->stream)) return 1;
    return launch_volume_fused(c, which, k0, k1, c->stream);
}


// fused H->E launch over the plain region: reads the current copies, writes the other copies (the caller flips)
static int launch_he(b200fdtd_ctx* c, cudaStream_t stream, int zc0 = -1, int zc1 = -1)
{
    const VolumePlan& P = c->plan;
    const bool clipped = zc0 >= 0;                           // z-slab ranks: interior planes [zc0, zc1) only
    if (P.nseg != 1) return clipped ? 0 : fail("fused H->E launch without a plain region");
    HeParams p;
    p.ein = cur_volt(c); p.hin = cur_curr(c); p.eout = oth_volt(c); p.hout = oth_curr(c);
    p.vv = c->vv; p.vi = c->vi; p.ii = c->ii; p.iv = c->iv;
    const bool cmp = c->cmp_meta[0] != nullptr && c->cmp_meta[1] != nullptr && (c->variant & 4) == 0;
    p.xv_e = c->cmp_xv[0]; p.meta_e = c->cmp_meta[0]; p.xv_h = c->cmp_xv[1]; p.meta_h = c->cmp_meta[1];
    p.ny = c->ny; p.px = c->px; p.sz = c->sz; p.cs = c->cs;
    p.X0 = P.has_lo ? P.xw0 : 0; p.X1 = P.has_hi ? P.xx1 : c->px; p.XT0 = P.has_hi ? P.xx1 + P.xw1 : c->px;
    p.Y0 = P.ym0; p.Y1 = P.ym1; p.Z0 = P.seg0[0]; p.Z1 = P.seg1[0];
    if (clipped) { if (p.Z0 < zc0) p.Z0 = zc0; if (p.Z1 > zc1) p.Z1 = zc1; if (p.Z1 <= p.Z0) return 0; }
    if (p.Y1 <= p.Y0 || p.Z1 <= p.Z0 || p.X1 <= p.X0) return fail("fused H->E launch over an empty region");
    const int ty = c->he_ty;
    int kz = c->he_kz; if (kz > p.Z1 - p.Z0) kz = p.Z1 - p.Z0;
    { const int n = (p.Z1 - p.Z0 + kz - 1) / kz; kz = (p.Z1 - p.Z0 + n - 1) / n; }      // chunks of equal length
    p.kz = kz;
    p.pf = (c->variant >> 16) & 3;                   // L2 prefetch distance in planes: 0 = default (1), 3 = off
    p.pf = p.pf == 0 ? 1 : (p.pf == 3 ? 0 : p.pf);
    dim3 block(32, ty + 1);
    dim3 grid((c->px - p.X0 + HE_SEG - 1) / HE_SEG, (p.Y1 - p.Y0 + ty - 1) / ty, (p.Z1 - p.Z0 + kz - 1) / kz);
    if (grid.y > 65535 || grid.z > 65535) return fail("grid too large for the fused launch");
    p.b_sz = 4 * p.sz; p.b_cs = 4 * p.cs; p.b_2cs = 8 * p.cs; p.b_sz_cs = 4 * (p.sz + p.cs); p.b_sz_2cs = 4 * (p.sz + 2 * p.cs);
    p.b_row = 4LL * p.px; p.b_row_2cs = 4 * (p.px + 2 * p.cs);
    for (int q = 0; q < 3; ++q) { p.b_pfe[q] = 4 * (2 * p.sz + q * p.cs); p.b_pfh[q] = 4 * (p.sz + q * p.cs); }   // one plane ahead
    p.xv_pitch = 4u * (unsigned)p.px; p.meta_step = 32 * p.ny;
    const bool staged = (c->variant & 256) != 0;
    p.nv_e = c->cmp_nvec[0]; p.nv_h = c->cmp_nvec[1];
    const size_t xs_bytes = (size_t)(p.nv_e + p.nv_h) * 32 * sizeof(float4);
    const bool v2 = cmp && (c->variant & 512) == 0 && xs_bytes <= 96 * 1024;
    const bool v3 = v2 && (c->variant & (1 << 18)) == 0;
#define LAUNCH_HE(TYV) do { \
        if (v3 && (c->variant & (1 << 19)) != 0 && sizeof(He4Smem<TYV>) + xs_bytes <= 220 * 1024) { const size_t sm4 = sizeof(He4Smem<TYV>) + xs_bytes; \
                  CK(cudaFuncSetAttribute(update_he4_kernel<TYV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4)); \
                  update_he4_kernel<TYV><<<grid, block, sm4, stream>>>(p); } \
        else if (v3 && (c->variant & (1 << 21)) == 0 && sizeof(He5Smem<TYV>) + xs_bytes <= 220 * 1024) { const size_t sm5 = sizeof(He5Smem<TYV>) + xs_bytes; \
                  CK(cudaFuncSetAttribute(update_he5_kernel<TYV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm5)); \
                  update_he5_kernel<TYV><<<grid, block, sm5, stream>>>(p); } \
        else if (v3 && sizeof(He3Smem<TYV>) + xs_bytes <= 220 * 1024) { const size_t sm3 = sizeof(He3Smem<TYV>) + xs_bytes; \
                  CK(cudaFuncSetAttribute(update_he3_kernel<TYV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3)); \
                  update_he3_kernel<TYV><<<grid, block, sm3, stream>>>(p); } \
        else if (v2 && 2 * sizeof(float4) * (TYV + 1) * 70 + xs_bytes <= 220 * 1024) { if (xs_bytes > 8 * 1024) CK(cudaFuncSetAttribute(update_he2_kernel<TYV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xs_bytes)); \
                  update_he2_kernel<TYV><<<grid, block, xs_bytes, stream>>>(p); } \
        else if (staged) { \
            const size_t sm = sizeof(HeSmem<TYV>); \
            if (cmp) { CK(cudaFuncSetAttribute(update_he_staged_kernel<TYV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
                       update_he_staged_kernel<TYV, true><<<grid, block, sm, stream>>>(p); } \
            else { CK(cudaFuncSetAttribute(update_he_staged_kernel<TYV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
                   update_he_staged_kernel<TYV, false><<<grid, block, sm, stream>>>(p); } \
        } else if (cmp) update_he_kernel<TYV, true><<<grid, block, 0, stream>>>(p); \
        else update_he_kernel<TYV, false><<<grid, block, 0, stream>>>(p); } while (0)
    switch (ty) {
        case 3: LAUNCH_HE(3); break;
        case 5: LAUNCH_HE(5); break;
        case 7: LAUNCH_HE(7); break;
        case 9: LAUNCH_HE(9); break;
        case 15: LAUNCH_HE(15); break;
        default: return fail("unsupported fused-launch tile height %d", ty);
    }
#undef LAUNCH_HE
    CKL();
    return 0;
}

// can one graph chunk of `steps` steps use the fused H->E launches?  (single slab, every PML box fused into the volume
// launches, a plain region, memory for the second copy of the fields)
static bool he_ready(b200fdtd_ctx* c, int steps)
{
    if ((c->variant & 128) || steps < 2 || !c->plan.valid || c->pml.n > 0 || c->plan.nseg != 1) return false;
    const VolumePlan& P = c->plan;
    if (P.ym1 <= P.ym0 || (P.has_hi ? P.xx1 : c->px) <= (P.has_lo ? P.xw0 : 0)) return false;
    if (!c->alt_volt) {
        const size_t bytes = sizeof(float) * 3 * (size_t)c->cs;
        if (cudaMalloc((void**)&c->alt_volt, bytes) != cudaSuccess) { cudaGetLastError(); c->alt_volt = nullptr; return false; }
        if (cudaMalloc((void**)&c->alt_curr, bytes) != cudaSuccess) { cudaGetLastError(); cudaFree(c->alt_volt); c->alt_volt = c->alt_curr = nullptr; return false; }
        cudaMemsetAsync(c->alt_volt, 0, bytes, c->stream);
        cudaMemsetAsync(c->alt_curr, 0, bytes, c->stream);
        c->alt_owned = true;
    }
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

// ------------------------------------------------------------------------------------
// narrow-band kernels
// ------------------------------------------------------------------------------------
// K5 excitation (Apply2Voltages): volt[idx] += amp * signal[ts - delay]
__global__ void excite_kernel(float* __restrict__ volt, const int64_t* __restrict__ idx,
                              const float* __restrict__ amp, const int* __restrict__ delay,
                              const float* __restrict__ sig, int siglen, int64_t n,
                              const int* __restrict__ d_ts, int ts_off)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int pos = *d_ts + ts_off - delay[e];
    if (pos < 0 || pos >= siglen) return;
    const int64_t q = idx[e];
    volt[q] = __fmaf_rn(amp[e], sig[pos], volt[q]);
}

// K3 Mur (App. A3): pre: tmp = volt[src] - k volt[dst]; post: tmp += k volt[src]; apply: volt[dst] = tmp
__global__ void mur_kernel(float* __restrict__ volt, const int64_t* __restrict__ dst,
                           const int64_t* __restrict__ src, const float* __restrict__ coeff,
                           float* __restrict__ tmp, int64_t n, int phase)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (phase == 0) {
        tmp[e] = __fmaf_rn(-coeff[e], volt[dst[e]], volt[src[e]]);
    } else if (phase == 1) {
        tmp[e] = __fmaf_rn(coeff[e], volt[src[e]], tmp[e]);
    } else {
        volt[dst[e]] = tmp[e];
    }
}

// K4 PML_8 (App. A4): split-flux UPML, pre and post passes over one slab box (the boxes that are not fused
// into the volume kernels: the narrow x-slabs).  blockIdx = (row chunk, z, component); threadIdx = (x, row).
//   pre : h = a*f - fo*flux ; f = flux ; flux = h
//   post: h = flux ; flux = f ; f = h + fn*flux
__global__ void pml_kernel(float* __restrict__ field, const PmlBoxDev B, int which, int post,
                           int px, long long sz, long long cs)
{
    const int y = blockIdx.x * blockDim.y + threadIdx.y;
    if (y >= B.by) return;
    const int z = blockIdx.y, comp = blockIdx.z;
    const long long lrow = (((long long)comp * B.bz + z) * B.by + y) * B.bx;
    const long long grow = comp * cs + (long long)(B.z0 + z + 1) * sz + (long long)(B.y0 + y) * px + B.x0;
    float* __restrict__ flux = which == 0 ? B.flux_v : B.flux_i;
    const float* __restrict__ a = which == 0 ? B.vv : B.ii;
    const float* __restrict__ fo = which == 0 ? B.vvfo : B.iifo;
    const float* __restrict__ fn = which == 0 ? B.vvfn : B.iifn;
    for (int x = threadIdx.x; x < B.bx; x += blockDim.x) {
        const long long l = lrow + x, q = grow + x;
        if (!post) {
            const float fl = flux[l];
            const float h = __fmaf_rn(a[l], field[q], -__fmul_rn(fo[l], fl));
            field[q] = fl;
            flux[l] = h;
        } else {
            const float h = flux[l];
            const float v = field[q];
            flux[l] = v;
            field[q] = __fmaf_rn(fn[l], v, h);
        }
    }
}

// tiny: advance the device step counter
__global__ void ts_add_kernel(int* d_ts, int n) { if (threadIdx.x == 0 && blockIdx.x == 0) *d_ts += n; }

// K6+K7 probes: weighted line/loop sums with a warp-shuffle reduction, time series and running DFT
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ volt, const float* __restrict__ curr,
        const int* __restrict__ kind, const int64_t* __restrict__ off, const int64_t* __restrict__ idx,
        const float* __restrict__ w, int interval, int max_samples, float* __restrict__ series,
        int nfreq, const double* __restrict__ freqs, float* __restrict__ dft, double dt,
        const int* __restrict__ d_ts, int ts_off)
{
    const int p = blockIdx.x;
    const int ts = *d_ts + ts_off;                 // completed steps
    const int s = ts / interval - 1;
    if (s < 0 || s >= max_samples) return;
    const float* fld = kind[p] == 0 ? volt : curr;
    float acc = 0.f;
    for (int64_t e = off[p] + threadIdx.x; e < off[p + 1]; e += blockDim.x) acc = __fmaf_rn(w[e], fld[idx[e]], acc);
    __shared__ float red[4];
    __shared__ float total;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        const float t = (red[0] + red[1]) + (red[2] + red[3]);
        total = t;
        series[(int64_t)p * max_samples + s] = t;
    }
    __syncthreads();
    const float val = total;
    const double tm = (kind[p] == 0 ? (double)ts : (double)ts + 0.5) * dt;
    for (int q = threadIdx.x; q < nfreq; q += blockDim.x) {
        double ph = freqs[q] * tm; ph -= floor(ph);
        double sn, cn; sincospi(2.0 * ph, &sn, &cn);
        float* a = dft + ((int64_t)p * nfreq + q) * 2;
        a[0] = __fmaf_rn(val, (float)cn, a[0]);
        a[1] = __fmaf_rn(-val, (float)sn, a[1]);
    }
}

// K8 NF2FF: running DFT of node-interpolated tangential E/H on the Huygens box faces (App. A6)
struct Nf2ffParams {
    const float* volt; const float* curr;
    int ny, px; long long sz, cs;
    const float* il[3]; const float* idl[3];     // inverse primal / dual edge lengths (z arrays offset by one entry)
    int nfreq; const double* freqs; double dt;
    const int* d_ts; int ts_off;
};
__global__ void __launch_bounds__(128) nf2ff_kernel(const FaceTable* __restrict__ tab, const Nf2ffParams P)
{
    extern __shared__ float tw[];                  // [nfreq][4] = cosE, sinE, cosH, sinH
    const FaceDev& F = tab->f[blockIdx.y];
    const int na = F.a1 - F.a0 + 1, nb = F.b1 - F.b0 + 1;
    const long long nn = (long long)na * nb;
    if ((long long)blockIdx.x * blockDim.x >= nn) return;
    const int ts = *P.d_ts + P.ts_off;
    for (int q = threadIdx.x; q < P.nfreq; q += blockDim.x) {
        double sn, cn;
        double ph = P.freqs[q] * ((double)ts * P.dt); ph -= floor(ph);
        sincospi(2.0 * ph, &sn, &cn); tw[4 * q] = (float)cn; tw[4 * q + 1] = (float)sn;
        ph = P.freqs[q] * (((double)ts + 0.5) * P.dt); ph -= floor(ph);
        sincospi(2.0 * ph, &sn, &cn); tw[4 * q + 2] = (float)cn; tw[4 * q + 3] = (float)sn;
    }
    __syncthreads();
    const long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= nn) return;
    const int n = F.normal, a = (n + 1) % 3, b = (n + 2) % 3;
    const int ia = F.a0 + (int)(node % na), ib = F.b0 + (int)(node / na);
    int co[3]; co[n] = F.plane; co[a] = ia; co[b] = ib;
    const long long st[3] = {1, (long long)P.px, P.sz};
    const long long lin0 = (long long)(co[2] + 1) * P.sz + (long long)co[1] * P.px + co[0];
    const int oa = (a == 2), ob = (b == 2);
    const float* va = P.volt + a * P.cs; const float* vb = P.volt + b * P.cs;
    const float* ca = P.curr + a * P.cs; const float* cb = P.curr + b * P.cs;
    const float Ea = 0.5f * (va[lin0] * P.il[a][ia + oa] + va[lin0 - st[a]] * P.il[a][ia + oa - 1]);
    const float Eb = 0.5f * (vb[lin0] * P.il[b][ib + ob] + vb[lin0 - st[b]] * P.il[b][ib + ob - 1]);
    const float Ha = 0.25f * P.idl[a][ia + oa] *
        ((ca[lin0] + ca[lin0 - st[b]]) + (ca[lin0 - st[n]] + ca[lin0 - st[b] - st[n]]));
    const float Hb = 0.25f * P.idl[b][ib + ob] *
        ((cb[lin0] + cb[lin0 - st[a]]) + (cb[lin0 - st[n]] + cb[lin0 - st[a] - st[n]]));
    const float v[4] = {Ea, Eb, Ha, Hb};
    float2* acc = reinterpret_cast<float2*>(F.acc);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        for (int q = 0; q < P.nfreq; ++q) {
            const float cn = tw[4 * q + (c >= 2 ? 2 : 0)], sn = tw[4 * q + (c >= 2 ? 3 : 1)];
            float2* d = acc + ((long long)c * P.nfreq + q) * nn + node;
            float2 t = *d;
            t.x = __fmaf_rn(v[c], cn, t.x);
            t.y = __fmaf_rn(-v[c], sn, t.y);
            *d = t;
        }
    }
}

// K9 energy: deterministic two-stage reduction of sum(f^2) over the owned planes
__global__ void __launch_bounds__(256) energy_partial_kernel(const float* __restrict__ volt, const float* __restrict__ curr,
        long long sz, long long cs, long long n_owned, double* __restrict__ partials)
{
    // partials[2*block + 0/1] = sum volt^2 / sum curr^2 of this block's grid-stride share
    double sv = 0.0, sc = 0.0;
    const long long n4 = n_owned / 4;              // n_owned = nz*sz, multiple of 4
    for (int c = 0; c < 3; ++c) {
        const float4* v = reinterpret_cast<const float4*>(volt + c * cs + sz);
        const float4* h = reinterpret_cast<const float4*>(curr + c * cs + sz);
        for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
            const float4 a = v[q], b = h[q];
            sv += (double)(a.x * a.x + a.y * a.y) + (double)(a.z * a.z + a.w * a.w);
            sc += (double)(b.x * b.x + b.y * b.y) + (double)(b.z * b.z + b.w * b.w);
        }
    }
    __shared__ double rv[8], rc[8];
    for (int o = 16; o > 0; o >>= 1) { sv += __shfl_down_sync(0xffffffffu, sv, o); sc += __shfl_down_sync(0xffffffffu, sc, o); }
    if ((threadIdx.x & 31) == 0) { rv[threadIdx.x >> 5] = sv; rc[threadIdx.x >> 5] = sc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int i = 0; i < 8; ++i) { a += rv[i]; b += rc[i]; }
        partials[2 * blockIdx.x] = a; partials[2 * blockIdx.x + 1] = b;
    }
}
__global__ void energy_final_kernel(const double* __restrict__ partials, int n, double* __restrict__ out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double a = 0, b = 0;
        for (int i = 0; i < n; ++i) { a += partials[2 * i]; b += partials[2 * i + 1]; }
        out[0] = a; out[1] = b;
    }
}

// K11 far field: N = sum J e^{jk r^.r'}, L = sum M e^{jk r^.r'} projected on theta^/phi^ (App. A6)
__global__ void __launch_bounds__(256) farfield_kernel(long long npts, const float* __restrict__ pos,
        const float* __restrict__ J, const float* __restrict__ M, double k, int ndir,
        const double* __restrict__ theta, const double* __restrict__ phi, float* __restrict__ out)
{
    const int d = blockIdx.x;
    if (d >= ndir) return;
    double st, ct, sp, cp;
    sincos(theta[d], &st, &ct); sincos(phi[d], &sp, &cp);
    const float ux = (float)(k * st * cp), uy = (float)(k * st * sp), uz = (float)(k * ct);
    double a[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = 0.0;
    for (long long q = threadIdx.x; q < npts; q += blockDim.x) {
        const float ph = ux * pos[q] + uy * pos[npts + q] + uz * pos[2 * npts + q];
        float sn, cn; sincosf(ph, &sn, &cn);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float jr = J[(c * npts + q) * 2], ji = J[(c * npts + q) * 2 + 1];
            const float mr = M[(c * npts + q) * 2], mi = M[(c * npts + q) * 2 + 1];
            a[2 * c] += (double)(jr * cn - ji * sn); a[2 * c + 1] += (double)(jr * sn + ji * cn);
            a[6 + 2 * c] += (double)(mr * cn - mi * sn); a[6 + 2 * c + 1] += (double)(mr * sn + mi * cn);
        }
    }
    __shared__ double red[8][12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        double v = a[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s[12];
        for (int i = 0; i < 12; ++i) { s[i] = 0; for (int wv = 0; wv < 8; ++wv) s[i] += red[wv][i]; }
        // theta^ = (ct cp, ct sp, -st), phi^ = (-sp, cp, 0)
        for (int part = 0; part < 2; ++part) {         // 0: N from J, 1: L from M
            const double* v = s + 6 * part;
            for (int ri = 0; ri < 2; ++ri) {
                const double vx = v[ri], vy = v[2 + ri], vz = v[4 + ri];
                out[((long long)d * 4 + 2 * part) * 2 + ri] = (float)(vx * ct * cp + vy * ct * sp - vz * st);
                out[((long long)d * 4 + 2 * part + 1) * 2 + ri] = (float)(-vx * sp + vy * cp);
            }
        }
    }
}

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
static void drop_graph(b200fdtd_ctx* c) { if (c->graph) { cudaGraphExecDestroy(c->graph); c->graph = nullptr; c->graph_steps = 0; } }


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
    cudaFree(c->exc_idx); cudaFree(c->exc_amp); cudaFree(c->exc_delay); cudaFree(c->exc_sig);
    cudaFree(c->mur_dst); cudaFree(c->mur_src); cudaFree(c->mur_coeff); cudaFree(c->mur_tmp);
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
        D.normal = F.normal; D.plane = F.plane; D.a0 = F.a0; D.a1 = F.a1; D.b0 = F.b0; D.b1 = F.b1; D.acc = F.acc;
        const int nn = (F.a1 - F.a0 + 1) * (F.b1 - F.b0 + 1);
        if (nn > max_nodes) max_nodes = nn;
    }
    c->nf_max_nodes = max_nodes; c->nf_nfreq = nfreq; c->nf_interval = interval; c->nf_dt = dt;
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

extern "C" int b200fdtd_get_timestep(b200fdtd_ctx* c, int64_t* ts) { if (!c || !ts) return fail("NULL argument"); *ts = c->ts; return 0; }

extern "C" int b200fdtd_set_timestep(b200fdtd_ctx* c, int64_t ts)
{
    if (!c) return fail("NULL ctx");
    if (ts < 0 || ts > 0x7fffffff) return fail("timestep out of range");
    CK(cudaSetDevice(c->device));
    int v = (int)ts;
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
        c->exc_delay, c->exc_sig, c->exc_siglen, c->n_exc, c->d_ts, ts_off);
    CKL();
    return 0;
}
static int launch_sampling(b200fdtd_ctx* c, int ts_off)
{
    if (c->n_probes > 0) {
        probe_kernel<<<c->n_probes, 128, 0, c->stream>>>(cur_volt(c), cur_curr(c), c->pr_kind, c->pr_off, c->pr_idx, c->pr_w,
            c->interval, c->max_samples, c->pr_series, c->pr_nfreq, c->pr_freqs, c->pr_dft, c->dt, c->d_ts, ts_off);
        CKL();
    }
    if (c->faces.n > 0) {
        Nf2ffParams P;
        P.volt = cur_volt(c); P.curr = cur_curr(c); P.ny = c->ny; P.px = c->px; P.sz = c->sz; P.cs = c->cs;
        for (int a = 0; a < 3; ++a) { P.il[a] = c->inv_len[a]; P.idl[a] = c->inv_dual[a]; }
        P.nfreq = c->nf_nfreq; P.freqs = c->nf_freqs; P.dt = c->nf_dt; P.d_ts = c->d_ts; P.ts_off = ts_off;
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

// E half step with device step offset `off` (host knows ts + off)
static int e_half(b200fdtd_ctx* c, int off)
{
    if (!c->plan.valid) if (build_plan(c)) return 1;
    const bool side = (c->plan.nfused > 0 || c->plan.xedge) && (c->variant & 2) == 0;
    if (launch_mur(c, 0)) return 1;              // Mur sees the true field, before any PML pass swaps in the flux
    if (side) { if (fork_side(c)) return 1; if (launch_volume_xslabs(c, 0, 0, c->nz, c->side)) return 1;
                if (launch_volume_fused(c, 0, 0, c->nz, slab_stream(c))) return 1; }
    if (launch_pml(c, 0, 0)) return 1;
    if (launch_volume_plain(c, 0, 0, c->nz, c->stream)) return 1;
    if (!side) { if (launch_volume_xslabs(c, 0, 0, c->nz, c->stream)) return 1; if (launch_volume_fused(c, 0, 0, c->nz, c->stream)) return 1; }
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
    if (side) { if (fork_side(c)) return 1; if (launch_volume_xslabs(c, 1, 0, c->nz, c->side)) return 1;
                if (launch_volume_fused(c, 1, 0, c->nz, slab_stream(c))) return 1; }
    if (launch_pml(c, 1, 0)) return 1;
    if (launch_volume_plain(c, 1, 0, c->nz, c->stream)) return 1;
    if (!side) { if (launch_volume_xslabs(c, 1, 0, c->nz, c->stream)) return 1; if (launch_volume_fused(c, 1, 0, c->nz, c->stream)) return 1; }
    if (launch_pml(c, 1, 1)) return 1;
    if (side) if (join_side(c)) return 1;
    if (c->flip) c->ccur ^= 1;
    return 0;
}

// H update of step n and E update of step n+1 in one sweep (graph chunks only; see update_he_kernel):
//   PML slabs: H update into the other copy  ->  fused H->E launch over the plain region  ->  PML slabs: E update
static int he_step(b200fdtd_ctx* c, int off)
{
    const bool slabs = c->plan.nfused > 0 || c->plan.xedge;
    const bool side = slabs && (c->variant & 2) == 0;
    if (launch_mur(c, 0)) return 1;              // Mur reads the old E
    c->flip = true;
    int rc = 0;
    const bool overlap = side && (c->variant & (1 << 20)) != 0;      // experiment (measured 1.5 % slower: the fused launch fills every SM)
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
        if (side) { if ((rc = fork_side(c))) break; }
        if ((rc = launch_volume_xslabs(c, 1, 0, c->nz, side ? c->side : c->stream))) break;
        if ((rc = launch_volume_fused(c, 1, 0, c->nz, side ? slab_stream(c) : c->stream))) break;
        if (side) { if ((rc = join_side(c))) break; }
        if ((rc = launch_he(c, c->stream))) break;
        c->ccur ^= 1;                            // H is new from here on
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
        return rc;
    }
    const bool odd = ((n - 1) & 1) != 0;
    c->vcur = c->ccur = 0;
    c->flip = odd; rc = e_half(c, 0); c->flip = false;
    for (int s = 1; s < n && !rc; ++s) rc = he_step(c, s);
    if (!rc) { c->flip = odd; rc = h_half(c); c->flip = false; }
    if (!rc && (c->vcur || c->ccur)) rc = fail("fused span did not return to the bound field arrays");
    c->vcur = c->ccur = 0; c->flip = false;
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

extern "C" int b200fdtd_run(b200fdtd_ctx* c, int64_t nsteps, int use_graph)
{
    if (!c) return fail("NULL ctx");
    if (nsteps < 0) return fail("nsteps < 0");
    if (!c->volt || !c->vv) return fail("fields/coefficients not bound");
    CK(cudaSetDevice(c->device));
    if (!c->plan.valid) if (build_plan(c)) return 1;        // never inside a stream capture
    c->he_fused = false;
    if (!use_graph || c->stream == nullptr) return run_eager(c, nsteps);   // the NULL stream cannot be captured
    const int iv = sample_interval(c);
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
        if (launch_ts_add(c, 1)) return 1;
        c->ts += 1;
        return 0;
    }
    if (phase == 2) {
        const int iv = sample_interval(c);
        if (iv > 0 && (c->ts % iv) == 0) return launch_sampling(c, 0);
        return 0;
    }
    return fail("phase must be 0, 1 or 2");
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
    if (launch_ts_add(c, 1)) return 1;
    c->ts += 1;
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
    const bool side = (c->plan.nfused > 0 || c->plan.xedge) && (c->variant & 2) == 0;    // interior slab launches side by side
    if (part == 0) {
        if (launch_mur(c, 0)) return 1;
        c->flip = true;
        if (side) rc = fork_side(c);
        if (!rc) rc = launch_volume_xslabs(c, 1, 1, nz - 1, side ? c->side : c->stream);
        if (!rc) rc = launch_volume_fused(c, 1, 1, nz - 1, side ? slab_stream(c) : c->stream);
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
        if (launch_ts_add(c, 1)) return 1;
        c->ts += 1;
        return 0;
    }
    c->flip = true;
    if (side) rc = fork_side(c);
    if (!rc) rc = launch_volume_xslabs(c, 0, 1, nz - 1, side ? c->side : c->stream);
    if (!rc) rc = launch_volume_fused(c, 0, 1, nz - 1, side ? slab_stream(c) : c->stream);
    if (!rc) rc = launch_volume(c, 0, 0, 1);
    if (!rc) rc = launch_volume(c, 0, nz - 1, nz);
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
    return 0;
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
    if (which < 2) return launch_volume(c, which, 0, c->nz);
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
    if (!(rows == 0 || rows == 3 || rows == 5 || rows == 7 || rows == 9 || rows == 15)) return fail("rows must be 0, 3, 5, 7, 9 or 15");
    if (planes < 0) return fail("planes must be >= 0");
    if (rows) c->he_ty = rows;
    if (planes) c->he_kz = planes;
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

extern "C" int b200fdtd_farfield(int device, void* stream, int64_t npts, const float* pos, const float* J,
                                  const float* M, double k, int ndir, const double* theta, const double* phi, float* out)
{
    if (npts <= 0 || ndir <= 0 || !pos || !J || !M || !theta || !phi || !out) return fail("bad far-field arguments");
    CK(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    double *d_th = nullptr, *d_ph = nullptr;
    CK(cudaMalloc((void**)&d_th, sizeof(double) * ndir));
    CK(cudaMalloc((void**)&d_ph, sizeof(double) * ndir));
    CK(cudaMemcpyAsync(d_th, theta, sizeof(double) * ndir, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_ph, phi, sizeof(double) * ndir, cudaMemcpyHostToDevice, s));
    farfield_kernel<<<ndir, 256, 0, s>>>(npts, pos, J, M, k, ndir, d_th, d_ph, out);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(s);
    cudaFree(d_th); cudaFree(d_ph);
    if (e != cudaSuccess) return fail("farfield launch failed: %s", cudaGetErrorString(e));
    return 0;
}
