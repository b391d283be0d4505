// kernels_volume.cuh - K1/K2: Yee E and H volume updates (plain rows, fused PML rows, narrow x-slabs), row compression
// Part of libb200fdtd (textually included by b200fdtd.cu; see that file for the data layout and the arithmetic contract).
#pragma once

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
    int kz;                         // planes between the starts of consecutive z-chunks
    int kspan;                      // planes a chunk marches (0 = kz); kspan < kz: a launch over scattered planes (the two boundary
                                    // planes of a z-slab rank in one launch: kz = nz - 1, kspan = 1)
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
    const float* flux;              // fused slab arrays [3][bz][by][px] (PML launches only): old flux is read here
    float* flux_out;                // ... and the new flux written here (== flux, or the other copy when the current flux ping-pongs)
    const float* a; const float* fo; const float* fn;
    const float* pxv; const unsigned char* pmeta;   // row compression of a/fo/fn (48-byte records per slab row) or NULL
    int y0, z0, by, bz;
    // narrow x-slabs (columns [0,xw0) and [xx1,xx1+xw1), multiples of 4): the plain launch (MODE 0) does not store
    // these columns; a narrow launch (MODE 2, bix = slab) owns them: a warp covers xs float4 columns of 32/xs
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
__device__ __forceinline__ void update_e_body(const VolParams& p, const RowParams& r, const unsigned bix, const unsigned biy, const unsigned biz)
{
    constexpr bool PML = MODE == 1;
    constexpr int KSTEP = 1;                                // the E march goes up in z
    const int lane = threadIdx.x;
    int i0, j;
    bool act;
    // narrow-slab launch state (MODE 2)
    const bool hi_slab = MODE == 2 && (bix == 1 || r.xw0 == 0);
    const int xw = hi_slab ? r.xw1 : r.xw0, xx0 = hi_slab ? r.xx1 : 0;
    if (MODE == 2) {
        const int xs = hi_slab ? r.xs1 : r.xs0;             // float4 slots per row, 32/xs rows per warp (12 columns: 10 rows, 2 idle lanes)
        const int c4 = lane % xs, rw = lane / xs;
        j = r.j0 + (biy * TY + threadIdx.y) * (32 / xs) + rw;
        i0 = xx0 + c4 * 4;
        if (rw >= 32 / xs || j >= r.j1 || c4 * 4 >= xw) return;   // per lane (no warp collectives in this mode)
        act = true;
    } else {
        i0 = (bix * 32 + lane) * 4;
        j = r.j0 + biy * TY + threadIdx.y;
        if (j >= r.j1) return;                              // warp-uniform
        if (MODE == 0) {
            if ((j >= r.sj0a && j < r.sj1a) || (j >= r.sj0b && j < r.sj1b)) return;
        }
        act = i0 < p.px;
    }
    // plain launch: columns owned by a narrow-slab launch are computed but not stored
    const bool own = MODE != 0 || !(i0 < r.xw0 || (i0 >= r.xx1 && i0 < r.xx1 + r.xw1));
    const int kbeg = p.k0 + biz * p.kz;
    const int kend = min(kbeg + (p.kspan ? p.kspan : p.kz), p.k1);
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
            if (act) { st4(r.flux_out + lb, fl0); st4(r.flux_out + lcs + lb, fl1); st4(r.flux_out + 2 * lcs + lb, fl2); }
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
__device__ __forceinline__ void update_h_body(const VolParams& p, const RowParams& r, const unsigned bix, const unsigned biy, const unsigned biz)
{
    constexpr bool PML = MODE == 1;
    constexpr int KSTEP = -1;                               // the H march goes down in z
    const int lane = threadIdx.x;
    int i0, j;
    bool act;
    // narrow-slab launch state (MODE 2)
    const bool hi_slab = MODE == 2 && (bix == 1 || r.xw0 == 0);
    const int xw = hi_slab ? r.xw1 : r.xw0, xx0 = hi_slab ? r.xx1 : 0;
    if (MODE == 2) {
        const int xs = hi_slab ? r.xs1 : r.xs0;             // float4 slots per row, 32/xs rows per warp (12 columns: 10 rows, 2 idle lanes)
        const int c4 = lane % xs, rw = lane / xs;
        j = r.j0 + (biy * TY + threadIdx.y) * (32 / xs) + rw;
        i0 = xx0 + c4 * 4;
        if (rw >= 32 / xs || j >= r.j1 || c4 * 4 >= xw) return;   // per lane (no warp collectives in this mode)
        act = true;
    } else {
        i0 = (bix * 32 + lane) * 4;
        j = r.j0 + biy * TY + threadIdx.y;
        if (j >= r.j1) return;                              // warp-uniform
        if (MODE == 0) {
            if ((j >= r.sj0a && j < r.sj1a) || (j >= r.sj0b && j < r.sj1b)) return;
        }
        act = i0 < p.px;
    }
    // plain launch: columns owned by a narrow-slab launch are computed but not stored
    const bool own = MODE != 0 || !(i0 < r.xw0 || (i0 >= r.xx1 && i0 < r.xx1 + r.xw1));
    const int kbeg = p.k0 + biz * p.kz;
    const int kend = min(kbeg + (p.kspan ? p.kspan : p.kz), p.k1);
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
            if (act) { st4(r.flux_out + lb, fl0); st4(r.flux_out + lcs + lb, fl1); st4(r.flux_out + 2 * lcs + lb, fl2); }
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

template <int TY, int MODE, bool CMP>
__global__ void __launch_bounds__(32 * TY, MODE ? (16 / TY > 0 ? 16 / TY : 1) : (24 / TY > 0 ? 24 / TY : 1)) update_e_kernel(const VolParams p, const RowParams r)
{
    update_e_body<TY, MODE, CMP>(p, r, blockIdx.x, blockIdx.y, blockIdx.z);
}
template <int TY, int MODE, bool CMP>
__global__ void __launch_bounds__(32 * TY, MODE ? (16 / TY > 0 ? 16 / TY : 1) : (24 / TY > 0 ? 24 / TY : 1)) update_h_kernel(const VolParams p, const RowParams r)
{
    update_h_body<TY, MODE, CMP>(p, r, blockIdx.x, blockIdx.y, blockIdx.z);
}

// All PML slab launches of one half step in ONE launch: the CTAs of up to six slabs (two z-slabs, two y-slabs, the narrow
// x-slab pair) are numbered consecutively; a CTA finds its slab and runs that slab's update exactly as its own launch
// would.  One grid that fills the machine instead of five small ones on five streams, and 2 instead of 10 slab launches
// per step (what matters most on small grids, where a step is launch-bound).
struct SlabEntry { RowParams r; int mode; int gx, gy, gz; int kz, k0, k1; int cta0; };
#define MAX_SLABS 12               // six slabs, each possibly once per boundary plane of a z-slab rank
struct SlabSet { int n; SlabEntry e[MAX_SLABS]; };
template <int WHICH, int TY, bool CMP>
__global__ void __launch_bounds__(32 * TY, (16 / TY > 0 ? 16 / TY : 1)) update_slabs_kernel(const VolParams p0, const __grid_constant__ SlabSet T)
{
    int q = 0;
#pragma unroll
    for (int u = 1; u < MAX_SLABS; ++u) if (u < T.n && (int)blockIdx.x >= T.e[u].cta0) q = u;
    const SlabEntry& E = T.e[q];
    const unsigned local = blockIdx.x - E.cta0;
    const unsigned bix = local % E.gx, biy = (local / E.gx) % E.gy, biz = local / (E.gx * E.gy);
    VolParams p = p0;
    p.kz = E.kz; p.kspan = 0; p.k0 = E.k0; p.k1 = E.k1;
    if (E.mode == 1) { if (WHICH == 0) update_e_body<TY, 1, CMP>(p, E.r, bix, biy, biz); else update_h_body<TY, 1, CMP>(p, E.r, bix, biy, biz); }
    else             { if (WHICH == 0) update_e_body<TY, 2, CMP>(p, E.r, bix, biy, biz); else update_h_body<TY, 2, CMP>(p, E.r, bix, biy, biz); }
}
