// kernels_fused.cuh - fused H->E launch (temporal blocking): update_he_kernel (plain fusion, any operator) and update_he6_kernel
// (row-compressed operator, whole tiles staged by the TMA engine; the production path)
// Part of libb200fdtd (textually included by b200fdtd.cu; see that file for the data layout and the arithmetic contract).
#pragma once

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
    int X0s;                        // column where the first x-segment starts (X0; 0 when whole-row PML slabs are swept too)
    int Y0, Y1, Z0, Z1;             // owned rows and planes
    int kz;
    int pf;                         // planes of L2 prefetch distance (0 = off)
    // byte offsets as launch constants (update_he6_kernel adds them to one per-thread byte offset: two integer instructions
    // per address instead of a 64-bit index computation)
    long long b_sz, b_cs, b_2cs, b_row, b_row_2cs;
    int meta_step;
    int nv_e, nv_h;                 // x-vectors of the E / H pass (update_he6_kernel keeps its 128-column slice of them in smem)
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


// ---- helpers of the staged kernel ----
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    const int n = valid ? 16 : 0;                            // 0 source bytes: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }


__device__ __forceinline__ float4 ldb4(const char* p) { return *reinterpret_cast<const float4*>(p); }
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


// ------------------------------------------------------------------------------------
// update_he6_kernel: the fused H->E launch of the production path
// ------------------------------------------------------------------------------------
// One CTA = (TY+1) warps = TY owned rows + the halo row below, 124 owned columns + the halo float4 on the left, marching
// up in z.  The row segments a CTA needs are contiguous in global memory (33 float4 of E, 32 of H per row and component):
// one elected lane per warp hands them to the TMA engine as 1-D bulk copies (cp.async.bulk, SASS UBLKCP) that complete on
// one mbarrier per ring slot.  E ring: planes k and k+1 in use + DE planes landing; H ring: plane k in use + one landing.
// A slot is refilled as soon as the CTA barrier in the middle of the iteration has passed (every read of plane k is
// before it), not at the top of the next iteration, so 1.5 / (DE + 0.5) planes are in flight per CTA: the launch is bound
// by bytes in flight per SM, not by the barrier (every tile shape of the previous generation ran at the same speed because
// they all kept one plane in flight per ring).  Out-of-grid parts of the rings are zeroed once and never written again.
//
// PML rows.  Rows inside a whole-row PML slab (z-slabs: all rows of some planes; y-slabs: some rows of the other planes)
// are updated by the same warps with the split-flux pre -> update -> post sequence of update_h/e_kernel MODE 1 on the
// registers they already hold (identical arithmetic per cell).  The halo recompute needs the OLD current flux of cells whose
// owner may already have written the new one, so the current flux is double buffered like the fields (gin -> gout); the
// voltage flux is touched by the owner of a cell only and stays in place.
struct HePmlBox {
    int y0, by, z0, bz;
    const float* gin; float* gout;          // current (H) flux [3][bz][by][px]: read copy / written copy
    float* fv;                              // voltage (E) flux, in place
    const float *ah, *foh, *fnh;            // ii, iifo, iifn
    const float *ae, *foe, *fne;            // vv, vvfo, vvfn
    const float *xvh, *xve; const unsigned char *mh, *me;     // row compression of the slab coefficients (or NULL)
};
struct HePml {
    int zlo, zhi, ym0, ym1;                 // plain planes [zlo,zhi), plain rows [ym0,ym1) of those planes
    int b_zlo, b_zhi, b_ylo, b_yhi;         // box index of each slab (-1: none)
    HePmlBox b[4];
};

// a, fo, fn of component c of a slab row; rec = the row's 48-byte record (9 scales, 9 ids, pad[0] = a slot is streamed in
// full; staged in shared memory one plane ahead) or NULL: stream the full arrays.
__device__ __forceinline__ void pml_coefs(const unsigned char* rec, int c, const float* A, const float* FO, const float* FN,
        const float* __restrict__ xv, long long lofs, int i0, int px, float4& a, float4& fo, float4& fn)
{
    if (rec == nullptr) { a = ld4_nc(A + lofs); fo = ld4_nc(FO + lofs); fn = ld4_nc(FN + lofs); return; }
    const float* sc = reinterpret_cast<const float*>(rec);   // staged in shared memory one plane ahead
    const float s0 = sc[c], s1 = sc[3 + c], s2 = sc[6 + c];
    const unsigned i0_ = rec[36 + c], i1_ = rec[39 + c], i2_ = rec[42 + c], full = rec[45];
    if (full == 0) {
        const char* pl = reinterpret_cast<const char*>(xv + i0); const unsigned pp = 4u * (unsigned)px;
        a = xvg4(pl, i0_, pp, s0); fo = xvg4(pl, i1_, pp, s1); fn = xvg4(pl, i2_, pp, s2);
    } else {
        a = pcoef4(i0_, s0, A + lofs, xv, i0, px); fo = pcoef4(i1_, s1, FO + lofs, xv, i0, px); fn = pcoef4(i2_, s2, FN + lofs, xv, i0, px);
    }
}

// one tiled TMA copy of a [3 components][rows][columns] box of a field into shared memory (4-D tensor map over
// [component][plane][row][column]; out-of-grid elements are zero-filled by the TMA engine)
__device__ __forceinline__ void tma_box4(void* smem, const void* tmap, int c0, int c1, int c2, int c3, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(smem_u32(smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}

template <int TY, int DE>
struct He6Smem {
    struct alignas(128) ESlot { float4 v[3][TY + 2][33]; };  // (a tiled TMA copy wants a 128-byte aligned destination)
    struct alignas(128) HSlot { float4 v[3][TY + 1][32]; };
    ESlot es_[2 + DE];
    HSlot hs_[2];
    float4 xb[2][TY][2][32];                                 // H_new (hz, hx) of rows 0..TY-1, for the row above
    float4 ms[2][TY + 1][4];
    float4 pms[2][TY + 1][6];                                // 48-byte records of a PML slab row: H pass, E pass
    unsigned long long ebar[2 + DE], hbar[2];
    float4 xs[1];                                            // [nv_h + nv_e][32], sized at launch
};

template <int TY, int DE, bool PML, bool TM>
__global__ void __launch_bounds__(32 * (TY + 1), 16 / (TY + 1)) update_he6_kernel(const HeParams p, const __grid_constant__ HePml Q,
        const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmH)
{
    constexpr int NS = 2 + DE;
    extern __shared__ __align__(1024) unsigned char he6_raw[];
    He6Smem<TY, DE>& S = *reinterpret_cast<He6Smem<TY, DE>*>(he6_raw);
    const int lane = threadIdx.x, r = threadIdx.y;
    const int i_seg = p.X0s - 4 + HE_SEG * (int)blockIdx.x;  // column of lane 0
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
        if (!TM) {                                           // (a tiled TMA copy zero-fills what lies outside the grid itself)
            float4* z = &S.es_[0].v[0][0][0];
            const int nz4 = (int)((sizeof(S.es_) + sizeof(S.hs_)) / sizeof(float4));
            for (int q = r * 32 + lane; q < nz4; q += 32 * (TY + 1)) z[q] = zero4();
        }
        if (r == 0 && lane == 0) {
            const unsigned arrivals = TM ? 1u : (unsigned)(TY + 1);      // TM: one thread issues the copies of the whole CTA
            for (int q = 0; q < NS; ++q) mbar_init(&S.ebar[q], arrivals);
            mbar_init(&S.hbar[0], arrivals); mbar_init(&S.hbar[1], arrivals);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zeros are ordered before the TMA writes
    }
    __syncthreads();
    const float4* xsh = S.xs + lane;
    const float4* xse = S.xs + p.nv_h * 32 + lane;
    const bool in_grid = col_ok && j >= 0 && j < p.Y1;
    const bool reg_x = (i0 >= p.X0 && i0 < p.X1) || i0 >= p.XT0;     // plain rows: columns of the launch region
    const bool row_ok = j >= 0 && j < p.Y1;
    // warp-uniform staging plan: in-grid float4 range [c_lo, c_hi) of the 33-wide E row segment (32-wide for H)
    const int c_lo = i_seg < 0 ? (-i_seg + 3) / 4 : 0;
    const int c_hi_e = min(33, (p.px - i_seg) / 4), c_hi_h = min(32, (p.px - i_seg) / 4);
    const bool e_row = j >= 0 && j < p.ny && j <= p.Y1 && c_hi_e > c_lo;         // this warp's own row of E
    const bool e_top = r == TY && j + 1 >= 0 && j + 1 < p.ny && j + 1 <= p.Y1 && c_hi_e > c_lo;   // +1 row (top warp)
    const bool h_row = j >= p.Y0 && j < p.Y1 && c_hi_h > c_lo;                    // H_old where H_new is computed (j >= Y0 >= 0)
    const unsigned nb_e = (unsigned)(c_hi_e - c_lo) * 16u, nb_h = (unsigned)(c_hi_h - c_lo) * 16u;
    const unsigned bytes_e = (e_row ? 3u * nb_e : 0u) + (e_top ? 2u * nb_e : 0u), bytes_h = h_row ? 3u * nb_h : 0u;
    const int kfirst = kbeg > p.Z0 ? kbeg - 1 : kbeg;
    // one byte offset per thread, advancing by a plane per iteration; every address is a launch-constant base plus it
    long long boff = 4 * ((long long)(kfirst + 1) * p.sz + (long long)j * p.px + i0);
    const long long soff = 16LL * (c_lo - lane);             // lane 0: from its own cell to the start of the in-grid part of the row segment
    const char* const ein_b = reinterpret_cast<const char*>(p.ein); const char* const hin_b = reinterpret_cast<const char*>(p.hin);
    char* const eout_b = reinterpret_cast<char*>(p.eout); char* const hout_b = reinterpret_cast<char*>(p.hout);
    const char* mrec = (lane < 2 ? reinterpret_cast<const char*>(p.meta_h) : reinterpret_cast<const char*>(p.meta_e))
                       + ((long long)(kfirst + 1) * p.ny + j) * 32 + (lane & 1) * 16;
    const char* safe = reinterpret_cast<const char*>(p.ein);

    int tcur = 0;                                            // iteration boff belongs to
    // lane 0 of every warp: hand this warp's rows of E_old(plane kfirst + d) to the TMA engine, into ring slot `slot`
    const int j_first = p.Y0 - 1 + TY * (int)blockIdx.y;     // row of warp 0
    auto stage_e = [&](int slot, int d) {
        if (TM) {      // one tiled copy: [3][TY+2 rows][132 columns] of plane kfirst + d
            if (r == 0 && lane == 0) {
                mbar_arrive_expect(&S.ebar[slot], 3u * (TY + 2) * 132u * 4u);
                tma_box4(&S.es_[slot], &tmE, i_seg, j_first, kfirst + 1 + d, 0, &S.ebar[slot]);
            }
            return;
        }
        if (lane != 0) return;
        unsigned long long* bar = &S.ebar[slot];
        mbar_arrive_expect(bar, bytes_e);
        const char* g = ein_b + (boff + soff) + (long long)(d - tcur) * p.b_sz;
        if (e_row) {
            bulk_g2s(&S.es_[slot].v[0][r][c_lo], g, nb_e, bar);
            bulk_g2s(&S.es_[slot].v[1][r][c_lo], g + p.b_cs, nb_e, bar);
            bulk_g2s(&S.es_[slot].v[2][r][c_lo], g + p.b_2cs, nb_e, bar);
        }
        if (e_top) {
            bulk_g2s(&S.es_[slot].v[0][TY + 1][c_lo], g + p.b_row, nb_e, bar);
            bulk_g2s(&S.es_[slot].v[2][TY + 1][c_lo], g + p.b_row_2cs, nb_e, bar);
        }
    };
    auto stage_h = [&](int slot, int d) {
        if (TM) {
            if (r == 0 && lane == 0) {
                mbar_arrive_expect(&S.hbar[slot], 3u * (TY + 1) * 128u * 4u);
                tma_box4(&S.hs_[slot], &tmH, i_seg, j_first, kfirst + 1 + d, 0, &S.hbar[slot]);
            }
            return;
        }
        if (lane != 0) return;
        unsigned long long* bar = &S.hbar[slot];
        mbar_arrive_expect(bar, bytes_h);
        if (h_row) {
            const char* g = hin_b + (boff + soff) + (long long)(d - tcur) * p.b_sz;
            bulk_g2s(&S.hs_[slot].v[0][r][c_lo], g, nb_h, bar);
            bulk_g2s(&S.hs_[slot].v[1][r][c_lo], g + p.b_cs, nb_h, bar);
            bulk_g2s(&S.hs_[slot].v[2][r][c_lo], g + p.b_2cs, nb_h, bar);
        }
    };
    // which whole-row PML slab row (j, k) lies in (-1: a plain row); warp-uniform
    auto pml_box = [&](int k) -> int {
        if (!PML) return -1;
        if (k < Q.zlo) return Q.b_zlo;
        if (k >= Q.zhi) return Q.b_zhi;
        if (j < Q.ym0) return Q.b_ylo;
        if (j >= Q.ym1) return Q.b_yhi;
        return -1;
    };

    // lanes 4..9: the two 48-byte slab row records of plane k (if (j, k) is a PML row) into pms[buf]
    auto stage_pml_rec = [&](int buf, int k) {
        if (!PML) return;
        const int pbk = pml_box(k);
        if (pbk < 0 || !row_ok || lane < 4 || lane >= 10) return;
        const HePmlBox& Bk = Q.b[pbk];
        const unsigned char* m = lane < 7 ? Bk.mh : Bk.me;
        if (m == nullptr) return;
        const long long lrow = (long long)(k - Bk.z0) * Bk.by + (j - Bk.y0);
        cp_async16(&S.pms[buf][r][lane - 4], m + lrow * 48 + ((lane - 4) % 3) * 16, true);
    };
    const int nplanes = kend - kfirst;                       // iterations; E planes kfirst .. kend are needed (nplanes + 1)
#pragma unroll
    for (int q = 0; q < NS; ++q) if (q <= nplanes) stage_e(q, q);
    stage_h(0, 0);
    if (nplanes > 1) stage_h(1, 1);
    if (lane < 4) cp_async16(&S.ms[0][r][lane], row_ok ? mrec : safe, row_ok);
    stage_pml_rec(0, kfirst);
    cp_async_commit();
    float4 hx_km = zero4(), hy_km = zero4();
    if (kfirst == kbeg && in_grid && lane >= 1 && r >= 1) { hx_km = ldb4(hout_b + boff - p.b_sz); hy_km = ldb4(hout_b + boff - p.b_sz + p.b_cs); }
    cp_async_wait<0>();
    __syncwarp();
    mbar_wait(&S.ebar[0], 0u);

    int se = 0;                                              // ring slot of plane k
    for (int t = 0; t < nplanes; ++t, boff += p.b_sz, mrec += p.meta_step) {
        const int k = kfirst + t;
        tcur = t;
        char* const qh = hout_b + boff;
        const bool pro = k < kbeg;
        const int se1 = se + 1 == NS ? 0 : se + 1, sh = t & 1, mb = t & 1;
        // row type of this iteration (warp-uniform) and what each lane does with its cells
        const int pb = pml_box(k);
        const bool regc = PML && pb >= 0 ? true : reg_x;
        const bool ext = in_grid && (!regc || j < p.Y0);     // H_new was written by a slab launch: read it
        const bool calc = in_grid && !ext;                   // H_new is computed here (owned cells and halo cells)
        const bool own = calc && lane >= 1 && r >= 1;        // ... and stored, together with E_new
        if (t + 1 < nplanes && lane < 4) cp_async16(&S.ms[mb ^ 1][r][lane], row_ok ? mrec + p.meta_step : safe, row_ok);
        if (t + 1 < nplanes) stage_pml_rec(mb ^ 1, k + 1);
        cp_async_commit();
        // PML row: slab-local offsets, old flux, row records (issued before the wait on the staged planes)
        long long lofs = 0, lcs = 0;
        float4 g0 = zero4(), g1 = zero4(), g2 = zero4();
        const unsigned char *rech = nullptr, *rece = nullptr;
        const HePmlBox* B = nullptr;
        if (PML && pb >= 0) {
            B = &Q.b[pb];
            const long long lrow = (long long)(k - B->z0) * B->by + (j - B->y0);
            lofs = lrow * p.px + i0; lcs = (long long)B->bz * B->by * p.px;
            if (B->mh) rech = reinterpret_cast<const unsigned char*>(&S.pms[mb][r][0]);
            if (B->me) rece = reinterpret_cast<const unsigned char*>(&S.pms[mb][r][3]);
            if (calc) { g0 = __ldcs(reinterpret_cast<const float4*>(B->gin + lofs)); g1 = __ldcs(reinterpret_cast<const float4*>(B->gin + lcs + lofs));
                        g2 = __ldcs(reinterpret_cast<const float4*>(B->gin + 2 * lcs + lofs)); }
        }
        if (PML) {                                           // pull the slab rows of plane k+2 into L2
            const int pb2 = k + 2 < kend ? pml_box(k + 2) : -1;
            if (pb2 >= 0 && in_grid) {
                const HePmlBox& B2 = Q.b[pb2];
                const long long l2 = ((long long)(k + 2 - B2.z0) * B2.by + (j - B2.y0)) * p.px + i0, c2 = (long long)B2.bz * B2.by * p.px;
                prefetch_l2(B2.gin + l2); prefetch_l2(B2.gin + c2 + l2); prefetch_l2(B2.gin + 2 * c2 + l2);
                prefetch_l2(B2.fv + l2); prefetch_l2(B2.fv + c2 + l2); prefetch_l2(B2.fv + 2 * c2 + l2);
            }
        }
        float4 hx = zero4(), hy = zero4(), hz = zero4();
        // H_new of cells a slab launch owns, where an owned cell of this tile needs it (the lane left of an owned lane, the
        // row below the first owned row).  Loaded into registers of their own and merged after the H phase: the loads (L2 or
        // DRAM round trips) then overlap the whole H phase instead of blocking the first instruction that reuses a register.
        const int own_right = __shfl_down_sync(0xffffffffu, own ? 1 : 0, 1);      // every lane takes part (no short-circuit around it)
        // (with PML rows inside, the row or plane above may own columns this row does not: then every such value is needed)
        const bool ext_need = ext && (PML || j < p.Y0 || own_right != 0);
        float4 hxe = zero4(), hye = zero4(), hze = zero4();
        if (ext_need) { hxe = ldb4(qh); hye = ldb4(qh + p.b_cs); hze = ldb4(qh + p.b_2cs); }
        mbar_wait(&S.ebar[se1], (unsigned)((t + 1) / NS) & 1u);  // plane k+1 has landed (plane k was waited for one iteration ago)
        mbar_wait(&S.hbar[sh], (unsigned)(t >> 1) & 1u);
        float4 ax = zero4(), ay = zero4(), az = zero4(), bx = zero4(), by = zero4(), bz = zero4();
        const float4 ex = S.es_[se].v[0][r][lane], ey = S.es_[se].v[1][r][lane], ez = S.es_[se].v[2][r][lane];
        float ez_r = __shfl_down_sync(0xffffffffu, ez.x, 1);
        float ey_r = __shfl_down_sync(0xffffffffu, ey.x, 1);
        if (lane == 31) { ez_r = S.es_[se].v[2][r][32].x; ey_r = S.es_[se].v[1][r][32].x; }
        if (calc) {
            hx = S.hs_[sh].v[0][r][lane]; hy = S.hs_[sh].v[1][r][lane]; hz = S.hs_[sh].v[2][r][lane];
            row_coefs(S.ms[mb][r][0], S.ms[mb][r][1], xsh, p.ii, p.iv, p.xv_h,
                      boff >> 2, p.cs, i0, p.px, ax, ay, az, bx, by, bz);
            const float4 ex1 = S.es_[se1].v[0][r][lane], ey1 = S.es_[se1].v[1][r][lane];
            const float4 ex_jp = S.es_[se].v[0][r + 1][lane], ez_jp = S.es_[se].v[2][r + 1][lane];
            const float4 ez_ip = make_float4(ez.y, ez.z, ez.w, ez_r);
            const float4 ey_ip = make_float4(ey.y, ey.z, ey.w, ey_r);
            if (PML && pb >= 0) {
                float4 a_, fo_, fn_, h_;
                pml_coefs(rech, 0, B->ah, B->foh, B->fnh, B->xvh, lofs, i0, p.px, a_, fo_, fn_);
                h_ = pml_pre4(a_, fo_, g0, hx); g0 = upd4(ax, g0, bx, ez, ez_jp, ey, ey1); hx = pml_post4(fn_, g0, h_);
                pml_coefs(rech, 1, B->ah, B->foh, B->fnh, B->xvh, lcs + lofs, i0, p.px, a_, fo_, fn_);
                h_ = pml_pre4(a_, fo_, g1, hy); g1 = upd4(ay, g1, by, ex, ex1, ez, ez_ip); hy = pml_post4(fn_, g1, h_);
                pml_coefs(rech, 2, B->ah, B->foh, B->fnh, B->xvh, 2 * lcs + lofs, i0, p.px, a_, fo_, fn_);
                h_ = pml_pre4(a_, fo_, g2, hz); g2 = upd4(az, g2, bz, ey, ey_ip, ex, ex_jp); hz = pml_post4(fn_, g2, h_);
            } else {
                hx = upd4(ax, hx, bx, ez, ez_jp, ey, ey1);
                hy = upd4(ay, hy, by, ex, ex1, ez, ez_ip);
                hz = upd4(az, hz, bz, ey, ey_ip, ex, ex_jp);
            }
        }
        if (ext) { hx = hxe; hy = hye; hz = hze; }
        float4 f0 = zero4(), f1 = zero4(), f2 = zero4();
        if (own && !pro) {
            stb4(qh, hx); stb4(qh + p.b_cs, hy); stb4(qh + p.b_2cs, hz);
            if (PML && pb >= 0) {
                float* go = B->gout + lofs;
                __stcs(reinterpret_cast<float4*>(go), g0); __stcs(reinterpret_cast<float4*>(go + lcs), g1); __stcs(reinterpret_cast<float4*>(go + 2 * lcs), g2);
                const float* fp = B->fv + lofs;              // old voltage flux: in flight while the CTA barrier is crossed
                f0 = __ldcs(reinterpret_cast<const float4*>(fp)); f1 = __ldcs(reinterpret_cast<const float4*>(fp + lcs));
                f2 = __ldcs(reinterpret_cast<const float4*>(fp + 2 * lcs));
            }
            if (!PML) row_coefs(S.ms[mb][r][2], S.ms[mb][r][3], xse, p.vv, p.vi, p.xv_e,
                                boff >> 2, p.cs, i0, p.px, ax, ay, az, bx, by, bz);
        }
        if (r < TY) { S.xb[mb][r][0][lane] = hz; S.xb[mb][r][1][lane] = hx; }
        cp_async_wait<0>();                                  // next plane's records (this warp's own copies)
        __syncthreads();                                     // H_new rows visible; every read of plane k (E and H rings) is done
        // refill the slots of plane k: E_old(k + NS), H_old(k + 2)
        if (t + NS <= nplanes) stage_e(se, t + NS);
        if (t + 2 < nplanes) stage_h(sh, t + 2);
        const float hz_l = __shfl_up_sync(0xffffffffu, hz.w, 1);
        const float hy_l = __shfl_up_sync(0xffffffffu, hy.w, 1);
        if (own && !pro) {
            if (PML) row_coefs(S.ms[mb][r][2], S.ms[mb][r][3], xse, p.vv, p.vi, p.xv_e,
                               boff >> 2, p.cs, i0, p.px, ax, ay, az, bx, by, bz);     // (after the barrier: fewer values live across it)
            const float4 hz_jm = S.xb[mb][r - 1][0][lane], hx_jm = S.xb[mb][r - 1][1][lane];
            const float4 hz_im = make_float4(hz_l, hz.x, hz.y, hz.z);
            const float4 hy_im = make_float4(hy_l, hy.x, hy.y, hy.z);
            float4 exn, eyn, ezn;
            if (PML && pb >= 0) {
                float4 a_, fo_, fn_, h_;
                float* fp = B->fv + lofs;                    // the voltage flux of a cell is touched by its owner only: in place
                pml_coefs(rece, 0, B->ae, B->foe, B->fne, B->xve, lofs, i0, p.px, a_, fo_, fn_);
                h_ = pml_pre4(a_, fo_, f0, ex); f0 = upd4(ax, f0, bx, hz, hz_jm, hy, hy_km); exn = pml_post4(fn_, f0, h_);
                pml_coefs(rece, 1, B->ae, B->foe, B->fne, B->xve, lcs + lofs, i0, p.px, a_, fo_, fn_);
                h_ = pml_pre4(a_, fo_, f1, ey); f1 = upd4(ay, f1, by, hx, hx_km, hz, hz_im); eyn = pml_post4(fn_, f1, h_);
                pml_coefs(rece, 2, B->ae, B->foe, B->fne, B->xve, 2 * lcs + lofs, i0, p.px, a_, fo_, fn_);
                h_ = pml_pre4(a_, fo_, f2, ez); f2 = upd4(az, f2, bz, hy, hy_im, hx, hx_jm); ezn = pml_post4(fn_, f2, h_);
                __stcs(reinterpret_cast<float4*>(fp), f0); __stcs(reinterpret_cast<float4*>(fp + lcs), f1); __stcs(reinterpret_cast<float4*>(fp + 2 * lcs), f2);
            } else {
                exn = upd4(ax, ex, bx, hz, hz_jm, hy, hy_km);
                eyn = upd4(ay, ey, by, hx, hx_km, hz, hz_im);
                ezn = upd4(az, ez, bz, hy, hy_im, hx, hx_jm);
            }
            char* const qe = eout_b + boff;
            stb4(qe, exn); stb4(qe + p.b_cs, eyn); stb4(qe + p.b_2cs, ezn);
        }
        hx_km = hx; hy_km = hy;
        se = se1;
    }
}
