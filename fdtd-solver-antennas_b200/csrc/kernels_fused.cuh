// kernels_fused.cuh - fused H->E launch (temporal blocking): every generation from the plain fusion to the TMA-staged default
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

