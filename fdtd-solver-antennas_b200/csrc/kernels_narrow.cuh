// kernels_narrow.cuh - K3..K11: excitation, Mur, separate PML passes, probes, NF2FF running DFT, energy, far field
// Part of libb200fdtd (textually included by b200fdtd.cu; see that file for the data layout and the arithmetic contract).
#pragma once

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

// K3 with implicit indices.  The edges of a Mur face come in long arithmetic runs (a row of a z- or y-face: stride 1; a
// column of an x-face: stride px), so the lists are stored as segments {first dst, first src, strides, count, first
// entry}: 40 bytes per run instead of 16 bytes of int64 indices per edge.  One warp per segment; same arithmetic per edge.
struct MurSeg { long long dst0, src0; int sd, ss; int count; int pad; long long e0; };
#define MUR_RUN_MAX 128            // 4 edges per lane: all loads of a run are in flight at once
__global__ void __launch_bounds__(128) mur_seg_kernel(float* __restrict__ volt, const MurSeg* __restrict__ segs, int nsegs,
                                                      const float* __restrict__ coeff, float* __restrict__ tmp, int phase)
{
    const int w = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= nsegs) return;
    const MurSeg S = segs[w];
    float a[MUR_RUN_MAX / 32], b[MUR_RUN_MAX / 32], k[MUR_RUN_MAX / 32];
#pragma unroll
    for (int u = 0; u < MUR_RUN_MAX / 32; ++u) {            // loads first ...
        const int q = lane + 32 * u;
        a[u] = b[u] = k[u] = 0.f;
        if (q < S.count) {
            const long long e = S.e0 + q, d = S.dst0 + (long long)q * S.sd, sidx = S.src0 + (long long)q * S.ss;
            if (phase == 0) { k[u] = coeff[e]; a[u] = volt[d]; b[u] = volt[sidx]; }
            else if (phase == 1) { k[u] = coeff[e]; a[u] = volt[sidx]; b[u] = tmp[e]; }
            else a[u] = tmp[e];
        }
    }
#pragma unroll
    for (int u = 0; u < MUR_RUN_MAX / 32; ++u) {            // ... then the same arithmetic per edge as mur_kernel
        const int q = lane + 32 * u;
        if (q < S.count) {
            const long long e = S.e0 + q, d = S.dst0 + (long long)q * S.sd;
            if (phase == 0) tmp[e] = __fmaf_rn(-k[u], a[u], b[u]);
            else if (phase == 1) tmp[e] = __fmaf_rn(k[u], a[u], b[u]);
            else volt[d] = a[u];
        }
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
    int interval, td_max;                        // time-domain store: sample s = ts/interval - 1 goes to td[s] if s < td_max
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
    if (F.td != nullptr) {                         // what openEMS writes to nf2ff_E/H_n.h5: the samples themselves
        const int s = ts / P.interval - 1;
        if (s >= 0 && s < P.td_max) {
            float* d = F.td + (long long)s * 4 * nn + node;
#pragma unroll
            for (int c = 0; c < 4; ++c) __stcs(d + c * nn, v[c]);
        }
    }
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

// K8b NF2FF at any frequency after the run: DFT of the stored face samples (openEMS's nf2ff reads its time-domain HDF5 dumps
// and does the same).  One thread per (component, node); twiddles of a chunk of samples are formed once per block in fp64
// and shared; accumulation in fp64.  td [ns][4][nn] f32, out [4][nfreq][nn][2] f32 (the layout of the running-DFT accumulators).
#define TD_CHUNK 128
__global__ void __launch_bounds__(256) nf2ff_td_dft_kernel(const float* __restrict__ td, long long nn, int ns, int interval, double dt,
        int nfreq, const double* __restrict__ freqs, float* __restrict__ out)
{
    __shared__ float2 tw[2][TD_CHUNK];             // [E | H time stamps][sample] = (cos, sin)
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over 4*nn
    const bool ok = q < 4 * nn;
    const int c = ok ? (int)(q / nn) : 0;
    const int hs = c >= 2 ? 1 : 0;
    for (int f = 0; f < nfreq; ++f) {
        double re = 0.0, im = 0.0;
        for (int s0 = 0; s0 < ns; s0 += TD_CHUNK) {
            __syncthreads();
            for (int t = threadIdx.x; t < 2 * TD_CHUNK; t += blockDim.x) {
                const int s = s0 + (t % TD_CHUNK), h = t / TD_CHUNK;
                const double tm = ((double)(s + 1) * interval + (h ? 0.5 : 0.0)) * dt;
                double ph = freqs[f] * tm; ph -= floor(ph);
                double sn, cn; sincospi(2.0 * ph, &sn, &cn);
                tw[h][t % TD_CHUNK] = make_float2((float)cn, (float)sn);
            }
            __syncthreads();
            if (ok) {
                const int m = min(TD_CHUNK, ns - s0);
                const float* src = td + (long long)s0 * 4 * nn + q;
                for (int s = 0; s < m; ++s) {
                    const float v = __ldcs(src + (long long)s * 4 * nn);
                    const float2 w = tw[hs][s];
                    re += (double)(v * w.x); im -= (double)(v * w.y);
                }
            }
        }
        if (ok) reinterpret_cast<float2*>(out)[((long long)c * nfreq + f) * nn + (q - (long long)c * nn)] = make_float2((float)re, (float)im);
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
    // one warp, fixed order: lane l adds partials l, l+32, ...; then a shuffle tree (deterministic)
    double a = 0, b = 0;
    for (int i = threadIdx.x; i < n; i += 32) { a += partials[2 * i]; b += partials[2 * i + 1]; }
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}

// K11a equivalent currents of one Huygens face from its spectra: J = n x H, M = -n x E (times dA and the DFT scale), node
// positions relative to the phase centre, and the face's share of Prad = 1/2 Re sum (E x H*) . n dA (fp64, one partial per
// block, added on the host in a fixed order).  acc = spectra of the chosen frequency: component c at acc + c * cstride.
struct SrcFace { int normal, side, na, nb; double coord; const float* acc; long long cstride; const double *xa, *xb, *wa, *wb; long long off; };
__global__ void __launch_bounds__(256) nf2ff_sources_kernel(const SrcFace F, double scale, double cx, double cy, double cz, long long npts,
        float* __restrict__ pos, float* __restrict__ J, float* __restrict__ M, double* __restrict__ prad_partial)
{
    const long long nn = (long long)F.na * F.nb;
    const long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double pr = 0.0;
    if (node < nn) {
        const int ia = (int)(node % F.na), ib = (int)(node / F.na);
        const int n = F.normal, a = (n + 1) % 3, b = (n + 2) % 3;
        const double s = F.side == 1 ? 1.0 : -1.0;
        const double dA = F.wb[ib] * F.wa[ia];
        const float w = (float)(dA * scale), sw = (float)s * w;
        const float2* A = reinterpret_cast<const float2*>(F.acc);
        const float2 Ea = A[node], Eb = A[F.cstride / 2 + node], Ha = A[F.cstride + node], Hb = A[3 * F.cstride / 2 + node];
        const double c3[3] = {cx, cy, cz};
        float P[3]; P[n] = (float)(F.coord - c3[n]); P[a] = (float)(F.xa[ia] - c3[a]); P[b] = (float)(F.xb[ib] - c3[b]);
        float2 Jv[3], Mv[3];
        Jv[n] = Mv[n] = make_float2(0.f, 0.f);
        Jv[a] = make_float2(-sw * Hb.x, -sw * Hb.y); Jv[b] = make_float2(sw * Ha.x, sw * Ha.y);      // J = n x H
        Mv[a] = make_float2(sw * Eb.x, sw * Eb.y);   Mv[b] = make_float2(-sw * Ea.x, -sw * Ea.y);    // M = -n x E
        const long long q = F.off + node;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            pos[c * npts + q] = P[c];
            reinterpret_cast<float2*>(J)[c * npts + q] = Jv[c];
            reinterpret_cast<float2*>(M)[c * npts + q] = Mv[c];
        }
        const double re = ((double)Ea.x * Hb.x + (double)Ea.y * Hb.y) - ((double)Eb.x * Ha.x + (double)Eb.y * Ha.y);
        pr = 0.5 * s * scale * scale * re * dA;
    }
    __shared__ double red[8];
    for (int o = 16; o > 0; o >>= 1) pr += __shfl_down_sync(0xffffffffu, pr, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = pr;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int i = 0; i < 8; ++i) t += red[i]; prad_partial[blockIdx.x] = t; }
}

// K11 far field: N = sum J e^{jk r^.r'}, L = sum M e^{jk r^.r'} projected on theta^/phi^ (App. A6)
// grid (ceil(ndir / 8), nsplit): a block sums its share of the surface points for 8 directions, one per warp.  The points
// are staged through shared memory in chunks of 256, so every point is fetched once per 8 directions (the sources of a
// 100 M-cell scene are 76 MB: per direction that read was the bound); fp64 phase and accumulation; farfield_final_kernel
// adds the shares in a fixed order (deterministic) and projects.
#define FF_CHUNK 256
__global__ void __launch_bounds__(256) farfield_kernel(long long npts, const float* __restrict__ pos,
        const float* __restrict__ J, const float* __restrict__ M, double k, int ndir,
        const double* __restrict__ theta, const double* __restrict__ phi, double* __restrict__ partial)
{
    __shared__ float sp_[3][FF_CHUNK];
    __shared__ float2 sj_[3][FF_CHUNK], sm_[3][FF_CHUNK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.x * 8 + warp;
    const bool live = d < ndir;
    double ux = 0, uy = 0, uz = 0;
    if (live) {
        double st, ct, sp, cp;
        sincos(theta[d], &st, &ct); sincos(phi[d], &sp, &cp);
        // phase in turns, formed and reduced in fp64 (k r can be hundreds of radians on a large box), then one fp32 sincospi
        ux = k * st * cp * 0.15915494309189535; uy = k * st * sp * 0.15915494309189535; uz = k * ct * 0.15915494309189535;
    }
    double a[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = 0.0;
    const long long per = (npts + gridDim.y - 1) / gridDim.y;
    const long long q0 = per * blockIdx.y, q1 = min(npts, q0 + per);
    for (long long c0 = q0; c0 < q1; c0 += FF_CHUNK) {
        const int m = (int)min((long long)FF_CHUNK, q1 - c0);
        __syncthreads();
        if ((int)threadIdx.x < m) {
            const long long q = c0 + threadIdx.x;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                sp_[c][threadIdx.x] = pos[c * npts + q];
                sj_[c][threadIdx.x] = reinterpret_cast<const float2*>(J)[c * npts + q];
                sm_[c][threadIdx.x] = reinterpret_cast<const float2*>(M)[c * npts + q];
            }
        }
        __syncthreads();
        if (live) {
            for (int t = lane; t < m; t += 32) {
                double tn = ux * (double)sp_[0][t] + uy * (double)sp_[1][t] + uz * (double)sp_[2][t];
                tn -= rint(tn);
                float sn, cn; sincospif(2.0f * (float)tn, &sn, &cn);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float2 jv = sj_[c][t], mv = sm_[c][t];
                    a[2 * c] += (double)(jv.x * cn - jv.y * sn); a[2 * c + 1] += (double)(jv.x * sn + jv.y * cn);
                    a[6 + 2 * c] += (double)(mv.x * cn - mv.y * sn); a[6 + 2 * c + 1] += (double)(mv.x * sn + mv.y * cn);
                }
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        double v = a[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) partial[((long long)d * gridDim.y + blockIdx.y) * 12 + i] = v;
    }
}
__global__ void farfield_final_kernel(int ndir, int nsplit, const double* __restrict__ partial,
        const double* __restrict__ theta, const double* __restrict__ phi, float* __restrict__ out)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= ndir) return;
    double s[12];
    for (int i = 0; i < 12; ++i) { s[i] = 0; for (int q = 0; q < nsplit; ++q) s[i] += partial[((long long)d * nsplit + q) * 12 + i]; }
    double st, ct, sp, cp;
    sincos(theta[d], &st, &ct); sincos(phi[d], &sp, &cp);
    // theta^ = (ct cp, ct sp, -st), phi^ = (-sp, cp, 0)
    for (int part = 0; part < 2; ++part) {             // 0: N from J, 1: L from M
        const double* v = s + 6 * part;
        for (int ri = 0; ri < 2; ++ri) {
            const double vx = v[ri], vy = v[2 + ri], vz = v[4 + ri];
            out[((long long)d * 4 + 2 * part) * 2 + ri] = (float)(vx * ct * cp + vy * ct * sp - vz * st);
            out[((long long)d * 4 + 2 * part + 1) * 2 + ri] = (float)(-vx * sp + vy * cp);
        }
    }
}
