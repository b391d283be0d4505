/*
 * fdtd_ref.c — CPU ORACLE of the FDTD time-stepping path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's library.  The product (fdtd-solver-antennas_b200) never does.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in openEMS (third-party, not
 * vendored, not pinned, not installable here: SURVEY.md §0, §8c).  This file restates
 * the published openEMS engine algorithm in plain scalar C:
 *   - engine update equations            : App. A1  (openEMS engine.cpp UpdateVoltages/UpdateCurrents)
 *   - excitation                         : App. A2/A5 (engine_ext_excitation.cpp Apply2Voltages)
 *   - Mur first order ABC                : App. A3  (engine_ext_mur_abc.cpp)
 *   - split-flux UPML                    : App. A4  (engine_ext_upml.cpp)
 *   - probe integrals / DFT              : App. A5  (processintegral.cpp, ports.py CalcPort)
 *   - NF2FF node interpolation / DFT     : App. A6  (processfields.cpp, nf2ff_calc.cpp)
 *   - energy estimate                    : App. A7  (engine_interface_fdtd.cpp CalcFastEnergy)
 * It is anchored on the reference's call sites for the path:
 *   antenna_sim/solver_fdtd_openems_microstrip_3d.py:82-93,176-179,214,225
 * and validated by analytic known-answer tests (tests/test_oracle_physics.py).
 *
 * Layout and linear indices are those of include/b200fdtd.h so the same inputs drive
 * this oracle and the CUDA engine:  lin(c,k,j,i) = ((c*(nz+2) + k+1)*ny + j)*px + i.
 * fp32 field arithmetic follows the engine contract:
 *   curl = ((a - b) - c) + d ;  f = fmaf(ca, f, cb*curl)
 * (compile with -ffp-contract=off so the compiler adds no contraction of its own).
 * DFT accumulators and the far field are double precision here (they are the reference
 * for the fp32 accumulators of the CUDA path).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t x0, y0, z0, bx, by, bz;
    float *flux_v, *flux_i;
    const float *vv, *vvfo, *vvfn, *ii, *iifo, *iifn;
} ref_pml_box;

typedef struct {
    int32_t normal, plane, a0, a1, b0, b1;
    double* acc;                 /* [4][nfreq][nb][na][2] double */
    float* td;                   /* optional time-domain store [td_max][4][nb][na] (openEMS: nf2ff_E/H_n.h5 dumps) */
    int32_t td_max;
} ref_face;

typedef struct {
    int32_t nx, ny, nz, px;
    float *volt, *curr;
    const float *vv, *vi, *ii, *iv;
    /* excitation */
    int64_t n_exc; const int64_t* exc_idx; const float* exc_amp; const int32_t* exc_delay;
    const float* exc_sig; int32_t exc_siglen;
    /* mur */
    int64_t n_mur; const int64_t* mur_dst; const int64_t* mur_src; const float* mur_coeff; float* mur_tmp;
    /* pml */
    int32_t n_pml; const ref_pml_box* pml;
    /* probes */
    int32_t n_probes; const int32_t* pr_kind; const int64_t* pr_off; const int64_t* pr_idx; const float* pr_w;
    int32_t interval; int32_t max_samples; double* pr_series; int32_t pr_nfreq; const double* pr_freqs;
    double* pr_dft; double dt;
    /* nf2ff */
    int32_t n_faces; const ref_face* faces; int32_t nf_nfreq; const double* nf_freqs;
    const float* inv_len[3]; const float* inv_dual[3];
    /* state */
    int64_t ts;
    int32_t threads;             /* OpenMP threads for the volume loops (<=0: default) */
} ref_engine;

static inline float upd1(float ca, float f, float cb, float a, float b, float c, float d)
{
    float curl = ((a - b) - c) + d;
    float t = cb * curl;
    return fmaf(ca, f, t);
}

/* App. A1, voltage update */
void ref_update_e(ref_engine* e)
{
    const int nx = e->nx, ny = e->ny, nz = e->nz, px = e->px;
    const int64_t sz = (int64_t)ny * px, cs = (int64_t)(nz + 2) * sz;
    float* v = e->volt; const float* h = e->curr;
    (void)nx;
#pragma omp parallel for collapse(2) schedule(static) num_threads(e->threads > 0 ? e->threads : 1)
    for (int k = 0; k < nz; ++k) {
        for (int j = 0; j < ny; ++j) {
            const int64_t row = (int64_t)(k + 1) * sz + (int64_t)j * px;
            for (int i = 0; i < px; ++i) {
                const int64_t q = row + i;
                const float hx = h[q], hy = h[cs + q], hz = h[2 * cs + q];
                const float hz_jm = j > 0 ? h[2 * cs + q - px] : 0.f;
                const float hx_jm = j > 0 ? h[q - px] : 0.f;
                const float hz_im = i > 0 ? h[2 * cs + q - 1] : 0.f;
                const float hy_im = i > 0 ? h[cs + q - 1] : 0.f;
                const float hx_km = h[q - sz], hy_km = h[cs + q - sz];
                v[q]          = upd1(e->vv[q],          v[q],          e->vi[q],          hz, hz_jm, hy, hy_km);
                v[cs + q]     = upd1(e->vv[cs + q],     v[cs + q],     e->vi[cs + q],     hx, hx_km, hz, hz_im);
                v[2 * cs + q] = upd1(e->vv[2 * cs + q], v[2 * cs + q], e->vi[2 * cs + q], hy, hy_im, hx, hx_jm);
            }
        }
    }
}

/* App. A1, current update */
void ref_update_h(ref_engine* e)
{
    const int ny = e->ny, nz = e->nz, px = e->px;
    const int64_t sz = (int64_t)ny * px, cs = (int64_t)(nz + 2) * sz;
    float* h = e->curr; const float* v = e->volt;
#pragma omp parallel for collapse(2) schedule(static) num_threads(e->threads > 0 ? e->threads : 1)
    for (int k = 0; k < nz; ++k) {
        for (int j = 0; j < ny; ++j) {
            const int64_t row = (int64_t)(k + 1) * sz + (int64_t)j * px;
            for (int i = 0; i < px; ++i) {
                const int64_t q = row + i;
                const float ex = v[q], ey = v[cs + q], ez = v[2 * cs + q];
                const float ez_jp = j + 1 < ny ? v[2 * cs + q + px] : 0.f;
                const float ex_jp = j + 1 < ny ? v[q + px] : 0.f;
                const float ez_ip = i + 1 < px ? v[2 * cs + q + 1] : 0.f;
                const float ey_ip = i + 1 < px ? v[cs + q + 1] : 0.f;
                const float ex_kp = v[q + sz], ey_kp = v[cs + q + sz];
                h[q]          = upd1(e->ii[q],          h[q],          e->iv[q],          ez, ez_jp, ey, ey_kp);
                h[cs + q]     = upd1(e->ii[cs + q],     h[cs + q],     e->iv[cs + q],     ex, ex_kp, ez, ez_ip);
                h[2 * cs + q] = upd1(e->ii[2 * cs + q], h[2 * cs + q], e->iv[2 * cs + q], ey, ey_ip, ex, ex_jp);
            }
        }
    }
}

/* App. A2/A5: soft voltage excitation, signal index = numTS - delay */
void ref_excite(ref_engine* e)
{
    for (int64_t n = 0; n < e->n_exc; ++n) {
        const int64_t pos = e->ts - e->exc_delay[n];
        if (pos < 0 || pos >= e->exc_siglen) continue;
        const int64_t q = e->exc_idx[n];
        e->volt[q] = fmaf(e->exc_amp[n], e->exc_sig[pos], e->volt[q]);
    }
}

/* App. A3: phase 0 pre-update, 1 post-update, 2 apply */
void ref_mur(ref_engine* e, int phase)
{
    float* v = e->volt;
#pragma omp parallel for schedule(static) num_threads(e->threads > 0 ? e->threads : 1)
    for (int64_t n = 0; n < e->n_mur; ++n) {
        if (phase == 0)      e->mur_tmp[n] = fmaf(-e->mur_coeff[n], v[e->mur_dst[n]], v[e->mur_src[n]]);
        else if (phase == 1) e->mur_tmp[n] = fmaf(e->mur_coeff[n], v[e->mur_src[n]], e->mur_tmp[n]);
        else                 v[e->mur_dst[n]] = e->mur_tmp[n];
    }
}

/* App. A4: split-flux UPML pre/post passes.  which 0 = voltages, 1 = currents */
void ref_pml(ref_engine* e, int which, int post)
{
    const int ny = e->ny, nz = e->nz, px = e->px;
    const int64_t sz = (int64_t)ny * px, cs = (int64_t)(nz + 2) * sz;
    float* fld = which == 0 ? e->volt : e->curr;
    for (int b = 0; b < e->n_pml; ++b) {
        const ref_pml_box* B = &e->pml[b];
        float* flux = which == 0 ? B->flux_v : B->flux_i;
        const float* a = which == 0 ? B->vv : B->ii;
        const float* fo = which == 0 ? B->vvfo : B->iifo;
        const float* fn = which == 0 ? B->vvfn : B->iifn;
#pragma omp parallel for collapse(2) schedule(static) num_threads(e->threads > 0 ? e->threads : 1)
        for (int c = 0; c < 3; ++c)
            for (int z = 0; z < B->bz; ++z)
                for (int y = 0; y < B->by; ++y)
                    for (int x = 0; x < B->bx; ++x) {
                        const int64_t l = (((int64_t)c * B->bz + z) * B->by + y) * B->bx + x;
                        const int64_t q = c * cs + (int64_t)(B->z0 + z + 1) * sz + (int64_t)(B->y0 + y) * px + (B->x0 + x);
                        if (!post) {
                            const float fl = flux[l];
                            const float t = fo[l] * fl;
                            const float hh = fmaf(a[l], fld[q], -t);
                            fld[q] = fl; flux[l] = hh;
                        } else {
                            const float hh = flux[l];
                            const float vv = fld[q];
                            flux[l] = vv;
                            fld[q] = fmaf(fn[l], vv, hh);
                        }
                    }
    }
}

static void twiddle(double f, double t, double* c, double* s)
{
    double ph = f * t; ph -= floor(ph);
    *c = cos(2.0 * M_PI * ph); *s = sin(2.0 * M_PI * ph);
}

/* App. A5: probe integrals, time series, DFT (double) */
void ref_probes(ref_engine* e)
{
    const int64_t s = e->ts / e->interval - 1;
    if (s < 0 || s >= e->max_samples) return;
    for (int p = 0; p < e->n_probes; ++p) {
        const float* fld = e->pr_kind[p] == 0 ? e->volt : e->curr;
        double acc = 0.0;
        for (int64_t n = e->pr_off[p]; n < e->pr_off[p + 1]; ++n) acc += (double)e->pr_w[n] * (double)fld[e->pr_idx[n]];
        e->pr_series[(int64_t)p * e->max_samples + s] = acc;
        const double tm = (e->pr_kind[p] == 0 ? (double)e->ts : (double)e->ts + 0.5) * e->dt;
        for (int q = 0; q < e->pr_nfreq; ++q) {
            double c, sn; twiddle(e->pr_freqs[q], tm, &c, &sn);
            e->pr_dft[((int64_t)p * e->pr_nfreq + q) * 2] += acc * c;
            e->pr_dft[((int64_t)p * e->pr_nfreq + q) * 2 + 1] -= acc * sn;
        }
    }
}

/* App. A6: node-interpolated tangential E/H on the Huygens faces, running DFT (double) */
void ref_nf2ff(ref_engine* e)
{
    const int ny = e->ny, nz = e->nz, px = e->px;
    const int64_t sz = (int64_t)ny * px, cs = (int64_t)(nz + 2) * sz;
    const int64_t st[3] = {1, px, sz};
    for (int fi = 0; fi < e->n_faces; ++fi) {
        const ref_face* F = &e->faces[fi];
        const int n = F->normal, a = (n + 1) % 3, b = (n + 2) % 3;
        const int na = F->a1 - F->a0 + 1, nb = F->b1 - F->b0 + 1;
        const int64_t nn = (int64_t)na * nb;
        const int oa = (a == 2), ob = (b == 2);
        const float* va = e->volt + a * cs; const float* vb = e->volt + b * cs;
        const float* ca = e->curr + a * cs; const float* cb = e->curr + b * cs;
        const int64_t smp = e->ts / e->interval - 1;
        if (F->td && smp >= 0 && smp < F->td_max) {
            /* the dump itself: fp32 node-interpolated samples, formed in fp32 exactly like the CUDA kernel forms them */
            for (int ib = F->b0; ib <= F->b1; ++ib)
                for (int ia = F->a0; ia <= F->a1; ++ia) {
                    int co[3]; co[n] = F->plane; co[a] = ia; co[b] = ib;
                    const int64_t l0 = (int64_t)(co[2] + 1) * sz + (int64_t)co[1] * px + co[0];
                    const int64_t node = (int64_t)(ib - F->b0) * na + (ia - F->a0);
                    float* d = F->td + smp * 4 * nn + node;
                    d[0] = 0.5f * (va[l0] * e->inv_len[a][ia + oa] + va[l0 - st[a]] * e->inv_len[a][ia + oa - 1]);
                    d[nn] = 0.5f * (vb[l0] * e->inv_len[b][ib + ob] + vb[l0 - st[b]] * e->inv_len[b][ib + ob - 1]);
                    d[2 * nn] = 0.25f * e->inv_dual[a][ia + oa] * ((ca[l0] + ca[l0 - st[b]]) + (ca[l0 - st[n]] + ca[l0 - st[b] - st[n]]));
                    d[3 * nn] = 0.25f * e->inv_dual[b][ib + ob] * ((cb[l0] + cb[l0 - st[a]]) + (cb[l0 - st[n]] + cb[l0 - st[a] - st[n]]));
                }
        }
        for (int q = 0; q < e->nf_nfreq; ++q) {
            double cE, sE, cH, sH;
            twiddle(e->nf_freqs[q], (double)e->ts * e->dt, &cE, &sE);
            twiddle(e->nf_freqs[q], ((double)e->ts + 0.5) * e->dt, &cH, &sH);
            for (int ib = F->b0; ib <= F->b1; ++ib)
                for (int ia = F->a0; ia <= F->a1; ++ia) {
                    int co[3]; co[n] = F->plane; co[a] = ia; co[b] = ib;
                    const int64_t l0 = (int64_t)(co[2] + 1) * sz + (int64_t)co[1] * px + co[0];
                    const int64_t node = (int64_t)(ib - F->b0) * na + (ia - F->a0);
                    const double Ea = 0.5 * ((double)va[l0] * e->inv_len[a][ia + oa] + (double)va[l0 - st[a]] * e->inv_len[a][ia + oa - 1]);
                    const double Eb = 0.5 * ((double)vb[l0] * e->inv_len[b][ib + ob] + (double)vb[l0 - st[b]] * e->inv_len[b][ib + ob - 1]);
                    const double Ha = 0.25 * e->inv_dual[a][ia + oa] *
                        ((double)ca[l0] + ca[l0 - st[b]] + ca[l0 - st[n]] + ca[l0 - st[b] - st[n]]);
                    const double Hb = 0.25 * e->inv_dual[b][ib + ob] *
                        ((double)cb[l0] + cb[l0 - st[a]] + cb[l0 - st[n]] + cb[l0 - st[a] - st[n]]);
                    const double v[4] = {Ea, Eb, Ha, Hb};
                    for (int c = 0; c < 4; ++c) {
                        double* d = F->acc + (((int64_t)c * e->nf_nfreq + q) * nn + node) * 2;
                        const double cn = c >= 2 ? cH : cE, sn = c >= 2 ? sH : sE;
                        d[0] += v[c] * cn; d[1] -= v[c] * sn;
                    }
                }
        }
    }
}

/* App. A7: CalcFastEnergy over the owned planes */
double ref_energy(const ref_engine* e)
{
    const int ny = e->ny, nz = e->nz, px = e->px;
    const int64_t sz = (int64_t)ny * px, cs = (int64_t)(nz + 2) * sz;
    double sv = 0.0, sc = 0.0;
    for (int c = 0; c < 3; ++c)
        for (int64_t q = sz; q < (int64_t)(nz + 1) * sz; ++q) {
            sv += (double)e->volt[c * cs + q] * e->volt[c * cs + q];
            sc += (double)e->curr[c * cs + q] * e->curr[c * cs + q];
        }
    return 0.5 * 8.85418781762e-12 * sv + 0.5 * 1.256637062e-6 * sc;
}

/* one full time step in openEMS order (App. A1): pre-E ext, E, post-E ext, apply, pre-H, H, post-H, ++ts, sample */
void ref_step(ref_engine* e)
{
    ref_mur(e, 0);
    ref_pml(e, 0, 0);
    ref_update_e(e);
    ref_pml(e, 0, 1);
    ref_mur(e, 1);
    ref_excite(e);
    ref_mur(e, 2);
    ref_pml(e, 1, 0);
    ref_update_h(e);
    ref_pml(e, 1, 1);
    e->ts += 1;
    if (e->interval > 0 && (e->ts % e->interval) == 0) {
        if (e->n_probes > 0) ref_probes(e);
        if (e->n_faces > 0) ref_nf2ff(e);
    }
}

void ref_run(ref_engine* e, int64_t nsteps) { for (int64_t s = 0; s < nsteps; ++s) ref_step(e); }

/* split phases (z-slab halo tests): phase 0 = E half step, 1 = H half step, 2 = sampling if due */
void ref_half_step(ref_engine* e, int phase)
{
    if (phase == 0) {
        ref_mur(e, 0); ref_pml(e, 0, 0); ref_update_e(e); ref_pml(e, 0, 1); ref_mur(e, 1); ref_excite(e); ref_mur(e, 2);
    } else if (phase == 1) {
        ref_pml(e, 1, 0); ref_update_h(e); ref_pml(e, 1, 1);
        e->ts += 1;
    } else {
        if (e->interval > 0 && (e->ts % e->interval) == 0) {
            if (e->n_probes > 0) ref_probes(e);
            if (e->n_faces > 0) ref_nf2ff(e);
        }
    }
}

/* the z-slab overlap protocol of the product cuts each half step in two (include/b200fdtd.h half_step_part).
 * The oracle keeps whole passes: part 0 does nothing, part 1 does the whole half step, which is equivalent
 * because the halo a part-1 launch needs has arrived by then. */
void ref_half_step_part(ref_engine* e, int phase, int part)
{
    if (part == 1) ref_half_step(e, phase);
}

/* App. A6 radiation integrals: out[ndir][4][2] = N_theta, N_phi, L_theta, L_phi (double) */
void ref_farfield(int64_t npts, const double* pos, const double* J, const double* M, double k,
                  int ndir, const double* theta, const double* phi, double* out)
{
    for (int d = 0; d < ndir; ++d) {
        const double st = sin(theta[d]), ct = cos(theta[d]), sp = sin(phi[d]), cp = cos(phi[d]);
        double s[12]; memset(s, 0, sizeof(s));
        for (int64_t q = 0; q < npts; ++q) {
            const double ph = k * (pos[q] * st * cp + pos[npts + q] * st * sp + pos[2 * npts + q] * ct);
            const double cn = cos(ph), sn = sin(ph);
            for (int c = 0; c < 3; ++c) {
                const double jr = J[(c * npts + q) * 2], ji = J[(c * npts + q) * 2 + 1];
                const double mr = M[(c * npts + q) * 2], mi = M[(c * npts + q) * 2 + 1];
                s[2 * c] += jr * cn - ji * sn; s[2 * c + 1] += jr * sn + ji * cn;
                s[6 + 2 * c] += mr * cn - mi * sn; s[6 + 2 * c + 1] += mr * sn + mi * cn;
            }
        }
        for (int part = 0; part < 2; ++part) {
            const double* v = s + 6 * part;
            for (int ri = 0; ri < 2; ++ri) {
                const double vx = v[ri], vy = v[2 + ri], vz = v[4 + ri];
                out[((int64_t)d * 4 + 2 * part) * 2 + ri] = vx * ct * cp + vy * ct * sp - vz * st;
                out[((int64_t)d * 4 + 2 * part + 1) * 2 + ri] = -vx * sp + vy * cp;
            }
        }
    }
}

int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
