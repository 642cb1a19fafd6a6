/*
 * cavity_oracle.c -- CPU oracle (plain C restatement).  TEST INFRASTRUCTURE ONLY; see the
 * header for the rules.  Build: oracle/Makefile (gcc -O2 -ffp-contract=off, i.e. separate
 * multiply and add like a baseline x86-64 HOOMD build; a second -mfma -ffp-contract=fast
 * build quantifies the contraction difference).
 */
#include "cavity_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* HOOMD's __scalar_as_int: the int that overlays the low 32 bits of the double
 * (used at src/CavityForceCompute.cc:83 and :190). */
static int type_bits(double w)
    {
    int32_t t;
    memcpy(&t, &w, sizeof(t));
    return (int)t;
    }

int orc_find_photon(const double* pos4, uint32_t N, uint32_t L_typeid)
    {
    /* src/CavityForceCompute.cc:81-88 */
    for (uint32_t i = 0; i < N; i++)
        if (type_bits(pos4[4 * (size_t)i + 3]) == (int)L_typeid)
            return (int)i;
    return -1;
    }

int orc_cavity_force(const double* pos4, const double* charge, const int32_t* image3,
                     double* force4, uint32_t N, double Lx, double Ly, double Lz,
                     uint32_t L_typeid, double omegac, double couplstr, double phmass,
                     double energies[3], double dipole[3])
    {
    /* src/CavityForceCompute.h:38-42: K = phmass * omegac * omegac (left to right) */
    const double K = phmass * omegac * omegac;
    const double g = couplstr;

    /* :145 every force entry starts as {+0,+0,+0,+0} */
    memset(force4, 0, sizeof(double) * 4 * (size_t)N);
    energies[0] = energies[1] = energies[2] = 0.0;
    dipole[0] = dipole[1] = dipole[2] = 0.0;

    /* :148-156 */
    const int photon = orc_find_photon(pos4, N, L_typeid);
    if (photon < 0)
        return -1;

    /* :91-111 unwrapped positions, a fresh N-element array per call exactly like :162 */
    double* u = (double*)malloc(sizeof(double) * 3 * (size_t)N);
    for (uint32_t i = 0; i < N; i++)
        {
        const double* p = pos4 + 4 * (size_t)i;
        const int32_t* im = image3 + 3 * (size_t)i;
        u[3 * (size_t)i + 0] = p[0] + (double)im[0] * Lx;
        u[3 * (size_t)i + 1] = p[1] + (double)im[1] * Ly;
        u[3 * (size_t)i + 2] = p[2] + (double)im[2] * Lz;
        }

    /* :113-129 strict index-ascending serial sum, skipping only index `photon` */
    double dx = 0.0, dy = 0.0, dz = 0.0;
    for (uint32_t i = 0; i < N; i++)
        {
        if ((int)i == photon)
            continue;
        const double c = charge[i];
        dx = dx + c * u[3 * (size_t)i + 0];
        dy = dy + c * u[3 * (size_t)i + 1];
        dz = dz + c * u[3 * (size_t)i + 2];
        }
    dipole[0] = dx;
    dipole[1] = dy;
    dipole[2] = dz;

    /* :169-171 */
    const double qx = u[3 * (size_t)photon + 0];
    const double qy = u[3 * (size_t)photon + 1];
    const double qz = u[3 * (size_t)photon + 2];

    /* :174-176; dot(a,b) = a.x*b.x + a.y*b.y + a.z*b.z, the xy vectors carry z = 0 */
    energies[0] = 0.5 * K * (qx * qx + qy * qy + qz * qz);
    energies[1] = g * (dx * qx + dy * qy + 0.0 * 0.0);
    energies[2] = 0.5 * (g * g / K) * (dx * dx + dy * dy + 0.0 * 0.0);

    /* :183 Dq = q_xy + (g/K) * d_xy */
    const double gk = g / K;
    const double Dqx = qx + gk * dx;
    const double Dqy = qy + gk * dy;

    /* :188-200 molecular particles (type != L): f = ((-g) * c) * Dq, z = +0 */
    for (uint32_t i = 0; i < N; i++)
        {
        if (type_bits(pos4[4 * (size_t)i + 3]) != (int)L_typeid)
            {
            const double s = -g * charge[i];
            force4[4 * (size_t)i + 0] = s * Dqx;
            force4[4 * (size_t)i + 1] = s * Dqy;
            force4[4 * (size_t)i + 2] = 0.0;
            }
        }

    /* :203-207 photon: -K*q - g*d_xy (z: -K*qz - g*0) */
    force4[4 * (size_t)photon + 0] = -K * qx - g * dx;
    force4[4 * (size_t)photon + 1] = -K * qy - g * dy;
    force4[4 * (size_t)photon + 2] = -K * qz - g * 0.0;
    force4[4 * (size_t)photon + 3] = 0.0; /* :180 */

    free(u);
    return photon;
    }

void orc_dipole_exact(const double* pos4, const double* charge, const int32_t* image3,
                      uint32_t N, double Lx, double Ly, double Lz, int photon_idx,
                      double dipole[3])
    {
    const double L[3] = {Lx, Ly, Lz};
    long double s[3] = {0, 0, 0}, comp[3] = {0, 0, 0};
    for (uint32_t i = 0; i < N; i++)
        {
        if ((int)i == photon_idx)
            continue;
        for (int c = 0; c < 3; c++)
            {
            /* the terms are the reference's double-rounded terms (:107-109, :124) */
            volatile double uu = pos4[4 * (size_t)i + c] + (double)image3[3 * (size_t)i + c] * L[c];
            volatile double term = charge[i] * uu;
            long double t = (long double)term;
            long double sum = s[c] + t;
            if (fabsl(s[c]) >= fabsl(t))
                comp[c] += (s[c] - sum) + t;
            else
                comp[c] += (t - sum) + s[c];
            s[c] = sum;
            }
        }
    for (int c = 0; c < 3; c++)
        dipole[c] = (double)(s[c] + comp[c]);
    }

double orc_bussi_rescale_factor(double K, double dof, double deltaT, double set_T, double tau,
                                double r_normal, double gamma_draw)
    {
    /* src/BussiReservoirThermostat.h:183-184 */
    if (dof == 0)
        return 1.0;

    /* :186-190 */
    double c = 0.0;
    if (tau != 0.0)
        c = exp(-deltaT / tau);

    /* :192-200: normal first, then (only when dof > 1) 2 * gamma((dof-1)/2, 1) */
    const double R = r_normal;
    double r_gamma = 0.0;
    if (dof > 1.0)
        r_gamma = 2.0 * gamma_draw;

    /* :202-208 */
    const double v = set_T / 2.0 / K;
    const double term1 = v * (1.0 - c) * (r_gamma + R * R);
    const double term2 = 2.0 * R * sqrt(v * (1.0 - c) * c);
    const double alpha2 = c + term1 + term2;
    const double alpha = sqrt(alpha2);

    /* :210-214 sign rule, Bussi 2009 eq. (A8) */
    const double K_bar = set_T * dof / 2.0;
    const double sign_term = R + sqrt(c * dof * K / ((1.0 - c) * K_bar));

    /* :217-224 */
    return sign_term >= 0.0 ? alpha : -alpha;
    }

double orc_kinetic_energy(const double* vel4, const uint32_t* idx, uint32_t n)
    {
    double ke = 0.0;
    for (uint32_t j = 0; j < n; j++)
        {
        const double* v = vel4 + 4 * (size_t)(idx ? idx[j] : j);
        ke += v[3] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        }
    return 0.5 * ke;
    }

void orc_rescale_velocities(double* vel4, const uint32_t* idx, uint32_t n, double alpha)
    {
    for (uint32_t j = 0; j < n; j++)
        {
        double* v = vel4 + 4 * (size_t)(idx ? idx[j] : j);
        v[0] *= alpha;
        v[1] *= alpha;
        v[2] *= alpha;
        }
    }

double orc_bussi_step(double* vel4, const uint32_t* idx, uint32_t n, double dof, double deltaT,
                      double set_T, double tau, double r_normal, double gamma_draw,
                      double reservoir[2], double* ke_out)
    {
    /* src/BussiReservoirThermostat.h:45-48 */
    if (deltaT == 0.0)
        return 1.0;
    /* :50-55 (ComputeThermo stand-in) */
    const double ke = orc_kinetic_energy(vel4, idx, n);
    if (ke_out)
        *ke_out = ke;
    /* :57-61 */
    if (dof != 0 && ke == 0)
        return NAN;
    /* :72-76 */
    const double alpha = orc_bussi_rescale_factor(ke, dof, deltaT, set_T, tau, r_normal, gamma_draw);
    /* :86-95 */
    const double delta = ke * (1.0 - alpha * alpha);
    reservoir[0] += delta;
    reservoir[1] = delta;
    orc_rescale_velocities(vel4, idx, n, alpha);
    return alpha;
    }

void orc_rhok(const double* pos, uint32_t stride, uint32_t N, const double* kvec, uint32_t K,
              double* rho_re, double* rho_im)
    {
    for (uint32_t k = 0; k < K; k++)
        {
        const double kx = kvec[3 * k + 0], ky = kvec[3 * k + 1], kz = kvec[3 * k + 2];
        long double sr = 0, si = 0;
        for (uint32_t j = 0; j < N; j++)
            {
            const double* r = pos + (size_t)stride * j;
            /* analysis.py:42 np.dot(positions, k): x*kx + y*ky + z*kz in double */
            const double kr = r[0] * kx + r[1] * ky + r[2] * kz;
            sr += (long double)cos(kr);
            si += (long double)sin(kr);
            }
        rho_re[k] = (double)sr;
        rho_im[k] = (double)si;
        }
    }

/* HOOMD's BoxDim::wrap for an orthorhombic box periodic in x, y, z: the box is [-L/2, L/2); a coordinate at or above
 * the upper face moves down by L and its image flag goes up by one, a coordinate below the lower face the other way;
 * one shift per direction.  [HOOMD-upstream: not in the reference tree, restated from its definition; it is the
 * "wrap" of SURVEY.md 8a row a11, the step after r <- r + v dt in TwoStepConstantVolume::integrateStepOne.] */
void orc_wrap(double* p, int32_t* img, const double L[3])
    {
    for (int c = 0; c < 3; c++)
        {
        const double hi = 0.5 * L[c];
        if (p[c] >= hi)
            {
            p[c] = p[c] - L[c];
            img[c] += 1;
            }
        else if (p[c] < -hi)
            {
            p[c] = p[c] + L[c];
            img[c] -= 1;
            }
        }
    }

void orc_nve_step_w(double* pos4, double* vel4, const double* charge, int32_t* image3,
                    double* force4, uint32_t N, double Lx, double Ly, double Lz, uint32_t L_typeid,
                    double omegac, double couplstr, double phmass, double dt, double energies[3], int wrap);

void orc_nve_step(double* pos4, double* vel4, const double* charge, const int32_t* image3,
                  double* force4, uint32_t N, double Lx, double Ly, double Lz, uint32_t L_typeid,
                  double omegac, double couplstr, double phmass, double dt, double energies[3])
    {
    orc_nve_step_w(pos4, vel4, charge, (int32_t*)image3, force4, N, Lx, Ly, Lz, L_typeid, omegac, couplstr, phmass, dt,
                   energies, 0);
    }

void orc_nve_step_w(double* pos4, double* vel4, const double* charge, int32_t* image3,
                    double* force4, uint32_t N, double Lx, double Ly, double Lz, uint32_t L_typeid,
                    double omegac, double couplstr, double phmass, double dt, double energies[3], int wrap)
    {
    double dip[3];
    const double Lbox[3] = {Lx, Ly, Lz};
    /* half kick + drift with the force from the previous call */
    for (uint32_t i = 0; i < N; i++)
        {
        double* v = vel4 + 4 * (size_t)i;
        double* p = pos4 + 4 * (size_t)i;
        const double* f = force4 + 4 * (size_t)i;
        const double hm = 0.5 * dt / v[3];
        for (int c = 0; c < 3; c++)
            {
            v[c] = v[c] + hm * f[c];
            p[c] = p[c] + dt * v[c];
            }
        if (wrap)
            orc_wrap(p, image3 + 3 * (size_t)i, Lbox);
        }
    orc_cavity_force(pos4, charge, image3, force4, N, Lx, Ly, Lz, L_typeid, omegac, couplstr,
                     phmass, energies, dip);
    for (uint32_t i = 0; i < N; i++)
        {
        double* v = vel4 + 4 * (size_t)i;
        const double* f = force4 + 4 * (size_t)i;
        const double hm = 0.5 * dt / v[3];
        for (int c = 0; c < 3; c++)
            v[c] = v[c] + hm * f[c];
        }
    }

double orc_nvt_step_w(double* pos4, double* vel4, const double* charge, int32_t* image3,
                      double* force4, uint32_t N, double Lx, double Ly, double Lz, uint32_t L_typeid,
                      double omegac, double couplstr, double phmass, double dt, uint32_t first, uint32_t n,
                      double dof, double set_T, double tau, double r_normal, double gamma_draw,
                      double reservoir[2], double* ke_io, double energies[3], int wrap);

double orc_nvt_step(double* pos4, double* vel4, const double* charge, const int32_t* image3,
                    double* force4, uint32_t N, double Lx, double Ly, double Lz, uint32_t L_typeid,
                    double omegac, double couplstr, double phmass, double dt, uint32_t first, uint32_t n,
                    double dof, double set_T, double tau, double r_normal, double gamma_draw,
                    double reservoir[2], double* ke_io, double energies[3])
    {
    return orc_nvt_step_w(pos4, vel4, charge, (int32_t*)image3, force4, N, Lx, Ly, Lz, L_typeid, omegac, couplstr, phmass, dt,
                          first, n, dof, set_T, tau, r_normal, gamma_draw, reservoir, ke_io, energies, 0);
    }

double orc_nvt_step_w(double* pos4, double* vel4, const double* charge, int32_t* image3,
                      double* force4, uint32_t N, double Lx, double Ly, double Lz, uint32_t L_typeid,
                      double omegac, double couplstr, double phmass, double dt, uint32_t first, uint32_t n,
                      double dof, double set_T, double tau, double r_normal, double gamma_draw,
                      double reservoir[2], double* ke_io, double energies[3], int wrap)
    {
    double dip[3];
    const double Lbox[3] = {Lx, Ly, Lz};
    double alpha = 1.0;
    const double ke = *ke_io;
    if (n > 0 && dof != 0.0)
        {
        if (ke == 0.0)
            return NAN; /* src/BussiReservoirThermostat.h:57-61 */
        alpha = orc_bussi_rescale_factor(ke, dof, dt, set_T, tau, r_normal, gamma_draw);
        const double delta = ke * (1.0 - alpha * alpha); /* :86-95 */
        reservoir[0] += delta;
        reservoir[1] = delta;
        }
    for (uint32_t i = 0; i < N; i++)
        {
        double* v = vel4 + 4 * (size_t)i;
        double* p = pos4 + 4 * (size_t)i;
        const double* f = force4 + 4 * (size_t)i;
        const double hm = 0.5 * dt / v[3];
        const int in = i >= first && i < first + n;
        for (int c = 0; c < 3; c++)
            {
            double vc = v[c];
            if (in)
                vc = vc * alpha;
            vc = vc + hm * f[c];
            v[c] = vc;
            p[c] = p[c] + dt * vc;
            }
        if (wrap)
            orc_wrap(p, image3 + 3 * (size_t)i, Lbox);
        }
    orc_cavity_force(pos4, charge, image3, force4, N, Lx, Ly, Lz, L_typeid, omegac, couplstr,
                     phmass, energies, dip);
    double acc = 0.0;
    for (uint32_t i = 0; i < N; i++)
        {
        double* v = vel4 + 4 * (size_t)i;
        const double* f = force4 + 4 * (size_t)i;
        const double hm = 0.5 * dt / v[3];
        for (int c = 0; c < 3; c++)
            v[c] = v[c] + hm * f[c];
        if (i >= first && i < first + n)
            acc += v[3] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        }
    *ke_io = 0.5 * acc;
    return alpha;
    }
