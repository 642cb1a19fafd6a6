/*
 * cavity_oracle.h -- CPU oracle for the cav-hoomd hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's CPU algorithm.  It exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can
 * check and time the CUDA product against it.  Nothing under cav_hoomd_b200/ (the product)
 * may include, link, import or call anything in oracle/.
 *
 * Parity pin: the reference's own tests hold NO golden vectors for this path
 * (SURVEY.md section 4 / 8c).  The restatement is pinned instead against the reference's own
 * translation unit src/CavityForceCompute.cc, compiled verbatim by path against a header
 * shim into oracle/_ref/ (see oracle/Makefile), and against tests/golden/ fixtures minted from
 * that build (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference checkout).
 */
#ifndef CAVITY_ORACLE_H
#define CAVITY_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/CavityForceCompute.cc:73-89 -- first i whose type id (low 32 bits of pos[i].w) is L. */
int orc_find_photon(const double* pos4, uint32_t N, uint32_t L_typeid);

/* src/CavityForceCompute.cc:134-208 (computeForces) with its helpers :91-111 (unwrap) and
 * :113-129 (dipole).  pos4/force4 are double4[N], image3 is int3[N], charge is double[N].
 * energies = {harmonic, coupling, dipole_self}; dipole = full 3-vector; returns photon index
 * (or -1, in which case forces and energies are zero, :149-156). */
int orc_cavity_force(const double* pos4, const double* charge, const int32_t* image3,
                     double* force4, uint32_t N, double Lx, double Ly, double Lz,
                     uint32_t L_typeid, double omegac, double couplstr, double phmass,
                     double energies[3], double dipole[3]);

/* Exact dipole (long double Neumaier sum of the double-rounded terms c_i*u_i) used to measure
 * the reduction-order error of BOTH the reference's serial sum and the CUDA reduction. */
void orc_dipole_exact(const double* pos4, const double* charge, const int32_t* image3,
                      uint32_t N, double Lx, double Ly, double Lz, int photon_idx,
                      double dipole[3]);

/* src/BussiReservoirThermostat.h:177-225 (compute_rescale_factor) with the two random draws
 * injected: r_normal = normal(rng), gamma_draw = gamma(rng) (the reference doubles it, :199).
 * Returns the signed rescale factor alpha. */
double orc_bussi_rescale_factor(double K, double dof, double deltaT, double set_T, double tau,
                                double r_normal, double gamma_draw);

/* Stand-in for HOOMD ComputeThermo (not in the reference tree): KE = 1/2 sum m |v|^2 over the
 * group, serial order.  vel4 is double4 {vx,vy,vz,mass}.  idx may be NULL (= 0..n-1). */
double orc_kinetic_energy(const double* vel4, const uint32_t* idx, uint32_t n);

/* v <- alpha v over the group (stand-in for the rescale inside HOOMD's step one). */
void orc_rescale_velocities(double* vel4, const uint32_t* idx, uint32_t n, double alpha);

/* src/BussiReservoirThermostat.h:43-98: one thermostat step given the draws.  Updates
 * reservoir[0] (cumulative translational) and reservoir[1] (instantaneous translational),
 * :86-95.  Returns alpha; dt == 0 returns 1 without touching anything (:45-48).
 * Returns NaN when dof != 0 and KE == 0 (the reference throws, :57-61). */
double orc_bussi_step(double* vel4, const uint32_t* idx, uint32_t n, double dof, double deltaT,
                      double set_T, double tau, double r_normal, double gamma_draw,
                      double reservoir[2], double* ke_out);

/* src/cavitymd/analysis.py:34-47: rho_k = sum_j cos(k.r_j) + i sum_j sin(k.r_j), all particles,
 * wrapped positions.  pos has `stride` doubles per particle (3 or 4).  Long-double accumulation
 * (ground truth for the K x N phase sum; the numpy restatement lives in oracle/oracle.py). */
void orc_rhok(const double* pos, uint32_t stride, uint32_t N, const double* kvec, uint32_t K,
              double* rho_re, double* rho_im);

/* One velocity-Verlet step of the minimal NVE harness used for the 10k-step drift comparison
 * (cavity force only, all particles incl. photon; masses in vel4.w; positions unwrapped, images
 * untouched).  Not a reference function: the harness both arms share. */
/* HOOMD BoxDim::wrap, orthorhombic fully periodic box [-L/2, L/2) (upstream code, restated; see the .c file) */
void orc_wrap(double* p, int32_t* img, const double L[3]);
/* the two harness steps with the drift followed by the box wrap + image update (wrap != 0) */
void orc_nve_step_w(double* pos4, double* vel4, const double* charge, int32_t* image3,
                    double* force4, uint32_t N, double Lx, double Ly, double Lz, uint32_t L_typeid,
                    double omegac, double couplstr, double phmass, double dt, double energies[3], int wrap);
double orc_nvt_step_w(double* pos4, double* vel4, const double* charge, int32_t* image3,
                      double* force4, uint32_t N, double Lx, double Ly, double Lz, uint32_t L_typeid,
                      double omegac, double couplstr, double phmass, double dt, uint32_t first, uint32_t n,
                      double dof, double set_T, double tau, double r_normal, double gamma_draw,
                      double reservoir[2], double* ke_io, double energies[3], int wrap);
void orc_nve_step(double* pos4, double* vel4, const double* charge, const int32_t* image3,
                  double* force4, uint32_t N, double Lx, double Ly, double Lz, uint32_t L_typeid,
                  double omegac, double couplstr, double phmass, double dt, double energies[3]);

/* One step of the thermostatted harness (velocity Verlet with the Bussi rescale folded into the first
 * half step and the kinetic energy for the NEXT step taken while the second half step writes the
 * velocities -- SURVEY.md 8f.1).  Not a reference function and not HOOMD's integrator: the harness both
 * arms share.  *ke_io holds 1/2 sum m|v|^2 of the thermostatted group [first, first+n) on entry (from
 * the previous step, or orc_kinetic_energy before the first) and on exit.  Sequence:
 *   alpha = compute_rescale_factor(ke) (src/BussiReservoirThermostat.h:177-225), reservoir bookkeeping
 *   (:86-95);  group: v <- alpha v;  all: v += dt/2 f/m, r += dt v;  cavity force;  all: v += dt/2 f/m;
 *   ke = 1/2 sum_group m|v|^2.  Returns alpha (NaN if ke == 0 with dof != 0, :57-61). */
double orc_nvt_step(double* pos4, double* vel4, const double* charge, const int32_t* image3,
                    double* force4, uint32_t N, double Lx, double Ly, double Lz, uint32_t L_typeid,
                    double omegac, double couplstr, double phmass, double dt, uint32_t first, uint32_t n,
                    double dof, double set_T, double tau, double r_normal, double gamma_draw,
                    double reservoir[2], double* ke_io, double energies[3]);

#ifdef __cplusplus
}
#endif
#endif
