// ref_driver.cc -- extern "C" driver around the REFERENCE's own translation units.
// TEST INFRASTRUCTURE ONLY (see oracle/cavity_oracle.h for the rules).
//
// oracle/Makefile compiles this file together with /root/reference/src/CavityForceCompute.cc
// (verbatim, by path -- never copied into this repo) against hoomd_shim/ into
// oracle/_ref/libcavref.so.  The reference's header-only src/BussiReservoirThermostat.h (and the
// vendored src/Thermostat.h it includes) are included below, again by path.  The result is the
// reference's arithmetic, operation order included, behind a C ABI that ctypes can call:
//   * to validate oracle/cavity_oracle.c (tests/test_oracle.py),
//   * to mint tests/golden/ fixtures (tests/golden/make_golden.py),
//   * as the timed CPU baseline (bench.py cpu_baseline.kind == "reference").
// What is NOT the reference here: hoomd_shim (stand-in for HOOMD's containers), the injected RNG
// draws, the KE loop of the shim ComputeThermo and the v *= alpha loop below (HOOMD upstream code
// that is not in the reference tree).
#include "CavityForceCompute.h"
#include "BussiReservoirThermostat.h"

#include <cstdio>

using namespace hoomd;

namespace
    {
struct CoutSilencer
    {
    // the reference's constructor prints a banner on std::cout (CavityForceCompute.cc:37-40);
    // bench.py needs stdout clean for its JSON line
    CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) { }
    ~CoutSilencer() { std::cout.rdbuf(old); }
    std::ostringstream sink;
    std::streambuf* old;
    };

std::vector<std::string> make_types(uint32_t L_typeid, uint32_t ntypes)
    {
    std::vector<std::string> t;
    for (uint32_t i = 0; i < ntypes; i++)
        t.push_back(i == L_typeid ? std::string("L") : ("T" + std::to_string(i)));
    return t;
    }

struct RefCavity
    {
    std::shared_ptr<ExecutionConfiguration> exec;
    std::shared_ptr<ParticleData> pdata;
    std::shared_ptr<SystemDefinition> sysdef;
    std::shared_ptr<cavitymd::CavityForceCompute> fc;
    };

struct RefBussi
    {
    std::shared_ptr<ExecutionConfiguration> exec;
    std::shared_ptr<ParticleData> pdata;
    std::shared_ptr<SystemDefinition> sysdef;
    std::shared_ptr<ParticleGroup> group;
    std::shared_ptr<md::ComputeThermo> thermo;
    std::shared_ptr<VariantConstant> kT;
    std::shared_ptr<md::BussiReservoirThermostat> th;
    };
    } // namespace

extern "C"
    {
// ntypes type names are generated; the one at index L_typeid is "L".  Pass L_typeid >= ntypes
// for a system without an 'L' type (getTypeByName then throws, CavityForceCompute.cc:79).
void* ref_cavity_create(uint32_t N, double Lx, double Ly, double Lz, uint32_t L_typeid,
                        uint32_t ntypes, double omegac, double couplstr, double phmass)
    {
    CoutSilencer quiet;
    auto* r = new RefCavity();
    r->exec = std::make_shared<ExecutionConfiguration>(false);
    r->pdata = std::make_shared<ParticleData>(N, BoxDim(Lx, Ly, Lz), make_types(L_typeid, ntypes), r->exec);
    r->sysdef = std::make_shared<SystemDefinition>(r->pdata);
    r->fc = std::make_shared<cavitymd::CavityForceCompute>(r->sysdef, omegac, couplstr, phmass);
    return r;
    }

void ref_cavity_destroy(void* h) { delete static_cast<RefCavity*>(h); }

void ref_cavity_load(void* h, const double* pos4, const double* charge, const int32_t* image3)
    {
    auto* r = static_cast<RefCavity*>(h);
    const unsigned int N = r->pdata->getN();
    ArrayHandle<Scalar4> p(r->pdata->getPositions(), access_location::host, access_mode::overwrite);
    ArrayHandle<Scalar> c(r->pdata->getCharges(), access_location::host, access_mode::overwrite);
    ArrayHandle<int3> im(r->pdata->getImages(), access_location::host, access_mode::overwrite);
    std::memcpy((void*)p.data, pos4, sizeof(double) * 4 * (size_t)N);
    std::memcpy((void*)c.data, charge, sizeof(double) * (size_t)N);
    std::memcpy((void*)im.data, image3, sizeof(int32_t) * 3 * (size_t)N);
    }

// runs ForceCompute::compute() `repeats` times; 0 on success, -2 when the reference threw
int ref_cavity_compute(void* h, uint32_t repeats)
    {
    auto* r = static_cast<RefCavity*>(h);
    try
        {
        for (uint32_t i = 0; i < repeats; i++)
            r->fc->compute(i);
        }
    catch (const std::exception&)
        {
        return -2;
        }
    return 0;
    }

void ref_cavity_read(void* h, double* force4, double energies[3])
    {
    auto* r = static_cast<RefCavity*>(h);
    const unsigned int N = r->pdata->getN();
    if (force4)
        {
        ArrayHandle<Scalar4> f(r->fc->getForceArray(), access_location::host, access_mode::read);
        std::memcpy(force4, (const void*)f.data, sizeof(double) * 4 * (size_t)N);
        }
    energies[0] = r->fc->getHarmonicEnergy();
    energies[1] = r->fc->getCouplingEnergy();
    energies[2] = r->fc->getDipoleSelfEnergy();
    }

// idx == NULL: group = 0..n-1
void* ref_bussi_create(uint32_t N, const double* vel4, const uint32_t* idx, uint32_t n, double dof,
                       double kT, double tau)
    {
    auto* r = new RefBussi();
    r->exec = std::make_shared<ExecutionConfiguration>(false);
    r->pdata = std::make_shared<ParticleData>(N, BoxDim(1, 1, 1), make_types(0, 1), r->exec);
    r->sysdef = std::make_shared<SystemDefinition>(r->pdata);
    std::vector<unsigned int> members(n);
    for (uint32_t j = 0; j < n; j++)
        members[j] = idx ? idx[j] : j;
    r->group = std::make_shared<ParticleGroup>(r->sysdef, members);
    r->group->setTranslationalDOF(dof);
    r->thermo = std::make_shared<md::ComputeThermo>(r->sysdef, r->group);
    r->kT = std::make_shared<VariantConstant>(kT);
    r->th = std::make_shared<md::BussiReservoirThermostat>(r->kT, r->group, r->thermo, r->sysdef, tau);
    ArrayHandle<Scalar4> v(r->pdata->getVelocities(), access_location::host, access_mode::overwrite);
    std::memcpy((void*)v.data, vel4, sizeof(double) * 4 * (size_t)N);
    return r;
    }

void ref_bussi_destroy(void* h) { delete static_cast<RefBussi*>(h); }

// One thermostat call + the rescale HOOMD's step one would apply.  out = {alpha, KE before,
// cumulative translational reservoir, instantaneous translational reservoir}.
// 0 ok, -2 the reference threw (zero kinetic energy, BussiReservoirThermostat.h:57-61).
int ref_bussi_step(void* h, uint64_t timestep, double deltaT, double r_normal, double gamma_draw,
                   double out[4])
    {
    auto* r = static_cast<RefBussi*>(h);
    auto& q = RandomGenerator::injected();
    q.clear();
    q.push_back(r_normal);
    q.push_back(gamma_draw);
    std::array<Scalar, 2> f;
    try
        {
        f = r->th->getRescalingFactorsOne(timestep, deltaT);
        }
    catch (const std::exception&)
        {
        q.clear();
        return -2;
        }
    q.clear();
    out[0] = f[0];
    out[1] = r->thermo->getTranslationalKineticEnergy();
    out[2] = r->th->getReservoirEnergyTranslational();
    out[3] = r->th->getInstantaneousReservoirTranslational();
    // HOOMD upstream stand-in: v <- alpha v over the group
    ArrayHandle<Scalar4> v(r->pdata->getVelocities(), access_location::host, access_mode::readwrite);
    for (unsigned int j = 0; j < r->group->getNumMembers(); j++)
        {
        Scalar4& vv = v.data[r->group->getMemberIndex(j)];
        vv.x *= f[0];
        vv.y *= f[0];
        vv.z *= f[0];
        }
    return 0;
    }

// The same with rotational degrees of freedom (reference src/BussiReservoirThermostat.h:53-55,77-81,87-95): the
// rotational dof / kinetic energy are injected (HOOMD's ComputeThermo forms them upstream); draws = {normal_t, gamma_t,
// normal_r, gamma_r} in the order the reference consumes them.  out = {alpha_t, alpha_r, KE_t, cumulative_t,
// instantaneous_t, cumulative_r, instantaneous_r}.  Velocities are not touched.
int ref_bussi_step_rot(void* h, uint64_t timestep, double deltaT, const double draws[4], double rot_dof, double rot_ke,
                       double out[7])
    {
    auto* r = static_cast<RefBussi*>(h);
    r->group->setRotationalDOF(rot_dof);
    r->thermo->setRotationalKineticEnergy(rot_ke);
    auto& q = RandomGenerator::injected();
    q.clear();
    for (int k = 0; k < 4; k++)
        q.push_back(draws[k]);
    std::array<Scalar, 2> f;
    try
        {
        f = r->th->getRescalingFactorsOne(timestep, deltaT);
        }
    catch (const std::exception&)
        {
        q.clear();
        return -2;
        }
    const int left = (int)q.size();
    q.clear();
    out[0] = f[0];
    out[1] = f[1];
    out[2] = r->thermo->getTranslationalKineticEnergy();
    out[3] = r->th->getReservoirEnergyTranslational();
    out[4] = r->th->getInstantaneousReservoirTranslational();
    out[5] = r->th->getReservoirEnergyRotational();
    out[6] = r->th->getInstantaneousReservoirRotational();
    return left; // draws NOT consumed (dof <= 1 draws no gamma, dof == 0 draws nothing)
    }

void ref_bussi_read(void* h, double* vel4)
    {
    auto* r = static_cast<RefBussi*>(h);
    ArrayHandle<Scalar4> v(r->pdata->getVelocities(), access_location::host, access_mode::read);
    std::memcpy(vel4, (const void*)v.data, sizeof(double) * 4 * (size_t)r->pdata->getN());
    }

void ref_bussi_reset(void* h) { static_cast<RefBussi*>(h)->th->resetReservoirEnergy(); }
    }
