// BussiReservoirThermostat.h -- B200 host class of the Bussi reservoir thermostat.
//
// Drop-in for the reference's class of the same name (reference src/BussiReservoirThermostat.h:22-226):
// same constructor (T variant, group, thermo, sysdef, tau), same pybind name and properties, still a
// hoomd::md::Thermostat whose getRescalingFactorsOne(timestep, deltaT) HOOMD's ConstantVolume method
// calls at the top of step one.
//
// Two modes:
//   fused_rescale = false (default)   the group's kinetic energy comes from ONE launch of
//       cavb200_bussi_ke instead of ComputeThermo::compute (a full thermo pass that also reduces
//       pressure and potential energy); alpha is formed on the host with the reference's own
//       formula (compute_rescale_factor below restates :177-225) from HOOMD's RandomGenerator, and
//       HOOMD's integrator applies it.  Results are the reference's for any HOOMD version.
//   fused_rescale = true              KE reduce, alpha, reservoir bookkeeping AND v <- alpha v happen
//       in one launch (cavb200_bussi) and the method returns {1, 1}, so the integrator does not
//       rescale a second time.  Valid when the integrator rescales BEFORE the half kick
//       (v <- alpha v; v += a dt/2), which is what TwoStepConstantVolume is believed to do
//       (unverified without HOOMD sources) -- hence opt-in.
#ifndef CAVB200_BUSSI_RESERVOIR_THERMOSTAT_H
#define CAVB200_BUSSI_RESERVOIR_THERMOSTAT_H

#include <hoomd/HOOMDMath.h>
#include <hoomd/ParticleGroup.h>
#include <hoomd/RNGIdentifiers.h>
#include <hoomd/RandomNumbers.h>
#include <hoomd/Variant.h>
#include <hoomd/md/ComputeThermo.h>
#include <hoomd/md/Thermostat.h>

#include "CavbHooks.h"

#include <array>
#include <cavb200.h>
#include <cmath>
#include <pybind11/pybind11.h>
#include <stdexcept>
#include <string>

namespace hoomd::md
    {
class PYBIND11_EXPORT BussiReservoirThermostat : public Thermostat, public CavbBussiSource
    {
    public:
    BussiReservoirThermostat(std::shared_ptr<Variant> T, std::shared_ptr<ParticleGroup> group,
                             std::shared_ptr<ComputeThermo> thermo, std::shared_ptr<SystemDefinition> sysdef, Scalar tau)
        : Thermostat(T, group, thermo, sysdef), m_tau(tau), m_handle(nullptr), m_fused(false), m_stale(false)
        {
#ifdef ENABLE_MPI
        // the kinetic energy is reduced over THIS rank's particles only, where ComputeThermo all-reduces
        // (reference :50-55 under domain decomposition): refuse rather than thermostat with a wrong alpha
        if (m_sysdef->isDomainDecomposed())
            throw std::runtime_error("BussiReservoirThermostat (cavb200): domain-decomposed (multi-rank) runs are not supported");
#endif
        int device = 0;
#ifdef ENABLE_HIP
        cudaGetDevice(&device);
#endif
        const int err = cavb200_create(&m_handle, device);
        if (err)
            throw std::runtime_error(std::string("BussiReservoirThermostat: cavb200_create: ") + cavb200_error_string(err));
        resetReservoirEnergy();
        }
    ~BussiReservoirThermostat() override { cavb200_destroy(m_handle); }

    std::array<Scalar, 2> getRescalingFactorsOne(uint64_t timestep, Scalar deltaT) override
        {
        if (deltaT == 0.0) // reference :45-48
            return {1.0, 1.0};

        const Scalar translational_dof = m_group->getTranslationalDOF();
        const Scalar rotational_dof = m_group->getRotationalDOF();

        // Rotational degrees of freedom (anisotropic bodies; 0 for the point particles of every shipped example):
        // the rotational kinetic energy is HOOMD's own quantity (angular momenta, moments of inertia), so it comes
        // from ComputeThermo exactly as in the reference (:50-55) -- only then is the full thermo pass paid for.
        Scalar rotational_kinetic_energy = 0.0;
        if (rotational_dof != 0)
            {
            m_thermo->compute(timestep);
            rotational_kinetic_energy = m_thermo->getRotationalKineticEnergy();
            if (rotational_kinetic_energy == 0) // reference :57-61
                throw std::runtime_error("Bussi thermostat requires non-zero initial momenta.");
            }

        // same generator, same seeding, same draw order as the reference (:63-67, :192-200: translational normal,
        // translational gamma, rotational normal, rotational gamma).  The draws do not depend on the kinetic
        // energies, so they are made before the launch.
        unsigned int instance_id = 0;
        if (m_group->getNumMembersGlobal() > 0)
            instance_id = m_group->getMemberTag(0);
        RandomGenerator rng(Seed(RNGIdentifier::BussiThermostat, timestep, m_sysdef->getSeed()), instance_id);
        const Scalar set_T = m_T->operator()(timestep);
        double r_normal = 0.0, gamma_draw = 0.0;
        draw(rng, translational_dof, r_normal, gamma_draw);
        double r_normal_rot = 0.0, gamma_draw_rot = 0.0;
        draw(rng, rotational_dof, r_normal_rot, gamma_draw_rot);

        Scalar rotational_factor = 1.0;
        if (rotational_dof != 0)
            {
            rotational_factor
                = compute_rescale_factor(rotational_kinetic_energy, rotational_dof, deltaT, set_T, r_normal_rot, gamma_draw_rot);
            const Scalar delta_rot = rotational_kinetic_energy * (1.0 - rotational_factor * rotational_factor); // :87
            m_reservoir_energy_rotational += delta_rot;                                                        // :91
            m_instantaneous_reservoir_rotational = delta_rot;                                                  // :95
            }
        else
            m_instantaneous_reservoir_rotational = 0.0; // :87 with KE = 0

        auto pdata = m_sysdef->getParticleData();
        ArrayHandle<Scalar4> d_vel(pdata->getVelocities(), access_location::device,
                                   m_fused ? access_mode::readwrite : access_mode::read);
        ArrayHandle<unsigned int> d_index(m_group->getIndexArray(), access_location::device, access_mode::read);
        const unsigned int n = m_group->getNumMembers();

        if (m_fused)
            {
            cavb200_bussi_args a = {set_T, m_tau, deltaT, translational_dof, r_normal, gamma_draw};
            check(cavb200_bussi(m_handle, reinterpret_cast<double*>(d_vel.data), d_index.data, 0, n, &a, nullptr));
            m_stale = true; // reservoir energies live on the device until a getter asks
            return {1.0, rotational_factor};
            }

        check(cavb200_bussi_ke(m_handle, reinterpret_cast<const double*>(d_vel.data), d_index.data, 0, n, nullptr));
        double out[5];
        check(cavb200_bussi_read(m_handle, out, nullptr)); // 8 bytes of KE back: alpha is needed on the host now
        const Scalar ke = out[0];
        if (translational_dof != 0 && ke == 0) // reference :57-61
            throw std::runtime_error("Bussi thermostat requires non-zero initial momenta.");
        const Scalar factor = compute_rescale_factor(ke, translational_dof, deltaT, set_T, r_normal, gamma_draw);
        const Scalar delta = ke * (1.0 - factor * factor); // reference :86
        m_host_cumulative += delta; // reference :90
        m_reservoir_energy_translational = m_host_cumulative + m_dev_cumulative;
        m_instantaneous_reservoir_translational = delta;
        return {factor, rotational_factor};
        }

    //! CavbBussiSource: the draws of one step for a fused integration method (TwoStepConstantVolumeCavity), which forms
    //! alpha on the device from the kinetic energy its own step two left there
    cavb200_bussi_args drawBussiArgs(uint64_t timestep, double deltaT) override
        {
        const Scalar translational_dof = m_group->getTranslationalDOF();
        if (m_group->getRotationalDOF() != 0)
            throw std::runtime_error("BussiReservoirThermostat (cavb200): the fused integration method handles point "
                                     "particles only (rotational degrees of freedom present)");
        unsigned int instance_id = 0;
        if (m_group->getNumMembersGlobal() > 0)
            instance_id = m_group->getMemberTag(0);
        RandomGenerator rng(Seed(RNGIdentifier::BussiThermostat, timestep, m_sysdef->getSeed()), instance_id);
        double r_normal = 0.0, gamma_draw = 0.0;
        draw(rng, translational_dof, r_normal, gamma_draw);
        cavb200_bussi_args a = {m_T->operator()(timestep), m_tau, deltaT, translational_dof, r_normal, gamma_draw};
        m_stale = true;
        return a;
        }
    void adoptHandle(cavb200_handle* h) override
        {
        m_read_handle = h ? h : m_handle;
        m_stale = true;
        }

    Scalar getTau() const { return m_tau; }
    void setTau(Scalar tau) { m_tau = tau; }
    bool getFusedRescale() const { return m_fused; }
    void setFusedRescale(bool f) { m_fused = f; }
    //! true: cooperative launches (safe when the GPU is shared with other streams / processes); false (default):
    //! programmatic dependent launches (include/cavb200.h, conventions)
    bool getCooperativeLaunch() const
        {
        int pdl = 1;
        cavb200_get_tuning(m_handle, "pdl", &pdl);
        return pdl == 0;
        }
    void setCooperativeLaunch(bool c) { cavb200_set_tuning(m_handle, "pdl", c ? 0 : 1); }
    unsigned long long getFaultCount() const { return cavb200_fault_count(m_handle); }

    Scalar getReservoirEnergyTranslational() { sync(); return m_reservoir_energy_translational; }
    Scalar getReservoirEnergyRotational() { return m_reservoir_energy_rotational; }
    Scalar getTotalReservoirEnergy() { sync(); return m_reservoir_energy_translational + m_reservoir_energy_rotational; }
    Scalar getInstantaneousReservoirTranslational() { sync(); return m_instantaneous_reservoir_translational; }
    Scalar getInstantaneousReservoirRotational() { return m_instantaneous_reservoir_rotational; }
    Scalar getInstantaneousReservoirTotal()
        {
        sync();
        return m_instantaneous_reservoir_translational + m_instantaneous_reservoir_rotational;
        }
    void resetReservoirEnergy()
        {
        m_reservoir_energy_translational = 0.0;
        m_reservoir_energy_rotational = 0.0;
        m_instantaneous_reservoir_translational = 0.0;
        m_instantaneous_reservoir_rotational = 0.0;
        m_host_cumulative = 0.0;
        m_dev_cumulative = 0.0;
        if (m_handle)
            cavb200_bussi_reset(m_handle, nullptr);
        if (m_read_handle && m_read_handle != m_handle)
            cavb200_bussi_reset(m_read_handle, nullptr);
        m_stale = false;
        }

    protected:
    //! reference :177-225 with the two draws already made (r_gamma = 2 * gamma_draw for dof > 1)
    Scalar compute_rescale_factor(Scalar K, double degrees_of_freedom, Scalar deltaT, Scalar set_T, double r_normal_one,
                                  double gamma_draw) const
        {
        if (degrees_of_freedom == 0)
            return Scalar(1.0);
        double time_decay_factor = 0.0;
        if (m_tau != 0.0)
            time_decay_factor = exp(-deltaT / m_tau);
        double r_gamma = 0.0;
        if (degrees_of_freedom > 1.0)
            r_gamma = 2.0 * gamma_draw;
        const double v = set_T / 2.0 / K;
        const double term1 = v * (1.0 - time_decay_factor) * (r_gamma + r_normal_one * r_normal_one);
        const double term2 = 2.0 * r_normal_one * sqrt(v * (1.0 - time_decay_factor) * time_decay_factor);
        const double alpha_magnitude = sqrt(time_decay_factor + term1 + term2);
        const double c = time_decay_factor;
        const double K_bar = set_T * degrees_of_freedom / 2.0;
        const double sign_term = r_normal_one + sqrt(c * degrees_of_freedom * K / ((1.0 - c) * K_bar));
        return sign_term >= 0.0 ? Scalar(alpha_magnitude) : Scalar(-alpha_magnitude);
        }

    private:
    //! one degree-of-freedom class's draws, in the reference's order (:192-200); nothing is drawn for dof == 0 (:181-184)
    static void draw(RandomGenerator& rng, Scalar dof, double& r_normal, double& gamma_draw)
        {
        if (dof == 0)
            return;
        NormalDistribution<double> normal(1.0);
        r_normal = normal(rng);
        if (dof > 1.0)
            {
            GammaDistribution<double> gamma((dof - 1.0) / Scalar(2.0), Scalar(1.0));
            gamma_draw = gamma(rng);
            }
        }
    void check(int err) const
        {
        if (err == 702) // cudaErrorLaunchTimeout: see include/cavb200.h, conventions
            throw std::runtime_error("BussiReservoirThermostat: an earlier thermostat kernel could not run with its whole grid "
                                     "resident (the GPU is shared with another stream or process), so that step's "
                                     "velocities were not rescaled; the thermostat now uses cooperative launches");
        if (err)
            throw std::runtime_error(std::string("BussiReservoirThermostat: ") + cavb200_error_string(err));
        }
    //! fused mode: pull {instantaneous, cumulative, error flag} off the device when a getter is called
    void sync()
        {
        if (!m_stale)
            return;
        double out[5];
        check(cavb200_bussi_read(m_read_handle ? m_read_handle : m_handle, out, nullptr));
        if (out[4] == 1.0)
            throw std::runtime_error("Bussi thermostat requires non-zero initial momenta.");
        m_instantaneous_reservoir_translational = out[2];
        m_dev_cumulative = out[3];
        m_reservoir_energy_translational = m_host_cumulative + m_dev_cumulative;
        m_stale = false;
        }

    Scalar m_tau;
    cavb200_handle* m_handle;
    cavb200_handle* m_read_handle = nullptr; //!< handle of a fused integration method that adopted this thermostat
    bool m_fused;
    bool m_stale;
    Scalar m_host_cumulative = 0.0; //!< deltas accumulated by the host path
    Scalar m_dev_cumulative = 0.0;  //!< running sum kept by the fused kernel (last read)
    Scalar m_reservoir_energy_translational = 0.0;
    Scalar m_reservoir_energy_rotational = 0.0;
    Scalar m_instantaneous_reservoir_translational = 0.0;
    Scalar m_instantaneous_reservoir_rotational = 0.0;
    };

void export_BussiReservoirThermostat(pybind11::module& m);
    } // namespace hoomd::md
#endif
