// hoomd_shim/hoomd/md/Thermostat.h -- the Thermostat base-class interface the plugin glue derives
// from (getRescalingFactorsOne/Two, advanceThermostat, getT/setT and the four protected members),
// written from its call sites in the reference's src/BussiReservoirThermostat.h.  See ../ShimCore.h.
#ifndef HOOMD_SHIM_MD_THERMOSTAT_H
#define HOOMD_SHIM_MD_THERMOSTAT_H
#include "../ShimCore.h"
#include <array>

namespace hoomd::md
    {
class Thermostat
    {
    public:
    Thermostat(std::shared_ptr<Variant> T, std::shared_ptr<ParticleGroup> group, std::shared_ptr<ComputeThermo> thermo,
               std::shared_ptr<SystemDefinition> sysdef)
        : m_group(group), m_thermo(thermo), m_T(T), m_sysdef(sysdef)
        {
        }
    virtual ~Thermostat() { }
    virtual std::array<Scalar, 2> getRescalingFactorsOne(uint64_t, Scalar) { return {Scalar(1.0), Scalar(1.0)}; }
    virtual std::array<Scalar, 2> getRescalingFactorsTwo(uint64_t, Scalar) { return {Scalar(1.0), Scalar(1.0)}; }
    virtual void advanceThermostat(uint64_t, Scalar, bool) { }
    std::shared_ptr<Variant> getT() { return m_T; }
    void setT(std::shared_ptr<Variant> T) { m_T = T; }
    protected:
    std::shared_ptr<ParticleGroup> m_group;
    std::shared_ptr<ComputeThermo> m_thermo;
    std::shared_ptr<Variant> m_T;
    std::shared_ptr<SystemDefinition> m_sysdef;
    };
    } // namespace hoomd::md
#endif
