// hoomd_shim/hoomd/md/IntegrationMethodTwoStep.h -- the base-class interface of HOOMD's two-step integration methods
// (constructor (sysdef, group), setDeltaT, virtual integrateStepOne / integrateStepTwo and the protected members a
// method uses), written from how hoomd.md.methods.ConstantVolume is used by the reference
// (examples/05_advanced_run.py:652).  HOOMD-upstream interface, not in the reference tree.  See ../ShimCore.h.
#ifndef HOOMD_SHIM_MD_INTEGRATION_METHOD_TWO_STEP_H
#define HOOMD_SHIM_MD_INTEGRATION_METHOD_TWO_STEP_H
#include "../ShimCore.h"

namespace hoomd::md
    {
class IntegrationMethodTwoStep
    {
    public:
    IntegrationMethodTwoStep(std::shared_ptr<SystemDefinition> sysdef, std::shared_ptr<ParticleGroup> group)
        : m_sysdef(sysdef), m_group(group), m_pdata(sysdef->getParticleData()), m_exec_conf(m_pdata->getExecConf()),
          m_deltaT(0)
        {
        }
    virtual ~IntegrationMethodTwoStep() { }
    virtual void integrateStepOne(uint64_t) { }
    virtual void integrateStepTwo(uint64_t) { }
    void setDeltaT(Scalar deltaT) { m_deltaT = deltaT; }
    Scalar getDeltaT() const { return m_deltaT; }
    std::shared_ptr<ParticleGroup> getGroup() const { return m_group; }
    protected:
    std::shared_ptr<SystemDefinition> m_sysdef;
    std::shared_ptr<ParticleGroup> m_group;
    std::shared_ptr<ParticleData> m_pdata;
    std::shared_ptr<ExecutionConfiguration> m_exec_conf;
    Scalar m_deltaT;
    };
    } // namespace hoomd::md
#endif
