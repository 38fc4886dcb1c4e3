// Method_GNEB: geodesic nudged elastic band over the images of a chain
// (core/src/engine/Method_GNEB.cpp:23-456, core/include/engine/Method_GNEB.hpp).
// The host keeps the control flow and the noi-sized scalar logic (Hermite interpolation of E(Rx));
// gradients, tangents, projections, spring / climbing / falling forces and the solver updates run
// as image-batched kernels (device/device_chain.cu).
#pragma once

#include "method.hpp"

namespace sb
{

class Method_GNEB : public Method
{
public:
    Method_GNEB( std::shared_ptr<Chain> chain, int solver, int idx_chain );
    ~Method_GNEB() override;

    void Iteration( bool hook_follows ) override;
    void Hook_Post_Iteration() override;
    void Finalize() override;
    void Save_Current( bool initial, bool final ) override;
    bool Converged() override;
    bool Iterations_Allowed() override
    {
        return chain->iteration_allowed;
    }
    std::string Name() override
    {
        return "GNEB";
    }
    std::vector<double> getTorqueMaxNorm_All() override
    {
        return max_torque_all;
    }
    void Lock() override
    {
        chain->Lock();
    }
    void Unlock() override
    {
        chain->Unlock();
    }
    void Sync_Host() override;
    void Sync_Device() override;

    std::shared_ptr<Chain> chain;
    std::vector<double> max_torque_all;

private:
    std::unique_ptr<dev::DeviceChain> device_;
    dev::ChainHookResult pending_;
    bool hook_pending_ = false;
    bool evaluated_    = false; // at least one force evaluation ran (the effective fields on the device are meaningful)
};

// Cubic Hermite interpolation of p(x) with slopes m (core/src/utility/Cubic_Hermite_Spline.cpp:11-48)
std::vector<std::vector<double>> cubic_hermite_interpolate(
    const std::vector<double> & x, const std::vector<double> & p, const std::vector<double> & m, int n_interpolations );

} // namespace sb
