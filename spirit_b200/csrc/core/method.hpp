// Iterative methods: the outer loop of Engine::Method (core/src/engine/Method.cpp:57-113) and the LLG
// and GNEB methods. The host keeps the control flow (stop criteria, amortisation, hooks, histories);
// every Iteration() is a handful of fused kernel launches on the image's stream
// (device/kernels.cuh). Save_Current keeps the histories and, for LLG with llg_output_any, writes the reference's
// spin (OVF) and energy files from the host copies at the log steps.
#pragma once

#include "state.hpp"

#include <chrono>
#include <deque>

namespace sb
{

class Method
{
public:
    Method( std::shared_ptr<Parameters_Method> parameters, int idx_image, int idx_chain );
    virtual ~Method() = default;

    // Method.cpp:57-113
    virtual void Iterate();

    // One solver iteration / the hooks around a block of iterations
    virtual void Iteration( bool hook_follows ) = 0;
    virtual void Hook_Pre_Iteration() {}
    virtual void Hook_Post_Iteration()           = 0;
    virtual void Finalize()                      = 0;
    virtual void Save_Current( bool initial, bool final );
    virtual bool Converged()          = 0;
    virtual bool Iterations_Allowed() = 0;
    virtual std::string Name()        = 0;
    virtual std::string SolverName();
    virtual std::string SolverFullName();
    virtual double get_simulated_time()
    {
        return 0;
    }
    virtual std::vector<double> getTorqueMaxNorm_All()
    {
        return { max_torque };
    }
    virtual void Lock()   = 0;
    virtual void Unlock() = 0;
    // Bring the host mirrors (spins, effective field) up to date with the device
    virtual void Sync_Host() = 0;
    // Push host spins to the device (before single shots, as callers may have written to the live pointer)
    virtual void Sync_Device() = 0;

    bool ContinueIterating();
    bool Walltime_Expired( double seconds ) const;
    double getIterationsPerSecond();
    std::int64_t getWallTime() const;

    int solver = 0;
    long iteration = 0, step = 0;
    long n_iterations = 0, n_iterations_log = 0, n_iterations_amortize = 1, n_log = 0;
    int idx_image, idx_chain;
    std::string starttime; // tag of output file names (llg_output_file_tag <time>)
    double max_torque = 0;
    std::shared_ptr<Parameters_Method> parameters;
    std::vector<int> history_iteration;
    std::vector<double> history_max_torque, history_energy;
    std::chrono::system_clock::time_point t_start, t_last;
    std::deque<std::chrono::system_clock::time_point> t_iterations;
};

// Method_LLG (core/src/engine/Method_LLG.cpp): one image, solvers Depondt / SIB / Heun / RK4 / VP
class Method_LLG : public Method
{
public:
    Method_LLG( std::shared_ptr<Spin_System> system, int solver, int idx_image, int idx_chain );

    void Iteration( bool hook_follows ) override;
    void Hook_Post_Iteration() override;
    void Finalize() override;
    void Save_Current( bool initial, bool final ) override;
    bool Converged() override;
    bool Iterations_Allowed() override
    {
        return system->iteration_allowed;
    }
    std::string Name() override
    {
        return "LLG";
    }
    double get_simulated_time() override
    {
        return picoseconds_passed;
    }
    void Lock() override
    {
        system->Lock();
    }
    void Unlock() override
    {
        system->Unlock();
    }
    void Sync_Host() override;
    void Sync_Device() override;

    // n iterations on spins that already live in HBM (no host<->device copies), one hook at the end. Returns the
    // elapsed milliseconds measured with CUDA events on the image's stream. Used by benchmarks (spirit_b200.h).
    static double Iterate_Device_Resident( const std::shared_ptr<Spin_System> & system, int solver, int n_iterations );
    // Kernel parameter block for `system` (Method_LLG.cpp:65-110,131-226)
    static dev::LLGParams make_params( const Spin_System & system, int solver );

    std::shared_ptr<Spin_System> system;
    double picoseconds_passed = 0;

private:
    dev::LLGParams llg_{};
    dev::HookResult pending_hook_{};
    bool hook_pending_ = false;
    bool force_converged_ = false;
};

} // namespace sb
