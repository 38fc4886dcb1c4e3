// Data model behind the opaque `State*` of the C API: State -> Chain -> Spin_System ("image").
// Follows the reference's structure (core/include/data/State.hpp:15-49, Spin_System.hpp:25-82,
// Spin_System_Chain.hpp) as far as the hot path needs it.
//
// Every image keeps AoS host mirrors of its spins and effective field because the C API hands out
// live `scalar*` views into them (System_Get_Spin_Directions / System_Get_Effective_Field,
// core/include/Spirit/System.h:31-43) which callers read and write between calls. The mirrors are
// pinned host memory when a device is present. The device copy (dev::DeviceImage) is created on
// first use and is authoritative only while a method iterates.
#pragma once

#include "geometry.hpp"
#include "hamiltonian.hpp"

#include "../device/runtime.hpp"

#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <vector>

namespace sb
{

// Parameters_Method (+_Solver) : core/include/data/Parameters_Method.hpp, Parameters_Method_Solver.hpp
struct Parameters_Method
{
    long n_iterations          = 1000000;
    long n_iterations_log      = 1000;
    long n_iterations_amortize = 1;
    long max_walltime_sec      = 0;
    double force_convergence   = 1e-10;
    std::string output_folder   = "output";
    std::string output_file_tag = "<time>";
    bool output_any             = false;
    bool output_initial         = false;
    bool output_final           = false;
    int output_vf_filetype      = 3; // IO_Fileformat_OVF_text
    double dt                   = 1e-3;
};

// core/include/data/Parameters_Method_LLG.hpp
struct Parameters_LLG : Parameters_Method
{
    double damping = 0.3;
    double beta    = 0;
    int rng_seed   = 2006;
    std::mt19937 prng{ 2006 };
    // Counter of the device-side Philox thermal-noise stream (one tick per solver iteration). Lives here, like the
    // reference's prng, so that consecutive simulations on the same image continue the stream.
    std::uint64_t philox_counter = 0;
    double temperature = 0;
    Vec3 temperature_gradient_direction{ 1, 0, 0 };
    double temperature_gradient_inclination = 0;
    bool stt_use_gradient                   = true;
    double stt_magnitude                    = 0;
    Vec3 stt_polarisation_normal{ 1, 0, 0 };
    bool direct_minimization                 = false;
    bool output_energy_step                  = false;
    bool output_energy_archive               = false;
    bool output_energy_spin_resolved         = false;
    bool output_energy_divide_by_nspins      = true;
    bool output_energy_add_readability_lines = false;
    bool output_configuration_step           = false;
    bool output_configuration_archive        = false;
};

// core/include/data/Parameters_Method_GNEB.hpp
struct Parameters_GNEB : Parameters_Method
{
    double spring_constant          = 1;
    double spring_force_ratio       = 0;
    double path_shortening_constant = 0;
    int n_E_interpolations          = 10;
    double temperature              = 0;
    int rng_seed                    = 2006;
    bool moving_endpoints           = false;
    bool translating_endpoints      = false;
    double equilibrium_delta_Rx_left  = 1.0;
    double equilibrium_delta_Rx_right = 1.0;
    bool escape_first                 = false;
    bool output_energies_step                  = false;
    bool output_energies_divide_by_nspins      = true;
    bool output_energies_add_readability_lines = false;
    bool output_energies_interpolated          = false;
    bool output_chain_step                     = false;
};

// AoS host field [n][3], pinned when a device is present
class HostField
{
public:
    HostField() = default;
    explicit HostField( std::size_t n );
    HostField( const HostField & other );
    HostField & operator=( const HostField & other );
    ~HostField();

    void resize( std::size_t n );
    std::size_t size() const
    {
        return n_;
    }
    Vec3 * data()
    {
        return data_;
    }
    const Vec3 * data() const
    {
        return data_;
    }
    Vec3 & operator[]( std::size_t i )
    {
        return data_[i];
    }
    const Vec3 & operator[]( std::size_t i ) const
    {
        return data_[i];
    }
    double * scalars()
    {
        return reinterpret_cast<double *>( data_ );
    }

private:
    Vec3 * data_   = nullptr;
    std::size_t n_ = 0;
    bool pinned_   = false;
};

class Method;

struct Spin_System
{
    Spin_System( std::shared_ptr<Hamiltonian> hamiltonian, std::shared_ptr<Geometry> geometry, std::shared_ptr<Parameters_LLG> llg );
    // Deep copy (Spin_System.cpp:38-72): own geometry, Hamiltonian and parameters; iteration_allowed = false
    Spin_System( const Spin_System & other );

    int nos = 0;
    HostField spins;
    HostField effective_field;
    std::shared_ptr<Hamiltonian> hamiltonian;
    std::shared_ptr<Geometry> geometry;
    std::shared_ptr<Parameters_LLG> llg_parameters;

    bool iteration_allowed  = false;
    bool singleshot_allowed = false;

    double E = 0;
    std::vector<std::pair<std::string, double>> E_array;
    Vec3 M{ 0, 0, 0 };

    // Device twin, created on first use. Throws if there is no CUDA device.
    dev::DeviceImage & device();
    bool has_device() const
    {
        return bool( device_ );
    }
    void drop_device()
    {
        if( device_ && effective_field_stale )
            refresh_effective_field_mirror(); // the field of the last run lives only there
        device_.reset();
    }
    // Push host spins / Hamiltonian to the device. While a method iterates on the device-resident spins the device copy is
    // the newer one (device_is_newer): then only the Hamiltonian tables are refreshed, so that a query from another
    // thread (Quantity_Get_*, System_Update_Data) evaluates the LIVE spins instead of rewinding the run to the host copy
    // of the last log step.
    void sync_to_device();
    bool device_is_newer = false; // set by the method's iterations, cleared when the host copy is refreshed (Sync_Host)
    // The host mirror of the effective field is refreshed LAZILY: a finished LLG run leaves the field of its last hook on the
    // device and marks the mirror stale; System_Get_Effective_Field (every call), System_Update_Data and the copy operations
    // bring it up to date. The spins are mirrored eagerly (they are the state); the field is a derived quantity that most
    // callers never read, and mirroring it doubles the device -> host traffic of every Simulation_*_Start call.
    bool effective_field_stale = false;
    void refresh_effective_field_mirror();

    // One-off evaluations on the device (Spin_System.cpp:115-141)
    void UpdateEnergy();
    void UpdateEffectiveField();

    void Lock()
    {
        mutex_.lock();
    }
    void Unlock()
    {
        mutex_.unlock();
    }

private:
    std::unique_ptr<dev::DeviceImage> device_;
    std::mutex mutex_;
};

enum class GNEB_Image_Type
{
    Normal     = 0,
    Climbing   = 1,
    Falling    = 2,
    Stationary = 3
};

struct Chain
{
    int noi = 0;
    std::vector<std::shared_ptr<Spin_System>> images;
    int idx_active_image = 0;
    std::shared_ptr<Parameters_GNEB> gneb_parameters;
    std::vector<GNEB_Image_Type> image_type;
    bool iteration_allowed  = false;
    bool singleshot_allowed = false;

    // Images sharded over several GPUs, one process each (include/spirit_b200.h (3)): this process holds the images
    // [shard_begin, shard_begin + noi) of a chain of shard_noi_global images (-1: not sharded)
    int shard_begin      = 0;
    int shard_noi_global = -1;

    std::vector<double> Rx, Rx_interpolated, E_interpolated;
    std::vector<std::vector<double>> E_array_interpolated;

    // (Re)size Rx and the interpolation arrays to noi (Spin_System_Chain.cpp:20-27, Chain.cpp:744-750)
    void Setup_Interpolation()
    {
        const int n_interp = noi + ( noi - 1 ) * gneb_parameters->n_E_interpolations;
        Rx.assign( noi, 0.0 );
        Rx_interpolated.assign( n_interp, 0.0 );
        E_interpolated.assign( n_interp, 0.0 );
        E_array_interpolated.assign( 7, std::vector<double>( n_interp, 0.0 ) );
    }

    // Locks the chain and all of its images (Spin_System_Chain.cpp:29-55)
    std::mutex mutex_;
    void Lock()
    {
        mutex_.lock();
        for( auto & image : images )
            image->Lock();
    }
    void Unlock()
    {
        for( auto & image : images )
            image->Unlock();
        mutex_.unlock();
    }
};

} // namespace sb

// The opaque struct of the C API (core/include/Spirit/State.h:61)
struct State
{
    std::shared_ptr<sb::Chain> chain;
    std::shared_ptr<sb::Spin_System> active_image;
    std::shared_ptr<sb::Spin_System> clipboard_image;
    std::shared_ptr<std::vector<sb::Vec3>> clipboard_spins;
    int nos              = 0;
    int noi              = 0;
    int idx_active_image = 0;
    std::vector<std::shared_ptr<sb::Method>> method_image;
    std::shared_ptr<sb::Method> method_chain;
    std::chrono::system_clock::time_point datetime_creation = std::chrono::system_clock::now();
    std::string datetime_creation_string;
    std::string config_file;
    bool quiet = false;
};

namespace sb
{
// Resolve (idx_image, idx_chain) the way the reference does (data/State.hpp:79-108): negative image index
// -> active image; index >= noi -> exception (the API layer turns it into a logged no-op).
void from_indices(
    const State * state, int & idx_image, int & idx_chain, std::shared_ptr<Spin_System> & image,
    std::shared_ptr<Chain> & chain );
} // namespace sb
