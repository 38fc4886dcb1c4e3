// Host-visible interface of the device layer (no CUDA headers needed by the callers).
//
// A DeviceImage is the HBM-resident state of one spin system ("image"): SoA fp64 spin buffers
// (current / predictor / next), work fields, reduction scratch, and the by-value parameter block
// of its Hamiltonian. All kernels of one image run on the image's own CUDA stream.
//
// There is NO CPU fallback: every entry point calls require_device() and throws if the CUDA
// runtime reports no usable device.
#pragma once

#include "params.hpp"

#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace sb
{
struct Hamiltonian;
struct Geometry;

namespace dev
{

bool device_available();
int device_count();    // 0 when the CUDA runtime reports no usable device
void require_device(); // throws std::runtime_error when there is no CUDA device
void set_device( int device ); // device used by the calling thread for images created afterwards
std::string device_name();

// Pinned host memory when a device is present (full-speed H2D/D2H of the AoS mirrors that the C API
// exposes), plain aligned memory otherwise (host-only use: config parsing, pair lists, tests on CPU).
void * host_alloc( std::size_t bytes, bool & pinned );
void host_free( void * ptr, bool pinned );

// Process-wide NCCL communicator for the slab decomposition (one process per GPU). The 128-byte unique id is created
// on rank 0 (comm_unique_id) and distributed by the launcher (bench.py / tests: torch.distributed broadcast).
void comm_unique_id( char id[128] );
void comm_init( int rank, int world, const char id[128] );
bool comm_active();
// low-level helpers on the process-wide communicator (stream = cudaStream_t): neighbour exchange of `count` doubles
// (send_low -> rank `lower`, send_high -> rank `upper`, receive into recv_low / recv_high; a negative rank = no
// neighbour) and an in-place all-reduce (sum or max) of `count` doubles
void comm_sendrecv( const double * send_low, const double * send_high, double * recv_low, double * recv_high, std::size_t count, int lower, int upper, void * stream );
void comm_allreduce( double * data, std::size_t count, bool max, void * stream );
void comm_group_begin();
void comm_group_end();
void comm_send( const double * data, std::size_t count, int peer, void * stream );
void comm_recv( double * data, std::size_t count, int peer, void * stream );
int comm_rank();
int comm_world();

struct DeviceBuffers; // opaque (device pointers, stream, events)
struct DDIPlan;       // opaque (device/ddi_fft.cu)

struct HookResult
{
    double energy     = 0; // energy of the configuration of the last force evaluation of the iteration
    double max_torque = 0; // max_i |Fv_i - (Fv_i.s_i) s_i| with Fv from the first stage, s the new spins
};

struct OsoState;
struct OsoStateDeleter
{
    void operator()( OsoState * p ) const;
};

class DeviceImage
{
public:
    DeviceImage( const Geometry & geometry );
    ~DeviceImage();
    DeviceImage( const DeviceImage & )             = delete;
    DeviceImage & operator=( const DeviceImage & ) = delete;

    int nos() const
    {
        return nos_;
    }

    // Slab decomposition along c: this image holds the planes [c_begin, c_begin + nc_local) of a lattice with Nc_global
    // planes (nc_local = the geometry's n_cells[2]). Neighbouring slabs live on ranks rank-1 / rank+1 of the process-wide
    // communicator; `halo` planes are exchanged after every kernel that writes a configuration. Must be called before
    // the spins are uploaded.
    void set_slab( int c_begin, int Nc_global );

    // (Re)build the stencil tables from the host Hamiltonian if its revision changed
    void set_hamiltonian( const Hamiltonian & ham );
    const StencilParams & stencil() const
    {
        return stencil_;
    }

    // AoS [nos][3] host <-> SoA device
    void upload_spins( const double * host_aos );
    void download_spins( double * host_aos );
    void download_effective_field( double * host_aos );

    // dipolar gradient field of an arbitrary configuration on this image's lattice (GNEB: one call per image); no-op
    // returning false without a dipolar plan
    bool ddi_gradient_of( const double * configuration_base, double * out_base, void * stream );

    // One-off evaluations on the device-resident spins (System_Update_Data, tests).
    // gradient_host_aos may be null. Hamiltonian_Heisenberg.cpp:670-766
    void gradient_and_energy( double * gradient_host_aos, double * energy );
    // -gradient into the effective-field buffer (Spin_System::UpdateEffectiveField)
    void update_effective_field();
    // Per-term energies: totals[n_terms]; per_spin_host (nullable) [n_terms][nos]. Hamiltonian_Heisenberg.cpp:262-404
    int energy_contributions( const Hamiltonian & ham, double * totals, double * per_spin_host );
    // mean of mu_s * s (Vectormath.cpp:495-502), or the plain mean of s if !weighted (Vectormath.cpp:150-160)
    void magnetization( double m[3], bool weighted );
    // topological charge of the plane c = 0 (one basis atom); density_host (nullable): [2][Na*Nb] charges per triangle
    double topological_charge( int diag, double sign0, double sign1, double * density_host );
    // any basis: n triangles per cell, vertex ids as in k_topological_charge_table; density_host (nullable): [n][Na*Nb]
    double topological_charge_table( int n, const int ( *vertex )[3], const double * sign, double * density_host );

    // n iterations of an LLG solver; if `hook` the last iteration also produces the quantities of
    // Method_LLG::Hook_Post_Iteration (Method_LLG.cpp:246-301) and the effective field buffer.
    // `llg.iteration` is advanced by n.
    void llg_iterate( int solver, LLGParams & llg, int n_iterations, bool hook, HookResult * result );
    // Same iterations with a CUDA event recorded between the stage kernels: stage_ms[k] receives the mean duration
    // (milliseconds) of stage k+1 over the n iterations (bench.py's per-kernel roofline). Returns the number of stages.
    int llg_profile_stages( int solver, LLGParams & llg, int n_iterations, double * stage_ms, int max_stages );
    // The constructor-time evaluation of Method_LLG (Method_LLG.cpp:57-62): force, virtual force, hook
    void llg_initial_hook( int solver, const LLGParams & llg, HookResult * result );
    // VP keeps velocity / previous force between iterations (Solver_VP.hpp:29-114)
    void vp_reset();
    // n iterations of VP_OSO / LBFGS_OSO (device_oso.cu; Solver_VP_OSO.hpp:34-115, Solver_LBFGS_OSO.hpp:39-77). `llg` must
    // carry the prefactor of the solver's virtual force (Method_LLG.cpp:163-171). oso_reset drops velocity / memory.
    void oso_iterate( int solver, LLGParams & llg, int n_iterations, bool hook, HookResult * result );
    void oso_reset();

    void synchronize();
    // test probe: 3 * count unit normal variates of the thermal field's generator (Philox counters 0 .. count - 1 of
    // iteration llg.iteration, the image's seed), as the stage kernels draw them
    void dump_thermal_variates( const LLGParams & llg, std::size_t count, float * host );

    // Timing of the enqueued work on this image's stream (CUDA events), milliseconds
    void timer_start();
    double timer_stop();

    std::uint64_t kernel_launches() const
    {
        return launches_;
    }

    // Which stage kernels serve this image: 1 nearest-neighbour marching kernels (sc6.cuh), 0 generic gather kernels.
    // Valid after the device tables are built.
    int stencil_variant() const;
    // true: Depondt / Heun / SIB iterations of this image run as ONE fused predictor + corrector kernel (sc6_fused.cuh)
    bool fused_usable( int solver, const LLGParams & l ) const;

    DeviceBuffers * buffers()
    {
        return buf_.get();
    }

    bool is_slab() const
    {
        return slab_;
    }
    void exchange_halo( void * device_field ); // DeviceField *; ordered after everything enqueued on the image's stream
    // the two halves of an overlapped exchange: start after the work enqueued so far (the boundary segments of a stage),
    // and make the image's stream wait for its completion
    void exchange_halo_begin( void * device_field, bool from_boundary_stream = false );
    void exchange_halo_end();
    // stream for the boundary segments of a stage; on return it waits for everything enqueued so far on the image's stream
    void * boundary_stream();

private:
    void ensure_work_fields( int solver );
    void slab_peer_setup(); // maps the neighbouring ranks' configuration buffers (CUDA IPC) for the in-kernel halo exchange
    void allreduce_scalars( int first, int count, bool max );
    // DDI gradient field of a configuration (0: spins -> ddi_s, 1: pred -> ddi_p, 2: pred2 -> ddi_p); no-op without DDI
    void compute_ddi_gradient( int which_config );

    int nos_ = 0;
    StencilParams stencil_{};
    std::uint64_t ham_revision_ = ~std::uint64_t( 0 );
    // pinned sites / defects (Geometry::site_flags) in storage order on the device; follows Geometry::site_revision
    void sync_site_flags( const Geometry & g );
    const double * ddi_operand( const double * conf_base, void * stream );
    unsigned char * site_flags_dev_ = nullptr;
    double * ddi_masked_            = nullptr; // configuration with the sites without moment zeroed: operand of the dipolar convolution
    std::uint64_t site_revision_    = 0;
    std::unique_ptr<DeviceBuffers> buf_;
    DDIPlan * ddi_ = nullptr; // owned; created by set_hamiltonian when ddi_method == fft
    std::uint64_t launches_ = 0;
    bool slab_                  = false;
    bool vp_initialized_        = false;
    bool vp_prev_projected_     = false; // the last VP iteration ran a hook: F_prev is the projected force (in Fv)
    bool effective_field_in_Fv_ = false;
    bool fused_disabled_        = false; // SPIRIT_B200_NO_FUSED=1 when the device tables were built
    std::vector<void *> * stage_events_ = nullptr; // when set, llg_iterate records an event after every stage kernel
    std::unique_ptr<OsoState, OsoStateDeleter> oso_; // fields and scalars of the OSO minimisers (device_oso.cu)
};

// GNEB parameters the device needs per force evaluation (Method_GNEB.cpp:87-258)
struct GNEBParams
{
    double spring_constant = 1;
    double dt              = 1e-3; // llg_dt of the images (VP uses it raw, Solver_VP.hpp:95-110)
    double dtg             = 0;    // dt * gamma / mu_B: Fv = dtg s x F (Method_GNEB.cpp:359-391)
    std::vector<int> image_type;   // 0 normal, 1 climbing, 2 falling, 3 stationary
    // Method_GNEB.cpp:137-170, 243-245: with a ratio > 0 the spring force equalises path lengths in the (Rx, E) plane
    double spring_force_ratio = 0;
    // :206-233: > 0 adds a force that shortens the path orthogonally to the gradient force and the tangent
    double path_shortening_constant = 0;
    // :261-355: the end images move too: rotational part of the gradient force, a spring that keeps their distance to the
    // neighbouring image, optionally a common translation; escape_first switches the rotational part off while the energy
    // rises along the tangent at the left end faster than at the right one
    bool moving_endpoints = false, translating_endpoints = false, escape_first = false;
    double equilibrium_delta_Rx_left = 1, equilibrium_delta_Rx_right = 1;
};

struct ChainHookResult
{
    bool degenerate = false; // two neighbouring images coincide (Method_GNEB.cpp:119-126)
    std::vector<double> energy, Rx, max_torque, dE_dRx; // of the local images
    double max_torque_chain = 0;                         // over ALL images of the (possibly sharded) chain
};

// HBM-resident state of a whole chain of images (GNEB): every field is one contiguous allocation [noi][field] and the
// kernels are batched over images (blockIdx.y = image). All images share one Hamiltonian.
struct DeviceChainBuffers;
class DeviceChain
{
public:
    // noi: images held by this process. Sharded over several GPUs (one process each): they are the images
    // [i_begin, i_begin + noi) of a chain of noi_global images; neighbouring ranks hold the neighbouring images.
    DeviceChain( const Geometry & geometry, int noi, int i_begin = 0, int noi_global = -1 );
    ~DeviceChain();
    DeviceChain( const DeviceChain & )             = delete;
    DeviceChain & operator=( const DeviceChain & ) = delete;

    int noi() const
    {
        return noi_;
    }
    void set_hamiltonian( const Hamiltonian & ham );
    void upload_image( int img, const double * host_aos );
    void download_image( int img, double * host_aos );
    // effective field (-gradient, unprojected) of image img from the last force evaluation
    void download_effective_field( int img, double * host_aos );

    // n iterations of solver VP / SIB / Depondt / Heun / VP_OSO / LBFGS_OSO / LBFGS_Atlas over all images; with `hook` the last iteration also produces
    // the quantities of Method_GNEB::Hook_Post_Iteration (Method_GNEB.cpp:410-456)
    void iterate( int solver, const GNEBParams & params, int n_iterations, bool hook, ChainHookResult * result );
    void vp_reset(); // also drops the velocity / L-BFGS memory of the OSO / atlas solvers
    void synchronize();
    std::uint64_t kernel_launches() const
    {
        return launches_;
    }

private:
    void evaluate_force( const GNEBParams & params, int which_configuration, int which_force );
    void exchange_halo_images( double * field_base );
    void share_slots( int first_slot, int n_slots, bool max );
    void reduce_to_slots( int first_slot, int n_slots, bool max );

    void ensure_ddi_field();
    int noi_ = 0, nos_ = 0, i_begin_ = 0, noi_global_ = 0;
    bool sharded_ = false;
    std::unique_ptr<DeviceImage> table_; // owns the stencil tables (shared Hamiltonian)
    std::unique_ptr<DeviceChainBuffers> buf_;
    std::uint64_t launches_  = 0;
    bool vp_prev_projected_ = false;
    std::unique_ptr<OsoState, OsoStateDeleter> oso_; // VP_OSO / LBFGS_OSO / LBFGS_Atlas over the chain (oso.cuh)
};

} // namespace dev
} // namespace sb
