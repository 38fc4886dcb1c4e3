// Private to the device layer: error checking and the buffer structs shared by device_image.cu and
// device_chain.cu. Not visible to the host library.
#pragma once

#include "kernels.cuh"
#include "sc6.cuh"
#include "sc6_fused.cuh"
#include "runtime.hpp"

#include <stdexcept>
#include <string>
#include <vector>

namespace sb
{
namespace dev
{

#define SB_CUDA_CHECK( expr )                                                                                          \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t err__ = ( expr );                                                                                  \
        if( err__ != cudaSuccess )                                                                                     \
            throw std::runtime_error(                                                                                  \
                std::string( "spirit_b200 CUDA error: " ) + cudaGetErrorString( err__ ) + " at " + __FILE__ + ":"      \
                + std::to_string( __LINE__ ) + " (" #expr ")" );                                                       \
    } while( 0 )

// one kernel per number of basis atoms known at compile time (1) or not (0); counts the launch
#define SB_DISPATCH_NB( KERNEL_CALL_NB1, KERNEL_CALL_NBX )                                                             \
    do                                                                                                                 \
    {                                                                                                                  \
        if( stencil_.NB == 1 )                                                                                         \
        {                                                                                                              \
            KERNEL_CALL_NB1;                                                                                           \
        }                                                                                                              \
        else                                                                                                           \
        {                                                                                                              \
            KERNEL_CALL_NBX;                                                                                           \
        }                                                                                                              \
        ++launches_;                                                                                                   \
        SB_CUDA_CHECK( cudaGetLastError() );                                                                           \
    } while( 0 )

struct DeviceField
{
    double * base = nullptr;
    std::size_t n = 0; // storage sites (padded planes, including halo planes); a multiple of 32
    // slab decomposition with peer-mapped memory: the same field of the rank below / above in THIS process' address space
    // (cudaIpcOpenMemHandle; travels with the field when two fields are swapped). Null: not mapped / no neighbour.
    double * peer_lo = nullptr;
    double * peer_hi = nullptr;

    void allocate( std::size_t n_ )
    {
        release();
        n = n_;
        SB_CUDA_CHECK( cudaMalloc( &base, 3 * n * sizeof( double ) ) );
    }
    void release()
    {
        if( base )
            cudaFree( base );
        base = nullptr;
        n    = 0;
        peer_lo = peer_hi = nullptr;
    }
    bool allocated() const
    {
        return base != nullptr;
    }
    Field3 f() const
    {
        Field3 r;
        r.base = base;
        return r;
    }
    ConstField3 c() const
    {
        ConstField3 r;
        r.base = base;
        return r;
    }
};

// device/device_image.cu: peer-mapped memory of the other ranks of the node (CUDA IPC over NVLink) and stream-ordered counters
// peer_map_all: COLLECTIVE over all ranks of the communicator. peers[r] = the buffer `local` of rank r in this process'
// address space (peers[rank] = local). Returns false -- on every rank alike -- when the mapping is not available.
bool peer_map_all( void * local, std::vector<void *> & peers, std::vector<void *> & opened, cudaStream_t stream );
bool peer_memops_available();
void peer_write32( cudaStream_t stream, unsigned * address, unsigned value );      // after the work enqueued so far, system-wide fence
void peer_wait_geq32( cudaStream_t stream, unsigned * address, unsigned value );   // the stream waits until *address >= value

// device/ddi_fft.cu
struct DDIPlan;
DDIPlan * ddi_plan_create( const Hamiltonian & ham, const StencilParams & sp, cudaStream_t stream );
void ddi_plan_destroy( DDIPlan * plan );
int ddi_gradient( DDIPlan & plan, ConstField3 spins, Field3 g_ddi, cudaStream_t stream );

struct DeviceBuffers
{
    SC6Launch sc6;
    FusedGeometry fused; // launch shape of the fused two-stage kernels (sc6_fused.cuh)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    // slab decomposition: halo exchanges run on their own stream so that they overlap the interior of a stage
    cudaStream_t comm_stream = nullptr; // NCCL halo exchange (high priority)
    cudaStream_t bnd_stream  = nullptr; // the boundary segments of a stage (high priority), concurrent with the interior
    cudaEvent_t ev_boundary = nullptr, ev_comm = nullptr, ev_ready = nullptr;
    // peer-mapped halo exchange (device_image.cu, slab_peer_setup): step counters written by the neighbouring ranks with
    // stream memory operations. flags[0]: steps finished by the rank below, flags[1]: by the rank above.
    unsigned * peer_flags    = nullptr; // own (cudaMalloc, 2 words)
    unsigned * peer_flag_lo  = nullptr; // flags[1] of the rank below (I am its upper neighbour), peer-mapped
    unsigned * peer_flag_hi  = nullptr; // flags[0] of the rank above
    int peer_state           = 0;       // 0 not tried, 1 ready, -1 unavailable (NCCL exchange stays)
    int peer_lo_nc           = 0;
    unsigned peer_step       = 0;       // fused iterations done with the peer-store exchange
    std::vector<void *> peer_opened;    // cudaIpcOpenMemHandle mappings to close

    LaunchGeom lg{};
    int nblocks            = 0;
    std::size_t n_storage  = 0; // storage sites per field (padded planes, with halo planes)
    int plane_sites        = 0; // real sites per plane: Na*NB*Nb

    DeviceField spins, pred, next; // configuration ping-pong
    DeviceField pred2;             // RK4 second predictor
    DeviceField acc;               // RK4 accumulator
    DeviceField F, Fv;             // force / virtual force (hook, VP); F doubles as the effective field
    DeviceField ddi_s, ddi_p;      // DDI gradient fields of s and of the predictor
    DeviceField scratch;           // gradient output for one-off evaluations
    float * xi = nullptr;          // thermal field of the running iteration as fp32 variates [3 * n_storage] (sc6.cuh)

    double * staging   = nullptr; // AoS staging [nos][3]
    double * partials  = nullptr; // [4][nblocks] reduction scratch
    double * scalars   = nullptr; // device scalars: [0..3] VP, [4] energy, [5] torque^2, [6..8] magnetisation
    double * h_scalars = nullptr; // pinned mirror
    double * terms     = nullptr; // per-term energies [6][nos]

    ~DeviceBuffers()
    {
        for( DeviceField * f : { &spins, &pred, &next, &pred2, &acc, &F, &Fv, &ddi_s, &ddi_p, &scratch } )
            f->release();
        for( void * q : peer_opened )
            cudaIpcCloseMemHandle( q );
        if( peer_flags )
            cudaFree( peer_flags );
        if( staging )
            cudaFree( staging );
        if( xi )
            cudaFree( xi );
        if( partials )
            cudaFree( partials );
        if( scalars )
            cudaFree( scalars );
        if( terms )
            cudaFree( terms );
        if( h_scalars )
            cudaFreeHost( h_scalars );
        if( ev_start )
            cudaEventDestroy( ev_start );
        if( ev_stop )
            cudaEventDestroy( ev_stop );
        if( ev_boundary )
            cudaEventDestroy( ev_boundary );
        if( ev_comm )
            cudaEventDestroy( ev_comm );
        if( ev_ready )
            cudaEventDestroy( ev_ready );
        if( comm_stream )
            cudaStreamDestroy( comm_stream );
        if( bnd_stream )
            cudaStreamDestroy( bnd_stream );
        if( stream )
            cudaStreamDestroy( stream );
    }
};

} // namespace dev
} // namespace sb
