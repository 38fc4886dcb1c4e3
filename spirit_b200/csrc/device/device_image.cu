// Host-side driver of the device layer: buffers, streams, kernel launches.
#include "device_buffers.cuh"

#include "../core/constants.hpp"
#include "../core/hamiltonian.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <stdexcept>
#include <tuple>

namespace sb
{
namespace dev
{

bool device_available()
{
    int n           = 0;
    cudaError_t err = cudaGetDeviceCount( &n );
    if( err != cudaSuccess )
    {
        cudaGetLastError(); // clear
        return false;
    }
    return n > 0;
}

int device_count()
{
    int n = 0;
    if( cudaGetDeviceCount( &n ) != cudaSuccess )
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void set_device( int device )
{
    require_device();
    SB_CUDA_CHECK( cudaSetDevice( device ) );
}

void require_device()
{
    if( !device_available() )
        throw std::runtime_error(
            "spirit_b200: no CUDA device available. This library has no CPU fallback: the gradient, energy and "
            "solver path runs only as sm_100a CUDA kernels." );
}

std::string device_name()
{
    require_device();
    int d = 0;
    SB_CUDA_CHECK( cudaGetDevice( &d ) );
    cudaDeviceProp prop;
    SB_CUDA_CHECK( cudaGetDeviceProperties( &prop, d ) );
    return prop.name;
}

void * host_alloc( std::size_t bytes, bool & pinned )
{
    pinned = false;
    if( bytes == 0 )
        bytes = 8;
    if( device_available() )
    {
        void * p = nullptr;
        if( cudaHostAlloc( &p, bytes, cudaHostAllocDefault ) == cudaSuccess )
        {
            pinned = true;
            return p;
        }
        cudaGetLastError();
    }
    void * p = nullptr;
    if( posix_memalign( &p, 64, bytes ) != 0 )
        throw std::bad_alloc();
    return p;
}

void host_free( void * ptr, bool pinned )
{
    if( !ptr )
        return;
    if( pinned )
        cudaFreeHost( ptr );
    else
        std::free( ptr );
}

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen) so that the library has no link-time dependency on it: single-GPU use never
// touches it. Prototypes as in nccl.h (2.x ABI).
// ---------------------------------------------------------------------------------------------
namespace
{
struct NcclId
{
    char internal[128];
};
using ncclComm_t = void *;
struct NcclApi
{
    void * lib                                                                                   = nullptr;
    int ( *GetUniqueId )( NcclId * )                                                             = nullptr;
    int ( *CommInitRank )( ncclComm_t *, int, NcclId, int )                                      = nullptr;
    int ( *CommDestroy )( ncclComm_t )                                                           = nullptr;
    int ( *Send )( const void *, std::size_t, int, int, ncclComm_t, cudaStream_t )               = nullptr;
    int ( *Recv )( void *, std::size_t, int, int, ncclComm_t, cudaStream_t )                     = nullptr;
    int ( *AllReduce )( const void *, void *, std::size_t, int, int, ncclComm_t, cudaStream_t ) = nullptr;
    int ( *AllGather )( const void *, void *, std::size_t, int, ncclComm_t, cudaStream_t )      = nullptr;
    int ( *GroupStart )()                                                                        = nullptr;
    int ( *GroupEnd )()                                                                          = nullptr;
    const char * ( *GetErrorString )( int )                                                      = nullptr;
    ncclComm_t comm                                                                              = nullptr;
    int rank = 0, world = 1;
};
NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8, NCCL_INT8 = 0, NCCL_SUM = 0, NCCL_MAX = 2;

void nccl_load()
{
    if( g_nccl.lib )
        return;
    const char * env          = std::getenv( "SPIRIT_B200_NCCL" );
    const char * candidates[] = { env, "libnccl.so.2", "libnccl.so" };
    for( const char * c : candidates )
    {
        if( !c || !*c )
            continue;
        g_nccl.lib = dlopen( c, RTLD_NOW | RTLD_GLOBAL );
        if( g_nccl.lib )
            break;
    }
    if( !g_nccl.lib )
        throw std::runtime_error( "spirit_b200: cannot load libnccl.so.2 (set SPIRIT_B200_NCCL to its path)" );
    auto sym = [&]( const char * name ) {
        void * f = dlsym( g_nccl.lib, name );
        if( !f )
            throw std::runtime_error( std::string( "spirit_b200: libnccl has no symbol " ) + name );
        return f;
    };
    g_nccl.GetUniqueId    = reinterpret_cast<decltype( g_nccl.GetUniqueId )>( sym( "ncclGetUniqueId" ) );
    g_nccl.CommInitRank   = reinterpret_cast<decltype( g_nccl.CommInitRank )>( sym( "ncclCommInitRank" ) );
    g_nccl.CommDestroy    = reinterpret_cast<decltype( g_nccl.CommDestroy )>( sym( "ncclCommDestroy" ) );
    g_nccl.Send           = reinterpret_cast<decltype( g_nccl.Send )>( sym( "ncclSend" ) );
    g_nccl.Recv           = reinterpret_cast<decltype( g_nccl.Recv )>( sym( "ncclRecv" ) );
    g_nccl.AllReduce      = reinterpret_cast<decltype( g_nccl.AllReduce )>( sym( "ncclAllReduce" ) );
    g_nccl.AllGather      = reinterpret_cast<decltype( g_nccl.AllGather )>( sym( "ncclAllGather" ) );
    g_nccl.GroupStart     = reinterpret_cast<decltype( g_nccl.GroupStart )>( sym( "ncclGroupStart" ) );
    g_nccl.GroupEnd       = reinterpret_cast<decltype( g_nccl.GroupEnd )>( sym( "ncclGroupEnd" ) );
    g_nccl.GetErrorString = reinterpret_cast<decltype( g_nccl.GetErrorString )>( sym( "ncclGetErrorString" ) );
}
void nccl_check( int result, const char * what )
{
    if( result != 0 )
        throw std::runtime_error( std::string( "spirit_b200 NCCL error in " ) + what + ": " + ( g_nccl.GetErrorString ? g_nccl.GetErrorString( result ) : "?" ) );
}
} // namespace

void comm_unique_id( char id[128] )
{
    nccl_load();
    NcclId nid;
    nccl_check( g_nccl.GetUniqueId( &nid ), "ncclGetUniqueId" );
    std::memcpy( id, nid.internal, 128 );
}
void comm_init( int rank, int world, const char id[128] )
{
    require_device();
    nccl_load();
    if( g_nccl.comm )
        throw std::runtime_error( "spirit_b200: the communicator is already initialised" );
    NcclId nid;
    std::memcpy( nid.internal, id, 128 );
    nccl_check( g_nccl.CommInitRank( &g_nccl.comm, world, nid, rank ), "ncclCommInitRank" );
    g_nccl.rank  = rank;
    g_nccl.world = world;
}
bool comm_active()
{
    return g_nccl.comm != nullptr;
}
void comm_sendrecv(
    const double * send_low, const double * send_high, double * recv_low, double * recv_high, std::size_t count, int lower, int upper,
    void * stream )
{
    if( !g_nccl.comm )
        throw std::runtime_error( "spirit_b200: communicator not initialised" );
    cudaStream_t st = cudaStream_t( stream );
    nccl_check( g_nccl.GroupStart(), "ncclGroupStart" );
    if( lower >= 0 )
        nccl_check( g_nccl.Send( send_low, count, NCCL_FLOAT64, lower, g_nccl.comm, st ), "ncclSend" );
    if( upper >= 0 )
        nccl_check( g_nccl.Send( send_high, count, NCCL_FLOAT64, upper, g_nccl.comm, st ), "ncclSend" );
    if( upper >= 0 )
        nccl_check( g_nccl.Recv( recv_high, count, NCCL_FLOAT64, upper, g_nccl.comm, st ), "ncclRecv" );
    if( lower >= 0 )
        nccl_check( g_nccl.Recv( recv_low, count, NCCL_FLOAT64, lower, g_nccl.comm, st ), "ncclRecv" );
    nccl_check( g_nccl.GroupEnd(), "ncclGroupEnd" );
}
void comm_group_begin()
{
    if( !g_nccl.comm )
        throw std::runtime_error( "spirit_b200: communicator not initialised" );
    nccl_check( g_nccl.GroupStart(), "ncclGroupStart" );
}
void comm_group_end()
{
    nccl_check( g_nccl.GroupEnd(), "ncclGroupEnd" );
}
void comm_send( const double * data, std::size_t count, int peer, void * stream )
{
    nccl_check( g_nccl.Send( data, count, NCCL_FLOAT64, peer, g_nccl.comm, cudaStream_t( stream ) ), "ncclSend" );
}
void comm_recv( double * data, std::size_t count, int peer, void * stream )
{
    nccl_check( g_nccl.Recv( data, count, NCCL_FLOAT64, peer, g_nccl.comm, cudaStream_t( stream ) ), "ncclRecv" );
}
void comm_allreduce( double * data, std::size_t count, bool max, void * stream )
{
    if( !g_nccl.comm )
        throw std::runtime_error( "spirit_b200: communicator not initialised" );
    nccl_check( g_nccl.AllReduce( data, data, count, NCCL_FLOAT64, max ? NCCL_MAX : NCCL_SUM, g_nccl.comm, cudaStream_t( stream ) ), "ncclAllReduce" );
}
int comm_rank()
{
    return g_nccl.rank;
}
int comm_world()
{
    return g_nccl.world;
}


// ---------------------------------------------------------------------------------------------
// Peer-mapped memory for the ranks of one node (NVLink / NVSwitch): the buffers of the neighbouring ranks are mapped into
// this process with CUDA IPC, kernels store halo planes straight into them, and the ranks keep in step with 32-bit
// counters written / awaited by stream memory operations (no kernel, no NCCL call per iteration).
// The driver entry points come from the runtime (cudaGetDriverEntryPoint): no link-time dependency on libcuda.
// ---------------------------------------------------------------------------------------------
namespace
{
struct StreamMemOps
{
    // CUresult cuStreamWriteValue32( CUstream, CUdeviceptr, cuuint32_t, unsigned flags ), cuStreamWaitValue32 alike
    int ( *write32 )( cudaStream_t, unsigned long long, unsigned, unsigned ) = nullptr;
    int ( *wait32 )( cudaStream_t, unsigned long long, unsigned, unsigned )  = nullptr;
    bool tried = false, ok = false;
};
StreamMemOps g_memops;
constexpr unsigned CU_WAIT_GEQ = 0x0; // CU_STREAM_WAIT_VALUE_GEQ

bool stream_memops_available()
{
    if( !g_memops.tried )
    {
        g_memops.tried = true;
        void *w = nullptr, *q = nullptr;
        cudaDriverEntryPointQueryResult r1, r2;
        if( cudaGetDriverEntryPoint( "cuStreamWriteValue32", &w, cudaEnableDefault, &r1 ) == cudaSuccess && r1 == cudaDriverEntryPointSuccess
            && cudaGetDriverEntryPoint( "cuStreamWaitValue32", &q, cudaEnableDefault, &r2 ) == cudaSuccess && r2 == cudaDriverEntryPointSuccess
            && w && q )
        {
            g_memops.write32 = reinterpret_cast<decltype( g_memops.write32 )>( w );
            g_memops.wait32  = reinterpret_cast<decltype( g_memops.wait32 )>( q );
            g_memops.ok      = true;
        }
        else
            cudaGetLastError();
    }
    return g_memops.ok;
}

// every rank contributes `bytes` bytes; all[r * bytes ...] = what rank r contributed (NCCL all-gather through device staging)
void allgather_bytes( const void * mine, std::size_t bytes, std::vector<char> & all, cudaStream_t stream )
{
    const int world = g_nccl.world;
    char *d_in = nullptr, *d_out = nullptr;
    SB_CUDA_CHECK( cudaMalloc( &d_in, bytes ) );
    SB_CUDA_CHECK( cudaMalloc( &d_out, bytes * world ) );
    SB_CUDA_CHECK( cudaMemcpyAsync( d_in, mine, bytes, cudaMemcpyHostToDevice, stream ) );
    nccl_check( g_nccl.AllGather( d_in, d_out, bytes, NCCL_INT8, g_nccl.comm, stream ), "ncclAllGather" );
    all.resize( bytes * world );
    SB_CUDA_CHECK( cudaMemcpyAsync( all.data(), d_out, bytes * world, cudaMemcpyDeviceToHost, stream ) );
    SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
    cudaFree( d_in );
    cudaFree( d_out );
}
} // namespace

bool peer_memops_available()
{
    return stream_memops_available();
}
void peer_write32( cudaStream_t stream, unsigned * address, unsigned value )
{
    if( !stream_memops_available() || g_memops.write32( stream, reinterpret_cast<unsigned long long>( address ), value, 0 ) != 0 )
        throw std::runtime_error( "spirit_b200: cuStreamWriteValue32 failed" );
}
void peer_wait_geq32( cudaStream_t stream, unsigned * address, unsigned value )
{
    if( !stream_memops_available() || g_memops.wait32( stream, reinterpret_cast<unsigned long long>( address ), value, CU_WAIT_GEQ ) != 0 )
        throw std::runtime_error( "spirit_b200: cuStreamWaitValue32 failed" );
}

bool peer_map_all( void * local, std::vector<void *> & peers, std::vector<void *> & opened, cudaStream_t stream )
{
    const int world = g_nccl.world, rank = g_nccl.rank;
    peers.assign( std::size_t( world ), nullptr );
    struct Record
    {
        cudaIpcMemHandle_t handle;
        int ok;
    } mine{};
    const char * off = std::getenv( "SPIRIT_B200_NO_PEER" );
    mine.ok = ( !( off && off[0] == '1' ) && g_nccl.comm && stream_memops_available() && local
                && cudaIpcGetMemHandle( &mine.handle, local ) == cudaSuccess )
                  ? 1
                  : 0;
    if( !mine.ok )
        cudaGetLastError();
    if( !g_nccl.comm )
        return false;
    std::vector<char> all;
    allgather_bytes( &mine, sizeof( Record ), all, stream );
    const Record * rec = reinterpret_cast<const Record *>( all.data() );
    for( int r = 0; r < world; ++r )
        if( !rec[r].ok )
            return false;
    int mapped = 1;
    for( int r = 0; r < world; ++r )
    {
        if( r == rank )
        {
            peers[r] = local;
            continue;
        }
        void * q = nullptr;
        if( cudaIpcOpenMemHandle( &q, rec[r].handle, cudaIpcMemLazyEnablePeerAccess ) != cudaSuccess )
        {
            cudaGetLastError();
            mapped = 0;
            break;
        }
        opened.push_back( q );
        peers[r] = q;
    }
    allgather_bytes( &mapped, sizeof( int ), all, stream );
    for( int r = 0; r < world; ++r )
        if( !reinterpret_cast<const int *>( all.data() )[r] )
            return false;
    return true;
}

// Collective over all ranks (every rank reaches its first fused iteration on a slab at the same point of the program):
// maps the two configuration buffers and the step counters of the neighbouring ranks. On any failure the image keeps the
// NCCL exchange (peer_state = -1) -- on ALL ranks, because the decision is made from gathered data.
void DeviceImage::slab_peer_setup()
{
    auto & b = *buf_;
    if( b.peer_state != 0 )
        return;
    b.peer_state = -1;
    const char * off = std::getenv( "SPIRIT_B200_NO_PEER" );
    const int world = g_nccl.world, rank = g_nccl.rank;
    struct Record
    {
        cudaIpcMemHandle_t spins, next, flags;
        int nc_local, ok, pid_lo, pid_hi;
    } mine{};
    bool ok = !( off && off[0] == '1' ) && stream_memops_available() && b.spins.allocated() && b.next.allocated();
    if( ok && !b.peer_flags )
    {
        ok = cudaMalloc( &b.peer_flags, 2 * sizeof( unsigned ) ) == cudaSuccess;
        if( ok )
            SB_CUDA_CHECK( cudaMemset( b.peer_flags, 0, 2 * sizeof( unsigned ) ) );
    }
    if( ok && world > 1 )
        ok = cudaIpcGetMemHandle( &mine.spins, b.spins.base ) == cudaSuccess && cudaIpcGetMemHandle( &mine.next, b.next.base ) == cudaSuccess
             && cudaIpcGetMemHandle( &mine.flags, b.peer_flags ) == cudaSuccess;
    if( !ok )
        cudaGetLastError();
    mine.nc_local = stencil_.nc_local;
    mine.ok       = ok ? 1 : 0;
    const bool periodic = stencil_.bc[2] != 0;
    const int lower = rank > 0 ? rank - 1 : ( periodic ? world - 1 : -1 );
    const int upper = rank < world - 1 ? rank + 1 : ( periodic ? 0 : -1 );
    if( world == 1 )
    {
        // one slab that is periodic in c: its own planes wrap around
        if( !ok )
            return;
        b.spins.peer_lo = b.spins.peer_hi = periodic ? b.spins.base : nullptr;
        b.next.peer_lo = b.next.peer_hi = periodic ? b.next.base : nullptr;
        b.peer_lo_nc                    = stencil_.nc_local;
        b.peer_state                    = 1;
        return;
    }
    std::vector<char> all;
    SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
    allgather_bytes( &mine, sizeof( Record ), all, b.stream );
    const Record * rec = reinterpret_cast<const Record *>( all.data() );
    for( int r = 0; r < world; ++r )
        if( !rec[r].ok )
            return; // same decision on every rank
    // open the neighbours' buffers (one mapping per remote allocation, also when both neighbours are the same rank)
    auto open = [&]( const cudaIpcMemHandle_t & h ) -> void * {
        void * q = nullptr;
        if( cudaIpcOpenMemHandle( &q, h, cudaIpcMemLazyEnablePeerAccess ) != cudaSuccess )
        {
            cudaGetLastError();
            return nullptr;
        }
        b.peer_opened.push_back( q );
        return q;
    };
    int mapped = 1;
    double *lo_spins = nullptr, *lo_next = nullptr, *hi_spins = nullptr, *hi_next = nullptr;
    unsigned *lo_flags = nullptr, *hi_flags = nullptr;
    if( lower >= 0 )
    {
        lo_spins = static_cast<double *>( open( rec[lower].spins ) );
        lo_next  = static_cast<double *>( open( rec[lower].next ) );
        lo_flags = static_cast<unsigned *>( open( rec[lower].flags ) );
        mapped   = mapped && lo_spins && lo_next && lo_flags;
    }
    if( upper >= 0 && upper == lower )
        hi_spins = lo_spins, hi_next = lo_next, hi_flags = lo_flags;
    else if( upper >= 0 )
    {
        hi_spins = static_cast<double *>( open( rec[upper].spins ) );
        hi_next  = static_cast<double *>( open( rec[upper].next ) );
        hi_flags = static_cast<unsigned *>( open( rec[upper].flags ) );
        mapped   = mapped && hi_spins && hi_next && hi_flags;
    }
    // second round: did every rank manage to map its neighbours?
    allgather_bytes( &mapped, sizeof( int ), all, b.stream );
    for( int r = 0; r < world; ++r )
        if( !reinterpret_cast<const int *>( all.data() )[r] )
            return;
    b.spins.peer_lo = lo_spins, b.spins.peer_hi = hi_spins;
    b.next.peer_lo = lo_next, b.next.peer_hi = hi_next;
    b.peer_flag_lo = lo_flags ? lo_flags + 1 : nullptr; // I am the upper neighbour of the rank below
    b.peer_flag_hi = hi_flags;                          // and the lower neighbour of the rank above
    b.peer_lo_nc   = lower >= 0 ? rec[lower].nc_local : 0;
    b.peer_step    = 0;
    b.peer_state   = 1;
}

namespace
{
int next_pow2( int v )
{
    int p = 1;
    while( p < v )
        p <<= 1;
    return p;
}

int env_int( const char * name, int fallback )
{
    const char * v = std::getenv( name );
    return ( v && *v ) ? std::atoi( v ) : fallback;
}

// Block shape and march length of the sc6 kernels. BX x BY threads own BX sites of BY rows; the grid's z dimension
// splits c into segments of lc planes. lc trades the two extra planes read per segment (and the pipeline fill of
// the march) against the tail of the last wave of CTAs. (SPIRIT_B200_SC6_BX / _BY / _LC override for tuning runs.)
SC6Geometry make_sc6_geometry( const StencilParams & p, int threads, int ctas_per_sm )
{
    SC6Geometry G;
    int bx = std::min( 128, ( ( p.Na + 31 ) / 32 ) * 32 );
    bx     = std::min( threads, env_int( "SPIRIT_B200_SC6_BX", bx ) );
    int by = std::max( 1, std::min( threads / bx, p.Nb ) );
    by     = std::max( 1, std::min( threads / bx, env_int( "SPIRIT_B200_SC6_BY", by ) ) );
    G.block = dim3( bx, by, 1 );
    const int gx = ( p.Na + bx - 1 ) / bx, gy = ( p.Nb + by - 1 ) / by;
    int n_sm = 148, dev = 0;
    if( cudaGetDevice( &dev ) == cudaSuccess )
        cudaDeviceGetAttribute( &n_sm, cudaDevAttrMultiProcessorCount, dev );
    const double slots = double( n_sm ) * ctas_per_sm;
    int best_lc = p.nc_local;
    double best = 1e300;
    for( int lc : { 4, 8, 16, 32, 64, 128 } )
    {
        if( lc > p.nc_local )
            lc = p.nc_local;
        const int nseg     = ( p.nc_local + lc - 1 ) / lc;
        const double waves = double( gx ) * gy * nseg / slots;
        // 3 plane-steps of fill / extra reads per segment (measured: lc 32 beats 16 beats 8 at 256^3)
        const double cost = ( 1.0 + 3.0 / lc ) * std::ceil( waves ) / waves;
        if( cost < best - 1e-12 )
        {
            best    = cost;
            best_lc = lc;
        }
    }
    G.lc   = std::max( 1, env_int( "SPIRIT_B200_SC6_LC", best_lc ) );
    G.lc   = std::max( 1, env_int( ctas_per_sm > 1 ? "SPIRIT_B200_SC6_LC1" : "SPIRIT_B200_SC6_LC2", G.lc ) ); // per launch shape
    G.grid = dim3( gx, gy, ( p.nc_local + G.lc - 1 ) / G.lc );
    return G;
}

SC6Launch make_sc6_launch( const StencilParams & p )
{
    SC6Launch L;
    L.one_window  = make_sc6_geometry( p, SC6_THREADS_1W, SC6_MINB_1W );
    L.two_windows = make_sc6_geometry( p, SC6_THREADS_2W, SC6_MINB_2W );
    return L;
}

// Tiles and march length of the fused two-stage kernels: one CTA per SM; a c-segment of lc planes costs lc corrector
// plane steps and lc + 2 predictor plane steps (~ 1.6 extra plane steps per segment); short segments fill the last wave.
FusedGeometry make_fused_geometry( const StencilParams & p )
{
    FusedGeometry G;
    const int gx = ( p.Na + FUSED_TX - 1 ) / FUSED_TX, gy = ( p.Nb + FUSED_TY - 1 ) / FUSED_TY;
    int n_sm = 148, dev = 0;
    if( cudaGetDevice( &dev ) == cudaSuccess )
        cudaDeviceGetAttribute( &n_sm, cudaDevAttrMultiProcessorCount, dev );
    int best_lc = p.nc_local;
    double best = 1e300;
    for( int nseg = 1; nseg <= p.nc_local; ++nseg )
    {
        const int lc = ( p.nc_local + nseg - 1 ) / nseg;
        if( lc < 4 && p.nc_local >= 4 )
            break;
        const int segs     = ( p.nc_local + lc - 1 ) / lc;
        const double waves = std::ceil( double( gx ) * gy * segs / n_sm );
        const double cost  = waves * ( lc + 1.6 );
        if( cost < best * ( 1 - 1e-9 ) )
        {
            best    = cost;
            best_lc = lc;
        }
    }
    G.lc   = std::max( 1, env_int( "SPIRIT_B200_FUSED_LC", best_lc ) );
    G.grid = dim3( gx, gy, ( p.nc_local + G.lc - 1 ) / G.lc );
    return G;
}

LaunchGeom make_geom( const StencilParams & p )
{
    LaunchGeom lg;
    lg.rowlen   = p.Na * p.NB;
    lg.rows     = p.Nb * p.nc_local;
    lg.bx       = std::min( 128, std::max( 32, next_pow2( lg.rowlen ) ) );
    lg.by       = BLOCK_THREADS / lg.bx;
    lg.blocks_x = ( lg.rowlen + lg.bx - 1 ) / lg.bx;
    lg.blocks_y = ( lg.rows + lg.by - 1 ) / lg.by;
    return lg;
}
} // namespace

DeviceImage::DeviceImage( const Geometry & g )
{
    require_device();
    nos_ = g.nos;
    if( g.n_cell_atoms > MAX_BASIS )
        throw std::runtime_error( "spirit_b200: more than 8 basis atoms per cell are not supported by the stencil kernels" );
    buf_ = std::make_unique<DeviceBuffers>();

    std::memset( &stencil_, 0, sizeof( stencil_ ) );
    stencil_.Na       = g.n_cells[0];
    stencil_.Nb       = g.n_cells[1];
    stencil_.Nc       = g.n_cells[2];
    stencil_.NB       = g.n_cell_atoms;
    stencil_.c_begin  = 0;
    stencil_.nc_local = g.n_cells[2];
    stencil_.halo     = 0;
    for( int ib = 0; ib < g.n_cell_atoms; ++ib )
        stencil_.mu_s[ib] = g.cell_mu_s[ib];

    auto & b          = *buf_;
    b.lg              = make_geom( stencil_ );
    b.nblocks         = b.lg.blocks_x * b.lg.blocks_y;
    stencil_.plane_stride = ( ( stencil_.Na * stencil_.NB * stencil_.Nb + FIELD_BLOCK - 1 ) / FIELD_BLOCK ) * FIELD_BLOCK;
    b.plane_sites         = stencil_.Na * stencil_.NB * stencil_.Nb;
    b.n_storage           = std::size_t( stencil_.plane_stride ) * ( stencil_.nc_local + 2 * stencil_.halo );
    b.sc6             = make_sc6_launch( stencil_ );
    b.fused           = make_fused_geometry( stencil_ );
    SB_CUDA_CHECK( cudaStreamCreateWithFlags( &b.stream, cudaStreamNonBlocking ) );
    SB_CUDA_CHECK( cudaEventCreate( &b.ev_start ) );
    SB_CUDA_CHECK( cudaEventCreate( &b.ev_stop ) );
    b.spins.allocate( b.n_storage );
    const std::size_t fused_ctas = std::size_t( b.fused.grid.x ) * b.fused.grid.y * b.fused.grid.z;
    SB_CUDA_CHECK( cudaMalloc( &b.partials, ( 4 * std::size_t( b.nblocks ) + 2 * fused_ctas ) * sizeof( double ) ) );
    SB_CUDA_CHECK( cudaMalloc( &b.scalars, 16 * sizeof( double ) ) );
    SB_CUDA_CHECK( cudaMemset( b.scalars, 0, 16 * sizeof( double ) ) );
    SB_CUDA_CHECK( cudaHostAlloc( &b.h_scalars, 16 * sizeof( double ), cudaHostAllocDefault ) );
}

DeviceImage::~DeviceImage()
{
    if( site_flags_dev_ )
        cudaFree( site_flags_dev_ );
    if( ddi_masked_ )
        cudaFree( ddi_masked_ );
    if( ddi_ )
        ddi_plan_destroy( ddi_ );
}

void DeviceImage::set_slab( int c_begin, int Nc_global )
{
    if( !comm_active() )
        throw std::runtime_error( "spirit_b200: set_slab needs an initialised communicator (SpiritB200_Comm_Init)" );
    auto & b = *buf_;
    if( b.F.allocated() || b.pred.allocated() )
        throw std::runtime_error( "spirit_b200: set_slab must be called before the image is used" );
    if( c_begin < 0 || c_begin + stencil_.nc_local > Nc_global )
        throw std::runtime_error( "spirit_b200: slab outside of the global lattice" );
    stencil_.Nc      = Nc_global;
    stencil_.c_begin = c_begin;
    // two halo planes per side: the fused predictor + corrector kernels recompute the predictor of the first halo plane from
    // the second (one exchange per iteration instead of one per stage); slabs thinner than that keep one
    stencil_.halo    = stencil_.nc_local >= 2 ? 2 : 1;
    slab_            = true;
    ham_revision_    = ~std::uint64_t( 0 );
    b.n_storage      = std::size_t( stencil_.plane_stride ) * ( stencil_.nc_local + 2 * stencil_.halo );
    b.spins.allocate( b.n_storage );
    SB_CUDA_CHECK( cudaMemset( b.spins.base, 0, 3 * b.n_storage * sizeof( double ) ) );
    if( !b.comm_stream )
    {
        int prio_low = 0, prio_high = 0;
        SB_CUDA_CHECK( cudaDeviceGetStreamPriorityRange( &prio_low, &prio_high ) );
        SB_CUDA_CHECK( cudaStreamCreateWithPriority( &b.comm_stream, cudaStreamNonBlocking, prio_high ) );
        SB_CUDA_CHECK( cudaStreamCreateWithPriority( &b.bnd_stream, cudaStreamNonBlocking, prio_high ) );
        SB_CUDA_CHECK( cudaEventCreateWithFlags( &b.ev_boundary, cudaEventDisableTiming ) );
        SB_CUDA_CHECK( cudaEventCreateWithFlags( &b.ev_comm, cudaEventDisableTiming ) );
        SB_CUDA_CHECK( cudaEventCreateWithFlags( &b.ev_ready, cudaEventDisableTiming ) );
    }
}

// Send the first / last `halo` owned planes of `field` to the lower / upper neighbour slab and receive their planes into
// the halo planes. Planes are contiguous (3 * plane_stride doubles). Open c: the chain ends have no neighbour.
void DeviceImage::exchange_halo( void * device_field )
{
    exchange_halo_begin( device_field );
    exchange_halo_end();
}

void * DeviceImage::boundary_stream()
{
    auto & b = *buf_;
    SB_CUDA_CHECK( cudaEventRecord( b.ev_ready, b.stream ) );
    SB_CUDA_CHECK( cudaStreamWaitEvent( b.bnd_stream, b.ev_ready, 0 ) );
    return b.bnd_stream;
}

void DeviceImage::exchange_halo_begin( void * device_field, bool from_boundary_stream )
{
    if( !slab_ )
        return;
    auto & b          = *buf_;
    DeviceField & f   = *static_cast<DeviceField *>( device_field );
    const auto & p    = stencil_;
    const int world = g_nccl.world, rank = g_nccl.rank;
    const std::size_t plane = 3 * std::size_t( p.plane_stride );
    const std::size_t count = plane * p.halo;
    const bool periodic     = p.bc[2] != 0;
    const int lower = rank > 0 ? rank - 1 : ( periodic ? world - 1 : -1 );
    const int upper = rank < world - 1 ? rank + 1 : ( periodic ? 0 : -1 );
    double * first_owned = f.base + plane * p.halo;
    double * last_owned  = f.base + plane * p.nc_local; // = halo + nc_local - halo
    double * low_halo    = f.base;
    double * high_halo   = f.base + plane * ( p.halo + p.nc_local );
    // the exchange starts when the work enqueued so far on the image's stream is done
    SB_CUDA_CHECK( cudaEventRecord( b.ev_boundary, from_boundary_stream ? b.bnd_stream : b.stream ) );
    SB_CUDA_CHECK( cudaStreamWaitEvent( b.comm_stream, b.ev_boundary, 0 ) );
    if( world == 1 )
    {
        // a single slab that is periodic in c: its own planes wrap around
        if( periodic )
        {
            SB_CUDA_CHECK( cudaMemcpyAsync( low_halo, last_owned, count * sizeof( double ), cudaMemcpyDeviceToDevice, b.comm_stream ) );
            SB_CUDA_CHECK( cudaMemcpyAsync( high_halo, first_owned, count * sizeof( double ), cudaMemcpyDeviceToDevice, b.comm_stream ) );
        }
    }
    else
    {
        nccl_check( g_nccl.GroupStart(), "ncclGroupStart" );
        if( lower >= 0 )
            nccl_check( g_nccl.Send( first_owned, count, NCCL_FLOAT64, lower, g_nccl.comm, b.comm_stream ), "ncclSend" );
        if( upper >= 0 )
            nccl_check( g_nccl.Send( last_owned, count, NCCL_FLOAT64, upper, g_nccl.comm, b.comm_stream ), "ncclSend" );
        // receive order matters when lower == upper (two slabs, periodic): the peer sends its FIRST planes first, and
        // those are my upper neighbours
        if( upper >= 0 )
            nccl_check( g_nccl.Recv( high_halo, count, NCCL_FLOAT64, upper, g_nccl.comm, b.comm_stream ), "ncclRecv" );
        if( lower >= 0 )
            nccl_check( g_nccl.Recv( low_halo, count, NCCL_FLOAT64, lower, g_nccl.comm, b.comm_stream ), "ncclRecv" );
        nccl_check( g_nccl.GroupEnd(), "ncclGroupEnd" );
    }
    SB_CUDA_CHECK( cudaEventRecord( b.ev_comm, b.comm_stream ) );
}

void DeviceImage::exchange_halo_end()
{
    if( !slab_ )
        return;
    SB_CUDA_CHECK( cudaStreamWaitEvent( buf_->stream, buf_->ev_comm, 0 ) );
}

void DeviceImage::allreduce_scalars( int first, int count, bool max )
{
    if( !slab_ || g_nccl.world == 1 )
        return;
    auto & b = *buf_;
    nccl_check(
        g_nccl.AllReduce( b.scalars + first, b.scalars + first, count, NCCL_FLOAT64, max ? NCCL_MAX : NCCL_SUM, g_nccl.comm, b.stream ),
        "ncclAllReduce" );
}

void DeviceImage::synchronize()
{
    SB_CUDA_CHECK( cudaStreamSynchronize( buf_->stream ) );
}

void DeviceImage::timer_start()
{
    // Slabs: the ranks' streams meet (an all-reduce of one unused scalar) right before the start event. Without it the skew with
    // which the processes leave their host-side barrier (a millisecond or two) sits inside the device-timed region of every rank
    // that waits for a neighbour's first halo planes: a tenth of a 20-iteration measurement.
    allreduce_scalars( 15, 1, true );
    SB_CUDA_CHECK( cudaEventRecord( buf_->ev_start, buf_->stream ) );
}
double DeviceImage::timer_stop()
{
    SB_CUDA_CHECK( cudaEventRecord( buf_->ev_stop, buf_->stream ) );
    SB_CUDA_CHECK( cudaEventSynchronize( buf_->ev_stop ) );
    float ms = 0;
    SB_CUDA_CHECK( cudaEventElapsedTime( &ms, buf_->ev_start, buf_->ev_stop ) );
    return ms;
}

// Build the merged neighbour table and on-site tables from the host Hamiltonian
// Site flags of a lattice with pinned sites or defects: one byte per storage index. With flags the nearest-neighbour
// kernels are not used (their tables are rebuilt: sc6 = 0) and the generic kernels test the flags (stencil.cuh).
void DeviceImage::sync_site_flags( const Geometry & g )
{
    if( g.site_revision == site_revision_ )
        return;
    site_revision_ = g.site_revision;
    ham_revision_  = ~std::uint64_t( 0 ); // the choice of kernels depends on the flags
    auto & b       = *buf_;
    bool any       = false;
    for( unsigned char f : g.site_flags )
        any = any || f != 0;
    if( !any )
    {
        if( site_flags_dev_ )
            SB_CUDA_CHECK( cudaFree( site_flags_dev_ ) );
        site_flags_dev_ = nullptr;
        return;
    }
    if( slab_ )
        throw std::runtime_error( "spirit_b200: pinned sites / defects are not supported on a slab decomposition" );
    std::vector<unsigned char> h( b.n_storage, 0 );
    const int plane_sites = stencil_.Na * stencil_.NB * stencil_.Nb;
    for( int i = 0; i < nos_; ++i )
    {
        const int c = i / plane_sites;
        h[std::size_t( i - c * plane_sites ) + std::size_t( stencil_.plane_stride ) * ( c + stencil_.halo )] = g.site_flags[i];
    }
    if( !site_flags_dev_ )
        SB_CUDA_CHECK( cudaMalloc( &site_flags_dev_, b.n_storage ) );
    SB_CUDA_CHECK( cudaMemcpyAsync( site_flags_dev_, h.data(), b.n_storage, cudaMemcpyHostToDevice, b.stream ) );
    SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
}

void DeviceImage::set_hamiltonian( const Hamiltonian & ham )
{
    sync_site_flags( *ham.geometry );
    if( ham.revision == ham_revision_ )
        return;
    const Geometry & g = *ham.geometry;
    stencil_.site_flags = site_flags_dev_;
    StencilParams & p  = stencil_;
    for( int d = 0; d < 3; ++d )
        p.bc[d] = ham.boundary_conditions[d];
    for( int ib = 0; ib < g.n_cell_atoms; ++ib )
        p.mu_s[ib] = g.cell_mu_s[ib];

    // Merge exchange and DMI pairs with identical (i, j, translations). Pairs that reach further than the
    // lattice are invalid for every site (idx_from_pair, Vectormath.hpp:451-453) and dropped here.
    using Key = std::tuple<int, int, int, int, int>;
    std::map<Key, Neighbour> merged;
    std::vector<Key> order;
    auto entry = [&]( const Pair & pr ) -> Neighbour * {
        const auto & t = pr.translations;
        if( std::abs( t[0] ) > g.n_cells[0] || std::abs( t[1] ) > g.n_cells[1] || std::abs( t[2] ) > p.Nc )
            return nullptr;
        if( slab_ && std::abs( t[2] ) > p.halo )
            throw std::runtime_error( "spirit_b200: pairs reaching further than one plane in c are not supported on a slab decomposition yet" );
        if( pr.i < 0 || pr.i >= g.n_cell_atoms || pr.j < 0 || pr.j >= g.n_cell_atoms )
            return nullptr;
        Key k{ pr.i, pr.j, t[0], t[1], t[2] };
        auto it = merged.find( k );
        if( it == merged.end() )
        {
            Neighbour nb{};
            nb.jb = pr.j;
            nb.da = t[0];
            nb.db = t[1];
            nb.dc = t[2];
            it    = merged.emplace( k, nb ).first;
            order.push_back( k );
        }
        return &it->second;
    };
    for( std::size_t i = 0; i < ham.exchange_pairs.size(); ++i )
        if( Neighbour * nb = entry( ham.exchange_pairs[i] ) )
            nb->J += ham.exchange_magnitudes[i];
    for( std::size_t i = 0; i < ham.dmi_pairs.size(); ++i )
        if( Neighbour * nb = entry( ham.dmi_pairs[i] ) )
        {
            nb->Dx += ham.dmi_magnitudes[i] * ham.dmi_normals[i].x;
            nb->Dy += ham.dmi_magnitudes[i] * ham.dmi_normals[i].y;
            nb->Dz += ham.dmi_magnitudes[i] * ham.dmi_normals[i].z;
        }
    if( order.size() > std::size_t( MAX_NEIGH ) )
        throw std::runtime_error( "spirit_b200: more than 160 distinct neighbour entries are not supported" );
    // group by basis atom i, keeping the reference's pair order inside a group
    p.n_neigh = 0;
    for( int ib = 0; ib < g.n_cell_atoms; ++ib )
    {
        p.neigh_begin[ib] = p.n_neigh;
        for( const Key & k : order )
            if( std::get<0>( k ) == ib )
                p.neigh[p.n_neigh++] = merged[k];
    }
    for( int ib = g.n_cell_atoms; ib <= MAX_BASIS; ++ib )
        p.neigh_begin[ib] = p.n_neigh;

    // Nearest-neighbour structure (sc6.cuh): one basis atom, neighbours only at +-a, +-b, +-c, symmetric J, antisymmetric D
    p.sc6 = 0;
    for( int d = 0; d < 3; ++d )
    {
        p.sc6_axis[d] = p.sc6_dflags[d] = 0;
        p.sc6_J[d] = p.sc6_nJ[d] = 0;
        for( int k = 0; k < 3; ++k )
            p.sc6_D[d][k] = 0;
    }
    if( g.n_cell_atoms == 1 )
    {
        bool ok = true;
        const Neighbour * plus[3]  = { nullptr, nullptr, nullptr };
        const Neighbour * minus[3] = { nullptr, nullptr, nullptr };
        for( int n = 0; n < p.n_neigh && ok; ++n )
        {
            const Neighbour & nb = p.neigh[n];
            const int t[3]       = { nb.da, nb.db, nb.dc };
            int axis = -1, nonzero = 0;
            for( int d = 0; d < 3; ++d )
                if( t[d] != 0 )
                {
                    ++nonzero;
                    axis = d;
                }
            if( nonzero != 1 || std::abs( t[axis] ) != 1 )
                ok = false;
            else
                ( t[axis] > 0 ? plus : minus )[axis] = &nb;
        }
        for( int d = 0; d < 3 && ok; ++d )
        {
            if( !plus[d] && !minus[d] )
                continue;
            if( !plus[d] || !minus[d] || plus[d]->J != minus[d]->J || plus[d]->Dx != -minus[d]->Dx
                || plus[d]->Dy != -minus[d]->Dy || plus[d]->Dz != -minus[d]->Dz )
            {
                ok = false;
                break;
            }
            p.sc6_axis[d]   = 1;
            p.sc6_J[d]      = plus[d]->J;
            p.sc6_nJ[d]     = -plus[d]->J;
            p.sc6_D[d][0]   = plus[d]->Dx;
            p.sc6_D[d][1]   = plus[d]->Dy;
            p.sc6_D[d][2]   = plus[d]->Dz;
            p.sc6_dflags[d] = ( plus[d]->Dx != 0 ? 1 : 0 ) | ( plus[d]->Dy != 0 ? 2 : 0 ) | ( plus[d]->Dz != 0 ? 4 : 0 );
        }
        const char * off = std::getenv( "SPIRIT_B200_GENERIC_STENCIL" ); // tests: force the generic gather kernels
        p.sc6            = ( ok && !( off && off[0] == '1' ) && !site_flags_dev_ ) ? 1 : 0;
    }
    {
        const char * off = std::getenv( "SPIRIT_B200_NO_FUSED" ); // tests / A-B runs: force the two-pass marching kernels
        fused_disabled_  = off && off[0] == '1';
    }

    // Uniaxial anisotropy
    if( ham.anisotropy_indices.size() > std::size_t( MAX_ANISO ) )
        throw std::runtime_error( "spirit_b200: more than 16 anisotropy entries are not supported" );
    p.n_aniso = int( ham.anisotropy_indices.size() );
    for( int i = 0; i < p.n_aniso; ++i )
    {
        p.aniso[i].ib = ham.anisotropy_indices[i];
        p.aniso[i].K  = ham.anisotropy_magnitudes[i];
        p.aniso[i].nx = ham.anisotropy_normals[i].x;
        p.aniso[i].ny = ham.anisotropy_normals[i].y;
        p.aniso[i].nz = ham.anisotropy_normals[i].z;
        p.aniso[i].flags = ( p.aniso[i].nx != 0 ? 1 : 0 ) | ( p.aniso[i].ny != 0 ? 2 : 0 ) | ( p.aniso[i].nz != 0 ? 4 : 0 );
    }
    // Cubic anisotropy (entries for the same atom add up)
    p.has_cubic = !ham.cubic_anisotropy_indices.empty();
    for( int ib = 0; ib < MAX_BASIS; ++ib )
        p.K4[ib] = 0;
    for( std::size_t i = 0; i < ham.cubic_anisotropy_indices.size(); ++i )
        if( ham.cubic_anisotropy_indices[i] >= 0 && ham.cubic_anisotropy_indices[i] < MAX_BASIS )
            p.K4[ham.cubic_anisotropy_indices[i]] += ham.cubic_anisotropy_magnitudes[i];
    // Zeeman
    p.has_zeeman = ham.idx_zeeman >= 0;
    for( int ib = 0; ib < MAX_BASIS; ++ib )
        for( int d = 0; d < 3; ++d )
            p.zeeman[ib][d]
                = ib < g.n_cell_atoms ? g.cell_mu_s[ib] * ham.external_field_magnitude * ham.external_field_normal[d] : 0.0;

    // sc6: on-site quadratic form, start value, and the template SPEC of the marching kernels
    for( int k = 0; k < 6; ++k )
        p.sc6_A[k] = 0;
    for( int i = 0; i < p.n_aniso; ++i )
        if( p.aniso[i].ib == 0 )
        {
            const double n[3] = { p.aniso[i].nx, p.aniso[i].ny, p.aniso[i].nz };
            const double f    = -2.0 * p.aniso[i].K;
            p.sc6_A[0] += f * n[0] * n[0];
            p.sc6_A[1] += f * n[1] * n[1];
            p.sc6_A[2] += f * n[2] * n[2];
            p.sc6_A[3] += f * n[0] * n[1];
            p.sc6_A[4] += f * n[0] * n[2];
            p.sc6_A[5] += f * n[1] * n[2];
        }
    for( int d = 0; d < 3; ++d )
        p.sc6_g0[d] = p.has_zeeman ? -p.zeeman[0][d] : 0.0;
    {
        int spec = 0;
        // a single open plane: the c-neighbours exist in the pair list but never contribute (idx_from_pair rejects them for
        // every site), so the marching kernels need not carry the c-direction at all (2-D systems: configs[0], GNEB images)
        const bool c_dead = !slab_ && p.Nc == 1 && !p.bc[2];
        if( p.sc6_axis[2] && !c_dead )
            spec |= SC6_HAS_C;
        for( int d = 0; d < 3; ++d )
            if( p.sc6_dflags[d] & ~( 1 << d ) )
                spec |= SC6_DMI_GENERAL;
        p.sc6_aniso_full = ( p.sc6_A[3] != 0 || p.sc6_A[4] != 0 || p.sc6_A[5] != 0 ) ? 1 : 0;
        buf_->sc6.spec = spec;
    }

    p.has_ddi = 0;
    if( ddi_ )
    {
        ddi_plan_destroy( ddi_ );
        ddi_ = nullptr;
    }
    if( ham.ddi_method == DDI_Method::FFT )
    {
        ddi_      = ddi_plan_create( ham, p, buf_->stream );
        p.has_ddi = 1;
        if( !buf_->ddi_s.allocated() )
        {
            buf_->ddi_s.allocate( buf_->n_storage );
            buf_->ddi_p.allocate( buf_->n_storage );
        }
    }
    else if( ham.ddi_method != DDI_Method::None )
        throw std::runtime_error( "spirit_b200: only ddi_method none/fft are in scope (SURVEY.md 2.2): the reference's cutoff / direct sums are oracle-side ground truth" );

    p.sc6_extras  = ( p.has_cubic || p.has_ddi || p.sc6_aniso_full ) ? 1 : 0;
    ham_revision_ = ham.revision;
}

// ---------------------------------------------------------------------------------------------
void DeviceImage::upload_spins( const double * host_aos )
{
    auto & b = *buf_;
    if( !b.staging )
        SB_CUDA_CHECK( cudaMalloc( &b.staging, 3 * std::size_t( nos_ ) * sizeof( double ) ) );
    SB_CUDA_CHECK( cudaMemcpyAsync(
        b.staging, host_aos, 3 * std::size_t( nos_ ) * sizeof( double ), cudaMemcpyHostToDevice, b.stream ) );
    k_aos_to_soa<<<( nos_ + BLOCK_THREADS - 1 ) / BLOCK_THREADS, BLOCK_THREADS, 0, b.stream>>>(
        b.staging, b.spins.f(), nos_, b.plane_sites, stencil_.plane_stride, stencil_.halo );
    ++launches_;
    SB_CUDA_CHECK( cudaGetLastError() );
    exchange_halo( &b.spins );
}

static void download_field(
    DeviceBuffers & b, const StencilParams & p, const DeviceField & f, double * host_aos, int nos, double scale,
    std::uint64_t & launches )
{
    if( !b.staging )
        SB_CUDA_CHECK( cudaMalloc( &b.staging, 3 * std::size_t( nos ) * sizeof( double ) ) );
    k_soa_to_aos<<<( nos + BLOCK_THREADS - 1 ) / BLOCK_THREADS, BLOCK_THREADS, 0, b.stream>>>(
        f.c(), b.staging, nos, b.plane_sites, p.plane_stride, p.halo, scale );
    ++launches;
    SB_CUDA_CHECK( cudaGetLastError() );
    SB_CUDA_CHECK( cudaMemcpyAsync(
        host_aos, b.staging, 3 * std::size_t( nos ) * sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
    SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
}

void DeviceImage::download_spins( double * host_aos )
{
    download_field( *buf_, stencil_, buf_->spins, host_aos, nos_, 1.0, launches_ );
}

void DeviceImage::download_effective_field( double * host_aos )
{
    if( !buf_->F.allocated() )
        throw std::runtime_error( "spirit_b200: effective field requested before it was computed" );
    download_field( *buf_, stencil_, effective_field_in_Fv_ ? buf_->Fv : buf_->F, host_aos, nos_, 1.0, launches_ );
}

// ---------------------------------------------------------------------------------------------
static void reduce_sum_to( DeviceBuffers & b, const double * partials, int slot, std::uint64_t & launches )
{
    k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( partials, b.nblocks, b.scalars + slot );
    ++launches;
}

void DeviceImage::gradient_and_energy( double * gradient_host_aos, double * energy )
{
    auto & b = *buf_;
    compute_ddi_gradient( 0 );
    if( !b.scratch.allocated() )
        b.scratch.allocate( b.n_storage );
    SB_DISPATCH_NB(
        ( k_gradient<1, true><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>(
            stencil_, b.lg, b.spins.c(), b.ddi_s.c(), b.scratch.f(), 1.0, b.partials ) ),
        ( k_gradient<0, true><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>(
            stencil_, b.lg, b.spins.c(), b.ddi_s.c(), b.scratch.f(), 1.0, b.partials ) ) );
    reduce_sum_to( b, b.partials, 4, launches_ );
    allreduce_scalars( 4, 1, false );
    SB_CUDA_CHECK( cudaMemcpyAsync( b.h_scalars + 4, b.scalars + 4, sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
    if( gradient_host_aos )
        download_field( b, stencil_, b.scratch, gradient_host_aos, nos_, 1.0, launches_ );
    SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
    if( energy )
        *energy = b.h_scalars[4];
}

void DeviceImage::update_effective_field()
{
    auto & b = *buf_;
    if( !b.F.allocated() )
        b.F.allocate( b.n_storage );
    effective_field_in_Fv_ = false;
    compute_ddi_gradient( 0 );
    SB_DISPATCH_NB(
        ( k_gradient<1, false><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>(
            stencil_, b.lg, b.spins.c(), b.ddi_s.c(), b.F.f(), -1.0, nullptr ) ),
        ( k_gradient<0, false><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>(
            stencil_, b.lg, b.spins.c(), b.ddi_s.c(), b.F.f(), -1.0, nullptr ) ) );
}

int DeviceImage::energy_contributions( const Hamiltonian & ham, double * totals, double * per_spin_host )
{
    auto & b = *buf_;
    set_hamiltonian( ham );
    const int n_terms = int( ham.contribution_names.size() );
    if( n_terms == 0 )
        return 0;
    if( !b.terms )
        SB_CUDA_CHECK( cudaMalloc( &b.terms, 6 * std::size_t( nos_ ) * sizeof( double ) ) );
    compute_ddi_gradient( 0 );
    EnergyTermPointers ptrs{};
    const int idx[6] = { ham.idx_zeeman, ham.idx_anisotropy, ham.idx_cubic_anisotropy, ham.idx_exchange, ham.idx_dmi, ham.idx_ddi };
    for( int t = 0; t < 6; ++t )
        ptrs.term[t] = idx[t] >= 0 ? b.terms + std::size_t( idx[t] ) * nos_ : nullptr;
    SB_DISPATCH_NB(
        ( k_energy_contributions<1><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>( stencil_, b.lg, b.spins.c(), b.ddi_s.c(), ptrs ) ),
        ( k_energy_contributions<0><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>( stencil_, b.lg, b.spins.c(), b.ddi_s.c(), ptrs ) ) );
    std::vector<double> h( static_cast<std::size_t>( n_terms ), 0.0 );
    for( int t = 0; t < n_terms; ++t )
    {
        const double * term = b.terms + std::size_t( t ) * nos_;
        // Simple and deterministic: one CTA folds the whole array (only used for one-off evaluations)
        k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( term, nos_, b.scalars + 9 );
        ++launches_;
        SB_CUDA_CHECK( cudaMemcpyAsync( b.h_scalars + 9, b.scalars + 9, sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
        SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
        h[t] = b.h_scalars[9];
    }
    for( int t = 0; t < n_terms; ++t )
        totals[t] = h[t];
    if( per_spin_host )
    {
        SB_CUDA_CHECK( cudaMemcpyAsync(
            per_spin_host, b.terms, std::size_t( n_terms ) * nos_ * sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
        SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
    }
    return n_terms;
}

void DeviceImage::magnetization( double m[3], bool weighted )
{
    auto & b = *buf_;
    k_magnetization<<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>( stencil_, b.lg, b.spins.c(), b.partials, b.nblocks, weighted ? 1 : 0 );
    ++launches_;
    for( int d = 0; d < 3; ++d )
    {
        k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( b.partials + std::size_t( d ) * b.nblocks, b.nblocks, b.scalars + 6 + d );
        ++launches_;
    }
    SB_CUDA_CHECK( cudaGetLastError() );
    SB_CUDA_CHECK( cudaMemcpyAsync( b.h_scalars + 6, b.scalars + 6, 3 * sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
    SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
    for( int d = 0; d < 3; ++d )
        m[d] = b.h_scalars[6 + d] / double( nos_ );
}

double DeviceImage::topological_charge( int diag, double sign0, double sign1, double * density_host )
{
    auto & b        = *buf_;
    const int cells = stencil_.Na * stencil_.Nb;
    const int nb    = ( cells + BLOCK_THREADS - 1 ) / BLOCK_THREADS;
    if( stencil_.NB != 1 )
        throw std::runtime_error( "spirit_b200: topological charge is implemented for one basis atom" );
    double *density = nullptr, *partials = nullptr;
    SB_CUDA_CHECK( cudaMalloc( &partials, std::size_t( nb ) * sizeof( double ) ) );
    if( density_host )
        SB_CUDA_CHECK( cudaMalloc( &density, 2 * std::size_t( cells ) * sizeof( double ) ) );
    k_topological_charge<<<nb, BLOCK_THREADS, 0, b.stream>>>( stencil_, b.spins.c(), diag, sign0, sign1, density, partials );
    k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( partials, nb, b.scalars + 6 );
    launches_ += 2;
    SB_CUDA_CHECK( cudaGetLastError() );
    SB_CUDA_CHECK( cudaMemcpyAsync( b.h_scalars + 6, b.scalars + 6, sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
    if( density_host )
        SB_CUDA_CHECK( cudaMemcpyAsync( density_host, density, 2 * std::size_t( cells ) * sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
    SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
    cudaFree( partials );
    if( density )
        cudaFree( density );
    return b.h_scalars[6];
}

double DeviceImage::topological_charge_table( int n, const int ( *vertex )[3], const double * sign, double * density_host )
{
    auto & b        = *buf_;
    const int cells = stencil_.Na * stencil_.Nb;
    const int nb    = ( cells + BLOCK_THREADS - 1 ) / BLOCK_THREADS;
    if( n < 1 || n > 16 )
        throw std::runtime_error( "spirit_b200: topological charge: unsupported number of triangles per cell" );
    TopologyTable t{};
    t.n = n;
    for( int k = 0; k < n; ++k )
    {
        for( int c = 0; c < 3; ++c )
            t.vertex[k][c] = vertex[k][c];
        t.sign[k] = sign[k];
    }
    double *density = nullptr, *partials = nullptr;
    SB_CUDA_CHECK( cudaMalloc( &partials, std::size_t( nb ) * sizeof( double ) ) );
    if( density_host )
        SB_CUDA_CHECK( cudaMalloc( &density, std::size_t( n ) * cells * sizeof( double ) ) );
    k_topological_charge_table<<<nb, BLOCK_THREADS, 0, b.stream>>>( stencil_, b.spins.c(), t, density, partials );
    k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( partials, nb, b.scalars + 6 );
    launches_ += 2;
    SB_CUDA_CHECK( cudaGetLastError() );
    SB_CUDA_CHECK( cudaMemcpyAsync( b.h_scalars + 6, b.scalars + 6, sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
    if( density_host )
        SB_CUDA_CHECK( cudaMemcpyAsync( density_host, density, std::size_t( n ) * cells * sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
    SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
    cudaFree( partials );
    if( density )
        cudaFree( density );
    return b.h_scalars[6];
}

// ---------------------------------------------------------------------------------------------
void DeviceImage::ensure_work_fields( int solver )
{
    auto & b = *buf_;
    auto need = [&]( DeviceField & f ) {
        if( !f.allocated() )
            f.allocate( b.n_storage );
    };
    need( b.F );
    need( b.Fv );
    if( solver != Solver_VP )
    {
        need( b.pred );
        need( b.next );
    }
    if( solver == Solver_RK4 )
    {
        need( b.pred2 );
        need( b.acc );
    }
}

void DeviceImage::vp_reset()
{
    vp_initialized_ = false;
    SB_CUDA_CHECK( cudaMemsetAsync( buf_->scalars, 0, 4 * sizeof( double ), buf_->stream ) );
}

int DeviceImage::stencil_variant() const
{
    if( !stencil_.sc6 )
        return 0;
    return 1;
}

// The fused two-stage kernels (sc6_fused.cuh) serve Depondt, Heun and SIB on the nearest-neighbour stencil when nothing
// global has to happen between predictor and corrector (no dipolar field) and the rare per-site terms are absent.
bool DeviceImage::fused_usable( int solver, const LLGParams & l ) const
{
    if( fused_disabled_ || ( slab_ && stencil_.halo < 2 ) )
        return false;
    if( solver != Solver_Depondt && solver != Solver_Heun && solver != Solver_SIB )
        return false;
    const StencilParams & p = stencil_;
    if( !p.sc6 || p.has_ddi || l.has_stt || l.has_tgrad )
        return false;
    if( ( buf_->sc6.spec & SC6_HAS_C ) && p.Nc < 2 )
        return false;
    return true;
}

namespace
{
template<int SOLVER, int STAGE>
void sc6_dispatch( const SC6Launch & sc6, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const StageArgs & a )
{
    if( SOLVER == Solver_Depondt )
        sc6_launch_depondt( STAGE, sc6, stream, p, l, a );
    else if( SOLVER == Solver_Heun )
        sc6_launch_heun( STAGE, sc6, stream, p, l, a );
    else if( SOLVER == Solver_SIB )
        sc6_launch_sib( STAGE, sc6, stream, p, l, a );
    else
        sc6_launch_rk4( STAGE, sc6, stream, p, l, a );
}

// Launches one solver stage that writes the configuration `out` and brings the halo planes of `out` up to date (slab
// decomposition). With the marching kernels the stage is split in two launches: the c-segments at the two slab ends
// first, then -- while their first / last plane travels to the neighbouring ranks on the communication stream -- the
// interior segments.
template<int SOLVER, int STAGE>
void launch_stage(
    bool nb1, bool hook, int nblocks, cudaStream_t stream, const StencilParams & p, const LaunchGeom & lg, const LLGParams & l,
    const StageArgs & a, const SC6Launch & sc6, DeviceImage & image, void * out )
{
    if( p.sc6 && !hook && !l.has_stt && !l.has_tgrad )
    {
        // nearest-neighbour structure: marching kernel (the hook iteration, which also stores F, Fv and reduces the
        // energy, goes through the generic kernel)
        const SC6Geometry & G = SC6Shape<SOLVER, STAGE>::two_windows ? sc6.two_windows : sc6.one_window;
        const int nseg        = int( G.grid.z );
        if( image.is_slab() && nseg >= 3 )
        {
            SC6Launch part         = sc6;
            SC6Geometry & Gp       = SC6Shape<SOLVER, STAGE>::two_windows ? part.two_windows : part.one_window;
            Gp.grid.z              = 2;
            Gp.seg_first           = 0;
            Gp.seg_stride          = nseg - 1;
            // the two end segments on the high-priority boundary stream, concurrently with the interior segments
            sc6_dispatch<SOLVER, STAGE>( part, cudaStream_t( image.boundary_stream() ), p, l, a );
            image.exchange_halo_begin( out, true );
            Gp.grid.z     = nseg - 2;
            Gp.seg_first  = 1;
            Gp.seg_stride = 1;
            sc6_dispatch<SOLVER, STAGE>( part, stream, p, l, a );
            image.exchange_halo_end();
            return;
        }
        sc6_dispatch<SOLVER, STAGE>( sc6, stream, p, l, a );
        image.exchange_halo( out );
        return;
    }
    if( nb1 )
    {
        if( hook )
            k_llg_stage<SOLVER, STAGE, 1, true><<<nblocks, BLOCK_THREADS, 0, stream>>>( p, lg, l, a );
        else
            k_llg_stage<SOLVER, STAGE, 1, false><<<nblocks, BLOCK_THREADS, 0, stream>>>( p, lg, l, a );
    }
    else
    {
        if( hook )
            k_llg_stage<SOLVER, STAGE, 0, true><<<nblocks, BLOCK_THREADS, 0, stream>>>( p, lg, l, a );
        else
            k_llg_stage<SOLVER, STAGE, 0, false><<<nblocks, BLOCK_THREADS, 0, stream>>>( p, lg, l, a );
    }
    image.exchange_halo( out );
}
} // namespace

void DeviceImage::llg_initial_hook( int solver, const LLGParams & llg, HookResult * result )
{
    auto & b = *buf_;
    ensure_work_fields( solver );
    compute_ddi_gradient( 0 );
    SB_DISPATCH_NB(
        ( k_force_and_virtual<1><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>(
            stencil_, b.lg, llg, b.spins.c(), b.ddi_s.c(), b.F.f(), b.Fv.f(), b.partials ) ),
        ( k_force_and_virtual<0><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>(
            stencil_, b.lg, llg, b.spins.c(), b.ddi_s.c(), b.F.f(), b.Fv.f(), b.partials ) ) );
    k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( b.partials, b.nblocks, b.scalars + 4 );
    k_hook<<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>( stencil_, b.lg, b.spins.c(), b.F.f(), b.Fv.c(), b.partials + b.nblocks );
    k_reduce_max<<<1, BLOCK_THREADS, 0, b.stream>>>( b.partials + b.nblocks, b.nblocks, b.scalars + 5 );
    launches_ += 3;
    SB_CUDA_CHECK( cudaGetLastError() );
    allreduce_scalars( 4, 1, false );
    allreduce_scalars( 5, 1, true );
    SB_CUDA_CHECK( cudaMemcpyAsync( b.h_scalars + 4, b.scalars + 4, 2 * sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
    SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
    if( result )
    {
        result->energy     = b.h_scalars[4];
        result->max_torque = std::sqrt( b.h_scalars[5] );
    }
    if( solver == Solver_VP )
    {
        // velocity = 0, F_prev = projected force of the constructor-time hook
        SB_CUDA_CHECK( cudaMemsetAsync( b.scalars, 0, 4 * sizeof( double ), b.stream ) );
        vp_initialized_    = true;
        vp_prev_projected_ = false; // k_hook projected F in place and ratio_prev = 0, so F serves as both
    }
    effective_field_in_Fv_ = false;
}

void DeviceImage::llg_iterate( int solver, LLGParams & llg, int n_iterations, bool hook, HookResult * result )
{
    auto & b = *buf_;
    ensure_work_fields( solver );
    const bool nb1 = stencil_.NB == 1;

    auto mark = [&]() {
        if( stage_events_ )
        {
            cudaEvent_t ev;
            SB_CUDA_CHECK( cudaEventCreate( &ev ) );
            SB_CUDA_CHECK( cudaEventRecord( ev, b.stream ) );
            stage_events_->push_back( ev );
        }
    };

    for( int it = 0; it < n_iterations; ++it )
    {
        const bool hk = hook && ( it == n_iterations - 1 );
        mark();
        StageArgs a{};
        a.s               = b.spins.c();
        a.ddi_s           = b.ddi_s.c();
        a.ddi_sp          = b.ddi_p.c();
        a.F_out           = b.F.f();
        a.Fv_out          = b.Fv.f();
        a.energy_partials = b.partials;
        if( llg.has_thermal && stencil_.sc6 && !b.xi )
            SB_CUDA_CHECK( cudaMalloc( &b.xi, 3 * b.n_storage * sizeof( float ) ) );
        a.xi = b.xi;

        compute_ddi_gradient( 0 );
        if( fused_usable( solver, llg ) )
        {
            // predictor and corrector in one launch: the spins are read once and written once (sc6_fused.cuh)
            const int ncta = int( b.fused.grid.x * b.fused.grid.y * b.fused.grid.z );
            FusedArgs fa{};
            fa.s               = b.spins.c();
            fa.out             = b.next.f();
            fa.F_out           = b.F.f();
            fa.energy_partials = b.partials;
            fa.torque_partials = b.partials + ncta;
            auto launch = [&]( const FusedGeometry & G, cudaStream_t st ) {
                if( solver == Solver_Depondt )
                    sc6_fused_depondt( hk, b.sc6.spec, G, st, stencil_, llg, fa );
                else if( solver == Solver_Heun )
                    sc6_fused_heun( hk, b.sc6.spec, G, st, stencil_, llg, fa );
                else
                    sc6_fused_sib( hk, b.sc6.spec, G, st, stencil_, llg, fa );
                ++launches_;
            };
            const int nseg = int( b.fused.grid.z );
            if( slab_ )
                slab_peer_setup();
            if( slab_ && b.peer_state == 1 )
            {
                // halo exchange inside the kernel: the CTAs at the slab ends store their planes into the neighbours' halo
                // planes (peer-mapped memory). Before the launch: the neighbours have finished the previous iteration (their
                // stores into my halo planes are complete, and they no longer read the buffer I am about to write into).
                const unsigned long long f = reinterpret_cast<unsigned long long>( b.peer_flags );
                if( b.peer_step > 0 && g_nccl.world > 1 )
                {
                    if( b.peer_flag_lo && g_memops.wait32( b.stream, f, b.peer_step, CU_WAIT_GEQ ) != 0 )
                        throw std::runtime_error( "spirit_b200: cuStreamWaitValue32 failed" );
                    if( b.peer_flag_hi && g_memops.wait32( b.stream, f + sizeof( unsigned ), b.peer_step, CU_WAIT_GEQ ) != 0 )
                        throw std::runtime_error( "spirit_b200: cuStreamWaitValue32 failed" );
                }
                fa.peer_lo_out = b.next.peer_lo;
                fa.peer_hi_out = b.next.peer_hi;
                fa.peer_lo_nc  = b.peer_lo_nc;
                launch( b.fused, b.stream );
                ++b.peer_step;
                if( g_nccl.world > 1 )
                {
                    if( b.peer_flag_lo && g_memops.write32( b.stream, reinterpret_cast<unsigned long long>( b.peer_flag_lo ), b.peer_step, 0 ) != 0 )
                        throw std::runtime_error( "spirit_b200: cuStreamWriteValue32 failed" );
                    if( b.peer_flag_hi && g_memops.write32( b.stream, reinterpret_cast<unsigned long long>( b.peer_flag_hi ), b.peer_step, 0 ) != 0 )
                        throw std::runtime_error( "spirit_b200: cuStreamWriteValue32 failed" );
                }
            }
            else if( slab_ && nseg >= 3 )
            {
                // slab: the two c-segments at the slab ends first (high-priority stream); while their first / last two planes
                // travel to the neighbouring ranks the interior segments run
                FusedGeometry part = b.fused;
                part.grid.z        = 2;
                part.seg_first     = 0;
                part.seg_stride    = nseg - 1;
                launch( part, cudaStream_t( boundary_stream() ) );
                exchange_halo_begin( &b.next, true );
                part.grid.z     = nseg - 2;
                part.seg_first  = 1;
                part.seg_stride = 1;
                launch( part, b.stream );
                exchange_halo_end();
            }
            else
            {
                launch( b.fused, b.stream );
                exchange_halo( &b.next );
            }
            mark();
            mark();
            if( hk )
            {
                k_reduce_hook<<<1, BLOCK_THREADS, 0, b.stream>>>( b.partials, b.partials + ncta, ncta, b.scalars + 4 );
                launches_ += 1;
            }
            std::swap( b.spins, b.next );
        }
        else if( solver == Solver_Depondt || solver == Solver_Heun || solver == Solver_SIB )
        {
            a.out = b.pred.f();
            if( solver == Solver_Depondt )
                launch_stage<Solver_Depondt, 1>( nb1, hk, b.nblocks, b.stream, stencil_, b.lg, llg, a, b.sc6, *this, &b.pred );
            else if( solver == Solver_Heun )
                launch_stage<Solver_Heun, 1>( nb1, hk, b.nblocks, b.stream, stencil_, b.lg, llg, a, b.sc6, *this, &b.pred );
            else
                launch_stage<Solver_SIB, 1>( nb1, hk, b.nblocks, b.stream, stencil_, b.lg, llg, a, b.sc6, *this, &b.pred );
            mark();
            compute_ddi_gradient( 1 );
            a.sp  = b.pred.c();
            a.out = b.next.f();
            if( solver == Solver_Depondt )
                launch_stage<Solver_Depondt, 2>( nb1, hk, b.nblocks, b.stream, stencil_, b.lg, llg, a, b.sc6, *this, &b.next );
            else if( solver == Solver_Heun )
                launch_stage<Solver_Heun, 2>( nb1, hk, b.nblocks, b.stream, stencil_, b.lg, llg, a, b.sc6, *this, &b.next );
            else
                launch_stage<Solver_SIB, 2>( nb1, hk, b.nblocks, b.stream, stencil_, b.lg, llg, a, b.sc6, *this, &b.next );
            mark();
            launches_ += 2;
            std::swap( b.spins, b.next );
        }
        else if( solver == Solver_RK4 )
        {
            a.acc = b.acc.f();
            a.out = b.pred.f();
            launch_stage<Solver_RK4, 1>( nb1, hk, b.nblocks, b.stream, stencil_, b.lg, llg, a, b.sc6, *this, &b.pred );
            mark();
            compute_ddi_gradient( 1 );
            a.sp  = b.pred.c();
            a.out = b.pred2.f();
            launch_stage<Solver_RK4, 2>( nb1, hk, b.nblocks, b.stream, stencil_, b.lg, llg, a, b.sc6, *this, &b.pred2 );
            mark();
            compute_ddi_gradient( 2 );
            a.sp  = b.pred2.c();
            a.out = b.pred.f();
            launch_stage<Solver_RK4, 3>( nb1, hk, b.nblocks, b.stream, stencil_, b.lg, llg, a, b.sc6, *this, &b.pred );
            mark();
            compute_ddi_gradient( 1 );
            a.sp  = b.pred.c();
            a.out = b.next.f();
            launch_stage<Solver_RK4, 4>( nb1, hk, b.nblocks, b.stream, stencil_, b.lg, llg, a, b.sc6, *this, &b.next );
            mark();
            launches_ += 4;
            std::swap( b.spins, b.next );
        }
        else if( solver == Solver_VP )
        {
            if( !vp_initialized_ )
                throw std::logic_error( "spirit_b200: VP iteration without the initial force evaluation" );
            double * pp = b.partials;
            double * pn = b.partials + b.nblocks;
            double * pe = b.partials + 2 * std::size_t( b.nblocks );
            const ConstField3 Fprev = vp_prev_projected_ ? b.Fv.c() : b.F.c();
            if( nb1 )
            {
                if( hk )
                    k_vp_a<1, true><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>( stencil_, b.lg, b.spins.c(), b.ddi_s.c(), b.F.f(), Fprev, b.scalars, pp, pn, pe );
                else
                    k_vp_a<1, false><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>( stencil_, b.lg, b.spins.c(), b.ddi_s.c(), b.F.f(), Fprev, b.scalars, pp, pn, pe );
            }
            else
            {
                if( hk )
                    k_vp_a<0, true><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>( stencil_, b.lg, b.spins.c(), b.ddi_s.c(), b.F.f(), Fprev, b.scalars, pp, pn, pe );
                else
                    k_vp_a<0, false><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>( stencil_, b.lg, b.spins.c(), b.ddi_s.c(), b.F.f(), Fprev, b.scalars, pp, pn, pe );
            }
            vp_prev_projected_ = hk;
            k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( pp, b.nblocks, b.scalars + 1 );
            k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( pn, b.nblocks, b.scalars + 2 );
            allreduce_scalars( 1, 2, false ); // projections are sums over ALL slabs
            k_vp_ratio<<<1, 1, 0, b.stream>>>( b.scalars );
            if( hk )
            {
                k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( pe, b.nblocks, b.scalars + 4 );
                k_vp_b<true><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>(
                    stencil_, b.lg, b.spins.f(), b.F.c(), b.Fv.f(), b.scalars, llg.dt, llg.dtg, b.partials + 3 * std::size_t( b.nblocks ) );
                k_reduce_max<<<1, BLOCK_THREADS, 0, b.stream>>>( b.partials + 3 * std::size_t( b.nblocks ), b.nblocks, b.scalars + 5 );
                launches_ += 2;
            }
            else
                k_vp_b<false><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>(
                    stencil_, b.lg, b.spins.f(), b.F.c(), b.Fv.f(), b.scalars, llg.dt, llg.dtg, nullptr );
            exchange_halo( &b.spins );
            launches_ += 5;
        }
        else
            throw std::runtime_error( "spirit_b200: solver id " + std::to_string( solver ) + " is not implemented" );

        if( hk && solver != Solver_VP && !fused_usable( solver, llg ) )
        {
            k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( b.partials, b.nblocks, b.scalars + 4 );
            k_hook<<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>(
                stencil_, b.lg, b.spins.c(), b.F.f(), b.Fv.c(), b.partials + b.nblocks );
            k_reduce_max<<<1, BLOCK_THREADS, 0, b.stream>>>( b.partials + b.nblocks, b.nblocks, b.scalars + 5 );
            launches_ += 3;
        }
        ++llg.iteration;
    }
    if( hook )
    {
        allreduce_scalars( 4, 1, false ); // energy
        allreduce_scalars( 5, 1, true );  // max torque^2
    }
    if( hook )
        effective_field_in_Fv_ = solver == Solver_VP;
    SB_CUDA_CHECK( cudaGetLastError() );
    if( hook )
    {
        SB_CUDA_CHECK( cudaMemcpyAsync( b.h_scalars + 4, b.scalars + 4, 2 * sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
        SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
        if( result )
        {
            result->energy     = b.h_scalars[4];
            result->max_torque = std::sqrt( b.h_scalars[5] );
        }
    }
}

int DeviceImage::llg_profile_stages( int solver, LLGParams & llg, int n_iterations, double * stage_ms, int max_stages )
{
    const int n_stages = solver == Solver_RK4 ? 4 : ( solver == Solver_VP ? 0 : 2 );
    if( n_stages == 0 || n_stages > max_stages )
        throw std::runtime_error( "spirit_b200: stage profiling is available for Depondt, Heun, SIB and RK4" );
    std::vector<void *> events;
    stage_events_ = &events;
    try
    {
        llg_iterate( solver, llg, n_iterations, false, nullptr );
    }
    catch( ... )
    {
        stage_events_ = nullptr;
        throw;
    }
    stage_events_ = nullptr;
    SB_CUDA_CHECK( cudaStreamSynchronize( buf_->stream ) );
    for( int k = 0; k < n_stages; ++k )
        stage_ms[k] = 0;
    const int per_it = n_stages + 1; // one event before stage 1 and one after every stage
    for( int it = 0; it < n_iterations; ++it )
        for( int k = 0; k < n_stages; ++k )
        {
            float ms = 0;
            SB_CUDA_CHECK( cudaEventElapsedTime(
                &ms, cudaEvent_t( events[it * per_it + k] ), cudaEvent_t( events[it * per_it + k + 1] ) ) );
            stage_ms[k] += ms / n_iterations;
        }
    for( void * ev : events )
        cudaEventDestroy( cudaEvent_t( ev ) );
    return n_stages;
}

bool DeviceImage::ddi_gradient_of( const double * configuration_base, double * out_base, void * stream )
{
    if( !stencil_.has_ddi || !ddi_ )
        return false;
    ConstField3 c;
    c.base = configuration_base;
    Field3 o;
    o.base = out_base;
    c.base = ddi_operand( configuration_base, stream );
    launches_ += ddi_gradient( *ddi_, c, o, cudaStream_t( stream ) );
    return true;
}

void DeviceImage::dump_thermal_variates( const LLGParams & llg, std::size_t count, float * host )
{
    auto & b    = *buf_;
    float * dev = nullptr;
    SB_CUDA_CHECK( cudaMalloc( &dev, 3 * count * sizeof( float ) ) );
    k_dump_variates<<<unsigned( ( count + BLOCK_THREADS - 1 ) / BLOCK_THREADS ), BLOCK_THREADS, 0, b.stream>>>( llg, count, dev );
    ++launches_;
    SB_CUDA_CHECK( cudaGetLastError() );
    SB_CUDA_CHECK( cudaMemcpyAsync( host, dev, 3 * count * sizeof( float ), cudaMemcpyDeviceToHost, b.stream ) );
    SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
    cudaFree( dev );
}

void DeviceImage::compute_ddi_gradient( int which )
{
    if( !stencil_.has_ddi || !ddi_ )
        return;
    auto & b = *buf_;
    const DeviceField & conf = which == 0 ? b.spins : ( which == 1 ? b.pred : b.pred2 );
    DeviceField & out        = which == 0 ? b.ddi_s : b.ddi_p;
    ConstField3 operand;
    operand.base = ddi_operand( conf.base, b.stream );
    launches_ += ddi_gradient( *ddi_, operand, out.f(), b.stream );
}

// Sites without a magnetic moment (defects: mu_s = 0, Geometry.cpp:80-81) do not enter the dipolar convolution: its operand
// is a copy of the configuration with those sites zeroed (their own dipolar gradient is dropped in site_gradient).
const double * DeviceImage::ddi_operand( const double * conf, void * stream_v )
{
    if( !site_flags_dev_ )
        return conf;
    auto & b            = *buf_;
    cudaStream_t stream = cudaStream_t( stream_v );
    if( !ddi_masked_ )
        SB_CUDA_CHECK( cudaMalloc( &ddi_masked_, 3 * b.n_storage * sizeof( double ) ) );
    const std::size_t n = 3 * b.n_storage;
    k_mask_moments<<<unsigned( ( n + BLOCK_THREADS - 1 ) / BLOCK_THREADS ), BLOCK_THREADS, 0, stream>>>( conf, ddi_masked_, site_flags_dev_, n );
    ++launches_;
    return ddi_masked_;
}

} // namespace dev
} // namespace sb
