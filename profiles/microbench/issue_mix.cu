// Issue-slot microbenchmark (B200): how many warp instructions per cycle does one SM sub-partition sustain for
// mixes of fp64 FMAs and integer ALU instructions? Answers whether an fp64 instruction (half rate: 16 lanes per
// sub-partition) also occupies the issue port for two cycles, i.e. whether "issue active" can reach 100 % in a kernel
// with 35 % fp64 instructions. Build: nvcc -arch=sm_100a -O3 -o issue_mix issue_mix.cu ; run: ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>

template<int ND, int NI>
__global__ void k_mix( double * out, unsigned * outi, long long * cycles, int iters )
{
    double a0 = threadIdx.x, a1 = 1.0 + threadIdx.x, a2 = 2.0, a3 = 3.0, a4 = 4.0, a5 = 5.0, a6 = 6.0, a7 = 7.0;
    unsigned i0 = threadIdx.x, i1 = 1, i2 = 2, i3 = 3, i4 = 4, i5 = 5, i6 = 6, i7 = 7;
    const double m = 1.0000001, c = 1e-9;
    const long long t0 = clock64();
    for( int it = 0; it < iters; ++it )
    {
#pragma unroll
        for( int u = 0; u < 8; ++u )
        {
            if( ND >= 1 ) { asm volatile( "fma.rn.f64 %0, %0, %1, %2;" : "+d"( a0 ) : "d"( m ), "d"( c ) ); asm volatile( "fma.rn.f64 %0, %0, %1, %2;" : "+d"( a1 ) : "d"( m ), "d"( c ) );
                            asm volatile( "fma.rn.f64 %0, %0, %1, %2;" : "+d"( a2 ) : "d"( m ), "d"( c ) ); asm volatile( "fma.rn.f64 %0, %0, %1, %2;" : "+d"( a3 ) : "d"( m ), "d"( c ) ); }
            if( NI >= 1 ) { asm volatile( "lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"( i0 ) : "r"( i4 ), "r"( i5 ) ); asm volatile( "lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"( i1 ) : "r"( i4 ), "r"( i5 ) );
                            asm volatile( "lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"( i2 ) : "r"( i4 ), "r"( i5 ) ); asm volatile( "lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"( i3 ) : "r"( i4 ), "r"( i5 ) ); }
            if( ND >= 2 ) { asm volatile( "fma.rn.f64 %0, %0, %1, %2;" : "+d"( a4 ) : "d"( m ), "d"( c ) ); asm volatile( "fma.rn.f64 %0, %0, %1, %2;" : "+d"( a5 ) : "d"( m ), "d"( c ) );
                            asm volatile( "fma.rn.f64 %0, %0, %1, %2;" : "+d"( a6 ) : "d"( m ), "d"( c ) ); asm volatile( "fma.rn.f64 %0, %0, %1, %2;" : "+d"( a7 ) : "d"( m ), "d"( c ) ); }
            if( NI >= 2 ) { asm volatile( "lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"( i4 ) : "r"( i0 ), "r"( i1 ) ); asm volatile( "lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"( i5 ) : "r"( i0 ), "r"( i1 ) );
                            asm volatile( "lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"( i6 ) : "r"( i0 ), "r"( i1 ) ); asm volatile( "lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"( i7 ) : "r"( i0 ), "r"( i1 ) ); }
            if( NI >= 3 ) { asm volatile( "lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"( i0 ) : "r"( i6 ), "r"( i7 ) ); asm volatile( "lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"( i1 ) : "r"( i6 ), "r"( i7 ) );
                            asm volatile( "lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"( i2 ) : "r"( i6 ), "r"( i7 ) ); asm volatile( "lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"( i3 ) : "r"( i6 ), "r"( i7 ) ); }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x]  = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    outi[blockIdx.x * blockDim.x + threadIdx.x] = i0 ^ i1 ^ i2 ^ i3 ^ i4 ^ i5 ^ i6 ^ i7;
    if( threadIdx.x == 0 )
        cycles[blockIdx.x] = t1 - t0;
}

template<int ND, int NI>
void run( int warps_per_sm )
{
    const int blocks = 148, threads = warps_per_sm * 32, iters = 2000;
    double * out;
    unsigned * outi;
    long long * cyc;
    cudaMalloc( &out, sizeof( double ) * blocks * threads );
    cudaMalloc( &outi, sizeof( unsigned ) * blocks * threads );
    cudaMalloc( &cyc, sizeof( long long ) * blocks );
    k_mix<ND, NI><<<blocks, threads>>>( out, outi, cyc, 10 );
    k_mix<ND, NI><<<blocks, threads>>>( out, outi, cyc, iters );
    long long h[148];
    cudaMemcpy( h, cyc, sizeof( h ), cudaMemcpyDeviceToHost );
    double mean = 0;
    for( int i = 0; i < blocks; ++i )
        mean += double( h[i] ) / blocks;
    const double per_warp_instr = double( iters ) * 8 * 4 * ( ND + NI );
    const double ipc_smsp       = per_warp_instr * ( warps_per_sm / 4.0 ) / mean;
    printf( "fp64:int = %d:%d  warps/SM %2d  cycles %.0f  warp-instr/cycle/sub-partition %.3f  (fp64 %.3f, int %.3f)\n", ND, NI,
            warps_per_sm, mean, ipc_smsp, ipc_smsp * ND / ( ND + NI ), ipc_smsp * NI / ( ND + NI ) );
    cudaFree( out );
    cudaFree( outi );
    cudaFree( cyc );
}

int main()
{
    for( int w : { 16, 32 } )
    {
        run<1, 0>( w );
        run<0, 1>( w );
        run<1, 1>( w );
        run<1, 2>( w );
        run<1, 3>( w );
        run<2, 1>( w );
    }
    return 0;
}
