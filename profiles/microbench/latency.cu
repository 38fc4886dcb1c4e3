// Dependent-issue latency of the instructions that make up the LLG stage kernels (B200): one warp, one chain.
// Build: nvcc -arch=sm_100a -O3 -o latency latency.cu
#include <cstdio>
#include <cuda_runtime.h>



template<int which>
__global__ void k_chain( double * out, long long * cycles, int iters )
{
    double a = threadIdx.x, m = 1.0000001, c = 1e-9;
    unsigned i0 = threadIdx.x, i1 = 0x9E3779B9u;
    unsigned long long w = threadIdx.x;
    float f = threadIdx.x * 0.001f + 0.5f;
    const long long t0 = clock64();
    for( int it = 0; it < iters; ++it )
    {
#pragma unroll
        for( int u = 0; u < 16; ++u )
        {
            if( which == 0 )
                asm volatile( "fma.rn.f64 %0, %0, %1, %2;" : "+d"( a ) : "d"( m ), "d"( c ) );
            else if( which == 1 )
                asm volatile( "add.rn.f64 %0, %0, %1;" : "+d"( a ) : "d"( c ) );
            else if( which == 2 )
                asm volatile( "mul.rn.f64 %0, %0, %1;" : "+d"( a ) : "d"( m ) );
            else if( which == 3 )
                asm volatile( "lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"( i0 ) : "r"( i1 ) );
            else if( which == 4 )
                asm volatile( "mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"( w ) : "r"( unsigned( w >> 7 ) ) );
            else if( which == 5 )
                asm volatile( "lg2.approx.ftz.f32 %0, %0;" : "+f"( f ) );
            else if( which == 6 )
                asm volatile( "fma.rn.f32 %0, %0, %1, %1;" : "+f"( f ) : "f"( 1.0001f ) );
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x]   = a + i0 + double( w ) + f;
    if( threadIdx.x == 0 )
        cycles[0] = t1 - t0;
}

int main()
{
    double * out;
    long long * cyc;
    cudaMalloc( &out, 32 * sizeof( double ) );
    cudaMalloc( &cyc, sizeof( long long ) );
    const char * names[] = { "DFMA", "DADD", "DMUL", "LOP3", "IMAD.WIDE (+shift)", "MUFU.LG2", "FFMA" };
    const int iters = 4000;
#define RUN( W )                                                                                                       \
    {                                                                                                                  \
        k_chain<W><<<1, 32>>>( out, cyc, 10 );                                                                         \
        k_chain<W><<<1, 32>>>( out, cyc, iters );                                                                      \
        long long h = 0;                                                                                               \
        cudaMemcpy( &h, cyc, sizeof( h ), cudaMemcpyDeviceToHost );                                                    \
        printf( "%-20s dependent-issue latency %.2f cycles\n", names[W], double( h ) / ( 16.0 * iters ) );             \
    }
    RUN( 0 ) RUN( 1 ) RUN( 2 ) RUN( 3 ) RUN( 4 ) RUN( 5 ) RUN( 6 )
    return 0;
}
