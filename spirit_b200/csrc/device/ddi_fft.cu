// Dipole-dipole interaction as a zero-padded FFT convolution, with hand-written batched FFT passes.
// Reference: Hamiltonian_Heisenberg::Prepare_DDI / FFT_Dipole_Matrices / FFT_Spins / Gradient_DDI_FFT
// (core/src/engine/Hamiltonian_Heisenberg.cpp:1373-1601, 920-1014), FFT::FFT_Plan (core/include/engine/FFT.hpp:171-265).
//
//   g_i -= mu_i * sum_j D(r_i - r_j) mu_j s_j,   D_ab(r) = C (3 r_a r_b / r^5 - delta_ab / r^3),  C = mu_0 mu_B^2 / (4 pi 1e-30)
// summed over `ddi_n_periodic_images` along periodic directions, evaluated as a circular convolution on the padded
// lattice P_d = 2 N_d (open or zero-padded periodic directions with N_d > 1), else N_d.
//
// The reference scatters mu s into a padded real buffer, runs a library 3-D R2C (FFTW / kissFFT / cuFFT), multiplies,
// runs a library C2R and un-pads in a serial loop: ~3 kB of memory traffic per spin. Here the 3-D transform is split
// into its three 1-D passes and each pass only touches data that is not known to be zero (pruning):
//   1 k_ddi_fwd_a      reads the spin field directly (x mu_s), length-Pa transforms of the Nb*Nc non-zero rows,
//                      writes the half spectrum  A[q][c<Nc][b<Nb][ka<=Pa/2]
//   2 k_fft_pass (b)   A -> B[q][c<Nc][kb<Pb][ka]                 (inputs b >= Nb are zero and never read)
//   3 k_ddi_c_mult     per (kb, ka): forward c-transform of the 3 NB components (inputs c >= Nc zero), multiply with
//                      the precomputed tensor spectrum D^(kc, kb, ka) (3x3 symmetric per sublattice pair), inverse
//                      c-transform, keep c < Nc:  B -> B (in place)
//   4 k_fft_pass (b)   inverse, keep b < Nb:  B -> A
//   5 k_ddi_inv_a      Hermitian extension, inverse a-transform, keep a < Na, g_ddi = -mu_s * result / P written as
//                      a field (AoSoA-32) that the stencil kernels add to the gradient
// Every pass is a batch of shared-memory FFTs, `ncol` adjacent transforms per CTA so that strided passes still move
// contiguous segments. Two families of kernels serve them:
//   * power-of-two lengths 64 ... 4096 (every zero-padded power-of-two lattice): k_ddi_fwd_a16 / k_ddi_inv_a16 (real rows as
//     complex sequences of half the length), k_fft_pass16r, k_ddi_c_mult16f -- instantiated per length (block_fft_ct: stage
//     loop unrolled at compile time), radix-8 butterflies in registers, IN PLACE in one padded shared buffer, twiddles of a
//     butterfly as powers of one table entry, tensor spectrum real (one sublattice) and stored in the c-pass's tile order;
//     the b- and c-passes start their first stage on the loaded registers and store from their last butterflies, the c-pass
//     multiplies with the tensor between its forward last and inverse first stage without leaving the registers
//     (k_fft_pass16 / k_ddi_c_mult16 are the same passes with every stage through shared memory: tuning builds with other
//     radices, SPIRIT_B200_FFT_PASS_REG=0 / SPIRIT_B200_DDI_C_FUSED=0); thin films (Pc <= 32) do the c-transforms in
//     registers (k_ddi_c_mult_small);
//   * any other length and any basis: k_ddi_fwd_a, k_fft_pass, k_ddi_c_mult, k_ddi_inv_a -- mixed-radix (4 / 2 / generic
//     prime) Stockham passes between two shared buffers.
// Slabs over several GPUs: the kb axis of B is cut into per-rank blocks, the all-to-alls travel one component at a time on
// their own stream under the passes of the other components (ddi_gradient_pipelined).
// cuFFT is not used by the product; the tests compare against the reference's FFT and direct-sum paths, every per-length
// instantiation included (tests/test_ddi_gpu.py).
#include "device_buffers.cuh"

#include "../core/constants.hpp"
#include "../core/hamiltonian.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace sb
{
namespace dev
{

namespace
{

#ifndef SB_FFT_THREADS
#define SB_FFT_THREADS 512 // tuning builds: 256 lifts the register cap of the fast kernels to 255 (radix-16 stages without spills)
#endif
constexpr int FFT_THREADS  = SB_FFT_THREADS;  // upper bound (128 registers: a radix-16 butterfly lives in 64); launches use fft_threads( work items )
constexpr int MAX_RADICES  = 16;
constexpr int MAX_SMEM_FFT = 200 * 1024;

struct FFTPlan1D
{
    int n       = 1;
    int n_radix = 0;
    int pow2    = 0; // n is a power of two (all strides are powers of two: shifts instead of divisions)
    int fast16  = 0; // power of two >= 64: radices 16 ... 16 [2 | 4 | 8], served by block_fft16
    int radix[MAX_RADICES];
    const double2 * twiddle = nullptr; // exp(-2 pi i k / n), k < n
};

__device__ __forceinline__ double2 cmul( const double2 & a, const double2 & b )
{
    return make_double2( a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x );
}
__device__ __forceinline__ double2 cadd( const double2 & a, const double2 & b )
{
    return make_double2( a.x + b.x, a.y + b.y );
}
__device__ __forceinline__ double2 csub( const double2 & a, const double2 & b )
{
    return make_double2( a.x - b.x, a.y - b.y );
}
// twiddle exp(-+2 pi i k/n): the table holds the forward sign
template<bool INVERSE>
__device__ __forceinline__ double2 tw( const FFTPlan1D & plan, int k )
{
    const double2 w = __ldg( plan.twiddle + k );
    return INVERSE ? make_double2( w.x, -w.y ) : w;
}

// exp(-2 pi i k / 32), k < 32: twiddles of the in-register transforms (lengths up to 32)
static __constant__ double2 TW32[32] = {
    { 1, 0 },
    { 0.98078528040323043, -0.19509032201612825 },
    { 0.92387953251128674, -0.38268343236508978 },
    { 0.83146961230254524, -0.55557023301960218 },
    { 0.70710678118654757, -0.70710678118654746 },
    { 0.55557023301960229, -0.83146961230254524 },
    { 0.38268343236508984, -0.92387953251128674 },
    { 0.19509032201612833, -0.98078528040323043 },
    { 0, -1 },
    { -0.19509032201612819, -0.98078528040323043 },
    { -0.38268343236508973, -0.92387953251128674 },
    { -0.55557023301960196, -0.83146961230254546 },
    { -0.70710678118654746, -0.70710678118654757 },
    { -0.83146961230254535, -0.55557023301960218 },
    { -0.92387953251128674, -0.38268343236508989 },
    { -0.98078528040323043, -0.19509032201612861 },
    { -1, 0 },
    { -0.98078528040323043, 0.19509032201612836 },
    { -0.92387953251128685, 0.38268343236508967 },
    { -0.83146961230254546, 0.55557023301960196 },
    { -0.70710678118654768, 0.70710678118654746 },
    { -0.55557023301960218, 0.83146961230254524 },
    { -0.38268343236509034, 0.92387953251128652 },
    { -0.19509032201612866, 0.98078528040323032 },
    { 0, 1 },
    { 0.1950903220161283, 0.98078528040323043 },
    { 0.38268343236509, 0.92387953251128663 },
    { 0.55557023301960184, 0.83146961230254546 },
    { 0.70710678118654735, 0.70710678118654768 },
    { 0.83146961230254524, 0.55557023301960218 },
    { 0.92387953251128652, 0.38268343236509039 },
    { 0.98078528040323032, 0.19509032201612872 } };


// ---------------------------------------------------------------------------------------------
// Power-of-two lengths >= 64: in-register radix-16 stages. n = 16 * 16 * ... * rem (rem = 2, 4, 8 or nothing): three
// stages for 4096 instead of six radix-4 ones. A thread owns 16 elements per stage (one radix-16 butterfly, or 16/R
// butterflies of the last radix R), computed as 4 x 4 (8: 4 x 2) small transforms in registers. The exchange between
// stages goes through ONE shared buffer in place: all threads read their inputs, barrier, all write their outputs.
// Between stages the buffer is padded by one element per 16 (index a -> a + a/16), which makes the stride-16 writes of
// the first stage conflict-free; the first stage reads and the last stage writes the caller's plain layout.
// Same Stockham indexing as the generic path: inputs a_i = x[q + s (p + m i)], outputs x[q + s (R p + i)].
// ---------------------------------------------------------------------------------------------
template<bool INVERSE>
__device__ __forceinline__ double2 tw32( int k )
{
    const double2 w = TW32[k];
    return INVERSE ? make_double2( w.x, -w.y ) : w;
}
template<bool INVERSE>
__device__ __forceinline__ void dft4_inplace( double2 & a0, double2 & a1, double2 & a2, double2 & a3 )
{
    const double2 b0 = cadd( a0, a2 ), b1 = csub( a0, a2 ), b2 = cadd( a1, a3 ), b3 = csub( a1, a3 );
    const double2 jb3 = INVERSE ? make_double2( -b3.y, b3.x ) : make_double2( b3.y, -b3.x ); // -+ i b3
    a0 = cadd( b0, b2 );
    a1 = cadd( b1, jb3 );
    a2 = csub( b0, b2 );
    a3 = csub( b1, jb3 );
}
// In-register transform of length R, natural order in and out.
template<bool INVERSE, int R>
__device__ __forceinline__ void dft_small( double2 ( &a )[R] )
{
    if( R == 2 )
    {
        const double2 t = a[0];
        a[0]            = cadd( t, a[1] );
        a[1]            = csub( t, a[1] );
    }
    else if( R == 4 )
        dft4_inplace<INVERSE>( a[0], a[1 % R], a[2 % R], a[3 % R] );
    else if( R == 8 )
    {
        // n = 2 n1 + n2, k = k1 + 4 k2: y[n2][k1] = DFT4_{n1} x[2 n1 + n2] (left in slot 2 k1 + n2), times W8^{n2 k1},
        // X[k1 + 4 k2] = DFT2_{n2} y[n2][k1] (left in slot 2 k1 + k2)
#pragma unroll
        for( int n2 = 0; n2 < 2; ++n2 )
            dft4_inplace<INVERSE>( a[n2 % R], a[( 2 + n2 ) % R], a[( 4 + n2 ) % R], a[( 6 + n2 ) % R] );
#pragma unroll
        for( int k1 = 1; k1 < 4; ++k1 )
            a[( 2 * k1 + 1 ) % R] = cmul( a[( 2 * k1 + 1 ) % R], tw32<INVERSE>( 4 * k1 ) );
        double2 o[R];
#pragma unroll
        for( int k1 = 0; k1 < 4; ++k1 )
        {
            o[k1 % R]         = cadd( a[( 2 * k1 ) % R], a[( 2 * k1 + 1 ) % R] );
            o[( k1 + 4 ) % R] = csub( a[( 2 * k1 ) % R], a[( 2 * k1 + 1 ) % R] );
        }
#pragma unroll
        for( int i = 0; i < R; ++i )
            a[i] = o[i];
    }
    else
    {
        // n = 4 n1 + n2, k = k1 + 4 k2: y[n2][k1] = DFT4_{n1} x[4 n1 + n2] (slot 4 k1 + n2), times W16^{n2 k1},
        // X[k1 + 4 k2] = DFT4_{n2} y[n2][k1] (slot 4 k1 + k2)
#pragma unroll
        for( int n2 = 0; n2 < 4; ++n2 )
            dft4_inplace<INVERSE>( a[n2 % R], a[( 4 + n2 ) % R], a[( 8 + n2 ) % R], a[( 12 + n2 ) % R] );
#pragma unroll
        for( int k1 = 1; k1 < 4; ++k1 )
#pragma unroll
            for( int n2 = 1; n2 < 4; ++n2 )
                a[( 4 * k1 + n2 ) % R] = cmul( a[( 4 * k1 + n2 ) % R], tw32<INVERSE>( 2 * n2 * k1 ) );
#pragma unroll
        for( int k1 = 0; k1 < 4; ++k1 )
            dft4_inplace<INVERSE>( a[( 4 * k1 ) % R], a[( 4 * k1 + 1 ) % R], a[( 4 * k1 + 2 ) % R], a[( 4 * k1 + 3 ) % R] );
        double2 o[R];
#pragma unroll
        for( int k1 = 0; k1 < 4; ++k1 )
#pragma unroll
            for( int k2 = 0; k2 < 4; ++k2 )
                o[( k1 + 4 * k2 ) % R] = a[( 4 * k1 + k2 ) % R];
#pragma unroll
        for( int i = 0; i < R; ++i )
            a[i] = o[i];
    }
}

#ifndef SB_FFT_LG_E
#define SB_FFT_LG_E 3 // tuning builds (spirit_b200/build.py SPIRIT_B200_DEFINES): 2 = radix-4 stages
#endif
#ifndef SB_FFT16_MINB
#define SB_FFT16_MINB 1 // CTAs per SM the fast kernels are compiled for (register cap 65536 / (512 * MINB))
#endif
constexpr int FFT_LG_E = SB_FFT_LG_E, FFT_E = 1 << FFT_LG_E; // elements a thread owns in the in-register stages (16: 3 stages for 4096 but
                                       // the twiddles push ptxas over 128 registers; 8: 4 stages, no spills)
template<bool INVERSE, int R>
__device__ __forceinline__ void fft16_stage(
    const FFTPlan1D & plan, double2 * x, const int ncol, const int lg_ncol, const int s, const int lg_s, const bool first, const bool last )
{
    constexpr int PER  = FFT_E / R; // butterflies per thread
    const int n        = plan.n;
    const int per_col  = n >> FFT_LG_E;        // threads per transform
    const int step     = n / R;                // s m: distance of the inputs of a butterfly
    const bool working = int( threadIdx.x ) < ( per_col << lg_ncol );
    const int col = threadIdx.x & ( ncol - 1 ), tt = threadIdx.x >> lg_ncol;
    double2 v[PER][R];
    if( working )
    {
#pragma unroll
        for( int jj = 0; jj < PER; ++jj )
        {
            const int item = tt + per_col * jj; // = q + s p
#pragma unroll
            for( int i = 0; i < R; ++i )
            {
                const int a   = item + step * i;
                const int idx = first ? a : a + ( a >> 4 );
                v[jj][i]      = x[( idx << lg_ncol ) + col];
            }
            dft_small<INVERSE, R>( v[jj] );
        }
    }
    __syncthreads();
    if( working )
    {
#pragma unroll
        for( int jj = 0; jj < PER; ++jj )
        {
            const int item = tt + per_col * jj;
            const int q = item & ( s - 1 ), p = item >> lg_s;
#pragma unroll
            for( int i = 0; i < R; ++i )
            {
                double2 val = v[jj][i];
                if( !last && i > 0 )
                    val = cmul( val, tw<INVERSE>( plan, i * p * s ) ); // w_{n/s}^{p i}; p = 0 in the last stage
                const int o   = q + s * ( R * p + i );
                const int idx = last ? o : o + ( o >> 4 );
                x[( idx << lg_ncol ) + col] = val;
            }
        }
    }
    __syncthreads();
}

// In place in x (which needs room for n * ncol * 17 / 16 elements). All threads of the CTA must call;
// blockDim.x >= n / FFT_E * ncol, ncol a power of two.
template<bool INVERSE>
__device__ double2 * block_fft16( const FFTPlan1D & plan, double2 * x, int ncol )
{
    const int lg_ncol = 31 - __clz( ncol );
    int s = 1, lg_s = 0;
    for( int stage = 0; stage < plan.n_radix; ++stage )
    {
        const int r      = plan.radix[stage];
        const bool first = stage == 0, last = stage == plan.n_radix - 1;
        if( r == 16 && FFT_E >= 16 )
            fft16_stage<INVERSE, FFT_E >= 16 ? 16 : FFT_E>( plan, x, ncol, lg_ncol, s, lg_s, first, last );
        else if( r == 8 && FFT_E >= 8 )
            fft16_stage<INVERSE, FFT_E >= 8 ? 8 : FFT_E>( plan, x, ncol, lg_ncol, s, lg_s, first, last );
        else if( r == 4 )
            fft16_stage<INVERSE, 4>( plan, x, ncol, lg_ncol, s, lg_s, first, last );
        else
            fft16_stage<INVERSE, 2>( plan, x, ncol, lg_ncol, s, lg_s, first, last );
        s *= r;
        lg_s += 31 - __clz( r );
    }
    return x;
}

// ---------------------------------------------------------------------------------------------
// The same transform with the length as a template parameter (n = 1 << LOGN, 64 <= n <= 4096): the stage loop is unrolled
// at compile time, strides, radices and the padding arithmetic are immediates. With a run-time length ptxas hoists the
// index arithmetic of every stage variant out of the stage loop (216 registers uncapped, spills under the 128 cap).
// Used by the fast pass kernels below, which are instantiated per length and selected on the host.
// ---------------------------------------------------------------------------------------------
#ifndef SB_FFT_TW_CHAIN_LOGN
#define SB_FFT_TW_CHAIN_LOGN 6 // from this length on the twiddles of a butterfly are powers of ONE table entry (measured: faster at every length, profiles/r1u)
#endif
template<bool INVERSE, int LOGN, int R, int LG_S, bool FIRST, bool LAST>
__device__ __forceinline__ void fft_stage_ct( const double2 * __restrict__ twiddle, double2 * x, const int lg_ncol, const int col, const int tt )
{
    constexpr int N = 1 << LOGN, PER = FFT_E / R, PER_COL = N >> FFT_LG_E, STEP = N / R, S = 1 << LG_S;
    // long transforms: R - 1 look-ups per butterfly are gathers of 16 bytes from as many cache lines per thread, which
    // cost more than the butterfly; there the powers w^2 .. w^(R-1) of the one entry w = w_{n/s}^p are multiplied up
    constexpr bool CHAIN = LOGN >= SB_FFT_TW_CHAIN_LOGN;
    double2 v[PER][R];
    double2 w1[PER];
    if( !LAST && CHAIN )
    {
#pragma unroll
        for( int jj = 0; jj < PER; ++jj )
            w1[jj] = __ldg( twiddle + ( ( ( tt + PER_COL * jj ) >> LG_S ) << LG_S ) );
    }
#pragma unroll
    for( int jj = 0; jj < PER; ++jj )
    {
        const int item = tt + PER_COL * jj; // = q + s p
#pragma unroll
        for( int i = 0; i < R; ++i )
        {
            const int a   = item + STEP * i;
            const int idx = FIRST ? a : a + ( a >> 4 );
            v[jj][i]      = x[( idx << lg_ncol ) + col];
        }
        dft_small<INVERSE, R>( v[jj] );
    }
    __syncthreads();
#pragma unroll
    for( int jj = 0; jj < PER; ++jj )
    {
        const int item = tt + PER_COL * jj;
        const int q = item & ( S - 1 ), p = item >> LG_S;
        double2 wc = make_double2( 1.0, 0.0 );
#pragma unroll
        for( int i = 0; i < R; ++i )
        {
            double2 val = v[jj][i];
            if( !LAST && i > 0 )
            {
                if( CHAIN )
                {
                    const double2 w = INVERSE ? make_double2( w1[jj].x, -w1[jj].y ) : w1[jj];
                    wc              = i == 1 ? w : cmul( wc, w );
                }
                else
                {
                    const double2 w = __ldg( twiddle + ( ( i * p ) << LG_S ) ); // w_{n/s}^{p i}
                    wc              = INVERSE ? make_double2( w.x, -w.y ) : w;
                }
                val = cmul( val, wc );
            }
            const int o   = q + S * ( R * p + i );
            const int idx = LAST ? o : o + ( o >> 4 );
            x[( idx << lg_ncol ) + col] = val;
        }
    }
    __syncthreads();
}
// stages: radix FFT_E while it divides what is left, then the rest (2 or 4) -- the plan of make_plan_1d for fast16 lengths
template<bool INVERSE, int LOGN, int LG_S = 0>
__device__ __forceinline__ void block_fft_ct( const double2 * __restrict__ twiddle, double2 * x, const int lg_ncol, const int col, const int tt )
{
    constexpr int LEFT = LOGN - LG_S;
    if constexpr( LEFT >= FFT_LG_E )
    {
        fft_stage_ct<INVERSE, LOGN, FFT_E, LG_S, LG_S == 0, LEFT == FFT_LG_E>( twiddle, x, lg_ncol, col, tt );
        if constexpr( LEFT > FFT_LG_E )
            block_fft_ct<INVERSE, LOGN, LG_S + FFT_LG_E>( twiddle, x, lg_ncol, col, tt );
    }
    else
        fft_stage_ct<INVERSE, LOGN, ( 1 << LEFT ), LG_S, LG_S == 0, true>( twiddle, x, lg_ncol, col, tt );
}

// In-shared-memory Stockham autosort FFT of `ncol` independent sequences of length plan.n stored as x[j * ncol + col].
// Returns the buffer (x or y) that holds the result. All threads of the CTA must call.
template<bool INVERSE>
__device__ double2 * block_fft( const FFTPlan1D & plan, double2 * x, double2 * y, int ncol )
{
    if( plan.fast16 && ( ncol & ( ncol - 1 ) ) == 0 && int( blockDim.x ) >= ( plan.n >> FFT_LG_E ) * ncol )
        return block_fft16<INVERSE>( plan, x, ncol );
    const int n = plan.n;
    int s       = 1; // stride = product of the radices already processed
    for( int stage = 0; stage < plan.n_radix; ++stage )
    {
        const int r = plan.radix[stage];
        const int m = n / ( s * r ); // remaining sub-transform length / r
        // one work item = (butterfly (p, q), column)
        const int items = m * s * ncol;
        // ncol and (for power-of-two lengths) s are powers of two: shifts and masks instead of integer divisions
        const int lg_ncol = 31 - __clz( ncol ), lg_s = 31 - __clz( s );
        const bool fast   = plan.pow2 && ( ncol & ( ncol - 1 ) ) == 0;
        for( int item = threadIdx.x; item < items; item += blockDim.x )
        {
            int col, q, p;
            if( fast )
            {
                col         = item & ( ncol - 1 );
                const int t = item >> lg_ncol;
                q           = t & ( s - 1 );
                p           = t >> lg_s;
            }
            else
            {
                col         = item % ncol;
                const int t = item / ncol;
                q           = t % s;
                p           = t / s;
            }
            // inputs a_i = x[q + s (p + m i)], outputs y[q + s (r p + i)] = (sum_k a_k w_r^{ik}) w_{n/s}^{p i}
            const int tw_step = p * s; // w_{n/s}^{p} = W_n^{p s}
            if( r == 2 )
            {
                const double2 a0 = x[( q + s * p ) * ncol + col];
                const double2 a1 = x[( q + s * ( p + m ) ) * ncol + col];
                y[( q + s * ( 2 * p ) ) * ncol + col]     = cadd( a0, a1 );
                y[( q + s * ( 2 * p + 1 ) ) * ncol + col] = cmul( csub( a0, a1 ), tw<INVERSE>( plan, tw_step ) );
            }
            else if( r == 4 )
            {
                const double2 a0 = x[( q + s * p ) * ncol + col];
                const double2 a1 = x[( q + s * ( p + m ) ) * ncol + col];
                const double2 a2 = x[( q + s * ( p + 2 * m ) ) * ncol + col];
                const double2 a3 = x[( q + s * ( p + 3 * m ) ) * ncol + col];
                const double2 b0 = cadd( a0, a2 ), b1 = csub( a0, a2 ), b2 = cadd( a1, a3 ), b3 = csub( a1, a3 );
                // -i b3 (forward) or +i b3 (inverse)
                const double2 jb3 = INVERSE ? make_double2( -b3.y, b3.x ) : make_double2( b3.y, -b3.x );
                y[( q + s * ( 4 * p ) ) * ncol + col]     = cadd( b0, b2 );
                y[( q + s * ( 4 * p + 1 ) ) * ncol + col] = cmul( cadd( b1, jb3 ), tw<INVERSE>( plan, tw_step ) );
                y[( q + s * ( 4 * p + 2 ) ) * ncol + col] = cmul( csub( b0, b2 ), tw<INVERSE>( plan, 2 * tw_step ) );
                y[( q + s * ( 4 * p + 3 ) ) * ncol + col] = cmul( csub( b1, jb3 ), tw<INVERSE>( plan, 3 * tw_step ) );
            }
            else
            {
                // generic prime radix: O(r^2) butterfly, w_r^{ik} = W_n^{(i k mod r) n / r}
                const int nr = n / r;
                for( int i = 0; i < r; ++i )
                {
                    double2 acc = make_double2( 0.0, 0.0 );
                    for( int k = 0; k < r; ++k )
                    {
                        const double2 a = x[( q + s * ( p + m * k ) ) * ncol + col];
                        acc             = cadd( acc, cmul( a, tw<INVERSE>( plan, ( ( i * k ) % r ) * nr ) ) );
                    }
                    y[( q + s * ( r * p + i ) ) * ncol + col] = cmul( acc, tw<INVERSE>( plan, int( ( std::int64_t( i ) * tw_step ) % n ) ) );
                }
            }
        }
        __syncthreads();
        double2 * t = x;
        x           = y;
        y           = t;
        s *= r;
    }
    return x;
}

// ---------------------------------------------------------------------------------------------
// Generic strided pass: transforms along j of  in[o * in_os + j * in_js + u]  for u < n_u (contiguous), o < n_o.
// Inputs j >= n_in are zero (not read), outputs j >= n_out are dropped. A CTA handles `ncol` consecutive u of one o.
// ---------------------------------------------------------------------------------------------
constexpr int DDI_MAX_PEERS = 8;
struct PassArgs
{
    const double2 * in;
    double2 * out;
    std::size_t in_os, in_js, out_os, out_js;
    int n_u, n_o, n_in, n_out, ncol;
    double scale;
    // optional blocked layout of the transform index j on the input / output side: element j lives at
    // (j / split) * split_stride + (j % split) * js. Used by the distributed convolution, where the kb axis is cut into
    // per-rank blocks so that the all-to-all moves one contiguous block per peer.
    int in_split, out_split;
    std::size_t in_split_stride, out_split_stride;
    // distributed convolution over peer-mapped memory (k_fft_pass16 only): block r of the split axis lives in the memory of
    // rank r. in_peer[r] / out_peer[r] point at this rank's block inside rank r's operand (in place of in / out +
    // r * split_stride); the transposes of the convolution are these remote loads / stores.
    int use_in_peer, use_out_peer;
    const double2 * in_peer[DDI_MAX_PEERS];
    double2 * out_peer[DDI_MAX_PEERS];
    // pencil decomposition (ka cut over the ranks): the OUTER index o = q * opeer_planes + c (c a global plane) selects the
    // destination: rank c / opeer_ncl, whose buffer out_peer[rank] holds element (q, c % opeer_ncl) at ((q * opeer_ncl) +
    // c % opeer_ncl) * out_os. The inverse b-pass stores its rows straight into the a-pass operand of the rank that owns the plane.
    int opeer_planes, opeer_ncl;
};
__device__ __forceinline__ std::size_t pass_offset( int j, std::size_t js, int split, std::size_t split_stride )
{
    return split ? std::size_t( j / split ) * split_stride + std::size_t( j % split ) * js : std::size_t( j ) * js;
}

template<bool INVERSE>
static __global__ void __launch_bounds__( FFT_THREADS ) k_fft_pass( const __grid_constant__ FFTPlan1D plan, const __grid_constant__ PassArgs a )
{
    extern __shared__ double2 smem[];
    const int n    = plan.n;
    double2 * x    = smem;
    double2 * y    = smem + std::size_t( n ) * a.ncol;
    const int tile = blockIdx.x, o = blockIdx.y;
    const int u0   = tile * a.ncol;
    const int nc   = min( a.ncol, a.n_u - u0 );
    for( int item = threadIdx.x; item < n * a.ncol; item += blockDim.x )
    {
        const int col = item % a.ncol, j = item / a.ncol;
        double2 v     = make_double2( 0.0, 0.0 );
        if( j < a.n_in && col < nc )
            v = a.in[std::size_t( o ) * a.in_os + pass_offset( j, a.in_js, a.in_split, a.in_split_stride ) + u0 + col];
        x[item] = v;
    }
    __syncthreads();
    const double2 * r = block_fft<INVERSE>( plan, x, y, a.ncol );
    for( int item = threadIdx.x; item < a.n_out * a.ncol; item += blockDim.x )
    {
        const int col = item % a.ncol, j = item / a.ncol;
        if( col < nc )
        {
            const double2 v = r[item];
            a.out[std::size_t( o ) * a.out_os + pass_offset( j, a.out_js, a.out_split, a.out_split_stride ) + u0 + col]
                = make_double2( a.scale * v.x, a.scale * v.y );
        }
    }
}

// The same pass for FFTPlan1D::fast16 lengths: ONE shared buffer (block_fft16 works in place), exactly one radix-8 work
// item per thread ((n / 8) << lg_ncol threads), no integer division (ncol and the per-rank split of the distributed
// layout are powers of two, lg = 31 stands for "no split"), a thread keeps its column for the whole pass. Half the
// shared memory and half the threads of k_fft_pass: two or more CTAs per SM, so that the loads of one overlap the
// butterflies of another.
__device__ __forceinline__ std::size_t pass_offset16( int j, std::size_t js, int lg_split, std::size_t split_stride )
{
    return std::size_t( j >> lg_split ) * split_stride + std::size_t( unsigned( j ) & ( ( 1u << lg_split ) - 1u ) ) * js;
}
// LG_SEQ: the CTA takes (1 << LG_SEQ) groups of ncol columns, adjacent in memory, loads and stores them together (so that
// a row of the tile is (ncol << LG_SEQ) x 16 contiguous bytes: for the long transforms, where ncol = 1, a full 32-byte
// sector instead of half of one) and transforms the groups one after the other, each in its own dense buffer.
template<bool INVERSE, int LOGN, int LG_SEQ>
static __global__ void __launch_bounds__( FFT_THREADS, SB_FFT16_MINB ) k_fft_pass16(
    const __grid_constant__ FFTPlan1D plan, const __grid_constant__ PassArgs a, const int lg_ncol, const int lg_in_split,
    const int lg_out_split )
{
    extern __shared__ double2 smem[];
    constexpr int NSEQ = 1 << LG_SEQ;
    const int lg_tile  = lg_ncol + LG_SEQ;                                          // columns of the CTA = 1 << lg_tile
    const int bufp     = ( ( 1 << LOGN ) << lg_ncol ) + ( ( ( 1 << LOGN ) << lg_ncol ) >> 4 ) + 1; // elements per group buffer
    const int o = blockIdx.y, u0 = blockIdx.x << lg_tile;
    const int ctile = threadIdx.x & ( ( 1 << lg_tile ) - 1 ), j0 = threadIdx.x >> lg_tile, jstep = blockDim.x >> lg_tile;
    const int col   = ctile & ( ( 1 << lg_ncol ) - 1 );
    double2 * mine  = smem + ( ctile >> lg_ncol ) * bufp; // buffer of this thread's column group (load / store phases)
    const bool valid   = u0 + ctile < a.n_u;
    const double2 * in = a.in + std::size_t( o ) * a.in_os + u0 + ctile;
    // exactly FFT_E * NSEQ elements per thread (blockDim = (n / FFT_E) << lg_ncol): FFT_E loads in flight before their stores
#pragma unroll
    for( int h = 0; h < NSEQ; ++h )
    {
        double2 v[FFT_E];
#pragma unroll
        for( int k = 0; k < FFT_E; ++k )
        {
            const int j = j0 + ( h * FFT_E + k ) * jstep;
            v[k]        = make_double2( 0.0, 0.0 );
            if( valid && j < a.n_in )
            {
                if( a.use_in_peer )
                    v[k] = a.in_peer[j >> lg_in_split]
                                    [std::size_t( o ) * a.in_os + u0 + ctile + std::size_t( unsigned( j ) & ( ( 1u << lg_in_split ) - 1u ) ) * a.in_js];
                else
                    v[k] = in[pass_offset16( j, a.in_js, lg_in_split, a.in_split_stride )];
            }
        }
#pragma unroll
        for( int k = 0; k < FFT_E; ++k )
            mine[( ( j0 + ( h * FFT_E + k ) * jstep ) << lg_ncol ) + col] = v[k];
    }
    __syncthreads();
#pragma unroll 1
    for( int g = 0; g < NSEQ; ++g )
        block_fft_ct<INVERSE, LOGN>(
            plan.twiddle, smem + g * bufp, lg_ncol, int( threadIdx.x ) & ( ( 1 << lg_ncol ) - 1 ), int( threadIdx.x ) >> lg_ncol );
    if( valid )
    {
        double2 * out = a.out + std::size_t( o ) * a.out_os + u0 + ctile;
        if( a.opeer_planes )
        {
            const int q = o / a.opeer_planes, c = o - q * a.opeer_planes, r = c / a.opeer_ncl;
            out         = a.out_peer[r] + std::size_t( q * a.opeer_ncl + ( c - r * a.opeer_ncl ) ) * a.out_os + u0 + ctile;
        }
#pragma unroll
        for( int k = 0; k < FFT_E * NSEQ; ++k )
        {
            const int j = j0 + k * jstep;
            if( j < a.n_out )
            {
                const double2 w = mine[( j << lg_ncol ) + col];
                if( a.use_out_peer )
                    a.out_peer[j >> lg_out_split]
                              [std::size_t( o ) * a.out_os + u0 + ctile + std::size_t( unsigned( j ) & ( ( 1u << lg_out_split ) - 1u ) ) * a.out_js]
                        = make_double2( a.scale * w.x, a.scale * w.y );
                else
                    out[pass_offset16( j, a.out_js, lg_out_split, a.out_split_stride )] = make_double2( a.scale * w.x, a.scale * w.y );
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The same pass as a PERSISTENT, software-pipelined kernel: every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... of
// the (o, column group) space with TWO tile buffers. The tile after the current one is fetched with 16-byte asynchronous
// copies (cp.async / LDGSTS: global -> shared without passing through registers, zero-filled where the input is known to be
// zero) issued before the butterflies of the current tile, so the memory system works under the radix stages instead of
// waiting for them. One barrier per tile on top of those of the stages: a thread waits for its own copies, the barrier makes
// everybody's visible AND says that the buffer of the previous tile has been stored and may be refilled.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16( void * smem_dst, const void * src, bool nonzero )
{
    const unsigned dst = unsigned( __cvta_generic_to_shared( smem_dst ) );
    const int bytes    = nonzero ? 16 : 0; // src-size 0: the 16 bytes are zero-filled, src is not read
    asm volatile( "cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"( dst ), "l"( src ), "r"( bytes ) : "memory" );
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile( "cp.async.commit_group;" ::: "memory" );
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile( "cp.async.wait_group 0;" ::: "memory" );
}

template<bool INVERSE, int LOGN>
static __global__ void __launch_bounds__( FFT_THREADS, SB_FFT16_MINB ) k_fft_pass16p(
    const __grid_constant__ FFTPlan1D plan, const __grid_constant__ PassArgs a, const int lg_ncol, const int lg_in_split,
    const int lg_out_split, const int n_tiles_u, const int n_tiles )
{
    extern __shared__ double2 smem[];
    const int bufp = ( ( 1 << LOGN ) << lg_ncol ) + ( ( ( 1 << LOGN ) << lg_ncol ) >> 4 ) + 1; // elements per tile buffer
    const int col = threadIdx.x & ( ( 1 << lg_ncol ) - 1 ), j0 = threadIdx.x >> lg_ncol, jstep = blockDim.x >> lg_ncol;
    // exactly FFT_E elements per thread and tile (blockDim = (n / FFT_E) << lg_ncol)
    auto fetch = [&]( int tile, double2 * buf )
    {
        const int o = tile / n_tiles_u, u = ( ( tile - o * n_tiles_u ) << lg_ncol ) + col;
        const bool valid        = u < a.n_u;
        const std::size_t base  = std::size_t( o ) * a.in_os + u;
#pragma unroll
        for( int k = 0; k < FFT_E; ++k )
        {
            const int j          = j0 + k * jstep;
            const bool nonzero   = valid && j < a.n_in;
            const double2 * src = a.in;
            if( nonzero )
                src = a.use_in_peer ? a.in_peer[j >> lg_in_split] + base + std::size_t( unsigned( j ) & ( ( 1u << lg_in_split ) - 1u ) ) * a.in_js
                                    : a.in + base + pass_offset16( j, a.in_js, lg_in_split, a.in_split_stride );
            cp_async16( buf + ( j << lg_ncol ) + col, src, nonzero );
        }
        cp_async_commit();
    };
    int tile  = blockIdx.x;
    int stage = 0;
    if( tile < n_tiles )
        fetch( tile, smem );
    for( ; tile < n_tiles; tile += gridDim.x, stage ^= 1 )
    {
        cp_async_wait_all();
        __syncthreads();
        if( tile + int( gridDim.x ) < n_tiles )
            fetch( tile + gridDim.x, smem + ( stage ^ 1 ) * bufp );
        double2 * x = smem + stage * bufp;
        block_fft_ct<INVERSE, LOGN>( plan.twiddle, x, lg_ncol, col, j0 );
        const int o = tile / n_tiles_u, u = ( ( tile - o * n_tiles_u ) << lg_ncol ) + col;
        if( u < a.n_u )
        {
            const std::size_t base = std::size_t( o ) * a.out_os + u;
            double2 * obase        = nullptr;
            if( a.opeer_planes )
            {
                const int q = o / a.opeer_planes, c = o - q * a.opeer_planes, r = c / a.opeer_ncl;
                obase       = a.out_peer[r] + std::size_t( q * a.opeer_ncl + ( c - r * a.opeer_ncl ) ) * a.out_os + u;
            }
#pragma unroll
            for( int k = 0; k < FFT_E; ++k )
            {
                const int j = j0 + k * jstep;
                if( j < a.n_out )
                {
                    const double2 w = x[( j << lg_ncol ) + col];
                    double2 * dst   = a.use_out_peer
                                          ? a.out_peer[j >> lg_out_split] + base + std::size_t( unsigned( j ) & ( ( 1u << lg_out_split ) - 1u ) ) * a.out_js
                                          : ( obase ? obase : a.out + base ) + pass_offset16( j, a.out_js, lg_out_split, a.out_split_stride );
                    *dst = make_double2( a.scale * w.x, a.scale * w.y );
                }
            }
        }
    }
}

// Pass a for a dense REAL input (setup of the tensor spectrum): real row of length n -> half spectrum
static __global__ void __launch_bounds__( FFT_THREADS ) k_fft_real_rows(
    const __grid_constant__ FFTPlan1D plan, const double * __restrict__ in, double2 * __restrict__ out, int Ha )
{
    extern __shared__ double2 smem[];
    const int n       = plan.n;
    double2 * x       = smem;
    double2 * y       = smem + n;
    const size_t row  = blockIdx.x;
    for( int j = threadIdx.x; j < n; j += blockDim.x )
        x[j] = make_double2( in[row * n + j], 0.0 );
    __syncthreads();
    const double2 * r = block_fft<false>( plan, x, y, 1 );
    for( int j = threadIdx.x; j < Ha; j += blockDim.x )
        out[row * Ha + j] = r[j];
}

// ---------------------------------------------------------------------------------------------
// DDI plan (device side)
// ---------------------------------------------------------------------------------------------
struct DDIDims
{
    int Na, Nb, Nc, NB;
    int Pa, Pb, Pc, Ha; // padded sizes, Ha = Pa/2 + 1
    int n_inter;
    int lookup[MAX_BASIS * MAX_BASIS]; // inter-sublattice index of (b1, b2)  (Hamiltonian_Heisenberg.cpp:1418-1428)
    int plane_stride, halo;
    double mu_s[MAX_BASIS];
    // layout of the c-pass operand: element (q, c, kb, ka) at
    //   (c / c_block) * block_stride + q * q_stride + (c % c_block) * Pb * Ha + kb * Ha + ka
    // single device: c_block = Nc, q_stride = Nc Pb Ha (plain [q][c][kb][ka]); distributed: one block per source rank
    int c_block;
    std::size_t block_stride, q_stride;
    // Mirror symmetries of the tensor spectrum of one sublattice on an orthorhombic, axis-aligned lattice: D_aa is even in
    // every k, D_ab (a != b) odd in k_a and k_b. Bit 0: only kc <= Pc / 2 is stored, bit 1: only kb <= Pb / 2 (one device);
    // the multiply kernels read the mirrored entry with the sign of the component. Verified numerically at setup.
    int mirror;
};
__device__ __host__ __forceinline__ int mirror_len( int P, bool on )
{
    return on ? P / 2 + 1 : P;
}
__device__ __forceinline__ std::size_t c_operand( const DDIDims & d, int q, int c )
{
    return std::size_t( c / d.c_block ) * d.block_stride + std::size_t( q ) * d.q_stride + std::size_t( c % d.c_block ) * ( std::size_t( d.Pb ) * d.Ha );
}

// 1: forward a-pass straight from the spin field. One CTA per (row = b + Nb c, component q = comp + 3 ib).
static __global__ void __launch_bounds__( FFT_THREADS ) k_ddi_fwd_a(
    const __grid_constant__ FFTPlan1D plan, const __grid_constant__ DDIDims d, ConstField3 spins, double2 * __restrict__ A )
{
    extern __shared__ double2 smem[];
    const int n   = plan.n;
    double2 * x   = smem;
    double2 * y   = smem + n;
    const int row = blockIdx.x; // b + Nb * c
    const int ib  = blockIdx.y;
    const int b = row % d.Nb, c = row / d.Nb;
    // the three components of basis atom ib: three transforms per CTA, one after the other (same row of the field)
    for( int comp = 0; comp < 3; ++comp )
    {
        for( int j = threadIdx.x; j < n; j += blockDim.x )
        {
            double v = 0.0;
            if( j < d.Na )
            {
                const std::size_t idx = std::size_t( ib + d.NB * j ) + std::size_t( d.Na ) * d.NB * b + std::size_t( d.plane_stride ) * ( c + d.halo );
                v                     = __ldg( spins.base + elem_offset( idx ) + comp * FIELD_BLOCK ) * d.mu_s[ib];
            }
            x[j] = make_double2( v, 0.0 );
        }
        __syncthreads();
        const double2 * r = block_fft<false>( plan, x, y, 1 );
        const int q       = comp + 3 * ib;
        double2 * out     = A + ( ( std::size_t( q ) * d.Nc + c ) * d.Nb + b ) * d.Ha;
        for( int j = threadIdx.x; j < d.Ha; j += blockDim.x )
            out[j] = r[j];
        __syncthreads();
    }
}

// 3: per (kb, ka-tile): forward c-transforms of all 3 NB components, tensor multiply, inverse c-transforms.
// B layout [q][c][kb][ka]; D^ layout [t][kc][kb][ka], t = comp6 + 6 * inter.
static __global__ void __launch_bounds__( FFT_THREADS ) k_ddi_c_mult(
    const __grid_constant__ FFTPlan1D plan, const __grid_constant__ DDIDims d, double2 * __restrict__ B,
    const double2 * __restrict__ Dhat, int ncol )
{
    extern __shared__ double2 smem[];
    const int n  = plan.n; // Pc
    const int nq = 3 * d.NB;
    // per component: two buffers of n * ncol
    const std::size_t buf = std::size_t( n ) * ncol;
    const int kb = blockIdx.y, u0 = blockIdx.x * ncol;
    const int nc = min( ncol, d.Ha - u0 );
    // load + forward transform every component
    __shared__ int result_in_y; // all transforms have the same number of stages: same final buffer
    for( int q = 0; q < nq; ++q )
    {
        double2 * x = smem + ( 2 * q ) * buf;
        for( int item = threadIdx.x; item < n * ncol; item += blockDim.x )
        {
            const int col = item % ncol, j = item / ncol;
            double2 v     = make_double2( 0.0, 0.0 );
            if( j < d.Nc && col < nc )
                v = B[c_operand( d, q, j ) + std::size_t( kb ) * d.Ha + u0 + col];
            x[item] = v;
        }
    }
    __syncthreads();
    for( int q = 0; q < nq; ++q )
    {
        double2 * x       = smem + ( 2 * q ) * buf;
        double2 * y       = x + buf;
        const double2 * r = block_fft<false>( plan, x, y, ncol );
        if( threadIdx.x == 0 && q == 0 )
            result_in_y = ( r == y ) ? 1 : 0;
    }
    __syncthreads();
    const int in_y = result_in_y;
    // multiply: F_{b1} = sum_{b2} D^(b1,b2) S_{b2}; results go to the OTHER buffer of each component
    for( int item = threadIdx.x; item < n * ncol; item += blockDim.x )
    {
        const int col = item % ncol, kc = item / ncol;
        if( col >= nc )
            continue;
        const std::size_t dk = ( std::size_t( kc ) * d.Pb + kb ) * d.Ha + u0 + col;
        const std::size_t dcomp = std::size_t( d.Pc ) * d.Pb * d.Ha;
        for( int b1 = 0; b1 < d.NB; ++b1 )
        {
            double2 fx = make_double2( 0, 0 ), fy = fx, fz = fx;
            for( int b2 = 0; b2 < d.NB; ++b2 )
            {
                const int inter    = d.lookup[b1 + b2 * d.NB];
                const double2 * Dp = Dhat + std::size_t( 6 * inter ) * dcomp + dk;
                const double2 Dxx = __ldg( Dp ), Dxy = __ldg( Dp + dcomp ), Dxz = __ldg( Dp + 2 * dcomp );
                const double2 Dyy = __ldg( Dp + 3 * dcomp ), Dyz = __ldg( Dp + 4 * dcomp ), Dzz = __ldg( Dp + 5 * dcomp );
                const double2 sx = smem[( 2 * ( 0 + 3 * b2 ) + in_y ) * buf + item];
                const double2 sy = smem[( 2 * ( 1 + 3 * b2 ) + in_y ) * buf + item];
                const double2 sz = smem[( 2 * ( 2 + 3 * b2 ) + in_y ) * buf + item];
                fx = cadd( fx, cadd( cmul( Dxx, sx ), cadd( cmul( Dxy, sy ), cmul( Dxz, sz ) ) ) );
                fy = cadd( fy, cadd( cmul( Dxy, sx ), cadd( cmul( Dyy, sy ), cmul( Dyz, sz ) ) ) );
                fz = cadd( fz, cadd( cmul( Dxz, sx ), cadd( cmul( Dyz, sy ), cmul( Dzz, sz ) ) ) );
            }
            smem[( 2 * ( 0 + 3 * b1 ) + 1 - in_y ) * buf + item] = fx;
            smem[( 2 * ( 1 + 3 * b1 ) + 1 - in_y ) * buf + item] = fy;
            smem[( 2 * ( 2 + 3 * b1 ) + 1 - in_y ) * buf + item] = fz;
        }
    }
    __syncthreads();
    // inverse transforms and store c < Nc
    for( int q = 0; q < nq; ++q )
    {
        double2 * x       = smem + ( 2 * q + 1 - in_y ) * buf;
        double2 * y       = smem + ( 2 * q + in_y ) * buf;
        const double2 * r = block_fft<true>( plan, x, y, ncol );
        for( int item = threadIdx.x; item < d.Nc * ncol; item += blockDim.x )
        {
            const int col = item % ncol, j = item / ncol;
            if( col < nc )
                B[c_operand( d, q, j ) + std::size_t( kb ) * d.Ha + u0 + col] = r[item];
        }
        __syncthreads();
    }
}


// 3 for one sublattice and a FFTPlan1D::fast16 length Pc: the three components transformed IN PLACE (block_fft16), the
// tensor multiply in place as well (one sublattice: F(k) needs S(k) of the same point only), (Pc / 8) << lg_ncol threads.
// The tensor spectrum is stored the way this kernel walks it, D^t[kb][ka tile][comp6][kc][col]: one contiguous block per
// CTA, read with unit stride (REAL_D: as doubles, the spectrum of a single sublattice is real).
template<bool REAL_D, int LOGN>
static __global__ void __launch_bounds__( FFT_THREADS, SB_FFT16_MINB ) k_ddi_c_mult16(
    const __grid_constant__ FFTPlan1D plan, const __grid_constant__ DDIDims d, double2 * __restrict__ B,
    const void * __restrict__ Dt_v, const int lg_ncol )
{
    extern __shared__ double2 smem[];
    constexpr int n = 1 << LOGN;
    const int ncol  = 1 << lg_ncol;
    const int tile_elems   = n << lg_ncol;
    const int bufp         = tile_elems + ( tile_elems >> 4 ) + 1; // room for the padded layout between the stages
    const int kb = blockIdx.y, u0 = blockIdx.x << lg_ncol;
    const int col = threadIdx.x & ( ncol - 1 ), j0 = threadIdx.x >> lg_ncol, jstep = blockDim.x >> lg_ncol;
    const bool valid       = u0 + col < d.Ha;
    const std::size_t plane = std::size_t( d.Pb ) * d.Ha;
    double2 * column       = B + std::size_t( kb ) * d.Ha + u0 + col;
    // mirrored tensor (REAL_D only): row kbr of the stored spectrum, n_kc entries along kc, signs of the odd components
    const bool mir_c = REAL_D && ( d.mirror & 1 ), mir_b = REAL_D && ( d.mirror & 2 ) && 2 * kb > d.Pb;
    const int kbr    = mir_b ? d.Pb - kb : kb;
    const double sgb = mir_b ? -1.0 : 1.0;
    const int tile_d = mirror_len( n, mir_c ) << lg_ncol; // elements of one component of the CTA's tensor block
    // pull the tensor block of this CTA towards L2 while the spectra are loaded and transformed
    {
        const std::size_t elem  = REAL_D ? sizeof( double ) : sizeof( double2 );
        const std::size_t bytes = 6 * std::size_t( tile_d ) * elem;
        const char * Dblock     = static_cast<const char *>( Dt_v ) + ( std::size_t( kbr ) * gridDim.x + blockIdx.x ) * bytes;
        for( std::size_t off = std::size_t( threadIdx.x ) * 128; off < bytes; off += std::size_t( blockDim.x ) * 128 )
            asm volatile( "prefetch.global.L2 [%0];" ::"l"( Dblock + off ) );
    }
    // element (q, c) of the column = c_operand; exactly FFT_E planes c = j0 + k jstep per thread and component
    // (recomputed where it is used: eight 64-bit offsets held across the transforms cost more in spills than this arithmetic)
    auto off_c = [&]( int k ) -> std::size_t
    {
        const int c = j0 + k * jstep;
        if( d.block_stride == 0 )
            return std::size_t( c ) * plane;
        const int blk = c / d.c_block;
        return std::size_t( blk ) * d.block_stride + std::size_t( c - blk * d.c_block ) * plane;
    };
    for( int q = 0; q < 3; ++q )
    {
        double2 * x            = smem + q * bufp;
        const double2 * column_q = column + std::size_t( q ) * d.q_stride;
        double2 v[FFT_E];
#pragma unroll
        for( int k = 0; k < FFT_E; ++k )
        {
            v[k] = make_double2( 0.0, 0.0 );
            if( valid && j0 + k * jstep < d.Nc )
                v[k] = column_q[off_c( k )];
        }
#pragma unroll
        for( int k = 0; k < FFT_E; ++k )
            x[( ( j0 + k * jstep ) << lg_ncol ) + col] = v[k];
    }
    __syncthreads();
#pragma unroll 1
    for( int q = 0; q < 3; ++q )
        block_fft_ct<false, LOGN>( plan.twiddle, smem + q * bufp, lg_ncol, col, j0 );
    {
        const std::size_t block = ( std::size_t( kbr ) * gridDim.x + blockIdx.x ) * 6 * std::size_t( tile_d );
#pragma unroll 4
        for( int item = threadIdx.x; item < tile_elems; item += blockDim.x )
        {
            const double2 sx = smem[item], sy = smem[bufp + item], sz = smem[2 * bufp + item];
            double2 fx, fy, fz;
            if( REAL_D )
            {
                const int kc       = item >> lg_ncol;
                const bool up      = mir_c && 2 * kc > n;
                const int item_d   = up ? ( ( n - kc ) << lg_ncol ) + ( item & ( ncol - 1 ) ) : item;
                const double sgc   = up ? -1.0 : 1.0;
                const double * Dp = static_cast<const double *>( Dt_v ) + block + item_d;
                const double Dxx = __ldg( Dp ), Dxy = sgb * __ldg( Dp + tile_d ), Dxz = sgc * __ldg( Dp + 2 * tile_d );
                const double Dyy = __ldg( Dp + 3 * tile_d ), Dyz = sgb * sgc * __ldg( Dp + 4 * tile_d ), Dzz = __ldg( Dp + 5 * tile_d );
                fx = make_double2( Dxx * sx.x + Dxy * sy.x + Dxz * sz.x, Dxx * sx.y + Dxy * sy.y + Dxz * sz.y );
                fy = make_double2( Dxy * sx.x + Dyy * sy.x + Dyz * sz.x, Dxy * sx.y + Dyy * sy.y + Dyz * sz.y );
                fz = make_double2( Dxz * sx.x + Dyz * sy.x + Dzz * sz.x, Dxz * sx.y + Dyz * sy.y + Dzz * sz.y );
            }
            else
            {
                const double2 * Dp = static_cast<const double2 *>( Dt_v ) + block + item;
                const double2 Dxx = __ldg( Dp ), Dxy = __ldg( Dp + tile_elems ), Dxz = __ldg( Dp + 2 * tile_elems );
                const double2 Dyy = __ldg( Dp + 3 * tile_elems ), Dyz = __ldg( Dp + 4 * tile_elems ), Dzz = __ldg( Dp + 5 * tile_elems );
                fx = cadd( cmul( Dxx, sx ), cadd( cmul( Dxy, sy ), cmul( Dxz, sz ) ) );
                fy = cadd( cmul( Dxy, sx ), cadd( cmul( Dyy, sy ), cmul( Dyz, sz ) ) );
                fz = cadd( cmul( Dxz, sx ), cadd( cmul( Dyz, sy ), cmul( Dzz, sz ) ) );
            }
            smem[item]            = fx;
            smem[bufp + item]     = fy;
            smem[2 * bufp + item] = fz;
        }
    }
    __syncthreads();
#pragma unroll 1
    for( int q = 0; q < 3; ++q )
        block_fft_ct<true, LOGN>( plan.twiddle, smem + q * bufp, lg_ncol, col, j0 );
    if( valid )
        for( int q = 0; q < 3; ++q )
        {
            const double2 * x  = smem + q * bufp;
            double2 * column_q = column + std::size_t( q ) * d.q_stride;
#pragma unroll
            for( int k = 0; k < FFT_E; ++k )
                if( j0 + k * jstep < d.Nc )
                    column_q[off_c( k )] = x[( ( j0 + k * jstep ) << lg_ncol ) + col];
        }
}

// ---------------------------------------------------------------------------------------------
// Halves of a stage, for kernels that keep the inputs of a transform's first stage or the outputs of its last stage in
// registers instead of passing them through shared memory: with the thread mapping of fft_stage_ct (thread tt of a column owns
// the radix-8 butterfly tt of every stage) the FFT_E elements j = tt + (n / 8) m, m < 8, are
//   * exactly the inputs of the thread's first-stage butterfly (a = tt + STEP i, STEP = n / 8), and
//   * exactly the outputs of its last-stage butterflies (o = item + (n / R) i with item = tt + (n / 8) jj: m = jj + (8 / R) i),
// forward and inverse alike. A pass that loads its column as "element m of thread tt" can therefore start the first butterfly on
// the loaded registers and store the last butterflies' results straight to their destination: two of the 2 x stages + 2
// shared-memory round trips and two CTA barriers per transform disappear, and the c-pass, whose inverse transform starts where
// the forward one ended, can do forward last stage -> tensor multiply -> inverse first stage without leaving the registers.
// ---------------------------------------------------------------------------------------------
template<int LOGN, int R, bool FIRST>
__device__ __forceinline__ void stage_fetch( const double2 * x, const int lg_ncol, const int col, const int tt, double2 ( &v )[FFT_E / R][R] )
{
    constexpr int N = 1 << LOGN, PER = FFT_E / R, PER_COL = N >> FFT_LG_E, STEP = N / R;
#pragma unroll
    for( int jj = 0; jj < PER; ++jj )
#pragma unroll
        for( int i = 0; i < R; ++i )
        {
            const int a   = tt + PER_COL * jj + STEP * i;
            const int idx = FIRST ? a : a + ( a >> 4 );
            v[jj][i]      = x[( idx << lg_ncol ) + col];
        }
}
// twiddles and stores of a stage whose butterflies (dft_small) have been done in v. No barrier: the caller orders it after
// the reads of the same buffer.
template<bool INVERSE, int LOGN, int R, int LG_S, bool LAST>
__device__ __forceinline__ void stage_emit(
    const double2 * __restrict__ twiddle, double2 * x, const int lg_ncol, const int col, const int tt, const double2 ( &v )[FFT_E / R][R] )
{
    constexpr int N = 1 << LOGN, PER = FFT_E / R, PER_COL = N >> FFT_LG_E, S = 1 << LG_S;
    constexpr bool CHAIN = LOGN >= SB_FFT_TW_CHAIN_LOGN;
#pragma unroll
    for( int jj = 0; jj < PER; ++jj )
    {
        const int item = tt + PER_COL * jj;
        const int q = item & ( S - 1 ), p = item >> LG_S;
        double2 w1 = make_double2( 1.0, 0.0 ), wc = make_double2( 1.0, 0.0 );
        if( !LAST && CHAIN )
        {
            w1 = __ldg( twiddle + ( p << LG_S ) );
            if( INVERSE )
                w1.y = -w1.y;
        }
#pragma unroll
        for( int i = 0; i < R; ++i )
        {
            double2 val = v[jj][i];
            if( !LAST && i > 0 )
            {
                if( CHAIN )
                    wc = i == 1 ? w1 : cmul( wc, w1 );
                else
                {
                    const double2 w = __ldg( twiddle + ( ( i * p ) << LG_S ) );
                    wc              = INVERSE ? make_double2( w.x, -w.y ) : w;
                }
                val = cmul( val, wc );
            }
            const int o   = q + S * ( R * p + i );
            const int idx = LAST ? o : o + ( o >> 4 );
            x[( idx << lg_ncol ) + col] = val;
        }
    }
}
// the radix-8 stages LG_S, LG_S + 3, ... that are neither the first nor the last of the transform
template<bool INVERSE, int LOGN, int LG_S, int LG_S_LAST>
__device__ __forceinline__ void middle_stages_ct( const double2 * __restrict__ twiddle, double2 * x, const int lg_ncol, const int col, const int tt )
{
    if constexpr( LG_S < LG_S_LAST )
    {
        fft_stage_ct<INVERSE, LOGN, FFT_E, LG_S, false, false>( twiddle, x, lg_ncol, col, tt );
        middle_stages_ct<INVERSE, LOGN, LG_S + FFT_LG_E, LG_S_LAST>( twiddle, x, lg_ncol, col, tt );
    }
}
// radix and first stride exponent of the last stage of block_fft_ct's plan
template<int LOGN>
struct LastStage
{
    static constexpr int REM = LOGN % FFT_LG_E, LG_R = REM ? REM : FFT_LG_E, R = 1 << LG_R, LG_S = LOGN - LG_R, PER = FFT_E / R;
};

// CTA shape the register-resident c-pass is compiled for, per length (measured, profiles/r2ze): up to 512 elements two columns
// per CTA are 128 threads, three CTAs per SM leave 168 registers (the 24 complex values fit without spills; a 128-register cap
// spills 250-300 bytes per thread); 1024 elements need 256 threads for two columns (one column per CTA moves half sectors:
// 512^3 c-pass 26 -> 32 ms), two CTAs per SM, 128 registers with those spills; longer columns: one column per CTA.
#ifndef SB_CF_MINB
#define SB_CF_MINB 3
#endif
template<int LOGN>
struct CFBounds
{
    static constexpr int T    = LOGN <= 9 ? 128 : ( LOGN <= 11 ? 256 : 512 );
    static constexpr int MINB = LOGN <= 9 ? SB_CF_MINB : ( LOGN <= 11 ? 2 : 1 );
};
inline int cf_threads( int n )
{
    return n <= 512 ? 128 : ( n <= 2048 ? 256 : 512 );
}

// k_fft_pass16 with the register-resident first / last stages: a thread loads the elements j0 + kk jstep, kk < 8 NSEQ, of its
// column (jstep = n / (8 NSEQ)), which are the inputs of the first-stage butterflies tt_b = j0 + b jstep, b < NSEQ (kk = NSEQ i + b),
// of that column's group; the last stage stores its results from the registers, one group after the other.
template<bool INVERSE, int LOGN, int LG_SEQ>
static __global__ void __launch_bounds__( FFT_THREADS, SB_FFT16_MINB ) k_fft_pass16r(
    const __grid_constant__ FFTPlan1D plan, const __grid_constant__ PassArgs a, const int lg_ncol, const int lg_in_split,
    const int lg_out_split )
{
    extern __shared__ double2 smem[];
    if constexpr( FFT_E != 8 )
        return;
    using L            = LastStage<LOGN>;
    constexpr int NSEQ = 1 << LG_SEQ, PER_COL = ( 1 << LOGN ) >> FFT_LG_E;
    const int lg_tile  = lg_ncol + LG_SEQ;
    const int bufp     = ( ( 1 << LOGN ) << lg_ncol ) + ( ( ( 1 << LOGN ) << lg_ncol ) >> 4 ) + 1;
    const int o = blockIdx.y, u0 = blockIdx.x << lg_tile;
    {
        const int ctile = threadIdx.x & ( ( 1 << lg_tile ) - 1 ), j0 = threadIdx.x >> lg_tile, jstep = blockDim.x >> lg_tile;
        double2 * mine  = smem + ( ctile >> lg_ncol ) * bufp;
        const bool valid   = u0 + ctile < a.n_u;
        const double2 * in = a.in + std::size_t( o ) * a.in_os + u0 + ctile;
        double2 v[NSEQ][1][FFT_E];
#pragma unroll
        for( int i = 0; i < FFT_E; ++i )
#pragma unroll
            for( int b = 0; b < NSEQ; ++b )
            {
                const int j = j0 + ( NSEQ * i + b ) * jstep;
                v[b][0][i]  = make_double2( 0.0, 0.0 );
                if( valid && j < a.n_in )
                {
                    if( a.use_in_peer )
                        v[b][0][i] = a.in_peer[j >> lg_in_split]
                                              [std::size_t( o ) * a.in_os + u0 + ctile + std::size_t( unsigned( j ) & ( ( 1u << lg_in_split ) - 1u ) ) * a.in_js];
                    else
                        v[b][0][i] = in[pass_offset16( j, a.in_js, lg_in_split, a.in_split_stride )];
                }
            }
#pragma unroll
        for( int b = 0; b < NSEQ; ++b )
        {
            dft_small<INVERSE, FFT_E>( v[b][0] );
            stage_emit<INVERSE, LOGN, FFT_E, 0, false>( plan.twiddle, mine, lg_ncol, ctile & ( ( 1 << lg_ncol ) - 1 ), j0 + b * jstep, v[b] );
        }
    }
    __syncthreads();
    const int col = int( threadIdx.x ) & ( ( 1 << lg_ncol ) - 1 ), tt = int( threadIdx.x ) >> lg_ncol;
#pragma unroll 1
    for( int g = 0; g < NSEQ; ++g )
        middle_stages_ct<INVERSE, LOGN, FFT_LG_E, L::LG_S>( plan.twiddle, smem + g * bufp, lg_ncol, col, tt );
    // last stage with the thread mapping of the load phase (butterflies tt_b of the thread's own column): a warp's store
    // instruction covers whole rows of the tile, (ncol << LG_SEQ) x 16 contiguous bytes
    {
        const int ctile = threadIdx.x & ( ( 1 << lg_tile ) - 1 ), j0 = threadIdx.x >> lg_tile, jstep = blockDim.x >> lg_tile;
        const double2 * mine = smem + ( ctile >> lg_ncol ) * bufp;
        const bool valid     = u0 + ctile < a.n_u;
        double2 * out        = a.out + std::size_t( o ) * a.out_os + u0 + ctile;
        if( a.opeer_planes )
        {
            const int q = o / a.opeer_planes, c = o - q * a.opeer_planes, r = c / a.opeer_ncl;
            out         = a.out_peer[r] + std::size_t( q * a.opeer_ncl + ( c - r * a.opeer_ncl ) ) * a.out_os + u0 + ctile;
        }
#pragma unroll 1
        for( int b = 0; b < NSEQ; ++b )
        {
            const int tb = j0 + b * jstep;
            double2 v[L::PER][L::R];
            stage_fetch<LOGN, L::R, false>( mine, lg_ncol, ctile & ( ( 1 << lg_ncol ) - 1 ), tb, v );
#pragma unroll
            for( int jj = 0; jj < L::PER; ++jj )
                dft_small<INVERSE, L::R>( v[jj] );
            if( valid )
            {
#pragma unroll
                for( int jj = 0; jj < L::PER; ++jj )
#pragma unroll
                    for( int i = 0; i < L::R; ++i )
                    {
                        const int j = tb + ( jj + L::PER * i ) * PER_COL;
                        if( j < a.n_out )
                        {
                            const double2 w = make_double2( a.scale * v[jj][i].x, a.scale * v[jj][i].y );
                            if( a.use_out_peer )
                                a.out_peer[j >> lg_out_split]
                                          [std::size_t( o ) * a.out_os + u0 + ctile + std::size_t( unsigned( j ) & ( ( 1u << lg_out_split ) - 1u ) ) * a.out_js]
                                    = w;
                            else
                                out[pass_offset16( j, a.out_js, lg_out_split, a.out_split_stride )] = w;
                        }
                    }
            }
        }
    }
}

// k_ddi_c_mult16 with the register-resident first / last stages described above: the first forward stage works on the loaded
// column, the last forward stage, the tensor multiply and the first inverse stage of the three components stay in registers
// (24 complex values per thread), the last inverse stage stores to global memory. Same launch shape, same tensor layout,
// the same arithmetic per butterfly (only where its operands wait differs).
template<bool REAL_D, int LOGN>
static __global__ void __launch_bounds__( CFBounds<LOGN>::T, CFBounds<LOGN>::MINB ) k_ddi_c_mult16f(
        const __grid_constant__ FFTPlan1D plan, const __grid_constant__ DDIDims d, double2 * __restrict__ B, const void * __restrict__ Dt_v,
        const int lg_ncol )
{
    extern __shared__ double2 smem[];
    if constexpr( FFT_E != 8 ) // (tuning builds with other radices keep k_ddi_c_mult16)
        return;
    using L               = LastStage<LOGN>;
    constexpr int n       = 1 << LOGN, PER_COL = n >> FFT_LG_E;
    const int ncol        = 1 << lg_ncol;
    const int tile_elems  = n << lg_ncol;
    const int bufp        = tile_elems + ( tile_elems >> 4 ) + 1;
    const int kb = blockIdx.y, u0 = blockIdx.x << lg_ncol;
    const int col = threadIdx.x & ( ncol - 1 ), tt = threadIdx.x >> lg_ncol; // blockDim.x = PER_COL << lg_ncol
    const bool valid        = u0 + col < d.Ha;
    const std::size_t plane = std::size_t( d.Pb ) * d.Ha;
    double2 * column        = B + std::size_t( kb ) * d.Ha + u0 + col;
    const bool mir_c = REAL_D && ( d.mirror & 1 ), mir_b = REAL_D && ( d.mirror & 2 ) && 2 * kb > d.Pb;
    const int kbr    = mir_b ? d.Pb - kb : kb;
    const double sgb = mir_b ? -1.0 : 1.0;
    const int tile_d = mirror_len( n, mir_c ) << lg_ncol;
    const std::size_t block = ( std::size_t( kbr ) * gridDim.x + blockIdx.x ) * 6 * std::size_t( tile_d );
    {
        const std::size_t elem  = REAL_D ? sizeof( double ) : sizeof( double2 );
        const std::size_t bytes = 6 * std::size_t( tile_d ) * elem;
        const char * Dblock     = static_cast<const char *>( Dt_v ) + block * elem;
        for( std::size_t off = std::size_t( threadIdx.x ) * 128; off < bytes; off += std::size_t( blockDim.x ) * 128 )
            asm volatile( "prefetch.global.L2 [%0];" ::"l"( Dblock + off ) );
    }
    auto off_c = [&]( int m ) -> std::size_t
    {
        const int c = tt + m * PER_COL;
        if( d.block_stride == 0 )
            return std::size_t( c ) * plane;
        const int blk = c / d.c_block;
        return std::size_t( blk ) * d.block_stride + std::size_t( c - blk * d.c_block ) * plane;
    };
    // forward, first stage: straight from global memory (planes c >= Nc are zero)
#pragma unroll 1
    for( int q = 0; q < 3; ++q )
    {
        const double2 * column_q = column + std::size_t( q ) * d.q_stride;
        double2 v[1][FFT_E];
#pragma unroll
        for( int m = 0; m < FFT_E; ++m )
        {
            v[0][m] = make_double2( 0.0, 0.0 );
            if( valid && tt + m * PER_COL < d.Nc )
                v[0][m] = column_q[off_c( m )];
        }
        dft_small<false, FFT_E>( v[0] );
        stage_emit<false, LOGN, FFT_E, 0, false>( plan.twiddle, smem + q * bufp, lg_ncol, col, tt, v );
    }
    __syncthreads();
#pragma unroll 1
    for( int q = 0; q < 3; ++q )
        middle_stages_ct<false, LOGN, FFT_LG_E, L::LG_S>( plan.twiddle, smem + q * bufp, lg_ncol, col, tt );
    // the tensor at the thread's point m of the column: kc = tt + m n / 8
    auto multiply = [&]( const int m, double2 & sx, double2 & sy, double2 & sz )
    {
        const int kc = tt + m * PER_COL;
        if( REAL_D )
        {
            const bool up     = mir_c && 2 * kc > n;
            const int item_d  = ( ( up ? n - kc : kc ) << lg_ncol ) + col;
            const double sgc  = up ? -1.0 : 1.0;
            const double * Dp = static_cast<const double *>( Dt_v ) + block + item_d;
            const double Dxx = __ldg( Dp ), Dxy = sgb * __ldg( Dp + tile_d ), Dxz = sgc * __ldg( Dp + 2 * tile_d );
            const double Dyy = __ldg( Dp + 3 * tile_d ), Dyz = sgb * sgc * __ldg( Dp + 4 * tile_d ), Dzz = __ldg( Dp + 5 * tile_d );
            const double2 fx = make_double2( Dxx * sx.x + Dxy * sy.x + Dxz * sz.x, Dxx * sx.y + Dxy * sy.y + Dxz * sz.y );
            const double2 fy = make_double2( Dxy * sx.x + Dyy * sy.x + Dyz * sz.x, Dxy * sx.y + Dyy * sy.y + Dyz * sz.y );
            const double2 fz = make_double2( Dxz * sx.x + Dyz * sy.x + Dzz * sz.x, Dxz * sx.y + Dyz * sy.y + Dzz * sz.y );
            sx = fx, sy = fy, sz = fz;
        }
        else
        {
            const double2 * Dp = static_cast<const double2 *>( Dt_v ) + block + ( kc << lg_ncol ) + col;
            const double2 Dxx = __ldg( Dp ), Dxy = __ldg( Dp + tile_elems ), Dxz = __ldg( Dp + 2 * tile_elems );
            const double2 Dyy = __ldg( Dp + 3 * tile_elems ), Dyz = __ldg( Dp + 4 * tile_elems ), Dzz = __ldg( Dp + 5 * tile_elems );
            const double2 fx = cadd( cmul( Dxx, sx ), cadd( cmul( Dxy, sy ), cmul( Dxz, sz ) ) );
            const double2 fy = cadd( cmul( Dxy, sx ), cadd( cmul( Dyy, sy ), cmul( Dyz, sz ) ) );
            const double2 fz = cadd( cmul( Dxz, sx ), cadd( cmul( Dyz, sy ), cmul( Dzz, sz ) ) );
            sx = fx, sy = fy, sz = fz;
        }
    };
    // forward last stage into registers, multiply, inverse first stage out of them
    double2 V[3][L::PER][L::R];
#pragma unroll
    for( int q = 0; q < 3; ++q )
    {
        stage_fetch<LOGN, L::R, false>( smem + q * bufp, lg_ncol, col, tt, V[q] );
#pragma unroll
        for( int jj = 0; jj < L::PER; ++jj )
            dft_small<false, L::R>( V[q][jj] );
    }
    __syncthreads(); // every thread has its inputs: the buffers may be overwritten
#pragma unroll
    for( int jj = 0; jj < L::PER; ++jj )
#pragma unroll
        for( int i = 0; i < L::R; ++i )
            multiply( jj + L::PER * i, V[0][jj][i], V[1][jj][i], V[2][jj][i] );
#pragma unroll
    for( int q = 0; q < 3; ++q )
    {
        double2 w[1][FFT_E];
#pragma unroll
        for( int jj = 0; jj < L::PER; ++jj )
#pragma unroll
            for( int i = 0; i < L::R; ++i )
                w[0][jj + L::PER * i] = V[q][jj][i];
        dft_small<true, FFT_E>( w[0] );
        stage_emit<true, LOGN, FFT_E, 0, false>( plan.twiddle, smem + q * bufp, lg_ncol, col, tt, w );
    }
    __syncthreads();
#pragma unroll 1
    for( int q = 0; q < 3; ++q )
        middle_stages_ct<true, LOGN, FFT_LG_E, L::LG_S>( plan.twiddle, smem + q * bufp, lg_ncol, col, tt );
    // inverse, last stage: results straight to global memory, planes c < Nc only
#pragma unroll 1
    for( int q = 0; q < 3; ++q )
    {
        double2 * column_q = column + std::size_t( q ) * d.q_stride;
        double2 v[L::PER][L::R];
        stage_fetch<LOGN, L::R, false>( smem + q * bufp, lg_ncol, col, tt, v );
#pragma unroll
        for( int jj = 0; jj < L::PER; ++jj )
            dft_small<true, L::R>( v[jj] );
        if( valid )
        {
#pragma unroll
            for( int jj = 0; jj < L::PER; ++jj )
#pragma unroll
                for( int i = 0; i < L::R; ++i )
                {
                    const int m = jj + L::PER * i;
                    if( tt + m * PER_COL < d.Nc )
                        column_q[off_c( m )] = v[jj][i];
                }
        }
    }
}

// D^[comp6][kc][kb][ka] -> D^t[kb][ka tile][comp6][kc][col] (setup). Columns past Ha in the last tile stay zero.
template<bool REAL_D>
static __global__ void k_ddi_tile_tensor( const __grid_constant__ DDIDims d, const double2 * __restrict__ Dhat, void * __restrict__ Dt_v, const int lg_ncol )
{
    const std::size_t half = std::size_t( d.Pc ) * d.Pb * d.Ha;
    const std::size_t i    = blockIdx.x * std::size_t( blockDim.x ) + threadIdx.x;
    if( i >= 6 * half )
        return;
    const int ka = int( i % d.Ha ), kb = int( ( i / d.Ha ) % d.Pb ), kc = int( ( i / ( std::size_t( d.Ha ) * d.Pb ) ) % d.Pc );
    const int comp = int( i / half );
    const int ntiles = ( d.Ha + ( 1 << lg_ncol ) - 1 ) >> lg_ncol;
    // mirrored storage (REAL_D; d.mirror bit 1 only on one device, where d.Pb is the whole axis): the upper halves are dropped
    if( REAL_D && ( ( ( d.mirror & 1 ) && 2 * kc > d.Pc ) || ( ( d.mirror & 2 ) && 2 * kb > d.Pb ) ) )
        return;
    const std::size_t tile_elems = std::size_t( mirror_len( d.Pc, REAL_D && ( d.mirror & 1 ) ) ) << lg_ncol;
    const std::size_t o = ( ( std::size_t( kb ) * ntiles + ( ka >> lg_ncol ) ) * 6 + comp ) * tile_elems + ( std::size_t( kc ) << lg_ncol )
                          + ( ka & ( ( 1 << lg_ncol ) - 1 ) );
    if( REAL_D )
        static_cast<double *>( Dt_v )[o] = Dhat[i].x;
    else
        static_cast<double2 *>( Dt_v )[o] = Dhat[i];
}

// 3': the same for small power-of-two Pc (thin films): ONE THREAD per (kb, ka) column does the length-PC transforms of
// all three components in registers -- no shared memory, no barriers, fully coalesced along ka. REAL_D: the tensor
// spectrum is real (single sublattice: D(-r) = D(r)) and stored as doubles, halving the dominant read stream.
template<int PC, bool INVERSE>
__device__ __forceinline__ void reg_fft( double2 ( &v )[PC] )
{
    // iterative radix-2 decimation in frequency, output in bit-reversed order -> reorder at the end
#pragma unroll
    for( int half = PC / 2; half >= 1; half /= 2 )
    {
#pragma unroll
        for( int base = 0; base < PC; base += 2 * half )
        {
#pragma unroll
            for( int j = 0; j < half; ++j )
            {
                const double2 a = v[base + j], b = v[base + j + half];
                v[base + j]     = cadd( a, b );
                const double2 d = csub( a, b );
                // twiddle exp(-+ 2 pi i j / (2 half)) (constant-bank operand after unrolling)
                const double2 t    = TW32[j * ( 16 / half )];
                v[base + j + half] = cmul( d, make_double2( t.x, INVERSE ? -t.y : t.y ) );
            }
        }
    }
    // bit reversal
    double2 w[PC];
#pragma unroll
    for( int i = 0; i < PC; ++i )
    {
        int r = 0;
#pragma unroll
        for( int bit = 1, rb = PC / 2; bit < PC; bit *= 2, rb /= 2 )
            if( i & bit )
                r |= rb;
        w[r] = v[i];
    }
#pragma unroll
    for( int i = 0; i < PC; ++i )
        v[i] = w[i];
}

template<int PC, bool REAL_D>
static __global__ void __launch_bounds__( 128 ) k_ddi_c_mult_small(
    const __grid_constant__ DDIDims d, double2 * __restrict__ B, const void * __restrict__ Dhat_v )
{
    const int ka = blockIdx.x * blockDim.x + threadIdx.x;
    const int kb = blockIdx.y;
    if( ka >= d.Ha )
        return;
    const std::size_t c_stride = std::size_t( d.Pb ) * d.Ha;
    const std::size_t col      = std::size_t( kb ) * d.Ha + ka;
    // tensor: [comp6][kc][kb][ka], with REAL_D possibly mirrored (DDIDims::mirror): [comp6][kc <= PC/2][kb <= Pb/2][ka]
    const bool mir_c = REAL_D && ( d.mirror & 1 ), mir_b = REAL_D && ( d.mirror & 2 ) && 2 * kb > d.Pb;
    const double sgb = mir_b ? -1.0 : 1.0;
    const std::size_t d_cstride = std::size_t( mirror_len( d.Pb, REAL_D && ( d.mirror & 2 ) ) ) * d.Ha;
    const std::size_t d_col     = std::size_t( mir_b ? d.Pb - kb : kb ) * d.Ha + ka;
    const std::size_t dcomp     = std::size_t( mirror_len( PC, mir_c ) ) * d_cstride;
    double2 sx[PC], sy[PC], sz[PC];
#pragma unroll
    for( int j = 0; j < PC; ++j )
    {
        const double2 zero = make_double2( 0.0, 0.0 );
        sx[j] = j < d.Nc ? B[c_operand( d, 0, j ) + col] : zero;
        sy[j] = j < d.Nc ? B[c_operand( d, 1, j ) + col] : zero;
        sz[j] = j < d.Nc ? B[c_operand( d, 2, j ) + col] : zero;
    }
    reg_fft<PC, false>( sx );
    reg_fft<PC, false>( sy );
    reg_fft<PC, false>( sz );
#pragma unroll
    for( int kc = 0; kc < PC; ++kc )
    {
        const bool up        = mir_c && 2 * kc > PC;
        const double sgc     = up ? -1.0 : 1.0;
        const std::size_t dk = std::size_t( up ? PC - kc : kc ) * d_cstride + d_col;
        double2 fx, fy, fz;
        if( REAL_D )
        {
            const double * Dp = static_cast<const double *>( Dhat_v ) + dk;
            const double Dxx = __ldg( Dp ), Dxy = sgb * __ldg( Dp + dcomp ), Dxz = sgc * __ldg( Dp + 2 * dcomp );
            const double Dyy = __ldg( Dp + 3 * dcomp ), Dyz = sgb * sgc * __ldg( Dp + 4 * dcomp ), Dzz = __ldg( Dp + 5 * dcomp );
            fx = make_double2( Dxx * sx[kc].x + Dxy * sy[kc].x + Dxz * sz[kc].x, Dxx * sx[kc].y + Dxy * sy[kc].y + Dxz * sz[kc].y );
            fy = make_double2( Dxy * sx[kc].x + Dyy * sy[kc].x + Dyz * sz[kc].x, Dxy * sx[kc].y + Dyy * sy[kc].y + Dyz * sz[kc].y );
            fz = make_double2( Dxz * sx[kc].x + Dyz * sy[kc].x + Dzz * sz[kc].x, Dxz * sx[kc].y + Dyz * sy[kc].y + Dzz * sz[kc].y );
        }
        else
        {
            const double2 * Dp = static_cast<const double2 *>( Dhat_v ) + dk;
            const double2 Dxx = __ldg( Dp ), Dxy = __ldg( Dp + dcomp ), Dxz = __ldg( Dp + 2 * dcomp );
            const double2 Dyy = __ldg( Dp + 3 * dcomp ), Dyz = __ldg( Dp + 4 * dcomp ), Dzz = __ldg( Dp + 5 * dcomp );
            fx = cadd( cmul( Dxx, sx[kc] ), cadd( cmul( Dxy, sy[kc] ), cmul( Dxz, sz[kc] ) ) );
            fy = cadd( cmul( Dxy, sx[kc] ), cadd( cmul( Dyy, sy[kc] ), cmul( Dyz, sz[kc] ) ) );
            fz = cadd( cmul( Dxz, sx[kc] ), cadd( cmul( Dyz, sy[kc] ), cmul( Dzz, sz[kc] ) ) );
        }
        sx[kc] = fx;
        sy[kc] = fy;
        sz[kc] = fz;
    }
    reg_fft<PC, true>( sx );
    reg_fft<PC, true>( sy );
    reg_fft<PC, true>( sz );
#pragma unroll
    for( int j = 0; j < PC; ++j )
        if( j < d.Nc )
        {
            B[c_operand( d, 0, j ) + col] = sx[j];
            B[c_operand( d, 1, j ) + col] = sy[j];
            B[c_operand( d, 2, j ) + col] = sz[j];
        }
}

// max |Im D^| and max |D^| (is the tensor spectrum real?) -- partial maxima per block
static __global__ void k_ddi_imag_check( const double2 * __restrict__ D, std::size_t n, double * __restrict__ out )
{
    double mi = 0, ma = 0;
    for( std::size_t i = blockIdx.x * std::size_t( blockDim.x ) + threadIdx.x; i < n; i += std::size_t( gridDim.x ) * blockDim.x )
    {
        mi = fmax( mi, fabs( D[i].y ) );
        ma = fmax( ma, fmax( fabs( D[i].x ), fabs( D[i].y ) ) );
    }
    mi = block_max( mi );
    if( threadIdx.x == 0 )
        out[2 * blockIdx.x] = mi;
    ma = block_max( ma );
    if( threadIdx.x == 0 )
        out[2 * blockIdx.x + 1] = ma;
}
// real parts of D^[comp6][kc][kb][ka] -> [comp6][kc < n_kc][kb < n_kb][ka] (n_k = P / 2 + 1 along a mirrored axis, else P)
static __global__ void k_ddi_take_real( const __grid_constant__ DDIDims d, const double2 * __restrict__ in, double * __restrict__ out )
{
    const int n_kc = mirror_len( d.Pc, d.mirror & 1 ), n_kb = mirror_len( d.Pb, d.mirror & 2 );
    const std::size_t n = std::size_t( 6 ) * n_kc * n_kb * d.Ha;
    const std::size_t i = blockIdx.x * std::size_t( blockDim.x ) + threadIdx.x;
    if( i >= n )
        return;
    const int ka = int( i % d.Ha ), kb = int( ( i / d.Ha ) % n_kb ), kc = int( ( i / ( std::size_t( d.Ha ) * n_kb ) ) % n_kc );
    const int comp = int( i / ( std::size_t( d.Ha ) * n_kb * n_kc ) );
    out[i] = in[( ( std::size_t( comp ) * d.Pc + kc ) * d.Pb + kb ) * d.Ha + ka].x;
}
// max | D^(.., P - k, ..) - sign D^(.., k, ..) | for the mirror along c (out[2 blk]) and along b (out[2 blk + 1]), sign = -1 for
// the components that are odd along that axis (xz, yz along c; xy, yz along b). d.Pb must be the whole b axis for the b check.
static __global__ void k_ddi_mirror_check( const __grid_constant__ DDIDims d, const double2 * __restrict__ D, double * __restrict__ out )
{
    const std::size_t half = std::size_t( d.Pc ) * d.Pb * d.Ha;
    double mc = 0, mb = 0;
    for( std::size_t i = blockIdx.x * std::size_t( blockDim.x ) + threadIdx.x; i < 6 * half; i += std::size_t( gridDim.x ) * blockDim.x )
    {
        const int ka = int( i % d.Ha ), kb = int( ( i / d.Ha ) % d.Pb ), kc = int( ( i / ( std::size_t( d.Ha ) * d.Pb ) ) % d.Pc );
        const int comp = int( i / half );
        const double v = D[i].x;
        const double sc = ( comp == 2 || comp == 4 ) ? -1.0 : 1.0, sb = ( comp == 1 || comp == 4 ) ? -1.0 : 1.0;
        const std::size_t ic = ( ( std::size_t( comp ) * d.Pc + ( d.Pc - kc ) % d.Pc ) * d.Pb + kb ) * d.Ha + ka;
        const std::size_t ib = ( ( std::size_t( comp ) * d.Pc + kc ) * d.Pb + ( d.Pb - kb ) % d.Pb ) * d.Ha + ka;
        mc = fmax( mc, fabs( D[ic].x - sc * v ) );
        mb = fmax( mb, fabs( D[ib].x - sb * v ) );
    }
    mc = block_max( mc );
    if( threadIdx.x == 0 )
        out[2 * blockIdx.x] = mc;
    mb = block_max( mb );
    if( threadIdx.x == 0 )
        out[2 * blockIdx.x + 1] = mb;
}

// 5: inverse a-pass (C2R through a complex transform of the Hermitian-extended row) fused with
//    g_ddi = -mu_s res / P  (Hamiltonian_Heisenberg.cpp:995-1013), written as a field
static __global__ void __launch_bounds__( FFT_THREADS ) k_ddi_inv_a(
    const __grid_constant__ FFTPlan1D plan, const __grid_constant__ DDIDims d, const double2 * __restrict__ A, Field3 g, double inv_P )
{
    extern __shared__ double2 smem[];
    const int n   = plan.n;
    double2 * x   = smem;
    double2 * y   = smem + n;
    const int row = blockIdx.x;
    const int ib  = blockIdx.y;
    const int b = row % d.Nb, c = row / d.Nb;
    for( int comp = 0; comp < 3; ++comp )
    {
        const int q        = comp + 3 * ib;
        const double2 * in = A + ( ( std::size_t( q ) * d.Nc + c ) * d.Nb + b ) * d.Ha;
        for( int j = threadIdx.x; j < n; j += blockDim.x )
        {
            double2 v;
            if( j < d.Ha )
                v = in[j];
            else
            {
                v   = in[n - j];
                v.y = -v.y;
            }
            x[j] = v;
        }
        __syncthreads();
        const double2 * r = block_fft<true>( plan, x, y, 1 );
        const double f    = -d.mu_s[ib] * inv_P;
        for( int j = threadIdx.x; j < d.Na; j += blockDim.x )
        {
            const std::size_t idx = std::size_t( ib + d.NB * j ) + std::size_t( d.Na ) * d.NB * b + std::size_t( d.plane_stride ) * ( c + d.halo );
            g.base[elem_offset( idx ) + comp * FIELD_BLOCK] = f * r[j].x;
        }
        __syncthreads();
    }
}

// 1 and 5 for even Pa with a FFTPlan1D::fast16 half length m = Pa / 2: the real row as a complex sequence of half the
// length, z_j = x_2j + i x_2j+1, one complex transform of length m (plan_h), and the split / merge step
//   forward   X_k = E_k + w^k O_k,  E_k = (Z_k + conj Z_{m-k}) / 2,  O_k = -i (Z_k - conj Z_{m-k}) / 2,  w = exp(-2 pi i / Pa)
//   inverse   Z_k = (X_k + conj X_{m-k}) + i w^{-k} (X_k - conj X_{m-k})   ->   Pa x_2j = Re z_j, Pa x_2j+1 = Im z_j
// instead of a length-Pa complex transform with zero imaginary parts: half the butterflies, half the shared memory.
// A CTA takes 1 << lg_nrow consecutive rows (b + Nb c) of one component as the "columns" of block_fft16;
// (m / 8) << lg_nrow threads. Thread -> (4 consecutive elements of a row, row, group of 4): global accesses are 64-byte
// segments along the row, shared-memory accesses of a warp stay inside one 512-byte window.
__device__ __forceinline__ void row_map( int lg_nrow, int & row, int & jlow, int & jhigh, int & jhigh_step )
{
    const int t = threadIdx.x;
    jlow        = t & 3;
    row         = ( t >> 2 ) & ( ( 1 << lg_nrow ) - 1 );
    jhigh       = t >> ( 2 + lg_nrow );
    jhigh_step  = blockDim.x >> ( 2 + lg_nrow );
}

// Pencil decomposition over several GPUs: the ka axis of the half spectrum is cut into per-rank blocks of `kblock` (the last rank
// also takes ka = Pa / 2), and the forward a-pass stores element ka of the row (q, global plane, b) straight into the b-pass
// operand of the rank that owns it (peer-mapped memory, NVLink): AT_r[q][c_global][b][ka - r kblock], rows of w[r] elements.
struct APush
{
    int on, lg_kblock, world, planes, c_begin;
    int w[DDI_MAX_PEERS];
    double2 * base[DDI_MAX_PEERS];
};
__device__ __forceinline__ double2 * apush_dst( const APush & ap, int q, int c, int b, int Nb, int k )
{
    const int r = min( k >> ap.lg_kblock, ap.world - 1 );
    return ap.base[r] + ( ( std::size_t( q ) * ap.planes + ap.c_begin + c ) * Nb + b ) * ap.w[r] + ( k - ( r << ap.lg_kblock ) );
}

template<int LOGM>
static __global__ void __launch_bounds__( FFT_THREADS, SB_FFT16_MINB ) k_ddi_fwd_a16(
    const __grid_constant__ FFTPlan1D plan_h, const double2 * __restrict__ tw_full, const __grid_constant__ DDIDims d,
    ConstField3 spins, double2 * __restrict__ A, const int lg_nrow, const int q0, const __grid_constant__ APush ap )
{
    extern __shared__ double2 smem[];
    constexpr int m = 1 << LOGM;
    int rl, jlow, jhigh, jstep;
    row_map( lg_nrow, rl, jlow, jhigh, jstep );
    const int q = q0 + blockIdx.y, comp = q % 3, ib = q / 3;
    const int row    = ( blockIdx.x << lg_nrow ) + rl;
    const bool valid = row < d.Nb * d.Nc;
    const int b = row % d.Nb, c = row / d.Nb;
    const std::size_t site0 = std::size_t( d.Na ) * d.NB * b + std::size_t( d.plane_stride ) * ( c + d.halo ) + ib;
    const double mu         = d.mu_s[ib];
    // exactly FFT_E elements per thread
    double2 zz[FFT_E];
#pragma unroll
    for( int k = 0; k < FFT_E; ++k )
    {
        const int j = 4 * ( jhigh + k * jstep ) + jlow;
        zz[k]       = make_double2( 0.0, 0.0 );
        if( valid && 2 * j < d.Na )
        {
            zz[k].x = __ldg( spins.base + elem_offset( site0 + std::size_t( d.NB ) * ( 2 * j ) ) + comp * FIELD_BLOCK );
            if( 2 * j + 1 < d.Na )
                zz[k].y = __ldg( spins.base + elem_offset( site0 + std::size_t( d.NB ) * ( 2 * j + 1 ) ) + comp * FIELD_BLOCK );
        }
    }
#pragma unroll
    for( int k = 0; k < FFT_E; ++k )
        smem[( ( 4 * ( jhigh + k * jstep ) + jlow ) << lg_nrow ) + rl] = make_double2( mu * zz[k].x, mu * zz[k].y );
    __syncthreads();
    block_fft_ct<false, LOGM>( plan_h.twiddle, smem, lg_nrow, threadIdx.x & ( ( 1 << lg_nrow ) - 1 ), threadIdx.x >> lg_nrow );
    if( !valid )
        return;
    double2 * out = A + ( ( std::size_t( q ) * d.Nc + c ) * d.Nb + b ) * d.Ha;
    for( int jh = jhigh; 4 * jh < m; jh += jstep )
    {
        const int k      = 4 * jh + jlow;
        const double2 Zk = smem[( k << lg_nrow ) + rl];
        const double2 Zm = smem[( ( ( m - k ) & ( m - 1 ) ) << lg_nrow ) + rl]; // Z_m = Z_0
        const double2 E  = make_double2( 0.5 * ( Zk.x + Zm.x ), 0.5 * ( Zk.y - Zm.y ) );
        const double2 O  = make_double2( 0.5 * ( Zk.y + Zm.y ), -0.5 * ( Zk.x - Zm.x ) );
        const double2 X  = cadd( E, cmul( __ldg( tw_full + k ), O ) );
        if( ap.on )
            *apush_dst( ap, q, c, b, d.Nb, k ) = X;
        else
            out[k] = X;
    }
    if( jhigh == 0 && jlow == 0 )
    {
        const double2 Z0 = smem[rl];
        const double2 X  = make_double2( Z0.x - Z0.y, 0.0 );
        if( ap.on )
            *apush_dst( ap, q, c, b, d.Nb, m ) = X;
        else
            out[m] = X;
    }
}

template<int LOGM>
static __global__ void __launch_bounds__( FFT_THREADS, SB_FFT16_MINB ) k_ddi_inv_a16(
    const __grid_constant__ FFTPlan1D plan_h, const double2 * __restrict__ tw_full, const __grid_constant__ DDIDims d,
    const double2 * __restrict__ A, Field3 g, const double inv_P, const int lg_nrow, const int q0, const __grid_constant__ APush ap )
{
    extern __shared__ double2 smem[];
    constexpr int m = 1 << LOGM;
    int rl, jlow, jhigh, jstep;
    row_map( lg_nrow, rl, jlow, jhigh, jstep );
    const int q = q0 + blockIdx.y, comp = q % 3, ib = q / 3;
    const int row    = ( blockIdx.x << lg_nrow ) + rl;
    const bool valid = row < d.Nb * d.Nc;
    const int b = row % d.Nb, c = row / d.Nb;
    const double2 * in = A + ( ( std::size_t( q ) * d.Nc + c ) * d.Nb + b ) * d.Ha;
#pragma unroll
    for( int kk = 0; kk < FFT_E; ++kk )
    {
        const int k = 4 * ( jhigh + kk * jstep ) + jlow;
        double2 Z   = make_double2( 0.0, 0.0 );
        if( valid )
        {
            // pencil decomposition: element k of the row lives in the operand of the rank that owns the ka block (pulled over NVLink)
            const double2 Xk = ap.on ? *apush_dst( ap, q, c, b, d.Nb, k ) : in[k];
            const double2 Xm = ap.on ? *apush_dst( ap, q, c, b, d.Nb, m - k ) : in[m - k];
            const double2 S  = make_double2( Xk.x + Xm.x, Xk.y - Xm.y ); // X_k + conj X_{m-k}
            const double2 D  = make_double2( Xk.x - Xm.x, Xk.y + Xm.y ); // X_k - conj X_{m-k}
            const double2 w  = __ldg( tw_full + k );                     // exp(-2 pi i k / Pa); the inverse needs its conjugate
            const double2 T  = cmul( make_double2( w.x, -w.y ), D );
            Z                = make_double2( S.x - T.y, S.y + T.x ); // S + i T
        }
        smem[( k << lg_nrow ) + rl] = Z;
    }
    __syncthreads();
    block_fft_ct<true, LOGM>( plan_h.twiddle, smem, lg_nrow, threadIdx.x & ( ( 1 << lg_nrow ) - 1 ), threadIdx.x >> lg_nrow );
    if( !valid )
        return;
    const std::size_t site0 = std::size_t( d.Na ) * d.NB * b + std::size_t( d.plane_stride ) * ( c + d.halo ) + ib;
    const double f          = -d.mu_s[ib] * inv_P;
    for( int jh = jhigh; 4 * jh < m; jh += jstep )
    {
        const int j = 4 * jh + jlow;
        if( 2 * j < d.Na )
        {
            const double2 z = smem[( j << lg_nrow ) + rl];
            g.base[elem_offset( site0 + std::size_t( d.NB ) * ( 2 * j ) ) + comp * FIELD_BLOCK] = f * z.x;
            if( 2 * j + 1 < d.Na )
                g.base[elem_offset( site0 + std::size_t( d.NB ) * ( 2 * j + 1 ) ) + comp * FIELD_BLOCK] = f * z.y;
        }
    }
}

// Persistent, software-pipelined forms of the two a-pass kernels above (one sublattice, even Na): tiles (row group, component)
// blockIdx.x, blockIdx.x + gridDim.x, ...; the rows of the NEXT tile are fetched with 16-byte asynchronous copies under the
// transform of the current one (see k_fft_pass16p). Forward: a pair (x_2j, x_2j+1) of the spin row IS the complex element z_j
// (adjacent doubles of an AoSoA block), copied straight into the transform buffer, mu_s applied to the output. Inverse: the raw
// half-spectrum rows land in a two-deep ring and the Hermitian merge step fills the transform buffer from there.
template<int LOGM>
static __global__ void __launch_bounds__( FFT_THREADS, SB_FFT16_MINB ) k_ddi_fwd_a16p(
    const __grid_constant__ FFTPlan1D plan_h, const double2 * __restrict__ tw_full, const __grid_constant__ DDIDims d,
    ConstField3 spins, double2 * __restrict__ A, const int lg_nrow, const int q0, const int n_rg, const int n_tiles,
    const __grid_constant__ APush ap )
{
    extern __shared__ double2 smem[];
    constexpr int m = 1 << LOGM;
    const int bufp  = ( m << lg_nrow ) + ( ( m << lg_nrow ) >> 4 ) + 1;
    int rl, jlow, jhigh, jstep;
    row_map( lg_nrow, rl, jlow, jhigh, jstep );
    const int rows = d.Nb * d.Nc;
    auto fetch = [&]( int tile, double2 * buf )
    {
        const int qi = tile / n_rg, row = ( ( tile - qi * n_rg ) << lg_nrow ) + rl;
        const bool valid = row < rows;
        const int b = row % d.Nb, c = row / d.Nb;
        const std::size_t site0 = std::size_t( d.Na ) * b + std::size_t( d.plane_stride ) * ( c + d.halo );
        const double * comp     = spins.base + ( ( q0 + qi ) % 3 ) * FIELD_BLOCK;
#pragma unroll
        for( int k = 0; k < FFT_E; ++k )
        {
            const int j        = 4 * ( jhigh + k * jstep ) + jlow;
            const bool nonzero = valid && 2 * j < d.Na;
            cp_async16( buf + ( j << lg_nrow ) + rl, nonzero ? comp + elem_offset( site0 + 2 * j ) : spins.base, nonzero );
        }
        cp_async_commit();
    };
    int tile = blockIdx.x, stage = 0;
    if( tile < n_tiles )
        fetch( tile, smem );
    for( ; tile < n_tiles; tile += gridDim.x, stage ^= 1 )
    {
        cp_async_wait_all();
        __syncthreads();
        if( tile + int( gridDim.x ) < n_tiles )
            fetch( tile + gridDim.x, smem + ( stage ^ 1 ) * bufp );
        double2 * x = smem + stage * bufp;
        block_fft_ct<false, LOGM>( plan_h.twiddle, x, lg_nrow, threadIdx.x & ( ( 1 << lg_nrow ) - 1 ), threadIdx.x >> lg_nrow );
        const int qi = tile / n_rg, row = ( ( tile - qi * n_rg ) << lg_nrow ) + rl;
        if( row < rows )
        {
            const int q = q0 + qi, b = row % d.Nb, c = row / d.Nb;
            const double h = 0.5 * d.mu_s[0];
            double2 * out  = A + ( ( std::size_t( q ) * d.Nc + c ) * d.Nb + b ) * d.Ha;
            for( int jh = jhigh; 4 * jh < m; jh += jstep )
            {
                const int k      = 4 * jh + jlow;
                const double2 Zk = x[( k << lg_nrow ) + rl];
                const double2 Zm = x[( ( ( m - k ) & ( m - 1 ) ) << lg_nrow ) + rl]; // Z_m = Z_0
                const double2 E  = make_double2( h * ( Zk.x + Zm.x ), h * ( Zk.y - Zm.y ) );
                const double2 O  = make_double2( h * ( Zk.y + Zm.y ), -h * ( Zk.x - Zm.x ) );
                const double2 X  = cadd( E, cmul( __ldg( tw_full + k ), O ) );
                if( ap.on )
                    *apush_dst( ap, q, c, b, d.Nb, k ) = X;
                else
                    out[k] = X;
            }
            if( jhigh == 0 && jlow == 0 )
            {
                const double2 Z0 = x[rl];
                const double2 X  = make_double2( d.mu_s[0] * ( Z0.x - Z0.y ), 0.0 );
                if( ap.on )
                    *apush_dst( ap, q, c, b, d.Nb, m ) = X;
                else
                    out[m] = X;
            }
        }
    }
}

template<int LOGM>
static __global__ void __launch_bounds__( FFT_THREADS, SB_FFT16_MINB ) k_ddi_inv_a16p(
    const __grid_constant__ FFTPlan1D plan_h, const double2 * __restrict__ tw_full, const __grid_constant__ DDIDims d,
    const double2 * __restrict__ A, Field3 g, const double inv_P, const int lg_nrow, const int q0, const int n_rg, const int n_tiles,
    const __grid_constant__ APush ap )
{
    extern __shared__ double2 smem[];
    constexpr int m = 1 << LOGM;
    const int rawp  = ( m + 1 ) << lg_nrow; // one raw tile: rows of m + 1 elements
    double2 * x     = smem + 2 * rawp;      // the transform buffer behind the two raw tiles
    int rl, jlow, jhigh, jstep;
    row_map( lg_nrow, rl, jlow, jhigh, jstep );
    const int rows = d.Nb * d.Nc;
    auto fetch = [&]( int tile, double2 * raw )
    {
        const int qi = tile / n_rg, row = ( ( tile - qi * n_rg ) << lg_nrow ) + rl;
        const bool valid = row < rows;
        const int b = row % d.Nb, c = row / d.Nb;
        const double2 * in = valid ? A + ( ( std::size_t( q0 + qi ) * d.Nc + c ) * d.Nb + b ) * d.Ha : A;
#pragma unroll
        for( int kk = 0; kk < FFT_E; ++kk )
        {
            const int k = 4 * ( jhigh + kk * jstep ) + jlow;
            const double2 * src = !valid ? A : ( ap.on ? apush_dst( ap, q0 + qi, c, b, d.Nb, k ) : in + k );
            cp_async16( raw + ( k << lg_nrow ) + rl, src, valid );
        }
        if( jhigh == 0 && jlow == 0 )
            cp_async16( raw + ( m << lg_nrow ) + rl, !valid ? A : ( ap.on ? apush_dst( ap, q0 + qi, c, b, d.Nb, m ) : in + m ), valid );
        cp_async_commit();
    };
    int tile = blockIdx.x, stage = 0;
    if( tile < n_tiles )
        fetch( tile, smem );
    for( ; tile < n_tiles; tile += gridDim.x, stage ^= 1 )
    {
        cp_async_wait_all();
        __syncthreads();
        if( tile + int( gridDim.x ) < n_tiles )
            fetch( tile + gridDim.x, smem + ( stage ^ 1 ) * rawp );
        const double2 * raw = smem + stage * rawp;
#pragma unroll
        for( int kk = 0; kk < FFT_E; ++kk )
        {
            const int k      = 4 * ( jhigh + kk * jstep ) + jlow;
            const double2 Xk = raw[( k << lg_nrow ) + rl], Xm = raw[( ( m - k ) << lg_nrow ) + rl];
            const double2 S  = make_double2( Xk.x + Xm.x, Xk.y - Xm.y ); // X_k + conj X_{m-k}
            const double2 D  = make_double2( Xk.x - Xm.x, Xk.y + Xm.y ); // X_k - conj X_{m-k}
            const double2 w  = __ldg( tw_full + k );                     // exp(-2 pi i k / Pa); the inverse needs its conjugate
            const double2 T  = cmul( make_double2( w.x, -w.y ), D );
            x[( k << lg_nrow ) + rl] = make_double2( S.x - T.y, S.y + T.x ); // S + i T
        }
        __syncthreads();
        block_fft_ct<true, LOGM>( plan_h.twiddle, x, lg_nrow, threadIdx.x & ( ( 1 << lg_nrow ) - 1 ), threadIdx.x >> lg_nrow );
        const int qi = tile / n_rg, row = ( ( tile - qi * n_rg ) << lg_nrow ) + rl;
        if( row < rows )
        {
            const int b = row % d.Nb, c = row / d.Nb;
            const std::size_t site0 = std::size_t( d.Na ) * b + std::size_t( d.plane_stride ) * ( c + d.halo );
            double * comp           = g.base + ( ( q0 + qi ) % 3 ) * FIELD_BLOCK;
            const double f          = -d.mu_s[0] * inv_P;
            for( int jh = jhigh; 4 * jh < m; jh += jstep )
            {
                const int j = 4 * jh + jlow;
                if( 2 * j < d.Na )
                {
                    const double2 z = x[( j << lg_nrow ) + rl];
                    *reinterpret_cast<double2 *>( comp + elem_offset( site0 + 2 * j ) ) = make_double2( f * z.x, f * z.y );
                }
            }
        }
    }
}

// Dipole tensor component `comp6` of sublattice pair (b1, b2) on the padded lattice, with periodic images
// (FFT_Dipole_Matrices, Hamiltonian_Heisenberg.cpp:1406-1499)
struct TensorGeom
{
    double ta[3], tb[3], tc[3]; // lattice_constant * bravais vectors
    double da, db, dc;          // cell_atoms[b1] - cell_atoms[b2] in lattice coordinates
    int img[3];
    double mult;
};
static __global__ void k_ddi_tensor( const __grid_constant__ DDIDims d, const __grid_constant__ TensorGeom t, int comp6, double * __restrict__ out )
{
    const std::size_t i = blockIdx.x * std::size_t( blockDim.x ) + threadIdx.x;
    const std::size_t P = std::size_t( d.Pa ) * d.Pb * d.Pc;
    if( i >= P )
        return;
    const int a = int( i % d.Pa ), b = int( ( i / d.Pa ) % d.Pb ), c = int( i / ( std::size_t( d.Pa ) * d.Pb ) );
    const int ai = a < d.Na ? a : a - d.Pa, bi = b < d.Nb ? b : b - d.Pb, ci = c < d.Nc ? c : c - d.Pc;
    // Offsets of -N along a padded direction never occur between two real sites (|difference| <= N - 1). The reference
    // fills them (with the value for -N, there is no slot for +N), which breaks the inversion symmetry of the padded
    // tensor; zeroed here, D(-r) = D(r) holds on the whole padded lattice and the spectrum of a single sublattice is
    // real. The convolution result at the real sites is identical.
    if( ( d.Pa > d.Na && ai == -d.Na ) || ( d.Pb > d.Nb && bi == -d.Nb ) || ( d.Pc > d.Nc && ci == -d.Nc ) )
    {
        out[i] = 0.0;
        return;
    }
    double D = 0.0;
    for( int pa = -t.img[0]; pa <= t.img[0]; ++pa )
        for( int pb = -t.img[1]; pb <= t.img[1]; ++pb )
            for( int pc = -t.img[2]; pc <= t.img[2]; ++pc )
            {
                const double fa = ai + pa * d.Na + t.da, fb = bi + pb * d.Nb + t.db, fc = ci + pc * d.Nc + t.dc;
                const double rx = fa * t.ta[0] + fb * t.tb[0] + fc * t.tc[0];
                const double ry = fa * t.ta[1] + fb * t.tb[1] + fc * t.tc[1];
                const double rz = fa * t.ta[2] + fb * t.tb[2] + fc * t.tc[2];
                const double r  = sqrt( rx * rx + ry * ry + rz * rz );
                if( r > 1e-10 )
                {
                    const double r3 = r * r * r, r5 = r * r * r * r * r;
                    double v;
                    switch( comp6 )
                    {
                        case 0: v = 3 * rx * rx / r5 - 1 / r3; break;
                        case 1: v = 3 * rx * ry / r5; break;
                        case 2: v = 3 * rx * rz / r5; break;
                        case 3: v = 3 * ry * ry / r5 - 1 / r3; break;
                        case 4: v = 3 * ry * rz / r5; break;
                        default: v = 3 * rz * rz / r5 - 1 / r3; break;
                    }
                    D += t.mult * v;
                }
            }
    out[i] = D;
}

std::vector<int> factorize( int n )
{
    std::vector<int> r;
    while( n % 4 == 0 )
    {
        r.push_back( 4 );
        n /= 4;
    }
    for( int p = 2; p * p <= n; ++p )
        while( n % p == 0 )
        {
            r.push_back( p );
            n /= p;
        }
    if( n > 1 )
        r.push_back( n );
    return r;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// launch shapes of the pipelined a-passes (k_ddi_fwd_a16p / k_ddi_inv_a16p)
struct APipe
{
    int ctas_fwd = 0, ctas_inv = 0;
    std::size_t smem_fwd = 0, smem_inv = 0;
};
struct DDIPlan
{
    DDIDims dims{};   // local planes: a- and b-passes
    DDIDims dims_g{}; // global lattice: tensor setup
    DDIDims dims_c{}; // c-pass operand: all planes, local kb range
    int world = 1, rank = 0;
    double2 * C = nullptr; // distributed: the c-pass operand after the all-to-all
    FFTPlan1D plan[3]; // a, b, c
    double2 * twiddle[3] = { nullptr, nullptr, nullptr };
    double2 * Dhat       = nullptr; // complex tensor spectrum, or
    double * Dhat_real   = nullptr; // its real part when the imaginary part vanishes (single sublattice)
    double2 * A          = nullptr;
    double2 * B          = nullptr;
    int ncol_b = 1, ncol_c = 1;
    std::size_t smem_a = 0, smem_b = 0, smem_c = 0;
    std::uint64_t launches_setup = 0;
    // fast variants (power-of-two lengths, see k_fft_pass16 / k_ddi_c_mult16 / k_ddi_fwd_a16)
    struct Fast
    {
        bool on = false;
        int lg = 0, threads = 0; // lg(columns or rows per CTA), CTA size
        int lg_seq = 0;          // b-pass: lg(column groups a CTA transforms one after the other)
        int pipe_ctas = 0;       // > 0: the persistent pipelined kernel (k_fft_pass16p) with this many CTAs (two tile buffers in smem)
        std::size_t smem = 0;
    } fast_a, fast_b, fast_c;
    APipe apipe;
    bool a_pipe = false;           // the a-passes run as the persistent pipelined kernels
    // pencil decomposition over ranks (ka cut into per-rank blocks): see ddi_gradient_pencil
    bool pencil = false, pencil_overlap = false;
    // pencil transposes by the copy engines (ddi_gradient_pencil_dma): the a-passes stay local, 2-D peer copies carry the ka blocks
    bool pencil_dma = false;
    double2 * A_recv = nullptr;    // [q][ncl][Nb][Ha]: the inverse a-pass operand, filled by the ranks' copies
    std::vector<void *> A_recv_peer;
    cudaStream_t copy_stream[DDI_MAX_PEERS] = {};
    cudaEvent_t ev_dma[4 * 3] = {};
    APush push{};                  // destinations of the forward a-pass / sources of the inverse a-pass (on = 0 without pencils)
    double2 * AT = nullptr;        // [q][Nc_global][Nb][w]: the b-pass operand of this rank's ka block, written by all ranks
    double2 * BT = nullptr;        // [q][Nc_global][Pb][w]: the c-pass operand
    std::vector<void *> AT_peer;
    FFTPlan1D plan_ah;             // length Pa / 2
    double2 * twiddle_ah = nullptr;
    int lg_split         = 31;     // lg of the per-rank kb block (distributed layout), 31: no split
    void * Dt            = nullptr; // tensor spectrum in the tile order of k_ddi_c_mult16
    bool Dt_real         = false;
    // distributed over peer-mapped memory (NVLink): no all-to-all at all. The forward b-pass stores its per-rank kb blocks
    // straight into the c-pass operand of their owners, the inverse b-pass loads its input from there; the c-pass operand
    // is double-buffered (evaluation e uses C2[e & 1]) and the ranks meet at two counters per evaluation.
    bool peer               = false;
    double2 * C2[2]         = { nullptr, nullptr };
    std::vector<void *> C2_peer[2];  // [buffer][rank]
    unsigned * flags        = nullptr; // [world]: flags[r] = barriers rank r has arrived at
    std::vector<void *> flags_peer;  // [rank]: the flags array of every rank
    std::vector<void *> peer_opened;
    unsigned barrier_count  = 0;
    unsigned evaluation     = 0;
    // distributed: the all-to-alls run on their own stream, one component at a time, under the passes of the others
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_q[3 * MAX_BASIS] = {};
    cudaEvent_t ev_all = nullptr;

    ~DDIPlan()
    {
        for( auto * t : twiddle )
            if( t )
                cudaFree( t );
        if( Dhat )
            cudaFree( Dhat );
        if( Dhat_real )
            cudaFree( Dhat_real );
        if( Dt )
            cudaFree( Dt );
        if( comm_stream )
            cudaStreamDestroy( comm_stream );
        for( auto e : ev_q )
            if( e )
                cudaEventDestroy( e );
        if( ev_all )
            cudaEventDestroy( ev_all );
        if( twiddle_ah )
            cudaFree( twiddle_ah );
        if( A )
            cudaFree( A );
        if( A_recv )
            cudaFree( A_recv );
        for( auto st : copy_stream )
            if( st )
                cudaStreamDestroy( st );
        for( auto e : ev_dma )
            if( e )
                cudaEventDestroy( e );
        if( AT )
            cudaFree( AT );
        if( BT )
            cudaFree( BT );
        if( B )
            cudaFree( B );
        if( C )
            cudaFree( C );
        for( void * q : peer_opened )
            cudaIpcCloseMemHandle( q );
        for( double2 * c : C2 )
            if( c )
                cudaFree( c );
        if( flags )
            cudaFree( flags );
    }
};

namespace
{
bool env_flag_off( const char * name )
{
    const char * v = std::getenv( name );
    return v && v[0] == '0';
}

void make_plan_1d( FFTPlan1D & plan, double2 *& table, int n )
{
    plan.n         = n;
    plan.fast16    = ( ( n & ( n - 1 ) ) == 0 && n >= 64 && !env_flag_off( "SPIRIT_B200_FFT16" ) ) ? 1 : 0;
    auto fac       = factorize( n );
    if( plan.fast16 )
    {
        fac.clear();
        int rest = n;
        while( rest % FFT_E == 0 )
        {
            fac.push_back( FFT_E );
            rest /= FFT_E;
        }
        if( rest > 1 )
            fac.push_back( rest );
    }
    if( fac.size() > std::size_t( MAX_RADICES ) )
        throw std::runtime_error( "spirit_b200: FFT length with too many factors" );
    plan.n_radix = int( fac.size() );
    plan.pow2    = ( n & ( n - 1 ) ) == 0 ? 1 : 0;
    for( int i = 0; i < plan.n_radix; ++i )
        plan.radix[i] = fac[i];
    std::vector<double2> w( n );
    for( int k = 0; k < n; ++k )
    {
        const long double ang = -2.0L * 3.14159265358979323846264338327950288L * k / n;
        w[k]                  = make_double2( double( cosl( ang ) ), double( sinl( ang ) ) );
    }
    SB_CUDA_CHECK( cudaMalloc( &table, std::size_t( n ) * sizeof( double2 ) ) );
    SB_CUDA_CHECK( cudaMemcpy( table, w.data(), std::size_t( n ) * sizeof( double2 ), cudaMemcpyHostToDevice ) );
    plan.twiddle = table;
}

// CTA size for a pass with `items` radix-4 work items per stage
int fft_threads( int items )
{
    return std::max( 64, std::min( FFT_THREADS, ( ( items + 31 ) / 32 ) * 32 ) );
}

int choose_ncol( int n, int n_components, int n_u )
{
    const std::size_t per_col = std::size_t( n ) * 2 * sizeof( double2 ) * n_components;
    int ncol                  = int( std::min<std::size_t>( 16, std::max<std::size_t>( 1, MAX_SMEM_FFT / per_col ) ) );
    ncol                      = std::min( ncol, std::max( 1, n_u ) );
    while( ncol & ( ncol - 1 ) ) // power of two: shift-based index arithmetic in block_fft
        ncol &= ncol - 1;
    if( per_col > std::size_t( MAX_SMEM_FFT ) )
        throw std::runtime_error( "spirit_b200: padded lattice dimension too long for the shared-memory FFT passes" );
    return ncol;
}

template<typename K>
void allow_smem( K kernel, std::size_t bytes )
{
    SB_CUDA_CHECK( cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int( bytes ) ) );
}
// Host-side selection of the per-length instantiations of the fast kernels (n = 64 .. 4096)
int ilog2( int n )
{
    return 31 - __builtin_clz( unsigned( n ) );
}
#define SB_FOR_LOGN( C ) C( 6 ) C( 7 ) C( 8 ) C( 9 ) C( 10 ) C( 11 ) C( 12 )
template<bool INVERSE>
void launch_pass16(
    const DDIPlan::Fast & f, dim3 grid, cudaStream_t stream, const FFTPlan1D & plan, const PassArgs & a, int lg_in, int lg_out, bool configure = false )
{
    if( f.pipe_ctas && !configure )
    {
        const int n_tiles_u = ( a.n_u + ( 1 << f.lg ) - 1 ) >> f.lg, n_tiles = n_tiles_u * a.n_o;
        const dim3 pgrid( std::min( n_tiles, f.pipe_ctas ) );
        switch( ilog2( plan.n ) )
        {
#define C( L )                                                                                                         \
    case L: k_fft_pass16p<INVERSE, L><<<pgrid, f.threads, f.smem, stream>>>( plan, a, f.lg, lg_in, lg_out, n_tiles_u, n_tiles ); break;
            SB_FOR_LOGN( C )
#undef C
            default: throw std::logic_error( "spirit_b200: no fast pass kernel for this length" );
        }
        return;
    }
    // register-resident first / last stages (k_fft_pass16r) unless SPIRIT_B200_FFT_PASS_REG=0
    const bool reg = FFT_E == 8 && !env_flag_off( "SPIRIT_B200_FFT_PASS_REG" );
    switch( ilog2( plan.n ) )
    {
#define C( L )                                                                                                         \
    case L:                                                                                                            \
        if( configure && f.lg_seq )                                                                                    \
        {                                                                                                              \
            allow_smem( k_fft_pass16<INVERSE, L, 1>, f.smem );                                                         \
            allow_smem( k_fft_pass16r<INVERSE, L, 1>, f.smem );                                                        \
        }                                                                                                              \
        else if( configure )                                                                                           \
        {                                                                                                              \
            allow_smem( k_fft_pass16<INVERSE, L, 0>, f.smem );                                                         \
            allow_smem( k_fft_pass16r<INVERSE, L, 0>, f.smem );                                                        \
        }                                                                                                              \
        else if( f.lg_seq && reg )                                                                                     \
            k_fft_pass16r<INVERSE, L, 1><<<grid, f.threads, f.smem, stream>>>( plan, a, f.lg, lg_in, lg_out );         \
        else if( f.lg_seq )                                                                                            \
            k_fft_pass16<INVERSE, L, 1><<<grid, f.threads, f.smem, stream>>>( plan, a, f.lg, lg_in, lg_out );          \
        else if( reg )                                                                                                 \
            k_fft_pass16r<INVERSE, L, 0><<<grid, f.threads, f.smem, stream>>>( plan, a, f.lg, lg_in, lg_out );         \
        else                                                                                                           \
            k_fft_pass16<INVERSE, L, 0><<<grid, f.threads, f.smem, stream>>>( plan, a, f.lg, lg_in, lg_out );          \
        break;
        SB_FOR_LOGN( C )
#undef C
        default: throw std::logic_error( "spirit_b200: no fast pass kernel for this length" );
    }
}
// The pipelined pass: two tile buffers; persistent grid = CTAs that fit on the device at once. Returns false (and leaves f
// unchanged) when two buffers do not fit into shared memory.
bool configure_pass16_pipelined( DDIPlan::Fast & f, int n )
{
    const std::size_t smem = 2 * ( ( std::size_t( n ) << f.lg ) * 17 / 16 + 1 ) * sizeof( double2 );
    if( smem > std::size_t( 220 * 1024 ) )
        return false;
    int dev = 0, sms = 0, per_sm_f = 0, per_sm_i = 0;
    SB_CUDA_CHECK( cudaGetDevice( &dev ) );
    SB_CUDA_CHECK( cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, dev ) );
    switch( ilog2( n ) )
    {
#define C( L )                                                                                                         \
    case L:                                                                                                            \
        allow_smem( k_fft_pass16p<false, L>, smem );                                                                   \
        allow_smem( k_fft_pass16p<true, L>, smem );                                                                    \
        SB_CUDA_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm_f, k_fft_pass16p<false, L>, f.threads, smem ) ); \
        SB_CUDA_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm_i, k_fft_pass16p<true, L>, f.threads, smem ) );  \
        break;
        SB_FOR_LOGN( C )
#undef C
        default: return false;
    }
    const int per_sm = std::min( per_sm_f, per_sm_i );
    if( per_sm < 1 )
        return false;
    f.smem      = smem;
    f.lg_seq    = 0;
    f.pipe_ctas = per_sm * sms;
    return true;
}
template<bool REAL_D>
void launch_c_mult16(
    const DDIPlan::Fast & f, dim3 grid, cudaStream_t stream, const FFTPlan1D & plan, const DDIDims & dc, double2 * operand, const void * Dt,
    bool configure = false )
{
    // register-resident first / last stages (k_ddi_c_mult16f) unless SPIRIT_B200_DDI_C_FUSED=0 or the CTA is larger than the
    // kernel was compiled for
    const bool fused = FFT_E == 8 && !env_flag_off( "SPIRIT_B200_DDI_C_FUSED" ) && f.threads <= cf_threads( plan.n );
    switch( ilog2( plan.n ) )
    {
#define C( L )                                                                                                         \
    case L:                                                                                                            \
        if( configure )                                                                                                \
        {                                                                                                              \
            allow_smem( k_ddi_c_mult16<REAL_D, L>, f.smem );                                                           \
            allow_smem( k_ddi_c_mult16f<REAL_D, L>, f.smem );                                                          \
        }                                                                                                              \
        else if( fused )                                                                                               \
            k_ddi_c_mult16f<REAL_D, L><<<grid, f.threads, f.smem, stream>>>( plan, dc, operand, Dt, f.lg );            \
        else                                                                                                           \
            k_ddi_c_mult16<REAL_D, L><<<grid, f.threads, f.smem, stream>>>( plan, dc, operand, Dt, f.lg );             \
        break;
        SB_FOR_LOGN( C )
#undef C
        default: throw std::logic_error( "spirit_b200: no fast c-pass kernel for this length" );
    }
}
// the pipelined a-passes: smem_inv_p bytes for the inverse (two raw tiles + transform buffer), 2 x f.smem for the forward
bool configure_a16_pipelined( const DDIPlan::Fast & f, int m, APipe & ap )
{
    ap.smem_fwd = 2 * f.smem;
    ap.smem_inv = 2 * ( ( std::size_t( m ) + 1 ) << f.lg ) * sizeof( double2 ) + f.smem;
    if( ap.smem_fwd > std::size_t( 220 * 1024 ) || ap.smem_inv > std::size_t( 220 * 1024 ) )
        return false;
    int dev = 0, sms = 0, pf = 0, pi = 0;
    SB_CUDA_CHECK( cudaGetDevice( &dev ) );
    SB_CUDA_CHECK( cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, dev ) );
    switch( ilog2( m ) )
    {
#define C( L )                                                                                                         \
    case L:                                                                                                            \
        allow_smem( k_ddi_fwd_a16p<L>, ap.smem_fwd );                                                                  \
        allow_smem( k_ddi_inv_a16p<L>, ap.smem_inv );                                                                  \
        SB_CUDA_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &pf, k_ddi_fwd_a16p<L>, f.threads, ap.smem_fwd ) ); \
        SB_CUDA_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &pi, k_ddi_inv_a16p<L>, f.threads, ap.smem_inv ) ); \
        break;
        SB_FOR_LOGN( C )
#undef C
        default: return false;
    }
    if( pf < 1 || pi < 1 )
        return false;
    ap.ctas_fwd = pf * sms;
    ap.ctas_inv = pi * sms;
    return true;
}
void launch_fwd_a16p(
    const DDIPlan::Fast & f, const APipe & ap, int n_rg, int nq, cudaStream_t stream, const FFTPlan1D & plan_h, const double2 * tw_full,
    const DDIDims & d, ConstField3 spins, double2 * A, int q0, const APush & push )
{
    const int n_tiles = n_rg * nq;
    const dim3 grid( std::min( n_tiles, ap.ctas_fwd ) );
    switch( ilog2( plan_h.n ) )
    {
#define C( L )                                                                                                         \
    case L: k_ddi_fwd_a16p<L><<<grid, f.threads, ap.smem_fwd, stream>>>( plan_h, tw_full, d, spins, A, f.lg, q0, n_rg, n_tiles, push ); break;
        SB_FOR_LOGN( C )
#undef C
        default: throw std::logic_error( "spirit_b200: no fast a-pass kernel for this length" );
    }
}
void launch_inv_a16p(
    const DDIPlan::Fast & f, const APipe & ap, int n_rg, int nq, cudaStream_t stream, const FFTPlan1D & plan_h, const double2 * tw_full,
    const DDIDims & d, const double2 * A, Field3 g, double inv_P, int q0, const APush & push )
{
    const int n_tiles = n_rg * nq;
    const dim3 grid( std::min( n_tiles, ap.ctas_inv ) );
    switch( ilog2( plan_h.n ) )
    {
#define C( L )                                                                                                         \
    case L: k_ddi_inv_a16p<L><<<grid, f.threads, ap.smem_inv, stream>>>( plan_h, tw_full, d, A, g, inv_P, f.lg, q0, n_rg, n_tiles, push ); break;
        SB_FOR_LOGN( C )
#undef C
        default: throw std::logic_error( "spirit_b200: no fast a-pass kernel for this length" );
    }
}
void launch_fwd_a16(
    const DDIPlan::Fast & f, dim3 grid, cudaStream_t stream, const FFTPlan1D & plan_h, const double2 * tw_full, const DDIDims & d,
    ConstField3 spins, double2 * A, int q0, bool configure = false, const APush & push = APush{} )
{
    switch( ilog2( plan_h.n ) )
    {
#define C( L )                                                                                                         \
    case L:                                                                                                            \
        if( configure )                                                                                                \
            allow_smem( k_ddi_fwd_a16<L>, f.smem );                                                                    \
        else                                                                                                           \
            k_ddi_fwd_a16<L><<<grid, f.threads, f.smem, stream>>>( plan_h, tw_full, d, spins, A, f.lg, q0, push );     \
        break;
        SB_FOR_LOGN( C )
#undef C
        default: throw std::logic_error( "spirit_b200: no fast a-pass kernel for this length" );
    }
}
void launch_inv_a16(
    const DDIPlan::Fast & f, dim3 grid, cudaStream_t stream, const FFTPlan1D & plan_h, const double2 * tw_full, const DDIDims & d,
    const double2 * A, Field3 g, double inv_P, int q0, bool configure = false, const APush & push = APush{} )
{
    switch( ilog2( plan_h.n ) )
    {
#define C( L )                                                                                                         \
    case L:                                                                                                            \
        if( configure )                                                                                                \
            allow_smem( k_ddi_inv_a16<L>, f.smem );                                                                    \
        else                                                                                                           \
            k_ddi_inv_a16<L><<<grid, f.threads, f.smem, stream>>>( plan_h, tw_full, d, A, g, inv_P, f.lg, q0, push );  \
        break;
        SB_FOR_LOGN( C )
#undef C
        default: throw std::logic_error( "spirit_b200: no fast a-pass kernel for this length" );
    }
}
} // namespace

void ddi_plan_destroy( DDIPlan * p )
{
    delete p;
}

// Prepare_DDI (Hamiltonian_Heisenberg.cpp:1501-1595): padded sizes, plans, tensor spectrum
DDIPlan * ddi_plan_create( const Hamiltonian & ham, const StencilParams & sp, cudaStream_t stream )
{
    const Geometry & g = *ham.geometry;
    auto plan          = std::unique_ptr<DDIPlan>( new DDIPlan );
    // Slab decomposition (one process per GPU): this rank holds ncl = n_cells[2] planes of a lattice with Nc_global
    // planes. The a- and b-passes run on the local planes; an all-to-all turns "my planes, all kb" into "all planes, my kb
    // range" for the c-pass + tensor multiply, and back. The tensor spectrum is stored for the local kb range only.
    const bool slab     = sp.halo > 0;
    const int world     = slab ? comm_world() : 1;
    const int rank      = slab ? comm_rank() : 0;
    const int ncl       = g.n_cells[2];
    const int Nc_global = slab ? sp.Nc : ncl;
    plan->world = world, plan->rank = rank;
    if( slab && ( Nc_global != ncl * world || sp.c_begin != rank * ncl ) )
        throw std::runtime_error( "spirit_b200: the distributed dipole convolution needs equal slabs in rank order" );

    DDIDims & d = plan->dims; // local passes (a, b)
    d.Na = g.n_cells[0], d.Nb = g.n_cells[1], d.Nc = ncl, d.NB = g.n_cell_atoms;
    const int N[3] = { g.n_cells[0], g.n_cells[1], Nc_global };
    int P[3];
    for( int i = 0; i < 3; ++i )
    {
        P[i] = N[i];
        if( N[i] > 1 && ( ham.boundary_conditions[i] == 0 || ham.ddi_pb_zero_padding ) )
            P[i] *= 2;
    }
    d.Pa = P[0], d.Pb = P[1], d.Pc = P[2];
    d.Ha           = d.Pa / 2 + 1;
    d.plane_stride = sp.plane_stride;
    d.halo         = sp.halo;
    for( int ib = 0; ib < d.NB; ++ib )
        d.mu_s[ib] = g.cell_mu_s[ib];
    // inter-sublattice lookup (same enumeration as the reference)
    d.n_inter = 0;
    for( int b1 = 0; b1 < d.NB; ++b1 )
        for( int b2 = 0; b2 < d.NB; ++b2 )
        {
            if( b1 == b2 && b1 != 0 )
            {
                d.lookup[b1 + b2 * d.NB] = 0;
                continue;
            }
            d.lookup[b1 + b2 * d.NB] = d.n_inter++;
        }
    if( d.Pb % world != 0 )
        throw std::runtime_error( "spirit_b200: the padded b dimension must be divisible by the number of ranks" );
    int kbl      = d.Pb / world; // kb range of this rank's c-pass (the whole axis on one device and with ka pencils)
    const int nq = 3 * d.NB;
    int ka0 = 0, Ha_loc = 0;     // ka pencils: this rank's block of the half spectrum
    d.c_block      = d.Nc;
    d.q_stride     = std::size_t( d.Nc ) * d.Pb * d.Ha;
    d.block_stride = 0;

    DDIDims & dg = plan->dims_g; // global lattice: the tensor
    dg           = d;
    dg.Nc        = Nc_global;
    DDIDims & dc = plan->dims_c; // c-pass operand: all planes, local kb range
    dc           = dg;
    dc.Pb        = kbl;
    if( world > 1 )
    {
        dc.c_block      = ncl;
        dc.q_stride     = std::size_t( ncl ) * kbl * d.Ha;
        dc.block_stride = std::size_t( nq ) * ncl * kbl * d.Ha;
    }
    else
    {
        dc.c_block  = Nc_global;
        dc.q_stride = std::size_t( Nc_global ) * kbl * d.Ha;
    }

    make_plan_1d( plan->plan[0], plan->twiddle[0], d.Pa );
    make_plan_1d( plan->plan[1], plan->twiddle[1], d.Pb );
    make_plan_1d( plan->plan[2], plan->twiddle[2], d.Pc );
    plan->ncol_b = choose_ncol( d.Pb, 1, d.Ha );
    // (the generic c-pass holds 3 NB components x 2 buffers of a column: lengths it cannot hold are served by the fast kernel only)
    const bool generic_c_fits = std::size_t( d.Pc ) * 2 * sizeof( double2 ) * nq <= std::size_t( MAX_SMEM_FFT );
    plan->ncol_c              = generic_c_fits ? choose_ncol( d.Pc, nq, d.Ha ) : 1;
    plan->smem_a = std::size_t( d.Pa ) * 2 * sizeof( double2 );
    plan->smem_b = std::size_t( d.Pb ) * 2 * sizeof( double2 ) * plan->ncol_b;
    // + one sixteenth of a buffer: the padded in-place stages of block_fft16 run past the LAST component's second buffer
    plan->smem_c = std::size_t( d.Pc ) * 2 * sizeof( double2 ) * plan->ncol_c * nq
                   + ( std::size_t( d.Pc ) * plan->ncol_c / 16 + 1 ) * sizeof( double2 );
    if( plan->smem_a > std::size_t( MAX_SMEM_FFT ) )
        throw std::runtime_error( "spirit_b200: padded lattice dimension a too long for the shared-memory FFT passes" );
    allow_smem( k_ddi_fwd_a, plan->smem_a );
    allow_smem( k_ddi_inv_a, plan->smem_a );
    allow_smem( k_fft_real_rows, plan->smem_a );
    allow_smem( k_fft_pass<false>, plan->smem_b ); // (the c-pass of the tensor setup raises it to its own need below)
    allow_smem( k_fft_pass<true>, plan->smem_b );
    if( generic_c_fits )
        allow_smem( k_ddi_c_mult, plan->smem_c );
    {
        // (n / 8) << lg threads, about 256 per CTA; one in-place buffer of n << lg elements (+ 1/16 padding) per transform set
        const bool allow = !env_flag_off( "SPIRIT_B200_FFT_FAST" );
        // `elements`: transform-set size a CTA aims at (n << lg <= elements). The passes are bound by barriers and fixed
        // latencies inside a CTA, not by bytes in flight: many small CTAs per SM (independent barrier domains) beat few large
        // ones (profiles/r2u_sweep.txt, r2v_sweep.txt: 256^3 12.2 -> 11.1 ms, 512^3 112 -> 107 ms per SIB iteration)
        auto shape       = []( DDIPlan::Fast & f, int n, int n_buffers, const char * tune, int elements = 2048 )
        {
            f.lg = 0;
            while( f.lg < 4 && ( n << ( f.lg + 1 ) ) <= elements )
                ++f.lg;
            while( ( ( n >> FFT_LG_E ) << f.lg ) < 32 ) // at least one warp
                ++f.lg;
            if( const char * v = std::getenv( tune ) ) // tuning runs (profiles/): columns per CTA
                f.lg = std::max( 0, std::min( 5, std::atoi( v ) ) );
            f.threads = ( n >> FFT_LG_E ) << f.lg;
            f.smem    = std::size_t( n_buffers ) * ( ( std::size_t( n ) << f.lg ) * 17 / 16 + 1 ) * sizeof( double2 );
            f.on      = f.threads >= 32 && f.threads <= FFT_THREADS && f.smem <= std::size_t( 220 * 1024 );
        };
        const bool pow2_world = ( world & ( world - 1 ) ) == 0;
        if( allow && plan->plan[1].fast16 && pow2_world )
        {
            // up to length 512: 128 threads (1024 elements) and one column group; longer: two column groups per CTA, loaded /
            // stored together and transformed in turn (rows of the tile twice as long: a whole sector where a CTA holds one
            // column, lengths >= 2048). Measured: profiles/r1y, r2v
            shape( plan->fast_b, d.Pb, 1, "SPIRIT_B200_FFT_LG_B", d.Pb <= 512 ? 1024 : 2048 );
            plan->fast_b.lg_seq = d.Pb <= 512 ? 0 : 1;
            if( const char * v = std::getenv( "SPIRIT_B200_FFT_SEQ_B" ) )
                plan->fast_b.lg_seq = std::atoi( v ) ? 1 : 0;
            if( plan->fast_b.lg_seq )
            {
                plan->fast_b.smem *= 2;
                if( plan->fast_b.smem > std::size_t( 220 * 1024 ) )
                    plan->fast_b.lg_seq = 0, plan->fast_b.smem /= 2;
            }
            if( world > 1 )
                plan->lg_split = 31 - __builtin_clz( unsigned( kbl ) );
        }
        if( allow && plan->plan[2].fast16 && d.NB == 1 )
            shape( plan->fast_c, d.Pc, 3, "SPIRIT_B200_FFT_LG_C", // the register-resident kernel: CTAs of cf_threads( Pc )
                   FFT_E == 8 && !env_flag_off( "SPIRIT_B200_DDI_C_FUSED" ) ? 8 * cf_threads( d.Pc ) : 2048 );
        if( allow && d.Pa % 2 == 0 && d.Pa >= 128 && ( d.Pa & ( d.Pa - 1 ) ) == 0 && !env_flag_off( "SPIRIT_B200_FFT16" ) )
        {
            make_plan_1d( plan->plan_ah, plan->twiddle_ah, d.Pa / 2 );
            if( plan->plan_ah.fast16 )
                shape( plan->fast_a, d.Pa / 2, 1, "SPIRIT_B200_FFT_LG_A", 256 ); // one or two warps per CTA
        }
        // lengths up to 4096 have instantiations
        plan->fast_b.on = plan->fast_b.on && d.Pb <= 4096;
        plan->fast_c.on = plan->fast_c.on && d.Pc <= 4096;
        plan->fast_a.on = plan->fast_a.on && d.Pa / 2 <= 4096;
        if( plan->fast_b.on && std::getenv( "SPIRIT_B200_FFT_PIPE_B" ) && !env_flag_off( "SPIRIT_B200_FFT_PIPE_B" ) )
        {
            // persistent pipelined b-pass; its own tile width (columns per CTA) may differ from the plain kernel's
            DDIPlan::Fast f = plan->fast_b;
            if( const char * v = std::getenv( "SPIRIT_B200_FFT_PIPE_LG_B" ) )
            {
                f.lg      = std::max( 0, std::min( 5, std::atoi( v ) ) );
                f.threads = ( d.Pb >> FFT_LG_E ) << f.lg;
            }
            if( f.threads >= 32 && f.threads <= FFT_THREADS && configure_pass16_pipelined( f, d.Pb ) )
                plan->fast_b = f;
        }
        if( plan->fast_b.on && !plan->fast_b.pipe_ctas )
        {
            launch_pass16<false>( plan->fast_b, dim3(), stream, plan->plan[1], PassArgs{}, 0, 0, true );
            launch_pass16<true>( plan->fast_b, dim3(), stream, plan->plan[1], PassArgs{}, 0, 0, true );
        }
        if( plan->fast_c.on )
        {
            launch_c_mult16<false>( plan->fast_c, dim3(), stream, plan->plan[2], plan->dims_c, nullptr, nullptr, true );
            launch_c_mult16<true>( plan->fast_c, dim3(), stream, plan->plan[2], plan->dims_c, nullptr, nullptr, true );
        }
        // pipelined (cp.async double-buffered, persistent) forms, where they measured faster than the plain kernels: the a-passes
        // of long rows (film, half length 2048: 0.54 / 0.65 -> 0.51 / 0.55 ms, profiles/r2t); short rows and the b-passes are
        // served as well or better by small plain CTAs (profiles/r2s, r2t, r2u). SPIRIT_B200_FFT_PIPE_A / _B = 1 / 0 force them.
        auto env_choice = []( const char * name, bool otherwise )
        {
            const char * v = std::getenv( name );
            return v ? v[0] != '0' : otherwise;
        };
        if( plan->fast_a.on && d.NB == 1 && d.Na % 2 == 0 && d.plane_stride % 2 == 0 && env_choice( "SPIRIT_B200_FFT_PIPE_A", d.Pa / 2 >= 1024 ) )
            plan->a_pipe = configure_a16_pipelined( plan->fast_a, d.Pa / 2, plan->apipe );
        if( plan->fast_a.on )
        {
            launch_fwd_a16( plan->fast_a, dim3(), stream, plan->plan_ah, nullptr, d, ConstField3{}, nullptr, 0, true );
            launch_inv_a16( plan->fast_a, dim3(), stream, plan->plan_ah, nullptr, d, nullptr, Field3{}, 0.0, 0, true );
        }
        // Several ranks, fast kernels, peer-mapped memory: ka PENCILS (ddi_gradient_pencil). The transposes move the a-pass output
        // (b not yet padded: half the volume of the kb-block transposes below) and are the stores / loads of the a-pass kernels.
        if( world > 1 && world <= DDI_MAX_PEERS && pow2_world && plan->fast_a.on && plan->fast_b.on && d.NB == 1 && ( d.Pa / 2 ) % world == 0
            && !env_flag_off( "SPIRIT_B200_DDI_PENCIL" ) && peer_memops_available() )
        {
            plan->pencil   = true;
            const int kblk = d.Pa / 2 / world;
            ka0            = rank * kblk;
            Ha_loc         = kblk + ( rank == world - 1 ? 1 : 0 );
            kbl            = d.Pb;
            dc             = dg;
            dc.Ha          = Ha_loc;
            dc.c_block     = Nc_global;
            dc.q_stride    = std::size_t( Nc_global ) * d.Pb * Ha_loc;
            dc.block_stride = 0;
            plan->lg_split = 31;
            plan->push.on        = 1;
            plan->push.lg_kblock = ilog2( kblk );
            plan->push.world     = world;
            plan->push.planes    = Nc_global;
            plan->push.c_begin   = rank * ncl;
            for( int r = 0; r < world; ++r )
                plan->push.w[r] = kblk + ( r == world - 1 ? 1 : 0 );
            // Overlapped schedule (ddi_gradient_pencil): the a-passes, whose speed is set by NVLink, run as PERSISTENT kernels on a
            // bounded number of CTAs per SM, so that the b- and c-passes of the other components (own stream) share the SMs with them
            if( const char * v = std::getenv( "SPIRIT_B200_DDI_PENCIL_DMA" ) )
                plan->pencil_dma = v[0] != '0';
            if( plan->pencil_dma )
            {
                plan->push.on = 0; // the a-passes read / write this rank's own staging buffers
                int least = 0, greatest = 0;
                SB_CUDA_CHECK( cudaDeviceGetStreamPriorityRange( &least, &greatest ) );
                SB_CUDA_CHECK( cudaStreamCreateWithPriority( &plan->comm_stream, cudaStreamNonBlocking, greatest ) );
                for( int r = 0; r < world; ++r )
                    SB_CUDA_CHECK( cudaStreamCreateWithPriority( &plan->copy_stream[r], cudaStreamNonBlocking, greatest ) );
                for( auto & e : plan->ev_dma )
                    SB_CUDA_CHECK( cudaEventCreateWithFlags( &e, cudaEventDisableTiming ) );
            }
            plan->pencil_overlap = !plan->pencil_dma && !env_flag_off( "SPIRIT_B200_DDI_PENCIL_OVERLAP" ) && d.Na % 2 == 0 && d.plane_stride % 2 == 0;
            if( plan->pencil_overlap )
            {
                plan->a_pipe = configure_a16_pipelined( plan->fast_a, d.Pa / 2, plan->apipe );
                plan->pencil_overlap = plan->a_pipe;
            }
            if( plan->pencil_overlap )
            {
                int cap = 4; // CTAs per SM of the persistent a-passes
                if( const char * v = std::getenv( "SPIRIT_B200_DDI_PENCIL_CTAS" ) )
                    cap = std::max( 1, std::atoi( v ) );
                int dev = 0, sms = 0;
                SB_CUDA_CHECK( cudaGetDevice( &dev ) );
                SB_CUDA_CHECK( cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, dev ) );
                plan->apipe.ctas_fwd = std::min( plan->apipe.ctas_fwd, cap * sms );
                plan->apipe.ctas_inv = std::min( plan->apipe.ctas_inv, cap * sms );
                int least = 0, greatest = 0;
                SB_CUDA_CHECK( cudaDeviceGetStreamPriorityRange( &least, &greatest ) );
                SB_CUDA_CHECK( cudaStreamCreateWithPriority( &plan->comm_stream, cudaStreamNonBlocking, greatest ) );
                for( int q = 0; q < 2 * nq; ++q )
                    SB_CUDA_CHECK( cudaEventCreateWithFlags( &plan->ev_q[q], cudaEventDisableTiming ) );
                SB_CUDA_CHECK( cudaEventCreateWithFlags( &plan->ev_all, cudaEventDisableTiming ) );
            }
        }
        if( world > 1 && !plan->pencil && plan->fast_a.on && plan->fast_b.on && !env_flag_off( "SPIRIT_B200_DDI_PIPELINE" ) )
        {
            int least = 0, greatest = 0;
            SB_CUDA_CHECK( cudaDeviceGetStreamPriorityRange( &least, &greatest ) );
            SB_CUDA_CHECK( cudaStreamCreateWithPriority( &plan->comm_stream, cudaStreamNonBlocking, greatest ) );
            for( int q = 0; q < nq; ++q )
                SB_CUDA_CHECK( cudaEventCreateWithFlags( &plan->ev_q[q], cudaEventDisableTiming ) );
            SB_CUDA_CHECK( cudaEventCreateWithFlags( &plan->ev_all, cudaEventDisableTiming ) );
        }
    }

    const std::size_t half       = std::size_t( d.Pc ) * d.Pb * d.Ha; // full half-spectrum of one tensor component
    if( !plan->pencil )
        Ha_loc = d.Ha;
    const std::size_t half_local = std::size_t( d.Pc ) * kbl * Ha_loc; // the local kb range (ka pencils: the local ka block) of it
    const std::size_t full       = std::size_t( d.Pc ) * d.Pb * d.Pa;
    const std::size_t n_B        = std::size_t( nq ) * ncl * d.Pb * d.Ha;
    SB_CUDA_CHECK( cudaMalloc( &plan->Dhat, std::size_t( 6 * d.n_inter ) * half_local * sizeof( double2 ) ) );
    if( plan->pencil )
    {
        // b- and c-pass operands of the local ka block; AT is written (forward a-pass) and read (inverse a-pass) by every rank
        SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
        SB_CUDA_CHECK( cudaMalloc( &plan->AT, std::size_t( nq ) * Nc_global * d.Nb * Ha_loc * sizeof( double2 ) ) );
        SB_CUDA_CHECK( cudaMalloc( &plan->BT, std::size_t( nq ) * Nc_global * d.Pb * Ha_loc * sizeof( double2 ) ) );
        SB_CUDA_CHECK( cudaMalloc( &plan->flags, std::size_t( world ) * sizeof( unsigned ) ) );
        SB_CUDA_CHECK( cudaMemset( plan->flags, 0, std::size_t( world ) * sizeof( unsigned ) ) );
        const bool m0 = peer_map_all( plan->AT, plan->AT_peer, plan->peer_opened, stream );
        const bool m1 = peer_map_all( plan->flags, plan->flags_peer, plan->peer_opened, stream );
        if( !m0 || !m1 )
            throw std::runtime_error( "spirit_b200: the distributed dipolar convolution could not map its operands into the peers' address "
                                      "spaces (CUDA IPC); SPIRIT_B200_DDI_PENCIL=0 selects the decomposition with an NCCL fallback" );
        for( int r = 0; r < world; ++r )
            plan->push.base[r] = static_cast<double2 *>( plan->AT_peer[r] );
        // (collective calls: every rank takes the same branch, the choice comes from the environment of the job)
        if( plan->pencil_dma )
        {
            const std::size_t n_A = std::size_t( nq ) * ncl * d.Nb * d.Ha;
            SB_CUDA_CHECK( cudaMalloc( &plan->A, n_A * sizeof( double2 ) ) );
            SB_CUDA_CHECK( cudaMalloc( &plan->A_recv, n_A * sizeof( double2 ) ) );
            if( !peer_map_all( plan->A_recv, plan->A_recv_peer, plan->peer_opened, stream ) )
                throw std::runtime_error( "spirit_b200: the distributed dipolar convolution could not map its operands into the peers' address spaces" );
        }
    }
    else
        SB_CUDA_CHECK( cudaMalloc( &plan->A, std::size_t( nq ) * ncl * d.Nb * d.Ha * sizeof( double2 ) ) );
    if( world > 1 && world <= DDI_MAX_PEERS && plan->fast_a.on && plan->fast_b.on && plan->lg_split != 31 )
    {
        // peer-mapped operands (collective; the same outcome on every rank)
        SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
        bool ok = peer_memops_available();
        for( int k = 0; k < 2; ++k )
            ok = cudaMalloc( &plan->C2[k], n_B * sizeof( double2 ) ) == cudaSuccess && ok;
        ok = cudaMalloc( &plan->flags, std::size_t( world ) * sizeof( unsigned ) ) == cudaSuccess && ok;
        if( !ok )
            cudaGetLastError();
        else
            SB_CUDA_CHECK( cudaMemset( plan->flags, 0, std::size_t( world ) * sizeof( unsigned ) ) );
        // every rank takes part in every collective call, whatever its own outcome so far
        const bool m0 = peer_map_all( ok ? plan->C2[0] : nullptr, plan->C2_peer[0], plan->peer_opened, stream );
        const bool m1 = peer_map_all( ok ? plan->C2[1] : nullptr, plan->C2_peer[1], plan->peer_opened, stream );
        const bool m2 = peer_map_all( ok ? plan->flags : nullptr, plan->flags_peer, plan->peer_opened, stream );
        plan->peer    = m0 && m1 && m2;
        if( !plan->peer )
        {
            for( double2 *& c : plan->C2 )
            {
                if( c )
                    cudaFree( c );
                c = nullptr;
            }
        }
    }
    if( !plan->peer && !plan->pencil )
    {
        SB_CUDA_CHECK( cudaMalloc( &plan->B, n_B * sizeof( double2 ) ) );
        if( world > 1 )
            SB_CUDA_CHECK( cudaMalloc( &plan->C, n_B * sizeof( double2 ) ) );
    }

    // tensor spectrum, one component at a time: real D -> rows (a) -> b -> c on the WHOLE padded lattice (every rank
    // computes it redundantly: the tensor is analytic, no communication), then the local kb range is kept
    double * Dreal = nullptr;
    double2 *tmp1 = nullptr, *tmp2 = nullptr;
    SB_CUDA_CHECK( cudaMalloc( &Dreal, full * sizeof( double ) ) );
    SB_CUDA_CHECK( cudaMalloc( &tmp1, half * sizeof( double2 ) ) );
    SB_CUDA_CHECK( cudaMalloc( &tmp2, half * sizeof( double2 ) ) );
    TensorGeom tg{};
    for( int k = 0; k < 3; ++k )
    {
        tg.ta[k] = g.lattice_constant * g.bravais_vectors[0][k];
        tg.tb[k] = g.lattice_constant * g.bravais_vectors[1][k];
        tg.tc[k] = g.lattice_constant * g.bravais_vectors[2][k];
        tg.img[k] = ham.boundary_conditions[k] == 0 ? 0 : ham.ddi_n_periodic_images[k];
    }
    tg.mult = constants::mu_0 * constants::mu_B * constants::mu_B / ( 4 * constants::Pi * 1e-30 );
    const int ncol_setup = plan->ncol_b;
    for( int b1 = 0; b1 < d.NB; ++b1 )
        for( int b2 = 0; b2 < d.NB; ++b2 )
        {
            if( b1 == b2 && b1 != 0 )
                continue;
            const int inter = d.lookup[b1 + b2 * d.NB];
            tg.da           = g.cell_atoms[b1][0] - g.cell_atoms[b2][0];
            tg.db           = g.cell_atoms[b1][1] - g.cell_atoms[b2][1];
            tg.dc           = g.cell_atoms[b1][2] - g.cell_atoms[b2][2];
            for( int comp6 = 0; comp6 < 6; ++comp6 )
            {
                double2 * Dout = plan->Dhat + std::size_t( 6 * inter + comp6 ) * half_local;
                k_ddi_tensor<<<unsigned( ( full + 255 ) / 256 ), 256, 0, stream>>>( dg, tg, comp6, Dreal );
                k_fft_real_rows<<<unsigned( std::size_t( d.Pb ) * d.Pc ), fft_threads( d.Pa / 4 ), plan->smem_a, stream>>>( plan->plan[0], Dreal, tmp1, d.Ha );
                // b: tmp1[c][b][ka] -> tmp2[c][kb][ka]
                PassArgs pb{};
                pb.in = tmp1, pb.out = tmp2;
                pb.in_os = pb.out_os = std::size_t( d.Pb ) * d.Ha;
                pb.in_js = pb.out_js = d.Ha;
                pb.n_u = d.Ha, pb.n_o = d.Pc, pb.n_in = d.Pb, pb.n_out = d.Pb, pb.ncol = ncol_setup, pb.scale = 1.0;
                k_fft_pass<false><<<dim3( ( d.Ha + ncol_setup - 1 ) / ncol_setup, d.Pc ), fft_threads( d.Pb / 4 * ncol_setup ), plan->smem_b, stream>>>( plan->plan[1], pb );
                // c: tmp2[c][kb][ka] -> tmp1[kc][kb][ka]; (kb, ka) is one contiguous index of length Pb*Ha
                PassArgs pc{};
                pc.in = tmp2, pc.out = tmp1;
                pc.in_os = pc.out_os = 0;
                pc.in_js = pc.out_js = std::size_t( d.Pb ) * d.Ha;
                pc.n_u = d.Pb * d.Ha, pc.n_o = 1, pc.n_in = d.Pc, pc.n_out = d.Pc, pc.scale = 1.0;
                pc.ncol = choose_ncol( d.Pc, 1, pc.n_u );
                const std::size_t smem_c1 = std::size_t( d.Pc ) * 2 * sizeof( double2 ) * pc.ncol;
                allow_smem( k_fft_pass<false>, std::max( plan->smem_b, smem_c1 ) );
                k_fft_pass<false><<<dim3( ( pc.n_u + pc.ncol - 1 ) / pc.ncol, 1 ), fft_threads( std::max( 1, d.Pc / 4 ) * pc.ncol ), smem_c1, stream>>>( plan->plan[2], pc );
                // keep the local kb range: [kc][kb in range][ka]   (ka pencils: [kc][kb][ka in block])
                if( plan->pencil )
                    SB_CUDA_CHECK( cudaMemcpy2DAsync(
                        Dout, std::size_t( Ha_loc ) * sizeof( double2 ), tmp1 + ka0, std::size_t( d.Ha ) * sizeof( double2 ),
                        std::size_t( Ha_loc ) * sizeof( double2 ), std::size_t( d.Pc ) * d.Pb, cudaMemcpyDeviceToDevice, stream ) );
                else
                SB_CUDA_CHECK( cudaMemcpy2DAsync(
                    Dout, std::size_t( kbl ) * d.Ha * sizeof( double2 ), tmp1 + std::size_t( rank ) * kbl * d.Ha,
                    std::size_t( d.Pb ) * d.Ha * sizeof( double2 ), std::size_t( kbl ) * d.Ha * sizeof( double2 ), d.Pc,
                    cudaMemcpyDeviceToDevice, stream ) );
                plan->launches_setup += 4;
            }
        }
    SB_CUDA_CHECK( cudaGetLastError() );
    SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
    cudaFree( Dreal );
    cudaFree( tmp1 );
    cudaFree( tmp2 );

    // Single sublattice: D(-r) = D(r), the spectrum is real. Verify numerically, then keep only the real parts.
    const bool small_c = d.NB == 1 && ( d.Pc & ( d.Pc - 1 ) ) == 0 && d.Pc <= 32;
    if( small_c )
        plan->fast_c.on = false; // the in-register kernel serves thin films
    if( !generic_c_fits && !plan->fast_c.on )
        throw std::runtime_error( "spirit_b200: padded lattice dimension c too long for the shared-memory FFT passes" );
    bool spectrum_real = false;
    if( small_c || plan->fast_c.on )
    {
        const std::size_t n_all = std::size_t( 6 * d.n_inter ) * half_local;
        const int blocks        = 1024;
        double * part           = nullptr;
        SB_CUDA_CHECK( cudaMalloc( &part, 2 * blocks * sizeof( double ) ) );
        k_ddi_imag_check<<<blocks, BLOCK_THREADS, 0, stream>>>( plan->Dhat, n_all, part );
        std::vector<double> h( 2 * blocks );
        SB_CUDA_CHECK( cudaMemcpyAsync( h.data(), part, h.size() * sizeof( double ), cudaMemcpyDeviceToHost, stream ) );
        SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
        double max_imag = 0, max_abs = 0;
        for( int i = 0; i < blocks; ++i )
        {
            max_imag = std::max( max_imag, h[2 * i] );
            max_abs  = std::max( max_abs, h[2 * i + 1] );
        }
        if( world > 1 )
        {
            // every rank takes the same decision: maxima over the whole spectrum, not over the local kb range
            const double mine[2] = { max_imag, max_abs };
            double all[2];
            SB_CUDA_CHECK( cudaMemcpyAsync( part, mine, sizeof( mine ), cudaMemcpyHostToDevice, stream ) );
            comm_allreduce( part, 2, true, stream );
            SB_CUDA_CHECK( cudaMemcpyAsync( all, part, sizeof( all ), cudaMemcpyDeviceToHost, stream ) );
            SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
            max_imag = all[0], max_abs = all[1];
        }
        cudaFree( part );
        if( std::getenv( "SPIRIT_B200_DDI_VERBOSE" ) )
            std::fprintf( stderr, "spirit_b200 ddi: max |Im D^| = %.3e, max |D^| = %.3e\n", max_imag, max_abs );
        // The imaginary parts are pure round-off of the forward transforms (a few ulp of the largest element times
        // log2 P): mathematically zero. 1e-12 relative is 4 orders above what is observed and far below any signal.
        spectrum_real = max_imag <= 1e-12 * max_abs;
        // Mirror symmetries (orthorhombic axis-aligned lattices, open or padded directions): checked on the spectrum itself,
        // along c on every rank's kb range (all ranks must agree), along b on one device only (a rank holds a kb block).
        if( spectrum_real && d.n_inter == 1 && !env_flag_off( "SPIRIT_B200_DDI_MIRROR" ) )
        {
            double * mpart = nullptr;
            SB_CUDA_CHECK( cudaMalloc( &mpart, 2 * blocks * sizeof( double ) ) );
            k_ddi_mirror_check<<<blocks, BLOCK_THREADS, 0, stream>>>( dc, plan->Dhat, mpart );
            SB_CUDA_CHECK( cudaMemcpyAsync( h.data(), mpart, h.size() * sizeof( double ), cudaMemcpyDeviceToHost, stream ) );
            SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
            double dev_c = 0, dev_b = 0;
            for( int i = 0; i < blocks; ++i )
            {
                dev_c = std::max( dev_c, h[2 * i] );
                dev_b = std::max( dev_b, h[2 * i + 1] );
            }
            if( world > 1 )
            {
                double all = dev_c;
                SB_CUDA_CHECK( cudaMemcpyAsync( mpart, &all, sizeof( all ), cudaMemcpyHostToDevice, stream ) );
                comm_allreduce( mpart, 1, true, stream );
                SB_CUDA_CHECK( cudaMemcpyAsync( &all, mpart, sizeof( all ), cudaMemcpyDeviceToHost, stream ) );
                SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
                dev_c = all;
            }
            cudaFree( mpart );
            int mirror = 0;
            if( dev_c <= 1e-12 * max_abs )
                mirror |= 1;
            if( ( world == 1 || plan->pencil ) && dev_b <= 1e-12 * max_abs ) // (needs the whole kb axis on this rank)
                mirror |= 2;
            if( std::getenv( "SPIRIT_B200_DDI_VERBOSE" ) )
                std::fprintf( stderr, "spirit_b200 ddi: mirror deviations c %.3e b %.3e -> mirror %d\n", dev_c, dev_b, mirror );
            plan->dims.mirror = plan->dims_c.mirror = dc.mirror = mirror;
        }
        if( spectrum_real && small_c )
        {
            const std::size_t n_kept = std::size_t( 6 ) * mirror_len( dc.Pc, dc.mirror & 1 ) * mirror_len( dc.Pb, dc.mirror & 2 ) * dc.Ha;
            SB_CUDA_CHECK( cudaMalloc( &plan->Dhat_real, n_kept * sizeof( double ) ) );
            k_ddi_take_real<<<unsigned( ( n_kept + 255 ) / 256 ), 256, 0, stream>>>( dc, plan->Dhat, plan->Dhat_real );
            SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
            cudaFree( plan->Dhat );
            plan->Dhat = nullptr;
        }
    }
    if( plan->fast_c.on )
    {
        // re-order the spectrum into the tiles k_ddi_c_mult16 reads (one sublattice: 6 components)
        const int lg            = plan->fast_c.lg;
        const std::size_t tiles = std::size_t( ( dc.Ha + ( 1 << lg ) - 1 ) >> lg );
        const std::size_t n_t   = std::size_t( mirror_len( kbl, dc.mirror & 2 ) ) * tiles * 6 * ( std::size_t( mirror_len( d.Pc, dc.mirror & 1 ) ) << lg );
        plan->Dt_real           = spectrum_real;
        const std::size_t bytes = n_t * ( spectrum_real ? sizeof( double ) : sizeof( double2 ) );
        SB_CUDA_CHECK( cudaMalloc( &plan->Dt, bytes ) );
        SB_CUDA_CHECK( cudaMemsetAsync( plan->Dt, 0, bytes, stream ) );
        const std::size_t n_in = 6 * half_local;
        if( spectrum_real )
            k_ddi_tile_tensor<true><<<unsigned( ( n_in + 255 ) / 256 ), 256, 0, stream>>>( dc, plan->Dhat, plan->Dt, lg );
        else
            k_ddi_tile_tensor<false><<<unsigned( ( n_in + 255 ) / 256 ), 256, 0, stream>>>( dc, plan->Dhat, plan->Dt, lg );
        SB_CUDA_CHECK( cudaGetLastError() );
        SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
        cudaFree( plan->Dhat );
        plan->Dhat = nullptr;
    }
    return plan.release();
}

namespace
{
// a-passes of components q0 .. q0 + nq - 1 (fast kernels), plain or pipelined
void run_fwd_a16( DDIPlan & plan, ConstField3 spins, int q0, int nq, cudaStream_t stream )
{
    const DDIDims & d = plan.dims;
    const int n_rg    = ( d.Nb * d.Nc + ( 1 << plan.fast_a.lg ) - 1 ) >> plan.fast_a.lg;
    if( plan.a_pipe )
        launch_fwd_a16p( plan.fast_a, plan.apipe, n_rg, nq, stream, plan.plan_ah, plan.plan[0].twiddle, d, spins, plan.A, q0, plan.push );
    else
        launch_fwd_a16( plan.fast_a, dim3( n_rg, nq ), stream, plan.plan_ah, plan.plan[0].twiddle, d, spins, plan.A, q0, false, plan.push );
}
void run_inv_a16( DDIPlan & plan, Field3 g_ddi, double inv_P, int q0, int nq, cudaStream_t stream )
{
    const DDIDims & d   = plan.dims;
    const double2 * src = plan.A_recv ? plan.A_recv : plan.A;
    const int n_rg    = ( d.Nb * d.Nc + ( 1 << plan.fast_a.lg ) - 1 ) >> plan.fast_a.lg;
    if( plan.a_pipe )
        launch_inv_a16p( plan.fast_a, plan.apipe, n_rg, nq, stream, plan.plan_ah, plan.plan[0].twiddle, d, src, g_ddi, inv_P, q0, plan.push );
    else
        launch_inv_a16( plan.fast_a, dim3( n_rg, nq ), stream, plan.plan_ah, plan.plan[0].twiddle, d, src, g_ddi, inv_P, q0, false, plan.push );
}
// step 3 (c-transforms + tensor multiply) on the operand in per-rank block layout
void launch_c_mult( DDIPlan & plan, double2 * operand, cudaStream_t stream )
{
    const DDIDims & d  = plan.dims;
    const DDIDims & dc = plan.dims_c;
    const int nq       = 3 * d.NB;
    const int kbl      = dc.Pb;
    const bool small_c = d.NB == 1 && ( dc.Pc & ( dc.Pc - 1 ) ) == 0 && dc.Pc <= 32;
    if( small_c )
    {
        const dim3 grid( ( dc.Ha + 127 ) / 128, kbl );
        const void * D = plan.Dhat_real ? static_cast<const void *>( plan.Dhat_real ) : static_cast<const void *>( plan.Dhat );
#define SB_DDI_SMALL( PC )                                                                                             \
    case PC:                                                                                                           \
        if( plan.Dhat_real )                                                                                           \
            k_ddi_c_mult_small<PC, true><<<grid, 128, 0, stream>>>( dc, operand, D );                                  \
        else                                                                                                           \
            k_ddi_c_mult_small<PC, false><<<grid, 128, 0, stream>>>( dc, operand, D );                                 \
        break;
        switch( dc.Pc )
        {
            SB_DDI_SMALL( 1 )
            SB_DDI_SMALL( 2 )
            SB_DDI_SMALL( 4 )
            SB_DDI_SMALL( 8 )
            SB_DDI_SMALL( 16 )
            SB_DDI_SMALL( 32 )
        }
#undef SB_DDI_SMALL
    }
    else if( plan.fast_c.on )
    {
        const dim3 grid( ( dc.Ha + ( 1 << plan.fast_c.lg ) - 1 ) >> plan.fast_c.lg, kbl );
        if( plan.Dt_real )
            launch_c_mult16<true>( plan.fast_c, grid, stream, plan.plan[2], dc, operand, plan.Dt );
        else
            launch_c_mult16<false>( plan.fast_c, grid, stream, plan.plan[2], dc, operand, plan.Dt );
    }
    else
    {
        if( !plan.Dhat )
            throw std::logic_error( "spirit_b200: real tensor spectrum with the shared-memory multiply kernel" );
        k_ddi_c_mult<<<dim3( ( dc.Ha + plan.ncol_c - 1 ) / plan.ncol_c, kbl ), fft_threads( std::max( 1, dc.Pc / 4 ) * plan.ncol_c * nq ), plan.smem_c, stream>>>(
            plan.plan[2], dc, operand, plan.Dhat, plan.ncol_c );
    }
}
} // namespace

namespace
{

// Distributed evaluation over peer-mapped memory: the transposes are the stores of the forward b-pass (into the c-pass
// operands of the ranks that own each kb block) and the loads of the inverse b-pass (from there). Two meetings of all ranks
// per evaluation, done with 32-bit counters and stream memory operations: (1) every rank has stored its blocks, before the
// c-passes; (2) every c-pass is done, before the inverse b-passes read. The operand is double-buffered, so the stores of the
// next evaluation cannot overtake the loads of this one (a rank reaches meeting (1) of evaluation e + 1 only after its
// loads of evaluation e).
void ddi_peer_barrier( DDIPlan & plan, cudaStream_t stream )
{
    const unsigned value = ++plan.barrier_count;
    for( int r = 0; r < plan.world; ++r )
        if( r != plan.rank )
            peer_write32( stream, static_cast<unsigned *>( plan.flags_peer[r] ) + plan.rank, value );
    for( int r = 0; r < plan.world; ++r )
        if( r != plan.rank )
            peer_wait_geq32( stream, plan.flags + r, value );
}

int ddi_gradient_peer( DDIPlan & plan, ConstField3 spins, Field3 g_ddi, cudaStream_t stream )
{
    const DDIDims & d  = plan.dims;
    const DDIDims & dc = plan.dims_c;
    const int world = plan.world, nq = 3 * d.NB, kbl = dc.Pb, ncl = d.Nc, rows = d.Nb * d.Nc;
    const int buf                = int( plan.evaluation++ & 1u );
    const std::size_t comp_elems = dc.q_stride; // one component inside a per-rank block: ncl * kbl * Ha
    const dim3 grid_a( ( rows + ( 1 << plan.fast_a.lg ) - 1 ) >> plan.fast_a.lg, 1 );
    const int lg_tile_b = plan.fast_b.lg + plan.fast_b.lg_seq;
    const dim3 grid_b( ( d.Ha + ( 1 << lg_tile_b ) - 1 ) >> lg_tile_b, ncl );
    for( int q = 0; q < nq; ++q )
    {
        run_fwd_a16( plan, spins, q, 1, stream );
        PassArgs pb{};
        pb.in    = plan.A + std::size_t( q ) * ncl * d.Nb * d.Ha;
        pb.in_os = std::size_t( d.Nb ) * d.Ha, pb.in_js = pb.out_js = d.Ha;
        pb.out_os = std::size_t( kbl ) * d.Ha, pb.out_split = kbl, pb.out_split_stride = dc.block_stride;
        pb.n_u = d.Ha, pb.n_o = ncl, pb.n_in = d.Nb, pb.n_out = d.Pb, pb.ncol = 1 << plan.fast_b.lg, pb.scale = 1.0;
        pb.use_out_peer = 1;
        for( int r = 0; r < world; ++r ) // my block inside rank r's operand: [source rank][component][plane][kb][ka]
            pb.out_peer[r] = static_cast<double2 *>( plan.C2_peer[buf][r] ) + std::size_t( plan.rank ) * dc.block_stride + std::size_t( q ) * comp_elems;
        launch_pass16<false>( plan.fast_b, grid_b, stream, plan.plan[1], pb, 31, plan.lg_split );
    }
    ddi_peer_barrier( plan, stream );
    launch_c_mult( plan, plan.C2[buf], stream );
    ddi_peer_barrier( plan, stream );
    const double inv_P = 1.0 / ( double( d.Pa ) * d.Pb * d.Pc );
    for( int q = 0; q < nq; ++q )
    {
        PassArgs ib{};
        ib.out    = plan.A + std::size_t( q ) * ncl * d.Nb * d.Ha;
        ib.out_os = std::size_t( d.Nb ) * d.Ha, ib.in_js = ib.out_js = d.Ha;
        ib.in_os = std::size_t( kbl ) * d.Ha, ib.in_split = kbl, ib.in_split_stride = dc.block_stride;
        ib.n_u = d.Ha, ib.n_o = ncl, ib.n_in = d.Pb, ib.n_out = d.Nb, ib.ncol = 1 << plan.fast_b.lg, ib.scale = 1.0;
        ib.use_in_peer = 1;
        for( int r = 0; r < world; ++r ) // the kb block of rank r for my planes: block `my rank` of rank r's operand
            ib.in_peer[r] = static_cast<const double2 *>( plan.C2_peer[buf][r] ) + std::size_t( plan.rank ) * dc.block_stride + std::size_t( q ) * comp_elems;
        launch_pass16<true>( plan.fast_b, grid_b, stream, plan.plan[1], ib, plan.lg_split, 31 );
        run_inv_a16( plan, g_ddi, inv_P, q, 1, stream );
    }
    SB_CUDA_CHECK( cudaGetLastError() );
    return 4 * nq + 1;
}

// Distributed evaluation with ka PENCILS. Every rank runs the a-passes on its planes and the b- and c-passes on its block of ka,
// for all planes of the lattice:
//   1 forward a-pass of the local planes; element ka of a row is stored into the operand AT of the rank that owns ka (peer-mapped
//     memory, NVLink): the transpose IS the store of the kernel, contiguous runs of a whole ka block per row
//   -- meeting of all ranks (32-bit counters, stream memory operations) --
//   2 b-pass, c-pass x tensor x inverse c-pass, inverse b-pass on [3][Nc_global][b][ka block]: the kernels of one device
//   -- meeting --
//   3 inverse a-pass of the local planes, its rows gathered from the ranks' AT (loads over NVLink)
// The transposed data is the a-pass output, where b is not yet zero-padded: half the volume of a transpose of the b-pass output
// (the kb-block decompositions below). One buffer suffices: a rank only ever touches the rows of its own planes in a peer's AT.
// SPIRIT_B200_DDI_TIMING=1: CUDA events around the phases of the pencil evaluation, averages printed every 8 evaluations
// (profiles/: where the time of a distributed evaluation goes; one stream, so the events serialise nothing)
struct PhaseTimer
{
    static constexpr int N = 8;
    cudaEvent_t ev[N] = {};
    double sum[N]     = {};
    int count = 0, on = -1;
    bool enabled()
    {
        if( on < 0 )
        {
            on = std::getenv( "SPIRIT_B200_DDI_TIMING" ) ? 1 : 0;
            if( on )
                for( auto & e : ev )
                    cudaEventCreate( &e );
        }
        return on == 1;
    }
    void mark( int i, cudaStream_t stream )
    {
        if( enabled() )
            cudaEventRecord( ev[i], stream );
    }
    void finish( int n_marks, int rank, cudaStream_t stream, const char * const * names )
    {
        if( !enabled() )
            return;
        cudaStreamSynchronize( stream );
        for( int i = 1; i < n_marks; ++i )
        {
            float ms = 0;
            cudaEventElapsedTime( &ms, ev[i - 1], ev[i] );
            sum[i] += ms;
        }
        if( ++count % 8 == 0 )
        {
            std::fprintf( stderr, "spirit_b200 ddi timing rank %d:", rank );
            for( int i = 1; i < n_marks; ++i )
            {
                std::fprintf( stderr, " %s %.3f", names[i], sum[i] / 8 );
                sum[i] = 0;
            }
            std::fprintf( stderr, " ms\n" );
        }
    }
};
PhaseTimer g_phase_timer;

int ddi_gradient_pencil( DDIPlan & plan, ConstField3 spins, Field3 g_ddi, cudaStream_t stream )
{
    const DDIDims & d  = plan.dims;
    const DDIDims & dc = plan.dims_c;
    const int nq = 3, w = dc.Ha, planes = dc.Nc;
    if( plan.pencil_dma )
    {
        // The a-passes stay local; the ka blocks travel as 2-D peer copies on the copy engines (one stream per destination), one
        // component at a time, under the passes of the other components: no SM takes part in the transposes.
        cudaStream_t cs     = plan.comm_stream;
        const int world = plan.world, rank = plan.rank, ncl = d.Nc, kblk = 1 << plan.push.lg_kblock;
        const int lg_tile_b = plan.fast_b.lg + plan.fast_b.lg_seq;
        const dim3 grid_b( ( w + ( 1 << lg_tile_b ) - 1 ) >> lg_tile_b, planes );
        const unsigned base = plan.barrier_count;
        plan.barrier_count += 2 * nq;
        const std::size_t at_q = std::size_t( planes ) * d.Nb * w, bt_q = std::size_t( planes ) * d.Pb * w;
        const std::size_t a_q  = std::size_t( ncl ) * d.Nb * d.Ha, rows = std::size_t( ncl ) * d.Nb, E = sizeof( double2 );
        auto wait_all = [&]( cudaStream_t st, unsigned value )
        {
            for( int r = 0; r < world; ++r )
                if( r != rank )
                    peer_wait_geq32( st, plan.flags + r, value );
        };
        for( int q = 0; q < nq; ++q )
        {
            run_fwd_a16( plan, spins, q, 1, stream ); // -> A[q][c][b][ka]
            SB_CUDA_CHECK( cudaEventRecord( plan.ev_dma[q], stream ) );
            for( int r = 0; r < world; ++r )
            {
                cudaStream_t st = plan.copy_stream[r];
                const int wr    = plan.push.w[r];
                SB_CUDA_CHECK( cudaStreamWaitEvent( st, plan.ev_dma[q], 0 ) );
                // my planes' rows of component q, columns of rank r's ka block -> AT of rank r
                double2 * dst = static_cast<double2 *>( plan.AT_peer[r] ) + ( std::size_t( q ) * planes + std::size_t( rank ) * ncl ) * d.Nb * wr;
                SB_CUDA_CHECK( cudaMemcpy2DAsync(
                    dst, wr * E, plan.A + q * a_q + std::size_t( r ) * kblk, d.Ha * E, wr * E, rows, cudaMemcpyDeviceToDevice, st ) );
                if( r != rank )
                    peer_write32( st, static_cast<unsigned *>( plan.flags_peer[r] ) + rank, base + 1 + q );
                else
                    SB_CUDA_CHECK( cudaEventRecord( plan.ev_dma[3 + q], st ) );
            }
            SB_CUDA_CHECK( cudaStreamWaitEvent( cs, plan.ev_dma[3 + q], 0 ) );
            wait_all( cs, base + 1 + q );
            PassArgs pb{};
            pb.in = plan.AT + q * at_q, pb.out = plan.BT + q * bt_q;
            pb.in_os = std::size_t( d.Nb ) * w, pb.out_os = std::size_t( d.Pb ) * w, pb.in_js = pb.out_js = w;
            pb.n_u = w, pb.n_o = planes, pb.n_in = d.Nb, pb.n_out = d.Pb, pb.ncol = 1 << plan.fast_b.lg, pb.scale = 1.0;
            launch_pass16<false>( plan.fast_b, grid_b, cs, plan.plan[1], pb, 31, 31 );
        }
        launch_c_mult( plan, plan.BT, cs );
        const double inv_P = 1.0 / ( double( d.Pa ) * d.Pb * d.Pc );
        for( int q = 0; q < nq; ++q )
        {
            PassArgs ib{};
            ib.in = plan.BT + q * bt_q, ib.out = plan.AT + q * at_q;
            ib.in_os = std::size_t( d.Pb ) * w, ib.out_os = std::size_t( d.Nb ) * w, ib.in_js = ib.out_js = w;
            ib.n_u = w, ib.n_o = planes, ib.n_in = d.Pb, ib.n_out = d.Nb, ib.ncol = 1 << plan.fast_b.lg, ib.scale = 1.0;
            launch_pass16<true>( plan.fast_b, grid_b, cs, plan.plan[1], ib, 31, 31 );
            SB_CUDA_CHECK( cudaEventRecord( plan.ev_dma[6 + q], cs ) );
            for( int r = 0; r < world; ++r )
            {
                cudaStream_t st = plan.copy_stream[r];
                SB_CUDA_CHECK( cudaStreamWaitEvent( st, plan.ev_dma[6 + q], 0 ) );
                // the rows of rank r's planes, my ka block -> the inverse a-pass operand of rank r
                double2 * dst = static_cast<double2 *>( plan.A_recv_peer[r] ) + q * a_q + std::size_t( rank ) * kblk;
                const double2 * src = plan.AT + ( std::size_t( q ) * planes + std::size_t( r ) * ncl ) * d.Nb * w;
                SB_CUDA_CHECK( cudaMemcpy2DAsync( dst, d.Ha * E, src, w * E, w * E, rows, cudaMemcpyDeviceToDevice, st ) );
                if( r != rank )
                    peer_write32( st, static_cast<unsigned *>( plan.flags_peer[r] ) + rank, base + 1 + nq + q );
                else
                    SB_CUDA_CHECK( cudaEventRecord( plan.ev_dma[9 + q], st ) );
            }
            SB_CUDA_CHECK( cudaStreamWaitEvent( stream, plan.ev_dma[9 + q], 0 ) );
            wait_all( stream, base + 1 + nq + q );
            run_inv_a16( plan, g_ddi, inv_P, q, 1, stream );
        }
        SB_CUDA_CHECK( cudaGetLastError() );
        return 4 * nq + 1;
    }
    if( plan.pencil_overlap && !g_phase_timer.enabled() )
    {
        // Per component: a-pass + push on `stream`, everything on the local ka block on the plan's own stream `cs`; a component's
        // b-pass starts when every rank has pushed that component (one meeting per component and direction), under the a-pass of
        // the next component; on the way back the pull + inverse a-pass of component q runs under the inverse b-pass of q + 1.
        cudaStream_t cs     = plan.comm_stream;
        const int lg_tile_b = plan.fast_b.lg + plan.fast_b.lg_seq;
        const dim3 grid_b( ( w + ( 1 << lg_tile_b ) - 1 ) >> lg_tile_b, planes );
        auto signal = [&]( cudaStream_t st, unsigned value )
        {
            for( int r = 0; r < plan.world; ++r )
                if( r != plan.rank )
                    peer_write32( st, static_cast<unsigned *>( plan.flags_peer[r] ) + plan.rank, value );
        };
        auto wait_all = [&]( cudaStream_t st, unsigned value )
        {
            for( int r = 0; r < plan.world; ++r )
                if( r != plan.rank )
                    peer_wait_geq32( st, plan.flags + r, value );
        };
        const unsigned base = plan.barrier_count;
        plan.barrier_count += 2 * nq;
        const std::size_t at_q = std::size_t( planes ) * d.Nb * w, bt_q = std::size_t( planes ) * d.Pb * w;
        for( int q = 0; q < nq; ++q )
        {
            run_fwd_a16( plan, spins, q, 1, stream );
            signal( stream, base + 1 + q );
            SB_CUDA_CHECK( cudaEventRecord( plan.ev_q[q], stream ) );
            SB_CUDA_CHECK( cudaStreamWaitEvent( cs, plan.ev_q[q], 0 ) );
            wait_all( cs, base + 1 + q );
            PassArgs pb{};
            pb.in = plan.AT + q * at_q, pb.out = plan.BT + q * bt_q;
            pb.in_os = std::size_t( d.Nb ) * w, pb.out_os = std::size_t( d.Pb ) * w, pb.in_js = pb.out_js = w;
            pb.n_u = w, pb.n_o = planes, pb.n_in = d.Nb, pb.n_out = d.Pb, pb.ncol = 1 << plan.fast_b.lg, pb.scale = 1.0;
            launch_pass16<false>( plan.fast_b, grid_b, cs, plan.plan[1], pb, 31, 31 );
        }
        launch_c_mult( plan, plan.BT, cs );
        const double inv_P = 1.0 / ( double( d.Pa ) * d.Pb * d.Pc );
        for( int q = 0; q < nq; ++q )
        {
            PassArgs ib{};
            ib.in = plan.BT + q * bt_q, ib.out = plan.AT + q * at_q;
            ib.in_os = std::size_t( d.Pb ) * w, ib.out_os = std::size_t( d.Nb ) * w, ib.in_js = ib.out_js = w;
            ib.n_u = w, ib.n_o = planes, ib.n_in = d.Pb, ib.n_out = d.Nb, ib.ncol = 1 << plan.fast_b.lg, ib.scale = 1.0;
            launch_pass16<true>( plan.fast_b, grid_b, cs, plan.plan[1], ib, 31, 31 );
            signal( cs, base + 1 + nq + q );
            SB_CUDA_CHECK( cudaEventRecord( plan.ev_q[nq + q], cs ) );
            SB_CUDA_CHECK( cudaStreamWaitEvent( stream, plan.ev_q[nq + q], 0 ) );
            wait_all( stream, base + 1 + nq + q );
            run_inv_a16( plan, g_ddi, inv_P, q, 1, stream );
        }
        SB_CUDA_CHECK( cudaGetLastError() );
        return 4 * nq + 1;
    }
    PhaseTimer & T = g_phase_timer;
    T.mark( 0, stream );
    run_fwd_a16( plan, spins, 0, nq, stream );
    T.mark( 1, stream );
    ddi_peer_barrier( plan, stream );
    T.mark( 2, stream );
    const int lg_tile_b = plan.fast_b.lg + plan.fast_b.lg_seq;
    const dim3 grid_b( ( w + ( 1 << lg_tile_b ) - 1 ) >> lg_tile_b, nq * planes );
    PassArgs pb{};
    pb.in = plan.AT, pb.out = plan.BT;
    pb.in_os = std::size_t( d.Nb ) * w, pb.out_os = std::size_t( d.Pb ) * w, pb.in_js = pb.out_js = w;
    pb.n_u = w, pb.n_o = nq * planes, pb.n_in = d.Nb, pb.n_out = d.Pb, pb.ncol = 1 << plan.fast_b.lg, pb.scale = 1.0;
    launch_pass16<false>( plan.fast_b, grid_b, stream, plan.plan[1], pb, 31, 31 );
    T.mark( 3, stream );
    launch_c_mult( plan, plan.BT, stream );
    T.mark( 4, stream );
    PassArgs ib{};
    ib.in = plan.BT, ib.out = plan.AT;
    ib.in_os = std::size_t( d.Pb ) * w, ib.out_os = std::size_t( d.Nb ) * w, ib.in_js = ib.out_js = w;
    ib.n_u = w, ib.n_o = nq * planes, ib.n_in = d.Pb, ib.n_out = d.Nb, ib.ncol = 1 << plan.fast_b.lg, ib.scale = 1.0;
    launch_pass16<true>( plan.fast_b, grid_b, stream, plan.plan[1], ib, 31, 31 );
    T.mark( 5, stream );
    ddi_peer_barrier( plan, stream );
    T.mark( 6, stream );
    const double inv_P = 1.0 / ( double( d.Pa ) * d.Pb * d.Pc );
    run_inv_a16( plan, g_ddi, inv_P, 0, nq, stream );
    T.mark( 7, stream );
    static const char * const names[8] = { "", "a+push", "wait", "b", "c", "b^-1", "wait", "pull+a^-1" };
    T.finish( 8, plan.rank, stream, names );
    SB_CUDA_CHECK( cudaGetLastError() );
    return 5;
}

// Distributed evaluation with the transposes hidden: component q travels (NCCL send / recv on the plan's own stream)
// while the a- and b-passes of component q + 1 run, and on the way back the inverse passes of component q run while
// component q + 1 travels. Same kernels, same per-rank block layout, same result as the plain sequence below.
int ddi_gradient_pipelined( DDIPlan & plan, ConstField3 spins, Field3 g_ddi, cudaStream_t stream )
{
    const DDIDims & d  = plan.dims;
    const DDIDims & dc = plan.dims_c;
    const int world = plan.world, nq = 3 * d.NB, kbl = dc.Pb, ncl = d.Nc, rows = d.Nb * d.Nc;
    cudaStream_t cs                  = plan.comm_stream;
    const std::size_t comp_elems     = dc.q_stride; // one component inside a per-rank block: ncl * kbl * Ha
    const dim3 grid_a( ( rows + ( 1 << plan.fast_a.lg ) - 1 ) >> plan.fast_a.lg, 1 );
    const int lg_tile_b = plan.fast_b.lg + plan.fast_b.lg_seq;
    const dim3 grid_b( ( d.Ha + ( 1 << lg_tile_b ) - 1 ) >> lg_tile_b, ncl );
    auto all_to_all = [&]( const double2 * from, double2 * to, int q )
    {
        comm_group_begin();
        for( int r = 0; r < world; ++r )
        {
            const std::size_t off = std::size_t( r ) * dc.block_stride + std::size_t( q ) * comp_elems;
            comm_send( reinterpret_cast<const double *>( from + off ), 2 * comp_elems, r, cs );
            comm_recv( reinterpret_cast<double *>( to + off ), 2 * comp_elems, r, cs );
        }
        comm_group_end();
    };
    for( int q = 0; q < nq; ++q )
    {
        run_fwd_a16( plan, spins, q, 1, stream );
        PassArgs pb{};
        pb.in    = plan.A + std::size_t( q ) * ncl * d.Nb * d.Ha;
        pb.out   = plan.B + std::size_t( q ) * comp_elems;
        pb.in_os = std::size_t( d.Nb ) * d.Ha, pb.in_js = pb.out_js = d.Ha;
        pb.out_os = std::size_t( kbl ) * d.Ha, pb.out_split = kbl, pb.out_split_stride = dc.block_stride;
        pb.n_u = d.Ha, pb.n_o = ncl, pb.n_in = d.Nb, pb.n_out = d.Pb, pb.ncol = 1 << plan.fast_b.lg, pb.scale = 1.0;
        launch_pass16<false>( plan.fast_b, grid_b, stream, plan.plan[1], pb, 31, plan.lg_split );
        SB_CUDA_CHECK( cudaEventRecord( plan.ev_q[q], stream ) );
        SB_CUDA_CHECK( cudaStreamWaitEvent( cs, plan.ev_q[q], 0 ) );
        all_to_all( plan.B, plan.C, q );
    }
    SB_CUDA_CHECK( cudaEventRecord( plan.ev_all, cs ) );
    SB_CUDA_CHECK( cudaStreamWaitEvent( stream, plan.ev_all, 0 ) );
    launch_c_mult( plan, plan.C, stream );
    SB_CUDA_CHECK( cudaEventRecord( plan.ev_all, stream ) );
    SB_CUDA_CHECK( cudaStreamWaitEvent( cs, plan.ev_all, 0 ) );
    const double inv_P = 1.0 / ( double( d.Pa ) * d.Pb * d.Pc );
    for( int q = 0; q < nq; ++q )
    {
        all_to_all( plan.C, plan.B, q );
        SB_CUDA_CHECK( cudaEventRecord( plan.ev_q[q], cs ) );
    }
    for( int q = 0; q < nq; ++q )
    {
        SB_CUDA_CHECK( cudaStreamWaitEvent( stream, plan.ev_q[q], 0 ) );
        PassArgs ib{};
        ib.in     = plan.B + std::size_t( q ) * comp_elems;
        ib.out    = plan.A + std::size_t( q ) * ncl * d.Nb * d.Ha;
        ib.out_os = std::size_t( d.Nb ) * d.Ha, ib.in_js = ib.out_js = d.Ha;
        ib.in_os = std::size_t( kbl ) * d.Ha, ib.in_split = kbl, ib.in_split_stride = dc.block_stride;
        ib.n_u = d.Ha, ib.n_o = ncl, ib.n_in = d.Pb, ib.n_out = d.Nb, ib.ncol = 1 << plan.fast_b.lg, ib.scale = 1.0;
        launch_pass16<true>( plan.fast_b, grid_b, stream, plan.plan[1], ib, plan.lg_split, 31 );
        run_inv_a16( plan, g_ddi, inv_P, q, 1, stream );
    }
    SB_CUDA_CHECK( cudaGetLastError() );
    return 4 * nq + 1;
}
} // namespace

// One DDI gradient evaluation: spins -> g_ddi field. Returns the number of kernels launched.
int ddi_gradient( DDIPlan & plan, ConstField3 spins, Field3 g_ddi, cudaStream_t stream )
{
    if( plan.pencil )
        return ddi_gradient_pencil( plan, spins, g_ddi, stream );
    if( plan.peer )
        return ddi_gradient_peer( plan, spins, g_ddi, stream );
    if( plan.comm_stream )
        return ddi_gradient_pipelined( plan, spins, g_ddi, stream );
    const DDIDims & d  = plan.dims;
    const DDIDims & dc = plan.dims_c;
    const int world    = plan.world;
    const int nq       = 3 * d.NB;
    const int kbl      = dc.Pb;
    const int rows     = d.Nb * d.Nc;
    if( plan.fast_a.on )
        run_fwd_a16( plan, spins, 0, nq, stream );
    else
        k_ddi_fwd_a<<<dim3( rows, d.NB ), fft_threads( d.Pa / 4 ), plan.smem_a, stream>>>( plan.plan[0], d, spins, plan.A );
    // forward b: A[q][c][b][ka] -> B; outer index o = q * ncl + c. Single device: B[o][kb][ka]; distributed: the kb axis
    // is cut into per-rank blocks, B[r][o][kb % kbl][ka], so that block r is what rank r needs
    PassArgs pb{};
    pb.in = plan.A, pb.out = plan.B;
    pb.in_os = std::size_t( d.Nb ) * d.Ha, pb.in_js = pb.out_js = d.Ha;
    pb.out_os = std::size_t( kbl ) * d.Ha * ( world > 1 ? 1 : world );
    if( world > 1 )
    {
        pb.out_split        = kbl;
        pb.out_split_stride = dc.block_stride;
    }
    else
        pb.out_os = std::size_t( d.Pb ) * d.Ha;
    pb.n_u = d.Ha, pb.n_o = nq * d.Nc, pb.n_in = d.Nb, pb.n_out = d.Pb, pb.ncol = plan.ncol_b, pb.scale = 1.0;
    const dim3 grid_b( ( d.Ha + plan.ncol_b - 1 ) / plan.ncol_b, pb.n_o );
    const int lg_tile_b = plan.fast_b.lg + plan.fast_b.lg_seq;
    const dim3 grid_b16( ( d.Ha + ( 1 << lg_tile_b ) - 1 ) >> lg_tile_b, pb.n_o );
    if( plan.fast_b.on )
        launch_pass16<false>( plan.fast_b, grid_b16, stream, plan.plan[1], pb, 31, plan.lg_split );
    else
        k_fft_pass<false><<<grid_b, fft_threads( d.Pb / 4 * plan.ncol_b ), plan.smem_b, stream>>>( plan.plan[1], pb );

    double2 * operand = plan.B;
    const std::size_t block_doubles = 2 * dc.block_stride; // one per-rank block, in doubles
    if( world > 1 )
    {
        // all-to-all: my block r -> rank r's block (my rank): "my planes, kb range of r" becomes "planes of r, my kb range"
        comm_group_begin();
        for( int r = 0; r < world; ++r )
        {
            comm_send( reinterpret_cast<const double *>( plan.B + std::size_t( r ) * dc.block_stride ), block_doubles, r, stream );
            comm_recv( reinterpret_cast<double *>( plan.C + std::size_t( r ) * dc.block_stride ), block_doubles, r, stream );
        }
        comm_group_end();
        operand = plan.C;
    }

    launch_c_mult( plan, operand, stream );

    if( world > 1 )
    {
        comm_group_begin();
        for( int r = 0; r < world; ++r )
        {
            comm_send( reinterpret_cast<const double *>( plan.C + std::size_t( r ) * dc.block_stride ), block_doubles, r, stream );
            comm_recv( reinterpret_cast<double *>( plan.B + std::size_t( r ) * dc.block_stride ), block_doubles, r, stream );
        }
        comm_group_end();
    }

    // inverse b: B -> A, keep b < Nb
    PassArgs ib{};
    ib.in = plan.B, ib.out = plan.A;
    ib.out_os = std::size_t( d.Nb ) * d.Ha, ib.in_js = ib.out_js = d.Ha;
    if( world > 1 )
    {
        ib.in_os           = std::size_t( kbl ) * d.Ha;
        ib.in_split        = kbl;
        ib.in_split_stride = dc.block_stride;
    }
    else
        ib.in_os = std::size_t( d.Pb ) * d.Ha;
    ib.n_u = d.Ha, ib.n_o = nq * d.Nc, ib.n_in = d.Pb, ib.n_out = d.Nb, ib.ncol = plan.ncol_b, ib.scale = 1.0;
    if( plan.fast_b.on )
        launch_pass16<true>( plan.fast_b, grid_b16, stream, plan.plan[1], ib, plan.lg_split, 31 );
    else
        k_fft_pass<true><<<grid_b, fft_threads( d.Pb / 4 * plan.ncol_b ), plan.smem_b, stream>>>( plan.plan[1], ib );
    const double inv_P = 1.0 / ( double( d.Pa ) * d.Pb * d.Pc );
    if( plan.fast_a.on )
        run_inv_a16( plan, g_ddi, inv_P, 0, nq, stream );
    else
        k_ddi_inv_a<<<dim3( rows, d.NB ), fft_threads( d.Pa / 4 ), plan.smem_a, stream>>>( plan.plan[0], d, plan.A, g_ddi, inv_P );
    SB_CUDA_CHECK( cudaGetLastError() );
    return 5;
}

} // namespace dev
} // namespace sb
