// Shared helpers of the C API layer. Every API function body is a function-try-block that ends in
// SB_API_CATCH: no exception crosses the C ABI, failures are logged and a neutral value is returned
// (the reference's convention, core/include/utility/Exception.hpp:119-121, core/src/Spirit/System.cpp:9-46).
#pragma once

#include "../core/config.hpp"
#include "../core/configurations.hpp"
#include "../core/constants.hpp"
#include "../core/logging.hpp"
#include "../core/method.hpp"
#include "../core/state.hpp"

#define SB_API_CATCH_VOID                                                                                              \
    catch( ... )                                                                                                       \
    {                                                                                                                  \
        sb::handle_exception_api( __func__, idx_image, idx_chain );                                                    \
    }
#define SB_API_CATCH_RET( value )                                                                                      \
    catch( ... )                                                                                                       \
    {                                                                                                                  \
        sb::handle_exception_api( __func__, idx_image, idx_chain );                                                    \
        return value;                                                                                                  \
    }

namespace sb
{
struct ImageRef
{
    std::shared_ptr<Spin_System> image;
    std::shared_ptr<Chain> chain;
};
inline ImageRef resolve( State * state, int & idx_image, int & idx_chain )
{
    ImageRef r;
    from_indices( state, idx_image, idx_chain, r.image, r.chain );
    return r;
}
// RAII lock of an image (setters lock the image, core/src/Spirit/Hamiltonian.cpp:64)
struct ImageLock
{
    explicit ImageLock( Spin_System & s ) : s_( s )
    {
        s_.Lock();
    }
    ~ImageLock()
    {
        s_.Unlock();
    }
    Spin_System & s_;
};
// Pinned sites are reset to their orientation after every change of a configuration through the API
// (Geometry::Apply_Pinning at the end of each Configuration_* / Transition_* call, core/src/Spirit/Configurations.cpp:151-637).
// Declared after the ImageLock of a call: runs before the lock is released.
struct PinGuard
{
    explicit PinGuard( Spin_System & s ) : s_( s ) {}
    ~PinGuard()
    {
        s_.geometry->apply_pinning( s_.spins.data() );
    }
    Spin_System & s_;
};
} // namespace sb
