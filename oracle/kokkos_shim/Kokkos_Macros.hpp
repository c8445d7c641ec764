// Kokkos_Macros.hpp (shim) -- preprocessor configuration only: the reference includes this header
// in the middle of function bodies (src/modules_force.h:40 is spliced into Input and ExaMiniMD::init).
#ifndef KOKKOS_SHIM_MACROS_HPP
#define KOKKOS_SHIM_MACROS_HPP
#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FORCEINLINE_FUNCTION inline __attribute__((always_inline))
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]
#define KOKKOS_VERSION 30000
#ifdef KOKKOS_SHIM_OPENMP
#define KOKKOS_ENABLE_OPENMP
#else
#define KOKKOS_ENABLE_SERIAL
#endif
#endif
