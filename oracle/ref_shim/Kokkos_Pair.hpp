// TEST INFRASTRUCTURE ONLY -- Kokkos::pair is std::pair in the shim.
#pragma once
#include "Kokkos_Core.hpp"
