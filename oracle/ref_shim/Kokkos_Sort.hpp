// TEST INFRASTRUCTURE ONLY -- nothing from Kokkos_Sort is used by the kernel headers.
#pragma once
#include "Kokkos_Core.hpp"
