// TEST INFRASTRUCTURE (oracle): the reference's pgens/streaming problem generator + a state dump hook.
#ifndef PROBLEM_GENERATOR_WRAP_H
#define PROBLEM_GENERATOR_WRAP_H
#define EB_REF_PGEN "/root/reference/pgens/streaming/pgen.hpp"
#include "../dump_common.hpp"
#endif
