// TEST INFRASTRUCTURE ONLY: second translation unit of libdsm_refmap.so -- the reference's fusion_functions.cpp,
// compiled where it lies, threads inlined (deterministic) like libdsm_ref_serial.so
#include "ref_common.hpp"
#include DSM_REF_SOURCE
