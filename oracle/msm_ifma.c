/* ORACLE (test infrastructure only).  AVX-512 IFMA 4-way vector backend of the variable-time MSM: see vec4_ifma.h / vec4_msm.h. */
#include "ge.h"
#include "vec4_ifma.h"
#include "vec4_msm.h"
