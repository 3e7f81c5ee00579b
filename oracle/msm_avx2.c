/* ORACLE (test infrastructure only).  AVX2 4-way vector backend of the variable-time MSM: see vec4_avx2.h / vec4_msm.h. */
#include "ge.h"
#include "vec4_avx2.h"
#include "vec4_msm.h"
