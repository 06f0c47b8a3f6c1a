// staged-input receive kernels for int16_t samples (see gf3_rx_staged.inc)
#define GF3_STAGED_T int16_t
#define GF3_STAGED_NAME i16
#define GF3_STAGED_ESTIMATE 1
#include "gf3_rx_staged.inc"
