// staged-input receive kernels for float samples (see gf3_rx_staged.inc)
#define GF3_STAGED_T float
#define GF3_STAGED_NAME f32
#include "gf3_rx_staged.inc"
