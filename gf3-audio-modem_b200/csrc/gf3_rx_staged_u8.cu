// staged-input receive kernels for uint8_t samples (see gf3_rx_staged.inc)
#define GF3_STAGED_T uint8_t
#define GF3_STAGED_NAME u8
#define GF3_STAGED_ESTIMATE 1
#include "gf3_rx_staged.inc"
