// gbxq_gemm_ts_grouped_sm100.cu -- the TMEM-operand GEMM for up to three projections that read the same activations
// (q_proj | k_proj | v_proj, gate_proj | up_proj: gbx_lm/models/qllama.py:76,115) and share bit width and group size, as
// ONE launch: the same kernel source as gbxq_gemm_ts_sm100.cu with the tiles of all segments in one grid.  At decode
// batches a launch costs ~5 us of prologue / epilogue whatever its size, and k_proj / v_proj alone fill 8 of 148 SMs.
#define GBXQ_TS_GROUPED 1
#include "gbxq_gemm_ts_sm100.cu"
