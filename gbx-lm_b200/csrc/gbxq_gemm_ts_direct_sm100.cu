// gbxq_gemm_ts_direct_sm100.cu -- the TMEM-operand GEMM for matrices whose row of scales is not a legal TMA row pitch
// (K / group_size not a multiple of 8, e.g. the tensor-parallel K shards of Qwen2.5-32B: o_proj K/4 = 1280 = 10 groups,
// down_proj K/2 = 13824 = 108 groups of 128): the same kernel source with the scales read from global memory by the
// dequant threads.  See gbxq_gemm_ts_sm100.cu.
#define GBXQ_TS_DIRECT 1
#include "gbxq_gemm_ts_sm100.cu"
